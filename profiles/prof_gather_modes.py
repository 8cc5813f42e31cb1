#!/usr/bin/env python
"""Numeric-phase time (k_gather_tile) of config 4 (256^3 H8 diffusion) and config 2 (128^3 H8 elasticity) for the pipelining mode
selected by FEGPU_GATHER_MODE (one process per mode: the knob is read once).  Prints one JSON line."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finetools_jl_b200 as fe

def run(n, ndn):
    fens, fes = fe.H8block(1.0, 1.0, 1.0, n, n, n)
    u = fe.NodalField(np.zeros((fens.count(), ndn))); fe.numberdofs(u)
    femm = fe.FEMMBase(fe.IntegDomain(fes, fe.GaussRule(3, 2)))
    geom = fe.NodalField(fens.xyz)
    a = fe.SysmatAssemblerSparseGPU(0.0); a.setnomatrixresult(True)
    if ndn == 1:
        kappa = np.array([[1.5, 0.2, 0.1], [0.2, 2.5, 0.3], [0.1, 0.3, 3.5]])
        call = lambda: fe.bilform_diffusion(femm, a, geom, u, fe.DataCache(kappa), raw=True)
    else:
        lam, mu = 0.3 / (1.3 * 0.4), 1 / 2.6
        C = np.zeros((6, 6)); C[:3, :3] = lam; C[np.arange(3), np.arange(3)] += 2 * mu; C[3:, 3:] = mu * np.eye(3)
        call = lambda: fe.bilform_lin_elastic(femm, a, geom, u, fe.DeforModelRed3D, fe.DataCache(C), raw=True)
    a.ctx.set_overlap(False)
    call()
    t = []
    for _ in range(6):
        call(); t.append(a.timings())
    ctx = a.ctx; a = None; ctx.release_meshes(); ctx.release_cache()
    med = lambda k: float(np.median([x[k] for x in t]))
    return {"numeric_ms": med("numeric_ms"), "integrate_ms": med("integrate_ms"), "total_cached_ms": med("total_ms")}

print(json.dumps({"mode": os.environ.get("FEGPU_GATHER_MODE", "default"), "c4": run(256, 1), "c2": run(128, 3)}), flush=True)
