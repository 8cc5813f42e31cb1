#!/usr/bin/env python
"""Per-rank cost of a P-way row-block partition, measured on ONE B200: the process plays rank r of P (slab owner map,
halo elements recomputed) on the full mesh, so the numbers are what that rank would see in a P-GPU run (assembly has no
collective).  Device-resident, fresh (pattern rebuilt) and cached steps, phase times with the overlap switched off.
usage: python profiles/emulate_rank.py <c2w|c4s> P r0 [r1 ...]
  c2w: weak scaling of the bench workload, H8 128x128x(128 P) lin_elastic
  c4s: strong scaling of BASELINE config 4, H8 256^3 diffusion split P ways"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finetools_jl_b200 as fe  # noqa: E402

KAPPA3 = np.array([[1.5, 0.2, 0.1], [0.2, 2.5, 0.3], [0.1, 0.3, 3.5]])


def iso():
    lam, mu = 0.3 / (1.3 * 0.4), 1 / 2.6
    C = np.zeros((6, 6)); C[:3, :3] = lam; C[np.arange(3), np.arange(3)] += 2 * mu; C[3:, 3:] = mu * np.eye(3)
    return C


def main():
    cfg, P = sys.argv[1], int(sys.argv[2])
    ranks = [int(x) for x in sys.argv[3:]] or [0]
    reps = int(os.environ.get("REPS", "5"))
    g32 = fe.GaussRule(3, 2)
    if cfg == "c2w":
        fens, fes = fe.H8block(1.0, 1.0, float(P), 128, 128, 128 * P)
        ndn, form, coef = 3, "elastic", iso()
    else:
        fens, fes = fe.H8block(1.0, 1.0, 1.0, 256, 256, 256)
        ndn, form, coef = 1, "diffusion", KAPPA3
    u = fe.NodalField(np.zeros((fens.count(), ndn))); fe.numberdofs(u)
    femm = fe.FEMMBase(fe.IntegDomain(fes, g32))
    geom = fe.NodalField(fens.xyz)
    owner = fe.slab_owner(fens.count(), P) if P > 1 else None
    for r in ranks:
        a = fe.SysmatAssemblerSparseGPU(0.0)
        a.setnomatrixresult(True)
        if form == "elastic":
            call = lambda: fe.bilform_lin_elastic(femm, a, geom, u, fe.DeforModelRed3D, fe.DataCache(coef), raw=True, node_owner=owner, my_rank=r)
        else:
            call = lambda: fe.bilform_diffusion(femm, a, geom, u, fe.DataCache(coef), raw=True, node_owner=owner, my_rank=r)
        call()
        ov, ser, cached = [], [], []
        for _ in range(reps):
            a.invalidate_patterns(); call(); ov.append(a.timings())
        a.ctx.set_overlap(False)
        for _ in range(reps):
            a.invalidate_patterns(); call(); ser.append(a.timings())
        a.ctx.set_overlap(True)
        for _ in range(reps):
            call(); cached.append(a.timings())
        med = lambda L, k: float(np.median([t[k] for t in L]))
        _, _, nnz = a.sizes()
        print(json.dumps({"config": cfg, "P": P, "rank": r, "elements_global": fes.count(), "nnz_rank": nnz,
                          "fresh_ms": med(ov, "total_ms"), "fresh_serial_ms": med(ser, "total_ms"),
                          "integrate_ms": med(ser, "integrate_ms"), "symbolic_ms": med(ser, "symbolic_ms"),
                          "numeric_ms": med(ser, "numeric_ms"), "cached_ms": med(cached, "total_ms")}), flush=True)
        ctx = a.ctx
        del a
        ctx.release_meshes()
        ctx.release_cache()


if __name__ == "__main__":
    main()
