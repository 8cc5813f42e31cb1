#!/usr/bin/env python
"""Secondary benchmark (not the driver's contract): device-resident timings of BASELINE.json configs 1-5 on ONE B200.
Prints one JSON object per config: elements/s fresh and cached, phase times, nnz/s of CSC construction.
usage: python profiles/bench_configs.py [c1 c2 c3 c4 c5]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finetools_jl_b200 as fe  # noqa: E402
from finetools_jl_b200 import _lib  # noqa: E402

KAPPA3 = np.array([[1.5, 0.2, 0.1], [0.2, 2.5, 0.3], [0.1, 0.3, 3.5]])


def iso():
    lam, mu = 0.3 / (1.3 * 0.4), 1 / 2.6
    C = np.zeros((6, 6)); C[:3, :3] = lam; C[np.arange(3), np.arange(3)] += 2 * mu; C[3:, 3:] = mu * np.eye(3)
    return C


def run(name, fens, fes, ndn, rule, form, coef, m=3, reps=5):
    u = fe.NodalField(np.zeros((fens.count(), ndn))); fe.numberdofs(u)
    a = fe.SysmatAssemblerSparseGPU(0.0)
    femm = fe.FEMMBase(fe.IntegDomain(fes, rule))
    geom = fe.NodalField(fens.xyz)
    a.setnomatrixresult(True)  # keep the CSC on the device: no D2H in this benchmark
    call = {"diffusion": lambda: fe.bilform_diffusion(femm, a, geom, u, fe.DataCache(coef), raw=True),
            "elastic": lambda: fe.bilform_lin_elastic(femm, a, geom, u, fe.DeforModelRed3D, fe.DataCache(coef), raw=True),
            "dot": lambda: fe.bilform_dot(femm, a, geom, u, fe.DataCache(coef), m=m, raw=True)}[form]
    call()
    fresh, fresh_ov, cached = [], [], []
    for _ in range(reps):  # as shipped: the symbolic phase on the second stream, concurrent with the integration
        a.invalidate_patterns(); call(); fresh_ov.append(a.timings())
    a.ctx.set_overlap(False)  # strictly serial phases for the per-phase columns
    for _ in range(reps):
        a.invalidate_patterns(); call(); fresh.append(a.timings())
    a.ctx.set_overlap(True)
    for _ in range(reps):
        call(); cached.append(a.timings())
    med = lambda L, k: float(np.median([t[k] for t in L]))
    _, _, nnz = a.sizes()
    nel = fes.count()
    out = {"config": name, "elements": nel, "nnz": nnz, "triplets": nel * (fes.nne * ndn) ** 2,
           "fresh_ms": med(fresh_ov, "total_ms"), "fresh_elements_per_s": nel / (med(fresh_ov, "total_ms") * 1e-3),
           "fresh_serial_ms": med(fresh, "total_ms"),
           "integrate_ms": med(fresh, "integrate_ms"), "symbolic_ms": med(fresh, "symbolic_ms"), "numeric_ms": med(fresh, "numeric_ms"),
           "cached_ms": med(cached, "total_ms"), "cached_elements_per_s": nel / (med(cached, "total_ms") * 1e-3),
           "csc_nnz_per_s_fresh": nnz / ((med(fresh, "symbolic_ms") + med(fresh, "numeric_ms")) * 1e-3),
           "csc_nnz_per_s_cached": nnz / (med(cached, "numeric_ms") * 1e-3)}
    print(json.dumps(out), flush=True)
    ctx = a.ctx
    del a
    ctx.release_meshes()
    ctx.release_cache()


def main():
    which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
    g32 = fe.GaussRule(3, 2)
    if "c1" in which:
        run("C1 H8 20^3 diffusion (kappa 3x3)", *fe.H8block(12.0, 1.1, 0.32, 20, 20, 20), 1, g32, "diffusion", KAPPA3)
    if "c2" in which:
        run("C2 H8 128^3 lin_elastic", *fe.H8block(1.0, 1.0, 1.0, 128, 128, 128), 3, g32, "elastic", iso())
    if "c3" in which:
        t = time.time()
        n = 100
        f4, s4 = fe.T4block(1.0, 1.0, 1.0, n, n, n)
        h, x = 1.0 / n, f4.xyz
        x0 = x.copy()
        x[:, 0] += 0.2 * h * np.sin(3 * np.pi * x0[:, 1]) * np.cos(2 * np.pi * x0[:, 2])
        x[:, 1] += 0.2 * h * np.sin(3 * np.pi * x0[:, 2]) * np.cos(2 * np.pi * x0[:, 0])
        x[:, 2] += 0.2 * h * np.sin(3 * np.pi * x0[:, 0]) * np.cos(2 * np.pi * x0[:, 1])
        fens, fes = fe.T4toT10(f4, s4)
        sys.stderr.write("T10 mesh generated in %.1f s\n" % (time.time() - t))
        run("C3 T10 distorted 6x100^3 mass (TetRule 4)", fens, fes, 1, fe.TetRule(4), "dot", np.eye(1))
    if "c4" in which:
        run("C4 H8 256^3 diffusion (kappa 3x3), 1 GPU", *fe.H8block(1.0, 1.0, 1.0, 256, 256, 256), 1, g32, "diffusion", KAPPA3)
    if "c5" in which:
        n = 96
        fens, vol = fe.H8block(1.0, 1.0, 1.0, n, n, n)
        run("C5a Q4 skin of 96^3 mass (Gauss 2x2, m=2)", fens, fe.meshboundary(vol), 1, fe.GaussRule(2, 2), "dot", np.eye(1), m=2)
        f4, v4 = fe.T4block(1.0, 1.0, 1.0, n, n, n)
        run("C5b T3 skin of 96^3 mass (TriRule 3, m=2)", f4, fe.meshboundary(v4), 1, fe.TriRule(3), "dot", np.eye(1), m=2)
        t = time.time()
        f20, s20 = fe.H8toH20(fens, vol)
        sys.stderr.write("H20 mesh generated in %.1f s\n" % (time.time() - t))
        run("C5c H20 96^3 lin_elastic (Gauss 3x3x3)", f20, s20, 3, fe.GaussRule(3, 3), "elastic", iso(), reps=3)


if __name__ == "__main__":
    main()
