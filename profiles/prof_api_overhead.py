#!/usr/bin/env python
"""Host-side cost of one public-API step of config 2 without the result transport: where do the milliseconds between the
device time (CUDA events) and the wall time go?  usage: python profiles/prof_api_overhead.py"""
import cProfile, io, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finetools_jl_b200 as fe
n = 128
fens, fes = fe.H8block(1.0, 1.0, 1.0, n, n, n)
u = fe.NodalField(np.zeros((fens.count(), 3))); fe.numberdofs(u)
lam, mu = 0.3 / (1.3 * 0.4), 1 / 2.6
C = np.zeros((6, 6)); C[:3, :3] = lam; C[np.arange(3), np.arange(3)] += 2 * mu; C[3:, 3:] = mu * np.eye(3)
a = fe.SysmatAssemblerSparseGPU(0.0)
femm = fe.FEMMBase(fe.IntegDomain(fes, fe.GaussRule(3, 2)))
geom = fe.NodalField(fens.xyz)
pinned_xyz = torch.empty(geom.values.shape[::-1], dtype=torch.float64, pin_memory=True).numpy().T
pinned_xyz[:] = geom.values
geom.values = pinned_xyz
cache = fe.DataCache(C)
a.setnomatrixresult(True)
def step():
    a.invalidate_patterns()
    fe.bilform_lin_elastic(femm, a, geom, u, fe.DeforModelRed3D, cache, raw=True)
for _ in range(3):
    step()
ts = []
for _ in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter(); step(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print("wall ms per step", sorted(ts), "device", a.timings())
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    step()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18); print(s.getvalue()[:3500])
