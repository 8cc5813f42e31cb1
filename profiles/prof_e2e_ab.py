#!/usr/bin/env python
"""A/B of the public-API step of config 2 (pinned xyz H2D, fresh pattern, full CSC into pinned host arrays): wall time of
REPS steps (min / median), to compare library knobs inside ONE gpurun call (the boxes of the pool differ by several ms).
usage: [FEGPU_QUEUED_FORMS=0] [FEGPU_EARLY_META=0] python profiles/prof_e2e_ab.py [edge=128] [reps=6]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finetools_jl_b200 as fe  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
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
fe.bilform_lin_elastic(femm, a, geom, u, fe.DeforModelRed3D, cache, raw=True)
a.setnomatrixresult(False)
m_, n_, nnz = a.sizes()
pin = lambda cnt, dt: torch.empty(cnt, dtype=dt, pin_memory=True).numpy()
out = (pin(n_ + 1, torch.int64), pin(nnz, torch.int64), pin(nnz, torch.float64))


def step():
    a.invalidate_patterns()
    fe.bilform_lin_elastic(femm, a, geom, u, fe.DeforModelRed3D, cache, raw=True, out=out)


step()
ts = []
for _ in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter(); step(); torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
tc = []
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); a._fetch(True, out); torch.cuda.synchronize()
    tc.append((time.perf_counter() - t0) * 1e3)
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("FEGPU_")}, "api_step_ms_min": min(ts),
                  "api_step_ms_median": float(np.median(ts)), "copy_only_ms_min": min(tc), "checksum": float(out[2][::1000003].sum())}))
