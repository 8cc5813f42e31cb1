#!/usr/bin/env python
"""Where does the end-to-end time of config 2 go?  Times the result transport (fegpu_makematrix_copy) into pinned and
pageable host arrays, against a plain pinned D2H copy of the same number of bytes, and one full public-API step.
usage: python profiles/prof_e2e.py [edge=128]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finetools_jl_b200 as fe  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
fens, fes = fe.H8block(1.0, 1.0, 1.0, n, n, n)
u = fe.NodalField(np.zeros((fens.count(), 3))); fe.numberdofs(u)
lam, mu = 0.3 / (1.3 * 0.4), 1 / 2.6
C = np.zeros((6, 6)); C[:3, :3] = lam; C[np.arange(3), np.arange(3)] += 2 * mu; C[3:, 3:] = mu * np.eye(3)
a = fe.SysmatAssemblerSparseGPU(0.0)
femm = fe.FEMMBase(fe.IntegDomain(fes, fe.GaussRule(3, 2)))
geom = fe.NodalField(fens.xyz)
pinned_xyz = torch.empty(geom.values.shape[::-1], dtype=torch.float64, pin_memory=True).numpy().T  # column-major, page-locked
pinned_xyz[:] = geom.values
geom.values = pinned_xyz
cache = fe.DataCache(C)
a.setnomatrixresult(True)
fe.bilform_lin_elastic(femm, a, geom, u, fe.DeforModelRed3D, cache, raw=True)
a.setnomatrixresult(False)
m_, n_, nnz = a.sizes()
pin = lambda cnt, dt: torch.empty(cnt, dtype=dt, pin_memory=True).numpy()
out_pin = (pin(n_ + 1, torch.int64), pin(nnz, torch.int64), pin(nnz, torch.float64))
out_page = (np.empty(n_ + 1, np.int64), np.empty(nnz, np.int64), np.empty(nnz, np.float64))
for o in out_page:
    o[:] = 0  # fault the pages in
res = {"edge": n, "nnz": nnz, "csc_bytes": (n_ + 1) * 8 + nnz * 16, "host_cores": os.cpu_count()}


def t(fn, reps=3):
    fn()
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


res["copy_pinned_ms"] = t(lambda: a._fetch(True, out_pin))
res["copy_pageable_ms"] = t(lambda: a._fetch(True, out_page))
res["copy_values_only_pinned_ms"] = t(lambda: a.fetch_values(out_pin[2]))
# parity of the transport itself: pinned and pageable results identical
assert np.array_equal(out_pin[1], out_page[1]) and np.array_equal(out_pin[2], out_page[2]) and np.array_equal(out_pin[0], out_page[0])
assert out_pin[1].min() >= 1 and out_pin[1].max() <= m_
# plain D2H of the same bytes for reference (what the old path cost), and of the reduced bytes
d = torch.empty(nnz * 2, dtype=torch.int64, device="cuda")
h = torch.empty(nnz * 2, dtype=torch.int64, pin_memory=True)
res["plain_d2h_16B_per_nnz_ms"] = t(lambda: h.copy_(d, non_blocking=False))
res["plain_d2h_GBps"] = nnz * 16 / (res["plain_d2h_16B_per_nnz_ms"] * 1e-3) / 1e9
res["full_api_step_ms"] = t(lambda: (a.invalidate_patterns(), fe.bilform_lin_elastic(femm, a, geom, u, fe.DeforModelRed3D, cache, raw=True, out=out_pin)))
res["api_step_no_transfer_ms"] = t(lambda: (a.setnomatrixresult(True), a.invalidate_patterns(),
                                            fe.bilform_lin_elastic(femm, a, geom, u, fe.DeforModelRed3D, cache, raw=True), a.setnomatrixresult(False)))
res["transfer_stats"] = a.ctx.transfer_stats()
print(json.dumps(res))
