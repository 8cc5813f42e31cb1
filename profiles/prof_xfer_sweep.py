#!/usr/bin/env python
"""Sweep of the result-transport knobs (host threads, SIMD width, chunk size, narrowing on/off) on config 2's CSC.
Each setting runs in a fresh process because the transport state is built once per context.
usage: python profiles/prof_xfer_sweep.py            (driver)   |   python profiles/prof_xfer_sweep.py child"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import torch
    import finetools_jl_b200 as fe
    n = 128
    fens, fes = fe.H8block(1.0, 1.0, 1.0, n, n, n)
    u = fe.NodalField(np.zeros((fens.count(), 3))); fe.numberdofs(u)
    lam, mu = 0.3 / (1.3 * 0.4), 1 / 2.6
    C = np.zeros((6, 6)); C[:3, :3] = lam; C[np.arange(3), np.arange(3)] += 2 * mu; C[3:, 3:] = mu * np.eye(3)
    a = fe.SysmatAssemblerSparseGPU(0.0)
    femm = fe.FEMMBase(fe.IntegDomain(fes, fe.GaussRule(3, 2)))
    a.setnomatrixresult(True)
    fe.bilform_lin_elastic(femm, a, fe.NodalField(fens.xyz), u, fe.DeforModelRed3D, fe.DataCache(C), raw=True)
    a.setnomatrixresult(False)
    m_, n_, nnz = a.sizes()
    pin = lambda cnt, dt: torch.empty(cnt, dtype=dt, pin_memory=True).numpy()
    out = (pin(n_ + 1, torch.int64), pin(nnz, torch.int64), pin(nnz, torch.float64))
    if os.environ.get("FEGPU_PAGEABLE") == "1":
        out = (np.zeros(n_ + 1, np.int64), np.zeros(nnz, np.int64), np.zeros(nnz, np.float64))
    a._fetch(True, out)
    ts = []
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter(); a._fetch(True, out); ts.append((time.perf_counter() - t0) * 1e3)
    print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("FEGPU_")}, "copy_ms": sorted(ts),
                      "stats": a.ctx.transfer_stats()}))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "compress":  # the compressed row-index transport against the int32 one
        settings = [{"FEGPU_XFER_COMPRESS": "0"}] + [{"FEGPU_HOST_THREADS": th} for th in ("4", "8", "12", "16")]
        for pg in ("0",):
            settings.append({"FEGPU_XFER_COMPRESS": "0", "FEGPU_PAGEABLE": "1"})
            settings.append({"FEGPU_HOST_THREADS": "8", "FEGPU_PAGEABLE": "1"})
        return sweep(settings)
    settings = [{"FEGPU_XFER_NARROW": "0"}]
    for th in ("4", "8", "16"):
        for simd in ("1", "2"):
            settings.append({"FEGPU_HOST_THREADS": th, "FEGPU_XFER_SIMD": simd})
    for mb in ("4", "8", "16"):
        settings.append({"FEGPU_HOST_THREADS": "8", "FEGPU_XFER_SIMD": "2", "FEGPU_XFER_CHUNK_MB": mb})
    settings.append({"FEGPU_XFER_NARROW": "0"})
    sweep(settings)


def sweep(settings):
    for s in settings:
        env = dict(os.environ); env.update(s)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED " + r.stderr[-300:], flush=True)


if __name__ == "__main__":
    child() if len(sys.argv) > 1 and sys.argv[1] == "child" else main()
