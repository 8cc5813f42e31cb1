#!/usr/bin/env python
"""bench.py -- elements/s assembled into CSC on B200 (BASELINE.json metric), strong scaling over 1/2/4/8 GPUs.

Headline workload: BASELINE.json configs[3], the north-star target -- bilform_diffusion (kappa 3x3) on a 256^3 H8 block
(16 777 216 elements, 64 triplets each, nnz = 769^3 = 454 756 609), GaussRule(3,2).  With N GPUs the SAME mesh is split into
N node-owned row blocks (z-slabs = contiguous node ranges); every rank integrates the elements that touch one of its nodes
(halo recomputed) and builds the CSC of its rows; there is no collective on the data path (scaling: "strong").

A "step" is one complete fresh assembly of the rank's block: symbolic phase (node -> element adjacency, neighbour lists,
colptr / rowval) -> element integration -> numeric gather-sum into nzval.  The sparsity-pattern cache is INVALIDATED before
every timed step, so no work is skipped; the cached re-assembly rate is reported separately under "cached".

  value  : inputs resident in HBM, CUDA-event time over K steps, max over ranks
  e2e    : the same step through the public API (the reference's call shape) with HOST buffers: coordinates go host -> device
           and colptr / rowval / nzval come back device -> host inside the timed region (pinned result arrays; the pageable
           figure a shim without page-locked buffers gets is reported beside it)
  config2: BASELINE configs[1] (128^3 H8 bilform_lin_elastic) measured the same way, as a secondary block of the line
  --impl reference : the CPU oracle (C restatement of FinEtools.jl's serial path; Julia is not available) on a bounded
           sample of the headline workload
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

_emit = None
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "elements/s assembled into CSC (H8 diffusion stiffness, fresh assembly incl. pattern build)"
UNIT = "elements/s"
KAPPA3 = np.array([[1.5, 0.2, 0.1], [0.2, 2.5, 0.3], [0.1, 0.3, 3.5]])


def isotropic_C(E=1.0, nu=0.3):
    lam = E * nu / ((1 + nu) * (1 - 2 * nu))
    mu = E / (2 * (1 + nu))
    C = np.zeros((6, 6))
    C[:3, :3] = lam
    C[np.arange(3), np.arange(3)] += 2 * mu
    C[3:, 3:] = mu * np.eye(3)
    return C


# Per-element figures of the two workloads.
#   flops_ref   : SURVEY.md 8(a), counted as the reference's loops execute
#   flops_exec  : FP64 operations the integration kernel executes per element (FMA = 2), from its SASS
#   vals        : doubles stored per element (compact upper block triangle of the symmetric element matrix)
# Compulsory bytes of THIS implementation are below SURVEY 8(d)'s 16 B / triplet figures (no keys, int32 indices, upper block
# triangle only), so every fraction is reported against the implementation's own compulsory traffic.
WORKLOADS = {
    "c4": dict(label="BASELINE configs[3]: bilform_diffusion (kappa 3x3), H8 block", edge=256, ndn=1, form="diffusion",
               flops_ref=6816, flops_exec=5600, vals=36, kernel="k_h8_diffusion"),
    "c2": dict(label="BASELINE configs[1]: bilform_lin_elastic (isotropic C), H8 block", edge=128, ndn=3, form="elastic",
               flops_ref=51936, flops_exec=25128, vals=324, kernel="k_h8_elastic",
               # isotropic (cubic-symmetry) C takes the outer-product formulation: 8 x (333 geometry + 4 x 168) + 36 x 32 for the blocks
               flops_exec_iso=9192),
}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_reference_run(n_edge, threads, form="diffusion"):
    """One assembly of an n_edge^3 H8 block with the CPU oracle.  threads == 1 is the reference's own serial path; threads > 1
    shards the ELEMENT LOOP over threads writing disjoint slices of one COO buffer (what FinEtoolsMultithreading does for
    the reference) and keeps the serial sparse()."""
    import finetools_jl_b200 as fe   # host-side mesh generator only
    from oracle import oracle as orc
    fens, fes = fe.H8block(1.0, 1.0, 1.0, n_edge, n_edge, n_edge)
    ndn = 1 if form == "diffusion" else 3
    u = fe.NodalField(np.zeros((fens.count(), ndn)))
    fe.numberdofs(u)
    rule = fe.GaussRule(3, 2)
    coef = KAPPA3 if form == "diffusion" else isotropic_C()
    fn = orc.bilform_diffusion_coo if form == "diffusion" else orc.bilform_lin_elastic_coo
    nall = u.nalldofs()
    nel = fes.count()
    em2 = (8 * ndn) ** 2
    t0 = time.perf_counter()
    if threads <= 1:
        I, J, V = fn("H8", fes.conn, fens.xyz, u.dofnums, nall, rule.param_coords, rule.weights, coef)
    else:
        from concurrent.futures import ThreadPoolExecutor
        nt = nel * em2
        I, J, V = np.empty(nt, np.int64), np.empty(nt, np.int64), np.empty(nt)
        bounds = np.linspace(0, nel, threads + 1).astype(np.int64)

        def work(k):
            lo, hi = int(bounds[k]), int(bounds[k + 1])
            sl = slice(lo * em2, hi * em2)
            fn("H8", fes.conn[lo:hi], fens.xyz, u.dofnums, nall, rule.param_coords, rule.weights, coef, out=(I[sl], J[sl], V[sl]))
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(work, range(threads)))
    t1 = time.perf_counter()
    colptr, rowval, nzval = orc.sparse(I, J, V, nall, nall)
    t2 = time.perf_counter()
    return nel, t2 - t0, t1 - t0, t2 - t1, nzval.size


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    threads = args.ref_threads if args.ref_threads > 0 else (os.cpu_count() or 1)
    n = args.ref_edge
    for _ in range(args.warmup):
        cpu_reference_run(n, threads)
    t = 0.0
    nel = 0
    for _ in range(args.steps):
        ne, dt, _, _, _ = cpu_reference_run(n, threads)
        t += dt
        nel += ne
    val = nel / t
    sample = "%d^3 H8 diffusion block (%d elements) per step; element loop on %d thread(s), serial sparse()" % (n, n ** 3, threads)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS["c4"]["label"] + ", GaussRule(3,2) -- bounded sample of the 256^3 block", "sample_edge": n},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "C restatement of FinEtools.jl v8.2.11 (oracle/fe_oracle.c); Julia is not installed on this image"}
    _emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
class Workload:
    """One BASELINE configuration on this rank: host mesh, fields, the public-API objects and the C-ABI handles."""

    def __init__(self, key, fe, ctx, world, rank, torch):
        from finetools_jl_b200 import _lib
        self.spec, self.fe, self.ctx, self.world, self.rank, self._lib, self.torch = WORKLOADS[key], fe, ctx, world, rank, _lib, torch
        n = self.spec["edge"]
        self.fens, self.fes = fe.H8block(1.0, 1.0, 1.0, n, n, n)
        self.u = fe.NodalField(np.zeros((self.fens.count(), self.spec["ndn"])))
        fe.numberdofs(self.u)
        self.rule = fe.GaussRule(3, 2)
        self.femm = fe.FEMMBase(fe.IntegDomain(self.fes, self.rule))
        self.geom = fe.NodalField(self.fens.xyz)
        # the step's input (node coordinates) lives in page-locked, column-major host memory, as the contract asks
        pinned = torch.empty(self.geom.values.shape[::-1], dtype=torch.float64, pin_memory=True).numpy().T
        pinned[:] = self.geom.values
        self.geom.values = pinned
        self.owner = fe.slab_owner(self.fens.count(), world) if world > 1 else None
        self.coef = KAPPA3 if self.spec["form"] == "diffusion" else isotropic_C()
        self.cache = fe.DataCache(self.coef)
        self.coef_f = np.asfortranarray(self.coef)
        self.a = fe.SysmatAssemblerSparseGPU(0.0, ctx=ctx)
        self.nelem = self.fes.count()
        # first call: uploads, builds everything (also the warm-up of the allocator)
        self.api_call(self.a, out=None, fetch=False)
        self.dmesh = ctx.device_mesh(self.fes)
        self.dof = self.dmesh.dofmap(self.u)
        self.m, self.n, self.nnz = self.a.sizes()
        self.win = self.dmesh.window()  # (lo, hi, nactive)

    def api_call(self, a, out, fetch=True):
        """The call a user makes: bilform_*(femm, assembler, geom, u, DataCache(...)) -> CSC in host arrays."""
        fe = self.fe
        if not fetch:
            a.setnomatrixresult(True)
        try:
            if self.spec["form"] == "diffusion":
                return fe.bilform_diffusion(self.femm, a, self.geom, self.u, self.cache, raw=True, node_owner=self.owner, my_rank=self.rank, out=out)
            return fe.bilform_lin_elastic(self.femm, a, self.geom, self.u, fe.DeforModelRed3D, self.cache, raw=True, node_owner=self.owner,
                                          my_rank=self.rank, out=out)
        finally:
            if not fetch:
                a.setnomatrixresult(False)

    def device_step(self, fresh=True):
        L, _lib = self._lib.lib(), self._lib
        if fresh:
            _lib.check(L.fegpu_pattern_invalidate(self.dof), self.ctx.handle)
        if self.spec["form"] == "diffusion":
            _lib.check(L.fegpu_bilform_diffusion(self.dmesh.handle, self.dof, 1, _lib.fptr(self.coef_f), self.a.handle), self.ctx.handle)
        else:
            _lib.check(L.fegpu_bilform_lin_elastic(self.dmesh.handle, self.dof, _lib.fptr(self.coef_f), self.a.handle), self.ctx.handle)

    def release(self):
        self.a = None
        self.ctx.release_meshes()
        self.ctx.release_cache()


def run_gpu(args):
    import torch
    import finetools_jl_b200 as fe

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = fe.GPUContext(local_rank, stream=torch.cuda.current_stream().cuda_stream)
    hbm_peak, peak_kind = measured_peaks()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if dist is None:
            return float(x)
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum_int(x):
        if dist is None:
            return int(x)
        t = torch.tensor([x], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        return int(t.item())

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        return allmax(ev0.elapsed_time(ev1))

    def wall(fn, steps):
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        barrier()
        return allmax(time.perf_counter() - t0)

    def measure(key, steps, warmup, sampler=None):
        """Everything measured for one workload; returns a dict (identical code path for headline and secondary block)."""
        W = Workload(key, fe, ctx, world, rank, torch)
        spec = W.spec
        ctx.set_async(True)
        for _ in range(warmup):
            W.device_step(True)
        if sampler is not None and rank == 0:
            sampler.start()
        launches0 = ctx.launch_count()
        ms_total = timed(lambda: W.device_step(True), steps)
        launches = ctx.launch_count() - launches0
        # per-kernel device times of ONE fresh step (named CUDA events on the launching stream), phases strictly serial: the
        # timed steps above overlap the element integration (second stream) with the symbolic phase, which would smear the
        # per-kernel durations the roofline figures are computed from
        ctx.set_overlap(False)
        phases, marks = [], []
        for _ in range(max(3, min(steps, 5))):
            ctx.marks_begin()
            W.device_step(True)
            ctx.synchronize()
            marks.append(dict(ctx.marks_read()))
            phases.append(W.a.timings())
        ph = {k: float(np.median([p[k] for p in phases])) for k in phases[0]}
        mk = {k: float(np.median([m.get(k, 0.0) for m in marks])) for k in marks[0]}
        ctx.set_overlap(True)
        # cached re-assembly (pattern reused: integration + gather-sum only)
        for _ in range(2):
            W.device_step(False)
        ms_cached = timed(lambda: W.device_step(False), steps)
        clocks = sampler.stop() if (sampler is not None and rank == 0) else None
        ctx.set_async(False)

        # ---- e2e through the public API with host buffers
        m_, n_, nnz_local = W.a.sizes()
        pin = lambda cnt, dt: torch.empty(max(cnt, 1), dtype=dt, pin_memory=True).numpy()[:cnt]
        out = (pin(n_ + 1, torch.int64), pin(nnz_local, torch.int64), pin(nnz_local, torch.float64))

        def e2e_fresh_pinned():
            W.a.invalidate_patterns()
            W.api_call(W.a, out)

        def e2e_fresh_pageable():  # what a shim that allocates fresh Vectors per makematrix! gets (FinEtoolsGPU.jl without reuse)
            W.a.invalidate_patterns()
            W.api_call(W.a, None)

        def e2e_cached_values():   # re-assembly on the cached pattern: only coordinates in, nzval out
            W.api_call(W.a, None, fetch=False)
            W.a.fetch_values(out[2])

        e2e_fresh_pinned()
        xfer0 = ctx.transfer_stats()
        e2e_steps = max(1, steps)
        e2e_s = wall(e2e_fresh_pinned, e2e_steps) / e2e_steps
        xfer1 = ctx.transfer_stats()
        side = max(1, min(steps, 3))
        e2e_pageable_s = wall(e2e_fresh_pageable, side) / side
        e2e_cached_values()
        e2e_cached_s = wall(e2e_cached_values, side) / side
        h2d = W.dmesh.h2d_bytes_last
        d2h = (n_ + 1) * 8 + nnz_local * 16
        nn = W.fens.count()
        if xfer1["compressed_results"] - xfer0["compressed_results"] >= e2e_steps:
            # rowval did not cross the link: colptr + per-node neighbour lists (int32 per node pair) + their offsets + the
            # int32 dof map + nzval did; the host threads decoded rowval from them (fegpu_transfer.cu)
            nd2 = spec["ndn"] ** 2
            link = (n_ + 1) * 8 + (nnz_local // nd2) * 4 + (nn + 1) * 8 + nn * spec["ndn"] * 4 + nnz_local * 8
            link_note = ("rowval is rebuilt on the host from the device's neighbour lists (int32 per node pair) + dof map by the library's "
                         "host threads while nzval is in flight")
        elif xfer1.get("stenciled_results", 0) - xfer0.get("stenciled_results", 0) >= e2e_steps:
            # rowval did not cross the link: colptr + one column-stencil id per column (uint32) + a dictionary of row-offset lists
            # (a few KB) + nzval did; the host threads rebuilt rowval (fe_col_stencils / expand_stencils)
            ncw = min(n_, (W.win[1] - W.win[0]) * spec["ndn"])   # the rank's non-empty column window: only its colptr / ids are shipped
            link = (ncw + 1) * 8 + ncw * 4 + nnz_local * 8
            link_note = ("rowval is rebuilt on the host from one column-stencil id per column + a dictionary of row-offset lists, built and "
                         "verified on the device, while nzval is in flight")
        else:
            link = d2h - 4 * nnz_local
            link_note = "rowval crosses the link as int32 and is widened by host threads"
        nnz_total = allsum_int(nnz_local)
        nact_total = allsum_int(W.win[2])
        nelem = W.nelem
        # ---- the optional exchange step: the ranks' row blocks gathered into ONE CSC in the HBM of rank 0 (NCCL over NVLink for the
        # counts and the rowval / nzval slabs + the library's plan / interleave kernels); reported beside the headline, not inside it
        gather = None
        if dist is not None and key == "c4":
            W.device_step(False)
            ctx.synchronize()
            fe.gather_row_blocks_device(W.a, dist, dst=0, ordered=True)   # warm-up (NCCL channels, staging buffers)
            reps = 3
            g_s = wall(lambda: fe.gather_row_blocks_device(W.a, dist, dst=0, ordered=True), reps) / reps
            nnz0 = nnz_local if rank == 0 else 0
            nnz0 = allsum_int(nnz0)
            wire = 16 * (nnz_total - nnz0) + 8 * n_ * world
            gather = {"ms": g_s * 1e3, "wire_bytes": int(wire), "wire_GBps_into_rank0": wire / g_s / 1e9,
                      "note": "wall time incl. host synchronisations, max over ranks; NVLink 5 into one GPU: 900 GB/s nominal per direction"}

        ms_step = ms_total / steps
        nact = W.win[2]                      # elements this rank integrates (its share + halo)
        nnodes_rank = W.win[1] - W.win[0]    # node window of the rank
        vals = spec["vals"]
        ndn = spec["ndn"]
        # compulsory bytes per launch of the three device stages on THIS rank; H8: 8 adjacency planes per node
        # (4 B each), 8 neighbour-slot words per node (8 B each: one byte per candidate), int32 degree / neighbour count
        per_node_planes = 8 * 4 + 8 * 8
        b_int = nact * (32 + 8 * vals) + nnodes_rank * 24                # int32 conn + values written + coordinates read once
        b_gather = nact * 8 * vals + nnz_local * 8 + nnodes_rank * (per_node_planes + 4 + 4 + 8 * ndn)   # values read once + nzval + planes + deg, nnbr, colptr
        # k_adj_place / k_adj_table (with the pre-fill of the planes): conn read, one 4-byte entry per (element, node) written over the fill
        b_adj = nact * (32 + 32) + nnodes_rank * 8 * 4
        # k_sym_tile: rowval + colptr written; per node the slot words written, the adjacency column read and re-written (sorted), degree,
        # nnbr, nbrptr; conn rows read (once from DRAM, the other seven visits are L2 hits by design)
        b_symtile = nnz_local * 8 + (n_ + 1) * 8 + nnodes_rank * (8 * 8 + 2 * 8 * 4 + 4 + 4 + 8) + nact * 32
        b_sym = b_adj + b_symtile
        peaks = ctx.measure_peaks()
        k_int_ms, k_gather_ms = mk.get("integrate", ph["integrate_ms"]), mk.get("gather", ph["numeric_ms"])
        sym_kernels = {k[4:]: v for k, v in mk.items() if k.startswith("sym:")}
        adj_ms = mk.get("sym:k_adj_place", mk.get("sym:k_adj_table", 0.0))
        iso_path = "flops_exec_iso" in spec and os.environ.get("FEGPU_ELASTIC_ISO", "1") != "0"  # the bench material is isotropic
        fl_exec = spec["flops_exec_iso"] if iso_path else spec["flops_exec"]
        kern = {
            spec["kernel"]: {"ms": k_int_ms, "bound": "fp64", "algorithmic_GBps": b_int / (k_int_ms * 1e-3) / 1e9,
                             "path": "outer products (cubic-symmetry D)" if iso_path else "general D",
                             "frac_hbm": b_int / (k_int_ms * 1e-3) / 1e9 / hbm_peak,
                             "executed_TFLOPs": fl_exec * nact / (k_int_ms * 1e-3) / 1e12,
                             "reference_count_TFLOPs": spec["flops_ref"] * nact / (k_int_ms * 1e-3) / 1e12,
                             "dfma_peak_TFLOPs_measured_here": peaks["dfma_tflops"],
                             "frac_fp64_executed": fl_exec * nact / (k_int_ms * 1e-3) / 1e12 / peaks["dfma_tflops"],
                             # element-integration roofline of the north star: the slower of flops / FP64 peak and bytes / HBM peak
                             "frac_of_integration_roofline": max(fl_exec * nact / (peaks["dfma_tflops"] * 1e12),
                                                                 b_int / (hbm_peak * 1e9)) / (k_int_ms * 1e-3)},
            "k_gather": {"ms": k_gather_ms, "bound": "hbm", "algorithmic_GBps": b_gather / (k_gather_ms * 1e-3) / 1e9,
                         "frac": b_gather / (k_gather_ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes": b_gather},
            "k_sym_tile": {"ms": mk.get("sym:k_sym_tile", 0.0), "bound": "hbm", "algorithmic_bytes": b_symtile,
                           "algorithmic_GBps": b_symtile / max(mk.get("sym:k_sym_tile", 0.0) * 1e-3, 1e-12) / 1e9,
                           "frac": b_symtile / max(mk.get("sym:k_sym_tile", 0.0) * 1e-3, 1e-12) / 1e9 / hbm_peak,
                           "note": "instruction-issue bound (sorting network in registers), reported against HBM all the same"},
            "k_adj": {"ms": adj_ms, "bound": "hbm", "algorithmic_bytes": b_adj, "algorithmic_GBps": b_adj / max(adj_ms * 1e-3, 1e-12) / 1e9,
                      "frac": b_adj / max(adj_ms * 1e-3, 1e-12) / 1e9 / hbm_peak,
                      "note": "k_adj_place (or k_adj_table) + the pre-fill of the adjacency planes + k_dof_affine: 4-byte scatter, L2-bound"},
            "symbolic": {"ms": ph["symbolic_ms"], "bound": "hbm", "algorithmic_GBps": b_sym / (ph["symbolic_ms"] * 1e-3) / 1e9,
                         "frac": b_sym / (ph["symbolic_ms"] * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes": b_sym, "kernels_ms": sym_kernels},
        }
        res = {
            "workload": "%s %d^3 (%d elements, %d nnz), GaussRule(3,2); %s" % (
                spec["label"], spec["edge"], nelem, nnz_total,
                "single GPU" if world == 1 else "%d node-owned row blocks (z-slabs), halo recomputed: %d element integrations in total" % (world, nact_total)),
            "value": nelem / (ms_step * 1e-3), "ms_per_step": ms_step, "launches": int(launches), "phases_ms": ph, "marks_ms": mk,
            "cached": {"value": nelem / (ms_cached / steps * 1e-3), "unit": UNIT, "ms_per_step": ms_cached / steps,
                       "note": "re-assembly on the cached pattern (integration + gather-sum), reported separately"},
            "nnz_per_s_csc_construction": nnz_total / ((ph["symbolic_ms"] + ph["numeric_ms"]) * 1e-3),
            "e2e": {"value": nelem / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                    "ms_per_step": e2e_s * 1e3,
                    "note": "per rank: coordinates H2D (pinned; a partitioned rank ships its node window) + full CSC (colptr,rowval,nzval) into "
                            "pinned host Int64/Float64 arrays (d2h_bytes_per_step = the bytes of those arrays); " + link_note,
                    "link_d2h_bytes_per_step": int(link),
                    "pageable": {"value": nelem / e2e_pageable_s, "ms_per_step": e2e_pageable_s * 1e3, "steps": side,
                                 "note": "same call with fresh pageable result arrays per step (no buffer reuse)"},
                    "cached_pattern_values_only": {"value": nelem / e2e_cached_s, "ms_per_step": e2e_cached_s * 1e3, "steps": side,
                                                   "note": "re-assembly on the cached pattern: coordinates in, nzval out (colptr/rowval kept by the caller)"},
                    "transfer_stats": xfer1},
            "gather_blocks": gather, "kernels": kern, "clocks": clocks, "rank0": {"active_elements": nact, "node_window": nnodes_rank, "nnz": nnz_local},
            "peaks": {"hbm_gbs": hbm_peak, "hbm_kind": peak_kind, "dfma_tflops_measured_here": peaks["dfma_tflops"], "copy_gbs_measured_here": peaks["copy_gbs"]},
        }
        W.release()
        return res

    def measure_c5(steps):
        """BASELINE configs[4] (mixed surface / volume run, as it states it for 8 GPUs): H20 96^3 bilform_lin_elastic with GaussRule(3,3)
        (the warp-per-element register-tiled kernel) plus the boundary mass bilform_dot on the Q4 / T3 skins of the 96^3 blocks,
        every FESet split into `world` node-owned row blocks.  Device-resident fresh and cached steps, max over ranks."""
        out = {}
        nb = 96
        cases = []
        fens, fes = fe.H20block(1.0, 1.0, 1.0, nb, nb, nb)
        cases.append(("H20 96^3 lin_elastic GaussRule(3,3)", fens, fes, 3, fe.GaussRule(3, 3), "elastic", isotropic_C(), 3))
        vf, vol = fe.H8block(1.0, 1.0, 1.0, nb, nb, nb)
        cases.append(("Q4 skin of H8 96^3, bilform_dot m=2, GaussRule(2,2)", vf, fe.meshboundary(vol), 1, fe.GaussRule(2, 2), "dot", np.array([[1.0]]), 2))
        vf, vol = fe.T4block(1.0, 1.0, 1.0, nb, nb, nb)
        cases.append(("T3 skin of T4 96^3, bilform_dot m=2, TriRule(3)", vf, fe.meshboundary(vol), 1, fe.TriRule(3), "dot", np.array([[1.0]]), 2))
        for name, fens, fes, ndn, rule, form, coef, m in cases:
            u = fe.NodalField(np.zeros((fens.count(), ndn)))
            fe.numberdofs(u)
            femm = fe.FEMMBase(fe.IntegDomain(fes, rule))
            geom = fe.NodalField(fens.xyz)
            # H20 / skin node numbers have no slab structure (mid-edge nodes are numbered after all corner nodes): partition the
            # nodes geometrically, as the reference does (pointpartitioning: recursive inertial bisection, MeshModificationModule.jl:1029)
            owner = None
            if world > 1:
                owner = (fe.pointpartitioning(fens.xyz, world) - 1).astype(np.int32) if (world & (world - 1)) == 0 else fe.slab_owner(fens.count(), world)
            a = fe.SysmatAssemblerSparseGPU(0.0, ctx=ctx)
            a.setnomatrixresult(True)
            cache = fe.DataCache(coef)
            if form == "elastic":
                call = lambda: fe.bilform_lin_elastic(femm, a, geom, u, fe.DeforModelRed3D, cache, raw=True, node_owner=owner, my_rank=rank)
            else:
                call = lambda: fe.bilform_dot(femm, a, geom, u, cache, m=m, raw=True, node_owner=owner, my_rank=rank)
            call()
            dmesh = ctx.device_mesh(fes)
            dof = dmesh.dofmap(u)
            from finetools_jl_b200 import _lib as L
            cf = np.asfortranarray(coef)

            def dev(fresh):
                if fresh:
                    L.check(L.lib().fegpu_pattern_invalidate(dof), ctx.handle)
                if form == "elastic":
                    L.check(L.lib().fegpu_bilform_lin_elastic(dmesh.handle, dof, L.fptr(cf), a.handle), ctx.handle)
                else:
                    L.check(L.lib().fegpu_bilform_dot(dmesh.handle, dof, L.fptr(cf), int(m), 1.0, a.handle), ctx.handle)

            ctx.set_async(True)
            for _ in range(3):
                dev(True)
            ms_f = timed(lambda: dev(True), steps) / steps
            dev(False)
            ms_c = timed(lambda: dev(False), steps) / steps
            ctx.set_async(False)
            ctx.set_overlap(False)
            dev(True)
            ctx.synchronize()
            ph = a.timings()
            ctx.set_overlap(True)
            nnz_rank = a.sizes()[2]
            nnz = allsum_int(nnz_rank)
            out[name] = {"elements": fes.count(), "nnz": nnz, "fresh_ms": ms_f, "cached_ms": ms_c, "fresh_elements_per_s": fes.count() / (ms_f * 1e-3),
                         "cached_elements_per_s": fes.count() / (ms_c * 1e-3), "phases_ms_rank0": ph}
            if form == "elastic":
                # rank 0's kernels against their rooflines: the gather reads the compact element records (nne (nne + 1) / 2 blocks of
                # 9 doubles) once and writes nzval; k_elastic_tiled executes (H20, 27 points) 27 x (385 geometry + 30 tiles x 72 DFMA) FMA
                # + 210 x 32 for the blocks with the outer-product formulation of an isotropic C, 27 x (1465 + 30 x 216) FMA for a general
                # one; the reference's own loop count is 812 430 flop per element
                nact0 = dmesh.window()[2]
                nne = fes.conn.shape[1]
                b_g = nact0 * (nne * (nne + 1) // 2) * 9 * 8 + nnz_rank * 8
                iso_path = os.environ.get("FEGPU_ELASTIC_ISO", "1") != "0"
                fl = (27 * (385 + 30 * 72) * 2 + 210 * 32) if iso_path else 27 * (1465 + 30 * 216) * 2
                pk = ctx.measure_peaks()
                out[name]["rank0_kernels"] = {
                    "k_gather": {"ms": ph["numeric_ms"], "algorithmic_bytes": b_g, "algorithmic_GBps": b_g / (ph["numeric_ms"] * 1e-3) / 1e9,
                                 "frac_hbm": b_g / (ph["numeric_ms"] * 1e-3) / 1e9 / hbm_peak},
                    "k_elastic_tiled": {"ms": ph["integrate_ms"], "path": "outer products (cubic-symmetry D)" if iso_path else "general D",
                                        "executed_TFLOPs": fl * nact0 / (ph["integrate_ms"] * 1e-3) / 1e12,
                                        "frac_fp64_executed": fl * nact0 / (ph["integrate_ms"] * 1e-3) / 1e12 / pk["dfma_tflops"],
                                        "reference_count_TFLOPs": 812430 * nact0 / (ph["integrate_ms"] * 1e-3) / 1e12}}
            a = None
            ctx.release_meshes()
            ctx.release_cache()
        return out

    sampler = ClockSampler(local_rank)
    r4 = measure("c4", args.steps, args.warmup, sampler)
    r2 = None
    if not args.no_secondary:
        r2 = measure("c2", max(3, min(args.steps, 10)), 3)
    r5 = None
    if args.config5 or (world == 8 and not args.no_secondary):
        r5 = measure_c5(max(3, min(args.steps, 5)))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # dominant device stage of the headline step by time
    stages = {k: r4["kernels"][k] for k in ("k_h8_diffusion", "k_gather", "k_sym_tile", "k_adj") if r4["kernels"][k]["ms"] > 0.0}
    if not stages:  # the general path built the pattern (no thread-per-node kernels): the whole symbolic stage stands for its kernels
        stages = {"symbolic": r4["kernels"]["symbolic"]}
    dom_name = max(stages, key=lambda k: stages[k]["ms"])
    dom = stages[dom_name]
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)
    tr = traffic.get("c4", {}).get(dom_name, {}).get("dram_bytes_per_launch") if world == 1 else None
    if dom["bound"] == "hbm":
        roofline = {"bound": "hbm", "achieved": dom["algorithmic_GBps"], "peak": hbm_peak, "unit": "GB/s", "frac": dom["frac"],
                    "traffic": tr, "algorithmic_bytes_per_launch": dom["algorithmic_bytes"]}
    else:  # the FP64 integration kernel: both terms of the north star's roofline
        roofline = {"bound": "fp64", "peak_source": "DFMA micro-benchmark (fegpu_measure_peaks: k_dfma_peak, dependent-free FMA chains on every SM) run "
                                                    "on this GPU in this process: MEASURED_PEAKS.json carries HBM and bf16 tensor peaks only, and this "
                                                    "kernel is FP64 CUDA-core work (8-36 wide contractions: no tensor-core shape)",
                    "achieved": dom["executed_TFLOPs"], "peak": dom["dfma_peak_TFLOPs_measured_here"], "unit": "TFLOP/s",
                    "frac": dom["frac_fp64_executed"], "traffic": tr, "hbm_GBps": dom["algorithmic_GBps"]}
    roofline.update({"kernel": dom_name, "peak_kind": r4["peaks"]["hbm_kind"], "traffic_source": traffic.get("source") if tr else None,
                     "kernels": r4["kernels"], "copy_gbs_measured_here": r4["peaks"]["copy_gbs_measured_here"],
                     "note": "dominant KERNEL of the fresh step by device time (marks_ms: CUDA events around every kernel of one serial step); "
                             "kernels.symbolic is the whole pattern build = k_adj + k_sym_tile"})

    cpu = None
    if world == 1:
        # bounded serial CPU baseline (the reference's own single-threaded path) on a sample of the headline workload
        from oracle import oracle as orc
        orc.build()
        ne, dt, t_form, t_sparse, _ = cpu_reference_run(args.cpu_edge, 1)
        cpu = {"value": ne / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "%d^3 H8 diffusion block (%d elements), one serial pass: element loop %.2f s + sparse() %.2f s; host has %d cores"
                         % (args.cpu_edge, ne, t_form, t_sparse, os.cpu_count() or 0)}

    line = {
        "metric": METRIC, "value": r4["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r4["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": r4["workload"],
                   "l2": "working set per rank (element values + CSC + pattern: GBs) >> 126 MB L2; no flush needed",
                   "step": "fresh assembly: pattern cache invalidated before every step; the element integration runs on a second "
                           "stream concurrently with the symbolic phase (phases_ms / marks_ms are measured with that overlap switched off)"},
        "clocks": r4["clocks"],
        "e2e": r4["e2e"],
        "gpu_launches": r4["launches"],
        "roofline": roofline,
        "cpu_baseline": cpu,
        "phases_ms": r4["phases_ms"], "marks_ms": r4["marks_ms"],
        "cached": r4["cached"],
        "nnz_per_s_csc_construction": r4["nnz_per_s_csc_construction"],
        "rank0": r4["rank0"],
    }
    if r4.get("gather_blocks"):
        g = r4["gather_blocks"]
        line["gather_blocks"] = dict(g, value_with_gather=r4["value"] * r4["ms_per_step"] / (r4["ms_per_step"] + g["ms"]),
                                     value_without_gather=r4["value"])
    if r2 is not None:
        line["config2"] = {k: r2[k] for k in ("workload", "value", "ms_per_step", "phases_ms", "marks_ms", "cached", "e2e", "kernels", "launches",
                                              "nnz_per_s_csc_construction")}
    if r5 is not None:
        line["config5"] = {"workload": "BASELINE configs[4]: H20 96^3 volume stiffness + Q4 / T3 surface mass, %d node-owned row blocks" % world, "parts": r5}
    if cpu is None:
        del line["cpu_baseline"]
    _emit(line)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--ref-threads", type=int, default=0, help="reference arm: threads for the element loop (0 = all cores)")
    ap.add_argument("--ref-edge", type=int, default=80, help="reference arm: block edge of the bounded sample")
    ap.add_argument("--cpu-edge", type=int, default=128, help="cpu_baseline leg: block edge of the bounded serial sample")
    ap.add_argument("--no-secondary", action="store_true", help="skip the config-2 block (and config 5 at 8 GPUs)")
    ap.add_argument("--config5", action="store_true", help="add the config-5 block (mixed surface / volume run) at any N; it is on by default at 8 GPUs")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    # stdout carries exactly ONE JSON line: everything else a library prints (e.g. "NCCL version ...") is sent to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global _emit

    def _emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
