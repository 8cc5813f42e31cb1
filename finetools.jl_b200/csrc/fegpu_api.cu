// C ABI of libfinegpu.so (include/fegpu.h): handles, uploads, the three bilinear forms, the generic assembler protocol,
// result access.  No CPU fallback anywhere: every compute call ends in a kernel launch on the context's stream.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "fegpu_internal.h"
#include "fegpu_pattern.h"

static thread_local std::string g_last_error;

int32_t fegpu_fail(fegpu_ctx *ctx, int32_t code, const std::string &msg) {
  g_last_error = msg;
  if (ctx) ctx->err = msg;
  return code;
}

namespace {

int nne_of(int et) {
  switch (et) {
    case FEGPU_T3: return 3; case FEGPU_Q4: return 4; case FEGPU_T4: return 4; case FEGPU_T10: return 10;
    case FEGPU_H8: return 8; case FEGPU_H20: return 20; case FEGPU_H27: return 27;
  }
  return -1;
}
int mdim_of(int et) { return (et == FEGPU_T3 || et == FEGPU_Q4) ? 2 : 3; }

__global__ void k_conn_convert(const int64_t *__restrict__ in, int32_t *__restrict__ out, int64_t n, int64_t nnodes, int *err) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t v = in[i];
  if (v < 1 || v > nnodes) {
    *err = 1;
    v = 1;
  }
  out[i] = (int32_t)(v - 1);
}

// dofnums -> 0-based int32; range checks of assemble! (AssemblyModule.jl:268-273), only for nodes used by elements
__global__ void k_dof_convert(const int64_t *__restrict__ in, int32_t *__restrict__ out, int64_t n, int64_t nnodes, int64_t row_nall,
                              int64_t col_nall, const uint8_t *__restrict__ used, int *err, int32_t *seen) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t v = in[i];
  const bool u = used[i % nnodes] != 0;
  if (u) {
    int code = 0;
    if (v < 1) code = 1;
    else if (v > col_nall) code = 2;
    else if (v > row_nall) code = 4;
    if (code) {
      atomicCAS(err, 0, code);
      v = 1;
    } else if (v <= INT32_MAX) {
      if (atomicAdd(&seen[v - 1], 1) > 0) err[1] = 1;  // two used (node, comp) slots share a dof number
    }
  } else if (v < 1 || v > INT32_MAX) {
    v = 1;
  }
  out[i] = (int32_t)(v - 1);
}

__global__ void k_mark_used(const int32_t *__restrict__ conn, int64_t n, uint8_t *__restrict__ used) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) used[conn[i]] = 1;
}

__global__ void k_rowowned(const int32_t *__restrict__ owner, int64_t nnodes, int32_t rank, uint8_t *__restrict__ owned) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nnodes) owned[i] = (owner[i] == rank) ? 1 : 0;
}

__global__ void k_elem_active(const int32_t *__restrict__ conn, int64_t nelem, int nne, const uint8_t *__restrict__ owned,
                              int32_t *__restrict__ flag) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nelem) return;
  int f = 0;
  for (int a = 0; a < nne; a++) f |= owned[conn[e * nne + a]];
  flag[e] = f;
}

// compact upper-block layout -> full element matrix, emission order (thread per full entry)
// perm (optional): output element r is the element of slot perm[r]
__global__ void k_expand_compact(const double *__restrict__ Vc, double *__restrict__ Vf, int64_t nelem, int nne, int ndn,
                                 const int32_t *__restrict__ perm, int64_t vstride) {
  const int EM = nne * ndn;
  const int64_t EM2 = (int64_t)EM * EM, CS = (int64_t)(nne * (nne + 1) / 2) * ndn * ndn;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nelem * EM2) return;
  const int64_t eo = i / EM2;
  const int64_t e = perm ? (int64_t)perm[eo] : eo;
  const int k = (int)(i - eo * EM2), c = k / EM, r = k - c * EM;
  const int li = r / ndn, p = r - li * ndn, lc = c / ndn, q = c - lc * ndn, nd2 = ndn * ndn;
  const int off = (li <= lc) ? nd2 * (lc * (lc + 1) / 2 + li) + q * ndn + p : nd2 * (li * (li + 1) / 2 + lc) + p * ndn + q;
  Vf[i] = vstride > 0 ? Vc[(int64_t)off * vstride + e] : Vc[e * CS + off];  // value planes: position off of slot e at off * vstride + e
}

// smallest and largest node id used by the active elements: win[0] = max(~node) (so that a zero-initialised slot means "no node"),
// win[1] = max(node + 1).  The symbolic phase runs its per-node passes over [lo, hi) only.
__global__ void __launch_bounds__(256) k_active_window(const int32_t *__restrict__ conn, int64_t nelem, int nne, const int32_t *__restrict__ flag,
                                                       int *__restrict__ win) {
  __shared__ int s_lo[8], s_hi[8];
  int lo = INT32_MAX, hi = 0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += (int64_t)gridDim.x * blockDim.x) {
    if (!flag[e]) continue;
    for (int a = 0; a < nne; a++) {
      const int n = conn[e * nne + a];
      lo = min(lo, n);
      hi = max(hi, n + 1);
    }
  }
  for (int d = 16; d > 0; d >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
  }
  if ((threadIdx.x & 31) == 0) {
    s_lo[threadIdx.x >> 5] = lo;
    s_hi[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; k++) {
      lo = min(lo, s_lo[k]);
      hi = max(hi, s_hi[k]);
    }
    if (hi > 0) {
      atomicMax(&win[0], INT32_MAX - lo);
      atomicMax(&win[1], hi);
    }
  }
}

__global__ void k_compact(const int32_t *__restrict__ flag, const int64_t *__restrict__ pos, int64_t nelem, int32_t *__restrict__ list) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nelem && flag[e]) list[pos[e]] = (int32_t)e;
}

// ---- roofline micro-benchmarks
__global__ void k_dfma_peak(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_copy(const double2 *__restrict__ in, double2 *__restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[i];
}

int32_t finish(fegpu_ctx *ctx) {
  if (!ctx->async) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return FEGPU_OK;
}

// the assembler's result borrows colptr / rowval of a pattern: hold a reference while it does
void asm_set_pattern(fegpu_asm *as, Pattern *p) {
  if (as->pat_src == p) return;
  if (p) fe_pattern_retain(p);
  if (as->pat_src) fe_pattern_free(as->pat_src);
  as->pat_src = p;
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

}  // namespace

// full element matrices permuted: out[r] = in[perm[r]]
// (nne, ndn only matter for the plane form: position k = c * EM + r of the full matrix lives in value plane (b * nne + a) * ndn^2 + j * ndn + i)
__global__ void k_permute_records(const double *__restrict__ in, double *__restrict__ out, int64_t nelem, int64_t rec, const int32_t *__restrict__ perm,
                                  int64_t vstride, int nne, int ndn) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nelem * rec) return;
  const int64_t eo = i / rec, k = i - eo * rec;
  if (vstride > 0) {
    const int EM = nne * ndn, c = (int)(k / EM), r = (int)(k - (int64_t)c * EM);
    const int b = c / ndn, j = c - b * ndn, a = r / ndn, ii = r - a * ndn;
    out[i] = in[((int64_t)(b * nne + a) * (ndn * ndn) + j * ndn + ii) * vstride + perm[eo]];
  } else {
    out[i] = in[(int64_t)perm[eo] * rec + k];
  }
}

int32_t fe_expand_compact(fegpu_ctx *ctx, const double *d_Vc, double *d_Vfull, int64_t nelem, int nne, int ndn, const int32_t *d_perm, int64_t vstride) {
  const int64_t n = nelem * (int64_t)(nne * ndn) * (nne * ndn);
  if (n == 0) return FEGPU_OK;
  k_expand_compact<<<grid_for(n, 256), 256, 0, ctx->stream>>>(d_Vc, d_Vfull, nelem, nne, ndn, d_perm, vstride);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return FEGPU_OK;
}

extern "C" {

const char *fegpu_last_error(fegpu_ctx *ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int32_t fegpu_create(fegpu_ctx **out, int32_t device) {
  if (!out) return fegpu_fail(nullptr, FEGPU_ERR_ARG, "ctx pointer is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fegpu_fail(nullptr, FEGPU_ERR_CUDA, std::string("no CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fegpu_fail(nullptr, FEGPU_ERR_ARG, "device index out of range");
  CUDA_TRY(nullptr, cudaSetDevice(device));
  CUDA_TRY(nullptr, cudaFree(0));
  fegpu_ctx *ctx = new fegpu_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  CUDA_TRY(nullptr, cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  {  // the symbolic phase of a fresh assembly runs on this stream at the highest priority (see run_bilform)
    int lo = 0, hi = 0;
    CUDA_TRY(nullptr, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUDA_TRY(nullptr, cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, hi));
  }
  CUDA_TRY(nullptr, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  if (const char *e = std::getenv("FEGPU_OVERLAP")) ctx->overlap = std::atoi(e) != 0;
  *out = ctx;
  return FEGPU_OK;
}

int32_t fegpu_set_overlap(fegpu_ctx *ctx, int32_t on) {
  if (!ctx) return FEGPU_ERR_ARG;
  ctx->overlap = on != 0;
  return FEGPU_OK;
}

int32_t fegpu_destroy(fegpu_ctx *ctx) {
  if (!ctx) return FEGPU_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->stream2) { cudaStreamSynchronize(ctx->stream2); cudaStreamDestroy(ctx->stream2); }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->xfer) fe_transfer_free(ctx->xfer);
  fe_dev_cache_destroy(ctx);
  delete ctx;
  return FEGPU_OK;
}

int32_t fegpu_set_stream(fegpu_ctx *ctx, void *s) {
  if (!ctx) return FEGPU_ERR_ARG;
  ctx->stream = (cudaStream_t)s;
  return FEGPU_OK;
}
int32_t fegpu_set_async(fegpu_ctx *ctx, int32_t on) {
  if (!ctx) return FEGPU_ERR_ARG;
  ctx->async = on != 0;
  return FEGPU_OK;
}
int32_t fegpu_cache_release(fegpu_ctx *ctx) {
  if (!ctx) return FEGPU_ERR_ARG;
  DeviceGuard g(ctx->device);
  fe_dev_cache_trim(ctx);
  return FEGPU_OK;
}
int32_t fegpu_synchronize(fegpu_ctx *ctx) {
  if (!ctx) return FEGPU_ERR_ARG;
  DeviceGuard g(ctx->device);
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return FEGPU_OK;
}
int64_t fegpu_launch_count(fegpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

int32_t fegpu_host_alloc(void **p, int64_t bytes) {
  if (!p || bytes < 0) return fegpu_fail(nullptr, FEGPU_ERR_ARG, "bad argument");
  *p = nullptr;
  cudaError_t e = cudaHostAlloc(p, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocPortable);
  if (e != cudaSuccess) return fegpu_fail(nullptr, FEGPU_ERR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
  return FEGPU_OK;
}

int32_t fegpu_host_free(void *p) {
  if (p) cudaFreeHost(p);
  return FEGPU_OK;
}

int32_t fegpu_marks_begin(fegpu_ctx *ctx) {
  if (!ctx) return FEGPU_ERR_ARG;
  ctx->marks_on = true;
  ctx->nmarks = 0;
  return FEGPU_OK;
}

int32_t fegpu_marks_read(fegpu_ctx *ctx, char *buf, int64_t cap) {
  if (!ctx || !buf || cap < 1) return fegpu_fail(ctx, FEGPU_ERR_ARG, "NULL argument");
  DeviceGuard g(ctx->device);
  ctx->marks_on = false;
  std::string out;
  if (ctx->nmarks > 0) CUDA_TRY(ctx, cudaEventSynchronize(ctx->marks[ctx->nmarks - 1].ev));
  for (int i = 1; i < ctx->nmarks; i++) {
    float ms = 0.f;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->marks[i - 1].ev, ctx->marks[i].ev));
    char tmp[160];
    std::snprintf(tmp, sizeof(tmp), "%s=%.6f;", ctx->marks[i].name, (double)ms);
    out += tmp;
  }
  ctx->nmarks = 0;
  const size_t n = std::min<size_t>(out.size(), (size_t)cap - 1);
  std::memcpy(buf, out.data(), n);
  buf[n] = 0;
  return FEGPU_OK;
}

int32_t fegpu_measure_peaks(fegpu_ctx *ctx, double *dfma_tflops, double *copy_gbs) {
  if (!ctx) return FEGPU_ERR_ARG;
  DeviceGuard g(ctx->device);
  cudaStream_t st = ctx->stream;
  cudaEvent_t e0, e1;
  CUDA_TRY(ctx, cudaEventCreate(&e0));
  CUDA_TRY(ctx, cudaEventCreate(&e1));
  float ms = 0;
  if (dfma_tflops) {
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 1 << 14;
    double *d = nullptr;
    CUDA_TRY(ctx, cudaMalloc((void **)&d, sizeof(double) * blocks * threads));
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
      CUDA_TRY(ctx, cudaEventRecord(e0, st));
      k_dfma_peak<<<blocks, threads, 0, st>>>(d, iters);
      ctx->launches++;
      CUDA_TRY(ctx, cudaEventRecord(e1, st));
      CUDA_TRY(ctx, cudaEventSynchronize(e1));
      CUDA_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
      double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
      if (rep > 0) best = std::max(best, tf);
    }
    *dfma_tflops = best;
    cudaFree(d);
  }
  if (copy_gbs) {
    const int64_t n = (int64_t)1 << 26;  // double2: 1 GiB read + 1 GiB write
    double2 *a = nullptr, *b = nullptr;
    CUDA_TRY(ctx, cudaMalloc((void **)&a, sizeof(double2) * n));
    CUDA_TRY(ctx, cudaMalloc((void **)&b, sizeof(double2) * n));
    CUDA_TRY(ctx, cudaMemsetAsync(a, 0, sizeof(double2) * n, st));
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
      CUDA_TRY(ctx, cudaEventRecord(e0, st));
      k_copy<<<ctx->sm_count * 16, 512, 0, st>>>(a, b, n);
      ctx->launches++;
      CUDA_TRY(ctx, cudaEventRecord(e1, st));
      CUDA_TRY(ctx, cudaEventSynchronize(e1));
      CUDA_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0) best = std::max(best, 2.0 * sizeof(double2) * n / (ms * 1e-3) / 1e9);
    }
    *copy_gbs = best;
    cudaFree(a);
    cudaFree(b);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return FEGPU_OK;
}

// ------------------------------------------------------------------------------------------------- mesh
int32_t fegpu_mesh_upload(fegpu_ctx *ctx, int32_t etype, int64_t nelem, const int64_t *conn, int64_t nnodes, int32_t sdim,
                          const double *xyz, fegpu_mesh **out) {
  if (!ctx || !out) return fegpu_fail(ctx, FEGPU_ERR_ARG, "NULL argument");
  *out = nullptr;
  const int nne = nne_of(etype);
  if (nne < 0) return fegpu_fail(ctx, FEGPU_ERR_ARG, "unknown element type");
  if (sdim < mdim_of(etype) || sdim > 3) return fegpu_fail(ctx, FEGPU_ERR_ARG, "space dimension does not fit the element manifold");
  if (nelem < 0 || nnodes < 0 || nnodes >= INT32_MAX || nelem >= INT32_MAX) return fegpu_fail(ctx, FEGPU_ERR_ARG, "mesh too large for int32 ids");
  if ((nelem > 0 && !conn) || (nnodes > 0 && !xyz)) return fegpu_fail(ctx, FEGPU_ERR_ARG, "NULL mesh arrays");
  DeviceGuard g(ctx->device);
  fegpu_mesh *m = new fegpu_mesh();
  m->ctx = ctx; m->etype = etype; m->nne = nne; m->mdim = mdim_of(etype); m->sdim = sdim; m->nelem = nelem; m->nnodes = nnodes;
  m->nactive = nelem;
  m->win_lo = 0;
  m->win_hi = nnodes;
  cudaStream_t st = ctx->stream;
  int64_t *d_c64 = nullptr;
  int *d_err = nullptr;
  auto fail = [&](int32_t code, const std::string &msg) {
    cudaFree(d_c64); cudaFree(d_err);
    fegpu_mesh_destroy(m);
    return fegpu_fail(ctx, code, msg);
  };
  for (int d = 0; d < sdim; d++) {  // bounding box (host pass over the coordinates the caller just handed us)
    double lo = nnodes ? xyz[(size_t)d * nnodes] : 0.0, hi = lo;
    for (int64_t i = 1; i < nnodes; i++) {
      const double v = xyz[(size_t)d * nnodes + i];
      lo = v < lo ? v : lo;
      hi = v > hi ? v : hi;
    }
    m->bbox_lo[d] = lo;
    m->bbox_hi[d] = hi;
  }
  const size_t nc = (size_t)nelem * nne;
  cudaError_t e;
#define MT(expr) if ((e = (expr)) != cudaSuccess) return fail(FEGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e))
  MT(cudaMalloc((void **)&m->d_conn, sizeof(int32_t) * std::max<size_t>(nc, 1)));
  MT(cudaMalloc((void **)&m->d_xyz, sizeof(double) * std::max<size_t>((size_t)nnodes * sdim, 1)));
  MT(cudaMalloc((void **)&d_c64, sizeof(int64_t) * std::max<size_t>(nc, 1)));
  MT(cudaMalloc((void **)&d_err, sizeof(int)));
  MT(cudaMemsetAsync(d_err, 0, sizeof(int), st));
  if (nc) MT(cudaMemcpyAsync(d_c64, conn, sizeof(int64_t) * nc, cudaMemcpyHostToDevice, st));
  if (nnodes) MT(cudaMemcpyAsync(m->d_xyz, xyz, sizeof(double) * (size_t)nnodes * sdim, cudaMemcpyHostToDevice, st));
  if (nc) {
    k_conn_convert<<<grid_for((int64_t)nc, 256), 256, 0, st>>>(d_c64, m->d_conn, (int64_t)nc, nnodes, d_err);
    ctx->launches++;
  }
  int h_err = 0;
  MT(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
  MT(cudaStreamSynchronize(st));
#undef MT
  if (h_err) return fail(FEGPU_ERR_ARG, "connectivity refers to a node outside 1..nnodes");
  cudaFree(d_c64);
  cudaFree(d_err);
  d_c64 = nullptr;
  d_err = nullptr;
  {  // internal element order: ascending smallest node id (locality of the element records with respect to the node-ordered output)
    const int32_t s = fe_order_elements(m);
    if (s != FEGPU_OK) { fegpu_mesh_destroy(m); return s; }
  }
  *out = m;
  return FEGPU_OK;
}

int32_t fegpu_mesh_destroy(fegpu_mesh *m) {
  if (!m) return FEGPU_OK;
  DeviceGuard g(m->ctx->device);
  cudaFree(m->d_uvel);
  cudaFree(m->d_orig);
  cudaFree(m->d_conn); cudaFree(m->d_xyz); cudaFree(m->d_tab); cudaFree(m->d_w); cudaFree(m->d_elem_list); cudaFree(m->d_rowowned);
  delete m;
  return FEGPU_OK;
}

int32_t fegpu_geom_update(fegpu_mesh *m, const double *xyz) {
  if (!m || !xyz) return fegpu_fail(m ? m->ctx : nullptr, FEGPU_ERR_ARG, "NULL argument");
  DeviceGuard g(m->ctx->device);
  CUDA_TRY(m->ctx, cudaMemcpyAsync(m->d_xyz, xyz, sizeof(double) * (size_t)m->nnodes * m->sdim, cudaMemcpyHostToDevice, m->ctx->stream));
  CUDA_TRY(m->ctx, cudaStreamSynchronize(m->ctx->stream));  // the host array may go away after return
  return FEGPU_OK;
}

int32_t fegpu_geom_update_window(fegpu_mesh *m, const double *xyz) {
  if (!m || !xyz) return fegpu_fail(m ? m->ctx : nullptr, FEGPU_ERR_ARG, "NULL argument");
  DeviceGuard g(m->ctx->device);
  const int64_t lo = m->win_lo, nw = m->win_hi - m->win_lo;
  if (nw > 0)
    for (int d = 0; d < m->sdim; d++)
      CUDA_TRY(m->ctx, cudaMemcpyAsync(m->d_xyz + (size_t)d * m->nnodes + lo, xyz + (size_t)d * m->nnodes + lo, sizeof(double) * (size_t)nw,
                                       cudaMemcpyHostToDevice, m->ctx->stream));
  CUDA_TRY(m->ctx, cudaStreamSynchronize(m->ctx->stream));  // the host array may go away after return
  return FEGPU_OK;
}

int32_t fegpu_mesh_window(fegpu_mesh *m, int64_t *lo, int64_t *hi, int64_t *nactive) {
  if (!m) return fegpu_fail(nullptr, FEGPU_ERR_ARG, "NULL mesh");
  if (lo) *lo = m->win_lo;
  if (hi) *hi = m->win_hi;
  if (nactive) *nactive = m->nactive;
  return FEGPU_OK;
}

int32_t fegpu_rule_set(fegpu_mesh *m, int32_t npts, const double *Ns, const double *gradNpar, const double *w) {
  if (!m || !Ns || !gradNpar || !w) return fegpu_fail(m ? m->ctx : nullptr, FEGPU_ERR_ARG, "NULL argument");
  fegpu_ctx *ctx = m->ctx;
  if (npts < 1 || npts > FEGPU_MAX_NPTS || (int64_t)npts * m->nne * (1 + m->mdim) > FEGPU_TAB_DOUBLES)
    return fegpu_fail(ctx, FEGPU_ERR_ARG, "quadrature rule too large for the on-chip tables");
  DeviceGuard g(ctx->device);
  const size_t nN = (size_t)npts * m->nne, nD = nN * m->mdim;
  m->h_tab.assign(Ns, Ns + nN);
  m->h_tab.insert(m->h_tab.end(), gradNpar, gradNpar + nD);
  m->h_w.assign(w, w + npts);
  cudaFree(m->d_tab);
  cudaFree(m->d_w);
  m->d_tab = nullptr;
  m->d_w = nullptr;
  CUDA_TRY(ctx, cudaMalloc((void **)&m->d_tab, sizeof(double) * (nN + nD)));
  CUDA_TRY(ctx, cudaMalloc((void **)&m->d_w, sizeof(double) * npts));
  CUDA_TRY(ctx, cudaMemcpyAsync(m->d_tab, m->h_tab.data(), sizeof(double) * (nN + nD), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(m->d_w, m->h_w.data(), sizeof(double) * npts, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  m->npts = npts;
  return FEGPU_OK;
}

int32_t fegpu_csys_set(fegpu_mesh *m, const double *csmat) {
  if (!m) return FEGPU_ERR_ARG;
  if (csmat && m->sdim != m->mdim) return fegpu_fail(m->ctx, FEGPU_ERR_ARG, "a material coordinate system matrix needs sdim == manifold dimension");
  m->use_rm = csmat != nullptr;
  const int n = m->sdim * m->mdim;
  for (int i = 0; i < 9; i++) m->rm[i] = (csmat && i < n) ? csmat[i] : 0.0;
  if (m->use_rm) {  // the identity takes the specialised kernels
    bool ident = true;
    for (int a = 0; a < m->sdim; a++)
      for (int b = 0; b < m->mdim; b++) ident = ident && m->rm[a + m->sdim * b] == (a == b ? 1.0 : 0.0);
    if (ident) m->use_rm = false;
  }
  return FEGPU_OK;
}

int32_t fegpu_otherdimension_set(fegpu_mesh *m, double otherdim) {
  if (!m) return FEGPU_ERR_ARG;
  m->otherdim = otherdim;
  return FEGPU_OK;
}

int32_t fegpu_partition_set(fegpu_mesh *m, const int32_t *node_owner, int32_t my_rank) {
  if (!m) return fegpu_fail(nullptr, FEGPU_ERR_ARG, "NULL mesh");
  fegpu_ctx *ctx = m->ctx;
  DeviceGuard g(ctx->device);
  cudaStream_t st = ctx->stream;
  m->topo_version++;
  cudaFree(m->d_elem_list);
  cudaFree(m->d_rowowned);
  m->d_elem_list = nullptr;
  m->elem_base = 0;
  m->d_rowowned = nullptr;
  m->partitioned = false;
  m->nactive = m->nelem;
  m->win_lo = 0;
  m->win_hi = m->nnodes;
  m->own_contig = false;
  m->own_lo = m->own_hi = 0;
  if (!node_owner) return FEGPU_OK;
  {  // are the owned nodes one contiguous range (slab partitions, meshes reordered by partition)?  Then the kernels test row
     // ownership with two comparisons instead of a byte load per candidate
    int64_t first = -1, last = -1, cnt = 0;
    for (int64_t i = 0; i < m->nnodes; i++)
      if (node_owner[i] == my_rank) {
        if (first < 0) first = i;
        last = i;
        cnt++;
      }
    m->own_contig = cnt > 0 && cnt == last - first + 1;
    m->own_lo = cnt > 0 ? first : 0;
    m->own_hi = cnt > 0 ? last + 1 : 0;
  }
  int32_t *d_owner = nullptr, *d_flag = nullptr;
  int64_t *d_pos = nullptr;
  int *d_win = nullptr;
  auto cleanup = [&]() { cudaFree(d_owner); cudaFree(d_flag); cudaFree(d_pos); cudaFree(d_win); };
  cudaError_t e;
#define PT(expr) if ((e = (expr)) != cudaSuccess) { cleanup(); return fegpu_fail(ctx, FEGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e)); }
  PT(cudaMalloc((void **)&d_owner, sizeof(int32_t) * std::max<int64_t>(m->nnodes, 1)));
  PT(cudaMalloc((void **)&m->d_rowowned, std::max<int64_t>(m->nnodes, 1)));
  PT(cudaMalloc((void **)&d_flag, sizeof(int32_t) * std::max<int64_t>(m->nelem, 1)));
  PT(cudaMalloc((void **)&d_pos, sizeof(int64_t) * (m->nelem + 1)));
  PT(cudaMemcpyAsync(d_owner, node_owner, sizeof(int32_t) * m->nnodes, cudaMemcpyHostToDevice, st));
  if (m->nnodes) k_rowowned<<<grid_for(m->nnodes, 256), 256, 0, st>>>(d_owner, m->nnodes, my_rank, m->d_rowowned);
  if (m->nelem) k_elem_active<<<grid_for(m->nelem, 256), 256, 0, st>>>(m->d_conn, m->nelem, m->nne, m->d_rowowned, d_flag);
  ctx->launches += 2;
  int64_t nact = 0;
  int32_t s = fe_exclusive_scan_i32_to_i64(ctx, d_flag, d_pos, m->nelem, 0, true, &nact);
  if (s != FEGPU_OK) { cleanup(); return s; }
  PT(cudaMalloc((void **)&m->d_elem_list, sizeof(int32_t) * std::max<int64_t>(nact, 1)));
  if (m->nelem) {
    k_compact<<<grid_for(m->nelem, 256), 256, 0, st>>>(d_flag, d_pos, m->nelem, m->d_elem_list);
    ctx->launches++;
  }
  // node window of the active elements (a z-slab of a block mesh owns a contiguous node range)
  int h_win[2] = {0, 0};
  PT(cudaMalloc((void **)&d_win, sizeof(int) * 2));
  PT(cudaMemsetAsync(d_win, 0, sizeof(int) * 2, st));
  if (m->nelem) {
    k_active_window<<<(unsigned)std::min<int64_t>(grid_for(m->nelem, 256), (int64_t)ctx->sm_count * 8), 256, 0, st>>>(m->d_conn, m->nelem, m->nne, d_flag, d_win);
    ctx->launches++;
  }
  PT(cudaMemcpyAsync(h_win, d_win, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
  int32_t h_ends[2] = {0, -1};
  if (nact > 0) {
    PT(cudaMemcpyAsync(&h_ends[0], m->d_elem_list, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    PT(cudaMemcpyAsync(&h_ends[1], m->d_elem_list + (nact - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  }
  PT(cudaStreamSynchronize(st));
#undef PT
  cleanup();
  if (nact > 0 && (int64_t)h_ends[1] - h_ends[0] + 1 == nact) {
    // the active elements are one contiguous range of the internal order (slabs of a mesh with node locality): slot s is element
    // elem_base + s, no list and no indirection in the kernels
    cudaFree(m->d_elem_list);
    m->d_elem_list = nullptr;
    m->elem_base = h_ends[0];
  }
  m->nactive = nact;
  m->partitioned = true;
  if (h_win[1] > 0) {
    m->win_lo = (int64_t)(INT32_MAX - h_win[0]);
    m->win_hi = (int64_t)h_win[1];
  } else {
    m->win_lo = m->win_hi = 0;  // no active element
  }
  return FEGPU_OK;
}

// ------------------------------------------------------------------------------------------------- dof map
int32_t fegpu_dofmap_upload(fegpu_ctx *ctx, fegpu_mesh *mesh, int32_t ndn, const int64_t *dofnums, int64_t row_nall, int64_t col_nall,
                            fegpu_dofmap **out) {
  if (!ctx || !mesh || !dofnums || !out) return fegpu_fail(ctx, FEGPU_ERR_ARG, "NULL argument");
  *out = nullptr;
  if (ndn < 1 || ndn > 6) return fegpu_fail(ctx, FEGPU_ERR_ARG, "dofs per node must be in 1..6");
  if (row_nall < 0 || col_nall < 0 || row_nall >= INT32_MAX || col_nall >= INT32_MAX)
    return fegpu_fail(ctx, FEGPU_ERR_ARG, "matrix dimension outside int32 range");
  DeviceGuard g(ctx->device);
  cudaStream_t st = ctx->stream;
  fegpu_dofmap *d = new fegpu_dofmap();
  d->ctx = ctx; d->mesh = mesh; d->ndn = ndn; d->row_nall = row_nall; d->col_nall = col_nall;
  const int64_t n = mesh->nnodes * ndn;
  int64_t *d64 = nullptr;
  uint8_t *d_used = nullptr;
  int *d_err = nullptr;
  int32_t *d_seen = nullptr;
  auto fail = [&](int32_t code, const std::string &msg) {
    cudaFree(d64); cudaFree(d_used); cudaFree(d_err); cudaFree(d_seen);
    fegpu_dofmap_destroy(d);
    return fegpu_fail(ctx, code, msg);
  };
  cudaError_t e;
#define DT(expr) if ((e = (expr)) != cudaSuccess) return fail(FEGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e))
  const int64_t nseen = std::max<int64_t>(std::max(row_nall, col_nall), 1);
  DT(cudaMalloc((void **)&d->d_dof, sizeof(int32_t) * std::max<int64_t>(n, 1)));
  DT(cudaMalloc((void **)&d64, sizeof(int64_t) * std::max<int64_t>(n, 1)));
  DT(cudaMalloc((void **)&d_used, std::max<int64_t>(mesh->nnodes, 1)));
  DT(cudaMalloc((void **)&d_err, sizeof(int) * 2));
  DT(cudaMalloc((void **)&d_seen, sizeof(int32_t) * nseen));
  DT(cudaMemsetAsync(d_used, 0, std::max<int64_t>(mesh->nnodes, 1), st));
  DT(cudaMemsetAsync(d_err, 0, sizeof(int) * 2, st));
  DT(cudaMemsetAsync(d_seen, 0, sizeof(int32_t) * nseen, st));
  if (n) DT(cudaMemcpyAsync(d64, dofnums, sizeof(int64_t) * n, cudaMemcpyHostToDevice, st));
  const int64_t nc = mesh->nelem * mesh->nne;
  if (nc) {
    k_mark_used<<<grid_for(nc, 256), 256, 0, st>>>(mesh->d_conn, nc, d_used);
    ctx->launches++;
  }
  if (n) {
    k_dof_convert<<<grid_for(n, 256), 256, 0, st>>>(d64, d->d_dof, n, mesh->nnodes, row_nall, col_nall, d_used, d_err, d_seen);
    ctx->launches++;
  }
  int h_err[2] = {0, 0};
  DT(cudaMemcpyAsync(h_err, d_err, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
  DT(cudaStreamSynchronize(st));
#undef DT
  if (h_err[0]) {
    static const char *msgs[] = {"", "Column degree of freedom < 1", "Column degree of freedom > size", "Row degree of freedom < 1",
                                 "Row degree of freedom > size"};
    return fail(FEGPU_ERR_COL_LT1 - (h_err[0] - 1), msgs[h_err[0]]);
  }
  d->injective = (h_err[1] == 0);
  cudaFree(d64); cudaFree(d_used); cudaFree(d_err); cudaFree(d_seen);
  *out = d;
  return FEGPU_OK;
}

int32_t fegpu_dofmap_destroy(fegpu_dofmap *d) {
  if (!d) return FEGPU_OK;
  DeviceGuard g(d->ctx->device);
  cudaFree(d->d_dof);
  if (d->pat) fe_pattern_free(d->pat);
  delete d;
  return FEGPU_OK;
}

int32_t fegpu_pattern_invalidate(fegpu_dofmap *d) {
  if (!d) return FEGPU_ERR_ARG;
  DeviceGuard g(d->ctx->device);
  if (d->pat) fe_pattern_free(d->pat);
  d->pat = nullptr;
  d->pat_topo_version = 0;
  return FEGPU_OK;
}

int32_t fegpu_pattern_path(fegpu_dofmap *d) {
  if (!d || !d->pat || d->pat_topo_version != d->mesh->topo_version) return 0;
  return d->pat->tile ? 2 : 1;
}

// ------------------------------------------------------------------------------------------------- assembler
int32_t fegpu_asm_create(fegpu_ctx *ctx, fegpu_asm **out) {
  if (!ctx || !out) return fegpu_fail(ctx, FEGPU_ERR_ARG, "NULL argument");
  DeviceGuard g(ctx->device);
  fegpu_asm *a = new fegpu_asm();
  a->ctx = ctx;
  for (auto &ev : a->ev) CUDA_TRY(ctx, cudaEventCreate(&ev));
  *out = a;
  return FEGPU_OK;
}

int32_t fegpu_asm_set_symmetric(fegpu_asm *as, int32_t on) {
  if (!as) return fegpu_fail(nullptr, FEGPU_ERR_ARG, "NULL assembler");
  if (as->started) return fegpu_fail(as->ctx, FEGPU_ERR_STATE, "cannot switch the assembler kind inside an assembly");
  as->symmetric = on != 0;
  return FEGPU_OK;
}

int32_t fegpu_asm_set_lumping(fegpu_asm *as, int32_t mode) {
  if (!as) return fegpu_fail(nullptr, FEGPU_ERR_ARG, "NULL assembler");
  if (mode < 0 || mode > 2) return fegpu_fail(as->ctx, FEGPU_ERR_ARG, "lumping mode must be 0 (none), 1 (diagonal) or 2 (HRZ)");
  if (as->started) return fegpu_fail(as->ctx, FEGPU_ERR_STATE, "cannot switch the assembler kind inside an assembly");
  as->lump = mode;
  return FEGPU_OK;
}

int32_t fegpu_asm_destroy(fegpu_asm *a) {
  if (!a) return FEGPU_OK;
  DeviceGuard g(a->ctx->device);
  asm_set_pattern(a, nullptr);
  cudaFree(a->d_V); cudaFree(a->d_nzval); cudaFree(a->own_colptr); cudaFree(a->own_rowval); cudaFree(a->d_F);
  cudaFree(a->view.own_colptr); cudaFree(a->view.own_rowval); cudaFree(a->view.own_nzval);
  for (auto &ev : a->ev)
    if (ev) cudaEventDestroy(ev);
  delete a;
  return FEGPU_OK;
}

static int32_t run_lumped(fegpu_mesh *mesh, fegpu_dofmap *dm, const FormArgs &fa, fegpu_asm *as);  // diagonal / HRZ assemblers

static int32_t run_bilform(fegpu_mesh *mesh, fegpu_dofmap *dm, const FormArgs &fa, fegpu_asm *as) {
  if (!mesh || !dm || !as) return fegpu_fail(mesh ? mesh->ctx : nullptr, FEGPU_ERR_ARG, "NULL argument");
  fegpu_ctx *ctx = mesh->ctx;
  if (dm->mesh != mesh || as->ctx != ctx || dm->ctx != ctx) return fegpu_fail(ctx, FEGPU_ERR_ARG, "handles belong to different meshes / contexts");
  if (dm->ndn != fa.ndn) return fegpu_fail(ctx, FEGPU_ERR_ARG, "Wrong size of matrix: dofs per node of the field do not fit the form");
  if (as->lump) return run_lumped(mesh, dm, fa, as);
  DeviceGuard g(ctx->device);
  cudaStream_t st = ctx->stream;
  const int EM = mesh->nne * fa.ndn;
  const int64_t ntrip = mesh->nactive * (int64_t)EM * EM;
  as->have_result = false;
  asm_set_pattern(as, nullptr);
  as->view.active = false;
  as->started = false;
  // 1. symbolic phase (cached in the dof map) and 2. element integration.  The symbolic phase decides the layout the
  // integration kernel writes (compact only on the structured path), but nothing else connects the two, so on a fresh
  // assembly the pattern build runs on the context's second, high-priority stream and the integration is launched on the
  // caller's stream as soon as the build knows it will take the structured path; the numeric phase waits for both.
  CUDA_TRY(ctx, cudaEventRecord(as->ev[0], st));
  fe_mark(ctx, "start");
  bool fast = fe_pattern_usable(dm);
  as->pattern_cached = false;
  FormArgs fa2 = fa;
  static const bool compact_off = std::getenv("FEGPU_COMPACT") && std::atoi(std::getenv("FEGPU_COMPACT")) == 0;  // A/B knob
  as->V_n = ntrip;
  as->last_EM = EM;
  static const bool planes_off = std::getenv("FEGPU_PLANES") && std::atoi(std::getenv("FEGPU_PLANES")) == 0;  // A/B knob
  const bool can_compact = fe_integrate_supports_compact(mesh, fa) && !compact_off;
  const bool can_planes = fe_integrate_supports_planes(mesh, fa) && !planes_off;
  bool integrated = false, sym_timed = false;
  auto integrate = [&](bool compact, bool planes) -> int32_t {  // always on the caller's stream
    if (integrated && fa2.compact == compact && fa2.planes == planes) return FEGPU_OK;
    fa2.compact = compact;
    fa2.planes = planes;
    fa2.vstride = planes ? ((mesh->nactive + 31) & ~(int64_t)31) : 0;
    const int64_t per_elem = compact ? fe_compact_size(mesh->nne, fa.ndn) : (int64_t)EM * EM;
    cudaStream_t cur = ctx->stream;
    ctx->stream = st;
    int32_t r = fe_asm_reserve(as, &as->d_V, &as->V_cap, (size_t)std::max<int64_t>(planes ? per_elem * fa2.vstride : mesh->nactive * per_elem, 1));
    as->V_compact = compact;
    as->V_planes = planes;
    as->V_stride = fa2.vstride;
    if (r == FEGPU_OK && cudaEventRecord(as->ev[4], st) != cudaSuccess) r = fegpu_fail(ctx, FEGPU_ERR_CUDA, "event record failed");
    if (r == FEGPU_OK) r = fe_integrate(mesh, fa2, as->d_V);
    if (r == FEGPU_OK && cudaEventRecord(as->ev[2], st) != cudaSuccess) r = fegpu_fail(ctx, FEGPU_ERR_CUDA, "event record failed");
    if (r == FEGPU_OK) fe_mark(ctx, "integrate");
    ctx->stream = cur;
    integrated = r == FEGPU_OK;
    return r;
  };
  if (fast) {
    if (!dm->pat || dm->pat_topo_version != mesh->topo_version) {
      const bool ov = ctx->overlap && ctx->stream2;
      cudaStream_t sym = ov ? ctx->stream2 : st;
      // called from inside the build (which runs on `sym`) as soon as it knows which structured path it takes; a second call
      // with another layout (the thread-per-node attempt failed its preconditions) integrates again
      std::function<int32_t(bool)> fork = [&](bool tile) -> int32_t { return integrate(can_compact, tile && can_planes); };
      if (ov) {  // the symbolic phase starts after everything queued on the caller's stream so far
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, st));
        CUDA_TRY(ctx, cudaStreamWaitEvent(sym, ctx->ev_fork, 0));
        ctx->stream = sym;
      }
      const int32_t bs = fe_pattern_build(dm, ov ? &fork : nullptr);
      ctx->stream = st;
      FE_TRY(bs);
      fast = fe_pattern_usable(dm) && dm->pat;  // the build may discover a degenerate mesh
      if (ov) {
        if (dm->pat) fe_pattern_set_stream(dm->pat, st);  // its stream-ordered frees follow the caller's stream from now on
        CUDA_TRY(ctx, cudaEventRecord(as->ev[1], sym));
        CUDA_TRY(ctx, cudaStreamWaitEvent(st, as->ev[1], 0));  // join
        sym_timed = true;
      }
    } else {
      as->pattern_cached = true;
    }
  }
  if (!sym_timed) CUDA_TRY(ctx, cudaEventRecord(as->ev[1], st));
  // the layout the numeric phase needs: compact / planes only on the structured paths, full element-major for the sort path
  FE_TRY(integrate(fast && can_compact, fast && can_planes && fe_pattern_is_tile(dm->pat)));
  CUDA_TRY(ctx, cudaEventRecord(as->ev[5], st));
  FE_TRACE("bilform: before numeric phase");
  // 3. numeric CSC phase
  if (fast) {
    const int64_t nnz = fe_pattern_nnz(dm->pat);
    FE_TRY(fe_asm_reserve(as, &as->d_nzval, &as->nz_cap, (size_t)std::max<int64_t>(nnz, 1)));
    FE_TRY(fe_gather(dm, as->d_V, fa2.compact, as->d_nzval, fa2.planes, fa2.vstride));
    fe_mark(ctx, "gather");
    as->nnz = nnz;
    as->nrows = dm->row_nall;
    as->ncols = dm->col_nall;
    as->d_colptr = fe_pattern_colptr(dm->pat);
    as->d_rowval = fe_pattern_rowval(dm->pat);
    asm_set_pattern(as, dm->pat);
  } else {
    if (mesh->partitioned) return fegpu_fail(ctx, FEGPU_ERR_ARG, "row-block partitioning needs an injective dof map and non-degenerate elements");
    int64_t *dI = nullptr, *dJ = nullptr;
    CUDA_TRY(ctx, cudaMalloc((void **)&dI, sizeof(int64_t) * std::max<int64_t>(ntrip, 1)));
    cudaError_t e = cudaMalloc((void **)&dJ, sizeof(int64_t) * std::max<int64_t>(ntrip, 1));
    if (e != cudaSuccess) { cudaFree(dI); return fegpu_fail(ctx, FEGPU_ERR_CUDA, cudaGetErrorString(e)); }
    int32_t s = fe_emit_ij(dm, dI, dJ);
    if (s == FEGPU_OK) s = fe_coo_to_csc(as, ntrip, dI, dJ, as->d_V, dm->row_nall, dm->col_nall);
    cudaFree(dI);
    cudaFree(dJ);
    FE_TRY(s);
  }
  CUDA_TRY(ctx, cudaEventRecord(as->ev[3], st));
  FE_TRACE("bilform: gather queued");
  as->ev_valid = true;
  as->have_result = true;
  if (as->symmetric) {
    // SysmatAssemblerSparseSymm: the element matrices of these forms are symmetric, so S + transpose(S) (diagonal halved) is
    // the full assembly up to summation order; what differs is the pattern: entries that sum to exactly 0.0 are not stored
    bool symm = fe_form_values_symmetric(fa.form);
    if (fa.form == FORM_DOT) {  // bilform_dot is symmetric exactly when its ndn x ndn coefficient is
      symm = true;
      for (int p = 0; p < fa.ndn; p++)
        for (int q = 0; q < p; q++) symm = symm && fa.coef[p + fa.ndn * q] == fa.coef[q + fa.ndn * p];
    }
    if (!symm) return fegpu_fail(ctx, FEGPU_ERR_ARG, "the symmetric assembler needs symmetric element matrices (non-symmetric coefficient)");
    FE_TRY(fe_csc_view(as, 1, as->nrows, 1, as->ncols, true));
  }
  return finish(ctx);
}

int32_t fegpu_bilform_diffusion(fegpu_mesh *mesh, fegpu_dofmap *dm, int32_t kappa_kind, const double *kappa, fegpu_asm *as) {
  if (!mesh || !kappa) return fegpu_fail(mesh ? mesh->ctx : nullptr, FEGPU_ERR_ARG, "NULL argument");
  if (mesh->sdim != mesh->mdim) return fegpu_fail(mesh->ctx, FEGPU_ERR_ARG, "bilform_diffusion needs sdim == manifold dimension");
  FormArgs fa;
  std::memset(&fa, 0, sizeof(fa));
  fa.form = (kappa_kind == 0) ? FORM_DIFF_ISO : FORM_DIFF_GEN;
  fa.ndn = 1;
  const int nk = (kappa_kind == 0) ? 1 : mesh->mdim * mesh->mdim;
  for (int i = 0; i < nk; i++) fa.coef[i] = kappa[i];
  fa.m = 3;
  fa.otherdim = mesh->otherdim;
  fa.use_rm = mesh->use_rm;
  for (int i = 0; i < 9; i++) fa.rm[i] = mesh->rm[i];
  return run_bilform(mesh, dm, fa, as);
}

int32_t fegpu_bilform_lin_elastic(fegpu_mesh *mesh, fegpu_dofmap *dm, const double *C, fegpu_asm *as) {
  if (!mesh || !C) return fegpu_fail(mesh ? mesh->ctx : nullptr, FEGPU_ERR_ARG, "NULL argument");
  if (mesh->mdim != 3 || mesh->sdim != 3) return fegpu_fail(mesh->ctx, FEGPU_ERR_ARG, "bilform_lin_elastic (DeforModelRed3D) needs 3-manifold elements in 3-D");
  FormArgs fa;
  std::memset(&fa, 0, sizeof(fa));
  fa.form = FORM_ELASTIC;
  fa.ndn = 3;
  for (int i = 0; i < 36; i++) fa.coef[i] = C[i];
  fa.m = 3;
  fa.otherdim = 1.0;
  fa.use_rm = mesh->use_rm;
  for (int i = 0; i < 9; i++) fa.rm[i] = mesh->rm[i];
  return run_bilform(mesh, dm, fa, as);
}

int32_t fegpu_bilform_dot(fegpu_mesh *mesh, fegpu_dofmap *dm, const double *c, int32_t m, double otherdim, fegpu_asm *as) {
  if (!mesh || !dm || !c) return fegpu_fail(mesh ? mesh->ctx : nullptr, FEGPU_ERR_ARG, "NULL argument");
  if (mesh->mdim == 3 && m != 3) return fegpu_fail(mesh->ctx, FEGPU_ERR_MANIFOLD, "That is the only acceptable option here.");
  if (mesh->mdim == 2 && (m < 2 || m > 3)) return fegpu_fail(mesh->ctx, FEGPU_ERR_MANIFOLD, "Those are the only acceptable options here.");
  // any number of dofs per node the dof map takes (1..6): above 3 the element matrices are formed as a Kronecker product
  FormArgs fa;
  std::memset(&fa, 0, sizeof(fa));
  fa.form = FORM_DOT;
  fa.ndn = dm->ndn;
  for (int i = 0; i < dm->ndn * dm->ndn; i++) fa.coef[i] = c[i];
  fa.m = m;
  fa.otherdim = otherdim;
  return run_bilform(mesh, dm, fa, as);
}

int32_t fegpu_bilform_convection(fegpu_mesh *mesh, fegpu_dofmap *dm, const double *uvel, double rho, fegpu_asm *as) {
  if (!mesh || !dm || !uvel) return fegpu_fail(mesh ? mesh->ctx : nullptr, FEGPU_ERR_ARG, "NULL argument");
  if (mesh->sdim != mesh->mdim) return fegpu_fail(mesh->ctx, FEGPU_ERR_ARG, "bilform_convection needs sdim == manifold dimension");
  fegpu_ctx *ctx = mesh->ctx;
  DeviceGuard g(ctx->device);
  const size_t n = (size_t)mesh->nnodes * mesh->sdim;
  if (!mesh->d_uvel) CUDA_TRY(ctx, cudaMalloc((void **)&mesh->d_uvel, sizeof(double) * std::max<size_t>(n, 1)));
  if (n) CUDA_TRY(ctx, cudaMemcpyAsync(mesh->d_uvel, uvel, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // the host array may go away after return
  FormArgs fa;
  std::memset(&fa, 0, sizeof(fa));
  fa.form = FORM_CONVECTION;
  fa.ndn = 1;
  fa.coef[0] = rho;  // evaluated by the reference but not used in its integrand (FEMMBaseModule.jl:1606-1617)
  fa.m = 3;
  fa.otherdim = mesh->otherdim;
  fa.d_uvel = mesh->d_uvel;
  return run_bilform(mesh, dm, fa, as);
}

int32_t fegpu_bilform_div_grad(fegpu_mesh *mesh, fegpu_dofmap *dm, double mu, fegpu_asm *as) {
  if (!mesh || !dm) return fegpu_fail(mesh ? mesh->ctx : nullptr, FEGPU_ERR_ARG, "NULL argument");
  if (mesh->sdim != mesh->mdim) return fegpu_fail(mesh->ctx, FEGPU_ERR_ARG, "bilform_div_grad needs sdim == manifold dimension");
  if (dm->ndn != mesh->sdim) return fegpu_fail(mesh->ctx, FEGPU_ERR_ARG, "bilform_div_grad needs one dof per space dimension at every node");
  FormArgs fa;
  std::memset(&fa, 0, sizeof(fa));
  fa.form = FORM_DIV_GRAD;
  fa.ndn = dm->ndn;
  fa.coef[0] = mu;
  fa.m = 3;
  fa.otherdim = mesh->otherdim;
  return run_bilform(mesh, dm, fa, as);
}

// ---- vectors: linform_dot / distribloads and the SysvecAssembler protocol
namespace {
// dense F[row_nall] from (row, value) pairs: the pairs go through the generic sort path as an n x 1 matrix (duplicates summed
// left to right, the reference's range errors), then the column is scattered into the zeroed vector
__global__ void k_scatter_column(const int64_t *__restrict__ rowval, const double *__restrict__ nzval, int64_t nnz, double *__restrict__ F) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nnz) F[rowval[i] - 1] = nzval[i];
}
__global__ void k_fill_i64(int64_t *p, int64_t n, int64_t v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void k_emit_rows(const int32_t *__restrict__ conn, const int32_t *__restrict__ elem_list, int64_t nactive, int nne, int ndn,
                            int64_t nnodes, const int32_t *__restrict__ dof, int64_t *__restrict__ I) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int EM = nne * ndn;
  if (i >= nactive * EM) return;
  const int64_t slot = i / EM;
  const int r = (int)(i - slot * EM);
  const int64_t e = elem_list ? elem_list[slot] : slot;
  I[i] = (int64_t)dof[(int64_t)(r % ndn) * nnodes + conn[e * nne + r / ndn]] + 1;
}

// Diagonal / HRZ-lumped assemblers (AssemblyModule.jl:599-794, 943-1141): one warp per square element matrix (column-major,
// size msize[e] or EM) -> its diagonal, for HRZ scaled by ffactor = sum(mat) / trace(mat) (:1085-1090).  The lanes sum the
// entries in memory order and combine by a fixed shuffle tree, so the result is reproducible.
__global__ void k_lump(const double *__restrict__ V, int64_t nmat, int EM, const int64_t *__restrict__ moff, const int64_t *__restrict__ doff,
                       const int32_t *__restrict__ msize, int hrz, double *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (e >= nmat) return;
  const int n = msize ? msize[e] : EM;
  const double *M = V + (moff ? moff[e] : e * (int64_t)EM * EM);
  double *o = out + (doff ? doff[e] : e * (int64_t)EM);
  double ff = 1.0;
  if (hrz) {
    double em2 = 0.0, dem2 = 0.0;
    for (int i = lane; i < n * n; i += 32) em2 += M[i];
    for (int i = lane; i < n; i += 32) dem2 += M[i + (int64_t)n * i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      em2 += __shfl_xor_sync(0xffffffffu, em2, d);
      dem2 += __shfl_xor_sync(0xffffffffu, dem2, d);
    }
    ff = em2 / dem2;
  }
  for (int j = lane; j < n; j += 32) o[j] = M[j + (int64_t)n * j] * ff;
}

// (I, J) of bilform_masslike's triplets in the reference's emission order: element e contributes an ndn x EM matrix whose rows
// are the element's own ndn global rows (e-1)*ndn + 1 .. e*ndn (FEMMBaseModule.jl:1907-1908)
__global__ void k_emit_masslike_ij(const int32_t *__restrict__ conn, const int32_t *__restrict__ orig, int64_t nelem, int nne, int ndn, int64_t nnodes,
                                   const int32_t *__restrict__ dof, int64_t *__restrict__ I, int64_t *__restrict__ J) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int per = ndn * nne * ndn;
  if (t >= nelem * per) return;
  const int64_t e = t / per;
  const int rem = (int)(t - e * per);
  const int p = rem % ndn, c = rem / ndn;
  I[t] = (orig ? (int64_t)orig[e] : e) * ndn + p + 1;  // rows are numbered by the CALLER's element ids
  J[t] = (int64_t)dof[(int64_t)(c % ndn) * nnodes + conn[e * nne + c / ndn]] + 1;
}

int32_t vector_from_pairs(fegpu_asm *as, int64_t n, const int64_t *dI, const double *dV, int64_t row_nall) {
  fegpu_ctx *ctx = as->ctx;
  cudaStream_t st = ctx->stream;
  int64_t *dJ = nullptr;
  CUDA_TRY(ctx, cudaMalloc((void **)&dJ, sizeof(int64_t) * std::max<int64_t>(n, 1)));
  if (n) k_fill_i64<<<grid_for(n, 256), 256, 0, st>>>(dJ, n, 1);
  int32_t s = fe_coo_to_csc(as, n, dI, dJ, dV, row_nall, 1);
  cudaStreamSynchronize(st);
  cudaFree(dJ);
  FE_TRY(s);
  FE_TRY(fe_asm_reserve(as, &as->d_F, &as->F_cap, (size_t)std::max<int64_t>(row_nall, 1)));
  CUDA_TRY(ctx, cudaMemsetAsync(as->d_F, 0, sizeof(double) * (size_t)std::max<int64_t>(row_nall, 1), st));
  if (as->nnz) {
    k_scatter_column<<<grid_for(as->nnz, 256), 256, 0, st>>>(as->d_rowval, as->d_nzval, as->nnz, as->d_F);
    ctx->launches++;
  }
  as->have_result = false;  // the n x 1 matrix was scaffolding
  as->F_n = row_nall;
  as->have_vector = true;
  return FEGPU_OK;
}
}  // namespace

// A bilinear form into a SysmatAssemblerSparseDiag / SysmatAssemblerSparseHRZLumpingSymm: full element matrices, their (scaled)
// diagonals, and sparse(I = J = dof, V) through the sort path -- exactly the reference's makematrix! (:770-778, :1123-1131), so
// only dofs that appear in an element get a stored entry.
static int32_t run_lumped(fegpu_mesh *mesh, fegpu_dofmap *dm, const FormArgs &fa, fegpu_asm *as) {
  fegpu_ctx *ctx = mesh->ctx;
  if (mesh->partitioned) return fegpu_fail(ctx, FEGPU_ERR_ARG, "the lumped / diagonal assemblers do not take row-block partitions");
  if (dm->row_nall != dm->col_nall) return fegpu_fail(ctx, FEGPU_ERR_ARG, "Row and column info do not agree");  // :689, :1041
  DeviceGuard g(ctx->device);
  cudaStream_t st = ctx->stream;
  const int EM = mesh->nne * fa.ndn;
  const int64_t nv = mesh->nactive * (int64_t)EM;
  as->have_result = false;
  asm_set_pattern(as, nullptr);
  as->view.active = false;
  as->started = false;
  as->pattern_cached = false;
  FormArgs fa2 = fa;
  fa2.compact = false;
  FE_TRY(fe_asm_reserve(as, &as->d_V, &as->V_cap, (size_t)std::max<int64_t>(nv * EM, 1)));
  as->V_n = nv * EM;
  as->last_EM = EM;
  as->V_compact = false;
  CUDA_TRY(ctx, cudaEventRecord(as->ev[0], st));
  CUDA_TRY(ctx, cudaEventRecord(as->ev[1], st));
  CUDA_TRY(ctx, cudaEventRecord(as->ev[4], st));
  FE_TRY(fe_integrate(mesh, fa2, as->d_V));
  CUDA_TRY(ctx, cudaEventRecord(as->ev[2], st));
  CUDA_TRY(ctx, cudaEventRecord(as->ev[5], st));
  int64_t *dI = nullptr;
  double *dD = nullptr;
  CUDA_TRY(ctx, cudaMalloc((void **)&dI, sizeof(int64_t) * std::max<int64_t>(nv, 1)));
  if (cudaMalloc((void **)&dD, sizeof(double) * std::max<int64_t>(nv, 1)) != cudaSuccess) {
    cudaFree(dI);
    return fegpu_fail(ctx, FEGPU_ERR_CUDA, "out of device memory");
  }
  if (nv) {
    k_emit_rows<<<grid_for(nv, 256), 256, 0, st>>>(mesh->conn_act(), mesh->d_elem_list, mesh->nactive, mesh->nne, fa.ndn, mesh->nnodes, dm->d_dof, dI);
    k_lump<<<grid_for(mesh->nactive * 32, 256), 256, 0, st>>>(as->d_V, mesh->nactive, EM, nullptr, nullptr, nullptr, as->lump == 2 ? 1 : 0, dD);
    ctx->launches += 2;
  }
  const int32_t s = fe_coo_to_csc(as, nv, dI, dI, dD, dm->row_nall, dm->col_nall);
  cudaStreamSynchronize(st);
  cudaFree(dI);
  cudaFree(dD);
  FE_TRY(s);
  CUDA_TRY(ctx, cudaEventRecord(as->ev[3], st));
  as->ev_valid = true;
  as->have_result = true;
  return finish(ctx);
}

int32_t fegpu_bilform_masslike(fegpu_mesh *mesh, fegpu_dofmap *dm, const double *c, int32_t m, double otherdim, fegpu_asm *as) {
  if (!mesh || !dm || !c || !as) return fegpu_fail(mesh ? mesh->ctx : nullptr, FEGPU_ERR_ARG, "NULL argument");
  fegpu_ctx *ctx = mesh->ctx;
  if (dm->mesh != mesh || as->ctx != ctx || dm->ctx != ctx) return fegpu_fail(ctx, FEGPU_ERR_ARG, "handles belong to different meshes / contexts");
  if (mesh->mdim == 3 && m != 3) return fegpu_fail(ctx, FEGPU_ERR_MANIFOLD, "That is the only acceptable option here.");
  if (mesh->mdim == 2 && (m < 2 || m > 3)) return fegpu_fail(ctx, FEGPU_ERR_MANIFOLD, "Those are the only acceptable options here.");
  if (dm->ndn > 3) return fegpu_fail(ctx, FEGPU_ERR_ARG, "bilform_masslike: up to 3 dofs per node");
  if (mesh->partitioned) return fegpu_fail(ctx, FEGPU_ERR_ARG, "bilform_masslike numbers its rows by element: no row-block partitions");
  if (as->symmetric || as->lump) return fegpu_fail(ctx, FEGPU_ERR_ARG, "bilform_masslike assembles a rectangular matrix: use the plain sparse assembler");
  DeviceGuard g(ctx->device);
  cudaStream_t st = ctx->stream;
  FormArgs fa;
  std::memset(&fa, 0, sizeof(fa));
  fa.form = FORM_MASSLIKE;
  fa.ndn = dm->ndn;
  for (int i = 0; i < dm->ndn * dm->ndn; i++) fa.coef[i] = c[i];
  fa.m = m;
  fa.otherdim = otherdim;
  const int per = dm->ndn * mesh->nne * dm->ndn;
  const int64_t n = mesh->nelem * (int64_t)per;
  as->have_result = false;
  asm_set_pattern(as, nullptr);
  as->view.active = false;
  as->started = false;
  as->pattern_cached = false;
  FE_TRY(fe_asm_reserve(as, &as->d_V, &as->V_cap, (size_t)std::max<int64_t>(n, 1)));
  as->V_n = 0;
  as->V_compact = false;
  for (int k : {0, 1, 4}) CUDA_TRY(ctx, cudaEventRecord(as->ev[k], st));
  FE_TRY(fe_integrate(mesh, fa, as->d_V));
  for (int k : {2, 5}) CUDA_TRY(ctx, cudaEventRecord(as->ev[k], st));
  int64_t *dI = nullptr, *dJ = nullptr;
  CUDA_TRY(ctx, cudaMalloc((void **)&dI, sizeof(int64_t) * std::max<int64_t>(n, 1)));
  if (cudaMalloc((void **)&dJ, sizeof(int64_t) * std::max<int64_t>(n, 1)) != cudaSuccess) {
    cudaFree(dI);
    return fegpu_fail(ctx, FEGPU_ERR_CUDA, "out of device memory");
  }
  if (n) {
    k_emit_masslike_ij<<<grid_for(n, 256), 256, 0, st>>>(mesh->d_conn, mesh->d_orig, mesh->nelem, mesh->nne, dm->ndn, mesh->nnodes, dm->d_dof, dI, dJ);
    ctx->launches++;
  }
  const int32_t s = fe_coo_to_csc(as, n, dI, dJ, as->d_V, mesh->nelem * dm->ndn, dm->col_nall);
  cudaStreamSynchronize(st);
  cudaFree(dI);
  cudaFree(dJ);
  FE_TRY(s);
  CUDA_TRY(ctx, cudaEventRecord(as->ev[3], st));
  as->ev_valid = true;
  as->have_result = true;
  return finish(ctx);
}

int32_t fegpu_linform_dot(fegpu_mesh *mesh, fegpu_dofmap *dm, const double *force, int32_t m, double otherdim, fegpu_asm *as) {
  if (!mesh || !dm || !force || !as) return fegpu_fail(mesh ? mesh->ctx : nullptr, FEGPU_ERR_ARG, "NULL argument");
  fegpu_ctx *ctx = mesh->ctx;
  if (dm->mesh != mesh || as->ctx != ctx || dm->ctx != ctx) return fegpu_fail(ctx, FEGPU_ERR_ARG, "handles belong to different meshes / contexts");
  if (mesh->mdim == 3 && m != 3) return fegpu_fail(ctx, FEGPU_ERR_MANIFOLD, "That is the only acceptable option here.");
  if (mesh->mdim == 2 && (m < 2 || m > 3)) return fegpu_fail(ctx, FEGPU_ERR_MANIFOLD, "Those are the only acceptable options here.");
  if (dm->ndn > 3) return fegpu_fail(ctx, FEGPU_ERR_ARG, "linform_dot: up to 3 dofs per node");
  DeviceGuard g(ctx->device);
  cudaStream_t st = ctx->stream;
  as->have_vector = false;
  FormArgs fa;
  std::memset(&fa, 0, sizeof(fa));
  fa.form = FORM_LINDOT;
  fa.ndn = dm->ndn;
  for (int i = 0; i < dm->ndn; i++) fa.coef[i] = force[i];
  fa.m = m;
  fa.otherdim = otherdim;
  const int EM = mesh->nne * dm->ndn;
  const int64_t nv = mesh->nactive * (int64_t)EM;
  FE_TRY(fe_asm_reserve(as, &as->d_V, &as->V_cap, (size_t)std::max<int64_t>(nv, 1)));
  as->V_n = 0;  // the element-matrix buffer now holds element vectors
  as->have_result = false;
  FE_TRY(fe_integrate(mesh, fa, as->d_V));
  bool fast = fe_pattern_usable(dm);
  if (fast && (!dm->pat || dm->pat_topo_version != mesh->topo_version)) {
    FE_TRY(fe_pattern_build(dm));
    fast = fe_pattern_usable(dm) && dm->pat;
  }
  if (fast) {
    FE_TRY(fe_asm_reserve(as, &as->d_F, &as->F_cap, (size_t)std::max<int64_t>(dm->row_nall, 1)));
    FE_TRY(fe_vec_gather(dm, as->d_V, as->d_F));
    as->F_n = dm->row_nall;
    as->have_vector = true;
  } else {
    if (mesh->partitioned) return fegpu_fail(ctx, FEGPU_ERR_ARG, "row-block partitioning needs an injective dof map and non-degenerate elements");
    int64_t *dI = nullptr;
    CUDA_TRY(ctx, cudaMalloc((void **)&dI, sizeof(int64_t) * std::max<int64_t>(nv, 1)));
    if (nv) k_emit_rows<<<grid_for(nv, 256), 256, 0, st>>>(mesh->conn_act(), mesh->d_elem_list, mesh->nactive, mesh->nne, dm->ndn, mesh->nnodes, dm->d_dof, dI);
    const int32_t s = vector_from_pairs(as, nv, dI, as->d_V, dm->row_nall);
    cudaStreamSynchronize(st);
    cudaFree(dI);
    FE_TRY(s);
  }
  return finish(ctx);
}

int32_t fegpu_vec_startassembly(fegpu_asm *as, int64_t row_nall) {
  if (!as) return fegpu_fail(nullptr, FEGPU_ERR_ARG, "NULL assembler");
  if (row_nall < 0 || row_nall >= INT32_MAX) return fegpu_fail(as->ctx, FEGPU_ERR_ARG, "vector length outside int32 range");
  as->hvI.clear();
  as->hvV.clear();
  as->v_row_nall = row_nall;
  as->vec_started = true;
  as->have_vector = false;
  return FEGPU_OK;
}

int32_t fegpu_vec_assemble(fegpu_asm *as, const double *vec, const int64_t *dofnums, int64_t n) {
  if (!as || (n > 0 && (!vec || !dofnums))) return fegpu_fail(as ? as->ctx : nullptr, FEGPU_ERR_ARG, "NULL argument");
  if (!as->vec_started) return fegpu_fail(as->ctx, FEGPU_ERR_STATE, "assemble! before startassembly!");
  for (int64_t i = 0; i < n; i++) {  // AssemblyModule.jl:899-906
    const int64_t gi = dofnums[i];
    if (gi < 1) return fegpu_fail(as->ctx, FEGPU_ERR_ROW_LT1, "Row degree of freedom < 1");
    if (gi > as->v_row_nall) return fegpu_fail(as->ctx, FEGPU_ERR_ROW_GT, "Row degree of freedom > size");
    as->hvI.push_back(gi);
    as->hvV.push_back(vec[i]);
  }
  return FEGPU_OK;
}

int32_t fegpu_makevector(fegpu_asm *as) {
  if (!as) return fegpu_fail(nullptr, FEGPU_ERR_ARG, "NULL assembler");
  fegpu_ctx *ctx = as->ctx;
  if (!as->vec_started) return fegpu_fail(ctx, FEGPU_ERR_STATE, "makevector! without startassembly!");
  DeviceGuard g(ctx->device);
  cudaStream_t st = ctx->stream;
  const int64_t n = (int64_t)as->hvV.size();
  int64_t *dI = nullptr;
  CUDA_TRY(ctx, cudaMalloc((void **)&dI, sizeof(int64_t) * std::max<int64_t>(n, 1)));
  int32_t s = fe_asm_reserve(as, &as->d_V, &as->V_cap, (size_t)std::max<int64_t>(n, 1));
  if (s == FEGPU_OK && n) {
    if (cudaMemcpyAsync(dI, as->hvI.data(), sizeof(int64_t) * n, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(as->d_V, as->hvV.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st) != cudaSuccess)
      s = fegpu_fail(ctx, FEGPU_ERR_CUDA, "upload of the staged vector entries failed");
  }
  as->V_n = 0;
  if (s == FEGPU_OK) s = vector_from_pairs(as, n, dI, as->d_V, as->v_row_nall);
  cudaStreamSynchronize(st);
  cudaFree(dI);
  FE_TRY(s);
  return finish(ctx);
}

int32_t fegpu_makevector_size(fegpu_asm *as, int64_t *n) {
  if (!as || !n) return fegpu_fail(as ? as->ctx : nullptr, FEGPU_ERR_ARG, "NULL argument");
  if (!as->have_vector) return fegpu_fail(as->ctx, FEGPU_ERR_STATE, "no assembled vector");
  *n = as->F_n;
  return FEGPU_OK;
}

int32_t fegpu_makevector_copy(fegpu_asm *as, double *F) {
  if (!as || !F) return fegpu_fail(as ? as->ctx : nullptr, FEGPU_ERR_ARG, "NULL argument");
  fegpu_ctx *ctx = as->ctx;
  if (!as->have_vector) return fegpu_fail(ctx, FEGPU_ERR_STATE, "no assembled vector");
  DeviceGuard g(ctx->device);
  if (as->F_n) CUDA_TRY(ctx, cudaMemcpyAsync(F, as->d_F, sizeof(double) * (size_t)as->F_n, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return FEGPU_OK;
}

// ---- generic protocol
int32_t fegpu_startassembly(fegpu_asm *as, int64_t nr, int64_t nc, int64_t nmats, int64_t row_nall, int64_t col_nall) {
  if (!as) return fegpu_fail(nullptr, FEGPU_ERR_ARG, "NULL assembler");
  if (nr < 0 || nc < 0 || nmats < 0 || row_nall < 0 || col_nall < 0) return fegpu_fail(as->ctx, FEGPU_ERR_ARG, "negative size");
  // like the reference (AssemblyModule.jl:221-229) sizes are only taken when no assembly is in flight
  if (as->lump && nr != nc) return fegpu_fail(as->ctx, FEGPU_ERR_ARG, "Diagonal sparse matrix is assumed to be assembled from square matrices");
  if (as->lump && row_nall != col_nall) return fegpu_fail(as->ctx, FEGPU_ERR_ARG, "Row and column info do not agree");
  if (!as->started) {
    as->hI.clear(); as->hJ.clear(); as->hV.clear(); as->hN.clear();
    const size_t expect = (size_t)(nr * nc * nmats);  // expectedntriples, :54-59
    as->hI.reserve(expect); as->hJ.reserve(expect); as->hV.reserve(expect);
    as->g_row_nall = row_nall;
    as->g_col_nall = col_nall;
    as->started = true;
  }
  return FEGPU_OK;
}

int32_t fegpu_assemble(fegpu_asm *as, const double *mat, const int64_t *dr, int64_t nrows, const int64_t *dc, int64_t ncols) {
  if (!as || !mat || !dr || !dc) return fegpu_fail(as ? as->ctx : nullptr, FEGPU_ERR_ARG, "NULL argument");
  if (!as->started) return fegpu_fail(as->ctx, FEGPU_ERR_STATE, "assemble! before startassembly!");
  if (as->symmetric && nrows != ncols) return fegpu_fail(as->ctx, FEGPU_ERR_MATSIZE, "Size mismatch");  // AssemblyModule.jl:510
  if (as->lump) {
    // diagonal / HRZ: the whole square matrix is staged (the scaling factor needs every entry), its column dofs, its size
    if (nrows != ncols) return fegpu_fail(as->ctx, FEGPU_ERR_MATSIZE, "Size mismatch");  // :724, :1077
    for (int64_t j = 0; j < ncols; j++) {
      if (dc[j] < 1) return fegpu_fail(as->ctx, FEGPU_ERR_COL_LT1, "Column degree of freedom < 1");
      if (dc[j] > as->g_col_nall) return fegpu_fail(as->ctx, FEGPU_ERR_COL_GT, "Column degree of freedom > size");
    }
    as->hV.insert(as->hV.end(), mat, mat + nrows * ncols);
    as->hJ.insert(as->hJ.end(), dc, dc + ncols);
    as->hN.push_back((int32_t)ncols);
    return FEGPU_OK;
  }
  for (int64_t j = 0; j < ncols; j++) {
    const int64_t dj = dc[j];
    if (dj < 1) return fegpu_fail(as->ctx, FEGPU_ERR_COL_LT1, "Column degree of freedom < 1");
    if (dj > as->g_col_nall) return fegpu_fail(as->ctx, FEGPU_ERR_COL_GT, "Column degree of freedom > size");
    for (int64_t i = as->symmetric ? j : 0; i < nrows; i++) {  // symmetric: lower triangle only, :521
      const int64_t di = dr[i];
      if (di < 1) return fegpu_fail(as->ctx, FEGPU_ERR_ROW_LT1, "Row degree of freedom < 1");
      if (di > as->g_row_nall) return fegpu_fail(as->ctx, FEGPU_ERR_ROW_GT, "Row degree of freedom > size");
      as->hV.push_back(mat[i + nrows * j]);
      as->hI.push_back(di);
      as->hJ.push_back(dj);
    }
  }
  return FEGPU_OK;
}

int32_t fegpu_triplets_append(fegpu_asm *as, int64_t n, const int64_t *I, const int64_t *J, const double *V) {
  if (!as || (n > 0 && (!I || !J || !V))) return fegpu_fail(as ? as->ctx : nullptr, FEGPU_ERR_ARG, "NULL argument");
  if (!as->started) return fegpu_fail(as->ctx, FEGPU_ERR_STATE, "triplets before startassembly!");
  as->hI.insert(as->hI.end(), I, I + n);
  as->hJ.insert(as->hJ.end(), J, J + n);
  as->hV.insert(as->hV.end(), V, V + n);
  return FEGPU_OK;
}

int32_t fegpu_makematrix(fegpu_asm *as) {
  if (!as) return fegpu_fail(nullptr, FEGPU_ERR_ARG, "NULL assembler");
  fegpu_ctx *ctx = as->ctx;
  if (!as->started) return fegpu_fail(ctx, FEGPU_ERR_STATE, "makematrix! without startassembly!");
  DeviceGuard g(ctx->device);
  cudaStream_t st = ctx->stream;
  if (as->lump) {
    // staged square matrices -> their (HRZ-scaled) diagonals on the device -> sparse(I = J = dof, V)
    const int64_t nmat = (int64_t)as->hN.size(), nval = (int64_t)as->hV.size(), nd = (int64_t)as->hJ.size();
    std::vector<int64_t> moff(nmat + 1, 0), doff(nmat + 1, 0);
    for (int64_t k = 0; k < nmat; k++) {
      moff[k + 1] = moff[k] + (int64_t)as->hN[k] * as->hN[k];
      doff[k + 1] = doff[k] + as->hN[k];
    }
    int64_t *dJ = nullptr, *dmo = nullptr, *ddo = nullptr;
    int32_t *dsz = nullptr;
    double *dD = nullptr;
    auto cleanup = [&]() { cudaFree(dJ); cudaFree(dmo); cudaFree(ddo); cudaFree(dsz); cudaFree(dD); };
    int32_t s = fe_asm_reserve(as, &as->d_V, &as->V_cap, (size_t)std::max<int64_t>(nval, 1));
    cudaError_t e = cudaSuccess;
    if (s == FEGPU_OK) {
      if ((e = cudaMalloc((void **)&dJ, sizeof(int64_t) * std::max<int64_t>(nd, 1))) == cudaSuccess &&
          (e = cudaMalloc((void **)&dmo, sizeof(int64_t) * (nmat + 1))) == cudaSuccess &&
          (e = cudaMalloc((void **)&ddo, sizeof(int64_t) * (nmat + 1))) == cudaSuccess &&
          (e = cudaMalloc((void **)&dsz, sizeof(int32_t) * std::max<int64_t>(nmat, 1))) == cudaSuccess &&
          (e = cudaMalloc((void **)&dD, sizeof(double) * std::max<int64_t>(nd, 1))) == cudaSuccess) {
        if (nval) cudaMemcpyAsync(as->d_V, as->hV.data(), sizeof(double) * nval, cudaMemcpyHostToDevice, st);
        if (nd) cudaMemcpyAsync(dJ, as->hJ.data(), sizeof(int64_t) * nd, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(dmo, moff.data(), sizeof(int64_t) * (nmat + 1), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(ddo, doff.data(), sizeof(int64_t) * (nmat + 1), cudaMemcpyHostToDevice, st);
        if (nmat) cudaMemcpyAsync(dsz, as->hN.data(), sizeof(int32_t) * nmat, cudaMemcpyHostToDevice, st);
        for (int k : {0, 1, 4, 2, 5}) cudaEventRecord(as->ev[k], st);
        if (nmat) {
          k_lump<<<grid_for(nmat * 32, 256), 256, 0, st>>>(as->d_V, nmat, 0, dmo, ddo, dsz, as->lump == 2 ? 1 : 0, dD);
          ctx->launches++;
        }
        s = fe_coo_to_csc(as, nd, dJ, dJ, dD, as->g_row_nall, as->g_col_nall);
        cudaStreamSynchronize(st);  // the host offset vectors go out of scope
      } else {
        s = fegpu_fail(ctx, FEGPU_ERR_CUDA, cudaGetErrorString(e));
      }
    }
    cleanup();
    FE_TRY(s);
    CUDA_TRY(ctx, cudaEventRecord(as->ev[3], st));
    as->ev_valid = true;
    as->have_result = true;
    asm_set_pattern(as, nullptr);
    as->pattern_cached = false;
    as->view.active = false;
    as->V_compact = false;
    as->started = false;
    as->V_n = 0;
    return finish(ctx);
  }
  if (as->symmetric) {
    // S + transpose(S) with the doubled diagonal halved (AssemblyModule.jl:576-579): the mirrored copy of every
    // off-diagonal triplet is appended; the diagonal stays single, which equals (2 S_jj) * 0.5 exactly
    const size_t n0 = as->hV.size();
    for (size_t k = 0; k < n0; k++)
      if (as->hI[k] != as->hJ[k]) {
        as->hI.push_back(as->hJ[k]);
        as->hJ.push_back(as->hI[k]);
        as->hV.push_back(as->hV[k]);
      }
  }
  const int64_t n = (int64_t)as->hV.size();
  int64_t *dI = nullptr, *dJ = nullptr;
  cudaError_t e;
  auto cleanup = [&]() { cudaFree(dI); cudaFree(dJ); };
#define GT(expr) if ((e = (expr)) != cudaSuccess) { cleanup(); return fegpu_fail(ctx, FEGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e)); }
  GT(cudaMalloc((void **)&dI, sizeof(int64_t) * std::max<int64_t>(n, 1)));
  GT(cudaMalloc((void **)&dJ, sizeof(int64_t) * std::max<int64_t>(n, 1)));
  int32_t s = fe_asm_reserve(as, &as->d_V, &as->V_cap, (size_t)std::max<int64_t>(n, 1));
  if (s != FEGPU_OK) { cleanup(); return s; }
  GT(cudaEventRecord(as->ev[0], st));
  GT(cudaEventRecord(as->ev[1], st));
  GT(cudaEventRecord(as->ev[4], st));
  if (n) {
    GT(cudaMemcpyAsync(dI, as->hI.data(), sizeof(int64_t) * n, cudaMemcpyHostToDevice, st));
    GT(cudaMemcpyAsync(dJ, as->hJ.data(), sizeof(int64_t) * n, cudaMemcpyHostToDevice, st));
    GT(cudaMemcpyAsync(as->d_V, as->hV.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
  }
  GT(cudaEventRecord(as->ev[2], st));
  GT(cudaEventRecord(as->ev[5], st));
  s = fe_coo_to_csc(as, n, dI, dJ, as->d_V, as->g_row_nall, as->g_col_nall);
  cleanup();
  FE_TRY(s);
  CUDA_TRY(ctx, cudaEventRecord(as->ev[3], st));
#undef GT
  as->ev_valid = true;
  as->have_result = true;
  asm_set_pattern(as, nullptr);
  as->pattern_cached = false;
  as->view.active = false;
  as->V_compact = false;
  as->started = false;  // "_buffer_pointer = 1": ready for the next startassembly!  (AssemblyModule.jl:327)
  as->V_n = 0;
  if (as->symmetric) FE_TRY(fe_csc_view(as, 1, as->nrows, 1, as->ncols, true));  // the sparse `+` keeps only non-zero sums
  return finish(ctx);
}

// ---- results
int32_t fegpu_makematrix_sizes(fegpu_asm *as, int64_t *nrows, int64_t *ncols, int64_t *nnz) {
  if (!as) return fegpu_fail(nullptr, FEGPU_ERR_ARG, "NULL assembler");
  if (!as->have_result) return fegpu_fail(as->ctx, FEGPU_ERR_STATE, "no assembled matrix");
  if (nrows) *nrows = as->r_nrows();
  if (ncols) *ncols = as->r_ncols();
  if (nnz) *nnz = as->r_nnz();
  return FEGPU_OK;
}

int32_t fegpu_makematrix_copy(fegpu_asm *as, int64_t *colptr, int64_t *rowval, double *nzval) {
  if (!as) return fegpu_fail(nullptr, FEGPU_ERR_ARG, "NULL assembler");
  fegpu_ctx *ctx = as->ctx;
  if (!as->have_result) return fegpu_fail(ctx, FEGPU_ERR_STATE, "no assembled matrix");
  DeviceGuard g(ctx->device);
  cudaStream_t st = ctx->stream;
  // large results go through the transport of fegpu_transfer.cu (int32 row indices on the link, host-thread widening)
  if (as->r_nnz() >= ((int64_t)1 << 20) && as->nrows < INT32_MAX) return fe_copy_result(as, colptr, rowval, nzval);
  if (colptr) CUDA_TRY(ctx, cudaMemcpyAsync(colptr, as->r_colptr(), sizeof(int64_t) * (as->r_ncols() + 1), cudaMemcpyDeviceToHost, st));
  if (rowval && as->r_nnz()) CUDA_TRY(ctx, cudaMemcpyAsync(rowval, as->r_rowval(), sizeof(int64_t) * as->r_nnz(), cudaMemcpyDeviceToHost, st));
  if (nzval && as->r_nnz()) CUDA_TRY(ctx, cudaMemcpyAsync(nzval, as->r_nzval(), sizeof(double) * as->r_nnz(), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(ctx, cudaStreamSynchronize(st));
  return FEGPU_OK;
}

int32_t fegpu_makematrix_view(fegpu_asm *as, int64_t row_first, int64_t row_last, int64_t col_first, int64_t col_last, int32_t drop_exact_zeros) {
  if (!as) return fegpu_fail(nullptr, FEGPU_ERR_ARG, "NULL assembler");
  if (!as->have_result) return fegpu_fail(as->ctx, FEGPU_ERR_STATE, "no assembled matrix");
  DeviceGuard g(as->ctx->device);
  return fe_csc_view(as, row_first, row_last, col_first, col_last, drop_exact_zeros != 0);
}

int32_t fegpu_makematrix_copy_values(fegpu_asm *as, double *nzval) { return fegpu_makematrix_copy(as, nullptr, nullptr, nzval); }

int32_t fegpu_makematrix_device(fegpu_asm *as, const int64_t **c, const int64_t **r, const double **v) {
  if (!as) return fegpu_fail(nullptr, FEGPU_ERR_ARG, "NULL assembler");
  if (!as->have_result) return fegpu_fail(as->ctx, FEGPU_ERR_STATE, "no assembled matrix");
  if (c) *c = as->r_colptr();
  if (r) *r = as->r_rowval();
  if (v) *v = as->r_nzval();
  return FEGPU_OK;
}

int32_t fegpu_coo_copy(fegpu_asm *as, fegpu_mesh *mesh, fegpu_dofmap *dm, int64_t *I, int64_t *J, double *V) {
  if (!as || !mesh || !dm) return fegpu_fail(nullptr, FEGPU_ERR_ARG, "NULL argument");
  fegpu_ctx *ctx = as->ctx;
  const int EM = mesh->nne * dm->ndn;
  const int64_t n = mesh->nactive * (int64_t)EM * EM;
  if (as->V_n != n || as->last_EM != EM) return fegpu_fail(ctx, FEGPU_ERR_STATE, "assembler does not hold this mesh's element matrices");
  DeviceGuard g(ctx->device);
  cudaStream_t st = ctx->stream;
  if (n == 0) return FEGPU_OK;
  // the element values sit in internal (slot) order; the reference emits element after element in ITS order
  int32_t *d_perm = nullptr;
  CUDA_TRY(ctx, cudaMalloc((void **)&d_perm, sizeof(int32_t) * (size_t)mesh->nactive));
  int32_t ps = fe_emission_order(mesh, d_perm);
  if (ps != FEGPU_OK) { cudaFree(d_perm); return ps; }
  if (I || J) {
    int64_t *dI = nullptr, *dJ = nullptr;
    cudaError_t e = cudaMalloc((void **)&dI, sizeof(int64_t) * n);
    if (e == cudaSuccess) e = cudaMalloc((void **)&dJ, sizeof(int64_t) * n);
    if (e != cudaSuccess) { cudaFree(dI); cudaFree(d_perm); return fegpu_fail(ctx, FEGPU_ERR_CUDA, cudaGetErrorString(e)); }
    int32_t s = fe_emit_ij(dm, dI, dJ, d_perm);
    if (s == FEGPU_OK && I && cudaMemcpyAsync(I, dI, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, st) != cudaSuccess) s = FEGPU_ERR_CUDA;
    if (s == FEGPU_OK && J && cudaMemcpyAsync(J, dJ, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, st) != cudaSuccess) s = FEGPU_ERR_CUDA;
    cudaStreamSynchronize(st);
    cudaFree(dI);
    cudaFree(dJ);
    if (s != FEGPU_OK) { cudaFree(d_perm); return s; }
  }
  if (V) {
    double *dfull = nullptr;
    cudaError_t e = cudaMalloc((void **)&dfull, sizeof(double) * n);
    if (e != cudaSuccess) { cudaFree(d_perm); return fegpu_fail(ctx, FEGPU_ERR_CUDA, cudaGetErrorString(e)); }
    int32_t s = FEGPU_OK;
    if (as->V_compact) {  // the fast path stored the compact symmetric layout: expand to full matrices on the way
      s = fe_expand_compact(ctx, as->d_V, dfull, mesh->nactive, mesh->nne, dm->ndn, d_perm, as->V_planes ? as->V_stride : 0);
    } else {
      k_permute_records<<<grid_for(n, 256), 256, 0, st>>>(as->d_V, dfull, mesh->nactive, (int64_t)EM * EM, d_perm, as->V_planes ? as->V_stride : 0,
                                                          mesh->nne, dm->ndn);
      ctx->launches++;
    }
    if (s == FEGPU_OK && cudaMemcpyAsync(V, dfull, sizeof(double) * n, cudaMemcpyDeviceToHost, st) != cudaSuccess) s = FEGPU_ERR_CUDA;
    cudaStreamSynchronize(st);
    cudaFree(dfull);
    if (s != FEGPU_OK) { cudaFree(d_perm); return s; }
  }
  CUDA_TRY(ctx, cudaStreamSynchronize(st));
  cudaFree(d_perm);
  return FEGPU_OK;
}

int32_t fegpu_last_timings(fegpu_asm *as, double ms[4]) {
  if (!as || !ms) return fegpu_fail(nullptr, FEGPU_ERR_ARG, "NULL argument");
  if (!as->ev_valid) return fegpu_fail(as->ctx, FEGPU_ERR_STATE, "no timed call yet");
  DeviceGuard g(as->ctx->device);
  CUDA_TRY(as->ctx, cudaEventSynchronize(as->ev[3]));
  float t;
  // reported order: integration (ev4 -> ev2, possibly on the second stream and concurrent with the symbolic phase),
  // symbolic (ev0 -> ev1), numeric (ev5 -> ev3)
  static const int first[3] = {4, 0, 5}, last[3] = {2, 1, 3};
  for (int i = 0; i < 3; i++) {
    CUDA_TRY(as->ctx, cudaEventElapsedTime(&t, as->ev[first[i]], as->ev[last[i]]));
    ms[i] = t;
  }
  CUDA_TRY(as->ctx, cudaEventElapsedTime(&t, as->ev[0], as->ev[3]));
  ms[3] = t;
  return FEGPU_OK;
}

int32_t fegpu_pattern_was_cached(fegpu_asm *as) { return (as && as->pattern_cached) ? 1 : 0; }

}  // extern "C"
