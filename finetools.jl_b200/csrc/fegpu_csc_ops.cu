// Device-side views of the assembled CSC: sub-blocks and zero-dropping.
//
//  * Block extraction replaces matrix_blocked_ff/fd/df/dd (MatrixUtilityModule.jl:675-793: A[1:nf, 1:nf], A[1:nf, nf+1:end], ...)
//    and SysmatAssemblerFFBlock's makematrix! (AssemblyModule.jl:1149-1231): only the rows/columns of the requested range are
//    kept (stored zeros included, as Julia's range indexing of a SparseMatrixCSC does), row indices are rebased to 1.
//  * Zero-dropping reproduces what SysmatAssemblerSparseSymm's makematrix! ends up with (AssemblyModule.jl:551-583): its
//    `S + transpose(S)` goes through SparseArrays' zero-preserving map, which stores only non-zero results.
// Three kernels: per-column count of the entries that survive, scan, order-preserving compaction.  The full result of the
// assembly stays untouched (several blocks can be cut from one assembly); the view owns its own colptr/rowval/nzval.
#include "fegpu_internal.h"

namespace {

constexpr int FL = 8;  // lanes per column

struct FilterParams {
  const int64_t *colptr, *rowval;  // source, 1-based
  const double *nzval;
  int64_t r0, r1, c0;              // kept rows r0..r1 (1-based, inclusive), first kept column
  int64_t ncols_out;
  int drop_zeros;
};

__device__ __forceinline__ bool keep(const FilterParams &F, int64_t row, double v) {
  return row >= F.r0 && row <= F.r1 && !(F.drop_zeros && v == 0.0);
}

__global__ void __launch_bounds__(256) k_filter_count(const FilterParams F, int64_t *__restrict__ count) {
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / FL;
  const int gl = threadIdx.x % FL;
  if (c >= F.ncols_out) return;
  const int64_t b = F.colptr[F.c0 - 1 + c] - 1, e = F.colptr[F.c0 + c] - 1;
  int n = 0;
  for (int64_t k = b + gl; k < e; k += FL) n += keep(F, F.rowval[k], F.nzval[k]) ? 1 : 0;
#pragma unroll
  for (int d = 1; d < FL; d <<= 1) n += __shfl_xor_sync(0xffffffffu, n, d);
  if (gl == 0) count[c] = n;
}

__global__ void __launch_bounds__(256) k_filter_write(const FilterParams F, const int64_t *__restrict__ colptr_out, int64_t *__restrict__ rowval_out,
                                                      double *__restrict__ nzval_out) {
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / FL;
  const int lane = threadIdx.x & 31, gl = threadIdx.x % FL;
  const unsigned gmask = ((1u << FL) - 1u) << (lane - gl);  // lanes of this column's group
  const bool live = c < F.ncols_out;
  const int64_t b = live ? F.colptr[F.c0 - 1 + c] - 1 : 0, e = live ? F.colptr[F.c0 + c] - 1 : 0;
  int64_t out = live ? colptr_out[c] - 1 : 0;
  // all groups of a warp iterate together so the ballots stay converged
  int64_t len = e - b;
#pragma unroll
  for (int d = FL; d < 32; d <<= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, d));
  for (int64_t k0 = 0; k0 < len; k0 += FL) {
    const int64_t k = b + k0 + gl;
    int64_t row = 0;
    double v = 0.0;
    bool kp = false;
    if (k < e) {
      row = F.rowval[k];
      v = F.nzval[k];
      kp = keep(F, row, v);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, kp) & gmask;
    if (kp) {
      const int64_t pos = out + __popc(bal & ((1u << lane) - 1u));
      rowval_out[pos] = row - F.r0 + 1;
      nzval_out[pos] = v;
    }
    out += __popc(bal);
  }
}

}  // namespace

int32_t fe_csc_view(fegpu_asm *as, int64_t r0, int64_t r1, int64_t c0, int64_t c1, bool drop_zeros) {
  fegpu_ctx *ctx = as->ctx;
  cudaStream_t st = ctx->stream;
  if (r0 < 1 || c0 < 1 || r1 > as->nrows || c1 > as->ncols || r1 < r0 - 1 || c1 < c0 - 1)
    return fegpu_fail(ctx, FEGPU_ERR_ARG, "block range outside the matrix");
  if (r0 == 1 && c0 == 1 && r1 == as->nrows && c1 == as->ncols && !drop_zeros) {
    as->view.active = false;
    return FEGPU_OK;
  }
  const int64_t nco = c1 - c0 + 1, nro = r1 - r0 + 1;
  size_t capb = as->view.colptr_cap * sizeof(int64_t);
  FE_TRY(fe_reserve_bytes(ctx, (void **)&as->view.own_colptr, &capb, sizeof(int64_t) * (size_t)(nco + 1)));
  as->view.colptr_cap = capb / sizeof(int64_t);
  CUDA_TRY(ctx, cudaMemsetAsync(as->view.own_colptr, 0, sizeof(int64_t) * (size_t)(nco + 1), st));
  FilterParams F{as->d_colptr, as->d_rowval, as->d_nzval, r0, r1, c0, nco, drop_zeros ? 1 : 0};
  int64_t tot = 1;
  if (nco > 0) {
    k_filter_count<<<grid_for(nco * FL, 256), 256, 0, st>>>(F, as->view.own_colptr);
    ctx->launches++;
  }
  FE_TRY(fe_exclusive_scan_i64(ctx, as->view.own_colptr, as->view.own_colptr, nco, 1, true, &tot));
  const int64_t nnz = tot - 1;
  capb = as->view.rowval_cap * sizeof(int64_t);
  FE_TRY(fe_reserve_bytes(ctx, (void **)&as->view.own_rowval, &capb, sizeof(int64_t) * (size_t)std::max<int64_t>(nnz, 1)));
  as->view.rowval_cap = capb / sizeof(int64_t);
  capb = as->view.nzval_cap * sizeof(double);
  FE_TRY(fe_reserve_bytes(ctx, (void **)&as->view.own_nzval, &capb, sizeof(double) * (size_t)std::max<int64_t>(nnz, 1)));
  as->view.nzval_cap = capb / sizeof(double);
  if (nco > 0 && nnz > 0) {
    k_filter_write<<<grid_for(nco * FL, 256), 256, 0, st>>>(F, as->view.own_colptr, as->view.own_rowval, as->view.own_nzval);
    ctx->launches++;
  }
  CUDA_TRY(ctx, cudaGetLastError());
  as->view.nrows = nro;
  as->view.ncols = nco;
  as->view.nnz = nnz;
  as->view.active = true;
  if (!ctx->async) CUDA_TRY(ctx, cudaStreamSynchronize(st));
  return FEGPU_OK;
}

// ------------------------------------------------------------------------------------------------ column stencils (transport codec)
// Matrices assembled on meshes repeat a few column shapes: the rows of column j are j + (a short list of offsets), and the same
// list serves most columns (one list for the interior of a structured block, a few dozen for its boundary).  For the result
// transport (fegpu_transfer.cu) the row indices are therefore described by ONE id per column + a dictionary of offset lists, 4 bytes
// per column instead of 4 bytes per non-zero, and the host threads rebuild rowval.  This is a codec of what the device produced:
// the lists are read off the device's rowval and every column is verified against its dictionary entry before anything is shipped;
// a matrix with too many distinct shapes (unstructured meshes) or a hash collision simply takes the int32 path.
namespace {

constexpr int ST_LANES = 8;  // lanes per column
constexpr unsigned long long ST_EMPTY = ~0ull;

__device__ __forceinline__ unsigned long long st_mix(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

// flags: [0] a column longer than maxlen, [1] table too full, [2] a column differs from its dictionary entry (hash collision)
__global__ void __launch_bounds__(256) k_stencil_insert(int64_t ncols, const int64_t *__restrict__ colptr, const int64_t *__restrict__ rowval, int maxlen,
                                                        int cap_mask, unsigned long long *keys, int32_t *rep, uint32_t *__restrict__ ids, int *flags) {
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / ST_LANES;
  const int gl = threadIdx.x % ST_LANES;
  const bool live = c < ncols;  // no early exit: every lane of the warp takes part in the shuffles
  const int64_t b = live ? colptr[c] - 1 : 0, e = live ? colptr[c + 1] - 1 : 0;
  const int64_t len = e - b;
  unsigned long long h = 0;
  for (int64_t k = b + gl; k < e; k += ST_LANES) {
    const unsigned long long off = (unsigned long long)(rowval[k] - 1 - c);  // row - column, 0-based (wraps for rows above the diagonal)
    h += st_mix(off * 0x9e3779b97f4a7c15ull + (unsigned long long)(k - b) * 0xd1b54a32d192ed03ull + 1ull);
  }
#pragma unroll
  for (int d = 1; d < ST_LANES; d <<= 1) h += __shfl_xor_sync(0xffffffffu, h, d);
  if (!live || gl != 0) return;
  if (*reinterpret_cast<volatile int *>(flags + 1)) { ids[c] = 0; return; }  // the table is already known to be too crowded: no dictionary
  if (len > maxlen) { flags[0] = 1; ids[c] = 0; return; }
  h = st_mix(h + (unsigned long long)len * 0x2545f4914f6cdd1dull);
  if (h == ST_EMPTY) h = 0;
  int slot = (int)(h & (unsigned long long)cap_mask);
  for (int probe = 0; probe <= cap_mask; probe++) {
    const unsigned long long old = atomicCAS(keys + slot, ST_EMPTY, h);
    if (old == ST_EMPTY) { atomicMin(rep + slot, (int32_t)c); ids[c] = (uint32_t)slot; return; }
    if (old == h) { atomicMin(rep + slot, (int32_t)c); ids[c] = (uint32_t)slot; return; }  // the representative: the lowest column of the shape
    slot = (slot + 1) & cap_mask;
    if (probe > 64) break;  // a table this crowded means the matrix has no small dictionary
  }
  flags[1] = 1;
  ids[c] = 0;
}

__global__ void __launch_bounds__(256) k_stencil_verify(int64_t ncols, const int64_t *__restrict__ colptr, const int64_t *__restrict__ rowval,
                                                        const int32_t *__restrict__ rep, const uint32_t *__restrict__ ids, int *flags) {
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / ST_LANES;
  const int gl = threadIdx.x % ST_LANES;
  if (c >= ncols) return;
  const int64_t r = rep[ids[c]];
  const int64_t b = colptr[c] - 1, e = colptr[c + 1] - 1, rb = colptr[r] - 1, re = colptr[r + 1] - 1;
  bool bad = (e - b) != (re - rb);
  if (!bad)
    for (int64_t k = gl; k < e - b; k += ST_LANES) bad = bad || (rowval[b + k] - c != rowval[rb + k] - r);
  if (bad) flags[2] = 1;
}

// dictionary entries, one per used slot: [slot, len, offsets (row - column) ...] at stride 2 + maxlen
__global__ void __launch_bounds__(256) k_stencil_dict(int cap, const unsigned long long *__restrict__ keys, const int32_t *__restrict__ rep,
                                                      const int64_t *__restrict__ colptr, const int64_t *__restrict__ rowval, int maxlen, int dmax,
                                                      int32_t *__restrict__ dict, int *nd) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= cap || keys[slot] == ST_EMPTY) return;
  const int idx = atomicAdd(nd, 1);
  if (idx >= dmax) return;
  const int64_t r = rep[slot], b = colptr[r] - 1, len = colptr[r + 1] - 1 - b;
  int32_t *d = dict + (size_t)idx * (2 + maxlen);
  d[0] = slot;
  d[1] = (int32_t)len;
  for (int64_t k = 0; k < len; k++) d[2 + k] = (int32_t)(rowval[b + k] - 1 - r);
}

// first and last non-empty column (colptr is monotone: two binary searches by one thread); range = {ncols, -1} for an empty matrix
__global__ void k_col_range(int64_t ncols, const int64_t *__restrict__ colptr, int *range) {
  const int64_t last = colptr[ncols];  // nnz + 1
  int64_t lo = 0, hi = ncols;          // first c with colptr[c + 1] > 1
  while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (colptr[mid + 1] > 1) hi = mid; else lo = mid + 1; }
  range[0] = (int)lo;
  lo = -1; hi = ncols - 1;             // last c with colptr[c] < last
  while (lo < hi) { const int64_t mid = (lo + hi + 1) >> 1; if (colptr[mid] < last) lo = mid; else hi = mid - 1; }
  range[1] = (int)lo;
}

}  // namespace

// ids[ncols] (uint32 slot numbers) and the dictionary for the CSC (colptr, rowval) on `stream`.  *ok = false: no small dictionary
// (the caller ships int32 row indices).  The caller frees *d_ids and *d_dict with fe_dev_free when *ok.  One host round trip.
int32_t fe_col_stencils(fegpu_ctx *ctx, int64_t ncols, const int64_t *d_colptr, const int64_t *d_rowval, cudaStream_t stream, uint32_t **d_ids,
                        int32_t **d_dict, int *ndict, int *maxlen_out, int *cap_out, int64_t *col_first, int64_t *col_last, bool *ok) {
  constexpr int CAP = 1 << 15, DMAX = 4096, MAXLEN = 126;
  *ok = false;
  *d_ids = nullptr;
  *d_dict = nullptr;
  if (ncols <= 0 || ncols >= ((int64_t)1 << 31)) return FEGPU_OK;
  unsigned long long *keys = nullptr;
  int32_t *rep = nullptr;
  int *flags = nullptr;  // [0..2] flags, [3] dictionary size, [4..5] first / last non-empty column
  auto drop = [&]() {
    if (keys) fe_dev_free(ctx, keys, stream);
    if (rep) fe_dev_free(ctx, rep, stream);
    if (flags) fe_dev_free(ctx, flags, stream);
  };
  auto fail = [&](int32_t s) { drop(); if (*d_ids) fe_dev_free(ctx, *d_ids, stream); if (*d_dict) fe_dev_free(ctx, *d_dict, stream); *d_ids = nullptr; *d_dict = nullptr; return s; };
#define ST_TRY(x) do { int32_t s_ = (x); if (s_ != FEGPU_OK) return fail(s_); } while (0)
#define ST_CUDA(x) do { if ((x) != cudaSuccess) return fail(fegpu_fail(ctx, FEGPU_ERR_CUDA, "column-stencil codec: CUDA call failed")); } while (0)
  ST_TRY(fe_dev_alloc(ctx, (void **)&keys, sizeof(unsigned long long) * CAP, stream));
  ST_TRY(fe_dev_alloc(ctx, (void **)&rep, sizeof(int32_t) * CAP, stream));
  ST_TRY(fe_dev_alloc(ctx, (void **)&flags, sizeof(int) * 6, stream));
  ST_TRY(fe_dev_alloc(ctx, (void **)d_ids, sizeof(uint32_t) * (size_t)ncols, stream));
  ST_TRY(fe_dev_alloc(ctx, (void **)d_dict, sizeof(int32_t) * (size_t)DMAX * (2 + MAXLEN), stream));
  ST_CUDA(cudaMemsetAsync(keys, 0xff, sizeof(unsigned long long) * CAP, stream));
  ST_CUDA(cudaMemsetAsync(rep, 0x7f, sizeof(int32_t) * CAP, stream));
  ST_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * 6, stream));
  const unsigned grid = grid_for(ncols * ST_LANES, 256);
  k_stencil_insert<<<grid, 256, 0, stream>>>(ncols, d_colptr, d_rowval, MAXLEN, CAP - 1, keys, rep, *d_ids, flags);
  k_stencil_verify<<<grid, 256, 0, stream>>>(ncols, d_colptr, d_rowval, rep, *d_ids, flags);
  k_stencil_dict<<<grid_for(CAP, 256), 256, 0, stream>>>(CAP, keys, rep, d_colptr, d_rowval, MAXLEN, DMAX, *d_dict, flags + 3);
  k_col_range<<<1, 1, 0, stream>>>(ncols, d_colptr, flags + 4);
  ctx->launches += 4;
  int h[6] = {0, 0, 0, 0, 0, 0};
  ST_CUDA(cudaMemcpyAsync(h, flags, sizeof(h), cudaMemcpyDeviceToHost, stream));
  ST_CUDA(cudaStreamSynchronize(stream));
  ST_CUDA(cudaGetLastError());
#undef ST_TRY
#undef ST_CUDA
  drop();
  if (h[0] || h[1] || h[2] || h[3] > DMAX || h[3] <= 0 || h[5] < h[4]) {
    fe_dev_free(ctx, *d_ids, stream);
    fe_dev_free(ctx, *d_dict, stream);
    *d_ids = nullptr;
    *d_dict = nullptr;
    return FEGPU_OK;
  }
  *ndict = h[3];
  *col_first = h[4];
  *col_last = h[5];
  *maxlen_out = MAXLEN;
  *cap_out = CAP;
  *ok = true;
  return FEGPU_OK;
}
