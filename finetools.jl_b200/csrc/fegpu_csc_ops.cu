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
