// bilform_lin_elastic (FEMMBaseModule.jl:1774-1813) for every 3-D element type and rule the dedicated H8 kernel does not take:
// T4, T10, H20, H27 (and H8 with rules other than 2x2x2).
//
// Register-tiled: the upper triangle of the nne x nne grid of 3x3 node blocks is cut into tiles of 4 row nodes x 2 column
// nodes, one lane per tile (H20: 30 tiles = one warp per element, T10: 9 tiles = three elements per warp, H27: 56 tiles = two
// warps per element).  Per quadrature point the lanes of an element first cooperate on the geometry (Jacobian entries, then
// one node each: gradN = gradNpar * inv(J) (gradN!, FESetModule.jl:507-544) and T_b = (Jac w) D B_b, the 6x3 slice of
// add_btdb_ut_only!'s DB (MatrixUtilityModule.jl:198-206; B_b has three non-zeros per column, DeforModelRedModule.jl:463-468)),
// leave them in shared memory, and then every lane accumulates its 8 blocks (72 FP64 accumulators) from 12 gradients and
// 2 x 18 T values: 216 DFMA per 48 shared loads, where the entry-per-thread kernel needed 6 shared loads per 3 DFMA and was
// bound by shared-memory bandwidth (H20: 3.7 TFLOP/s).  G and T are double buffered: two barriers per point.
// Output: the compact upper-block layout (fegpu_internal.h) on the mesh-structured path, else full matrices in the
// reference's emission order with the lower triangle mirrored (complete_lt!, MatrixUtilityModule.jl:164).
#include "fegpu_internal.h"

namespace {

constexpr int TR = 4, TC = 2;  // tile: row nodes x column nodes

struct TileTab {
  uint8_t I[64], J[64];  // tile -> (row tile, column tile)
  int nt;
};

constexpr int count_tiles(int nne) {
  int n = 0;
  for (int J = 0; J < (nne + TC - 1) / TC; J++)
    for (int I = 0; I < (nne + TR - 1) / TR; I++)
      if (TR * I <= TC * J + TC - 1) n++;  // the tile holds at least one block with row node <= column node
  return n;
}
template <int NNE>
struct Tiles {
  static constexpr int NT = count_tiles(NNE);  // T4 2, H8 6, T10 9, H20 30, H27 56
  static constexpr int BLOCK = (NT <= 32) ? 128 : ((NT + 31) / 32) * 32;
};

struct ElParams {
  const int32_t *conn;
  const double *xyz;
  int64_t nnodes;
  const int32_t *elem_list;
  int64_t nactive;
  const double *dN;  // [npts][3][NNE]
  const double *w;   // [npts]
  int npts;
  double *V;
  int compact;
  double C[36];  // 6x6 col-major
  TileTab tab;
};

// k9[i + 3 j] += B_a[:, i] . T[:, j]   (rows ascending, the zeros of B_a skipped)
__device__ __forceinline__ void block_acc(double *k9, const double *ga, const double *T) {
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const double *d = T + 6 * j;
    // three DFMA straight into the accumulator (instead of DMUL + 2 DFMA + DADD)
    k9[0 + 3 * j] = fma(ga[2], d[4], fma(ga[1], d[3], fma(ga[0], d[0], k9[0 + 3 * j])));
    k9[1 + 3 * j] = fma(ga[2], d[5], fma(ga[0], d[3], fma(ga[1], d[1], k9[1 + 3 * j])));
    k9[2 + 3 * j] = fma(ga[1], d[5], fma(ga[0], d[4], fma(ga[2], d[2], k9[2 + 3 * j])));
  }
}

// cubic-symmetry D (fe_elastic_cubic, fegpu_internal.h): accumulate the outer products P = sum Jw g_a g_b' only ...
__device__ __forceinline__ void outer_acc(double *p9, const double *ga, const double *hb) {
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int i = 0; i < 3; i++) p9[i + 3 * j] = fma(ga[i], hb[j], p9[i + 3 * j]);
}
// ... and form the block from P at the end: K_ij = lam P_ij + mu P_ji (i != j), K_ii = D00 P_ii + mu (tr P - P_ii); c = {D00, lam, mu}
__device__ __forceinline__ void iso_block(double *k9, const double *c) {
  const double tr = k9[0] + k9[4] + k9[8];
  double out[9];
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int i = 0; i < 3; i++)
      out[i + 3 * j] = (i == j) ? fma(c[0], k9[i + 3 * i], c[2] * (tr - k9[i + 3 * i])) : fma(c[1], k9[i + 3 * j], c[2] * k9[j + 3 * i]);
#pragma unroll
  for (int i = 0; i < 9; i++) k9[i] = out[i];
}

template <int NNE, bool ISO>
__global__ void __launch_bounds__(Tiles<NNE>::BLOCK) k_elastic_tiled(const ElParams P) {
  constexpr int NT = Tiles<NNE>::NT;
  constexpr bool WARP = NT <= 32;                       // groups live inside a warp
  constexpr int GS = WARP ? NT : ((NT + 31) / 32) * 32;  // lanes cooperating on one element
  constexpr int EPW = WARP ? 32 / NT : 1;               // elements per warp
  constexpr int GPB = WARP ? 4 * EPW : 1;               // elements per block (4 warps, or one multi-warp element)
  constexpr int TSTR = 19;                              // T stride (18 + pad: lanes reading different b hit different banks)
  constexpr int GRP = NNE * 3 + 10 + 2 * NNE * 3 + 2 * NNE * TSTR;
  constexpr int EM = NNE * 3;

  extern __shared__ double smem[];
  double *sdN = smem;                       // [npts][3][NNE]
  double *sw = sdN + P.npts * 3 * NNE;      // [npts]
  double *sgrp = sw + P.npts;
  int g, gl;
  if (WARP) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    g = w * EPW + lane / NT;
    gl = lane % NT;
    if (lane >= EPW * NT) { g = w * EPW; gl = -1; }  // spare lanes of the warp: they only take part in the barriers
  } else {
    g = 0;
    gl = threadIdx.x;
  }
  double *sX = sgrp + (size_t)g * GRP;  // [NNE][3]
  double *sJ = sX + NNE * 3;            // 9 (+1 pad)
  double *sG = sJ + 10;                 // [2][NNE][3]
  double *sT = sG + 2 * NNE * 3;        // [2][NNE][TSTR]
  auto gsync = [&]() {
    if (WARP) __syncwarp();
    else __syncthreads();
  };

  for (int i = threadIdx.x; i < P.npts * 3 * NNE; i += blockDim.x) sdN[i] = P.dN[i];
  for (int i = threadIdx.x; i < P.npts; i += blockDim.x) sw[i] = P.w[i];
  __syncthreads();

  const bool worker = gl >= 0 && gl < NT;
  const int tI = worker ? P.tab.I[gl] : 0, tJ = worker ? P.tab.J[gl] : 0;
  const int64_t ngroups_total = (int64_t)gridDim.x * GPB;
  const int64_t iters = (P.nactive + ngroups_total - 1) / ngroups_total;
  for (int64_t it = 0; it < iters; it++) {
    const int64_t slot_raw = (it * gridDim.x + blockIdx.x) * GPB + g;
    const bool live = slot_raw < P.nactive;
    const int64_t slot = live ? slot_raw : P.nactive - 1;
    const int64_t e = P.elem_list ? P.elem_list[slot] : slot;
    const int32_t *conn = P.conn + e * NNE;
    gsync();  // the previous element's shared data is no longer read
    if (gl >= 0)
      for (int i = gl; i < NNE * 3; i += GS) {
        const int a = i % NNE, s = i / NNE;
        sX[a * 3 + s] = P.xyz[(int64_t)s * P.nnodes + conn[a]];
      }
    gsync();

    double K[TR][TC][9];
#pragma unroll
    for (int k = 0; k < TR; k++)
#pragma unroll
      for (int b = 0; b < TC; b++)
#pragma unroll
        for (int i = 0; i < 9; i++) K[k][b][i] = 0.0;

    for (int j = 0; j < P.npts; j++) {
      const double *dN = sdN + j * 3 * NNE;  // [3][NNE]
      double *G = sG + (j & 1) * NNE * 3, *T = sT + (j & 1) * NNE * TSTR;
      // 1. Jacobian entries J[s + 3 d] = sum_a X[a][s] dN[d][a]   (locjac!, MatrixUtilityModule.jl:38-68)
      if (gl >= 0)
        for (int k = gl; k < 9; k += GS) {
          const int s = k % 3, d = k / 3;
          double acc = 0.0;
          for (int a = 0; a < NNE; a++) acc += sX[a * 3 + s] * dN[d * NNE + a];
          sJ[k] = acc;
        }
      gsync();
      // 2. every lane: inverse and determinant (cheap, redundant); one node per lane: gradient and T_b
      double Jm[9], inv[9];
#pragma unroll
      for (int k = 0; k < 9; k++) Jm[k] = sJ[k];
#define R(i, jj) Jm[(i - 1) + 3 * (jj - 1)]
      const double det = R(1, 1) * (R(2, 2) * R(3, 3) - R(3, 2) * R(2, 3)) - R(1, 2) * (R(2, 1) * R(3, 3) - R(2, 3) * R(3, 1)) +
                         R(1, 3) * (R(2, 1) * R(3, 2) - R(2, 2) * R(3, 1));
      const double invdet = 1.0 / det;
      inv[0] = (R(2, 2) * R(3, 3) - R(3, 2) * R(2, 3)) * invdet;
      inv[3] = -(R(1, 2) * R(3, 3) - R(1, 3) * R(3, 2)) * invdet;
      inv[6] = (R(1, 2) * R(2, 3) - R(1, 3) * R(2, 2)) * invdet;
      inv[1] = -(R(2, 1) * R(3, 3) - R(2, 3) * R(3, 1)) * invdet;
      inv[4] = (R(1, 1) * R(3, 3) - R(1, 3) * R(3, 1)) * invdet;
      inv[7] = -(R(1, 1) * R(2, 3) - R(2, 1) * R(1, 3)) * invdet;
      inv[2] = (R(2, 1) * R(3, 2) - R(3, 1) * R(2, 2)) * invdet;
      inv[5] = -(R(1, 1) * R(3, 2) - R(3, 1) * R(1, 2)) * invdet;
      inv[8] = (R(1, 1) * R(2, 2) - R(2, 1) * R(1, 2)) * invdet;
#undef R
      const double Jw = det * sw[j];
      if (gl >= 0)
        for (int a = gl; a < NNE; a += GS) {
          double gq[3];
#pragma unroll
          for (int c = 0; c < 3; c++) gq[c] = dN[a] * inv[0 + 3 * c] + dN[NNE + a] * inv[1 + 3 * c] + dN[2 * NNE + a] * inv[2 + 3 * c];
          G[a * 3 + 0] = gq[0]; G[a * 3 + 1] = gq[1]; G[a * 3 + 2] = gq[2];
          double *Ta = T + a * TSTR;
          if (ISO) {
            Ta[0] = Jw * gq[0]; Ta[1] = Jw * gq[1]; Ta[2] = Jw * gq[2];
            continue;
          }
          // column comp x of B_a: rows 0 (g0), 3 (g1), 4 (g2); comp y: rows 1 (g1), 3 (g0), 5 (g2); comp z: rows 2 (g2), 4 (g0), 5 (g1)
#pragma unroll
          for (int mx = 0; mx < 6; mx++) {
            Ta[mx] = Jw * (P.C[mx + 6 * 0] * gq[0] + P.C[mx + 6 * 3] * gq[1] + P.C[mx + 6 * 4] * gq[2]);
            Ta[6 + mx] = Jw * (P.C[mx + 6 * 1] * gq[1] + P.C[mx + 6 * 3] * gq[0] + P.C[mx + 6 * 5] * gq[2]);
            Ta[12 + mx] = Jw * (P.C[mx + 6 * 2] * gq[2] + P.C[mx + 6 * 4] * gq[0] + P.C[mx + 6 * 5] * gq[1]);
          }
        }
      gsync();
      // 3. the lane's tile: 4 row nodes x 2 column nodes
      if (worker) {
        double ga[TR][3];
#pragma unroll
        for (int k = 0; k < TR; k++) {
          const int a = min(TR * tI + k, NNE - 1);
          ga[k][0] = G[a * 3 + 0]; ga[k][1] = G[a * 3 + 1]; ga[k][2] = G[a * 3 + 2];
        }
#pragma unroll
        for (int bb = 0; bb < TC; bb++) {
          const int b = min(TC * tJ + bb, NNE - 1);
          if (ISO) {
            double hb[3];
#pragma unroll
            for (int i = 0; i < 3; i++) hb[i] = T[b * TSTR + i];
#pragma unroll
            for (int k = 0; k < TR; k++) outer_acc(K[k][bb], ga[k], hb);
          } else {
            double Tb[18];
#pragma unroll
            for (int i = 0; i < 18; i++) Tb[i] = T[b * TSTR + i];
#pragma unroll
            for (int k = 0; k < TR; k++) block_acc(K[k][bb], ga[k], Tb);
          }
        }
      }
    }
    // emission
    if (live && worker) {
#pragma unroll
      for (int k = 0; k < TR; k++)
#pragma unroll
        for (int bb = 0; bb < TC; bb++) {
          const int a = TR * tI + k, b = TC * tJ + bb;
          if (a < NNE && b < NNE && a <= b) {
            const bool diag = a == b;
            if (ISO) iso_block(K[k][bb], P.C);
            if (P.compact) {
              double *Vb = P.V + slot * (int64_t)(NNE * (NNE + 1) / 2 * 9) + 9 * (b * (b + 1) / 2 + a);
#pragma unroll
              for (int jx = 0; jx < 3; jx++)
#pragma unroll
                for (int ix = 0; ix < 3; ix++)
                  if (!diag || ix <= jx) {
                    const double v = K[k][bb][ix + 3 * jx];
                    Vb[jx * 3 + ix] = v;
                    if (diag && ix != jx) Vb[ix * 3 + jx] = v;
                  }
            } else {
              double *Ve = P.V + slot * (int64_t)(EM * EM);
#pragma unroll
              for (int jx = 0; jx < 3; jx++)
#pragma unroll
                for (int ix = 0; ix < 3; ix++)
                  if (!diag || ix <= jx) {
                    const double v = K[k][bb][ix + 3 * jx];
                    Ve[(3 * b + jx) * EM + 3 * a + ix] = v;
                    Ve[(3 * a + ix) * EM + 3 * b + jx] = v;  // complete_lt!
                  }
            }
          }
        }
    }
  }
}

template <int NNE>
int32_t launch_tiled(fegpu_mesh *mesh, const FormArgs &fa, double *d_V) {
  fegpu_ctx *ctx = mesh->ctx;
  constexpr int NT = Tiles<NNE>::NT;
  static_assert(NT <= 64, "tile table holds 64 tiles");
  constexpr bool WARP = NT <= 32;
  constexpr int EPW = WARP ? 32 / NT : 1;
  constexpr int GPB = WARP ? 4 * EPW : 1;
  constexpr int BLOCK = WARP ? 128 : ((NT + 31) / 32) * 32;
  constexpr int GRP = NNE * 3 + 10 + 2 * NNE * 3 + 2 * NNE * 19;
  if (mesh->nactive == 0) return FEGPU_OK;
  ElParams P;
  P.conn = mesh->conn_act(); P.xyz = mesh->d_xyz; P.nnodes = mesh->nnodes; P.elem_list = mesh->d_elem_list; P.nactive = mesh->nactive;
  P.dN = mesh->d_tab + (size_t)mesh->npts * NNE;  // the table holds N [npts][NNE] first
  P.w = mesh->d_w; P.npts = mesh->npts; P.V = d_V; P.compact = fa.compact ? 1 : 0;
  for (int i = 0; i < 36; i++) P.C[i] = fa.coef[i];
  double cub[3];
  const bool iso = fe_elastic_cubic(fa.coef, cub);
  if (iso)
    for (int i = 0; i < 3; i++) P.C[i] = cub[i];
  int n = 0;
  for (int J = 0; J < (NNE + TC - 1) / TC; J++)
    for (int I = 0; I < (NNE + TR - 1) / TR; I++)
      if (TR * I <= TC * J + TC - 1) { P.tab.I[n] = (uint8_t)I; P.tab.J[n] = (uint8_t)J; n++; }
  P.tab.nt = n;
  const size_t smem = sizeof(double) * ((size_t)mesh->npts * 3 * NNE + mesh->npts + (size_t)GPB * GRP);
  auto kern = iso ? k_elastic_tiled<NNE, true> : k_elastic_tiled<NNE, false>;
  if (smem > 48 * 1024) CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t want = (mesh->nactive + GPB - 1) / GPB;
  const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)ctx->sm_count * 16);
  kern<<<grid, BLOCK, smem, ctx->stream>>>(P);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return FEGPU_OK;
}

}  // namespace

int32_t fe_integrate_elastic_tiled(fegpu_mesh *mesh, const FormArgs &fa, double *d_V, bool *handled) {
  *handled = true;
  switch (mesh->etype) {
    case FEGPU_T4: return launch_tiled<4>(mesh, fa, d_V);
    case FEGPU_T10: return launch_tiled<10>(mesh, fa, d_V);
    case FEGPU_H8: return launch_tiled<8>(mesh, fa, d_V);
    case FEGPU_H20: return launch_tiled<20>(mesh, fa, d_V);
    case FEGPU_H27: return launch_tiled<27>(mesh, fa, d_V);
  }
  *handled = false;
  return FEGPU_OK;
}
