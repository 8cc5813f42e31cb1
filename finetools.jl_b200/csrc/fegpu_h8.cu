// Specialised H8 kernels (the benchmark configurations C1/C2/C4 of BASELINE.json).
//
//  k_h8_diffusion : thread per element.  Coordinates, Jacobian, gradients and the 36-entry upper triangle live in
//                   registers; the quadrature tables sit in __constant__ memory (uniform operands of the DFMAs).
//  k_h8_elastic   : 32 elements per CTA, 4 warps.  Phase A: thread (element, 2 quadrature points) computes Jacobian,
//                   inverse and the 8 nodal gradients into shared memory (element index fastest => conflict free).
//                   Phase B: warp t owns the node-column pair (t, 7-t) of the 8x8 grid of 3x3 blocks -- exactly 9 upper
//                   blocks for every t, so no lane divergence -- and lane = element.  D*B_b (DB, add_btdb_ut_only!
//                   MatrixUtilityModule.jl:198-206) is formed once per column node and point and reused by its blocks.
//                   The 24x24 matrix is staged through shared memory and written with coalesced 128-bit stores in the
//                   reference's emission order.
// Only GaussRule(3,2) (8 points) takes these paths; other rules use the generic kernel.
#include <cstdlib>
#include <cstring>

#include "fegpu_internal.h"

namespace {

// The quadrature tables and the coefficient travel as KERNEL PARAMETERS (1.9 KB of the 4 KB parameter space): they land in
// the launch's own constant bank, i.e. they are uniform operands of the DFMAs exactly like __constant__ data, but they belong
// to this launch -- two contexts on one device (or two queued forms on different streams) cannot overwrite each other's
// tables between upload and kernel, and no cudaMemcpyToSymbol precedes the launch.
struct H8Params {
  const int32_t *conn;
  const double *xyz;
  int64_t nnodes;
  const int32_t *elem_list;
  int64_t nactive;
  double *V;
  int64_t vstride;       // > 0: plane layout, value k of slot s at V[k * vstride + s] (FormArgs::planes)
  double dN[8 * 3 * 8];  // [point][dim][node]
  double w[8];
  double coef[36];
};

__device__ __forceinline__ void inv3(const double *J, double *inv, double &det) {
#define R(i, j) J[(i - 1) + 3 * (j - 1)]
  det = R(1, 1) * (R(2, 2) * R(3, 3) - R(3, 2) * R(2, 3)) - R(1, 2) * (R(2, 1) * R(3, 3) - R(2, 3) * R(3, 1)) +
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
}

// COMPACT: only the upper triangle (36 values, packed by columns: entry (r <= c) at c(c+1)/2 + r) is written -- the layout
// k_gather reads for symmetric forms on the mesh-structured path; otherwise the full 8x8 matrix in emission order.
template <bool GENERAL, bool COMPACT>
__global__ void __launch_bounds__(128) k_h8_diffusion(const __grid_constant__ H8Params P) {
  const double *c_dN = P.dN, *c_w = P.w, *c_coef = P.coef;
  const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= P.nactive) return;
  const int64_t e = P.elem_list ? P.elem_list[slot] : slot;
  int nd[8];
  {
    const int4 *c4 = reinterpret_cast<const int4 *>(P.conn + e * 8);
    int4 a = __ldg(c4), b = __ldg(c4 + 1);
    nd[0] = a.x; nd[1] = a.y; nd[2] = a.z; nd[3] = a.w; nd[4] = b.x; nd[5] = b.y; nd[6] = b.z; nd[7] = b.w;
  }
  double X[8][3];
#pragma unroll
  for (int a = 0; a < 8; a++)
#pragma unroll
    for (int s = 0; s < 3; s++) X[a][s] = __ldg(P.xyz + (int64_t)s * P.nnodes + nd[a]);

  double acc[36];
#pragma unroll
  for (int i = 0; i < 36; i++) acc[i] = 0.0;

#pragma unroll 1
  for (int j = 0; j < 8; j++) {
    const double *dN = c_dN + j * 24;
    double J[9];
#pragma unroll
    for (int i = 0; i < 9; i++) J[i] = 0.0;
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
      for (int d = 0; d < 3; d++)
#pragma unroll
        for (int s = 0; s < 3; s++) J[s + 3 * d] += X[a][s] * dN[d * 8 + a];
    double inv[9], det;
    inv3(J, inv, det);
    const double Jw = det * c_w[j];
    double G[8][3];
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
      for (int c = 0; c < 3; c++) G[a][c] = dN[a] * inv[0 + 3 * c] + dN[8 + a] * inv[1 + 3 * c] + dN[16 + a] * inv[2 + 3 * c];
    if (GENERAL) {
      // kG[px][nx] = Jw * sum_q kappa[px,q] G[nx][q]   (add_gkgt_ut_only!, MatrixUtilityModule.jl:134-142)
#pragma unroll
      for (int nx = 0; nx < 8; nx++) {
        double kg[3];
#pragma unroll
        for (int mx = 0; mx < 3; mx++) {
          double a = 0.0;
#pragma unroll
          for (int px = 0; px < 3; px++) a += c_coef[mx + 3 * px] * G[nx][px];
          kg[mx] = Jw * a;
        }
#pragma unroll
        for (int mx = 0; mx <= nx; mx++) {  // three DFMA straight into the accumulator
          double a = acc[nx * (nx + 1) / 2 + mx];
#pragma unroll
          for (int px = 0; px < 3; px++) a = fma(G[mx][px], kg[px], a);
          acc[nx * (nx + 1) / 2 + mx] = a;
        }
      }
    } else {
      const double mult = c_coef[0] * det * c_w[j];  // (c * Jac * w[j])  FEMMBaseModule.jl:1528
#pragma unroll
      for (int nx = 0; nx < 8; nx++)
#pragma unroll
        for (int px = 0; px < 3; px++) {
          const double a = mult * G[nx][px];
#pragma unroll
          for (int mx = 0; mx <= nx; mx++) acc[nx * (nx + 1) / 2 + mx] += G[mx][px] * a;
        }
    }
  }
  if (P.vstride > 0) {  // planes: the lanes of a warp (consecutive slots) write consecutive words of every plane
    double *out = P.V + slot;
    if (COMPACT) {
#pragma unroll
      for (int i = 0; i < 36; i++) out[(int64_t)i * P.vstride] = acc[i];
    } else {
#pragma unroll
      for (int c = 0; c < 8; c++)
#pragma unroll
        for (int r = 0; r < 8; r++) out[(int64_t)(c * 8 + r) * P.vstride] = (r <= c) ? acc[c * (c + 1) / 2 + r] : acc[r * (r + 1) / 2 + c];
    }
    return;
  }
  if (COMPACT) {
    double2 *out = reinterpret_cast<double2 *>(P.V + slot * 36);
#pragma unroll
    for (int i = 0; i < 36; i += 2) out[i >> 1] = make_double2(acc[i], acc[i + 1]);
    return;
  }
  // complete_lt! + emission order: V[slot][c*8 + r]
  double2 *out = reinterpret_cast<double2 *>(P.V + slot * 64);
#pragma unroll
  for (int c = 0; c < 8; c++)
#pragma unroll
    for (int r = 0; r < 8; r += 2) {
      const int r0 = r, r1 = r + 1;
      const double v0 = (r0 <= c) ? acc[c * (c + 1) / 2 + r0] : acc[r0 * (r0 + 1) / 2 + c];
      const double v1 = (r1 <= c) ? acc[c * (c + 1) / 2 + r1] : acc[r1 * (r1 + 1) / 2 + c];
      out[(c * 8 + r) >> 1] = make_double2(v0, v1);
    }
}

// ------------------------------------------------------------------------------------------------ elasticity
constexpr int EL_EPB = 32;                 // elements per block
constexpr int EL_GSTRIDE = 25;             // doubles per (element, point): 24 gradients + Jw
// shared: G [8 pts][25][32 elems] doubles = 51200 B; staging for 16 elements aliases G
constexpr int EL_SMEM_FULL = 16 * 577 * 8;          // staging is the larger user
// compact layout: all 32 element matrices are staged at once (32 * 325 * 8 = 83 200 B, two CTAs per SM still fit): one staging
// pass with every lane active instead of two passes with half of them (FEGPU_ELASTIC_STAGE=2 keeps the two-pass version, 51 200 B)
constexpr int EL_SMEM_COMPACT1 = 32 * 325 * 8;
constexpr int EL_SMEM_COMPACT = 8 * EL_GSTRIDE * EL_EPB * 8;  // G is (16 * 325 * 8 = 41600 B of staging fits inside)

__device__ __forceinline__ void db_col(const double *g, double Jw, double DB[18], const double *c_coef) {
  // DB[:, j] = D * (Jw * B_b[:, j]), B_b column j has 3 non-zeros (DeforModelRedModule.jl:463-468, Rm = I)
  // comp x: rows 0(g0) 3(g1) 4(g2); comp y: rows 1(g1) 3(g0) 5(g2); comp z: rows 2(g2) 4(g0) 5(g1)
  // The scalar Jac*w multiplies the three gradients once (3 DMUL) instead of the 18 entries: 57 instead of 72 FP64
  // instructions per column node and point; the result differs from (Jac*w)*(D*B) by rounding only.
  const double h0 = Jw * g[0], h1 = Jw * g[1], h2 = Jw * g[2];
#pragma unroll
  for (int mx = 0; mx < 6; mx++) {
    DB[mx] = fma(c_coef[mx + 6 * 4], h2, fma(c_coef[mx + 6 * 3], h1, c_coef[mx + 6 * 0] * h0));
    DB[6 + mx] = fma(c_coef[mx + 6 * 5], h2, fma(c_coef[mx + 6 * 3], h0, c_coef[mx + 6 * 1] * h1));
    DB[12 + mx] = fma(c_coef[mx + 6 * 5], h1, fma(c_coef[mx + 6 * 4], h0, c_coef[mx + 6 * 2] * h2));
  }
}

__device__ __forceinline__ void block_acc(double *k9, const double *ga, const double DB[18]) {
  // k9[i + 3*j] += B_a[:, i] . DB[:, j]   (rows ascending, as add_btdb_ut_only! sums px = 1..6 skipping the zeros); the
  // three products are folded straight into the accumulator (3 DFMA instead of DMUL + 2 DFMA + DADD)
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const double *d = DB + 6 * j;
    k9[0 + 3 * j] = fma(ga[2], d[4], fma(ga[1], d[3], fma(ga[0], d[0], k9[0 + 3 * j])));
    k9[1 + 3 * j] = fma(ga[2], d[5], fma(ga[0], d[3], fma(ga[1], d[1], k9[1 + 3 * j])));
    k9[2 + 3 * j] = fma(ga[1], d[5], fma(ga[0], d[4], fma(ga[2], d[2], k9[2 + 3 * j])));
  }
}

// Cubic-symmetry D (isotropic materials are the common case, MatDeforElastIsoModule: D = [lam + a on the normal diagonal, lam
// off it] (+) mu I_3, nothing else): B_a(:,i)' D B_b(:,j) collapses to  lam g_a,i g_b,j + mu g_a,j g_b,i  (i != j)  and
// D00 g_a,i g_b,i + mu sum_{k != i} g_a,k g_b,k  (i == j), so the quadrature only has to accumulate the 3 x 3 outer products
// P = sum_pts Jw g_a g_b' (9 DFMA per block and point instead of 27 + the D B columns) and the block is formed from P once at the
// end.  Same numbers as the general kernel up to rounding (different summation order), 4.1 x fewer FP64 instructions.
__device__ __forceinline__ void outer_acc(double *p9, const double *ga, const double hb[3]) {
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int i = 0; i < 3; i++) p9[i + 3 * j] = fma(ga[i], hb[j], p9[i + 3 * j]);
}
// in place: P (row comp i, col comp j at i + 3 j) -> K; coef = {D00, lam, mu}
__device__ __forceinline__ void iso_block(double *k9, const double *coef) {
  const double d00 = coef[0], lam = coef[1], mu = coef[2];
  const double tr = k9[0] + k9[4] + k9[8];
  double out[9];
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int i = 0; i < 3; i++)
      out[i + 3 * j] = (i == j) ? fma(d00, k9[i + 3 * i], mu * (tr - k9[i + 3 * i])) : fma(lam, k9[i + 3 * j], mu * k9[j + 3 * i]);
#pragma unroll
  for (int i = 0; i < 9; i++) k9[i] = out[i];
}

// staged element-matrix strides (doubles).  Odd => the 16 lanes of a half-warp (one element each, same offset) hit 16
// distinct 8-byte bank pairs: conflict-free staging stores.
constexpr int EL_MSTRIDE_FULL = 577;      // 576 values, emission order
constexpr int EL_MSTRIDE_COMPACT = 325;   // 36 upper 3x3 blocks: block (a <= b) at 9*(b(b+1)/2 + a), column-major inside
constexpr int EL_MSTRIDE_BULK = 326;      // the same, 16-byte aligned records for the bulk copy engine
constexpr int EL_SMEM_BULK = 32 * EL_MSTRIDE_BULK * 8;

// Phase B + staging for the warp owning column nodes B1 = t and B2 = 7 - t.  All warps run the SAME code (t is a
// warp-uniform runtime value): slots 0..4 always belong to column B2, slot 8 always to B1, slots 5..7 to B2 iff
// s < 8 - t.  One DB column is live at a time.  (A fully static 4-way instantiation was 61 KB of SASS and starved the
// 32 KB instruction cache: 49 % of the stall samples were "no_instructions".)
// CTA-wide barrier of phase B spelled as a named barrier (id 1, 128 threads).
__device__ __forceinline__ void block_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ void load3(double *d, const double *g, int node) {
  d[0] = g[(node * 3 + 0) * EL_EPB];
  d[1] = g[(node * 3 + 1) * EL_EPB];
  d[2] = g[(node * 3 + 2) * EL_EPB];
}

// Bulk asynchronous copy shared -> global (the TMA engine's non-tensor form, SASS UBLKCP): the staged element matrices leave
// the SM without a load / store loop of the CTA's own warps, which go on with the next tile meanwhile.
__device__ __forceinline__ void bulk_store(double *dst, const double *src_shared, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"((uint32_t)__cvta_generic_to_shared(src_shared)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// NPASS: staging passes (2: 16 elements at a time, 1: all 32 at once -- needs 32 * MSTRIDE doubles of shared memory)
// BULK (compact, one pass): the staged matrices are handed to the copy engine, one 2592-byte bulk store per element (stride 326
// doubles: 16-byte aligned sources; the even stride costs a 2-way bank conflict on the staging stores); the caller waits for the
// engine to have READ the staging area before it overwrites it
template <bool COMPACT, int NPASS, bool BULK = false, bool ISO = false>
__device__ __forceinline__ void elastic_phase_b(double *sm, const int lane, const int t, const int64_t slot0, const H8Params &P) {
  static_assert(!BULK || (COMPACT && NPASS == 1), "bulk stores: compact layout, single staging pass");
  constexpr int EPP = 32 / NPASS;  // elements per staging pass
  constexpr int MSTRIDE = BULK ? EL_MSTRIDE_BULK : (COMPACT ? EL_MSTRIDE_COMPACT : EL_MSTRIDE_FULL);
  constexpr int MSIZE = COMPACT ? 324 : 576;
  const int B1 = t, B2 = 7 - t, NB2 = 8 - t;
  double K[9][9];
#pragma unroll
  for (int s = 0; s < 9; s++)
#pragma unroll
    for (int i = 0; i < 9; i++) K[s][i] = 0.0;
#pragma unroll 1
  for (int j = 0; j < 8; j++) {
    const double *g = sm + (size_t)j * EL_GSTRIDE * EL_EPB + lane;
    const double Jw = g[24 * EL_EPB];
    double gb[3], ga[3];
    if (ISO) {
      double hb[3];
      load3(gb, g, B2);
      hb[0] = Jw * gb[0]; hb[1] = Jw * gb[1]; hb[2] = Jw * gb[2];
#pragma unroll
      for (int s = 0; s < 5; s++) {
        load3(ga, g, s);
        outer_acc(K[s], ga, hb);
      }
#pragma unroll
      for (int s = 5; s < 8; s++)
        if (s < NB2) {
          load3(ga, g, s);
          outer_acc(K[s], ga, hb);
        }
      load3(gb, g, B1);
      hb[0] = Jw * gb[0]; hb[1] = Jw * gb[1]; hb[2] = Jw * gb[2];
#pragma unroll
      for (int s = 5; s < 8; s++)
        if (s >= NB2) {
          load3(ga, g, s - NB2);
          outer_acc(K[s], ga, hb);
        }
      load3(ga, g, t);
      outer_acc(K[8], ga, hb);
    } else {
      double DB[18];
      load3(gb, g, B2);
      db_col(gb, Jw, DB, P.coef);
#pragma unroll
      for (int s = 0; s < 5; s++) {
        load3(ga, g, s);
        block_acc(K[s], ga, DB);
      }
#pragma unroll
      for (int s = 5; s < 8; s++)
        if (s < NB2) {
          load3(ga, g, s);
          block_acc(K[s], ga, DB);
        }
      load3(gb, g, B1);
      db_col(gb, Jw, DB, P.coef);
#pragma unroll
      for (int s = 5; s < 8; s++)
        if (s >= NB2) {
          load3(ga, g, s - NB2);
          block_acc(K[s], ga, DB);
        }
      load3(ga, g, t);  // slot 8: block (a = t, b = B1), the diagonal block of column B1
      block_acc(K[8], ga, DB);
    }
  }
  if (ISO) {
#pragma unroll
    for (int s = 0; s < 9; s++) iso_block(K[s], P.coef);
  }
  block_bar();  // everyone is done reading G: the staging buffer may overwrite it

  // ---- stage + write: two halves of 16 elements
  for (int half = 0; half < NPASS; half++) {
    if (NPASS == 1 || (lane >> 4) == half) {
      double *M = sm + (size_t)(lane & (EPP - 1)) * MSTRIDE;
#pragma unroll
      for (int s = 0; s < 9; s++) {
        const int a = (s < NB2) ? s : s - NB2;
        const int b = (s < NB2) ? B2 : B1;
        const bool diag = (a == b);
        if (COMPACT) {
          double *Mb = M + 9 * (b * (b + 1) / 2 + a);  // 3x3 block (row node a, column node b), column-major
#pragma unroll
          for (int jx = 0; jx < 3; jx++)
#pragma unroll
            for (int ix = 0; ix < 3; ix++)
              if (!diag || ix <= jx) {  // diagonal block: its upper triangle is the reference's value, mirrored (complete_lt!)
                const double v = K[s][ix + 3 * jx];
                Mb[jx * 3 + ix] = v;
                if (diag && ix != jx) Mb[ix * 3 + jx] = v;
              }
        } else {
          double *Mu = M + (b * 3) * 24 + a * 3;  // (row a-block, column b-block)
          double *Ml = M + (a * 3) * 24 + b * 3;  // mirrored block
#pragma unroll
          for (int jx = 0; jx < 3; jx++)
#pragma unroll
            for (int ix = 0; ix < 3; ix++)
              if (!diag || ix <= jx) {  // diagonal block: only its upper triangle is the reference's value
                const double v = K[s][ix + 3 * jx];
                Mu[jx * 24 + ix] = v;
                Ml[ix * 24 + jx] = v;  // complete_lt!
              }
        }
      }
    }
    if (BULK) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the staging stores become visible to the copy engine
      block_bar();
      if (t == 0 && slot0 + lane < P.nactive) bulk_store(P.V + (slot0 + lane) * MSIZE, sm + (size_t)lane * MSTRIDE, MSIZE * 8);
      return;
    }
    block_bar();
    const int64_t sbase = slot0 + half * EPP;
    const int64_t nvalid = min((int64_t)EPP, P.nactive - sbase);
    if (P.vstride > 0) {
      // value planes: warp t writes planes t, t + 4, ...; lane = element (256-byte coalesced stores, conflict-free reads: odd stride)
      if (lane < nvalid)
        for (int k = t; k < MSIZE; k += 4) P.V[(int64_t)k * P.vstride + sbase + lane] = sm[(size_t)lane * MSTRIDE + k];
    } else if (nvalid > 0) {
      const int nval = (int)(nvalid * MSIZE);  // contiguous slots are contiguous in V
      double *dst = P.V + sbase * MSIZE;
      for (int i = threadIdx.x; i < nval; i += 128) {
        const int el = i / MSIZE, k = i - el * MSIZE;
        dst[i] = sm[(size_t)el * MSTRIDE + k];
      }
    }
    block_bar();
  }
}

// MINB: CTAs per SM the register allocation aims at (2: 244 registers, no spills; 3: 168 registers with ~380 B of spills --
// FEGPU_ELASTIC_CTAS=3 selects it for A/B measurements)
template <bool COMPACT, int MINB, int NPASS, bool ISO = false>
__global__ void __launch_bounds__(128, MINB) k_h8_elastic(const __grid_constant__ H8Params P) {
  const double *c_dN = P.dN, *c_w = P.w;
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, t = threadIdx.x >> 5;  // lane = element in block, t = column pair
  const int64_t slot0 = (int64_t)blockIdx.x * EL_EPB;
  const int64_t slot_raw = slot0 + lane;
  const bool live = slot_raw < P.nactive;
  const int64_t slot = live ? slot_raw : P.nactive - 1;
  const int64_t e = P.elem_list ? P.elem_list[slot] : slot;

  // ---- phase A: gradients for points 2t, 2t+1 of element `lane`
  {
    int nd[8];
    const int4 *c4 = reinterpret_cast<const int4 *>(P.conn + e * 8);
    int4 a = __ldg(c4), b = __ldg(c4 + 1);
    nd[0] = a.x; nd[1] = a.y; nd[2] = a.z; nd[3] = a.w; nd[4] = b.x; nd[5] = b.y; nd[6] = b.z; nd[7] = b.w;
    double X[8][3];
#pragma unroll
    for (int n = 0; n < 8; n++)
#pragma unroll
      for (int s = 0; s < 3; s++) X[n][s] = __ldg(P.xyz + (int64_t)s * P.nnodes + nd[n]);
#pragma unroll 1
    for (int jj = 0; jj < 2; jj++) {
      const int j = 2 * t + jj;
      const double *dN = c_dN + j * 24;
      double J[9];
#pragma unroll
      for (int i = 0; i < 9; i++) J[i] = 0.0;
#pragma unroll
      for (int n = 0; n < 8; n++)
#pragma unroll
        for (int d = 0; d < 3; d++)
#pragma unroll
          for (int s = 0; s < 3; s++) J[s + 3 * d] += X[n][s] * dN[d * 8 + n];
      double inv[9], det;
      inv3(J, inv, det);
      double *g = sm + (size_t)j * EL_GSTRIDE * EL_EPB + lane;
#pragma unroll
      for (int n = 0; n < 8; n++)
#pragma unroll
        for (int c = 0; c < 3; c++)
          g[(n * 3 + c) * EL_EPB] = dN[n] * inv[0 + 3 * c] + dN[8 + n] * inv[1 + 3 * c] + dN[16 + n] * inv[2 + 3 * c];
      g[24 * EL_EPB] = det * c_w[j];
    }
  }
  __syncthreads();
  elastic_phase_b<COMPACT, NPASS, false, ISO>(sm, lane, t, slot0, P);
}

// Persistent version with bulk stores: a CTA walks tiles of 32 elements; the copy engine drains tile i's staged matrices while the
// warps fetch the coordinates of tile i+1 (the staging area aliases G, so G is written only after the engine has read it).
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_h8_elastic_bulk(const __grid_constant__ H8Params P, const int64_t ntiles) {
  const double *c_dN = P.dN, *c_w = P.w;
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, t = threadIdx.x >> 5;
#pragma unroll 1
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t slot0 = tile * EL_EPB;
    const int64_t slot_raw = slot0 + lane;
    const int64_t slot = slot_raw < P.nactive ? slot_raw : P.nactive - 1;
    const int64_t e = P.elem_list ? P.elem_list[slot] : slot;
    int nd[8];
    const int4 *c4 = reinterpret_cast<const int4 *>(P.conn + e * 8);
    int4 a = __ldg(c4), b = __ldg(c4 + 1);
    nd[0] = a.x; nd[1] = a.y; nd[2] = a.z; nd[3] = a.w; nd[4] = b.x; nd[5] = b.y; nd[6] = b.z; nd[7] = b.w;
    double X[8][3];
#pragma unroll
    for (int n = 0; n < 8; n++)
#pragma unroll
      for (int s = 0; s < 3; s++) X[n][s] = __ldg(P.xyz + (int64_t)s * P.nnodes + nd[n]);
    if (t == 0) bulk_wait_read();  // warp 0 issued the previous tile's stores
    __syncthreads();
#pragma unroll 1
    for (int jj = 0; jj < 2; jj++) {
      const int j = 2 * t + jj;
      const double *dN = c_dN + j * 24;
      double J[9];
#pragma unroll
      for (int i = 0; i < 9; i++) J[i] = 0.0;
#pragma unroll
      for (int n = 0; n < 8; n++)
#pragma unroll
        for (int d = 0; d < 3; d++)
#pragma unroll
          for (int s = 0; s < 3; s++) J[s + 3 * d] += X[n][s] * dN[d * 8 + n];
      double inv[9], det;
      inv3(J, inv, det);
      double *g = sm + (size_t)j * EL_GSTRIDE * EL_EPB + lane;
#pragma unroll
      for (int n = 0; n < 8; n++)
#pragma unroll
        for (int c = 0; c < 3; c++)
          g[(n * 3 + c) * EL_EPB] = dN[n] * inv[0 + 3 * c] + dN[8 + n] * inv[1 + 3 * c] + dN[16 + n] * inv[2 + 3 * c];
      g[24 * EL_EPB] = det * c_w[j];
    }
    __syncthreads();
    elastic_phase_b<true, 1, true>(sm, lane, t, slot0, P);
  }
  if (t == 0) bulk_wait_all();  // shared memory must outlive the engine's reads
}

}  // namespace

int32_t fe_integrate_h8(fegpu_mesh *mesh, const FormArgs &fa, double *d_V, bool *handled) {
  *handled = false;
  fegpu_ctx *ctx = mesh->ctx;
  if (mesh->npts != 8 || mesh->sdim != 3) return FEGPU_OK;
  const bool diff = (fa.form == FORM_DIFF_ISO || fa.form == FORM_DIFF_GEN);
  const bool elast = (fa.form == FORM_ELASTIC);
  if (!diff && !elast) return FEGPU_OK;
  if (mesh->nactive == 0) { *handled = true; return FEGPU_OK; }
  H8Params P{mesh->conn_act(), mesh->d_xyz, mesh->nnodes, mesh->d_elem_list, mesh->nactive, d_V, fa.planes ? fa.vstride : 0, {0}, {0}, {0}};
  // dN part of the host table: [npts][3][8] starting after N [npts][8]
  std::memcpy(P.dN, mesh->h_tab.data() + 8 * 8, sizeof(double) * 8 * 24);
  std::memcpy(P.w, mesh->h_w.data(), sizeof(double) * 8);
  std::memcpy(P.coef, fa.coef, sizeof(double) * 36);
  if (diff) {
    unsigned grid = grid_for(mesh->nactive, 128);
    // (Co-residency with k_sym_tile was measured: holding this kernel at one CTA per SM so that two CTAs of the symbolic kernel fit
    // beside it made the fresh step slower, 11.2 against 9.2 ms on config 4 -- the two kernels compete for issue slots.)
    if (fa.form == FORM_DIFF_GEN) {
      if (fa.compact) k_h8_diffusion<true, true><<<grid, 128, 0, ctx->stream>>>(P);
      else k_h8_diffusion<true, false><<<grid, 128, 0, ctx->stream>>>(P);
    } else {
      if (fa.compact) k_h8_diffusion<false, true><<<grid, 128, 0, ctx->stream>>>(P);
      else k_h8_diffusion<false, false><<<grid, 128, 0, ctx->stream>>>(P);
    }
  } else {
    // the attribute is per device (a process may hold contexts on several): set it on every launch, like every other kernel here
    static const bool three = std::getenv("FEGPU_ELASTIC_CTAS") && std::atoi(std::getenv("FEGPU_ELASTIC_CTAS")) == 3;
    unsigned grid = grid_for(mesh->nactive, EL_EPB);
    static const bool two_pass = std::getenv("FEGPU_ELASTIC_STAGE") && std::atoi(std::getenv("FEGPU_ELASTIC_STAGE")) == 2;
    double cub[3];
    const bool iso = fe_elastic_cubic(fa.coef, cub);  // cubic-symmetry D: the outer-product formulation, P.coef = {D00, lam, mu}
    if (iso) std::memcpy(P.coef, cub, sizeof(cub));
#define EL_LAUNCH_I(C_, M_, NP_, SM_, I_)                                                                                       \
  do {                                                                                                                          \
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_h8_elastic<C_, M_, NP_, I_>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_));       \
    k_h8_elastic<C_, M_, NP_, I_><<<grid, 128, SM_, ctx->stream>>>(P);                                                          \
  } while (0)
#define EL_LAUNCH(C_, M_, NP_, SM_)                    \
  do {                                                 \
    if (iso) EL_LAUNCH_I(C_, M_, NP_, SM_, true);      \
    else EL_LAUNCH_I(C_, M_, NP_, SM_, false);         \
  } while (0)
    // persistent CTAs + copy-engine stores (element-major records only): measured 2.62 ms against 2.53 ms for the plain kernel on
    // config 2 (profiles/r02_ncu_final_c2.txt: the copy-out loop was not what keeps the FP64 pipe at 55 %) -- A/B knob, off
    static const bool bulk_on = std::getenv("FEGPU_ELASTIC_BULK") && std::atoi(std::getenv("FEGPU_ELASTIC_BULK")) == 1;
    if (fa.compact && bulk_on && !iso && !fa.planes && !three && !two_pass) {
      const int64_t ntiles = (mesh->nactive + EL_EPB - 1) / EL_EPB;
      const unsigned pgrid = (unsigned)std::min<int64_t>(ntiles, (int64_t)ctx->sm_count * 2);
      CUDA_TRY(ctx, cudaFuncSetAttribute(k_h8_elastic_bulk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, EL_SMEM_BULK));
      k_h8_elastic_bulk<2><<<pgrid, 128, EL_SMEM_BULK, ctx->stream>>>(P, ntiles);
    } else if (fa.compact) {
      if (three) EL_LAUNCH(true, 3, 2, EL_SMEM_COMPACT);
      else if (two_pass) EL_LAUNCH(true, 2, 2, EL_SMEM_COMPACT);
      else EL_LAUNCH(true, 2, 1, EL_SMEM_COMPACT1);
    } else {
      if (three) EL_LAUNCH(false, 3, 2, EL_SMEM_FULL);
      else EL_LAUNCH(false, 2, 2, EL_SMEM_FULL);
    }
#undef EL_LAUNCH
#undef EL_LAUNCH_I
  }
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  *handled = true;
  return FEGPU_OK;
}
