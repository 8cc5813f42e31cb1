// Element integration kernels (FP64 on the CUDA cores; no tensor cores: the contractions are 8-81 wide).
//
// Generic kernel k_integrate<...>: TPE cooperating threads per element (8 for the small scalar elements, 32 = one
// warp for T10 / H8-elastic / H20-scalar, 128 = one CTA for H20/H27 vector forms).  Per quadrature point the group
//   1. evaluates the isoparametric Jacobian from the element's coordinates (shared memory) and the caller's
//      gradNparams table               -- locjac!  MatrixUtilityModule.jl:60, Jacobian FESetModule.jl:426-489
//   2. forms gradN = gradNparams * inv(J) (adjugate / det)            -- gradN!  FESetModule.jl:453-544
//   3. forms kappa*gradN' (add_gkgt_ut_only! :120) or D*B (blmat! DeforModelRedModule.jl:447, add_btdb_ut_only! :189)
//   4. every thread accumulates its share of the upper-triangle (or full, for bilform_dot) element matrix entries
//      in registers, in the reference's summation order.
// At the end the triangle is mirrored (complete_lt! :164) and written in the reference's emission order
// (AssemblyModule.jl:261-280: column-major, p = e*EM*EM + (j-1)*EM + i) as values only -- the (I,J) keys are
// implied by conn + dofnums and never materialised on the fast path.
#include <cstdlib>

#include "fegpu_internal.h"

namespace {

struct IntegParams {
  const int32_t *conn;       // [nelem][NNE] 0-based
  const double *xyz;         // [SDIM][nnodes]
  int64_t nnodes;
  const int32_t *elem_list;  // active element ids or nullptr
  int64_t nactive;
  const double *tab;         // N [npts][NNE], then dN [npts][MDIM][NNE]
  const double *w;           // [npts]
  int npts;
  double *V;                 // [nactive][EM*EM], or [nactive][fe_compact_size] when compact
  int compact;
  int64_t vstride;           // > 0: plane layout, value k of slot s at V[k * vstride + s] (FormArgs::planes)
  double coef[36];
  int m;
  double otherdim;
  const double *uvel;        // [SDIM][nnodes] convective velocity (FORM_CONVECTION)
  double rm[9];              // constant material coordinate system matrix Rm, SDIM x MDIM column-major
  int use_rm;                // 0 = identity
};

template <int TPE>
__device__ __forceinline__ void group_sync() {
  if (TPE <= 32)
    __syncwarp();
  else
    __syncthreads();
}

// Jacobian determinant / surface measure.  J is SDIM x MDIM, column-major.
template <int SDIM, int MDIM>
__device__ __forceinline__ double jac_measure(const double *J) {
  if (SDIM == 3 && MDIM == 3) {
    return J[0] * (J[4] * J[8] - J[5] * J[7]) - J[3] * (J[1] * J[8] - J[7] * J[2]) + J[6] * (J[1] * J[5] - J[4] * J[2]);
  } else if (SDIM == 2 && MDIM == 2) {
    return J[0] * J[3] - J[1] * J[2];
  } else {  // SDIM == 3, MDIM == 2 : |J1 x J2|
    double c0 = J[1] * J[5] - J[2] * J[4];
    double c1 = J[2] * J[3] - J[0] * J[5];
    double c2 = J[0] * J[4] - J[1] * J[3];
    return sqrt(c0 * c0 + c1 * c1 + c2 * c2);
  }
}

// inverse of the MDIM x MDIM Jacobian, adjugate times 1/det (FESetModule.jl:513-530)
template <int MDIM>
__device__ __forceinline__ void jac_inverse(const double *J, double *inv) {
  if (MDIM == 3) {
#define R(i, j) J[(i - 1) + 3 * (j - 1)]
    double invdet = 1.0 / (R(1, 1) * (R(2, 2) * R(3, 3) - R(3, 2) * R(2, 3)) - R(1, 2) * (R(2, 1) * R(3, 3) - R(2, 3) * R(3, 1)) +
                           R(1, 3) * (R(2, 1) * R(3, 2) - R(2, 2) * R(3, 1)));
    inv[0] = (R(2, 2) * R(3, 3) - R(3, 2) * R(2, 3)) * invdet;   // 11
    inv[3] = -(R(1, 2) * R(3, 3) - R(1, 3) * R(3, 2)) * invdet;  // 12
    inv[6] = (R(1, 2) * R(2, 3) - R(1, 3) * R(2, 2)) * invdet;   // 13
    inv[1] = -(R(2, 1) * R(3, 3) - R(2, 3) * R(3, 1)) * invdet;  // 21
    inv[4] = (R(1, 1) * R(3, 3) - R(1, 3) * R(3, 1)) * invdet;   // 22
    inv[7] = -(R(1, 1) * R(2, 3) - R(2, 1) * R(1, 3)) * invdet;  // 23
    inv[2] = (R(2, 1) * R(3, 2) - R(3, 1) * R(2, 2)) * invdet;   // 31
    inv[5] = -(R(1, 1) * R(3, 2) - R(3, 1) * R(1, 2)) * invdet;  // 32
    inv[8] = (R(1, 1) * R(2, 2) - R(2, 1) * R(1, 2)) * invdet;   // 33
#undef R
  } else {
    double invdet = 1.0 / (J[0] * J[3] - J[2] * J[1]);
    inv[0] = J[3] * invdet;
    inv[2] = -J[2] * invdet;
    inv[1] = -J[1] * invdet;
    inv[3] = J[0] * invdet;
  }
}

// Nonzero rows of column (node, comp) of the 3-D strain-displacement matrix (DeforModelRedModule.jl:463-468 with Rm = I):
// comp x: rows xx(g1) xy(g2) xz(g3); comp y: yy(g2) xy(g1) yz(g3); comp z: zz(g3) xz(g1) yz(g2).  Rows ascending.
// Column (node, comp) of B for a constant non-identity material coordinate system (DeforModelRedModule.jl:463-468): all six rows
__device__ __forceinline__ void bcol_rm(int comp, const double *g, const double *rm, double b[6]) {
  const double r1 = rm[comp], r2 = rm[comp + 3], r3 = rm[comp + 6];  // Rm[j, 1..3]
  b[0] = g[0] * r1;
  b[1] = g[1] * r2;
  b[2] = g[2] * r3;
  b[3] = g[1] * r1 + g[0] * r2;
  b[4] = g[2] * r1 + g[0] * r3;
  b[5] = g[2] * r2 + g[1] * r3;
}

__device__ __forceinline__ void bcol(int comp, const double *g, int rows[3], double vals[3]) {
  if (comp == 0) {
    rows[0] = 0; vals[0] = g[0]; rows[1] = 3; vals[1] = g[1]; rows[2] = 4; vals[2] = g[2];
  } else if (comp == 1) {
    rows[0] = 1; vals[0] = g[1]; rows[1] = 3; vals[1] = g[0]; rows[2] = 5; vals[2] = g[2];
  } else {
    rows[0] = 2; vals[0] = g[2]; rows[1] = 4; vals[1] = g[0]; rows[2] = 5; vals[2] = g[1];
  }
}

template <int NNE, int MDIM, int SDIM, int NDN, int FORM, int TPE>
__global__ void __launch_bounds__(TPE <= 32 ? 128 : TPE) k_integrate(const IntegParams P) {
  constexpr int EM = NNE * NDN;
  constexpr bool SYM = (FORM == FORM_DIFF_ISO || FORM == FORM_DIFF_GEN || FORM == FORM_ELASTIC);
  // linform_dot: a vector; bilform_masslike: an NDN x EM matrix
  constexpr int NENT = (FORM == FORM_LINDOT) ? EM : (FORM == FORM_MASSLIKE ? NDN * EM : (SYM ? EM * (EM + 1) / 2 : EM * EM));
  constexpr int EPT = (NENT + TPE - 1) / TPE;
  constexpr int GPB = (TPE <= 32) ? 128 / TPE : 1;  // element groups per block
  constexpr int NAUX = (FORM == FORM_ELASTIC) ? 6 * EM : (FORM == FORM_DIFF_GEN ? MDIM * NNE : (FORM == FORM_CONVECTION ? NNE * SDIM : 1));

  extern __shared__ double smem[];
  // layout: tables [npts*NNE*(1+MDIM)] | w [npts] | per group: X [NNE*SDIM], G [NNE*MDIM], AUX [NAUX]
  double *sN = smem;
  double *sdN = sN + P.npts * NNE;
  double *sw = sdN + P.npts * NNE * MDIM;
  double *sgrp = sw + P.npts;
  const int g = threadIdx.x / TPE, t = threadIdx.x % TPE;
  double *sX = sgrp + g * (NNE * SDIM + NNE * MDIM + NAUX);
  double *sG = sX + NNE * SDIM;
  double *sA = sG + NNE * MDIM;

  for (int i = threadIdx.x; i < P.npts * NNE * (1 + MDIM); i += blockDim.x) sN[i] = P.tab[i];
  for (int i = threadIdx.x; i < P.npts; i += blockDim.x) sw[i] = P.w[i];
  __syncthreads();

  // entry -> (r, c) decode, once per thread
  int er[EPT], ec[EPT];
#pragma unroll
  for (int k = 0; k < EPT; k++) {
    int idx = t + k * TPE;
    if (idx >= NENT) idx = NENT - 1;  // clamp: duplicates recompute the last entry, stores are predicated
    if (SYM) {
      int c = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
      while (c * (c + 1) / 2 > idx) c--;
      while ((c + 1) * (c + 2) / 2 <= idx) c++;
      er[k] = idx - c * (c + 1) / 2;
      ec[k] = c;
    } else if (FORM == FORM_MASSLIKE) {
      er[k] = idx % NDN;  // row p of the NDN x EM element matrix
      ec[k] = idx / NDN;  // column (b, q)
    } else {
      er[k] = idx % EM;
      ec[k] = idx / EM;  // 0 for the element vector of linform_dot
    }
  }

  const int64_t ngroups_total = (int64_t)gridDim.x * GPB;
  // all groups of a block iterate the same number of times so the barriers stay aligned
  const int64_t iters = (P.nactive + ngroups_total - 1) / ngroups_total;
  for (int64_t it = 0; it < iters; it++) {
    const int64_t slot_raw = (it * gridDim.x + blockIdx.x) * GPB + g;
    const bool live = slot_raw < P.nactive;
    const int64_t slot = live ? slot_raw : P.nactive - 1;
    const int64_t e = P.elem_list ? P.elem_list[slot] : slot;
    const int32_t *conn = P.conn + e * NNE;
    group_sync<TPE>();  // previous element's smem reads are done
    for (int i = t; i < NNE * SDIM; i += TPE) {
      int a = i % NNE, s = i / NNE;
      sX[a * SDIM + s] = P.xyz[(int64_t)s * P.nnodes + conn[a]];
      if (FORM == FORM_CONVECTION) sA[a * SDIM + s] = P.uvel[(int64_t)s * P.nnodes + conn[a]];  // gathervalues_asmat!(u, eus, conn)
    }
    group_sync<TPE>();

    double acc[EPT];
#pragma unroll
    for (int k = 0; k < EPT; k++) acc[k] = 0.0;

    for (int j = 0; j < P.npts; j++) {
      const double *dN = sdN + j * NNE * MDIM;  // [MDIM][NNE]
      const double *N = sN + j * NNE;
      // J = X' * dN  (SDIM x MDIM, column-major), every thread redundantly
      double J[SDIM * MDIM];
#pragma unroll
      for (int i = 0; i < SDIM * MDIM; i++) J[i] = 0.0;
      for (int a = 0; a < NNE; a++) {
#pragma unroll
        for (int d = 0; d < MDIM; d++) {
          double dn = dN[d * NNE + a];
#pragma unroll
          for (int s = 0; s < SDIM; s++) J[s + SDIM * d] += sX[a * SDIM + s] * dn;
        }
      }
      double Jac = jac_measure<SDIM, MDIM>(J);
      if (FORM == FORM_MASSLIKE) {
        // factor = Ns[b] * Jac * w ; elmat[p, (b, q)] += factor * c[p, q]               FEMMBaseModule.jl:1896-1903
        if (MDIM == 2 && P.m == 3) Jac = Jac * P.otherdim;
#pragma unroll
        for (int k = 0; k < EPT; k++) {
          const int p = er[k], c = ec[k];
          acc[k] += (N[c / NDN] * Jac * sw[j]) * P.coef[p + NDN * (c % NDN)];
        }
      } else if (FORM == FORM_LINDOT) {
        // elvec[rx] += (Ns[kx] * (Jac * w)) * force[mx]                                FEMMBaseModule.jl:1230-1239
        if (MDIM == 2 && P.m == 3) Jac = Jac * P.otherdim;
        const double Factor = Jac * sw[j];
#pragma unroll
        for (int k = 0; k < EPT; k++) {
          const int r = er[k];
          acc[k] += (N[r / NDN] * Factor) * P.coef[r % NDN];
        }
      } else if (FORM == FORM_DOT) {
        if (MDIM == 2 && P.m == 3) Jac = Jac * P.otherdim;
        const double Jw = Jac * sw[j];
#pragma unroll
        for (int k = 0; k < EPT; k++) {
          const int r = er[k], c = ec[k];
          const int kn = r / NDN, pp = r % NDN, mn = c / NDN, qq = c % NDN;
          // factor = (Ns[k]*Ns[m]*Jac*w) ; elmat += factor*c[p,q]      FEMMBaseModule.jl:1356-1359
          const double factor = N[kn] * N[mn] * Jac * sw[j];
          acc[k] += factor * P.coef[pp + NDN * qq];
        }
        (void)Jw;
      } else {
        // gradN rows
        double inv[MDIM * MDIM];
        if (SDIM == MDIM) {
          if (P.use_rm && (FORM == FORM_DIFF_GEN || FORM == FORM_ELASTIC)) {
            // RmTJ = Rm' * J (mulCAtB! / At_mul_B!, FEMMBaseModule.jl:1496, 1802): gradients in the material directions
            double R[MDIM * MDIM];
#pragma unroll
            for (int a = 0; a < MDIM; a++)
#pragma unroll
              for (int b = 0; b < MDIM; b++) {
                double acc2 = 0.0;
#pragma unroll
                for (int k2 = 0; k2 < SDIM; k2++) acc2 += P.rm[k2 + SDIM * a] * J[k2 + SDIM * b];
                R[a + MDIM * b] = acc2;
              }
            jac_inverse<MDIM>(R, inv);
          } else {
            jac_inverse<MDIM>(J, inv);
          }
        }
        if (MDIM == 2) Jac = Jac * P.otherdim;  // Jacobianvolume of a 2-manifold: surface Jacobian x other dimension (IntegDomainModule.jl:504-517)
        const double Jw = Jac * sw[j];
        group_sync<TPE>();  // previous point's G / AUX reads are done
        for (int a = t; a < NNE; a += TPE) {
#pragma unroll
          for (int c = 0; c < MDIM; c++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < MDIM; k++) s += dN[k * NNE + a] * inv[k + MDIM * c];
            sG[a * MDIM + c] = s;
          }
        }
        group_sync<TPE>();
        double us[SDIM];
        if (FORM == FORM_CONVECTION) {
          // u_s = sum_q Ns[q] * eus[q, s]                                            FEMMBaseModule.jl:1611-1614
#pragma unroll
          for (int s = 0; s < SDIM; s++) {
            double a = 0.0;
            for (int q = 0; q < NNE; q++) a += N[q] * sA[q * SDIM + s];
            us[s] = a;
          }
        }
        if (FORM == FORM_DIFF_GEN) {
          // kappa_bargradNT[mx, nx] = Jac_w * sum_px kappa[mx,px]*gradN[nx,px]     MatrixUtilityModule.jl:134-142
          for (int i = t; i < MDIM * NNE; i += TPE) {
            const int nx = i / MDIM, mx = i % MDIM;
            double a = 0.0;
#pragma unroll
            for (int px = 0; px < MDIM; px++) a += P.coef[mx + MDIM * px] * sG[nx * MDIM + px];
            sA[mx + MDIM * nx] = Jw * a;
          }
          group_sync<TPE>();
        } else if (FORM == FORM_ELASTIC) {
          // DB[:, c] = Jac_w * D * B[:, c]                                          MatrixUtilityModule.jl:198-206
          for (int c = t; c < EM; c += TPE) {
            if (P.use_rm) {
              double b6[6];
              bcol_rm(c % 3, sG + (c / 3) * 3, P.rm, b6);
#pragma unroll
              for (int mx = 0; mx < 6; mx++) {
                double a = 0.0;
#pragma unroll
                for (int px = 0; px < 6; px++) a += P.coef[mx + 6 * px] * b6[px];
                sA[mx + 6 * c] = Jw * a;
              }
            } else {
              int rows[3];
              double vals[3];
              bcol(c % 3, sG + (c / 3) * 3, rows, vals);
#pragma unroll
              for (int mx = 0; mx < 6; mx++) {
                double a = 0.0;
#pragma unroll
                for (int q = 0; q < 3; q++) a += P.coef[mx + 6 * rows[q]] * vals[q];
                sA[mx + 6 * c] = Jw * a;
              }
            }
          }
          group_sync<TPE>();
        }
#pragma unroll
        for (int k = 0; k < EPT; k++) {
          const int r = er[k], c = ec[k];
          if (FORM == FORM_DIFF_ISO) {
            // Ke[mx,nx] += gradN[mx,px] * (mult*gradN[nx,px]), px ascending          MatrixUtilityModule.jl:90-96
            const double mult = P.coef[0] * Jac * sw[j];
#pragma unroll
            for (int px = 0; px < MDIM; px++) acc[k] += sG[r * MDIM + px] * (mult * sG[c * MDIM + px]);
          } else if (FORM == FORM_DIFF_GEN) {
            double a = 0.0;
#pragma unroll
            for (int px = 0; px < MDIM; px++) a += sG[r * MDIM + px] * sA[px + MDIM * c];
            acc[k] += a;
          } else if (FORM == FORM_CONVECTION) {
            // elmat[p, r] += Ns[p] * (sum_s u_s * gradN[r, s]) * (Jac * w)               FEMMBaseModule.jl:1608-1618
            double a = 0.0;
#pragma unroll
            for (int s = 0; s < SDIM; s++) a += us[s] * sG[c * MDIM + s];
            acc[k] += N[r] * a * Jw;
          } else if (FORM == FORM_DIV_GRAD) {
            // p = (a, s), r = (b, t): factor * (delta_st * sum_q gradN[a,q] gradN[b,q] + gradN[a,t] gradN[b,s])   :1694-1707
            const double factor = P.coef[0] * Jw;
            const int na = r / NDN, s = r % NDN, nb = c / NDN, tt = c % NDN;
            if (s == tt) {
#pragma unroll
              for (int q = 0; q < NDN; q++) acc[k] += factor * sG[na * MDIM + q] * sG[nb * MDIM + q];
            }
            acc[k] += factor * sG[na * MDIM + tt] * sG[nb * MDIM + s];
          } else if (P.use_rm) {  // FORM_ELASTIC with a material coordinate system: dense columns of B
            double b6[6];
            bcol_rm(r % 3, sG + (r / 3) * 3, P.rm, b6);
            double a = 0.0;
#pragma unroll
            for (int px = 0; px < 6; px++) a += b6[px] * sA[px + 6 * c];
            acc[k] += a;
          } else {  // FORM_ELASTIC: accum = sum_px B[px,mx]*DB[px,nx]
            int rows[3];
            double vals[3];
            bcol(r % 3, sG + (r / 3) * 3, rows, vals);
            double a = 0.0;
#pragma unroll
            for (int q = 0; q < 3; q++) a += vals[q] * sA[rows[q] + 6 * c];
            acc[k] += a;
          }
        }
      }
    }
    // emission: V[slot][c*EM + r]; the mirrored entry is complete_lt!
    if (FORM == FORM_LINDOT || FORM == FORM_MASSLIKE) {
      if (live) {  // entry idx = t + k*TPE is the position in the element vector / the column-major NDN x EM matrix
#pragma unroll
        for (int k = 0; k < EPT; k++)
          if (t + k * TPE < NENT) P.V[slot * (int64_t)NENT + t + k * TPE] = acc[k];
      }
    } else if (live && SYM && P.compact) {
      // compact upper-block layout (fegpu_internal.h): block (a <= b) at NDN^2 * (b(b+1)/2 + a), column-major inside;
      // element-major records, or planes (value index x vstride + slot) for the thread-per-node numeric kernel
      constexpr int ND2 = NDN * NDN;
      constexpr int64_t REC = (int64_t)(NNE * (NNE + 1) / 2 * ND2);
#pragma unroll
      for (int k = 0; k < EPT; k++) {
        if (t + k * TPE < NENT) {
          const int r = er[k], c = ec[k];
          const int a = r / NDN, i = r % NDN, b = c / NDN, j = c % NDN;
          const int blk = b * (b + 1) / 2 + a;
          if (P.vstride > 0) {  // value planes: entry e of block blk of slot s at ((blk * ND2 + e) * vstride + s)
            double *Vb = P.V + (int64_t)blk * ND2 * P.vstride + slot;
            Vb[(int64_t)(j * NDN + i) * P.vstride] = acc[k];
            if (a == b && i != j) Vb[(int64_t)(i * NDN + j) * P.vstride] = acc[k];
          } else {
            double *Vb = P.V + slot * REC + ND2 * blk;
            Vb[j * NDN + i] = acc[k];
            if (a == b && i != j) Vb[i * NDN + j] = acc[k];
          }
        }
      }
    } else if (live) {
      // full matrix, emission order; planes: block (column node b, row node a) = plane b * NNE + a, entry j * NDN + i
      constexpr int ND2 = NDN * NDN;
      double *Ve = P.V + slot * (int64_t)(EM * EM);
#pragma unroll
      for (int k = 0; k < EPT; k++) {
        if (t + k * TPE < NENT) {
          const int r = er[k], c = ec[k];
          if (P.vstride > 0) {
            const int a = r / NDN, i = r % NDN, b = c / NDN, j = c % NDN;
            P.V[((int64_t)(b * NNE + a) * ND2 + j * NDN + i) * P.vstride + slot] = acc[k];
            if (SYM && r != c) P.V[((int64_t)(a * NNE + b) * ND2 + i * NDN + j) * P.vstride + slot] = acc[k];
          } else {
            Ve[c * EM + r] = acc[k];
            if (SYM && r != c) Ve[r * EM + c] = acc[k];
          }
        }
      }
    }
  }
}

template <int NNE, int MDIM, int SDIM, int NDN, int FORM, int TPE>
int32_t launch_generic(fegpu_mesh *mesh, const FormArgs &fa, double *d_V) {
  fegpu_ctx *ctx = mesh->ctx;
  constexpr int EM = NNE * NDN;
  constexpr int GPB = (TPE <= 32) ? 128 / TPE : 1;
  constexpr int NAUX = (FORM == FORM_ELASTIC) ? 6 * EM : (FORM == FORM_DIFF_GEN ? MDIM * NNE : (FORM == FORM_CONVECTION ? NNE * SDIM : 1));
  IntegParams P;
  P.conn = mesh->conn_act(); P.xyz = mesh->d_xyz; P.nnodes = mesh->nnodes; P.elem_list = mesh->d_elem_list;
  P.nactive = mesh->nactive; P.tab = mesh->d_tab; P.w = mesh->d_w; P.npts = mesh->npts; P.V = d_V;
  P.compact = (fa.compact && fe_form_symmetric(FORM)) ? 1 : 0;
  P.vstride = (fa.planes && FORM != FORM_LINDOT && FORM != FORM_MASSLIKE) ? fa.vstride : 0;
  for (int i = 0; i < 36; i++) P.coef[i] = fa.coef[i];
  P.m = fa.m; P.otherdim = fa.otherdim; P.uvel = fa.d_uvel;
  for (int i = 0; i < 9; i++) P.rm[i] = fa.rm[i];
  P.use_rm = fa.use_rm ? 1 : 0;
  if (mesh->nactive == 0) return FEGPU_OK;
  size_t smem = sizeof(double) * ((size_t)mesh->npts * NNE * (1 + MDIM) + mesh->npts + (size_t)GPB * (NNE * SDIM + NNE * MDIM + NAUX));
  auto kern = k_integrate<NNE, MDIM, SDIM, NDN, FORM, TPE>;
  if (smem > 48 * 1024) CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t want = (mesh->nactive + GPB - 1) / GPB;
  int64_t cap = (int64_t)ctx->sm_count * 16;
  unsigned grid = (unsigned)std::min<int64_t>(want, cap);
  kern<<<grid, (TPE <= 32 ? 128 : TPE), smem, ctx->stream>>>(P);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return FEGPU_OK;
}

template <int NNE, int MDIM, int SDIM, int TPE_S, int TPE_V>
int32_t dispatch_form(fegpu_mesh *mesh, const FormArgs &fa, double *d_V) {
  switch (fa.form) {
    case FORM_DIFF_ISO:
      if (SDIM != MDIM) break;
      return launch_generic<NNE, MDIM, (SDIM == MDIM ? SDIM : MDIM), 1, FORM_DIFF_ISO, TPE_S>(mesh, fa, d_V);
    case FORM_DIFF_GEN:
      if (SDIM != MDIM) break;
      return launch_generic<NNE, MDIM, (SDIM == MDIM ? SDIM : MDIM), 1, FORM_DIFF_GEN, TPE_S>(mesh, fa, d_V);
    case FORM_ELASTIC:
      if (SDIM != 3 || MDIM != 3) break;
      return launch_generic<NNE, 3, 3, 3, FORM_ELASTIC, TPE_V>(mesh, fa, d_V);
    case FORM_DOT:
      if (fa.ndn == 1) return launch_generic<NNE, MDIM, SDIM, 1, FORM_DOT, TPE_S>(mesh, fa, d_V);
      if (fa.ndn == 2) return launch_generic<NNE, MDIM, SDIM, 2, FORM_DOT, TPE_V>(mesh, fa, d_V);
      if (fa.ndn == 3) return launch_generic<NNE, MDIM, SDIM, 3, FORM_DOT, TPE_V>(mesh, fa, d_V);
      return fegpu_fail(mesh->ctx, FEGPU_ERR_ARG, "bilform_dot: 1, 2 or 3 dofs per node are supported");
    case FORM_LINDOT:
      if (fa.ndn == 1) return launch_generic<NNE, MDIM, SDIM, 1, FORM_LINDOT, TPE_S>(mesh, fa, d_V);
      if (fa.ndn == 2) return launch_generic<NNE, MDIM, SDIM, 2, FORM_LINDOT, TPE_S>(mesh, fa, d_V);
      if (fa.ndn == 3) return launch_generic<NNE, MDIM, SDIM, 3, FORM_LINDOT, TPE_S>(mesh, fa, d_V);
      return fegpu_fail(mesh->ctx, FEGPU_ERR_ARG, "linform_dot: 1, 2 or 3 dofs per node are supported");
    case FORM_MASSLIKE:
      if (fa.ndn == 1) return launch_generic<NNE, MDIM, SDIM, 1, FORM_MASSLIKE, TPE_S>(mesh, fa, d_V);
      if (fa.ndn == 2) return launch_generic<NNE, MDIM, SDIM, 2, FORM_MASSLIKE, TPE_S>(mesh, fa, d_V);
      if (fa.ndn == 3) return launch_generic<NNE, MDIM, SDIM, 3, FORM_MASSLIKE, TPE_V>(mesh, fa, d_V);
      return fegpu_fail(mesh->ctx, FEGPU_ERR_ARG, "bilform_masslike: 1, 2 or 3 dofs per node are supported");
    case FORM_CONVECTION:
      if (SDIM != MDIM) break;
      return launch_generic<NNE, MDIM, (SDIM == MDIM ? SDIM : MDIM), 1, FORM_CONVECTION, TPE_S>(mesh, fa, d_V);
    case FORM_DIV_GRAD:
      if (SDIM != MDIM) break;
      return launch_generic<NNE, MDIM, (SDIM == MDIM ? SDIM : MDIM), (SDIM == MDIM ? SDIM : MDIM), FORM_DIV_GRAD, TPE_V>(mesh, fa, d_V);
  }
  return fegpu_fail(mesh->ctx, FEGPU_ERR_ARG, "form not defined for this element manifold / space dimension");
}

}  // namespace

int32_t fe_integrate_h8(fegpu_mesh *mesh, const FormArgs &fa, double *d_V, bool *handled);             // fegpu_h8.cu
int32_t fe_integrate_elastic_tiled(fegpu_mesh *mesh, const FormArgs &fa, double *d_V, bool *handled);  // fegpu_elastic.cu

bool fe_dot_scalar_applies(const fegpu_mesh *mesh, const FormArgs &fa);              // fegpu_dot.cu
int32_t fe_integrate_dot_scalar(fegpu_mesh *mesh, const FormArgs &fa, double *d_V);  // fegpu_dot.cu

bool fe_integrate_supports_compact(const fegpu_mesh *mesh, const FormArgs &fa) {
  return fe_form_symmetric(fa.form) || fe_dot_scalar_applies(mesh, fa);  // a 1 x 1 coefficient makes bilform_dot symmetric
}

// Which kernels write the plane layout: the H8 kernels, the scalar mass kernel and the generic entry-per-thread kernel -- i.e.
// every bilinear form on the element types the thread-per-node path takes, except the register-tiled elasticity kernel (T4
// elasticity keeps element-major records) and the Kronecker path of bilform_dot with more than 3 dofs per node.
bool fe_integrate_supports_planes(const fegpu_mesh *mesh, const FormArgs &fa) {
  if (fa.form == FORM_LINDOT || fa.form == FORM_MASSLIKE) return false;
  // Scalar fields, and elasticity on H8 with the 8-point rule (k_h8_elastic writes value planes from its staging buffer with
  // coalesced 256-byte stores).  The other vector-field kernels keep element-major records: the NDN lanes of a node read one
  // contiguous block of the record, and strided plane stores from registers cost more than they gave
  // (profiles/r02_bench_n1_planes_variants.txt).
  const bool rotated = fa.use_rm && (fa.form == FORM_DIFF_GEN || fa.form == FORM_ELASTIC);
  static const bool vec_planes_off = std::getenv("FEGPU_VEC_PLANES") && std::atoi(std::getenv("FEGPU_VEC_PLANES")) == 0;  // A/B knob
  if (fa.ndn != 1) return !vec_planes_off && fa.form == FORM_ELASTIC && fa.ndn == 3 && mesh->etype == FEGPU_H8 && mesh->npts == 8 && mesh->sdim == 3 && !rotated;
  if (fa.form == FORM_ELASTIC && mesh->sdim == 3 && mesh->mdim == 3 && !rotated && !(mesh->etype == FEGPU_H8 && mesh->npts == 8)) {
    static const bool tiled_off = std::getenv("FEGPU_ELASTIC_TILED") && std::atoi(std::getenv("FEGPU_ELASTIC_TILED")) == 0;
    if (!tiled_off) return false;  // k_elastic_tiled
  }
  return true;
}

namespace {
// bilform_dot with more than 3 dofs per node (the reference has no cap, FEMMBaseModule.jl:1355-1360): the element matrix is the
// Kronecker product of the scalar one with the coefficient, elmat[(k,p),(m,q)] = (sum_j N_k N_m Jac w_j) c[p,q].  The reference
// multiplies by c inside the quadrature loop; summing first differs from that by rounding only (c is constant).
__global__ void k_kron_coef(const double *__restrict__ Ms, double *__restrict__ V, int64_t nactive, int nne, int ndn, const double *__restrict__ cdev) {
  const int EM = nne * ndn;
  const int64_t EM2 = (int64_t)EM * EM;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nactive * EM2) return;
  const int64_t e = t / EM2;
  const int loc = (int)(t - e * EM2);
  const int col = loc / EM, row = loc - col * EM;  // emission order: column-major
  const int m = col / ndn, q = col - m * ndn, k = row / ndn, p = row - k * ndn;
  V[t] = Ms[e * (int64_t)(nne * nne) + (int64_t)m * nne + k] * cdev[p + ndn * q];
}
}  // namespace

int32_t fe_integrate(fegpu_mesh *mesh, const FormArgs &fa, double *d_V) {
  if (mesh->npts <= 0) return fegpu_fail(mesh->ctx, FEGPU_ERR_STATE, "no quadrature rule set (fegpu_rule_set)");
  if (mesh->nactive <= 0) return FEGPU_OK;  // empty FESet, or a rank that owns no node: nothing to integrate
  if (fa.form == FORM_DOT && fa.ndn > 3) {
    fegpu_ctx *ctx = mesh->ctx;
    cudaStream_t st = ctx->stream;
    FormArgs f1 = fa;
    f1.ndn = 1;
    f1.coef[0] = 1.0;
    f1.compact = false;
    double *d_Ms = nullptr, *d_c = nullptr;
    const int nne = mesh->nne;
    FE_TRY(fe_dev_alloc(ctx, (void **)&d_Ms, sizeof(double) * (size_t)mesh->nactive * nne * nne, st));
    int32_t s = fe_dev_alloc(ctx, (void **)&d_c, sizeof(double) * 36, st);
    if (s == FEGPU_OK) s = fe_integrate(mesh, f1, d_Ms);  // scalar mass matrices, full layout
    if (s == FEGPU_OK && cudaMemcpyAsync(d_c, fa.coef, sizeof(double) * fa.ndn * fa.ndn, cudaMemcpyHostToDevice, st) != cudaSuccess)
      s = fegpu_fail(ctx, FEGPU_ERR_CUDA, "coefficient upload failed");
    if (s == FEGPU_OK) {
      const int64_t n = mesh->nactive * (int64_t)(nne * fa.ndn) * (nne * fa.ndn);
      k_kron_coef<<<grid_for(n, 256), 256, 0, st>>>(d_Ms, d_V, mesh->nactive, nne, fa.ndn, d_c);
      ctx->launches++;
      if (cudaGetLastError() != cudaSuccess) s = fegpu_fail(ctx, FEGPU_ERR_CUDA, "k_kron_coef launch failed");
    }
    if (s == FEGPU_OK && cudaStreamSynchronize(st) != cudaSuccess) s = fegpu_fail(ctx, FEGPU_ERR_CUDA, "synchronize failed");  // fa.coef is the caller's
    if (d_c) fe_dev_free(ctx, d_c, st);
    fe_dev_free(ctx, d_Ms, st);
    return s;
  }
  if (fe_dot_scalar_applies(mesh, fa)) return fe_integrate_dot_scalar(mesh, fa, d_V);
  const bool rotated = fa.use_rm && (fa.form == FORM_DIFF_GEN || fa.form == FORM_ELASTIC);  // only the generic kernel knows Rm
  if (mesh->etype == FEGPU_H8 && !rotated) {
    bool handled = false;
    FE_TRY(fe_integrate_h8(mesh, fa, d_V, &handled));
    if (handled) return FEGPU_OK;
  }
  if (fa.form == FORM_ELASTIC && mesh->sdim == 3 && mesh->mdim == 3 && !rotated) {
    // register-tiled kernel (fegpu_elastic.cu); FEGPU_ELASTIC_TILED=0 keeps the entry-per-thread kernel for A/B measurements
    static const bool tiled_off = std::getenv("FEGPU_ELASTIC_TILED") && std::atoi(std::getenv("FEGPU_ELASTIC_TILED")) == 0;
    if (!tiled_off) {
      bool handled = false;
      FE_TRY(fe_integrate_elastic_tiled(mesh, fa, d_V, &handled));
      if (handled) return FEGPU_OK;
    }
  }
  switch (mesh->etype) {
    case FEGPU_T3:
      if (mesh->sdim == 2) return dispatch_form<3, 2, 2, 8, 8>(mesh, fa, d_V);
      return dispatch_form<3, 2, 3, 8, 8>(mesh, fa, d_V);
    case FEGPU_Q4:
      if (mesh->sdim == 2) return dispatch_form<4, 2, 2, 8, 8>(mesh, fa, d_V);
      return dispatch_form<4, 2, 3, 8, 8>(mesh, fa, d_V);
    case FEGPU_T4: return dispatch_form<4, 3, 3, 8, 32>(mesh, fa, d_V);
    case FEGPU_T10: return dispatch_form<10, 3, 3, 32, 32>(mesh, fa, d_V);
    case FEGPU_H8: return dispatch_form<8, 3, 3, 8, 32>(mesh, fa, d_V);
    case FEGPU_H20: return dispatch_form<20, 3, 3, 32, 128>(mesh, fa, d_V);
    case FEGPU_H27: return dispatch_form<27, 3, 3, 32, 128>(mesh, fa, d_V);
  }
  return fegpu_fail(mesh->ctx, FEGPU_ERR_ARG, "unknown element type");
}
