/*
 * fe_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C99) of the
 * FinEtools.jl v8.2.11 element-integration-and-assembly path.  It is the checker
 * for the CUDA library and the timed serial CPU baseline; nothing in the product
 * path (finetools.jl_b200/) may call it.
 *
 * Parity status: the reference is Julia-only and cannot run in this environment,
 * so this oracle is pinned to the reference's own known-answer tests (gradN golden
 * vectors, the 7x7 assembler matrix, the integral identities of test/test_forms.jl)
 * -- see tests/test_oracle_*.py.  Entry-by-entry stiffness values are NOT pinned by
 * any reference test: "parity unpinned" at the entry level.
 *
 * Every function cites the reference file:line (relative to /root/reference/src)
 * whose loops it follows.  Build with -ffp-contract=off: Julia does not contract
 * a*b+c into FMA unless asked, so neither may this file.
 *
 * Array conventions = the reference's (Julia): matrices are column-major,
 * indices crossing the API are 1-based int64.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

enum { ET_T3 = 1, ET_Q4 = 2, ET_T4 = 3, ET_T10 = 4, ET_H8 = 5, ET_H20 = 6, ET_H27 = 7 };

ORC_API int orc_nne(int et) {
  switch (et) {
    case ET_T3: return 3;  case ET_Q4: return 4;  case ET_T4: return 4;  case ET_T10: return 10;
    case ET_H8: return 8;  case ET_H20: return 20; case ET_H27: return 27;
  }
  return -1;
}
ORC_API int orc_mdim(int et) { return (et == ET_T3 || et == ET_Q4) ? 2 : 3; }

/* ------------------------------------------------------------------ basis */
/* N[nne]; FESetModule.jl:665 (T3) :713 (Q4) :1344 (T4) :1395 (T10) :955 (H8) :1030 (H20) :1217 (H27) */
ORC_API int orc_bfun(int et, const double *pc, double *N) {
  switch (et) {
    case ET_T3:
      N[0] = (1 - pc[0] - pc[1]); N[1] = pc[0]; N[2] = pc[1];
      return 0;
    case ET_Q4:
      N[0] = 0.25 * (1.0 - pc[0]) * (1.0 - pc[1]);
      N[1] = 0.25 * (1.0 + pc[0]) * (1.0 - pc[1]);
      N[2] = 0.25 * (1.0 + pc[0]) * (1.0 + pc[1]);
      N[3] = 0.25 * (1.0 - pc[0]) * (1.0 + pc[1]);
      return 0;
    case ET_T4:
      N[0] = (1 - pc[0] - pc[1] - pc[2]); N[1] = pc[0]; N[2] = pc[1]; N[3] = pc[2];
      return 0;
    case ET_T10: {
      double r = pc[0], s = pc[1], t = pc[2];
      N[0] = (1 - r - s - t) * (2 * (1 - r - s - t) - 1);
      N[1] = r * (2 * r - 1);
      N[2] = s * (2 * s - 1);
      N[3] = t * (2 * t - 1);
      N[4] = 4 * (1 - r - s - t) * r;
      N[5] = 4 * r * s;
      N[6] = 4 * s * (1 - r - s - t);
      N[7] = 4 * (1 - r - s - t) * t;
      N[8] = 4 * r * t;
      N[9] = 4 * s * t;
      return 0;
    }
    case ET_H8: {
      double omx = (1.0 - pc[0]), ome = (1.0 - pc[1]), omt = (1.0 - pc[2]);
      double opx = (1.0 + pc[0]), ope = (1.0 + pc[1]), opt = (1.0 + pc[2]);
      N[0] = omx * ome * omt / 8.0; N[1] = opx * ome * omt / 8.0;
      N[2] = opx * ope * omt / 8.0; N[3] = omx * ope * omt / 8.0;
      N[4] = omx * ome * opt / 8.0; N[5] = opx * ome * opt / 8.0;
      N[6] = opx * ope * opt / 8.0; N[7] = omx * ope * opt / 8.0;
      return 0;
    }
    case ET_H20: {
      double x = pc[0], y = pc[1], z = pc[2];
      double xim = (-1 + x), etam = (-1 + y), zetam = (-1 + z);
      double xip = (1 + x), etap = (1 + y), zetap = (1 + z);
      N[0] = 1.0 / 8 * xim * etam * zetam * (2 + x + y + z);
      N[1] = -1.0 / 8 * xip * etam * zetam * (2 - x + y + z);
      N[2] = 1.0 / 8 * xip * etap * zetam * (2 - x - y + z);
      N[3] = -1.0 / 8 * xim * etap * zetam * (2 + x - y + z);
      N[4] = 1.0 / 8 * xim * etam * zetap * (-2 - x - y + z);
      N[5] = -1.0 / 8 * xip * etam * zetap * (-2 + x - y + z);
      N[6] = 1.0 / 8 * xip * etap * zetap * (-2 + x + y + z);
      N[7] = -1.0 / 8 * xim * etap * zetap * (-2 - x + y + z);
      N[8] = -1.0 / 4 * xim * xip * etam * zetam;
      N[9] = 1.0 / 4 * etam * etap * xip * zetam;
      N[10] = 1.0 / 4 * xim * xip * etap * zetam;
      N[11] = -1.0 / 4 * etam * etap * xim * zetam;
      N[12] = 1.0 / 4 * xim * xip * etam * zetap;
      N[13] = -1.0 / 4 * etam * etap * xip * zetap;
      N[14] = -1.0 / 4 * xim * xip * etap * zetap;
      N[15] = 1.0 / 4 * etam * etap * xim * zetap;
      N[16] = -1.0 / 4 * zetam * zetap * xim * etam;
      N[17] = 1.0 / 4 * zetam * zetap * xip * etam;
      N[18] = -1.0 / 4 * zetam * zetap * xip * etap;
      N[19] = 1.0 / 4 * zetam * zetap * xim * etap;
      return 0;
    }
    case ET_H27: {
      double xi = pc[0], eta = pc[1], zet = pc[2];
      double x1 = (xi - 1), y1 = (eta - 1), z1 = (zet - 1);
      double x2 = (xi + 1), y2 = (eta + 1), z2 = (zet + 1);
      N[0] = 1.0 / 8.0 * z1 * zet * x1 * xi * y1 * eta;
      N[1] = 1.0 / 8.0 * z1 * zet * x2 * xi * y1 * eta;
      N[2] = 1.0 / 8.0 * z1 * zet * x2 * xi * y2 * eta;
      N[3] = 1.0 / 8.0 * z1 * zet * x1 * xi * y2 * eta;
      N[4] = 1.0 / 8.0 * z2 * zet * x1 * xi * y1 * eta;
      N[5] = 1.0 / 8.0 * z2 * zet * x2 * xi * y1 * eta;
      N[6] = 1.0 / 8.0 * z2 * zet * x2 * xi * y2 * eta;
      N[7] = 1.0 / 8.0 * z2 * zet * x1 * xi * y2 * eta;
      N[8] = 1.0 / 4.0 * z1 * zet * (-x2) * x1 * y1 * eta;
      N[9] = 1.0 / 4.0 * z1 * zet * x2 * xi * (-y2) * y1;
      N[10] = 1.0 / 4.0 * z1 * zet * (-x2) * x1 * y2 * eta;
      N[11] = 1.0 / 4.0 * z1 * zet * x1 * xi * (-y2) * y1;
      N[12] = 1.0 / 4.0 * z2 * zet * (-x2) * x1 * y1 * eta;
      N[13] = 1.0 / 4.0 * z2 * zet * x2 * xi * (-y2) * y1;
      N[14] = 1.0 / 4.0 * z2 * zet * (-x2) * x1 * y2 * eta;
      N[15] = 1.0 / 4.0 * z2 * zet * x1 * xi * (-y2) * y1;
      N[16] = 1.0 / 4.0 * (-z2) * z1 * x1 * xi * y1 * eta;
      N[17] = 1.0 / 4.0 * (-z2) * z1 * x2 * xi * y1 * eta;
      N[18] = 1.0 / 4.0 * (-z2) * z1 * x2 * xi * y2 * eta;
      N[19] = 1.0 / 4.0 * (-z2) * z1 * x1 * xi * y2 * eta;
      N[20] = 1.0 / 2.0 * z1 * zet * (-x2) * x1 * (-y2) * y1;
      N[21] = 1.0 / 2.0 * (-z2) * z1 * (-x2) * x1 * y1 * eta;
      N[22] = 1.0 / 2.0 * (-z2) * z1 * x2 * xi * (-y2) * y1;
      N[23] = 1.0 / 2.0 * (-z2) * z1 * (-x2) * x1 * y2 * eta;
      N[24] = 1.0 / 2.0 * (-z2) * z1 * x1 * xi * (-y2) * y1;
      N[25] = 1.0 / 2.0 * z2 * zet * (-x2) * x1 * (-y2) * y1;
      N[26] = (-z2) * z1 * (-x2) * x1 * (-y2) * y1;
      return 0;
    }
  }
  return -1;
}

/* dN[nne x mdim] column-major. FESetModule.jl:678 :724 :1355 :1415 :977 :1095 :1260 */
ORC_API int orc_bfundpar(int et, const double *pc, double *dN) {
  int nne = orc_nne(et);
#define D(r, c) dN[(r) + (size_t)nne * (c)]
  switch (et) {
    case ET_T3:
      D(0, 0) = -1.0; D(0, 1) = -1.0; D(1, 0) = +1.0; D(1, 1) = 0.0; D(2, 0) = 0.0; D(2, 1) = +1.0;
      return 0;
    case ET_Q4:
      D(0, 0) = -(1.0 - pc[1]) * 0.25; D(0, 1) = -(1.0 - pc[0]) * 0.25;
      D(1, 0) = (1.0 - pc[1]) * 0.25;  D(1, 1) = -(1.0 + pc[0]) * 0.25;
      D(2, 0) = (1.0 + pc[1]) * 0.25;  D(2, 1) = (1.0 + pc[0]) * 0.25;
      D(3, 0) = -(1.0 + pc[1]) * 0.25; D(3, 1) = (1.0 - pc[0]) * 0.25;
      return 0;
    case ET_T4:
      D(0, 0) = -1.0; D(0, 1) = -1.0; D(0, 2) = -1.0;
      D(1, 0) = +1.0; D(1, 1) = 0.0;  D(1, 2) = 0.0;
      D(2, 0) = 0.0;  D(2, 1) = +1.0; D(2, 2) = 0.0;
      D(3, 0) = 0.0;  D(3, 1) = 0.0;  D(3, 2) = +1.0;
      return 0;
    case ET_T10: {
      double r = pc[0], s = pc[1], t = pc[2];
      double c0[10] = {-3 + 4 * r + 4 * s + 4 * t, 4 * r - 1, 0, 0, -8 * r + 4 - 4 * s - 4 * t, 4 * s, -4 * s, -4 * t, 4 * t, 0};
      double c1[10] = {-3 + 4 * r + 4 * s + 4 * t, 0, 4 * s - 1, 0, -4 * r, 4 * r, 4 - 4 * r - 8 * s - 4 * t, -4 * t, 0, 4 * t};
      double c2[10] = {-3 + 4 * r + 4 * s + 4 * t, 0, 0, 4 * t - 1, -4 * r, 0, -4 * s, -8 * t + 4 - 4 * r - 4 * s, 4 * r, 4 * s};
      for (int i = 0; i < 10; i++) { D(i, 0) = c0[i]; D(i, 1) = c1[i]; D(i, 2) = c2[i]; }
      return 0;
    }
    case ET_H8: {
      double omxi = (1.0 - pc[0]), ometa = (1.0 - pc[1]), omtheta = (1.0 - pc[2]);
      double opxi = (1.0 + pc[0]), opeta = (1.0 + pc[1]), optheta = (1.0 + pc[2]);
      double c0[8] = {-ometa * omtheta, ometa * omtheta, opeta * omtheta, -opeta * omtheta,
                      -ometa * optheta, ometa * optheta, opeta * optheta, -opeta * optheta};
      double c1[8] = {-omxi * omtheta, -opxi * omtheta, opxi * omtheta, omxi * omtheta,
                      -omxi * optheta, -opxi * optheta, opxi * optheta, omxi * optheta};
      double c2[8] = {-omxi * ometa, -opxi * ometa, -opxi * opeta, -omxi * opeta,
                      omxi * ometa, opxi * ometa, opxi * opeta, omxi * opeta};
      for (int i = 0; i < 8; i++) { D(i, 0) = c0[i] / 8.0; D(i, 1) = c1[i] / 8.0; D(i, 2) = c2[i] / 8.0; }
      return 0;
    }
    case ET_H20: {
      double x = pc[0], y = pc[1], z = pc[2];
      /* NB the sign flip relative to orc_bfun: FESetModule.jl:1097-1099 */
      double xim = -(-1 + x), etam = -(-1 + y), zetam = -(-1 + z);
      double xip = (1 + x), etap = (1 + y), zetap = (1 + z);
      double twoppp = (2 + x + y + z), twompp = (2 - x + y + z), twopmp = (2 + x - y + z), twoppm = (2 + x + y - z);
      double twommp = (2 - x - y + z), twopmm = (2 + x - y - z), twompm = (2 - x + y - z), twommm = (2 - x - y - z);
      double a[20] = {
          1.0 / 8 * etam * zetam * twoppp - 1.0 / 8 * xim * etam * zetam,
          -1.0 / 8 * etam * zetam * twompp + 1.0 / 8 * xip * etam * zetam,
          -1.0 / 8 * etap * zetam * twommp + 1.0 / 8 * xip * etap * zetam,
          1.0 / 8 * etap * zetam * twopmp - 1.0 / 8 * xim * etap * zetam,
          1.0 / 8 * etam * zetap * twoppm - 1.0 / 8 * xim * etam * zetap,
          -1.0 / 8 * etam * zetap * twompm + 1.0 / 8 * xip * etam * zetap,
          -1.0 / 8 * etap * zetap * twommm + 1.0 / 8 * xip * etap * zetap,
          1.0 / 8 * etap * zetap * twopmm - 1.0 / 8 * xim * etap * zetap,
          -1.0 / 4 * xip * etam * zetam + 1.0 / 4 * xim * etam * zetam,
          1.0 / 4 * etam * etap * zetam,
          -1.0 / 4 * xip * etap * zetam + 1.0 / 4 * xim * etap * zetam,
          -1.0 / 4 * etam * etap * zetam,
          -1.0 / 4 * xip * etam * zetap + 1.0 / 4 * xim * etam * zetap,
          1.0 / 4 * etam * etap * zetap,
          -1.0 / 4 * xip * etap * zetap + 1.0 / 4 * xim * etap * zetap,
          -1.0 / 4 * etam * etap * zetap,
          -1.0 / 4 * zetam * zetap * etam,
          1.0 / 4 * zetam * zetap * etam,
          1.0 / 4 * zetam * zetap * etap,
          -1.0 / 4 * zetam * zetap * etap};
      double b[20] = {
          1.0 / 8 * xim * zetam * twoppp - 1.0 / 8 * xim * etam * zetam,
          1.0 / 8 * xip * zetam * twompp - 1.0 / 8 * xip * etam * zetam,
          -1.0 / 8 * xip * zetam * twommp + 1.0 / 8 * xip * etap * zetam,
          -1.0 / 8 * xim * zetam * twopmp + 1.0 / 8 * xim * etap * zetam,
          1.0 / 8 * xim * zetap * twoppm - 1.0 / 8 * xim * etam * zetap,
          1.0 / 8 * xip * zetap * twompm - 1.0 / 8 * xip * etam * zetap,
          -1.0 / 8 * xip * zetap * twommm + 1.0 / 8 * xip * etap * zetap,
          -1.0 / 8 * xim * zetap * twopmm + 1.0 / 8 * xim * etap * zetap,
          -1.0 / 4 * xim * xip * zetam,
          -1.0 / 4 * xip * etap * zetam + 1.0 / 4 * xip * etam * zetam,
          1.0 / 4 * xim * xip * zetam,
          -1.0 / 4 * xim * etap * zetam + 1.0 / 4 * xim * etam * zetam,
          -1.0 / 4 * xim * xip * zetap,
          -1.0 / 4 * xip * etap * zetap + 1.0 / 4 * xip * etam * zetap,
          1.0 / 4 * xim * xip * zetap,
          -1.0 / 4 * xim * etap * zetap + 1.0 / 4 * xim * etam * zetap,
          -1.0 / 4 * zetam * zetap * xim,
          -1.0 / 4 * zetam * zetap * xip,
          1.0 / 4 * zetam * zetap * xip,
          1.0 / 4 * zetam * zetap * xim};
      double c[20] = {
          1.0 / 8 * xim * etam * twoppp - 1.0 / 8 * xim * etam * zetam,
          1.0 / 8 * xip * etam * twompp - 1.0 / 8 * xip * etam * zetam,
          1.0 / 8 * xip * etap * twommp - 1.0 / 8 * xip * etap * zetam,
          1.0 / 8 * xim * etap * twopmp - 1.0 / 8 * xim * etap * zetam,
          -1.0 / 8 * xim * etam * twoppm + 1.0 / 8 * xim * etam * zetap,
          -1.0 / 8 * xip * etam * twompm + 1.0 / 8 * xip * etam * zetap,
          -1.0 / 8 * xip * etap * twommm + 1.0 / 8 * xip * etap * zetap,
          -1.0 / 8 * xim * etap * twopmm + 1.0 / 8 * xim * etap * zetap,
          -1.0 / 4 * xim * xip * etam,
          -1.0 / 4 * etam * etap * xip,
          -1.0 / 4 * xim * xip * etap,
          -1.0 / 4 * etam * etap * xim,
          1.0 / 4 * xim * xip * etam,
          1.0 / 4 * etam * etap * xip,
          1.0 / 4 * xim * xip * etap,
          1.0 / 4 * etam * etap * xim,
          -1.0 / 4 * xim * etam * zetap + 1.0 / 4 * xim * etam * zetam,
          -1.0 / 4 * xip * etam * zetap + 1.0 / 4 * xip * etam * zetam,
          -1.0 / 4 * xip * etap * zetap + 1.0 / 4 * xip * etap * zetam,
          -1.0 / 4 * xim * etap * zetap + 1.0 / 4 * xim * etap * zetam};
      for (int i = 0; i < 20; i++) { D(i, 0) = a[i]; D(i, 1) = b[i]; D(i, 2) = c[i]; }
      return 0;
    }
    case ET_H27: {
      double xi = pc[0], eta = pc[1], zet = pc[2];
      double x1 = (xi - 1.0 / 2.0), x2 = (xi + 1.0 / 2.0), x3 = (xi - 1.0), x4 = (xi + 1.0);
      double z1 = (zet - 1.0), z2 = (zet - 1.0 / 2.0), z3 = (zet + 1.0), z4 = (zet + 1.0 / 2.0);
      double y1 = (eta - 1.0), y2 = (eta - 1.0 / 2.0), y3 = (eta + 1.0), y4 = (eta + 1.0 / 2.0);
      double v[27][3] = {
          {1.0 / 4.0 * z1 * zet * x1 * y1 * eta, 1.0 / 4.0 * z1 * zet * x3 * xi * y2, 1.0 / 4.0 * z2 * x3 * xi * y1 * eta},
          {1.0 / 4.0 * z1 * zet * x2 * y1 * eta, 1.0 / 4.0 * z1 * zet * x4 * xi * y2, 1.0 / 4.0 * z2 * x4 * xi * y1 * eta},
          {1.0 / 4.0 * z1 * zet * x2 * y3 * eta, 1.0 / 4.0 * z1 * zet * x4 * xi * y4, 1.0 / 4.0 * z2 * x4 * xi * y3 * eta},
          {1.0 / 4.0 * z1 * zet * x1 * y3 * eta, 1.0 / 4.0 * z1 * zet * x3 * xi * y4, 1.0 / 4.0 * z2 * x3 * xi * y3 * eta},
          {1.0 / 4.0 * z3 * zet * x1 * y1 * eta, 1.0 / 4.0 * z3 * zet * x3 * xi * y2, 1.0 / 4.0 * z4 * x3 * xi * y1 * eta},
          {1.0 / 4.0 * z3 * zet * x2 * y1 * eta, 1.0 / 4.0 * z3 * zet * x4 * xi * y2, 1.0 / 4.0 * z4 * x4 * xi * y1 * eta},
          {1.0 / 4.0 * z3 * zet * x2 * y3 * eta, 1.0 / 4.0 * z3 * zet * x4 * xi * y4, 1.0 / 4.0 * z4 * x4 * xi * y3 * eta},
          {1.0 / 4.0 * z3 * zet * x1 * y3 * eta, 1.0 / 4.0 * z3 * zet * x3 * xi * y4, 1.0 / 4.0 * z4 * x3 * xi * y3 * eta},
          {-1.0 / 2.0 * z1 * zet * xi * y1 * eta, 1.0 / 2.0 * z1 * zet * (-x4) * x3 * y2, 1.0 / 2.0 * z2 * (-x4) * x3 * y1 * eta},
          {1.0 / 2.0 * z1 * zet * x2 * (-y3) * y1, -1.0 / 2.0 * z1 * zet * x4 * xi * eta, 1.0 / 2.0 * z2 * x4 * xi * (-y3) * y1},
          {-1.0 / 2.0 * z1 * zet * xi * y3 * eta, 1.0 / 2.0 * z1 * zet * (-x4) * x3 * y4, 1.0 / 2.0 * z2 * (-x4) * x3 * y3 * eta},
          {1.0 / 2.0 * z1 * zet * x1 * (-y3) * y1, -1.0 / 2.0 * z1 * zet * x3 * xi * eta, 1.0 / 2.0 * z2 * x3 * xi * (-y3) * y1},
          {-1.0 / 2.0 * z3 * zet * xi * y1 * eta, 1.0 / 2.0 * z3 * zet * (-x4) * x3 * y2, 1.0 / 2.0 * z4 * (-x4) * x3 * y1 * eta},
          {1.0 / 2.0 * z3 * zet * x2 * (-y3) * y1, -1.0 / 2.0 * z3 * zet * x4 * xi * eta, 1.0 / 2.0 * z4 * x4 * xi * (-y3) * y1},
          {-1.0 / 2.0 * z3 * zet * xi * y3 * eta, 1.0 / 2.0 * z3 * zet * (-x4) * x3 * y4, 1.0 / 2.0 * z4 * (-x4) * x3 * y3 * eta},
          {1.0 / 2.0 * z3 * zet * x1 * (-y3) * y1, -1.0 / 2.0 * z3 * zet * x3 * xi * eta, 1.0 / 2.0 * z4 * x3 * xi * (-y3) * y1},
          {1.0 / 2.0 * (-z3) * z1 * x1 * y1 * eta, 1.0 / 2.0 * (-z3) * z1 * x3 * xi * y2, -1.0 / 2.0 * zet * x3 * xi * y1 * eta},
          {1.0 / 2.0 * (-z3) * z1 * x2 * y1 * eta, 1.0 / 2.0 * (-z3) * z1 * x4 * xi * y2, -1.0 / 2.0 * zet * x4 * xi * y1 * eta},
          {1.0 / 2.0 * (-z3) * z1 * x2 * y3 * eta, 1.0 / 2.0 * (-z3) * z1 * x4 * xi * y4, -1.0 / 2.0 * zet * x4 * xi * y3 * eta},
          {1.0 / 2.0 * (-z3) * z1 * x1 * y3 * eta, 1.0 / 2.0 * (-z3) * z1 * x3 * xi * y4, -1.0 / 2.0 * zet * x3 * xi * y3 * eta},
          {-z1 * zet * xi * (-y3) * y1, -z1 * zet * (-x4) * x3 * eta, z2 * (-x4) * x3 * (-y3) * y1},
          {-(-z3) * z1 * xi * y1 * eta, (-z3) * z1 * (-x4) * x3 * y2, -zet * (-x4) * x3 * y1 * eta},
          {(-z3) * z1 * x2 * (-y3) * y1, -(-z3) * z1 * x4 * xi * eta, -zet * x4 * xi * (-y3) * y1},
          {-(-z3) * z1 * xi * y3 * eta, (-z3) * z1 * (-x4) * x3 * y4, -zet * (-x4) * x3 * y3 * eta},
          {(-z3) * z1 * x1 * (-y3) * y1, -(-z3) * z1 * x3 * xi * eta, -zet * x3 * xi * (-y3) * y1},
          {-z3 * zet * xi * (-y3) * y1, -z3 * zet * (-x4) * x3 * eta, z4 * (-x4) * x3 * (-y3) * y1},
          {-2.0 * (-z3) * z1 * xi * (-y3) * y1, -2.0 * (-z3) * z1 * (-x4) * x3 * eta, -2.0 * zet * (-x4) * x3 * (-y3) * y1}};
      for (int i = 0; i < 27; i++) { D(i, 0) = v[i][0]; D(i, 1) = v[i][1]; D(i, 2) = v[i][2]; }
      return 0;
    }
  }
#undef D
  return -1;
}

/* ------------------------------------------------------------------ rules */
/* 1-D Gauss tables: IntegRuleModule.jl:208-223 (orders 1..4 restated) */
static int gauss1d(int order, double *x, double *w) {
  switch (order) {
    case 1: x[0] = 0.0; w[0] = 2.0; return 0;
    case 2: x[0] = -0.577350269189626; x[1] = 0.577350269189626; w[0] = 1.0; w[1] = 1.0; return 0;
    case 3:
      x[0] = -0.774596669241483; x[1] = 0.0; x[2] = 0.774596669241483;
      w[0] = 0.5555555555555556; w[1] = 0.8888888888888889; w[2] = 0.5555555555555556;
      return 0;
    case 4:
      x[0] = -0.86113631159405; x[1] = -0.33998104358486; x[2] = 0.33998104358486; x[3] = 0.86113631159405;
      w[0] = 0.34785484513745; w[1] = 0.65214515486255; w[2] = 0.65214515486255; w[3] = 0.34785484513745;
      return 0;
  }
  return -1;
}

/* pc[npts x dim] column-major, w[npts]; tensor order i outer .. k inner: IntegRuleModule.jl:354-390 */
ORC_API int orc_gauss_rule(int dim, int order, double *pc, double *w) {
  double x1[8], w1[8];
  if (dim < 1 || dim > 3 || gauss1d(order, x1, w1)) return -1;
  int npts = 1;
  for (int d = 0; d < dim; d++) npts *= order;
  int r = 0;
  if (dim == 1) {
    for (int i = 0; i < order; i++) { pc[i] = x1[i]; w[i] = w1[i]; }
  } else if (dim == 2) {
    for (int i = 0; i < order; i++)
      for (int j = 0; j < order; j++) {
        pc[r] = x1[i]; pc[r + npts] = x1[j];
        w[r] = w1[i] * w1[j];
        r++;
      }
  } else {
    for (int i = 0; i < order; i++)
      for (int j = 0; j < order; j++)
        for (int k = 0; k < order; k++) {
          pc[r] = x1[i]; pc[r + npts] = x1[j]; pc[r + 2 * npts] = x1[k];
          w[r] = w1[i] * w1[j] * w1[k];
          r++;
        }
  }
  return npts;
}

/* IntegRuleModule.jl:483-516 */
ORC_API int orc_tet_rule(int npts, double *pc, double *w) {
  if (npts == 1) {
    pc[0] = 0.25; pc[1] = 0.25; pc[2] = 0.25; w[0] = 1.0 / 6.0;
    return 1;
  } else if (npts == 4) {
    const double a = 0.13819660, b = 0.58541020;
    double p[4][3] = {{a, a, a}, {b, a, a}, {a, b, a}, {a, a, b}};
    for (int i = 0; i < 4; i++) {
      for (int d = 0; d < 3; d++) pc[i + 4 * d] = p[i][d];
      w[i] = 0.041666666666666666667;
    }
    return 4;
  } else if (npts == 5) {
    double a = 1.0 / 6.0, b = 0.25, c = 0.5, d = -0.8, e = 0.45;
    double p[5][3] = {{b, b, b}, {c, a, a}, {a, c, a}, {a, a, c}, {a, a, a}};
    double ww[5] = {d, e, e, e, e};
    for (int i = 0; i < 5; i++) {
      for (int k = 0; k < 3; k++) pc[i + 5 * k] = p[i][k];
      w[i] = ww[i] / 6;
    }
    return 5;
  }
  return -1;
}

/* IntegRuleModule.jl:41-47 */
ORC_API int orc_tri_rule(int npts, double *pc, double *w) {
  if (npts == 1) {
    pc[0] = 1.0 / 3.0; pc[1] = 1.0 / 3.0; w[0] = 1.0 / 2.0;
    return 1;
  } else if (npts == 3) {
    double p[3][2] = {{2.0 / 3, 1.0 / 6}, {1.0 / 6, 2.0 / 3}, {1.0 / 6, 1.0 / 6}};
    for (int i = 0; i < 3; i++) {
      pc[i] = p[i][0]; pc[i + 3] = p[i][1];
      w[i] = (1.0 / 3) / 2;
    }
    return 3;
  }
  return -1;
}

/* ------------------------------------------------------- small dense kernels */
/* C[MxN] = A[KxM]' * B[KxN]   MatrixUtilityModule.jl:478-492 (the @avx loop, taken in plain k order) */
static void mulCAtB(double *C, int M, int N, const double *A, const double *B, int K) {
  for (int n = 0; n < N; n++)
    for (int m = 0; m < M; m++) {
      double Cmn = 0.0;
      for (int k = 0; k < K; k++) Cmn += A[k + (size_t)K * m] * B[k + (size_t)K * n];
      C[m + (size_t)M * n] = Cmn;
    }
}

/* loc[1 x sdim] = N' * X ; J[sdim x mdim] = X' * dN    MatrixUtilityModule.jl:24-68 */
static void locjac(double *loc, double *J, const double *X, const double *N, const double *dN, int nne, int sdim, int mdim) {
  mulCAtB(loc, 1, sdim, N, X, nne);
  mulCAtB(J, sdim, mdim, X, dN, nne);
}

/* FESetModule.jl:479-489 (3-manifold), :426-435 (2-manifold) */
static double jacobian3(const double *J) {
#define Jm(r, c) J[(r - 1) + 3 * (c - 1)]
  return (Jm(1, 1) * (Jm(2, 2) * Jm(3, 3) - Jm(3, 2) * Jm(2, 3)) - Jm(1, 2) * (Jm(2, 1) * Jm(3, 3) - Jm(2, 3) * Jm(3, 1)) +
          Jm(1, 3) * (Jm(2, 1) * Jm(3, 2) - Jm(2, 2) * Jm(3, 1)));
#undef Jm
}
static double jacobian2(const double *J, int sdim) {
  if (sdim == 2) return (J[0] * J[3] - J[1] * J[2]);
  /* norm(cross(J[:,1], J[:,2])) -- LinearAlgebra.norm of a 3-vector: sqrt of sum of squares */
  const double *a = J, *b = J + 3;
  double c0 = a[1] * b[2] - a[2] * b[1];
  double c1 = a[2] * b[0] - a[0] * b[2];
  double c2 = a[0] * b[1] - a[1] * b[0];
  return sqrt(c0 * c0 + c1 * c1 + c2 * c2);
}

/* FESetModule.jl:507-544 */
static void gradN3(double *g, const double *dN, const double *R, int nne) {
#define r_(i, j) R[(i - 1) + 3 * (j - 1)]
  double invdet = 1.0 / (+r_(1, 1) * (r_(2, 2) * r_(3, 3) - r_(3, 2) * r_(2, 3)) - r_(1, 2) * (r_(2, 1) * r_(3, 3) - r_(2, 3) * r_(3, 1)) +
                         r_(1, 3) * (r_(2, 1) * r_(3, 2) - r_(2, 2) * r_(3, 1)));
  double i11 = (r_(2, 2) * r_(3, 3) - r_(3, 2) * r_(2, 3)) * invdet;
  double i12 = -(r_(1, 2) * r_(3, 3) - r_(1, 3) * r_(3, 2)) * invdet;
  double i13 = (r_(1, 2) * r_(2, 3) - r_(1, 3) * r_(2, 2)) * invdet;
  double i21 = -(r_(2, 1) * r_(3, 3) - r_(2, 3) * r_(3, 1)) * invdet;
  double i22 = (r_(1, 1) * r_(3, 3) - r_(1, 3) * r_(3, 1)) * invdet;
  double i23 = -(r_(1, 1) * r_(2, 3) - r_(2, 1) * r_(1, 3)) * invdet;
  double i31 = (r_(2, 1) * r_(3, 2) - r_(3, 1) * r_(2, 2)) * invdet;
  double i32 = -(r_(1, 1) * r_(3, 2) - r_(3, 1) * r_(1, 2)) * invdet;
  double i33 = (r_(1, 1) * r_(2, 2) - r_(2, 1) * r_(1, 2)) * invdet;
#undef r_
  for (int r = 0; r < nne; r++) {
    double a = dN[r], b = dN[r + nne], c = dN[r + 2 * nne];
    g[r] = a * i11 + b * i21 + c * i31;
    g[r + nne] = a * i12 + b * i22 + c * i32;
    g[r + 2 * nne] = a * i13 + b * i23 + c * i33;
  }
}
/* FESetModule.jl:453-470 */
static void gradN2(double *g, const double *dN, const double *R, int nne) {
  double invdet = 1.0 / (R[0] * R[3] - R[2] * R[1]);
  double i11 = (R[3]) * invdet, i12 = -(R[2]) * invdet, i21 = -(R[1]) * invdet, i22 = (R[0]) * invdet;
  for (int r = 0; r < nne; r++) {
    g[r] = dN[r] * i11 + dN[r + nne] * i21;
    g[r + nne] = dN[r] * i12 + dN[r + nne] * i22;
  }
}

/* MatrixUtilityModule.jl:84-98 */
ORC_API void orc_add_mggt_ut_only(double *Ke, const double *gradN, double mult, int nne, int mdim) {
  for (int nx = 0; nx < nne; nx++)
    for (int px = 0; px < mdim; px++) {
      double a = (mult)*gradN[nx + nne * px];
      for (int mx = 0; mx <= nx; mx++) Ke[mx + nne * nx] += gradN[mx + nne * px] * a;
    }
}

/* MatrixUtilityModule.jl:120-153; scratch kg is mdim x nne */
ORC_API void orc_add_gkgt_ut_only(double *Ke, const double *gradN, double Jac_w, const double *kappa, double *kg, int nne, int mdim) {
  for (int nx = 0; nx < nne; nx++)
    for (int mx = 0; mx < mdim; mx++) {
      double accum = 0.0;
      for (int px = 0; px < mdim; px++) accum += kappa[mx + mdim * px] * gradN[nx + nne * px];
      kg[mx + mdim * nx] = Jac_w * accum;
    }
  for (int nx = 0; nx < nne; nx++)
    for (int mx = 0; mx <= nx; mx++) {
      double accum = 0.0;
      for (int px = 0; px < mdim; px++) accum += gradN[mx + nne * px] * kg[px + mdim * nx];
      Ke[mx + nne * nx] += accum;
    }
}

/* MatrixUtilityModule.jl:189-216; B is nstr x K, D nstr x nstr, DB scratch nstr x K */
ORC_API void orc_add_btdb_ut_only(double *Ke, const double *B, double Jac_w, const double *D, double *DB, int nstr, int K) {
  for (int nx = 0; nx < K; nx++)
    for (int mx = 0; mx < nstr; mx++) {
      double accum = 0.0;
      for (int px = 0; px < nstr; px++) accum += D[mx + nstr * px] * B[px + nstr * nx];
      DB[mx + nstr * nx] = Jac_w * accum;
    }
  for (int nx = 0; nx < K; nx++)
    for (int mx = 0; mx <= nx; mx++) {
      double accum = 0.0;
      for (int px = 0; px < nstr; px++) accum += B[px + nstr * mx] * DB[px + nstr * nx];
      Ke[mx + (size_t)K * nx] += accum;
    }
}

/* MatrixUtilityModule.jl:164-173 */
ORC_API void orc_complete_lt(double *Ke, int n) {
  for (int nx = 0; nx < n; nx++)
    for (int mx = nx + 1; mx < n; mx++) Ke[mx + (size_t)n * nx] = Ke[nx + (size_t)n * mx];
}

/* DeforModelRedModule.jl:447-472 with Rm = identity handed in explicitly (6 x 3*nne) */
static void blmat3d(double *B, const double *g, const double *Rm, int nne) {
  memset(B, 0, sizeof(double) * 6 * 3 * nne);
#define Rm_(i, j) Rm[(i - 1) + 3 * (j - 1)]
  for (int i = 0; i < nne; i++) {
    double g1 = g[i], g2 = g[i + nne], g3 = g[i + 2 * nne];
    for (int j = 1; j <= 3; j++) {
      double *b = B + 6 * (3 * i + (j - 1));
      b[0] = g1 * Rm_(j, 1);
      b[1] = g2 * Rm_(j, 2);
      b[2] = g3 * Rm_(j, 3);
      b[3] = g2 * Rm_(j, 1) + g1 * Rm_(j, 2);
      b[4] = g3 * Rm_(j, 1) + g1 * Rm_(j, 3);
      b[5] = g3 * Rm_(j, 2) + g2 * Rm_(j, 3);
    }
  }
#undef Rm_
}

/* --------------------------------------------------------------- assembler */
/* AssemblyModule.jl:250-282: column-major walk, three stores per entry, range checks.
 * Returns 0, or the error code of the first violated check:
 *   1 "Column degree of freedom < 1"  2 "Column degree of freedom > size"
 *   3 "Row degree of freedom < 1"     4 "Row degree of freedom > size"            */
static int assemble(int64_t *I, int64_t *J, double *V, int64_t *p, const double *mat, const int64_t *dr, int nr, const int64_t *dc,
                    int nc, int64_t row_nall, int64_t col_nall) {
  int64_t q = *p;
  for (int j = 0; j < nc; j++) {
    int64_t dj = dc[j];
    if (dj < 1) return 1;
    if (dj > col_nall) return 2;
    for (int i = 0; i < nr; i++) {
      int64_t di = dr[i];
      if (di < 1) return 3;
      if (di > row_nall) return 4;
      V[q] = mat[i + (size_t)nr * j];
      I[q] = di;
      J[q] = dj;
      q++;
    }
  }
  *p = q;
  return 0;
}

ORC_API int orc_assemble(int64_t *I, int64_t *J, double *V, int64_t *p, const double *mat, const int64_t *dr, int nr,
                         const int64_t *dc, int nc, int64_t row_nall, int64_t col_nall) {
  return assemble(I, J, V, p, mat, dr, nr, dc, nc, row_nall, col_nall);
}

/* ------------------------------------------------------------------ forms */
typedef struct {
  int et, nne, mdim, sdim, ndn, npts;
  int64_t nelem, nnodes, nalldofs;
  const int64_t *conn;    /* [nelem][nne] 1-based */
  const double *xyz;      /* nnodes x sdim col-major */
  const int64_t *dofnums; /* nnodes x ndn col-major */
  double *Ns, *dNs;       /* per point: nne, nne*mdim */
  const double *w;
} formctx;

static int form_setup(formctx *f, int et, int64_t nelem, const int64_t *conn, int64_t nnodes, int sdim, const double *xyz, int ndn,
                      const int64_t *dofnums, int64_t nalldofs, int npts, const double *pc, const double *w) {
  f->et = et; f->nne = orc_nne(et); f->mdim = orc_mdim(et); f->sdim = sdim; f->ndn = ndn; f->npts = npts;
  f->nelem = nelem; f->nnodes = nnodes; f->nalldofs = nalldofs; f->conn = conn; f->xyz = xyz; f->dofnums = dofnums; f->w = w;
  if (f->nne < 0) return -1;
  /* integrationdata: IntegDomainModule.jl:631-648 */
  f->Ns = (double *)malloc(sizeof(double) * npts * f->nne);
  f->dNs = (double *)malloc(sizeof(double) * npts * f->nne * f->mdim);
  for (int j = 0; j < npts; j++) {
    double p[3] = {0, 0, 0};
    for (int d = 0; d < f->mdim; d++) p[d] = pc[j + (size_t)npts * d];
    orc_bfun(et, p, f->Ns + (size_t)j * f->nne);
    orc_bfundpar(et, p, f->dNs + (size_t)j * f->nne * f->mdim);
  }
  return 0;
}
static void form_free(formctx *f) { free(f->Ns); free(f->dNs); }

/* gathervalues_asmat! FieldModule.jl:263-275 ; gatherdofnums! :304-314 */
static void gather_elem(const formctx *f, int64_t e, double *ecoords, int64_t *dofs) {
  const int64_t *c = f->conn + e * f->nne;
  for (int i = 0; i < f->nne; i++)
    for (int j = 0; j < f->sdim; j++) ecoords[i + f->nne * j] = f->xyz[(c[i] - 1) + f->nnodes * j];
  int en = 0;
  for (int i = 0; i < f->nne; i++)
    for (int j = 0; j < f->ndn; j++) dofs[en++] = f->dofnums[(c[i] - 1) + f->nnodes * j];
}

static const double IDENT3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
static const double IDENT2[4] = {1, 0, 0, 1};

/* bilform_diffusion: FEMMBaseModule.jl:1462-1535.  kappa_kind 0 = scalar (_iso :1508), 1 = mdim x mdim (_general :1476).
 * Emits nelem*nne^2 triplets into I,J,V (reference emission order).  If elmats != NULL also stores element matrices. */
ORC_API int orc_bilform_diffusion(int et, int64_t nelem, const int64_t *conn, int64_t nnodes, int sdim, const double *xyz,
                                  const int64_t *dofnums, int64_t nalldofs, int npts, const double *pc, const double *w,
                                  int kappa_kind, const double *kappa, double otherdim, const double *Rm, int64_t *I, int64_t *J,
                                  double *V) {
  formctx f;
  if (form_setup(&f, et, nelem, conn, nnodes, sdim, xyz, 1, dofnums, nalldofs, npts, pc, w)) return -1;
  int nne = f.nne, mdim = f.mdim;
  if (sdim != mdim) { form_free(&f); return -2; }
  double ecoords[27 * 3], loc[3], Jm[9], RmTJ[9], gradN[27 * 3], kg[3 * 27];
  double *elmat = (double *)malloc(sizeof(double) * nne * nne);
  int64_t dofs[27];
  int64_t p = 0;
  int rc = 0;
  for (int64_t i = 0; i < nelem && !rc; i++) {
    gather_elem(&f, i, ecoords, dofs);
    memset(elmat, 0, sizeof(double) * nne * nne);
    for (int j = 0; j < npts; j++) {
      const double *N = f.Ns + (size_t)j * nne, *dN = f.dNs + (size_t)j * nne * mdim;
      locjac(loc, Jm, ecoords, N, dN, nne, sdim, mdim);
      /* Jacobianvolume: IntegDomainModule.jl:567 (3-manifold), :504 (2-manifold x the constant otherdimension) */
      double Jac = (mdim == 3) ? jacobian3(Jm) : jacobian2(Jm, sdim) * otherdim;
      if (kappa_kind == 1) {
        /* mulCAtB!(RmTJ, csmat(self.mcsys), J) FEMMBaseModule.jl:1496; csmat: the constant mcsys matrix (NULL = identity) */
        mulCAtB(RmTJ, mdim, mdim, Rm ? Rm : ((mdim == 3) ? IDENT3 : IDENT2), Jm, mdim);
        if (mdim == 3) gradN3(gradN, dN, RmTJ, nne); else gradN2(gradN, dN, RmTJ, nne);
        orc_add_gkgt_ut_only(elmat, gradN, (Jac * w[j]), kappa, kg, nne, mdim);
      } else {
        if (mdim == 3) gradN3(gradN, dN, Jm, nne); else gradN2(gradN, dN, Jm, nne);
        orc_add_mggt_ut_only(elmat, gradN, (kappa[0] * Jac * w[j]), nne, mdim);
      }
    }
    orc_complete_lt(elmat, nne);
    rc = assemble(I, J, V, &p, elmat, dofs, nne, dofs, nne, nalldofs, nalldofs);
  }
  free(elmat);
  form_free(&f);
  return rc;
}

/* bilform_lin_elastic with DeforModelRed3D: FEMMBaseModule.jl:1774-1813.  C is 6x6 col-major. */
ORC_API int orc_bilform_lin_elastic(int et, int64_t nelem, const int64_t *conn, int64_t nnodes, int sdim, const double *xyz,
                                    const int64_t *dofnums, int64_t nalldofs, int npts, const double *pc, const double *w,
                                    const double *C, const double *Rm, int64_t *I, int64_t *J, double *V) {
  formctx f;
  if (form_setup(&f, et, nelem, conn, nnodes, sdim, xyz, 3, dofnums, nalldofs, npts, pc, w)) return -1;
  int nne = f.nne, mdim = f.mdim;
  if (sdim != 3 || mdim != 3) { form_free(&f); return -2; }
  int K = 3 * nne;
  double ecoords[27 * 3], loc[3], Jm[9], RmTJ[9], gradN[27 * 3];
  double *B = (double *)malloc(sizeof(double) * 6 * K), *DB = (double *)malloc(sizeof(double) * 6 * K);
  double *elmat = (double *)malloc(sizeof(double) * K * K);
  int64_t dofs[81];
  int64_t p = 0;
  int rc = 0;
  for (int64_t i = 0; i < nelem && !rc; i++) {
    gather_elem(&f, i, ecoords, dofs);
    memset(elmat, 0, sizeof(double) * K * K);
    for (int j = 0; j < npts; j++) {
      const double *N = f.Ns + (size_t)j * nne, *dN = f.dNs + (size_t)j * nne * mdim;
      locjac(loc, Jm, ecoords, N, dN, nne, sdim, mdim);
      double Jac = jacobian3(Jm);
      mulCAtB(RmTJ, 3, 3, Rm ? Rm : IDENT3, Jm, 3); /* At_mul_B!(RmTJ, csmat, J) :1802 */
      gradN3(gradN, dN, RmTJ, nne);
      blmat3d(B, gradN, Rm ? Rm : IDENT3, nne);
      orc_add_btdb_ut_only(elmat, B, Jac * w[j], C, DB, 6, K);
    }
    orc_complete_lt(elmat, K);
    rc = assemble(I, J, V, &p, elmat, dofs, K, dofs, K, nalldofs, nalldofs);
  }
  free(B); free(DB); free(elmat);
  form_free(&f);
  return rc;
}

/* bilform_dot: FEMMBaseModule.jl:1335-1366.  c is ndn x ndn col-major; m = manifold dimension of the Jacobian;
 * otherdim = constant "other dimension" (IntegDomainModule.jl:150-152 gives 1.0) */
ORC_API int orc_bilform_dot(int et, int64_t nelem, const int64_t *conn, int64_t nnodes, int sdim, const double *xyz, int ndn,
                            const int64_t *dofnums, int64_t nalldofs, int npts, const double *pc, const double *w, const double *c,
                            int m, double otherdim, int64_t *I, int64_t *J, double *V) {
  formctx f;
  if (form_setup(&f, et, nelem, conn, nnodes, sdim, xyz, ndn, dofnums, nalldofs, npts, pc, w)) return -1;
  int nne = f.nne, mdim = f.mdim;
  int K = ndn * nne;
  if ((mdim == 3 && m != 3) || (mdim == 2 && (m < 2 || m > 3))) { form_free(&f); return -3; } /* IntegDomainModule.jl:543,600 */
  double ecoords[27 * 3], loc[3], Jm[9];
  double *elmat = (double *)malloc(sizeof(double) * K * K);
  int64_t *dofs = (int64_t *)malloc(sizeof(int64_t) * K);
  int64_t p = 0;
  int rc = 0;
  for (int64_t i = 0; i < nelem && !rc; i++) {
    gather_elem(&f, i, ecoords, dofs);
    memset(elmat, 0, sizeof(double) * K * K);
    for (int j = 0; j < npts; j++) {
      const double *N = f.Ns + (size_t)j * nne, *dN = f.dNs + (size_t)j * nne * mdim;
      locjac(loc, Jm, ecoords, N, dN, nne, sdim, mdim);
      double Jac;
      if (mdim == 3) Jac = jacobian3(Jm);
      else { Jac = jacobian2(Jm, sdim); if (m == 3) Jac = Jac * otherdim; }
      for (int k = 0; k < nne; k++)
        for (int mm = 0; mm < nne; mm++) {
          double factor = (N[k] * N[mm] * Jac * w[j]);
          for (int pp = 0; pp < ndn; pp++)
            for (int q = 0; q < ndn; q++) elmat[(k * ndn + pp) + (size_t)K * (mm * ndn + q)] += factor * c[pp + ndn * q];
        }
    }
    rc = assemble(I, J, V, &p, elmat, dofs, K, dofs, K, nalldofs, nalldofs);
  }
  free(elmat); free(dofs);
  form_free(&f);
  return rc;
}

/* bilform_convection: FEMMBaseModule.jl:1583-1625.  Scalar field Q (1 dof/node), convective velocity u given at the nodes
 * (uvals nnodes x nsd col-major, nsd = sdim), rho constant.  Non-symmetric: the full element matrix is formed. */
ORC_API int orc_bilform_convection(int et, int64_t nelem, const int64_t *conn, int64_t nnodes, int sdim, const double *xyz,
                                   const double *uvals, int nsd, const int64_t *dofnums, int64_t nalldofs, int npts, const double *pc,
                                   const double *w, double rho, double otherdim, int64_t *I, int64_t *J, double *V) {
  formctx f;
  if (form_setup(&f, et, nelem, conn, nnodes, sdim, xyz, 1, dofnums, nalldofs, npts, pc, w)) return -1;
  int nne = f.nne, mdim = f.mdim;
  if (sdim != mdim || nsd != sdim) { form_free(&f); return -2; }
  double ecoords[27 * 3], eus[27 * 3], loc[3], Jm[9], gradN[27 * 3];
  double *elmat = (double *)malloc(sizeof(double) * nne * nne);
  int64_t dofs[27];
  int64_t p = 0;
  int rc = 0;
  (void)rho; /* the reference evaluates rhof but never uses the value: FEMMBaseModule.jl:1606-1617 */
  for (int64_t i = 0; i < nelem && !rc; i++) {
    gather_elem(&f, i, ecoords, dofs);
    const int64_t *c = f.conn + i * nne;
    for (int a = 0; a < nne; a++)
      for (int s = 0; s < nsd; s++) eus[a + nne * s] = uvals[(c[a] - 1) + nnodes * s]; /* gathervalues_asmat!(u, eus, conn) :1599 */
    memset(elmat, 0, sizeof(double) * nne * nne);
    for (int j = 0; j < npts; j++) {
      const double *N = f.Ns + (size_t)j * nne, *dN = f.dNs + (size_t)j * nne * mdim;
      locjac(loc, Jm, ecoords, N, dN, nne, sdim, mdim);
      double Jac = (mdim == 3) ? jacobian3(Jm) : jacobian2(Jm, sdim) * otherdim;
      if (mdim == 3) gradN3(gradN, dN, Jm, nne); else gradN2(gradN, dN, Jm, nne);
      for (int pp = 0; pp < nne; pp++)
        for (int r = 0; r < nne; r++) {
          double accum = 0.0;
          for (int s = 0; s < nsd; s++) {
            double u_s = 0.0;
            for (int q = 0; q < nne; q++) u_s += N[q] * eus[q + nne * s];
            accum += u_s * gradN[r + nne * s];
          }
          elmat[pp + (size_t)nne * r] += N[pp] * accum * (Jac * w[j]);
        }
    }
    rc = assemble(I, J, V, &p, elmat, dofs, nne, dofs, nne, nalldofs, nalldofs);
  }
  free(elmat);
  form_free(&f);
  return rc;
}

/* bilform_div_grad: FEMMBaseModule.jl:1672-1713.  Vector field u with ndn = sdim dofs per node, constant viscosity mu. */
ORC_API int orc_bilform_div_grad(int et, int64_t nelem, const int64_t *conn, int64_t nnodes, int sdim, const double *xyz, int ndn,
                                 const int64_t *dofnums, int64_t nalldofs, int npts, const double *pc, const double *w, double mu,
                                 double otherdim, int64_t *I, int64_t *J, double *V) {
  formctx f;
  if (form_setup(&f, et, nelem, conn, nnodes, sdim, xyz, ndn, dofnums, nalldofs, npts, pc, w)) return -1;
  int nne = f.nne, mdim = f.mdim;
  if (sdim != mdim || ndn != sdim) { form_free(&f); return -2; }
  int K = ndn * nne;
  double ecoords[27 * 3], loc[3], Jm[9], gradN[27 * 3];
  double *elmat = (double *)malloc(sizeof(double) * K * K);
  int64_t dofs[81];
  int64_t p = 0;
  int rc = 0;
  for (int64_t i = 0; i < nelem && !rc; i++) {
    gather_elem(&f, i, ecoords, dofs);
    memset(elmat, 0, sizeof(double) * K * K);
    for (int j = 0; j < npts; j++) {
      const double *N = f.Ns + (size_t)j * nne, *dN = f.dNs + (size_t)j * nne * mdim;
      locjac(loc, Jm, ecoords, N, dN, nne, sdim, mdim);
      double Jac = (mdim == 3) ? jacobian3(Jm) : jacobian2(Jm, sdim) * otherdim;
      if (mdim == 3) gradN3(gradN, dN, Jm, nne); else gradN2(gradN, dN, Jm, nne);
      double factor = mu * (Jac * w[j]);
      for (int a = 0; a < nne; a++)
        for (int b = 0; b < nne; b++)
          for (int s = 0; s < ndn; s++) {
            int pr = a * ndn + s, r = b * ndn + s;
            for (int q = 0; q < ndn; q++) elmat[pr + (size_t)K * r] += factor * gradN[a + nne * q] * gradN[b + nne * q];
            for (int q = 0; q < ndn; q++) {
              r = b * ndn + q;
              elmat[pr + (size_t)K * r] += factor * gradN[a + nne * q] * gradN[b + nne * s];
            }
          }
    }
    rc = assemble(I, J, V, &p, elmat, dofs, K, dofs, K, nalldofs, nalldofs);
  }
  free(elmat);
  form_free(&f);
  return rc;
}

/* bilform_masslike: FEMMBaseModule.jl:1865-1912.  Rows are numbered by element: element i owns rows (i-1)*ndn+1 .. i*ndn;
 * c is ndn x ndn col-major; emits nelem * ndn * (nne*ndn) triplets. */
ORC_API int orc_bilform_masslike(int et, int64_t nelem, const int64_t *conn, int64_t nnodes, int sdim, const double *xyz, int ndn,
                                 const int64_t *dofnums, int64_t nalldofs, int npts, const double *pc, const double *w, const double *c,
                                 int m, double otherdim, int64_t *I, int64_t *J, double *V) {
  formctx f;
  if (form_setup(&f, et, nelem, conn, nnodes, sdim, xyz, ndn, dofnums, nalldofs, npts, pc, w)) return -1;
  int nne = f.nne, mdim = f.mdim;
  int elrows = ndn, elcols = nne * ndn;
  if ((mdim == 3 && m != 3) || (mdim == 2 && (m < 2 || m > 3))) { form_free(&f); return -3; }
  double ecoords[27 * 3], loc[3], Jm[9];
  double *elmat = (double *)malloc(sizeof(double) * elrows * elcols);
  int64_t *dofs = (int64_t *)malloc(sizeof(int64_t) * elcols);
  int64_t rowdofs[6];
  int64_t p = 0;
  int rc = 0;
  for (int64_t i = 0; i < nelem && !rc; i++) {
    gather_elem(&f, i, ecoords, dofs);
    memset(elmat, 0, sizeof(double) * elrows * elcols);
    for (int j = 0; j < npts; j++) {
      const double *N = f.Ns + (size_t)j * nne, *dN = f.dNs + (size_t)j * nne * mdim;
      locjac(loc, Jm, ecoords, N, dN, nne, sdim, mdim);
      double Jac;
      if (mdim == 3) Jac = jacobian3(Jm);
      else { Jac = jacobian2(Jm, sdim); if (m == 3) Jac = Jac * otherdim; }
      for (int b = 0; b < nne; b++) {
        double factor = N[b] * Jac * w[j];
        for (int pp = 0; pp < ndn; pp++)
          for (int q = 0; q < ndn; q++) elmat[pp + (size_t)elrows * (b * ndn + q)] += factor * c[pp + ndn * q];
      }
    }
    for (int r = 0; r < elrows; r++) rowdofs[r] = i * elrows + r + 1; /* collect(((i-1)*elrows + 1):(i*elrows)) */
    rc = assemble(I, J, V, &p, elmat, rowdofs, elrows, dofs, elcols, nelem * elrows, nalldofs);
  }
  free(elmat); free(dofs);
  form_free(&f);
  return rc;
}

/* linform_dot (= distribloads with a constant ForceIntensity): FEMMBaseModule.jl:1207-1244, SysvecAssembler
 * AssemblyModule.jl:853-917.  force is ndn values; F has nalldofs entries (zeroed here like startassembly!). */
ORC_API int orc_linform_dot(int et, int64_t nelem, const int64_t *conn, int64_t nnodes, int sdim, const double *xyz, int ndn,
                            const int64_t *dofnums, int64_t nalldofs, int npts, const double *pc, const double *w, const double *force,
                            int m, double otherdim, double *F) {
  formctx f;
  if (form_setup(&f, et, nelem, conn, nnodes, sdim, xyz, ndn, dofnums, nalldofs, npts, pc, w)) return -1;
  int nne = f.nne, mdim = f.mdim;
  int K = ndn * nne;
  if ((mdim == 3 && m != 3) || (mdim == 2 && (m < 2 || m > 3))) { form_free(&f); return -3; }
  double ecoords[27 * 3], loc[3], Jm[9];
  double *elvec = (double *)malloc(sizeof(double) * K);
  int64_t *dofs = (int64_t *)malloc(sizeof(int64_t) * K);
  for (int64_t k = 0; k < nalldofs; k++) F[k] = 0.0;
  int rc = 0;
  for (int64_t i = 0; i < nelem && !rc; i++) {
    gather_elem(&f, i, ecoords, dofs);
    for (int k = 0; k < K; k++) elvec[k] = 0.0;
    for (int j = 0; j < npts; j++) {
      const double *N = f.Ns + (size_t)j * nne, *dN = f.dNs + (size_t)j * nne * mdim;
      locjac(loc, Jm, ecoords, N, dN, nne, sdim, mdim);
      double Jac;
      if (mdim == 3) Jac = jacobian3(Jm);
      else { Jac = jacobian2(Jm, sdim); if (m == 3) Jac = Jac * otherdim; }
      double Factor = (Jac * w[j]);
      int rx = 0;
      for (int kx = 0; kx < nne; kx++) {
        double NkxF = N[kx] * Factor;
        for (int mx = 0; mx < ndn; mx++) { elvec[rx] = elvec[rx] + NkxF * force[mx]; rx++; }
      }
    }
    for (int k = 0; k < K && !rc; k++) { /* assemble!(::SysvecAssembler, vec, dofnums) :899-906 */
      int64_t gi = dofs[k];
      if (gi < 1) rc = 3;
      else if (gi > nalldofs) rc = 4;
      else F[gi - 1] += elvec[k];
    }
  }
  free(elvec); free(dofs);
  form_free(&f);
  return rc;
}

/* ------------------------------------------------------ sparse(I,J,V,m,n) */
/* Restatement of Julia's SparseArrays.sparse!(I,J,V,m,n,+) (stdlib, pinned by Julia ^1.12, not vendored in the
 * reference; call site AssemblyModule.jl:319-325).  Published algorithm (after Tim Davis' CSparse / HALFPERM):
 *   1. count entries per row, 2. counting-sort the triplets into CSR keeping input order inside a row,
 *   3. sweep each row, folding repeated columns into the first occurrence with + (left to right) and counting
 *      the surviving entries per column, 4. prefix-sum the column counts, 5. transpose CSR -> CSC row by row,
 *      which leaves row indices strictly increasing inside each column.  Explicit zeros are kept.
 * Returns nnz (>= 0), or -1 for an out-of-range index.  colptr has n+1 entries, rowval/nzval need capacity ntrip
 * (call with rowval == NULL to get nnz only).  All indices 1-based. */
ORC_API int64_t orc_sparse(int64_t ntrip, const int64_t *I, const int64_t *J, const double *V, int64_t m, int64_t n, int64_t *colptr,
                           int64_t *rowval, double *nzval) {
  int64_t *rowptr = (int64_t *)calloc((size_t)m + 2, sizeof(int64_t));
  int64_t *ccol = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ntrip > 0 ? ntrip : 1));
  double *cval = (double *)malloc(sizeof(double) * (size_t)(ntrip > 0 ? ntrip : 1));
  int64_t *klast = (int64_t *)calloc((size_t)n + 1, sizeof(int64_t));
  int64_t *cptr = (int64_t *)calloc((size_t)n + 2, sizeof(int64_t));
  int64_t nnz = -1;
  for (int64_t k = 0; k < ntrip; k++) {
    if (I[k] < 1 || I[k] > m || J[k] < 1 || J[k] > n) goto done;
    rowptr[I[k] + 1]++;
  }
  /* rowptr[i+1] := first slot of row i (1-based rows); shifted by one so the scatter can bump it */
  {
    int64_t acc = 0;
    for (int64_t i = 1; i <= m + 1; i++) { int64_t c = rowptr[i]; rowptr[i] = acc; acc += c; }
  }
  for (int64_t k = 0; k < ntrip; k++) {
    int64_t q = rowptr[I[k] + 1]++;
    ccol[q] = J[k];
    cval[q] = V[k];
  }
  /* now rowptr[i] = start of row i, rowptr[i+1] = end  (i = 1..m) */
  {
    int64_t w = 0; /* write cursor of the compacted CSR */
    for (int64_t i = 1; i <= m; i++) {
      int64_t start = rowptr[i], stop = rowptr[i + 1], newstart = w;
      for (int64_t k = start; k < stop; k++) {
        int64_t j = ccol[k];
        if (klast[j] > newstart) { /* column j already seen in this row: fold */
          cval[klast[j] - 1] = cval[klast[j] - 1] + cval[k];
        } else {
          ccol[w] = j;
          cval[w] = cval[k];
          w++;
          klast[j] = w; /* position + 1 */
          cptr[j + 1]++;
        }
      }
      rowptr[i] = newstart;
    }
    rowptr[m + 1] = w;
    nnz = w;
  }
  colptr[0] = 1;
  for (int64_t j = 1; j <= n; j++) colptr[j] = colptr[j - 1] + cptr[j + 1];
  if (rowval) {
    for (int64_t j = 1; j <= n; j++) cptr[j] = colptr[j - 1] - 1; /* 0-based fill cursors */
    for (int64_t i = 1; i <= m; i++)
      for (int64_t k = rowptr[i]; k < rowptr[i + 1]; k++) {
        int64_t q = cptr[ccol[k]]++;
        rowval[q] = i;
        nzval[q] = cval[k];
      }
  }
done:
  free(rowptr); free(ccol); free(cval); free(klast); free(cptr);
  return nnz;
}
