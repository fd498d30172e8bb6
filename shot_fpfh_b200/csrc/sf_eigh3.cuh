// 3x3 symmetric eigen-decomposition that follows LAPACK's dsyevd code path for n = 3 step by step, so that the
// eigenvector SIGNS come out as np.linalg.eigh returns them (shot.py:36).
//
// Why the signs matter: the reference disambiguates an axis by flipping it when strictly more neighbours project
// negatively than non-negatively (shot.py:40-45), and the query point itself always projects to exactly 0, which
// counts as non-negative. When the other neighbours split evenly (|#neg - #pos| <= 1, 5-9 % of the queries at
// K ~ 70-100) neither v nor -v gets flipped, and the frame the reference returns is whatever sign LAPACK happened
// to produce. An eigen-solver with another sign convention (Jacobi, analytic) disagrees on those queries by a
// mirrored x/y axis — a completely different descriptor. Hence this port.
//
// Path taken by numpy.linalg.eigh(a) (UPLO='L', jobz='V') for a 3x3 matrix, LAPACK >= 3.10 (OpenBLAS >= 0.3.20,
// i.e. both the NumPy 1.26.4 wheels the reference pins and the NumPy 2.3 of this image):
//   dsyevd -> dsytrd -> dsytd2('L')      one Householder reflector H (dlarfg) on rows/cols 2..3
//          -> dstedc('I') -> dsteqr('I') implicit QL / QR with Wilkinson shift, dlaev2 on 2x2 blocks, dlartg
//                                        rotations applied by dlasr, final selection sort (ascending)
//          -> dormtr('L','L','N')        Z := H * Z on rows 2..3
// No scaling branch is taken for |a| in (1e-146, 1e+146) (dsyevd) / (1e-122, 1e+153) (dsteqr); inputs outside
// that range are scaled into it first, which LAPACK also does (by a different constant — irrelevant to signs).
#pragma once
#include <math.h>

#ifndef SF_HD
#if defined(__CUDACC__)
#define SF_HD __host__ __device__ __forceinline__
#else
#define SF_HD inline
#endif
#endif

namespace sf {
namespace lapack3 {

SF_HD double sign(double a, double b) { return copysign(fabs(a), b); }  // Fortran SIGN, signed zero included

SF_HD double dlapy2(double x, double y) {
  const double xa = fabs(x), ya = fabs(y);
  const double w = fmax(xa, ya), z = fmin(xa, ya);
  if (z == 0.0) return w;
  const double q = z / w;
  return w * sqrt(1.0 + q * q);
}

// LAPACK 3.10+ dlartg (la_lartg.f90): c >= 0, r carries the sign of f.
SF_HD void dlartg(double f, double g, double& c, double& s, double& r) {
  const double safmin = 2.2250738585072014e-308, safmax = 4.4942328371557898e+307;
  const double rtmin = 1.4916681462400413e-154, rtmax = 4.7403759540545887e+153;
  const double f1 = fabs(f), g1 = fabs(g);
  if (g == 0.0) {
    c = 1.0; s = 0.0; r = f;
  } else if (f == 0.0) {
    c = 0.0; s = sign(1.0, g); r = g1;
  } else if (f1 > rtmin && f1 < rtmax && g1 > rtmin && g1 < rtmax) {
    const double d = sqrt(f * f + g * g);
    c = f1 / d;
    r = sign(d, f);
    s = g / r;
  } else {
    const double u = fmin(safmax, fmax(safmin, fmax(f1, g1)));
    const double fs = f / u, gs = g / u;
    const double d = sqrt(fs * fs + gs * gs);
    c = fabs(fs) / d;
    r = sign(d, f);
    s = gs / r;
    r = r * u;
  }
}

// dlaev2: eigen-decomposition of [[a, b], [b, c]]; rt1 has the larger absolute value, (cs1, sn1) is its vector.
SF_HD void dlaev2(double a, double b, double c, double& rt1, double& rt2, double& cs1, double& sn1) {
  const double sm = a + c, df = a - c, adf = fabs(df), tb = b + b, ab = fabs(tb);
  double acmx, acmn;
  if (fabs(a) > fabs(c)) { acmx = a; acmn = c; } else { acmx = c; acmn = a; }
  double rt;
  if (adf > ab) { const double q = ab / adf; rt = adf * sqrt(1.0 + q * q); }
  else if (adf < ab) { const double q = adf / ab; rt = ab * sqrt(1.0 + q * q); }
  else rt = ab * sqrt(2.0);
  int sgn1;
  if (sm < 0.0) { rt1 = 0.5 * (sm - rt); sgn1 = -1; rt2 = (acmx / rt1) * acmn - (b / rt1) * b; }
  else if (sm > 0.0) { rt1 = 0.5 * (sm + rt); sgn1 = 1; rt2 = (acmx / rt1) * acmn - (b / rt1) * b; }
  else { rt1 = 0.5 * rt; rt2 = -0.5 * rt; sgn1 = 1; }
  int sgn2;
  double cs;
  if (df >= 0.0) { cs = df + rt; sgn2 = 1; } else { cs = df - rt; sgn2 = -1; }
  const double acs = fabs(cs);
  if (acs > ab) {
    const double ct = -tb / cs;
    sn1 = 1.0 / sqrt(1.0 + ct * ct);
    cs1 = ct * sn1;
  } else if (ab == 0.0) {
    cs1 = 1.0; sn1 = 0.0;
  } else {
    const double tn = -cs / tb;
    cs1 = 1.0 / sqrt(1.0 + tn * tn);
    sn1 = tn * cs1;
  }
  if (sgn1 == sgn2) { const double tn = cs1; cs1 = -sn1; sn1 = tn; }
}

// dlasr(SIDE='R', PIVOT='V', DIRECT=forward?'F':'B') on the 3 x ncols block of z starting at column col0 (0-based),
// rotations (cw[j], sw[j]), j = 0..ncols-2.
SF_HD void dlasr_rv(bool forward, int ncols, const double* cw, const double* sw, double z[3][3], int col0) {
  for (int t = 0; t < ncols - 1; ++t) {
    const int j = forward ? t : ncols - 2 - t;
    const double ct = cw[j], st = sw[j];
    if (ct != 1.0 || st != 0.0) {
      for (int i = 0; i < 3; ++i) {
        const double temp = z[i][col0 + j + 1];
        z[i][col0 + j + 1] = ct * temp - st * z[i][col0 + j];
        z[i][col0 + j] = st * temp + ct * z[i][col0 + j];
      }
    }
  }
}

// dsteqr(COMPZ='I') for n = 3, transcribed loop for loop (run-time indices into small arrays). Kept as the statement of
// what dsteqr3 below must compute — tests/host_math compares the two bit for bit —; the kernels call dsteqr3.
// d[0..2], e[0..1] are overwritten; z receives the eigenvectors in COLUMNS.
SF_HD void dsteqr3_generic(double* d, double* e, double z[3][3]) {
  const int n = 3;
  const double eps = 1.1102230246251565e-16, eps2 = eps * eps, safmin = 2.2250738585072014e-308;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) z[i][j] = i == j ? 1.0 : 0.0;
  double work_c[2], work_s[2];
  const int nmaxit = n * 30;
  int jtot = 0;
  // 1-based indices as in the Fortran source; D(i) = d[i-1], E(i) = e[i-1].
#define D_(i) d[(i)-1]
#define E_(i) e[(i)-1]
  int l1 = 1;
  const int nm1 = n - 1;
  while (true) {  // label 10
    if (l1 > n) break;
    if (l1 > 1) E_(l1 - 1) = 0.0;
    int m = n;
    if (l1 <= nm1) {
      for (int mm = l1; mm <= nm1; ++mm) {
        const double tst = fabs(E_(mm));
        if (tst == 0.0) { m = mm; break; }
        if (tst <= (sqrt(fabs(D_(mm))) * sqrt(fabs(D_(mm + 1)))) * eps) { E_(mm) = 0.0; m = mm; break; }
      }
    }
    int l = l1;
    const int lsv = l;
    int lend = m;
    const int lendsv = lend;
    l1 = m + 1;
    if (lend == l) continue;
    // (no scaling: see the header comment)
    double anorm = 0.0;
    for (int i = l; i <= lend; ++i) anorm = fmax(anorm, fabs(D_(i)));
    for (int i = l; i < lend; ++i) anorm = fmax(anorm, fabs(E_(i)));
    if (anorm == 0.0) continue;
    if (fabs(D_(lend)) < fabs(D_(l))) { lend = lsv; l = lendsv; }
    if (lend > l) {
      // ---- QL iteration ----
      while (true) {  // label 40
        int mq = lend;
        if (l != lend) {
          for (int mm = l; mm <= lend - 1; ++mm) {
            const double tst = fabs(E_(mm)) * fabs(E_(mm));
            if (tst <= (eps2 * fabs(D_(mm))) * fabs(D_(mm + 1)) + safmin) { mq = mm; break; }
          }
        }
        if (mq < lend) E_(mq) = 0.0;
        double p = D_(l);
        if (mq == l) {  // label 80: eigenvalue found
          D_(l) = p;
          l = l + 1;
          if (l <= lend) continue;
          break;
        }
        if (mq == l + 1) {
          double rt1, rt2, c, s;
          dlaev2(D_(l), E_(l), D_(l + 1), rt1, rt2, c, s);
          work_c[0] = c; work_s[0] = s;
          dlasr_rv(false, 2, work_c, work_s, z, l - 1);
          D_(l) = rt1; D_(l + 1) = rt2; E_(l) = 0.0;
          l = l + 2;
          if (l <= lend) continue;
          break;
        }
        if (jtot == nmaxit) break;
        ++jtot;
        double g = (D_(l + 1) - p) / (2.0 * E_(l));
        double r = dlapy2(g, 1.0);
        g = D_(mq) - p + (E_(l) / (g + sign(r, g)));
        double s = 1.0, c = 1.0;
        p = 0.0;
        for (int i = mq - 1; i >= l; --i) {
          const double f = s * E_(i), b = c * E_(i);
          dlartg(g, f, c, s, r);
          if (i != mq - 1) E_(i + 1) = r;
          g = D_(i + 1) - p;
          r = (D_(i) - g) * s + 2.0 * c * b;
          p = s * r;
          D_(i + 1) = g + p;
          g = c * r - b;
          work_c[i - l] = c;      // WORK(I), rebased to the block start L
          work_s[i - l] = -s;     // WORK(N-1+I)
        }
        dlasr_rv(false, mq - l + 1, work_c, work_s, z, l - 1);
        D_(l) = D_(l) - p;
        E_(l) = g;
      }
    } else {
      // ---- QR iteration ----
      while (true) {  // label 90
        int mq = lend;
        if (l != lend) {
          for (int mm = l; mm >= lend + 1; --mm) {
            const double tst = fabs(E_(mm - 1)) * fabs(E_(mm - 1));
            if (tst <= (eps2 * fabs(D_(mm))) * fabs(D_(mm - 1)) + safmin) { mq = mm; break; }
          }
        }
        if (mq > lend) E_(mq - 1) = 0.0;
        double p = D_(l);
        if (mq == l) {  // label 130
          D_(l) = p;
          l = l - 1;
          if (l >= lend) continue;
          break;
        }
        if (mq == l - 1) {
          double rt1, rt2, c, s;
          dlaev2(D_(l - 1), E_(l - 1), D_(l), rt1, rt2, c, s);
          work_c[0] = c; work_s[0] = s;
          dlasr_rv(true, 2, work_c, work_s, z, l - 2);
          D_(l - 1) = rt1; D_(l) = rt2; E_(l - 1) = 0.0;
          l = l - 2;
          if (l >= lend) continue;
          break;
        }
        if (jtot == nmaxit) break;
        ++jtot;
        double g = (D_(l - 1) - p) / (2.0 * E_(l - 1));
        double r = dlapy2(g, 1.0);
        g = D_(mq) - p + (E_(l - 1) / (g + sign(r, g)));
        double s = 1.0, c = 1.0;
        p = 0.0;
        for (int i = mq; i <= l - 1; ++i) {
          const double f = s * E_(i), b = c * E_(i);
          dlartg(g, f, c, s, r);
          if (i != mq) E_(i - 1) = r;
          g = D_(i) - p;
          r = (D_(i + 1) - g) * s + 2.0 * c * b;
          p = s * r;
          D_(i) = g + p;
          g = c * r - b;
          work_c[i - mq] = c;     // WORK(I), rebased to the block start M
          work_s[i - mq] = s;
        }
        dlasr_rv(true, l - mq + 1, work_c, work_s, z, mq - 1);
        D_(l) = D_(l) - p;
        E_(l - 1) = g;
      }
    }
    if (jtot >= nmaxit) break;
  }
  // selection sort, ascending (label 160)
  for (int ii = 2; ii <= n; ++ii) {
    const int i = ii - 1;
    int k = i;
    double p = D_(i);
    for (int j = ii; j <= n; ++j)
      if (D_(j) < p) { k = j; p = D_(j); }
    if (k != i) {
      D_(k) = D_(i);
      D_(i) = p;
      for (int r = 0; r < 3; ++r) { const double t = z[r][i - 1]; z[r][i - 1] = z[r][k - 1]; z[r][k - 1] = t; }
    }
  }
#undef D_
#undef E_
}

// ---- the same computation with every index resolved at compile time ---------------------------------------------
// dsteqr's control flow for n = 3 has few shapes: the matrix splits (or not) at a negligible off-diagonal entry into
// blocks of 1, 2 or 3; a 3-block is iterated from the top (QL) or from the bottom (QR) until its first eigenvalue
// separates, which leaves a 2-block; a 2-block is solved by dlaev2 (identically in QL and QR). Written out per shape,
// D, E and Z are plain scalars: in registers on the device, where the transcription above kept them in local memory
// (430 local loads/stores per thread of lrf_eigen_kernel, the largest stall of that latency-bound kernel).
struct Steqr3 {
  double d1, d2, d3, e1, e2;
  double z11, z21, z31, z12, z22, z32, z13, z23, z33;  // z<row><column>
  int jtot;
};

// One plane rotation of dlasr(SIDE='R', PIVOT='V') on columns (a, b) = (j, j + 1), given by reference.
SF_HD void rot2(double ct, double st, double& a1, double& a2, double& a3, double& b1, double& b2, double& b3) {
  if (ct != 1.0 || st != 0.0) {
    const double t1 = b1, t2 = b2, t3 = b3;
    b1 = ct * t1 - st * a1; a1 = st * t1 + ct * a1;
    b2 = ct * t2 - st * a2; a2 = st * t2 + ct * a2;
    b3 = ct * t3 - st * a3; a3 = st * t3 + ct * a3;
  }
}
#define SF_ROT12(t, c, s) rot2(c, s, t.z11, t.z21, t.z31, t.z12, t.z22, t.z32)
#define SF_ROT23(t, c, s) rot2(c, s, t.z12, t.z22, t.z32, t.z13, t.z23, t.z33)

SF_HD bool ql_small(double e, double da, double db) {  // the deflation test inside the QL / QR iterations
  const double eps = 1.1102230246251565e-16, eps2 = eps * eps, safmin = 2.2250738585072014e-308;
  const double tst = fabs(e) * fabs(e);
  return tst <= (eps2 * fabs(da)) * fabs(db) + safmin;
}

// A 2-block on rows/columns (1,2) or (2,3): dlaev2 + one rotation (what both the QL and the QR iteration do with it).
// `bottom_up`: the block is iterated as QR (|d_last| < |d_first|), which only changes the order of the factors in the test.
SF_HD void block2_12(Steqr3& t, bool test_first, bool bottom_up = false) {
  if (test_first && (bottom_up ? ql_small(t.e1, t.d2, t.d1) : ql_small(t.e1, t.d1, t.d2))) { t.e1 = 0.0; return; }
  double rt1, rt2, c, s;
  dlaev2(t.d1, t.e1, t.d2, rt1, rt2, c, s);
  SF_ROT12(t, c, s);
  t.d1 = rt1; t.d2 = rt2; t.e1 = 0.0;
}
SF_HD void block2_23(Steqr3& t, bool test_first, bool bottom_up = false) {
  if (test_first && (bottom_up ? ql_small(t.e2, t.d3, t.d2) : ql_small(t.e2, t.d2, t.d3))) { t.e2 = 0.0; return; }
  double rt1, rt2, c, s;
  dlaev2(t.d2, t.e2, t.d3, rt1, rt2, c, s);
  SF_ROT23(t, c, s);
  t.d2 = rt1; t.d3 = rt2; t.e2 = 0.0;
}

// The 3-block iterated from the top (QL, l = 1, lend = 3).
SF_HD void block3_ql(Steqr3& t) {
  const int nmaxit = 90;
  while (true) {  // l = 1
    int mq = 3;
    if (ql_small(t.e1, t.d1, t.d2)) mq = 1;
    else if (ql_small(t.e2, t.d2, t.d3)) mq = 2;
    if (mq == 1) {  // first eigenvalue found: l = 2, the 2-block (2,3) remains (its own test comes first)
      t.e1 = 0.0;
      block2_23(t, true);
      return;
    }
    if (mq == 2) {  // 2-block (1,2) by dlaev2, then l = 3 = lend: done
      t.e2 = 0.0;
      block2_12(t, false);
      return;
    }
    if (t.jtot == nmaxit) return;
    ++t.jtot;
    double p = t.d1;
    double g = (t.d2 - p) / (2.0 * t.e1);
    double r = dlapy2(g, 1.0);
    g = t.d3 - p + (t.e1 / (g + sign(r, g)));
    double s = 1.0, c = 1.0;
    p = 0.0;
    // i = 2
    double f = s * t.e2, b = c * t.e2;
    dlartg(g, f, c, s, r);
    g = t.d3 - p;
    r = (t.d2 - g) * s + 2.0 * c * b;
    p = s * r;
    t.d3 = g + p;
    g = c * r - b;
    const double c2 = c, s2 = -s;
    // i = 1
    f = s * t.e1; b = c * t.e1;
    dlartg(g, f, c, s, r);
    t.e2 = r;
    g = t.d2 - p;
    r = (t.d1 - g) * s + 2.0 * c * b;
    p = s * r;
    t.d2 = g + p;
    g = c * r - b;
    const double c1 = c, s1 = -s;
    // dlasr backward over the three columns: rotation 2 on (2,3), then rotation 1 on (1,2)
    SF_ROT23(t, c2, s2);
    SF_ROT12(t, c1, s1);
    t.d1 = t.d1 - p;
    t.e1 = g;
  }
}

// The 3-block iterated from the bottom (QR, l = 3, lend = 1).
SF_HD void block3_qr(Steqr3& t) {
  const int nmaxit = 90;
  while (true) {  // l = 3
    int mq = 1;
    if (ql_small(t.e2, t.d3, t.d2)) mq = 3;
    else if (ql_small(t.e1, t.d2, t.d1)) mq = 2;
    if (mq == 3) {  // last eigenvalue found: l = 2, the 2-block (1,2) remains (its own test comes first)
      t.e2 = 0.0;
      if (ql_small(t.e1, t.d2, t.d1)) { t.e1 = 0.0; return; }
      block2_12(t, false);
      return;
    }
    if (mq == 2) {  // 2-block (2,3) by dlaev2, then l = 1 = lend: done
      t.e1 = 0.0;
      block2_23(t, false);
      return;
    }
    if (t.jtot == nmaxit) return;
    ++t.jtot;
    double p = t.d3;
    double g = (t.d2 - p) / (2.0 * t.e2);
    double r = dlapy2(g, 1.0);
    g = t.d1 - p + (t.e2 / (g + sign(r, g)));
    double s = 1.0, c = 1.0;
    p = 0.0;
    // i = 1
    double f = s * t.e1, b = c * t.e1;
    dlartg(g, f, c, s, r);
    g = t.d1 - p;
    r = (t.d2 - g) * s + 2.0 * c * b;
    p = s * r;
    t.d1 = g + p;
    g = c * r - b;
    const double c1 = c, s1 = s;
    // i = 2
    f = s * t.e2; b = c * t.e2;
    dlartg(g, f, c, s, r);
    t.e1 = r;
    g = t.d2 - p;
    r = (t.d3 - g) * s + 2.0 * c * b;
    p = s * r;
    t.d2 = g + p;
    g = c * r - b;
    const double c2 = c, s2 = s;
    // dlasr forward over the three columns: rotation 1 on (1,2), then rotation 2 on (2,3)
    SF_ROT12(t, c1, s1);
    SF_ROT23(t, c2, s2);
    t.d3 = t.d3 - p;
    t.e2 = g;
  }
}

// dsteqr(COMPZ='I') for n = 3. d[0..2], e[0..1] are overwritten; z receives the eigenvectors in COLUMNS.
SF_HD void dsteqr3(double* d_io, double* e_io, double z[3][3]) {
  const double eps = 1.1102230246251565e-16;
  Steqr3 t;
  t.d1 = d_io[0]; t.d2 = d_io[1]; t.d3 = d_io[2]; t.e1 = e_io[0]; t.e2 = e_io[1];
  t.z11 = 1.0; t.z21 = 0.0; t.z31 = 0.0; t.z12 = 0.0; t.z22 = 1.0; t.z32 = 0.0; t.z13 = 0.0; t.z23 = 0.0; t.z33 = 1.0;
  t.jtot = 0;
  // the split of the matrix into blocks (label 10): the first negligible off-diagonal entry ends a block
  const double a1 = fabs(t.e1), a2 = fabs(t.e2);
  bool split1 = a1 == 0.0;
  if (!split1 && a1 <= (sqrt(fabs(t.d1)) * sqrt(fabs(t.d2))) * eps) { t.e1 = 0.0; split1 = true; }
  bool split2 = a2 == 0.0;
  if (!split2 && a2 <= (sqrt(fabs(t.d2)) * sqrt(fabs(t.d3))) * eps) { t.e2 = 0.0; split2 = true; }
  // (blocks whose largest entry is 0 are skipped: anorm == 0)
  if (!split1 && !split2) {
    if (fabs(t.d3) < fabs(t.d1)) block3_qr(t); else block3_ql(t);   // anorm > 0: e1 != 0
  } else if (split1 && !split2) {
    t.e1 = 0.0;
    block2_23(t, true, fabs(t.d3) < fabs(t.d2));  // blocks (1) and (2,3); anorm > 0: e2 != 0
  } else if (!split1 && split2) {
    block2_12(t, true, fabs(t.d2) < fabs(t.d1));  // blocks (1,2) and (3)
    t.e2 = 0.0;
  } else {
    t.e1 = 0.0; t.e2 = 0.0;
  }
  // selection sort, ascending (label 160), for n = 3
  {
    int k = 1;
    double p = t.d1;
    if (t.d2 < p) { k = 2; p = t.d2; }
    if (t.d3 < p) { k = 3; p = t.d3; }
    if (k == 2) {
      t.d2 = t.d1; t.d1 = p;
      double w = t.z11; t.z11 = t.z12; t.z12 = w;
      w = t.z21; t.z21 = t.z22; t.z22 = w;
      w = t.z31; t.z31 = t.z32; t.z32 = w;
    } else if (k == 3) {
      t.d3 = t.d1; t.d1 = p;
      double w = t.z11; t.z11 = t.z13; t.z13 = w;
      w = t.z21; t.z21 = t.z23; t.z23 = w;
      w = t.z31; t.z31 = t.z33; t.z33 = w;
    }
    if (t.d3 < t.d2) {
      const double q = t.d3; t.d3 = t.d2; t.d2 = q;
      double w = t.z12; t.z12 = t.z13; t.z13 = w;
      w = t.z22; t.z22 = t.z23; t.z23 = w;
      w = t.z32; t.z32 = t.z33; t.z33 = w;
    }
  }
  d_io[0] = t.d1; d_io[1] = t.d2; d_io[2] = t.d3; e_io[0] = t.e1; e_io[1] = t.e2;
  z[0][0] = t.z11; z[0][1] = t.z12; z[0][2] = t.z13;
  z[1][0] = t.z21; z[1][1] = t.z22; z[1][2] = t.z23;
  z[2][0] = t.z31; z[2][1] = t.z32; z[2][2] = t.z33;
}
#undef SF_ROT12
#undef SF_ROT23

}  // namespace lapack3

// The decomposition in two halves, so that a kernel can regroup its problems between them (lrf_eigen_kernel sorts the
// problems of a block by the branch dsteqr is about to take: threads of one warp then run the same code).
//   eigh3_tridiagonal : scaling + dsytd2('L')  -> t.d, t.e, the reflector (tau, v2), the scale
//   eigh3_finish      : dsteqr('I') + dormtr   -> eval ascending, evec[c][.] = eigenvector of eval[c]
struct Tridiagonal3 {
  double d[3], e[2], tau, v2, scale;
};

// m = {a11, a21, a31, a22, a32, a33} (the lower triangle, which is what LAPACK reads with UPLO='L').
SF_HD void eigh3_tridiagonal(const double m_in[6], Tridiagonal3& t) {
  using namespace lapack3;
  double m[6];
  double amax = 0.0;
  for (int i = 0; i < 6; ++i) amax = fmax(amax, fabs(m_in[i]));
  const double scale = (amax > 0.0 && (amax < 1e-100 || amax > 1e100)) ? 1.0 / amax : 1.0;
  for (int i = 0; i < 6; ++i) m[i] = m_in[i] * scale;
  const double a11 = m[0], a21 = m[1], a31 = m[2], a22 = m[3], a32 = m[4], a33 = m[5];
  // ---- dsytd2('L'), i = 1: dlarfg(2, a21, a31) -----------------------------------------------------------
  double tau = 0.0, v2 = 0.0;
  t.d[0] = a11;
  double b22 = a22, b32 = a32, b33 = a33;
  const double xnorm = fabs(a31);
  if (xnorm == 0.0) {
    t.e[0] = a21;
  } else {
    const double beta = -sign(dlapy2(a21, xnorm), a21);
    tau = (beta - a21) / beta;
    v2 = a31 * (1.0 / (a21 - beta));  // dscal by 1/(alpha - beta); v = (1, v2)
    t.e[0] = beta;
    // x := tau * A22 * v (dsymv, lower) ; alpha := -1/2 tau (x . v) ; w := x + alpha v ; A22 -= v w^T + w v^T
    const double x1 = tau * (a22 + a32 * v2);
    const double x2 = tau * (a32 + a33 * v2);
    const double alpha = -0.5 * tau * (x1 + x2 * v2);
    const double w1 = x1 + alpha, w2 = x2 + alpha * v2;
    b22 = a22 - (w1 + w1);
    b32 = a32 - (v2 * w1 + w2);
    b33 = a33 - (v2 * w2 + w2 * v2);
  }
  t.d[1] = b22;
  t.e[1] = b32;  // i = 2: dlarfg(1, ...) -> tau = 0
  t.d[2] = b33;
  t.tau = tau;
  t.v2 = v2;
  t.scale = scale;
}

// True when dsteqr will work on the whole matrix from the bottom up (QR iteration), false from the top down (QL) —
// for the common case of no negligible off-diagonal entry; only used to GROUP problems, never to decide anything.
SF_HD bool eigh3_bottom_up(const Tridiagonal3& t) { return fabs(t.d[2]) < fabs(t.d[0]); }

SF_HD void eigh3_finish(Tridiagonal3& t, double eval[3], double evec[3][3]) {
  using namespace lapack3;
  // ---- dsteqr('I') ---------------------------------------------------------------------------------------
  double z[3][3];
  dsteqr3(t.d, t.e, z);
  // ---- dormtr: rows 2..3 of Z := H * rows 2..3, H = I - tau v v^T -------------------------------------------
  if (t.tau != 0.0) {
    for (int c = 0; c < 3; ++c) {
      const double s = t.tau * (z[1][c] + t.v2 * z[2][c]);
      z[1][c] -= s;
      z[2][c] -= s * t.v2;
    }
  }
  for (int c = 0; c < 3; ++c) {
    eval[c] = t.d[c] / t.scale;
    for (int k = 0; k < 3; ++k) evec[c][k] = z[k][c];
  }
}

// eval ascending; evec[c][.] = eigenvector of eval[c] (np.linalg.eigh's column c), LAPACK's sign.
SF_HD void eigh3(const double m_in[6], double eval[3], double evec[3][3]) {
  Tridiagonal3 t;
  eigh3_tridiagonal(m_in, t);
  eigh3_finish(t, eval, evec);
}

}  // namespace sf
