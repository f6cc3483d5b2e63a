/*
 * ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of `kepler.solve(M, ecc)` from the third-party package
 * kepler.py == 0.0.7 (dfm/kepler.py; pinned in the reference at
 * requirements.txt:143 and pyproject.toml:21).  The package is NOT vendored in
 * /root/reference and cannot be installed here (no network), so this file
 * restates its published algorithm:
 *
 *   1. wrap M into [0, 2pi) with NumPy/Python remainder semantics,
 *   2. reflect M > pi to 2pi - M,
 *   3. Markley (1995) cubic starter,
 *   4. ONE high-order (Nijenhuis 1991 style) refinement that uses
 *      E - sin E and 1 - cos E from a 10-term nested series,
 *   5. undo the reflection.
 *
 * PARITY STATUS: "parity unpinned" — the reference's own tests never call
 * kepler.solve with fixed inputs; the solver is pinned
 * here by its residual |E - e sin E - M| instead (tests/test_oracle.py::test_kepler_residual_grid).
 *
 * Reference call sites this stands in for:
 *   support/models/kep00.model:6, kep01.model:13, kep02.model:20,
 *   kep03.model:5, kep04.model:14, kep06.model:8, kep07.model:16,
 *   akep00.model:5, emp_model.py:1325.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: one IEEE rounding per
 * written operation, so results do not depend on the host's FMA support).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#include <math.h>
#include <stdint.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#ifndef M_PI_2
#define M_PI_2 1.57079632679489661923
#endif
#ifndef M_PI_4
#define M_PI_4 0.78539816339744830962
#endif

/* NumPy `np.mod` / Python `%` for doubles: result carries the divisor's sign. */
static double py_mod(double a, double b) {
  double m = fmod(a, b);
  if (b == 0.0) return m;
  if (m != 0.0) {
    if ((b < 0.0) != (m < 0.0)) m += b;
  } else {
    m = copysign(0.0, b);
  }
  return m;
}

/* x - sin(x) and 1 - cos(x) for x in [0, pi], nested 10-term series after
 * folding x into [0, pi/4] (Nijenhuis 1991 reduction). */
static void sin_cos_reduc(double x, double *sn_reduc, double *cs_reduc) {
  static const double s[10] = {1.0 / 6,   1.0 / 20,  1.0 / 42,  1.0 / 72,  1.0 / 110,
                               1.0 / 156, 1.0 / 210, 1.0 / 272, 1.0 / 342, 1.0 / 420};
  static const double c[10] = {0.5,       1.0 / 12,  1.0 / 30,  1.0 / 56,  1.0 / 90,
                               1.0 / 132, 1.0 / 182, 1.0 / 240, 1.0 / 306, 1.0 / 380};
  int bigg = x > M_PI_2;
  double u = bigg ? M_PI - x : x;
  int big = u > M_PI_4;
  double v = big ? M_PI_2 - u : u;
  double w = v * v;

  double ss = 1.0, cc = 1.0;
  for (int i = 9; i >= 1; --i) {
    ss = 1.0 - w * s[i] * ss;
    cc = 1.0 - w * c[i] * cc;
  }
  ss *= v * w * s[0];
  cc *= w * c[0];

  double sn, cs;
  if (big) {
    sn = u - 1.0 + cc;
    cs = 1.0 - M_PI_2 + u + ss;
  } else {
    sn = ss;
    cs = cc;
  }
  if (bigg) {
    sn = 2.0 * x - M_PI + sn;
    cs = 2.0 - cs;
  }
  *sn_reduc = sn;
  *cs_reduc = cs;
}

/* Markley (1995) starter; M in [0, pi]. */
static double markley_starter(double M, double ecc, double ome) {
  const double FACTOR1 = 3.0 * M_PI / (M_PI - 6.0 / M_PI);
  const double FACTOR2 = 1.6 / (M_PI - 6.0 / M_PI);
  double M2 = M * M;
  double alpha = FACTOR1 + FACTOR2 * (M_PI - M) / (1.0 + ecc);
  double d = 3.0 * ome + alpha * ecc;
  double alphad = alpha * d;
  double r = (3.0 * alphad * (d - ome) + M2) * M;
  double q = 2.0 * alphad * ome - M2;
  double q2 = q * q;
  double w = pow(fabs(r) + sqrt(q2 * q + r * r), 2.0 / 3.0);
  return (2.0 * r * w / (w * w + w * q + q2) + M) / d;
}

static double refine_estimate(double M, double ecc, double ome, double E) {
  double sE, cE;
  sin_cos_reduc(E, &sE, &cE); /* sE = E - sin E, cE = 1 - cos E */
  double f_0 = ecc * sE + E * ome - M;
  double f_1 = ecc * cE + ome;
  double f_2 = ecc * (E - sE);
  double f_3 = 1.0 - f_1;
  double d_3 = -f_0 / (f_1 - 0.5 * f_0 * f_2 / f_1);
  double d_4 = -f_0 / (f_1 + 0.5 * d_3 * f_2 + (d_3 * d_3) * f_3 / 6.0);
  double d_42 = d_4 * d_4;
  double dE = -f_0 / (f_1 + 0.5 * d_4 * f_2 + d_4 * d_4 * f_3 / 6.0 - d_42 * d_4 * f_2 / 24.0);
  return E + dE;
}

double emp_oracle_kepler_solve1(double M, double ecc) {
  const double two_pi = 2.0 * M_PI;
  M = py_mod(M, two_pi);
  int high = M > M_PI;
  if (high) M = two_pi - M;
  double ome = 1.0 - ecc;
  double E = markley_starter(M, ecc, ome);
  E = refine_estimate(M, ecc, ome, E);
  if (high) E = two_pi - E;
  return E;
}

/* Vectorised entry: E[i] = solve(M[i], ecc[i]). */
void emp_oracle_kepler_solve(const double *M, const double *ecc, double *E, int64_t n) {
  for (int64_t i = 0; i < n; ++i) E[i] = emp_oracle_kepler_solve1(M[i], ecc[i]);
}

/* Diagnostics used by tests: the un-refined Markley starter on the folded M. */
double emp_oracle_kepler_starter(double M, double ecc) {
  const double two_pi = 2.0 * M_PI;
  M = py_mod(M, two_pi);
  if (M > M_PI) M = two_pi - M;
  return markley_starter(M, ecc, 1.0 - ecc);
}
