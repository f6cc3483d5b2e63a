// emp_device.cuh — device-side building blocks of the EMPEROR hot path (sm_100a).
//
// What the reference computes per (walker, datapoint) — SURVEY.md §8a rows A3-A9:
//   M = freq*t + phase ; E = kepler.solve(M, e) ; f = 2 atan(sqrt((1+e)/(1-e)) tan(E/2)) ;
//   model += A (cos(f+w) + e cos w)                     support/models/kep00.model:4-8
// Here (B200-first, FP64 CUDA cores, no tensor cores — this is not a contraction):
//   * M is formed with the reference's two roundings and reduced mod 2pi EXACTLY
//     (rint + one FMA), so the solver sees bit-identical input to NumPy's fmod path;
//   * default solver (EMP_SOLVER_GRID): the grid-anchored core below (kep_grid_a / kep_grid_c) —
//     FP32 starter, sin/cos of the nearest grid point E_h = k 2^-7 from a shared-memory table, one FP32
//     Halley step in delta-space, one FP64 Newton correction.  Same root as kepler.solve to ~1 ulp
//     (Kepler's equation has one root), ~half the FP64 instructions of kepler.py's refinement;
//   * EMP_SOLVER_KEPLERPY and eccentricities above kGridEccMax: the Markley starter and the single
//     high-order refinement of kepler.py (oracle/kepler_oracle.c), `kepler_refined`, evaluated with FMAs;
//   * the RV term is evaluated as [b1 cos E + a2 sin E] / (1 - e cos E), b1 = A cos w (1 - e^2),
//     a2 = -A sin w sqrt(1 - e^2): algebraically identical to the template's tan/atan/cos chain.
// Element-wise parity of both solvers: tests/test_kepler_gpu.py; SASS evidence: tests/test_sass_evidence.py
// and profiles/r02_logl_sass.txt.
#pragma once
#include <stdint.h>
#include "../../include/emperor_b200.h"

namespace emp {

constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 2.0 * 3.14159265358979323846;  // == np.float64(2*np.pi)
constexpr double kPi2 = 1.57079632679489661923;
constexpr double kPi4 = 0.78539816339744830962;
constexpr double kInvTwoPi = 0.15915494309189533577;
constexpr double kMagic = 6755399441055744.0;  // 1.5 * 2^52: (x + kMagic) - kMagic == rint(x)
// Markley (1995) constants, as in kepler.py
constexpr double kF1 = 3.0 * kPi / (kPi - 6.0 / kPi);
constexpr double kF2 = 1.6 / (kPi - 6.0 / kPi);

// Grid-anchored Kepler core (likelihood kernel v6): sin/cos of the grid points
// k * 2^-7, k = 0..511 (covers [0, 4) > pi), as correctly rounded FP64 pairs and FP32 pairs.
constexpr int kGridN = 512;
constexpr double kGridEccMax = 0.98;  // beyond it the walker/planet takes the kepler.py-style refinement

// Per-(walker, planet) starter table (likelihood kernel v10).  The eccentricity is a constant of the walker, so
// E(M; e) on [0, pi] is ONE smooth curve per planet that every one of the N datapoints looks up: it is tabulated
// once per walker in the kernel prologue (kStartN cells of pi/kStartN, a quadratic through the cell's two edges and
// its centre: |error| <= 7e-4 for e <= 0.8, 1.5e-4 for e <= 0.7 with 52 cells — the Markley starter's own error is
// 4e-4; what the refinement needs is |delta| < ~6e-3 around the grid point, see kep_grid_a) and replaces the Markley
// starter's 17 FP32 + 4 MUFU instructions per point by 5 FP32 instructions and 3 shared-memory reads.  Amortised
// over N >= 2k points the build (105 FP32 solves per planet) is < 1 % of the walker's work.  Measured on C4
// (profiles/r02_variants.log): 6.28 -> 5.62 ms per launch; 5 planets x 52 cells is what fits next to the tile ring
// with two CTAs per SM (64 cells x 5 planets drops to one CTA per SM: 6.74 ms).
#ifndef EMP_START_N
#define EMP_START_N 52
#endif
#ifndef EMP_START_PLANETS
#define EMP_START_PLANETS 5
#endif
constexpr int kStartN = EMP_START_N;    // cells; nodes 0 .. kStartN (<= 127: the node index is masked with 127)
constexpr int kStartStride = kStartN + 1;
constexpr int kStartPlanets = EMP_START_PLANETS;  // planets per walker that get a table (shared-memory budget)
constexpr float kStartEccMax = 0.8f;
constexpr int kStartFloats = kStartPlanets * 3 * kStartStride;  // per walker

// Hot-loop FP64 literals travel in the KERNEL PARAMETER bank (c[0x0]) so that DFMA/DADD take them
// as a direct constant operand: a 64-bit immediate costs two UMOVs per use and a user
// __constant__ array (bank 3) costs an LDC per use (profiles/r02_logl_sass.txt shows the c[0x0][..] operands).
//   (v - sin v)/v^3 = 1/3! - w/5! + w^2/7! - ...   (8 terms: next term < 1e-18 relative on [0, pi/4])
//   (1 - cos v)/v^2 = 1/2! - w/4! + w^2/6! - ...   (9 terms), both stored highest degree first
struct HotConsts {
  double sinc[8];
  double cosc[9];
  double c[10];  // [0] pi [1] 2pi [2] pi/2 [3] pi/4 [4] 1/(2pi) [5] rint magic [6] F1 [7] 1/6 [8] 1/24 [9] 1-pi/2
  double g[4];   // grid core: [0] 1/120 [1] 1/720 [2] 2^45 [3] spare
  double ex[13];  // exp_neg: [0] log2(e) [1] -ln2_hi [2] -ln2_lo [3..12] 1/2! .. 1/11!
};
__host__ __device__ inline HotConsts make_hot_consts() {
  HotConsts h = {{-1.0 / 355687428096000.0, 1.0 / 1307674368000.0, -1.0 / 6227020800.0, 1.0 / 39916800.0,
                  -1.0 / 362880.0, 1.0 / 5040.0, -1.0 / 120.0, 1.0 / 6.0},
                 {1.0 / 6402373705728000.0, -1.0 / 20922789888000.0, 1.0 / 87178291200.0, -1.0 / 479001600.0,
                  1.0 / 3628800.0, -1.0 / 40320.0, 1.0 / 720.0, -1.0 / 24.0, 0.5},
                 {kPi, kTwoPi, kPi2, kPi4, kInvTwoPi, kMagic, kF1, 1.0 / 6.0, 1.0 / 24.0, 1.0 - kPi2},
                 {1.0 / 120.0, 1.0 / 720.0, 35184372088832.0, 0.0},
                 {1.4426950408889634074, -6.93147180369123816490e-01, -1.90821492927058770002e-10, 1.0 / 2.0,
                  1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0, 1.0 / 40320.0, 1.0 / 362880.0,
                  1.0 / 3628800.0, 1.0 / 39916800.0}};
  return h;
}

// Per (walker, Keplerian) constants, computed once in the kernel prologue.
struct KepConst {
  double freq;    // 2 pi / per
  double tpv;     // t_p for kep03/kep04, else 0:   M = freq*(t - tpv) + phv reproduces both
  double phv;     // phase, or 0 for kep03/kep04    templates' roundings (x - 0 and x + 0 are exact)
  double e;       // eccentricity
  double ome;     // 1 - e
  double c2;      // kF2 / (1 + e)
  double ome3;    // 3 (1 - e)
  double a1;      // A cos w
  double a2;      // -A sin w sqrt(1 - e^2)
  double a3;      // A e cos w
  double b1;      // A cos w (1 - e^2): RV = (b1 cos E + a2 sin E) / (1 - e cos E)   (grid core)
  float ef, omef, c2f, ome3f;  // FP32 copies for the starter
  float ef3, om23f, c2f3, efh; // e/3, 2(1-e)/3, 3 c2, e/2: constant factors folded for the grid core
  int slow_mod;                // |M| may exceed 1e12 somewhere in the data set: use the fmod path
  int robust;                  // e outside [0, kGridEccMax]: the grid-anchored core is not used
  int tab;                     // the walker's starter table for this planet was built (e <= kStartEccMax)
  int _pad;
};
constexpr int kKepConstDoubles = sizeof(KepConst) / sizeof(double);

// ---- parameter transforms: support/models/kep0{0,1,2,3,4,6,7}.model, akep00.model ----
__device__ inline void kep_elements(int model, const double* th, double& per, double& A, double& ph,
                                    double& e, double& w, bool& use_tp) {
  use_tp = false;
  double S = 0.0, C = 0.0, thr = 0.0;
  bool sc = false;
  switch (model) {
    case EMP_KEP00:
    case EMP_AKEP00:
      per = th[0]; A = th[1]; ph = th[2]; e = th[3]; w = th[4];
      break;
    case EMP_KEP01:
      per = th[0]; A = th[1]; ph = th[2]; S = th[3]; C = th[4]; sc = true; thr = 1e-6;
      break;
    case EMP_KEP02: {
      per = exp(th[0]);
      double As = th[1], Ac = th[2];
      A = __dadd_rn(__dmul_rn(As, As), __dmul_rn(Ac, Ac));
      S = th[3]; C = th[4]; sc = true; thr = 1e-5;
      ph = acos(Ac / sqrt(A));
      if (As < 0.0) ph = kTwoPi - ph;
      break;
    }
    case EMP_KEP03:
      per = th[0]; A = th[1]; ph = th[2]; e = th[3]; w = th[4]; use_tp = true;
      break;
    case EMP_KEP04:
      per = th[0]; A = th[1]; ph = th[2]; S = th[3]; C = th[4]; sc = true; thr = 1e-5; use_tp = true;
      break;
    case EMP_KEP06:
      per = exp(th[0]); A = th[1]; ph = th[2]; e = th[3]; w = th[4];
      break;
    case EMP_KEP07:
    default:
      per = exp(th[0]); A = th[1]; ph = th[2]; S = th[3]; C = th[4]; sc = true; thr = 1e-6;
      break;
  }
  if (sc) {
    e = __dadd_rn(__dmul_rn(S, S), __dmul_rn(C, C));
    if (e < thr) {
      w = 0.0;
    } else {
      w = acos(C / sqrt(e));
      if (S < 0.0) w = kTwoPi - w;
    }
  }
}

__device__ inline void kep_constants(int model, const double* th, double t_absmax, KepConst& k) {
  double per, A, ph, e, w;
  bool use_tp;
  kep_elements(model, th, per, A, ph, e, w, use_tp);
  double sw, cw;
  sincos(w, &sw, &cw);
  double ome = 1.0 - e;
  k.freq = kTwoPi / per;
  k.tpv = use_tp ? ph : 0.0;
  k.phv = use_tp ? 0.0 : ph;
  k.e = e;
  k.ome = ome;
  k.c2 = kF2 / (1.0 + e);
  k.ome3 = 3.0 * ome;
  k.a1 = A * cw;
  k.a2 = -A * sw * sqrt(ome * (1.0 + e));
  k.a3 = A * e * cw;
  k.b1 = k.a1 * (ome * (1.0 + e));
  k.ef3 = float(e / 3.0);
  k.om23f = float(2.0 * ome / 3.0);
  k.c2f3 = float(3.0 * k.c2);
  k.efh = float(0.5 * e);
  k.ef = float(e);
  k.omef = float(ome);
  k.c2f = float(k.c2);
  k.ome3f = float(k.ome3);
  // the exact rint/FMA reduction needs |M| < 2^51 * 2pi; decide once per (walker, planet)
  const double m_bound = fabs(k.freq) * (t_absmax + fabs(k.tpv)) + fabs(k.phv);
  k.slow_mod = (m_bound < 1.0e12) ? 0 : 1;  // also catches NaN / inf parameters
  k.robust = (e >= 0.0 && e <= kGridEccMax) ? 0 : 1;
  k.tab = 0;
  k._pad = 0;
}

// ---- mean anomaly, reduced to [0, pi] exactly like NumPy's remainder ------------------
__device__ __forceinline__ double mean_anomaly(const KepConst& k, double t) {
  // `freq * X_ + phase` (kep00.model:5) / `freq * (X_ - tp)` (kep03.model:4) with the same roundings:
  // t - 0 and x + 0 are exact, so one branch-free form serves both templates
  return __dadd_rn(__dmul_rn(k.freq, __dsub_rn(t, k.tpv)), k.phv);
}

__device__ __noinline__ double mod_two_pi_slow(double M) {
  double m = fmod(M, kTwoPi);
  if (m != 0.0) {
    if (m < 0.0) m += kTwoPi;
  } else {
    m = 0.0;
  }
  return m;
}

// r = M mod 2pi in [0, 2pi] with Python semantics. Every step is exact: k = rint(M/2pi),
// M - k*c has at most 53 significant bits (multiple of ulp(c), |.| < 8), so the FMA and
// the conditional +c reproduce fmod()+fix-up bit for bit.
__device__ __forceinline__ double mod_two_pi(double M, const HotConsts& H) {
  double kd = __dsub_rn(__dadd_rn(__dmul_rn(M, H.c[4]), H.c[5]), H.c[5]);
  double r = __fma_rn(-kd, H.c[1], M);
  if (r < 0.0) r = __dadd_rn(r, H.c[1]);
  return r;  // valid for |M| < 1e12 (KepConst::slow_mod routes everything else to fmod)
}

// Folded mean anomaly Mr in [0, pi] and the sign bit of the fold (0x80000000 when the solver's
// E must be reflected to 2pi - E).  The centred remainder r = M - rint(M/2pi)*2pi is exact, and
// kepler.py's "wrap to [0, 2pi), reflect if > pi" is exactly (|r|, r < 0): for r < 0 the wrap
// gives r + 2pi (exact) and the reflection 2pi - (r + 2pi) = -r (exact).  No compare, no select.
__device__ __forceinline__ double fold_anomaly(double M, const HotConsts& H, int& sign_hi) {
  const double kd = __dsub_rn(__dadd_rn(__dmul_rn(M, H.c[4]), H.c[5]), H.c[5]);
  const double r = __fma_rn(-kd, H.c[1], M);
  sign_hi = __double2hiint(r) & 0x80000000;
  return fabs(r);
}
__device__ __forceinline__ double fold_anomaly_slow(double M, int& sign_hi) {  // generic fmod path
  const double r0 = mod_two_pi_slow(M);
  const bool high = r0 > kPi;
  sign_hi = high ? 0x80000000 : 0;
  return high ? __dsub_rn(kTwoPi, r0) : r0;
}
__device__ __forceinline__ double flip_sign(double x, int sign_hi) {
  return __hiloint2double(__double2hiint(x) ^ sign_hi, __double2loint(x));
}

// ---- x - sin x and 1 - cos x on [0, pi] (Nijenhuis-style folding, Taylor core) ----------
__device__ __forceinline__ void sin_cos_reduc(double x, double& sn, double& cs, const HotConsts& H) {
  // (selects, not c - |x - c|: the folded value must stay EXACT for small x, where E - sin E
  // needs its full relative accuracy)
  const bool bigg = x > H.c[2];
  const double u = bigg ? H.c[0] - x : x;
  const bool big = u > H.c[3];
  const double v = big ? H.c[2] - u : u;
  double w = v * v;
  // Estrin evaluation (depth 4 instead of 8): coefficients are stored highest degree first
  const double w2 = w * w, w4 = w2 * w2;
  const double s01 = fma(H.sinc[6], w, H.sinc[7]), s23 = fma(H.sinc[4], w, H.sinc[5]);
  const double s45 = fma(H.sinc[2], w, H.sinc[3]), s67 = fma(H.sinc[0], w, H.sinc[1]);
  const double ps = fma(fma(s67, w2, s45), w4, fma(s23, w2, s01));
  const double c01 = fma(H.cosc[7], w, H.cosc[8]), c23 = fma(H.cosc[5], w, H.cosc[6]);
  const double c45 = fma(H.cosc[3], w, H.cosc[4]), c67 = fma(H.cosc[1], w, H.cosc[2]);
  const double pc = fma(fma(H.cosc[0], w4, fma(c67, w2, c45)), w4, fma(c23, w2, c01));
  double ss = ps * (v * w);
  double cc = pc * w;
  double s1 = big ? (u - 1.0) + cc : ss;
  double c1 = big ? (H.c[9] + u) + ss : cc;
  sn = bigg ? fma(2.0, x, -H.c[0]) + s1 : s1;
  cs = bigg ? 2.0 - c1 : c1;
}

// ---- approximate-reciprocal helpers (MUFU seeds + Newton steps, no slow-path branches) -----
__device__ __forceinline__ double rcp_seed(double x) {  // MUFU.RCP64H, relative error 2^-23
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
template <int N>
__device__ __forceinline__ double rcp_nr(double x) {  // N Newton steps: 2^-46, 2^-92 (-> ~1 ulp)
  double y = rcp_seed(x);
#pragma unroll
  for (int i = 0; i < N; ++i) y = fma(y, fma(-x, y, 1.0), y);
  return y;
}
__device__ __forceinline__ float f32_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float f32_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float f32_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float f32_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Markley (1995) starter in FP64 (cold path: M ~ 0 or a non-finite FP32 result)
__device__ __noinline__ double markley_starter_f64(double Mr, double e, double ome) {
  const double M2 = Mr * Mr;
  const double alpha = fma(kF2 / (1.0 + e), kPi - Mr, kF1);
  const double d = fma(alpha, e, 3.0 * ome);
  const double ad = alpha * d;
  const double r = fma(3.0 * ad, d - ome, M2) * Mr;
  const double q = fma(2.0 * ad, ome, -M2);
  const double q2 = q * q;
  const double cb = cbrt(fabs(r) + sqrt(fma(q2, q, r * r)));
  const double w = cb * cb;
  const double den0 = fma(w, w + q, q2);
  return fma(2.0 * r, w, Mr * den0) / (den0 * d);
}

// Same starter in FP32 on the FMA/MUFU pipes (they issue in the slots the half-rate FP64 pipe
// leaves free).  The starter is only an initial guess with an intrinsic error of ~4e-4; its
// FP32 rounding (1e-7) changes the refined root by < 1e-18 (tests/test_kepler_gpu.py).
__device__ __forceinline__ double markley_starter(double Mr, const KepConst& k, bool& bad) {
  const float M = __double2float_rn(Mr);
  const float M2 = M * M;
  const float alpha = fmaf(k.c2f, 3.14159274f - M, 7.64804745f /* F1 */);
  const float d = fmaf(alpha, k.ef, k.ome3f);
  const float ad = alpha * d;
  const float r = fmaf(3.0f * ad, d - k.omef, M2) * M;
  const float q = fmaf(2.0f * ad, k.omef, -M2);
  const float q2 = q * q;
  const float x = fabsf(r) + f32_sqrt(fmaf(q2, q, r * r));
  const float w = f32_ex2(0.666666687f * f32_lg2(x));  // x^(2/3)
  const float den0 = fmaf(w, w + q, q2);
  const float E0f = fmaf(2.0f * r, w, M * den0) * f32_rcp(den0 * d);
  // no branch here: a slow-path call in the middle of the point's arithmetic is a scheduling
  // barrier that keeps the compiler from interleaving the two points a lane works on.  The caller
  // redoes flagged points (M ~ 0, NaN, out of range) on the cold FP64 path afterwards.
  bad = !(M > 1e-15f) || !(fabsf(E0f) < 4.0f);
  return double(E0f);
}

// Kepler's equation for a folded mean anomaly Mr in [0, pi]: returns the pre-refinement
// estimate E0 and the correction dE (E = E0 + dE), plus sin E and 1 - cos E of the refined E.
template <bool kCold>
__device__ __forceinline__ void kepler_refined(double Mr, const KepConst& k, const HotConsts& H, double& E0_out,
                                               double& dE_out, double& s1_out, double& cE1_out, bool& bad);

// One Keplerian's RV at time t (everything of kep00.model:4-8 for one point).
// kCold = false: branch-free fast path, `bad` tells the caller to redo the point with kCold = true
// (FP64 starter, fmod reduction where the walker needs it).
template <bool kCold>
__device__ __forceinline__ double kep_rv(const KepConst& k, double t, const HotConsts& H, bool& bad) {
  const double M = mean_anomaly(k, t);
  int sign_hi;
  const double Mr = (kCold && k.slow_mod) ? fold_anomaly_slow(M, sign_hi) : fold_anomaly(M, H, sign_hi);
  double E0, dE, s1, cE1;
  kepler_refined<kCold>(Mr, k, H, E0, dE, s1, cE1, bad);
  const double den = fma(k.e, cE1, k.ome);                     // 1 - e cos E1
  const double s1s = flip_sign(s1, sign_hi);                   // sin(2pi - E) = -sin E
  const double num = fma(k.a1, k.ome - cE1, k.a2 * s1s);
  return fma(num, rcp_nr<2>(den), k.a3);
}

// the eccentricity-dependent constants alone (solver entry points: no amplitudes, no frequency)
__device__ __forceinline__ void kep_constants_ecc(double ecc, KepConst& k) {
  k.freq = 1.0; k.tpv = 0.0; k.phv = 0.0;
  k.e = ecc;
  k.ome = 1.0 - ecc;
  k.c2 = kF2 / (1.0 + ecc);
  k.ome3 = 3.0 * k.ome;
  k.a1 = k.a2 = k.a3 = k.b1 = 0.0;
  k.ef = float(ecc);
  k.omef = float(k.ome);
  k.c2f = float(k.c2);
  k.ome3f = float(k.ome3);
  k.ef3 = float(ecc / 3.0);
  k.om23f = float(2.0 * k.ome / 3.0);
  k.c2f3 = float(3.0 * k.c2);
  k.efh = float(0.5 * ecc);
  k.slow_mod = 0;
  k.robust = (ecc >= 0.0 && ecc <= kGridEccMax) ? 0 : 1;
  k.tab = 0;
  k._pad = 0;
}

// kepler.solve(M, ecc) for one element (A13 of SURVEY.md §8a): E in [0, 2pi]
__device__ __forceinline__ double kepler_solve(double M, double ecc, const HotConsts& H) {
  KepConst k;
  k.e = ecc;
  k.ome = 1.0 - ecc;
  k.ef = float(ecc);
  k.omef = float(k.ome);
  k.c2f = float(kF2 / (1.0 + ecc));
  k.ome3f = 3.0f * k.omef;
  int sign_hi;
  const double Mr = (fabs(M) < 1.0e12) ? fold_anomaly(M, H, sign_hi) : fold_anomaly_slow(M, sign_hi);
  double E0, dE, s1, cE1;
  bool bad = false;
  kepler_refined<false>(Mr, k, H, E0, dE, s1, cE1, bad);
  if (bad) kepler_refined<true>(Mr, k, H, E0, dE, s1, cE1, bad);
  const double E = E0 + dE;
  return sign_hi ? H.c[1] - E : E;
}

template <bool kCold>
__device__ __forceinline__ void kepler_refined(double Mr, const KepConst& k, const HotConsts& H, double& E0_out,
                                               double& dE_out, double& s1_out, double& cE1_out, bool& bad) {
  const double E0 = kCold ? markley_starter_f64(Mr, k.e, k.ome) : markley_starter(Mr, k, bad);

  // single high-order refinement (kepler.py refine_estimate).  d3 and d4 only feed small
  // correction terms: a 2^-23 reciprocal for d3 and a 2^-46 one for d4 leave dE exact to
  // < 1e-19; dE itself and the final quotient use a fully converged reciprocal.
  double sE, cE;
  sin_cos_reduc(E0, sE, cE, H);
  const double s0 = E0 - sE;  // sin E0
  const double f0 = fma(k.e, sE, fma(E0, k.ome, -Mr));
  const double f1 = fma(k.e, cE, k.ome);
  const double f2 = k.e * s0;
  const double f3 = 1.0 - f1;
  const double d3 = -f0 * f1 * rcp_seed(fma(f1, f1, -0.5 * f0 * f2));
  const double f36 = f3 * H.c[7], f22 = 0.5 * f2;
  const double d4 = -f0 * rcp_nr<1>(fma(d3 * d3, f36, fma(d3, f22, f1)));
  const double d42 = d4 * d4;
  const double dE = -f0 * rcp_nr<2>(fma(-d42 * d4, f2 * H.c[8], fma(d42, f36, fma(d4, f22, f1))));

  // rotate (sin E0, 1 - cos E0) by dE: |dE| <= 5e-4, 4th order is exact to < 1e-20
  const double dE2 = dE * dE;
  const double sd = fma(-dE * dE2, H.c[7], dE);                 // sin dE
  const double cdm = dE2 * fma(dE2, -H.c[8], 0.5);              // 1 - cos dE
  const double c0 = 1.0 - cE;                                  // cos E0
  const double s1 = s0 + fma(c0, sd, -s0 * cdm);               // sin E1
  const double cE1 = cE + fma(s0, sd, c0 * cdm);               // 1 - cos E1
  E0_out = E0;
  dE_out = dE;
  s1_out = s1;
  cE1_out = cE1;
}

// cold path: FP64 starter (M ~ 0 / non-finite FP32 starter) and, for walkers with absurd
// frequencies (|M| >= 1e12), the generic fmod reduction.  Same arithmetic otherwise.
__device__ __noinline__ double kep_rv_cold(const KepConst& k, double t) {
  const HotConsts H = make_hot_consts();  // literals: keeps the caller free of a stack copy
  bool bad = false;
  return kep_rv<true>(k, t, H, bad);
}

// fast path + immediate fix-up (for the kernels that are not throughput critical)
__device__ __forceinline__ double kep_rv_checked(const KepConst& k, double t, const HotConsts& H) {
  bool bad = false;
  double r = kep_rv<false>(k, t, H, bad);
  if (bad || k.slow_mod) r = kep_rv_cold(k, t);
  return r;
}

// walkers/planets the grid core does not take (e > kGridEccMax, e < 0, NaN): kepler.py-style
// refinement, out of line so that its registers do not weigh on the hot loop
__device__ __noinline__ double kep_rv_robust(const KepConst& k, double t) {
  const HotConsts H = make_hot_consts();
  return kep_rv_checked(k, t, H);
}
// the same for the four points of a lane at once: the four refinements are independent chains, so the
// out-of-line path keeps the instruction-level parallelism of the inlined one (EMP_SOLVER_KEPLERPY runs here)
struct Quad { double v0, v1, v2, v3; };
__device__ __noinline__ Quad kep_rv_robust4(const KepConst& k, double t0, double t1, double t2, double t3) {
  const HotConsts H = make_hot_consts();
  bool b0 = false, b1 = false, b2 = false, b3 = false;
  Quad r;
  r.v0 = kep_rv<false>(k, t0, H, b0);
  r.v1 = kep_rv<false>(k, t1, H, b1);
  r.v2 = kep_rv<false>(k, t2, H, b2);
  r.v3 = kep_rv<false>(k, t3, H, b3);
  if (b0 || k.slow_mod) r.v0 = kep_rv_cold(k, t0);  // rare: M ~ 0, non-finite FP32 starter, |M| >= 1e12
  if (b1 || k.slow_mod) r.v1 = kep_rv_cold(k, t1);
  if (b2 || k.slow_mod) r.v2 = kep_rv_cold(k, t2);
  if (b3 || k.slow_mod) r.v3 = kep_rv_cold(k, t3);
  return r;
}

// ---- grid-anchored Kepler core (likelihood kernel v6) ----------------------------------------
// Same root as kepler.solve (Kepler's equation has ONE root), reached with
// ~half the FP64 instructions of the series-based refinement:
//   1. FP32 Markley starter E0 (error <= 4e-4), snapped to the grid: k = rint(128 E0), El = E0 - k/128
//      (exact in FP32, |El| <= 2^-8).  sin/cos of the grid point Eh = k/128 come from a table in shared
//      memory (FP64 pair + FP32 pair).
//   2. In delta-space (E = Eh + delta) Kepler's equation is a polynomial with small terms,
//        g(delta) = delta - a sin(delta) + b (1 - cos(delta)) - c,  a = e cos Eh, b = e sin Eh,
//        c = (M - Eh) + b   <- the only cancellation, done once in FP64,
//      so one FP32 Halley step from El lands within ~1e-9 of the root (no cancellation left for
//      FP32 to spoil: every term is O(delta)).
//   3. FP64: sin/cos(delta) by 3-term polynomials (|delta| < 4.4e-3), residual and slope at delta,
//      one Newton correction dd (|dd| ~ 1e-9, so a 2^-46 reciprocal is exact enough) and a
//      first-order rotation by dd give sin E, cos E of the root to ~1 ulp.
//   4. RV term [b1 cos E + a2 sin E] / (1 - e cos E), b1 = A cos w (1 - e^2) (the template's
//      A (cos(f+w) + e cos w) with the constant folded in); the reciprocal of the denominator is one
//      Newton step from the slope's reciprocal (they differ by ~1e-8 relative).
// CPU emulation vs an 80-bit solution, e in [0, 0.98]: max |dE| 1.1e-15, max |dRV/A| 6e-15
// (the oracle's own figures: 7.5e-16 and 5.2e-15) — tests/tools/kepler_grid_emulation.py.
__device__ __forceinline__ float markley_starter_f32(float M, const KepConst& k) {
  // the starter of kepler.py with the constant factors 3 and 2 folded into per-walker constants
  const float M2 = M * M;
  const float alpha3 = fmaf(k.c2f3, 3.14159274f - M, 22.9441414f /* 3 F1 */);  // 3 alpha
  const float d = fmaf(alpha3, k.ef3, k.ome3f);                                 // 3 (1-e) + alpha e
  const float ad3 = alpha3 * d;                                                 // 3 alpha d
  const float r = fmaf(ad3, d - k.omef, M2) * M;
  const float q = fmaf(ad3, k.om23f, -M2);                                      // 2 alpha d (1-e) - M^2
  const float q2 = q * q;
  const float x = fabsf(r) + f32_sqrt(fmaf(q2, q, r * r));
  const float w = f32_ex2(0.666666687f * f32_lg2(x));  // x^(2/3)
  const float den0 = fmaf(w, w + q, q2);
  return fmaf(r + r, w, M * den0) * f32_rcp(den0 * d);
}

// The core is split in two stages so that the likelihood kernel can software-pipeline them across
// planets (stage A of planet k+1 is issued next to stage C of planet k: A is FP32/MUFU/LDS heavy, C is
// pure FP64, and a warp's in-order instruction stream then feeds both pipes all the time).
struct GridStage {   // what stage A hands to stage C, per point
  double df;         // FP32-accurate delta (E = Eh + delta), widened
  double c;          // (M - Eh) + e sin Eh
  double sh, ch;     // sin Eh, cos Eh
  double eh;         // the grid point itself (only the solver entry point reads it)
  int sign_hi;       // sign bit of the centred remainder: the root is reflected to 2pi - E
};

// Stage A: mean anomaly, exact reduction, FP32 starter, grid lookup, FP32 Halley step in delta-space.
// tabf holds (sin, cos, sin/2, cos/6)(Eh) in FP32.  No validity flag: for e in [0, kGridEccMax] and a
// finite mean anomaly the FP32 starter is finite and inside [0, pi + 1e-3] (M = 0 gives E0 = 0), and
// the table index is masked.
// kTab: the starter comes from the walker's table `st` ([3][kStartStride] floats: value, slope, curvature per node).
template <bool kTab = false>
__device__ __forceinline__ void kep_grid_a(const KepConst& k, double t, const HotConsts& H,
                                           const double2* __restrict__ tab, const float4* __restrict__ tabf,
                                           GridStage& S, const float* __restrict__ st = nullptr) {
  const double M = mean_anomaly(k, t);
  // centred remainder r = M - rint(M/2pi) 2pi (exact for ANY integer near M/2pi, so the FMA in the rint is free)
  const double kd = fma(M, H.c[4], H.c[5]) - H.c[5];
  const double rr = fma(-kd, H.c[1], M);
  S.sign_hi = __double2hiint(rr) & 0x80000000;
  const double Mr = fabs(rr);
  const float Mf = __double2float_rn(Mr);
  float E0f;
  if (kTab) {
    // node kn = rint(M kStartN/pi) from the low mantissa bits of x + 1.5*2^23, dx = x - kn in [-0.5, 0.5]
    const float xs = fmaf(Mf, float(kStartN / 3.14159265358979323846), 12582912.0f);
    const int kn = __float_as_int(xs) & 127;
    const float dx = fmaf(Mf, float(kStartN / 3.14159265358979323846), 12582912.0f - xs);
    E0f = fmaf(dx, fmaf(dx, st[2 * kStartStride + kn], st[kStartStride + kn]), st[kn]);
  } else {
    E0f = markley_starter_f32(Mf, k);
  }
  // grid point: the low mantissa bits of E0*128 + 1.5*2^23 are rint(128 E0)
  const float km = fmaf(E0f, 128.0f, 12582912.0f);
  const float El = fmaf(km - 12582912.0f, -0.0078125f, E0f);
  const int ki = __float_as_int(km) & (kGridN - 1);
  const double2 sc = tab[ki];  // (sin Eh, cos Eh)
  const float4 q = tabf[ki];
  const double Eh = __hiloint2double(0x42C00000, ki) - H.g[2];  // (2^45 + k 2^-7) - 2^45, exact
  const double c = fma(k.e, sc.x, Mr - Eh);
  // FP32 Halley step in delta-space:  g = d - e d [cos Eh - d (sin Eh/2 + d cos Eh/6)] - c
  const float cf = __double2float_rn(c);
  const float P = fmaf(-El, fmaf(El, q.w, q.z), q.y);
  const float g0 = fmaf(-(k.ef * El), P, El - cf);
  const float Q = fmaf(-El, fmaf(0.5f * El, q.y, q.x), q.y);
  const float g1 = fmaf(-k.ef, Q, 1.0f);
  const float g2h = k.efh * fmaf(El, q.y, q.x);  // g'' / 2
  const float r = f32_rcp(g1);
  const float dn = g0 * r;
  const float d1 = fmaf(-dn, dn * (g2h * r), El - dn);
  S.df = double(d1);
  S.c = c;
  S.sh = sc.x;
  S.ch = sc.y;
  S.eh = Eh;
}

// Stage C, first half: the root.  sin E, cos E of the refined E = Eh + (df + dd), and 1/(1 - e cos E_f).
__device__ __forceinline__ void kep_grid_root(const KepConst& k, const GridStage& S, const HotConsts& H, double& sE,
                                              double& cE, double& dd, double& y1) {
  const double df = S.df;
  const double d2 = df * df;
  const double sl = fma(df * d2, fma(d2, H.g[0], -H.c[7]), df);          // sin(delta)
  const double cm = d2 * fma(d2, -H.c[8], 0.5);                           // 1 - cos(delta): d^6/720 < 1.1e-17
  const double w1 = fma(S.ch, sl, -S.sh * cm);                            // sin E_f - sin Eh
  const double w2 = fma(S.sh, sl, S.ch * cm);                             // cos Eh - cos E_f
  const double sEf = S.sh + w1, cEf = S.ch - w2;
  const double g = fma(-k.e, w1, df - S.c);                               // E_f - e sin E_f - M
  const double gp = fma(-k.e, cEf, 1.0);                                  // 1 - e cos E_f
  y1 = rcp_nr<1>(gp);
  dd = -g * y1;
  sE = fma(cEf, dd, sEf);
  cE = fma(-sEf, dd, cEf);
}

// Stage C: FP64 correction and the RV term, ADDED to acc.
__device__ __forceinline__ double kep_grid_c(const KepConst& k, const GridStage& S, double acc, const HotConsts& H) {
  double sE, cE, dd, y1;
  kep_grid_root(k, S, H, sE, cE, dd, y1);
  const double den = fma(-k.e, cE, 1.0);
  const double y2 = fma(y1, fma(-den, y1, 1.0), y1);                      // 1/den: Newton from 1/gp
  // A (cos(f+w) + e cos w) = [A cos w (1-e^2) cos E - A sin w sqrt(1-e^2) sin E] / (1 - e cos E)
  const double num = fma(k.b1, cE, k.a2 * flip_sign(sE, S.sign_hi));      // sin(2pi - E) = -sin E
  return fma(num, y2, acc);
}

template <bool kTab = false>
__device__ __forceinline__ double kep_rv_grid(const KepConst& k, double t, double acc, const HotConsts& H,
                                              const double2* __restrict__ tab, const float4* __restrict__ tabf,
                                              const float* __restrict__ st = nullptr) {
  GridStage S;
  kep_grid_a<kTab>(k, t, H, tab, tabf, S, st);
  return kep_grid_c(k, S, acc, H);
}

// Build one planet's starter table (all 32 lanes of the walker's warp).  E at the 2*kStartN + 1 half-spaced points
// M_j = j pi/(2 kStartN) by the FP32 Markley starter + two FP32 Newton steps (MUFU sin/cos: ~2e-6), plus the two
// ghost points outside [0, pi] by symmetry; node n gets the quadratic through (n - 1/2, n, n + 1/2).
// scratch: 2*kStartN + 3 floats of shared memory owned by the warp.
__device__ __forceinline__ void build_start_table(const KepConst& k, float* st, float* scratch, int lane) {
  const float hh = float(3.14159265358979323846 / (2 * kStartN));
  for (int j = lane; j <= 2 * kStartN; j += 32) {
    const float M = hh * float(j);
    float E = (j == 0) ? 0.0f : markley_starter_f32(M, k);
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const float sE = __sinf(E), cE = __cosf(E);
      E -= (fmaf(-k.ef, sE, E) - M) * f32_rcp(fmaf(-k.ef, cE, 1.0f));
    }
    if (j == 0) E = 0.0f;
    if (j == 2 * kStartN) E = 3.14159274f;
    scratch[j + 1] = E;
  }
  __syncwarp();
  if (lane == 0) {
    scratch[0] = -scratch[2];                                            // E(-M) = -E(M)
    scratch[2 * kStartN + 2] = 6.28318548f - scratch[2 * kStartN];       // E(pi + d) = 2 pi - E(pi - d)
  }
  __syncwarp();
  for (int n = lane; n <= kStartN; n += 32) {
    const float em = scratch[2 * n], e0 = scratch[2 * n + 1], ep = scratch[2 * n + 2];
    st[n] = e0;
    st[kStartStride + n] = ep - em;                       // slope per cell
    st[2 * kStartStride + n] = 2.0f * ((ep - e0) - (e0 - em));  // curvature: 2 (E+ + E- - 2 E0)
  }
  __syncwarp();
}

// exp(x) for x <= 0 (decay factors of the MA block): n = rint(x log2 e), r = x - n ln2 (two-part),
// degree-11 Taylor polynomial by Estrin (|r| <= 0.347: truncation 6e-15 relative), scaled by 2^n in the
// exponent field.  Coefficients come from the kernel-parameter bank.  x < -708 saturates at ~1e-308.
__device__ __forceinline__ double exp_neg(double x, const HotConsts& H) {
  const double kd = fma(x, H.ex[0], H.c[5]);
  const double n = kd - H.c[5];
  const double r = fma(n, H.ex[2], fma(n, H.ex[1], x));
  const double r2 = r * r, r4 = r2 * r2;
  const double p01 = 1.0 + r;
  const double p23 = fma(r, H.ex[4], H.ex[3]);
  const double p45 = fma(r, H.ex[6], H.ex[5]);
  const double p67 = fma(r, H.ex[8], H.ex[7]);
  const double p89 = fma(r, H.ex[10], H.ex[9]);
  const double pab = fma(r, H.ex[12], H.ex[11]);
  const double lo = fma(r2, p23, p01), mid = fma(r2, p67, p45), hi = fma(r2, pab, p89);
  const double p = fma(r4, fma(r4, hi, mid), lo);
  const int ni = max(__double2loint(kd), -1021);
  return __hiloint2double(__double2hiint(p) + (ni << 20), __double2loint(p));
}

// ---- A cos(freq t + phase): support/models/sinusoid00.model, magneticcycle00.model -------------
struct PeriodicTerm {
  double freq, phase, amp;
  double slow;  // != 0: |freq t + phase| may exceed 1e12 -> library cos()
};
__device__ inline PeriodicTerm make_periodic(double freq, double phase, double amp, double t_absmax) {
  PeriodicTerm p;
  p.freq = freq; p.phase = phase; p.amp = amp;
  p.slow = (fabs(freq) * t_absmax + fabs(phase) < 1.0e12) ? 0.0 : 1.0;
  return p;
}
__device__ __forceinline__ double periodic_value(const PeriodicTerm& p, double t, const HotConsts& H) {
  const double M = __dadd_rn(__dmul_rn(p.freq, t), p.phase);  // freq * X_ + phase
  if (p.slow != 0.0) return p.amp * cos(M);
  int sign_hi;
  const double r = fold_anomaly(M, H, sign_hi);  // cos(M) = cos(|r|), r in [0, pi] exactly reduced
  double sn, cs;
  sin_cos_reduc(r, sn, cs, H);                   // cs = 1 - cos r
  return p.amp * (1.0 - cs);
}

// ---- priors: support/priors/{Uniform,Normal,Jeffreys,Isotropic,Fixed}.prior ------------
__device__ inline double prior_value(const EmpPriorOp& o, double x) {
  if (o.prior == EMP_PRIOR_FIXED) return 0.0;
  if (!(o.lo <= x && x <= o.hi)) return -INFINITY;
  switch (o.prior) {
    case EMP_PRIOR_UNIFORM:
    case EMP_PRIOR_JEFFREYS:
      return o.a0;
    case EMP_PRIOR_NORMAL: {
      // -0.5*((x - mu)/s)**2 - np.log(s*np.sqrt(2*np.pi)) - logZ   (Normal.prior:8), same roundings
      double z = __ddiv_rn(__dsub_rn(x, o.a0), o.a1);
      double v = __dmul_rn(-0.5, __dmul_rn(z, z));
      return __dsub_rn(__dsub_rn(v, o.a3), o.a2);
    }
    case EMP_PRIOR_ISOTROPIC:
      return __dsub_rn(log(__dmul_rn(0.5, sin(x))), o.a0);  // Isotropic.prior:4
    default:
      return NAN;
  }
}

// my_prior (emp.py:190-254) as the straight-line program in the descriptor.
__device__ inline double prior_program(const EmpPriorOp* ops, int n_ops, const double* th) {
  double lp = 0.0;
  for (int i = 0; i < n_ops; ++i) {
    const EmpPriorOp& o = ops[i];
    if (o.op == EMP_POP_PARAM) {
      lp = __dadd_rn(lp, prior_value(o, th[o.i0]));
    } else if (o.op == EMP_POP_CHECK) {
      if (lp == -INFINITY) return lp;
    } else {
      double a = th[o.i0], b = th[o.i1];
      double x = __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b));
      lp = __dadd_rn(lp, prior_value(o, x));
    }
  }
  return lp;
}

// The same program evaluated by a whole warp: every lane evaluates the priors of the ops it owns (they are
// independent), lane 0 adds them in program order.  The reference's early `return lp` at the end of a block
// whose running sum is -inf is the `break` below (values behind it are never added), so the result is
// bit-identical to prior_program.  vals: EMP_MAX_PRIOR_OPS doubles of shared scratch owned by the warp.
__device__ __forceinline__ double prior_program_warp(const EmpModelDesc* __restrict__ d, const double* th,
                                                     double* vals, int lane) {
  const int n_ops = d->n_prior_ops;
  for (int k = lane; k < n_ops; k += 32) {
    const EmpPriorOp& op = d->prior_ops[k];
    double v = 0.0;
    if (op.op == EMP_POP_PARAM) {
      v = prior_value(op, th[op.i0]);
    } else if (op.op == EMP_POP_SUMSQ) {
      const double a = th[op.i0], b = th[op.i1];
      v = prior_value(op, __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)));
    }
    vals[k] = v;
  }
  __syncwarp();
  double lp = 0.0;
  if (lane == 0) {
    for (int k = 0; k < n_ops; ++k) {
      if (d->prior_ops[k].op == EMP_POP_CHECK) {
        if (lp == -INFINITY) break;
      } else {
        lp = __dadd_rn(lp, vals[k]);
      }
    }
  }
  return __shfl_sync(0xffffffffu, lp, 0);
}

// ---- small helpers -----------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// mbarrier + 1-D bulk TMA (cp.async.bulk; SASS: UBLKCP) -----------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u)  // suspend-time hint: the waiting thread sleeps in hardware
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

}  // namespace emp
