// emp_am.cuh — Hipparcos-Gaia astrometric block on the device.
//
// Restates `loglike_AM` and its helpers as the reference generates them
// (ReddModel._write_model_AM, emp_model.py:1232-1672; constants support/astrometry/constants.scr):
//   reflex orbit of every Keplerian at the 108 Hipparcos + 40 Gaia (GOST) epochs through the
//   Thiele-Innes constants, barycentre = GDR3 catalogue - offsets propagated linearly in 3-D
//   (obs_lin_prop_PA) to the Hipparcos epoch, along-scan projection against the Hipparcos IAD
//   residuals (iid Gaussian with jitter J_H), Gaia: 5-parameter linear refit of the synthetic
//   along-scan signal (solution vectors GSV) against the GDR2 / GDR3 catalogue differences
//   (multivariate normal, covariance inflated by J_G^2).
// One warp per evaluation (only rows inside the prior support: the compact list the RV launch
// built); lanes stride the epochs, lanes 0/1 propagate the barycentre, lanes 0..4 do the refit.
//
// Precision: the reference mixes FP64 with x87 80-bit `np.longdouble` where it forms
// `(dec - ref_dec) * 3.6e6` (constants.scr:5-6,13-14; emp_model.py:1423-1431) and propagates ABSOLUTE angles in
// FP64 before differencing them against the catalogues: its value is only defined to ~4e-9 relative (a 40-digit
// evaluation of the same formulas, tests/tools/am_truth_mpmath.py, differs from the reference's float by up to
// 3.9e-9).  A GPU has no 80-bit type and does not need one: every catalogue difference is formed here from
// offsets that never cancel (lin_prop_offsets), which agrees with the 40-digit value to ~1e-14 relative.  The
// device is therefore tested against the EXACT value (<= 1e-10, north_star's bar) and against the reference at
// the reference's own accuracy (tests/test_am.py).
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "../../include/emperor_b200.h"
#include "emp_device.cuh"
#include "emp_pt.cuh"

static int fail(int code, const std::string& msg);

namespace emp {

constexpr int kAmWarps = 4;
constexpr int kAmMaxGost = 256;
constexpr double kLog2Pi = 1.8378770664093454835606594728112;
constexpr double kPcPerKpc = 1e3, kDayPerYear = 365.25, kPc2Au = 206265.0, kAuyr2Kms = 4.74047;
constexpr double kDeg2Rad = 0.017453292519943295769, kRad2Deg = 57.295779513082320877;

struct AmDevice {
  int enabled = 0;
  int n_hipp = 0, n_gost = 0, n_mask2 = 0, n_mask3 = 0;
  // device arrays
  double* t_rel = nullptr;  // [n_hipp + n_gost] time_iad_all - common_t
  double *cpsi_h = nullptr, *spsi_h = nullptr, *epoch_h = nullptr, *parf_h = nullptr, *res_h = nullptr,
         *sres2_h = nullptr;
  double *tg_ref = nullptr, *cpsi_g = nullptr, *spsi_g = nullptr, *parf_g = nullptr;
  int32_t *idx2 = nullptr, *idx3 = nullptr;
  double *gsv2 = nullptr, *gsv3 = nullptr;
  // small constants by value
  double inv_cov[3][25];
  double log_det_cov[3];
  double astro_gost[2][5];
  double catalogs[3][7];
  double times_refed[3];
};

struct AmParams {
  const EmpModelDesc* desc;
  const double* theta;
  const int32_t* eval_index;
  const int32_t* n_active;
  double* logl;
  int n_hipp, n_gost, n_mask2, n_mask3;
  const double* t_rel;
  const double *cpsi_h, *spsi_h, *epoch_h, *parf_h, *res_h, *sres2_h;
  const double *tg_ref, *cpsi_g, *spsi_g, *parf_g;
  const int32_t *idx2, *idx3;
  const double *gsv2, *gsv3;
  double inv_cov[3][25];
  double log_det_cov[3];
  double astro_gost[2][5];
  double cat_h[7];    // Hipparcos catalogue row (AM_catalogs_[0])
  double cat_ref[7];  // GDR3 row (AM_catalogs_[-1]); [1:] = AM_catalogs_obs_ref
  double t_ref_h;     // AM_catalogs_times_refed[0]
  HotConsts H;
  PtAccept pt;        // PT step: Metropolis accept once the joint RV + AM likelihood is known (emp_pt.cuh)
};

struct AmPlanet {
  KepConst k;       // only the solver fields are used
  double pha, sq;   // phase, sqrt(1 - e^2)
  double beta, A, B, F, G, C, Hc, plxfac;
};

struct AmWarp {
  double th[EMP_MAX_DIM];
  AmPlanet pl[EMP_MAX_KEP];
  double bary_h[6];
  double bary_g[6];
  double deltas[5];
  double abs_g[kAmMaxGost];
  double prm[2][5];
};

// obs_lin_prop_PA for ONE catalogue epoch (emp_model.py:1498-1573), formulated on OFFSETS.
// The reference propagates absolute angles (deg -> rad -> xyz -> atan2 -> deg) and then forms
// (RA_bary - RA_catalogue) * 3.6e6: a difference of two ~70 degree numbers that agree to 1e-5, which leaves the
// value of loglike_AM defined only to ~4e-9 relative (tests/test_am.py, tests/tools/am_truth_mpmath.py) — even in
// its x87 80-bit arithmetic.  In the star's local frame (r, north, east) the propagated position is exactly
//   p1 = (D + vr tf) r + (vde tf) n + (vra tf) e,         D = 1e3/plx,
// so the CHANGE of the angles has a closed form without any cancellation:
//   dRA = atan2(E, R cd - N sd),   sin(dDE) = (N - sd A kappa)/|p1|,  A = R cd - N sd, kappa = sqrt(1 + (E/A)^2) - 1.
// Every difference against a catalogue position is then a sum of small, accurately known terms.  In plain FP64
// this agrees with a 40-digit evaluation of the reference's formulas to 2e-15 relative (the reference: 4e-9).
//   obs: ra, de [rad], plx, pmra, pmde, rv of the un-propagated barycentre; time_refed [d]
//   out: dRA, dDE [deg] (position change), plx1, pmra1, pmde1, de1 [deg]
__device__ inline void lin_prop_offsets(double ra, double de, double plx, double pmra, double pmde, double rv,
                                        double time_refed, double* out) {
  double sinde, cosde, sinra, cosra;
  sincos(de, &sinde, &cosde);
  sincos(ra, &sinra, &cosra);
  const double d = 1.0 / plx;
  const double D = d * kPcPerKpc;
  const double vra = pmra * d, vde = pmde * d, vr = rv / kAuyr2Kms;
  const double tf = time_refed / (kDayPerYear * kPc2Au);
  const double R = D + vr * tf, N = vde * tf, E = vra * tf;
  const double A = R * cosde - N * sinde;
  const double dra = atan2(E, A);
  const double eps = (E / A) * (E / A);
  const double kappa = eps / (1.0 + sqrt(1.0 + eps));
  const double L = sqrt(R * R + N * N + E * E);
  const double dde = asin((N - sinde * A * kappa) / L);
  // proper motions / parallax at the new epoch: the velocity in the local frame of the NEW direction
  const double vx = vr * cosde * cosra - vde * sinde * cosra - vra * sinra;
  const double vy = vr * cosde * sinra - vde * sinde * sinra + vra * cosra;
  const double vz = vr * sinde + vde * cosde;
  double sinra1, cosra1, sinde1, cosde1;
  sincos(ra + dra, &sinra1, &cosra1);
  sincos(de + dde, &sinde1, &cosde1);
  const double r0 = cosra1 * vx + sinra1 * vy;
  const double r1 = -sinra1 * vx + cosra1 * vy;
  const double v2 = -sinde1 * r0 + cosde1 * vz;
  const double d1 = L * 1e-3;
  out[0] = dra * kRad2Deg;
  out[1] = dde * kRad2Deg;
  out[2] = 1.0 / d1;
  out[3] = r1 / d1;
  out[4] = v2 / d1;
  out[5] = (de + dde) * kRad2Deg;
}

// reflex-orbit offsets (mas) of one planet at relative time t: rows 0..2 of calc_astro_new
__device__ __forceinline__ void am_orbit(const AmPlanet& p, double t, const HotConsts& H, double& ras, double& dec,
                                         double& plx) {
  const double M = __dadd_rn(__dmul_rn(p.k.freq, t), p.pha);
  int sign_hi;
  const double Mr = (fabs(M) < 1.0e12) ? fold_anomaly(M, H, sign_hi) : fold_anomaly_slow(M, sign_hi);
  double E0, dE, s1, cE1;
  bool bad = false;
  kepler_refined<false>(Mr, p.k, H, E0, dE, s1, cE1, bad);
  if (bad) kepler_refined<true>(Mr, p.k, H, E0, dE, s1, cE1, bad);
  const double X = p.k.ome - cE1;              // cos E - e
  const double Y = p.sq * flip_sign(s1, sign_hi);  // sqrt(1-e^2) sin E
  ras += p.beta * (p.B * X + p.G * Y);
  dec += p.beta * (p.A * X + p.F * Y);
  plx += p.plxfac * (p.C * X + p.Hc * Y);
}

__global__ void __launch_bounds__(kAmWarps * 32) am_logl_kernel(const AmParams P) {
  __shared__ AmWarp ws[kAmWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * kAmWarps + warp;
  if (e >= *P.n_active) return;
  const int64_t slot = P.eval_index[e];
  const EmpModelDesc* __restrict__ d = P.desc;
  AmWarp& w = ws[warp];
  // the reference indexes the UN-expanded theta with full-theta slices (a00.like:7); the host
  // rejects AM models with fixed parameters, so free == full here
  for (int i = lane; i < d->ndim_free; i += 32) w.th[i] = P.theta[slot * d->ndim_free + i];
  __syncwarp();
  const double* off = w.th + d->am_offset_off;  // dra, dde, dplx, dpmra, dpmde offsets
  const double J_H = w.th[d->am_jitter_off], J_G = w.th[d->am_jitter_off + 1];
  const double plx0 = P.cat_ref[3] - off[2];  // AM_PLX_ref - theta_am_off[2]
  const int K = d->n_kep;

  if (lane < K) {
    const double* th = w.th + d->kep_off[lane];
    const double per = th[0], Kamp = th[1], pha = th[2], ecc = th[3], omega = th[4], inc = th[5], Om = th[6];
    AmPlanet p;
    p.k.freq = kTwoPi / per;
    p.k.e = ecc;
    p.k.ome = 1.0 - ecc;
    p.k.ef = float(ecc);
    p.k.omef = float(p.k.ome);
    p.k.c2f = float(kF2 / (1.0 + ecc));
    p.k.ome3f = 3.0f * p.k.omef;
    p.pha = pha;
    double sinI, cosI, sinOM, cosOM, sinom, cosom;
    sincos(inc, &sinI, &cosI);
    sincos(Om, &sinOM, &cosOM);
    sincos(omega, &sinom, &cosom);
    p.sq = sqrt(1.0 - ecc * ecc);
    p.A = cosom * cosOM - sinom * sinOM * cosI;
    p.B = cosom * sinOM + sinom * cosOM * cosI;
    p.F = -sinom * cosOM - cosom * sinOM * cosI;
    p.G = -sinom * sinOM + cosom * cosOM * cosI;
    p.C = sinom * sinI;
    p.Hc = cosom * sinI;
    const double beta0 = per / kDayPerYear * (Kamp / kPcPerKpc / kAuyr2Kms) * p.sq / kTwoPi / sinI;
    p.beta = -beta0 * plx0;
    p.plxfac = -p.beta * plx0 / 206265e3;
    w.pl[lane] = p;
  }
  // barycentre (model_barycenter, emp_model.py:1483-1495): lane 0 -> Hipparcos epoch, lane 1 -> GDR3 epoch.
  // Positions are carried as offsets from the GDR3 catalogue position (lin_prop_offsets).
  const double dec_ref = P.cat_ref[2];
  const double dra_obs = -(off[0] / 3.6e6) / cos(dec_ref * kDeg2Rad);  // RA_obs - RA_ref [deg]
  const double dde_obs = -off[1] / 3.6e6;                              // DE_obs - DE_ref [deg]
  if (lane < 2) {
    const double ra = (P.cat_ref[1] + dra_obs) * kDeg2Rad, de = (P.cat_ref[2] + dde_obs) * kDeg2Rad;
    double o[6];
    lin_prop_offsets(ra, de, P.cat_ref[3] - off[2], P.cat_ref[4] - off[3], P.cat_ref[5] - off[4], P.cat_ref[6],
                     lane == 0 ? P.t_ref_h : 0.0, o);
    if (lane == 0) {
      // get_deltas_HIPP (emp_model.py:1403-1415): (RA_ref - RA_hip) is an exact difference of nearby doubles
      const double mean_dec = 0.5 * (P.cat_h[2] + o[5]);
      w.deltas[0] = ((P.cat_ref[1] - P.cat_h[1]) + dra_obs + o[0]) * cos(mean_dec * kDeg2Rad) * 3.6e6;
      w.deltas[1] = ((P.cat_ref[2] - P.cat_h[2]) + dde_obs + o[1]) * 3.6e6;
      w.deltas[2] = o[2] - P.cat_h[3];
      w.deltas[3] = o[3] - P.cat_h[4];
      w.deltas[4] = o[4] - P.cat_h[5];
    } else {
      w.bary_g[0] = de;    // declination of the barycentre at the reference epoch [rad]
      w.bary_g[2] = o[2];  // parallax, proper motions at the reference epoch
      w.bary_g[3] = o[3];
      w.bary_g[4] = o[4];
    }
  }
  __syncwarp();

  // ---- Hipparcos IAD (compute_abs_signal_hipp + gaussian_loglike_iid) ------------------------
  double acc = 0.0;
  const double JH2 = J_H * J_H;
  for (int i = lane; i < P.n_hipp; i += 32) {
    double ras = 0.0, dec = 0.0, plx = 0.0;
    const double t = P.t_rel[i];
    for (int k = 0; k < K; ++k) am_orbit(w.pl[k], t, P.H, ras, dec, plx);
    const double dra0 = ras + w.deltas[0], dde0 = dec + w.deltas[1];
    const double ep = P.epoch_h[i];
    const double ab = P.cpsi_h[i] * (dra0 + w.deltas[3] * ep) + P.spsi_h[i] * (dde0 + w.deltas[4] * ep) +
                      P.parf_h[i] * w.deltas[2];
    const double res = P.res_h[i] - ab;
    const double var = P.sres2_h[i] + JH2;
    acc += res * res / var + log(var);
  }
  acc = warp_sum(acc);
  double ll = -0.5 * (acc + double(P.n_hipp) * kLog2Pi);

  // ---- Gaia GOST epochs (_prepare_gost_inputs, obs_lin_prop_simple, get_deltas_GOST) --------
  {
    const double* bg = w.bary_g;  // barycentre at the reference epoch: [0] DEC [rad], [2] PLX, [3] PMRA, [4] PMDEC
    const double de = bg[0];
    const double ref_dec = P.cat_ref[2];
    for (int j = lane; j < P.n_gost; j += 32) {
      double ras = 0.0, dec = 0.0, plx = 0.0;
      const double t = P.t_rel[P.n_hipp + j];
      for (int k = 0; k < K; ++k) am_orbit(w.pl[k], t, P.H, ras, dec, plx);
      const double tg = P.tg_ref[j];
      // obs_lin_prop_simple (emp_model.py:1576-1594) as offsets from the catalogue position [deg]
      const double decs = de + bg[4] * tg / kDayPerYear / 206265e3;
      const double dra_g = dra_obs + (bg[3] * tg / kDayPerYear / cos(decs) / 206265e3) * kRad2Deg;
      const double dde_g = dde_obs + (bg[4] * tg / kDayPerYear / 206265e3) * kRad2Deg;
      // get_deltas_GOST (emp_model.py:1418-1431)
      const double dec_deg = ref_dec + dde_g + dec / 3.6e6;
      const double cos_dec = cos(dec_deg * kDeg2Rad);
      const double dra = dra_g * cos_dec * 3.6e6 + ras;
      const double ddec = dde_g * 3.6e6 + dec;
      const double dplx = bg[2] + plx;
      w.abs_g[j] = P.spsi_g[j] * dra + P.cpsi_g[j] * ddec + P.parf_g[j] * dplx;
    }
  }
  __syncwarp();
  // params = GSV[cat] @ abs_gost[mask]: lanes 0..4 -> GDR2 rows, lanes 8..12 -> GDR3 rows
  if (lane < 5) {
    double s = 0.0;
    for (int m = 0; m < P.n_mask2; ++m) s += P.gsv2[lane * P.n_mask2 + m] * w.abs_g[P.idx2[m]];
    w.prm[0][lane] = P.astro_gost[0][lane] - s;
  } else if (lane >= 8 && lane < 13) {
    const int r = lane - 8;
    double s = 0.0;
    for (int m = 0; m < P.n_mask3; ++m) s += P.gsv3[r * P.n_mask3 + m] * w.abs_g[P.idx3[m]];
    w.prm[1][r] = P.astro_gost[1][r] - s;
  }
  __syncwarp();
  if (lane == 0) {
    const double JG2 = J_G * J_G;
    const double ljg = log(JG2);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const double* ic = P.inv_cov[c + 1];
      double quad = 0.0;
      for (int a = 0; a < 5; ++a) {
        double row = 0.0;
        for (int b = 0; b < 5; ++b) row += w.prm[c][b] * ic[b * 5 + a];  // (res @ inv_cov)[a]
        quad += row * w.prm[c][a];
      }
      quad /= JG2;
      ll += -0.5 * (quad + 5.0 * ljg + P.log_det_cov[c + 1] + 5.0 * kLog2Pi);
    }
    ll = P.logl[slot] + ll;  // a00.like:8  ll1 + ll2
    if (!P.pt.enabled) P.logl[slot] = ll;
  }
  if (P.pt.enabled) pt_accept_row(P.pt, d->ndim_free, slot, __shfl_sync(0xffffffffu, ll, 0), lane);
}

// ---- host side ---------------------------------------------------------------------------------
template <typename T>
static int am_up(T** dst, const T* src, size_t n) {
  if (cudaMalloc(dst, n * sizeof(T)) != cudaSuccess) return fail(EMP_ENOMEM, "cudaMalloc (astrometry)");
  if (cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess)
    return fail(EMP_ECUDA, "cudaMemcpy (astrometry)");
  return EMP_OK;
}

inline int am_upload(const EmpAmData* a, AmDevice* d) {
  if (!a->time_hipp || !a->time_gost || !a->catalogs || !a->gsv2 || !a->gsv3 || !a->inv_cov)
    return fail(EMP_EINVAL, "astrometry data has NULL arrays");
  if (a->n_hipp < 1 || a->n_gost < 1 || a->n_gost > kAmMaxGost)
    return fail(EMP_EINVAL, "astrometry: need 1 <= n_gost <= 256 and n_hipp >= 1");
  for (int m = 0; m < a->n_mask2; ++m)
    if (a->idx_mask2[m] < 0 || a->idx_mask2[m] >= a->n_gost) return fail(EMP_EINVAL, "idx_mask2 out of range");
  for (int m = 0; m < a->n_mask3; ++m)
    if (a->idx_mask3[m] < 0 || a->idx_mask3[m] >= a->n_gost) return fail(EMP_EINVAL, "idx_mask3 out of range");
  d->n_hipp = a->n_hipp; d->n_gost = a->n_gost; d->n_mask2 = a->n_mask2; d->n_mask3 = a->n_mask3;
  const int n_iad = a->n_hipp + a->n_gost;
  std::vector<double> trel(n_iad), sres2(a->n_hipp), tg(a->n_gost);
  const double ref_epoch = a->catalogs[2 * 7 + 0];
  for (int i = 0; i < a->n_hipp; ++i) {
    trel[i] = a->time_hipp[i] - a->common_t;           // (time_iad_all - common_t), emp_model.py:1322
    sres2[i] = a->sres_hipp[i] * a->sres_hipp[i];      // SRES_HIPP_**2
  }
  for (int j = 0; j < a->n_gost; ++j) {
    trel[a->n_hipp + j] = a->time_gost[j] - a->common_t;
    tg[j] = a->time_gost[j] - ref_epoch;               // time_iad_gost_refed
  }
  int rc;
  if ((rc = am_up(&d->t_rel, trel.data(), n_iad))) return rc;
  if ((rc = am_up(&d->cpsi_h, a->cpsi_hipp, a->n_hipp))) return rc;
  if ((rc = am_up(&d->spsi_h, a->spsi_hipp, a->n_hipp))) return rc;
  if ((rc = am_up(&d->epoch_h, a->epoch_hipp, a->n_hipp))) return rc;
  if ((rc = am_up(&d->parf_h, a->parf_hipp, a->n_hipp))) return rc;
  if ((rc = am_up(&d->res_h, a->res_hipp, a->n_hipp))) return rc;
  if ((rc = am_up(&d->sres2_h, sres2.data(), a->n_hipp))) return rc;
  if ((rc = am_up(&d->tg_ref, tg.data(), a->n_gost))) return rc;
  if ((rc = am_up(&d->cpsi_g, a->cpsi_gost, a->n_gost))) return rc;
  if ((rc = am_up(&d->spsi_g, a->spsi_gost, a->n_gost))) return rc;
  if ((rc = am_up(&d->parf_g, a->parf_gost, a->n_gost))) return rc;
  if ((rc = am_up(&d->idx2, a->idx_mask2, a->n_mask2 > 0 ? a->n_mask2 : 1))) return rc;
  if ((rc = am_up(&d->idx3, a->idx_mask3, a->n_mask3 > 0 ? a->n_mask3 : 1))) return rc;
  if ((rc = am_up(&d->gsv2, a->gsv2, size_t(5) * (a->n_mask2 > 0 ? a->n_mask2 : 1)))) return rc;
  if ((rc = am_up(&d->gsv3, a->gsv3, size_t(5) * (a->n_mask3 > 0 ? a->n_mask3 : 1)))) return rc;
  for (int c = 0; c < 3; ++c) {
    for (int k = 0; k < 25; ++k) d->inv_cov[c][k] = a->inv_cov[c * 25 + k];
    d->log_det_cov[c] = a->log_det_cov[c];
    for (int k = 0; k < 7; ++k) d->catalogs[c][k] = a->catalogs[c * 7 + k];
    d->times_refed[c] = a->catalogs[c * 7] - ref_epoch;
  }
  for (int c = 0; c < 2; ++c)
    for (int k = 0; k < 5; ++k) d->astro_gost[c][k] = a->astro_gost[c * 5 + k];
  d->enabled = 1;
  return EMP_OK;
}

inline void am_free(AmDevice* d) {
  cudaFree(d->t_rel); cudaFree(d->cpsi_h); cudaFree(d->spsi_h); cudaFree(d->epoch_h); cudaFree(d->parf_h);
  cudaFree(d->res_h); cudaFree(d->sres2_h); cudaFree(d->tg_ref); cudaFree(d->cpsi_g); cudaFree(d->spsi_g);
  cudaFree(d->parf_g); cudaFree(d->idx2); cudaFree(d->idx3); cudaFree(d->gsv2); cudaFree(d->gsv3);
  *d = AmDevice();
}

// adds loglike_AM to logl for every row of the compact list (built by prior_compact_kernel)
inline int am_launch(AmDevice* d, const EmpModelDesc* d_desc, const double* theta_dev, int64_t n_eval,
                     const int32_t* eval_index, const int32_t* n_active, double* logl_dev, cudaStream_t stream,
                     int64_t* launches, const PtAccept* accept) {
  if (!d->enabled) return fail(EMP_EINVAL, "astrometry data not uploaded");
  AmParams P;
  P.desc = d_desc; P.theta = theta_dev; P.eval_index = eval_index; P.n_active = n_active; P.logl = logl_dev;
  P.n_hipp = d->n_hipp; P.n_gost = d->n_gost; P.n_mask2 = d->n_mask2; P.n_mask3 = d->n_mask3;
  P.t_rel = d->t_rel;
  P.cpsi_h = d->cpsi_h; P.spsi_h = d->spsi_h; P.epoch_h = d->epoch_h; P.parf_h = d->parf_h;
  P.res_h = d->res_h; P.sres2_h = d->sres2_h;
  P.tg_ref = d->tg_ref; P.cpsi_g = d->cpsi_g; P.spsi_g = d->spsi_g; P.parf_g = d->parf_g;
  P.idx2 = d->idx2; P.idx3 = d->idx3; P.gsv2 = d->gsv2; P.gsv3 = d->gsv3;
  memcpy(P.inv_cov, d->inv_cov, sizeof(P.inv_cov));
  memcpy(P.log_det_cov, d->log_det_cov, sizeof(P.log_det_cov));
  memcpy(P.astro_gost, d->astro_gost, sizeof(P.astro_gost));
  memcpy(P.cat_h, d->catalogs[0], sizeof(P.cat_h));
  memcpy(P.cat_ref, d->catalogs[2], sizeof(P.cat_ref));
  P.t_ref_h = d->times_refed[0];
  P.H = make_hot_consts();
  P.pt = PtAccept();
  P.pt.enabled = 0;
  if (accept) P.pt = *accept;
  const unsigned grid = unsigned((n_eval + kAmWarps - 1) / kAmWarps);
  am_logl_kernel<<<grid, kAmWarps * 32, 0, stream>>>(P);
  *launches += 1;
  if (cudaGetLastError() != cudaSuccess) return fail(EMP_ECUDA, "am_logl_kernel launch failed");
  return EMP_OK;
}

}  // namespace emp
