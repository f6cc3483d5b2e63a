// emp_logl.cuh — batched my_likelihood + my_prior kernels.
//
// Two launches per batch:
//   1. prior_compact_kernel — one warp per evaluation: re-insert the fixed parameters, run the
//      prior program, write logp; rows with prior == -inf get logl = -inf and are dropped
//      (emcee never evaluates the likelihood there), the others are appended to a compact
//      index list.  Without this pass ~40 % of the proposals of a young chain sit outside the
//      prior box and their warps would idle next to working ones.
//   2. logl_rv_kernel — one WARP per surviving evaluation; lanes own consecutive pairs of
//      datapoints, so the Keplerian constants are warp-uniform and every shared-memory read is
//      a conflict-free 128-bit access.  A CTA is kWalkerWarps walkers that consume the same
//      stream of data tiles (t | y | yerr^2 | instrument id), staged from L2 into a
//      kStages-deep shared-memory ring by 1-D bulk TMA (cp.async.bulk + mbarrier complete_tx):
//      each tile is fetched once per kWalkerWarps evaluations.  There is no producer warp:
//      lane 0 of warp 0 refills the stage that every warp released one tile ago, so it only
//      ever blocks on a warp that is two tiles behind.
//      chi^2 and sum(log err2) are accumulated per lane and reduced with warp shuffles in a
//      fixed order: results are bit-reproducible run to run and independent of the slot.
#pragma once
#include "emp_device.cuh"
#include "emp_pt.cuh"

namespace emp {

constexpr int kTilePoints = 512;                       // datapoints per TMA tile
constexpr int kTileBytes = kTilePoints * (3 * 8 + 4);  // t, y, e2 (FP64) + ins (int32)
#ifndef EMP_STAGES
#define EMP_STAGES 3
#endif
constexpr int kStages = EMP_STAGES;
constexpr int kWalkerWarps = 8;                        // walkers per CTA
constexpr int kLoglThreads = kWalkerWarps * 32;

// per-walker constants in shared memory (one slot per walker warp)
struct WalkerConst {
  double th[EMP_MAX_DIM];
  KepConst kep[EMP_MAX_KEP];
  double gamma[EMP_MAX_INS];
  double jit2[EMP_MAX_INS];
  double acc[EMP_MAX_ACC];
  double ma[2 * EMP_MAX_MA];
  double ma_itau[EMP_MAX_MA];              // 1 / tau_c
  double ma_rh[EMP_MAX_MA], ma_thist[EMP_MAX_MA];  // MA(order >= 2) history, newest first (warp-uniform)
  double sai[EMP_MAX_INS * EMP_MAX_SAI];    // activity coefficient of (instrument, column), 0 where absent
  PeriodicTerm per[2 * EMP_MAX_PERIODIC];  // A cos(freq t + phase) terms of Sinusoid / MagneticCycle blocks
  int n_per, _pad;
};

struct LoglParams {
  const EmpModelDesc* desc;   // device copy
  const char* tiles;          // packed tiles, kTileBytes each
  int64_t n_points;
  int32_t n_tiles;
  const double* theta;        // [n_eval, ndim_free]
  const int32_t* eval_index;  // compact list of rows to evaluate
  const int32_t* n_active;    // number of entries in eval_index (device)
  double* logl;
  double t0;                  // X_[0] (acc.model uses X_ - X_[0])
  double ll_const;            // -0.5*log(2*pi)*ndat  (00.like:1)
  double t_absmax;            // max |t|: bounds the mean anomaly per (walker, planet)
  const double2* grid_sc;     // [kGridN] (sin, cos)(k 2^-7), correctly rounded FP64
  const float4* grid_scf;     // (sin, cos, sin/2, cos/6) of the same points in FP32
  uint32_t tile_bytes;        // kTileBytes + sai_cols * kTilePoints * 8 (activity columns follow the instrument ids)
  int32_t sai_cols;           // activity columns per point (max over the instruments), 0 = none
  int32_t solver;             // EMP_SOLVER_GRID (default) | EMP_SOLVER_KEPLERPY: every planet takes kep_rv_robust
  int32_t* zero_counter;      // PT step: the compact-list counter of the OTHER half, zeroed here for its next use
  PtAccept pt;                // PT step: Metropolis accept in the epilogue (pt.enabled), emp_pt.cuh
  HotConsts H;                // FP64 literals of the hot loop, read as c[0x0][..] operands
};

// dynamic shared memory: tile ring | mbarriers | sin/cos grid (FP64 pairs, FP32 pairs) | walker constants
// (the ring's size depends on the number of activity columns, so the offsets are functions of tile_bytes)
__host__ __device__ constexpr size_t smem_bar_off(uint32_t tile_bytes) { return size_t(kStages) * tile_bytes; }
__host__ __device__ constexpr size_t smem_tab_off(uint32_t tile_bytes) { return smem_bar_off(tile_bytes) + 64; }
__host__ __device__ constexpr size_t smem_tabf_off(uint32_t tile_bytes) {
  return smem_tab_off(tile_bytes) + kGridN * sizeof(double2);
}
__host__ __device__ constexpr size_t smem_walker_off(uint32_t tile_bytes) {
  return smem_tabf_off(tile_bytes) + kGridN * sizeof(float4);
}
__host__ __device__ constexpr size_t smem_start_off(uint32_t tile_bytes) {
  return smem_walker_off(tile_bytes) + kWalkerWarps * sizeof(WalkerConst);
}
#ifndef EMP_STARTER_TABLE
#define EMP_STARTER_TABLE 1  // per-walker starter tables (emp_device.cuh); 0 builds the Markley-only kernel (A/B)
#endif
// the starter tables are used when the rest leaves room for them with two CTAs per SM (no activity columns)
__host__ __device__ constexpr bool logl_use_start_tables(uint32_t tile_bytes) {
  return EMP_STARTER_TABLE && tile_bytes == kTileBytes;
}
__host__ __device__ constexpr size_t logl_smem_bytes(uint32_t tile_bytes) {
  return smem_start_off(tile_bytes) +
         (logl_use_start_tables(tile_bytes) ? size_t(kWalkerWarps) * kStartFloats * sizeof(float) : 0);
}
constexpr uint32_t kTileBytesMax = kTileBytes + EMP_MAX_SAI * kTilePoints * 8;
static_assert(2 * kStages * sizeof(uint64_t) <= 64 && kTileBytes % 16 == 0 && (kTilePoints * 8) % 16 == 0,
              "shared-memory layout");

// theta[ndim_free] -> full theta in shared memory (emp_model.py:709-711)
__device__ __forceinline__ void load_full_theta(const EmpModelDesc* __restrict__ d,
                                                const double* __restrict__ theta_row, double* th, int lane) {
  const int nfull = d->ndim_full, nfree = d->ndim_free;
  for (int i = lane; i < nfull; i += 32) th[i] = d->full_init[i];
  __syncwarp();
  for (int j = lane; j < nfree; j += 32) th[d->free_to_full[j]] = theta_row[j];
  __syncwarp();
}

// Keplerian / instrument constants of one walker from its full theta
__device__ __forceinline__ void walker_constants(const EmpModelDesc* __restrict__ d, WalkerConst& wc, int lane,
                                                 double t_absmax) {
  if (lane < d->n_kep) {
    KepConst kc;
    kep_constants(d->kep_model[lane], wc.th + d->kep_off[lane], t_absmax, kc);
    wc.kep[lane] = kc;
  }
  if (lane < d->n_ins) {
    wc.gamma[lane] = wc.th[d->offset_off + lane];
    double s = d->has_jitter ? wc.th[d->jitter_off + lane] : 0.0;
    wc.jit2[lane] = __dmul_rn(s, s);  // theta ** 2 (jitter00.model:3)
  }
  if (lane < d->acc_order) wc.acc[lane] = wc.th[d->acc_off + lane];
  if (lane < 2 * d->ma_order) wc.ma[lane] = wc.th[d->ma_off + lane];
  if (lane < d->ma_order) wc.ma_itau[lane] = 1.0 / wc.th[d->ma_off + 2 * lane + 1];
  if (lane < EMP_MAX_MA) { wc.ma_rh[lane] = 0.0; wc.ma_thist[lane] = 0.0; }
  if (d->n_sai > 0) {
    // theta_sa[j] of sai00.model, regrouped per instrument: column c of instrument i is j = sum(count[:i]) + c
    for (int idx = lane; idx < d->n_ins * EMP_MAX_SAI; idx += 32) {
      const int i = idx / EMP_MAX_SAI, c = idx % EMP_MAX_SAI;
      int base = 0;
      for (int q = 0; q < i; ++q) base += d->sai_count[q];
      wc.sai[idx] = (c < d->sai_count[i]) ? wc.th[d->sai_off + base + c] : 0.0;
    }
  }
  if (lane == 0) {
    int n = 0;
    for (int b = 0; b < d->n_periodic; ++b) {
      const double* th = wc.th + d->periodic_off[b];
      if (d->periodic_kind[b] == 0) {  // sinusoid00.model: per, A, phase
        wc.per[n++] = make_periodic(kTwoPi / th[0], th[2], th[1], t_absmax);
      } else {                         // magneticcycle00.model: per, A1, A2, phase1, phase2
        wc.per[n++] = make_periodic(kPi / th[0], th[3], th[1], t_absmax);
        wc.per[n++] = make_periodic(kTwoPi / th[0], th[4], th[2], t_absmax);
      }
    }
    wc.n_per = n;
  }
  __syncwarp();
}

// ---- launch 1: prior + compaction ---------------------------------------------------------
constexpr int kPriorWarps = 8;
__global__ void __launch_bounds__(kPriorWarps * 32)
prior_compact_kernel(const EmpModelDesc* __restrict__ d, const double* __restrict__ theta, int64_t n_eval,
                     double* __restrict__ logl, double* __restrict__ logp, int32_t* __restrict__ eval_index,
                     int32_t* __restrict__ n_active) {
  __shared__ double th_s[kPriorWarps][EMP_MAX_DIM];
  __shared__ double val_s[kPriorWarps][EMP_MAX_PRIOR_OPS];
  __shared__ int s_flag[kPriorWarps];
  __shared__ int s_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t e = int64_t(blockIdx.x) * kPriorWarps + warp;
  bool inside = false;
  if (e < n_eval) {
    load_full_theta(d, theta + e * d->ndim_free, th_s[warp], lane);
    const double lp = prior_program_warp(d, th_s[warp], val_s[warp], lane);
    inside = !(lp == -INFINITY);
    if (lane == 0) {
      logp[e] = lp;
      if (!inside) logl[e] = -INFINITY;
    }
  }
  // rows inside the prior support are appended to the compact list: one atomic per CTA
  if (lane == 0) s_flag[warp] = inside ? 1 : 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = 0;
    for (int k = 0; k < kPriorWarps; ++k) n += s_flag[k];
    s_base = n ? atomicAdd(n_active, n) : 0;
  }
  __syncthreads();
  if (inside && lane == 0) {
    int r = 0;
    for (int k = 0; k < warp; ++k) r += s_flag[k];
    eval_index[s_base + r] = int32_t(e);
  }
}

// np.polyval([a_n .. a_1, 0], x) with NumPy's Horner roundings (acc.model:2)
__device__ __forceinline__ double accel_term(const double* acc, int order, double x) {
  double y = acc[0];
  for (int j = 1; j < order; ++j) y = __dadd_rn(__dmul_rn(y, x), acc[j]);
  return __dadd_rn(__dmul_rn(y, x), 0.0);
}

// ---- launch 2: likelihood -------------------------------------------------------------------
// per-lane running sums and the warp-uniform MA(1) state
struct LaneAcc {
  double chi;              // sum r^2 / err2
  double prod;             // running product of err2, mantissa kept in [1, 2)
  int esum;                // ... and the sum of the exponent fields taken out of it
  double r_carry, t_prev;  // MA(1) carry: previous residual and timestamp
};

// one step of the inclusive warp scan over affine maps r -> A r + B, predicated instead of selected
__device__ __forceinline__ void scan_step(double& A, double& B, double Ap, double Bp, int lane, int off) {
  asm("{\n"
      ".reg .pred p;\n"
      "setp.ge.s32 p, %4, %5;\n"
      "@p fma.rn.f64 %1, %0, %3, %1;\n"
      "@p mul.rn.f64 %0, %0, %2;\n"
      "}\n"
      : "+d"(A), "+d"(B)
      : "d"(Ap), "d"(Bp), "r"(lane), "r"(off));
}

// Model features the kernel is specialised on (one instantiation per combination the descriptor can ask
// for): the tail runs once per point, and its warp-uniform `if (acc_order)`, `if (ma_order == 1)` ... tests
// and loops cost a fifth of its instructions when they are decided at run time.
constexpr int kFeatAcc = 1;    // AccelerationBlock
constexpr int kFeatMa1 = 2;    // global MA(1) (moav01.model, order 1): warp scan
constexpr int kFeatMaN = 4;    // global MA of order >= 2: serial recurrence
constexpr int kFeatPost = 8;   // StellarActivity / Sinusoid / MagneticCycle terms after the MA block
constexpr int kNumFeat = 16;

// Everything after the Keplerian sum for one segment of 128 points (lane owns the 4 CONSECUTIVE points
// 4*lane .. 4*lane+3 of the segment, so that the MA recurrence needs one warp scan per 128 points):
// acceleration, offsets, jitter, MA recurrence, activity / periodic terms, chi^2 and log-det.
// kFull: every point of the tile is a data point (all tiles but the last one): no validity selects.
template <int kFeat, bool kFull>
__device__ __forceinline__ void tail128(WalkerConst& wc, const LoglParams& P, LaneAcc& A, int lane, int seg, int cnt,
                                        int64_t base, const unsigned char* tb, const double (&t)[4], double (&m)[4],
                                        int acc_order, int ma_order, int n_per) {
  const double2* ys = reinterpret_cast<const double2*>(tb + kTilePoints * 8);
  const double2* es = reinterpret_cast<const double2*>(tb + kTilePoints * 16);
  const int4* is = reinterpret_cast<const int4*>(tb + kTilePoints * 24);
  const int li = seg * 64 + 2 * lane;  // double2 index of the lane's first pair
  const int p0 = seg * 128 + 4 * lane;
  bool v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = kFull || (p0 + j) < cnt;
  if (kFeat & kFeatAcc) {
#pragma unroll
    for (int j = 0; j < 4; ++j) m[j] += accel_term(wc.acc, acc_order, __dsub_rn(t[j], P.t0));
  }
  const int4 in4 = is[seg * 32 + lane];
  const int in[4] = {in4.x, in4.y, in4.z, in4.w};
  const double2 ya = ys[li], yb = ys[li + 1], ea = es[li], eb = es[li + 1];
  const double y[4] = {ya.x, ya.y, yb.x, yb.y};
  const double e2[4] = {ea.x, ea.y, eb.x, eb.y};
  double d[4], w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    m[j] += wc.gamma[in[j]];
    d[j] = v[j] ? y[j] - m[j] : 0.0;
    w[j] = v[j] ? e2[j] + wc.jit2[in[j]] : 1.0;
  }

  if (kFeat & kFeatMa1) {
    // moav01.model: r_i = d_i - phi*exp(-|t_i - t_{i-1}|/tau) * r_{i-1}, sequential in i.
    // Evaluated as a warp scan over the affine maps r -> a*r + b (exact algebra).  The very first point
    // has no MA term (`if i > c`): its predecessor residual is the initial carry 0, so the term vanishes
    // by itself (a is finite: t_prev starts at 0).
    const double phi = wc.ma[0], itau = wc.ma_itau[0];
    const double tl = __shfl_up_sync(0xffffffffu, t[3], 1);
    double tp = (lane == 0) ? A.t_prev : tl;
    double a[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double ex = exp_neg(-fabs(t[j] - tp) * itau, P.H);
      a[j] = v[j] ? -phi * ex : 0.0;
      tp = t[j];
    }
    // compose the lane's four maps, then inclusive scan across lanes
    double Am = a[0], Bm = d[0];
#pragma unroll
    for (int j = 1; j < 4; ++j) { Bm = fma(a[j], Bm, d[j]); Am = a[j] * Am; }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const double Ap = __shfl_up_sync(0xffffffffu, Am, off);
      const double Bp = __shfl_up_sync(0xffffffffu, Bm, off);
      scan_step(Am, Bm, Ap, Bp, lane, off);
    }
    const double r_last = fma(Am, A.r_carry, Bm);  // residual at this lane's last point
    double r_prev = __shfl_up_sync(0xffffffffu, r_last, 1);
    if (lane == 0) r_prev = A.r_carry;
#pragma unroll
    for (int j = 0; j < 4; ++j) { d[j] = fma(a[j], r_prev, d[j]); r_prev = d[j]; }
    // carry to the next segment: last VALID point of this one
    if (kFull) {
      A.r_carry = __shfl_sync(0xffffffffu, d[3], 31);
      A.t_prev = __shfl_sync(0xffffffffu, t[3], 31);
    } else {
      const int last = min(127, cnt - seg * 128 - 1);  // >= 0: the caller skips empty segments
      const int lj = last & 3;
      const double dv = lj == 0 ? d[0] : lj == 1 ? d[1] : lj == 2 ? d[2] : d[3];
      const double tv = lj == 0 ? t[0] : lj == 1 ? t[1] : lj == 2 ? t[2] : t[3];
      A.r_carry = __shfl_sync(0xffffffffu, dv, last >> 2);
      A.t_prev = __shfl_sync(0xffffffffu, tv, last >> 2);
    }
  } else if (kFeat & kFeatMaN) {
    // general order: serial recurrence over the 128 points (rare configuration), every lane runs the
    // same uniform loop on shuffled values; the warp-uniform history lives in the walker's slot
    for (int j = 0; j < 128; ++j) {
      const int src = j >> 2, cj = j & 3;
      const double dj = __shfl_sync(0xffffffffu, cj == 0 ? d[0] : cj == 1 ? d[1] : cj == 2 ? d[2] : d[3], src);
      const double tj = __shfl_sync(0xffffffffu, cj == 0 ? t[0] : cj == 1 ? t[1] : cj == 2 ? t[2] : t[3], src);
      const bool vj = (seg * 128 + j) < cnt;
      const int64_t gi = base + seg * 128 + j;
      double r = dj;
      if (vj) {
        for (int c = 0; c < ma_order; ++c)
          if (gi > c) r -= wc.ma[2 * c] * exp(-fabs(tj - wc.ma_thist[c]) / wc.ma[2 * c + 1]) * wc.ma_rh[c];
        __syncwarp();
        if (lane == 0) {
          for (int c = EMP_MAX_MA - 1; c > 0; --c) { wc.ma_rh[c] = wc.ma_rh[c - 1]; wc.ma_thist[c] = wc.ma_thist[c - 1]; }
          wc.ma_rh[0] = r;
          wc.ma_thist[0] = tj;
        }
        __syncwarp();
        if (src == lane) {
          if (cj == 0) d[0] = r; else if (cj == 1) d[1] = r; else if (cj == 2) d[2] = r; else d[3] = r;
        }
      }
    }
  }

  // StellarActivity, Sinusoid and MagneticCycle blocks come after the MA block in the reference's model
  // (emp.py:2636-2650): they are not part of the MA residuals, only of the final one.
  // sai00.model: model0 += theta_sa[j] * SAI{j}_ — the tile carries, per point, the columns of ITS instrument
  if (kFeat & kFeatPost) {
    for (int c = 0; c < P.sai_cols; ++c) {
      const double2* sc = reinterpret_cast<const double2*>(tb + kTileBytes + size_t(c) * kTilePoints * 8);
      const double2 sa = sc[li], sb = sc[li + 1];
      d[0] = fma(-wc.sai[in[0] * EMP_MAX_SAI + c], sa.x, d[0]);  // padding rows carry 0
      d[1] = fma(-wc.sai[in[1] * EMP_MAX_SAI + c], sa.y, d[1]);
      d[2] = fma(-wc.sai[in[2] * EMP_MAX_SAI + c], sb.x, d[2]);
      d[3] = fma(-wc.sai[in[3] * EMP_MAX_SAI + c], sb.y, d[3]);
    }
    for (int q = 0; q < n_per; ++q) {
      const PeriodicTerm& pt = wc.per[q];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (v[j]) d[j] -= periodic_value(pt, t[j], P.H);
    }
  }

  // chi^2: r0^2/w0 + r1^2/w1 over the common denominator (one reciprocal per pair);
  // log-det: sum(log err2) (00.like:5) as the log of a running product whose exponent is moved into
  // an integer sum after every segment, so the only log() is the one after the last tile
  const double w01 = w[0] * w[1], w23 = w[2] * w[3];
  A.chi = fma(fma(d[0] * d[0], w[1], (d[1] * d[1]) * w[0]), rcp_nr<2>(w01), A.chi);
  A.chi = fma(fma(d[2] * d[2], w[3], (d[3] * d[3]) * w[2]), rcp_nr<2>(w23), A.chi);
  const double pr = A.prod * (w01 * w23);
  const int hi = __double2hiint(pr);
  A.esum += hi >> 20;
  A.prod = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(pr));
}

// kGroups = groups of 64 points a warp works on at once: 2*kGroups independent Kepler chains per lane (A/B
// builds showed 2 is the sweet spot; the tail is written for exactly that: 128-point segments)
template <int kGroups, int kFeat>
__global__ void __launch_bounds__(kLoglThreads, 2) logl_rv_kernel(const LoglParams P) {
  static_assert(kGroups == 2, "the tail works on 128-point segments");
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* tiles_s = smem;
  const uint32_t tile_bytes = P.tile_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + smem_bar_off(tile_bytes));
  uint64_t* empty_bar = full_bar + kStages;
  double2* tab = reinterpret_cast<double2*>(smem + smem_tab_off(tile_bytes));
  float4* tabf = reinterpret_cast<float4*>(smem + smem_tabf_off(tile_bytes));
  WalkerConst* wcs = reinterpret_cast<WalkerConst*>(smem + smem_walker_off(tile_bytes));
  const bool use_tab = logl_use_start_tables(tile_bytes);
  float* start_tab = reinterpret_cast<float*>(smem + smem_start_off(tile_bytes)) +
                     size_t(threadIdx.x >> 5) * kStartFloats;

  const int n_active = *P.n_active;
  const int first = blockIdx.x * kWalkerWarps;
  if (P.zero_counter && blockIdx.x == 0 && threadIdx.x == 0) *P.zero_counter = 0;
  if (first >= n_active) return;  // whole CTA beyond the compact list

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const EmpModelDesc* __restrict__ d = P.desc;
  const int n_tiles = P.n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kWalkerWarps);
    }
    fence_barrier_init();
    // prime the ring
    const int pre = n_tiles < kStages ? n_tiles : kStages;
    for (int i = 0; i < pre; ++i) {
      mbar_arrive_expect_tx(&full_bar[i], tile_bytes);
      tma_bulk_g2s(tiles_s + size_t(i) * tile_bytes, P.tiles + size_t(i) * tile_bytes, tile_bytes, &full_bar[i]);
    }
  }
  // sin/cos grid of the Kepler core (12 KB, L2-resident)
  for (int i = threadIdx.x; i < kGridN; i += kLoglThreads) {
    tab[i] = P.grid_sc[i];
    tabf[i] = P.grid_scf[i];
  }

  // ---- prologue: per-walker constants -------------------------------------------------------
  const bool active = (first + warp) < n_active;
  int64_t slot = -1;
  WalkerConst& wc = wcs[warp];
  if (active) {
    slot = P.eval_index[first + warp];
    load_full_theta(d, P.theta + slot * d->ndim_free, wc.th, lane);
    walker_constants(d, wc, lane, P.t_absmax);
    if (P.solver == EMP_SOLVER_KEPLERPY && lane < d->n_kep) wc.kep[lane].robust = 1;
    __syncwarp();
    if (use_tab) {  // starter tables of the planets the grid core serves (wc.th is free again: scratch)
      const int nk = d->n_kep < kStartPlanets ? d->n_kep : kStartPlanets;
      for (int k = 0; k < nk; ++k) {
        KepConst& kc = wc.kep[k];
        const bool ok = (kc.slow_mod | kc.robust) == 0 && kc.ef <= kStartEccMax;  // warp-uniform
        if (ok) build_start_table(kc, start_tab + k * 3 * kStartStride, reinterpret_cast<float*>(wc.th), lane);
        if (lane == 0) kc.tab = ok ? 1 : 0;
      }
      __syncwarp();
    }
  }
  __syncthreads();  // publishes the barrier inits and the grid to all warps

  const int K = d->n_kep;
  const int acc_order = d->acc_order;
  const int ma_order = (d->ma_mode == EMP_MA_GLOBAL) ? d->ma_order : 0;
  const int n_per = active ? wc.n_per : 0;

  LaneAcc A;
  A.chi = 0.0; A.prod = 1.0; A.esum = 0; A.r_carry = 0.0; A.t_prev = 0.0;
  int n_pairs = 0;  // segments folded into the running product (warp-uniform)

  for (int i = 0; i < n_tiles; ++i) {
    const int s = i % kStages;
    // refill: the stage of tile i-2 was released by every warp unless one lags two tiles behind
    // -> load tile i+1 into it (one tile of prefetch distance; a tile is >100 us of math)
    if (threadIdx.x == 0 && i >= 2 && (i - 2 + kStages) < n_tiles) {
      const int sp = (i - 2) % kStages;
      mbar_wait(&empty_bar[sp], ((i - 2) / kStages) & 1);
      mbar_arrive_expect_tx(&full_bar[sp], tile_bytes);
      tma_bulk_g2s(tiles_s + size_t(sp) * tile_bytes, P.tiles + size_t(i - 2 + kStages) * tile_bytes, tile_bytes,
                   &full_bar[sp]);
    }
    __syncwarp();
    mbar_wait(&full_bar[s], (i / kStages) & 1);
    if (active) {
      const unsigned char* tb = tiles_s + size_t(s) * tile_bytes;
      const double2* ts = reinterpret_cast<const double2*>(tb);
      const int64_t base = int64_t(i) * kTilePoints;
      const int64_t rem = P.n_points - base;
      const int cnt = rem < kTilePoints ? int(rem) : kTilePoints;
      const int segs = (cnt + 127) >> 7;
      for (int seg = 0; seg < segs; ++seg) {
        // lane owns points 4*lane .. 4*lane+3 of the segment: two 128-bit reads 32 bytes apart (a 2-way bank
        // conflict on 4 reads per segment; the tile's padding replicates the last timestamp)
        const double2 ta = ts[seg * 64 + 2 * lane], tc = ts[seg * 64 + 2 * lane + 1];
        const double t[4] = {ta.x, ta.y, tc.x, tc.y};
        double m[4] = {0.0, 0.0, 0.0, 0.0};
        for (int k = 0; k < K; ++k) {
          const KepConst& kc = wc.kep[k];
          if (kc.tab) {  // warp-uniform: starter from the walker's table
            const float* st = start_tab + k * 3 * kStartStride;
#pragma unroll
            for (int j = 0; j < 4; ++j) m[j] = kep_rv_grid<true>(kc, t[j], m[j], P.H, tab, tabf, st);
          } else if ((kc.slow_mod | kc.robust) == 0) {  // warp-uniform
#pragma unroll
            for (int j = 0; j < 4; ++j)  // straight-line code: the scheduler interleaves the four chains
              m[j] = kep_rv_grid(kc, t[j], m[j], P.H, tab, tabf);
          } else {
            const Quad r = kep_rv_robust4(kc, t[0], t[1], t[2], t[3]);
            m[0] += r.v0; m[1] += r.v1; m[2] += r.v2; m[3] += r.v3;
          }
        }
        if (cnt == kTilePoints)  // warp-uniform: every tile but the last
          tail128<kFeat, true>(wc, P, A, lane, seg, cnt, base, tb, t, m, acc_order, ma_order, n_per);
        else
          tail128<kFeat, false>(wc, P, A, lane, seg, cnt, base, tb, t, m, acc_order, ma_order, n_per);
        ++n_pairs;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }

  if (active) {
    const double lsum = fma(double(A.esum - 1023 * n_pairs), 0.693147180559945309417, log(A.prod));
    const double tot = warp_sum(A.chi + lsum);  // xor butterfly: every lane holds the same sum
    const double ll = fma(-0.5, tot, P.ll_const);
    if (P.pt.enabled) pt_accept_row(P.pt, d->ndim_free, slot, ll, lane);  // slot = proposal index of the half
    else if (lane == 0) P.logl[slot] = ll;
  }
}

// host side: the feature mask of a descriptor and the instantiation that serves it
inline int logl_features(const EmpModelDesc& d) {
  int f = 0;
  if (d.acc_order > 0) f |= kFeatAcc;
  if (d.ma_mode == EMP_MA_GLOBAL && d.ma_order == 1) f |= kFeatMa1;
  if (d.ma_mode == EMP_MA_GLOBAL && d.ma_order >= 2) f |= kFeatMaN;
  if (d.n_periodic > 0 || d.n_sai > 0) f |= kFeatPost;
  return f;
}
using LoglKernel = void (*)(const LoglParams);
template <int kGroups, int kFeat>
struct LoglKernelTable {
  static void fill(LoglKernel* tab) {
    // MA(1) and MA(order >= 2) exclude each other: those slots stay null
    if constexpr ((kFeat & kFeatMa1) && (kFeat & kFeatMaN)) tab[kFeat] = nullptr;
    else tab[kFeat] = logl_rv_kernel<kGroups, kFeat>;
    LoglKernelTable<kGroups, kFeat + 1>::fill(tab);
  }
};
template <int kGroups>
struct LoglKernelTable<kGroups, kNumFeat> {
  static void fill(LoglKernel*) {}
};

// my_model(theta) for one theta: model0[n], err20[n]  (emp_model.py:706-781); thread per point
// for the first pass, MA recurrence (if any) applied serially by one thread afterwards.
__global__ void model_rv_kernel(const EmpModelDesc* __restrict__ d, const double* __restrict__ theta,
                                const double* __restrict__ t, const double* __restrict__ y,
                                const double* __restrict__ e2, const int32_t* __restrict__ ins, int64_t n,
                                double t0, double t_absmax, double* __restrict__ model, double* __restrict__ err2,
                                const double* __restrict__ sai, int sai_cols, const HotConsts H) {
  __shared__ WalkerConst wc;
  if (threadIdx.x < 32) {
    load_full_theta(d, theta, wc.th, threadIdx.x);
    walker_constants(d, wc, threadIdx.x, t_absmax);
  }
  __syncthreads();
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    double m = 0.0;
    for (int k = 0; k < d->n_kep; ++k) m += kep_rv_checked(wc.kep[k], t[i], H);
    if (d->acc_order > 0) m += accel_term(wc.acc, d->acc_order, __dsub_rn(t[i], t0));
    m += wc.gamma[ins[i]];
    if (d->ma_mode != EMP_MA_GLOBAL || d->ma_order == 0) {  // else added after the MA pass (model_periodic_kernel)
      for (int c = 0; c < sai_cols; ++c) m += wc.sai[ins[i] * EMP_MAX_SAI + c] * sai[size_t(c) * n + i];
      for (int q = 0; q < wc.n_per; ++q) m += periodic_value(wc.per[q], t[i], H);
    }
    model[i] = m;
    err2[i] = e2[i] + wc.jit2[ins[i]];
  }
}

// activity / periodic terms of a model with a global MA block: added after model_ma_kernel
__global__ void model_periodic_kernel(const EmpModelDesc* __restrict__ d, const double* __restrict__ theta,
                                      const double* __restrict__ t, const int32_t* __restrict__ ins, int64_t n,
                                      double t_absmax, double* __restrict__ model, const double* __restrict__ sai,
                                      int sai_cols, const HotConsts H) {
  __shared__ WalkerConst wc;
  if (threadIdx.x < 32) {
    load_full_theta(d, theta, wc.th, threadIdx.x);
    walker_constants(d, wc, threadIdx.x, t_absmax);
  }
  __syncthreads();
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    for (int c = 0; c < sai_cols; ++c) model[i] += wc.sai[ins[i] * EMP_MAX_SAI + c] * sai[size_t(c) * n + i];
    for (int q = 0; q < wc.n_per; ++q) model[i] += periodic_value(wc.per[q], t[i], H);
  }
}

__global__ void model_ma_kernel(const EmpModelDesc* __restrict__ d, const double* __restrict__ theta,
                                const double* __restrict__ t, const double* __restrict__ y, int64_t n,
                                double* __restrict__ model) {
  // moav01.model:3-15, serial (post-processing helper, one theta)
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const int order = d->ma_order;
  double th_ma[2 * EMP_MAX_MA];
  double full[EMP_MAX_DIM];
  for (int i = 0; i < d->ndim_full; ++i) full[i] = d->full_init[i];
  for (int j = 0; j < d->ndim_free; ++j) full[d->free_to_full[j]] = theta[j];
  for (int c = 0; c < 2 * order; ++c) th_ma[c] = full[d->ma_off + c];
  double rh[EMP_MAX_MA], thist[EMP_MAX_MA];
  for (int c = 0; c < EMP_MAX_MA; ++c) { rh[c] = 0.0; thist[c] = 0.0; }
  for (int64_t i = 0; i < n; ++i) {
    double r = y[i] - model[i];
    double m = model[i];
    for (int c = 0; c < order; ++c) {
      if (i > c) {
        const double ma = th_ma[2 * c] * exp(-fabs(t[i] - thist[c]) / th_ma[2 * c + 1]) * rh[c];
        m += ma;
        r -= ma;
      }
    }
    for (int c = EMP_MAX_MA - 1; c > 0; --c) { rh[c] = rh[c - 1]; thist[c] = thist[c - 1]; }
    rh[0] = r;
    thist[0] = t[i];
    model[i] = m;
  }
}

}  // namespace emp
