// emp_pt.cuh — parallel-tempering step kernels (§8f row N1).
//
// The sampler arithmetic of the reference lives in reddemcee / emcee 3.1.6 (not vendored);
// semantics restated in oracle/pt_oracle.py.  All random numbers are host-supplied so that
// the accept / swap decisions are a pure function of (state, draws):
//   stretch move  : q = c[rint] - (c[rint] - s) * zz          (emcee StretchMove.get_proposal)
//   accept        : factors + (beta*ll' + lp') - (beta*ll + lp) > ln u   (emcee RedBlueMove.propose,
//                   tempered as in ptemcee/reddemcee)
//   swap (i,i-1)  : (beta_{i-1} - beta_i) * (ll_i[perm_i] - ll_{i-1}[perm_{i-1}]) > ln u, hot -> cold
//   ladder        : Vousden, Farr & Mandel (2016) dynamics with reddemcee's adapt_tau / adapt_nu
// Every floating-point expression below uses explicit single-rounding intrinsics (no FMA
// contraction) so that, given identical log-likelihoods, the decisions are bit-identical to
// the NumPy oracle.
//
// A sweep with nsteps = 1 is SIX launches (round 1: 13 + torch glue + a host synchronisation):
//   per half-ensemble  pt_propose_prior_kernel    proposal + prior program + compaction
//                      logl_rv_kernel             likelihood + Metropolis accept in its epilogue
//   per sweep          pt_swap_plan_chain_kernel  hot->cold swap plan as W independent chains + ladder adaptation
//                                                 + histories (the single-CTA shared-memory kernels pt_swap_plan_*
//                                                 remain for pairs that are not listed by their warmer slot)
//                      pt_apply_plan_kernel       row gather (local or peer HBM over NVLink) + swap mean
//                                                 distance + chain store
// A ladder sharded over several GPUs adds ONE launch and no NCCL call: pt_publish_kernel writes the rank's rows of
// logL and of the swap draws into every peer's gathered block over NVLink and raises a flag; the plan kernel waits
// for the flags of the sweep.
#pragma once
#include <stdint.h>
#include "emp_device.cuh"

namespace emp {

constexpr int kMaxPeers = 16;  // GPUs a ladder can be sharded over (one NVSwitch domain)

// ---- Metropolis accept, run by the warp that just finished a proposal's likelihood ---------------
struct PtAccept {
  int32_t enabled;
  int32_t T, W, split;        // local temperatures, walkers, half being updated
  double* p;                  // [T, W, ndim] state (updated in place)
  double* logl;               // [T, W]
  double* logp;               // [T, W]
  const double* q;            // [T*H, ndim] proposals of this half
  const double* lpq;          // [T*H] their log-priors
  const int32_t* half_idx;    // [T, 2, H]
  const double* betas;        // beta of local row t = betas[beta_off + t*beta_stride] (the whole ladder's array)
  int32_t beta_off, beta_stride;
  const double* factors;      // [T, 2, H] (ndim-1) ln zz
  const double* lnu;          // [T, 2, H] ln u
  uint8_t* accepted;          // [T, W]
  int32_t* n_accepted;        // [T, W] accepted moves per walker since the run began (may be NULL)
  unsigned long long* cnt;    // [0] proposals [1] inside the prior support [2] accepted
  uint32_t* n_nan;
};

// tj: index of the proposal inside the half (t*H + j); ll_new known to every lane
__device__ __forceinline__ void pt_accept_row(const PtAccept& A, int ndim, int64_t tj, double ll_new, int lane) {
  const int32_t H = A.W / 2;
  const int32_t j = int32_t(tj % H), t = int32_t(tj / H);
  const int64_t o = (int64_t(t) * 2 + A.split) * H + j;
  const int32_t i = A.half_idx[o];
  const int64_t w = int64_t(t) * A.W + i;
  const double beta = A.betas[A.beta_off + t * A.beta_stride];
  const double lp_new = A.lpq[tj];
  const double ll_old = A.logl[w], lp_old = A.logp[w];
  // lnpdiff = f + (beta*ll' + lp') - (beta*ll + lp)
  const double post_new = __dadd_rn(__dmul_rn(beta, ll_new), lp_new);
  const double post_old = __dadd_rn(__dmul_rn(beta, ll_old), lp_old);
  const double lnpdiff = __dsub_rn(__dadd_rn(A.factors[o], post_new), post_old);
  const bool acc = lnpdiff > A.lnu[o];
  if (lane == 0) {
    A.accepted[w] = acc ? 1 : 0;
    if (ll_new != ll_new) atomicAdd(A.n_nan, 1u);  // emcee raises on a NaN likelihood; here: rejected + counted
    if (acc) {
      atomicAdd(&A.cnt[2], 1ull);
      if (A.n_accepted) A.n_accepted[w] += 1;
    }
  }
  if (acc) {  // red/blue: only this warp touches walker w during this half-step
    for (int dd = lane; dd < ndim; dd += 32) A.p[w * ndim + dd] = A.q[tj * ndim + dd];
    if (lane == 0) { A.logl[w] = ll_new; A.logp[w] = lp_new; }
  }
}

// ---- proposal + prior + compaction ------------------------------------------------------------------
struct PtPropose {
  const EmpModelDesc* desc;
  const double* p;            // [T, W, ndim]
  int32_t T, W, split;
  const int32_t* half_idx;    // [T, 2, H]
  const double* zz;           // [T, 2, H]
  const int32_t* rint;        // [T, 2, H]
  double* q;                  // [T*H, ndim] out
  double* lpq;                // [T*H] out
  int32_t* eval_index;        // compact list of the proposals inside the prior support
  int32_t* n_active;          // its length (zero on entry; the likelihood kernel of the OTHER half re-zeroes it)
  uint8_t* accepted;          // [T, W]: 0 written for the proposals that are never evaluated
  unsigned long long* cnt;
  long long* step_counter;    // stretch steps begun so far (device scalar; split 0 increments it)
};

constexpr int kProposeWarps = 8;

// One warp per proposal.  q = c[rint] - (c[rint] - s) zz with the reference's roundings; the prior
// program is evaluated lane-parallel (one op per lane) and summed in program order by lane 0 — the
// reference's early `return lp` after a block whose sum is -inf is reproduced by the ordered sum, so
// logP is bit-identical to my_prior.  Proposals outside the prior support are never evaluated (emcee
// semantics): they get accepted = 0 here; the others are appended to the compact list with one atomic
// per CTA.
__global__ void __launch_bounds__(kProposeWarps * 32) pt_propose_prior_kernel(const PtPropose A) {
  __shared__ double th_s[kProposeWarps][EMP_MAX_DIM];
  __shared__ double val_s[kProposeWarps][EMP_MAX_PRIOR_OPS];
  __shared__ int s_flag[kProposeWarps];
  __shared__ int s_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const EmpModelDesc* __restrict__ d = A.desc;
  const int32_t H = A.W / 2, ndim = d->ndim_free;
  const int64_t n_prop = int64_t(A.T) * H;
  const int64_t tj = int64_t(blockIdx.x) * kProposeWarps + warp;
  const bool valid = tj < n_prop;
  bool inside = false;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    atomicAdd(&A.cnt[0], (unsigned long long)n_prop);
    if (A.split == 0 && A.step_counter) *A.step_counter += 1;
  }
  if (valid) {
    const int32_t j = int32_t(tj % H), t = int32_t(tj / H);
    const int64_t o = (int64_t(t) * 2 + A.split) * H + j;
    const int32_t i = A.half_idx[o];
    const int32_t partner = A.half_idx[(int64_t(t) * 2 + (1 - A.split)) * H + A.rint[o]];
    const double z = A.zz[o];
    double* th = th_s[warp];
    for (int k = lane; k < d->ndim_full; k += 32) th[k] = d->full_init[k];
    __syncwarp();
    for (int dd = lane; dd < ndim; dd += 32) {
      const double s = A.p[(int64_t(t) * A.W + i) * ndim + dd];
      const double c = A.p[(int64_t(t) * A.W + partner) * ndim + dd];
      const double qv = __dsub_rn(c, __dmul_rn(__dsub_rn(c, s), z));  // c[rint] - (c[rint] - s) * zz[:, None]
      A.q[tj * ndim + dd] = qv;
      th[d->free_to_full[dd]] = qv;
    }
    __syncwarp();
    const double lp = prior_program_warp(d, th, val_s[warp], lane);
    if (lane == 0) A.lpq[tj] = lp;
    inside = !(lp == -INFINITY);
    if (!inside && lane == 0) A.accepted[int64_t(t) * A.W + i] = 0;
  }
  if (lane == 0) s_flag[warp] = (valid && inside) ? 1 : 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = 0;
    for (int k = 0; k < kProposeWarps; ++k) n += s_flag[k];
    s_base = n ? atomicAdd(A.n_active, n) : 0;
    if (n) atomicAdd(&A.cnt[1], (unsigned long long)n);
  }
  __syncthreads();
  if (valid && inside && lane == 0) {
    int r = 0;
    for (int k = 0; k < warp; ++k) r += s_flag[k];
    A.eval_index[s_base + r] = int32_t(tj);
  }
}

// ---- deterministic exp for the ladder adaptation ---------------------------------------------------
// exp(x) as a FIXED sequence of individually rounded IEEE operations (no FMA), restated operation by
// operation in oracle/pt_oracle.py::exp_det, so that the device ladder equals the oracle's bit for bit.
// |error| <= 1 ulp; NumPy's own exp is neither correctly rounded nor the same on every host (SVML/AVX-512
// vs libm builds), so "the reference's np.exp" is only defined to that level anyway.
__device__ __forceinline__ double exp_det(double x) {
  const double kLog2e = 1.4426950408889634074, kL1 = 6.93147180369123816490e-01, kL2 = 1.90821492927058770002e-10;
  const double n = rint(__dmul_rn(x, kLog2e));
  const double r = __dsub_rn(__dsub_rn(x, __dmul_rn(n, kL1)), __dmul_rn(n, kL2));
  // 1/13! ... 1/2!
  const double c[12] = {1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0,
                        1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5};
  double pz = c[0];
#pragma unroll
  for (int k = 1; k < 12; ++k) pz = __dadd_rn(__dmul_rn(pz, r), c[k]);
  const double y = __dadd_rn(1.0, __dadd_rn(r, __dmul_rn(__dmul_rn(r, r), pz)));
  return scalbn(y, int(n));
}

// ---- swap plan + ladder adaptation -------------------------------------------------------------------
struct PtPlan {
  int32_t T, W;
  const double* logl;       // [T, W] log-likelihood of the whole ladder
  double* betas;            // [T] in: ladder of this sweep; out: adapted ladder (when adapt != 0)
  const int32_t* perm;      // [T-1, 2, W]
  const double* lnu;        // [T-1, W]
  int32_t* src;             // [T, W] out: flat index of the slot whose walker ends in (t, w)
  int32_t* n_acc;           // [T-1] out: accepted swaps per pair
  int32_t adapt;            // 1: Vousden ladder dynamics after the sweep (ntemps > 2)
  int32_t hot_sorted;       // 1: perm[j, 0, :] is the identity (pairs listed by their slot in the warmer row)
  double adapt_tau, adapt_nu;
  long long* sweep_counter; // device scalar: sweeps finished (the reference's `time`); incremented here
  double* beta_hist;        // [hist_cap, T] ladder after each sweep (may be NULL)
  int32_t* nacc_hist;       // [hist_cap, T-1] swap counts of each sweep (may be NULL)
  long long hist_cap;
  // sharded ladder with the peer-push exchange (pt_publish_kernel): logl / lnu / the partner rows are read from
  // the gathered block of the sweep's parity once every rank's flag says its rows have landed (chain kernel only)
  const unsigned char* gath[2];
  int32_t n_ranks;
};

// ---- gathered block of a sharded ladder (peer-push exchange) ---------------------------------------------------
// Every rank owns two such blocks (sweep parity), mapped into every peer with CUDA IPC.  After the stretch phase a
// rank WRITES its rows of logL and of the swap draws straight into every peer's block over NVLink
// (pt_publish_kernel) and then raises its flag there; the plan kernel of a rank waits for all flags of the
// sweep.  No NCCL call, no staging copy, rows land in ladder order.
//   [logl  T_all x W f64 | lnu  T_all x W f64 | partner  T_all x W i32 | flags  kMaxPeers i64]
__host__ __device__ inline size_t gath_lnu_off(int T, int W) { return size_t(T) * W * 8; }
__host__ __device__ inline size_t gath_partner_off(int T, int W) { return size_t(T) * W * 16; }
__host__ __device__ inline size_t gath_flags_off(int T, int W) { return (size_t(T) * W * 20 + 127) / 128 * 128; }
__host__ __device__ inline size_t gath_bytes(int T, int W) { return gath_flags_off(T, W) + 16 * 8; }

__device__ __forceinline__ long long ld_acquire_sys(const long long* p) {
  long long v;
  asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(long long* p, long long v) {
  asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct PtPublish {
  int32_t T_loc, W, T_all, G, rank, strided;
  const double* logl;        // [T_loc, W] this rank's log-likelihoods after the stretch phase
  const int32_t* perm;       // [T_loc, 2, W] the swap pairs this rank drew (row t: pair index = its temperature)
  const double* lnu;         // [T_loc, W]
  unsigned char* peer[2][kMaxPeers];  // gathered blocks (parity 0 / 1) of every rank, own block included
  const long long* sweep_counter;     // sweeps finished: parity and flag value of this sweep
  uint32_t* ticket;
};

constexpr int kPublishThreads = 256;
__global__ void __launch_bounds__(kPublishThreads) pt_publish_kernel(const PtPublish A) {
  __shared__ int s_last;
  const long long k = *A.sweep_counter;
  const int q = int(k & 1);
  const int t = blockIdx.y;
  const int w = blockIdx.x * kPublishThreads + threadIdx.x;
  const int j = A.strided ? t * A.G + A.rank : A.rank * A.T_loc + t;  // temperature (= pair index) of local row t
  if (w < A.W) {
    const double ll = A.logl[int64_t(t) * A.W + w];
    const double lu = A.lnu[int64_t(t) * A.W + w];
    const int32_t pb = A.perm[(int64_t(t) * 2 + 1) * A.W + w];
    const size_t o = size_t(j) * A.W + w;
    const size_t o_lnu = gath_lnu_off(A.T_all, A.W), o_par = gath_partner_off(A.T_all, A.W);
    for (int r = 0; r < A.G; ++r) {
      unsigned char* g = A.peer[q][(A.rank + r) % A.G];  // start at home, then round the peers
      reinterpret_cast<double*>(g)[o] = ll;
      reinterpret_cast<double*>(g + o_lnu)[o] = lu;
      reinterpret_cast<int32_t*>(g + o_par)[o] = pb;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0)
    s_last = (atomicAdd(A.ticket, 1u) == gridDim.x * gridDim.y - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (threadIdx.x < A.G) {
    long long* f = reinterpret_cast<long long*>(A.peer[q][threadIdx.x] + gath_flags_off(A.T_all, A.W)) + A.rank;
    st_release_sys(f, k + 1);
  }
  if (threadIdx.x == 0) *A.ticket = 0;
}

// The adaptation of oracle/pt_oracle.py::adapt_ladder (same operations, same order):
//   decay = tau/(time+tau); kappa = decay/nu; dS_i = kappa (A_i - A_{i+1}); dT_i = (1/b_{i+1} - 1/b_i) exp(dS_i);
//   b_{i+1} = 1/(cumsum(dT)_i + 1/b_0),  i = 0 .. T-3
// s_x: T doubles of shared scratch.  Called by all threads of the single plan CTA after the last pair.
__device__ __forceinline__ void plan_tail(const PtPlan& A, double* s_x, const int32_t* s_nacc) {
  const int tid = threadIdx.x, nt = blockDim.x, T = A.T;
  const long long row = A.sweep_counter ? *A.sweep_counter : 0;
  const long long time = row + 1;  // `self.time += 1` precedes the adaptation
  __syncthreads();
  if (A.adapt && T > 2) {
    const double decay = __ddiv_rn(A.adapt_tau, __dadd_rn(double(time), A.adapt_tau));
    const double kappa = __ddiv_rn(decay, A.adapt_nu);
    const double Wd = double(A.W);
    for (int i = tid; i < T - 2; i += nt) {
      const double r0 = __ddiv_rn(double(s_nacc[i]), Wd), r1 = __ddiv_rn(double(s_nacc[i + 1]), Wd);
      const double dS = __dmul_rn(kappa, __dsub_rn(r0, r1));
      const double dT = __dsub_rn(__ddiv_rn(1.0, A.betas[i + 1]), __ddiv_rn(1.0, A.betas[i]));
      s_x[i] = __dmul_rn(dT, exp_det(dS));
    }
    __syncthreads();
    if (tid == 0) {  // np.cumsum: strictly sequential
      double c = 0.0;
      for (int i = 0; i < T - 2; ++i) {
        c = (i == 0) ? s_x[0] : __dadd_rn(c, s_x[i]);
        s_x[i] = c;
      }
    }
    __syncthreads();
    const double t0 = __ddiv_rn(1.0, A.betas[0]);
    for (int i = tid; i < T - 2; i += nt) A.betas[i + 1] = __ddiv_rn(1.0, __dadd_rn(s_x[i], t0));
    __syncthreads();
  }
  if (row < A.hist_cap) {
    if (A.beta_hist)
      for (int i = tid; i < T; i += nt) A.beta_hist[row * T + i] = A.betas[i];
    if (A.nacc_hist)
      for (int i = tid; i < T - 1; i += nt) A.nacc_hist[row * (T - 1) + i] = s_nacc[i];
  }
  if (tid == 0 && A.sweep_counter) *A.sweep_counter = row + 1;
}

constexpr int kPlanMaxT = 2048;  // ladder length limit: the sweep's swap counts live in shared memory for the adaptation

// Swap plan: sequential over temperature pairs (hot -> cold), parallel over walkers, single CTA: every rank
// of a sharded ladder replays all T-1 pairs, so this kernel's latency is the serial (Amdahl) part of a sweep.
// The active logL rows and their source-index rows live in shared memory.
//   kBuf = 3 (36 W bytes): ONE block barrier per pair.  The colder row of the NEXT pair is staged into the spare
//     buffer while the current pair is decided (nothing reads it yet), the finished hot row is written back and
//     recycled as the next spare by the thread that owns the same elements.
//   kBuf = 2 (24 W bytes, W > 5600): the cold row is staged in place, which costs a second barrier per pair.
// The perm rows / uniforms of the next pair and the logL row after it are prefetched into registers under the
// decision of the current pair; swap counts are warp-aggregated (one shared-memory atomic per warp).
// kR = elements per thread = ceil(W / 1024).
template <int kR, int kBuf>
__global__ void __launch_bounds__(1024) pt_swap_plan_kernel(const PtPlan A) {
  extern __shared__ __align__(128) unsigned char plan_smem[];
  const int32_t T = A.T, W = A.W;
  // roles rotate by pointer: hot = warmer row of the current pair, cold = its colder row, spare = staging
  double* lh = reinterpret_cast<double*>(plan_smem);
  double* lc = lh + W;
  double* ls = lc + (kBuf == 3 ? W : 0);
  int32_t* sh = reinterpret_cast<int32_t*>(ls + W);
  int32_t* sc = sh + W;
  int32_t* ss = sc + (kBuf == 3 ? W : 0);
  __shared__ int32_t s_count[2];
  __shared__ int32_t s_nacc[kPlanMaxT];
  __shared__ double s_x[kPlanMaxT];
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
  const double* __restrict__ logl = A.logl;
  const int32_t* __restrict__ perm = A.perm;
  const double* __restrict__ lnu = A.lnu;

  int32_t ca[kR], cb[kR];
  double cu[kR], nl[kR];
  // hot row = temperature T-1; cold row of the first pair = temperature T-2
  for (int w = tid; w < W; w += nt) {
    lh[w] = logl[int64_t(T - 1) * W + w];
    sh[w] = (T - 1) * W + w;
    if (T > 1) {
      lc[w] = logl[int64_t(T - 2) * W + w];
      sc[w] = (T - 2) * W + w;
    }
  }
  if (tid < 2) s_count[tid] = 0;
  if (T > 1) {
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      const int k = tid + r * nt;
      if (k < W) {
        ca[r] = perm[(int64_t(T - 2) * 2 + 0) * W + k];
        cb[r] = perm[(int64_t(T - 2) * 2 + 1) * W + k];
        cu[r] = lnu[int64_t(T - 2) * W + k];
      }
    }
  }
  __syncthreads();
  // pair j couples temperature j+1 (hot) with j (cold)
  for (int j = T - 2; j >= 0; --j) {
    int32_t na[kR], nb[kR];
    double nu[kR];
    // prefetch: draws of pair j-1, logL row of temperature j-1 (the cold row of pair j-1)
    if (j >= 1) {
#pragma unroll
      for (int r = 0; r < kR; ++r) {
        const int k = tid + r * nt;
        if (k < W) {
          na[r] = perm[(int64_t(j - 1) * 2 + 0) * W + k];
          nb[r] = perm[(int64_t(j - 1) * 2 + 1) * W + k];
          nu[r] = lnu[int64_t(j - 1) * W + k];
          nl[r] = logl[int64_t(j - 1) * W + k];
        }
      }
    }
    const double dbeta = __dsub_rn(A.betas[j], A.betas[j + 1]);
    int local = 0;
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      const int k = tid + r * nt;
      if (k < W) {
        const int a = ca[r], b = cb[r];  // slot a of temp j+1 <-> slot b of temp j; perm rows are permutations
        const double la = lh[a], lb = lc[b];
        if (__dmul_rn(dbeta, __dsub_rn(la, lb)) > cu[r]) {
          const int32_t sa = sh[a];
          sh[a] = sc[b];
          sc[b] = sa;
          lh[a] = lb;
          lc[b] = la;
          ++local;
        }
      }
    }
    local = __reduce_add_sync(0xffffffffu, local);
    if (lane == 0 && local) atomicAdd(&s_count[j & 1], local);
    if (kBuf == 3 && j >= 1) {  // stage the next cold row into the spare buffer (nobody reads it in this phase)
#pragma unroll
      for (int r = 0; r < kR; ++r) {
        const int k = tid + r * nt;
        if (k < W) {
          ls[k] = nl[r];
          ss[k] = (j - 1) * W + k;
        }
      }
    }
    __syncthreads();
    // row j+1 of the plan is final: write it back (each thread the elements it re-initialises later)
    for (int w = tid; w < W; w += nt) A.src[int64_t(j + 1) * W + w] = sh[w];
    if (tid == 0) {
      const int32_t c = s_count[j & 1];
      A.n_acc[j] = c;
      s_nacc[j] = c;
      s_count[j & 1] = 0;
    }
    if (kBuf == 3) {
      double* tl = lh; lh = lc; lc = ls; ls = tl;
      int32_t* ts = sh; sh = sc; sc = ss; ss = ts;
    } else {
      double* tl = lh; lh = lc; lc = tl;
      int32_t* ts = sh; sh = sc; sc = ts;
      if (j >= 1) {
#pragma unroll
        for (int r = 0; r < kR; ++r) {
          const int k = tid + r * nt;
          if (k < W) {
            lc[k] = nl[r];
            sc[k] = (j - 1) * W + k;
          }
        }
        __syncthreads();
      }
    }
#pragma unroll
    for (int r = 0; r < kR; ++r) { ca[r] = na[r]; cb[r] = nb[r]; cu[r] = nu[r]; }
  }
  if (A.src)
    for (int w = tid; w < W; w += nt) A.src[w] = sh[w];
  plan_tail(A, s_x, s_nacc);
}

// The same plan when the pairs of every sweep are listed BY THEIR SLOT IN THE WARMER ROW (perm[j, 0, :] = identity;
// draws.py relabels the reference's (iperm, i1perm, u) triples that way — the set of (a, b, u) triples, hence every
// decision, is unchanged).  Thread k then owns slot k of the warmer row for the whole pair: that row lives in
// REGISTERS, and only the colder row is accessed at random.  The plan kernel is bound by shared-memory bank
// conflicts of exactly those random accesses (8 per element in the generic kernel: 2.2 us per pair at W = 2048);
// here there are 4, the rest is conflict-free: 24 W bytes of shared memory, one barrier per pair.
template <int kR>
__global__ void __launch_bounds__(1024) pt_swap_plan_sorted_kernel(const PtPlan A) {
  extern __shared__ __align__(128) unsigned char plan_smem[];
  const int32_t T = A.T, W = A.W;
  double* lc = reinterpret_cast<double*>(plan_smem);  // colder row of the current pair
  double* ls = lc + W;                                 // staging: colder row of the next pair
  int32_t* sc = reinterpret_cast<int32_t*>(ls + W);
  int32_t* ss = sc + W;
  __shared__ int32_t s_count[2];
  __shared__ int32_t s_nacc[kPlanMaxT];
  __shared__ double s_x[kPlanMaxT];
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
  const double* __restrict__ logl = A.logl;
  const int32_t* __restrict__ perm = A.perm;
  const double* __restrict__ lnu = A.lnu;

  double hl[kR], cu[kR], nl[kR];
  int32_t hs[kR], cb[kR];
#pragma unroll
  for (int r = 0; r < kR; ++r) {
    const int k = tid + r * nt;
    if (k < W) {
      hl[r] = logl[int64_t(T - 1) * W + k];
      hs[r] = (T - 1) * W + k;
      if (T > 1) {
        lc[k] = logl[int64_t(T - 2) * W + k];
        sc[k] = (T - 2) * W + k;
        cb[r] = perm[(int64_t(T - 2) * 2 + 1) * W + k];
        cu[r] = lnu[int64_t(T - 2) * W + k];
      }
    }
  }
  if (tid < 2) s_count[tid] = 0;
  __syncthreads();
  for (int j = T - 2; j >= 0; --j) {  // pair j: temperature j+1 (warmer, in registers) with j (colder, shared)
    int32_t nb[kR];
    double nu[kR];
    if (j >= 1) {
#pragma unroll
      for (int r = 0; r < kR; ++r) {
        const int k = tid + r * nt;
        if (k < W) {
          nb[r] = perm[(int64_t(j - 1) * 2 + 1) * W + k];
          nu[r] = lnu[int64_t(j - 1) * W + k];
          nl[r] = logl[int64_t(j - 1) * W + k];
        }
      }
    }
    const double dbeta = __dsub_rn(A.betas[j], A.betas[j + 1]);
    int local = 0;
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      const int k = tid + r * nt;
      if (k < W) {
        const int b = cb[r];  // slot k of temp j+1 <-> slot b of temp j; b runs over a permutation
        const double lb = lc[b];
        if (__dmul_rn(dbeta, __dsub_rn(hl[r], lb)) > cu[r]) {
          const int32_t sb = sc[b];
          sc[b] = hs[r];
          lc[b] = hl[r];
          hl[r] = lb;
          hs[r] = sb;
          ++local;
        }
      }
    }
    local = __reduce_add_sync(0xffffffffu, local);
    if (lane == 0 && local) atomicAdd(&s_count[j & 1], local);
    if (j >= 1) {  // stage the next colder row (nobody reads the staging buffer in this phase)
#pragma unroll
      for (int r = 0; r < kR; ++r) {
        const int k = tid + r * nt;
        if (k < W) {
          ls[k] = nl[r];
          ss[k] = (j - 1) * W + k;
        }
      }
    }
    __syncthreads();
    // row j+1 of the plan is final (registers); row j, as the swaps left it, becomes the warmer row
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      const int k = tid + r * nt;
      if (k < W) {
        A.src[int64_t(j + 1) * W + k] = hs[r];
        hl[r] = lc[k];
        hs[r] = sc[k];
        cb[r] = nb[r];
        cu[r] = nu[r];
      }
    }
    if (tid == 0) {
      const int32_t c = s_count[j & 1];
      A.n_acc[j] = c;
      s_nacc[j] = c;
      s_count[j & 1] = 0;
    }
    double* tl = lc; lc = ls; ls = tl;
    int32_t* ts = sc; sc = ss; ss = ts;
  }
  if (A.src) {
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      const int k = tid + r * nt;
      if (k < W) A.src[k] = hs[r];
    }
  }
  plan_tail(A, s_x, s_nacc);
}

// The sorted plan fed by the TMA engine (W <= 2048).  Measured (profiles/r02_plan_bench.log): the register-prefetch
// kernels take 0.9 us per pair at W = 512 and +0.9 us per further 1024 walkers.  They are bound by INSTRUCTION
// ISSUE of the one SM they run on: ~108 SASS instructions per walker and pair (64-bit address arithmetic of the
// global prefetches, range predicates, staging copies) x 32 warps / 4 issue slots.  Here every pair's three rows are
// contiguous in global memory, so ONE thread fetches them with three 1-D bulk copies (cp.async.bulk -> mbarrier
// complete_tx) into a ring of n_stages stages, n_stages - 1 pairs ahead of their use, and the 1024 threads only
// touch shared memory with 32-bit indices: ~30 instructions per walker and pair.  Arrays are padded to kR * 1024
// entries whose uniform is +inf (a padded slot never swaps), so the pair loop carries no range predicate.
// A stage is [logL row (becomes the colder row, modified in place) | uniforms | partner indices | source indices]
// = 24 * kR * 1024 bytes; W % 4 == 0 (16-byte rows for the bulk copies).
template <int kR>
__global__ void __launch_bounds__(1024) pt_swap_plan_tma_kernel(const PtPlan A, int n_stages) {
  extern __shared__ __align__(128) unsigned char plan_smem[];
  __shared__ __align__(8) uint64_t s_bar[8];
  __shared__ int32_t s_count[2];
  __shared__ int32_t s_nacc[kPlanMaxT];
  __shared__ double s_x[kPlanMaxT];
  constexpr int kPad = kR * 1024;
  constexpr uint32_t kStage = 24u * kPad;
  const int32_t T = A.T, W = A.W;
  const int tid = threadIdx.x, lane = tid & 31;
  auto st_l = [&](int s) { return reinterpret_cast<double*>(plan_smem + s * kStage); };
  auto st_u = [&](int s) { return reinterpret_cast<double*>(plan_smem + s * kStage + 8u * kPad); };
  auto st_b = [&](int s) { return reinterpret_cast<int32_t*>(plan_smem + s * kStage + 16u * kPad); };
  auto st_s = [&](int s) { return reinterpret_cast<int32_t*>(plan_smem + s * kStage + 20u * kPad); };
  auto issue = [&](int jj) {  // thread 0: the three rows of pair jj -> stage jj % n_stages
    const int s = jj % n_stages;
    fence_proxy_async();
    mbar_arrive_expect_tx(&s_bar[s], uint32_t(W) * 20u);
    tma_bulk_g2s(st_l(s), A.logl + int64_t(jj) * W, uint32_t(W) * 8u, &s_bar[s]);
    tma_bulk_g2s(st_u(s), A.lnu + int64_t(jj) * W, uint32_t(W) * 8u, &s_bar[s]);
    tma_bulk_g2s(st_b(s), A.perm + (int64_t(jj) * 2 + 1) * W, uint32_t(W) * 4u, &s_bar[s]);
  };
  if (tid == 0) {
    for (int s = 0; s < n_stages; ++s) mbar_init(&s_bar[s], 1);
    fence_barrier_init();
  }
  if (tid < 2) s_count[tid] = 0;
  // padding of every stage (the bulk copies only ever write the first W entries)
  for (int s = 0; s < n_stages; ++s) {
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      const int k = tid + r * 1024;
      if (k >= W) {
        st_l(s)[k] = 0.0;
        st_u(s)[k] = INFINITY;
        st_b(s)[k] = k;
      }
    }
  }
  __syncthreads();
  // prologue: the first n_stages - 1 pairs are in flight before anything is decided.  Stage (T-1) % n_stages stays
  // empty: it is the one the first iteration refills (the stage "of pair T-1", which does not exist)
  for (int q = 0; q < n_stages - 1; ++q) {
    const int jj = T - 2 - q;
    if (jj < 0) break;
    if (tid == 0) issue(jj);
    int32_t* sc = st_s(jj % n_stages);
#pragma unroll
    for (int r = 0; r < kR; ++r) sc[tid + r * 1024] = jj * W + tid + r * 1024;
  }
  // the warmest row lives in registers
  double hl[kR];
  int32_t hs[kR];
  bool valid[kR];
#pragma unroll
  for (int r = 0; r < kR; ++r) {
    const int k = tid + r * 1024;
    valid[r] = k < W;
    hl[r] = valid[r] ? A.logl[int64_t(T - 1) * W + k] : 0.0;
    hs[r] = (T - 1) * W + k;
  }
  int32_t* src_row = A.src + int64_t(T - 1) * W + tid;  // row j+1 of the plan, this thread's first slot
  for (int i = tid; i < T; i += 1024) s_x[i] = A.betas[i];  // the ladder, read once (s_x is free until plan_tail)
  __syncthreads();
  // ring bookkeeping without divisions: stage of pair j, stage of pair j+1, parity of the stage's current use
  int s = (T - 2) % n_stages, s_prev = (T - 1) % n_stages, uses = 0;
  uint32_t parity = 0;
  for (int j = T - 2; j >= 0; --j) {  // pair j: temperature j+1 (warmer, registers) with j (colder, its stage)
    mbar_wait(&s_bar[s], parity);
    double* lc = st_l(s);
    const double* us = st_u(s);
    const int32_t* bs = st_b(s);
    int32_t* sc = st_s(s);
    const double dbeta = __dsub_rn(s_x[j], s_x[j + 1]);
    int local = 0;
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      const int k = tid + r * 1024;
      const int b = bs[k];  // slot k of temp j+1 <-> slot b of temp j; b runs over a permutation
      const double lb = lc[b];
      if (__dmul_rn(dbeta, __dsub_rn(hl[r], lb)) > us[k]) {
        const int32_t sb = sc[b];
        sc[b] = hs[r];
        lc[b] = hl[r];
        hl[r] = lb;
        hs[r] = sb;
        ++local;
      }
    }
    local = __reduce_add_sync(0xffffffffu, local);
    if (lane == 0 && local) atomicAdd(&s_count[j & 1], local);
    __syncthreads();
    // row j+1 of the plan is final (registers); row j, as the swaps left it, becomes the warmer row
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      const int k = tid + r * 1024;
      if (valid[r]) src_row[r * 1024] = hs[r];
      hl[r] = lc[k];
      hs[r] = sc[k];
    }
    src_row -= W;
    if (tid == 0) {
      const int32_t c = s_count[j & 1];
      A.n_acc[j] = c;
      s_nacc[j] = c;
      s_count[j & 1] = 0;
    }
    // the stage of pair j+1 was read back before this barrier by every thread: refill it, n_stages - 1 pairs ahead
    const int jj = j + 1 - n_stages;
    if (jj >= 0) {
      if (tid == 0) {
        fence_proxy_async();
        mbar_arrive_expect_tx(&s_bar[s_prev], uint32_t(W) * 20u);
        tma_bulk_g2s(st_l(s_prev), A.logl + int64_t(jj) * W, uint32_t(W) * 8u, &s_bar[s_prev]);
        tma_bulk_g2s(st_u(s_prev), A.lnu + int64_t(jj) * W, uint32_t(W) * 8u, &s_bar[s_prev]);
        tma_bulk_g2s(st_b(s_prev), A.perm + (int64_t(jj) * 2 + 1) * W, uint32_t(W) * 4u, &s_bar[s_prev]);
      }
      int32_t* sn = st_s(s_prev);
#pragma unroll
      for (int r = 0; r < kR; ++r) sn[tid + r * 1024] = jj * W + tid + r * 1024;
    }
    s_prev = s;
    s = (s == 0) ? n_stages - 1 : s - 1;
    if (++uses == n_stages) { uses = 0; parity ^= 1u; }
  }
  if (A.src) {
#pragma unroll
    for (int r = 0; r < kR; ++r)
      if (valid[r]) A.src[tid + r * 1024] = hs[r];
  }
  plan_tail(A, s_x, s_nacc);
}

// ---- swap plan by chains (default when the pairs are listed by their slot in the warmer row) ------------------
// The hot -> cold sweep looks sequential over the T-1 pairs AND coupled over the walkers, but it is W INDEPENDENT
// CHAINS.  Pair j couples slot s of the warmer row j+1 with slot b = perm_j[s] of the colder row j, and row j is
// untouched until pair j is decided; so the element that pair j-1 finds in slot b of row j is either the walker the
// chain carried down to (j+1, s) or the original walker of (j, b) — it depends on ONE decision, which depends on the
// one before it along the path s_{T-1} = k, s_j = perm_j[s_{j+1}].  The path does not depend on the decisions at
// all.  One thread per chain follows its path from the hottest row down with the carried walker (source index,
// logL) in registers: no shared-memory exchange and no barrier between pairs, chains spread over many SMs.  Per
// pair a thread issues three independent L2 gathers at the new slot (the colder row's logL, the next partner, the
// next uniform) and decides the PREVIOUS pair while they are in flight: the critical path is one L2 latency per
// pair (~0.13 us; the single-CTA shared-memory kernels above need 1.24 us per pair at W = 2048, all of it
// bank-conflict replays and one block barrier).  Decisions are the same expressions on the same operands as in
// those kernels: the plan is bit-identical.  Swap counts: warp ballot -> shared-memory counters per CTA -> global
// integer atomics (order-independent); the last CTA to finish (ticket) publishes n_acc, resets the scratch for the
// next launch and runs the ladder adaptation + histories (plan_tail).
constexpr int kChainThreads = 64;
__global__ void __launch_bounds__(kChainThreads) pt_swap_plan_chain_kernel(const PtPlan A, int32_t* __restrict__ g_cnt,
                                                                            uint32_t* __restrict__ g_ticket) {
  __shared__ int32_t s_nacc[kPlanMaxT];
  __shared__ double s_x[kPlanMaxT];
  __shared__ int s_last;
  const int32_t T = A.T, W = A.W;
  const int tid = threadIdx.x, lane = tid & 31;
  const int k = blockIdx.x * kChainThreads + tid;
  const bool valid = k < W;
  for (int i = tid; i < T; i += kChainThreads) { s_x[i] = A.betas[i]; s_nacc[i] = 0; }
  // operands: the caller's arrays, or (sharded ladder, peer-push exchange) the gathered block of this sweep's parity
  const double* __restrict__ g_logl = A.logl;
  const double* __restrict__ g_lnu = A.lnu;
  const int32_t* __restrict__ g_par = A.perm + W;  // partner row of pair j at g_par + j * par_stride
  int64_t par_stride = 2 * int64_t(W);
  if (A.gath[0]) {
    const long long sweep = *A.sweep_counter;      // read by every CTA before the last one increments it
    const unsigned char* g = A.gath[sweep & 1];
    g_logl = reinterpret_cast<const double*>(g);
    g_lnu = reinterpret_cast<const double*>(g + gath_lnu_off(T, W));
    g_par = reinterpret_cast<const int32_t*>(g + gath_partner_off(T, W));
    par_stride = W;
    if (tid < A.n_ranks) {  // every rank's rows of this sweep have landed (bounded spin: a dead peer must not hang us)
      const long long* f = reinterpret_cast<const long long*>(g + gath_flags_off(T, W)) + tid;
      const long long t0 = clock64();
      while (ld_acquire_sys(f) < sweep + 1)
        if (clock64() - t0 > 60000000000ll) __trap();
    }
  }
  __syncthreads();
  if (T >= 2) {
    // carried walker of the chain that starts in slot k of the hottest row
    int s = valid ? k : 0;                                     // slot in row j+1
    double hl = __ldcg(g_logl + int64_t(T - 1) * W + s);
    int32_t hs = (T - 1) * W + s;
    // pair T-2: partner slot and uniform
    int b = __ldcg(g_par + int64_t(T - 2) * par_stride + s);
    double u = __ldcg(g_lnu + int64_t(T - 2) * W + s);
    // The decision of pair j is taken one iteration late, under the gathers of pair j-1: its operands arrived
    // together with the partner slot those gathers needed.
    auto decide = [&](int j, int s_w, int s_c, double lb, double uu) {
      // pair j: temperature j+1 (carried walker in slot s_w) with temperature j (slot s_c, logL lb)
      const double dbeta = __dsub_rn(s_x[j], s_x[j + 1]);
      const bool acc = valid && (__dmul_rn(dbeta, __dsub_rn(hl, lb)) > uu);
      const int32_t cold = j * W + s_c;
      if (valid) A.src[int64_t(j + 1) * W + s_w] = acc ? cold : hs;  // row j+1 of the plan is final in slot s_w
      if (!acc) { hl = lb; hs = cold; }                               // the chain goes on with whoever sits in (j, s_c)
      const unsigned m = __ballot_sync(0xffffffffu, acc);
      if (lane == 0 && m) atomicAdd(&s_nacc[j], __popc(m));
    };
    int p_s = 0, p_b = 0;
    double p_lb = 0.0, p_u = 0.0;
    for (int j = T - 2; j >= 0; --j) {
      // gathers at slot b of row j: its logL (pair j), the partner and the uniform of pair j-1
      const double lb_new = __ldcg(g_logl + int64_t(j) * W + b);
      int b_next = 0;
      double u_next = 0.0;
      if (j > 0) {
        b_next = __ldcg(g_par + int64_t(j - 1) * par_stride + b);
        u_next = __ldcg(g_lnu + int64_t(j - 1) * W + b);
      }
      if (j < T - 2) decide(j + 1, p_s, p_b, p_lb, p_u);
      p_s = s; p_b = b; p_lb = lb_new; p_u = u;
      s = b; b = b_next; u = u_next;
    }
    decide(0, p_s, p_b, p_lb, p_u);
    if (valid) A.src[s] = hs;  // row 0
  } else if (valid && A.src) {
    A.src[k] = k;
  }
  __syncthreads();
  // per-CTA counts -> global; the last CTA owns the totals
  for (int i = tid; i < T - 1; i += kChainThreads)
    if (s_nacc[i]) atomicAdd(&g_cnt[i], s_nacc[i]);
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(g_ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int i = tid; i < T - 1; i += kChainThreads) {
    const int32_t c = atomicExch(&g_cnt[i], 0);   // read the total and leave the scratch zeroed for the next launch
    s_nacc[i] = c;
    if (A.n_acc) A.n_acc[i] = c;
  }
  if (tid == 0) *g_ticket = 0;
  for (int i = tid; i < T; i += kChainThreads) s_x[i] = 0.0;
  plan_tail(A, s_x, s_nacc);
}

// Fallback for ensembles whose rows do not fit in shared memory (W > 8192): rows in global scratch.
__global__ void __launch_bounds__(1024) pt_swap_plan_global_kernel(const PtPlan A, double* __restrict__ ll_work) {
  __shared__ int32_t s_count;
  __shared__ int32_t s_nacc[kPlanMaxT];
  __shared__ double s_x[kPlanMaxT];
  const int32_t T = A.T, W = A.W;
  double* ll_hot = ll_work;
  double* ll_cold = ll_work + W;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int64_t g = tid; g < int64_t(T) * W; g += nt) A.src[g] = int32_t(g);
  for (int w = tid; w < W; w += nt) ll_hot[w] = A.logl[int64_t(T - 1) * W + w];
  __syncthreads();
  for (int i = T - 1; i >= 1; --i) {
    if (tid == 0) s_count = 0;
    for (int w = tid; w < W; w += nt) ll_cold[w] = A.logl[int64_t(i - 1) * W + w];
    __syncthreads();
    const double dbeta = __dsub_rn(A.betas[i - 1], A.betas[i]);
    const int32_t* pi = A.perm + (int64_t(i - 1) * 2 + 0) * W;
    const int32_t* pi1 = A.perm + (int64_t(i - 1) * 2 + 1) * W;
    const double* u = A.lnu + int64_t(i - 1) * W;
    int local = 0;
    for (int k = tid; k < W; k += nt) {
      const int a = pi[k], b = pi1[k];
      const double la = ll_hot[a], lb = ll_cold[b];
      if (__dmul_rn(dbeta, __dsub_rn(la, lb)) > u[k]) {
        const int64_t ga = int64_t(i) * W + a, gb = int64_t(i - 1) * W + b;
        const int32_t sa = A.src[ga];
        A.src[ga] = A.src[gb];
        A.src[gb] = sa;
        ll_hot[a] = lb;
        ll_cold[b] = la;
        ++local;
      }
    }
    if (local) atomicAdd(&s_count, local);
    __syncthreads();
    if (tid == 0) { A.n_acc[i - 1] = s_count; s_nacc[i - 1] = s_count; }
    double* tmp = ll_hot; ll_hot = ll_cold; ll_cold = tmp;
    __syncthreads();
  }
  plan_tail(A, s_x, s_nacc);
}

// ---- plan application: row gather + swap mean distance + chain store ------------------------------------
struct PtApply {
  int32_t T_loc, W, ndim;
  int32_t T_all, G, rank, strided;   // ladder size, ranks, this rank, layout (1: rank r holds r, r+G, ...)
  const int32_t* src;                // [T_all, W] plan (global flat indices t*W + w); NULL = identity (no swap sweep)
  // CURRENT buffers of every rank ([T_loc, W, ndim] / [T_loc, W]); entry `rank` is the local one, the others are
  // peer HBM mapped through CUDA IPC and read over NVLink
  const double* p_in[kMaxPeers];
  const double* ll_in[kMaxPeers];
  const double* lp_in[kMaxPeers];
  double* p_out; double* ll_out; double* lp_out;  // local target buffers (NULL when src == NULL)
  // swap mean distance (consumers emp.py:961-965, 1985-1990): per destination temperature, mean over the slots
  // that received a walker from a hotter rung of |x_new - x_old| in units of the prior widths D_
  const double* D;                   // [ndim] or NULL
  double* smd_part;                  // [T_loc, blocks_per_temp, 2] partial (sum, count)
  uint32_t* smd_ticket;              // [T_loc]
  double* smd_hist;                  // [hist_cap, T_loc]
  const long long* sweep_counter;    // already incremented by the plan kernel: row = *counter - 1
  long long hist_cap;
  // chain store: slot = (*step_counter - 1) / thin when (*step_counter - 1) % thin == 0
  double* chain; double* ch_ll; double* ch_lp;  // [cap, T_loc, W, ndim] / [cap, T_loc, W]; NULL = no store
  const long long* step_counter;
  long long store_cap, store_ring;   // capacity in samples; ring != 0: slot wraps (host streaming sink)
  int32_t thin;
};

constexpr int kApplyWarps = 8;
constexpr int kApplyRowsPerWarp = 4;
constexpr int kApplyRows = kApplyWarps * kApplyRowsPerWarp;  // rows of one temperature per CTA

__device__ __forceinline__ long long store_slot(const long long* step_counter, int thin, long long cap, long long ring) {
  if (!step_counter) return -1;
  const long long n = *step_counter - 1;
  if (n < 0 || (n % thin) != 0) return -1;
  long long s = n / thin;
  if (ring) s %= cap;
  return s < cap ? s : -1;
}

// grid = (blocks_per_temp, T_loc); a CTA owns kApplyRows consecutive walkers of one local temperature
__global__ void __launch_bounds__(kApplyWarps * 32) pt_apply_plan_kernel(const PtApply A) {
  __shared__ double s_sum[kApplyWarps];
  __shared__ int s_cnt[kApplyWarps];
  __shared__ int s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tl = blockIdx.y;
  const int tg = A.strided ? tl * A.G + A.rank : A.rank * A.T_loc + tl;
  const int ndim = A.ndim, W = A.W;
  const long long slot = A.chain ? store_slot(A.step_counter, A.thin, A.store_cap, A.store_ring) : -1;
  double wsum = 0.0;
  int wcnt = 0;
  for (int rr = 0; rr < kApplyRowsPerWarp; ++rr) {
    const int w = (blockIdx.x * kApplyWarps + warp) * kApplyRowsPerWarp + rr;
    if (w >= W) break;
    const int64_t r = int64_t(tl) * W + w;
    int owner = A.rank;
    int64_t srow = r;
    int st = tg;
    if (A.src) {
      const int32_t s = A.src[int64_t(tg) * W + w];
      st = s / W;
      const int sw = s - st * W;
      owner = A.strided ? st % A.G : st / A.T_loc;
      const int sl = A.strided ? st / A.G : st % A.T_loc;
      srow = int64_t(sl) * W + sw;
    }
    const double* __restrict__ pin = A.p_in[owner];
    double dist2 = 0.0;
    const bool down = A.D && st > tg;
    for (int dd = lane; dd < ndim; dd += 32) {
      const double v = pin[srow * ndim + dd];
      if (A.p_out) A.p_out[r * ndim + dd] = v;
      if (slot >= 0) A.chain[(slot * A.T_loc * W + r) * ndim + dd] = v;
      if (down) {
        const double z = __ddiv_rn(__dsub_rn(v, A.p_in[A.rank][r * ndim + dd]), A.D[dd]);
        dist2 = __dadd_rn(dist2, __dmul_rn(z, z));
      }
    }
    if (lane == 0) {
      const double l = A.ll_in[owner][srow], q = A.lp_in[owner][srow];
      if (A.ll_out) { A.ll_out[r] = l; A.lp_out[r] = q; }
      if (slot >= 0) { A.ch_ll[slot * A.T_loc * W + r] = l; A.ch_lp[slot * A.T_loc * W + r] = q; }
    }
    if (down) {  // fixed-order butterfly: every lane ends with the same sum
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) dist2 = __dadd_rn(dist2, __shfl_xor_sync(0xffffffffu, dist2, off));
      wsum = __dadd_rn(wsum, __dsqrt_rn(dist2));
      ++wcnt;
    }
  }
  if (!A.D || !A.smd_hist) return;
  // deterministic reduction: warps in order inside the CTA, CTAs in order by the last one to finish
  if (lane == 0) { s_sum[warp] = wsum; s_cnt[warp] = wcnt; }
  __syncthreads();
  const int nb = gridDim.x;
  if (threadIdx.x == 0) {
    double s = 0.0;
    int c = 0;
    for (int k = 0; k < kApplyWarps; ++k) { s = __dadd_rn(s, s_sum[k]); c += s_cnt[k]; }
    A.smd_part[(int64_t(tl) * nb + blockIdx.x) * 2 + 0] = s;
    A.smd_part[(int64_t(tl) * nb + blockIdx.x) * 2 + 1] = double(c);
    __threadfence();
    const unsigned t = atomicAdd(&A.smd_ticket[tl], 1u);
    s_last = (t == unsigned(nb - 1));
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    double s = 0.0, c = 0.0;
    const volatile double* part = A.smd_part;
    for (int b = 0; b < nb; ++b) {
      s = __dadd_rn(s, part[(int64_t(tl) * nb + b) * 2 + 0]);
      c += part[(int64_t(tl) * nb + b) * 2 + 1];
    }
    const long long row = A.sweep_counter ? *A.sweep_counter - 1 : 0;
    if (row >= 0 && row < A.hist_cap) A.smd_hist[row * A.T_loc + tl] = __ddiv_rn(s, c > 1.0 ? c : 1.0);
    A.smd_ticket[tl] = 0;  // ready for the next sweep
  }
}

// Legacy row gather (emp_pt_gather_rows): rows of the *_in arrays picked by `src`.
__global__ void pt_gather_rows_kernel(int64_t n_rows, int32_t ndim, const int32_t* __restrict__ src,
                                      const double* __restrict__ p_in, const double* __restrict__ ll_in,
                                      const double* __restrict__ lp_in, double* __restrict__ p_out,
                                      double* __restrict__ ll_out, double* __restrict__ lp_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  const int64_t n_warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp_global; r < n_rows; r += n_warps) {
    const int64_t s = src[r];
    for (int dd = lane; dd < ndim; dd += 32) p_out[r * ndim + dd] = p_in[s * ndim + dd];
    if (lane == 0) { ll_out[r] = ll_in[s]; lp_out[r] = lp_in[s]; }
  }
}

}  // namespace emp
