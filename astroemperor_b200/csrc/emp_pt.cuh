// emp_pt.cuh — parallel-tempering step kernels.
//
// The sampler arithmetic of the reference lives in reddemcee / emcee 3.1.6 (not vendored);
// semantics restated in oracle/pt_oracle.py.  All random numbers are host-supplied so that
// the accept / swap decisions are a pure function of (state, draws):
//   stretch move  : q = c[rint] - (c[rint] - s) * zz          (emcee StretchMove.get_proposal)
//   accept        : factors + (beta*ll' + lp') - (beta*ll + lp) > ln u   (emcee RedBlueMove.propose,
//                   tempered as in ptemcee/reddemcee)
//   swap (i,i-1)  : (beta_{i-1} - beta_i) * (ll_i[perm_i] - ll_{i-1}[perm_{i-1}]) > ln u, hot -> cold
// Every floating-point expression below uses explicit single-rounding intrinsics (no FMA
// contraction) so that, given identical log-likelihoods, the decisions are bit-identical to
// the NumPy oracle.
#pragma once
#include <stdint.h>

namespace emp {

// Proposal for one half of every temperature: q[t, j, :] for j in [0, H)
//   half_idx [T, 2, H]: walkers of split 0 / split 1 in ascending order
//   zz, rint [T, 2, H]: draws of the walkers of split s, in that order
__global__ void pt_propose_kernel(const double* __restrict__ p, int32_t T, int32_t W, int32_t ndim,
                                  int32_t split, const int32_t* __restrict__ half_idx,
                                  const double* __restrict__ zz, const int32_t* __restrict__ rint,
                                  double* __restrict__ q) {
  const int32_t H = W / 2;
  const int64_t total = int64_t(T) * H * ndim;
  for (int64_t g = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; g < total;
       g += int64_t(gridDim.x) * blockDim.x) {
    const int32_t dd = int32_t(g % ndim);
    const int64_t tj = g / ndim;
    const int32_t j = int32_t(tj % H), t = int32_t(tj / H);
    const int64_t o = (int64_t(t) * 2 + split) * H + j;
    const int32_t i = half_idx[o];
    const int32_t partner = half_idx[(int64_t(t) * 2 + (1 - split)) * H + rint[o]];
    const double s = p[(int64_t(t) * W + i) * ndim + dd];
    const double c = p[(int64_t(t) * W + partner) * ndim + dd];
    // c[rint] - (c[rint] - s) * zz[:, None]
    q[g] = __dsub_rn(c, __dmul_rn(__dsub_rn(c, s), zz[o]));
  }
}

// Metropolis accept for one half; updates p / logl / logp in place.
__global__ void pt_accept_kernel(double* __restrict__ p, double* __restrict__ logl, double* __restrict__ logp,
                                 int32_t T, int32_t W, int32_t ndim, int32_t split,
                                 const int32_t* __restrict__ half_idx, const double* __restrict__ betas,
                                 const double* __restrict__ factors, const double* __restrict__ lnu,
                                 const double* __restrict__ q, const double* __restrict__ llq,
                                 const double* __restrict__ lpq, uint8_t* __restrict__ accepted,
                                 uint32_t* __restrict__ n_nan, unsigned long long* __restrict__ cnt) {
  const int32_t H = W / 2;
  const int64_t total = int64_t(T) * H;
  // one warp per proposal so the coordinate copy is coalesced
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  const int64_t n_warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t tj = warp_global; tj < total; tj += n_warps) {
    const int32_t j = int32_t(tj % H), t = int32_t(tj / H);
    const int64_t o = (int64_t(t) * 2 + split) * H + j;
    const int32_t i = half_idx[o];
    const int64_t w = int64_t(t) * W + i;
    const double beta = betas[t];
    const double ll_new = llq[tj], lp_new = lpq[tj];
    const double ll_old = logl[w], lp_old = logp[w];
    // lnpdiff = f + (beta*ll' + lp') - (beta*ll + lp)
    const double post_new = __dadd_rn(__dmul_rn(beta, ll_new), lp_new);
    const double post_old = __dadd_rn(__dmul_rn(beta, ll_old), lp_old);
    const double lnpdiff = __dsub_rn(__dadd_rn(factors[o], post_new), post_old);
    const bool acc = lnpdiff > lnu[o];
    if (lane == 0) {
      accepted[w] = acc ? 1 : 0;
      if (ll_new != ll_new && lp_new != -INFINITY) atomicAdd(n_nan, 1u);
      atomicAdd(&cnt[0], 1ull);
      if (lp_new != -INFINITY) atomicAdd(&cnt[1], 1ull);
      if (acc) atomicAdd(&cnt[2], 1ull);
    }
    if (acc) {
      for (int dd = lane; dd < ndim; dd += 32) p[w * ndim + dd] = q[tj * ndim + dd];
      if (lane == 0) { logl[w] = ll_new; logp[w] = lp_new; }
    }
  }
}

// Swap plan: sequential over temperature pairs (hot -> cold), parallel over walkers.
// Single CTA; cur[] holds, for the hotter temperature of the current pair, the flat source
// index and log-likelihood of what presently sits in each slot.
//   src [T, W]: flat index of the ORIGINAL slot whose walker ends in (t, w)
__global__ void pt_swap_plan_kernel(int32_t T, int32_t W, const double* __restrict__ logl,
                                    const double* __restrict__ betas, const int32_t* __restrict__ perm,
                                    const double* __restrict__ lnu, int32_t* __restrict__ src,
                                    int32_t* __restrict__ n_acc, double* __restrict__ ll_work) {
  // ll_work [2, W] scratch in global memory (W may exceed what fits comfortably in smem)
  __shared__ int32_t s_count;
  double* ll_hot = ll_work;       // current content of temperature i
  double* ll_cold = ll_work + W;  // current content of temperature i-1
  const int tid = threadIdx.x, nt = blockDim.x;
  // identity plan
  for (int64_t g = tid; g < int64_t(T) * W; g += nt) src[g] = int32_t(g);
  for (int w = tid; w < W; w += nt) ll_hot[w] = logl[int64_t(T - 1) * W + w];
  __syncthreads();
  for (int i = T - 1; i >= 1; --i) {
    if (tid == 0) s_count = 0;
    for (int w = tid; w < W; w += nt) ll_cold[w] = logl[int64_t(i - 1) * W + w];
    __syncthreads();
    const double dbeta = __dsub_rn(betas[i - 1], betas[i]);
    const int32_t* pi = perm + (int64_t(i - 1) * 2 + 0) * W;
    const int32_t* pi1 = perm + (int64_t(i - 1) * 2 + 1) * W;
    const double* u = lnu + int64_t(i - 1) * W;
    int local = 0;
    for (int k = tid; k < W; k += nt) {
      const int a = pi[k], b = pi1[k];  // slot a of temp i  <->  slot b of temp i-1
      const double la = ll_hot[a], lb = ll_cold[b];
      const double paccept = __dmul_rn(dbeta, __dsub_rn(la, lb));
      if (paccept > u[k]) {
        // perm rows are permutations: no two k touch the same slot
        const int64_t ga = int64_t(i) * W + a, gb = int64_t(i - 1) * W + b;
        const int32_t sa = src[ga];
        src[ga] = src[gb];
        src[gb] = sa;
        ll_hot[a] = lb;
        ll_cold[b] = la;
        ++local;
      }
    }
    if (local) atomicAdd(&s_count, local);
    __syncthreads();
    if (tid == 0) n_acc[i - 1] = s_count;
    // temperature i-1 becomes the hot side of the next pair
    double* tmp = ll_hot; ll_hot = ll_cold; ll_cold = tmp;
    __syncthreads();
  }
}

// Same plan with the two active logL rows and their source-index rows resident in shared memory
// (24*W bytes, W <= 8192).  The sweep is sequential over pairs and every rank of a sharded ladder replays
// all of them, so this kernel's latency is the serial (Amdahl) part of a sweep.  Per pair: the colder
// logL row, the two permutation rows and the uniforms of the NEXT pair are prefetched into registers while
// the current pair is decided out of shared memory (software pipeline: the global-load latency, ~1.5 us of
// the former 2.9 us per pair, is hidden), two block barriers, one coalesced write-back of the finished row
// of `src` by the threads that own the same elements in the next pair (no barrier needed for it).
// kR = elements per thread = ceil(W / 1024).
template <int kR>
__global__ void __launch_bounds__(1024) pt_swap_plan_smem_kernel(int32_t T, int32_t W, const double* __restrict__ logl,
                                                                const double* __restrict__ betas,
                                                                const int32_t* __restrict__ perm,
                                                                const double* __restrict__ lnu,
                                                                int32_t* __restrict__ src, int32_t* __restrict__ n_acc) {
  extern __shared__ __align__(16) unsigned char plan_smem[];
  double* ll_hot = reinterpret_cast<double*>(plan_smem);
  double* ll_cold = ll_hot + W;
  int32_t* sr_hot = reinterpret_cast<int32_t*>(ll_cold + W);
  int32_t* sr_cold = sr_hot + W;
  __shared__ int32_t s_count;
  const int tid = threadIdx.x, nt = blockDim.x;
  int32_t pa[kR], pb[kR];
  double pu[kR], pl[kR];
  // prefetch of pair j (temperatures j+1 and j): perm rows, uniforms, the colder logL row
  auto prefetch = [&](int j) {
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      const int k = tid + r * nt;
      if (k < W) {
        pa[r] = perm[(int64_t(j) * 2 + 0) * W + k];
        pb[r] = perm[(int64_t(j) * 2 + 1) * W + k];
        pu[r] = lnu[int64_t(j) * W + k];
        pl[r] = logl[int64_t(j) * W + k];
      }
    }
  };
  for (int w = tid; w < W; w += nt) {
    ll_hot[w] = logl[int64_t(T - 1) * W + w];
    sr_hot[w] = (T - 1) * W + w;
  }
  if (tid == 0) s_count = 0;
  if (T > 1) prefetch(T - 2);
  for (int i = T - 1; i >= 1; --i) {
    int32_t ca[kR], cb[kR];
    double cu[kR];
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      const int k = tid + r * nt;
      ca[r] = pa[r]; cb[r] = pb[r]; cu[r] = pu[r];
      if (k < W) {
        ll_cold[k] = pl[r];
        sr_cold[k] = (i - 1) * W + k;
      }
    }
    const double dbeta = __dsub_rn(betas[i - 1], betas[i]);
    __syncthreads();
    if (i >= 2) prefetch(i - 2);  // in flight while this pair is decided
    int local = 0;
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      const int k = tid + r * nt;
      if (k < W) {
        const int a = ca[r], b = cb[r];  // slot a of temp i  <->  slot b of temp i-1; perm rows are permutations
        const double la = ll_hot[a], lb = ll_cold[b];
        if (__dmul_rn(dbeta, __dsub_rn(la, lb)) > cu[r]) {
          const int32_t sa = sr_hot[a];
          sr_hot[a] = sr_cold[b];
          sr_cold[b] = sa;
          ll_hot[a] = lb;
          ll_cold[b] = la;
          ++local;
        }
      }
    }
    if (local) atomicAdd(&s_count, local);
    __syncthreads();
    // row i of the plan is final: write it back (each thread the elements it overwrites next iteration);
    // row i-1 becomes the hot side
    for (int w = tid; w < W; w += nt) src[int64_t(i) * W + w] = sr_hot[w];
    if (tid == 0) { n_acc[i - 1] = s_count; s_count = 0; }
    double* tl = ll_hot; ll_hot = ll_cold; ll_cold = tl;
    int32_t* ts = sr_hot; sr_hot = sr_cold; sr_cold = ts;
  }
  __syncthreads();
  for (int w = tid; w < W; w += nt) src[w] = sr_hot[w];
}

// Apply a plan to rows [t0, t0 + T_loc) held locally; sources must be local too
// (single-GPU, or after the cross-rank exchange staged remote rows into `p_in`).
//   src_local: flat index into the *_in arrays (already translated by the caller)
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
