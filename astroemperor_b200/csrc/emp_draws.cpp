// emp_draws.cpp — host-side supply of the random draws of a parallel-tempering sweep.
//
// BASELINE.json north_star: "the stretch-move proposal, the Metropolis accept and the temperature swap are done
// on device from host-supplied random draws".  The draws are what the reference stack consumes from its
// numpy.random.RandomState (emcee 3.1.6 RedBlueMove.propose + StretchMove.get_proposal, then the swap sweep of
// the ptemcee lineage; draw ORDER documented in astroemperor_b200/draws.py).  Generating them with NumPy costs
// ~100 Python-level calls per sweep (0.3 ms for 10 temperatures, 5 ms for 32 x 2048 walkers), which made the
// end-to-end rate of the small BASELINE configs host-bound.  This file restates the legacy RandomState
// algorithms those calls run — MT19937, random_sample (53-bit doubles), shuffle / permutation (Fisher-Yates over
// random_interval with masked rejection), randint (masked rejection on 32-bit words) — so that every stream
// yields BIT-IDENTICAL draws to `numpy.random.RandomState` (tests/test_host_logic.py checks it against NumPy
// itself), one stream per temperature and per adjacent swap pair, streams processed in parallel threads.
// Transformations that involve libm (log) stay in NumPy so that thresholds are the same bits on both paths.
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../include/emperor_b200.h"

namespace {

struct Mt19937 {
  uint32_t key[624];
  int pos;

  void gen() {
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MAT = 0x9908b0dfu;
    int kk;
    uint32_t y;
    for (kk = 0; kk < 624 - 397; ++kk) {
      y = (key[kk] & UPPER) | (key[kk + 1] & LOWER);
      key[kk] = key[kk + 397] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MAT);
    }
    for (; kk < 623; ++kk) {
      y = (key[kk] & UPPER) | (key[kk + 1] & LOWER);
      key[kk] = key[kk + (397 - 624)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MAT);
    }
    y = (key[623] & UPPER) | (key[0] & LOWER);
    key[623] = key[396] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MAT);
    pos = 0;
  }
  inline uint32_t next32() {
    if (pos == 624) gen();
    uint32_t y = key[pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  // numpy legacy random_sample / rand / uniform(0, 1): 53-bit double from two words
  inline double next_double() {
    const int32_t a = next32() >> 5, b = next32() >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
  }
  // numpy random_interval(max): uniform integer in [0, max], masked rejection (32-bit words for max < 2^32)
  inline uint32_t interval(uint32_t max) {
    if (max == 0) return 0;
    uint32_t mask = max;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t v;
    while ((v = (next32() & mask)) > max) {}
    return v;
  }
  // RandomState.shuffle / permutation on a 1-D array: for i = n-1 .. 1: j = interval(i); swap(x[i], x[j])
  template <typename T>
  void shuffle(T* x, int n) {
    for (int i = n - 1; i >= 1; --i) {
      const uint32_t j = interval(uint32_t(i));
      const T t = x[i]; x[i] = x[j]; x[j] = t;
    }
  }
};

}  // namespace

// A small persistent worker pool: the draws of a sweep are T + (T-1) independent streams of ~20-100 us each, and a
// sweep of a small ensemble lasts 0.2 ms on the device, so spawning threads per call would cost more than it buys.
class WorkerPool {
 public:
  explicit WorkerPool(int n_workers) {
    for (int i = 0; i < n_workers; ++i) th_.emplace_back([this]() { loop(); });
  }
  ~WorkerPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
      ++gen_;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  // runs fn(0..n-1) on the workers and the calling thread; returns when all are done.
  // Every run owns its counters (a Job on the caller's stack) and workers pick the job pointer up UNDER the mutex,
  // together with their "active" mark: a worker that wakes late — after the run it was woken for has finished, maybe
  // while the next one is being set up — either finds no job, or a job whose owner cannot return before the worker
  // has left it.  (Shared counters + an unsynchronised function pointer let such a worker claim item 0 of the NEXT
  // run with the stale null pointer of the previous one and drop it: the caller then waited for ever.)
  void run(int n, const std::function<void(int)>& fn) {
    if (th_.empty() || n <= 1) {
      for (int i = 0; i < n; ++i) fn(i);
      return;
    }
    Job job;
    job.fn = &fn;
    job.n = n;
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = &job;
      ++gen_;
    }
    cv_.notify_all();
    work(job);
    while (job.done.load(std::memory_order_acquire) < n) std::this_thread::yield();
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = nullptr;
    }
    while (active_.load(std::memory_order_acquire) > 0) std::this_thread::yield();
  }

 private:
  struct Job {
    const std::function<void(int)>* fn = nullptr;
    int n = 0;
    std::atomic<int> next{0}, done{0};
  };
  static void work(Job& j) {
    for (;;) {
      const int i = j.next.fetch_add(1);
      if (i >= j.n) break;
      (*j.fn)(i);
      j.done.fetch_add(1, std::memory_order_release);
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      Job* j = nullptr;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&]() { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        j = job_;
        if (j) active_.fetch_add(1);
      }
      if (j) {
        work(*j);
        active_.fetch_sub(1, std::memory_order_release);
      }
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_;
  Job* job_ = nullptr;            // guarded by m_
  std::atomic<int> active_{0};
  uint64_t gen_ = 0;
  bool stop_ = false;
};

struct EmpDrawStreams {
  std::vector<Mt19937> s;
  std::unique_ptr<WorkerPool> pool;
};

static int dfail(int code) { return code; }

extern "C" int emp_draws_create(int32_t n_streams, const uint32_t* keys, const int32_t* pos, int32_t n_threads,
                                EmpDrawStreams** out) {
  if (!out || !keys || !pos || n_streams < 1) return dfail(EMP_EINVAL);
  EmpDrawStreams* d = new (std::nothrow) EmpDrawStreams();
  if (!d) return dfail(EMP_ENOMEM);
  d->s.resize(n_streams);
  for (int i = 0; i < n_streams; ++i) {
    memcpy(d->s[i].key, keys + size_t(i) * 624, 624 * sizeof(uint32_t));
    if (pos[i] < 0 || pos[i] > 624) { delete d; return dfail(EMP_EINVAL); }
    d->s[i].pos = pos[i];
  }
  n_threads = std::max(1, std::min(n_threads, 64));
  d->pool.reset(new WorkerPool(n_threads - 1));  // the calling thread works too
  *out = d;
  return EMP_OK;
}

extern "C" int emp_draws_destroy(EmpDrawStreams* d) {
  delete d;
  return EMP_OK;
}

extern "C" int emp_draws_get_state(EmpDrawStreams* d, int32_t stream, uint32_t* key624, int32_t* pos) {
  if (!d || !key624 || !pos || stream < 0 || stream >= int(d->s.size())) return dfail(EMP_EINVAL);
  memcpy(key624, d->s[stream].key, 624 * sizeof(uint32_t));
  *pos = d->s[stream].pos;
  return EMP_OK;
}

// Stretch draws of one temperature stream for nsteps RedBlue steps, emcee's order per step: shuffle(arange(W) % 2);
// then per split: rand(H) [-> zz], randint(H, size=H), rand(H) [-> ln u].  Outputs are [nsteps, n_temps, 2, H]; j is
// the temperature's position in that layout.
static void stretch_stream(Mt19937& g, int j, int n_temps, int W, int nsteps, int32_t* half_idx, double* u_zz,
                           int32_t* rint, double* u_acc) {
  const int H = W / 2;
  std::vector<uint8_t> inds(W);
  for (int s = 0; s < nsteps; ++s) {
    const size_t base = (size_t(s) * n_temps + j) * 2 * H;
    for (int i = 0; i < W; ++i) inds[i] = uint8_t(i & 1);
    g.shuffle(inds.data(), W);
    int n0 = 0, n1 = 0;
    for (int i = 0; i < W; ++i) {
      if (inds[i] == 0) half_idx[base + n0++] = i;
      else half_idx[base + H + n1++] = i;
    }
    for (int split = 0; split < 2; ++split) {
      const size_t o = base + size_t(split) * H;
      for (int i = 0; i < H; ++i) u_zz[o + i] = g.next_double();
      for (int i = 0; i < H; ++i) rint[o + i] = int32_t(g.interval(uint32_t(H - 1)));  // randint(H): [0, H-1]
      for (int i = 0; i < H; ++i) u_acc[o + i] = g.next_double();
    }
  }
}

// Swap draws of one adjacent pair, consumed hot -> cold: permutation(W), permutation(W), uniform(size=W) — then
// RELABELLED: the reference pairs slot iperm[k] of the warmer row with slot i1perm[k] of the colder one under the
// uniform u[k]; the same set of (a, b, u) triples is handed over listed by a (row 0 = identity, row 1 = b(a),
// u = u(a)), which lets the plan kernel keep the warmer row in registers (emp_pt.cuh).  No decision changes.
static void swap_stream(Mt19937* g, int W, int32_t* p0, double* uk) {
  if (!g) {  // padding row of a sharded ladder
    memset(p0, 0, 2 * size_t(W) * sizeof(int32_t));
    for (int i = 0; i < W; ++i) uk[i] = 1.0;  // ln 1 = 0
    return;
  }
  std::vector<int32_t> a(W), b(W);
  std::vector<double> u(W);
  for (int i = 0; i < W; ++i) a[i] = i;
  g->shuffle(a.data(), W);
  for (int i = 0; i < W; ++i) b[i] = i;
  g->shuffle(b.data(), W);
  for (int i = 0; i < W; ++i) u[i] = g->next_double();
  int32_t* p1 = p0 + W;
  for (int k = 0; k < W; ++k) {
    p0[a[k]] = a[k];
    p1[a[k]] = b[k];
    uk[a[k]] = u[k];
  }
}

// All draws of one sweep in one parallel region: the stretch draws of the temperatures temp_streams[n_temps]
// ([nsteps, n_temps, 2, H]: half_idx, u_zz = uniform behind the stretch factor, rint, u_acc = accept uniform) and the
// swap draws of the pairs pair_streams[n_rows] (stream index < 0: padding row; perm [n_rows, 2, W], u_swap [n_rows, W]).
// The uniforms are returned raw: NumPy applies ((a-1)u+1)^2/a and the logs.
extern "C" int emp_draws_sweep(EmpDrawStreams* d, const int32_t* temp_streams, int32_t n_temps, int32_t W,
                               int32_t nsteps, int32_t* half_idx, double* u_zz, int32_t* rint, double* u_acc,
                               const int32_t* pair_streams, int32_t n_rows, int32_t* perm, double* u_swap) {
  if (!d || n_temps < 0 || n_rows < 0 || W < 2 || (W & 1) || nsteps < 0) return dfail(EMP_EINVAL);
  if (n_temps > 0 && (!temp_streams || !half_idx || !u_zz || !rint || !u_acc)) return dfail(EMP_EINVAL);
  if (n_rows > 0 && (!pair_streams || !perm || !u_swap)) return dfail(EMP_EINVAL);
  const int ns = int(d->s.size());
  for (int j = 0; j < n_temps; ++j)
    if (temp_streams[j] < 0 || temp_streams[j] >= ns) return dfail(EMP_EINVAL);
  for (int k = 0; k < n_rows; ++k)
    if (pair_streams[k] >= ns) return dfail(EMP_EINVAL);
  const std::function<void(int)> item = [=](int i) {
    if (i < n_temps) {
      stretch_stream(d->s[temp_streams[i]], i, n_temps, W, nsteps, half_idx, u_zz, rint, u_acc);
    } else {
      const int k = i - n_temps;
      swap_stream(pair_streams[k] < 0 ? nullptr : &d->s[pair_streams[k]], W, perm + size_t(k) * 2 * W,
                  u_swap + size_t(k) * W);
    }
  };
  d->pool->run(n_temps + n_rows, item);
  return EMP_OK;
}

// k consecutive sweeps in ONE parallel region: sweep q's arrays start q * stride_bytes behind the given pointers
// (the chunk layout of `emp_pt_sweep_chunk`).  A stream's k sweeps are drawn by one worker in order, so the draws are
// exactly those of k successive emp_draws_sweep calls — but small ladders (2 temperatures at BASELINE config 1) pay
// one pool dispatch per chunk instead of one per sweep.
extern "C" int emp_draws_sweeps(EmpDrawStreams* d, int32_t k, int64_t stride_bytes, const int32_t* temp_streams,
                                int32_t n_temps, int32_t W, int32_t nsteps, int32_t* half_idx, double* u_zz,
                                int32_t* rint, double* u_acc, const int32_t* pair_streams, int32_t n_rows,
                                int32_t* perm, double* u_swap) {
  if (!d || k < 1 || stride_bytes < 0 || n_temps < 0 || n_rows < 0 || W < 2 || (W & 1) || nsteps < 0)
    return dfail(EMP_EINVAL);
  if (n_temps > 0 && (!temp_streams || !half_idx || !u_zz || !rint || !u_acc)) return dfail(EMP_EINVAL);
  if (n_rows > 0 && (!pair_streams || !perm || !u_swap)) return dfail(EMP_EINVAL);
  const int ns = int(d->s.size());
  for (int j = 0; j < n_temps; ++j)
    if (temp_streams[j] < 0 || temp_streams[j] >= ns) return dfail(EMP_EINVAL);
  for (int r = 0; r < n_rows; ++r)
    if (pair_streams[r] >= ns) return dfail(EMP_EINVAL);
  auto at = [stride_bytes](auto* p, int q) {
    using T = decltype(p);
    return reinterpret_cast<T>(reinterpret_cast<char*>(p) + size_t(q) * size_t(stride_bytes));
  };
  const std::function<void(int)> item = [=](int i) {
    for (int q = 0; q < k; ++q) {
      if (i < n_temps) {
        stretch_stream(d->s[temp_streams[i]], i, n_temps, W, nsteps, at(half_idx, q), at(u_zz, q), at(rint, q),
                       at(u_acc, q));
      } else {
        const int r = i - n_temps;
        swap_stream(pair_streams[r] < 0 ? nullptr : &d->s[pair_streams[r]], W, at(perm, q) + size_t(r) * 2 * W,
                    at(u_swap, q) + size_t(r) * W);
      }
    }
  };
  d->pool->run(n_temps + n_rows, item);
  return EMP_OK;
}
