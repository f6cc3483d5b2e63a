// emp_abi.cu — C-ABI of the B200-native EMPEROR hot path (include/emperor_b200.h).
// Host side: handle lifetime, data packing/upload, kernel launches.  No CPU fallback.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <new>
#include <string>
#include <vector>

#include "../../include/emperor_b200.h"
#include "emp_device.cuh"
#include "emp_pt.cuh"
#include "emp_logl.cuh"
#include "emp_am.cuh"

using namespace emp;

#ifndef EMP_LOGL_GROUPS
#define EMP_LOGL_GROUPS 2  // 64-point groups a likelihood warp works on at once (A/B: scripts/build_variant.sh)
#endif

static thread_local std::string g_last_error;

static int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return fail(EMP_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));           \
  } while (0)

constexpr size_t kPlanSmemMax = 196 * 1024;  // dynamic shared memory of the swap-plan kernel (+ 24 KB static)
constexpr int kGraphCache = 8;

struct GraphEntry {
  EmpPtSweep key;
  cudaGraphExec_t exec = nullptr;
  int64_t launches = 0;
};
// a chunk of k sweeps (+ the upload of their draws) captured as ONE graph (emp_pt_sweep_chunk)
constexpr int kChunkGraphCache = 4;
struct ChunkGraphEntry {
  std::vector<unsigned char> key;  // the k argument blocks + the copy parameters
  cudaGraphExec_t exec = nullptr;
  int64_t launches = 0;
};

struct EmpHandle {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  EmpModelDesc desc;
  EmpModelDesc* d_desc = nullptr;
  int64_t n = 0;
  int32_t n_tiles = 0;
  char* d_tiles = nullptr;
  double *d_t = nullptr, *d_y = nullptr, *d_e2 = nullptr;
  int32_t* d_ins = nullptr;
  // host copies kept for re-packing the tiles when activity columns are attached
  std::vector<double> h_t, h_y, h_e2;
  std::vector<int32_t> h_ins;
  double* d_sai = nullptr;   // [sai_cols][n] activity columns per point (of the point's own instrument)
  int32_t sai_cols = 0;
  bool sai_attached = false;
  uint32_t tile_bytes = kTileBytes;
  double2* d_grid_sc = nullptr;  // sin/cos grid of the Kepler core (emp_device.cuh kep_rv_grid)
  float4* d_grid_scf = nullptr;
  double t0 = 0.0;
  double t_absmax = 0.0;
  double ll_const = 0.0;
  // scratch for the host-buffer entry points and the PT step
  double *d_theta = nullptr, *d_ll = nullptr, *d_lp = nullptr;
  int64_t cap_eval = 0;
  double *d_q = nullptr, *d_llq = nullptr, *d_lpq = nullptr;
  int64_t cap_q = 0;
  double* d_llwork = nullptr;
  int64_t cap_llwork = 0;
  int32_t* d_index = nullptr;  // compact list of evaluations inside the prior support
  int32_t* d_nact = nullptr;   // its length
  int32_t* d_nact2 = nullptr;  // PT step: one counter per half (the likelihood kernel of one half re-zeroes the other)
  bool plan_attr_set = false;
  bool plan_no_tma = false;          // EMP_PLAN_NO_TMA=1: keep the register-prefetch plan kernels (A/B measurements)
  double* d_smd_part = nullptr;      // swap-mean-distance partial sums of the plan application
  uint32_t* d_smd_ticket = nullptr;
  int32_t* d_plan_cnt = nullptr;     // chain plan kernel: swap counts of all CTAs ([kPlanMaxT] + the ticket behind them)
  bool plan_no_chain = false;        // EMP_PLAN_NO_CHAIN=1: the single-CTA shared-memory plan kernels (A/B measurements)
  int64_t cap_smd = 0;
  double *d_model = nullptr, *d_err2 = nullptr;  // emp_model_host scratch
  GraphEntry graphs[kGraphCache];    // captured sweeps (emp_pt_sweep, use_graph)
  int graph_next = 0;
  int64_t graph_captures = 0;
  ChunkGraphEntry chunk_graphs[kChunkGraphCache];
  int chunk_next = 0;
  cudaStream_t cap_stream = nullptr;
  int64_t cap_index = 0;
  uint32_t* d_nan = nullptr;          // [0] NaN proposals
  unsigned long long* d_cnt = nullptr; // [0] proposals, [1] proposals inside the prior support, [2] accepted
  AmDevice am;
  int num_sms = 148;
  int64_t launches = 0;
  bool timing = false;
  int solver = EMP_SOLVER_GRID;
  LoglKernel logl_kernels[kNumFeat] = {};  // one instantiation per model-feature mask (emp_logl.cuh)
  LoglKernel logl_kernel = nullptr;         // the one this handle's descriptor selects
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // per-launch timing of the likelihood kernel (bench.py roofline): ring of event pairs
  std::vector<cudaEvent_t> tev;
  size_t tev_used = 0;
};

static void make_grid_tables(std::vector<double2>& sc, std::vector<float4>& scf);
static int create_fill(EmpHandle* h, const EmpModelDesc* desc, const double* t, const double* y, const double* yerr,
                       const int32_t* flag, int64_t n, const EmpAmData* am, int device, int num_sms);

extern "C" const char* emp_last_error(void) { return g_last_error.c_str(); }
extern "C" int emp_abi_version(void) { return EMP_ABI_VERSION; }

static int validate_desc(const EmpModelDesc* d) {
  if (!d) return fail(EMP_EINVAL, "desc is NULL");
  if (d->abi_version != EMP_ABI_VERSION)
    return fail(EMP_EINVAL, "descriptor abi_version " + std::to_string(d->abi_version) + " != library " +
                                std::to_string(EMP_ABI_VERSION));
  if (d->ndim_full < 1 || d->ndim_full > EMP_MAX_DIM || d->ndim_free < 1 || d->ndim_free > d->ndim_full)
    return fail(EMP_EINVAL, "bad ndim");
  if (d->n_kep < 0 || d->n_kep > EMP_MAX_KEP) return fail(EMP_EINVAL, "bad n_kep");
  if (d->n_ins < 1 || d->n_ins > EMP_MAX_INS) return fail(EMP_EINVAL, "bad n_ins");
  if (d->acc_order < 0 || d->acc_order > EMP_MAX_ACC) return fail(EMP_EINVAL, "bad acc_order");
  if (d->ma_order < 0 || d->ma_order > EMP_MAX_MA) return fail(EMP_EINVAL, "bad ma_order");
  if (d->ma_mode < EMP_MA_NONE || d->ma_mode > EMP_MA_GLOBAL) return fail(EMP_EINVAL, "bad ma_mode");
  if (d->n_prior_ops < 0 || d->n_prior_ops > EMP_MAX_PRIOR_OPS) return fail(EMP_EINVAL, "bad n_prior_ops");
  if (d->n_periodic < 0 || d->n_periodic > EMP_MAX_PERIODIC) return fail(EMP_EINVAL, "bad n_periodic");
  for (int b = 0; b < d->n_periodic; ++b) {
    const int np = d->periodic_kind[b] == 0 ? 3 : 5;
    if (d->periodic_kind[b] < 0 || d->periodic_kind[b] > 1 || d->periodic_off[b] < 0 ||
        d->periodic_off[b] + np > d->ndim_full)
      return fail(EMP_EINVAL, "bad periodic block");
  }
  for (int k = 0; k < d->n_kep; ++k) {
    int m = d->kep_model[k];
    if (m < 0 || m > 7) return fail(EMP_EINVAL, "bad kep_model");
    int np = (m == EMP_AKEP00) ? 7 : 5;
    if (d->kep_off[k] < 0 || d->kep_off[k] + np > d->ndim_full) return fail(EMP_EINVAL, "bad kep_off");
  }
  if (d->n_sai < 0) return fail(EMP_EINVAL, "bad n_sai");
  if (d->n_sai > 0) {
    int tot = 0;
    for (int i = 0; i < d->n_ins; ++i) {
      if (d->sai_count[i] < 0 || d->sai_count[i] > EMP_MAX_SAI) return fail(EMP_EINVAL, "bad sai_count");
      tot += d->sai_count[i];
    }
    if (tot != d->n_sai || d->sai_off < 0 || d->sai_off + d->n_sai > d->ndim_full)
      return fail(EMP_EINVAL, "bad stellar-activity block");
  }
  if (d->offset_off < 0 || d->offset_off + d->n_ins > d->ndim_full) return fail(EMP_EINVAL, "bad offset_off");
  if (d->has_jitter && (d->jitter_off < 0 || d->jitter_off + d->n_ins > d->ndim_full))
    return fail(EMP_EINVAL, "bad jitter_off");
  for (int j = 0; j < d->ndim_free; ++j)
    if (d->free_to_full[j] < 0 || d->free_to_full[j] >= d->ndim_full) return fail(EMP_EINVAL, "bad free_to_full");
  for (int i = 0; i < d->n_prior_ops; ++i) {
    const EmpPriorOp& o = d->prior_ops[i];
    if (o.op < EMP_POP_PARAM || o.op > EMP_POP_SUMSQ) return fail(EMP_EINVAL, "bad prior op");
    if (o.op != EMP_POP_CHECK) {
      if (o.prior < EMP_PRIOR_UNIFORM || o.prior > EMP_PRIOR_FIXED)
        return fail(EMP_EUNSUPPORTED, "prior kind not implemented on the device path");
      if (o.i0 < 0 || o.i0 >= d->ndim_full || o.i1 < 0 || o.i1 >= d->ndim_full)
        return fail(EMP_EINVAL, "bad prior index");
    }
  }
  return EMP_OK;
}

// tiles: [t | y | yerr^2 | ins | activity columns...] per tile; padding replicates the last timestamp
static int pack_tiles(EmpHandle* h, const double* sai_compact, int32_t sai_cols) {
  const int64_t n = h->n;
  const uint32_t tile_bytes = kTileBytes + uint32_t(sai_cols) * kTilePoints * 8;
  std::vector<char> packed(size_t(h->n_tiles) * tile_bytes);
  for (int32_t tix = 0; tix < h->n_tiles; ++tix) {
    char* base = packed.data() + size_t(tix) * tile_bytes;
    double* pt = reinterpret_cast<double*>(base);
    double* py = pt + kTilePoints;
    double* pe = py + kTilePoints;
    int32_t* pi = reinterpret_cast<int32_t*>(pe + kTilePoints);
    double* ps = reinterpret_cast<double*>(base + kTileBytes);
    for (int j = 0; j < kTilePoints; ++j) {
      int64_t i = int64_t(tix) * kTilePoints + j;
      bool ok = i < n;
      pt[j] = ok ? h->h_t[i] : h->h_t[n - 1];
      py[j] = ok ? h->h_y[i] : 0.0;
      pe[j] = ok ? h->h_e2[i] : 1.0;
      pi[j] = ok ? h->h_ins[i] : 0;
      for (int c = 0; c < sai_cols; ++c) ps[size_t(c) * kTilePoints + j] = ok ? sai_compact[size_t(c) * n + i] : 0.0;
    }
  }
  cudaFree(h->d_tiles);
  h->d_tiles = nullptr;
  CUDA_TRY(cudaMalloc(&h->d_tiles, packed.size()));
  CUDA_TRY(cudaMemcpy(h->d_tiles, packed.data(), packed.size(), cudaMemcpyHostToDevice));
  h->tile_bytes = tile_bytes;
  h->sai_cols = sai_cols;
  return EMP_OK;
}

extern "C" int emp_attach_sai(EmpHandle* h, const double* sai_host, int32_t n_sai) {
  if (!h || !sai_host) return fail(EMP_EINVAL, "NULL argument");
  if (n_sai != h->desc.n_sai || n_sai < 1) return fail(EMP_EINVAL, "n_sai does not match the model descriptor");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  int cols = 0, base[EMP_MAX_INS];
  for (int i = 0, acc = 0; i < h->desc.n_ins; ++i) {
    base[i] = acc;
    acc += h->desc.sai_count[i];
    cols = std::max(cols, int(h->desc.sai_count[i]));
  }
  const int64_t n = h->n;
  std::vector<double> compact(size_t(cols) * n, 0.0);
  for (int64_t i = 0; i < n; ++i) {
    const int in = h->h_ins[i];
    for (int c = 0; c < h->desc.sai_count[in]; ++c) compact[size_t(c) * n + i] = sai_host[size_t(base[in] + c) * n + i];
  }
  int rc = pack_tiles(h, compact.data(), cols);
  if (rc) return rc;
  cudaFree(h->d_sai);
  h->d_sai = nullptr;
  CUDA_TRY(cudaMalloc(&h->d_sai, compact.size() * sizeof(double)));
  CUDA_TRY(cudaMemcpy(h->d_sai, compact.data(), compact.size() * sizeof(double), cudaMemcpyHostToDevice));
  h->sai_attached = true;
  return EMP_OK;
}

extern "C" int emp_create(const EmpModelDesc* desc, const double* t, const double* y, const double* yerr,
                          const int32_t* flag, int64_t n, const EmpAmData* am, int device, EmpHandle** out) {
  if (!out) return fail(EMP_EINVAL, "out is NULL");
  *out = nullptr;
  int rc = validate_desc(desc);
  if (rc) return rc;
  if (!t || !y || !yerr || !flag || n < 1) return fail(EMP_EINVAL, "bad data arrays");
  if (desc->am_enabled && !am) return fail(EMP_EINVAL, "am_enabled but am data is NULL");
  for (int64_t i = 0; i < n; ++i)
    if (flag[i] < 1 || flag[i] > desc->n_ins) return fail(EMP_EINVAL, "flag out of 1..n_ins");

  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(EMP_ENODEV, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                " (there is no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(EMP_EINVAL, "device index out of range");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)  // the fatbin holds sm_100a code only: an sm_90 or sm_120 part cannot run it
    return fail(EMP_ENODEV, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                                std::to_string(prop.minor) + "; this library is built for sm_100a only");

  EmpHandle* h = new (std::nothrow) EmpHandle();
  if (!h) return fail(EMP_ENOMEM, "host allocation failed");
  rc = create_fill(h, desc, t, y, yerr, flag, n, am, device, prop.multiProcessorCount);
  if (rc) {  // every failure after this point releases what was allocated so far
    const std::string msg = g_last_error;
    emp_destroy(h);
    return fail(rc, msg);
  }
  *out = h;
  return EMP_OK;
}

static int create_fill(EmpHandle* h, const EmpModelDesc* desc, const double* t, const double* y, const double* yerr,
                       const int32_t* flag, int64_t n, const EmpAmData* am, int device, int num_sms) {
  int rc = EMP_OK;
  h->device = device;
  h->num_sms = num_sms;
  h->desc = *desc;
  h->n = n;
  h->n_tiles = int32_t((n + kTilePoints - 1) / kTilePoints);
  h->t0 = t[0];
  for (int64_t i = 0; i < n; ++i) h->t_absmax = fmax(h->t_absmax, fabs(t[i]));
  h->ll_const = -0.5 * log(2.0 * M_PI) * double(n);  // 00.like:1

  CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  h->owns_stream = true;
  CUDA_TRY(cudaEventCreate(&h->ev0));
  CUDA_TRY(cudaEventCreate(&h->ev1));

  h->h_t.assign(t, t + n);
  h->h_y.assign(y, y + n);
  h->h_e2.resize(n);
  h->h_ins.resize(n);
  for (int64_t i = 0; i < n; ++i) {
    h->h_e2[i] = yerr[i] * yerr[i];  // err20 = YERR_ ** 2 (emp_model.py:714)
    h->h_ins[i] = flag[i] - 1;
  }
  const std::vector<double>& e2 = h->h_e2;
  const std::vector<int32_t>& ins = h->h_ins;
  rc = pack_tiles(h, nullptr, 0);
  if (rc) return rc;
  CUDA_TRY(cudaMalloc(&h->d_t, n * sizeof(double)));
  CUDA_TRY(cudaMalloc(&h->d_y, n * sizeof(double)));
  CUDA_TRY(cudaMalloc(&h->d_e2, n * sizeof(double)));
  CUDA_TRY(cudaMalloc(&h->d_ins, n * sizeof(int32_t)));
  CUDA_TRY(cudaMemcpy(h->d_t, t, n * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(h->d_y, y, n * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(h->d_e2, e2.data(), n * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(h->d_ins, ins.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(&h->d_desc, sizeof(EmpModelDesc)));
  CUDA_TRY(cudaMemcpy(h->d_desc, desc, sizeof(EmpModelDesc), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(&h->d_nan, sizeof(uint32_t)));
  CUDA_TRY(cudaMemset(h->d_nan, 0, sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&h->d_cnt, 4 * sizeof(unsigned long long)));
  CUDA_TRY(cudaMemset(h->d_cnt, 0, 4 * sizeof(unsigned long long)));

  {
    std::vector<double2> sc;
    std::vector<float4> scf;
    make_grid_tables(sc, scf);
    CUDA_TRY(cudaMalloc(&h->d_grid_sc, kGridN * sizeof(double2)));
    CUDA_TRY(cudaMalloc(&h->d_grid_scf, kGridN * sizeof(float4)));
    CUDA_TRY(cudaMemcpy(h->d_grid_sc, sc.data(), kGridN * sizeof(double2), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(h->d_grid_scf, scf.data(), kGridN * sizeof(float4), cudaMemcpyHostToDevice));
  }
  LoglKernelTable<EMP_LOGL_GROUPS, 0>::fill(h->logl_kernels);
  h->logl_kernel = h->logl_kernels[logl_features(*desc)];
  if (!h->logl_kernel) return fail(EMP_EINVAL, "no likelihood kernel for this feature combination");
  CUDA_TRY(cudaFuncSetAttribute((const void*)h->logl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                int(logl_smem_bytes(kTileBytesMax))));
  if (desc->am_enabled) {
    rc = am_upload(am, &h->am);
    if (rc) return rc;
  }
  return EMP_OK;
}

extern "C" int emp_destroy(EmpHandle* h) {
  if (!h) return EMP_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->d_tiles); cudaFree(h->d_t); cudaFree(h->d_y); cudaFree(h->d_e2); cudaFree(h->d_ins);
  cudaFree(h->d_sai); cudaFree(h->d_grid_sc); cudaFree(h->d_grid_scf); cudaFree(h->d_index); cudaFree(h->d_nact); cudaFree(h->d_cnt);
  for (cudaEvent_t e : h->tev) cudaEventDestroy(e);
  cudaFree(h->d_desc); cudaFree(h->d_theta); cudaFree(h->d_ll); cudaFree(h->d_lp);
  cudaFree(h->d_q); cudaFree(h->d_llq); cudaFree(h->d_lpq); cudaFree(h->d_llwork); cudaFree(h->d_nan);
  cudaFree(h->d_nact2); cudaFree(h->d_smd_part); cudaFree(h->d_smd_ticket); cudaFree(h->d_model); cudaFree(h->d_err2);
  cudaFree(h->d_plan_cnt);
  for (GraphEntry& g : h->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  for (ChunkGraphEntry& g : h->chunk_graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  am_free(&h->am);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->owns_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return EMP_OK;
}

extern "C" int emp_stream(EmpHandle* h, void** stream) {
  if (!h || !stream) return fail(EMP_EINVAL, "NULL argument");
  *stream = (void*)h->stream;
  return EMP_OK;
}

extern "C" int emp_set_stream(EmpHandle* h, void* stream) {
  if (!h) return fail(EMP_EINVAL, "NULL handle");
  CUDA_TRY(cudaSetDevice(h->device));
  if (h->stream) CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (h->owns_stream && h->stream) CUDA_TRY(cudaStreamDestroy(h->stream));
  h->stream = (cudaStream_t)stream;
  h->owns_stream = false;
  return EMP_OK;
}

extern "C" int emp_synchronize(EmpHandle* h) {
  if (!h) return fail(EMP_EINVAL, "NULL handle");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return EMP_OK;
}

static LoglParams base_logl_params(EmpHandle* h) {
  LoglParams P;
  P.desc = h->d_desc;
  P.tiles = h->d_tiles;
  P.n_points = h->n;
  P.n_tiles = h->n_tiles;
  P.theta = nullptr;
  P.eval_index = nullptr;
  P.n_active = nullptr;
  P.logl = nullptr;
  P.t0 = h->t0;
  P.ll_const = h->ll_const;
  P.t_absmax = h->t_absmax;
  P.grid_sc = h->d_grid_sc;
  P.grid_scf = h->d_grid_scf;
  P.tile_bytes = h->tile_bytes;
  P.sai_cols = h->sai_cols;
  P.solver = h->solver;
  P.zero_counter = nullptr;
  P.pt = PtAccept();
  P.pt.enabled = 0;
  P.H = make_hot_consts();
  return P;
}

// the likelihood kernel over the compact list (n_eval = upper bound of its length), optionally bracketed by
// timing events
static int launch_logl_kernel(EmpHandle* h, const LoglParams& P, int64_t n_eval, cudaStream_t st) {
  if (n_eval > 2147483647LL - kWalkerWarps) return fail(EMP_EINVAL, "n_eval too large for one launch");
  const unsigned grid = unsigned((n_eval + kWalkerWarps - 1) / kWalkerWarps);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (h->timing) {
    if (h->tev_used + 2 > h->tev.size()) {
      for (int k = 0; k < 2; ++k) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreate(&e));
        h->tev.push_back(e);
      }
    }
    e0 = h->tev[h->tev_used];
    e1 = h->tev[h->tev_used + 1];
    h->tev_used += 2;
    CUDA_TRY(cudaEventRecord(e0, st));
  }
  h->logl_kernel<<<grid, kLoglThreads, logl_smem_bytes(h->tile_bytes), st>>>(P);
  if (h->timing) CUDA_TRY(cudaEventRecord(e1, st));
  h->launches += 1;
  return EMP_OK;
}

static int launch_logl(EmpHandle* h, const double* theta_dev, int64_t n_eval, double* logl_dev,
                       double* logp_dev) {
  if (n_eval == 0) return EMP_OK;
  if (h->desc.n_sai > 0 && !h->sai_attached)
    return fail(EMP_EINVAL, "the model has a StellarActivityBlock: call emp_attach_sai first");
  if (n_eval > 2147483647LL - kWalkerWarps) return fail(EMP_EINVAL, "n_eval too large for one launch");
  if (n_eval > h->cap_index) {
    cudaFree(h->d_index);
    h->d_index = nullptr;
    h->cap_index = 0;
    CUDA_TRY(cudaMalloc(&h->d_index, size_t(n_eval) * sizeof(int32_t)));
    h->cap_index = n_eval;
  }
  if (!h->d_nact) CUDA_TRY(cudaMalloc(&h->d_nact, sizeof(int32_t)));
  CUDA_TRY(cudaMemsetAsync(h->d_nact, 0, sizeof(int32_t), h->stream));
  const unsigned grid_p = unsigned((n_eval + kPriorWarps - 1) / kPriorWarps);
  prior_compact_kernel<<<grid_p, kPriorWarps * 32, 0, h->stream>>>(h->d_desc, theta_dev, n_eval, logl_dev, logp_dev,
                                                                 h->d_index, h->d_nact);
  h->launches += 1;
  LoglParams P = base_logl_params(h);
  P.theta = theta_dev;
  P.eval_index = h->d_index;
  P.n_active = h->d_nact;
  P.logl = logl_dev;
  int rc = launch_logl_kernel(h, P, n_eval, h->stream);
  if (rc) return rc;
  CUDA_TRY(cudaGetLastError());
  if (h->desc.am_enabled) {
    rc = am_launch(&h->am, h->d_desc, theta_dev, n_eval, h->d_index, h->d_nact, logl_dev, h->stream,
                   &h->launches, nullptr);
    if (rc) return rc;
  }
  return EMP_OK;
}

extern "C" int emp_logl_batch(EmpHandle* h, const double* theta_dev, int64_t n_eval, double* logl_dev,
                              double* logp_dev) {
  if (!h || !theta_dev || !logl_dev || !logp_dev || n_eval < 0) return fail(EMP_EINVAL, "bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  return launch_logl(h, theta_dev, n_eval, logl_dev, logp_dev);
}

static int ensure_eval_scratch(EmpHandle* h, int64_t n_eval) {
  if (n_eval <= h->cap_eval) return EMP_OK;
  cudaFree(h->d_theta); cudaFree(h->d_ll); cudaFree(h->d_lp);
  h->d_theta = h->d_ll = h->d_lp = nullptr;
  h->cap_eval = 0;
  CUDA_TRY(cudaMalloc(&h->d_theta, size_t(n_eval) * h->desc.ndim_free * sizeof(double)));
  CUDA_TRY(cudaMalloc(&h->d_ll, size_t(n_eval) * sizeof(double)));
  CUDA_TRY(cudaMalloc(&h->d_lp, size_t(n_eval) * sizeof(double)));
  h->cap_eval = n_eval;
  return EMP_OK;
}

extern "C" int emp_logl_batch_host(EmpHandle* h, const double* theta_host, int64_t n_eval, double* logl_host,
                                   double* logp_host) {
  if (!h || !theta_host || !logl_host || !logp_host || n_eval < 0) return fail(EMP_EINVAL, "bad argument");
  if (n_eval == 0) return EMP_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  int rc = ensure_eval_scratch(h, n_eval);
  if (rc) return rc;
  const size_t nb = size_t(n_eval) * h->desc.ndim_free * sizeof(double);
  CUDA_TRY(cudaMemcpyAsync(h->d_theta, theta_host, nb, cudaMemcpyHostToDevice, h->stream));
  rc = launch_logl(h, h->d_theta, n_eval, h->d_ll, h->d_lp);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(logl_host, h->d_ll, n_eval * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaMemcpyAsync(logp_host, h->d_lp, n_eval * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return EMP_OK;
}

extern "C" int emp_model_host(EmpHandle* h, const double* theta_host, double* model_host, double* err2_host) {
  if (!h || !theta_host || !model_host || !err2_host) return fail(EMP_EINVAL, "bad argument");
  if (h->desc.n_sai > 0 && !h->sai_attached)
    return fail(EMP_EINVAL, "the model has a StellarActivityBlock: call emp_attach_sai first");
  CUDA_TRY(cudaSetDevice(h->device));
  int rc = ensure_eval_scratch(h, 1);
  if (rc) return rc;
  if (!h->d_model) CUDA_TRY(cudaMalloc(&h->d_model, h->n * sizeof(double)));
  if (!h->d_err2) CUDA_TRY(cudaMalloc(&h->d_err2, h->n * sizeof(double)));
  double *d_model = h->d_model, *d_err2 = h->d_err2;
  CUDA_TRY(cudaMemcpyAsync(h->d_theta, theta_host, h->desc.ndim_free * sizeof(double), cudaMemcpyHostToDevice,
                           h->stream));
  int blocks = int((h->n + 255) / 256);
  if (blocks > 4 * h->num_sms) blocks = 4 * h->num_sms;
  model_rv_kernel<<<blocks, 256, 0, h->stream>>>(h->d_desc, h->d_theta, h->d_t, h->d_y, h->d_e2, h->d_ins, h->n,
                                                 h->t0, h->t_absmax, d_model, d_err2, h->d_sai, h->sai_cols,
                                                 make_hot_consts());
  h->launches += 1;
  if (h->desc.ma_mode == EMP_MA_GLOBAL && h->desc.ma_order > 0) {
    model_ma_kernel<<<1, 32, 0, h->stream>>>(h->d_desc, h->d_theta, h->d_t, h->d_y, h->n, d_model);
    h->launches += 1;
    if (h->desc.n_periodic > 0 || h->sai_cols > 0) {
      model_periodic_kernel<<<blocks, 256, 0, h->stream>>>(h->d_desc, h->d_theta, h->d_t, h->d_ins, h->n, h->t_absmax,
                                                           d_model, h->d_sai, h->sai_cols, make_hot_consts());
      h->launches += 1;
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(model_host, d_model, h->n * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(err2_host, d_err2, h->n * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) return fail(EMP_ECUDA, cudaGetErrorString(e));
  return EMP_OK;
}

// ---- parallel tempering ------------------------------------------------------------------

static int ensure_q_scratch(EmpHandle* h, int64_t n_prop) {
  if (!h->d_nact2) {
    CUDA_TRY(cudaMalloc(&h->d_nact2, 2 * sizeof(int32_t)));
    CUDA_TRY(cudaMemset(h->d_nact2, 0, 2 * sizeof(int32_t)));
  }
  if (n_prop > h->cap_index) {
    cudaFree(h->d_index);
    h->d_index = nullptr;
    h->cap_index = 0;
    CUDA_TRY(cudaMalloc(&h->d_index, size_t(n_prop) * sizeof(int32_t)));
    h->cap_index = n_prop;
  }
  if (n_prop <= h->cap_q) return EMP_OK;
  cudaFree(h->d_q); cudaFree(h->d_llq); cudaFree(h->d_lpq);
  h->d_q = h->d_llq = h->d_lpq = nullptr;
  h->cap_q = 0;
  CUDA_TRY(cudaMalloc(&h->d_q, size_t(n_prop) * h->desc.ndim_free * sizeof(double)));
  CUDA_TRY(cudaMalloc(&h->d_llq, size_t(n_prop) * sizeof(double)));
  CUDA_TRY(cudaMalloc(&h->d_lpq, size_t(n_prop) * sizeof(double)));
  h->cap_q = n_prop;
  return EMP_OK;
}

// what one stretch step needs (a view of EmpPtSweep for step `s`, or of emp_pt_stretch_step's arguments)
struct StretchArgs {
  int32_t T, W;
  double *p, *logl, *logp;
  const double* betas;
  int32_t beta_off, beta_stride;
  const int32_t* half_idx;
  const double* zz;
  const int32_t* rint;
  const double* factors;
  const double* lnu;
  uint8_t* accepted;
  int32_t* n_accepted;
  long long* step_counter;
};

// One red/blue stretch step: per half a proposal + prior kernel and the likelihood kernel(s) with the
// Metropolis accept in the epilogue.  Everything it touches was allocated by ensure_q_scratch (capturable).
static int enqueue_stretch_step(EmpHandle* h, const StretchArgs& a, cudaStream_t st) {
  const int32_t H = a.W / 2;
  const int64_t n_prop = int64_t(a.T) * H;
  for (int split = 0; split < 2; ++split) {
    PtPropose pp;
    pp.desc = h->d_desc;
    pp.p = a.p; pp.T = a.T; pp.W = a.W; pp.split = split;
    pp.half_idx = a.half_idx; pp.zz = a.zz; pp.rint = a.rint;
    pp.q = h->d_q; pp.lpq = h->d_lpq;
    pp.eval_index = h->d_index;
    pp.n_active = h->d_nact2 + split;
    pp.accepted = a.accepted;
    pp.cnt = h->d_cnt;
    pp.step_counter = a.step_counter;
    const unsigned grid_p = unsigned((n_prop + kProposeWarps - 1) / kProposeWarps);
    pt_propose_prior_kernel<<<grid_p, kProposeWarps * 32, 0, st>>>(pp);
    h->launches += 1;

    PtAccept pa;
    pa.enabled = 1;
    pa.T = a.T; pa.W = a.W; pa.split = split;
    pa.p = a.p; pa.logl = a.logl; pa.logp = a.logp;
    pa.q = h->d_q; pa.lpq = h->d_lpq;
    pa.half_idx = a.half_idx;
    pa.betas = a.betas; pa.beta_off = a.beta_off; pa.beta_stride = a.beta_stride;
    pa.factors = a.factors; pa.lnu = a.lnu;
    pa.accepted = a.accepted; pa.n_accepted = a.n_accepted;
    pa.cnt = h->d_cnt; pa.n_nan = h->d_nan;

    LoglParams P = base_logl_params(h);
    P.theta = h->d_q;
    P.eval_index = h->d_index;
    P.n_active = h->d_nact2 + split;
    P.zero_counter = h->d_nact2 + (1 - split);
    P.logl = h->d_llq;  // only written when the accept is left to the astrometric kernel
    P.pt = pa;
    if (h->desc.am_enabled) P.pt.enabled = 0;
    int rc = launch_logl_kernel(h, P, n_prop, st);
    if (rc) return rc;
    if (h->desc.am_enabled) {
      rc = am_launch(&h->am, h->d_desc, h->d_q, n_prop, h->d_index, h->d_nact2 + split, h->d_llq, st, &h->launches,
                     &pa);
      if (rc) return rc;
    }
  }
  return EMP_OK;
}

extern "C" int emp_pt_stretch_step(EmpHandle* h, int32_t T, int32_t W, double* p, double* logl, double* logp,
                                   const double* betas, const int32_t* half_idx, const double* zz,
                                   const int32_t* rint, const double* factors, const double* lnu,
                                   uint8_t* accepted) {
  if (!h || !p || !logl || !logp || !betas || !half_idx || !zz || !rint || !factors || !lnu || !accepted)
    return fail(EMP_EINVAL, "NULL argument");
  if (T < 1 || W < 2 || (W & 1)) return fail(EMP_EINVAL, "need T >= 1 and an even number of walkers");
  if (h->desc.n_sai > 0 && !h->sai_attached)
    return fail(EMP_EINVAL, "the model has a StellarActivityBlock: call emp_attach_sai first");
  CUDA_TRY(cudaSetDevice(h->device));
  int rc = ensure_q_scratch(h, int64_t(T) * (W / 2));
  if (rc) return rc;
  StretchArgs a = {T, W, p, logl, logp, betas, 0, 1, half_idx, zz, rint, factors, lnu, accepted, nullptr, nullptr};
  rc = enqueue_stretch_step(h, a, h->stream);
  if (rc) return rc;
  CUDA_TRY(cudaGetLastError());
  return EMP_OK;
}

static int enqueue_plan(EmpHandle* h, const PtPlan& A, cudaStream_t st) {
  const int32_t W = A.W;
  const size_t smem3 = size_t(W) * 36, smem2 = size_t(W) * 24;
  using PlanKernel = void (*)(const PtPlan);
  PlanKernel k = nullptr;
  size_t smem = 0;
  // TMA-fed kernel: arrays padded to kR * 1024 entries, at least 3 ring stages -> W <= 2048
  if (A.hot_sorted && A.T >= 2 && (!h->plan_no_chain || A.gath[0])) {  // gathered operands: chain kernel only
    const unsigned grid = unsigned((W + kChainThreads - 1) / kChainThreads);
    pt_swap_plan_chain_kernel<<<grid, kChainThreads, 0, st>>>(A, h->d_plan_cnt,
                                                              reinterpret_cast<uint32_t*>(h->d_plan_cnt + kPlanMaxT));
    h->launches += 1;
    return EMP_OK;
  }
  const int tma_r = (W + 1023) / 1024;
  const int tma_stages = tma_r <= 2 ? int(std::min<size_t>(kPlanSmemMax / (size_t(tma_r) * 1024 * 24), 8)) : 0;
  if (A.hot_sorted && (W % 4) == 0 && tma_stages >= 3 && A.T >= 2 && !h->plan_no_tma) {
    const size_t bytes = size_t(tma_stages) * tma_r * 1024 * 24;
    if (tma_r == 1) pt_swap_plan_tma_kernel<1><<<1, 1024, bytes, st>>>(A, tma_stages);
    else pt_swap_plan_tma_kernel<2><<<1, 1024, bytes, st>>>(A, tma_stages);
    h->launches += 1;
    return EMP_OK;
  }
  if (A.hot_sorted && smem2 <= kPlanSmemMax && W <= 8192) {
    k = W <= 1024 ? pt_swap_plan_sorted_kernel<1> : W <= 2048 ? pt_swap_plan_sorted_kernel<2>
      : W <= 4096 ? pt_swap_plan_sorted_kernel<4> : pt_swap_plan_sorted_kernel<8>;
    smem = smem2;
  } else if (smem3 <= kPlanSmemMax) {
    k = W <= 1024 ? pt_swap_plan_kernel<1, 3> : W <= 2048 ? pt_swap_plan_kernel<2, 3>
      : W <= 4096 ? pt_swap_plan_kernel<4, 3> : pt_swap_plan_kernel<6, 3>;
    smem = smem3;
  } else if (smem2 <= kPlanSmemMax && W <= 8192) {
    k = pt_swap_plan_kernel<8, 2>;
    smem = smem2;
  }
  if (k) {
    k<<<1, 1024, smem, st>>>(A);
  } else {
    pt_swap_plan_global_kernel<<<1, 1024, 0, st>>>(A, h->d_llwork);
  }
  h->launches += 1;
  return EMP_OK;
}

static int ensure_plan_scratch(EmpHandle* h, int32_t W) {
  if (!h->plan_attr_set) {  // once per handle, not per call
    const char* env = getenv("EMP_PLAN_NO_TMA");
    h->plan_no_tma = env && env[0] == '1';
    const char* env2 = getenv("EMP_PLAN_NO_CHAIN");
    h->plan_no_chain = env2 && env2[0] == '1';
    CUDA_TRY(cudaMalloc(&h->d_plan_cnt, (kPlanMaxT + 2) * sizeof(int32_t)));  // counts | plan ticket | publish ticket
    CUDA_TRY(cudaMemset(h->d_plan_cnt, 0, (kPlanMaxT + 2) * sizeof(int32_t)));
    const void* ks[] = {(const void*)pt_swap_plan_kernel<1, 3>, (const void*)pt_swap_plan_kernel<2, 3>,
                        (const void*)pt_swap_plan_kernel<4, 3>, (const void*)pt_swap_plan_kernel<6, 3>,
                        (const void*)pt_swap_plan_kernel<8, 2>, (const void*)pt_swap_plan_sorted_kernel<1>,
                        (const void*)pt_swap_plan_sorted_kernel<2>, (const void*)pt_swap_plan_sorted_kernel<4>,
                        (const void*)pt_swap_plan_sorted_kernel<8>, (const void*)pt_swap_plan_tma_kernel<1>,
                        (const void*)pt_swap_plan_tma_kernel<2>};
    for (const void* k : ks)
      CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kPlanSmemMax)));
    h->plan_attr_set = true;
  }
  if (size_t(W) * 24 > kPlanSmemMax || W > 8192) {
    if (2 * int64_t(W) > h->cap_llwork) {
      cudaFree(h->d_llwork);
      h->d_llwork = nullptr;
      h->cap_llwork = 0;
      CUDA_TRY(cudaMalloc(&h->d_llwork, 2 * size_t(W) * sizeof(double)));
      h->cap_llwork = 2 * int64_t(W);
    }
  }
  return EMP_OK;
}

extern "C" int emp_pt_swap_plan(EmpHandle* h, int32_t T, int32_t W, const double* logl_all, const double* betas,
                                const int32_t* perm, const double* lnu, int32_t* src, int32_t* n_acc) {
  if (!h || !logl_all || !betas || !src || !n_acc) return fail(EMP_EINVAL, "NULL argument");
  if (T < 1 || W < 1 || T > kPlanMaxT) return fail(EMP_EINVAL, "bad T/W (T <= 2048)");
  if (T > 1 && (!perm || !lnu)) return fail(EMP_EINVAL, "NULL draws");
  CUDA_TRY(cudaSetDevice(h->device));
  int rc = ensure_plan_scratch(h, W);
  if (rc) return rc;
  PtPlan A = {};
  A.T = T; A.W = W; A.logl = logl_all; A.betas = const_cast<double*>(betas); A.perm = perm; A.lnu = lnu;
  A.src = src; A.n_acc = n_acc;
  A.adapt = 0; A.adapt_tau = 1.0; A.adapt_nu = 1.0;
  rc = enqueue_plan(h, A, h->stream);
  if (rc) return rc;
  CUDA_TRY(cudaGetLastError());
  return EMP_OK;
}

extern "C" int emp_pt_gather_rows(EmpHandle* h, int64_t n_rows, int32_t ndim, const int32_t* src,
                                  const double* p_in, const double* logl_in, const double* logp_in, double* p_out,
                                  double* logl_out, double* logp_out) {
  if (!h || !src || !p_in || !logl_in || !logp_in || !p_out || !logl_out || !logp_out)
    return fail(EMP_EINVAL, "NULL argument");
  if (n_rows == 0) return EMP_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  const int threads = 256;
  int blocks = int(std::min<int64_t>((n_rows * 32 + threads - 1) / threads, int64_t(h->num_sms) * 8));
  pt_gather_rows_kernel<<<blocks, threads, 0, h->stream>>>(n_rows, ndim, src, p_in, logl_in, logp_in, p_out,
                                                           logl_out, logp_out);
  h->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return EMP_OK;
}

// ---- whole sweeps ------------------------------------------------------------------------------------------
static int validate_sweep(EmpHandle* h, const EmpPtSweep* s, bool need_swap) {
  if (!h || !s) return fail(EMP_EINVAL, "NULL argument");
  if (s->T_loc < 1 || s->W < 2 || (s->W & 1) || s->nsteps < 0) return fail(EMP_EINVAL, "bad T_loc / W / nsteps");
  if (s->n_ranks < 1 || s->n_ranks > EMP_MAX_PEERS || s->rank < 0 || s->rank >= s->n_ranks)
    return fail(EMP_EINVAL, "bad n_ranks / rank");
  if (s->T_all != s->T_loc * s->n_ranks) return fail(EMP_EINVAL, "T_all != T_loc * n_ranks");
  if (s->T_all > kPlanMaxT) return fail(EMP_EINVAL, "ladder longer than 2048 temperatures");
  if (!s->p || !s->logl || !s->logp || !s->betas || !s->accepted) return fail(EMP_EINVAL, "NULL state");
  if (s->nsteps > 0 && (!s->half_idx || !s->zz || !s->rint || !s->factors || !s->lnu))
    return fail(EMP_EINVAL, "NULL stretch draws");
  if (s->thin < 1) return fail(EMP_EINVAL, "thin must be >= 1");
  if (need_swap && s->T_all > 1) {
    if (!s->perm || !s->lnu_swap || !s->src || !s->n_acc || !s->p_alt || !s->logl_alt || !s->logp_alt)
      return fail(EMP_EINVAL, "NULL swap draws / plan / alternate buffers");
    if (s->n_ranks > 1) {
      const bool push = s->peer_gath[0][s->rank] != nullptr;
      if (push) {
        for (int q = 0; q < 2; ++q)
          for (int r = 0; r < s->n_ranks; ++r)
            if (!s->peer_gath[q][r]) return fail(EMP_EINVAL, "NULL gathered block of a peer");
        if (!s->perm_hot_sorted) return fail(EMP_EINVAL, "the peer-push exchange needs perm_hot_sorted");
      } else if (!s->logl_all) return fail(EMP_EINVAL, "sharded ladder: logl_all is NULL");
      for (int r = 0; r < s->n_ranks; ++r)
        if (!s->peer_p[r] || !s->peer_logl[r] || !s->peer_logp[r]) return fail(EMP_EINVAL, "NULL peer buffer");
    }
  }
  if (h->desc.n_sai > 0 && !h->sai_attached)
    return fail(EMP_EINVAL, "the model has a StellarActivityBlock: call emp_attach_sai first");
  return EMP_OK;
}

static int ensure_apply_scratch(EmpHandle* h, const EmpPtSweep* s) {
  const int nb = (s->W + kApplyRows - 1) / kApplyRows;
  const int64_t need = int64_t(s->T_loc) * nb * 2;
  if (need > h->cap_smd) {
    cudaFree(h->d_smd_part); cudaFree(h->d_smd_ticket);
    h->d_smd_part = nullptr; h->d_smd_ticket = nullptr; h->cap_smd = 0;
    CUDA_TRY(cudaMalloc(&h->d_smd_part, size_t(need) * sizeof(double)));
    CUDA_TRY(cudaMalloc(&h->d_smd_ticket, size_t(s->T_loc) * sizeof(uint32_t)));
    CUDA_TRY(cudaMemset(h->d_smd_ticket, 0, size_t(s->T_loc) * sizeof(uint32_t)));
    h->cap_smd = need;
  }
  return EMP_OK;
}

static void fill_apply(EmpHandle* h, const EmpPtSweep* s, bool with_plan, PtApply& A) {
  A = PtApply();
  A.T_loc = s->T_loc; A.W = s->W; A.ndim = h->desc.ndim_free;
  A.T_all = s->T_all; A.G = s->n_ranks; A.rank = s->rank; A.strided = s->strided;
  A.src = with_plan ? s->src : nullptr;
  for (int r = 0; r < kMaxPeers; ++r) { A.p_in[r] = nullptr; A.ll_in[r] = nullptr; A.lp_in[r] = nullptr; }
  if (s->n_ranks > 1 && with_plan) {
    for (int r = 0; r < s->n_ranks; ++r) { A.p_in[r] = s->peer_p[r]; A.ll_in[r] = s->peer_logl[r]; A.lp_in[r] = s->peer_logp[r]; }
  }
  A.p_in[s->rank] = s->p; A.ll_in[s->rank] = s->logl; A.lp_in[s->rank] = s->logp;
  A.p_out = with_plan ? s->p_alt : nullptr;
  A.ll_out = with_plan ? s->logl_alt : nullptr;
  A.lp_out = with_plan ? s->logp_alt : nullptr;
  const bool smd = with_plan && s->D && s->smd_hist;
  A.D = smd ? s->D : nullptr;
  A.smd_part = h->d_smd_part; A.smd_ticket = h->d_smd_ticket;
  A.smd_hist = smd ? s->smd_hist : nullptr;
  A.sweep_counter = (const long long*)s->sweep_counter;
  A.hist_cap = s->hist_cap;
  A.chain = s->chain; A.ch_ll = s->chain_ll; A.ch_lp = s->chain_lp;
  A.step_counter = (const long long*)s->step_counter;
  A.store_cap = s->store_cap; A.store_ring = s->store_ring; A.thin = s->thin;
}

static int enqueue_apply(EmpHandle* h, const PtApply& A, cudaStream_t st) {
  const dim3 grid((A.W + kApplyRows - 1) / kApplyRows, A.T_loc);
  pt_apply_plan_kernel<<<grid, kApplyWarps * 32, 0, st>>>(A);
  h->launches += 1;
  return EMP_OK;
}

static int enqueue_stretch_phase(EmpHandle* h, const EmpPtSweep* s, bool swap_follows, cudaStream_t st) {
  const int64_t per_step = int64_t(s->T_loc) * 2 * (s->W / 2);
  for (int k = 0; k < s->nsteps; ++k) {
    StretchArgs a;
    a.T = s->T_loc; a.W = s->W; a.p = s->p; a.logl = s->logl; a.logp = s->logp;
    a.betas = s->betas;
    a.beta_off = s->strided ? s->rank : s->rank * s->T_loc;
    a.beta_stride = s->strided ? s->n_ranks : 1;
    a.half_idx = s->half_idx + k * per_step; a.zz = s->zz + k * per_step; a.rint = s->rint + k * per_step;
    a.factors = s->factors + k * per_step; a.lnu = s->lnu + k * per_step;
    a.accepted = s->accepted; a.n_accepted = s->n_accepted;
    a.step_counter = (long long*)s->step_counter;
    int rc = enqueue_stretch_step(h, a, st);
    if (rc) return rc;
    // every stretch step is a stored sample (reddemcee: nsweeps*nsteps samples per run); the last one of a
    // sweep is stored after the swap by the plan application
    const bool last = (k == s->nsteps - 1);
    if (s->chain && (!last || !swap_follows)) {
      PtApply A;
      fill_apply(h, s, false, A);
      rc = enqueue_apply(h, A, st);
      if (rc) return rc;
    }
  }
  return EMP_OK;
}

static int enqueue_swap_phase(EmpHandle* h, const EmpPtSweep* s, cudaStream_t st) {
  PtPlan P = {};
  P.T = s->T_all; P.W = s->W;
  const bool push = s->n_ranks > 1 && s->peer_gath[0][s->rank] != nullptr;
  if (push) {
    // this rank's rows of logL and of the swap draws go straight into every peer's gathered block over NVLink
    PtPublish U = {};
    U.T_loc = s->T_loc; U.W = s->W; U.T_all = s->T_all; U.G = s->n_ranks; U.rank = s->rank; U.strided = s->strided;
    U.logl = s->logl; U.perm = s->perm; U.lnu = s->lnu_swap;
    for (int q = 0; q < 2; ++q)
      for (int r = 0; r < s->n_ranks; ++r) U.peer[q][r] = static_cast<unsigned char*>(s->peer_gath[q][r]);
    U.sweep_counter = (const long long*)s->sweep_counter;
    U.ticket = reinterpret_cast<uint32_t*>(h->d_plan_cnt + kPlanMaxT + 1);
    const dim3 grid((s->W + kPublishThreads - 1) / kPublishThreads, s->T_loc);
    pt_publish_kernel<<<grid, kPublishThreads, 0, st>>>(U);
    h->launches += 1;
    P.gath[0] = static_cast<const unsigned char*>(s->peer_gath[0][s->rank]);
    P.gath[1] = static_cast<const unsigned char*>(s->peer_gath[1][s->rank]);
    P.n_ranks = s->n_ranks;
  }
  P.logl = (s->n_ranks > 1) ? s->logl_all : s->logl;
  P.betas = s->betas; P.perm = s->perm; P.lnu = s->lnu_swap; P.src = s->src; P.n_acc = s->n_acc;
  P.adapt = s->adapt; P.adapt_tau = s->adapt_tau; P.adapt_nu = s->adapt_nu;
  P.hot_sorted = s->perm_hot_sorted;
  P.sweep_counter = (long long*)s->sweep_counter;
  P.beta_hist = s->beta_hist; P.nacc_hist = s->nacc_hist; P.hist_cap = s->hist_cap;
  int rc = enqueue_plan(h, P, st);
  if (rc) return rc;
  PtApply A;
  fill_apply(h, s, true, A);
  return enqueue_apply(h, A, st);
}

static int enqueue_sweep(EmpHandle* h, const EmpPtSweep* s, cudaStream_t st) {
  const bool swap = s->T_all > 1;
  int rc = enqueue_stretch_phase(h, s, swap, st);
  if (rc) return rc;
  if (swap) return enqueue_swap_phase(h, s, st);
  // a single temperature: no swap sweep, but the sweep counter and the beta history still advance
  PtPlan P = {};
  P.T = 1; P.W = s->W; P.logl = s->logl; P.betas = s->betas;
  P.sweep_counter = (long long*)s->sweep_counter;
  P.beta_hist = s->beta_hist; P.hist_cap = s->hist_cap;
  return enqueue_plan(h, P, st);
}

static int prepare_sweep(EmpHandle* h, const EmpPtSweep* s, bool need_swap) {
  int rc = validate_sweep(h, s, need_swap);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(h->device));
  rc = ensure_q_scratch(h, int64_t(s->T_loc) * (s->W / 2));
  if (rc) return rc;
  rc = ensure_plan_scratch(h, s->W);
  if (rc) return rc;
  return ensure_apply_scratch(h, s);
}

extern "C" int emp_pt_sweep(EmpHandle* h, const EmpPtSweep* s) {
  int rc = prepare_sweep(h, s, true);
  if (rc) return rc;
  if (s->n_ranks != 1 && !s->peer_gath[0][s->rank])
    return fail(EMP_EINVAL, "emp_pt_sweep runs a sharded sweep only with the peer-push exchange (peer_gath); "
                            "use emp_pt_sweep_stretch/_swap around the NCCL all-gather otherwise");
  if (s->n_ranks != 1 && !s->sweep_counter) return fail(EMP_EINVAL, "the peer-push exchange needs sweep_counter");
  if (!s->use_graph || h->timing) {
    rc = enqueue_sweep(h, s, h->stream);
    if (rc) return rc;
    CUDA_TRY(cudaGetLastError());
    return EMP_OK;
  }
  // CUDA graph: the sweep is captured once per distinct argument block (the caller alternates between two: the
  // state and the draw staging are double-buffered) on a private stream, then replayed on the handle's stream
  GraphEntry* hit = nullptr;
  for (GraphEntry& g : h->graphs)
    if (g.exec && memcmp(&g.key, s, sizeof(EmpPtSweep)) == 0) { hit = &g; break; }
  if (!hit) {
    if (!h->cap_stream) CUDA_TRY(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    const int64_t l0 = h->launches;
    CUDA_TRY(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
    rc = enqueue_sweep(h, s, h->cap_stream);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(h->cap_stream, &graph);
    const int64_t n_launch = h->launches - l0;
    h->launches = l0;
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return fail(EMP_ECUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(EMP_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    GraphEntry& slot = h->graphs[h->graph_next];
    h->graph_next = (h->graph_next + 1) % kGraphCache;
    if (slot.exec) cudaGraphExecDestroy(slot.exec);
    slot.key = *s;
    slot.exec = exec;
    slot.launches = n_launch;
    hit = &slot;
    h->graph_captures += 1;
  }
  CUDA_TRY(cudaGraphLaunch(hit->exec, h->stream));
  h->launches += hit->launches;
  return EMP_OK;
}

// k whole sweeps in ONE graph launch: small ensembles (BASELINE configs 1-3) finish a sweep in tens of
// microseconds, so a host that enqueues sweep by sweep (draw, stage, upload, launch) is the bottleneck.  The
// caller draws the k sweeps into one pinned block; the graph uploads it with a single copy node and then runs the
// k sweeps back to back (each with its own argument block: the state parity alternates, the draw pointers advance).
extern "C" int emp_pt_sweep_chunk(EmpHandle* h, const EmpPtSweep* s, int32_t k, void* draws_dev,
                                  const void* draws_host, int64_t draws_bytes) {
  if (!h || !s || k < 1) return fail(EMP_EINVAL, "bad chunk");
  if ((draws_host != nullptr) != (draws_dev != nullptr) || (draws_host && draws_bytes < 1))
    return fail(EMP_EINVAL, "draws_host / draws_dev / draws_bytes must come together");
  for (int j = 0; j < k; ++j) {
    int rc = prepare_sweep(h, s + j, true);
    if (rc) return rc;
    if (s[j].n_ranks != 1) return fail(EMP_EINVAL, "emp_pt_sweep_chunk is a single-GPU entry point");
  }
  if (h->timing) {  // per-launch event timing: plain stream launches
    if (draws_host) CUDA_TRY(cudaMemcpyAsync(draws_dev, draws_host, size_t(draws_bytes), cudaMemcpyHostToDevice, h->stream));
    for (int j = 0; j < k; ++j) {
      int rc = enqueue_sweep(h, s + j, h->stream);
      if (rc) return rc;
    }
    CUDA_TRY(cudaGetLastError());
    return EMP_OK;
  }
  std::vector<unsigned char> key(size_t(k) * sizeof(EmpPtSweep) + 3 * sizeof(int64_t));
  memcpy(key.data(), s, size_t(k) * sizeof(EmpPtSweep));
  const int64_t tail[3] = {int64_t(reinterpret_cast<uintptr_t>(draws_dev)),
                           int64_t(reinterpret_cast<uintptr_t>(draws_host)), draws_bytes};
  memcpy(key.data() + size_t(k) * sizeof(EmpPtSweep), tail, sizeof(tail));
  ChunkGraphEntry* hit = nullptr;
  for (ChunkGraphEntry& g : h->chunk_graphs)
    if (g.exec && g.key == key) { hit = &g; break; }
  if (!hit) {
    if (!h->cap_stream) CUDA_TRY(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    const int64_t l0 = h->launches;
    CUDA_TRY(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
    int rc = EMP_OK;
    cudaError_t ec = cudaSuccess;
    if (draws_host)
      ec = cudaMemcpyAsync(draws_dev, draws_host, size_t(draws_bytes), cudaMemcpyHostToDevice, h->cap_stream);
    for (int j = 0; j < k && rc == EMP_OK && ec == cudaSuccess; ++j) rc = enqueue_sweep(h, s + j, h->cap_stream);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(h->cap_stream, &graph);
    const int64_t n_launch = h->launches - l0;
    h->launches = l0;
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ec != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      return fail(EMP_ECUDA, std::string("cudaMemcpyAsync (capture; the draws must be in pinned memory): ") +
                                 cudaGetErrorString(ec));
    }
    if (e != cudaSuccess) return fail(EMP_ECUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(EMP_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    ChunkGraphEntry& slot = h->chunk_graphs[h->chunk_next];
    h->chunk_next = (h->chunk_next + 1) % kChunkGraphCache;
    if (slot.exec) cudaGraphExecDestroy(slot.exec);
    slot.key = std::move(key);
    slot.exec = exec;
    slot.launches = n_launch;
    hit = &slot;
    h->graph_captures += 1;
  }
  CUDA_TRY(cudaGraphLaunch(hit->exec, h->stream));
  h->launches += hit->launches;
  return EMP_OK;
}

extern "C" int emp_pt_sweep_stretch(EmpHandle* h, const EmpPtSweep* s) {
  int rc = prepare_sweep(h, s, false);
  if (rc) return rc;
  rc = enqueue_stretch_phase(h, s, s->T_all > 1, h->stream);
  if (rc) return rc;
  CUDA_TRY(cudaGetLastError());
  return EMP_OK;
}

extern "C" int emp_pt_sweep_swap(EmpHandle* h, const EmpPtSweep* s) {
  int rc = prepare_sweep(h, s, true);
  if (rc) return rc;
  if (s->T_all < 2) return EMP_OK;
  rc = enqueue_swap_phase(h, s, h->stream);
  if (rc) return rc;
  CUDA_TRY(cudaGetLastError());
  return EMP_OK;
}

extern "C" int emp_gather_block_bytes(int32_t T_all, int32_t W, int64_t* bytes) {
  if (!bytes || T_all < 1 || W < 1) return fail(EMP_EINVAL, "bad argument");
  *bytes = int64_t(gath_bytes(T_all, W));
  return EMP_OK;
}

// ---- shareable device memory (CUDA IPC) -------------------------------------------------------------------
extern "C" int emp_dev_alloc(int device, int64_t bytes, void** ptr) {
  if (!ptr || bytes < 1) return fail(EMP_EINVAL, "bad argument");
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaMalloc(ptr, size_t(bytes)));
  return EMP_OK;
}
extern "C" int emp_dev_free(int device, void* ptr) {
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaFree(ptr));
  return EMP_OK;
}
extern "C" int emp_ipc_export(int device, void* ptr, unsigned char handle64[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  if (!ptr || !handle64) return fail(EMP_EINVAL, "NULL argument");
  CUDA_TRY(cudaSetDevice(device));
  cudaIpcMemHandle_t hd;
  CUDA_TRY(cudaIpcGetMemHandle(&hd, ptr));
  memcpy(handle64, &hd, 64);
  return EMP_OK;
}
extern "C" int emp_ipc_open(int device, const unsigned char handle64[64], void** ptr) {
  if (!ptr || !handle64) return fail(EMP_EINVAL, "NULL argument");
  CUDA_TRY(cudaSetDevice(device));
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle64, 64);
  CUDA_TRY(cudaIpcOpenMemHandle(ptr, hd, cudaIpcMemLazyEnablePeerAccess));
  return EMP_OK;
}
extern "C" int emp_ipc_close(int device, void* ptr) {
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return EMP_OK;
}

extern "C" int emp_nan_count(EmpHandle* h, uint32_t* count) {
  if (!h || !count) return fail(EMP_EINVAL, "NULL argument");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemcpyAsync(count, h->d_nan, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return EMP_OK;
}

// ---- introspection ---------------------------------------------------------------------------
extern "C" int emp_launch_count(EmpHandle* h, int64_t* count) {
  if (!h || !count) return fail(EMP_EINVAL, "NULL argument");
  *count = h->launches;
  return EMP_OK;
}

extern "C" int emp_graph_captures(EmpHandle* h, int64_t* count) {
  if (!h || !count) return fail(EMP_EINVAL, "NULL argument");
  *count = h->graph_captures;
  return EMP_OK;
}

extern "C" int emp_set_solver(EmpHandle* h, int solver) {
  if (!h) return fail(EMP_EINVAL, "NULL handle");
  if (solver != EMP_SOLVER_GRID && solver != EMP_SOLVER_KEPLERPY) return fail(EMP_EINVAL, "unknown solver");
  h->solver = solver;
  return EMP_OK;
}

extern "C" int emp_set_timing(EmpHandle* h, int enable) {
  if (!h) return fail(EMP_EINVAL, "NULL handle");
  h->timing = enable != 0;
  h->tev_used = 0;
  return EMP_OK;
}

extern "C" int emp_timing_collect(EmpHandle* h, double* total_ms, int64_t* n_launches) {
  if (!h || !total_ms || !n_launches) return fail(EMP_EINVAL, "NULL argument");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  double tot = 0.0;
  for (size_t i = 0; i + 1 < h->tev_used; i += 2) {
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->tev[i], h->tev[i + 1]));
    tot += ms;
  }
  *total_ms = tot;
  *n_launches = int64_t(h->tev_used / 2);
  h->tev_used = 0;
  return EMP_OK;
}

extern "C" int emp_counters(EmpHandle* h, uint64_t* out4) {
  if (!h || !out4) return fail(EMP_EINVAL, "NULL argument");
  CUDA_TRY(cudaSetDevice(h->device));
  unsigned long long c[4];
  uint32_t nn = 0;
  CUDA_TRY(cudaMemcpyAsync(c, h->d_cnt, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaMemcpyAsync(&nn, h->d_nan, sizeof(nn), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  out4[0] = c[0]; out4[1] = c[1]; out4[2] = c[2]; out4[3] = nn;
  return EMP_OK;
}

// FP64 FMA peak of the device: the roofline denominator of this path (SURVEY.md §8d row D3;
// MEASURED_PEAKS.json carries no FP64 entry).  8 independent DFMA chains per thread.
__global__ void fp64_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
         x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

extern "C" int emp_fp64_peak(int device, double* tflops) {
  if (!tflops) return fail(EMP_EINVAL, "NULL argument");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  const int threads = 512, blocks = prop.multiProcessorCount * 4, iters = 1 << 16;
  double* d_out = nullptr;
  CUDA_TRY(cudaMalloc(&d_out, size_t(blocks) * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CUDA_TRY(cudaEventRecord(e0));
    fp64_peak_kernel<<<blocks, threads>>>(d_out, iters, 0.999999, 1e-7);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * 8.0 * double(iters) * double(blocks) * threads;
    const double tf = fl / (ms * 1e-3) * 1e-12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  *tflops = best;
  return EMP_OK;
}

// ---- kepler.solve drop-in --------------------------------------------
__global__ void kepler_solve_kernel(const double* __restrict__ M, const double* __restrict__ ecc, int64_t n,
                                    int ecc_scalar, double* __restrict__ E, const HotConsts H) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    E[i] = kepler_solve(M[i], ecc[ecc_scalar ? 0 : i], H);
}

extern "C" int emp_kepler_solve_host(const double* M, const double* ecc, int64_t n, int ecc_is_scalar, double* E,
                                     int device) {
  if (!M || !ecc || !E || n < 0) return fail(EMP_EINVAL, "bad argument");
  if (n == 0) return EMP_OK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(EMP_ENODEV, "no CUDA device (there is no CPU fallback)");
  CUDA_TRY(cudaSetDevice(device));
  double *dM = nullptr, *de = nullptr, *dE = nullptr;
  const int64_t ne = ecc_is_scalar ? 1 : n;
  cudaError_t e = cudaMalloc(&dM, n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&de, ne * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dE, n * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(dM, M, n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(de, ecc, ne * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    int blocks = int(std::min<int64_t>((n + 255) / 256, 148 * 8));
    kepler_solve_kernel<<<blocks, 256>>>(dM, de, n, ecc_is_scalar, dE, make_hot_consts());
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(E, dE, n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(dM); cudaFree(de); cudaFree(dE);
  if (e != cudaSuccess) return fail(EMP_ECUDA, cudaGetErrorString(e));
  return EMP_OK;
}

// ---- the likelihood kernel's own solver, exposed for parity tests ------------------------------
__global__ void kepler_grid_kernel(const double* __restrict__ M, const double* __restrict__ ecc, int64_t n,
                                   int ecc_scalar, const double2* __restrict__ tab, const float4* __restrict__ tabf,
                                   double* __restrict__ E, double* __restrict__ sinE, double* __restrict__ cosE,
                                   const HotConsts H) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    KepConst k;
    kep_constants_ecc(ecc[ecc_scalar ? 0 : i], k);
    const double m = M[i];
    double e_out, s_out, c_out;
    if (k.robust || !(fabs(m) < 1.0e12)) {  // the kernel routes these to the kepler.py-style refinement
      e_out = kepler_solve(m, k.e, H);
      sincos(e_out, &s_out, &c_out);
    } else {
      GridStage S;
      kep_grid_a(k, m, H, tab, tabf, S);  // freq = 1, tp = phase = 0: the mean anomaly is m itself
      double sE, cE, dd, y1;
      kep_grid_root(k, S, H, sE, cE, dd, y1);
      const double Er = S.eh + (S.df + dd);
      e_out = S.sign_hi ? H.c[1] - Er : Er;
      s_out = flip_sign(sE, S.sign_hi);
      c_out = cE;
    }
    E[i] = e_out;
    if (sinE) sinE[i] = s_out;
    if (cosE) cosE[i] = c_out;
  }
}

// the same with the per-walker starter table (one eccentricity for the whole array, like a walker's planet):
// every CTA builds the table in shared memory exactly like the likelihood kernel's prologue does
__global__ void kepler_grid_table_kernel(const double* __restrict__ M, double ecc, int64_t n,
                                         const double2* __restrict__ tab, const float4* __restrict__ tabf,
                                         double* __restrict__ E, double* __restrict__ sinE, double* __restrict__ cosE,
                                         const HotConsts H) {
  __shared__ float st[3 * kStartStride];
  __shared__ float scratch[2 * kStartN + 3];
  KepConst k;
  kep_constants_ecc(ecc, k);
  if (threadIdx.x < 32) build_start_table(k, st, scratch, threadIdx.x);
  __syncthreads();
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    GridStage S;
    kep_grid_a<true>(k, M[i], H, tab, tabf, S, st);
    double sE, cE, dd, y1;
    kep_grid_root(k, S, H, sE, cE, dd, y1);
    const double Er = S.eh + (S.df + dd);
    E[i] = S.sign_hi ? H.c[1] - Er : Er;
    if (sinE) sinE[i] = flip_sign(sE, S.sign_hi);
    if (cosE) cosE[i] = cE;
  }
}

static void make_grid_tables(std::vector<double2>& sc, std::vector<float4>& scf) {
  // (sin, cos)(k 2^-7) correctly rounded from long double; the FP32 entries are rounded from those
  sc.resize(kGridN);
  scf.resize(kGridN);
  for (int k = 0; k < kGridN; ++k) {
    const long double x = (long double)k / 128.0L;
    sc[k] = make_double2(double(sinl(x)), double(cosl(x)));
    scf[k] = make_float4(float(sc[k].x), float(sc[k].y), float(0.5 * sc[k].x), float(sc[k].y / 6.0));
  }
}

static int kepler_grid_host_impl(const double* M, const double* ecc, int64_t n, int ecc_is_scalar, double* E,
                                 double* sinE, double* cosE, int device, bool use_table);

extern "C" int emp_kepler_grid_host(const double* M, const double* ecc, int64_t n, int ecc_is_scalar, double* E,
                                    double* sinE, double* cosE, int device) {
  return kepler_grid_host_impl(M, ecc, n, ecc_is_scalar, E, sinE, cosE, device, false);
}

extern "C" int emp_kepler_grid_table_host(const double* M, double ecc, int64_t n, double* E, double* sinE,
                                          double* cosE, int device) {
  if (!(ecc >= 0.0 && ecc <= double(kStartEccMax)) || fabs(ecc) > 1.0)
    return fail(EMP_EINVAL, "the starter table serves eccentricities in [0, 0.8]");
  return kepler_grid_host_impl(M, &ecc, n, 1, E, sinE, cosE, device, true);
}

static int kepler_grid_host_impl(const double* M, const double* ecc, int64_t n, int ecc_is_scalar, double* E,
                                 double* sinE, double* cosE, int device, bool use_table) {
  if (!M || !ecc || !E || n < 0) return fail(EMP_EINVAL, "bad argument");
  if (n == 0) return EMP_OK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(EMP_ENODEV, "no CUDA device (there is no CPU fallback)");
  CUDA_TRY(cudaSetDevice(device));
  std::vector<double2> sc;
  std::vector<float4> scf;
  make_grid_tables(sc, scf);
  double *dM = nullptr, *de = nullptr, *dE = nullptr, *dS = nullptr, *dC = nullptr;
  double2* dtab = nullptr;
  float4* dtabf = nullptr;
  const int64_t ne = ecc_is_scalar ? 1 : n;
  cudaError_t e = cudaMalloc(&dM, n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&de, ne * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dE, n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dS, n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dC, n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dtab, kGridN * sizeof(double2));
  if (e == cudaSuccess) e = cudaMalloc(&dtabf, kGridN * sizeof(float4));
  if (e == cudaSuccess) e = cudaMemcpy(dM, M, n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(de, ecc, ne * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dtab, sc.data(), kGridN * sizeof(double2), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dtabf, scf.data(), kGridN * sizeof(float4), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    int blocks = int(std::min<int64_t>((n + 255) / 256, 148 * 8));
    if (use_table)
      kepler_grid_table_kernel<<<blocks, 256>>>(dM, ecc[0], n, dtab, dtabf, dE, dS, dC, make_hot_consts());
    else
      kepler_grid_kernel<<<blocks, 256>>>(dM, de, n, ecc_is_scalar, dtab, dtabf, dE, dS, dC, make_hot_consts());
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(E, dE, n * sizeof(double), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && sinE) e = cudaMemcpy(sinE, dS, n * sizeof(double), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && cosE) e = cudaMemcpy(cosE, dC, n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(dM); cudaFree(de); cudaFree(dE); cudaFree(dS); cudaFree(dC); cudaFree(dtab); cudaFree(dtabf);
  if (e != cudaSuccess) return fail(EMP_ECUDA, cudaGetErrorString(e));
  return EMP_OK;
}
