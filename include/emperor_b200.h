/*
 * emperor_b200.h — C-ABI of the B200-native EMPEROR hot path.
 *
 * The reference (ReddTea/astroemperor 0.9.10) is pure Python: its hot path is
 * the *generated* functions my_model / my_likelihood / my_prior that
 * reddemcee.PTSampler calls once per walker per step.  There is
 * no FFI in the reference; the entry points below are what a ctypes binding for
 * that path binds (INTEGRATION.md shows the stub).  Each entry cites the
 * reference interface it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - every function returns 0 on success or a negative EMP_E* code; the text of
 *     the last error of the calling thread is available from emp_last_error().
 *   - "dev" pointers are CUDA device pointers on the handle's device, "host"
 *     pointers are ordinary host memory.  All arrays are C-contiguous FP64 unless
 *     stated.
 *   - a handle owns one CUDA stream; calls on one handle are not re-entrant,
 *     different handles may be used from different threads.
 *   - there is NO CPU fallback: if no CUDA device is usable emp_create fails.
 */
#ifndef EMPEROR_B200_H
#define EMPEROR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMP_ABI_VERSION 6

#define EMP_MAX_KEP 10   /* Keplerian blocks                                  */
#define EMP_MAX_INS 16   /* instruments (offset / jitter entries)             */
#define EMP_MAX_DIM 128  /* length of the full parameter vector (free+fixed)  */
#define EMP_MAX_ACC 4    /* polynomial acceleration order                     */
#define EMP_MAX_MA 4     /* moving-average order                              */
#define EMP_MAX_PERIODIC 4 /* Sinusoid / MagneticCycle blocks                  */
#define EMP_MAX_SAI 4    /* stellar-activity index columns per instrument     */
#define EMP_MAX_PRIOR_OPS (2 * EMP_MAX_DIM + 2 * EMP_MAX_KEP + 8)

/* error codes */
#define EMP_OK 0
#define EMP_EINVAL (-1)   /* bad argument / unsupported descriptor            */
#define EMP_ECUDA (-2)    /* CUDA runtime error (see emp_last_error)          */
#define EMP_ENODEV (-3)   /* no usable sm_100 device                          */
#define EMP_ENOMEM (-4)
#define EMP_EUNSUPPORTED (-5) /* feature exists in the reference but not here */

/* Keplerian templates: support/models/kep0{0,1,2,3,4,6,7}.model, akep00.model */
enum EmpKepModel {
  EMP_KEP00 = 0, /* per, A, phase, ecc, w        ; ((1+e)/(1-e))**0.5          */
  EMP_KEP01 = 1, /* per, A, phase, S, C          ; ecc<1e-6 -> w=0             */
  EMP_KEP02 = 2, /* lnP, As, Ac, S, C            ; ecc<1e-5 -> w=0             */
  EMP_KEP03 = 3, /* per, A, tp, ecc, w           ; M = freq*(t-tp)             */
  EMP_KEP04 = 4, /* per, A, tp, S, C             ; ecc<1e-5                    */
  EMP_AKEP00 = 5,/* per, A, pha, ecc, w, I, Om   ; RV part ignores I, Om       */
  EMP_KEP06 = 6, /* lnP, A, phase, ecc, w                                      */
  EMP_KEP07 = 7  /* lnP, A, phase, S, C          ; ecc<1e-6                    */
};

/* support/priors/{Uniform,Normal,Jeffreys,Isotropic,Fixed}.prior */
enum EmpPriorKind {
  EMP_PRIOR_UNIFORM = 0,   /* a0 = logZ                                        */
  EMP_PRIOR_NORMAL = 1,    /* a0 = mu, a1 = s, a2 = logZ, a3 = log(s*sqrt(2pi))*/
  EMP_PRIOR_JEFFREYS = 2,  /* implemented as uniform in the reference; a0=logZ */
  EMP_PRIOR_ISOTROPIC = 3, /* log(0.5 sin x) - a0                              */
  EMP_PRIOR_FIXED = 4      /* contributes 0                                    */
};

/* my_prior is a straight-line program (emp.py:182-254); one op per line. */
enum EmpPriorOpKind {
  EMP_POP_PARAM = 0,  /* lp += Prior(theta_full[i0])                           */
  EMP_POP_CHECK = 1,  /* if lp == -inf: return lp   (end of each block)        */
  EMP_POP_SUMSQ = 2   /* x = theta_full[i0]**2 + theta_full[i1]**2; lp += Prior(x) */
};

typedef struct EmpPriorOp {
  int32_t op;    /* EmpPriorOpKind */
  int32_t prior; /* EmpPriorKind   */
  int32_t i0, i1;
  double lo, hi;
  double a0, a1, a2, a3;
} EmpPriorOp;

/* moving-average handling */
enum EmpMaMode {
  EMP_MA_NONE = 0,
  EMP_MA_REFERENCE_NOOP = 1, /* support/models/moav00.model: parameters only enter the prior */
  EMP_MA_GLOBAL = 2          /* support/models/moav01.model: recurrence over all points      */
};

/*
 * Model descriptor: what emp_model.py:706-781 (_write_model_RV) and
 * emp.py:182-254 (_write_prior_reddemcee) bake into the generated script.
 * All *_off fields index the FULL theta (after the fixed parameters were
 * re-inserted, emp_model.py:709-711).
 */
typedef struct EmpModelDesc {
  int32_t abi_version; /* = EMP_ABI_VERSION */
  int32_t ndim_free;   /* sampler dimension (model.ndim__)                     */
  int32_t ndim_full;   /* len(theta) after np.insert of the fixed values       */
  int32_t n_kep;
  int32_t kep_model[EMP_MAX_KEP]; /* EmpKepModel */
  int32_t kep_off[EMP_MAX_KEP];
  int32_t acc_order;   /* 0 = no AccelerationBlock (support/models/acc.model)  */
  int32_t acc_off;
  int32_t n_ins;       /* instruments; Flag values are 1..n_ins                */
  int32_t offset_off;  /* support/models/offset00.model                        */
  int32_t has_jitter;  /* support/models/jitter00.model                        */
  int32_t jitter_off;
  int32_t ma_mode;     /* EmpMaMode */
  int32_t ma_order;
  int32_t ma_off;
  int32_t am_enabled;  /* Hipparcos-Gaia block (emp_model.py:1232-1672)        */
  int32_t am_offset_off; /* AstrometryOffsetBlock: 5 params                    */
  int32_t am_jitter_off; /* AstrometryJitterBlock: J_H, J_G                    */
  int32_t n_prior_ops;
  /* periodic blocks, evaluated AFTER the MA block like the reference orders them
   * (emp.py:2628-2651): support/models/sinusoid00.model (per, A, phase) and
   * magneticcycle00.model (per, A1, A2, phase1, phase2) */
  int32_t n_periodic;
  int32_t periodic_kind[EMP_MAX_PERIODIC]; /* 0 = sinusoid, 1 = magnetic cycle */
  int32_t periodic_off[EMP_MAX_PERIODIC];
  /* StellarActivityBlock (support/models/sai00.model, emitted once per activity column,
   * emp_model.py:736-745): model0 += theta_sa[j] * SAI{j+1}_, also AFTER the MA block.  Column j
   * belongs to one instrument (its entries are 0 elsewhere, qol_utils.py:74-95, 198), columns are
   * numbered instrument by instrument: instrument i owns sai_count[i] columns starting at
   * sum(sai_count[:i]); theta_sa = theta_full[sai_off : sai_off + n_sai]. */
  int32_t n_sai;
  int32_t sai_off;
  int32_t sai_count[EMP_MAX_INS];
  int32_t free_to_full[EMP_MAX_DIM]; /* full index of free parameter j         */
  double full_init[EMP_MAX_DIM];     /* fixed values at their full index, 0 elsewhere */
  EmpPriorOp prior_ops[EMP_MAX_PRIOR_OPS];
} EmpModelDesc;

/*
 * Hipparcos-Gaia constants: the arrays emp_model.py:610-702 (_write_data_AM)
 * loads into the generated script (qol_utils.py:305-446 computes them once on
 * the host).  All pointers are HOST pointers; emp_create copies them.
 */
typedef struct EmpAmData {
  int32_t n_hipp;      /* len(data_iad_hipp)                                   */
  int32_t n_gost;      /* len(data_iad_gost) after filtering                   */
  int32_t n_mask2;     /* sum(mask_GDR2)                                       */
  int32_t n_mask3;     /* sum(mask_GDR3)                                       */
  double common_t;     /* RV time origin                                       */
  const double *time_hipp;   /* [n_hipp] BJD                                   */
  const double *cpsi_hipp, *spsi_hipp, *epoch_hipp, *parf_hipp, *res_hipp, *sres_hipp; /* [n_hipp] */
  const double *time_gost;   /* [n_gost] BJD                                   */
  const double *cpsi_gost, *spsi_gost, *parf_gost; /* [n_gost]                 */
  const int32_t *idx_mask2;  /* [n_mask2] indices into the gost arrays         */
  const int32_t *idx_mask3;  /* [n_mask3]                                      */
  const double *gsv2;        /* [5, n_mask2] AM_GSV['GDR2']                    */
  const double *gsv3;        /* [5, n_mask3] AM_GSV['GDR3']                    */
  const double *inv_cov;     /* [3,5,5] AM_inv_COV                             */
  const double *log_det_cov; /* [3]                                            */
  const double *astro_gost;  /* [2,5] AM_astro_gost.values (dra,ddec,plx,pmra,pmdec) */
  const double *catalogs;    /* [3,7] ref_epoch, ra, dec, parallax, pmra, pmdec, rv */
} EmpAmData;

typedef struct EmpHandle EmpHandle;

/* ---- lifecycle ---------------------------------------------------------- */

/* Replaces the data/constant section of the generated script
 * (emp_model.py:406-433 _write_data_RV: X_, Y_, YERR_, mask{i}).
 * t, y, yerr, flag: HOST arrays of n points in time order; flag in 1..n_ins.
 * am: NULL unless desc->am_enabled. */
int emp_create(const EmpModelDesc *desc, const double *t, const double *y, const double *yerr,
               const int32_t *flag, int64_t n, const EmpAmData *am, int device, EmpHandle **out);
int emp_destroy(EmpHandle *h);
const char *emp_last_error(void);
int emp_abi_version(void);
/* The CUDA stream (cudaStream_t as void*) all work of this handle is queued on. */
int emp_stream(EmpHandle *h, void **stream);
/* Queue all further work of this handle on a caller-owned stream (e.g. torch's current
 * stream, so that torch copies and these kernels are ordered without extra syncs). */
int emp_set_stream(EmpHandle *h, void *stream);
int emp_synchronize(EmpHandle *h);

/* Stellar-activity columns of a model with desc->n_sai > 0 (the SAI{j}_ arrays of the generated
 * script, emp_model.py:425-433): sai_host[j*n + i] = column j at point i, j < n_sai, same point
 * order as emp_create's arrays.  Must be called once before the first evaluation. */
int emp_attach_sai(EmpHandle *h, const double *sai_host, int32_t n_sai);

/* ---- likelihood / prior -------------------------------------------------- */

/* Batched my_likelihood + my_prior (support/likelihoods/00.like:3-5, a00.like:3-8,
 * emp.py:190-254): for each of n_eval rows of theta_dev[n_eval, ndim_free] writes
 * logl_dev[i] and logp_dev[i].  Rows whose prior is -inf get logl = -inf without
 * evaluating the model (emcee semantics).  Asynchronous on the handle's stream. */
int emp_logl_batch(EmpHandle *h, const double *theta_dev, int64_t n_eval, double *logl_dev,
                   double *logp_dev);
/* Same with host buffers: H2D copy, kernel, D2H copy, synchronises. */
int emp_logl_batch_host(EmpHandle *h, const double *theta_host, int64_t n_eval, double *logl_host,
                        double *logp_host);
/* my_model(theta) for ONE theta (emp_model.py:706-781): model0[n], err20[n] to host.
 * Used by the reference's post-processing (emp.py:1546-1558). */
int emp_model_host(EmpHandle *h, const double *theta_host, double *model_host, double *err2_host);

/* kepler.solve(M, ecc) (kepler.py 0.0.7; call sites support/models/kep00.model:6 ... akep00.model:5,
 * emp_model.py:1325): eccentric anomaly E[i] of M[i], ecc[i] (or ecc[0] if ecc_is_scalar), host
 * buffers.  Runs kepler.py's own scheme (Markley starter + its single high-order refinement): the solver
 * the likelihood kernel uses under EMP_SOLVER_KEPLERPY and for eccentricities above 0.98. */
int emp_kepler_solve_host(const double *M, const double *ecc, int64_t n, int ecc_is_scalar, double *E,
                          int device);

/* The solver the likelihood kernel actually runs per (walker, planet, datapoint) — the grid-anchored
 * core of emp_device.cuh (sin/cos table + FP32 Halley step + FP64 Newton correction), exposed so that
 * parity tests can pin it element by element against the same kepler.solve call sites: E[i] in
 * [0, 2pi] plus, if the pointers are not NULL, the sin E and cos E the RV term is built from.  Elements
 * with ecc outside [0, 0.98] or |M| >= 1e12 take the kepler.py-style refinement, as in the kernel. */
int emp_kepler_grid_host(const double *M, const double *ecc, int64_t n, int ecc_is_scalar, double *E,
                         double *sinE, double *cosE, int device);

/* The same core with the per-walker STARTER TABLE the likelihood kernel builds in its prologue (emp_device.cuh,
 * kernel v10) instead of the Markley starter: one eccentricity in [0, 0.8] for the whole array, |M| < 1e12.  Exists
 * so that the production path of table-served planets can be pinned element by element too. */
int emp_kepler_grid_table_host(const double *M, double ecc, int64_t n, double *E, double *sinE, double *cosE,
                               int device);

/* ---- parallel-tempering step -------------------------------------------- */

/* One emcee RedBlue stretch-move step of every temperature held by this handle
 * (reddemcee.PTSampler / emcee 3.1.6 RedBlueMove + StretchMove(a=2), SURVEY.md §3.3 and §8a
 * row A15), from host-supplied draws that were already copied to the device.  H = W/2.
 *   p        [T, W, ndim]  walker positions (updated in place)
 *   logl     [T, W], logp [T, W]  (updated in place)
 *   betas    [T]
 *   half_idx [T, 2, H] int32  walkers of split 0 / split 1, ascending (the shuffled
 *                             `arange(W) % 2` of RedBlueMove.propose, as index lists)
 *   zz       [T, 2, H]     stretch factor ((a-1)u+1)^2/a of the j-th walker of split s
 *   rint     [T, 2, H] int32  partner index inside the complementary half
 *   factors  [T, 2, H]     (ndim-1)*ln zz   (host computes the log so the decision is bit-exact)
 *   lnu      [T, 2, H]     log of the accept uniform
 *   accepted [T, W] uint8  (output; 1 where the proposal was accepted)
 * Split 0 is proposed/evaluated/accepted first, then split 1 (emcee order): per half one
 * proposal + prior kernel and one likelihood kernel whose epilogue does the Metropolis accept
 * (joint RV + astrometry models: a third launch, the astrometric kernel, accepts). */
int emp_pt_stretch_step(EmpHandle *h, int32_t T, int32_t W, double *p, double *logl, double *logp,
                        const double *betas, const int32_t *half_idx, const double *zz,
                        const int32_t *rint, const double *factors, const double *lnu,
                        uint8_t *accepted);

/* Adjacent-temperature swap sweep, hot -> cold (ptemcee lineage).
 *   logl_all [T, W]   log-likelihood of every temperature of the ladder (read only)
 *   betas    [T]
 *   perm     [T-1, 2, W] int32  pair j couples temperature i=j+1 with i-1=j:
 *                               perm[j,0,:] indexes walkers of i, perm[j,1,:] of i-1
 *   lnu      [T-1, W]
 * Outputs:
 *   src      [T, W] int32  flat index (t*W + w) of the slot whose walker ends up in (t, w)
 *   n_acc    [T-1] int32   accepted swaps per pair
 * Every rank of a sharded ladder replays this identically from the all-gathered logl.
 * (No ladder adaptation: that is emp_pt_sweep / emp_pt_sweep_swap.)  T <= 2048. */
int emp_pt_swap_plan(EmpHandle *h, int32_t T, int32_t W, const double *logl_all,
                     const double *betas, const int32_t *perm, const double *lnu, int32_t *src,
                     int32_t *n_acc);
/* Row gather that applies a swap plan: for r in [0, n_rows):
 *   p_out[r,:] = p_in[src[r],:], logl_out[r] = logl_in[src[r]], logp_out[r] = logp_in[src[r]].
 * `src` indexes rows of the *_in arrays (the caller translates global plan indices when the
 * ladder is sharded and remote rows were staged behind the local ones). */
int emp_pt_gather_rows(EmpHandle *h, int64_t n_rows, int32_t ndim, const int32_t *src,
                       const double *p_in, const double *logl_in, const double *logp_in,
                       double *p_out, double *logl_out, double *logp_out);
/* Proposals whose log-likelihood came out NaN (rejected; emcee would raise). */
int emp_nan_count(EmpHandle *h, uint32_t *count);

/* One whole sweep of `sampler.run_mcmc(p1, nsweeps=, nsteps=)` (support/endit_reddemcee.scr:3): nsteps
 * stretch steps of every local temperature, the hot -> cold swap sweep, the ladder adaptation
 * (reddemcee adapt_tau / adapt_nu / adapt_mode 0, defaults emp.py:2370-2380), the
 * tsw / smd / beta histories and the chain store — all on the device, no host synchronisation.
 * All pointers are DEVICE pointers on the handle's device unless stated. */
#define EMP_MAX_PEERS 16
typedef struct EmpPtSweep {
  int32_t T_loc, W, nsteps; /* local temperatures, walkers (even), stretch steps per sweep        */
  int32_t T_all;            /* temperatures of the whole ladder (= T_loc unless sharded)          */
  int32_t n_ranks, rank;    /* ladder sharded over n_ranks GPUs; this handle is `rank`            */
  int32_t strided;          /* 1: rank r holds temperatures r, r+G, ...; 0: contiguous blocks     */
  int32_t use_graph;        /* 1: the sweep is captured once into a CUDA graph and replayed       */
  double *p, *logl, *logp;             /* state [T_loc, W, ndim], [T_loc, W], [T_loc, W]          */
  double *p_alt, *logl_alt, *logp_alt; /* the swap writes the new state here (caller swaps roles) */
  double *betas;            /* [T_all] ladder; adapted in place when adapt != 0                   */
  const int32_t *half_idx;  /* stretch draws [nsteps, T_loc, 2, H], see emp_pt_stretch_step        */
  const double *zz;
  const int32_t *rint;
  const double *factors;
  const double *lnu;
  const int32_t *perm;      /* swap draws of the whole ladder [T_all-1, 2, W] (NULL if T_all == 1) */
  const double *lnu_swap;   /* [T_all-1, W]                                                        */
  uint8_t *accepted;        /* [T_loc, W] accept mask of the last stretch step                     */
  int32_t *n_accepted;      /* [T_loc, W] accepted moves per walker, accumulated (may be NULL)     */
  int32_t *src;             /* [T_all, W] swap plan of this sweep (output)                         */
  int32_t *n_acc;           /* [T_all-1] accepted swaps per pair (output)                          */
  int32_t adapt;            /* 1: adapt the ladder after the swap sweep (needs T_all > 2)          */
  int32_t thin;             /* store every thin-th stretch step (>= 1)                             */
  double adapt_tau, adapt_nu;
  int64_t *sweep_counter;   /* device scalar: sweeps done (reddemcee's `time`); incremented        */
  int64_t *step_counter;    /* device scalar: stretch steps begun; incremented                     */
  double *beta_hist;        /* [hist_cap, T_all] ladder after every sweep (may be NULL)            */
  int32_t *nacc_hist;       /* [hist_cap, T_all-1] swap counts of every sweep (may be NULL)        */
  double *smd_hist;         /* [hist_cap, T_loc] swap mean distance per sweep (needs D; may be NULL) */
  int64_t hist_cap;
  const double *D;          /* [ndim] prior widths `sampler.D_` (emp.py:595-602) or NULL           */
  double *chain, *chain_ll, *chain_lp; /* [store_cap, T_loc, W, ndim] / [store_cap, T_loc, W] or NULL */
  int64_t store_cap;
  int32_t store_ring;       /* 1: slot index wraps at store_cap (host streams the ring out)        */
  int32_t perm_hot_sorted;  /* 1: perm[j, 0, :] is the identity for every pair (the host lists the pairs by
                             * their slot in the warmer row): the plan kernel then keeps that row in registers */
  /* sharded ladder only: CURRENT buffers (p, logl, logp) of every rank, peer HBM mapped with
   * emp_ipc_open; entry `rank` must equal p / logl / logp */
  const double *peer_p[EMP_MAX_PEERS], *peer_logl[EMP_MAX_PEERS], *peer_logp[EMP_MAX_PEERS];
  const double *logl_all;   /* [T_all, W] all-gathered logL (sharded, NCCL exchange) or NULL          */
  /* sharded ladder, peer-push exchange (no NCCL, emp_pt_sweep runs the whole sweep): the gathered blocks
   * (emp_gather_block_bytes, zero-initialised, one per sweep parity) of every rank, peers mapped with
   * emp_ipc_open.  After the stretch phase a rank writes its rows of logL and of the swap draws (perm /
   * lnu_swap then hold THIS rank's T_loc pair rows: row t = the pair whose index is the rank's t-th temperature)
   * into every block over NVLink and raises its flag there; the plan kernel waits for the flags of the sweep.
   * All NULL: not used. */
  void *peer_gath[2][EMP_MAX_PEERS];
} EmpPtSweep;
/* Size of a gathered block of the peer-push exchange: [logL | swap uniforms | partner slots] of the whole ladder
 * in ladder order + one arrival flag per rank. */
int emp_gather_block_bytes(int32_t T_all, int32_t W, int64_t *bytes);
/* The whole sweep: single GPU (6 launches at nsteps = 1), or a sharded ladder with the peer-push exchange
 * (7 launches: + pt_publish_kernel; nothing but this library's kernels, so the sweep is graph-capturable). */
int emp_pt_sweep(EmpHandle *h, const EmpPtSweep *s);
/* k consecutive single-GPU sweeps replayed from ONE graph launch (argument blocks s[0..k-1]; the state parity
 * alternates and the draw pointers advance from block to block).  When draws_host is not NULL the graph begins
 * with one copy node that uploads draws_bytes from the PINNED host block draws_host to draws_dev — the k sweeps'
 * draws, which the blocks point into.  For ensembles whose sweep takes tens of microseconds (BASELINE configs
 * 1-3) this removes the per-sweep host work of `sampler.run_mcmc(p1, nsweeps=, nsteps=)`
 * (support/endit_reddemcee.scr:3): draw, stage, upload and launch happen once per chunk. */
int emp_pt_sweep_chunk(EmpHandle *h, const EmpPtSweep *s, int32_t k, void *draws_dev, const void *draws_host,
                       int64_t draws_bytes);
/* Sharded ladder: the stretch phase, then — after the caller all-gathered logL (and the swap draws)
 * over NCCL — the replicated swap plan + adaptation and the row gather straight from the owners' HBM. */
int emp_pt_sweep_stretch(EmpHandle *h, const EmpPtSweep *s);
int emp_pt_sweep_swap(EmpHandle *h, const EmpPtSweep *s);

/* Device memory that can be shared with the other ranks of the node (CUDA IPC over NVLink/NVSwitch):
 * the sharded swap reads the rows it needs from the owners' ensembles instead of all-gathering them. */
int emp_dev_alloc(int device, int64_t bytes, void **ptr);
int emp_dev_free(int device, void *ptr);
int emp_ipc_export(int device, void *ptr, unsigned char handle64[64]);
int emp_ipc_open(int device, const unsigned char handle64[64], void **ptr);
int emp_ipc_close(int device, void *ptr);

/* ---- host-supplied random draws ------------------------------------------- */
/* The draws of a sweep are what the reference stack consumes from numpy.random.RandomState (emcee 3.1.6
 * RedBlueMove.propose + StretchMove.get_proposal per temperature, the swap sweep's permutations and uniforms;
 * order: astroemperor_b200/draws.py).  These entry points restate the legacy RandomState algorithms (MT19937,
 * random_sample, shuffle / permutation, randint) bit for bit, one stream per temperature and per adjacent swap
 * pair, and fill caller-owned HOST buffers (e.g. the pinned staging buffer) with several threads.  Host only: no
 * CUDA call is made. */
typedef struct EmpDrawStreams EmpDrawStreams;
/* keys [n_streams, 624], pos [n_streams]: `RandomState.get_state()[1:3]` of every stream; n_threads: size of
 * the persistent worker pool (the calling thread included). */
int emp_draws_create(int32_t n_streams, const uint32_t *keys, const int32_t *pos, int32_t n_threads,
                     EmpDrawStreams **out);
int emp_draws_destroy(EmpDrawStreams *d);
int emp_draws_get_state(EmpDrawStreams *d, int32_t stream, uint32_t *key624, int32_t *pos);
/* All draws of one sweep, streams in parallel.  Stretch: nsteps RedBlue steps of the temperatures whose stream
 * indices are temp_streams[n_temps]; outputs [nsteps, n_temps, 2, H]: half_idx (walkers of each split, ascending),
 * u_zz (uniform behind the stretch factor), rint (partner index), u_acc (accept uniform).  Swap: the pairs
 * pair_streams[n_rows] (index < 0 = padding row): perm [n_rows, 2, W], u_swap [n_rows, W]. */
int emp_draws_sweep(EmpDrawStreams *d, const int32_t *temp_streams, int32_t n_temps, int32_t W, int32_t nsteps,
                    int32_t *half_idx, double *u_zz, int32_t *rint, double *u_acc, const int32_t *pair_streams,
                    int32_t n_rows, int32_t *perm, double *u_swap);

/* The same for k consecutive sweeps in one call: sweep q's arrays start q * stride_bytes behind the given
 * pointers (the chunk layout of emp_pt_sweep_chunk); identical to k successive emp_draws_sweep calls. */
int emp_draws_sweeps(EmpDrawStreams *d, int32_t k, int64_t stride_bytes, const int32_t *temp_streams,
                     int32_t n_temps, int32_t W, int32_t nsteps, int32_t *half_idx, double *u_zz, int32_t *rint,
                     double *u_acc, const int32_t *pair_streams, int32_t n_rows, int32_t *perm, double *u_swap);

/* ---- introspection -------------------------------------------------------- */
/* Number of kernels this handle has launched since creation (bench.py gpu_launches). */
int emp_launch_count(EmpHandle *h, int64_t *count);
/* Number of CUDA graphs emp_pt_sweep has captured so far (a steady run replays two: the state and the
 * draw staging are double-buffered). */
int emp_graph_captures(EmpHandle *h, int64_t *count);
/* Kepler solver of the likelihood kernel.  EMP_SOLVER_GRID (default): Markley starter, then the
 * grid-anchored refinement of emp_device.cuh (same root as kepler.solve to ~1 ulp, about half the FP64
 * instructions).  EMP_SOLVER_KEPLERPY: Markley starter + the single high-order refinement of kepler.py
 * 0.0.7 (the iteration the reference's kepler.solve call sites run) for every planet; the grid solver
 * already hands eccentricities above 0.98 to it.  Both agree to ~1e-14 relative on logL (tests). */
#define EMP_SOLVER_GRID 0
#define EMP_SOLVER_KEPLERPY 1
int emp_set_solver(EmpHandle *h, int solver);

/* Per-launch device timing of the likelihood kernel: while enabled every launch is bracketed
 * by CUDA events on the handle's stream; emp_timing_collect synchronises, returns the summed
 * kernel time and the number of launches since the last collect, and resets. */
int emp_set_timing(EmpHandle *h, int enable);
int emp_timing_collect(EmpHandle *h, double *total_ms, int64_t *n_launches);
/* Diagnostics of the PT step since creation: out4 = { proposals, proposals inside the prior
 * support (the ones whose likelihood was evaluated), accepted, NaN likelihoods }. */
int emp_counters(EmpHandle *h, uint64_t *out4);
/* Measured FP64 FMA throughput of the device in TFLOP/s (8 independent DFMA chains per
 * thread): the roofline denominator of this FP64-pipe-bound path. */
int emp_fp64_peak(int device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* EMPEROR_B200_H */
