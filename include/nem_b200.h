/*
 * nem_b200.h -- public C ABI of libnem_b200.so, the B200-native NEM partitioning engine.
 *
 * Plain C types only (no torch, no C++): this is the boundary a maintainer of
 * labgem/pangenomeNEM binds instead of the reference's ppanggolin/NEM C sources.
 * File:line citations refer to the reference tree (ppanggolin/...).
 *
 * Layers:
 *   1. nem()            drop-in for the reference's only exported symbol (NEM/nem_exe.h:23-35,
 *                       NEM/nem_exe.c:239-704), what NEM/nem.pyx wraps and
 *                       ppanggolin.py:1814-1826 calls.
 *   2. nem_b200_ex()    same, plus the knobs nem() hard-codes (site update, tie rule, E-step
 *                       sweeps, device) -- the reference only reaches them through its historic
 *                       CLI (NEM/nem_hlp.c:109-289).
 *   3. nemb_*           in-memory API: load a pangenome once (packed or dense host buffers),
 *                       run fits, read posteriors -- replaces the text-file round trip of
 *                       ppanggolin.py:821-930 + 1886-1972 for callers that can link it, and is
 *                       what the parity tests and bench.py drive.  Stage entry points expose
 *                       each kernel of the path for per-function parity tests.
 *
 * The library has NO CPU fallback: without a CUDA device every compute entry point fails
 * with NEMB_E_CUDA and nem() returns EXIT_E_SYSTEM (5).
 */
#ifndef NEM_B200_H
#define NEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------------------
 * 1. drop-in entry point  -- replaces NEM/nem_exe.c:239-251 (prototype NEM/nem_exe.h:23-35)
 * Return value: the reference's ExitET (NEM/lib_io.h:22-34, mapping nem_exe.c:637-703):
 *   0 OK, 1 W_RESULT (empty class, no output files), 2 E_ARGS, 3 E_FILE, 4 E_MEMORY,
 *   5 E_SYSTEM (no usable CUDA device), 6 E_BUG.
 * Inputs  <Fname>.str .dat .nei(type S) .m(init_mode 2); outputs <Fname>.uf|.cf, .mf and with
 * dolog .log and .stderr (nem_exe.c:272-282, 437-461, 1596-1781).
 * ------------------------------------------------------------------------------------- */
int nem(const char *Fname, const int nk, const char *algo, const float beta,
        const char *convergence, const float convergence_th, const char *format,
        const int it_max, const int dolog, const char *model_family, const char *proportion,
        const char *dispersion, const int init_mode);

/* knobs the reference fixes inside nem(): NemPara.SiteUpdate / TieRule / NbEIters / Seed
 * (nem_exe.c:351-361, nem_typ.h:330-341).  Zero-initialise for the reference's behaviour. */
typedef struct {
    int32_t update;      /* 0 = seq (UPDATE_SEQ, reference default), 1 = para (UPDATE_PARA) */
    int32_t sweep_impl;  /* seq+ncem only: 0 auto, 1 level-scheduled, 2 speculative fixed point */
    int32_t device;      /* CUDA ordinal, -1 = env NEM_B200_DEVICE or current device */
    int32_t n_random_inits; /* init_mode 1: number of random starts, 0 = 50 (nem_typ.h:94) */
    int64_t seed;        /* init_mode 1 RNG seed, 0 = 42 (the reference uses time(NULL)) */
    /* beta estimation, the CLI's -B / -G / -H (nem_hlp.c:220-245, defaults nem_typ.h:70-79):
     * beta_mode 0 fix, 1 psgrad, 2 heu_d, 3 heu_l (BetaET, nem_typ.h:128-135).  A zero
     * parameter means the reference's default. */
    int32_t beta_mode;
    int32_t grad_n_iter;                 /* -G nit  [1]  */
    float   grad_conv, grad_step;        /* -G conv [0.001] step [0 = Newton-like step] */
    float   heu_step, heu_max;           /* -H bstep [0.1] bmax [2.0] */
    float   heu_ddrop, heu_dloss;        /* -H ddrop [0.8] dloss [0.5] */
    float   heu_lloss;                   /* -H lloss [0.02] */
    int32_t reserved[7];
} nem_b200_extra;

int nem_b200_ex(const char *Fname, const int nk, const char *algo, const float beta,
                const char *convergence, const float convergence_th, const char *format,
                const int it_max, const int dolog, const char *model_family,
                const char *proportion, const char *dispersion, const int init_mode,
                const nem_b200_extra *extra);

/* Forked callers (ppanggolin.py:1039, command_line.py:262-281, 618): a process that was forked
 * after CUDA had been initialised cannot use CUDA; nem() then forwards the call to a helper process
 * (`nem_exe --serve`, started once per calling process).  nem_b200_helper_pid() = pid of the live
 * helper of this process (0: none); nem_b200_serve() is the helper's loop. */
int nem_b200_helper_pid(void);
int nem_b200_serve(int fd_in, int fd_out);

/* ---------------------------------------------------------------------------------------
 * 3. in-memory API
 * ------------------------------------------------------------------------------------- */
enum { NEMB_OK = 0, NEMB_W_EMPTYCLASS = 1, NEMB_E_ARG = 2, NEMB_E_FILE = 3, NEMB_E_MEMORY = 4,
       NEMB_E_CUDA = 5, NEMB_E_BUG = 6 };

enum { NEMB_ALGO_NEM = 0, NEMB_ALGO_NCEM = 1 };                 /* nem_typ.h:120-126 */
enum { NEMB_UPDATE_SEQ = 0, NEMB_UPDATE_PARA = 1 };             /* nem_typ.h:256-261 */
enum { NEMB_CONV_NONE = 0, NEMB_CONV_CLAS = 1, NEMB_CONV_CRIT = 2 }; /* nem_typ.h:272-278 */
enum { NEMB_PROP_EQUAL = 0, NEMB_PROP_K = 1 };                  /* nem_typ.h:200-205 */
enum { NEMB_DISP___ = 0, NEMB_DISP_K_ = 1, NEMB_DISP__D = 2, NEMB_DISP_KD = 3 }; /* :191-198 */
enum { NEMB_SWEEP_AUTO = 0, NEMB_SWEEP_LEVEL = 1, NEMB_SWEEP_SPEC = 2 };
enum { NEMB_BETA_FIX = 0, NEMB_BETA_PSGRAD = 1, NEMB_BETA_HEUD = 2, NEMB_BETA_HEUL = 3 }; /* nem_typ.h:128-135 */

typedef struct nemb_handle nemb_handle;

typedef struct {
    int32_t k;            /* number of classes, 1..16 */
    int32_t algo, update, conv, prop, disp;
    int32_t it_max;
    int32_t param_fixed;  /* .m flag 2: M-step skipped (nem_alg.c:1806) */
    int32_t dolog;        /* evaluate the criteria before/after every sweep (nem_alg.c:2361,2398) */
    int32_t sweep_impl;
    int32_t profile;      /* 1: time every stage with CUDA events (adds event overhead) */
    float   beta, conv_thr;
    /* BETA_PSGRAD (EstimBeta, nem_alg.c:2120-2230): beta is re-estimated after every M-step by
     * gradient ascent on the pseudo-likelihood of the current classification.  beta_mode is
     * NEMB_BETA_FIX or NEMB_BETA_PSGRAD here; the two heuristics are nemb_fit_beta_heuristic(). */
    int32_t beta_mode;
    int32_t grad_n_iter;  /* gradient iterations per EM iteration; 0 = 1 (nem_typ.h:76) */
    float   grad_conv;    /* stop when |gradient| < grad_conv * N; 0 = 0.001 (nem_typ.h:77) */
    float   grad_step;    /* > 0: beta += grad*step/N; 0: beta += grad / max(4 dsec, N/10) (nem_typ.h:78) */
    int32_t reserved[4];
} nemb_options;

typedef struct {
    int32_t status;       /* NEMB_OK or NEMB_W_EMPTYCLASS */
    int32_t iters;        /* EM iterations executed (nem_alg.c:1842) */
    int32_t converged;
    int32_t empty_class;  /* 1-based, 0 = none */
    double  U, D, L, M, Z, G;      /* nem_alg.c:2702-2751 */
    int64_t n_allnul;     /* families whose every class has zero density (last sweep) */
    int64_t n_ties;       /* ncem: exact arg-max ties in the last sweep */
    int64_t fixup_rounds; /* speculative sweep: fix-up rounds summed over all sweeps */
    int64_t kernel_launches;   /* kernels this fit enqueued */
    float   fit_ms;       /* device time of the whole fit, CUDA events on the engine's stream */
    float   ms_density, ms_sweep, ms_mstep, ms_criteria;  /* profile=1 only */
    int32_t n_density, n_sweep, n_mstep, n_criteria;      /* launches behind those sums */
    int32_t best_start;   /* nemb_fit_random: 1-based index of the retained start, else 0 */
    int32_t n_success;    /* nemb_fit_random: starts that ended without an empty class */
    /* profile=1: ms_density/n_density count only the passes that READ X; the iterations whose
     * class centres did not move rebuild logpf from the cached Hamming counts (ms_density_cached).
     * ms_mstep/n_mstep count the full X^T recounts; ms_mstep_delta the incremental updates. */
    float   ms_density_cached, ms_mstep_delta;
    int32_t n_density_cached, n_mstep_delta;
    int64_t exchanges;    /* row shards: all-gathers this fit issued */
    int64_t n_kept;       /* site evaluations the margin cache of the dense sweep saved (all sweeps) */
    float   beta;         /* beta of the last sweep and of the criteria (ModelParaT.Beta on return):
                             the option's, or the estimate of psgrad / of a heuristic */
    int32_t n_beta_tested;   /* nemb_fit_beta_heuristic: fits that ended without an empty class */
    /* persistent EM kernel (one cooperative launch per fit): launches of it, the
     * device-wide barriers they executed, the X passes and X^T recounts done inside them */
    int32_t pk_launches, pk_barriers, pk_x_passes, pk_recounts;
} nemb_result;

/* Per-iteration trace for the .log writer (nem_alg.c:1995-2052, 2620-2646). */
typedef void (*nemb_iter_cb)(void *user, int iter, const double crit_before[6],
                             const double crit_after[6], const float *prop, const float *center,
                             const float *disp, const float *nk);

int  nemb_create(nemb_handle **out, int device);
void nemb_destroy(nemb_handle *h);
const char *nemb_last_error(const nemb_handle *h);
int  nemb_set_stream(nemb_handle *h, void *cuda_stream);   /* default: an own non-blocking stream */

/* Load one pangenome; all pointers are HOST buffers, copied to HBM inside the call.
 * x_packed: uint32[n][words_per_row], genome d = bit d%32 of word d/32, padding bits zero.
 * row_ptr/col/wgt: CSR of the neighbour graph (NULL row_ptr = non-spatial, beta forced to 0,
 * nem_exe.c:570-574); order inside a row = order in the .nei file. */
int nemb_load_packed(nemb_handle *h, int n, int d, int words_per_row, const uint32_t *x_packed,
                     const int32_t *row_ptr, const int32_t *col, const float *wgt);
int nemb_load_dense_u8(nemb_handle *h, int n, int d, const uint8_t *x, const int32_t *row_ptr,
                       const int32_t *col, const float *wgt);
/* X already resident in HBM (device pointer, packed layout as above); CSR still from host. */
int nemb_load_packed_device(nemb_handle *h, int n, int d, int words_per_row,
                            const uint32_t *x_packed_dev, const int32_t *row_ptr,
                            const int32_t *col, const float *wgt);

/* ---------------------------------------------------------------------------------------
 * Row-sharded fits (one process per GPU).  The reference has no counterpart: its engine is one
 * single-threaded process.  Families are split into `world` contiguous
 * id ranges of shard_len = ceil(n_glob / world) rows; rank r owns rows [r*shard_len,
 * min(n_glob, (r+1)*shard_len)).  X is sharded; the neighbour graph (5 % of the bytes), theta
 * and the 1-byte labels are replicated.  Per EM iteration the ranks exchange
 *   - the M-step sufficient statistics S[K*D] + n[K] (all-gather + rank-ordered sum = a
 *     deterministic all-reduce), and
 *   - after every sweep round, the label slices (halo exchange of the 1-byte hard labels, or of
 *     the float posteriors for algo nem), so that the SEQUENTIAL sweep keeps its exact
 *     single-process result (speculative fixed point across ranks, DESIGN.md "Multi-GPU").
 * The only primitive the engine needs is an all-gather on device pointers, supplied through
 * this vtable (NCCL binding below; an in-process test double for single-GPU tests).
 * ------------------------------------------------------------------------------------- */
typedef struct nemb_comm {
    void   *ctx;
    int32_t rank, world;
    /* every rank contributes bytes_per_rank bytes from `send`; `recv` receives world blocks in
     * rank order.  `send` may be recv + rank*bytes_per_rank (in place).  Device pointers; the
     * work is enqueued on cuda_stream.  Returns 0 on success. */
    int   (*allgather)(void *ctx, const void *send, void *recv, size_t bytes_per_rank,
                       void *cuda_stream);
    void  (*destroy)(void *ctx);
} nemb_comm;

/* NCCL binding (libnccl.so.2 is dlopen'ed; NEM_B200_NCCL_LIB overrides the path).  Rank 0 makes
 * the 128-byte id and ships it to the other ranks by any means (torch.distributed broadcast in
 * pangenomenem_b200/sharded.py). */
int  nemb_nccl_unique_id(uint8_t id_out[128]);
int  nemb_comm_create_nccl(nemb_comm **out, const uint8_t id[128], int rank, int world);
/* In-process test double: `world` communicators for `world` host threads driving `world`
 * handles on ONE device; collectives are staged device copies ordered by events. */
int  nemb_comm_create_local(nemb_comm **out_array /*[world]*/, int world);
void nemb_comm_destroy(nemb_comm *c);
/* Attach before loading; NULL detaches (single GPU).  The handle does not own the comm. */
int  nemb_set_comm(nemb_handle *h, nemb_comm *comm);

/* Load this rank's shard: x_packed = rows [row0, row0+n_loc) (host buffer), graph = the GLOBAL
 * CSR over n_glob families (every rank passes the same).  Needs a comm with world > 1 or
 * row0 == 0 && n_loc == n_glob. */
int nemb_load_shard(nemb_handle *h, int n_glob, int row0, int n_loc, int d, int words_per_row,
                    const uint32_t *x_packed, const int32_t *row_ptr, const int32_t *col,
                    const float *wgt);
int nemb_load_shard_device(nemb_handle *h, int n_glob, int row0, int n_loc, int d,
                           int words_per_row, const uint32_t *x_packed_dev, const int32_t *row_ptr,
                           const int32_t *col, const float *wgt);
/* shard geometry helper: shard_len and this rank's [row0, row0+n_loc) */
void nemb_shard_range(int n_glob, int world, int rank, int *shard_len, int *row0, int *n_loc);

/* One fit = ClassifyByNemOneBeta's INIT_PARAM_FILE branch (nem_alg.c:1151-1169): blind sweep,
 * beta sweep, EM loop, final criteria.  theta (prop[K], center[K*D], disp[K*D], float32, host)
 * is the starting point on entry and the last M-step's estimate on return. */
int nemb_fit(nemb_handle *h, const nemb_options *opt, float *prop, float *center, float *disp,
             nemb_result *res);
int nemb_fit_logged(nemb_handle *h, const nemb_options *opt, float *prop, float *center,
                    float *disp, nemb_result *res, nemb_iter_cb cb, void *user);
/* init_mode 1 (RandNemAlgo, nem_alg.c:1574-1742): n_starts random starts, best by criterion. */
int nemb_fit_random(nemb_handle *h, const nemb_options *opt, int n_starts, int64_t seed,
                    float *prop, float *center, float *disp, nemb_result *res);

/* INIT_FILE (ClassifyByNemOneBeta, nem_alg.c:1091-1113): start NemAlgo from a given
 * classification t_init[n*k] (host; under ncem every row must be one-hot) instead of the two
 * initial sweeps.  The first M-step estimates theta from it (InitPara + MakeParaFromLabeled +
 * the first EstimPara of NemAlgo); a class without any family ends the call with
 * NEMB_W_EMPTYCLASS ("Class %d has no labeled observation", nem_alg.c:1338-1345).
 * prop/center/disp are outputs only.  Single GPU. */
/* Phase profile of the persistent EM kernel during the last fit: nanoseconds CTA 0 spent in
 * 0 init, 1 changed-rows scan, 2 delta statistics, 3 full recount, 4 closed forms + tables,
 * 5 X pass, 6 margin test, 7 active-list evaluation, 8 dense Jacobi round, 9 fix-up rounds
 * (barrier waits included); [10] = fix-up rounds executed. */
int nemb_get_persist_profile(nemb_handle *h, unsigned long long out12[12]);
/* Per EM iteration (first 12 of the last launch): ns of scan, delta/recount, closed forms, margin
 * test, evaluation, fix-up rounds; active sites of the evaluation (-1 = dense round); rounds. */
int nemb_get_persist_trace(nemb_handle *h, long long out96[96]);
int nemb_fit_from_partition(nemb_handle *h, const nemb_options *opt, const float *t_init,
                            float *prop, float *center, float *disp, nemb_result *res);

/* One evaluation of EstimBeta (nem_alg.c:2120-2230) on a classification t[n*k] (host): returns
 * the new beta in *beta_io and, when sums3 != NULL, the log pseudo-likelihood, its gradient and
 * minus its second derivative of the last gradient iteration.  Stage entry point of psgrad. */
int nemb_stage_estim_beta(nemb_handle *h, const nemb_options *opt, const float *t, float *beta_io,
                          double *sums3);

/* ClassifyByNemHeuBeta (nem_alg.c:731-992): estimate beta by a sweep of complete fits at
 * beta = 0, step, 2 step, ... <= max.  mode NEMB_BETA_HEUD: Hathaway criterion D, stop at the
 * first drop of its slope below -ddrop * N, else threshold its total loss at dloss;
 * NEMB_BETA_HEUL: mixture likelihood L, stop when it falls lloss * N under its maximum.  Like
 * the reference every fit starts from the all-zero classification and from the PARAMETERS THE
 * PREVIOUS FIT LEFT; the final fit at the estimate starts from the classification saved before
 * the drop (INIT_FILE) or, for heu_d without a detected drop, from scratch.  A zero field of
 * `hp` means the reference's default (nem_typ.h:71-75).  beta_trace/crit_trace (nullable, cap
 * entries): the tested betas and their criterion.  res->beta = the estimate.  Single GPU. */
typedef struct { float step, max, ddrop, dloss, lloss; } nemb_beta_heuristic;
int nemb_fit_beta_heuristic(nemb_handle *h, const nemb_options *opt, int mode,
                            const nemb_beta_heuristic *hp, float *prop, float *center,
                            float *disp, nemb_result *res, float *beta_trace, float *crit_trace,
                            int cap);

/* The pieces of nemb_fit_random, exposed so that every start can be checked on its own:
 * InitPara's whole-sample dispersion (nem_alg.c:1253-1265: the M-step with every family in the
 * first class), and MakeRandomPara for start number `start` (nem_alg.c:1381-1473: centres =
 * distinct random data rows, dispersion = sample dispersion / K, equal proportions).  Start s
 * draws from its own splitmix64 stream derived from (seed, s): reproducible and independent of
 * the order the starts run in. */
int nemb_sample_dispersion(nemb_handle *h, const nemb_options *opt, float *disp_sample /*[d]*/);
int nemb_random_start(nemb_handle *h, int k, int64_t seed, int start, const float *disp_sample,
                      float *prop, float *center, float *disp);
/* nemb_fit_random over n_workers host threads / streams sharing the resident pangenome (the
 * result does not depend on n_workers; nemb_fit_random uses 4, env NEM_B200_RANDOM_WORKERS) */
int nemb_fit_random_workers(nemb_handle *h, const nemb_options *opt, int n_starts, int64_t seed,
                            int n_workers, float *prop, float *center, float *disp,
                            nemb_result *res);

/* n = the GLOBAL number of families (every rank of a sharded fit holds all labels) */
int nemb_get_posteriors(nemb_handle *h, float *t_out /*[n*k]*/);
int nemb_get_labels(nemb_handle *h, int32_t *label_out /*[n]*/);   /* MAP, first maximum */
/* the same for families [first, first + count) only -- a rank of a sharded fit reads back its
 * own rows (nemb_shard_range) instead of the whole pangenome */
int nemb_get_labels_rows(nemb_handle *h, int first, int count, int32_t *label_out /*[count]*/);

/* loader products, for bit-exact packing / indexing tests */
int nemb_get_dims(const nemb_handle *h, int *n, int *d, int *words_per_row, int *nwt,
                  int *depth, int *nnz);
int nemb_get_packed(nemb_handle *h, uint32_t *out /*[n*wpr]*/);
int nemb_get_transposed(nemb_handle *h, uint32_t *out /*[d*nwt]*/);
/* the CSR as resident in HBM (any pointer may be NULL): row_ptr[n+1], col[nnz], wgt[nnz] */
int nemb_get_graph(nemb_handle *h, int32_t *row_ptr, int32_t *col, float *wgt);
int nemb_get_levels(nemb_handle *h, int32_t *level_of_site /*[n]*/);

/* stage entry points (each runs exactly the kernels nemb_fit uses for that step) */
int nemb_stage_density(nemb_handle *h, int k, const float *prop, const float *center,
                       const float *disp, int force_general, double *logpf_out /*[n*k]*/,
                       int32_t *hamming_out /*[n*k] or NULL (popcount path only)*/,
                       int *used_uniform);
int nemb_stage_sweep(nemb_handle *h, const nemb_options *opt, const double *logpf /*[n*k]*/,
                     float beta, float *t_inout /*[n*k]*/, int32_t *label_out /*[n] or NULL*/,
                     int64_t *fixup_rounds);
int nemb_stage_mstep(nemb_handle *h, const nemb_options *opt, const float *t /*[n*k]*/,
                     float *prop, float *center, float *disp, double *nk_out /*[k]*/,
                     double *skd_out /*[k*d]*/, int *empty_class);
int nemb_stage_criteria(nemb_handle *h, const nemb_options *opt, const double *logpf,
                        const float *t, float beta, double *crit6 /*U D L M Z G*/);

/* ---------------------------------------------------------------------------------------
 * 4. resample driver -- NEM on organism subsets without the text round trip.
 * Replaces, for callers that can link it, the per-sample work of the chunk loop of partition()
 * (ppanggolin.py:995-1105: sample `chunck_size` organisms, __write_nem_input_files
 * ppanggolin.py:821-930, nem(), vote) and of the evolution-curve workers
 * (command_line.py:262-281, 599-619).  The subsample is built ON THE DEVICE from the resident
 * pangenome exactly as the reference writes it: families without a selected organism are dropped
 * and the others renumbered in order (ppanggolin.py:847-852); an edge is kept when it exists in
 * at least one selected organism and its weight is that count (ppanggolin.py:862-880).
 * ------------------------------------------------------------------------------------- */
/* genome_mask: host, ceil(D/32) words, bit d = genome d selected.  edge_presence_dev: DEVICE
 * uint32[nnz][words_per_row] -- bit d of row e = CSR entry e exists in genome d -- or NULL for
 * the co-presence model (the edge exists wherever both families are present).  `dst` (same
 * device, single GPU) becomes a loaded pangenome of n_eff families x d_eff genomes whose genomes
 * are the selected ones in ascending order; fit it with nemb_fit().  The source graph must be
 * symmetric.  Buffers of dst are sized by the source, so repeated calls do not allocate. */
int nemb_subsample(nemb_handle *src, nemb_handle *dst, const uint32_t *genome_mask,
                   const uint32_t *edge_presence_dev, int *n_eff, int *d_eff);
/* original family id of every row of a device-built subsample (nem_file.index of the reference) */
int nemb_get_family_index(nemb_handle *dst, int32_t *index_out /*[n_eff]*/);

typedef struct {
    int32_t n_runs;          /* samples fitted */
    int32_t n_ok;            /* fits whose classes map to (persistent, shell, cloud) */
    int32_t n_inconsistent;  /* class/parameter consistency check failed (ppanggolin.py:1956-1957):
                                every family of the sample voted "undefined" */
    int32_t n_failed;        /* empty class: the reference writes no .uf, all families undefined */
    int64_t family_iterations;   /* sum over samples of n_eff x EM iterations */
    int64_t kernel_launches;
    double  fit_ms_sum;      /* device time of the fits (they overlap across workers) */
} nemb_batch_stats;

/* n_runs independent fits (K = 3, ncem, PPanGGOLiN's default initial parameters) on the genome
 * subsets genome_masks[n_runs][ceil(D/32)], run r with beta betas[r] (NULL: opt->beta).
 * n_workers host threads, each with its own stream and scratch handle, pull runs from a shared
 * counter so the latency-bound small fits overlap on the device.  votes_out (host, nullable):
 * int32[N][4] = how many samples put each family in persistent / shell / cloud / undefined
 * (cpt_partition of ppanggolin.py:997-1037).  iters_out (nullable): EM iterations of every run. */
int nemb_resample_batch(nemb_handle *src, int n_runs, const uint32_t *genome_masks,
                        const float *betas, const nemb_options *opt, int n_workers,
                        const uint32_t *edge_presence_dev, int32_t *votes_out, int32_t *iters_out,
                        nemb_batch_stats *stats);

/* ---------------------------------------------------------------------------------------
 * host-side loader / writers (no GPU needed): the NEM file contract as in-memory buffers.
 * Replaces ReadStrFile / ReadMatrixFile / ReadParamFile / ReadNeiFile (nem_exe.c:739-898,
 * 973-1091, 1278-1478) and SaveResults (nem_exe.c:1596-1781).
 * ------------------------------------------------------------------------------------- */
typedef struct {
    int32_t  n, d, words_per_row;
    int32_t  spatial;        /* 1 = type S (.nei read), 0 = type N */
    int32_t  nnz, max_neigh;
    int32_t  m_flag;         /* .m first token: 1 = initial, 2 = fixed; 0 = .m not read */
    uint32_t *x_packed;      /* [n][words_per_row] */
    int32_t  *row_ptr, *col; /* CSR, NULL when not spatial */
    float    *wgt;
    float    *prop, *center, *disp;   /* [k], [k*d], [k*d] when k > 0 */
} nemb_host_problem;

/* k = 0 skips <base>.m.  Returns a NEMB_* code; messages go to stderr. */
int  nemb_read_files(const char *base, int k, nemb_host_problem *out);
void nemb_free_host_problem(nemb_host_problem *p);
int  nemb_write_uf(const char *path, int n, int k, const float *t);
int  nemb_write_cf(const char *path, int n, const int32_t *label);
int  nemb_write_mf(const char *path, int k, int d, double U, double D, double L, double M,
                   float beta, const float *prop, const float *center, const float *disp);

const char *nemb_version(void);

#ifdef __cplusplus
}
#endif
#endif
