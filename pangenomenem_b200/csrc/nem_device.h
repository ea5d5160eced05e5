/*
 * nem_device.h -- internal thin C-ABI layer between the C host (nem_fit.c, nem_io.c,
 * nem_api.c) and the CUDA kernels (nem_kernels.cu).  Every nemk_* function enqueues work on
 * the given stream and returns; all pointers are DEVICE pointers.  Plain C types only.
 */
#ifndef NEM_DEVICE_H
#define NEM_DEVICE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NEMB_MAX_K 16

/* Per-class coefficients derived from theta, read by the density kernels.
 * uniform path: logpf = forb ? (H ? -inf : lp) : lp - (a*H + base)
 * general path: logpf = nul ? -inf : lp - (base + sum_{d: x=1} delta[k][d]) */
typedef struct {
    double lp[NEMB_MAX_K];
    double a[NEMB_MAX_K];
    double base[NEMB_MAX_K];
    int32_t forb[NEMB_MAX_K];   /* eps_k <= EPSILON: any mismatch => zero density */
    int32_t kind[NEMB_MAX_K];   /* popcount path: 0 general, 1 centre 0 everywhere (H = popc x),
                                   2 centre 1 everywhere (H = D - popc x), 3 all centres 1/2 (H = 0) */
    int32_t uniform_ok;         /* 1 when every class is popcount-eligible */
    int32_t empty_class;        /* 1-based index of an empty class (M-step) or 0 */
    int32_t halt;               /* set by nemk_iter_end when the fit is over (converged, or an empty
                                   class): every kernel of a later, speculatively enqueued iteration
                                   returns at once.  Must follow empty_class (the `skip` pair). */
    int32_t mu_changed;         /* the class bit masks differ from the previous tables (or forced):
                                   the cached Hamming counts are stale, the density pass must run */
    /* margin cache of the dense sweep (nemk_margins): dstep[k] bounds how far class k's score
     * lp - (a*H + base) can have moved, for any H in [0, D], between the previous tables and
     * these (+inf: unknown); drift accumulates 2*max_k dstep over the sweeps since the margins
     * were stored */
    double  dstep[NEMB_MAX_K];
    double  drift;
    /* persistent EM kernel: class k's bit masks differ from the previous tables (written, not
     * OR-ed, by the CTA that builds class k's tables: needs no reset between iterations) */
    int32_t mu_moved_k[NEMB_MAX_K];
} nemk_coef;

/* Margin cache of the ncem speculative sweep (exact shortcut).  m[i] = (best - second best score
 * of site i at its last evaluation) + the drift at that time.  While the class bit masks do not
 * move, theta only changes the per-class coefficients: no score can move by more than dstep, so a
 * site whose margin exceeds the drift accumulated since and none of whose later-or-equal
 * neighbours changed label in the previous sweep (stale flags) keeps its label: the dense round
 * copies it without touching the CSR.  m == NULL: feature off. */
typedef struct {
    float   *m;            /* [n_loc] */
    uint8_t *stale_cur;    /* [lab_len] set during the previous sweep, consumed (cleared) by this one */
    uint8_t *stale_next;   /* [lab_len] set by this sweep's label changes */
    int32_t  on;           /* host: this sweep may skip (same beta as the previous sweep) */
} nemk_margins;

/* Device scalars of one sweep / one iteration (host reads them back in one copy). */
typedef struct {
    int32_t changed;     /* ncem: number of labels != previous iteration's */
    int32_t nfix;        /* speculative sweep: number of fix-up rounds the sweep needed */
    int32_t allnul;      /* rows whose every class has zero density */
    int32_t ties;        /* ncem rows whose arg max was an exact tie */
    float   maxdiff;     /* nem: max |t - t_old| */
    int32_t pending;     /* row-sharded sweep: (reader, moved label) pairs across ranks found by the
                            last label exchange -- the same number on every rank; 0 = settled */
    int32_t changed_glob; /* row-sharded sweep: labels of ALL families != previous iteration's,
                            counted by every rank from the exchanged labels (same on every rank) */
    int32_t kept;        /* dense round: sites whose label was copied thanks to the margin cache */
} nemk_counters;

/* Device status block of one sweep / iteration, and the copy nemk_iter_end publishes into
 * mapped pinned host memory (no memcpy, no stream synchronisation: the host polls `seq`). */
typedef struct {
    nemk_counters cnt;
    double crit_before[6];
    double crit_after[6];
} nemk_iter_status;

typedef struct {
    nemk_counters cnt;             /* summed over the ranks (nfix, maxdiff: maximum) */
    double crit_before[6];
    double crit_after[6];
    int32_t empty_class, mu_changed, halt, pad;
    unsigned long long seq;        /* written last, after a system-wide fence */
} nemk_host_status;

typedef void *nemk_stream;

/* Where the consumers of log p_k f_k(x_i) (ncem sweeps, criteria) get it from: the materialised
 * array logpf[n_loc][K], or -- popcount density path -- the cached Hamming counts ham[n_loc][K]
 * plus the class coefficients: logpf = lp - (a*H + base), evaluated in registers with the density
 * kernels' own expression (same bits), so no N*K*8-byte array is written or read. */
typedef struct {
    const double *logpf;
    const int32_t *ham;      /* non-NULL selects the Hamming source */
    const nemk_coef *coef;
    int32_t wsum_any_order;  /* every edge weight is a small integer (checked at load): the fp64
                                context sums are exact in ANY order, so hubs may add them in
                                parallel; 0 = keep the file order of SumNeighsOfClass */
} nemk_lpsrc;

/* ---- loader */
void nemk_pack_u8(nemk_stream s, const uint8_t *x, int n, int d, int wpr, uint32_t *out);
void nemk_transpose_bits(nemk_stream s, const uint32_t *x, int n, int wpr, int d, int nwt,
                         uint32_t *xt);
/* the rows [row_base, row_base+rows) of x only (row_base % 256 == 0) into an already zeroed xt:
 * lets the loader transpose chunk by chunk behind the host->device copy */
void nemk_transpose_bits_rows(nemk_stream s, const uint32_t *x, int row_base, int rows, int wpr,
                              int d, int nwt, uint32_t *xt);

/* ---- theta -> tables */
void nemk_theta_tables(nemk_stream s, int k, int d, int wpr, const float *prop,
                       const float *center, const float *disp, nemk_coef *coef,
                       uint32_t *mask_xor, uint32_t *mask_valid, uint32_t *mask_f0,
                       uint32_t *mask_f1, double *delta, int force_mu_changed);

/* ---- E-step density */
/* cached != 0: `hamming` is the engine's persistent H cache; the X pass only runs when
 * coef->mu_changed, otherwise logpf is rebuilt from the cached counts (same formula, same bits) */
/* logpf may be NULL (only the Hamming counts are produced) */
void nemk_density_uniform(nemk_stream s, int k, int d, const uint32_t *x, int n, int wpr,
                          const nemk_coef *coef, const uint32_t *mask_xor,
                          const uint32_t *mask_valid, double *logpf, int32_t *hamming, int cached);
/* pop[i] = popc(x_i) (theta-independent; rows [0,n) of the x pointer given) */
void nemk_row_popcount(nemk_stream s, const uint32_t *x, int n, int wpr, int32_t *pop);
/* every class has a constant centre (coef->kind 1/2/3): H = P, D - P or 0 without reading X */
void nemk_ham_from_pop(nemk_stream s, int k, int n, int d, const nemk_coef *coef,
                       const int32_t *pop, int32_t *ham);
/* runs only when !coef->mu_changed: logpf from the cached Hamming counts */
void nemk_logpf_from_cache(nemk_stream s, int k, int n, const nemk_coef *coef,
                           const int32_t *hamming, double *logpf);
void nemk_density_general(nemk_stream s, int k, const uint32_t *x, int n, int d, int wpr,
                          const nemk_coef *coef, const uint32_t *mask_f0, const uint32_t *mask_f1,
                          const double *delta, double *logpf);

/* ---- loader: graph validation.  flags2[0] bit0 row_ptr broken, bit1 neighbour out of range,
 * bit2 some edge i->j has no j->i (reader lists differ from neighbour lists), bit3 some weight is
 * not an integer of magnitude <= 2^20 (context sums are then order dependent); flags2[1] = max degree */
void nemk_graph_check(nemk_stream s, int n, int nnz, const int32_t *row_ptr, const int32_t *col,
                      const float *wgt, int32_t *flags2);

/* index-sorted list of this rank's hubs (degree > 16), evaluated one warp per site by the sweeps
 * and the criteria; block_counts needs ceil(n_loc/1024) ints, list up to n_loc ints */
void nemk_heavy_list(nemk_stream s, int row0, int n_loc, const int32_t *row_ptr,
                     int32_t *block_counts, int32_t *list, int32_t *total);

/* ---- end of a sweep / an EM iteration: sum the ranks' counters, and when `decide` apply the
 * convergence test on the device (HasConverged `clas`, nem_alg.c:2075-2089: ncem = no label
 * changed, nem = max |t - t_old| < thr; conv: 0 none, 1 clas) and raise coef->halt when the fit is
 * over (converged or empty class), then publish everything to the mapped host slot. */
/* The same work fused into the last kernel of a sweep (nemk_sweep_ncem_fixup): host == NULL = not
 * fused. */
typedef struct {
    int32_t world;
    const nemk_counters *cnt_all;
    const nemk_iter_status *st;
    nemk_coef *coef;
    int32_t decide, ncem, conv;
    float thr;
    nemk_host_status *host;
    unsigned long long seq;
} nemk_iter_end_args;

/* world == 0: cnt_all is this rank's counter block of a row-sharded speculative sweep; its
 * changed_glob / pending are already global (nemk_mark_remote), the other fields stay local */
void nemk_iter_end(nemk_stream s, int world, const nemk_counters *cnt_all,
                   const nemk_iter_status *st, nemk_coef *coef, int decide, int ncem, int conv,
                   float thr, nemk_host_status *host_slot, unsigned long long seq);

/* ---- E-step sweeps.  label 255 = unlabelled (the reference's calloc'd ClassifM row).
 * `skip` (nullable) points at coef->empty_class: skip[0] = an empty class was found by the M-step
 * (the reference does not run the E-step then), skip[1] = coef->halt; either non-zero => the
 * kernel returns at once.
 * Row sharding: this rank owns the global rows [row0, row0+n_loc); labels, t, CSR, dirty flags and
 * work lists are indexed by GLOBAL family id, logpf by local row.  One GPU: row0 = 0, n_loc = N.
 * nemk_sweep_ncem_jacobi with copy_ranks > 1 also copies the other ranks' label slices
 * lab_in -> lab_out (slices of shard_len labels), so the fix-up rounds read the previous labels of
 * remote families until the first exchange.
 * Work lists: wl_a/wl_b used alternately, wl_cnt[4] rotating counters (round r: list r&1). */
void nemk_sweep_ncem_jacobi(nemk_stream s, int k, int row0, int n_loc, nemk_lpsrc lps,
                            const int32_t *row_ptr, const int32_t *col, const float *wgt,
                            double beta, const uint8_t *lab_in, uint8_t *lab_out, int32_t *dirty,
                            int32_t *wl, int32_t *wl_count, const int32_t *rrow_ptr,
                            const int32_t *rcol, const int32_t *heavy, int n_heavy,
                            nemk_counters *cnt, const int32_t *skip, int copy_ranks, int shard_len,
                            nemk_margins mg);
void nemk_sweep_ncem_fixup(nemk_stream s, int k, int row0, int n_loc, nemk_lpsrc lps,
                           const int32_t *row_ptr, const int32_t *col, const float *wgt, double beta,
                           const uint8_t *lab_old, uint8_t *lab_cur, int32_t *dirty, int32_t *wl_a,
                           int32_t *wl_b, int32_t *wl_cnt, int round, const int32_t *rrow_ptr,
                           const int32_t *rcol, nemk_counters *cnt, const int32_t *skip,
                           const nemk_iter_end_args *fused /* nullable: publish the status too */,
                           nemk_margins mg);
void nemk_sweep_ncem_fixup_round(nemk_stream s, int k, int row0, int n_loc, nemk_lpsrc lps,
                                 const int32_t *row_ptr, const int32_t *col, const float *wgt,
                                 double beta, const uint8_t *lab_old, uint8_t *lab_cur,
                                 int32_t *dirty, int32_t *wl_a, int32_t *wl_b, int32_t *wl_cnt,
                                 int round, const int32_t *rrow_ptr, const int32_t *rcol,
                                 nemk_counters *cnt, const int32_t *skip, nemk_margins mg);
/* after a label exchange: every rank scans ALL families.  A label that differs from the one seen
 * at the previous exchange (seen_in; the first exchange of a sweep passes the sweep's input labels)
 * queues this rank's later readers and counts, for every rank alike, the cross-rank (reader,
 * label) pairs -> cnt->pending; cnt->changed_glob = labels != lab_in.  seen_out receives lab_cur. */
void nemk_mark_remote(nemk_stream s, int n_glob, int row0, int n_loc, int shard_len,
                      const uint8_t *lab_cur, const uint8_t *lab_in, const uint8_t *seen_in,
                      uint8_t *seen_out, int32_t *dirty, int32_t *wl, int32_t *wl_count,
                      const int32_t *rrow_ptr, const int32_t *rcol, nemk_counters *cnt,
                      const int32_t *skip);
void nemk_sweep_ncem_level(nemk_stream s, int k, nemk_lpsrc lps, const int32_t *row_ptr,
                           const int32_t *col, const float *wgt, double beta, uint8_t *lab,
                           const int32_t *sites, const int32_t *level_ptr, int lv_lo, int lv_hi,
                           int grid_ctas, nemk_counters *cnt, const int32_t *skip);
void nemk_sweep_nem_jacobi(nemk_stream s, int k, int row0, int n_loc, const double *logpf,
                           const int32_t *row_ptr, const int32_t *col, const float *wgt, double beta,
                           const float *t_in, float *t_out, nemk_counters *cnt, const int32_t *skip);
void nemk_sweep_nem_level(nemk_stream s, int k, const double *logpf, const int32_t *row_ptr,
                          const int32_t *col, const float *wgt, double beta, float *t,
                          const int32_t *sites, const int32_t *level_ptr, int lv_lo, int lv_hi,
                          int grid_ctas, nemk_counters *cnt, const int32_t *skip);

/* ---- M-step */
/* lab_m (nullable): receives a copy of the labels = the state the statistics now describe */
void nemk_label_masks(nemk_stream s, int k, int n, int nwt, const uint8_t *lab, uint32_t *cm,
                      int32_t *nk_int, uint8_t *lab_m, const int32_t *halt);
/* incremental ncem statistics: rows whose label differs from lab_m (the labels S and n describe)
 * move their bits from S[old] to S[new] (exact integer updates).  list: n ints + 1 counter */
void nemk_mstep_delta(nemk_stream s, int k, int n, int d, int wpr, const uint32_t *x,
                      const uint8_t *lab, const uint8_t *lab_m, int32_t *list, int32_t *count,
                      int32_t *s_int, int32_t *nk_int, const int32_t *halt);
void nemk_mstep_ncem(nemk_stream s, int k, int d, int nwt, const uint32_t *xt, const uint32_t *cm,
                     int32_t *s_int, const int32_t *halt);
void nemk_mstep_nem(nemk_stream s, int k, int n, int d, int wpr, const uint32_t *x, const float *t,
                    int rows_per_chunk, double *partial_s, double *partial_n, double *s_dbl,
                    double *nk_dbl);
/* M-step closed forms (mu, eps, p from S and n) + the density tables of the new theta, one CTA
 * per class */
void nemk_mstep_finalize_tables(nemk_stream s, int k, int n, int d, int wpr, int prop_model,
                                int disp_model, const int32_t *s_int, const int32_t *nk_int,
                                const double *s_dbl, const double *nk_dbl, float *prop,
                                float *center, float *disp, nemk_coef *coef, uint32_t *mask_xor,
                                uint32_t *mask_valid, uint32_t *mask_f0, uint32_t *mask_f1,
                                double *delta, int force_mu_changed);

/* ---- criteria: per-rank partial sums (exactly nblocks rows of 4 doubles: D G L Z), then the
 * final fixed-order sum over the partial rows of every rank -> U D L M Z G */
int  nemk_criteria_partial(nemk_stream s, int k, int row0, int n_loc, nemk_lpsrc lps,
                           const int32_t *row_ptr, const int32_t *col, const float *wgt, double beta,
                           const uint8_t *lab, const float *t, const int32_t *heavy, int n_heavy,
                           double *partials, int nblocks);
void nemk_criteria_final(nemk_stream s, int nblocks_total, const double *partials, double beta,
                         double *crit6);
/* EstimBeta (nem_alg.c:2120-2230): same walk, partial rows = (crit, grad, dsec, 0); finish with
 * nemk_criteria_final(beta = NaN) which then stores the raw sums in crit6[0..3] */
int  nemk_betagrad_partial(nemk_stream s, int k, int row0, int n_loc, const int32_t *row_ptr,
                           const int32_t *col, const float *wgt, double beta, const uint8_t *lab,
                           const float *t, const int32_t *heavy, int n_heavy, double *partials,
                           int nblocks);

/* ---- row-sharded sweep, sparse label exchange (nem_kernels.cu "SPARSE label exchange"): a block is
 * 2 + 2*cap int32 words; blocks = the all-gathered [world] blocks.  After nemk_delta_apply
 * cnt->pending is the global number of cross-rank (reader, moved label) pairs, or -1 when some
 * rank moved more than cap labels (nothing was applied: do a full exchange instead);
 * cnt->changed_glob = labels changed over all ranks. */
void nemk_delta_pack(nemk_stream s, int row0, int n_loc, int cap, const uint8_t *lab_cur,
                     const uint8_t *seen, const nemk_counters *cnt, int32_t *block, const int32_t *skip);
void nemk_delta_apply(nemk_stream s, int world, int cap, const int32_t *blocks, int row0, int n_loc,
                      int shard_len, uint8_t *lab_cur, uint8_t *seen, int32_t *dirty, int32_t *wl,
                      int32_t *wl_count, const int32_t *rrow_ptr, const int32_t *rcol,
                      nemk_counters *cnt, const int32_t *skip);

/* ---- rank-ordered sums over an all-gathered stage [world][count] */
void nemk_sum_ranks_i32(nemk_stream s, int world, size_t count, const int32_t *stage, int32_t *out);
void nemk_sum_ranks_f64(nemk_stream s, int world, size_t count, const double *stage, double *out);

/* ---- resample driver (nem_sub_kernels.cu): genome subsample of the resident pangenome */
/* flag[n] = family has a selected genome; new_id[n+1] = exclusive scan; *n_eff = total */
void nemk_sub_active(nemk_stream s, int n, int wpr, const uint32_t *x, const uint32_t *mask,
                     int32_t *flag, int32_t *new_id, int32_t *block_tmp, int32_t *n_eff);
/* x_new[new_id[i]] = the d_eff selected columns (cols[], ascending) of row i, re-packed; index */
void nemk_sub_gather(nemk_stream s, int n, int wpr, int d_eff, int wpr_new, const uint32_t *x,
                     const uint32_t *mask, const int32_t *flag, const int32_t *new_id,
                     uint32_t *x_new, int32_t *index);
/* w_tmp[e] = popc(E_e & mask) (E = edge_bits, or x_i & x_j when NULL); cnt[n_cnt] kept entries per
 * NEW row; new_row_ptr[n_cnt+1] = their exclusive scan; *nnz_new, *maxdeg */
void nemk_sub_edges(nemk_stream s, int n, int wpr, const uint32_t *x, const uint32_t *mask,
                    const uint32_t *edge_bits, const int32_t *row_ptr, const int32_t *col,
                    const int32_t *flag, const int32_t *new_id, float *w_tmp, int32_t *cnt,
                    int32_t *new_row_ptr, int32_t *block_tmp, int32_t *nnz_new, int32_t *maxdeg,
                    int n_cnt);
void nemk_sub_fill(nemk_stream s, int n, const int32_t *row_ptr, const int32_t *col,
                   const float *w_tmp, const int32_t *flag, const int32_t *new_id,
                   const int32_t *new_row_ptr, int32_t *col_new, float *wgt_new);
/* votes[index[i]*4 + cls] += 1, cls = c_label (0 P, 1 S, 2 C) or 3 (U) */
void nemk_sub_vote(nemk_stream s, int n_eff, const int32_t *index, const uint8_t *lab, int c0,
                   int c1, int c2, int all_u, int32_t *votes);


/* ---- persistent EM kernel (nem_persist.cuh): ONE cooperative launch runs the whole ncem fit on the
 * popcount density path -- tables, densities, the two initial sweeps, then per EM iteration the
 * incremental (or full) M-step statistics, the closed forms + tables, the density pass, the
 * speculative sequential sweep (margin test, evaluation, fix-up rounds) and the convergence test
 * -- with a device-wide barrier between phases instead of a launch, and no host round trip until
 * the fit is over.  A problem too large for the in-kernel (L2-sized) X / X^T passes leaves the
 * kernel for exactly those two HBM-bound passes (exit codes below) and re-enters it. */
enum { NEMK_PK_ENTRY_INIT = 0,        /* tables of the theta in prop/center/disp, then the fit */
       NEMK_PK_ENTRY_INIT_SWEEPS = 1, /* densities ready (host ran the X pass): blind + beta sweep */
       NEMK_PK_ENTRY_MSTEP = 2,       /* start of an EM iteration */
       NEMK_PK_ENTRY_FINALIZE = 3,    /* statistics ready (host ran the full recount) */
       NEMK_PK_ENTRY_SWEEP = 4 };     /* densities of this iteration ready (host ran the X pass) */
enum { NEMK_PK_EXIT_DONE = 0, NEMK_PK_EXIT_NEED_DENSITY = 1, NEMK_PK_EXIT_NEED_RECOUNT = 2,
       NEMK_PK_EXIT_PEER_TIMEOUT = 3 };

#define NEMK_PK_MAX_WORLD 8

typedef struct {
    int32_t exit_code, resume_entry;    /* NEED_*: run the pass, re-enter at resume_entry */
    int32_t iters, converged, empty_class;
    int32_t cur, stale_par, last_changed, stats_valid, cnt_par, delta_mode, flags_stale, mu_changed;
    int32_t n_allnul, n_ties;           /* of the last sweep */
    uint32_t xepoch;                    /* cross-rank barrier epoch when the kernel was left */
    int32_t decide_pending, chg_local;
    int32_t xerror;                     /* a cross-rank barrier timed out (a peer died) */
    int32_t sweeps, x_passes, recounts, barriers;
    long long kept, fixup_rounds;       /* summed over the sweeps of this launch */
    /* nanoseconds CTA 0 spent in each phase, barrier waits included (globaltimer): 0 init (prep,
     * tables, first densities) 1 changed-rows scan 2 delta statistics 3 full recount 4 closed forms
     * + tables 5 X pass 6 margin test 7 evaluation of the active list 8 dense Jacobi round
     * 9 fix-up rounds 10 fix-up rounds executed (count) 11 spare */
    unsigned long long phase_ns[12];
    /* the same per EM iteration (first 12 iterations of the launch): [it][0..5] = ns of scan, delta
     * (or recount), finalize, margin test, evaluation, fix-up rounds; [it][6] = active sites of the
     * evaluation (-1: dense round), [it][7] = fix-up rounds */
    long long trace[12][8];
    /* final criteria U D L M Z G (ComputeCrit, nem_alg.c:2678-2757) when have_crit: computed by
     * the kernel itself when it is left for good */
    double crit[6];
    int32_t have_crit, pad1;
    unsigned long long seq;             /* written last, after a system-wide fence */
} nemk_persist_out;

typedef struct {
    /* problem (one GPU: row0 = 0, n = all families) */
    int32_t K, n, D, wpr, nwt;
    int32_t prop_model, disp_model, conv, it_max;
    float   conv_thr;
    double  beta;
    int32_t seq_sweep;        /* graph + beta != 0 + update=seq: speculative sequential sweep */
    int32_t use_graph;        /* graph + beta != 0 (update=para: one Jacobi round per sweep) */
    int32_t n_heavy, wsum_any_order, use_margins;
    int32_t x_in_kernel;      /* X and X^T passes may run inside the kernel (L2-sized problem) */
    int32_t init_from_pop;    /* ENTRY_INIT: every class of theta0 has a constant centre */
    /* entry state */
    int32_t entry, iter0, cur, stale_par, stats_valid, last_changed, margins_on, cnt_par, delta_mode, flags_stale, mu_changed;
    int32_t decide_pending, chg_local;   /* row shards: a sweep's convergence test waits for the ranks' sum */
    int32_t n_allnul, n_ties;   /* of the last sweep so far (carried over a re-entry) */
    unsigned long long seq;
    /* buffers (device) */
    const uint32_t *x, *xt;
    const int32_t *pop;
    const int32_t *row_ptr, *col, *rrow_ptr, *rcol, *heavy;
    const float *wgt;
    float *prop, *center, *disp;
    nemk_coef *coef;
    uint32_t *mxor, *mval, *f0, *f1, *cm;
    double *delta;
    int32_t *ham, *stat;      /* stat = S[K*D] then n[K] */
    uint8_t *lab[2], *stale[2];
    float *margin;
    int32_t *dirty, *wl[2], *wl_cnt;   /* wl_cnt[8]: [0..3] rotating fix-up counters, [4] changed rows */
    int32_t *hub_list;        /* unused (kept for layout) */
    int32_t *wlist[2];        /* fix-up work lists (duplicates allowed), wl_cap entries each */
    int32_t wl_cap, pad0;
    uint8_t *evflag;          /* [n] flags of every site's LAST evaluation: bit0 all-null, bit1 tie */
    int32_t *scratch;         /* [16] ints, zero at rest: [3] all-null rows [4] ties of the last sweep,
                                 [8..11] overflow flags of the rotating work lists */
    nemk_counters *cnt2;      /* [2] alternating per-sweep counter blocks, zero at entry */
    double *crit_partials;    /* [grid][4] per-CTA partial sums of the final criteria */
    int32_t want_crit, spatial;   /* evaluate the final criteria inside; the problem has a graph */
    int32_t no_shortcuts, pad2;   /* worst case: X pass and full X^T recount EVERY iteration (margin cache off too) */
    /* ---- row shards (world > 1): this rank owns the families [row0, row0 + n) of n_glob.  Labels,
     * Hamming counts, margins, flags, CSR and work lists are indexed by GLOBAL family id; x, xt, pop
     * and `stat` (this rank's S and n) by local row.  Every rank maps every rank's exchange block
     * (CUDA IPC, peer memory over NVLink): label moves and re-evaluation requests are stored
     * straight into the peers' blocks by the kernel, and the ranks meet at cross-rank barriers
     * (epoch flags in the blocks) -- one per M-step, one or two per sweep. */
    int32_t world, rank, row0, shard_len, n_glob, xcap;
    uint32_t xepoch, pad3;            /* cross-rank barrier epoch at entry */
    char *peer[NEMK_PK_MAX_WORLD];    /* exchange block of every rank (own included) */
    long long off_lab[2], off_stale[2], off_xflag, off_tot, off_incnt, off_inbox, off_stat, off_crit;
    int32_t stat_len, pad4;           /* ints a rank publishes per M-step: S[K*D], n[K], changed, all-null, ties, pad */
    int32_t *out_cnt;                 /* [world] local: requests queued for every rank in this super-round */
    int32_t *stat_glob;               /* [K*D + K] statistics summed over the ranks in rank order */
    unsigned *bar;            /* [2] device-wide barrier state (count, generation) */
    nemk_persist_out *out;    /* mapped pinned host memory */
} nemk_persist_args;

/* grid the cooperative launch may use for this K (CTAs), 0 = unsupported device */
int  nemk_persist_max_grid(int k);
/* value-preserving touch of every 4 KB of a peer's exchange block (fresh CUDA IPC mapping) */
void nemk_touch_peer(nemk_stream s, void *p, size_t bytes);
/* enqueue the kernel with `grid` CTAs (<= nemk_persist_max_grid) */
void nemk_persist_launch(nemk_stream s, const nemk_persist_args *a, int grid);

/* ---- helpers */
void nemk_labels_to_t(nemk_stream s, int k, int n, const uint8_t *lab, float *t);
void nemk_t_to_labels(nemk_stream s, int k, int n, const float *t, uint8_t *lab);
void nemk_fill_u8(nemk_stream s, uint8_t *p, int v, size_t n);
int  nemk_last_error(char *buf, int len);

#ifdef __cplusplus
}
#endif
#endif
