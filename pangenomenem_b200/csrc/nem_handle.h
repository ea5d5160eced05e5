/*
 * nem_handle.h -- internal: the engine handle shared by the C host files (nem_fit.c: buffers, EM
 * control flow; nem_resample.c: device-side subsample builder + resample driver).  Not installed.
 */
#ifndef NEM_HANDLE_H
#define NEM_HANDLE_H
#include "nem_b200.h"
#include "nem_device.h"

#include <cuda_runtime_api.h>
#include <stddef.h>
#include <stdint.h>

#define NEMB_VERSION "nem-b200 0.2 (NEM 1.08-a compatible)"
#define NARROW_LEVEL 2048   /* levels at most this wide are walked by one CTA */
#define MAX_LEVEL_GRID 2368 /* 148 SMs x 16 CTAs of 256 threads */
#define MAX_WORLD 64
#define CRIT_BLOCKS_MAX 1184 /* 148 SMs x 8 */
#define CRIT_BLOCKS_SHARD 296
#define DELTA_CAP 2048       /* labels a rank may publish per sparse exchange */
#define RING 4              /* status slots in mapped host memory */

typedef struct { int lo, hi, grid; } sweep_step;

typedef nemk_iter_status iter_status;

enum { ST_DENSITY = 0, ST_SWEEP = 1, ST_MSTEP = 2, ST_CRIT = 3, ST_DENSITY_CACHED = 4, ST_MSTEP_DELTA = 5, ST_NB = 6 };

typedef struct { void *p; size_t cap; } dbuf;   /* grow-only device buffer */

struct nemb_handle {
    int device;
    cudaStream_t stream;
    int own_stream;
    char err[512];
    nemb_comm *comm;
    int rank, world;
    /* problem.  n = rows this rank owns, n_glob = families of the whole pangenome, row0 = global
     * id of local row 0, shard_len = rows per rank slot (lab_len = world * shard_len >= n_glob) */
    int n, n_glob, row0, shard_len, lab_len;
    int d, wpr, nwt, nnz, spatial, symmetric, max_neigh, loaded;
    int wgt_integral;      /* every edge weight is an integer of magnitude <= 2^20 (exact sums) */
    uint32_t *d_x, *d_xt;
    int x_owned, have_xt;
    dbuf b_x, b_xt, b_row_ptr, b_col, b_wgt, b_rrow_ptr, b_rcol, b_sites, b_level_ptr, b_flags, b_heavy;
    dbuf b_pop;            /* row popcounts of X (lazy, or taken behind the chunked upload) */
    int32_t *d_pop;
    int have_pop;
    dbuf b_sub, b_index;   /* resample driver: builder scratch; original family id of every row */
    int32_t *d_index;      /* non-NULL when this problem is a device-built subsample */
    cudaStream_t copy_stream;   /* loader: chunked upload of X, overlapped with the transposes */
    cudaEvent_t copy_ev[16];
    uint8_t *h_lab_stage;       /* pinned staging of the labels read-back */
    size_t h_lab_cap;
    int poll_relaxed;      /* status polling sleeps between probes (resample workers share the cores) */
    int32_t *d_heavy;      /* index-sorted hubs among this rank's rows */
    int n_heavy;
    int32_t *d_row_ptr, *d_col, *d_rrow_ptr, *d_rcol, *d_sites, *d_level_ptr;
    float *d_wgt;
    int have_levels, depth;
    int32_t *h_level;      /* host: level of every site (lazy) */
    sweep_step *steps;
    int n_steps;
    /* per-K buffers: one slab */
    int k_alloc;
    dbuf b_slab, b_t[2], b_nem;
    float *d_prop, *d_center, *d_disp, *d_iner;
    nemk_coef *d_coef;
    uint32_t *d_mxor, *d_mval, *d_f0, *d_f1;
    double *d_delta;
    double *d_logpf;
    uint8_t *d_lab[4];     /* two sweep buffers, the labels last seen from remote ranks, and the
                              labels the M-step statistics currently describe (local rows) */
    int32_t *d_ham;        /* cached Hamming counts H[n][K] of the popcount density path */
    int ham_valid, stats_valid, tables_forced;
    float *d_margin;       /* margin cache of the dense sweep (nemk_margins), local rows */
    uint8_t *d_stale[2];   /* stale flags, swapped every speculative sweep */
    int stale_par, sweep_same_beta;
    nemk_margins mg;       /* the margins of the sweep being enqueued (m == NULL: off) */
    int32_t *d_xchg, *d_xchg_all;     /* row shards: sparse label-exchange blocks (own, all ranks) */
    int prev_valid;        /* d_lab[cur ^ 1] holds the input labels of the last sweep (it flipped) */
    int lp_from_ham;       /* the consumers rebuild logpf from d_ham in registers (no logpf array) */
    int64_t last_changed;  /* labels moved by the last sweep (all ranks), -1 = unknown */
    float *d_t[2];
    int cur, state_labels;
    int32_t *d_dirty, *d_wl[2], *d_wl_counts;
    uint32_t *d_cm;
    int32_t *d_stat_loc;                      /* this rank's S[K*D] then n[K] (kept incrementally) */
    int32_t *d_stat_int, *d_stat_int_stage;   /* all ranks' sum; stage = [world][K*D+K] */
    double *d_stat_dbl, *d_stat_dbl_stage;
    double *d_partial_s, *d_partial_n;
    int rows_per_chunk, nchunks;
    double *d_crit_partials;
    int crit_blocks;
    iter_status *d_status, *h_status;
    nemk_counters *d_cnt_all, *h_cnt_all;     /* [world] gathered sweep counters */
    int32_t *h_empty;
    nemk_host_status *ring, *d_ring;          /* mapped pinned status slots, host / device view */
    unsigned long long seq;                   /* last sequence number handed to nemk_iter_end */
    /* persistent EM kernel: scratch (hub list, counters, barrier), mapped status block */
    dbuf b_pk;
    int32_t *d_pk_hub, *d_pk_scratch, *d_pk_wl[2];
    uint8_t *d_pk_evflag;
    int pk_wl_cap, sm_khz;
    nemk_counters *d_pk_cnt2;
    unsigned *d_pk_bar;
    double *d_pk_crit;          /* per-CTA partial sums of the in-kernel criteria */
    int pk_have_crit;           /* the last persistent fit evaluated its final criteria itself */
    double pk_crit[6];
    float *h_theta_stage;       /* pinned staging of theta (one copy each way) */
    size_t h_theta_cap;
    nemk_persist_out *pk_out, *d_pk_out;
    unsigned long long pk_seq;
    int pk_cnt_par, pk_ready_n, pk_ready_heavy;
    long long pk_trace[12][8];            /* per-iteration trace of the LAST launch of the last fit */
    unsigned long long pk_phase_ns[12];   /* of the last fit (nemk_persist_out.phase_ns, summed over its launches) */
    int persist_local, xpeer_local; /* debugging: peer-memory kernel between the ranks of the in-process test communicator */
    int pk_shard_max_world; /* row shards: largest world size served by the peer-memory kernel */
    int pk_grid_limit;     /* > 0: CTAs a fit of this handle may use (concurrent fits share the GPU) */
    /* environment knobs (tests / A-B runs), read ONCE per fit by read_env_knobs() instead of a
     * getenv() per sweep */
    int no_persist, keep_logpf, no_margins, no_popcache, no_spec, full_mstep, full_exchange;
    int medium_list, pk_grid_env;
    int no_shortcuts;      /* NEM_B200_NO_SHORTCUTS: X pass + full recount every iteration, margin cache off */
    size_t pk_xlimit;      /* X up to this many bytes: the X / X^T passes run inside the persistent kernel */
    /* row shards, persistent kernel: this rank's exchange block (labels, stale flags, barrier flags,
     * inboxes, staging areas) and the peers' blocks mapped through CUDA IPC */
    dbuf b_xblk;
    size_t xblk_bytes;
    int xblk_k, xblk_lab_len, xblk_ok;      /* what the mapped blocks were laid out for; usable */
    char *xpeer[NEMK_PK_MAX_WORLD];
    long long xoff[12];                     /* off_lab[2], off_stale[2], xflag, tot, incnt, inbox, stat, crit */
    int xstat_len, xcap;
    unsigned pk_xepoch;
    int32_t *d_pk_out_cnt;
    int no_persist_shard;
    /* fit bookkeeping */
    int64_t launches, fixup_rounds, exchanges;
    int profile;
    cudaEvent_t *ev;
    int *ev_kind;
    int ev_cap, ev_n, ev_last_density, ev_last_cached;
};

#define CK(call)                                                                           \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return fail(h, e_ == cudaErrorMemoryAllocation ? NEMB_E_MEMORY : NEMB_E_CUDA,  \
                        "%s: %s", #call, cudaGetErrorString(e_));                          \
    } while (0)

#define CKK()                                                                  \
    do {                                                                       \
        char b_[256];                                                          \
        if (nemk_last_error(b_, sizeof b_))                                    \
            return fail(h, NEMB_E_CUDA, "kernel launch failed: %s", b_);       \
    } while (0)


/* helpers defined in nem_fit.c */
int  nemb_i_fail(nemb_handle *h, int code, const char *fmt, ...);
int  nemb_i_reserve(nemb_handle *h, dbuf *b, size_t bytes);
void nemb_i_reset_problem(nemb_handle *h);
#define fail nemb_i_fail

#endif
