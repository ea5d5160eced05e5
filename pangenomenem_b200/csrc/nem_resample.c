/*
 * nem_resample.c -- the resample driver (include/nem_b200.h, layer 4).
 *
 * PPanGGOLiN runs NEM on organism subsets in two places: the chunk loop of partition() when there
 * are more than `chunck_size` organisms (ppanggolin.py:995-1105: sample 500 organisms, write the
 * five text files, run nem(), vote per family until every family is validated) and the
 * evolution-curve workers (command_line.py:262-281, 599-619).  Every sample re-serialises the
 * pangenome through __write_nem_input_files (ppanggolin.py:821-930).  Here the pangenome stays in
 * HBM: nemb_subsample() builds the subsample ON THE DEVICE into a second handle (kernels in
 * nem_sub_kernels.cu), the unchanged engine fits it, and nemb_resample_batch() runs many samples
 * over a few worker streams and accumulates the P/S/C/U votes on the device.
 */
#include "nem_handle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <sys/prctl.h>

static int round_up4(int v) { return (v + 3) / 4 * 4; }
static size_t carve(size_t *off, size_t bytes)
{
    size_t at = (*off + 255) & ~(size_t)255;
    *off = at + bytes;
    return at;
}

int nemb_subsample(nemb_handle *src, nemb_handle *h, const uint32_t *genome_mask,
                   const uint32_t *edge_presence_dev, int *n_eff_out, int *d_eff_out)
{
    if (!src || !h || !genome_mask) return NEMB_E_ARG;
    if (src == h) return fail(h, NEMB_E_ARG, "nemb_subsample: source and destination must differ");
    if (!src->loaded) return fail(h, NEMB_E_ARG, "nemb_subsample: no pangenome loaded in the source");
    if (src->world > 1 || h->world > 1) return fail(h, NEMB_E_ARG, "nemb_subsample: single-GPU handles only");
    if (src->device != h->device) return fail(h, NEMB_E_ARG, "nemb_subsample: handles on different devices");
    if (src->spatial && !src->symmetric)
        return fail(h, NEMB_E_ARG, "nemb_subsample: the neighbour graph must be symmetric");
    CK(cudaSetDevice(h->device));
    nemb_i_reset_problem(h);
    const int n = src->n, D = src->d, wpr = src->wpr, nnz = src->spatial ? src->nnz : 0;
    const int wm = (D + 31) / 32;

    /* host: padded mask (the ascending list of the selected genomes is implicit in it) */
    uint32_t *mask = calloc((size_t)wpr, sizeof(uint32_t));
    if (!mask) return fail(h, NEMB_E_MEMORY, "host alloc");
    int d_eff = 0;
    for (int w = 0; w < wm; w++) {
        uint32_t m = genome_mask[w];
        if (w == wm - 1 && (D & 31)) m &= (1u << (D & 31)) - 1u;
        mask[w] = m;
        d_eff += __builtin_popcount(m);
    }
    if (d_eff == 0) { free(mask); return fail(h, NEMB_E_ARG, "nemb_subsample: empty genome mask"); }
    const int wpr_new = round_up4((d_eff + 31) / 32);

    /* scratch (sized by the SOURCE so that a worker handle never reallocates between samples) */
    const int nb = (n + 1023) / 1024 + 2;
    size_t off = 0;
    size_t o_mask = carve(&off, sizeof(uint32_t) * wpr);
    size_t o_flag = carve(&off, sizeof(int32_t) * n), o_id = carve(&off, sizeof(int32_t) * ((size_t)n + 1));
    size_t o_cnt = carve(&off, sizeof(int32_t) * ((size_t)n + 1)), o_blk = carve(&off, sizeof(int32_t) * nb);
    size_t o_tot = carve(&off, sizeof(int32_t) * 8), o_w = carve(&off, sizeof(float) * (size_t)(nnz ? nnz : 1));
    int rc;
#define BAIL(code) do { free(mask); return (code); } while (0)
    if ((rc = nemb_i_reserve(h, &h->b_sub, off)) != NEMB_OK) BAIL(rc);
    char *base = h->b_sub.p;
    uint32_t *d_mask = (uint32_t *)(base + o_mask);
    int32_t *d_flag = (int32_t *)(base + o_flag);
    int32_t *d_id = (int32_t *)(base + o_id), *d_cnt = (int32_t *)(base + o_cnt);
    int32_t *d_blk = (int32_t *)(base + o_blk), *d_tot = (int32_t *)(base + o_tot);
    float *d_w = (float *)(base + o_w);
    if ((rc = nemb_i_reserve(h, &h->b_x, sizeof(uint32_t) * (size_t)n * round_up4(wm))) != NEMB_OK) BAIL(rc);
    if ((rc = nemb_i_reserve(h, &h->b_index, sizeof(int32_t) * (size_t)n)) != NEMB_OK) BAIL(rc);
    if (src->spatial) {
        if ((rc = nemb_i_reserve(h, &h->b_row_ptr, sizeof(int32_t) * ((size_t)n + 2))) != NEMB_OK) BAIL(rc);
        if ((rc = nemb_i_reserve(h, &h->b_col, sizeof(int32_t) * (size_t)(nnz ? nnz : 1))) != NEMB_OK) BAIL(rc);
        if ((rc = nemb_i_reserve(h, &h->b_wgt, sizeof(float) * (size_t)(nnz ? nnz : 1))) != NEMB_OK) BAIL(rc);
        size_t hl_blocks = ((size_t)n + 1023) / 1024 + 1;
        if ((rc = nemb_i_reserve(h, &h->b_heavy, sizeof(int32_t) * (hl_blocks + (size_t)n + 1))) != NEMB_OK) BAIL(rc);
        if ((rc = nemb_i_reserve(h, &h->b_flags, 64)) != NEMB_OK) BAIL(rc);
    }
    /* (pageable source: the runtime stages it before the call returns, so it can be freed at once) */
    cudaError_t ce = cudaMemcpyAsync(d_mask, mask, sizeof(uint32_t) * wpr, cudaMemcpyHostToDevice, h->stream);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(d_tot, 0, sizeof(int32_t) * 8, h->stream);
    free(mask);
#undef BAIL
    if (ce != cudaSuccess) return fail(h, NEMB_E_CUDA, "nemb_subsample upload: %s", cudaGetErrorString(ce));

    h->d_x = h->b_x.p;
    h->x_owned = 1;
    h->d_index = h->b_index.p;
    nemk_sub_active(h->stream, n, wpr, src->d_x, d_mask, d_flag, d_id, d_blk, &d_tot[0]);
    nemk_sub_gather(h->stream, n, wpr, d_eff, wpr_new, src->d_x, d_mask, d_flag, d_id, h->d_x, h->d_index);
    if (src->spatial) {
        nemk_sub_edges(h->stream, n, wpr, src->d_x, d_mask, edge_presence_dev, src->d_row_ptr, src->d_col,
                       d_flag, d_id, d_w, d_cnt, h->b_row_ptr.p, d_blk, &d_tot[1], &d_tot[2], n);
        nemk_sub_fill(h->stream, n, src->d_row_ptr, src->d_col, d_w, d_flag, d_id, h->b_row_ptr.p,
                      h->b_col.p, h->b_wgt.p);
    }
    if (src->spatial) {
        /* hub list of the new graph, over the OLD row count (the rows beyond n_eff have no entry):
         * its size comes back with the totals in one synchronisation */
        size_t hl_blocks = ((size_t)n + 1023) / 1024 + 1;
        nemk_heavy_list(h->stream, 0, n, h->b_row_ptr.p, h->b_heavy.p, (int32_t *)h->b_heavy.p + hl_blocks,
                        &d_tot[3]);
    }
    CKK();
    int32_t tot[4] = {0, 0, 0, 0};
    CK(cudaMemcpyAsync(tot, d_tot, sizeof tot, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const int n_eff = tot[0];
    if (n_eff <= 0) return fail(h, NEMB_E_ARG, "nemb_subsample: no family has a selected genome");

    h->n = h->n_glob = n_eff; h->row0 = 0; h->shard_len = n_eff; h->lab_len = n_eff;
    h->d = d_eff; h->wpr = wpr_new;
    h->nwt = round_up4((n_eff + 31) / 32);
    if (h->nwt < 4) h->nwt = 4;
    h->spatial = src->spatial;
    h->nnz = 0; h->symmetric = 1; h->max_neigh = 0; h->n_heavy = 0; h->d_heavy = NULL;
    h->d_row_ptr = h->d_col = h->d_rrow_ptr = h->d_rcol = NULL; h->d_wgt = NULL;
    if (h->spatial) {
        h->nnz = tot[1]; h->max_neigh = tot[2];
        h->wgt_integral = !getenv("NEM_B200_ORDERED_SUMS");   /* coverages are popcounts */
        h->d_row_ptr = h->b_row_ptr.p; h->d_col = h->b_col.p; h->d_wgt = h->b_wgt.p;
        h->d_rrow_ptr = h->d_row_ptr; h->d_rcol = h->d_col;
        size_t hl_blocks = ((size_t)n + 1023) / 1024 + 1;
        h->d_heavy = (int32_t *)h->b_heavy.p + hl_blocks;
        h->n_heavy = tot[3];
    }
    h->loaded = 1;
    if (n_eff_out) *n_eff_out = n_eff;
    if (d_eff_out) *d_eff_out = d_eff;
    return NEMB_OK;
}

int nemb_get_family_index(nemb_handle *h, int32_t *index_out)
{
    if (!h || !index_out) return NEMB_E_ARG;
    if (!h->loaded || !h->d_index) return fail(h, NEMB_E_ARG, "not a device-built subsample");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(index_out, h->d_index, sizeof(int32_t) * (size_t)h->n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return NEMB_OK;
}

/* PPanGGOLiN's default .m (ppanggolin.py:893-901) as ReadParamFile stores it (nem_exe.c:1022-1034) */
static void default_theta3(int d, float *prop, float *center, float *disp)
{
    prop[0] = 0.33333f; prop[1] = 0.33333f;
    prop[2] = (1.0f - prop[0]) - prop[1];
    for (int j = 0; j < d; j++) {
        center[j] = 1.0f; center[d + j] = 0.5f; center[2 * d + j] = 0.0f;
        disp[j] = 0.1f; disp[d + j] = 0.5f; disp[2 * d + j] = 0.1f;
    }
}

/* run_partitioning's class -> P/S/C map (ppanggolin.py:1925-1957): persistent = largest number of
 * non-zero centres, shell = largest summed dispersion, and the result must be (0,1,2) = (P,S,C),
 * otherwise every family of the sample is "undefined" */
static int psc_consistent(int d, const float *center, const float *disp)
{
    int best_mu = -1, pk = 0, sk = 0;
    double best_eps = -1.0;
    for (int k = 0; k < 3; k++) {
        int nz = 0;
        double se = 0.0;
        for (int j = 0; j < d; j++) { nz += center[k * d + j] != 0.0f; se += (double)disp[k * d + j]; }
        if (nz > best_mu) { best_mu = nz; pk = k; }
        if (se > best_eps) { best_eps = se; sk = k; }
    }
    return pk == 0 && sk == 1;
}

typedef struct {
    nemb_handle *src;
    int n_runs, wm, failed_rc, relaxed;
    const uint32_t *masks;
    const float *betas;
    const nemb_options *opt;
    const uint32_t *edge_bits;
    int32_t *d_votes;
    int32_t *iters_out;
    int next;                 /* atomic run counter */
    int pk_grid_limit;        /* CTAs a worker's persistent fit kernel may use */
    int fit_slots, build_slots;   /* fits / subsample builds in flight at once (0 = no limit) */
    int fits_now, builds_now;
    pthread_mutex_t gate_mu;
    pthread_cond_t gate_cv;
    pthread_mutex_t mu;
    nemb_batch_stats st;
    char err[256];
} batch_ctx;

/* counting gates: how many fits (persistent cooperative kernels that spin at device-wide barriers)
 * and how many builders (bandwidth-bound grids of thousands of CTAs) share the GPU at a time */
static void gate_enter(batch_ctx *c, int *now, int slots)
{
    if (slots <= 0) return;
    pthread_mutex_lock(&c->gate_mu);
    while (*now >= slots) pthread_cond_wait(&c->gate_cv, &c->gate_mu);
    (*now)++;
    pthread_mutex_unlock(&c->gate_mu);
}
static void gate_leave(batch_ctx *c, int *now, int slots)
{
    if (slots <= 0) return;
    pthread_mutex_lock(&c->gate_mu);
    (*now)--;
    pthread_cond_broadcast(&c->gate_cv);
    pthread_mutex_unlock(&c->gate_mu);
}

static void *batch_worker(void *arg)
{
    batch_ctx *c = arg;
    nemb_handle *h = NULL;
    int rc = nemb_create(&h, c->src->device);
    if (rc == NEMB_OK && c->relaxed) { h->poll_relaxed = 1; prctl(PR_SET_TIMERSLACK, 1000UL, 0, 0, 0); }
    /* the fits of concurrent workers are persistent cooperative kernels (one launch per fit): each
     * takes its share of the SMs so that they run side by side instead of one after the other */
    if (rc == NEMB_OK) h->pk_grid_limit = c->pk_grid_limit;
    float *theta = NULL;
    int cap_d = 0;
    nemb_batch_stats loc;
    memset(&loc, 0, sizeof loc);
    while (rc == NEMB_OK) {
        int r = __atomic_fetch_add(&c->next, 1, __ATOMIC_RELAXED);
        if (r >= c->n_runs) break;
        int n_eff = 0, d_eff = 0;
        gate_enter(c, &c->builds_now, c->build_slots);
        rc = nemb_subsample(c->src, h, c->masks + (size_t)r * c->wm, c->edge_bits, &n_eff, &d_eff);
        gate_leave(c, &c->builds_now, c->build_slots);
        if (rc != NEMB_OK) break;
        if (d_eff > cap_d) {
            free(theta);
            theta = malloc(sizeof(float) * (3 + 6 * (size_t)d_eff));
            cap_d = d_eff;
            if (!theta) { rc = NEMB_E_MEMORY; break; }
        }
        float *prop = theta, *center = theta + 3, *disp = theta + 3 + 3 * (size_t)d_eff;
        default_theta3(d_eff, prop, center, disp);
        nemb_options o = *c->opt;
        if (c->betas) o.beta = c->betas[r];
        o.profile = 0; o.dolog = 0;
        nemb_result res;
        gate_enter(c, &c->fits_now, c->fit_slots);
        int frc = nemb_fit(h, &o, prop, center, disp, &res);
        gate_leave(c, &c->fits_now, c->fit_slots);
        int all_u = 0;
        if (frc == NEMB_W_EMPTYCLASS) { all_u = 1; loc.n_failed++; }   /* no .uf => all undefined */
        else if (frc != NEMB_OK) { rc = frc; break; }
        else if (!psc_consistent(d_eff, center, disp)) { all_u = 1; loc.n_inconsistent++; }
        else loc.n_ok++;
        nemk_sub_vote(h->stream, n_eff, h->d_index, h->d_lab[h->cur], 0, 1, 2, all_u, c->d_votes);
        loc.n_runs++;
        loc.family_iterations += (int64_t)n_eff * res.iters;
        loc.kernel_launches += res.kernel_launches + 8;
        loc.fit_ms_sum += res.fit_ms;
        if (c->iters_out) c->iters_out[r] = res.iters;
    }
    if (h) { cudaStreamSynchronize(h->stream); }
    pthread_mutex_lock(&c->mu);
    if (rc != NEMB_OK && c->failed_rc == NEMB_OK) {
        c->failed_rc = rc;
        snprintf(c->err, sizeof c->err, "%s", h ? nemb_last_error(h) : "nemb_create failed");
    }
    c->st.n_runs += loc.n_runs; c->st.n_ok += loc.n_ok; c->st.n_inconsistent += loc.n_inconsistent;
    c->st.n_failed += loc.n_failed; c->st.family_iterations += loc.family_iterations;
    c->st.kernel_launches += loc.kernel_launches; c->st.fit_ms_sum += loc.fit_ms_sum;
    pthread_mutex_unlock(&c->mu);
    free(theta);
    if (h) nemb_destroy(h);
    return NULL;
}

int nemb_resample_batch(nemb_handle *src, int n_runs, const uint32_t *genome_masks, const float *betas,
                        const nemb_options *opt, int n_workers, const uint32_t *edge_presence_dev,
                        int32_t *votes_out, int32_t *iters_out, nemb_batch_stats *stats)
{
    if (!src || !genome_masks || !opt || n_runs < 0) return NEMB_E_ARG;
    nemb_handle *h = src;
    if (!src->loaded) return fail(h, NEMB_E_ARG, "nemb_resample_batch: no pangenome loaded");
    if (opt->k != 3) return fail(h, NEMB_E_ARG, "nemb_resample_batch: K must be 3 (persistent/shell/cloud)");
    if (opt->algo != NEMB_ALGO_NCEM) return fail(h, NEMB_E_ARG, "nemb_resample_batch: algo must be ncem");
    if (n_workers < 1) n_workers = 8;
    if (n_workers > 32) n_workers = 32;
    long cores = sysconf(_SC_NPROCESSORS_ONLN);   /* every worker polls its status slot: leave cores free */
    if (cores > 0 && n_workers > cores) n_workers = (int)cores;
    /* more pollers than spare cores (several ranks per node count too: NEM_B200_POLL=relaxed) =>
     * the workers sleep ~2 us between status probes instead of spinning */
    const char *pe = getenv("NEM_B200_POLL");
    int relaxed = pe ? !strcmp(pe, "relaxed") : (cores > 0 && n_workers > cores / 2);
    if (n_workers > n_runs && n_runs > 0) n_workers = n_runs;
    CK(cudaSetDevice(src->device));
    CK(cudaStreamSynchronize(src->stream));   /* the source is read-only from here on */
    batch_ctx c;
    memset(&c, 0, sizeof c);
    c.src = src; c.n_runs = n_runs; c.wm = (src->d + 31) / 32; c.masks = genome_masks; c.betas = betas;
    c.opt = opt; c.edge_bits = edge_presence_dev; c.relaxed = relaxed; c.iters_out = iters_out; c.failed_rc = NEMB_OK;
    {
        /* fits in flight, builders in flight, CTAs per fit: a fit is latency-bound (device-wide
         * barriers) and gains from running beside others on a share of the SMs; every resident fit
         * CTA pins half an SM's registers while it spins, so the fits together get at most about
         * half of the CTA slots and the builders of the other workers the rest */
        const char *e = getenv("NEM_B200_BATCH_FITS");
        int total = nemk_persist_max_grid(opt->k);
        c.fit_slots = e && *e ? atoi(e) : (n_workers > 1 ? 4 : 0);
        e = getenv("NEM_B200_BATCH_BUILDS");
        c.build_slots = e && *e ? atoi(e) : 0;
        int conc = c.fit_slots > 0 && c.fit_slots < n_workers ? c.fit_slots : n_workers;
        e = getenv("NEM_B200_BATCH_GRID");      /* CTAs per worker fit (0 = all) */
        c.pk_grid_limit = e && *e ? atoi(e) : (conc > 1 && total > 0 ? (total / (2 * conc) < 16 ? 16 : total / (2 * conc)) : 0);
    }
    pthread_mutex_init(&c.gate_mu, NULL);
    pthread_cond_init(&c.gate_cv, NULL);
    pthread_mutex_init(&c.mu, NULL);
    size_t vbytes = sizeof(int32_t) * 4 * (size_t)src->n;
    CK(cudaMalloc((void **)&c.d_votes, vbytes));
    CK(cudaMemset(c.d_votes, 0, vbytes));
    pthread_t th[32];
    int started = 0;
    for (int w = 0; w < n_workers && n_runs > 0; w++)
        if (pthread_create(&th[started], NULL, batch_worker, &c) == 0) started++;
    for (int w = 0; w < started; w++) pthread_join(th[w], NULL);
    pthread_mutex_destroy(&c.mu);
    pthread_mutex_destroy(&c.gate_mu);
    pthread_cond_destroy(&c.gate_cv);
    int rc = c.failed_rc;
    if (n_runs > 0 && started == 0) rc = fail(h, NEMB_E_BUG, "could not start a worker thread");
    else if (rc != NEMB_OK) fail(h, rc, "resample worker: %s", c.err);
    cudaSetDevice(src->device);
    if (rc == NEMB_OK && votes_out) {
        cudaError_t e = cudaMemcpy(votes_out, c.d_votes, vbytes, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(h, NEMB_E_CUDA, "votes copy: %s", cudaGetErrorString(e));
    }
    cudaFree(c.d_votes);
    if (stats) *stats = c.st;
    return rc;
}
