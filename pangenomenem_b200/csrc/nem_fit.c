/*
 * nem_fit.c -- C host of the in-memory API (include/nem_b200.h, layer 3): device buffers, the
 * loader (CSR upload, device-side validation, lazy sweep level schedule), the EM control flow
 * for one GPU and for row shards, and the stage entry points.  All numerics run in the CUDA
 * kernels of nem_kernels.cu through nem_device.h; this file only sequences them.  No CPU
 * fallback exists: every entry point needs a CUDA device.
 *
 * Control flow restated from the reference (root ppanggolin/NEM):
 *   ClassifyByNemOneBeta, INIT_PARAM_FILE branch   nem_alg.c:1151-1169
 *   ComputePartitionFromPara                       nem_alg.c:1951-1989
 *   NemAlgo                                        nem_alg.c:1746-1879
 *   HasConverged                                   nem_alg.c:2056-2112
 *   RandNemAlgo / MakeRandomPara / InitPara        nem_alg.c:1574-1742, 1381-1473, 1200-1280
 */
#include "nem_handle.h"

#include <cuda_runtime_api.h>
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

/* ------------------------------------------------------------------ errors */
int nemb_i_fail(nemb_handle *h, int code, const char *fmt, ...)
{
    if (h) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(h->err, sizeof h->err, fmt, ap);
        va_end(ap);
    }
    return code;
}

const char *nemb_version(void) { return NEMB_VERSION; }
const char *nemb_last_error(const nemb_handle *h) { return h ? h->err : "null handle"; }

static int reserve(nemb_handle *h, dbuf *b, size_t bytes)
{
    if (bytes < 256) bytes = 256;
    if (b->cap >= bytes) return NEMB_OK;
    if (b->p) { cudaFree(b->p); b->p = NULL; b->cap = 0; }
    size_t want = bytes + bytes / 16;   /* a little slack: reloads of slightly larger problems */
    cudaError_t e = cudaMalloc(&b->p, want);
    if (e != cudaSuccess) { cudaGetLastError(); want = bytes; e = cudaMalloc(&b->p, want); }
    if (e != cudaSuccess)
        return fail(h, e == cudaErrorMemoryAllocation ? NEMB_E_MEMORY : NEMB_E_CUDA,
                    "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
    b->cap = want;
    return NEMB_OK;
}
int nemb_i_reserve(nemb_handle *h, dbuf *b, size_t bytes) { return reserve(h, b, bytes); }
static void release(dbuf *b) { if (b->p) cudaFree(b->p); b->p = NULL; b->cap = 0; }

/* ------------------------------------------------------------------ lifetime */
/* pid of the process in which THIS address space first touched CUDA (0: never).  A forked child
 * inherits it, and with it a CUDA state it cannot use (CUDA does not survive fork): nem() then
 * serves the child through a helper process (nem_api.c). */
static pid_t g_cuda_pid;
int nemb_i_cuda_owner_pid(void) { return (int)g_cuda_pid; }

int nemb_create(nemb_handle **out, int device)
{
    if (!out) return NEMB_E_ARG;
    *out = NULL;
    if (g_cuda_pid && g_cuda_pid != getpid()) {
        fprintf(stderr, "nem_b200: this process was forked from one that had already initialised CUDA; "
                        "a CUDA context does not survive fork() -- use nem() (it serves forked callers "
                        "through a helper process), or fork before the first engine call\n");
        return NEMB_E_CUDA;
    }
    if (!g_cuda_pid) g_cuda_pid = getpid();
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        fprintf(stderr, "nem_b200: no CUDA device available (%s); this engine has no CPU path\n",
                e != cudaSuccess ? cudaGetErrorString(e) : "0 devices");
        return NEMB_E_CUDA;
    }
    if (device < 0) {
        const char *env = getenv("NEM_B200_DEVICE");
        if (env && *env) device = atoi(env);
        else if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= count) device = device % count;
    nemb_handle *h = calloc(1, sizeof *h);
    if (!h) return NEMB_E_MEMORY;
    h->device = device;
    h->world = 1;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        free(h);
        return NEMB_E_CUDA;
    }
    h->own_stream = 1;
    *out = h;
    return NEMB_OK;
}

int nemb_set_stream(nemb_handle *h, void *cuda_stream)
{
    if (!h) return NEMB_E_ARG;
    if (h->own_stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    h->stream = (cudaStream_t)cuda_stream;
    h->own_stream = 0;
    return NEMB_OK;
}

int nemb_set_comm(nemb_handle *h, nemb_comm *comm)
{
    if (!h) return NEMB_E_ARG;
    if (comm && (comm->world < 1 || comm->world > MAX_WORLD || comm->rank < 0 ||
                 comm->rank >= comm->world || !comm->allgather))
        return fail(h, NEMB_E_ARG, "bad communicator (rank %d of %d)", comm->rank, comm->world);
    h->comm = comm;
    h->rank = comm ? comm->rank : 0;
    h->world = comm ? comm->world : 1;
    h->loaded = 0;
    h->k_alloc = 0;
    return NEMB_OK;
}

void nemb_shard_range(int n_glob, int world, int rank, int *shard_len, int *row0, int *n_loc)
{
    if (world < 1) world = 1;
    /* ceil(N / world) rounded up to 16 families when there are several ranks: a shard then starts on
     * a 16-byte boundary of the label arrays (vector accesses of the persistent kernel) */
    int sl = (n_glob + world - 1) / world;
    if (world > 1 && sl >= 1024) sl = (sl + 15) / 16 * 16;      /* (tiny pangenomes keep ceil(N / world)) */
    long long r0 = (long long)rank * sl;
    int nl = r0 >= n_glob ? 0 : (int)((long long)n_glob - r0 < sl ? (long long)n_glob - r0 : sl);
    if (shard_len) *shard_len = sl;
    if (row0) *row0 = (int)(r0 > n_glob ? n_glob : r0);
    if (n_loc) *n_loc = nl;
}

/* logical reset; device buffers are kept for the next load (grow-only) */
static void reset_problem(nemb_handle *h)
{
    if (!h->x_owned) h->d_x = NULL;
    free(h->h_level); free(h->steps);
    h->h_level = NULL; h->steps = NULL;
    h->n_steps = 0; h->loaded = 0; h->have_xt = 0; h->have_levels = 0; h->depth = 0;
    h->d_index = NULL; h->have_pop = 0;
    h->k_alloc = 0;
}

void nemb_i_reset_problem(nemb_handle *h) { reset_problem(h); }

void nemb_destroy(nemb_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    reset_problem(h);
    dbuf *all[] = {&h->b_x, &h->b_xt, &h->b_row_ptr, &h->b_col, &h->b_wgt, &h->b_rrow_ptr,
                   &h->b_rcol, &h->b_sites, &h->b_level_ptr, &h->b_flags, &h->b_heavy, &h->b_sub, &h->b_index, &h->b_pop,
                   &h->b_slab, &h->b_t[0],
                   &h->b_t[1], &h->b_nem, &h->b_pk};
    for (size_t i = 0; i < sizeof all / sizeof all[0]; i++) release(all[i]);
    if (h->h_status) cudaFreeHost(h->h_status);
    if (h->h_cnt_all) cudaFreeHost(h->h_cnt_all);
    if (h->h_empty) cudaFreeHost(h->h_empty);
    if (h->ring) cudaFreeHost(h->ring);
    if (h->pk_out) cudaFreeHost(h->pk_out);
    for (int p = 0; p < NEMK_PK_MAX_WORLD; p++) if (h->xpeer[p] && p != h->rank && !h->xpeer_local) cudaIpcCloseMemHandle(h->xpeer[p]);
    if (h->b_xblk.p) cudaFree(h->b_xblk.p);
    if (h->h_lab_stage) cudaFreeHost(h->h_lab_stage);
    if (h->h_theta_stage) cudaFreeHost(h->h_theta_stage);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (int i = 0; i < 16; i++) if (h->copy_ev[i]) cudaEventDestroy(h->copy_ev[i]);
    for (int i = 0; i < h->ev_cap; i++) cudaEventDestroy(h->ev[i]);
    free(h->ev); free(h->ev_kind);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    free(h);
}

/* ------------------------------------------------------------------ loader: graph */
/* The CSR goes to the device as is; k_graph_check validates it there (row_ptr monotone,
 * neighbours in range) and tells whether every edge has its reverse, in which case the reader
 * lists of the speculative sweep ARE the neighbour lists.  Only a directed .nei file makes the
 * host build the transpose. */
static int upload(nemb_handle *h, dbuf *b, const void *src, size_t bytes)
{
    int rc = reserve(h, b, bytes);
    if (rc != NEMB_OK) return rc;
    if (bytes) CK(cudaMemcpyAsync(b->p, src, bytes, cudaMemcpyHostToDevice, h->stream));
    return NEMB_OK;
}

static int gather(nemb_handle *h, const void *send, void *recv, size_t bytes_per_rank);

/* A host array EVERY rank of a sharded fit holds (the global graph): each rank pushes one slice
 * through its own PCIe link and the slices are all-gathered over NVLink, so a rank uploads
 * bytes / world instead of bytes (8 GPUs, C4: 36 MB instead of 288 MB of graph per rank).  The
 * call is collective: all ranks pass the same array (nemb_load_shard's contract). */
static int upload_replicated(nemb_handle *h, dbuf *b, const void *src, size_t bytes)
{
    const char *mn = getenv("NEM_B200_SLICED_UPLOAD_MIN");     /* bytes; tests lower it */
    size_t min_bytes = mn && *mn ? (size_t)strtoull(mn, NULL, 10) : ((size_t)1 << 20);
    if (h->world <= 1 || bytes < min_bytes || bytes == 0 || getenv("NEM_B200_FULL_GRAPH_UPLOAD"))
        return upload(h, b, src, bytes);
    size_t slice = ((bytes + (size_t)h->world - 1) / (size_t)h->world + 255) & ~(size_t)255;
    int rc = reserve(h, b, slice * (size_t)h->world);
    if (rc != NEMB_OK) return rc;
    size_t off = slice * (size_t)h->rank;
    size_t mine = off < bytes ? (bytes - off < slice ? bytes - off : slice) : 0;
    if (mine) CK(cudaMemcpyAsync((char *)b->p + off, (const char *)src + off, mine, cudaMemcpyHostToDevice, h->stream));
    return gather(h, (char *)b->p + off, b->p, slice);
}

static int load_graph(nemb_handle *h, int n, const int32_t *row_ptr, const int32_t *col,
                      const float *wgt)
{
    int rc;
    h->spatial = row_ptr != NULL;
    h->nnz = 0; h->symmetric = 1; h->max_neigh = 0; h->n_heavy = 0; h->d_heavy = NULL;
    h->wgt_integral = 0;
    h->d_row_ptr = h->d_col = h->d_rrow_ptr = h->d_rcol = NULL; h->d_wgt = NULL;
    if (!h->spatial) return NEMB_OK;
    if (row_ptr[0] != 0) return fail(h, NEMB_E_ARG, "row_ptr[0] must be 0");
    int nnz = row_ptr[n];
    if (nnz < 0) return fail(h, NEMB_E_ARG, "row_ptr[n] < 0");
    if (nnz > 0 && (!col || !wgt)) return fail(h, NEMB_E_ARG, "col/wgt missing");
    h->nnz = nnz;
    if ((rc = upload_replicated(h, &h->b_row_ptr, row_ptr, sizeof(int32_t) * ((size_t)n + 1))) != NEMB_OK) return rc;
    if ((rc = upload_replicated(h, &h->b_col, col, sizeof(int32_t) * (size_t)nnz)) != NEMB_OK) return rc;
    if ((rc = upload_replicated(h, &h->b_wgt, wgt, sizeof(float) * (size_t)nnz)) != NEMB_OK) return rc;
    if ((rc = reserve(h, &h->b_flags, 64)) != NEMB_OK) return rc;
    h->d_row_ptr = h->b_row_ptr.p; h->d_col = h->b_col.p; h->d_wgt = h->b_wgt.p;
    nemk_graph_check(h->stream, n, nnz, h->d_row_ptr, h->d_col, h->d_wgt, (int32_t *)h->b_flags.p);
    CKK();
    /* hubs of this rank's rows (only meaningful if row_ptr is sane; re-checked below) */
    size_t hl_blocks = ((size_t)h->n + 1023) / 1024 + 1;
    if ((rc = reserve(h, &h->b_heavy, sizeof(int32_t) * (hl_blocks + (size_t)h->n + 1))) != NEMB_OK) return rc;
    h->d_heavy = (int32_t *)h->b_heavy.p + hl_blocks;
    nemk_heavy_list(h->stream, h->row0, h->n, h->d_row_ptr, h->b_heavy.p, h->d_heavy,
                    (int32_t *)h->b_flags.p + 2);
    CKK();
    int32_t flags[3] = {0, 0, 0};
    CK(cudaMemcpyAsync(flags, h->b_flags.p, sizeof flags, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (flags[0] & 1) return fail(h, NEMB_E_ARG, "row_ptr is not monotone");
    if (flags[0] & 2) return fail(h, NEMB_E_ARG, "neighbour index out of range");
    h->max_neigh = flags[1];
    h->n_heavy = flags[2];
    h->symmetric = !(flags[0] & 4);
    h->wgt_integral = !(flags[0] & 8) && !getenv("NEM_B200_ORDERED_SUMS");
    if (h->symmetric) { h->d_rrow_ptr = h->d_row_ptr; h->d_rcol = h->d_col; return NEMB_OK; }

    /* directed graph: reader lists = transpose of the CSR */
    int32_t *rrow = calloc((size_t)n + 1, sizeof(int32_t));
    int32_t *rcol = malloc(sizeof(int32_t) * (size_t)(nnz ? nnz : 1));
    int32_t *fill = malloc(sizeof(int32_t) * (size_t)(n ? n : 1));
    if (!rrow || !rcol || !fill) { free(rrow); free(rcol); free(fill); return fail(h, NEMB_E_MEMORY, "host alloc"); }
    for (int e = 0; e < nnz; e++) rrow[col[e] + 1]++;
    for (int i = 0; i < n; i++) rrow[i + 1] += rrow[i];
    for (int i = 0; i < n; i++) fill[i] = rrow[i];
    for (int i = 0; i < n; i++)
        for (int e = row_ptr[i]; e < row_ptr[i + 1]; e++) rcol[fill[col[e]]++] = i;
    rc = upload(h, &h->b_rrow_ptr, rrow, sizeof(int32_t) * ((size_t)n + 1));
    if (rc == NEMB_OK) rc = upload(h, &h->b_rcol, rcol, sizeof(int32_t) * (size_t)nnz);
    if (rc == NEMB_OK && cudaStreamSynchronize(h->stream) != cudaSuccess) rc = fail(h, NEMB_E_CUDA, "graph upload");
    free(rrow); free(rcol); free(fill);
    h->d_rrow_ptr = h->b_rrow_ptr.p; h->d_rcol = h->b_rcol.p;
    return rc;
}

/* Gauss-Seidel levels of the index-order sweep (i after every lower-index site it reads or is
 * read by), sites sorted by (level, index) and the launch schedule of the level-scheduled sweep.
 * Built on first use (level sweep, sequential fuzzy sweep, nemb_get_levels): the default
 * speculative sweep does not need it. */
static int ensure_levels(nemb_handle *h)
{
    if (h->have_levels || !h->spatial) return NEMB_OK;
    if (h->world > 1) return fail(h, NEMB_E_ARG, "the level schedule is not available on row shards");
    int n = h->n_glob, nnz = h->nnz;
    int32_t *row_ptr = malloc(sizeof(int32_t) * ((size_t)n + 1));
    int32_t *col = malloc(sizeof(int32_t) * (size_t)(nnz ? nnz : 1));
    int32_t *rrow = NULL, *rcol = NULL;
    if (!row_ptr || !col) { free(row_ptr); free(col); return fail(h, NEMB_E_MEMORY, "host alloc"); }
    CK(cudaMemcpyAsync(row_ptr, h->d_row_ptr, sizeof(int32_t) * ((size_t)n + 1), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(col, h->d_col, sizeof(int32_t) * (size_t)nnz, cudaMemcpyDeviceToHost, h->stream));
    if (!h->symmetric) {
        rrow = malloc(sizeof(int32_t) * ((size_t)n + 1));
        rcol = malloc(sizeof(int32_t) * (size_t)(nnz ? nnz : 1));
        CK(cudaMemcpyAsync(rrow, h->d_rrow_ptr, sizeof(int32_t) * ((size_t)n + 1), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(rcol, h->d_rcol, sizeof(int32_t) * (size_t)nnz, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    const int32_t *rr = rrow ? rrow : row_ptr, *rc_ = rcol ? rcol : col;

    int32_t *level = calloc((size_t)(n ? n : 1), sizeof(int32_t));
    int32_t *pend = calloc((size_t)(n ? n : 1), sizeof(int32_t));
    int depth = 0;
    for (int i = 0; i < n; i++) {
        int lv = pend[i];
        for (int e = row_ptr[i]; e < row_ptr[i + 1]; e++) {
            int j = col[e];
            if (j < i && level[j] > lv) lv = level[j];
        }
        level[i] = lv + 1;
        for (int e = row_ptr[i]; e < row_ptr[i + 1]; e++) {
            int j = col[e];
            if (j > i && level[i] > pend[j]) pend[j] = level[i];
        }
        for (int e = rr[i]; e < rr[i + 1]; e++) { /* sites that read i and come later */
            int j = rc_[e];
            if (j > i && level[i] > pend[j]) pend[j] = level[i];
        }
        if (level[i] > depth) depth = level[i];
    }
    free(pend);
    h->depth = depth;
    h->h_level = level;
    int32_t *lptr = calloc((size_t)depth + 2, sizeof(int32_t));
    int32_t *sites = malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) lptr[level[i]]++;          /* level l -> slot l (1-based) */
    for (int l = 1; l <= depth; l++) lptr[l] += lptr[l - 1];
    /* lptr[l] = number of sites with level <= l; start of level l (1-based) = lptr[l-1] */
    int32_t *pos = calloc((size_t)depth + 1, sizeof(int32_t));
    for (int l = 1; l <= depth; l++) pos[l] = lptr[l - 1];
    for (int i = 0; i < n; i++) sites[pos[level[i]]++] = i;
    free(pos);
    /* device level_ptr is 0-based over levels: level_ptr[q] = start of level q+1 */
    h->steps = malloc(sizeof(sweep_step) * (size_t)(depth ? depth : 1));
    h->n_steps = 0;
    for (int q = 0; q < depth;) {
        int width = lptr[q + 1] - lptr[q];
        if (width > NARROW_LEVEL) {
            int grid = (width + 255) / 256;
            if (grid > MAX_LEVEL_GRID) grid = MAX_LEVEL_GRID;
            if (grid < 2) grid = 2;
            h->steps[h->n_steps++] = (sweep_step){q, q + 1, grid};
            q++;
        } else {
            int q2 = q;
            while (q2 < depth && lptr[q2 + 1] - lptr[q2] <= NARROW_LEVEL) q2++;
            h->steps[h->n_steps++] = (sweep_step){q, q2, 1};
            q = q2;
        }
    }
    int rc = upload(h, &h->b_sites, sites, sizeof(int32_t) * (size_t)n);
    if (rc == NEMB_OK) rc = upload(h, &h->b_level_ptr, lptr, sizeof(int32_t) * ((size_t)depth + 1));
    if (rc == NEMB_OK && cudaStreamSynchronize(h->stream) != cudaSuccess) rc = fail(h, NEMB_E_CUDA, "level upload");
    h->d_sites = h->b_sites.p; h->d_level_ptr = h->b_level_ptr.p;
    free(row_ptr); free(col); free(rrow); free(rcol); free(lptr); free(sites);
    if (rc == NEMB_OK) h->have_levels = 1;
    return rc;
}

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

/* shard geometry + graph; X is already on its way to the device */
static int load_common(nemb_handle *h, int n_glob, int row0, int n_loc, int d,
                       const int32_t *row_ptr, const int32_t *col, const float *wgt)
{
    if (n_glob <= 0 || d <= 0) return fail(h, NEMB_E_ARG, "n and d must be > 0 (n=%d d=%d)", n_glob, d);
    int sl, r0, nl;
    nemb_shard_range(n_glob, h->world, h->rank, &sl, &r0, &nl);
    if (row0 != r0 || n_loc != nl)
        return fail(h, NEMB_E_ARG, "rank %d of %d must own rows [%d,%d) of %d (got [%d,%d))", h->rank,
                    h->world, r0, r0 + nl, n_glob, row0, row0 + n_loc);
    h->n = n_loc; h->n_glob = n_glob; h->row0 = row0; h->shard_len = sl;
    h->lab_len = sl * h->world;
    h->d = d;
    h->nwt = round_up((n_loc + 31) / 32, 4);
    if (h->nwt < 4) h->nwt = 4;
    int rc = load_graph(h, n_glob, row_ptr, col, wgt);
    if (rc == NEMB_OK) h->loaded = 1;
    return rc;
}

int nemb_load_shard(nemb_handle *h, int n_glob, int row0, int n_loc, int d, int wpr,
                    const uint32_t *x, const int32_t *row_ptr, const int32_t *col, const float *wgt)
{
    if (!h || (!x && n_loc > 0)) return NEMB_E_ARG;
    CK(cudaSetDevice(h->device));
    reset_problem(h);
    if (n_loc < 0 || d <= 0) return fail(h, NEMB_E_ARG, "bad shape");
    if (wpr < (d + 31) / 32) return fail(h, NEMB_E_ARG, "words_per_row %d too small for d=%d", wpr, d);
    int wpr_dev = round_up(wpr, 4), rc;
    size_t bytes = sizeof(uint32_t) * (size_t)n_loc * wpr_dev;
    if ((rc = reserve(h, &h->b_x, bytes)) != NEMB_OK) return rc;
    h->d_x = h->b_x.p;
    h->x_owned = 1;
    int pre_xt = 0;
    if (n_loc > 0) {
        if (wpr_dev == wpr && bytes >= ((size_t)64 << 20) && !getenv("NEM_B200_PLAIN_UPLOAD")) {
            /* large X: upload in chunks on a copy stream and transpose every chunk (X^T feeds the
             * first M-step) on the engine's stream as soon as it has landed, so the transpose and
             * the graph upload/validation below hide behind the PCIe transfer */
            int nwt = round_up((n_loc + 31) / 32, 4);
            if (nwt < 4) nwt = 4;
            if ((rc = reserve(h, &h->b_xt, sizeof(uint32_t) * (size_t)d * nwt)) != NEMB_OK) return rc;
            if ((rc = reserve(h, &h->b_pop, sizeof(int32_t) * (size_t)n_loc)) != NEMB_OK) return rc;
            if (!h->copy_stream) CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
            CK(cudaMemsetAsync(h->b_xt.p, 0, sizeof(uint32_t) * (size_t)d * nwt, h->stream));
            int chunks = 8;
            int rows_per = round_up((n_loc + chunks - 1) / chunks, 256);
            for (int c = 0, r0 = 0; r0 < n_loc; c++, r0 += rows_per) {
                int rows = n_loc - r0 < rows_per ? n_loc - r0 : rows_per;
                if (!h->copy_ev[c]) CK(cudaEventCreateWithFlags(&h->copy_ev[c], cudaEventDisableTiming));
                CK(cudaMemcpyAsync(h->d_x + (size_t)r0 * wpr, x + (size_t)r0 * wpr,
                                   sizeof(uint32_t) * (size_t)rows * wpr, cudaMemcpyHostToDevice, h->copy_stream));
                CK(cudaEventRecord(h->copy_ev[c], h->copy_stream));
            }
            pre_xt = rows_per;
        } else if (wpr_dev == wpr) {
            CK(cudaMemcpyAsync(h->d_x, x, bytes, cudaMemcpyHostToDevice, h->stream));
        } else {
            CK(cudaMemsetAsync(h->d_x, 0, bytes, h->stream));
            CK(cudaMemcpy2DAsync(h->d_x, sizeof(uint32_t) * (size_t)wpr_dev, x,
                                 sizeof(uint32_t) * (size_t)wpr, sizeof(uint32_t) * (size_t)wpr, n_loc,
                                 cudaMemcpyHostToDevice, h->stream));
        }
    }
    h->wpr = wpr_dev;
    /* the graph goes up and is validated on the engine's stream while X is still in flight */
    rc = load_common(h, n_glob, row0, n_loc, d, row_ptr, col, wgt);
    if (pre_xt) {
        for (int c = 0, r0 = 0; r0 < n_loc; c++, r0 += pre_xt) {
            int rows = n_loc - r0 < pre_xt ? n_loc - r0 : pre_xt;
            cudaStreamWaitEvent(h->stream, h->copy_ev[c], 0);
            nemk_transpose_bits_rows(h->stream, h->d_x, r0, rows, wpr_dev, d, h->nwt, h->b_xt.p);
            if (h->b_pop.p) nemk_row_popcount(h->stream, h->d_x + (size_t)r0 * wpr_dev, rows, wpr_dev, (int32_t *)h->b_pop.p + r0);
        }
        /* the caller may reuse its buffer when we return */
        cudaError_t e = cudaStreamSynchronize(h->copy_stream);
        if (e != cudaSuccess && rc == NEMB_OK) rc = fail(h, NEMB_E_CUDA, "X upload: %s", cudaGetErrorString(e));
        if (rc == NEMB_OK) { h->d_xt = h->b_xt.p; h->have_xt = 1; }
        if (rc == NEMB_OK && h->b_pop.p) { h->d_pop = h->b_pop.p; h->have_pop = 1; }
    }
    return rc;
}

int nemb_load_shard_device(nemb_handle *h, int n_glob, int row0, int n_loc, int d, int wpr,
                           const uint32_t *x_dev, const int32_t *row_ptr, const int32_t *col,
                           const float *wgt)
{
    if (!h || (!x_dev && n_loc > 0)) return NEMB_E_ARG;
    CK(cudaSetDevice(h->device));
    reset_problem(h);
    if (wpr % 4 || wpr < (d + 31) / 32)
        return fail(h, NEMB_E_ARG, "device X needs words_per_row %% 4 == 0 and >= ceil(d/32)");
    h->d_x = (uint32_t *)x_dev;
    h->x_owned = 0;
    h->wpr = wpr;
    return load_common(h, n_glob, row0, n_loc, d, row_ptr, col, wgt);
}

static int need_single(nemb_handle *h, const char *what)
{
    if (h->world > 1) return fail(h, NEMB_E_ARG, "%s is a single-GPU entry point; use nemb_load_shard", what);
    return NEMB_OK;
}

int nemb_load_packed(nemb_handle *h, int n, int d, int wpr, const uint32_t *x,
                     const int32_t *row_ptr, const int32_t *col, const float *wgt)
{
    if (!h || !x) return NEMB_E_ARG;
    int rc = need_single(h, "nemb_load_packed");
    if (rc != NEMB_OK) return rc;
    return nemb_load_shard(h, n, 0, n, d, wpr, x, row_ptr, col, wgt);
}

int nemb_load_packed_device(nemb_handle *h, int n, int d, int wpr, const uint32_t *x_dev,
                            const int32_t *row_ptr, const int32_t *col, const float *wgt)
{
    if (!h || !x_dev) return NEMB_E_ARG;
    int rc = need_single(h, "nemb_load_packed_device");
    if (rc != NEMB_OK) return rc;
    return nemb_load_shard_device(h, n, 0, n, d, wpr, x_dev, row_ptr, col, wgt);
}

int nemb_load_dense_u8(nemb_handle *h, int n, int d, const uint8_t *x, const int32_t *row_ptr,
                       const int32_t *col, const float *wgt)
{
    if (!h || !x) return NEMB_E_ARG;
    int rc = need_single(h, "nemb_load_dense_u8");
    if (rc != NEMB_OK) return rc;
    CK(cudaSetDevice(h->device));
    reset_problem(h);
    if (n <= 0 || d <= 0) return fail(h, NEMB_E_ARG, "n and d must be > 0");
    int wpr = round_up((d + 31) / 32, 4);
    /* the dense bytes are staged in the (not yet needed) transposed-bits buffer */
    if ((rc = reserve(h, &h->b_xt, (size_t)n * d)) != NEMB_OK) return rc;
    if ((rc = reserve(h, &h->b_x, sizeof(uint32_t) * (size_t)n * wpr)) != NEMB_OK) return rc;
    h->d_x = h->b_x.p;
    h->x_owned = 1;
    CK(cudaMemcpyAsync(h->b_xt.p, x, (size_t)n * d, cudaMemcpyHostToDevice, h->stream));
    nemk_pack_u8(h->stream, h->b_xt.p, n, d, wpr, h->d_x);
    CKK();
    h->wpr = wpr;
    return load_common(h, n, 0, n, d, row_ptr, col, wgt);
}

/* ------------------------------------------------------------------ per-K buffers */
static size_t carve(size_t *off, size_t bytes)
{
    size_t at = (*off + 255) & ~(size_t)255;
    *off = at + bytes;
    return at;
}

static int ensure_k(nemb_handle *h, int k)
{
    if (h->k_alloc == k) return NEMB_OK;
    size_t n = h->n, d = h->d, kd = (size_t)k * d, kw = (size_t)k * h->wpr, L = h->lab_len;
    size_t W = h->world, stat = kd + k;
    h->crit_blocks = h->world > 1 ? CRIT_BLOCKS_SHARD : (int)((n + 255) / 256);
    if (h->crit_blocks > CRIT_BLOCKS_MAX) h->crit_blocks = CRIT_BLOCKS_MAX;
    if (h->crit_blocks < 1) h->crit_blocks = 1;
    size_t off = 0;
    size_t o_prop = carve(&off, sizeof(float) * k), o_center = carve(&off, sizeof(float) * kd);
    size_t o_disp = carve(&off, sizeof(float) * kd), o_iner = carve(&off, sizeof(float) * kd);
    size_t o_coef = carve(&off, sizeof(nemk_coef));
    size_t o_mxor = carve(&off, 4 * kw), o_mval = carve(&off, 4 * kw);
    size_t o_f0 = carve(&off, 4 * kw), o_f1 = carve(&off, 4 * kw);
    size_t o_delta = carve(&off, sizeof(double) * (kd + k));
    size_t o_logpf = carve(&off, sizeof(double) * n * k);
    size_t o_lab0 = carve(&off, L), o_lab1 = carve(&off, L), o_lab2 = carve(&off, L);
    size_t o_lab3 = carve(&off, n + 1), o_ham = carve(&off, 4 * n * k + 4), o_sl = carve(&off, 4 * (kd + k));
    size_t o_margin = carve(&off, 4 * n + 4), o_stale0 = carve(&off, L), o_stale1 = carve(&off, L);
    size_t delta_b = sizeof(int32_t) * (2 + 2 * DELTA_CAP);
    size_t o_xchg = carve(&off, delta_b), o_xchg_all = carve(&off, delta_b * W);
    size_t o_dirty = carve(&off, 4 * L);
    size_t o_wl0 = carve(&off, 4 * (n + 1)), o_wl1 = carve(&off, 4 * (n + 1));
    size_t o_wlc = carve(&off, 4 * 8);
    size_t o_cm = carve(&off, 4 * (size_t)k * h->nwt);
    size_t o_si = carve(&off, 4 * stat), o_sis = carve(&off, 4 * stat * W);
    size_t o_sd = carve(&off, 8 * stat), o_sds = carve(&off, 8 * stat * W);
    size_t o_crit = carve(&off, sizeof(double) * 4 * h->crit_blocks * W);
    size_t o_status = carve(&off, sizeof(iter_status));
    size_t o_cnt = carve(&off, sizeof(nemk_counters) * W);
    int rc = reserve(h, &h->b_slab, off);
    if (rc != NEMB_OK) return rc;
    char *base = h->b_slab.p;
    h->d_prop = (float *)(base + o_prop); h->d_center = (float *)(base + o_center);
    h->d_disp = (float *)(base + o_disp); h->d_iner = (float *)(base + o_iner);
    h->d_coef = (nemk_coef *)(base + o_coef);
    h->d_mxor = (uint32_t *)(base + o_mxor); h->d_mval = (uint32_t *)(base + o_mval);
    h->d_f0 = (uint32_t *)(base + o_f0); h->d_f1 = (uint32_t *)(base + o_f1);
    h->d_delta = (double *)(base + o_delta); h->d_logpf = (double *)(base + o_logpf);
    h->d_lab[0] = (uint8_t *)(base + o_lab0); h->d_lab[1] = (uint8_t *)(base + o_lab1);
    h->d_lab[2] = (uint8_t *)(base + o_lab2); h->d_lab[3] = (uint8_t *)(base + o_lab3);
    h->d_ham = (int32_t *)(base + o_ham); h->d_stat_loc = (int32_t *)(base + o_sl);
    h->d_margin = (float *)(base + o_margin);
    h->d_xchg = (int32_t *)(base + o_xchg); h->d_xchg_all = (int32_t *)(base + o_xchg_all);
    h->d_stale[0] = (uint8_t *)(base + o_stale0); h->d_stale[1] = (uint8_t *)(base + o_stale1);
    h->ham_valid = 0; h->stats_valid = 0; h->last_changed = -1;
    h->d_dirty = (int32_t *)(base + o_dirty);
    h->d_wl[0] = (int32_t *)(base + o_wl0); h->d_wl[1] = (int32_t *)(base + o_wl1);
    h->d_wl_counts = (int32_t *)(base + o_wlc);
    h->d_cm = (uint32_t *)(base + o_cm);
    h->d_stat_int = (int32_t *)(base + o_si); h->d_stat_int_stage = (int32_t *)(base + o_sis);
    h->d_stat_dbl = (double *)(base + o_sd); h->d_stat_dbl_stage = (double *)(base + o_sds);
    h->d_crit_partials = (double *)(base + o_crit);
    h->d_status = (iter_status *)(base + o_status);
    h->d_cnt_all = (nemk_counters *)(base + o_cnt);
    CK(cudaMemsetAsync(h->d_coef, 0, sizeof(nemk_coef), h->stream));
    CK(cudaMemsetAsync(h->d_dirty, 0, 4 * L, h->stream));
    CK(cudaMemsetAsync(h->d_wl_counts, 0, 4 * 8, h->stream));
    CK(cudaMemsetAsync(h->d_status, 0, sizeof(iter_status), h->stream));
    h->d_t[0] = h->d_t[1] = NULL;
    h->d_partial_s = h->d_partial_n = NULL;
    if (!h->h_status) CK(cudaMallocHost((void **)&h->h_status, sizeof(iter_status)));
    if (!h->h_cnt_all) CK(cudaMallocHost((void **)&h->h_cnt_all, sizeof(nemk_counters) * MAX_WORLD));
    if (!h->h_empty) CK(cudaMallocHost((void **)&h->h_empty, 2 * sizeof(int32_t)));
    if (!h->ring) {
        CK(cudaHostAlloc((void **)&h->ring, sizeof(nemk_host_status) * RING, cudaHostAllocMapped | cudaHostAllocPortable));
        memset(h->ring, 0, sizeof(nemk_host_status) * RING);
        CK(cudaHostGetDevicePointer((void **)&h->d_ring, h->ring, 0));
    }
    h->k_alloc = k;
    return NEMB_OK;
}

/* float posteriors [lab_len][K]: only algo nem (both buffers) and nemb_get_posteriors (one) */
static int ensure_t(nemb_handle *h, int k, int both)
{
    size_t bytes = sizeof(float) * (size_t)h->lab_len * k;
    for (int b = 0; b < (both ? 2 : 1); b++) {
        int rc = reserve(h, &h->b_t[b], bytes);
        if (rc != NEMB_OK) return rc;
        h->d_t[b] = h->b_t[b].p;
    }
    return NEMB_OK;
}

/* fuzzy M-step scratch: per-chunk partial sums */
static int ensure_nem_scratch(nemb_handle *h, int k)
{
    size_t kd = (size_t)k * h->d;
    /* chunks of 1024 families (4096 beyond 512k: the partial sums are K*D doubles per chunk): with
     * 512 genomes per CTA the grid needs that many chunks to fill the SMs */
    h->rows_per_chunk = h->n <= (1 << 19) ? 1024 : 4096;
    h->nchunks = (h->n + h->rows_per_chunk - 1) / h->rows_per_chunk;
    if (h->nchunks < 1) h->nchunks = 1;
    size_t off = 0;
    size_t o_s = carve(&off, sizeof(double) * kd * h->nchunks);
    size_t o_n = carve(&off, sizeof(double) * (size_t)k * h->nchunks);
    int rc = reserve(h, &h->b_nem, off);
    if (rc != NEMB_OK) return rc;
    h->d_partial_s = (double *)((char *)h->b_nem.p + o_s);
    h->d_partial_n = (double *)((char *)h->b_nem.p + o_n);
    return NEMB_OK;
}

static int ensure_xt(nemb_handle *h)
{
    if (h->have_xt) return NEMB_OK;
    int rc = reserve(h, &h->b_xt, sizeof(uint32_t) * (size_t)h->d * h->nwt);
    if (rc != NEMB_OK) return rc;
    h->d_xt = h->b_xt.p;
    nemk_transpose_bits(h->stream, h->d_x, h->n, h->wpr, h->d, h->nwt, h->d_xt);
    CKK();
    h->have_xt = 1;
    return NEMB_OK;
}

/* ------------------------------------------------------------------ collectives */
static int gather(nemb_handle *h, const void *send, void *recv, size_t bytes_per_rank)
{
    if (h->world <= 1) {
        if (send != recv) CK(cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, h->stream));
        return NEMB_OK;
    }
    int rc = h->comm->allgather(h->comm->ctx, send, recv, bytes_per_rank, (void *)h->stream);
    h->exchanges++;
    if (rc) return fail(h, NEMB_E_CUDA, "all-gather failed on rank %d (code %d)", h->rank, rc);
    return NEMB_OK;
}

/* ------------------------------------------------------------------ stage timing */
static void stage_mark(nemb_handle *h, int kind_or_end)
{
    if (!h->profile) return;
    if (h->ev_n >= h->ev_cap) {
        int nc = h->ev_cap ? h->ev_cap * 2 : 256;
        h->ev = realloc(h->ev, sizeof(cudaEvent_t) * nc);
        h->ev_kind = realloc(h->ev_kind, sizeof(int) * nc);
        for (int i = h->ev_cap; i < nc; i++) cudaEventCreate(&h->ev[i]);
        h->ev_cap = nc;
    }
    h->ev_kind[h->ev_n] = kind_or_end;
    if (kind_or_end == ST_DENSITY) h->ev_last_density = h->ev_n;
    if (kind_or_end == ST_DENSITY_CACHED) h->ev_last_cached = h->ev_n;
    cudaEventRecord(h->ev[h->ev_n++], h->stream);
}
#define STAGE_BEGIN(kind) stage_mark(h, (kind))
#define STAGE_END() stage_mark(h, -1)

/* popcount-path eligibility of theta, evaluated on the host copy (mirrors k_theta_tables):
 * 0 = general path, 1 = popcount path, 2 = popcount path and every class has a constant centre
 * (all 0, all 1 or all 1/2: the Hamming counts follow from the row popcounts alone) */
static int theta_uniform(int k, int d, const float *center, const float *disp)
{
    int all_const = 1;
    for (int c = 0; c < k; c++) {
        float e0 = disp[(size_t)c * d];
        int n_valid = 0, n_x1 = 0;
        for (int j = 0; j < d; j++) {
            float mu = center[(size_t)c * d + j], e = disp[(size_t)c * d + j];
            int m0 = abs((int)(0.0f - mu)), m1 = abs((int)(1.0f - mu));
            if (memcmp(&e, &e0, sizeof e) || m0 > 1 || m1 > 1) return 0;
            n_valid += m0 != m1; n_x1 += m0 == 1 && m1 == 0;
        }
        if (!(n_valid == 0 || (n_valid == d && (n_x1 == 0 || n_x1 == d)))) all_const = 0;
    }
    return all_const ? 2 : 1;
}

static int ensure_pop(nemb_handle *h)
{
    if (h->have_pop) return NEMB_OK;
    int rc = reserve(h, &h->b_pop, sizeof(int32_t) * (size_t)(h->n ? h->n : 1));
    if (rc != NEMB_OK) return rc;
    h->d_pop = h->b_pop.p;
    nemk_row_popcount(h->stream, h->d_x, h->n, h->wpr, h->d_pop);
    h->launches++;
    CKK();
    h->have_pop = 1;
    return NEMB_OK;
}

/* ------------------------------------------------------------------ environment knobs */
/* Verified on real NVLink boxes at 2 and 4 ranks (tests/nccl_check.py, bench strong_scaling:
 * identical to the single-GPU fit).  At 8 ranks the C4-sized fit hung in its first beta sweep
 * (profiles/r2_8gpu_hang.txt: two ranks never reached cross-rank barrier 6, each stuck at a LOCAL
 * barrier of a fix-up round, i.e. the CTAs of a rank disagreed on a round's item count); until
 * that is understood, 8 ranks take the NCCL protocol of the launch-per-stage loop. */
#ifndef PK_SHARD_MAX_WORLD_DEFAULT
#define PK_SHARD_MAX_WORLD_DEFAULT 4
#endif
static int env_flag(const char *name) { const char *e = getenv(name); return e && *e; }
static void read_env_knobs(nemb_handle *h)
{
    const char *e;
    h->no_persist = env_flag("NEM_B200_NO_PERSIST");
    h->no_persist_shard = env_flag("NEM_B200_NO_PERSIST_SHARD");
    h->keep_logpf = env_flag("NEM_B200_KEEP_LOGPF");
    h->no_margins = env_flag("NEM_B200_NO_MARGINS");
    h->no_popcache = env_flag("NEM_B200_NO_POPCACHE");
    h->no_spec = env_flag("NEM_B200_NO_SPEC");
    h->full_mstep = env_flag("NEM_B200_FULL_MSTEP");
    h->full_exchange = env_flag("NEM_B200_FULL_EXCHANGE");
    /* worst case of the exact shortcuts (bench.py value_no_shortcuts): every iteration reads X (the
     * cached Hamming counts are declared void), recounts S through X^T and evaluates every site */
    h->no_shortcuts = env_flag("NEM_B200_NO_SHORTCUTS");
    if (h->no_shortcuts) { h->no_margins = 1; h->full_mstep = 1; h->no_popcache = 1; }
    e = getenv("NEM_B200_MEDIUM_LIST");
    h->medium_list = e && *e ? atoi(e) : 32768;
    /* largest world size the peer-memory kernel serves (beyond it: the NCCL protocol of the
     * launch-per-stage loop) */
    h->persist_local = env_flag("NEM_B200_PERSIST_LOCAL");
    e = getenv("NEM_B200_PERSIST_SHARD_MAX");
    h->pk_shard_max_world = e && *e ? atoi(e) : PK_SHARD_MAX_WORLD_DEFAULT;
    if (h->pk_shard_max_world > NEMK_PK_MAX_WORLD) h->pk_shard_max_world = NEMK_PK_MAX_WORLD;
    e = getenv("NEM_B200_PK_GRID");
    h->pk_grid_env = e && *e ? atoi(e) : 0;
    e = getenv("NEM_B200_PK_XLIMIT");           /* bytes of X the in-kernel X / X^T passes accept */
    h->pk_xlimit = e && *e ? (size_t)strtoull(e, NULL, 10) : ((size_t)64 << 20);
}

/* ------------------------------------------------------------------ steps of one fit */
static nemk_lpsrc lpsrc(const nemb_handle *h)
{
    nemk_lpsrc s = {h->d_logpf, h->lp_from_ham ? h->d_ham : NULL, h->d_coef, h->wgt_integral};
    return s;
}

/* next_uniform: the density pass that follows takes the popcount path; its cached Hamming
 * counts stay valid while the class bit masks do not move (k_theta_tables detects that) */
static int run_tables(nemb_handle *h, int k, int next_uniform)
{
    int force = !(h->ham_valid && next_uniform) || h->no_shortcuts;
    h->tables_forced = force;
    nemk_theta_tables(h->stream, k, h->d, h->wpr, h->d_prop, h->d_center, h->d_disp, h->d_coef,
                      h->d_mxor, h->d_mval, h->d_f0, h->d_f1, h->d_delta, force);
    h->launches++;
    CKK();
    return NEMB_OK;
}

/* lean: ncem fits on the popcount path keep only the Hamming counts; the sweeps and the criteria
 * rebuild logpf from them in registers (nemk_lpsrc), so neither the N*K*8-byte array nor the
 * separate cached-rebuild launch exists */
static int run_density(nemb_handle *h, int k, int uniform, int32_t *d_hamming, int lean)
{
    if (uniform == 2 && lean && d_hamming == NULL && !h->no_popcache) {
        /* every class has a constant centre: H from the cached row popcounts, X is not read */
        int rc = ensure_pop(h);
        if (rc != NEMB_OK) return rc;
        STAGE_BEGIN(ST_DENSITY_CACHED);
        nemk_ham_from_pop(h->stream, k, h->n, h->d, h->d_coef, h->d_pop, h->d_ham);
        STAGE_END();
        h->launches++;
        CKK();
        h->ham_valid = 1; h->lp_from_ham = 1;
        h->ev_last_density = h->ev_last_cached = -1;
        return NEMB_OK;
    }
    STAGE_BEGIN(ST_DENSITY);
    h->lp_from_ham = 0;
    if (uniform) {
        /* d_hamming == NULL: the fit's own pass, through the persistent H cache */
        lean = lean && d_hamming == NULL;
        nemk_density_uniform(h->stream, k, h->d, h->d_x, h->n, h->wpr, h->d_coef, h->d_mxor, h->d_mval,
                             lean ? NULL : h->d_logpf, d_hamming ? d_hamming : h->d_ham, d_hamming == NULL);
        h->ham_valid = d_hamming == NULL;
        h->lp_from_ham = lean;
        if (lean) {
            /* the kernel returns at once when the masks did not move (coef->mu_changed == 0):
             * read_status relabels this stage as "cached" then */
            h->ev_last_cached = -1;
            if (h->profile && h->tables_forced) h->ev_last_density = -1;   /* the X pass certainly ran */
        } else if (h->ham_valid) {   /* exactly one of the two kernels does work (coef->mu_changed) */
            STAGE_END();
            STAGE_BEGIN(ST_DENSITY_CACHED);
            nemk_logpf_from_cache(h->stream, k, h->n, h->d_coef, h->d_ham, h->d_logpf);
            h->launches++;
            if (h->profile && h->tables_forced) {   /* the X pass certainly ran */
                h->ev_kind[h->ev_last_cached] = -2;
                h->ev_last_density = h->ev_last_cached = -1;
            }
        }
    } else {
        nemk_density_general(h->stream, k, h->d_x, h->n, h->d, h->wpr, h->d_coef, h->d_f0, h->d_f1,
                             h->d_delta, h->d_logpf);
        h->ham_valid = 0;
    }
    STAGE_END();
    h->launches++;
    CKK();
    return NEMB_OK;
}

/* End of a sweep / iteration: (all-gather of the ranks' counters,) one tiny kernel that sums them,
 * decides convergence when `o` is given (EM iteration; it raises the device halt flag) and writes
 * the status block into a mapped pinned slot.  Returns the sequence number to wait for. */
static int publish_status(nemb_handle *h, const nemb_options *o, unsigned long long *seq_out)
{
    const nemk_counters *cnt_all = &h->d_status->cnt;
    if (h->world > 1) {
        int rc = gather(h, &h->d_status->cnt, h->d_cnt_all, sizeof(nemk_counters));
        if (rc != NEMB_OK) return rc;
        cnt_all = h->d_cnt_all;
    }
    unsigned long long seq = ++h->seq;
    int decide = o && o->conv != NEMB_CONV_CRIT;
    nemk_iter_end(h->stream, h->world, cnt_all, h->d_status, h->d_coef, decide,
                  o ? o->algo == NEMB_ALGO_NCEM : 0, o ? o->conv : 0, o ? o->conv_thr : 0.f,
                  h->d_ring + (seq % RING), seq);
    h->launches++;
    CKK();
    *seq_out = seq;
    return NEMB_OK;
}

/* spin on the slot's sequence number (the kernel wrote it after a system-wide fence) */
static int wait_status(nemb_handle *h, unsigned long long seq)
{
    volatile nemk_host_status *s = &h->ring[seq % RING];
    for (unsigned spins = 1; s->seq != seq; spins++) {
        if ((spins & 0xfff) == 0) {
            cudaError_t e = cudaStreamQuery(h->stream);
            if (e != cudaSuccess && e != cudaErrorNotReady)
                return fail(h, NEMB_E_CUDA, "device error while waiting for the status: %s", cudaGetErrorString(e));
            if (e == cudaSuccess && s->seq != seq)
                return fail(h, NEMB_E_BUG, "stream drained without status %llu (slot holds %llu)", seq, (unsigned long long)s->seq);
        }
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
        if (h->poll_relaxed && spins > 64) {      /* resample workers: other streams keep the GPU busy */
            struct timespec ts = {0, 2000};
            nanosleep(&ts, NULL);
        } else if ((spins & 0xff) == 0) sched_yield();   /* several engines may poll on few cores */
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    const nemk_host_status *r = &h->ring[seq % RING];
    h->h_status->cnt = r->cnt;
    memcpy(h->h_status->crit_before, r->crit_before, sizeof r->crit_before);
    memcpy(h->h_status->crit_after, r->crit_after, sizeof r->crit_after);
    h->h_empty[0] = r->empty_class;
    h->h_empty[1] = r->mu_changed;
    /* h_empty[1] = coef->mu_changed of the last tables: 0 => the last popcount density pass did
     * not read X (cached Hamming counts) */
    if (h->profile && h->ham_valid && h->ev_last_density >= 0 && h->ev_last_cached >= 0) {
        /* discard the stage whose kernel returned at once */
        h->ev_kind[h->h_empty[1] ? h->ev_last_cached : h->ev_last_density] = -2;
        h->ev_last_density = h->ev_last_cached = -1;
    } else if (h->profile && h->lp_from_ham && h->ev_last_density >= 0) {
        if (!h->h_empty[1]) h->ev_kind[h->ev_last_density] = ST_DENSITY_CACHED;   /* X not read */
        h->ev_last_density = -1;
    }
    return NEMB_OK;
}

/* counters of the last sweep, summed over the ranks, on the host */
static int read_status(nemb_handle *h)
{
    unsigned long long seq;
    int rc = publish_status(h, NULL, &seq);
    if (rc != NEMB_OK) return rc;
    return wait_status(h, seq);
}

/* jacobi round 0 left the first work list in list 0 / counter 0: three grid-wide rounds, then
 * one CTA walks the tail to exhaustion (and leaves the four counters at 0) */
enum { GRID_ROUNDS = 3, SHORT_LIST = 32768, MEDIUM_LIST = 32768 };
static void local_fixups(nemb_handle *h, int k, double beta, const uint8_t *in, uint8_t *out,
                         const int32_t *rp, const int32_t *skip, const nemk_iter_end_args *fused)
{
    /* few labels moved last iteration => the work lists are short: the tail cluster (8 CTAs) walks
     * them all; a medium list gets ONE grid-wide round first (the whole GPU takes the bulk, the
     * tail the rest); long or unknown lists get GRID_ROUNDS */
    const int medium = h->medium_list;
    int grid_rounds = (h->last_changed >= 0 && h->last_changed < SHORT_LIST)
                          ? (h->last_changed >= medium ? 1 : 0) : GRID_ROUNDS;
    for (int r = 0; r < grid_rounds; r++)
        nemk_sweep_ncem_fixup_round(h->stream, k, h->row0, h->n, lpsrc(h), rp, h->d_col, h->d_wgt,
                                    beta, in, out, h->d_dirty, h->d_wl[0], h->d_wl[1],
                                    h->d_wl_counts, r, h->d_rrow_ptr, h->d_rcol, &h->d_status->cnt,
                                    skip, h->mg);
    nemk_sweep_ncem_fixup(h->stream, k, h->row0, h->n, lpsrc(h), rp, h->d_col, h->d_wgt, beta, in,
                          out, h->d_dirty, h->d_wl[0], h->d_wl[1], h->d_wl_counts, grid_rounds,
                          h->d_rrow_ptr, h->d_rcol, &h->d_status->cnt, skip, fused, h->mg);
    h->launches += 1 + grid_rounds;
}

/* one E-step sweep; *flipped tells whether the state moved to the other buffer */
/* decide (nullable): the options of the EM iteration this sweep ends -- the sharded sweep then
 * publishes the iteration's status itself; *status_read tells the caller it did */
static int run_sweep(nemb_handle *h, const nemb_options *o, double beta, int *flipped,
                     const nemb_options *decide, int *status_read)
{
    int k = o->k, n = h->n, row0 = h->row0, rc;
    const int32_t *skip = &h->d_coef->empty_class;
    const int32_t *rp = h->spatial ? h->d_row_ptr : NULL;
    int seq = o->update == NEMB_UPDATE_SEQ && h->spatial && beta != 0.0;
    size_t SL = h->shard_len;
    *flipped = 0;
    if (status_read) *status_read = 0;
    memset(&h->mg, 0, sizeof h->mg);
    CK(cudaMemsetAsync(&h->d_status->cnt, 0, sizeof(nemk_counters), h->stream));
    STAGE_BEGIN(ST_SWEEP);
    if (o->algo == NEMB_ALGO_NCEM) {
        uint8_t *in = h->d_lab[h->cur], *out = h->d_lab[h->cur ^ 1], *seen = h->d_lab[2];
        int impl = o->sweep_impl == NEMB_SWEEP_AUTO ? NEMB_SWEEP_SPEC : o->sweep_impl;
        if (!seq) {
            nemk_sweep_ncem_jacobi(h->stream, k, row0, n, lpsrc(h), rp, h->d_col, h->d_wgt, beta, in,
                                   out, NULL, NULL, NULL, NULL, NULL, h->d_heavy, h->n_heavy, &h->d_status->cnt, skip,
                                   0, 0, h->mg);
            h->launches++;
            /* halo exchange of the hard labels: every rank's slice, 1 byte per family */
            if (h->world > 1) {
                if ((rc = gather(h, out + (size_t)h->rank * SL, out, SL)) != NEMB_OK) return rc;
                CK(cudaMemcpyAsync(seen, out, h->lab_len, cudaMemcpyDeviceToDevice, h->stream));   /* invariant: seen == labels */
            }
            *flipped = 1;
        } else if (impl == NEMB_SWEEP_SPEC) {
            /* margin cache (one GPU, Hamming-count scores): sites whose margin exceeds what theta
             * can have moved and whose later neighbours kept their label are copied, not evaluated */
            if (h->world == 1 && h->lp_from_ham && !h->no_margins) {
                h->mg.m = h->d_margin;
                h->mg.stale_cur = h->d_stale[h->stale_par];
                h->mg.stale_next = h->d_stale[h->stale_par ^ 1];
                h->mg.on = h->sweep_same_beta;
                h->stale_par ^= 1;
            }
            /* row shards: the jacobi kernel also copies the other ranks' slices in -> out (remote
             * labels start the sweep at their previous value) */
            nemk_sweep_ncem_jacobi(h->stream, k, row0, n, lpsrc(h), rp, h->d_col, h->d_wgt, beta, in,
                                   out, h->d_dirty, h->d_wl[0], &h->d_wl_counts[0], h->d_rrow_ptr,
                                   h->d_rcol, h->d_heavy, h->n_heavy, &h->d_status->cnt, skip,
                                   h->world, (int)SL, h->mg);
            h->launches++;
            if (h->world == 1 && decide) {
                /* one GPU: the tail of the fix-up rounds ends the iteration -- it also decides
                 * convergence and publishes the status (no separate nemk_iter_end launch) */
                nemk_iter_end_args fa = {1, &h->d_status->cnt, h->d_status, h->d_coef,
                                         decide->conv != NEMB_CONV_CRIT, 1, decide->conv, decide->conv_thr,
                                         NULL, 0};
                fa.seq = ++h->seq;
                fa.host = h->d_ring + (fa.seq % RING);
                local_fixups(h, k, beta, in, out, rp, skip, &fa);
                if (status_read) *status_read = 1;
            } else
                local_fixups(h, k, beta, in, out, rp, skip, NULL);
            if (h->world > 1) {
                /* speculative fixed point ACROSS ranks: exchange label slices, queue the local
                 * readers of every remote label that moved, fix up, until no rank queues anything
                 * (then every rank holds the sequential sweep's labels for all families).  ONE
                 * all-gather (the labels) and one host poll per round: every rank derives the
                 * global pending / changed counts from the exchanged labels themselves
                 * (k_mark_remote), and the last round's status -- convergence decided on the
                 * device when `decide` is given -- doubles as the iteration's status. */
                const int dec = decide && decide->conv != NEMB_CONV_CRIT;
                const size_t delta_b = sizeof(int32_t) * (2 + 2 * DELTA_CAP);
                for (int guard = 0;; guard++) {
                    unsigned long long seq;
                    /* sparse round first when few labels are expected to have moved: blocks of
                     * (family, label) pairs instead of whole slices.  `seen` holds, on every
                     * rank, the labels as of the previous exchange (== the input labels at the
                     * start of a sweep). */
                    int sparse = !h->full_exchange &&
                                 (guard > 0 || (h->last_changed >= 0 && h->last_changed <= DELTA_CAP));
                    int settled = -1;
                    if (sparse) {
                        nemk_delta_pack(h->stream, row0, n, DELTA_CAP, out, seen, &h->d_status->cnt,
                                        h->d_xchg, skip);
                        if ((rc = gather(h, h->d_xchg, h->d_xchg_all, delta_b)) != NEMB_OK) return rc;
                        CK(cudaMemsetAsync(&h->d_status->cnt.pending, 0, 2 * sizeof(int32_t), h->stream));   /* + changed_glob */
                        nemk_delta_apply(h->stream, h->world, DELTA_CAP, h->d_xchg_all, row0, n, (int)SL, out,
                                         seen, h->d_dirty, h->d_wl[0], &h->d_wl_counts[0], h->d_rrow_ptr,
                                         h->d_rcol, &h->d_status->cnt, skip);
                        seq = ++h->seq;
                        nemk_iter_end(h->stream, 0, &h->d_status->cnt, h->d_status, h->d_coef, dec, 1,
                                      decide ? decide->conv : 0, decide ? decide->conv_thr : 0.f,
                                      h->d_ring + (seq % RING), seq);
                        h->launches += 3;
                        CKK();
                        if ((rc = wait_status(h, seq)) != NEMB_OK) return rc;
                        settled = h->h_status->cnt.pending;   /* < 0: some rank moved too many labels */
                    }
                    if (settled < 0) {
                        if ((rc = gather(h, out + (size_t)h->rank * SL, out, SL)) != NEMB_OK) return rc;
                        CK(cudaMemsetAsync(&h->d_status->cnt.pending, 0, 2 * sizeof(int32_t), h->stream));   /* + changed_glob */
                        nemk_mark_remote(h->stream, h->n_glob, row0, n, (int)SL, out, in, seen, seen,
                                         h->d_dirty, h->d_wl[0], &h->d_wl_counts[0], h->d_rrow_ptr,
                                         h->d_rcol, &h->d_status->cnt, skip);
                        seq = ++h->seq;
                        nemk_iter_end(h->stream, 0, &h->d_status->cnt, h->d_status, h->d_coef, dec, 1,
                                      decide ? decide->conv : 0, decide ? decide->conv_thr : 0.f,
                                      h->d_ring + (seq % RING), seq);
                        h->launches += 2;
                        CKK();
                        if ((rc = wait_status(h, seq)) != NEMB_OK) return rc;
                    }
                    if (h->h_status->cnt.pending == 0 || *h->h_empty) break;
                    if (guard > h->n_glob) return fail(h, NEMB_E_BUG, "sharded sweep did not settle");
                    local_fixups(h, k, beta, in, out, rp, skip, NULL);
                }
                if (status_read) *status_read = 1;
            }
            *flipped = 1;
        } else {
            if ((rc = ensure_levels(h)) != NEMB_OK) return rc;
            for (int s = 0; s < h->n_steps; s++) {
                nemk_sweep_ncem_level(h->stream, k, lpsrc(h), rp, h->d_col, h->d_wgt, beta, in,
                                      h->d_sites, h->d_level_ptr, h->steps[s].lo, h->steps[s].hi,
                                      h->steps[s].grid, &h->d_status->cnt, skip);
                h->launches++;
            }
        }
    } else {
        float *in = h->d_t[h->cur], *out = h->d_t[h->cur ^ 1];
        if (!seq) {
            nemk_sweep_nem_jacobi(h->stream, k, row0, n, h->d_logpf, rp, h->d_col, h->d_wgt, beta, in,
                                  out, &h->d_status->cnt, skip);
            h->launches++;
            size_t slice = sizeof(float) * SL * k;   /* halo exchange of the posteriors */
            if (h->world > 1 &&
                (rc = gather(h, (char *)out + (size_t)h->rank * slice, out, slice)) != NEMB_OK) return rc;
            *flipped = 1;
        } else {
            if ((rc = ensure_levels(h)) != NEMB_OK) return rc;
            for (int s = 0; s < h->n_steps; s++) {
                nemk_sweep_nem_level(h->stream, k, h->d_logpf, rp, h->d_col, h->d_wgt, beta, in,
                                     h->d_sites, h->d_level_ptr, h->steps[s].lo, h->steps[s].hi,
                                     h->steps[s].grid, &h->d_status->cnt, skip);
                h->launches++;
            }
        }
    }
    STAGE_END();
    CKK();
    if (*flipped) h->cur ^= 1;
    h->prev_valid = *flipped;
    return NEMB_OK;
}

static int run_mstep(nemb_handle *h, const nemb_options *o, int next_uniform)
{
    int k = o->k, rc;
    size_t kd = (size_t)k * h->d, stat = kd + k;
    int incremental = o->algo == NEMB_ALGO_NCEM && h->stats_valid && h->prev_valid && h->last_changed >= 0 &&
                      h->last_changed <= h->n / 8 && !h->full_mstep;
    STAGE_BEGIN(incremental ? ST_MSTEP_DELTA : ST_MSTEP);
    if (o->algo == NEMB_ALGO_NCEM) {
        /* exact integer counts S_kd, n_k of this rank's rows: a full recount through X^T, or --
         * when the last sweep moved few labels -- an update from the rows that changed class */
        const uint8_t *lab_loc = h->d_lab[h->cur] + h->row0;
        if (incremental) {
            /* the statistics describe the labels the last sweep started from: its other buffer */
            nemk_mstep_delta(h->stream, k, h->n, h->d, h->wpr, h->d_x, lab_loc,
                             h->d_lab[h->cur ^ 1] + h->row0,
                             h->d_wl[1], &h->d_wl_counts[4], h->d_stat_loc, h->d_stat_loc + kd,
                             &h->d_coef->halt);
        } else {
            if ((rc = ensure_xt(h)) != NEMB_OK) return rc;
            nemk_label_masks(h->stream, k, h->n, h->nwt, lab_loc, h->d_cm, h->d_stat_loc + kd,
                             NULL, &h->d_coef->halt);
            nemk_mstep_ncem(h->stream, k, h->d, h->nwt, h->d_xt, h->d_cm, h->d_stat_loc,
                            &h->d_coef->halt);
            h->stats_valid = 1;
        }
        h->launches += 2;
        const int32_t *stat_glob = h->d_stat_loc;
        if (h->world > 1) {
            if ((rc = gather(h, h->d_stat_loc, h->d_stat_int_stage, sizeof(int32_t) * stat)) != NEMB_OK) return rc;
            nemk_sum_ranks_i32(h->stream, h->world, stat, h->d_stat_int_stage, h->d_stat_int);
            h->launches++;
            stat_glob = h->d_stat_int;
        }
        nemk_mstep_finalize_tables(h->stream, k, h->n_glob, h->d, h->wpr, o->prop, o->disp, stat_glob,
                                   stat_glob + kd, NULL, NULL, h->d_prop, h->d_center, h->d_disp,
                                   h->d_coef, h->d_mxor, h->d_mval, h->d_f0, h->d_f1, h->d_delta,
                                   !(h->ham_valid && next_uniform) || h->no_shortcuts);
        h->tables_forced = !(h->ham_valid && next_uniform) || h->no_shortcuts;
        h->launches++;
    } else {
        if ((rc = ensure_nem_scratch(h, k)) != NEMB_OK) return rc;
        nemk_mstep_nem(h->stream, k, h->n, h->d, h->wpr, h->d_x,
                       h->d_t[h->cur] + (size_t)h->row0 * k, h->rows_per_chunk, h->d_partial_s,
                       h->d_partial_n, h->d_stat_dbl, h->d_stat_dbl + kd);
        h->launches += 2;
        if (h->world > 1) {   /* rank-ordered float64 sum: same bits on every rank, every run */
            if ((rc = gather(h, h->d_stat_dbl, h->d_stat_dbl_stage, sizeof(double) * stat)) != NEMB_OK) return rc;
            nemk_sum_ranks_f64(h->stream, h->world, stat, h->d_stat_dbl_stage, h->d_stat_dbl);
            h->launches++;
        }
        nemk_mstep_finalize_tables(h->stream, k, h->n_glob, h->d, h->wpr, o->prop, o->disp, NULL, NULL,
                                   h->d_stat_dbl, h->d_stat_dbl + kd, h->d_prop, h->d_center,
                                   h->d_disp, h->d_coef, h->d_mxor, h->d_mval, h->d_f0, h->d_f1,
                                   h->d_delta, !(h->ham_valid && next_uniform));
        h->tables_forced = !(h->ham_valid && next_uniform);
        h->launches++;
    }
    STAGE_END();
    CKK();
    return NEMB_OK;
}

static int run_criteria(nemb_handle *h, const nemb_options *o, double beta, double *d_out)
{
    int rc;
    STAGE_BEGIN(ST_CRIT);
    size_t mine = (size_t)h->rank * h->crit_blocks * 4;
    nemk_criteria_partial(h->stream, o->k, h->row0, h->n, lpsrc(h), h->spatial ? h->d_row_ptr : NULL,
                          h->d_col, h->d_wgt, beta, o->algo == NEMB_ALGO_NCEM ? h->d_lab[h->cur] : NULL,
                          o->algo == NEMB_ALGO_NCEM ? NULL : h->d_t[h->cur], h->d_heavy, h->n_heavy,
                          h->d_crit_partials + mine, h->crit_blocks);
    if (h->world > 1 && (rc = gather(h, h->d_crit_partials + mine, h->d_crit_partials,
                                     sizeof(double) * 4 * h->crit_blocks)) != NEMB_OK) return rc;
    nemk_criteria_final(h->stream, h->crit_blocks * h->world, h->d_crit_partials, beta, d_out);
    STAGE_END();
    h->launches += 2;
    CKK();
    return NEMB_OK;
}

/* EstimBeta, BETA_PSGRAD (nem_alg.c:2120-2230): gradient ascent on the log pseudo-likelihood of
 * the current classification (the one the M-step just used).  The site sums run on the device
 * (float64, fixed two-stage order: the criteria kernel's walk); beta and its update are float on
 * the host like *BetaP.  One stream synchronisation per gradient iteration: this mode is off the
 * PPanGGOLiN path (the reference reaches it through its CLI only, nem_hlp.c:220-236). */
static int run_estim_beta(nemb_handle *h, const nemb_options *o, float *beta_io, double *sums3)
{
    int rc;
    if (!h->spatial) return NEMB_OK;                    /* TYPE_NONSPATIAL: nem_alg.c:2150-2151 */
    int nit = o->grad_n_iter > 0 ? o->grad_n_iter : 1;                 /* nem_typ.h:76 */
    float cvt = o->grad_conv > 0.f ? o->grad_conv : 0.001f;            /* nem_typ.h:77 */
    float beta = *beta_io;
    int npt = h->n_glob, conv = 0;
    size_t mine = (size_t)h->rank * h->crit_blocks * 4;
    double sums[4] = {0, 0, 0, 0};
    for (int it = 0; it < nit && !conv; it++) {
        nemk_betagrad_partial(h->stream, o->k, h->row0, h->n, h->d_row_ptr, h->d_col, h->d_wgt,
                              (double)beta, h->state_labels ? h->d_lab[h->cur] : NULL,
                              h->state_labels ? NULL : h->d_t[h->cur], h->d_heavy, h->n_heavy,
                              h->d_crit_partials + mine, h->crit_blocks);
        if (h->world > 1 && (rc = gather(h, h->d_crit_partials + mine, h->d_crit_partials,
                                         sizeof(double) * 4 * h->crit_blocks)) != NEMB_OK) return rc;
        nemk_criteria_final(h->stream, h->crit_blocks * h->world, h->d_crit_partials, (double)NAN,
                            h->d_status->crit_before);
        h->launches += 2;
        CKK();
        CK(cudaMemcpyAsync(sums, h->d_status->crit_before, sizeof sums, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        float grad = (float)sums[1], dsec = (float)sums[2];
        if (o->grad_step <= 0.0f) {                     /* nem_alg.c:2194-2200 */
            dsec = dsec * 4;
            if (dsec < (float)(npt / 10)) dsec = (float)(npt / 10);
            beta += grad / dsec;
        } else {
            beta += grad * (o->grad_step / npt);
        }
        conv = fabsf(grad) < (cvt * npt);
    }
    if (sums3) { sums3[0] = sums[0]; sums3[1] = sums[1]; sums3[2] = sums[2]; }
    if (beta > 5.0f) beta = 5.0f;                       /* MAX_BETA / MIN_BETA, nem_alg.c:84-85 */
    else if (beta < -5.0f) beta = -5.0f;
    else if (isnan(beta)) beta = 0.0f;
    *beta_io = beta;
    return NEMB_OK;
}

static int check_options(nemb_handle *h, const nemb_options *o)
{
    read_env_knobs(h);
    if (!h->loaded) return fail(h, NEMB_E_ARG, "no pangenome loaded");
    if (o->k < 1 || o->k > NEMB_MAX_K) return fail(h, NEMB_E_ARG, "k must be in 1..%d (here %d)", NEMB_MAX_K, o->k);
    if (o->algo != NEMB_ALGO_NEM && o->algo != NEMB_ALGO_NCEM) return fail(h, NEMB_E_ARG, "bad algo %d", o->algo);
    if (o->update != NEMB_UPDATE_SEQ && o->update != NEMB_UPDATE_PARA) return fail(h, NEMB_E_ARG, "bad update %d", o->update);
    if (o->conv < 0 || o->conv > 2) return fail(h, NEMB_E_ARG, "bad convergence test %d", o->conv);
    if (o->conv != NEMB_CONV_NONE && !(o->conv_thr > 0)) return fail(h, NEMB_E_ARG, "conv threshold must be > 0");
    if (o->prop < 0 || o->prop > 1 || o->disp < 0 || o->disp > 3) return fail(h, NEMB_E_ARG, "bad model");
    if (o->it_max < 0) return fail(h, NEMB_E_ARG, "it_max must be >= 0");
    if (o->beta_mode != NEMB_BETA_FIX && o->beta_mode != NEMB_BETA_PSGRAD)
        return fail(h, NEMB_E_ARG, "beta_mode %d: the heuristics are nemb_fit_beta_heuristic()", o->beta_mode);
    if (h->world > 1 && h->spatial && o->update == NEMB_UPDATE_SEQ) {
        if (o->algo == NEMB_ALGO_NEM)
            return fail(h, NEMB_E_ARG, "row shards: the sequential fuzzy sweep is single-GPU only (use update=para)");
        if (o->sweep_impl == NEMB_SWEEP_LEVEL)
            return fail(h, NEMB_E_ARG, "row shards: the level-scheduled sweep is single-GPU only");
    }
    return NEMB_OK;
}

static int init_state(nemb_handle *h, const nemb_options *o)
{
    int rc;
    h->cur = 0;
    h->state_labels = o->algo == NEMB_ALGO_NCEM;
    h->ham_valid = 0; h->stats_valid = 0; h->last_changed = -1; h->prev_valid = 0;
    h->stale_par = 0; h->sweep_same_beta = 0;
    memset(&h->mg, 0, sizeof h->mg);
    CK(cudaMemsetAsync(h->d_stale[0], 0, h->lab_len, h->stream));
    CK(cudaMemsetAsync(h->d_stale[1], 0, h->lab_len, h->stream));
    if (h->state_labels) CK(cudaMemsetAsync(h->d_lab[0], 255, h->lab_len, h->stream));
    if (h->state_labels && h->world > 1) CK(cudaMemsetAsync(h->d_lab[2], 255, h->lab_len, h->stream));   /* labels all ranks last saw */
    else {
        if ((rc = ensure_t(h, o->k, 1)) != NEMB_OK) return rc;
        CK(cudaMemsetAsync(h->d_t[0], 0, sizeof(float) * (size_t)h->lab_len * o->k, h->stream));
        CK(cudaMemsetAsync(h->d_t[1], 0, sizeof(float) * (size_t)h->lab_len * o->k, h->stream));
    }
    CK(cudaMemsetAsync(&h->d_coef->empty_class, 0, 2 * sizeof(int32_t), h->stream));   /* + halt */
    return NEMB_OK;
}

/* theta (prop[K], center[K*D], disp[K*D]) lives in one span of the per-K slab: one copy each way
 * through a pinned staging buffer instead of three pageable copies */
static int theta_stage(nemb_handle *h, size_t bytes)
{
    if (h->h_theta_cap >= bytes) return NEMB_OK;
    if (h->h_theta_stage) cudaFreeHost(h->h_theta_stage);
    h->h_theta_stage = NULL; h->h_theta_cap = 0;
    CK(cudaMallocHost((void **)&h->h_theta_stage, bytes + 256));
    h->h_theta_cap = bytes + 256;
    return NEMB_OK;
}
static int theta_h2d(nemb_handle *h, int k, const float *prop, const float *center, const float *disp)
{
    const size_t kd = (size_t)k * h->d;
    const size_t o_c = (size_t)((char *)h->d_center - (char *)h->d_prop), o_d = (size_t)((char *)h->d_disp - (char *)h->d_prop);
    const size_t span = o_d + sizeof(float) * kd;
    int rc = theta_stage(h, span);
    if (rc != NEMB_OK) return rc;
    char *st = (char *)h->h_theta_stage;
    memcpy(st, prop, sizeof(float) * k);
    memcpy(st + o_c, center, sizeof(float) * kd);
    memcpy(st + o_d, disp, sizeof(float) * kd);
    CK(cudaMemcpyAsync(h->d_prop, st, span, cudaMemcpyHostToDevice, h->stream));
    return NEMB_OK;
}
static int theta_d2h(nemb_handle *h, int k, float *prop, float *center, float *disp)
{
    const size_t kd = (size_t)k * h->d;
    const size_t o_c = (size_t)((char *)h->d_center - (char *)h->d_prop), o_d = (size_t)((char *)h->d_disp - (char *)h->d_prop);
    const size_t span = o_d + sizeof(float) * kd;
    int rc = theta_stage(h, span);
    if (rc != NEMB_OK) return rc;
    char *st = (char *)h->h_theta_stage;
    CK(cudaMemcpyAsync(st, h->d_prop, span, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    memcpy(prop, st, sizeof(float) * k);
    memcpy(center, st + o_c, sizeof(float) * kd);
    memcpy(disp, st + o_d, sizeof(float) * kd);
    return NEMB_OK;
}

/* ------------------------------------------------------------------ persistent EM kernel */
/* Which fits run as ONE cooperative launch (nem_persist.cuh): ncem on the popcount density path
 * with the per-class dispersion models, one GPU, started from theta -- PPanGGOLiN's call
 * (ppanggolin.py:1814-1826) without the per-iteration log.  Everything else keeps the
 * launch-per-stage loop of em_core. */
int nemb_i_comm_is_nccl(const nemb_comm *c);

static int persist_eligible(nemb_handle *h, const nemb_options *o, int uniform0, int has_cb,
                            const float *t_init)
{
    if (t_init || has_cb || o->dolog || h->no_persist) return 0;
    if (h->world > 1) {
        /* row shards: the ranks' kernels wait on each other through peer memory, so every rank must
         * own a GPU (NCCL communicator = one process per GPU), the shards must start on 16-family
         * boundaries (vector label accesses), and the sweep must be the sequential one */
        /* (NEM_B200_PERSIST_LOCAL=1, debugging only: the in-process test communicator too -- all its
         * ranks share ONE device, so every rank's cooperative grid gets 1/world of the CTA slots and
         * the kernels can wait on each other only if all of them are resident at once) */
        if (h->no_persist_shard || h->world > h->pk_shard_max_world) return 0;
        if (!nemb_i_comm_is_nccl(h->comm) && !h->persist_local) return 0;
        if (h->shard_len % 16 || !h->spatial || o->update != NEMB_UPDATE_SEQ) return 0;
    }
    if (o->algo != NEMB_ALGO_NCEM || o->param_fixed || o->conv == NEMB_CONV_CRIT) return 0;
    if (o->beta_mode == NEMB_BETA_PSGRAD && h->spatial) return 0;
    if (uniform0 < 1 || !(o->disp == NEMB_DISP_K_ || o->disp == NEMB_DISP___)) return 0;
    if (o->sweep_impl != NEMB_SWEEP_AUTO && o->sweep_impl != NEMB_SWEEP_SPEC) return 0;
    if (h->keep_logpf) return 0;
    return nemk_persist_max_grid(o->k) > 0;
}

/* Row shards: lay out this rank's exchange block and map every peer's (CUDA IPC).  Collective: all
 * ranks call it with the same K and pangenome.  Returns NEMB_OK with h->xblk_ok = 0 when peer
 * memory is not available on this box (the caller then keeps the launch-per-stage loop; the
 * verdict is the same on every rank). */
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static int ensure_xblk(nemb_handle *h, int k)
{
    if (h->xblk_ok && h->xblk_k == k && h->xblk_lab_len == h->lab_len) return NEMB_OK;
    const double t_begin = now_s();
    const int W = h->world;
    const size_t L = ((size_t)h->lab_len + 255) & ~(size_t)255;
    h->xcap = 1 << 16;
    h->xstat_len = (int)(((size_t)k * h->d + k + 4 + 63) & ~(size_t)63);
    size_t off = 0;
    long long *o = h->xoff;
    o[0] = (long long)carve(&off, L); o[1] = (long long)carve(&off, L);            /* lab[2] */
    o[2] = (long long)carve(&off, L); o[3] = (long long)carve(&off, L);            /* stale[2] */
    o[4] = (long long)carve(&off, sizeof(unsigned) * NEMK_PK_MAX_WORLD);           /* xflag */
    o[5] = (long long)carve(&off, sizeof(int32_t) * 2 * NEMK_PK_MAX_WORLD);        /* tot */
    o[6] = (long long)carve(&off, sizeof(int32_t) * 2 * NEMK_PK_MAX_WORLD);        /* incnt */
    o[7] = (long long)carve(&off, sizeof(int32_t) * 2 * (size_t)W * h->xcap);      /* inbox */
    o[8] = (long long)carve(&off, sizeof(int32_t) * (size_t)h->xstat_len);         /* stat (this rank's) */
    o[9] = (long long)carve(&off, sizeof(double) * (size_t)W * 8);                 /* crit */
    /* drop the mappings of a previous layout */
    for (int p = 0; p < NEMK_PK_MAX_WORLD; p++) {
        if (h->xpeer[p] && p != h->rank && !h->xpeer_local) cudaIpcCloseMemHandle(h->xpeer[p]);
        h->xpeer[p] = NULL;
    }
    h->xblk_ok = 0;
    h->xpeer_local = !nemb_i_comm_is_nccl(h->comm);
    CK(cudaStreamSynchronize(h->stream));
    if (h->b_xblk.p) { cudaFree(h->b_xblk.p); h->b_xblk.p = NULL; h->b_xblk.cap = 0; }
    CK(cudaMalloc(&h->b_xblk.p, off));
    h->b_xblk.cap = off; h->xblk_bytes = off;
    CK(cudaMemsetAsync(h->b_xblk.p, 0, off, h->stream));
    /* handles (and a go/no-go flag) travel through the communicator's all-gather */
    struct { cudaIpcMemHandle_t hd; int ok; int pad; void *raw; int pad2[12]; } mine, all[NEMK_PK_MAX_WORLD];
    const int local = !nemb_i_comm_is_nccl(h->comm);     /* one process, one device: plain pointers */
    memset(&mine, 0, sizeof mine);
    mine.raw = h->b_xblk.p;
    mine.ok = local ? 1 : cudaIpcGetMemHandle(&mine.hd, h->b_xblk.p) == cudaSuccess;
    if (!mine.ok) cudaGetLastError();
    dbuf tmp = {NULL, 0};
    int rc = reserve(h, &tmp, sizeof mine * (size_t)(W + 1));
    if (rc != NEMB_OK) return rc;
    char *d_all = tmp.p, *d_mine = d_all + sizeof mine * (size_t)h->rank;
    CK(cudaMemcpyAsync(d_mine, &mine, sizeof mine, cudaMemcpyHostToDevice, h->stream));
    if ((rc = gather(h, d_mine, d_all, sizeof mine)) != NEMB_OK) { release(&tmp); return rc; }
    CK(cudaMemcpyAsync(all, d_all, sizeof mine * (size_t)W, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    int ok = 1;
    for (int p = 0; p < W; p++) ok &= all[p].ok;
    for (int p = 0; p < W && ok; p++) {
        if (p == h->rank) { h->xpeer[p] = h->b_xblk.p; continue; }
        void *ptr = NULL;
        if (local) { h->xpeer[p] = all[p].raw; continue; }
        if (cudaIpcOpenMemHandle(&ptr, all[p].hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
        } else
            h->xpeer[p] = ptr;
    }
    /* touch every peer block once from the host side of this context: the first access through a
     * freshly opened mapping is slow (peer access is enabled lazily), and it must not happen inside
     * the kernel, where a rank would keep its peers waiting at the first cross-rank barrier */
    for (int p = 0; p < W && ok; p++) {
        if (p == h->rank) continue;
        unsigned probe = 0;
        if (cudaMemcpyAsync(&probe, h->xpeer[p] + h->xoff[4], sizeof probe, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
            cudaStreamSynchronize(h->stream) != cudaSuccess) { cudaGetLastError(); ok = 0; continue; }
        /* (NEM_B200_TOUCH_PEERS=1: every page of it from the device side as well -- atomicOr with 0,
         * the peer's data stand.  Tried against the 8-rank hang: no effect, so off by default) */
        if (getenv("NEM_B200_TOUCH_PEERS")) {
            nemk_touch_peer(h->stream, h->xpeer[p], off);
            if (cudaStreamSynchronize(h->stream) != cudaSuccess) { cudaGetLastError(); ok = 0; }
        }
    }
    /* second round: every rank must have mapped every block */
    mine.ok = ok;
    CK(cudaMemcpyAsync(d_mine, &mine, sizeof mine, cudaMemcpyHostToDevice, h->stream));
    if ((rc = gather(h, d_mine, d_all, sizeof mine)) != NEMB_OK) { release(&tmp); return rc; }
    CK(cudaMemcpyAsync(all, d_all, sizeof mine * (size_t)W, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    release(&tmp);
    for (int p = 0; p < W; p++) ok &= all[p].ok;
    h->xblk_ok = ok;
    if (getenv("NEM_B200_DEBUG_SHARD"))
        fprintf(stderr, "[nem_b200 rank %d] exchange blocks mapped: ok=%d, %.1f ms, %zu bytes\n", h->rank, ok,
                1e3 * (now_s() - t_begin), off);
    h->xblk_k = k; h->xblk_lab_len = h->lab_len;
    h->pk_xepoch = 0;
    if (!ok)
        for (int p = 0; p < W; p++) {
            if (h->xpeer[p] && p != h->rank && !h->xpeer_local) cudaIpcCloseMemHandle(h->xpeer[p]);
            h->xpeer[p] = NULL;
        }
    return NEMB_OK;
}

static int ensure_persist(nemb_handle *h)
{
    if (!h->pk_out) {
        CK(cudaHostAlloc((void **)&h->pk_out, sizeof(nemk_persist_out), cudaHostAllocMapped | cudaHostAllocPortable));
        memset(h->pk_out, 0, sizeof(nemk_persist_out));
        CK(cudaHostGetDevicePointer((void **)&h->d_pk_out, h->pk_out, 0));
    }
    size_t off = 0;
    size_t cap = (size_t)(h->nnz > h->n ? h->nnz : h->n) + 64;   /* work-list entries (duplicates allowed) */
    size_t o_wl0 = carve(&off, sizeof(int32_t) * cap), o_wl1 = carve(&off, sizeof(int32_t) * cap);
    size_t o_evf = carve(&off, (size_t)h->n + 64);
    size_t o_hub = carve(&off, sizeof(int32_t) * 4);
    size_t o_scr = carve(&off, sizeof(int32_t) * 64);      /* [16..23]: heartbeat of the row-sharded kernel */
    size_t o_cnt = carve(&off, sizeof(nemk_counters) * 2);
    size_t o_bar = carve(&off, sizeof(unsigned) * 4);
    size_t o_crit = carve(&off, sizeof(double) * 4 * 2048);
    size_t o_outc = carve(&off, sizeof(int32_t) * 16);
    int fresh = h->b_pk.cap < off;
    int rc = reserve(h, &h->b_pk, off);
    if (rc != NEMB_OK) return rc;
    char *base = h->b_pk.p;
    h->d_pk_hub = (int32_t *)(base + o_hub);
    h->d_pk_wl[0] = (int32_t *)(base + o_wl0); h->d_pk_wl[1] = (int32_t *)(base + o_wl1);
    h->d_pk_evflag = (uint8_t *)(base + o_evf);
    h->pk_wl_cap = (int)(cap > 0x7fffffff ? 0x7fffffff : cap);
    int32_t *scr = (int32_t *)(base + o_scr);
    nemk_counters *cnt2 = (nemk_counters *)(base + o_cnt);
    unsigned *bar = (unsigned *)(base + o_bar);
    if (fresh || scr != h->d_pk_scratch || cnt2 != h->d_pk_cnt2 || bar != h->d_pk_bar) {
        /* zero at rest: the kernel leaves its counters, lists and the barrier clean */
        CK(cudaMemsetAsync(base + o_scr, 0, o_crit - o_scr, h->stream));
        h->pk_cnt_par = 0;
    }
    h->d_pk_scratch = scr; h->d_pk_cnt2 = cnt2; h->d_pk_bar = bar;
    h->d_pk_crit = (double *)(base + o_crit);
    h->d_pk_out_cnt = (int32_t *)(base + o_outc);
    return NEMB_OK;
}

static int wait_persist(nemb_handle *h, unsigned long long seq)
{
    volatile nemk_persist_out *s = h->pk_out;
    const double t_wait = h->world > 1 ? now_s() : 0.0;
    int reported = 0;
    for (unsigned spins = 1; s->seq != seq; spins++) {
        if ((spins & 0xfff) == 0) {
            if (h->world > 1 && !reported && now_s() - t_wait > 3.0) {
                /* a row-sharded kernel that has not come back after 3 s: say where it is (the
                 * heartbeat words of its scratch block, read on a side stream while it runs) */
                int32_t hb[24];
                cudaStream_t side = NULL;
                reported = 1;
                if (cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) == cudaSuccess) {
                    if (cudaMemcpyAsync(hb, h->d_pk_scratch, sizeof hb, cudaMemcpyDeviceToHost, side) == cudaSuccess &&
                        cudaStreamSynchronize(side) == cudaSuccess)
                        fprintf(stderr, "[nem_b200 rank %d] persistent kernel still running after 3 s: round %d, items %d, "
                                "source %d, cascade epoch %d, barrier entered %d / passed %d, timeout word %d\n", h->rank,
                                hb[16], hb[17], hb[18], hb[19], hb[20], hb[21], hb[13]);
                    cudaStreamDestroy(side);
                }
                cudaGetLastError();
            }
            cudaError_t e = cudaStreamQuery(h->stream);
            if (e != cudaSuccess && e != cudaErrorNotReady)
                return fail(h, NEMB_E_CUDA, "device error inside the persistent EM kernel: %s", cudaGetErrorString(e));
            if (e == cudaSuccess && s->seq != seq)
                return fail(h, NEMB_E_BUG, "persistent EM kernel ended without status %llu (slot holds %llu)", seq,
                            (unsigned long long)s->seq);
        }
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
        if (h->poll_relaxed && spins > 64) {
            struct timespec ts = {0, 2000};
            nanosleep(&ts, NULL);
        } else if ((spins & 0xff) == 0) sched_yield();
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    return NEMB_OK;
}

static int run_density(nemb_handle *h, int k, int uniform, int32_t *d_hamming, int lean);

/* the whole fit up to (not including) the final criteria; fills iter / converged / status / empty
 * the way em_core's loop does */
static int em_persist(nemb_handle *h, const nemb_options *o, int uniform0, nemb_result *res,
                      int *iter_out, int *converged_out, int *status_out, int *empty_out)
{
    int k = o->k, rc;
    const float betaf = h->spatial ? o->beta : 0.0f;          /* nem_exe.c:570-574 */
    const size_t kd = (size_t)k * h->d;
    if ((rc = ensure_persist(h)) != NEMB_OK) return rc;
    nemk_persist_args a;
    memset(&a, 0, sizeof a);
    a.world = 1;
    a.evflag = h->d_pk_evflag;
    a.margin = h->d_margin;
    a.lab[0] = h->d_lab[0]; a.lab[1] = h->d_lab[1];
    a.stale[0] = h->d_stale[0]; a.stale[1] = h->d_stale[1];
    if (h->world > 1) {
        /* labels and stale flags live in the exchange block the peers store into; margins, Hamming
         * counts and evaluation flags stay rank-local, addressed by GLOBAL family id (-row0) */
        char *blk = h->b_xblk.p;
        a.world = h->world; a.rank = h->rank; a.row0 = h->row0; a.shard_len = h->shard_len;
        a.n_glob = h->n_glob; a.xcap = h->xcap; a.xepoch = h->pk_xepoch;
        for (int p = 0; p < h->world; p++) a.peer[p] = h->xpeer[p];
        a.off_lab[0] = h->xoff[0]; a.off_lab[1] = h->xoff[1]; a.off_stale[0] = h->xoff[2]; a.off_stale[1] = h->xoff[3];
        a.off_xflag = h->xoff[4]; a.off_tot = h->xoff[5]; a.off_incnt = h->xoff[6]; a.off_inbox = h->xoff[7];
        a.off_stat = h->xoff[8]; a.off_crit = h->xoff[9];
        a.stat_len = h->xstat_len; a.out_cnt = h->d_pk_out_cnt; a.stat_glob = h->d_stat_int;
        /* this rank's statistics live in its block (the peers read them after the M-step barrier) */
        h->d_stat_loc = (int32_t *)(blk + h->xoff[8]);
        h->d_lab[0] = (uint8_t *)(blk + h->xoff[0]); h->d_lab[1] = (uint8_t *)(blk + h->xoff[1]);
        a.lab[0] = h->d_lab[0]; a.lab[1] = h->d_lab[1];
        a.stale[0] = (uint8_t *)(blk + h->xoff[2]); a.stale[1] = (uint8_t *)(blk + h->xoff[3]);
        a.margin = h->d_margin - h->row0;
        a.evflag = h->d_pk_evflag - h->row0;
    }
    a.K = k; a.n = h->n; a.D = h->d; a.wpr = h->wpr; a.nwt = h->nwt;
    a.prop_model = o->prop; a.disp_model = o->disp; a.conv = o->conv; a.it_max = o->it_max;
    a.conv_thr = o->conv_thr;
    a.beta = (double)betaf;
    a.use_graph = h->spatial && a.beta != 0.0;
    a.seq_sweep = a.use_graph && o->update == NEMB_UPDATE_SEQ;
    a.n_heavy = h->n_heavy; a.wsum_any_order = h->wgt_integral;
    a.use_margins = !h->no_margins;
    size_t xbytes = sizeof(uint32_t) * (size_t)h->n * h->wpr;
    a.x_in_kernel = xbytes <= h->pk_xlimit && h->world == 1;
    a.init_from_pop = uniform0 == 2 && !h->no_popcache;
    if (a.init_from_pop && (rc = ensure_pop(h)) != NEMB_OK) return rc;
    if (a.x_in_kernel && (rc = ensure_xt(h)) != NEMB_OK) return rc;
    a.x = h->d_x; a.xt = h->d_xt; a.pop = h->d_pop;
    a.row_ptr = h->d_row_ptr; a.col = h->d_col; a.rrow_ptr = h->d_rrow_ptr; a.rcol = h->d_rcol;
    a.heavy = h->d_heavy; a.wgt = h->d_wgt;
    a.prop = h->d_prop; a.center = h->d_center; a.disp = h->d_disp; a.coef = h->d_coef;
    a.mxor = h->d_mxor; a.mval = h->d_mval; a.f0 = h->d_f0; a.f1 = h->d_f1; a.cm = h->d_cm;
    a.delta = h->d_delta; a.ham = h->d_ham; a.stat = h->d_stat_loc;
    a.dirty = h->d_dirty; a.wl[0] = h->d_wl[0]; a.wl[1] = h->d_wl[1];
    a.wl_cnt = h->d_wl_counts;
    a.wlist[0] = h->d_pk_wl[0]; a.wlist[1] = h->d_pk_wl[1]; a.wl_cap = h->pk_wl_cap;
    a.hub_list = h->d_pk_hub; a.scratch = h->d_pk_scratch; a.cnt2 = h->d_pk_cnt2; a.bar = h->d_pk_bar;
    a.out = h->d_pk_out;
    a.crit_partials = h->d_pk_crit; a.want_crit = 1; a.spatial = h->spatial;
    a.no_shortcuts = h->no_shortcuts;
    h->pk_have_crit = 0;
    /* host state of a fresh fit (the kernel's prep phase writes the device side) */
    h->cur = 0; h->state_labels = 1;
    h->ham_valid = 0; h->stats_valid = 0; h->last_changed = -1; h->prev_valid = 0;
    h->stale_par = 0; h->sweep_same_beta = 0; h->lp_from_ham = 1;
    memset(&h->mg, 0, sizeof h->mg);
    a.entry = NEMK_PK_ENTRY_INIT; a.iter0 = 0; a.cur = 0; a.stale_par = 0; a.stats_valid = 0;
    a.last_changed = -1; a.margins_on = 0;
    int grid = nemk_persist_max_grid(k);
    int want = (h->n + 511) / 512;
    if (want < 16) want = 16;
    if (h->pk_grid_limit > 0 && want > h->pk_grid_limit) want = h->pk_grid_limit;
    if (h->pk_grid_env > 0) want = h->pk_grid_env;
    if (h->world > 1 && h->xpeer_local && want > grid / h->world) want = grid / h->world;   /* ranks share the device */
    if (grid > want) grid = want;
    if (grid > 2048) grid = 2048;      /* d_pk_crit rows */
    if (grid < 1) grid = 1;
    if (h->world > 1 && h->xpeer_local) {
        /* one-device debugging mode: the ranks' kernels can only meet if they are resident together,
         * so the host threads line up first (a collective on the test communicator + a drained stream) */
        if ((rc = gather(h, &h->d_status->cnt, h->d_cnt_all, sizeof(nemk_counters))) != NEMB_OK) return rc;
        CK(cudaStreamSynchronize(h->stream));
    }
    int done = 0, guard = 0;
    memset(h->pk_phase_ns, 0, sizeof h->pk_phase_ns);
    nemk_persist_out out;
    memset(&out, 0, sizeof out);
    while (!done) {
        if (++guard > 4 * (o->it_max + 4)) return fail(h, NEMB_E_BUG, "persistent EM kernel keeps leaving");
        a.cnt_par = h->pk_cnt_par;
        a.seq = ++h->pk_seq;
        const double t_launch = now_s();
        nemk_persist_launch(h->stream, &a, grid);
        h->launches++;
        res->pk_launches++;
        CKK();
        if ((rc = wait_persist(h, a.seq)) != NEMB_OK) return rc;
        memcpy(&out, (const void *)h->pk_out, sizeof out);
        if (h->world > 1 && getenv("NEM_B200_DEBUG_SHARD"))
            fprintf(stderr, "[nem_b200 rank %d] launch %d entry %d: %.2f ms, exit %d, epoch %u, %d barriers\n", h->rank,
                    res->pk_launches, a.entry, 1e3 * (now_s() - t_launch), out.exit_code, out.xepoch, out.barriers);
        h->pk_cnt_par = out.cnt_par;
        /* the kernel counts SM cycles: convert with the SM clock (kHz) */
        if (!h->sm_khz) { int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->device); h->sm_khz = khz > 0 ? khz : 1965000; }
        for (int q = 0; q < 10; q++) h->pk_phase_ns[q] += out.phase_ns[q] * 1000000ull / (unsigned long long)h->sm_khz;
        h->pk_phase_ns[10] += out.phase_ns[10];
        memcpy(h->pk_trace, out.trace, sizeof h->pk_trace);
        for (int r = 0; r < 12; r++)
            for (int q = 0; q < 6; q++) h->pk_trace[r][q] = h->pk_trace[r][q] * 1000000ll / (long long)h->sm_khz;
        h->cur = out.cur; h->stale_par = out.stale_par; h->last_changed = out.last_changed;
        h->stats_valid = out.stats_valid; h->prev_valid = out.sweeps > 0 || a.entry != NEMK_PK_ENTRY_INIT;
        h->ham_valid = 1;
        h->fixup_rounds += out.fixup_rounds;
        res->n_kept += out.kept;
        res->pk_barriers += out.barriers; res->pk_x_passes += out.x_passes; res->pk_recounts += out.recounts;
        res->n_allnul = out.n_allnul; res->n_ties = out.n_ties;
        h->pk_xepoch = out.xepoch;
        if (out.exit_code == NEMK_PK_EXIT_PEER_TIMEOUT || out.xerror)
            return fail(h, NEMB_E_CUDA, "row-sharded fit: rank %d waited for a peer that never arrived (epoch %u, "
                        "iteration %d, %d sweeps and %d barriers into launch %d, entry %d; missing rank %d whose "
                        "epoch flag here read %d)", h->rank, out.xepoch,
                        out.iters, out.sweeps, out.barriers, res->pk_launches, a.entry,
                        (out.xerror >> 4) & 15, (out.xerror >> 8) & 0xfffff);
        if (out.exit_code == NEMK_PK_EXIT_DONE) { done = 1; break; }
        a.entry = out.resume_entry; a.iter0 = out.iters; a.cur = out.cur; a.stale_par = out.stale_par;
        a.stats_valid = out.stats_valid; a.last_changed = out.last_changed;
        a.delta_mode = out.delta_mode; a.n_allnul = out.n_allnul; a.n_ties = out.n_ties;
        a.flags_stale = out.flags_stale; a.mu_changed = out.mu_changed;
        a.xepoch = out.xepoch; a.decide_pending = out.decide_pending; a.chg_local = out.chg_local;
        a.margins_on = out.sweeps > 0 || a.margins_on;
        if (out.exit_code == NEMK_PK_EXIT_NEED_DENSITY) {
            /* the class masks moved and X does not fit the L2: the TMA-tiled X pass (HBM-bound) */
            if ((rc = run_density(h, k, 1, NULL, 1)) != NEMB_OK) return rc;
            res->pk_x_passes++;
        } else if (out.exit_code == NEMK_PK_EXIT_NEED_RECOUNT) {
            /* full recount of S = X^T T through the transposed bits (HBM-bound) */
            if ((rc = ensure_xt(h)) != NEMB_OK) return rc;
            STAGE_BEGIN(ST_MSTEP);
            nemk_label_masks(h->stream, k, h->n, h->nwt, h->d_lab[h->cur] + h->row0, h->d_cm, h->d_stat_loc + kd,
                             NULL, &h->d_coef->halt);
            nemk_mstep_ncem(h->stream, k, h->d, h->nwt, h->d_xt, h->d_cm, h->d_stat_loc, &h->d_coef->halt);
            CK(cudaMemsetAsync(&h->d_coef->uniform_ok, 1, sizeof(int32_t), h->stream));
            CK(cudaMemsetAsync(&h->d_coef->mu_changed, 0, sizeof(int32_t), h->stream));
            STAGE_END();
            h->launches += 2;
            res->pk_recounts++;
            CKK();
        } else
            return fail(h, NEMB_E_BUG, "persistent EM kernel: unknown exit code %d", out.exit_code);
    }
    h->pk_have_crit = out.have_crit;
    memcpy(h->pk_crit, out.crit, sizeof h->pk_crit);
    h->sweep_same_beta = 1;
    /* the kernel's incremental statistics follow its own convention (they may already include the
     * last sweep's moves): a launch-per-stage M-step after it must recount */
    h->stats_valid = 0;
    *iter_out = out.iters; *converged_out = out.converged; *empty_out = out.empty_class;
    *status_out = out.empty_class ? NEMB_W_EMPTYCLASS : NEMB_OK;
    return NEMB_OK;
}

/* EM from the theta already resident on the device (d_prop/d_center/d_disp). */
static int upload_state(nemb_handle *h, const nemb_options *o, const float *t);

/* t_init != NULL: INIT_FILE (nem_alg.c:1091-1113) -- NemAlgo starts from that classification, the
 * two initial sweeps are not run and theta comes from the first M-step. */
/* CK / CKK inside em_core leave through its cleanup label */
#define CKO(call)                                                                          \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            rc = fail(h, e_ == cudaErrorMemoryAllocation ? NEMB_E_MEMORY : NEMB_E_CUDA,    \
                      "%s: %s", #call, cudaGetErrorString(e_));                            \
            goto out;                                                                      \
        }                                                                                  \
    } while (0)
static int em_core(nemb_handle *h, const nemb_options *o, int uniform0, nemb_result *res,
                   nemb_iter_cb cb, void *user, float *prop, float *center, float *disp,
                   const float *t_init)
{
    int k = o->k, rc, flipped = 0;
    float betaf = h->spatial ? o->beta : 0.0f;          /* nem_exe.c:570-574 */
    double beta = (double)betaf;
    double swept_beta = NAN;                            /* beta of the previous sweep */
    const int psgrad = o->beta_mode == NEMB_BETA_PSGRAD && h->spatial;
    int uniform_m = o->param_fixed ? uniform0 : (o->disp == NEMB_DISP_K_ || o->disp == NEMB_DISP___);   /* 1: theta lives on the device */
    int want_crit_each = o->dolog || o->conv == NEMB_CONV_CRIT;
    int lean = o->algo == NEMB_ALGO_NCEM && !h->keep_logpf;
    size_t kd = (size_t)k * h->d;
    float *nk_host = cb ? malloc(sizeof(float) * k) : NULL;
    int iter = 0, enq = 0, converged = 0, status = NEMB_OK, empty = 0;
    int flips[2] = {0, 0};
    unsigned long long seqs[2] = {0, 0};
    double oldcrit = 0.0;
    int can_spec;
    /* one cooperative launch for the whole fit where it applies (nem_persist.cuh) */
    int persist = lean && persist_eligible(h, o, uniform0, cb != NULL, t_init);
    if (cb && !nk_host) { rc = fail(h, NEMB_E_MEMORY, "host memory"); goto out; }
    if (persist && h->world > 1) {
        /* the peers' exchange blocks (CUDA IPC); without peer memory every rank falls back alike */
        if ((rc = ensure_xblk(h, k)) != NEMB_OK) goto out;
        if (!h->xblk_ok) persist = 0;
    }
    if (persist) {
        if ((rc = em_persist(h, o, uniform0, res, &iter, &converged, &status, &empty)) != NEMB_OK) goto out;
        goto tail;
    }

    if ((rc = init_state(h, o)) != NEMB_OK) goto out;
    if (t_init) {
        if ((rc = upload_state(h, o, t_init)) != NEMB_OK) goto out;
    } else {
    if ((rc = run_tables(h, k, 0)) != NEMB_OK) goto out;
    if ((rc = run_density(h, k, uniform0, NULL, lean)) != NEMB_OK) goto out;
    /* ComputePartitionFromPara(Needinit=1): blind sweep then beta sweep (nem_alg.c:1970-1981) */
    if ((rc = run_sweep(h, o, 0.0, &flipped, NULL, NULL)) != NEMB_OK) goto out;
    if (o->dolog && (rc = run_criteria(h, o, beta, h->d_status->crit_before)) != NEMB_OK) goto out;
    if ((rc = run_sweep(h, o, beta, &flipped, NULL, NULL)) != NEMB_OK) goto out;
    if (o->dolog && (rc = run_criteria(h, o, beta, h->d_status->crit_after)) != NEMB_OK) goto out;
    swept_beta = beta;
    }
    if (!t_init && (o->dolog || cb || o->it_max == 0)) {
        if ((rc = read_status(h)) != NEMB_OK) goto out;
        h->fixup_rounds += h->h_status->cnt.nfix;
        res->n_allnul = h->h_status->cnt.allnul;
        res->n_ties = h->h_status->cnt.ties;
        oldcrit = h->h_status->crit_after[3];
        if (cb) {
            for (int c = 0; c < k; c++) nk_host[c] = 0.0f;   /* NbObs_KD not estimated yet: calloc zeros */
            cb(user, 0, h->h_status->crit_before, h->h_status->crit_after, prop, center, disp, nk_host);
        }
    }

    /* EM loop.  One iteration = M-step, densities, sweep, then nemk_iter_end: the convergence
     * test runs on the device and the status lands in mapped host memory, so the host neither
     * copies nor synchronises -- and once the fit is in its steady state (incremental M-step) the
     * NEXT iteration is enqueued before this one's status is known.  If this one ends the fit, the
     * device halt flag turns every kernel of the speculative iteration into a no-op and its
     * host-side effects (buffer flip) are undone below. */
    can_spec = !o->dolog && !cb && !want_crit_each && !h->profile && h->world == 1 && !psgrad &&
                   o->algo == NEMB_ALGO_NCEM && !o->param_fixed && !h->no_spec;
    while (iter < o->it_max && !converged && status == NEMB_OK) {
        int want = iter + 1;
        if (can_spec && want < o->it_max && h->stats_valid && h->last_changed >= 0 &&
            h->last_changed <= h->n / 8)
            want++;
        for (; enq < want; enq++) {
            if (!o->param_fixed && (rc = run_mstep(h, o, uniform_m)) != NEMB_OK) goto out;
            if (psgrad) {            /* EstimBeta follows the M-step (nem_alg.c:1810-1812) */
                if ((rc = run_estim_beta(h, o, &betaf, NULL)) != NEMB_OK) goto out;
                beta = (double)betaf;
            }
            if ((rc = run_density(h, k, uniform_m, NULL, lean)) != NEMB_OK) goto out;
            if (o->dolog && (rc = run_criteria(h, o, beta, h->d_status->crit_before)) != NEMB_OK) goto out;
            int status_read = 0;
            h->sweep_same_beta = swept_beta == beta;   /* the previous sweep used this beta: cached margins may apply */
            swept_beta = beta;
            if ((rc = run_sweep(h, o, beta, &flipped, (want_crit_each || cb) ? NULL : o, &status_read)) != NEMB_OK) goto out;
            if (want_crit_each && (rc = run_criteria(h, o, beta, h->d_status->crit_after)) != NEMB_OK) goto out;
            if (cb) {
                CKO(cudaMemcpyAsync(prop, h->d_prop, sizeof(float) * k, cudaMemcpyDeviceToHost, h->stream));
                CKO(cudaMemcpyAsync(center, h->d_center, sizeof(float) * kd, cudaMemcpyDeviceToHost, h->stream));
                CKO(cudaMemcpyAsync(disp, h->d_disp, sizeof(float) * kd, cudaMemcpyDeviceToHost, h->stream));
            }
            flips[enq & 1] = flipped;
            if (status_read && !want_crit_each && !cb) seqs[enq & 1] = h->seq;   /* the sweep's last status */
            else if ((rc = publish_status(h, o, &seqs[enq & 1])) != NEMB_OK) goto out;
        }
        if ((rc = wait_status(h, seqs[iter & 1])) != NEMB_OK) goto out;
        if (cb) CKO(cudaStreamSynchronize(h->stream));   /* theta copies of this iteration */
        flipped = flips[iter & 1];
        iter++;
        h->last_changed = o->algo == NEMB_ALGO_NCEM ? h->h_status->cnt.changed : -1;
        empty = *h->h_empty;
        if (empty) {              /* nem_alg.c:1831-1838: E-step not run, loop ends */
            status = NEMB_W_EMPTYCLASS;
            if (flipped) h->cur ^= 1;
            continue;
        }
        h->fixup_rounds += h->h_status->cnt.nfix;
        res->n_allnul = h->h_status->cnt.allnul;
        res->n_ties = h->h_status->cnt.ties;
        res->n_kept += h->h_status->cnt.kept;
        if (o->conv == NEMB_CONV_CLAS) {
            float md = o->algo == NEMB_ALGO_NCEM ? (h->h_status->cnt.changed ? 1.0f : 0.0f)
                                                 : h->h_status->cnt.maxdiff;
            converged = md < o->conv_thr;
        } else if (o->conv == NEMB_CONV_CRIT) {
            double curc = h->h_status->crit_after[3];
            float dif = curc != 0 ? (float)fabs((curc - oldcrit) / curc) : FLT_MAX;
            converged = dif < o->conv_thr;
        }
        if (want_crit_each) oldcrit = h->h_status->crit_after[3];
        if (cb) {
            if (o->param_fixed) {
                for (int c = 0; c < k; c++) nk_host[c] = 0.0f;   /* NbObs_KD not estimated yet: calloc zeros */
            } else if (o->algo == NEMB_ALGO_NCEM) {
                int32_t ni[NEMB_MAX_K];
                CKO(cudaMemcpy(ni, (h->world > 1 ? h->d_stat_int : h->d_stat_loc) + kd, sizeof(int32_t) * k, cudaMemcpyDeviceToHost));
                for (int c = 0; c < k; c++) nk_host[c] = (float)ni[c];
            } else {
                double nd[NEMB_MAX_K];
                CKO(cudaMemcpy(nd, h->d_stat_dbl + kd, sizeof(double) * k, cudaMemcpyDeviceToHost));
                for (int c = 0; c < k; c++) nk_host[c] = (float)nd[c];
            }
            cb(user, iter, h->h_status->crit_before, h->h_status->crit_after, prop, center, disp, nk_host);
        }
    }
    if (enq > iter) {
        /* the speculative iteration behind the last one: a no-op on the device if the fit ended
         * (halt), otherwise (it_max cannot be hit here: want < it_max) it must not exist */
        if (!(converged || status != NEMB_OK)) { rc = fail(h, NEMB_E_BUG, "speculative iteration left over"); goto out; }
        if (flips[(enq - 1) & 1]) h->cur ^= 1;
    }
tail:
    if (h->world > 1 && !persist && status == NEMB_OK && iter > 0) {
        /* row shards: the per-iteration statuses of the speculative sweep carry this rank's
         * all-null / tie counters only; one counter all-gather gives the totals of the last sweep */
        if ((rc = read_status(h)) != NEMB_OK) goto out;
        res->n_allnul = h->h_status->cnt.allnul;
        res->n_ties = h->h_status->cnt.ties;
    }
    res->beta = betaf;
    if (t_init && status == NEMB_W_EMPTYCLASS && iter == 1) {
        /* MakeParaFromLabeled: "Class %d has no labeled observation" (nem_alg.c:1338-1345) --
         * NemAlgo is not entered */
        res->status = status; res->iters = 0; res->empty_class = empty;
        rc = NEMB_OK;
        goto out;
    }
    if (iter == 0) { /* nem_alg.c:1845-1851 */
        if ((rc = run_mstep(h, o, uniform_m)) != NEMB_OK) goto out;
        if ((rc = run_density(h, k, uniform_m, NULL, lean)) != NEMB_OK) goto out;
    }
    if (persist && h->pk_have_crit) {
        /* the kernel evaluated the criteria and reported the empty class: theta comes back in ONE
         * copy through pinned memory */
        if ((rc = theta_d2h(h, k, prop, center, disp)) != NEMB_OK) goto out;
        memcpy(h->h_status->crit_after, h->pk_crit, sizeof h->pk_crit);
        *h->h_empty = empty;
    } else {
    if ((rc = run_criteria(h, o, beta, h->d_status->crit_after)) != NEMB_OK) goto out;
    CKO(cudaMemcpyAsync(prop, h->d_prop, sizeof(float) * k, cudaMemcpyDeviceToHost, h->stream));
    CKO(cudaMemcpyAsync(center, h->d_center, sizeof(float) * kd, cudaMemcpyDeviceToHost, h->stream));
    CKO(cudaMemcpyAsync(disp, h->d_disp, sizeof(float) * kd, cudaMemcpyDeviceToHost, h->stream));
    CKO(cudaMemcpyAsync(h->h_status, h->d_status, sizeof(iter_status), cudaMemcpyDeviceToHost, h->stream));
    CKO(cudaMemcpyAsync(h->h_empty, &h->d_coef->empty_class, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CKO(cudaStreamSynchronize(h->stream));
    }
    if (iter == 0 && *h->h_empty) { status = NEMB_W_EMPTYCLASS; empty = *h->h_empty; }
    res->status = status; res->iters = iter; res->converged = converged; res->empty_class = empty;
    const double *c6 = h->h_status->crit_after;
    res->U = c6[0]; res->D = c6[1]; res->L = c6[2]; res->M = c6[3]; res->Z = c6[4]; res->G = c6[5];
    rc = NEMB_OK;
out:
    free(nk_host);
    return rc;
}

static void collect_profile(nemb_handle *h, nemb_result *res, cudaEvent_t e0, cudaEvent_t e1)
{
    cudaEventElapsedTime(&res->fit_ms, e0, e1);
    res->kernel_launches = h->launches;
    res->fixup_rounds = h->fixup_rounds;
    res->exchanges = h->exchanges;
    if (!h->profile) return;
    float *ms[ST_NB] = {&res->ms_density, &res->ms_sweep, &res->ms_mstep, &res->ms_criteria,
                        &res->ms_density_cached, &res->ms_mstep_delta};
    int32_t *cn[ST_NB] = {&res->n_density, &res->n_sweep, &res->n_mstep, &res->n_criteria,
                          &res->n_density_cached, &res->n_mstep_delta};
    for (int i = 0; i + 1 < h->ev_n; i++) {
        int kind = h->ev_kind[i];
        if (kind < 0 || h->ev_kind[i + 1] != -1) continue;
        float t = 0.f;
        cudaEventElapsedTime(&t, h->ev[i], h->ev[i + 1]);
        *ms[kind] += t;
        *cn[kind] += 1;
    }
}

static int fit_common(nemb_handle *h, const nemb_options *o, float *prop, float *center, float *disp,
                      nemb_result *res, nemb_iter_cb cb, void *user, const float *t_init)
{
    if (!h || !o || !prop || !center || !disp || !res) return NEMB_E_ARG;
    int rc;
    CK(cudaSetDevice(h->device));
    if ((rc = check_options(h, o)) != NEMB_OK) return rc;
    if (t_init) {
        if ((rc = need_single(h, "nemb_fit_from_partition")) != NEMB_OK) return rc;
        if (o->param_fixed) return fail(h, NEMB_E_ARG, "a starting classification needs the M-step (param_fixed = 0)");
        if (o->algo == NEMB_ALGO_NCEM) {      /* hardened rows (nem_alg.c:2385-2391) */
            for (int i = 0; i < h->n; i++) {
                int ones = 0, other = 0;
                for (int c = 0; c < o->k; c++) {
                    float v = t_init[(size_t)i * o->k + c];
                    if (v == 1.0f) ones++; else if (v != 0.0f) other++;
                }
                if (ones != 1 || other) return fail(h, NEMB_E_ARG, "ncem: row %d of the starting classification is not one-hot", i);
            }
        }
    }
    if ((rc = ensure_k(h, o->k)) != NEMB_OK) return rc;
    memset(res, 0, sizeof *res);
    h->launches = 0; h->fixup_rounds = 0; h->exchanges = 0; h->profile = o->profile; h->ev_n = 0;
    h->ev_last_density = h->ev_last_cached = -1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, h->stream));
    int uniform0 = 0;
    if (!t_init) {
        if ((rc = theta_h2d(h, o->k, prop, center, disp)) != NEMB_OK) return rc;
        uniform0 = theta_uniform(o->k, h->d, center, disp);
    }
    rc = em_core(h, o, uniform0, res, cb, user, prop, center, disp, t_init);
    cudaEventRecord(e1, h->stream);
    cudaStreamSynchronize(h->stream);
    if (rc == NEMB_OK) collect_profile(h, res, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (rc != NEMB_OK) return rc;
    return res->status;
}

int nemb_fit_logged(nemb_handle *h, const nemb_options *o, float *prop, float *center, float *disp,
                    nemb_result *res, nemb_iter_cb cb, void *user)
{
    return fit_common(h, o, prop, center, disp, res, cb, user, NULL);
}

int nemb_fit(nemb_handle *h, const nemb_options *o, float *prop, float *center, float *disp,
             nemb_result *res)
{
    return fit_common(h, o, prop, center, disp, res, NULL, NULL, NULL);
}

int nemb_fit_from_partition(nemb_handle *h, const nemb_options *o, const float *t_init, float *prop,
                            float *center, float *disp, nemb_result *res)
{
    if (!t_init) return NEMB_E_ARG;
    return fit_common(h, o, prop, center, disp, res, NULL, NULL, t_init);
}

/* ClassifyByNemHeuBeta (nem_alg.c:731-992).  Host control flow over complete fits; the float
 * arithmetic on the criteria (criV, slopes, thresholds) repeats the reference's. */
int nemb_fit_beta_heuristic(nemb_handle *h, const nemb_options *o, int mode,
                            const nemb_beta_heuristic *hp0, float *prop, float *center, float *disp,
                            nemb_result *res, float *beta_trace, float *crit_trace, int cap)
{
    if (!h || !o || !prop || !center || !disp || !res) return NEMB_E_ARG;
    if (mode != NEMB_BETA_HEUD && mode != NEMB_BETA_HEUL) return fail(h, NEMB_E_ARG, "heuristic mode must be heu_d or heu_l");
    if (o->beta_mode != NEMB_BETA_FIX) return fail(h, NEMB_E_ARG, "a beta heuristic cannot be combined with psgrad");
    if (o->param_fixed) return fail(h, NEMB_E_ARG, "the beta heuristics need the M-step (param_fixed = 0)");
    int rc;
    if ((rc = need_single(h, "nemb_fit_beta_heuristic")) != NEMB_OK) return rc;
    nemb_beta_heuristic hp = {0.1f, 2.0f, 0.8f, 0.5f, 0.02f};            /* nem_typ.h:71-75 */
    if (hp0) {
        if (hp0->step > 0) hp.step = hp0->step;
        if (hp0->max > 0) hp.max = hp0->max;
        if (hp0->ddrop > 0) hp.ddrop = hp0->ddrop;
        if (hp0->dloss > 0) hp.dloss = hp0->dloss;
        if (hp0->lloss > 0) hp.lloss = hp0->lloss;
    }
    const int n = h->n, k = o->k;
    const size_t nk = (size_t)n * k;
    int nbtamax = (int)(hp.max / hp.step) + 1;
    if (nbtamax < 1 || nbtamax > (1 << 20)) return fail(h, NEMB_E_ARG, "heuristic: bad step/max");
    float *btaV = calloc((size_t)nbtamax + 2, sizeof(float)), *criV = calloc((size_t)nbtamax + 2, sizeof(float));
    float *best = malloc(sizeof(float) * nk);
    if (!btaV || !criV || !best) { free(btaV); free(criV); free(best); return fail(h, NEMB_E_MEMORY, "heuristic: host memory"); }
    int nbta = 0, stop = 0, Dincreas = 0, Ddrop = 0, Lfound = 0, have_best = 0;
    float Dmin = 0.0f, prevSlope = NAN, thisSlope = NAN, Lmax = NAN, btaEst = NAN;
    const float DdropThres = -hp.ddrop * n, LlossThres = hp.lloss * n;
    nemb_options ob = *o;
    int64_t launches = 0, fixups = 0, kept = 0;
    float ms = 0.f;
    rc = NEMB_OK;
#define SAVE_BEST() do { if ((rc = nemb_get_posteriors(h, best)) != NEMB_OK) goto done; have_best = 1; } while (0)
    for (float bt = 0.0f; bt <= hp.max && !stop; bt += hp.step) {
        ob.beta = bt;
        int frc = fit_common(h, &ob, prop, center, disp, res, NULL, NULL, NULL);
        launches += res->kernel_launches; fixups += res->fixup_rounds; kept += res->n_kept; ms += res->fit_ms;
        if (frc != NEMB_OK) {                /* an empty class skips this beta (nem_alg.c:836-838) */
            if (frc != NEMB_W_EMPTYCLASS) { rc = frc; goto done; }
            continue;
        }
        if (nbta > nbtamax) break;
        nbta++;
        btaV[nbta] = bt;
        if (mode == NEMB_BETA_HEUD) {
            criV[nbta] = (float)res->D;
            if (criV[nbta] < Dmin) Dmin = criV[nbta];
            if (nbta >= 2) {
                prevSlope = thisSlope;
                thisSlope = (criV[nbta] - criV[nbta - 1]) / (btaV[nbta] - btaV[nbta - 1]);
                if (thisSlope >= 0.5 * n) Dincreas = 1;
            }
            if (nbta >= 3) {
                if (!Ddrop && !Dincreas) {
                    if ((thisSlope - prevSlope) < DdropThres) { Ddrop = 1; stop = 1; btaEst = btaV[nbta - 1]; }
                    else SAVE_BEST();
                }
            } else SAVE_BEST();
        } else {
            criV[nbta] = (float)res->L;
            if (nbta < 2) { Lmax = criV[nbta]; SAVE_BEST(); }
            else {
                if (criV[nbta] > Lmax) Lmax = criV[nbta];
                if (!Lfound) {
                    if (criV[nbta] < Lmax - LlossThres) { Lfound = 1; stop = 1; btaEst = btaV[nbta - 1]; }
                    else SAVE_BEST();
                }
            }
        }
    }
#undef SAVE_BEST
    {
        int from_partition;
        if (mode == NEMB_BETA_HEUD && !Ddrop) {        /* loss thresholding (nem_alg.c:931-954) */
            float DThres = criV[1] - (criV[1] - Dmin) * hp.dloss;
            int ibta, found = 0;
            for (ibta = 1; ibta <= nbta && !found; ibta++) found = criV[ibta] <= DThres;
            btaEst = found ? btaV[ibta - 2] : 0.0f;    /* "heuristic failed to detect beta" */
            from_partition = 0;
        } else
            from_partition = have_best;                /* restore the saved classification, INIT_FILE */
        if (mode == NEMB_BETA_HEUL && !Lfound) btaEst = btaV[nbta];
        if (isnan(btaEst)) btaEst = 0.0f;              /* every tested beta failed */
        for (int i = 0; i < nbta && i < cap; i++) {
            if (beta_trace) beta_trace[i] = btaV[i + 1];
            if (crit_trace) crit_trace[i] = criV[i + 1];
        }
        ob.beta = btaEst;
        rc = fit_common(h, &ob, prop, center, disp, res, NULL, NULL, from_partition ? best : NULL);
        res->kernel_launches += launches; res->fixup_rounds += fixups; res->n_kept += kept; res->fit_ms += ms;
        res->n_beta_tested = nbta;
    }
done:
    free(btaV); free(criV); free(best);
    return rc;
}

int nemb_stage_estim_beta(nemb_handle *h, const nemb_options *o, const float *t, float *beta_io,
                          double *sums3)
{
    if (!h || !o || !t || !beta_io) return NEMB_E_ARG;
    int rc;
    CK(cudaSetDevice(h->device));
    if ((rc = check_options(h, o)) != NEMB_OK) return rc;
    if ((rc = ensure_k(h, o->k)) != NEMB_OK) return rc;
    h->profile = 0;
    if ((rc = upload_state(h, o, t)) != NEMB_OK) return rc;
    return run_estim_beta(h, o, beta_io, sums3);
}

/* nanoseconds CTA 0 of the persistent EM kernel spent per phase during the last fit (nem_device.h
 * nemk_persist_out.phase_ns; entry 10 = fix-up rounds executed) */
int nemb_get_persist_profile(nemb_handle *h, unsigned long long out12[12])
{
    if (!h || !out12) return NEMB_E_ARG;
    memcpy(out12, h->pk_phase_ns, sizeof h->pk_phase_ns);
    return NEMB_OK;
}

/* per-iteration trace of the last launch of the persistent kernel (nemk_persist_out.trace) */
int nemb_get_persist_trace(nemb_handle *h, long long out96[96])
{
    if (!h || !out96) return NEMB_E_ARG;
    memcpy(out96, h->pk_trace, sizeof h->pk_trace);
    return NEMB_OK;
}

/* ------------------------------------------------------------------ results */
int nemb_get_posteriors(nemb_handle *h, float *t_out)
{
    if (!h || !t_out || !h->k_alloc) return NEMB_E_ARG;
    CK(cudaSetDevice(h->device));
    int k = h->k_alloc, rc;
    float *src = h->d_t[h->cur];
    if (h->state_labels) {
        if ((rc = ensure_t(h, k, 0)) != NEMB_OK) return rc;
        src = h->d_t[0];
        nemk_labels_to_t(h->stream, k, h->n_glob, h->d_lab[h->cur], src);
        CKK();
    }
    CK(cudaMemcpyAsync(t_out, src, sizeof(float) * (size_t)h->n_glob * k, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return NEMB_OK;
}

int nemb_get_labels_rows(nemb_handle *h, int first, int count, int32_t *label_out)
{
    if (!h || !label_out || !h->k_alloc) return NEMB_E_ARG;
    if (first < 0 || count < 0 || (int64_t)first + count > h->n_glob)
        return fail(h, NEMB_E_ARG, "label rows [%d, %d) outside 0..%d", first, first + count, h->n_glob);
    CK(cudaSetDevice(h->device));
    uint8_t *src = h->d_lab[h->cur];
    int n = h->n_glob;
    if (!h->state_labels) {
        src = h->d_lab[0];
        nemk_t_to_labels(h->stream, h->k_alloc, n, h->d_t[h->cur], src);
        CKK();
    }
    if (count == 0) return NEMB_OK;
    if (h->h_lab_cap < (size_t)count) {      /* pinned staging, grow-only */
        if (h->h_lab_stage) cudaFreeHost(h->h_lab_stage);
        h->h_lab_stage = NULL; h->h_lab_cap = 0;
        CK(cudaMallocHost((void **)&h->h_lab_stage, (size_t)count + 64));
        h->h_lab_cap = (size_t)count + 64;
    }
    const uint8_t *tmp = h->h_lab_stage;
    CK(cudaMemcpyAsync(h->h_lab_stage, src + first, (size_t)count, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < count; i++) label_out[i] = tmp[i] == 255 ? -1 : (int32_t)tmp[i];
    return NEMB_OK;
}

int nemb_get_labels(nemb_handle *h, int32_t *label_out)
{
    if (!h) return NEMB_E_ARG;
    return nemb_get_labels_rows(h, 0, h->n_glob, label_out);
}

/* n = rows this rank holds; depth = Gauss-Seidel DAG depth (builds the level schedule on first
 * request; 0 on row shards and for non-spatial data) */
int nemb_get_dims(const nemb_handle *hc, int *n, int *d, int *wpr, int *nwt, int *depth, int *nnz)
{
    nemb_handle *h = (nemb_handle *)hc;
    if (!h || !h->loaded) return NEMB_E_ARG;
    if (depth && h->spatial && h->world == 1 && !h->have_levels) {
        if (cudaSetDevice(h->device) != cudaSuccess) return NEMB_E_CUDA;
        int rc = ensure_levels(h);
        if (rc != NEMB_OK) return rc;
    }
    if (n) *n = h->n;
    if (d) *d = h->d;
    if (wpr) *wpr = h->wpr;
    if (nwt) *nwt = h->nwt;
    if (depth) *depth = h->depth;
    if (nnz) *nnz = h->nnz;
    return NEMB_OK;
}

int nemb_get_packed(nemb_handle *h, uint32_t *out)
{
    if (!h || !h->loaded || !out) return NEMB_E_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(out, h->d_x, sizeof(uint32_t) * (size_t)h->n * h->wpr, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return NEMB_OK;
}

int nemb_get_graph(nemb_handle *h, int32_t *row_ptr, int32_t *col, float *wgt)
{
    if (!h || !h->loaded) return NEMB_E_ARG;
    if (!h->spatial) return fail(h, NEMB_E_ARG, "no neighbour graph loaded");
    CK(cudaSetDevice(h->device));
    if (row_ptr) CK(cudaMemcpyAsync(row_ptr, h->d_row_ptr, sizeof(int32_t) * ((size_t)h->n_glob + 1), cudaMemcpyDeviceToHost, h->stream));
    if (col && h->nnz) CK(cudaMemcpyAsync(col, h->d_col, sizeof(int32_t) * (size_t)h->nnz, cudaMemcpyDeviceToHost, h->stream));
    if (wgt && h->nnz) CK(cudaMemcpyAsync(wgt, h->d_wgt, sizeof(float) * (size_t)h->nnz, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return NEMB_OK;
}

int nemb_get_transposed(nemb_handle *h, uint32_t *out)
{
    if (!h || !h->loaded || !out) return NEMB_E_ARG;
    CK(cudaSetDevice(h->device));
    int rc = ensure_xt(h);
    if (rc != NEMB_OK) return rc;
    CK(cudaMemcpyAsync(out, h->d_xt, sizeof(uint32_t) * (size_t)h->d * h->nwt, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return NEMB_OK;
}

int nemb_get_levels(nemb_handle *h, int32_t *level)
{
    if (!h || !h->loaded || !level) return NEMB_E_ARG;
    if (!h->spatial) { for (int i = 0; i < h->n; i++) level[i] = 1; return NEMB_OK; }
    CK(cudaSetDevice(h->device));
    int rc = ensure_levels(h);
    if (rc != NEMB_OK) return rc;
    memcpy(level, h->h_level, sizeof(int32_t) * (size_t)h->n);
    return NEMB_OK;
}

/* ------------------------------------------------------------------ stage entry points */
static int upload_theta(nemb_handle *h, int k, const float *prop, const float *center, const float *disp)
{
    size_t kd = (size_t)k * h->d;
    CK(cudaMemcpyAsync(h->d_prop, prop, sizeof(float) * k, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_center, center, sizeof(float) * kd, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_disp, disp, sizeof(float) * kd, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(&h->d_coef->empty_class, 0, 2 * sizeof(int32_t), h->stream));   /* + halt */
    return NEMB_OK;
}

int nemb_stage_density(nemb_handle *h, int k, const float *prop, const float *center,
                       const float *disp, int force_general, double *logpf_out,
                       int32_t *hamming_out, int *used_uniform)
{
    if (!h || !h->loaded || !prop || !center || !disp || !logpf_out) return NEMB_E_ARG;
    if (k < 1 || k > NEMB_MAX_K) return fail(h, NEMB_E_ARG, "bad k");
    int rc;
    if ((rc = need_single(h, "nemb_stage_density")) != NEMB_OK) return rc;
    CK(cudaSetDevice(h->device));
    if ((rc = ensure_k(h, k)) != NEMB_OK) return rc;
    h->profile = 0;
    if ((rc = upload_theta(h, k, prop, center, disp)) != NEMB_OK) return rc;
    int uniform = !force_general && theta_uniform(k, h->d, center, disp);
    int32_t *d_ham = NULL;
    size_t nk = (size_t)h->n * k;
    if (hamming_out && uniform) CK(cudaMalloc((void **)&d_ham, sizeof(int32_t) * nk));
    h->ham_valid = 0;
    if ((rc = run_tables(h, k, 0)) != NEMB_OK) return rc;
    if ((rc = run_density(h, k, uniform, d_ham, 0)) != NEMB_OK) return rc;
    CK(cudaMemcpyAsync(logpf_out, h->d_logpf, sizeof(double) * nk, cudaMemcpyDeviceToHost, h->stream));
    if (d_ham) CK(cudaMemcpyAsync(hamming_out, d_ham, sizeof(int32_t) * nk, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(d_ham);
    if (used_uniform) *used_uniform = uniform;
    return NEMB_OK;
}

static int upload_state(nemb_handle *h, const nemb_options *o, const float *t)
{
    size_t nk = (size_t)h->n * o->k;
    int rc;
    if ((rc = need_single(h, "stage entry points")) != NEMB_OK) return rc;
    if ((rc = ensure_t(h, o->k, 1)) != NEMB_OK) return rc;
    h->cur = 0;
    h->state_labels = o->algo == NEMB_ALGO_NCEM;
    h->stats_valid = 0; h->ham_valid = 0; h->last_changed = -1;
    CK(cudaMemcpyAsync(h->d_t[h->state_labels ? 1 : 0], t, sizeof(float) * nk, cudaMemcpyHostToDevice, h->stream));
    if (h->state_labels) {
        nemk_t_to_labels(h->stream, o->k, h->n, h->d_t[1], h->d_lab[0]);
        CKK();
    }
    return NEMB_OK;
}

int nemb_stage_sweep(nemb_handle *h, const nemb_options *o, const double *logpf, float beta,
                     float *t_inout, int32_t *label_out, int64_t *fixup_rounds)
{
    if (!h || !o || !logpf || !t_inout) return NEMB_E_ARG;
    int rc, flipped;
    CK(cudaSetDevice(h->device));
    if ((rc = check_options(h, o)) != NEMB_OK) return rc;
    if ((rc = ensure_k(h, o->k)) != NEMB_OK) return rc;
    h->profile = 0;
    size_t nk = (size_t)h->n * o->k;
    CK(cudaMemcpyAsync(h->d_logpf, logpf, sizeof(double) * nk, cudaMemcpyHostToDevice, h->stream));
    h->lp_from_ham = 0;
    CK(cudaMemsetAsync(&h->d_coef->empty_class, 0, 2 * sizeof(int32_t), h->stream));   /* + halt */
    if ((rc = upload_state(h, o, t_inout)) != NEMB_OK) return rc;
    if ((rc = run_sweep(h, o, h->spatial ? (double)beta : 0.0, &flipped, NULL, NULL)) != NEMB_OK) return rc;
    if ((rc = read_status(h)) != NEMB_OK) return rc;
    if (fixup_rounds) *fixup_rounds = h->h_status->cnt.nfix;
    if ((rc = nemb_get_posteriors(h, t_inout)) != NEMB_OK) return rc;
    if (label_out && (rc = nemb_get_labels(h, label_out)) != NEMB_OK) return rc;
    return NEMB_OK;
}

int nemb_stage_mstep(nemb_handle *h, const nemb_options *o, const float *t, float *prop,
                     float *center, float *disp, double *nk_out, double *skd_out, int *empty_class)
{
    if (!h || !o || !t || !prop || !center || !disp) return NEMB_E_ARG;
    int rc, k = o->k;
    CK(cudaSetDevice(h->device));
    if ((rc = check_options(h, o)) != NEMB_OK) return rc;
    if ((rc = ensure_k(h, k)) != NEMB_OK) return rc;
    h->profile = 0;
    size_t kd = (size_t)k * h->d;
    if ((rc = upload_theta(h, k, prop, center, disp)) != NEMB_OK) return rc;
    if ((rc = upload_state(h, o, t)) != NEMB_OK) return rc;
    if ((rc = run_mstep(h, o, 0)) != NEMB_OK) return rc;
    CK(cudaMemcpyAsync(prop, h->d_prop, sizeof(float) * k, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(center, h->d_center, sizeof(float) * kd, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(disp, h->d_disp, sizeof(float) * kd, cudaMemcpyDeviceToHost, h->stream));
    if ((rc = read_status(h)) != NEMB_OK) return rc;
    if (empty_class) *empty_class = *h->h_empty;
    if (nk_out || skd_out) {
        if (o->algo == NEMB_ALGO_NCEM) {
            int32_t *ti = malloc(sizeof(int32_t) * (kd + k));
            CK(cudaMemcpy(ti, h->d_stat_loc, sizeof(int32_t) * (kd + k), cudaMemcpyDeviceToHost));
            if (skd_out) for (size_t q = 0; q < kd; q++) skd_out[q] = ti[q];
            if (nk_out) for (int c = 0; c < k; c++) nk_out[c] = ti[kd + c];
            free(ti);
        } else {
            if (skd_out) CK(cudaMemcpy(skd_out, h->d_stat_dbl, sizeof(double) * kd, cudaMemcpyDeviceToHost));
            if (nk_out) CK(cudaMemcpy(nk_out, h->d_stat_dbl + kd, sizeof(double) * k, cudaMemcpyDeviceToHost));
        }
    }
    return NEMB_OK;
}

int nemb_stage_criteria(nemb_handle *h, const nemb_options *o, const double *logpf, const float *t,
                        float beta, double *crit6)
{
    if (!h || !o || !logpf || !t || !crit6) return NEMB_E_ARG;
    int rc;
    CK(cudaSetDevice(h->device));
    if ((rc = check_options(h, o)) != NEMB_OK) return rc;
    if ((rc = ensure_k(h, o->k)) != NEMB_OK) return rc;
    h->profile = 0;
    size_t nk = (size_t)h->n * o->k;
    CK(cudaMemcpyAsync(h->d_logpf, logpf, sizeof(double) * nk, cudaMemcpyHostToDevice, h->stream));
    h->lp_from_ham = 0;
    if ((rc = upload_state(h, o, t)) != NEMB_OK) return rc;
    if ((rc = run_criteria(h, o, h->spatial ? (double)beta : 0.0, h->d_status->crit_after)) != NEMB_OK) return rc;
    if ((rc = read_status(h)) != NEMB_OK) return rc;
    memcpy(crit6, h->h_status->crit_after, sizeof(double) * 6);
    return NEMB_OK;
}

/* ------------------------------------------------------------------ init_mode 1: random starts */
/* RandNemAlgo (nem_alg.c:1574-1742): InitPara's whole-sample dispersion (1247-1265: an M-step
 * with every family in class 1), then n_starts x { MakeRandomPara (1381-1473): eps = sample
 * dispersion / K, p = 1/K, centres = distinct random families ; blind+beta sweeps ; NemAlgo },
 * keep the start with the best criterion M (DEFAULT_CRIT, nem_typ.h:80), final M-step on the
 * best partition (1715).  The reference draws with libc random() seeded by the wall clock
 * (nem_exe.c:353,621); we use a documented splitmix64 stream, so parity is statistical. */
static uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

/* ------------------------------------------------------------------ random starts (init_mode 1)
 * RandNemAlgo (nem_alg.c:1574-1742) = InitPara (1200-1280: dispersion of the whole sample from an
 * all-in-class-1 M-step) + n_starts x { MakeRandomPara (1381-1473: centres = distinct random data
 * rows, dispersion = sample dispersion / K, equal proportions); fit } + keep the best start by
 * criterion M (first maximum) + a final EstimPara on the best partition.  The reference draws from
 * libc random() seeded with time(NULL); here start s has its own splitmix64 stream derived from
 * (seed, s), so the starts are reproducible and independent of the order they run in -- which lets
 * several worker streams fit them concurrently. */
int nemb_sample_dispersion(nemb_handle *h, const nemb_options *o, float *disp_sample)
{
    if (!h || !o || !disp_sample) return NEMB_E_ARG;
    int rc, k = o->k;
    CK(cudaSetDevice(h->device));
    if ((rc = check_options(h, o)) != NEMB_OK) return rc;
    if ((rc = need_single(h, "nemb_sample_dispersion")) != NEMB_OK) return rc;
    if ((rc = ensure_k(h, k)) != NEMB_OK) return rc;
    if (o->algo != NEMB_ALGO_NCEM && (rc = ensure_t(h, k, 1)) != NEMB_OK) return rc;
    size_t kd = (size_t)k * h->d;
    float *theta = malloc(sizeof(float) * (k + 2 * kd));
    if (!theta) return fail(h, NEMB_E_MEMORY, "host alloc");
    for (int c = 0; c < k; c++) theta[c] = (float)(1.0 / k);
    for (size_t q = 0; q < 2 * kd; q++) theta[k + q] = 0.5f;
    rc = upload_theta(h, k, theta, theta + k, theta + k + kd);
    free(theta);
    if (rc != NEMB_OK) return rc;
    h->cur = 0; h->state_labels = o->algo == NEMB_ALGO_NCEM;
    CK(cudaMemsetAsync(h->d_lab[0], 0, h->n, h->stream));          /* every family in class 1 */
    if (!h->state_labels) { nemk_labels_to_t(h->stream, k, h->n, h->d_lab[0], h->d_t[0]); CKK(); }
    h->stats_valid = 0; h->ham_valid = 0; h->prev_valid = 0; h->last_changed = -1;
    CK(cudaMemsetAsync(&h->d_coef->empty_class, 0, 2 * sizeof(int32_t), h->stream));   /* + halt */
    if ((rc = run_mstep(h, o, 0)) != NEMB_OK) return rc;
    CK(cudaMemcpyAsync(disp_sample, h->d_disp, sizeof(float) * h->d, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return NEMB_OK;
}

int nemb_random_start(nemb_handle *h, int k, int64_t seed, int start, const float *disp_sample,
                      float *prop, float *center, float *disp)
{
    if (!h || !h->loaded || !disp_sample || !prop || !center || !disp || k < 1 || k > NEMB_MAX_K || start < 0)
        return NEMB_E_ARG;
    CK(cudaSetDevice(h->device));
    int n = h->n, d = h->d, wpr = h->wpr;
    uint64_t rng = (uint64_t)(seed ? seed : 42) + 0x632be59bd9b4e019ULL * (uint64_t)(start + 1);
    uint32_t *rows = malloc(sizeof(uint32_t) * (size_t)k * wpr);
    if (!rows) return fail(h, NEMB_E_MEMORY, "host alloc");
    for (int c = 0; c < k; c++) {
        prop[c] = (float)(1.0 / k);
        for (int j = 0; j < d; j++) disp[(size_t)c * d + j] = disp_sample[j] / (float)k;
        for (int draw = 0, again = 1; again && draw < 100; draw++) {
            int ipt = (int)(splitmix64(&rng) % (uint64_t)n);
            cudaError_t e = cudaMemcpyAsync(rows + (size_t)c * wpr, h->d_x + (size_t)ipt * wpr,
                                            sizeof(uint32_t) * wpr, cudaMemcpyDeviceToHost, h->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) { free(rows); return fail(h, NEMB_E_CUDA, "%s", cudaGetErrorString(e)); }
            again = 0;   /* drawn again while identical to a previous centre (nem_alg.c:1421-1449) */
            for (int p = 0; p < c && !again; p++)
                if (!memcmp(rows + (size_t)p * wpr, rows + (size_t)c * wpr, sizeof(uint32_t) * wpr))
                    again = 1;
        }
        for (int j = 0; j < d; j++)
            center[(size_t)c * d + j] = (float)((rows[(size_t)c * wpr + (j >> 5)] >> (j & 31)) & 1u);
    }
    free(rows);
    return NEMB_OK;
}

/* a second handle on the SAME resident pangenome (X, X^T, graph, hub list are borrowed, never
 * freed or written through `w`); its per-fit state is its own */
static void attach_problem(nemb_handle *w, const nemb_handle *src)
{
    reset_problem(w);
    w->n = src->n; w->n_glob = src->n_glob; w->row0 = src->row0; w->shard_len = src->shard_len;
    w->lab_len = src->lab_len; w->d = src->d; w->wpr = src->wpr; w->nwt = src->nwt;
    w->nnz = src->nnz; w->spatial = src->spatial; w->symmetric = src->symmetric;
    w->max_neigh = src->max_neigh; w->wgt_integral = src->wgt_integral;
    w->d_x = src->d_x; w->x_owned = 0;
    w->d_xt = src->d_xt; w->have_xt = src->have_xt;
    w->d_pop = src->d_pop; w->have_pop = src->have_pop;
    w->d_row_ptr = src->d_row_ptr; w->d_col = src->d_col; w->d_wgt = src->d_wgt;
    w->d_rrow_ptr = src->d_rrow_ptr; w->d_rcol = src->d_rcol;
    w->d_heavy = src->d_heavy; w->n_heavy = src->n_heavy;
    w->loaded = 1;
}

typedef struct {
    nemb_handle *src;
    const nemb_options *opt;
    const float *sam;
    int64_t seed;
    int n_starts, next, failed_rc;
    size_t state_bytes;
    pthread_mutex_t mu;
    /* per worker results */
    struct rs_best { int start, nsucc, last_status; nemb_result res; float *theta; void *state; nemb_handle *w; } *best;
    int64_t launches;
    char err[256];
} rs_ctx;

/* fits the starts it pulls from ctx->next on handle w; keeps its best (first maximum of M) */
static int rs_run(rs_ctx *c, nemb_handle *w, struct rs_best *b)
{
    const nemb_options *o = c->opt;
    int k = o->k, d = w->d, rc = NEMB_OK;
    size_t kd = (size_t)k * d;
    float *theta = malloc(sizeof(float) * (k + 2 * kd));
    if (!theta) return NEMB_E_MEMORY;
    b->start = -1; b->nsucc = 0; b->last_status = NEMB_W_EMPTYCLASS;
    for (;;) {
        int s = __atomic_fetch_add(&c->next, 1, __ATOMIC_RELAXED);
        if (s >= c->n_starts) break;
        float *prop = theta, *center = theta + k, *disp = theta + k + kd;
        if ((rc = nemb_random_start(w, k, c->seed, s, c->sam, prop, center, disp)) != NEMB_OK) break;
        nemb_options oo = *o;
        oo.profile = 0;
        nemb_result cur;
        int frc = nemb_fit(w, &oo, prop, center, disp, &cur);
        if (frc != NEMB_OK && frc != NEMB_W_EMPTYCLASS) { rc = frc; break; }
        __atomic_fetch_add(&c->launches, cur.kernel_launches, __ATOMIC_RELAXED);
        b->last_status = cur.status;
        if (cur.status != NEMB_OK) continue;
        b->nsucc++;
        /* workers pull starts in increasing order: a strict > keeps the first maximum */
        if (b->start < 0 || cur.M > b->res.M) {
            b->res = cur; b->start = s;
            memcpy(b->theta, theta, sizeof(float) * (k + 2 * kd));
            cudaError_t e = cudaMemcpyAsync(b->state, o->algo == NEMB_ALGO_NCEM ? (void *)w->d_lab[w->cur]
                                                                                : (void *)w->d_t[w->cur],
                                            c->state_bytes, cudaMemcpyDeviceToDevice, w->stream);
            if (e != cudaSuccess) { rc = NEMB_E_CUDA; break; }
        }
    }
    cudaStreamSynchronize(w->stream);
    free(theta);
    return rc;
}

static void *rs_worker(void *arg)
{
    rs_ctx *c = ((void **)arg)[0];
    struct rs_best *b = ((void **)arg)[1];
    int rc = rs_run(c, b->w, b);
    if (rc != NEMB_OK) {
        pthread_mutex_lock(&c->mu);
        if (c->failed_rc == NEMB_OK) { c->failed_rc = rc; snprintf(c->err, sizeof c->err, "%s", nemb_last_error(b->w)); }
        pthread_mutex_unlock(&c->mu);
    }
    return NULL;
}

int nemb_fit_random_workers(nemb_handle *h, const nemb_options *o, int n_starts, int64_t seed,
                            int n_workers, float *prop, float *center, float *disp, nemb_result *res)
{
    if (!h || !o || !prop || !center || !disp || !res) return NEMB_E_ARG;
    int rc, k = o->k;
    CK(cudaSetDevice(h->device));
    if ((rc = check_options(h, o)) != NEMB_OK) return rc;
    if ((rc = need_single(h, "nemb_fit_random")) != NEMB_OK) return rc;
    if (n_starts <= 0) n_starts = 50;                       /* DEFAULT_NBRANDINITS, nem_typ.h:94 */
    if (n_workers < 1) n_workers = 1;
    if (n_workers > 16) n_workers = 16;
    if (n_workers > n_starts) n_workers = n_starts;
    int n = h->n, d = h->d;
    size_t kd = (size_t)k * d;
    memset(res, 0, sizeof *res);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, h->stream));
    float *sam = malloc(sizeof(float) * d);
    if (!sam) return fail(h, NEMB_E_MEMORY, "host alloc");
    if ((rc = nemb_sample_dispersion(h, o, sam)) != NEMB_OK) { free(sam); return rc; }
    /* the first full M-step built X^T; the workers borrow it with the rest of the problem */
    rs_ctx c;
    memset(&c, 0, sizeof c);
    c.src = h; c.opt = o; c.sam = sam; c.seed = seed; c.n_starts = n_starts; c.failed_rc = NEMB_OK;
    c.state_bytes = o->algo == NEMB_ALGO_NCEM ? (size_t)n : sizeof(float) * (size_t)n * k;
    pthread_mutex_init(&c.mu, NULL);
    c.best = calloc((size_t)n_workers, sizeof *c.best);
    void *args[16][2];
    pthread_t th[16];
    int made = 0;
    for (int w = 0; w < n_workers && rc == NEMB_OK; w++, made++) {
        struct rs_best *b = &c.best[w];
        b->theta = malloc(sizeof(float) * (k + 2 * kd));
        if (cudaMalloc(&b->state, c.state_bytes) != cudaSuccess || !b->theta) { rc = fail(h, NEMB_E_MEMORY, "alloc"); made++; break; }
        if (w == 0) b->w = h;                  /* the caller's handle fits starts too */
        else {
            if ((rc = nemb_create(&b->w, h->device)) != NEMB_OK) { made++; break; }
            attach_problem(b->w, h);
        }
    }
    if (rc == NEMB_OK) {
        CK(cudaStreamSynchronize(h->stream));
        int started = 0;
        for (int w = 1; w < n_workers; w++) {
            args[w][0] = &c; args[w][1] = &c.best[w];
            if (pthread_create(&th[w], NULL, rs_worker, args[w]) == 0) started |= 1 << w;
        }
        int r0 = rs_run(&c, h, &c.best[0]);
        for (int w = 1; w < n_workers; w++) if (started & (1 << w)) pthread_join(th[w], NULL);
        rc = r0 != NEMB_OK ? r0 : c.failed_rc;
        if (rc != NEMB_OK && r0 == NEMB_OK) fail(h, rc, "random-start worker: %s", c.err);
    }
    int best_w = -1, nsucc = 0, last_status = NEMB_W_EMPTYCLASS;
    if (rc == NEMB_OK) {
        for (int w = 0; w < n_workers; w++) {
            struct rs_best *b = &c.best[w];
            nsucc += b->nsucc;
            if (b->nsucc == 0) { if (b->last_status != NEMB_W_EMPTYCLASS) last_status = b->last_status; continue; }
            /* global first maximum: larger M, or the same M from an earlier start */
            if (best_w < 0 || b->res.M > c.best[best_w].res.M ||
                (b->res.M == c.best[best_w].res.M && b->start < c.best[best_w].start))
                best_w = w;
        }
        if (best_w >= 0) {
            struct rs_best *b = &c.best[best_w];
            memcpy(prop, b->theta, sizeof(float) * k);
            memcpy(center, b->theta + k, sizeof(float) * kd);
            memcpy(disp, b->theta + k + kd, sizeof(float) * kd);
            rc = ensure_k(h, k);
            if (rc == NEMB_OK && o->algo != NEMB_ALGO_NCEM) rc = ensure_t(h, k, 1);
            if (rc == NEMB_OK) rc = upload_theta(h, k, prop, center, disp);
            if (rc == NEMB_OK) {
                h->cur = 0; h->state_labels = o->algo == NEMB_ALGO_NCEM;
                cudaMemcpyAsync(o->algo == NEMB_ALGO_NCEM ? (void *)h->d_lab[0] : (void *)h->d_t[0], b->state,
                                c.state_bytes, cudaMemcpyDeviceToDevice, h->stream);
                h->stats_valid = 0; h->ham_valid = 0; h->prev_valid = 0; h->last_changed = -1;
                cudaMemsetAsync(&h->d_coef->empty_class, 0, 2 * sizeof(int32_t), h->stream);   /* + halt */
                rc = run_mstep(h, o, 0);   /* final EstimPara on the best partition, nem_alg.c:1715 */
            }
            if (rc == NEMB_OK) {
                cudaMemcpyAsync(prop, h->d_prop, sizeof(float) * k, cudaMemcpyDeviceToHost, h->stream);
                cudaMemcpyAsync(center, h->d_center, sizeof(float) * kd, cudaMemcpyDeviceToHost, h->stream);
                cudaMemcpyAsync(disp, h->d_disp, sizeof(float) * kd, cudaMemcpyDeviceToHost, h->stream);
                *res = b->res;
                res->status = NEMB_OK;
                res->best_start = b->start + 1;
                res->n_success = nsucc;
            }
        } else {
            res->status = last_status;
        }
    }
    cudaEventRecord(e1, h->stream);
    cudaStreamSynchronize(h->stream);
    cudaEventElapsedTime(&res->fit_ms, e0, e1);
    res->kernel_launches = c.launches;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    for (int w = 0; w < made && w < n_workers; w++) {
        struct rs_best *b = &c.best[w];
        if (b->state) cudaFree(b->state);
        free(b->theta);
        if (w > 0 && b->w) nemb_destroy(b->w);
    }
    free(c.best); free(sam);
    pthread_mutex_destroy(&c.mu);
    if (rc != NEMB_OK) return rc;
    return res->status;
}

int nemb_fit_random(nemb_handle *h, const nemb_options *o, int n_starts, int64_t seed, float *prop,
                    float *center, float *disp, nemb_result *res)
{
    const char *e = getenv("NEM_B200_RANDOM_WORKERS");
    return nemb_fit_random_workers(h, o, n_starts, seed, e ? atoi(e) : 4, prop, center, disp, res);
}
