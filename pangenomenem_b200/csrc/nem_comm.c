/*
 * nem_comm.c -- the two communicators shipped with the engine for row-sharded fits
 * (include/nem_b200.h "Row-sharded fits"):
 *
 *   nccl   one rank per process/GPU; libnccl.so.2 is dlopen'ed so that the library itself has no
 *          link-time dependency on NCCL (single-GPU users never load it).  The only primitive
 *          the engine needs is ncclAllGather on the engine's stream.
 *   local  `world` ranks = `world` host threads driving `world` handles on ONE device: each
 *          rank copies its block into a shared staging buffer on its own stream, events order
 *          the copies, host barriers order the ranks.  No kernel ever waits on another rank's
 *          kernel, so the ranks may run serialised on the one GPU (B200_PROFILING.md rule).
 *          Used by the single-GPU parity tests of the sharded path.
 *
 * The reference has no counterpart (its engine is one single-threaded process).
 */
#define _GNU_SOURCE
#include "nem_b200.h"

#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ NCCL (dlopen) */
typedef struct { char internal[128]; } nccl_uid;
typedef void *nccl_comm_t;
typedef int (*fn_get_uid)(nccl_uid *);
typedef int (*fn_init_rank)(nccl_comm_t *, int, nccl_uid, int);
typedef int (*fn_allgather)(const void *, void *, size_t, int /*dtype*/, nccl_comm_t, cudaStream_t);
typedef int (*fn_destroy)(nccl_comm_t);
typedef const char *(*fn_errstr)(int);

static struct {
    void *lib;
    fn_get_uid get_uid;
    fn_init_rank init_rank;
    fn_allgather allgather;
    fn_destroy destroy;
    fn_errstr errstr;
} g_nccl;

static int nccl_load(void)
{
    if (g_nccl.lib) return 0;
    const char *cands[] = {getenv("NEM_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so", NULL};
    void *lib = NULL;
    for (int i = 0; i < 4 && !lib; i++)
        if (cands[i] && *cands[i]) lib = dlopen(cands[i], RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
        fprintf(stderr, "nem_b200: cannot load libnccl.so.2 (%s); set NEM_B200_NCCL_LIB\n", dlerror());
        return -1;
    }
    g_nccl.get_uid = (fn_get_uid)dlsym(lib, "ncclGetUniqueId");
    g_nccl.init_rank = (fn_init_rank)dlsym(lib, "ncclCommInitRank");
    g_nccl.allgather = (fn_allgather)dlsym(lib, "ncclAllGather");
    g_nccl.destroy = (fn_destroy)dlsym(lib, "ncclCommDestroy");
    g_nccl.errstr = (fn_errstr)dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.allgather || !g_nccl.destroy) {
        fprintf(stderr, "nem_b200: libnccl lacks a required symbol\n");
        dlclose(lib);
        return -1;
    }
    g_nccl.lib = lib;
    return 0;
}

typedef struct { nccl_comm_t comm; } nccl_ctx;

static int nccl_allgather_cb(void *ctx, const void *send, void *recv, size_t bytes, void *stream)
{
    nccl_ctx *c = ctx;
    int rc = g_nccl.allgather(send, recv, bytes, 0 /* ncclInt8 / ncclChar */, c->comm, (cudaStream_t)stream);
    if (rc != 0)
        fprintf(stderr, "nem_b200: ncclAllGather: %s\n", g_nccl.errstr ? g_nccl.errstr(rc) : "error");
    return rc;
}

static void nccl_destroy_cb(void *ctx)
{
    nccl_ctx *c = ctx;
    if (c && c->comm) g_nccl.destroy(c->comm);
    free(c);
}

int nemb_nccl_unique_id(uint8_t id_out[128])
{
    if (!id_out) return NEMB_E_ARG;
    if (nccl_load()) return NEMB_E_CUDA;
    nccl_uid uid;
    memset(&uid, 0, sizeof uid);
    if (g_nccl.get_uid(&uid) != 0) return NEMB_E_CUDA;
    memcpy(id_out, &uid, 128);
    return NEMB_OK;
}

/* one process and one GPU per rank: kernels of different ranks may wait on each other (the
 * persistent row-sharded EM kernel); the in-process test double runs its ranks on ONE device, where
 * they may be serialised, and must never do that */
int nemb_i_comm_is_nccl(const nemb_comm *c) { return c && c->allgather == nccl_allgather_cb; }

int nemb_comm_create_nccl(nemb_comm **out, const uint8_t id[128], int rank, int world)
{
    if (!out || !id || world < 1 || rank < 0 || rank >= world) return NEMB_E_ARG;
    *out = NULL;
    if (nccl_load()) return NEMB_E_CUDA;
    nemb_comm *c = calloc(1, sizeof *c);
    nccl_ctx *x = calloc(1, sizeof *x);
    if (!c || !x) { free(c); free(x); return NEMB_E_MEMORY; }
    nccl_uid uid;
    memcpy(&uid, id, 128);
    int rc = g_nccl.init_rank(&x->comm, world, uid, rank);   /* binds the CURRENT cuda device */
    if (rc != 0) {
        fprintf(stderr, "nem_b200: ncclCommInitRank(rank %d of %d): %s\n", rank, world,
                g_nccl.errstr ? g_nccl.errstr(rc) : "error");
        free(c); free(x);
        return NEMB_E_CUDA;
    }
    c->ctx = x; c->rank = rank; c->world = world;
    c->allgather = nccl_allgather_cb;
    c->destroy = nccl_destroy_cb;
    *out = c;
    return NEMB_OK;
}

/* ------------------------------------------------------------------ local (threads, one device) */
typedef struct {
    int world, refs;
    pthread_barrier_t bar;
    pthread_mutex_t mu;
    void *stage;            /* device staging buffer, world * bytes_per_rank */
    size_t stage_cap;
    cudaEvent_t *wrote, *read;   /* per rank */
    int failed;
} local_shared;

typedef struct { local_shared *sh; int rank; } local_ctx;

static int local_allgather_cb(void *ctx, const void *send, void *recv, size_t bytes, void *stream)
{
    local_ctx *c = ctx;
    local_shared *sh = c->sh;
    cudaStream_t st = (cudaStream_t)stream;
    int W = sh->world, r = c->rank;
    /* (1) rank 0 sizes the staging buffer once every rank's previous reads are enqueued+done */
    pthread_barrier_wait(&sh->bar);
    if (r == 0 && sh->stage_cap < bytes * W) {
        for (int q = 0; q < W; q++) cudaEventSynchronize(sh->read[q]);
        if (sh->stage) cudaFree(sh->stage);
        sh->stage = NULL; sh->stage_cap = 0;
        if (cudaMalloc(&sh->stage, bytes * W) == cudaSuccess) sh->stage_cap = bytes * W;
        else sh->failed = 1;
    }
    pthread_barrier_wait(&sh->bar);
    if (sh->failed) return -1;
    /* (2) my block -> stage, after every rank finished reading the previous round */
    for (int q = 0; q < W; q++) cudaStreamWaitEvent(st, sh->read[q], 0);
    cudaMemcpyAsync((char *)sh->stage + (size_t)r * bytes, send, bytes, cudaMemcpyDeviceToDevice, st);
    cudaEventRecord(sh->wrote[r], st);
    pthread_barrier_wait(&sh->bar);
    /* (3) everybody's block -> my recv */
    for (int q = 0; q < W; q++) cudaStreamWaitEvent(st, sh->wrote[q], 0);
    cudaError_t e = cudaMemcpyAsync(recv, sh->stage, bytes * W, cudaMemcpyDeviceToDevice, st);
    cudaEventRecord(sh->read[r], st);
    pthread_barrier_wait(&sh->bar);
    return e == cudaSuccess ? 0 : -1;
}

static void local_destroy_cb(void *ctx)
{
    local_ctx *c = ctx;
    local_shared *sh = c->sh;
    pthread_mutex_lock(&sh->mu);
    int left = --sh->refs;
    pthread_mutex_unlock(&sh->mu);
    if (left == 0) {
        for (int q = 0; q < sh->world; q++) { cudaEventDestroy(sh->wrote[q]); cudaEventDestroy(sh->read[q]); }
        if (sh->stage) cudaFree(sh->stage);
        pthread_barrier_destroy(&sh->bar);
        pthread_mutex_destroy(&sh->mu);
        free(sh->wrote); free(sh->read); free(sh);
    }
    free(c);
}

int nemb_comm_create_local(nemb_comm **out_array, int world)
{
    if (!out_array || world < 1 || world > 64) return NEMB_E_ARG;
    local_shared *sh = calloc(1, sizeof *sh);
    if (!sh) return NEMB_E_MEMORY;
    sh->world = world; sh->refs = world;
    pthread_barrier_init(&sh->bar, NULL, (unsigned)world);
    pthread_mutex_init(&sh->mu, NULL);
    sh->wrote = calloc(world, sizeof(cudaEvent_t));
    sh->read = calloc(world, sizeof(cudaEvent_t));
    for (int q = 0; q < world; q++) {
        if (cudaEventCreateWithFlags(&sh->wrote[q], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&sh->read[q], cudaEventDisableTiming) != cudaSuccess)
            return NEMB_E_CUDA;
    }
    for (int r = 0; r < world; r++) {
        nemb_comm *c = calloc(1, sizeof *c);
        local_ctx *x = calloc(1, sizeof *x);
        x->sh = sh; x->rank = r;
        c->ctx = x; c->rank = r; c->world = world;
        c->allgather = local_allgather_cb;
        c->destroy = local_destroy_cb;
        out_array[r] = c;
    }
    return NEMB_OK;
}

void nemb_comm_destroy(nemb_comm *c)
{
    if (!c) return;
    if (c->destroy) c->destroy(c->ctx);
    free(c);
}
