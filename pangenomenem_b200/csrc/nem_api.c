/*
 * nem_api.c -- the reference-facing boundary: int nem(...) with the reference's 13 arguments
 * (NEM/nem_exe.h:23-35, NEM/nem_exe.c:239-704), reading <Fname>.str/.dat/.nei/.m and writing
 * <Fname>.uf|.cf, .mf and, with dolog, .log and .stderr exactly where and how the reference
 * does, so that ppanggolin.py:1814-1826 + 1886-1972 drives it unchanged.
 *
 * What differs on purpose (INTEGRATION.md "Behavioural notes"):
 *   - arguments are validated BEFORE any file is read and a bad value returns EXIT_E_ARGS;
 *     the reference overwrites its own error flag (nem_exe.c:371-431 vs 472) and carries on
 *     with enum value -1;
 *   - every error returns an ExitET code; the reference's early returns leak raw StatusET
 *     values (nem_exe.c:301,309,476,520);
 *   - stderr is never closed when dolog=0 (nem_exe.c:274,657 closes the process's stderr);
 *   - MAP ties go to the first class (TIE_FIRST); the reference draws them with a wall-clock
 *     seeded random() (nem_exe.c:353,361,621);
 *   - norm/lapl families, gem, init modes 0/3/4 and image data are off the PPanGGOLiN path
 *      and return EXIT_E_ARGS with a message.
 */
#define _GNU_SOURCE   /* dladdr */
#include "nem_b200.h"
#include "nem_io.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <dlfcn.h>
#include <errno.h>
#include <signal.h>
#include <spawn.h>
#include <sys/types.h>
#include <sys/wait.h>

enum { EXIT_OK_ = 0, EXIT_W_RESULT_ = 1, EXIT_E_ARGS_ = 2, EXIT_E_FILE_ = 3, EXIT_E_MEMORY_ = 4,
       EXIT_E_SYSTEM_ = 5, EXIT_E_BUG_ = 6 };   /* ExitET, NEM/lib_io.h:22-34 */

static const char *kVersion = "1.08-a";          /* NemVersionStrC, nem_exe.c:233 */
static const char *kAlgoDes[] = {"NEM", "NCEM (C-step)", "GEM (Monte-Carlo at E-step)"};
static const char *kPropDes[] = {"P_", "Pk"};
static const char *kDispDes[] = {"S__", "SK_", "S_D", "S_KD"};

static int find_str(const char *s, const char *const *tab, int n)
{
    if (!s) return -1;
    for (int i = 0; i < n; i++)
        if (!strcmp(s, tab[i])) return i;      /* GetEnum, nem_exe.c:1784-1809: exact match */
    return -1;
}

typedef struct {
    FILE *flog, *ferr;
    int k, d, n;
    double mult;
    float beta;
} log_ctx;

/* WriteLogCrit (nem_alg.c:2620-2646) x2 + WriteLogClasses (nem_alg.c:1995-2052) */
static void log_iteration(void *user, int iter, const double cb[6], const double ca[6],
                          const float *prop, const float *center, const float *disp,
                          const float *nk)
{
    log_ctx *L = user;
    if (L->ferr) {
        if (iter > 0) fprintf(L->ferr, "\b\b\b\b\b%4d ", iter);   /* nem_alg.c:1795-1796 */
    }
    if (!L->flog) return;
    FILE *f = L->flog;
    fprintf(f, "%4d ", iter);
    fprintf(f, " %5.0f %5.0f %5.3f", (double)(float)cb[0] * L->mult, (double)(float)cb[3] * L->mult, (double)NAN);
    fprintf(f, " %5.0f %5.0f %5.3f", (double)(float)ca[0] * L->mult, (double)(float)ca[3] * L->mult, (double)NAN);
    fprintf(f, "  %5.3f ", (double)L->beta);
    for (int c = 0; c < L->k; c++) fprintf(f, " %5.3f", (double)prop[c]);
    fprintf(f, " ");
    for (int q = 0; q < L->k * L->d; q++) fprintf(f, " %7.3f", (double)center[q]);
    fprintf(f, " ");
    for (int q = 0; q < L->k * L->d; q++) fprintf(f, " %7.3f", (double)disp[q]);
    fprintf(f, " ");
    for (int c = 0; c < L->k; c++)
        for (int j = 0; j < L->d; j++) fprintf(f, " %7.1f", (double)nk[c]);
    fprintf(f, "\n");
    if (iter == 0) { /* Needinit blank line, then WriteLogHeader (nem_alg.c:1986-1987, 1883-1946) */
        fprintf(f, "\n");
        fprintf(f, "%4s  %5s %5s %5s", "It", "UM", "PM", "Er");
        fprintf(f, " %3s%-2d %3s%-2d %3s%-2d", "UE", 1, "PE", 1, "Er", 1);
        fprintf(f, "  %5s ", "Beta");
        for (int c = 0; c < L->k; c++) fprintf(f, " %3s%02d", "P", c + 1);
        fprintf(f, " ");
        for (int c = 0; c < L->k; c++)
            for (int j = 0; j < L->d; j++) fprintf(f, " %3s%02d_%1d", "M", c + 1, j + 1);
        fprintf(f, " ");
        for (int c = 0; c < L->k; c++)
            for (int j = 0; j < L->d; j++) fprintf(f, " %3s%02d_%1d", "D", c + 1, j + 1);
        fprintf(f, " ");
        for (int c = 0; c < L->k; c++)
            for (int j = 0; j < L->d; j++) fprintf(f, " %3s%02d_%1d", "n", c + 1, j + 1);
        fprintf(f, "\n");
    }
}

static int finish(FILE *ferr, int own_err, int code)
{
    switch (code) {                      /* messages of nem_exe.c:637-703 */
    case EXIT_W_RESULT_: fprintf(ferr, "*** NEM warning status : empty class\n"); break;
    case EXIT_E_ARGS_:   fprintf(ferr, "*** NEM error status : bad arguments\n"); break;
    case EXIT_E_MEMORY_: fprintf(ferr, "*** NEM error status : not enough memory\n"); break;
    case EXIT_E_FILE_:   fprintf(ferr, "*** NEM error status : wrong file format\n"); break;
    case EXIT_E_SYSTEM_: fprintf(ferr, "*** NEM error status : no usable CUDA device (no CPU path)\n"); break;
    case EXIT_E_BUG_:    fprintf(ferr, "*** NEM internal error\n"); break;
    default: break;
    }
    if (own_err) fclose(ferr); else fflush(ferr);
    return code;
}

static int map_status(int nemb)
{
    switch (nemb) {
    case NEMB_OK: return EXIT_OK_;
    case NEMB_W_EMPTYCLASS: return EXIT_W_RESULT_;
    case NEMB_E_ARG: return EXIT_E_ARGS_;
    case NEMB_E_FILE: return EXIT_E_FILE_;
    case NEMB_E_MEMORY: return EXIT_E_MEMORY_;
    case NEMB_E_CUDA: return EXIT_E_SYSTEM_;
    default: return EXIT_E_BUG_;
    }
}

/* ------------------------------------------------------------------ engine handle of nem()
 * PPanGGOLiN calls nem() again and again from one process (the chunk loop ppanggolin.py:1045-1095,
 * the evolution-curve workers command_line.py:262-281).  Creating an engine per call costs more
 * than the fit itself (device buffers, pinned status blocks, streams: ~45 ms of an 81 ms call on
 * 100 000 x 500), so one engine per process and device is kept between calls -- its buffers are
 * grow-only.  A second concurrent call (threads) or a call from a forked child gets a private
 * engine for the call.  NEM_B200_NO_HANDLE_CACHE=1 restores create/destroy per call. */
static pthread_mutex_t g_cache_mu = PTHREAD_MUTEX_INITIALIZER;
static nemb_handle *g_cache_h;
static int g_cache_dev = -2, g_cache_busy;
static pid_t g_cache_pid;

static int acquire_engine(int device, nemb_handle **out, int *cached)
{
    *cached = 0;
    if (device < 0) {                                   /* same rule as nemb_create */
        const char *env = getenv("NEM_B200_DEVICE");
        if (env && *env) device = atoi(env);
    }
    if (!getenv("NEM_B200_NO_HANDLE_CACHE")) {
        pthread_mutex_lock(&g_cache_mu);
        if (g_cache_h && g_cache_pid != getpid()) {     /* forked child: the parent's context is */
            g_cache_h = NULL; g_cache_busy = 0;         /* not ours -- forget it, do not free it */
        }
        if (!g_cache_busy) {
            if (g_cache_h && g_cache_dev != device) { nemb_destroy(g_cache_h); g_cache_h = NULL; }
            int rc = NEMB_OK;
            if (!g_cache_h) {
                rc = nemb_create(&g_cache_h, device);
                g_cache_dev = device; g_cache_pid = getpid();
            }
            if (rc == NEMB_OK) { g_cache_busy = 1; *out = g_cache_h; *cached = 1; }
            pthread_mutex_unlock(&g_cache_mu);
            return rc;
        }
        pthread_mutex_unlock(&g_cache_mu);
    }
    return nemb_create(out, device);
}

static void release_engine(nemb_handle *h, int cached)
{
    if (!h) return;
    if (!cached) { nemb_destroy(h); return; }
    pthread_mutex_lock(&g_cache_mu);
    g_cache_busy = 0;
    pthread_mutex_unlock(&g_cache_mu);
}


/* ------------------------------------------------------------------ forked callers
 * PPanGGOLiN calls nem() in the parent (ppanggolin.py:1125, 1207) and then again from forked pool
 * workers (ppanggolin.py:1039; command_line.py:262-281, 618).  A CUDA context does not survive
 * fork(), and the caller ignores nem()'s return value, so a failing child would silently turn every
 * family into "undefined".  A process that inherited an initialised CUDA state therefore never
 * touches CUDA: it starts ONE helper process (`nem_exe --serve`, next to this library; posix_spawn,
 * i.e. fork + immediate exec) and forwards its nem() calls to it over a pipe pair.  The helper owns
 * its own context and engine cache, so a pool worker pays the context creation once.
 * NEM_B200_FORCE_HELPER=1 routes every call that way (tests). */
extern char **environ;
int nemb_i_cuda_owner_pid(void);
static pthread_mutex_t g_helper_mu = PTHREAD_MUTEX_INITIALIZER;
static pid_t g_helper_pid, g_helper_owner;
static int g_helper_wr = -1, g_helper_rd = -1;

#define NEM_HELPER_MAGIC 0x4e454d42
typedef struct {
    int32_t magic, nk, it_max, dolog, init_mode, has_extra;
    float beta, thr;
    int32_t len[7];          /* Fname algo convergence format family proportion dispersion */
    nem_b200_extra extra;
} helper_req;

static int write_all(int fd, const void *p, size_t n)
{
    const char *c = p;
    while (n) {
        ssize_t w = write(fd, c, n);
        if (w < 0) { if (errno == EINTR) continue; return -1; }
        c += w; n -= (size_t)w;
    }
    return 0;
}
static int read_all(int fd, void *p, size_t n)
{
    char *c = p;
    while (n) {
        ssize_t r = read(fd, c, n);
        if (r < 0) { if (errno == EINTR) continue; return -1; }
        if (r == 0) return -1;
        c += r; n -= (size_t)r;
    }
    return 0;
}

static int must_use_helper(void)
{
    const char *f = getenv("NEM_B200_FORCE_HELPER");
    if (f && *f && getenv("NEM_B200_IS_HELPER") == NULL) return 1;
    int owner = nemb_i_cuda_owner_pid();
    return owner != 0 && owner != (int)getpid();
}

int nem_b200_helper_pid(void) { return g_helper_owner == getpid() ? (int)g_helper_pid : 0; }

static void helper_close(void)
{
    if (g_helper_wr >= 0) close(g_helper_wr);
    if (g_helper_rd >= 0) close(g_helper_rd);
    g_helper_wr = g_helper_rd = -1;
    if (g_helper_pid > 0 && g_helper_owner == getpid()) { int st; waitpid(g_helper_pid, &st, 0); }
    g_helper_pid = 0;
}
static void helper_atexit(void) { if (g_helper_owner == getpid()) helper_close(); }

static int helper_start(void)
{
    if (g_helper_pid > 0 && g_helper_owner == getpid()) return 0;
    if (g_helper_owner != getpid()) {            /* descriptors inherited from the parent's helper */
        if (g_helper_wr >= 0) close(g_helper_wr);
        if (g_helper_rd >= 0) close(g_helper_rd);
        g_helper_wr = g_helper_rd = -1; g_helper_pid = 0;
    }
    char path[4096];
    Dl_info info;
    const char *exe = getenv("NEM_B200_HELPER");
    if (!exe || !*exe) {
        if (!dladdr((void *)&helper_start, &info) || !info.dli_fname) return -1;
        snprintf(path, sizeof path, "%s", info.dli_fname);
        char *slash = strrchr(path, '/');
        if (!slash) return -1;
        snprintf(slash + 1, sizeof path - (size_t)(slash + 1 - path), "nem_exe");
        exe = path;
    }
    int to[2], from[2];
    if (pipe(to) || pipe(from)) return -1;
    posix_spawn_file_actions_t fa;
    posix_spawn_file_actions_init(&fa);
    posix_spawn_file_actions_adddup2(&fa, to[0], 3);
    posix_spawn_file_actions_adddup2(&fa, from[1], 4);
    posix_spawn_file_actions_addclose(&fa, to[1]);
    posix_spawn_file_actions_addclose(&fa, from[0]);
    char *argv[] = {(char *)exe, (char *)"--serve", NULL};
    /* the helper must use CUDA itself whatever this process was told */
    size_t ne = 0;
    while (environ[ne]) ne++;
    char **env = malloc(sizeof(char *) * (ne + 2));
    if (!env) return -1;
    memcpy(env, environ, sizeof(char *) * ne);
    env[ne] = (char *)"NEM_B200_IS_HELPER=1"; env[ne + 1] = NULL;
    pid_t pid = 0;
    int rc = posix_spawn(&pid, exe, &fa, NULL, argv, env);
    free(env);
    posix_spawn_file_actions_destroy(&fa);
    close(to[0]); close(from[1]);
    if (rc != 0) { close(to[1]); close(from[0]); fprintf(stderr, "nem_b200: cannot start the helper %s: %s\n", exe, strerror(rc)); return -1; }
    static int registered;
    if (!registered) { atexit(helper_atexit); registered = 1; }
    signal(SIGPIPE, SIG_IGN);
    g_helper_pid = pid; g_helper_owner = getpid(); g_helper_wr = to[1]; g_helper_rd = from[0];
    return 0;
}

static int helper_call(const char *Fname, int nk, const char *algo, float beta, const char *conv, float thr,
                       const char *format, int it_max, int dolog, const char *family, const char *prop,
                       const char *disp, int init_mode, const nem_b200_extra *extra)
{
    const char *str[7] = {Fname, algo, conv, format, family, prop, disp};
    helper_req rq;
    memset(&rq, 0, sizeof rq);
    rq.magic = NEM_HELPER_MAGIC; rq.nk = nk; rq.it_max = it_max; rq.dolog = dolog; rq.init_mode = init_mode;
    rq.beta = beta; rq.thr = thr; rq.has_extra = extra != NULL;
    if (extra) rq.extra = *extra;
    for (int i = 0; i < 7; i++) rq.len[i] = str[i] ? (int32_t)strlen(str[i]) : -1;
    int32_t rc = EXIT_E_SYSTEM_;
    pthread_mutex_lock(&g_helper_mu);
    for (int attempt = 0; attempt < 2; attempt++) {
        if (helper_start() != 0) break;
        int bad = write_all(g_helper_wr, &rq, sizeof rq);
        for (int i = 0; i < 7 && !bad; i++)
            if (rq.len[i] > 0) bad = write_all(g_helper_wr, str[i], (size_t)rq.len[i]);
        if (!bad) bad = read_all(g_helper_rd, &rc, sizeof rc);
        if (!bad) break;
        rc = EXIT_E_SYSTEM_;
        helper_close();                 /* the helper died: one fresh start */
    }
    pthread_mutex_unlock(&g_helper_mu);
    return rc;
}

/* `nem_exe --serve` (nem_cli.c): requests on descriptor 3, return codes on descriptor 4 */
int nem_b200_serve(int fd_in, int fd_out)
{
    for (;;) {
        helper_req rq;
        if (read_all(fd_in, &rq, sizeof rq) != 0) return 0;        /* the caller is gone */
        if (rq.magic != NEM_HELPER_MAGIC) return 2;
        char *str[7];
        for (int i = 0; i < 7; i++) {
            str[i] = NULL;
            if (rq.len[i] >= 0) {
                if (rq.len[i] > (1 << 20)) return 2;
                str[i] = calloc((size_t)rq.len[i] + 1, 1);
                if (!str[i] || (rq.len[i] > 0 && read_all(fd_in, str[i], (size_t)rq.len[i]) != 0)) return 2;
            }
        }
        int32_t rc = nem_b200_ex(str[0], rq.nk, str[1], rq.beta, str[2], rq.thr, str[3], rq.it_max, rq.dolog,
                                 str[4], str[5], str[6], rq.init_mode, rq.has_extra ? &rq.extra : NULL);
        for (int i = 0; i < 7; i++) free(str[i]);
        if (write_all(fd_out, &rc, sizeof rc) != 0) return 0;
    }
}

int nem_b200_ex(const char *Fname, const int nk, const char *algo, const float beta,
                const char *convergence, const float convergence_th, const char *format,
                const int it_max, const int dolog, const char *model_family,
                const char *proportion, const char *dispersion, const int init_mode,
                const nem_b200_extra *extra)
{
    static const char *algoS[] = {"nem", "ncem", "gem"};
    static const char *convS[] = {"none", "clas", "crit"};
    static const char *fmtS[] = {"hard", "fuzzy"};
    static const char *famS[] = {"norm", "lapl", "bern"};
    static const char *propS[] = {"p_", "pk"};
    static const char *dispS[] = {"s__", "sk_", "s_d", "skd"};
    nem_b200_extra ex;
    memset(&ex, 0, sizeof ex);
    ex.device = -1;
    if (extra) ex = *extra;
    else {
        const char *u = getenv("NEM_B200_UPDATE");           /* "para" | "seq" */
        if (u && !strcmp(u, "para")) ex.update = NEMB_UPDATE_PARA;
        const char *si = getenv("NEM_B200_SWEEP");           /* "level" | "spec" */
        if (si && !strcmp(si, "level")) ex.sweep_impl = NEMB_SWEEP_LEVEL;
        if (si && !strcmp(si, "spec")) ex.sweep_impl = NEMB_SWEEP_SPEC;
        const char *sd = getenv("NEM_B200_SEED");
        if (sd && *sd) ex.seed = atoll(sd);
    }

    if (!Fname) return EXIT_E_ARGS_;
    if (must_use_helper())
        return helper_call(Fname, nk, algo, beta, convergence, convergence_th, format, it_max, dolog,
                           model_family, proportion, dispersion, init_mode, extra);
    FILE *ferr = stderr;
    int own_err = 0;
    char path[4200];
    if (dolog) {                                             /* nem_exe.c:272-282 */
        snprintf(path, sizeof path, "%s.stderr", Fname);
        FILE *f = fopen(path, "w");
        if (f) { ferr = f; own_err = 1; }
    }
    fprintf(ferr, " * * * NEM (spatial data clustering) v%s * * *\n", kVersion);

    /* ---- arguments (nem_exe.c:296-435) */
    int bad = 0;
    if (nk <= 0) { fprintf(ferr, "Nb of classes must be > 0 (here %d)\n", nk); bad = 1; }
    else if (nk > 16) { fprintf(ferr, "Nb of classes must be <= 16 in this engine (here %d)\n", nk); bad = 1; }
    int a = find_str(algo, algoS, 3), cv = find_str(convergence, convS, 3);
    int fm = find_str(format, fmtS, 2), fa = find_str(model_family, famS, 3);
    int pr = find_str(proportion, propS, 2), di = find_str(dispersion, dispS, 4);
    if (a < 0) { fprintf(ferr, " Unknown type of algorithm %s\n", algo ? algo : "(null)"); bad = 1; }
    if (cv < 0) { fprintf(ferr, " Unknown convergence test %s\n", convergence ? convergence : "(null)"); bad = 1; }
    else if (cv != 0 && !(convergence_th > 0)) { fprintf(ferr, " Conv threshold must be > 0 (here %f)\n", convergence_th); bad = 1; }
    if (fm < 0) { fprintf(ferr, " Unknown format %s\n", format ? format : "(null)"); bad = 1; }
    if (it_max < 0) { fprintf(ferr, "Nb iterations must be >= 0 (here %d)\n", it_max); bad = 1; }
    if (fa < 0) { fprintf(ferr, " Unknown family %s\n", model_family ? model_family : "(null)"); bad = 1; }
    if (pr < 0) { fprintf(ferr, " Unknown proportion %s\n", proportion ? proportion : "(null)"); bad = 1; }
    if (di < 0) { fprintf(ferr, " Unknown dispersion %s\n", dispersion ? dispersion : "(null)"); bad = 1; }
    if (a == 2) { fprintf(ferr, " Algorithm gem is not available in the B200 engine\n"); bad = 1; }
    if (fa == 0 || fa == 1) { fprintf(ferr, " Family %s is not available in the B200 engine (bern only)\n", model_family); bad = 1; }
    if (init_mode != 1 && init_mode != 2) {
        fprintf(ferr, "Initialization mode %d is not available in the B200 engine "
                      "(1 = random starts, 2 = parameter file)\n", init_mode);
        bad = 1;
    }
    if (ex.beta_mode < NEMB_BETA_FIX || ex.beta_mode > NEMB_BETA_HEUL) {
        fprintf(ferr, " Unknown beta estimation mode %d\n", ex.beta_mode); bad = 1;
    } else if (ex.beta_mode >= NEMB_BETA_HEUD && init_mode != 2) {
        fprintf(ferr, " The beta heuristics need the parameter file initialisation (init_mode 2) in the B200 engine\n");
        bad = 1;
    } else if (ex.beta_mode != NEMB_BETA_FIX && init_mode != 2) {
        fprintf(ferr, " Beta estimation needs the parameter file initialisation (init_mode 2) in the B200 engine\n");
        bad = 1;
    }
    if (bad) return finish(ferr, own_err, EXIT_E_ARGS_);

    /* ---- files */
    char type = 0, datadesc[2048] = "", neidesc[2048] = "";
    int n = 0, d = 0, rc;
    if ((rc = nemio_read_str(Fname, ferr, &type, &n, &d, datadesc, sizeof datadesc)) != NEMB_OK)
        return finish(ferr, own_err, map_status(rc));
    if (type == 'I') {
        fprintf(ferr, "Image data (type I) is not available in the B200 engine\n");
        return finish(ferr, own_err, EXIT_E_ARGS_);
    }
    int wpr = ((d + 31) / 32 + 3) / 4 * 4;
    uint32_t *xp = NULL;
    int32_t *row_ptr = NULL, *col = NULL;
    float *wgt = NULL;
    size_t kd = (size_t)nk * d;
    float *prop = calloc(nk, sizeof(float)), *center = calloc(kd, sizeof(float)),
          *disp = calloc(kd, sizeof(float));
    float *t = NULL;
    int32_t *label = NULL;
    nemb_handle *h = NULL;
    int h_cached = 0;
    int code = EXIT_OK_, flag = 1, max_neigh = 0;
    log_ctx L;
    memset(&L, 0, sizeof L);
    if (!prop || !center || !disp) { code = EXIT_E_MEMORY_; goto done; }

    fprintf(ferr, "Reading points ...\n");
    long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    snprintf(path, sizeof path, "%s.dat", Fname);
    if ((rc = nemio_read_dat(path, ferr, n, d, wpr, &xp, ncpu > 16 ? 16 : (int)ncpu)) != NEMB_OK) {
        code = map_status(rc); goto done;
    }
    if (init_mode == 2) {
        fprintf(ferr, "Reading parameter file ...\n");
        snprintf(path, sizeof path, "%s.m", Fname);
        if ((rc = nemio_read_m(path, ferr, nk, d, &flag, prop, center, disp)) != NEMB_OK) {
            code = map_status(rc); goto done;
        }
    }
    if (type == 'S') {
        fprintf(ferr, "Reading neighborhood information ...\n");
        if ((rc = nemio_read_nei(Fname, ferr, n, &row_ptr, &col, &wgt, &max_neigh, neidesc,
                                 sizeof neidesc)) != NEMB_OK) {
            code = map_status(rc); goto done;
        }
    }
    float beta_eff = type == 'S' ? beta : 0.0f;                 /* nem_exe.c:570-574 */

    /* ---- banner (nem_exe.c:576-614) */
    fprintf(ferr, "\nData : ");
    if (datadesc[0]) fprintf(ferr, "%s\n", datadesc); else fprintf(ferr, "\n");
    fprintf(ferr, "  file names =  %10s   |   nb points   = %10d\n", Fname, n);
    fprintf(ferr, "  type       =  %10s   |   dim         = %10d\n", type == 'S' ? "Spatial" : "NoSpatial", d);
    if (type == 'S') {
        fprintf(ferr, "Neighborhood system :\n  max neighb =  %10d\n", max_neigh);
        fprintf(ferr, "%s\n", neidesc);
    }
    fprintf(ferr, "\nNEM parameters :\nType of algorithm : '%s'\n", kAlgoDes[a]);
    fprintf(ferr, "  beta       =  %10.2f   |   nk                    = %3d\n", (double)beta_eff, nk);
    fprintf(ferr, "                %10s   |   model                 = %s, %s %s\n", " ", "Bernoulli",
            kPropDes[pr], kDispDes[di]);
    fprintf(ferr, "\n");

    /* ---- engine */
    if ((rc = acquire_engine(ex.device, &h, &h_cached)) != NEMB_OK) { code = map_status(rc); goto done; }
    if ((rc = nemb_load_packed(h, n, d, wpr, xp, row_ptr, col, wgt)) != NEMB_OK) {
        fprintf(ferr, "%s\n", nemb_last_error(h));
        code = map_status(rc); goto done;
    }
    free(xp); xp = NULL;
    nemb_options o;
    memset(&o, 0, sizeof o);
    o.k = nk; o.algo = a; o.update = ex.update; o.conv = cv; o.prop = pr; o.disp = di;
    o.it_max = it_max; o.param_fixed = (init_mode == 2 && flag == 2); o.dolog = dolog != 0;
    o.sweep_impl = ex.sweep_impl; o.beta = beta; o.conv_thr = convergence_th;
    const int beta_mode = type == 'S' ? ex.beta_mode : NEMB_BETA_FIX;
    if (beta_mode == NEMB_BETA_PSGRAD) {
        o.beta_mode = NEMB_BETA_PSGRAD; o.grad_n_iter = ex.grad_n_iter;
        o.grad_conv = ex.grad_conv; o.grad_step = ex.grad_step;
    }
    nemb_result res;
    L.ferr = ferr; L.k = nk; L.d = d; L.n = n; L.beta = beta_eff;
    if (dolog) {                                                /* StartLogFile, nem_alg.c:1478-1498 */
        snprintf(path, sizeof path, "%s.log", Fname);
        L.flog = fopen(path, "w");
        if (!L.flog) fprintf(ferr, "Could not open file '%s' in write mode\n", path);
        else {
            time_t timer = time(NULL);
            fprintf(L.flog, "NEM log file  -  %s\n", asctime(localtime(&timer)));
            L.mult = exp(-((int)(log(n / 1000.) / log(10))) * log(10));
            fprintf(L.flog, "  Criteria are multiplied by %f\n\n", L.mult);
        }
    }
    if (init_mode == 2) {
        fprintf(ferr, "Initializing parameters from given value\n");
        if (L.flog) fprintf(L.flog, "Initializing parameters from given value :\n");
        fprintf(ferr, "  Iterations : %4d ", 0);
        if (beta_mode >= NEMB_BETA_HEUD) {
            /* ClassifyByNemHeuBeta (nem_alg.c:731-992); the .log keeps the header only (the
             * reference rewrites it for every tested beta, 1014-1023) */
            nemb_beta_heuristic hp = {ex.heu_step, ex.heu_max, ex.heu_ddrop, ex.heu_dloss, ex.heu_lloss};
            float bt[64], ct[64];
            fprintf(ferr, "\n* * Starting heuristic * *\n");
            rc = nemb_fit_beta_heuristic(h, &o, beta_mode, &hp, prop, center, disp, &res, bt, ct, 64);
            if (rc == NEMB_OK || rc == NEMB_W_EMPTYCLASS) {
                for (int i = 0; i < res.n_beta_tested && i < 64; i++)
                    fprintf(ferr, " * * Tested beta = %5.2f : %s = %10.1f * *\n", (double)bt[i],
                            beta_mode == NEMB_BETA_HEUD ? "D" : "L", (double)ct[i]);
                fprintf(ferr, "\n * * *  Estimated beta : %3.2f * * *\n", (double)res.beta);
            }
        } else
            rc = nemb_fit_logged(h, &o, prop, center, disp, &res, dolog ? log_iteration : NULL, &L);
        if (rc == NEMB_OK || rc == NEMB_W_EMPTYCLASS) beta_eff = res.beta;
    } else {
        fprintf(ferr, "Random initial partitions (%d starts)\n", ex.n_random_inits ? ex.n_random_inits : 50);
        rc = nemb_fit_random(h, &o, ex.n_random_inits, ex.seed, prop, center, disp, &res);
        if (rc == NEMB_OK)
            fprintf(ferr, "Best start was %d (%s = %g)\n", res.best_start, "M", res.M);
    }
    if (L.flog) { fclose(L.flog); L.flog = NULL; }
    if (rc != NEMB_OK && rc != NEMB_W_EMPTYCLASS) {
        fprintf(ferr, "\n%s\n", nemb_last_error(h));
        code = map_status(rc); goto done;
    }
    fprintf(ferr, "\n  criterion NEM = %6.3f / Ps-Like = %6.3f / Lmix = %6.3f\n", res.U, res.M, res.L);
    if (rc == NEMB_W_EMPTYCLASS) {
        fprintf(ferr, "Class %d empty at iteration %d\n", res.empty_class, res.iters);
        code = EXIT_W_RESULT_; goto done;                       /* no result files, nem_exe.c:624-631 */
    }
    if (cv != 0 && init_mode == 2)
        fprintf(ferr, res.converged ? "  NEM converged after %d iterations\n"
                                    : "  NEM did not converge after %d iterations\n", res.iters);

    /* ---- SaveResults (nem_exe.c:1596-1781) */
    fprintf(ferr, "Saving results ...\n");
    char outname[4200];
    snprintf(outname, sizeof outname, "%s%s", Fname, fm == 0 ? ".cf" : ".uf");
    if (fm == 0) {
        label = malloc(sizeof(int32_t) * (size_t)n);
        if (!label) { code = EXIT_E_MEMORY_; goto done; }
        if ((rc = nemb_get_labels(h, label)) != NEMB_OK) { code = map_status(rc); goto done; }
        rc = nemio_write_cf(outname, ferr, n, label);
    } else {
        t = malloc(sizeof(float) * (size_t)n * nk);
        if (!t) { code = EXIT_E_MEMORY_; goto done; }
        if ((rc = nemb_get_posteriors(h, t)) != NEMB_OK) { code = map_status(rc); goto done; }
        rc = nemio_write_uf(outname, ferr, n, nk, t);
    }
    if (rc != NEMB_OK) { code = map_status(rc); goto done; }
    snprintf(path, sizeof path, "%s.mf", Fname);
    double crit4[4] = {res.U, res.D, res.L, res.M};
    if ((rc = nemio_write_mf_mode(path, ferr, nk, d, crit4, beta_eff, beta_mode, prop, center, disp)) != NEMB_OK) {
        code = map_status(rc); goto done;
    }
    fprintf(ferr, "NEM completed, classification in %s\n", outname);
    fprintf(ferr, " criteria and parameters in %s%s\n", Fname, ".mf");
    if (dolog) fprintf(ferr, "Log of detailed running in %s.log\n", Fname);

done:
    if (L.flog) fclose(L.flog);
    release_engine(h, h_cached);
    free(xp); free(row_ptr); free(col); free(wgt); free(prop); free(center); free(disp);
    free(t); free(label);
    return finish(ferr, own_err, code);
}

int nem(const char *Fname, const int nk, const char *algo, const float beta,
        const char *convergence, const float convergence_th, const char *format,
        const int it_max, const int dolog, const char *model_family, const char *proportion,
        const char *dispersion, const int init_mode)
{
    return nem_b200_ex(Fname, nk, algo, beta, convergence, convergence_th, format, it_max, dolog,
                       model_family, proportion, dispersion, init_mode, NULL);
}

/* ------------------------------------------------------------------ host loader / writers */
int nemb_read_files(const char *base, int k, nemb_host_problem *out)
{
    if (!base || !out || k < 0) return NEMB_E_ARG;
    memset(out, 0, sizeof *out);
    char type = 0, path[4200];
    int n = 0, d = 0, rc;
    if ((rc = nemio_read_str(base, stderr, &type, &n, &d, NULL, 0)) != NEMB_OK) return rc;
    if (type == 'I') { fprintf(stderr, "Image data (type I) is not available in the B200 engine\n"); return NEMB_E_ARG; }
    out->n = n; out->d = d; out->spatial = type == 'S';
    out->words_per_row = ((d + 31) / 32 + 3) / 4 * 4;
    long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    snprintf(path, sizeof path, "%s.dat", base);
    if ((rc = nemio_read_dat(path, stderr, n, d, out->words_per_row, &out->x_packed,
                             ncpu > 16 ? 16 : (int)ncpu)) != NEMB_OK) goto fail;
    if (k > 0) {
        size_t kd = (size_t)k * d;
        out->prop = calloc(k, sizeof(float)); out->center = calloc(kd, sizeof(float));
        out->disp = calloc(kd, sizeof(float));
        if (!out->prop || !out->center || !out->disp) { rc = NEMB_E_MEMORY; goto fail; }
        snprintf(path, sizeof path, "%s.m", base);
        int flag = 0;
        if ((rc = nemio_read_m(path, stderr, k, d, &flag, out->prop, out->center, out->disp)) != NEMB_OK) goto fail;
        out->m_flag = flag;
    }
    if (out->spatial) {
        int mx = 0;
        if ((rc = nemio_read_nei(base, stderr, n, &out->row_ptr, &out->col, &out->wgt, &mx, NULL, 0)) != NEMB_OK) goto fail;
        out->max_neigh = mx; out->nnz = out->row_ptr[n];
    }
    return NEMB_OK;
fail:
    nemb_free_host_problem(out);
    return rc;
}

void nemb_free_host_problem(nemb_host_problem *p)
{
    if (!p) return;
    free(p->x_packed); free(p->row_ptr); free(p->col); free(p->wgt);
    free(p->prop); free(p->center); free(p->disp);
    memset(p, 0, sizeof *p);
}

int nemb_write_uf(const char *path, int n, int k, const float *t) { return nemio_write_uf(path, stderr, n, k, t); }
int nemb_write_cf(const char *path, int n, const int32_t *label) { return nemio_write_cf(path, stderr, n, label); }
int nemb_write_mf(const char *path, int k, int d, double U, double D, double L, double M, float beta,
                  const float *prop, const float *center, const float *disp)
{
    double c4[4] = {U, D, L, M};
    return nemio_write_mf(path, stderr, k, d, c4, beta, prop, center, disp);
}
