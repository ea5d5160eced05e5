/*
 * oracle/nem_oracle.h -- TEST INFRASTRUCTURE.  CPU restatement (float64, log-domain) of the
 * reference NEM hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; the product library never does.
 */
#ifndef NEM_ORACLE_H
#define NEM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { NEMO_ALGO_NEM = 0, NEMO_ALGO_NCEM = 1 };
enum { NEMO_UPDATE_SEQ = 0, NEMO_UPDATE_PARA = 1 };
enum { NEMO_CONV_NONE = 0, NEMO_CONV_CLAS = 1, NEMO_CONV_CRIT = 2 };
enum { NEMO_PROP_EQUAL = 0, NEMO_PROP_K = 1 };
enum { NEMO_DISP___ = 0, NEMO_DISP_K_ = 1, NEMO_DISP__D = 2, NEMO_DISP_KD = 3 };
enum { NEMO_OK = 0, NEMO_EMPTYCLASS = 1 };

typedef struct {
    int n, d, k;
    const uint8_t *x;        /* [n*d] 0/1, row = family */
    const int32_t *row_ptr;  /* [n+1] or NULL (non-spatial => beta forced to 0) */
    const int32_t *col;      /* [nnz] */
    const float   *wgt;      /* [nnz] */
    int    algo, update, conv, prop, disp;
    float  beta, conv_thr;
    int    it_max;
    int    param_fixed;      /* .m flag 2: skip the M-step (nem_alg.c:1806) */
    int    dolog;            /* only affects which criterion value `crit` convergence sees */
} nemo_problem;

typedef struct {
    int    status;           /* NEMO_OK / NEMO_EMPTYCLASS */
    int    iters;            /* EM iterations executed */
    int    converged;
    double U, D, L, M, Z, G; /* nem_alg.c:2702-2751 */
    int64_t n_allnul;        /* rows where every class has zero density (reference: uniform 1/K) */
    int64_t n_ties;          /* ncem rows whose final argmax was an exact tie */
} nemo_result;

/* theta = (prop[K], center[K*D], disp[K*D]) float32 in/out; t[N*K] float32 out; label[N] out. */
int nemo_fit(const nemo_problem *pb, float *prop, float *center, float *disp,
             float *t, int32_t *label, nemo_result *res);

/* beta estimation (reachable in the reference through the CLI only) */
typedef struct { int n_iter; float conv_thr, step; } nemo_psgrad;     /* BtaPsGradT, nem_typ.h:300-308 */
typedef struct { float step, max, ddrop, dloss, lloss; } nemo_heu;    /* nem_typ.h:319-323 */
float nemo_estim_beta(const nemo_problem *pb, const nemo_psgrad *g, const float *t, float beta,
                      double *out3 /*crit grad dsec, nullable*/);
/* from_partition: t holds the starting classification (INIT_FILE); g != NULL: BETA_PSGRAD */
int nemo_fit_ex(const nemo_problem *pb, int from_partition, const nemo_psgrad *g, float *prop,
                float *center, float *disp, float *t, int32_t *label, nemo_result *res,
                float *beta_out);
/* mode 0 = heu_d, 1 = heu_l (ClassifyByNemHeuBeta, nem_alg.c:731-992) */
int nemo_fit_heuristic(const nemo_problem *pb, int mode, const nemo_heu *hp, float *prop,
                       float *center, float *disp, float *t, int32_t *label, nemo_result *res,
                       float *beta_est, int *n_tested, float *beta_trace, float *crit_trace,
                       int cap);

/* Stage functions (same arithmetic as nemo_fit uses), for per-kernel parity tests. */
void nemo_pack(const uint8_t *x, int n, int d, int words_per_row, uint32_t *out);
void nemo_hamming(const nemo_problem *pb, const float *center, const float *disp,
                  int32_t *h /*[n*k]*/);
void nemo_logpf(const nemo_problem *pb, const float *prop, const float *center,
                const float *disp, double *logpf /*[n*k], -inf when density is 0*/);
void nemo_sweep(const nemo_problem *pb, const double *logpf, double beta, float *t, int32_t *label);
int  nemo_mstep(const nemo_problem *pb, const float *t, float *prop, float *center, float *disp,
                double *nk_out /*[k]*/, double *skd_out /*[k*d]*/);
void nemo_criteria(const nemo_problem *pb, const double *logpf, const float *t, double beta,
                   double *crit6 /*U D L M Z G*/);
int  nemo_levels(int n, const int32_t *row_ptr, const int32_t *col, int32_t *level /*[n]*/);

#ifdef __cplusplus
}
#endif
#endif
