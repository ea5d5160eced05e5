/*
 * oracle/nem_oracle.c -- TEST INFRASTRUCTURE (oracle #2).  NOT product code: the shipped
 * library (pangenomenem_b200/csrc) never links, loads or calls this file.
 *
 * A plain-C, float64, log-domain restatement of the NEM partitioning hot path of
 * labgem/pangenomeNEM (reference root /root/reference/ppanggolin/NEM).  Every function cites
 * the reference lines it follows.  Parity status: the reference ships NO tests, fixtures or
 * golden vectors, so this oracle is pinned against OUTPUTS OF THE
 * REFERENCE ITSELF: tests/test_oracle_vs_reference.py runs oracle/_ref (the unmodified
 * reference compiled in place) when present, and tests/golden/ holds vectors generated from
 * it by tests/golden/make_golden.py.
 *
 * Numerical contract (shared with the CUDA engine; DESIGN.md "Numerical contract"):
 *   - theta = (p_k, mu_kd, eps_kd) is float32, exactly what the reference stores
 *     (nem_typ.h:434-445) and every expression on theta alone repeats the reference's float
 *     expression bit for bit: (1-eps)/eps and 1-eps in float, log() in double
 *     (nem_mod.c:656-661).
 *   - sums over genomes, families and neighbours are float64 (the reference uses float32
 *     running sums: nem_mod.c:631,661,1298-1312,1674-1683; nem_alg.c:2860-2874,2734-2745),
 *   - posteriors are formed in the log domain (the reference multiplies p_k*f_k in linear
 *     double and underflows for D >~ 500: nem_alg.c:2282,2581-2613) and stored as float32
 *     like ClassifM,
 *   - MAP ties go to the first maximum (TIE_FIRST, nem_alg.c:641; the reference's default
 *     TIE_RANDOM is seeded by the wall clock, nem_exe.c:353,621, and cannot be reproduced).
 *
 * OpenMP: only loops whose iterations are independent (one family per iteration, every sum
 * inside it in the same order as the sequential loop) or whose sums are exact integers (the
 * ncem statistics: t in {0,1}) are threaded, so the results do not depend on the thread count;
 * it lets the whole-fit parity tests run at the BASELINE shapes (250 000 x 1000, D = 5000).
 */
#include "nem_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NEMO_EPSILON 1e-20 /* nem_typ.h:65 */

/* ------------------------------------------------------------------ loader: bit packing */
/* genome d -> bit d%32 (LSB first) of word d/32, zero padded. */
void nemo_pack(const uint8_t *x, int n, int d, int wpr, uint32_t *out)
{
    memset(out, 0, (size_t)n * wpr * sizeof(uint32_t));
    for (int i = 0; i < n; i++)
        for (int j = 0; j < d; j++)
            if (x[(size_t)i * d + j]) out[(size_t)i * wpr + (j >> 5)] |= 1u << (j & 31);
}

/* ------------------------------------------------------------------ Bernoulli density */
/* Per (k,d) tables of DensBernoulli (nem_mod.c:649-674):
 *   absdif = abs((int)(x - mu))                       -> m0 (x=0), m1 (x=1)
 *   disp > EPSILON : term = absdif*log((1-disp)/disp) - log(1-disp)
 *   else           : absdif != 0 -> zero density ; absdif == 0 -> no term              */
typedef struct {
    double *cost0, *cost1; /* [d] */
    uint8_t *forb0, *forb1;
    int *m0, *m1;
} class_tab;

static void tab_alloc(class_tab *t, int d)
{
    t->cost0 = malloc(sizeof(double) * d); t->cost1 = malloc(sizeof(double) * d);
    t->forb0 = malloc(d); t->forb1 = malloc(d);
    t->m0 = malloc(sizeof(int) * d); t->m1 = malloc(sizeof(int) * d);
}
static void tab_free(class_tab *t)
{
    free(t->cost0); free(t->cost1); free(t->forb0); free(t->forb1); free(t->m0); free(t->m1);
}
static void tab_fill(class_tab *t, int d, const float *mu, const float *eps)
{
    for (int j = 0; j < d; j++) {
        float e = eps[j];
        int m0 = abs((int)(0.0f - mu[j]));
        int m1 = abs((int)(1.0f - mu[j]));
        t->m0[j] = m0; t->m1[j] = m1;
        if ((double)e > NEMO_EPSILON) {
            float ratio = (1.0f - e) / e;       /* float, as `( 1 - disp ) / disp` */
            float om = 1.0f - e;
            double a = log((double)ratio), c = -log((double)om);
            t->cost0[j] = m0 * a + c; t->cost1[j] = m1 * a + c;
            t->forb0[j] = t->forb1[j] = 0;
        } else {
            t->cost0[j] = t->cost1[j] = 0.0;
            t->forb0[j] = (m0 != 0); t->forb1[j] = (m1 != 0);
        }
    }
}

/* integer part of the density: H_ik = sum_d absdif (nem_mod.c:657-658) */
void nemo_hamming(const nemo_problem *pb, const float *center, const float *disp, int32_t *h)
{
    int n = pb->n, d = pb->d, K = pb->k;
    class_tab tb; tab_alloc(&tb, d);
    for (int k = 0; k < K; k++) {
        tab_fill(&tb, d, center + (size_t)k * d, disp + (size_t)k * d);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) {
            const uint8_t *xi = pb->x + (size_t)i * d;
            int32_t s = 0;
            for (int j = 0; j < d; j++) s += xi[j] ? tb.m1[j] : tb.m0[j];
            h[(size_t)i * K + k] = s;
        }
    }
    tab_free(&tb);
}

/* ComputePkFkiM (nem_alg.c:2260-2285) + DensBernoulli (nem_mod.c:619-690), log domain:
 * logpf_ik = log p_k + log f_k(x_i); -inf where the reference has density 0. */
void nemo_logpf(const nemo_problem *pb, const float *prop, const float *center,
                const float *disp, double *logpf)
{
    int n = pb->n, d = pb->d, K = pb->k;
    class_tab tb; tab_alloc(&tb, d);
    for (int k = 0; k < K; k++) {
        double pk = prop[k];
        double lp = (pk > NEMO_EPSILON) ? log(pk) : -INFINITY; /* nem_alg.c:2265-2271 */
        tab_fill(&tb, d, center + (size_t)k * d, disp + (size_t)k * d);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) {
            const uint8_t *xi = pb->x + (size_t)i * d;
            double dk = 0.0; int nul = 0;
            for (int j = 0; j < d; j++) {
                if (xi[j]) { dk += tb.cost1[j]; nul |= tb.forb1[j]; }
                else       { dk += tb.cost0[j]; nul |= tb.forb0[j]; }
            }
            logpf[(size_t)i * K + k] = nul ? -INFINITY : lp - dk;
        }
    }
    tab_free(&tb);
}

/* ------------------------------------------------------------------ E-step sweep */
/* SumNeighsOfClass (nem_alg.c:2850-2884): ctx_k = sum_j w_ij * t_jk, file order, un-normalised */
static void context(const nemo_problem *pb, int i, const float *t, double *ctx)
{
    int K = pb->k;
    for (int k = 0; k < K; k++) ctx[k] = 0.0;
    if (!pb->row_ptr) return;
    for (int e = pb->row_ptr[i]; e < pb->row_ptr[i + 1]; e++) {
        const float *tj = t + (size_t)pb->col[e] * K;
        double w = pb->wgt[e];
        for (int k = 0; k < K; k++) ctx[k] += w * (double)tj[k];
    }
}

/* One E-step sweep: ComputePartitionNEM (nem_alg.c:2364-2395) with ComputeLocalProba
 * (nem_alg.c:2576-2613) in the log domain and, for ncem, ComputeMAP + LabelToClassVector
 * (nem_alg.c:2386-2391, 603-615, 659-663).  seq = in place in index order (UPDATE_SEQ,
 * ORDER_DIRECT: nem_exe.c:359-360, 726-727); para reads the pre-sweep copy. */
static void sweep_impl(const nemo_problem *pb, const double *logpf, double beta, float *t,
                       int32_t *label, int64_t *n_allnul, int64_t *n_ties)
{
    int n = pb->n, K = pb->k;
    float *src = t, *copy = NULL;
    double *ctx = malloc(sizeof(double) * K), *sc = malloc(sizeof(double) * K);
    if (pb->update == NEMO_UPDATE_PARA) {
        copy = malloc(sizeof(float) * (size_t)n * K);
        memcpy(copy, t, sizeof(float) * (size_t)n * K);
        src = copy;
    }
    int64_t alln = 0, ties = 0;
    for (int i = 0; i < n; i++) {
        context(pb, i, src, ctx);
        const double *lp = logpf + (size_t)i * K;
        float *ti = t + (size_t)i * K;
        double mx = -INFINITY; int kmax = 0;
        for (int k = 0; k < K; k++) {
            sc[k] = lp[k] + beta * ctx[k];
            if (sc[k] > mx) { mx = sc[k]; kmax = k; }
        }
        if (mx == -INFINITY) { /* cumnum == 0 branch (nem_alg.c:2603-2612): uniform */
            alln++;
            if (pb->algo == NEMO_ALGO_NCEM) {
                for (int k = 0; k < K; k++) ti[k] = 0.f;
                ti[0] = 1.f; if (label) label[i] = 0;
            } else {
                for (int k = 0; k < K; k++) ti[k] = (float)(1.0 / K);
                if (label) label[i] = 0;
            }
            continue;
        }
        if (pb->algo == NEMO_ALGO_NCEM) {
            for (int k = kmax + 1; k < K; k++) if (sc[k] == mx) { ties++; break; }
            for (int k = 0; k < K; k++) ti[k] = 0.f;
            ti[kmax] = 1.f;
        } else {
            double z = 0.0;
            for (int k = 0; k < K; k++) { sc[k] = exp(sc[k] - mx); z += sc[k]; }
            for (int k = 0; k < K; k++) ti[k] = (float)(sc[k] / z);
        }
        if (label) label[i] = kmax;
    }
    if (n_allnul) *n_allnul = alln;
    if (n_ties) *n_ties = ties;
    free(ctx); free(sc); free(copy);
}

void nemo_sweep(const nemo_problem *pb, const double *logpf, double beta, float *t, int32_t *label)
{
    sweep_impl(pb, logpf, beta, t, label, NULL, NULL);
}

/* ------------------------------------------------------------------ M-step */
/* EstimPara for the Bernoulli family (nem_mod.c:446-465) = Laplace estimator on 0/1 data:
 *   EstimSizes        (nem_mod.c:1275-1317)  n_k = sum_i t_ik            (n_kd = n_k, no NaN)
 *   ComputeMedian     (nem_mod.c:1422-1479)  weighted median of a 0/1 column:
 *                                            mu = 1 if S_kd > n_k/2, 0 if <, 1/2 if equal
 *   EstimLaplaceIner  (nem_mod.c:1646-1704)  iner_kd = sum_i t_ik |x_id - mu_kd|
 *   InerToDisp*       (nem_mod.c:922-1174)   eps from iner, four dispersion models
 *   proportions       (nem_mod.c:455-465)
 * Returns NEMO_EMPTYCLASS when some n_k <= EPSILON (nem_mod.c:1363,1404-1409); like the
 * reference the centre of an empty class is kept. */
int nemo_mstep(const nemo_problem *pb, const float *t, float *prop, float *center, float *disp,
               double *nk_out, double *skd_out)
{
    int n = pb->n, d = pb->d, K = pb->k, status = NEMO_OK;
    double *nk = calloc(K, sizeof(double)), *s = calloc((size_t)K * d, sizeof(double));
    int hard = 1;   /* every t_ik is 0 or 1: n_k and S_kd are integer counts, exact in any order */
    for (size_t q = 0; q < (size_t)n * K && hard; q++) hard = t[q] == 0.0f || t[q] == 1.0f;
    if (hard && (double)n < 9.0e15) {
#pragma omp parallel
        {
            double *nk_t = calloc(K, sizeof(double)), *s_t = calloc((size_t)K * d, sizeof(double));
#pragma omp for schedule(static) nowait
            for (int i = 0; i < n; i++) {
                const uint8_t *xi = pb->x + (size_t)i * d;
                const float *ti = t + (size_t)i * K;
                for (int k = 0; k < K; k++) {
                    if (ti[k] == 0.0f) continue;
                    nk_t[k] += 1.0;
                    double *sk = s_t + (size_t)k * d;
                    for (int j = 0; j < d; j++) sk[j] += (double)xi[j];
                }
            }
#pragma omp critical
            {
                for (int k = 0; k < K; k++) nk[k] += nk_t[k];
                for (size_t q = 0; q < (size_t)K * d; q++) s[q] += s_t[q];
            }
            free(nk_t); free(s_t);
        }
    } else
    for (int i = 0; i < n; i++) {
        const uint8_t *xi = pb->x + (size_t)i * d;
        const float *ti = t + (size_t)i * K;
        for (int k = 0; k < K; k++) {
            double tik = ti[k];
            if (tik == 0.0) continue;
            nk[k] += tik;
            double *sk = s + (size_t)k * d;
            for (int j = 0; j < d; j++) if (xi[j]) sk[j] += tik;
        }
    }
    float *nkf = malloc(sizeof(float) * K), *iner = malloc(sizeof(float) * (size_t)K * d);
    for (int k = 0; k < K; k++) {
        nkf[k] = (float)nk[k];
        for (int j = 0; j < d; j++) {
            double skd = s[(size_t)k * d + j], half = 0.5 * nk[k], in;
            if ((double)nkf[k] > NEMO_EPSILON) {
                float mu = skd > half ? 1.0f : (skd < half ? 0.0f : 0.5f);
                center[(size_t)k * d + j] = mu;
            } else {
                status = NEMO_EMPTYCLASS;   /* centre kept (nem_mod.c:1405) */
            }
            float mu = center[(size_t)k * d + j];
            /* sum_i t_ik |x_id - mu| for x in {0,1}: S*|1-mu| + (n-S)*|0-mu| */
            in = skd * fabs(1.0 - (double)mu) + (nk[k] - skd) * fabs((double)mu);
            iner[(size_t)k * d + j] = (float)in;
        }
    }
    switch (pb->disp) {
    case NEMO_DISP_KD: /* nem_mod.c:1152-1170, MISSING_IGNORE branch */
        for (int k = 0; k < K; k++)
            if ((double)nkf[k] > NEMO_EPSILON)
                for (int j = 0; j < d; j++)
                    disp[(size_t)k * d + j] = iner[(size_t)k * d + j] / nkf[k];
        break;
    case NEMO_DISP_K_: /* nem_mod.c:1043-1073 */
        for (int k = 0; k < K; k++)
            if (nkf[k] > 0) {
                double si = 0.0, sn = 0.0;
                for (int j = 0; j < d; j++) { si += iner[(size_t)k * d + j]; sn += nkf[k]; }
                float dk = (float)si / (float)sn;
                for (int j = 0; j < d; j++) disp[(size_t)k * d + j] = dk;
            }
        break;
    case NEMO_DISP__D: /* nem_mod.c:1104-1126 */
        for (int j = 0; j < d; j++) {
            float si = 0.f, sn = 0.f;
            for (int k = 0; k < K; k++) { sn += nkf[k]; si += iner[(size_t)k * d + j]; }
            float dd = si / sn;
            for (int k = 0; k < K; k++) disp[(size_t)k * d + j] = dd;
        }
        break;
    default: { /* NEMO_DISP___ : nem_mod.c:988-1015 */
        double si = 0.0, sn = 0.0;
        for (int k = 0; k < K; k++)
            if (nkf[k] > 0)
                for (int j = 0; j < d; j++) { si += iner[(size_t)k * d + j]; sn += nkf[k]; }
        float v = (float)si / (float)sn;
        for (int i = 0; i < K * d; i++) disp[i] = v;
    } }
    for (int k = 0; k < K; k++) /* nem_mod.c:456-465 */
        prop[k] = pb->prop == NEMO_PROP_K ? nkf[k] / (float)n : (float)(1.0 / K);
    if (nk_out) memcpy(nk_out, nk, sizeof(double) * K);
    if (skd_out) memcpy(skd_out, s, sizeof(double) * (size_t)K * d);
    free(nk); free(s); free(nkf); free(iner);
    return status;
}

/* ------------------------------------------------------------------ criteria */
/* ComputeCrit (nem_alg.c:2678-2757) with float64 sums and log-domain L and Z. */
void nemo_criteria(const nemo_problem *pb, const double *logpf, const float *t, double beta,
                   double *crit6)
{
    int n = pb->n, K = pb->k;
    double D = 0, G = 0, L = 0, Z = 0;
    double *ctx = malloc(sizeof(double) * K);
    for (int i = 0; i < n; i++) {
        context(pb, i, t, ctx);
        double lmx = -INFINITY, zmx = -INFINITY;
        for (int k = 0; k < K; k++) {
            double l = logpf[(size_t)i * K + k];
            if (l > lmx) lmx = l;
            if (beta * ctx[k] > zmx) zmx = beta * ctx[k];
        }
        double fs = 0, zs = 0;
        for (int k = 0; k < K; k++) {
            float cik = t[(size_t)i * K + k];
            double l = logpf[(size_t)i * K + k];
            if (cik > FLT_MIN) { /* MINFLOAT, nem_alg.c:2727 */
                double lc = (l == -INFINITY) ? -(double)FLT_MAX : l; /* nem_mod.c:685 */
                D += (double)cik * (lc - log((double)cik));
                G += (double)cik * ctx[k];
            }
            if (lmx > -INFINITY) fs += exp(l - lmx);
            zs += exp(beta * ctx[k] - zmx);
        }
        L += (lmx > -INFINITY) ? lmx + log(fs) : -INFINITY;
        Z -= zmx + log(zs);
    }
    crit6[0] = D + 0.5 * beta * G; /* U */
    crit6[1] = D; crit6[2] = L;
    crit6[3] = D + beta * G + Z;   /* M */
    crit6[4] = Z; crit6[5] = G;
    free(ctx);
}

/* ------------------------------------------------------------------ sweep DAG levels */
/* Level schedule equivalent to the in-place index-order sweep (SURVEY.md hard part #1):
 * i must run after every lower-index site it reads AND after every lower-index site that
 * reads it (anti-dependency for asymmetric files).  Returns the depth. */
int nemo_levels(int n, const int32_t *row_ptr, const int32_t *col, int32_t *level)
{
    int depth = 0;
    int32_t *pend = calloc(n, sizeof(int32_t)); /* max level of lower readers of i */
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
        if (level[i] > depth) depth = level[i];
    }
    free(pend);
    return depth;
}

/* ------------------------------------------------------------------ EM driver */
static float max_abs_diff(const float *a, const float *b, size_t m)
{
    float mx = 0.f; /* HasConverged CVTEST_CLAS, nem_alg.c:2075-2089 */
    for (size_t i = 0; i < m; i++) {
        float df = a[i] - b[i];
        if (df < 0) df = -df;
        if (df > mx) mx = df;
    }
    return mx;
}

/* ------------------------------------------------------------------ beta estimation */
/* EstimBeta, BETA_PSGRAD (nem_alg.c:2120-2230): gradient ascent on the log pseudo-likelihood of
 * the classification t under the Potts prior,
 *     crit = sum_i [ beta sum_k t_ik c_ik - log sum_k exp(beta c_ik) ],  c_ik = sum_j w_ij t_jk,
 *     grad = sum_i [ sum_k t_ik c_ik - E_i(c) ],   dsec = sum_i Var_i(c)   (softmax(beta c_i.) law),
 * step <= 0: beta += grad / max(4 dsec, N/10) (integer N/10, nem_alg.c:2194-2200), else
 * beta += grad * step / N; stop after n_iter iterations or when |grad| < conv_thr * N; clamp to
 * [-5, 5] (MAX_BETA/MIN_BETA, nem_alg.c:84-85), NaN -> 0.
 * float64 site sums in index order (the reference: float32 running sums), softmax moments taken
 * relative to max_k(beta c_ik) -- the same numbers as nem_alg.c:2172-2191 wherever the
 * reference's float exp() does not overflow (there it yields NaN and beta = 0).  beta itself and
 * its update stay float like *BetaP.  out3 (nullable) = crit, grad, dsec of the LAST iteration. */
float nemo_estim_beta(const nemo_problem *pb, const nemo_psgrad *g, const float *t, float beta,
                      double *out3)
{
    int n = pb->n, K = pb->k;
    if (!pb->row_ptr) return beta;                      /* TYPE_NONSPATIAL: nem_alg.c:2150-2151 */
    double *ctx = malloc(sizeof(double) * K);
    double crit = 0, grad = 0, dsec = 0;
    int conv = 0;
    for (int it = 0; it < g->n_iter && !conv; it++) {
        crit = grad = dsec = 0;
        for (int i = 0; i < n; i++) {
            context(pb, i, t, ctx);
            const float *ti = t + (size_t)i * K;
            double b = (double)beta, mx = -INFINITY;
            for (int k = 0; k < K; k++) if (b * ctx[k] > mx) mx = b * ctx[k];
            double se = 0, sce = 0, sc2e = 0, stc = 0;
            for (int k = 0; k < K; k++) {
                double e = exp(b * ctx[k] - mx);
                se += e; sce += ctx[k] * e; sc2e += ctx[k] * ctx[k] * e;
                stc += (double)ti[k] * ctx[k];
            }
            crit += b * stc - (mx + log(se));
            grad += stc - sce / se;
            dsec += (sc2e * se - sce * sce) / (se * se);
        }
        float gradf = (float)grad, dsecf = (float)dsec;
        if (g->step <= 0.0f) {
            dsecf = dsecf * 4;
            if (dsecf < (float)(n / 10)) dsecf = (float)(n / 10);
            beta += gradf / dsecf;
        } else {
            beta += gradf * (g->step / n);
        }
        conv = fabs(gradf) < (g->conv_thr * n);
    }
    free(ctx);
    if (out3) { out3[0] = crit; out3[1] = grad; out3[2] = dsec; }
    if (beta > 5.0f) return 5.0f;
    if (beta < -5.0f) return -5.0f;
    if (isnan(beta)) return 0.0f;
    return beta;
}

/* ------------------------------------------------------------------ EM driver (general) */
/* ClassifyByNemOneBeta (nem_alg.c:997-1192) for
 *   from_partition = 0: INIT_PARAM_FILE (1151-1169): ComputePartitionFromPara(Needinit=1) from the
 *                       all-zero classification, then NemAlgo;
 *   from_partition = 1: INIT_FILE (1091-1113): t holds the starting classification; InitPara +
 *                       MakeParaFromLabeled = one EstimPara of it (an empty class ends the call,
 *                       1338-1345), whose result NemAlgo's first M-step recomputes; then NemAlgo.
 * g != NULL: BETA_PSGRAD, beta re-estimated after every M-step (nem_alg.c:1810-1812) from the
 * classification the M-step used, and carried on (ParaP->Beta). */
int nemo_fit_ex(const nemo_problem *pb0, int from_partition, const nemo_psgrad *g, float *prop,
                float *center, float *disp, float *t, int32_t *label, nemo_result *res,
                float *beta_out)
{
    nemo_problem pbv = *pb0; const nemo_problem *pb = &pbv;
    int n = pb->n, K = pb->k;
    size_t nk = (size_t)n * K;
    float betaf = pb->row_ptr ? pb->beta : 0.0f; /* nem_exe.c:570-574 */
    double *logpf = malloc(sizeof(double) * nk);
    float *told = malloc(sizeof(float) * nk);
    double crit[6] = {0, 0, 0, 0, 0, 0};
    memset(res, 0, sizeof *res);
    int iter, converged = 0, status = NEMO_OK;

    if (from_partition) {
        status = nemo_mstep(pb, t, prop, center, disp, NULL, NULL);
        if (status != NEMO_OK) {                    /* "Class %d has no labeled observation" */
            res->status = status;
            if (beta_out) *beta_out = betaf;
            free(logpf); free(told);
            return status;
        }
    } else {
        memset(t, 0, sizeof(float) * nk);              /* calloc'd ClassifM, nem_exe.c:524-526 */
        nemo_logpf(pb, prop, center, disp, logpf);
        sweep_impl(pb, logpf, 0.0, t, label, NULL, NULL);          /* blind, nem_alg.c:1970-1977 */
        sweep_impl(pb, logpf, (double)betaf, t, label, &res->n_allnul, &res->n_ties);
        if (pb->dolog) nemo_criteria(pb, logpf, t, (double)betaf, crit);    /* WriteLogCrit, nem_alg.c:2398 */
    }

    for (iter = 1; iter <= pb->it_max && !converged && status == NEMO_OK; iter++) {
        double oldcrit = crit[3]; /* ChosenCrit(CRIT_M), nem_alg.c:1802, nem_exe.c:346 */
        memcpy(told, t, sizeof(float) * nk);
        if (!pb->param_fixed) status = nemo_mstep(pb, t, prop, center, disp, NULL, NULL);
        if (g) betaf = nemo_estim_beta(pb, g, t, betaf, NULL);
        if (status != NEMO_OK) continue; /* empty class: loop condition ends the run */
        nemo_logpf(pb, prop, center, disp, logpf);
        sweep_impl(pb, logpf, (double)betaf, t, label, &res->n_allnul, &res->n_ties);
        if (pb->conv == NEMO_CONV_CLAS) {
            converged = max_abs_diff(t, told, nk) < pb->conv_thr;
        } else if (pb->conv == NEMO_CONV_CRIT) {
            nemo_criteria(pb, logpf, t, (double)betaf, crit);
            double cur = crit[3];
            float dif = cur != 0 ? (float)fabs((cur - oldcrit) / cur) : FLT_MAX;
            converged = dif < pb->conv_thr;
        } else if (pb->dolog) {
            nemo_criteria(pb, logpf, t, (double)betaf, crit);
        }
    }
    iter -= 1;
    if (iter == 0) { /* nem_alg.c:1845-1851 */
        nemo_mstep(pb, t, prop, center, disp, NULL, NULL);
        nemo_logpf(pb, prop, center, disp, logpf);
    }
    nemo_criteria(pb, logpf, t, (double)betaf, crit);
    if (label) { /* MAP (first max); for ncem the one-hot index */
        for (int i = 0; i < n; i++) {
            int km = 0;
            for (int k = 1; k < K; k++) if (t[(size_t)i * K + k] > t[(size_t)i * K + km]) km = k;
            label[i] = km;
        }
    }
    res->status = status; res->iters = iter; res->converged = converged;
    res->U = crit[0]; res->D = crit[1]; res->L = crit[2]; res->M = crit[3];
    res->Z = crit[4]; res->G = crit[5];
    if (beta_out) *beta_out = betaf;
    free(logpf); free(told);
    return status;
}

/* ClassifyByNemOneBeta INIT_PARAM_FILE branch (nem_alg.c:1151-1169):
 * ComputePartitionFromPara(Needinit=1) (nem_alg.c:1951-1989) then NemAlgo (1746-1879). */
int nemo_fit(const nemo_problem *pb0, float *prop, float *center, float *disp,
             float *t, int32_t *label, nemo_result *res)
{
    return nemo_fit_ex(pb0, 0, NULL, prop, center, disp, t, label, res, NULL);
}

/* ------------------------------------------------------------------ beta heuristics */
/* ClassifyByNemHeuBeta (nem_alg.c:731-992), BETA_HEUD (mode 0, Hathaway criterion D: stop at the
 * first drop of its slope below -ddrop*N, else threshold its total loss) and BETA_HEUL (mode 1,
 * mixture likelihood L: stop when it falls lloss*N under its maximum).  One complete fit per
 * tested beta = 0, step, 2 step ... <= max, each from the all-zero classification but from the
 * PARAMETERS THE PREVIOUS FIT LEFT (StatModelP->Para is in/out, 826-833); a fit that ends with an
 * empty class is skipped.  Final fit at the estimated beta: from the classification saved before
 * the drop (InitMode = INIT_FILE, 958-963), or -- D heuristic without a detected drop -- again
 * from the all-zero classification (952-954).  Criteria compared in float like criV/slopes.
 * trace (nullable, cap entries each): tested betas and their criterion. */
int nemo_fit_heuristic(const nemo_problem *pb0, int mode, const nemo_heu *hp, float *prop,
                       float *center, float *disp, float *t, int32_t *label, nemo_result *res,
                       float *beta_est, int *n_tested, float *beta_trace, float *crit_trace,
                       int cap)
{
    nemo_problem pb = *pb0;
    int n = pb.n, K = pb.k;
    size_t nk = (size_t)n * K;
    int nbtamax = (int)(hp->max / hp->step) + 1;
    float *btaV = calloc(nbtamax + 2, sizeof(float)), *criV = calloc(nbtamax + 2, sizeof(float));
    float *best = calloc(nk, sizeof(float));
    int nbta = 0, stop = 0, Dincreas = 0, Ddrop = 0, Lfound = 0;
    float Dmin = 0.0f, prevSlope = NAN, thisSlope = NAN, Lmax = NAN, btaEst = NAN;
    float DdropThres = -hp->ddrop * n, LlossThres = hp->lloss * n;

    for (float bt = 0.0f; bt <= hp->max && !stop; bt += hp->step) {
        pb.beta = bt;
        if (nemo_fit_ex(&pb, 0, NULL, prop, center, disp, t, label, res, NULL) != NEMO_OK) continue;
        if (nbta > nbtamax) break;                 /* cannot happen (float steps), guards the arrays */
        nbta++;
        btaV[nbta] = bt;
        if (mode == 0) {
            criV[nbta] = (float)res->D;
            if (criV[nbta] < Dmin) Dmin = criV[nbta];
            if (nbta >= 2) {
                prevSlope = thisSlope;
                thisSlope = (criV[nbta] - criV[nbta - 1]) / (btaV[nbta] - btaV[nbta - 1]);
                if (thisSlope >= 0.5 * n) Dincreas = 1;
            }
            if (nbta >= 3) {
                if (!Ddrop && !Dincreas) {
                    if ((thisSlope - prevSlope) < DdropThres) {
                        Ddrop = 1; stop = 1;
                        btaEst = btaV[nbta - 1];
                    } else
                        memcpy(best, t, sizeof(float) * nk);
                }
            } else
                memcpy(best, t, sizeof(float) * nk);
        } else {
            criV[nbta] = (float)res->L;
            if (nbta < 2) {
                Lmax = criV[nbta];
                memcpy(best, t, sizeof(float) * nk);
            } else {
                if (criV[nbta] > Lmax) Lmax = criV[nbta];
                if (!Lfound) {
                    if (criV[nbta] < Lmax - LlossThres) {
                        Lfound = 1; stop = 1;
                        btaEst = btaV[nbta - 1];
                    } else
                        memcpy(best, t, sizeof(float) * nk);
                }
            }
        }
    }
    int from_partition;
    if (mode == 0 && !Ddrop) {
        float DThres = criV[1] - (criV[1] - Dmin) * hp->dloss;
        int ibta, found = 0;
        for (ibta = 1; ibta <= nbta && !found; ibta++) found = criV[ibta] <= DThres;
        btaEst = found ? btaV[ibta - 2] : 0.0f;    /* "heuristic failed to detect beta" */
        from_partition = 0;
    } else {
        memcpy(t, best, sizeof(float) * nk);
        from_partition = 1;
    }
    if (mode == 1 && !Lfound) btaEst = btaV[nbta];
    if (n_tested) *n_tested = nbta;
    for (int i = 0; i < nbta && i < cap; i++) {
        if (beta_trace) beta_trace[i] = btaV[i + 1];
        if (crit_trace) crit_trace[i] = criV[i + 1];
    }
    if (beta_est) *beta_est = btaEst;
    pb.beta = btaEst;
    int rc = nemo_fit_ex(&pb, from_partition, NULL, prop, center, disp, t, label, res, NULL);
    free(btaV); free(criV); free(best);
    return rc;
}
