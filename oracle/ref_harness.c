/*
 * oracle/ref_harness.c  --  TEST INFRASTRUCTURE (oracle #1 in-process harness), not product code.
 *
 * Calls the reference's linkable entry ClassifyByNem()
 * (/root/reference/ppanggolin/NEM/nem_alg.h:10-18, nem_alg.c:546-584) directly with a
 * hand-filled NemParaT so that the knobs nem() hides are reachable:
 *     SiteUpdate  (nem_typ.h:340, default UPDATE_SEQ   nem_exe.c:360)
 *     TieRule     (nem_typ.h:341, default TIE_RANDOM   nem_exe.c:361)
 *     Seed        (nem_typ.h:332, default time(NULL)   nem_exe.c:353)
 *     NbEIters    (nem_typ.h:330)
 * and dumps ClassifM / parameters / criteria at full float precision instead of the
 * "%5.3f" text of SaveResults (nem_exe.c:1677).
 *
 * The file readers below are OURS (the reference's are static in nem_exe.c); they read
 * exactly the files PPanGGOLiN writes (ppanggolin.py:829-930).  Only the reference's
 * headers are included, its .c files are compiled in place by oracle/Makefile.
 *
 * usage: nem_ref_harness base K algo beta conv thr itmax family prop disp init
 *                        update(seq|para) tie(random|first) seed out_prefix
 *                        [betamode(fix|psgrad|heu_d|heu_l)  [p1 p2 p3 p4 p5]]
 *        betamode (StatModelT.Spec.BetaModel, nem_typ.h:128-135; the CLI's -B, nem_hlp.c:220-225)
 *        with psgrad: p1..p3 = nit conv step (-G, BtaPsGradT nem_typ.h:300-308)
 *        with heu_*:  p1..p5 = bstep bmax ddrop dloss lloss (-H, nem_typ.h:319-323)
 * writes out_prefix.cm.f32   raw float32 ClassifM [N*K]
 *        out_prefix.par.f32  raw float32: Prop_K[K] Center_KD[K*D] Disp_KD[K*D]
 *        out_prefix.txt      "status U D L M Z G beta"   (beta = StatModelT.Para.Beta on return)
 */
#include "nem_typ.h"
#include "nem_alg.h"
#include "genmemo.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

static int find(const char *s, const char *const *tab, int n)
{
    for (int i = 0; i < n; i++)
        if (!strcmp(s, tab[i])) return i;
    fprintf(stderr, "harness: unknown option value '%s'\n", s);
    exit(98);
}

static void die(const char *msg, const char *arg)
{
    fprintf(stderr, "harness: %s %s\n", msg, arg ? arg : "");
    exit(97);
}

int main(int argc, char **argv)
{
    if (argc < 16) die("need 15 arguments", NULL);
    const char *base = argv[1];
    int K = atoi(argv[2]);
    static NemParaT P;
    static SpatialT S;
    static DataT X;
    static StatModelT M;
    static CriterT C;
    char name[1024];

    /* reference progress text (incl. "NEM converged after %d iterations") */
    snprintf(name, sizeof name, "%s.stderr", argv[15]);
    out_stderr = fopen(name, "w");
    if (!out_stderr) die("cannot open", name);

    memset(&P, 0, sizeof P);
    P.Algo = find(argv[3], AlgoStrVC, ALGO_NB);
    M.Para.Beta = (float)atof(argv[4]);
    P.CvTest = find(argv[5], CvTestStrVC, CVTEST_NB);
    P.CvThres = (float)atof(argv[6]);
    P.NbIters = atoi(argv[7]);
    M.Spec.K = K;
    M.Spec.ClassFamily = find(argv[8], FamilyStrVC, FAMILY_NB);
    M.Spec.ClassPropor = find(argv[9], ProporStrVC, PROPOR_NB);
    M.Spec.ClassDisper = find(argv[10], DisperStrVC, DISPER_NB);
    M.Spec.BetaModel = BETA_FIX;
    P.InitMode = atoi(argv[11]);
    P.SiteUpdate = find(argv[12], UpdateStrVC, UPDATE_NB);
    P.TieRule = find(argv[13], TieStrVC, TIE_NB);
    P.Seed = atol(argv[14]);
    const char *outp = argv[15];

    /* remaining defaults as nem() sets them (nem_exe.c:331-362) */
    P.BtaHeuStep = DEFAULT_BTAHEUSTEP; P.BtaHeuMax = DEFAULT_BTAHEUMAX;
    P.BtaHeuDDrop = DEFAULT_BTAHEUDDROP; P.BtaHeuDLoss = DEFAULT_BTAHEUDLOSS;
    P.BtaHeuLLoss = DEFAULT_BTAHEULLOSS;
    P.BtaPsGrad.NbIter = DEFAULT_BTAGRADNIT; P.BtaPsGrad.ConvThres = DEFAULT_BTAGRADCVTH;
    P.BtaPsGrad.Step = DEFAULT_BTAGRADSTEP; P.BtaPsGrad.RandInit = DEFAULT_BTAGRADRAND;
    P.Crit = DEFAULT_CRIT; P.DoLog = FALSE; P.NbEIters = DEFAULT_NBEITERS;
    P.NbRandomInits = DEFAULT_NBRANDINITS; P.Format = FORMAT_FUZZY;
    P.MissMode = DEFAULT_MISSING; P.ParamFileMode = NO_PARAM_FILE;
    P.SortedVar = DEFAULT_SORTEDVAR; P.NeighSpec = NEIGH_FILE;
    P.VisitOrder = ORDER_DIRECT; P.Debug = FALSE;
    if (argc > 16) {
        M.Spec.BetaModel = find(argv[16], BtaStrVC, BETA_NB);
        if (M.Spec.BetaModel == BETA_PSGRAD && argc > 19) {
            P.BtaPsGrad.NbIter = atoi(argv[17]); P.BtaPsGrad.ConvThres = (float)atof(argv[18]);
            P.BtaPsGrad.Step = (float)atof(argv[19]);
        }
        if ((M.Spec.BetaModel == BETA_HEUD || M.Spec.BetaModel == BETA_HEUL) && argc > 21) {
            P.BtaHeuStep = (float)atof(argv[17]); P.BtaHeuMax = (float)atof(argv[18]);
            P.BtaHeuDDrop = (float)atof(argv[19]); P.BtaHeuDLoss = (float)atof(argv[20]);
            P.BtaHeuLLoss = (float)atof(argv[21]);
        }
    }

    /* .str : "S|N  N  D" */
    int N, D; char type[64];
    snprintf(name, sizeof name, "%s.str", base);
    FILE *f = fopen(name, "r"); if (!f) die("cannot open", name);
    if (fscanf(f, "%63s %d %d", type, &N, &D) != 3) die("bad str file", name);
    fclose(f);
    S.Type = (type[0] == 'N' || type[0] == 'n') ? TYPE_NONSPATIAL : TYPE_SPATIAL;
    X.NbPts = N; X.NbVars = D; X.NbMiss = 0; X.LabelV = NULL;

    /* .dat */
    X.PointsM = malloc(sizeof(float) * (size_t)N * D);
    snprintf(name, sizeof name, "%s.dat", base);
    f = fopen(name, "r"); if (!f) die("cannot open", name);
    for (size_t i = 0; i < (size_t)N * D; i++)
        if (fscanf(f, "%f", &X.PointsM[i]) != 1) die("short dat file", name);
    fclose(f);
    X.SiteVisitV = malloc(sizeof(int) * N);
    for (int i = 0; i < N; i++) X.SiteVisitV[i] = i;

    /* model arrays */
    M.Para.Prop_K = calloc(K, sizeof(float));
    M.Para.Disp_KD = calloc((size_t)K * D, sizeof(float));
    M.Para.Center_KD = calloc((size_t)K * D, sizeof(float));
    M.Para.NbObs_K = calloc(K, sizeof(float));
    M.Para.NbObs_KD = calloc((size_t)K * D, sizeof(float));
    M.Para.Iner_KD = calloc((size_t)K * D, sizeof(float));
    M.Desc.DispSam_D = calloc(D, sizeof(float));
    M.Desc.MiniSam_D = calloc(D, sizeof(float));
    M.Desc.MaxiSam_D = calloc(D, sizeof(float));

    /* .m (init_mode 2) : flag, K-1 proportions, K*D centres, K*D dispersions */
    if (P.InitMode == INIT_PARAM_FILE) {
        snprintf(name, sizeof name, "%s.m", base);
        f = fopen(name, "r"); if (!f) die("cannot open", name);
        int flag; float pk = 1.f;
        if (fscanf(f, "%d", &flag) != 1) die("bad m file", name);
        P.ParamFileMode = (flag == 2) ? PARAM_FILE_FIX : PARAM_FILE_INIT;
        for (int k = 0; k < K - 1; k++) {
            if (fscanf(f, "%f", &M.Para.Prop_K[k]) != 1) die("bad m file", name);
            pk = pk - M.Para.Prop_K[k];
        }
        M.Para.Prop_K[K - 1] = pk;
        for (int i = 0; i < K * D; i++)
            if (fscanf(f, "%f", &M.Para.Center_KD[i]) != 1) die("bad m file", name);
        for (int i = 0; i < K * D; i++)
            if (fscanf(f, "%f", &M.Para.Disp_KD[i]) != 1) die("bad m file", name);
        fclose(f);
    }

    /* .nei : weighted flag, then "id nb n_1..n_nb [w_1..w_nb]" (1-based) */
    S.MaxNeighs = 0;
    if (S.Type == TYPE_SPATIAL) {
        PtNeighsT *pn = calloc(N, sizeof(PtNeighsT));
        S.NeighData.PtsNeighsV = pn;
        snprintf(name, sizeof name, "%s.nei", base);
        f = fopen(name, "r"); if (!f) die("cannot open", name);
        int weighted, id, nb;
        if (fscanf(f, "%d", &weighted) != 1) die("bad nei file", name);
        while (fscanf(f, "%d %d", &id, &nb) == 2) {
            NeighT *v = calloc(nb > 0 ? nb : 1, sizeof(NeighT));
            int *ids = malloc(sizeof(int) * (nb > 0 ? nb : 1));
            for (int j = 0; j < nb; j++)
                if (fscanf(f, "%d", &ids[j]) != 1) die("bad nei file", name);
            int nv = 0;
            for (int j = 0; j < nb; j++) {
                float w = 1.f;
                if (weighted && fscanf(f, "%g", &w) != 1) die("bad nei file", name);
                if (ids[j] >= 1 && ids[j] <= N && w != 0.f) {
                    v[nv].Index = ids[j] - 1; v[nv].Weight = w; nv++;
                }
            }
            free(ids);
            pn[id - 1].NeighsV = v; pn[id - 1].NbNeigh = nv;
            if (nv > S.MaxNeighs) S.MaxNeighs = nv;
        }
        fclose(f);
    } else {
        M.Para.Beta = 0.f;
    }

    float *CM = calloc((size_t)N * K, sizeof(float));
    memset(&C, 0, sizeof C);
    C.Errinfo.Kc = K; C.Errinfo.Kr = 0; C.Errinfo.Km = K; C.Errinfo.TieRule = P.TieRule;

    srandom(P.Seed);
    int sts = ClassifyByNem(&P, &S, &X, &M, CM, &C);

    snprintf(name, sizeof name, "%s.cm.f32", outp);
    f = fopen(name, "wb"); fwrite(CM, sizeof(float), (size_t)N * K, f); fclose(f);
    snprintf(name, sizeof name, "%s.par.f32", outp);
    f = fopen(name, "wb");
    fwrite(M.Para.Prop_K, sizeof(float), K, f);
    fwrite(M.Para.Center_KD, sizeof(float), (size_t)K * D, f);
    fwrite(M.Para.Disp_KD, sizeof(float), (size_t)K * D, f);
    fclose(f);
    snprintf(name, sizeof name, "%s.txt", outp);
    f = fopen(name, "w");
    fprintf(f, "%d %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", sts, C.U, C.D, C.L, C.M, C.Z, C.G,
            M.Para.Beta);
    fclose(f);
    fclose(out_stderr);
    return 0;
}
