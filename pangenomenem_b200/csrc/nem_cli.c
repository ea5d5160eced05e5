/*
 * nem_cli.c -- `nem_exe file K [options]`: the historical command line of NEM, which survives
 * in the reference only as help text (NEM/nem_hlp.c:109-289).  Thin wrapper over nem_b200_ex();
 * it also reaches the knobs nem() hides (-U update, -S seed) -- SURVEY.md section 8b "CLI".
 *
 * Supported: -a {nem ncem}  -b beta  -c {none clas crit} [thr]  -f {hard fuzzy}  -i itmax
 *            -l {y n}  -m bern {p_ pk} {s__ sk_ s_d skd}  -s m <ignored file, uses file.m> | -s r <n>
 *            -t first  -U {seq para}  -S seed  -W {auto level spec}  -g device
 *            -B {fix psgrad heu_d heu_l}  -G nit conv step rand  -H bstep bmax ddrop dloss lloss
 *            (beta estimation, nem_hlp.c:220-245; `rand` must be 0)
 * Anything else of the 1.08 syntax (gem, norm/lapl, image data, -o) is off the PPanGGOLiN path
 * and is rejected.
 */
#include "nem_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int usage(const char *argv0)
{
    fprintf(stderr,
            "usage: %s file K [-a nem|ncem] [-b beta] [-c none|clas|crit thr] [-f hard|fuzzy]\n"
            "          [-i itmax] [-l y|n] [-m bern p_|pk s__|sk_|s_d|skd] [-s m file | -s r n]\n"
            "          [-t first] [-U seq|para] [-S seed] [-W auto|level|spec] [-g device]\n"
            "          [-B fix|psgrad|heu_d|heu_l] [-G nit conv step 0] [-H bstep bmax ddrop dloss lloss]\n"
            "%s\n", argv0, nemb_version());
    return 2;
}

int main(int argc, char **argv)
{
    /* helper of a forked nem() caller (csrc/nem_api.c "forked callers") */
    if (argc == 2 && !strcmp(argv[1], "--serve")) return nem_b200_serve(3, 4);
    if (argc < 3) return usage(argv[0]);
    const char *file = argv[1];
    int k = atoi(argv[2]);
    const char *algo = "nem", *conv = "clas", *fmt = "hard", *fam = "bern", *prop = "p_", *disp = "s__";
    float beta = 1.0f, thr = 0.01f;       /* DEFAULT_BETA, DEFAULT_CVTHRES (nem_typ.h:69,81) */
    int itmax = 100, dolog = 0, init = 2;
    nem_b200_extra ex;
    memset(&ex, 0, sizeof ex);
    ex.device = -1;
    for (int i = 3; i < argc; i++) {
        const char *a = argv[i];
#define ARG() ((i + 1 < argc) ? argv[++i] : (usage(argv[0]), exit(2), ""))
        if (!strcmp(a, "-a")) algo = ARG();
        else if (!strcmp(a, "-b")) beta = (float)atof(ARG());
        else if (!strcmp(a, "-c")) { conv = ARG(); if (strcmp(conv, "none")) thr = (float)atof(ARG()); }
        else if (!strcmp(a, "-f")) fmt = ARG();
        else if (!strcmp(a, "-i")) itmax = atoi(ARG());
        else if (!strcmp(a, "-l")) dolog = ARG()[0] == 'y';
        else if (!strcmp(a, "-m")) { fam = ARG(); prop = ARG(); disp = ARG(); }
        else if (!strcmp(a, "-s")) {
            const char *m = ARG();
            if (!strcmp(m, "m")) { init = 2; (void)ARG(); }
            else if (!strcmp(m, "r")) { init = 1; ex.n_random_inits = atoi(ARG()); }
            else { fprintf(stderr, "init mode -s %s is not available in the B200 engine\n", m); return 2; }
        }
        else if (!strcmp(a, "-t")) { if (strcmp(ARG(), "first")) { fprintf(stderr, "only -t first is available (ties are deterministic)\n"); return 2; } }
        else if (!strcmp(a, "-U")) ex.update = !strcmp(ARG(), "para");
        else if (!strcmp(a, "-S")) ex.seed = atoll(ARG());
        else if (!strcmp(a, "-W")) { const char *w = ARG(); ex.sweep_impl = !strcmp(w, "level") ? 1 : !strcmp(w, "spec") ? 2 : 0; }
        else if (!strcmp(a, "-g")) ex.device = atoi(ARG());
        else if (!strcmp(a, "-B")) {
            const char *b = ARG();
            ex.beta_mode = !strcmp(b, "fix") ? 0 : !strcmp(b, "psgrad") ? 1 : !strcmp(b, "heu_d") ? 2 : !strcmp(b, "heu_l") ? 3 : -1;
            if (ex.beta_mode < 0) { fprintf(stderr, "unknown beta estimation mode %s\n", b); return 2; }
        }
        else if (!strcmp(a, "-G")) {
            ex.grad_n_iter = atoi(ARG()); ex.grad_conv = (float)atof(ARG()); ex.grad_step = (float)atof(ARG());
            if (atoi(ARG()) != 0) { fprintf(stderr, "-G ... rand: a random initial beta is not available\n"); return 2; }
        }
        else if (!strcmp(a, "-H")) {
            ex.heu_step = (float)atof(ARG()); ex.heu_max = (float)atof(ARG()); ex.heu_ddrop = (float)atof(ARG());
            ex.heu_dloss = (float)atof(ARG()); ex.heu_lloss = (float)atof(ARG());
        }
        else if (!strcmp(a, "-v")) { printf("%s\n", nemb_version()); return 0; }
        else { fprintf(stderr, "option %s is not available in the B200 engine\n", a); return usage(argv[0]); }
#undef ARG
    }
    return nem_b200_ex(file, k, algo, beta, conv, thr, fmt, itmax, dolog, fam, prop, disp, init, &ex);
}
