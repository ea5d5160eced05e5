/*
 * oracle/ref_cli_main.c  --  TEST INFRASTRUCTURE (oracle #1 driver), not product code.
 *
 * A main() that forwards argv to the reference's own entry point
 *   int nem(Fname, nk, algo, beta, convergence, convergence_th, format, it_max,
 *           dolog, model_family, proportion, dispersion, init_mode)
 * declared at /root/reference/ppanggolin/NEM/nem_exe.h:23-35.  The reference
 * snapshot ships no main(); this file is ours and is linked against the reference
 * objects compiled in place by oracle/Makefile (outputs only under oracle/_ref/).
 *
 * usage: nem_ref_cli Fname K algo beta conv thr format itmax dolog family prop disp init
 */
#include <stdio.h>
#include <stdlib.h>

extern int nem(const char *Fname, const int nk, const char *algo, const float beta,
               const char *convergence, const float convergence_th, const char *format,
               const int it_max, const int dolog, const char *model_family,
               const char *proportion, const char *dispersion, const int init_mode);

int main(int argc, char **argv)
{
    if (argc < 14) {
        fprintf(stderr,
                "usage: %s Fname K algo beta conv thr format itmax dolog family prop disp init\n",
                argv[0]);
        return 99;
    }
    return nem(argv[1], atoi(argv[2]), argv[3], (float)atof(argv[4]), argv[5],
               (float)atof(argv[6]), argv[7], atoi(argv[8]), atoi(argv[9]), argv[10],
               argv[11], argv[12], atoi(argv[13]));
}
