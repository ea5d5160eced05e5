/*
 * nem_io.h -- internal: readers of the NEM input files and writers of its result files
 * (host C; formats documented in the reference's NEM/nem_hlp.c:309-513).
 */
#ifndef NEM_IO_H
#define NEM_IO_H
#include <stdint.h>
#include <stdio.h>

/* all functions return a NEMB_* status code (include/nem_b200.h) and print the reason to `err` */
int nemio_read_str(const char *base, FILE *err, char *type /*'S','N','I'*/, int *n, int *d,
                   char *comment, int comment_len);
int nemio_read_dat(const char *path, FILE *err, int n, int d, int wpr, uint32_t **packed_out,
                   int n_threads);
int nemio_read_nei(const char *base, FILE *err, int n, int32_t **row_ptr, int32_t **col,
                   float **wgt, int *max_neigh, char *comment, int comment_len);
int nemio_read_m(const char *path, FILE *err, int k, int d, int *flag, float *prop, float *center,
                 float *disp);

int nemio_write_uf(const char *path, FILE *err, int n, int k, const float *t);
int nemio_write_cf(const char *path, FILE *err, int n, const int32_t *label);
int nemio_write_mf(const char *path, FILE *err, int k, int d, const double crit_udlm[4], float beta,
                   const float *prop, const float *center, const float *disp);
/* beta_mode 0..3 = fix, psgrad, heu_d, heu_l: the "Beta (...)" line (nem_exe.c:1714) */
int nemio_write_mf_mode(const char *path, FILE *err, int k, int d, const double crit_udlm[4],
                        float beta, int beta_mode, const float *prop, const float *center,
                        const float *disp);
#endif
