/* Test-only harness: runs the host readers of csrc/nem_io.c on <base>.{str,dat,nei,m} and prints the
 * status of each (compiled with -fsanitize=address,undefined by tests/test_host_io_fuzz.py). */
#include "nem_io.h"
#include "nem_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char **argv)
{
    if (argc < 4) return 64;
    const char *base = argv[1];
    int k = atoi(argv[2]), threads = atoi(argv[3]);
    char type = 0, comment[64], path[512];
    int n = 0, d = 0, rc;
    rc = nemio_read_str(base, stderr, &type, &n, &d, comment, sizeof comment);
    printf("str %d %c %d %d\n", rc, type ? type : '?', n, d);
    if (rc != NEMB_OK) return 0;
    if (n > 2000000 || d > 200000) return 0;   /* a harness guard, not a library limit */
    int wpr = ((d + 31) / 32 + 3) & ~3;
    uint32_t *x = NULL;
    snprintf(path, sizeof path, "%s.dat", base);
    rc = nemio_read_dat(path, stderr, n, d, wpr, &x, threads);
    printf("dat %d\n", rc);
    if (rc == NEMB_OK) {
        unsigned long long sum = 0;
        for (size_t i = 0; i < (size_t)n * wpr; i++) sum += x[i];   /* touch every word */
        printf("datsum %llu\n", sum);
    }
    free(x);
    int32_t *rp = NULL, *col = NULL;
    float *w = NULL;
    int mx = 0;
    rc = nemio_read_nei(base, stderr, n, &rp, &col, &w, &mx, comment, sizeof comment);
    printf("nei %d %d\n", rc, mx);
    if (rc == NEMB_OK) {
        double s = 0;
        for (int i = 0; i < n; i++)
            for (int e = rp[i]; e < rp[i + 1]; e++) {
                if (col[e] < 0 || col[e] >= n) { printf("BAD col\n"); return 70; }
                s += w[e];
            }
        printf("neisum %d %g\n", rp[n], s);
    }
    free(rp); free(col); free(w);
    float *prop = calloc(k, sizeof(float)), *cen = calloc((size_t)k * d + 1, sizeof(float)),
          *dis = calloc((size_t)k * d + 1, sizeof(float));
    int flag = 0;
    snprintf(path, sizeof path, "%s.m", base);
    rc = nemio_read_m(path, stderr, k, d, &flag, prop, cen, dis);
    printf("m %d %d\n", rc, flag);
    free(prop); free(cen); free(dis);
    return 0;
}
