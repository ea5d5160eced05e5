/*
 * nem_io.c -- the NEM loader (host side) and result writers.
 *
 * Readers restate, with the same tolerance, the reference's
 *   ReadOpeningComments  NEM/lib_io.c:32-87     leading '#' comment lines
 *   ReadStrFile          NEM/nem_exe.c:739-830  "S|N|I  N  D"
 *   ReadMatrixFile       NEM/nem_exe.c:834-898  whitespace-separated "%f" tokens, short file = error
 *   ReadParamFile        NEM/nem_exe.c:973-1091 flag, K-1 proportions, K*D centres, K*D dispersions
 *   ReadPtsNeighs        NEM/nem_exe.c:1342-1478 weighted flag, then "id nb n_1..n_nb [w_1..w_nb]"
 * but produce the engine's HBM layout directly: X is bit-packed while it is parsed (never a
 * float matrix), the neighbour lists become one CSR in file order.
 * Writers restate SaveResults (NEM/nem_exe.c:1596-1781) format for format.
 *
 * Deliberate deviations (all outside what PPanGGOLiN can emit, ppanggolin.py:829-930):
 *   - .dat cells must be 0 or 1 (Bernoulli data); NaN / other values are rejected with E_FILE
 *     instead of being carried as floats;
 *   - a .nei entry is dropped when its index is out of 1..N OR its weight is 0; the reference
 *     compacts the two lists independently (nem_exe.c:1416-1446) which misaligns such files;
 *   - a record for a point id outside 1..N is an error (the reference writes out of bounds).
 */
#include "nem_io.h"
#include "nem_b200.h"

#include <ctype.h>
#include <errno.h>
#include <fcntl.h>
#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

/* ------------------------------------------------------------------ mapped files + tokens */
typedef struct {
    const char *p, *end, *base;
    size_t len;
} cursor;

static int map_file(const char *path, cursor *c)
{
    int fd = open(path, O_RDONLY);
    if (fd < 0) return -1;
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return -1; }
    c->len = (size_t)st.st_size;
    if (c->len == 0) {
        c->base = c->p = c->end = "";
        close(fd);
        return 0;
    }
    void *m = mmap(NULL, c->len, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return -1;
    madvise(m, c->len, MADV_SEQUENTIAL);
    c->base = c->p = m;
    c->end = c->p + c->len;
    return 0;
}

static void unmap_file(cursor *c)
{
    if (c->len) munmap((void *)c->base, c->len);
    c->len = 0;
}

static inline int is_ws(char ch) { return ch == ' ' || (ch >= '\t' && ch <= '\r'); }

static inline int next_token(cursor *c, const char **tok, size_t *len)
{
    const char *p = c->p, *e = c->end;
    while (p < e && is_ws(*p)) p++;
    if (p >= e) { c->p = p; return 0; }
    const char *q = p;
    while (q < e && !is_ws(*q)) q++;
    *tok = p; *len = (size_t)(q - p);
    c->p = q;
    return 1;
}

/* plain decimal tokens (what PPanGGOLiN writes: ids, counts, integer coverages) are converted
 * in place; anything else goes through strtol / strtof like the reference's fscanf */
static inline int tok_digits(const char *t, size_t len, long *out)
{
    size_t i = 0;
    int neg = 0;
    if (len > 1 && (t[0] == '-' || t[0] == '+')) { neg = t[0] == '-'; i = 1; }
    if (len - i == 0 || len - i > 15) return 0;
    long v = 0;
    for (; i < len; i++) {
        unsigned dgt = (unsigned)(t[i] - '0');
        if (dgt > 9) return 0;
        v = v * 10 + (long)dgt;
    }
    *out = neg ? -v : v;
    return 1;
}

static int tok_int(const char *t, size_t len, long *out)
{
    if (tok_digits(t, len, out)) return 1;
    char buf[64];
    if (len == 0 || len >= sizeof buf) return 0;
    memcpy(buf, t, len); buf[len] = 0;
    char *endp;
    errno = 0;
    long v = strtol(buf, &endp, 10);
    if (endp == buf || *endp) return 0;
    *out = v;
    return 1;
}

static int tok_float(const char *t, size_t len, float *out)
{
    long iv;
    if (tok_digits(t, len, &iv)) { *out = (float)iv; return 1; }   /* < 10^15: one correct rounding, as strtof */
    char buf[128];
    if (len == 0 || len >= sizeof buf) return 0;
    memcpy(buf, t, len); buf[len] = 0;
    char *endp;
    float v = strtof(buf, &endp);
    if (endp == buf || *endp) return 0;
    *out = v;
    return 1;
}

/* skip the opening lines that start with '#', keep their text (lib_io.c:32-87) */
static void skip_comments(cursor *c, char *comment, int comment_len)
{
    if (comment && comment_len > 0) comment[0] = 0;
    while (c->p < c->end && *c->p == '#') {
        const char *q = c->p;
        while (q < c->end && *q != '\n') q++;
        if (q < c->end) q++;
        if (comment) {
            size_t have = strlen(comment), add = (size_t)(q - c->p - 1);
            if (have + add + 1 < (size_t)comment_len) {
                memcpy(comment + have, c->p + 1, add);
                comment[have + add] = 0;
            }
        }
        c->p = q;
    }
}

/* ------------------------------------------------------------------ .str */
int nemio_read_str(const char *base, FILE *err, char *type, int *n, int *d, char *comment,
                   int comment_len)
{
    char path[4200];
    snprintf(path, sizeof path, "%s.str", base);
    cursor c;
    if (map_file(path, &c) != 0) {
        fprintf(err, "File str %s does not exist\n", path);
        return NEMB_E_FILE;
    }
    skip_comments(&c, comment, comment_len);
    const char *t; size_t len; long v[3]; int nf = 0;
    char ty = 0;
    if (next_token(&c, &t, &len)) { ty = (char)toupper((unsigned char)t[0]); if (len != 1) ty = 0; }
    while (nf < 3 && next_token(&c, &t, &len) && tok_int(t, len, &v[nf])) nf++;
    unmap_file(&c);
    if (!ty || nf < 2) {
        fprintf(err, "Structure file (%s) not enough fields\n", path);
        return NEMB_E_FILE;
    }
    /* N and D are int in the engine's ABI (nem_exe.h:23-35): a header beyond that is an error,
     * not a silent truncation */
    for (int q = 0; q < nf; q++)
        if (v[q] < 0 || v[q] > INT_MAX || (ty == 'I' && q == 1 && v[0] * v[1] > INT_MAX)) {
            fprintf(err, "Structure file (%s): sizes out of range\n", path);
            return NEMB_E_FILE;
        }
    if (ty == 'I') {
        if (nf < 3) { fprintf(err, "Structure file (%s) not enough fields\n", path); return NEMB_E_FILE; }
        *type = 'I'; *n = (int)(v[0] * v[1]); *d = (int)v[2];
        return NEMB_OK;
    }
    if (ty != 'S' && ty != 'N') {
        fprintf(err, "Data type %c unknown in file %s\n", ty, path);
        return NEMB_E_FILE;
    }
    *type = ty; *n = (int)v[0]; *d = (int)v[1];
    if (*n <= 0 || *d <= 0) {
        fprintf(err, "Structure file (%s): N and D must be > 0\n", path);
        return NEMB_E_FILE;
    }
    return NEMB_OK;
}

/* ------------------------------------------------------------------ .dat -> packed bits */
typedef struct {
    const char *src; int d, wpr, r0, r1; uint32_t *out; int bad;
} dat_job;

#if defined(__SSE2__)
#include <emmintrin.h>
static uint8_t even_bits_lut[256];      /* bits 0,2,4,6 of the index packed into bits 0..3 */
static pthread_once_t even_bits_once = PTHREAD_ONCE_INIT;
static void even_bits_init(void)
{
    for (int v = 0; v < 256; v++)
        even_bits_lut[v] = (uint8_t)((v & 1) | ((v >> 1) & 2) | ((v >> 2) & 4) | ((v >> 3) & 8));
}
/* 32 cells = 64 bytes "c s c s ..." -> one word; *bad is raised unless every c is '0'/'1' and
 * every s is white space (space or \t..\r) */
static inline uint32_t dat_word_sse2(const unsigned char *q, unsigned *bad)
{
    const __m128i one = _mm_set1_epi8('1'), even = _mm_set1_epi16(0x00ff), lsb = _mm_set1_epi16(0x0001);
    const __m128i sp = _mm_set1_epi8(' '), nine = _mm_set1_epi8(9), four = _mm_set1_epi8(4);
    uint32_t word = 0;
    unsigned ok = 0xffffu;
    for (int part = 0; part < 4; part++) {
        __m128i v = _mm_loadu_si128((const __m128i *)(q + 16 * part));
        unsigned m1 = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(v, one));           /* byte == '1' */
        __m128i cell_ok = _mm_cmpeq_epi8(_mm_or_si128(v, lsb), one);                /* even bytes: '0' | 1 == '1' */
        __m128i off = _mm_sub_epi8(v, nine);                                        /* odd bytes: \t..\r or ' ' */
        __m128i ws_ok = _mm_or_si128(_mm_cmpeq_epi8(_mm_min_epu8(off, four), off), _mm_cmpeq_epi8(v, sp));
        __m128i good = _mm_or_si128(_mm_and_si128(cell_ok, even), _mm_andnot_si128(even, ws_ok));
        ok &= (unsigned)_mm_movemask_epi8(good);
        word |= (uint32_t)(even_bits_lut[m1 & 0xff] | (even_bits_lut[m1 >> 8] << 4)) << (8 * part);
    }
    *bad |= ok ^ 0xffffu;
    return word;
}
#endif

/* fast path: PPanGGOLiN's exact layout, one char per cell + one separator (ppanggolin.py:850) */
static void *dat_worker(void *arg)
{
    dat_job *j = arg;
    int d = j->d, wpr = j->wpr;
    size_t stride = (size_t)2 * d;
#if defined(__SSE2__)
    pthread_once(&even_bits_once, even_bits_init);
#endif
    for (int r = j->r0; r < j->r1; r++) {
        const unsigned char *p = (const unsigned char *)j->src + (size_t)r * stride;
        uint32_t *o = j->out + (size_t)r * wpr;
        unsigned bad = 0;
        for (int w = 0; w * 32 < d; w++) {
            int nb = d - w * 32 < 32 ? d - w * 32 : 32;
            uint32_t word = 0;
            const unsigned char *q = p + (size_t)w * 64;
#if defined(__SSE2__)
            if (nb == 32) { o[w] = dat_word_sse2(q, &bad); continue; }
#endif
            for (int b = 0; b < nb; b++) {
                unsigned ch = q[2 * b];
                bad |= (ch | 1u) ^ '1';
                bad |= !is_ws((char)q[2 * b + 1]);
                word |= (uint32_t)(ch & 1u) << b;
            }
            o[w] = word;
        }
        if (bad) { j->bad = 1; return NULL; }
    }
    return NULL;
}

int nemio_read_dat(const char *path, FILE *err, int n, int d, int wpr, uint32_t **packed_out,
                   int n_threads)
{
    cursor c;
    *packed_out = NULL;
    if (map_file(path, &c) != 0) {
        fprintf(err, "File matrix %s does not exist\n", path);
        return NEMB_E_FILE;
    }
    uint32_t *out = calloc((size_t)n * wpr, sizeof(uint32_t));
    if (!out) { unmap_file(&c); return NEMB_E_MEMORY; }
    size_t want = (size_t)n * 2 * d;
    int done = 0;
    if (c.len == want || c.len + 1 == want) {
        /* the last separator may be missing: pad-check it separately */
        int last_missing = (c.len + 1 == want);
        if (n_threads < 1) n_threads = 1;
        if (n_threads > 64) n_threads = 64;
        if (n < 4096) n_threads = 1;
        pthread_t th[64]; dat_job jobs[64]; int started[64];
        int rows = last_missing ? n - 1 : n;
        for (int t = 0; t < n_threads; t++) {
            jobs[t] = (dat_job){c.base, d, wpr, (int)((long long)rows * t / n_threads),
                                (int)((long long)rows * (t + 1) / n_threads), out, 0};
            /* a thread that cannot be created is run inline, like nei_read_lines does */
            started[t] = n_threads > 1 && pthread_create(&th[t], NULL, dat_worker, &jobs[t]) == 0;
            if (!started[t]) dat_worker(&jobs[t]);
        }
        int bad = 0;
        for (int t = 0; t < n_threads; t++) {
            if (started[t]) pthread_join(th[t], NULL);
            bad |= jobs[t].bad;
        }
        done = !bad && !last_missing;
    }
    if (!done) { /* general tokenizer: any whitespace, any float spelling of 0 / 1 */
        memset(out, 0, (size_t)n * wpr * sizeof(uint32_t));
        c.p = c.base;
        const char *t; size_t len;
        for (int i = 0; i < n; i++) {
            for (int j = 0; j < d; j++) {
                float v;
                if (!next_token(&c, &t, &len)) {
                    fprintf(err, "%s : short file (%d/%d lines and %d/%d columns)\n", path, i, n, j, d);
                    free(out); unmap_file(&c);
                    return NEMB_E_FILE;
                }
                int bit;
                if (len == 1 && (t[0] == '0' || t[0] == '1')) bit = t[0] - '0';
                else if (tok_float(t, len, &v) && (v == 0.0f || v == 1.0f)) bit = (v == 1.0f);
                else {
                    fprintf(err, "%s : value '%.*s' at line %d column %d is not 0 or 1 "
                                 "(Bernoulli engine: binary data without missing values only)\n",
                            path, (int)(len > 32 ? 32 : len), t, i + 1, j + 1);
                    free(out); unmap_file(&c);
                    return NEMB_E_FILE;
                }
                if (bit) out[(size_t)i * wpr + (j >> 5)] |= 1u << (j & 31);
            }
        }
    }
    unmap_file(&c);
    *packed_out = out;
    return NEMB_OK;
}

/* ------------------------------------------------------------------ .m */
int nemio_read_m(const char *path, FILE *err, int k, int d, int *flag, float *prop, float *center,
                 float *disp)
{
    cursor c;
    if (map_file(path, &c) != 0) {
        fprintf(err, "File param %s does not exist\n", path);
        return NEMB_E_FILE;
    }
    int nbvalues = 1 + (k - 1) + 2 * k * d, rc = NEMB_OK;
    const char *t; size_t len; long fl; float v;
#define NEED_TOKEN()                                                                          \
    if (!next_token(&c, &t, &len)) {                                                          \
        fprintf(err, "The file %s needs at least %d values (%d missing)\n", path,             \
                1 + (k - 1) + 2 * k * d, nbvalues);                                           \
        unmap_file(&c);                                                                       \
        return NEMB_E_FILE;                                                                   \
    }                                                                                         \
    nbvalues--
    NEED_TOKEN();
    if (!tok_int(t, len, &fl) || (fl != 1 && fl != 2)) {
        fprintf(err, "First line of file %s must be 1 (parameters at beginning) or 2 (fixed "
                     "parameters throughout the clustering process) \n", path);
        unmap_file(&c);
        return NEMB_E_FILE;
    }
    *flag = (int)fl;
    float pk = 1.0f;
    for (int i = 0; i < k - 1; i++) {
        NEED_TOKEN();
        if (!tok_float(t, len, &v)) goto bad;
        prop[i] = v;
        pk = pk - prop[i];                      /* float, nem_exe.c:1032 */
    }
    prop[k - 1] = pk;
    if (pk <= 0.0f) { fprintf(err, "Last class has pK = %5.2f <= 0\n", pk); rc = NEMB_E_FILE; }
    for (int i = 0; i < k * d; i++) {
        NEED_TOKEN();
        if (!tok_float(t, len, &v)) goto bad;
        center[i] = v;
    }
    for (int i = 0; i < k * d; i++) {
        NEED_TOKEN();
        if (!tok_float(t, len, &v)) goto bad;
        disp[i] = v;
        if (!(v > 0)) {
            fprintf(err, "Dispersion(k=%d, d=%d) = %5.3f <= 0\n", i / d + 1, i % d + 1, v);
            rc = NEMB_E_FILE;
        }
    }
    if (next_token(&c, &t, &len)) {
        fprintf(err, "The file %s needs do not need more than %d values\n", path, 1 + (k - 1) + 2 * k * d);
        rc = NEMB_E_FILE;
    }
    unmap_file(&c);
    return rc;
bad:
    fprintf(err, "The file %s holds a token that is not a number: '%.*s'\n", path,
            (int)(len > 32 ? 32 : len), t);
    unmap_file(&c);
    return NEMB_E_FILE;
#undef NEED_TOKEN
}

/* ------------------------------------------------------------------ .nei, line-structured fast path
 * PPanGGOLiN writes one record per line (ppanggolin.py:862-888).  When every non-blank line holds
 * exactly one record -- id, nb, nb neighbour ids and, in a weighted file, nb weights -- the lines
 * are split over the host threads; each thread parses its lines into its own pool and the pools
 * are merged in file order ("the last record of a point wins", like the sequential reader).  Any
 * line that does not fit (a record spanning lines, a short line, a bad token) makes the caller
 * fall back to the sequential token reader, which has the reference's fscanf semantics and
 * reports the error. */
typedef struct { int32_t id, nv; int64_t off; } nei_rec;
typedef struct {
    const char *beg, *end; int n, weighted;
    nei_rec *rec; size_t nrec, rec_cap;
    int32_t *pc; float *pw; size_t used, cap;
    int fallback;
} nei_job;

static void *nei_worker(void *arg)
{
    nei_job *j = arg;
    const char *p = j->beg;
    while (p < j->end) {
        const char *eol = memchr(p, '\n', (size_t)(j->end - p));
        if (!eol) eol = j->end;
        cursor c = {p, eol, p, 0};                    /* p, end, base, len: one line */
        p = eol < j->end ? eol + 1 : j->end;
        const char *t; size_t len; long id, nb;
        if (!next_token(&c, &t, &len)) continue;      /* blank line */
        if (!tok_int(t, len, &id) || !next_token(&c, &t, &len) || !tok_int(t, len, &nb) ||
            id < 1 || id > j->n || nb < 0 || nb > (1 << 28) ||
            nb > (long)(c.end - c.p)) { j->fallback = 1; return NULL; }   /* nb tokens cannot fit on the line */
        if (j->used + (size_t)nb > j->cap) {
            size_t cap = j->cap ? j->cap : 4096;
            while (j->used + (size_t)nb > cap) cap *= 2;
            int32_t *pc = realloc(j->pc, cap * sizeof(int32_t));
            float *pw = pc ? realloc(j->pw, cap * sizeof(float)) : NULL;
            if (pc) j->pc = pc;
            if (pw) j->pw = pw;
            if (!pc || !pw) { j->fallback = 1; return NULL; }
            j->cap = cap;
        }
        size_t s0 = j->used;
        for (long q = 0; q < nb; q++) {
            long nbr;
            if (!next_token(&c, &t, &len) || !tok_int(t, len, &nbr)) { j->fallback = 1; return NULL; }
            j->pc[s0 + q] = nbr < INT32_MIN || nbr > INT32_MAX ? 0 : (int32_t)nbr;
            j->pw[s0 + q] = 1.0f;
        }
        if (j->weighted)
            for (long q = 0; q < nb; q++) {
                float w;
                if (!next_token(&c, &t, &len) || !tok_float(t, len, &w)) { j->fallback = 1; return NULL; }
                j->pw[s0 + q] = w;
            }
        if (next_token(&c, &t, &len)) { j->fallback = 1; return NULL; }   /* trailing tokens */
        size_t nv = 0;
        for (long q = 0; q < nb; q++) {
            int32_t nbr = j->pc[s0 + q]; float w = j->pw[s0 + q];
            if (nbr >= 1 && nbr <= j->n && w != 0.0f) { j->pc[s0 + nv] = nbr - 1; j->pw[s0 + nv] = w; nv++; }
        }
        if (j->nrec == j->rec_cap) {
            size_t cap = j->rec_cap ? 2 * j->rec_cap : 4096;
            nei_rec *r = realloc(j->rec, cap * sizeof(nei_rec));
            if (!r) { j->fallback = 1; return NULL; }
            j->rec = r; j->rec_cap = cap;
        }
        j->rec[j->nrec++] = (nei_rec){(int32_t)id, (int32_t)nv, (int64_t)s0};
        j->used = s0 + nv;
    }
    return NULL;
}

/* returns 1 when the CSR was built, 0 when the caller must use the sequential reader, < 0 = -NEMB_E_* */
static int nei_read_lines(const char *beg, const char *end, int n, int weighted, int n_threads,
                          int32_t **row_ptr_out, int32_t **col_out, float **wgt_out, int *max_neigh)
{
    if (n_threads > 64) n_threads = 64;
    if (n_threads < 1 || end - beg < (1 << 16)) n_threads = 1;
    nei_job jobs[64];
    pthread_t th[64];
    int started[64] = {0};
    const char *cut = beg;
    for (int t = 0; t < n_threads; t++) {
        const char *stop = end;
        if (t + 1 < n_threads) {
            stop = beg + (size_t)(end - beg) * (size_t)(t + 1) / (size_t)n_threads;
            if (stop < cut) stop = cut;
            const char *nl = memchr(stop, '\n', (size_t)(end - stop));
            stop = nl ? nl + 1 : end;
        }
        memset(&jobs[t], 0, sizeof jobs[t]);
        jobs[t].beg = cut; jobs[t].end = stop; jobs[t].n = n; jobs[t].weighted = weighted;
        cut = stop;
    }
    for (int t = 1; t < n_threads; t++) started[t] = pthread_create(&th[t], NULL, nei_worker, &jobs[t]) == 0;
    nei_worker(&jobs[0]);
    for (int t = 1; t < n_threads; t++) {
        if (started[t]) pthread_join(th[t], NULL);
        else nei_worker(&jobs[t]);
    }
    int ret = 1;
    for (int t = 0; t < n_threads; t++) if (jobs[t].fallback) ret = 0;
    int32_t *rp = NULL, *cl = NULL; float *wg = NULL;
    int32_t *last_t = NULL; int64_t *last_r = NULL;
    if (ret == 1) {
        last_t = malloc(sizeof(int32_t) * (size_t)n);
        last_r = malloc(sizeof(int64_t) * (size_t)n);
        rp = malloc(sizeof(int32_t) * ((size_t)n + 1));
        if (!last_t || !last_r || !rp) ret = -NEMB_E_MEMORY;
    }
    if (ret == 1) {
        for (int i = 0; i < n; i++) last_t[i] = -1;
        for (int t = 0; t < n_threads; t++)
            for (size_t r = 0; r < jobs[t].nrec; r++) {
                int id = jobs[t].rec[r].id - 1;
                last_t[id] = t; last_r[id] = (int64_t)r;
            }
        int64_t nnz = 0;
        int mx = 0;
        for (int i = 0; i < n; i++) {
            rp[i] = (int32_t)nnz;
            if (last_t[i] >= 0) {
                int nv = jobs[last_t[i]].rec[last_r[i]].nv;
                nnz += nv;
                if (nv > mx) mx = nv;
            }
            if (nnz > 2147483647LL) { ret = -NEMB_E_FILE; break; }
        }
        if (ret == 1) {
            rp[n] = (int32_t)nnz;
            cl = malloc(sizeof(int32_t) * (size_t)(nnz ? nnz : 1));
            wg = malloc(sizeof(float) * (size_t)(nnz ? nnz : 1));
            if (!cl || !wg) ret = -NEMB_E_MEMORY;
        }
        if (ret == 1) {
            for (int i = 0; i < n; i++) {
                if (last_t[i] < 0) continue;
                const nei_job *j = &jobs[last_t[i]];
                const nei_rec *r = &j->rec[last_r[i]];
                if (r->nv == 0) continue;   /* a thread without entries has no pool */
                memcpy(cl + rp[i], j->pc + r->off, sizeof(int32_t) * (size_t)r->nv);
                memcpy(wg + rp[i], j->pw + r->off, sizeof(float) * (size_t)r->nv);
            }
            *row_ptr_out = rp; *col_out = cl; *wgt_out = wg; *max_neigh = mx;
            rp = NULL; cl = NULL; wg = NULL;
        }
    }
    free(rp); free(cl); free(wg); free(last_t); free(last_r);
    for (int t = 0; t < n_threads; t++) { free(jobs[t].rec); free(jobs[t].pc); free(jobs[t].pw); }
    return ret;
}

/* ------------------------------------------------------------------ .nei -> CSR */
int nemio_read_nei(const char *base, FILE *err, int n, int32_t **row_ptr_out, int32_t **col_out,
                   float **wgt_out, int *max_neigh, char *comment, int comment_len)
{
    char path[4200];
    snprintf(path, sizeof path, "%s.nei", base);
    cursor c;
    *row_ptr_out = NULL; *col_out = NULL; *wgt_out = NULL; *max_neigh = 0;
    if (map_file(path, &c) != 0) {
        fprintf(err, "File Neigh %s File does not exist\n", path);
        return NEMB_E_FILE;
    }
    skip_comments(&c, comment, comment_len);
    const char *t; size_t len; long v;
    int weighted = 0, rc = NEMB_OK;
    if (next_token(&c, &t, &len) && tok_int(t, len, &v)) weighted = v != 0;
    {
        long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
        int fast = nei_read_lines(c.p, c.end, n, weighted, ncpu > 16 ? 16 : (int)ncpu, row_ptr_out,
                                  col_out, wgt_out, max_neigh);
        if (fast == 1) { unmap_file(&c); return NEMB_OK; }
        if (fast < 0) {
            if (fast == -NEMB_E_FILE) fprintf(err, "neighbourhood too large\n");
            unmap_file(&c);
            return -fast;
        }
    }

    /* records land in a scratch pool; start[i]/cnt[i] point at the LAST record of point i */
    size_t cap = 1 << 16, used = 0;
    int32_t *pool_c = malloc(cap * sizeof(int32_t));
    float *pool_w = malloc(cap * sizeof(float));
    int64_t *start = malloc(sizeof(int64_t) * (size_t)n);
    int32_t *cnt = calloc((size_t)n, sizeof(int32_t));
    if (!pool_c || !pool_w || !start || !cnt) { rc = NEMB_E_MEMORY; goto out; }
    for (int i = 0; i < n; i++) start[i] = 0;
    int line = 0;
    while (next_token(&c, &t, &len)) {
        long id, nb;
        if (!tok_int(t, len, &id)) break;                 /* reference: loop just ends */
        if (!next_token(&c, &t, &len) || !tok_int(t, len, &nb)) break;
        if (id < 1 || id > n || nb < 0) {
            fprintf(err, "Error in neighb. file l.%d : point id %ld out of 1..%d\n", line, id, n);
            rc = NEMB_E_FILE; goto out;
        }
        if (nb > (long)(c.end - c.p)) {   /* more neighbours announced than bytes left in the file */
            fprintf(err, "Error in neighb. file l.%d : neighbor %ld\n", line, (long)(c.end - c.p));
            rc = NEMB_E_FILE; goto out;
        }
        if (used + (size_t)nb > cap) {
            while (used + (size_t)nb > cap) cap *= 2;
            int32_t *nc = realloc(pool_c, cap * sizeof(int32_t));
            if (nc) pool_c = nc;
            float *nw = nc ? realloc(pool_w, cap * sizeof(float)) : NULL;
            if (nw) pool_w = nw;
            if (!nc || !nw) { rc = NEMB_E_MEMORY; goto out; }
        }
        size_t s0 = used;
        for (long q = 0; q < nb; q++) {
            long nbr;
            if (!next_token(&c, &t, &len) || !tok_int(t, len, &nbr)) {
                fprintf(err, "Error in neighb. file l.%d : neighbor %ld\n", line, q);
                rc = NEMB_E_FILE; goto out;
            }
            pool_c[s0 + q] = nbr < INT32_MIN || nbr > INT32_MAX ? 0 : (int32_t)nbr;   /* out of 1..n: skipped below */
            pool_w[s0 + q] = 1.0f;
        }
        if (weighted) {
            for (long q = 0; q < nb; q++) {
                float w;
                if (!next_token(&c, &t, &len) || !tok_float(t, len, &w)) {
                    fprintf(err, "Error in neighb. file l.%d : weight %ld\n", line, q);
                    rc = NEMB_E_FILE; goto out;
                }
                pool_w[s0 + q] = w;
            }
        }
        /* compact: keep entries with a valid index and a non-zero weight */
        size_t nv = 0;
        for (long q = 0; q < nb; q++) {
            int32_t nbr = pool_c[s0 + q]; float w = pool_w[s0 + q];
            if (nbr >= 1 && nbr <= n && w != 0.0f) {
                pool_c[s0 + nv] = nbr - 1; pool_w[s0 + nv] = w; nv++;
            }
        }
        start[id - 1] = (int64_t)s0; cnt[id - 1] = (int32_t)nv;
        used = s0 + nv;
        line++;
    }
    {
        int32_t *rp = malloc(sizeof(int32_t) * ((size_t)n + 1));
        int64_t nnz = 0;
        for (int i = 0; i < n; i++) nnz += cnt[i];
        if (nnz > 2147483647LL) { fprintf(err, "neighbourhood too large\n"); rc = NEMB_E_FILE; free(rp); goto out; }
        int32_t *cl = malloc(sizeof(int32_t) * (size_t)(nnz ? nnz : 1));
        float *wg = malloc(sizeof(float) * (size_t)(nnz ? nnz : 1));
        if (!rp || !cl || !wg) { rc = NEMB_E_MEMORY; free(rp); free(cl); free(wg); goto out; }
        int32_t o = 0, mx = 0;
        for (int i = 0; i < n; i++) {
            rp[i] = o;
            memcpy(cl + o, pool_c + start[i], sizeof(int32_t) * (size_t)cnt[i]);
            memcpy(wg + o, pool_w + start[i], sizeof(float) * (size_t)cnt[i]);
            o += cnt[i];
            if (cnt[i] > mx) mx = cnt[i];
        }
        rp[n] = o;
        *row_ptr_out = rp; *col_out = cl; *wgt_out = wg; *max_neigh = mx;
    }
out:
    free(pool_c); free(pool_w); free(start); free(cnt);
    unmap_file(&c);
    return rc;
}

/* ------------------------------------------------------------------ writers */
/* " %5.3f " of a value in [0,1], rounded like printf (exact product, ties to even) */
static inline char *put_uf(char *p, float v)
{
    if (!(v >= 0.0f && v <= 1.0f)) return p + sprintf(p, " %5.3f ", (double)v);
    double y = (double)v * 1000.0;            /* exact: 24-bit x 10-bit mantissas */
    double fl = floor(y), fr = y - fl;
    long q = (long)fl;
    if (fr > 0.5 || (fr == 0.5 && (q & 1))) q++;
    p[0] = ' '; p[1] = (char)('0' + q / 1000); p[2] = '.';
    p[3] = (char)('0' + (q / 100) % 10); p[4] = (char)('0' + (q / 10) % 10);
    p[5] = (char)('0' + q % 10); p[6] = ' ';
    return p + 7;
}

int nemio_write_uf(const char *path, FILE *err, int n, int k, const float *t)
{
    FILE *f = fopen(path, "w");
    if (!f) { fprintf(err, "Could not open file '%s' in write mode\n", path); return NEMB_E_FILE; }
    size_t per = (size_t)k * 32 + 2, rows = 4096;
    char *buf = malloc(per * rows);
    if (!buf) { fclose(f); return NEMB_E_MEMORY; }
    for (int i0 = 0; i0 < n; i0 += (int)rows) {
        char *p = buf;
        int i1 = i0 + (int)rows < n ? i0 + (int)rows : n;
        for (int i = i0; i < i1; i++) {
            for (int c = 0; c < k; c++) p = put_uf(p, t[(size_t)i * k + c]);
            *p++ = '\n';
        }
        fwrite(buf, 1, (size_t)(p - buf), f);
    }
    free(buf);
    fclose(f);
    return NEMB_OK;
}

int nemio_write_cf(const char *path, FILE *err, int n, const int32_t *label)
{
    FILE *f = fopen(path, "w");
    if (!f) { fprintf(err, "Could not open file '%s' in write mode\n", path); return NEMB_E_FILE; }
    for (int i = 0; i < n; i++) fprintf(f, "%d ", label[i] + 1);   /* nem_exe.c:1652 */
    fprintf(f, "\n");
    fclose(f);
    return NEMB_OK;
}

int nemio_write_mf(const char *path, FILE *err, int k, int d, const double crit[4], float beta,
                   const float *prop, const float *center, const float *disp)
{
    return nemio_write_mf_mode(path, err, k, d, crit, beta, 0, prop, center, disp);
}

int nemio_write_mf_mode(const char *path, FILE *err, int k, int d, const double crit[4], float beta,
                        int beta_mode, const float *prop, const float *center, const float *disp)
{
    /* BetaDesVC, nem_typ.h:540-543 */
    static const char *const beta_des[4] = {"fixed", "pseudo-likelihood gradient",
                                            "heuristic Hathaway crit", "heuristic mixture likelihood"};
    FILE *f = fopen(path, "w");
    if (!f) { fprintf(err, "Could not open file '%s' in write mode\n", path); return NEMB_E_FILE; }
    /* nem_exe.c:1708-1773; criteria are float in the reference (CriterT), error rate is NaN */
    fprintf(f, "Criteria U=NEM, D=Hathaway, L=mixture, M=markov ps-like, error\n\n");
    fprintf(f, "  %g    %g    %g    %g   %g\n\n", (double)(float)crit[0], (double)(float)crit[1],
            (double)(float)crit[2], (double)(float)crit[3], (double)NAN);
    fprintf(f, "Beta (%s)\n", beta_des[beta_mode >= 0 && beta_mode < 4 ? beta_mode : 0]);
    fprintf(f, "  %6.4f\n", (double)beta);
    fprintf(f, "Mu (%d), Pk, and disp (%d) of the %d classes\n\n", d, d, k);
    for (int c = 0; c < k; c++) {
        for (int j = 0; j < d; j++) fprintf(f, " %10.3g ", (double)center[(size_t)c * d + j]);
        fprintf(f, "  %5.3g  ", (double)prop[c]);
        for (int j = 0; j < d; j++) fprintf(f, " %10g ", (double)disp[(size_t)c * d + j]);
        fprintf(f, "\n");
    }
    fclose(f);
    return NEMB_OK;
}
