// nem_kernels.cu -- hand-written CUDA kernels (sm_100a) of the NEM partitioning hot path and the
// thin extern "C" launch layer declared in nem_device.h.
//
// Data layout in HBM (DESIGN.md "Layout"):
//   X    uint32 [N][wpr]   genome d of family i = bit d%32 of word d/32; wpr multiple of 4 (uint4)
//   XT   uint32 [D][nwt]   transposed bits: family i = bit i%32 of word i/32 of column d
//   CSR  int32 row_ptr[N+1], int32 col[nnz], float wgt[nnz]  (file order inside a row)
//   logpf double [N][K]    log p_k + log f_k(x_i)   (-inf = zero density)
//   lab  uint8 [N]         ncem hard labels, 255 = unlabelled;   t float [N][K] for nem
//
// Reference functions restated by each kernel are cited at the kernel (reference root
// /root/reference/ppanggolin/NEM).  All kernels are HBM/latency bound integer/bit work: no
// tensor cores by design (K=3 gives ~3 popc per 4 bytes of X).
#include "nem_device.h"

// compile-time tuning knobs (defaults = the measured configuration)
#ifndef FX_CLUSTER
#define FX_CLUSTER 8   // CTAs of the fix-up tail's cluster (> 8 needs the non-portable opt-in)
#endif
#ifndef JAC_MINB
#define JAC_MINB 1     // min resident CTAs per SM asked of the dense sweep round (register cap)
#endif
#ifndef JAC_SPT
#define JAC_SPT 4      // sub-tiles of 256 sites per CTA of the dense sweep round (sites per thread)
#endif
#ifndef JAC_MINB_SMALLK
#define JAC_MINB_SMALLK 6   // K <= 4 (the pangenome case): 40 registers, six CTAs per SM
#endif
#ifndef FIXUP_HOIST
#define FIXUP_HOIST 0  // fix-up rounds: label-independent loads issued before the dirty-flag fence (measured neutral on C4)
#endif
#ifndef FIXUP_ACQ
#define FIXUP_ACQ 0    // fix-up rounds: flag clear as an ACQUIRE exchange instead of exchange + fence (untested knob)
#endif
#ifndef JAC_HUB_GUARD
#define JAC_HUB_GUARD 1   // dense round: light threads never copy a hub's label (see the kernel).  Default since the
                          // end of round 1 (the race it closes is real: tests/test_sweep_protocol_model.py); the
                          // GPU-measured numbers under profiles/r1_* are of the =0 build, same registers/spills
#endif
#define JAC_TILE (256 * JAC_SPT)
#ifndef JAC_SPARSE_MAX
#define JAC_SPARSE_MAX 256   // at most this many sites left in a CTA: compacted path
#endif

#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <math_constants.h>
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NEM_EPSILON 1e-20  // nem_typ.h:65
#define FULL 0xffffffffu

static __device__ __forceinline__ double neg_inf() { return -CUDART_INF; }
static __device__ __forceinline__ void prefetch_l1(const void *p) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// log p_k f_k(x_i) of the popcount path from the Hamming count (ComputePkFkiM nem_alg.c:2260-2285 +
// DensBernoulli nem_mod.c:649-674).  ONE expression shared by the density epilogues, the cached
// rebuild and the consumers that evaluate it in registers (nemk_lpsrc), so the bits are the same
// wherever it is computed.
static __device__ __forceinline__ double logpf_of_h(const nemk_coef *__restrict__ coef, int k, int h) {
    double lp = coef->lp[k];
    if (coef->forb[k]) return h ? neg_inf() : lp;
    return lp - (coef->a[k] * (double)h + coef->base[k]);
}
template <int KT>
static __device__ __forceinline__ void load_lp(const nemk_lpsrc &src, int K, size_t il, double (&lp)[KT]) {
#pragma unroll
    for (int k = 0; k < KT; k++) {
        if (k < K) lp[k] = src.ham ? logpf_of_h(src.coef, k, src.ham[il * K + k]) : src.logpf[il * K + k];
        else lp[k] = neg_inf();
    }
}

// ---------------------------------------------------------------------------------------------
// error bookkeeping for the launch layer
static cudaError_t g_last_err = cudaSuccess;
static void note_launch() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess && g_last_err == cudaSuccess) g_last_err = e;
}
extern "C" int nemk_last_error(char *buf, int len) {
    cudaError_t e = g_last_err;
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) return 0;
    if (buf && len > 0) snprintf(buf, len, "%s", cudaGetErrorString(e));
    g_last_err = cudaSuccess;
    return (int)e;
}
static inline cudaStream_t S(nemk_stream s) { return (cudaStream_t)s; }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// deterministic block-wide double sum (fixed shuffle tree + fixed smem order)
template <int THREADS>
static __device__ __forceinline__ double block_sum(double v, double *sh /*[32]*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = (l < THREADS / 32) ? sh[l] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(FULL, r, o);
    }
    if (threadIdx.x == 0) sh[0] = r;
    __syncthreads();
    r = sh[0];
    return r;
}

// =============================================================================================
// Loader kernels (the reference keeps X as float[N*D], nem_exe.c:834-898)
// =============================================================================================
// one warp per family: lane reads byte 32w+lane (coalesced), ballot -> one packed word
__global__ void k_pack_u8(const uint8_t *__restrict__ x, int n, int d, int wpr,
                          uint32_t *__restrict__ out) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const uint8_t *row = x + (size_t)warp * d;
    for (int w = 0; w < wpr; w++) {
        int j = w * 32 + lane;
        unsigned bit = (j < d) ? (row[j] != 0) : 0u;
        unsigned word = __ballot_sync(FULL, bit);
        if (lane == 0) out[(size_t)warp * wpr + w] = word;
    }
}

// Bit-matrix transpose X -> XT.  A CTA owns 256 families x 32 words (1024 genomes), staged in
// shared memory with coalesced 128-byte row reads.  Each thread then transposes 32x32 bit tiles IN
// REGISTERS (5 butterfly stages of masked swaps, ~200 integer ops per tile instead of 32 warp
// ballots) and every genome's 8 output words (256 families = one 32-byte sector) leave as two
// 16-byte stores.
static __device__ __forceinline__ void transpose32(uint32_t (&a)[32]) {
    // a[r] bit c  ->  a[c] bit r   (Hacker's Delight 7-3, LSB-first variant)
    uint32_t m = 0x0000ffffu;
#pragma unroll
    for (int j = 16; j != 0; j >>= 1, m ^= (m << j)) {
#pragma unroll
        for (int k = 0; k < 32; k = (k + j + 1) & ~j) {
            uint32_t t = ((a[k] >> j) ^ a[k + j]) & m;
            a[k] ^= t << j;
            a[k + j] ^= t;
        }
    }
}

#define TB_ROWS 256
__global__ void __launch_bounds__(256)
k_transpose_bits(const uint32_t *__restrict__ x, int row_base, int n, int wpr, int d, int nwt,
                 uint32_t *__restrict__ xt) {
    __shared__ uint32_t smem_t[1024 * 9];   // >= TB_ROWS*33; reused as out[1024][9]
    uint32_t (*tile)[33] = reinterpret_cast<uint32_t (*)[33]>(smem_t);
    const int r0 = row_base + blockIdx.x * TB_ROWS, w0 = blockIdx.y * 32;   // row_base: multiple of TB_ROWS
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;  // 8 warps
    for (int rr = wid; rr < TB_ROWS; rr += 8) {
        int r = r0 + rr, w = w0 + lane;
        tile[rr][lane] = (r < n && w < wpr) ? x[(size_t)r * wpr + w] : 0u;
    }
    __syncthreads();
    // thread (g = wid, c = lane): the 32x32 tile of row group g (families r0+32g..+31), word
    // column c.  tile[32g + r][c]: the 32 lanes of a warp read 32 consecutive columns of one row
    uint32_t a[32];
#pragma unroll
    for (int r = 0; r < 32; r++) a[r] = tile[wid * 32 + r][lane];
    transpose32(a);
    __syncthreads();
    // a[b] = families 32g..32g+31 of genome (w0+lane)*32 + b.  Re-stage so that one thread owns
    // one genome's 8 words: out[slot(genome)][g]
    uint32_t (*outw)[9] = reinterpret_cast<uint32_t (*)[9]>(smem_t);
#pragma unroll
    for (int b = 0; b < 32; b++) outw[b * 32 + lane][wid] = a[b];   // slot b*32+c: conflict-free
    __syncthreads();
    const int g8 = (r0 >> 5);        // first output word of this CTA's 256 families
    for (int q = threadIdx.x; q < 1024; q += 256) {
        int dd = (w0 + (q & 31)) * 32 + (q >> 5);   // slot q = b*32 + c  ->  genome (w0+c)*32 + b
        if (dd >= d) continue;
        uint32_t *dst = xt + (size_t)dd * nwt + g8;
        if (g8 + 8 <= nwt) {
            *reinterpret_cast<uint4 *>(dst) = make_uint4(outw[q][0], outw[q][1], outw[q][2], outw[q][3]);
            *reinterpret_cast<uint4 *>(dst + 4) = make_uint4(outw[q][4], outw[q][5], outw[q][6], outw[q][7]);
        } else {
            for (int g = 0; g < 8 && g8 + g < nwt; g++) dst[g] = outw[q][g];
        }
    }
}

// =============================================================================================
// theta -> per-class tables.  DensBernoulli's per-variable term (nem_mod.c:656-670):
//   absdif = abs((int)(x - mu));  disp > EPSILON: absdif*log((1-disp)/disp) - log(1-disp)
//   else absdif != 0 -> zero density.   (1-disp)/disp and 1-disp are FLOAT expressions there.
// one CTA per class, variables in parallel.
// =============================================================================================
#define TT_THREADS 1024
// Per-thread part of the tables of class k over the genomes [32*w_lo, 32*w_hi): writes the bit
// masks and delta, returns the thread's partial sums.  The two log() of a variable are only
// evaluated when its dispersion differs from the class's first one (never, for the
// per-class-constant dispersion models).
struct TablesPartial { double base_u, base_g; int notok, mu_moved, n_valid, n_x1; };
struct ClassCoef { float e0; bool live0; double a0, c0; };

static __device__ __forceinline__ ClassCoef class_coef(float e0) {
    ClassCoef cc;
    cc.e0 = e0;
    cc.live0 = (double)e0 > NEM_EPSILON;
    cc.a0 = 0.0; cc.c0 = 0.0;
    if (cc.live0) {
        float ratio = __fdiv_rn(__fsub_rn(1.0f, e0), e0);
        float om = __fsub_rn(1.0f, e0);
        cc.a0 = log((double)ratio);
        cc.c0 = -log((double)om);
    }
    return cc;
}

static __device__ __forceinline__ TablesPartial tables_words(
    int k, int D, int wpr, int w_lo, int w_hi, const ClassCoef &cc, const float *center,
    const float *disp, uint32_t *mask_xor, uint32_t *mask_valid, uint32_t *mask_f0,
    uint32_t *mask_f1, double *delta) {
    TablesPartial p = {0.0, 0.0, 0, 0, 0, 0};
    const int tid = threadIdx.x, lane = tid & 31;
    for (int j0 = w_lo * 32; j0 < w_hi * 32; j0 += blockDim.x) {
        int j = j0 + tid;  // whole warps stay together: the range starts and ends on a word
        bool in = j < D && j < w_hi * 32;
        float mu = in ? center[(size_t)k * D + j] : 0.5f;
        float e = in ? disp[(size_t)k * D + j] : cc.e0;
        int m0 = abs((int)(0.0f - mu)), m1 = abs((int)(1.0f - mu));
        bool same = __float_as_uint(e) == __float_as_uint(cc.e0);
        bool live = same ? cc.live0 : ((double)e > NEM_EPSILON);
        double a = same ? cc.a0 : 0.0, c = same ? cc.c0 : 0.0;
        if (in && live && !same) {
            float ratio = __fdiv_rn(__fsub_rn(1.0f, e), e);
            float om = __fsub_rn(1.0f, e);
            a = log((double)ratio);
            c = -log((double)om);
        }
        double cost0 = in ? (m0 * a + c) : 0.0, cost1 = in ? (m1 * a + c) : 0.0;
        if (in) {
            delta[(size_t)k * D + j] = cost1 - cost0;
            p.base_g += cost0;
            p.base_u += c;
            if (!same || m0 > 1 || m1 > 1) p.notok = 1;
        }
        unsigned bx = __ballot_sync(FULL, in && m0 == 1 && m1 == 0);
        unsigned bv = __ballot_sync(FULL, in && m0 != m1);
        unsigned b0 = __ballot_sync(FULL, in && !live && m0 != 0);
        unsigned b1 = __ballot_sync(FULL, in && !live && m1 != 0);
        if (lane == 0 && (j >> 5) < w_hi) {
            size_t o = (size_t)k * wpr + (j >> 5);
            p.n_valid += __popc(bv); p.n_x1 += __popc(bx);
            if (mask_xor[o] != bx || mask_valid[o] != bv) p.mu_moved = 1;
            mask_xor[o] = bx; mask_valid[o] = bv; mask_f0[o] = b0; mask_f1[o] = b1;
        }
    }
    return p;
}

static __device__ __forceinline__ void tables_commit(int k, int K, int D, const float *prop,
                                                     nemk_coef *coef, double *delta,
                                                     const ClassCoef &cc, double base_u,
                                                     double base_g, bool ok, int n_valid, int n_x1,
                                                     bool have_prev) {
    // previous coefficients of the class (margin cache: how far can its score have moved?)
    const double lp_o = coef->lp[k], a_o = coef->a[k], base_o = coef->base[k];
    const int forb_o = coef->forb[k];
    // popcount-path class kind: 3 = no genome counts (all centres 1/2), 1 = centre 0 everywhere
    // (H = popc(x)), 2 = centre 1 everywhere (H = D - popc(x)), 0 = general
    coef->kind[k] = n_valid == 0 ? 3 : (n_valid == D && n_x1 == 0) ? 1 : (n_valid == D && n_x1 == D) ? 2 : 0;
    double pk = prop[k];
    coef->lp[k] = (pk > NEM_EPSILON) ? log(pk) : neg_inf();  // nem_alg.c:2265-2271
    if (ok) {
        coef->a[k] = cc.live0 ? cc.a0 : 0.0;
        coef->base[k] = cc.live0 ? base_u : 0.0;
        coef->forb[k] = cc.live0 ? 0 : 1;
    } else {
        coef->a[k] = 0.0; coef->base[k] = base_g; coef->forb[k] = 0;
    }
    delta[(size_t)K * D + k] = base_g;  // general-path base: sum_d cost0_kd
    if (!ok) atomicAnd(&coef->uniform_ok, 0);   // preset to non-zero by the launcher
    // |score_new - score_old| <= |d(lp - base)| + |d a| * D for every H in [0, D]; a forbidding
    // class scores lp or -inf.  Anything else (first tables, general path, flag flips, infinities)
    // = unknown.
    double step = CUDART_INF;
    if (have_prev && ok && forb_o == coef->forb[k]) {
        double d0 = (coef->lp[k] - coef->base[k]) - (lp_o - base_o), da = coef->a[k] - a_o;
        if (coef->forb[k]) { d0 = coef->lp[k] - lp_o; da = 0.0; }
        double b = fabs(d0) + fabs(da) * (double)D;
        if (b == b && b < CUDART_INF) step = b;   // finite, not NaN
    }
    coef->dstep[k] = step;
}

// tables of the theta the caller supplied (start of a fit): one CTA per class
__global__ void __launch_bounds__(TT_THREADS)
k_theta_tables(int K, int D, int wpr, const float *__restrict__ prop,
               const float *__restrict__ center, const float *__restrict__ disp, nemk_coef *coef,
               uint32_t *mask_xor, uint32_t *mask_valid, uint32_t *mask_f0, uint32_t *mask_f1,
               double *delta) {
    __shared__ double sh[32];
    const int k = blockIdx.x;
    ClassCoef cc = class_coef(disp[(size_t)k * D]);
    TablesPartial p = tables_words(k, D, wpr, 0, wpr, cc, center, disp, mask_xor, mask_valid,
                                   mask_f0, mask_f1, delta);
    if (p.mu_moved) atomicOr(&coef->mu_changed, 1);   // preset by the launcher (0, or 1 = forced)
    double base_u = block_sum<TT_THREADS>(p.base_u, sh);
    double base_g = block_sum<TT_THREADS>(p.base_g, sh);
    double notok = block_sum<TT_THREADS>((double)p.notok, sh);
    double nv = block_sum<TT_THREADS>((double)p.n_valid, sh);
    double nx = block_sum<TT_THREADS>((double)p.n_x1, sh);
    if (threadIdx.x == 0)
        tables_commit(k, K, D, prop, coef, delta, cc, base_u, base_g, notok == 0.0, (int)nv, (int)nx, false);
}

// =============================================================================================
// E-step density, popcount path.  ComputePkFkiM (nem_alg.c:2260-2285) + DensBernoulli
// (nem_mod.c:619-690) for classes whose eps is constant over the genomes (sk_, s__ and
// PPanGGOLiN's default .m):  H_ik = popc((x_i ^ M1_k) & V_k),  logpf = lp - (a*H + base).
// LPR lanes cooperate on one family; uint4 loads; masks staged in shared memory.
// =============================================================================================
template <int KT, int LPR>
__global__ void __launch_bounds__(256)
k_density_uniform(int K, const uint4 *__restrict__ x, int n, int wpr4,
                  const nemk_coef *__restrict__ coef, const uint4 *__restrict__ mxor,
                  const uint4 *__restrict__ mval, double *__restrict__ logpf,
                  int32_t *__restrict__ hamming, int cached) {
    if (coef->empty_class | coef->halt) return;  // M-step found an empty class (E-step not run) / fit over
    if (cached && !coef->mu_changed) return;
    extern __shared__ uint4 smem[];
    uint4 *sx = smem, *sv = smem + (size_t)KT * wpr4;
    for (int i = threadIdx.x; i < K * wpr4; i += blockDim.x) { sx[i] = mxor[i]; sv[i] = mval[i]; }
    __syncthreads();
    const int rows_per_block = blockDim.x / LPR;
    const int sub = threadIdx.x % LPR;
    for (long long row = (long long)blockIdx.x * rows_per_block + threadIdx.x / LPR; row < n;
         row += (long long)gridDim.x * rows_per_block) {
        int h[KT];
#pragma unroll
        for (int k = 0; k < KT; k++) h[k] = 0;
        const uint4 *xr = x + (size_t)row * wpr4;
        for (int c = sub; c < wpr4; c += LPR) {
            uint4 v = __ldg(xr + c);
#pragma unroll
            for (int k = 0; k < KT; k++) {
                if (k < K) {
                    uint4 a = sx[k * wpr4 + c], b = sv[k * wpr4 + c];
                    h[k] += __popc((v.x ^ a.x) & b.x) + __popc((v.y ^ a.y) & b.y) +
                            __popc((v.z ^ a.z) & b.z) + __popc((v.w ^ a.w) & b.w);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < KT; k++)
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) h[k] += __shfl_xor_sync(FULL, h[k], o);
        if (sub == 0) {
#pragma unroll
            for (int k = 0; k < KT; k++) {
                if (k < K) {
                    if (logpf) logpf[(size_t)row * K + k] = logpf_of_h(coef, k, h[k]);
                    if (hamming) hamming[(size_t)row * K + k] = h[k];
                }
            }
        }
    }
}

// =============================================================================================
// E-step density, popcount path, v2: TMA-staged row tiles + one thread per family.
//
// v1 above spends its time on shared-memory mask reads (6 LDS.128 per 16 B of X) and on the
// quarter-rate POPC (XU) pipe (ncu: profiles/r1_c4_v1_ncu_summary.csv).  Here
//   - a tile of ROWS families is copied to shared memory by cp.async.bulk (TMA, one bulk copy
//     per family, 16-byte granules, completion on an mbarrier), double buffered, so the HBM
//     stream never waits on the ALUs;
//   - the shared row stride is an ODD number of uint4, which makes "lane r reads chunk c of
//     row r" a conflict-free LDS.128;
//   - every lane is at the same chunk at the same time, so the class masks are warp-uniform
//     (broadcast LDS.128, one wavefront);
//   - popcounts go through carry-save adders: per class  ones' = ones^a^b, carry = maj(ones,a,b)
//     (2 LOP3 on the full-rate ALU pipe) and ONE popc(carry) per two words instead of two.
// H_ik = 2*sum popc(carry) + popc(ones).  Same arithmetic result as v1 (integers).
// =============================================================================================
static __device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
static __device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
static __device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
static __device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
static __device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes,
                                                uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

static __device__ __forceinline__ uint32_t lop_xor3(uint32_t a, uint32_t b, uint32_t c) {
    return a ^ b ^ c;
}
static __device__ __forceinline__ uint32_t lop_maj3(uint32_t a, uint32_t b, uint32_t c) {
    return (a & b) | (c & (a | b));
}

// ROWS families per tile, T column groups per family: thread (r, t) owns the chunks
// [t*wpr4/T, (t+1)*wpr4/T) of family r, so that t (hence the mask address) is warp-uniform and
// lane r reads chunk c of row r (odd stride => conflict-free).  Partial counts meet in shared
// memory through integer atomics; ONE block barrier per tile; the epilogue of tile i (log-density
// + store, one OUTPUT per thread: contiguous 8-byte stores) runs after that barrier, overlapped
// with the next tile's counting.
//
// Class kinds (coef->kind, set with the tables): a class whose centre is 0 for every genome has
// H = P = popc(x), one whose centre is 1 everywhere has H = D - P, one whose centres are all 1/2
// has H = 0; only "general" classes (kind 0) pay the per-word mask work.  P is one mask-free
// carry-save stream shared by all constant classes.  In a pangenome the persistent class is all
// ones and the cloud class all zeros, so typically ONE class of three is general.
template <int KT, int ROWS, int T>
__global__ void __launch_bounds__(ROWS *T)
k_density_tma(int K, int D, const uint4 *__restrict__ x, int n, int wpr4, int stride4, int n_tiles,
              int n_stages, const nemk_coef *__restrict__ coef, const uint4 *__restrict__ mxor,
              const uint4 *__restrict__ mval, double *__restrict__ logpf,
              int32_t *__restrict__ hamming, int cached) {
    if (coef->empty_class | coef->halt) return;
    if (cached && !coef->mu_changed) return;   // H cache still valid: k_logpf_from_h does the work
    extern __shared__ __align__(128) uint4 dsm[];
    __shared__ __align__(8) uint64_t bars[16];
    __shared__ int hsum[3][ROWS * (KT + 1)];   // per row: KG general counts then P; 3 rotating slots
    __shared__ int s_gidx[KT], s_kind[KT], s_slot[KT];
    __shared__ int s_kg, s_needp;
    const int tid = threadIdx.x;
    const int r = tid % ROWS, t = tid / ROWS;
    if (tid == 0) {
        int kg = 0, needp = 0;
        for (int k = 0; k < K; k++) {
            int kind = coef->kind[k];
            s_kind[k] = kind; s_slot[k] = -1;
            if (kind == 0) { s_gidx[kg] = k; s_slot[k] = kg; kg++; }
            else if (kind != 3) needp = 1;
        }
        s_kg = kg; s_needp = needp;
    }
    for (int i = tid; i < 3 * ROWS * (KT + 1); i += ROWS * T) (&hsum[0][0])[i] = 0;
    if (tid == 0) {
        for (int q = 0; q < n_stages; q++) mbar_init(&bars[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int KG = s_kg;
    const bool needp = s_needp != 0;
    uint4 *smask = dsm;                                // [wpr4][KG][2] (xor, valid), general classes
    uint4 *tiles = dsm + (size_t)2 * KT * wpr4;        // n_stages x ROWS x stride4 ring
    for (int i = tid; i < KG * wpr4; i += ROWS * T) {
        int c = i / KG, g = i % KG, k = s_gidx[g];
        smask[2 * i] = mxor[k * wpr4 + c];
        smask[2 * i + 1] = mval[k * wpr4 + c];
    }
    __syncthreads();
    const uint32_t row_bytes = (uint32_t)wpr4 * 16u;
    const int c_lo = (int)((long long)wpr4 * t / T), c_hi = (int)((long long)wpr4 * (t + 1) / T);
    auto issue = [&](int tile, int stage) {
        long long r0 = (long long)tile * ROWS;
        int rows = (int)min((long long)ROWS, (long long)n - r0);
        if (tid == 0) mbar_expect_tx(&bars[stage], (uint32_t)rows * row_bytes);
        // every warp issues its share of the tile's bulk copies (32/T rows per warp): one warp
        // issuing them all arrives late at the tile barrier (profiles/r1_c4_density_stalls.txt).
        // A copy may complete before thread 0's expect_tx: the transaction count goes negative,
        // the phase still cannot complete before the arrival.
        constexpr int RPW = 32 / T;
        const int lane = tid & 31, rr = (tid >> 5) * RPW + lane;
        if (lane < RPW && rr < rows)
            bulk_g2s(tiles + ((size_t)stage * ROWS + rr) * stride4, x + (size_t)(r0 + rr) * wpr4,
                     row_bytes, &bars[stage]);
    };
    auto epilogue = [&](int tile, int slot) {   // one output (row, class) per thread
        for (int o = tid; o < ROWS * K; o += ROWS * T) {
            int rr = o / K, k = o - rr * K;
            long long row = (long long)tile * ROWS + rr;
            if (row >= n) break;
            const int *hs = &hsum[slot][rr * (KT + 1)];
            int kind = s_kind[k], P = hs[KT];
            int h = kind == 0 ? hs[s_slot[k]] : kind == 1 ? P : kind == 2 ? D - P : 0;
            if (logpf) logpf[(size_t)row * K + k] = logpf_of_h(coef, k, h);
            if (hamming) hamming[(size_t)row * K + k] = h;
        }
    };
    // ring of n_stages tiles: n_stages-1 bulk loads stay in flight behind the tile being counted
    int tile = blockIdx.x, it = 0, prev_tile = -1;
    for (int q = 0; q < n_stages - 1; q++) {
        long long tq = (long long)tile + (long long)q * gridDim.x;
        if (tq < n_tiles) issue((int)tq, q);
    }
    for (; tile < n_tiles; tile += gridDim.x, it++) {
        int stage = it % n_stages;
        long long next = (long long)tile + (long long)(n_stages - 1) * gridDim.x;
        if (next < n_tiles) issue((int)next, (it + n_stages - 1) % n_stages);
        // slots rotate over 3: tile `it` accumulates into it%3, the epilogue of tile it-1 reads
        // (it-1)%3, and (it+1)%3 -- last read one barrier ago -- is cleared for the next tile
        if (T > 1)
            for (int i = tid; i < ROWS * (KT + 1); i += ROWS * T) hsum[(it + 1) % 3][i] = 0;
        if (prev_tile >= 0) epilogue(prev_tile, (it + 2) % 3);
        mbar_wait(&bars[stage], (uint32_t)((it / n_stages) & 1));
        long long row = (long long)tile * ROWS + r;
        uint32_t ones[KT], cnt2[KT], onesP = 0u, cntP = 0u;
#pragma unroll
        for (int k = 0; k < KT; k++) { ones[k] = 0u; cnt2[k] = 0u; }
        if (row < n) {
            const uint4 *xr = tiles + ((size_t)stage * ROWS + r) * stride4;
#pragma unroll 2
            for (int c = c_lo; c < c_hi; c++) {
                uint4 v = xr[c];
                if (needp) {
                    uint32_t c0 = lop_maj3(onesP, v.x, v.y);
                    uint32_t o1 = lop_xor3(onesP, v.x, v.y);
                    uint32_t c1 = lop_maj3(o1, v.z, v.w);
                    onesP = lop_xor3(o1, v.z, v.w);
                    cntP += __popc(c0) + __popc(c1);
                }
                const uint4 *mk = smask + (size_t)2 * KG * c;
#pragma unroll
                for (int g = 0; g < KT; g++) {
                    if (g < KG) {
                        uint4 a = mk[2 * g], b = mk[2 * g + 1];
                        uint32_t m0 = (v.x ^ a.x) & b.x, m1 = (v.y ^ a.y) & b.y;
                        uint32_t m2 = (v.z ^ a.z) & b.z, m3 = (v.w ^ a.w) & b.w;
                        uint32_t c0 = lop_maj3(ones[g], m0, m1);
                        uint32_t o1 = lop_xor3(ones[g], m0, m1);
                        uint32_t c1 = lop_maj3(o1, m2, m3);
                        ones[g] = lop_xor3(o1, m2, m3);
                        cnt2[g] += __popc(c0) + __popc(c1);
                    }
                }
            }
        }
        int *hs = &hsum[it % 3][r * (KT + 1)];
#pragma unroll
        for (int g = 0; g < KT; g++) {
            if (g < KG) {
                int h = (int)(2u * cnt2[g]) + __popc(ones[g]);
                if (T > 1) { if (h) atomicAdd(&hs[g], h); }
                else hs[g] = h;
            }
        }
        if (needp) {
            int h = (int)(2u * cntP) + __popc(onesP);
            if (T > 1) { if (h) atomicAdd(&hs[KT], h); }
            else hs[KT] = h;
        }
        prev_tile = tile;
        __syncthreads();  // the stage is free again and hsum[it%3] is complete
    }
    if (prev_tile >= 0) epilogue(prev_tile, (it + 2) % 3);
}

// =============================================================================================
// Row popcounts P_i = popc(x_i): theta-independent, taken once per load (behind the upload).  When
// EVERY class has a constant centre -- PPanGGOLiN's initial parameters: 1 / 1/2 / 0 -- the Hamming
// counts are H = P, D - P or 0 and the density pass does not read X at all.
// =============================================================================================
__global__ void __launch_bounds__(256)
k_row_popcount(const uint4 *__restrict__ x, int n, int wpr4, int32_t *__restrict__ pop) {
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= n) return;
    int c = 0;
    for (int q = lane; q < wpr4; q += 32) {
        uint4 v = __ldg(x + (size_t)row * wpr4 + q);
        c += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
    }
    c = __reduce_add_sync(FULL, c);
    if (lane == 0) pop[row] = c;
}

template <int KT>
__global__ void __launch_bounds__(256)
k_ham_from_pop(int K, int n, int D, const nemk_coef *__restrict__ coef,
               const int32_t *__restrict__ pop, int32_t *__restrict__ ham) {
    if (coef->empty_class | coef->halt) return;
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    int P = pop[row];
#pragma unroll
    for (int k = 0; k < KT; k++) {
        if (k < K) {
            int kind = coef->kind[k];   // 1: centre 0 everywhere, 2: centre 1 everywhere, 3: all 1/2
            ham[(size_t)row * K + k] = kind == 1 ? P : kind == 2 ? D - P : 0;
        }
    }
}

// =============================================================================================
// E-step density, general path (skd / s_d / arbitrary .m): per-genome weights.
//   logf = -(base_k + sum_{d: x_id=1} delta_kd), zero density if a forbidden cell mismatches.
// One warp per family, lanes over words, set bits walked with ffs; fp64, fixed order.
// =============================================================================================
template <int KT>
__global__ void __launch_bounds__(256)
k_density_general(int K, const uint32_t *__restrict__ x, int n, int D, int wpr,
                  const nemk_coef *__restrict__ coef, const uint32_t *__restrict__ mxor,
                  const uint32_t *__restrict__ f0, const uint32_t *__restrict__ f1,
                  const double *__restrict__ delta, const double *__restrict__ base_g,
                  double *__restrict__ logpf) {
    if (coef->empty_class | coef->halt) return;
    int lane = threadIdx.x & 31;
    long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    int wreal = (D + 31) >> 5;
    for (long long row = warp; row < n; row += nwarps) {
        double acc[KT];
        unsigned nul[KT];
#pragma unroll
        for (int k = 0; k < KT; k++) { acc[k] = 0.0; nul[k] = 0u; }
        for (int w = lane; w < wreal; w += 32) {
            uint32_t v = x[(size_t)row * wpr + w];
            uint32_t live = (w == wreal - 1 && (D & 31)) ? ((1u << (D & 31)) - 1u) : FULL;
#pragma unroll
            for (int k = 0; k < KT; k++)
                if (k < K) nul[k] |= (v & f1[k * wpr + w]) | (~v & live & f0[k * wpr + w]);
            uint32_t bits = v & live;
            while (bits) {
                int b = __ffs(bits) - 1;
                bits &= bits - 1;
                int j = w * 32 + b;
#pragma unroll
                for (int k = 0; k < KT; k++)
                    if (k < K) acc[k] += delta[(size_t)k * D + j];
            }
        }
#pragma unroll
        for (int k = 0; k < KT; k++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                acc[k] += __shfl_xor_sync(FULL, acc[k], o);
                nul[k] |= __shfl_xor_sync(FULL, nul[k], o);
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < KT; k++)
                if (k < K)
                    logpf[(size_t)row * K + k] =
                        nul[k] ? neg_inf() : coef->lp[k] - (base_g[k] + acc[k]);
        }
    }
}

// =============================================================================================
// E-step density, general path, tiled (the default for skd / s_d / arbitrary .m when the delta
// table fits shared memory).  DensBernoulli (nem_mod.c:619-690) summed as
//   logf_ik = -(base_k + sum_{d: x_id = 1} delta_kd)
// with LANES OVER FAMILIES and the genomes walked in ascending order: the delta row of a genome is
// the same for every lane, so it comes out of shared memory as a broadcast (no divergent gathers --
// the warp-per-family kernel above spends its time in 32-way divergent L1 accesses), and the sum
// becomes K predicated fp64 adds per (family, genome).  A lane carries DG_R families so that one
// broadcast feeds DG_R * K adds: the kernel is bound by the fp64 pipe, not by the load/store unit.
// Shared memory: delta re-laid as [genome][KT] (zero beyond D).  fp64, fixed order (ascending d).
// =============================================================================================
template <int KT> struct DgR { static constexpr int R = KT <= 4 ? 4 : (KT <= 8 ? 2 : 1); };

template <int KT>
__global__ void __launch_bounds__(256)
k_density_general_tiled(int K, const uint32_t *__restrict__ x, int n, int D, int wpr,
                        const nemk_coef *__restrict__ coef, const uint32_t *__restrict__ f0,
                        const uint32_t *__restrict__ f1, const double *__restrict__ delta,
                        const double *__restrict__ base_g, double *__restrict__ logpf) {
    constexpr int R = DgR<KT>::R;
    extern __shared__ __align__(16) double dg_sd[];          // [wreal * 32][KT]
    __shared__ int s_forb, s_nonfin;
    if (coef->empty_class | coef->halt) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int wreal = (D + 31) >> 5;
    if (threadIdx.x == 0) { s_forb = 0; s_nonfin = 0; }
    __syncthreads();
    for (int idx = threadIdx.x; idx < wreal * 32 * KT; idx += blockDim.x) {
        int d = idx / KT, k = idx - d * KT;
        const double dv = (d < D && k < K) ? delta[(size_t)k * D + d] : 0.0;
        dg_sd[idx] = 0.5 * dv;                               // halved: see the fma below
        if (!(fabs(dv) <= DBL_MAX)) s_nonfin = 1;
    }
    {   // does any class forbid a cell (eps = 0)?  almost never: the test is skipped then
        unsigned any = 0;
        for (int idx = threadIdx.x; idx < K * wpr; idx += blockDim.x) any |= f0[idx] | f1[idx];
        if (any) s_forb = 1;
    }
    __syncthreads();
    const bool forb = s_forb != 0, nonfin = s_nonfin != 0;
    const long long ntiles = ((long long)n + 32 * R - 1) / (32 * R);
    for (long long t = (long long)blockIdx.x * nw + warp; t < ntiles; t += (long long)gridDim.x * nw) {
        const uint32_t *xr[R];
        long long row[R];
        double acc[R][KT];
        unsigned nul[R], v[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            row[r] = t * (32 * R) + r * 32 + lane;
            xr[r] = x + (size_t)(row[r] < n ? row[r] : n - 1) * wpr;
            nul[r] = 0u;
            v[r] = xr[r][0];
#pragma unroll
            for (int k = 0; k < KT; k++) acc[r][k] = 0.0;
        }
        for (int w = 0; w < wreal; w++) {
            const uint32_t live = (w == wreal - 1 && (D & 31)) ? ((1u << (D & 31)) - 1u) : FULL;
            unsigned cur[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                cur[r] = v[r] & live;
                if (w + 1 < wreal) v[r] = xr[r][w + 1];      // next word: in flight during the adds
            }
            if (forb) {
                for (int k = 0; k < K; k++) {
                    const uint32_t m0 = f0[k * wpr + w], m1 = f1[k * wpr + w];
#pragma unroll
                    for (int r = 0; r < R; r++)
                        if ((cur[r] & m1) | (~cur[r] & live & m0)) nul[r] |= 1u << k;
                }
            }
            const double2 *sp = reinterpret_cast<const double2 *>(dg_sd + (size_t)w * 32 * KT);
            if (nonfin) {       // a delta of +-inf (eps = 1 in a .m file): literal conditional adds
#pragma unroll 1
                for (int b = 0; b < 32; b++)
#pragma unroll
                    for (int r = 0; r < R; r++)
                        if ((cur[r] >> b) & 1u) {
#pragma unroll
                            for (int k = 0; k < KT; k++) acc[r][k] += 2.0 * dg_sd[((size_t)w * 32 + b) * KT + k];
                        }
                continue;
            }
#pragma unroll
            for (int b = 0; b < 32; b += 2) {
                double dd[2 * KT];                           // HALF the deltas of genomes b and b + 1
#pragma unroll
                for (int q = 0; q < KT; q++) {
                    double2 t2 = sp[(b >> 1) * KT + q];
                    dd[2 * q] = t2.x; dd[2 * q + 1] = t2.y;
                }
#pragma unroll
                for (int r = 0; r < R; r++) {
                    // bit -> the double 2.0 or 0.0 (high word 0x40000000 or 0): the conditional add
                    // becomes ONE fma(2 or 0, delta / 2, acc) on the fp64 pipe -- same value as
                    // acc + delta, no select instructions (a select costs two per add)
                    const unsigned h0 = (b <= 30 ? cur[r] << (30 - b) : cur[r] >> (b - 30)) & 0x40000000u;
                    const unsigned h1 = (b + 1 <= 30 ? cur[r] << (29 - b) : cur[r] >> (b - 29)) & 0x40000000u;
                    const double m0 = __hiloint2double((int)h0, 0), m1 = __hiloint2double((int)h1, 0);
#pragma unroll
                    for (int k = 0; k < KT; k++) {
                        acc[r][k] = fma(m0, dd[k], acc[r][k]);
                        acc[r][k] = fma(m1, dd[KT + k], acc[r][k]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (row[r] < n) {
#pragma unroll
                for (int k = 0; k < KT; k++)
                    if (k < K)
                        logpf[(size_t)row[r] * K + k] =
                            ((nul[r] >> k) & 1u) ? neg_inf() : coef->lp[k] - (base_g[k] + acc[r][k]);
            }
        }
    }
}

// =============================================================================================
// E-step site update.  ComputeLocalProba (nem_alg.c:2546-2616) in the log domain,
// SumNeighsOfClass (nem_alg.c:2850-2884), ComputeMAP first-max (nem_alg.c:603-615).
// =============================================================================================
template <int KT>
struct SiteCtx {
    double v[KT];
};

// context from hard labels: ctx_k = sum_j w_ij [lab_j == k]; `pick(j)` returns neighbour j's label.
// Neighbours are fetched four at a time (indices, weights, then the four label gathers together) so
// a site costs ~3 dependent memory round trips instead of 2 per neighbour; the sum order stays the
// file order (SumNeighsOfClass, nem_alg.c:2865-2875).
template <int KT, typename Pick>
static __device__ __forceinline__ void ctx_labels(int K, int i, const int32_t *__restrict__ row_ptr,
                                                  const int32_t *__restrict__ col,
                                                  const float *__restrict__ wgt, Pick pick,
                                                  double *ctx) {
#pragma unroll
    for (int k = 0; k < KT; k++) ctx[k] = 0.0;
    if (!row_ptr) return;
    int lo = row_ptr[i], hi = row_ptr[i + 1];
    for (int e = lo; e < hi; e += 4) {
        int j[4];
        float w[4];
        unsigned l[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            bool in = e + q < hi;
            j[q] = in ? col[e + q] : -1;
            w[q] = in ? wgt[e + q] : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) l[q] = j[q] >= 0 ? pick(j[q]) : 255u;
#pragma unroll
        for (int q = 0; q < 4; q++) {
#pragma unroll
            for (int k = 0; k < KT; k++)
                if (l[q] == (unsigned)k) ctx[k] += (double)w[q];
        }
    }
}

// returns arg max (first max); flags: bit0 = all classes have zero density, bit1 = exact tie
template <int KT>
static __device__ __forceinline__ int site_argmax(int K, const double *__restrict__ lp,
                                                  const double *ctx, double beta, int &flags,
                                                  double &margin) {
    double mx = neg_inf();
    int km = 0;
    double sc[KT];
#pragma unroll
    for (int k = 0; k < KT; k++) {
        sc[k] = (k < K) ? lp[k] + beta * ctx[k] : neg_inf();
        if (sc[k] > mx) { mx = sc[k]; km = k; }
    }
    flags = 0;
    margin = neg_inf();   // never skipped
    if (mx == neg_inf()) { flags = 1; return 0; }
    double second = neg_inf();
#pragma unroll
    for (int k = 0; k < KT; k++) {
        if (k != km && k < K) second = fmax(second, sc[k]);
        if (k > km && k < K && sc[k] == mx) flags |= 2;
    }
    margin = mx - second;   // 0 on an exact tie, +inf when every other class has zero density
    return km;
}
template <int KT>
static __device__ __forceinline__ int site_argmax(int K, const double *__restrict__ lp,
                                                  const double *ctx, double beta, int &flags) {
    double margin;
    return site_argmax<KT>(K, lp, ctx, beta, flags, margin);
}

// ---- margin cache (nemk_margins).  All kernels of one sweep derive the same two numbers from the
// coefficients: `test` (a stored margin must exceed it for the site to be skipped) and `store`
// (added to the margins stored by this sweep; it becomes coef->drift when the sweep ends).
struct SweepThr { double test, store; };
static __device__ __forceinline__ SweepThr sweep_thr(int K, const nemk_coef *__restrict__ coef,
                                                     const nemk_margins &mg) {
    SweepThr t;
    t.test = CUDART_INF; t.store = 0.0;
    if (!mg.m || !mg.on || coef->mu_changed) return t;
    double step = 0.0;
    for (int k = 0; k < K; k++) step = fmax(step, coef->dstep[k]);
    if (!(step < CUDART_INF)) return t;   // +inf or NaN: unknown move, evaluate everything
    t.store = coef->drift + 2.0 * step;
    t.test = t.store + 1e-6;              // slack for the rounding of the scores themselves
    return t;
}
static __device__ __forceinline__ void store_margin(const nemk_margins &mg, int il, double margin,
                                                    double store) {
    if (mg.m) mg.m[il] = __double2float_rd(margin + store);   // rounded down: conservative
}

// Speculative sequential sweep bookkeeping: every site that READS i and is visited later
// (larger index) must be re-evaluated when i's label moves.  dirty[] de-duplicates, wl[] is
// the work list of the next round.
// The claims (atomicExch on dirty[]) of four readers are issued back to back and only then
// looked at, and the work-list slots of a group come from ONE atomicAdd: a thread pays one L2
// round trip per four readers instead of two per reader (the fix-up tail is a chain of such
// dependent round trips).
static __device__ __forceinline__ void mark_readers(int i, const int32_t *__restrict__ rrow_ptr,
                                                    const int32_t *__restrict__ rcol,
                                                    int32_t *dirty, int32_t *wl, int32_t *wl_count,
                                                    int row0, int row1, uint8_t *stale_next = nullptr) {
    int lo = rrow_ptr[i], hi = rrow_ptr[i + 1];
    for (int e = lo; e < hi; e += 4) {
        int j[4], was[4];
#pragma unroll
        for (int q = 0; q < 4; q++) j[q] = e + q < hi ? rcol[e + q] : -1;
        int nclaim = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            // only sites this rank owns ([row0,row1), the whole graph on one GPU) are queued here;
            // the owner of a remote reader queues it when it sees i's new label (k_mark_remote)
            const bool own = j[q] >= 0 && j[q] >= row0 && j[q] < row1;
            // readers visited before i (or i itself) keep this sweep's evaluation, which saw i's
            // OLD label: their cached margin is void for the next sweep
            if (stale_next && own && j[q] <= i) stale_next[j[q]] = 1;
            was[q] = (own && j[q] > i) ? atomicExch(&dirty[j[q]], 1) : 1;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) nclaim += was[q] == 0;
        if (nclaim) {
            int base = atomicAdd(wl_count, nclaim);
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (was[q] == 0) wl[base++] = j[q];
        }
    }
}

// ---- ncem, parallel (Jacobi) update; also round 0 of the speculative sequential sweep
// (dirty != nullptr): changed sites queue their later readers for the fix-up rounds.
// ---- high-degree sites ("hubs": backbone families carry ~100 island anchors).  One thread walking
// a hub's neighbour list serialises ~3 memory round trips per 4 neighbours and becomes the tail of
// the whole sweep, so sites with more than HEAVY_DEG neighbours are evaluated by a full warp: the
// 32 lanes fetch 32 neighbours (index, weight, label) at once, then the weights are ADDED IN FILE
// ORDER through shuffles, so the float64 sum is bit-identical to the one-thread sum
// (SumNeighsOfClass order, nem_alg.c:2865-2875).  Every lane ends with the same ctx.
#define HEAVY_DEG 16
template <int KT, typename Pick>
static __device__ __forceinline__ void ctx_labels_warp(int K, int i, const int32_t *__restrict__ row_ptr,
                                                       const int32_t *__restrict__ col,
                                                       const float *__restrict__ wgt, Pick pick,
                                                       double *ctx, bool any_order = false) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < KT; k++) ctx[k] = 0.0;
    if (!row_ptr) return;
    int lo = row_ptr[i], hi = row_ptr[i + 1];
    if (any_order) {
        // integer weights: every partial sum is exact, so the lanes add their strided share (four
        // independent loads in flight per lane) and a butterfly adds the lanes -- same bits as the
        // file-order sum, without its serial chain of dependent loads
        for (int e0 = lo + lane; e0 < hi; e0 += 128) {
            int j[4];
            float w[4];
            unsigned l[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                int e = e0 + 32 * q;
                bool in = e < hi;
                j[q] = in ? col[e] : -1;
                w[q] = in ? wgt[e] : 0.f;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) l[q] = j[q] >= 0 ? pick(j[q]) : 255u;
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
                for (int k = 0; k < KT; k++)
                    if (l[q] == (unsigned)k) ctx[k] += (double)w[q];
        }
#pragma unroll
        for (int k = 0; k < KT; k++)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ctx[k] += __shfl_xor_sync(FULL, ctx[k], o);
        return;
    }
    for (int e0 = lo; e0 < hi; e0 += 32) {
        int e = e0 + lane;
        bool in = e < hi;
        int j = in ? col[e] : -1;
        float w = in ? wgt[e] : 0.f;
        unsigned l = in ? pick(j) : 255u;
        int cnt = min(32, hi - e0);
        for (int q = 0; q < cnt; q++) {
            unsigned lq = __shfl_sync(FULL, l, q);
            float wq = __shfl_sync(FULL, w, q);
#pragma unroll
            for (int k = 0; k < KT; k++)
                if (lq == (unsigned)k) ctx[k] += (double)wq;
        }
    }
}

// ---- context of 32 CONSECUTIVE sites by one warp (the dense sweeps and the criteria): the CSR
// segments of consecutive rows are contiguous, so the warp streams [row_ptr[i0], row_ptr[i0+32])
// with coalesced loads (index, weight, then the label gather: 32 entries per instruction) into a
// shared-memory stage, and every lane then adds ITS entries in file order (fp64, same sum as
// SumNeighsOfClass, nem_alg.c:2865-2875).  Replaces 8 scattered 4-byte loads per site by ~1/4 of
// the L1 sectors and a third of the instructions.  Lanes with an empty range (lo == hi: out of
// range, or a hub handled elsewhere) only help loading; chunks no live lane intersects are skipped.
// Must be called by all 32 lanes.
#define COOP_C 4                    // entries per lane per chunk
#define COOP_CHUNK (32 * COOP_C)
template <int KT>
static __device__ __forceinline__ void ctx_labels_coop(int lo, int hi, int seg_lo, int seg_hi,
                                                       const int32_t *__restrict__ col,
                                                       const float *__restrict__ wgt,
                                                       const uint8_t *__restrict__ lab,
                                                       float *s_w /*[COOP_CHUNK]*/,
                                                       uint8_t *s_l /*[COOP_CHUNK]*/, double *ctx) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < KT; k++) ctx[k] = 0.0;
    for (int chunk = seg_lo; chunk < seg_hi; chunk += COOP_CHUNK) {
        const int chunk_end = min(chunk + COOP_CHUNK, seg_hi);
        if (!__any_sync(FULL, lo < chunk_end && hi > chunk)) continue;
        int j[COOP_C];
        float w[COOP_C];
#pragma unroll
        for (int c = 0; c < COOP_C; c++) {
            int e = chunk + c * 32 + lane;
            bool in = e < chunk_end;
            j[c] = in ? col[e] : -1;
            w[c] = in ? wgt[e] : 0.f;
        }
#pragma unroll
        for (int c = 0; c < COOP_C; c++) {
            s_w[c * 32 + lane] = w[c];
            s_l[c * 32 + lane] = j[c] >= 0 ? lab[j[c]] : (uint8_t)255;
        }
        __syncwarp();
        const int a = max(lo, chunk) - chunk, b = min(hi, chunk_end) - chunk;
        for (int q = a; q < b; q++) {
            unsigned l = s_l[q];
            double wq = (double)s_w[q];
#pragma unroll
            for (int k = 0; k < KT; k++)
                if (l == (unsigned)k) ctx[k] += wq;
        }
        __syncwarp();
    }
}

static __device__ __forceinline__ void mark_readers_warp(int i, const int32_t *__restrict__ rrow_ptr,
                                                         const int32_t *__restrict__ rcol,
                                                         int32_t *dirty, int32_t *wl,
                                                         int32_t *wl_count, int row0, int row1,
                                                         uint8_t *stale_next = nullptr) {
    int lo = rrow_ptr[i], hi = rrow_ptr[i + 1];
    for (int e = lo + (threadIdx.x & 31); e < hi; e += 32) {
        int j = rcol[e];
        if (stale_next && j <= i && j >= row0 && j < row1) stale_next[j] = 1;
        if (j > i && j >= row0 && j < row1 && atomicExch(&dirty[j], 1) == 0)
            wl[atomicAdd(wl_count, 1)] = j;
    }
}

// Rows [row0, row0+n_loc) of the GLOBAL graph are this rank's (row0 = 0, n_loc = N on one GPU);
// labels, CSR, dirty flags and work lists are indexed by global family id, logpf by local row.
template <int KT>
__global__ void __launch_bounds__(256, KT <= 4 ? JAC_MINB_SMALLK : JAC_MINB)
k_sweep_ncem_jacobi(int K, int row0, int n_loc, const nemk_lpsrc lps,
                    const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                    const float *__restrict__ wgt, double beta, const uint8_t *__restrict__ lab_in,
                    uint8_t *__restrict__ lab_out, int32_t *dirty, int32_t *wl, int32_t *wl_count,
                    const int32_t *__restrict__ rrow_ptr, const int32_t *__restrict__ rcol,
                    const int32_t *__restrict__ heavy, int n_heavy, int heavy_blocks,
                    nemk_counters *cnt, const int32_t *__restrict__ skip, int copy_ranks,
                    int shard_len, const nemk_margins mg) {
    if (skip && (skip[0] | skip[1])) return;
    const int lane = threadIdx.x & 31;
    int changed = 0, flags = 0, kept_site = 0;
    if ((int)blockIdx.x >= heavy_blocks && copy_ranks > 1) {
        // row shards: the other ranks' labels start the sweep at their previous value
        const int mine = row0 / shard_len;
        for (int s = 0; s < JAC_SPT; s++) {
            int q = (blockIdx.x - heavy_blocks) * JAC_TILE + s * 256 + threadIdx.x;
            if (q >= shard_len) break;
            for (int r = 0; r < copy_ranks; r++)
                if (r != mine) lab_out[(size_t)r * shard_len + q] = lab_in[(size_t)r * shard_len + q];
        }
    }
    const SweepThr thr = sweep_thr(K, lps.coef, mg);
    const bool may_skip = mg.m && thr.test < CUDART_INF;
    if ((int)blockIdx.x < heavy_blocks) {
        // hubs first (longest work): one warp per site
        int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        if (wid >= n_heavy) return;
        int i = heavy[wid], il = i - row0;
        if (may_skip && !mg.stale_cur[i] && (double)mg.m[il] > thr.test && lab_in[i] != 255) {
            if (lane == 0) {
                lab_out[i] = lab_in[i];   // margin > possible move, context unchanged
#if JAC_HUB_GUARD
                atomicAdd(&cnt->kept, 1);   // hubs are counted here, not by the light threads
#endif
            }
            return;
        }
        double ctx[KT];
        ctx_labels_warp<KT>(K, i, row_ptr, col, wgt, [&](int j) { return (unsigned)lab_in[j]; }, ctx,
                            lps.wsum_any_order != 0);
        double lpv[KT], margin;
        load_lp<KT>(lps, K, (size_t)il, lpv);
        int km = site_argmax<KT>(K, lpv, ctx, beta, flags, margin);
        int ch = (km != (int)lab_in[i]);
        if (lane == 0) {
            lab_out[i] = (uint8_t)km;
            store_margin(mg, il, margin, thr.store);
            if (mg.m && mg.stale_cur[i]) mg.stale_cur[i] = 0;
        }
        if (ch && dirty)
            mark_readers_warp(i, rrow_ptr, rcol, dirty, wl, wl_count, row0, row0 + n_loc, mg.stale_next);
        if (lane == 0) {
            if (ch) atomicAdd(&cnt->changed, 1);
            if (flags & 1) atomicAdd(&cnt->allnul, 1);
            if (flags & 2) atomicAdd(&cnt->ties, 1);
        }
        return;
    }
    // Light sites: a CTA covers JAC_SPT sub-tiles of 256 consecutive sites; thread t owns site t of
    // every sub-tile.  The margin test of all its sites is issued at once (3 * JAC_SPT independent
    // loads in flight per thread), so the dense round is ~one wave of CTAs instead of four chains
    // of dependent latencies back to back.
    __shared__ float s_w[8][COOP_CHUNK];
    __shared__ uint8_t s_l[8][COOP_CHUNK];
    __shared__ int s_act[JAC_TILE];
    __shared__ int s_nact;
    const int tile0 = (blockIdx.x - heavy_blocks) * JAC_TILE;
    const int32_t *rp = beta != 0.0 ? row_ptr : nullptr;
    unsigned actm = 0u;   // bit s: this thread's site of sub-tile s has to be evaluated
    int sparse = 0;
    if (may_skip) {
        // margin cache: the site keeps its label when its stored margin exceeds everything theta
        // can have moved since and no later-or-equal neighbour changed in the previous sweep
        uint8_t st[JAC_SPT], lb[JAC_SPT];
        float mv[JAC_SPT];
#if JAC_HUB_GUARD
        // A hub's label, margin and stale flag belong to its warp in the hub blocks, which may have
        // evaluated the site and cleared its flag before this thread looks: a "kept" hub is not
        // copied here (the hub warp writes lab_out in every case), or a late copy of the old label
        // could land on top of the new one.
        unsigned hubm = 0u;
        const bool hubs = heavy_blocks && rp;
#endif
#pragma unroll
        for (int s = 0; s < JAC_SPT; s++) {
            const int sl = tile0 + s * 256 + (int)threadIdx.x;
            const bool valid = sl < n_loc;
            st[s] = valid ? mg.stale_cur[row0 + sl] : (uint8_t)1;
            mv[s] = valid ? mg.m[sl] : 0.f;
            lb[s] = valid ? lab_in[row0 + sl] : (uint8_t)255;
#if JAC_HUB_GUARD
            if (hubs && valid && rp[row0 + sl + 1] - rp[row0 + sl] > HEAVY_DEG) hubm |= 1u << s;
#endif
        }
        if (threadIdx.x == 0) s_nact = 0;
        __syncthreads();
        unsigned ba[JAC_SPT];
        int wtotal = 0;
#pragma unroll
        for (int s = 0; s < JAC_SPT; s++) {
            const int sl = tile0 + s * 256 + (int)threadIdx.x;
            const bool valid = sl < n_loc;
            const bool keep = valid && !st[s] && (double)mv[s] > thr.test && lb[s] != 255;
#if JAC_HUB_GUARD
            if (keep && !((hubm >> s) & 1u)) lab_out[row0 + sl] = lb[s];
            kept_site += keep && !((hubm >> s) & 1u);
#else
            if (keep) lab_out[row0 + sl] = lb[s];
            kept_site += keep;
#endif
            if (valid && !keep) actm |= 1u << s;
            ba[s] = __ballot_sync(FULL, valid && !keep);
            wtotal += __popc(ba[s]);
        }
        // How many sites of this CTA are left?  Few (steady state: a few per cent): compact them so
        // that a few warps walk the dependent loads and the others retire at once.  Many: the
        // warp-cooperative segment path below.
        int base = 0;
        if (lane == 0 && wtotal) base = atomicAdd(&s_nact, wtotal);
        base = __shfl_sync(FULL, base, 0);
#pragma unroll
        for (int s = 0; s < JAC_SPT; s++) {
            if ((actm >> s) & 1u)
                s_act[base + __popc(ba[s] & ((1u << lane) - 1u))] = tile0 + s * 256 + (int)threadIdx.x;
            base += __popc(ba[s]);
        }
        __syncthreads();
        sparse = s_nact <= JAC_SPARSE_MAX;
    } else {
#pragma unroll
        for (int s = 0; s < JAC_SPT; s++)
            if (tile0 + s * 256 + (int)threadIdx.x < n_loc) actm |= 1u << s;
    }
    if (sparse) {
        const int nact = s_nact;
        for (int q = threadIdx.x; q < nact; q += blockDim.x) {
            const int sl = s_act[q], si = row0 + sl;
            const bool hv = heavy_blocks && rp && (rp[si + 1] - rp[si] > HEAVY_DEG);
            if (hv) continue;   // evaluated by the hub blocks
            double ctx[KT], lpv[KT], margin;
            load_lp<KT>(lps, K, (size_t)sl, lpv);   // requested before the neighbour walk, not after
            const int lin = (int)lab_in[si];
            ctx_labels<KT>(K, si, rp, col, wgt, [&](int j) { return (unsigned)lab_in[j]; }, ctx);
            int fl;
            int km = site_argmax<KT>(K, lpv, ctx, beta, fl, margin);
            lab_out[si] = (uint8_t)km;
            store_margin(mg, sl, margin, thr.store);
            if (mg.stale_cur[si]) mg.stale_cur[si] = 0;
            int ch = (km != lin);
            if (ch && dirty)
                mark_readers(si, rrow_ptr, rcol, dirty, wl, wl_count, row0, row0 + n_loc, mg.stale_next);
            changed += ch;          // a thread may take several sites here: counted directly
            if (fl & 1) atomicAdd(&cnt->allnul, 1);
            if (fl & 2) atomicAdd(&cnt->ties, 1);
        }
        if (changed) atomicAdd(&cnt->changed, changed);
        kept_site = __reduce_add_sync(FULL, kept_site);
        if (lane == 0 && kept_site) atomicAdd(&cnt->kept, kept_site);
        return;
    }
    int nallnul = 0, nties = 0;
    for (int s = 0; s < JAC_SPT; s++) {
        const int il = tile0 + s * 256 + (int)threadIdx.x, i = row0 + il;
        if (il - lane >= n_loc) break;   // warps entirely out of range have nothing to do
        const bool act = (actm >> s) & 1u;
        int lo = 0, hi = 0;
        if (rp && act) { lo = rp[i]; hi = rp[i + 1]; }
        const bool is_heavy = heavy_blocks && (hi - lo > HEAVY_DEG);
        double ctx[KT], lpv[KT];
#pragma unroll
        for (int k = 0; k < KT; k++) ctx[k] = 0.0;
        if (act) load_lp<KT>(lps, K, (size_t)il, lpv);   // requested before the neighbour walk
        if (rp) {
            int seg_lo = __reduce_min_sync(FULL, act ? lo : 0x7fffffff);
            int seg_hi = __reduce_max_sync(FULL, act ? hi : 0);
            if (is_heavy) lo = hi = 0;   // evaluated by the hub blocks
            if (seg_lo < seg_hi)
                ctx_labels_coop<KT>(lo, hi, seg_lo, seg_hi, col, wgt, lab_in, s_w[threadIdx.x >> 5],
                                    s_l[threadIdx.x >> 5], ctx);
        }
        if (act && !is_heavy) {
            double margin;
            int fl;
            int km = site_argmax<KT>(K, lpv, ctx, beta, fl, margin);
            lab_out[i] = (uint8_t)km;
            store_margin(mg, il, margin, thr.store);
            if (mg.m && mg.stale_cur[i]) mg.stale_cur[i] = 0;
            const int ch = (km != (int)lab_in[i]);
            if (ch && dirty)
                mark_readers(i, rrow_ptr, rcol, dirty, wl, wl_count, row0, row0 + n_loc, mg.stale_next);
            changed += ch;
            nallnul += fl & 1;
            nties += (fl >> 1) & 1;
        }
    }
    changed = __reduce_add_sync(FULL, changed);
    nallnul = __reduce_add_sync(FULL, nallnul);
    nties = __reduce_add_sync(FULL, nties);
    kept_site = __reduce_add_sync(FULL, kept_site);
    if (lane == 0) {
        if (changed) atomicAdd(&cnt->changed, changed);
        if (nallnul) atomicAdd(&cnt->allnul, nallnul);
        if (nties) atomicAdd(&cnt->ties, nties);
        if (kept_site) atomicAdd(&cnt->kept, kept_site);
    }
}

// ---- ncem, speculative sequential sweep, fix-up rounds (no host round trips).
// The in-place index-order sweep (UPDATE_SEQ, nem_alg.c:2378-2383) defines
//     cur_i = F_i( cur_j for j<i , old_j for j>=i )
// a triangular system with a unique solution.  Round 0 (the Jacobi kernel) evaluates F with old
// everywhere; each later round re-evaluates only the sites one of whose lower-index inputs
// moved, until the work list is empty (at most DAG-depth rounds, in practice a handful because
// the data term dominates).  A site clears its dirty flag BEFORE reading its inputs and every
// change re-queues its later readers AFTER publishing the new label, so no update is lost and
// the fixed point reached is the sequential sweep's result whatever the interleaving.
// The first rounds (long work lists) run grid-wide, one launch per round; the tail runs in one
// CTA that loops until the list is empty.
// Re-evaluate site i.  NOTE (kept as a warning): letting the thread that changed i go straight on
// with a reader it has just claimed ("chasing" the chain i -> i+1 -> ... inside one round) is NOT
// safe: the claimed reader may be in the middle of its evaluation by another thread of the same
// round (it cleared its flag before reading), and the two evaluations would race on lab_cur[j] --
// the stale one can land last and nobody re-queues j.  Rounds separated by a barrier are what
// guarantees that a site is evaluated by one thread at a time; measured, the chase bought nothing
// anyway (the tail is bound by the dependent loads of a link, not by the barriers).
template <int KT>
static __device__ __forceinline__ int fixup_site(int K, int i, int row0, int row1,
                                                 const nemk_lpsrc &lps,
                                                 const int32_t *__restrict__ row_ptr,
                                                 const int32_t *__restrict__ col,
                                                 const float *__restrict__ wgt, double beta,
                                                 const uint8_t *__restrict__ lab_old,
                                                 uint8_t *lab_cur, int32_t *dirty, int32_t *next_list,
                                                 int32_t *next_cnt,
                                                 const int32_t *__restrict__ rrow_ptr,
                                                 const int32_t *__restrict__ rcol,
                                                 const nemk_margins &mg, double thr_store) {
    // Everything that does not depend on the neighbours' labels is requested BEFORE the flag is
    // cleared, so that it travels together with the fence instead of after it: the site's own
    // current label (only this thread writes lab_cur[i] during the round), its data term (constant
    // during a sweep) and the first sectors of its neighbour list (immutable; prefetched into L1).
    if (FIXUP_HOIST && row_ptr) {
        const int lo = row_ptr[i], hi = row_ptr[i + 1];
        if (lo < hi) {
            prefetch_l1(col + lo); prefetch_l1(wgt + lo);
            prefetch_l1(col + hi - 1); prefetch_l1(wgt + hi - 1);
        }
    }
    int was = 0;
    double lpv[KT];
    if (FIXUP_HOIST) {
        was = __ldcg(lab_cur + i);
        load_lp<KT>(lps, K, (size_t)(i - row0), lpv);
    }
#if FIXUP_ACQ
    // Every access to dirty[] is a read-modify-write, so the flag's modification order is total:
    // either the changer's claim (store label; fence; exchange 1) precedes this exchange -- then
    // this ACQUIRE exchange synchronises with it and the label loads below see the new label -- or
    // it follows, finds 0 and re-queues the site.  No full fence on this side.
    {
        int32_t was_flag;
        asm volatile("atom.acquire.gpu.global.exch.b32 %0, [%1], %2;"
                     : "=r"(was_flag) : "l"(dirty + i), "r"(0) : "memory");
        (void)was_flag;
    }
#else
    atomicExch(&dirty[i], 0);
    __threadfence();
#endif
    double ctx[KT];
    ctx_labels<KT>(K, i, row_ptr, col, wgt,
                   [&](int j) { return (unsigned)(j < i ? __ldcg(lab_cur + j) : lab_old[j]); }, ctx);
    int flags;
    double margin;
    if (!FIXUP_HOIST) load_lp<KT>(lps, K, (size_t)(i - row0), lpv);
    int km = site_argmax<KT>(K, lpv, ctx, beta, flags, margin);
    store_margin(mg, i - row0, margin, thr_store);   // the LAST evaluation of a site is its final one
    if (!FIXUP_HOIST) was = __ldcg(lab_cur + i);
    if (km == was) return 0;
    lab_cur[i] = (uint8_t)km;
    __threadfence();
    mark_readers(i, rrow_ptr, rcol, dirty, next_list, next_cnt, row0, row1, mg.stale_next);
    int old = lab_old[i];
    return (km != old) - (was != old);
}

// the same for a hub, by a whole warp (every lane returns the same delta)
template <int KT>
static __device__ __forceinline__ int fixup_site_warp(int K, int i, int row0, int row1,
                                                      const nemk_lpsrc &lps,
                                                      const int32_t *__restrict__ row_ptr,
                                                      const int32_t *__restrict__ col,
                                                      const float *__restrict__ wgt, double beta,
                                                      const uint8_t *__restrict__ lab_old,
                                                      uint8_t *lab_cur, int32_t *dirty,
                                                      int32_t *next_list, int32_t *next_cnt,
                                                      const int32_t *__restrict__ rrow_ptr,
                                                      const int32_t *__restrict__ rcol,
                                                      const nemk_margins &mg, double thr_store) {
    const int lane = threadIdx.x & 31;
    double lpv[KT];
    int was = (int)__ldcg(lab_cur + i);   // only this warp writes lab_cur[i] during the round
    load_lp<KT>(lps, K, (size_t)(i - row0), lpv);
    if (lane == 0) { atomicExch(&dirty[i], 0); __threadfence(); }
    __syncwarp();
    double ctx[KT];
    ctx_labels_warp<KT>(K, i, row_ptr, col, wgt,
                        [&](int j) { return (unsigned)(j < i ? __ldcg(lab_cur + j) : lab_old[j]); }, ctx,
                        lps.wsum_any_order != 0);
    int flags;
    double margin;
    int km = site_argmax<KT>(K, lpv, ctx, beta, flags, margin);
    if (lane == 0) store_margin(mg, i - row0, margin, thr_store);
    was = __shfl_sync(FULL, was, 0);
    if (km == was) return 0;
    if (lane == 0) { lab_cur[i] = (uint8_t)km; __threadfence(); }
    __syncwarp();
    mark_readers_warp(i, rrow_ptr, rcol, dirty, next_list, next_cnt, row0, row1, mg.stale_next);
    int old = lab_old[i];
    return (km != old) - (was != old);
}

// one work-list slot per lane; light sites by their lane, hubs afterwards by the whole warp.
// `idx` must be warp-uniform in validity order (idx = base + lane).  Returns this lane's delta.
template <int KT>
static __device__ __forceinline__ int fixup_items(int K, int idx, int count,
                                                  const int32_t *__restrict__ cur_list, int row0,
                                                  int row1, const nemk_lpsrc &lps,
                                                  const int32_t *__restrict__ row_ptr,
                                                  const int32_t *__restrict__ col,
                                                  const float *__restrict__ wgt, double beta,
                                                  const uint8_t *__restrict__ lab_old,
                                                  uint8_t *lab_cur, int32_t *dirty,
                                                  int32_t *next_list, int32_t *next_cnt,
                                                  const int32_t *__restrict__ rrow_ptr,
                                                  const int32_t *__restrict__ rcol,
                                                  const nemk_margins &mg, double thr_store) {
    const int lane = threadIdx.x & 31;
    int i = idx < count ? cur_list[idx] : -1;
    bool hub = i >= 0 && (row_ptr[i + 1] - row_ptr[i] > HEAVY_DEG);
    int d = 0;
    if (i >= 0 && !hub)
        d = fixup_site<KT>(K, i, row0, row1, lps, row_ptr, col, wgt, beta, lab_old, lab_cur, dirty,
                           next_list, next_cnt, rrow_ptr, rcol, mg, thr_store);
    unsigned hm = __ballot_sync(FULL, hub);
    while (hm) {
        int src = __ffs(hm) - 1;
        hm &= hm - 1;
        int site = __shfl_sync(FULL, i, src);
        int r = fixup_site_warp<KT>(K, site, row0, row1, lps, row_ptr, col, wgt, beta, lab_old,
                                    lab_cur, dirty, next_list, next_cnt, rrow_ptr, rcol, mg, thr_store);
        if (lane == src) d = r;
    }
    return d;
}

// Work lists: two lists used alternately, FOUR rotating counters.  Round r consumes list r&1 with
// count wl_cnt[r&3], appends to list (r+1)&1 through wl_cnt[(r+1)&3] and clears wl_cnt[(r+2)&3]
// (idle during round r), so no memset sits between rounds.  The tail kernel leaves all four at 0.
template <int KT>
__global__ void __launch_bounds__(256)
k_sweep_ncem_fixup_round(int K, int row0, int row1, const nemk_lpsrc lps,
                         const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                         const float *__restrict__ wgt, double beta,
                         const uint8_t *__restrict__ lab_old, uint8_t *lab_cur, int32_t *dirty,
                         int32_t *wl_a, int32_t *wl_b, int32_t *wl_cnt, int round,
                         const int32_t *__restrict__ rrow_ptr, const int32_t *__restrict__ rcol,
                         nemk_counters *cnt, const int32_t *__restrict__ skip, const nemk_margins mg) {
    if (skip && (skip[0] | skip[1])) return;
    const double thr_store = sweep_thr(K, lps.coef, mg).store;
    const int32_t *cur_list = (round & 1) ? wl_b : wl_a;
    int32_t *next_list = (round & 1) ? wl_a : wl_b;
    int count = wl_cnt[round & 3];
    int32_t *next_cnt = &wl_cnt[(round + 1) & 3];
    if (blockIdx.x == 0 && threadIdx.x == 0) wl_cnt[(round + 2) & 3] = 0;
    int dchanged = 0;
    for (int base = blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < count;
         base += gridDim.x * blockDim.x)
        dchanged += fixup_items<KT>(K, base + (threadIdx.x & 31), count, cur_list, row0, row1, lps,
                                    row_ptr, col, wgt, beta, lab_old, lab_cur, dirty, next_list,
                                    next_cnt, rrow_ptr, rcol, mg, thr_store);
    if (dchanged) atomicAdd(&cnt->changed, dchanged);
    if (blockIdx.x == 0 && threadIdx.x == 0 && count) atomicAdd(&cnt->nfix, 1);
}

static __device__ void iter_end_body(int world, const nemk_counters *cnt_all,
                                     const nemk_iter_status *st, nemk_coef *coef, int decide,
                                     int ncem, int conv, float thr, nemk_host_status *host,
                                     unsigned long long seq);

// Tail of the fix-up rounds in ONE launch: a thread-block cluster of FX_CLUSTER CTAs loops over the
// rounds with a hardware cluster barrier between them (release/acquire at cluster scope, after a
// device fence for the label/work-list stores) until the work list is empty.
/* FX_CLUSTER: CTAs of the cluster (compile-time knob, top of the file) */
template <int KT>
__global__ void __cluster_dims__(FX_CLUSTER, 1, 1) __launch_bounds__(1024)
k_sweep_ncem_fixup(int K, int row0, int row1, const nemk_lpsrc lps,
                   const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                   const float *__restrict__ wgt, double beta, const uint8_t *__restrict__ lab_old,
                   uint8_t *lab_cur, int32_t *dirty, int32_t *wl_a, int32_t *wl_b,
                   int32_t *wl_cnt, int round, const int32_t *__restrict__ rrow_ptr,
                   const int32_t *__restrict__ rcol, nemk_counters *cnt,
                   const int32_t *__restrict__ skip, const nemk_iter_end_args fused,
                   const nemk_margins mg) {
    namespace cgx = cooperative_groups;
    cgx::cluster_group cluster = cgx::this_cluster();
    const int crank = (int)cluster.block_rank();
    if (skip && (skip[0] | skip[1])) {   // nothing ran: the status is still due
        if (fused.host && crank == 0 && threadIdx.x == 0)
            iter_end_body(fused.world, fused.cnt_all, fused.st, fused.coef, fused.decide, fused.ncem,
                          fused.conv, fused.thr, fused.host, fused.seq);
        return;
    }
    int rounds = 0, dchanged = 0;
    const double thr_store = sweep_thr(K, lps.coef, mg).store;
    for (;; round++) {
        int32_t *cur_list = (round & 1) ? wl_b : wl_a, *next_list = (round & 1) ? wl_a : wl_b;
        int32_t *next_cnt = &wl_cnt[(round + 1) & 3];
        // nobody appends to list round&1 during this round: every thread reads the same count
        int count = *(volatile int32_t *)&wl_cnt[round & 3];
        if (crank == 0 && threadIdx.x == 0) wl_cnt[(round + 2) & 3] = 0;   // idle counter
        if (count == 0) break;
        rounds++;
        for (int base = crank * 1024 + (threadIdx.x & ~31); base < count; base += FX_CLUSTER * 1024)
            dchanged += fixup_items<KT>(K, base + (threadIdx.x & 31), count, cur_list, row0, row1,
                                        lps, row_ptr, col, wgt, beta, lab_old, lab_cur, dirty,
                                        next_list, next_cnt, rrow_ptr, rcol, mg, thr_store);
        __threadfence();
        cluster.sync();
    }
    cluster.sync();   // every CTA has read the last (zero) count
    if (crank == 0 && threadIdx.x < 4) wl_cnt[threadIdx.x] = 0;
    // the sweep is over: its margins carry thr_store, the drift the next sweep starts from
    if (mg.m && crank == 0 && threadIdx.x == 0) const_cast<nemk_coef *>(lps.coef)->drift = thr_store;
    if (dchanged) atomicAdd(&cnt->changed, dchanged);
    if (crank == 0 && threadIdx.x == 0 && rounds) atomicAdd(&cnt->nfix, rounds);
    if (fused.host) {   // end of the sweep = end of the iteration: decide + publish here (k_iter_end)
        __threadfence();
        cluster.sync();
        if (crank == 0 && threadIdx.x == 0)
            iter_end_body(fused.world, fused.cnt_all, fused.st, fused.coef, fused.decide, fused.ncem,
                          fused.conv, fused.thr, fused.host, fused.seq);
    }
}

// Row-sharded sweep, after a label exchange.  EVERY rank scans ALL families (1 byte each):
//  - a label that differs from the one seen at the previous exchange has "moved": its later
//    readers on OTHER ranks must be re-evaluated by their owners.  This rank queues the ones it
//    owns; every (reader, moved label) pair across ranks is counted, so all ranks -- which hold the
//    same labels and the same `seen` -- compute the SAME pending total without exchanging counters;
//  - changed = labels that differ from the sweep's input labels, over all families (the
//    convergence test of the iteration, again identical on every rank).
__global__ void __launch_bounds__(256)
k_mark_remote(int n_glob, int row0, int row1, int shard_len, const uint8_t *__restrict__ lab_cur,
              const uint8_t *__restrict__ lab_in, const uint8_t *seen_in, uint8_t *seen_out,
              int32_t *dirty, int32_t *wl, int32_t *wl_count, const int32_t *__restrict__ rrow_ptr,
              const int32_t *__restrict__ rcol, nemk_counters *cnt,
              const int32_t *__restrict__ skip) {
    if (skip && (skip[0] | skip[1])) return;   // the sweep did not run: lab_cur holds nothing new
    int pend = 0, changed = 0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_glob; j += gridDim.x * blockDim.x) {
        uint8_t c = lab_cur[j];
        changed += c != lab_in[j];
        bool moved = c != seen_in[j];
        if (moved || seen_in != seen_out) seen_out[j] = c;
        if (!moved) continue;
        const int own_lo = (j / shard_len) * shard_len, own_hi = own_lo + shard_len;   // j's rank
        const bool j_mine = j >= row0 && j < row1;
        int lo = rrow_ptr[j], hi = rrow_ptr[j + 1];
        for (int e = lo; e < hi; e++) {
            int i = rcol[e];
            if (i <= j || (i >= own_lo && i < own_hi)) continue;   // earlier, or same rank as j
            pend++;
            if (!j_mine && i >= row0 && i < row1 && atomicExch(&dirty[i], 1) == 0)
                wl[atomicAdd(wl_count, 1)] = i;
        }
    }
    pend = __reduce_add_sync(FULL, pend);
    changed = __reduce_add_sync(FULL, changed);
    if ((threadIdx.x & 31) == 0) {
        if (pend) atomicAdd(&cnt->pending, pend);
        if (changed) atomicAdd(&cnt->changed_glob, changed);
    }
}

// ---- row-sharded sweep, SPARSE label exchange.  In the steady state a rank moves a few hundred
// labels per sweep: instead of all-gathering whole 1-byte-per-family slices (and scanning them),
// every rank packs (family, label) pairs for its own rows whose label differs from the one all
// ranks last saw, the fixed-size blocks are all-gathered, and every rank applies ALL blocks:
// remote labels are written into its copy, `seen` is refreshed, the later readers it owns are
// queued, and the cross-rank (reader, label) pairs are counted -- the same number on every rank.
// Block layout (int32 words): [0] entries wanted (> cap = overflow: nothing is applied and the
// host falls back to a full exchange), [1] this rank's net changed-label count, then cap x
// (family, label).
__global__ void __launch_bounds__(256)
k_delta_pack(int row0, int n_loc, int cap, const uint8_t *__restrict__ lab_cur,
             const uint8_t *__restrict__ seen, const nemk_counters *__restrict__ cnt, int32_t *block,
             const int32_t *__restrict__ skip) {
    if (skip && (skip[0] | skip[1])) return;
    int il = blockIdx.x * blockDim.x + threadIdx.x;
    if (il == 0) block[1] = cnt->changed;
    bool mv = false;
    int i = row0 + il;
    uint8_t c = 0;
    if (il < n_loc) { c = lab_cur[i]; mv = c != seen[i]; }
    unsigned m = __ballot_sync(FULL, mv);
    int lane = threadIdx.x & 31, base = 0;
    if (m && lane == 0) base = atomicAdd(&block[0], __popc(m));
    base = __shfl_sync(FULL, base, 0);
    if (mv) {
        int pos = base + __popc(m & ((1u << lane) - 1u));
        if (pos < cap) { block[2 + 2 * pos] = i; block[3 + 2 * pos] = (int)c; }
    }
}

__global__ void __launch_bounds__(256)
k_delta_apply(int world, int cap, int block_words, const int32_t *__restrict__ blocks, int row0,
              int row1, int shard_len, uint8_t *lab_cur, uint8_t *seen, int32_t *dirty, int32_t *wl,
              int32_t *wl_count, const int32_t *__restrict__ rrow_ptr,
              const int32_t *__restrict__ rcol, nemk_counters *cnt,
              const int32_t *__restrict__ skip) {
    if (skip && (skip[0] | skip[1])) return;
    bool overflow = false;
    int changed = 0;
    for (int r = 0; r < world; r++) {
        int c = blocks[(size_t)r * block_words];
        overflow |= c > cap;
        changed += blocks[(size_t)r * block_words + 1];
    }
    if (overflow) {   // uniform over the grid and over the ranks: the host sees pending < 0
        if (blockIdx.x == 0 && threadIdx.x == 0) cnt->pending = -1;
        return;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) cnt->changed_glob = changed;
    int pend = 0;
    for (int r = blockIdx.y; r < world; r += gridDim.y) {
        const int32_t *b = blocks + (size_t)r * block_words;
        const int count = b[0];
        const int own_lo = r * shard_len, own_hi = own_lo + shard_len;
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
            const int j = b[2 + 2 * e];
            const uint8_t l = (uint8_t)b[3 + 2 * e];
            const bool j_mine = j >= row0 && j < row1;
            if (!j_mine) lab_cur[j] = l;
            seen[j] = l;
            int lo = rrow_ptr[j], hi = rrow_ptr[j + 1];
            for (int q = lo; q < hi; q++) {
                int i = rcol[q];
                if (i <= j || (i >= own_lo && i < own_hi)) continue;   // earlier, or same rank as j
                pend++;
                if (!j_mine && i >= row0 && i < row1 && atomicExch(&dirty[i], 1) == 0)
                    wl[atomicAdd(wl_count, 1)] = i;
            }
        }
    }
    pend = __reduce_add_sync(FULL, pend);
    if ((threadIdx.x & 31) == 0 && pend) atomicAdd(&cnt->pending, pend);
}

// ---- ncem, level-scheduled exact sequential sweep (reference order), in place.
// sites[] is sorted by (level, index); no two sites of a level are neighbours, so a level is
// updated in parallel; levels run in order (one launch per wide level, or one CTA walking a run
// of narrow levels with __syncthreads between them).
template <int KT>
__global__ void __launch_bounds__(1024)
k_sweep_ncem_level(int K, const nemk_lpsrc lps, const int32_t *__restrict__ row_ptr,
                   const int32_t *__restrict__ col, const float *__restrict__ wgt, double beta,
                   uint8_t *lab, const int32_t *__restrict__ sites,
                   const int32_t *__restrict__ level_ptr, int lv_lo, int lv_hi, int single_cta,
                   nemk_counters *cnt, const int32_t *__restrict__ skip) {
    if (skip && (skip[0] | skip[1])) return;
    const volatile uint8_t *vlab = lab;
    int changed = 0, nul = 0, ties = 0;
    for (int lv = lv_lo; lv < lv_hi; lv++) {
        int lo = level_ptr[lv], hi = level_ptr[lv + 1];
        int start = single_cta ? threadIdx.x : blockIdx.x * blockDim.x + threadIdx.x;
        int stride = single_cta ? blockDim.x : gridDim.x * blockDim.x;
        for (int s = lo + start; s < hi; s += stride) {
            int i = sites[s];
            double ctx[KT];
            ctx_labels<KT>(K, i, row_ptr, col, wgt, [&](int j) { return (unsigned)vlab[j]; }, ctx);
            int flags;
            double lpv[KT];
            load_lp<KT>(lps, K, (size_t)i, lpv);
            int km = site_argmax<KT>(K, lpv, ctx, beta, flags);
            changed += (km != (int)vlab[i]);
            nul += flags & 1;
            ties += (flags >> 1) & 1;
            lab[i] = (uint8_t)km;
        }
        if (single_cta) __syncthreads();
    }
    if (changed) atomicAdd(&cnt->changed, changed);
    if (nul) atomicAdd(&cnt->allnul, nul);
    if (ties) atomicAdd(&cnt->ties, ties);
}

// ---- nem (fuzzy) site update: t_i = softmax_k(logpf_ik + beta*ctx_ik), ctx from float t.
template <int KT>
static __device__ __forceinline__ void ctx_fuzzy(int K, int i, const int32_t *__restrict__ row_ptr,
                                                 const int32_t *__restrict__ col,
                                                 const float *__restrict__ wgt,
                                                 const volatile float *t, double *ctx) {
#pragma unroll
    for (int k = 0; k < KT; k++) ctx[k] = 0.0;
    if (!row_ptr) return;
    int lo = row_ptr[i], hi = row_ptr[i + 1];
    for (int e = lo; e < hi; e++) {
        const volatile float *tj = t + (size_t)col[e] * K;
        double w = (double)wgt[e];
#pragma unroll
        for (int k = 0; k < KT; k++)
            if (k < K) ctx[k] += w * (double)tj[k];
    }
}

template <int KT>
static __device__ __forceinline__ float site_softmax(int K, const double *__restrict__ lp,
                                                     const double *ctx, double beta,
                                                     const volatile float *t_old, float *t_new,
                                                     int &allnul) {
    double sc[KT], mx = neg_inf();
#pragma unroll
    for (int k = 0; k < KT; k++) {
        sc[k] = (k < K) ? lp[k] + beta * ctx[k] : neg_inf();
        mx = fmax(mx, sc[k]);
    }
    float md = 0.f;
    allnul = (mx == neg_inf());
    double z = 0.0;
#pragma unroll
    for (int k = 0; k < KT; k++) {
        sc[k] = allnul ? 1.0 : exp(sc[k] - mx);  // exp(-inf)=0 for zero-density classes
        if (k < K) z += sc[k];
    }
#pragma unroll
    for (int k = 0; k < KT; k++) {
        if (k < K) {
            float v = (float)(sc[k] / z);
            float df = fabsf(__fsub_rn(v, t_old[k]));  // HasConverged, nem_alg.c:2082-2085
            md = fmaxf(md, df);
            t_new[k] = v;
        }
    }
    return md;
}

static __device__ __forceinline__ void atomic_max_float(float *addr, float v) {
    // v >= 0: integer ordering equals float ordering
    atomicMax((int *)addr, __float_as_int(v));
}

template <int KT>
__global__ void __launch_bounds__(256)
k_sweep_nem_jacobi(int K, int row0, int n_loc, const double *__restrict__ logpf,
                   const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                   const float *__restrict__ wgt, double beta, const float *__restrict__ t_in,
                   float *__restrict__ t_out, nemk_counters *cnt,
                   const int32_t *__restrict__ skip) {
    if (skip && (skip[0] | skip[1])) return;
    int il = blockIdx.x * blockDim.x + threadIdx.x;
    int i = row0 + il;
    float md = 0.f;
    int allnul = 0;
    if (il < n_loc) {
        double ctx[KT];
        ctx_fuzzy<KT>(K, i, beta != 0.0 ? row_ptr : nullptr, col, wgt, t_in, ctx);
        float tn[KT];
        md = site_softmax<KT>(K, logpf + (size_t)il * K, ctx, beta, t_in + (size_t)i * K, tn, allnul);
#pragma unroll
        for (int k = 0; k < KT; k++)
            if (k < K) t_out[(size_t)i * K + k] = tn[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) md = fmaxf(md, __shfl_xor_sync(FULL, md, o));
    unsigned bn = __ballot_sync(FULL, allnul);
    if ((threadIdx.x & 31) == 0) {
        if (md > 0.f) atomic_max_float(&cnt->maxdiff, md);
        if (bn) atomicAdd(&cnt->allnul, __popc(bn));
    }
}

template <int KT>
__global__ void __launch_bounds__(1024)
k_sweep_nem_level(int K, const double *__restrict__ logpf, const int32_t *__restrict__ row_ptr,
                  const int32_t *__restrict__ col, const float *__restrict__ wgt, double beta,
                  float *t, const int32_t *__restrict__ sites, const int32_t *__restrict__ level_ptr,
                  int lv_lo, int lv_hi, int single_cta, nemk_counters *cnt,
                  const int32_t *__restrict__ skip) {
    if (skip && (skip[0] | skip[1])) return;
    float md = 0.f;
    int nul = 0;
    for (int lv = lv_lo; lv < lv_hi; lv++) {
        int lo = level_ptr[lv], hi = level_ptr[lv + 1];
        int start = single_cta ? threadIdx.x : blockIdx.x * blockDim.x + threadIdx.x;
        int stride = single_cta ? blockDim.x : gridDim.x * blockDim.x;
        for (int s = lo + start; s < hi; s += stride) {
            int i = sites[s];
            double ctx[KT];
            ctx_fuzzy<KT>(K, i, row_ptr, col, wgt, t, ctx);
            float tn[KT];
            int an;
            md = fmaxf(md, site_softmax<KT>(K, logpf + (size_t)i * K, ctx, beta,
                                            t + (size_t)i * K, tn, an));
            nul += an;
#pragma unroll
            for (int k = 0; k < KT; k++)
                if (k < K) t[(size_t)i * K + k] = tn[k];
        }
        if (single_cta) __syncthreads();
    }
    if (md > 0.f) atomic_max_float(&cnt->maxdiff, md);
    if (nul) atomicAdd(&cnt->allnul, nul);
}

// =============================================================================================
// M-step sufficient statistics.  EstimSizes / ComputeMedian / EstimLaplaceIner
// (nem_mod.c:1275-1317, 1422-1479, 1646-1704) all reduce to n_k = sum_i t_ik and
// S_kd = sum_i t_ik x_id  (X^T.T): mu_kd and iner_kd are closed forms of (n_k, S_kd).
// =============================================================================================
// ncem: class bit masks cm[k][w] (bit i%32 of word i/32 set iff lab_i == k) + n_k
template <int KT>
__global__ void __launch_bounds__(256)
k_label_masks(int K, int n, int nwt, const uint8_t *__restrict__ lab, uint32_t *__restrict__ cm,
              int32_t *nk, uint8_t *__restrict__ lab_m, const int32_t *__restrict__ halt) {
    __shared__ int snk[KT];
    if (halt && *halt) return;
    if (threadIdx.x < KT) snk[threadIdx.x] = 0;
    __syncthreads();
    int lane = threadIdx.x & 31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nwt * 32; i += gridDim.x * blockDim.x) {
        unsigned l = (i < n) ? lab[i] : 255u;
        if (lab_m && i < n) lab_m[i] = (uint8_t)l;
#pragma unroll
        for (int k = 0; k < KT; k++) {
            if (k < K) {
                unsigned m = __ballot_sync(FULL, l == (unsigned)k);
                if (lane == 0) {
                    cm[(size_t)k * nwt + (i >> 5)] = m;
                    if (m) atomicAdd(&snk[k], __popc(m));
                }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < K && snk[threadIdx.x]) atomicAdd(&nk[threadIdx.x], snk[threadIdx.x]);
}

// ncem: S_kd = sum_w popc(XT[d][w] & cm[k][w]).  A warp owns 32 x MU uint4 of a column (4096 x MU
// families), keeps its class masks in registers and walks a chunk of genomes.  Popcounts go through
// carry-save adders (ones' = ones^a^b, carry = maj(ones,a,b): 2 LOP3 on the full-rate ALU pipe, one
// quarter-rate POPC per TWO words); integer adds => exact and order-free (run-to-run reproducible).
template <int KT, int MU>
__global__ void __launch_bounds__(256)
k_mstep_ncem(int K, int D, int nwt4, const uint4 *__restrict__ xt, const uint4 *__restrict__ cm,
             int dchunk, int32_t *S, const int32_t *__restrict__ halt) {
    if (halt && *halt) return;
    int lane = threadIdx.x & 31;
    int warp_in_block = threadIdx.x >> 5;
    int wg = blockIdx.x * (blockDim.x >> 5) + warp_in_block;  // word group: 32*MU uint4
    int base = wg * 32 * MU;
    if (base >= nwt4) return;
    uint4 m[KT][MU];
#pragma unroll
    for (int u = 0; u < MU; u++) {
        int c = base + u * 32 + lane;
#pragma unroll
        for (int k = 0; k < KT; k++)
            m[k][u] = (c < nwt4 && k < K) ? cm[(size_t)k * nwt4 + c] : make_uint4(0, 0, 0, 0);
    }
    int d0 = blockIdx.y * dchunk, d1 = min(D, d0 + dchunk);
#pragma unroll 2
    for (int dd = d0; dd < d1; dd++) {
        uint4 v[MU];
#pragma unroll
        for (int u = 0; u < MU; u++) {
            int c = base + u * 32 + lane;
            v[u] = (c < nwt4) ? __ldg(xt + (size_t)dd * nwt4 + c) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int k = 0; k < KT; k++) {
            if (k < K) {
                uint32_t ones = 0u, cnt2 = 0u;
#pragma unroll
                for (int u = 0; u < MU; u++) {
                    uint32_t a0 = v[u].x & m[k][u].x, a1 = v[u].y & m[k][u].y;
                    uint32_t a2 = v[u].z & m[k][u].z, a3 = v[u].w & m[k][u].w;
                    uint32_t c0 = lop_maj3(ones, a0, a1);
                    uint32_t o1 = lop_xor3(ones, a0, a1);
                    uint32_t c1 = lop_maj3(o1, a2, a3);
                    ones = lop_xor3(o1, a2, a3);
                    cnt2 += __popc(c0) + __popc(c1);
                }
                int s = (int)(2u * cnt2) + __popc(ones);
                s = __reduce_add_sync(FULL, s);
                if (lane == k && s) atomicAdd(&S[(size_t)k * D + dd], s);
            }
        }
    }
}

// ncem, incremental statistics.  After the first iterations only a few hundred families change
// class per sweep (the centres stop moving), so instead of re-reading X^T
// (N*D/8 bytes) the statistics are UPDATED: every row whose label differs from lab_m (the labels
// S and n currently describe = the input labels of the last sweep, which the sweep leaves intact in
// its other buffer) moves its bits from S[old] to S[new].  Integer adds: exact, and the result
// equals the full recount whatever the order.
__global__ void __launch_bounds__(256)
k_changed_rows(int n, const uint8_t *__restrict__ lab, const uint8_t *__restrict__ lab_m,
               int32_t *list, int32_t *count, const int32_t *__restrict__ halt) {
    if (halt && *halt) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool ch = i < n && lab[i] != lab_m[i];
    unsigned m = __ballot_sync(FULL, ch);
    int lane = threadIdx.x & 31, base = 0;
    if (m && lane == 0) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(FULL, base, 0);
    if (ch) list[base + __popc(m & ((1u << lane) - 1u))] = i;
}

// the same scan, 16 families per thread (one uint4 of labels from each buffer): 1/16 of the CTAs,
// one work-list atomicAdd per 512 families.  Needs 16-byte aligned label pointers (one GPU, or a
// shard that starts on a multiple of 16); the order of the list does not matter (integer adds).
#ifndef CHANGED_ROWS_VEC
#define CHANGED_ROWS_VEC 1
#endif
static __device__ __forceinline__ uint32_t nonzero_bytes(uint32_t x) {   // bit j: byte j of x != 0
    uint32_t m = (x | (x >> 4)) & 0x0f0f0f0fu;
    m = (m | (m >> 2)) & 0x03030303u;
    m = (m | (m >> 1)) & 0x01010101u;
    return (m | (m >> 7) | (m >> 14) | (m >> 21)) & 0xfu;
}
__global__ void __launch_bounds__(256)
k_changed_rows16(int n, const uint8_t *__restrict__ lab, const uint8_t *__restrict__ lab_m,
                 int32_t *list, int32_t *count, const int32_t *__restrict__ halt) {
    if (halt && *halt) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    const int i0 = t * 16;
    uint32_t dm = 0u;   // bit j: family i0 + j changed class
    if (i0 + 16 <= n) {
        const uint4 a = __ldg((const uint4 *)lab + t), b = __ldg((const uint4 *)lab_m + t);
        dm = nonzero_bytes(a.x ^ b.x) | (nonzero_bytes(a.y ^ b.y) << 4) |
             (nonzero_bytes(a.z ^ b.z) << 8) | (nonzero_bytes(a.w ^ b.w) << 12);
    } else {
        for (int j = 0; i0 + j < n; j++) dm |= (uint32_t)(lab[i0 + j] != lab_m[i0 + j]) << j;
    }
    const int c = __popc(dm);
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    if (!total) return;   // warp-uniform
    int base = 0;
    if (lane == 31) base = atomicAdd(count, total);
    base = __shfl_sync(FULL, base, 31);
    int pos = base + incl - c;
    while (dm) {
        list[pos++] = i0 + __ffs(dm) - 1;
        dm &= dm - 1;
    }
}

// work item = (group of 32 changed rows, chunk of DELTA_CHUNK words): lane = row, its DELTA_CHUNK/4
// uint4 loads are issued together; per 32-genome word a ballot transpose gives lane b the 32-row
// column of genome 32w+b, which meets the rows' old/new class masks by popcount.  Items are spread
// over the whole grid, so a few hundred changed rows cost microseconds.
#define DELTA_CHUNK 4
// the items of the incremental update, spread over the warps of the whole grid (shared by the
// standalone kernel and the persistent EM kernel)
template <int KT>
static __device__ __forceinline__ void mstep_delta_items(int K, int D, int wpr, const uint32_t *x,
                                                         const uint8_t *lab, const uint8_t *lab_m,
                                                         const int32_t *list, int total, int32_t *S,
                                                         int32_t *nk) {
    const int lane = threadIdx.x & 31;
    const int groups = (total + 31) >> 5, warps = (gridDim.x * blockDim.x) >> 5;
    const int wreal = (D + 31) >> 5;
    const int chunks = (wreal + DELTA_CHUNK - 1) / DELTA_CHUNK;
    const long long items = (long long)groups * chunks;
    for (long long it = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < items; it += warps) {
        int g = (int)(it / chunks), c = (int)(it % chunks);
        int idx = g * 32 + lane;
        int row = idx < total ? list[idx] : -1;
        unsigned lo = row >= 0 ? lab_m[row] : 255u, ln = row >= 0 ? lab[row] : 255u;
        unsigned plus[KT], minus[KT];
#pragma unroll
        for (int k = 0; k < KT; k++) {
            plus[k] = __ballot_sync(FULL, ln == (unsigned)k);
            minus[k] = __ballot_sync(FULL, lo == (unsigned)k);
        }
        if (c == 0 && lane < K) {
            int dn = 0;
#pragma unroll
            for (int k = 0; k < KT; k++)
                if (lane == k) dn = __popc(plus[k]) - __popc(minus[k]);
            if (dn) atomicAdd(&nk[lane], dn);
        }
        const uint32_t *xr = x + (size_t)(row >= 0 ? row : 0) * wpr + (size_t)c * DELTA_CHUNK;
        uint4 v4[DELTA_CHUNK / 4];
#pragma unroll
        for (int u = 0; u < DELTA_CHUNK / 4; u++)   // wpr is a multiple of 4: whole uint4 stay in the row
            v4[u] = (row >= 0 && c * DELTA_CHUNK + u * 4 < wpr) ? __ldg((const uint4 *)xr + u)
                                                              : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < DELTA_CHUNK / 4; u++) {
            uint32_t vv[4] = {v4[u].x, v4[u].y, v4[u].z, v4[u].w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                int w = c * DELTA_CHUNK + u * 4 + q;
                if (w >= wreal) break;
                uint32_t word = vv[q], mine = 0;
#pragma unroll
                for (int b = 0; b < 32; b++) {
                    uint32_t col = __ballot_sync(FULL, (word >> b) & 1u);
                    if (lane == b) mine = col;
                }
                int d = w * 32 + lane;
                if (d < D) {
#pragma unroll
                    for (int k = 0; k < KT; k++) {
                        if (k < K) {
                            int dv = __popc(mine & plus[k]) - __popc(mine & minus[k]);
                            if (dv) atomicAdd(&S[(size_t)k * D + d], dv);
                        }
                    }
                }
            }
        }
    }
}

template <int KT>
__global__ void __launch_bounds__(256)
k_mstep_delta(int K, int D, int wpr, const uint32_t *__restrict__ x, const uint8_t *__restrict__ lab,
              const uint8_t *__restrict__ lab_m, const int32_t *__restrict__ list,
              const int32_t *__restrict__ count, int32_t *S, int32_t *nk,
              const int32_t *__restrict__ halt) {
    if (halt && *halt) return;
    mstep_delta_items<KT>(K, D, wpr, x, lab, lab_m, list, *count, S, nk);
}

// nem (fuzzy): partial sums S_kd = sum_i t_ik x_id over a chunk of families, fp64, fixed order
// (families ascending inside the chunk, chunks summed in order by k_mstep_nem_reduce).
// Lanes over genomes: a CTA of 128 threads owns 512 genomes (16 words of the packed row); warp w
// takes words 4w..4w+3 and lane b their bit b, so ONE broadcast 16-byte shared-memory load of the
// row's words and K broadcast loads of the posteriors feed 4*K fused multiply-adds per thread.
// The conditional add is fma(2 or 0, t/2, acc): the bit is rotated to position 30 of the high word
// of a double (2.0 or 0.0), t is staged halved, so the value equals acc + t without a select.
template <int KT>
__global__ void __launch_bounds__(128)
k_mstep_nem_partial(int K, int n, int D, int wpr, const uint32_t *__restrict__ x,
                    const float *__restrict__ t, int rows_per_chunk, double *__restrict__ partial_s,
                    double *__restrict__ partial_n) {
    constexpr int TILE = 128;   // families staged per pass
    __shared__ __align__(16) double st[TILE * KT];       // t / 2
    __shared__ __align__(16) uint32_t sx[TILE * 16];
    __shared__ double shn[4 * KT];
    const int chunk = blockIdx.x, db = blockIdx.y;
    const int r0 = chunk * rows_per_chunk, r1 = min(n, r0 + rows_per_chunk);
    const int w0 = db * 16;     // 512 genomes = 16 words
    const int wsel = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rot = (30 - lane) & 31;
    double acc[4][KT], accn[KT];
#pragma unroll
    for (int k = 0; k < KT; k++) {
        accn[k] = 0.0;
#pragma unroll
        for (int q = 0; q < 4; q++) acc[q][k] = 0.0;
    }
    for (int base = r0; base < r1; base += TILE) {
        const int cnt = min(TILE, r1 - base);
        __syncthreads();
        for (int q = threadIdx.x; q < cnt * KT; q += 128) {
            int r = q / KT, k = q - r * KT;
            st[q] = k < K ? 0.5 * (double)t[(size_t)(base + r) * K + k] : 0.0;
        }
        for (int q = threadIdx.x; q < cnt * 16; q += 128) {
            int r = q >> 4, w = w0 + (q & 15);
            sx[q] = (w < wpr) ? x[(size_t)(base + r) * wpr + w] : 0u;
        }
        __syncthreads();
        const uint4 *sx4 = reinterpret_cast<const uint4 *>(sx) + wsel;
#pragma unroll 4
        for (int r = 0; r < cnt; r++) {
            const uint4 xv = sx4[r * 4];
            double td[KT];
#pragma unroll
            for (int k = 0; k < KT; k++) td[k] = st[r * KT + k];
            const uint32_t wv[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const double m = __hiloint2double((int)(__funnelshift_l(wv[q], wv[q], rot) & 0x40000000u), 0);
#pragma unroll
                for (int k = 0; k < KT; k++) acc[q][k] = fma(m, td[k], acc[q][k]);
            }
        }
        if (db == 0) {  // n_k partial: thread handles family base+threadIdx.x
            if (threadIdx.x < cnt) {
#pragma unroll
                for (int k = 0; k < KT; k++)
                    if (k < K) accn[k] += 2.0 * st[threadIdx.x * KT + k];
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int d = (w0 + wsel * 4 + q) * 32 + lane;
        if (d < D) {
#pragma unroll
            for (int k = 0; k < KT; k++)
                if (k < K) partial_s[((size_t)chunk * K + k) * D + d] = acc[q][k];
        }
    }
    if (db == 0) {
        int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
        for (int k = 0; k < KT; k++) {
            double v = accn[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
            if (l == 0) shn[w * KT + k] = v;
        }
        __syncthreads();
        if (threadIdx.x < K) {
            int k = threadIdx.x;
            partial_n[(size_t)chunk * K + k] = ((shn[0 * KT + k] + shn[1 * KT + k]) + shn[2 * KT + k]) + shn[3 * KT + k];
        }
    }
}

// fixed-order reduction over chunks: S[k][d] and n[k]
__global__ void k_mstep_nem_reduce(int K, int D, int nchunks, const double *__restrict__ partial_s,
                                   const double *__restrict__ partial_n, double *__restrict__ S,
                                   double *__restrict__ nk) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < K * D) {
        double s = 0.0;
        for (int c = 0; c < nchunks; c++) s += partial_s[(size_t)c * K * D + q];
        S[q] = s;
    }
    if (q < K) {
        double s = 0.0;
        for (int c = 0; c < nchunks; c++) s += partial_n[(size_t)c * K + q];
        nk[q] = s;
    }
}

// =============================================================================================
// M-step closed forms (one CTA).  EstimLaplaceCenters / ComputeMedian: mu = 1 | 0 | 1/2 for
// S > | < | = n/2 (nem_mod.c:1422-1479);  EstimLaplaceIner: iner = S|1-mu| + (n-S)|mu|
// (nem_mod.c:1674-1683);  InerToDisp{__,K_,_D,KD} (nem_mod.c:965-1174, MISSING_IGNORE branch
// forced for Bernoulli, nem_mod.c:446-447);  proportions (nem_mod.c:455-465).
// theta is float32 like the reference; divisions are float divisions.
// =============================================================================================
// M-step closed forms + E-step tables of the new theta in ONE launch: a thread-block CLUSTER of
// FT_CLUSTER CTAs per class, each CTA owning a slice of the genomes (whole 32-genome words).  The
// sums over genomes (sk_/s__ inertia, the classes' base terms) are block-reduced, left in each
// CTA's shared memory and combined through distributed shared memory in CTA-rank order (fixed
// order => reproducible), with two cluster barriers instead of a second kernel launch.
// The models that pool classes (s_d, s__) recompute the other classes' inertia from S and n.
namespace cg = cooperative_groups;

static __device__ __forceinline__ float iner_of(double s, double n, bool nonempty, float mu_keep) {
    // EstimLaplaceCenters / ComputeMedian then EstimLaplaceIner; an empty class keeps its centre
    // (nem_mod.c:1363) and has n = S = 0, hence zero inertia
    double half = 0.5 * n;
    float mu = nonempty ? (s > half ? 1.0f : (s < half ? 0.0f : 0.5f)) : mu_keep;
    return (float)(s * fabs(1.0 - (double)mu) + (n - s) * fabs((double)mu));
}

// sum of v[0..NV) over the cluster: block sums (the fixed shuffle tree + fixed smem order of
// block_sum, all NV values through ONE pair of block barriers), left in each CTA's `slot`, then
// added in CTA-rank order by the threads that `need` the totals -- the CL remote reads of a value
// are independent loads issued back to back (CL is a compile-time constant), not a chain.
template <int NV, int TH, int CL>
static __device__ __forceinline__ void cluster_sum(cg::cluster_group &cluster, double (&v)[NV],
                                                   double *sh /*[NV * 32]*/, double *slot /*[NV] smem*/,
                                                   bool need) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
    for (int q = 0; q < NV; q++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(FULL, v[q], o);
    }
    __syncthreads();   // sh may still be read from a previous sum
    if (l == 0) {
#pragma unroll
        for (int q = 0; q < NV; q++) sh[q * 32 + w] = v[q];
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int q = 0; q < NV; q++) {
            double r = (l < TH / 32) ? sh[q * 32 + l] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(FULL, r, o);
            if (l == 0) slot[q] = r;
        }
    }
    cluster.sync();
    if (need) {
        double part[NV][CL];
#pragma unroll
        for (int r = 0; r < CL; r++) {
            const double *rs = cluster.map_shared_rank(slot, r);
#pragma unroll
            for (int q = 0; q < NV; q++) part[q][r] = rs[q];
        }
#pragma unroll
        for (int q = 0; q < NV; q++) {
            double t = 0.0;
#pragma unroll
            for (int r = 0; r < CL; r++) t += part[q][r];
            v[q] = t;
        }
    }
    cluster.sync();   // slots may be rewritten
}

// CL CTAs per class (launched as a cluster of CL through cudaLaunchKernelEx), TH threads per CTA
template <int CL, int TH>
__global__ void __launch_bounds__(TH)
k_mstep_finalize_tables(int K, int N, int D, int wpr, int prop_model, int disp_model,
                        const int32_t *__restrict__ s_int, const int32_t *__restrict__ nk_int,
                        const double *__restrict__ s_dbl, const double *__restrict__ nk_dbl,
                        float *prop, float *center, float *disp, nemk_coef *coef,
                        uint32_t *mask_xor, uint32_t *mask_valid, uint32_t *mask_f0,
                        uint32_t *mask_f1, double *delta) {
    __shared__ double sh[5 * 32];
    __shared__ double slot[8];
    __shared__ float nkf[NEMB_MAX_K];
    __shared__ double nkd[NEMB_MAX_K];
    if (coef->halt) return;   // uniform over the launch: only k_iter_end (another launch) writes it
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, k = blockIdx.x / CL, part = blockIdx.x % CL;
    const int wreal = (D + 31) >> 5, wper = (wreal + CL - 1) / CL;
    const int w_lo = min(wreal, part * wper), w_hi = min(wreal, w_lo + wper);
    const int j_lo = w_lo * 32, j_hi = min(D, w_hi * 32);
    if (tid < K) {
        double v = s_int ? (double)nk_int[tid] : nk_dbl[tid];
        nkd[tid] = v;
        nkf[tid] = (float)v;
    }
    __syncthreads();
    if (blockIdx.x == 0 && tid == 0) {
        int empty = 0;
        for (int c = 0; c < K; c++)
            if (!((double)nkf[c] > NEM_EPSILON)) empty = c + 1;  // nem_mod.c:1363,1404-1409
        coef->empty_class = empty;  // like the reference: the LAST empty class, 1-based
    }
    auto S_of = [&](int c, int j) -> double {
        size_t q = (size_t)c * D + j;
        return s_int ? (double)s_int[q] : s_dbl[q];
    };
    const bool nonempty = (double)nkf[k] > NEM_EPSILON;
    // centres of this class (this CTA's genomes)
    if (nonempty)
        for (int j = j_lo + tid; j < j_hi; j += TH) {
            double s = S_of(k, j), half = 0.5 * nkd[k];
            center[(size_t)k * D + j] = s > half ? 1.0f : (s < half ? 0.0f : 0.5f);
        }
    // dispersions of this class
    if (disp_model == 3) {  // skd: nem_mod.c:1152-1170
        if (nonempty)
            for (int j = j_lo + tid; j < j_hi; j += TH)
                disp[(size_t)k * D + j] = __fdiv_rn(iner_of(S_of(k, j), nkd[k], true, 0.f), nkf[k]);
    } else if (disp_model == 2) {  // s_d: nem_mod.c:1104-1126, float sums over the classes in order
        for (int j = j_lo + tid; j < j_hi; j += TH) {
            float si = 0.f, sn = 0.f;
            for (int c = 0; c < K; c++) {
                bool ne = (double)nkf[c] > NEM_EPSILON;
                sn = __fadd_rn(sn, nkf[c]);
                si = __fadd_rn(si, iner_of(S_of(c, j), nkd[c], ne, ne ? 0.f : center[(size_t)c * D + j]));
            }
            disp[(size_t)k * D + j] = __fdiv_rn(si, sn);
        }
    } else {  // sk_ (nem_mod.c:1043-1073) and s__ (nem_mod.c:988-1015): inertia summed over genomes
        double v[1] = {0.0};
        double sn = 0.0;
        if (disp_model == 1) {
            for (int j = j_lo + tid; j < j_hi; j += TH)
                v[0] += (double)iner_of(S_of(k, j), nkd[k], nonempty, nonempty ? 0.f : center[(size_t)k * D + j]);
            sn = (double)nkf[k] * (double)D;
        } else {
            for (int c = 0; c < K; c++) {
                if (nkf[c] > 0.f) {
                    bool ne = (double)nkf[c] > NEM_EPSILON;
                    for (int j = j_lo + tid; j < j_hi; j += TH)
                        v[0] += (double)iner_of(S_of(c, j), nkd[c], ne, ne ? 0.f : center[(size_t)c * D + j]);
                    sn += (double)nkf[c] * (double)D;
                }
            }
        }
        cluster_sum<1, TH, CL>(cluster, v, sh, slot, true);
        if (disp_model == 0 || nkf[k] > 0.f) {
            float dk = __fdiv_rn((float)v[0], (float)sn);
            for (int j = j_lo + tid; j < j_hi; j += TH) disp[(size_t)k * D + j] = dk;
        }
    }
    if (part == 0 && tid == 0)  // nem_mod.c:456-465
        prop[k] = prop_model == 1 ? __fdiv_rn(nkf[k], (float)N) : (float)(1.0 / (double)K);
    // the class's first dispersion (the popcount path's reference value) is written by part 0
    __threadfence();
    cluster.sync();
    ClassCoef cc = class_coef(__ldcg(&disp[(size_t)k * D]));
    TablesPartial p = tables_words(k, D, wpr, w_lo, w_hi, cc, center, disp, mask_xor, mask_valid,
                                   mask_f0, mask_f1, delta);
    // padding words [wreal, wpr) of the masks carry no genome: part 0 keeps them clean
    if (part == 0)
        for (int w = wreal + tid; w < wpr; w += TH) {
            size_t o = (size_t)k * wpr + w;
            mask_xor[o] = 0u; mask_valid[o] = 0u; mask_f0[o] = 0u; mask_f1[o] = 0u;
        }
    if (p.mu_moved) atomicOr(&coef->mu_changed, 1);   // preset by the launcher (0, or 1 = forced)
    double v[5] = {p.base_u, p.base_g, (double)p.notok, (double)p.n_valid, (double)p.n_x1};
    cluster_sum<5, TH, CL>(cluster, v, sh, slot, part == 0 && tid == 0);
    if (part == 0 && tid == 0)
        tables_commit(k, K, D, prop, coef, delta, cc, v[0], v[1], v[2] == 0.0, (int)v[3], (int)v[4], true);
}

// =============================================================================================
// Criteria.  ComputeCrit (nem_alg.c:2678-2757): D, G, L, Z per family then
// U = D + beta/2 G, M = D + beta G + Z; float64, log-domain L and Z, deterministic two-stage sum.
// =============================================================================================
template <int KT>
static __device__ __forceinline__ void crit_site(int K, const double *__restrict__ lp,
                                                 const double *ctx, const float *ti, double beta,
                                                 double &cD, double &cG, double &cL, double &cZ) {
    double lmx = neg_inf(), zmx = neg_inf();
#pragma unroll
    for (int k = 0; k < KT; k++) {
        if (k < K) { lmx = fmax(lmx, lp[k]); zmx = fmax(zmx, beta * ctx[k]); }
    }
    double fs = 0, zs = 0;
#pragma unroll
    for (int k = 0; k < KT; k++) {
        if (k < K) {
            double l = lp[k];
            float cik = ti[k];
            if (cik > FLT_MIN) {  // MINFLOAT, nem_alg.c:2727
                double lc = (l == neg_inf()) ? -(double)FLT_MAX : l;  // nem_mod.c:685
                cD += (double)cik * (lc - log((double)cik));
                cG += (double)cik * ctx[k];
            }
            if (lmx > neg_inf()) fs += exp(l - lmx);
            zs += exp(beta * ctx[k] - zmx);
        }
    }
    cL += (lmx > neg_inf()) ? lmx + log(fs) : neg_inf();
    cZ -= zmx + log(zs);
}

// EstimBeta's per-site terms (nem_alg.c:2163-2191): log pseudo-likelihood of the classification
// under the Potts prior, its first derivative in beta and minus its second derivative.  The
// softmax moments of c_ik = ctx[k] are taken relative to max_k(beta c_ik): the reference's values
// wherever its float exp() does not overflow.
template <int KT>
static __device__ __forceinline__ void grad_site(int K, const double *ctx, const float *ti, double beta,
                                                 double &crit, double &grad, double &dsec) {
    double mx = neg_inf();
#pragma unroll
    for (int k = 0; k < KT; k++)
        if (k < K) mx = fmax(mx, beta * ctx[k]);
    double se = 0, sce = 0, sc2e = 0, stc = 0;
#pragma unroll
    for (int k = 0; k < KT; k++) {
        if (k < K) {
            double e = exp(beta * ctx[k] - mx);
            se += e; sce += ctx[k] * e; sc2e += ctx[k] * ctx[k] * e;
            stc += (double)ti[k] * ctx[k];
        }
    }
    crit += beta * stc - (mx + log(se));
    grad += stc - sce / se;
    dsec += (sc2e * se - sce * sce) / (se * se);
}

// blocks [0, heavy_blocks): hubs of the (index-sorted) heavy list, one warp per site;
// the other blocks: the remaining sites, one thread per site.  Fixed assignment => deterministic.
// GRAD: the same walk accumulates EstimBeta's (crit, grad, dsec) instead of (D, G, L, Z).
template <int KT, bool GRAD>
__global__ void __launch_bounds__(256)
k_criteria_partial(int K, int row0, int n_loc, const nemk_lpsrc lps,
                   const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                   const float *__restrict__ wgt, double beta, const uint8_t *__restrict__ lab,
                   const float *__restrict__ t, const int32_t *__restrict__ heavy, int n_heavy,
                   int heavy_blocks, double *__restrict__ partials) {
    __shared__ double sh[32];
    double cD = 0, cG = 0, cL = 0, cZ = 0;
    if ((int)blockIdx.x < heavy_blocks) {
        const int lane = threadIdx.x & 31;
        int warps = heavy_blocks * (blockDim.x >> 5);
        for (int wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi < n_heavy; wi += warps) {
            int i = heavy[wi];
            double ctx[KT];
            float ti[KT];
            ctx_labels_warp<KT>(K, i, row_ptr, col, wgt, [&](int j) { return (unsigned)lab[j]; }, ctx,
                                lps.wsum_any_order != 0);
            unsigned l = lab[i];
#pragma unroll
            for (int k = 0; k < KT; k++) ti[k] = (l == (unsigned)k) ? 1.f : 0.f;
            if (lane == 0) {
                if (GRAD) grad_site<KT>(K, ctx, ti, beta, cD, cG, cL);
                else {
                    double lpv[KT];
                    load_lp<KT>(lps, K, (size_t)(i - row0), lpv);
                    crit_site<KT>(K, lpv, ctx, ti, beta, cD, cG, cL, cZ);
                }
            }
        }
    } else {
        __shared__ float s_w[8][COOP_CHUNK];
        __shared__ uint8_t s_l[8][COOP_CHUNK];
        const int nb = gridDim.x - heavy_blocks, lane = threadIdx.x & 31;
        for (int base = (blockIdx.x - heavy_blocks) * blockDim.x + (threadIdx.x & ~31); base < n_loc;
             base += nb * blockDim.x) {
            const int il = base + lane, i = row0 + il;
            const bool valid = il < n_loc;
            int lo = 0, hi = 0;
            if (row_ptr && valid) { lo = row_ptr[i]; hi = row_ptr[i + 1]; }
            const bool is_heavy = heavy_blocks && (hi - lo > HEAVY_DEG);
            double ctx[KT];
            float ti[KT];
            if (lab) {
                if (row_ptr) {
                    int seg_lo = __shfl_sync(FULL, lo, 0), seg_hi = __reduce_max_sync(FULL, hi);
                    if (is_heavy) lo = hi = 0;
                    ctx_labels_coop<KT>(lo, hi, seg_lo, seg_hi, col, wgt, lab, s_w[threadIdx.x >> 5],
                                        s_l[threadIdx.x >> 5], ctx);
                } else {
#pragma unroll
                    for (int k = 0; k < KT; k++) ctx[k] = 0.0;
                }
                if (!valid || is_heavy) continue;
                unsigned l = lab[i];
#pragma unroll
                for (int k = 0; k < KT; k++) ti[k] = (l == (unsigned)k) ? 1.f : 0.f;
            } else {
                if (!valid || is_heavy) continue;
                ctx_fuzzy<KT>(K, i, row_ptr, col, wgt, t, ctx);
#pragma unroll
                for (int k = 0; k < KT; k++) ti[k] = (k < K) ? t[(size_t)i * K + k] : 0.f;
            }
            if (GRAD) grad_site<KT>(K, ctx, ti, beta, cD, cG, cL);
            else {
                double lpv[KT];
                load_lp<KT>(lps, K, (size_t)il, lpv);
                crit_site<KT>(K, lpv, ctx, ti, beta, cD, cG, cL, cZ);
            }
        }
    }
    cD = block_sum<256>(cD, sh);
    cG = block_sum<256>(cG, sh);
    cL = block_sum<256>(cL, sh);
    cZ = block_sum<256>(cZ, sh);
    if (threadIdx.x == 0) {
        partials[blockIdx.x * 4 + 0] = cD; partials[blockIdx.x * 4 + 1] = cG;
        partials[blockIdx.x * 4 + 2] = cL; partials[blockIdx.x * 4 + 3] = cZ;
    }
}

__global__ void __launch_bounds__(256)
k_criteria_final(int nblocks, const double *__restrict__ partials, double beta, double *crit6) {
    __shared__ double sh[32];
    double v[4] = {0, 0, 0, 0};
    for (int b = threadIdx.x; b < nblocks; b += 256)   // fixed assignment => deterministic
#pragma unroll
        for (int q = 0; q < 4; q++) v[q] += partials[b * 4 + q];
#pragma unroll
    for (int q = 0; q < 4; q++) v[q] = block_sum<256>(v[q], sh);
    if (threadIdx.x == 0 && beta != beta) {   // NaN beta = raw sums (EstimBeta's crit, grad, dsec)
#pragma unroll
        for (int q = 0; q < 4; q++) crit6[q] = v[q];
    } else if (threadIdx.x == 0) {
        double D = v[0], G = v[1], L = v[2], Z = v[3];
        crit6[0] = D + 0.5 * beta * G;
        crit6[1] = D; crit6[2] = L;
        crit6[3] = D + beta * G + Z;
        crit6[4] = Z; crit6[5] = G;
    }
}

// out[q] = sum over ranks of stage[r][q], ranks added in order 0..W-1 (deterministic; used to build
// all-reduces of the M-step statistics on top of a single all-gather)
__global__ void k_sum_ranks_i32(int world, size_t count, const int32_t *__restrict__ stage,
                                int32_t *__restrict__ out) {
    size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= count) return;
    int32_t s = 0;
    for (int r = 0; r < world; r++) s += stage[(size_t)r * count + q];
    out[q] = s;
}
__global__ void k_sum_ranks_f64(int world, size_t count, const double *__restrict__ stage,
                                double *__restrict__ out) {
    size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= count) return;
    double s = 0.0;
    for (int r = 0; r < world; r++) s += stage[(size_t)r * count + q];
    out[q] = s;
}

// =============================================================================================
// Loader: index-sorted list of the hubs (degree > HEAVY_DEG) among this rank's rows -- a
// deterministic 3-kernel compaction (count per block, scan of the block counts, fill).
// =============================================================================================
#define HL_THREADS 1024
__global__ void __launch_bounds__(HL_THREADS)
k_heavy_count(int row0, int n_loc, const int32_t *__restrict__ row_ptr, int32_t *block_counts) {
    int il = blockIdx.x * HL_THREADS + threadIdx.x;
    int i = row0 + il;
    int hv = (il < n_loc) && (row_ptr[i + 1] - row_ptr[i] > HEAVY_DEG);
    int c = __syncthreads_count(hv);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}
__global__ void __launch_bounds__(HL_THREADS)
k_heavy_scan(int nblocks, int32_t *block_counts, int32_t *total) {
    __shared__ int wsum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int b0 = 0; b0 < nblocks; b0 += HL_THREADS) {
        int b = b0 + threadIdx.x;
        int v = b < nblocks ? block_counts[b] : 0, x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(FULL, x, o); if (lane >= o) x += y; }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        if (w == 0) {
            int ws = wsum[lane], z = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(FULL, z, o); if (lane >= o) z += y; }
            wsum[lane] = z - ws;   // exclusive prefix of the warp totals
        }
        __syncthreads();
        int excl = carry + wsum[w] + x - v;
        if (b < nblocks) block_counts[b] = excl;
        __syncthreads();
        if (threadIdx.x == HL_THREADS - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(HL_THREADS)
k_heavy_fill(int row0, int n_loc, const int32_t *__restrict__ row_ptr,
             const int32_t *__restrict__ block_offsets, int32_t *list) {
    __shared__ int wsum[32];
    int il = blockIdx.x * HL_THREADS + threadIdx.x;
    int i = row0 + il;
    int hv = (il < n_loc) && (row_ptr[i + 1] - row_ptr[i] > HEAVY_DEG);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned m = __ballot_sync(FULL, hv);
    if (lane == 0) wsum[w] = __popc(m);
    __syncthreads();
    if (w == 0) {
        int ws = wsum[lane], z = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(FULL, z, o); if (lane >= o) z += y; }
        wsum[lane] = z - ws;
    }
    __syncthreads();
    if (hv) list[block_offsets[blockIdx.x] + wsum[w] + __popc(m & ((1u << lane) - 1u))] = i;
}

// =============================================================================================
// Loader: graph validation on the device (replaces the checks of ReadPtsNeighs,
// nem_exe.c:1342-1478, and decides whether the reader lists equal the neighbour lists).
// flags[0] |= 1 row_ptr not monotone, |= 2 neighbour out of range, |= 4 some edge i->j has no j->i;
// flags[1] = max degree.
// =============================================================================================
__global__ void __launch_bounds__(256)
k_graph_check(int n, int nnz, const int32_t *__restrict__ row_ptr,
              const int32_t *__restrict__ col, const float *__restrict__ wgt, int32_t *flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = row_ptr[i], hi = row_ptr[i + 1];
    int bad = 0;
    if (hi < lo || lo < 0 || hi > nnz) { atomicOr(&flags[0], 1); return; }
    for (int e = lo; e < hi; e++) {
        int j = col[e];
        float w = wgt[e];
        if (!(w == rintf(w) && fabsf(w) <= 1048576.0f)) bad |= 8;   // sums of < 2^31 such terms are exact
        if (j < 0 || j >= n) { bad |= 2; continue; }
        if (j == i) continue;
        int jl = row_ptr[j], jh = row_ptr[j + 1];
        if (jl < 0 || jh > nnz) continue;   // reported by row j itself
        bool found = false;
        for (int f = jl; f < jh && !found; f++) found = (col[f] == i);
        if (!found) bad |= 4;
    }
    if (bad) atomicOr(&flags[0], bad);
    atomicMax(&flags[1], hi - lo);
}

// =============================================================================================
// End of a sweep / an EM iteration: counters summed over the ranks, the `clas` convergence test
// (HasConverged, nem_alg.c:2075-2089) decided on the device, coef->halt raised when the fit is over,
// and the whole status block published into mapped pinned host memory.  The host polls `seq`: no
// memcpy, no stream synchronisation, and it may already have enqueued the next iteration (whose
// kernels all return at once when halt is set).
// =============================================================================================
static __device__ void iter_end_body(int world, const nemk_counters *cnt_all,
                                     const nemk_iter_status *st, nemk_coef *coef, int decide,
                                     int ncem, int conv, float thr, nemk_host_status *host,
                                     unsigned long long seq) {
    nemk_counters tot;
    tot.changed = 0; tot.nfix = 0; tot.allnul = 0; tot.ties = 0; tot.maxdiff = 0.f; tot.pending = 0;
    tot.changed_glob = 0; tot.kept = 0;
    const volatile nemk_counters *vc = cnt_all;   // written by atomics of this very launch when fused
    if (world == 0) {   // row-sharded speculative sweep: changed / pending are already global
        tot.changed = vc[0].changed_glob; tot.nfix = vc[0].nfix; tot.allnul = vc[0].allnul;
        tot.ties = vc[0].ties; tot.maxdiff = vc[0].maxdiff; tot.pending = vc[0].pending;
        tot.changed_glob = vc[0].changed_glob; tot.kept = vc[0].kept;
    }
    for (int r = 0; r < world; r++) {
        tot.changed += vc[r].changed; tot.allnul += vc[r].allnul; tot.ties += vc[r].ties;
        tot.pending += vc[r].pending; tot.kept += vc[r].kept;
        tot.nfix = max(tot.nfix, vc[r].nfix);
        tot.maxdiff = fmaxf(tot.maxdiff, vc[r].maxdiff);
    }
    int halt = coef->halt, empty = coef->empty_class;
    if (decide && !halt && tot.pending == 0) {   // pending != 0: a row-sharded sweep is not settled yet
        bool converged = false;
        if (conv == 1) {
            float md = ncem ? (tot.changed ? 1.0f : 0.0f) : tot.maxdiff;
            converged = md < thr;
        }
        if (converged || empty) { halt = 1; coef->halt = 1; }
    }
    host->cnt = tot;
#pragma unroll
    for (int q = 0; q < 6; q++) { host->crit_before[q] = st->crit_before[q]; host->crit_after[q] = st->crit_after[q]; }
    host->empty_class = empty; host->mu_changed = coef->mu_changed; host->halt = halt; host->pad = 0;
    __threadfence_system();
    *(volatile unsigned long long *)&host->seq = seq;
}

__global__ void k_iter_end(int world, const nemk_counters *__restrict__ cnt_all,
                           const nemk_iter_status *__restrict__ st, nemk_coef *coef, int decide,
                           int ncem, int conv, float thr, nemk_host_status *host,
                           unsigned long long seq) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    iter_end_body(world, cnt_all, st, coef, decide, ncem, conv, thr, host, seq);
}

// =============================================================================================
// small helpers
// =============================================================================================
__global__ void k_labels_to_t(int K, int n, const uint8_t *__restrict__ lab, float *__restrict__ t) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n * K) t[q] = (lab[q / K] == (unsigned)(q % K)) ? 1.f : 0.f;
}
__global__ void k_t_to_labels(int K, int n, const float *__restrict__ t, uint8_t *__restrict__ lab) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int km = 0;
    float mx = t[(size_t)i * K];
    for (int k = 1; k < K; k++) {
        float v = t[(size_t)i * K + k];
        if (v > mx) { mx = v; km = k; }
    }
    lab[i] = (mx > 0.f) ? (uint8_t)km : (uint8_t)255;  // all-zero row = unlabelled (calloc'd ClassifM)
}

#include "nem_persist.cuh"

// =============================================================================================
// launch layer
// =============================================================================================
#ifdef NEMK_DEV_K3   /* development builds: K = 3 only (compile time) */
#define DISPATCH_K(K, CALL) do { constexpr int KT = 3; CALL; } while (0)
#else
#define DISPATCH_K(K, CALL)                      \
    do {                                         \
        if ((K) <= 2) { constexpr int KT = 2; CALL; }        \
        else if ((K) == 3) { constexpr int KT = 3; CALL; }   \
        else if ((K) == 4) { constexpr int KT = 4; CALL; }   \
        else if ((K) <= 8) { constexpr int KT = 8; CALL; }   \
        else { constexpr int KT = 16; CALL; }                \
    } while (0)
#endif

static int g_num_sms = 0;
static int num_sms() {
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

extern "C" void nemk_pack_u8(nemk_stream s, const uint8_t *x, int n, int d, int wpr, uint32_t *out) {
    if (n <= 0) return;
    k_pack_u8<<<cdiv((long long)n * 32, 256), 256, 0, S(s)>>>(x, n, d, wpr, out);
    note_launch();
}

extern "C" void nemk_transpose_bits(nemk_stream s, const uint32_t *x, int n, int wpr, int d,
                                    int nwt, uint32_t *xt) {
    cudaMemsetAsync(xt, 0, (size_t)d * nwt * sizeof(uint32_t), S(s));
    if (n <= 0) return;
    dim3 grid(cdiv(n, TB_ROWS), cdiv(wpr, 32));
    k_transpose_bits<<<grid, 256, 0, S(s)>>>(x, 0, n, wpr, d, nwt, xt);
    note_launch();
}

// rows [row_base, row_base + rows) only (row_base a multiple of 256); xt must have been zeroed
extern "C" void nemk_transpose_bits_rows(nemk_stream s, const uint32_t *x, int row_base, int rows,
                                         int wpr, int d, int nwt, uint32_t *xt) {
    if (rows <= 0) return;
    dim3 grid(cdiv(rows, TB_ROWS), cdiv(wpr, 32));
    k_transpose_bits<<<grid, 256, 0, S(s)>>>(x, row_base, row_base + rows, wpr, d, nwt, xt);
    note_launch();
}

extern "C" void nemk_theta_tables(nemk_stream s, int k, int d, int wpr, const float *prop,
                                  const float *center, const float *disp, nemk_coef *coef,
                                  uint32_t *mask_xor, uint32_t *mask_valid, uint32_t *mask_f0,
                                  uint32_t *mask_f1, double *delta, int force_mu_changed) {
    cudaMemsetAsync(&coef->uniform_ok, 1, sizeof(int32_t), S(s));
    cudaMemsetAsync(&coef->mu_changed, force_mu_changed ? 1 : 0, sizeof(int32_t), S(s));
    k_theta_tables<<<k, TT_THREADS, 0, S(s)>>>(k, d, wpr, prop, center, disp, coef, mask_xor,
                                              mask_valid, mask_f0, mask_f1, delta);
    note_launch();
}

template <int KT>
static void launch_density_uniform(cudaStream_t st, int K, const uint32_t *x, int n, int wpr,
                                   const nemk_coef *coef, const uint32_t *mx, const uint32_t *mv,
                                   double *logpf, int32_t *hamming, int cached) {
    int wpr4 = wpr / 4;
    size_t smem = (size_t)2 * KT * wpr4 * sizeof(uint4);
    int lpr = 1;
    while (lpr < wpr4 && lpr < 32) lpr <<= 1;
    int rows_per_block = 256 / lpr;
    int grid = cdiv(n, rows_per_block);
    int cap = num_sms() * 16;
    if (grid > cap) grid = cap;
#define LAUNCH_DU(L)                                                                              \
    do {                                                                                          \
        if (smem > 48 * 1024)                                                                     \
            cudaFuncSetAttribute(k_density_uniform<KT, L>,                                        \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
        k_density_uniform<KT, L><<<grid, 256, smem, st>>>(K, (const uint4 *)x, n, wpr4, coef,     \
                                                          (const uint4 *)mx, (const uint4 *)mv,   \
                                                          logpf, hamming, cached);                \
    } while (0)
    switch (lpr) {
    case 1: LAUNCH_DU(1); break;
    case 2: LAUNCH_DU(2); break;
    case 4: LAUNCH_DU(4); break;
    case 8: LAUNCH_DU(8); break;
    case 16: LAUNCH_DU(16); break;
    default: LAUNCH_DU(32); break;
    }
#undef LAUNCH_DU
}

template <int KT, int T, int ROWS>
static bool launch_density_tma_t(cudaStream_t st, int K, int D, const uint32_t *x, int n, int wpr,
                                 const nemk_coef *coef, const uint32_t *mx, const uint32_t *mv,
                                 double *logpf, int32_t *hamming, int cached) {
    int wpr4 = wpr / 4;
    int stride4 = wpr4 | 1;  // odd number of uint4 per shared row => conflict-free LDS.128
    size_t mask_b = (size_t)2 * KT * wpr4 * sizeof(uint4);
    size_t tile_b = (size_t)ROWS * stride4 * sizeof(uint4);
    size_t stat = (size_t)3 * ROWS * (KT + 1) * 4 + 512;
    const size_t budget = 224 * 1024;
    if (mask_b + 2 * tile_b + stat > budget) return false;
    static int force_s = -1, force_c = -1;
    if (force_s < 0) { const char *e = getenv("NEM_B200_DENSITY_STAGES"); force_s = e ? atoi(e) : 0; }
    if (force_c < 0) { const char *e = getenv("NEM_B200_DENSITY_CTAS"); force_c = e ? atoi(e) : 0; }
    // CTAs per SM: enough threads for the ALUs (>= 1024 threads/SM when they fit), and the rest of
    // the shared memory goes to pipeline depth
    int by_threads = 2048 / (ROWS * T);
    int per_sm = force_c ? force_c : (1024 + ROWS * T - 1) / (ROWS * T);
    if (per_sm > by_threads) per_sm = by_threads;
    if (per_sm < 1) per_sm = 1;
    while (per_sm > 1 && (mask_b + 2 * tile_b + stat + 1024) * per_sm > budget) per_sm--;
    int n_stages = (int)((budget / per_sm - mask_b - stat - 1024) / tile_b);
    if (force_s) n_stages = force_s;
    if (n_stages > 16) n_stages = 16;
    if (n_stages < 2) n_stages = 2;
    size_t smem = mask_b + (size_t)n_stages * tile_b;
    if (smem + stat > budget) return false;
    int n_tiles = cdiv(n, ROWS);
    static size_t attr_set = 0;
    if (smem > attr_set) {
        cudaFuncSetAttribute(k_density_tma<KT, ROWS, T>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(budget - stat));
        attr_set = budget;
    }
    int grid = num_sms() * per_sm;
    if (grid > n_tiles) grid = n_tiles;
    k_density_tma<KT, ROWS, T><<<grid, ROWS * T, smem, st>>>(
        K, D, (const uint4 *)x, n, wpr4, stride4, n_tiles, n_stages, coef, (const uint4 *)mx,
        (const uint4 *)mv, logpf, hamming, cached);
    return true;
}

template <int KT>
static bool launch_density_tma(cudaStream_t st, int K, int D, const uint32_t *x, int n, int wpr,
                               const nemk_coef *coef, const uint32_t *mx, const uint32_t *mv,
                               double *logpf, int32_t *hamming, int cached) {
    int wpr4 = wpr / 4;
    static int force_t = -1;
    if (force_t < 0) { const char *e = getenv("NEM_B200_DENSITY_T"); force_t = e ? atoi(e) : 0; }
    int t = force_t ? force_t : (wpr4 >= 32 ? 4 : wpr4 >= 16 ? 2 : 1);   // measured on B200, DESIGN.md
    static int force_r = -1;
    if (force_r < 0) { const char *e = getenv("NEM_B200_DENSITY_ROWS"); force_r = e ? atoi(e) : 0; }
    // two CTAs of 64 families x T column groups per SM: their tile barriers interleave (C4: 121 us
    // against 128 us for one CTA of 128 families)
    int rows = force_r ? force_r : (t >= 4 ? 64 : 128);
    if (KT > 4) t = t > 2 ? 2 : t;   // keep the static partial buffer small for large K
#define DT(TT, RR) launch_density_tma_t<KT, TT, RR>(st, K, D, x, n, wpr, coef, mx, mv, logpf, hamming, cached)
    if constexpr (KT <= 4) {
        if (t >= 8) return rows <= 32 ? DT(8, 32) : rows <= 64 ? DT(8, 64) : DT(8, 128);
        if (t >= 4) return rows <= 32 ? DT(4, 32) : rows <= 64 ? DT(4, 64) : DT(4, 128);
    }
    if (t >= 2) return rows <= 64 ? DT(2, 64) : DT(2, 128);
    return DT(1, 128);
#undef DT
}

static int g_density_impl = -1;  // 0 = v1 (plain loads), 1 = v2 (TMA tiles); env NEM_B200_DENSITY=v1
// logpf from the cached Hamming counts (the epilogue of the density kernels, same bits); runs
// when the class bit masks did not move since the counts were taken
template <int KT>
__global__ void __launch_bounds__(256)
k_logpf_from_h(int K, int n, const nemk_coef *__restrict__ coef, const int32_t *__restrict__ hamming,
               double *__restrict__ logpf) {
    if (coef->empty_class | coef->halt | coef->mu_changed) return;
    size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (size_t)n * K) return;
    int k = (int)(q % K);
    logpf[q] = logpf_of_h(coef, k, hamming[q]);
}

extern "C" void nemk_density_uniform(nemk_stream s, int k, int d, const uint32_t *x, int n, int wpr,
                                     const nemk_coef *coef, const uint32_t *mask_xor,
                                     const uint32_t *mask_valid, double *logpf, int32_t *hamming,
                                     int cached) {
    if (n <= 0) return;
    if (!hamming) cached = 0;
    if (g_density_impl < 0) {
        const char *e = getenv("NEM_B200_DENSITY");
        g_density_impl = (e && !strcmp(e, "v1")) ? 0 : 1;
    }
    bool done = false;
    if (g_density_impl >= 1)
        DISPATCH_K(k, (done = launch_density_tma<KT>(S(s), k, d, x, n, wpr, coef, mask_xor, mask_valid,
                                                     logpf, hamming, cached)));
    if (!done)
        DISPATCH_K(k, (launch_density_uniform<KT>(S(s), k, x, n, wpr, coef, mask_xor, mask_valid,
                                                  logpf, hamming, cached)));
    note_launch();
}

extern "C" void nemk_row_popcount(nemk_stream s, const uint32_t *x, int n, int wpr, int32_t *pop) {
    if (n <= 0) return;
    k_row_popcount<<<cdiv((long long)n * 32, 256), 256, 0, S(s)>>>((const uint4 *)x, n, wpr / 4, pop);
    note_launch();
}

extern "C" void nemk_ham_from_pop(nemk_stream s, int k, int n, int d, const nemk_coef *coef,
                                  const int32_t *pop, int32_t *ham) {
    if (n <= 0) return;
    DISPATCH_K(k, (k_ham_from_pop<KT><<<cdiv(n, 256), 256, 0, S(s)>>>(k, n, d, coef, pop, ham)));
    note_launch();
}

extern "C" void nemk_logpf_from_cache(nemk_stream s, int k, int n, const nemk_coef *coef,
                                      const int32_t *hamming, double *logpf) {
    if (n <= 0) return;
    DISPATCH_K(k, (k_logpf_from_h<KT><<<cdiv((long long)n * k, 256), 256, 0, S(s)>>>(
                      k, n, coef, hamming, logpf)));
    note_launch();
}

// launches the tiled general density when its delta table fits shared memory; false = not launched
template <int KT>
static bool density_general_tiled_k(nemk_stream s, int k, const uint32_t *x, int n, int d, int wpr,
                                    const nemk_coef *coef, const uint32_t *mask_f0,
                                    const uint32_t *mask_f1, const double *delta,
                                    const double *base_g, double *logpf) {
    static int smem_max = -1, warp_only = -1;
    if (warp_only < 0) { const char *e = getenv("NEM_B200_DENSITY_WARP"); warp_only = e && *e; }
    if (warp_only) return false;
    if (smem_max < 0) {
        int dev = 0, v = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        smem_max = v;
    }
    const size_t smem = (size_t)((d + 31) / 32) * 32 * KT * sizeof(double);
    if (smem + 1024 > (size_t)smem_max) return false;
    static size_t attr_set = 0;
    if (smem > attr_set) {
        if (cudaFuncSetAttribute(k_density_general_tiled<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess) { cudaGetLastError(); return false; }
        attr_set = smem;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_density_general_tiled<KT>, 256, smem)
            != cudaSuccess || per_sm < 1) { cudaGetLastError(); return false; }
    if (per_sm > 2) per_sm = 2;
    constexpr int R = DgR<KT>::R;
    long long ntiles = ((long long)n + 32 * R - 1) / (32 * R);
    long long grid = (ntiles + 7) / 8;
    if (grid > (long long)num_sms() * per_sm) grid = (long long)num_sms() * per_sm;
    k_density_general_tiled<KT><<<(int)grid, 256, smem, S(s)>>>(k, x, n, d, wpr, coef, mask_f0, mask_f1,
                                                               delta, base_g, logpf);
    return true;
}
static bool density_general_tiled(nemk_stream s, int k, const uint32_t *x, int n, int d, int wpr,
                                  const nemk_coef *coef, const uint32_t *mask_f0,
                                  const uint32_t *mask_f1, const double *delta,
                                  const double *base_g, double *logpf) {
    bool ok = false;
    DISPATCH_K(k, (ok = density_general_tiled_k<KT>(s, k, x, n, d, wpr, coef, mask_f0, mask_f1, delta,
                                                    base_g, logpf)));
    return ok;
}

extern "C" void nemk_density_general(nemk_stream s, int k, const uint32_t *x, int n, int d, int wpr,
                                     const nemk_coef *coef, const uint32_t *mask_f0,
                                     const uint32_t *mask_f1, const double *delta, double *logpf) {
    if (n <= 0) return;
    // delta[K*D] is followed by base_g[K] (written by k_theta_tables)
    const double *base_g = delta + (size_t)k * d;
    if (density_general_tiled(s, k, x, n, d, wpr, coef, mask_f0, mask_f1, delta, base_g, logpf)) {
        note_launch();
        return;
    }
    // delta does not fit shared memory (or NEM_B200_DENSITY_WARP is set): warp-per-family walk
    int grid = cdiv((long long)n * 32, 256);
    int cap = num_sms() * 16;
    if (grid > cap) grid = cap;
    DISPATCH_K(k, (k_density_general<KT><<<grid, 256, 0, S(s)>>>(k, x, n, d, wpr, coef, nullptr,
                                                                mask_f0, mask_f1, delta, base_g,
                                                                logpf)));
    note_launch();
}

extern "C" void nemk_sweep_ncem_jacobi(nemk_stream s, int k, int row0, int n_loc,
                                       nemk_lpsrc lps, const int32_t *row_ptr,
                                       const int32_t *col, const float *wgt, double beta,
                                       const uint8_t *lab_in, uint8_t *lab_out, int32_t *dirty,
                                       int32_t *wl, int32_t *wl_count, const int32_t *rrow_ptr,
                                       const int32_t *rcol, const int32_t *heavy, int n_heavy,
                                       nemk_counters *cnt, const int32_t *skip, int copy_ranks,
                                       int shard_len, nemk_margins mg) {
    if (n_loc <= 0 && copy_ranks <= 1) return;
    int hb = (row_ptr && beta != 0.0 && heavy && n_loc > 0) ? cdiv((long long)n_heavy * 32, 256) : 0;
    int cover = copy_ranks > 1 && shard_len > n_loc ? shard_len : n_loc;   // the copy spans a full slice
    DISPATCH_K(k, (k_sweep_ncem_jacobi<KT><<<hb + cdiv(cover, JAC_TILE), 256, 0, S(s)>>>(
                      k, row0, n_loc, lps, row_ptr, col, wgt, beta, lab_in, lab_out, dirty, wl,
                      wl_count, rrow_ptr, rcol, heavy, n_heavy, hb, cnt, skip, copy_ranks,
                      shard_len, mg)));
    note_launch();
}

extern "C" void nemk_heavy_list(nemk_stream s, int row0, int n_loc, const int32_t *row_ptr,
                                int32_t *block_counts, int32_t *list, int32_t *total) {
    cudaMemsetAsync(total, 0, sizeof(int32_t), S(s));
    if (n_loc <= 0) return;
    int nb = cdiv(n_loc, HL_THREADS);
    k_heavy_count<<<nb, HL_THREADS, 0, S(s)>>>(row0, n_loc, row_ptr, block_counts);
    k_heavy_scan<<<1, HL_THREADS, 0, S(s)>>>(nb, block_counts, total);
    k_heavy_fill<<<nb, HL_THREADS, 0, S(s)>>>(row0, n_loc, row_ptr, block_counts, list);
    note_launch();
}

extern "C" void nemk_sweep_ncem_fixup(nemk_stream s, int k, int row0, int n_loc, nemk_lpsrc lps,
                                      const int32_t *row_ptr, const int32_t *col, const float *wgt,
                                      double beta, const uint8_t *lab_old, uint8_t *lab_cur,
                                      int32_t *dirty, int32_t *wl_a, int32_t *wl_b, int32_t *wl_cnt,
                                      int round, const int32_t *rrow_ptr, const int32_t *rcol,
                                      nemk_counters *cnt, const int32_t *skip,
                                      const nemk_iter_end_args *fused, nemk_margins mg) {
    nemk_iter_end_args fa;
    memset(&fa, 0, sizeof fa);
    if (fused) fa = *fused;
    if (n_loc <= 0) {
        if (fa.host)
            k_iter_end<<<1, 32, 0, S(s)>>>(fa.world, fa.cnt_all, fa.st, fa.coef, fa.decide, fa.ncem,
                                           fa.conv, fa.thr, fa.host, fa.seq);
        note_launch();
        return;
    }
#if FX_CLUSTER > 8
    DISPATCH_K(k, (cudaFuncSetAttribute(k_sweep_ncem_fixup<KT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)));
#endif
    DISPATCH_K(k, (k_sweep_ncem_fixup<KT><<<FX_CLUSTER, 1024, 0, S(s)>>>(
                      k, row0, row0 + n_loc, lps, row_ptr, col, wgt, beta, lab_old, lab_cur, dirty,
                      wl_a, wl_b, wl_cnt, round, rrow_ptr, rcol, cnt, skip, fa, mg)));
    note_launch();
}

extern "C" void nemk_sweep_ncem_fixup_round(nemk_stream s, int k, int row0, int n_loc,
                                            nemk_lpsrc lps, const int32_t *row_ptr,
                                            const int32_t *col, const float *wgt, double beta,
                                            const uint8_t *lab_old, uint8_t *lab_cur, int32_t *dirty,
                                            int32_t *wl_a, int32_t *wl_b, int32_t *wl_cnt, int round,
                                            const int32_t *rrow_ptr, const int32_t *rcol,
                                            nemk_counters *cnt, const int32_t *skip, nemk_margins mg) {
    if (n_loc <= 0) return;
    int grid = num_sms() * 8;   // CTAs beyond the list return at once; long lists need the threads
    DISPATCH_K(k, (k_sweep_ncem_fixup_round<KT><<<grid, 256, 0, S(s)>>>(
                      k, row0, row0 + n_loc, lps, row_ptr, col, wgt, beta, lab_old, lab_cur, dirty,
                      wl_a, wl_b, wl_cnt, round, rrow_ptr, rcol, cnt, skip, mg)));
    note_launch();
}

extern "C" void nemk_mark_remote(nemk_stream s, int n_glob, int row0, int n_loc, int shard_len,
                                 const uint8_t *lab_cur, const uint8_t *lab_in,
                                 const uint8_t *seen_in, uint8_t *seen_out, int32_t *dirty,
                                 int32_t *wl, int32_t *wl_count, const int32_t *rrow_ptr,
                                 const int32_t *rcol, nemk_counters *cnt, const int32_t *skip) {
    if (n_glob <= 0) return;
    int grid = cdiv(n_glob, 256);
    if (grid > num_sms() * 8) grid = num_sms() * 8;
    k_mark_remote<<<grid, 256, 0, S(s)>>>(n_glob, row0, row0 + n_loc, shard_len, lab_cur, lab_in,
                                          seen_in, seen_out, dirty, wl, wl_count, rrow_ptr, rcol, cnt, skip);
    note_launch();
}

extern "C" void nemk_delta_pack(nemk_stream s, int row0, int n_loc, int cap, const uint8_t *lab_cur,
                                const uint8_t *seen, const nemk_counters *cnt, int32_t *block,
                                const int32_t *skip) {
    cudaMemsetAsync(block, 0, 2 * sizeof(int32_t), S(s));
    if (n_loc <= 0) return;
    k_delta_pack<<<cdiv(n_loc, 256), 256, 0, S(s)>>>(row0, n_loc, cap, lab_cur, seen, cnt, block, skip);
    note_launch();
}

extern "C" void nemk_delta_apply(nemk_stream s, int world, int cap, const int32_t *blocks, int row0,
                                 int n_loc, int shard_len, uint8_t *lab_cur, uint8_t *seen,
                                 int32_t *dirty, int32_t *wl, int32_t *wl_count,
                                 const int32_t *rrow_ptr, const int32_t *rcol, nemk_counters *cnt,
                                 const int32_t *skip) {
    dim3 grid(8, world < 32 ? world : 32);
    k_delta_apply<<<grid, 256, 0, S(s)>>>(world, cap, 2 + 2 * cap, blocks, row0, row0 + n_loc, shard_len,
                                          lab_cur, seen, dirty, wl, wl_count, rrow_ptr, rcol, cnt, skip);
    note_launch();
}

extern "C" void nemk_sum_ranks_i32(nemk_stream s, int world, size_t count, const int32_t *stage,
                                   int32_t *out) {
    if (!count) return;
    k_sum_ranks_i32<<<cdiv((long long)count, 256), 256, 0, S(s)>>>(world, count, stage, out);
    note_launch();
}
extern "C" void nemk_sum_ranks_f64(nemk_stream s, int world, size_t count, const double *stage,
                                   double *out) {
    if (!count) return;
    k_sum_ranks_f64<<<cdiv((long long)count, 256), 256, 0, S(s)>>>(world, count, stage, out);
    note_launch();
}

extern "C" void nemk_graph_check(nemk_stream s, int n, int nnz, const int32_t *row_ptr,
                                 const int32_t *col, const float *wgt, int32_t *flags2) {
    cudaMemsetAsync(flags2, 0, 2 * sizeof(int32_t), S(s));
    if (n <= 0) return;
    k_graph_check<<<cdiv(n, 256), 256, 0, S(s)>>>(n, nnz, row_ptr, col, wgt, flags2);
    note_launch();
}

extern "C" void nemk_sweep_ncem_level(nemk_stream s, int k, nemk_lpsrc lps,
                                      const int32_t *row_ptr, const int32_t *col, const float *wgt,
                                      double beta, uint8_t *lab, const int32_t *sites,
                                      const int32_t *level_ptr, int lv_lo, int lv_hi, int grid_ctas,
                                      nemk_counters *cnt, const int32_t *skip) {
    int single_cta = grid_ctas <= 1;
    int grid = single_cta ? 1 : grid_ctas, threads = single_cta ? 1024 : 256;
    DISPATCH_K(k, (k_sweep_ncem_level<KT><<<grid, threads, 0, S(s)>>>(
                      k, lps, row_ptr, col, wgt, beta, lab, sites, level_ptr, lv_lo, lv_hi,
                      single_cta, cnt, skip)));
    note_launch();
}

extern "C" void nemk_sweep_nem_jacobi(nemk_stream s, int k, int row0, int n_loc, const double *logpf,
                                      const int32_t *row_ptr, const int32_t *col, const float *wgt,
                                      double beta, const float *t_in, float *t_out,
                                      nemk_counters *cnt, const int32_t *skip) {
    if (n_loc <= 0) return;
    DISPATCH_K(k, (k_sweep_nem_jacobi<KT><<<cdiv(n_loc, 256), 256, 0, S(s)>>>(
                      k, row0, n_loc, logpf, row_ptr, col, wgt, beta, t_in, t_out, cnt, skip)));
    note_launch();
}

extern "C" void nemk_sweep_nem_level(nemk_stream s, int k, const double *logpf,
                                     const int32_t *row_ptr, const int32_t *col, const float *wgt,
                                     double beta, float *t, const int32_t *sites,
                                     const int32_t *level_ptr, int lv_lo, int lv_hi, int grid_ctas,
                                     nemk_counters *cnt, const int32_t *skip) {
    int single_cta = grid_ctas <= 1;
    int grid = single_cta ? 1 : grid_ctas, threads = single_cta ? 1024 : 256;
    DISPATCH_K(k, (k_sweep_nem_level<KT><<<grid, threads, 0, S(s)>>>(
                      k, logpf, row_ptr, col, wgt, beta, t, sites, level_ptr, lv_lo, lv_hi,
                      single_cta, cnt, skip)));
    note_launch();
}

extern "C" void nemk_iter_end(nemk_stream s, int world, const nemk_counters *cnt_all,
                              const nemk_iter_status *st, nemk_coef *coef, int decide, int ncem,
                              int conv, float thr, nemk_host_status *host_slot,
                              unsigned long long seq) {
    k_iter_end<<<1, 32, 0, S(s)>>>(world, cnt_all, st, coef, decide, ncem, conv, thr, host_slot, seq);
    note_launch();
}

extern "C" void nemk_label_masks(nemk_stream s, int k, int n, int nwt, const uint8_t *lab,
                                 uint32_t *cm, int32_t *nk_int, uint8_t *lab_m, const int32_t *halt) {
    cudaMemsetAsync(nk_int, 0, (size_t)k * sizeof(int32_t), S(s));
    if (n <= 0) return;
    int grid = cdiv((long long)nwt * 32, 256);
    if (grid > num_sms() * 8) grid = num_sms() * 8;
    DISPATCH_K(k, (k_label_masks<KT><<<grid, 256, 0, S(s)>>>(k, n, nwt, lab, cm, nk_int, lab_m, halt)));
    note_launch();
}

extern "C" void nemk_mstep_ncem(nemk_stream s, int k, int d, int nwt, const uint32_t *xt,
                                const uint32_t *cm, int32_t *s_int, const int32_t *halt) {
    cudaMemsetAsync(s_int, 0, (size_t)k * d * sizeof(int32_t), S(s));
    int nwt4 = nwt / 4;
    if (nwt4 <= 0 || d <= 0) return;
    static int mu_env = -1;
    if (mu_env < 0) { const char *e = getenv("NEM_B200_MSTEP_MU"); mu_env = e ? atoi(e) : 0; }
    int mu = mu_env ? mu_env : (k <= 4 && nwt4 >= 2048 ? 4 : nwt4 >= 512 ? 2 : 1);
    if (k > 4 && mu > 2) mu = 2;   // registers: KT x MU uint4 of masks
    if (k > 8) mu = 1;
    int wgroups = cdiv(nwt4, 32 * mu);
    int gx = cdiv(wgroups, 8);
    int want = num_sms() * 16;   // CTAs: fill the machine, genomes split into chunks
    int ny = cdiv(want, gx);
    if (ny < 1) ny = 1;
    if (ny > d) ny = d;
    int dchunk = cdiv(d, ny);
    ny = cdiv(d, dchunk);
    dim3 grid(gx, ny);
#define MS(KTT, MUU) k_mstep_ncem<KTT, MUU><<<grid, 256, 0, S(s)>>>(k, d, nwt4, (const uint4 *)xt, (const uint4 *)cm, dchunk, s_int, halt)
    if (mu >= 4) { if (k <= 2) MS(2, 4); else if (k == 3) MS(3, 4); else MS(4, 4); }
    else if (mu == 2) { if (k <= 2) MS(2, 2); else if (k == 3) MS(3, 2); else if (k == 4) MS(4, 2); else MS(8, 2); }
    else { DISPATCH_K(k, (MS(KT, 1))); }
#undef MS
    note_launch();
}

extern "C" void nemk_mstep_delta(nemk_stream s, int k, int n, int d, int wpr, const uint32_t *x,
                                 const uint8_t *lab, const uint8_t *lab_m, int32_t *list, int32_t *count,
                                 int32_t *s_int, int32_t *nk_int, const int32_t *halt) {
    if (n <= 0) return;
    cudaMemsetAsync(count, 0, sizeof(int32_t), S(s));
    if (CHANGED_ROWS_VEC && ((((uintptr_t)lab) | ((uintptr_t)lab_m)) & 15) == 0)
        k_changed_rows16<<<cdiv(cdiv(n, 16), 256), 256, 0, S(s)>>>(n, lab, lab_m, list, count, halt);
    else
        k_changed_rows<<<cdiv(n, 256), 256, 0, S(s)>>>(n, lab, lab_m, list, count, halt);
    note_launch();
    int grid = num_sms() * 4;
    DISPATCH_K(k, (k_mstep_delta<KT><<<grid, 256, 0, S(s)>>>(k, d, wpr, x, lab, lab_m, list, count,
                                                            s_int, nk_int, halt)));
    note_launch();
}

extern "C" void nemk_mstep_nem(nemk_stream s, int k, int n, int d, int wpr, const uint32_t *x,
                               const float *t, int rows_per_chunk, double *partial_s,
                               double *partial_n, double *s_dbl, double *nk_dbl) {
    int nchunks = cdiv(n, rows_per_chunk);
    dim3 grid(nchunks, cdiv(d, 512));
    DISPATCH_K(k, (k_mstep_nem_partial<KT><<<grid, 128, 0, S(s)>>>(k, n, d, wpr, x, t, rows_per_chunk,
                                                                  partial_s, partial_n)));
    note_launch();
    k_mstep_nem_reduce<<<cdiv((long long)k * d, 256), 256, 0, S(s)>>>(k, d, nchunks, partial_s,
                                                                     partial_n, s_dbl, nk_dbl);
    note_launch();
}

extern "C" void nemk_mstep_finalize_tables(nemk_stream s, int k, int n, int d, int wpr,
                                           int prop_model, int disp_model, const int32_t *s_int,
                                           const int32_t *nk_int, const double *s_dbl,
                                           const double *nk_dbl, float *prop, float *center,
                                           float *disp, nemk_coef *coef, uint32_t *mask_xor,
                                           uint32_t *mask_valid, uint32_t *mask_f0,
                                           uint32_t *mask_f1, double *delta, int force_mu_changed) {
    cudaMemsetAsync(&coef->uniform_ok, 1, sizeof(int32_t), S(s));
    cudaMemsetAsync(&coef->mu_changed, force_mu_changed ? 1 : 0, sizeof(int32_t), S(s));
    // the work is K*D elements: what costs is the chain of barriers.  Always a cluster of 8 CTAs of
    // 512 threads per class (DSMEM sums); NEM_B200_FT_CLUSTER=1|2 selects one or two CTAs of 1024
    static int force_cl = -1;
    if (force_cl < 0) { const char *e = getenv("NEM_B200_FT_CLUSTER"); force_cl = e ? atoi(e) : 0; }
    int cl = force_cl ? force_cl : 8;   // measured on C4 (D = 5000): 8-CTA clusters beat one CTA per class
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(k * cl); cfg.stream = S(s); cfg.attrs = at; cfg.numAttrs = 1;
#define FT_LAUNCH(CL, TH)                                                                        \
    do {                                                                                         \
        cfg.blockDim = dim3(TH);                                                                 \
        cudaLaunchKernelEx(&cfg, k_mstep_finalize_tables<CL, TH>, k, n, d, wpr, prop_model,      \
                           disp_model, s_int, nk_int, s_dbl, nk_dbl, prop, center, disp, coef,   \
                           mask_xor, mask_valid, mask_f0, mask_f1, delta);                       \
    } while (0)
    if (cl == 1) FT_LAUNCH(1, 1024);
    else if (cl == 2) { cfg.gridDim = dim3(k * 2); FT_LAUNCH(2, 1024); }
    else { cfg.gridDim = dim3(k * 8); FT_LAUNCH(8, 512); }
#undef FT_LAUNCH
    note_launch();
}

extern "C" int nemk_criteria_partial(nemk_stream s, int k, int row0, int n_loc, nemk_lpsrc lps,
                                     const int32_t *row_ptr, const int32_t *col, const float *wgt,
                                     double beta, const uint8_t *lab, const float *t,
                                     const int32_t *heavy, int n_heavy, double *partials,
                                     int nblocks) {
    // always exactly `nblocks` partial rows (idle blocks write zeros) so that the ranks of a
    // sharded fit can gather equal-sized buffers; the first `hb` blocks take the hubs
    int hb = (row_ptr && lab && heavy && n_heavy > 0) ? cdiv((long long)n_heavy * 32, 256) : 0;
    if (hb > nblocks / 2) hb = nblocks / 2;
    if (nblocks < 2) hb = 0;
    DISPATCH_K(k, (k_criteria_partial<KT, false><<<nblocks, 256, 0, S(s)>>>(k, row0, n_loc, lps, row_ptr,
                                                                           col, wgt, beta, lab, t, heavy,
                                                                           n_heavy, hb, partials)));
    note_launch();
    return nblocks;
}
// EstimBeta's sums (nem_alg.c:2157-2191) with the criteria kernel's walk; the 4th column is 0.
// Finish with nemk_criteria_final(beta = NaN): out[0..2] = crit, grad, dsec.
extern "C" int nemk_betagrad_partial(nemk_stream s, int k, int row0, int n_loc, const int32_t *row_ptr,
                                     const int32_t *col, const float *wgt, double beta,
                                     const uint8_t *lab, const float *t, const int32_t *heavy,
                                     int n_heavy, double *partials, int nblocks) {
    int hb = (row_ptr && lab && heavy && n_heavy > 0) ? cdiv((long long)n_heavy * 32, 256) : 0;
    if (hb > nblocks / 2) hb = nblocks / 2;
    if (nblocks < 2) hb = 0;
    nemk_lpsrc lps;
    memset(&lps, 0, sizeof lps);
    DISPATCH_K(k, (k_criteria_partial<KT, true><<<nblocks, 256, 0, S(s)>>>(k, row0, n_loc, lps, row_ptr,
                                                                          col, wgt, beta, lab, t, heavy,
                                                                          n_heavy, hb, partials)));
    note_launch();
    return nblocks;
}
extern "C" void nemk_criteria_final(nemk_stream s, int nblocks_total, const double *partials,
                                    double beta, double *crit6) {
    k_criteria_final<<<1, 256, 0, S(s)>>>(nblocks_total, partials, beta, crit6);
    note_launch();
}

extern "C" void nemk_labels_to_t(nemk_stream s, int k, int n, const uint8_t *lab, float *t) {
    if (n <= 0) return;
    k_labels_to_t<<<cdiv((long long)n * k, 256), 256, 0, S(s)>>>(k, n, lab, t);
    note_launch();
}
extern "C" void nemk_t_to_labels(nemk_stream s, int k, int n, const float *t, uint8_t *lab) {
    if (n <= 0) return;
    k_t_to_labels<<<cdiv(n, 256), 256, 0, S(s)>>>(k, n, t, lab);
    note_launch();
}
extern "C" void nemk_fill_u8(nemk_stream s, uint8_t *p, int v, size_t n) {
    cudaMemsetAsync(p, v, n, S(s));
}

// ---- persistent EM kernel (nem_persist.cuh)
template <int KT>
static int persist_max_grid_t() {
    int dev = 0, coop = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_em_persist<KT>, PK_THREADS, 0) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return per_sm * num_sms();
}
// row shards: one value-preserving system-scope read-modify-write per 4 KB of a peer's exchange
// block, so that every page of a freshly opened CUDA IPC mapping has been reached from this device
// BEFORE the persistent kernel depends on it (a first access inside the kernel keeps the peers
// waiting at a cross-rank barrier)
__global__ void k_touch_peer(unsigned *p, size_t bytes) {
    const size_t stride = 4096 / sizeof(unsigned), n = bytes / 4096;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x)
        atomicOr_system(p + q * stride, 0u);
}
extern "C" void nemk_touch_peer(nemk_stream s, void *p, size_t bytes) {
    if (!p || bytes < 4096) return;
    k_touch_peer<<<64, 256, 0, S(s)>>>((unsigned *)p, bytes);
    note_launch();
}

extern "C" int nemk_persist_max_grid(int k) {
    static int cache[5] = {-1, -1, -1, -1, -1};
    int slot = k <= 2 ? 0 : k == 3 ? 1 : k == 4 ? 2 : k <= 8 ? 3 : 4;
    if (cache[slot] < 0) DISPATCH_K(k, (cache[slot] = persist_max_grid_t<KT>()));
    return cache[slot];
}
extern "C" void nemk_persist_launch(nemk_stream s, const nemk_persist_args *a, int grid) {
    nemk_persist_args copy = *a;
    void *params[1] = {&copy};
    DISPATCH_K(a->K, (cudaLaunchCooperativeKernel((void *)k_em_persist<KT>, dim3(grid), dim3(PK_THREADS),
                                                  params, 0, S(s))));
    note_launch();
}
