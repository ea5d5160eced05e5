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

#include <cuda_runtime.h>
#include <math_constants.h>
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#define NEM_EPSILON 1e-20  // nem_typ.h:65
#define FULL 0xffffffffu

static __device__ __forceinline__ double neg_inf() { return -CUDART_INF; }

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

// 32x32 bit-tile transpose with ballots: block = 32 families x 32 words
__global__ void k_transpose_bits(const uint32_t *__restrict__ x, int n, int wpr, int d, int nwt,
                                 uint32_t *__restrict__ xt) {
    __shared__ uint32_t tile[32][33];
    int r0 = blockIdx.x * 32, w0 = blockIdx.y * 32;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;  // 32 warps
    int r = r0 + wid, w = w0 + lane;
    tile[wid][lane] = (r < n && w < wpr) ? x[(size_t)r * wpr + w] : 0u;
    __syncthreads();
    // warp `wid` now transposes word column w0+wid: lane = family r0+lane
    uint32_t word = tile[lane][wid];
    int wcol = w0 + wid;
    uint32_t mine = 0;
#pragma unroll
    for (int b = 0; b < 32; b++) {
        uint32_t v = __ballot_sync(FULL, (word >> b) & 1u);
        if (lane == b) mine = v;
    }
    int dd = wcol * 32 + lane;
    if (wcol < wpr && dd < d) xt[(size_t)dd * nwt + blockIdx.x] = mine;
}

// =============================================================================================
// theta -> per-class tables.  DensBernoulli's per-variable term (nem_mod.c:656-670):
//   absdif = abs((int)(x - mu));  disp > EPSILON: absdif*log((1-disp)/disp) - log(1-disp)
//   else absdif != 0 -> zero density.   (1-disp)/disp and 1-disp are FLOAT expressions there.
// one CTA; class after class, variables in parallel.
// =============================================================================================
#define TT_THREADS 1024
__global__ void __launch_bounds__(TT_THREADS)
k_theta_tables(int K, int D, int wpr, const float *__restrict__ prop,
               const float *__restrict__ center, const float *__restrict__ disp, nemk_coef *coef,
               uint32_t *mask_xor, uint32_t *mask_valid, uint32_t *mask_f0, uint32_t *mask_f1,
               double *delta) {
    __shared__ double sh[32];
    __shared__ int sh_ok;
    int tid = threadIdx.x, lane = tid & 31;
    int all_ok = 1;
    for (int k = 0; k < K; k++) {
        if (tid == 0) sh_ok = 1;
        __syncthreads();
        const float e0 = disp[(size_t)k * D];
        double base_u = 0.0, base_g = 0.0;
        int ok = 1;
        for (int j0 = 0; j0 < wpr * 32; j0 += TT_THREADS) {
            int j = j0 + tid;  // wpr*32 is a multiple of 32, so whole warps stay together
            bool in = j < D;
            float mu = in ? center[(size_t)k * D + j] : 0.5f;
            float e = in ? disp[(size_t)k * D + j] : e0;
            int m0 = abs((int)(0.0f - mu)), m1 = abs((int)(1.0f - mu));
            bool live = (double)e > NEM_EPSILON;
            double a = 0.0, c = 0.0;
            if (in && live) {
                float ratio = __fdiv_rn(__fsub_rn(1.0f, e), e);
                float om = __fsub_rn(1.0f, e);
                a = log((double)ratio);
                c = -log((double)om);
            }
            double cost0 = in ? (m0 * a + c) : 0.0, cost1 = in ? (m1 * a + c) : 0.0;
            if (in) {
                delta[(size_t)k * D + j] = cost1 - cost0;
                base_g += cost0;
                base_u += c;
                if (__float_as_uint(e) != __float_as_uint(e0) || m0 > 1 || m1 > 1) ok = 0;
            }
            unsigned bx = __ballot_sync(FULL, in && m0 == 1 && m1 == 0);
            unsigned bv = __ballot_sync(FULL, in && m0 != m1);
            unsigned b0 = __ballot_sync(FULL, in && !live && m0 != 0);
            unsigned b1 = __ballot_sync(FULL, in && !live && m1 != 0);
            if (lane == 0 && (j >> 5) < wpr) {
                size_t o = (size_t)k * wpr + (j >> 5);
                mask_xor[o] = bx; mask_valid[o] = bv; mask_f0[o] = b0; mask_f1[o] = b1;
            }
        }
        if (!ok) sh_ok = 0;  // benign race: every writer stores 0
        base_u = block_sum<TT_THREADS>(base_u, sh);
        base_g = block_sum<TT_THREADS>(base_g, sh);
        __syncthreads();
        if (tid == 0) {
            double pk = prop[k];
            coef->lp[k] = (pk > NEM_EPSILON) ? log(pk) : neg_inf();  // nem_alg.c:2265-2271
            bool live = (double)e0 > NEM_EPSILON;
            if (sh_ok) {
                float ratio = __fdiv_rn(__fsub_rn(1.0f, e0), e0);
                coef->a[k] = live ? log((double)ratio) : 0.0;
                coef->base[k] = live ? base_u : 0.0;
                coef->forb[k] = live ? 0 : 1;
            } else {
                coef->a[k] = 0.0; coef->base[k] = base_g; coef->forb[k] = 0;
            }
            delta[(size_t)K * D + k] = base_g;  // general-path base: sum_d cost0_kd
        }
        all_ok &= sh_ok;
        __syncthreads();
    }
    if (tid == 0) coef->uniform_ok = all_ok;
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
                  int32_t *__restrict__ hamming) {
    if (coef->empty_class) return;  // M-step found an empty class: E-step is not run
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
                    double lp = coef->lp[k], v;
                    if (coef->forb[k]) v = h[k] ? neg_inf() : lp;
                    else v = lp - (coef->a[k] * (double)h[k] + coef->base[k]);
                    logpf[(size_t)row * K + k] = v;
                    if (hamming) hamming[(size_t)row * K + k] = h[k];
                }
            }
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
    if (coef->empty_class) return;
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
// E-step site update.  ComputeLocalProba (nem_alg.c:2546-2616) in the log domain,
// SumNeighsOfClass (nem_alg.c:2850-2884), ComputeMAP first-max (nem_alg.c:603-615).
// =============================================================================================
template <int KT>
struct SiteCtx {
    double v[KT];
};

// context from hard labels: ctx_k = sum_j w_ij [lab_j == k]; `pick(j)` returns neighbour j's label
template <int KT, typename Pick>
static __device__ __forceinline__ void ctx_labels(int K, int i, const int32_t *__restrict__ row_ptr,
                                                  const int32_t *__restrict__ col,
                                                  const float *__restrict__ wgt, Pick pick,
                                                  double *ctx) {
#pragma unroll
    for (int k = 0; k < KT; k++) ctx[k] = 0.0;
    if (!row_ptr) return;
    int lo = row_ptr[i], hi = row_ptr[i + 1];
    for (int e = lo; e < hi; e++) {
        int j = col[e];
        unsigned l = pick(j);
        double w = (double)wgt[e];
#pragma unroll
        for (int k = 0; k < KT; k++)
            if (l == (unsigned)k) ctx[k] += w;
    }
}

// returns arg max (first max); flags: bit0 = all classes have zero density, bit1 = exact tie
template <int KT>
static __device__ __forceinline__ int site_argmax(int K, const double *__restrict__ lp,
                                                  const double *ctx, double beta, int &flags) {
    double mx = neg_inf();
    int km = 0;
    double sc[KT];
#pragma unroll
    for (int k = 0; k < KT; k++) {
        sc[k] = (k < K) ? lp[k] + beta * ctx[k] : neg_inf();
        if (sc[k] > mx) { mx = sc[k]; km = k; }
    }
    flags = 0;
    if (mx == neg_inf()) { flags = 1; return 0; }
#pragma unroll
    for (int k = 0; k < KT; k++)
        if (k > km && k < K && sc[k] == mx) flags |= 2;
    return km;
}

// Speculative sequential sweep bookkeeping: every site that READS i and is visited later
// (larger index) must be re-evaluated when i's label moves.  dirty[] de-duplicates, wl[] is
// the work list of the next round.
static __device__ __forceinline__ void mark_readers(int i, const int32_t *__restrict__ rrow_ptr,
                                                    const int32_t *__restrict__ rcol,
                                                    int32_t *dirty, int32_t *wl, int32_t *wl_count) {
    int lo = rrow_ptr[i], hi = rrow_ptr[i + 1];
    for (int e = lo; e < hi; e++) {
        int j = rcol[e];
        if (j > i && atomicExch(&dirty[j], 1) == 0) wl[atomicAdd(wl_count, 1)] = j;
    }
}

// ---- ncem, parallel (Jacobi) update; also round 0 of the speculative sequential sweep
// (dirty != nullptr): changed sites queue their later readers for the fix-up rounds.
template <int KT>
__global__ void __launch_bounds__(256)
k_sweep_ncem_jacobi(int K, int n, const double *__restrict__ logpf,
                    const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                    const float *__restrict__ wgt, double beta, const uint8_t *__restrict__ lab_in,
                    uint8_t *__restrict__ lab_out, int32_t *dirty, int32_t *wl, int32_t *wl_count,
                    const int32_t *__restrict__ rrow_ptr, const int32_t *__restrict__ rcol,
                    nemk_counters *cnt, const int32_t *__restrict__ skip) {
    if (skip && *skip) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int changed = 0, flags = 0;
    if (i < n) {
        double ctx[KT];
        ctx_labels<KT>(K, i, beta != 0.0 ? row_ptr : nullptr, col, wgt,
                       [&](int j) { return (unsigned)lab_in[j]; }, ctx);
        int km = site_argmax<KT>(K, logpf + (size_t)i * K, ctx, beta, flags);
        lab_out[i] = (uint8_t)km;
        changed = (km != (int)lab_in[i]);
        if (changed && dirty) mark_readers(i, rrow_ptr, rcol, dirty, wl, wl_count);
    }
    unsigned bc = __ballot_sync(FULL, changed), bn = __ballot_sync(FULL, flags & 1),
             bt = __ballot_sync(FULL, flags & 2);
    if ((threadIdx.x & 31) == 0) {
        if (bc) atomicAdd(&cnt->changed, __popc(bc));
        if (bn) atomicAdd(&cnt->allnul, __popc(bn));
        if (bt) atomicAdd(&cnt->ties, __popc(bt));
    }
}

// ---- ncem, speculative sequential sweep, fix-up rounds (ONE CTA, no host round trips).
// The in-place index-order sweep (UPDATE_SEQ, nem_alg.c:2378-2383) defines
//     cur_i = F_i( cur_j for j<i , old_j for j>=i )
// a triangular system with a unique solution.  Round 0 (the Jacobi kernel) evaluates F with old
// everywhere; each later round re-evaluates only the sites one of whose lower-index inputs
// moved, until the work list is empty (at most DAG-depth rounds, in practice a handful because
// the data term dominates).  A site clears its dirty flag BEFORE reading its inputs and every
// change re-queues its later readers AFTER publishing the new label, so no update is lost and
// the fixed point reached is the sequential sweep's result whatever the interleaving.
template <int KT>
__global__ void __launch_bounds__(1024)
k_sweep_ncem_fixup(int K, int n, const double *__restrict__ logpf,
                   const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                   const float *__restrict__ wgt, double beta, const uint8_t *__restrict__ lab_old,
                   uint8_t *lab_cur, int32_t *dirty, int32_t *wl_a, int32_t *wl_b,
                   int32_t *wl_counts /*[2]*/, const int32_t *__restrict__ rrow_ptr,
                   const int32_t *__restrict__ rcol, nemk_counters *cnt,
                   const int32_t *__restrict__ skip) {
    if (skip && *skip) return;
    __shared__ int s_count;
    const volatile uint8_t *vcur = lab_cur;
    int32_t *cur_list = wl_a, *next_list = wl_b;
    int32_t *cur_cnt = &wl_counts[0], *next_cnt = &wl_counts[1];
    int rounds = 0, dchanged = 0;
    for (;;) {
        if (threadIdx.x == 0) { s_count = *(volatile int32_t *)cur_cnt; *next_cnt = 0; }
        __syncthreads();
        int count = s_count;
        if (count == 0) break;
        rounds++;
        for (int idx = threadIdx.x; idx < count; idx += blockDim.x) {
            int i = cur_list[idx];
            atomicExch(&dirty[i], 0);
            __threadfence_block();
            double ctx[KT];
            ctx_labels<KT>(K, i, row_ptr, col, wgt,
                           [&](int j) { return (unsigned)(j < i ? vcur[j] : lab_old[j]); }, ctx);
            int flags;
            int km = site_argmax<KT>(K, logpf + (size_t)i * K, ctx, beta, flags);
            int was = vcur[i];
            if (km != was) {
                lab_cur[i] = (uint8_t)km;
                __threadfence_block();
                mark_readers(i, rrow_ptr, rcol, dirty, next_list, next_cnt);
                int old = lab_old[i];
                dchanged += (km != old) - (was != old);
            }
        }
        __syncthreads();
        int32_t *tl = cur_list; cur_list = next_list; next_list = tl;
        int32_t *tc = cur_cnt; cur_cnt = next_cnt; next_cnt = tc;
    }
    if (dchanged) atomicAdd(&cnt->changed, dchanged);
    if (threadIdx.x == 0) {
        cnt->nfix = rounds;
        wl_counts[0] = 0; wl_counts[1] = 0;
    }
}

// ---- ncem, level-scheduled exact sequential sweep (reference order), in place.
// sites[] is sorted by (level, index); no two sites of a level are neighbours, so a level is
// updated in parallel; levels run in order (one launch per wide level, or one CTA walking a run
// of narrow levels with __syncthreads between them).
template <int KT>
__global__ void __launch_bounds__(1024)
k_sweep_ncem_level(int K, const double *__restrict__ logpf, const int32_t *__restrict__ row_ptr,
                   const int32_t *__restrict__ col, const float *__restrict__ wgt, double beta,
                   uint8_t *lab, const int32_t *__restrict__ sites,
                   const int32_t *__restrict__ level_ptr, int lv_lo, int lv_hi, int single_cta,
                   nemk_counters *cnt, const int32_t *__restrict__ skip) {
    if (skip && *skip) return;
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
            int km = site_argmax<KT>(K, logpf + (size_t)i * K, ctx, beta, flags);
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
k_sweep_nem_jacobi(int K, int n, const double *__restrict__ logpf,
                   const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                   const float *__restrict__ wgt, double beta, const float *__restrict__ t_in,
                   float *__restrict__ t_out, nemk_counters *cnt,
                   const int32_t *__restrict__ skip) {
    if (skip && *skip) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float md = 0.f;
    int allnul = 0;
    if (i < n) {
        double ctx[KT];
        ctx_fuzzy<KT>(K, i, beta != 0.0 ? row_ptr : nullptr, col, wgt, t_in, ctx);
        float tn[KT];
        md = site_softmax<KT>(K, logpf + (size_t)i * K, ctx, beta, t_in + (size_t)i * K, tn, allnul);
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
    if (skip && *skip) return;
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
              int32_t *nk) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int lane = threadIdx.x & 31;
    unsigned l = (i < n) ? lab[i] : 255u;
#pragma unroll
    for (int k = 0; k < KT; k++) {
        if (k < K) {
            unsigned m = __ballot_sync(FULL, l == (unsigned)k);
            if (lane == 0 && (i >> 5) < nwt) {
                cm[(size_t)k * nwt + (i >> 5)] = m;
                if (m) atomicAdd(&nk[k], __popc(m));
            }
        }
    }
}

// ncem: S_kd = sum_w popc(XT[d][w] & cm[k][w]).  A warp owns 32 uint4 (4096 families), keeps its
// class masks in registers and walks a chunk of genomes; integer adds => exact and order-free.
template <int KT>
__global__ void __launch_bounds__(256)
k_mstep_ncem(int K, int D, int nwt4, const uint4 *__restrict__ xt, const uint4 *__restrict__ cm,
             int dchunk, int32_t *S) {
    int lane = threadIdx.x & 31;
    int warp_in_block = threadIdx.x >> 5;
    int wg = blockIdx.x * (blockDim.x >> 5) + warp_in_block;  // word group
    int c = wg * 32 + lane;
    bool in = c < nwt4;
    uint4 m[KT];
#pragma unroll
    for (int k = 0; k < KT; k++)
        m[k] = (in && k < K) ? cm[(size_t)k * nwt4 + c] : make_uint4(0, 0, 0, 0);
    int d0 = blockIdx.y * dchunk, d1 = min(D, d0 + dchunk);
    if (wg * 32 >= nwt4) return;
#pragma unroll 4
    for (int dd = d0; dd < d1; dd++) {
        uint4 v = in ? __ldg(xt + (size_t)dd * nwt4 + c) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < KT; k++) {
            if (k < K) {
                int s = __popc(v.x & m[k].x) + __popc(v.y & m[k].y) + __popc(v.z & m[k].z) +
                        __popc(v.w & m[k].w);
                s = __reduce_add_sync(FULL, s);
                if (lane == k && s) atomicAdd(&S[(size_t)k * D + dd], s);
            }
        }
    }
}

// nem (fuzzy): partial sums over a chunk of families, thread = genome, fp64, fixed order.
template <int KT>
__global__ void __launch_bounds__(128)
k_mstep_nem_partial(int K, int n, int D, int wpr, const uint32_t *__restrict__ x,
                    const float *__restrict__ t, int rows_per_chunk, double *__restrict__ partial_s,
                    double *__restrict__ partial_n) {
    const int TILE = 128;  // families staged per pass
    __shared__ float st[TILE * KT];
    __shared__ uint32_t sx[TILE * 4];
    __shared__ double shn[4 * KT];
    int chunk = blockIdx.x, db = blockIdx.y;
    int r0 = chunk * rows_per_chunk, r1 = min(n, r0 + rows_per_chunk);
    int d = db * 128 + threadIdx.x;
    int w0 = db * 4;  // 128 genomes = 4 words
    double acc[KT], accn[KT];
#pragma unroll
    for (int k = 0; k < KT; k++) { acc[k] = 0.0; accn[k] = 0.0; }
    for (int base = r0; base < r1; base += TILE) {
        int cnt = min(TILE, r1 - base);
        __syncthreads();
        for (int q = threadIdx.x; q < cnt * K; q += 128) st[(q / K) * KT + (q % K)] = t[(size_t)base * K + q];
        for (int q = threadIdx.x; q < cnt * 4; q += 128) {
            int r = q >> 2, w = w0 + (q & 3);
            sx[q] = (w < wpr) ? x[(size_t)(base + r) * wpr + w] : 0u;
        }
        __syncthreads();
        int wsel = threadIdx.x >> 5, b = threadIdx.x & 31;
        for (int r = 0; r < cnt; r++) {
            unsigned bit = (sx[r * 4 + wsel] >> b) & 1u;
#pragma unroll
            for (int k = 0; k < KT; k++)
                if (bit && k < K) acc[k] += (double)st[r * KT + k];
        }
        if (db == 0) {  // n_k partial: thread handles family base+threadIdx.x
            if (threadIdx.x < cnt) {
#pragma unroll
                for (int k = 0; k < KT; k++)
                    if (k < K) accn[k] += (double)st[threadIdx.x * KT + k];
            }
        }
    }
    if (d < D) {
#pragma unroll
        for (int k = 0; k < KT; k++)
            if (k < K) partial_s[((size_t)chunk * K + k) * D + d] = acc[k];
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
#define FIN_THREADS 1024
__global__ void __launch_bounds__(FIN_THREADS)
k_mstep_finalize(int K, int N, int D, int prop_model, int disp_model,
                 const int32_t *__restrict__ s_int, const int32_t *__restrict__ nk_int,
                 const double *__restrict__ s_dbl, const double *__restrict__ nk_dbl, float *prop,
                 float *center, float *disp, float *iner, nemk_coef *coef) {
    __shared__ double sh[32];
    __shared__ float nkf[NEMB_MAX_K];
    __shared__ double nkd[NEMB_MAX_K];
    __shared__ int empty;
    int tid = threadIdx.x;
    if (tid == 0) empty = 0;
    __syncthreads();
    if (tid < K) {
        double v = s_int ? (double)nk_int[tid] : nk_dbl[tid];
        nkd[tid] = v;
        nkf[tid] = (float)v;
    }
    __syncthreads();
    if (tid == 0) {
        for (int k = 0; k < K; k++)
            if (!((double)nkf[k] > NEM_EPSILON)) empty = k + 1;  // nem_mod.c:1363,1404-1409
        coef->empty_class = empty;  // like the reference: the LAST empty class, 1-based
    }
    for (int q = tid; q < K * D; q += FIN_THREADS) {
        int k = q / D;
        double s = s_int ? (double)s_int[q] : s_dbl[q];
        double n = nkd[k], half = 0.5 * n;
        float mu = center[q];
        if ((double)nkf[k] > NEM_EPSILON) {
            mu = s > half ? 1.0f : (s < half ? 0.0f : 0.5f);
            center[q] = mu;
        }
        double in = s * fabs(1.0 - (double)mu) + (n - s) * fabs((double)mu);
        iner[q] = (float)in;
    }
    __syncthreads();
    if (disp_model == 3) {  // skd: nem_mod.c:1152-1170
        for (int q = tid; q < K * D; q += FIN_THREADS) {
            int k = q / D;
            if ((double)nkf[k] > NEM_EPSILON) disp[q] = __fdiv_rn(iner[q], nkf[k]);
        }
    } else if (disp_model == 1) {  // sk_: nem_mod.c:1043-1073
        for (int k = 0; k < K; k++) {
            double si = 0.0;
            for (int j = tid; j < D; j += FIN_THREADS) si += (double)iner[(size_t)k * D + j];
            si = block_sum<FIN_THREADS>(si, sh);
            if (nkf[k] > 0.f) {
                double sn = (double)nkf[k] * (double)D;
                float dk = __fdiv_rn((float)si, (float)sn);
                for (int j = tid; j < D; j += FIN_THREADS) disp[(size_t)k * D + j] = dk;
            }
        }
    } else if (disp_model == 2) {  // s_d: nem_mod.c:1104-1126, float sums over k in order
        for (int j = tid; j < D; j += FIN_THREADS) {
            float si = 0.f, sn = 0.f;
            for (int k = 0; k < K; k++) {
                sn = __fadd_rn(sn, nkf[k]);
                si = __fadd_rn(si, iner[(size_t)k * D + j]);
            }
            float dd = __fdiv_rn(si, sn);
            for (int k = 0; k < K; k++) disp[(size_t)k * D + j] = dd;
        }
    } else {  // s__: nem_mod.c:988-1015
        double si = 0.0, sn = 0.0;
        for (int k = 0; k < K; k++) {
            if (nkf[k] > 0.f) {
                for (int j = tid; j < D; j += FIN_THREADS) si += (double)iner[(size_t)k * D + j];
                sn += (double)nkf[k] * (double)D;
            }
        }
        si = block_sum<FIN_THREADS>(si, sh);
        float v = __fdiv_rn((float)si, (float)sn);
        for (int q = tid; q < K * D; q += FIN_THREADS) disp[q] = v;
    }
    if (tid < K)  // nem_mod.c:456-465
        prop[tid] = prop_model == 1 ? __fdiv_rn(nkf[tid], (float)N) : (float)(1.0 / (double)K);
}

// =============================================================================================
// Criteria.  ComputeCrit (nem_alg.c:2678-2757): D, G, L, Z per family then
// U = D + beta/2 G, M = D + beta G + Z; float64, log-domain L and Z, deterministic two-stage sum.
// =============================================================================================
template <int KT>
__global__ void __launch_bounds__(256)
k_criteria_partial(int K, int n, const double *__restrict__ logpf,
                   const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                   const float *__restrict__ wgt, double beta, const uint8_t *__restrict__ lab,
                   const float *__restrict__ t, double *__restrict__ partials) {
    __shared__ double sh[32];
    double cD = 0, cG = 0, cL = 0, cZ = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double ctx[KT];
        float ti[KT];
        if (lab) {
            ctx_labels<KT>(K, i, row_ptr, col, wgt, [&](int j) { return (unsigned)lab[j]; }, ctx);
            unsigned l = lab[i];
#pragma unroll
            for (int k = 0; k < KT; k++) ti[k] = (l == (unsigned)k) ? 1.f : 0.f;
        } else {
            ctx_fuzzy<KT>(K, i, row_ptr, col, wgt, t, ctx);
#pragma unroll
            for (int k = 0; k < KT; k++) ti[k] = (k < K) ? t[(size_t)i * K + k] : 0.f;
        }
        const double *lp = logpf + (size_t)i * K;
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
    cD = block_sum<256>(cD, sh);
    cG = block_sum<256>(cG, sh);
    cL = block_sum<256>(cL, sh);
    cZ = block_sum<256>(cZ, sh);
    if (threadIdx.x == 0) {
        partials[blockIdx.x * 4 + 0] = cD; partials[blockIdx.x * 4 + 1] = cG;
        partials[blockIdx.x * 4 + 2] = cL; partials[blockIdx.x * 4 + 3] = cZ;
    }
}

__global__ void k_criteria_final(int nblocks, const double *__restrict__ partials, double beta,
                                 double *crit6) {
    if (threadIdx.x || blockIdx.x) return;
    double D = 0, G = 0, L = 0, Z = 0;
    for (int b = 0; b < nblocks; b++) {
        D += partials[b * 4 + 0]; G += partials[b * 4 + 1];
        L += partials[b * 4 + 2]; Z += partials[b * 4 + 3];
    }
    crit6[0] = D + 0.5 * beta * G;
    crit6[1] = D; crit6[2] = L;
    crit6[3] = D + beta * G + Z;
    crit6[4] = Z; crit6[5] = G;
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

// =============================================================================================
// launch layer
// =============================================================================================
#define DISPATCH_K(K, CALL)                      \
    do {                                         \
        if ((K) <= 2) { constexpr int KT = 2; CALL; }        \
        else if ((K) == 3) { constexpr int KT = 3; CALL; }   \
        else if ((K) == 4) { constexpr int KT = 4; CALL; }   \
        else if ((K) <= 8) { constexpr int KT = 8; CALL; }   \
        else { constexpr int KT = 16; CALL; }                \
    } while (0)

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
    if (n <= 0) return;
    cudaMemsetAsync(xt, 0, (size_t)d * nwt * sizeof(uint32_t), S(s));
    dim3 grid(cdiv(n, 32), cdiv(wpr, 32));
    k_transpose_bits<<<grid, 1024, 0, S(s)>>>(x, n, wpr, d, nwt, xt);
    note_launch();
}

extern "C" void nemk_theta_tables(nemk_stream s, int k, int d, int wpr, const float *prop,
                                  const float *center, const float *disp, nemk_coef *coef,
                                  uint32_t *mask_xor, uint32_t *mask_valid, uint32_t *mask_f0,
                                  uint32_t *mask_f1, double *delta) {
    k_theta_tables<<<1, TT_THREADS, 0, S(s)>>>(k, d, wpr, prop, center, disp, coef, mask_xor,
                                              mask_valid, mask_f0, mask_f1, delta);
    note_launch();
}

template <int KT>
static void launch_density_uniform(cudaStream_t st, int K, const uint32_t *x, int n, int wpr,
                                   const nemk_coef *coef, const uint32_t *mx, const uint32_t *mv,
                                   double *logpf, int32_t *hamming) {
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
                                                          logpf, hamming);                        \
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

extern "C" void nemk_density_uniform(nemk_stream s, int k, const uint32_t *x, int n, int wpr,
                                     const nemk_coef *coef, const uint32_t *mask_xor,
                                     const uint32_t *mask_valid, double *logpf, int32_t *hamming) {
    if (n <= 0) return;
    DISPATCH_K(k, (launch_density_uniform<KT>(S(s), k, x, n, wpr, coef, mask_xor, mask_valid,
                                              logpf, hamming)));
    note_launch();
}

extern "C" void nemk_density_general(nemk_stream s, int k, const uint32_t *x, int n, int d, int wpr,
                                     const nemk_coef *coef, const uint32_t *mask_f0,
                                     const uint32_t *mask_f1, const double *delta, double *logpf) {
    if (n <= 0) return;
    // delta[K*D] is followed by base_g[K] (written by k_theta_tables)
    const double *base_g = delta + (size_t)k * d;
    int grid = cdiv((long long)n * 32, 256);
    int cap = num_sms() * 16;
    if (grid > cap) grid = cap;
    DISPATCH_K(k, (k_density_general<KT><<<grid, 256, 0, S(s)>>>(k, x, n, d, wpr, coef, nullptr,
                                                                mask_f0, mask_f1, delta, base_g,
                                                                logpf)));
    note_launch();
}

extern "C" void nemk_sweep_ncem_jacobi(nemk_stream s, int k, int n, const double *logpf,
                                       const int32_t *row_ptr, const int32_t *col, const float *wgt,
                                       double beta, const uint8_t *lab_in, uint8_t *lab_out,
                                       int32_t *dirty, int32_t *wl, int32_t *wl_count,
                                       const int32_t *rrow_ptr, const int32_t *rcol,
                                       nemk_counters *cnt, const int32_t *skip) {
    if (n <= 0) return;
    DISPATCH_K(k, (k_sweep_ncem_jacobi<KT><<<cdiv(n, 256), 256, 0, S(s)>>>(
                      k, n, logpf, row_ptr, col, wgt, beta, lab_in, lab_out, dirty, wl, wl_count,
                      rrow_ptr, rcol, cnt, skip)));
    note_launch();
}

extern "C" void nemk_sweep_ncem_fixup(nemk_stream s, int k, int n, const double *logpf,
                                      const int32_t *row_ptr, const int32_t *col, const float *wgt,
                                      double beta, const uint8_t *lab_old, uint8_t *lab_cur,
                                      int32_t *dirty, int32_t *wl_a, int32_t *wl_b,
                                      int32_t *wl_counts, const int32_t *rrow_ptr,
                                      const int32_t *rcol, nemk_counters *cnt, const int32_t *skip) {
    if (n <= 0) return;
    DISPATCH_K(k, (k_sweep_ncem_fixup<KT><<<1, 1024, 0, S(s)>>>(
                      k, n, logpf, row_ptr, col, wgt, beta, lab_old, lab_cur, dirty, wl_a, wl_b,
                      wl_counts, rrow_ptr, rcol, cnt, skip)));
    note_launch();
}

extern "C" void nemk_sweep_ncem_level(nemk_stream s, int k, const double *logpf,
                                      const int32_t *row_ptr, const int32_t *col, const float *wgt,
                                      double beta, uint8_t *lab, const int32_t *sites,
                                      const int32_t *level_ptr, int lv_lo, int lv_hi, int grid_ctas,
                                      nemk_counters *cnt, const int32_t *skip) {
    int single_cta = grid_ctas <= 1;
    int grid = single_cta ? 1 : grid_ctas, threads = single_cta ? 1024 : 256;
    DISPATCH_K(k, (k_sweep_ncem_level<KT><<<grid, threads, 0, S(s)>>>(
                      k, logpf, row_ptr, col, wgt, beta, lab, sites, level_ptr, lv_lo, lv_hi,
                      single_cta, cnt, skip)));
    note_launch();
}

extern "C" void nemk_sweep_nem_jacobi(nemk_stream s, int k, int n, const double *logpf,
                                      const int32_t *row_ptr, const int32_t *col, const float *wgt,
                                      double beta, const float *t_in, float *t_out,
                                      nemk_counters *cnt, const int32_t *skip) {
    if (n <= 0) return;
    DISPATCH_K(k, (k_sweep_nem_jacobi<KT><<<cdiv(n, 256), 256, 0, S(s)>>>(
                      k, n, logpf, row_ptr, col, wgt, beta, t_in, t_out, cnt, skip)));
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

extern "C" void nemk_label_masks(nemk_stream s, int k, int n, int nwt, const uint8_t *lab,
                                 uint32_t *cm, int32_t *nk_int) {
    cudaMemsetAsync(cm, 0, (size_t)k * nwt * sizeof(uint32_t), S(s));
    cudaMemsetAsync(nk_int, 0, (size_t)k * sizeof(int32_t), S(s));
    if (n <= 0) return;
    DISPATCH_K(k, (k_label_masks<KT><<<cdiv(n, 256), 256, 0, S(s)>>>(k, n, nwt, lab, cm, nk_int)));
    note_launch();
}

extern "C" void nemk_mstep_ncem(nemk_stream s, int k, int d, int nwt, const uint32_t *xt,
                                const uint32_t *cm, int32_t *s_int) {
    cudaMemsetAsync(s_int, 0, (size_t)k * d * sizeof(int32_t), S(s));
    int nwt4 = nwt / 4;
    if (nwt4 <= 0 || d <= 0) return;
    int wgroups = cdiv(nwt4, 32);
    int gx = cdiv(wgroups, 8);
    // enough CTAs to fill the machine: split the genomes into chunks
    int want = num_sms() * 8;
    int ny = cdiv(want, gx);
    if (ny < 1) ny = 1;
    if (ny > d) ny = d;
    int dchunk = cdiv(d, ny);
    ny = cdiv(d, dchunk);
    dim3 grid(gx, ny);
    DISPATCH_K(k, (k_mstep_ncem<KT><<<grid, 256, 0, S(s)>>>(k, d, nwt4, (const uint4 *)xt,
                                                           (const uint4 *)cm, dchunk, s_int)));
    note_launch();
}

extern "C" void nemk_mstep_nem(nemk_stream s, int k, int n, int d, int wpr, const uint32_t *x,
                               const float *t, int rows_per_chunk, double *partial_s,
                               double *partial_n, double *s_dbl, double *nk_dbl) {
    int nchunks = cdiv(n, rows_per_chunk);
    dim3 grid(nchunks, cdiv(d, 128));
    DISPATCH_K(k, (k_mstep_nem_partial<KT><<<grid, 128, 0, S(s)>>>(k, n, d, wpr, x, t, rows_per_chunk,
                                                                  partial_s, partial_n)));
    note_launch();
    k_mstep_nem_reduce<<<cdiv((long long)k * d, 256), 256, 0, S(s)>>>(k, d, nchunks, partial_s,
                                                                     partial_n, s_dbl, nk_dbl);
    note_launch();
}

extern "C" void nemk_mstep_finalize(nemk_stream s, int k, int n, int d, int prop_model,
                                    int disp_model, const int32_t *s_int, const int32_t *nk_int,
                                    const double *s_dbl, const double *nk_dbl, float *prop,
                                    float *center, float *disp, float *iner_scratch,
                                    nemk_coef *coef) {
    k_mstep_finalize<<<1, FIN_THREADS, 0, S(s)>>>(k, n, d, prop_model, disp_model, s_int, nk_int,
                                                 s_dbl, nk_dbl, prop, center, disp, iner_scratch,
                                                 coef);
    note_launch();
}

extern "C" void nemk_criteria(nemk_stream s, int k, int n, const double *logpf,
                              const int32_t *row_ptr, const int32_t *col, const float *wgt,
                              double beta, const uint8_t *lab, const float *t, double *partials,
                              int nblocks_cap, double *crit6) {
    int nb = cdiv(n, 256);
    if (nb > nblocks_cap) nb = nblocks_cap;
    if (nb < 1) nb = 1;
    DISPATCH_K(k, (k_criteria_partial<KT><<<nb, 256, 0, S(s)>>>(k, n, logpf, row_ptr, col, wgt, beta,
                                                               lab, t, partials)));
    note_launch();
    k_criteria_final<<<1, 32, 0, S(s)>>>(nb, partials, beta, crit6);
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
