// nem_sub_kernels.cu -- device-side builder of a GENOME SUBSAMPLE of the resident pangenome
// (sm_100a), and the vote accumulation of the resample driver.
//
// Reference behaviour restated: PPanGGOLiN re-serialises the pangenome for every organism subset
// (ppanggolin.py:821-930, called from the chunk loop ppanggolin.py:1045-1095 and from the
// evolution-curve workers command_line.py:262-281): families with no organism in the subset are
// dropped and the rest renumbered in order (ppanggolin.py:847-852), a neighbour is kept when the
// edge exists in at least one selected organism and its weight is that count, "coverage"
// (ppanggolin.py:862-880).  Here the same subsample is built from the packed X already in HBM:
//   active_i  = (x_i & mask) != 0                      new id = exclusive scan of the flags
//   X'        = the selected genome columns of the active rows, re-packed (bit gather)
//   w'_ij     = popc(E_e & mask)  with E_e the edge x organism presence bits, or -- when no
//               explicit E is supplied -- the co-presence model E_e = x_i & x_j; w' = 0 => dropped
// All of it is HBM-bound bit work: one pass over X for the flags, one for the gather, one over the
// CSR (+ the two endpoint rows) for the weights.
#include "nem_device.h"

#include <cuda_runtime.h>
#include <stdint.h>

#define FULL 0xffffffffu
static inline cudaStream_t S(nemk_stream s) { return (cudaStream_t)s; }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// one warp per family: any selected genome present?
__global__ void __launch_bounds__(256)
k_sub_active(int n, int wpr4, const uint4 *__restrict__ x, const uint4 *__restrict__ mask,
             int32_t *__restrict__ flag) {
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= n) return;
    uint32_t any = 0u;
    for (int c = lane; c < wpr4; c += 32) {
        uint4 v = __ldg(x + (size_t)row * wpr4 + c), m = __ldg(mask + c);
        any |= (v.x & m.x) | (v.y & m.y) | (v.z & m.z) | (v.w & m.w);
    }
    any = __reduce_or_sync(FULL, any);
    if (lane == 0) flag[row] = any != 0u;
}

// ---- exclusive scan of int32 values, three kernels (per-block sums, scan of the block sums by one
// CTA, fill); out may alias in.  total receives the grand sum.
#define SC_THREADS 1024
__global__ void __launch_bounds__(SC_THREADS)
k_scan_block_sums(int n, const int32_t *__restrict__ in, int32_t *__restrict__ block_sums) {
    __shared__ int wsum[32];
    int i = blockIdx.x * SC_THREADS + threadIdx.x;
    int v = i < n ? in[i] : 0;
    v = __reduce_add_sync(FULL, v);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        int s = __reduce_add_sync(FULL, wsum[threadIdx.x]);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = s;
    }
}
__global__ void __launch_bounds__(SC_THREADS)
k_scan_blocks(int nblocks, int32_t *block_sums, int32_t *total) {
    __shared__ int wsum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int b0 = 0; b0 < nblocks; b0 += SC_THREADS) {
        int b = b0 + threadIdx.x;
        int v = b < nblocks ? block_sums[b] : 0, x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(FULL, x, o); if (lane >= o) x += y; }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        if (w == 0) {
            int ws = wsum[lane], z = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(FULL, z, o); if (lane >= o) z += y; }
            wsum[lane] = z - ws;
        }
        __syncthreads();
        int excl = carry + wsum[w] + x - v;
        if (b < nblocks) block_sums[b] = excl;
        __syncthreads();
        if (threadIdx.x == SC_THREADS - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(SC_THREADS)
k_scan_fill(int n, const int32_t *in, const int32_t *__restrict__ block_offsets, int32_t *out) {
    __shared__ int wsum[32];
    int i = blockIdx.x * SC_THREADS + threadIdx.x;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int v = i < n ? in[i] : 0, x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(FULL, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    if (w == 0) {
        int ws = wsum[lane], z = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(FULL, z, o); if (lane >= o) z += y; }
        wsum[lane] = z - ws;
    }
    __syncthreads();
    if (i < n) out[i] = block_offsets[blockIdx.x] + wsum[w] + x - v;
    if (i == n - 1) out[n] = block_offsets[blockIdx.x] + wsum[w] + x;   // out has n+1 entries
}

static void exclusive_scan(cudaStream_t st, int n, const int32_t *in, int32_t *out /*[n+1]*/,
                           int32_t *block_tmp, int32_t *total) {
    int nb = cdiv(n, SC_THREADS);
    k_scan_block_sums<<<nb, SC_THREADS, 0, st>>>(n, in, block_tmp);
    k_scan_blocks<<<1, SC_THREADS, 0, st>>>(nb, block_tmp, total);
    k_scan_fill<<<nb, SC_THREADS, 0, st>>>(n, in, block_tmp, out);
}

// bit gather: output word q of the new row = bits cols[32q .. 32q+31] of the old row.
// A CTA owns SG_ROWS consecutive old rows: they are staged in shared memory with coalesced loads
// (inactive rows too -- they are cheap and keep the loads regular), the column list is staged
// once, and the (row, output word) items are spread over all threads.
#define SG_ROWS 64
__global__ void __launch_bounds__(256)
k_sub_gather(int n, int wpr, int d_eff, int wpr_new, const uint32_t *__restrict__ x,
             const int32_t *__restrict__ cols, const int32_t *__restrict__ flag,
             const int32_t *__restrict__ new_id, uint32_t *__restrict__ x_new,
             int32_t *__restrict__ index) {
    extern __shared__ int32_t s_dyn[];
    int32_t *s_cols = s_dyn;                                   // wpr_new * 32 entries (-1 = padding)
    uint32_t *s_rows = (uint32_t *)(s_dyn + wpr_new * 32);     // SG_ROWS x (wpr + 1)
    __shared__ int s_nid[SG_ROWS];
    const int r0 = blockIdx.x * SG_ROWS, rows = min(SG_ROWS, n - r0), stride = wpr + 1;
    for (int q = threadIdx.x; q < wpr_new * 32; q += blockDim.x) s_cols[q] = q < d_eff ? cols[q] : -1;
    for (int q = threadIdx.x; q < rows * wpr; q += blockDim.x) {
        int rr = q / wpr, w = q - rr * wpr;
        s_rows[rr * stride + w] = x[(size_t)(r0 + rr) * wpr + w];
    }
    if (threadIdx.x < rows) {
        int row = r0 + threadIdx.x;
        int nid = flag[row] ? new_id[row] : -1;
        s_nid[threadIdx.x] = nid;
        if (nid >= 0) index[nid] = row;
    }
    __syncthreads();
    for (int it = threadIdx.x; it < rows * wpr_new; it += blockDim.x) {
        int rr = it / wpr_new, q = it - rr * wpr_new;
        int nid = s_nid[rr];
        if (nid < 0) continue;
        const uint32_t *xr = s_rows + rr * stride;
        const int32_t *cq = s_cols + q * 32;
        uint32_t out = 0u;
#pragma unroll 8
        for (int b = 0; b < 32; b++) {
            int c = cq[b];
            if (c >= 0) out |= ((xr[c >> 5] >> (c & 31)) & 1u) << b;
        }
        x_new[(size_t)nid * wpr_new + q] = out;
    }
}

// edge weights under the mask: one warp per (old) row; the warp is split into groups of LPE lanes
// (LPE = the uint4 chunks of a row, rounded up to a power of two, at most 32) and every group takes
// one CSR entry at a time.  w_tmp[e] = coverage (0 = dropped); cnt[new_id[i]] = kept entries of the
// row; maxdeg via atomicMax.
template <int LPE>
__global__ void __launch_bounds__(256)
k_sub_edges(int n, int wpr4, const uint4 *__restrict__ x, const uint4 *__restrict__ mask,
            const uint4 *__restrict__ edge_bits /* [nnz][wpr4] or null */,
            const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
            const int32_t *__restrict__ flag, const int32_t *__restrict__ new_id,
            float *__restrict__ w_tmp, int32_t *__restrict__ cnt, int32_t *maxdeg) {
    constexpr int G = 32 / LPE;
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= n) return;
    const int g = lane / LPE, sub = lane % LPE;
    int lo = row_ptr[row], hi = row_ptr[row + 1];
    bool act = flag[row] != 0;
    int kept = 0;
    for (int e0 = lo; e0 < hi; e0 += G) {      // warp-uniform trip count
        int e = e0 + g;
        int c = 0;
        if (e < hi && act) {
            int j = col[e];
            if (flag[j]) {
                for (int q = sub; q < wpr4; q += LPE) {
                    uint4 m = __ldg(mask + q), a;
                    if (edge_bits) a = __ldg(edge_bits + (size_t)e * wpr4 + q);
                    else {
                        uint4 u = __ldg(x + (size_t)row * wpr4 + q), v = __ldg(x + (size_t)j * wpr4 + q);
                        a = make_uint4(u.x & v.x, u.y & v.y, u.z & v.z, u.w & v.w);
                    }
                    c += __popc(a.x & m.x) + __popc(a.y & m.y) + __popc(a.z & m.z) + __popc(a.w & m.w);
                }
            }
        }
#pragma unroll
        for (int o = LPE / 2; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
        if (sub == 0 && e < hi) { w_tmp[e] = (float)c; kept += c > 0; }
    }
    kept = __reduce_add_sync(FULL, kept);
    if (lane == 0 && act) {
        cnt[new_id[row]] = kept;
        if (kept) atomicMax(maxdeg, kept);
    }
}

// compaction of the kept entries, file order preserved: one thread per old row
__global__ void __launch_bounds__(256)
k_sub_fill(int n, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
           const float *__restrict__ w_tmp, const int32_t *__restrict__ flag,
           const int32_t *__restrict__ new_id, const int32_t *__restrict__ new_row_ptr,
           int32_t *__restrict__ col_new, float *__restrict__ wgt_new) {
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n || !flag[row]) return;
    int o = new_row_ptr[new_id[row]];
    for (int e = row_ptr[row]; e < row_ptr[row + 1]; e++) {
        float w = w_tmp[e];
        if (w > 0.f) { col_new[o] = new_id[col[e]]; wgt_new[o] = w; o++; }
    }
}

// votes[index[i]][cls[label_i]] += 1   (cls: class index -> 0 P, 1 S, 2 C, 3 U; all_u: every family U)
__global__ void __launch_bounds__(256)
k_sub_vote(int n_eff, const int32_t *__restrict__ index, const uint8_t *__restrict__ lab,
           int c0, int c1, int c2, int all_u, int32_t *votes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_eff) return;
    int l = lab[i];
    int cls = all_u ? 3 : (l == 0 ? c0 : l == 1 ? c1 : l == 2 ? c2 : 3);
    atomicAdd(&votes[(size_t)index[i] * 4 + cls], 1);
}

extern "C" void nemk_sub_active(nemk_stream s, int n, int wpr, const uint32_t *x, const uint32_t *mask,
                                int32_t *flag, int32_t *new_id, int32_t *block_tmp, int32_t *n_eff) {
    if (n <= 0) return;
    k_sub_active<<<cdiv((long long)n * 32, 256), 256, 0, S(s)>>>(n, wpr / 4, (const uint4 *)x,
                                                                (const uint4 *)mask, flag);
    exclusive_scan(S(s), n, flag, new_id, block_tmp, n_eff);
}

extern "C" void nemk_sub_gather(nemk_stream s, int n, int wpr, int d_eff, int wpr_new, const uint32_t *x,
                                const int32_t *cols, const int32_t *flag, const int32_t *new_id,
                                uint32_t *x_new, int32_t *index) {
    if (n <= 0) return;
    size_t smem = sizeof(int32_t) * ((size_t)wpr_new * 32 + (size_t)SG_ROWS * (wpr + 1));
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(k_sub_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_sub_gather<<<cdiv(n, SG_ROWS), 256, smem, S(s)>>>(n, wpr, d_eff, wpr_new, x, cols, flag, new_id,
                                                       x_new, index);
}

extern "C" void nemk_sub_edges(nemk_stream s, int n, int wpr, const uint32_t *x, const uint32_t *mask,
                               const uint32_t *edge_bits, const int32_t *row_ptr, const int32_t *col,
                               const int32_t *flag, const int32_t *new_id, float *w_tmp, int32_t *cnt,
                               int32_t *new_row_ptr, int32_t *block_tmp, int32_t *nnz_new,
                               int32_t *maxdeg, int n_cnt) {
    if (n <= 0) return;
    cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (size_t)n_cnt, S(s));
    cudaMemsetAsync(maxdeg, 0, sizeof(int32_t), S(s));
    const int wpr4 = wpr / 4, grid = cdiv((long long)n * 32, 256);
#define SE(L) k_sub_edges<L><<<grid, 256, 0, S(s)>>>(n, wpr4, (const uint4 *)x, (const uint4 *)mask, \
        (const uint4 *)edge_bits, row_ptr, col, flag, new_id, w_tmp, cnt, maxdeg)
    if (wpr4 <= 2) SE(2); else if (wpr4 <= 4) SE(4); else if (wpr4 <= 8) SE(8);
    else if (wpr4 <= 16) SE(16); else SE(32);
#undef SE
    // the scan runs over the OLD row count (an upper bound of n_eff; the tail counts are zero)
    exclusive_scan(S(s), n_cnt, cnt, new_row_ptr, block_tmp, nnz_new);
}

extern "C" void nemk_sub_fill(nemk_stream s, int n, const int32_t *row_ptr, const int32_t *col,
                              const float *w_tmp, const int32_t *flag, const int32_t *new_id,
                              const int32_t *new_row_ptr, int32_t *col_new, float *wgt_new) {
    if (n <= 0) return;
    k_sub_fill<<<cdiv(n, 256), 256, 0, S(s)>>>(n, row_ptr, col, w_tmp, flag, new_id, new_row_ptr,
                                               col_new, wgt_new);
}

extern "C" void nemk_sub_vote(nemk_stream s, int n_eff, const int32_t *index, const uint8_t *lab,
                              int c0, int c1, int c2, int all_u, int32_t *votes) {
    if (n_eff <= 0) return;
    k_sub_vote<<<cdiv(n_eff, 256), 256, 0, S(s)>>>(n_eff, index, lab, c0, c1, c2, all_u, votes);
}
