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

// bit gather: the selected genome columns of every active row, re-packed.  The selected columns are
// an ascending list, so the new row is the concatenation over the source words w of
// compress(x_w, mask_w) (the bits of x_w under mask_w moved to the low end, "pext"), word w's bits
// starting at bit prefix_w = popc(mask_0..w-1).  The compress network of a word (Hacker's Delight
// 7-4: five masks) depends on the mask only: warp 0 of the CTA derives it once, then every
// (row, source word) item costs ~16 integer instructions and one or two shared-memory atomicOr
// into the staged output row -- instead of 32 dependent bit lookups per OUTPUT word.
// A CTA owns SG_ROWS consecutive old rows: staged in shared memory with coalesced loads (inactive
// rows too: cheap, and the loads stay regular); the new rows leave through shared memory as well.
#define SG_ROWS 64
__global__ void __launch_bounds__(256)
k_sub_gather(int n, int wpr, int d_eff, int wpr_new, const uint32_t *__restrict__ x,
             const uint32_t *__restrict__ mask, const int32_t *__restrict__ flag,
             const int32_t *__restrict__ new_id, uint32_t *__restrict__ x_new,
             int32_t *__restrict__ index) {
    extern __shared__ uint32_t s_dyn[];
    uint32_t *s_net = s_dyn;                                   // [8][wpr]: m0, mv0..mv4, prefix, count
    uint32_t *s_rows = s_dyn + 8 * wpr;                        // SG_ROWS x (wpr + 1)
    uint32_t *s_out = s_rows + SG_ROWS * (wpr + 1);            // SG_ROWS x (wpr_new + 1)
    __shared__ int s_nid[SG_ROWS];
    const int r0 = blockIdx.x * SG_ROWS, rows = min(SG_ROWS, n - r0), stride = wpr + 1, ostride = wpr_new + 1;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 32) {          // the compress networks and the bit offsets of the source words
        int carry = 0;
        for (int w0 = 0; w0 < wpr; w0 += 32) {
            int w = w0 + lane;
            uint32_t m = w < wpr ? mask[w] : 0u;
            int c = __popc(m), incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += y; }
            if (w < wpr) {
                s_net[0 * wpr + w] = m;
                s_net[6 * wpr + w] = (uint32_t)(carry + incl - c);
                s_net[7 * wpr + w] = (uint32_t)c;
                uint32_t mk = ~m << 1, mm = m;
#pragma unroll
                for (int i = 0; i < 5; i++) {
                    uint32_t mp = mk ^ (mk << 1);
                    mp ^= mp << 2; mp ^= mp << 4; mp ^= mp << 8; mp ^= mp << 16;
                    uint32_t mv = mp & mm;
                    s_net[(1 + i) * wpr + w] = mv;
                    mm = (mm ^ mv) | (mv >> (1 << i));
                    mk &= ~mp;
                }
            }
            carry += __shfl_sync(FULL, incl, 31);
        }
    }
    for (int q = threadIdx.x; q < rows * wpr; q += blockDim.x) {
        int rr = q / wpr, w = q - rr * wpr;
        s_rows[rr * stride + w] = x[(size_t)(r0 + rr) * wpr + w];
    }
    for (int q = threadIdx.x; q < rows * ostride; q += blockDim.x) s_out[q] = 0u;
    if (threadIdx.x < rows) {
        int row = r0 + threadIdx.x;
        int nid = flag[row] ? new_id[row] : -1;
        s_nid[threadIdx.x] = nid;
        if (nid >= 0) index[nid] = row;
    }
    __syncthreads();
    for (int it = threadIdx.x; it < rows * wpr; it += blockDim.x) {
        int rr = it / wpr, w = it - rr * wpr;
        const int cnt = (int)s_net[7 * wpr + w];
        if (s_nid[rr] < 0 || cnt == 0) continue;
        uint32_t v = s_rows[rr * stride + w] & s_net[w];
#pragma unroll
        for (int i = 0; i < 5; i++) {
            uint32_t t = v & s_net[(1 + i) * wpr + w];
            v = (v ^ t) | (t >> (1 << i));
        }
        if (v == 0u) continue;
        const int pos = (int)s_net[6 * wpr + w], sh = pos & 31;
        uint32_t *o = s_out + rr * ostride + (pos >> 5);
        atomicOr(o, v << sh);
        if (sh + cnt > 32) atomicOr(o + 1, v >> (32 - sh));
    }
    __syncthreads();
    for (int it = threadIdx.x; it < rows * wpr_new; it += blockDim.x) {
        int rr = it / wpr_new, q = it - rr * wpr_new;
        int nid = s_nid[rr];
        if (nid >= 0) x_new[(size_t)nid * wpr_new + q] = s_out[rr * ostride + q];
    }
}

// edge weights under the mask.  A warp owns 32 consecutive (old) rows and streams their contiguous
// CSR segment 32 entries at a time: neighbour ids with one coalesced load, the row of an entry by a
// five-step search over the lanes' row starts (shuffles), then groups of LPE lanes (LPE = the uint4
// chunks of a row, rounded up to a power of two, at most 32) take one entry each -- the loads of the
// unrolled batch are independent and go out together, so the kernel is bound by L2 throughput, not
// by the row_ptr -> col -> flag -> x chain of latencies a warp per row would pay for four entries.
// w_tmp[e] = coverage (0 = dropped); cnt[new_id[i]] = kept entries of the row; maxdeg via atomicMax.
// (Co-presence model: an inactive neighbour has no selected genome, so its coverage is 0 without
// looking at its flag.)
template <int LPE>
__global__ void __launch_bounds__(256)
k_sub_edges(int n, int wpr4, const uint4 *__restrict__ x, const uint4 *__restrict__ mask,
            const uint4 *__restrict__ edge_bits /* [nnz][wpr4] or null */,
            const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
            const int32_t *__restrict__ flag, const int32_t *__restrict__ new_id,
            float *__restrict__ w_tmp, int32_t *__restrict__ cnt, int32_t *maxdeg) {
    constexpr int G = 32 / LPE;
    __shared__ int s_cnt[8][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / LPE, sub = lane % LPE;
    const int r0 = (blockIdx.x * 8 + warp) * 32;
    if (r0 >= n) return;                                   // warp-uniform
    const int row_l = r0 + lane;
    const int lo_l = row_ptr[min(row_l, n)], hi_l = row_ptr[min(row_l + 1, n)];
    const int act_l = row_l < n && flag[row_l] != 0;
    s_cnt[warp][lane] = 0;
    __syncwarp();
    const int seg_lo = __shfl_sync(FULL, lo_l, 0), seg_hi = __shfl_sync(FULL, hi_l, 31);
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    const uint4 m_own = sub < wpr4 ? __ldg(mask + sub) : zero4;
    for (int e0 = seg_lo; e0 < seg_hi; e0 += 32) {         // warp-uniform trip count
        const int e_l = e0 + lane;
        const bool in_l = e_l < seg_hi;
        const int j_l = in_l ? col[e_l] : -1;
        int ri = 0;                                        // last lane whose first entry is <= e_l
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int lo_c = __shfl_sync(FULL, lo_l, ri + step);   // ri + step <= 31
            if (lo_c <= e_l) ri += step;
        }
        const int act_r = __shfl_sync(FULL, act_l, ri);
        int ok_l = in_l && act_r;
        if (edge_bits && ok_l) ok_l = flag[j_l] != 0;
#pragma unroll
        for (int t = 0; t < LPE; t++) {
            const int idx = t * G + g;
            const int j = __shfl_sync(FULL, j_l, idx), rr = __shfl_sync(FULL, ri, idx);
            const int ok = __shfl_sync(FULL, ok_l, idx);
            const int js = j < 0 ? 0 : j;
            int c = 0;
            for (int q = sub; q < wpr4; q += LPE) {
                const uint4 m = q == sub ? m_own : __ldg(mask + q);
                uint4 a;
                if (edge_bits) a = ok ? __ldg(edge_bits + (size_t)(e0 + idx) * wpr4 + q) : zero4;
                else {
                    const uint4 u = __ldg(x + (size_t)(r0 + rr) * wpr4 + q), v = __ldg(x + (size_t)js * wpr4 + q);
                    a = make_uint4(u.x & v.x, u.y & v.y, u.z & v.z, u.w & v.w);
                }
                c += __popc(a.x & m.x) + __popc(a.y & m.y) + __popc(a.z & m.z) + __popc(a.w & m.w);
            }
            if (!ok) c = 0;
#pragma unroll
            for (int o = LPE / 2; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
            if (sub == 0 && j >= 0) {
                w_tmp[e0 + idx] = (float)c;
                if (c > 0) atomicAdd(&s_cnt[warp][rr], 1);
            }
        }
    }
    __syncwarp();
    const int kept = s_cnt[warp][lane];
    if (act_l) cnt[new_id[row_l]] = kept;
    const int mx = __reduce_max_sync(FULL, kept);
    if (lane == 0 && mx > 0) atomicMax(maxdeg, mx);
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
                                const uint32_t *mask, const int32_t *flag, const int32_t *new_id,
                                uint32_t *x_new, int32_t *index) {
    if (n <= 0) return;
    size_t smem = sizeof(uint32_t) * ((size_t)8 * wpr + (size_t)SG_ROWS * (wpr + 1) + (size_t)SG_ROWS * (wpr_new + 1));
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(k_sub_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_sub_gather<<<cdiv(n, SG_ROWS), 256, smem, S(s)>>>(n, wpr, d_eff, wpr_new, x, mask, flag, new_id,
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
    const int wpr4 = wpr / 4, grid = cdiv(n, 256);      // a warp per 32 rows, 8 warps per CTA
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
