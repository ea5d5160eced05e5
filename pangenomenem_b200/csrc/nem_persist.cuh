// nem_persist.cuh -- the persistent EM kernel (included by nem_kernels.cu after the device
// functions it reuses; interface: nemk_persist_args in nem_device.h).
//
// ONE cooperative launch runs a whole ncem fit on the popcount density path.  Reference control
// flow restated: ClassifyByNemOneBeta INIT_PARAM_FILE (nem_alg.c:1151-1169) =
// ComputePartitionFromPara(Needinit=1) (1951-1989: blind sweep, beta sweep) + NemAlgo (1746-1879:
// M-step, E-step, HasConverged 2056-2112).  Each box below is a PHASE; phases are separated by a
// device-wide barrier (all CTAs are co-resident: cooperative launch), not by a kernel launch:
//
//   init      prep | tables(theta0) | H from row popcounts or X pass | blind sweep | beta sweep
//   iteration [scan changed rows | delta statistics]  or  [zero | class masks | X^T recount]
//             | closed forms + tables (one CTA per class) | X pass when the class masks moved
//             | margin test -> active list | evaluate (Jacobi) | fix-up rounds ... | decide
//
// Every CTA takes the same decisions from the same device counters (read after a barrier), so
// there is no host round trip inside a fit.  Problems whose X does not fit the L2 leave the kernel
// for the two HBM-bound passes (TMA density kernel, X^T recount) and re-enter it: exit codes
// NEMK_PK_EXIT_NEED_DENSITY / NEED_RECOUNT.
//
// Memory visibility: data another CTA wrote in an earlier phase is read with ordinary loads after
// the barrier (fence + atomic arrival + fence, the cooperative-groups grid-sync pattern); nothing
// mutable is read through __ldg / const __restrict__ kernel parameters here (the kernel takes ONE
// by-value struct, so the compiler cannot infer non-coherent loads), and lab_cur inside a fix-up
// round is read with __ldcg like in the standalone kernels.

#define PK_THREADS 512

static __device__ __forceinline__ void pk_grid_sync(unsigned *bar, unsigned nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned *vgen = bar + 1;
        const unsigned gen = *vgen;     // cannot advance before this CTA arrives
        __threadfence();                // release: this CTA's stores of the phase
        if (atomicAdd(bar, 1u) == nblocks - 1u) {
            bar[0] = 0u;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            while (*vgen == gen) { }
        }
        __threadfence();                // acquire: the other CTAs' stores
    }
    __syncthreads();
}

struct PkSweepCnt { int changed, allnul, ties; };

// ---- one site of a Jacobi round taken from a list (margin-cached round): every input is an OLD
// label, the new one goes to lab_out; a change queues the later readers for the fix-up rounds
template <int KT>
static __device__ __forceinline__ void pk_eval_site(int K, int i, const nemk_lpsrc &lps,
                                                    const int32_t *rp, const int32_t *col,
                                                    const float *wgt, double beta,
                                                    const uint8_t *lab_in, uint8_t *lab_out,
                                                    int32_t *dirty, int32_t *wl, int32_t *wl_count,
                                                    const int32_t *rrow_ptr, const int32_t *rcol, int n,
                                                    const nemk_margins &mg, double thr_store,
                                                    PkSweepCnt &c) {
    double ctx[KT], lpv[KT], margin;
    load_lp<KT>(lps, K, (size_t)i, lpv);
    const int lin = (int)lab_in[i];
    ctx_labels<KT>(K, i, rp, col, wgt, [&](int j) { return (unsigned)lab_in[j]; }, ctx);
    int fl;
    const int km = site_argmax<KT>(K, lpv, ctx, beta, fl, margin);
    lab_out[i] = (uint8_t)km;
    store_margin(mg, i, margin, thr_store);
    if (mg.m) mg.stale_cur[i] = 0;
    if (km != lin) {
        c.changed++;
        if (dirty) mark_readers(i, rrow_ptr, rcol, dirty, wl, wl_count, 0, n, mg.stale_next);
    }
    c.allnul += fl & 1;
    c.ties += (fl >> 1) & 1;
}

// the same for a hub, by a whole warp (counters on lane 0)
template <int KT>
static __device__ __forceinline__ void pk_eval_hub(int K, int i, const nemk_lpsrc &lps,
                                                   const int32_t *rp, const int32_t *col,
                                                   const float *wgt, double beta,
                                                   const uint8_t *lab_in, uint8_t *lab_out,
                                                   int32_t *dirty, int32_t *wl, int32_t *wl_count,
                                                   const int32_t *rrow_ptr, const int32_t *rcol, int n,
                                                   const nemk_margins &mg, double thr_store,
                                                   PkSweepCnt &c) {
    const int lane = threadIdx.x & 31;
    double ctx[KT], lpv[KT], margin;
    load_lp<KT>(lps, K, (size_t)i, lpv);
    ctx_labels_warp<KT>(K, i, rp, col, wgt, [&](int j) { return (unsigned)lab_in[j]; }, ctx,
                        lps.wsum_any_order != 0);
    int fl;
    const int km = site_argmax<KT>(K, lpv, ctx, beta, fl, margin);
    const int ch = km != (int)lab_in[i];
    if (lane == 0) {
        lab_out[i] = (uint8_t)km;
        store_margin(mg, i, margin, thr_store);
        if (mg.m) mg.stale_cur[i] = 0;
        c.changed += ch;
        c.allnul += fl & 1;
        c.ties += (fl >> 1) & 1;
    }
    if (ch && dirty) mark_readers_warp(i, rrow_ptr, rcol, dirty, wl, wl_count, 0, n, mg.stale_next);
}

// ---- one E-step sweep (ComputePartitionNEM nem_alg.c:2330-2405 for ncem).  All CTAs call it.
// use_graph: context term on; seq: in-place index-order semantics through the speculative fixed
// point; margins (mg.m != NULL) only with seq.  Returns the number of labels that
// differ from lab_in (every thread gets the same value).
template <int KT>
static __device__ int pk_sweep(const nemk_persist_args &a, const nemk_lpsrc &lps, double beta,
                               bool use_graph, bool seq, const uint8_t *lab_in, uint8_t *lab_out,
                               nemk_margins mg, nemk_counters *cnt, nemk_counters *cnt_next,
                               float *s_w, uint8_t *s_l, int &barriers, long long &kept_total,
                               int &nfix_total) {
    const int K = a.K, n = a.n;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nthreads = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nwarps = nthreads >> 5, gwarp = gtid >> 5;
    const int32_t *rp = use_graph ? a.row_ptr : nullptr;
    int32_t *dirty = seq ? a.dirty : nullptr;
    int32_t *wl0 = a.wl[0], *wl_cnt = a.wl_cnt;
    const SweepThr thr = sweep_thr(K, lps.coef, mg);
    const bool may_skip = mg.m && thr.test < CUDART_INF;
    PkSweepCnt c = {0, 0, 0};
    int kept = 0;

    if (may_skip) {
        // ---- phase: margin test of every site (streaming: 6 bytes per site, nothing dependent).
        // A site whose stored margin exceeds what theta can have moved and none of whose
        // later-or-equal neighbours changed in the previous sweep keeps its label; the others go
        // to the active lists (light sites: wl[1], hubs: hub_list).
        int32_t *al = a.wl[1], *al_cnt = a.scratch;
        const int ngroups = (n + 3) >> 2;
        for (int g0 = (gtid & ~31); g0 < ngroups; g0 += nthreads) {
            const int g = g0 + lane, i0 = g * 4;
            unsigned actm = 0u, hubm = 0u;
            int nv = 0;
            if (g < ngroups) {
                uint8_t st[4], lb[4];
                float mv[4];
                nv = min(4, n - i0);
                if (nv == 4) {
                    const uchar4 s4 = *reinterpret_cast<const uchar4 *>(mg.stale_cur + i0);
                    const uchar4 l4 = *reinterpret_cast<const uchar4 *>(lab_in + i0);
                    const float4 m4 = *reinterpret_cast<const float4 *>(mg.m + i0);
                    st[0] = s4.x; st[1] = s4.y; st[2] = s4.z; st[3] = s4.w;
                    lb[0] = l4.x; lb[1] = l4.y; lb[2] = l4.z; lb[3] = l4.w;
                    mv[0] = m4.x; mv[1] = m4.y; mv[2] = m4.z; mv[3] = m4.w;
                    *reinterpret_cast<uchar4 *>(lab_out + i0) = l4;   // active sites are rewritten below
                } else {
                    for (int q = 0; q < 4; q++) {
                        const bool v = q < nv;
                        st[q] = v ? mg.stale_cur[i0 + q] : (uint8_t)0;
                        lb[q] = v ? lab_in[i0 + q] : (uint8_t)0;
                        mv[q] = v ? mg.m[i0 + q] : CUDART_INF_F;
                        if (v) lab_out[i0 + q] = lb[q];
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const bool keep = !st[q] && (double)mv[q] > thr.test && lb[q] != 255;
                    if (q < nv && !keep) actm |= 1u << q;
                }
                kept += nv - __popc(actm);
                if (actm && rp && a.n_heavy) {
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (((actm >> q) & 1u) && rp[i0 + q + 1] - rp[i0 + q] > HEAVY_DEG) hubm |= 1u << q;
                }
            }
            // warp-aggregated append: one atomic per warp and list
            const int cl = __popc(actm & ~hubm), chb = __popc(hubm);
            int incl = cl, inch = chb;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL, incl, o), w = __shfl_up_sync(FULL, inch, o);
                if (lane >= o) { incl += v; inch += w; }
            }
            const int tl = __shfl_sync(FULL, incl, 31), th = __shfl_sync(FULL, inch, 31);
            int bl = 0, bh = 0;
            if (lane == 31) {
                if (tl) bl = atomicAdd(&al_cnt[0], tl);
                if (th) bh = atomicAdd(&al_cnt[1], th);
            }
            bl = __shfl_sync(FULL, bl, 31) + incl - cl;
            bh = __shfl_sync(FULL, bh, 31) + inch - chb;
            unsigned m = actm;
            while (m) {
                const int q = __ffs(m) - 1;
                m &= m - 1;
                if ((hubm >> q) & 1u) a.hub_list[bh++] = i0 + q;
                else al[bl++] = i0 + q;
            }
        }
        pk_grid_sync(a.bar, gridDim.x); barriers++;
        // ---- phase: evaluate the active sites, spread evenly over the grid (hubs first)
        const int nl = *(volatile int32_t *)&al_cnt[0], nh = *(volatile int32_t *)&al_cnt[1];
        for (int wi = gwarp; wi < nh; wi += nwarps)
            pk_eval_hub<KT>(K, a.hub_list[wi], lps, rp, a.col, a.wgt, beta, lab_in, lab_out, dirty, wl0,
                            &wl_cnt[0], a.rrow_ptr, a.rcol, n, mg, thr.store, c);
        for (int q = gtid; q < nl; q += nthreads)
            pk_eval_site<KT>(K, al[q], lps, rp, a.col, a.wgt, beta, lab_in, lab_out, dirty, wl0,
                             &wl_cnt[0], a.rrow_ptr, a.rcol, n, mg, thr.store, c);
    } else {
        // ---- phase: dense Jacobi round, every site evaluated.  Hubs by one warp each (longest
        // work first), then warps over 32 consecutive sites whose contiguous CSR segment is
        // streamed cooperatively (ctx_labels_coop)
        const bool hubs = rp && a.n_heavy > 0;
        if (hubs)
            for (int wi = gwarp; wi < a.n_heavy; wi += nwarps)
                pk_eval_hub<KT>(K, a.heavy[wi], lps, rp, a.col, a.wgt, beta, lab_in, lab_out, dirty,
                                wl0, &wl_cnt[0], a.rrow_ptr, a.rcol, n, mg, thr.store, c);
        for (int base = gwarp * 32; base < n; base += nwarps * 32) {
            const int i = base + lane;
            const bool valid = i < n;
            int lo = 0, hi = 0;
            if (rp && valid) { lo = rp[i]; hi = rp[i + 1]; }
            const bool is_heavy = hubs && (hi - lo > HEAVY_DEG);
            double ctx[KT], lpv[KT];
#pragma unroll
            for (int k = 0; k < KT; k++) ctx[k] = 0.0;
            if (valid) load_lp<KT>(lps, K, (size_t)i, lpv);
            if (rp) {
                const int seg_lo = __reduce_min_sync(FULL, valid ? lo : 0x7fffffff);
                const int seg_hi = __reduce_max_sync(FULL, valid ? hi : 0);
                if (is_heavy) lo = hi = 0;
                if (seg_lo < seg_hi)
                    ctx_labels_coop<KT>(lo, hi, seg_lo, seg_hi, a.col, a.wgt, lab_in,
                                        s_w + wib * COOP_CHUNK, s_l + wib * COOP_CHUNK, ctx);
            }
            if (valid && !is_heavy) {
                double margin;
                int fl;
                const int km = site_argmax<KT>(K, lpv, ctx, beta, fl, margin);
                lab_out[i] = (uint8_t)km;
                store_margin(mg, i, margin, thr.store);
                if (mg.m) mg.stale_cur[i] = 0;
                if (km != (int)lab_in[i]) {
                    c.changed++;
                    if (dirty) mark_readers(i, a.rrow_ptr, a.rcol, dirty, wl0, &wl_cnt[0], 0, n, mg.stale_next);
                }
                c.allnul += fl & 1;
                c.ties += (fl >> 1) & 1;
            }
        }
    }
    {
        const int ch = __reduce_add_sync(FULL, c.changed), an = __reduce_add_sync(FULL, c.allnul);
        const int ti = __reduce_add_sync(FULL, c.ties), kp = __reduce_add_sync(FULL, kept);
        if (lane == 0) {
            if (ch) atomicAdd(&cnt->changed, ch);
            if (an) atomicAdd(&cnt->allnul, an);
            if (ti) atomicAdd(&cnt->ties, ti);
            if (kp) atomicAdd(&cnt->kept, kp);
        }
    }
    pk_grid_sync(a.bar, gridDim.x); barriers++;
    if (gtid == 0) {
        if (may_skip) { a.scratch[0] = 0; a.scratch[1] = 0; }   // consumed: zero at rest
        // the counter block of the NEXT sweep: last read right after the previous sweep's final
        // barrier, and every CTA has passed a barrier of this sweep since
        cnt_next->changed = 0; cnt_next->nfix = 0; cnt_next->allnul = 0; cnt_next->ties = 0;
        cnt_next->maxdiff = 0.f; cnt_next->pending = 0; cnt_next->changed_glob = 0; cnt_next->kept = 0;
    }

    // ---- fix-up rounds (speculative sequential sweep): re-evaluate the sites one of whose
    // lower-index inputs moved until the work list is empty; one barrier per round
    int rounds = 0;
    if (seq) {
        int dchanged = 0;
        for (int round = 0;; round++) {
            int32_t *cur_list = (round & 1) ? a.wl[1] : a.wl[0], *next_list = (round & 1) ? a.wl[0] : a.wl[1];
            int32_t *next_cnt = &wl_cnt[(round + 1) & 3];
            const int count = *(volatile int32_t *)&wl_cnt[round & 3];
            if (gtid == 0) wl_cnt[(round + 2) & 3] = 0;   // idle during this round
            if (count == 0) break;
            rounds++;
            for (int base = (gtid & ~31); base < count; base += nthreads)
                dchanged += fixup_items<KT>(K, base + lane, count, cur_list, 0, n, lps, rp, a.col, a.wgt,
                                            beta, lab_in, lab_out, dirty, next_list, next_cnt,
                                            a.rrow_ptr, a.rcol, mg, thr.store);
            if (dchanged) { atomicAdd(&cnt->changed, dchanged); dchanged = 0; }
            pk_grid_sync(a.bar, gridDim.x); barriers++;
        }
        // every CTA has read the last (zero) count before anybody can append again: the next
        // appends happen at least one barrier later (the next sweep's evaluation phase)
        if (gtid == 0) {
            wl_cnt[0] = 0; wl_cnt[1] = 0; wl_cnt[2] = 0; wl_cnt[3] = 0;
            if (mg.m) const_cast<nemk_coef *>(lps.coef)->drift = thr.store;
        }
    }
    nfix_total += rounds;
    const volatile nemk_counters *vc = cnt;
    kept_total += vc->kept;
    return vc->changed;
}

// ---- closed forms + tables of class k by ONE CTA (k_mstep_finalize_tables with block sums instead
// of cluster sums; same float expressions: nem_mod.c:455-465, 965-1174, 1422-1479, 1646-1704)
template <int TH>
static __device__ void pk_finalize_class(int k, const nemk_persist_args &a, double *sh,
                                         float *nkf, double *nkd) {
    const int K = a.K, D = a.D, N = a.n, tid = threadIdx.x;
    const int32_t *s_int = a.stat, *nk_int = a.stat + (size_t)K * D;
    float *center = a.center, *disp = a.disp;
    const int wreal = (D + 31) >> 5;
    __syncthreads();
    if (tid < K) {
        const double v = (double)nk_int[tid];
        nkd[tid] = v;
        nkf[tid] = (float)v;
    }
    __syncthreads();
    if (k == 0 && tid == 0) {
        int empty = 0;
        for (int c = 0; c < K; c++)
            if (!((double)nkf[c] > NEM_EPSILON)) empty = c + 1;   // nem_mod.c:1363,1404-1409
        a.coef->empty_class = empty;
    }
    auto S_of = [&](int c, int j) -> double { return (double)s_int[(size_t)c * D + j]; };
    const bool nonempty = (double)nkf[k] > NEM_EPSILON;
    if (nonempty)
        for (int j = tid; j < D; j += TH) {
            const double s = S_of(k, j), half = 0.5 * nkd[k];
            center[(size_t)k * D + j] = s > half ? 1.0f : (s < half ? 0.0f : 0.5f);
        }
    if (a.disp_model == 3) {
        if (nonempty)
            for (int j = tid; j < D; j += TH)
                disp[(size_t)k * D + j] = __fdiv_rn(iner_of(S_of(k, j), nkd[k], true, 0.f), nkf[k]);
    } else if (a.disp_model == 2) {
        for (int j = tid; j < D; j += TH) {
            float si = 0.f, sn = 0.f;
            for (int c = 0; c < K; c++) {
                const bool ne = (double)nkf[c] > NEM_EPSILON;
                sn = __fadd_rn(sn, nkf[c]);
                si = __fadd_rn(si, iner_of(S_of(c, j), nkd[c], ne, ne ? 0.f : center[(size_t)c * D + j]));
            }
            disp[(size_t)k * D + j] = __fdiv_rn(si, sn);
        }
    } else {
        double v = 0.0, sn = 0.0;
        if (a.disp_model == 1) {
            for (int j = tid; j < D; j += TH)
                v += (double)iner_of(S_of(k, j), nkd[k], nonempty, nonempty ? 0.f : center[(size_t)k * D + j]);
            sn = (double)nkf[k] * (double)D;
        } else {
            for (int c = 0; c < K; c++) {
                if (nkf[c] > 0.f) {
                    const bool ne = (double)nkf[c] > NEM_EPSILON;
                    for (int j = tid; j < D; j += TH)
                        v += (double)iner_of(S_of(c, j), nkd[c], ne, ne ? 0.f : center[(size_t)c * D + j]);
                    sn += (double)nkf[c] * (double)D;
                }
            }
        }
        v = block_sum<TH>(v, sh);
        if (a.disp_model == 0 || nkf[k] > 0.f) {
            const float dk = __fdiv_rn((float)v, (float)sn);
            for (int j = tid; j < D; j += TH) disp[(size_t)k * D + j] = dk;
        }
    }
    if (tid == 0)
        a.prop[k] = a.prop_model == 1 ? __fdiv_rn(nkf[k], (float)N) : (float)(1.0 / (double)K);
    __threadfence_block();
    __syncthreads();
    const ClassCoef cc = class_coef(__ldcg(&disp[(size_t)k * D]));
    TablesPartial p = tables_words(k, D, a.wpr, 0, wreal, cc, center, disp, a.mxor, a.mval, a.f0, a.f1, a.delta);
    for (int w = wreal + tid; w < a.wpr; w += TH) {
        const size_t o = (size_t)k * a.wpr + w;
        a.mxor[o] = 0u; a.mval[o] = 0u; a.f0[o] = 0u; a.f1[o] = 0u;
    }
    if (p.mu_moved) atomicOr(&a.coef->mu_changed, 1);
    const double base_u = block_sum<TH>(p.base_u, sh), base_g = block_sum<TH>(p.base_g, sh);
    const double notok = block_sum<TH>((double)p.notok, sh);
    const double nv = block_sum<TH>((double)p.n_valid, sh), nx = block_sum<TH>((double)p.n_x1, sh);
    if (tid == 0)
        tables_commit(k, K, D, a.prop, a.coef, a.delta, cc, base_u, base_g, notok == 0.0, (int)nv, (int)nx, true);
}

// tables of the theta the caller supplied (k_theta_tables), class k by one CTA
template <int TH>
static __device__ void pk_tables_class(int k, const nemk_persist_args &a, double *sh) {
    const ClassCoef cc = class_coef(a.disp[(size_t)k * a.D]);
    TablesPartial p = tables_words(k, a.D, a.wpr, 0, a.wpr, cc, a.center, a.disp, a.mxor, a.mval,
                                   a.f0, a.f1, a.delta);
    if (p.mu_moved) atomicOr(&a.coef->mu_changed, 1);
    const double base_u = block_sum<TH>(p.base_u, sh), base_g = block_sum<TH>(p.base_g, sh);
    const double notok = block_sum<TH>((double)p.notok, sh);
    const double nv = block_sum<TH>((double)p.n_valid, sh), nx = block_sum<TH>((double)p.n_x1, sh);
    if (threadIdx.x == 0)
        tables_commit(k, a.K, a.D, a.prop, a.coef, a.delta, cc, base_u, base_g, notok == 0.0, (int)nv,
                      (int)nx, false);
}

// ---- in-kernel X pass for L2-sized problems (k_density_uniform's arithmetic, Hamming counts
// only): lpr lanes per family, uint4 loads, class masks through L1
template <int KT>
static __device__ void pk_density_pass(const nemk_persist_args &a) {
    const int K = a.K, n = a.n, wpr4 = a.wpr >> 2, lane = threadIdx.x & 31;
    int lpr = 1;
    while (lpr < wpr4 && lpr < 32) lpr <<= 1;
    const int rpw = 32 / lpr, sub = lane % lpr;
    const int nwarps = (gridDim.x * blockDim.x) >> 5, gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint4 *x = reinterpret_cast<const uint4 *>(a.x);
    const uint4 *mx = reinterpret_cast<const uint4 *>(a.mxor), *mv = reinterpret_cast<const uint4 *>(a.mval);
    int kind[KT];
#pragma unroll
    for (int k = 0; k < KT; k++) kind[k] = k < K ? a.coef->kind[k] : 3;
    for (int rbase = gwarp * rpw; rbase < n; rbase += nwarps * rpw) {
        const int row = rbase + lane / lpr;
        int h[KT], P = 0;
#pragma unroll
        for (int k = 0; k < KT; k++) h[k] = 0;
        if (row < n) {
            const uint4 *xr = x + (size_t)row * wpr4;
            for (int c = sub; c < wpr4; c += lpr) {
                const uint4 v = __ldg(xr + c);
                P += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
#pragma unroll
                for (int k = 0; k < KT; k++) {
                    if (k < K && kind[k] == 0) {
                        const uint4 p = mx[k * wpr4 + c], q = mv[k * wpr4 + c];
                        h[k] += __popc((v.x ^ p.x) & q.x) + __popc((v.y ^ p.y) & q.y) +
                                __popc((v.z ^ p.z) & q.z) + __popc((v.w ^ p.w) & q.w);
                    }
                }
            }
        }
        for (int o = lpr >> 1; o > 0; o >>= 1) {
            P += __shfl_xor_sync(FULL, P, o);
#pragma unroll
            for (int k = 0; k < KT; k++) h[k] += __shfl_xor_sync(FULL, h[k], o);
        }
        if (row < n && sub == 0) {
#pragma unroll
            for (int k = 0; k < KT; k++)
                if (k < K)
                    a.ham[(size_t)row * K + k] = kind[k] == 0 ? h[k] : kind[k] == 1 ? P : kind[k] == 2 ? a.D - P : 0;
        }
    }
}

// ---- in-kernel full recount for L2-sized problems: class masks + n_k, then S = popc(X^T & cm)
template <int KT>
static __device__ void pk_label_masks(const nemk_persist_args &a, const uint8_t *lab) {
    const int K = a.K, n = a.n, nwt = a.nwt, lane = threadIdx.x & 31;
    int32_t *nk = a.stat + (size_t)K * a.D;
    const int nthreads = gridDim.x * blockDim.x;
    int cnt[KT];
#pragma unroll
    for (int k = 0; k < KT; k++) cnt[k] = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nwt * 32; i += nthreads) {
        const unsigned l = (i < n) ? lab[i] : 255u;
#pragma unroll
        for (int k = 0; k < KT; k++) {
            if (k < K) {
                const unsigned m = __ballot_sync(FULL, l == (unsigned)k);
                if (lane == 0) { a.cm[(size_t)k * nwt + (i >> 5)] = m; cnt[k] += __popc(m); }
            }
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < KT; k++)
            if (k < K && cnt[k]) atomicAdd(&nk[k], cnt[k]);
    }
}

template <int KT>
static __device__ void pk_recount(const nemk_persist_args &a) {
    const int K = a.K, D = a.D, nwt4 = a.nwt >> 2, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5, gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint4 *xt = reinterpret_cast<const uint4 *>(a.xt);
    const uint4 *cm = reinterpret_cast<const uint4 *>(a.cm);
    const int wgroups = (nwt4 + 31) >> 5;
    // items = (word group of 32 uint4, chunk of genomes); ~4 items per warp
    int ny = max(1, (4 * nwarps) / wgroups);
    if (ny > D) ny = D;
    const int dchunk = (D + ny - 1) / ny;
    ny = (D + dchunk - 1) / dchunk;
    for (int it = gwarp; it < wgroups * ny; it += nwarps) {
        const int wg = it % wgroups, dy = it / wgroups;
        const int c = wg * 32 + lane;
        uint4 m[KT];
#pragma unroll
        for (int k = 0; k < KT; k++) m[k] = (c < nwt4 && k < K) ? cm[(size_t)k * nwt4 + c] : make_uint4(0, 0, 0, 0);
        const int d1 = min(D, (dy + 1) * dchunk);
#pragma unroll 2
        for (int dd = dy * dchunk; dd < d1; dd++) {
            const uint4 v = (c < nwt4) ? __ldg(xt + (size_t)dd * nwt4 + c) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int k = 0; k < KT; k++) {
                if (k < K) {
                    int s = __popc(v.x & m[k].x) + __popc(v.y & m[k].y) + __popc(v.z & m[k].z) + __popc(v.w & m[k].w);
                    s = __reduce_add_sync(FULL, s);
                    if (lane == k && s) atomicAdd(&a.stat[(size_t)k * D + dd], s);
                }
            }
        }
    }
}

// ---- scan of the rows whose label differs from the one the statistics describe (16 per thread)
static __device__ void pk_scan_changed(int n, const uint8_t *lab, const uint8_t *lab_m, int32_t *list,
                                       int32_t *count) {
    const int nthreads = gridDim.x * blockDim.x, lane = threadIdx.x & 31;
    const int ngroups = (n + 15) >> 4;
    for (int t0 = ((blockIdx.x * blockDim.x + threadIdx.x) & ~31); t0 < ngroups; t0 += nthreads) {
        const int t = t0 + lane, i0 = t * 16;
        uint32_t dm = 0u;
        if (t < ngroups) {
            if (i0 + 16 <= n) {
                const uint4 p = *(reinterpret_cast<const uint4 *>(lab) + t);
                const uint4 q = *(reinterpret_cast<const uint4 *>(lab_m) + t);
                dm = nonzero_bytes(p.x ^ q.x) | (nonzero_bytes(p.y ^ q.y) << 4) |
                     (nonzero_bytes(p.z ^ q.z) << 8) | (nonzero_bytes(p.w ^ q.w) << 12);
            } else {
                for (int j = 0; i0 + j < n; j++) dm |= (uint32_t)(lab[i0 + j] != lab_m[i0 + j]) << j;
            }
        }
        const int c = __popc(dm);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(FULL, incl, 31);
        if (!total) continue;
        int base = 0;
        if (lane == 31) base = atomicAdd(count, total);
        base = __shfl_sync(FULL, base, 31);
        int pos = base + incl - c;
        while (dm) {
            list[pos++] = i0 + __ffs(dm) - 1;
            dm &= dm - 1;
        }
    }
}

// =============================================================================================
template <int KT>
__global__ void __launch_bounds__(PK_THREADS, 2)
k_em_persist(const nemk_persist_args a) {
    __shared__ double sh[32];
    __shared__ float nkf[NEMB_MAX_K];
    __shared__ double nkd[NEMB_MAX_K];
    __shared__ float s_w[(PK_THREADS / 32) * COOP_CHUNK];
    __shared__ uint8_t s_l[(PK_THREADS / 32) * COOP_CHUNK];
    const int K = a.K, n = a.n, D = a.D;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const unsigned nb = gridDim.x;
    nemk_lpsrc lps;
    lps.logpf = nullptr; lps.ham = a.ham; lps.coef = a.coef; lps.wsum_any_order = a.wsum_any_order;
    const bool use_graph = a.use_graph != 0, seq = a.seq_sweep != 0;
    const bool margins = seq && a.use_margins;
    int state = a.entry, it = a.iter0, cur = a.cur, stale_par = a.stale_par;
    int stats_valid = a.stats_valid, last_changed = a.last_changed, margins_on = a.margins_on;
    int barriers = 0, sweeps = 0, x_passes = 0, recounts = 0, nfix = 0, cnt_par = a.cnt_par;
    long long kept = 0;
    int exit_code = NEMK_PK_EXIT_DONE, resume = NEMK_PK_ENTRY_MSTEP, converged = 0, empty = 0;
    int n_allnul = 0, n_ties = 0;
#define PK_SYNC() do { pk_grid_sync(a.bar, nb); barriers++; } while (0)

    // the entry codes plus the two initial sweeps as states of their own, so that the sweep has
    // ONE (inlined) call site
    enum { S_INIT = NEMK_PK_ENTRY_INIT, S_BLIND = NEMK_PK_ENTRY_INIT_SWEEPS, S_MSTEP = NEMK_PK_ENTRY_MSTEP,
           S_FINALIZE = NEMK_PK_ENTRY_FINALIZE, S_SWEEP = NEMK_PK_ENTRY_SWEEP, S_BETA0 = 5 };
    for (;;) {
        if (state == S_INIT) {
            // ---- prep: unlabelled state (the reference's calloc'd ClassifM), clean flags, presets
            for (int i = gtid; i < (n + 3) / 4; i += nthreads) {
                reinterpret_cast<uint32_t *>(a.lab[0])[i] = 0xffffffffu;
                reinterpret_cast<uint32_t *>(a.stale[0])[i] = 0u;
                reinterpret_cast<uint32_t *>(a.stale[1])[i] = 0u;
            }
            if (gtid == 0) {
                a.coef->uniform_ok = 1; a.coef->mu_changed = 1; a.coef->empty_class = 0; a.coef->halt = 0;
                // zero at rest from here on (a launch-per-stage fit on the same handle does not keep
                // the delta-list counter clean)
                for (int q = 0; q < 8; q++) a.wl_cnt[q] = 0;
                a.scratch[0] = 0; a.scratch[1] = 0;
            }
            PK_SYNC();
            for (int k = blockIdx.x; k < K; k += gridDim.x) pk_tables_class<PK_THREADS>(k, a, sh);
            PK_SYNC();
            if (a.init_from_pop) {
                // every class has a constant centre: H = P, D - P or 0 (k_ham_from_pop)
                for (int row = gtid; row < n; row += nthreads) {
                    const int P = a.pop[row];
#pragma unroll
                    for (int k = 0; k < KT; k++) {
                        if (k < K) {
                            const int kind = a.coef->kind[k];
                            a.ham[(size_t)row * K + k] = kind == 1 ? P : kind == 2 ? D - P : 0;
                        }
                    }
                }
                PK_SYNC();
            } else if (a.x_in_kernel) {
                pk_density_pass<KT>(a);
                x_passes++;
                PK_SYNC();
            } else {
                exit_code = NEMK_PK_EXIT_NEED_DENSITY; resume = NEMK_PK_ENTRY_INIT_SWEEPS;
                break;
            }
            state = S_BLIND;
        } else if (state == S_MSTEP) {
            if (it >= a.it_max) break;
            const bool incremental = stats_valid && last_changed >= 0 && last_changed <= n / 8;
            if (incremental) {
                // statistics describe lab[cur ^ 1] (the input of the last sweep)
                if (gtid == 0) { a.coef->uniform_ok = 1; a.coef->mu_changed = 0; }
                pk_scan_changed(n, a.lab[cur], a.lab[cur ^ 1], a.wl[1], &a.wl_cnt[4]);
                PK_SYNC();
                const int total = *(volatile int32_t *)&a.wl_cnt[4];
                mstep_delta_items<KT>(K, D, a.wpr, a.x, a.lab[cur], a.lab[cur ^ 1], a.wl[1], total,
                                      a.stat, a.stat + (size_t)K * D);
                PK_SYNC();
                if (gtid == 0) a.wl_cnt[4] = 0;   // consumed
            } else if (a.x_in_kernel) {
                for (int q = gtid; q < K * D + K; q += nthreads) a.stat[q] = 0;
                if (gtid == 0) { a.coef->uniform_ok = 1; a.coef->mu_changed = 0; }
                PK_SYNC();
                pk_label_masks<KT>(a, a.lab[cur]);
                PK_SYNC();
                pk_recount<KT>(a);
                recounts++;
                PK_SYNC();
            } else {
                exit_code = NEMK_PK_EXIT_NEED_RECOUNT; resume = NEMK_PK_ENTRY_FINALIZE;
                break;
            }
            state = S_FINALIZE;
        } else if (state == S_FINALIZE) {
            stats_valid = 1;
            for (int k = blockIdx.x; k < K; k += gridDim.x) pk_finalize_class<PK_THREADS>(k, a, sh, nkf, nkd);
            PK_SYNC();
            const volatile nemk_coef *vc = a.coef;
            empty = vc->empty_class;
            if (empty) {   // nem_alg.c:1831-1838: the E-step is not run, the loop ends
                it++;
                break;
            }
            state = S_SWEEP;
            if (vc->mu_changed) {
                if (a.x_in_kernel) {
                    pk_density_pass<KT>(a);
                    x_passes++;
                    PK_SYNC();
                } else {
                    exit_code = NEMK_PK_EXIT_NEED_DENSITY; resume = NEMK_PK_ENTRY_SWEEP;
                    break;
                }
            }
        } else {
            // ---- one E-step sweep: S_BLIND (beta = 0, ComputePartitionFromPara's first sweep),
            // S_BETA0 (the beta sweep that follows it) or S_SWEEP (the sweep of an EM iteration).
            // Two counter blocks alternate: a sweep clears the block of the NEXT one (cnt_par
            // survives the launch, so a re-entry finds its block clean).
            const bool blind = state == S_BLIND;
            nemk_counters *cnt = a.cnt2 + cnt_par, *cnt_next = a.cnt2 + (cnt_par ^ 1);
            cnt_par ^= 1;
            nemk_margins mg;
            mg.m = nullptr; mg.stale_cur = nullptr; mg.stale_next = nullptr; mg.on = 0;
            if (margins && !blind) {
                mg.m = a.margin; mg.stale_cur = a.stale[stale_par]; mg.stale_next = a.stale[stale_par ^ 1];
                mg.on = state == S_BETA0 ? 0 : margins_on;
                stale_par ^= 1;
            }
            const int ch = pk_sweep<KT>(a, lps, blind ? 0.0 : a.beta, use_graph && !blind, seq && !blind,
                                        a.lab[cur], a.lab[cur ^ 1], mg, cnt, cnt_next, s_w, s_l, barriers,
                                        kept, nfix);
            const volatile nemk_counters *vcn = cnt;
            n_allnul = vcn->allnul; n_ties = vcn->ties;
            cur ^= 1;
            sweeps++;
            if (blind) { state = S_BETA0; continue; }
            last_changed = ch;
            margins_on = 1;   // the next sweep uses the same beta
            if (state == S_SWEEP) {
                it++;
                if (a.conv == 1) {   // HasConverged `clas` under ncem: no label changed
                    const float md = ch ? 1.0f : 0.0f;
                    if (md < a.conv_thr) { converged = 1; break; }
                }
            }
            state = S_MSTEP;
        }
    }
    if (gtid == 0) {
        nemk_persist_out *o = a.out;
        o->exit_code = exit_code; o->resume_entry = resume;
        o->iters = it; o->converged = converged; o->empty_class = empty;
        o->cnt_par = cnt_par;
        o->cur = cur; o->stale_par = stale_par; o->last_changed = last_changed; o->stats_valid = stats_valid;
        o->n_allnul = n_allnul; o->n_ties = n_ties;
        o->sweeps = sweeps; o->x_passes = x_passes; o->recounts = recounts; o->barriers = barriers;
        o->kept = kept; o->fixup_rounds = nfix;
        __threadfence_system();
        *(volatile unsigned long long *)&o->seq = a.seq;
    }
#undef PK_SYNC
}
