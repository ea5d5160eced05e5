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
//   iteration [changed rows: scan + statistics update]  or  [zero | class masks | X^T recount]
//             | closed forms + tables (one CTA per class) | X pass when the class masks moved
//             | margin test + evaluation of what is left (Jacobi) | fix-up rounds ... | scan
//
// Every CTA takes the same decisions from the same device counters (read after a barrier), so
// there is no host round trip inside a fit.  Problems whose X does not fit the L2 leave the kernel
// for the two HBM-bound passes (TMA density kernel, X^T recount) and re-enter it: exit codes
// NEMK_PK_EXIT_NEED_DENSITY / NEED_RECOUNT.
//
// Memory visibility: data another CTA wrote in an earlier phase is read with ordinary loads after
// the barrier (fence + atomic arrival + fence, the cooperative-groups grid-sync pattern; the
// gpu-scope fence also drops the SM's L1); nothing mutable is read through __ldg / const
// __restrict__ kernel parameters here (the kernel takes ONE by-value struct, so the compiler cannot
// infer non-coherent loads), and labels that move INSIDE a fix-up round are read with __ldcg.
//
// Fix-up rounds without flags or fences (measured: a gpu-scope fence per site costs more than the
// evaluation, profiles/r2_pk_phases_*.txt).  The in-place index-order sweep (UPDATE_SEQ,
// nem_alg.c:2378-2383) is the unique solution of cur_i = F_i(cur_j, j<i; old_j, j>=i).  Round 0
// evaluates F with old labels everywhere; every evaluation that CHANGES a label appends all later
// readers of the site to the next round's list -- no de-duplication, a site may be listed (and
// evaluated) several times in a round.  Invariant after the barrier that ends round r: a site is
// either in the next list or none of its lower inputs moved during round r, in which case every
// evaluation of it in round r saw the same inputs and left the same label, F_i(current inputs).
// An empty list therefore means the fixed point = the sequential sweep's labels, whatever the
// interleaving; a racing duplicate can only cause a redundant evaluation, never a lost update,
// because the site whose store caused the race listed the reader again.  The number of labels that
// differ from the sweep's input is counted afterwards by the scan that also feeds the M-step.

#define PK_THREADS 512
#ifndef PK_TRACK
#define PK_TRACK 0
#endif
#define PK_TRACK_MAX 4096      // at most this many labels moved last sweep: the statistics follow the
                               // label moves of this sweep at once (no scan / update phases)
// PK_CHASE > 1 would let a fix-up evaluation that moved a label go straight on with the first later
// reader instead of queueing it.  NOT exact (GPU test, round 2): another evaluation of that reader
// that started before the move can store its stale label last, and nobody re-queues the reader.
// Kept at 1 (= off) as a warning, like the note in nem_kernels.cu.
#ifndef PK_CHASE
#define PK_CHASE 1
#endif

// device-wide barrier: one atomic per CTA on a counter whose top bit flips when all have arrived
// (the cooperative-groups scheme; 1.2 us for 296 CTAs on B200, profiles/pk_barrier_bench.cu)
static __device__ __forceinline__ void pk_grid_sync(unsigned *bar, unsigned nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned add = blockIdx.x == 0 ? 0x80000000u - (nblocks - 1u) : 1u;
        __threadfence();                // release: this CTA's stores of the phase
        const unsigned old = atomicAdd(bar, add);
        while ((((old ^ *(volatile unsigned *)bar) & 0x80000000u) == 0u)) { }
        __threadfence();                // acquire: the other CTAs' stores (drops this SM's L1)
    }
    __syncthreads();
}

// ---- row shards: the exchange block of rank p (peer memory), typed views into it
template <typename T>
static __device__ __forceinline__ T *pk_peer(const nemk_persist_args &a, int p, long long off) {
    return reinterpret_cast<T *>(a.peer[p] + off);
}
// cross-rank barrier.  All CTAs of all ranks call it the same number of times (every decision that
// leads here is taken from numbers that are identical on every rank).  Local device-wide barrier
// (every thread that stored into a peer's block fenced at system scope before), then thread 0
// raises this rank's epoch flag in every peer's block and waits for theirs in its own, then a
// second local barrier releases the grid.  A peer that never arrives (dead process) ends the wait
// after ~4 s with *xerr set instead of hanging the GPU.
static __device__ void pk_xbarrier(const nemk_persist_args &a, unsigned &epoch, int &xerr) {
    pk_grid_sync(a.bar, gridDim.x);
    epoch++;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *(volatile int *)&a.scratch[20] = (int)epoch;       // heartbeat: barrier entered
        // (every thread that stored into a peer fenced at system scope itself, before the local
        // barrier above: nothing of this rank is still in flight.  The flag goes out as a
        // system-scope reduction -- a plain store may sit in a write buffer while this thread polls)
        for (int p = 0; p < a.world; p++)
            if (p != a.rank) atomicMax_system(pk_peer<unsigned>(a, p, a.off_xflag) + a.rank, epoch);
        volatile unsigned *mine = pk_peer<unsigned>(a, a.rank, a.off_xflag);
        const long long t0 = clock64();
        for (int q = 0; q < a.world; q++) {
            if (q == a.rank) continue;
            while ((int)(mine[q] - epoch) < 0) {
                if (clock64() - t0 > 8000000000ll) {     // who was missing and what its flag said (diagnostics)
                    *(volatile int *)&a.scratch[13] = 1 | (q << 4) | (int)((mine[q] & 0xfffffu) << 8);
                    break;
                }
            }
        }
        // what the peers stored lives in THIS rank's memory: dropping the SM's L1 (gpu-scope fence of
        // the barrier below) is enough to read it
        *(volatile int *)&a.scratch[21] = (int)epoch;       // heartbeat: barrier passed
    }
    pk_grid_sync(a.bar, gridDim.x);
    if (*(volatile int *)&a.scratch[13]) xerr = *(volatile int *)&a.scratch[13];
}
// label of own family i := km in the label buffer `buf` of every OTHER rank
static __device__ __forceinline__ void pk_push_label(const nemk_persist_args &a, int buf, int i, int km) {
    for (int p = 0; p < a.world; p++)
        if (p != a.rank) pk_peer<uint8_t>(a, p, a.off_lab[buf])[i] = (uint8_t)km;
    __threadfence_system();
}
// family j of a LOWER rank read the old label of a family that moved: its cached margin is void for
// its owner's next sweep
static __device__ __forceinline__ void pk_push_stale(const nemk_persist_args &a, int buf, int j) {
    pk_peer<uint8_t>(a, j / a.shard_len, a.off_stale[buf])[j] = 1;
    __threadfence_system();
}
// family j, owned by another rank, must be re-evaluated there: into that rank's inbox segment of
// this rank (parity = the super-round being built); counted locally, published at the barrier
static __device__ __forceinline__ void pk_push_remote(const nemk_persist_args &a, int par, int j) {
    const int owner = j / a.shard_len;
    const int slot = atomicAdd(pk_peer<int32_t>(a, a.rank, a.off_incnt) + par * NEMK_PK_MAX_WORLD + owner, 1);
    if (slot < a.xcap) {
        pk_peer<int32_t>(a, owner, a.off_inbox)[((size_t)par * a.world + a.rank) * a.xcap + slot] = j;
        __threadfence_system();
    } else
        a.scratch[12] = 1;      // overflow: every rank then re-evaluates all its families
}

// phase timer of CTA 0 / thread 0 (nemk_persist_out.phase_ns)
// (SM cycle counter: CTA 0 never migrates; a %globaltimer read is far slower and would sit on the
// critical path of every phase; the host converts with the SM clock rate)
static __device__ __forceinline__ unsigned long long pk_now() { return (unsigned long long)clock64(); }
struct PkProf { unsigned long long t_last, ns[12]; long long *trace; };   // trace: row of the current iteration or NULL
static __device__ __forceinline__ int pk_trace_col(int idx) {   // phase index -> trace column
    return idx == 1 ? 0 : (idx == 2 || idx == 3) ? 1 : idx == 4 ? 2 : idx == 6 ? 3 : (idx == 7 || idx == 8) ? 4 : idx == 9 ? 5 : -1;
}
#define PK_MARK(P, IDX) do { if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t_ = pk_now(); \
    (P).ns[IDX] += t_ - (P).t_last; \
    if ((P).trace && pk_trace_col(IDX) >= 0) (P).trace[pk_trace_col(IDX)] += (long long)(t_ - (P).t_last); \
    (P).t_last = t_; } } while (0)

// where the evaluations of a sweep append the sites that must be (re-)evaluated in the next round
struct PkNext { int32_t *list, *cnt, *ovf; int cap; const nemk_persist_args *a; int lo, hi, par, stale_buf; };   // [lo, hi) own families, par: inbox parity, stale_buf: index of stale_next

// label of site i moved: every later reader goes to the next round (no de-duplication), earlier-or-
// equal readers keep this sweep's evaluation, which saw the OLD label: their cached margin is void
static __device__ __forceinline__ void pk_push_readers(int i, const int32_t *rrow_ptr, const int32_t *rcol,
                                                       const PkNext &nx, uint8_t *stale_next,
                                                       int *chase = nullptr) {
    // chase (nullable, in: -1): receives the FIRST later reader in row order, which is then not
    // appended -- the caller goes on with it in the same round
    const int lo = rrow_ptr[i], hi = rrow_ptr[i + 1];
    for (int e = lo; e < hi; e += 8) {
        int j[8], nl = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) j[q] = e + q < hi ? rcol[e + q] : -1;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            if (j[q] < 0) continue;
            if (j[q] <= i) {
                if (stale_next) {
                    if (j[q] >= nx.lo) stale_next[j[q]] = 1;
                    else pk_push_stale(*nx.a, nx.stale_buf, j[q]);     // a lower rank's family
                }
                j[q] = -1;
            }
            else if (j[q] >= nx.hi) { pk_push_remote(*nx.a, nx.par, j[q]); j[q] = -1; }   // another rank's family
            else if (chase && *chase < 0) { *chase = j[q]; j[q] = -1; }
            else nl++;
        }
        if (nl) {
            int base = atomicAdd(nx.cnt, nl);
            if (base + nl > nx.cap) { *nx.ovf = 1; continue; }   // the next round then takes every site
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (j[q] >= 0) nx.list[base++] = j[q];
        }
    }
}
static __device__ __forceinline__ void pk_push_readers_warp(int i, const int32_t *rrow_ptr, const int32_t *rcol,
                                                            const PkNext &nx, uint8_t *stale_next) {
    const int lane = threadIdx.x & 31;
    const int lo = rrow_ptr[i], hi = rrow_ptr[i + 1];
    for (int e0 = lo; e0 < hi; e0 += 32) {
        const int e = e0 + lane;
        int j = e < hi ? rcol[e] : -1;
        if (j >= 0 && j <= i && stale_next) {
            if (j >= nx.lo) stale_next[j] = 1;
            else pk_push_stale(*nx.a, nx.stale_buf, j);
        }
        if (j >= nx.hi) { pk_push_remote(*nx.a, nx.par, j); j = -1; }
        const unsigned later = __ballot_sync(FULL, j > i);
        if (!later) continue;
        int base = 0;
        if (lane == 0) base = atomicAdd(nx.cnt, __popc(later));
        base = __shfl_sync(FULL, base, 0);
        if (base + __popc(later) > nx.cap) { if (lane == 0) *nx.ovf = 1; continue; }
        if (j > i) nx.list[base + __popc(later & ((1u << lane) - 1u))] = j;
    }
}

// label byte i := km, returns the label it REPLACED (CAS on the aligned word): with duplicates racing
// in a fix-up round, the chain of replaced labels telescopes, so statistics moved "from the replaced
// label to the new one" stay exact
static __device__ __forceinline__ int pk_xchg_label(uint8_t *lab, int i, int km) {
    unsigned *w = reinterpret_cast<unsigned *>(lab + (i & ~3));
    const int sh = (i & 3) * 8;
    unsigned old = __ldcg(w), assumed;
    do {
        assumed = old;
        old = atomicCAS(w, assumed, (assumed & ~(0xffu << sh)) | ((unsigned)km << sh));
    } while (old != assumed);
    return (int)((old >> sh) & 0xffu);
}

// family `row` leaves class `from` for class `to`: its bits move between the integer statistics
// S[from], S[to] and n (exact, order-free).  By a whole warp.
static __device__ __forceinline__ void pk_move_row(const nemk_persist_args &a, int row, int from, int to) {
    const int lane = threadIdx.x & 31, D = a.D, wreal = (D + 31) >> 5;
    int32_t *S = a.stat, *nk = a.stat + (size_t)a.K * D;
    if (lane == 0) { atomicAdd(&nk[from], -1); atomicAdd(&nk[to], 1); }
    const uint32_t *xr = a.x + (size_t)row * a.wpr;
    for (int w = lane; w < wreal; w += 32) {
        uint32_t bits = __ldg(xr + w);
        while (bits) {
            const int d = w * 32 + __ffs(bits) - 1;
            bits &= bits - 1;
            atomicAdd(&S[(size_t)from * D + d], -1);
            atomicAdd(&S[(size_t)to * D + d], 1);
        }
    }
}
// the rows whose lanes hold chg != 0 move from `from` to `to` (all lanes must call)
static __device__ __forceinline__ void pk_move_rows_warp(const nemk_persist_args &a, int chg, int row,
                                                        int from, int to) {
    unsigned cm = __ballot_sync(FULL, chg != 0);
    while (cm) {
        const int src = __ffs(cm) - 1;
        cm &= cm - 1;
        pk_move_row(a, __shfl_sync(FULL, row, src), __shfl_sync(FULL, from, src), __shfl_sync(FULL, to, src));
    }
}

// ---- one site of the Jacobi round (every input is an OLD label), new label to lab_out
template <int KT>
static __device__ __forceinline__ int pk_jac_site(int K, int i, const nemk_lpsrc &lps, const int32_t *rp,
                                                   const int32_t *col, const float *wgt, double beta,
                                                   const uint8_t *lab_in, uint8_t *lab_out, bool seq,
                                                   const int32_t *rrow_ptr, const int32_t *rcol,
                                                   const PkNext &nx, const nemk_margins &mg,
                                                   double thr_store, uint8_t *evflag) {
    double ctx[KT], lpv[KT], margin;
    load_lp<KT>(lps, K, (size_t)i, lpv);
    const int lin = (int)lab_in[i];
    ctx_labels<KT>(K, i, rp, col, wgt, [&](int j) { return (unsigned)lab_in[j]; }, ctx);
    int fl;
    const int km = site_argmax<KT>(K, lpv, ctx, beta, fl, margin);
    lab_out[i] = (uint8_t)km;
    evflag[i] = (uint8_t)fl;
    store_margin(mg, i, margin, thr_store);
    if (mg.m) mg.stale_cur[i] = 0;
    if (seq && km != lin) pk_push_readers(i, rrow_ptr, rcol, nx, mg.stale_next);
    return km != lin ? km + 1 : 0;      // 0 = unchanged, else new label + 1
}

// a hub of the Jacobi round, by a whole warp (update=para only: the sequential sweep defers its hubs
// to the first fix-up round, which spreads them one per warp)
template <int KT>
static __device__ __forceinline__ void pk_jac_hub(int K, int i, const nemk_lpsrc &lps, const int32_t *rp,
                                                  const int32_t *col, const float *wgt, double beta,
                                                  const uint8_t *lab_in, uint8_t *lab_out, uint8_t *evflag) {
    double ctx[KT], lpv[KT], margin;
    load_lp<KT>(lps, K, (size_t)i, lpv);
    ctx_labels_warp<KT>(K, i, rp, col, wgt, [&](int j) { return (unsigned)lab_in[j]; }, ctx,
                        lps.wsum_any_order != 0);
    int fl;
    const int km = site_argmax<KT>(K, lpv, ctx, beta, fl, margin);
    if ((threadIdx.x & 31) == 0) { lab_out[i] = (uint8_t)km; evflag[i] = (uint8_t)fl; }
}

// ---- re-evaluation of site i in a fix-up round: lower inputs are CURRENT labels (they move during
// the round: __ldcg), the others old.  Returns 0 when the label stands, else new label + 1;
// `replaced` = the label the store replaced (track: exchanged, so racing duplicates telescope),
// `chase` = the first later reader (not queued: the caller goes on with it in this round).
template <int KT>
static __device__ __forceinline__ int pk_fix_site(int K, int i, const nemk_lpsrc &lps, const int32_t *rp,
                                                  const int32_t *col, const float *wgt, double beta,
                                                  const uint8_t *lab_old, uint8_t *lab_cur,
                                                  const int32_t *rrow_ptr, const int32_t *rcol,
                                                  const PkNext &nx, const nemk_margins &mg,
                                                  double thr_store, uint8_t *evflag, bool track,
                                                  int &replaced, int &chase) {
    double ctx[KT], lpv[KT], margin;
    const int was = (int)__ldcg(lab_cur + i);
    load_lp<KT>(lps, K, (size_t)i, lpv);
    ctx_labels<KT>(K, i, rp, col, wgt,
                   [&](int j) { return (unsigned)(j < i ? __ldcg(lab_cur + j) : lab_old[j]); }, ctx);
    int fl;
    const int km = site_argmax<KT>(K, lpv, ctx, beta, fl, margin);
    store_margin(mg, i, margin, thr_store);     // the LAST evaluation of a site is its final one
    evflag[i] = (uint8_t)fl;
    chase = -1;      // (no chasing: see PK_CHASE above -- every later reader is queued)
    if (km == was) return 0;
    if (track) replaced = pk_xchg_label(lab_cur, i, km);
    else { lab_cur[i] = (uint8_t)km; replaced = was; }
    pk_push_readers(i, rrow_ptr, rcol, nx, mg.stale_next);
    return km + 1;
}
// a hub, by a whole warp; statistics and the net changed count included
template <int KT>
static __device__ __forceinline__ void pk_fix_hub(const nemk_persist_args &a, int K, int i, const nemk_lpsrc &lps,
                                                  const int32_t *rp, const int32_t *col, const float *wgt,
                                                  double beta, const uint8_t *lab_old, uint8_t *lab_cur,
                                                  int out_buf, const int32_t *rrow_ptr, const int32_t *rcol,
                                                  const PkNext &nx, const nemk_margins &mg,
                                                  double thr_store, uint8_t *evflag, bool track) {
    const int lane = threadIdx.x & 31;
    double ctx[KT], lpv[KT], margin;
    int was = (int)__ldcg(lab_cur + i);
    load_lp<KT>(lps, K, (size_t)i, lpv);
    ctx_labels_warp<KT>(K, i, rp, col, wgt,
                        [&](int j) { return (unsigned)(j < i ? __ldcg(lab_cur + j) : lab_old[j]); }, ctx,
                        lps.wsum_any_order != 0);
    int fl;
    const int km = site_argmax<KT>(K, lpv, ctx, beta, fl, margin);
    was = __shfl_sync(FULL, was, 0);
    int replaced = was;
    if (lane == 0) {
        store_margin(mg, i, margin, thr_store);
        evflag[i] = (uint8_t)fl;
        if (km != was) {
            if (track) replaced = pk_xchg_label(lab_cur, i, km);
            else lab_cur[i] = (uint8_t)km;
        }
    }
    if (km == was) return;
    if (a.world > 1 && lane == 0) pk_push_label(a, out_buf, i, km);
    pk_push_readers_warp(i, rrow_ptr, rcol, nx, mg.stale_next);
    if (track) {
        replaced = __shfl_sync(FULL, replaced, 0);
        if (replaced != km) {
            pk_move_row(a, i, replaced, km);
            const int old = (int)lab_old[i];
            const int dn = (km != old) - (replaced != old);
            if (lane == 0 && dn) atomicAdd(&a.scratch[6], dn);
        }
    }
}

// ---- one E-step sweep (ComputePartitionNEM nem_alg.c:2330-2405 for ncem).  All CTAs call it.
// use_graph: context term on; seq: in-place index-order semantics through the speculative fixed
// point; margins (mg.m != NULL) only with seq.  cnt->kept receives the sites the margin cache
// saved.  The caller counts the changed labels afterwards (pk_scan).
// Row shards: this rank evaluates its own families [row0, row0 + n); a label that moves is stored
// into every peer's copy of the output buffer at once, a later reader owned by another rank is
// queued in that rank's inbox.  A rank runs its fix-up rounds to exhaustion with LOCAL barriers
// only; the ranks then meet (one cross-rank barrier), and as long as some inbox is not empty they
// start another cascade from their inboxes.  Any interleaving reaches the same fixed point (header).
template <int KT>
static __device__ bool pk_sweep(const nemk_persist_args &a, const nemk_lpsrc &lps, double beta,
                               bool use_graph, bool seq, const uint8_t *lab_in, uint8_t *lab_out,
                               int out_buf, nemk_margins mg, int stale_next_buf, nemk_counters *cnt,
                               nemk_counters *cnt_next, float *s_w, uint8_t *s_l, int &barriers,
                               long long &kept_total, int &nfix_total, PkProf &prof, bool track_ok,
                               int mu_changed, unsigned &xepoch, int &xerr) {
    const int K = a.K, n = a.n, lo_i = a.row0, hi_i = a.row0 + a.n;
    const bool sharded = a.world > 1;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nthreads = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nwarps = nthreads >> 5, gwarp = gtid >> 5;
    const int32_t *rp = use_graph ? a.row_ptr : nullptr;
    int32_t *wl_cnt = a.wl_cnt, *ovf = a.scratch + 8;
    // (sweep_thr with the kernel's own copy of mu_changed: coef->mu_changed is only kept for the
    // kernels that run outside, and nothing resets it between iterations here)
    SweepThr thr;
    thr.test = CUDART_INF; thr.store = 0.0;
    if (mg.m && mg.on && !mu_changed) {
        double step = 0.0;
        for (int k = 0; k < K; k++) step = fmax(step, lps.coef->dstep[k]);
        if (step < CUDART_INF) { thr.store = lps.coef->drift + 2.0 * step; thr.test = thr.store + 1e-6; }
    }
    const bool may_skip = mg.m && thr.test < CUDART_INF;
    const bool hubs = rp && a.n_heavy > 0;
    // track: the integer statistics S, n (which describe lab_in when the sweep starts) follow every
    // label move of this sweep at once (compile-time knob PK_TRACK; one GPU only)
    const bool track = PK_TRACK && track_ok && may_skip && !sharded;
    int par = (int)(xepoch & 1u);      // inbox parity of the requests queued until the next cross-rank barrier
    const PkNext nx0 = {a.wlist[0], &wl_cnt[0], &ovf[0], a.wl_cap, &a, lo_i, hi_i, par, stale_next_buf};
    int kept = 0;
    // counters of the last scan: every CTA read them right after the scan's barrier and has passed
    // another barrier since (closed forms); the next scan starts after this sweep's barriers
    if (gtid == 0) { wl_cnt[4] = 0; a.scratch[3] = 0; a.scratch[4] = 0; a.scratch[6] = 0; a.scratch[7] = track; }

    if (may_skip) {
        // ---- phase: margin test + evaluation, one warp per 128 consecutive sites.  A site whose
        // stored margin exceeds what theta can have moved and none of whose later-or-equal
        // neighbours changed in the previous sweep keeps its label (6 streamed bytes per site);
        // the warp compacts what is left and evaluates it at once (no list, no barrier, and ONE
        // visitor per site).  Active hubs are deferred to the first fix-up round.
        int *sm_act = reinterpret_cast<int *>(s_w) + wib * COOP_CHUNK;
        const int ngroups = (n + 3) >> 2;
        int nact = 0;
        for (int g0 = gwarp * 32; g0 < ngroups; g0 += nwarps * 32) {
            const int g = g0 + lane, i0 = lo_i + g * 4;
            unsigned actm = 0u, hubm = 0u;
            if (g < ngroups) {
                uint8_t st[4], lb[4];
                float mv[4];
                const int nv = min(4, hi_i - i0);
                if (nv == 4) {
                    const uchar4 s4 = *reinterpret_cast<const uchar4 *>(mg.stale_cur + i0);
                    const uchar4 l4 = *reinterpret_cast<const uchar4 *>(lab_in + i0);
                    const float4 m4 = *reinterpret_cast<const float4 *>(mg.m + i0);
                    st[0] = s4.x; st[1] = s4.y; st[2] = s4.z; st[3] = s4.w;
                    lb[0] = l4.x; lb[1] = l4.y; lb[2] = l4.z; lb[3] = l4.w;
                    mv[0] = m4.x; mv[1] = m4.y; mv[2] = m4.z; mv[3] = m4.w;
                } else {
                    for (int q = 0; q < 4; q++) {
                        const bool v = q < nv;
                        st[q] = v ? mg.stale_cur[i0 + q] : (uint8_t)0;
                        lb[q] = v ? lab_in[i0 + q] : (uint8_t)0;
                        mv[q] = v ? mg.m[i0 + q] : CUDART_INF_F;
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const bool keep = !st[q] && (double)mv[q] > thr.test && lb[q] != 255;
                    if (q < nv && !keep) actm |= 1u << q;
                }
                kept += nv - __popc(actm);
                if (actm && hubs) {
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (((actm >> q) & 1u) && rp[i0 + q + 1] - rp[i0 + q] > HEAVY_DEG) hubm |= 1u << q;
                }
                // labels of the kept sites and of the deferred hubs (old label until their round);
                // the sites evaluated below are written by their evaluation only
                const unsigned wr = (~actm | hubm) & ((1u << nv) - 1u);
                if (wr == 0xfu) *reinterpret_cast<uchar4 *>(lab_out + i0) = make_uchar4(lb[0], lb[1], lb[2], lb[3]);
                else {
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if ((wr >> q) & 1u) lab_out[i0 + q] = lb[q];
                }
                if (hubm) {
                    int base = atomicAdd(nx0.cnt, __popc(hubm));
                    if (base + __popc(hubm) > nx0.cap) *nx0.ovf = 1;
                    else {
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            if ((hubm >> q) & 1u) { nx0.list[base++] = i0 + q; mg.stale_cur[i0 + q] = 0; }
                    }
                }
            }
            const unsigned lm = actm & ~hubm;
            const int c = __popc(lm);
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            const int total = __shfl_sync(FULL, incl, 31);
            if (!total) continue;
            int pos = incl - c;
            unsigned m = lm;
            while (m) {
                const int q = __ffs(m) - 1;
                m &= m - 1;
                sm_act[pos++] = i0 + q;
            }
            __syncwarp();
            for (int r0 = 0; r0 < total; r0 += 32) {
                const int r = r0 + lane, site = r < total ? sm_act[r] : -1;
                int res = 0;
                if (site >= 0) {
                    res = pk_jac_site<KT>(K, site, lps, rp, a.col, a.wgt, beta, lab_in, lab_out, seq, a.rrow_ptr,
                                          a.rcol, nx0, mg, thr.store, a.evflag);
                    if (res && sharded) pk_push_label(a, out_buf, site, res - 1);
                }
                if (track) {
                    const int from = site >= 0 ? (int)lab_in[site] : 0;
                    pk_move_rows_warp(a, res, site, from, res - 1);
                    const int nch = __popc(__ballot_sync(FULL, res != 0));
                    if (lane == 0 && nch) atomicAdd(&a.scratch[6], nch);
                }
            }
            __syncwarp();
            nact += total;
        }
        if (lane == 0 && nact) atomicAdd(&a.scratch[5], nact);   // sites evaluated (trace)
    } else {
        // ---- phase: dense Jacobi round, every site evaluated.  Warps over 32 consecutive sites
        // whose contiguous CSR segment is streamed cooperatively (ctx_labels_coop).  Hubs: the
        // sequential sweep defers them to the first fix-up round (old label until then, one warp
        // each there); update=para evaluates them here, one warp each.
        if (gtid == 0 && prof.trace) prof.trace[6] = -1;
        if (hubs && seq) {
            for (int q0 = (gtid & ~31); q0 < a.n_heavy; q0 += nthreads) {
                const int q = q0 + lane;
                const bool v = q < a.n_heavy;
                const int i = v ? a.heavy[q] : 0;
                if (v) { lab_out[i] = lab_in[i]; if (mg.m) mg.stale_cur[i] = 0; }
                const unsigned bm = __ballot_sync(FULL, v);
                int base = 0;
                if (lane == 0) base = atomicAdd(nx0.cnt, __popc(bm));
                base = __shfl_sync(FULL, base, 0);
                if (base + __popc(bm) > nx0.cap) { if (lane == 0) *nx0.ovf = 1; }
                else if (v) nx0.list[base + lane] = i;
            }
        } else if (hubs) {
            for (int wi = gwarp; wi < a.n_heavy; wi += nwarps) {
                const int i = a.heavy[wi];
                pk_jac_hub<KT>(K, i, lps, rp, a.col, a.wgt, beta, lab_in, lab_out, a.evflag);
            }
        }
        for (int base = lo_i + gwarp * 32; base < hi_i; base += nwarps * 32) {
            const int i = base + lane;
            const bool valid = i < hi_i;
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
            int newlab = valid ? (int)lab_in[i] : 0;      // a deferred hub keeps its old label for now
            if (valid && !is_heavy) {
                double margin;
                int fl;
                const int km = site_argmax<KT>(K, lpv, ctx, beta, fl, margin);
                lab_out[i] = (uint8_t)km;
                a.evflag[i] = (uint8_t)fl;
                store_margin(mg, i, margin, thr.store);
                if (mg.m) mg.stale_cur[i] = 0;
                if (seq && km != newlab) pk_push_readers(i, a.rrow_ptr, a.rcol, nx0, mg.stale_next);
                newlab = km;
            }
            if (sharded) {
                // the 32 labels of the warp go to every peer as whole words (8 stores per peer)
                unsigned w = (unsigned)newlab & 0xffu;
                w |= (__shfl_down_sync(FULL, (unsigned)newlab & 0xffu, 1) << 8);
                w |= (__shfl_down_sync(FULL, (unsigned)newlab & 0xffu, 2) << 16);
                w |= (__shfl_down_sync(FULL, (unsigned)newlab & 0xffu, 3) << 24);
                if ((lane & 3) == 0 && valid) {
                    if (i + 4 <= hi_i) {
                        for (int p = 0; p < a.world; p++)
                            if (p != a.rank) *reinterpret_cast<unsigned *>(pk_peer<uint8_t>(a, p, a.off_lab[out_buf]) + i) = w;
                    } else {
                        for (int q = 0; i + q < hi_i; q++)
                            for (int p = 0; p < a.world; p++)
                                if (p != a.rank) pk_peer<uint8_t>(a, p, a.off_lab[out_buf])[i + q] = (uint8_t)(w >> (8 * q));
                    }
                    __threadfence_system();
                }
            }
        }
        if (sharded && hubs && !seq) {
            // update=para: the hubs were evaluated above by their warps; publish their labels
            pk_grid_sync(a.bar, gridDim.x); barriers++;
            for (int q = gtid; q < a.n_heavy; q += nthreads) pk_push_label(a, out_buf, a.heavy[q], lab_out[a.heavy[q]]);
        }
    }
    {
        const int kp = __reduce_add_sync(FULL, kept);
        if (lane == 0 && kp) atomicAdd(&cnt->kept, kp);
    }
    pk_grid_sync(a.bar, gridDim.x); barriers++;
    PK_MARK(prof, may_skip ? 7 : 8);
    if (gtid == 0) {
        // the counter block of the NEXT sweep: every CTA has passed a barrier since it last read it
        cnt_next->changed = 0; cnt_next->nfix = 0; cnt_next->allnul = 0; cnt_next->ties = 0;
        cnt_next->maxdiff = 0.f; cnt_next->pending = 0; cnt_next->changed_glob = 0; cnt_next->kept = 0;
        if (may_skip) { if (prof.trace) prof.trace[6] = a.scratch[5]; a.scratch[5] = 0; }
    }
    if (sharded && !may_skip) {
        // a dense round stored ALL labels into the peers: the fix-up rounds must not read a peer's
        // family before its label of this round has arrived (a margin-cached round needs no such
        // meeting: the peers' copies already hold the previous labels, and a label that moves
        // comes with a request to re-evaluate its readers)
        pk_xbarrier(a, xepoch, xerr); barriers += 2;
        if (xerr) return false;
    }

    // ---- fix-up rounds: the listed sites are re-evaluated (duplicates allowed) until no label
    // moves; one LOCAL barrier per round.  Round r reads list r&1 / counter r&3, appends to list
    // (r+1)&1 / counter (r+1)&3 and clears counter (r+2)&3 (idle during the round).  Items are dealt
    // round-robin over the warps so that a run of hubs lands on different warps.
    int rounds = 0;
    if (seq || sharded) {
        int round = 0, src_mode = 0, inbox_total = 0;    // src_mode: 0 local list, 1 inbox, 2 every own family
        int in_pref[NEMK_PK_MAX_WORLD + 1];
        for (int q = 0; q <= NEMK_PK_MAX_WORLD; q++) in_pref[q] = 0;
        for (;;) {   // cascades, separated by cross-rank barriers (one cascade on one GPU)
            for (;; round++) {
                const int32_t *cur_list = a.wlist[round & 1];
                const PkNext nx = {a.wlist[(round + 1) & 1], &wl_cnt[(round + 1) & 3], &ovf[(round + 1) & 3],
                                   a.wl_cap, &a, lo_i, hi_i, par, stale_next_buf};
                int count = 0;
                bool all = src_mode == 2;
                if (src_mode == 0) {
                    count = seq ? *(volatile int32_t *)&wl_cnt[round & 3] : 0;
                    all = *(volatile int32_t *)&ovf[round & 3] != 0;   // a list overflowed: every own family
                } else if (src_mode == 1)
                    count = inbox_total;
                if (gtid == 0) { wl_cnt[(round + 2) & 3] = 0; ovf[(round + 2) & 3] = 0; }
                const int items = all ? n : count;
                if (sharded && gtid == 0) {                 // heartbeat (read by the host if the fit hangs)
                    *(volatile int *)&a.scratch[16] = round; *(volatile int *)&a.scratch[17] = items;
                    *(volatile int *)&a.scratch[18] = all ? 2 : src_mode; *(volatile int *)&a.scratch[19] = (int)xepoch;
                }
                if (items == 0) break;
                rounds++;
                const int32_t *inbox = pk_peer<int32_t>(a, a.rank, a.off_inbox) + (size_t)(par ^ 1) * a.world * a.xcap;
                for (int base = gwarp; base < items; base += 32 * nwarps) {
                    const int idx = base + lane * nwarps;
                    int i = -1;
                    if (idx < items) {
                        if (all) i = lo_i + idx;
                        else if (src_mode == 1) {
                            int sseg = 0;
                            while (idx >= in_pref[sseg + 1]) sseg++;
                            i = inbox[(size_t)sseg * a.xcap + (idx - in_pref[sseg])];
                        } else
                            i = cur_list[idx];
                    }
                    // hubs by the whole warp
                    const bool hub = i >= 0 && (rp[i + 1] - rp[i] > HEAVY_DEG);
                    unsigned hm = __ballot_sync(FULL, hub);
                    while (hm) {
                        const int src = __ffs(hm) - 1;
                        hm &= hm - 1;
                        pk_fix_hub<KT>(a, K, __shfl_sync(FULL, i, src), lps, rp, a.col, a.wgt, beta, lab_in, lab_out,
                                       out_buf, a.rrow_ptr, a.rcol, nx, mg, thr.store, a.evflag, track);
                    }
                    if (hub) i = -1;
                    int res = 0, replaced = 0, chase = -1;
                    if (i >= 0) {
                        res = pk_fix_site<KT>(K, i, lps, rp, a.col, a.wgt, beta, lab_in, lab_out, a.rrow_ptr, a.rcol,
                                              nx, mg, thr.store, a.evflag, track, replaced, chase);
                        if (res && sharded) pk_push_label(a, out_buf, i, res - 1);
                    }
                    if (track) {
                        const int moved = res != 0 && replaced != res - 1;
                        pk_move_rows_warp(a, moved, i, replaced, res - 1);
                        if (moved) {
                            const int old = (int)lab_in[i];
                            const int dn = ((res - 1) != old) - (replaced != old);
                            if (dn) atomicAdd(&a.scratch[6], dn);
                        }
                    }
                }
                pk_grid_sync(a.bar, gridDim.x); barriers++;
                src_mode = 0;
            }
            if (!sharded) break;
            // ---- the ranks meet: requests sent to every rank (their inbox counts) and this rank's
            // total, stored into the peers' blocks; after the barrier every rank adds up the same
            // totals.  Nothing in flight anywhere = the sweep is over.
            // (counts stay in THIS rank's block -- row `par` of its out table, written by local
            // atomics during the cascade -- and the peers read them after the barrier)
            int32_t *my_out = pk_peer<int32_t>(a, a.rank, a.off_incnt) + par * NEMK_PK_MAX_WORLD;
            if (gtid == 0 && *(volatile int32_t *)&a.scratch[12]) { my_out[a.rank] = -1; a.scratch[12] = 0; }
            pk_xbarrier(a, xepoch, xerr);
            barriers += 2;
            if (xerr) break;
            // requests in flight anywhere (the same sum on every rank) and this rank's inbox counts:
            // read through peer memory by ONE thread per CTA, handed to the others in shared memory
            __shared__ int s_meet[NEMK_PK_MAX_WORLD + 2];
            if (threadIdx.x == 0) {
                int tot_ = 0, over_ = 0;
                for (int q = 0; q < a.world; q++) {
                    const int32_t *row = pk_peer<int32_t>(a, q, a.off_incnt) + par * NEMK_PK_MAX_WORLD;
                    int mine_ = 0;
                    for (int p = 0; p < a.world; p++) {
                        int c = __ldcv(row + p);
                        if (p == q) { if (c < 0) over_ = 1; continue; }
                        if (c > a.xcap) c = a.xcap;
                        tot_ += c;
                        if (p == a.rank) mine_ = c;
                    }
                    s_meet[q] = q == a.rank ? 0 : mine_;
                }
                s_meet[NEMK_PK_MAX_WORLD] = tot_; s_meet[NEMK_PK_MAX_WORLD + 1] = over_;
            }
            __syncthreads();
            const int total = s_meet[NEMK_PK_MAX_WORLD], over = s_meet[NEMK_PK_MAX_WORLD + 1];
            int inc[NEMK_PK_MAX_WORLD];
            for (int q = 0; q < NEMK_PK_MAX_WORLD; q++) inc[q] = q < a.world ? s_meet[q] : 0;
            __syncthreads();
            if (total == 0 && !over) break;
            in_pref[0] = 0;
            for (int q = 0; q < a.world; q++) in_pref[q + 1] = in_pref[q] + inc[q];
            for (int q = a.world; q < NEMK_PK_MAX_WORLD; q++) in_pref[q + 1] = in_pref[a.world];
            inbox_total = in_pref[a.world];
            src_mode = over ? 2 : (inbox_total ? 1 : 0);   // nothing for this rank: it still meets again
            // requests queued from now on go to the other inbox / counter row; that row was read by
            // the peers two barriers ago at the latest
            par ^= 1;
            if (gtid == 0)
                for (int q = 0; q < NEMK_PK_MAX_WORLD; q++) pk_peer<int32_t>(a, a.rank, a.off_incnt)[par * NEMK_PK_MAX_WORLD + q] = 0;
            pk_grid_sync(a.bar, gridDim.x); barriers++;
        }
        PK_MARK(prof, 9);
        if (gtid == 0) {
            prof.ns[10] += rounds;
            if (prof.trace) prof.trace[7] = rounds;
            // every CTA has read the last (zero) count before anybody can append again: the next
            // appends happen at least one barrier later (the next sweep's evaluation phase)
            for (int q = 0; q < 4; q++) { wl_cnt[q] = 0; ovf[q] = 0; }
            if (mg.m) const_cast<nemk_coef *>(lps.coef)->drift = thr.store;
        }
    }
    nfix_total += rounds;
    kept_total += ((const volatile nemk_counters *)cnt)->kept;
    return track;
}

// ---- after a sweep: the rows whose label differs from the sweep's input.  Counts them (the
// `clas` convergence test, nem_alg.c:2075-2089, and the size of the M-step's update), sums the
// all-null / tie flags of every site's last evaluation, and feeds the incremental statistics:
// mode 0 count only, 1 append the rows to `list` (the M-step updates S and n from the list after
// the barrier, 32 rows per ballot transpose).  16 labels per lane.  Sweeps that tracked their label
// moves (pk_sweep `track`) need no scan; their flags are summed once, when the kernel is left.
template <int KT>
static __device__ void pk_scan(const nemk_persist_args &a, const uint8_t *lab, const uint8_t *lab_m,
                               int mode, int32_t *list, int32_t *count, int fix_buf) {
    // row shards: own families only (row0 is a multiple of 16).  fix_buf >= 0: the peers' copies of
    // label buffer fix_buf (the one the NEXT sweep writes) still hold, for this rank's families, the
    // labels of one sweep ago: the rows that moved are brought up to date there.
    const int n = a.n, lo_i = a.row0;
    const int nthreads = gridDim.x * blockDim.x, lane = threadIdx.x & 31;
    const int ngroups = (n + 15) >> 4;
    int nul = 0, ties = 0;
    for (int t0 = ((blockIdx.x * blockDim.x + threadIdx.x) & ~31); t0 < ngroups; t0 += nthreads) {
        const int t = t0 + lane, i0 = lo_i + t * 16;
        uint32_t dm = 0u;
        if (t < ngroups) {
            if (i0 + 16 <= lo_i + n) {
                const uint4 p = *reinterpret_cast<const uint4 *>(lab + i0);
                const uint4 q = *reinterpret_cast<const uint4 *>(lab_m + i0);
                const uint4 f = *reinterpret_cast<const uint4 *>(a.evflag + i0);
                dm = nonzero_bytes(p.x ^ q.x) | (nonzero_bytes(p.y ^ q.y) << 4) |
                     (nonzero_bytes(p.z ^ q.z) << 8) | (nonzero_bytes(p.w ^ q.w) << 12);
                nul += __popc(f.x & 0x01010101u) + __popc(f.y & 0x01010101u) + __popc(f.z & 0x01010101u) +
                       __popc(f.w & 0x01010101u);
                ties += __popc(f.x & 0x02020202u) + __popc(f.y & 0x02020202u) + __popc(f.z & 0x02020202u) +
                        __popc(f.w & 0x02020202u);
            } else {
                for (int j = 0; i0 + j < lo_i + n; j++) {
                    dm |= (uint32_t)(lab[i0 + j] != lab_m[i0 + j]) << j;
                    nul += a.evflag[i0 + j] & 1;
                    ties += (a.evflag[i0 + j] >> 1) & 1;
                }
            }
        }
        if (fix_buf >= 0 && dm) {
            uint32_t m = dm;
            while (m) {
                const int i = i0 + __ffs(m) - 1;
                m &= m - 1;
                for (int p = 0; p < a.world; p++)
                    if (p != a.rank) pk_peer<uint8_t>(a, p, a.off_lab[fix_buf])[i] = lab[i];
            }
            __threadfence_system();
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
        if (mode == 1) {
            base = __shfl_sync(FULL, base, 31);
            int pos = base + incl - c;
            uint32_t m = dm;
            while (m) {
                list[pos++] = i0 + __ffs(m) - 1;
                m &= m - 1;
            }
        }
    }
    nul = __reduce_add_sync(FULL, nul);
    ties = __reduce_add_sync(FULL, ties);
    if (lane == 0) {
        if (nul) atomicAdd(&a.scratch[3], nul);
        if (ties) atomicAdd(&a.scratch[4], ties);
    }
}

// four block-wide double sums through ONE pair of block barriers (fixed shuffle tree + fixed
// shared-memory order: deterministic)
template <int TH>
static __device__ __forceinline__ void pk_block_sum4(double (&v)[4], double *sh /*[4 * 32]*/) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(FULL, v[q], o);
    __syncthreads();
    if (l == 0) {
#pragma unroll
        for (int q = 0; q < 4; q++) sh[q * 32 + w] = v[q];
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            double r = (l < TH / 32) ? sh[q * 32 + l] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(FULL, r, o);
            if (l == 0) sh[q * 32] = r;
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; q++) v[q] = sh[q * 32];
}

// ---- closed forms + tables of class k by ONE CTA (same float expressions as
// k_mstep_finalize_tables: nem_mod.c:455-465, 965-1174, 1422-1479, 1646-1704; DensBernoulli's terms
// nem_mod.c:656-670).  This runs every EM iteration while the rest of the grid waits at the
// barrier, so it is written for latency: the statistics of a thread's genomes are loaded together,
// the centres stay in registers as 2-bit codes between the two passes (nothing written is read
// back), the previous mask words are fetched up front (one pair per lane, handed to lane 0 by a
// shuffle) and the four table sums share one pair of block barriers.  A class without families, or
// more than 32 * TH genomes, takes the general code path.
template <int TH>
static __device__ void pk_finalize_class(int k, const nemk_persist_args &a, double *sh,
                                         float *nkf, double *nkd) {
    const int K = a.K, D = a.D, N = a.world > 1 ? a.n_glob : a.n, tid = threadIdx.x, lane = tid & 31, wpr = a.wpr;
    const int32_t *__restrict__ s_int = a.world > 1 ? a.stat_glob : a.stat;      // summed over the ranks
    const int32_t *__restrict__ nk_int = s_int + (size_t)K * D;
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
    const int niter = (D + TH - 1) / TH;           // genomes per thread
    const bool fast = nonempty && niter <= 32 && (a.disp_model == 0 || a.disp_model == 1);
    if (fast) {
        // previous mask words of this warp's iterations: lane q fetches the pair of iteration q
        uint32_t oldx = 0u, oldv = 0u;
        {
            const int j = lane * TH + (tid & ~31);
            if (lane < niter && j < D) {
                const size_t o = (size_t)k * wpr + (j >> 5);
                oldx = a.mxor[o]; oldv = a.mval[o];
            }
        }
        // pass A: centres (2-bit codes) and inertia, four genomes per thread in flight
        unsigned long long codes = 0ull;
        double v = 0.0, sn = 0.0;
        const double half = 0.5 * nkd[k];
        for (int q0 = 0; q0 < niter; q0 += 4) {
            double s[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int j = (q0 + q) * TH + tid;
                s[q] = (q0 + q < niter && j < D) ? (double)s_int[(size_t)k * D + j] : -1.0;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int j = (q0 + q) * TH + tid;
                if (s[q] < 0.0) continue;
                const unsigned code = s[q] > half ? 1u : (s[q] < half ? 0u : 2u);
                codes |= (unsigned long long)code << (2 * (q0 + q));
                center[(size_t)k * D + j] = code == 1u ? 1.0f : (code == 0u ? 0.0f : 0.5f);
                if (a.disp_model == 1) v += (double)iner_of(s[q], nkd[k], true, 0.f);
            }
        }
        if (a.disp_model == 1) sn = (double)nkf[k] * (double)D;
        else {
            // s__ pools the classes: every class's inertia (nem_mod.c:988-1015)
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
        const float dk = __fdiv_rn((float)v, (float)sn);
        if (tid == 0)
            a.prop[k] = a.prop_model == 1 ? __fdiv_rn(nkf[k], (float)N) : (float)(1.0 / (double)K);
        const ClassCoef cc = class_coef(dk);
        // pass B: dispersions, the per-genome table terms and the bit masks, from the codes
        double sums[4] = {0.0, 0.0, 0.0, 0.0};   // base_u, base_g, n_valid, n_x1
        int mu_moved = 0;
        for (int q = 0; q < niter; q++) {
            const int j = q * TH + tid;
            const bool in = j < D;
            const unsigned code = (unsigned)(codes >> (2 * q)) & 3u;
            const int m0 = code == 1u, m1 = code == 0u;       // abs((int)(0 - mu)), abs((int)(1 - mu))
            if (in) {
                disp[(size_t)k * D + j] = dk;
                const double cost0 = m0 * cc.a0 + cc.c0, cost1 = m1 * cc.a0 + cc.c0;
                a.delta[(size_t)k * D + j] = cost1 - cost0;
                sums[1] += cost0;
                sums[0] += cc.c0;
            }
            const unsigned bx = __ballot_sync(FULL, in && code == 1u);
            const unsigned bv = __ballot_sync(FULL, in && code != 2u);
            const unsigned b0 = __ballot_sync(FULL, in && !cc.live0 && m0 != 0);
            const unsigned b1 = __ballot_sync(FULL, in && !cc.live0 && m1 != 0);
            const uint32_t ox = __shfl_sync(FULL, oldx, q), ov = __shfl_sync(FULL, oldv, q);
            if (lane == 0 && (j >> 5) < wreal) {
                const size_t o = (size_t)k * wpr + (j >> 5);
                sums[2] += __popc(bv); sums[3] += __popc(bx);
                if (ox != bx || ov != bv) mu_moved = 1;
                a.mxor[o] = bx; a.mval[o] = bv; a.f0[o] = b0; a.f1[o] = b1;
            }
        }
        for (int w = wreal + tid; w < wpr; w += TH) {
            const size_t o = (size_t)k * wpr + w;
            a.mxor[o] = 0u; a.mval[o] = 0u; a.f0[o] = 0u; a.f1[o] = 0u;
        }
        const int any_moved = __syncthreads_or(mu_moved);
        pk_block_sum4<TH>(sums, sh);
        if (tid == 0) {
            a.coef->mu_moved_k[k] = any_moved;
            tables_commit(k, K, D, a.prop, a.coef, a.delta, cc, sums[0], sums[1], true, (int)sums[2], (int)sums[3], true);
        }
        return;
    }
    // ---- general path
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
    TablesPartial p = tables_words(k, D, wpr, 0, wreal, cc, center, disp, a.mxor, a.mval, a.f0, a.f1, a.delta);
    for (int w = wreal + tid; w < wpr; w += TH) {
        const size_t o = (size_t)k * wpr + w;
        a.mxor[o] = 0u; a.mval[o] = 0u; a.f0[o] = 0u; a.f1[o] = 0u;
    }
    const int any_moved = __syncthreads_or(p.mu_moved);
    double sums[4] = {p.base_u, p.base_g, (double)p.n_valid, (double)p.n_x1};
    const double notok = block_sum<TH>((double)p.notok, sh);
    pk_block_sum4<TH>(sums, sh);
    if (tid == 0) {
        a.coef->mu_moved_k[k] = any_moved;
        tables_commit(k, K, D, a.prop, a.coef, a.delta, cc, sums[0], sums[1], notok == 0.0, (int)sums[2], (int)sums[3], true);
    }
}

// tables of the theta the caller supplied (k_theta_tables), class k by one CTA
template <int TH>
static __device__ void pk_tables_class(int k, const nemk_persist_args &a, double *sh) {
    const ClassCoef cc = class_coef(a.disp[(size_t)k * a.D]);
    TablesPartial p = tables_words(k, a.D, a.wpr, 0, a.wpr, cc, a.center, a.disp, a.mxor, a.mval,
                                   a.f0, a.f1, a.delta);
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

// ---- final criteria (ComputeCrit, nem_alg.c:2678-2757: D, G, L, Z per family, U = D + beta/2 G,
// M = D + beta G + Z) from the hard labels, the densities of the last tables and the whole graph
// (even when beta = 0): k_criteria_partial's walk -- hubs by a warp, the others 32 consecutive
// sites per warp -- with per-CTA partial sums; CTA 0 adds them in CTA order after the barrier
// (fixed order: reproducible for a given grid).
template <int KT>
static __device__ void pk_criteria_partial(const nemk_persist_args &a, const nemk_lpsrc &lps,
                                           const uint8_t *lab, float *s_w, uint8_t *s_l, double *sh) {
    const int K = a.K, n = a.n, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nthreads = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nwarps = nthreads >> 5, gwarp = gtid >> 5;
    const int32_t *rp = a.spatial ? a.row_ptr : nullptr;
    const bool hubs = rp && a.n_heavy > 0;
    const double beta = a.beta;
    double c[4] = {0.0, 0.0, 0.0, 0.0};   // D G L Z
    if (hubs)
        for (int wi = gwarp; wi < a.n_heavy; wi += nwarps) {
            const int i = a.heavy[wi];
            double ctx[KT], lpv[KT];
            float ti[KT];
            ctx_labels_warp<KT>(K, i, rp, a.col, a.wgt, [&](int j) { return (unsigned)lab[j]; }, ctx,
                                lps.wsum_any_order != 0);
            const unsigned l = lab[i];
#pragma unroll
            for (int k = 0; k < KT; k++) ti[k] = (l == (unsigned)k) ? 1.f : 0.f;
            if (lane == 0) {
                load_lp<KT>(lps, K, (size_t)i, lpv);
                crit_site<KT>(K, lpv, ctx, ti, beta, c[0], c[1], c[2], c[3]);
            }
        }
    for (int base = a.row0 + gwarp * 32; base < a.row0 + n; base += nwarps * 32) {
        const int i = base + lane;
        const bool valid = i < a.row0 + n;
        int lo = 0, hi = 0;
        if (rp && valid) { lo = rp[i]; hi = rp[i + 1]; }
        const bool is_heavy = hubs && (hi - lo > HEAVY_DEG);
        double ctx[KT], lpv[KT];
        float ti[KT];
#pragma unroll
        for (int k = 0; k < KT; k++) ctx[k] = 0.0;
        if (valid) load_lp<KT>(lps, K, (size_t)i, lpv);
        if (rp) {
            const int seg_lo = __reduce_min_sync(FULL, valid ? lo : 0x7fffffff);
            const int seg_hi = __reduce_max_sync(FULL, valid ? hi : 0);
            if (is_heavy) lo = hi = 0;
            if (seg_lo < seg_hi)
                ctx_labels_coop<KT>(lo, hi, seg_lo, seg_hi, a.col, a.wgt, lab, s_w + wib * COOP_CHUNK,
                                    s_l + wib * COOP_CHUNK, ctx);
        }
        if (!valid || is_heavy) continue;
        const unsigned l = lab[i];
#pragma unroll
        for (int k = 0; k < KT; k++) ti[k] = (l == (unsigned)k) ? 1.f : 0.f;
        crit_site<KT>(K, lpv, ctx, ti, beta, c[0], c[1], c[2], c[3]);
    }
    pk_block_sum4<PK_THREADS>(c, sh);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < 4; q++) a.crit_partials[(size_t)blockIdx.x * 4 + q] = c[q];
    }
}
// this rank's sums D G L Z: its CTAs' partial rows added in CTA order (one CTA calls this)
static __device__ void pk_criteria_rank_sums(const nemk_persist_args &a, double *sh, double (&v)[4]) {
    for (int q = 0; q < 4; q++) v[q] = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += PK_THREADS)   // fixed assignment: deterministic
#pragma unroll
        for (int q = 0; q < 4; q++) v[q] += a.crit_partials[(size_t)b * 4 + q];
    pk_block_sum4<PK_THREADS>(v, sh);
}
static __device__ __forceinline__ void pk_criteria_combine(double beta, const double *v, double *crit6) {
    const double D = v[0], G = v[1], L = v[2], Z = v[3];
    crit6[0] = D + 0.5 * beta * G; crit6[1] = D; crit6[2] = L;
    crit6[3] = D + beta * G + Z; crit6[4] = Z; crit6[5] = G;
}

// ---- row shards, M-step: every rank's statistics (S, n, and the number of its families that
// moved) live in its exchange block; after the cross-rank barrier every rank reads all of them
// through peer memory and adds them in RANK ORDER (integers: exact, the same on every rank).  A
// rank changes its statistics again only after the next cross-rank barrier.
static __device__ int pk_stats_sum(const nemk_persist_args &a) {
    const int nstat = a.K * a.D + a.K;
    const int nthreads = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int q = gtid; q < nstat; q += nthreads) {
        int v = 0;
        for (int r = 0; r < a.world; r++) v += __ldcv(pk_peer<int32_t>(a, r, a.off_stat) + q);
        a.stat_glob[q] = v;
    }
    __shared__ int s_changed;
    if (threadIdx.x == 0) {
        int changed = 0;
        for (int r = 0; r < a.world; r++) changed += __ldcv(pk_peer<int32_t>(a, r, a.off_stat) + nstat);
        s_changed = changed;
    }
    __syncthreads();
    const int changed = s_changed;
    __syncthreads();
    return changed;
}

// =============================================================================================
template <int KT>
__global__ void __launch_bounds__(PK_THREADS, 2)
k_em_persist(const nemk_persist_args a) {
    __shared__ double sh[5 * 32];
    __shared__ float nkf[NEMB_MAX_K];
    __shared__ double nkd[NEMB_MAX_K];
    __shared__ float s_w[(PK_THREADS / 32) * COOP_CHUNK];
    __shared__ uint8_t s_l[(PK_THREADS / 32) * COOP_CHUNK];
    const int K = a.K, n = a.n, D = a.D;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const unsigned nb = gridDim.x;
    nemk_lpsrc lps;
    // (Hamming counts, margins and evaluation flags are stored for this rank's rows only; the host
    // passes margin / evflag already shifted by -row0, the Hamming counts are shifted here, so that
    // all of them are indexed by GLOBAL family id like the labels)
    lps.logpf = nullptr; lps.ham = a.ham - (size_t)a.row0 * a.K; lps.coef = a.coef; lps.wsum_any_order = a.wsum_any_order;
    const bool use_graph = a.use_graph != 0, seq = a.seq_sweep != 0;
    const bool margins = seq && a.use_margins;
    int state = a.entry, it = a.iter0, cur = a.cur, stale_par = a.stale_par;
    int stats_valid = a.stats_valid, last_changed = a.last_changed, margins_on = a.margins_on;
    int barriers = 0, sweeps = 0, x_passes = 0, recounts = 0, nfix = 0, cnt_par = a.cnt_par;
    int delta_mode = a.delta_mode;      // statistics vs lab[cur]: 0 stale (recount), 1 update from the scan's
                                        // list pending, 2 current (the sweep tracked its label moves)
    int flags_stale = a.flags_stale;    // n_allnul / n_ties do not describe the last sweep yet
    int mu_changed = a.mu_changed;      // the class masks moved in the last tables (cached Hamming counts void)
    long long kept = 0;
    const bool sharded = a.world > 1;
    unsigned xepoch = a.xepoch;          // cross-rank barriers passed so far (all ranks count alike)
    int xerr = 0, chg_local = a.chg_local, decide_pending = a.decide_pending;
    const uint32_t *xg = a.x - (size_t)a.row0 * a.wpr;      // X row of GLOBAL family i = xg + i * wpr
    int exit_code = NEMK_PK_EXIT_DONE, resume = NEMK_PK_ENTRY_MSTEP, converged = 0, empty = 0;
    int n_allnul = a.n_allnul, n_ties = a.n_ties;
    PkProf prof;
    for (int q = 0; q < 12; q++) prof.ns[q] = 0;
    prof.t_last = pk_now();
    prof.trace = nullptr;
    if (gtid == 0)
        for (int q = 0; q < 12 * 8; q++) (&a.out->trace[0][0])[q] = 0;
#define PK_SYNC() do { pk_grid_sync(a.bar, nb); barriers++; } while (0)

    // the entry codes plus the two initial sweeps as states of their own, so that the sweep has
    // ONE (inlined) call site
    enum { S_INIT = NEMK_PK_ENTRY_INIT, S_BLIND = NEMK_PK_ENTRY_INIT_SWEEPS, S_MSTEP = NEMK_PK_ENTRY_MSTEP,
           S_FINALIZE = NEMK_PK_ENTRY_FINALIZE, S_SWEEP = NEMK_PK_ENTRY_SWEEP, S_BETA0 = 5 };
    for (;;) {
        if (state == S_INIT) {
            // ---- prep: unlabelled state (the reference's calloc'd ClassifM), clean flags, presets
            // (row shards: every rank clears its own copy of the whole-pangenome arrays)
            const int n_all = sharded ? a.shard_len * a.world : n;
            for (int i = gtid; i < (n_all + 3) / 4; i += nthreads) {
                reinterpret_cast<uint32_t *>(a.lab[0])[i] = 0xffffffffu;
                reinterpret_cast<uint32_t *>(a.stale[0])[i] = 0u;
                reinterpret_cast<uint32_t *>(a.stale[1])[i] = 0u;
            }
            // (the evaluation flags exist for this rank's families only)
            for (int i = gtid; i < (n + 3) / 4; i += nthreads)
                reinterpret_cast<uint32_t *>(a.evflag + a.row0)[i] = 0u;
            mu_changed = 1;
            if (gtid == 0) {
                a.coef->uniform_ok = 1; a.coef->mu_changed = 1; a.coef->empty_class = 0; a.coef->halt = 0;
                // zero at rest from here on (a launch-per-stage fit on the same handle does not keep
                // the shared counters clean)
                for (int q = 0; q < 8; q++) a.wl_cnt[q] = 0;
                for (int q = 0; q < 16; q++) a.scratch[q] = 0;
                if (sharded)
                    for (int q = 0; q < 2 * NEMK_PK_MAX_WORLD; q++) pk_peer<int32_t>(a, a.rank, a.off_incnt)[q] = 0;
            }
            PK_SYNC();
            if (sharded) {
                // no rank stores a label into a peer before every rank has left its previous fit
                // and cleared its arrays
                pk_xbarrier(a, xepoch, xerr); barriers += 2;
                if (xerr) { exit_code = NEMK_PK_EXIT_PEER_TIMEOUT; break; }
            }
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
            PK_MARK(prof, 0);
            state = S_BLIND;
        } else if (state == S_MSTEP) {
            if (!decide_pending && it >= a.it_max) break;
            prof.trace = (gtid == 0 && it - a.iter0 < 12) ? a.out->trace[it - a.iter0] : nullptr;
            // statistics of the labels lab[cur]: already updated by the sweep (delta_mode 2), updated
            // from the scan's list (1), or recounted.  Row shards always update from the list once a
            // first recount exists (a uniform choice: every rank leaves the kernel, or none).
            const int moved = sharded ? chg_local : last_changed;      // rows of THIS rank that moved
            const bool incremental = !a.no_shortcuts && stats_valid && moved >= 0 && (sharded || moved <= n / 8);
            if (stats_valid && delta_mode == 2) {
                // the sweep tracked its label moves: nothing left to do
            } else if (incremental && delta_mode == 1) {
                mstep_delta_items<KT>(K, D, a.wpr, xg, a.lab[cur], a.lab[cur ^ 1], a.wlist[1], moved, a.stat,
                                      a.stat + (size_t)K * D);
                PK_SYNC();
                PK_MARK(prof, 2);
            } else if (a.x_in_kernel) {
                for (int q = gtid; q < K * D + K; q += nthreads) a.stat[q] = 0;
                PK_SYNC();
                pk_label_masks<KT>(a, a.lab[cur]);
                PK_SYNC();
                pk_recount<KT>(a);
                recounts++;
                PK_SYNC();
                PK_MARK(prof, 3);
            } else {
                exit_code = NEMK_PK_EXIT_NEED_RECOUNT; resume = NEMK_PK_ENTRY_FINALIZE;
                break;
            }
            state = S_FINALIZE;
        } else if (state == S_FINALIZE) {
            stats_valid = 1; delta_mode = 2;    // S and n describe lab[cur] (updated, recounted, or tracked)
            if (sharded) {
                // ---- the ranks' statistics meet: cross-rank barrier, then a rank-ordered sum of every
                // rank's block (this rank's a.stat IS its block's statistics area)
                if (gtid == 0) a.stat[K * D + K] = chg_local;
                pk_xbarrier(a, xepoch, xerr); barriers += 2;
                if (xerr) { exit_code = NEMK_PK_EXIT_PEER_TIMEOUT; break; }
                const int chg_glob = pk_stats_sum(a);
                // the request counters of the last sweep's meetings: every peer read them before it
                // came to the barrier above
                if (gtid == 0)
                    for (int q = 0; q < 2 * NEMK_PK_MAX_WORLD; q++) pk_peer<int32_t>(a, a.rank, a.off_incnt)[q] = 0;
                PK_SYNC();
                if (decide_pending) {        // the convergence test of the iteration whose sweep just ended
                    decide_pending = 0;
                    last_changed = chg_glob;
                    it++;
                    if (a.conv == 1) {
                        const float md = last_changed ? 1.0f : 0.0f;
                        if (md < a.conv_thr) { converged = 1; state = S_MSTEP; break; }
                    }
                    if (it >= a.it_max) { state = S_MSTEP; break; }
                }
            }
            for (int k = blockIdx.x; k < K; k += gridDim.x) pk_finalize_class<PK_THREADS>(k, a, sh, nkf, nkd);
            PK_SYNC();
            PK_MARK(prof, 4);
            const volatile nemk_coef *vc = a.coef;
            empty = vc->empty_class;
            mu_changed = 0;
            for (int k = 0; k < K; k++) mu_changed |= vc->mu_moved_k[k];
            mu_changed |= a.no_shortcuts;
            if (gtid == 0) a.coef->mu_changed = mu_changed;     // for the kernels that run outside
            if (empty) {   // nem_alg.c:1831-1838: the E-step is not run, the loop ends
                it++;
                break;
            }
            state = S_SWEEP;
            if (mu_changed) {
                if (a.x_in_kernel) {
                    pk_density_pass<KT>(a);
                    x_passes++;
                    PK_SYNC();
                    PK_MARK(prof, 5);
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
            // (PK_TRACK: statistics follow the label moves inside the sweep -- measured slower than
            // the scan + list update: moved rows cluster in a few warps and a row costs its set bits
            // in atomics; kept as a compile-time knob)
            const bool track_ok = PK_TRACK && state == S_SWEEP && stats_valid && delta_mode == 2 &&
                                  it + 1 < a.it_max && last_changed >= 0 && last_changed <= PK_TRACK_MAX;
            const int stale_next_buf = stale_par;       // after the flip above: index of mg.stale_next
            const bool tracked = pk_sweep<KT>(a, lps, blind ? 0.0 : a.beta, use_graph && !blind, seq && !blind,
                                              a.lab[cur], a.lab[cur ^ 1], cur ^ 1, mg, stale_next_buf, cnt, cnt_next,
                                              s_w, s_l, barriers, kept, nfix, prof, track_ok, mu_changed, xepoch, xerr);
            if (xerr) { exit_code = NEMK_PK_EXIT_PEER_TIMEOUT; break; }
            cur ^= 1;
            sweeps++;
            flags_stale = 1;
            if (blind) { state = S_BETA0; continue; }
            margins_on = 1;   // the next sweep uses the same beta
            if (state == S_BETA0) {   // NemAlgo starts with a full recount: nothing to scan
                if (sharded) {
                    // ... but the peers' copies of the other label buffer must follow (pk_scan fix_buf)
                    pk_scan<KT>(a, a.lab[cur], a.lab[cur ^ 1], 0, a.wlist[1], &a.wl_cnt[5], cur ^ 1);
                    PK_SYNC();
                }
                last_changed = -1; delta_mode = 0;
                state = S_MSTEP;
                continue;
            }
            if (tracked) {
                // the statistics followed the sweep: they describe lab[cur]; net number of moved labels
                last_changed = *(volatile int32_t *)&a.scratch[6];
                delta_mode = 2;
            } else {
                // ---- the rows this sweep moved: count (convergence test), flags of the last
                // evaluations, and the list the next M-step updates its statistics from
                const int want = (sharded ? !stats_valid : (it + 1 >= a.it_max || !stats_valid)) || a.no_shortcuts ? 0 : 1;
                pk_scan<KT>(a, a.lab[cur], a.lab[cur ^ 1], want, a.wlist[1], &a.wl_cnt[4], sharded ? (cur ^ 1) : -1);
                PK_SYNC();
                PK_MARK(prof, 1);
                last_changed = *(volatile int32_t *)&a.wl_cnt[4];
                n_allnul = *(volatile int32_t *)&a.scratch[3];
                n_ties = *(volatile int32_t *)&a.scratch[4];
                flags_stale = 0;
                delta_mode = want;
                if (want == 0) stats_valid = 0;   // neither current nor one list away from it
            }
            state = S_MSTEP;
            if (sharded) {
                // the number of moved labels is this rank's: the test waits for the ranks' sum, which
                // travels with the statistics (S_FINALIZE)
                chg_local = last_changed;
                decide_pending = 1;
                continue;
            }
            it++;
            if (a.conv == 1) {   // HasConverged `clas` under ncem: no label changed
                const float md = last_changed ? 1.0f : 0.0f;
                if (md < a.conv_thr) { converged = 1; break; }
            }
        }
    }
    if (exit_code == NEMK_PK_EXIT_DONE && flags_stale && sweeps > 0) {
        // all-null rows / exact ties of the last sweep: flags of every site's last evaluation
        // (scratch[3], [4] are zero: the sweep cleared them and no scan followed)
        pk_scan<KT>(a, a.lab[cur], a.lab[cur], 0, nullptr, &a.wl_cnt[5], -1);
        PK_SYNC();
        n_allnul = *(volatile int32_t *)&a.scratch[3];
        n_ties = *(volatile int32_t *)&a.scratch[4];
        flags_stale = 0;
    }
    // the fit is over: the final criteria, from the densities of the last tables and the labels after
    // the last sweep (nem_alg.c:1844-1853).  it == 0 (no EM iteration: the reference estimates
    // theta once more first) is left to the host.
    double crit6[6] = {0, 0, 0, 0, 0, 0};
    int have_crit = 0;
    if (exit_code == NEMK_PK_EXIT_DONE && a.want_crit && it > 0) {
        pk_criteria_partial<KT>(a, lps, a.lab[cur], s_w, s_l, sh);
        PK_SYNC();
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        if (blockIdx.x == 0) pk_criteria_rank_sums(a, sh, v);
        if (sharded) {
            // the ranks' sums (and their all-null / tie counts) meet like the statistics do: slot of
            // this rank in every rank's staging area, cross-rank barrier, rank-ordered sum
            if (gtid == 0) {
                for (int p = 0; p < a.world; p++) {
                    double *dst = pk_peer<double>(a, p, a.off_crit) + (size_t)a.rank * 8;
                    for (int q = 0; q < 4; q++) dst[q] = v[q];
                    dst[4] = (double)n_allnul; dst[5] = (double)n_ties;
                }
                __threadfence_system();
            }
            pk_xbarrier(a, xepoch, xerr); barriers += 2;
            if (xerr) exit_code = NEMK_PK_EXIT_PEER_TIMEOUT;
            const volatile double *st = pk_peer<double>(a, a.rank, a.off_crit);
            double w6[6] = {0, 0, 0, 0, 0, 0};
            for (int r = 0; r < a.world; r++)
                for (int q = 0; q < 6; q++) w6[q] += st[(size_t)r * 8 + q];
            for (int q = 0; q < 4; q++) v[q] = w6[q];
            n_allnul = (int)w6[4]; n_ties = (int)w6[5];
        }
        pk_criteria_combine(a.beta, v, crit6);
        have_crit = exit_code == NEMK_PK_EXIT_DONE;
        PK_MARK(prof, 11);
    }
    if (gtid == 0) {
        nemk_persist_out *o = a.out;
        for (int q = 0; q < 6; q++) o->crit[q] = crit6[q];
        o->have_crit = have_crit;
        o->exit_code = exit_code; o->resume_entry = resume;
        o->iters = it; o->converged = converged; o->empty_class = empty;
        o->cnt_par = cnt_par; o->delta_mode = delta_mode; o->flags_stale = flags_stale; o->mu_changed = mu_changed;
        o->cur = cur; o->stale_par = stale_par; o->last_changed = last_changed; o->stats_valid = stats_valid;
        o->n_allnul = n_allnul; o->n_ties = n_ties;
        o->xepoch = xepoch; o->xerror = xerr; o->decide_pending = decide_pending; o->chg_local = chg_local;
        o->sweeps = sweeps; o->x_passes = x_passes; o->recounts = recounts; o->barriers = barriers;
        o->kept = kept; o->fixup_rounds = nfix;
        for (int q = 0; q < 12; q++) o->phase_ns[q] = prof.ns[q];
        __threadfence_system();
        *(volatile unsigned long long *)&o->seq = a.seq;
    }
#undef PK_SYNC
}
