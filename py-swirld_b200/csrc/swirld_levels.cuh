// swirld_levels.cuh -- k_divide, level-scheduled (Node.divide_rounds, swirld.py:187-222).
//
// The only true dependency of an event is on its two parents, and `height`
// (swirld.py:114-120: 0 for roots, 1 + max over the parents) is exactly the length of
// the longest dependency chain below it: all events of one height are independent.
// So the chunk is first put in (height, arrival) order by a counting sort
// (k_lvl_hist / k_lvl_scan / k_lvl_scatter / k_lvl_desc, all parallel), and one
// persistent CTA then walks the levels:
//
//   * positions: every event gets a position g in processing order (gpos); the last
//     RING positions' results (row, T matrix, round) live in a shared-memory ring,
//     slot = g % RING, because parents are almost always a few levels back;
//   * warps 0-27 COMPUTE: a level has at most M <= 64 events, one warp per event, lane =
//     member column (NC columns per lane); a level usually fits one pass.  Levels are
//     separated by a named barrier over the compute warps.  Per event: read both parents'
//     slots, can_see row = max, T = OR (same-round parents only), stake-weighted
//     popcounts + ballots -> promotion, witness registration, own term, write the slot,
//     then stream the HBM copy (row, T, round, witness flag, SM) straight from registers.
//     Measured on B200: with so few events per level the walk is bound by the dependent
//     instruction latency of ONE event, so the widest split of an event (a full warp)
//     and the most warps in flight win over packing several events into a warp;
//   * warps 28-31 PREPARE the next batch of levels (<= 64 events) while the current one
//     computes: read the sorted descriptors, resolve each parent to a ring slot or copy
//     an older parent from L2/HBM into a staging slot with cp.async.
//   Both roles meet at one __syncthreads per batch.
#pragma once
#include "swirld_kernels.cuh"

#define LV_RING 192
#define LV_STAGE 32          // staging slots per buffer (2 buffers)
#define LV_BATCH 64          // events per batch (and the widest possible level)
#define LV_MAXLEV 32         // levels per batch
#define LV_THREADS 1024
#define LV_PREP_WARPS 4
#define LV_COMP_WARPS (LV_THREADS / 32 - LV_PREP_WARPS)
#define LV_PREP_THREADS (LV_PREP_WARPS * 32)
#define LV_COMP_THREADS (LV_COMP_WARPS * 32)

struct __align__(16) GDesc {  // one per position, written by k_lvl_desc
    int32_t h, cr, pa, pb, ga, gb, pad0, pad1;
};

struct LvlParams {
    int first, n;             // chunk [first, first+n) in arrival order == positions [first, first+n)
    int hmin, nbins;
    const int32_t *height, *p0, *p1, *creator;
    int32_t *hist;            // [nbins + 1]
    int32_t *cursor;          // [nbins]
    int32_t *order;           // [cap]  position -> event
    int32_t *gpos;            // [cap]  event -> position
    int32_t *lvl_start;       // [n + 1] positions where the chunk's levels start; [nl] = first + n
    int32_t *scal;            // SC_NLEV
    GDesc *gdesc;             // [cap]
};

__global__ void k_lvl_hist(LvlParams P) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < P.n; j += gridDim.x * blockDim.x)
        atomicAdd(&P.hist[P.height[P.first + j] - P.hmin], 1);
}

// exclusive scan of the histogram (one CTA) + compaction of the non-empty bins into lvl_start
__global__ void __launch_bounds__(1024, 1) k_lvl_scan(LvlParams P) {
    __shared__ int wsum_[32], wlev_[32];
    __shared__ int carry, carry_lev;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { carry = 0; carry_lev = 0; }
    __syncthreads();
    for (int base = 0; base < P.nbins; base += 1024) {
        const int i = base + tid;
        const int v = i < P.nbins ? P.hist[i] : 0;
        const int ne = v > 0 ? 1 : 0;
        int s = v, q = ne;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, s, o), b = __shfl_up_sync(0xffffffffu, q, o);
            if (lane >= o) { s += a; q += b; }
        }
        if (lane == 31) { wsum_[warp] = s; wlev_[warp] = q; }
        __syncthreads();
        if (warp == 0) {
            int a = wsum_[lane], b = wlev_[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int x = __shfl_up_sync(0xffffffffu, a, o), y = __shfl_up_sync(0xffffffffu, b, o);
                if (lane >= o) { a += x; b += y; }
            }
            wsum_[lane] = a; wlev_[lane] = b;
        }
        __syncthreads();
        const int excl = carry + (warp ? wsum_[warp - 1] : 0) + s - v;
        const int lev = carry_lev + (warp ? wlev_[warp - 1] : 0) + q - ne;
        if (i < P.nbins) {
            P.cursor[i] = P.first + excl;
            if (ne) P.lvl_start[lev] = P.first + excl;
        }
        __syncthreads();
        if (tid == 1023) { carry += wsum_[31]; carry_lev += wlev_[31]; }
        __syncthreads();
    }
    if (tid == 0) { P.lvl_start[carry_lev] = P.first + P.n; P.scal[SC_NSEG] = carry_lev; }
}

__global__ void k_lvl_scatter(LvlParams P) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < P.n; j += gridDim.x * blockDim.x) {
        const int h = P.first + j;
        const int pos = atomicAdd(&P.cursor[P.height[h] - P.hmin], 1);
        P.order[pos] = h;
        P.gpos[h] = pos;
    }
}

__global__ void k_lvl_desc(LvlParams P) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < P.n; j += gridDim.x * blockDim.x) {
        const int g = P.first + j;
        const int h = P.order[g];
        GDesc d;
        d.h = h; d.cr = P.creator[h]; d.pa = P.p0[h]; d.pb = P.p1[h];
        d.ga = d.pa >= 0 ? P.gpos[d.pa] : -1;
        d.gb = d.pb >= 0 ? P.gpos[d.pb] : -1;
        d.pad0 = d.pad1 = 0;
        P.gdesc[g] = d;
    }
}

// ---------------------------------------------------------------- the level walker
// ROWS = the walker also computes the can_see rows (fused); otherwise they were produced by
// the k_cs_* kernels (swirld_cansee.cuh) and the ring only carries T and the round.
template <int NC, bool ROWS>
struct __align__(16) LvSlot {
    int32_t row[NC * 32];
    u64 T[NC * 32];
    int32_t round;
    int32_t pad[3];
};
template <int NC>
struct __align__(16) LvSlot<NC, false> {
    u64 T[NC * 32];
    int32_t round;
    int32_t pad[3];
};
template <int NC, bool ROWS> struct LvOwnRows { int32_t r[3][LV_BATCH][NC * 32]; };   // row(h) of the batch's events
template <int NC> struct LvOwnRows<NC, true> { int32_t r[1][1][4]; };

struct __align__(16) LvDesc {  // per event of a batch, in shared memory
    int32_t h, cr, pa, pb;
    int32_t la, lb;            // parent location: >= 0 slot index, -1 none (root), -2 read from HBM
    int32_t pad0, pad1;
};

template <int NC, bool ROWS>
struct LvSmem {
    LvSlot<NC, ROWS> slot[LV_RING + 2 * LV_STAGE];
    __align__(16) LvOwnRows<NC, ROWS> own;
    LvDesc desc[3][LV_BATCH];
    int32_t lv_off[3][LV_MAXLEV + 1];
    int32_t nlev[3], bsize[3], bstart[3], nact[3];
    int32_t stage_cnt[2];
    int32_t stage_p[2][LV_STAGE];
    __align__(16) int32_t Wc[SW_WC][NC * 32];
    __align__(16) i64 stake[NC * 32];
    int rmaxp[2];
};

struct Div4Params {
    DivParams d;              // tables, stake, thresholds, scal
    const GDesc *gdesc;
    const int32_t *lvl_start;  // levels of this chunk; their count is scal[SC_NSEG]
};

__device__ __forceinline__ unsigned lv_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lv_cp_async4(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(lv_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void lv_cp_async8(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(lv_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void lv_cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void bar_named(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}

// ---- rare paths
template <int NC, bool ROWS>
__device__ __forceinline__ void lv_parent_from_hbm(const DivParams &P, int p, int lane, int &r, int (&v)[NC], u64 (&t)[NC]) {
    r = __ldcg(P.round + p);
#pragma unroll
    for (int j = 0; j < NC; j++) {
        const int c = lane + 32 * j;
        if (ROWS) v[j] = c < P.M ? __ldcg(P.row + (size_t)p * P.M + c) : -1;
        t[j] = c < P.M ? __ldcg(P.T + (size_t)p * P.M + c) : 0ull;
    }
}
template <int NC>
__device__ __forceinline__ void lv_witnesses_from_hbm(const DivParams &P, int rh, int lane, int (&w)[NC]) {
#pragma unroll
    for (int j = 0; j < NC; j++) {
        const int c = lane + 32 * j;
        w[j] = c < P.M ? __ldcg(P.W + (size_t)rh * P.M + c) : -1;
    }
}

// FULL: M == 32*NC exactly (no padded member columns) -- drops the per-column predicates.
template <int NC, bool UNIT, bool ROWS, bool FULL>
__global__ void __launch_bounds__(LV_THREADS, 1) k_divide_levels(Div4Params Q) {
    extern __shared__ __align__(16) unsigned char smraw[];
    LvSmem<NC, ROWS> &S = *reinterpret_cast<LvSmem<NC, ROWS> *>(smraw);
    const DivParams &P = Q.d;
    constexpr int MS = NC * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = P.M;
    const i64 thr = P.tot2 / 3;             // 3*x > 2*tot  <=>  x > floor(2*tot/3) for integers
    const bool is_compute = warp < LV_COMP_WARPS;
    const int first = P.first, last = P.first + P.n;
    const int nl = P.scal[SC_NSEG];         // written by k_lvl_scan
    bool act[NC];
#pragma unroll
    for (int j = 0; j < NC; j++) act[j] = FULL || lane + 32 * j < M;

    // ---- one-time init
    if (tid < MS) S.stake[tid] = tid < M ? P.stake[tid] : 0;
    for (int i = tid; i < (LV_RING + 2 * LV_STAGE) * MS; i += LV_THREADS) {   // padded columns stay (-1, 0)
        if constexpr (ROWS) S.slot[i / MS].row[i % MS] = -1;
        S.slot[i / MS].T[i % MS] = 0;
    }
    if constexpr (!ROWS)
        for (int i = tid; i < 3 * LV_BATCH * MS; i += LV_THREADS) S.own.r[i / (LV_BATCH * MS)][(i / MS) % LV_BATCH][i % MS] = -1;
    int rmax = P.scal[SC_MAX_ROUND];
    int wbase = max(0, rmax - (SW_WC / 2 - 1));
    if (tid < 2) S.rmaxp[tid] = rmax;
    if (tid < 3) { S.nlev[tid] = 0; S.bsize[tid] = 0; S.bstart[tid] = first; }
    for (int i = tid; i < SW_WC * MS; i += LV_THREADS) {
        const int slot = i / MS, c = i % MS;
        const int r = wbase + ((slot - (wbase % SW_WC) + SW_WC) % SW_WC);
        S.Wc[slot][c] = (c < M && r < P.Rcap) ? P.W[(size_t)r * M + c] : -1;
    }
    __syncthreads();

    // ---- prepare batch `b` into descriptor buffer b % 3 and staging buffer b & 1 (prep warps)
    int lv_cursor = 0;                      // next level of the chunk (uniform over the prep threads)
    auto prep = [&](int b) {
        const int db = b % 3, sb = b & 1;
        const int pt = tid - LV_COMP_THREADS;       // 0..LV_PREP_THREADS-1
        if (pt < 32) {                      // batch extent: as many whole levels as fit in LV_BATCH events
            const int gb0 = lv_cursor < nl ? __ldcg(Q.lvl_start + lv_cursor) : last;
            const int li = lv_cursor + 1 + lane;
            const int v = li <= nl ? __ldcg(Q.lvl_start + li) : 0x7fffffff;
            const unsigned fits = __ballot_sync(0xffffffffu, v - gb0 <= LV_BATCH && li <= nl);
            // levels are contiguous, so `fits` is a prefix mask
            const int cnt = lv_cursor < nl ? __popc(fits) : 0;
            if (lane == 0) {
                S.nlev[db] = cnt; S.bstart[db] = gb0; S.stage_cnt[sb] = 0;
                S.lv_off[db][0] = 0;
                if (cnt == 0) S.bsize[db] = 0;
                if (lv_cursor < nl && cnt == 0) atomicMin(&P.scal[SC_ERR], -8);   // a level wider than LV_BATCH
            }
            if (lane < cnt) {
                S.lv_off[db][lane + 1] = v - gb0;
                if (lane == cnt - 1) S.bsize[db] = v - gb0;
            }
            // widest level of the batch -> how many compute warps take part in its level barriers
            const int prevv = __shfl_up_sync(0xffffffffu, v, 1);
            int wdt = lane < cnt ? v - (lane == 0 ? gb0 : prevv) : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) wdt = max(wdt, __shfl_xor_sync(0xffffffffu, wdt, o));
            if (lane == 0) S.nact[db] = min(LV_COMP_WARPS, max(wdt, 1));
        }
        bar_named(2, LV_PREP_THREADS);
        const int cnt = S.nlev[db], gb0 = S.bstart[db], bs = S.bsize[db], ge = gb0 + bs;
        lv_cursor += cnt;
        if (pt < bs) {
            const GDesc g = Q.gdesc[gb0 + pt];
            LvDesc d;
            d.h = g.h; d.cr = g.cr; d.pa = g.pa; d.pb = g.pb; d.pad0 = d.pad1 = 0;
            d.la = d.lb = -1;
            if (g.pa >= 0) {
                // in the ring iff computed in this launch and not overwritten by this batch
                if (g.ga >= first && g.ga >= ge - LV_RING) d.la = g.ga % LV_RING;
                else {
                    const int s = atomicAdd(&S.stage_cnt[sb], 1);
                    if (s < LV_STAGE) { S.stage_p[sb][s] = g.pa; d.la = LV_RING + sb * LV_STAGE + s; }
                    else d.la = -2;
                }
                if (g.gb >= first && g.gb >= ge - LV_RING) d.lb = g.gb % LV_RING;
                else {
                    const int s = atomicAdd(&S.stage_cnt[sb], 1);
                    if (s < LV_STAGE) { S.stage_p[sb][s] = g.pb; d.lb = LV_RING + sb * LV_STAGE + s; }
                    else d.lb = -2;
                }
            }
            S.desc[db][pt] = d;
        }
        bar_named(2, LV_PREP_THREADS);
        // copy the staged (old) parents from the HBM tables, one prep warp per parent
        const int ns = min(S.stage_cnt[sb], LV_STAGE);
        for (int s = warp - LV_COMP_WARPS; s < ns; s += LV_PREP_WARPS) {
            const int p = S.stage_p[sb][s];
            LvSlot<NC, ROWS> *dst = &S.slot[LV_RING + sb * LV_STAGE + s];
            for (int c = lane; c < M; c += 32) {
                if constexpr (ROWS) lv_cp_async4(&dst->row[c], P.row + (size_t)p * M + c);
                lv_cp_async8(&dst->T[c], P.T + (size_t)p * M + c);
            }
            if (lane == 0) lv_cp_async4(&dst->round, P.round + p);
        }
        if constexpr (!ROWS) {              // can_see rows (made by k_cs_*) of the batch's FIRST level;
            const int n0 = cnt > 0 ? S.lv_off[db][1] : 0;   // later levels prefetch theirs from the compute warps
            for (int k = warp - LV_COMP_WARPS; k < n0; k += LV_PREP_WARPS) {
                const int h = S.desc[db][k].h;
                for (int c = lane; c < M; c += 32) lv_cp_async4(&S.own.r[db][k][c], P.row + (size_t)h * M + c);
            }
        }
    };

    // ---- one event, one warp (lane = member column, NC columns per lane)
    auto process = [&](const LvDesc *dp, int slot_h, unsigned parity, const int (&own_row)[NC]) {
        const int4 d0 = reinterpret_cast<const int4 *>(dp)[0];        // h, cr, pa, pb
        const int2 d1 = reinterpret_cast<const int2 *>(dp)[2];        // la, lb
        const int eh = d0.x, cr = d0.y, pa = d0.z, pb = d0.w, la = d1.x, lb = d1.y;
        int rowh[NC];
        u64 t[NC];
        int r = -1, ra = -1;
        bool promoted = true;                               // a root: round 0, own term only
#pragma unroll
        for (int j = 0; j < NC; j++) { rowh[j] = -1; t[j] = 0; }
        if (pa >= 0) {
            int va[NC], vb[NC];
            u64 ta[NC], tb[NC];
            int rb;
#pragma unroll
            for (int j = 0; j < NC; j++) { va[j] = -1; vb[j] = -1; }
            if (la >= 0) {
                const LvSlot<NC, ROWS> &A = S.slot[la];
                ra = A.round;
#pragma unroll
                for (int j = 0; j < NC; j++) {
                    if constexpr (ROWS) va[j] = A.row[lane + 32 * j];
                    ta[j] = A.T[lane + 32 * j];
                }
            } else lv_parent_from_hbm<NC, ROWS>(P, pa, lane, ra, va, ta);   // staging overflow
            if (lb >= 0) {
                const LvSlot<NC, ROWS> &B = S.slot[lb];
                rb = B.round;
#pragma unroll
                for (int j = 0; j < NC; j++) {
                    if constexpr (ROWS) vb[j] = B.row[lane + 32 * j];
                    tb[j] = B.T[lane + 32 * j];
                }
            } else lv_parent_from_hbm<NC, ROWS>(P, pb, lane, rb, vb, tb);
            r = max(ra, rb);                                                // swirld.py:200
            const u64 ka = ra == r ? ~0ull : 0ull, kb = rb == r ? ~0ull : 0ull;
            int cnt = 0;
#pragma unroll
            for (int j = 0; j < NC; j++) {
                if constexpr (ROWS) rowh[j] = max(va[j], vb[j]);            // swirld.py:203-205
                t[j] = (ta[j] & ka) | (tb[j] & kb);
                bool ss;                                                    // swirld.py:209-214
                if (UNIT) ss = __popcll(t[j]) > (int)thr;
                else ss = wsum(t[j], 0, S.stake) > thr;
                cnt += __popc(__ballot_sync(0xffffffffu, ss));
            }
            promoted = (i64)cnt > thr;                                      // swirld.py:216 (quirk Q3)
        }
        int rh = r + (promoted ? 1 : 0);                                    // swirld.py:217-219
        const bool wit = rh > ra;                                           // swirld.py:221 / 196-197
        if (rh >= P.Rcap) { if (lane == 0) atomicMin(&P.scal[SC_ERR], -5); rh = P.Rcap - 1; }
        const bool w_sm = rh >= wbase && rh < wbase + SW_WC;
        const int wslot = rh % SW_WC;
        int w[NC];
        if (w_sm) {
#pragma unroll
            for (int j = 0; j < NC; j++) w[j] = S.Wc[wslot][lane + 32 * j];
        } else lv_witnesses_from_hbm<NC>(P, rh, lane, w);            // a round outside the cached window
        if (wit && lane == 0) {                                              // swirld.py:222
            if (w_sm) S.Wc[wslot][cr] = eh;
            P.W[(size_t)rh * M + cr] = eh;
            atomicMax(&S.rmaxp[parity], rh);
        }
        if constexpr (!ROWS) {                              // the finished row, from k_cs_*
#pragma unroll
            for (int j = 0; j < NC; j++) rowh[j] = own_row[j];
        }
        const u64 keep = promoted ? 0ull : ~0ull, own = 1ull << cr;
        u64 smask = 0;
        LvSlot<NC, ROWS> &H = S.slot[slot_h];
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const int c = lane + 32 * j;
            const bool mine = c == cr;
            if (mine) rowh[j] = eh;                                          // swirld.py:220
            const int wk = (mine && wit) ? eh : w[j];
            const bool sm = wk >= 0 && rowh[j] >= wk;
            smask |= (u64)__ballot_sync(0xffffffffu, sm) << (32 * j);
            t[j] = (t[j] & keep) | (sm ? own : 0ull);
            if constexpr (ROWS) H.row[c] = rowh[j];         // publish: what the next level reads
            H.T[c] = t[j];
        }
        if (lane == 0) H.round = rh;
        // ---- stream the HBM copy straight from registers (fire and forget)
        {
            const size_t o = (size_t)eh * M + lane;
            int32_t *grow = P.row + o;
            u64 *gT = P.T + o;
#pragma unroll
            for (int j = 0; j < NC; j++)
                if (act[j]) {
                    if constexpr (ROWS) grow[32 * j] = rowh[j];
                    gT[32 * j] = t[j];
                }
        }
        if (lane == 0) {
            P.round[eh] = rh;
            P.wit[eh] = wit ? 1 : 0;
            P.SM[eh] = smask;
        }
    };

    // ---- pipeline: prep(0) | compute(b) || prep(b+1) | ...
    if (!is_compute) { prep(0); lv_cp_async_wait_all(); }
    __syncthreads();
    long long c_proc = 0, c_lbar = 0, c_bbar = 0, c_prep = 0, c_wait = 0, c_nproc = 0, c_nlev = 0, c_nbatch = 0;
    for (int b = 0;; ++b) {
        const int db = b % 3;
        const int nlev = S.nlev[db];
        if (nlev == 0) break;                               // uniform: written before the barrier
        const unsigned parity = (unsigned)b & 1u;
        long long t0 = clock64();
        const int nact = S.nact[db];
        if (is_compute && warp < nact) {            // the other compute warps have nothing in this batch
            const int gb0 = S.bstart[db];
            int nrow[NC];                                   // own row prefetched for the next level
            bool have_n = false;
#pragma unroll
            for (int j = 0; j < NC; j++) nrow[j] = -1;
            for (int l = 0; l < nlev; ++l) {
                const int lo = S.lv_off[db][l], hi = S.lv_off[db][l + 1];
                int prow[NC];
                bool have_p = false;
#pragma unroll
                for (int j = 0; j < NC; j++) prow[j] = -1;
                if constexpr (!ROWS) {
                    if (l + 1 < nlev && hi + warp < S.lv_off[db][l + 2]) {   // my event of the next level
                        const int eh2 = S.desc[db][hi + warp].h;
#pragma unroll
                        for (int j = 0; j < NC; j++)
                            if (act[j]) prow[j] = __ldcg(P.row + (size_t)eh2 * M + lane + 32 * j);
                        have_p = true;
                    }
                }
                for (int k = lo + warp; k < hi; k += nact) {
                    int orow[NC];
#pragma unroll
                    for (int j = 0; j < NC; j++) orow[j] = -1;
                    if constexpr (!ROWS) {
                        if (l == 0) {
#pragma unroll
                            for (int j = 0; j < NC; j++) orow[j] = S.own.r[db][k][lane + 32 * j];
                        } else if (k == lo + warp && have_n) {
#pragma unroll
                            for (int j = 0; j < NC; j++) orow[j] = nrow[j];
                        } else {
                            const int eh1 = S.desc[db][k].h;
#pragma unroll
                            for (int j = 0; j < NC; j++)
                                if (act[j]) orow[j] = __ldcg(P.row + (size_t)eh1 * M + lane + 32 * j);
                        }
                    }
                    process(&S.desc[db][k], (gb0 + k) % LV_RING, parity, orow);
                    c_nproc++;
                }
#pragma unroll
                for (int j = 0; j < NC; j++) nrow[j] = prow[j];
                have_n = have_p;
                __syncwarp();
                long long t1 = clock64(); c_proc += t1 - t0;
                bar_named(1, nact * 32);
                t0 = clock64(); c_lbar += t0 - t1; c_nlev++;
            }
        } else if (!is_compute) {
            prep(b + 1);
            long long t2 = clock64(); c_prep += t2 - t0;
            lv_cp_async_wait_all();
            t0 = clock64(); c_wait += t0 - t2;
        }
        __syncthreads();
        c_bbar += clock64() - t0; c_nbatch++;
        rmax = max(rmax, S.rmaxp[parity]);
        const int nbase = max(wbase, rmax - (SW_WC / 2 - 1));
        if (nbase != wbase) {                               // uniform: every thread sees the same rmax
            for (int i = tid; i < SW_WC * MS; i += LV_THREADS) {
                const int slot = i / MS, c = i % MS;
                const int ro = wbase + ((slot - (wbase % SW_WC) + SW_WC) % SW_WC);
                const int rn = nbase + ((slot - (nbase % SW_WC) + SW_WC) % SW_WC);
                if (rn != ro)
                    S.Wc[slot][c] = (c < M && rn < P.Rcap) ? __ldcg(P.W + (size_t)rn * M + c) : -1;
            }
            wbase = nbase;
            __syncthreads();
        }
        if (S.nlev[(b + 1) % 3] == 0) break;                // that was the last batch
    }
    if (tid == 0) P.scal[SC_MAX_ROUND] = rmax;
    if (P.dbg && lane == 0 && (warp == 0 || warp == LV_COMP_WARPS)) {   // cycle accounting, one warp per role
        unsigned long long *o = (unsigned long long *)P.dbg + (warp == 0 ? 0 : 8);
        if (warp == 0) { atomicAdd(&o[0], (unsigned long long)c_proc); atomicAdd(&o[1], (unsigned long long)c_lbar);
                         atomicAdd(&o[2], (unsigned long long)c_bbar); atomicAdd(&o[3], (unsigned long long)c_nproc);
                         atomicAdd(&o[4], (unsigned long long)c_nlev); atomicAdd(&o[5], (unsigned long long)c_nbatch); }
        else { atomicAdd(&o[1], (unsigned long long)c_prep); atomicAdd(&o[2], (unsigned long long)c_wait);
               atomicAdd(&o[3], (unsigned long long)c_bbar); }
    }
}
