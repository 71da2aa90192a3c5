// swirld_rounds.cuh -- round numbers for a chunk of events on the WHOLE GPU
// (Node.divide_rounds, swirld.py:187-222; the can_see rows come from swirld_cansee.cuh).
//
// The reference walks the events one by one because `round[h]` reads `round[parent]`.
// On a fork-free graph the same numbers follow from a monotone predicate instead:
//
//   Wf_r[c]   := member c's FIRST event with round >= r                  (-1: none yet)
//   P_r(h)    := #{c_ : hits_r(h)[c_] > 2T/3} > 2T/3, with
//   hits_r(h)[c_] = sum over members c of stake[c] * [pre(h)[c] >= Wf_r[c] >= 0]
//                                                 * [row(pre(h)[c])[c_] >= Wf_r[c_] >= 0]
//   (pre(h) = can_see row of h with the own column set back to the self-parent, swirld.py:203-205, 220)
//
//   round[h] >= r+1  <=>  P_r(h)          and P_r is monotone along every member's chain.
//
// (=>: an event is promoted from r exactly when the reference's strongly-sees count
// passes, swirld.py:209-219, and every descendant keeps seeing at least as much; <=: P_r(h)
// needs a parent of round >= r.  Using the FIRST event of round >= r instead of the round-r
// witness changes nothing, because seeing a skipping member's event already implies a
// higher round.  tests/test_rounds_model.py keeps the executable proof against the oracle.)
//
// So one cooperative kernel advances all chains round by round: in every step each member
// chain that is waiting at the lowest open round r tests P_r on its next L pending events
// (one warp per event, all SMs), a grid barrier follows, and every CTA applies the same
// bookkeeping: events before the chain's first hit are final at round r, the first hit opens
// round r+1 for that chain (Wf_{r+1}[c]).  The depth of the computation is the number of
// ROUNDS in the chunk (~1 per 690 events at 64 members), not the number of DAG levels.
//
// S_r(k) = {c_ : row(k)[c_] >= Wf_r[c_] >= 0} of an event k is cached per event with the
// round it was computed for (a 16-byte {mask, mask ^ key(r)} entry, so a torn or stale read
// is detected and simply recomputed).
#pragma once
#include <cooperative_groups.h>
#include "swirld_kernels.cuh"

namespace cg = cooperative_groups;

#define RB_WR 32            // rounds of Wf mirrored in shared memory
#define RB_LMAX 64          // pending events tested per chain and step (at most)
#define RB_THREADS 512       // 16 warps per CTA, one CTA per SM
#define RB_RING 256          // per-member ring of recent events (what precedes the chunk)
#define RB_MAXMISS 8         // an event with more unprepared S_r masks than this waits for a later step

struct RbParams {
    int M, first, n, Rcap, L;
    const int32_t *row, *p0, *creator, *seq;
    int32_t *round;             // [cap] out
    int32_t *Wf;                // [Rcap][M] first event of round >= r per member
    ulonglong2 *sc;             // [cap] {S_r(k), S_r(k) ^ key(r)}
    int32_t *cev;               // [cap] chunk events grouped by creator, region [first, first+n)
    int32_t *ccnt, *cmin;       // [64] per-member count / smallest seq inside the chunk
    int32_t *ctot;              // [64] events of the member so far (1 + its largest seq)
    int32_t *gchain;            // [64][RB_RING] the member's most recent events by seq % RB_RING
    int32_t *coff;              // [65]
    uint8_t *res;               // [2][64 * RB_LMAX]
    const i64 *stake;
    i64 tot2;
    int32_t *scal;
    long long *dbg;             // cycle counters (profiling), may be NULL
    uint8_t *wit;               // [cap]   (finish kernels)
    int32_t *W;                 // [Rcap][M] the reference's witnesses table
    u64 *SM;                    // [cap]
};

// ---- per-member event lists of the chunk
__global__ void k_rb_count(RbParams P) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < P.n; j += gridDim.x * blockDim.x) {
        const int h = P.first + j, c = P.creator[h];
        const int sq = P.seq[h];
        atomicAdd(&P.ccnt[c], 1);
        atomicMin(&P.cmin[c], sq);
        atomicMax(&P.ctot[c], sq + 1);
    }
}
__global__ void k_rb_offsets(RbParams P) {       // one warp
    const int lane = threadIdx.x;
    int a = lane < P.M ? P.ccnt[lane] : 0, b = lane + 32 < P.M ? P.ccnt[lane + 32] : 0;
    int sa = a, sb = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int x = __shfl_up_sync(0xffffffffu, sa, o), y = __shfl_up_sync(0xffffffffu, sb, o);
        if (lane >= o) { sa += x; sb += y; }
    }
    const int tot_a = __shfl_sync(0xffffffffu, sa, 31);
    P.coff[lane] = sa - a;
    P.coff[lane + 32] = tot_a + sb - b;
    if (lane == 31) P.coff[64] = tot_a + sb;
}
__global__ void k_rb_scatter(RbParams P) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < P.n; j += gridDim.x * blockDim.x) {
        const int h = P.first + j, c = P.creator[h];
        P.cev[P.first + P.coff[c] + P.seq[h] - P.cmin[c]] = h;
    }
}

// after the chunk: remember each member's most recent RB_RING events for the next chunk
__global__ void k_rb_tail(RbParams P) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < P.n; j += gridDim.x * blockDim.x) {
        const int h = P.first + j, c = P.creator[h], sq = P.seq[h];
        if (P.ctot[c] - sq <= RB_RING) P.gchain[c * RB_RING + (sq & (RB_RING - 1))] = h;
    }
}

__device__ __forceinline__ u64 rb_key(int r) { return (u64)(r + 1) * 0x9E3779B97F4A7C15ull; }

template <int NC, bool UNIT>
__global__ void __launch_bounds__(RB_THREADS, 1) k_rounds_batch(RbParams P) {
    cg::grid_group grid = cg::this_grid();
    __shared__ int cur[64], pos[64], len[64], off[64];
    __shared__ int Wl[RB_WR][64];
    __shared__ i64 stake_s[64];
    __shared__ int s_rmin, s_newtop, s_nfin[64], s_base[64];
    __shared__ int cmin_s[64], ctot_s[64], ra_lo[64], ra_pre[65];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = P.M, L = P.L;
    const int gw = blockIdx.x * (blockDim.x >> 5) + warp, nw = gridDim.x * (blockDim.x >> 5);
    const i64 thr = P.tot2 / 3;
    const bool lead = blockIdx.x == 0;

    int rtop = max(P.scal[SC_MAX_ROUND], 0);
    for (int i = tid; i < RB_WR * 64; i += blockDim.x) {
        const int slot = i >> 6, c = i & 63;
        // the round of (rtop-RB_WR, rtop] that maps to this slot
        const int r = rtop - ((rtop - slot) & (RB_WR - 1));
        Wl[slot][c] = (c < M && r >= 0 && r < P.Rcap) ? __ldcg(P.Wf + (size_t)r * M + c) : -1;
    }
    if (tid < 64) {
        const int c = tid;
        stake_s[c] = c < M ? P.stake[c] : 0;
        int o = 0, l = 0, cu = 0x7fffffff;
        if (c < M) {
            o = P.first + P.coff[c]; l = P.coff[c + 1] - P.coff[c];
            if (l > 0) {
                const int h0 = P.cev[o], pa = P.p0[h0];
                cu = pa < 0 ? 0 : P.round[pa];
            }
        }
        off[c] = o; len[c] = l; pos[c] = 0; cur[c] = cu;
        cmin_s[c] = c < M ? P.cmin[c] : 0; ctot_s[c] = c < M ? P.ctot[c] : 0;
    }
    __syncthreads();
    if (tid < M && len[tid] > 0 && cur[tid] == 0) {          // a member's root opens round 0 for it
        const int h0 = P.cev[off[tid]];
        if (P.p0[h0] < 0) {
            if (rtop < RB_WR) Wl[0][tid] = h0;
            if (lead) P.Wf[tid] = h0;
        }
    }
    grid.sync();                                              // (late roots write the global table only)
    if (rtop >= RB_WR && tid < M) { /* round 0 is outside the mirror: nothing to refresh */ }

    auto wrow = [&](int r, int c) -> int {                    // Wf_r[c]
        if (r > rtop - RB_WR && r <= rtop) return Wl[r & (RB_WR - 1)][c];
        return __ldcg(P.Wf + (size_t)r * M + c);
    };

    long long c_miss = 0, c_lastmiss = 0;
    // ---- P_r(h) by one warp
    auto eval = [&](int h, int r, bool may_defer) -> int {
        const int cr = P.creator[h], pa = P.p0[h];
        int W[NC], pre[NC];
        bool live[NC];
        u64 m[NC];
        i64 lv = 0;
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const int c = lane + 32 * j;
            W[j] = c < M ? wrow(r, c) : -1;
            pre[j] = c < M ? __ldcg(P.row + (size_t)h * M + c) : -1;
            if (c == cr) pre[j] = pa;
            live[j] = W[j] >= 0 && pre[j] >= W[j];
            m[j] = 0;
            if (UNIT) lv += __popc(__ballot_sync(0xffffffffu, live[j]));
            else {
                i64 s = live[j] ? stake_s[c] : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                lv += s;
            }
        }
        if (lv <= thr) return 0;                              // hits[c_] <= stake of the live members
        const u64 key = rb_key(r);
        bool valid[NC];
#pragma unroll
        for (int j = 0; j < NC; j++) {
            valid[j] = true;
            if (live[j]) {
                const ulonglong2 e = __ldcg(P.sc + pre[j]);
                valid[j] = (e.x ^ e.y) == key;
                m[j] = e.x;
            }
        }
        {   // an event far ahead of the tested windows sees events nobody prepared: leave it for a later step
            int nm = 0;
#pragma unroll
            for (int j = 0; j < NC; j++) nm += __popc(__ballot_sync(0xffffffffu, live[j] && !valid[j]));
            c_lastmiss = nm;
            if (may_defer && nm > RB_MAXMISS) return 2;       // (never the chain's first pending event: progress)
        }
#pragma unroll
        for (int jj = 0; jj < NC; jj++) {                     // cache misses: compute S_r(k) together,
            unsigned miss = __ballot_sync(0xffffffffu, live[jj] && !valid[jj]);
            c_miss += __popc(miss);
            while (miss) {                                    // four rows in flight per trip
                int kk[4], ll[4], v[4][NC];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    ll[u] = miss ? __ffs(miss) - 1 : -1;
                    if (miss) miss &= miss - 1;
                    kk[u] = __shfl_sync(0xffffffffu, pre[jj], ll[u] < 0 ? 0 : ll[u]);
#pragma unroll
                    for (int j = 0; j < NC; j++) {
                        const int c = lane + 32 * j;
                        v[u][j] = (ll[u] >= 0 && c < M) ? __ldcg(P.row + (size_t)kk[u] * M + c) : -1;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (ll[u] < 0) continue;                  // uniform
                    u64 mask = 0;
#pragma unroll
                    for (int j = 0; j < NC; j++)
                        mask |= (u64)__ballot_sync(0xffffffffu, W[j] >= 0 && v[u][j] >= W[j]) << (32 * j);
                    if (lane == 0) P.sc[kk[u]] = make_ulonglong2(mask, mask ^ key);
                    if (lane == ll[u]) m[jj] = mask;
                }
            }
        }
        i64 hits[NC];
#pragma unroll
        for (int j = 0; j < NC; j++) hits[j] = 0;
#pragma unroll
        for (int jj = 0; jj < NC; jj++) {                     // column sums over the live members' masks
            const u64 mine = live[jj] ? m[jj] : 0ull;         // a member that is not live contributes nothing
#pragma unroll 8
            for (int l = 0; l < 32; l++) {
                const u64 mm = __shfl_sync(0xffffffffu, mine, l);
                const i64 st = UNIT ? 1 : stake_s[jj * 32 + l];
#pragma unroll
                for (int j = 0; j < NC; j++) hits[j] += ((mm >> (lane + 32 * j)) & 1) ? st : 0;
            }
        }
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < NC; j++) cnt += __popc(__ballot_sync(0xffffffffu, hits[j] > thr));
        return (i64)cnt > thr ? 1 : 0;
    };

    long long c_eval = 0, c_sync = 0, c_upd = 0, c_nev = 0, c_steps = 0, c_maxev = 0, c_sumev = 0;
    for (int step = 0;; ++step) {
        const long long t0 = clock64();
        // ---- lowest open round
        if (warp == 0) {
            int a = 0x7fffffff;
            for (int c = lane; c < M; c += 32) if (pos[c] < len[c]) a = min(a, cur[c]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a = min(a, __shfl_xor_sync(0xffffffffu, a, o));
            if (lane == 0) s_rmin = a;
        }
        __syncthreads();
        const int rmin = s_rmin;
        if (rmin == 0x7fffffff) break;                        // every chain is done
        uint8_t *res = P.res + (size_t)(step & 1) * 64 * RB_LMAX;
        // ---- S_rmin(k) of every event the tests below can meet: per member, its events from
        //      Wf_rmin[c] up to the end of its pending window (one warp per event, all SMs)
        if (tid < 64) {
            int lo = 0, cnt = 0;
            if (tid < M) {
                const int w = wrow(rmin, tid);
                if (w >= 0) {
                    lo = __ldcg(P.seq + w);
                    const int hi = len[tid] > 0 ? cmin_s[tid] + min(len[tid], pos[tid] + L) : ctot_s[tid];
                    cnt = max(0, min(hi - lo, RB_RING));
                }
            }
            ra_lo[tid] = lo; ra_pre[tid + 1] = cnt;
        }
        __syncthreads();
        if (tid == 0) { ra_pre[0] = 0; for (int c = 0; c < 64; c++) ra_pre[c + 1] += ra_pre[c]; }
        __syncthreads();
        {
            const int total = ra_pre[64];
            const u64 key = rb_key(rmin);
            for (int i = gw; i < total; i += nw) {
                int c = 0;                                     // member whose range holds candidate i
#pragma unroll
                for (int b = 32; b > 0; b >>= 1) if (c + b < 64 && ra_pre[c + b] <= i) c += b;
                const int sq = ra_lo[c] + (i - ra_pre[c]);
                int k;
                if (len[c] > 0 && sq >= cmin_s[c]) k = P.cev[off[c] + sq - cmin_s[c]];
                else {      // before the chunk: the ring holds the member's last RB_RING events of the earlier chunks
                    const int before = len[c] > 0 ? cmin_s[c] : ctot_s[c];
                    k = (before - sq <= RB_RING) ? __ldcg(P.gchain + c * RB_RING + (sq & (RB_RING - 1))) : -1;
                }
                if (k < 0) continue;
                u64 mask = 0;
#pragma unroll
                for (int j = 0; j < NC; j++) {
                    const int cc = lane + 32 * j;
                    const int wv = cc < M ? wrow(rmin, cc) : -1;
                    const int v = cc < M ? __ldcg(P.row + (size_t)k * M + cc) : -1;
                    mask |= (u64)__ballot_sync(0xffffffffu, wv >= 0 && v >= wv) << (32 * j);
                }
                if (lane == 0) P.sc[k] = make_ulonglong2(mask, mask ^ key);
            }
        }
        grid.sync();
        // ---- test the pending windows
        for (int e = gw; e < M * L; e += nw) {
            const int c = e / L, j = e % L;
            if (pos[c] < len[c] && cur[c] == rmin && pos[c] + j < len[c]) {
                const int h = P.cev[off[c] + pos[c] + j];
                const long long e0 = clock64();
                const int hit = P.p0[h] >= 0 ? eval(h, rmin, j > 0) : 0;
                const long long e1 = clock64() - e0;
                c_maxev = e1 > c_maxev ? e1 : c_maxev; c_sumev += e1;
                if (P.dbg && lane == 0 && e1 > 20000) {        // census of the slow tests
                    atomicAdd((unsigned long long *)P.dbg + 9, 1ull);
                    atomicAdd((unsigned long long *)P.dbg + 10, (unsigned long long)j);
                    atomicAdd((unsigned long long *)P.dbg + 11, (unsigned long long)(hit == 2));
                    atomicAdd((unsigned long long *)P.dbg + 12, (unsigned long long)c_lastmiss);
                }
                if (lane == 0) res[c * RB_LMAX + j] = (uint8_t)hit;
                c_nev++;
            }
        }
        const long long t1 = clock64();
        grid.sync();
        const long long t2 = clock64();
        // ---- identical bookkeeping in every CTA (only CTA 0 writes the global results)
        int ft = -1, win = 0;
        bool mine = false;
        if (tid < M && pos[tid] < len[tid] && cur[tid] == rmin) {
            mine = true;
            win = min(L, len[tid] - pos[tid]);
            int unk = win;                                     // first position left untested (result 2)
#pragma unroll
            for (int q4 = RB_LMAX / 16 - 1; q4 >= 0; q4--) {
                const uint4 a = __ldcg(reinterpret_cast<const uint4 *>(res + tid * RB_LMAX) + q4);
                const unsigned wds[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int q = 3; q >= 0; q--)
#pragma unroll
                    for (int bb = 3; bb >= 0; bb--) {
                        const int idx = q4 * 16 + q * 4 + bb;
                        const unsigned v = (wds[q] >> (8 * bb)) & 0xff;
                        if (idx < win && v == 1) ft = idx;
                        if (idx < win && v == 2) unk = idx;
                    }
            }
            if (ft >= 0 && unk < ft) ft = -1;                  // an untested event precedes the first hit
            if (ft < 0) win = unk;                             // only the tested prefix is final
        }
        if (tid == 0) s_newtop = 0;
        if (tid < 64) { s_nfin[tid] = 0; s_base[tid] = 0; }
        __syncthreads();
        if (mine && ft >= 0 && rmin + 1 > rtop) s_newtop = 1;
        __syncthreads();
        if (s_newtop) {                                        // open the shared-memory row of round rmin+1
            rtop = rmin + 1;
            if (tid < 64) Wl[rtop & (RB_WR - 1)][tid] = -1;
            if (rtop >= P.Rcap && tid == 0 && lead) atomicMin(&P.scal[SC_ERR], -5);
        }
        __syncthreads();
        if (mine) {
            const int c = tid, o = off[c] + pos[c];
            const int nfinal = ft >= 0 ? ft : win;
            s_nfin[c] = nfinal; s_base[c] = o;
            pos[c] += nfinal;
            if (ft >= 0) {
                const int hnew = P.cev[o + ft];
                cur[c] = rmin + 1;
                if (rmin + 1 < P.Rcap) {
                    Wl[(rmin + 1) & (RB_WR - 1)][c] = hnew;
                    if (lead) P.Wf[(size_t)(rmin + 1) * M + c] = hnew;
                }
            }
        }
        __syncthreads();
        if (lead)                                              // final rounds of the events before the first hit
            for (int i = tid; i < M * RB_LMAX; i += blockDim.x) {
                const int c = i / RB_LMAX, j = i % RB_LMAX;
                if (j < s_nfin[c]) P.round[P.cev[s_base[c] + j]] = rmin;
            }
        __syncthreads();
        c_eval += t1 - t0; c_sync += t2 - t1; c_upd += clock64() - t2; c_steps++;
    }
    if (P.dbg && lane == 0) {
        atomicMax((unsigned long long *)P.dbg + 5, (unsigned long long)c_maxev);
        atomicAdd((unsigned long long *)P.dbg + 6, (unsigned long long)c_miss);
        atomicAdd((unsigned long long *)P.dbg + 7, (unsigned long long)c_sumev);
        atomicAdd((unsigned long long *)P.dbg + 13, (unsigned long long)c_nev);
    }
    if (P.dbg && lane == 0 && warp == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x / 2)) {
        unsigned long long *o = (unsigned long long *)P.dbg + (blockIdx.x == 0 ? 0 : 8);
        atomicAdd(&o[0], (unsigned long long)c_eval); atomicAdd(&o[1], (unsigned long long)c_sync);
        atomicAdd(&o[2], (unsigned long long)c_upd); atomicAdd(&o[3], (unsigned long long)c_nev);
        atomicAdd(&o[4], (unsigned long long)c_steps);
    }
    if (lead && tid == 0 && P.n > 0) P.scal[SC_MAX_ROUND] = rtop;
}

// ---- the reference's witness flags / witnesses table from the finished rounds (swirld.py:221-222, 196-197)
__global__ void k_rb_witness(RbParams P) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < P.n; j += gridDim.x * blockDim.x) {
        const int h = P.first + j, pa = P.p0[h], r = P.round[h];
        const bool wit = pa < 0 || r > P.round[pa];
        P.wit[h] = wit ? 1 : 0;
        if (wit && r >= 0 && r < P.Rcap) P.W[(size_t)r * P.M + P.creator[h]] = h;
    }
}
// ---- SM(h) = {c_ : W[round h][c_] >= 0 and row(h)[c_] >= W[round h][c_]}, one warp per event
template <int NC>
__global__ void __launch_bounds__(256) k_rb_seenmask(RbParams P) {
    const int lane = threadIdx.x & 31;
    const int j0 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j0 >= P.n) return;
    const int h = P.first + j0, r = P.round[h], M = P.M;
    u64 mask = 0;
#pragma unroll
    for (int j = 0; j < NC; j++) {
        const int c = lane + 32 * j;
        int w = -1, v = -1;
        if (c < M && r >= 0 && r < P.Rcap) { w = P.W[(size_t)r * M + c]; v = P.row[(size_t)h * M + c]; }
        mask |= (u64)__ballot_sync(0xffffffffu, w >= 0 && v >= w) << (32 * j);
    }
    if (lane == 0) P.SM[h] = mask;
}
