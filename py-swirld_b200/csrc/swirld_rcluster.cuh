// swirld_rcluster.cuh -- the round numbers of a chunk (Node.divide_rounds, swirld.py:187-222) inside ONE thread-block
// cluster, M <= 64.
//
// swirld_rounds.cuh advances all member chains one round per grid-wide step; a step there is ~8 dependent trips
// through L2 (rows, mask cache, atomics, the grid barrier, the results): 6 us, 1461 times per million events.  The
// work of a step is small (a few thousand masks and tests), so this kernel keeps everything a step touches in the
// shared memory of 16 CTAs and pays distributed-shared-memory latency (~200 cycles) and two cluster barriers instead:
//
//   * rows in SEQ space: rs(h)[c] = the chain position (swirld.py's implicit per-creator sequence number) of the
//     event of member c that h sees (k_rc_seqrows, from the can_see table).  Every comparison of the scheme
//     ("row(k)[c_] >= Wf_r[c_]") holds in seq space as it does in index space -- a member's events are ordered the
//     same way in both -- and a mask S_r(k) is addressed by (member, position) without any lookup.
//   * CTA q owns chains 4q..4q+3: a window of RC_WN consecutive rows per chain in its shared memory, filled one
//     step ahead by cp.async.
//   * a step (round r = lowest open round):
//       a  the owner computes S_r of its members' events [Wls_r[m], mend[m]) from its window and stores each mask
//          into the mask table of ALL 16 CTAs (st.shared::cluster);                       cluster barrier
//       b  the owner finds each chain's first pending event that passes P_r: P_r, and "an event that sees beyond
//          the prepared masks", are monotone along a chain, so a 5-ary search with the chain's 4 warps needs 3
//          passes of one test per warp (the test of swirld_rounds.cuh, masks from LOCAL shared memory);
//       c  the results go to every CTA (64 words);                                       cluster barrier
//       d  identical bookkeeping in every CTA (positions, rounds, the seq-space mirror of Wf), the owner stores the
//          final rounds, the window slides.
//   * whatever the windows cannot decide -- no progress for RC_STALL steps, a chain more than RB_WR rounds behind,
//     rows further before the chunk than the ring -- hands the REST of the chunk to k_rounds_batch through `cont`
//     (positions and rounds per chain).  tests/test_rounds_cluster_model.py is the executable model.
#pragma once
#include "swirld_rounds.cuh"

#define RC_CS 16            // CTAs per cluster
#define RC_CPC 4            // chains per CTA
#define RC_WPC 4            // warps per chain
#define RC_THREADS 512
#define RC_LW 32            // pending events searched per chain and step
#define RC_WN 128           // rows per chain in shared memory (power of two)
#define RC_MR 64            // masks per member and step
#define RC_PF 32            // rows loaded beyond the searched window
#define RC_PASSES 3         // (RC_WPC + 1) ^ RC_PASSES >= RC_LW + 1
#define RC_STALL 3
#define RC_REACH (RC_WN / 2) // rows before the chunk that a launch may need (from the ring)
#define RC_INF 0x7fffffff

struct RcParams {
    RbParams R;
    const int32_t *rsg;      // [n][64] seq-space rows of the chunk's events in cev order (k_rc_seqrows)
    int32_t *cont;           // [0,64) positions, [64,128) rounds, [128] 1 = k_rounds_batch has work left
};

#define RC_SMEM_ROWS ((size_t)RC_CPC * RC_WN * 64 * 4)
#define RC_SMEM_MASK ((size_t)64 * RC_MR * 8)
#define RC_SMEM_WLS ((size_t)RB_WR * 64 * 4)
#define RC_SMEM_BYTES (RC_SMEM_ROWS + RC_SMEM_MASK + RC_SMEM_WLS + 512 + 8192)

// seq-space rows of the chunk, grouped by creator like cev: one warp per event
__global__ void __launch_bounds__(256) k_rc_seqrows(RbParams P, int32_t *rsg) {
    const int lane = threadIdx.x & 31;
    for (int j = blockIdx.x * 8 + (threadIdx.x >> 5); j < P.n; j += gridDim.x * 8) {
        const int h = P.cev[P.first + j];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int c = lane + 32 * u;
            const int v = c < P.M ? P.row[(size_t)h * P.M + c] : -1;
            rsg[(size_t)j * 64 + c] = v < 0 ? -1 : P.seq[v];
        }
    }
}

__device__ __forceinline__ unsigned rc_cta_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void rc_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned rc_map(const void *p, unsigned rank) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ void rc_st_u64(unsigned addr, u64 v) { asm volatile("st.shared::cluster.u64 [%0], %1;" :: "r"(addr), "l"(v) : "memory"); }
__device__ __forceinline__ void rc_st_u32(unsigned addr, unsigned v) { asm volatile("st.shared::cluster.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void rc_cp16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

__device__ __forceinline__ void rc_cp4(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void rc_st_v4(unsigned addr, uint4 v) {
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// carry-save adder over 64 bit columns: a + b + c = 2 * h + l
__device__ __forceinline__ void rc_csa(u64 &h, u64 &l, u64 a, u64 b, u64 c) {
    const u64 u = a ^ b;
    h = (a & b) | (u & c);
    l = u ^ c;
}
// bit-sliced a[0..NB) += the same number of the lane `delta` away; the sum has NB + 1 bits
template <int NB>
__device__ __forceinline__ void rc_vadd_xor(u64 (&a)[8], int delta) {
    u64 carry = 0;
#pragma unroll
    for (int k = 0; k < NB; k++) {
        const u64 b = __shfl_xor_sync(0xffffffffu, a[k], delta);
        u64 h, l;
        rc_csa(h, l, a[k], b, carry);
        a[k] = l; carry = h;
    }
    a[NB] = carry;
}

template <bool UNIT>
__device__ __forceinline__ void rounds_cluster_body(const RcParams &Q) {
    const RbParams &P = Q.R;
    extern __shared__ __align__(16) unsigned char rc_smem[];
    int (*rsw)[RC_WN][64] = reinterpret_cast<int (*)[RC_WN][64]>(rc_smem);                       // [chain][slot][member]
    u64 (*maskbuf)[RC_MR] = reinterpret_cast<u64 (*)[RC_MR]>(rc_smem + RC_SMEM_ROWS);            // [member][offset]
    int (*Wls)[64] = reinterpret_cast<int (*)[64]>(rc_smem + RC_SMEM_ROWS + RC_SMEM_MASK);       // seq of Wf_r[c], -1: none
    i64 *stake_s = reinterpret_cast<i64 *>(rc_smem + RC_SMEM_ROWS + RC_SMEM_MASK + RC_SMEM_WLS);
    int *iv = reinterpret_cast<int *>(stake_s + 64);
    int *cur = iv, *pos = iv + 64, *len = iv + 128, *off = iv + 192, *cmin_s = iv + 256, *ctot_s = iv + 320;
    int *wlo = iv + 384, *wld = iv + 448, *wrd = iv + 512, *slo = iv + 576, *smend = iv + 640, *swin = iv + 704;
    int *xres = iv + 768, *s_nfin = iv + 832, *s_base = iv + 896, *s_old = iv + 960, *coff_s = iv + 1024;
    int *sa = iv + 1088, *sb = sa + RC_CPC, *svb = sb + RC_CPC, *tres = svb + RC_CPC;             // tres[RC_CPC][RC_WPC]
    int (*cevw)[RC_WN] = reinterpret_cast<int (*)[RC_WN]>(iv + 1152);                           // event index of every window row
    int *vres = iv + 1152 + RC_CPC * RC_WN;                                                      // [RC_CPC][RC_LW] test results

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = P.M;
    const int bx = (int)rc_cta_rank();
    const bool lead = bx == 0;
    const i64 thr = P.tot2 / 3;

    int rtop = max(P.scal[SC_MAX_ROUND], 0);
    for (int i = tid; i < RB_WR * 64; i += RC_THREADS) {
        const int slot = i >> 6, c = i & 63;
        const int r = rtop - ((rtop - slot) & (RB_WR - 1));
        const int w = (c < M && r >= 0 && r < P.Rcap) ? __ldcg(P.Wf + (size_t)r * M + c) : -1;
        Wls[slot][c] = w >= 0 ? P.seq[w] : -1;
    }
    if (tid < 64) {
        const int c = tid;
        stake_s[c] = c < M ? P.stake[c] : 0;
        int o = 0, l = 0, cu = RC_INF, co = 0;
        if (c < M) {
            co = P.coff[c]; o = P.first + co; l = P.coff[c + 1] - co;
            if (l > 0) {
                const int h0 = P.cev[o], pa = P.p0[h0];
                cu = pa < 0 ? 0 : P.round[pa];
            }
        }
        off[c] = o; len[c] = l; pos[c] = 0; cur[c] = cu; coff_s[c] = co;
        cmin_s[c] = (c < M && l > 0) ? P.cmin[c] : 0; ctot_s[c] = c < M ? P.ctot[c] : 0;
    }
    __syncthreads();
    if (tid < M && len[tid] > 0 && cur[tid] == 0) {            // a member's root opens round 0 for it
        const int h0 = P.cev[off[tid]];
        if (P.p0[h0] < 0) {
            if (rtop < RB_WR) Wls[0][tid] = 0;
            if (lead) P.Wf[tid] = h0;
        }
    }
    __syncthreads();

    auto lowest_open = [&]() -> int {                           // every warp for itself: the state is identical in all CTAs
        int r = RC_INF;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int c = lane + 32 * j;
            if (pos[c] < len[c]) r = min(r, cur[c]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r = min(r, __shfl_xor_sync(0xffffffffu, r, o));
        return r;
    };
    auto seqpos = [&](int c) -> int { return len[c] > 0 ? cmin_s[c] + pos[c] : ctot_s[c]; };

    int handed = 0;
    // ---- the windows at launch: [wlo, wld) from the chunk's seq rows, what precedes the chunk through the ring
    {
        const int rmin = lowest_open();
        bool bad = false;
        if (rmin != RC_INF && rmin <= rtop - RB_WR) bad = true;
        if (tid < 64 && rmin != RC_INF && !bad) {
            const int c = tid, sp = seqpos(c);
            const int lo = Wls[rmin & (RB_WR - 1)][c];
            const int wl = lo >= 0 ? min(lo, sp) : sp;
            const int before = len[c] > 0 ? cmin_s[c] : ctot_s[c];
            if (before - wl > RC_REACH) bad = true;
            wlo[c] = wl;
            wld[c] = wrd[c] = min(min(ctot_s[c], wl + RC_WN), sp + RC_LW + RC_PF);
        }
        if (__syncthreads_or(bad)) handed = 1;
        if (!handed && rmin != RC_INF) {
            for (int cl = 0; cl < RC_CPC; cl++) {
                const int c = bx * RC_CPC + cl;
                const int before = len[c] > 0 ? cmin_s[c] : ctot_s[c];
                for (int sq = wlo[c] + warp; sq < wld[c]; sq += RC_THREADS / 32) {
                    int v0, v1;
                    if (sq >= before) {
                        const int32_t *src = Q.rsg + (size_t)(coff_s[c] + sq - cmin_s[c]) * 64;
                        v0 = __ldcg(src + lane); v1 = __ldcg(src + lane + 32);
                        if (lane == 0) cevw[cl][sq & (RC_WN - 1)] = P.cev[off[c] + sq - cmin_s[c]];
                    } else {
                        const int h = __ldcg(P.gchain + c * RB_RING + (sq & (RB_RING - 1)));
                        const int a0 = lane < M ? __ldcg(P.row + (size_t)h * M + lane) : -1;
                        const int a1 = lane + 32 < M ? __ldcg(P.row + (size_t)h * M + lane + 32) : -1;
                        v0 = a0 < 0 ? -1 : P.seq[a0]; v1 = a1 < 0 ? -1 : P.seq[a1];
                    }
                    rsw[cl][sq & (RC_WN - 1)][lane] = v0; rsw[cl][sq & (RC_WN - 1)][lane + 32] = v1;
                }
            }
        }
        __syncthreads();
    }
    rc_cluster_sync();                                          // every CTA of the cluster runs before any remote store

    long long c_t[6] = {0, 0, 0, 0, 0, 0}, c_steps = 0, c_tests = 0, c_unk = 0;
    int stall = 0;
    while (!handed) {
        const long long t0 = clock64();
        const int rmin = lowest_open();
        if (rmin == RC_INF) break;                              // every chain is done
        if (rmin <= rtop - RB_WR) { handed = 1; break; }
        const int slot = rmin & (RB_WR - 1);
        if (tid < 64) {
            const int c = tid, lo = Wls[slot][c], sp = seqpos(c);
            const bool open = pos[c] < len[c];
            int me = -1;
            if (lo >= 0) me = min(min(wrd[c], open ? sp + RC_LW : ctot_s[c]), lo + RC_MR);
            slo[c] = lo; smend[c] = me;
            swin[c] = (open && cur[c] == rmin) ? max(0, min(min(RC_LW, len[c] - pos[c]), wrd[c] - sp)) : -1;
        }
        __syncthreads();
        const long long t1 = clock64();
        // ---- a: the masks of my members' ranges into my own table, then each member's masks to every other CTA as
        //         one wide store per (member, CTA)
        {
            const int w0 = slo[lane], w1 = slo[lane + 32];
            int cnt[RC_CPC], total = 0;
#pragma unroll
            for (int cl = 0; cl < RC_CPC; cl++) {
                const int c = bx * RC_CPC + cl;
                cnt[cl] = max(0, smend[c] - slo[c]);
                total += cnt[cl];
            }
            for (int item = warp; item < total; item += RC_THREADS / 32) {
                int cl = 0, i = item;
#pragma unroll
                for (int q = 0; q < RC_CPC - 1; q++) if (cl == q && i >= cnt[q]) { i -= cnt[q]; cl = q + 1; }
                const int c = bx * RC_CPC + cl, s = slo[c] + i;
                const int *row = rsw[cl][s & (RC_WN - 1)];
                const int v0 = row[lane], v1 = row[lane + 32];
                const u64 mask = (u64)__ballot_sync(0xffffffffu, w0 >= 0 && v0 >= w0) |
                                 (u64)__ballot_sync(0xffffffffu, w1 >= 0 && v1 >= w1) << 32;
                if (lane == 0) maskbuf[c][i] = mask;
            }
            __syncthreads();
            for (int pair = warp; pair < RC_CPC * RC_CS; pair += RC_THREADS / 32) {
                const int cl = pair & (RC_CPC - 1), r = pair / RC_CPC, c = bx * RC_CPC + cl;
                if (r == bx || 2 * lane >= cnt[cl]) continue;
                const uint4 val = *reinterpret_cast<const uint4 *>(&maskbuf[c][2 * lane]);
                rc_st_v4(rc_map(&maskbuf[c][2 * lane], (unsigned)r), val);
            }
        }
        const long long t2 = clock64();
        rc_cluster_sync();
        const long long t3 = clock64();
        // ---- b: first pending event with P_r (1) or beyond the masks (2), per chain
        if (UNIT) {
            // every position of the window at once, bit-sliced: 8 tests per warp, 4 lanes x 16 members per test.  A lane
            // adds the masks of its live members into a vertical counter (one bit plane per power of two, 64 columns
            // wide), the 4 lanes of a test add their counters, and the column counts are compared with the threshold
            // plane by plane -- no transposes, ~50 instructions per test.
            const int cl = warp / RC_WPC, c = bx * RC_CPC + cl;
            const int q = lane >> 2, pp = lane & 3, t = (warp % RC_WPC) * 8 + q;
            const int win = swin[c];
            const bool act = t < win;
            const int thr_i = (int)thr;
            int v = 0;
            if (__any_sync(0xffffffffu, act)) {
                c_tests++;
                const int sq = cmin_s[c] + pos[c] + t;
                const int *row = rsw[cl][sq & (RC_WN - 1)];
                const int rot = 2 * q + (pp >> 1);             // (the 32 lanes read 32 different banks)
                int offs[16];
                int lv = 0;
                bool unk = false;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int m = 16 * pp + ((i + rot) & 15);
                    const int pr = act ? (m == c ? sq - 1 : row[m]) : -1;   // the own column is set back to the self-parent
                    const int W = slo[m];
                    const bool live = act && W >= 0 && pr >= W;
                    if (live && pr >= smend[m]) unk = true;
                    offs[i] = live ? pr - W : -1;
                    lv += live ? 1 : 0;
                }
                lv += __shfl_xor_sync(0xffffffffu, lv, 1);
                lv += __shfl_xor_sync(0xffffffffu, lv, 2);
                const unsigned ub = __ballot_sync(0xffffffffu, unk);
                unk = ((ub >> (lane & ~3)) & 0xfu) != 0;
                const bool need = act && lv > thr_i && !unk;   // (lv <= thr: hits[c_] <= the live members)
                if (act && lv > thr_i && unk) { v = 2; c_unk++; }
                if (__any_sync(0xffffffffu, need)) {
                    u64 x[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const int m = 16 * pp + ((i + rot) & 15);
                        x[i] = (need && offs[i] >= 0) ? maskbuf[m][offs[i]] : 0ull;
                    }
                    u64 a[8], t2a, t2b, t4a, t4b, t8a, t8b;
                    a[0] = a[1] = a[2] = a[3] = 0;
                    rc_csa(t2a, a[0], a[0], x[0], x[1]);   rc_csa(t2b, a[0], a[0], x[2], x[3]);   rc_csa(t4a, a[1], a[1], t2a, t2b);
                    rc_csa(t2a, a[0], a[0], x[4], x[5]);   rc_csa(t2b, a[0], a[0], x[6], x[7]);   rc_csa(t4b, a[1], a[1], t2a, t2b);
                    rc_csa(t8a, a[2], a[2], t4a, t4b);
                    rc_csa(t2a, a[0], a[0], x[8], x[9]);   rc_csa(t2b, a[0], a[0], x[10], x[11]); rc_csa(t4a, a[1], a[1], t2a, t2b);
                    rc_csa(t2a, a[0], a[0], x[12], x[13]); rc_csa(t2b, a[0], a[0], x[14], x[15]); rc_csa(t4b, a[1], a[1], t2a, t2b);
                    rc_csa(t8b, a[2], a[2], t4a, t4b);
                    rc_csa(a[4], a[3], a[3], t8a, t8b);       // 16 a[4] + 8 a[3] + 4 a[2] + 2 a[1] + a[0] = members per column
                    rc_vadd_xor<5>(a, 1);
                    rc_vadd_xor<6>(a, 2);                      // 7 planes: 0..64 per column
                    u64 gt = 0, eq = ~0ull;
#pragma unroll
                    for (int k = 6; k >= 0; k--) {
                        const u64 tk = ((thr_i >> k) & 1) ? ~0ull : 0ull;
                        gt |= eq & a[k] & ~tk;
                        eq &= ~(a[k] ^ tk);
                    }
                    if (need) v = __popcll(gt) > thr_i ? 1 : 0; // a COUNT of members against the STAKE threshold (quirk Q3)
                }
            }
            if (pp == 0) vres[cl * RC_LW + t] = act ? v : 0;
            __syncthreads();
            if (warp < RC_CPC) {
                const int w2 = swin[bx * RC_CPC + warp];
                const int x = vres[warp * RC_LW + lane];
                const unsigned inwin = w2 >= 32 ? 0xffffffffu : (w2 > 0 ? (1u << w2) - 1u : 0u);
                const unsigned nz = __ballot_sync(0xffffffffu, x != 0) & inwin;
                const int f = nz ? __ffs(nz) - 1 : max(w2, 0);
                const int vf = __shfl_sync(0xffffffffu, x, f & 31);
                if (lane == 0) { sb[warp] = f; svb[warp] = nz ? vf : 0; }
            }
            __syncthreads();
        } else {
            // integer stakes: a 5-ary search with the chain's 4 warps, one test of swirld_rounds.cuh's kind per warp and pass
            if (tid < RC_CPC) { sa[tid] = -1; sb[tid] = max(swin[bx * RC_CPC + tid], 0); svb[tid] = 0; }
            __syncthreads();
            for (int pass = 0; pass < RC_PASSES; pass++) {
                const int cl = warp / RC_WPC, i = warp % RC_WPC, c = bx * RC_CPC + cl;
                const int a = sa[cl], b = sb[cl], nun = b - a - 1;
                int t = -1;
                if (swin[c] > 0 && nun > 0) {
                    if (nun < RC_WPC) { if (i < nun) t = a + 1 + i; }
                    else t = a + ((i + 1) * (nun + 1)) / (RC_WPC + 1);
                }
                int v = 0;
                if (t >= 0) {                                       // (warp-uniform)
                    c_tests++;
                    const int sq = cmin_s[c] + pos[c] + t;
                    const int *row = rsw[cl][sq & (RC_WN - 1)];
                    int pre[2], W[2];
                    bool live[2];
                    i64 lv = 0;
    #pragma unroll
                    for (int j = 0; j < 2; j++) {
                        const int m = lane + 32 * j;
                        pre[j] = m == c ? sq - 1 : row[m];          // the own column is set back to the self-parent
                        W[j] = slo[m];
                        live[j] = W[j] >= 0 && pre[j] >= W[j];
                        if (UNIT) lv += __popc(__ballot_sync(0xffffffffu, live[j]));
                        else {
                            i64 s = live[j] ? stake_s[m] : 0;
    #pragma unroll
                            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                            lv += s;
                        }
                    }
                    if (lv > thr) {                                 // else: hits[c_] <= stake of the live members
                        const bool unk = (live[0] && pre[0] >= smend[lane]) || (live[1] && pre[1] >= smend[lane + 32]);
                        if (__any_sync(0xffffffffu, unk)) { v = 2; c_unk++; }
                        else {
                            u64 mm[2];
    #pragma unroll
                            for (int j = 0; j < 2; j++) mm[j] = live[j] ? maskbuf[lane + 32 * j][pre[j] - W[j]] : 0ull;
                            unsigned T[2][2];                       // T[jj][j]: bit b = member jj*32+b sees column j*32+lane
    #pragma unroll
                            for (int jj = 0; jj < 2; jj++)
    #pragma unroll
                                for (int j = 0; j < 2; j++) T[jj][j] = rb_transpose32((unsigned)(mm[jj] >> (32 * j)), lane);
                            int cntc = 0;
    #pragma unroll
                            for (int j = 0; j < 2; j++) {
                                i64 hits = 0;
                                if (UNIT) hits = __popc(T[0][j]) + __popc(T[1][j]);
                                else {
    #pragma unroll
                                    for (int jj = 0; jj < 2; jj++)
    #pragma unroll 8
                                        for (int bq = 0; bq < 32; bq++) hits += ((T[jj][j] >> bq) & 1) ? stake_s[jj * 32 + bq] : 0;
                                }
                                cntc += __popc(__ballot_sync(0xffffffffu, hits > thr));
                            }
                            v = (i64)cntc > thr ? 1 : 0;           // a COUNT of members against the STAKE threshold (quirk Q3)
                        }
                    }
                }
                if (lane == 0) tres[cl * RC_WPC + i] = t >= 0 ? (t << 2 | v) : -1;
                __syncthreads();
                if (tid < RC_CPC) {
                    int a2 = sa[tid], b2 = sb[tid], vb2 = svb[tid];
    #pragma unroll
                    for (int q = 0; q < RC_WPC; q++) {
                        const int x = tres[tid * RC_WPC + q];
                        if (x < 0) continue;
                        const int tq = x >> 2, vq = x & 3;
                        if (vq == 0) a2 = max(a2, tq);
                        else if (tq < b2) { b2 = tq; vb2 = vq; }
                    }
                    sa[tid] = a2; sb[tid] = b2; svb[tid] = vb2;
                }
                __syncthreads();
            }
        }
        const long long t4 = clock64();
        // ---- c: (first position, what it is) of my chains to every CTA
        if (tid < RC_CPC * RC_CS) {
            const int cl = tid & (RC_CPC - 1), rank = tid / RC_CPC, c = bx * RC_CPC + cl;
            unsigned x = 0;
            if (swin[c] >= 0) { const int f = sb[cl]; x = (unsigned)(f << 2 | (f < swin[c] ? svb[cl] : 0)); }
            rc_st_u32(rc_map(&xres[c], (unsigned)rank), x);
        }
        rc_cluster_sync();
        const long long t5 = clock64();
        // ---- d: identical bookkeeping in every CTA
        bool tested = false;
        int f = 0, vf = 0;
        if (tid < 64 && swin[tid] >= 0) { tested = true; const int x = xres[tid]; f = x >> 2; vf = x & 3; }
        if (__syncthreads_or(tested && vf == 1 && rmin + 1 > rtop)) {       // open the mirror row of round rmin+1
            rtop = rmin + 1;
            if (tid < 64) Wls[rtop & (RB_WR - 1)][tid] = -1;
            if (rtop >= P.Rcap && tid == 0 && lead) atomicMin(&P.scal[SC_ERR], -5);
            __syncthreads();
        }
        if (tid < 64) {
            const int c = tid;
            int nf = 0, base = 0;
            if (tested) {
                const int op = pos[c];
                base = cmin_s[c] + op; nf = f;
                if (vf == 1) {
                    cur[c] = rmin + 1;
                    if (rmin + 1 < P.Rcap) {
                        Wls[(rmin + 1) & (RB_WR - 1)][c] = cmin_s[c] + op + f;
                        if (c / RC_CPC == bx) P.Wf[(size_t)(rmin + 1) * M + c] = cevw[c % RC_CPC][(cmin_s[c] + op + f) & (RC_WN - 1)];
                    }
                }
                pos[c] = op + f;
            }
            s_nfin[c] = nf; s_base[c] = base;
        }
        const bool prog = __syncthreads_or(tested && (f > 0 || vf == 1));
        if (tid < RC_CPC * RC_LW) {                             // final rounds of my chains' events before the first hit
            const int cl = tid / RC_LW, j = tid % RC_LW, c = bx * RC_CPC + cl;
            if (j < s_nfin[c]) P.round[cevw[cl][(s_base[c] + j) & (RC_WN - 1)]] = rmin;
        }
        // ---- the windows: what was issued a step ago has arrived; issue the next rows
        asm volatile("cp.async.wait_all;" ::: "memory");
        const int rnext = lowest_open();
        bool moved = false;
        if (tid < 64) {
            const int c = tid, sp = seqpos(c);
            moved = wrd[c] != wld[c];
            wrd[c] = wld[c];
            if (rnext != RC_INF && rnext > rtop - RB_WR) {
                const int l2 = Wls[rnext & (RB_WR - 1)][c];
                wlo[c] = max(wlo[c], l2 >= 0 ? min(l2, sp) : sp);
            }
            const int hi = min(min(ctot_s[c], wlo[c] + RC_WN), sp + RC_LW + RC_PF);
            s_old[c] = wld[c];
            if (hi > wld[c]) { wld[c] = hi; moved = true; }
        }
        const bool mv = __syncthreads_or(moved);                // (also: the arrived rows are visible to all warps)
#pragma unroll
        for (int cl = 0; cl < RC_CPC; cl++) {
            const int c = bx * RC_CPC + cl, lo2 = s_old[c], nrow = wld[c] - lo2;
            for (int i = tid; i < nrow * 16; i += RC_THREADS) {
                const int sq = lo2 + (i >> 4), q = i & 15;
                rc_cp16(&rsw[cl][sq & (RC_WN - 1)][q * 4], Q.rsg + (size_t)(coff_s[c] + sq - cmin_s[c]) * 64 + q * 4);
                if (q == 0) rc_cp4(&cevw[cl][sq & (RC_WN - 1)], P.cev + off[c] + sq - cmin_s[c]);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        stall = (prog || mv) ? 0 : stall + 1;
        if (stall >= RC_STALL) handed = 1;
        const long long t6 = clock64();
        c_t[0] += t1 - t0; c_t[1] += t2 - t1; c_t[2] += t3 - t2; c_t[3] += t4 - t3; c_t[4] += t5 - t4; c_t[5] += t6 - t5;
        c_steps++;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    if (lead) {
        if (tid < 64) { Q.cont[tid] = pos[tid]; Q.cont[64 + tid] = cur[tid]; }
        if (tid == 0) {
            Q.cont[128] = handed;
            if (P.n > 0) P.scal[SC_MAX_ROUND] = rtop;
        }
    }
    if (P.dbg && lane == 0) {
        atomicAdd((unsigned long long *)P.dbg + 11, (unsigned long long)c_tests);     // (the slots of k_rounds_batch's counters)
        atomicAdd((unsigned long long *)P.dbg + 12, (unsigned long long)c_unk);
    }
    if (P.dbg && tid == 0 && lead) {
        unsigned long long *o = (unsigned long long *)P.dbg;
        for (int i = 0; i < 6; i++) atomicAdd(&o[i], (unsigned long long)c_t[i]);
        atomicAdd(&o[6], (unsigned long long)c_steps);
        atomicAdd(&o[7], 1ull);                                 // launches, and how many handed work back
        atomicAdd(&o[15], (unsigned long long)handed);
    }
    rc_cluster_sync();                                          // nobody leaves while its shared memory may still be written
}

template <bool UNIT>
__global__ void __launch_bounds__(RC_THREADS, 1) k_rounds_cluster(RcParams Q) { rounds_cluster_body<UNIT>(Q); }

// several independent node-views (swirld_rounds.cuh, k_rounds_batch_views): one cluster per view, as many side by side
// as the device holds -- the clusters never talk to each other, so this is an ordinary (non-cooperative) launch
template <bool UNIT>
__global__ void __launch_bounds__(RC_THREADS, 1) k_rounds_cluster_views(const RcParams *Qv) {
    __shared__ RcParams Qs;
    if (threadIdx.x == 0) Qs = Qv[blockIdx.x / RC_CS];
    __syncthreads();
    rounds_cluster_body<UNIT>(Qs);
}
