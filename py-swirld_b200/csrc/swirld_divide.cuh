// swirld_divide.cuh -- k_divide: can_see rows + round numbers + witness registration
// (Node.divide_rounds, swirld.py:187-222) for a chunk of events, one CTA.
//
// Execution model.  The chunk is walked in windows of 32 consecutive event indices,
// one warp per event, lane = member column (NC columns per lane).  The only true
// dependency of an event is on its two parents (always lower indices): a parent inside
// the same window is waited for on that event's mbarrier (hardware-suspended wait, no
// polling); parents from earlier windows are complete by the end-of-window barrier.
//
// Shared memory holds 256 "slots" (row, T matrix, round of one event):
//   slots [0,128)   ring of the last 4 windows, slot = event & 127 -- parents are almost
//                   always recent chain heads;
//   slots [128,256) staging, [window parity][warp][parent]: an older parent is copied
//                   from L2/HBM with cp.async one window AHEAD, off the critical path.
// So the compute path reads both parents from shared memory whatever their age.  An
// event publishes its slot, releases its mbarrier, and only then streams row / T /
// round / flags to HBM.
#pragma once
#include "swirld_kernels.cuh"

#define SW_RING 128
#define SW_RING_WINS (SW_RING / 32)

template <int NC>
struct __align__(16) DivSlot {
    int32_t row[NC * 32];
    u64 T[NC * 32];
    int32_t round;
    int32_t pad[3];
};

template <int NC>
struct DivSmem {
    DivSlot<NC> slot[SW_RING + 2 * 32 * 2];
    int32_t Wc[SW_WC][NC * 32];
    i64 stake[NC * 32];
    u64 mbar[32];
    int rmaxp[2];
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64 *b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_release(u64 *b) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.release.cta.shared::cta.b64 st, [%0];\n\t}"
                 :: "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire(u64 *b, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n\t}"
        :: "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// copy event p's (row, T, round) from the global tables into a staging slot
template <int NC>
__device__ __forceinline__ void stage_parent(const DivParams &P, DivSlot<NC> *dst, int p, int lane, const bool *act) {
#pragma unroll
    for (int j = 0; j < NC; j++)
        if (act[j]) {
            const size_t o = (size_t)p * P.M + lane + 32 * j;
            cp_async4(&dst->row[lane + 32 * j], P.row + o);
            cp_async8(&dst->T[lane + 32 * j], P.T + o);
        }
    if (lane == 0) cp_async4(&dst->round, P.round + p);
}

template <int NC, bool UNIT>
__global__ void __launch_bounds__(1024, 1) k_divide(DivParams P) {
    extern __shared__ __align__(16) unsigned char smraw[];
    DivSmem<NC> &S = *reinterpret_cast<DivSmem<NC> *>(smraw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = P.M;
    constexpr int MS = NC * 32;
    bool act[NC];
#pragma unroll
    for (int j = 0; j < NC; j++) act[j] = lane + 32 * j < M;
    const i64 thr = P.tot2 / 3;             // 3*x > 2*tot  <=>  x > floor(2*tot/3) for integers

    if (tid < 32) mbar_init(&S.mbar[tid], 1);
    if (tid < MS) S.stake[tid] = tid < M ? P.stake[tid] : 0;
    // padded columns of every slot stay (-1, 0) for the whole launch
    for (int i = tid; i < (SW_RING + 128) * MS; i += 1024) {
        S.slot[i / MS].row[i % MS] = -1;
        S.slot[i / MS].T[i % MS] = 0;
    }
    int rmax = P.scal[SC_MAX_ROUND];
    int wbase = max(0, rmax - (SW_WC / 2 - 1));
    if (tid < 2) S.rmaxp[tid] = rmax;
    for (int i = tid; i < SW_WC * MS; i += 1024) {
        const int slot = i / MS, c = i % MS;
        const int r = wbase + ((slot - (wbase % SW_WC) + SW_WC) % SW_WC);
        S.Wc[slot][c] = (c < M && r < P.Rcap) ? P.W[(size_t)r * M + c] : -1;
    }
    __syncthreads();

    const int first = P.first, last = P.first + P.n;
    const int w_first = first >> 5, w_last = (last - 1) >> 5;

    // software pipeline: (pa,pb,cr) of the current window, (pa1,..) of the next one,
    // loads for the one after that in flight
    int h = w_first * 32 + warp;
    int pa = -1, pb = -1, cr = 0, pa1 = -1, pb1 = -1, cr1 = 0;
    bool valid = h >= first && h < last;
    if (valid) { pa = __ldg(P.p0 + h); pb = __ldg(P.p1 + h); cr = __ldg(P.creator + h); }
    bool valid1 = w_first < w_last && h + 32 < last;
    if (valid1) { pa1 = __ldg(P.p0 + h + 32); pb1 = __ldg(P.p1 + h + 32); cr1 = __ldg(P.creator + h + 32); }
    // staging index of an out-of-ring parent of the event in window `w`, else -1
    auto stage_of = [&](int p, int w, int which) -> int {
        if (p < 0) return -1;
        const bool in_ring = p >= first && (w - (p >> 5)) < SW_RING_WINS;
        return in_ring ? -1 : SW_RING + ((w & 1) * 32 + warp) * 2 + which;
    };
    int sa = valid ? stage_of(pa, w_first, 0) : -1, sb = valid ? stage_of(pb, w_first, 1) : -1;
    if (sa >= 0) stage_parent<NC>(P, &S.slot[sa], pa, lane, act);
    if (sb >= 0) stage_parent<NC>(P, &S.slot[sb], pb, lane, act);
    cp_async_commit();

    for (int win = w_first; win <= w_last; ++win) {
        // ---- stage the NEXT window's old parents, fetch the one after that
        const int sa1 = valid1 ? stage_of(pa1, win + 1, 0) : -1, sb1 = valid1 ? stage_of(pb1, win + 1, 1) : -1;
        if (sa1 >= 0) stage_parent<NC>(P, &S.slot[sa1], pa1, lane, act);
        if (sb1 >= 0) stage_parent<NC>(P, &S.slot[sb1], pb1, lane, act);
        cp_async_commit();
        const int h2 = h + 64;
        const bool valid2 = win + 2 <= w_last && h2 < last;
        int pa2 = -1, pb2 = -1, cr2 = 0;
        if (valid2) { pa2 = __ldg(P.p0 + h2); pb2 = __ldg(P.p1 + h2); cr2 = __ldg(P.creator + h2); }

        const unsigned parity = (unsigned)(win - w_first) & 1u;
        if (valid) {
            int rowh[NC];
            u64 t[NC];
            int r = -1, ra = -1;
            bool promoted = true;                           // a root: round 0, own term only
#pragma unroll
            for (int j = 0; j < NC; j++) { rowh[j] = -1; t[j] = 0; }
            if (pa >= 0) {
                if (sa >= 0 || sb >= 0) { cp_async_wait_1(); __syncwarp(); }
                if (sa < 0 && (pa >> 5) == win) mbar_wait_acquire(&S.mbar[pa & 31], parity);
                if (sb < 0 && (pb >> 5) == win) mbar_wait_acquire(&S.mbar[pb & 31], parity);
                const DivSlot<NC> &A = S.slot[sa >= 0 ? sa : (pa & (SW_RING - 1))];
                const DivSlot<NC> &B = S.slot[sb >= 0 ? sb : (pb & (SW_RING - 1))];
                ra = A.round;
                const int rb = B.round;
                r = max(ra, rb);                                            // swirld.py:200
                int cnt = 0;
#pragma unroll
                for (int j = 0; j < NC; j++) {
                    rowh[j] = max(A.row[lane + 32 * j], B.row[lane + 32 * j]);   // swirld.py:203-205
                    const u64 ta = A.T[lane + 32 * j], tb = B.T[lane + 32 * j];
                    t[j] = (ra == r ? ta : 0ull) | (rb == r ? tb : 0ull);
                    bool ss;                                                // swirld.py:209-214
                    if (UNIT) ss = __popcll(t[j]) > (int)thr;
                    else ss = wsum(t[j], 0, S.stake) > thr;
                    cnt += __popc(__ballot_sync(0xffffffffu, ss));
                }
                promoted = (i64)cnt > thr;                                  // swirld.py:216 (quirk Q3)
            }
            int rh = r + (promoted ? 1 : 0);                                // swirld.py:217-219
            const bool wit = rh > ra;                                       // swirld.py:221 / 196-197
            if (rh >= P.Rcap) { if (lane == 0) atomicMin(&P.scal[SC_ERR], -5); rh = P.Rcap - 1; }
            const bool w_sm = rh >= wbase && rh < wbase + SW_WC;
            const int wslot = rh % SW_WC;
            if (wit && lane == 0) {                                          // swirld.py:222
                if (w_sm) S.Wc[wslot][cr] = h;
                else P.W[(size_t)rh * M + cr] = h;       // uncached round: must precede the release
                atomicMax(&S.rmaxp[parity], rh);
            }
            __syncwarp();
            DivSlot<NC> &H = S.slot[h & (SW_RING - 1)];
            u64 th[NC];
            u64 smask = 0;
#pragma unroll
            for (int j = 0; j < NC; j++) {
                const int c = lane + 32 * j;
                if (c == cr) rowh[j] = h;                                    // swirld.py:220
                int w = -1;
                if (w_sm) w = S.Wc[wslot][c];
                else if (act[j]) w = (c == cr && wit) ? h : __ldcg(P.W + (size_t)rh * M + c);
                const bool sm = w >= 0 && rowh[j] >= w;
                smask |= (u64)__ballot_sync(0xffffffffu, sm) << (32 * j);
                th[j] = (promoted ? 0ull : t[j]) | (sm ? (1ull << cr) : 0ull);
                H.row[c] = rowh[j];
                H.T[c] = th[j];
            }
            if (lane == 0) H.round = rh;
            __syncwarp();
            if (lane == 0) mbar_arrive_release(&S.mbar[warp]);
            // ---- off the critical path: stream the results to HBM
#pragma unroll
            for (int j = 0; j < NC; j++)
                if (act[j]) {
                    const size_t o = (size_t)h * M + lane + 32 * j;
                    P.row[o] = rowh[j];
                    P.T[o] = th[j];
                }
            if (lane == 0) {
                if (wit && w_sm) P.W[(size_t)rh * M + cr] = h;
                P.round[h] = rh;
                P.wit[h] = wit ? 1 : 0;
                P.SM[h] = smask;
            }
        } else if (lane == 0) {
            mbar_arrive_release(&S.mbar[warp]);        // keep every barrier's phase in step
        }
        __syncthreads();
        rmax = max(rmax, S.rmaxp[parity]);
        const int nbase = max(wbase, rmax - (SW_WC / 2 - 1));
        if (nbase != wbase) {                           // uniform: every thread sees the same rmax
            for (int i = tid; i < SW_WC * MS; i += 1024) {
                const int slot = i / MS, c = i % MS;
                const int ro = wbase + ((slot - (wbase % SW_WC) + SW_WC) % SW_WC);
                const int rn = nbase + ((slot - (nbase % SW_WC) + SW_WC) % SW_WC);
                if (rn != ro)
                    S.Wc[slot][c] = (c < M && rn < P.Rcap) ? __ldcg(P.W + (size_t)rn * M + c) : -1;
            }
            wbase = nbase;
            __syncthreads();
        }
        h += 32;
        valid = valid1; pa = pa1; pb = pb1; cr = cr1; sa = sa1; sb = sb1;
        valid1 = valid2; pa1 = pa2; pb1 = pb2; cr1 = cr2;
    }
    if (tid == 0) P.scal[SC_MAX_ROUND] = rmax;
}
