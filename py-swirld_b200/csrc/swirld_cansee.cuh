// swirld_cansee.cuh -- the can_see table (swirld.py:72, 203-205, 220) as its own,
// bandwidth-bound kernel family.
//
// row(h)[c] = max(row(p0)[c], row(p1)[c]) with the own column := h is a max-plus LINEAR
// recurrence over the DAG and every member column is independent of the others, so it
// does not have to be walked in dependency order:
//
//   A  k_cs_local     the range is cut into blocks of B consecutive events; one CTA per
//                     block, one THREAD per member column, events in index order.  A parent
//                     outside the block is a leaf (it contributes only itself, in its
//                     creator's column), so the blocks are independent.  PR(h)[c] = the latest
//                     event of member c that is an in-block ancestor (or self) of h, or a
//                     direct out-of-block parent of one.  A column whose value is in-block is
//                     final (in-block indices are larger than anything else).
//   B  k_cs_boundary  one CTA walks the blocks in order and finishes the few rows later
//                     blocks depend on: the events referenced from later blocks ("exported")
//                     and each member's last event of the block (the next block-start heads).
//   C  k_cs_local<2>  every block again, now with the final rows of its out-of-block parents (the heads
//                     of B are preloaded): the exact recurrence, blocks in parallel.
//
// Finishing a row in B: for every member m, the out-of-block ancestor that represents it is
//   e_m = Q[m] (member m's head at the start of the block) if h sees an in-block m-event,
//   else PR(h)[m] (a direct out-of-block parent, possibly older than the head), else none;
// row(h)[c] = max(PR(h)[c], max_m row(e_m)[c]) for the columns that are not in-block.  A row
// whose out-of-block columns already show the heads themselves is final as it is (Q[c] is the
// largest value column c can take below the block).
//
// (tests/test_cansee_model.py keeps an executable model of exactly this scheme against the oracle.)
#pragma once
#include "swirld_kernels.cuh"

struct CsParams {
    int M, first, n, B, nb;     // events [first, first+n) in nb blocks of B
    const int32_t *p0, *p1, *creator;
    int32_t *row;               // [cap][M]; rows < first are final
    uint8_t *exported;          // [cap] flags, zero on entry for [first, first+n)
    int32_t *exp_list;          // [cap]: block j's list lives at [first + j*B, ...)
    int32_t *exp_m;             // [cap]: for a list entry that is its member's last event of the block, the member; else -1
    int32_t *exp_cnt;           // [nb]
    int32_t *last;              // [nb][M] last event of member m inside block j, -1 none
    int32_t *Qtab;              // [nb+1][M] head of member m at the start of block j
    int32_t *carry;             // [M] heads before `first` (in/out: updated to the heads after the range)
};

#define CS_TILE 256

// ---- A / C: the in-block recurrence, one CTA per block, blockDim.x = 32*NC threads = member columns.
// PASS 1: partial rows (out-of-block parents are leaves).  PASS 2 (after k_cs_boundary): the exact
// rows -- the same walk, but an out-of-block parent contributes its FINAL row: the block-start heads
// are preloaded into the per-member cache, anything older is read from the table.
// Per member the cache holds its latest event so far and that event's value in this thread's column;
// in a gossip graph both parents of a new event are such "latest" events, so the walk runs out of
// shared memory.  Two consecutive events are processed together when the second does not have the
// first as a parent (its loads are issued before the first one's store: ~2x on the latency chain).
template <int NC, int PASS>
__global__ void __launch_bounds__(NC * 32) k_cs_local(CsParams P) {
    constexpr int MS = NC * 32;
    __shared__ int2 tv[MS][MS];              // [member][column]: (event, its cached value); one LDS.64
    __shared__ int4 meta[CS_TILE + 2];       // (p0, p1, creator | dep << 16, creator of p1)
    const int c = threadIdx.x, M = P.M;
    const int s = P.first + blockIdx.x * P.B, e = min(s + P.B, P.first + P.n);
    __shared__ int32_t qs[MS];
    if (PASS == 2) {
        qs[c] = c < M ? P.Qtab[(size_t)blockIdx.x * M + c] : -1;
        __syncthreads();
    }
#pragma unroll 8
    for (int m = 0; m < MS; m++) {           // private to this thread's column
        const int q = PASS == 2 ? qs[m] : -1;
        const bool on = q >= 0 && c < M;
        const int val = on ? P.row[(size_t)(on ? q : 0) * M + (on ? c : 0)] : -1;   // independent loads, no branches
        tv[m][c] = make_int2(q, val);
    }
    for (int t0 = s; t0 < e; t0 += CS_TILE) {
        const int tn = min(CS_TILE, e - t0);
        __syncthreads();
        for (int i = c; i < tn + 2; i += MS) {
            int4 mt = make_int4(-1, -1, 1 << 16, 0);             // padding: "depends on its predecessor"
            if (i < tn) {
                const int a = P.p0[t0 + i], b = P.p1[t0 + i];
                const int dep = (a == t0 + i - 1 || b == t0 + i - 1) ? 1 : 0;
                mt = make_int4(a, b, P.creator[t0 + i] | (dep << 16), b >= 0 ? P.creator[b] : 0);
                if (PASS == 1) {
                    if (a >= P.first && a < s) P.exported[a] = 1;   // referenced from a later block
                    if (b >= P.first && b < s) P.exported[b] = 1;
                }
            }
            meta[i] = mt;
        }
        __syncthreads();
        if (c >= M) continue;
        // One parent's contribution to column c, branch-free: the cached head of the parent's member if it
        // IS the parent, else (pass 1) the leaf value of an out-of-block parent; `need` = a table read is due.
        auto contrib = [&](int p, int pc, int2 cached, bool &need) -> int {
            const bool hit = cached.x == p;
            need = (PASS == 1 ? p >= s : p >= 0) && !hit;
            const int leaf = (PASS == 1 && p >= 0 && p < s && c == pc) ? p : -1;
            return hit ? cached.y : leaf;
        };
        int i = 0;
        int4 m0 = meta[0];
        while (i < tn) {
            const int4 m1 = meta[i + 1];
            const int h = t0 + i;
            const int cr0 = m0.z & 0xffff, cr1 = m1.z & 0xffff;
            if ((m1.z >> 16) == 0) {                                            // independent pair (i, i+1)
                const int2 a0 = tv[cr0][c], b0 = tv[m0.w][c], a1 = tv[cr1][c], b1 = tv[m1.w][c];
                const int4 m2 = meta[i + 2];
                // (all four loads in flight before the first use: keeps them out of the branches below)
                asm volatile("" :: "r"(a0.x), "r"(a0.y), "r"(b0.x), "r"(b0.y), "r"(a1.x), "r"(a1.y), "r"(b1.x), "r"(b1.y));
                bool na0, nb0, na1, nb1;
                int x0 = contrib(m0.x, cr0, a0, na0), y0 = contrib(m0.y, m0.w, b0, nb0);
                int x1 = contrib(m1.x, cr1, a1, na1), y1 = contrib(m1.y, m1.w, b1, nb1);
                if (na0 | nb0 | na1 | nb1) {                                    // rare: a parent that is not its member's latest event
                    if (na0) x0 = P.row[(size_t)m0.x * M + c];
                    if (nb0) y0 = P.row[(size_t)m0.y * M + c];
                    if (na1) x1 = P.row[(size_t)m1.x * M + c];
                    if (nb1) y1 = P.row[(size_t)m1.y * M + c];
                }
                const int v0 = c == cr0 ? h : max(x0, y0), v1 = c == cr1 ? h + 1 : max(x1, y1);
                tv[cr0][c] = make_int2(h, v0);
                tv[cr1][c] = make_int2(h + 1, v1);
                P.row[(size_t)h * M + c] = v0;
                P.row[(size_t)(h + 1) * M + c] = v1;
                m0 = m2; i += 2;
            } else {
                const int2 a0 = tv[cr0][c], b0 = tv[m0.w][c];
                asm volatile("" :: "r"(a0.x), "r"(a0.y), "r"(b0.x), "r"(b0.y));
                bool na0, nb0;
                int x0 = contrib(m0.x, cr0, a0, na0), y0 = contrib(m0.y, m0.w, b0, nb0);
                if (na0 | nb0) {
                    if (na0) x0 = P.row[(size_t)m0.x * M + c];
                    if (nb0) y0 = P.row[(size_t)m0.y * M + c];
                }
                const int v0 = c == cr0 ? h : max(x0, y0);
                tv[cr0][c] = make_int2(h, v0);
                P.row[(size_t)h * M + c] = v0;
                m0 = m1; i += 1;
            }
        }
    }
    if (PASS == 1 && c < M) P.last[(size_t)blockIdx.x * M + c] = tv[c][c].x;  // member c's last event of the block
}

// per-block work lists of k_cs_boundary: the exported events and each member's last event of the block
// (entry = event, and the member whose block-end head it is, or -1)
__global__ void k_cs_collect(CsParams P) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < P.n; j += gridDim.x * blockDim.x) {
        const int x = P.first + j, blk = j / P.B, cr = P.creator[x];
        const bool is_last = P.last[(size_t)blk * P.M + cr] == x;
        if (P.exported[x] || is_last) {
            const int pos = atomicAdd(&P.exp_cnt[blk], 1);
            P.exp_list[P.first + blk * P.B + pos] = x;
            P.exp_m[P.first + blk * P.B + pos] = is_last ? cr : -1;
            P.exported[x] = 1;
        }
    }
}

// Finish one row (one warp, lanes = columns, NC per lane).  pr: in = the partial row, out = the final row.
// Q[m] the block-start heads, S[m][c] their final rows (shared memory).
template <int NC>
__device__ __forceinline__ void cs_complete_row(const CsParams &P, int x, int lim, int lane, const int32_t *Q,
                                                const int32_t (*S)[NC * 32], int (&pr)[NC]) {
    const int M = P.M;
    int q[NC];
    bool inb[NC];
    bool all_in = true, fast = true;
#pragma unroll
    for (int j = 0; j < NC; j++) {
        const int c = lane + 32 * j;
        if (c >= M) pr[j] = 0x7fffffff;                             // padded columns count as in-block
        q[j] = c < M ? Q[c] : -1;
        inb[j] = pr[j] >= lim;
        all_in &= inb[j];
        // every out-of-block column already shows its member's head: nothing older can add to it
        // (the head of c is the largest value column c can take outside the block)
        fast &= inb[j] || pr[j] == q[j];
    }
    if (__all_sync(0xffffffffu, all_in | fast)) return;
    int acc[NC];
#pragma unroll
    for (int j = 0; j < NC; j++) acc[j] = pr[j];
#pragma unroll
    for (int jj = 0; jj < NC; jj++) {
        const int cj = lane + 32 * jj;
        // members whose chain below the block is entered at the head (h sees an in-block event of the
        // member, or has the head itself as a parent) / at an older event (a direct out-of-block parent)
        unsigned hm = __ballot_sync(0xffffffffu, cj < M && q[jj] >= 0 && (inb[jj] || pr[jj] == q[jj]));
        unsigned om = __ballot_sync(0xffffffffu, cj < M && !inb[jj] && pr[jj] >= 0 && pr[jj] != q[jj]);
        while (hm) {
            const int m = jj * 32 + __ffs(hm) - 1;
            hm &= hm - 1;
#pragma unroll
            for (int j = 0; j < NC; j++) acc[j] = max(acc[j], S[m][lane + 32 * j]);
        }
        while (om) {
            const int l = __ffs(om) - 1;
            om &= om - 1;
            const int ev = __shfl_sync(0xffffffffu, pr[jj], l);
#pragma unroll
            for (int j = 0; j < NC; j++) {
                const int c = lane + 32 * j;
                if (c < M) acc[j] = max(acc[j], __ldcg(P.row + (size_t)ev * M + c));
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NC; j++) {
        const int c = lane + 32 * j;
        if (c < M && !inb[j] && acc[j] != pr[j]) P.row[(size_t)x * M + c] = acc[j];
        if (!inb[j]) pr[j] = acc[j];
    }
}

// ---- B: block by block, the rows later blocks depend on.  One CTA; per block: finish the listed rows
// (one warp per row, the partial rows of a warp's batch are fetched together), then install the block's
// last events as the new heads.  The next block's list is fetched while this one is processed.
#define CS_LIST 512
template <int NC>
__global__ void __launch_bounds__(1024, 1) k_cs_boundary(CsParams P) {
    constexpr int MS = NC * 32;
    __shared__ int32_t Q[MS], newq[MS];
    __shared__ int32_t S[MS][MS], Snew[MS][MS];
    __shared__ int32_t lst[3][CS_LIST], lstm[3][CS_LIST];       // lists of blocks blk, blk+1, blk+2
    __shared__ int cnt_s[3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, M = P.M;
    if (tid < MS) { Q[tid] = tid < M ? P.carry[tid] : -1; newq[tid] = -1; }
    if (tid < 2) cnt_s[tid] = tid < P.nb ? P.exp_cnt[tid] : 0;
    __syncthreads();
    for (int i = tid; i < MS * MS; i += 1024) {
        const int m = i / MS, c = i % MS;
        S[m][c] = (m < M && c < M && Q[m] >= 0) ? P.row[(size_t)Q[m] * M + c] : -1;
    }
    for (int b = 0; b < 2 && b < P.nb; b++)
        if (tid < min(cnt_s[b], CS_LIST)) {
            lst[b][tid] = P.exp_list[P.first + b * P.B + tid]; lstm[b][tid] = P.exp_m[P.first + b * P.B + tid];
        }
    __syncthreads();
    // a warp's first four rows of a block (entries warp, warp+32, ...) are fetched one block ahead
    int xc[4], mc[4], prc[4][NC];
    auto fetch = [&](int blk, int b, int (&x)[4], int (&mm)[4], int (&pr)[4][NC]) {
        const int cnt = blk < P.nb ? cnt_s[b] : 0;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = warp + 32 * u;
            x[u] = -1; mm[u] = -1;
            if (i < cnt) {
                x[u] = i < CS_LIST ? lst[b][i] : P.exp_list[P.first + blk * P.B + i];
                mm[u] = i < CS_LIST ? lstm[b][i] : P.exp_m[P.first + blk * P.B + i];
            }
#pragma unroll
            for (int j = 0; j < NC; j++) {
                const int c = lane + 32 * j;
                pr[u][j] = (x[u] >= 0 && c < M) ? P.row[(size_t)x[u] * M + c] : 0;
            }
        }
    };
    fetch(0, 0, xc, mc, prc);
    for (int blk = 0; blk < P.nb; blk++) {
        const int b0 = blk % 3, b1 = (blk + 1) % 3, b2 = (blk + 2) % 3;
        const int lim = P.first + blk * P.B, cnt = cnt_s[b0];
        if (tid < M) P.Qtab[(size_t)blk * M + tid] = Q[tid];             // heads for pass 2
        // loads for later blocks first: the list of blk+2, the first rows of blk+1 (partial rows of
        // pass 1: nothing in this loop writes them before their own turn)
        int ncnt = 0, nx = -1, nm = -1;
        if (blk + 2 < P.nb) {
            ncnt = P.exp_cnt[blk + 2];
            if (tid < min(ncnt, CS_LIST)) { nx = P.exp_list[lim + 2 * P.B + tid]; nm = P.exp_m[lim + 2 * P.B + tid]; }
        }
        int xn[4], mn[4], prn[4][NC];
        fetch(blk + 1, b1, xn, mn, prn);
        auto finish = [&](int x, int mm, int (&pr)[NC]) {
            cs_complete_row<NC>(P, x, lim, lane, Q, S, pr);
            if (mm >= 0) {                                               // the member's head after this block
#pragma unroll
                for (int j = 0; j < NC; j++) if (lane + 32 * j < M) Snew[mm][lane + 32 * j] = pr[j];
                if (lane == 0) newq[mm] = x;
            }
        };
#pragma unroll
        for (int u = 0; u < 4; u++) if (xc[u] >= 0) finish(xc[u], mc[u], prc[u]);   // (uniform per warp)
        for (int i = warp + 128; i < cnt; i += 32) {                     // long lists: the rest on demand
            const int x = i < CS_LIST ? lst[b0][i] : P.exp_list[lim + i];
            const int mm = i < CS_LIST ? lstm[b0][i] : P.exp_m[lim + i];
            int pr[NC];
#pragma unroll
            for (int j = 0; j < NC; j++) pr[j] = lane + 32 * j < M ? P.row[(size_t)x * M + lane + 32 * j] : 0;
            finish(x, mm, pr);
        }
        __syncthreads();
        for (int i = tid; i < MS * MS; i += 1024) {
            const int m = i / MS, c = i % MS;
            if (newq[m] >= 0) S[m][c] = Snew[m][c];
        }
        if (tid < min(ncnt, CS_LIST)) { lst[b2][tid] = nx; lstm[b2][tid] = nm; }
        if (tid == 0) cnt_s[b2] = ncnt;
        __syncthreads();
        if (tid < MS && newq[tid] >= 0) { Q[tid] = newq[tid]; newq[tid] = -1; }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            xc[u] = xn[u]; mc[u] = mn[u];
#pragma unroll
            for (int j = 0; j < NC; j++) prc[u][j] = prn[u][j];
        }
        __syncthreads();
    }
    if (tid < M) { P.carry[tid] = Q[tid]; P.Qtab[(size_t)P.nb * M + tid] = Q[tid]; }
}
