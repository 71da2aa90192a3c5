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
//   C  k_cs_fix       every other event with a column that is not in-block is finished in
//                     parallel.
//
// Finishing a row: for every member m, the out-of-block ancestor that represents it is
//   e_m = Q[m] (member m's head at the start of the block) if h sees an in-block m-event,
//   else PR(h)[m] (a direct out-of-block parent, possibly older than the head), else none;
// row(h)[c] = max(PR(h)[c], max_m row(e_m)[c]) for the columns that are not in-block.  When
// every e_m is the head Q[m] (the usual case once the block has mixed) that inner max is the
// per-block constant CM[c] = max_m row(Q[m])[c]: one compare per column, no gathers.
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
    int32_t *exp_cnt;           // [nb]
    int32_t *last;              // [nb][M] last event of member m inside block j, -1 none
    int32_t *Qtab;              // [nb+1][M] head of member m at the start of block j
    int32_t *CM;                // [nb][M]  column-wise max of the rows of Qtab[j][*]
    int32_t *carry;             // [M] heads before `first` (in/out: updated to the heads after the range)
};

#define CS_TILE 256

// ---- A: per-block partial rows.  blockDim.x = 32*NC threads = member columns.
template <int NC>
__global__ void __launch_bounds__(NC * 32) k_cs_local(CsParams P) {
    constexpr int MS = NC * 32;
    __shared__ int2 tv[MS][MS];              // [member][column]: (event, its cached value); one LDS.64
    __shared__ int32_t sp0[CS_TILE], sp1[CS_TILE], scr[CS_TILE], scb[CS_TILE];
    const int c = threadIdx.x, M = P.M;
    const int s = P.first + blockIdx.x * P.B, e = min(s + P.B, P.first + P.n);
    for (int m = 0; m < MS; m++) tv[m][c] = make_int2(-1, -1);   // private to this thread's column
    for (int t0 = s; t0 < e; t0 += CS_TILE) {
        const int tn = min(CS_TILE, e - t0);
        __syncthreads();
        for (int i = c; i < tn; i += MS) {
            const int a = P.p0[t0 + i], b = P.p1[t0 + i];
            sp0[i] = a; sp1[i] = b; scr[i] = P.creator[t0 + i];
            scb[i] = b >= 0 ? P.creator[b] : 0;
            if (a >= P.first && a < s) P.exported[a] = 1;   // referenced from a later block
            if (b >= P.first && b < s) P.exported[b] = 1;
        }
        __syncthreads();
        if (c >= M) continue;
        for (int i = 0; i < tn; i++) {
            const int h = t0 + i, pa = sp0[i], pb = sp1[i], cr = scr[i], cb = scb[i];
            int v = -1;
            if (pa >= 0) {
                const int2 ca = tv[cr][c], cbv = tv[cb][c];   // both cached heads at once
                int a, b;
                if (pa >= s) a = ca.x == pa ? ca.y : P.row[(size_t)pa * M + c];
                else a = c == cr ? pa : -1;                  // out-of-block parent: a leaf
                if (pb >= s) b = cbv.x == pb ? cbv.y : P.row[(size_t)pb * M + c];
                else b = c == cb ? pb : -1;
                v = max(a, b);
            }
            if (c == cr) v = h;
            tv[cr][c] = make_int2(h, v);
            P.row[(size_t)h * M + c] = v;
        }
    }
    if (c < M) P.last[(size_t)blockIdx.x * M + c] = tv[c][c].x;  // member c's last event of the block
}

// per-block lists of the exported events
__global__ void k_cs_collect(CsParams P) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < P.n; j += gridDim.x * blockDim.x) {
        const int x = P.first + j;
        if (P.exported[x]) {
            const int blk = j / P.B;
            const int pos = atomicAdd(&P.exp_cnt[blk], 1);
            P.exp_list[P.first + blk * P.B + pos] = x;
        }
    }
}

// Finish one row (one warp, lanes = columns, NC per lane).  Q/S/CMs live in shared memory:
// Q[m] the block-start heads, S[m][c] their final rows, CMs[c] the column maxima (NULL: no fast path).
template <int NC>
__device__ __forceinline__ void cs_complete_row(const CsParams &P, int x, int lim, int lane,
                                                const int32_t *Q, const int32_t (*S)[NC * 32], const int32_t *CMs) {
    const int M = P.M;
    int pr[NC], q[NC];
    bool inb[NC];
    bool all_in = true, fast = CMs != nullptr;
#pragma unroll
    for (int j = 0; j < NC; j++) {
        const int c = lane + 32 * j;
        pr[j] = c < M ? P.row[(size_t)x * M + c] : 0x7fffffff;     // padded columns count as in-block
        q[j] = c < M ? Q[c] : -1;
        inb[j] = pr[j] >= lim;
        all_in &= inb[j];
        // fast path: every member with a head is represented by exactly that head
        fast &= inb[j] ? true : (pr[j] == q[j]);                    // covers "no head, nothing seen" (-1 == -1)
    }
    if (__all_sync(0xffffffffu, all_in)) return;                    // nothing outside the block: already final
    int acc[NC];
    if (__all_sync(0xffffffffu, fast)) {
#pragma unroll
        for (int j = 0; j < NC; j++) acc[j] = inb[j] ? pr[j] : max(pr[j], CMs[lane + 32 * j]);
    } else {
#pragma unroll
        for (int j = 0; j < NC; j++) acc[j] = pr[j];
#pragma unroll
        for (int jj = 0; jj < NC; jj++) {
            for (int l = 0; l < 32; l++) {
                const int m = jj * 32 + l;
                if (m >= M) break;
                int ev = __shfl_sync(0xffffffffu, pr[jj], l);
                const int qm = __shfl_sync(0xffffffffu, q[jj], l);
                if (ev >= lim) ev = qm;                             // sees an in-block m-event: its chain reaches the head
                if (ev < 0) continue;
#pragma unroll
                for (int j = 0; j < NC; j++) {
                    const int c = lane + 32 * j;
                    if (c < M) acc[j] = max(acc[j], ev == qm ? S[m][c] : __ldcg(P.row + (size_t)ev * M + c));
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NC; j++) {
        const int c = lane + 32 * j;
        if (c < M && !inb[j] && acc[j] != pr[j]) P.row[(size_t)x * M + c] = acc[j];
    }
}

// ---- B: block by block, the rows later blocks depend on
template <int NC>
__global__ void __launch_bounds__(1024, 1) k_cs_boundary(CsParams P) {
    constexpr int MS = NC * 32;
    __shared__ int32_t Q[MS], CMs[MS];
    __shared__ int32_t S[MS][MS];
    __shared__ int cnt_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, M = P.M;
    if (tid < MS) Q[tid] = tid < M ? P.carry[tid] : -1;
    __syncthreads();
    for (int i = tid; i < MS * MS; i += 1024) {
        const int m = i / MS, c = i % MS;
        S[m][c] = (m < M && c < M && Q[m] >= 0) ? P.row[(size_t)Q[m] * M + c] : -1;
    }
    __syncthreads();
    for (int blk = 0; blk < P.nb; blk++) {
        const int lim = P.first + blk * P.B;
        int32_t *list = P.exp_list + lim;
        if (tid < MS) {                                      // tables for k_cs_fix
            int cm = -1;
            for (int m = 0; m < M; m++) cm = max(cm, S[m][tid]);
            CMs[tid] = cm;
            if (tid < M) { P.CM[(size_t)blk * M + tid] = cm; P.Qtab[(size_t)blk * M + tid] = Q[tid]; }
        }
        if (tid == 0) cnt_s = P.exp_cnt[blk];
        __syncthreads();
        if (tid < M) {                                       // the block's last event of every member
            const int x = P.last[(size_t)blk * M + tid];
            if (x >= 0 && !P.exported[x]) { P.exported[x] = 1; list[atomicAdd(&cnt_s, 1)] = x; }
        }
        __syncthreads();
        const int cnt = cnt_s;
        for (int i = warp; i < cnt; i += 32) cs_complete_row<NC>(P, list[i], lim, lane, Q, S, CMs);
        __syncthreads();
        for (int i = tid; i < MS * MS; i += 1024) {          // heads for the next block
            const int m = i / MS, c = i % MS;
            const int x = m < M ? P.last[(size_t)blk * M + m] : -1;
            if (x >= 0 && c < M) S[m][c] = P.row[(size_t)x * M + c];
        }
        __syncthreads();
        if (tid < M) { const int x = P.last[(size_t)blk * M + tid]; if (x >= 0) Q[tid] = x; }
        __syncthreads();
    }
    if (tid < M) { P.carry[tid] = Q[tid]; P.Qtab[(size_t)P.nb * M + tid] = Q[tid]; }
}

// ---- C: everything else, in parallel.  grid = (tiles per block, nb), 8 warps, 512 events per CTA.
#define CS_FIX_EVENTS 512
template <int NC>
__global__ void __launch_bounds__(256) k_cs_fix(CsParams P) {
    constexpr int MS = NC * 32;
    __shared__ int32_t Q[MS], CMs[MS];
    __shared__ int32_t S[MS][MS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, M = P.M;
    const int blk = blockIdx.y;
    const int lim = P.first + blk * P.B, bend = min(lim + P.B, P.first + P.n);
    const int t0 = lim + blockIdx.x * CS_FIX_EVENTS, t1 = min(t0 + CS_FIX_EVENTS, bend);
    if (t0 >= bend) return;
    if (tid < MS) { Q[tid] = tid < M ? P.Qtab[(size_t)blk * M + tid] : -1; CMs[tid] = tid < M ? P.CM[(size_t)blk * M + tid] : -1; }
    __syncthreads();
    for (int i = tid; i < MS * MS; i += 256) {
        const int m = i / MS, c = i % MS;
        S[m][c] = (m < M && c < M && Q[m] >= 0) ? P.row[(size_t)Q[m] * M + c] : -1;
    }
    __syncthreads();
    for (int x = t0 + warp; x < t1; x += 8) {
        if (P.exported[x]) continue;                         // finished by k_cs_boundary
        cs_complete_row<NC>(P, x, lim, lane, Q, S, CMs);
    }
}
