// swirld_cansee.cuh -- the can_see table (swirld.py:72, 203-205, 220) for any member count.
//
// row(h)[c] = max(row(p0)[c], row(p1)[c]) with the own column := h is a max-plus LINEAR recurrence
// over the DAG and every member column is independent of the others.  So the table is cut twice:
// into BLOCKS of B consecutive events and into COLUMN TILES of 32 members; one warp (= one CTA) owns
// one (block, tile) and walks the block's events in index order with one thread per column.  The only
// state of the walk is, per member, the value of that member's latest event in this thread's column
// (`val[M][32]` ints of shared memory): in a gossip graph both parents of a new event are such latest
// events, so the walk never leaves shared memory and never waits for another thread.
//
//   prep    per event: packed (p0, p1, creator, creator of p1, flags) for the walks; which rows anything
//           else will read (`wr`), which rows a LATER block reads (`xb`), each member's last event per
//           block; groups of four consecutive events with pairwise disjoint members are flagged: their
//           eight cache reads are issued together (the walk is a latency chain, this is its pipelining).
//   pass 1  parents outside the block are leaves (they contribute only themselves, in their creator's
//           column).  Nothing but the `wr` rows and each member's last row of the block is written.
//   heads   Q[j][m] = member m's latest event below block j (a scan of `last` over the blocks).
//   check   every listed row (xb, or a member's last row), all blocks AT ONCE: a row whose out-of-block
//           columns all show the member's head Q[j][c] itself is already final -- the head is the largest
//           value column c can take below the block.  With B >= 16 M that is all but ~0.1 % of the listed
//           rows of a gossip graph; the others go to a list and
//   slow    are finished, in dependency waves, from the final rows of the events they enter the lower
//           blocks through (one CTA, a handful of rows or an empty launch).
//   pass 2  the exact rows: the same walk with the cache preloaded with the final rows of the block-start
//           heads; a stale other-parent (not its member's latest event: `stale`, from sw_append) is read
//           from the table.  This is the only pass that writes the table.
//
// tests/test_cansee2_model.py is the executable model of exactly this scheme against the oracle.
#pragma once
#include "swirld_kernels.cuh"

struct CsParams {
    int M, first, n, B, nb, first_al;   // events [first, first+n); block 0 = [first, first_al+B), block j = [first_al+jB, ..)
    int CT;                     // columns per tile: 32, or 16 / 8 when the per-member cache would leave a SM too few CTAs
    int SV;                     // prefetch slots in shared memory: CS_SV when the range holds stale other-parents, else 0
    // several GPUs (sw_peer_connect): the column tiles are split over the ranks and every row segment a walk produces
    // is stored straight into EVERY rank's table over NVLink (the all-gather of the can_see table, fused into the walk)
    int tile_lo, npeer;         // first tile of this rank; 0 = single GPU
    int32_t *prow[8];           // peer p's table (own: row)
    const int32_t *p0, *p1, *creator;
    const uint8_t *stale;       // [cap] from sw_append: the other-parent is not its member's latest event
    int32_t *row;               // [cap][M]; rows < first are final
    int4 *meta;                 // [cap] (p0, p1, creator | stale << 16 | clean4 << 17, creator of p1)
    uint8_t *wr, *xb;           // [cap] zero on entry for the range
    int32_t *last;              // [nb][M] last event of member m inside block j, -1 none (filled with -1 on entry)
    int32_t *Qtab;              // [nb+1][M] head of member m at the start of block j
    int32_t *carry;             // [M] heads before `first` (in/out)
    int32_t *slow_list;         // [cap] the listed rows that failed the check (any order)
    int32_t *slow_cnt;          // [4]: slow rows, slow rows still pending, rows on xlist; zero on entry
    int32_t *xlist;             // [cap] the rows a later block reads (each once)
    int32_t *slow_blk;          // [cap] the failed rows again, per block: block j's at [start(j), ..)
    int32_t *blk_cnt;           // [nb] their counts, zero on entry
    uint8_t *sflag;             // [cap] 0 final, 1 pending slow row, 2 + w finished in wave w of k_cs_slow_*
};

#define CS_TILE 128
#define CS_SV 64                // prefetch slots for the stale other-parents of a tile (the rest is read in the walk)
#define CS_CT 32                // threads per CTA (one warp); columns per tile = P.CT <= 32

// (the last block takes whatever is left: a short tail block would fail the finality check often, so the host
//  folds a tail shorter than 3B/4 into the block before it)
__device__ __forceinline__ int cs_block_of(const CsParams &P, int h) {
    return h < P.first_al + P.B ? 0 : min((h - P.first_al) / P.B, P.nb - 1);
}
__device__ __forceinline__ int cs_start(const CsParams &P, int j) { return j == 0 ? P.first : P.first_al + j * P.B; }
__device__ __forceinline__ int cs_end(const CsParams &P, int j) { return j == P.nb - 1 ? P.first + P.n : P.first_al + (j + 1) * P.B; }

__global__ void __launch_bounds__(256) k_cs_prep(CsParams P) {
    const int end = P.first + P.n;
    for (int h = P.first + blockIdx.x * blockDim.x + threadIdx.x; h < end; h += gridDim.x * blockDim.x) {
        const int a = P.p0[h], b = P.p1[h], c = P.creator[h];
        const int cb = b >= 0 ? P.creator[b] : 0;
        const int st = (b >= 0 && P.stale[h]) ? 1 : 0;
        const int bh = cs_block_of(P, h);
        if (P.nb > 1) {
            auto mark = [&](int p) {                   // referenced from a later block: listed for the finality check, once
                P.wr[p] = 1;
                const unsigned sh = 8u * (p & 3);
                const unsigned old = atomicOr(reinterpret_cast<unsigned *>(P.xb + (p & ~3)), 1u << sh);
                if (!((old >> sh) & 1u)) P.xlist[atomicAdd(&P.slow_cnt[2], 1)] = p;
            };
            if (a >= P.first && cs_block_of(P, a) != bh) mark(a);
            if (b >= P.first) {
                if (cs_block_of(P, b) != bh) mark(b);
                else if (st) P.wr[b] = 1;
            }
            atomicMax(&P.last[(size_t)bh * P.M + c], h);
        } else atomicMax(&P.last[c], h);
        int clean = 0;
        if ((h & 3) == 0 && h + 3 < cs_end(P, bh)) {       // h, h+1, h+2, h+3: no member written by one is read by a later one
            // (a stale other-parent is read from the table, not from the cache: it takes part when it lies before the
            //  group's tile, where the walk prefetches it)
            const int s0 = cs_start(P, bh), t0 = s0 + ((h - s0) / CS_TILE) * CS_TILE;
            int cr[4], co[4], sti[4];
            cr[0] = c; co[0] = b >= 0 ? cb : c; sti[0] = st;
            bool ok = !st || b < t0;
#pragma unroll
            for (int i = 1; i < 4; i++) {
                const int bi = P.p1[h + i];
                cr[i] = P.creator[h + i];
                co[i] = bi >= 0 ? P.creator[bi] : cr[i];
                sti[i] = (bi >= 0 && P.stale[h + i]) ? 1 : 0;
                ok &= !sti[i] || bi < t0;
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = i + 1; j < 4; j++) ok &= cr[i] != cr[j] && (sti[j] || cr[i] != co[j]);
            clean = ok ? 1 : 0;
        }
        P.meta[h] = make_int4(a, b, c | (st << 16) | (clean << 17), cb);
    }
}

// a row segment goes to this rank's table, or to every rank's (P2P stores over NVLink)
template <bool MULTI>
__device__ __forceinline__ void cs_store(const CsParams &P, size_t idx, int v) {
    if (MULTI) { for (int p = 0; p < P.npeer; p++) P.prow[p][idx] = v; }
    else P.row[idx] = v;
}

// One (block, column tile) per warp.  PASS 1: leaves outside the block, sparse writes.  PASS 2: exact.
// STALE: the range holds stale other-parents (sw_append counts them); without, the prefetch machinery compiles away
template <int PASS, bool STALE, bool MULTI>
__global__ void __launch_bounds__(CS_CT) k_cs_pass(CsParams P) {
    extern __shared__ int cs_smem[];
    const int CT = P.CT;
    int4 *meta = reinterpret_cast<int4 *>(cs_smem);                                // [CS_TILE]
    uint8_t *wrt = reinterpret_cast<uint8_t *>(meta + CS_TILE);                    // [CS_TILE]
    uint8_t *slist = wrt + CS_TILE;                                                // 2 * CS_TILE bytes: the prefetched parents' indices
    int *svb = cs_smem + CS_TILE * 4 + 3 * CS_TILE / 4;                            // [CS_SV][CT] prefetched rows of stale other-parents
    int *valb = svb + P.SV * CT;                                                   // [M][CT]
    const int tl = threadIdx.x, M = P.M, blk = blockIdx.x;
    const int lane = tl & (CT - 1);                                                // lanes >= CT shadow lane % CT (their stores are off)
    const int c = (blockIdx.y + (MULTI ? P.tile_lo : 0)) * CT + lane;
    const bool own = tl < CT;                                                      // (a shadow lane never stores)
    const bool col = own && c < M;
    if (PASS == 2 && blockIdx.x == 0 && blockIdx.y == 0 && !MULTI)           // the next launch's carry heads (several ranks: k_cs_carry)
        for (int m = tl; m < M; m += CS_CT) P.carry[m] = P.Qtab[(size_t)P.nb * M + m];
#define val(m) (valb + (size_t)(m) * CT)
    const int s = cs_start(P, blk), e = cs_end(P, blk);
    int32_t *rowc = P.row + (col ? c : 0);
    const size_t cc = col ? c : 0;                                                 // (stores: cs_store<MULTI>(P, h * M + cc, v))
    if (PASS == 1) {
        for (int m = 0; m < M; m++) if (own) val(m)[lane] = -1;
    } else {
        const int32_t *Q = P.Qtab + (size_t)blk * M;
#pragma unroll 8
        for (int m = 0; m < M; m++) {
            const int q = Q[m];
            const int v0 = (q >= 0 && col) ? rowc[(size_t)q * M] : -1;
            if (own) val(m)[lane] = v0;
        }
    }
    for (int t0 = s; t0 < e; t0 += CS_TILE) {
        const int tn = min(CS_TILE, e - t0);
        __syncwarp();
        for (int i = tl; i < tn; i += CS_CT) {
            meta[i] = P.meta[t0 + i];
            if (PASS == 1) wrt[i] = P.wr[t0 + i];
        }
        __syncwarp();
        // a stale other-parent (not its member's latest event) is read from the table; when it lies before this tile its
        // value is there already: fetch all of the tile's now, eight loads in flight, instead of one memory round trip
        // per event in the walk
        if (STALE) {
            // prefetch slot k <- the stale other-parent of tile position i: its index goes to sb[k] and the tile's copy of
            // the event is re-coded (p1 := -(k + 2)) so that the walk finds the slot without another lookup
            int *sb = reinterpret_cast<int *>(slist);           // [CS_SV] (the two byte arrays' space: 2 * CS_TILE bytes)
            int ns = 0;
            for (int i0 = 0; i0 < tn; i0 += CS_CT) {
                const int i = i0 + tl;
                bool pf = false;
                int b = -1;
                if (i < tn) { const int4 mi = meta[i]; b = mi.y; pf = ((mi.z >> 16) & 1) && b < t0 && (PASS == 2 || b >= s); }
                const unsigned bal = __ballot_sync(0xffffffffu, pf);
                const int slot = ns + __popc(bal & ((1u << tl) - 1));
                if (pf && slot < CS_SV) { sb[slot] = b; meta[i].y = -(slot + 2); }
                ns += __popc(bal);
            }
            ns = min(ns, CS_SV);
            __syncwarp();
            // every lane copies its own column of every prefetched row straight into shared memory: all in flight at once
            if (own) {
#pragma unroll 4
                for (int k = 0; k < ns; k++) {
                    int *dst = svb + k * CT + lane;
                    if (col) {
                        const int32_t *src = rowc + (size_t)sb[k] * M;
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
                    } else *dst = -1;
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
        }
        int i = 0;
        while (i < tn) {
            const int4 m0 = meta[i];
            const int h = t0 + i;
            bool fast = ((m0.z >> 17) & 1) && i + 4 <= tn;
            if (STALE && fast) {                           // (a stale member of the group needs its prefetch slot)
                const unsigned stm = ((m0.z >> 16) & 1) | (((meta[i + 1].z >> 16) & 1) << 1) | (((meta[i + 2].z >> 16) & 1) << 2) | (((meta[i + 3].z >> 16) & 1) << 3);
                if (stm) {
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (((stm >> u) & 1) && meta[i + u].y > -2 && !(PASS == 1 && meta[i + u].y < s)) fast = false;
                }
            }
            if (fast) {
                // four events with pairwise disjoint members: all cache reads first
                const int4 m1 = meta[i + 1], m2 = meta[i + 2], m3 = meta[i + 3];
                const int c0 = m0.z & 0xffff, c1 = m1.z & 0xffff, c2 = m2.z & 0xffff, c3 = m3.z & 0xffff;
                int x0, x1, x2, x3, y0, y1, y2, y3;
                auto other = [&](const int4 &m, int k) -> int {      // the other-parent's contribution in this column
                    if (STALE && ((m.z >> 16) & 1)) {
                        if (m.y <= -2) return svb[(-m.y - 2) * CT + lane];
                        return (c == m.w) ? m.y : -1;                // (pass 1, a stale parent below the block: a leaf)
                    }
                    if (PASS == 1) return m.y >= s ? val(m.w)[lane] : ((m.y >= 0 && c == m.w) ? m.y : -1);
                    return m.y >= 0 ? val(m.w)[lane] : -1;
                };
                if (PASS == 1) {
                    x0 = m0.x >= s ? val(c0)[lane] : -1; x1 = m1.x >= s ? val(c1)[lane] : -1;
                    x2 = m2.x >= s ? val(c2)[lane] : -1; x3 = m3.x >= s ? val(c3)[lane] : -1;
                } else {
                    x0 = val(c0)[lane]; x1 = val(c1)[lane]; x2 = val(c2)[lane]; x3 = val(c3)[lane];
                }
                y0 = other(m0, 0); y1 = other(m1, 1); y2 = other(m2, 2); y3 = other(m3, 3);
                const int v0 = c == c0 ? h : max(x0, y0), v1 = c == c1 ? h + 1 : max(x1, y1);
                const int v2 = c == c2 ? h + 2 : max(x2, y2), v3 = c == c3 ? h + 3 : max(x3, y3);
                if (own) { val(c0)[lane] = v0; val(c1)[lane] = v1; val(c2)[lane] = v2; val(c3)[lane] = v3; }
                if (PASS == 2) {
                    if (col) {
                        cs_store<MULTI>(P, (size_t)h * M + cc, v0); cs_store<MULTI>(P, (size_t)(h + 1) * M + cc, v1);
                        cs_store<MULTI>(P, (size_t)(h + 2) * M + cc, v2); cs_store<MULTI>(P, (size_t)(h + 3) * M + cc, v3);
                    }
                } else if (col) {
                    if (wrt[i]) cs_store<MULTI>(P, (size_t)h * M + cc, v0);
                    if (wrt[i + 1]) cs_store<MULTI>(P, (size_t)(h + 1) * M + cc, v1);
                    if (wrt[i + 2]) cs_store<MULTI>(P, (size_t)(h + 2) * M + cc, v2);
                    if (wrt[i + 3]) cs_store<MULTI>(P, (size_t)(h + 3) * M + cc, v3);
                }
                i += 4;
            } else {
                const int a = m0.x, b = m0.y, cr = m0.z & 0xffff, st = (m0.z >> 16) & 1, cb = m0.w;
                int x, y;
                if (PASS == 1) {
                    x = a >= s ? val(cr)[lane] : -1;
                    if (STALE && st && b <= -2) y = svb[(-b - 2) * CT + lane];
                    else if (b >= s) y = (STALE && st) ? (col ? rowc[(size_t)b * M] : -1) : val(cb)[lane];
                    else y = (b >= 0 && c == cb) ? b : -1;
                } else {
                    x = val(cr)[lane];
                    if (STALE && st && b <= -2) y = svb[(-b - 2) * CT + lane];
                    else y = b < 0 ? -1 : ((STALE && st) ? (col ? rowc[(size_t)b * M] : -1) : val(cb)[lane]);
                }
                const int v = c == cr ? h : max(x, y);
                if (own) val(cr)[lane] = v;
                if (col && (PASS == 2 || wrt[i])) cs_store<MULTI>(P, (size_t)h * M + cc, v);
                i += 1;
            }
        }
    }
    if (PASS == 1 && col) {                          // each member's last row of the block, from the cache
        const int32_t *__restrict__ L = P.last + (size_t)blk * M;
        for (int m0 = 0; m0 < M; m0 += 8) {          // (eight loads in flight, then the stores: they may alias for the compiler)
            int l[8];
#pragma unroll
            for (int u = 0; u < 8; u++) l[u] = m0 + u < M ? __ldg(L + m0 + u) : -1;
#pragma unroll
            for (int u = 0; u < 8; u++) if (l[u] >= 0) cs_store<MULTI>(P, (size_t)l[u] * M + cc, val(m0 + u)[lane]);
        }
    }
#undef val
}

// heads at the start of every block: Q[j][m] = member m's latest event below block j = its last event of the nearest
// earlier block that has one (with B >= 16 M that is the block before, almost always), else the carry head.  Row nb is
// the carry of the next launch (installed by k_cs_pass<2>).
__global__ void k_cs_heads(CsParams P) {
    const size_t tot = (size_t)(P.nb + 1) * P.M;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(i / P.M), m = (int)(i % P.M);
        int q = -1;
        int jj = j - 1;
        for (; jj >= 0; jj--) { q = __ldg(P.last + (size_t)jj * P.M + m); if (q >= 0) break; }
        if (jj < 0) q = P.carry[m];
        P.Qtab[i] = q;
    }
}

// finality check of the listed rows of all blocks at once (see the header): the items are every member's last event
// of every block, and the rows a later block reads (xlist; one of those that is also a last event is taken once)
__global__ void __launch_bounds__(256) k_cs_check(CsParams P) {
    const int lane = threadIdx.x & 31, M = P.M;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int nlast = P.nb * M, total = nlast + P.slow_cnt[2];
    for (int i = wid; i < total; i += nw) {
        int x;
        if (i < nlast) x = P.last[i];
        else {
            x = P.xlist[i - nlast];
            if (P.last[(size_t)cs_block_of(P, x) * M + P.creator[x]] == x) x = -1;
        }
        if (x < 0) continue;
        const int bx = cs_block_of(P, x), lim = cs_start(P, bx);
        const int32_t *Q = P.Qtab + (size_t)bx * M;
        bool ok = true;
        for (int c = lane; c < M; c += 32) {
            const int pr = P.row[(size_t)x * M + c];
            ok &= pr >= lim || pr == Q[c];
        }
        if (!__all_sync(0xffffffffu, ok) && lane == 0) {
            P.slow_list[atomicAdd(&P.slow_cnt[0], 1)] = x;
            atomicAdd(&P.slow_cnt[1], 1);
            P.slow_blk[lim + atomicAdd(&P.blk_cnt[bx], 1)] = x;
            P.sflag[x] = 1;
        }
    }
}

// The rows that failed the check: row(x)[c] = max(partial, max over members m of the FINAL row of the event through
// which x enters m's chain below the block: the head Q[m] if x sees an in-block event of m (or the head itself),
// else the direct out-of-block parent it shows) -- for the columns that can still grow (neither in-block nor the head
// itself).  Those entry events are listed rows of earlier blocks; their column in question is final unless the entry is
// itself a row in flux AND that column is one of its open ones -- so the rows are finished in dependency WAVES, column
// by column: an open column whose entries are all settled is written now; a row with no open column left is stamped
// 2 + wave; the others wait for the next wave (the pending rows of the lowest block always finish).  Waves 1 and 2 are
// grid-wide launches, whatever is left (normally nothing) is finished by one CTA.
#define CS_SLOW_WARPS 16
#define CS_COLPAR 8             // a row with more open columns than this is finished with one lane per column

// number of open columns of row x (neither in-block nor the head itself)
__device__ __forceinline__ int cs_open_count(const CsParams &P, int x, int lim, const int32_t *Q, int lane) {
    int n = 0;
    for (int c0 = 0; c0 < P.M; c0 += 32) {
        const int c = c0 + lane;
        bool bad = false;
        if (c < P.M) { const int pr = P.row[(size_t)x * P.M + c]; bad = pr < lim && pr != Q[c]; }
        n += __popc(__ballot_sync(0xffffffffu, bad));
    }
    return n;
}
// all open columns of row x at once: lane = column, the entry events in turn (their rows are final)
__device__ __forceinline__ void cs_finish_row_columns(const CsParams &P, int x, int lim, const int32_t *Q, const int *ent, int lane) {
    const int M = P.M;
    for (int c = lane; c < M; c += 32) {
        const int pr = P.row[(size_t)x * M + c];
        if (!(pr < lim && pr != Q[c])) continue;
        int acc = pr;
#pragma unroll 4
        for (int m = 0; m < M; m++) {
            const int ev = ent[m];
            if (ev >= 0) acc = max(acc, __ldcg(P.row + (size_t)ev * M + c));
        }
        if (acc != pr) P.row[(size_t)x * M + c] = acc;
    }
}
__device__ __forceinline__ int cs_slow_wave(const CsParams &P, int wave, int w0, int nwarps, int *ent, int lane) {
    const int M = P.M, cnt = P.slow_cnt[0];
    int done = 0;
    for (int i = w0; i < cnt; i += nwarps) {
        const int x = P.slow_list[i];
        if (P.sflag[x] != 1) continue;
        const int bx = cs_block_of(P, x), lim = cs_start(P, bx);
        const int32_t *Q = P.Qtab + (size_t)bx * M;
        for (int m = lane; m < M; m += 32) {
            const int pr = P.row[(size_t)x * M + m], q = Q[m];
            ent[m] = (pr >= lim || pr == q) ? q : pr;
        }
        __syncwarp();
        bool complete = true;
        if (cs_open_count(P, x, lim, Q, lane) > CS_COLPAR) {
            // many open columns: all at once, one lane per column -- when no entry event is itself a row in flux
            bool flux = false;
            for (int m = lane; m < M; m += 32) {
                const int ev = ent[m];
                if (ev >= P.first) { const int f = P.sflag[ev]; flux |= f == 1 || f == min(2 + wave, 250); }
            }
            if (__any_sync(0xffffffffu, flux)) continue;
            cs_finish_row_columns(P, x, lim, Q, ent, lane);
            __syncwarp();
            __threadfence();
            if (lane == 0) { P.sflag[x] = (uint8_t)min(2 + wave, 250); atomicSub(&P.slow_cnt[1], 1); }
            done++;
            continue;
        }
        for (int c0 = 0; c0 < M; c0 += 32) {
            const int c = c0 + lane;
            bool bad = false;
            if (c < M) { const int pr = P.row[(size_t)x * M + c]; bad = pr < lim && pr != Q[c]; }
            unsigned todo = __ballot_sync(0xffffffffu, bad);
            while (todo) {
                const int cb = c0 + __ffs(todo) - 1;
                todo &= todo - 1;
                // column cb of x = max over the entry events of THEIR column cb -- which is final unless the entry is
                // itself a row in flux (pending, or finished in this very wave) whose column cb is one of its open ones
                int acc = -1;
                bool wait = false;
                for (int m = lane; m < M; m += 32) {
                    const int ev = ent[m];
                    if (ev < 0) continue;
                    const int v = __ldcg(P.row + (size_t)ev * M + cb);
                    acc = max(acc, v);
                    if (ev >= P.first) {
                        const int f = P.sflag[ev];
                        if (f == 1 || f == min(2 + wave, 250)) {
                            const int be = cs_block_of(P, ev);
                            wait |= v < cs_start(P, be) && v != P.Qtab[(size_t)be * M + cb];
                        }
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc = max(acc, __shfl_xor_sync(0xffffffffu, acc, o));
                if (__any_sync(0xffffffffu, wait)) { complete = false; continue; }
                if (lane == 0 && acc > P.row[(size_t)x * M + cb]) P.row[(size_t)x * M + cb] = acc;
            }
        }
        __syncwarp();
        if (!complete) continue;                                 // some column waits for another row: next wave
        __threadfence();
        if (lane == 0) { P.sflag[x] = (uint8_t)min(2 + wave, 250); atomicSub(&P.slow_cnt[1], 1); }
        done++;
    }
    return done;
}
__global__ void __launch_bounds__(CS_SLOW_WARPS * 32) k_cs_slow_wave(CsParams P, int wave) {
    extern __shared__ int cs_smem[];
    if (P.slow_cnt[1] <= 0) return;
    const int warp = threadIdx.x >> 5;
    cs_slow_wave(P, wave, blockIdx.x * CS_SLOW_WARPS + warp, gridDim.x * CS_SLOW_WARPS, cs_smem + (size_t)warp * P.M, threadIdx.x & 31);
}
// What two grid-wide waves left pending (long dependency chains: graphs in which a member's last event of a block often
// does not see every head, e.g. two cliques with rare cross links): block after block -- a row depends only on rows of
// EARLIER blocks, so every pending row of block j can be finished once the blocks below are done.  One CTA, the rows of a
// block across its warps.
#define CS_REST_WARPS 32
__global__ void __launch_bounds__(CS_REST_WARPS * 32) k_cs_slow_rest(CsParams P) {
    extern __shared__ int cs_smem[];
    if (P.slow_cnt[1] <= 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, M = P.M;
    int *ent = cs_smem + (size_t)warp * M;
    for (int blk = 0; blk < P.nb; blk++) {
        const int cnt = P.blk_cnt[blk];
        if (cnt == 0) continue;                                  // (uniform)
        const int lim = cs_start(P, blk);
        const int32_t *Q = P.Qtab + (size_t)blk * M;
        for (int i = warp; i < cnt; i += CS_REST_WARPS) {
            const int x = P.slow_blk[lim + i];
            if (P.sflag[x] != 1) continue;
            for (int m = lane; m < M; m += 32) {
                const int pr = P.row[(size_t)x * M + m], q = Q[m];
                ent[m] = (pr >= lim || pr == q) ? q : pr;
            }
            __syncwarp();
            if (cs_open_count(P, x, lim, Q, lane) > CS_COLPAR) {
                cs_finish_row_columns(P, x, lim, Q, ent, lane);
                __syncwarp();
                if (lane == 0) P.sflag[x] = 251;
                continue;
            }
            for (int c0 = 0; c0 < M; c0 += 32) {
                const int c = c0 + lane;
                bool bad = false;
                if (c < M) { const int pr = P.row[(size_t)x * M + c]; bad = pr < lim && pr != Q[c]; }
                unsigned todo = __ballot_sync(0xffffffffu, bad);
                while (todo) {
                    const int cb = c0 + __ffs(todo) - 1;
                    todo &= todo - 1;
                    int acc = -1;
                    for (int m = lane; m < M; m += 32) {
                        const int ev = ent[m];
                        if (ev >= 0) acc = max(acc, __ldcg(P.row + (size_t)ev * M + cb));
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc = max(acc, __shfl_xor_sync(0xffffffffu, acc, o));
                    if (lane == 0 && acc > P.row[(size_t)x * M + cb]) P.row[(size_t)x * M + cb] = acc;
                }
            }
            __syncwarp();
            if (lane == 0) P.sflag[x] = 251;
        }
        __threadfence();
        __syncthreads();                                         // block blk is final before the next one reads it
    }
}

// a handful of new events (the reference's own cadence: one sync per call): straight from the parents' rows
__global__ void __launch_bounds__(1024) k_cs_small(CsParams P) {
    const int M = P.M;
    for (int h = P.first; h < P.first + P.n; h++) {
        const int a = P.p0[h], b = P.p1[h], cr = P.creator[h];
        for (int c = threadIdx.x; c < M; c += blockDim.x) {     // (a thread re-reads only columns it wrote itself)
            const int x = a >= 0 ? P.row[(size_t)a * M + c] : -1, y = b >= 0 ? P.row[(size_t)b * M + c] : -1;
            P.row[(size_t)h * M + c] = c == cr ? h : max(x, y);
        }
        if (threadIdx.x == 0) P.carry[cr] = h;
    }
}

// Cross-GPU barrier between two kernels of the scan (several ranks): one warp; rank r stores `count` into slot r of every
// peer's flag row (after a system-scope fence: everything this rank stored into the peers' tables before is visible first),
// then waits until its own row shows `count` from every rank.  Bounded: ~4 s, then the engine's error flag.
__global__ void k_xbarrier(unsigned *const *peer_flags, int rank, int nranks, unsigned count, int32_t *scal) {
    const int t = threadIdx.x;
    __threadfence_system();
    if (t < nranks) asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(peer_flags[t] + rank), "r"(count) : "memory");
    if (t < nranks) {
        const long long t0 = clock64();
        unsigned v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(peer_flags[rank] + t) : "memory");
            if ((int)(v - count) >= 0) break;
            if (clock64() - t0 > 8000000000ll) { atomicMin(&scal[SC_ERR], -4); break; }
        }
    }
    __threadfence_system();
}

// several ranks: a rank may own no column tile at all, so the carry heads are installed by their own little kernel
__global__ void k_cs_carry(CsParams P) {
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < P.M; m += gridDim.x * blockDim.x) P.carry[m] = P.Qtab[(size_t)P.nb * P.M + m];
}
