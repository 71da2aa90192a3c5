// swirld_b200.cu -- host side of libswirld_b200.so: the C ABI of include/swirld_b200.h
// over the kernels in swirld_kernels.cuh.  No torch, no CPU compute path: every
// consensus result is produced by a kernel; the host only validates the graph shape
// on append (what Node.is_valid_event checks, swirld.py:104-108), keeps the
// creator/height/chain-position mirrors it needs for that, and moves bytes.
#include "swirld_kernels.cuh"
#include "swirld_divide.cuh"
#include "swirld_levels.cuh"
#include "swirld_cansee.cuh"
#include "swirld_rounds.cuh"

#include <cstdlib>
#include "../../include/swirld_b200.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

std::string g_create_error;

struct TimedSpan { cudaEvent_t a, b; int cat; };

}  // namespace

struct sw_engine {
    int M = 0, NC = 1, cap = 0, C = 6, device = 0, Rcap = 0;
    bool unit = true;
    i64 tot = 0;
    std::vector<i64> h_stake;
    // host mirrors for validation / views
    std::vector<int32_t> h_creator, h_head, h_count;
    int32_t *h_height = nullptr, *h_seq = nullptr;   // pinned, cap entries: sources of asynchronous copies
    cudaStream_t copy_stream = nullptr;              // sw_append's copies run beside the kernels of earlier chunks
    struct PendingAppend { int base; cudaEvent_t done; };
    std::vector<PendingAppend> appends;              // copies (+ eager can_see scans) the compute stream has not waited for yet
    int n_events = 0, n_divided = 0, n_tx = 0;
    // device columns
    int32_t *d_p0 = nullptr, *d_p1 = nullptr, *d_creator = nullptr, *d_seq = nullptr, *d_height = nullptr;
    int32_t *d_hist = nullptr, *d_cursor = nullptr, *d_order = nullptr, *d_gpos = nullptr, *d_lvl_start = nullptr;
    GDesc *d_gdesc = nullptr;
    long long *d_dbg = nullptr;
    unsigned rb_epoch = 0;        // launches of k_rounds_batch (mask-cache key)
    int divide_impl = 5;          // 5 = round-batch on the whole GPU (default), 4 = level walker, 3 = per-event flags
    int n_sm = 0;
    int32_t *d_Wf = nullptr, *d_cev = nullptr, *d_rbmeta = nullptr, *d_rbtot = nullptr, *d_gchain = nullptr;   // round-batch state
    ulonglong2 *d_sc = nullptr;
    uint8_t *d_res = nullptr;
    int cansee_scan = 0;          // 1 = can_see by the blocked scan k_cs_* (SW_CANSEE_IMPL=scan), 0 = fused into the walker
    int n_rowed = 0;              // events whose can_see row is complete (cansee_scan)
    uint8_t *d_exported = nullptr;
    int32_t *d_exp_list = nullptr, *d_exp_m = nullptr, *d_exp_cnt = nullptr, *d_cs_last = nullptr, *d_cs_Q = nullptr, *d_cs_carry = nullptr;
    double *d_t = nullptr;
    uint8_t *d_sig = nullptr;
    int32_t *d_row = nullptr, *d_round = nullptr;
    u64 *d_T = nullptr, *d_SM = nullptr;
    uint8_t *d_wit = nullptr;
    int8_t *d_famous_ev = nullptr;
    // per round
    int32_t *d_W = nullptr, *d_rem = nullptr, *d_newc = nullptr;
    u64 *d_S = nullptr;
    int8_t *d_famous = nullptr;
    uint8_t *d_consensus = nullptr, *d_done = nullptr, *d_coin = nullptr;
    i64 *d_stake = nullptr;
    int32_t *d_scal = nullptr;
    // find_order
    int32_t *d_lastord = nullptr, *d_tx = nullptr, *d_idx = nullptr, *d_batch_ev = nullptr,
            *d_batch_seg = nullptr, *d_seg_start = nullptr, *d_seg_fw = nullptr, *d_seg_nf = nullptr,
            *d_perm = nullptr, *d_rounds_in = nullptr, *d_plan = nullptr;
    uint8_t *d_seg_white = nullptr;
    double *d_ts = nullptr;
    u64 *d_key = nullptr;
    int seg_cap = 0;
    void *d_flush = nullptr;
    size_t flush_bytes = 0;
    int32_t *h_scal = nullptr;    // pinned
    int32_t *h_newc = nullptr;    // pinned, Rcap
    cudaStream_t stream = nullptr;
    cudaEvent_t user_ev[16] = {nullptr};
    std::vector<TimedSpan> spans;
    std::vector<cudaEvent_t> pool;
    sw_stats_t stats{};
    std::string err;
};

namespace {

int fail(sw_engine *e, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (e) e->err = buf; else g_create_error = buf;
    return code;
}

#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t _s = (call);                                                         \
        if (_s != cudaSuccess)                                                           \
            return fail(e, SW_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(_s),   \
                        __FILE__, __LINE__);                                             \
    } while (0)

template <typename T>
cudaError_t dalloc(T **p, size_t n) { return cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T)); }

cudaEvent_t get_event(sw_engine *e) {
    if (!e->pool.empty()) { cudaEvent_t ev = e->pool.back(); e->pool.pop_back(); return ev; }
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    return ev;
}

struct Span {
    sw_engine *e; TimedSpan s;
    Span(sw_engine *e_, int cat) : e(e_) { s.a = get_event(e); s.b = get_event(e); s.cat = cat; cudaEventRecord(s.a, e->stream); }
    ~Span() { cudaEventRecord(s.b, e->stream); e->spans.push_back(s); }
};

void fold_spans(sw_engine *e) {
    std::vector<TimedSpan> pending;
    for (auto &s : e->spans) {
        if (cudaEventQuery(s.b) == cudaErrorNotReady) { pending.push_back(s); continue; }   // (a scan still running on the copy stream)
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
            if (s.cat == 0) e->stats.ms_divide_rounds += ms;
            else if (s.cat == 1) e->stats.ms_decide_fame += ms;
            else if (s.cat == 2) e->stats.ms_find_order += ms;
            else if (s.cat == 3) { e->stats.ms_can_see += ms; e->stats.ms_divide_rounds += ms; }
            else if (s.cat == 4) e->stats.ms_rounds_kernel += ms;
        }
        e->pool.push_back(s.a); e->pool.push_back(s.b);
    }
    e->spans.swap(pending);
}

// Make the compute stream wait for the appended batches that start below `upto` (all of them: upto < 0).
int wait_appends(sw_engine *e, int upto) {
    size_t k = 0;
    for (auto &a : e->appends) {
        if (upto >= 0 && a.base >= upto) { e->appends[k++] = a; continue; }
        CK(cudaStreamWaitEvent(e->stream, a.done, 0));
        e->pool.push_back(a.done);           // (re-recorded only after later work was enqueued behind the wait)
    }
    e->appends.resize(k);
    return 0;
}

int device_error(sw_engine *e) {     // after a sync: did a kernel flag an error?
    int code = e->h_scal[SC_ERR];
    if (code < 0) {
        const char *what = code == SW_E_CAPACITY ? "round table exhausted"
                         : code == SW_E_INDEX ? "list index out of range (swirld.py:305: a single seer)"
                         : code == SW_E_KEY ? "KeyError (undecided witness in a consensus round)" : "device error";
        return fail(e, code, "%s", what);
    }
    return 0;
}

int reset_state(sw_engine *e, bool keep_events = false) {
    const size_t RM = (size_t)e->Rcap * e->M;
    k_fill_i32<<<256, 256, 0, e->stream>>>(e->d_W, -1, RM);
    CK(cudaMemsetAsync(e->d_famous, 0xff, RM, e->stream));
    CK(cudaMemsetAsync(e->d_S, 0, RM * sizeof(u64), e->stream));
    CK(cudaMemsetAsync(e->d_consensus, 0, e->Rcap, e->stream));
    CK(cudaMemsetAsync(e->d_famous_ev, 0xff, e->cap, e->stream));
    k_fill_i32<<<256, 256, 0, e->stream>>>(e->d_idx, -1, (size_t)e->cap);
    k_fill_i32<<<1, 64, 0, e->stream>>>(e->d_lastord, -1, (size_t)e->M);
    k_fill_i32<<<1, 64, 0, e->stream>>>(e->d_cs_carry, -1, (size_t)64);
    k_fill_i32<<<256, 256, 0, e->stream>>>(e->d_Wf, -1, RM);
    CK(cudaMemsetAsync(e->d_rbtot, 0, sizeof(int32_t) * 64, e->stream));
    k_fill_i32<<<64, 256, 0, e->stream>>>(e->d_gchain, -1, (size_t)64 * RB_RING);
    CK(cudaMemsetAsync(e->d_sc, 0, sizeof(ulonglong2) * (size_t)e->cap, e->stream));
    int32_t sc[SC_COUNT] = {0};
    sc[SC_MAX_ROUND] = -1;
    CK(cudaMemcpyAsync(e->d_scal, sc, sizeof sc, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->stats.kernel_launches += 3;
    e->n_divided = e->n_tx = 0;
    e->n_rowed = 0;
    if (!keep_events) {
        e->n_events = 0;
        std::fill(e->h_head.begin(), e->h_head.end(), -1);
        std::fill(e->h_count.begin(), e->h_count.end(), 0);
    }
    memset(e->h_scal, 0, sizeof(int32_t) * SC_COUNT);
    return 0;
}

template <int NC, bool UNIT>
int launch_divide(sw_engine *e, const DivParams &P) {
    const size_t smem = sizeof(DivSmem<NC>);
    CK(cudaFuncSetAttribute(k_divide<NC, UNIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_divide<NC, UNIT><<<1, 1024, smem, e->stream>>>(P);
    CK(cudaGetLastError());
    return 0;
}

template <int NC, bool UNIT, bool ROWS, bool FULL>
int launch_levels4(sw_engine *e, const Div4Params &Q) {
    const size_t smem = sizeof(LvSmem<NC, ROWS>);
    CK(cudaFuncSetAttribute(k_divide_levels<NC, UNIT, ROWS, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        cudaEvent_t a = get_event(e), b = get_event(e);
        cudaEventRecord(a, e->stream);
        k_divide_levels<NC, UNIT, ROWS, FULL><<<1, LV_THREADS, smem, e->stream>>>(Q);
        cudaEventRecord(b, e->stream);
        e->spans.push_back(TimedSpan{a, b, 4});
    }
    CK(cudaGetLastError());
    return 0;
}
template <int NC, bool UNIT, bool ROWS>
int launch_levels(sw_engine *e, const Div4Params &Q) {
    return e->M == NC * 32 ? launch_levels4<NC, UNIT, ROWS, true>(e, Q) : launch_levels4<NC, UNIT, ROWS, false>(e, Q);
}

// can_see rows of every appended event that does not have one yet: blocked scan (k_cs_*)
// `st`: the compute stream (lazily, from sw_divide_rounds) or the copy stream (eagerly, from sw_append:
// the scan of a new chunk then runs beside the round kernel of the previous one)
template <int NC>
int cansee_scan(sw_engine *e, cudaStream_t st, int upto) {
    const int first = e->n_rowed, n = upto - e->n_rowed;
    if (n <= 0) return 0;
    CsParams C{};
    C.M = e->M; C.first = first; C.n = n;
    // block length: the in-block walks are serial in B, the boundary pass in n/B (SW_CS_B overrides)
    C.B = std::min(n, n >= 200000 ? 4096 : 1024);
    if (const char *v = getenv("SW_CS_B")) C.B = std::max(64, std::min(n, atoi(v)));
    C.nb = (n + C.B - 1) / C.B;
    C.p0 = e->d_p0; C.p1 = e->d_p1; C.creator = e->d_creator; C.row = e->d_row;
    C.exported = e->d_exported; C.exp_list = e->d_exp_list; C.exp_m = e->d_exp_m; C.exp_cnt = e->d_exp_cnt;
    C.last = e->d_cs_last; C.Qtab = e->d_cs_Q; C.carry = e->d_cs_carry;
    CK(cudaMemsetAsync(e->d_exported + first, 0, (size_t)n, st));
    CK(cudaMemsetAsync(e->d_exp_cnt, 0, sizeof(int32_t) * (size_t)C.nb, st));
    cudaEvent_t a = get_event(e), b = get_event(e);
    cudaEventRecord(a, st);
    k_cs_local<NC, 1><<<C.nb, NC * 32, 0, st>>>(C);
    k_cs_collect<<<std::max(1, std::min(296, (n + 255) / 256)), 256, 0, st>>>(C);
    k_cs_boundary<NC><<<1, 1024, 0, st>>>(C);
    k_cs_local<NC, 2><<<C.nb, NC * 32, 0, st>>>(C);
    cudaEventRecord(b, st);
    e->spans.push_back(TimedSpan{a, b, 3});
    CK(cudaGetLastError());
    e->stats.kernel_launches += 4;
    e->n_rowed = upto;
    return 0;
}

// counting sort of the chunk by height, then the level-scheduled walk
int divide_levels(sw_engine *e, const DivParams &P) {
    int hmin = e->h_height[P.first], hmax = hmin;
    for (int i = P.first; i < P.first + P.n; i++) { hmin = std::min(hmin, e->h_height[i]); hmax = std::max(hmax, e->h_height[i]); }
    LvlParams L{};
    L.first = P.first; L.n = P.n; L.hmin = hmin; L.nbins = hmax - hmin + 1;
    L.height = e->d_height; L.p0 = e->d_p0; L.p1 = e->d_p1; L.creator = e->d_creator;
    L.hist = e->d_hist; L.cursor = e->d_cursor; L.order = e->d_order; L.gpos = e->d_gpos;
    L.lvl_start = e->d_lvl_start; L.scal = e->d_scal; L.gdesc = e->d_gdesc;
    CK(cudaMemsetAsync(e->d_hist, 0, sizeof(int32_t) * (size_t)L.nbins, e->stream));
    const int blocks = std::max(1, std::min(296, (P.n + 255) / 256));
    k_lvl_hist<<<blocks, 256, 0, e->stream>>>(L);
    k_lvl_scan<<<1, 1024, 0, e->stream>>>(L);
    k_lvl_scatter<<<blocks, 256, 0, e->stream>>>(L);
    k_lvl_desc<<<blocks, 256, 0, e->stream>>>(L);
    CK(cudaGetLastError());
    Div4Params Q{};
    Q.d = P; Q.gdesc = e->d_gdesc; Q.lvl_start = e->d_lvl_start;
    e->stats.kernel_launches += 4;
    if (e->cansee_scan) {
        if (e->NC == 1) return e->unit ? launch_levels<1, true, false>(e, Q) : launch_levels<1, false, false>(e, Q);
        return e->unit ? launch_levels<2, true, false>(e, Q) : launch_levels<2, false, false>(e, Q);
    }
    if (e->NC == 1) return e->unit ? launch_levels<1, true, true>(e, Q) : launch_levels<1, false, true>(e, Q);
    return e->unit ? launch_levels<2, true, true>(e, Q) : launch_levels<2, false, true>(e, Q);
}

// rounds of the chunk by the cooperative round-batch kernel (swirld_rounds.cuh)
template <int NC, bool UNIT>
int divide_round_batch(sw_engine *e, const DivParams &D) {
    RbParams R{};
    R.M = e->M; R.first = D.first; R.n = D.n; R.Rcap = e->Rcap;
    // a few SMs stay free for the can_see scan of the next chunk, which runs beside this kernel (SW_RB_FREE_SMS)
    int free_sms = 16;
    if (const char *v = getenv("SW_RB_FREE_SMS")) free_sms = std::max(0, atoi(v));
    const int grid = std::max(e->n_sm / 2, e->n_sm - free_sms);
    R.L = std::max(1, std::min(RB_LMAX, grid * (RB_THREADS / 32) / e->M));
    R.maxmiss = RB_MAXMISS;
    R.epoch = ++e->rb_epoch;
    if (const char *v = getenv("SW_RB_L")) R.L = std::max(1, std::min(R.L, atoi(v)));          // tuning knobs
    if (const char *v = getenv("SW_RB_MAXMISS")) R.maxmiss = std::max(0, atoi(v));
    R.row = e->d_row; R.p0 = e->d_p0; R.creator = e->d_creator; R.seq = e->d_seq; R.round = e->d_round;
    R.Wf = e->d_Wf; R.sc = e->d_sc; R.cev = e->d_cev;
    R.ccnt = e->d_rbmeta; R.cmin = e->d_rbmeta + 64; R.coff = e->d_rbmeta + 128; R.bar = reinterpret_cast<unsigned *>(e->d_rbmeta + 224);
    R.ctot = e->d_rbtot; R.gchain = e->d_gchain;
    R.res = e->d_res; R.stake = e->d_stake; R.tot2 = D.tot2; R.scal = e->d_scal;
    R.wit = e->d_wit; R.W = e->d_W; R.SM = e->d_SM; R.dbg = e->d_dbg;
    R.wlist = e->d_cev + e->cap; R.wcnt = e->d_rbmeta + 225;
    CK(cudaMemsetAsync(R.ccnt, 0, sizeof(int32_t) * 64, e->stream));
    CK(cudaMemsetAsync(R.cmin, 0x7f, sizeof(int32_t) * 64, e->stream));
    const int blocks = std::max(1, std::min(296, (D.n + 255) / 256));
    k_rb_count<<<blocks, 256, 0, e->stream>>>(R);
    k_rb_offsets<<<1, 32, 0, e->stream>>>(R);
    k_rb_scatter<<<blocks, 256, 0, e->stream>>>(R);
    CK(cudaGetLastError());
    void *args[] = {(void *)&R};
    {
        cudaEvent_t a = get_event(e), b = get_event(e);
        cudaEventRecord(a, e->stream);
        CK(cudaLaunchCooperativeKernel((void *)k_rounds_batch<NC, UNIT>, dim3(grid), dim3(RB_THREADS), args, 0, e->stream));
        cudaEventRecord(b, e->stream);
        e->spans.push_back(TimedSpan{a, b, 4});
    }
    k_rb_tail<<<blocks, 256, 0, e->stream>>>(R);
    k_rb_witness<<<blocks, 256, 0, e->stream>>>(R);
    k_rb_seenmask<NC><<<(D.n + 7) / 8, 256, 0, e->stream>>>(R);
    CK(cudaGetLastError());
    e->stats.kernel_launches += 7;
    return 0;
}

}  // namespace

extern "C" {

int sw_version(void) { return 100; }

const char *sw_last_error(const sw_engine *e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int sw_create(int M, int capacity_events, const int64_t *stake, int coin_period, int device,
              sw_engine **out) {
    sw_engine *e = nullptr;
    if (!out) return fail(e, SW_E_ARG, "out is NULL");
    *out = nullptr;
    if (M < 1 || capacity_events < 1 || coin_period < 1) return fail(e, SW_E_ARG, "bad M / capacity / coin period");
    if (M > SW_MAX_MEMBERS) return fail(e, SW_E_UNSUPPORTED, "M=%d > %d members not supported by this build", M, SW_MAX_MEMBERS);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail(e, SW_E_CUDA, "no CUDA device (this engine has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(e, SW_E_ARG, "device %d out of range (%d devices)", device, ndev);
    e = new sw_engine();
    e->M = M; e->NC = (M + 31) / 32; e->cap = capacity_events; e->C = coin_period; e->device = device;
    e->h_stake.resize(M);
    for (int c = 0; c < M; c++) {
        e->h_stake[c] = stake ? stake[c] : 1;
        if (e->h_stake[c] < 0) { delete e; return fail(nullptr, SW_E_ARG, "negative stake"); }
        if (e->h_stake[c] != 1) e->unit = false;
        e->tot += e->h_stake[c];
    }
    // a round other than the last needs more than 2*tot/3 members with a witness (quirk Q3)
    i64 per = std::min<i64>(M, (2 * e->tot) / 3 + 1);
    if (per < 1) per = 1;
    e->Rcap = (int)std::min<i64>((i64)e->cap + 2, (i64)e->cap / per + 16);
    if (const char *impl = getenv("SW_DIVIDE_IMPL")) { int v = atoi(impl); e->divide_impl = (v == 3 || v == 4) ? v : 5; }
    if (const char *impl = getenv("SW_CANSEE_IMPL")) e->cansee_scan = strcmp(impl, "scan") == 0 ? 1 : 0;
    if (e->divide_impl == 3) e->cansee_scan = 0;
    if (e->divide_impl == 5) e->cansee_scan = 1;     // the round-batch kernel reads finished can_see rows
    e->h_head.assign(M, -1);
    e->h_count.assign(M, 0);
    e->h_creator.reserve(e->cap);
    int rc = [&]() -> int {
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
        const size_t cap = e->cap, RM = (size_t)e->Rcap * M;
        CK(cudaMallocHost((void **)&e->h_height, sizeof(int32_t) * cap));
        CK(cudaMallocHost((void **)&e->h_seq, sizeof(int32_t) * cap));
        CK(dalloc(&e->d_p0, cap)); CK(dalloc(&e->d_p1, cap)); CK(dalloc(&e->d_creator, cap)); CK(dalloc(&e->d_seq, cap));
        CK(dalloc(&e->d_t, cap)); CK(dalloc(&e->d_sig, cap * 64)); CK(dalloc(&e->d_height, cap));
        CK(dalloc(&e->d_hist, cap + 2)); CK(dalloc(&e->d_cursor, cap + 2)); CK(dalloc(&e->d_order, cap));
        CK(dalloc(&e->d_gpos, cap)); CK(dalloc(&e->d_lvl_start, cap + 2)); CK(dalloc(&e->d_gdesc, cap));
        CK(dalloc(&e->d_exported, cap)); CK(dalloc(&e->d_exp_list, cap + 8192)); CK(dalloc(&e->d_exp_m, cap + 8192)); CK(dalloc(&e->d_exp_cnt, cap / 64 + 4));
        CK(dalloc(&e->d_cs_last, (cap / 64 + 4) * (size_t)M)); CK(dalloc(&e->d_cs_Q, (cap / 64 + 5) * (size_t)M));
        CK(dalloc(&e->d_cs_carry, (size_t)64));
        CK(dalloc(&e->d_Wf, RM)); CK(dalloc(&e->d_cev, 2 * cap)); /* + the witness list of the current chunk */ CK(dalloc(&e->d_rbmeta, (size_t)256));
        CK(dalloc(&e->d_sc, cap)); CK(dalloc(&e->d_res, (size_t)2 * 64 * RB_LMAX));
        CK(dalloc(&e->d_rbtot, (size_t)64)); CK(dalloc(&e->d_gchain, (size_t)64 * RB_RING));
        CK(cudaDeviceGetAttribute(&e->n_sm, cudaDevAttrMultiProcessorCount, device));
        CK(dalloc(&e->d_dbg, (size_t)40)); CK(cudaMemsetAsync(e->d_dbg, 0, sizeof(long long) * 40, e->stream));
        CK(dalloc(&e->d_row, cap * M)); CK(dalloc(&e->d_SM, cap));
        if (e->divide_impl != 5) CK(dalloc(&e->d_T, cap * M));           // strongly-sees matrices: level walker only
        CK(dalloc(&e->d_round, cap)); CK(dalloc(&e->d_wit, cap)); CK(dalloc(&e->d_famous_ev, cap));
        CK(dalloc(&e->d_W, RM)); CK(dalloc(&e->d_S, RM)); CK(dalloc(&e->d_famous, RM));
        CK(dalloc(&e->d_consensus, (size_t)e->Rcap)); CK(dalloc(&e->d_done, (size_t)e->Rcap)); CK(dalloc(&e->d_coin, RM));
        CK(dalloc(&e->d_rem, (size_t)e->Rcap)); CK(dalloc(&e->d_newc, (size_t)e->Rcap));
        CK(dalloc(&e->d_stake, (size_t)M)); CK(dalloc(&e->d_scal, (size_t)SC_COUNT));
        CK(dalloc(&e->d_lastord, (size_t)64)); CK(dalloc(&e->d_tx, cap)); CK(dalloc(&e->d_idx, cap));
        CK(dalloc(&e->d_batch_ev, cap)); CK(dalloc(&e->d_batch_seg, cap)); CK(dalloc(&e->d_perm, 2 * cap));
        CK(dalloc(&e->d_ts, cap)); CK(dalloc(&e->d_key, cap * 8));
        CK(cudaMallocHost((void **)&e->h_scal, sizeof(int32_t) * SC_COUNT));
        CK(cudaMallocHost((void **)&e->h_newc, sizeof(int32_t) * e->Rcap));
        CK(cudaMemcpyAsync(e->d_stake, e->h_stake.data(), sizeof(i64) * M, cudaMemcpyHostToDevice, e->stream));
        return reset_state(e);
    }();
    if (rc < 0) { g_create_error = e->err; sw_destroy(e); return rc; }
    memset(&e->stats, 0, sizeof e->stats);
    *out = e;
    return SW_OK;
}

void sw_destroy(sw_engine *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) { wait_appends(e, -1); cudaStreamSynchronize(e->stream); }
    fold_spans(e);
    for (auto ev : e->pool) cudaEventDestroy(ev);
    for (auto ev : e->user_ev) if (ev) cudaEventDestroy(ev);
    void *ptrs[] = {e->d_rbtot, e->d_gchain, e->d_Wf, e->d_cev, e->d_rbmeta, e->d_sc, e->d_res, e->d_cs_last, e->d_cs_Q, e->d_cs_carry, e->d_exported, e->d_exp_list, e->d_exp_m, e->d_exp_cnt, e->d_coin, e->d_dbg, e->d_height, e->d_hist, e->d_cursor, e->d_order, e->d_gpos, e->d_lvl_start, e->d_gdesc,
                    e->d_p0, e->d_p1, e->d_creator, e->d_seq, e->d_t, e->d_sig, e->d_row, e->d_T, e->d_SM,
                    e->d_round, e->d_wit, e->d_famous_ev, e->d_W, e->d_S, e->d_famous, e->d_consensus,
                    e->d_done, e->d_rem, e->d_newc, e->d_stake, e->d_scal, e->d_lastord, e->d_tx, e->d_idx,
                    e->d_batch_ev, e->d_batch_seg, e->d_perm, e->d_ts, e->d_key, e->d_seg_start, e->d_seg_fw,
                    e->d_seg_nf, e->d_seg_white, e->d_rounds_in, e->d_plan, e->d_flush};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (e->h_scal) cudaFreeHost(e->h_scal);
    if (e->h_newc) cudaFreeHost(e->h_newc);
    if (e->h_height) cudaFreeHost(e->h_height);
    if (e->h_seq) cudaFreeHost(e->h_seq);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

int sw_reset(sw_engine *e) {
    if (!e) return SW_E_ARG;
    CK(cudaSetDevice(e->device));
    if (wait_appends(e, -1) < 0) return SW_E_CUDA;
    CK(cudaStreamSynchronize(e->stream));
    fold_spans(e);
    e->h_creator.clear();
    int rc = reset_state(e);
    memset(&e->stats, 0, sizeof e->stats);
    return rc;
}

int sw_rewind(sw_engine *e) {
    if (!e) return SW_E_ARG;
    CK(cudaSetDevice(e->device));
    if (wait_appends(e, -1) < 0) return SW_E_CUDA;
    CK(cudaStreamSynchronize(e->stream));
    fold_spans(e);
    return reset_state(e, true);
}

int sw_event_record(sw_engine *e, int slot) {
    if (!e || slot < 0 || slot >= 16) return fail(e, SW_E_ARG, "bad event slot");
    CK(cudaSetDevice(e->device));
    if (!e->user_ev[slot]) CK(cudaEventCreate(&e->user_ev[slot]));
    CK(cudaEventRecord(e->user_ev[slot], e->stream));
    return SW_OK;
}

int sw_event_elapsed_ms(sw_engine *e, int a, int b, double *ms_out) {
    if (!e || a < 0 || a >= 16 || b < 0 || b >= 16 || !ms_out || !e->user_ev[a] || !e->user_ev[b])
        return fail(e, SW_E_ARG, "bad event slot");
    CK(cudaSetDevice(e->device));
    CK(cudaEventSynchronize(e->user_ev[b]));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e->user_ev[a], e->user_ev[b]));
    *ms_out = ms;
    return SW_OK;
}

int sw_append(sw_engine *e, int n, const int32_t *p0, const int32_t *p1, const int32_t *creator,
              const double *t, const uint8_t *sig) {
    if (!e || n < 0 || (n > 0 && (!p0 || !p1 || !creator || !t || !sig))) return fail(e, SW_E_ARG, "bad argument");
    if (n == 0) return SW_OK;
    if ((i64)e->n_events + n > e->cap) return fail(e, SW_E_CAPACITY, "capacity_events=%d exceeded", e->cap);
    CK(cudaSetDevice(e->device));
    const int base = e->n_events;
    // graph-shape checks of is_valid_event (swirld.py:104-108) + the fork-free contract
    std::vector<int32_t> head_save(e->h_head), count_save(e->h_count);
    e->h_creator.resize((size_t)base + n);
    int rc = SW_OK;
    for (int j = 0; j < n && rc == SW_OK; j++) {
        const int i = base + j, c = creator[j], a = p0[j], b = p1[j];
        if (c < 0 || c >= e->M) { rc = fail(e, SW_E_ARG, "event %d: creator %d out of range", i, c); break; }
        if (a < 0 && b < 0) {
            if (e->h_head[c] >= 0) { rc = fail(e, SW_E_FORK, "event %d: second root of member %d", i, c); break; }
            e->h_height[i] = 0;                                          // swirld.py:117-118
        } else {
            if (a < 0 || b < 0 || a >= i || b >= i) { rc = fail(e, SW_E_PARENT, "event %d: parents (%d,%d) unknown", i, a, b); break; }
            if (e->h_creator[a] != c) { rc = fail(e, SW_E_PARENT, "event %d: self-parent %d has another creator", i, a); break; }
            if (e->h_creator[b] == c) { rc = fail(e, SW_E_PARENT, "event %d: other-parent %d has the same creator", i, b); break; }
            if (e->h_head[c] != a) { rc = fail(e, SW_E_FORK, "event %d: self-parent %d is not member %d's latest event (fork)", i, a, c); break; }
            e->h_height[i] = std::max(e->h_height[a], e->h_height[b]) + 1;   // swirld.py:120
        }
        e->h_creator[i] = c;
        e->h_head[c] = i;
        e->h_seq[i] = e->h_count[c]++;
    }
    if (rc != SW_OK) {
        e->h_head = head_save; e->h_count = count_save;
        e->h_creator.resize(base);
        return rc;
    }
    // The copies go to their own stream: they touch only the new rows, so they overlap the kernels of
    // earlier chunks still running on the compute stream; later compute work waits for them.
    cudaStream_t cs = e->copy_stream;
    CK(cudaMemcpyAsync(e->d_p0 + base, p0, sizeof(int32_t) * n, cudaMemcpyHostToDevice, cs));
    CK(cudaMemcpyAsync(e->d_p1 + base, p1, sizeof(int32_t) * n, cudaMemcpyHostToDevice, cs));
    CK(cudaMemcpyAsync(e->d_creator + base, creator, sizeof(int32_t) * n, cudaMemcpyHostToDevice, cs));
    CK(cudaMemcpyAsync(e->d_t + base, t, sizeof(double) * n, cudaMemcpyHostToDevice, cs));
    CK(cudaMemcpyAsync(e->d_sig + (size_t)base * 64, sig, (size_t)64 * n, cudaMemcpyHostToDevice, cs));
    CK(cudaMemcpyAsync(e->d_seq + base, e->h_seq + base, sizeof(int32_t) * n, cudaMemcpyHostToDevice, cs));
    CK(cudaMemcpyAsync(e->d_height + base, e->h_height + base, sizeof(int32_t) * n, cudaMemcpyHostToDevice, cs));
    // rows are up to date and the batch is big: scan it now, beside the kernels of the previous chunk
    const bool eager = e->cansee_scan && e->n_rowed == base && n >= 4096;
    e->stats.h2d_bytes += (i64)n * (5 * 4 + 8 + 64);
    e->stats.events += n;
    e->n_events += n;
    if (eager) {
        int rc2 = e->NC == 1 ? cansee_scan<1>(e, cs, e->n_events) : cansee_scan<2>(e, cs, e->n_events);
        if (rc2 < 0) return rc2;
    }
    // the compute stream waits for this batch only when a call first touches it (wait_appends)
    cudaEvent_t done = get_event(e);
    CK(cudaEventRecord(done, cs));
    e->appends.push_back({base, done});
    return SW_OK;
}

int sw_divide_rounds(sw_engine *e, int first, int n) {
    if (!e || n < 0) return fail(e, SW_E_ARG, "bad argument");
    if (n == 0) return SW_OK;
    if (first != e->n_divided) return fail(e, SW_E_ARG, "divide_rounds: first=%d but %d events are divided (events must arrive in order)", first, e->n_divided);
    if (first + n > e->n_events) return fail(e, SW_E_KEY, "divide_rounds: events [%d,%d) not appended", first, first + n);
    CK(cudaSetDevice(e->device));
    if (wait_appends(e, first + n) < 0) return SW_E_CUDA;
    DivParams P{};
    P.M = e->M; P.first = first; P.n = n; P.Rcap = e->Rcap;
    P.p0 = e->d_p0; P.p1 = e->d_p1; P.creator = e->d_creator;
    P.row = e->d_row; P.T = e->d_T; P.SM = e->d_SM; P.round = e->d_round; P.wit = e->d_wit; P.W = e->d_W;
    P.stake = e->d_stake; P.tot2 = 2 * e->tot; P.unit = e->unit ? 1 : 0; P.scal = e->d_scal; P.dbg = e->d_dbg;
    StrongParams Q{};
    Q.M = e->M; Q.first = first; Q.n = n; Q.Rcap = e->Rcap; Q.creator = e->d_creator; Q.row = e->d_row;
    Q.round = e->d_round; Q.wit = e->d_wit; Q.SM = e->d_SM; Q.S = e->d_S; Q.stake = e->d_stake; Q.tot2 = 2 * e->tot;
    Q.coin = e->d_coin; Q.sig = e->d_sig; Q.unit = e->unit ? 1 : 0;
    if (e->cansee_scan && first + n > e->n_rowed) {
        // rows are behind (small appends, or after sw_rewind): scan everything appended so far, here.
        // (Scanning only [n_rowed, first+n) here and the rest on the copy stream beside the round kernels was
        // measured: 1% faster on average at C3, but with a visible run-to-run spread.)
        int rc = e->NC == 1 ? cansee_scan<1>(e, e->stream, e->n_events) : cansee_scan<2>(e, e->stream, e->n_events);
        if (rc < 0) return rc;
    }
    {
        Span sp(e, 0);
        int rc;
        if (e->divide_impl == 5)
            rc = e->NC == 1 ? (e->unit ? divide_round_batch<1, true>(e, P) : divide_round_batch<1, false>(e, P))
                            : (e->unit ? divide_round_batch<2, true>(e, P) : divide_round_batch<2, false>(e, P));
        else if (e->divide_impl == 4) rc = divide_levels(e, P);
        else rc = e->NC == 1 ? (e->unit ? launch_divide<1, true>(e, P) : launch_divide<1, false>(e, P))
                             : (e->unit ? launch_divide<2, true>(e, P) : launch_divide<2, false>(e, P));
        if (rc < 0) return rc;
        const int wpb = 8;
        int blocks = (n + wpb - 1) / wpb;
        if (e->divide_impl == 5) {                       // the round-batch path leaves the list of the chunk's witnesses
            Q.list = e->d_cev + e->cap; Q.list_n = e->d_rbmeta + 225;
            blocks = std::min(blocks, 4 * e->n_sm);
        }
        if (e->NC == 1) k_strong<1><<<blocks, wpb * 32, 0, e->stream>>>(Q);
        else k_strong<2><<<blocks, wpb * 32, 0, e->stream>>>(Q);
        CK(cudaGetLastError());
    }
    e->stats.kernel_launches += e->divide_impl == 5 ? 1 : 2;      // k_strong (+ the walker; the round-batch kernels count themselves)
    e->stats.events_divided += n;
    e->n_divided += n;
    return SW_OK;
}

int sw_decide_fame(sw_engine *e, int32_t *new_c_out, int cap) {
    if (!e || cap < 0 || (cap > 0 && !new_c_out)) return fail(e, SW_E_ARG, "bad argument");
    CK(cudaSetDevice(e->device));
    if (e->n_divided == 0) return fail(e, SW_E_ARG, "decide_fame: no witnesses yet (max() of an empty dict, swirld.py:225)");
    FameParams P{};
    P.M = e->M; P.Rcap = e->Rcap; P.C = e->C; P.W = e->d_W; P.S = e->d_S; P.famous = e->d_famous;
    P.famous_ev = e->d_famous_ev; P.consensus = e->d_consensus; P.done = e->d_done; P.rem = e->d_rem;
    P.coin = e->d_coin; P.stake = e->d_stake; P.tot2 = 2 * e->tot; P.unit = e->unit ? 1 : 0; P.newc = e->d_newc; P.scal = e->d_scal;
    {
        Span sp(e, 1);
        k_fame_begin<<<1, 32, 0, e->stream>>>(P);
        k_fame_rounds<<<2 * e->n_sm, 256, 0, e->stream>>>(P);
        k_fame_finish<<<1, 1024, 0, e->stream>>>(P);
        CK(cudaGetLastError());
    }
    e->stats.kernel_launches += 3;
    CK(cudaMemcpyAsync(e->h_scal, e->d_scal, sizeof(int32_t) * SC_COUNT, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    fold_spans(e);
    e->stats.d2h_bytes += sizeof(int32_t) * SC_COUNT;
    int rc = device_error(e);
    if (rc < 0) return rc;
    const int cnt = e->h_scal[SC_NEWC];
    if (cnt > cap) return fail(e, SW_E_ARG, "decide_fame: %d new consensus rounds do not fit cap=%d", cnt, cap);
    if (cnt > 0) {
        CK(cudaMemcpyAsync(e->h_newc, e->d_newc, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        memcpy(new_c_out, e->h_newc, sizeof(int32_t) * cnt);
        e->stats.d2h_bytes += sizeof(int32_t) * cnt;
    }
    return cnt;
}

int sw_find_order(sw_engine *e, const int32_t *new_c, int n) {
    if (!e || n < 0 || (n > 0 && !new_c)) return fail(e, SW_E_ARG, "bad argument");
    if (n == 0) return 0;
    CK(cudaSetDevice(e->device));
    std::vector<int32_t> rs(new_c, new_c + n);
    std::sort(rs.begin(), rs.end());                                  // sorted(new_c), swirld.py:283
    for (int r : rs) if (r < 0 || r >= e->Rcap) return fail(e, SW_E_KEY, "find_order: unknown round %d", r);
    if (n > e->seg_cap) {
        int nc = std::max(n, std::max(64, 2 * e->seg_cap));
        for (void *p : {(void *)e->d_seg_start, (void *)e->d_seg_fw, (void *)e->d_seg_nf, (void *)e->d_seg_white, (void *)e->d_rounds_in, (void *)e->d_plan})
            if (p) cudaFree(p);
        CK(dalloc(&e->d_seg_start, (size_t)nc + 1)); CK(dalloc(&e->d_seg_fw, (size_t)nc * 64));
        CK(dalloc(&e->d_seg_nf, (size_t)nc)); CK(dalloc(&e->d_seg_white, (size_t)nc * 64));
        CK(dalloc(&e->d_rounds_in, (size_t)nc)); CK(dalloc(&e->d_plan, (size_t)nc * 64 * 8));
        e->seg_cap = nc;
    }
    CK(cudaMemcpyAsync(e->d_rounds_in, rs.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, e->stream));
    OrderParams P{};
    P.M = e->M; P.Rcap = e->Rcap; P.nrounds = n; P.rounds = e->d_rounds_in; P.W = e->d_W; P.famous = e->d_famous;
    P.row = e->d_row; P.p0 = e->d_p0; P.creator = e->d_creator; P.seq = e->d_seq; P.t = e->d_t; P.sig = e->d_sig;
    P.stake = e->d_stake; P.tot = e->tot; P.lastord = e->d_lastord; P.batch_ev = e->d_batch_ev; P.batch_seg = e->d_batch_seg;
    P.seg_start = e->d_seg_start; P.seg_fw = e->d_seg_fw; P.seg_nf = e->d_seg_nf; P.seg_white = e->d_seg_white;
    P.ts = e->d_ts; P.key = e->d_key; P.perm = e->d_perm; P.tx = e->d_tx; P.idx = e->d_idx; P.tx_base = e->n_tx; P.scal = e->d_scal;
    P.plan = e->d_plan; P.plan_stride = e->seg_cap * 64;
    cudaEvent_t a = get_event(e), b = get_event(e);
    cudaEventRecord(a, e->stream);
    k_order_rounds<<<n, 1024, 0, e->stream>>>(P);
    k_order_cuts<<<1, 64, 0, e->stream>>>(P);
    k_order_list<<<std::max(1, std::min(4 * e->n_sm, (n * 64 + 255) / 256)), 256, 0, e->stream>>>(P);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(e->h_scal, e->d_scal, sizeof(int32_t) * SC_COUNT, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));      // rs (host vector) was consumed by the copy above
    e->stats.kernel_launches += 3;
    e->stats.h2d_bytes += sizeof(int32_t) * n;
    e->stats.d2h_bytes += sizeof(int32_t) * SC_COUNT;
    int rc = device_error(e);
    const int nbatch = e->h_scal[SC_BATCH];
    if (rc == 0 && nbatch > 0) {
        const int wpb = 8;
        k_order_times<<<(nbatch + wpb - 1) / wpb, wpb * 32, 0, e->stream>>>(P, nbatch);
        k_order_sort<<<n, 1024, 0, e->stream>>>(P);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(e->h_scal, e->d_scal, sizeof(int32_t) * SC_COUNT, cudaMemcpyDeviceToHost, e->stream));
        e->stats.kernel_launches += 2;
    }
    cudaEventRecord(b, e->stream);
    e->spans.push_back(TimedSpan{a, b, 2});
    CK(cudaStreamSynchronize(e->stream));
    fold_spans(e);
    if (rc == 0) rc = device_error(e);
    if (rc < 0) return rc;
    e->n_tx += nbatch;
    return nbatch;
}

int sw_n_events(const sw_engine *e) { return e ? e->n_events : SW_E_ARG; }
int sw_n_divided(const sw_engine *e) { return e ? e->n_divided : SW_E_ARG; }
int sw_n_transactions(const sw_engine *e) { return e ? e->n_tx : SW_E_ARG; }

int sw_sync(sw_engine *e) {
    if (!e) return SW_E_ARG;
    CK(cudaSetDevice(e->device));
    if (wait_appends(e, -1) < 0) return SW_E_CUDA;
    CK(cudaMemcpyAsync(e->h_scal, e->d_scal, sizeof(int32_t) * SC_COUNT, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    fold_spans(e);
    return device_error(e);
}

int sw_max_round(sw_engine *e) {
    int rc = sw_sync(e);
    if (rc < 0) return rc;
    return e->h_scal[SC_MAX_ROUND];
}

int sw_stats(sw_engine *e, sw_stats_t *out) {
    if (!e || !out) return SW_E_ARG;
    int rc = sw_sync(e);
    *out = e->stats;
    return rc;
}

#define GETTER(NAME, TYPE, SRC, LIMIT, WIDTH)                                                        \
    int NAME(sw_engine *e, int first, int n, TYPE *out) {                                            \
        if (!e || first < 0 || n < 0 || (n > 0 && !out)) return fail(e, SW_E_ARG, "bad argument");  \
        if (first + n > (LIMIT)) return fail(e, SW_E_KEY, #NAME ": [%d,%d) out of range", first, first + n); \
        if (n == 0) return SW_OK;                                                                    \
        CK(cudaSetDevice(e->device));                                                                \
        if (wait_appends(e, -1) < 0) return SW_E_CUDA;                                               \
        CK(cudaMemcpyAsync(out, (SRC) + (size_t)first * (WIDTH), sizeof(TYPE) * (size_t)n * (WIDTH), \
                           cudaMemcpyDeviceToHost, e->stream));                                     \
        CK(cudaStreamSynchronize(e->stream));                                                        \
        e->stats.d2h_bytes += sizeof(TYPE) * (size_t)n * (WIDTH);                                    \
        return SW_OK;                                                                                \
    }

GETTER(sw_get_round, int32_t, e->d_round, e->n_divided, 1)
GETTER(sw_get_witness_flags, uint8_t, e->d_wit, e->n_divided, 1)
GETTER(sw_get_famous, int8_t, e->d_famous_ev, e->n_events, 1)
GETTER(sw_get_can_see, int32_t, e->d_row, e->n_divided, e->M)
GETTER(sw_get_witness_table, int32_t, e->d_W, e->Rcap, e->M)
GETTER(sw_get_transactions, int32_t, e->d_tx, e->n_tx, 1)
GETTER(sw_get_idx, int32_t, e->d_idx, e->n_events, 1)

int sw_get_height(sw_engine *e, int first, int n, int32_t *out) {
    if (!e || first < 0 || n < 0 || (n > 0 && !out)) return fail(e, SW_E_ARG, "bad argument");
    if (first + n > e->n_events) return fail(e, SW_E_KEY, "sw_get_height: out of range");
    memcpy(out, e->h_height + first, sizeof(int32_t) * n);
    return SW_OK;
}

int sw_get_consensus(sw_engine *e, int32_t *out, int cap) {
    if (!e || cap < 0) return fail(e, SW_E_ARG, "bad argument");
    int mr = sw_max_round(e);
    if (mr < -1) return mr;
    std::vector<uint8_t> flags((size_t)mr + 2);
    if (mr >= 0) {
        CK(cudaMemcpyAsync(flags.data(), e->d_consensus, (size_t)mr + 1, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        e->stats.d2h_bytes += mr + 1;
    }
    int cnt = 0;
    for (int r = 0; r <= mr; r++)
        if (flags[r]) { if (cnt < cap) out[cnt] = r; cnt++; }
    return cnt;
}

int sw_debug_counters(sw_engine *e, int64_t *out16, int clear) {
    if (!e || !out16) return SW_E_ARG;
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(out16, e->d_dbg, sizeof(long long) * 16, cudaMemcpyDeviceToHost));
    if (clear) CK(cudaMemset(e->d_dbg, 0, sizeof(long long) * 40));
    return SW_OK;
}

int sw_flush_l2(sw_engine *e, int64_t bytes) {
    if (!e || bytes <= 0) return SW_E_ARG;
    CK(cudaSetDevice(e->device));
    if ((size_t)bytes > e->flush_bytes) {
        if (e->d_flush) cudaFree(e->d_flush);
        e->d_flush = nullptr;
        CK(cudaMalloc(&e->d_flush, (size_t)bytes));
        e->flush_bytes = (size_t)bytes;
    }
    CK(cudaMemsetAsync(e->d_flush, 0x5a, (size_t)bytes, e->stream));
    return SW_OK;
}

}  // extern "C"
