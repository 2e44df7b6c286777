// The C ABI of include/mofreak_b200.h: context, device memory, launch sequences.  No torch, no CPU fallback.
//
// Device memory of a context (W x H frames, store of N frames, batches of up to F pairs):
//   frame store            N * W*H                u8     the ring that keeps t and t-5 resident
//   per lane (x4: consecutive launch sequences rotate through complete buffer sets, each on its own streams)
//     maps img/S/seen/status 4 * F * pyr_px       u8     layout in geom.h (pyr_px ~ 1.9 W*H)
//     integral             F * (W+1)(H+1)         i32
//     hit masks            F * rows * ceil(W/64)  u64    detector hits, one word per (pair, layer row, 64-column tile)
//     candidates           cand_cap * 30 B               ordered list + keypoint records + scans, work queues, tie lists
//   rows                   row_cap * 32 B                accumulated mfk_row output
//   FREAK tables           8.4 MB                        64 x 256 x 43 receptive fields (float x, y, sigma)
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/mofreak_b200.h"
#include "../../include/mofreak_b200_debug.h"
#include "kernels.h"

using namespace mfk;

static_assert(sizeof(mfk_row) == 32, "mfk_row must be 32 bytes");

namespace {

thread_local std::string g_create_error;

enum ProfId { P_FRONT, P_ABSDIFF, P_PYRAMID, P_SCORES, P_SCAN, P_FILL, P_INTEGRAL, P_PHASE_A, P_PHASE_B, P_COMPACT, P_DESCRIBE,
              P_BOW, P_MIP, P_TEXT, P_COUNT };
const char* const kProfNames[P_COUNT] = {"front", "absdiff", "pyramid", "scores", "scan", "fill_candidates", "integral", "phase_a",
                                         "phase_b", "compact", "describe", "bow_assign", "mip_only", "text"};
// kernels launched per call of each launcher family (for gpu_launches accounting)
struct ProfSpan { int id; cudaEvent_t a, b; };

}  // namespace

enum { DEFAULT_LANES = 4 };

struct mfk_ctx {
    int device = 0, W = 0, H = 0, store_frames = 0, max_pairs = 0;
    int cand_cap_pair = 0, row_cap_pair = 0;
    cudaStream_t stream = nullptr;        // all kernels, row fetches
    cudaStream_t copy_stream = nullptr;   // frame-ring uploads, so that H2D of the next frames overlaps the kernels of these
    cudaEvent_t upload_done = nullptr;
    bool upload_pending = false;
    enum { MAX_LANES = 8 };
    // slots a queued launch sequence still reads, as a few [lo, hi] intervals (frame rings: one interval per stream), so that
    // an upload only waits for the kernels that really read the slots it overwrites
    struct ReadRec {
        enum { MAX_IV = 16 };
        int lo[MAX_IV], hi[MAX_IV], n_iv = 0;
        cudaEvent_t ev[MAX_LANES] = {}; bool on[MAX_LANES] = {}; bool valid = false;
        bool overlaps(int a, int b) const { for (int i = 0; i < n_iv; i++) if (lo[i] <= b && hi[i] >= a) return true; return false; }
    };
    ReadRec readers[8];                   // which slot ranges recent launch sequences still read
    int reader_next = 0;
    // the mirror image: which slots recent uploads fill, each with its own event on the copy stream, so that a launch sequence
    // waits only for the uploads that carry ITS frames (a caller may queue all uploads of a step first and the extractions
    // after them without serialising the first extraction behind the last upload).  Records leave the ring in upload order;
    // `writers_folded` is the event of the newest record that left: the copy stream is in order, so waiting for it covers
    // every upload that is no longer tracked.
    struct WriteRec { int lo = 0, hi = -1; cudaEvent_t ev = nullptr; bool valid = false; };
    enum { N_WRITERS = 256 };
    WriteRec writers[N_WRITERS];
    int writer_next = 0;
    cudaEvent_t writers_folded = nullptr;
    bool writers_folded_valid = false;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    // geometry template (F filled per batch)
    int lw[N_LAYERS], lh[N_LAYERS], rows = 0, pyr_px = 0;
    // device buffers
    uint8_t* store = nullptr;
    // Complete sets of per-batch buffers ("lanes"), each with its own streams: consecutive launch sequences
    // rotate through them, so the latency-bound sparse kernels of one batch run beside the issue-bound dense
    // kernels of the next.  Only the row accumulation (counters, rows_buf, pair_rows) is shared; it is handed from
    // batch to batch through `advanced` events.
    struct Lane {
        Maps maps{};
        int32_t* integral = nullptr;
        int32_t* integral_left = nullptr;
        unsigned long long* hitmask = nullptr;  // detector hits, one 64-bit word per (pair, layer row, 64-column tile)     // per image row: sums left of each 32-column strip
        CandBuf cb{};
        ScoreMapCache score_maps;
        int32_t* pairs = nullptr;                // this batch's pair descriptions: cur | prev | label, max_pairs each
        uint32_t* scan_scratch = nullptr;
        // stream: the lane's spine (small kernels, descriptor kernel; high priority).  dense_stream: k_front + k_scores, the two
        // kernels whose grids fill every SM for their whole duration (low priority); aux_stream: the integral image (low).
        cudaStream_t stream = nullptr, dense_stream = nullptr, aux_stream = nullptr;
        cudaEvent_t aux_fork = nullptr, aux_join = nullptr, advanced = nullptr, done = nullptr, prepared = nullptr, scored = nullptr;
        bool used = false;
        int seen_F = -1;                         // batch size the `seen` map was last cleared for (layer bases depend on it)
    };
    Lane lane[MAX_LANES];
    int n_lanes = DEFAULT_LANES, lanes_built = DEFAULT_LANES;
    bool reset_pending = false;           // mfk_reset_rows is applied in-line by the next launch sequence (or by settle())
    bool main_wait_pending = false;       // the next launch sequence orders its shared-row part after main_mark
    bool serial = false;
    bool prioritised = true;              // MFK_PRIORITY=0: one stream per lane (plus the integral side stream), no priorities
    bool early_losers = true;             // MFK_EARLY_LOSERS=0: every detector hit becomes a candidate (k_scores drops none)
    bool tail_overlap = false;            // MFK_TAIL_OVERLAP=1: descriptor kernels of consecutive batches may overlap
    bool early_compact = true;            // MFK_EARLY_COMPACT=0: the whole second compaction waits for the previous tail
    bool fused_front = true;              // MFK_FUSED_FRONT=0 forces the separate dense kernels (A/B measurements, wide frames use them anyway)
    uint64_t batches = 0;
    int last_lane = 0;
    uint32_t* counters = nullptr;
    cudaEvent_t main_mark = nullptr;
    int cand_cap = 0;
    mfk_row* rows_buf = nullptr;
    size_t row_cap = 0;
    uint32_t* pair_rows = nullptr;
    size_t pair_cap = 0, pairs_queued = 0;
    DeviceTables tb{};
    void* tb_alloc[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint8_t* codebook = nullptr;
    int K = 0;
    uint32_t* bow_counts = nullptr;
    int bow_counts_cap = 0;
    void* bow_tmp = nullptr;
    size_t bow_tmp_bytes = 0;
    // `.mofreak` text formatting: line lengths/offsets (+ scan scratch, + flag word), the text itself, staged rows
    uint32_t* text_off = nullptr;
    size_t text_off_cap = 0;
    char* text_buf = nullptr;
    size_t text_buf_cap = 0;
    mfk_row* text_rows = nullptr;
    size_t text_rows_cap = 0;
    uint8_t* bgr_stage = nullptr;         // BGR frames on their way into the gray store (mfk_upload_frames_bgr)
    size_t bgr_stage_bytes = 0;
    // last batch (debug access)
    Geom last{};
    bool have_last = false;
    // accounting
    uint64_t launches = 0;
    bool prof_on = false;
    std::vector<ProfSpan> spans;
    std::vector<cudaEvent_t> event_pool;
    uint32_t prof_launches[P_COUNT] = {};
    float prof_ms[P_COUNT] = {};
    std::string err;
};

namespace {

int fail(mfk_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CU(ctx, call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) return fail(ctx, MFK_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

cudaEvent_t take_event(mfk_ctx* c) {
    if (!c->event_pool.empty()) { cudaEvent_t e = c->event_pool.back(); c->event_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

struct Span {          // brackets one launcher call with events when profiling is on
    mfk_ctx* c; int id; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
    Span(mfk_ctx* c_, int id_, int kernels, cudaStream_t st_ = nullptr) : c(c_), id(id_), st(st_ ? st_ : c_->stream) {
        c->launches += kernels;
        c->prof_launches[id] += kernels;
        if (c->prof_on) { a = take_event(c); b = take_event(c); cudaEventRecord(a, st); }
    }
    ~Span() {
        if (a) { cudaEventRecord(b, st); c->spans.push_back(ProfSpan{id, a, b}); }
    }
};

void drain_spans(mfk_ctx* c) {
    for (const ProfSpan& s : c->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) c->prof_ms[s.id] += ms;
        c->event_pool.push_back(s.a);
        c->event_pool.push_back(s.b);
    }
    c->spans.clear();
}

Geom make_geom(const mfk_ctx* c, int F) {
    Geom g{};
    size_t below = 0;
    int row0 = 0;
    static const float scale[N_LAYERS] = {1.0f, 1.5f, 2.0f, 3.0f, 4.0f, 6.0f};
    for (int l = 0; l < N_LAYERS; l++) {
        g.L[l].w = c->lw[l]; g.L[l].h = c->lh[l]; g.L[l].px = c->lw[l] * c->lh[l];
        g.L[l].row0 = row0; g.L[l].base = (size_t)F * below;
        g.L[l].scale = scale[l];
        g.L[l].offset = (float)(0.5 * scale[l] - 0.5);
        below += (size_t)g.L[l].px; row0 += c->lh[l];
    }
    g.W = c->W; g.H = c->H; g.F = F; g.rows = c->rows; g.pyr_px = c->pyr_px;
    g.cand_cap = c->cand_cap; g.row_cap = (int)c->row_cap;
    return g;
}

template <class T> cudaError_t dalloc(T** p, size_t count) { return cudaMalloc((void**)p, count * sizeof(T)); }

int upload_tables(mfk_ctx* c) {
    const Tables& T = tables();
    const size_t npts = (size_t)FREAK_SCALES * FREAK_ORIENTATIONS * FREAK_POINTS;
    FreakPoint* pts; int32_t* ps; float* thr; OrientPair* op; UsedPair* up;
    CU(c, dalloc(&pts, npts)); c->tb_alloc[0] = pts;
    CU(c, dalloc(&ps, FREAK_SCALES)); c->tb_alloc[1] = ps;
    CU(c, dalloc(&thr, FREAK_SCALES)); c->tb_alloc[2] = thr;
    CU(c, dalloc(&op, FREAK_ORIENT_PAIRS)); c->tb_alloc[3] = op;
    CU(c, dalloc(&up, FREAK_USED_PAIRS)); c->tb_alloc[4] = up;
    CU(c, cudaMemcpy(pts, T.points, npts * sizeof(FreakPoint), cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(ps, T.pattern_size, sizeof T.pattern_size, cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(thr, T.scale_threshold, sizeof T.scale_threshold, cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(op, T.orient, sizeof T.orient, cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(up, T.used, sizeof T.used, cudaMemcpyHostToDevice));
    uint16_t* mn;
    CU(c, dalloc(&mn, 2 * 361)); c->tb_alloc[5] = mn;
    CU(c, cudaMemcpy(mn, T.mip_needed, sizeof T.mip_needed, cudaMemcpyHostToDevice));
    c->tb.mip_needed = mn; c->tb.n_mip_needed = T.n_mip_needed;
    {
        uint32_t plan[2 * 361];
        for (int i = 0; i < T.n_mip_needed; i++) {
            const uint32_t e = T.mip_needed[i], fr = e >> 15, dx = e & 31, dy = (e >> 5) & 31;
            plan[i] = describe_patch_offset((int)fr, (int)dx, (int)dy) | (16u * dx) << 10 | (16u * dy) << 19 | fr << 28;
        }
        uint32_t* mp;
        CU(c, dalloc(&mp, 2 * 361)); c->tb_alloc[6] = mp;
        CU(c, cudaMemcpy(mp, plan, sizeof(uint32_t) * T.n_mip_needed, cudaMemcpyHostToDevice));
        c->tb.mip_plan = mp;
    }
    c->tb.points = pts; c->tb.pattern_size = ps; c->tb.scale_threshold = thr; c->tb.orient = op; c->tb.used = up;
    return MFK_OK;
}

// the copy stream has just been given the upload of slots [lo, hi]: remember it with an event of its own
int note_upload(mfk_ctx* c, int lo, int hi) {
    mfk_ctx::WriteRec& w = c->writers[c->writer_next];
    c->writer_next = (c->writer_next + 1) % mfk_ctx::N_WRITERS;
    if (w.valid) { std::swap(w.ev, c->writers_folded); c->writers_folded_valid = true; }      // the oldest record leaves the ring
    CU(c, cudaEventRecord(w.ev, c->copy_stream));
    w.lo = lo; w.hi = hi; w.valid = true;
    return MFK_OK;
}

// orders everything queued on the lanes before whatever is queued on the main stream next
int join_lanes(mfk_ctx* c) {
    for (int i = 0; i < c->lanes_built; i++) {
        mfk_ctx::Lane& L = c->lane[i];
        if (!L.used) continue;
        CU(c, cudaEventRecord(L.done, L.stream));
        CU(c, cudaStreamWaitEvent(c->stream, L.done, 0));
    }
    return MFK_OK;
}

// join_lanes + apply a row reset that no launch sequence has consumed yet: the state every reader of the shared row
// accumulation (counters, rows_buf, pair_rows) starts from
int settle(mfk_ctx* c) {
    { const int rc = join_lanes(c); if (rc != MFK_OK) return rc; }
    if (c->reset_pending) {
        CU(c, cudaMemsetAsync(c->counters, 0, 4 * sizeof(uint32_t), c->stream));
        c->reset_pending = false;
    }
    return MFK_OK;
}

int sync_all(mfk_ctx* c) {
    CU(c, cudaStreamSynchronize(c->copy_stream));
    for (int i = 0; i < c->lanes_built; i++) {
        if (c->lane[i].aux_stream) CU(c, cudaStreamSynchronize(c->lane[i].aux_stream));
        if (c->lane[i].dense_stream) CU(c, cudaStreamSynchronize(c->lane[i].dense_stream));
        if (c->lane[i].stream) CU(c, cudaStreamSynchronize(c->lane[i].stream));
    }
    CU(c, cudaStreamSynchronize(c->stream));
    return MFK_OK;
}

int grow_pair_rows(mfk_ctx* c, size_t need) {
    if (need <= c->pair_cap) return MFK_OK;
    size_t cap = std::max(need, c->pair_cap * 2);
    cap = std::max<size_t>(cap, 1024);
    { const int rc = sync_all(c); if (rc != MFK_OK) return rc; }
    uint32_t* pr;
    CU(c, dalloc(&pr, cap));
    if (c->pairs_queued) CU(c, cudaMemcpy(pr, c->pair_rows, c->pairs_queued * 4, cudaMemcpyDeviceToDevice));
    cudaFree(c->pair_rows);
    c->pair_rows = pr; c->pair_cap = cap;
    return MFK_OK;
}

// one launch sequence over `F` pairs (host descriptions cur/prev/label), on the next lane; `first` is the index of its
// first pair among the pairs queued since the last row reset
int run_batch(mfk_ctx* c, const int32_t* cur, const int32_t* prev, const int32_t* label, size_t first, int F) {
    const Geom g = make_geom(c, F);
    const int li = (int)(c->batches % (uint64_t)c->n_lanes);
    mfk_ctx::Lane& L = c->lane[li];
    mfk_ctx::Lane& P = c->lane[c->last_lane];            // lane of the previous batch
    cudaStream_t st = L.stream;
    PairList pl{L.pairs, L.pairs + c->max_pairs, L.pairs + 2 * (size_t)c->max_pairs};
    c->launches += launch_set_pairs(st, cur, prev, label, F, c->store_frames, pl);
    // write-back marks only ever land on layers 1..5 (they come from scans issued from the layer below), so layer 0's
    // share of the map stays zero as long as the lane keeps its geometry
    if (L.seen_F == F) CU(c, cudaMemsetAsync(L.maps.seen + g.L[1].base, 0, (size_t)F * g.pyr_px - g.L[1].base, st));
    else { CU(c, cudaMemsetAsync(L.maps.seen, 0, (size_t)c->max_pairs * c->pyr_px, st)); L.seen_F = F; }
    CU(c, cudaMemsetAsync(L.cb.pend_count, 0, (size_t)F * N_LAYERS * sizeof(uint32_t), st));
    CU(c, cudaMemsetAsync(L.cb.q_count, 0, 4 * sizeof(uint32_t), st));
    // dense front: one fused kernel (difference, pyramid, row sums) for frames up to 2048 wide, the separate kernels beyond.
    // It and the score kernel run on the lane's LOW-priority stream: their grids fill every SM for their whole duration, and
    // at equal priority the block scheduler hands the small dependent kernels of the other launch sequences (scans,
    // compactions, NMS phases: each a few blocks that wait on the previous one) their SM slots only when such a grid has been
    // dispatched completely - measured with mfk_profile_timeline: 1.1 ms for the five tiny kernels of the second compaction,
    // which run in 0.08 ms on an idle GPU, on the path that serialises the launch sequences.
    const bool fused = c->fused_front && front_supported(g);
    cudaStream_t ds = (c->serial || !L.dense_stream) ? st : L.dense_stream;
    if (ds != st) { CU(c, cudaEventRecord(L.prepared, st)); CU(c, cudaStreamWaitEvent(ds, L.prepared, 0)); }
    if (fused) { Span s(c, P_FRONT, 1, ds); CU(c, launch_front(ds, g, c->store, c->store_frames, pl, L.maps.img, L.integral_left)); }
    else { Span s(c, P_ABSDIFF, 1, ds);   launch_absdiff(ds, g, c->store, c->store_frames, pl, L.maps.img); }
    // the integral image only needs layer 0: fork it onto the side stream, join before the descriptor kernel
    if (c->serial) {
        Span s(c, P_INTEGRAL, fused ? 1 : 2, st);  launch_integral(st, g, L.maps.img, L.integral, L.integral_left, fused);
    } else {
        CU(c, cudaEventRecord(L.aux_fork, ds));
        CU(c, cudaStreamWaitEvent(L.aux_stream, L.aux_fork, 0));
        { Span s(c, P_INTEGRAL, fused ? 1 : 2, L.aux_stream);  launch_integral(L.aux_stream, g, L.maps.img, L.integral, L.integral_left, fused); }
        CU(c, cudaEventRecord(L.aux_join, L.aux_stream));
    }
    if (!fused) { Span s(c, P_PYRAMID, 5, ds);   launch_pyramid(ds, g, L.maps.img); }
    { Span s(c, P_SCORES, 1, ds);    launch_scores(ds, g, L.maps.img, L.maps.S, c->early_losers ? L.maps.status : nullptr, L.hitmask, L.score_maps); }
    if (ds != st) { CU(c, cudaEventRecord(L.scored, ds)); CU(c, cudaStreamWaitEvent(st, L.scored, 0)); }
    { Span s(c, P_SCAN, 4, st);      launch_row_counts(st, g, L.hitmask, L.cb.row_count);
                                     launch_exclusive_scan(st, L.cb.row_count, F * g.rows, nullptr, L.scan_scratch); }
    { Span s(c, P_FILL, 1, st);      launch_fill_candidates(st, g, L.hitmask, L.cb); }
    { Span s(c, P_PHASE_A, 4, st);   launch_phase_a(st, g, L.maps, L.cb, c->tb); }
    { Span s(c, P_PHASE_B, 1, st);   launch_phase_b(st, g, L.maps, L.cb, c->tb); }
    // second ordered compaction, lane-private part: keep flags of the batch's candidates and their exclusive scan
    auto keep_scan = [&]() {
        Span s(c, P_COMPACT, 4, st);
        launch_keep_flags(st, g, L.cb);
        launch_exclusive_scan(st, L.cb.keep_scan, c->cand_cap, L.cb.live, L.scan_scratch);
    };
    if (c->early_compact) keep_scan();
    // From here on batches run one after the other.  Only the advance step of k_scatter_rows (this batch's base in the shared row accumulation, the
    // counters) and the per-pair row counts need that; the descriptor kernel is lane-private, but letting those of several
    // batches overlap measured 1-2 % slower (they gather from per-batch candidate arrays and integral images, and together
    // thrash L2), so the tail waits for the previous batch's.  Also ordered here: whatever the caller queued on the
    // main stream before this call (bag-of-words reads of the rows, timers), and a pending row reset.
    if (c->batches > 0 && c->last_lane != li) CU(c, cudaStreamWaitEvent(st, P.advanced, 0));
    if (c->main_wait_pending) { CU(c, cudaStreamWaitEvent(st, c->main_mark, 0)); c->main_wait_pending = false; }
    if (c->reset_pending) { CU(c, cudaMemsetAsync(c->counters, 0, 4 * sizeof(uint32_t), st)); c->reset_pending = false; }
    if (!c->early_compact) keep_scan();
    // kept keypoints -> row slots of this batch, per-pair row counts, and this batch's base in the shared accumulation
    { Span s(c, P_COMPACT, 1, st);   launch_scatter_rows(st, g, L.cb, c->pair_rows + first); }
    if (c->tail_overlap) CU(c, cudaEventRecord(L.advanced, st));       // the next batch's tail may start under this descriptor kernel
    if (!c->serial) CU(c, cudaStreamWaitEvent(st, L.aux_join, 0));
    { Span s(c, P_DESCRIBE, 1, st);
      launch_describe(st, g, c->store, c->store_frames, pl, L.maps.img, L.integral, L.cb, c->tb, c->rows_buf);
    }
    if (!c->tail_overlap) CU(c, cudaEventRecord(L.advanced, st));
    CU(c, cudaGetLastError());
    L.used = true;
    c->last = g; c->have_last = true; c->last_lane = li;
    c->batches++;
    return MFK_OK;
}

void destroy(mfk_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) sync_all(c);
    drain_spans(c);
    for (cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
    cudaFree(c->store);
    for (auto& L : c->lane) {
        cudaFree(L.maps.img); cudaFree(L.maps.S); cudaFree(L.maps.seen); cudaFree(L.maps.status); cudaFree(L.integral); cudaFree(L.integral_left); cudaFree(L.hitmask);
        cudaFree(L.cb.row_count); cudaFree(L.cb.xy); cudaFree(L.cb.meta); cudaFree(L.cb.kx); cudaFree(L.cb.ky); cudaFree(L.cb.ksize);
        cudaFree(L.cb.kresp); cudaFree(L.cb.flags); cudaFree(L.cb.scale_idx); cudaFree(L.cb.keep_scan); cudaFree(L.cb.row_src);
        cudaFree(L.cb.pend_count); cudaFree(L.cb.pend_list); cudaFree(L.cb.live);
        cudaFree(L.cb.q_count); cudaFree(L.cb.q0); cudaFree(L.cb.q1); cudaFree(L.cb.q2); cudaFree(L.cb.q1_up); cudaFree(L.cb.q2_scan); cudaFree(L.scan_scratch); cudaFree(L.pairs);
        for (cudaEvent_t e : {L.aux_fork, L.aux_join, L.advanced, L.done, L.prepared, L.scored}) if (e) cudaEventDestroy(e);
        if (L.aux_stream) cudaStreamDestroy(L.aux_stream);
        if (L.dense_stream) cudaStreamDestroy(L.dense_stream);
        if (L.stream) cudaStreamDestroy(L.stream);
    }
    cudaFree(c->counters);
    cudaFree(c->rows_buf); cudaFree(c->pair_rows);
    for (void* p : c->tb_alloc) cudaFree(p);
    cudaFree(c->codebook); cudaFree(c->bow_counts); cudaFree(c->bow_tmp);
    cudaFree(c->text_off); cudaFree(c->text_buf); cudaFree(c->text_rows); cudaFree(c->bgr_stage);
    if (c->t0) cudaEventDestroy(c->t0);
    if (c->t1) cudaEventDestroy(c->t1);
    if (c->upload_done) cudaEventDestroy(c->upload_done);
    if (c->main_mark) cudaEventDestroy(c->main_mark);
    for (auto& r : c->readers) for (cudaEvent_t e : r.ev) if (e) cudaEventDestroy(e);
    for (auto& wr : c->writers) if (wr.ev) cudaEventDestroy(wr.ev);
    if (c->writers_folded) cudaEventDestroy(c->writers_folded);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int ensure_bow_tmp(mfk_ctx* c, size_t bytes) {
    if (bytes <= c->bow_tmp_bytes) return MFK_OK;
    CU(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->bow_tmp); c->bow_tmp = nullptr; c->bow_tmp_bytes = 0;
    CU(c, cudaMalloc(&c->bow_tmp, bytes));
    c->bow_tmp_bytes = bytes;
    return MFK_OK;
}

// `.mofreak` text of n device rows on the main stream, copied to the caller's HOST buffer
int format_device_rows(mfk_ctx* c, const mfk_row* d_rows, size_t n, char* text, size_t cap, size_t* n_bytes) {
    if (n_bytes) *n_bytes = 0;
    if (n == 0) return MFK_OK;
    if (n > 0xfffffff0u / TEXT_ROW_MAX) return fail(c, MFK_ERR_ARG, "%zu rows exceed the 4 GiB text limit of one call", n);
    // layout of text_off: [0, n] lengths -> offsets (slot n = total) | scan scratch | flag word
    const size_t scratch_words = (n + 1) / 2048 + 8;
    const size_t need = n + 1 + scratch_words + 1;
    if (need > c->text_off_cap) {
        CU(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->text_off); c->text_off = nullptr; c->text_off_cap = 0;
        CU(c, dalloc(&c->text_off, need + need / 4));
        c->text_off_cap = need + need / 4;
    }
    uint32_t* off = c->text_off;
    uint32_t* scratch = off + n + 1;
    uint32_t* flags = scratch + scratch_words;
    CU(c, cudaMemsetAsync(flags, 0, 4, c->stream));
    uint32_t h[2] = {0, 0};      // total bytes, flags
    { Span s(c, P_TEXT, 4);
      launch_text_lengths(c->stream, d_rows, (uint32_t)n, off, flags);
      launch_exclusive_scan(c->stream, off, (int)(n + 1), nullptr, scratch); }
    CU(c, cudaMemcpyAsync(&h[0], off + n, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(&h[1], flags, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (h[1]) return fail(c, MFK_ERR_RANGE, "a row holds a float that \"%%g\" prints in exponent notation; format these rows on the host");
    if (n_bytes) *n_bytes = h[0];
    if (h[0] > cap || !text) return fail(c, MFK_ERR_CAPACITY, "%u bytes of text do not fit the caller's buffer of %zu", h[0], cap);
    if (h[0] > c->text_buf_cap) {
        cudaFree(c->text_buf); c->text_buf = nullptr; c->text_buf_cap = 0;
        const size_t want = (size_t)h[0] + h[0] / 4 + 256;
        CU(c, cudaMalloc((void**)&c->text_buf, want));
        c->text_buf_cap = want;
    }
    { Span s(c, P_TEXT, 1); launch_text_write(c->stream, d_rows, (uint32_t)n, off, c->text_buf); }
    CU(c, cudaMemcpyAsync(text, c->text_buf, h[0], cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaGetLastError());
    return MFK_OK;
}

}  // namespace

extern "C" {

const char* mfk_version(void) { return "mofreak_b200 0.1 (sm_100a)"; }

int mfk_freak_scale_index(float kp_size) { return scale_index_from_table(kp_size); }

int mfk_create(mfk_ctx** out, int device, int width, int height, int store_frames, int max_pairs) {
    return mfk_create_ex(out, device, width, height, store_frames, max_pairs, 0, 0);
}

int mfk_create_ex(mfk_ctx** out, int device, int width, int height, int store_frames, int max_pairs, int cand_cap_pair,
                  int row_cap_pair) {
    if (!out) return fail(nullptr, MFK_ERR_ARG, "out is NULL");
    *out = nullptr;
    // the top pyramid layer must still hold a 7x7 neighbourhood, and coordinates are packed in 16 bits
    if (width < 48 || height < 48 || width > 8192 || height > 8192)
        return fail(nullptr, MFK_ERR_ARG, "frame size %dx%d outside 48..8192", width, height);
    if (store_frames < MFK_FRAME_GAP + 1 || max_pairs < 1) return fail(nullptr, MFK_ERR_ARG, "store_frames >= 6 and max_pairs >= 1 required");
    if (max_pairs > 65535) return fail(nullptr, MFK_ERR_ARG, "max_pairs <= 65535 (pairs index a grid dimension); longer requests are split into launch sequences anyway");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, MFK_ERR_CUDA, "no usable CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, MFK_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    mfk_ctx* c = new mfk_ctx();
    c->device = device; c->W = width; c->H = height; c->store_frames = store_frames; c->max_pairs = max_pairs;
#define CUC(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            fail(nullptr, MFK_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__));                    \
            destroy(c);                                                                                 \
            return MFK_ERR_CUDA;                                                                        \
        }                                                                                               \
    } while (0)
    CUC(cudaSetDevice(device));
    CUC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUC(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CUC(cudaEventCreateWithFlags(&c->upload_done, cudaEventDisableTiming));
    CUC(cudaEventCreateWithFlags(&c->main_mark, cudaEventDisableTiming));
    // tuning knob (documented in the header at mfk_create): number of buffer sets, 1..4
    if (const char* e = getenv("MFK_LANES")) c->lanes_built = std::max(1, std::min((int)mfk_ctx::MAX_LANES, atoi(e)));
    c->n_lanes = c->lanes_built;
    if (const char* e = getenv("MFK_FUSED_FRONT")) c->fused_front = atoi(e) != 0;
    if (const char* e = getenv("MFK_PRIORITY")) c->prioritised = atoi(e) != 0;
    if (const char* e = getenv("MFK_EARLY_COMPACT")) c->early_compact = atoi(e) != 0;
    if (const char* e = getenv("MFK_TAIL_OVERLAP")) c->tail_overlap = atoi(e) != 0;
    if (const char* e = getenv("MFK_EARLY_LOSERS")) c->early_losers = atoi(e) != 0;
    int prio_low = 0, prio_high = 0;
    CUC(cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high));
    for (int li = 0; li < c->lanes_built; li++) {
        mfk_ctx::Lane& L = c->lane[li];
        if (c->prioritised) {
            CUC(cudaStreamCreateWithPriority(&L.stream, cudaStreamNonBlocking, prio_high));
            CUC(cudaStreamCreateWithPriority(&L.dense_stream, cudaStreamNonBlocking, prio_low));
            CUC(cudaStreamCreateWithPriority(&L.aux_stream, cudaStreamNonBlocking, prio_low));
        } else {
            CUC(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
            CUC(cudaStreamCreateWithFlags(&L.aux_stream, cudaStreamNonBlocking));
        }
        CUC(cudaEventCreateWithFlags(&L.prepared, cudaEventDisableTiming));
        CUC(cudaEventCreateWithFlags(&L.scored, cudaEventDisableTiming));
        CUC(cudaEventCreateWithFlags(&L.aux_fork, cudaEventDisableTiming));
        CUC(cudaEventCreateWithFlags(&L.aux_join, cudaEventDisableTiming));
        CUC(cudaEventCreateWithFlags(&L.advanced, cudaEventDisableTiming));
        CUC(cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming));
    }
    for (auto& r : c->readers) for (auto& e : r.ev) CUC(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& wr : c->writers) CUC(cudaEventCreateWithFlags(&wr.ev, cudaEventDisableTiming));
    CUC(cudaEventCreateWithFlags(&c->writers_folded, cudaEventDisableTiming));
    CUC(cudaEventCreate(&c->t0));
    CUC(cudaEventCreate(&c->t1));
    // pyramid geometry, brisk.cpp:1660, 1666
    c->lw[0] = width; c->lh[0] = height;
    c->lw[1] = 2 * (width / 3); c->lh[1] = 2 * (height / 3);
    for (int l = 2; l < N_LAYERS; l++) { c->lw[l] = c->lw[l - 2] / 2; c->lh[l] = c->lh[l - 2] / 2; }
    for (int l = 0; l < N_LAYERS; l++) { c->rows += c->lh[l]; c->pyr_px += c->lw[l] * c->lh[l]; }
    c->cand_cap_pair = cand_cap_pair > 0 ? cand_cap_pair : std::max(4096, c->pyr_px / 6);
    c->row_cap_pair = row_cap_pair > 0 ? row_cap_pair : std::max(1024, c->pyr_px / 48);
    const size_t F = (size_t)max_pairs;
    const size_t cand_cap = std::min<size_t>(F * c->cand_cap_pair, (size_t)1 << 30);
    c->cand_cap = (int)cand_cap;
    c->row_cap = F * c->row_cap_pair * 4;       // several batches may accumulate before a fetch
    const size_t px = (size_t)width * height;
    CUC(dalloc(&c->store, (size_t)store_frames * px + 64));
    // slack: register windows (brisk_device.cuh Window4 / load5) read up to 4 rows + 8 bytes past a layer's end
    const size_t slack = (size_t)4 * width + 64;
    const size_t scan_n = std::max<size_t>(cand_cap, F * c->rows);
    CUC(dalloc(&c->counters, 8));
    CUC(cudaMemset(c->counters, 0, 8 * sizeof(uint32_t)));
    // a context that only ever sees one launch sequence per call (max_pairs >= the caller's batches) still gets all
    // lanes: consecutive calls rotate through them as well
    for (int li = 0; li < c->lanes_built; li++) {
        mfk_ctx::Lane& L = c->lane[li];
        CUC(dalloc(&L.maps.img, F * c->pyr_px + slack));
        CUC(dalloc(&L.maps.S, F * c->pyr_px + slack));
        CUC(dalloc(&L.maps.seen, F * c->pyr_px + slack));
        CUC(dalloc(&L.maps.status, F * c->pyr_px + slack));
        CUC(dalloc(&L.integral, F * (size_t)(width + 1) * (height + 1)));
        CUC(dalloc(&L.integral_left, F * (size_t)height * ((width + 31) / 32)));
        CUC(dalloc(&L.hitmask, F * (size_t)c->rows * score_mask_words(width)));
        CUC(dalloc(&L.cb.row_count, F * c->rows + 2));
        CUC(dalloc(&L.cb.xy, cand_cap)); CUC(dalloc(&L.cb.meta, cand_cap));
        CUC(dalloc(&L.cb.kx, cand_cap)); CUC(dalloc(&L.cb.ky, cand_cap));
        CUC(dalloc(&L.cb.ksize, cand_cap)); CUC(dalloc(&L.cb.kresp, cand_cap));
        CUC(dalloc(&L.cb.flags, cand_cap)); CUC(dalloc(&L.cb.scale_idx, cand_cap));
        CUC(dalloc(&L.cb.keep_scan, cand_cap + 2)); CUC(dalloc(&L.cb.row_src, cand_cap));
        CUC(dalloc(&L.cb.pend_count, F * N_LAYERS));
        CUC(dalloc(&L.cb.pend_list, F * N_LAYERS * PEND_CAP));
        L.cb.q_cap = (uint32_t)std::max<size_t>(cand_cap / 4, 4096);
        CUC(dalloc(&L.cb.q_count, 4));
        CUC(dalloc(&L.cb.q0, L.cb.q_cap)); CUC(dalloc(&L.cb.q1, L.cb.q_cap)); CUC(dalloc(&L.cb.q2, L.cb.q_cap));
        CUC(dalloc(&L.cb.q1_up, (size_t)3 * L.cb.q_cap)); CUC(dalloc(&L.cb.q2_scan, (size_t)6 * L.cb.q_cap));
        CUC(dalloc(&L.cb.live, 4));
        CUC(cudaMemset(L.cb.live, 0, 4 * sizeof(uint32_t)));
        L.cb.counters = c->counters;
        CUC(dalloc(&L.scan_scratch, scan_n / 2048 + 8));
        CUC(dalloc(&L.pairs, 3 * F));
    }
    CUC(dalloc(&c->rows_buf, c->row_cap));
    if (upload_tables(c) != MFK_OK) { g_create_error = c->err; destroy(c); return MFK_ERR_CUDA; }
    if (grow_pair_rows(c, (size_t)max_pairs) != MFK_OK) { g_create_error = c->err; destroy(c); return MFK_ERR_CUDA; }
#undef CUC
    *out = c;
    return MFK_OK;
}

void mfk_destroy(mfk_ctx* ctx) { destroy(ctx); }

const char* mfk_last_error(const mfk_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int mfk_sync(mfk_ctx* c) {
    if (!c) return MFK_ERR_ARG;
    return sync_all(c);
}

int mfk_host_alloc(void** ptr, size_t bytes, int write_combined) {
    if (!ptr || bytes == 0) return fail(nullptr, MFK_ERR_ARG, "bad argument to mfk_host_alloc");
    *ptr = nullptr;
    const cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0));
    if (e != cudaSuccess) return fail(nullptr, MFK_ERR_NOMEM, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    return MFK_OK;
}

int mfk_host_free(void* ptr) {
    if (!ptr) return MFK_OK;
    return cudaFreeHost(ptr) == cudaSuccess ? MFK_OK : MFK_ERR_CUDA;
}

int mfk_upload_frames(mfk_ctx* c, const uint8_t* gray, size_t stride, int n_frames, int slot0) {
    if (!c || !gray || n_frames < 0 || slot0 < 0) return fail(c, MFK_ERR_ARG, "bad argument to mfk_upload_frames");
    if (stride < (size_t)c->W) return fail(c, MFK_ERR_ARG, "stride %zu < width %d", stride, c->W);
    if (n_frames > c->store_frames) return fail(c, MFK_ERR_ARG, "%d frames do not fit a store of %d", n_frames, c->store_frames);
    CU(c, cudaSetDevice(c->device));
    const size_t px = (size_t)c->W * c->H;
    int done = 0;
    while (done < n_frames) {
        const int slot = (slot0 + done) % c->store_frames;
        const int run = std::min(n_frames - done, c->store_frames - slot);     // contiguous slots before the wrap
        // slots still being read by queued launch sequences must not be overwritten under them
        for (auto& r : c->readers)
            if (r.valid && r.overlaps(slot, slot + run - 1))
                for (int k = 0; k < mfk_ctx::MAX_LANES; k++) if (r.on[k]) CU(c, cudaStreamWaitEvent(c->copy_stream, r.ev[k], 0));
        const uint8_t* src = gray + (size_t)done * stride * c->H;
        if (stride == (size_t)c->W)
            CU(c, cudaMemcpyAsync(c->store + (size_t)slot * px, src, (size_t)run * px, cudaMemcpyHostToDevice, c->copy_stream));
        else
            CU(c, cudaMemcpy2DAsync(c->store + (size_t)slot * px, c->W, src, stride, c->W, (size_t)run * c->H,
                                    cudaMemcpyHostToDevice, c->copy_stream));
        { const int rc = note_upload(c, slot, slot + run - 1); if (rc != MFK_OK) return rc; }
        done += run;
    }
    CU(c, cudaEventRecord(c->upload_done, c->copy_stream));
    c->upload_pending = true;
    return MFK_OK;
}

int mfk_upload_frames_bgr(mfk_ctx* c, const uint8_t* bgr, size_t stride, int n_frames, int slot0, int coeff_bits) {
    if (!c || !bgr || n_frames < 0 || slot0 < 0 || (coeff_bits != 14 && coeff_bits != 15))
        return fail(c, MFK_ERR_ARG, "bad argument to mfk_upload_frames_bgr (coeff_bits is 14 or 15)");
    if (stride < (size_t)c->W * 3) return fail(c, MFK_ERR_ARG, "stride %zu < 3 * width %d", stride, c->W);
    if (n_frames > c->store_frames) return fail(c, MFK_ERR_ARG, "%d frames do not fit a store of %d", n_frames, c->store_frames);
    CU(c, cudaSetDevice(c->device));
    const size_t px = (size_t)c->W * c->H;
    const int chunk = 32;                             // frames per staging round
    if (c->bgr_stage_bytes < (size_t)chunk * px * 3) {
        CU(c, cudaStreamSynchronize(c->copy_stream));
        cudaFree(c->bgr_stage); c->bgr_stage = nullptr; c->bgr_stage_bytes = 0;
        CU(c, cudaMalloc((void**)&c->bgr_stage, (size_t)chunk * px * 3 + 16));
        c->bgr_stage_bytes = (size_t)chunk * px * 3;
    }
    int done = 0;
    while (done < n_frames) {
        const int slot = (slot0 + done) % c->store_frames;
        const int run = std::min(std::min(n_frames - done, c->store_frames - slot), chunk);
        for (auto& r : c->readers)
            if (r.valid && r.overlaps(slot, slot + run - 1))
                for (int k = 0; k < mfk_ctx::MAX_LANES; k++) if (r.on[k]) CU(c, cudaStreamWaitEvent(c->copy_stream, r.ev[k], 0));
        const uint8_t* src = bgr + (size_t)done * stride * c->H;
        // the staging buffer is reused in stream order: the previous round's conversion kernel runs before this copy
        CU(c, cudaMemcpy2DAsync(c->bgr_stage, (size_t)c->W * 3, src, stride, (size_t)c->W * 3, (size_t)run * c->H,
                                cudaMemcpyHostToDevice, c->copy_stream));
        uint8_t* dst = c->store + (size_t)slot * px;
        launch_bgr2gray(c->copy_stream, c->bgr_stage, dst, (size_t)run * px, coeff_bits);
        c->launches++;
        { const int rc = note_upload(c, slot, slot + run - 1); if (rc != MFK_OK) return rc; }
        done += run;
    }
    CU(c, cudaEventRecord(c->upload_done, c->copy_stream));
    c->upload_pending = true;
    CU(c, cudaGetLastError());
    return MFK_OK;
}

int mfk_frame_store(mfk_ctx* c, uint8_t** device_base, int* store_frames) {
    if (!c) return MFK_ERR_ARG;
    if (device_base) *device_base = c->store;
    if (store_frames) *store_frames = c->store_frames;
    return MFK_OK;
}

int mfk_extract_pairs(mfk_ctx* c, const int32_t* cur_slot, const int32_t* prev_slot, const int32_t* label, int n_pairs) {
    if (!c || n_pairs < 0 || (n_pairs && (!cur_slot || !prev_slot || !label))) return fail(c, MFK_ERR_ARG, "bad argument to mfk_extract_pairs");
    if (n_pairs == 0) return MFK_OK;
    for (int i = 0; i < n_pairs; i++)
        if (cur_slot[i] < 0 || prev_slot[i] < 0) return fail(c, MFK_ERR_ARG, "pair %d names a negative store slot", i);
    CU(c, cudaSetDevice(c->device));
    int rc = grow_pair_rows(c, c->pairs_queued + (size_t)n_pairs);
    if (rc != MFK_OK) return rc;
    const size_t first = c->pairs_queued;
    c->upload_pending = false;
    // the slots this call reads, as maximal runs of consecutive slots; more runs than a record holds collapse into one
    // conservative [min, max] range
    std::vector<int> slots;
    slots.reserve(2 * (size_t)n_pairs);
    for (int i = 0; i < n_pairs; i++) { slots.push_back(cur_slot[i] % c->store_frames); slots.push_back(prev_slot[i] % c->store_frames); }
    std::sort(slots.begin(), slots.end());
    slots.erase(std::unique(slots.begin(), slots.end()), slots.end());
    int iv_lo[mfk_ctx::ReadRec::MAX_IV], iv_hi[mfk_ctx::ReadRec::MAX_IV], n_iv = 0;
    bool too_many = false;
    for (size_t i = 0; i < slots.size();) {
        size_t j = i;
        while (j + 1 < slots.size() && slots[j + 1] == slots[j] + 1) j++;
        if (n_iv == mfk_ctx::ReadRec::MAX_IV) { too_many = true; break; }
        iv_lo[n_iv] = slots[i]; iv_hi[n_iv] = slots[j]; n_iv++;
        i = j + 1;
    }
    if (too_many) { iv_lo[0] = slots.front(); iv_hi[0] = slots.back(); n_iv = 1; }
    // The lane-private front of a launch sequence (difference image, pyramid, scores, NMS) only waits for the uploads that
    // carry its frames; its shared-row part is additionally ordered after whatever the caller queued on the main stream
    // before this call (run_batch, main_wait_pending).  So the front of this call overlaps the tail of the previous one.
    CU(c, cudaEventRecord(c->main_mark, c->stream));
    c->main_wait_pending = true;
    bool touched[mfk_ctx::MAX_LANES] = {};
    for (int done = 0; done < n_pairs; done += c->max_pairs) {
        const int F = std::min(c->max_pairs, n_pairs - done);
        const int li = (int)(c->batches % (uint64_t)c->n_lanes);
        if (!touched[li]) {
            if (c->writers_folded_valid) CU(c, cudaStreamWaitEvent(c->lane[li].stream, c->writers_folded, 0));
            for (const mfk_ctx::WriteRec& w : c->writers) {
                if (!w.valid) continue;
                bool hit = false;
                for (int k = 0; k < n_iv && !hit; k++) hit = w.lo <= iv_hi[k] && w.hi >= iv_lo[k];
                if (hit) CU(c, cudaStreamWaitEvent(c->lane[li].stream, w.ev, 0));
            }
            touched[li] = true;
        }
        rc = run_batch(c, cur_slot + done, prev_slot + done, label + done, first + done, F);
        if (rc != MFK_OK) return rc;
    }
    c->pairs_queued += (size_t)n_pairs;
    mfk_ctx::ReadRec& rec = c->readers[c->reader_next];
    c->reader_next = (c->reader_next + 1) % 8;
    if (rec.valid)          // recycled record: keep its ordering
        for (int k = 0; k < mfk_ctx::MAX_LANES; k++) if (rec.on[k]) CU(c, cudaStreamWaitEvent(c->copy_stream, rec.ev[k], 0));
    rec.n_iv = n_iv; rec.valid = true;
    for (int i = 0; i < n_iv; i++) { rec.lo[i] = iv_lo[i]; rec.hi[i] = iv_hi[i]; }
    for (int k = 0; k < mfk_ctx::MAX_LANES; k++) {
        rec.on[k] = touched[k];
        if (touched[k]) CU(c, cudaEventRecord(rec.ev[k], c->lane[k].stream));
    }
    return MFK_OK;
}

static int read_counters(mfk_ctx* c, uint32_t h[4]) {
    { const int rc = settle(c); if (rc != MFK_OK) return rc; }
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaMemcpy(h, c->counters, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return MFK_OK;
}

int mfk_reset_rows(mfk_ctx* c) {
    if (!c) return MFK_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    // applied by the next launch sequence just before it touches the shared rows (so that its lane-private front may
    // start under the tail of the previous sequence), or by the next reader of the counters (settle)
    c->reset_pending = true;
    c->pairs_queued = 0;
    return MFK_OK;
}

int mfk_rows_device(mfk_ctx* c, const mfk_row** device_rows, size_t* n_rows) {
    if (!c) return MFK_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    uint32_t h[4];
    int rc = read_counters(c, h);
    if (rc != MFK_OK) return rc;
    if (device_rows) *device_rows = c->rows_buf;
    if (n_rows) *n_rows = h[1];
    if (h[0]) return fail(c, MFK_ERR_BUDGET, "budget overflow (%s%s); recreate the context with larger budgets (mfk_create_ex)",
                          (h[0] & 1) ? "candidates " : "", (h[0] & 2) ? "rows" : "");
    return MFK_OK;
}

int mfk_fetch_rows(mfk_ctx* c, mfk_row* rows, size_t cap, size_t* n_rows, uint32_t* rows_per_pair, size_t pairs_cap) {
    if (!c) return MFK_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    uint32_t h[4];
    int rc = read_counters(c, h);
    if (rc != MFK_OK) return rc;
    const size_t n = h[1];
    if (n_rows) *n_rows = n;
    const size_t ncopy = std::min(n, cap);
    if (rows && ncopy) CU(c, cudaMemcpy(rows, c->rows_buf, ncopy * sizeof(mfk_row), cudaMemcpyDeviceToHost));
    if (rows_per_pair) {
        const size_t np = std::min(pairs_cap, c->pairs_queued);
        if (np) CU(c, cudaMemcpy(rows_per_pair, c->pair_rows, np * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    }
    const uint32_t overflow = h[0];
    rc = mfk_reset_rows(c);
    if (rc != MFK_OK) return rc;
    if (overflow) return fail(c, MFK_ERR_BUDGET, "budget overflow (%s%s); recreate the context with larger budgets (mfk_create_ex)",
                              (overflow & 1) ? "candidates " : "", (overflow & 2) ? "rows" : "");
    if (n > cap) return fail(c, MFK_ERR_CAPACITY, "%zu rows do not fit the caller's buffer of %zu", n, cap);
    return MFK_OK;
}

int mfk_rows_text(mfk_ctx* c, char* text, size_t cap, size_t* n_bytes) {
    if (!c || (cap && !text)) return fail(c, MFK_ERR_ARG, "bad argument to mfk_rows_text");
    CU(c, cudaSetDevice(c->device));
    uint32_t h[4];
    int rc = read_counters(c, h);
    if (rc != MFK_OK) return rc;
    if (h[0]) return fail(c, MFK_ERR_BUDGET, "budget overflow (%s%s); recreate the context with larger budgets (mfk_create_ex)",
                          (h[0] & 1) ? "candidates " : "", (h[0] & 2) ? "rows" : "");
    return format_device_rows(c, c->rows_buf, h[1], text, cap, n_bytes);
}

int mfk_format_rows(mfk_ctx* c, const mfk_row* rows, size_t n, char* text, size_t cap, size_t* n_bytes) {
    if (!c || (n && !rows) || (cap && !text)) return fail(c, MFK_ERR_ARG, "bad argument to mfk_format_rows");
    CU(c, cudaSetDevice(c->device));
    if (n > c->text_rows_cap) {
        CU(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->text_rows); c->text_rows = nullptr; c->text_rows_cap = 0;
        CU(c, dalloc(&c->text_rows, n + n / 4));
        c->text_rows_cap = n + n / 4;
    }
    if (n) CU(c, cudaMemcpyAsync(c->text_rows, rows, n * sizeof(mfk_row), cudaMemcpyHostToDevice, c->stream));
    return format_device_rows(c, c->text_rows, n, text, cap, n_bytes);
}

static int enqueue_clip(mfk_ctx* c, const uint8_t* gray, size_t stride, int n_frames);

int mfk_extract_clip(mfk_ctx* c, const uint8_t* gray, size_t stride, int n_frames, mfk_row* rows, size_t cap, size_t* n_rows) {
    if (!c || !gray || n_frames < 0) return fail(c, MFK_ERR_ARG, "bad argument to mfk_extract_clip");
    if (n_rows) *n_rows = 0;
    const int rc = enqueue_clip(c, gray, stride, n_frames);
    if (rc != MFK_OK) return rc;
    return mfk_fetch_rows(c, rows, cap, n_rows, nullptr, 0);
}

int mfk_extract_clip_text(mfk_ctx* c, const uint8_t* gray, size_t stride, int n_frames, char* text, size_t text_cap, size_t* n_bytes,
                          mfk_row* rows, size_t cap, size_t* n_rows) {
    if (!c || !gray || n_frames < 0 || (text_cap && !text)) return fail(c, MFK_ERR_ARG, "bad argument to mfk_extract_clip_text");
    if (n_rows) *n_rows = 0;
    if (n_bytes) *n_bytes = 0;
    int rc = enqueue_clip(c, gray, stride, n_frames);
    if (rc != MFK_OK) return rc;
    uint32_t h[4];
    rc = read_counters(c, h);
    if (rc != MFK_OK) return rc;
    if (h[0]) {                    // sizes of an overflowed run mean nothing: report none, so that no caller re-reads it
        mfk_reset_rows(c);
        return fail(c, MFK_ERR_BUDGET, "budget overflow (%s%s); recreate the context with larger budgets (mfk_create_ex)",
                    (h[0] & 1) ? "candidates " : "", (h[0] & 2) ? "rows" : "");
    }
    if (n_rows) *n_rows = h[1];
    const int rc_text = format_device_rows(c, c->rows_buf, h[1], text, text_cap, n_bytes);
    if (rc_text != MFK_OK && rc_text != MFK_ERR_RANGE) return rc_text;       // a text buffer too small keeps the accumulation
    if (rows && h[1] > cap)
        return fail(c, MFK_ERR_CAPACITY, "%u rows do not fit the caller's buffer of %zu (accumulation kept)", h[1], cap);
    const std::string text_err = c->err;
    rc = rows ? mfk_fetch_rows(c, rows, cap, n_rows, nullptr, 0) : mfk_reset_rows(c);
    if (rc != MFK_OK) return rc;
    if (rc_text != MFK_OK) c->err = text_err;
    return rc_text;
}

static int enqueue_clip(mfk_ctx* c, const uint8_t* gray, size_t stride, int n_frames) {
    int rc = mfk_reset_rows(c);
    if (rc != MFK_OK) return rc;
    if (n_frames <= MFK_FRAME_GAP) return MFK_OK;
    // stream the clip through the store in windows that overlap by the 5-frame look-back
    const int window = c->store_frames;
    std::vector<int32_t> cur, prev, label;
    int uploaded = 0;      // frames [0, uploaded) have been copied; frame k lives in slot k % window
    int next = MFK_FRAME_GAP;
    while (next < n_frames) {
        // frames needed for pairs next .. last: next-5 .. last; keep at most `window` frames live
        const int last = std::min(n_frames - 1, next - MFK_FRAME_GAP + window - 1);
        if (uploaded <= last) {
            rc = mfk_upload_frames(c, gray + (size_t)uploaded * stride * c->H, stride, last + 1 - uploaded, uploaded % window);
            if (rc != MFK_OK) return rc;
            uploaded = last + 1;
        }
        cur.clear(); prev.clear(); label.clear();
        for (int k = next; k <= last; k++) { cur.push_back(k % window); prev.push_back((k - MFK_FRAME_GAP) % window); label.push_back(k - 1); }
        rc = mfk_extract_pairs(c, cur.data(), prev.data(), label.data(), (int)cur.size());
        if (rc != MFK_OK) return rc;
        next = last + 1;      // the next upload reuses slots; it is ordered after these kernels by the stream
    }
    return MFK_OK;
}

int mfk_mip_describe(mfk_ctx* c, int cur_slot, int prev_slot, const float* kp_xys, int n, uint8_t* motion) {
    if (!c || n < 0 || (n && (!kp_xys || !motion)) || cur_slot < 0 || prev_slot < 0) return fail(c, MFK_ERR_ARG, "bad argument to mfk_mip_describe");
    if (n == 0) return MFK_OK;
    CU(c, cudaSetDevice(c->device));
    const size_t err_off = (((size_t)n * 20 + 15) & ~(size_t)15);      // n x {3 floats} | n x 8 bytes | error word
    int rc = ensure_bow_tmp(c, err_off + 16);
    if (rc != MFK_OK) return rc;
    float* d_kp = (float*)c->bow_tmp;
    uint8_t* d_mot = (uint8_t*)c->bow_tmp + (size_t)n * 12;
    uint32_t* d_err = (uint32_t*)((uint8_t*)c->bow_tmp + err_off);
    CU(c, cudaStreamWaitEvent(c->stream, c->upload_done, 0));
    CU(c, cudaMemcpyAsync(d_kp, kp_xys, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemsetAsync(d_err, 0, 4, c->stream));
    const size_t px = (size_t)c->W * c->H;
    { Span s(c, P_MIP, 1);
      launch_mip_only(c->stream, c->W, c->H, c->store + (size_t)(cur_slot % c->store_frames) * px,
                      c->store + (size_t)(prev_slot % c->store_frames) * px, d_kp, n, d_mot, d_err, c->tb); }
    uint32_t herr = 0;
    CU(c, cudaMemcpyAsync(motion, d_mot, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(&herr, d_err, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (herr) return fail(c, MFK_ERR_ARG, "a keypoint's region of interest leaves the frame (cv::Mat ROI assertion in the reference)");
    return MFK_OK;
}

// ---- bag of words --------------------------------------------------------------------------------------

int mfk_bow_set_codebook(mfk_ctx* c, const uint8_t* codewords, int K) {
    if (!c || !codewords || K < 1) return fail(c, MFK_ERR_ARG, "bad argument to mfk_bow_set_codebook");
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->codebook); c->codebook = nullptr;
    cudaFree(c->bow_counts); c->bow_counts = nullptr;
    CU(c, cudaMalloc((void**)&c->codebook, (size_t)K * 16));
    CU(c, dalloc(&c->bow_counts, (size_t)K));
    CU(c, cudaMemcpy(c->codebook, codewords, (size_t)K * 16, cudaMemcpyHostToDevice));
    c->K = K; c->bow_counts_cap = K;
    return MFK_OK;
}

int mfk_bow_assign_device(mfk_ctx* c, const uint8_t* d_desc, size_t N, int32_t* d_idx) {
    if (!c || (N && (!d_desc || !d_idx))) return fail(c, MFK_ERR_ARG, "bad argument to mfk_bow_assign_device");
    if (!c->codebook) return fail(c, MFK_ERR_STATE, "no codebook set");
    CU(c, cudaSetDevice(c->device));
    if (((uintptr_t)d_desc & 15) != 0) return fail(c, MFK_ERR_ARG, "d_desc must be 16-byte aligned (descriptors are read as 128-bit words)");
    { Span s(c, P_BOW, 1); CU(c, launch_bow_assign(c->stream, d_desc, 16, N, nullptr, c->codebook, c->K, d_idx, nullptr)); }
    return MFK_OK;
}

int mfk_bow_assign(mfk_ctx* c, const uint8_t* desc, size_t N, int32_t* idx) {
    if (!c || (N && (!desc || !idx))) return fail(c, MFK_ERR_ARG, "bad argument to mfk_bow_assign");
    if (!c->codebook) return fail(c, MFK_ERR_STATE, "no codebook set");
    if (N == 0) return MFK_OK;
    CU(c, cudaSetDevice(c->device));
    int rc = ensure_bow_tmp(c, N * 20);
    if (rc != MFK_OK) return rc;
    uint8_t* d_desc = (uint8_t*)c->bow_tmp;
    int32_t* d_idx = (int32_t*)((uint8_t*)c->bow_tmp + N * 16);
    CU(c, cudaMemcpyAsync(d_desc, desc, N * 16, cudaMemcpyHostToDevice, c->stream));
    { Span s(c, P_BOW, 1); CU(c, launch_bow_assign(c->stream, d_desc, 16, N, nullptr, c->codebook, c->K, d_idx, nullptr)); }
    CU(c, cudaMemcpyAsync(idx, d_idx, N * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return MFK_OK;
}

int mfk_bow_histogram(mfk_ctx* c, const uint8_t* desc, size_t N, uint32_t* counts) {
    if (!c || !counts || (N && !desc)) return fail(c, MFK_ERR_ARG, "bad argument to mfk_bow_histogram");
    if (!c->codebook) return fail(c, MFK_ERR_STATE, "no codebook set");
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaMemsetAsync(c->bow_counts, 0, (size_t)c->K * 4, c->stream));
    if (N) {
        int rc = ensure_bow_tmp(c, N * 16);
        if (rc != MFK_OK) return rc;
        CU(c, cudaMemcpyAsync(c->bow_tmp, desc, N * 16, cudaMemcpyHostToDevice, c->stream));
        { Span s(c, P_BOW, 1); CU(c, launch_bow_assign(c->stream, (const uint8_t*)c->bow_tmp, 16, N, nullptr, c->codebook, c->K, nullptr, c->bow_counts)); }
    }
    CU(c, cudaMemcpyAsync(counts, c->bow_counts, (size_t)c->K * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return MFK_OK;
}

// Counts of the accumulated rows into c->bow_counts (device), queued on the main stream without a host round trip: the
// kernel takes the live row count from the device counters.  finish_counts() reads counts and overflow flags back.
static int histogram_rows_device(mfk_ctx* c) {
    { const int rc = settle(c); if (rc != MFK_OK) return rc; }
    CU(c, cudaMemsetAsync(c->bow_counts, 0, (size_t)c->K * 4, c->stream));
    Span s(c, P_BOW, 1);
    // descriptor = appearance | motion = bytes 16..31 of each 32-byte row
    CU(c, launch_bow_assign(c->stream, (const uint8_t*)c->rows_buf + 16, 32, c->row_cap, c->counters + 1, c->codebook, c->K, nullptr, c->bow_counts));
    return MFK_OK;
}

static int finish_counts(mfk_ctx* c, uint32_t* counts) {
    uint32_t h[4] = {0, 0, 0, 0};
    CU(c, cudaMemcpyAsync(counts, c->bow_counts, (size_t)c->K * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(h, c->counters, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (h[0]) return fail(c, MFK_ERR_BUDGET, "budget overflow (%s%s); recreate the context with larger budgets (mfk_create_ex)",
                          (h[0] & 1) ? "candidates " : "", (h[0] & 2) ? "rows" : "");
    return MFK_OK;
}

int mfk_bow_histogram_rows(mfk_ctx* c, uint32_t* counts) {
    if (!c || !counts) return fail(c, MFK_ERR_ARG, "bad argument to mfk_bow_histogram_rows");
    if (!c->codebook) return fail(c, MFK_ERR_STATE, "no codebook set");
    CU(c, cudaSetDevice(c->device));
    { const int rc = histogram_rows_device(c); if (rc != MFK_OK) return rc; }
    return finish_counts(c, counts);
}

// NCCL is resolved at run time so that the library loads on machines without it and shares the copy a host
// process (e.g. torch) has already loaded.
namespace {
struct NcclId { char internal[128]; };
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_getid_fn)(NcclId*);
typedef int (*nccl_init_fn)(void**, int, NcclId, int);
typedef int (*nccl_destroy_fn)(void*);
void* nccl_symbol(const char* name) {
    static void* h = nullptr;
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    return h ? dlsym(h, name) : nullptr;
}
}  // namespace

int mfk_allreduce_counts(mfk_ctx* c, uint32_t* d_counts, size_t n, void* nccl_comm) {
    if (!c || !d_counts || !nccl_comm) return fail(c, MFK_ERR_ARG, "bad argument to mfk_allreduce_counts");
    static nccl_allreduce_fn fn = (nccl_allreduce_fn)nccl_symbol("ncclAllReduce");
    if (!fn) return fail(c, MFK_ERR_STATE, "NCCL is not available (libnccl.so.2 not found)");
    CU(c, cudaSetDevice(c->device));
    { const int rc = join_lanes(c); if (rc != MFK_OK) return rc; }
    const int ncclUint32 = 3, ncclSum = 0;      // nccl.h: ncclDataType_t / ncclRedOp_t
    const int r = fn(d_counts, d_counts, n, ncclUint32, ncclSum, nccl_comm, c->stream);
    if (r != 0) return fail(c, MFK_ERR_CUDA, "ncclAllReduce failed with code %d", r);
    return MFK_OK;
}

int mfk_bow_histogram_videos(mfk_ctx* c, const uint32_t* pairs_per_video, int n_videos, uint32_t* counts, void* nccl_comm) {
    if (!c || !counts || n_videos < 1) return fail(c, MFK_ERR_ARG, "bad argument to mfk_bow_histogram_videos");
    if (!c->codebook) return fail(c, MFK_ERR_STATE, "no codebook set");
    CU(c, cudaSetDevice(c->device));
    if (!pairs_per_video && n_videos != 1) return fail(c, MFK_ERR_ARG, "pairs_per_video is required for more than one video");
    std::vector<uint32_t> start((size_t)n_videos + 1, 0);
    if (pairs_per_video) for (int v = 0; v < n_videos; v++) start[v + 1] = start[v] + pairs_per_video[v];
    else start[1] = (uint32_t)c->pairs_queued;
    if (start[n_videos] != c->pairs_queued)
        return fail(c, MFK_ERR_ARG, "pairs_per_video sums to %u, but %zu pairs were queued since the last reset", start[n_videos], c->pairs_queued);
    const size_t cells = (size_t)n_videos * c->K;
    if (cells > (size_t)c->bow_counts_cap) {
        CU(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->bow_counts); c->bow_counts = nullptr; c->bow_counts_cap = 0;
        CU(c, dalloc(&c->bow_counts, cells));
        c->bow_counts_cap = (int)cells;
    }
    // scratch: pair_start [n+1] | row offsets [n+2] | scan scratch
    const size_t nv1 = (size_t)n_videos + 1;
    { const int rc = ensure_bow_tmp(c, (2 * nv1 + 1 + nv1 / 2048 + 8) * sizeof(uint32_t)); if (rc != MFK_OK) return rc; }
    uint32_t* d_start = (uint32_t*)c->bow_tmp;
    uint32_t* d_off = d_start + nv1;
    uint32_t* d_scratch = d_off + nv1 + 1;
    { const int rc = settle(c); if (rc != MFK_OK) return rc; }
    // pageable source: the runtime stages it before returning, so `start` may go out of scope
    CU(c, cudaMemcpyAsync(d_start, start.data(), nv1 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemsetAsync(c->bow_counts, 0, cells * 4, c->stream));
    { Span s(c, P_BOW, 5);
      launch_video_rows(c->stream, c->pair_rows, d_start, n_videos, d_off);
      launch_exclusive_scan(c->stream, d_off, n_videos, nullptr, d_scratch);
      // descriptor = appearance | motion = bytes 16..31 of each 32-byte row; the live row count is read on the device
      CU(c, launch_bow_assign(c->stream, (const uint8_t*)c->rows_buf + 16, 32, c->row_cap, c->counters + 1, c->codebook, c->K, nullptr,
                              c->bow_counts, d_off, n_videos)); }
    if (nccl_comm) { const int rc = mfk_allreduce_counts(c, c->bow_counts, cells, nccl_comm); if (rc != MFK_OK) return rc; }
    uint32_t h[4] = {0, 0, 0, 0};
    CU(c, cudaMemcpyAsync(counts, c->bow_counts, cells * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(h, c->counters, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (h[0]) return fail(c, MFK_ERR_BUDGET, "budget overflow (%s%s); recreate the context with larger budgets (mfk_create_ex)",
                          (h[0] & 1) ? "candidates " : "", (h[0] & 2) ? "rows" : "");      // this rank's own flags
    return MFK_OK;
}

int mfk_nccl_unique_id(uint8_t id_out[128]) {
    static nccl_getid_fn fn = (nccl_getid_fn)nccl_symbol("ncclGetUniqueId");
    if (!fn || !id_out) return fail(nullptr, MFK_ERR_STATE, "NCCL is not available (libnccl.so.2 not found)");
    NcclId id;
    const int r = fn(&id);
    if (r != 0) return fail(nullptr, MFK_ERR_CUDA, "ncclGetUniqueId failed with code %d", r);
    memcpy(id_out, id.internal, 128);
    return MFK_OK;
}

int mfk_nccl_comm_init(mfk_ctx* c, void** nccl_comm_out, int n_ranks, int rank, const uint8_t id_bytes[128]) {
    if (!c || !nccl_comm_out || !id_bytes) return fail(c, MFK_ERR_ARG, "bad argument to mfk_nccl_comm_init");
    static nccl_init_fn fn = (nccl_init_fn)nccl_symbol("ncclCommInitRank");
    if (!fn) return fail(c, MFK_ERR_STATE, "NCCL is not available (libnccl.so.2 not found)");
    CU(c, cudaSetDevice(c->device));
    NcclId id;
    memcpy(id.internal, id_bytes, 128);
    const int r = fn(nccl_comm_out, n_ranks, id, rank);
    if (r != 0) return fail(c, MFK_ERR_CUDA, "ncclCommInitRank failed with code %d", r);
    return MFK_OK;
}

int mfk_nccl_comm_destroy(void* nccl_comm) {
    static nccl_destroy_fn fn = (nccl_destroy_fn)nccl_symbol("ncclCommDestroy");
    if (!fn || !nccl_comm) return MFK_ERR_ARG;
    return fn(nccl_comm) == 0 ? MFK_OK : MFK_ERR_CUDA;
}

// ---- measurement ----------------------------------------------------------------------------------------

int mfk_timer_start(mfk_ctx* c) {
    if (!c) return MFK_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    // the lane-private front of the next launch sequence does not wait for the main stream, so start from an idle device
    { const int rc = sync_all(c); if (rc != MFK_OK) return rc; }
    CU(c, cudaEventRecord(c->t0, c->stream));
    return MFK_OK;
}

int mfk_timer_stop(mfk_ctx* c, float* elapsed_ms) {
    if (!c || !elapsed_ms) return MFK_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    { const int rc = join_lanes(c); if (rc != MFK_OK) return rc; }
    CU(c, cudaEventRecord(c->t1, c->stream));
    CU(c, cudaEventSynchronize(c->t1));
    CU(c, cudaEventElapsedTime(elapsed_ms, c->t0, c->t1));
    return MFK_OK;
}

int mfk_profile_enable(mfk_ctx* c, int enable) {
    if (!c) return MFK_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    { const int rc = sync_all(c); if (rc != MFK_OK) return rc; }
    drain_spans(c);
    c->prof_on = enable != 0;
    memset(c->prof_launches, 0, sizeof c->prof_launches);
    memset(c->prof_ms, 0, sizeof c->prof_ms);
    return MFK_OK;
}

int mfk_profile_read(mfk_ctx* c, mfk_prof_entry* entries, int cap, int* n_entries) {
    if (!c || !entries || !n_entries) return MFK_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    { const int rc = sync_all(c); if (rc != MFK_OK) return rc; }
    drain_spans(c);
    int n = 0;
    for (int i = 0; i < P_COUNT && n < cap; i++) {
        entries[n].name = kProfNames[i];
        entries[n].launches = c->prof_launches[i];
        entries[n].total_ms = c->prof_ms[i];
        n++;
    }
    *n_entries = n;
    return MFK_OK;
}

int mfk_profile_timeline(mfk_ctx* c, mfk_span* spans, int cap, int* n_spans) {
    if (!c || !spans || !n_spans || cap < 0) return MFK_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    { const int rc = sync_all(c); if (rc != MFK_OK) return rc; }
    int n = 0;
    for (const ProfSpan& s : c->spans) {
        if (n >= cap) break;
        float a = 0.f, b = 0.f;
        if (cudaEventElapsedTime(&a, c->t0, s.a) != cudaSuccess || cudaEventElapsedTime(&b, c->t0, s.b) != cudaSuccess) continue;
        spans[n].name = kProfNames[s.id];
        spans[n].start_ms = a; spans[n].end_ms = b;
        n++;
    }
    *n_spans = n;
    return MFK_OK;
}

int mfk_set_serial(mfk_ctx* c, int serial) {
    if (!c) return MFK_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    { const int rc = sync_all(c); if (rc != MFK_OK) return rc; }
    c->serial = serial != 0;
    c->n_lanes = c->serial ? 1 : c->lanes_built;
    return MFK_OK;
}

uint64_t mfk_launch_count(const mfk_ctx* c) { return c ? c->launches : 0; }

// ---- stage access for parity tests ---------------------------------------------------------------------

int mfk_debug_layer(mfk_ctx* c, int what, int layer, int pair, uint8_t* out, int* dims) {
    if (!c || !out || layer < 0 || layer >= N_LAYERS || what < 0 || what > 1) return fail(c, MFK_ERR_ARG, "bad argument to mfk_debug_layer");
    if (!c->have_last || pair < 0 || pair >= c->last.F) return fail(c, MFK_ERR_STATE, "no such pair in the last batch");
    CU(c, cudaSetDevice(c->device));
    { const int rc = sync_all(c); if (rc != MFK_OK) return rc; }
    const LayerGeom& L = c->last.L[layer];
    const Maps& maps = c->lane[c->last_lane].maps;
    const uint8_t* base = (what == 0 ? maps.img : maps.S) + L.base + (size_t)pair * L.px;
    CU(c, cudaMemcpy(out, base, (size_t)L.px, cudaMemcpyDeviceToHost));
    if (dims) { dims[0] = L.w; dims[1] = L.h; }
    return MFK_OK;
}

int mfk_debug_counters(mfk_ctx* c, uint32_t out[8]) {
    if (!c || !out) return MFK_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    { const int rc = sync_all(c); if (rc != MFK_OK) return rc; }
    CU(c, cudaMemcpy(out, c->counters, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return MFK_OK;
}

int mfk_debug_keypoints(mfk_ctx* c, int pair, float* out, int cap, int* n) {
    if (!c || !n) return fail(c, MFK_ERR_ARG, "bad argument to mfk_debug_keypoints");
    if (!c->have_last || pair < 0 || pair >= c->last.F) return fail(c, MFK_ERR_STATE, "no such pair in the last batch");
    CU(c, cudaSetDevice(c->device));
    { const int rc = sync_all(c); if (rc != MFK_OK) return rc; }
    const Geom& g = c->last;
    const CandBuf& cb = c->lane[c->last_lane].cb;
    uint32_t range[2];
    CU(c, cudaMemcpy(&range[0], cb.row_count + (size_t)pair * g.rows, 4, cudaMemcpyDeviceToHost));
    CU(c, cudaMemcpy(&range[1], cb.row_count + (size_t)(pair + 1) * g.rows, 4, cudaMemcpyDeviceToHost));
    range[0] = std::min<uint32_t>(range[0], (uint32_t)g.cand_cap);
    range[1] = std::min<uint32_t>(range[1], (uint32_t)g.cand_cap);
    const size_t m = range[1] - range[0];
    std::vector<uint8_t> flags(m);
    std::vector<uint32_t> meta(m);
    std::vector<float> kx(m), ky(m), ks(m), kr(m);
    if (m) {
        CU(c, cudaMemcpy(flags.data(), cb.flags + range[0], m, cudaMemcpyDeviceToHost));
        CU(c, cudaMemcpy(meta.data(), cb.meta + range[0], m * 4, cudaMemcpyDeviceToHost));
        CU(c, cudaMemcpy(kx.data(), cb.kx + range[0], m * 4, cudaMemcpyDeviceToHost));
        CU(c, cudaMemcpy(ky.data(), cb.ky + range[0], m * 4, cudaMemcpyDeviceToHost));
        CU(c, cudaMemcpy(ks.data(), cb.ksize + range[0], m * 4, cudaMemcpyDeviceToHost));
        CU(c, cudaMemcpy(kr.data(), cb.kresp + range[0], m * 4, cudaMemcpyDeviceToHost));
    }
    int count = 0;
    for (size_t i = 0; i < m; i++) {
        if (!(flags[i] & ST_KEYPOINT)) continue;
        if (out && count < cap) {
            out[5 * count] = kx[i]; out[5 * count + 1] = ky[i]; out[5 * count + 2] = ks[i];
            out[5 * count + 3] = kr[i]; out[5 * count + 4] = (float)(meta[i] & 7);
        }
        count++;
    }
    *n = count;
    return MFK_OK;
}

}  // extern "C"
