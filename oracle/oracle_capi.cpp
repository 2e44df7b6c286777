// TEST INFRASTRUCTURE ONLY (see oracle.h).  extern "C" surface of the oracle for tests/ and for the
// cpu_baseline / --impl reference legs of bench.py (ctypes).  All buffers are caller owned.
#include "oracle.h"

#include <omp.h>
#include <string.h>
#include <chrono>

using namespace orc;

#ifdef MFK_ORACLE_WITH_REF
extern "C" int ref_brisk_detect(const uint8_t* img, int w, int h, int thresh, int octaves, float* out, int cap);
namespace orc {
void brisk_ref_detect(const uint8_t* img, int w, int h, std::vector<Keypoint>& out) {
    std::vector<float> buf(5 * 4096);
    int n = ref_brisk_detect(img, w, h, 30, 3, buf.data(), (int)buf.size() / 5);
    if (n > (int)buf.size() / 5) {
        buf.resize((size_t)5 * n);
        n = ref_brisk_detect(img, w, h, 30, 3, buf.data(), n);
    }
    out.resize(n);
    for (int i = 0; i < n; i++) {
        out[i].x = buf[5 * i]; out[i].y = buf[5 * i + 1]; out[i].size = buf[5 * i + 2];
        out[i].response = buf[5 * i + 3]; out[i].layer = (int)buf[5 * i + 4];
    }
}
}
#endif

static DetectFn pick_detector(int use_ref) {
#ifdef MFK_ORACLE_WITH_REF
    if (use_ref) return brisk_ref_detect;
#else
    (void)use_ref;
#endif
    return brisk_port_detect;
}

extern "C" {

// 1 when this library carries the reference's own detector object code (oracle/_ref build)
int orc_has_reference(void) {
#ifdef MFK_ORACLE_WITH_REF
    return 1;
#else
    return 0;
#endif
}

// out: rows of {x, y, size, response, layer}; returns the keypoint count (writes at most cap)
int orc_brisk_detect(const uint8_t* img, int w, int h, int thresh, int jacobi_reverse, float* out, int cap) {
    std::vector<Keypoint> k;
    brisk_port_detect_ex(img, w, h, thresh, jacobi_reverse != 0, k);
    for (size_t i = 0; i < k.size() && (int)i < cap; i++) {
        out[5 * i] = k[i].x; out[5 * i + 1] = k[i].y; out[5 * i + 2] = k[i].size;
        out[5 * i + 3] = k[i].response; out[5 * i + 4] = (float)k[i].layer;
    }
    return (int)k.size();
}
void orc_halfsample(const uint8_t* src, int w, int h, uint8_t* dst) { port_halfsample(src, w, h, dst); }
void orc_twothirdsample(const uint8_t* src, int w, int h, uint8_t* dst) { port_twothirdsample(src, w, h, dst); }
int  orc_oast_score(const uint8_t* img, int w, int x, int y) { return port_oast_score(img, w, x, y); }
int  orc_agast58_score(const uint8_t* img, int w, int x, int y) { return port_agast58_score(img, w, x, y); }
void orc_score_map(const uint8_t* img, int w, int h, uint8_t* S) { port_score_map(img, w, h, S); }

void orc_integral(const uint8_t* img, int w, int h, int32_t* sum) { integral_u8(img, w, h, sum); }
int  orc_freak_pattern_size(int scale_idx) { return freak_pattern_size(freak_pattern(), scale_idx); }
int  orc_freak_scale_index(float size) { return freak_scale_index(size); }

// kps_io: n rows of {x, y, size, response, layer}, compacted in place to the survivors of FREAK's
// border filter; desc: 64 bytes per survivor; theta: orientation bin per survivor.  Returns the count.
int orc_freak_compute(const uint8_t* img, int w, int h, float* kps_io, int n, uint8_t* desc, int* theta) {
    std::vector<Keypoint> k(n);
    for (int i = 0; i < n; i++) {
        k[i].x = kps_io[5 * i]; k[i].y = kps_io[5 * i + 1]; k[i].size = kps_io[5 * i + 2];
        k[i].response = kps_io[5 * i + 3]; k[i].layer = (int)kps_io[5 * i + 4];
    }
    std::vector<uint8_t> d;
    std::vector<int> t;
    freak_compute(freak_pattern(), img, w, h, k, d, &t);
    for (size_t i = 0; i < k.size(); i++) {
        kps_io[5 * i] = k[i].x; kps_io[5 * i + 1] = k[i].y; kps_io[5 * i + 2] = k[i].size;
        kps_io[5 * i + 3] = k[i].response; kps_io[5 * i + 4] = (float)k[i].layer;
        if (theta) theta[i] = t[i];
    }
    if (desc && !d.empty()) memcpy(desc, d.data(), d.size());
    return (int)k.size();
}

void orc_resize19(const uint8_t* src, int stride, int s, uint8_t* dst) { resize_linear_u8_19(src, stride, s, dst); }
unsigned orc_mip_byte(const uint8_t* cur19, const uint8_t* prev19, int x, int y) { return mip_byte(cur19, prev19, x, y); }
void orc_mip_extract(const uint8_t* cur, const uint8_t* prev, int w, int h, float size, int x, int y, uint8_t* motion) {
    mip_extract(cur, prev, w, h, size, x, y, motion);
}

// One clip.  rows_out: up to cap 32-byte rows; returns the row count (may exceed cap).
long orc_extract_clip(const uint8_t* frames, int w, int h, int n_frames, int use_ref, int as_written,
                      void* rows_out, long cap) {
    ExtractOptions opt; opt.as_written = as_written != 0; opt.detect = pick_detector(use_ref);
    std::vector<Row> rows;
    extract_clip(frames, w, h, n_frames, opt, rows);
    const long n = (long)rows.size();
    if (rows_out && n) memcpy(rows_out, rows.data(), sizeof(Row) * (size_t)(n < cap ? n : cap));
    return n;
}

// Timing leg: n_clips clips of n_frames each, one clip per OpenMP thread as the fair multi-core form
// of the reference's serial per-video loop (main.cpp:854-924).  Returns wall seconds; totals[0] =
// processed frames, totals[1] = rows.
double orc_time_extract(const uint8_t* clips, int w, int h, int n_frames, int n_clips, int use_ref,
                        int as_written, int threads, long* totals) {
    ExtractOptions opt; opt.as_written = as_written != 0; opt.detect = pick_detector(use_ref);
    freak_pattern();
    long rows_total = 0;
    const size_t clip_bytes = (size_t)w * h * n_frames;
    if (threads < 1) threads = 1;
    const auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(+ : rows_total)
    for (int c = 0; c < n_clips; c++) {
        std::vector<Row> rows;
        extract_clip(clips + c * clip_bytes, w, h, n_frames, opt, rows);
        rows_total += (long)rows.size();
    }
    const auto t1 = std::chrono::steady_clock::now();
    totals[0] = (long)n_clips * (n_frames > 5 ? n_frames - 5 : 0);
    totals[1] = rows_total;
    return std::chrono::duration<double>(t1 - t0).count();
}

// text: malloc-free two-call protocol — returns the length; copies at most cap bytes into out.
long orc_format_mofreak(const void* rows, long n, char* out, long cap) {
    const std::string s = format_mofreak((const Row*)rows, (size_t)n);
    if (out && cap > 0) memcpy(out, s.data(), (size_t)((long)s.size() < cap ? (long)s.size() : cap));
    return (long)s.size();
}

unsigned orc_hamming_u8(unsigned a, unsigned b) { return hamming_u8((unsigned char)a, (unsigned char)b); }
void orc_bow_assign(const uint8_t* desc, long n, const uint8_t* codebook, int K, int32_t* idx) {
    for (long i = 0; i < n; i++) idx[i] = brute_force_match(desc + 16 * i, codebook, K);
}
double orc_time_bow_assign(const uint8_t* desc, long n, const uint8_t* codebook, int K, int threads, int32_t* idx) {
    if (threads < 1) threads = 1;
    const auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(static) num_threads(threads)
    for (long i = 0; i < n; i++) idx[i] = brute_force_match(desc + 16 * i, codebook, K);
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}
// returns the number of codewords parsed; copies at most cap_words*16 bytes
int orc_load_clusters(const char* text, long len, uint8_t* out, int cap_words) {
    const std::vector<uint8_t> c = load_clusters_text(std::string(text, (size_t)len));
    const int K = (int)(c.size() / 16);
    if (out) memcpy(out, c.data(), (size_t)16 * (K < cap_words ? K : cap_words));
    return K;
}
int orc_build_histogram(const char* text, long len, const uint8_t* codebook, int K, float* hist) {
    std::vector<float> h;
    const bool ok = build_histogram(std::string(text, (size_t)len), codebook, K, h);
    memcpy(hist, h.data(), sizeof(float) * K);
    return ok ? 1 : 0;
}

// TRECVID sliding window (sliding_restate.cpp): two-call protocol, returns the text length
long orc_sliding_bow(const char* text, long len, const uint8_t* codebook, int K, int alpha, int label, char* out, long cap) {
    const std::string s = sliding_bag_of_words(std::string(text, (size_t)len), codebook, K, alpha, label);
    if (out && cap > 0) memcpy(out, s.data(), (size_t)((long)s.size() < cap ? (long)s.size() : cap));
    return (long)s.size();
}

// codebook selection (cluster_restate.cpp): data is shuffled in place; returns the random draws consumed
long orc_random_clusters(uint8_t* data, int rows, int dim, int num_clusters, int num_classes, const int* classes, int n_classes,
                         int only_one_class, const int* rands, uint8_t* centers, int* center_row) {
    return (long)random_clusters(data, rows, dim, num_clusters, num_classes, classes, n_classes, only_one_class != 0, rands, centers, center_row);
}
long orc_clusters_text(const uint8_t* centers, int num_clusters, int dim, char* out, long cap) {
    const std::string s = clusters_text(centers, num_clusters, dim);
    if (out && cap > 0) memcpy(out, s.data(), (size_t)((long)s.size() < cap ? (long)s.size() : cap));
    return (long)s.size();
}

}  // extern "C"
