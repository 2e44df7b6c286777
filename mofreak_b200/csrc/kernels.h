// Launchers of the sm_100a kernels (definitions in kernels_*.cu).  All take the stream to launch on and
// raw device pointers; none synchronises.  Buffers are described in context.cu / DESIGN.md.
#pragma once
#include <cuda_runtime.h>

#include "geom.h"
#include "tables.h"

namespace mfk {

struct PairList {                 // per-batch pair description (device arrays, F entries each)
    const int32_t* cur_slot;
    const int32_t* prev_slot;
    const int32_t* label;
};

struct Maps {                     // F * pyr_px bytes each, layout in geom.h
    uint8_t* img;                 // pyramid; layer 0 is |I_t - I_{t-5}|
    uint8_t* S;                   // closed-form OAST-9/16 scores
    uint8_t* seen;                // getScoreMaxAbove write-back marks (layers 1..5 used)
    uint8_t* status;              // ST_* per candidate pixel (valid only where S >= threshold)
};

struct CandBuf {                  // ordered candidate list of the batch + per-candidate results
    uint32_t* row_count;          // [F*rows + 1] hits per (pair, layer row); scanned in place -> offsets
    uint32_t* xy;                 // x | y << 16
    uint32_t* meta;               // pair << 3 | layer
    float* kx; float* ky; float* ksize; float* kresp;   // refined keypoint of surviving candidates
    uint8_t* flags;               // ST_* (ST_SURVIVE+score test passed = keypoint, ST_KEEP = output row)
    uint8_t* scale_idx;           // FREAK scale index of kept candidates
    uint32_t* keep_scan;          // [cand_cap + 1] exclusive scan of ST_KEEP
    uint32_t* row_src;            // [cand_cap] candidate index of the batch's r-th output row
    uint32_t* pend_count;         // [F*6] tie candidates per (pair, layer), appended by phase A
    uint32_t* pend_list;          // [F*6][PEND_CAP] their candidate indices (unordered)
    // phase A work queues (unordered; device-wide re-packing between the steps of phase A)
    uint32_t* q_count;            // [3] entries in q0 / q1 / q2
    uint32_t* q0;                 // [q_cap] candidates that passed the 8-neighbour test without a tie
    uint32_t* q1; float* q1_up;   // [q_cap] after the scan above, with its (max, dx, dy)
    uint32_t* q2; float* q2_scan; // [q_cap] after the scan below, with (up.max, up.dx, up.dy, dn.max, dn.dx, dn.dy)
    uint32_t q_cap;
    uint32_t* counters;           // shared by all batches: [0] overflow flags (1 candidates, 2 rows), [1] rows accumulated
                                  // so far, [2] pairs accumulated so far, [3] live candidates of the last batch (copy)
    uint32_t* live;               // [0] live candidate count of THIS batch (per buffer set), [1] its overflow flags until the advance step,
                                  // [2] row ticket counter of k_describe, [3] rows_base of this batch (both set by the advance step of k_scatter_rows)
};

enum { PEND_CAP = 2048 };         // per (pair, layer); beyond it phase B falls back to an ordered sweep

struct DeviceTables {             // device copies of mfk::Tables
    const FreakPoint* points;
    const int32_t* pattern_size;      // [64]
    const float* scale_threshold;     // [64]
    const OrientPair* orient;         // [45]
    const UsedPair* used;             // [64]
    const uint16_t* mip_needed;       // patch pixels the MIP tests read (tables.h)
    int n_mip_needed;
    const uint32_t* mip_plan;         // the same pixels as the descriptor kernel's resize loop wants them: byte offset of the pixel in
                                      // the warp's [current | previous] patch buffers (10 bits) | 16 * column << 10 | 16 * row << 19 | frame << 28
};

// --- dense stage ------------------------------------------------------------------------------------
// interleaved BGR -> gray as cv::cvtColor computes it (coeff_bits 14: OpenCV 2.4, 15: OpenCV >= 3)
void launch_bgr2gray(cudaStream_t, const uint8_t* bgr, uint8_t* gray, size_t n_px, int coeff_bits);
// pair descriptions travel as kernel parameters (no staging copy, no host stall): 256 per launch; returns launches made
int launch_set_pairs(cudaStream_t, const int32_t* cur_host, const int32_t* prev_host, const int32_t* label_host, int n, int store_frames, PairList dst);   // slots land reduced modulo the store size
void launch_absdiff(cudaStream_t, const Geom&, const uint8_t* store, int store_frames, PairList, uint8_t* img);
void launch_pyramid(cudaStream_t, const Geom&, uint8_t* img);                                  // layers 1..5
int  score_mask_words(int frame_width);        // 64-bit hit-mask words per (pair, layer row)
// k_scores' tensor maps of one lane's pyramid buffer (opaque here: CUtensorMap needs cuda.h), rebuilt when F or img change
struct ScoreMapCache {
    int F = -1;
    const uint8_t* img = nullptr;
    alignas(64) unsigned char blob[N_LAYERS * 128 + 64];
};
void launch_scores(cudaStream_t, const Geom&, const uint8_t* img, uint8_t* S, uint8_t* status_or_null, unsigned long long* hitmask, ScoreMapCache& cache);
void launch_row_counts(cudaStream_t, const Geom&, const unsigned long long* hitmask, uint32_t* row_count);
void launch_fill_candidates(cudaStream_t, const Geom&, const unsigned long long* hitmask, CandBuf);
// left_done: the per-row strip sums were already written by launch_front
void launch_integral(cudaStream_t, const Geom&, const uint8_t* img /*layer 0*/, int32_t* integral,
                     int32_t* left /*F*H*ceil(W/32) scratch*/, bool left_done);
// the fused front (kernels_front.cu): difference image, pyramid layers 1..5 and the integral's row strip sums in ONE kernel
// from TMA-staged frame bands; replaces launch_absdiff + launch_pyramid + the first pass of launch_integral for frames up
// to front_supported() wide
bool front_supported(const Geom&);
cudaError_t launch_front(cudaStream_t, const Geom&, const uint8_t* store, int store_frames, PairList, uint8_t* img, int32_t* left);
// --- scans --------------------------------------------------------------------------------------------
// exclusive scan of data[0..n) in place, data[n] = total; n is a host upper bound, *d_n (optional
// device value <= n) the live length; scratch holds ceil(n/2048)+1 words
void launch_exclusive_scan(cudaStream_t, uint32_t* data, int n, const uint32_t* d_n, uint32_t* scratch);
// --- sparse stage -----------------------------------------------------------------------------------
void launch_phase_a(cudaStream_t, const Geom&, Maps, CandBuf, DeviceTables);
void launch_phase_b(cudaStream_t, const Geom&, Maps, CandBuf, DeviceTables);
void launch_keep_flags(cudaStream_t, const Geom&, CandBuf);           // flags -> keep_scan input
void launch_scatter_rows(cudaStream_t, const Geom&, CandBuf, uint32_t* pair_rows);   // after the scan: row_src, per-pair counts (pair_rows = this batch's slice)
void launch_describe(cudaStream_t, const Geom&, const uint8_t* store, int store_frames, PairList, const uint8_t* img,
                     const int32_t* integral, CandBuf, DeviceTables, void* rows /*mfk_row*/);
// byte offset of patch pixel (column dx, row dy) of frame fr (0 current, 1 previous) in a warp's patch buffers (kernels_desc.cu)
uint32_t describe_patch_offset(int fr, int dx, int dy);
void launch_mip_only(cudaStream_t, int W, int H, const uint8_t* cur, const uint8_t* prev, const float* kp_xys, int n,
                     uint8_t* motion, uint32_t* err_flag, DeviceTables);
// --- bag of words --------------------------------------------------------------------------------------
// returns the launch status (the shared-memory opt-in is per device and is set on every call)
// n descriptors, or min(n, *d_n) when d_n (a device word) is given.  With seg_off (n_seg + 1 ascending row offsets) the
// counts are per segment: counts[seg * K + word]
cudaError_t launch_bow_assign(cudaStream_t, const uint8_t* desc, size_t stride_bytes, size_t n, const uint32_t* d_n /*nullable*/,
                              const uint8_t* codebook, int K, int32_t* idx, uint32_t* counts /*nullable*/,
                              const uint32_t* seg_off = nullptr, int n_seg = 0);
void launch_video_rows(cudaStream_t, const uint32_t* pair_rows, const uint32_t* pair_start, int n_videos, uint32_t* out);

// `.mofreak` text (kernels_text.cu): line lengths into len[0..n] (slot n = 0, the scan makes it the total; flags |= 1 if a float
// needs exponent notation), then the lines at their scanned offsets
void launch_text_lengths(cudaStream_t, const void* rows, uint32_t n, uint32_t* len, uint32_t* flags);
void launch_text_write(cudaStream_t, const void* rows, uint32_t n, const uint32_t* off, char* text);

}  // namespace mfk
