/* mofreak_b200 — C ABI of the B200 (sm_100a) MoFREAK extraction + bag-of-words hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.  The reference
 * (ChrisWhiten/MoFREAK) has no FFI; the functions it would bind for this path are ordinary C++
 * members, cited per entry point below as src/MoFREAK/<file>:<lines>.  The host classes in
 * mofreak_b200/host/ (MoFREAKUtilities, MotionInterchangePatterns, BagOfWordsRepresentation) keep the
 * reference's call surface and are implemented on top of exactly these symbols.
 *
 * Conventions: every function returns 0 (MFK_OK) or a negative MFK_ERR_* code and never throws; the
 * message of the last failure on a context is mfk_last_error().  One context per GPU; a context is not
 * thread safe (calls on one context must be serialised by the caller, as the reference's single
 * global MoFREAKUtilities object is, main.cpp:47).  All device work is queued on the context's own
 * non-blocking CUDA stream; functions documented "async" return before it completes.  There is no
 * CPU fallback: if no CUDA device is usable mfk_create() fails.
 *
 * Fixed algorithm constants (literals in the reference): frame gap 5 (MoFREAKUtilities.cpp:378),
 * BRISK threshold 30 and 3 octaves (:420-421, brisk.h:293), FREAK defaults (:427), 8 + 8 descriptor
 * bytes (MoFREAKUtilities.h:20-21), MIP theta 288 on 19x19 patches (:48, :300-316).
 */
#ifndef MOFREAK_B200_H_
#define MOFREAK_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFK_OK              0
#define MFK_ERR_ARG        -1   /* bad argument */
#define MFK_ERR_CUDA       -2   /* a CUDA call failed (message has the CUDA error string) */
#define MFK_ERR_CAPACITY   -3   /* a CALLER buffer is too small; the call says how much is needed, device state is kept */
#define MFK_ERR_NOMEM      -4
#define MFK_ERR_STATE      -5   /* call out of order (e.g. fetch before extract) */
#define MFK_ERR_RANGE      -6   /* a value the device formatter does not cover (see mfk_rows_text) */
#define MFK_ERR_BUDGET     -7   /* a DEVICE per-pair candidate/row budget overflowed: results of the run are void and
                                   no sizes are reported; recreate the context with mfk_create_ex and larger budgets */

#define MFK_FRAME_GAP       5

typedef struct mfk_ctx mfk_ctx;

/* One .mofreak row: the payload of struct MoFREAKFeature (MoFREAKUtilities.h:23-53) as written by
 * writeMoFREAKFeaturesToFile (MoFREAKUtilities.cpp:691-719): x y frame scale [0 0] a0..a7 m0..m7.
 * motion_x / motion_y are constant 0 in the reference (:476-477) and are not stored. 32 bytes. */
typedef struct mfk_row {
    float   x, y;
    int32_t frame;
    float   scale;
    uint8_t appearance[8];
    uint8_t motion[8];
} mfk_row;

/* ---- lifetime ---------------------------------------------------------------------------------- */

/* Creates a context for width x height gray frames on CUDA device `device`.
 * store_frames: capacity of the device-resident frame store (the "frame ring"), in frames.
 * max_pairs:    how many (frame t, frame t-5) pairs one internal launch sequence processes.
 * Replaces: new MoFREAKUtilities(dataset) + the per-frame new BriskFeatureDetector(30) / cv::FREAK
 * construction of MoFREAKUtilities.cpp:420-427 (all tables are built once here).
 * Device memory: the store plus four sets ("lanes") of per-launch-sequence buffers, about 11 MB per pair and set at
 * 720x576 (27 B per frame pixel), i.e. ~11 GB for max_pairs = 256; the environment variable MFK_LANES = 1..8 sets
 * the number of sets (default 4; fewer = less memory and less overlap between consecutive launch sequences; more measured no gain). */
int  mfk_create(mfk_ctx** out, int device, int width, int height, int store_frames, int max_pairs);
/* Same, with explicit per-pair budgets (0 = default) for detector candidates and output rows. */
int  mfk_create_ex(mfk_ctx** out, int device, int width, int height, int store_frames, int max_pairs,
                   int cand_cap_per_pair, int row_cap_per_pair);
void mfk_destroy(mfk_ctx* ctx);
const char* mfk_last_error(const mfk_ctx* ctx);   /* ctx may be NULL: error of the last failed mfk_create */
int  mfk_sync(mfk_ctx* ctx);                      /* waits for the context's stream */

/* ---- frame store (the ring that keeps t and t-5 resident in HBM) ------------------------------------ */

/* Copies n_frames gray u8 images (row pitch `stride` bytes, >= width) from HOST memory into store
 * slots slot0 .. slot0+n_frames-1 (indices taken modulo store_frames).  Async when `gray` is pinned.
 * Replaces: frame_queue.push(frame.clone()) of MoFREAKUtilities.cpp:391-399, 485-487. */
int  mfk_upload_frames(mfk_ctx* ctx, const uint8_t* gray, size_t stride, int n_frames, int slot0);
/* Page-locked HOST memory for frame and row buffers: uploads from it are true asynchronous DMA (pageable memory is
 * staged by the driver and stalls the caller).  write_combined != 0 asks for write-combined pages: faster for the
 * device to read, slow for the CPU to read back - for buffers the host only fills (decoded frames).  Portable across
 * contexts/devices.  No counterpart in the reference (its frames are cv::Mat in pageable memory). */
int  mfk_host_alloc(void** ptr, size_t bytes, int write_combined);
int  mfk_host_free(void* ptr);
/* Frame ingest from interleaved 8-bit BGR: uploads n_frames frames (stride = bytes per source row,
 * >= 3 * width) and converts them to gray ON THE DEVICE into slots slot0.. exactly as cv::cvtColor(CV_BGR2GRAY)
 * does for the reference (MoFREAKUtilities.cpp:395, 410): (B*cb + G*cg + R*cr + half) >> bits.  coeff_bits = 14
 * selects OpenCV 2.4's coefficients 1868/9617/4899 (the reference's version; recalled, not verifiable offline),
 * 15 those of OpenCV >= 3 (3735/19235/9798; verified against cv2 4.13).  Async, like mfk_upload_frames. */
int  mfk_upload_frames_bgr(mfk_ctx* ctx, const uint8_t* bgr, size_t stride, int n_frames, int slot0, int coeff_bits);
/* Device address of slot 0 and the store capacity, for producers that already live on the GPU
 * (decoder output, synthetic generators).  Slots are dense: slot k starts at base + k*width*height. */
int  mfk_frame_store(mfk_ctx* ctx, uint8_t** device_base, int* store_frames);

/* ---- extraction --------------------------------------------------------------------------------- */

/* Queues extraction for n_pairs frame pairs.  Pair i compares store slot cur_slot[i] with
 * prev_slot[i] and labels its rows label[i] (the reference labels frame c of a clip c-1).
 * Rows accumulate on the device in pair order, each pair's rows in the reference's keypoint order
 * (BRISK layer-major, raster within a layer, minus FREAK's border drops).  Async.
 * Replaces the loop body of MoFREAKUtilities::computeMoFREAKFromFile, MoFREAKUtilities.cpp:413-483:
 * absdiff, BriskFeatureDetector::detect (brisk.cpp:549-704), cv::FREAK::compute,
 * extractMotionByMotionInterchangePatterns (:288-325) and row assembly (:447-480).
 * Slots must be >= 0 (taken modulo store_frames); MFK_ERR_ARG otherwise. */
int  mfk_extract_pairs(mfk_ctx* ctx, const int32_t* cur_slot, const int32_t* prev_slot,
                       const int32_t* label, int n_pairs);
/* Waits, then copies the rows accumulated since the last fetch/reset to HOST memory.
 * rows_per_pair (optional, may be NULL) receives one count per queued pair, in queue order, at most
 * pairs_cap entries.  *n_rows is the total; if it exceeds cap only cap rows are copied and
 * MFK_ERR_CAPACITY is returned.  MFK_ERR_BUDGET if a device budget overflowed (rows incomplete).  Resets the accumulation. */
int  mfk_fetch_rows(mfk_ctx* ctx, mfk_row* rows, size_t cap, size_t* n_rows,
                    uint32_t* rows_per_pair, size_t pairs_cap);
/* Device-side view of the accumulated rows (valid until the next extract/fetch/reset); syncs. */
int  mfk_rows_device(mfk_ctx* ctx, const mfk_row** device_rows, size_t* n_rows);
/* Drops the accumulated rows.  Does not wait: the reset takes effect in stream order just before the next launch
 * sequence appends its rows (so the front of that sequence may already run under the tail of the previous one), or
 * at the next call that reads the accumulation.  Rows fetched earlier are unaffected. */
int  mfk_reset_rows(mfk_ctx* ctx);

/* Whole clip from HOST frames, the shape of one computeMoFREAKFromFile call
 * (MoFREAKUtilities.cpp:374-498): uploads n_frames frames, processes frames 5..n-1 against c-5 with
 * labels c-1, fetches the rows.  Includes H2D and D2H; this is the end-to-end path. */
int  mfk_extract_clip(mfk_ctx* ctx, const uint8_t* gray, size_t stride, int n_frames,
                      mfk_row* rows, size_t cap, size_t* n_rows);

/* ---- `.mofreak` text ------------------------------------------------------------------------------- */

/* The rows accumulated since the last fetch/reset as `.mofreak` text in HOST memory, byte for byte what
 * MoFREAKUtilities::writeMoFREAKFeaturesToFile (MoFREAKUtilities.cpp:691-719) writes for them: per row
 * "x y frame scale 0 0 a0..a7 m0..m7", each token followed by one space, then '\n', floats as the default
 * ostream operator<< prints them ("%g").  Formatted on the device; does NOT reset the accumulation.
 * *n_bytes is the length of the text; if it exceeds cap nothing is copied and MFK_ERR_CAPACITY is returned
 * (call with text = NULL, cap = 0 to size a buffer).  MFK_ERR_RANGE if a row holds a float outside
 * 0 | [1e-4, 1e6) (exponent notation; never the case for extracted rows) - format such rows on the host. */
int  mfk_rows_text(mfk_ctx* ctx, char* text, size_t cap, size_t* n_bytes);
/* The same for n caller-supplied HOST rows (e.g. a binary row file turned back into text). */
int  mfk_format_rows(mfk_ctx* ctx, const mfk_row* rows, size_t n, char* text, size_t cap, size_t* n_bytes);
/* mfk_extract_clip that also returns the clip's `.mofreak` text: one call = one computeMoFREAKFromFile
 * (MoFREAKUtilities.cpp:374-498 incl. the write at :490-491).  rows may be NULL when only the text is wanted.
 * If the text or the rows do not fit, *n_bytes / *n_rows hold the sizes needed, MFK_ERR_CAPACITY is returned and
 * the accumulation is KEPT: read it with mfk_rows_text + mfk_fetch_rows into larger buffers.  Otherwise resets.
 * MFK_ERR_BUDGET (a device budget overflowed): *n_bytes = *n_rows = 0, the accumulation is dropped; redo the clip on a
 * context with larger budgets (mfk_create_ex). */
int  mfk_extract_clip_text(mfk_ctx* ctx, const uint8_t* gray, size_t stride, int n_frames,
                           char* text, size_t text_cap, size_t* n_bytes,
                           mfk_row* rows, size_t cap, size_t* n_rows);

/* ---- MIP alone (the MotionInterchangePatterns facade) ------------------------------------------------ */

/* n keypoints on one frame pair held in store slots cur_slot / prev_slot: kp_xys = n x {x, y, size}
 * (HOST floats; x, y truncated to int exactly as at MoFREAKUtilities.cpp:460).  motion = n x 8 HOST bytes.
 * Replaces MoFREAKUtilities::extractMotionByMotionInterchangePatterns, MoFREAKUtilities.cpp:288-325. */
int  mfk_mip_describe(mfk_ctx* ctx, int cur_slot, int prev_slot, const float* kp_xys, int n, uint8_t* motion);

/* ---- bag of words ----------------------------------------------------------------------------------- */

/* codewords: K x 16 HOST bytes (8 appearance then 8 motion).  Replaces the in-memory result of
 * BagOfWordsRepresentation::loadClusters, BagOfWordsRepresentation.cpp:140-173. */
int  mfk_bow_set_codebook(mfk_ctx* ctx, const uint8_t* codewords, int K);
/* desc: N x 16 HOST bytes -> idx: N HOST int32, arg-min Hamming distance, lowest index on ties.
 * Replaces BagOfWordsRepresentation::bruteForceMatch / hammingDistance, :22-72. */
int  mfk_bow_assign(mfk_ctx* ctx, const uint8_t* desc, size_t N, int32_t* idx);
/* Same on DEVICE buffers (async).  d_desc must be 16-byte aligned (descriptors are read as 128-bit words). */
int  mfk_bow_assign_device(mfk_ctx* ctx, const uint8_t* d_desc, size_t N, int32_t* d_idx);
/* Assigns the rows currently accumulated on the device (descriptor = appearance|motion) and adds
 * the per-word counts into counts[K] (HOST uint32, overwritten).  The float normalisation of
 * buildHistogram (:125-135) is count / total on the host.  Replaces buildHistogram, :74-138. */
int  mfk_bow_histogram_rows(mfk_ctx* ctx, uint32_t* counts);
/* desc: N x 16 HOST bytes -> counts[K] HOST uint32 (overwritten). */
int  mfk_bow_histogram(mfk_ctx* ctx, const uint8_t* desc, size_t N, uint32_t* counts);
/* Optional exchange step when one video's rows live on several GPUs: sums d_counts[n] (DEVICE uint32)
 * over all ranks of `nccl_comm` (an ncclComm_t) on the context's stream.  Async.  Returns
 * MFK_ERR_STATE when the library was built without NCCL.  No counterpart in the reference. */
int  mfk_allreduce_counts(mfk_ctx* ctx, uint32_t* d_counts, size_t n, void* nccl_comm);
/* One word-count vector per VIDEO from the rows accumulated on the device - what the reference computes per `.mofreak`
 * file (BagOfWordsRepresentation::buildHistogram, :74-138, called once per file from convertFileToBOWFeature :960-1031).
 * The pairs queued since the last reset are split into n_videos consecutive groups of pairs_per_video[v] pairs
 * (NULL with n_videos == 1: all of them); counts = n_videos x K HOST uint32, overwritten.
 * nccl_comm != NULL: the videos' frame ranges are sharded over the ranks of that communicator (every rank queued ITS
 * frames of the same n_videos videos); the count vectors are summed over the ranks (ncclAllReduce, uint32 sum, over
 * NVLink/NVSwitch) and every rank receives the totals.  Collective then: every rank must call it.  Assignment kernel,
 * collective and read-back are queued back to back on the context's stream; the host waits once. */
int  mfk_bow_histogram_videos(mfk_ctx* ctx, const uint32_t* pairs_per_video, int n_videos, uint32_t* counts, void* nccl_comm);

/* Helpers to build the communicator for mfk_allreduce_counts without linking NCCL (same run-time resolution): rank 0 calls
 * mfk_nccl_unique_id and ships the 128 bytes to the other ranks by any means (torch.distributed, MPI, a file); every rank
 * then calls mfk_nccl_comm_init with its context's device current. */
int  mfk_nccl_unique_id(uint8_t id_out[128]);
int  mfk_nccl_comm_init(mfk_ctx* ctx, void** nccl_comm_out, int n_ranks, int rank, const uint8_t id[128]);
int  mfk_nccl_comm_destroy(void* nccl_comm);

/* ---- utilities ---------------------------------------------------------------------------------------- */

/* Number of kernel launches issued by this context since creation. */
uint64_t mfk_launch_count(const mfk_ctx* ctx);
/* CUDA-event stopwatch on the context's streams (the reference times a stage with one clock() pair, main.cpp:929-1011). */
int  mfk_timer_start(mfk_ctx* ctx);
int  mfk_timer_stop(mfk_ctx* ctx, float* elapsed_ms);   /* syncs */
const char* mfk_version(void);

/* Per-kernel accounting, stage dumps for parity tests and the synthetic clip generator are declared in
 * mofreak_b200_debug.h: they are test and measurement instruments, not part of the drop-in surface. */

#ifdef __cplusplus
}
#endif
#endif /* MOFREAK_B200_H_ */
