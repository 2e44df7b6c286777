// Sparse (per-candidate) stage: BriskScaleSpace::getKeypoints, reference brisk.cpp:641-703, with isMax2D
// (:838-934) and refine3D (:937-1103), restructured so that it is order independent.
//
// The reference walks candidates sequentially; its lazily memoised score map makes the tie-break of isMax2D
// depend on which scores earlier candidates happened to touch.  Here the raw map it would hold is modelled as
//     raw(q) = S(q)                       if S(q) >= threshold (written by getAgastPoints), or
//                                            q was read by a getScoreMaxAbove scan from the layer below, or
//                                            q lies in the 3x3 patch of a raster-earlier candidate that survived
//              0                           otherwise
// (validated against the reference's object code through oracle/brisk_port.cpp, which has this same structure).
//   phase A  one thread per candidate: 8-neighbour maximum test; candidates without an equal neighbour are
//            decided completely (scans above/below, refinement, keypoint record);
//   phase B  one block per pair: the few candidates with ties, layer by layer, by a fixed point whose only
//            ordering requirement is "every raster-earlier undecided tie candidate within 3 px is final".
#include <stdio.h>
#include <stdlib.h>

#include "sparse_device.cuh"

namespace mfk {

// ---------------------------------------------------------------------------------------------------
// phase A
// ---------------------------------------------------------------------------------------------------
// Four kernels, re-packing the survivors device-wide between them so that every expensive step runs on full warps
// and no warp waits at a block barrier for a slower one:
//   k_pa_test   every candidate against its 8 neighbours (3 aligned row loads); losers are final, candidates with an
//               equal neighbour go to their pair's tie list, the rest to queue 0;
//   k_pa_above  scan of the layer above (stage_above) -> queue 1;   k_pa_below  scan of the layer below -> queue 2;
//   k_pa_emit   refinement and keypoint record (stage_emit).
// Queues are unordered (warp-aggregated atomic appends): candidates are independent of each other in phase A.
constexpr int PA_THREADS = 128;

// all 32 lanes must call; returns the slot for lanes with pred
__device__ __forceinline__ uint32_t warp_append(uint32_t* counter, bool pred) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (!m) return 0;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(PA_THREADS) k_pa_test(const __grid_constant__ Geom g, Maps m, CandBuf cb) {
    const uint32_t n = min(cb.row_count[g.F * g.rows], (uint32_t)g.cand_cap);
    const uint32_t n_round = (n + 31u) & ~31u;
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n_round; c += gridDim.x * blockDim.x) {
        bool pass = false;
        if (c < n) {
            const uint32_t xy = cb.xy[c], meta = cb.meta[c];
            const int x = xy & 0xffff, y = xy >> 16, li = meta & 7, f = meta >> 3;
            const LayerGeom& L = g.L[li];
            const size_t off = L.base + (size_t)f * L.px;
            const uint8_t* S = m.S + off + y * L.w + x;
            // isMax2D, order-independent part (brisk.cpp:838-869): neighbours below the threshold can never exceed
            // or equal a candidate, so the dense map can be read directly
            const uint32_t r0 = load4(S - L.w - 1), r1 = load4(S - 1), r2 = load4(S + L.w - 1);
            const int c0 = (r1 >> 8) & 0xff;
            const uint32_t m01 = __vmaxu4(r0 & 0xffffffu, r2 & 0xffffffu);          // rows above and below, 3 cells each
            const uint32_t m02 = __vmaxu4(m01, r1 & 0xff00ffu);                      // left and right neighbour
            const int mx = max(max((int)(m02 & 0xff), (int)((m02 >> 8) & 0xff)), (int)((m02 >> 16) & 0xff));
            uint8_t* status = m.status + off + y * L.w + x;
            if (mx > c0) { *status = 0; cb.flags[c] = 0; }
            else if (mx == c0) {
                *status = ST_PENDING; cb.flags[c] = ST_PENDING;
                const uint32_t slot = atomicAdd(&cb.pend_count[f * N_LAYERS + li], 1u);
                if (slot < PEND_CAP) cb.pend_list[(size_t)(f * N_LAYERS + li) * PEND_CAP + slot] = c;
            } else pass = true;
        }
        const uint32_t slot = warp_append(&cb.q_count[0], pass);
        if (pass) { if (slot < cb.q_cap) cb.q0[slot] = c; else atomicOr(&cb.live[1], 1u); }
    }
}

__global__ void __launch_bounds__(PA_THREADS, 10) k_pa_above(const __grid_constant__ Geom g, Maps m, CandBuf cb) {
    const uint32_t n = min(cb.q_count[0], cb.q_cap);
    const uint32_t n_round = (n + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        bool pass = false;
        uint32_t c = 0;
        int li = 0;
        Scan up = neutral_scan();
        if (i < n) {
            c = cb.q0[i];
            const uint32_t xy = cb.xy[c], meta = cb.meta[c];
            PairView pv; pv.g = &g; pv.m = m; pv.f = meta >> 3;
            li = meta & 7;
            pass = stage_above(pv, cb, li, xy & 0xffff, xy >> 16, c, up);
        }
        // layer-0 candidates have no layer below (the AGAST-5/8 stand-in is part of the emit step): they go straight to the
        // emit queue, so that k_pa_below runs on full warps of candidates that really scan
        const bool to_emit = pass && li == 0, to_below = pass && li != 0;
        const uint32_t slot = warp_append(&cb.q_count[1], to_below);
        if (to_below) {
            if (slot < cb.q_cap) { cb.q1[slot] = c; cb.q1_up[3 * (size_t)slot] = up.max; cb.q1_up[3 * (size_t)slot + 1] = up.dx; cb.q1_up[3 * (size_t)slot + 2] = up.dy; }
            else atomicOr(&cb.live[1], 1u);
        }
        const uint32_t slot2 = warp_append(&cb.q_count[2], to_emit);
        if (to_emit) {
            if (slot2 < cb.q_cap) {
                cb.q2[slot2] = c;
                float* o = cb.q2_scan + 6 * (size_t)slot2;
                o[0] = up.max; o[1] = up.dx; o[2] = up.dy; o[3] = 0.f; o[4] = 0.f; o[5] = 0.f;      // neutral_scan() below
            } else atomicOr(&cb.live[1], 1u);
        }
    }
}

__global__ void __launch_bounds__(PA_THREADS, 10) k_pa_below(const __grid_constant__ Geom g, Maps m, CandBuf cb) {
    const uint32_t n = min(cb.q_count[1], cb.q_cap);
    const uint32_t n_round = (n + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        bool pass = false;
        uint32_t c = 0;
        Scan dn = neutral_scan();
        if (i < n) {
            c = cb.q1[i];
            const uint32_t xy = cb.xy[c], meta = cb.meta[c];
            PairView pv; pv.g = &g; pv.m = m; pv.f = meta >> 3;
            pass = stage_below(pv, cb, meta & 7, xy & 0xffff, xy >> 16, c, dn);
        }
        const uint32_t slot = warp_append(&cb.q_count[2], pass);
        if (pass) {
            if (slot < cb.q_cap) {
                cb.q2[slot] = c;
                float* o = cb.q2_scan + 6 * (size_t)slot;
                const float* u = cb.q1_up + 3 * (size_t)i;
                o[0] = u[0]; o[1] = u[1]; o[2] = u[2]; o[3] = dn.max; o[4] = dn.dx; o[5] = dn.dy;
            } else atomicOr(&cb.live[1], 1u);
        }
    }
}

__global__ void __launch_bounds__(PA_THREADS) k_pa_emit(const __grid_constant__ Geom g, Maps m, CandBuf cb, DeviceTables tb) {
    const uint32_t n = min(cb.q_count[2], cb.q_cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t c = cb.q2[i];
        const float* o = cb.q2_scan + 6 * (size_t)i;
        Scan up = neutral_scan(), dn = neutral_scan();
        up.max = o[0]; up.dx = o[1]; up.dy = o[2]; dn.max = o[3]; dn.dx = o[4]; dn.dy = o[5];
        const uint32_t xy = cb.xy[c], meta = cb.meta[c];
        PairView pv; pv.g = &g; pv.m = m; pv.f = meta >> 3;
        stage_emit(pv, cb, tb, meta & 7, xy & 0xffff, xy >> 16, c, up, dn);
    }
}

// blocks per SM of the four kernels (MFK_PA_BLOCKS="t,a,b,e" overrides, for measurements)
static const int* pa_blocks() {
    static int b[4] = {8, 8, 8, 8};
    static bool parsed = false;
    if (!parsed) {
        parsed = true;
        if (const char* e = getenv("MFK_PA_BLOCKS")) {
            int v[4];
            if (sscanf(e, "%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3]) == 4)
                for (int i = 0; i < 4; i++) if (v[i] >= 1 && v[i] <= 16) b[i] = v[i];
        }
    }
    return b;
}

void launch_phase_a(cudaStream_t st, const Geom& g, Maps m, CandBuf cb, DeviceTables tb) {
    const int* b = pa_blocks();
    k_pa_test<<<148 * b[0], PA_THREADS, 0, st>>>(g, m, cb);
    k_pa_above<<<148 * b[1], PA_THREADS, 0, st>>>(g, m, cb);
    k_pa_below<<<148 * b[2], PA_THREADS, 0, st>>>(g, m, cb);
    k_pa_emit<<<148 * b[3], PA_THREADS, 0, st>>>(g, m, cb, tb);
}

// ---------------------------------------------------------------------------------------------------
// phase B: a group of G threads per pair, 256 / G pairs per block
// ---------------------------------------------------------------------------------------------------
// The tie candidates of a pair are few (about pyramid pixels / 1000: ~36 at 160x120, ~760 at 720x576) and their fixed
// point is a chain of short rounds, i.e. latency: what a pair needs is cheap barriers and many pairs in flight, not
// threads.  So the group size follows the frame size - one WARP per pair for small frames (__syncwarp, eight pairs per
// block), named barriers for 64 / 128 threads, the whole block (__syncthreads) for large frames.
constexpr int PB_THREADS = 256;
#ifndef PB_BLOCKS_PER_SM
#define PB_BLOCKS_PER_SM 3
#endif
constexpr int PB_LIST = PEND_CAP;  // tie candidates of one (pair, layer) staged in shared memory per sweep, for G = 256

template <int G>
__device__ __forceinline__ void group_sync(int grp) {
    if (G == 32) __syncwarp();
    else if (G == PB_THREADS) __syncthreads();
    else asm volatile("bar.sync %0, %1;" :: "r"(grp + 1), "r"(G) : "memory");
}

// Shared memory of one block (dynamic: 27 bytes per list slot, PB_LIST slots shared by the block's pairs).  The candidates'
// coordinates and, for the winners of a round, their own scores travel with the lists, so that a round costs four dependent
// global-memory round trips (tie window; scan above; scan below; own patch) instead of seven.
//
// The groups of a block advance IN STEP: block-wide barriers align the phases of a round (verdicts | re-packing | scans of the
// winners), and a group without work in a round waits.  The kernel is 6 K instructions (~100 KB) and every group follows its own
// pair; left to themselves the warps of an SM sit at different places of that code and the kernel stalls on instruction fetch
// (ncu, one warp per pair, free-running: 74 % of the stall samples "no instruction").  In step, the warps of a block execute
// the same code region at the same time.
template <int G>
__global__ void __launch_bounds__(PB_THREADS, PB_BLOCKS_PER_SM) k_phase_b(const __grid_constant__ Geom g, Maps m, CandBuf cb, DeviceTables tb) {
    constexpr int PER = PB_THREADS / G;          // pairs per block
    constexpr int LIST = PB_LIST / PER;          // list capacity of one pair
    extern __shared__ __align__(16) uint32_t pb_smem[];
    __shared__ int ctl[PER][4];                  // n_list, n_next, n_pass, more
    __shared__ int warp_cnt[PB_THREADS / 32];
    const int grp = threadIdx.x / G, t = threadIdx.x % G;
    const int f = blockIdx.x * PER + grp;
    const bool active = f < g.F;                 // groups beyond the last pair only keep the block's barriers company
    // per pair: list[2][LIST] | list_xy[2][LIST] | passed[LIST] | passed_xy[LIST] | verdict[LIST] | center[LIST] | passed_center[LIST]
    uint32_t* mine = pb_smem + (size_t)grp * (6 * LIST + 3 * LIST / 4);
    uint32_t (*list)[LIST] = reinterpret_cast<uint32_t (*)[LIST]>(mine);            // undecided tie candidates of this round / of the next
    uint32_t (*lxy)[LIST] = reinterpret_cast<uint32_t (*)[LIST]>(mine + 2 * LIST);  // their x | y << 16
    uint32_t* passed = mine + 4 * LIST;           // this round's winners, re-packed so that the scans run on full warps
    uint32_t* passed_xy = mine + 5 * LIST;
    uint8_t* verdict = reinterpret_cast<uint8_t*>(mine + 6 * LIST);
    uint8_t* center = verdict + LIST;
    uint8_t* passed_center = center + LIST;
    volatile int& n_list = ctl[grp][0];
    volatile int& n_next = ctl[grp][1];
    volatile int& n_pass = ctl[grp][2];
    volatile int& more = ctl[grp][3];
    int* const wc = warp_cnt + grp * (G / 32);
    PairView pv; pv.g = &g; pv.m = m; pv.f = active ? f : 0;
    const uint32_t cap = (uint32_t)g.cand_cap;
    // which layers have tie candidates at all: one load for the six counts instead of a dependent round trip per layer
    uint32_t pending_of[N_LAYERS];
#pragma unroll
    for (int li = 0; li < N_LAYERS; li++) pending_of[li] = active ? cb.pend_count[pv.f * N_LAYERS + li] : 0u;
    for (int li = 0; li < N_LAYERS; li++) {
        const uint32_t n_pending = pending_of[li];
        if (!__syncthreads_or(n_pending != 0)) continue;                 // no pair of this block has ties on this layer
        const LayerGeom& L = g.L[li];
        const size_t r0 = (size_t)pv.f * g.rows + L.row0;
        uint32_t cursor = 0, end = 0;
        if (n_pending) { cursor = min(cb.row_count[r0], cap); end = min(cb.row_count[r0 + L.h], cap); }
        const bool listed = n_pending <= (uint32_t)min((int)PEND_CAP, LIST);      // phase A's list holds them all and they fit (order is irrelevant then)
        bool chunks_left = n_pending != 0;       // this group still has a chunk of tie candidates to stage
        int nl = 0, cur = 0;
        for (;;) {
            // ---- a group whose list is empty stages its next chunk: candidates are in raster order, so taking the pending
            // ones chunk by chunk keeps every dependency (always raster-earlier) inside the current or a finished chunk ----
            if (nl == 0 && chunks_left) {
                if (t == 0) { n_list = 0; more = 0; }
                group_sync<G>(grp);
                uint32_t next_cursor = end;
                if (listed) {
                    const uint32_t* src = cb.pend_list + (size_t)(pv.f * N_LAYERS + li) * PEND_CAP;
                    for (uint32_t i = t; i < n_pending; i += G) { const uint32_t c = src[i]; list[0][i] = c; lxy[0][i] = cb.xy[c]; }
                    if (t == 0) n_list = (int)n_pending;
                } else
                // gather up to LIST pending candidates starting at cursor, preserving order
                for (uint32_t base = cursor; base < end; base += G) {
                    const uint32_t c = base + t;
                    const bool pend = c < end && (cb.flags[c] & ST_PENDING);
                    // ordered append: warp ballots + per-warp offsets
                    const unsigned bal = __ballot_sync(0xffffffffu, pend);
                    if ((t & 31) == 0) wc[t >> 5] = __popc(bal);
                    group_sync<G>(grp);
                    int before = n_list;
                    for (int w = 0; w < (int)(t >> 5); w++) before += wc[w];
                    int total = 0;
                    for (int w = 0; w < G / 32; w++) total += wc[w];
                    const int slot = before + __popc(bal & ((1u << (t & 31)) - 1u));
                    const bool fits = n_list + total <= LIST;
                    if (pend && fits) { list[0][slot] = c; lxy[0][slot] = cb.xy[c]; }
                    group_sync<G>(grp);
                    if (!fits) { next_cursor = base; if (t == 0) more = 1; break; }
                    if (t == 0) n_list += total;
                    group_sync<G>(grp);
                }
                group_sync<G>(grp);
                nl = n_list; cur = 0;
                cursor = next_cursor;
                chunks_left = more != 0;
                group_sync<G>(grp);              // `more` and n_list were read by every thread before thread 0 may clear them again
#ifdef MFK_PB_STATS
                if (t == 0) atomicAdd(&cb.counters[5], (uint32_t)nl);
#endif
            }
            if (!__syncthreads_or(nl > 0)) break;                        // every pair of the block is done with this layer
            // ---- one round of the fixed point: every candidate whose raster-earlier tie neighbours are final (from the
            // previous round's statuses) is decided; winners and still-waiting ones are re-packed; the winners' scans run ----
#ifdef MFK_PB_STATS
            if (t == 0 && nl > 0) { atomicAdd(&cb.counters[4], 1u); atomicAdd(&cb.counters[7], (uint32_t)nl); }
#endif
            if (t == 0) { n_next = 0; n_pass = 0; }
            for (int i = t; i < nl; i += G) {
                const uint32_t xy = lxy[cur][i];
                TieWindow tw;
                tw.load(pv, li, xy & 0xffff, xy >> 16);
                verdict[i] = earlier_pending_nearby(tw) ? 0 : (tie_break(tw) ? 2 : 1);
                center[i] = (uint8_t)packed_byte(tw.S[3], 3);
            }
            __syncthreads();
            for (int i = t; i < nl; i += G) {
                const uint32_t c = list[cur][i], xy = lxy[cur][i];
                const uint8_t v = verdict[i];
                if (v == 0) { const int k = atomicAdd(&ctl[grp][1], 1); list[cur ^ 1][k] = c; lxy[cur ^ 1][k] = xy; }   // still waiting (order is irrelevant)
                else if (v == 2) { const int k = atomicAdd(&ctl[grp][2], 1); passed[k] = c; passed_xy[k] = xy; passed_center[k] = center[i]; }
                else { pv.m.status[pv.off(li) + (xy >> 16) * L.w + (xy & 0xffff)] = 0; cb.flags[c] = 0; }
            }
            __syncthreads();
            const int np = nl > 0 ? n_pass : 0;
            for (int i = t; i < np; i += G) {
                const uint32_t xy = passed_xy[i];
                after_max2d(pv, cb, tb, li, xy & 0xffff, xy >> 16, passed[i], passed_center[i]);
            }
            __threadfence_block();               // statuses and seen marks written by this group are read by it in the next round / layer
            __syncthreads();
            if (nl > 0) { nl = n_next; cur ^= 1; }
            __syncthreads();                     // n_next was read before thread 0 clears it
        }
    }
}

int phase_b_group(const Geom& g) {
    if (const char* e = getenv("MFK_PB_GROUP")) { const int v = atoi(e); if (v == 32 || v == 64 || v == 128 || v == 256) return v; }
    // Measured in the pipelined bench (four launch sequences in flight): small groups win at every frame size, although a
    // pair then takes longer on its own - 720x576: G = 256 106.2 k frames/s, 128 112.4 k, 64 116.5 k, 32 116.3 k; 320x240: 64
    // 541 k, 32 572 k; 160x120: 64 1.83 M, 32 2.04 M.  The kernel waits on memory most of the time; fewer resident warps
    // per pair leave the SMs' registers and issue slots to the dense kernels of the other launch sequences.
    const int ties = g.pyr_px / 1000;            // expected tie candidates per pair
    return ties <= 400 ? 32 : 64;
}

template <int G>
static void launch_phase_b_g(cudaStream_t st, const Geom& g, Maps m, CandBuf cb, DeviceTables tb) {
    constexpr int PER = PB_THREADS / G;
    const size_t smem = (size_t)PB_LIST * 27;                              // 6 words + 3 bytes per list slot
    cudaFuncSetAttribute(k_phase_b<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     // per device: set on every launch
    k_phase_b<G><<<(g.F + PER - 1) / PER, PB_THREADS, smem, st>>>(g, m, cb, tb);
}

void launch_phase_b(cudaStream_t st, const Geom& g, Maps m, CandBuf cb, DeviceTables tb) {
    switch (phase_b_group(g)) {
        case 32:  launch_phase_b_g<32>(st, g, m, cb, tb); break;
        case 64:  launch_phase_b_g<64>(st, g, m, cb, tb); break;
        case 128: launch_phase_b_g<128>(st, g, m, cb, tb); break;
        default:  launch_phase_b_g<256>(st, g, m, cb, tb); break;
    }
}

// ---------------------------------------------------------------------------------------------------
// keep flags -> scan input
// ---------------------------------------------------------------------------------------------------
__global__ void k_keep_flags(Geom g, CandBuf cb) {
    const uint32_t n = min(cb.row_count[g.F * g.rows], (uint32_t)g.cand_cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) cb.live[0] = n;                       // live length for the scan
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        cb.keep_scan[i] = (cb.flags[i] & ST_KEEP) ? 1u : 0u;
}

void launch_keep_flags(cudaStream_t st, const Geom& g, CandBuf cb) {
    k_keep_flags<<<148 * 8, 256, 0, st>>>(g, cb);
}

}  // namespace mfk

namespace mfk {

// after the scan of the keep flags: row r of the batch comes from candidate row_src[r]; rows of pair f are the
// kept candidates between the pair's first and one-past-last candidate.  One thread also performs the one serialised step
// between batches: this batch's base in the shared row accumulation, counters advanced (nothing in this kernel reads what
// that step writes; the descriptor kernel, launched next, does).
__global__ void k_scatter_rows(Geom g, CandBuf cb, uint32_t* __restrict__ pair_rows) {
    const uint32_t n = cb.live[0];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    if (tid == 0) {
        const uint32_t rows = cb.keep_scan[n];
        const uint32_t have = cb.counters[1];
        cb.live[3] = have;              // rows_base of this batch for k_describe
        cb.live[2] = 0;                 // fresh row tickets for k_describe
        // overflow seen by this batch's lane-private kernels joins the shared flags here, behind any row reset
        if (cb.live[1]) { cb.counters[0] |= cb.live[1]; cb.live[1] = 0; }
        if (have + rows > (uint32_t)g.row_cap) { cb.counters[0] |= 2u; cb.counters[1] = (uint32_t)g.row_cap; }
        else cb.counters[1] = have + rows;
        cb.counters[2] += (uint32_t)g.F;
        cb.counters[3] = n;
    }
    for (uint32_t i = tid; i < n; i += nthreads)
        if (cb.flags[i] & ST_KEEP) cb.row_src[cb.keep_scan[i]] = i;
    for (uint32_t i = tid; i < (uint32_t)g.F; i += nthreads) {
        const uint32_t b = min(cb.row_count[(size_t)i * g.rows], n), e = min(cb.row_count[(size_t)(i + 1) * g.rows], n);
        pair_rows[i] = cb.keep_scan[e] - cb.keep_scan[b];
    }
}

void launch_scatter_rows(cudaStream_t st, const Geom& g, CandBuf cb, uint32_t* pair_rows) {
    k_scatter_rows<<<148 * 8, 256, 0, st>>>(g, cb, pair_rows);
}

}  // namespace mfk
