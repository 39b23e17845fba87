// KF, TMA + warp-specialised variant of fused.cu for level planes whose rows are 16-byte aligned
// (h*w % 4 == 0): same arithmetic and outputs (shared finish_tile), different data movement.
//
// The register-resident kernel of fused.cu is latency-bound (0.74 of the HBM copy rate): a warp's
// loads, its compare/select chains and the epilogue's global round trips (slot atomics, box / corner
// reads of the survivors) all sit in one instruction stream.  Here they are three concurrent roles of
// one persistent CTA per SM, over tiles of 32 positions ([277 class channels][32] = 35,456 B) in a
// five-deep shared-memory ring:
//   warps 21-22, lane 0  producers (each owns alternate ring slots): eight 3-D TMA box loads per tile
//                        (UTMALDG.3D, one per class tensor: all its channels x 32 positions of image
//                        b), completion on the slot's `full` mbarrier, reuse gated by `empty`.  A
//                        thread needs ~100 cycles per UTMALDG, hence two issuers;
//   warps 0-15           scanners, warp w owns half (w & 1) of class group w >> 1: lanes along
//                        positions (conflict-free LDS), running maximum as an FMNMX chain with the
//                        first-index / runner-up selects hanging off it, two interleaved chains;
//                        leaves (maximum logit, its first index, runner-up before it) per position
//                        in a small exchange buffer and bar.arrive's (non-blocking) on the slot's
//                        named barrier -- scanners only ever wait for data.  (Eight scanners, one
//                        whole group each, needed ~1100 cycles per tile -- a dependent-issue-bound
//                        instruction stream -- against a ~1300-cycle HBM budget: too close.)
//   warps 16-20          finishers, one per ring slot: bar.sync on the slot's barrier, merge the two
//                        halves of every group, release the slot, one sigmoid of the maximum per
//                        group (fused.cu explains why that is exact) plus one for the tie test, then
//                        finish_tile (threshold, slot claim, key, the survivors' finished rows): its
//                        global round trips (~2000 cycles) are off everybody else's path.
// Ownership is by SLOT everywhere (producer q: slots with s % 2 == q; finisher f: slot f), so every
// waiter on a slot's mbarrier phase has itself seen the previous phase.  Two earlier assignments by
// TILE were wrong in ways only a perturbed schedule shows (a cold workspace; foreign kernels sharing
// the SMs, tools/pipeline_stress.py): a free finisher rotation let a fast finisher's bar.sync complete on
// the scanner arrivals of the slot's previous tile, and with phase waits added, a waiter two phases
// ahead passed mbarrier.try_wait.parity (which only distinguishes odd from even phases).
// The kernel is instantiated for fp32 and for fp16 level tensors (kHalf: FLOAT16 tensor maps, 64-byte
// stage rows with every group starting on an even row, exact upcast in the scanners and finishers).
// Tiles are assigned statically (tile = blockIdx.x + it * gridDim.x): every tile costs the same here.
// Round 2 tried to lift this kernel from 0.80 to 0.90 of the HBM rate and measured four re-designs, all
// bit-identical and all SLOWER on the cfg2 shape (56 us here): K1's structure -- six warps, each the only
// producer and consumer of its own stage, maxima-only reduction, argmax recovered for survivors only --
// 64 us, and 246 us on dense eval thresholds (six warps cannot hide the survivors' work and the ~1000
// cycles a warp needs to issue eight UTMALDGs); the same with dedicated producer warps 70 us; this
// kernel with maxima-only scanners and finisher-side argmax recovery 75 us (the finisher then holds the
// slot ~2000 cycles longer, and five slots is all that fits); this kernel with four producer warps 68 us.
// A per-role clock profile of THIS kernel then showed that its steady state is already at K1's HBM pace
// (a tile per 1256 cycles per CTA on 116 SMs, K1: 1220): the 0.81 of a launch timed alone is its ramp
// (two producer lanes need ~2500 cycles to request the first five tiles) and its tail (a finisher's tile
// takes ~6000 cycles from slot barrier to last row store, during which HBM idles at the end of the
// launch) -- exactly what back-to-back launches on alternating streams overlap.  What did help, mostly
// on dense tiles (150 -> 127 us), is mbar_wait_warp below.
// Positions past the end of a level are zero-filled by the TMA unit and masked by `valid`.
#include <type_traits>

#include "fused_tile.cuh"

namespace lp {

constexpr int KT_RING = 5;
constexpr int KT_SCANNERS = 2 * NGROUP;              // warps 0..15: warp w scans half (w & 1) of group w >> 1
constexpr int KT_FINISHERS = KT_RING;                // warps 16..20: finisher f owns ring slot f
constexpr int KT_PRODUCERS = 2;                      // warps 21..22: producer q owns the slots with s % 2 == q
constexpr int KT_THREADS = (KT_SCANNERS + KT_FINISHERS + KT_PRODUCERS) * 32;
constexpr int KT_STAGE_ELEMS = (ROW - 13) * DEC_TILE;    // class planes only; 4 (2) bytes each
constexpr int KT_STAGE_FLOATS = KT_STAGE_ELEMS;
constexpr int KT_PART_WORDS = KT_SCANNERS * DEC_TILE;    // per ring slot: one word per scanner and position
constexpr int KT_SMEM = KT_RING * (KT_STAGE_FLOATS + 3 * KT_PART_WORDS) * 4 + 2 * KT_RING * 8;   // fp16 stages use half of theirs

// Named barrier of ring slot s (1..5): the sixteen scanner warps arrive, the tile's finisher warp waits.
// (best, arg, before) of a range, followed by the same of a LATER range of the same group
__device__ __forceinline__ void merge_later(float& best, int& arg, float& before, float b1, int a1, float p1) {
    const bool later = b1 > best;
    before = later ? fmaxf(best, p1) : before;
    arg = later ? a1 : arg;
    best = fmaxf(best, b1);
}

__device__ __forceinline__ void slot_arrive(int s) {
    asm volatile("bar.arrive %0, %1;" ::"r"(1 + s), "n"((KT_SCANNERS + 1) * 32) : "memory");
}
__device__ __forceinline__ void slot_wait(int s) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + s), "n"((KT_SCANNERS + 1) * 32) : "memory");
}

// Channels [C0, C1) of one group of one anchor from the stage: maximum logit, its FIRST index (within
// the group), and the largest logit of the range that precedes that index (see fused.cu).  NC
// independent chains over consecutive sub-ranges, interleaved step by step (a warp issues in order: a
// single chain leaves it stalled on every compare -> select dependency), then folded left to right: a
// later sub-range wins only with a strictly larger maximum, and then everything in the earlier ones
// precedes its index.  Within a chain the running maximum is an FMNMX; the compare that drives the
// selects hangs off it.
// First channel row of class group g inside a stage.  A TMA box must land on a 128-byte boundary:
// fp32 rows are 128 B, so the groups pack densely; fp16 rows are 64 B, so every group starts on an even
// row (odd-width groups are followed by one unused row: 284 rows instead of 277).
template <bool kHalf>
__host__ __device__ constexpr int stage_row(int g) {
    if (!kHalf) return group_begin(g) - 13;
    int r = 0;
    for (int i = 0; i < g; ++i) r += ((group_begin(i + 1) - group_begin(i)) + 1) & ~1;
    return r;
}
constexpr int KT_STAGE_ELEMS_H = stage_row<true>(NGROUP) * DEC_TILE;   // 284 rows of 32 halves
static_assert(KT_STAGE_ELEMS_H * 2 % 128 == 0 && KT_STAGE_ELEMS_H * 2 <= KT_STAGE_FLOATS * 4, "fp16 stage layout");

// T = float or __half (stage element); a half is upcast exactly
__device__ __forceinline__ float stage_val(const float* p) { return *p; }
__device__ __forceinline__ float stage_val(const __half* p) { return __half2float(*p); }

template <int C0, int C1, class T>
__device__ __forceinline__ void range_scan_smem(const T* col0, float& best, int& arg, float& before) {
    constexpr int NC = 2, WIDTH = C1 - C0;
    constexpr int Q = (WIDTH + NC - 1) / NC;   // chain k covers [C0 + k*Q, min(C0 + (k+1)*Q, C1))
    float bm[NC], pm[NC];
    int am[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        bm[k] = stage_val(col0 + (C0 + k * Q) * DEC_TILE);
        pm[k] = -INFINITY;
        am[k] = C0 + k * Q;
    }
#pragma unroll
    for (int c = 1; c < Q; ++c) {
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            if (k * Q + c < WIDTH) {
                const float v = stage_val(col0 + (C0 + k * Q + c) * DEC_TILE);
                const bool up = v > bm[k];   // strict: the first occurrence of the maximum wins (torch.max)
                pm[k] = up ? bm[k] : pm[k];
                am[k] = up ? C0 + k * Q + c : am[k];
                bm[k] = fmaxf(bm[k], v);
            }
        }
    }
    best = bm[0];
    arg = am[0];
    before = pm[0];
#pragma unroll
    for (int k = 1; k < NC; ++k) merge_later(best, arg, before, bm[k], am[k], pm[k]);
}
template <int WIDTH, class T>
__device__ __forceinline__ void half_scan_smem(const T* col0, int half, float& best, int& arg, float& before) {
    constexpr int H = WIDTH / 2;
    if (half == 0) range_scan_smem<0, H>(col0, best, arg, before);
    else range_scan_smem<H, WIDTH>(col0, best, arg, before);
}

__device__ __forceinline__ void locate(const LevelsFilterParams& p, int tile, int& b, int& l, int& p0) {
    b = tile / p.tiles_per_image;
    const int r = tile - b * p.tiles_per_image;
    l = 0;
#pragma unroll
    for (int i = 1; i < LP_MAX_LEVELS; ++i)
        if (i < p.n_levels && r >= p.lv[i].tile_off) l = i;
    p0 = (r - p.lv[l].tile_off) * DEC_TILE;
}

#ifdef LP_KF_PROFILE
#define LP_PF_DECL long long pf[4] = {0, 0, 0, 0}, pf_t = clock64()
#define LP_PF(k) do { const long long t1_ = clock64(); pf[k] += t1_ - pf_t; pf_t = t1_; } while (0)
#define LP_PF_OUT() do { if (p.timing != nullptr && lane == 0) for (int k = 0; k < 4; ++k) p.timing[(blockIdx.x * 23 + warp) * 4 + k] = pf[k]; } while (0)
#else
#define LP_PF_DECL do { } while (0)
#define LP_PF(k) do { } while (0)
#define LP_PF_OUT() do { } while (0)
#endif

// -DLP_KF_ASSERT builds (liblpnms_kfassert.so, tests/test_gpu_kf_assert.py) tag every hand-off with the
// tile sequence number it belongs to and trap on a mismatch -- the three schedule-dependent bugs listed
// above were each a warp acting on a slot in the wrong PHASE, which only showed as rare wrong keys:
//   tag_issued[s]    producer, before the loads of tile `it` are issued into slot s
//   tag_scanned[s][w] scanner w, with its results of tile `it`
//   tag_released[s]  finisher, before it hands slot s back
// checks: scanner after `full` (the data in the slot is tile it's), finisher after the slot barrier (all
// sixteen results are tile it's), producer after `empty` (the tile it overwrites, it - RING, was released).
#ifdef LP_KF_ASSERT
#define LP_KA(cond) do { if (!(cond)) __trap(); } while (0)
#else
#define LP_KA(cond) do { } while (0)
#endif

template <bool kHalf>
__global__ void __launch_bounds__(KT_THREADS, 1) levels_filter_tma_kernel(const LevelsFilterParams p,
                                                                            const __grid_constant__ DecodeMaps maps) {
    using T = typename std::conditional<kHalf, __half, float>::type;
    constexpr int STAGE_ELEMS = kHalf ? KT_STAGE_ELEMS_H : KT_STAGE_ELEMS;
    constexpr int stage_rows[NGROUP] = {stage_row<kHalf>(0), stage_row<kHalf>(1), stage_row<kHalf>(2), stage_row<kHalf>(3),
                                        stage_row<kHalf>(4), stage_row<kHalf>(5), stage_row<kHalf>(6), stage_row<kHalf>(7)};
    extern __shared__ __align__(128) unsigned char smem[];
    T* stage0 = reinterpret_cast<T*>(smem);
    float* part_b = reinterpret_cast<float*>(smem) + KT_RING * KT_STAGE_FLOATS;   // same offsets for both element types                          // [ring][scanner][position] maximum logit
    float* part_p = part_b + KT_RING * KT_PART_WORDS;                            // ... largest logit before its index
    int* part_a = reinterpret_cast<int*>(part_p + KT_RING * KT_PART_WORDS);      // ... its first index in the group
    uint64_t* full = reinterpret_cast<uint64_t*>(part_a + KT_RING * KT_PART_WORDS);
    uint64_t* empty = full + KT_RING;
#ifdef LP_KF_ASSERT
    __shared__ volatile int tag_issued[KT_RING], tag_released[KT_RING], tag_scanned[KT_RING][KT_SCANNERS];
    if (threadIdx.x < KT_RING) { tag_issued[threadIdx.x] = -1; tag_released[threadIdx.x] = -1; }
    if (threadIdx.x < KT_RING * KT_SCANNERS) tag_scanned[threadIdx.x / KT_SCANNERS][threadIdx.x % KT_SCANNERS] = -1;
#endif

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < KT_RING; ++k) { mbar_init(&full[k], 1); mbar_init(&empty[k], 1); }
        mbar_fence_init();
    }
    __syncthreads();   // the only CTA-wide barrier: the roles never meet again

    const int first = blockIdx.x, step = gridDim.x;
    const int n_my = first < p.n_tiles ? (p.n_tiles - first + step - 1) / step : 0;

    if (warp >= KT_SCANNERS + KT_FINISHERS) {            // ---- producers
        if (lane != 0) return;
        LP_PF_DECL;
        // A producer owns whole ring slots, not alternate tiles: whoever waits on a slot's `empty` phase u
        // must have seen phase u-1 itself, or try_wait's parity test mistakes "two phases behind" for
        // "done" (the slot's previous tile not even taken over yet) and the stage is overwritten early.
        const int q = warp - (KT_SCANNERS + KT_FINISHERS);
        for (int it = 0; it < n_my; ++it) {
            const int s = it % KT_RING, use = it / KT_RING;
            if (s % KT_PRODUCERS != q) continue;
            int b, l, p0;
            locate(p, first + it * step, b, l, p0);
            LP_PF(0);
            if (use > 0) mbar_wait_relaxed(&empty[s], (use - 1) & 1);
#ifdef LP_KF_ASSERT
            LP_KA(tag_released[s] == (use > 0 ? it - KT_RING : -1));   // the slot's previous tile was handed back, no other
            LP_KA(tag_issued[s] == (use > 0 ? it - KT_RING : -1));
            tag_issued[s] = it;
            __threadfence_block();
#endif
            LP_PF(1);
            T* stage = stage0 + s * STAGE_ELEMS;
            fence_proxy_async_smem();   // the stage was last read through the generic proxy
            mbar_expect_tx(&full[s], KT_STAGE_ELEMS * (int)sizeof(T));   // the boxes' bytes (padding rows are not written)
#pragma unroll
            for (int g = 0; g < NGROUP; ++g)
                tma_load_3d(stage + stage_row<kHalf>(g) * DEC_TILE, &maps.m[l][g], p0, 0, b, &full[s]);
            LP_PF(2);
        }
        LP_PF_OUT();
    } else if (warp >= KT_SCANNERS) {                    // ---- finishers
        LP_PF_DECL;
        for (int it = warp - KT_SCANNERS; it < n_my; it += KT_FINISHERS) {   // it % KT_RING == this finisher's slot
            const int s = it % KT_RING;
            int b, l, p0;
            locate(p, first + it * step, b, l, p0);
            const DecodeLevel& lv = p.lv[l];
            LP_PF(0);
            // (One finisher per SLOT: its tiles reach it in order, so its bar.sync can never complete on
            // the scanner arrivals of the slot's previous tile -- which a free rotation allowed, first
            // through plain overtaking, then, with an `empty`-phase wait in front, through the parity
            // aliasing described at the producers.)
            slot_wait(s);
#ifdef LP_KF_ASSERT
            if (lane < KT_SCANNERS) LP_KA(tag_scanned[s][lane] == it);   // every scanner's result is THIS tile's
            LP_KA(tag_issued[s] == it);
#endif
            LP_PF(1);               // all sixteen half-group results of tile `it` are in the exchange buffer
            float c[NGROUP];
            unsigned long long args = 0;
            unsigned ties = 0;
            float best[NGROUP], before[NGROUP];
#pragma unroll
            for (int g = 0; g < NGROUP; ++g) {   // the two halves of a group, merged with first-index semantics
                const int q = (s * KT_SCANNERS + 2 * g) * DEC_TILE + lane;
                best[g] = part_b[q];
                before[g] = part_p[q];
                int arg = part_a[q];
                merge_later(best[g], arg, before[g], part_b[q + DEC_TILE], part_a[q + DEC_TILE], part_p[q + DEC_TILE]);
                args |= (unsigned long long)arg << (6 * g);
            }
            __syncwarp();               // every lane has its copy: the slot (stage + exchange) may be refilled
#ifdef LP_KF_ASSERT
            if (lane == 0) { tag_released[s] = it; __threadfence_block(); }
#endif
            if (lane == 0) mbar_arrive(&empty[s]);
            //   score: sigmoid of the maximum logit == maximum of the sigmoids (monotone device sigmoid);
            //   tie:   arg is the first index of the maximum LOGIT; the reference takes the first index of
            //          the maximum SIGMOID, which is earlier iff a smaller logit before it rounds to the
            //          same value -- checked exactly with one more sigmoid.
#pragma unroll
            for (int g = 0; g < NGROUP; ++g) {
                c[g] = __fmul_rn(sigmoid_f32(best[g]), 1.0f);   // cls * obj, obj == 1 (nms.py:76)
                if (((args >> (6 * g)) & 63u) != 0 && sigmoid_f32(before[g]) == c[g]) ties |= 1u << g;
            }
            const int pos = p0 + lane;
            finish_tile<kHalf>(p, lv, b, pos, pos < lv.hw, c, args, ties, lane);
            LP_PF(2);
        }
        LP_PF_OUT();
    } else {                                             // ---- scanners: warp w owns half (w & 1) of group w >> 1
        LP_PF_DECL;
        for (int it = 0; it < n_my; ++it) {
            const int s = it % KT_RING;
            LP_PF(0);
            mbar_wait_warp(&full[s], (it / KT_RING) & 1, lane);   // one poller, and the warp leaves CONVERGED (common.cuh)
#ifdef LP_KF_ASSERT
            LP_KA(tag_issued[s] == it);                                  // the slot holds THIS tile's planes
            LP_KA(tag_scanned[s][warp] == (it >= KT_RING ? it - KT_RING : -1));
#endif
            LP_PF(1);               // a real barrier for the compiler too: no stage read may move above it
            const int g = warp >> 1, half = warp & 1;
            const T* col0 = stage0 + s * STAGE_ELEMS + stage_rows[g] * DEC_TILE + lane;
            float best, before;
            int arg;
            if (g == 0) half_scan_smem<31>(col0, half, best, arg, before);
            else if (g == 1) half_scan_smem<24>(col0, half, best, arg, before);
            else half_scan_smem<37>(col0, half, best, arg, before);
            const int q = (s * KT_SCANNERS + warp) * DEC_TILE + lane;
            part_b[q] = best;
            part_p[q] = before;
            part_a[q] = arg;
#ifdef LP_KF_ASSERT
            LP_KA(tag_issued[s] == it);                                  // ... and still does after the scan
            __syncwarp();
            if (lane == 0) { tag_scanned[s][warp] = it; __threadfence_block(); }
#endif
            slot_arrive(s);             // non-blocking
            LP_PF(2);
        }
        LP_PF_OUT();
    }
}

cudaError_t launch_levels_filter_tma(const LevelsFilterParams& p, const DecodeMaps& maps, int num_ctas, cudaStream_t stream) {
    static_assert(KT_SMEM <= 227 * 1024, "KF stages exceed shared memory");
    static_assert(KT_RING + 1 <= 16, "one named barrier per ring slot");
    static_assert(KT_FINISHERS == KT_RING, "a slot's consecutive tiles must share their finisher");
    static bool configured[64] = {false}, configured_h[64] = {false};
    cudaError_t e = p.half_levels ? configure_smem_once(levels_filter_tma_kernel<true>, KT_SMEM, configured_h)
                                  : configure_smem_once(levels_filter_tma_kernel<false>, KT_SMEM, configured);
    if (e != cudaSuccess) return e;
    const int grid = p.n_tiles < num_ctas ? p.n_tiles : num_ctas;
    if (p.half_levels) levels_filter_tma_kernel<true><<<grid, KT_THREADS, KT_SMEM, stream>>>(p, maps);
    else levels_filter_tma_kernel<false><<<grid, KT_THREADS, KT_SMEM, stream>>>(p, maps);
    return cudaGetLastError();
}

}  // namespace lp
