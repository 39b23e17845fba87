// KF, TMA + warp-specialised variant of fused.cu for level planes whose rows are 16-byte aligned
// (h*w % 4 == 0): same arithmetic and outputs (shared finish_tile), different data movement.
//
// The register-resident kernel of fused.cu is latency-bound (0.74 of the HBM copy rate): a warp's
// loads, its compare/select chains and the epilogue's global round trips (slot atomics, box / corner
// reads of the survivors) all sit in one instruction stream.  Here they are three concurrent roles of
// one persistent CTA per SM, over tiles of 32 positions ([277 class channels][32] = 35,456 B) in a
// six-deep shared-memory ring:
//   warps 14-15, lane 0  producers (tiles alternate between them): eight 3-D TMA box loads per tile
//                        (UTMALDG.3D, one per class tensor: all its channels x 32 positions of image
//                        b), completion on the slot's `full` mbarrier, reuse gated by `empty`.  A
//                        thread needs ~150 cycles per UTMALDG, hence two issuers;
//   warps 0-7            scanners, warp g owns class group g: lanes along positions (conflict-free
//                        LDS), running maximum as an FMNMX chain with the first-index / runner-up
//                        selects hanging off it, two chains per group; one sigmoid of the maximum
//                        (fused.cu explains why that is exact); leaves score, argmax and tie flag per
//                        position in a small exchange buffer and bar.arrive's (non-blocking) on the
//                        tile's named barrier -- scanners only ever wait for data;
//   warps 8-13           finishers, one per ring slot: bar.sync on the slot's barrier, pull the eight
//                        groups' results into registers, release the slot, then finish_tile
//                        (threshold, slot claim, key, the survivors' finished rows): its ~3000 cycles
//                        of global latency are off everybody else's critical path.  (One finisher
//                        per SLOT, not a free rotation: with four finishers over six slots a fast
//                        finisher's bar.sync completed a barrier whose 256 scanner arrivals belonged
//                        to the slot's previous tile, still waiting for its own slow finisher.)
// Tiles are assigned statically (tile = blockIdx.x + it * gridDim.x): every tile costs the same here.
// Positions past the end of a level are zero-filled by the TMA unit and masked by `valid`.
#include "fused_tile.cuh"

namespace lp {

constexpr int KT_RING = 6;
constexpr int KT_SCANNERS = NGROUP;                  // warps 0..7
constexpr int KT_FINISHERS = 12;                     // warps 8..19: tile t belongs to finisher t % 12
constexpr int KT_PRODUCERS = 2;                      // warps 20..21
constexpr int KT_THREADS = (KT_SCANNERS + KT_FINISHERS + KT_PRODUCERS) * 32;
constexpr int KT_STAGE_FLOATS = (ROW - 13) * DEC_TILE;   // class planes only
constexpr int KT_PART_WORDS = NGROUP * DEC_TILE;         // per ring slot: one word per group and position
constexpr int KT_SMEM = KT_RING * (KT_STAGE_FLOATS + 2 * KT_PART_WORDS) * 4 + 2 * KT_RING * 8;

// Named barrier of ring slot s (1..6): the eight scanner warps arrive, the tile's finisher warp waits.
__device__ __forceinline__ void slot_arrive(int s) {
    asm volatile("bar.arrive %0, %1;" ::"r"(1 + s), "n"((KT_SCANNERS + 1) * 32) : "memory");
}
__device__ __forceinline__ void slot_wait(int s) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + s), "n"((KT_SCANNERS + 1) * 32) : "memory");
}

// One group of one anchor from the stage: maximum logit, its FIRST index, and the largest logit that
// precedes that index (see fused.cu).  Four independent chains over consecutive quarters of the
// channels, interleaved step by step (a warp issues in order: a single chain leaves it stalled on
// every compare -> select dependency), then folded left to right: a later quarter wins only with a
// strictly larger maximum, and then everything in the earlier quarters precedes its index.
// Within a chain the running maximum is an FMNMX; the compare that drives the selects hangs off it.
template <int WIDTH>
__device__ __forceinline__ void group_scan_smem(const float* col0, float& best, int& arg, float& before) {
    constexpr int NC = 4;
    constexpr int Q = (WIDTH + NC - 1) / NC;   // chain k covers [k*Q, min((k+1)*Q, WIDTH))
    float bm[NC], pm[NC];
    int am[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        bm[k] = col0[k * Q * DEC_TILE];
        pm[k] = -INFINITY;
        am[k] = k * Q;
    }
#pragma unroll
    for (int c = 1; c < Q; ++c) {
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            if (k * Q + c < WIDTH) {
                const float v = col0[(k * Q + c) * DEC_TILE];
                const bool up = v > bm[k];   // strict: the first occurrence of the maximum wins (torch.max)
                pm[k] = up ? bm[k] : pm[k];
                am[k] = up ? k * Q + c : am[k];
                bm[k] = fmaxf(bm[k], v);
            }
        }
    }
    best = bm[0];
    arg = am[0];
    before = pm[0];
#pragma unroll
    for (int k = 1; k < NC; ++k) {
        const bool later = bm[k] > best;
        before = later ? fmaxf(best, pm[k]) : before;
        arg = later ? am[k] : arg;
        best = fmaxf(best, bm[k]);
    }
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
#define LP_PF_OUT() do { if (p.timing != nullptr && lane == 0) for (int k = 0; k < 4; ++k) p.timing[(blockIdx.x * 22 + warp) * 4 + k] = pf[k]; } while (0)
#else
#define LP_PF_DECL do { } while (0)
#define LP_PF(k) do { } while (0)
#define LP_PF_OUT() do { } while (0)
#endif

__global__ void __launch_bounds__(KT_THREADS, 1) levels_filter_tma_kernel(const LevelsFilterParams p,
                                                                            const __grid_constant__ DecodeMaps maps) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* stage0 = reinterpret_cast<float*>(smem);
    float* part_c = stage0 + KT_RING * KT_STAGE_FLOATS;                              // [ring][group][position] score
    unsigned* part_a = reinterpret_cast<unsigned*>(part_c + KT_RING * KT_PART_WORDS); // argmax | tie << 8
    uint64_t* full = reinterpret_cast<uint64_t*>(part_a + KT_RING * KT_PART_WORDS);
    uint64_t* empty = full + KT_RING;

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
        for (int it = warp - (KT_SCANNERS + KT_FINISHERS); it < n_my; it += KT_PRODUCERS) {
            const int s = it % KT_RING, use = it / KT_RING;
            int b, l, p0;
            locate(p, first + it * step, b, l, p0);
            LP_PF(0);
            if (use > 0) mbar_wait_relaxed(&empty[s], (use - 1) & 1);
            LP_PF(1);
            float* stage = stage0 + s * KT_STAGE_FLOATS;
            fence_proxy_async_smem();   // the stage was last read through the generic proxy
            mbar_expect_tx(&full[s], KT_STAGE_FLOATS * 4);
#pragma unroll
            for (int g = 0; g < NGROUP; ++g)
                tma_load_3d(stage + (group_begin(g) - 13) * DEC_TILE, &maps.m[l][g], p0, 0, b, &full[s]);
            LP_PF(2);
        }
        LP_PF_OUT();
    } else if (warp >= KT_SCANNERS) {                    // ---- finishers
        LP_PF_DECL;
        for (int it = warp - KT_SCANNERS; it < n_my; it += KT_FINISHERS) {
            const int s = it % KT_RING;
            int b, l, p0;
            locate(p, first + it * step, b, l, p0);
            const DecodeLevel& lv = p.lv[l];
            LP_PF(0);
            // Not before the slot's previous tile has been taken over by ITS finisher: a bar.sync issued
            // earlier would complete on the 256 scanner arrivals that belong to that tile.
            if (it >= KT_RING) {
                if (lane == 0) mbar_wait_relaxed(&empty[s], (it / KT_RING - 1) & 1);
                __syncwarp();
            }
            slot_wait(s);
            LP_PF(1);               // all eight groups of tile `it` are in the exchange buffer
            float c[NGROUP];
            unsigned long long args = 0;
            unsigned ties = 0;
#pragma unroll
            for (int g = 0; g < NGROUP; ++g) {
                c[g] = part_c[(s * NGROUP + g) * DEC_TILE + lane];
                const unsigned a = part_a[(s * NGROUP + g) * DEC_TILE + lane];
                args |= (unsigned long long)(a & 63u) << (6 * g);
                ties |= (a >> 8) << g;
            }
            __syncwarp();               // every lane has its copy: the slot (stage + exchange) may be refilled
            if (lane == 0) mbar_arrive(&empty[s]);
            const int pos = p0 + lane;
            finish_tile(p, lv, b, pos, pos < lv.hw, c, args, ties, lane);
            LP_PF(2);
        }
        LP_PF_OUT();
    } else {                                             // ---- scanners: warp g owns class group g
        LP_PF_DECL;
        for (int it = 0; it < n_my; ++it) {
            const int s = it % KT_RING;
            LP_PF(0);
            if (lane == 0) mbar_wait(&full[s], (it / KT_RING) & 1);
            __syncwarp();
            LP_PF(1);               // a real barrier for the compiler too: no stage read may move above it
            const float* col0 = stage0 + s * KT_STAGE_FLOATS + (group_begin(warp) - 13) * DEC_TILE + lane;
            float best, before;
            int arg;
            if (warp == 0) group_scan_smem<31>(col0, best, arg, before);
            else if (warp == 1) group_scan_smem<24>(col0, best, arg, before);
            else group_scan_smem<37>(col0, best, arg, before);
            //   score: sigmoid of the maximum logit == maximum of the sigmoids (monotone device sigmoid);
            //   tie:   arg is the first index of the maximum LOGIT; the reference takes the first index of
            //          the maximum SIGMOID, which is earlier iff a smaller logit before it rounds to the
            //          same value -- checked exactly with one more sigmoid.
            const float cg = __fmul_rn(sigmoid_f32(best), 1.0f);   // cls * obj, obj == 1 (nms.py:76)
            const unsigned tie = (arg > 0 && sigmoid_f32(before) == cg) ? 1u : 0u;
            part_c[(s * NGROUP + warp) * DEC_TILE + lane] = cg;
            part_a[(s * NGROUP + warp) * DEC_TILE + lane] = (unsigned)arg | (tie << 8);
            slot_arrive(s);             // non-blocking
            LP_PF(2);
        }
        LP_PF_OUT();
    }
}

cudaError_t launch_levels_filter_tma(const LevelsFilterParams& p, const DecodeMaps& maps, int num_ctas, cudaStream_t stream) {
    static_assert(KT_SMEM <= 227 * 1024, "KF stages exceed shared memory");
    static_assert(KT_RING + 1 <= 16, "one named barrier per ring slot");
    cudaError_t e = cudaFuncSetAttribute(levels_filter_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KT_SMEM);
    if (e != cudaSuccess) return e;
    const int grid = p.n_tiles < num_ctas ? p.n_tiles : num_ctas;
    levels_filter_tma_kernel<<<grid, KT_THREADS, KT_SMEM, stream>>>(p, maps);
    return cudaGetLastError();
}

}  // namespace lp
