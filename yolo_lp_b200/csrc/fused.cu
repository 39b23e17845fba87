// KF -- fused head decode + score + threshold + compact, straight from the raw per-level
// prediction-conv outputs (NCHW), without materialising the [B, A, 290] head tensor.
//
// Restates effidehead.py:251-258 (sigmoid), nms.py:76 (cls * obj with obj == 1.0, effidehead.py:290),
// :81-88 (eight group maxima), :90-91 (the buggy filter mean) and :120 (NMS score) per anchor and
// emits the same 64-bit sort key as K1.  Because the device sigmoid is monotone non-decreasing
// (checked exhaustively, tools/sigmoid_monotone.py), max_j sigmoid(x_j) == sigmoid(max_j x_j)
// bit for bit, so a group maximum costs one sigmoid instead of 31/24/37 and the result is
// identical to decode kernel -> K1 on the same level tensors.
//
// HBM-bound, no shared memory: a warp owns a tile of 32 consecutive positions of one level of one
// image; lanes run along positions, so every channel read is one fully coalesced 128-byte line and
// a lane keeps a whole group (<= 37 independent loads) in flight.  Only the 277 class planes are
// streamed (1108 B per anchor); box and corner planes are touched for candidates only.
// Traffic per anchor: 1108 B instead of 1156 + 1160 (decode) + 1160 (K1) = 3476 B.
// (A variant on the decode kernel's cp.async shared-memory ring was measured too: 118 us against
// 62 us for this register-resident form on the cfg2 shape.  The TMA-fed, warp-specialised
// fused_tma.cu is the one that beats it -- 57 us -- and is the default for 16-byte aligned planes;
// this kernel serves every other shape.)
//
// The same pass tracks, per group, the first index of the maximum (torch.max semantics,
// nms.py:81-88); a surviving lane then decodes its own box and corners (effidehead.py:283-286,
// nms.py:79) and stores the finished 28-float detection row by slot, so that K2 only has to order,
// suppress and copy 112-byte records.  The one subtlety -- two different logits rounding to the
// same sigmoid, where the reference's argmax is the earlier index -- is detected exactly (one more
// sigmoid per group) and resolved by a rare warp-cooperative pass.
#include "fused_tile.cuh"

namespace lp {

static_assert(DEC_TILE == 32, "a KF tile is one warp wide");
constexpr int KF_THREADS = 768;   // one CTA of 24 warps per SM: whole SMs can then be left to K2 (see launch)
constexpr int KF_WARPS = KF_THREADS / 32;

// One class group of one anchor (lane).  load: all of the group's logits in flight at once.
// scan: maximum logit, its FIRST index, and the largest logit that precedes that index (needed to
// detect ties in sigmoid space, see below).
template <int WIDTH>
__device__ __forceinline__ void group_load(float (&v)[37], const float* __restrict__ base, size_t hw, bool valid) {
#pragma unroll
    for (int c = 0; c < WIDTH; ++c) v[c] = valid ? __ldg(base + c * hw) : 0.0f;
}
template <int WIDTH>
__device__ __forceinline__ void group_scan(const float (&v)[37], float& best, int& arg, float& before) {
    best = v[0];
    arg = 0;
    before = -INFINITY;
#pragma unroll
    for (int c = 1; c < WIDTH; ++c) {
        const bool up = v[c] > best;   // strict: the first occurrence of the maximum wins (torch.max)
        before = up ? best : before;   // best so far == max of everything before index c
        arg = up ? c : arg;
        best = up ? v[c] : best;
    }
}

__global__ void __launch_bounds__(KF_THREADS, 1) levels_filter_kernel(const LevelsFilterParams p) {
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * KF_WARPS + (threadIdx.x >> 5);
    const int n_warps = gridDim.x * KF_WARPS;
    // tiles: the first one is static (the global warp id), later ones are claimed from a global
    // counter one claim ahead of use -- keeps every warp busy until the last tile whatever the grid
    int next = 0x7fffffff;
    if (lane == 0 && gw < p.n_tiles) next = n_warps + (int)atomicAdd(p.tile_counter, 1u);
    for (int tile = gw; tile < p.n_tiles;) {
        const int b = tile / p.tiles_per_image, r = tile - b * p.tiles_per_image;
        int l = 0;
#pragma unroll
        for (int i = 1; i < LP_MAX_LEVELS; ++i)
            if (i < p.n_levels && r >= p.lv[i].tile_off) l = i;
        const DecodeLevel& lv = p.lv[l];
        const int pos = (r - lv.tile_off) * DEC_TILE + lane;
        const bool valid = pos < lv.hw;
        const size_t hw = (size_t)lv.hw;
        const size_t off = (size_t)b * hw;  // image offset of a 1-channel plane; times the group width below

        // Per-group state is folded as soon as the group is scanned (score, 6-bit argmax packed into
        // `args`, tie bit) so that only one group of loads plus ~12 registers of state stay live:
        // that keeps three CTAs (24 warps) resident per SM, which this latency-bound kernel needs.
        //   score: sigmoid of the maximum logit == maximum of the sigmoids (monotone device sigmoid);
        //   tie:   arg is the first index of the maximum LOGIT; the reference takes the first index of
        //          the maximum SIGMOID, which is earlier iff a smaller logit before it rounds to the
        //          same value -- checked exactly with one more sigmoid.
        float c[NGROUP];
        unsigned long long args = 0;
        unsigned ties = 0;
        float v[37];
#define LP_GROUP(G, W)                                                                  \
        {                                                                               \
            group_load<W>(v, lv.cls[G] + off * W + pos, hw, valid);                     \
            float best, before;                                                         \
            int arg;                                                                    \
            group_scan<W>(v, best, arg, before);                                        \
            c[G] = __fmul_rn(sigmoid_f32(best), 1.0f); /* cls * obj, obj == 1 (nms.py:76) */ \
            if (arg > 0 && sigmoid_f32(before) == c[G]) ties |= 1u << G;                \
            args |= (unsigned long long)arg << (6 * G);                                 \
        }
        LP_GROUP(0, 31) LP_GROUP(1, 24) LP_GROUP(2, 37) LP_GROUP(3, 37)
        LP_GROUP(4, 37) LP_GROUP(5, 37) LP_GROUP(6, 37) LP_GROUP(7, 37)
#undef LP_GROUP
        finish_tile(p, lv, b, pos, valid, c, args, ties, lane);
        tile = __shfl_sync(0xffffffffu, next, 0);
        if (lane == 0 && tile < p.n_tiles) next = n_warps + (int)atomicAdd(p.tile_counter, 1u);
    }
}

cudaError_t launch_levels_filter(const LevelsFilterParams& p, const DecodeMaps* maps, int num_ctas, cudaStream_t stream) {
    if (p.n_tiles <= 0) return cudaSuccess;
    if (maps != nullptr) return launch_levels_filter_tma(p, *maps, num_ctas, stream);
    int grid = (p.n_tiles + KF_WARPS - 1) / KF_WARPS;
    if (grid > num_ctas) grid = num_ctas;
    levels_filter_kernel<<<grid, KF_THREADS, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace lp
