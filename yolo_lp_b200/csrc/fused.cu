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
// read here (1108 B per anchor); box and corner planes are touched by K2 for candidates only.
// Traffic per anchor: 1108 B instead of 1156 + 1160 (decode) + 1160 (K1) = 3476 B.
#include "kernels.cuh"

namespace lp {

static_assert(DEC_TILE == 32, "a KF tile is one warp wide");
constexpr int KF_THREADS = 256;
constexpr int KF_WARPS = KF_THREADS / 32;

template <int WIDTH>
__device__ __forceinline__ float group_max_logit(const float* __restrict__ base, size_t hw, bool valid) {
    float v[WIDTH];
#pragma unroll
    for (int c = 0; c < WIDTH; ++c) v[c] = valid ? __ldg(base + c * hw) : 0.0f;
    float m = v[0];
#pragma unroll
    for (int c = 1; c < WIDTH; ++c) m = fmaxf(m, v[c]);
    return m;
}

__global__ void __launch_bounds__(KF_THREADS, 3) levels_filter_kernel(const LevelsFilterParams p) {
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * KF_WARPS + (threadIdx.x >> 5);
    const int n_warps = gridDim.x * KF_WARPS;
    // (image, tile-in-image) walked without a division per tile
    int b = gw / p.tiles_per_image, r = gw - b * p.tiles_per_image;
    for (int tile = gw; tile < p.n_tiles; tile += n_warps) {
        int l = 0;
#pragma unroll
        for (int i = 1; i < LP_MAX_LEVELS; ++i)
            if (i < p.n_levels && r >= p.lv[i].tile_off) l = i;
        const DecodeLevel& lv = p.lv[l];
        const int pos = (r - lv.tile_off) * DEC_TILE + lane;
        const bool valid = pos < lv.hw;
        const size_t hw = (size_t)lv.hw;
        const size_t off = (size_t)b * hw;  // image offset in units of one channel plane... times width below

        float c[NGROUP];
        c[0] = group_max_logit<31>(lv.cls[0] + off * 31 + pos, hw, valid);
        c[1] = group_max_logit<24>(lv.cls[1] + off * 24 + pos, hw, valid);
        c[2] = group_max_logit<37>(lv.cls[2] + off * 37 + pos, hw, valid);
        c[3] = group_max_logit<37>(lv.cls[3] + off * 37 + pos, hw, valid);
        c[4] = group_max_logit<37>(lv.cls[4] + off * 37 + pos, hw, valid);
        c[5] = group_max_logit<37>(lv.cls[5] + off * 37 + pos, hw, valid);
        c[6] = group_max_logit<37>(lv.cls[6] + off * 37 + pos, hw, valid);
        c[7] = group_max_logit<37>(lv.cls[7] + off * 37 + pos, hw, valid);
#pragma unroll
        for (int g = 0; g < NGROUP; ++g) c[g] = __fmul_rn(sigmoid_f32(c[g]), 1.0f);  // cls * obj, obj == 1 (nms.py:76)
        float filt, score;
        lp_means(c, filt, score);

        const bool pass = valid && (filt >= p.conf);
        const unsigned m = __ballot_sync(0xffffffffu, pass);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(p.counts + b, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (pass)
                p.keys[(size_t)b * p.key_stride + base + __popc(m & ((1u << lane) - 1u))] =
                    make_key(score, (unsigned)(lv.anchor_off + pos));
        }
        r += n_warps;
        while (r >= p.tiles_per_image) {
            r -= p.tiles_per_image;
            ++b;
        }
    }
}

cudaError_t launch_levels_filter(const LevelsFilterParams& p, int num_sms, cudaStream_t stream) {
    if (p.n_tiles <= 0) return cudaSuccess;
    int grid = (p.n_tiles + KF_WARPS - 1) / KF_WARPS;
    const int cap = num_sms * 3;
    if (grid > cap) grid = cap;
    levels_filter_kernel<<<grid, KF_THREADS, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace lp
