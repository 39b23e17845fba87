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
// (A variant on the decode kernel's 6-stage cp.async shared-memory ring was measured too: 118 us
// against 62 us for this register-resident form on the cfg2 shape.)
//
// The same pass tracks, per group, the first index of the maximum (torch.max semantics,
// nms.py:81-88); a surviving lane then decodes its own box and corners (effidehead.py:283-286,
// nms.py:79) and stores the finished 28-float detection row by slot, so that K2 only has to order,
// suppress and copy 112-byte records.  The one subtlety -- two different logits rounding to the
// same sigmoid, where the reference's argmax is the earlier index -- is detected exactly (one more
// sigmoid per group) and resolved by a rare warp-cooperative pass.
#include "kernels.cuh"

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

// Exact first argmax in sigmoid space for one group of one anchor, warp-cooperative (lanes along the
// group's columns).  Only reached when two different logits round to the same sigmoid.
__device__ __noinline__ int group_argmax_exact(const float* plane, size_t hw, int width, int lane) {
    constexpr int kInvalid = 1 << 20;
    float best = -INFINITY;
    int bi = kInvalid;
    if (lane < width) {
        best = sigmoid_f32(__ldg(plane + (size_t)lane * hw));
        bi = lane;
    }
    if (lane + 32 < width) {
        const float v = sigmoid_f32(__ldg(plane + (size_t)(lane + 32) * hw));
        if (v > best) { best = v; bi = lane + 32; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    return bi;
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
        float filt, score;
        lp_means(c, filt, score);

        const bool pass = valid && (filt >= p.conf);
        const unsigned m = __ballot_sync(0xffffffffu, pass);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(p.counts + b, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            const unsigned slot = base + __popc(m & ((1u << lane) - 1u));
            // rare: resolve sigmoid-space ties exactly, one survivor and group at a time
            for (unsigned todo = __ballot_sync(0xffffffffu, pass && ties != 0); todo; todo &= todo - 1) {
                const int src = __ffs(todo) - 1;
                const int cpos = __shfl_sync(0xffffffffu, pos, src);
                const unsigned tg = __shfl_sync(0xffffffffu, ties, src);
#pragma unroll
                for (int g = 0; g < NGROUP; ++g) {
                    if (!((tg >> g) & 1u)) continue;  // warp-uniform
                    const int width = group_begin(g + 1) - group_begin(g);
                    const int exact = group_argmax_exact(lv.cls[g] + off * width + cpos, hw, width, lane);
                    if (lane == src) args = (args & ~(63ull << (6 * g))) | ((unsigned long long)exact << (6 * g));
                }
            }
            if (pass) {
                // this lane finishes its own row: box (nms.py:79 on effidehead.py:283,285), corners (:284,286)
                const unsigned anchor = (unsigned)(lv.anchor_off + pos);
                p.keys[(size_t)b * p.key_stride + slot] = make_key(score, anchor);
                p.slot_of[(size_t)b * p.A + anchor] = slot;
                const float* reg = lv.reg + off * 4 + pos;
                const float* cor = lv.cor + off * 8 + pos;
                const int y = pos / lv.w, x = pos - y * lv.w;
                const float ax = anchor_coord(x), ay = anchor_coord(y);
                const float4 q = decode_box(ax, ay, __ldg(reg), __ldg(reg + hw), __ldg(reg + 2 * hw), __ldg(reg + 3 * hw), lv.stride);
                float4* row = reinterpret_cast<float4*>(p.rec + ((size_t)b * p.A + slot) * OUTW);
                float k[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) k[i] = decode_corner(i, ax, ay, __ldg(cor + i * hw), lv.stride);
                row[0] = xywh_to_xyxy(q.x, q.y, q.z, q.w);
                row[1] = make_float4(k[0], k[1], k[2], k[3]);
                row[2] = make_float4(k[4], k[5], k[6], k[7]);
                row[3] = make_float4(c[0], c[1], c[2], c[3]);
                row[4] = make_float4(c[4], c[5], c[6], c[7]);
                float a[NGROUP];
#pragma unroll
                for (int g = 0; g < NGROUP; ++g) a[g] = (float)(unsigned)((args >> (6 * g)) & 63u);
                row[5] = make_float4(a[0], a[1], a[2], a[3]);
                row[6] = make_float4(a[4], a[5], a[6], a[7]);
            }
        }
        tile = __shfl_sync(0xffffffffu, next, 0);
        if (lane == 0 && tile < p.n_tiles) next = n_warps + (int)atomicAdd(p.tile_counter, 1u);
    }
}

cudaError_t launch_levels_filter(const LevelsFilterParams& p, int num_ctas, cudaStream_t stream) {
    if (p.n_tiles <= 0) return cudaSuccess;
    int grid = (p.n_tiles + KF_WARPS - 1) / KF_WARPS;
    if (grid > num_ctas) grid = num_ctas;
    levels_filter_kernel<<<grid, KF_THREADS, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace lp
