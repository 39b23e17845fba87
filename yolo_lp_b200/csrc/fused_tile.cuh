// Shared by the two KF kernels (fused.cu: register-resident loads, fused_tma.cu: TMA-staged loads):
// what happens to a tile of 32 anchors once the eight group maxima / argmaxes are known.
#pragma once
#include <cuda_fp16.h>

#include "kernels.cuh"

namespace lp {

// Element idx of a level tensor.  kHalf: the DecodeLevel pointers address IEEE halves (fp16 level
// tensors, the reference's --half mode) and every value is upcast exactly on load.
template <bool kHalf>
__device__ __forceinline__ float ld_level(const float* base, size_t idx) {
    return kHalf ? __half2float(__ldg(reinterpret_cast<const __half*>(base) + idx)) : __ldg(base + idx);
}

// Exact first argmax in sigmoid space for one group of one anchor, warp-cooperative (lanes along the
// group's columns).  Only reached when two different logits round to the same sigmoid.
template <bool kHalf>
static __device__ __noinline__ int group_argmax_exact(const float* base, size_t idx0, size_t hw, int width, int lane) {
    constexpr int kInvalid = 1 << 20;
    float best = -INFINITY;
    int bi = kInvalid;
    if (lane < width) {
        best = sigmoid_f32(ld_level<kHalf>(base, idx0 + (size_t)lane * hw));
        bi = lane;
    }
    if (lane + 32 < width) {
        const float v = sigmoid_f32(ld_level<kHalf>(base, idx0 + (size_t)(lane + 32) * hw));
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

// c[g] = group scores, args = eight 6-bit first-argmax indices of the maximum LOGIT, ties = groups
// whose argmax must be re-derived in sigmoid space.  Filter (nms.py:90-91), slot claim, key, and
// for the survivors the finished 28-float row.  Warp-collective: all 32 lanes call it.
template <bool kHalf = false>
__device__ __forceinline__ void finish_tile(const LevelsFilterParams& p, const DecodeLevel& lv, int b, int pos, bool valid,
                                            const float (&c)[NGROUP], unsigned long long args, unsigned ties, int lane) {
    const size_t hw = (size_t)lv.hw;
    const size_t off = (size_t)b * hw;
    float filt, score;
    lp_means(c, filt, score);

    const bool pass = valid && (filt >= p.conf);
    const unsigned m = __ballot_sync(0xffffffffu, pass);
    if (m) {
        // Survivors read their box / corner distances first, so that this round trip to memory overlaps
        // the slot claim's (an L2 atomic) instead of following it.
        float d[12];
        if (pass) {
#pragma unroll
            for (int i = 0; i < 4; ++i) d[i] = ld_level<kHalf>(lv.reg, off * 4 + pos + i * hw);
#pragma unroll
            for (int i = 0; i < 8; ++i) d[4 + i] = ld_level<kHalf>(lv.cor, off * 8 + pos + i * hw);
        }
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
                const int exact = group_argmax_exact<kHalf>(lv.cls[g], off * width + cpos, hw, width, lane);
                if (lane == src) args = (args & ~(63ull << (6 * g))) | ((unsigned long long)exact << (6 * g));
            }
        }
        if (pass) {
            // this lane finishes its own row: box (nms.py:79 on effidehead.py:283,285), corners (:284,286)
            const unsigned anchor = (unsigned)(lv.anchor_off + pos);
            p.keys[(size_t)b * p.key_stride + slot] = make_key(score, anchor);
            p.slot_of[(size_t)b * p.A + anchor] = slot;
            const int y = pos / lv.w, x = pos - y * lv.w;
            const float ax = anchor_coord(x), ay = anchor_coord(y);
            const float4 q = decode_box(ax, ay, d[0], d[1], d[2], d[3], lv.stride);
            float4* row = reinterpret_cast<float4*>(p.rec + ((size_t)b * p.A + slot) * OUTW);
            float k[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) k[i] = decode_corner(i, ax, ay, d[4 + i], lv.stride);
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
}

// p.half_levels selects the fp16 instantiation (maps then describe FLOAT16 tensors)
cudaError_t launch_levels_filter_tma(const LevelsFilterParams& p, const DecodeMaps& maps, int num_ctas, cudaStream_t stream);

}  // namespace lp
