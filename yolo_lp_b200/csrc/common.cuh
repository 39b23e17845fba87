// Shared device helpers for liblpnms (sm_100a only).
//
// Bit-exactness rules (SURVEY.md §7 "hard parts"): the library is compiled with
// -fmad=false and every value that feeds a comparison is built from the
// __f*_rn intrinsics in the reference's association order, so no FMA
// contraction can change a kept set.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lpnms.h"

namespace lp {

constexpr int ROW = LP_ROW;      // 290 floats, 1160 B per anchor row
constexpr int OUTW = LP_OUT;     // 28 floats per detection
constexpr int NGROUP = 8;

// Column groups of a head row: province | alphabet | six characters (nms.py:81-88).
__host__ __device__ constexpr int group_begin(int g) {
    return g == 0 ? 13 : g == 1 ? 44 : g == 2 ? 68 : g == 3 ? 105 : g == 4 ? 142 : g == 5 ? 179 : g == 6 ? 216 : g == 7 ? 253 : 290;
}
__host__ __device__ constexpr int group_of(int col) {
    return col < 44 ? 0 : col < 68 ? 1 : col < 105 ? 2 : col < 142 ? 3 : col < 179 ? 4 : col < 216 ? 5 : col < 253 ? 6 : 7;
}

// ---- order-preserving float <-> uint key -----------------------------------------------------
// ascending unsigned order == ascending float order (all finite values, -0 mapped before +0).
__device__ __forceinline__ uint32_t float_sortable(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// 64-bit candidate key: ascending key == (score descending, anchor ascending), which is the
// order torchvision's stable descending sort produces on the anchor-ordered compaction
// (nms.py:97,121).
__device__ __forceinline__ unsigned long long make_key(float score, uint32_t anchor) {
    // score + 0.0f folds -0.0 into +0.0 so the two compare equal like in a float sort
    return ((unsigned long long)(~float_sortable(__fadd_rn(score, 0.0f))) << 32) | anchor;
}

// ---- the 8-term means (nms.py:90-91 and :120) --------------------------------------------------
// filter mean adds ad4 (c[6]) twice and never ad5 (c[7]) -- a reference bug that must be kept;
// the NMS score adds all eight.  Both are left-to-right fp32 sums divided by 8.
__device__ __forceinline__ void lp_means(const float (&c)[NGROUP], float& filt, float& score) {
    float s = __fadd_rn(c[0], c[1]);
    s = __fadd_rn(s, c[2]);
    s = __fadd_rn(s, c[3]);
    s = __fadd_rn(s, c[4]);
    s = __fadd_rn(s, c[5]);
    s = __fadd_rn(s, c[6]);
    filt = __fdiv_rn(__fadd_rn(s, c[6]), 8.0f);
    score = __fdiv_rn(__fadd_rn(s, c[7]), 8.0f);
}

// ---- box helpers ---------------------------------------------------------------------------------
// xywh2xyxy, nms.py:21-28 (w/2 is exact, so *0.5f is the same fp32 value)
__device__ __forceinline__ float4 xywh_to_xyxy(float cx, float cy, float w, float h) {
    const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);
    return make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
}
__device__ __forceinline__ float box_area(const float4& b) {
    return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}
// torchvision CPU nms_kernel_impl: ovr = inter / (area_i + area_j - inter); suppress iff
// (double)ovr > thr.  `thr_floor` is the largest float <= thr, for which the float compare
// ovr > thr_floor is equivalent.  0/0 -> NaN -> not suppressed.
__device__ __forceinline__ bool iou_exceeds(const float4& a, float area_a, const float4& b, float area_b, float thr_floor) {
    const float w = fmaxf(0.0f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
    const float h = fmaxf(0.0f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
    const float inter = __fmul_rn(w, h);
    // inter is +0 for disjoint boxes (the common case): 0/u is 0, -0 or NaN, never > thr (thr >= 0),
    // so the IEEE division -- whose zero-numerator case runs the slow path -- is skipped exactly.
    if (!(inter > 0.0f)) return false;
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
    return ovr > thr_floor;
}

// Inferer.rescale on one coordinate (inferer.py:210-225) + optional caller .round() (:100)
__device__ __forceinline__ float rescale_coord(float v, float pad, float ratio, float hi, int do_round) {
    v = __fdiv_rn(__fsub_rn(v, pad), ratio);
    v = fminf(fmaxf(v, 0.0f), hi);
    return do_round ? rintf(v) : v;
}

// ---- mbarrier / bulk-copy (TMA 1-D) PTX ----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// generic-proxy reads of a buffer must be ordered before the async proxy overwrites it
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// global -> shared bulk async copy (SASS: UBLKCP); bytes % 16 == 0, both addresses 16-B aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

}  // namespace lp
