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
    // Disjoint boxes are the common case and K2 is issue-bound: decide them on the signs of the two
    // extents alone.  inter = max(0, dx) * max(0, dy) is +0 unless both are positive, and 0 / u is
    // 0, -0 or NaN, never > thr (thr >= 0) -- so this is exactly the reference's result.
    const float dx = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
    const float dy = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
    if (!(dx > 0.0f && dy > 0.0f)) return false;
    const float inter = __fmul_rn(dx, dy);
    if (!(inter > 0.0f)) return false;   // the product of two tiny extents may still round to zero
    const float u = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    // The IEEE division costs ~25 instructions: decide without it whenever the quotient is clearly off
    // the threshold.  With tu = RN(thr * u),
    //   inter > tu * (1 + 1e-6)  =>  inter / u > thr * (1 + 5e-7) > succ(thr)  =>  RN(inter / u) > thr
    //   inter < tu * (1 - 1e-6)  =>  inter / u < thr * (1 - 5e-7) < pred(thr)  =>  RN(inter / u) < thr
    // (two roundings of 2^-24 each against a margin of 1e-6; tu well inside the normal range, so the
    // relative bounds hold).  Only the band in between -- and the degenerate cases u <= 0, u or tu not
    // finite, denormal-sized products -- take the exact quotient.
    // succ(thr) <= thr * (1 + 2^-23) needs a normal thr as well (a denormal threshold is legal, if absurd).
    const float tu = __fmul_rn(thr_floor, u);
    if (thr_floor >= 1e-30f && tu > 1e-30f && tu < 3.0e38f) {   // implies u > 0, both finite
        if (inter > __fmul_rn(tu, 1.000001f)) return true;
        if (inter < __fmul_rn(tu, 0.999999f)) return false;
    }
    const float ovr = __fdiv_rn(inter, u);
    return ovr > thr_floor;
}

// ---- head decode arithmetic (effidehead.py:283-286) ------------------------------------------------
// anchor point of grid cell (x, y): generate_anchors(is_eval=True), anchor_generator.py:13-14
__device__ __forceinline__ float anchor_coord(int cell) { return __fadd_rn((float)cell, 0.5f); }
// dist2bbox 'xywh' (general.py:31-38) followed by *= stride (effidehead.py:285): returns cx, cy, w, h
__device__ __forceinline__ float4 decode_box(float ax, float ay, float l, float t, float r, float b, float stride) {
    const float x1 = __fsub_rn(ax, l), y1 = __fsub_rn(ay, t), x2 = __fadd_rn(ax, r), y2 = __fadd_rn(ay, b);
    return make_float4(__fmul_rn(__fmul_rn(__fadd_rn(x1, x2), 0.5f), stride), __fmul_rn(__fmul_rn(__fadd_rn(y1, y2), 0.5f), stride),
                       __fmul_rn(__fsub_rn(x2, x1), stride), __fmul_rn(__fsub_rn(y2, y1), stride));
}
// dist2cor (general.py:51-66) followed by *= stride (effidehead.py:286): corner k of 8 (TL, BL, BR, TR as x,y pairs)
__device__ __forceinline__ float decode_corner(int k, float ax, float ay, float d, float stride) {
    // signs: x: - - + +   y: - + + -
    const bool is_y = k & 1;
    const bool plus = is_y ? (k == 3 || k == 5) : (k >= 4);
    const float a = is_y ? ay : ax;
    return __fmul_rn(plus ? __fadd_rn(a, d) : __fsub_rn(a, d), stride);
}

// sigmoid(x) = 1 / (1 + 2^(-x log2 e)) on the SFU (MUFU.EX2 + MUFU.RCP): relative error <= ~2.5e-6
// for |x| <= 30 (ex2.approx 2^-22.5, the rounded exponent |x| * 6e-8, rcp.approx 1 ulp), inside the
// 1e-5 bar of the path; larger magnitudes (saturated scores, denormal results) take the libm route.
// Verified monotone non-decreasing over every finite fp32 input (tools/sigmoid_monotone.py), which
// is what makes max_j sigmoid(x_j) == sigmoid(max_j x_j) exact in the fused path.
static __device__ __noinline__ float sigmoid_slow(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }
// the SFU route alone (valid for |x| <= 30), branch-free
__device__ __forceinline__ float sigmoid_sfu(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmul_rn(x, -1.4426950408889634f)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fadd_rn(1.0f, e)));
    return r;
}
__device__ __forceinline__ float sigmoid_f32(float x) {
    float r = sigmoid_sfu(x);
    if (fabsf(x) > 30.0f) r = sigmoid_slow(x);
    return r;
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// for roles that are off the critical path: a spinning try_wait competes for issue slots with the
// warps that do the work
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(128);
}
// Warp-collective wait with ONE poller: lane 0 polls, a vote ends the loop and -- the point -- leaves the
// warp converged.  `if (lane == 0) mbar_wait(...); __syncwarp();` does not: WARPSYNC synchronises the
// lanes but they keep executing as two groups afterwards (ncu on round 2's KF: lane 0 ran whole tiles
// apart from lanes 1-31, avg 20 threads per instruction, every later ballot / shuffle on the slow
// divergent path).  All 32 lanes must call it.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int lane) {
    for (;;) {
        bool ok = false;
        if (lane == 0) ok = mbar_try_wait(bar, parity);
        if (__any_sync(0xffffffffu, ok)) break;
    }
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

// 3-D TMA box load (SASS: UTMALDG): the box lands densely, innermost dimension first, at dst_smem
__device__ __forceinline__ void tma_load_3d(void* dst_smem, const void* tensor_map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst_smem)),
        "l"(reinterpret_cast<uint64_t>(tensor_map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// Host side: opt a kernel into its dynamic shared-memory size once per device (the call is idempotent,
// so a race between two host threads only repeats it).
template <class Kernel>
inline cudaError_t configure_smem_once(Kernel kernel, int bytes, bool (&done)[64]) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
    return e;
}

}  // namespace lp
