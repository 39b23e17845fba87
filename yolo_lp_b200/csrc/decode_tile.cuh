// Shared by the two decode kernels (decode.cu: LSU loads, decode_tma.cu: TMA loads + warp
// specialisation): the out-tile geometry, the bulk-store PTX and the per-tile transposition.
#pragma once
#include <cuda_fp16.h>

#include "level_tiles.cuh"

namespace lp {

constexpr int DEC_THREADS = LT_THREADS;              // threads that transpose (the consumers)
constexpr int DEC_WARPS = DEC_THREADS / 32;
constexpr int OUT_FLOATS = DEC_TILE * ROW;           // [32 positions][290 columns]

__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// A class score as the head tensor holds it.  half_scores: rounded to the nearest half and widened again --
// what the reference's model.half() forward stores (torch.sigmoid of a half tensor computes in fp32 and
// rounds once; torch.cat with the fp32 geometry columns then promotes it, effidehead.py:251-258,288-301).
__device__ __forceinline__ float class_score(float logit, bool half_scores) {
    const float s = sigmoid_f32(logit);
    return half_scores ? __half2float(__float2half_rn(s)) : s;
}

// One staged tile (channel-major, 32 positions per channel row) -> finished rows in `outt`.
__device__ __forceinline__ void transpose_tile(const float* stage, float* outt, const TileInfo& t, const DecodeLevel& lv,
                                               int warp, int lane, bool half_scores) {
    // class columns: lanes along positions (conflict-free stage reads), sigmoid, transposed write
    // Two adjacent columns per lane and one 64-bit store: with the packed 290-word row pitch a
    // 32-bit store per lane is a 2-way bank conflict (290 = 2 mod 32), a 64-bit one is conflict-free
    // (145 = 1 mod 16 eight-byte units per half-warp phase) and halves the store count.
    if (lane < t.n) {
        float* orow = outt + lane * ROW;
#pragma unroll 4
        for (int col = 14 + 2 * warp; col < ROW; col += 2 * DEC_WARPS) {
            float2 v;
            v.x = class_score(stage[col * DEC_TILE + lane], half_scores);
            v.y = class_score(stage[(col + 1) * DEC_TILE + lane], half_scores);
            *reinterpret_cast<float2*>(orow + col) = v;
        }
        if (warp == 0) orow[13] = class_score(stage[13 * DEC_TILE + lane], half_scores);
    }
    // box / objectness / corner columns: one thread per position
    if (warp == DEC_WARPS - 1 && lane < t.n) {
        const int pos = t.p0 + lane;
        const int y = pos / lv.w, x = pos - y * lv.w;
        const float ax = anchor_coord(x), ay = anchor_coord(y);
        const float sd = lv.stride;
        const float* st = stage + lane;
        float* r = outt + lane * ROW;
        const float4 bx = decode_box(ax, ay, st[0 * DEC_TILE], st[1 * DEC_TILE], st[2 * DEC_TILE], st[3 * DEC_TILE], sd);
        r[0] = bx.x; r[1] = bx.y; r[2] = bx.z; r[3] = bx.w;
        r[4] = 1.0f;                                             // effidehead.py:290
#pragma unroll
        for (int k = 0; k < 8; ++k) r[5 + k] = decode_corner(k, ax, ay, st[(5 + k) * DEC_TILE], sd);
    }
}

cudaError_t launch_decode_tma(const DecodeParams& p, const DecodeMaps& maps, int num_sms, cudaStream_t stream);

}  // namespace lp
