// Shared machinery of the kernels that stream raw per-level conv outputs (NCHW) tile by tile:
// a tile is DEC_TILE = 32 consecutive positions of one level of one image, staged column-major
// ([290 output columns][32 positions], column 4 unused) in shared memory by 16-byte cp.async copies
// issued by all LT_THREADS threads.  Used by decode.cu and fused.cu.
#pragma once
#include "kernels.cuh"

namespace lp {

constexpr int LT_THREADS = 512;
constexpr int STAGE_FLOATS = ROW * DEC_TILE;         // [290 columns][32 positions], column 4 unused
constexpr int LT_SLOTS = ((ROW - 1) * (DEC_TILE / 4) + LT_THREADS - 1) / LT_THREADS;  // 16-byte copies per thread and tile

struct TileInfo {
    int b, l, p0, n;
};

// Walks the tiles tile0, tile0 + step, ... of one CTA without dividing: (image, tile-in-image).
struct TileWalker {
    int b, r;
    template <class P>
    __device__ __forceinline__ void init(const P& p, int tile) {
        b = tile / p.tiles_per_image;
        r = tile - b * p.tiles_per_image;
    }
    template <class P>
    __device__ __forceinline__ void advance(const P& p, int step) {
        r += step;
        while (r >= p.tiles_per_image) {
            r -= p.tiles_per_image;
            ++b;
        }
    }
    template <class P>
    __device__ __forceinline__ TileInfo info(const P& p) const {
        TileInfo t;
        t.b = b;
        t.l = 0;
#pragma unroll
        for (int i = 1; i < LP_MAX_LEVELS; ++i)
            if (i < p.n_levels && r >= p.lv[i].tile_off) t.l = i;
        t.p0 = (r - p.lv[t.l].tile_off) * DEC_TILE;
        t.n = min(DEC_TILE, p.lv[t.l].hw - t.p0);
        return t;
    }
};

__device__ __forceinline__ void cp_async_16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Which source tensor / channel feeds output column `col` (col != 4): tensor 0..7 = class groups,
// 8 = reg, 9 = cor.  Fixed per thread and copy slot, so it is decoded once, outside the tile loop.
__device__ __forceinline__ void column_source(int col, int& tensor, int& ch, int& width) {
    if (col < 4) { tensor = 8; ch = col; width = 4; }
    else if (col < 13) { tensor = 9; ch = col - 5; width = 8; }
    else { tensor = group_of(col); ch = col - group_begin(tensor); width = group_begin(tensor + 1) - group_begin(tensor); }
}
__device__ __forceinline__ const float* tensor_base(const DecodeLevel& lv, int tensor) {
    return tensor == 8 ? lv.reg : tensor == 9 ? lv.cor : lv.cls[tensor];
}


// all threads: queue the copies that bring tile `t` into `stage` (column-major [col][32]).
// Element offsets inside one source tensor fit 32 bits (checked by lp_detect_decode_f32).
template <class P>
__device__ __forceinline__ void load_tile(const P& p, const TileInfo& t, float* stage, int tid,
                                          const unsigned (&slot)[LT_SLOTS]) {
    const DecodeLevel& lv = p.lv[t.l];
    if (p.bulk_in) {  // rows 16-byte aligned: 8 chunks of 4 positions per row
        const unsigned hw = (unsigned)lv.hw, b = (unsigned)t.b, p0 = (unsigned)t.p0;
#pragma unroll
        for (int k = 0; k < LT_SLOTS; ++k) {
            const unsigned sl = slot[k];  // col | chunk << 9 | tensor << 13 | ch << 17 | width << 23, ~0 = none
            if (sl == 0xffffffffu) continue;
            const unsigned col = sl & 511, chunk4 = ((sl >> 9) & 15) * 4, tensor = (sl >> 13) & 15, ch = (sl >> 17) & 63,
                           width = sl >> 23;
            if ((int)chunk4 < t.n)
                cp_async_16(stage + col * DEC_TILE + chunk4, tensor_base(lv, tensor) + ((b * width + ch) * hw + p0 + chunk4));
        }
    } else {
        for (int ci = tid; ci < (ROW - 1) * DEC_TILE; ci += LT_THREADS) {
            const int r = ci / DEC_TILE, q = ci - r * DEC_TILE;
            const int col = r + (r >= 4);
            if (q < t.n) cp_async_4(stage + col * DEC_TILE + q, column_src(lv, t.b, col) + t.p0 + q);
        }
    }
}


// per-thread copy slots: col | chunk << 9 | tensor << 13 | ch << 17 | width << 23, ~0 = none
__device__ __forceinline__ void make_slots(unsigned (&slot)[LT_SLOTS], int tid) {
#pragma unroll
    for (int k = 0; k < LT_SLOTS; ++k) {
        const int ci = tid + k * LT_THREADS;
        slot[k] = 0xffffffffu;
        if (ci < (ROW - 1) * (DEC_TILE / 4)) {
            const int r = ci / (DEC_TILE / 4), chunk = ci - r * (DEC_TILE / 4);
            const int col = r + (r >= 4);
            int tensor, ch, width;
            column_source(col, tensor, ch, width);
            slot[k] = (unsigned)col | (unsigned)chunk << 9 | (unsigned)tensor << 13 | (unsigned)ch << 17 | (unsigned)width << 23;
        }
    }
}

}  // namespace lp
