// K1 for fp16 head tensors (`--half`, inferer.py:46-50, evaler.py:116):
// pred[B*A, 290] stored as IEEE half, every value upcast exactly to fp32 on load, then the same
// arithmetic as filter.cu -- so the result is bit for bit the fp32 path's on `pred.float()`, which is
// the parity contract of this row (the reference's own half mode runs torchvision's CUDA kernel in
// half arithmetic behind an unstable sort).  Half the bytes of the HBM-bound stage.
//
// Same data movement as filter.cu with the row pitch halved: a tile is 64 consecutive rows =
// 37 120 B (again a multiple of 128 B for any A; single 580-byte rows are only 4-byte aligned), one
// TMA 1-D bulk copy per tile into the warp's own stage, six stages per SM.  A lane owns rows l and
// l + 32 of the tile: row pitch 145 words and 145 mod 32 = 17 is odd, so the 32 lanes of an LDS.32
// hit 32 distinct banks.  Each 32-bit load carries two columns.
#include <cuda_fp16.h>

#include "kernels.cuh"

namespace lp {

constexpr int HTILE_ROWS = 64;
constexpr int HROW_BYTES = ROW * 2;                    // 580
constexpr int HTILE_BYTES = HTILE_ROWS * HROW_BYTES;   // 37120
constexpr int HFILTER_WARPS = 6;
constexpr int HFILTER_THREADS = HFILTER_WARPS * 32;
constexpr int HFILTER_SMEM = HFILTER_WARPS * HTILE_BYTES + HFILTER_WARPS * 8;

__device__ __forceinline__ void issue_tile_h(const FilterParams& p, unsigned tile, unsigned char* buf, uint64_t* bar,
                                             uint64_t policy) {
    const unsigned row0 = tile * HTILE_ROWS;
    const unsigned rows = min((unsigned)HTILE_ROWS, p.total_rows - row0);
    const unsigned bytes = rows * HROW_BYTES;
    const unsigned bulk = bytes & ~15u;  // rows % 4 != 0 -> up to 12 trailing bytes moved by hand
    const char* src = reinterpret_cast<const char*>(p.pred) + (size_t)row0 * HROW_BYTES;
    mbar_expect_tx(bar, bulk);
    bulk_g2s(buf, src, bulk, bar, policy);
    for (unsigned o = bulk; o < bytes; o += 4)
        *reinterpret_cast<unsigned*>(buf + o) = *reinterpret_cast<const unsigned*>(src + o);
}

// the eight group maxima of one row (columns 13..289 times the objectness in column 4), from 145 words
__device__ __forceinline__ void row_scores(const unsigned* r, float& filt, float& score) {
    const float obj = __low2float(*reinterpret_cast<const __half2*>(r + 2));   // column 4 = low half of word 2
    float m[NGROUP];
#pragma unroll
    for (int g = 0; g < NGROUP; ++g) m[g] = -INFINITY;
#pragma unroll
    for (int pi = 6; pi < ROW / 2; ++pi) {  // pairs covering columns 12..289
        const float2 v = __half22float2(*reinterpret_cast<const __half2*>(r + pi));
        if (2 * pi >= 13) m[group_of(2 * pi)] = fmaxf(m[group_of(2 * pi)], __fmul_rn(v.x, obj));
        m[group_of(2 * pi + 1)] = fmaxf(m[group_of(2 * pi + 1)], __fmul_rn(v.y, obj));
    }
    lp_means(m, filt, score);
}

__global__ void __launch_bounds__(HFILTER_THREADS, 1) filter_half_kernel(const FilterParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* buf = smem + warp * HTILE_BYTES;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + HFILTER_WARPS * HTILE_BYTES) + warp;

    // tile scheduling as in filter.cu: first tile static, later ones claimed one ahead of use
    const unsigned n_warps = gridDim.x * HFILTER_WARPS;
    unsigned tile = blockIdx.x * HFILTER_WARPS + warp;
    unsigned next = 0xffffffffu;
    uint64_t policy = 0;
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        policy = l2_evict_first_policy();
        if (tile < p.n_tiles) {
            issue_tile_h(p, tile, buf, bar, policy);
            next = n_warps + atomicAdd(p.tile_counter, 1u);
        }
    }
    __syncwarp();

    uint32_t parity = 0;
    while (tile < p.n_tiles) {
        mbar_wait(bar, parity);
        parity ^= 1;
        __syncwarp();  // orders lane 0's hand-copied tail (ragged last tile) before the reads

        // lanes past the end of a ragged last tile read stale smem; masked below
        float filt[2], score[2];
#pragma unroll
        for (int h = 0; h < 2; ++h)
            row_scores(reinterpret_cast<const unsigned*>(buf + (size_t)(lane + 32 * h) * HROW_BYTES), filt[h], score[h]);
        // all lanes have consumed the stage: hand it back to the async proxy and refill
        __syncwarp();
        next = __shfl_sync(0xffffffffu, next, 0);
        unsigned claim = 0xffffffffu;
        if (lane == 0 && next < p.n_tiles) {
            fence_proxy_async_smem();
            issue_tile_h(p, next, buf, bar, policy);
            claim = n_warps + atomicAdd(p.tile_counter, 1u);
        }

#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const unsigned row = tile * HTILE_ROWS + 32 * h + lane;
            const bool pass = row < p.total_rows && (filt[h] >= p.conf);
            const unsigned img = row / p.A;
            const unsigned anchor = row - img * p.A;
            unsigned todo = __ballot_sync(0xffffffffu, pass);
            while (todo) {  // at most two images per half tile unless A < 32
                const int leader = __ffs(todo) - 1;
                const unsigned limg = __shfl_sync(0xffffffffu, img, leader);
                const unsigned grp = __ballot_sync(0xffffffffu, pass && img == limg);
                int base = 0;
                if (lane == leader) base = atomicAdd(p.counts + limg, __popc(grp));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (pass && img == limg) {
                    const unsigned slot = base + __popc(grp & ((1u << lane) - 1u));
                    p.keys[(size_t)limg * p.key_stride + slot] = make_key(score[h], anchor);
                }
                todo &= ~grp;
            }
        }
        tile = next;
        next = claim;
    }
}

cudaError_t launch_filter_half(const FilterParams& p, int num_sms, cudaStream_t stream) {
    static_assert(HFILTER_SMEM <= 227 * 1024, "filter stages exceed shared memory");
    static bool configured[64] = {false};
    cudaError_t e = configure_smem_once(filter_half_kernel, HFILTER_SMEM, configured);
    if (e != cudaSuccess) return e;
    const unsigned n_tiles = (p.total_rows + HTILE_ROWS - 1) / HTILE_ROWS;
    FilterParams q = p;
    q.n_tiles = n_tiles;
    unsigned grid = (n_tiles + HFILTER_WARPS - 1) / HFILTER_WARPS;
    if (grid > (unsigned)num_sms) grid = num_sms;
    if (grid == 0) return cudaSuccess;
    filter_half_kernel<<<grid, HFILTER_THREADS, HFILTER_SMEM, stream>>>(q);
    return cudaGetLastError();
}

}  // namespace lp
