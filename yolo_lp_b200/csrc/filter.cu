// K1 -- fused score / threshold / compact over the head tensor pred[B*A, 290] (fp32).
//
// Restates, for every anchor row, nms.py:76 (cls *= obj), :81-88 (eight group maxima),
// :90-91 (the buggy 8-term filter mean) and :120 (the NMS score), and emits one 64-bit sort key
// per surviving row.  Nothing else is written: boxes, corners and argmaxes of the (few) kept
// rows are recomputed by K2 straight from `pred`, so this kernel's HBM traffic is the
// algorithmic minimum -- every row read exactly once, 8 B written per survivor.
//
// Data movement (the roofline kernel, HBM-bound):
//   * `pred` is treated as one flat array of rows.  A tile is 32 consecutive rows = 37 120 B,
//     which is 128-B aligned for any A (row pitch 1160 B is only 8-B aligned, so single rows
//     cannot be moved with 128-bit/TMA transfers; 32-row tiles can).
//   * every warp owns one smem stage and streams tiles into it with ONE TMA 1-D bulk copy
//     (cp.async.bulk + mbarrier complete_tx; SASS UBLKCP), L2 evict-first.  6 warps x 37 KB =
//     217.5 KB in flight per SM, no register staging, no producer warp.
//   * compute is one lane per row out of shared memory with 64-bit loads: row pitch 290 words
//     == 145 8-byte units, and 145 mod 16 == 1, so the 16 lanes of each LDS.64 phase hit 16
//     distinct bank pairs -- conflict-free without padding.
//   * survivors are compacted with warp ballots + one atomicAdd per (warp, image); slot order
//     is irrelevant because the key carries the anchor index (sort restores reference order).
#include "kernels.cuh"

namespace lp {

constexpr int TILE_ROWS = 32;
constexpr int TILE_BYTES = TILE_ROWS * ROW * 4;  // 37120
constexpr int FILTER_WARPS = 6;
constexpr int FILTER_THREADS = FILTER_WARPS * 32;
constexpr int FILTER_SMEM = FILTER_WARPS * TILE_BYTES + FILTER_WARPS * 8;

__device__ __forceinline__ void issue_tile(const FilterParams& p, unsigned tile, float* buf, uint64_t* bar, uint64_t policy) {
    const unsigned row0 = tile * TILE_ROWS;
    const unsigned rows = min((unsigned)TILE_ROWS, p.total_rows - row0);
    const unsigned bytes = rows * (ROW * 4);
    const unsigned bulk = bytes & ~15u;  // rows odd -> 8 trailing bytes moved by hand
    const char* src = reinterpret_cast<const char*>(p.pred) + (size_t)row0 * (ROW * 4);
    mbar_expect_tx(bar, bulk);
    bulk_g2s(buf, src, bulk, bar, policy);
    if (bulk != bytes) {
        const float2 t = *reinterpret_cast<const float2*>(src + bulk);
        *reinterpret_cast<float2*>(reinterpret_cast<char*>(buf) + bulk) = t;
    }
}

__global__ void __launch_bounds__(FILTER_THREADS, 1) filter_kernel(const FilterParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* buf = reinterpret_cast<float*>(smem + warp * TILE_BYTES);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + FILTER_WARPS * TILE_BYTES) + warp;

    // Tile scheduling: the first tile of every warp is static (its global warp id), all later
    // ones are claimed from a global counter, one claim ahead of use so the atomic's round trip
    // hides behind the load in flight.  Dynamic claims keep every SM busy until the last tile and
    // let late-starting CTAs (SMs still occupied by the previous batch's K2) simply take less.
    const unsigned n_warps = gridDim.x * FILTER_WARPS;
    unsigned tile = blockIdx.x * FILTER_WARPS + warp;
    unsigned next = 0xffffffffu;
    uint64_t policy = 0;
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        policy = l2_evict_first_policy();
        if (tile < p.n_tiles) {
            issue_tile(p, tile, buf, bar, policy);
            next = n_warps + atomicAdd(p.tile_counter, 1u);
        }
    }
    __syncwarp();

    uint32_t parity = 0;
    while (tile < p.n_tiles) {
        mbar_wait(bar, parity);
        parity ^= 1;
        __syncwarp();  // orders lane 0's hand-copied tail (odd last tile) before the reads

        const unsigned row = tile * TILE_ROWS + lane;
        const bool valid = row < p.total_rows;
        float filt = 0.0f, score = 0.0f;
        {
            // lanes past the end of a ragged last tile read stale (finite or not) smem; masked below
            const float2* r = reinterpret_cast<const float2*>(buf + lane * ROW);
            const float obj = r[2].x;  // column 4
            float m[NGROUP];
#pragma unroll
            for (int g = 0; g < NGROUP; ++g) m[g] = -INFINITY;
#pragma unroll
            for (int pi = 6; pi < ROW / 2; ++pi) {  // pairs covering columns 12..289
                const float2 v = r[pi];
                if (2 * pi >= 13) m[group_of(2 * pi)] = fmaxf(m[group_of(2 * pi)], __fmul_rn(v.x, obj));
                m[group_of(2 * pi + 1)] = fmaxf(m[group_of(2 * pi + 1)], __fmul_rn(v.y, obj));
            }
            lp_means(m, filt, score);
        }
        // all lanes have consumed the stage: hand it back to the async proxy and refill
        __syncwarp();
        next = __shfl_sync(0xffffffffu, next, 0);
        unsigned claim = 0xffffffffu;
        if (lane == 0 && next < p.n_tiles) {
            fence_proxy_async_smem();
            issue_tile(p, next, buf, bar, policy);
            claim = n_warps + atomicAdd(p.tile_counter, 1u);
        }

        const bool pass = valid && (filt >= p.conf);
        const unsigned img = row / p.A;
        const unsigned anchor = row - img * p.A;
        unsigned todo = __ballot_sync(0xffffffffu, pass);
        while (todo) {  // at most two images per tile unless A < 32
            const int leader = __ffs(todo) - 1;
            const unsigned limg = __shfl_sync(0xffffffffu, img, leader);
            const unsigned grp = __ballot_sync(0xffffffffu, pass && img == limg);
            int base = 0;
            if (lane == leader) base = atomicAdd(p.counts + limg, __popc(grp));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (pass && img == limg) {
                const unsigned slot = base + __popc(grp & ((1u << lane) - 1u));
                p.keys[(size_t)limg * p.key_stride + slot] = make_key(score, anchor);
            }
            todo &= ~grp;
        }
        tile = next;
        next = claim;
    }
}

cudaError_t launch_filter(const FilterParams& p, int num_sms, cudaStream_t stream) {
    static_assert(FILTER_SMEM <= 227 * 1024, "filter stages exceed shared memory");
    static bool configured[64] = {false};  // per device; idempotent, so a race only repeats the call
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FILTER_SMEM);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    unsigned grid = (p.n_tiles + FILTER_WARPS - 1) / FILTER_WARPS;
    if (grid > (unsigned)num_sms) grid = num_sms;
    if (grid == 0) return cudaSuccess;
    filter_kernel<<<grid, FILTER_THREADS, FILTER_SMEM, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace lp
