// Detect.forward eval tail (yolov6/models/effidehead.py:247-301), TMA + warp-specialised variant
// of decode.cu for level planes whose rows are 16-byte aligned (h*w % 4 == 0, every letterboxed
// YOLO input).  Same arithmetic, same output bits: the transposition is the shared transpose_tile.
//
// A phase profile of the all-threads-do-everything kernel (clock64, per 32-position tile) showed the
// serial sections, not HBM, set its pace: ~1400 cycles issuing the tile's loads (18 LDGSTS per
// thread, or ten UTMALDG from one thread), ~2000 transposing (MUFU-bound: 2 SFU ops per sigmoid),
// ~370 issuing the bulk store -- one after the other between CTA-wide barriers, 3900 cycles per tile
// against a ~2900-cycle HBM budget.  Here the three jobs run concurrently:
//   warp 16, lane 0   load producer: per tile ten 3-D TMA box loads (UTMALDG.3D, one per source
//                     tensor: [c channels][32 positions] of image b) into a 3-deep stage ring,
//                     completion on the stage's `full` mbarrier, reuse gated by `empty`;
//   warp 17, lane 0   store producer: per tile one TMA bulk store (UBLKCP) of the finished rows from a
//                     3-deep out ring, gated by `ofull`; frees buffers through `oempty` once
//                     cp.async.bulk.wait_group.read says the engine has read them;
//   warp 18           release warp: waits on the tile's named barrier for the sixteen consumer warps,
//                     then arrives on `empty` (stage free) and `ofull` (out tile ready);
//   warps 0-15        consumers: wait full/oempty, transpose + sigmoid + box/corner decode, then
//                     bar.arrive (non-blocking) on the tile's named barrier and straight on to the next
//                     tile: no consumer ever blocks on another, the rings bound the drift.
// (Having each consumer warp arrive on the mbarriers itself was slower: mbarrier.arrive after the
// proxy fence costs several hundred cycles in the arriving warp.)
// Positions past the end of a level are zero-filled by the TMA unit (and never stored).
// With every h*w a multiple of 4 (the precondition for the tensor maps) each tile's first row index
// and row count are even, so every tile's destination and size are 16-byte multiples (row pitch
// 1160 B = 8 mod 16) and the bulk store needs no fallback; `out` itself must be 16-byte aligned.
#include "decode_tile.cuh"

namespace lp {

constexpr int DT_STAGES = 3;
constexpr int DT_OUTS = 3;
constexpr int DT_THREADS = DEC_THREADS + 96;   // + load producer, store producer and release warps
constexpr int DT_SMEM = (DT_STAGES * STAGE_FLOATS + DT_OUTS * OUT_FLOATS) * 4 + (2 * DT_STAGES + 2 * DT_OUTS) * 8;

// first output column fed by source tensor k (0..7 class groups, 8 reg, 9 cor) and its channel count
__device__ __forceinline__ int tensor_first_col(int k) { return k == 8 ? 0 : k == 9 ? 5 : group_begin(k); }

// Named barriers: 1..3 = "tile it%3 is transposed" (consumers arrive without blocking, the release warp
// waits).  Three of them because a fast warp may run up to two tiles ahead of a slow one -- not three:
// the stage ring stops it.
__device__ __forceinline__ void tile_done_arrive(int it) {
    asm volatile("bar.arrive %0, %1;" ::"r"(1 + it % 3), "n"(DEC_THREADS + 32) : "memory");
}
__device__ __forceinline__ void tile_done_wait(int it) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + it % 3), "n"(DEC_THREADS + 32) : "memory");
}

#ifdef LP_DEC_PROFILE
#define LP_PF_DECL long long pf[4] = {0, 0, 0, 0}, pf_t = clock64()
#define LP_PF(k) do { const long long t1_ = clock64(); pf[k] += t1_ - pf_t; pf_t = t1_; } while (0)
#define LP_PF_OUT(role) do { if (p.timing != nullptr) for (int k = 0; k < 4; ++k) p.timing[(blockIdx.x * 3 + role) * 4 + k] = pf[k]; } while (0)
#else
#define LP_PF_DECL do { } while (0)
#define LP_PF(k) do { } while (0)
#define LP_PF_OUT(role) do { } while (0)
#endif

__global__ void __launch_bounds__(DT_THREADS, 1) decode_tma_kernel(const DecodeParams p, const __grid_constant__ DecodeMaps maps) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* stage0 = reinterpret_cast<float*>(smem);
    float* out0 = stage0 + DT_STAGES * STAGE_FLOATS;
    uint64_t* full = reinterpret_cast<uint64_t*>(out0 + DT_OUTS * OUT_FLOATS);
    uint64_t* empty = full + DT_STAGES;
    uint64_t* ofull = empty + DT_STAGES;
    uint64_t* oempty = ofull + DT_OUTS;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < DT_STAGES; ++k) { mbar_init(&full[k], 1); mbar_init(&empty[k], 1); }
#pragma unroll
        for (int k = 0; k < DT_OUTS; ++k) { mbar_init(&ofull[k], 1); mbar_init(&oempty[k], 1); }
        mbar_fence_init();
    }
    __syncthreads();   // the only CTA-wide barrier: the roles never meet again

    const int first = blockIdx.x, step = gridDim.x;
    const int n_my = first < p.n_tiles ? (p.n_tiles - first + step - 1) / step : 0;
    TileWalker w;
    w.init(p, first);

    if (warp == DEC_WARPS) {            // ---- load producer
        if (lane != 0) return;
        LP_PF_DECL;
        for (int it = 0; it < n_my; ++it, w.advance(p, step)) {
            const int s = it % DT_STAGES, use = it / DT_STAGES;
            LP_PF(0);
            if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
            LP_PF(1);
            const TileInfo t = w.info(p);
            float* stage = stage0 + s * STAGE_FLOATS;
            fence_proxy_async_smem();   // the stage was last read through the generic proxy
            mbar_expect_tx(&full[s], (ROW - 1) * DEC_TILE * 4);
#pragma unroll
            for (int k = 0; k < DEC_TENSORS; ++k)
                tma_load_3d(stage + tensor_first_col(k) * DEC_TILE, &maps.m[t.l][k], t.p0, 0, t.b, &full[s]);
            LP_PF(2);
        }
        LP_PF_OUT(0);
    } else if (warp == DEC_WARPS + 1) { // ---- store producer
        if (lane != 0) return;
        LP_PF_DECL;
        for (int it = 0; it < n_my; ++it, w.advance(p, step)) {
            const int o = it % DT_OUTS;
            const TileInfo t = w.info(p);
            float* dst = p.out + ((size_t)t.b * p.A + p.lv[t.l].anchor_off + t.p0) * ROW;
            const uint32_t bytes = (uint32_t)t.n * (ROW * 4);
            LP_PF(0);
            mbar_wait(&ofull[o], (it / DT_OUTS) & 1);
            LP_PF(1);
            bulk_s2g(dst, out0 + o * OUT_FLOATS, bytes);   // 16-byte aligned: see the note on top
            bulk_commit();
            LP_PF(2);
            bulk_wait_read1();          // the store of tile it-1 has finished reading its buffer
            if (it > 0) mbar_arrive(&oempty[(it - 1) % DT_OUTS]);
            LP_PF(3);
        }
        bulk_wait0();
        LP_PF_OUT(1);
    } else if (warp == DEC_WARPS + 2) { // ---- release warp
        for (int it = 0; it < n_my; ++it) {
            tile_done_wait(it);         // all sixteen consumer warps have finished tile `it`
            if (lane == 0) {
                mbar_arrive(&empty[it % DT_STAGES]);
                mbar_arrive(&ofull[it % DT_OUTS]);
            }
        }
    } else {                            // ---- consumers
        LP_PF_DECL;
        for (int it = 0; it < n_my; ++it, w.advance(p, step)) {
            const int s = it % DT_STAGES, o = it % DT_OUTS;
            const float* stage = stage0 + s * STAGE_FLOATS;
            float* outt = out0 + o * OUT_FLOATS;
            const TileInfo t = w.info(p);
            const DecodeLevel& lv = p.lv[t.l];
            LP_PF(3);
            // one poller per warp (512 threads spinning on try_wait slowed the arrivals on the same barriers),
            // through mbar_wait_warp so that the warp is converged again when it starts on the tile
            if (it >= DT_OUTS) mbar_wait_warp(&oempty[o], (it / DT_OUTS - 1) & 1, lane);
            LP_PF(0);
            mbar_wait_warp(&full[s], (it / DT_STAGES) & 1, lane);
            LP_PF(1);
            transpose_tile(stage, outt, t, lv, warp, lane, p.half_scores != 0);
            fence_proxy_async_smem();   // generic-proxy writes of outt -> visible to the bulk store
            LP_PF(2);
            tile_done_arrive(it);       // non-blocking; the release warp passes the tile on
        }
        if (tid == 0) LP_PF_OUT(2);
    }
}

cudaError_t launch_decode_tma(const DecodeParams& p, const DecodeMaps& maps, int num_sms, cudaStream_t stream) {
    static_assert(DT_SMEM <= 227 * 1024, "decode stages exceed shared memory");
    static bool configured[64] = {false};
    cudaError_t e = configure_smem_once(decode_tma_kernel, DT_SMEM, configured);
    if (e != cudaSuccess) return e;
    const int grid = p.n_tiles < num_sms ? p.n_tiles : num_sms;
    decode_tma_kernel<<<grid, DT_THREADS, DT_SMEM, stream>>>(p, maps);
    return cudaGetLastError();
}

}  // namespace lp
