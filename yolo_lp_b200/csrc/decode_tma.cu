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
//   warps 0-15        consumers: wait full/oempty, transpose + sigmoid + box/corner decode, one named
//                     barrier (bar.sync 1, 512), then one thread releases the stage and hands the out
//                     tile over.
// Positions past the end of a level are zero-filled by the TMA unit (and never stored).
#include "decode_tile.cuh"

namespace lp {

constexpr int DT_STAGES = 3;
constexpr int DT_OUTS = 3;
constexpr int DT_THREADS = DEC_THREADS + 64;   // + load producer warp + store producer warp
constexpr int DT_SMEM = (DT_STAGES * STAGE_FLOATS + DT_OUTS * OUT_FLOATS) * 4 + (2 * DT_STAGES + 2 * DT_OUTS) * 8;

// first output column fed by source tensor k (0..7 class groups, 8 reg, 9 cor) and its channel count
__device__ __forceinline__ int tensor_first_col(int k) { return k == 8 ? 0 : k == 9 ? 5 : group_begin(k); }

__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(DEC_THREADS) : "memory"); }

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
        for (int it = 0; it < n_my; ++it, w.advance(p, step)) {
            const int s = it % DT_STAGES, use = it / DT_STAGES;
            if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
            const TileInfo t = w.info(p);
            float* stage = stage0 + s * STAGE_FLOATS;
            fence_proxy_async_smem();   // the stage was last read through the generic proxy
            mbar_expect_tx(&full[s], (ROW - 1) * DEC_TILE * 4);
#pragma unroll
            for (int k = 0; k < DEC_TENSORS; ++k)
                tma_load_3d(stage + tensor_first_col(k) * DEC_TILE, &maps.m[t.l][k], t.p0, 0, t.b, &full[s]);
        }
    } else if (warp == DEC_WARPS + 1) { // ---- store producer
        if (lane != 0) return;
        for (int it = 0; it < n_my; ++it, w.advance(p, step)) {
            const int o = it % DT_OUTS;
            const TileInfo t = w.info(p);
            float* dst = p.out + ((size_t)t.b * p.A + p.lv[t.l].anchor_off + t.p0) * ROW;
            const uint32_t bytes = (uint32_t)t.n * (ROW * 4);
            mbar_wait(&ofull[o], (it / DT_OUTS) & 1);
            // 8-byte-aligned-only tiles (odd row index or odd row count) were stored by the consumers
            if (((reinterpret_cast<uintptr_t>(dst) | bytes) & 15u) == 0) bulk_s2g(dst, out0 + o * OUT_FLOATS, bytes);
            bulk_commit();              // one (possibly empty) group per tile keeps wait_group.read 1 exact
            bulk_wait_read1();          // the store of tile it-1 has finished reading its buffer
            if (it > 0) mbar_arrive(&oempty[(it - 1) % DT_OUTS]);
        }
        bulk_wait0();
    } else {                            // ---- consumers
        for (int it = 0; it < n_my; ++it, w.advance(p, step)) {
            const int s = it % DT_STAGES, o = it % DT_OUTS;
            const float* stage = stage0 + s * STAGE_FLOATS;
            float* outt = out0 + o * OUT_FLOATS;
            const TileInfo t = w.info(p);
            const DecodeLevel& lv = p.lv[t.l];
            if (it >= DT_OUTS) mbar_wait(&oempty[o], (it / DT_OUTS - 1) & 1);
            mbar_wait(&full[s], (it / DT_STAGES) & 1);
            transpose_tile(stage, outt, t, lv, warp, lane);
            fence_proxy_async_smem();   // generic-proxy writes of outt -> visible to the bulk store
            consumer_barrier();
            float* dst = p.out + ((size_t)t.b * p.A + lv.anchor_off + t.p0) * ROW;
            const uint32_t bytes = (uint32_t)t.n * (ROW * 4);
            if (((reinterpret_cast<uintptr_t>(dst) | bytes) & 15u) != 0) {
                float2* d2 = reinterpret_cast<float2*>(dst);
                const float2* s2 = reinterpret_cast<const float2*>(outt);
                for (int i = tid; i < t.n * (ROW / 2); i += DEC_THREADS) d2[i] = s2[i];
                consumer_barrier();     // nobody still reads outt when it is handed over
            }
            if (tid == 0) {
                mbar_arrive(&empty[s]);
                mbar_arrive(&ofull[o]);
            }
        }
    }
}

cudaError_t launch_decode_tma(const DecodeParams& p, const DecodeMaps& maps, int num_sms, cudaStream_t stream) {
    static_assert(DT_SMEM <= 227 * 1024, "decode stages exceed shared memory");
    cudaError_t e = cudaFuncSetAttribute(decode_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM);
    if (e != cudaSuccess) return e;
    const int grid = p.n_tiles < num_sms ? p.n_tiles : num_sms;
    decode_tma_kernel<<<grid, DT_THREADS, DT_SMEM, stream>>>(p, maps);
    return cudaGetLastError();
}

}  // namespace lp
