// Detect.forward eval tail (yolov6/models/effidehead.py:247-301, use_dfl=False): raw per-level
// prediction-conv outputs (NCHW) -> head tensor out[B, A, 290].
//
// Fuses, in one pass over the data: generate_anchors(is_eval=True) (anchor_generator.py:11-31,
// computed from the anchor index, never materialised), 8x sigmoid (:251-258), the
// reshape/cat/permute NCHW -> anchor-major transpose (:260-280), dist2bbox 'xywh' (:283,
// general.py:29-40), dist2cor (:284, general.py:51-66), the stride multiply (:285-286), the
// constant objectness column (:290) and the final concat (:287-301).
// Algorithmic traffic: 1156 B read + 1160 B written per anchor (the reference moves the payload
// roughly three times each way through ~30 launches).
//
// This is the general-shape kernel: all 512 threads load, transpose and store in turn.  Inputs
// whose level planes have 16-byte aligned rows (h*w % 4 == 0) take decode_tma.cu instead (TMA box
// loads + warp specialisation, 0.93 of the measured HBM peak against 0.80 here).
//
// HBM-bound transpose, persistent CTAs (one per SM), software-pipelined over tiles of 32 anchor
// positions of one level of one image:
//   in    289 channel rows of 32 positions (128 B each) land channel-major in a shared-memory
//         stage through 16-byte cp.async copies (LDGSTS.128, no register staging) issued by all
//         threads; four stages, so the loads of tiles t+1..t+3 fly while tile t is transposed
//         (111 KB in flight per SM).  Shapes whose rows are not 16-byte aligned use 4-byte cp.async.
//   xpose transpose_tile (decode_tile.cuh), shared with decode_tma.cu;
//   out   the tile's rows are contiguous in the output: one TMA bulk store (UBLKCP) when the
//         destination is 16-byte aligned, coalesced 64-bit stores otherwise.
// Per-tile phase profile (clock64): ~1400 cycles issuing the loads, ~2000 transposing, ~370 issuing
// the store, serialised by the two CTA-wide barriers: 3900 cycles against a ~2900-cycle HBM budget.
#include "decode_tile.cuh"

namespace lp {

constexpr int DEC_STAGES = 4;
constexpr int DEC_SMEM = (DEC_STAGES * STAGE_FLOATS + 2 * OUT_FLOATS) * 4;

__global__ void __launch_bounds__(DEC_THREADS, 1) decode_kernel(const DecodeParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* stage0 = reinterpret_cast<float*>(smem);
    float* out0 = stage0 + DEC_STAGES * STAGE_FLOATS;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned slot[LT_SLOTS];
    make_slots(slot, tid);
    // prologue: DEC_STAGES-1 tiles in flight
    const int first = blockIdx.x, step = gridDim.x;
    TileWalker cur, ahead;
    cur.init(p, first);
    ahead = cur;
#pragma unroll
    for (int k = 0; k < DEC_STAGES - 1; ++k) {
        if (first + k * step < p.n_tiles) load_tile(p, ahead.info(p), stage0 + k * STAGE_FLOATS, tid, slot);
        cp_async_commit();
        ahead.advance(p, step);
    }
    int it = 0;
    for (int tile = first; tile < p.n_tiles; tile += step, ++it) {
        const float* stage = stage0 + (it % DEC_STAGES) * STAGE_FLOATS;
        float* outt = out0 + (it & 1) * OUT_FLOATS;
        // the stage of tile it-1 was released by the __syncthreads that ended the previous iteration
        if (tile + (DEC_STAGES - 1) * step < p.n_tiles)
            load_tile(p, ahead.info(p), stage0 + ((it + DEC_STAGES - 1) % DEC_STAGES) * STAGE_FLOATS, tid, slot);
        cp_async_commit();
        ahead.advance(p, step);
        const TileInfo t = cur.info(p);
        cur.advance(p, step);
        const DecodeLevel& lv = p.lv[t.l];
        cp_async_wait<DEC_STAGES - 1>();  // this thread's copies of the current tile have landed
        // the bulk store of tile it-2 must have finished READING this out buffer before it is rewritten
        if (tid == 0) bulk_wait_read1();
        __syncthreads();                  // ... and everybody else's copies

        transpose_tile(stage, outt, t, lv, warp, lane, p.half_scores != 0);
        fence_proxy_async_smem();  // generic-proxy writes of outt -> visible to the bulk store
        __syncthreads();           // also releases this stage for the copies queued next iteration

        // t.n finished rows are contiguous in the output
        float* dst = p.out + ((size_t)t.b * p.A + lv.anchor_off + t.p0) * ROW;
        const uint32_t bytes = (uint32_t)t.n * (ROW * 4);
        if (((reinterpret_cast<uintptr_t>(dst) | bytes) & 15u) == 0) {
            if (tid == 0) bulk_s2g(dst, outt, bytes);
        } else {  // 8-byte aligned only (odd row index or odd row count)
            float2* d2 = reinterpret_cast<float2*>(dst);
            const float2* s2 = reinterpret_cast<const float2*>(outt);
            for (int i = tid; i < t.n * (ROW / 2); i += DEC_THREADS) d2[i] = s2[i];
        }
        if (tid == 0) bulk_commit();  // one (possibly empty) group per tile keeps wait_group.read 1 exact
    }
    cp_async_wait<0>();
    if (tid == 0) bulk_wait0();
}

// the device sigmoid on a flat array (debug / property tests: monotonicity, accuracy)
__global__ void sigmoid_kernel(const float* __restrict__ in, long long n, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = sigmoid_f32(in[i]);
}
cudaError_t launch_sigmoid(const float* in, long long n, float* out, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    sigmoid_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(in, n, out);
    return cudaGetLastError();
}

cudaError_t launch_decode(const DecodeParams& p, const DecodeMaps* maps, int num_sms, cudaStream_t stream) {
    if (p.n_tiles <= 0) return cudaSuccess;
    if (p.bulk_in == 2 && maps != nullptr) return launch_decode_tma(p, *maps, num_sms, stream);
    static_assert(DEC_SMEM <= 227 * 1024, "decode stages exceed shared memory");
    static bool configured[64] = {false};
    cudaError_t e = configure_smem_once(decode_kernel, DEC_SMEM, configured);
    if (e != cudaSuccess) return e;
    const int grid = p.n_tiles < num_sms ? p.n_tiles : num_sms;
    decode_kernel<<<grid, DEC_THREADS, DEC_SMEM, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace lp
