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
// A CTA transposes a tile of kTile consecutive positions of one level through shared memory:
// channel-major coalesced reads (one 128-B line per channel per 32 positions), row-major
// coalesced 64-bit writes of the finished 1160-B rows.
#include "kernels.cuh"

namespace lp {

constexpr int DEC_THREADS = 256;
constexpr int DEC_WARPS = DEC_THREADS / 32;
constexpr int N_CLS = ROW - 13;   // 277 sigmoid columns

__device__ __forceinline__ float sigmoid_f32(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

__global__ void __launch_bounds__(DEC_THREADS) decode_kernel(const DecodeParams p) {
    __shared__ __align__(16) float tile[DEC_TILE * ROW];   // finished rows, packed like the output
    __shared__ float raw[12][DEC_TILE];                     // ltrb + 8 corner distances

    const int b = blockIdx.y;
    int l = 0;
#pragma unroll
    for (int i = 1; i < LP_MAX_LEVELS; ++i)
        if (i < p.n_levels && (int)blockIdx.x >= p.lv[i].tile_off) l = i;
    const DecodeLevel& lv = p.lv[l];
    const int p0 = ((int)blockIdx.x - lv.tile_off) * DEC_TILE;
    const int n = min(DEC_TILE, lv.hw - p0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool in = lane < n;

    // class channels: column 13+c of the row; source tensor = group_of(13+c)
    for (int c = warp; c < N_CLS; c += DEC_WARPS) {
        const int col = 13 + c;
        const int g = group_of(col);
        const int ch = col - group_begin(g);
        const int width = group_begin(g + 1) - group_begin(g);
        if (in) {
            const float x = __ldg(lv.cls[g] + ((size_t)b * width + ch) * lv.hw + p0 + lane);
            tile[lane * ROW + col] = sigmoid_f32(x);
        }
    }
    for (int c = warp; c < 12; c += DEC_WARPS) {
        if (in) {
            const float* src = c < 4 ? lv.reg + ((size_t)b * 4 + c) * lv.hw : lv.cor + ((size_t)b * 8 + (c - 4)) * lv.hw;
            raw[c][lane] = __ldg(src + p0 + lane);
        }
    }
    __syncthreads();
    if (threadIdx.x < n) {
        const int t = threadIdx.x;
        const int pos = p0 + t;
        const int y = pos / lv.w, x = pos - y * lv.w;
        const float ax = __fadd_rn((float)x, 0.5f), ay = __fadd_rn((float)y, 0.5f);  // anchor_generator.py:13-14
        const float s = lv.stride;
        float* r = tile + t * ROW;
        // dist2bbox 'xywh' (general.py:31-38) then *= stride (effidehead.py:285)
        const float x1 = __fsub_rn(ax, raw[0][t]), y1 = __fsub_rn(ay, raw[1][t]);
        const float x2 = __fadd_rn(ax, raw[2][t]), y2 = __fadd_rn(ay, raw[3][t]);
        r[0] = __fmul_rn(__fmul_rn(__fadd_rn(x1, x2), 0.5f), s);
        r[1] = __fmul_rn(__fmul_rn(__fadd_rn(y1, y2), 0.5f), s);
        r[2] = __fmul_rn(__fsub_rn(x2, x1), s);
        r[3] = __fmul_rn(__fsub_rn(y2, y1), s);
        r[4] = 1.0f;                                             // effidehead.py:290
        // dist2cor (general.py:51-66) then *= stride (effidehead.py:286)
        r[5] = __fmul_rn(__fsub_rn(ax, raw[4][t]), s);
        r[6] = __fmul_rn(__fsub_rn(ay, raw[5][t]), s);
        r[7] = __fmul_rn(__fsub_rn(ax, raw[6][t]), s);
        r[8] = __fmul_rn(__fadd_rn(ay, raw[7][t]), s);
        r[9] = __fmul_rn(__fadd_rn(ax, raw[8][t]), s);
        r[10] = __fmul_rn(__fadd_rn(ay, raw[9][t]), s);
        r[11] = __fmul_rn(__fadd_rn(ax, raw[10][t]), s);
        r[12] = __fmul_rn(__fsub_rn(ay, raw[11][t]), s);
    }
    __syncthreads();
    // n finished rows are contiguous in the output (8-byte aligned: 1160 = 8 * 145)
    float2* dst = reinterpret_cast<float2*>(p.out + ((size_t)b * p.A + lv.anchor_off + p0) * ROW);
    const float2* src = reinterpret_cast<const float2*>(tile);
    for (int i = threadIdx.x; i < n * (ROW / 2); i += DEC_THREADS) dst[i] = src[i];
}

cudaError_t launch_decode(const DecodeParams& p, int n_tiles, int B, cudaStream_t stream) {
    if (n_tiles <= 0 || B <= 0) return cudaSuccess;
    decode_kernel<<<dim3(n_tiles, B), DEC_THREADS, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace lp
