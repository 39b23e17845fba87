// Stand-alone geometry entry points that keep the reference's helper signatures alive
// (they are also called by the training losses, which stay untouched):
//   generate_anchors(is_eval=True, mode='af')   yolov6/assigners/anchor_generator.py:11-31
//   dist2bbox                                   yolov6/utils/general.py:29-40
//   dist2cor                                    yolov6/utils/general.py:51-66
//   xywh2xyxy                                   yolov6/utils/nms.py:21-28
//   Inferer.rescale (+ caller's .round())       yolov6/core/inferer.py:203-228, :100
// All are tiny element-wise kernels (launch-latency bound); the hot path never calls them --
// K-decode and K2 fuse the same arithmetic.
#include "kernels.cuh"

namespace lp {

__global__ void anchors_kernel(const AnchorLevels lv, float* __restrict__ points, float* __restrict__ strides) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= lv.A) return;
    int l = 0;
#pragma unroll
    for (int i = 1; i < LP_MAX_LEVELS; ++i)
        if (i < lv.n_levels && a >= lv.off[i]) l = i;
    const int pos = a - lv.off[l];
    const int y = pos / lv.w[l], x = pos - y * lv.w[l];
    points[2 * a] = __fadd_rn((float)x, lv.offset);
    points[2 * a + 1] = __fadd_rn((float)y, lv.offset);
    strides[a] = lv.stride[l];
}

__global__ void dist2bbox_kernel(const float4* __restrict__ dist, const float2* __restrict__ ap, long long total, int A,
                                 int xywh, float4* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float2 a = ap[i % A];
    const float4 d = dist[i];
    const float x1 = __fsub_rn(a.x, d.x), y1 = __fsub_rn(a.y, d.y), x2 = __fadd_rn(a.x, d.z), y2 = __fadd_rn(a.y, d.w);
    out[i] = xywh ? make_float4(__fmul_rn(__fadd_rn(x1, x2), 0.5f), __fmul_rn(__fadd_rn(y1, y2), 0.5f),
                                __fsub_rn(x2, x1), __fsub_rn(y2, y1))
                  : make_float4(x1, y1, x2, y2);
}

__global__ void dist2cor_kernel(const float4* __restrict__ dist, const float2* __restrict__ ap, long long total, int A,
                                float4* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float2 a = ap[i % A];
    const float4 d0 = dist[2 * i], d1 = dist[2 * i + 1];
    out[2 * i] = make_float4(__fsub_rn(a.x, d0.x), __fsub_rn(a.y, d0.y), __fsub_rn(a.x, d0.z), __fadd_rn(a.y, d0.w));
    out[2 * i + 1] = make_float4(__fadd_rn(a.x, d1.x), __fadd_rn(a.y, d1.y), __fadd_rn(a.x, d1.z), __fsub_rn(a.y, d1.w));
}

__global__ void xywh2xyxy_kernel(const float* __restrict__ in, long long n, long long in_stride, float* out, long long out_stride) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* r = in + i * in_stride;
    const float4 b = xywh_to_xyxy(r[0], r[1], r[2], r[3]);
    float* o = out + i * out_stride;
    o[0] = b.x; o[1] = b.y; o[2] = b.z; o[3] = b.w;
}

__global__ void rescale_kernel(float* rows, long long k, long long row_stride, float pad_x, float pad_y, float ratio,
                               float w0, float h0, int do_round) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k * 12) return;
    const long long r = i / 12;
    const int c = (int)(i - r * 12);
    float* v = rows + r * row_stride + c;
    *v = (c & 1) ? rescale_coord(*v, pad_y, ratio, h0, do_round) : rescale_coord(*v, pad_x, ratio, w0, do_round);
}

__global__ void rescale_batch_kernel(float* det, const int* __restrict__ counts, int max_det,
                                     const float* __restrict__ params, int do_round) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = min(counts[b], max_det);
    if (i >= n * 12) return;
    const int r = i / 12, c = i - r * 12;
    const float* q = params + (size_t)b * 5;
    float* v = det + ((size_t)b * max_det + r) * OUTW + c;
    *v = (c & 1) ? rescale_coord(*v, q[1], q[2], q[4], do_round) : rescale_coord(*v, q[0], q[2], q[3], do_round);
}

// --save-txt record of Inferer.infer (inferer.py:92-93,103-119), one thread per detection:
// [8 class ids | box_convert(xyxy) / (W0,H0,W0,H0) | 8 corners / (W0,H0)x4 | mean(row[12:19])]
__global__ void txt_records_kernel(const float* __restrict__ det, const int* __restrict__ counts, int max_det,
                                   const float* __restrict__ src_wh, float* __restrict__ rec) {
    const int b = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= min(counts[b], max_det)) return;
    const float* r = det + ((size_t)b * max_det + k) * OUTW;
    float* o = rec + ((size_t)b * max_det + k) * 21;
    const float w0 = src_wh[2 * b], h0 = src_wh[2 * b + 1];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = r[20 + i];                               // inferer.py:117
    // box_convert (inferer.py:309-316) then / gn (:115); true fp32 divisions
    o[8] = __fdiv_rn(__fmul_rn(__fadd_rn(r[0], r[2]), 0.5f), w0);
    o[9] = __fdiv_rn(__fmul_rn(__fadd_rn(r[1], r[3]), 0.5f), h0);
    o[10] = __fdiv_rn(__fsub_rn(r[2], r[0]), w0);
    o[11] = __fdiv_rn(__fsub_rn(r[3], r[1]), h0);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[12 + i] = __fdiv_rn(r[4 + i], (i & 1) ? h0 : w0);   // :116
    float s = r[12];
#pragma unroll
    for (int i = 13; i < 19; ++i) s = __fadd_rn(s, r[i]);                       // :113, seven of the eight groups
    o[20] = __fdiv_rn(s, 7.0f);
}

// Label prep of Evaler.predict (yolov6/core/evaler.py:119-127), one thread per target row:
// in[T,21] = image | 8 ids | normalised xywh | 8 normalised corners; out[T,20] = 8 ids | xyxy px | corners px.
// Rows keep their input order; out_image[T] carries the image index for the host-side grouping.
__global__ void prepare_targets_kernel(const float* __restrict__ in, int T, float w, float h, float* __restrict__ out,
                                       int* __restrict__ out_image) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const float* r = in + (size_t)t * 21;
    float* o = out + (size_t)t * 20;
    out_image[t] = (int)r[0];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = r[1 + i];
    const float4 b = xywh_to_xyxy(r[9], r[10], r[11], r[12]);                      // :120 (nms.py:21-28)
    o[8] = __fmul_rn(b.x, w); o[9] = __fmul_rn(b.y, h); o[10] = __fmul_rn(b.z, w); o[11] = __fmul_rn(b.w, h);   // :123-125
#pragma unroll
    for (int i = 0; i < 8; ++i) o[12 + i] = __fmul_rn(r[13 + i], (i & 1) ? h : w);
}

// Per-target matching of Evaler.eval (yolov6/core/evaler.py:183-229), one warp per target:
// IoU of the target box with every prediction of its image (box_iou, general.py:93-115), first
// maximum (torch.max, :190), then the corner test (:218) and the 8-character test (:223-226)
// against the matched prediction.  match[t] = { t_iou, match index, is_cor, is_cls };
// t_iou = -1 marks an image without predictions (the reference skips it, :188).
__global__ void eval_match_kernel(const float* __restrict__ det, const int* __restrict__ counts, int B, int max_det,
                                  const float* __restrict__ targets, const int* __restrict__ target_image, int T,
                                  float* __restrict__ match) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= T) return;
    const int b = target_image[t];
    float* o = match + (size_t)t * 4;
    // a target whose image index is outside the batch reads nothing (the host accumulator rejects it)
    const int n = (b >= 0 && b < B) ? min(counts[b], max_det) : 0;
    const float* tg = targets + (size_t)t * 20;
    const float* rows = det + (size_t)(n > 0 ? b : 0) * max_det * OUTW;
    if (n <= 0) {
        if (lane == 0) { o[0] = -1.0f; o[1] = 0.0f; o[2] = 0.0f; o[3] = 0.0f; }
        return;
    }
    const float tx1 = tg[8], ty1 = tg[9], tx2 = tg[10], ty2 = tg[11];
    const float tarea = __fmul_rn(__fsub_rn(tx2, tx1), __fsub_rn(ty2, ty1));
    float best = -INFINITY;
    int bi = 1 << 30;
    for (int i = lane; i < n; i += 32) {
        const float4 p = *reinterpret_cast<const float4*>(rows + (size_t)i * OUTW);
        const float parea = __fmul_rn(__fsub_rn(p.z, p.x), __fsub_rn(p.w, p.y));
        const float w = fmaxf(__fsub_rn(fminf(p.z, tx2), fmaxf(p.x, tx1)), 0.0f);
        const float h = fmaxf(__fsub_rn(fminf(p.w, ty2), fmaxf(p.y, ty1)), 0.0f);
        const float inter = __fmul_rn(w, h);
        const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(parea, tarea), inter));
        if (iou > best) { best = iou; bi = i; }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {  // max, ties -> lowest index (torch.max on CPU)
        const float ov = __shfl_xor_sync(0xffffffffu, best, s);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, s);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) {
        if (bi >= n) bi = 0;  // every IoU was NaN: torch.max returns index 0
        const float* pr = rows + (size_t)bi * OUTW;
        float s = fabsf(__fsub_rn(pr[4], tg[12]));
#pragma unroll
        for (int i = 1; i < 8; ++i) s = __fadd_rn(s, fabsf(__fsub_rn(pr[4 + i], tg[12 + i])));
        const bool is_cor = __fdiv_rn(s, 8.0f) < __fmul_rn(0.1f, __fsqrt_rn(tarea));
        bool is_cls = true;
#pragma unroll
        for (int i = 0; i < 8; ++i) is_cls = is_cls && ((int)pr[20 + i] == (int)tg[i]);
        o[0] = best;
        o[1] = (float)bi;
        o[2] = is_cor ? 1.0f : 0.0f;
        o[3] = is_cls ? 1.0f : 0.0f;
    }
}

static inline unsigned blocks_for(long long n, int t) { return (unsigned)((n + t - 1) / t); }

cudaError_t launch_anchors(const AnchorLevels& lv, float* points, float* strides, cudaStream_t s) {
    if (lv.A <= 0) return cudaSuccess;
    anchors_kernel<<<blocks_for(lv.A, 256), 256, 0, s>>>(lv, points, strides);
    return cudaGetLastError();
}
cudaError_t launch_dist2bbox(const float* d, const float* ap, long long n, int A, int xywh, float* out, cudaStream_t s) {
    const long long total = n * A;
    if (total <= 0) return cudaSuccess;
    dist2bbox_kernel<<<blocks_for(total, 256), 256, 0, s>>>((const float4*)d, (const float2*)ap, total, A, xywh, (float4*)out);
    return cudaGetLastError();
}
cudaError_t launch_dist2cor(const float* d, const float* ap, long long n, int A, float* out, cudaStream_t s) {
    const long long total = n * A;
    if (total <= 0) return cudaSuccess;
    dist2cor_kernel<<<blocks_for(total, 256), 256, 0, s>>>((const float4*)d, (const float2*)ap, total, A, (float4*)out);
    return cudaGetLastError();
}
cudaError_t launch_xywh2xyxy(const float* in, long long n, long long is, float* out, long long os, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    xywh2xyxy_kernel<<<blocks_for(n, 256), 256, 0, s>>>(in, n, is, out, os);
    return cudaGetLastError();
}
cudaError_t launch_rescale(float* rows, long long k, long long rs, float px, float py, float ratio, float w0, float h0,
                           int do_round, cudaStream_t s) {
    if (k <= 0) return cudaSuccess;
    rescale_kernel<<<blocks_for(k * 12, 256), 256, 0, s>>>(rows, k, rs, px, py, ratio, w0, h0, do_round);
    return cudaGetLastError();
}
cudaError_t launch_rescale_batch(float* det, const int* counts, int B, int max_det, const float* params, int do_round,
                                 cudaStream_t s) {
    if (B <= 0 || max_det <= 0) return cudaSuccess;
    rescale_batch_kernel<<<dim3(blocks_for((long long)max_det * 12, 256), B), 256, 0, s>>>(det, counts, max_det, params, do_round);
    return cudaGetLastError();
}

cudaError_t launch_prepare_targets(const float* in, int T, float w, float h, float* out, int* out_image, cudaStream_t s) {
    if (T <= 0) return cudaSuccess;
    prepare_targets_kernel<<<blocks_for(T, 128), 128, 0, s>>>(in, T, w, h, out, out_image);
    return cudaGetLastError();
}

cudaError_t launch_eval_match(const float* det, const int* counts, int B, int max_det, const float* targets,
                              const int* target_image, int T, float* match, cudaStream_t s) {
    if (T <= 0) return cudaSuccess;
    eval_match_kernel<<<blocks_for((long long)T * 32, 128), 128, 0, s>>>(det, counts, B, max_det, targets, target_image, T, match);
    return cudaGetLastError();
}

cudaError_t launch_txt_records(const float* det, const int* counts, int B, int max_det, const float* src_wh, float* rec,
                               cudaStream_t s) {
    if (B <= 0 || max_det <= 0) return cudaSuccess;
    txt_records_kernel<<<dim3(blocks_for(max_det, 128), B), 128, 0, s>>>(det, counts, max_det, src_wh, rec);
    return cudaGetLastError();
}

}  // namespace lp
