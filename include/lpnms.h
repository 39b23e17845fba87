/*
 * lpnms.h -- C ABI of the B200-native YOLO-LP post-processing library
 *            (liblpnms.so, hand-written sm_100a CUDA, no torch types).
 *
 * The reference (KyleHuang9/YOLO-LP) exposes this path as plain Python
 * callables, not as an FFI; the entry points below are what a ctypes binding
 * for each of those callables binds (see INTEGRATION.md for the stubs):
 *
 *   lp_nms_f32               <- yolov6/utils/nms.py:31-130  non_max_suppression
 *                               (+ torchvision.ops.nms, call site nms.py:121,
 *                                + xywh2xyxy nms.py:21-28)
 *   lp_detect_decode_f32     <- yolov6/models/effidehead.py:247-301
 *                               Detect.forward eval tail (after the convs);
 *                               lp_detect_decode_half_scores_f32: the same under model.half()
 *   lp_generate_anchors_f32  <- yolov6/assigners/anchor_generator.py:11-31
 *   lp_dist2bbox_f32         <- yolov6/utils/general.py:29-40
 *   lp_dist2cor_f32          <- yolov6/utils/general.py:51-66
 *   lp_xywh2xyxy_f32         <- yolov6/utils/nms.py:21-28
 *   lp_rescale_f32           <- yolov6/core/inferer.py:203-228 (+ .round() :100)
 *   lp_rescale_batch_f32     <- same, one launch for a whole [B,max_det,28] batch
 *   lp_detect_postprocess_f32 <- effidehead.py:247-301 + nms.py:31-130 fused (no head tensor);
 *                               lp_detect_pipelined_to_host_f32: one pipelined step of it + the D2H copy
 *                               of the detections (the post-model flow of Inferer.infer, inferer.py:82-120)
 *   lp_txt_records_f32 / lp_txt_lines_host <- yolov6/core/inferer.py:92-93,103-120 (--save-txt records)
 *   lp_eval_match_f32 / lp_eval_accumulate_host <- yolov6/core/evaler.py:153-283 (LP metric)
 *   lp_prepare_targets_f32   <- yolov6/core/evaler.py:119-127 (Evaler.predict label prep)
 *   lp_nms_f16, lp_detect_postprocess_f16 (+ stage / pipelined twins)
 *                            <- the same calls in the reference's --half mode
 *                               (yolov6/core/inferer.py:46-50, evaler.py:116)
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless the
 *     parameter name ends in _host;
 *   - the library never allocates, frees or synchronises; all work is queued on
 *     the cudaStream_t passed as `stream` (void* here so C callers need no CUDA
 *     headers); it keeps no global state and is re-entrant: tuning / debug knobs
 *     are a per-call `const lp_opts_t* opts` HOST pointer (NULL = defaults), the
 *     last parameter of every entry that launches K1 / KF / K2 / decode;
 *   - every function returns int: 0 = ok, <0 = LP_E_* argument error,
 *     >0 = cudaError_t of a failed launch; nothing throws, aborts or exits.
 */
#ifndef LPNMS_H_
#define LPNMS_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LP_VERSION 200            /* major*10000 + minor*100 + patch */
#define LP_ROW 290                /* floats per head row: 4 box | 1 obj | 8 corners | 31 | 24 | 6x37 */
#define LP_OUT 28                 /* floats per detection: 4 xyxy | 8 corners | 8 conf | 8 argmax */
#define LP_MAX_LEVELS 4
#define LP_MAX_NMS_DEFAULT 30000  /* yolov6/utils/nms.py:62 */

#define LP_OK 0
#define LP_E_NULL (-1)            /* required pointer is NULL */
#define LP_E_SIZE (-2)            /* negative / zero / overflowing size */
#define LP_E_ALIGN (-3)           /* pointer not aligned as documented */
#define LP_E_WORKSPACE (-4)       /* workspace smaller than lp_nms_workspace_bytes() */
#define LP_E_THRESHOLD (-5)       /* conf/iou threshold outside [0,1] (nms.py:57-58) */
#define LP_E_ARG (-6)             /* any other bad argument */

#if defined(__GNUC__)
#define LP_API __attribute__((visibility("default")))
#else
#define LP_API
#endif

typedef void* lp_stream_t;        /* cudaStream_t */

/* One FPN level of raw prediction-conv outputs, NCHW fp32 contiguous:
 * cls[0..7] = pro[B,31,h,w] alp[B,24,h,w] ad0..ad5[B,37,h,w]; reg[B,4,h,w]; cor[B,8,h,w]. */
typedef struct lp_level {
    const float* cls[8];
    const float* reg;
    const float* cor;
    int h;
    int w;
    float stride;
} lp_level_t;

/* Per-call knobs; NULL or all-zero = production behaviour.  Results never depend on them. */
typedef struct lp_opts {
    int filter_ctas;     /* CTA count of K1 (lp::filter_kernel) / KF (lp::levels_filter_*); 0 = heuristic */
    int no_tma;          /* != 0: the cp.async / register-resident kernels of the decode and fused paths
                            instead of their TMA variants (same results) */
    long long* timing;   /* debug: DEVICE buffer [B,16] of int64 that K2 fills with clock64() stamps of its
                            phases (the decode / fused kernels of -DLP_DEC_PROFILE / -DLP_KF_PROFILE builds
                            write per-role cycle sums to it); NULL = off */
} lp_opts_t;

LP_API int lp_version(void);
LP_API const char* lp_error_string(int code);

/* Bytes of device scratch lp_nms_f32 needs for (B images, A anchors, max_det). */
LP_API int lp_nms_workspace_bytes(int B, int A, int max_det, size_t* out_bytes);

/*
 * Confidence filter + greedy NMS over pred[B,A,290] (contiguous, 16-byte
 * aligned, NOT modified -- the reference's in-place `x[:,13:] *= x[:,4:5]`
 * side effect, nms.py:76, is not reproduced).
 *   conf_thres  compared in fp32 as (float)conf_thres       (nms.py:90-91)
 *   iou_thres   compared as (double)iou_f32 > iou_thres     (torchvision CPU kernel)
 *   max_det     rows kept per image                         (nms.py:122-123)
 *   max_nms     candidates entering NMS, 30000 in nms.py:62; ties at the cut are
 *               broken (score desc, anchor asc) where the reference is unstable
 * Outputs: out[B,max_det,28] (rows >= counts[b] untouched), counts[B],
 * kept_anchor[B,max_det] (anchor index of every kept row; may be NULL).
 * rescale (may be NULL): [B,5] fp32 = pad_x, pad_y, ratio, W0, H0 per image;
 * when given, columns 0..11 of every kept row are mapped back to source
 * coordinates exactly as lp_rescale_f32 does (round applied iff do_round).
 */
LP_API int lp_nms_f32(const float* pred, int B, int A, double conf_thres, double iou_thres,
               int max_det, int max_nms, void* workspace, size_t workspace_bytes,
               float* out, int* counts, int* kept_anchor,
               const float* rescale, int do_round, lp_stream_t stream,
        const lp_opts_t* opts);

/*
 * The two stages of lp_nms_f32, separately launchable (profiling, per-stage timing):
 *   lp_nms_filter_f32    K1: nms.py:76-97 + :120 -- scores every row of pred once, leaves the
 *                        surviving candidates' sort keys and per-image counts in `workspace`;
 *   lp_nms_suppress_f32  K2: sort + torchvision.ops.nms (nms.py:121) + keep[:max_det] (:122-123)
 *                        + output rows, reading what K1 left in the same `workspace`.
 * lp_nms_f32 == lp_nms_filter_f32 followed by lp_nms_suppress_f32 on the same stream.
 */
LP_API int lp_nms_filter_f32(const float* pred, int B, int A, double conf_thres, void* workspace,
                             size_t workspace_bytes, lp_stream_t stream,
        const lp_opts_t* opts);
LP_API int lp_nms_suppress_f32(const float* pred, int B, int A, double iou_thres, int max_det, int max_nms,
                               void* workspace, size_t workspace_bytes, float* out, int* counts, int* kept_anchor,
                               const float* rescale, int do_round, lp_stream_t stream,
        const lp_opts_t* opts);

/*
 * One pipelined step driven from (at least) two streams of the caller (native equivalent of
 * yolo_lp_b200.nms.NmsPipeline.submit): K1 on filter_stream, K2 on nms_stream, ordered by
 * filtered_event; workspace_free_event (may be NULL) is the done_event of the step that last used
 * this workspace; done_event / time_*_event (may be NULL) are recorded after K2 / round K1.
 * All events are cudaEvent_t handles owned by the caller.
 * filter_stream may differ from step to step: alternating between two filter streams (one per
 * workspace) removes the stream order between consecutive K1 launches, so the first CTAs of step
 * i+1 move onto SMs as K2 of step i-1 / K1 of step i let go of them and K1's drain and ramp-up
 * overlap (cfg2: 48.2 -> 44.5 us per step); yolo_lp_b200.nms.NmsPipeline does this.
 * K2 of a pipelined step zeroes the workspace's candidate counters once it has read them; a
 * non-NULL workspace_free_event therefore also asserts that the LAST operation on this workspace
 * was such a step WITH THE SAME B, and lets the entry skip the memset node in front of K1.  Pass
 * NULL whenever the workspace is fresh, was last touched by any other entry point, is used with a
 * different B, or the previous step on it returned an error.  Every argument of both stages is
 * validated before anything is queued, so a failing call leaves the workspace as it was.
 */
LP_API int lp_nms_pipelined_f32(const float* pred, int B, int A, double conf_thres, double iou_thres, int max_det,
                                int max_nms, void* workspace, size_t workspace_bytes, float* out, int* counts,
                                int* kept_anchor, const float* rescale, int do_round, lp_stream_t filter_stream,
                                lp_stream_t nms_stream, void* workspace_free_event, void* filtered_event,
                                void* done_event, void* time_begin_event, void* time_end_event,
        const lp_opts_t* opts);

/*
 * fp16 head tensors (the reference's --half mode: inferer.py:46-50, evaler.py:116; SURVEY §8-f rank
 * 3).  `pred` holds B*A*290 IEEE halves, contiguous, 16-byte aligned; every value is upcast exactly
 * on load and all arithmetic is the f32 entries', so the results are bit for bit those of the f32
 * entry on the upcast tensor -- at half the bytes of the HBM- (and PCIe-) bound stage.  (The
 * reference's own half mode computes in half precision on CUDA behind an unstable sort and is not
 * reproducible across implementations; see DESIGN.md.)  Outputs, workspace and knobs as above.
 */
LP_API int lp_nms_f16(const void* pred, int B, int A, double conf_thres, double iou_thres, int max_det, int max_nms,
                      void* workspace, size_t workspace_bytes, float* out, int* counts, int* kept_anchor,
                      const float* rescale, int do_round, lp_stream_t stream,
        const lp_opts_t* opts);
LP_API int lp_nms_filter_f16(const void* pred, int B, int A, double conf_thres, void* workspace,
                             size_t workspace_bytes, lp_stream_t stream,
        const lp_opts_t* opts);
LP_API int lp_nms_suppress_f16(const void* pred, int B, int A, double iou_thres, int max_det, int max_nms,
                               void* workspace, size_t workspace_bytes, float* out, int* counts, int* kept_anchor,
                               const float* rescale, int do_round, lp_stream_t stream,
        const lp_opts_t* opts);
LP_API int lp_nms_pipelined_f16(const void* pred, int B, int A, double conf_thres, double iou_thres, int max_det,
                                int max_nms, void* workspace, size_t workspace_bytes, float* out, int* counts,
                                int* kept_anchor, const float* rescale, int do_round, lp_stream_t filter_stream,
                                lp_stream_t nms_stream, void* workspace_free_event, void* filtered_event,
                                void* done_event, void* time_begin_event, void* time_end_event,
        const lp_opts_t* opts);

/* Debug / property tests: the decode kernel's sigmoid evaluated on a flat device array. */
LP_API int lp_debug_sigmoid_f32(const float* in, long long n, float* out, lp_stream_t stream);

/* Detect.forward eval tail: raw per-level conv outputs -> out[B,A,290],
 * A = sum h*w, levels in order; anchors are computed from the index, never
 * materialised. */
LP_API int lp_detect_decode_f32(const lp_level_t* levels_host, int n_levels, int B, float* out, lp_stream_t stream,
        const lp_opts_t* opts);

/*
 * The same for a model.half() forward (inferer.py:46-50): the caller upcasts the half conv outputs (exact)
 * and every class score is rounded to the nearest IEEE half before it is stored as fp32 -- which is what the
 * reference's head tensor holds in that mode: torch.sigmoid on half tensors rounds once, and torch.cat
 * promotes the result to fp32 together with the box / corner columns, which the reference computes in
 * fp32 even then because its anchor points are fp32 (effidehead.py:251-258, 283-301).
 */
LP_API int lp_detect_decode_half_scores_f32(const lp_level_t* levels_host, int n_levels, int B, float* out,
                                            lp_stream_t stream, const lp_opts_t* opts);

/*
 * Fused head tail + NMS: raw per-level conv outputs -> detections, without materialising the
 * [B,A,290] head tensor (effidehead.py:247-301 followed by nms.py:31-130 in one pass over the class
 * planes).  Same outputs and knobs as lp_nms_f32; workspace from lp_detect_workspace_bytes
 * (A = sum h*w; it additionally holds one finished 28-float row per candidate); results are bit-identical to lp_detect_decode_f32 followed by lp_nms_f32 on the same
 * level tensors.  lp_detect_filter_f32 / lp_detect_suppress_f32 are its two stages (KF, K2).
 */
LP_API int lp_detect_workspace_bytes(int B, int A, int max_det, size_t* out_bytes);
LP_API int lp_detect_postprocess_f32(const lp_level_t* levels_host, int n_levels, int B, double conf_thres,
                                     double iou_thres, int max_det, int max_nms, void* workspace,
                                     size_t workspace_bytes, float* out, int* counts, int* kept_anchor,
                                     const float* rescale, int do_round, lp_stream_t stream,
        const lp_opts_t* opts);
LP_API int lp_detect_filter_f32(const lp_level_t* levels_host, int n_levels, int B, double conf_thres, int max_det,
                                void* workspace, size_t workspace_bytes, lp_stream_t stream,
        const lp_opts_t* opts);
LP_API int lp_detect_suppress_f32(const lp_level_t* levels_host, int n_levels, int B, double iou_thres, int max_det,
                                  int max_nms, void* workspace, size_t workspace_bytes, float* out, int* counts,
                                  int* kept_anchor, const float* rescale, int do_round, lp_stream_t stream,
        const lp_opts_t* opts);

/* lp_nms_pipelined_f32 for the fused path: KF on filter_stream, K2 on nms_stream. */
LP_API int lp_detect_pipelined_f32(const lp_level_t* levels_host, int n_levels, int B, double conf_thres,
                                   double iou_thres, int max_det, int max_nms, void* workspace,
                                   size_t workspace_bytes, float* out, int* counts, int* kept_anchor,
                                   const float* rescale, int do_round, lp_stream_t filter_stream,
                                   lp_stream_t nms_stream, void* workspace_free_event, void* filtered_event,
                                   void* done_event, void* time_begin_event, void* time_end_event,
        const lp_opts_t* opts);

/*
 * lp_detect_pipelined_f32 followed by the copy of the step's results to pinned HOST memory, in one call --
 * the production flow when the head runs on this GPU: level tensors resident in HBM, detections wanted on
 * the host (Inferer.infer's per-image loop, inferer.py:100-120).  After K2, copy_stream waits for done_event
 * (required), copies counts[B] -> counts_host_pinned and out[B,max_det,28] -> out_host_pinned with
 * cudaMemcpyAsync, and records copied_event (cudaEvent_t, required): wait for it, then read the rows of
 * image b as out_host_pinned[b, :counts_host_pinned[b]].  No kept_anchor / timing events here.
 */
LP_API int lp_detect_pipelined_to_host_f32(const lp_level_t* levels_host, int n_levels, int B, double conf_thres,
                                           double iou_thres, int max_det, int max_nms, void* workspace,
                                           size_t workspace_bytes, float* out, int* counts, const float* rescale,
                                           int do_round, lp_stream_t filter_stream, lp_stream_t nms_stream,
                                           void* workspace_free_event, void* filtered_event, void* done_event,
                                           float* out_host_pinned, int* counts_host_pinned, lp_stream_t copy_stream,
                                           void* copied_event, const lp_opts_t* opts);
LP_API int lp_detect_pipelined_to_host_f16(const lp_level_t* levels_host, int n_levels, int B, double conf_thres,
                                           double iou_thres, int max_det, int max_nms, void* workspace,
                                           size_t workspace_bytes, float* out, int* counts, const float* rescale,
                                           int do_round, lp_stream_t filter_stream, lp_stream_t nms_stream,
                                           void* workspace_free_event, void* filtered_event, void* done_event,
                                           float* out_host_pinned, int* counts_host_pinned, lp_stream_t copy_stream,
                                           void* copied_event, const lp_opts_t* opts);

/*
 * The fused path on fp16 level tensors (model.half(): the prediction convs emit halves).  The
 * pointers in lp_level_t then address IEEE halves; every value is upcast exactly on load, so the
 * results are bit for bit those of the f32 entries on the upcast tensors.  TMA kernel only: returns
 * LP_E_ARG unless every level has h*w % 8 == 0 and 16-byte aligned tensors -- upcast such inputs and
 * call the f32 entry.  K2 is shared: use lp_detect_suppress_f32 after lp_detect_filter_f16.
 * (Scores are NOT rounded to half here: these entries equal lp_detect_decode_f32 + lp_nms_f32 on the
 * upcast tensors, not lp_detect_decode_half_scores_f32 + lp_nms_f32 -- the fused call has no counterpart
 * in the reference, so there is no half-mode head tensor of the reference's to match.)
 */
LP_API int lp_detect_postprocess_f16(const lp_level_t* levels_host, int n_levels, int B, double conf_thres,
                                     double iou_thres, int max_det, int max_nms, void* workspace,
                                     size_t workspace_bytes, float* out, int* counts, int* kept_anchor,
                                     const float* rescale, int do_round, lp_stream_t stream,
        const lp_opts_t* opts);
LP_API int lp_detect_filter_f16(const lp_level_t* levels_host, int n_levels, int B, double conf_thres, int max_det,
                                void* workspace, size_t workspace_bytes, lp_stream_t stream,
        const lp_opts_t* opts);
LP_API int lp_detect_pipelined_f16(const lp_level_t* levels_host, int n_levels, int B, double conf_thres,
                                   double iou_thres, int max_det, int max_nms, void* workspace,
                                   size_t workspace_bytes, float* out, int* counts, int* kept_anchor,
                                   const float* rescale, int do_round, lp_stream_t filter_stream,
                                   lp_stream_t nms_stream, void* workspace_free_event, void* filtered_event,
                                   void* done_event, void* time_begin_event, void* time_end_event,
        const lp_opts_t* opts);

/* generate_anchors(is_eval=True, mode='af'): anchor_points[A,2], stride_tensor[A]. */
LP_API int lp_generate_anchors_f32(const int* h_host, const int* w_host, const float* stride_host, int n_levels,
                            float grid_cell_offset, float* anchor_points, float* stride_tensor, lp_stream_t stream);

/* dist2bbox: distance[n,A,4] (ltrb) + anchor_points[A,2] -> out[n,A,4]; xywh != 0 selects 'xywh'. */
LP_API int lp_dist2bbox_f32(const float* distance, const float* anchor_points, long long n, int A, int xywh,
                     float* out, lp_stream_t stream);

/* dist2cor: distance[n,A,8] + anchor_points[A,2] -> out[n,A,8] (TL, BL, BR, TR). */
LP_API int lp_dist2cor_f32(const float* distance, const float* anchor_points, long long n, int A,
                    float* out, lp_stream_t stream);

/* xywh2xyxy on n rows; in/out row strides in floats (>= 4). in == out allowed. */
LP_API int lp_xywh2xyxy_f32(const float* in, long long n, long long in_stride, float* out, long long out_stride,
                     lp_stream_t stream);

/* Inferer.rescale on k rows of 12 coords (row stride in floats), in place:
 * v = (v - pad) / ratio (true fp32 division), clamp x to [0,W0], y to [0,H0],
 * then round-half-even iff do_round.  pad/ratio are the reference's Python
 * doubles rounded to fp32 by the caller. */
LP_API int lp_rescale_f32(float* rows, long long k, long long row_stride, float pad_x, float pad_y, float ratio,
                   float w0, float h0, int do_round, lp_stream_t stream);

/* Same over det[B,max_det,28] with counts[B] and params[B,5] = pad_x,pad_y,ratio,W0,H0. */
LP_API int lp_rescale_batch_f32(float* det, const int* counts, int B, int max_det, const float* params,
                         int do_round, lp_stream_t stream);

/*
 * Caller-side records of Inferer.infer (yolov6/core/inferer.py:92-93,103-120), batched:
 *   lp_txt_records_f32  det[B,max_det,28] (already rescaled + rounded) + counts[B] + src_wh[B,2]
 *                       (W0,H0 of the source images) -> records[B,max_det,21] =
 *                       8 class ids | box_convert(xyxy)/(W0,H0,W0,H0) | 8 corners/(W0,H0)x4 | conf,
 *                       conf = mean(row[12:19]) (seven of the eight groups, as the reference does);
 *   lp_txt_lines_host   HOST helper: the --save-txt text of n records ('%g ' * 20, newline-ended),
 *                       written to buf; LP_E_WORKSPACE if buf is too small.
 */
LP_API int lp_txt_records_f32(const float* det, const int* counts, int B, int max_det, const float* src_wh,
                              float* records, lp_stream_t stream);
LP_API int lp_txt_lines_host(const float* records_host, long long n, char* buf, size_t buf_bytes, size_t* written);

/*
 * LP evaluation metric, Evaler.eval (yolov6/core/evaler.py:153-283):
 *   lp_eval_match_f32        per target (one warp each): box_iou against every prediction of its
 *                            image (general.py:93-115), first maximum, corner test (:218), 8-character
 *                            test (:223-226).  targets[T,20] = 8 class ids | xyxy | 8 corners, grouped
 *                            by image in ascending image order; match[T,4] = t_iou, match index,
 *                            is_cor, is_cls (t_iou = -1: image without predictions).
 *   lp_eval_accumulate_host  HOST: the reference's counters (counters[42] = true_cnt, pred_cnt,
 *                            pred_cnts[10], cor_right[10], cls_right[10], right[10]) and summary[25] =
 *                            mAP, mAP@.5, mAP@.75, mAP@.5:.95, recall, mAP_list[10], recall_list[10],
 *                            including the reference's stale-bin quirk for an IoU of exactly 1.0.
 */
/* Label prep of Evaler.predict (evaler.py:119-127): targets[T,21] = image | 8 ids | normalised xywh | 8
 * normalised corners -> out[T,20] = 8 ids | xyxy px | 8 corners px (same row order), out_image[T]. */
LP_API int lp_prepare_targets_f32(const float* targets, int T, float w, float h, float* out, int* out_image,
                                  lp_stream_t stream);
LP_API int lp_eval_match_f32(const float* det, const int* counts, int B, int max_det, const float* targets,
                             const int* target_image, int T, float* match, lp_stream_t stream);
LP_API int lp_eval_accumulate_host(const float* match_host, const int* target_image_host, const int* counts_host, int B,
                                   int T, long long* counters, double* summary);

#ifdef __cplusplus
}
#endif
#endif /* LPNMS_H_ */
