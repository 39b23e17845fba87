// Kernel parameter blocks and host-side launchers shared by the translation units of liblpnms.
#pragma once
#include <cuda.h>   // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "common.cuh"

namespace lp {

// ---- K1 filter.cu ------------------------------------------------------------------------------
struct FilterParams {
    const float* pred;            // [total_rows, 290]
    unsigned total_rows;          // B*A
    unsigned A;
    unsigned n_tiles;
    float conf;                   // (float)conf_thres
    unsigned long long* keys;     // [B, key_stride]
    int* counts;                  // [B], zeroed before launch
    unsigned key_stride;
    unsigned* tile_counter;       // dynamic tile scheduler, zeroed before launch
};
cudaError_t launch_filter(const FilterParams& p, int num_sms, cudaStream_t stream);
// the same for a head tensor stored as IEEE half (p.pred then points at halves); filter_half.cu
cudaError_t launch_filter_half(const FilterParams& p, int num_sms, cudaStream_t stream);

// ---- decode.cu ---------------------------------------------------------------------------------
constexpr int DEC_TILE = 32;      // anchor positions per tile
struct DecodeLevel {
    const float* cls[8];
    const float* reg;
    const float* cor;
    int w, hw;
    int anchor_off;   // first anchor index of this level
    int tile_off;     // first tile index of this level
    float stride;
};
struct DecodeParams {
    DecodeLevel lv[LP_MAX_LEVELS];
    int n_levels;
    int A;
    int tiles_per_image;
    int n_tiles;      // B * tiles_per_image
    int bulk_in;      // 0: 4-byte cp.async (unaligned planes), 1: 16-byte cp.async, 2: 3-D TMA boxes (DecodeMaps)
    int half_scores;  // round every class score to the nearest IEEE half (the reference's model.half() head tensor)
    float* out;
    long long* timing;  // debug only (-DLP_DEC_PROFILE builds): per-CTA cycle sums, see decode_tma.cu
};
// One tensor map per level and source tensor (0..7 class groups, 8 reg, 9 cor): the [B, C, h*w] tensor
// seen as a 3-D array (h*w innermost), box = 32 positions x all C channels x 1 image.
constexpr int DEC_TENSORS = 10;
struct alignas(64) DecodeMaps {
    CUtensorMap m[LP_MAX_LEVELS][DEC_TENSORS];
};
cudaError_t launch_decode(const DecodeParams& p, const DecodeMaps* maps, int num_sms, cudaStream_t stream);
cudaError_t launch_sigmoid(const float* in, long long n, float* out, cudaStream_t stream);

// source plane of output column `col` (col != 4) for image b of level lv
__device__ __forceinline__ const float* column_src(const DecodeLevel& lv, int b, int col) {
    if (col < 4) return lv.reg + ((size_t)b * 4 + col) * lv.hw;
    if (col < 13) return lv.cor + ((size_t)b * 8 + (col - 5)) * lv.hw;
    const int g = group_of(col);
    const int width = group_begin(g + 1) - group_begin(g);
    return lv.cls[g] + ((size_t)b * width + (col - group_begin(g))) * lv.hw;
}

// ---- KF fused.cu -------------------------------------------------------------------------------
struct LevelsFilterParams {
    DecodeLevel lv[LP_MAX_LEVELS];
    int n_levels;
    int tiles_per_image;          // tiles of DEC_TILE positions
    int n_tiles;                  // B * tiles_per_image
    float conf;
    unsigned long long* keys;     // [B, key_stride]
    int* counts;                  // [B], zeroed before launch
    unsigned key_stride;
    int A;                        // anchors per image
    float* rec;                   // [B, A, 28] finished detection rows of the candidates, by slot
    unsigned* slot_of;            // [B, A] slot of a candidate anchor (only candidates are written)
    unsigned* tile_counter;       // dynamic tile scheduler, zeroed before launch
    int half_levels;              // the level tensors hold IEEE halves (TMA kernel only)
    long long* timing;            // debug only (-DLP_KF_PROFILE builds): per-warp cycle sums, see fused_tma.cu
};
cudaError_t launch_levels_filter(const LevelsFilterParams& p, const DecodeMaps* maps, int num_ctas, cudaStream_t stream);



// ---- K2 nms.cu ---------------------------------------------------------------------------------
struct NmsParams {
    const float* pred;          // [B, A, 290]
    unsigned A;
    unsigned long long* keys;   // [B, key_stride]
    unsigned key_stride;
    const int* counts;          // [B] candidates per image (from K1)
    float iou_floor;            // largest float <= iou_thres
    int max_det;
    int max_nms;
    float4* kept_box;           // [B, max_det] workspace
    int* kept_anchor_ws;        // [B, max_det] workspace
    float* out;                 // [B, max_det, 28]
    int* out_counts;            // [B]
    int* kept_anchor;           // [B, max_det] or null
    const float* rescale;       // [B, 5] or null
    int do_round;
    int sort_smem_keys;         // capacity of the shared-memory sort buffer (power of two)
    long long* timing;          // debug only: [B, 16] clock64 stamps per phase, or null
    // fused path only (from_levels != 0): KF left the finished rows of every candidate, pred is null
    int from_levels;
    int half_input;             // pred holds IEEE halves (580-byte rows), upcast exactly on load
    int rearm;                  // pipelined entries: zero counts[b] / the tile counter once read, so the
                                // next filter launch on this workspace needs no memset node
    const float* rec;           // [B, A, 28] by slot
    const unsigned* slot_of;    // [B, A]
};
cudaError_t launch_nms(const NmsParams& p, int B, cudaStream_t stream);
int nms_sort_smem_keys(unsigned A);

// ---- geometry.cu -------------------------------------------------------------------------------
struct AnchorLevels {
    int w[LP_MAX_LEVELS], hw[LP_MAX_LEVELS], off[LP_MAX_LEVELS];
    float stride[LP_MAX_LEVELS];
    int n_levels, A;
    float offset;
};
cudaError_t launch_anchors(const AnchorLevels& lv, float* points, float* strides, cudaStream_t s);
cudaError_t launch_dist2bbox(const float* d, const float* ap, long long n, int A, int xywh, float* out, cudaStream_t s);
cudaError_t launch_dist2cor(const float* d, const float* ap, long long n, int A, float* out, cudaStream_t s);
cudaError_t launch_xywh2xyxy(const float* in, long long n, long long is, float* out, long long os, cudaStream_t s);
cudaError_t launch_rescale(float* rows, long long k, long long rs, float px, float py, float ratio, float w0, float h0,
                           int do_round, cudaStream_t s);
cudaError_t launch_rescale_batch(float* det, const int* counts, int B, int max_det, const float* params, int do_round,
                                 cudaStream_t s);

cudaError_t launch_prepare_targets(const float* in, int T, float w, float h, float* out, int* out_image, cudaStream_t s);
cudaError_t launch_eval_match(const float* det, const int* counts, int B, int max_det, const float* targets,
                              const int* target_image, int T, float* match, cudaStream_t s);
cudaError_t launch_txt_records(const float* det, const int* counts, int B, int max_det, const float* src_wh, float* rec,
                               cudaStream_t s);

}  // namespace lp
