// extern "C" boundary of liblpnms.so (see include/lpnms.h): argument validation, workspace
// carving and kernel launches.  No allocation, no synchronisation; the only process-global state
// is idempotent per-device caches (SM count, kernel attributes, the tensor-map encoder).  Tuning and
// debug knobs travel with the call (lp_opts_t), so concurrent callers never see each other's.
#include <math.h>
#include <stdio.h>

#include "kernels.cuh"

namespace lp {

constexpr size_t WS_ALIGN = 256;

static inline size_t align_up(size_t x) { return (x + WS_ALIGN - 1) / WS_ALIGN * WS_ALIGN; }
static inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

static unsigned pow2_at_least(unsigned n) {
    unsigned p = 1;
    while (p < n) p <<= 1;
    return p;
}

// Per-image key stride: A when the candidates always fit the shared-memory sort, else the next
// power of two (the global-memory bitonic path pads in place).
static unsigned key_stride_for(unsigned A) {
    return pow2_at_least(A) <= (unsigned)nms_sort_smem_keys(A) ? A : pow2_at_least(A);
}

struct WsLayout {
    size_t counts, keys, kept_box, kept_anchor, total;
    size_t slot_of, rec, total_fused;  // fused path only, after `total`
};
static WsLayout ws_layout(int B, int A, int max_det) {
    WsLayout w;
    size_t off = 0;
    w.counts = off;      off = align_up(off + sizeof(int) * ((size_t)B + 1));  // + the tile counter of K1
    w.keys = off;        off = align_up(off + sizeof(unsigned long long) * (size_t)B * key_stride_for(A));
    w.kept_box = off;    off = align_up(off + sizeof(float4) * (size_t)B * (size_t)max_det);
    w.kept_anchor = off; off = align_up(off + sizeof(int) * (size_t)B * (size_t)max_det);
    w.total = off;
    w.slot_of = off;     off = align_up(off + sizeof(unsigned) * (size_t)B * (size_t)A);
    w.rec = off;         off = align_up(off + sizeof(float) * LP_OUT * (size_t)B * (size_t)A);
    w.total_fused = off;
    return w;
}

static int num_sms_cached() {
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cache[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev] = n;
    }
    return cache[dev];
}

static bool size_ok(int B, int A, int max_det) {
    if (B <= 0 || A <= 0 || max_det < 0) return false;
    if ((long long)B * A >= (1ll << 31) - 64) return false;         // row index fits 32 bits
    if ((long long)B * (long long)max_det >= (1ll << 31)) return false;
    return true;
}

}  // namespace lp

using namespace lp;

extern "C" {

LP_API int lp_version(void) { return LP_VERSION; }

LP_API const char* lp_error_string(int code) {
    switch (code) {
        case LP_OK: return "ok";
        case LP_E_NULL: return "required pointer is NULL";
        case LP_E_SIZE: return "bad size";
        case LP_E_ALIGN: return "pointer is not aligned as required";
        case LP_E_WORKSPACE: return "workspace too small";
        case LP_E_THRESHOLD: return "threshold outside [0, 1]";
        case LP_E_ARG: return "bad argument";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

LP_API int lp_nms_workspace_bytes(int B, int A, int max_det, size_t* out_bytes) {
    if (!out_bytes) return LP_E_NULL;
    if (!size_ok(B, A, max_det)) return LP_E_SIZE;
    *out_bytes = ws_layout(B, A, max_det).total;
    return LP_OK;
}

// Per-call knobs (lp_opts_t, NULL = production defaults), resolved once per entry.
struct Opts {
    int ctas = 0;                 // CTA count of K1 / KF; 0 = the heuristics below
    bool tma = true;              // TMA variants of the decode / fused kernels when the shapes allow
    long long* timing = nullptr;  // debug: device buffer for clock64 stamps
};
static Opts resolve(const lp_opts_t* o) {
    Opts r;
    if (o) {
        r.ctas = o->filter_ctas > 0 ? o->filter_ctas : 0;
        r.tma = o->no_tma == 0;
        r.timing = o->timing;
    }
    return r;
}

// shared validation + parameter setup of the two NMS stages
static int nms_setup(const float* pred, int B, int A, int max_det, void* workspace, size_t workspace_bytes,
                     FilterParams& f, NmsParams& n, const Opts& o, bool need_pred = true) {
    if ((need_pred && !pred) || !workspace) return LP_E_NULL;
    if (!size_ok(B, A, max_det)) return LP_E_SIZE;
    if ((need_pred && !aligned(pred, 16)) || !aligned(workspace, WS_ALIGN)) return LP_E_ALIGN;
    const WsLayout w = ws_layout(B, A, max_det);
    if (workspace_bytes < w.total) return LP_E_WORKSPACE;
    char* ws = static_cast<char*>(workspace);
    f.pred = pred;
    f.total_rows = (unsigned)B * (unsigned)A;
    f.A = (unsigned)A;
    f.n_tiles = (f.total_rows + 31u) / 32u;
    f.conf = 0.0f;
    f.keys = reinterpret_cast<unsigned long long*>(ws + w.keys);
    f.counts = reinterpret_cast<int*>(ws + w.counts);
    f.key_stride = key_stride_for((unsigned)A);
    f.tile_counter = reinterpret_cast<unsigned*>(f.counts + B);
    n.pred = pred;
    n.A = (unsigned)A;
    n.keys = f.keys;
    n.key_stride = f.key_stride;
    n.counts = f.counts;
    n.iou_floor = 0.0f;
    n.max_det = max_det;
    n.max_nms = LP_MAX_NMS_DEFAULT;
    n.kept_box = reinterpret_cast<float4*>(ws + w.kept_box);
    n.kept_anchor_ws = reinterpret_cast<int*>(ws + w.kept_anchor);
    n.out = nullptr;
    n.out_counts = nullptr;
    n.kept_anchor = nullptr;
    n.rescale = nullptr;
    n.do_round = 0;
    n.sort_smem_keys = nms_sort_smem_keys((unsigned)A);
    n.timing = o.timing;
    n.from_levels = 0;
    n.half_input = 0;
    n.rearm = 0;
    n.rec = reinterpret_cast<const float*>(ws + w.rec);           // only valid in a fused-size workspace
    n.slot_of = reinterpret_cast<const unsigned*>(ws + w.slot_of);
    return LP_OK;
}

// Everything lp_nms_* can reject, checked before anything is queued (a step that failed half way
// would leave the workspace's candidate counters non-zero behind a caller who believes it re-armed).
static int nms_validate(const void* pred, int B, int A, double conf_thres, double iou_thres, int max_det, int max_nms,
                        const void* workspace, size_t workspace_bytes, const float* out, const int* counts) {
    if (!pred || !workspace || !counts || (!out && max_det > 0)) return LP_E_NULL;
    if (!size_ok(B, A, max_det) || max_nms <= 0) return LP_E_SIZE;
    if (!(conf_thres >= 0.0 && conf_thres <= 1.0) || !(iou_thres >= 0.0 && iou_thres <= 1.0)) return LP_E_THRESHOLD;
    if (!aligned(pred, 16) || !aligned(workspace, WS_ALIGN) || !aligned(out, 4) || !aligned(counts, 4)) return LP_E_ALIGN;
    if (workspace_bytes < ws_layout(B, A, max_det).total) return LP_E_WORKSPACE;
    return LP_OK;
}

// K1's grid: it saturates HBM with roughly half the SMs (one 217 KB CTA each); the rest is left free
// so that K2 of the previous batch (one CTA per image, driven from a second stream) can run
// concurrently instead of queueing behind K1's persistent CTAs.  Another ninth of the SMs (16 of 148)
// is left free beyond that: with the filter alternating between two streams the next batch's first
// CTAs start there at once, and K1 alone is no slower (measured at cfg2 / cfg5: 100 CTAs 43.2 /
// 169.9 us per pipelined step, 116 CTAs 44.4 / 171.9).
static int filter_ctas(int B, const Opts& o) {
    const int sms = num_sms_cached();
    if (o.ctas > 0) return o.ctas < sms ? o.ctas : sms;
    int ctas = sms - (B < sms / 2 ? B : sms / 2) - sms / 9;
    if (ctas < sms - sms * 7 / 16) ctas = sms - sms * 7 / 16;   // 84 of 148 still saturate HBM
    return ctas;
}

// armed: the last kernel that ran on this workspace was a K2 launched with rearm (it zeroed the
// candidate counts and the tile counter), so the memset node in front of the filter kernel is skipped
static int nms_filter(const float* pred, int B, int A, double conf_thres, void* workspace, size_t workspace_bytes,
                      lp_stream_t stream, const Opts& o, bool armed, bool half) {
    if (!(conf_thres >= 0.0 && conf_thres <= 1.0)) return LP_E_THRESHOLD;
    FilterParams f;
    NmsParams n;
    const int rc = nms_setup(pred, B, A, 0, workspace, (size_t)-1, f, n, o);
    if (rc != LP_OK) return rc;
    // the layout up to the keys does not depend on max_det; require at least counts + keys
    const WsLayout w = ws_layout(B, A, 0);
    if (workspace_bytes < w.kept_box) return LP_E_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!armed) {
        const cudaError_t e = cudaMemsetAsync(f.counts, 0, sizeof(int) * ((size_t)B + 1), s);  // counts + tile counter
        if (e != cudaSuccess) return (int)e;
    }
    f.conf = (float)conf_thres;  // tensor >= python-scalar compares in fp32
    const int ctas = filter_ctas(B, o);
    return (int)(half ? launch_filter_half(f, ctas, s) : launch_filter(f, ctas, s));
}

static int nms_suppress(const float* pred, int B, int A, double iou_thres, int max_det, int max_nms, void* workspace,
                        size_t workspace_bytes, float* out, int* counts, int* kept_anchor, const float* rescale,
                        int do_round, lp_stream_t stream, const Opts& o, bool rearm, bool half) {
    if (!counts || (!out && max_det > 0)) return LP_E_NULL;
    if (max_nms <= 0) return LP_E_SIZE;
    if (!(iou_thres >= 0.0 && iou_thres <= 1.0)) return LP_E_THRESHOLD;
    if (!aligned(out, 4) || !aligned(counts, 4)) return LP_E_ALIGN;
    FilterParams f;
    NmsParams n;
    const int rc = nms_setup(pred, B, A, max_det, workspace, workspace_bytes, f, n, o);
    if (rc != LP_OK) return rc;
    // (double)ovr > iou_thres  <=>  ovr > largest float <= iou_thres
    float iou_floor = (float)iou_thres;
    if ((double)iou_floor > iou_thres) iou_floor = nextafterf(iou_floor, -INFINITY);
    n.iou_floor = iou_floor;
    n.max_nms = max_nms;
    n.out = out;
    n.out_counts = counts;
    n.kept_anchor = kept_anchor;
    n.rescale = rescale;
    n.do_round = do_round;
    n.rearm = rearm ? 1 : 0;
    n.half_input = half ? 1 : 0;
    return (int)launch_nms(n, B, static_cast<cudaStream_t>(stream));
}

// The event choreography of one pipelined step, shared by the NMS and the fused entries: `filter`
// queues K1 / KF on the filter stream, `suppress` K2 on the NMS stream.
extern "C++" {
template <class Filter, class Suppress>
static int pipelined_step(lp_stream_t filter_stream, lp_stream_t nms_stream, void* workspace_free_event,
                          void* filtered_event, void* done_event, void* time_begin_event, void* time_end_event,
                          Filter filter, Suppress suppress) {
    if (!filtered_event) return LP_E_NULL;
    cudaStream_t sf = static_cast<cudaStream_t>(filter_stream), sn = static_cast<cudaStream_t>(nms_stream);
    cudaError_t e;
    if (workspace_free_event) {
        e = cudaStreamWaitEvent(sf, static_cast<cudaEvent_t>(workspace_free_event), 0);
        if (e != cudaSuccess) return (int)e;
    }
    if (time_begin_event) {
        e = cudaEventRecord(static_cast<cudaEvent_t>(time_begin_event), sf);
        if (e != cudaSuccess) return (int)e;
    }
    // a workspace that comes with the done_event of its previous step was re-armed by that step's K2
    int rc = filter(workspace_free_event != nullptr);
    if (rc != LP_OK) return rc;
    if (time_end_event) {
        e = cudaEventRecord(static_cast<cudaEvent_t>(time_end_event), sf);
        if (e != cudaSuccess) return (int)e;
    }
    e = cudaEventRecord(static_cast<cudaEvent_t>(filtered_event), sf);
    if (e != cudaSuccess) return (int)e;
    e = cudaStreamWaitEvent(sn, static_cast<cudaEvent_t>(filtered_event), 0);
    if (e != cudaSuccess) return (int)e;
    rc = suppress();
    if (rc != LP_OK) return rc;
    if (done_event) {
        e = cudaEventRecord(static_cast<cudaEvent_t>(done_event), sn);
        if (e != cudaSuccess) return (int)e;
    }
    return LP_OK;
}
}  // extern "C++"

static int nms_pipelined(const float* pred, int B, int A, double conf_thres, double iou_thres, int max_det, int max_nms,
                         void* workspace, size_t workspace_bytes, float* out, int* counts, int* kept_anchor,
                         const float* rescale, int do_round, lp_stream_t filter_stream, lp_stream_t nms_stream,
                         void* workspace_free_event, void* filtered_event, void* done_event, void* time_begin_event,
                         void* time_end_event, const lp_opts_t* opts, bool half) {
    const int rc = nms_validate(pred, B, A, conf_thres, iou_thres, max_det, max_nms, workspace, workspace_bytes, out, counts);
    if (rc != LP_OK) return rc;
    const Opts o = resolve(opts);
    return pipelined_step(
        filter_stream, nms_stream, workspace_free_event, filtered_event, done_event, time_begin_event, time_end_event,
        [&](bool armed) { return nms_filter(pred, B, A, conf_thres, workspace, workspace_bytes, filter_stream, o, armed, half); },
        [&]() {
            return nms_suppress(pred, B, A, iou_thres, max_det, max_nms, workspace, workspace_bytes, out, counts, kept_anchor,
                                rescale, do_round, nms_stream, o, true, half);
        });
}

static int nms_serial(const float* pred, int B, int A, double conf_thres, double iou_thres, int max_det, int max_nms,
                      void* workspace, size_t workspace_bytes, float* out, int* counts, int* kept_anchor,
                      const float* rescale, int do_round, lp_stream_t stream, const lp_opts_t* opts, bool half) {
    int rc = nms_validate(pred, B, A, conf_thres, iou_thres, max_det, max_nms, workspace, workspace_bytes, out, counts);
    if (rc != LP_OK) return rc;
    const Opts o = resolve(opts);
    rc = nms_filter(pred, B, A, conf_thres, workspace, workspace_bytes, stream, o, false, half);
    if (rc != LP_OK) return rc;
    return nms_suppress(pred, B, A, iou_thres, max_det, max_nms, workspace, workspace_bytes, out, counts, kept_anchor, rescale,
                        do_round, stream, o, false, half);
}

LP_API int lp_nms_f32(const float* pred, int B, int A, double conf_thres, double iou_thres, int max_det, int max_nms,
                      void* workspace, size_t workspace_bytes, float* out, int* counts, int* kept_anchor,
                      const float* rescale, int do_round, lp_stream_t stream, const lp_opts_t* opts) {
    return nms_serial(pred, B, A, conf_thres, iou_thres, max_det, max_nms, workspace, workspace_bytes, out, counts,
                      kept_anchor, rescale, do_round, stream, opts, false);
}
LP_API int lp_nms_filter_f32(const float* pred, int B, int A, double conf_thres, void* workspace,
                             size_t workspace_bytes, lp_stream_t stream, const lp_opts_t* opts) {
    return nms_filter(pred, B, A, conf_thres, workspace, workspace_bytes, stream, resolve(opts), false, false);
}
LP_API int lp_nms_suppress_f32(const float* pred, int B, int A, double iou_thres, int max_det, int max_nms,
                               void* workspace, size_t workspace_bytes, float* out, int* counts, int* kept_anchor,
                               const float* rescale, int do_round, lp_stream_t stream, const lp_opts_t* opts) {
    return nms_suppress(pred, B, A, iou_thres, max_det, max_nms, workspace, workspace_bytes, out, counts, kept_anchor,
                        rescale, do_round, stream, resolve(opts), false, false);
}
LP_API int lp_nms_pipelined_f32(const float* pred, int B, int A, double conf_thres, double iou_thres, int max_det,
                                int max_nms, void* workspace, size_t workspace_bytes, float* out, int* counts,
                                int* kept_anchor, const float* rescale, int do_round, lp_stream_t filter_stream,
                                lp_stream_t nms_stream, void* workspace_free_event, void* filtered_event,
                                void* done_event, void* time_begin_event, void* time_end_event, const lp_opts_t* opts) {
    return nms_pipelined(pred, B, A, conf_thres, iou_thres, max_det, max_nms, workspace, workspace_bytes, out, counts,
                         kept_anchor, rescale, do_round, filter_stream, nms_stream, workspace_free_event, filtered_event,
                         done_event, time_begin_event, time_end_event, opts, false);
}

// ---- fp16 head tensors: pred holds IEEE halves, results == the f32 entries on pred.float()
LP_API int lp_nms_f16(const void* pred, int B, int A, double conf_thres, double iou_thres, int max_det, int max_nms,
                      void* workspace, size_t workspace_bytes, float* out, int* counts, int* kept_anchor,
                      const float* rescale, int do_round, lp_stream_t stream, const lp_opts_t* opts) {
    return nms_serial(static_cast<const float*>(pred), B, A, conf_thres, iou_thres, max_det, max_nms, workspace,
                      workspace_bytes, out, counts, kept_anchor, rescale, do_round, stream, opts, true);
}
LP_API int lp_nms_filter_f16(const void* pred, int B, int A, double conf_thres, void* workspace, size_t workspace_bytes,
                             lp_stream_t stream, const lp_opts_t* opts) {
    return nms_filter(static_cast<const float*>(pred), B, A, conf_thres, workspace, workspace_bytes, stream, resolve(opts),
                      false, true);
}
LP_API int lp_nms_suppress_f16(const void* pred, int B, int A, double iou_thres, int max_det, int max_nms, void* workspace,
                               size_t workspace_bytes, float* out, int* counts, int* kept_anchor, const float* rescale,
                               int do_round, lp_stream_t stream, const lp_opts_t* opts) {
    return nms_suppress(static_cast<const float*>(pred), B, A, iou_thres, max_det, max_nms, workspace, workspace_bytes, out,
                        counts, kept_anchor, rescale, do_round, stream, resolve(opts), false, true);
}
LP_API int lp_nms_pipelined_f16(const void* pred, int B, int A, double conf_thres, double iou_thres, int max_det,
                                int max_nms, void* workspace, size_t workspace_bytes, float* out, int* counts,
                                int* kept_anchor, const float* rescale, int do_round, lp_stream_t filter_stream,
                                lp_stream_t nms_stream, void* workspace_free_event, void* filtered_event,
                                void* done_event, void* time_begin_event, void* time_end_event, const lp_opts_t* opts) {
    return nms_pipelined(static_cast<const float*>(pred), B, A, conf_thres, iou_thres, max_det, max_nms, workspace,
                         workspace_bytes, out, counts, kept_anchor, rescale, do_round, filter_stream, nms_stream,
                         workspace_free_event, filtered_event, done_event, time_begin_event, time_end_event, opts, true);
}

LP_API int lp_detect_workspace_bytes(int B, int A, int max_det, size_t* out_bytes) {
    if (!out_bytes) return LP_E_NULL;
    if (!size_ok(B, A, max_det)) return LP_E_SIZE;
    *out_bytes = ws_layout(B, A, max_det).total_fused;
    return LP_OK;
}

// cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
        tried = true;
    }
    return fn;
}

// tensor maps of every level tensor for the decode kernel's TMA mode; false -> use the cp.async modes
static bool build_decode_maps(const DecodeLevel (&lv)[LP_MAX_LEVELS], int n_levels, int B, DecodeMaps& maps,
                              int n_tensors = DEC_TENSORS, bool half = false) {
    EncodeTiledFn encode = tensor_map_encoder();
    if (!encode) return false;
    for (int l = 0; l < n_levels; ++l) {
        for (int k = 0; k < n_tensors; ++k) {
            const int C = k == 8 ? 4 : k == 9 ? 8 : group_begin(k + 1) - group_begin(k);
            const float* base = k == 8 ? lv[l].reg : k == 9 ? lv[l].cor : lv[l].cls[k];
            const cuuint64_t dims[3] = {(cuuint64_t)lv[l].hw, (cuuint64_t)C, (cuuint64_t)B};
            const cuuint64_t esz = half ? 2 : 4;
            const cuuint64_t strides[2] = {(cuuint64_t)lv[l].hw * esz, (cuuint64_t)lv[l].hw * esz * C};
            const cuuint32_t box[3] = {(cuuint32_t)DEC_TILE, (cuuint32_t)C, 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            if (encode(&maps.m[l][k], half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return false;
        }
    }
    return true;
}

// validates a level table and lays it out for the kernels (anchor / tile offsets per level)
static int build_levels(const lp_level_t* levels, int n_levels, int B, DecodeLevel (&lv)[LP_MAX_LEVELS], int& A_out,
                        int& tiles_out, bool& bulk_out) {
    if (!levels) return LP_E_NULL;
    if (n_levels <= 0 || n_levels > LP_MAX_LEVELS || B <= 0) return LP_E_SIZE;
    long long A = 0;
    int tiles = 0;
    bool bulk = true;
    for (int l = 0; l < n_levels; ++l) {
        const lp_level_t& src = levels[l];
        if (src.h <= 0 || src.w <= 0) return LP_E_SIZE;
        if (!src.reg || !src.cor) return LP_E_NULL;
        DecodeLevel& d = lv[l];
        for (int g = 0; g < 8; ++g) {
            if (!src.cls[g]) return LP_E_NULL;
            if (!aligned(src.cls[g], 4)) return LP_E_ALIGN;
            d.cls[g] = src.cls[g];
            bulk = bulk && aligned(src.cls[g], 16);
        }
        if (!aligned(src.reg, 4) || !aligned(src.cor, 4)) return LP_E_ALIGN;
        d.reg = src.reg;
        d.cor = src.cor;
        d.w = src.w;
        d.hw = src.h * src.w;
        d.anchor_off = (int)A;
        d.tile_off = tiles;
        d.stride = src.stride;
        bulk = bulk && aligned(src.reg, 16) && aligned(src.cor, 16) && (d.hw % 4 == 0);
        A += d.hw;
        tiles += (d.hw + DEC_TILE - 1) / DEC_TILE;
        // row indices, tile indices and per-tensor element offsets (<= 37 channels) all fit 32 bits
        if (A * (long long)B >= (1ll << 31) || (long long)tiles * B >= (1ll << 31) ||
            (long long)B * 37 * d.hw >= (1ll << 31)) return LP_E_SIZE;
    }
    for (int l = n_levels; l < LP_MAX_LEVELS; ++l) lv[l] = lv[0];
    A_out = (int)A;
    tiles_out = tiles;
    bulk_out = bulk;
    return LP_OK;
}

static int detect_decode(const lp_level_t* levels, int n_levels, int B, float* out, lp_stream_t stream,
                         const lp_opts_t* opts, bool half_scores) {
    if (!out) return LP_E_NULL;
    if (!aligned(out, 8)) return LP_E_ALIGN;
    const Opts o = resolve(opts);
    DecodeParams p;
    int A = 0, tiles = 0;
    bool bulk = false;
    const int rc = build_levels(levels, n_levels, B, p.lv, A, tiles, bulk);
    if (rc != LP_OK) return rc;
    p.n_levels = n_levels;
    p.A = A;
    p.tiles_per_image = tiles;
    p.n_tiles = tiles * B;
    p.bulk_in = bulk ? 1 : 0;
    p.half_scores = half_scores ? 1 : 0;
    p.out = out;
    p.timing = o.timing;
    DecodeMaps maps;
    if (bulk && o.tma && aligned(out, 16) && build_decode_maps(p.lv, n_levels, B, maps)) p.bulk_in = 2;
    return (int)launch_decode(p, p.bulk_in == 2 ? &maps : nullptr, num_sms_cached(), static_cast<cudaStream_t>(stream));
}

LP_API int lp_detect_decode_f32(const lp_level_t* levels, int n_levels, int B, float* out, lp_stream_t stream,
                                const lp_opts_t* opts) {
    return detect_decode(levels, n_levels, B, out, stream, opts, false);
}
// The head tensor of the reference's model.half() forward from the (exactly upcast) half conv outputs: box /
// obj / corner columns as above -- the reference computes them in fp32 too, its anchors being fp32 -- and
// every class score rounded to the nearest IEEE half (torch.sigmoid on a half tensor) before it is stored.
LP_API int lp_detect_decode_half_scores_f32(const lp_level_t* levels, int n_levels, int B, float* out, lp_stream_t stream,
                                            const lp_opts_t* opts) {
    return detect_decode(levels, n_levels, B, out, stream, opts, true);
}

// Everything the fused entries can reject (apart from the fp16 shape rule, which needs the tensor
// maps), checked before anything is queued.
static int detect_validate(const lp_level_t* levels, int n_levels, int B, double conf_thres, double iou_thres, int max_det,
                           int max_nms, const void* workspace, size_t workspace_bytes, const float* out, const int* counts,
                           bool half) {
    if (!workspace || !counts || (!out && max_det > 0)) return LP_E_NULL;
    if (max_nms <= 0 || max_det < 0) return LP_E_SIZE;
    if (!(conf_thres >= 0.0 && conf_thres <= 1.0) || !(iou_thres >= 0.0 && iou_thres <= 1.0)) return LP_E_THRESHOLD;
    DecodeLevel lv[LP_MAX_LEVELS];
    int A = 0, tiles = 0;
    bool bulk = false;
    const int rc = build_levels(levels, n_levels, B, lv, A, tiles, bulk);
    if (rc != LP_OK) return rc;
    if (!size_ok(B, A, max_det)) return LP_E_SIZE;
    if (!aligned(workspace, WS_ALIGN) || !aligned(out, 4) || !aligned(counts, 4)) return LP_E_ALIGN;
    if (workspace_bytes < ws_layout(B, A, max_det).total_fused) return LP_E_WORKSPACE;
    if (half) {
        if (!bulk) return LP_E_ARG;
        for (int l = 0; l < n_levels; ++l)
            if (lv[l].hw % 8 != 0) return LP_E_ARG;
    }
    return LP_OK;
}

// overlapped: the caller runs K2 of another batch concurrently (the pipelined entry and the stand-alone
// stage entry, which exists for exactly that); false for the serial one-call path
static int detect_filter(const lp_level_t* levels, int n_levels, int B, double conf_thres, int max_det, void* workspace,
                         size_t workspace_bytes, lp_stream_t stream, const Opts& o, bool overlapped, bool armed, bool half) {
    if (max_det < 0) return LP_E_SIZE;
    if (!(conf_thres >= 0.0 && conf_thres <= 1.0)) return LP_E_THRESHOLD;
    LevelsFilterParams k;
    int A = 0, tiles = 0;
    bool bulk = false;
    int rc = build_levels(levels, n_levels, B, k.lv, A, tiles, bulk);
    if (rc != LP_OK) return rc;
    FilterParams f;
    NmsParams n;
    rc = nms_setup(nullptr, B, A, 0, workspace, (size_t)-1, f, n, o, false);
    if (rc != LP_OK) return rc;
    // the fused layout depends on max_det through the kept_* arrays that precede slot_of / rec
    const WsLayout w = ws_layout(B, A, max_det);
    if (workspace_bytes < w.total_fused) return LP_E_WORKSPACE;
    k.A = A;
    k.rec = reinterpret_cast<float*>(static_cast<char*>(workspace) + w.rec);
    k.slot_of = reinterpret_cast<unsigned*>(static_cast<char*>(workspace) + w.slot_of);
    k.n_levels = n_levels;
    k.tiles_per_image = tiles;
    k.n_tiles = tiles * B;
    k.conf = (float)conf_thres;
    k.keys = f.keys;
    k.counts = f.counts;
    k.key_stride = f.key_stride;
    k.tile_counter = f.tile_counter;
    k.timing = o.timing;
    DecodeMaps maps;
    k.half_levels = half ? 1 : 0;
    bool tma;
    if (half) {
        // fp16 level tensors exist as a TMA kernel only: 16-byte aligned tensors and row strides
        // (h*w % 8 == 0 at every level); the caller upcasts anything else and takes the f32 entry
        for (int l = 0; l < n_levels; ++l)
            if (k.lv[l].hw % 8 != 0) return LP_E_ARG;
        if (!bulk || !build_decode_maps(k.lv, n_levels, B, maps, NGROUP, true)) return LP_E_ARG;
        tma = true;
    } else {
        tma = bulk && o.tma && build_decode_maps(k.lv, n_levels, B, maps, NGROUP);
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!armed) {   // nothing is queued before the last check above has passed
        const cudaError_t e = cudaMemsetAsync(f.counts, 0, sizeof(int) * ((size_t)B + 1), s);
        if (e != cudaSuccess) return (int)e;
    }
    // like K1: one persistent CTA per SM, some SMs left free for K2 of the previous batch.  The
    // register-resident kernel is latency-bound and wants more SMs than K1 (a sixth left free is the
    // measured sweet spot; K2's CTAs also fit beside it on a shared SM).  The TMA kernel fills its
    // SM's shared memory, so K2 only runs on the SMs it leaves: one per image, up to half of them
    // (measured: B=32 -> 116 CTAs, B=64 -> 84 CTAs are the best pipelined points).  With nothing to
    // overlap it still runs best a little short of all SMs (B=64: 110 us on 124, 120 us on 148).
    int ctas = num_sms_cached();
    const int spare = tma && overlapped ? ctas / 2 : ctas / 6;
    ctas -= B < spare ? B : spare;
    if (o.ctas > 0) ctas = o.ctas < num_sms_cached() ? o.ctas : num_sms_cached();
    return (int)launch_levels_filter(k, tma ? &maps : nullptr, ctas, s);
}

static int detect_suppress(const lp_level_t* levels, int n_levels, int B, double iou_thres, int max_det, int max_nms,
                           void* workspace, size_t workspace_bytes, float* out, int* counts, int* kept_anchor,
                           const float* rescale, int do_round, lp_stream_t stream, const Opts& o, bool rearm) {
    if (!counts || (!out && max_det > 0)) return LP_E_NULL;
    if (max_nms <= 0) return LP_E_SIZE;
    if (!(iou_thres >= 0.0 && iou_thres <= 1.0)) return LP_E_THRESHOLD;
    if (!aligned(out, 4) || !aligned(counts, 4)) return LP_E_ALIGN;
    FilterParams f;
    NmsParams n;
    DecodeLevel lv[LP_MAX_LEVELS];
    int A = 0, tiles = 0;
    bool bulk = false;
    int rc = build_levels(levels, n_levels, B, lv, A, tiles, bulk);
    if (rc != LP_OK) return rc;
    rc = nms_setup(nullptr, B, A, max_det, workspace, workspace_bytes, f, n, o, false);
    if (rc != LP_OK) return rc;
    if (workspace_bytes < ws_layout(B, A, max_det).total_fused) return LP_E_WORKSPACE;
    float iou_floor = (float)iou_thres;
    if ((double)iou_floor > iou_thres) iou_floor = nextafterf(iou_floor, -INFINITY);
    n.iou_floor = iou_floor;
    n.max_nms = max_nms;
    n.out = out;
    n.out_counts = counts;
    n.kept_anchor = kept_anchor;
    n.rescale = rescale;
    n.do_round = do_round;
    n.from_levels = 1;
    n.rearm = rearm ? 1 : 0;
    return (int)launch_nms(n, B, static_cast<cudaStream_t>(stream));
}

static int detect_postprocess(const lp_level_t* levels, int n_levels, int B, double conf_thres, double iou_thres,
                              int max_det, int max_nms, void* workspace, size_t workspace_bytes, float* out, int* counts,
                              int* kept_anchor, const float* rescale, int do_round, lp_stream_t stream,
                              const lp_opts_t* opts, bool half) {
    int rc = detect_validate(levels, n_levels, B, conf_thres, iou_thres, max_det, max_nms, workspace, workspace_bytes, out,
                             counts, half);
    if (rc != LP_OK) return rc;
    const Opts o = resolve(opts);
    rc = detect_filter(levels, n_levels, B, conf_thres, max_det, workspace, workspace_bytes, stream, o, false, false, half);
    if (rc != LP_OK) return rc;
    return detect_suppress(levels, n_levels, B, iou_thres, max_det, max_nms, workspace, workspace_bytes, out, counts,
                           kept_anchor, rescale, do_round, stream, o, false);
}

static int detect_pipelined(const lp_level_t* levels, int n_levels, int B, double conf_thres, double iou_thres, int max_det,
                            int max_nms, void* workspace, size_t workspace_bytes, float* out, int* counts,
                            int* kept_anchor, const float* rescale, int do_round, lp_stream_t filter_stream,
                            lp_stream_t nms_stream, void* workspace_free_event, void* filtered_event, void* done_event,
                            void* time_begin_event, void* time_end_event, const lp_opts_t* opts, bool half) {
    const int rc = detect_validate(levels, n_levels, B, conf_thres, iou_thres, max_det, max_nms, workspace, workspace_bytes,
                                   out, counts, half);
    if (rc != LP_OK) return rc;
    const Opts o = resolve(opts);
    return pipelined_step(
        filter_stream, nms_stream, workspace_free_event, filtered_event, done_event, time_begin_event, time_end_event,
        [&](bool armed) {
            return detect_filter(levels, n_levels, B, conf_thres, max_det, workspace, workspace_bytes, filter_stream, o, true,
                                 armed, half);
        },
        [&]() {
            return detect_suppress(levels, n_levels, B, iou_thres, max_det, max_nms, workspace, workspace_bytes, out, counts,
                                   kept_anchor, rescale, do_round, nms_stream, o, true);
        });
}

LP_API int lp_detect_filter_f32(const lp_level_t* levels, int n_levels, int B, double conf_thres, int max_det,
                                void* workspace, size_t workspace_bytes, lp_stream_t stream, const lp_opts_t* opts) {
    return detect_filter(levels, n_levels, B, conf_thres, max_det, workspace, workspace_bytes, stream, resolve(opts), true,
                         false, false);
}
LP_API int lp_detect_filter_f16(const lp_level_t* levels, int n_levels, int B, double conf_thres, int max_det,
                                void* workspace, size_t workspace_bytes, lp_stream_t stream, const lp_opts_t* opts) {
    return detect_filter(levels, n_levels, B, conf_thres, max_det, workspace, workspace_bytes, stream, resolve(opts), true,
                         false, true);
}
LP_API int lp_detect_suppress_f32(const lp_level_t* levels, int n_levels, int B, double iou_thres, int max_det,
                                  int max_nms, void* workspace, size_t workspace_bytes, float* out, int* counts,
                                  int* kept_anchor, const float* rescale, int do_round, lp_stream_t stream,
                                  const lp_opts_t* opts) {
    return detect_suppress(levels, n_levels, B, iou_thres, max_det, max_nms, workspace, workspace_bytes, out, counts,
                           kept_anchor, rescale, do_round, stream, resolve(opts), false);
}
LP_API int lp_detect_postprocess_f32(const lp_level_t* levels, int n_levels, int B, double conf_thres, double iou_thres,
                                     int max_det, int max_nms, void* workspace, size_t workspace_bytes, float* out,
                                     int* counts, int* kept_anchor, const float* rescale, int do_round,
                                     lp_stream_t stream, const lp_opts_t* opts) {
    return detect_postprocess(levels, n_levels, B, conf_thres, iou_thres, max_det, max_nms, workspace, workspace_bytes, out,
                              counts, kept_anchor, rescale, do_round, stream, opts, false);
}
// fp16 level tensors (the pointers of lp_level_t then address IEEE halves): results == the f32 entry on
// the upcast tensors.  LP_E_ARG when a level's h*w is not a multiple of 8 or a tensor is not 16-byte
// aligned (the fp16 path is a TMA kernel only): upcast and call the f32 entry.
LP_API int lp_detect_postprocess_f16(const lp_level_t* levels, int n_levels, int B, double conf_thres, double iou_thres,
                                     int max_det, int max_nms, void* workspace, size_t workspace_bytes, float* out,
                                     int* counts, int* kept_anchor, const float* rescale, int do_round,
                                     lp_stream_t stream, const lp_opts_t* opts) {
    return detect_postprocess(levels, n_levels, B, conf_thres, iou_thres, max_det, max_nms, workspace, workspace_bytes, out,
                              counts, kept_anchor, rescale, do_round, stream, opts, true);
}
LP_API int lp_detect_pipelined_f32(const lp_level_t* levels, int n_levels, int B, double conf_thres, double iou_thres,
                                   int max_det, int max_nms, void* workspace, size_t workspace_bytes, float* out,
                                   int* counts, int* kept_anchor, const float* rescale, int do_round,
                                   lp_stream_t filter_stream, lp_stream_t nms_stream, void* workspace_free_event,
                                   void* filtered_event, void* done_event, void* time_begin_event,
                                   void* time_end_event, const lp_opts_t* opts) {
    return detect_pipelined(levels, n_levels, B, conf_thres, iou_thres, max_det, max_nms, workspace, workspace_bytes, out,
                            counts, kept_anchor, rescale, do_round, filter_stream, nms_stream, workspace_free_event,
                            filtered_event, done_event, time_begin_event, time_end_event, opts, false);
}
LP_API int lp_detect_pipelined_f16(const lp_level_t* levels, int n_levels, int B, double conf_thres, double iou_thres,
                                   int max_det, int max_nms, void* workspace, size_t workspace_bytes, float* out,
                                   int* counts, int* kept_anchor, const float* rescale, int do_round,
                                   lp_stream_t filter_stream, lp_stream_t nms_stream, void* workspace_free_event,
                                   void* filtered_event, void* done_event, void* time_begin_event,
                                   void* time_end_event, const lp_opts_t* opts) {
    return detect_pipelined(levels, n_levels, B, conf_thres, iou_thres, max_det, max_nms, workspace, workspace_bytes, out,
                            counts, kept_anchor, rescale, do_round, filter_stream, nms_stream, workspace_free_event,
                            filtered_event, done_event, time_begin_event, time_end_event, opts, true);
}

// lp_detect_pipelined_* plus the copy of the step's results into pinned HOST buffers: the production flow
// when the head runs on this GPU (level tensors in HBM, detections wanted on the host) as ONE call per step.
static int detect_pipelined_to_host(const lp_level_t* levels, int n_levels, int B, double conf_thres, double iou_thres,
                                    int max_det, int max_nms, void* workspace, size_t workspace_bytes, float* out, int* counts,
                                    const float* rescale, int do_round, lp_stream_t filter_stream, lp_stream_t nms_stream,
                                    void* workspace_free_event, void* filtered_event, void* done_event, float* out_host,
                                    int* counts_host, lp_stream_t copy_stream, void* copied_event, const lp_opts_t* opts,
                                    bool half) {
    if (!done_event || !out_host || !counts_host || !copied_event) return LP_E_NULL;
    int rc = detect_pipelined(levels, n_levels, B, conf_thres, iou_thres, max_det, max_nms, workspace, workspace_bytes, out,
                              counts, nullptr, rescale, do_round, filter_stream, nms_stream, workspace_free_event,
                              filtered_event, done_event, nullptr, nullptr, opts, half);
    if (rc != LP_OK) return rc;
    cudaStream_t sc = static_cast<cudaStream_t>(copy_stream);
    cudaError_t e = cudaStreamWaitEvent(sc, static_cast<cudaEvent_t>(done_event), 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(counts_host, counts, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost, sc);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(out_host, out, sizeof(float) * LP_OUT * (size_t)B * (size_t)max_det, cudaMemcpyDeviceToHost, sc);
    if (e == cudaSuccess) e = cudaEventRecord(static_cast<cudaEvent_t>(copied_event), sc);
    return (int)e;
}
LP_API int lp_detect_pipelined_to_host_f32(const lp_level_t* levels, int n_levels, int B, double conf_thres, double iou_thres,
                                           int max_det, int max_nms, void* workspace, size_t workspace_bytes, float* out,
                                           int* counts, const float* rescale, int do_round, lp_stream_t filter_stream,
                                           lp_stream_t nms_stream, void* workspace_free_event, void* filtered_event,
                                           void* done_event, float* out_host, int* counts_host, lp_stream_t copy_stream,
                                           void* copied_event, const lp_opts_t* opts) {
    return detect_pipelined_to_host(levels, n_levels, B, conf_thres, iou_thres, max_det, max_nms, workspace, workspace_bytes,
                                    out, counts, rescale, do_round, filter_stream, nms_stream, workspace_free_event,
                                    filtered_event, done_event, out_host, counts_host, copy_stream, copied_event, opts, false);
}
LP_API int lp_detect_pipelined_to_host_f16(const lp_level_t* levels, int n_levels, int B, double conf_thres, double iou_thres,
                                           int max_det, int max_nms, void* workspace, size_t workspace_bytes, float* out,
                                           int* counts, const float* rescale, int do_round, lp_stream_t filter_stream,
                                           lp_stream_t nms_stream, void* workspace_free_event, void* filtered_event,
                                           void* done_event, float* out_host, int* counts_host, lp_stream_t copy_stream,
                                           void* copied_event, const lp_opts_t* opts) {
    return detect_pipelined_to_host(levels, n_levels, B, conf_thres, iou_thres, max_det, max_nms, workspace, workspace_bytes,
                                    out, counts, rescale, do_round, filter_stream, nms_stream, workspace_free_event,
                                    filtered_event, done_event, out_host, counts_host, copy_stream, copied_event, opts, true);
}

LP_API int lp_debug_sigmoid_f32(const float* in, long long n, float* out, lp_stream_t stream) {
    if (!in || !out) return LP_E_NULL;
    if (n < 0) return LP_E_SIZE;
    return (int)launch_sigmoid(in, n, out, static_cast<cudaStream_t>(stream));
}

LP_API int lp_generate_anchors_f32(const int* h, const int* w, const float* stride, int n_levels, float grid_cell_offset,
                            float* anchor_points, float* stride_tensor, lp_stream_t stream) {
    if (!h || !w || !stride || !anchor_points || !stride_tensor) return LP_E_NULL;
    if (n_levels <= 0 || n_levels > LP_MAX_LEVELS) return LP_E_SIZE;
    AnchorLevels lv;
    long long A = 0;
    for (int l = 0; l < LP_MAX_LEVELS; ++l) {
        const int s = l < n_levels ? l : 0;
        if (h[s] <= 0 || w[s] <= 0) return LP_E_SIZE;
        lv.w[l] = w[s];
        lv.hw[l] = h[s] * w[s];
        lv.off[l] = (int)A;
        lv.stride[l] = stride[s];
        if (l < n_levels) A += lv.hw[l];
        if (A >= (1ll << 31)) return LP_E_SIZE;
    }
    lv.n_levels = n_levels;
    lv.A = (int)A;
    lv.offset = grid_cell_offset;
    return (int)launch_anchors(lv, anchor_points, stride_tensor, static_cast<cudaStream_t>(stream));
}

LP_API int lp_dist2bbox_f32(const float* distance, const float* anchor_points, long long n, int A, int xywh, float* out,
                     lp_stream_t stream) {
    if (!distance || !anchor_points || !out) return LP_E_NULL;
    if (n < 0 || A <= 0) return LP_E_SIZE;
    if (!aligned(distance, 16) || !aligned(out, 16) || !aligned(anchor_points, 8)) return LP_E_ALIGN;
    return (int)launch_dist2bbox(distance, anchor_points, n, A, xywh, out, static_cast<cudaStream_t>(stream));
}

LP_API int lp_dist2cor_f32(const float* distance, const float* anchor_points, long long n, int A, float* out, lp_stream_t stream) {
    if (!distance || !anchor_points || !out) return LP_E_NULL;
    if (n < 0 || A <= 0) return LP_E_SIZE;
    if (!aligned(distance, 16) || !aligned(out, 16) || !aligned(anchor_points, 8)) return LP_E_ALIGN;
    return (int)launch_dist2cor(distance, anchor_points, n, A, out, static_cast<cudaStream_t>(stream));
}

LP_API int lp_xywh2xyxy_f32(const float* in, long long n, long long in_stride, float* out, long long out_stride, lp_stream_t stream) {
    if (!in || !out) return LP_E_NULL;
    if (n < 0 || in_stride < 4 || out_stride < 4) return LP_E_SIZE;
    return (int)launch_xywh2xyxy(in, n, in_stride, out, out_stride, static_cast<cudaStream_t>(stream));
}

LP_API int lp_rescale_f32(float* rows, long long k, long long row_stride, float pad_x, float pad_y, float ratio, float w0,
                   float h0, int do_round, lp_stream_t stream) {
    if (!rows && k > 0) return LP_E_NULL;
    if (k < 0 || row_stride < 12) return LP_E_SIZE;
    if (!(ratio > 0.0f)) return LP_E_ARG;
    return (int)launch_rescale(rows, k, row_stride, pad_x, pad_y, ratio, w0, h0, do_round, static_cast<cudaStream_t>(stream));
}

LP_API int lp_rescale_batch_f32(float* det, const int* counts, int B, int max_det, const float* params, int do_round,
                         lp_stream_t stream) {
    if (!det || !counts || !params) return LP_E_NULL;
    if (B <= 0 || B > 65535 || max_det < 0) return LP_E_SIZE;
    return (int)launch_rescale_batch(det, counts, B, max_det, params, do_round, static_cast<cudaStream_t>(stream));
}

LP_API int lp_txt_records_f32(const float* det, const int* counts, int B, int max_det, const float* src_wh,
                              float* records, lp_stream_t stream) {
    if (!det || !counts || !src_wh || !records) return LP_E_NULL;
    if (B <= 0 || B > 65535 || max_det < 0) return LP_E_SIZE;
    return (int)launch_txt_records(det, counts, B, max_det, src_wh, records, static_cast<cudaStream_t>(stream));
}

LP_API int lp_txt_lines_host(const float* records_host, long long n, char* buf, size_t buf_bytes, size_t* written) {
    if ((!records_host && n > 0) || !buf || !written) return LP_E_NULL;
    if (n < 0) return LP_E_SIZE;
    size_t off = 0;
    for (long long r = 0; r < n; ++r) {
        const float* v = records_host + r * 21;
        for (int i = 0; i < 20; ++i) {  // ('%g ' * 20).rstrip() % line, inferer.py:120
            const int m = snprintf(buf + off, off < buf_bytes ? buf_bytes - off : 0, i == 19 ? "%g\n" : "%g ", (double)v[i]);
            if (m < 0) return LP_E_ARG;
            off += (size_t)m;
            if (off >= buf_bytes) return LP_E_WORKSPACE;
        }
    }
    *written = off;
    return LP_OK;
}

LP_API int lp_prepare_targets_f32(const float* targets, int T, float w, float h, float* out, int* out_image,
                                  lp_stream_t stream) {
    if (T == 0) return LP_OK;
    if (!targets || !out || !out_image) return LP_E_NULL;
    if (T < 0) return LP_E_SIZE;
    return (int)launch_prepare_targets(targets, T, w, h, out, out_image, static_cast<cudaStream_t>(stream));
}

LP_API int lp_eval_match_f32(const float* det, const int* counts, int B, int max_det, const float* targets,
                             const int* target_image, int T, float* match, lp_stream_t stream) {
    if (T == 0) return LP_OK;
    if (!det || !counts || !targets || !target_image || !match) return LP_E_NULL;
    if (B <= 0 || max_det <= 0 || T < 0) return LP_E_SIZE;
    if (!aligned(det, 16)) return LP_E_ALIGN;
    return (int)launch_eval_match(det, counts, B, max_det, targets, target_image, T, match, static_cast<cudaStream_t>(stream));
}

// Counters and summary of Evaler.eval (yolov6/core/evaler.py:160-283) from the per-target matches,
// in image / target order.  Quirks kept: the fp32 IoU is compared against the Python doubles
// 0.5 + 0.05 n rounded to fp32; an IoU that falls into no bin (exactly 1.0) reuses the previous
// target's bin for the correctness counters (stale `iou_idx`) and is skipped by pred_cnts.
LP_API int lp_eval_accumulate_host(const float* match_host, const int* target_image_host, const int* counts_host, int B,
                                   int T, long long* counters, double* summary) {
    if ((T > 0 && (!match_host || !target_image_host)) || !counts_host || !counters || !summary) return LP_E_NULL;
    if (B <= 0 || T < 0) return LP_E_SIZE;
    double iou_list[10];
    float lo[10], hi[10];
    for (int n = 0; n < 10; ++n) {
        iou_list[n] = 0.5 + n * 0.05;
        lo[n] = (float)iou_list[n];
        hi[n] = (float)(iou_list[n] + 0.05);
    }
    long long true_cnt = 0, pred_cnt = 0, pred_cnts[10] = {0}, cor_right[10] = {0}, cls_right[10] = {0}, right[10] = {0};
    int iou_idx = -1;
    int t = 0;
    while (t < T) {
        const int b = target_image_host[t];
        if (b < 0 || b >= B) return LP_E_ARG;
        int e = t;
        while (e < T && target_image_host[e] == b) ++e;
        if (e < T && target_image_host[e] < b) return LP_E_ARG;  // must be grouped in image order
        true_cnt += e - t;
        if (counts_host[b] > 0) {
            for (int k = t; k < e; ++k) {
                const float v = match_host[4 * k];
                if (v < 0.5f) continue;
                if (v >= 0.7f) ++pred_cnt;
                for (int n = 0; n < 10; ++n)
                    if (v >= lo[n] && v < hi[n]) { iou_idx = n; break; }
                if (iou_idx < 0) return LP_E_ARG;  // the reference raises NameError here
                const bool is_cor = match_host[4 * k + 2] != 0.0f, is_cls = match_host[4 * k + 3] != 0.0f;
                if (is_cor) ++cor_right[iou_idx];
                if (is_cls) ++cls_right[iou_idx];
                if (is_cor && is_cls) ++right[iou_idx];
            }
            for (int k = t; k < e; ++k) {
                const float v = match_host[4 * k];
                if (v < 0.5f) continue;
                for (int n = 0; n < 10; ++n)
                    if (v >= lo[n] && v < hi[n]) { ++pred_cnts[n]; break; }
            }
        }
        t = e;
    }
    counters[0] = true_cnt;
    counters[1] = pred_cnt;
    for (int n = 0; n < 10; ++n) {
        counters[2 + n] = pred_cnts[n];
        counters[12 + n] = cor_right[n];
        counters[22 + n] = cls_right[n];
        counters[32 + n] = right[n];
    }
    // evaler.py:247-283
    double m5095 = 0.0;
    long long valid = 0, right_50 = 0, pred_50 = 0, right_75 = 0, pred_75 = 0, t_right = 0;
    for (int i = 0; i < 10; ++i) {
        const double m = pred_cnts[i] > 0 ? (double)right[i] / (double)pred_cnts[i] : -(double)(right[i] == pred_cnts[i]);
        summary[5 + i] = m;
        if (m != -1.0) { m5095 += m; ++valid; }
        right_50 += right[i];
        pred_50 += pred_cnts[i];
        if (iou_list[i] >= 0.75) { right_75 += right[i]; pred_75 += pred_cnts[i]; }
        if (iou_list[i] >= 0.7) t_right += right[i];
    }
    summary[0] = pred_cnt > 0 ? (double)t_right / (double)pred_cnt : 0.0;
    summary[1] = pred_50 > 0 ? (double)right_50 / (double)pred_50 : 0.0;
    summary[2] = pred_75 > 0 ? (double)right_75 / (double)pred_75 : 0.0;
    summary[3] = valid > 0 ? m5095 / (double)valid : 0.0;
    long long acc = 0;
    for (int i = 0; i < 10; ++i) {
        acc += right[i];
        summary[15 + i] = true_cnt > 0 ? (double)acc / (double)true_cnt : 0.0;
    }
    summary[4] = true_cnt > 0 ? (double)acc / (double)true_cnt : 0.0;  // the reference divides by zero here
    return LP_OK;
}

}  // extern "C"
