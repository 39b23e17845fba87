// K2 -- per-image sort -> exact greedy NMS -> fused gather (+ optional rescale/round).
//
// One 1024-thread CTA per image; images are independent (nms.py:68 loop body), so a batch is
// B concurrent CTAs and the stage's latency is that of the slowest image.
//
//   sort    keys (score desc, anchor asc) from K1, bitonic in shared memory (<= 16384 keys) or, for
//           larger candidate sets, in the global workspace.  Ascending key order reproduces
//           torchvision's stable descending sort over the anchor-ordered compaction
//           (nms.py:97,121); more than max_nms candidates are cut to the first max_nms
//           (nms.py:115-116).
//   NMS     torchvision.ops.nms CPU semantics (call site nms.py:121), bit-exact fp32 IoU, evaluated
//           lazily: walk the sorted list in 1024-wide windows; a window is first tested against
//           every box kept so far, then resolved in order -- each newly kept box suppresses the
//           rest of its window (one IoU per thread, one __syncthreads per kept box).  Work is
//           bounded by max_det * N IoUs and stops as soon as max_det rows are kept, which the
//           reference's truncation keep[:max_det] (nms.py:122-123) makes legal.
//   gather  one warp per kept row re-reads the row from `pred`, recomputes nms.py:76-96 (group
//           maxima with first-index argmax, xyxy box, corners) and writes the 28-float output row,
//           optionally mapped back to source coordinates (inferer.py:203-228, :100).
#include "kernels.cuh"

namespace lp {

constexpr int NMS_THREADS = 1024;
constexpr int NMS_WARPS = NMS_THREADS / 32;
constexpr int SORT_SMEM_KEYS = 16384;  // 128 KB
constexpr int KEPT_SMEM = 1024;        // kept boxes cached in shared memory (rest via L2)

__device__ __forceinline__ unsigned next_pow2(unsigned n) { return n <= 1 ? 1u : 1u << (32 - __clz(n - 1)); }

// In-place ascending bitonic sort of n (power of two) keys; `keys` is shared or global memory.
template <bool kGlobal>
__device__ void bitonic_sort(unsigned long long* keys, unsigned n) {
    for (unsigned k = 2; k <= n; k <<= 1) {
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned t = threadIdx.x; t < (n >> 1); t += NMS_THREADS) {
                const unsigned i = 2 * t - (t & (j - 1));  // bit j of i is clear
                const unsigned l = i | j;
                unsigned long long a, b;
                if (kGlobal) {
                    a = __ldcg(keys + i);
                    b = __ldcg(keys + l);
                } else {
                    a = keys[i];
                    b = keys[l];
                }
                const bool up = (i & k) == 0;
                if ((a > b) == up) {
                    if (kGlobal) {
                        __stcg(keys + i, b);
                        __stcg(keys + l, a);
                    } else {
                        keys[i] = b;
                        keys[l] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(NMS_THREADS, 1) nms_kernel(const NmsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* skeys = reinterpret_cast<unsigned long long*>(smem_raw);  // [sort_smem_keys]
    __shared__ float4 wbox[NMS_THREADS];      // boxes of the current window
    __shared__ float4 kbox[KEPT_SMEM];        // first KEPT_SMEM kept boxes
    __shared__ unsigned words[2][NMS_WARPS];  // alive bitmask of the window, double-buffered

    const unsigned b = blockIdx.x;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* pred = p.pred + (size_t)b * p.A * ROW;
    float4* kept_box = p.kept_box + (size_t)b * p.max_det;
    int* kept_anchor = p.kept_anchor_ws + (size_t)b * p.max_det;

    unsigned N = (unsigned)p.counts[b];
    if (N > p.A) N = p.A;
    int n_keep = 0;

    if (N > 0 && p.max_det > 0) {
        // ------------------------------------------------------------------ sort
        const unsigned npad = next_pow2(N);
        unsigned long long* gkeys = p.keys + (size_t)b * p.key_stride;
        const unsigned long long* sorted;
        if (npad <= (unsigned)p.sort_smem_keys) {
            for (unsigned i = tid; i < npad; i += NMS_THREADS) skeys[i] = i < N ? __ldcg(gkeys + i) : ~0ull;
            __syncthreads();
            bitonic_sort<false>(skeys, npad);
            sorted = skeys;
        } else {  // key_stride >= npad is guaranteed by lp_nms_workspace_bytes
            for (unsigned i = N + tid; i < npad; i += NMS_THREADS) __stcg(gkeys + i, ~0ull);
            __syncthreads();
            bitonic_sort<true>(gkeys, npad);
            sorted = gkeys;
        }
        const bool sorted_global = sorted != skeys;
        if (N > (unsigned)p.max_nms) N = p.max_nms;  // nms.py:115-116

        // ------------------------------------------------------------------ windowed lazy NMS
        for (unsigned w0 = 0; w0 < N && n_keep < p.max_det; w0 += NMS_THREADS) {
            const unsigned pos = w0 + tid;
            bool alive = pos < N;
            unsigned anchor = 0;
            float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
            float area = 0.f;
            if (alive) {
                anchor = (unsigned)(sorted_global ? __ldcg(sorted + pos) : sorted[pos]);
                const float2* r = reinterpret_cast<const float2*>(pred + (size_t)anchor * ROW);
                const float2 c = __ldg(r), s = __ldg(r + 1);
                box = xywh_to_xyxy(c.x, c.y, s.x, s.y);  // nms.py:79
                area = box_area(box);
            }
            wbox[tid] = box;
            // suppression by boxes kept in earlier windows
            for (int k = 0; k < n_keep; ++k) {
                const float4 kb = k < KEPT_SMEM ? kbox[k] : __ldcg(kept_box + k);
                if (alive && iou_exceeds(kb, box_area(kb), box, area, p.iou_floor)) alive = false;
            }
            unsigned word = __ballot_sync(0xffffffffu, alive);
            if (lane == 0) words[0][warp] = word;
            __syncthreads();

            int cur = -1;
            unsigned parity = 0;
            while (true) {
                // first alive position after `cur` (every warp computes it redundantly)
                unsigned wl = words[parity][lane];
                if (cur >= 0) {
                    const unsigned cw = (unsigned)cur >> 5, cb = (unsigned)cur & 31;
                    if (lane < cw) wl = 0;
                    else if (lane == cw) wl &= ~(0xffffffffu >> (31 - cb));
                }
                const unsigned nz = __ballot_sync(0xffffffffu, wl != 0);
                if (!nz) break;
                const int fw = __ffs(nz) - 1;
                const unsigned fwl = __shfl_sync(0xffffffffu, wl, fw);
                const int nxt = fw * 32 + (__ffs(fwl) - 1);
                if ((int)tid == nxt) {  // this thread's candidate is kept
                    if (n_keep < KEPT_SMEM) kbox[n_keep] = box;
                    kept_box[n_keep] = box;
                    kept_anchor[n_keep] = (int)anchor;
                }
                ++n_keep;
                cur = nxt;
                if (n_keep >= p.max_det) break;
                const float4 kb = wbox[nxt];
                if (alive && (int)tid > nxt && iou_exceeds(kb, box_area(kb), box, area, p.iou_floor)) alive = false;
                parity ^= 1;
                word = __ballot_sync(0xffffffffu, alive);
                if (lane == 0) words[parity][warp] = word;
                __syncthreads();
            }
            __syncthreads();  // kbox / kept_* visible, wbox and words free for the next window
        }
    }
    __syncthreads();

    // ---------------------------------------------------------------------- gather
    if (tid == 0) p.out_counts[b] = n_keep;
    float pad_x = 0.f, pad_y = 0.f, ratio = 1.f, w0f = 0.f, h0f = 0.f;
    const bool do_rescale = p.rescale != nullptr;
    if (do_rescale) {
        const float* rp = p.rescale + (size_t)b * 5;
        pad_x = rp[0]; pad_y = rp[1]; ratio = rp[2]; w0f = rp[3]; h0f = rp[4];
    }
    for (int k = warp; k < n_keep; k += NMS_WARPS) {
        const int anchor = kept_anchor[k];
        const float* row = pred + (size_t)anchor * ROW;
        const float obj = __ldg(row + 4);
        float val = 0.f;
        if (lane < 4) {
            const float4 bx = k < KEPT_SMEM ? kbox[k] : kept_box[k];
            val = lane == 0 ? bx.x : lane == 1 ? bx.y : lane == 2 ? bx.z : bx.w;
        } else if (lane < 12) {
            val = __ldg(row + lane + 1);  // corners: columns 5..12 -> output 4..11 (nms.py:94)
        }
#pragma unroll
        for (int g = 0; g < NGROUP; ++g) {
            constexpr int kInvalid = 1 << 20;
            const int s = group_begin(g), e = group_begin(g + 1);
            float best = -INFINITY;
            int bi = kInvalid;
            // widths are 31, 24 or 37: at most two columns per lane
            if (s + (int)lane < e) {
                best = __fmul_rn(__ldg(row + s + lane), obj);  // nms.py:76
                bi = lane;
            }
            if (s + 32 + (int)lane < e) {
                const float v = __fmul_rn(__ldg(row + s + 32 + lane), obj);
                if (v > best) { best = v; bi = 32 + lane; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {  // max, ties -> lowest index (torch.max on CPU)
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if ((int)lane == 12 + g) val = best;
            if ((int)lane == 20 + g) val = (float)bi;
        }
        if (do_rescale && lane < 12) {
            val = (lane & 1) ? rescale_coord(val, pad_y, ratio, h0f, p.do_round)
                             : rescale_coord(val, pad_x, ratio, w0f, p.do_round);
        }
        if (lane < OUTW) p.out[((size_t)b * p.max_det + k) * OUTW + lane] = val;
        if (p.kept_anchor != nullptr && lane == 0) p.kept_anchor[(size_t)b * p.max_det + k] = anchor;
    }
}

cudaError_t launch_nms(const NmsParams& p, int B, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    const size_t smem = (size_t)p.sort_smem_keys * sizeof(unsigned long long);
    cudaError_t e = cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    nms_kernel<<<B, NMS_THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

int nms_sort_smem_keys(unsigned A) {
    unsigned n = 32;
    while (n < A && n < (unsigned)SORT_SMEM_KEYS) n <<= 1;
    return (int)n;
}

}  // namespace lp
