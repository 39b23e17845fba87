// K2 -- per-image sort -> exact greedy NMS -> fused gather (+ optional rescale/round).
//
// One 1024-thread CTA per image; images are independent (nms.py:68 loop body), so a batch is
// B concurrent CTAs and the stage's latency is that of the slowest image.  Only the warps that
// own candidates take part in the sort/NMS barriers (named barrier over `t_act` threads).
//
//   sort    keys (score desc, anchor asc) from K1: rank sort (<= 512 keys), bitonic in shared
//           memory (<= 16384 keys) or bitonic in the global workspace beyond that.  Ascending key
//           order reproduces torchvision's stable descending sort over the anchor-ordered
//           compaction (nms.py:97,121); more than max_nms candidates are cut to the first
//           max_nms (nms.py:115-116).
//   NMS     torchvision.ops.nms CPU semantics (call site nms.py:121), bit-exact fp32 IoU, evaluated
//           lazily and chunk-wise: the sorted list is walked in 1024-wide windows (one candidate
//           per thread); a window is first tested against every box kept so far, then resolved
//           32 candidates (one warp's chunk) at a time -- every later candidate collects the
//           bitmask S of chunk members that would suppress it, the chunk's own warp settles which
//           members are kept with a ballot fixed-point (exactly the greedy order), and the kept
//           mask K is broadcast: candidates with S & K die.  Serial depth is the number of
//           non-empty chunks (<= N/32), not the number of kept boxes; IoU work is only done
//           against still-alive chunk members, and everything stops as soon as max_det rows are
//           kept, which the reference's truncation keep[:max_det] (nms.py:122-123) makes legal.
//   gather  one half-warp per kept row re-reads the row from `pred` (prefetched into L2 when it
//           was kept), recomputes nms.py:76-96 (group maxima with first-index argmax, xyxy box,
//           corners) and writes the 28-float output row, optionally mapped back to source
//           coordinates (inferer.py:203-228, :100).
#include "kernels.cuh"

namespace lp {

constexpr int NMS_THREADS = 1024;
constexpr int NMS_WARPS = NMS_THREADS / 32;
constexpr int SORT_SMEM_KEYS = 16384;  // 128 KB
constexpr int RANK_SORT_MAX = 512;     // rank sort needs 2 * RANK_SORT_MAX keys of shared memory
constexpr int KEPT_SMEM = 1024;        // kept boxes / anchors cached in shared memory (rest via L2)

__device__ __forceinline__ unsigned next_pow2(unsigned n) { return n <= 1 ? 1u : 1u << (32 - __clz(n - 1)); }

// barrier over the first `nthreads` threads of the CTA (a multiple of 32); id 1, id 0 is __syncthreads
__device__ __forceinline__ void bar_active(unsigned nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// In-place ascending bitonic sort of n (power of two) keys; `keys` is shared or global memory.
template <bool kGlobal>
__device__ void bitonic_sort(unsigned long long* keys, unsigned n, unsigned t_act) {
    for (unsigned k = 2; k <= n; k <<= 1) {
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned t = threadIdx.x; t < (n >> 1); t += t_act) {
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
            bar_active(t_act);
        }
    }
}

__global__ void __launch_bounds__(NMS_THREADS, 1) nms_kernel(const NmsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* skeys = reinterpret_cast<unsigned long long*>(smem_raw);  // [sort_smem_keys]
    __shared__ float4 wbox[NMS_THREADS];      // boxes of the current window
    __shared__ float4 kbox[KEPT_SMEM];        // first KEPT_SMEM kept boxes
    __shared__ int kanchor[KEPT_SMEM];        // and their anchors
    __shared__ unsigned words[NMS_WARPS];     // alive bitmask of the window, one word per chunk
    __shared__ unsigned kmask;                // kept mask of the chunk being resolved
    __shared__ int s_nkeep;

    const unsigned b = blockIdx.x;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* pred = p.pred + (size_t)b * p.A * ROW;
    float4* kept_box = p.kept_box + (size_t)b * p.max_det;
    int* kept_anchor = p.kept_anchor_ws + (size_t)b * p.max_det;

    unsigned N = (unsigned)p.counts[b];
    if (N > p.A) N = p.A;
    if (tid < NMS_WARPS) words[tid] = 0;
    if (tid == 0) s_nkeep = 0;
    __syncthreads();

    // threads that own a candidate in a window; the others go straight to the gather
    const unsigned t_act = N >= NMS_THREADS ? NMS_THREADS : ((N + 31u) & ~31u);
    if (tid < t_act && p.max_det > 0) {
        int n_keep = 0;
        // ------------------------------------------------------------------ sort
        unsigned long long* gkeys = p.keys + (size_t)b * p.key_stride;
        const unsigned long long* sorted;
        bool sorted_global = false;
        if (N <= RANK_SORT_MAX) {
            // rank sort: keys are distinct (they embed the anchor), rank = number of smaller keys
            unsigned long long key = ~0ull;
            if (tid < N) key = __ldcg(gkeys + tid);
            skeys[tid] = key;
            bar_active(t_act);
            unsigned rank = 0;
#pragma unroll 4
            for (unsigned i = 0; i < N; ++i) rank += skeys[i] < key;
            if (tid < N) skeys[RANK_SORT_MAX + rank] = key;
            bar_active(t_act);
            sorted = skeys + RANK_SORT_MAX;
        } else {
            const unsigned npad = next_pow2(N);
            if (npad <= (unsigned)p.sort_smem_keys) {
                for (unsigned i = tid; i < npad; i += t_act) skeys[i] = i < N ? __ldcg(gkeys + i) : ~0ull;
                bar_active(t_act);
                bitonic_sort<false>(skeys, npad, t_act);
                sorted = skeys;
            } else {  // key_stride >= npad is guaranteed by lp_nms_workspace_bytes
                for (unsigned i = N + tid; i < npad; i += t_act) __stcg(gkeys + i, ~0ull);
                bar_active(t_act);
                bitonic_sort<true>(gkeys, npad, t_act);
                sorted = gkeys;
                sorted_global = true;
            }
        }
        if (N > (unsigned)p.max_nms) N = p.max_nms;  // nms.py:115-116

        // ------------------------------------------------------------------ windowed, chunked greedy NMS
        const unsigned lower = (1u << lane) - 1u;
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
            {
                const unsigned word = __ballot_sync(0xffffffffu, alive);
                if (lane == 0) words[warp] = word;
            }
            bar_active(t_act);

            int c = -1;
            while (true) {
                // next chunk (warp) that still has alive members; every warp computes it redundantly
                const unsigned wl = (int)lane > c ? words[lane] : 0u;
                const unsigned nz = __ballot_sync(0xffffffffu, wl != 0);
                if (!nz) break;
                c = __ffs(nz) - 1;
                const unsigned A = __shfl_sync(0xffffffffu, wl, c);
                // S: members of chunk c that suppress this thread's candidate if they are kept
                unsigned S = 0;
                if ((int)warp >= c && alive) {
                    const float4* cb = wbox + c * 32;
                    for (unsigned m = A; m; m &= m - 1) {
                        const int i = __ffs(m) - 1;
                        const float4 kb = cb[i];
                        if (iou_exceeds(kb, box_area(kb), box, area, p.iou_floor)) S |= 1u << i;
                    }
                    if ((int)warp == c) S &= lower;  // only earlier members of the own chunk count
                }
                if ((int)warp == c) {
                    // greedy inside the chunk: K_j = A_j and no kept earlier member suppresses j.
                    // Iterating K <- F(K) fixes one more leading member per round; the unique
                    // fixed point is the sequential result.
                    const bool in_a = (A >> lane) & 1u;
                    unsigned K = A;
                    while (true) {
                        const unsigned K2 = __ballot_sync(0xffffffffu, in_a && (S & K) == 0);
                        if (K2 == K) break;
                        K = K2;
                    }
                    const int room = p.max_det - n_keep;
                    if (__popc(K) > room) K &= (1u << __fns(K, 0, room + 1)) - 1u;  // first `room` members only
                    if ((K >> lane) & 1u) {
                        const int k = n_keep + __popc(K & lower);
                        if (k < KEPT_SMEM) { kbox[k] = box; kanchor[k] = (int)anchor; }
                        kept_box[k] = box;
                        kept_anchor[k] = (int)anchor;
                        const char* row = reinterpret_cast<const char*>(pred + (size_t)anchor * ROW);
#pragma unroll
                        for (int l = 0; l < ROW * 4; l += 128) prefetch_l2(row + l);  // for the gather
                    }
                    if (lane == 0) kmask = K;
                }
                bar_active(t_act);
                const unsigned K = kmask;
                n_keep += __popc(K);
                if ((S & K) != 0 || (int)warp == c) alive = false;
                if (n_keep >= p.max_det) break;
                if ((int)warp > c) {
                    const unsigned word = __ballot_sync(0xffffffffu, alive);
                    if (lane == 0) words[warp] = word;
                }
                bar_active(t_act);
            }
            bar_active(t_act);  // kbox / kept_* visible, wbox / words / kmask free for the next window
        }
        if (tid == 0) s_nkeep = n_keep;
    }
    __syncthreads();

    // ---------------------------------------------------------------------- gather
    const int n_keep = s_nkeep;
    if (tid == 0) p.out_counts[b] = n_keep;
    float pad_x = 0.f, pad_y = 0.f, ratio = 1.f, w0f = 0.f, h0f = 0.f;
    const bool do_rescale = p.rescale != nullptr;
    if (do_rescale) {
        const float* rp = p.rescale + (size_t)b * 5;
        pad_x = rp[0]; pad_y = rp[1]; ratio = rp[2]; w0f = rp[3]; h0f = rp[4];
    }
    const unsigned hl = tid & 15;  // lane inside the half-warp that owns a row
    for (int k0 = 2 * warp; k0 < n_keep; k0 += NMS_THREADS / 16) {  // warp-uniform trip count (shuffles inside)
        const bool live = k0 + (int)(lane >> 4) < n_keep;
        const int k = live ? k0 + (int)(lane >> 4) : k0;  // an idle upper half-warp mirrors the lower one, stores masked
        const int anchor = k < KEPT_SMEM ? kanchor[k] : kept_anchor[k];
        const float* row = pred + (size_t)anchor * ROW;
        // issue every load of the row before the first use: one memory round trip per row
        const float obj = __ldg(row + 4);
        const float corner = (hl >= 4 && hl < 12) ? __ldg(row + hl + 1) : 0.f;  // cols 5..12 -> out 4..11
        float v[NGROUP][3];
#pragma unroll
        for (int g = 0; g < NGROUP; ++g) {
            const int s = group_begin(g), e = group_begin(g + 1);
#pragma unroll
            for (int q = 0; q < 3; ++q) v[g][q] = (s + 16 * q + (int)hl < e) ? __ldg(row + s + 16 * q + hl) : 0.f;
        }
        const float4 bx = k < KEPT_SMEM ? kbox[k] : kept_box[k];
        float o0 = hl == 0 ? bx.x : hl == 1 ? bx.y : hl == 2 ? bx.z : hl == 3 ? bx.w : corner;  // out[hl]
        float o1 = 0.f;                                                                           // out[16 + hl]
#pragma unroll
        for (int g = 0; g < NGROUP; ++g) {
            constexpr int kInvalid = 1 << 20;
            const int s = group_begin(g), e = group_begin(g + 1);
            float best = -INFINITY;
            int bi = kInvalid;
#pragma unroll
            for (int q = 0; q < 3; ++q) {  // widths 31 / 24 / 37: at most three columns per lane
                if (s + 16 * q + (int)hl < e) {
                    const float x = __fmul_rn(v[g][q], obj);  // nms.py:76
                    if (x > best) { best = x; bi = 16 * q + hl; }
                }
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {  // max, ties -> lowest index (torch.max on CPU)
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if ((int)hl == 12 + g) o0 = best;            // out 12..15 = conf 0..3
            if ((int)hl == g - 4) o1 = best;             // out 16..19 = conf 4..7
            if ((int)hl == 4 + g) o1 = (float)bi;        // out 20..27 = argmax 0..7
        }
        if (do_rescale && hl < 12) {
            o0 = (hl & 1) ? rescale_coord(o0, pad_y, ratio, h0f, p.do_round)
                          : rescale_coord(o0, pad_x, ratio, w0f, p.do_round);
        }
        if (live) {
            float* dst = p.out + ((size_t)b * p.max_det + k) * OUTW;
            dst[hl] = o0;
            if (hl < 12) dst[16 + hl] = o1;
            if (p.kept_anchor != nullptr && hl == 0) p.kept_anchor[(size_t)b * p.max_det + k] = anchor;
        }
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
    unsigned n = 2 * RANK_SORT_MAX;  // the rank sort needs an input and an output buffer
    while (n < A && n < (unsigned)SORT_SMEM_KEYS) n <<= 1;
    return (int)n;
}

}  // namespace lp
