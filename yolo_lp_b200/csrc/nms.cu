// K2 -- per-image order-by-score -> exact greedy NMS -> fused gather (+ optional rescale/round).
//
// One 1024-thread CTA per image; images are independent (nms.py:68 loop body), so a batch is
// B concurrent CTAs and the stage's latency is that of the slowest image.
//
//   order   keys (score desc, anchor asc) from K1.  Ascending key order reproduces torchvision's
//           stable descending sort over the anchor-ordered compaction (nms.py:97,121); more than
//           max_nms candidates are cut to the first max_nms (nms.py:115-116).
//             N <= 64     rank sort in shared memory (rank = number of smaller keys)
//             larger      SEGMENTED: a 1024-bin histogram over the score bits splits the candidates
//                         into score-ordered segments of >= 512 keys; only the segments the greedy
//                         walk actually reaches are ordered (with max_det = 300 that is usually the
//                         first one or two).  A segment whose bins all hold <= 128 keys -- the normal
//                         case -- is a counting sort: keys are scattered to their bin's slice and
//                         ranked inside it (<= 128 comparisons per key); otherwise it is compacted and
//                         sorted by a bitonic network (block_sort: shuffles + 15 smem exchanges up to
//                         1024 keys).  A pathological histogram (one bin holding nearly everything,
//                         e.g. all scores equal) falls back to a full sort, in shared memory up to
//                         16384 keys and in the global workspace beyond.  The image's keys are cached
//                         in shared memory (N <= 12288) for the three-plus passes this mode makes.
//   NMS     torchvision.ops.nms CPU semantics (call site nms.py:121), bit-exact fp32 IoU, evaluated
//           lazily: the ordered list is walked one chunk of 32 candidates per step, warp w owning
//           member w.  Its lanes run over the boxes kept so far (stopping at the first suppressor)
//           and over the chunk's earlier members; warp 0 then settles the chunk's greedy result from
//           the 32 "suppressed by an earlier member" masks (ballot fixed point) and appends the kept
//           boxes.  IoU work is therefore only done for candidates the walk actually reaches --
//           32 x (kept so far + 16) tests per step -- and everything stops as soon as max_det rows
//           are kept, which the reference's truncation keep[:max_det] (nms.py:122-123) makes legal.
//           (The first version resolved 512-wide windows eagerly, every later candidate against every
//           alive member of the current chunk: 2-3x the IoU tests and three barriers per step;
//           cfg2 35 -> 28, cfg4 87 -> 60, cfg5 228 -> 110 thousand cycles for this phase.
//           Round 2 measured two "suppression matrix first, resolve after" designs for this phase, both
//           bit-exact and both slower than the chunk walk's 28 thousand cycles on cfg2: warp per member with
//           one ballot per 32 predecessors + a 32-members-per-step resolve on one warp, 44 thousand (every
//           (member, word) item is a ~500-cycle dependent chain of LDS -> IoU -> ballot -> store); and
//           2-8 threads per member running serial IoU loops + the greedy set as a parallel fixed point,
//           78 thousand (the lanes of a warp then test unrelated pairs and diverge through the IoU
//           shortcuts).  The walk does a third of the IoU tests of either and keeps a warp's lanes on one
//           candidate.  Two kept boxes per lane and vote, the kept rows' L2 prefetch moved off the settle
//           warp, and a two-chain re-score in the gather changed nothing measurable either (31 us); nor did
//           ONE barrier per step -- every warp settling the chunk redundantly, the previous step's kept
//           members tested straight from the window: the settle then takes 1.4 thousand cycles in each
//           of the 32 warps instead of 0.7 in warp 0 plus a barrier, 2.9 against 2.7 thousand per step.
//           A gather with one WARP per kept row -- ten coalesced loads per lane straight from the head
//           tensor, group maximum / first argmax by REDUX + ballots, no staging, no barrier -- took 14.5
//           thousand cycles where the staged thread-per-(row, group) re-score takes 10.9: ~200 warp
//           instructions per row against ~50.)
//   gather  kept rows are staged in shared memory, one thread per (row, group) recomputes
//           nms.py:76-96 (group maxima with first-index argmax), one per (row, coordinate) emits the
//           xyxy box and the corners, optionally mapped back to source coordinates
//           (inferer.py:203-228, :100).
#include <cuda_fp16.h>

#include "kernels.cuh"

namespace lp {

constexpr int NMS_THREADS = 1024;
constexpr int WIN = 512;               // candidates per window
constexpr int SORT_SMEM_KEYS = 16384;  // 128 KB sort buffer (later the row staging area)
constexpr int BLOCK_SORT_MAX = 1024;   // block_sort: input keys [0, n), output keys [BLOCK_SORT_MAX, BLOCK_SORT_MAX + n)
constexpr int RANK_SORT_N = 64;        // up to here the n^2 rank sort (one pass, three barriers); above, the histogram path
constexpr int HIST_BINS = 1024;
constexpr int SEG_TARGET = 512;        // minimum candidates per segment
constexpr int COUNTING_BIN_MAX = 128;  // a segment whose bins are all this small is ordered by counting sort
constexpr int COUNTING_SEG_MAX = 2048; // ... if it fits skeys[0, 2048) -> skeys[2048, 4096)
constexpr int KEY_CACHE_AT = 4 * BLOCK_SORT_MAX;  // segmented mode: the image's keys are cached in skeys[KEY_CACHE_AT, +N)
constexpr int KEPT_SMEM = 1024;        // kept boxes / anchors cached in shared memory (rest via L2)

__device__ __forceinline__ unsigned next_pow2(unsigned n) { return n <= 1 ? 1u : 1u << (32 - __clz(n - 1)); }

// barrier over the first `nthreads` threads of the CTA (a multiple of 32); id 1, id 0 is __syncthreads
__device__ __forceinline__ void bar_active(unsigned nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

// In-place ascending bitonic sort of n (power of two) keys; `keys` is shared or global memory.
template <bool kGlobal>
__device__ void bitonic_sort(unsigned long long* keys, unsigned n, unsigned nthreads) {
    for (unsigned k = 2; k <= n; k <<= 1) {
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned t = threadIdx.x; t < (n >> 1); t += nthreads) {
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
            bar_active(nthreads);
        }
    }
}

__device__ __forceinline__ void cp_async_4(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
// (a < b) as 0/1 through the borrow of a 64-bit subtraction: 3 integer instructions
__device__ __forceinline__ unsigned lt_u64(unsigned long long a, unsigned long long b) {
    unsigned r;
    asm("{\n\t.reg .u32 t;\n\t"
        "sub.cc.u32 t, %1, %3;\n\t"
        "subc.cc.u32 t, %2, %4;\n\t"
        "subc.u32 %0, 0, 0;\n\t}"
        : "=r"(r)
        : "r"((unsigned)a), "r"((unsigned)(a >> 32)), "r"((unsigned)b), "r"((unsigned)(b >> 32)));
    return r & 1u;  // 0 - 0 - borrow = 0xffffffff when a < b
}

// Rank sort of n <= RANK_SORT_N distinct keys (they embed the anchor) by the first `nthreads`
// threads: rank = number of smaller keys.  Keys sit in keys[0, n); the sorted list is written to
// keys[BLOCK_SORT_MAX, BLOCK_SORT_MAX + n).  The n^2 comparisons are spread over all threads: slot
// jj of replica rr counts over the rr-th slice of the keys, two per 128-bit shared-memory load, and
// the partial ranks are summed in `acc`.  Every one of the `nthreads` threads must call this.
__device__ void rank_sort(unsigned long long* keys, unsigned* acc, unsigned n, unsigned nthreads) {
    const unsigned tid = threadIdx.x;
    const unsigned wn = (n + 31u) & ~31u;            // slots, padded with ~0 (never smaller than a key)
    const unsigned rn = nthreads / wn, jj = tid % wn, rr = tid / wn;
    const bool work = rr < rn;
    for (unsigned i = n + tid; i < wn; i += nthreads) keys[i] = ~0ull;
    if (tid < wn) acc[tid] = 0;
    bar_active(nthreads);
    unsigned long long key = 0;
    if (work) {
        key = keys[jj];
        const unsigned pairs = (wn / 2 + rn - 1) / rn;
        const unsigned p0 = rr * pairs, p1 = min(p0 + pairs, wn / 2);
        const ulonglong2* kp = reinterpret_cast<const ulonglong2*>(keys);
        unsigned rank = 0;
#pragma unroll 4
        for (unsigned i = p0; i < p1; ++i) {
            const ulonglong2 kk = kp[i];
            rank += lt_u64(kk.x, key) + lt_u64(kk.y, key);
        }
        atomicAdd(&acc[jj], rank);
    }
    bar_active(nthreads);
    if (work && rr == 0 && jj < n) keys[BLOCK_SORT_MAX + acc[jj]] = key;
    bar_active(nthreads);
}

// Ascending sort of n <= BLOCK_SORT_MAX keys by the whole CTA, one key per thread: keys[0, n) ->
// keys[BLOCK_SORT_MAX, BLOCK_SORT_MAX + n).  A bitonic network whose exchanges at distance < 32 are
// warp shuffles (no barrier, no shared memory); only the 15 (of 55) steps at distance >= 32 go
// through a double-buffered exchange area keys[2 * BLOCK_SORT_MAX, 4 * BLOCK_SORT_MAX) with one
// barrier each.  (The first version was a rank sort -- n^2 comparisons -- that took 22 thousand
// cycles for a 534-key segment; this takes about four.)  Every thread of the CTA must call it.
__device__ void block_sort(unsigned long long* keys, unsigned n) {
    const unsigned i = threadIdx.x;
    const unsigned npad = n <= 32 ? 32u : next_pow2(n);
    const bool active = i < npad;   // warps beyond the padded size only keep the barriers company
    unsigned long long key = i < n ? keys[i] : ~0ull;  // pads sort to the end
    unsigned long long* xbuf = keys + 2 * BLOCK_SORT_MAX;
    unsigned cur = 0;
    for (unsigned k = 2; k <= npad; k <<= 1) {
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            unsigned long long other = key;
            if (j >= 32) {
                if (active) xbuf[cur * BLOCK_SORT_MAX + i] = key;
                __syncthreads();
                if (active) other = xbuf[cur * BLOCK_SORT_MAX + (i ^ j)];
                cur ^= 1u;  // the next exchange writes the other buffer: no second barrier needed
            } else if (active) {  // warp-uniform: npad is a multiple of 32
                other = __shfl_xor_sync(0xffffffffu, key, j);
            }
            // the lower index of a pair keeps the minimum in an ascending block, the maximum otherwise
            const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
            key = keep_min ? (other < key ? other : key) : (other > key ? other : key);
        }
    }
    if (i < n) keys[BLOCK_SORT_MAX + i] = key;
    __syncthreads();
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// xyxy box of a candidate (nms.py:79): from the head tensor, or -- fused path -- from the finished
// row KF stored for it
// (kHalf: the head tensor is stored as IEEE half, 580-byte rows, every value upcast exactly on load)
template <bool kLevels, bool kHalf>
__device__ __forceinline__ float4 candidate_box(const NmsParams& p, const float* pred, unsigned b, unsigned anchor) {
    if (!kLevels && kHalf) {
        const __half2* r = reinterpret_cast<const __half2*>(reinterpret_cast<const __half*>(pred) + (size_t)anchor * ROW);
        const float2 c = __half22float2(__ldg(r)), s = __half22float2(__ldg(r + 1));
        return xywh_to_xyxy(c.x, c.y, s.x, s.y);
    }
    if (!kLevels) {
        const float2* r = reinterpret_cast<const float2*>(pred + (size_t)anchor * ROW);
        const float2 c = __ldg(r), s = __ldg(r + 1);
        return xywh_to_xyxy(c.x, c.y, s.x, s.y);
    }
    const unsigned slot = __ldg(p.slot_of + (size_t)b * p.A + anchor);
    return __ldg(reinterpret_cast<const float4*>(p.rec + ((size_t)b * p.A + slot) * OUTW));
}

template <bool kLevels, bool kHalf>
__global__ void __launch_bounds__(NMS_THREADS, 1) nms_kernel(const NmsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* skeys = reinterpret_cast<unsigned long long*>(smem_raw);  // sort buffer, later row staging
    __shared__ float4 wbox[WIN];              // boxes of the current window
    __shared__ unsigned rank_acc[RANK_SORT_N]; // rank-sort partial ranks
    __shared__ float4 kbox[KEPT_SMEM];        // first KEPT_SMEM kept boxes
    __shared__ int kanchor[KEPT_SMEM];        // and their anchors
    __shared__ int wanchor[WIN];              // and their anchors
    __shared__ unsigned s_alive[32], s_S[32]; // per chunk member: survives the kept list / earlier members that suppress it
    __shared__ unsigned scratch[HIST_BINS];   // score histogram (inclusive scan), lives across segments
    __shared__ unsigned fill[HIST_BINS];      // counting sort: keys placed so far per bin
    __shared__ unsigned red[32];              // cross-warp reductions
    __shared__ unsigned s_misc[4];            // [0] segment end bin, [1] segment fill counter, [2] kept mask of the chunk
    __shared__ int s_nkeep;

    const unsigned b = blockIdx.x;
    const unsigned tid = threadIdx.x, lane = tid & 31;
    // image base; with kHalf `pred` points at halves, so the element offset counts 2-byte units
    const float* pred = kLevels ? nullptr
                                : kHalf ? reinterpret_cast<const float*>(reinterpret_cast<const __half*>(p.pred) + (size_t)b * p.A * ROW)
                                        : p.pred + (size_t)b * p.A * ROW;
    constexpr int ELEM = kHalf ? 2 : 4;
    float4* kept_box = p.kept_box + (size_t)b * p.max_det;
    int* kept_anchor = p.kept_anchor_ws + (size_t)b * p.max_det;

#define LP_STAMP(i) do { if (p.timing != nullptr && tid == 0) p.timing[(size_t)b * 16 + (i)] = clock64(); } while (0)
    LP_STAMP(0);
    unsigned N = (unsigned)p.counts[b];
    if (N > p.A) N = p.A;
    if (tid == 0) s_nkeep = 0;
    __syncthreads();
    if (p.rearm && tid == 0) {  // every thread has read counts[b]; the filter kernel of this workspace is done
        const_cast<int*>(p.counts)[b] = 0;
        if (b == 0) const_cast<int*>(p.counts)[gridDim.x] = 0;  // the filter kernels' tile counter
    }

    constexpr unsigned P = NMS_THREADS;   // every phase runs on the whole CTA
    unsigned long long* gkeys = p.keys + (size_t)b * p.key_stride;
    const unsigned long long* sorted = skeys;
    bool sorted_global = false, segmented = false;
    unsigned n_seg = N;         // candidates in the current (sorted) segment
    unsigned kmin = 0, shift = 0;
    // Segmented mode passes over all keys three times or more (range, histogram, one compaction per
    // segment): cache them behind the sort area when they fit (N <= 12288) instead of going to L2.
    bool key_cache = N > (unsigned)RANK_SORT_N && N + KEY_CACHE_AT <= (unsigned)p.sort_smem_keys;
    unsigned long long* ckeys = skeys + KEY_CACHE_AT;

    if (N > 0 && p.max_det > 0) {

        // ------------------------------------------------------------------ ordering mode
        if (N <= (unsigned)RANK_SORT_N) {
            for (unsigned i = tid; i < N; i += P) skeys[i] = __ldcg(gkeys + i);
            bar_active(P);
            // W slots (N rounded up to a warp) times R replicas share the comparisons
            const unsigned W = (N + 31u) & ~31u, Pr = (NMS_THREADS / W) * W;
            if (tid < Pr) rank_sort(skeys, rank_acc, N, Pr);
            __syncthreads();
            sorted = skeys + BLOCK_SORT_MAX;
        } else {
            {
                // ---- histogram of the score bits (upper key word); P == 1024 == HIST_BINS here
                unsigned lo = 0xffffffffu, hi = 0;
#pragma unroll 4
                for (unsigned i = tid; i < N; i += P) {
                    const unsigned long long key = __ldcg(gkeys + i);
                    if (key_cache) ckeys[i] = key;
                    const unsigned k = (unsigned)(key >> 32);
                    lo = min(lo, k);
                    hi = max(hi, k);
                }
                lo = __reduce_min_sync(0xffffffffu, lo);
                hi = __reduce_max_sync(0xffffffffu, hi);
                if (lane == 0) { red[tid >> 5] = lo; scratch[tid >> 5] = hi; }
                bar_active(P);
                kmin = __reduce_min_sync(0xffffffffu, red[lane]);
                const unsigned kmax = __reduce_max_sync(0xffffffffu, scratch[lane]);
                const unsigned d = kmax - kmin;
                shift = d < HIST_BINS ? 0 : (32 - __clz(d)) - 10;  // (d >> shift) < 1024, monotone in the key
                bar_active(P);
                scratch[tid] = 0;
                bar_active(P);
                for (unsigned i = tid; i < N; i += P)
                    atomicAdd(&scratch[((unsigned)((key_cache ? ckeys[i] : __ldcg(gkeys + i)) >> 32) - kmin) >> shift], 1u);
                bar_active(P);
                // inclusive scan over the 1024 bins (one per thread)
                const unsigned cnt = scratch[tid];
                unsigned inc = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
                    if ((int)lane >= o) inc += t;
                }
                if (lane == 31) red[tid >> 5] = inc;
                const unsigned big = __reduce_max_sync(0xffffffffu, cnt);
                bar_active(P);
                unsigned wsum = red[lane];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned t = __shfl_up_sync(0xffffffffu, wsum, o);
                    if ((int)lane >= o) wsum += t;
                }
                const unsigned before = __shfl_sync(0xffffffffu, wsum, (tid >> 5) ? (tid >> 5) - 1 : 0);
                inc += (tid >> 5) ? before : 0u;
                bar_active(P);
                scratch[tid] = inc;
                if (lane == 0) red[tid >> 5] = big;
                bar_active(P);
                const unsigned biggest = __reduce_max_sync(0xffffffffu, red[lane]);
                // a segment is the shortest run of bins reaching SEG_TARGET keys: it fits the sort
                // buffer unless a single bin is huge
                segmented = biggest <= (unsigned)(p.sort_smem_keys - SEG_TARGET - BLOCK_SORT_MAX);
                // a segment longer than the sort area in front of the cache would overwrite it
                key_cache = key_cache && segmented && biggest + SEG_TARGET <= (unsigned)KEY_CACHE_AT;
            }
            if (!segmented) {
                const unsigned npad = next_pow2(N);
                if (N <= (unsigned)BLOCK_SORT_MAX) {
                    for (unsigned i = tid; i < N; i += P) skeys[i] = __ldcg(gkeys + i);
                    bar_active(P);
                    block_sort(skeys, N);
                    sorted = skeys + BLOCK_SORT_MAX;
                } else if (npad <= (unsigned)p.sort_smem_keys) {
                    for (unsigned i = tid; i < npad; i += P) skeys[i] = i < N ? __ldcg(gkeys + i) : ~0ull;
                    bar_active(P);
                    bitonic_sort<false>(skeys, npad, P);
                } else {  // key_stride >= npad is guaranteed by lp_nms_workspace_bytes
                    for (unsigned i = N + tid; i < npad; i += P) __stcg(gkeys + i, ~0ull);
                    bar_active(P);
                    bitonic_sort<true>(gkeys, npad, P);
                    sorted = gkeys;
                    sorted_global = true;
                }
            }
        }
        bar_active(P);
        LP_STAMP(2);  // ordered (or histogram ready)

    }
    __syncthreads();

    // ---------------------------------------------------------------------- NMS, all 32 warps
    // The ordered list is walked 32 candidates (one chunk) per step, warp w owning member w:
    //   test     lanes run over the boxes kept so far (kbox, 16-byte conflict-free LDS) plus the
    //            chunk's earlier members: `alive` = no kept box suppresses the member, S = mask of
    //            earlier chunk members that would (if they end up kept);
    //   settle   warp 0, lane i = member i: K = the greedy result inside the chunk (ballot fixed
    //            point over the S masks), cut to the room left under max_det; kept members append
    //            their box / anchor to the kept list.
    // Two CTA barriers per step; work is only ever done for candidates the walk actually reaches and
    // stops as soon as max_det rows are kept (nms.py:122-123 makes that legal).  Candidate boxes are
    // gathered WIN at a time into shared memory.
    const unsigned warp = tid >> 5;
    const unsigned lower = (1u << lane) - 1u;
    const float iou_floor = p.iou_floor;
    int n_keep = 0;
    unsigned consumed = 0;   // candidates of earlier segments
    unsigned seg_bin0 = 0;   // first histogram bin of the next segment
    bool first_window = true;
#ifdef LP_NMS_PROFILE
    long long pf_hdr = 0, pf_iou = 0, pf_b0 = 0, pf_fix = 0, pf_b2 = 0, pf_n = 0, pf_t = clock64();
    // only the last warp reads the clock: 32 warps doing so right after a barrier serialise on CS2R
#define LP_PF(acc) do { if (warp == 31) { const long long t1_ = clock64(); acc += t1_ - pf_t; pf_t = t1_; } } while (0)
#else
#define LP_PF(acc) do { } while (0)
#endif
    while (p.max_det > 0 && consumed < N && consumed < (unsigned)p.max_nms && n_keep < p.max_det) {
        if (segmented) {
            // ---- next segment: bins [seg_bin0, e] with e the first bin reaching SEG_TARGET keys
            const unsigned base_cnt = seg_bin0 ? scratch[seg_bin0 - 1] : 0u;
            if (tid == 0) { s_misc[0] = HIST_BINS - 1; s_misc[1] = 0; }
            bar_active(P);
            if (tid >= seg_bin0 && scratch[tid] - base_cnt >= (unsigned)SEG_TARGET) atomicMin(&s_misc[0], tid);
            bar_active(P);
            const unsigned e = s_misc[0];
            n_seg = scratch[e] - base_cnt;
            // largest bin of the segment
            if (tid == 0) s_misc[3] = 0;
            fill[tid] = 0;
            bar_active(P);
            if (tid >= seg_bin0 && tid <= e) atomicMax(&s_misc[3], scratch[tid] - (tid ? scratch[tid - 1] : 0u));
            bar_active(P);
            const bool counting = s_misc[3] <= (unsigned)COUNTING_BIN_MAX && n_seg <= (unsigned)COUNTING_SEG_MAX;
            // Compaction.  Counting mode: the histogram already orders the keys by bin, so a key goes
            // straight to its bin's slice (slot within the bin from an atomic counter, arbitrary);
            // otherwise warp-aggregated appends in arbitrary order.
            for (unsigned i0 = 0; i0 < N; i0 += P) {
                const unsigned i = i0 + tid;
                unsigned long long key = 0;
                bool take = false;
                unsigned bin = 0;
                if (i < N) {
                    key = key_cache ? ckeys[i] : __ldcg(gkeys + i);
                    bin = ((unsigned)(key >> 32) - kmin) >> shift;
                    take = bin >= seg_bin0 && bin <= e;
                }
                if (counting) {
                    if (take) skeys[(bin ? scratch[bin - 1] : 0u) - base_cnt + atomicAdd(&fill[bin], 1u)] = key;
                } else {
                    const unsigned m = __ballot_sync(0xffffffffu, take);
                    unsigned base = 0;
                    if (lane == 0 && m) base = atomicAdd(&s_misc[1], (unsigned)__popc(m));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (take) skeys[base + __popc(m & lower)] = key;
                }
            }
            bar_active(P);
            if (consumed == 0) LP_STAMP(1);  // first segment compacted
            if (counting) {
                // rank inside the bin's slice: at most COUNTING_BIN_MAX comparisons per key instead of a
                // sorting network over the whole segment (which took 15-20 thousand cycles for ~500 keys)
                unsigned long long* out = skeys + COUNTING_SEG_MAX;
                for (unsigned q = tid; q < n_seg; q += P) {
                    const unsigned long long key = skeys[q];
                    const unsigned bin = ((unsigned)(key >> 32) - kmin) >> shift;
                    const unsigned s0 = (bin ? scratch[bin - 1] : 0u) - base_cnt, s1 = scratch[bin] - base_cnt;
                    unsigned rank = 0;
                    for (unsigned r = s0; r < s1; ++r) rank += lt_u64(skeys[r], key);
                    out[s0 + rank] = key;
                }
                bar_active(P);
                sorted = out;
            } else if (n_seg <= (unsigned)BLOCK_SORT_MAX) {
                block_sort(skeys, n_seg);
                sorted = skeys + BLOCK_SORT_MAX;
            } else {
                const unsigned npad = next_pow2(n_seg);
                for (unsigned i = n_seg + tid; i < npad; i += P) skeys[i] = ~0ull;
                bar_active(P);
                bitonic_sort<false>(skeys, npad, P);
                sorted = skeys;
            }
            seg_bin0 = e + 1;
            if (consumed == 0) { LP_STAMP(7); if (p.timing != nullptr && tid == 0) p.timing[(size_t)b * 16 + 14] = n_seg; }
        }
        const unsigned n_use = min(n_seg, (unsigned)p.max_nms - consumed);  // nms.py:115-116


        for (unsigned w0 = 0; w0 < n_use && n_keep < p.max_det; w0 += WIN) {
            const unsigned n_win = min((unsigned)WIN, n_use - w0);
            if (tid < n_win) {
                const unsigned pos = w0 + tid;
                const unsigned anchor = (unsigned)(sorted_global ? __ldcg(sorted + pos) : sorted[pos]);
                wbox[tid] = candidate_box<kLevels, kHalf>(p, pred, b, anchor);  // nms.py:79
                wanchor[tid] = (int)anchor;
            }
            __syncthreads();
            if (first_window) { LP_STAMP(3); first_window = false; }  // first window loaded
            LP_PF(pf_hdr);
            for (unsigned c0 = 0; c0 < n_win && n_keep < p.max_det; c0 += 32) {
                // ---- test: warp w <-> member c0 + w
                const unsigned idx = c0 + warp;
                if (idx < n_win) {  // warp-uniform
                    const float4 box = wbox[idx];
                    const float area = box_area(box);
                    // kept boxes in shared memory first (all of them unless max_det > KEPT_SMEM), in a loop
                    // without the global fall-back's address arithmetic; stops at the first suppressor
                    bool dead = false;
                    const int n_s = min(n_keep, KEPT_SMEM);
                    const float4* kb_lane = kbox + lane;
                    for (int k0 = 0; k0 < n_s && !dead; k0 += 32) {
                        bool hit = false;
                        if (k0 + (int)lane < n_s) {
                            const float4 kb = kb_lane[k0];
                            hit = iou_exceeds(kb, box_area(kb), box, area, iou_floor);
                        }
                        dead = __any_sync(0xffffffffu, hit);
                    }
                    for (int k0 = KEPT_SMEM; k0 < n_keep && !dead; k0 += 32) {
                        bool hit = false;
                        if (k0 + (int)lane < n_keep) {
                            const float4 kb = __ldcg(kept_box + k0 + lane);
                            hit = iou_exceeds(kb, box_area(kb), box, area, iou_floor);
                        }
                        dead = __any_sync(0xffffffffu, hit);
                    }
                    unsigned S = 0;
                    if (!dead) {  // earlier members of the chunk (all of them exist: c0 + lane < idx)
                        bool over = false;
                        if (lane < warp) {
                            const float4 mb = wbox[c0 + lane];
                            over = iou_exceeds(mb, box_area(mb), box, area, iou_floor);
                        }
                        S = __ballot_sync(0xffffffffu, over);
                    }
                    if (lane == 0) { s_alive[warp] = dead ? 0u : 1u; s_S[warp] = S; }
                } else if (lane == 0) {
                    s_alive[warp] = 0u;
                    s_S[warp] = 0u;
                }
                LP_PF(pf_iou);
                __syncthreads();
                LP_PF(pf_b0);
                // ---- settle: K_i = alive_i and no kept earlier member suppresses i.  K <- F(K) fixes at
                // least one more leading member per round; the unique fixed point is the greedy result.
                if (warp == 0) {
                    const bool in_a = s_alive[lane] != 0u;
                    const unsigned sc = s_S[lane];
                    unsigned K = __ballot_sync(0xffffffffu, in_a);
                    while (true) {
                        const unsigned K2 = __ballot_sync(0xffffffffu, in_a && (sc & K) == 0);
                        if (K2 == K) break;
                        K = K2;
                    }
                    const int room = p.max_det - n_keep;
                    if (__popc(K) > room) K &= (1u << __fns(K, 0, room + 1)) - 1u;  // first `room` members only
                    if ((K >> lane) & 1u) {
                        const int k = n_keep + __popc(K & lower);
                        const float4 box = wbox[c0 + lane];
                        const int anchor = wanchor[c0 + lane];
                        if (k < KEPT_SMEM) { kbox[k] = box; kanchor[k] = anchor; }
                        else { kept_box[k] = box; kept_anchor[k] = anchor; }
                        // the gather will want this row: start pulling it into L2 now (K1 streamed the
                        // head tensor with an evict-first policy, so it is most likely back in HBM)
                        if (!kLevels) {
                            const char* row = reinterpret_cast<const char*>(pred) + (size_t)anchor * ROW * ELEM;
#pragma unroll
                            for (int o = 0; o < ROW * ELEM + 127; o += 128) prefetch_l2(row + o);
                        }
                    }
                    if (lane == 0) s_misc[2] = K;
                }
                LP_PF(pf_fix);
                __syncthreads();
                n_keep += __popc(s_misc[2]);
                LP_PF(pf_b2);
#ifdef LP_NMS_PROFILE
                ++pf_n;
#endif
            }
        }
        consumed += n_seg;
        if (!segmented) break;
    }
#ifdef LP_NMS_PROFILE
    if (p.timing != nullptr && tid == NMS_THREADS - 1) {  // last warp
        long long* q = p.timing + (size_t)b * 16 + 8;
        q[0] = pf_hdr; q[1] = pf_iou; q[2] = pf_b0; q[3] = pf_fix; q[4] = pf_b2; q[5] = pf_n;
    }
#endif
#undef LP_PF
    if (tid == 0) s_nkeep = n_keep;
    LP_STAMP(4);  // NMS done
    __syncthreads();

    // ---------------------------------------------------------------------- gather
    // Kept rows are staged in shared memory (cp.async, 8-byte granules: rows are only 8-byte
    // aligned), then one thread per (row, group) scans its <= 37 class scores in order -- strict
    // '>' keeps the first maximum like torch.max on CPU -- and one thread per (row, coordinate)
    // emits the box / corner columns.
    n_keep = s_nkeep;
    if (tid == 0) p.out_counts[b] = n_keep;
    // Staged rows: pitch 1168 B (16-byte multiple); a row starts 8 bytes into its slot when its
    // index in the head tensor (b * A + anchor) is odd, so that source (row index * 1160 B from a
    // 16-byte aligned base: 16-byte aligned only for even rows) and destination agree mod 16 and all
    // but 8 of the 1160 bytes move as 16-byte cp.async copies.
    // (kHalf: 580-byte rows, 4-byte aligned; slots of 584 B filled with 4-byte copies.)
    constexpr int SROW_BYTES = kHalf ? 584 : 1168;
    unsigned char* sbytes = smem_raw;
    const int cap_rows = (int)(((size_t)p.sort_smem_keys * sizeof(unsigned long long)) / SROW_BYTES);
    const int img_parity = (int)(((size_t)b * p.A) & 1);
#define LP_SROW(r, anchor) (sbytes + (size_t)(r) * SROW_BYTES + (kHalf ? 0 : 8 * (((anchor) + img_parity) & 1)))
    // column c of a staged row, upcast exactly when the tensor is half
    auto col = [](const unsigned char* row, int c) -> float {
        return kHalf ? __half2float(reinterpret_cast<const __half*>(row)[c]) : reinterpret_cast<const float*>(row)[c];
    };
    const bool do_rescale = p.rescale != nullptr;
    float pad_x = 0.f, pad_y = 0.f, ratio = 1.f, w0f = 0.f, h0f = 0.f;
    if (do_rescale) {
        const float* rp = p.rescale + (size_t)b * 5;
        pad_x = rp[0]; pad_y = rp[1]; ratio = rp[2]; w0f = rp[3]; h0f = rp[4];
    }
    for (int base = 0; base < n_keep; base += cap_rows) {
        const int nb = min(cap_rows, n_keep - base);
        if (!kLevels) {
            for (int r = (int)(tid >> 5); r < nb; r += NMS_THREADS / 32) {  // one warp per row
                const int k = base + r;
                const int anchor = k < KEPT_SMEM ? kanchor[k] : kept_anchor[k];
                const unsigned char* src = reinterpret_cast<const unsigned char*>(pred) + (size_t)anchor * ROW * ELEM;
                if (kHalf) {
                    unsigned char* dst = sbytes + (size_t)r * SROW_BYTES;
                    for (int c = (int)lane; c < ROW / 2; c += 32) cp_async_4(dst + 4 * c, src + 4 * c);
                    continue;
                }
                const int odd = (anchor + img_parity) & 1;
                unsigned char* dst = sbytes + (size_t)r * SROW_BYTES + 8 * odd;
                // 72 16-byte chunks from the first 16-byte boundary of the row, 8 bytes before (odd) or after (even)
                for (int c = (int)lane; c < 72; c += 32) cp_async_16(dst + 8 * odd + 16 * c, src + 8 * odd + 16 * c);
                if (lane == 0) cp_async_8(dst + (odd ? 0 : 1152), src + (odd ? 0 : 1152));
            }
            cp_async_wait_all();
            __syncthreads();
        } else {
            // fused path: KF already finished the row of every candidate -- copy the kept ones
            for (int i = tid; i < nb * OUTW; i += NMS_THREADS) {
                const int r = i / OUTW, c = i - r * OUTW;
                const int k = base + r;
                const int anchor = k < KEPT_SMEM ? kanchor[k] : kept_anchor[k];
                const unsigned slot = __ldg(p.slot_of + (size_t)b * p.A + anchor);
                float val = __ldg(p.rec + ((size_t)b * p.A + slot) * OUTW + c);
                if (do_rescale && c < 12)
                    val = (c & 1) ? rescale_coord(val, pad_y, ratio, h0f, p.do_round) : rescale_coord(val, pad_x, ratio, w0f, p.do_round);
                p.out[((size_t)b * p.max_det + k) * OUTW + c] = val;
                if (c == 0 && p.kept_anchor != nullptr) p.kept_anchor[(size_t)b * p.max_det + k] = anchor;
            }
            continue;
        }
        if (base == 0) LP_STAMP(6);  // rows staged
        for (int t = tid; t < nb * NGROUP; t += NMS_THREADS) {
            const int r = t >> 3, g = t & 7;
            const int ka = base + r < KEPT_SMEM ? kanchor[base + r] : kept_anchor[base + r];
            const unsigned char* row = LP_SROW(r, ka);
            const float obj = col(row, 4);
            const int s = group_begin(g), e = group_begin(g + 1);
            float best = __fmul_rn(col(row, s), obj);  // nms.py:76
            int bi = 0;
            // The SM is issue-bound here (32 warps, ~280 element visits each): keep the visit short.
            // Columns past the group's end are replaced by the group's first value -- never '>' the
            // running maximum -- so the unrolled body needs no bounds predicate; the maximum itself is
            // an FMNMX and only the index hangs on the compare.  (x > best with best = max so far is
            // exactly torch.max's first-occurrence rule; NaN scores are outside the head's range.)
            const float first = best;
            for (int i0 = s + 1; i0 < e; i0 += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + u;
                    const float v = i < e ? __fmul_rn(col(row, i), obj) : first;
                    bi = v > best ? i - s : bi;
                    best = fmaxf(best, v);
                }
            }
            float* dst = p.out + ((size_t)b * p.max_det + base + r) * OUTW;
            dst[12 + g] = best;
            dst[20 + g] = (float)bi;
        }
        if (base == 0) LP_STAMP(15);  // first batch: groups re-scored
        for (int t = tid; t < nb * 12; t += NMS_THREADS) {
            const int r = t / 12, c = t - r * 12;
            const int k = base + r;
            float val;
            if (c < 4) {
                const float4 bx = k < KEPT_SMEM ? kbox[k] : kept_box[k];
                val = c == 0 ? bx.x : c == 1 ? bx.y : c == 2 ? bx.z : bx.w;
            } else {
                const int ka = k < KEPT_SMEM ? kanchor[k] : kept_anchor[k];
                val = col(LP_SROW(r, ka), c + 1);  // corners: columns 5..12 -> output 4..11 (nms.py:94)
            }
            if (do_rescale)
                val = (c & 1) ? rescale_coord(val, pad_y, ratio, h0f, p.do_round) : rescale_coord(val, pad_x, ratio, w0f, p.do_round);
            p.out[((size_t)b * p.max_det + k) * OUTW + c] = val;
        }
        if (p.kept_anchor != nullptr)
            for (int t = tid; t < nb; t += NMS_THREADS) {
                const int k = base + t;
                p.kept_anchor[(size_t)b * p.max_det + k] = k < KEPT_SMEM ? kanchor[k] : kept_anchor[k];
            }
        __syncthreads();
    }
    LP_STAMP(5);  // gather done
#undef LP_SROW
#undef LP_STAMP
}

cudaError_t launch_nms(const NmsParams& p, int B, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    const size_t smem = (size_t)p.sort_smem_keys * sizeof(unsigned long long);
    static bool configured[64] = {false};  // per device; idempotent, so a race only repeats the calls
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(nms_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(nms_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(nms_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    if (p.from_levels) nms_kernel<true, false><<<B, NMS_THREADS, smem, stream>>>(p);
    else if (p.half_input) nms_kernel<false, true><<<B, NMS_THREADS, smem, stream>>>(p);
    else nms_kernel<false, false><<<B, NMS_THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

int nms_sort_smem_keys(unsigned A) {
    (void)A;  // always the full buffer: it doubles as the row staging area of the gather
    return SORT_SMEM_KEYS;
}

}  // namespace lp
