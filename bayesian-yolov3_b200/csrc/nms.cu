// K3: class-agnostic greedy NMS + gather, one CTA per image, everything in shared memory.
// Replaces tf.image.non_max_suppression(boxes[:, :4], boxes[:, obj_idx], 1000) + tf.gather of
// /root/reference/inference_epistemic.py:99-128, inference_aleatoric.py:104-145, inference_standard_yolov3.py:104-145
// (and the tf.while_loop over the batch, :137-143: here the batch is the grid and results are padded to max_out).
//
// Semantics (bit-exact with oracle/nms_ref.c): candidates ordered by (score desc, index asc); a candidate is kept
// iff IoU(candidate, s) <= thr for every already kept s; stop at max_out.  IoU in fp32 with every operation rounded
// separately (__f*_rn, no FMA contraction): corners re-ordered by min/max, area (y2-y1)*(x2-x1), 0 if an area <= 0.
//
// Phases:  0. (N > 4096) 65536-bin histogram of the upper key bits picks a score cut that keeps 3072..4096 candidates;
//             only those are sorted; if the scan runs out of them before max_out boxes are kept, the kernel redoes
//             the image over all candidates (exact either way);
//          1. keys -> smem as (ordered score bits : u32, index : u16), bitonic sort of the padded power of two;
//          2. batches of 512 candidates in sorted order: (A) test against the boxes kept so far, (B) 512x512
//             suppression bit-matrix inside the batch, (C) one warp resolves the batch sequentially with ballots;
//          3. gather the kept rows, zero-fill the tail, write indices and count.
// Because "any earlier kept box overlaps" is order independent, the parallel evaluation selects exactly the boxes
// the sequential algorithm selects.
#include <algorithm>

#include "common.cuh"

namespace byolo {

constexpr int kNmsThreads = 1024;
constexpr int kBatch = 512;
constexpr int kWords = kBatch / 32;
constexpr int kMaxN = 32768;
constexpr int kMaxOut = 2048;

__device__ __forceinline__ uint32_t ordered_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct BoxA { float ymin, xmin, ymax, xmax, area; };
constexpr size_t kAuxBytes = sizeof(BoxA) * (kBatch + kMaxOut) + 4u * (kBatch * kWords + kWords) + 4u * kMaxOut;

__device__ __forceinline__ BoxA load_box(const float* r) {
    BoxA b;
    b.ymin = fminf(r[0], r[2]);
    b.xmin = fminf(r[1], r[3]);
    b.ymax = fmaxf(r[0], r[2]);
    b.xmax = fmaxf(r[1], r[3]);
    b.area = __fmul_rn(__fsub_rn(b.ymax, b.ymin), __fsub_rn(b.xmax, b.xmin));
    return b;
}

__device__ __forceinline__ bool iou_gt(const BoxA& a, const BoxA& b, float thr) {
    if (a.area <= 0.f || b.area <= 0.f) return false;
    const float ih = fmaxf(__fsub_rn(fminf(a.ymax, b.ymax), fmaxf(a.ymin, b.ymin)), 0.f);
    const float iw = fmaxf(__fsub_rn(fminf(a.xmax, b.xmax), fmaxf(a.xmin, b.xmin)), 0.f);
    const float inter = __fmul_rn(ih, iw);
    if (inter == 0.f) return false;                       // IoU == 0 exactly; skips the division for disjoint boxes
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(a.area, b.area), inter)) > thr;
}

constexpr int kTopK = 4096;          // candidates sorted on the fast path
constexpr int kTopTarget = 3072;     // the score cut keeps at least this many (if N allows)
constexpr int kBins = 65536;         // histogram over the upper 16 bits of the ordered score key

// Bitonic sort of (key_hi desc, key_lo asc) over NP (power of two) entries.
__device__ void bitonic_sort(uint32_t* key_hi, uint16_t* key_lo, int NP) {
    const int tid = threadIdx.x;
    for (int k = 2; k <= NP; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int pidx = tid; pidx < (NP >> 1); pidx += kNmsThreads) {
                const int i = ((pidx & ~(j - 1)) << 1) | (pidx & (j - 1));
                const int l = i | j;
                const uint32_t hi_i = key_hi[i], hi_l = key_hi[l];
                const uint16_t lo_i = key_lo[i], lo_l = key_lo[l];
                const bool i_first = (hi_i > hi_l) || (hi_i == hi_l && lo_i < lo_l);   // i belongs before l in final order
                const bool up = (i & k) == 0;                                          // this block sorts into final order
                if (i_first != up) {
                    key_hi[i] = hi_l; key_hi[l] = hi_i;
                    key_lo[i] = lo_l; key_lo[l] = lo_i;
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(kNmsThreads, 1)
nms_kernel(const float* __restrict__ rows_all, int N, int D, int obj_idx, float thr, int max_out, int NP,
           size_t region0, float* __restrict__ out_rows, int* __restrict__ out_idx, int* __restrict__ out_count) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* rows = rows_all + (size_t)blockIdx.x * N * D;
    // region 0: score histogram (top-K selection), then score keys during the sort, then the scan-phase scratch;
    // region 1: the candidate indices, which the sort leaves in selection-priority order.
    uint32_t* key_hi = reinterpret_cast<uint32_t*>(sm);                 // [NP]
    uint16_t* key_lo = reinterpret_cast<uint16_t*>(sm + region0);       // [NP]
    uint16_t* hist = reinterpret_cast<uint16_t*>(sm);                   // [kBins] (N <= 32768 < 65536 fits u16)
    uint8_t* aux = sm;
    BoxA* cand = reinterpret_cast<BoxA*>(aux);                          // [kBatch]
    BoxA* kept = cand + kBatch;                                         // [max_out]
    uint32_t* mask = reinterpret_cast<uint32_t*>(kept + kMaxOut);       // [kBatch][kWords]
    uint32_t* dead = mask + kBatch * kWords;                            // [kWords]
    int* kept_idx = reinterpret_cast<int*>(dead + kWords);              // [max_out]
    __shared__ int s_kept, s_cut_bin, s_ncand, s_fill;
    __shared__ int s_warp_sum[32];

    // ---- 0. fast path: a score cut that keeps ~kTopTarget..kTopK candidates (exact: the cut is a key prefix) ----
    bool full = (N <= kTopK);
    if (!full) {
        for (int i = tid; i < kBins / 2; i += kNmsThreads) reinterpret_cast<uint32_t*>(hist)[i] = 0u;
        if (tid == 0) { s_cut_bin = -1; s_ncand = 0; s_fill = 0; }
        __syncthreads();
        for (int i = tid; i < N; i += kNmsThreads) {
            const uint32_t bin = ordered_key(rows[(size_t)i * D + obj_idx]) >> 16;
            // 16-bit counters packed in 32-bit words: add into the right half-word (no overflow: counts <= N < 65536)
            atomicAdd(reinterpret_cast<uint32_t*>(hist) + (bin >> 1), (bin & 1) ? 0x10000u : 1u);
        }
        __syncthreads();
        // thread t owns bins [64t, 64t+64); suffix sums from the top bin downwards
        constexpr int per = kBins / kNmsThreads;
        int local = 0;
        for (int b = 0; b < per; ++b) local += hist[tid * per + b];
        int incl = local;                                  // inclusive suffix scan over threads (higher tid = higher keys)
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_down_sync(0xFFFFFFFFu, incl, o);
            if (lane + o < 32) incl += v;
        }
        if (lane == 0) s_warp_sum[warp] = incl;
        __syncthreads();
        int above = 0;                                     // candidates in warps above mine
        for (int w = warp + 1; w < 32; ++w) above += s_warp_sum[w];
        const int above_me = above + incl - local;         // candidates in bins above my range
        if (above_me < kTopTarget && above_me + local >= kTopTarget) {
            int cum = above_me;
            for (int b = per - 1; b >= 0; --b) {
                cum += hist[tid * per + b];
                if (cum >= kTopTarget) { s_cut_bin = tid * per + b; s_ncand = cum; break; }
            }
        }
        __syncthreads();
        if (s_cut_bin < 0 || s_ncand > kTopK) full = true;  // fewer than the target in total, or one huge tie bin
    }

    for (int pass = 0; pass < 2; ++pass) {
        // ---- 1. keys + bitonic sort (descending score, ascending index) ----
        int n_cand, NPs;
        if (full) {
            n_cand = N;
            NPs = NP;
            for (int i = tid; i < NP; i += kNmsThreads) {
                key_hi[i] = (i < N) ? ordered_key(rows[(size_t)i * D + obj_idx]) : 0u;
                key_lo[i] = (i < N) ? (uint16_t)i : (uint16_t)0xFFFF;
            }
        } else {
            n_cand = s_ncand;
            NPs = kTopK;
            const uint32_t cut = (uint32_t)s_cut_bin;
            __syncthreads();                               // everyone has read the histogram results: region 0 is free
            for (int i = tid; i < kTopK; i += kNmsThreads) { key_hi[i] = 0u; key_lo[i] = (uint16_t)0xFFFF; }
            __syncthreads();
            for (int i = tid; i < N; i += kNmsThreads) {
                const uint32_t key = ordered_key(rows[(size_t)i * D + obj_idx]);
                if ((key >> 16) >= cut) {
                    const int pos = atomicAdd(&s_fill, 1);
                    key_hi[pos] = key;
                    key_lo[pos] = (uint16_t)i;
                }
            }
        }
        if (tid == 0) s_kept = 0;
        __syncthreads();
        bitonic_sort(key_hi, key_lo, NPs);

        // ---- 2. batched greedy scan ----
        for (int base = 0; base < n_cand; base += kBatch) {
            const int kept_before = s_kept;
            if (kept_before >= max_out) break;
            const int nb = min(kBatch, n_cand - base);
            if (tid < kBatch) {
                BoxA b;
                b.ymin = b.xmin = b.ymax = b.xmax = 0.f;
                b.area = -1.f;
                if (tid < nb) b = load_box(rows + (size_t)key_lo[base + tid] * D);
                cand[tid] = b;
            }
            if (tid < kWords) dead[tid] = 0u;
            __syncthreads();
            {   // (A) two threads per candidate, each scanning half of the kept list
                const int c = tid & (kBatch - 1), part = tid >> 9;
                const int half = (kept_before + 1) >> 1;
                const int j0 = part * half, j1 = min(kept_before, j0 + half);
                const BoxA me = cand[c];
                bool hit = false;
                if (c < nb)
                    for (int j = j0; j < j1; ++j)
                        if (iou_gt(me, kept[j], thr)) { hit = true; break; }
                if (hit || c >= nb) atomicOr(&dead[c >> 5], 1u << (c & 31));
            }
            __syncthreads();
            // (B) mask[k][w] bit i: candidate (32w+i) > k is suppressed by candidate k
            for (int wid = tid; wid < kBatch * kWords; wid += kNmsThreads) {
                const int k = wid / kWords, w = wid - k * kWords;
                uint32_t bits = 0u;
                if (w * 32 + 31 > k && !((dead[k >> 5] >> (k & 31)) & 1u)) {
                    const BoxA bk = cand[k];
                    const uint32_t dw = dead[w];
                    for (int i = 0; i < 32; ++i) {
                        const int c = w * 32 + i;
                        if (c > k && !((dw >> i) & 1u) && iou_gt(cand[c], bk, thr)) bits |= 1u << i;
                    }
                }
                mask[wid] = bits;
            }
            __syncthreads();
            // (C) sequential resolution by one warp: lane l < kWords owns candidates [32l, 32l+32)
            if (warp == 0) {
                uint32_t alive = (lane < kWords) ? ~dead[lane] : 0u;
                uint32_t removed = 0u;
                int n_new = 0;
                while (true) {
                    const uint32_t cur = alive & ~removed;
                    const uint32_t vote = __ballot_sync(0xFFFFFFFFu, cur != 0u);
                    if (!vote) break;
                    const int src = __ffs(vote) - 1;
                    const uint32_t wv = __shfl_sync(0xFFFFFFFFu, cur, src);
                    const int bit = __ffs(wv) - 1;
                    const int k = src * 32 + bit;
                    if (lane == 0) {
                        kept[kept_before + n_new] = cand[k];
                        kept_idx[kept_before + n_new] = key_lo[base + k];
                    }
                    ++n_new;
                    if (kept_before + n_new >= max_out) break;
                    if (lane < kWords) removed |= mask[k * kWords + lane];
                    if (lane == src) alive &= ~(1u << bit);
                }
                if (lane == 0) s_kept = kept_before + n_new;
            }
            __syncthreads();
        }
        // the cut was sufficient iff the cap was reached or nothing was left below it
        if (full || s_kept >= max_out || n_cand == N) break;
        __syncthreads();
        full = true;                                       // rare: redo exactly over all candidates
    }

    // ---- 3. gather ----
    const int n_kept = s_kept;
    float* orow = out_rows + (size_t)blockIdx.x * max_out * D;
    for (int e = tid; e < max_out * D; e += kNmsThreads) {
        const int k = e / D, c = e - k * D;
        orow[e] = (k < n_kept) ? rows[(size_t)kept_idx[k] * D + c] : 0.f;
    }
    if (out_idx)
        for (int k = tid; k < max_out; k += kNmsThreads) out_idx[(size_t)blockIdx.x * max_out + k] = (k < n_kept) ? kept_idx[k] : -1;
    if (tid == 0 && out_count) out_count[blockIdx.x] = n_kept;
}

size_t nms_workspace_bytes(int, int) { return 0; }

int launch_nms(const float* rows, int B, int N, int D, int obj_idx, float iou_thr, int max_out, float* out_rows, int* out_idx,
               int* out_count, void*, size_t, cudaStream_t st) {
    BY_REQUIRE(N >= 0 && N <= kMaxN, "NMS kernel handles up to 32768 candidates per image");
    BY_REQUIRE(max_out >= 1 && max_out <= kMaxOut, "max_out must be in [1, 2048]");
    BY_REQUIRE(obj_idx >= 4 && obj_idx < D, "obj_idx out of range");
    if (B == 0) return 0;
    int NP = 2;
    while (NP < N) NP <<= 1;
    const size_t region0 = std::max({(size_t)NP * 4, kAuxBytes, N > 4096 ? (size_t)65536 * 2 : (size_t)0});
    const size_t smem = region0 + (size_t)NP * 2;
    static bool attr_done = false;
    if (!attr_done) {
        BY_CUDA(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
        attr_done = true;
    }
    BY_REQUIRE(smem <= 227 * 1024 - 1024, "NMS shared memory budget exceeded");
    nms_kernel<<<B, kNmsThreads, smem, st>>>(rows, N, D, obj_idx, iou_thr, max_out, NP, region0, out_rows, out_idx, out_count);
    BY_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace byolo
