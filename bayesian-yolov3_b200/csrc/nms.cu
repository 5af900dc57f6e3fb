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
//             suppression bit-matrix inside the batch, one ballot per 32 pairs, (C) candidates without a possible
//             suppressor or victim inside the batch are kept outright, one warp walks the others sequentially;
//          3. gather the kept rows, zero-fill the tail, write indices and count.
// Because "any earlier kept box overlaps" is order independent, the parallel evaluation selects exactly the boxes
// the sequential algorithm selects.
//
// A batch of B images would occupy only B of the 148 SMs, so an image is handled by a thread-block CLUSTER of CS CTAs
// (CS = 1, 2, 4 or 8): every CTA keeps the full state (sorted keys, kept list) and runs the cheap phases redundantly,
// while the pair tests - phase (A) over the kept list and the rows of the bit matrix in (B), >half of the time at CS = 1 -
// are split across the cluster and exchanged through distributed shared memory (two cluster barriers per batch).
#include <algorithm>
#include <array>
#include <map>
#include <mutex>
#include <tuple>

#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

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

// Boxes are kept as (ymin, xmin, ymax, xmax) + area.  A box whose area is <= 0 never overlaps anything in TF's IoU
// (it returns 0 before looking at the other box); it is stored as (+inf, +inf, -inf, -inf) so that the first
// interval test of iou_gt rejects it without a separate area check.
struct __align__(16) Box4 { float ymin, xmin, ymax, xmax; };
constexpr size_t kAuxBytes = (sizeof(Box4) + 4) * (kBatch + kMaxOut) + 4u * (kBatch * kWords + 9 * kWords) + 4u * kMaxOut;

__device__ __forceinline__ void load_box(const float* r, Box4* b, float* area) {
    Box4 t;
    t.ymin = fminf(r[0], r[2]);
    t.xmin = fminf(r[1], r[3]);
    t.ymax = fmaxf(r[0], r[2]);
    t.xmax = fmaxf(r[1], r[3]);
    const float a = __fmul_rn(__fsub_rn(t.ymax, t.ymin), __fsub_rn(t.xmax, t.xmin));
    if (!(a > 0.f)) {
        t.ymin = t.xmin = __int_as_float(0x7f800000);
        t.ymax = t.xmax = __int_as_float(0xff800000);
    }
    *b = t;
    *area = a;
}

// IoU(a, b) > thr (thr >= 0) with TF's fp32 operation order.  Disjoint boxes (the common case) leave after one or two
// interval tests: their intersection is 0, hence IoU = 0 <= thr.
__device__ __forceinline__ bool iou_gt(const Box4& a, float area_a, const Box4& b, float area_b, float thr) {
    const float ih = __fsub_rn(fminf(a.ymax, b.ymax), fmaxf(a.ymin, b.ymin));
    if (!(ih > 0.f)) return false;
    const float iw = __fsub_rn(fminf(a.xmax, b.xmax), fmaxf(a.xmin, b.xmin));
    if (!(iw > 0.f)) return false;
    const float inter = __fmul_rn(ih, iw);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter)) > thr;
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

// The scan phase shared by both kernels: the n_cand candidates listed in key_lo[] (selection-priority order) are walked in
// batches of 512 against the kept list, which grows in place.  All CTAs of a cluster call it with identical state.
struct ScanBuf {
    Box4* cand;            // [kBatch]
    float* cand_area;      // [kBatch]
    uint32_t* mask;        // [kBatch][kWords] row k: later candidates k suppresses
    uint32_t* dead;        // [kWords] suppressed by kept boxes of earlier batches / padding
    uint32_t* contested;   // [kWords] has a potential suppressor inside the batch
    uint32_t* has_row;     // [kWords] suppresses somebody inside the batch
    uint32_t* selw;        // [2 * kWords] final selection of the batch + rank offsets
    uint32_t* dpart;       // [kWords] phase (A) hits found by THIS CTA
    Box4* kept;            // [max_out]
    float* kept_area;      // [max_out]
    int* kept_idx;         // [max_out]
};

template <int CS, typename IdxT>
__device__ __forceinline__ void scan_candidates(const float* __restrict__ rows, int D, const IdxT* key_lo, int n_cand, float thr,
                                                int max_out, const ScanBuf& sb, int* s_kept, cg::cluster_group& cluster, int crank) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Box4* cand = sb.cand;
    float* cand_area = sb.cand_area;
    uint32_t* mask = sb.mask;
    uint32_t* dead = sb.dead;
    uint32_t* contested = sb.contested;
    uint32_t* has_row = sb.has_row;
    uint32_t* selw = sb.selw;
    uint32_t* dpart = sb.dpart;
    Box4* kept = sb.kept;
    float* kept_area = sb.kept_area;
    int* kept_idx = sb.kept_idx;
    for (int base = 0; base < n_cand; base += kBatch) {
        const int kept_before = (*s_kept);
        if (kept_before >= max_out) break;
        const int nb = min(kBatch, n_cand - base);
        if (tid < kBatch) {
            Box4 b;
            float a = -1.f;
            b.ymin = b.xmin = __int_as_float(0x7f800000);
            b.ymax = b.xmax = __int_as_float(0xff800000);
            if (tid < nb) load_box(rows + (size_t)key_lo[base + tid] * D, &b, &a);
            cand[tid] = b;
            cand_area[tid] = a;
        }
        if (tid < kWords) { dead[tid] = 0u; contested[tid] = 0u; has_row[tid] = 0u; dpart[tid] = 0u; }
        if (CS > 1)                                    // peers only deliver the non-zero words of the bit matrix
            for (int i = tid; i < kBatch * kWords; i += kNmsThreads) mask[i] = 0u;
        __syncthreads();
        {   // (A) two threads per candidate and CTA, each scanning one of the 2*CS parts of the kept list
            const int c = tid & (kBatch - 1), part = crank * 2 + (tid >> 9);
            const int chunk = (kept_before + 2 * CS - 1) / (2 * CS);
            const int j0 = part * chunk, j1 = min(kept_before, j0 + chunk);
            const Box4 me = cand[c];
            const float my_area = cand_area[c];
            bool hit = false;
            if (c < nb)
                for (int j = j0; j < j1; ++j)
                    if (iou_gt(me, my_area, kept[j], kept_area[j], thr)) { hit = true; break; }
            if (hit || c >= nb) atomicOr(&dpart[c >> 5], 1u << (c & 31));
        }
        __syncthreads();
        if (CS > 1) {
            cluster.sync();                            // every CTA's hits are in its dpart[]; all masks are zeroed
            if (tid < kWords * CS) atomicOr(&dead[tid & (kWords - 1)], cluster.map_shared_rank(dpart, tid / kWords)[tid & (kWords - 1)]);
        } else if (tid < kWords) {
            dead[tid] = dpart[tid];
        }
        __syncthreads();
        // (B) suppression matrix, one ballot per 32 pairs: row k, word w, lane l <-> candidate c = 32w + l.  The rows
        // are dealt round-robin to the CTAs of the cluster; a CTA delivers its non-zero words to every CTA.
        {
            uint32_t* mask_to = mask;
            uint32_t* contested_to = contested;
            uint32_t* has_row_to = has_row;
            if (CS > 1 && lane < CS) {                 // lane l of every warp writes to CTA l
                mask_to = cluster.map_shared_rank(mask, lane);
                contested_to = cluster.map_shared_rank(contested, lane);
                has_row_to = cluster.map_shared_rank(has_row, lane);
            }
            const bool writer = CS > 1 ? (lane < CS) : (lane == 0);
            for (int it = 0, k = warp; k < kBatch; k += kNmsThreads / 32, ++it) {
                if (CS > 1 && (it % CS) != crank) continue;              // warp-uniform
                const bool k_alive = !((dead[k >> 5] >> (k & 31)) & 1u);
                const int w0 = k >> 5;
                if (CS == 1 && (lane < w0 || !k_alive)) mask[k * kWords + (lane & (kWords - 1))] = 0u;
                if (!k_alive) continue;                                  // warp-uniform
                const Box4 bk = cand[k];
                const float ak = cand_area[k];
                uint32_t any = 0u;
                for (int w = w0; w < kWords; ++w) {
                    const int c = w * 32 + lane;
                    const bool bit = (c > k) && !((dead[w] >> lane) & 1u) && iou_gt(cand[c], cand_area[c], bk, ak, thr);
                    const uint32_t word = __ballot_sync(0xFFFFFFFFu, bit);
                    if (writer && (CS == 1 || word)) mask_to[k * kWords + w] = word;
                    any |= word;
                    if (word && writer) atomicOr(&contested_to[w], word);
                }
                if (any && writer) atomicOr(&has_row_to[k >> 5], 1u << (k & 31));
            }
        }
        if (CS > 1) cluster.sync(); else __syncthreads();
        // (C) resolve.  A candidate that nobody in the batch can suppress and that suppresses nobody is kept without
        // looking at the order; only the others (bit in `contested` or `has_row`) go through the sequential walk.
        if (warp == 0) {
            const uint32_t alive = (lane < kWords) ? ~dead[lane] : 0u;
            uint32_t walk = (lane < kWords) ? (alive & (contested[lane] | has_row[lane])) : 0u;
            uint32_t sel = alive & ~walk;
            uint32_t removed = 0u;
            while (true) {
                const uint32_t cur = walk & ~removed;
                const uint32_t vote = __ballot_sync(0xFFFFFFFFu, cur != 0u);
                if (!vote) break;
                const int src = __ffs(vote) - 1;
                const uint32_t wv = __shfl_sync(0xFFFFFFFFu, cur, src);
                const int bit = __ffs(wv) - 1;
                const int k = src * 32 + bit;
                if (lane < kWords) removed |= mask[k * kWords + lane];
                if (lane == src) { walk &= ~(1u << bit); sel |= 1u << bit; }
            }
            // cap at max_out in selection order: prefix counts over the 16 words
            int cnt = __popc(sel), pre = cnt;
            for (int o = 1; o < kWords; o <<= 1) {
                const int v = __shfl_up_sync(0xFFFFFFFFu, pre, o);
                if (lane >= o) pre += v;
            }
            const int before = pre - cnt;                            // selected in lower words
            const int room = max_out - kept_before;
            if (lane < kWords) {
                uint32_t keepw = sel;
                if (before >= room) keepw = 0u;
                else if (before + cnt > room) {                      // keep only the first (room - before) set bits
                    int need = room - before;
                    uint32_t t = sel, out = 0u;
                    while (need-- > 0) { const uint32_t low = t & (0u - t); out |= low; t ^= low; }
                    keepw = out;
                }
                selw[lane] = keepw;
                selw[kWords + lane] = (uint32_t)min(before, room);  // rank offset of this word
            }
            const int total = __shfl_sync(0xFFFFFFFFu, pre, kWords - 1);
            if (lane == 0) (*s_kept) = kept_before + min(total, room);
        }
        __syncthreads();
        if (tid < kBatch) {                                          // append in order, in parallel
            const uint32_t wsel = selw[tid >> 5];
            if ((wsel >> (tid & 31)) & 1u) {
                const int pos = kept_before + (int)selw[kWords + (tid >> 5)] + __popc(wsel & ((1u << (tid & 31)) - 1u));
                kept[pos] = cand[tid];
                kept_area[pos] = cand_area[tid];
                kept_idx[pos] = (int)key_lo[base + tid];
            }
        }
        __syncthreads();
    }
}

template <int CS>
__global__ void __launch_bounds__(kNmsThreads, 1)
nms_kernel(const float* __restrict__ rows_all, int N, int D, int obj_idx, float thr, int max_out, int NP,
           size_t region0, float* __restrict__ out_rows, int* __restrict__ out_idx, int* __restrict__ out_count, int out_img_rows) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = CS > 1 ? (int)cluster.block_rank() : 0;      // my share of the pair tests
    const int img = blockIdx.x / CS;
    const float* rows = rows_all + (size_t)img * N * D;
    // region 0: score histogram (top-K selection), then score keys during the sort, then the scan-phase scratch;
    // region 1: the candidate indices, which the sort leaves in selection-priority order.
    uint32_t* key_hi = reinterpret_cast<uint32_t*>(sm);                 // [NP]
    uint16_t* key_lo = reinterpret_cast<uint16_t*>(sm + region0);       // [NP]
    uint16_t* hist = reinterpret_cast<uint16_t*>(sm);                   // [kBins] (N <= 32768 < 65536 fits u16)
    uint8_t* aux = sm;
    Box4* cand = reinterpret_cast<Box4*>(aux);                          // [kBatch]
    Box4* kept = cand + kBatch;                                         // [max_out]
    float* cand_area = reinterpret_cast<float*>(kept + kMaxOut);        // [kBatch]
    float* kept_area = cand_area + kBatch;                              // [max_out]
    uint32_t* mask = reinterpret_cast<uint32_t*>(kept_area + kMaxOut);  // [kBatch][kWords] row k: later candidates k suppresses
    uint32_t* dead = mask + kBatch * kWords;                            // [kWords] suppressed by earlier batches / padding
    uint32_t* contested = dead + kWords;                                // [kWords] has a potential suppressor inside the batch
    uint32_t* has_row = contested + kWords;                             // [kWords] suppresses somebody inside the batch
    uint32_t* selw = has_row + kWords;                                  // [2 * kWords] final selection of the batch + rank offsets
    uint32_t* dpart = selw + 2 * kWords;                                // [kWords] phase (A) hits found by THIS CTA
    int* kept_idx = reinterpret_cast<int*>(selw + 6 * kWords);          // [max_out]
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
        {
            const ScanBuf sb{cand, cand_area, mask, dead, contested, has_row, selw, dpart, kept, kept_area, kept_idx};
            scan_candidates<CS, uint16_t>(rows, D, key_lo, n_cand, thr, max_out, sb, &s_kept, cluster, crank);
        }
        // the cut was sufficient iff the cap was reached or nothing was left below it
        if (full || s_kept >= max_out || n_cand == N) break;
        __syncthreads();
        full = true;                                       // rare: redo exactly over all candidates
    }

    // ---- 3. gather ----
    const int n_kept = s_kept;
    // out_img_rows > max_out ("packed" output, the all-gather message of byolo/dist.py): one more row per image whose first
    // element is the count, so that detections + count travel as ONE fp32 block
    float* orow = out_rows + (size_t)img * out_img_rows * D;
    if (out_img_rows > max_out && crank == 0 && tid < D) orow[(size_t)max_out * D + tid] = tid == 0 ? (float)n_kept : 0.f;
    for (int e = crank * kNmsThreads + tid; e < max_out * D; e += kNmsThreads * CS) {      // every CTA holds the full result
        const int k = e / D, c = e - k * D;
        orow[e] = (k < n_kept) ? rows[(size_t)kept_idx[k] * D + c] : 0.f;
    }
    if (out_idx && crank == 0)
        for (int k = tid; k < max_out; k += kNmsThreads) out_idx[(size_t)img * max_out + k] = (k < n_kept) ? kept_idx[k] : -1;
    if (tid == 0 && crank == 0 && out_count) out_count[img] = n_kept;
    if (CS > 1) cluster.sync();                            // nobody leaves while a peer may still touch its shared memory
}

// ----------------------------------------------------------------------------------------------------------------------
// Chunked variant for N > 32768 candidates per image (e.g. the reference's own ECP geometry, 1024 x 1920: N = 120960).
// The candidates cannot all be sorted in shared memory, so the greedy scan walks the global (score desc, index asc) order
// in CHUNKS of at most 4096 candidates: a radix select over the 64-bit composite (score key : ~index) finds a lower bound
// below the previous chunk such that at most 4096 candidates fall in between (one histogram pass over the scores per
// level, 15 bits per level; the first level almost always suffices), the chunk is compacted, sorted and scanned exactly
// like a sorted prefix in nms_kernel, and the kept list persists across chunks.  Stops at max_out kept or when every
// candidate has been visited, so the result is the exact greedy selection for any N and any number of score ties.
// ----------------------------------------------------------------------------------------------------------------------
constexpr int kSelBits = 15;
constexpr int kSelBins = 1 << kSelBits;
constexpr size_t kChunkScratchOff = (size_t)kTopK * 4;                                           // after key_hi
constexpr size_t kChunkScratchBytes = (sizeof(Box4) + 4) * kBatch + 4u * (kBatch * kWords + 9 * kWords);
constexpr size_t kChunkRegionA = (size_t)kSelBins * 4;                                           // histogram | key_hi + scan scratch
static_assert(kChunkScratchOff + kChunkScratchBytes <= kChunkRegionA, "scan scratch must fit beside the keys");
constexpr size_t kChunkSmem = kChunkRegionA + (size_t)kTopK * 4 + (sizeof(Box4) + 4 + 4) * kMaxOut;

__device__ __forceinline__ unsigned long long composite(uint32_t key, uint32_t idx) {      // larger = earlier in selection order
    return ((unsigned long long)key << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
}

__device__ void bitonic_sort32(uint32_t* key_hi, uint32_t* key_lo, int NP) {
    const int tid = threadIdx.x;
    for (int k = 2; k <= NP; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int pidx = tid; pidx < (NP >> 1); pidx += kNmsThreads) {
                const int i = ((pidx & ~(j - 1)) << 1) | (pidx & (j - 1));
                const int l = i | j;
                const uint32_t hi_i = key_hi[i], hi_l = key_hi[l];
                const uint32_t lo_i = key_lo[i], lo_l = key_lo[l];
                const bool i_first = (hi_i > hi_l) || (hi_i == hi_l && lo_i < lo_l);
                const bool up = (i & k) == 0;
                if (i_first != up) {
                    key_hi[i] = hi_l; key_hi[l] = hi_i;
                    key_lo[i] = lo_l; key_lo[l] = lo_i;
                }
            }
            __syncthreads();
        }
    }
}

template <int CS>
__global__ void __launch_bounds__(kNmsThreads, 1)
nms_chunked_kernel(const float* __restrict__ rows_all, int N, int D, int obj_idx, float thr, int max_out,
                   float* __restrict__ out_rows, int* __restrict__ out_idx, int* __restrict__ out_count, int out_img_rows) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = CS > 1 ? (int)cluster.block_rank() : 0;
    const int img = blockIdx.x / CS;
    const float* rows = rows_all + (size_t)img * N * D;
    uint32_t* hist = reinterpret_cast<uint32_t*>(sm);                                   // [kSelBins]
    uint32_t* key_hi = reinterpret_cast<uint32_t*>(sm);                                 // [kTopK] (after the selection)
    uint8_t* aux = sm + kChunkScratchOff;
    Box4* cand = reinterpret_cast<Box4*>(aux);                                          // [kBatch]
    float* cand_area = reinterpret_cast<float*>(cand + kBatch);                         // [kBatch]
    uint32_t* mask = reinterpret_cast<uint32_t*>(cand_area + kBatch);                   // [kBatch][kWords]
    uint32_t* dead = mask + kBatch * kWords;
    uint32_t* contested = dead + kWords;
    uint32_t* has_row = contested + kWords;
    uint32_t* selw = has_row + kWords;                                                  // [2 * kWords]
    uint32_t* dpart = selw + 2 * kWords;                                                // [kWords]
    uint32_t* key_lo = reinterpret_cast<uint32_t*>(sm + kChunkRegionA);                 // [kTopK] candidate indices of the chunk
    Box4* kept = reinterpret_cast<Box4*>(key_lo + kTopK);                               // [kMaxOut]  persists across chunks
    float* kept_area = reinterpret_cast<float*>(kept + kMaxOut);                        // [kMaxOut]
    int* kept_idx = reinterpret_cast<int*>(kept_area + kMaxOut);                        // [kMaxOut]
    __shared__ int s_kept, s_fill, s_bin, s_above;
    __shared__ int s_warp_sum[32];

    if (tid == 0) s_kept = 0;
    unsigned long long upper = ~0ull;          // exclusive: candidates with composite >= upper have been visited
    bool upper_open = true;                    // nothing visited yet (even composite ~0 would be eligible)
    int visited = 0;
    __syncthreads();

    while (visited < N && s_kept < max_out) {
        // ---- radix select: lower = largest bound such that #{lower <= c < upper} <= kTopK and >= 1 --------------------
        unsigned long long prefix = 0ull;      // digits fixed so far (high bits), all lower bits zero
        int fixed_bits = 0;
        unsigned long long lower = 0ull;
        while (true) {
            const int shift = max(64 - fixed_bits - kSelBits, 0);
            const int digit_bits = min(kSelBits, 64 - fixed_bits);
            for (int i = tid; i < kSelBins; i += kNmsThreads) hist[i] = 0u;
            __syncthreads();
            for (int i = tid; i < N; i += kNmsThreads) {
                const unsigned long long c = composite(ordered_key(rows[(size_t)i * D + obj_idx]), (uint32_t)i);
                const bool below = upper_open || c < upper;
                const bool match = fixed_bits == 0 || (c >> (64 - fixed_bits)) == (prefix >> (64 - fixed_bits));
                if (below && match) atomicAdd(&hist[(uint32_t)(c >> shift) & ((1u << digit_bits) - 1u)], 1u);
            }
            __syncthreads();
            // thread t owns bins [32t, 32t+32); suffix sums from the top bin downwards
            constexpr int per = kSelBins / kNmsThreads;
            int local = 0;
            for (int b = 0; b < per; ++b) local += (int)hist[tid * per + b];
            int incl = local;
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_down_sync(0xFFFFFFFFu, incl, o);
                if (lane + o < 32) incl += v;
            }
            if (lane == 0) s_warp_sum[warp] = incl;
            if (tid == 0) { s_bin = -1; s_above = 0; }
            __syncthreads();
            int above = 0;
            for (int w = warp + 1; w < 32; ++w) above += s_warp_sum[w];
            const int above_me = above + incl - local;                 // matching candidates in bins above my range
            // the crossing bin: the first bin (from the top) whose inclusive count exceeds kTopK, or - if the total fits -
            // "below the lowest bin" (s_bin stays -1)
            if (above_me <= kTopK && above_me + local > kTopK) {
                int cum = above_me;
                for (int b = per - 1; b >= 0; --b) {
                    const int h = (int)hist[tid * per + b];
                    if (cum + h > kTopK) { s_bin = tid * per + b; s_above = cum; break; }
                    cum += h;
                }
            }
            __syncthreads();
            const int bin = s_bin, above_bin = s_above;
            if (bin < 0) {                     // everything that matches fits in one chunk
                lower = prefix;                // (prefix has zeros below the fixed digits: the smallest composite with this prefix)
                break;
            }
            if (above_bin >= 1) {              // take the bins above the crossing bin; the crossing bin waits for the next chunk
                lower = prefix | ((unsigned long long)(bin + 1) << shift);
                break;
            }
            // the top remaining bin alone exceeds kTopK: fix its digit and look at the next digit inside it
            prefix |= (unsigned long long)bin << shift;
            fixed_bits += digit_bits;
            __syncthreads();
        }

        // ---- compaction of the chunk, sort ------------------------------------------------------------------------------
        __syncthreads();
        for (int i = tid; i < kTopK; i += kNmsThreads) { key_hi[i] = 0u; key_lo[i] = 0xFFFFFFFFu; }
        if (tid == 0) s_fill = 0;
        __syncthreads();
        for (int i = tid; i < N; i += kNmsThreads) {
            const uint32_t key = ordered_key(rows[(size_t)i * D + obj_idx]);
            const unsigned long long c = composite(key, (uint32_t)i);
            if ((upper_open || c < upper) && c >= lower) {
                const int pos = atomicAdd(&s_fill, 1);
                key_hi[pos] = key;
                key_lo[pos] = (uint32_t)i;
            }
        }
        __syncthreads();
        const int n_cand = s_fill;             // 1 .. kTopK by construction of `lower`
        bitonic_sort32(key_hi, key_lo, kTopK);

        // ---- batched greedy scan over the chunk (same code as nms_kernel) ------------------------------------------------
        {
            const ScanBuf sb{cand, cand_area, mask, dead, contested, has_row, selw, dpart, kept, kept_area, kept_idx};
            scan_candidates<CS, uint32_t>(rows, D, key_lo, n_cand, thr, max_out, sb, &s_kept, cluster, crank);
        }
        visited += n_cand;
        upper = lower;
        upper_open = false;
        __syncthreads();
    }

    const int n_kept = s_kept;
    // out_img_rows > max_out ("packed" output, the all-gather message of byolo/dist.py): one more row per image whose first
    // element is the count, so that detections + count travel as ONE fp32 block
    float* orow = out_rows + (size_t)img * out_img_rows * D;
    if (out_img_rows > max_out && crank == 0 && tid < D) orow[(size_t)max_out * D + tid] = tid == 0 ? (float)n_kept : 0.f;
    for (int e = crank * kNmsThreads + tid; e < max_out * D; e += kNmsThreads * CS) {
        const int k = e / D, c = e - k * D;
        orow[e] = (k < n_kept) ? rows[(size_t)kept_idx[k] * D + c] : 0.f;
    }
    if (out_idx && crank == 0)
        for (int k = tid; k < max_out; k += kNmsThreads) out_idx[(size_t)img * max_out + k] = (k < n_kept) ? kept_idx[k] : -1;
    if (tid == 0 && crank == 0 && out_count) out_count[img] = n_kept;
    if (CS > 1) cluster.sync();
}

// Per-class NMS support (the variant "used to produce the results for the paper", inference_epistemic.py:104-126: one NMS per
// class over the rows whose score of that class is strictly greater than every other class score).  A copy of the rows in
// which every OTHER row is neutralised - score -inf (sorts last) and a zero-area box (IoU 0 with everything, so it neither
// suppresses nor gets suppressed) - makes the class-agnostic kernel select exactly the subset's boxes first.
__global__ void class_filter_kernel(const float* __restrict__ rows, float* __restrict__ out, long long total, int D, int obj_idx,
                                    int cls_start, int cls_cnt, int cls) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= total) return;
    const float* src = rows + r * D;
    float* dst = out + r * D;
    const float mine = src[cls_start + cls];
    bool keep = true;
    for (int c = 0; c < cls_cnt; ++c)
        if (c != cls && !(mine > src[cls_start + c])) keep = false;      // tf.greater: strict, ties belong to no class
    for (int c = 0; c < D; ++c) dst[c] = src[c];
    if (!keep) {
        dst[0] = dst[1] = dst[2] = dst[3] = 0.f;
        dst[obj_idx] = __int_as_float(0xff800000);
    }
}

int launch_class_filter(const float* rows, long long total_rows, int D, int obj_idx, int cls_start, int cls_cnt, int cls, float* out,
                        cudaStream_t st) {
    BY_REQUIRE(obj_idx >= 4 && obj_idx < D && cls_start >= 4 && cls_cnt >= 1 && cls_start + cls_cnt <= D && cls >= 0 && cls < cls_cnt,
               "class filter: column indices out of range");
    if (total_rows == 0) return 0;
    class_filter_kernel<<<(unsigned)((total_rows + 255) / 256), 256, 0, st>>>(rows, out, total_rows, D, obj_idx, cls_start, cls_cnt, cls);
    BY_CUDA(cudaGetLastError());
    return 0;
}

int launch_nms(const float* rows, int B, int N, int D, int obj_idx, float iou_thr, int max_out, float* out_rows, int* out_idx,
               int* out_count, const NmsOptions& opt, cudaStream_t st) {
    BY_REQUIRE(N >= 0 && (long long)N * D < (1ll << 31), "NMS: candidate rows must be 32-bit indexable");
    BY_REQUIRE(max_out >= 1 && max_out <= kMaxOut, "max_out must be in [1, 2048]");
    BY_REQUIRE(iou_thr >= 0.f, "iou_thr must be >= 0");
    BY_REQUIRE(obj_idx >= 4 && obj_idx < D, "obj_idx out of range");
    BY_REQUIRE(opt.force_cs == 0 || opt.force_cs == 1 || opt.force_cs == 2 || opt.force_cs == 4 || opt.force_cs == 8,
               "cluster size must be 1, 2, 4 or 8");
    if (B == 0) return 0;
    const bool chunked = N > kMaxN || opt.force_chunked;      // tests force it on small N
    int out_img_rows = max_out + (opt.packed ? 1 : 0);
    int NP = 2;
    while (NP < N) NP <<= 1;
    size_t region0 = std::max({(size_t)NP * 4, kAuxBytes, N > 4096 ? (size_t)65536 * 2 : (size_t)0});
    const size_t smem = chunked ? kChunkSmem : region0 + (size_t)NP * 2;
    BY_REQUIRE(smem <= 227 * 1024 - 1024, "NMS shared memory budget exceeded");
    // cluster size: as many CTAs per image as keep all images resident at once (148 SMs, one CTA per SM).
    // Co-resident clusters of size 1 << i: queried once per (kernel family, device, smem size); the first query on a device
    // also opts the kernels in to the large dynamic shared memory.
    static std::mutex occ_mutex;
    static std::map<std::tuple<int, int, size_t>, std::array<int, 4>> occ_cache;
    int dev = 0;
    BY_CUDA(cudaGetDevice(&dev));
    auto make_cfg = [&](cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int blocks, int cs) {
        cfg.gridDim = dim3(blocks);
        cfg.blockDim = dim3(kNmsThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    };
    auto kernel_of = [&](int cs) -> const void* {
        if (chunked)
            return cs == 8 ? (const void*)nms_chunked_kernel<8> : cs == 4 ? (const void*)nms_chunked_kernel<4>
                 : cs == 2 ? (const void*)nms_chunked_kernel<2> : (const void*)nms_chunked_kernel<1>;
        return cs == 8 ? (const void*)nms_kernel<8> : cs == 4 ? (const void*)nms_kernel<4>
             : cs == 2 ? (const void*)nms_kernel<2> : (const void*)nms_kernel<1>;
    };
    auto occupancy = [&](int cs) -> int {
        if (cudaFuncSetAttribute(kernel_of(cs), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024) != cudaSuccess) return 0;
        cudaLaunchConfig_t cfg{};
        cudaLaunchAttribute attr[1];
        make_cfg(cfg, attr, cs * 64, cs);
        cfg.stream = nullptr;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kernel_of(cs), &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
        return n;
    };
    std::array<int, 4> max_clusters;
    {
        std::lock_guard<std::mutex> lock(occ_mutex);
        const auto key = std::make_tuple(chunked ? 1 : 0, dev, smem);
        auto it = occ_cache.find(key);
        if (it == occ_cache.end())
            it = occ_cache.emplace(key, std::array<int, 4>{occupancy(1), occupancy(2), occupancy(4), occupancy(8)}).first;
        max_clusters = it->second;
    }
    int cs = 1;
    for (int i = 3; i >= 1; --i)
        if (max_clusters[i] >= B) { cs = 1 << i; break; }
    if (opt.force_cs) cs = opt.force_cs;
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    make_cfg(cfg, attr, B * cs, cs);
    if (chunked) {
        void* args[] = {(void*)&rows, (void*)&N, (void*)&D, (void*)&obj_idx, (void*)&iou_thr, (void*)&max_out,
                        (void*)&out_rows, (void*)&out_idx, (void*)&out_count, (void*)&out_img_rows};
        BY_CUDA(cudaLaunchKernelExC(&cfg, kernel_of(cs), args));
    } else {
        void* args[] = {(void*)&rows, (void*)&N, (void*)&D, (void*)&obj_idx, (void*)&iou_thr, (void*)&max_out, (void*)&NP,
                        (void*)&region0, (void*)&out_rows, (void*)&out_idx, (void*)&out_count, (void*)&out_img_rows};
        BY_CUDA(cudaLaunchKernelExC(&cfg, kernel_of(cs), args));
    }
    BY_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace byolo
