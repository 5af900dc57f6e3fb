// K1: fused conv(1x1 | 3x3, stride 1 | 2) [+ dropout] + BN-shift + LeakyReLU [+ residual] as an implicit GEMM on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), operands staged by TMA into 128B/64B-swizzled
// shared memory.  Replaces the Conv2D + FusedBatchNorm + LeakyRelu (+ RandomUniform/Floor/Mul dropout, + Add) node
// groups that /root/reference/lib_yolo/layers.py:545-575 (conv), :505-507 (residual), :521-524 (dropout) create.
//
// GEMM view:  D[M = output pixels, N = cout] = A[M, K] * W[N, K]^T,  K = taps * (C1 + C2), fp16 operands, fp32 accumulate.
// Activations are dense NHWC (common.cuh), GEMM row m = output pixel (s, y, x) in row-major order.  The A tile of a
// 128-pixel M tile is fetched by the TMA unit itself:
//   * 3x3 convs: an IM2COL tensor map over [C, W, H, S] (bounding box corners -1/-1, traversal stride = conv stride):
//     one load per filter tap (r, s) = "128 consecutive output pixels starting at pixel m0, displaced by (s, r)"; the
//     unit walks across image rows and samples and zero-fills taps that fall outside the image, which is exactly SAME
//     padding at stride 1 and the explicit pad-1-on-all-sides of the darknet downsample convs (layers.py:616-635).
//     No padded buffers, no border rows in the GEMM, no im2col matrix.
//   * 1x1 convs: plain 2D boxes of the [rows, C] matrix; a channel concat [in1, in2] (route, layers.py:583-592) reads
//     its K range from two maps.
//   * MC stacking (stack_feature_map, layers.py:595-597) is a 5D im2col map [C, W, H, T, B] whose T stride is ZERO:
//     sample s = b*T + t of the stacked tensor reads image b of the cached backbone map.  No copy.
//
// Structure: persistent CTAs (one per SM), 384 threads = warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM
// allocator, warps 4-11 epilogue (TMEM lane quadrant = warp % 4, two warps per quadrant split the columns); smem ring of `num_stages` {A,B} tiles with
// full/empty mbarriers; two TMEM accumulators so the epilogue of tile i overlaps the main loop of tile i+1.
#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "conv_umma.cuh"

namespace byolo {

// --------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// ---- 2-CTA (cta_group::2) forms: the TMA of either CTA signals the LEADER's barrier (peer bit 24 cleared), the
// leader's MMA commits are multicast to the barrier at the same offset in both CTAs.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"((uint64_t)map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
// im2col loads: {c, w, h, n} = first channel and the (input-space) position of the filter's top-left corner for the first
// output pixel of the tile; {ow, oh} = filter tap.  CG = 2: completion lands on the leader CTA's barrier.
template <int CG>
__device__ __forceinline__ void tma_im2col_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h, int n, int ow, int oh) {
    if constexpr (CG == 2) {
        asm volatile(
            "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
            "l"((uint64_t)map), "r"(bar & kPeerBitMask), "r"(c), "r"(w), "r"(h), "r"(n), "h"((uint16_t)ow), "h"((uint16_t)oh)
            : "memory");
    } else {
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
            "l"((uint64_t)map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"((uint16_t)ow), "h"((uint16_t)oh)
            : "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tma_im2col_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h, int d, int n) {
    if constexpr (CG == 2) {
        asm volatile(
            "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2], {%8, %8, %8};" ::"r"(dst),
            "l"((uint64_t)map), "r"(bar & kPeerBitMask), "r"(c), "r"(w), "r"(h), "r"(d), "r"(n), "h"((uint16_t)0)
            : "memory");
    } else {
        asm volatile(
            "cp.async.bulk.tensor.5d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2], {%8, %8, %8};" ::"r"(dst),
            "l"((uint64_t)map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(d), "r"(n), "h"((uint16_t)0)
            : "memory");
    }
}
__device__ __forceinline__ void umma_f16_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {     // arrive on CTA `cta`'s copy of `bar`
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
        "}" ::"r"(bar),
        "r"(cta)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)map), "r"(src),
                 "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// One lane of a converged warp (elect.sync): tcgen05/TMA operands live in uniform registers, so issuing from a
// divergent `lane == 0` branch makes the compiler wrap every instruction in an ELECT/R2UR.BROADCAST loop.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(pred));
    return pred != 0;
}
// MMA with the smem descriptors given as (low word, shared high word): only the start-address field changes per call.
template <int CG>
__device__ __forceinline__ void umma_f16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                              uint32_t accumulate) {
    if constexpr (CG == 1) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            ".reg .b64 da, db;\n"
            "mov.b64 da, {%1, %3};\n"
            "mov.b64 db, {%2, %3};\n"
            "setp.ne.b32 p, %5, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n"
            "}" ::"r"(d_tmem),
            "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            ".reg .b64 da, db;\n"
            "mov.b64 da, {%1, %3};\n"
            "mov.b64 db, {%2, %3};\n"
            "setp.ne.b32 p, %5, 0;\n"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n"
            "}" ::"r"(d_tmem),
            "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor (tcgen05 "smem descriptor"): start address, stride between 8-row
// groups (SBO), version 1, swizzle mode.  The leading-dimension offset is unused for swizzled K-major tiles.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout_type << 61;
    return d;
}

constexpr int kCtlWarps = 4;            // TMA producer (A), MMA issuer, TMEM allocator, TMA producer (B)
constexpr int kMaxBias = 1024;         // floats of BN shift / bias staged in shared memory
constexpr int kStageOutBytes = 2048;   // per epilogue warp: 32 rows x 64 B output chunk for the TMA store
constexpr int kTileM = 128;
constexpr int kMaxStages = 8;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;
constexpr int kChunkStride = 128;      // split mode: N tile <= 128 columns, four chunk buffers

struct SmemCtl {
    uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    uint64_t acc_full[4];       // fp16 mode: two 256-column accumulators; split mode: four 128-column chunk buffers
    uint64_t acc_empty[4];
    uint32_t tmem_base;
    uint32_t pad[3];
    alignas(16) float bias[kMaxBias];      // read as float4 by the epilogue
};
static_assert(offsetof(SmemCtl, bias) % 16 == 0, "bias must be 16-byte aligned");

__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) { return f.d == 1 ? n : (__umulhi(n, f.mul) >> f.shr); }

template <int CG>
__device__ __forceinline__ void tma_a2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    if constexpr (CG == 2) tma_load_2d_2cta(dst, map, bar, c0, c1);
    else tma_load_2d(dst, map, bar, c0, c1);
}

// ------------------------------------------------------------------------------------------------------ epilogue math
// One chunk = 32 accumulator columns of one row (thread = TMEM lane = pixel): [dropout] + shift + leaky [+ residual]
// -> 32 fp16 = 64 output bytes.  Specialised at compile time: the run-time flags of the old generic loop cost more
// issue slots than the arithmetic.  Dropout: keep iff the element's 16 random bits >= thr16, written as 32-bit
// compares ((w << 16) >= thr32 for the low half, w >= thr32 for the high half: identical decisions, no extraction),
// and the mask multiplier goes through one FFMA with the shift.
struct DropRow {
    uint32_t group0;         // (elem index of this row's channel 0) >> 3
    uint32_t t, image;
};

template <bool DROP, bool RES, bool X3>
__device__ __forceinline__ void chunk_f16(const uint32_t (&raw)[32], const float* __restrict__ bias_s, const uint4 (&res)[X3 ? 8 : 4],
                                          uint4 (&o)[X3 ? 8 : 4], const Dropout& d, const DropRow& dr, uint32_t group_c) {
    const uint32_t thr32 = d.thr16 << 16;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        float v[8];
        const float4 b0 = *reinterpret_cast<const float4*>(bias_s + 8 * g), b1 = *reinterpret_cast<const float4*>(bias_s + 8 * g + 4);
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        if constexpr (DROP) {
            const uint4 r = philox4x32_10(make_uint4(dr.group0 + group_c + (uint32_t)g, (uint32_t)d.layer_id, dr.t, dr.image), d.seed_lo, d.seed_hi);
            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // dropped: the shift alone; kept: x * 1/(1-p) + shift in one FFMA (predicated on the compare)
                v[2 * i] = b[2 * i];
                v[2 * i + 1] = b[2 * i + 1];
                if ((w[i] << 16) >= thr32) v[2 * i] = fmaf(__uint_as_float(raw[8 * g + 2 * i]), d.keep_scale, b[2 * i]);
                if (w[i] >= thr32) v[2 * i + 1] = fmaf(__uint_as_float(raw[8 * g + 2 * i + 1]), d.keep_scale, b[2 * i + 1]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(raw[8 * g + j]) + b[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.1f * v[j]);
        if constexpr (RES) {
            const __half2* rh = reinterpret_cast<const __half2*>(&res[g]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 f = __half22float2(rh[j]);
                if constexpr (X3) {                      // shortcut value = hi + lo (exact in fp32)
                    const float2 fl = __half22float2(reinterpret_cast<const __half2*>(&res[4 + g])[j]);
                    f.x += fl.x;
                    f.y += fl.y;
                }
                v[2 * j] += f.x;
                v[2 * j + 1] += f.y;
            }
        }
        __half2* oh = reinterpret_cast<__half2*>(&o[g]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            oh[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
            if constexpr (X3) {                          // lo = fp16(v - hi): the value is carried as hi + lo (22 significant bits)
                const float2 h = __half22float2(oh[j]);
                reinterpret_cast<__half2*>(&o[4 + g])[j] = __floats2half2_rn(v[2 * j] - h.x, v[2 * j + 1] - h.y);
            }
        }
    }
}

// Epilogue of the 8 epilogue warps.  KIND selects the store path and the fused extras (EpiKind in conv_umma.cuh).
// X3 (split-fp16 mode): activations are carried as hi + lo fp16 pairs, pixel layout [hi C | lo C]; the epilogue splits its
// fp32 result and stores both halves (second staging block, second TMA store), the shortcut is read as hi + lo.
template <int CG, int EW, int KIND, bool X3>
__device__ __forceinline__ void run_epilogue(const CUtensorMap* map_o, const UmmaParams& p, SmemCtl* ctl, uint32_t tmem_base,
                                             uint32_t out_stage, int warp, int lane, uint32_t rank, int first_tile, int tile_step) {
    constexpr bool kF32 = KIND == EPI_F32;
    constexpr bool kDropT = KIND == EPI_F16_DROP_T;     // T-invariant conv: one accumulator row -> T masked output samples
    constexpr bool kDrop = KIND == EPI_F16_DROP || kDropT;
    constexpr bool kRes = KIND == EPI_F16_RES;
    constexpr bool kUp = KIND == EPI_UPSAMPLE;
    constexpr int CH = kF32 ? 16 : 32;
    constexpr int NO = (X3 && !kF32) ? 8 : 4;    // uint4 per row and chunk: hi [+ lo]
    constexpr int NP = NO / 4;                   // planes written
    const int quad = warp & 3;
    constexpr int kSplit = EW / 4;               // warps sharing a TMEM lane quadrant: they interleave the column chunks
    const int hsel = (warp - kCtlWarps) >> 2;
    const Epilogue& ep = p.ep;
    const int Ho = p.gout.H, Wo = p.gout.W;
    const uint32_t out_rows = (uint32_t)p.out_rows;
    const int nchunks = p.BN / CH;
    const int nnt_shift = p.nnt_shift, BN = p.BN, ldc = ep.ldc;
    const int pitch = X3 && !kF32 ? 2 * ldc : ldc;           // halves per pixel of the output / shortcut buffers
    const uint32_t stg = out_stage + (uint32_t)(warp - kCtlWarps) * kStageOutBytes;
    const uint32_t stg_row = stg + lane * 64;
    const uint32_t swz = (uint32_t)((lane >> 1) & 3);
    const uint32_t acc_full0 = smem_u32(&ctl->acc_full[0]), acc_empty0 = smem_u32(&ctl->acc_empty[0]);
    const int i = quad * 32 + lane;              // row of the tile == TMEM lane
    uint4 rnext[NO] = {};
    auto res_prefetch = [&](int t, int ch) {     // residual (shortcut) values of chunk `ch` of tile `t` for this thread's row
        if (t >= p.num_tiles || ch >= nchunks) return;
        const uint32_t r = (uint32_t)((t >> nnt_shift) * CG + (int)rank) * kTileM + (uint32_t)i;
        if (r >= out_rows) return;
        const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(ep.residual) + (size_t)r * pitch +
                                                         (t & ((1 << nnt_shift) - 1)) * BN + ch * CH);
#pragma unroll
        for (int j = 0; j < 4; ++j) rnext[j] = __ldg(rp + j);
        if constexpr (NO == 8) {
            const uint4* rl = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(rp) + ldc);
#pragma unroll
            for (int j = 0; j < 4; ++j) rnext[4 + j] = __ldg(rl + j);
        }
    };
    uint32_t tile_it = 0, chunk_it = 0;
    for (int tile = first_tile; tile < p.num_tiles; tile += tile_step, ++tile_it) {
        const uint32_t as = tile_it & 1, aphase = (tile_it >> 1) & 1;
        const int m_tile = (tile >> nnt_shift) * CG + (int)rank;
        const int n0 = (tile & ((1 << nnt_shift) - 1)) * BN;
        const uint32_t row = (uint32_t)m_tile * kTileM + (uint32_t)i;       // GEMM row == output pixel index
        const bool valid = row < out_rows;
        DropRow dr{0u, 0u, 0u};
        int us = 0, uy = 0, ux = 0;
        if constexpr (kDrop || kUp) {
            const uint32_t su = fdiv(row, p.fd_plane), rem = row - su * p.fd_plane.d;        // sample, pixel inside the map
            if constexpr (kDropT) {                  // GEMM rows run over the B images; t is the loop below
                dr.image = (uint32_t)ep.drop.image0 + su;
                dr.group0 = (rem * (uint32_t)ep.cout + (uint32_t)n0) >> 3;
                us = (int)su;
                ux = (int)rem;
            } else if constexpr (kDrop) {
                const uint32_t im = fdiv(su, p.fd_T);
                dr.t = su - im * p.fd_T.d;
                dr.image = (uint32_t)ep.drop.image0 + im;
                dr.group0 = (rem * (uint32_t)ep.cout + (uint32_t)n0) >> 3;
            } else {
                us = (int)su;
                uy = (int)fdiv(rem, p.fd_w);
                ux = (int)rem - uy * Wo;
            }
        }
        uint32_t up_q[4] = {0u, 0u, 0u, 0u};          // kUp: top-left destination pixel of rows (lane >> 2) + 8k, or ~0 if past the end
        if constexpr (kUp) {
            const uint32_t q00 = valid ? (uint32_t)((us * 2 * Ho + 2 * uy) * (2 * Wo) + 2 * ux) : 0xFFFFFFFFu;
#pragma unroll
            for (int k = 0; k < 4; ++k) up_q[k] = __shfl_sync(0xFFFFFFFFu, q00, (lane >> 2) + 8 * k);
        }
        if constexpr (kRes && !X3) {
            if (tile_it == 0) res_prefetch(tile, hsel);      // later tiles: requested during the previous tile's last chunk
        }
        // Split mode: the accumulator arrives in chunks (see the MMA issuer); every chunk is added - round to nearest - to
        // register sums of this warp's <= 2 column chunks and its TMEM buffer is handed back at once.
        constexpr int kSums = X3 ? 2 : 1;
        float sums[kSums][CH];
        if constexpr (X3) {
            for (int c = 0; c < p.chunks_per_tile; ++c, ++chunk_it) {
                const uint32_t cas = chunk_it & 3, cphase = (chunk_it >> 2) & 1;
                mbar_wait(acc_full0 + 8 * cas, cphase);
                tc_fence_after();
                const uint32_t ctaddr = tmem_base + cas * kChunkStride + ((uint32_t)(quad * 32) << 16);
#pragma unroll
                for (int k = 0; k < kSums; ++k) {
                    const int ch = hsel + kSplit * k;
                    if (ch < nchunks) {
                        uint32_t t[32];
                        if constexpr (kF32) tmem_ld16(ctaddr + ch * CH, t); else tmem_ld32(ctaddr + ch * CH, t);
                        tmem_ld_wait();
                        if (c == 0) {
#pragma unroll
                            for (int j = 0; j < CH; ++j) sums[k][j] = __uint_as_float(t[j]);
                        } else {
#pragma unroll
                            for (int j = 0; j < CH; ++j) sums[k][j] += __uint_as_float(t[j]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CG == 1) mbar_arrive(acc_empty0 + 8 * cas);
                    else mbar_arrive_cluster(acc_empty0 + 8 * cas, 0);
                }
            }
        } else {
            mbar_wait(acc_full0 + 8 * as, aphase);
            tc_fence_after();
        }
        const uint32_t taddr = tmem_base + as * kAccStride + ((uint32_t)(quad * 32) << 16);
        // one column chunk of this row: accumulator values -> epilogue math -> store(s)
        auto do_chunk = [&](const int ch, auto&& fetch) {
            const int c0 = ch * CH;                  // column inside the tile
            uint32_t raw[32];
            fetch(raw, c0);
            uint4 rcur[NO] = {};
            if constexpr (kRes) {
                if constexpr (X3) res_prefetch(tile, ch);        // split mode: the epilogue is off the critical path, load at use
#pragma unroll
                for (int j = 0; j < NO; ++j) rcur[j] = rnext[j];
                // the shortcut values of this warp's NEXT chunk - of this tile or, after the last one, of its next tile -
                // are requested now, a whole chunk (or an accumulator wait) ahead of their use
                if constexpr (!X3) {
                    if (ch + kSplit < nchunks) res_prefetch(tile, ch + kSplit);
                    else res_prefetch(tile + tile_step, hsel);
                }
            }
            if constexpr (!X3) tmem_ld_wait();
            const int c = n0 + c0;
            uint4 o4[NO];                            // the 64 output bytes of this row (x2: hi, lo)
            if constexpr (kF32) {
                const float* bs = ctl->bias + c;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 b = *reinterpret_cast<const float4*>(bs + 4 * j);
                    o4[j] = make_uint4(__float_as_uint(__uint_as_float(raw[4 * j]) + b.x), __float_as_uint(__uint_as_float(raw[4 * j + 1]) + b.y),
                                       __float_as_uint(__uint_as_float(raw[4 * j + 2]) + b.z), __float_as_uint(__uint_as_float(raw[4 * j + 3]) + b.w));
                }
            } else if constexpr (!kDropT) {
                chunk_f16<kDrop, kRes, X3>(raw, ctl->bias + c, rcur, o4, ep.drop, dr, (uint32_t)c0 >> 3);
            }
            if constexpr (kDropT) {
                // conv "75" (yolov3.py:538-544): its input is the MC-stacked backbone map, identical for the T samples of an
                // image, so the GEMM ran once per image; here the same accumulator values get the T dropout masks and go to
                // the T output samples: GEMM row (b, pixel) -> output rows ((b*T + t)*plane + pixel).  A warp's 32-row block
                // that lies inside one image is 32 contiguous output rows (TMA store, as everywhere); the few blocks that
                // run across an image boundary or past the last row are written with per-thread 16-byte stores.
                const int plane = (int)p.fd_plane.d, T = (int)p.fd_T.d;
                const int row0 = m_tile * kTileM + quad * 32;
                const int b0 = (int)fdiv((uint32_t)row0, p.fd_plane), pix0 = row0 - b0 * plane;
                const bool inside = pix0 + 32 <= plane && (uint32_t)(row0 + 32) <= out_rows;
                __half* ob = reinterpret_cast<__half*>(ep.out);
                for (int t = 0; t < T; ++t) {
                    dr.t = (uint32_t)t;
                    chunk_f16<true, false, X3>(raw, ctl->bias + c, rcur, o4, ep.drop, dr, (uint32_t)c0 >> 3);
#pragma unroll
                    for (int pl = 0; pl < NP; ++pl) {
                        if (inside) {
                            if (lane == 0) bulk_wait_read0();
                            __syncwarp();
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint32_t dst = stg_row + (((uint32_t)j ^ swz) << 4);
                                const uint4 v = o4[4 * pl + j];
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                            }
                            fence_async_smem();
                            __syncwarp();
                            if (lane == 0) {
                                tma_store_2d(map_o, stg, pl * ldc + c, (b0 * T + t) * plane + pix0);
                                bulk_commit();
                            }
                        } else if (valid) {
                            uint4* dst = reinterpret_cast<uint4*>(ob + ((size_t)(us * T + t) * plane + ux) * pitch + pl * ldc + c);
#pragma unroll
                            for (int j = 0; j < 4; ++j) dst[j] = o4[4 * pl + j];
                        }
                    }
                }
            } else if constexpr (!kUp) {
                // the warp's 32 x 64 B block(s) -> 64B-swizzled staging -> TMA store (rows past the end are clipped by the unit)
                // (split mode: hi then lo through the same block - its main loop is 3x longer, the epilogue has time to spare,
                // and a second block would cost a pipeline stage)
#pragma unroll
                for (int pl = 0; pl < NP; ++pl) {
                    if (lane == 0) bulk_wait_read0();    // the previous store of this warp has drained the staging block
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t dst = stg_row + (((uint32_t)j ^ swz) << 4);
                        const uint4 v = o4[4 * pl + j];
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(map_o, stg, pl * ldc + c, m_tile * kTileM + quad * 32);
                        bulk_commit();
                    }
                }
            } else {
                // nearest-neighbour x2: every row goes to four destination pixels.  The chunk is transposed through the
                // staging block so that one store instruction writes 8 rows x 64 contiguous bytes (full 32 B sectors).
                __half* ob = reinterpret_cast<__half*>(ep.out);
#pragma unroll
                for (int pl = 0; pl < NP; ++pl) {
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t dst = stg_row + (((uint32_t)j ^ swz) << 4);
                        const uint4 v = o4[4 * pl + j];
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                    }
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t rr = (uint32_t)(lane >> 2) + 8u * k, jj = (uint32_t)lane & 3u;
                        uint4 v;
                        const uint32_t src = stg + rr * 64 + ((jj ^ ((rr >> 1) & 3u)) << 4);
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src));
                        if (up_q[k] != 0xFFFFFFFFu) {
#pragma unroll
                            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                                for (int dx = 0; dx < 2; ++dx) {
                                    const size_t q = (size_t)up_q[k] + (size_t)(dy * 2 * Wo + dx);
                                    *reinterpret_cast<uint4*>(ob + q * pitch + pl * ldc + c + jj * 8) = v;
                                }
                        }
                    }
                }
            }
        };
#ifdef BYOLO_DBG_HOOKS
        if (!(p.dbg & 2))
#endif
        {
            if constexpr (X3) {
#pragma unroll
                for (int k = 0; k < kSums; ++k) {
                    const int ch = hsel + kSplit * k;
                    if (ch < nchunks)
                        do_chunk(ch, [&](uint32_t (&raw)[32], int) {
#pragma unroll
                            for (int j = 0; j < CH; ++j) raw[j] = __float_as_uint(sums[k][j]);
                        });
                }
            } else {
                for (int ch = hsel; ch < nchunks; ch += kSplit)
                    do_chunk(ch, [&](uint32_t (&raw)[32], int c0) {
                        if constexpr (kF32) tmem_ld16(taddr + c0, raw); else tmem_ld32(taddr + c0, raw);
                    });
            }
        }
        if constexpr (!X3) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CG == 1) mbar_arrive(acc_empty0 + 8 * as);
                else mbar_arrive_cluster(acc_empty0 + 8 * as, 0);      // the leader's MMA thread waits for both CTAs
            }
        }
    }
    if (!kUp && lane == 0) bulk_wait0();          // all output writes complete before the CTA retires
}

// CG = 1: one CTA per 128-row tile.  CG = 2: a CTA pair (cluster of 2) works on 256 rows x BN: each CTA stages its own
// 128 A rows and HALF of the B tile, the leader CTA issues tcgen05.mma.cta_group::2 (M = 256) for both, every CTA runs
// the epilogue of its own 128 accumulator rows.  Per SM and MMA cycle this moves 2/3 of the bytes of CG = 1.
// AM: how the producer addresses the A operand (AMode in conv_umma.cuh).
// X3: split-fp16 mode.  Every operand is a hi + lo fp16 pair (activations: pixel layout [hi C | lo C]; weights: rows
// [0, cout_pad) hi, [cout_pad, 2 cout_pad) lo); a K block stages {A_hi, A_lo, B_hi, B_lo} and issues three MMAs per K step
// into the same accumulator: hi*hi + hi*lo + lo*hi (lo*lo is below fp32 resolution).
template <int CG, int AM, int EW, bool X3>
__global__ void __launch_bounds__((kCtlWarps + EW) * 32, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_a2,
                 const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_o, const UmmaParams p) {
    extern __shared__ uint8_t smem_raw[];
    // ring of {A,B} tiles, 1024B aligned (SWIZZLE_128B atoms are 1024B); output staging and control block behind it
    const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
    constexpr uint32_t NPL = X3 ? 2 : 1;                          // operand planes (hi [, lo])
    const uint32_t a_span = p.a_bytes * NPL;                      // A_hi [A_lo], then B_hi [B_lo]
    const uint32_t kb_bytes = a_span + p.b_bytes * NPL;           // one K block of {A, B}
    const uint32_t stage_bytes = kb_bytes * p.kbs;                // a stage holds kbs K blocks (8 MMAs per commit when kbs = 2)
    const uint32_t out_stage = ring + p.num_stages * stage_bytes;
    SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem_raw + (ring - smem_u32(smem_raw)) + (size_t)p.num_stages * stage_bytes +
                                              EW * kStageOutBytes);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;     // 0 = leader (issues the MMAs)
    const int first_tile = blockIdx.x / CG, tile_step = gridDim.x / CG;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a1);
        tma_prefetch_desc(&map_a2);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_o);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.num_stages; ++s) {
            mbar_init(smem_u32(&ctl->full[s]), p.bsplit ? 2 : 1);       // split: the A and the B producer warp each post their byte count
            mbar_init(smem_u32(&ctl->empty[s]), 1);
        }
        for (int s = 0; s < 4; ++s) {
            mbar_init(smem_u32(&ctl->acc_full[s]), 1);
            mbar_init(smem_u32(&ctl->acc_empty[s]), EW * CG);      // CG = 2: both CTAs' epilogues release the leader
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if constexpr (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)),
                         "r"(kTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)),
                         "r"(kTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    for (int i = threadIdx.x; i < (p.BN << p.nnt_shift); i += (kCtlWarps + EW) * 32) ctl->bias[i] = __ldg(p.ep.bias + i);
    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();           // the peer's barriers exist before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = ctl->tmem_base;
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch, bias
    // staging - none of it produced by the previous kernel) may overlap the tail of the previous grid; from here on we
    // read its output.  The next grid may start its own prologue as soon as our CTAs retire.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (p.clk && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long g;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
        p.clk[0] = clock64();
        p.clk[1] = g;
    }

    // Both control loops below are latency chains of ONE warp: every instruction between two TMA / MMA issues is on the
    // critical path of the short layers (ncu, profiles/r01: ~1200 cycles per stage with divisions and parameter
    // re-loads in the loop), so all loop state lives in registers and advances by additions only.
    const int num_kb = p.taps * (p.kb1 + p.kb2);
    const int num_stages = p.num_stages, kbs = p.kbs, num_tiles = p.num_tiles, nnt_shift = p.nnt_shift;
    const uint32_t full0 = smem_u32(&ctl->full[0]), empty0 = smem_u32(&ctl->empty[0]);

    if (warp == 0) {
        // ================================ TMA producer, A operand (whole warp converged, one elected lane issues) =====
        const int BK = p.BK, kbt = p.kb1 + p.kb2, kb1 = p.kb1, cstride = p.stride;
        const uint32_t a_bytes = p.a_bytes, b_bytes = p.b_bytes;
        const int c1_lo = p.c1_lo, c2_lo = p.c2_lo, b_lo_row = p.b_lo_row;
        const bool bsplit = p.bsplit != 0;
        uint32_t stage = 0, phase = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
            const int m_tile = (tile >> nnt_shift) * CG + (int)rank;
            const int m0 = m_tile * kTileM;
            int cw = 0, chh = 0, cn = 0, ct = 0;          // first output pixel of the tile as tensor-map coordinates
            if constexpr (AM != A_TILED) {
                const uint32_t q = fdiv((uint32_t)m0, p.fd_w);                // m0 = (n * Ho + ho) * Wo + wo
                const uint32_t n = fdiv(q, p.fd_h);
                cw = m0 - (int)(q * p.fd_w.d);
                chh = (int)(q - n * p.fd_h.d);
                cn = (int)n;
                if constexpr (AM == A_IM2COL) {           // top-left corner of the filter window in input space (pad 1)
                    cw = cw * cstride - 1;
                    chh = chh * cstride - 1;
                } else {                                  // stacked source: sample n = b*T + t reads image b
                    const uint32_t b = fdiv(n, p.fd_T);
                    ct = (int)(n - b * p.fd_T.d);
                    cn = (int)b;
                }
            }
            int cb = 0, r = 0, s = 0;                     // channel block, filter tap (r, s)
            int b_k = 0;
            const int n0b = (tile & ((1 << nnt_shift) - 1)) * p.BN + (int)rank * p.b_rows * (CG - 1);
            for (int kb = 0; kb < num_kb; kb += kbs) {
                const int nkb = min(kbs, num_kb - kb);
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                const uint32_t full = full0 + 8 * stage;
                const bool leader_lane = elect_one();
                uint32_t sa = ring + stage * stage_bytes;
#ifdef BYOLO_DBG_HOOKS
                if (p.dbg & 1) {                       // experiment: barrier traffic without any bytes moving
                    if (leader_lane && rank == 0) mbar_arrive(full);
                    __syncwarp();
                    if (++stage == (uint32_t)num_stages) { stage = 0; phase ^= 1; }
                    continue;
                }
#endif
                if (leader_lane && rank == 0) mbar_expect_tx(full, (bsplit ? a_span : kb_bytes) * nkb * CG);    // bytes of both CTAs land on the leader's barrier
                for (int j = 0; j < nkb; ++j) {
                    if (leader_lane) {
#pragma unroll
                        for (uint32_t pl = 0; pl < NPL; ++pl) {       // plane 0 = hi (or the only one), plane 1 = lo
                            const uint32_t da = sa + pl * a_bytes;
                            const int o1 = pl ? c1_lo : 0, o2 = pl ? c2_lo : 0;
                            if constexpr (AM == A_IM2COL) {
                                tma_im2col_4d<CG>(da, &map_a1, full, cb * BK + o1, cw, chh, cn, s, r);
                            } else if constexpr (AM == A_STACK1) {
                                tma_im2col_5d<CG>(da, &map_a1, full, cb * BK + o1, cw, chh, ct, cn);
                            } else if constexpr (AM == A_STACK2) {
                                if (cb < kb1) tma_a2d<CG>(da, &map_a1, full, cb * BK + o1, m0);
                                else tma_im2col_5d<CG>(da, &map_a2, full, (cb - kb1) * BK + o2, cw, chh, ct, cn);
                            } else {
                                if (cb < kb1) tma_a2d<CG>(da, &map_a1, full, cb * BK + o1, m0);
                                else tma_a2d<CG>(da, &map_a2, full, (cb - kb1) * BK + o2, m0);
                            }
                            if (!bsplit) tma_a2d<CG>(sa + a_span + pl * b_bytes, &map_b, full, b_k, n0b + (pl ? b_lo_row : 0));
                        }
                    }
                    sa += kb_bytes;
                    b_k += BK;
                    if (++cb == kbt) {                    // next filter tap
                        cb = 0;
                        if (++s == 3) { s = 0; ++r; }
                    }
                }
                __syncwarp();
                if (++stage == (uint32_t)num_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (leader CTA only when CG = 2; warp converged, one lane issues) ====
        if (rank == 0) {
            const uint32_t desc_hi = (uint32_t)(make_smem_desc(0, p.sbo_bytes, p.layout_type) >> 32);
            const uint32_t lo_fixed = 1u << 16;                      // leading-dimension field (unused for swizzled K-major)
            const uint32_t idesc = p.idesc, a16 = p.a_bytes >> 4, kb16 = kb_bytes >> 4, as16 = a_span >> 4, b16 = p.b_bytes >> 4;
            const bool bk64 = p.BK == 64;
            const uint32_t acc_full0 = smem_u32(&ctl->acc_full[0]), acc_empty0 = smem_u32(&ctl->acc_empty[0]);
            uint32_t stage = 0, phase = 0, tile_it = 0;
          if constexpr (!X3) {
            for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++tile_it) {
                const uint32_t as = tile_it & 1, aphase = (tile_it >> 1) & 1;
                mbar_wait(acc_empty0 + 8 * as, aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * kAccStride;
                for (int kb = 0; kb < num_kb; kb += kbs) {
                    const int nkb = min(kbs, num_kb - kb);
                    mbar_wait(full0 + 8 * stage, phase);
                    tc_fence_after();
                    if (elect_one()) {
                        uint32_t a_lo = ((ring + stage * stage_bytes) >> 4) | lo_fixed;      // smem < 256 KB: the field never overflows
                        for (int j = 0; j < nkb; ++j) {
#ifdef BYOLO_DBG_HOOKS
                            if (!(p.dbg & 4))
#endif
                            {
                                // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the (addr >> 4) field
                                const uint32_t b_lo = a_lo + as16;
                                umma_f16_lohi<CG>(d_tmem, a_lo, b_lo, desc_hi, idesc, (kb | j) != 0);
                                umma_f16_lohi<CG>(d_tmem, a_lo + 2, b_lo + 2, desc_hi, idesc, 1);
                                if (bk64) {
                                    umma_f16_lohi<CG>(d_tmem, a_lo + 4, b_lo + 4, desc_hi, idesc, 1);
                                    umma_f16_lohi<CG>(d_tmem, a_lo + 6, b_lo + 6, desc_hi, idesc, 1);
                                }
                            }
                            a_lo += kb16;
                        }
                        // one commit per stage: frees the smem slot (in both CTAs when CG = 2) once these MMAs retire.  A
                        // commit after only 4 MMAs leaves the tensor pipe idle ~200 cycles (tools/micro), hence kbs = 2.
                        if constexpr (CG == 1) umma_commit(empty0 + 8 * stage);
                        else umma_commit_2cta(empty0 + 8 * stage);
                    }
                    __syncwarp();
                    if (++stage == (uint32_t)num_stages) { stage = 0; phase ^= 1; }
                }
                // accumulator complete -> epilogue(s)
                if (elect_one()) {
                    if constexpr (CG == 1) umma_commit(acc_full0 + 8 * as);
                    else umma_commit_2cta(acc_full0 + 8 * as);
                }
                __syncwarp();
            }
          } else {
            // Split mode.  The tensor core adds into its fp32 accumulator with TRUNCATION (measured, tools/x3_probe.py: a
            // signed bias of -2.8e-9 * K relative, -1.3e-5 for a 3x3x512 conv - it compounds over 75 layers), so an
            // accumulator only ever takes one CHUNK of `chunk_stages` pipeline stages (24 MMAs); the epilogue warps drain
            // each chunk into round-to-nearest fp32 register sums while the next chunk runs in the other accumulator.
            const int chunk_stages = p.chunk_stages;
            uint32_t chunk_it = 0;
            for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
                for (int kb = 0; kb < num_kb; ++chunk_it) {
                    const uint32_t as = chunk_it & 3, aphase = (chunk_it >> 2) & 1;      // four chunk buffers of 128 columns:
                    mbar_wait(acc_empty0 + 8 * as, aphase ^ 1);                            // the issuer runs up to 3 chunks ahead of the drain
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + as * kChunkStride;
                    for (int cs = 0; cs < chunk_stages && kb < num_kb; ++cs, kb += kbs) {
                        const int nkb = min(kbs, num_kb - kb);
                        mbar_wait(full0 + 8 * stage, phase);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t a0 = ((ring + stage * stage_bytes) >> 4) | lo_fixed;
                            const int ksteps = bk64 ? 4 : 2;
                            // the small cross terms first, the hi * hi terms last: every add onto a large accumulator is one
                            // truncation at ITS ulp, whatever the size of the addend
                            uint32_t a_hi = a0;
                            for (int j = 0; j < nkb; ++j, a_hi += kb16) {
                                const uint32_t b_hi = a_hi + as16, a_lo = a_hi + a16, b_lo = b_hi + b16;      // planes: A_hi A_lo B_hi B_lo
                                for (int ks = 0; ks < ksteps; ++ks) {
                                    const uint32_t o = 2u * ks;      // 16 elements (32 B) along K inside the swizzle atom
                                    umma_f16_lohi<CG>(d_tmem, a_hi + o, b_lo + o, desc_hi, idesc, (cs | j | ks) != 0);   // hi * lo
                                    umma_f16_lohi<CG>(d_tmem, a_lo + o, b_hi + o, desc_hi, idesc, 1);                     // lo * hi
                                }
                            }
                            a_hi = a0;
                            for (int j = 0; j < nkb; ++j, a_hi += kb16) {
                                const uint32_t b_hi = a_hi + as16;
                                for (int ks = 0; ks < ksteps; ++ks) umma_f16_lohi<CG>(d_tmem, a_hi + 2u * ks, b_hi + 2u * ks, desc_hi, idesc, 1);   // hi * hi
                            }
                            if constexpr (CG == 1) umma_commit(empty0 + 8 * stage);
                            else umma_commit_2cta(empty0 + 8 * stage);
                        }
                        __syncwarp();
                        if (++stage == (uint32_t)num_stages) { stage = 0; phase ^= 1; }
                    }
                    if (elect_one()) {          // chunk complete -> the epilogue warps drain it
                        if constexpr (CG == 1) umma_commit(acc_full0 + 8 * as);
                        else umma_commit_2cta(acc_full0 + 8 * as);
                    }
                    __syncwarp();
                }
            }
          }
        }
    } else if (warp == 3 && p.bsplit) {
        // ================================ TMA producer, B operand (weights): same ring, same barriers =================
        const int BK = p.BK, BN = p.BN;
        const int n_rank = (int)rank * p.b_rows * (CG - 1);                  // CG = 2: my half of B
        const uint32_t b_bytes = p.b_bytes;
        const int b_lo_row = p.b_lo_row;
        uint32_t stage = 0, phase = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
            const int n0 = (tile & ((1 << nnt_shift) - 1)) * BN + n_rank;
            int b_k = 0;
            for (int kb = 0; kb < num_kb; kb += kbs) {
                const int nkb = min(kbs, num_kb - kb);
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                const uint32_t full = full0 + 8 * stage;
                const bool leader_lane = elect_one();
                uint32_t sb = ring + stage * stage_bytes + a_span;
#ifdef BYOLO_DBG_HOOKS
                if (p.dbg & 1) {
                    if (leader_lane && rank == 0) mbar_arrive(full);
                    __syncwarp();
                    if (++stage == (uint32_t)num_stages) { stage = 0; phase ^= 1; }
                    continue;
                }
#endif
                if (leader_lane && rank == 0) mbar_expect_tx(full, b_bytes * NPL * nkb * CG);
                for (int j = 0; j < nkb; ++j) {
                    if (leader_lane) {
                        tma_a2d<CG>(sb, &map_b, full, b_k, n0);
                        if constexpr (X3) tma_a2d<CG>(sb + b_bytes, &map_b, full, b_k, n0 + b_lo_row);
                    }
                    sb += kb_bytes;
                    b_k += BK;
                }
                __syncwarp();
                if (++stage == (uint32_t)num_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= kCtlWarps) {
        // ================================ epilogue ================================
        // 8 warps: TMEM lane quadrant = warp % 4 (hardware rule), the two warps of a quadrant interleave the column
        // chunks.  A chunk is 64 bytes of output per row (32 fp16 or 16 fp32 columns): TMEM -> registers,
        // [dropout] + shift + leaky [+ residual], convert, then
        //   * the warp's 32 x 64 B block goes to 64B-swizzled shared memory and out with one TMA store;
        //   * the two upsampling layers: 16-byte stores straight from registers (four destination pixels per row).
        // The residual of the next chunk is requested before the current one is processed (the first one before the
        // accumulator barrier), so its latency overlaps TMEM traffic and math.
        switch (p.epi_kind) {
            case EPI_F16: run_epilogue<CG, EW, EPI_F16, X3>(&map_o, p, ctl, tmem_base, out_stage, warp, lane, rank, first_tile, tile_step); break;
            case EPI_F16_RES: run_epilogue<CG, EW, EPI_F16_RES, X3>(&map_o, p, ctl, tmem_base, out_stage, warp, lane, rank, first_tile, tile_step); break;
            case EPI_F16_DROP: run_epilogue<CG, EW, EPI_F16_DROP, X3>(&map_o, p, ctl, tmem_base, out_stage, warp, lane, rank, first_tile, tile_step); break;
            case EPI_F16_DROP_T:
                if constexpr (AM == A_TILED)
                    run_epilogue<CG, EW, EPI_F16_DROP_T, X3>(&map_o, p, ctl, tmem_base, out_stage, warp, lane, rank, first_tile, tile_step);
                break;
            default:
                if constexpr (CG == 1 && AM == A_TILED) {
                    if (p.epi_kind == EPI_F32) run_epilogue<CG, EW, EPI_F32, X3>(&map_o, p, ctl, tmem_base, out_stage, warp, lane, rank, first_tile, tile_step);
                    else run_epilogue<CG, EW, EPI_UPSAMPLE, X3>(&map_o, p, ctl, tmem_base, out_stage, warp, lane, rank, first_tile, tile_step);
                }
                break;
        }
    }

    // ---------------- teardown ----------------
    tc_fence_before();
    __syncthreads();
    if (p.clk && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long g;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
        p.clk[2] = clock64();
        p.clk[3] = g;
    }
    if constexpr (CG == 2) cluster_sync_all();           // the peer may not retire while the leader's MMAs read its smem
    if (warp == 2) {
        if constexpr (CG == 1)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// --------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*,
                                   const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void* driver_fn(const char* name) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) return p;
    return nullptr;
}
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(driver_fn("cuTensorMapEncodeTiled"));
    return fn;
}
static EncodeIm2colFn encode_im2col_fn() {
    static EncodeIm2colFn fn = reinterpret_cast<EncodeIm2colFn>(driver_fn("cuTensorMapEncodeIm2col"));
    return fn;
}

static int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, const uint32_t* estr, int swizzle_bytes, bool f32 = false) {
    EncodeTiledFn fn = encode_fn();
    BY_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], e[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = estr[i]; }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    CUresult r = fn(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, e,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return -3;
    }
    return 0;
}

// im2col map over a dense fp16 NHWC tensor.  rank 4: dims {C, W, H, S}, corner = -pad / -pad (3x3, pad 1) or 0 / 0.
// rank 5: dims {C, W, H, T, B} with a ZERO byte stride on T (MC stacking), 1x1 only.  Semantics verified on the device with
// tools/micro/im2col_probe.cu: the unit walks `pixels` consecutive filter-window positions (W fastest, then H, then the
// outer dims), `tstride` apart, adds the per-instruction tap offset, and zero-fills whatever falls outside the tensor.
static int make_im2col_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, int corner,
                           int tstride, int channels, int pixels, int swizzle_bytes) {
    EncodeIm2colFn fn = encode_im2col_fn();
    BY_REQUIRE(fn != nullptr, "cuTensorMapEncodeIm2col not available from the driver");
    cuuint64_t d[5], s[4];
    cuuint32_t e[5] = {1, (cuuint32_t)tstride, (cuuint32_t)tstride, 1, 1};
    int lo[3] = {corner, corner, 0}, up[3] = {corner, corner, 0};
    for (int i = 0; i < rank; ++i) d[i] = dims[i];
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, s, lo, up, (cuuint32_t)channels,
                    (cuuint32_t)pixels, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeIm2col failed with CUresult " + std::to_string((int)r));
        return -3;
    }
    return 0;
}

static FastDiv make_fastdiv(uint32_t d) {
    FastDiv f{d, 0u, 0u};
    if (d <= 1) { f.d = 1; return f; }
    uint32_t lg = 0;
    while ((1ull << lg) < d) ++lg;                        // ceil(log2 d)
    const uint32_t pw = 31 + lg;
    f.mul = (uint32_t)(((1ull << pw) + d - 1) / d);
    f.shr = pw - 32;
    return f;
}

// Experiment switches (BYOLO_BN / BYOLO_CG / BYOLO_DBG / BYOLO_BSPLIT / BYOLO_KBS, profiles/r01/exp_v8_switches.txt) exist only
// in builds with -DBYOLO_DBG_HOOKS; the product library never reads the environment.
static int dbg_env(const char* name, int dflt = 0) {
#ifdef BYOLO_DBG_HOOKS
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
#else
    (void)name;
    return dflt;
#endif
}

static const void* kernel_variant(int cg, int amode, int x3) {
#define BYOLO_KV(CG_, AM_) (x3 ? (const void*)conv_umma_kernel<CG_, AM_, 8, true> : (const void*)conv_umma_kernel<CG_, AM_, 8, false>)
    if (cg == 2) {
        switch (amode) {
            case A_IM2COL: return BYOLO_KV(2, A_IM2COL);
            case A_STACK1: return BYOLO_KV(2, A_STACK1);
            case A_STACK2: return BYOLO_KV(2, A_STACK2);
            default: return BYOLO_KV(2, A_TILED);
        }
    }
    switch (amode) {
        case A_IM2COL: return BYOLO_KV(1, A_IM2COL);
        case A_STACK1: return BYOLO_KV(1, A_STACK1);
        case A_STACK2: return BYOLO_KV(1, A_STACK2);
        default: return BYOLO_KV(1, A_TILED);
    }
#undef BYOLO_KV
}

int umma_prepare(const ConvProblem& q, UmmaLaunch* L) {
    std::memset(L, 0, sizeof(*L));
    UmmaParams& p = L->p;
    const Geom& g = q.gin;
    const int C1 = g.C, C2 = q.in2 ? q.c2 : 0;
    const int t1 = std::max(q.t1, 1), t2 = std::max(q.t2, 1);
    BY_REQUIRE(q.k == 1 || q.k == 3, "kernel size must be 1 or 3 (layers.py:528)");
    BY_REQUIRE(q.stride == 1 || (q.stride == 2 && q.k == 3 && !q.in2), "stride 2 only as the 3x3 downsample conv");
    BY_REQUIRE(!(q.in2 && q.k != 1), "channel-concat input only for 1x1 convs");
    BY_REQUIRE(C1 % 32 == 0 && C2 % 32 == 0, "channel counts must be multiples of 32");
    BY_REQUIRE(q.cout_pad % 16 == 0 && q.cout_pad <= kMaxBias, "padded cout must be a multiple of 16 and <= 1024");
    BY_REQUIRE(g.H % q.stride == 0 && g.W % q.stride == 0, "stride-2 convs need even maps");
    BY_REQUIRE((t1 == 1 && t2 == 1) || q.k == 1, "MC-stacked sources only feed 1x1 convs");
    BY_REQUIRE(!(t1 > 1 && q.in2) && !(t2 > 1 && !q.in2), "stacked source: in1 alone, or in2 of a concat");
    BY_REQUIRE(g.S % t1 == 0 && g.S % t2 == 0, "sample count must be a multiple of the stacking factor");
    const int t_out = std::max(q.t_out, 1);            // > 1: T-invariant conv, every GEMM row is stored as t_out masked samples
    BY_REQUIRE(t_out == 1 || (q.k == 1 && !q.in2 && t1 == 1 && q.ep.drop.enabled && q.ep.drop.T == t_out && q.ep.out_mode == OUT_DENSE),
               "t_out: plain 1x1 dropout conv over un-stacked images only");
    p.BK = (C1 % 64 == 0 && C2 % 64 == 0) ? 64 : 32;
    p.taps = q.k * q.k;
    p.kb1 = C1 / p.BK;
    p.kb2 = C2 / p.BK;
    p.BN = std::min(q.cout_pad, 256);
    const bool x3 = q.x3 != 0;
    p.x3 = x3 ? 1 : 0;
    const int npl = x3 ? 2 : 1;                                   // operand planes: hi [, lo]
    // split mode: N tiles of <= 128 columns - each epilogue thread keeps round-to-nearest register sums of its <= 64 columns
    // (chunked accumulation, see the MMA issuer) and {A_hi, A_lo, B_hi, B_lo} stages stay <= 64 KB
    if (x3) p.BN = std::min(p.BN, 128);
    {
        // Wave quantisation: a layer runs ceil(tiles / resident tiles) rounds of tiles.  Small-M layers (19x19, 38x38 maps
        // of a 16-image batch) fill 1.24 / 2.46 rounds with 256-wide tiles; 128-wide tiles cost the same MMA rate
        // (tools/micro/mma_rate: 4093 MAC/cycle/SM at N = 128 and 256) and waste less of the last round.
        int dev_ = 0, sms_ = 148;
        if (cudaGetDevice(&dev_) == cudaSuccess) cudaDeviceGetAttribute(&sms_, cudaDevAttrMultiProcessorCount, dev_);
        const long long m_rows = (long long)g.S * (g.H / q.stride) * (g.W / q.stride);
        const bool pair = q.k == 3;                                                          // CTA pairs (decided below)
        const long long m_units = (m_rows + (pair ? 255 : 127)) / (pair ? 256 : 128);
        const int slots = pair ? sms_ / 2 : sms_;
        const int bn_env = dbg_env("BYOLO_BN");                                            // 256: never narrow the tile
        if (p.BN == 256 && q.cout_pad % 256 == 0 && bn_env != 256) {
            const long long r256 = (m_units * (q.cout_pad / 256) + slots - 1) / slots * 2;      // cost in 128-column units
            const long long r128 = (m_units * (q.cout_pad / 128) + slots - 1) / slots;
            if (r128 * 10 <= r256 * 9) p.BN = 128;       // only when a tenth of the rounds goes away (narrow tiles reload A twice as often)
        }
    }
    BY_REQUIRE(q.cout_pad % p.BN == 0, "cout_pad must be a multiple of the N tile");
    p.num_n_tiles = q.cout_pad / p.BN;
    BY_REQUIRE((p.num_n_tiles & (p.num_n_tiles - 1)) == 0, "the number of N tiles must be a power of two");
    for (p.nnt_shift = 0; (1 << p.nnt_shift) < p.num_n_tiles; ++p.nnt_shift) {}
    p.amode = q.k == 3 ? A_IM2COL : (t1 > 1 ? A_STACK1 : (t2 > 1 ? A_STACK2 : A_TILED));
    p.stride = q.stride;
    // CTA pairs for the feed-bound shapes: 3x3 stride-1 convs with wide N tiles
    const int cg_env = dbg_env("BYOLO_CG");                                          // 1 = force off, 2 = default policy
    // (measured per layer, profiles/r01/exp_v8_switches.txt: pairs win on every 3x3 conv, 5-23% fewer cycles)
    p.cg = (cg_env != 1 && q.k == 3 && p.BN >= 64) ? 2 : 1;
    // wide (N tile 256) 1x1 convs: pairs halve the weight traffic per SM: -10..-12% cycles on the 512/768 -> 256 and
    // 1024 -> 512 layers (exp_v8_switches.txt)
    if (cg_env != 1 && q.k == 1 && p.BN >= 256 && p.BK == 64 && q.ep.out_mode == OUT_DENSE) p.cg = 2;
    if (cg_env == 5 && q.k == 1 && p.BN >= 128 && p.BK == 64 && q.ep.out_mode == OUT_DENSE) p.cg = 2;      // experiment: N tile 128 too
    if (x3 && q.k == 1 && p.BN >= 128 && p.BK == 64 && q.ep.out_mode == OUT_DENSE) p.cg = 2;                // split mode: N tile is capped at 128
    p.b_rows = p.BN / p.cg;
    p.dbg = dbg_env("BYOLO_DBG");
    p.gout.S = g.S;
    p.gout.H = g.H / q.stride;
    p.gout.W = g.W / q.stride;
    p.gout.C = q.ep.cout;
    p.out_rows = (long long)p.gout.S * p.gout.H * p.gout.W;
    BY_REQUIRE(p.out_rows < (1ll << 31) && g.rows() < (1ll << 31), "activation map too large for 32-bit row indices");
    if (q.ep.out_mode != OUT_DENSE_F32) BY_REQUIRE(q.ep.cout % 32 == 0 && q.ep.ldc % 32 == 0, "fp16 outputs need cout % 32 == 0");
    else BY_REQUIRE(q.ep.ldc == q.cout_pad && q.stride == 1, "fp32 (detection) outputs are stored cout_pad wide, stride 1 only");
    p.ep = q.ep;
    p.ep.drop.thr16 = std::min<uint32_t>(p.ep.drop.thr16, 65535u);
    if (q.ep.out_mode == OUT_DENSE_F32) p.epi_kind = EPI_F32;
    else if (q.ep.out_mode == OUT_UPSAMPLE2) p.epi_kind = EPI_UPSAMPLE;
    else if (q.ep.drop.enabled) p.epi_kind = t_out > 1 ? EPI_F16_DROP_T : EPI_F16_DROP;
    else if (q.ep.residual) p.epi_kind = EPI_F16_RES;
    else p.epi_kind = EPI_F16;
    BY_REQUIRE((p.epi_kind == EPI_F32) == !q.ep.leaky, "fp16 outputs are conv+BN+leaky layers, the fp32 output is the linear detection conv");
    BY_REQUIRE(!(q.ep.drop.enabled && (q.ep.residual || (p.epi_kind != EPI_F16_DROP && p.epi_kind != EPI_F16_DROP_T))), "dropout only on plain convs");
    BY_REQUIRE(!(q.ep.residual && p.epi_kind != EPI_F16_RES), "residual only on plain convs");
    BY_REQUIRE(!((p.epi_kind == EPI_F32 || p.epi_kind == EPI_UPSAMPLE) && p.amode != A_TILED), "detection / upsampling convs are plain 1x1 convs");
    p.fd_plane = make_fastdiv((uint32_t)(p.gout.H * p.gout.W));
    p.fd_w = make_fastdiv((uint32_t)p.gout.W);
    p.fd_h = make_fastdiv((uint32_t)p.gout.H);
    // A_STACK*: sample -> (image, t) of the stacked source; dropout: sample -> (image, t) of the mask stream (same T)
    const int Tdiv = p.amode == A_STACK1 ? t1 : (p.amode == A_STACK2 ? t2 : std::max(q.ep.drop.T, 1));
    BY_REQUIRE(!(q.ep.drop.enabled && p.amode >= A_STACK1 && Tdiv != q.ep.drop.T), "stacking factor and dropout T differ");
    p.fd_T = make_fastdiv((uint32_t)Tdiv);
    const int swz = p.BK * 2;                                     // 128B or 64B rows
    p.sbo_bytes = 8 * swz;
    p.layout_type = (swz == 128) ? 2u : 4u;                       // SWIZZLE_128B / SWIZZLE_64B
    p.a_bytes = kTileM * swz;
    p.b_bytes = p.b_rows * swz;
    // instruction descriptor: D=f32, A=B=f16, both K-major, N>>3 at bit 17, M>>4 at bit 24 (M = 256 for a CTA pair)
    p.idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)((kTileM * p.cg) >> 4) << 24);
    const int bsplit_env = dbg_env("BYOLO_BSPLIT", 1);
    // 8 epilogue warps.  12 and 16 were measured: no gain / slower (16 caps the block at 96 registers per thread and the
    // extra staging blocks cost a pipeline stage), profiles/r01/exp_v8_switches.txt.
    p.epi_warps = 8;
    p.bsplit = bsplit_env;
    const int budget = 227 * 1024 - 1024 - (int)sizeof(SmemCtl) - p.epi_warps * kStageOutBytes - 64;
    const int kbs_env = dbg_env("BYOLO_KBS");
    const int kb_bytes = npl * (p.a_bytes + p.b_bytes);           // one K block: {A, B} or {A_hi, A_lo, B_hi, B_lo}
    const int num_kb = p.taps * (p.kb1 + p.kb2);
    // K blocks per stage: every stage costs one barrier round trip and one tcgen05.commit (~400 cycles of issue
    // overhead measured on the short layers), so a stage should carry >= 8 MMAs: 2 blocks of 64 channels or 4 of 32 -
    // as long as >= 3 stages still fit.
    p.kbs = 1;
    if (kbs_env >= 2) {                                   // experiment: force it wherever two stages still fit
        if (num_kb >= kbs_env && budget / (kbs_env * kb_bytes) >= 2) p.kbs = kbs_env;
        else if (num_kb >= 2 && budget / (2 * kb_bytes) >= 3) p.kbs = 2;
    } else if (kbs_env != 1) {
        // split mode issues 3 MMAs per K step: one 64-channel block (or two 32-channel blocks) already carries 12
        for (int cand = (p.BK == 64 ? 2 : 4) / npl; cand >= 2; cand /= 2)
            if (num_kb >= cand && budget / (cand * kb_bytes) >= 3) { p.kbs = cand; break; }
    }
    const int stage_bytes = p.kbs * kb_bytes;
    BY_REQUIRE(budget / stage_bytes >= 2, "conv tile does not fit two pipeline stages in shared memory");
    // split mode: 2 stages = 24 MMAs per accumulator chunk (per stage the 8 cross terms first, then the 4 hi * hi terms: 16
    // adds land on a large accumulator).  One stage per chunk halves the truncation bias again but makes the kernel
    // TMEM-read bound - every chunk is 64 KB of tcgen05.ld per CTA - 698 vs 787 img/s (profiles/r02/exp_x3_chunk.txt)
    p.chunk_stages = 2;
    if (dbg_env("BYOLO_CHUNK") > 0) p.chunk_stages = dbg_env("BYOLO_CHUNK");
    p.chunks_per_tile = ((num_kb + p.kbs - 1) / p.kbs + p.chunk_stages - 1) / p.chunk_stages;
    p.num_stages = std::min(kMaxStages, budget / stage_bytes);
    L->smem_bytes = p.num_stages * stage_bytes + 1024 + sizeof(SmemCtl) + p.epi_warps * kStageOutBytes + 64;

    const uint32_t one[5] = {1, 1, 1, 1, 1};
    p.num_m_tiles = (int)((p.out_rows + kTileM - 1) / kTileM);
    const uint32_t box_a[2] = {(uint32_t)p.BK, (uint32_t)kTileM};
    if (p.amode == A_IM2COL) {
        const uint64_t P1 = (uint64_t)C1 * npl;                    // halves per pixel: [hi C1 | lo C1] in split mode
        const uint64_t d[4] = {P1, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.S};
        const uint64_t st[3] = {P1 * 2, P1 * 2 * g.W, P1 * 2 * g.W * g.H};
        if (int e = make_im2col_map(&L->a1, q.in1, 4, d, st, -1, q.stride, p.BK, kTileM, swz)) return e;
        L->a2 = L->a1;
    } else {
        auto stacked = [&](CUtensorMap* m, const void* base, int C, int T) {
            const uint64_t P = (uint64_t)C * npl;
            const uint64_t d[5] = {P, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)T, (uint64_t)(g.S / T)};
            const uint64_t st[4] = {P * 2, P * 2 * g.W, 0ull, P * 2 * g.W * g.H};
            return make_im2col_map(m, base, 5, d, st, 0, 1, p.BK, kTileM, swz);
        };
        auto tiled = [&](CUtensorMap* m, const void* base, int C) {
            const uint64_t d[2] = {(uint64_t)C * npl, (uint64_t)g.rows()}, st[1] = {(uint64_t)C * npl * 2};
            return make_map(m, base, 2, d, st, box_a, one, swz);
        };
        if (p.amode == A_STACK1) { if (int e = stacked(&L->a1, q.in1, C1, t1)) return e; }
        else if (int e = tiled(&L->a1, q.in1, C1)) return e;
        if (!q.in2) L->a2 = L->a1;
        else if (p.amode == A_STACK2) { if (int e = stacked(&L->a2, q.in2, C2, t2)) return e; }
        else if (int e = tiled(&L->a2, q.in2, C2)) return e;
    }
    {
        const uint64_t K = (uint64_t)p.taps * (C1 + C2);
        uint64_t d[2] = {K, (uint64_t)q.cout_pad * npl}, s[1] = {K * 2};      // split mode: lo rows follow the cout_pad hi rows
        uint32_t box[2] = {(uint32_t)p.BK, (uint32_t)p.b_rows};
        if (int e = make_map(&L->b, q.w16, 2, d, s, box, one, swz)) return e;
    }
    L->o = L->b;
    if (p.epi_kind != EPI_UPSAMPLE) {
        // output map: [rows, ldc] of the dense output, 32 x 64 B boxes
        const bool f32 = p.epi_kind == EPI_F32;
        const uint64_t pitch = (uint64_t)q.ep.ldc * (f32 ? 1 : npl);
        uint64_t d[2] = {pitch, (uint64_t)p.out_rows * t_out}, st[1] = {pitch * (f32 ? 4 : 2)};      // t_out > 1: S * t_out output samples
        uint32_t box[2] = {(uint32_t)(f32 ? 16 : 32), 32u};
        if (int e = make_map(&L->o, q.ep.out, 2, d, st, box, one, 64, f32)) return e;
    }
    p.c1_lo = C1;
    p.c2_lo = C2;
    p.b_lo_row = q.cout_pad;
    p.num_tiles = ((p.num_m_tiles + p.cg - 1) / p.cg) * p.num_n_tiles;      // CG = 2: tiles are pairs of M tiles
    int dev = 0, sms = 0;
    BY_CUDA(cudaGetDevice(&dev));
    BY_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    L->grid = std::min(p.num_tiles * p.cg, sms / p.cg * p.cg);
    // the opt-in to > 48 KB of dynamic shared memory is a per-device attribute: done once for every device that is used
    static std::mutex attr_mutex;
    static bool attr_done[64] = {};
    cudaError_t attr_err = cudaSuccess;
    {
        std::lock_guard<std::mutex> lock(attr_mutex);
        if (dev < 0 || dev >= 64 || !attr_done[dev]) {
            for (int v = 0; v < 16 && attr_err == cudaSuccess; ++v)
                attr_err = cudaFuncSetAttribute(kernel_variant(v & 1 ? 2 : 1, (v >> 1) & 3, v >> 3), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (attr_err == cudaSuccess && dev >= 0 && dev < 64) attr_done[dev] = true;
        }
    }
    BY_CUDA(attr_err);
    return 0;
}

int umma_launch(const UmmaLaunch& L, cudaStream_t st) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(L.grid);
    cfg.blockDim = dim3((kCtlWarps + L.p.epi_warps) * 32);
    cfg.dynamicSmemBytes = L.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (L.p.cg == 2) {
        attr[1].id = cudaLaunchAttributeClusterDimension;
        attr[1].val.clusterDim.x = 2;
        attr[1].val.clusterDim.y = 1;
        attr[1].val.clusterDim.z = 1;
        cfg.numAttrs = 2;
    }
    void* args[5] = {(void*)&L.a1, (void*)&L.a2, (void*)&L.b, (void*)&L.o, (void*)&L.p};
    BY_CUDA(cudaLaunchKernelExC(&cfg, kernel_variant(L.p.cg, L.p.amode, L.p.x3), args));
    return 0;
}

int launch_conv_umma(const ConvProblem& q, cudaStream_t st) {
    UmmaLaunch L;
    if (int e = umma_prepare(q, &L)) return e;
    return umma_launch(L, st);
}

}  // namespace byolo
