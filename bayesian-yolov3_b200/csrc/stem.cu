// K1b (product path): stem conv 3 -> 32, 3x3 stride 1 SAME + BN shift + leaky (darknet.py:10 -> layers.py:545-575),
// reading the fp32 image [B,H,W,3] in [0,1) (dataset_utils.py:6-11) and writing the dense NHWC fp16 map.
//
// K = 27 is far too shallow for a tcgen05 pipeline (one 128x32x32 tile per 128 pixels, nothing to overlap), and the
// layer is bound by its 64 B/pixel output, so the contraction runs on warp-level mma.sync m16n8k16 (fp16 operands,
// fp32 accumulate - the same rounding points as every other conv of the fp16 path) and the work goes into the data
// movement: a CTA stages an 18 x 34 pixel halo tile as fp16 in shared memory (patch row r of pixel x is then 9
// consecutive halves, so K index k = r*9 + (q*3 + ch) is exactly the engine's weight order), each warp produces
// 16-pixel x 32-channel tiles whose 1024 output bytes are contiguous in HBM and leave as 16-byte stores.
// Algorithmic bytes: 12 B read + 64 B written per pixel.
#include <algorithm>

#include "common.cuh"

namespace byolo {

namespace {
constexpr int kTW = 32, kTH = 16;            // output tile; image sizes are multiples of 32 (yolov3.py:207-211)
constexpr int kInW = (kTW + 2) * 3;          // 102 halves of input per staged row
constexpr int kPitch = 104;                  // halves
constexpr int kOutPitch = 40;                // halves per staged output row (80 B: conflict-free fragment stores)
constexpr int kWarps = 8;

__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(__half lo, __half hi) {
    return (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
}
}  // namespace

// X3 (split-fp16 mode): image and weights are hi + lo fp16 pairs, three MMAs (hi*hi + hi*lo + lo*hi) per K step, and the
// output pixel is stored as [hi 32 halves | lo 32 halves].
template <bool X3>
__global__ void __launch_bounds__(kWarps * 32, X3 ? 2 : 4)
stem_mma_kernel(const float* __restrict__ img, int H, int W, int num_tiles, const __half* __restrict__ w16 /*[32 (+32 lo)][27]*/,
                const float* __restrict__ bias, __half* __restrict__ out) {
    constexpr int NP = X3 ? 2 : 1;
    __shared__ __align__(16) __half sin[NP][(kTH + 2) * kPitch];
    __shared__ __align__(16) __half sout[kWarps][NP][16 * kOutPitch];
    const int tiles_x = W / kTW, tiles_y = H / kTH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;

    // B fragments (weights): b[s][nt] = {W[n = nt*8+g][k = s*16 + 2t, +1], W[n][k + 8, + 9]}, zero for k >= 27
    uint32_t bf[NP][2][4][2];
#pragma unroll
    for (int pl = 0; pl < NP; ++pl)
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const int k = s * 16 + p * 8 + 2 * t, n = nt * 8 + g + 32 * pl;
                const __half lo = k < 27 ? w16[n * 27 + k] : __float2half(0.f);
                const __half hi = k + 1 < 27 ? w16[n * 27 + k + 1] : __float2half(0.f);
                bf[pl][s][nt][p] = pack2(lo, hi);
            }
    float bs[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        bs[nt][0] = __ldg(bias + nt * 8 + 2 * t);
        bs[nt][1] = __ldg(bias + nt * 8 + 2 * t + 1);
    }
    // offsets (in halves, relative to the pixel's patch origin) of this thread's 8 K indices
    int koff[2][2][2];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = s * 16 + p * 8 + 2 * t + e;
                koff[s][p][e] = k < 27 ? (k / 9) * kPitch + (k % 9) : 0;      // k >= 27 meets a zero weight
            }

  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
    const int x0 = tx * kTW, y0 = ty * kTH;
    // stage the halo tile: rows y0-1 .. y0+16, columns (x0-1)*3 .. (x0+33)*3, zero outside the image (SAME padding)
    // all of a thread's loads are issued before the first conversion (one load in flight per thread left the kernel at
    // 0.39 of the HBM rate: F2FP waiting on the scoreboard, profiles/r01/ncu_v7_single_stem.txt)
    constexpr int kStageIters = ((kTH + 2) * kPitch + kWarps * 32 - 1) / (kWarps * 32);
    float sv[kStageIters];
#pragma unroll
    for (int it = 0; it < kStageIters; ++it) {
        const int i = threadIdx.x + it * kWarps * 32;
        const int r = i / kPitch, c = i - r * kPitch;
        const int iy = y0 - 1 + r, ix = x0 - 1 + c / 3;
        sv[it] = 0.f;
        if (i < (kTH + 2) * kPitch && c < kInW && iy >= 0 && iy < H && ix >= 0 && ix < W)
            sv[it] = __ldg(img + ((long long)(b * H + iy) * W + (x0 - 1)) * 3 + c);
    }
#pragma unroll
    for (int it = 0; it < kStageIters; ++it) {
        const int i = threadIdx.x + it * kWarps * 32;
        if (i < (kTH + 2) * kPitch) {
            const __half hi = __float2half_rn(sv[it]);
            sin[0][i] = hi;
            if constexpr (X3) sin[1][i] = __float2half_rn(sv[it] - __half2float(hi));
        }
    }
    __syncthreads();

#pragma unroll 1
    for (int mt = 0; mt < 4; ++mt) {                     // warp: rows 2*warp, 2*warp+1; two 16-pixel tiles per row
        const int ly = 2 * warp + (mt >> 1), lx = (mt & 1) * 16;
        uint32_t a[NP][2][4];
#pragma unroll
        for (int pl = 0; pl < NP; ++pl) {
            const __half* p0 = sin[pl] + ly * kPitch + (lx + g) * 3;      // pixel g of the tile; pixel g+8 is 24 halves further
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    a[pl][s][2 * p] = pack2(p0[koff[s][p][0]], p0[koff[s][p][1]]);
                    a[pl][s][2 * p + 1] = pack2(p0[24 + koff[s][p][0]], p0[24 + koff[s][p][1]]);
                }
        }
        float acc[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[nt][j] = 0.f;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                mma16816(acc[nt], a[0][s], bf[0][s][nt][0], bf[0][s][nt][1]);
                if constexpr (X3) {
                    mma16816(acc[nt], a[0][s], bf[1][s][nt][0], bf[1][s][nt][1]);     // hi * lo
                    mma16816(acc[nt], a[1][s], bf[0][s][nt][0], bf[0][s][nt][1]);     // lo * hi
                }
            }
        }
        __syncwarp();                                    // the previous tile has left the staging block
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            float v0 = acc[nt][0] + bs[nt][0], v1 = acc[nt][1] + bs[nt][1];
            float v2 = acc[nt][2] + bs[nt][0], v3 = acc[nt][3] + bs[nt][1];
            v0 = fmaxf(v0, 0.1f * v0); v1 = fmaxf(v1, 0.1f * v1);
            v2 = fmaxf(v2, 0.1f * v2); v3 = fmaxf(v3, 0.1f * v3);
            const __half2 h01 = __floats2half2_rn(v0, v1), h23 = __floats2half2_rn(v2, v3);
            *reinterpret_cast<__half2*>(sout[warp][0] + g * kOutPitch + nt * 8 + 2 * t) = h01;
            *reinterpret_cast<__half2*>(sout[warp][0] + (g + 8) * kOutPitch + nt * 8 + 2 * t) = h23;
            if constexpr (X3) {
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                *reinterpret_cast<__half2*>(sout[warp][1] + g * kOutPitch + nt * 8 + 2 * t) = __floats2half2_rn(v0 - f01.x, v1 - f01.y);
                *reinterpret_cast<__half2*>(sout[warp][1] + (g + 8) * kOutPitch + nt * 8 + 2 * t) = __floats2half2_rn(v2 - f23.x, v3 - f23.y);
            }
        }
        __syncwarp();
        // 16 pixels x 64 B (x2 in split mode: [hi | lo] per pixel), contiguous in the output row
        __half* o = out + (((long long)b * H + (y0 + ly)) * W + (x0 + lx)) * (32 * NP);
#pragma unroll
        for (int j = 0; j < 2 * NP; ++j) {
            const int idx = lane + 32 * j, row = idx / (4 * NP), q = idx % (4 * NP), pl = q >> 2;
            const uint4 v = *reinterpret_cast<const uint4*>(sout[warp][pl] + row * kOutPitch + (q & 3) * 8);
            *reinterpret_cast<uint4*>(o + row * 32 * NP + q * 8) = v;
        }
    }
    __syncthreads();                                     // everyone is done with the halo tile before it is overwritten
  }
}

int launch_stem_mma(const float* img, int B, int H, int W, const __half* w16, const float* bias, void* out, bool x3, cudaStream_t st) {
    BY_REQUIRE(H % kTH == 0 && W % kTW == 0, "stem: image size must be a multiple of 32");
    const long long tiles = (long long)B * (H / kTH) * (W / kTW);
    BY_REQUIRE(tiles < (1ll << 31), "stem: too many tiles");
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)std::min<long long>(tiles, (long long)sms * (x3 ? 2 : 4));   // resident CTAs per SM, each walks its tiles
    if (x3) stem_mma_kernel<true><<<grid, kWarps * 32, 0, st>>>(img, H, W, (int)tiles, w16, bias, (__half*)out);
    else stem_mma_kernel<false><<<grid, kWarps * 32, 0, st>>>(img, H, W, (int)tiles, w16, bias, (__half*)out);
    BY_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace byolo
