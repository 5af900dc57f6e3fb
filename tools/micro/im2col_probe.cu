// Probe: semantics of TMA im2col loads (cuTensorMapEncodeIm2col + cp.async.bulk.tensor.4d...im2col) on sm_100a.
// Tensor NHWC fp16 [N=3, H=6, W=6, C=64], every channel of pixel (n,h,w) holds id = 1 + n*100 + h*10 + w (0 = OOB fill).
// Each case loads 32 pixels x 64 channels (SWIZZLE_128B) and prints the pixel id found in every shared-memory row.
//   nvcc -gencode arch=compute_100a,code=sm_100a im2col_probe.cu -o im2col_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*,
                                   const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kPix = 32;

__global__ void probe(const __grid_constant__ CUtensorMap map, int c, int w, int h, int n, int ow, int oh, float* out) {
    __shared__ __align__(1024) uint8_t tile[kPix * 128];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < kPix * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(tile)[i] = 0xFFFFFFFFu;   // NaN pattern
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(kPix * 128) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(
                smem_u32(tile)),
            "l"((uint64_t)&map), "r"(smem_u32(&bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"((uint16_t)ow), "h"((uint16_t)oh)
            : "memory");
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                : "=r"(done)
                : "r"(smem_u32(&bar))
                : "memory");
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kPix; i += blockDim.x) {
        // channel 0 lives in 16-byte chunk 0, stored at chunk (0 ^ (row & 7)) under SWIZZLE_128B
        const __half* p = reinterpret_cast<const __half*>(tile + i * 128 + ((i & 7) << 4));
        out[i] = __half2float(p[0]);
    }
}

__global__ void probe5(const __grid_constant__ CUtensorMap map, int c, int w, int h, int d, int n, float* out) {
    __shared__ __align__(1024) uint8_t tile[kPix * 128];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < kPix * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(tile)[i] = 0xFFFFFFFFu;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(kPix * 128) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.5d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2], {%8, %8, %8};" ::"r"(
                smem_u32(tile)),
            "l"((uint64_t)&map), "r"(smem_u32(&bar)), "r"(c), "r"(w), "r"(h), "r"(d), "r"(n), "h"((uint16_t)0)
            : "memory");
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                : "=r"(done)
                : "r"(smem_u32(&bar))
                : "memory");
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kPix; i += blockDim.x) {
        const __half* p = reinterpret_cast<const __half*>(tile + i * 128 + ((i & 7) << 4));
        out[i] = __half2float(p[0]);
    }
}

// compile-only: the cta_group::2 form of the im2col load (never launched)
__global__ void compile_check_cg2(const __grid_constant__ CUtensorMap map, uint32_t dst, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %3, %3, %3}], [%2], {%4, %4};" ::"r"(dst),
        "l"((uint64_t)&map), "r"(bar), "r"(0), "h"((uint16_t)0)
        : "memory");
}

int main() {
    const int N = 3, H = 6, W = 6, C = 64;
    std::vector<__half> host((size_t)N * H * W * C);
    for (int n = 0; n < N; ++n)
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w)
                for (int c = 0; c < C; ++c) host[(((size_t)n * H + h) * W + w) * C + c] = __float2half((float)(1 + n * 100 + h * 10 + w));
    __half* dev;
    cudaMalloc(&dev, host.size() * 2);
    cudaMemcpy(dev, host.data(), host.size() * 2, cudaMemcpyHostToDevice);
    float* out;
    cudaMalloc(&out, kPix * 4);
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) { printf("no entry point\n"); return 1; }
    EncodeIm2colFn enc = (EncodeIm2colFn)fp;

    struct Case { const char* name; int stride; int lo, up; int w, h, n, ow, oh; };
    const Case cases[] = {
        {"s1 start(-1,-1,0) off(0,0): top-left tap", 1, -1, -1, -1, -1, 0, 0, 0},
        {"s1 start(-1,-1,0) off(1,1): centre tap", 1, -1, -1, -1, -1, 0, 1, 1},
        {"s1 start(-1,-1,0) off(2,2): bottom-right tap", 1, -1, -1, -1, -1, 0, 2, 2},
        {"s1 start(2,3,1) off(2,1)", 1, -1, -1, 2, 3, 1, 2, 1},
        {"s1 start(0,0,0) off(0,0) lo=0 up=0 (1x1 conv)", 1, 0, 0, 0, 0, 0, 0, 0},
        {"s2 start(-1,-1,0) off(1,1)", 2, -1, -1, -1, -1, 0, 1, 1},
        {"s2 start(-1,-1,0) off(0,0)", 2, -1, -1, -1, -1, 0, 0, 0},
        {"s2 start(1,1,0) off(2,2)", 2, -1, -1, 1, 1, 0, 2, 2},
        {"s2 start(3,-1,0) off(1,1)", 2, -1, -1, 3, -1, 0, 1, 1},
    };
    for (const Case& cs : cases) {
        alignas(64) CUtensorMap map;
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        int lo[2] = {cs.lo, cs.lo}, up[2] = {cs.up, cs.up};
        cuuint32_t es[4] = {1, (cuuint32_t)cs.stride, (cuuint32_t)cs.stride, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, dev, dims, strides, lo, up, 64, kPix, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("== %s   (encode rc %d)\n", cs.name, (int)r);
        if (r != CUDA_SUCCESS) continue;
        cudaMemset(out, 0, kPix * 4);
        probe<<<1, 64>>>(map, 0, cs.w, cs.h, cs.n, cs.ow, cs.oh, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("   kernel error: %s\n", cudaGetErrorString(e)); return 2; }
        float res[kPix];
        cudaMemcpy(res, out, sizeof(res), cudaMemcpyDeviceToHost);
        for (int i = 0; i < kPix; ++i) {
            const int id = (int)res[i];
            if (res[i] != res[i]) printf("  NaN");
            else if (id == 0) printf("    .");
            else printf(" %d%d%d%d", 0, (id - 1) / 100, ((id - 1) / 10) % 10, (id - 1) % 10);      // 0 n h w
            if (i % 8 == 7) printf("\n");
        }
    }
    {   // 5D map with a ZERO stride on the D (= MC sample) axis: [B=2, T=3 (virtual), H=3, W=4, C=64] over the first 2 images
        alignas(64) CUtensorMap map;
        for (cuuint64_t tstride : {(cuuint64_t)0, (cuuint64_t)16}) {
            cuuint64_t dims[5] = {64, 4, 3, 3, 2};
            cuuint64_t strides[4] = {128, 4 * 128, tstride, (cuuint64_t)H * W * C * 2};
            int lo[3] = {0, 0, 0}, up[3] = {0, 0, 0};
            cuuint32_t es[5] = {1, 1, 1, 1, 1};
            CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, dev, dims, strides, lo, up, 64, kPix, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            printf("== 5D, T stride %d bytes, start (0,0,1,0) (encode rc %d)\n", (int)tstride, (int)r);
            if (r != CUDA_SUCCESS) continue;
            cudaMemset(out, 0, kPix * 4);
            probe5<<<1, 64>>>(map, 0, 0, 0, 1, 0, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("   kernel error: %s\n", cudaGetErrorString(e)); return 2; }
            float res[kPix];
            cudaMemcpy(res, out, sizeof(res), cudaMemcpyDeviceToHost);
            for (int i = 0; i < kPix; ++i) {
                const int id = (int)res[i];
                if (res[i] != res[i]) printf("  NaN");
                else if (id == 0) printf("    .");
                else printf(" %d%d%d%d", 0, (id - 1) / 100, ((id - 1) / 10) % 10, (id - 1) % 10);
                if (i % 8 == 7) printf("\n");
            }
        }
    }
    return 0;
}
