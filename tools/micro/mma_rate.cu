// Micro-benchmark: cycles per tcgen05.mma (kind::f16, SS operands from shared memory) with no data movement,
// for cta_group::1 (M=128) and cta_group::2 (M=256), N in {64,128,256}.   nvcc -arch=sm_100a mma_rate.cu -o mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
template <int CG>
__global__ void __launch_bounds__(128, 1) k(int N, int iters, long long* out, int concurrent_ld) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const uint32_t base = (smem_u32(sm) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm + (base - smem_u32(sm)))[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
    if (warp == 1 && lane == 0 && rank == 0) {
        const uint64_t ad = make_desc(base), bd = make_desc(base + 16384);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint64_t a = ad + 2 * (i & 3), b = bd + 2 * (i & 3);
            if (CG == 1)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tm), "l"(a), "l"(b), "r"(idesc), "r"(i) : "memory");
            else
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tm), "l"(a), "l"(b), "r"(idesc), "r"(i) : "memory");
        }
        if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
        asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    } else if (warp >= 2 && concurrent_ld) {
        // epilogue-like TMEM reads running concurrently with the MMAs (second accumulator half)
        uint32_t v[32];
        uint32_t acc = 0;
        for (int i = 0; i < iters / 8; ++i) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(tm + 256 + ((uint32_t)((warp & 3) * 32) << 16) + (i & 7) * 32));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += v[0] + v[31];
        }
        if (acc == 0x12345) out[1] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    if (warp == 0) {
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
    }
}

// mode 1: commit after every `grp` MMAs to a ring of barriers nobody waits on.
// mode 2: full producer/consumer ping-pong like conv_umma (producer thread: wait empty -> arrive full; MMA thread: wait full -> grp MMAs -> commit empty)
__global__ void __launch_bounds__(128, 1) ring(int N, int kblocks, int grp, int stages, int mode, int fence, long long* out, int pair = 1) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t full[8], empty[8];
    __shared__ uint32_t tmem_base;
    const uint32_t base = (smem_u32(sm) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm + (base - smem_u32(sm)))[i] = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 8; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[s])), "r"(1));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
#define WAIT(bar, par) asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(bar), "r"(par) : "memory")
    if (warp == 0 && lane == 0 && mode == 2) {          // producer
        for (int it = 0; it < kblocks; ++it) {
            const int st = it % stages, ph = (it / stages) & 1;
            WAIT(smem_u32(&empty[st]), ph ^ 1);
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[st])) : "memory");
        }
    } else if (warp == 1 && lane == 0) {
        const uint64_t ad = make_desc(base), bd = make_desc(base + 16384);
        long long t0 = clock64();
        for (int it = 0; it < kblocks; ++it) {
            const int st = it % stages, ph = (it / stages) & 1;
            if (mode == 2) { WAIT(smem_u32(&full[st]), ph); }
            if (fence) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int k = 0; k < grp; ++k) {
                const uint64_t a = ad + 2 * (k & 3), b = bd + 2 * (k & 3);
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tm), "l"(a), "l"(b), "r"(idesc), "r"(it | k) : "memory");
            }
            if ((it % pair) == pair - 1)
                for (int q = pair - 1; q >= 0; --q)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[(it - q) % stages])) : "memory");
        }
        // drain: wait for the last commit
        {
            const int it = kblocks - 1; const int st = it % stages, ph = (it / stages) & 1;
            if (mode == 2) { /* producer consumed phases; wait on own view */ }
            long long spin = 0;
            (void)spin; (void)st; (void)ph;
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&full[7])) : "memory");
        if (mode == 1 || stages < 8) { WAIT(smem_u32(&full[7]), 0); }
        long long t1 = clock64();
        out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}
void run_ring(int N, int grp, int stages, int mode, int fence, int pair = 1) {
    long long* d; cudaMalloc(&d, 16); cudaMemset(d, 0, 16);
    const int kblocks = 2048, smem = 50 * 1024 + 1024;
    cudaFuncSetAttribute(ring, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; ++rep) ring<<<1, 128, smem>>>(N, kblocks, grp, stages, mode, fence, d, pair);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("ring N=%d grp=%d stages=%d mode=%d fence=%d pair=%d : %.1f cycles/MMA (%s)\n", N, grp, stages, mode, fence, pair, (double)h / (kblocks * grp), cudaGetErrorString(e));
    cudaFree(d);
}

template <int CG> void run(int N, int grid, int ld) {
    long long* d; cudaMalloc(&d, 16); cudaMemset(d, 0, 16);
    const int iters = 8192, smem = 50 * 1024 + 1024;
    cudaFuncSetAttribute(k<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, k<CG>, N, iters, d, ld);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("CG=%d N=%3d grid=%3d tmem_ld=%d : %.1f cycles/MMA  (%s)  -> %.0f MACs/cycle/SM\n", CG, N, grid, ld, (double)h / iters, cudaGetErrorString(e),
           128.0 * N * 16 / ((double)h / iters));
    cudaFree(d);
}
int main() {
    for (int N : {64, 128, 256}) { run<1>(N, 1, 0); run<2>(N, 2, 0); }
    run<1>(256, 148, 0); run<2>(256, 148, 0);
    run<1>(256, 148, 1); run<2>(256, 148, 1);
    for (int mode : {1, 2}) for (int grp : {4, 8}) for (int st : {4, 6}) for (int f : {0, 1}) run_ring(256, grp, st, mode, f);
    run_ring(256, 1, 4, 2, 1); run_ring(256, 2, 4, 2, 1); run_ring(256, 16, 4, 2, 1);
    for (int mode : {1, 2}) for (int st : {4, 6, 8}) for (int pair : {2, 3, 4}) if (pair < st) run_ring(256, 4, st, mode, 1, pair);
    for (int N : {128, 64}) for (int grp : {4, 8, 16}) run_ring(N, grp, 6, 2, 1, 1);
    run_ring(128, 4, 8, 2, 1, 2); run_ring(128, 4, 8, 2, 1, 4); run_ring(64, 4, 8, 2, 1, 4);
    return 0;
}
