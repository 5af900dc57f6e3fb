// Micro-benchmark 2: MMA ring (no data movement) for cta_group::1 and ::2 with different commit flavours.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61;
    return d;
}
#define WAIT(bar, par) asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(bar), "r"(par) : "memory")
// commit: 0 = cta_group::1 plain, 1 = cta_group::2 multicast to both, 2 = cta_group::2 leader only (no multicast)
template <int CG, int STAGES>
__global__ void __launch_bounds__(128, 1) ring(int N, int kblocks, int grp, int commit_kind, int pingpong, long long* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t full[STAGES], empty[STAGES], done;
    __shared__ uint32_t tmem_base;
    const uint32_t base = (smem_u32(sm) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm + (base - smem_u32(sm)))[i] = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[s])), "r"(1));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&done)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (CG == 1) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512) : "memory");
                       asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
        else { asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512) : "memory");
               asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
    if (warp == 0 && lane == 0 && pingpong && rank == 0) {          // producer (leader only; the peer would mirror it)
        for (int it = 0; it < kblocks; ++it) {
            const int st = it % STAGES, ph = (it / STAGES) & 1;
            WAIT(smem_u32(&empty[st]), ph ^ 1);
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[st])) : "memory");
        }
    } else if (warp == 1 && lane == 0 && rank == 0) {
        const uint64_t ad = make_desc(base), bd = make_desc(base + 16384);
        long long t0 = clock64();
        for (int it = 0; it < kblocks; ++it) {
            const int st = it % STAGES, ph = (it / STAGES) & 1;
            if (pingpong) { WAIT(smem_u32(&full[st]), ph); }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int k = 0; k < grp; ++k) {
                const uint64_t a = ad + 2 * (k & 3), b = bd + 2 * (k & 3);
                if (CG == 1) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tm), "l"(a), "l"(b), "r"(idesc), "r"(it | k) : "memory");
                else asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tm), "l"(a), "l"(b), "r"(idesc), "r"(it | k) : "memory");
            }
            const uint32_t eb = smem_u32(&empty[st]);
            if (commit_kind == 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(eb) : "memory");
            else if (commit_kind == 1) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(eb), "h"((uint16_t)3) : "memory");
            else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(eb) : "memory");
        }
        const uint32_t db = smem_u32(&done);
        if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(db) : "memory");
        else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(db) : "memory");
        WAIT(db, 0);
        long long t1 = clock64();
        out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    if (warp == 0) {
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
    }
}
template <int CG, int STAGES> void run(int N, int grp, int commit_kind, int pingpong, int grid) {
    long long* d; cudaMalloc(&d, 16); cudaMemset(d, 0, 16);
    const int kblocks = 4096 / grp * 4, smem = 50 * 1024 + 1024;
    cudaFuncSetAttribute(ring<CG, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    double best = 1e30;
    for (int rep = 0; rep < 3; ++rep) {
        cudaLaunchKernelEx(&cfg, ring<CG, STAGES>, N, kblocks, grp, commit_kind, pingpong, d);
        cudaDeviceSynchronize();
        long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        best = h / (double)(kblocks * grp) < best ? h / (double)(kblocks * grp) : best;
    }
    printf("CG=%d stages=%d N=%d grp=%2d commit=%d pingpong=%d grid=%3d : %.1f cycles/MMA (%s)\n", CG, STAGES, N, grp, commit_kind, pingpong, grid, best, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}
int main() {
    for (int pp : {0, 1}) for (int grp : {4, 8, 16}) { run<1, 4>(256, grp, 0, pp, 1); run<2, 4>(256, grp, 1, pp, 2); run<2, 4>(256, grp, 2, pp, 2); }
    for (int grp : {8, 16}) { run<1, 3>(256, grp, 0, 1, 148); run<2, 3>(256, grp, 1, 1, 148); run<2, 3>(256, grp, 2, 1, 148); }
    return 0;
}
