// CUDA-core kernels: the exact fp32 conv path (precision modes "fp32" / "fp16-simt", used to cross-check the
// tensor-core kernel layer by layer and to give bit-stable rows to the NMS end-to-end tests), the 3-channel stem
// conv (K1b) of those two modes, and fp32 <-> activation-type conversion for the test hooks.
// Same semantics as conv_umma.cu: conv -> [dropout] -> + BN shift -> leaky -> [+ residual]
// (/root/reference/lib_yolo/layers.py:545-575, :505-507, :521-524).
#include <algorithm>

#include "common.cuh"

namespace byolo {

template <typename T> __device__ __forceinline__ float ld_act(const T* p);
template <> __device__ __forceinline__ float ld_act<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_act<__half>(const __half* p) { return __half2float(__ldg(p)); }
template <typename T> __device__ __forceinline__ void st_act(T* p, float v);
template <> __device__ __forceinline__ void st_act<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_act<__half>(__half* p, float v) { *p = __float2half_rn(v); }

// one thread = one output pixel x 4 consecutive output channels
template <typename T>
__global__ void __launch_bounds__(256)
conv_simt_kernel(const T* __restrict__ in1, const T* __restrict__ in2, Geom g, int c2, int t1, int t2, int k, int stride, int cout_pad,
                 const float* __restrict__ w /*[K, cout_pad]*/, Epilogue ep, long long total) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int groups = cout_pad >> 2;
    const int cg = (int)(gid % groups);
    const long long pix = gid / groups;
    const int Ho = g.H / stride, Wo = g.W / stride;
    const int x = (int)(pix % Wo);
    const int y = (int)((pix / Wo) % Ho);
    const int s = (int)(pix / ((long long)Wo * Ho));
    const int c = cg * 4;
    const int C1 = g.C;
    const int pad = k / 2;         // 3x3: SAME at stride 1 (Appendix B-1), explicit pad 1 on all sides at stride 2 (layers.py:616-635)
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int kk = 0;
    for (int r = 0; r < k; ++r)
        for (int q = 0; q < k; ++q) {
            const int iy = y * stride + r - pad, ix = x * stride + q - pad;
            if (iy < 0 || iy >= g.H || ix < 0 || ix >= g.W) { kk += C1 + (in2 ? c2 : 0); continue; }      // zero padding
            const T* a1 = in1 + (((long long)(s / t1) * g.H + iy) * g.W + ix) * C1;
            for (int ci = 0; ci < C1; ++ci, ++kk) {
                const float a = ld_act<T>(a1 + ci);
                const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (long long)kk * cout_pad + c));
                acc[0] = fmaf(a, wv.x, acc[0]);
                acc[1] = fmaf(a, wv.y, acc[1]);
                acc[2] = fmaf(a, wv.z, acc[2]);
                acc[3] = fmaf(a, wv.w, acc[3]);
            }
            if (in2) {
                const T* a2 = in2 + (((long long)(s / t2) * g.H + iy) * g.W + ix) * c2;
                for (int ci = 0; ci < c2; ++ci, ++kk) {
                    const float a = ld_act<T>(a2 + ci);
                    const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (long long)kk * cout_pad + c));
                    acc[0] = fmaf(a, wv.x, acc[0]);
                    acc[1] = fmaf(a, wv.y, acc[1]);
                    acc[2] = fmaf(a, wv.z, acc[2]);
                    acc[3] = fmaf(a, wv.w, acc[3]);
                }
            }
        }
    if (c >= ep.cout) return;
    if (ep.drop.enabled) {
        float v8[8];
        const int c8 = c & ~7;
#pragma unroll
        for (int j = 0; j < 8; ++j) v8[j] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) v8[(c - c8) + j] = acc[j];
        const uint32_t e = (uint32_t)(y * Wo + x) * (uint32_t)ep.cout + (uint32_t)c8;
        dropout8(v8, ep.drop, e >> 3, s % ep.drop.T, ep.drop.image0 + s / ep.drop.T);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = v8[(c - c8) + j];
    }
    const long long opix = ((long long)s * Ho + y) * Wo + x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (c + j >= ep.cout) break;
        float v = acc[j] + __ldg(ep.bias + c + j);
        if (ep.leaky) v = fmaxf(v, 0.1f * v);
        if (ep.out_mode == OUT_DENSE_F32) {
            reinterpret_cast<float*>(ep.out)[opix * ep.ldc + c + j] = v;
            continue;
        }
        if (ep.residual) v += ld_act<T>(reinterpret_cast<const T*>(ep.residual) + opix * ep.ldc + c + j);
        T* ob = reinterpret_cast<T*>(ep.out);
        if (ep.out_mode == OUT_DENSE) {
            st_act<T>(ob + opix * ep.ldc + c + j, v);
        } else {
            for (int dy = 0; dy < 2; ++dy)
                for (int dx = 0; dx < 2; ++dx) {
                    const long long qq = ((long long)s * (2 * Ho) + (2 * y + dy)) * (2 * Wo) + (2 * x + dx);
                    st_act<T>(ob + qq * ep.ldc + c + j, v);
                }
        }
    }
}

int launch_conv_simt(const ConvProblem& p, bool act_half, cudaStream_t st) {
    BY_REQUIRE(p.cout_pad % 4 == 0, "cout_pad % 4");
    BY_REQUIRE((p.t1 <= 1 && p.t2 <= 1) || p.k == 1, "MC-stacked sources only feed 1x1 convs");
    const int Ho = p.gin.H / p.stride, Wo = p.gin.W / p.stride;
    const long long total = (long long)p.gin.S * Ho * Wo * (p.cout_pad / 4);
    const int grid = (int)((total + 255) / 256);
    if (act_half)
        conv_simt_kernel<__half><<<grid, 256, 0, st>>>((const __half*)p.in1, (const __half*)p.in2, p.gin, p.c2, std::max(p.t1, 1),
                                                       std::max(p.t2, 1), p.k, p.stride, p.cout_pad, p.w32, p.ep, total);
    else
        conv_simt_kernel<float><<<grid, 256, 0, st>>>((const float*)p.in1, (const float*)p.in2, p.gin, p.c2, std::max(p.t1, 1),
                                                      std::max(p.t2, 1), p.k, p.stride, p.cout_pad, p.w32, p.ep, total);
    BY_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// K1b: stem conv 3 -> 32, 3x3 stride 1 SAME (darknet.py:10), reading the fp32 image [B,H,W,3] in [0,1) directly
// (dataset_utils.py:6-11).  K = 27 is too shallow for the tensor pipe; the layer is bound by its 32-channel output.
// ------------------------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ void st_act16(T* p, const float* v);     // 16 consecutive channels
template <> __device__ __forceinline__ void st_act16<float>(float* p, const float* v) {
#pragma unroll
    for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(p)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
template <> __device__ __forceinline__ void st_act16<__half>(__half* p, const float* v) {
    uint4 o[2];
    __half2* h = reinterpret_cast<__half2*>(o);
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    reinterpret_cast<uint4*>(p)[0] = o[0];
    reinterpret_cast<uint4*>(p)[1] = o[1];
}

// One thread = two horizontally adjacent output pixels x 32 channels: 36 input values and 64 accumulators in
// registers, weights read as broadcast float4 from shared memory (8 FMAs per shared load), 128-byte stores.
// T = __half ("fp16-simt", the CUDA-core twin of the tensor-core path): image and weights are rounded to fp16 first,
// exactly the operands stem.cu feeds to the tensor cores.
template <typename T>
__global__ void __launch_bounds__(128)
stem_kernel(const float* __restrict__ img, int B, int H, int W, const float* __restrict__ w, const float* __restrict__ bias,
            T* __restrict__ out) {
    __shared__ __align__(16) float sw[27 * 32];
    __shared__ float sb[32];
    constexpr bool kRound = sizeof(T) == 2;
    for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) sw[i] = kRound ? __half2float(__float2half_rn(w[i])) : w[i];
    if (threadIdx.x < 32) sb[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const int W2 = W >> 1;
    const long long pid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pid >= (long long)B * H * W2) return;
    const int x = 2 * (int)(pid % W2), y = (int)((pid / W2) % H), b = (int)(pid / ((long long)W2 * H));
    float in[3][4][3];                                   // rows y-1..y+1, cols x-1..x+2
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int iy = y + r - 1;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int ix = x + q - 1;
            const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
            const float* p = img + (((long long)b * H + iy) * W + ix) * 3;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float v = ok ? __ldg(p + ch) : 0.f;
                in[r][q][ch] = kRound ? __half2float(__float2half_rn(v)) : v;
            }
        }
    }
    T* o = out + (((long long)b * H + y) * W + x) * 32;
#pragma unroll
    for (int half = 0; half < 2; ++half) {               // 16 output channels at a time keeps the register count down
        float a0[16], a1[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) { a0[c] = 0.f; a1[c] = 0.f; }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int q = 0; q < 3; ++q)
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const float v0 = in[r][q][ch], v1 = in[r][q + 1][ch];
                    const float4* wp = reinterpret_cast<const float4*>(sw + ((r * 3 + q) * 3 + ch) * 32 + half * 16);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 wv = wp[j];
                        a0[4 * j] = fmaf(v0, wv.x, a0[4 * j]);         a1[4 * j] = fmaf(v1, wv.x, a1[4 * j]);
                        a0[4 * j + 1] = fmaf(v0, wv.y, a0[4 * j + 1]); a1[4 * j + 1] = fmaf(v1, wv.y, a1[4 * j + 1]);
                        a0[4 * j + 2] = fmaf(v0, wv.z, a0[4 * j + 2]); a1[4 * j + 2] = fmaf(v1, wv.z, a1[4 * j + 2]);
                        a0[4 * j + 3] = fmaf(v0, wv.w, a0[4 * j + 3]); a1[4 * j + 3] = fmaf(v1, wv.w, a1[4 * j + 3]);
                    }
                }
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            a0[c] += sb[half * 16 + c]; a0[c] = fmaxf(a0[c], 0.1f * a0[c]);
            a1[c] += sb[half * 16 + c]; a1[c] = fmaxf(a1[c], 0.1f * a1[c]);
        }
        st_act16<T>(o + half * 16, a0);
        st_act16<T>(o + 32 + half * 16, a1);
    }
}

int launch_stem(const float* img, int B, int H, int W, const float* w32, const float* bias, void* out, bool act_half,
                cudaStream_t st) {
    BY_REQUIRE(W % 2 == 0, "stem: width must be even");
    const long long total = (long long)B * H * (W / 2);
    const int grid = (int)((total + 127) / 128);
    if (act_half)
        stem_kernel<__half><<<grid, 128, 0, st>>>(img, B, H, W, w32, bias, (__half*)out);
    else
        stem_kernel<float><<<grid, 128, 0, st>>>(img, B, H, W, w32, bias, (float*)out);
    BY_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// fp32 [S,H,W,C]  <->  activation type T, same dense layout   (test hooks / activation read-back only)
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void convert_kernel(const float* __restrict__ f32, T* __restrict__ act, long long total, int to_act) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        if (to_act)
            st_act<T>(act + i, f32[i]);
        else
            const_cast<float*>(f32)[i] = ld_act<T>(act + i);
    }
}

// split fp16: pixel = [hi C halves | lo C halves], value = hi + lo (both conversions exact to ~2^-22 relative)
__global__ void convert_hilo_kernel(const float* __restrict__ f32, __half* __restrict__ act, long long total, int C, int to_act) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long pix = i / C;
        const int c = (int)(i - pix * C);
        __half* hp = act + pix * 2 * C + c;
        if (to_act) {
            const float v = f32[i];
            const __half hi = __float2half_rn(v);
            hp[0] = hi;
            hp[C] = __float2half_rn(v - __half2float(hi));
        } else {
            const_cast<float*>(f32)[i] = __half2float(hp[0]) + __half2float(hp[C]);
        }
    }
}

static int convert(const float* dense, void* act, Geom g, int act_fmt, int to_act, cudaStream_t st) {
    const long long total = g.rows() * g.C;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    if (act_fmt == ACT_F16_HILO)
        convert_hilo_kernel<<<grid, 256, 0, st>>>(dense, (__half*)act, total, g.C, to_act);
    else if (act_fmt == ACT_F16)
        convert_kernel<__half><<<grid, 256, 0, st>>>(dense, (__half*)act, total, to_act);
    else
        convert_kernel<float><<<grid, 256, 0, st>>>(dense, (float*)act, total, to_act);
    BY_CUDA(cudaGetLastError());
    return 0;
}

int launch_pack(const float* dense, void* act, Geom g, int act_fmt, cudaStream_t st) { return convert(dense, act, g, act_fmt, 1, st); }

int launch_unpack(const void* act, float* dense, Geom g, int act_fmt, cudaStream_t st) {
    return convert(dense, const_cast<void*>(act), g, act_fmt, 0, st);
}

}  // namespace byolo
