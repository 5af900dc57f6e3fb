// Shared declarations of libbyolo (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace byolo {

// ----------------------------------------------------------------------------------------------------------
// Error plumbing: every C-ABI entry returns 0 or a negative code; the message is kept per thread.
// ----------------------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
#define BY_CUDA(expr)                                                                                      \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess) {                                                                           \
            ::byolo::set_error(std::string(#expr) + " -> " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                               std::to_string(__LINE__) + ")");                                            \
            return -2;                                                                                     \
        }                                                                                                  \
    } while (0)
#define BY_REQUIRE(cond, msg)                                                                              \
    do {                                                                                                   \
        if (!(cond)) {                                                                                     \
            ::byolo::set_error(std::string(msg) + " [" #cond "] (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
            return -1;                                                                                     \
        }                                                                                                  \
    } while (0)

// ----------------------------------------------------------------------------------------------------------
// Activation layout: dense NHWC.  A map of S samples x H x W x C is the row-major matrix [S*H*W, C]; GEMM row m is
// pixel (s, y, x) = (m / (H*W), (m / W) % H, m % W).  SAME / explicit padding is never materialised: the tensor-core
// path reads its A operand through TMA im2col tensor maps (out-of-image taps are zero-filled by the hardware), the
// CUDA-core path tests the bounds.
// ----------------------------------------------------------------------------------------------------------
struct Geom {
    int S, H, W, C;
    __host__ __device__ long long rows() const { return (long long)S * H * W; }
};

enum OutMode : int {
    OUT_DENSE = 0,       // T   [S,Ho,Wo,ldc]
    OUT_DENSE_F32 = 1,   // f32 [S,Ho,Wo,ldc]     detection conv -> raw head output, ldc = cout padded to 16
    OUT_UPSAMPLE2 = 2,   // T   [S,2Ho,2Wo,ldc]   nearest x2 (layers.py:578-580) fused into the store
};

struct Dropout {
    int enabled;             // 0/1
    uint32_t seed_lo, seed_hi;
    int layer_id;            // 0..14
    int T;                   // samples per image: sample s = image*T + t
    int image0;              // global index of image 0 of this batch (rank offset)
    uint32_t thr16;          // keep iff r16 >= thr16
    float keep_scale;        // 1/(1-p)
};

// Epilogue description shared by the tensor-core and the CUDA-core conv kernels.
struct Epilogue {
    const float* bias;       // [cout_pad]  BN shift (scale is folded into the weights) or detection bias
    const void* residual;    // T, same geometry as the output, or nullptr
    void* out;
    int out_mode;
    int ldc;                 // channels of the output buffer
    int cout;                // valid output channels
    int leaky;               // 1: max(x, 0.1x)
    Dropout drop;
};

// ----------------------------------------------------------------------------------------------------------
// Philox4x32-10 dropout stream (specification: oracle/philox.py).
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

// Applies the mask of 8 consecutive channels (element indices 8*group .. 8*group+7) to x[0..7].
__device__ __forceinline__ void dropout8(float* x, const Dropout& d, uint32_t group, int t, int image) {
    const uint4 r = philox4x32_10(make_uint4(group, (uint32_t)d.layer_id, (uint32_t)t, (uint32_t)image), d.seed_lo,
                                  d.seed_hi);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        x[2 * i] = ((w[i] & 0xFFFFu) >= d.thr16) ? x[2 * i] * d.keep_scale : 0.f;
        x[2 * i + 1] = ((w[i] >> 16) >= d.thr16) ? x[2 * i + 1] * d.keep_scale : 0.f;
    }
}

// ----------------------------------------------------------------------------------------------------------
// Kernel launchers (definitions in the respective .cu files). All return 0 / negative error.
// ----------------------------------------------------------------------------------------------------------
struct ConvProblem {
    // input(s): dense half/float buffers
    const void* in1;
    const void* in2;         // second K-slice of a 1x1 conv over a channel concat [in1, in2] (layers.py:583-592) or null
    Geom gin;                // geometry of in1 AS THE CONV SEES IT (S = samples of the output; C = channels of in1)
    int c2;                  // channels of in2
    int t1, t2;              // MC stacking (layers.py:595-597) without a copy: sample s of the conv reads sample s / t1 of
                             // in1 (s / t2 of in2); 1 = the buffer holds every sample itself.  1x1 convs only.
    int t_out;               // > 1 (tensor-core path only): gin.S counts IMAGES and every output row is stored t_out times, once per MC
                             // sample with that sample's dropout mask - the conv over a stacked input without the T-fold GEMM
    int k, stride;           // 1|3, 1|2
    int cout_pad;            // rows of the weight matrix (multiple of 16)
    const __half* w16;       // [cout_pad, K] K-major, K = k*k*(C1+C2) ordered (tap, channel); BN scale folded
    const float* w32;        // [K, cout_pad] fp32, same folding (CUDA-core path)
    int x3;                  // split-fp16 mode (BYOLO_PREC_FP16X3): activations [rows][hi C | lo C], w16 = [2 cout_pad, K] (hi rows, lo rows)
    Epilogue ep;
};

int launch_conv_umma(const ConvProblem& p, cudaStream_t st);                 // conv_umma.cu (fp16 operands, tcgen05)
int launch_conv_simt(const ConvProblem& p, bool act_half, cudaStream_t st);  // conv_simt.cu
int launch_stem(const float* img, int B, int H, int W, const float* w32 /*[27,32]*/, const float* bias,
                void* out, bool act_half, cudaStream_t st);                  // conv_simt.cu
int launch_stem_mma(const float* img, int B, int H, int W, const __half* w16 /*[32 (x2: hi, lo)][27]*/, const float* bias, void* out,
                    bool x3, cudaStream_t st);                               // stem.cu (fp16 operands, mma.sync)
// activation storage formats: fp32 | fp16 | split fp16 (pixel = [hi C halves | lo C halves], value = hi + lo)
enum ActFmt : int { ACT_F32 = 0, ACT_F16 = 1, ACT_F16_HILO = 2 };
int launch_pack(const float* dense, void* act, Geom g, int act_fmt, cudaStream_t st);
int launch_unpack(const void* act, float* dense, Geom g, int act_fmt, cudaStream_t st);

struct DecodeProblem {
    int variant;             // 0 standard, 1 aleatoric, 2 epistemic
    int B, T;                // images, samples per image (1 unless epistemic)
    int cls_cnt;
    int gh[3], gw[3];        // grids, stride 32/16/8
    const float* raw[3];     // [B*T, gh(+2), gw(+2), ld] fp32; padded = 1: one-pixel border around each map (unused by the engine)
    int ld[3];
    int padded;
    float prior_h[9], prior_w[9];
    float* rows;             // [B, N, D]
    int N, D;
};
int launch_decode(const DecodeProblem& p, cudaStream_t st);                  // decode.cu

struct NmsOptions {
    int packed = 0;          // 1: out_rows is [B, max_out + 1, D]; row max_out of every image = (count, 0, ...)
    int force_cs = 0;        // test hook: cluster size 1 | 2 | 4 | 8 instead of the occupancy-based choice
    int force_chunked = 0;   // test hook: take the N > 32768 path for any N
};
int launch_nms(const float* rows, int B, int N, int D, int obj_idx, float iou_thr, int max_out, float* out_rows,
               int* out_idx, int* out_count, const NmsOptions& opt, cudaStream_t st);  // nms.cu
int launch_class_filter(const float* rows, long long total_rows, int D, int obj_idx, int cls_start, int cls_cnt, int cls, float* out,
                        cudaStream_t st);                                      // nms.cu

}  // namespace byolo
