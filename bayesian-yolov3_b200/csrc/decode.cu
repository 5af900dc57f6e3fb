// K2: head-output split + per-sample transforms + reduction over the T Monte-Carlo samples + anchor decode, writing
// each row at its final concat_bbox position (staged in shared memory, stored coalesced).  One thread per (image, anchor).
//   split              /root/reference/lib_yolo/layers.py:11-84
//   standard decode    layers.py:191-258      row = [y0,x0,y1,x1, obj, cls..]
//   aleatoric decode   layers.py:261-346      row = [y0,x0,y1,x1, var*4, prod var, obj, H(obj), cls.., H(cls), layer, prior]
//   epistemic stats    layers.py:361-411, decode layers.py:414-502
//                      row = [y0,x0,y1,x1, diag cov*4, ale var*4, det cov, sum ale, obj, MI, H, cls.., MI, H, layer, prior]
//   concat order       inference_epistemic.py:173-184: scale (32,16,8) -> prior -> row -> col
// fp32 throughout, operations in the order the reference graph applies them; compiled with -fmad=false so the
// covariance E[xx^T] - E[x]E[x]^T and the entropies round like separate TF ops.  p*log(p) at p in {0,1} yields
// NaN exactly as the reference does (layers.py:349-358).
#include "common.cuh"

namespace byolo {

constexpr int kMaxCls = 16;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float logistic_entropy(float s) { return -((1.f - s) * logf(1.f - s) + s * logf(s)); }

__device__ __forceinline__ void softmax_(const float* z, int n, float* out) {
    float m = z[0];
    for (int i = 1; i < n; ++i) m = fmaxf(m, z[i]);
    float sum = 0.f;
    for (int i = 0; i < n; ++i) { out[i] = expf(z[i] - m); sum += out[i]; }
    for (int i = 0; i < n; ++i) out[i] = out[i] / sum;
}
__device__ __forceinline__ float softmax_entropy(const float* s, int n) {
    float acc = 0.f;
    for (int i = 0; i < n; ++i) acc += s[i] * logf(s[i]);
    return -acc;
}

// determinant of a 4x4 by LU with partial pivoting (what tf.linalg.det / LAPACK getrf do)
__device__ float det4(float a[4][4]) {
    float det = 1.f;
    for (int k = 0; k < 4; ++k) {
        int piv = k;
        float best = fabsf(a[k][k]);
        for (int i = k + 1; i < 4; ++i)
            if (fabsf(a[i][k]) > best) { best = fabsf(a[i][k]); piv = i; }
        if (piv != k) {
            for (int j = 0; j < 4; ++j) { const float t = a[k][j]; a[k][j] = a[piv][j]; a[piv][j] = t; }
            det = -det;
        }
        det *= a[k][k];
        if (a[k][k] == 0.f) return 0.f;
        for (int i = k + 1; i < 4; ++i) {
            const float f = a[i][k] / a[k][k];
            for (int j = k + 1; j < 4; ++j) a[i][j] -= f * a[k][j];
        }
    }
    return det;
}

constexpr int kDecThreads = 128;
constexpr int kMaxD = 21 + kMaxCls;          // widest row (epistemic)

// Loads the 2*(5+C) (or 5+C) raw values of one anchor of one sample.  The block is 8-byte aligned for the aleatoric /
// epistemic layouts (an even number of floats per prior, maps 16 floats wide), so it is read as float2.
template <int NMAX>
__device__ __forceinline__ void load_block(const float* __restrict__ v, int n, bool vec2, float* dst) {
    if (vec2) {
#pragma unroll
        for (int i = 0; i < NMAX; i += 2)
            if (i < n) {
                const float2 f = __ldg(reinterpret_cast<const float2*>(v + i));
                dst[i] = f.x;
                dst[i + 1] = f.y;
            }
    } else {
#pragma unroll
        for (int i = 0; i < NMAX; ++i)
            if (i < n) dst[i] = __ldg(v + i);
    }
}

// CC = compile-time class count (2: the reference default, everything stays in registers) or 0 = run-time p.cls_cnt.
template <int CC>
__global__ void __launch_bounds__(kDecThreads) decode_kernel(const DecodeProblem p) {
    // rows of the block's 128 consecutive anchors are contiguous in the output: they are staged in shared memory
    // (row pitch D is odd for every variant at cls_cnt 2 -> conflict-free) and leave as coalesced stores
    extern __shared__ float srows[];
    const long long gid0 = (long long)blockIdx.x * kDecThreads;
    const long long gid = gid0 + threadIdx.x;
    const long long total = (long long)p.B * p.N;
    const bool active = gid < total;
    float* out = srows + threadIdx.x * p.D;
    if (active) {
    const int b = (int)(gid / p.N);
    int a = (int)(gid % p.N);
    int j = 0;
    for (; j < 2; ++j) {
        const int n = 3 * p.gh[j] * p.gw[j];
        if (a < n) break;
        a -= n;
    }
    const int lh = p.gh[j], lw = p.gw[j], pad = p.padded;
    const int prior = a / (lh * lw);
    const int cell = a - prior * lh * lw;
    const int row = cell / lw, col = cell - row * lw;
    const int C = CC ? CC : p.cls_cnt;
    constexpr int kC = CC ? CC : kMaxCls;      // array extents
    const int block = p.variant == 0 ? 5 + C : 2 * (5 + C);
    const bool vec2 = (block % 2 == 0) && (p.ld[j] % 2 == 0);
    const float pw = p.prior_w[j * 3 + prior], ph = p.prior_h[j * 3 + prior];
    float tx, ty, tw, th;      // (mean) t-space location
    const long long plane = (long long)(lh + 2 * pad) * (lw + 2 * pad);
    const float* v0 = p.raw[j] + ((long long)(row + pad) * (lw + 2 * pad) + col + pad) * p.ld[j] + prior * block;

    if (p.variant != 2) {
        float v[2 * (5 + kC)];
        load_block<2 * (5 + kC)>(v0 + (long long)b * plane * p.ld[j], block, vec2, v);
        tx = v[0]; ty = v[1]; tw = v[2]; th = v[3];
        float cls[kC];
        if (p.variant == 0) {
            out[4] = sigmoidf_(v[4]);
            softmax_(v + 5, C, cls);
            for (int i = 0; i < C; ++i) out[5 + i] = cls[i];
        } else {
            float prod = 1.f;
            for (int i = 0; i < 4; ++i) {
                const float var = expf(v[4 + i]);
                out[4 + i] = var;
                prod = (i == 0) ? var : prod * var;
            }
            out[8] = prod;
            const float obj = sigmoidf_(v[8]);
            out[9] = obj;
            out[10] = logistic_entropy(obj);
            softmax_(v + 10, C, cls);
            for (int i = 0; i < C; ++i) out[11 + i] = cls[i];
            out[11 + C] = softmax_entropy(cls, C);
            out[12 + C] = (float)j;
            out[13 + C] = (float)prior;
        }
    } else {
        const int T = p.T;
        float s_loc[4] = {0, 0, 0, 0}, s_var[4] = {0, 0, 0, 0}, s_out[4][4];
        float s_obj = 0.f, s_obj_h = 0.f, s_cls[kC], s_cls_h = 0.f;
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 4; ++k) s_out[i][k] = 0.f;
        for (int i = 0; i < C; ++i) s_cls[i] = 0.f;
        // the next sample's values are requested before the current one is processed (the loop is latency bound)
        float v[2 * (5 + kC)], vn[2 * (5 + kC)];
        const float* vt = v0 + (long long)b * T * plane * p.ld[j];
        load_block<2 * (5 + kC)>(vt, block, vec2, vn);
        for (int t = 0; t < T; ++t) {
#pragma unroll
            for (int i = 0; i < 2 * (5 + kC); ++i) v[i] = vn[i];
            if (t + 1 < T) load_block<2 * (5 + kC)>(vt + (long long)(t + 1) * plane * p.ld[j], block, vec2, vn);
            float loc[4];
            for (int i = 0; i < 4; ++i) { loc[i] = v[i]; s_loc[i] += loc[i]; s_var[i] += expf(v[4 + i]); }
            for (int i = 0; i < 4; ++i)
                for (int k = i; k < 4; ++k) s_out[i][k] += loc[i] * loc[k];
            const float obj = sigmoidf_(v[8]);
            s_obj += obj;
            s_obj_h += logistic_entropy(obj);
            float cls[kC];
            softmax_(v + 10, C, cls);
            for (int i = 0; i < C; ++i) s_cls[i] += cls[i];
            s_cls_h += softmax_entropy(cls, C);
        }
        const float fT = (float)T;
        float ev[4], cov[4][4];
        for (int i = 0; i < 4; ++i) ev[i] = s_loc[i] / fT;
        for (int i = 0; i < 4; ++i)
            for (int k = i; k < 4; ++k) {
                const float c = s_out[i][k] / fT - ev[i] * ev[k];
                cov[i][k] = c;
                cov[k][i] = c;
            }
        tx = ev[0]; ty = ev[1]; tw = ev[2]; th = ev[3];
        float ale_sum = 0.f;
        for (int i = 0; i < 4; ++i) {
            out[4 + i] = cov[i][i];
            const float av = s_var[i] / fT;
            out[8 + i] = av;
            ale_sum = (i == 0) ? av : ale_sum + av;
        }
        out[12] = det4(cov);
        out[13] = ale_sum;
        const float obj_mean = s_obj / fT;
        const float obj_h = logistic_entropy(obj_mean);
        out[14] = obj_mean;
        out[15] = obj_h - s_obj_h / fT;
        out[16] = obj_h;
        float cm[kC];
        for (int i = 0; i < C; ++i) { cm[i] = s_cls[i] / fT; out[17 + i] = cm[i]; }
        const float cls_h = softmax_entropy(cm, C);
        out[17 + C] = cls_h - s_cls_h / fT;
        out[18 + C] = cls_h;
        out[19 + C] = (float)j;
        out[20 + C] = (float)prior;
    }
    const float x = ((float)col + sigmoidf_(tx)) / (float)lw;
    const float y = ((float)row + sigmoidf_(ty)) / (float)lh;
    const float w2 = (expf(tw) * pw) / 2.f;
    const float h2 = (expf(th) * ph) / 2.f;
    out[0] = y - h2;
    out[1] = x - w2;
    out[2] = y + h2;
    out[3] = x + w2;
    }
    __syncthreads();
    const long long nvals = min((long long)kDecThreads, total - gid0) * p.D;
    float* dst = p.rows + gid0 * p.D;
    for (int i = threadIdx.x; i < nvals; i += kDecThreads) dst[i] = srows[i];
}

int launch_decode(const DecodeProblem& p, cudaStream_t st) {
    BY_REQUIRE(p.cls_cnt >= 1 && p.cls_cnt <= kMaxCls, "cls_cnt out of range");
    const long long total = (long long)p.B * p.N;
    BY_REQUIRE(p.D <= kMaxD, "row width out of range");
    const int grid = (int)((total + kDecThreads - 1) / kDecThreads);
    const size_t smem = sizeof(float) * kDecThreads * p.D;
    if (p.cls_cnt == 2) decode_kernel<2><<<grid, kDecThreads, smem, st>>>(p);
    else decode_kernel<0><<<grid, kDecThreads, smem, st>>>(p);
    BY_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace byolo
