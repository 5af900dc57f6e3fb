// libbyolo engine: weight folding/upload, per-batch execution plans (buffers, tensor maps, launch list) and the C ABI
// declared in include/byolo.h.  The plan is this repo's replacement for the graph that
// /root/reference/lib_yolo/yolov3.py:232-310 / 370-451 / 518-628 builds through model.ModelBuilder (model.py:20-185).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <vector>

#include "../../include/byolo.h"
#include "common.cuh"
#include "conv_umma.cuh"

namespace byolo {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

struct LayerDef { int k, s, cin, cout, bn, dropout; };

// layer table in creation order of the reference graph (darknet.py:7-39, yolov3.py:543-622)
static std::vector<LayerDef> layer_table(int variant, int cls_cnt) {
    std::vector<LayerDef> t;
    t.push_back({3, 1, 3, 32, 1, 0});
    int c = 32;
    const int stage[5][2] = {{32, 1}, {64, 2}, {128, 8}, {256, 8}, {512, 4}};
    for (auto& st : stage) {
        t.push_back({3, 2, c, 2 * st[0], 1, 0});
        c = 2 * st[0];
        for (int b = 0; b < st[1]; ++b) {
            t.push_back({1, 1, c, st[0], 1, 0});
            t.push_back({3, 1, st[0], c, 1, 0});
        }
    }
    const int det_ch = variant == BYOLO_STANDARD ? 3 * (5 + cls_cnt) : 3 * 2 * (5 + cls_cnt);
    const int mc = variant == BYOLO_EPISTEMIC;
    const int head[3][2] = {{512, 1024}, {256, 768}, {128, 384}};
    for (int j = 0; j < 3; ++j) {
        const int f = head[j][0];
        if (j) t.push_back({1, 1, 2 * f, f, 1, 0});
        t.push_back({1, 1, head[j][1], f, 1, mc});
        t.push_back({3, 1, f, 2 * f, 1, mc});
        t.push_back({1, 1, 2 * f, f, 1, mc});
        t.push_back({3, 1, f, 2 * f, 1, mc});
        t.push_back({1, 1, 2 * f, f, 1, mc});
        t.push_back({3, 1, f, 2 * f, 1, 0});
        t.push_back({1, 1, 2 * f, det_ch, 0, 0});
    }
    return t;
}

struct LayerWeights {
    int K = 0, cout_pad = 0;
    float* bias = nullptr;     // [cout_pad] BN shift or detection bias
    __half* w16 = nullptr;     // [cout_pad, K]
    float* w32 = nullptr;      // [K, cout_pad]
    void release() {
        cudaFree(bias); cudaFree(w16); cudaFree(w32);
        bias = nullptr; w16 = nullptr; w32 = nullptr;
    }
};

// Folds BN into (weights, shift) and uploads both layouts.  kernel: HWIO fp32; bn = beta,gamma,mean,var | nullptr.
// split: the fp16 matrix carries 2*cout_pad rows - fp16(v) in rows [0, cout_pad), fp16(v - hi) in the rows behind them
// (split-fp16 mode, BYOLO_PREC_FP16X3).
static int fold_and_upload(const LayerDef& d, const float* kernel, const float* bn, const float* bias, LayerWeights* out,
                           bool split = false) {
    const int K = d.k * d.k * d.cin;
    const int cp = (d.cout + 15) / 16 * 16;
    std::vector<float> scale(d.cout, 1.f), shift(cp, 0.f);
    for (int c = 0; c < d.cout; ++c) {
        if (bn) {
            const float beta = bn[c], gamma = bn[d.cout + c], mean = bn[2 * d.cout + c], var = bn[3 * d.cout + c];
            scale[c] = gamma / std::sqrt(var + 1e-5f);              // layers.py:511,516: epsilon=1e-05
            shift[c] = beta - mean * scale[c];
        } else {
            shift[c] = bias ? bias[c] : 0.f;
        }
    }
    std::vector<float> w32((size_t)K * cp, 0.f);
    std::vector<__half> w16((size_t)cp * K * (split ? 2 : 1), __float2half(0.f));
    for (int kk = 0; kk < K; ++kk)
        for (int c = 0; c < d.cout; ++c) {
            const float v = kernel[(size_t)kk * d.cout + c] * scale[c];   // HWIO flattened == [K][cout]
            w32[(size_t)kk * cp + c] = v;
            const __half hi = __float2half_rn(v);
            w16[(size_t)c * K + kk] = hi;
            if (split) w16[((size_t)cp + c) * K + kk] = __float2half_rn(v - __half2float(hi));
        }
    out->K = K;
    out->cout_pad = cp;
    auto upload = [&]() -> int {
        BY_CUDA(cudaMalloc(&out->bias, sizeof(float) * cp));
        BY_CUDA(cudaMalloc(&out->w32, sizeof(float) * w32.size()));
        BY_CUDA(cudaMalloc(&out->w16, sizeof(__half) * w16.size()));
        BY_CUDA(cudaMemcpy(out->bias, shift.data(), sizeof(float) * cp, cudaMemcpyHostToDevice));
        BY_CUDA(cudaMemcpy(out->w32, w32.data(), sizeof(float) * w32.size(), cudaMemcpyHostToDevice));
        BY_CUDA(cudaMemcpy(out->w16, w16.data(), sizeof(__half) * w16.size(), cudaMemcpyHostToDevice));
        return 0;
    };
    const int rc = upload();
    if (rc) out->release();          // no partial allocations survive a failure
    return rc;
}

struct Buffer {
    void* ptr = nullptr;
    Geom g{0, 0, 0, 0};            // dense NHWC geometry; C = stored channels
    bool f32 = false;              // raw detection maps are fp32 (C = cout padded to 16), everything else is T
    int valid_c = 0;               // channels that carry data (== C except for the raw detection maps)
    size_t bytes = 0;
};

enum StepKind { STEP_STEM, STEP_CONV };

struct Step {
    StepKind kind;
    int layer = -1;
    ConvProblem prob{};
    UmmaLaunch ul{};
    int out_buf = -1;
};

struct Plan {
    int B = 0;
    std::vector<Buffer> bufs;
    std::vector<Step> steps;
    std::vector<int> conv_out;      // conv index -> buffer id
    int raw_buf[3] = {-1, -1, -1};
    float* rows_scratch = nullptr;  // [B,N,D] when the caller passes no rows buffer
    float* img_stage = nullptr;     // device staging for byolo_detect_host
    float* out_stage = nullptr;
    int* cnt_stage = nullptr;
    int stage_max_out = 0;
    unsigned long long* clk = nullptr;   // profiling: 4 x u64 per step (conv kernels write SM clock / global timer pairs)
    std::vector<cudaEvent_t> ev;    // profiling: steps.size() + 3 events (one before each launch, decode, nms, end)
    bool ev_recorded = false;
    // coarse profiling (byolo_profile(h, 2)): 4 events per byolo_detect in a ring - start, after the stem launch, before the
    // decode launch, after the NMS launch.  Nothing sits between the conv launches, so they overlap exactly as without it.
    static constexpr int kCoarseRing = 256;
    std::vector<cudaEvent_t> cev;
    int coarse_n = 0;
    ~Plan() {
        for (auto& x : cev) cudaEventDestroy(x);
        for (auto& x : ev) cudaEventDestroy(x);
        cudaFree(clk);
        for (auto& b : bufs) cudaFree(b.ptr);
        cudaFree(rows_scratch); cudaFree(img_stage); cudaFree(out_stage); cudaFree(cnt_stage);
    }
};

}  // namespace byolo

using namespace byolo;

struct byolo_engine {
    byolo_config cfg{};
    std::vector<LayerDef> layers;
    std::vector<LayerWeights> weights;
    bool loaded = false;
    std::map<int, std::unique_ptr<Plan>> plans;
    Plan* last_plan = nullptr;     // plan of the most recent forward (byolo_get_activation)
    bool profiling = false;
    bool coarse = false;           // byolo_profile(h, 2)
    // pipelined host entry (byolo_submit_host / byolo_wait_host): two slots, separate H2D and D2H copy streams
    struct HostSlot {
        float* img = nullptr; float* out = nullptr; int* cnt = nullptr;
        size_t img_bytes = 0, out_bytes = 0;
        cudaEvent_t h2d = nullptr, done = nullptr, d2h = nullptr;
        bool used = false;
    } slot[2];
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    int N = 0, D = 0, obj_idx = 0, cls_start = 0;
    int gh[3], gw[3];
    bool umma() const { return cfg.precision == BYOLO_PREC_FP16 || cfg.precision == BYOLO_PREC_FP16X3; }
    bool x3() const { return cfg.precision == BYOLO_PREC_FP16X3; }
    int act_fmt() const { return cfg.precision == BYOLO_PREC_FP32 ? ACT_F32 : (x3() ? ACT_F16_HILO : ACT_F16); }
    size_t esize() const { return act_fmt() == ACT_F16 ? 2 : 4; }      // bytes per activation element (hi + lo pair: 4)
    bool mc() const { return cfg.variant == BYOLO_EPISTEMIC; }
    int samples(int B) const { return mc() ? B * cfg.T : B; }
    ~byolo_engine() {
        plans.clear();
        for (auto& w : weights) w.release();
        for (auto& s : slot) {
            cudaFree(s.img); cudaFree(s.out); cudaFree(s.cnt);
            if (s.h2d) cudaEventDestroy(s.h2d);
            if (s.done) cudaEventDestroy(s.done);
            if (s.d2h) cudaEventDestroy(s.d2h);
        }
        if (s_h2d) cudaStreamDestroy(s_h2d);
        if (s_d2h) cudaStreamDestroy(s_d2h);
    }
};

namespace byolo {

static int new_buffer(byolo_engine* e, Plan* pl, int S, int H, int W, int C, bool f32, int* id) {
    Buffer b;
    b.g = Geom{S, H, W, C};
    b.f32 = f32;
    b.valid_c = C;
    b.bytes = (size_t)b.g.rows() * C * (f32 ? 4 : e->esize());
    BY_CUDA(cudaMalloc(&b.ptr, b.bytes));
    BY_CUDA(cudaMemset(b.ptr, 0, b.bytes));
    pl->bufs.push_back(b);
    *id = (int)pl->bufs.size() - 1;
    return 0;
}

// Appends the conv step of layer `li`.  in2 < 0: single input.  residual < 0: none.  t1 / t2 > 1: that input is a cached
// backbone map of B images which the conv reads as B*T stacked samples (stack_feature_map, layers.py:595-597).
static int add_conv(byolo_engine* e, Plan* pl, int li, int in1, int in2, int residual, int out_mode, int drop_id, int* out_id,
                    int t1 = 1, int t2 = 1) {
    const LayerDef& d = e->layers[li];
    const LayerWeights& w = e->weights[li];
    Buffer bin = pl->bufs[in1];
    // T-invariant conv (conv "75", yolov3.py:538-544: a 1x1 dropout conv whose ONLY input is a stacked backbone map): on the
    // tensor-core paths the GEMM runs once per image and the epilogue writes the T masked samples (EPI_F16_DROP_T)
    int t_out = 1;
    if (e->umma() && t1 > 1 && in2 < 0 && d.k == 1 && drop_id >= 0 && out_mode == OUT_DENSE) { t_out = t1; t1 = 1; }
    bin.g.S *= t1;                                  // geometry as the conv sees it
    BY_REQUIRE(bin.g.C + (in2 >= 0 ? pl->bufs[in2].g.C : 0) == d.cin, "plan wiring: channel mismatch");
    BY_REQUIRE(in2 < 0 || pl->bufs[in2].g.S * t2 == bin.g.S, "plan wiring: sample count mismatch");
    const int Ho = bin.g.H / d.s, Wo = bin.g.W / d.s;
    int ob;
    if (out_mode == OUT_DENSE_F32) {
        if (int r = new_buffer(e, pl, bin.g.S, Ho, Wo, w.cout_pad, true, &ob)) return r;
        pl->bufs[ob].valid_c = d.cout;
    } else if (out_mode == OUT_UPSAMPLE2) {
        if (int r = new_buffer(e, pl, bin.g.S, 2 * Ho, 2 * Wo, d.cout, false, &ob)) return r;
    } else {
        if (int r = new_buffer(e, pl, bin.g.S * t_out, Ho, Wo, d.cout, false, &ob)) return r;
    }
    Step st;
    st.kind = STEP_CONV;
    st.layer = li;
    st.out_buf = ob;
    ConvProblem& p = st.prob;
    p.in1 = bin.ptr;
    p.in2 = in2 >= 0 ? pl->bufs[in2].ptr : nullptr;
    p.gin = bin.g;
    p.c2 = in2 >= 0 ? pl->bufs[in2].g.C : 0;
    p.t1 = t1;
    p.t2 = t2;
    p.t_out = t_out;
    p.k = d.k;
    p.stride = d.s;
    p.cout_pad = w.cout_pad;
    p.w16 = w.w16;
    p.w32 = w.w32;
    p.x3 = e->x3();
    p.ep.bias = w.bias;
    p.ep.residual = residual >= 0 ? pl->bufs[residual].ptr : nullptr;
    p.ep.out = pl->bufs[ob].ptr;
    p.ep.out_mode = out_mode;
    p.ep.ldc = out_mode == OUT_DENSE_F32 ? w.cout_pad : d.cout;
    p.ep.cout = d.cout;
    p.ep.leaky = d.bn;
    p.ep.drop.enabled = drop_id >= 0;
    p.ep.drop.layer_id = drop_id;
    p.ep.drop.T = e->cfg.T;
    p.ep.drop.thr16 = (uint32_t)std::lround((double)e->cfg.drop_prob * 65536.0);
    p.ep.drop.keep_scale = 1.0f / (1.0f - e->cfg.drop_prob);
    if (e->umma())
        if (int r = umma_prepare(p, &st.ul)) return r;
    pl->steps.push_back(st);
    pl->conv_out[li] = ob;
    *out_id = ob;
    return 0;
}

static int build_plan(byolo_engine* e, int B, Plan* pl) {
    const byolo_config& c = e->cfg;
    pl->B = B;
    pl->conv_out.assign(e->layers.size(), -1);
    const int T = e->mc() ? c.T : 1;
    int li = 0, cur, tmp;
    // ---- darknet53 (darknet.py:7-39) ----
    if (int r = new_buffer(e, pl, B, c.height, c.width, 32, false, &cur)) return r;
    {
        Step st;
        st.kind = STEP_STEM;
        st.layer = 0;
        st.out_buf = cur;
        pl->steps.push_back(st);
        pl->conv_out[0] = cur;
        li = 1;
    }
    const int blocks[5] = {1, 2, 8, 8, 4};
    int taps[5];
    for (int sidx = 0; sidx < 5; ++sidx) {
        if (int r = add_conv(e, pl, li++, cur, -1, -1, OUT_DENSE, -1, &cur)) return r;            // downsample
        for (int b = 0; b < blocks[sidx]; ++b) {                                                   // residual block
            if (int r = add_conv(e, pl, li++, cur, -1, -1, OUT_DENSE, -1, &tmp)) return r;
            if (int r = add_conv(e, pl, li++, tmp, -1, cur, OUT_DENSE, -1, &cur)) return r;       // + shortcut (model.py:96-99)
        }
        taps[sidx] = cur;
    }
    const int l36 = taps[2], l61 = taps[3], l74 = taps[4];
    // ---- det_net_1..3 (yolov3.py:543-622) ----
    int drop_next = 0;
    auto drop_id = [&](int layer) {
        return (e->layers[layer].dropout && e->mc() && !c.standard_test_dropout) ? drop_next++ : -1;
    };
    int route_src = -1;
    const int srcs[3] = {l74, l61, l36};
    for (int j = 0; j < 3; ++j) {
        // stack_feature_map(-1 | 61 | 36, T) is not materialised: the first conv of each head reads the cached backbone
        // map of B images as B*T samples (t1 / t2 = T)
        int x1 = srcs[0], x2 = -1;
        if (j > 0) {
            if (int r = add_conv(e, pl, li++, route_src, -1, -1, OUT_UPSAMPLE2, -1, &x1)) return r;  // conv 84/96 + upsample
            x2 = srcs[j];
        }
        int x = x1, c5 = -1;
        for (int i = 0; i < 6; ++i) {
            const int l = li++;
            if (i == 0) {
                if (int r = add_conv(e, pl, l, x, x2, -1, OUT_DENSE, drop_id(l), &x, j == 0 ? T : 1, j == 0 ? 1 : T)) return r;
            } else if (int r = add_conv(e, pl, l, x, -1, -1, OUT_DENSE, drop_id(l), &x)) {
                return r;
            }
            if (i == 4) c5 = x;
        }
        if (int r = add_conv(e, pl, li++, x, -1, -1, OUT_DENSE_F32, -1, &pl->raw_buf[j])) return r;  // detection conv
        route_src = c5;                                                                            // route([-3])
    }
    BY_REQUIRE(li == (int)e->layers.size(), "plan did not consume all layers");
    BY_CUDA(cudaMalloc(&pl->rows_scratch, sizeof(float) * (size_t)B * e->N * e->D));
    return 0;
}

static int get_plan(byolo_engine* e, int B, Plan** out) {
    BY_REQUIRE(e->loaded, "byolo_load_weights has not been called");
    BY_REQUIRE(B >= 1 && B <= e->cfg.max_batch, "batch size out of range");
    auto it = e->plans.find(B);
    if (it == e->plans.end()) {
        std::unique_ptr<Plan> pl(new Plan());
        if (int r = build_plan(e, B, pl.get())) return r;
        it = e->plans.emplace(B, std::move(pl)).first;
    }
    *out = it->second.get();
    e->last_plan = *out;
    return 0;
}

static int run_forward(byolo_engine* e, Plan* pl, const float* img, uint64_t seed, int image0, float* rows, cudaStream_t st) {
    const byolo_config& c = e->cfg;
    const bool prof = e->profiling;
    if (prof && pl->ev.empty()) {
        pl->ev.resize(pl->steps.size() + 3);
        for (auto& x : pl->ev) BY_CUDA(cudaEventCreate(&x));
        BY_CUDA(cudaMalloc(&pl->clk, sizeof(unsigned long long) * 4 * pl->steps.size()));
        BY_CUDA(cudaMemset(pl->clk, 0, sizeof(unsigned long long) * 4 * pl->steps.size()));
    }
    const bool coarse = e->coarse && !prof;
    cudaEvent_t* ce = nullptr;
    if (coarse) {
        if (pl->cev.empty()) {
            pl->cev.resize(4 * Plan::kCoarseRing);
            for (auto& x : pl->cev) BY_CUDA(cudaEventCreate(&x));
        }
        ce = &pl->cev[4 * (pl->coarse_n % Plan::kCoarseRing)];
        BY_CUDA(cudaEventRecord(ce[0], st));
    }
    size_t ei = 0;
    for (Step& s : pl->steps) {
        if (coarse && &s == &pl->steps[1]) BY_CUDA(cudaEventRecord(ce[1], st));       // the stem is the first launch
        if (prof) BY_CUDA(cudaEventRecord(pl->ev[ei++], st));
        if (s.kind == STEP_STEM) {
            const LayerWeights& w = e->weights[0];
            if (e->umma()) {
                if (int r = launch_stem_mma(img, pl->B, c.height, c.width, w.w16, w.bias, pl->bufs[s.out_buf].ptr, e->x3(), st)) return r;
            } else if (int r = launch_stem(img, pl->B, c.height, c.width, w.w32, w.bias, pl->bufs[s.out_buf].ptr, e->act_fmt() == ACT_F16, st)) {
                return r;
            }
        } else {
            if (e->umma()) {
                Dropout& d = s.ul.p.ep.drop;
                d.seed_lo = (uint32_t)seed; d.seed_hi = (uint32_t)(seed >> 32); d.image0 = image0;
                s.ul.p.clk = prof ? pl->clk + 4 * (ei - 1) : nullptr;
                if (int r = umma_launch(s.ul, st)) return r;
            } else {
                Dropout& d = s.prob.ep.drop;
                d.seed_lo = (uint32_t)seed; d.seed_hi = (uint32_t)(seed >> 32); d.image0 = image0;
                if (int r = launch_conv_simt(s.prob, e->act_fmt() == ACT_F16, st)) return r;
            }
        }
    }
    DecodeProblem dp{};
    dp.variant = c.variant;
    dp.B = pl->B;
    dp.T = e->mc() ? c.T : 1;
    dp.cls_cnt = c.cls_cnt;
    for (int j = 0; j < 3; ++j) {
        dp.gh[j] = e->gh[j]; dp.gw[j] = e->gw[j];
        dp.raw[j] = (const float*)pl->bufs[pl->raw_buf[j]].ptr;
        dp.ld[j] = pl->bufs[pl->raw_buf[j]].g.C;
    }
    dp.padded = 0;
    for (int i = 0; i < 9; ++i) { dp.prior_h[i] = c.prior_h[i]; dp.prior_w[i] = c.prior_w[i]; }
    dp.rows = rows;
    dp.N = e->N;
    dp.D = e->D;
    if (prof) BY_CUDA(cudaEventRecord(pl->ev[ei++], st));
    if (coarse) BY_CUDA(cudaEventRecord(ce[2], st));
    if (int r = launch_decode(dp, st)) return r;
    if (prof) { BY_CUDA(cudaEventRecord(pl->ev[ei++], st)); pl->ev_recorded = false; }
    return 0;
}

}  // namespace byolo

// =====================================================================================================================
//                                                       C ABI
// =====================================================================================================================
extern "C" {

int byolo_version(void) { return 2; }
const char* byolo_last_error(void) { return g_err.c_str(); }

int byolo_create(const byolo_config* cfg, byolo_handle* out) {
    BY_REQUIRE(cfg && out, "null argument");
    BY_REQUIRE(cfg->variant >= 0 && cfg->variant <= 2, "unknown variant");
    BY_REQUIRE(cfg->height > 0 && cfg->width > 0 && cfg->height % 32 == 0 && cfg->width % 32 == 0,
               "image size must be a multiple of 32 (yolov3.py:207-211)");
    BY_REQUIRE(cfg->cls_cnt >= 1 && cfg->cls_cnt <= 16, "cls_cnt must be in [1,16]");
    BY_REQUIRE(cfg->max_batch >= 1, "max_batch must be >= 1");
    BY_REQUIRE(cfg->precision >= 0 && cfg->precision <= 3, "unknown precision");
    BY_REQUIRE(cfg->variant != BYOLO_EPISTEMIC || cfg->T >= 1, "epistemic model needs T >= 1 (yolov3.py:467-468)");
    BY_REQUIRE(cfg->drop_prob >= 0.f && cfg->drop_prob < 1.f, "drop_prob must be in [0,1)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("no CUDA device: libbyolo has no CPU fallback");
        return -2;
    }
    byolo_engine* e = new byolo_engine();
    e->cfg = *cfg;
    if (cfg->variant != BYOLO_EPISTEMIC) e->cfg.T = 1;
    e->layers = layer_table(cfg->variant, cfg->cls_cnt);
    e->N = 0;
    for (int j = 0; j < 3; ++j) {
        e->gh[j] = cfg->height / (32 >> j);
        e->gw[j] = cfg->width / (32 >> j);
        e->N += 3 * e->gh[j] * e->gw[j];
    }
    const int C = cfg->cls_cnt;
    if (cfg->variant == BYOLO_STANDARD) { e->D = 5 + C; e->obj_idx = 4; e->cls_start = 5; }
    if (cfg->variant == BYOLO_ALEATORIC) { e->D = 14 + C; e->obj_idx = 9; e->cls_start = 11; }
    if (cfg->variant == BYOLO_EPISTEMIC) { e->D = 21 + C; e->obj_idx = 14; e->cls_start = 17; }
    *out = e;
    return 0;
}

int byolo_destroy(byolo_handle h) {
    if (h) { cudaDeviceSynchronize(); delete h; }
    return 0;
}

int byolo_load_weights(byolo_handle h, const void* blob, size_t bytes) {
    BY_REQUIRE(h && blob, "null argument");
    const int32_t* hdr = (const int32_t*)blob;
    BY_REQUIRE(bytes >= 8 && hdr[0] == 0x31575942, "not a BYW1 weight blob");
    const int n = hdr[1];
    BY_REQUIRE(n == (int)h->layers.size(), "weight blob has the wrong number of layers");
    BY_REQUIRE(bytes >= 8 + (size_t)24 * n, "weight blob truncated");
    const int32_t* tab = hdr + 2;
    // validate the whole blob before touching the engine: a bad blob leaves the previous weights in place
    size_t off = 8 + (size_t)24 * n;
    for (int i = 0; i < n; ++i) {
        const LayerDef& d = h->layers[i];
        BY_REQUIRE(tab[6 * i] == d.k && tab[6 * i + 1] == d.s && tab[6 * i + 2] == d.cin && tab[6 * i + 3] == d.cout &&
                       tab[6 * i + 4] == d.bn,
                   "weight blob layer table does not match the model variant");
        const size_t nv = (size_t)(d.bn ? 4 : 1) * d.cout, nk = (size_t)d.k * d.k * d.cin * d.cout;
        BY_REQUIRE(off + 4 * (nv + nk) <= bytes, "weight blob truncated");
        off += 4 * (nv + nk);
    }
    BY_REQUIRE(off == bytes, "weight blob has trailing bytes");     // darknet.py:66 `assert ptr == len(weights)`
    // upload into a fresh table and swap on success; on a CUDA failure the partial table is released and the engine is
    // left without weights (loaded = false) rather than with dangling pointers
    h->loaded = false;
    h->plans.clear();
    h->last_plan = nullptr;
    for (auto& w : h->weights) w.release();
    h->weights.clear();
    std::vector<LayerWeights> fresh(n);
    off = 8 + (size_t)24 * n;
    int rc = 0;
    for (int i = 0; i < n && !rc; ++i) {
        const LayerDef& d = h->layers[i];
        const size_t nv = (size_t)(d.bn ? 4 : 1) * d.cout, nk = (size_t)d.k * d.k * d.cin * d.cout;
        const float* vec = (const float*)((const char*)blob + off);
        rc = fold_and_upload(d, vec + nv, d.bn ? vec : nullptr, d.bn ? nullptr : vec, &fresh[i], h->x3());
        off += 4 * (nv + nk);
    }
    if (!rc && cudaDeviceSynchronize() != cudaSuccess) { set_error("byolo_load_weights: device synchronisation failed"); rc = -2; }
    if (rc) {
        for (auto& w : fresh) w.release();
        return rc;
    }
    h->weights.swap(fresh);
    h->loaded = true;
    return 0;
}

int byolo_output_shape(byolo_handle h, int32_t* N, int32_t* D, int32_t* obj_idx, int32_t* cls_start_idx) {
    BY_REQUIRE(h, "null handle");
    if (N) *N = h->N;
    if (D) *D = h->D;
    if (obj_idx) *obj_idx = h->obj_idx;
    if (cls_start_idx) *cls_start_idx = h->cls_start;
    return 0;
}

int byolo_forward(byolo_handle h, const float* img_dev, int32_t B, uint64_t seed, int32_t image_index0, float* rows_dev,
                  void* stream) {
    BY_REQUIRE(h && img_dev && rows_dev, "null argument");
    Plan* pl;
    if (int r = get_plan(h, B, &pl)) return r;
    return run_forward(h, pl, img_dev, seed, image_index0, rows_dev, (cudaStream_t)stream);
}

int byolo_nms(const float* rows_dev, int32_t B, int32_t N, int32_t D, int32_t obj_idx, float iou_thr, int32_t max_out,
              float* out_rows_dev, int32_t* out_idx_dev, int32_t* out_count_dev, void* stream) {
    BY_REQUIRE(rows_dev && out_rows_dev && out_count_dev, "null argument");
    return launch_nms(rows_dev, B, N, D, obj_idx, iou_thr, max_out, out_rows_dev, out_idx_dev, out_count_dev, NmsOptions{},
                      (cudaStream_t)stream);
}

int byolo_nms_ex(const float* rows_dev, int32_t B, int32_t N, int32_t D, int32_t obj_idx, float iou_thr, int32_t max_out,
                 float* out_rows_dev, int32_t* out_idx_dev, int32_t* out_count_dev, int32_t packed, int32_t force_cluster_size,
                 int32_t force_chunked, void* stream) {
    BY_REQUIRE(rows_dev && out_rows_dev, "null argument");
    BY_REQUIRE(packed || out_count_dev, "pass out_count_dev or ask for the packed layout");
    NmsOptions opt;
    opt.packed = packed != 0;
    opt.force_cs = force_cluster_size;
    opt.force_chunked = force_chunked != 0;
    return launch_nms(rows_dev, B, N, D, obj_idx, iou_thr, max_out, out_rows_dev, out_idx_dev, out_count_dev, opt, (cudaStream_t)stream);
}

int byolo_class_filter(const float* rows_dev, int32_t B, int32_t N, int32_t D, int32_t obj_idx, int32_t cls_start_idx, int32_t cls_cnt,
                       int32_t cls, float* out_rows_dev, void* stream) {
    BY_REQUIRE(rows_dev && out_rows_dev && B >= 0 && N >= 0, "bad argument");
    return launch_class_filter(rows_dev, (long long)B * N, D, obj_idx, cls_start_idx, cls_cnt, cls, out_rows_dev, (cudaStream_t)stream);
}

static int detect_impl(byolo_handle h, const float* img_dev, int32_t B, uint64_t seed, int32_t image_index0, float iou_thr,
                       int32_t max_out, float* rows_dev, float* out_rows_dev, int32_t* out_idx_dev, int32_t* out_count_dev,
                       bool packed, void* stream) {
    Plan* pl;
    if (int r = get_plan(h, B, &pl)) return r;
    float* rows = rows_dev ? rows_dev : pl->rows_scratch;
    if (int r = run_forward(h, pl, img_dev, seed, image_index0, rows, (cudaStream_t)stream)) return r;
    NmsOptions opt;
    opt.packed = packed;
    if (int r = launch_nms(rows, B, h->N, h->D, h->obj_idx, iou_thr, max_out, out_rows_dev, out_idx_dev, out_count_dev, opt,
                           (cudaStream_t)stream))
        return r;
    if (h->profiling) { BY_CUDA(cudaEventRecord(pl->ev.back(), (cudaStream_t)stream)); pl->ev_recorded = true; }
    if (h->coarse && !h->profiling && !pl->cev.empty()) {
        BY_CUDA(cudaEventRecord(pl->cev[4 * (pl->coarse_n % Plan::kCoarseRing) + 3], (cudaStream_t)stream));
        ++pl->coarse_n;
    }
    return 0;
}

int byolo_detect(byolo_handle h, const float* img_dev, int32_t B, uint64_t seed, int32_t image_index0, float iou_thr,
                 int32_t max_out, float* rows_dev, float* out_rows_dev, int32_t* out_idx_dev, int32_t* out_count_dev,
                 void* stream) {
    BY_REQUIRE(h && img_dev && out_rows_dev && out_count_dev, "null argument");
    return detect_impl(h, img_dev, B, seed, image_index0, iou_thr, max_out, rows_dev, out_rows_dev, out_idx_dev, out_count_dev, false, stream);
}

int byolo_detect_packed(byolo_handle h, const float* img_dev, int32_t B, uint64_t seed, int32_t image_index0, float iou_thr,
                        int32_t max_out, float* out_packed_dev, int32_t* out_idx_dev, void* stream) {
    BY_REQUIRE(h && img_dev && out_packed_dev, "null argument");
    return detect_impl(h, img_dev, B, seed, image_index0, iou_thr, max_out, nullptr, out_packed_dev, out_idx_dev, nullptr, true, stream);
}

int byolo_profile(byolo_handle h, int32_t enable) {
    BY_REQUIRE(h, "null handle");
    h->profiling = enable == 1;
    h->coarse = enable == 2;
    if (h->coarse)
        for (auto& kv : h->plans) kv.second->coarse_n = 0;
    return 0;
}

int byolo_profile_read_coarse(byolo_handle h, float* stem_ms, float* conv_ms, float* tail_ms, int32_t capacity) {
    BY_REQUIRE(h && stem_ms && conv_ms && tail_ms, "null argument");
    Plan* pl = h->last_plan;
    BY_REQUIRE(pl && !pl->cev.empty(), "no byolo_detect has run in coarse profiling mode");
    const int n = std::min(std::min(pl->coarse_n, (int)Plan::kCoarseRing), (int)capacity);
    for (int i = 0; i < n; ++i) {
        const int slot = (pl->coarse_n - n + i) % Plan::kCoarseRing;
        cudaEvent_t* ce = &pl->cev[4 * slot];
        BY_CUDA(cudaEventSynchronize(ce[3]));
        BY_CUDA(cudaEventElapsedTime(&stem_ms[i], ce[0], ce[1]));
        BY_CUDA(cudaEventElapsedTime(&conv_ms[i], ce[1], ce[2]));
        BY_CUDA(cudaEventElapsedTime(&tail_ms[i], ce[2], ce[3]));
    }
    return n;
}

int byolo_profile_read(byolo_handle h, float* ms, int32_t* kind, int32_t* layer, double* flops, float* sm_mhz, int32_t capacity) {
    BY_REQUIRE(h && ms && kind && layer && flops && sm_mhz, "null argument");
    Plan* pl = h->last_plan;
    BY_REQUIRE(pl && pl->ev_recorded, "no profiled byolo_detect has completed");
    const int n = (int)pl->steps.size() + 2;
    BY_REQUIRE(capacity >= n, "capacity too small");
    BY_CUDA(cudaEventSynchronize(pl->ev.back()));
    std::vector<unsigned long long> clk(4 * pl->steps.size(), 0ull);
    BY_CUDA(cudaMemcpy(clk.data(), pl->clk, sizeof(unsigned long long) * clk.size(), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) {
        BY_CUDA(cudaEventElapsedTime(&ms[i], pl->ev[i], pl->ev[i + 1]));
        flops[i] = 0.0;
        sm_mhz[i] = 0.f;
        if (i < (int)pl->steps.size() && clk[4 * i + 3] > clk[4 * i + 1])
            sm_mhz[i] = (float)((double)(clk[4 * i + 2] - clk[4 * i]) * 1e3 / (double)(clk[4 * i + 3] - clk[4 * i + 1]));
        if (i < (int)pl->steps.size()) {
            const Step& s = pl->steps[i];
            kind[i] = (int)s.kind;
            layer[i] = s.layer;
            const LayerDef& d = h->layers[s.layer];
            const Buffer& ob = pl->bufs[s.out_buf];
            const int up = (s.kind == STEP_CONV && s.prob.ep.out_mode == OUT_UPSAMPLE2) ? 4 : 1;
            flops[i] = 2.0 * d.k * d.k * d.cin * d.cout * (double)ob.g.S * ob.g.H * ob.g.W / up;
        } else {
            kind[i] = 3 + (i - (int)pl->steps.size());      // 3 decode, 4 nms
            layer[i] = -1;
        }
    }
    return n;
}

int byolo_detect_host(byolo_handle h, const float* img_host, int32_t B, uint64_t seed, int32_t image_index0, float iou_thr,
                      int32_t max_out, float* out_rows_host, int32_t* out_count_host, void* stream) {
    BY_REQUIRE(h && img_host && out_rows_host && out_count_host, "null argument");
    Plan* pl;
    if (int r = get_plan(h, B, &pl)) return r;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t img_bytes = sizeof(float) * (size_t)B * h->cfg.height * h->cfg.width * 3;
    if (!pl->img_stage) BY_CUDA(cudaMalloc(&pl->img_stage, img_bytes));
    if (pl->stage_max_out < max_out) {
        cudaFree(pl->out_stage); cudaFree(pl->cnt_stage);
        pl->out_stage = nullptr; pl->cnt_stage = nullptr;
        BY_CUDA(cudaMalloc(&pl->out_stage, sizeof(float) * (size_t)B * max_out * h->D));
        BY_CUDA(cudaMalloc(&pl->cnt_stage, sizeof(int) * B));
        pl->stage_max_out = max_out;
    }
    BY_CUDA(cudaMemcpyAsync(pl->img_stage, img_host, img_bytes, cudaMemcpyHostToDevice, st));
    if (int r = byolo_detect(h, pl->img_stage, B, seed, image_index0, iou_thr, max_out, nullptr, pl->out_stage, nullptr,
                             pl->cnt_stage, stream))
        return r;
    BY_CUDA(cudaMemcpyAsync(out_rows_host, pl->out_stage, sizeof(float) * (size_t)B * max_out * h->D, cudaMemcpyDeviceToHost, st));
    BY_CUDA(cudaMemcpyAsync(out_count_host, pl->cnt_stage, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
    BY_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int byolo_submit_host(byolo_handle h, const float* img_host, int32_t B, uint64_t seed, int32_t image_index0, float iou_thr,
                      int32_t max_out, float* out_rows_host, int32_t* out_count_host, int32_t slot, void* stream) {
    BY_REQUIRE(h && img_host && out_rows_host && out_count_host, "null argument");
    BY_REQUIRE(slot == 0 || slot == 1, "slot must be 0 or 1");
    cudaStream_t st = (cudaStream_t)stream;
    if (!h->s_h2d) {
        BY_CUDA(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
        BY_CUDA(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
    }
    byolo_engine::HostSlot& s = h->slot[slot];
    if (!s.h2d) {
        BY_CUDA(cudaEventCreateWithFlags(&s.h2d, cudaEventDisableTiming));
        BY_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        BY_CUDA(cudaEventCreateWithFlags(&s.d2h, cudaEventDisableTiming));
    }
    const size_t img_bytes = sizeof(float) * (size_t)B * h->cfg.height * h->cfg.width * 3;
    const size_t out_bytes = sizeof(float) * (size_t)B * max_out * h->D;
    if (s.img_bytes < img_bytes) {
        BY_CUDA(cudaDeviceSynchronize());
        cudaFree(s.img); s.img = nullptr;
        BY_CUDA(cudaMalloc(&s.img, img_bytes));
        s.img_bytes = img_bytes;
    }
    if (s.out_bytes < out_bytes) {
        BY_CUDA(cudaDeviceSynchronize());
        cudaFree(s.out); cudaFree(s.cnt); s.out = nullptr; s.cnt = nullptr;
        BY_CUDA(cudaMalloc(&s.out, out_bytes));
        BY_CUDA(cudaMalloc(&s.cnt, sizeof(int) * h->cfg.max_batch));
        s.out_bytes = out_bytes;
    }
    // H2D of this step may run while the previous step computes; it only has to wait for the last compute that READ
    // this slot's staging buffer
    if (s.used) BY_CUDA(cudaStreamWaitEvent(h->s_h2d, s.done, 0));
    BY_CUDA(cudaMemcpyAsync(s.img, img_host, img_bytes, cudaMemcpyHostToDevice, h->s_h2d));
    BY_CUDA(cudaEventRecord(s.h2d, h->s_h2d));
    BY_CUDA(cudaStreamWaitEvent(st, s.h2d, 0));
    if (s.used) BY_CUDA(cudaStreamWaitEvent(st, s.d2h, 0));          // previous results of this slot have left the device
    if (int r = byolo_detect(h, s.img, B, seed, image_index0, iou_thr, max_out, nullptr, s.out, nullptr, s.cnt, stream)) return r;
    BY_CUDA(cudaEventRecord(s.done, st));
    BY_CUDA(cudaStreamWaitEvent(h->s_d2h, s.done, 0));
    BY_CUDA(cudaMemcpyAsync(out_rows_host, s.out, out_bytes, cudaMemcpyDeviceToHost, h->s_d2h));
    BY_CUDA(cudaMemcpyAsync(out_count_host, s.cnt, sizeof(int) * B, cudaMemcpyDeviceToHost, h->s_d2h));
    BY_CUDA(cudaEventRecord(s.d2h, h->s_d2h));
    s.used = true;
    return 0;
}

int byolo_wait_host(byolo_handle h, int32_t slot) {
    BY_REQUIRE(h && (slot == 0 || slot == 1), "bad argument");
    BY_REQUIRE(h->slot[slot].used, "nothing was submitted on this slot");
    BY_CUDA(cudaEventSynchronize(h->slot[slot].d2h));
    return 0;
}

int byolo_decode(byolo_handle h, const float* raw0_dev, const float* raw1_dev, const float* raw2_dev, int32_t B,
                 float* rows_dev, void* stream) {
    BY_REQUIRE(h && raw0_dev && raw1_dev && raw2_dev && rows_dev, "null argument");
    const byolo_config& c = h->cfg;
    DecodeProblem dp{};
    dp.variant = c.variant;
    dp.B = B;
    dp.T = h->mc() ? c.T : 1;
    dp.cls_cnt = c.cls_cnt;
    const float* raws[3] = {raw0_dev, raw1_dev, raw2_dev};
    const int det_ch = h->layers.back().cout;
    for (int j = 0; j < 3; ++j) { dp.gh[j] = h->gh[j]; dp.gw[j] = h->gw[j]; dp.raw[j] = raws[j]; dp.ld[j] = det_ch; }
    dp.padded = 0;
    for (int i = 0; i < 9; ++i) { dp.prior_h[i] = c.prior_h[i]; dp.prior_w[i] = c.prior_w[i]; }
    dp.rows = rows_dev;
    dp.N = h->N;
    dp.D = h->D;
    return launch_decode(dp, (cudaStream_t)stream);
}

int byolo_conv_layer(int32_t precision, const float* in1_dev, const float* in2_dev, int32_t S, int32_t H, int32_t W,
                     int32_t cin1, int32_t cin2, int32_t k, int32_t stride, int32_t cout, const float* kernel_host,
                     const float* bn_host, const float* bias_host, const float* residual_dev, int32_t upsample,
                     int32_t dropout_layer, int32_t T, uint64_t seed, int32_t image_index0, float drop_prob, int32_t t1, int32_t t2,
                     float* out_dev, void* stream) {
    BY_REQUIRE(in1_dev && kernel_host && out_dev, "null argument");
    t1 = t1 > 1 ? t1 : 1;
    t2 = t2 > 1 ? t2 : 1;
    BY_REQUIRE(S % t1 == 0 && S % t2 == 0, "S must be a multiple of the stacking factors");
    BY_REQUIRE((t1 == 1 && t2 == 1) || k == 1, "MC-stacked sources only feed 1x1 convs");
    BY_REQUIRE(!(t1 > 1 && in2_dev) && !(t2 > 1 && !in2_dev), "stacked source: in1 alone, or in2 of a concat");
    BY_REQUIRE(precision >= 0 && precision <= 3, "unknown precision");
    BY_REQUIRE((bn_host != nullptr) != (bias_host != nullptr), "pass exactly one of bn / bias");
    cudaStream_t st = (cudaStream_t)stream;
    const bool x3 = precision == BYOLO_PREC_FP16X3, umma = x3 || precision == BYOLO_PREC_FP16;
    const int half = precision == BYOLO_PREC_FP32 ? ACT_F32 : (x3 ? ACT_F16_HILO : ACT_F16);      // activation format
    const size_t es = half == ACT_F16 ? 2 : 4;
    LayerDef d{k, stride, cin1 + cin2, cout, bn_host != nullptr, dropout_layer >= 0};
    LayerWeights w;
    int rc = fold_and_upload(d, kernel_host, bn_host, bias_host, &w, x3);
    void *p1 = nullptr, *p2 = nullptr, *pr = nullptr, *po = nullptr;
    const Geom g1{S / t1, H, W, cin1}, g2{S / t2, H, W, cin2};      // what the caller's buffers hold
    const Geom gc{S, H, W, cin1};                                   // what the conv sees
    const int Ho = H / stride, Wo = W / stride;
    const bool dense = bn_host == nullptr;
    const Geom go{S, upsample ? 2 * Ho : Ho, upsample ? 2 * Wo : Wo, dense ? w.cout_pad : cout};
    float* tmp = nullptr;          // detection conv: output is cout_pad wide
    auto alloc0 = [&](void** p, size_t bytes) {
        if (rc) return;
        if (cudaMalloc(p, bytes) != cudaSuccess || cudaMemsetAsync(*p, 0, bytes, st) != cudaSuccess) {
            set_error("byolo_conv_layer: allocation failed");
            rc = -2;
        }
    };
    const bool stem = cin1 == 3;                  // darknet.py:10: the image conv, fed dense fp32 [S,H,W,3] as it is
    if (stem && !(k == 3 && stride == 1 && cout == 32 && !in2_dev && !residual_dev && !upsample && bn_host && dropout_layer < 0 &&
                  H % 32 == 0 && W % 32 == 0)) {
        set_error("byolo_conv_layer: a 3-channel input is only supported as the stem conv (3x3, stride 1, 32 filters, BN)");
        w.release();
        return -1;
    }
    if (!stem) alloc0(&p1, (size_t)g1.rows() * cin1 * es);
    if (in2_dev) alloc0(&p2, (size_t)g2.rows() * cin2 * es);
    if (residual_dev) alloc0(&pr, (size_t)go.rows() * cout * es);
    alloc0(&po, (size_t)go.rows() * go.C * (dense ? 4 : es));
    if (dense) alloc0((void**)&tmp, (size_t)S * Ho * Wo * go.C * 4);
    if (!rc && !stem) rc = launch_pack(in1_dev, p1, g1, half, st);
    if (!rc && in2_dev) rc = launch_pack(in2_dev, p2, g2, half, st);
    if (!rc && residual_dev) rc = launch_pack(residual_dev, pr, go, half, st);
    if (!rc && stem) {
        rc = umma ? launch_stem_mma(in1_dev, S, H, W, w.w16, w.bias, po, x3, st)
                  : launch_stem(in1_dev, S, H, W, w.w32, w.bias, po, half == ACT_F16, st);
    } else if (!rc) {
        ConvProblem p{};
        // a lone stacked input of a dropout conv takes the T-invariant route on the tensor-core paths, exactly as in the engine's plan
        const bool tinv = umma && t1 > 1 && !in2_dev && k == 1 && dropout_layer >= 0 && !upsample && !dense;
        p.in1 = p1; p.in2 = p2; p.gin = tinv ? g1 : gc; p.c2 = cin2; p.k = k; p.stride = stride;
        p.t_out = tinv ? t1 : 1;
        if (tinv) t1 = 1;
        p.cout_pad = w.cout_pad; p.w16 = w.w16; p.w32 = w.w32; p.x3 = x3;
        p.t1 = t1; p.t2 = t2;
        p.ep.bias = w.bias; p.ep.residual = pr; p.ep.out = po;
        p.ep.out_mode = dense ? OUT_DENSE_F32 : (upsample ? OUT_UPSAMPLE2 : OUT_DENSE);
        p.ep.ldc = go.C; p.ep.cout = cout; p.ep.leaky = d.bn;
        p.ep.drop.enabled = dropout_layer >= 0;
        p.ep.drop.layer_id = dropout_layer; p.ep.drop.T = T > 0 ? T : 1; p.ep.drop.image0 = image_index0;
        p.ep.drop.seed_lo = (uint32_t)seed; p.ep.drop.seed_hi = (uint32_t)(seed >> 32);
        p.ep.drop.thr16 = (uint32_t)std::lround((double)drop_prob * 65536.0);
        p.ep.drop.keep_scale = 1.0f / (1.0f - drop_prob);
        rc = umma ? launch_conv_umma(p, st) : launch_conv_simt(p, half == ACT_F16, st);
    }
    if (!rc) {
        if (dense) {
            rc = launch_unpack(po, tmp, go, ACT_F32, st);
            if (!rc && cudaMemcpy2DAsync(out_dev, (size_t)cout * 4, tmp, (size_t)go.C * 4, (size_t)cout * 4, (size_t)S * Ho * Wo,
                                         cudaMemcpyDeviceToDevice, st) != cudaSuccess)
                rc = -2;
        } else {
            rc = launch_unpack(po, out_dev, go, half, st);
        }
    }
    const cudaError_t se = cudaStreamSynchronize(st);
    if (!rc && se != cudaSuccess) {
        set_error(std::string("byolo_conv_layer: ") + cudaGetErrorString(se));
        rc = -2;
    }
    cudaFree(tmp);
    cudaFree(p1); cudaFree(p2); cudaFree(pr); cudaFree(po);
    w.release();
    return rc;
}

int byolo_get_activation(byolo_handle h, int32_t conv_index, float* dst_dev, size_t capacity, int32_t shape[4], void* stream) {
    BY_REQUIRE(h && dst_dev && shape, "null argument");
    BY_REQUIRE(h->last_plan != nullptr, "no forward pass has run yet");
    Plan* pl = h->last_plan;
    BY_REQUIRE(conv_index >= 0 && conv_index < (int)pl->conv_out.size(), "conv index out of range");
    const Buffer& b = pl->bufs[pl->conv_out[conv_index]];
    shape[0] = b.g.S; shape[1] = b.g.H; shape[2] = b.g.W; shape[3] = b.valid_c;
    const size_t n = (size_t)b.g.S * b.g.H * b.g.W * b.valid_c;
    BY_REQUIRE(capacity >= n, "destination too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (!b.f32) return launch_unpack(b.ptr, dst_dev, b.g, h->act_fmt(), st);
    // raw detection map: fp32, stored cout_pad wide -> drop the padding channels
    float* tmp = nullptr;
    const size_t pix = (size_t)b.g.S * b.g.H * b.g.W;
    BY_CUDA(cudaMalloc(&tmp, pix * b.g.C * 4));
    int rc = launch_unpack(b.ptr, tmp, b.g, ACT_F32, st);
    if (!rc && cudaMemcpy2DAsync(dst_dev, (size_t)b.valid_c * 4, tmp, (size_t)b.g.C * 4, (size_t)b.valid_c * 4, pix,
                                 cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        rc = -2;
    cudaStreamSynchronize(st);
    cudaFree(tmp);
    return rc;
}

int byolo_launch_count(byolo_handle h, int32_t B) {
    BY_REQUIRE(h, "null handle");
    Plan* pl;
    if (int r = get_plan(h, B, &pl)) return r;
    return (int)pl->steps.size() + 2;      // + decode + nms
}

static double flops_per_image(byolo_handle h, bool executed);

double byolo_flops_per_image(byolo_handle h) { return h ? flops_per_image(h, false) : 0.0; }
double byolo_flops_per_image_executed(byolo_handle h) { return h ? flops_per_image(h, true) : 0.0; }

static double flops_per_image(byolo_handle h, bool executed) {
    double backbone = 0, head = 0;
    int Hc = h->cfg.height, Wc = h->cfg.width;
    // spatial size per layer follows the plan: backbone strides, head grids
    int li = 0;
    auto fl = [&](const LayerDef& d, int Ho, int Wo) { return 2.0 * d.k * d.k * d.cin * d.cout * Ho * Wo; };
    backbone += fl(h->layers[li++], Hc, Wc);
    const int blocks[5] = {1, 2, 8, 8, 4};
    for (int s = 0; s < 5; ++s) {
        Hc /= 2; Wc /= 2;
        backbone += fl(h->layers[li++], Hc, Wc);
        for (int b = 0; b < 2 * blocks[s]; ++b) backbone += fl(h->layers[li++], Hc, Wc);
    }
    double once = 0;        // executed only: conv "75" reads a T-invariant input, its GEMM runs once per image on the tensor-core paths
    for (int j = 0; j < 3; ++j) {
        if (j) head += fl(h->layers[li++], h->gh[j - 1], h->gw[j - 1]);
        for (int i = 0; i < 7; ++i) {
            const double f = fl(h->layers[li], h->gh[j], h->gw[j]);
            const bool tinv = executed && j == 0 && i == 0 && h->mc() && h->umma() && h->cfg.T > 1 && h->layers[li].dropout &&
                              !h->cfg.standard_test_dropout;
            (tinv ? once : head) += f;
            ++li;
        }
    }
    // executed MMA work of the split mode: three tensor-core passes per product
    return (backbone + once + head * (h->mc() ? h->cfg.T : 1)) * (executed && h->x3() ? 3.0 : 1.0);
}

}  // extern "C"
