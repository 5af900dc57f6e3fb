// Launch record of the tcgen05 conv kernel (conv_umma.cu): built once per layer of a plan, launched every forward.
#pragma once
#include "common.cuh"

namespace byolo {

// epilogue specialisations of conv_umma_kernel (run_epilogue<KIND>)
enum EpiKind : int {
    EPI_F16 = 0,        // shift + leaky -> fp16, TMA store
    EPI_F16_RES = 1,    //   ... + residual shortcut (layers.py:505-507)
    EPI_F16_DROP = 2,   //   dropout before the shift (layers.py:521-524, 560-570)
    EPI_F32 = 3,        // + bias, linear -> fp32 raw detection map, TMA store (layers.py:600-613)
    EPI_UPSAMPLE = 4,
    EPI_F16_DROP_T = 5, // T-invariant dropout conv (conv "75": input = MC-stacked backbone map, yolov3.py:538-544): the GEMM runs once
                        // per IMAGE, the epilogue applies the T masks and stores the T samples (3D output map)   // shift + leaky -> fp16 stored to the four pixels of the nearest x2 upsample (layers.py:578-580)
};

// how the TMA producer fetches the A operand (activations)
enum AMode : int {
    A_TILED = 0,        // 1x1 conv: plain 2D boxes of the [rows, C] matrix (in1, then in2 for a channel concat)
    A_IM2COL = 1,       // 3x3 conv, stride 1 or 2: 4D im2col map over [C, W, H, S], one load per filter tap, padding = OOB zero fill
    A_STACK1 = 2,       // 1x1 conv whose in1 is an MC-stacked map (layers.py:595-597): 5D im2col map [C, W, H, T, B], T stride 0
    A_STACK2 = 3,       // 1x1 conv over [in1 (2D boxes), in2 = MC-stacked map (5D, T stride 0)]
};

// n / d for n < 2^31 as umulhi(n, mul) >> shr (d == 1: identity)
struct FastDiv {
    uint32_t d, mul, shr;
};

struct UmmaParams {
    int num_m_tiles, num_n_tiles, num_tiles;
    int BN, BK, num_stages;
    int kbs;                       // K blocks per pipeline stage (1 or 2): one barrier round trip / tcgen05.commit per stage
    int cg;                        // CTAs per tile: 1, or 2 (cta_group::2 pair, M = 256)
    int epi_warps;                 // 8 | 16 epilogue warps (block = 4 control warps + these)
    int bsplit;                    // 1: warp 3 issues the B (weight) loads, warp 0 only A; 0: warp 0 issues both
    int b_rows;                    // rows of B each CTA stages (BN / cg)
    int a_bytes, b_bytes;          // bytes of one A / B stage tile
    int taps;                      // 1 | 9
    int kb1, kb2;                  // K blocks per tap read from in1 / in2
    int amode;                     // AMode
    int stride;                    // 1 | 2 (3x3 only)
    Geom gout;                     // output geometry (S samples x H x W, C = valid output channels)
    long long out_rows;            // S*H*W of the output
    Epilogue ep;
    uint32_t idesc;                // tcgen05 instruction descriptor
    uint32_t sbo_bytes, layout_type;
    int epi_kind;                  // EpiKind
    int nnt_shift;                 // log2(num_n_tiles)
    FastDiv fd_plane, fd_w, fd_h, fd_T;   // output row -> (sample, y, x); sample -> (image, MC sample t)
    int x3;                        // split-fp16 mode: operands are hi + lo fp16 pairs, 3 MMAs per K step
    int c1_lo, c2_lo;              // channel coordinate of the lo plane in the in1 / in2 tensor maps (= C1 / C2)
    int b_lo_row;                  // row of the first lo weight row in the B tensor map (= cout_pad)
    int chunk_stages;              // split mode: pipeline stages per accumulator chunk (the epilogue sums the chunks in fp32, RN)
    int chunks_per_tile;
    int dbg;                       // experiments only (-DBYOLO_DBG_HOOKS, BYOLO_DBG): 1 = no operand TMA loads, 2 = no epilogue work, 4 = no MMAs
    unsigned long long* clk;       // profiling: {clock64, globaltimer} at start and end of CTA 0 (effective SM clock), or null
};

struct UmmaLaunch {
    CUtensorMap a1, a2, b, o;
    UmmaParams p;
    int grid;
    int smem_bytes;
};

int umma_prepare(const ConvProblem& q, UmmaLaunch* out);
int umma_launch(const UmmaLaunch& L, cudaStream_t st);

}  // namespace byolo
