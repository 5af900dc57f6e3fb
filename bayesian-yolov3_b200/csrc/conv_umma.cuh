// Launch record of the tcgen05 conv kernel (conv_umma.cu): built once per layer of a plan, launched every forward.
#pragma once
#include "common.cuh"

namespace byolo {

struct UmmaParams {
    int num_m_tiles, num_n_tiles, num_tiles;
    int BN, BK, num_stages;
    int kbs;                       // K blocks per pipeline stage (1 or 2): one barrier round trip / tcgen05.commit per stage
    int cg;                        // CTAs per tile: 1, or 2 (cta_group::2 pair, M = 256)
    int b_rows;                    // rows of B each CTA stages (BN / cg)
    int a_bytes, b_bytes;          // bytes of one A / B stage tile
    int taps;                      // 1 | 9
    int kb1, kb2;                  // K blocks per tap read from in1 / in2
    int in_PW;                     // padded width of the input (row shift of one filter row)
    int s2;                        // stride-2 patch mode
    int BW, BH, BI, tiles_x, tiles_y;
    Geom gout;                     // un-padded output geometry
    Epilogue ep;
    uint32_t idesc;                // tcgen05 instruction descriptor
    uint32_t sbo_bytes, layout_type;
    int dbg;                       // experiments only (BYOLO_DBG): 1 = no operand TMA loads, 2 = no epilogue work, 4 = no MMAs
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
