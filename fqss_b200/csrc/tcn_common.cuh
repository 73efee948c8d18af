// tcn_common.cuh -- the recompute-on-load chains shared by forward and backward of the fused ConvBlock.
// The SAME device functions produce the quantised activations in forward and re-derive them in
// backward, so the STE masks of backward see exactly the codes forward used.
#pragma once
#include "fqss_common.cuh"

namespace fqss {

constexpr int ROW_THREADS = 256;
constexpr float GLN_EPS = 1e-8f;      // convtasnetq.py:8

struct GlnRow {
    float mu, rstd, gamma, scale, shift;
};

__device__ __forceinline__ GlnRow load_gln_row(const double* __restrict__ stats, int b, double n_elems, const float* __restrict__ gw,
                                               const float* __restrict__ gb, int c) {
    GlnRow g;
    const double mean = stats[2 * b] / n_elems;
    double var = stats[2 * b + 1] / n_elems - mean * mean;
    var = var > 0.0 ? var : 0.0;
    g.mu = (float)mean;
    g.rstd = (float)(1.0 / sqrt(var + (double)GLN_EPS));
    g.gamma = __ldg(gw + c);
    g.scale = __fmul_rn(g.rstd, g.gamma);
    g.shift = __fadd_rn(__fmul_rn(-g.scale, g.mu), __ldg(gb + c));
    return g;
}

__device__ __forceinline__ float prelu_f(float y, float a) { return y > 0.f ? y : __fmul_rn(a, y); }

// ---- y1 -> a1 = FQ1(PReLU(y1)) -> n1 = gLN1(a1) -> a2 = FQ2(n1)
struct Hidden1 {
    int quant;
    float slope;
    ActQF q1, q2;
    GlnRow g;
};

__device__ __forceinline__ Hidden1 load_hidden1(const fqss_tcn_block& p, int b, int c) {
    Hidden1 h;
    h.quant = p.quant;
    h.slope = __ldg(p.slope1);
    if (p.quant) {
        h.q1 = load_actqf(p.q1.rmin, p.q1.rmax, 8);
        h.q2 = load_actqf(p.q2.rmin, p.q2.rmax, 8);
    }
    h.g = load_gln_row(p.stats1, b, (double)p.Chid * (double)p.M, p.gn1_w, p.gn1_b, c);
    return h;
}
__device__ __forceinline__ float hidden1_a1(const Hidden1& h, float y) {
    float z = prelu_f(y, h.slope);
    return h.quant ? actqf_fq(h.q1, z) : z;
}
__device__ __forceinline__ float hidden1_n1(const Hidden1& h, float a1) { return __fadd_rn(__fmul_rn(a1, h.g.scale), h.g.shift); }
__device__ __forceinline__ float hidden1_a2(const Hidden1& h, float y) {
    float n1 = hidden1_n1(h, hidden1_a1(h, y));
    return h.quant ? actqf_fq(h.q2, n1) : n1;
}

// ---- y3 -> a3 = FQ3(PReLU(y3)) -> n3 = gLN2(a3) -> a4 = FQ4(n3)
struct Hidden3 {
    int quant;
    float slope;
    ActQF q3, q4;
    GlnRow g;
};

__device__ __forceinline__ Hidden3 load_hidden3(const fqss_tcn_block& p, int b, int c) {
    Hidden3 h;
    h.quant = p.quant;
    h.slope = __ldg(p.slope3);
    if (p.quant) {
        h.q3 = load_actqf(p.q3.rmin, p.q3.rmax, 8);
        h.q4 = load_actqf(p.q4.rmin, p.q4.rmax, 8);
    }
    h.g = load_gln_row(p.stats3, b, (double)p.Chid * (double)p.M, p.gn2_w, p.gn2_b, c);
    return h;
}
__device__ __forceinline__ float hidden3_a3(const Hidden3& h, float y) {
    float z = prelu_f(y, h.slope);
    return h.quant ? actqf_fq(h.q3, z) : z;
}
__device__ __forceinline__ float hidden3_n3(const Hidden3& h, float a3) { return __fadd_rn(__fmul_rn(a3, h.g.scale), h.g.shift); }
// GEMM operand of the res/skip conv: the integer code of FQ4 (quant) or the value itself (float model)
__device__ __forceinline__ float hidden3_op(const Hidden3& h, float y) {
    float n3 = hidden3_n3(h, hidden3_a3(h, y));
    return h.quant ? actqf_code(h.q4, n3) : n3;
}

int tcn_validate_block(const fqss_tcn_block* p, const char* who);

}  // namespace fqss
