// tcn_common.cuh -- the recompute-on-load chains shared by forward and backward of the fused ConvBlock.
// The SAME device functions produce the quantised activations in forward and re-derive them in
// backward, so the STE masks of backward see exactly the codes forward used.
//
// Code-indexed tables.  Everything downstream of an 8-bit quantiser is, per (sample, channel) row, a
// function of the 8-bit code alone: a1 = FQ1(.) takes 256 values, hence so do gLN1(a1), FQ2(gLN1(a1)),
// the STE mask and range-gradient weight of FQ2, the normalised value xhat ...  Each row CTA therefore
// tabulates those chains once (256 entries, exact slow-path arithmetic) and the per-element work is
// "quantise the continuous input, look the rest up" -- one LDS instead of ~15 dependent FP32 ops.
#pragma once
#include "fqss_common.cuh"

namespace fqss {

constexpr int ROW_THREADS = 256;     // default CTA width of the row kernels (one CTA per (sample, channel) row).  Kernels
                                      // whose per-row fixed cost (constants, table build, reductions) dominated at 16 frames
                                      // per thread are instantiated with NTH = 128 instead (measured per kernel on B200).
constexpr float GLN_EPS = 1e-8f;      // convtasnetq.py:8

struct GlnRow {
    float mu, rstd, gamma, scale, shift;
};

// Row constants (written once per block by tcn_rowconst_kernel, tcn_fwd.cu): rc[0..3] / rc[4..7] / rc[8..11] = {min,
// delta, 1/delta, levels} of the quantiser before the gLN / after it / after the next op; rc[12], rc[13] = the exact
// clipping thresholds of the FIRST quantiser in the domain of its input z: z_lo = min{z : (z - min)/delta >= -0.5},
// z_hi = min{z : (z - min)/delta >= levels + 0.5} (the quotient is monotone in z, so lo <= z < hi is the forward's STE
// mask -0.5 <= t < levels + 0.5 bit for bit, without a division per element); rc[RC_HDR+2b], rc[RC_HDR+1+2b] = mean,
// rstd of sample b.  Loading them costs a handful of uniform LDGs per thread instead of fp64 divisions and a square root.
constexpr int RC_HDR = 16;
__device__ __forceinline__ ActQF load_actqf_rc(const float* __restrict__ rc) {
    ActQF q;
    q.mn = __ldg(rc);
    q.delta = __ldg(rc + 1);
    q.inv = __ldg(rc + 2);
    q.levels = __ldg(rc + 3);
    return q;
}
__device__ __forceinline__ GlnRow load_gln_row(const float* __restrict__ rc, int b, const float* __restrict__ gw,
                                               const float* __restrict__ gb, int c) {
    GlnRow g;
    g.mu = __ldg(rc + RC_HDR + 2 * b);
    g.rstd = __ldg(rc + RC_HDR + 1 + 2 * b);
    g.gamma = __ldg(gw + c);
    g.scale = __fmul_rn(g.rstd, g.gamma);
    g.shift = __fadd_rn(__fmul_rn(-g.scale, g.mu), __ldg(gb + c));
    return g;
}

// smallest float z with (z - min) / delta >= target (IEEE division; bit-identical to actqf_t wherever that is finite):
// bisection over the order-preserving integer image of the floats
__device__ __forceinline__ unsigned f2key(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
__device__ inline float fq_threshold(float mn, float delta, float target) {
    unsigned lo = f2key(-3.0e38f), hi = f2key(3.0e38f);
    while (hi - lo > 1u) {
        const unsigned mid = lo + ((hi - lo) >> 1);
        const float t = __fdiv_rn(__fsub_rn(key2f(mid), mn), delta);
        if (t >= target) hi = mid; else lo = mid;
    }
    return key2f(hi);
}

__device__ __forceinline__ float prelu_f(float y, float a) { return y > 0.f ? y : __fmul_rn(a, y); }
__device__ __forceinline__ float gln_apply(const GlnRow& g, float a) { return __fadd_rn(__fmul_rn(a, g.scale), g.shift); }
__device__ __forceinline__ float gln_xhat(const GlnRow& g, float a) { return (a - g.mu) * g.rstd; }

// The row constants are produced by the LAST CTA of the kernel that completes the statistics (ticket counter in the
// slot after the 2*B statistics doubles, zeroed together with them): no extra launch between producer and consumer.
struct RowConstJob {
    double* stats;                       // [2*B] sums + 1 slot used as the arrival counter
    float* rc;                           // [RC_HDR + 2*B]
    int B;
    double n_elems;
    const float *qa_min, *qa_max, *qb_min, *qb_max, *qc_min, *qc_max;   // NULL pairs are skipped
};

// Call from ALL threads of every CTA after the CTA's last contribution to job.stats has been issued.
// `contributed`: this thread issued global atomics on job.stats (only those threads pay for the fence).
__device__ __forceinline__ void rowconst_last_cta(const RowConstJob& j, unsigned total_ctas, bool contributed) {
    __shared__ int is_last;
    if (contributed) __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(reinterpret_cast<unsigned*>(j.stats + 2 * j.B), 1u);
        is_last = (t == total_ctas - 1u);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int t = threadIdx.x;
    if (t < 3) {
        const float* mn = t == 0 ? j.qa_min : (t == 1 ? j.qb_min : j.qc_min);
        const float* mx = t == 0 ? j.qa_max : (t == 1 ? j.qb_max : j.qc_max);
        if (mn) {
            const ActQF q = load_actqf(mn, mx, 8);
            j.rc[4 * t] = q.mn; j.rc[4 * t + 1] = q.delta; j.rc[4 * t + 2] = q.inv; j.rc[4 * t + 3] = q.levels;
        }
    }
    if ((t == 32 || t == 33) && j.qa_min) {      // clipping thresholds of the first quantiser (a warp of its own: 32 dependent divisions)
        const ActQF q = load_actqf(j.qa_min, j.qa_max, 8);
        j.rc[12 + (t - 32)] = fq_threshold(q.mn, q.delta, t == 32 ? -0.5f : q.levels + 0.5f);
    }
    for (int b = t; b < j.B; b += blockDim.x) {
        const double mean = __ldcg(j.stats + 2 * b) / j.n_elems;
        double var = __ldcg(j.stats + 2 * b + 1) / j.n_elems - mean * mean;
        var = var > 0.0 ? var : 0.0;
        j.rc[RC_HDR + 2 * b] = (float)mean;
        j.rc[RC_HDR + 1 + 2 * b] = (float)(1.0 / sqrt(var + (double)GLN_EPS));
    }
}

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
        h.q1 = load_actqf_rc(p.rc1);
        h.q2 = load_actqf_rc(p.rc1 + 4);
    }
    h.g = load_gln_row(p.rc1, b, p.gn1_w, p.gn1_b, c);
    return h;
}
__device__ __forceinline__ float hidden1_a1(const Hidden1& h, float y) {
    float z = prelu_f(y, h.slope);
    return h.quant ? actqf_fq(h.q1, z) : z;
}
__device__ __forceinline__ float hidden1_n1(const Hidden1& h, float a1) { return gln_apply(h.g, a1); }
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
        h.q3 = load_actqf_rc(p.rc3);
        h.q4 = load_actqf_rc(p.rc3 + 4);
    }
    h.g = load_gln_row(p.rc3, b, p.gn2_w, p.gn2_b, c);
    return h;
}
__device__ __forceinline__ float hidden3_a3(const Hidden3& h, float y) {
    float z = prelu_f(y, h.slope);
    return h.quant ? actqf_fq(h.q3, z) : z;
}
__device__ __forceinline__ float hidden3_n3(const Hidden3& h, float a3) { return gln_apply(h.g, a3); }
// GEMM operand of the res/skip conv: the integer code of FQ4 (quant) or the value itself (float model)
__device__ __forceinline__ float hidden3_op(const Hidden3& h, float y) {
    float n3 = hidden3_n3(h, hidden3_a3(h, y));
    return h.quant ? actqf_code(h.q4, n3) : n3;
}

// ---------------------------------------------------------------------------------------------
// code-indexed tables (256 entries; built by the first 256 threads of a row CTA, then __syncthreads)
// ---------------------------------------------------------------------------------------------
// index of the table entry for a continuous pre-quantiser value z (w.r.t. quantiser q); also returns t
__device__ __forceinline__ unsigned code_index(const ActQF& q, float z, float& t, float& biased) {
    t = actqf_t(q, z);
    biased = actqf_biased(q, t);
    return actqf_index(biased);
}
__device__ __forceinline__ unsigned code_index(const ActQF& q, float z) {
    float t, bz;
    return code_index(q, z, t, bz);
}

// forward chain after quantiser A (code i): a = decode_A(i); n = gLN(a); value of FQ_B(n)
__device__ __forceinline__ float chain_fq_value(const ActQF& qa, const GlnRow& g, const ActQF& qb, int i) {
    return actqf_fq(qb, gln_apply(g, actqf_decode(qa, (float)i)));
}
__device__ __forceinline__ float chain_fq_code(const ActQF& qa, const GlnRow& g, const ActQF& qb, int i) {
    return actqf_code(qb, gln_apply(g, actqf_decode(qa, (float)i)));
}
// backward view of the same chain: {value or xhat, STE mask of FQ_B (1/0), range weight D_B, xhat of a}
__device__ __forceinline__ float4 chain_bwd_entry(const ActQF& qa, const GlnRow& g, const ActQF& qb, int i, bool want_value) {
    const float a = actqf_decode(qa, (float)i);
    const float n = gln_apply(g, a);
    const float t = actqf_t(qb, n);
    const bool in = actqf_inside(qb, t);
    const float c = actqf_unbias(actqf_biased(qb, t));
    float4 e;
    e.x = want_value ? actqf_decode(qb, c) : 0.f;
    e.y = in ? 1.f : 0.f;
    e.z = in ? (c - t) : c;
    e.w = gln_xhat(g, a);
    return e;
}

// backward tables: tabX[i] = xhat(decode_A(i)) with the STE mask of FQ_B in the mantissa LSB, tabD[i] = range weight
// D_B (c - t inside, c outside).  One LDS.32 each per element instead of ~20 dependent FP32 ops.
__device__ __forceinline__ void chain_bwd_tables(const ActQF& qa, const GlnRow& g, const ActQF& qb, int i, float* tabX, float* tabD) {
    const float4 e = chain_bwd_entry(qa, g, qb, i, false);
    tabX[i] = __uint_as_float((__float_as_uint(e.w) & ~1u) | (e.y != 0.f ? 1u : 0u));
    tabD[i] = e.z;
}
__device__ __forceinline__ bool tab_mask(float xm) { return (__float_as_uint(xm) & 1u) != 0u; }

// float per-thread partials -> warp sums in fp32 -> cross-warp sums in fp64; result valid in thread 0.
// `sh` needs NV*32 doubles.
template <int NV>
__device__ __forceinline__ void block_sum_fd(const float (&s)[NV], double (&v)[NV], double* sh) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    float w[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) w[i] = warp_sum(s[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) sh[i * 32 + wid] = (double)w[i];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double x = lane < nw ? sh[i * 32 + lane] : 0.0;
            v[i] = warp_sum(x);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// small vector helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 bf16x4_to_float4(uint2 v) {
    const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&v.x);
    const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
    const float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ uint2 float4_to_bf16x4(float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 v;
    v.x = *reinterpret_cast<uint32_t*>(&lo);
    v.y = *reinterpret_cast<uint32_t*>(&hi);
    return v;
}
__device__ __forceinline__ uint2 ldg_bf16x4(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }

// ---------------------------------------------------------------------------------------------
// shared-memory table lookups of the row kernels: tables sit on a (256 << SH)-byte boundary, so the address of the entry
// of `byte k of w` is SHF + LOP3
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float2 lds64(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
// shared-memory address of table entry `byte k of w` : ((w >> s) & (255 << SH)) | base   (base aligned to 256 << SH)
template <int SH>
__device__ __forceinline__ uint32_t tab_addr(uint32_t w, int k, uint32_t base) {
    const int s = 8 * k - SH;
    return ((s >= 0 ? (w >> s) : (w << -s)) & (255u << SH)) | base;
}
// keep a CTA-uniform value in a VECTOR register: the element loops use these as FSEL / LOP3 operands, which cannot read the
// uniform register file, so a uniform-register copy costs one extra move per use
__device__ __forceinline__ float vreg(float x) {
    asm volatile("mov.b32 %0, %0;" : "+f"(x));
    return x;
}
__device__ __forceinline__ uint32_t vreg(uint32_t x) {
    asm volatile("mov.b32 %0, %0;" : "+r"(x));
    return x;
}
// ---------------------------------------------------------------------------------------------
// dilated 3-tap reads from a shared-memory row with a zero halo of dw_pad(d) floats on both sides
// ---------------------------------------------------------------------------------------------
template <int DMODE>
__device__ __forceinline__ void dw_taps(const float* row, int v, int d, float4& L, float4& C, float4& R) {
    const float4* r4 = reinterpret_cast<const float4*>(row);
    C = r4[v];
    if (DMODE == 0) {                    // d % 4 == 0
        L = r4[v - (d >> 2)];
        R = r4[v + (d >> 2)];
    } else if (DMODE == 1) {             // d == 1
        const float4 a = r4[v - 1], b = r4[v + 1];
        L = make_float4(a.w, C.x, C.y, C.z);
        R = make_float4(C.y, C.z, C.w, b.x);
    } else if (DMODE == 2) {             // d == 2
        const float4 a = r4[v - 1], b = r4[v + 1];
        L = make_float4(a.z, a.w, C.x, C.y);
        R = make_float4(C.z, C.w, b.x, b.y);
    } else {                             // any other dilation: scalar reads
        const int m = 4 * v;
        L = make_float4(row[m - d], row[m + 1 - d], row[m + 2 - d], row[m + 3 - d]);
        R = make_float4(row[m + d], row[m + 1 + d], row[m + 2 + d], row[m + 3 + d]);
    }
}

__host__ __device__ inline int dw_pad(int dil) { return (dil + 3) & ~3; }
__host__ __device__ inline int dw_mode(int dil) { return (dil & 3) == 0 ? 0 : (dil == 1 ? 1 : (dil == 2 ? 2 : 3)); }

int tcn_validate_block(const fqss_tcn_block* p, const char* who);

}  // namespace fqss
