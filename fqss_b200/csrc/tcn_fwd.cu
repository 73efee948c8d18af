// tcn_fwd.cu -- forward of the fused TCN ConvBlock (convtasnetq.py:11-42 after quantize_model).
//
//   K1  tcgen05 GEMM  x_op -> y1 = W1q x + b1            epilogue: exact FQ1 codes (code1, 1 B/frame) and the gLN
//                     statistics of a1 = FQ1(PReLU(y1)) as integer code sums; last CTA -> per-sample constants
//   K2  this file     code1 -> table[gLN1, FQ2] -> depthwise dilated 3-tap FIR + bias -> y3, code3
//                     one CTA per (sample, channel) row; the whole row (M <= 8K frames) is staged in shared memory
//                     so the dilation halo costs no extra HBM traffic; epilogue: gLN statistics of
//                     a3 = FQ3(PReLU(y3)).  HBM bytes: code1 1 + y3 4 + code3 1 = 6 B/frame.
//                     Float model: y1 -> PReLU -> gLN1 -> FIR -> y3 (or, with the second gLN folded into the
//                     res/skip conv, a3 = PReLU(y3) straight into the [hi ; lo] GEMM operand).
//   K3a this file     code3 -> table[gLN2, FQ4] -> bf16 operand (integer code)          3 B/frame
//   K3  tcgen05 GEMM  a4_op -> res_y / skip_y, x_out = FQ(x + FQ(res)), skip_out = FQ(skip_in + FQ(skip))
//
// Stored per block: the pre-activations backward re-reads (y1, y3, res_y, skip_y; skipped in inference), the
// 8-bit codes of the two hidden quantisers and the 128-wide block outputs; every other quantised activation is
// looked up / recomputed by its consumer (and again in backward) from the same device functions.
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "fqss_common.cuh"
#include "gemm_tc.cuh"
#include "tcn_common.cuh"

namespace fqss {

// ---------------------------------------------------------------------------------------------
// weight preparation (tiny; one CTA per output channel)
// ---------------------------------------------------------------------------------------------
__global__ void tcn_prep_kernel(const float* __restrict__ W, const float* __restrict__ wmin, const float* __restrict__ wmax,
                                const float* __restrict__ bias, const float* __restrict__ amin, const float* __restrict__ amax,
                                __nv_bfloat16* __restrict__ Wc, __nv_bfloat16* __restrict__ WcT, float* __restrict__ s1,
                                float* __restrict__ s0, float* __restrict__ dws, int K, int Ntot, int n_off, int split) {
    __shared__ double sh[32];
    const int o = blockIdx.x;
    const bool quant = wmin != nullptr;
    WQ q;
    if (quant) q = make_wq(wmin[o], wmax[o], 8);
    double rsum = 0.0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float w = W[(int64_t)o * K + k];
        float c = quant ? wq_code(q, w) : w;
        __nv_bfloat16 cb = __float2bfloat16_rn(c);          // codes -128..127 are exact in bf16
        if (split) {                                        // float model, fp32-grade: [hi | hi | lo]
            __nv_bfloat16* row = Wc + (int64_t)(n_off + o) * 3 * K;
            row[k] = cb;
            row[K + k] = cb;
            row[2 * K + k] = __float2bfloat16_rn(c - __bfloat162float(cb));
        } else {
            Wc[(int64_t)(n_off + o) * K + k] = cb;
        }
        if (WcT) WcT[(int64_t)k * Ntot + n_off + o] = cb;
        rsum += (double)c;
    }
    double v[1] = {rsum};
    block_sum<1>(v, sh);
    if (threadIdx.x == 0) {
        float dw = quant ? q.delta : 1.f;
        float da = 1.f, mn = 0.f;
        if (amin) {
            mn = *amin;
            da = __fdiv_rn(__fsub_rn(*amax, mn), 255.f);
        }
        s1[n_off + o] = dw * da;
        s0[n_off + o] = dw * mn * (float)v[0] + (bias ? bias[o] : 0.f);
        dws[n_off + o] = dw;
    }
}

constexpr int PREP_BATCH = 32;
struct PrepBatch {
    fqss_prep_item it[PREP_BATCH];
};
__global__ void tcn_prep_batch_kernel(const __grid_constant__ PrepBatch b) {
    const fqss_prep_item& t = b.it[blockIdx.y];
    if ((int)blockIdx.x >= t.N) return;
    __shared__ double sh[32];
    const int o = blockIdx.x, K = t.K;
    const bool quant = t.wmin != nullptr;
    WQ q;
    if (quant) q = make_wq(t.wmin[o], t.wmax[o], 8);
    __nv_bfloat16* Wc = reinterpret_cast<__nv_bfloat16*>(t.Wc);
    __nv_bfloat16* WcT = reinterpret_cast<__nv_bfloat16*>(t.WcT);
    double rsum = 0.0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float w = t.W[(int64_t)o * K + k];
        const float c = quant ? wq_code(q, w) : w;
        const __nv_bfloat16 cb = __float2bfloat16_rn(c);
        if (t.split) {
            __nv_bfloat16* row = Wc + (int64_t)(t.n_off + o) * 3 * K;
            row[k] = cb;
            row[K + k] = cb;
            row[2 * K + k] = __float2bfloat16_rn(c - __bfloat162float(cb));
        } else {
            Wc[(int64_t)(t.n_off + o) * K + k] = cb;
        }
        if (WcT) WcT[(int64_t)k * t.Ntot + t.n_off + o] = cb;
        rsum += (double)c;
    }
    double v[1] = {rsum};
    block_sum<1>(v, sh);
    if (threadIdx.x == 0) {
        const float dw = quant ? q.delta : 1.f;
        float da = 1.f, mn = 0.f;
        if (t.amin) {
            mn = *t.amin;
            da = __fdiv_rn(__fsub_rn(*t.amax, mn), 255.f);
        }
        t.s1[t.n_off + o] = dw * da;
        t.s0[t.n_off + o] = dw * mn * (float)v[0] + (t.bias ? t.bias[o] : 0.f);
        t.dws[t.n_off + o] = dw;
    }
}

// Float (teacher) model, inference: the second gLN is folded into the res/skip conv it feeds.  With a3 = PReLU(y3),
// n3[c] = rstd_b * gamma_c * (a3[c] - mu_b) + beta_c and y[o] = sum_c W[o,c] n3[c] + bias[o]:
//   y[o] = rstd_b * (sum_c (W[o,c] gamma_c) a3[c]  -  mu_b * u[o]) + v[o],   u[o] = sum_c W[o,c] gamma_c,
//                                                                         v[o] = bias[o] + sum_c W[o,c] beta_c
// so the depthwise kernel can hand a3 to the GEMM directly (no normalisation pass over the hidden tensor).
// Wc = [hi | hi | lo] split of W*gamma; u -> s1, v -> s0 (fp64 sums).
__global__ void tcn_prep_fold_kernel(const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, __nv_bfloat16* __restrict__ Wc, float* __restrict__ u,
                                     float* __restrict__ v, int K, int n_off) {
    __shared__ double sh[2 * 32];
    const int o = blockIdx.x;
    double su = 0.0, sv = 0.0;
    __nv_bfloat16* row = Wc + (int64_t)(n_off + o) * 3 * K;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float w = W[(int64_t)o * K + k];
        const float wg = __fmul_rn(w, gamma[k]);
        const __nv_bfloat16 hi = __float2bfloat16_rn(wg);
        row[k] = hi;
        row[K + k] = hi;
        row[2 * K + k] = __float2bfloat16_rn(wg - __bfloat162float(hi));
        su += (double)w * (double)gamma[k];
        sv += (double)w * (double)beta[k];
    }
    double vv[2] = {su, sv};
    block_sum<2>(vv, sh);
    if (threadIdx.x == 0) {
        u[n_off + o] = (float)vv[0];
        v[n_off + o] = (float)(vv[1] + (bias ? (double)bias[o] : 0.0));
    }
}

// ---------------------------------------------------------------------------------------------
// K2: depthwise kernel.  One CTA per (sample, channel) row.
//   phase 0  (quantised model) tabulate code1 -> a2 = FQ2(gLN1(decode1(code1)))        256 entries
//   phase 1  y1 (128-bit loads) -> PReLU -> code1 -> table -> a2 row in shared memory, zero halo of
//            `dil` frames on both sides (so the taps need no boundary predicates)
//   phase 2  3-tap dilated FIR from shared memory (128-bit reads) + bias -> y3 (128-bit stores);
//            gLN statistics of a3 = FQ3(PReLU(y3)) for the next normalisation
// HBM bytes: read y1 + write y3 = 8 B/element.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4& v) { return make_float2(v.z, v.w); }

template <bool QUANT, int DMODE, int NTH>
__global__ void __launch_bounds__(NTH) tcn_dw_fwd_kernel(const fqss_tcn_block p) {
    // dynamic shared memory: [slack to a 1 KB boundary | table 1 KB | dpad | ld | dpad]: a2 row with a zero halo
    extern __shared__ __align__(16) uint8_t dsm_raw[];
    __shared__ double sh[2 * 32];
    __shared__ unsigned shi[2 * 8];
    const int64_t r = blockIdx.x;
    const int b = (int)(r / p.Chid), c = (int)(r % p.Chid);
    const int M = p.M, d = p.dil, dpad = dw_pad(d);
    const int ld = (int)p.ld;
    const uint32_t raw_s = smem_addr(dsm_raw);
    const uint32_t tab_s = (raw_s + 1023u) & ~1023u;
    float* lut = reinterpret_cast<float*>(dsm_raw + (tab_s - raw_s));
    float* dsm = lut + 256;
    float* row = dsm + dpad;
    const Hidden1 h = load_hidden1(p, b, c);
    for (int i = threadIdx.x; i < dpad; i += NTH) {
        dsm[i] = 0.f;
        row[ld + i] = 0.f;
    }
    if (QUANT) {
        for (int i = threadIdx.x; i < 256; i += NTH) lut[i] = chain_fq_value(h.q1, h.g, h.q2, i);
        __syncthreads();
    }
    const float* y1 = p.y1 + r * p.ld;
    const uint32_t* code1 = reinterpret_cast<const uint32_t*>(p.code1 + r * p.ld);      // written by the expand GEMM's epilogue
    const int nvec = ld >> 2, nfull = M >> 2;
    auto put = [&](int v, float4 a) {
        if (4 * v + 3 >= M) {            // ragged tail / pad columns: the FIR must see zeros there
            if (4 * v + 0 >= M) a.x = 0.f;
            if (4 * v + 1 >= M) a.y = 0.f;
            if (4 * v + 2 >= M) a.z = 0.f;
            a.w = 0.f;
        }
        *reinterpret_cast<float4*>(row + 4 * v) = a;
    };
    if (QUANT) {
        // code words first (4 in flight per thread: a row is ~1000 of them), then one table lookup per frame (SHF + LOP3 +
        // LDS); the full quads run without frame checks, the ragged / pad quads (at most two) go through put()
        constexpr int NQ = 8;
        const uint32_t tb = vreg(tab_s);
        auto look = [&](uint32_t w) {
            return make_float4(lds32(tab_addr<2>(w, 0, tb)), lds32(tab_addr<2>(w, 1, tb)), lds32(tab_addr<2>(w, 2, tb)),
                               lds32(tab_addr<2>(w, 3, tb)));
        };
        for (int base = threadIdx.x; base < nfull; base += NQ * NTH) {
            uint32_t cw[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) cw[q] = (base + q * NTH < nfull) ? __ldg(code1 + base + q * NTH) : 0u;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int v = base + q * NTH;
                if (v < nfull) *reinterpret_cast<float4*>(row + 4 * v) = look(cw[q]);
            }
        }
        for (int v = nfull + threadIdx.x; v < nvec; v += NTH) put(v, look(__ldg(code1 + v)));
    } else {
        constexpr int NQ = 8;
        auto norm = [&](const float4 y) {
            return make_float4(gln_apply(h.g, prelu_f(y.x, h.slope)), gln_apply(h.g, prelu_f(y.y, h.slope)),
                               gln_apply(h.g, prelu_f(y.z, h.slope)), gln_apply(h.g, prelu_f(y.w, h.slope)));
        };
        for (int base = threadIdx.x; base < nfull; base += NQ * NTH) {
            float4 yv[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                yv[q] = (base + q * NTH < nfull) ? ldg4_stream(y1 + 4 * (base + q * NTH)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int v = base + q * NTH;
                if (v < nfull) *reinterpret_cast<float4*>(row + 4 * v) = norm(yv[q]);
            }
        }
        for (int v = nfull + threadIdx.x; v < nvec; v += NTH) put(v, norm(ldg4_stream(y1 + 4 * v)));
    }
    __syncthreads();
    const float2 w0 = f2s(__ldg(p.wdw + c * 3)), w1 = f2s(__ldg(p.wdw + c * 3 + 1)), w2 = f2s(__ldg(p.wdw + c * 3 + 2));
    const float2 bias = f2s(__ldg(p.bdw + c));
    const float slope3 = __ldg(p.slope3);
    ActQF q3;
    if (QUANT) q3 = load_actqf_rc(p.rc1 + 8);
    float* y3 = p.y3 + r * p.ld;
    uint32_t* code3 = (QUANT && p.code3) ? reinterpret_cast<uint32_t*>(p.code3 + r * p.ld) : nullptr;
    const bool fold = !QUANT && p.split == 2;
    __nv_bfloat16* a3_hi = reinterpret_cast<__nv_bfloat16*>(p.a4_op) + ((int64_t)b * 2 * p.Chid + c) * p.ld;
    __nv_bfloat16* a3_lo = a3_hi + (int64_t)p.Chid * p.ld;
    // statistics of a3 = FQ3(PReLU(y3)): the quantised model accumulates the integer codes (sum c, sum c^2, exact; one
    // IDP.4A each on the packed code word), a3 = delta*c + min is expanded at the end; the float model sums the values
    float s = 0.f, ss = 0.f;
    unsigned sc = 0u, scc = 0u;
    const bool st_y3 = p.y3 != nullptr;      // NULL in quantised inference: only backward reads y3 (the codes carry on)
    auto body = [&](int v, auto tail_tag) {
        constexpr bool TAIL = decltype(tail_tag)::value;
        float4 L, C, R;
        dw_taps<DMODE>(row, v, d, L, C, R);
        const float2 o01 = __ffma2_rn(w2, lo2(R), __ffma2_rn(w1, lo2(C), __ffma2_rn(w0, lo2(L), bias)));
        const float2 o23 = __ffma2_rn(w2, hi2(R), __ffma2_rn(w1, hi2(C), __ffma2_rn(w0, hi2(L), bias)));
        const float2 z01 = make_float2(prelu_f(o01.x, slope3), prelu_f(o01.y, slope3));
        const float2 z23 = make_float2(prelu_f(o23.x, slope3), prelu_f(o23.y, slope3));
        if (!QUANT && fold) {
            // gLN2 is folded into the res/skip GEMM (tcn_prep_fold_kernel): its operand is a3 itself, as a [hi ; lo] pair
            const uint2 pk = float4_to_bf16x4(z01.x, z01.y, z23.x, z23.y);
            const float4 hv = bf16x4_to_float4(pk);
            *reinterpret_cast<uint2*>(a3_hi + 4 * v) = pk;
            *reinterpret_cast<uint2*>(a3_lo + 4 * v) = float4_to_bf16x4(z01.x - hv.x, z01.y - hv.y, z23.x - hv.z, z23.y - hv.w);
        } else if (st_y3) {
            stg4(y3 + 4 * v, make_float4(o01.x, o01.y, o23.x, o23.y));
        }
        if (QUANT) {
            // exact codes of FQ3 (same arithmetic as every later consumer): saved here so that the hidden quantiser and
            // the backward sums read 1 B/frame instead of re-deriving the code from y3
            const float2 t01 = actqf_t2(q3, z01), t23 = actqf_t2(q3, z23);
            unsigned c0 = code_u8(t01.x), c1 = code_u8(t01.y), c2 = code_u8(t23.x), c3 = code_u8(t23.y);
            if (TAIL) {                  // statistics over valid frames only; codes at frames >= M are stored as 0
                const int nval = M - 4 * v;
                c3 = 0u;
                if (nval < 3) c2 = 0u;
                if (nval < 2) c1 = 0u;
                if (nval < 1) c0 = 0u;
            }
            const unsigned word = __byte_perm(__byte_perm(c0, c1, 0x1140), __byte_perm(c2, c3, 0x1140), 0x5410);
            if (code3) code3[v] = word;
            sc = __dp4a(word, 0x01010101u, sc);
            scc = __dp4a(word, word, scc);
        } else {
            const int nval = TAIL ? M - 4 * v : 4;
            float z[4] = {z01.x, z01.y, z23.x, z23.y};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float a3 = k < nval ? z[k] : 0.f;
                s += a3;
                ss = fmaf(a3, a3, ss);
            }
        }
    };
    for (int v = threadIdx.x; v < nfull; v += NTH) body(v, std::false_type{});
    for (int v = nfull + threadIdx.x; v < nvec; v += NTH) body(v, std::true_type{});
    if (QUANT) {
        sc = __reduce_add_sync(0xffffffffu, sc);
        scc = __reduce_add_sync(0xffffffffu, scc);
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) { shi[wid] = sc; shi[8 + wid] = scc; }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long tc = 0, tcc = 0;
#pragma unroll
            for (int w = 0; w < NTH / 32; ++w) { tc += shi[w]; tcc += shi[8 + w]; }
            const double dl = (double)q3.delta, mn = (double)q3.mn, n = (double)M;
            atomicAdd(p.stats3 + 2 * b, dl * (double)tc + n * mn);
            atomicAdd(p.stats3 + 2 * b + 1, dl * dl * (double)tcc + 2.0 * dl * mn * (double)tc + n * mn * mn);
        }
    } else {
        double vv[2] = {(double)s, (double)ss};
        block_sum<2>(vv, sh);
        if (threadIdx.x == 0) {
            atomicAdd(p.stats3 + 2 * b, vv[0]);
            atomicAdd(p.stats3 + 2 * b + 1, vv[1]);
        }
    }
}

// Row constants after K2: with 16 K short-lived row CTAs a per-CTA arrival ticket costs more than this one-warp launch
// (measured: +20 us on the depthwise kernel vs 6.5 us here); the GEMM (148 persistent CTAs) finalises in its last CTA.
__global__ void tcn_rowconst_kernel(RowConstJob j) {
    const int t = threadIdx.x;
    if (t < 3) {
        const float* mn = t == 0 ? j.qa_min : (t == 1 ? j.qb_min : j.qc_min);
        const float* mx = t == 0 ? j.qa_max : (t == 1 ? j.qb_max : j.qc_max);
        if (mn) {
            const ActQF q = load_actqf(mn, mx, 8);
            j.rc[4 * t] = q.mn; j.rc[4 * t + 1] = q.delta; j.rc[4 * t + 2] = q.inv; j.rc[4 * t + 3] = q.levels;
        }
    }
    if ((t == 32 || t == 33) && j.qa_min) {      // clipping thresholds of the first quantiser (tcn_common.cuh)
        const ActQF q = load_actqf(j.qa_min, j.qa_max, 8);
        j.rc[12 + (t - 32)] = fq_threshold(q.mn, q.delta, t == 32 ? -0.5f : q.levels + 0.5f);
    }
    for (int b = t; b < j.B; b += blockDim.x) {
        const double mean = j.stats[2 * b] / j.n_elems;
        double var = j.stats[2 * b + 1] / j.n_elems - mean * mean;
        var = var > 0.0 ? var : 0.0;
        j.rc[RC_HDR + 2 * b] = (float)mean;
        j.rc[RC_HDR + 1 + 2 * b] = (float)(1.0 / sqrt(var + (double)GLN_EPS));
    }
}

// ---------------------------------------------------------------------------------------------
// K3a: hidden quantiser -> a4 operand (bf16 code / value).  Quantised model: code3 (saved by K2) -> per-row table
// (code3 -> code4 = FQ4-code(gLN2(decode3(code3)))): 3 B/element.  Float model: y3 -> PReLU -> gLN2 -> [hi ; lo] pair.
// ---------------------------------------------------------------------------------------------
template <bool QUANT, int NTH>
__global__ void __launch_bounds__(NTH) tcn_hidden_fq_kernel(const fqss_tcn_block p) {
    __shared__ float lut[256];
    const int64_t r = blockIdx.x;
    const int b = (int)(r / p.Chid), c = (int)(r % p.Chid);
    const Hidden3 h = load_hidden3(p, b, c);
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.a4_op) + r * p.ld;
    if (QUANT) {
        // code3 (saved by the depthwise kernel) -> code4 through the per-row table: 1 B in, 2 B out per frame.
        // The table holds the bf16 bit pattern of code4, so a frame costs one LDS + one PRMT.
        unsigned short* lut16 = reinterpret_cast<unsigned short*>(lut);
        for (int i = threadIdx.x; i < 256; i += NTH)
            lut16[i] = __bfloat16_as_ushort(__float2bfloat16_rn(chain_fq_code(h.q3, h.g, h.q4, i)));
        __syncthreads();
        const uint2* c3 = reinterpret_cast<const uint2*>(p.code3 + r * p.ld);       // rows are 8-byte aligned (ld % 8 == 0)
        uint4* o16 = reinterpret_cast<uint4*>(out);
        const int n8 = (p.M + 7) >> 3;
        constexpr int NQ = 4;                       // code words in flight per thread (a row is only ~500 of them)
        for (int base = threadIdx.x; base < n8; base += NQ * NTH) {
            uint2 cw[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int v = base + q * NTH;
                cw[q] = v < n8 ? __ldg(c3 + v) : make_uint2(0u, 0u);
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int v = base + q * NTH;
                if (v >= n8) continue;
                const uint32_t w[2] = {cw[q].x, cw[q].y};
                uint32_t o[4];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const uint32_t a0 = lut16[w[k] & 255u], a1 = lut16[(w[k] >> 8) & 255u];
                    const uint32_t a2 = lut16[(w[k] >> 16) & 255u], a3 = lut16[w[k] >> 24];
                    o[2 * k] = a0 | (a1 << 16);
                    o[2 * k + 1] = a2 | (a3 << 16);
                }
                o16[v] = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
        return;
    }
    const float* y3 = p.y3 + r * p.ld;
    const int nvec = (p.M + 3) >> 2;
    for (int v = threadIdx.x; v < nvec; v += NTH) {
        const float4 y = ldg4_stream(y3 + 4 * v);
        const float o0 = gln_apply(h.g, prelu_f(y.x, h.slope));
        const float o1 = gln_apply(h.g, prelu_f(y.y, h.slope));
        const float o2 = gln_apply(h.g, prelu_f(y.z, h.slope));
        const float o3 = gln_apply(h.g, prelu_f(y.w, h.slope));
        const uint2 pk = float4_to_bf16x4(o0, o1, o2, o3);
        if (!p.split) {
            *reinterpret_cast<uint2*>(out + 4 * v) = pk;
        } else {
            // [hi ; lo] pair: sample b owns rows [2*b*Chid, 2*(b+1)*Chid); residual = value - bf16(value)
            __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(p.a4_op) + ((int64_t)b * 2 * p.Chid + c) * p.ld;
            __nv_bfloat16* ol = oh + (int64_t)p.Chid * p.ld;
            *reinterpret_cast<uint2*>(oh + 4 * v) = pk;
            const float4 hv = bf16x4_to_float4(pk);
            *reinterpret_cast<uint2*>(ol + 4 * v) = float4_to_bf16x4(o0 - hv.x, o1 - hv.y, o2 - hv.z, o3 - hv.w);
        }
    }
}

// fp32 -> [hi | hi | lo] (weights, layout 0) or per-sample [hi rows ; lo rows] (activations, layout 1)
__global__ void __launch_bounds__(ROW_THREADS) split_bf16_kernel(const float* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ out,
                                                                int64_t ldo, int cols, int C, int layout) {
    const int64_t r = blockIdx.x;
    const float* src = x + r * ldx;
    if (layout == 0) {
        __nv_bfloat16* row = out + r * 3 * (int64_t)cols;
        for (int k = threadIdx.x; k < cols; k += ROW_THREADS) {
            const float v = src[k];
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            row[k] = hi;
            row[cols + k] = hi;
            row[2 * cols + k] = __float2bfloat16_rn(v - __bfloat162float(hi));
        }
    } else {
        const int64_t b = r / C, c = r % C;
        __nv_bfloat16* oh = out + (b * 2 * C + c) * ldo;
        __nv_bfloat16* ol = oh + (int64_t)C * ldo;
        for (int m = threadIdx.x; m < cols; m += ROW_THREADS) {
            const float v = src[m];
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            oh[m] = hi;
            ol[m] = __float2bfloat16_rn(v - __bfloat162float(hi));
        }
    }
}

// out[r][m] = bf16(scale[r % C] * g[r][m]), rowsum[r % C] += sum_m g[r][m] (fp64, unscaled): turns an fp32 output
// gradient into the pre-scaled bf16 operand of the dgrad / wgrad GEMMs of a code-operand 1x1 conv and its bias grad
__global__ void __launch_bounds__(ROW_THREADS) rowscale_bf16_kernel(const float* __restrict__ g, int64_t ldg, __nv_bfloat16* __restrict__ out,
                                                                   int64_t ldo, int M, int C, const float* __restrict__ scale,
                                                                   double* __restrict__ rowsum) {
    __shared__ double sh[32];
    const int64_t r = blockIdx.x;
    const int ch = (int)(r % C);
    const float sc = __ldg(scale + ch);
    const float4* src = reinterpret_cast<const float4*>(g + r * ldg);
    uint2* dst = reinterpret_cast<uint2*>(out + r * ldo);
    float s = 0.f;
    const int nfull = M >> 2;
    for (int v = threadIdx.x; v < nfull; v += ROW_THREADS) {
        const float4 x = ldg4_stream(reinterpret_cast<const float*>(src + v));
        s += (x.x + x.y) + (x.z + x.w);
        dst[v] = float4_to_bf16x4(x.x * sc, x.y * sc, x.z * sc, x.w * sc);
    }
    if ((M & 3) && (int)threadIdx.x == (nfull % ROW_THREADS)) {
        float x[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < (M & 3); ++k) x[k] = g[r * ldg + 4 * nfull + k];
        s += (x[0] + x[1]) + (x[2] + x[3]);
        dst[nfull] = float4_to_bf16x4(x[0] * sc, x[1] * sc, x[2] * sc, x[3] * sc);
    }
    double v[1] = {(double)warp_sum(s)};
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sh[wid] = v[0];
    __syncthreads();
    if (threadIdx.x == 0 && rowsum) {
        double t = 0.0;
        for (int w = 0; w < ROW_THREADS / 32; ++w) t += sh[w];
        atomicAdd(rowsum + ch, t);
    }
}

// values -> GEMM operand of the first block: integer code w.r.t. the producer's quantiser, or the value itself
__global__ void __launch_bounds__(ROW_THREADS) tcn_encode_kernel(const float* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ out,
                                                                int64_t ldo, int M, const float* rmin, const float* rmax) {
    const int64_t r = blockIdx.x;
    ActQF q;
    const bool quant = rmin != nullptr;
    if (quant) q = load_actqf(rmin, rmax, 8);
    const float* xr = x + r * ldx;
    __nv_bfloat16* orow = out + r * ldo;
    // 128-bit loads / 64-bit stores over the full frame quads of aligned rows (the scalar loop ran at ~0.3 of HBM), scalar tail
    const bool vec = (ldx & 3) == 0 && (ldo & 3) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 7) == 0;
    const int nfull = vec ? (M >> 2) : 0;
    for (int v4 = threadIdx.x; v4 < nfull; v4 += ROW_THREADS) {
        const float4 v = ldg4(xr + 4 * v4);
        const float c0 = quant ? actqf_code(q, v.x) : v.x, c1 = quant ? actqf_code(q, v.y) : v.y;
        const float c2 = quant ? actqf_code(q, v.z) : v.z, c3 = quant ? actqf_code(q, v.w) : v.w;
        *reinterpret_cast<uint2*>(orow + 4 * v4) = float4_to_bf16x4(c0, c1, c2, c3);
    }
    for (int m = 4 * nfull + threadIdx.x; m < M; m += ROW_THREADS) {
        float v = xr[m];
        orow[m] = __float2bfloat16_rn(quant ? actqf_code(q, v) : v);
    }
}

static int validate_block(const fqss_tcn_block* p, const char* who) {
    FQSS_REQUIRE(p, -1, "%s: null block", who);
    FQSS_REQUIRE(p->B > 0 && p->M > 0 && p->dil >= 1 && p->Cio > 0 && p->Chid > 0, -1, "%s: bad shape", who);
    FQSS_REQUIRE(p->Cio % 64 == 0 && p->Chid % 128 == 0 && p->Cio % 128 == 0, -1,
                 "%s: fused path needs Cio %% 128 == 0 and Chid %% 128 == 0 (got %d, %d)", who, p->Cio, p->Chid);
    FQSS_REQUIRE(p->ld >= p->M && p->ld % 8 == 0, -2, "%s: ld must be >= M and a multiple of 8", who);
    FQSS_REQUIRE(((size_t)p->ld * 2 + 4 * (size_t)((p->dil + 3) & ~3) + 1280) * sizeof(float) + (size_t)p->ld <= 200 * 1024, -1,
                 "%s: row too long for shared-memory staging (M=%d, dil=%d)", who, p->M, p->dil);
    FQSS_REQUIRE(p->Wc1 && p->Wc2 && p->s1_1 && p->s0_1 && p->s1_2 && p->s0_2 && p->wdw && p->bdw, -1, "%s: block not prepared", who);
    FQSS_REQUIRE(p->no_skip == 0 || (p->no_skip == 1 && p->has_res), -1, "%s: no_skip must be 0 or 1, and a skip-less block needs the residual path", who);
    FQSS_REQUIRE(p->slope1 && p->slope3 && p->gn1_w && p->gn1_b && p->gn2_w && p->gn2_b, -1, "%s: missing layer parameters", who);
    FQSS_REQUIRE(p->x_op && p->x_in && (p->y1 || p->quant) && (p->y3 || p->split == 2 || p->quant) && p->stats1 && p->stats3 && p->a4_op && (p->skip_out || p->no_skip) && p->rc1 && p->rc3, -1,
                 "%s: missing activation buffers", who);
    FQSS_REQUIRE(!p->split || !p->quant, -1, "%s: split operands are a float-model (quant == 0) feature", who);
    FQSS_REQUIRE(p->split >= 0 && p->split <= 2, -1, "%s: split must be 0, 1 or 2", who);
    if (p->has_res) FQSS_REQUIRE(p->x_out && p->x_out_op, -1, "%s: missing residual buffers", who);
    if (!p->first_block && !p->no_skip) FQSS_REQUIRE(p->skip_in, -1, "%s: missing skip_in", who);
    if (p->quant) {
        FQSS_REQUIRE(p->code1 && p->code3, -1, "%s: the quantised path needs the code buffers (the expand GEMM hands FQ1's codes to the depthwise kernel through code1, the depthwise kernel FQ3's codes to the hidden quantiser through code3)", who);
        const fqss_qrange* qs[] = {&p->q1, &p->q2, &p->q3, &p->q4, &p->qskip};
        for (auto q : qs) FQSS_REQUIRE((q->rmin && q->rmax) || (q == &p->qskip && p->no_skip), -1, "%s: missing quantiser range", who);
        if (p->has_res) FQSS_REQUIRE(p->qres.rmin && p->qadd.rmin, -1, "%s: missing residual quantisers", who);
        if (!p->first_block && !p->no_skip) FQSS_REQUIRE(p->qadds.rmin, -1, "%s: missing skip-sum quantiser", who);
    }
    return 0;
}

int tcn_validate_block(const fqss_tcn_block* p, const char* who) { return validate_block(p, who); }

}  // namespace fqss

using namespace fqss;

extern "C" {

int fqss_tcn_prep(const float* W, const float* wmin, const float* wmax, const float* bias, const float* amin, const float* amax,
                  void* Wc, void* WcT, float* s1, float* s0, float* dws, int N, int K, int Ntot, int n_off, int split, void* stream) {
    FQSS_REQUIRE(W && Wc && s1 && s0 && dws && N > 0 && K > 0 && n_off >= 0 && n_off + N <= Ntot, -1, "tcn_prep: bad argument");
    FQSS_REQUIRE(WcT || split, -1, "tcn_prep: WcT may be NULL only for split (inference) operands");
    FQSS_REQUIRE(!split || (wmin == nullptr && amin == nullptr), -1, "tcn_prep: split operands are for the float model");
    FQSS_REQUIRE((wmin == nullptr) == (wmax == nullptr) && (amin == nullptr) == (amax == nullptr), -1, "tcn_prep: ranges come in pairs");
    FQSS_PROF("tcn_prep", stream);
    tcn_prep_kernel<<<N, 128, 0, (cudaStream_t)stream>>>(W, wmin, wmax, bias, amin, amax, (__nv_bfloat16*)Wc, (__nv_bfloat16*)WcT, s1,
                                                         s0, dws, K, Ntot, n_off, split);
    return check_launch("tcn_prep");
}

int fqss_tcn_prep_fold(const float* W, const float* bias, const float* gamma, const float* beta, void* Wc, float* u, float* v, int N,
                       int K, int Ntot, int n_off, void* stream) {
    FQSS_REQUIRE(W && gamma && beta && Wc && u && v && N > 0 && K > 0 && n_off >= 0 && n_off + N <= Ntot, -1, "tcn_prep_fold: bad argument");
    FQSS_PROF("tcn_prep", stream);
    tcn_prep_fold_kernel<<<N, 128, 0, (cudaStream_t)stream>>>(W, bias, gamma, beta, (__nv_bfloat16*)Wc, u, v, K, n_off);
    return check_launch("tcn_prep_fold");
}

int fqss_tcn_prep_batch(const fqss_prep_item* items, int n, void* stream) {
    FQSS_REQUIRE(items && n > 0, -1, "tcn_prep_batch: empty batch");
    cudaStream_t s = (cudaStream_t)stream;
    for (int i0 = 0; i0 < n; i0 += PREP_BATCH) {
        PrepBatch b;
        const int m = n - i0 < PREP_BATCH ? n - i0 : PREP_BATCH;
        int maxn = 0;
        for (int i = 0; i < m; ++i) {
            const fqss_prep_item& t = items[i0 + i];
            FQSS_REQUIRE(t.W && t.Wc && t.s1 && t.s0 && t.dws && t.N > 0 && t.K > 0 && t.n_off >= 0 && t.n_off + t.N <= t.Ntot, -1,
                         "tcn_prep_batch: bad item %d", i0 + i);
            FQSS_REQUIRE(t.WcT || t.split, -1, "tcn_prep_batch: WcT may be NULL only for split (inference) operands");
            FQSS_REQUIRE(!t.split || (t.wmin == nullptr && t.amin == nullptr), -1, "tcn_prep_batch: split operands are for the float model");
            FQSS_REQUIRE((t.wmin == nullptr) == (t.wmax == nullptr) && (t.amin == nullptr) == (t.amax == nullptr), -1,
                         "tcn_prep_batch: ranges come in pairs");
            b.it[i] = t;
            if (t.N > maxn) maxn = t.N;
        }
        FQSS_PROF("tcn_prep(batch)", s);
        tcn_prep_batch_kernel<<<dim3(maxn, m), 128, 0, s>>>(b);
    }
    return check_launch("tcn_prep_batch");
}

int fqss_tcn_encode(const float* x, int64_t ldx, void* out_bf16, int64_t ldo, int64_t rows, int M, const float* rmin,
                    const float* rmax, void* stream) {
    FQSS_REQUIRE(x && out_bf16 && rows > 0 && M > 0 && ldx >= M && ldo >= M, -1, "tcn_encode: bad argument");
    FQSS_PROF("tcn_encode", stream);
    tcn_encode_kernel<<<(unsigned)rows, ROW_THREADS, 0, (cudaStream_t)stream>>>(x, ldx, (__nv_bfloat16*)out_bf16, ldo, M, rmin, rmax);
    return check_launch("tcn_encode");
}

int fqss_split_bf16(const float* x, int64_t ldx, void* out_bf16, int64_t ldo, int64_t rows, int cols, int C, int layout, void* stream) {
    FQSS_REQUIRE(x && out_bf16 && rows > 0 && cols > 0 && ldx >= cols && (layout == 0 || (layout == 1 && C > 0 && rows % C == 0 && ldo >= cols)),
                 -1, "split_bf16: bad argument");
    FQSS_PROF("split_bf16", stream);
    split_bf16_kernel<<<(unsigned)rows, ROW_THREADS, 0, (cudaStream_t)stream>>>(x, ldx, (__nv_bfloat16*)out_bf16, ldo, cols, C, layout);
    return check_launch("split_bf16");
}

int fqss_rowscale_bf16(const float* g, int64_t ldg, void* out_bf16, int64_t ldo, int64_t rows, int M, int C, const float* scale,
                       double* rowsum, void* stream) {
    FQSS_REQUIRE(g && out_bf16 && scale && rows > 0 && M > 0 && C > 0 && rows % C == 0, -1, "rowscale_bf16: bad argument");
    FQSS_REQUIRE(ldg >= M && ldo >= ((M + 3) & ~3) && ldg % 4 == 0 && ldo % 4 == 0 && aligned16(g) && aligned16(out_bf16), -2,
                 "rowscale_bf16: rows must be 16-byte aligned with room for whole quads");
    cudaStream_t s = (cudaStream_t)stream;
    FQSS_PROFN("rowscale_bf16", s, 1);
    if (rowsum) cudaMemsetAsync(rowsum, 0, (size_t)C * sizeof(double), s);
    rowscale_bf16_kernel<<<(unsigned)rows, ROW_THREADS, 0, s>>>(g, ldg, (__nv_bfloat16*)out_bf16, ldo, M, C, scale, rowsum);
    return check_launch("rowscale_bf16");
}

int fqss_tcn_block_fwd(const fqss_tcn_block* p, void* stream) {
    int rc = validate_block(p, "tcn_block_fwd");
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int rows = p->B * p->Chid;
    cudaMemsetAsync(p->stats1, 0, ((size_t)p->B * 2 + 1) * sizeof(double), s);     // sums + the arrival counter of the finaliser
    cudaMemsetAsync(p->stats3, 0, ((size_t)p->B * 2 + 1) * sizeof(double), s);
    // K1
    tcg::Args a{};
    a.B = p->B; a.M = p->M; a.K = p->Cio; a.N = p->Chid; a.ld = p->ld; a.s1 = p->s1_1; a.s0 = p->s0_1; a.quant = p->quant;
    if (p->split) { a.K = 3 * p->Cio; a.a_rows = 2 * p->Cio; }
    a.out_f32 = p->y1; a.code1 = p->quant ? p->code1 : nullptr; a.slope = p->slope1; a.q1_min = p->q1.rmin; a.q1_max = p->q1.rmax; a.stats = p->stats1;
    a.rc = p->rc1; a.n_elems = (double)p->Chid * (double)p->M;
    a.q2_min = p->q2.rmin; a.q2_max = p->q2.rmax; a.q3_min = p->q3.rmin; a.q3_max = p->q3.rmax;
    rc = tcg::run(tcg::EPI_EXPAND, p->x_op, p->Wc1, a, s);
    if (rc) return rc;
    // K2 / K3a
    {
        const int dpad = dw_pad(p->dil);
        const size_t smem = 2048 + ((size_t)p->ld + 2 * dpad) * sizeof(float);      // slack to a 1 KB boundary + table + row with halo
        FQSS_REQUIRE(smem <= 200 * 1024, -1, "tcn_block_fwd: row + dilation halo do not fit shared memory (M=%d, dil=%d)", p->M, p->dil);
        static bool cfg = false;
        if (!cfg) {
#define FQSS_DW_ATTR(Q, D, T) cudaFuncSetAttribute(tcn_dw_fwd_kernel<Q, D, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)
            FQSS_DW_ATTR(true, 0, 256); FQSS_DW_ATTR(true, 1, 256); FQSS_DW_ATTR(true, 2, 256); FQSS_DW_ATTR(true, 3, 256);
            FQSS_DW_ATTR(true, 0, 128); FQSS_DW_ATTR(true, 1, 128); FQSS_DW_ATTR(true, 2, 128); FQSS_DW_ATTR(true, 3, 128);
            FQSS_DW_ATTR(false, 0, 128); FQSS_DW_ATTR(false, 1, 128); FQSS_DW_ATTR(false, 2, 128); FQSS_DW_ATTR(false, 3, 128);
            FQSS_DW_ATTR(false, 0, 256); FQSS_DW_ATTR(false, 1, 256); FQSS_DW_ATTR(false, 2, 256); FQSS_DW_ATTR(false, 3, 256);
#undef FQSS_DW_ATTR
            cfg = true;
        }
        // CTA width: 128 threads for the speech rows (M = 3 999: measured 80 vs 87 us quantised, 98 vs 132 us float), 256 for
        // long rows (music, M = 7 999: 49.3 vs 51.8 us, 57.4 vs 59.5 us); FQSS_DWF_TH / FQSS_DWFF_TH override
        static const int env_q = getenv("FQSS_DWF_TH") ? atoi(getenv("FQSS_DWF_TH")) : 0;
        static const int env_f = getenv("FQSS_DWFF_TH") ? atoi(getenv("FQSS_DWFF_TH")) : 0;
        const int th_q = env_q ? env_q : (p->M > 6000 ? 256 : 128);
        const int th_f = env_f ? env_f : (p->M > 6000 ? 256 : 128);
#define FQSS_DW_LAUNCH(Q, D)                                                                                   \
    do {                                                                                                       \
        FQSS_PROF(Q ? "tcn_dw_fwd" : "tcn_dw_fwd(float)", s);                                                  \
        if ((Q ? th_q : th_f) == 256) tcn_dw_fwd_kernel<Q, D, 256><<<rows, 256, smem, s>>>(*p);                \
        else tcn_dw_fwd_kernel<Q, D, 128><<<rows, 128, smem, s>>>(*p);                                         \
    } while (0)
        RowConstJob job3;
        job3.stats = p->stats3; job3.rc = p->rc3; job3.B = p->B; job3.n_elems = (double)p->Chid * (double)p->M;
        job3.qa_min = p->q3.rmin; job3.qa_max = p->q3.rmax; job3.qb_min = p->q4.rmin; job3.qb_max = p->q4.rmax;
        job3.qc_min = nullptr; job3.qc_max = nullptr;
#define FQSS_RC3_LAUNCH() do { FQSS_PROF("tcn_rowconst", s); tcn_rowconst_kernel<<<1, 64, 0, s>>>(job3); } while (0)
        const int mode = dw_mode(p->dil);
        if (p->quant) {
            if (mode == 0) FQSS_DW_LAUNCH(true, 0); else if (mode == 1) FQSS_DW_LAUNCH(true, 1);
            else if (mode == 2) FQSS_DW_LAUNCH(true, 2); else FQSS_DW_LAUNCH(true, 3);
            FQSS_RC3_LAUNCH();
            {
                FQSS_PROF("tcn_hidden_fq", s);
                static const int hth = getenv("FQSS_HFQ_TH") ? atoi(getenv("FQSS_HFQ_TH")) : 128;
                if (hth == 256) tcn_hidden_fq_kernel<true, 256><<<rows, 256, 0, s>>>(*p);
                else if (hth == 64) tcn_hidden_fq_kernel<true, 64><<<rows, 64, 0, s>>>(*p);
                else tcn_hidden_fq_kernel<true, 128><<<rows, 128, 0, s>>>(*p);
            }
        } else {
            if (mode == 0) FQSS_DW_LAUNCH(false, 0); else if (mode == 1) FQSS_DW_LAUNCH(false, 1);
            else if (mode == 2) FQSS_DW_LAUNCH(false, 2); else FQSS_DW_LAUNCH(false, 3);
            if (p->split != 2) {
                FQSS_RC3_LAUNCH();
                { FQSS_PROF("tcn_hidden_fq(float)", s); tcn_hidden_fq_kernel<false, 256><<<rows, 256, 0, s>>>(*p); }
            }
        }
#undef FQSS_DW_LAUNCH
#undef FQSS_RC3_LAUNCH
    }
    rc = check_launch("tcn_block_fwd(K2/K3a)");
    if (rc) return rc;
    // K3
    tcg::Args k{};
    k.B = p->B; k.M = p->M; k.K = p->Chid; k.N = (p->has_res ? p->Cio : 0) + (p->no_skip ? 0 : p->Cio); k.ld = p->ld;
    if (p->split) { k.K = 3 * p->Chid; k.a_rows = 2 * p->Chid; k.split = 1; }
    if (p->split == 2) { k.fold_stats = p->stats3; k.n_elems = (double)p->Chid * (double)p->M; }
    k.s1 = p->s1_2; k.s0 = p->s0_2; k.quant = p->quant;
    k.n_res = p->has_res ? p->Cio : 0; k.first_block = p->first_block;
    k.res_y = p->res_y; k.skip_y = p->skip_y; k.x_in = p->x_in; k.x_out = p->x_out; k.x_out_op = (__nv_bfloat16*)p->x_out_op;
    k.skip_in = p->skip_in; k.skip_out = p->skip_out;
    k.qres_min = p->qres.rmin; k.qres_max = p->qres.rmax; k.qskip_min = p->qskip.rmin; k.qskip_max = p->qskip.rmax;
    k.qadd_min = p->qadd.rmin; k.qadd_max = p->qadd.rmax; k.qadds_min = p->qadds.rmin; k.qadds_max = p->qadds.rmax;
    return tcg::run(tcg::EPI_RESSKIP, p->a4_op, p->Wc2, k, s);
}

}  // extern "C"
