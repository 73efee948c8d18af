// tcn_bwd.cu -- backward of the fused TCN ConvBlock.  Every quantised activation is re-derived from what forward
// saved (8-bit codes code1 / code3 where the code is enough, the pre-activations y1, y3, res_y, skip_y where the
// continuous value matters) with the same device functions forward used, so the straight-through masks match the
// forward codes exactly.  Stages (one launch each):
//
//   T    tail          g_x_out, g_skip_out -> dY2 (bf16, res|skip rows, pre-scaled by delta_w), g_xd, g_skip_in
//   G    dgrad GEMM    dY2 x Wc2T -> g_a4 (bf16)                                    [tcgen05]
//   W    wgrad GEMM    dY2 x a4_op -> dW2q            [tcgen05, split-K; side stream, next to P1]
//   P1   gLN2 sums     g_a4, code3 -> per-row {sum g_n3, sum g_n3*xhat3}, range sums of FQ4
//   R    reduce        per-sample S1,S2 and per-channel dgamma,dbeta
//   P2D  fused         g_a4, y3 -> gLN2/FQ3/PReLU3 backward -> g_y3 row in SHARED memory -> depthwise + FQ2 backward
//                      (taps from shared memory, code1) -> g_n1 (bf16), dW_dw, db_dw, gLN1 row sums, range sums
//                      (the two-kernel variant P2 + D is kept behind FQSS_SPLIT_P2D for A/B runs)
//   R    reduce
//   Q    gLN1/FQ1/PReLU backward  g_n1, y1 -> dY1 (bf16, pre-scaled), db1
//   G    dgrad GEMM    dY1 x Wc1T (+ g_xd) -> g_x_in (fp32)                          [tcgen05]
//   W    wgrad GEMM    dY1 x x_op -> dW1q                                           [tcgen05, split-K]
//   F    finalise      fp64 accumulators -> fp32 parameter gradients
//
// The 512-wide gradient tensors travel in bf16 (the "1e-2 bf16 GEMM path"); the residual-stream and
// skip-sum gradients (128-wide) stay fp32 end to end.
#include <stdlib.h>
#include <type_traits>

#include "fqss_common.cuh"
#include "gemm_tc.cuh"
#include "tcn_common.cuh"
#include "tcn_bwd_common.cuh"
#include "wgrad_tc.cuh"

namespace fqss {

// ---------------------------------------------------------------------------------------------
// T: tail backward over the 128-wide tensors.  grid = B*Cio rows.
// ---------------------------------------------------------------------------------------------
struct TailQ {
    ActQF qres, qskip, qadd, qadds;
};

// HAS_SKIP = false: skip-less block (fqss_tcn_block.no_skip) -- the residual branch alone moves half the bytes per trip, so
// NQ = 2 frame quads are requested before the first is consumed (bytes in flight per SM decide: 80 -> 56 us at M = 7 999)
template <bool HAS_SKIP, int NQ>
__global__ void __launch_bounds__(ROW_THREADS) tcn_tail_bwd_kernel(const fqss_tcn_block p, const fqss_tcn_block_grads g, double* acc) {
    __shared__ double sh[10 * 32];
    __shared__ TailQ tq;
    const AccLayout L(p.B, p.Cio, p.Chid);
    const int64_t r = blockIdx.x;
    const int b = (int)(r / p.Cio), o = (int)(r % p.Cio);
    const int M = p.M;
    constexpr bool has_skip = HAS_SKIP;
    const int n2 = (p.has_res ? p.Cio : 0) + (has_skip ? p.Cio : 0);
    if (p.quant && threadIdx.x == 0) {
        if (has_skip) tq.qskip = load_actqf(p.qskip.rmin, p.qskip.rmax, 8);
        if (p.has_res) {
            tq.qres = load_actqf(p.qres.rmin, p.qres.rmax, 8);
            tq.qadd = load_actqf(p.qadd.rmin, p.qadd.rmax, 8);
        }
        if (!p.first_block && has_skip) tq.qadds = load_actqf(p.qadds.rmin, p.qadds.rmax, 8);
    }
    __syncthreads();
    const ActQF qres = tq.qres, qskip = tq.qskip, qadd = tq.qadd, qadds = tq.qadds;
    const float sc_res = p.has_res ? __ldg(p.dws2 + o) : 0.f;
    const float sc_skip = has_skip ? __ldg(p.dws2 + (p.has_res ? p.Cio : 0) + o) : 0.f;
    __nv_bfloat16* dY = reinterpret_cast<__nv_bfloat16*>(g.dY2);
    __nv_bfloat16* dres = dY + ((int64_t)b * n2 + o) * p.ld;
    __nv_bfloat16* dskip = dY + ((int64_t)b * n2 + (p.has_res ? p.Cio : 0) + o) * p.ld;
    float s[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) s[i] = 0.f;      // 0,1 qadd | 2,3 qres | 4,5 qadds | 6,7 qskip | 8 db_res | 9 db_skip
    const int nvec = (int)(p.ld >> 2);
    const int64_t rb = r * p.ld;
    struct TailLd { float4 gxo, ry, xin, gso, sy, sin; };
    // issue every load of a trip before the first use
    auto load = [&](int v) {
        TailLd t;
        const int64_t i = rb + 4 * v;
        if (p.has_res) {
            t.gxo = ld_f4(g.g_x_out + i);
            t.ry = ldg4_stream(p.res_y + i);
            if (p.quant) t.xin = ldg4_stream(p.x_in + i);
        }
        if (has_skip) {
            t.gso = ld_f4(g.g_skip_out + i);
            t.sy = ldg4_stream(p.skip_y + i);
            if (p.quant && !p.first_block) t.sin = ldg4_stream(p.skip_in + i);
        }
        return t;
    };
    auto body = [&](int v, const TailLd& t) {
        const int m0 = 4 * v;
        const int64_t i = rb + m0;
        const float4 gxo = t.gxo, ry = t.ry, xin = t.xin, gso = t.gso, sy = t.sy, sin = t.sin;
        if (p.has_res) {
            float gz[4], gr[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float gk = (m0 + k < M) ? f4_get(gxo, k) : 0.f;
                if (p.quant) {
                    const float ryk = f4_get(ry, k);
                    const float rq = actqf_fq(qres, ryk);
                    const float z = __fadd_rn(f4_get(xin, k), rq);
                    gz[k] = actqf_bwd(qadd, z, gk, s[0], s[1]);
                    gr[k] = actqf_bwd(qres, ryk, gz[k], s[2], s[3]);
                } else {
                    gz[k] = gk;
                    gr[k] = gk;
                }
                s[8] += gr[k];
            }
            stg4(g.g_xd + i, make_float4(gz[0], gz[1], gz[2], gz[3]));
            st_bf16x4(dres + m0, gr[0] * sc_res, gr[1] * sc_res, gr[2] * sc_res, gr[3] * sc_res);
        }
        if (has_skip) {
            float gz[4], gs[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float gk = (m0 + k < M) ? f4_get(gso, k) : 0.f;
                if (p.quant) {
                    const float syk = f4_get(sy, k);
                    gz[k] = gk;
                    if (!p.first_block) {
                        const float sq = actqf_fq(qskip, syk);
                        const float z = __fadd_rn(f4_get(sin, k), sq);
                        gz[k] = actqf_bwd(qadds, z, gk, s[4], s[5]);
                    }
                    gs[k] = actqf_bwd(qskip, syk, gz[k], s[6], s[7]);
                } else {
                    gz[k] = gk;
                    gs[k] = gk;
                }
                s[9] += gs[k];
            }
            if (!p.first_block) stg4(g.g_skip_in + i, make_float4(gz[0], gz[1], gz[2], gz[3]));
            st_bf16x4(dskip + m0, gs[0] * sc_skip, gs[1] * sc_skip, gs[2] * sc_skip, gs[3] * sc_skip);
        }
    };
    for (int base = threadIdx.x; base < nvec; base += NQ * ROW_THREADS) {
        TailLd d[NQ];
#pragma unroll
        for (int k = 0; k < NQ; ++k)
            if (base + k * ROW_THREADS < nvec) d[k] = load(base + k * ROW_THREADS);
#pragma unroll
        for (int k = 0; k < NQ; ++k)
            if (base + k * ROW_THREADS < nvec) body(base + k * ROW_THREADS, d[k]);
    }
    double v[10];
    block_sum_fd<10>(s, v, sh);
    if (threadIdx.x == 0) {
        if (p.quant) {
            if (p.has_res) {
                atomicAdd(acc + L.qs(o) + 2 * QADD, v[0]); atomicAdd(acc + L.qs(o) + 2 * QADD + 1, v[1]);
                atomicAdd(acc + L.qs(o) + 2 * QRES, v[2]); atomicAdd(acc + L.qs(o) + 2 * QRES + 1, v[3]);
            }
            if (!p.first_block && has_skip) { atomicAdd(acc + L.qs(o) + 2 * QADDS, v[4]); atomicAdd(acc + L.qs(o) + 2 * QADDS + 1, v[5]); }
            if (has_skip) { atomicAdd(acc + L.qs(o) + 2 * QSKIP, v[6]); atomicAdd(acc + L.qs(o) + 2 * QSKIP + 1, v[7]); }
        }
        if (p.has_res) atomicAdd(acc + L.db2 + o, v[8]);
        if (has_skip) atomicAdd(acc + L.db2 + (p.has_res ? p.Cio : 0) + o, v[9]);
    }
}

// ---------------------------------------------------------------------------------------------
// P1 / P2: gLN2 + FQ4 (+ FQ3 + PReLU3 in P2).  grid = B*Chid rows.  g_a4 in g_hid_a (bf16).
// Everything downstream of FQ3 is a function of the 8-bit code of a3: tabX = xhat3 | mask4, tabD = D4.
// ---------------------------------------------------------------------------------------------
template <int PHASE, bool QUANT, int NTH, int NQ>
__global__ void __launch_bounds__(NTH) tcn_gln2_bwd_kernel(const fqss_tcn_block p, const fqss_tcn_block_grads g, double* acc) {
    __shared__ double sh[4 * 32];
    __shared__ float tabX[256], tabD[256];
    const AccLayout L(p.B, p.Cio, p.Chid);
    const int64_t r = blockIdx.x;
    const int b = (int)(r / p.Chid), c = (int)(r % p.Chid);
    const Hidden3 h = load_hidden3(p, b, c);
    if (QUANT) {
        for (int i = threadIdx.x; i < 256; i += NTH) {
            chain_bwd_tables(h.q3, h.g, h.q4, i, tabX, tabD);
            if (PHASE == 1)         // P1 reads only tabD: fold the mask into its LSB
                tabD[i] = __uint_as_float((__float_as_uint(tabD[i]) & ~1u) | (tab_mask(tabX[i]) ? 1u : 0u));
        }
        __syncthreads();
    }
    const float2 xa3 = f2s(QUANT ? h.q3.delta * h.g.rstd : 0.f), xb3 = f2s(QUANT ? (h.q3.mn - h.g.mu) * h.g.rstd : 0.f);
    const int M = p.M;
    const float4* y3 = reinterpret_cast<const float4*>(p.y3 + r * p.ld);
    const uint2* ga4 = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(g.g_hid_a) + r * p.ld);
    uint2* gy3 = reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.g_hid_b) + r * p.ld);
    float2 A = f2s(0.f), nBc = f2s(0.f), nCc = f2s(0.f);
    if (PHASE == 2) {
        const float invN = __fdividef(1.f, (float)p.Chid * (float)p.M);
        A = f2s(h.g.rstd * h.g.gamma);
        double sm1, sm2;
        samp_get(acc, L, 2, b, sm1, sm2);
        nBc = f2s(-h.g.rstd * invN * (float)sm1);
        nCc = f2s(-h.g.rstd * invN * (float)sm2);
    }
    const float slope = h.slope;
    // P1: a0 = sum g*D4, a1 = sum g*(1-m4), a2 = sum g*m4, a3 = sum g*m4*xhat     (q4: sD = a0, sZ = a1)
    // P2: a0 = sum ga3*D3, a1 = sum ga3*(1-m3), a3 = sum min(y,0)*gz              (q3: sD = a0, sZ = a1; slope3)
    float2 a0 = f2s(0.f), a1 = f2s(0.f), a2 = f2s(0.f), a3 = f2s(0.f);
    constexpr bool CODES = (PHASE == 1 && QUANT);      // everything P1 needs is a function of the saved code of a3
    const uint32_t* c3 = reinterpret_cast<const uint32_t*>(p.code3 + r * p.ld);
    struct Ld { float4 y; uint32_t packed; uint2 g; };
    auto load = [&](int v) {
        Ld d;
        d.y = make_float4(0.f, 0.f, 0.f, 0.f);
        d.packed = 0;
        if (CODES) d.packed = __ldg(c3 + v);
        else d.y = __ldg(y3 + v);
        d.g = __ldg(ga4 + v);
        return d;
    };
    auto body = [&](int v, const Ld& d, auto tail_tag) {
        constexpr bool TAIL = decltype(tail_tag)::value;
        const float4 y = d.y;
        const uint32_t packed = d.packed;
        const float4 gi = bf16x4_to_float4(d.g);
        float2 gg[2] = {lo2(gi), hi2(gi)};
        if (TAIL) mask_tail(gg[0], gg[1], M - 4 * v);
        const float2 yy[2] = {lo2(y), hi2(y)};
        float2 o[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float2 z = f2s(0.f);
            if (!CODES) z = make_float2(prelu_f(yy[j].x, slope), prelu_f(yy[j].y, slope));
            float2 xh, gn, t = f2s(0.f);
            unsigned ix = 0, iy = 0;
            if (CODES) {
                // one LDS.32 per element (D4 with the STE mask in its LSB); xhat3 is one FMA on the code: shared-memory
                // bandwidth under random bank conflicts, not issue slots, is what the table lookups cost
                ix = (packed >> (16 * j)) & 255u;
                iy = (packed >> (16 * j + 8)) & 255u;
                const float2 dm = make_float2(tabD[ix], tabD[iy]);
                xh = __ffma2_rn(make_float2((float)ix, (float)iy), xa3, xb3);
                gn = make_float2(tab_mask(dm.x) ? gg[j].x : 0.f, tab_mask(dm.y) ? gg[j].y : 0.f);
                a0 = __ffma2_rn(gg[j], dm, a0);
                a1 = __fadd2_rn(a1, __fadd2_rn(gg[j], neg2(gn)));
            } else if (QUANT) {
                t = actqf_t2(h.q3, z);
                ix = code_u8(t.x);
                iy = code_u8(t.y);
                xh = make_float2(tabX[ix], tabX[iy]);
                gn = make_float2(tab_mask(xh.x) ? gg[j].x : 0.f, tab_mask(xh.y) ? gg[j].y : 0.f);
            } else {
                xh = make_float2(gln_xhat(h.g, z.x), gln_xhat(h.g, z.y));
                gn = gg[j];
            }
            if (PHASE == 1) {
                if (QUANT && !CODES) {
                    a0 = __ffma2_rn(gg[j], make_float2(tabD[ix], tabD[iy]), a0);
                    a1 = __fadd2_rn(a1, __fadd2_rn(gg[j], neg2(gn)));      // g*(1-m): exactly g or 0, no cancellation
                }
                a2 = __fadd2_rn(a2, gn);
                a3 = __ffma2_rn(gn, xh, a3);
            } else {
                float2 ga3 = __ffma2_rn(A, gn, __ffma2_rn(xh, nCc, nBc));
                if (TAIL) {
                    if (4 * v + 2 * j + 1 >= M) ga3.y = 0.f;
                    if (4 * v + 2 * j >= M) ga3.x = 0.f;
                }
                float2 gz = ga3;
                if (QUANT) {
                    const float2 cf = make_float2((float)ix, (float)iy);
                    const float2 dd = __fadd2_rn(cf, neg2(t));
                    const float2 th = __fadd2_rn(t, f2s(0.5f));
                    const bool inx = inside_u8(th.x), iny = inside_u8(th.y);
                    a0 = __ffma2_rn(ga3, make_float2(inx ? dd.x : cf.x, iny ? dd.y : cf.y), a0);
                    gz = make_float2(inx ? ga3.x : 0.f, iny ? ga3.y : 0.f);
                    a1 = __fadd2_rn(a1, __fadd2_rn(ga3, neg2(gz)));
                }
                o[j] = __fmul2_rn(gz, make_float2(yy[j].x > 0.f ? 1.f : slope, yy[j].y > 0.f ? 1.f : slope));
                a3 = __ffma2_rn(make_float2(fminf(yy[j].x, 0.f), fminf(yy[j].y, 0.f)), gz, a3);
            }
        }
        if (PHASE == 2) gy3[v] = float4_to_bf16x4(o[0].x, o[0].y, o[1].x, o[1].y);
    };
    FQSS_ROW_LOOP_BATCH(NTH, NQ, load, body, M);
    const float s[4] = {hsum(a0), hsum(a1), PHASE == 1 ? hsum(a2) : hsum(a3), hsum(a3)};
    double v[4];
    block_sum_fd<4>(s, v, sh);
    if (threadIdx.x == 0) {
        if (PHASE == 1) {
            if (QUANT) { atomicAdd(acc + L.qs(c) + 2 * Q4, v[0]); atomicAdd(acc + L.qs(c) + 2 * Q4 + 1, v[1]); }
            acc[L.row2 + 2 * r] = v[2];
            acc[L.row2 + 2 * r + 1] = v[3];
        } else {
            if (QUANT) { atomicAdd(acc + L.qs(c) + 2 * Q3, v[0]); atomicAdd(acc + L.qs(c) + 2 * Q3 + 1, v[1]); }
            atomicAdd(acc + L.qs(c) + AccLayout::SLOPE_OFF + 1, v[2]);
        }
    }
}

// P1 of the quantised model on the saved codes: 3 B/element of HBM traffic, so the kernel lives or dies by the number
// of loads in flight.  Every thread first issues the loads of NQ quads (code word + bf16 gradient quad each), then
// consumes them; quads at or beyond ceil(M/4) are predicated off, the ragged quad masks its gradient lanes.
template <int NQ, int NTH>
__global__ void __launch_bounds__(NTH) tcn_gln2_sums_codes_kernel(const fqss_tcn_block p, const fqss_tcn_block_grads g, double* acc) {
    __shared__ double sh[4 * 32];
    __shared__ float tabD[256];
    const AccLayout L(p.B, p.Cio, p.Chid);
    const int64_t r = blockIdx.x;
    const int b = (int)(r / p.Chid), c = (int)(r % p.Chid);
    const int M = p.M;
    const int nq = (M + 3) >> 2;
    const uint32_t* c3 = reinterpret_cast<const uint32_t*>(p.code3 + r * p.ld);
    const uint2* ga4 = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(g.g_hid_a) + r * p.ld);
    uint32_t cw[NQ];
    uint2 gw[NQ];
    auto issue = [&](int base) {
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
            const int v = base + k * NTH;
            const bool ok = v < nq;
            cw[k] = ok ? __ldg(c3 + v) : 0u;
            gw[k] = ok ? __ldg(ga4 + v) : make_uint2(0u, 0u);
        }
    };
    // the first batch of row data is requested BEFORE the per-row constants and the table are built: one CTA handles one
    // short row, so serialising "constants -> table -> barrier -> data" would expose two full memory latencies per CTA
    issue(threadIdx.x);
    const Hidden3 h = load_hidden3(p, b, c);
    for (int i = threadIdx.x; i < 256; i += NTH) {
        const float4 e = chain_bwd_entry(h.q3, h.g, h.q4, i, false);        // {-, mask4, D4, xhat3}
        tabD[i] = __uint_as_float((__float_as_uint(e.z) & ~1u) | (e.y != 0.f ? 1u : 0u));
    }
    __syncthreads();
    const float2 xa3 = f2s(h.q3.delta * h.g.rstd), xb3 = f2s((h.q3.mn - h.g.mu) * h.g.rstd);
    // a0 = sum g*D4, a1 = sum g*(1-m4), a2 = sum g*m4, a3 = sum g*m4*xhat3
    float2 a0 = f2s(0.f), a1 = f2s(0.f), a2 = f2s(0.f), a3 = f2s(0.f);
    for (int base = threadIdx.x; base < nq; base += NQ * NTH) {
        if (base != (int)threadIdx.x) issue(base);
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
            const int v = base + k * NTH;
            const float4 gi = bf16x4_to_float4(gw[k]);
            float2 gg[2] = {lo2(gi), hi2(gi)};
            if (4 * v + 3 >= M) mask_tail(gg[0], gg[1], M - 4 * v);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const unsigned ix = (cw[k] >> (16 * j)) & 255u, iy = (cw[k] >> (16 * j + 8)) & 255u;
                const float2 dm = make_float2(tabD[ix], tabD[iy]);
                const float2 xh = __ffma2_rn(make_float2((float)ix, (float)iy), xa3, xb3);
                const float2 gn = make_float2(tab_mask(dm.x) ? gg[j].x : 0.f, tab_mask(dm.y) ? gg[j].y : 0.f);
                a0 = __ffma2_rn(gg[j], dm, a0);
                a1 = __fadd2_rn(a1, __fadd2_rn(gg[j], neg2(gn)));
                a2 = __fadd2_rn(a2, gn);
                a3 = __ffma2_rn(gn, xh, a3);
            }
        }
    }
    const float s[4] = {hsum(a0), hsum(a1), hsum(a2), hsum(a3)};
    double v[4];
    block_sum_fd<4>(s, v, sh);
    if (threadIdx.x == 0) {
        atomicAdd(acc + L.qs(c) + 2 * Q4, v[0]);
        atomicAdd(acc + L.qs(c) + 2 * Q4 + 1, v[1]);
        acc[L.row2 + 2 * r] = v[2];
        acc[L.row2 + 2 * r + 1] = v[3];
    }
}

// R: per-sample {S1,S2} and per-channel dgamma/dbeta from the per-row sums.  CTAs [0,B): one sample each;
// CTAs [B, B + ceil(C/32)): 32 channels each, 8 thread groups split the samples (the loads of one thread are independent,
// so the whole reduction is one memory round trip deep instead of B).
__global__ void __launch_bounds__(256) tcn_gln_reduce_kernel(const double* __restrict__ rowacc, int B, int C, const float* __restrict__ gamma,
                                                            float* __restrict__ g_gamma, float* __restrict__ g_beta, double* __restrict__ samp) {
    __shared__ double sh[2 * 256];
    if ((int)blockIdx.x < B) {
        const int b = blockIdx.x;
        double s1 = 0.0, s2 = 0.0;
        for (int c = threadIdx.x; c < C; c += 256) {
            const double gm = (double)__ldg(gamma + c);
            const double2 v = *reinterpret_cast<const double2*>(rowacc + 2 * ((int64_t)b * C + c));
            s1 += gm * v.x;
            s2 += gm * v.y;
        }
        double v[2] = {s1, s2};
        block_sum<2>(v, sh);
        if (threadIdx.x == 0) { samp[2 * b] = v[0]; samp[2 * b + 1] = v[1]; }
    } else {
        const int c = (blockIdx.x - B) * 32 + (threadIdx.x & 31), grp = threadIdx.x >> 5;
        double gb = 0.0, gg = 0.0;
        if (c < C) {
#pragma unroll 4
            for (int b = grp; b < B; b += 8) {
                const double2 v = *reinterpret_cast<const double2*>(rowacc + 2 * ((int64_t)b * C + c));
                gb += v.x;
                gg += v.y;
            }
        }
        sh[threadIdx.x] = gb;
        sh[256 + threadIdx.x] = gg;
        __syncthreads();
        if (grp == 0 && c < C) {
#pragma unroll
            for (int k = 1; k < 8; ++k) { gb += sh[32 * k + threadIdx.x]; gg += sh[256 + 32 * k + threadIdx.x]; }
            g_beta[c] = (float)gb;
            g_gamma[c] = (float)gg;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// D: depthwise backward + FQ2 backward + gLN1 row sums.  One CTA per row, streaming (no staging, no barrier in
// the element loop):
//   g_a2[m] = w0 g[m+d] + w1 g[m] + w2 g[m-d]   -- the three taps of g_y3 (bf16) are read straight from global memory;
//             the row is 8 KB, so the two shifted reads hit L1/L2 and DRAM still sees each byte once
//   FQ2: t2 = pre-rounding value of FQ2 as a function of the saved code of a1 (one 32-bit table lookup); code2, a2,
//             the STE mask and the range weight follow arithmetically (bit-identical to forward); xhat1 is one FMA
//   taps:     dW_k = sum_m g[m] a2[m+(k-1)d] = sum_m a2[m] g[m-(k-1)d]  (g is zero outside [0,M)), so only the
//             centre a2 is needed
// The float model derives a2 / xhat1 from y1 directly.
// ---------------------------------------------------------------------------------------------
template <bool QUANT, int DMODE, int NTH, int NQ>
__global__ void __launch_bounds__(NTH) tcn_dw_bwd_kernel(const fqss_tcn_block p, const fqss_tcn_block_grads g, double* acc) {
    __shared__ double sh[8 * 32];
    __shared__ float tabT[256];                       // QUANT: t2 = (gLN1(decode1(code1)) - min2) / delta2
    const AccLayout L(p.B, p.Cio, p.Chid);
    const int64_t r = blockIdx.x;
    const int b = (int)(r / p.Chid), c = (int)(r % p.Chid);
    const int M = p.M, d = p.dil;
    const Hidden1 h = load_hidden1(p, b, c);
    if (QUANT) {
        for (int i = threadIdx.x; i < 256; i += NTH) tabT[i] = actqf_t(h.q2, gln_apply(h.g, actqf_decode(h.q1, (float)i)));
        __syncthreads();
    }
    const float4* y1 = reinterpret_cast<const float4*>(p.y1 + r * p.ld);
    const uint32_t* c1 = reinterpret_cast<const uint32_t*>(p.code1 + r * p.ld);
    const __nv_bfloat16* gy3 = reinterpret_cast<const __nv_bfloat16*>(g.g_hid_b) + r * p.ld;
    const uint2* gq = reinterpret_cast<const uint2*>(gy3);
    const int nq = (M + 3) >> 2;                      // quads holding valid frames; the ragged one carries zeros beyond M
    const float slope = h.slope;
    const float2 w0 = f2s(__ldg(p.wdw + c * 3)), w1 = f2s(__ldg(p.wdw + c * 3 + 1)), w2 = f2s(__ldg(p.wdw + c * 3 + 2));
    const float2 xa1 = f2s(QUANT ? h.q1.delta * h.g.rstd : 0.f), xb1 = f2s(QUANT ? (h.q1.mn - h.g.mu) * h.g.rstd : 0.f);
    uint2* gn1o = reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.g_hid_a) + r * p.ld);
    auto ld1 = [&](int m) -> float { return (m >= 0 && m < M) ? __bfloat162float(gy3[m]) : 0.f; };
    // b0 = sum ga2*D2, b1 = sum ga2*(1-m2), b2 = sum gn1, b3 = sum gn1*xhat1 | taps: d0,d1,d2 = sum a2*g[+d,0,-d], d3 = sum g
    float2 b0 = f2s(0.f), b1 = f2s(0.f), b2 = f2s(0.f), b3 = f2s(0.f), d0 = f2s(0.f), d1 = f2s(0.f), d2 = f2s(0.f), d3 = f2s(0.f);
    auto ldraw = [&](int vq) -> uint2 {
        if (vq < 0 || vq >= nq) return make_uint2(0u, 0u);
        return __ldg(gq + vq);
    };
    struct Ld { float4 y; uint32_t packed; uint2 c, a, e; };      // a / e: the quads holding the left / right taps
    auto load = [&](int v) {
        Ld t;
        t.y = make_float4(0.f, 0.f, 0.f, 0.f);
        t.packed = 0;
        if (QUANT) t.packed = __ldg(c1 + v);
        else t.y = __ldg(y1 + v);
        t.c = ldraw(v);
        const int sh = DMODE == 0 ? (d >> 2) : 1;
        t.a = make_uint2(0u, 0u);
        t.e = make_uint2(0u, 0u);
        if (DMODE != 3) {
            t.a = ldraw(v - sh);
            t.e = ldraw(v + sh);
        }
        return t;
    };
    auto body = [&](int v, const Ld& t, auto tail_tag) {
        constexpr bool TAIL = decltype(tail_tag)::value;
        const uint32_t packed = t.packed;
        const float4 y = t.y;
        const float4 gC = bf16x4_to_float4(t.c);
        float4 gL, gR;
        if (DMODE == 0) {                      // d % 4 == 0: whole quads
            gL = bf16x4_to_float4(t.a);
            gR = bf16x4_to_float4(t.e);
        } else if (DMODE == 1) {
            const float4 a = bf16x4_to_float4(t.a), e = bf16x4_to_float4(t.e);
            gL = make_float4(a.w, gC.x, gC.y, gC.z);
            gR = make_float4(gC.y, gC.z, gC.w, e.x);
        } else if (DMODE == 2) {
            const float4 a = bf16x4_to_float4(t.a), e = bf16x4_to_float4(t.e);
            gL = make_float4(a.z, a.w, gC.x, gC.y);
            gR = make_float4(gC.z, gC.w, e.x, e.y);
        } else {
            const int m = 4 * v;
            gL = make_float4(ld1(m - d), ld1(m + 1 - d), ld1(m + 2 - d), ld1(m + 3 - d));
            gR = make_float4(ld1(m + d), ld1(m + 1 + d), ld1(m + 2 + d), ld1(m + 3 + d));
        }
        float2 o[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float2 gc = j ? hi2(gC) : lo2(gC), gl = j ? hi2(gL) : lo2(gL), gr = j ? hi2(gR) : lo2(gR);
            // y3[m'] = sum_k w_k a2[m' + (k-1)d]  =>  d/da2[m] = w0 g[m+d] + w1 g[m] + w2 g[m-d]
            float2 ga2 = __ffma2_rn(w0, gr, __ffma2_rn(w1, gc, __fmul2_rn(w2, gl)));
            float2 a2, xh, gn1;
            if (QUANT) {
                const unsigned ix = (packed >> (16 * j)) & 255u, iy = (packed >> (16 * j + 8)) & 255u;
                const float2 t2 = make_float2(tabT[ix], tabT[iy]);
                const float2 c2 = make_float2((float)code_u8(t2.x), (float)code_u8(t2.y));
                a2 = __fadd2_rn(__fmul2_rn(f2s(h.q2.delta), c2), f2s(h.q2.mn));
                xh = __ffma2_rn(make_float2((float)ix, (float)iy), xa1, xb1);
                if (TAIL) {                                   // frames >= M: no gradient, and a2 there is not part of the row
                    if (4 * v + 2 * j + 1 >= M) { ga2.y = 0.f; a2.y = 0.f; }
                    if (4 * v + 2 * j >= M) { ga2.x = 0.f; a2.x = 0.f; }
                }
                const float2 th = __fadd2_rn(t2, f2s(0.5f));
                const bool inx = inside_u8(th.x), iny = inside_u8(th.y);
                const float2 dd = __fadd2_rn(c2, neg2(t2));
                gn1 = make_float2(inx ? ga2.x : 0.f, iny ? ga2.y : 0.f);
                b0 = __ffma2_rn(ga2, make_float2(inx ? dd.x : c2.x, iny ? dd.y : c2.y), b0);
                b1 = __fadd2_rn(b1, __fadd2_rn(ga2, neg2(gn1)));
            } else {
                const float2 yj = j ? hi2(y) : lo2(y);
                const float2 z = make_float2(prelu_f(yj.x, slope), prelu_f(yj.y, slope));
                a2 = make_float2(gln_apply(h.g, z.x), gln_apply(h.g, z.y));
                xh = make_float2(gln_xhat(h.g, z.x), gln_xhat(h.g, z.y));
                if (TAIL) {
                    if (4 * v + 2 * j + 1 >= M) { ga2.y = 0.f; a2.y = 0.f; }
                    if (4 * v + 2 * j >= M) { ga2.x = 0.f; a2.x = 0.f; }
                }
                gn1 = ga2;
            }
            d0 = __ffma2_rn(a2, gr, d0);      // dW_0 = sum_m a2[m] g[m+d]
            d1 = __ffma2_rn(a2, gc, d1);
            d2 = __ffma2_rn(a2, gl, d2);      // dW_2 = sum_m a2[m] g[m-d]
            d3 = __fadd2_rn(d3, gc);
            b2 = __fadd2_rn(b2, gn1);
            b3 = __ffma2_rn(gn1, xh, b3);
            o[j] = gn1;
        }
        gn1o[v] = float4_to_bf16x4(o[0].x, o[0].y, o[1].x, o[1].y);
    };
    FQSS_ROW_LOOP_BATCH(NTH, NQ, load, body, M);
    const float s[8] = {hsum(b0), hsum(b1), hsum(b2), hsum(b3), hsum(d0), hsum(d1), hsum(d2), hsum(d3)};
    double v[8];
    block_sum_fd<8>(s, v, sh);
    if (threadIdx.x == 0) {
        if (QUANT) { atomicAdd(acc + L.qs(c) + 2 * Q2, v[0]); atomicAdd(acc + L.qs(c) + 2 * Q2 + 1, v[1]); }
        acc[L.row1 + 2 * r] = v[2];
        acc[L.row1 + 2 * r + 1] = v[3];
        atomicAdd(acc + L.dwdw + 3 * c, v[4]);
        atomicAdd(acc + L.dwdw + 3 * c + 1, v[5]);
        atomicAdd(acc + L.dwdw + 3 * c + 2, v[6]);
        atomicAdd(acc + L.dbdw + c, v[7]);
    }
}

// ---------------------------------------------------------------------------------------------
// P2 + D fused: both stages are local to one (sample, channel) row once the per-sample gLN2 sums exist, so one CTA
//   phase A  g_a4 (bf16), y3 -> gLN2 / FQ3 / PReLU3 backward -> g_y3 row in SHARED memory (fp32, zero halo of `dil`)
//   phase B  taps of g_y3 from shared memory + saved code of a1 -> depthwise / FQ2 backward -> g_n1 (bf16), sums
// g_y3 never goes to HBM (was: 2 B/frame written + read back three times through L1), the bf16 rounding of g_y3
// disappears, and the two block reductions become one.  HBM bytes: y3 4 + g_a4 2 + code1 1 + g_n1 2 = 9 B/frame.
// ---------------------------------------------------------------------------------------------
template <bool QUANT, int DMODE, int NTH, int NQ>
__global__ void __launch_bounds__(NTH) tcn_gln2_dw_bwd_kernel(const fqss_tcn_block p, const fqss_tcn_block_grads g, double* acc) {
    extern __shared__ __align__(16) float dsm[];      // [dpad | ld | dpad]
    __shared__ double sh[11 * 32];
    __shared__ float tabX[256], tabT[256];
    const AccLayout L(p.B, p.Cio, p.Chid);
    const int64_t r = blockIdx.x;
    const int b = (int)(r / p.Chid), c = (int)(r % p.Chid);
    const int M = p.M, d = p.dil, dpad = dw_pad(d), ld = (int)p.ld;
    float* row = dsm + dpad;
    const Hidden3 h3 = load_hidden3(p, b, c);
    const Hidden1 h1 = load_hidden1(p, b, c);
    // zero the halo and every frame from the ragged quad's end up to the pitch
    for (int i = threadIdx.x; i < dpad; i += NTH) {
        dsm[i] = 0.f;
        row[ld + i] = 0.f;
    }
    for (int i = ((M + 3) & ~3) + threadIdx.x; i < ld; i += NTH) row[i] = 0.f;
    if (QUANT) {
        for (int i = threadIdx.x; i < 256; i += NTH) {
            const float4 e = chain_bwd_entry(h3.q3, h3.g, h3.q4, i, false);      // {-, mask4, D4, xhat3}
            tabX[i] = __uint_as_float((__float_as_uint(e.w) & ~1u) | (e.y != 0.f ? 1u : 0u));
            tabT[i] = actqf_t(h1.q2, gln_apply(h1.g, actqf_decode(h1.q1, (float)i)));
        }
    }
    __syncthreads();
    // ---------------- phase A: gLN2 + FQ3 + PReLU3 backward -> row ----------------
    float2 a0 = f2s(0.f), a1 = f2s(0.f), a3 = f2s(0.f);      // q3: sD, sZ; slope3
    {
        const float4* y3 = reinterpret_cast<const float4*>(p.y3 + r * p.ld);
        const uint2* ga4 = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(g.g_hid_a) + r * p.ld);
        const float invN = __fdividef(1.f, (float)p.Chid * (float)p.M);
        const float2 A = f2s(h3.g.rstd * h3.g.gamma);
        double sm1, sm2;
        samp_get(acc, L, 2, b, sm1, sm2);
        const float2 nBc = f2s(-h3.g.rstd * invN * (float)sm1);
        const float2 nCc = f2s(-h3.g.rstd * invN * (float)sm2);
        const float slope = h3.slope;
        struct Ld { float4 y; uint2 g; };
        auto load = [&](int v) {
            Ld t;
            t.y = __ldg(y3 + v);
            t.g = __ldg(ga4 + v);
            return t;
        };
        auto body = [&](int v, const Ld& t, auto tail_tag) {
            constexpr bool TAIL = decltype(tail_tag)::value;
            const float4 gi = bf16x4_to_float4(t.g);
            const float2 gg[2] = {lo2(gi), hi2(gi)};
            const float2 yy[2] = {lo2(t.y), hi2(t.y)};
            float2 o[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float2 z = make_float2(prelu_f(yy[j].x, slope), prelu_f(yy[j].y, slope));
                float2 xh, gn, tq = f2s(0.f);
                unsigned ix = 0, iy = 0;
                if (QUANT) {
                    tq = actqf_t2(h3.q3, z);
                    ix = code_u8(tq.x);
                    iy = code_u8(tq.y);
                    xh = make_float2(tabX[ix], tabX[iy]);
                    gn = make_float2(tab_mask(xh.x) ? gg[j].x : 0.f, tab_mask(xh.y) ? gg[j].y : 0.f);
                } else {
                    xh = make_float2(gln_xhat(h3.g, z.x), gln_xhat(h3.g, z.y));
                    gn = gg[j];
                }
                float2 ga3 = __ffma2_rn(A, gn, __ffma2_rn(xh, nCc, nBc));
                if (TAIL) {
                    if (4 * v + 2 * j + 1 >= M) ga3.y = 0.f;
                    if (4 * v + 2 * j >= M) ga3.x = 0.f;
                }
                float2 gz = ga3;
                if (QUANT) {
                    const float2 cf = make_float2((float)ix, (float)iy);
                    const float2 dd = __fadd2_rn(cf, neg2(tq));
                    const float2 th = __fadd2_rn(tq, f2s(0.5f));
                    const bool inx = inside_u8(th.x), iny = inside_u8(th.y);
                    a0 = __ffma2_rn(ga3, make_float2(inx ? dd.x : cf.x, iny ? dd.y : cf.y), a0);
                    gz = make_float2(inx ? ga3.x : 0.f, iny ? ga3.y : 0.f);
                    a1 = __fadd2_rn(a1, __fadd2_rn(ga3, neg2(gz)));
                }
                o[j] = __fmul2_rn(gz, make_float2(yy[j].x > 0.f ? 1.f : slope, yy[j].y > 0.f ? 1.f : slope));
                a3 = __ffma2_rn(make_float2(fminf(yy[j].x, 0.f), fminf(yy[j].y, 0.f)), gz, a3);
            }
            *reinterpret_cast<float4*>(row + 4 * v) = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);
        };
        FQSS_ROW_LOOP_BATCH(NTH, NQ, load, body, M);
    }
    __syncthreads();
    // ---------------- phase B: depthwise + FQ2 backward, gLN1 row sums ----------------
    const float4* y1 = reinterpret_cast<const float4*>(p.y1 + r * p.ld);
    const uint32_t* c1 = reinterpret_cast<const uint32_t*>(p.code1 + r * p.ld);
    const float slope1 = h1.slope;
    const float2 w0 = f2s(__ldg(p.wdw + c * 3)), w1 = f2s(__ldg(p.wdw + c * 3 + 1)), w2 = f2s(__ldg(p.wdw + c * 3 + 2));
    const float2 xa1 = f2s(QUANT ? h1.q1.delta * h1.g.rstd : 0.f), xb1 = f2s(QUANT ? (h1.q1.mn - h1.g.mu) * h1.g.rstd : 0.f);
    uint2* gn1o = reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.g_hid_a) + r * p.ld);
    // b0 = sum ga2*D2, b1 = sum ga2*(1-m2), b2 = sum gn1, b3 = sum gn1*xhat1 | taps: d0,d1,d2 = sum a2*g[+d,0,-d], d3 = sum g
    float2 b0 = f2s(0.f), b1 = f2s(0.f), b2 = f2s(0.f), b3 = f2s(0.f), d0 = f2s(0.f), d1 = f2s(0.f), d2 = f2s(0.f), d3 = f2s(0.f);
    {
        struct Ld { float4 y; uint32_t packed; };
        auto load = [&](int v) {
            Ld t;
            t.y = make_float4(0.f, 0.f, 0.f, 0.f);
            t.packed = 0;
            if (QUANT) t.packed = __ldg(c1 + v);
            else t.y = __ldg(y1 + v);
            return t;
        };
        auto body = [&](int v, const Ld& t, auto tail_tag) {
            constexpr bool TAIL = decltype(tail_tag)::value;
            const uint32_t packed = t.packed;
            float4 gL, gC, gR;
            dw_taps<DMODE>(row, v, d, gL, gC, gR);
            float2 o[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float2 gc = j ? hi2(gC) : lo2(gC), gl = j ? hi2(gL) : lo2(gL), gr = j ? hi2(gR) : lo2(gR);
                // y3[m'] = sum_k w_k a2[m' + (k-1)d]  =>  d/da2[m] = w0 g[m+d] + w1 g[m] + w2 g[m-d]
                float2 ga2 = __ffma2_rn(w0, gr, __ffma2_rn(w1, gc, __fmul2_rn(w2, gl)));
                float2 a2, xh, gn1;
                if (QUANT) {
                    const unsigned ix = (packed >> (16 * j)) & 255u, iy = (packed >> (16 * j + 8)) & 255u;
                    const float2 t2 = make_float2(tabT[ix], tabT[iy]);
                    const float2 c2 = make_float2((float)code_u8(t2.x), (float)code_u8(t2.y));
                    a2 = __ffma2_rn(f2s(h1.q2.delta), c2, f2s(h1.q2.mn));      // feeds only the tap-gradient sums
                    xh = __ffma2_rn(make_float2((float)ix, (float)iy), xa1, xb1);
                    if (TAIL) {                                   // frames >= M: no gradient, and a2 there is not part of the row
                        if (4 * v + 2 * j + 1 >= M) { ga2.y = 0.f; a2.y = 0.f; }
                        if (4 * v + 2 * j >= M) { ga2.x = 0.f; a2.x = 0.f; }
                    }
                    const float2 th = __fadd2_rn(t2, f2s(0.5f));
                    const bool inx = inside_u8(th.x), iny = inside_u8(th.y);
                    const float2 dd = __fadd2_rn(c2, neg2(t2));
                    gn1 = make_float2(inx ? ga2.x : 0.f, iny ? ga2.y : 0.f);
                    b0 = __ffma2_rn(ga2, make_float2(inx ? dd.x : c2.x, iny ? dd.y : c2.y), b0);
                    b1 = __fadd2_rn(b1, __fadd2_rn(ga2, neg2(gn1)));
                } else {
                    const float2 yj = j ? hi2(t.y) : lo2(t.y);
                    const float2 z = make_float2(prelu_f(yj.x, slope1), prelu_f(yj.y, slope1));
                    a2 = make_float2(gln_apply(h1.g, z.x), gln_apply(h1.g, z.y));
                    xh = make_float2(gln_xhat(h1.g, z.x), gln_xhat(h1.g, z.y));
                    if (TAIL) {
                        if (4 * v + 2 * j + 1 >= M) { ga2.y = 0.f; a2.y = 0.f; }
                        if (4 * v + 2 * j >= M) { ga2.x = 0.f; a2.x = 0.f; }
                    }
                    gn1 = ga2;
                }
                d0 = __ffma2_rn(a2, gr, d0);      // dW_0 = sum_m a2[m] g[m+d]
                d1 = __ffma2_rn(a2, gc, d1);
                d2 = __ffma2_rn(a2, gl, d2);      // dW_2 = sum_m a2[m] g[m-d]
                d3 = __fadd2_rn(d3, gc);
                b2 = __fadd2_rn(b2, gn1);
                b3 = __ffma2_rn(gn1, xh, b3);
                o[j] = gn1;
            }
            gn1o[v] = float4_to_bf16x4(o[0].x, o[0].y, o[1].x, o[1].y);
        };
        FQSS_ROW_LOOP_BATCH(NTH, NQ, load, body, M);
    }
    const float s[11] = {hsum(a0), hsum(a1), hsum(a3), hsum(b0), hsum(b1), hsum(b2), hsum(b3), hsum(d0), hsum(d1), hsum(d2), hsum(d3)};
    double v[11];
    block_sum_fd<11>(s, v, sh);
    if (threadIdx.x == 0) {
        if (QUANT) {
            atomicAdd(acc + L.qs(c) + 2 * Q3, v[0]); atomicAdd(acc + L.qs(c) + 2 * Q3 + 1, v[1]);
            atomicAdd(acc + L.qs(c) + 2 * Q2, v[3]); atomicAdd(acc + L.qs(c) + 2 * Q2 + 1, v[4]);
        }
        atomicAdd(acc + L.qs(c) + AccLayout::SLOPE_OFF + 1, v[2]);
        acc[L.row1 + 2 * r] = v[5];
        acc[L.row1 + 2 * r + 1] = v[6];
        atomicAdd(acc + L.dwdw + 3 * c, v[7]);
        atomicAdd(acc + L.dwdw + 3 * c + 1, v[8]);
        atomicAdd(acc + L.dwdw + 3 * c + 2, v[9]);
        atomicAdd(acc + L.dbdw + c, v[10]);
    }
}


}  // namespace fqss

#include "tcn_rows.cuh"

namespace fqss {

// ---------------------------------------------------------------------------------------------
// Q: gLN1 + FQ1 + PReLU1 backward: g_n1 (bf16, g_hid_a), y1 -> dY1 (bf16, pre-scaled by delta_w1), db1
// ---------------------------------------------------------------------------------------------
template <bool QUANT, int NTH, int NQ>
__global__ void __launch_bounds__(NTH) tcn_gln1_bwd_kernel(const fqss_tcn_block p, const fqss_tcn_block_grads g, double* acc) {
    __shared__ double sh[4 * 32];
    const AccLayout L(p.B, p.Cio, p.Chid);
    const int64_t r = blockIdx.x;
    const int b = (int)(r / p.Chid), c = (int)(r % p.Chid);
    const Hidden1 h = load_hidden1(p, b, c);
    const int M = p.M;
    const float4* y1 = reinterpret_cast<const float4*>(p.y1 + r * p.ld);
    const uint2* gn1 = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(g.g_hid_a) + r * p.ld);
    uint2* dY1 = reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.dY1) + r * p.ld);
    const float invN = __fdividef(1.f, (float)p.Chid * (float)p.M);
    const float2 A2 = f2s(h.g.rstd * h.g.gamma);
    double sm1, sm2;
    samp_get(acc, L, 1, b, sm1, sm2);
    const float2 nBc = f2s(-h.g.rstd * invN * (float)sm1);
    const float2 nCc = f2s(-h.g.rstd * invN * (float)sm2);
    // xhat1 = (a1 - mu) * rstd with a1 = delta1 * code + min1  ->  one FMA on the code
    const float2 xa = f2s(QUANT ? h.q1.delta * h.g.rstd : 0.f);
    const float2 xb = f2s(QUANT ? (h.q1.mn - h.g.mu) * h.g.rstd : 0.f);
    const float2 sc = f2s(__ldg(p.dws1 + c));
    const float slope = h.slope;
    // a0 = sum ga1*D1, a1 = sum ga1*(1-m1), a3 = sum min(y,0)*gz, a4 = sum gy
    float2 a0 = f2s(0.f), a1 = f2s(0.f), a3 = f2s(0.f), a4 = f2s(0.f);
    struct Ld { float4 y; uint2 g; };
    auto load = [&](int v) {
        Ld d;
        d.y = __ldg(y1 + v);
        d.g = __ldg(gn1 + v);
        return d;
    };
    auto body = [&](int v, const Ld& d, auto tail_tag) {
        constexpr bool TAIL = decltype(tail_tag)::value;
        const float4 y = d.y;
        const float4 gi = bf16x4_to_float4(d.g);
        const float2 yy[2] = {lo2(y), hi2(y)};
        const float2 gg[2] = {lo2(gi), hi2(gi)};
        float2 o[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float2 z = make_float2(prelu_f(yy[j].x, slope), prelu_f(yy[j].y, slope));
            float2 gz;
            if (QUANT) {
                const float2 t = actqf_t2(h.q1, z);
                const float2 cf = make_float2((float)code_u8(t.x), (float)code_u8(t.y));
                const float2 xh = __ffma2_rn(cf, xa, xb);
                float2 ga1 = __ffma2_rn(A2, gg[j], __ffma2_rn(xh, nCc, nBc));
                if (TAIL) {
                    if (4 * v + 2 * j + 1 >= M) ga1.y = 0.f;
                    if (4 * v + 2 * j >= M) ga1.x = 0.f;
                }
                const float2 dd = __fadd2_rn(cf, neg2(t));
                const float2 th = __fadd2_rn(t, f2s(0.5f));
                    const bool inx = inside_u8(th.x), iny = inside_u8(th.y);
                a0 = __ffma2_rn(ga1, make_float2(inx ? dd.x : cf.x, iny ? dd.y : cf.y), a0);
                gz = make_float2(inx ? ga1.x : 0.f, iny ? ga1.y : 0.f);
                a1 = __fadd2_rn(a1, __fadd2_rn(ga1, neg2(gz)));
            } else {
                const float2 xh = make_float2(gln_xhat(h.g, z.x), gln_xhat(h.g, z.y));
                gz = __ffma2_rn(A2, gg[j], __ffma2_rn(xh, nCc, nBc));
                if (TAIL) {
                    if (4 * v + 2 * j + 1 >= M) gz.y = 0.f;
                    if (4 * v + 2 * j >= M) gz.x = 0.f;
                }
            }
            const float2 gy = __fmul2_rn(gz, make_float2(yy[j].x > 0.f ? 1.f : slope, yy[j].y > 0.f ? 1.f : slope));
            a3 = __ffma2_rn(make_float2(fminf(yy[j].x, 0.f), fminf(yy[j].y, 0.f)), gz, a3);
            a4 = __fadd2_rn(a4, gy);
            o[j] = __fmul2_rn(gy, sc);
        }
        dY1[v] = float4_to_bf16x4(o[0].x, o[0].y, o[1].x, o[1].y);
    };
    FQSS_ROW_LOOP_BATCH(NTH, NQ, load, body, M);
    const float s[4] = {hsum(a0), hsum(a1), hsum(a3), hsum(a4)};
    double v[4];
    block_sum_fd<4>(s, v, sh);
    if (threadIdx.x == 0) {
        if (QUANT) { atomicAdd(acc + L.qs(c) + 2 * Q1, v[0]); atomicAdd(acc + L.qs(c) + 2 * Q1 + 1, v[1]); }
        atomicAdd(acc + L.qs(c) + AccLayout::SLOPE_OFF, v[2]);
        atomicAdd(acc + L.db1 + c, v[3]);
    }
}

// F: fp64 accumulators -> fp32 outputs
__global__ void tcn_bwd_finalize_kernel(const fqss_tcn_block p, const fqss_tcn_block_grads g, const double* __restrict__ acc, int gln_from_acc) {
    const AccLayout L(p.B, p.Cio, p.Chid);
    const int n2 = (p.has_res ? p.Cio : 0) + (p.no_skip ? 0 : p.Cio);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (gln_from_acc && i < p.Chid) {      // the lean row kernels accumulate dgamma / dbeta here (no reduce launch)
        g.g_gn1_b[i] = (float)acc[L.gln1 + 2 * i];
        g.g_gn1_w[i] = (float)acc[L.gln1 + 2 * i + 1];
        g.g_gn2_b[i] = (float)acc[L.gln2 + 2 * i];
        g.g_gn2_w[i] = (float)acc[L.gln2 + 2 * i + 1];
    }
    if (i < 8 && p.quant) {
        double2 v[AccLayout::NSLOT];          // all slot loads in flight at once (one L2 round trip, not NSLOT)
#pragma unroll
        for (int k = 0; k < AccLayout::NSLOT; ++k) v[k] = *reinterpret_cast<const double2*>(acc + L.qs(k) + 2 * i);
        double sD = 0.0, sZ = 0.0;
#pragma unroll
        for (int k = 0; k < AccLayout::NSLOT; ++k) {
            sD += v[k].x;
            sZ += v[k].y;
        }
        g.g_q[2 * i] = (float)(sZ - sD / 255.0);      // d/d min_range
        g.g_q[2 * i + 1] = (float)(sD / 255.0);       // d/d max_range
    }
    if (i == 8 || i == 9) {
        double w[AccLayout::NSLOT];
#pragma unroll
        for (int k = 0; k < AccLayout::NSLOT; ++k) w[k] = acc[L.qs(k) + AccLayout::SLOPE_OFF + (i - 8)];
        double sl = 0.0;
#pragma unroll
        for (int k = 0; k < AccLayout::NSLOT; ++k) sl += w[k];
        (i == 8 ? g.g_slope1 : g.g_slope3)[0] = (float)sl;
    }
    if (i < p.Chid) {
        g.db1[i] = (float)acc[L.db1 + i];
        g.dbdw[i] = (float)acc[L.dbdw + i];
        for (int k = 0; k < 3; ++k) g.dwdw[3 * i + k] = (float)acc[L.dwdw + 3 * i + k];
    }
    if (i < n2) g.db2[i] = (float)acc[L.db2 + i];
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int g_wgrad_overlap = -1;      // -1: take FQSS_WGRAD_OVERLAP from the environment (default on) at first use

// side stream + fork / join events of the backward pass (created once per process; one process drives one GPU)
static int g_side_device = -1;      // the device the side stream / events were created on
static cudaStream_t side_stream() {
    static cudaStream_t st = nullptr;
    if (!st) {
        cudaGetDevice(&g_side_device);
        cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    }
    return st;
}
// The statics above belong to ONE device (the launch contract is one process per GPU): a call from another device would
// record events on a foreign stream -- refuse it loudly instead.
static bool side_device_ok() {
    if (g_side_device < 0) return true;
    int dev = -1;
    cudaGetDevice(&dev);
    return dev == g_side_device;
}
static cudaEvent_t side_event(int i) {      // 0 / 1: fork / done of the dW2 wgrad; 2: fork of the block tail; 3, 4: tail done (per accumulator)
    static cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (!ev[i]) cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
    return ev[i];
}

// "Tail on the side stream" (fqss_set_bwd_tail_side): the second weight-gradient GEMM of a block and the finalise kernel
// leave the main stream as well, so the next block's backward starts right after this block's last dgrad GEMM.  Nothing on
// the chain consumes their outputs (parameter gradients, read after the whole stack: fqss_tcn_bwd_join orders that); the
// reuse hazards of the shared scratch are ordered by events: dY2 (next block's tail kernel vs this block's dW2 wgrad),
// dY1 (next block's gLN1 kernel vs this block's dW1 wgrad), and the fp64 accumulator block, which is double-buffered.
static int g_tail_side = 0;
static int g_flip = 0;
static bool g_pend_w2 = false, g_pend_tail[2] = {false, false};

// development knob: loads-in-flight batch of a row kernel, from the environment (read once per name)
static int tune_nq(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

}  // namespace fqss

using namespace fqss;

extern "C" {

int fqss_set_wgrad_overlap(int on) {
    const int prev = g_wgrad_overlap;
    g_wgrad_overlap = on;
    return prev;
}

int fqss_set_bwd_tail_side(int on) {
    const int prev = g_tail_side;
    g_tail_side = on;
    return prev;
}

// Orders everything the backward calls since the last join left on the side stream before later work on `stream`.
int fqss_tcn_bwd_join(void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (g_pend_w2) cudaStreamWaitEvent(s, side_event(1), 0);
    for (int i = 0; i < 2; ++i)
        if (g_pend_tail[i]) cudaStreamWaitEvent(s, side_event(3 + i), 0);
    g_pend_w2 = g_pend_tail[0] = g_pend_tail[1] = false;
    return check_launch("tcn_bwd_join");
}

size_t fqss_tcn_ws_bytes(int B, int Cio, int Chid) {
    AccLayout L(B, Cio, Chid);
    size_t acc = 2 * align_up((size_t)L.total * sizeof(double), 256);      // two accumulator blocks (alternating per call)
    size_t cst = align_up((size_t)2 * 1024 * sizeof(float), 256);
    // the partial buffer is sized for the longest sequence the row kernels accept (M <= 16K); the actual need
    // depends only on the split count, which is capped by the SM count
    size_t p1 = (size_t)160 * Chid * Cio * sizeof(float);
    size_t p2 = (size_t)160 * 2 * Cio * Chid * sizeof(float);
    return acc + cst + (p1 > p2 ? p1 : p2) + 1024;
}

int fqss_tcn_block_bwd(const fqss_tcn_block* p, const fqss_tcn_block_grads* g, void* stream) {
    int rc = tcn_validate_block(p, "tcn_block_bwd");
    if (rc) return rc;
    FQSS_REQUIRE(!p->quant || (p->code1 && p->code3), -1, "tcn_block_bwd: forward did not save the activation codes (code1 / code3)");
    FQSS_REQUIRE(!p->split && p->y1 && p->y3 && (p->skip_y || p->no_skip) && (!p->has_res || p->res_y) && p->Wc1T && p->Wc2T, -1,
                 "tcn_block_bwd: block was run in inference mode (split operands / no saved pre-activations)");
    FQSS_REQUIRE(g && (g->g_skip_out || p->no_skip) && g->g_x_in && g->dY2 && g->g_hid_a && g->dY1 && g->ws, -1, "tcn_block_bwd: null buffer");
    FQSS_REQUIRE(!p->has_res || (g->g_x_out && g->g_xd), -1, "tcn_block_bwd: residual path needs g_x_out / g_xd");
    FQSS_REQUIRE(p->first_block || p->no_skip || g->g_skip_in, -1, "tcn_block_bwd: g_skip_in missing");
    FQSS_REQUIRE(g->dW1q && g->db1 && g->dW2q && g->db2 && g->dwdw && g->dbdw && g->g_gn1_w && g->g_gn1_b && g->g_gn2_w && g->g_gn2_b &&
                     g->g_slope1 && g->g_slope3 && g->g_q, -1, "tcn_block_bwd: null parameter-gradient output");
    FQSS_REQUIRE(g->ws_bytes >= fqss_tcn_ws_bytes(p->B, p->Cio, p->Chid), -3, "tcn_block_bwd: workspace too small");
    FQSS_REQUIRE(p->Chid <= 1024 && 2 * p->Cio <= 1024, -1, "tcn_block_bwd: channel count too large");
    FQSS_REQUIRE(side_device_ok(), -1, "tcn_block_bwd: the library's side stream belongs to device %d (one process drives one GPU)", g_side_device);
    cudaStream_t s = (cudaStream_t)stream;
    const AccLayout L(p->B, p->Cio, p->Chid);
    const size_t acc_bytes = align_up((size_t)L.total * sizeof(double), 256);
    const int overlap = g_wgrad_overlap < 0 ? (g_wgrad_overlap = tune_nq("FQSS_WGRAD_OVERLAP", 1)) : g_wgrad_overlap;
    const bool tail_side = overlap && g_tail_side;
    int flip = 0;
    if (tail_side) {
        flip = g_flip;
        g_flip ^= 1;
        if (g_pend_w2) cudaStreamWaitEvent(s, side_event(1), 0);               // previous block's dW2 wgrad still reads dY2
        if (g_pend_tail[flip]) cudaStreamWaitEvent(s, side_event(3 + flip), 0);  // this accumulator's previous finalise
    }
    double* acc = (double*)((char*)g->ws + (size_t)flip * acc_bytes);
    float* part = (float*)((char*)g->ws + 2 * acc_bytes + align_up((size_t)2 * 1024 * sizeof(float), 256));
    const size_t part_cap = g->ws_bytes - ((char*)part - (char*)g->ws);
    const int n2 = (p->has_res ? p->Cio : 0) + (p->no_skip ? 0 : p->Cio);
    const int rows_h = p->B * p->Chid, rows_io = p->B * p->Cio;

    cudaMemsetAsync(acc, 0, (size_t)L.total * sizeof(double), s);
    // T
    {
        FQSS_PROF("tcn_tail_bwd", s);
        if (p->no_skip) tcn_tail_bwd_kernel<false, 2><<<rows_io, ROW_THREADS, 0, s>>>(*p, *g, acc);
        else tcn_tail_bwd_kernel<true, 1><<<rows_io, ROW_THREADS, 0, s>>>(*p, *g, acc);
    }
    rc = check_launch("tcn_block_bwd(tail)");
    if (rc) return rc;
    // G: g_a4 = Wc2T-GEMM(dY2)   (K = n2, N = Chid) -> bf16
    {
        tcg::Args a{};
        a.B = p->B; a.M = p->M; a.K = n2; a.N = p->Chid; a.ld = p->ld; a.s1 = nullptr; a.s0 = nullptr; a.out_bf16 = (__nv_bfloat16*)g->g_hid_a;
        rc = tcg::run(tcg::EPI_BF16, g->dY2, p->Wc2T, a, s);
        if (rc) return rc;
    }
    // W: dW2q.  The split-K wgrad (one 200 KB CTA per SM, TMA + tcgen05, SM pipes nearly idle, ~47 % of HBM) and the
    // gLN2 row-sum kernel that follows (ALU-bound, ~40 % of HBM) do not depend on each other and fit on an SM together:
    // fork the wgrad to a side stream (inputs dY2 / db2 sums are complete after T), join before the second wgrad, which
    // reuses the partial-tile buffer.  Works the same under stream capture (fork / join through events).
    cudaStream_t sw = s;
    if (overlap) {
        sw = side_stream();
        cudaEventRecord(side_event(0), s);
        cudaStreamWaitEvent(sw, side_event(0), 0);
    }
    rc = tcw::run(g->dY2, p->a4_op, p->B, p->M, p->ld, n2, p->Chid, part, part_cap, p->quant ? p->q4.rmin : nullptr,
                  p->quant ? p->q4.rmax : nullptr, p->dws2, acc + L.db2, g->dW2q, sw);
    if (overlap) cudaEventRecord(side_event(1), sw);      // recorded even on failure: the main stream must not wait forever
    if (tail_side) g_pend_w2 = true;
    if (rc) return rc;
    // P1, R, P2
    static const int split_p2d = tune_nq("FQSS_SPLIT_P2D", 0);
    static const int lean_env = tune_nq("FQSS_LEAN_P2D", 1);
    // quantised model: the instruction-lean row kernels (tcn_rows.cuh); they add their gLN sums straight into the
    // accumulator block, so the two reduce launches disappear.  FQSS_LEAN_P2D=0 keeps the first versions for A/B runs.
    const bool lean = p->quant && lean_env && !split_p2d;
    if (lean) {
        FQSS_PROF("tcn_gln2_sums", s);
        tcn_gln2_sums_lean_kernel<128, 8><<<rows_h, 128, P1_LEAN_SMEM, s>>>(*p, *g, acc);
    } else {
        {
            FQSS_PROF("tcn_gln2_bwd<1>", s);
            static const int p1_th = tune_nq("FQSS_P1_TH", 128);
            if (p->quant) {
                if (p1_th == 64) tcn_gln2_sums_codes_kernel<4, 64><<<rows_h, 64, 0, s>>>(*p, *g, acc);
                else tcn_gln2_sums_codes_kernel<4, 128><<<rows_h, 128, 0, s>>>(*p, *g, acc);
            }
            else tcn_gln2_bwd_kernel<1, false, 128, 2><<<rows_h, 128, 0, s>>>(*p, *g, acc);
        }
        { FQSS_PROF("tcn_gln_reduce", s); tcn_gln_reduce_kernel<<<p->B + (p->Chid + 31) / 32, 256, 0, s>>>(acc + L.row2, p->B, p->Chid, p->gn2_w, g->g_gn2_w, g->g_gn2_b,
                                                                         acc + L.samp2); }
    }
    // P2 + D (one kernel; FQSS_SPLIT_P2D=1 runs the two separate kernels instead -- development / A-B knob)
    if (lean) {
        const size_t smem = p2d_lean_smem(p->ld, p->dil);
        FQSS_REQUIRE(smem <= 200 * 1024, -1, "tcn_block_bwd: row + dilation halo do not fit shared memory (M=%d, dil=%d)", p->M, p->dil);
        const int mode = dw_mode(p->dil);
        FQSS_PROF("tcn_gln2_dw_bwd", s);
#define FQSS_FL_LAUNCH1(D, NQv, MB)                                                                                         \
    do {                                                                                                                   \
        if (smem > 48 * 1024)                                                                                              \
            cudaFuncSetAttribute(tcn_gln2_dw_bwd_lean_kernel<D, 128, NQv, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
        tcn_gln2_dw_bwd_lean_kernel<D, 128, NQv, MB><<<rows_h, 128, smem, s>>>(*p, *g, acc);                               \
    } while (0)
        // development knob FQSS_FL_VAR: quads in flight / CTAs per SM = 0: 4 / 7, 1: 4 / 6, 2: 2 / 7
        static const int fl_var = tune_nq("FQSS_FL_VAR", 0);
#define FQSS_FL_LAUNCH(D)                                                                  \
    do {                                                                                   \
        if (fl_var == 1) FQSS_FL_LAUNCH1(D, 4, 6); else if (fl_var == 2) FQSS_FL_LAUNCH1(D, 2, 7); \
        else FQSS_FL_LAUNCH1(D, 4, 7);                                                     \
    } while (0)
        if (mode == 0) FQSS_FL_LAUNCH(0); else if (mode == 1) FQSS_FL_LAUNCH(1); else if (mode == 2) FQSS_FL_LAUNCH(2); else FQSS_FL_LAUNCH(3);
#undef FQSS_FL_LAUNCH1
#undef FQSS_FL_LAUNCH
    } else if (!split_p2d) {
        const int dpad = dw_pad(p->dil);
        const size_t smem = ((size_t)p->ld + 2 * dpad) * sizeof(float);
        FQSS_REQUIRE(smem <= 200 * 1024, -1, "tcn_block_bwd: row + dilation halo do not fit shared memory (M=%d, dil=%d)", p->M, p->dil);
        const int mode = dw_mode(p->dil);
        static const int nqf = tune_nq("FQSS_NQ_F", 4);
        FQSS_PROF("tcn_gln2_dw_bwd", s);
        static const int f_th = tune_nq("FQSS_F_TH", 128);
#define FQSS_F_LAUNCH(Q, D, NQv)                                                                                         \
    do {                                                                                                                 \
        if (smem > 48 * 1024) {                                                                                          \
            cudaFuncSetAttribute(tcn_gln2_dw_bwd_kernel<Q, D, 128, NQv>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
            cudaFuncSetAttribute(tcn_gln2_dw_bwd_kernel<Q, D, 64, NQv>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);  \
        }                                                                                                                \
        if (f_th == 64) tcn_gln2_dw_bwd_kernel<Q, D, 64, NQv><<<rows_h, 64, smem, s>>>(*p, *g, acc);                      \
        else tcn_gln2_dw_bwd_kernel<Q, D, 128, NQv><<<rows_h, 128, smem, s>>>(*p, *g, acc);                               \
    } while (0)
#define FQSS_F_MODE(Q, NQv)                                                                          \
    do {                                                                                             \
        if (mode == 0) FQSS_F_LAUNCH(Q, 0, NQv); else if (mode == 1) FQSS_F_LAUNCH(Q, 1, NQv);       \
        else if (mode == 2) FQSS_F_LAUNCH(Q, 2, NQv); else FQSS_F_LAUNCH(Q, 3, NQv);                 \
    } while (0)
        if (p->quant) {
            if (nqf == 1) FQSS_F_MODE(true, 1); else if (nqf == 2) FQSS_F_MODE(true, 2); else FQSS_F_MODE(true, 4);
        } else {
            FQSS_F_MODE(false, 2);
        }
#undef FQSS_F_MODE
#undef FQSS_F_LAUNCH
    } else {
        FQSS_REQUIRE(g->g_hid_b, -1, "tcn_block_bwd: the two-kernel gLN2 / depthwise path needs the g_hid_b scratch");
        {
            FQSS_PROF("tcn_gln2_bwd<2>", s);
            if (p->quant) {
                const int nqv = tune_nq("FQSS_NQ_P2", 4);
                if (nqv == 1) tcn_gln2_bwd_kernel<2, true, 128, 1><<<rows_h, 128, 0, s>>>(*p, *g, acc);
                else if (nqv == 2) tcn_gln2_bwd_kernel<2, true, 128, 2><<<rows_h, 128, 0, s>>>(*p, *g, acc);
                else tcn_gln2_bwd_kernel<2, true, 128, 4><<<rows_h, 128, 0, s>>>(*p, *g, acc);
            } else tcn_gln2_bwd_kernel<2, false, 128, 2><<<rows_h, 128, 0, s>>>(*p, *g, acc);
        }
        // D, R, Q
        {
            FQSS_PROF("tcn_dw_bwd", s);
            const int nqd = tune_nq("FQSS_NQ_DW", 4);
#define FQSS_DWB_LAUNCH(Q, D)                                                                      \
        do {                                                                                           \
            if (nqd == 1) tcn_dw_bwd_kernel<Q, D, 128, 1><<<rows_h, 128, 0, s>>>(*p, *g, acc);          \
            else if (nqd == 2) tcn_dw_bwd_kernel<Q, D, 128, 2><<<rows_h, 128, 0, s>>>(*p, *g, acc);     \
            else tcn_dw_bwd_kernel<Q, D, 128, 4><<<rows_h, 128, 0, s>>>(*p, *g, acc);                   \
        } while (0)
            const int mode = dw_mode(p->dil);
            if (p->quant) {
                if (mode == 0) FQSS_DWB_LAUNCH(true, 0); else if (mode == 1) FQSS_DWB_LAUNCH(true, 1);
                else if (mode == 2) FQSS_DWB_LAUNCH(true, 2); else FQSS_DWB_LAUNCH(true, 3);
            } else {
                if (mode == 0) FQSS_DWB_LAUNCH(false, 0); else if (mode == 1) FQSS_DWB_LAUNCH(false, 1);
                else if (mode == 2) FQSS_DWB_LAUNCH(false, 2); else FQSS_DWB_LAUNCH(false, 3);
            }
#undef FQSS_DWB_LAUNCH
        }
    }
    if (!lean) {
        FQSS_PROF("tcn_gln_reduce", s);
        tcn_gln_reduce_kernel<<<p->B + (p->Chid + 31) / 32, 256, 0, s>>>(acc + L.row1, p->B, p->Chid, p->gn1_w, g->g_gn1_w, g->g_gn1_b, acc + L.samp1);
    }
    if (tail_side && g_pend_tail[flip ^ 1]) cudaStreamWaitEvent(s, side_event(3 + (flip ^ 1)), 0);   // previous block's dW1 wgrad still reads dY1
    {
        FQSS_PROF("tcn_gln1_bwd", s);
        if (p->quant) {
            const int nqv = tune_nq("FQSS_NQ_Q", 14);
            if (nqv == 1) tcn_gln1_bwd_kernel<true, 256, 1><<<rows_h, 256, 0, s>>>(*p, *g, acc);
            else if (nqv == 2) tcn_gln1_bwd_kernel<true, 256, 2><<<rows_h, 256, 0, s>>>(*p, *g, acc);
            else if (nqv == 4) tcn_gln1_bwd_kernel<true, 256, 4><<<rows_h, 256, 0, s>>>(*p, *g, acc);
            else if (nqv == 12) tcn_gln1_bwd_kernel<true, 128, 2><<<rows_h, 128, 0, s>>>(*p, *g, acc);
            else tcn_gln1_bwd_kernel<true, 128, 4><<<rows_h, 128, 0, s>>>(*p, *g, acc);
        } else tcn_gln1_bwd_kernel<false, 256, 2><<<rows_h, 256, 0, s>>>(*p, *g, acc);
    }
    rc = check_launch("tcn_block_bwd(hidden)");
    if (rc) return rc;
    const int nf = p->Chid > n2 ? p->Chid : n2;
    if (tail_side) {
        // W (dW1q) + F on the side stream, behind the dW2 wgrad (same partial-tile buffer, stream order)
        cudaEventRecord(side_event(2), s);
        cudaStreamWaitEvent(sw, side_event(2), 0);
        rc = tcw::run(g->dY1, p->x_op, p->B, p->M, p->ld, p->Chid, p->Cio, part, part_cap, p->quant ? p->q_in.rmin : nullptr,
                      p->quant ? p->q_in.rmax : nullptr, p->dws1, acc + L.db1, g->dW1q, sw);
        { FQSS_PROF("tcn_bwd_misc", sw); tcn_bwd_finalize_kernel<<<(nf + 255) / 256, 256, 0, sw>>>(*p, *g, acc, lean ? 1 : 0); }
        cudaEventRecord(side_event(3 + flip), sw);
        g_pend_tail[flip] = true;
        if (rc) return rc;
    }
    // G: g_x_in = Wc1T-GEMM(dY1) (+ g_xd)   (K = Chid, N = Cio) -> fp32
    {
        tcg::Args a{};
        a.B = p->B; a.M = p->M; a.K = p->Chid; a.N = p->Cio; a.ld = p->ld; a.s1 = nullptr; a.s0 = nullptr; a.out_f32 = g->g_x_in;
        a.addend = p->has_res ? g->g_xd : nullptr;
        rc = tcg::run(p->has_res ? tcg::EPI_ADD : tcg::EPI_STORE, g->dY1, p->Wc1T, a, s);
        if (rc) return rc;
    }
    if (tail_side) return check_launch("tcn_block_bwd(tail on the side stream)");
    // W: dW1q
    if (overlap) cudaStreamWaitEvent(s, side_event(1), 0);      // join: dW2q is final, the partial-tile buffer is free again
    rc = tcw::run(g->dY1, p->x_op, p->B, p->M, p->ld, p->Chid, p->Cio, part, part_cap, p->quant ? p->q_in.rmin : nullptr,
                  p->quant ? p->q_in.rmax : nullptr, p->dws1, acc + L.db1, g->dW1q, s);
    if (rc) return rc;
    // F
    { FQSS_PROF("tcn_bwd_misc", s); tcn_bwd_finalize_kernel<<<(nf + 255) / 256, 256, 0, s>>>(*p, *g, acc, lean ? 1 : 0); }
    return check_launch("tcn_block_bwd(finalize)");
}

}  // extern "C"
