// tcn_rows.cuh -- the fused gLN2 / FQ3 / PReLU3 / depthwise / FQ2 backward row kernel of the quantised ConvBlock, written
// against the SASS (the kernel is issue-bound: ~70 instructions per frame in its first version, HBM needs <= ~40):
//
//  * table entries are PAIRS read with one LDS.64: phase A {xhat3*nC + nB, mask4 ? rstd3*gamma2 : 0} turns gLN2 + FQ4
//    backward into one FMA; phase B {D2', a2} replaces the conversions and selects that rebuilt them from t2.  Tables sit
//    on a 2 KB boundary of shared memory, so "extract byte, scale, add base" is SHF + LOP3 ((x & 0x7f8) | base);
//  * the float of the table ADDRESS doubles as the float of the code (base + 8c is exact in fp32): de-quantised value
//    and xhat are FMAs on it, the base is removed from the sums once per row in fp64;
//  * clipped codes carry their range weight D shifted by MASK_OFF: the STE mask is ONE compare, and
//    sum g*D = sum g*D' - MASK_OFF * sum g*(1-m), the second sum being needed anyway;
//  * phase A reads the saved code of a3 (1 B/frame).  The STE mask of FQ3 is two compares of z = PReLU(y3) against the
//    exact thresholds the forward left in rc3[12..13] (tcn_common.cuh), feeding one select; the range weight c - t is
//    (decode(c) - z) / delta with the division pulled out of the sum; clipped-high elements contribute 255 * g through
//    sum u * c (u = g - g*mask is zero inside);
//  * the twelve row sums are reduced through shared memory (transposed: 12 stores + 3 short warp reductions per
//    thread instead of 12 full warp reductions).
//
// A persistent variant (one resident wave of CTAs walking over rows, inputs by cp.async) was measured slower on B200:
// with ~100 registers per thread it cannot keep enough loads in flight; one short-lived CTA per row with 4 quads of
// loads in flight per thread and 7 CTAs per SM can (A/B runs: profiles/gpu_r02[b-j].sh).
#pragma once
#include "tcn_bwd_common.cuh"

namespace fqss {

// x if lo <= z < hi else 0 -- two compares feeding ONE select (setp.ge, then setp.lt.and)
__device__ __forceinline__ float sel_in(float x, float z, float lo, float hi) {
    float r;
    asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %2, %3;\n\tsetp.lt.and.f32 p, %2, %4, p;\n\tselp.f32 %0, %1, 0f00000000, p;\n\t}"
        : "=f"(r) : "f"(x), "f"(z), "f"(lo), "f"(hi));
    return r;
}
// (a, b) as ONE aligned register pair (the packed FP32x2 instructions need one; without this the compiler re-packs per use)
__device__ __forceinline__ float2 pair2(float a, float b) {
    unsigned long long v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a), "f"(b));
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
// Masked table entries: D (range weight: c - t inside, c outside) is stored as D + MASK_OFF for clipped codes, so the STE
// mask is ONE compare (D' < MASK_CUT) and sum g*D = sum g*D' - MASK_OFF * sum g*(1 - m), the second sum being needed anyway.
constexpr float MASK_OFF = 512.f, MASK_CUT = 256.f;
__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// transposed block reduction of NV per-thread floats: red[k][tid] in shared memory, warp w sums the values k = w, w+NW, ...
// (fp32 over the NTH partials), lane 0 of that warp gets them in out[] (index j <-> k = w + j*NW)
template <int NTH, int NV>
__device__ __forceinline__ void block_sums_t(const float (&s)[NV], float* red, float (&out)[(NV + NTH / 32 - 1) / (NTH / 32)]) {
    constexpr int NW = NTH / 32;
#pragma unroll
    for (int k = 0; k < NV; ++k) red[k * NTH + threadIdx.x] = s[k];
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < (NV + NW - 1) / NW; ++j) {
        const int k = w + j * NW;
        float v = 0.f;
        if (k < NV) {
#pragma unroll
            for (int i = 0; i < NW; ++i) v += red[k * NTH + lane + 32 * i];
        }
        out[j] = warp_sum(v);
    }
}

// dynamic shared memory: [slack to a 2 KB boundary | tabA 2 KB | tabB 2 KB | dpad | ld | dpad floats]
// (the row buffer doubles as the scratch of the final reduction: at least 12 x 128 floats behind the left halo)
__host__ __device__ inline size_t p2d_lean_smem(int64_t ld, int dil) {
    size_t rowf = (size_t)ld + 2 * dw_pad(dil);
    const size_t need = (size_t)12 * 128 + dw_pad(dil);
    if (rowf < need) rowf = need;
    return 2048 + 4096 + rowf * sizeof(float);
}

struct LdA { float4 y; uint2 g; uint32_t cw; };

template <int DMODE, int NTH, int NQ, int MINB>
__global__ void __launch_bounds__(NTH, MINB) tcn_gln2_dw_bwd_lean_kernel(const fqss_tcn_block p, const fqss_tcn_block_grads g, double* acc) {
    extern __shared__ __align__(16) uint8_t dsm_raw[];
    const AccLayout L(p.B, p.Cio, p.Chid);
    const int64_t r = blockIdx.x;
    const int b = (int)(r / p.Chid), c = (int)(r % p.Chid);
    const int M = p.M, d = p.dil, dpad = dw_pad(d), ld = (int)p.ld;
    const int nfull = M >> 2;
    const float4* y3 = reinterpret_cast<const float4*>(p.y3 + r * p.ld);
    const uint32_t* c3 = reinterpret_cast<const uint32_t*>(p.code3 + r * p.ld);
    const uint2* ga4 = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(g.g_hid_a) + r * p.ld);
    const uint32_t* c1 = reinterpret_cast<const uint32_t*>(p.code1 + r * p.ld);
    auto loadA = [&](int v) {
        LdA t;
        t.y = __ldg(y3 + v);
        t.g = __ldg(ga4 + v);
        t.cw = __ldg(c3 + v);
        return t;
    };
    // the first trip's loads are in flight while constants and tables are built: a CTA lives for ONE short row, so
    // "constants -> tables -> barrier -> data" in sequence would expose two full memory latencies per CTA
    LdA first[NQ];
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
        const int v = threadIdx.x + k * NTH;
        if (v < nfull) first[k] = loadA(v);
    }
    const uint32_t raw_s = smem_addr(dsm_raw);
    const uint32_t tab_s = (raw_s + 2047u) & ~2047u;
    uint8_t* tab_g = dsm_raw + (tab_s - raw_s);
    float2* tabA = reinterpret_cast<float2*>(tab_g);          // {xhat3*nC + nB, mask4 ? rstd3*gamma2 : 0}
    float2* tabB = tabA + 256;                                // {D2 (+ MASK_OFF where FQ2 clips), a2}
    float* dsm = reinterpret_cast<float*>(tab_g + 4096);
    float* row = dsm + dpad;
    const uint32_t ta = vreg(tab_s), tbB = vreg(tab_s + 2048u);
    const Hidden3 h3 = load_hidden3(p, b, c);
    const Hidden1 h1 = load_hidden1(p, b, c);
    const float zlo = vreg(__ldg(p.rc3 + 12)), zhi = vreg(__ldg(p.rc3 + 13));
    // zero the halo and every frame from the ragged quad's end up to the pitch
    for (int i = threadIdx.x; i < dpad; i += NTH) {
        dsm[i] = 0.f;
        row[ld + i] = 0.f;
    }
    for (int i = ((M + 3) & ~3) + threadIdx.x; i < ld; i += NTH) row[i] = 0.f;
    {
        const float invN = __fdividef(1.f, (float)p.Chid * (float)p.M);
        const float A = h3.g.rstd * h3.g.gamma;
        double sm1, sm2;
        samp_get(acc, L, 2, b, sm1, sm2);
        const float nB = -h3.g.rstd * invN * (float)sm1;
        const float nC = -h3.g.rstd * invN * (float)sm2;
        for (int i = threadIdx.x; i < 256; i += NTH) {
            const float4 e = chain_bwd_entry(h3.q3, h3.g, h3.q4, i, false);            // {-, mask4, D4, xhat3}
            tabA[i] = make_float2(fmaf(e.w, nC, nB), e.y != 0.f ? A : 0.f);
            const float n1 = gln_apply(h1.g, actqf_decode(h1.q1, (float)i));
            const float t2 = actqf_t(h1.q2, n1);
            const bool in2 = actqf_inside(h1.q2, t2);
            const float c2 = actqf_unbias(actqf_biased(h1.q2, t2));
            const float D2 = in2 ? (c2 - t2) : c2;
            tabB[i] = make_float2(in2 ? D2 : D2 + MASK_OFF, actqf_decode(h1.q2, c2));
        }
    }
    __syncthreads();
    // ---------------- phase A: gLN2 + FQ4 + FQ3 + PReLU3 backward -> row ----------------
    //   a0 = sum gz*(decode3(c) - z)  (x 1/delta3), a0c = sum (ga3 - gz)*(ta + 8c), a1 = sum (ga3 - gz), a3 = sum min(y,0)*gz
    float2 a0 = f2s(0.f), a0c = f2s(0.f), a1 = f2s(0.f), a3 = f2s(0.f);
    {
        const float slope3 = vreg(h3.slope);
        const float2 slope3v = f2s(slope3);
        const float dec3s = 0.125f * h3.q3.delta;
        // decode3(c) = delta3*c + min3 from the float of the table address (ta + 8c)
        const float2 dec3a = f2s(dec3s), dec3b = f2s(fmaf(-dec3s, (float)ta, h3.q3.mn));
        auto body = [&](int v, const LdA& t, auto tail_tag) {
            constexpr bool TAIL = decltype(tail_tag)::value;
            const float2 yy[2] = {lo2(t.y), hi2(t.y)};
            float2 o[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint32_t w = j ? t.g.y : t.g.x;
                const uint32_t ax = tab_addr<3>(t.cw, 2 * j, ta), ay = tab_addr<3>(t.cw, 2 * j + 1, ta);
                const float2 ex = lds64(ax), ey = lds64(ay);
                float2 ga3 = make_float2(fmaf(ex.y, bf16lo(w), ex.x), fmaf(ey.y, bf16hi(w), ey.x));
                if (TAIL) {
                    if (4 * v + 2 * j + 1 >= M) ga3.y = 0.f;
                    if (4 * v + 2 * j >= M) ga3.x = 0.f;
                }
                const float2 ys = __fmul2_rn(yy[j], slope3v);
                const bool px = yy[j].x > 0.f, py = yy[j].y > 0.f;
                const float2 z = make_float2(px ? yy[j].x : ys.x, py ? yy[j].y : ys.y);
                const float2 gz = make_float2(sel_in(ga3.x, z.x, zlo, zhi), sel_in(ga3.y, z.y, zlo, zhi));
                const float2 u = __fadd2_rn(ga3, neg2(gz));
                const float2 cf = make_float2((float)ax, (float)ay);
                a1 = __fadd2_rn(a1, u);
                a0c = __ffma2_rn(u, cf, a0c);
                const float2 dq = __ffma2_rn(cf, dec3a, dec3b);      // decode3(code)
                a0 = __ffma2_rn(gz, __fadd2_rn(dq, neg2(z)), a0);
                o[j] = __fmul2_rn(gz, make_float2(px ? 1.f : slope3, py ? 1.f : slope3));
                a3 = __ffma2_rn(make_float2(fminf(yy[j].x, 0.f), fminf(yy[j].y, 0.f)), gz, a3);
            }
            *reinterpret_cast<float4*>(row + 4 * v) = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);
        };
        // first trip (already loaded), then the batched loop
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
            const int v = threadIdx.x + k * NTH;
            if (v < nfull) body(v, first[k], std::false_type{});
        }
        for (int base = threadIdx.x + NQ * NTH; base < nfull; base += NQ * NTH) {
            LdA dd[NQ];
#pragma unroll
            for (int k = 0; k < NQ; ++k) {
                const int v = base + k * NTH;
                if (v < nfull) dd[k] = loadA(v);
            }
#pragma unroll
            for (int k = 0; k < NQ; ++k) {
                const int v = base + k * NTH;
                if (v < nfull) body(v, dd[k], std::false_type{});
            }
        }
        if ((M & 3) && (int)threadIdx.x == (nfull % NTH)) body(nfull, loadA(nfull), std::true_type{});
    }
    __syncthreads();
    // ---------------- phase B: depthwise + FQ2 backward, gLN1 row sums ----------------
    //   b0 = sum ga2*D2' (two scalar lanes), b1 = sum ga2*(1-m2), s2 = sum gn1, scf = sum gn1*(tbB + 8 c1)
    //   taps: d0,d1,d2 = sum a2*g[+d,0,-d], d3 = sum g
    float2 b1 = f2s(0.f), s2 = f2s(0.f), scf = f2s(0.f), d0 = f2s(0.f), d1 = f2s(0.f), d2 = f2s(0.f), d3 = f2s(0.f);
    float b0x = 0.f, b0y = 0.f;
    {
        const float2 w0 = f2s(__ldg(p.wdw + c * 3)), w1 = f2s(__ldg(p.wdw + c * 3 + 1)), w2 = f2s(__ldg(p.wdw + c * 3 + 2));
        uint2* gn1o = reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.g_hid_a) + r * p.ld);
        auto load = [&](int v) { return __ldg(c1 + v); };
        auto body = [&](int v, const uint32_t cw, auto tail_tag) {
            constexpr bool TAIL = decltype(tail_tag)::value;
            float4 gL, gC, gR;
            dw_taps<DMODE>(row, v, d, gL, gC, gR);
            float2 o[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float2 gc = j ? hi2(gC) : lo2(gC), gl = j ? hi2(gL) : lo2(gL), gr = j ? hi2(gR) : lo2(gR);
                // y3[m'] = sum_k w_k a2[m' + (k-1)d]  =>  d/da2[m] = w0 g[m+d] + w1 g[m] + w2 g[m-d]
                float2 ga2 = __ffma2_rn(w0, gr, __ffma2_rn(w1, gc, __fmul2_rn(w2, gl)));
                const uint32_t ax = tab_addr<3>(cw, 2 * j, tbB), ay = tab_addr<3>(cw, 2 * j + 1, tbB);
                const float2 ex = lds64(ax), ey = lds64(ay);
                float2 a2 = TAIL ? make_float2(ex.y, ey.y) : pair2(ex.y, ey.y);
                if (TAIL) {                                   // frames >= M: no gradient, and a2 there is not part of the row
                    if (4 * v + 2 * j + 1 >= M) { ga2.y = 0.f; a2.y = 0.f; }
                    if (4 * v + 2 * j >= M) { ga2.x = 0.f; a2.x = 0.f; }
                }
                const float2 gn1 = make_float2(ex.x < MASK_CUT ? ga2.x : 0.f, ey.x < MASK_CUT ? ga2.y : 0.f);
                b0x = fmaf(ga2.x, ex.x, b0x);
                b0y = fmaf(ga2.y, ey.x, b0y);
                b1 = __fadd2_rn(b1, __fadd2_rn(ga2, neg2(gn1)));
                d0 = __ffma2_rn(a2, gr, d0);      // dW_0 = sum_m a2[m] g[m+d]
                d1 = __ffma2_rn(a2, gc, d1);
                d2 = __ffma2_rn(a2, gl, d2);      // dW_2 = sum_m a2[m] g[m-d]
                d3 = __fadd2_rn(d3, gc);
                s2 = __fadd2_rn(s2, gn1);
                scf = __ffma2_rn(gn1, make_float2((float)ax, (float)ay), scf);
                o[j] = gn1;
            }
            gn1o[v] = float4_to_bf16x4(o[0].x, o[0].y, o[1].x, o[1].y);
        };
        FQSS_ROW_LOOP_BATCH(NTH, NQ, load, body, M);
    }
    __syncthreads();      // the row buffer becomes the reduction scratch
    // value k is reduced by warp k % 4: warp 0 {a0, a0c, a1} -> q3, warp 1 {a3, b0, b1} -> slope3, q2,
    // warp 2 {s2, scf, d3} -> gLN1 row sums, bias, warp 3 {d0, d1, d2} -> taps
    const float sv[12] = {hsum(a0), hsum(a3), hsum(s2), hsum(d0), hsum(a0c), b0x + b0y, hsum(scf), hsum(d1), hsum(a1), hsum(b1), hsum(d3), hsum(d2)};
    float t3[3];
    static_assert(NTH == 128, "p2d_lean_smem sizes the reduction scratch for 128 threads");
    block_sums_t<NTH, 12>(sv, row, t3);
    if ((threadIdx.x & 31) == 0) {
        const int w = threadIdx.x >> 5;
        const double x0 = (double)t3[0], x1 = (double)t3[1], x2 = (double)t3[2];
        if (w == 0) {
            // sD(q3) = sum_in gz*(c - t) + 255 * sum_above (ga3 - gz);  sum u*(ta + 8c) = ta * sum u + 8 * 255 * sum_above u
            atomicAdd(acc + L.qs(c) + 2 * Q3, x0 * (double)h3.q3.inv + 0.125 * (x1 - (double)ta * x2));
            atomicAdd(acc + L.qs(c) + 2 * Q3 + 1, x2);
        } else if (w == 1) {
            atomicAdd(acc + L.qs(c) + AccLayout::SLOPE_OFF + 1, x0);
            atomicAdd(acc + L.qs(c) + 2 * Q2, x1 - (double)MASK_OFF * x2);
            atomicAdd(acc + L.qs(c) + 2 * Q2 + 1, x2);
        } else if (w == 2) {
            // xhat1 = (delta1*c + min1 - mu1) * rstd1 = xa * (8c) + xb
            const double xa = 0.125 * (double)h1.q1.delta * (double)h1.g.rstd, xb = ((double)h1.q1.mn - (double)h1.g.mu) * (double)h1.g.rstd;
            const double s3 = xa * (x1 - (double)tbB * x0) + xb * x0;      // sum gn1 * xhat1
            const double gm = (double)h1.g.gamma;
            atomicAdd(acc + L.ss(L.sslot1, b, c), gm * x0);               // per-sample gLN1 sums (no reduce launch)
            atomicAdd(acc + L.ss(L.sslot1, b, c) + 1, gm * s3);
            atomicAdd(acc + L.gln1 + 2 * c, x0);                           // dbeta1, dgamma1
            atomicAdd(acc + L.gln1 + 2 * c + 1, s3);
            atomicAdd(acc + L.dbdw + c, x2);
        } else {
            atomicAdd(acc + L.dwdw + 3 * c, x0);
            atomicAdd(acc + L.dwdw + 3 * c + 1, x1);
            atomicAdd(acc + L.dwdw + 3 * c + 2, x2);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// P1 (quantised): g_a4 (bf16), code3 -> per-sample gLN2 sums, dgamma2 / dbeta2, range sums of FQ4.  Everything is a
// function of the saved code of a3: one LDS.32 (D4', shifted by MASK_OFF where FQ4 clips) per frame; xhat3 is an FMA
// on the float of the table address.  3 B/frame of HBM traffic, so the first trip's loads are issued before the
// constants and the table.  dynamic shared memory: [slack to a 1 KB boundary | table 1 KB | 4 x 128 floats]
// ---------------------------------------------------------------------------------------------
constexpr size_t P1_LEAN_SMEM = 1024 + 1024 + 4 * 128 * sizeof(float);

template <int NTH, int NQ>
__global__ void __launch_bounds__(NTH) tcn_gln2_sums_lean_kernel(const fqss_tcn_block p, const fqss_tcn_block_grads g, double* acc) {
    extern __shared__ __align__(16) uint8_t dsm_raw[];
    static_assert(NTH == 128, "P1_LEAN_SMEM sizes the reduction scratch for 128 threads");
    const AccLayout L(p.B, p.Cio, p.Chid);
    const int64_t r = blockIdx.x;
    const int b = (int)(r / p.Chid), c = (int)(r % p.Chid);
    const int M = p.M;
    const uint32_t* c3 = reinterpret_cast<const uint32_t*>(p.code3 + r * p.ld);
    const uint2* ga4 = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(g.g_hid_a) + r * p.ld);
    const int nfull = M >> 2;
    uint32_t cw[NQ];
    uint2 gw[NQ];
    // quads beyond the last full one load as zeros: a zero gradient adds nothing to any of the four sums (every table entry
    // is finite), so the element loop runs without frame checks; the ragged quad (M % 4 frames) is handled once, below
    auto issue = [&](int base) {
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
            const int v = base + k * NTH;
            const bool ok = v < nfull;
            cw[k] = ok ? __ldg(c3 + v) : 0u;
            gw[k] = ok ? __ldg(ga4 + v) : make_uint2(0u, 0u);
        }
    };
    issue(threadIdx.x);
    const uint32_t raw_s = smem_addr(dsm_raw);
    const uint32_t tab_s = (raw_s + 1023u) & ~1023u;
    float* tabD = reinterpret_cast<float*>(dsm_raw + (tab_s - raw_s));
    float* red = tabD + 256;
    const uint32_t tb = vreg(tab_s);
    const Hidden3 h = load_hidden3(p, b, c);
    for (int i = threadIdx.x; i < 256; i += NTH) {
        const float4 e = chain_bwd_entry(h.q3, h.g, h.q4, i, false);        // {-, mask4, D4, xhat3}
        tabD[i] = e.y != 0.f ? e.z : e.z + MASK_OFF;
    }
    __syncthreads();
    // a0 = sum g*D4', a1 = sum g*(1-m4), s2 = sum gn, scf = sum gn*(tb + 4c)  -- all as frame pairs (packed FP32x2)
    float2 a0 = f2s(0.f), a1 = f2s(0.f), s2 = f2s(0.f), scf = f2s(0.f);
    auto quad = [&](const uint32_t cq, const uint2 gq, const int nval) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const uint32_t w = j ? gq.y : gq.x;
            float2 gg = make_float2(bf16lo(w), bf16hi(w));
            if (2 * j + 1 >= nval) gg.y = 0.f;        // nval = 4 in the loop: folded away
            if (2 * j >= nval) gg.x = 0.f;
            const uint32_t ax = tab_addr<2>(cq, 2 * j, tb), ay = tab_addr<2>(cq, 2 * j + 1, tb);
            const float2 dd = make_float2(lds32(ax), lds32(ay));
            const float2 gn = make_float2(dd.x < MASK_CUT ? gg.x : 0.f, dd.y < MASK_CUT ? gg.y : 0.f);
            a0 = __ffma2_rn(gg, dd, a0);
            a1 = __fadd2_rn(a1, __fadd2_rn(gg, neg2(gn)));
            s2 = __fadd2_rn(s2, gn);
            scf = __ffma2_rn(gn, make_float2((float)ax, (float)ay), scf);
        }
    };
    for (int base = threadIdx.x; base < nfull; base += NQ * NTH) {
        if (base != (int)threadIdx.x) issue(base);
#pragma unroll
        for (int k = 0; k < NQ; ++k) quad(cw[k], gw[k], 4);
    }
    if ((M & 3) && (int)threadIdx.x == (nfull % NTH)) quad(__ldg(c3 + nfull), __ldg(ga4 + nfull), M & 3);
    const float sv[4] = {hsum(a0), hsum(a1), hsum(s2), hsum(scf)};
    float t1[1];
    block_sums_t<NTH, 4>(sv, red, t1);
    // every warp now holds ONE of the four totals in lane 0; combine through shared memory
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t1[0];
    __syncthreads();
    if (threadIdx.x == 0) {
        const double x0 = (double)red[0], x1 = (double)red[1], x2 = (double)red[2], x3 = (double)red[3];
        atomicAdd(acc + L.qs(c) + 2 * Q4, x0 - (double)MASK_OFF * x1);
        atomicAdd(acc + L.qs(c) + 2 * Q4 + 1, x1);
        // xhat3 = (delta3*c + min3 - mu) * rstd = xa * (4c) + xb
        const double xa = 0.25 * (double)h.q3.delta * (double)h.g.rstd, xb = ((double)h.q3.mn - (double)h.g.mu) * (double)h.g.rstd;
        const double s3 = xa * (x3 - (double)tb * x2) + xb * x2;
        const double gm = (double)h.g.gamma;
        atomicAdd(acc + L.ss(L.sslot2, b, c), gm * x2);
        atomicAdd(acc + L.ss(L.sslot2, b, c) + 1, gm * s3);
        atomicAdd(acc + L.gln2 + 2 * c, x2);
        atomicAdd(acc + L.gln2 + 2 * c + 1, s3);
    }
}

}  // namespace fqss
