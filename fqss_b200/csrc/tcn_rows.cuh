// tcn_rows.cuh -- persistent-row backward kernels of the fused ConvBlock (quantised model):
//   P1   tcn_gln2_sums_rows_kernel     g_a4 (bf16), code3 -> per-sample gLN2 sums, dgamma2 / dbeta2, range sums of FQ4
//   P2D  tcn_gln2_dw_bwd_rows_kernel   gLN2 / FQ3 / PReLU3 backward -> g_y3 row in shared memory -> depthwise + FQ2
//                                      backward -> g_n1 (bf16), tap / bias / range / slope sums, gLN1 sums
//
// One CTA per (sample, channel) row (16 384 CTAs at batch 32) spends a quarter of its instructions on per-CTA fixed work:
// constants, halo zeroing, eleven block reductions, the exit.  Here the grid is exactly ONE resident wave; jobs =
// (channel, RJ consecutive samples) are dealt round-robin.  Channel constants are loaded once per job, the halo is
// zeroed once per CTA, the global sums (range gradients, slope) stay in registers for the CTA's lifetime, the
// per-channel sums (taps, bias, dgamma, dbeta) for a job, and the per-row work shrinks to the two code tables plus a
// two-value reduction.  The per-sample gLN sums go straight to the accumulator block (fp64 atomics, 2 per row), so the
// separate reduce launches disappear.  While a row is processed the CTA's next row is prefetched into L2.
//
// Element loops (the kernels are issue-bound, so they are written against the SASS):
//  * table entries are PAIRS read with one LDS.64: phase A {xhat3*nC + nB, mask4 ? rstd3*gamma2 : 0} turns gLN2 + FQ4
//    backward into one FMA; phase B {D2 | mask2, a2} replaces the conversions and selects that rebuilt them from t2.
//    Tables are 2 KB-aligned in shared memory, so "extract byte, scale, add base" is SHF + LOP3 ((x & mask) | base);
//  * the float of the table ADDRESS doubles as the float of the code (base + 8c is exact in fp32): de-quantised value
//    and xhat are FMAs on it, the base is removed from the sums once per row / CTA in fp64;
//  * phase A reads the saved code of a3 (1 B/frame).  The STE mask of FQ3 is two compares of z = PReLU(y3) against the
//    exact thresholds z_lo = min{z : t(z) >= -0.5}, z_hi = min{z : t(z) >= 255.5} (t is monotone in z, so this is the
//    forward's -0.5 <= t < 255.5 bit for bit), and the range weight c - t is (decode(c) - z) / delta with the division
//    pulled out of the sum; clipped-high elements contribute 255 * g through sum u * c (u = g - g*mask is zero inside).
#pragma once
#include "tcn_bwd_common.cuh"

namespace fqss {

__device__ __forceinline__ unsigned f2key(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
// smallest float z with (z - min) / delta >= target (IEEE division; bit-identical to actqf_t wherever that is finite)
__device__ inline float fq_threshold(const ActQF& q, float target) {
    unsigned lo = f2key(-3.0e38f), hi = f2key(3.0e38f);
    while (hi - lo > 1u) {
        const unsigned mid = lo + ((hi - lo) >> 1);
        const float t = __fdiv_rn(__fsub_rn(key2f(mid), q.mn), q.delta);
        if (t >= target) hi = mid; else lo = mid;
    }
    return key2f(hi);
}

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
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// the CTA prefetches [p, p + bytes) into L2, one 128 B line per thread and trip
template <int NTH>
__device__ __forceinline__ void prefetch_row(const void* p, int bytes) {
    const char* c = reinterpret_cast<const char*>(p);
    for (int o = threadIdx.x * 128; o < bytes; o += NTH * 128) prefetch_l2(c + o);
}

template <int NTH>
__device__ __forceinline__ void row_partials_store(float a, float b, float2* slot) {
    a = warp_sum(a);
    b = warp_sum(b);
    if ((threadIdx.x & 31) == 0) slot[threadIdx.x >> 5] = make_float2(a, b);
}
template <int NTH>
__device__ __forceinline__ double2 row_partials_load(const float2* slot) {
    double x = 0.0, y = 0.0;
#pragma unroll
    for (int w = 0; w < NTH / 32; ++w) { x += (double)slot[w].x; y += (double)slot[w].y; }
    return make_double2(x, y);
}

// Row loop: unpredicated trips of NQ quads per thread, then single-quad trips, then the ragged quad (M % 4 frames) by
// one thread through the same body with TAIL = true.  `load(v)` returns a plain struct of raw words.
#define FQSS_ROWS_LOOP(NTH_, NQ_, load, body, M)                                                          \
    do {                                                                                                  \
        const int nfull_ = (M) >> 2;                                                                      \
        int v_ = threadIdx.x;                                                                             \
        for (; v_ + ((NQ_)-1) * (NTH_) < nfull_; v_ += (NQ_) * (NTH_)) {                                  \
            decltype(load(0)) d_[NQ_];                                                                    \
            _Pragma("unroll") for (int k_ = 0; k_ < (NQ_); ++k_) d_[k_] = load(v_ + k_ * (NTH_));         \
            _Pragma("unroll") for (int k_ = 0; k_ < (NQ_); ++k_) body(v_ + k_ * (NTH_), d_[k_], std::false_type{}); \
        }                                                                                                 \
        for (; v_ < nfull_; v_ += (NTH_)) body(v_, load(v_), std::false_type{});                          \
        if (((M)&3) && (int)threadIdx.x == (nfull_ % (NTH_))) body(nfull_, load(nfull_), std::true_type{}); \
    } while (0)

struct RowJob {
    int c, b_lo, b_hi;
};
__device__ __forceinline__ RowJob row_job(int job, int Chid, int B, int RJ) {
    RowJob j;
    j.c = job % Chid;
    j.b_lo = (job / Chid) * RJ;
    j.b_hi = min(B, j.b_lo + RJ);
    return j;
}

// ---------------------------------------------------------------------------------------------
// P1
// ---------------------------------------------------------------------------------------------
template <int NTH, int NQ, int MINB>
__global__ void __launch_bounds__(NTH, MINB) tcn_gln2_sums_rows_kernel(const fqss_tcn_block p, const fqss_tcn_block_grads g, double* acc, int RJ) {
    __shared__ __align__(1024) float tabD[256];     // D4 (+ MASK_OFF where FQ4 clips)
    __shared__ double sh[3 * 32];
    __shared__ float2 part[2][NTH / 32];
    const AccLayout L(p.B, p.Cio, p.Chid);
    const int M = p.M;
    const int jobs_per_ch = (p.B + RJ - 1) / RJ, njobs = jobs_per_ch * p.Chid;
    const ActQF q3 = load_actqf_rc(p.rc3), q4 = load_actqf_rc(p.rc3 + 4);
    const uint32_t tb = vreg(smem_addr(tabD));
    const double tbase = (double)tb;
    const __nv_bfloat16* ga_all = reinterpret_cast<const __nv_bfloat16*>(g.g_hid_a);
    // global: a0 = sum g*D4 (two scalar lanes), a1 = sum g*(1-m4)
    float a0x = 0.f, a0y = 0.f;
    float2 a1 = f2s(0.f);
    unsigned it = 0;
    for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
        const RowJob J = row_job(job, p.Chid, p.B, RJ);
        const int c = J.c;
        const float gamma = __ldg(p.gn2_w + c), beta = __ldg(p.gn2_b + c);
        double dbet = 0.0, dgam = 0.0;                  // thread 0: per-channel sums over this job's samples
        float xa_prev = 0.f, xb_prev = 0.f;
        for (int b = J.b_lo; b < J.b_hi; ++b, ++it) {
            const int64_t r = (int64_t)b * p.Chid + c;
            const uint32_t* c3 = reinterpret_cast<const uint32_t*>(p.code3 + r * p.ld);
            const uint2* ga4 = reinterpret_cast<const uint2*>(ga_all + r * p.ld);
            {   // next row of this CTA -> L2
                int nb = b + 1, nc = c;
                if (nb >= J.b_hi) {
                    const int nj = job + gridDim.x;
                    if (nj < njobs) { const RowJob N = row_job(nj, p.Chid, p.B, RJ); nb = N.b_lo; nc = N.c; } else nb = -1;
                }
                if (nb >= 0) {
                    const int64_t nr = (int64_t)nb * p.Chid + nc;
                    prefetch_row<NTH>(p.code3 + nr * p.ld, M);
                    prefetch_row<NTH>(ga_all + nr * p.ld, 2 * M);
                }
            }
            GlnRow gr;
            gr.mu = __ldg(p.rc3 + RC_HDR + 2 * b);
            gr.rstd = __ldg(p.rc3 + RC_HDR + 1 + 2 * b);
            gr.gamma = gamma;
            gr.scale = __fmul_rn(gr.rstd, gamma);
            gr.shift = __fadd_rn(__fmul_rn(-gr.scale, gr.mu), beta);
            __syncthreads();                              // the previous row is done with tabD; its partial sums are visible
            if (b > J.b_lo && threadIdx.x == 0) {
                const double2 v = row_partials_load<NTH>(part[(it - 1u) & 1u]);
                // sum gn, sum gn * xhat3 = xa * (sum gn*(base + 4c) - base * sum gn) + xb * sum gn
                const double s2 = v.x, s3 = (double)xa_prev * (v.y - tbase * v.x) + (double)xb_prev * v.x;
                atomicAdd(acc + L.samp2 + 2 * (b - 1), (double)gamma * s2);
                atomicAdd(acc + L.samp2 + 2 * (b - 1) + 1, (double)gamma * s3);
                dbet += s2;
                dgam += s3;
            }
            for (int i = threadIdx.x; i < 256; i += NTH) {
                const float4 e = chain_bwd_entry(q3, gr, q4, i, false);        // {-, mask4, D4, xhat3}
                tabD[i] = e.y != 0.f ? e.z : e.z + MASK_OFF;
            }
            __syncthreads();
            float2 s2 = f2s(0.f), scf = f2s(0.f);         // per row: sum gn, sum gn * (base + 4c)
            struct Ld { uint2 g; uint32_t cw; };
            auto load = [&](int v) {
                Ld t;
                t.g = __ldg(ga4 + v);
                t.cw = __ldg(c3 + v);
                return t;
            };
            auto body = [&](int v, const Ld& t, auto tail_tag) {
                constexpr bool TAIL = decltype(tail_tag)::value;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint32_t w = j ? t.g.y : t.g.x;
                    float gx = bf16lo(w), gy = bf16hi(w);
                    if (TAIL) {
                        if (4 * v + 2 * j + 1 >= M) gy = 0.f;
                        if (4 * v + 2 * j >= M) gx = 0.f;
                    }
                    const uint32_t ax = tab_addr<2>(t.cw, 2 * j, tb), ay = tab_addr<2>(t.cw, 2 * j + 1, tb);
                    const float dx = lds32(ax), dy = lds32(ay);
                    const float2 gg = make_float2(gx, gy);
                    const float2 gn = make_float2(dx < MASK_CUT ? gx : 0.f, dy < MASK_CUT ? gy : 0.f);
                    a0x = fmaf(gx, dx, a0x);
                    a0y = fmaf(gy, dy, a0y);
                    a1 = __fadd2_rn(a1, __fadd2_rn(gg, neg2(gn)));
                    s2 = __fadd2_rn(s2, gn);
                    scf = __ffma2_rn(gn, make_float2((float)ax, (float)ay), scf);
                }
            };
            FQSS_ROWS_LOOP(NTH, NQ, load, body, M);
            // xhat3 = (delta3*c + min3 - mu) * rstd = xa * (4c) + xb
            xa_prev = 0.25f * q3.delta * gr.rstd;
            xb_prev = (q3.mn - gr.mu) * gr.rstd;
            row_partials_store<NTH>(hsum(s2), hsum(scf), part[it & 1u]);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const double2 v = row_partials_load<NTH>(part[(it - 1u) & 1u]);
            const double s2 = v.x, s3 = (double)xa_prev * (v.y - tbase * v.x) + (double)xb_prev * v.x;
            atomicAdd(acc + L.samp2 + 2 * (J.b_hi - 1), (double)gamma * s2);
            atomicAdd(acc + L.samp2 + 2 * (J.b_hi - 1) + 1, (double)gamma * s3);
            atomicAdd(acc + L.gln2 + 2 * c, dbet + s2);
            atomicAdd(acc + L.gln2 + 2 * c + 1, dgam + s3);
        }
    }
    const float s[3] = {a0x, a0y, hsum(a1)};
    double v[3];
    block_sum_fd<3>(s, v, sh);
    if (threadIdx.x == 0) {
        atomicAdd(acc + L.q + 2 * Q4, v[0] + v[1] - (double)MASK_OFF * v[2]);
        atomicAdd(acc + L.q + 2 * Q4 + 1, v[2]);
    }
}

// ---------------------------------------------------------------------------------------------
// P2D
// ---------------------------------------------------------------------------------------------
// Inputs reach the CTA through cp.async (global -> shared, no registers, no scoreboard stalls): while phase B of row i
// runs, the phase-A inputs of the CTA's next row (y3, g_a4, code3) stream into `bufA`; while phase A runs, the row's
// code1 streams into `bufB`.  Per row: tables -> [wait A] barrier X -> issue code1 -> phase A -> [wait code1] barrier Y
// -> issue next A -> phase B.  Two table sets (row parity) make the table build independent of the previous phase B.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// the CTA copies `bytes` (multiple of 16 / 8) from src to shared address dst
template <int NTH>
__device__ __forceinline__ void cta_copy16(uint32_t dst, const void* src, int bytes) {
    const char* c = reinterpret_cast<const char*>(src);
    for (int o = threadIdx.x * 16; o < bytes; o += NTH * 16) cp_async16(dst + o, c + o);
}
template <int NTH>
__device__ __forceinline__ void cta_copy8(uint32_t dst, const void* src, int bytes) {
    const char* c = reinterpret_cast<const char*>(src);
    for (int o = threadIdx.x * 8; o < bytes; o += NTH * 8) cp_async8(dst + o, c + o);
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds64u(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds32u(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

// dynamic shared memory of P2D:
//   [slack to a 2 KB boundary | 2 x {tabA 2 KB, tabB 2 KB} | dpad | ld | dpad floats | bufA: y3 4*ld, g 2*ld, code3 ld | bufB: code1 ld]
__host__ __device__ inline size_t p2d_rows_smem(int64_t ld, int dil) {
    return 2048 + 8192 + ((size_t)ld + 2 * dw_pad(dil)) * sizeof(float) + (size_t)ld * 8;
}

template <int DMODE, int NTH, int MINB>
__global__ void __launch_bounds__(NTH, MINB) tcn_gln2_dw_bwd_rows_kernel(const fqss_tcn_block p, const fqss_tcn_block_grads g, double* acc, int RJ) {
    extern __shared__ __align__(16) uint8_t dsm_raw[];
    __shared__ double sh[7 * 32];
    __shared__ float2 part[2][NTH / 32];
    __shared__ float zthr[2];
    const AccLayout L(p.B, p.Cio, p.Chid);
    const int M = p.M, d = p.dil, dpad = dw_pad(d), ld = (int)p.ld;
    const int nfull = M >> 2;
    const int jobs_per_ch = (p.B + RJ - 1) / RJ, njobs = jobs_per_ch * p.Chid;
    // two table sets (row parity) on a 2 KB boundary: "byte -> entry address" is SHF + LOP3 ((x & 0x7f8) | base)
    const uint32_t raw_s = smem_addr(dsm_raw);
    const uint32_t tab_s = (raw_s + 2047u) & ~2047u;
    uint8_t* tab_g = dsm_raw + (tab_s - raw_s);
    float* dsm = reinterpret_cast<float*>(tab_g + 8192);
    float* row = dsm + dpad;
    const uint32_t bufY = tab_s + 8192u + (uint32_t)(ld + 2 * dpad) * 4u;      // 16-byte aligned: ld % 8 == 0, dpad % 4 == 0
    const uint32_t bufG = bufY + 4u * ld, bufC3 = bufG + 2u * ld, bufC1 = bufC3 + (uint32_t)ld;
    const ActQF q1 = load_actqf_rc(p.rc1), q2 = load_actqf_rc(p.rc1 + 4), q3 = load_actqf_rc(p.rc3), q4 = load_actqf_rc(p.rc3 + 4);
    const __nv_bfloat16* ga_all = reinterpret_cast<const __nv_bfloat16*>(g.g_hid_a);
    const int bytesM = ((M + 3) >> 2) << 2;      // frames covered by the quads that are read (<= ld)
    auto issueA = [&](int64_t r) {
        cta_copy16<NTH>(bufY, p.y3 + r * p.ld, 4 * bytesM);
        cta_copy8<NTH>(bufG, ga_all + r * p.ld, 2 * bytesM);
        cta_copy8<NTH>(bufC3, p.code3 + r * p.ld, (bytesM + 7) & ~7);
        cp_async_commit();
    };
    if ((int)blockIdx.x < njobs) {
        const RowJob J0 = row_job(blockIdx.x, p.Chid, p.B, RJ);
        issueA((int64_t)J0.b_lo * p.Chid + J0.c);
    }
    // zero the halo and every frame from the ragged quad's end up to the pitch: written once, never touched by the row loop
    for (int i = threadIdx.x; i < dpad; i += NTH) {
        dsm[i] = 0.f;
        row[ld + i] = 0.f;
    }
    for (int i = ((M + 3) & ~3) + threadIdx.x; i < ld; i += NTH) row[i] = 0.f;
    if (threadIdx.x < 2) zthr[threadIdx.x] = fq_threshold(q3, threadIdx.x ? 255.5f : -0.5f);
    const float slope3 = vreg(__ldg(p.slope3));
    const float2 slope3v = f2s(slope3);
    const float invN = __fdividef(1.f, (float)p.Chid * (float)p.M);
    const float dec3s = 0.125f * q3.delta;
    const float2 dec3a = f2s(dec3s);
    // global sums, kept in registers across all rows of this CTA
    //   a0 = sum gz*(decode3(c) - z)  (x 1/delta3), a0c = sum (ga3 - gz)*(8c), a1 = sum (ga3 - gz), a3 = sum min(y,0)*gz
    //   b0 = sum ga2*D2' (two scalar lanes), b1 = sum ga2*(1-m2)
    float2 a0 = f2s(0.f), a0c = f2s(0.f), a1 = f2s(0.f), a3 = f2s(0.f), b1 = f2s(0.f);
    float b0x = 0.f, b0y = 0.f;
    unsigned it = 0;
    for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
        const RowJob J = row_job(job, p.Chid, p.B, RJ);
        const int c = J.c;
        const float gam1 = __ldg(p.gn1_w + c), bet1 = __ldg(p.gn1_b + c), gam2 = __ldg(p.gn2_w + c), bet2 = __ldg(p.gn2_b + c);
        const float2 w0 = f2s(__ldg(p.wdw + c * 3)), w1 = f2s(__ldg(p.wdw + c * 3 + 1)), w2 = f2s(__ldg(p.wdw + c * 3 + 2));
        // per-channel sums of this job: taps d0,d1,d2 = sum a2*g[+d,0,-d], d3 = sum g
        float2 d0 = f2s(0.f), d1 = f2s(0.f), d2 = f2s(0.f), d3 = f2s(0.f);
        double dbet = 0.0, dgam = 0.0;
        float xa_prev = 0.f, xb_prev = 0.f, tb_prev = 0.f;
        for (int b = J.b_lo; b < J.b_hi; ++b, ++it) {
            const int64_t r = (int64_t)b * p.Chid + c;
            int64_t r_next = -1;      // next row of this CTA
            {
                int nb = b + 1, nc = c;
                if (nb >= J.b_hi) {
                    const int nj = job + gridDim.x;
                    if (nj < njobs) { const RowJob N = row_job(nj, p.Chid, p.B, RJ); nb = N.b_lo; nc = N.c; } else nb = -1;
                }
                if (nb >= 0) r_next = (int64_t)nb * p.Chid + nc;
            }
            GlnRow g1, g3;
            g1.mu = __ldg(p.rc1 + RC_HDR + 2 * b); g1.rstd = __ldg(p.rc1 + RC_HDR + 1 + 2 * b); g1.gamma = gam1;
            g1.scale = __fmul_rn(g1.rstd, gam1); g1.shift = __fadd_rn(__fmul_rn(-g1.scale, g1.mu), bet1);
            g3.mu = __ldg(p.rc3 + RC_HDR + 2 * b); g3.rstd = __ldg(p.rc3 + RC_HDR + 1 + 2 * b); g3.gamma = gam2;
            g3.scale = __fmul_rn(g3.rstd, gam2); g3.shift = __fadd_rn(__fmul_rn(-g3.scale, g3.mu), bet2);
            const float A = g3.rstd * gam2;
            const float nB = -g3.rstd * invN * (float)acc[L.samp2 + 2 * b];
            const float nC = -g3.rstd * invN * (float)acc[L.samp2 + 2 * b + 1];
            // this row's table set: never the one phase B of the previous row may still be reading
            const uint32_t ta = vreg(tab_s + ((it & 1u) << 12)), tbB = ta + 2048u;
            float2* tabA = reinterpret_cast<float2*>(tab_g + ((it & 1u) << 12));
            float2* tabB = tabA + 256;
            for (int i = threadIdx.x; i < 256; i += NTH) {
                const float4 e = chain_bwd_entry(q3, g3, q4, i, false);            // {-, mask4, D4, xhat3}
                tabA[i] = make_float2(fmaf(e.w, nC, nB), e.y != 0.f ? A : 0.f);
                const float n1 = gln_apply(g1, actqf_decode(q1, (float)i));
                const float t2 = actqf_t(q2, n1);
                const bool in2 = actqf_inside(q2, t2);
                const float c2 = actqf_unbias(actqf_biased(q2, t2));
                const float D2 = in2 ? (c2 - t2) : c2;
                tabB[i] = make_float2(in2 ? D2 : D2 + MASK_OFF, actqf_decode(q2, c2));
            }
            cp_async_wait_all();      // this thread's share of the row's phase-A inputs has landed
            __syncthreads();          // X: tables + inputs visible; every warp is done with the previous row's phase B
            cta_copy8<NTH>(bufC1, p.code1 + r * p.ld, (bytesM + 7) & ~7);
            cp_async_commit();
            if (b > J.b_lo && threadIdx.x == 0) {
                const double2 v = row_partials_load<NTH>(part[(it - 1u) & 1u]);
                const double s2 = v.x, s3 = (double)xa_prev * (v.y - (double)tb_prev * v.x) + (double)xb_prev * v.x;
                atomicAdd(acc + L.samp1 + 2 * (b - 1), (double)gam1 * s2);
                atomicAdd(acc + L.samp1 + 2 * (b - 1) + 1, (double)gam1 * s3);
                dbet += s2;
                dgam += s3;
            }
            const float zlo = vreg(zthr[0]), zhi = vreg(zthr[1]);
            // ---------------- phase A ----------------
            {
                // decode3(c) = delta3*c + min3 from the float of the table address (ta + 8c)
                const float taf = (float)ta;
                const float2 dec3b = f2s(fmaf(-dec3s, taf, q3.mn)), tafv = f2s(-taf);
                auto body = [&](int v, auto tail_tag) {
                    constexpr bool TAIL = decltype(tail_tag)::value;
                    const float4 y = lds128(bufY + 16u * v);
                    const uint2 gw = lds64u(bufG + 8u * v);
                    const uint32_t cw = lds32u(bufC3 + 4u * v);
                    const float2 yy[2] = {lo2(y), hi2(y)};
                    float2 o[2];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint32_t w = j ? gw.y : gw.x;
                        const uint32_t ax = tab_addr<3>(cw, 2 * j, ta), ay = tab_addr<3>(cw, 2 * j + 1, ta);
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
                        a0c = __ffma2_rn(u, __fadd2_rn(cf, tafv), a0c);       // u * 8c (the table base changes with the row parity)
                        const float2 dq = __ffma2_rn(cf, dec3a, dec3b);      // decode3(code)
                        a0 = __ffma2_rn(gz, __fadd2_rn(dq, neg2(z)), a0);
                        o[j] = __fmul2_rn(gz, make_float2(px ? 1.f : slope3, py ? 1.f : slope3));
                        a3 = __ffma2_rn(make_float2(fminf(yy[j].x, 0.f), fminf(yy[j].y, 0.f)), gz, a3);
                    }
                    *reinterpret_cast<float4*>(row + 4 * v) = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);
                };
                int v = threadIdx.x;
                for (; v + NTH < nfull; v += 2 * NTH) {
                    body(v, std::false_type{});
                    body(v + NTH, std::false_type{});
                }
                if (v < nfull) body(v, std::false_type{});
                if ((M & 3) && (int)threadIdx.x == (nfull % NTH)) body(nfull, std::true_type{});
            }
            cp_async_wait_all();      // code1 of this row
            __syncthreads();          // Y: g_y3 row + code1 visible; every warp is done reading bufA
            if (r_next >= 0) issueA(r_next);
            // ---------------- phase B ----------------
            float2 s2 = f2s(0.f), scf = f2s(0.f);         // per row: sum gn1, sum gn1 * (tbB + 8 c1)
            {
                uint2* gn1o = reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.g_hid_a) + r * p.ld);
                auto body = [&](int v, auto tail_tag) {
                    constexpr bool TAIL = decltype(tail_tag)::value;
                    const uint32_t cw = lds32u(bufC1 + 4u * v);
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
                int v = threadIdx.x;
                for (; v + NTH < nfull; v += 2 * NTH) {
                    body(v, std::false_type{});
                    body(v + NTH, std::false_type{});
                }
                if (v < nfull) body(v, std::false_type{});
                if ((M & 3) && (int)threadIdx.x == (nfull % NTH)) body(nfull, std::true_type{});
            }
            // xhat1 = (delta1*c + min1 - mu1) * rstd1 = xa * (8c) + xb
            xa_prev = 0.125f * q1.delta * g1.rstd;
            xb_prev = (q1.mn - g1.mu) * g1.rstd;
            tb_prev = (float)tbB;
            row_partials_store<NTH>(hsum(s2), hsum(scf), part[it & 1u]);
        }
        // per-channel flush of this job
        __syncthreads();
        if (threadIdx.x == 0) {
            const double2 v = row_partials_load<NTH>(part[(it - 1u) & 1u]);
            const double s2 = v.x, s3 = (double)xa_prev * (v.y - (double)tb_prev * v.x) + (double)xb_prev * v.x;
            atomicAdd(acc + L.samp1 + 2 * (J.b_hi - 1), (double)gam1 * s2);
            atomicAdd(acc + L.samp1 + 2 * (J.b_hi - 1) + 1, (double)gam1 * s3);
            atomicAdd(acc + L.gln1 + 2 * c, dbet + s2);
            atomicAdd(acc + L.gln1 + 2 * c + 1, dgam + s3);
        }
        const float sd[4] = {hsum(d0), hsum(d1), hsum(d2), hsum(d3)};
        double vd[4];
        block_sum_fd<4>(sd, vd, sh);
        if (threadIdx.x == 0) {
            atomicAdd(acc + L.dwdw + 3 * c, vd[0]);
            atomicAdd(acc + L.dwdw + 3 * c + 1, vd[1]);
            atomicAdd(acc + L.dwdw + 3 * c + 2, vd[2]);
            atomicAdd(acc + L.dbdw + c, vd[3]);
        }
    }
    cp_async_wait_all();
    const float s[7] = {hsum(a0), hsum(a0c), hsum(a1), hsum(a3), b0x, b0y, hsum(b1)};
    double v[7];
    block_sum_fd<7>(s, v, sh);
    if (threadIdx.x == 0) {
        // sD(q3) = sum_in gz*(c - t) + 255 * sum_above (ga3 - gz);  sum u*(8c) = 8 * 255 * sum_above u
        atomicAdd(acc + L.q + 2 * Q3, v[0] * (double)q3.inv + 0.125 * v[1]);
        atomicAdd(acc + L.q + 2 * Q3 + 1, v[2]);
        atomicAdd(acc + L.slope + 1, v[3]);
        atomicAdd(acc + L.q + 2 * Q2, v[4] + v[5] - (double)MASK_OFF * v[6]);
        atomicAdd(acc + L.q + 2 * Q2 + 1, v[6]);
    }
}

}  // namespace fqss
