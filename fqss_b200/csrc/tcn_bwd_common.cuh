// tcn_bwd_common.cuh -- accumulator layout, row-loop macros and small vector helpers shared by the backward kernels
// of the fused ConvBlock (tcn_bwd.cu, tcn_rows.cuh).
#pragma once
#include <type_traits>

#include "fqss_common.cuh"
#include "tcn_common.cuh"

namespace fqss {

int num_sms();

// fp64 accumulator block inside the workspace
struct AccLayout {
    // Sums every row CTA of a launch adds to (the range sums of the eight activation quantisers, the two PReLU slopes) are
    // spread over NSLOT copies, one per 256-byte line, chosen by the CTA's channel: 16 K fp64 atomics of one launch on ONE
    // address serialise in L2 (measured on B200: 18 us of the 71 us gLN2 sums kernel, and the backlog slows the next kernel).
    // The finalise kernel adds the copies in slot order.
    static constexpr int NSLOT = 32, SLOT_STRIDE = 32, SLOPE_OFF = 16;
    // The per-sample gLN sums {S1, S2} of the lean row kernels are split the same way over NSS copies chosen by the channel
    // (sslot1 / sslot2; the 512 row CTAs of a sample run back to back, so their atomics queue on one sector).  The readers --
    // every CTA of the next row kernel, at its start -- add samp + all copies (samp_get), so NSS stays small: measured on one
    // B200 in one session (scratch/gpu_ab.sh), gLN2 sums / fused gLN2+depthwise / gLN1 kernels: 1 copy 55.4 / 166.4 / 95.6 us,
    // 2 copies 48.9 / 158.9 / 94.9 us, 8 copies 49.0 / 169.0 / 98.0 us.  The non-lean path's reduce kernel writes samp1 / samp2
    // and leaves the copies zero.
    static constexpr int NSS = 2;
    int64_t glob, db1, db2, dbdw, dwdw, row1, row2, samp1, samp2, sslot1, sslot2, gln1, gln2, total;
    __host__ __device__ int64_t ss(int64_t base, int b, int slot) const { return base + ((int64_t)b * NSS + (slot & (NSS - 1))) * 4; }
    __host__ __device__ int64_t qs(int slot) const { return glob + (int64_t)(slot & (NSLOT - 1)) * SLOT_STRIDE; }
    __host__ __device__ AccLayout(int B, int Cio, int Chid) {
        int64_t o = 0;
        glob = o; o += (int64_t)NSLOT * SLOT_STRIDE;      // per slot: [16 range sums | 2 slope sums | pad]
        db1 = o; o += Chid;
        db2 = o; o += 2 * Cio;
        dbdw = o; o += Chid;
        dwdw = o; o += 3 * Chid;
        row1 = o; o += 2 * (int64_t)B * Chid;
        row2 = o; o += 2 * (int64_t)B * Chid;
        samp1 = o; o += 2 * B;
        samp2 = o; o += 2 * B;
        o = (o + 3) & ~(int64_t)3;                   // 32-byte aligned: the readers use 16-byte loads
        sslot1 = o; o += (int64_t)B * NSS * 4;      // per (sample, copy): {S1, S2, pad, pad} -- one 32-byte sector each
        sslot2 = o; o += (int64_t)B * NSS * 4;
        gln1 = o; o += 2 * Chid;          // {dbeta, dgamma} per channel, accumulated by the quantised row kernels (no reduce launch)
        gln2 = o; o += 2 * Chid;
        total = o;
    }
};

// per-sample gLN sums of sample b: the reduce kernel's entry plus the NSS atomic copies (independent loads, one round trip)
__device__ __forceinline__ void samp_get(const double* acc, const AccLayout& L, int which, int b, double& s1, double& s2) {
    const int64_t base = which == 1 ? L.samp1 : L.samp2, sb = which == 1 ? L.sslot1 : L.sslot2;
    double a = acc[base + 2 * b], c = acc[base + 2 * b + 1];
    double2 v[AccLayout::NSS];
#pragma unroll
    for (int k = 0; k < AccLayout::NSS; ++k) v[k] = *reinterpret_cast<const double2*>(acc + L.ss(sb, b, k));
#pragma unroll
    for (int k = 0; k < AccLayout::NSS; ++k) { a += v[k].x; c += v[k].y; }
    s1 = a;
    s2 = c;
}

enum { Q1 = 0, Q2 = 1, Q3 = 2, Q4 = 3, QRES = 4, QSKIP = 5, QADD = 6, QADDS = 7 };

__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }

// ---------------------------------------------------------------------------------------------
// Row kernels.  One CTA per (sample, channel) row, every thread handles 4 consecutive frames per trip
// (128-bit fp32 / 64-bit bf16 accesses), per-row constants and the code-indexed tables are built once
// per CTA, partial sums are reduced fp32 -> warp -> fp64.  Frames m >= M (row padding up to ld) carry
// no gradient: inputs are masked on load, outputs there are written as zeros.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_bf16x4(__nv_bfloat16* p, float a, float b, float c, float d) {
    *reinterpret_cast<uint2*>(p) = float4_to_bf16x4(a, b, c, d);
}
__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float f4_get(const float4& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }

__device__ __forceinline__ float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4& v) { return make_float2(v.z, v.w); }
__device__ __forceinline__ float2 neg2(float2 v) { return make_float2(-v.x, -v.y); }
__device__ __forceinline__ float hsum(float2 v) { return v.x + v.y; }
// zero the components of a frame quad that lie at or beyond M (only the last quad of a row is affected)
__device__ __forceinline__ void mask_tail(float2& a01, float2& a23, int nval) {
    if (nval < 4) {
        a23.y = 0.f;
        if (nval < 3) a23.x = 0.f;
        if (nval < 2) a01.y = 0.f;
        if (nval < 1) a01.x = 0.f;
    }
}

// Row loops run mask-free over the full frame quads (v < M/4); the one ragged quad of a row (M % 4 frames) is
// handled once, by one thread, through the same body with TAIL = true.  Quads that are entirely padding
// (ld - M >= 4) are never read or written.
#define FQSS_ROW_LOOPN(NTH_, body, M)                                                                           \
    do {                                                                                                  \
        const int nfull_ = (M) >> 2;                                                                      \
        for (int v_ = threadIdx.x; v_ < nfull_; v_ += (NTH_)) body(v_, std::false_type{});                \
        if (((M)&3) && (int)threadIdx.x == (nfull_ % (NTH_))) body(nfull_, std::true_type{});             \
    } while (0)

// Batched variant: every thread first issues the loads of NQ quads (`load(v)` returns a plain struct of raw words), then
// consumes them.  One row is only ~1000 quads, i.e. a handful of trips per thread, so without the batch every trip
// exposes a full DRAM latency (ncu: > 40 % of the stall samples on the first use of the loaded word) and the bytes in
// flight per SM stay far below what HBM needs.
#define FQSS_ROW_LOOP_BATCH(NTH_, NQ_, load, body, M)                                                     \
    do {                                                                                                  \
        const int nfull_ = (M) >> 2;                                                                      \
        for (int base_ = threadIdx.x; base_ < nfull_; base_ += (NQ_) * (NTH_)) {                          \
            decltype(load(0)) d_[NQ_];                                                                    \
            _Pragma("unroll") for (int k_ = 0; k_ < (NQ_); ++k_) {                                        \
                const int v_ = base_ + k_ * (NTH_);                                                       \
                if (v_ < nfull_) d_[k_] = load(v_);                                                       \
            }                                                                                             \
            _Pragma("unroll") for (int k_ = 0; k_ < (NQ_); ++k_) {                                        \
                const int v_ = base_ + k_ * (NTH_);                                                       \
                if (v_ < nfull_) body(v_, d_[k_], std::false_type{});                                     \
            }                                                                                             \
        }                                                                                                 \
        if (((M)&3) && (int)threadIdx.x == (nfull_ % (NTH_))) body(nfull_, load(nfull_), std::true_type{}); \
    } while (0)

}  // namespace fqss
