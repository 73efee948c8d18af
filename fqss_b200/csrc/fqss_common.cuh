// fqss_b200 -- common device helpers for the sm_100a kernels.
//
// Arithmetic contract (verified against qat_quant.py:88-147):
// every fake-quant op is a separately rounded fp32 operation (PyTorch eager never contracts to
// FMA), so the helpers below use the explicit round-to-nearest intrinsics; a plain `a*b+c` would be
// fused by nvcc and flip codes that sit on a rounding boundary.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/fqss.h"

namespace fqss {

// ---------------------------------------------------------------------------------------------
// host side: error convention
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define FQSS_REQUIRE(cond, code, ...)          \
    do {                                       \
        if (!(cond)) {                         \
            ::fqss::set_error(__VA_ARGS__);    \
            return (code);                     \
        }                                      \
    } while (0)

// Launch accounting + optional CUDA-event timing on the launching stream (prof.cu).  Every kernel
// launch site sits inside one: FQSS_PROF("kernel_class", stream) or FQSS_PROFN(name, stream, n_kernels).
class ProfScope {
public:
    ProfScope(const char* name, cudaStream_t s, int nkernels = 1);
    ~ProfScope();
    ProfScope(const ProfScope&) = delete;
    ProfScope& operator=(const ProfScope&) = delete;

private:
    cudaStream_t stream_;
    void* e1_;
    int slot_;
};
#define FQSS_PROF_CAT2(a, b) a##b
#define FQSS_PROF_CAT(a, b) FQSS_PROF_CAT2(a, b)
#define FQSS_PROF(name, stream) ::fqss::ProfScope FQSS_PROF_CAT(fqss_prof_, __LINE__)(name, (cudaStream_t)(stream))
#define FQSS_PROFN(name, stream, n) ::fqss::ProfScope FQSS_PROF_CAT(fqss_prof_, __LINE__)(name, (cudaStream_t)(stream), (n))

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------------------------------------
// quantiser parameter blocks, computed once per thread from the device-resident range tensors
// ---------------------------------------------------------------------------------------------
struct ActQ {
    float mn;      // zero point = min_range
    float delta;   // (max - min) / levels
    float levels;  // 2^bits - 1
};

__device__ __forceinline__ ActQ load_actq(const float* __restrict__ rmin, const float* __restrict__ rmax, int n_bits) {
    ActQ q;
    q.mn = __ldg(rmin);
    q.levels = (float)((1 << n_bits) - 1);
    q.delta = __fdiv_rn(__fsub_rn(__ldg(rmax), q.mn), q.levels);   // qat_quant.py:140
    return q;
}

// t = (x - min) / delta, X = rint(t)  (qat_quant.py:144; round_ste value == rint, half-to-even)
__device__ __forceinline__ float actq_t(const ActQ& q, float x) { return __fdiv_rn(__fsub_rn(x, q.mn), q.delta); }

__device__ __forceinline__ float actq_code(const ActQ& q, float x) {
    float X = rintf(actq_t(q, x));
    return fminf(fmaxf(X, 0.f), q.levels);
}

// y = delta * code + min, two roundings (qat_quant.py:145)
__device__ __forceinline__ float actq_decode(const ActQ& q, float code) {
    return __fadd_rn(__fmul_rn(q.delta, code), q.mn);
}

__device__ __forceinline__ float actq_fq(const ActQ& q, float x) { return actq_decode(q, actq_code(q, x)); }

// Backward pieces for one element.  Returns the gradient w.r.t. the pre-quant value and
// accumulates the two range sums:  sD += g * (clip(X) - m*t),  sZ += g * (1 - m).
//   g_max = sD / levels ;  g_min = sZ - sD / levels         
__device__ __forceinline__ float actq_bwd(const ActQ& q, float x, float g, float& sD, float& sZ) {
    float t = actq_t(q, x);
    float X = rintf(t);
    bool in = (X >= 0.f) && (X <= q.levels);
    float c = fminf(fmaxf(X, 0.f), q.levels);
    float D = in ? __fsub_rn(X, t) : c;
    sD = fmaf(g, D, sD);
    sZ += in ? 0.f : g;
    // ((g * delta) * m) / delta : the exact op chain autograd runs (mul, clip-mask, div)
    return in ? __fdiv_rn(__fmul_rn(g, q.delta), q.delta) : 0.f;
}

// ---------------------------------------------------------------------------------------------
// Exact activation quantiser for the fused (ALU-bound) kernels -- same bits as actq_*(), ~1/3 of the
// issue slots and no MUFU / conversion-pipe instructions:
//  * IEEE quotient without a division: with inv = RN(1/delta), q = RN(d*inv), r = d - q*delta (exact
//    in one FMA), t = RN(q + r*inv) is the correctly rounded d/delta (Markstein's correction step);
//    3 FMA-pipe instructions, branch-free.
//  * rint (half-to-even) by the 1.5*2^23 magic constant on a value pre-clamped to [0, levels]:
//    rint(clamp(t)) == clamp(rint(t)).  The low byte of (clamp(t) + MAGIC) is the integer code.
//  * STE mask without rounding: 0 <= rint(t) <= levels  <=>  -0.5 <= t < levels + 0.5 (ties go to
//    even: -0.5 -> -0 is inside, levels + 0.5 -> levels + 1 is outside).
// ---------------------------------------------------------------------------------------------
constexpr float RINT_MAGIC = 12582912.f;      // 1.5 * 2^23

struct ActQF {
    float mn, delta, inv, levels;
};

__device__ __forceinline__ ActQF load_actqf(const float* __restrict__ rmin, const float* __restrict__ rmax, int n_bits) {
    ActQF q;
    q.mn = __ldg(rmin);
    q.levels = (float)((1 << n_bits) - 1);
    q.delta = __fdiv_rn(__fsub_rn(__ldg(rmax), q.mn), q.levels);
    q.inv = __fdiv_rn(1.f, q.delta);
    return q;
}

__device__ __forceinline__ float exact_div(float d, float delta, float inv) {
    const float q = __fmul_rn(d, inv);
    const float r = __fmaf_rn(-q, delta, d);
    return __fmaf_rn(r, inv, q);
}

// t = (x - min) / delta, bit-identical to actq_t()
__device__ __forceinline__ float actqf_t(const ActQF& q, float x) { return exact_div(__fsub_rn(x, q.mn), q.delta, q.inv); }

// clamp(t) + MAGIC: low mantissa bits hold the integer code
__device__ __forceinline__ float actqf_biased(const ActQF& q, float t) {
    return __fadd_rn(fminf(fmaxf(t, 0.f), q.levels), RINT_MAGIC);
}
__device__ __forceinline__ unsigned actqf_index(float biased) { return __float_as_uint(biased) & 0xFFFFu; }
__device__ __forceinline__ float actqf_unbias(float biased) { return __fsub_rn(biased, RINT_MAGIC); }

__device__ __forceinline__ float actqf_code(const ActQF& q, float x) { return actqf_unbias(actqf_biased(q, actqf_t(q, x))); }
__device__ __forceinline__ bool actqf_inside(const ActQF& q, float t) { return t >= -0.5f && t < q.levels + 0.5f; }

__device__ __forceinline__ float actqf_decode(const ActQF& q, float c) { return __fadd_rn(__fmul_rn(q.delta, c), q.mn); }
__device__ __forceinline__ float actqf_fq(const ActQF& q, float x) { return actqf_decode(q, actqf_code(q, x)); }

// statistics-only variant (never decides a stored code): reciprocal multiply, fused decode
__device__ __forceinline__ float actqf_fq_approx(const ActQF& q, float x) {
    const float t = (x - q.mn) * q.inv;
    const float c = (fminf(fmaxf(t, 0.f), q.levels) + RINT_MAGIC) - RINT_MAGIC;
    return fmaf(c, q.delta, q.mn);
}

// backward through the quantiser given t = actqf_t(q, x): returns g * mask, accumulates
// sD += g*(clip(X) - m*t), sZ += g*(1-m)   (g_max = sD/levels, g_min = sZ - sD/levels)
__device__ __forceinline__ float actqf_bwd_t(const ActQF& q, float t, float g, float& sD, float& sZ) {
    const bool in = actqf_inside(q, t);
    const float c = actqf_unbias(actqf_biased(q, t));
    sD = fmaf(g, in ? (c - t) : c, sD);
    sZ += in ? 0.f : g;
    return in ? g : 0.f;
}
__device__ __forceinline__ float actqf_bwd(const ActQF& q, float x, float g, float& sD, float& sZ) {
    return actqf_bwd_t(q, actqf_t(q, x), g, sD, sZ);
}

// ---------------------------------------------------------------------------------------------
// Packed-pair variants for the issue-bound row kernels (sm_100 FADD2 / FMUL2 / FFMA2 process two fp32
// lanes per issue slot; every op below is the same IEEE-rounded fp32 operation as its scalar twin):
//   actqf_t2   : t = (x - min) / delta for two elements, bit-identical to actqf_t()
//   code_u8    : clamp(rint(t), 0, 255) as an integer in ONE conversion (cvt.rni.sat.u8.f32, ties to even,
//                NaN -> 0) -- identical to the fmaxf/fminf/magic-constant sequence for 8-bit quantisers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 f2s(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 exact_div2(float2 d, float delta, float inv) {
    const float2 q = __fmul2_rn(d, f2s(inv));
    const float2 r = __ffma2_rn(q, f2s(-delta), d);
    return __ffma2_rn(r, f2s(inv), q);
}
__device__ __forceinline__ float2 actqf_t2(const ActQF& q, float2 x) { return exact_div2(__fadd2_rn(x, f2s(-q.mn)), q.delta, q.inv); }
__device__ __forceinline__ unsigned code_u8(float t) {
    unsigned r;
    asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(t));
    return r;
}
// integer code (0..255) -> float without the conversion pipe: OR into the mantissa of 2^23, subtract 2^23 (exact)
__device__ __forceinline__ float2 u8x2_to_float2(unsigned a, unsigned b) {
    return __fadd2_rn(make_float2(__uint_as_float(a | 0x4B000000u), __uint_as_float(b | 0x4B000000u)), f2s(-8388608.f));
}
// STE mask of an 8-bit quantiser: inside <=> -0.5 <= t < 255.5 <=> 0 <= t + 0.5 < 256 (the sum is exact wherever it
// could cross either bound; -0.5 -> +0 is inside, negatives have the sign bit set, NaN is outside): one unsigned compare
__device__ __forceinline__ bool inside_u8(float t_plus_half) { return __float_as_uint(t_plus_half) < 0x43800000u; }

struct WQ {
    float delta;
    float lo, hi;
};

__device__ __forceinline__ WQ make_wq(float rmin, float rmax, int n_bits) {
    WQ q;
    float a = fmaxf(fabsf(rmin), fabsf(rmax));
    float levels = (float)((1 << n_bits) - 1);
    q.delta = __fdiv_rn(__fmul_rn(2.f, a), levels);   // qat_quant.py:131
    q.lo = -(float)(1 << (n_bits - 1));
    q.hi = (float)((1 << (n_bits - 1)) - 1);
    return q;
}

__device__ __forceinline__ float wq_code(const WQ& q, float w) {
    float X = rintf(__fdiv_rn(w, q.delta));
    return fminf(fmaxf(X, q.lo), q.hi);
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum of NV values per thread; result valid in thread 0.  `sh` needs NV*32 doubles.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* sh) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) sh[i * 32 + wid] = v[i];
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
// 128-bit global access
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stg4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// streaming variants: data touched exactly once should not displace L2-resident operands
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

}  // namespace fqss
