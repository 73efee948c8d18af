// export.cu -- the export-time quantisers of the reference (qat_quant.py:15-72): once training is over the learned ranges
// are turned into (scale, zero-point) pairs and the model is evaluated through torch.fake_quantize_per_{tensor,channel}_
// affine.  These kernels restate that ATen arithmetic on sm_100a (bit-exact against the reference classes on CPU, see
// oracle/fqss_oracle_export.py):
//
//     inv = 1.0f / scale ;  q = nearbyint(x * inv) + zero_point ;  y = (clamp(q, qmin, qmax) - zero_point) * scale
//
// plus the straight-through mask (qmin <= q <= qmax) of the op's backward, and the integer codes themselves (int32: the
// clamped q), which is what a deployment toolchain consumes.  HBM-bound elementwise passes: 128-bit accesses on the
// per-tensor form, grid = a multiple of the SM count.
#include "fqss_common.cuh"

namespace fqss {

int num_sms();

struct AffQ { float scale, inv, zp; float qmin, qmax; };

__device__ __forceinline__ float affq_code(const AffQ& q, float x, bool* in) {
    // nearbyint(x * inv) + zp evaluated in fp32 is exact: |rint| < 2^24 for every value that can land inside [qmin, qmax],
    // and values beyond clamp identically; the comparison against the range happens BEFORE the clamp (ATen's mask)
    float v = __fadd_rn(rintf(__fmul_rn(x, q.inv)), q.zp);
    *in = v >= q.qmin && v <= q.qmax;
    return fminf(fmaxf(v, q.qmin), q.qmax);
}
__device__ __forceinline__ float affq_decode(const AffQ& q, float c) { return __fmul_rn(__fsub_rn(c, q.zp), q.scale); }

constexpr int EX_THREADS = 256;

// per-tensor: x, y [n]; optional mask (u8) and codes (i32)
__global__ void __launch_bounds__(EX_THREADS) fq_affine_tensor_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                     uint8_t* __restrict__ mask, int32_t* __restrict__ code, int64_t n,
                                                                     float scale, int zero_point, int qmin, int qmax) {
    AffQ q;
    q.scale = scale; q.inv = __fdiv_rn(1.0f, scale); q.zp = (float)zero_point; q.qmin = (float)qmin; q.qmax = (float)qmax;
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * EX_THREADS;
    for (int64_t i = (int64_t)blockIdx.x * EX_THREADS + threadIdx.x; i < n4; i += stride) {
        const float4 v = ldg4_stream(x + 4 * i);
        bool m0, m1, m2, m3;
        const float c0 = affq_code(q, v.x, &m0), c1 = affq_code(q, v.y, &m1), c2 = affq_code(q, v.z, &m2), c3 = affq_code(q, v.w, &m3);
        stg4(y + 4 * i, make_float4(affq_decode(q, c0), affq_decode(q, c1), affq_decode(q, c2), affq_decode(q, c3)));
        if (mask) *reinterpret_cast<uchar4*>(mask + 4 * i) = make_uchar4(m0, m1, m2, m3);
        if (code) *reinterpret_cast<int4*>(code + 4 * i) = make_int4((int)c0, (int)c1, (int)c2, (int)c3);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t j = (n4 << 2) + threadIdx.x;
        bool m;
        const float c = affq_code(q, x[j], &m);
        y[j] = affq_decode(q, c);
        if (mask) mask[j] = m;
        if (code) code[j] = (int)c;
    }
}

// per-channel: x viewed as [outer][ch][inner]; scales[ch], zero-points 0 (TorchWeightFakeQuantize: symmetric)
__global__ void __launch_bounds__(EX_THREADS) fq_affine_channel_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                      uint8_t* __restrict__ mask, int32_t* __restrict__ code, int outer,
                                                                      int ch, int inner, const float* __restrict__ scales, int qmin,
                                                                      int qmax) {
    const int64_t n = (int64_t)outer * ch * inner;
    for (int64_t i = (int64_t)blockIdx.x * EX_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * EX_THREADS) {
        const int c = (int)((i / inner) % ch);
        AffQ q;
        q.scale = __ldg(scales + c); q.inv = __fdiv_rn(1.0f, q.scale); q.zp = 0.f; q.qmin = (float)qmin; q.qmax = (float)qmax;
        bool m;
        const float v = affq_code(q, x[i], &m);
        y[i] = affq_decode(q, v);
        if (mask) mask[i] = m;
        if (code) code[i] = (int)v;
    }
}

__global__ void __launch_bounds__(EX_THREADS) mask_mul_kernel(const float* __restrict__ g, const uint8_t* __restrict__ mask,
                                                             float* __restrict__ gx, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * EX_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * EX_THREADS)
        gx[i] = mask[i] ? g[i] : 0.f;
}

static int ex_grid(int64_t work) {
    int64_t g = (work + EX_THREADS - 1) / EX_THREADS;
    const int64_t cap = (int64_t)num_sms() * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace fqss

using namespace fqss;

extern "C" {

int fqss_fq_affine_tensor(const float* x, float* y, uint8_t* mask, int32_t* code, int64_t n, float scale, int zero_point, int qmin,
                          int qmax, void* stream) {
    FQSS_REQUIRE(x && y && n >= 0, -1, "fq_affine_tensor: null argument");
    FQSS_REQUIRE(scale > 0.f && qmin <= qmax, -1, "fq_affine_tensor: scale must be positive and qmin <= qmax");
    // same refusal as ATen (the reference's export of a range that does not contain zero fails exactly here)
    FQSS_REQUIRE(zero_point >= qmin && zero_point <= qmax, -1, "`zero_point` must be between `quant_min` and `quant_max`.");
    FQSS_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && (!code || ((uintptr_t)code & 15) == 0) &&
                     (!mask || ((uintptr_t)mask & 3) == 0), -2, "fq_affine_tensor: buffers must be 16-byte aligned");
    if (n == 0) return 0;
    FQSS_PROF("fq_affine", (cudaStream_t)stream);
    fq_affine_tensor_kernel<<<ex_grid(n >> 2), EX_THREADS, 0, (cudaStream_t)stream>>>(x, y, mask, code, n, scale, zero_point, qmin, qmax);
    return check_launch("fq_affine_tensor");
}

int fqss_fq_affine_channel(const float* x, float* y, uint8_t* mask, int32_t* code, int outer, int ch, int inner, const float* scales,
                           int qmin, int qmax, void* stream) {
    FQSS_REQUIRE(x && y && scales && outer > 0 && ch > 0 && inner > 0 && qmin <= qmax, -1, "fq_affine_channel: bad argument");
    FQSS_PROF("fq_affine", (cudaStream_t)stream);
    fq_affine_channel_kernel<<<ex_grid((int64_t)outer * ch * inner), EX_THREADS, 0, (cudaStream_t)stream>>>(x, y, mask, code, outer, ch,
                                                                                                          inner, scales, qmin, qmax);
    return check_launch("fq_affine_channel");
}

int fqss_fq_affine_bwd(const float* g, const uint8_t* mask, float* gx, int64_t n, void* stream) {
    FQSS_REQUIRE(g && mask && gx && n >= 0, -1, "fq_affine_bwd: null argument");
    if (n == 0) return 0;
    FQSS_PROF("fq_affine_bwd", (cudaStream_t)stream);
    mask_mul_kernel<<<ex_grid(n), EX_THREADS, 0, (cudaStream_t)stream>>>(g, mask, gx, n);
    return check_launch("fq_affine_bwd");
}

}  // extern "C"
