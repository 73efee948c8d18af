// attention.cu -- the attention core of MultiheadAttentionQ (reference qat_layers.py:926-939) for the small heads of the
// dual-path separators (DPTNetQ: 4 heads of 16, chunks of 250 frames; SepformerQ: 8 heads of 32): between the quantiser of
// q / sqrt(d) and the quantiser of the head outputs the reference computes  softmax(q k^T) v  with three library calls that
// materialise the [batch*heads, L, L] score tensor twice (hundreds of MB per layer).  Here it is one pass per direction:
//
//   forward   one thread per query row, K and V of the (batch, head) staged once in shared memory and read as warp-wide
//             broadcasts; two sweeps over the keys (row maximum, then exp / sum / weighted V) -- the scores never leave the
//             registers.  Saves the row log-sum-exp for backward.
//   backward  p_ij is recomputed from the log-sum-exp.  dQ: thread per query (same sweep as forward); dK, dV: thread per key
//             with Q, dO, lse, delta = <dO_i, O_i> of the head in shared memory -- every output row has one owner, no atomics,
//             bit-reproducible.
//
// fp32 throughout (FMA accumulation), exp through expf.  Shapes: q, o, dO, dq [BH][Lq][HD]; k, v, dk, dv [BH][Lk][HD], contiguous.
#include "fqss_common.cuh"

namespace fqss {

constexpr int AT_THREADS = 128;

// Rows live in registers as float2 pairs: sm_100's packed FFMA2 retires two fp32 FMAs per issue slot, and these kernels are
// issue-bound (per (query, key) pair: 3 * HD FMAs against HD * 3 / 4 broadcast LDS.128).
template <int HD>
struct Row {
    float2 p[HD / 2];
};
template <int HD>
__device__ __forceinline__ void load_row(Row<HD>& r, const float* __restrict__ src) {
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
        const float4 t = *reinterpret_cast<const float4*>(src + d);
        r.p[d / 2] = make_float2(t.x, t.y);
        r.p[d / 2 + 1] = make_float2(t.z, t.w);
    }
}
template <int HD>
__device__ __forceinline__ void zero_row(Row<HD>& r) {
#pragma unroll
    for (int d = 0; d < HD / 2; ++d) r.p[d] = make_float2(0.f, 0.f);
}
template <int HD>
__device__ __forceinline__ float dot_row(const Row<HD>& a, const float* __restrict__ s) {
    float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);      // two chains
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
        const float4 t = *reinterpret_cast<const float4*>(s + d);
        acc0 = __ffma2_rn(a.p[d / 2], make_float2(t.x, t.y), acc0);
        acc1 = __ffma2_rn(a.p[d / 2 + 1], make_float2(t.z, t.w), acc1);
    }
    return (acc0.x + acc1.x) + (acc0.y + acc1.y);
}
template <int HD>
__device__ __forceinline__ void axpy_row(Row<HD>& y, float a, const float* __restrict__ s) {
    const float2 aa = make_float2(a, a);
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
        const float4 t = *reinterpret_cast<const float4*>(s + d);
        y.p[d / 2] = __ffma2_rn(aa, make_float2(t.x, t.y), y.p[d / 2]);
        y.p[d / 2 + 1] = __ffma2_rn(aa, make_float2(t.z, t.w), y.p[d / 2 + 1]);
    }
}
template <int HD>
__device__ __forceinline__ void store_row(float* __restrict__ dst, const Row<HD>& r, float scale) {
#pragma unroll
    for (int d = 0; d < HD; d += 4)
        *reinterpret_cast<float4*>(dst + d) = make_float4(r.p[d / 2].x * scale, r.p[d / 2].y * scale, r.p[d / 2 + 1].x * scale, r.p[d / 2 + 1].y * scale);
}
__device__ __forceinline__ void stage(float* dst, const float* __restrict__ src, int n) {      // n floats, n % 4 == 0, 16-byte aligned
    for (int i = threadIdx.x * 4; i < n; i += blockDim.x * 4) *reinterpret_cast<float4*>(dst + i) = __ldg(reinterpret_cast<const float4*>(src + i));
}

// grid (ceil(Lq / 128), BH); dynamic smem: K [Lk][HD], V [Lk][HD]
template <int HD>
__global__ void __launch_bounds__(AT_THREADS) attn_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                                              float* __restrict__ o, float* __restrict__ lse, int Lq, int Lk) {
    extern __shared__ __align__(16) float sm[];
    float* Ks = sm;
    float* Vs = sm + (size_t)Lk * HD;
    const int bh = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    stage(Ks, k + (size_t)bh * Lk * HD, Lk * HD);
    stage(Vs, v + (size_t)bh * Lk * HD, Lk * HD);
    __syncthreads();
    if (i >= Lq) return;
    Row<HD> qi, acc;
    load_row<HD>(qi, q + ((size_t)bh * Lq + i) * HD);
    float m = -3.0e38f;
    for (int j = 0; j < Lk; ++j) m = fmaxf(m, dot_row<HD>(qi, Ks + j * HD));
    float l = 0.f;
    zero_row<HD>(acc);
    for (int j = 0; j < Lk; ++j) {
        const float p = expf(dot_row<HD>(qi, Ks + j * HD) - m);
        l += p;
        axpy_row<HD>(acc, p, Vs + j * HD);
    }
    store_row<HD>(o + ((size_t)bh * Lq + i) * HD, acc, 1.f / l);
    lse[(size_t)bh * Lq + i] = m + logf(l);
}

// dQ and delta: grid (ceil(Lq / 128), BH); dynamic smem: K, V
template <int HD>
__global__ void __launch_bounds__(AT_THREADS) attn_bwd_dq_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                                                 const float* __restrict__ o, const float* __restrict__ dO,
                                                                 const float* __restrict__ lse, float* __restrict__ dq, float* __restrict__ delta,
                                                                 int Lq, int Lk) {
    extern __shared__ __align__(16) float sm[];
    float* Ks = sm;
    float* Vs = sm + (size_t)Lk * HD;
    const int bh = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    stage(Ks, k + (size_t)bh * Lk * HD, Lk * HD);
    stage(Vs, v + (size_t)bh * Lk * HD, Lk * HD);
    __syncthreads();
    if (i >= Lq) return;
    const size_t row = (size_t)bh * Lq + i;
    Row<HD> qi, gi, acc;
    load_row<HD>(qi, q + row * HD);
    load_row<HD>(gi, dO + row * HD);
    const float dl = dot_row<HD>(gi, o + row * HD);
    const float ls = lse[row];
    zero_row<HD>(acc);
    for (int j = 0; j < Lk; ++j) {
        const float p = expf(dot_row<HD>(qi, Ks + j * HD) - ls);
        const float ds = p * (dot_row<HD>(gi, Vs + j * HD) - dl);
        axpy_row<HD>(acc, ds, Ks + j * HD);
    }
    store_row<HD>(dq + row * HD, acc, 1.f);
    delta[row] = dl;
}

// dK and dV: grid (ceil(Lk / 128), BH); dynamic smem: Q [Lq][HD], dO [Lq][HD], lse [Lq], delta [Lq]
template <int HD>
__global__ void __launch_bounds__(AT_THREADS) attn_bwd_dkv_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                                                  const float* __restrict__ dO, const float* __restrict__ lse,
                                                                  const float* __restrict__ delta, float* __restrict__ dk, float* __restrict__ dv,
                                                                  int Lq, int Lk, int Lq4) {
    extern __shared__ __align__(16) float sm[];
    float* Qs = sm;
    float* Gs = Qs + (size_t)Lq * HD;
    float* Ls = Gs + (size_t)Lq * HD;
    float* Ds = Ls + Lq4;
    const int bh = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    stage(Qs, q + (size_t)bh * Lq * HD, Lq * HD);
    stage(Gs, dO + (size_t)bh * Lq * HD, Lq * HD);
    for (int i = threadIdx.x; i < Lq; i += blockDim.x) {
        Ls[i] = __ldg(lse + (size_t)bh * Lq + i);
        Ds[i] = __ldg(delta + (size_t)bh * Lq + i);
    }
    __syncthreads();
    if (j >= Lk) return;
    const size_t row = (size_t)bh * Lk + j;
    Row<HD> kj, vj, ak, av;
    load_row<HD>(kj, k + row * HD);
    load_row<HD>(vj, v + row * HD);
    zero_row<HD>(ak);
    zero_row<HD>(av);
    for (int i = 0; i < Lq; ++i) {
        const float p = expf(dot_row<HD>(kj, Qs + i * HD) - Ls[i]);
        axpy_row<HD>(av, p, Gs + i * HD);
        const float ds = p * (dot_row<HD>(vj, Gs + i * HD) - Ds[i]);
        axpy_row<HD>(ak, ds, Qs + i * HD);
    }
    store_row<HD>(dk + row * HD, ak, 1.f);
    store_row<HD>(dv + row * HD, av, 1.f);
}

constexpr size_t AT_MAX_SMEM = 200 * 1024;

// CTA width for L rows: the narrowest of {32, 64, 128} threads that wastes the fewest lanes (short inter-chunk sequences)
static int at_threads(int L) {
    int best = 128, waste = ((L + 127) / 128) * 128 - L;
    for (int t = 64; t >= 32; t >>= 1) {
        const int w = ((L + t - 1) / t) * t - L;
        if (w < waste) { waste = w; best = t; }
    }
    return best;
}

template <typename F>
static int set_smem(F fn, size_t bytes) {
    if (bytes <= 48 * 1024) return 0;
    return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AT_MAX_SMEM) == cudaSuccess ? 0 : -4;
}

}  // namespace fqss

using namespace fqss;

extern "C" {

// largest of the three kernels' shared-memory needs; > 200 KB: no kernel (the caller keeps the library path)
size_t fqss_attn_smem_bytes(int Lq, int Lk, int HD) {
    const size_t a = (size_t)2 * Lk * HD * sizeof(float);
    const size_t b = ((size_t)2 * Lq * HD + 2 * (size_t)((Lq + 3) & ~3)) * sizeof(float);
    return a > b ? a : b;
}

int fqss_attn_fwd(const float* q, const float* k, const float* v, float* o, float* lse, int BH, int Lq, int Lk, int HD, void* stream) {
    FQSS_REQUIRE(q && k && v && o && lse && BH > 0 && BH < 65536 && Lq > 0 && Lk > 0, -1, "attn_fwd: bad argument");
    FQSS_REQUIRE(HD == 8 || HD == 16 || HD == 32, -1, "attn_fwd: head size %d has no kernel (8, 16, 32)", HD);
    FQSS_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o), -2, "attn_fwd: operands must be 16-byte aligned");
    const size_t smem = (size_t)2 * Lk * HD * sizeof(float);
    FQSS_REQUIRE(smem <= AT_MAX_SMEM, -1, "attn_fwd: K and V of one head (%zu B) do not fit shared memory", smem);
    cudaStream_t s = (cudaStream_t)stream;
    const int th = at_threads(Lq);
    const dim3 grid((Lq + th - 1) / th, BH);
    FQSS_PROF("attn_fwd", s);
#define FQSS_AT_F(D_)                                                                  \
    do {                                                                               \
        if (set_smem(attn_fwd_kernel<D_>, smem)) { set_error("attn_fwd: cannot set shared memory"); return -4; } \
        attn_fwd_kernel<D_><<<grid, th, smem, s>>>(q, k, v, o, lse, Lq, Lk);   \
    } while (0)
    if (HD == 8) FQSS_AT_F(8); else if (HD == 16) FQSS_AT_F(16); else FQSS_AT_F(32);
#undef FQSS_AT_F
    return check_launch("attn_fwd");
}

int fqss_attn_bwd(const float* q, const float* k, const float* v, const float* o, const float* dO, const float* lse, float* dq, float* dk,
                  float* dv, float* delta_ws, int BH, int Lq, int Lk, int HD, void* stream) {
    FQSS_REQUIRE(q && k && v && o && dO && lse && dq && dk && dv && delta_ws && BH > 0 && BH < 65536 && Lq > 0 && Lk > 0, -1, "attn_bwd: bad argument");
    FQSS_REQUIRE(HD == 8 || HD == 16 || HD == 32, -1, "attn_bwd: head size %d has no kernel (8, 16, 32)", HD);
    FQSS_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o) && aligned16(dO) && aligned16(dq) && aligned16(dk) && aligned16(dv), -2,
                 "attn_bwd: operands must be 16-byte aligned");
    const size_t smem1 = (size_t)2 * Lk * HD * sizeof(float);
    const int Lq4 = (Lq + 3) & ~3;
    const size_t smem2 = ((size_t)2 * Lq * HD + 2 * (size_t)Lq4) * sizeof(float);
    FQSS_REQUIRE(smem1 <= AT_MAX_SMEM && smem2 <= AT_MAX_SMEM, -1, "attn_bwd: one head does not fit shared memory");
    cudaStream_t s = (cudaStream_t)stream;
    const int t1 = at_threads(Lq), t2 = at_threads(Lk);
    const dim3 g1((Lq + t1 - 1) / t1, BH), g2((Lk + t2 - 1) / t2, BH);
    FQSS_PROFN("attn_bwd", s, 2);
#define FQSS_AT_B(D_)                                                                                              \
    do {                                                                                                           \
        if (set_smem(attn_bwd_dq_kernel<D_>, smem1) || set_smem(attn_bwd_dkv_kernel<D_>, smem2)) { set_error("attn_bwd: cannot set shared memory"); return -4; } \
        attn_bwd_dq_kernel<D_><<<g1, t1, smem1, s>>>(q, k, v, o, dO, lse, dq, delta_ws, Lq, Lk);           \
        attn_bwd_dkv_kernel<D_><<<g2, t2, smem2, s>>>(q, k, v, dO, lse, delta_ws, dk, dv, Lq, Lk, Lq4);    \
    } while (0)
    if (HD == 8) FQSS_AT_B(8); else if (HD == 16) FQSS_AT_B(16); else FQSS_AT_B(32);
#undef FQSS_AT_B
    return check_launch("attn_bwd");
}

}  // extern "C"
