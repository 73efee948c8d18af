// edge_tc.cu -- small producer / consumer kernels that put the filterbank edges (Conv1dEncoderQ qat_layers.py:993-1046,
// ConvTr1dDecoderQ :1305-1361, ResidualErrorBlock :1105-1220) on the tcgen05 GEMM of gemm_tc.cu:
//
//   decoder  y[r,t] = sum_{o,k: 8m+k=t} w[o,k] Y[r,o,m]   =  overlap-add of  frames[r,k,m] = sum_o w[o,k] Y[r,o,m]
//            -> a GEMM over the 512 filters with the 16 taps as output channels (integer codes on both sides: Y is the
//               output of an 8-bit quantiser, w is fake-quantised per tensor) followed by fqss_ola_fwd
//   encoder  y[r,o,m] = sum_{c,k} w[o,c,k] x[r,c,8m+k]    =  GEMM over the 16*C taps of the framed input
//
//   fqss_frames_split    g [R][T]  -> bf16 [R][128][ld]: rows 0..15 = hi(g[r, 8m+k]), rows 16..31 = mid, 32..47 = lo, others 0; the
//                        framed output gradient as a three-term bf16 operand (hi + mid + lo = the fp32 value: 24 mantissa bits) of the
//                        decoder's dgrad / wgrad GEMMs, plus the fp64 row sums the wgrad affine needs
//   fqss_frames_encode   x [R][C][T] on an 8-bit grid -> bf16 codes [R][KP][ld]: row c*16+k = code(x[r,c,8m+k]), others 0
//   fqss_sub_fq_codes    Y1 = FQ(Y - Yq) (RQB, qat_layers.py:1195) as bf16 codes only: the decoder GEMM's next operand
//   fqss_dec_wgrad_fold  the [128][F] result of fqss_wgrad_codes on split frames -> dWq[F][16] (hi + lo rows, transposed)
#include "fqss_common.cuh"
#include "tcn_common.cuh"

namespace fqss {

int num_sms();

constexpr int ET_THREADS = 256;

// grid (ceil(M/256), 3L or 128, R): one thread = one (row, frame)
__global__ void __launch_bounds__(ET_THREADS) frames_split_kernel(const float* __restrict__ g, int64_t ldg, __nv_bfloat16* __restrict__ out,
                                                                 int64_t ldo, int M, int L, int H, double* __restrict__ rowsum) {
    __shared__ double sh[32];
    const int64_t r = blockIdx.z;
    const int j = (int)blockIdx.y;                   // j < 128
    const int64_t row = r * 128 + j;
    const int m = blockIdx.x * ET_THREADS + threadIdx.x;
    float v = 0.f;
    if (j < 3 * L && m < M) {
        // x = hi + mid + lo, three bf16 terms: 24 mantissa bits, i.e. the fp32 value exactly (up to denormal tails)
        const int part = j / L;
        const float x = __ldg(g + r * ldg + (int64_t)m * H + (j - part * L));
        const float hi = __bfloat162float(__float2bfloat16_rn(x));
        const float mid = __bfloat162float(__float2bfloat16_rn(x - hi));
        v = part == 0 ? hi : (part == 1 ? mid : __bfloat162float(__float2bfloat16_rn((x - hi) - mid)));
    }
    if (m < M) out[row * ldo + m] = __float2bfloat16_rn(v);        // exact: v is a bf16 value
    if (rowsum && j < 3 * L) {                                     // block-uniform branch
        double s = (double)warp_sum(v);
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) sh[wid] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < ET_THREADS / 32; ++w) t += sh[w];
            atomicAdd(rowsum + j, t);
        }
    }
}

__global__ void __launch_bounds__(ET_THREADS) frames_encode_kernel(const float* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ out,
                                                                  int64_t ldo, int C, int M, int L, int H, int KP, const float* rmin,
                                                                  const float* rmax) {
    const int64_t r = blockIdx.z;
    const int j = (int)blockIdx.y;                   // j < KP
    const int64_t row = r * KP + j;
    const int m = blockIdx.x * ET_THREADS + threadIdx.x;
    if (m >= M) return;
    float c = 0.f;
    if (j < C * L) {
        const ActQF q = load_actqf(rmin, rmax, 8);
        const int ch = j / L, k = j - ch * L;
        c = actqf_code(q, __ldg(x + (r * C + ch) * ldx + (int64_t)m * H + k));
    }
    out[row * ldo + m] = __float2bfloat16_rn(c);
}

__global__ void __launch_bounds__(ET_THREADS) sub_fq_codes_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ b,
                                                                 int64_t ldb, __nv_bfloat16* __restrict__ out, int64_t ldo, int M,
                                                                 const float* rmin, const float* rmax) {
    const int64_t row = blockIdx.y;
    const ActQF q = load_actqf(rmin, rmax, 8);
    const int nq = (M + 3) >> 2;
    for (int v = blockIdx.x * ET_THREADS + threadIdx.x; v < nq; v += gridDim.x * ET_THREADS) {
        const float4 x = ldg4_stream(a + row * lda + 4 * v), y = ldg4_stream(b + row * ldb + 4 * v);
        const float c0 = actqf_code(q, __fsub_rn(x.x, y.x)), c1 = actqf_code(q, __fsub_rn(x.y, y.y));
        const float c2 = actqf_code(q, __fsub_rn(x.z, y.z)), c3 = actqf_code(q, __fsub_rn(x.w, y.w));
        *reinterpret_cast<uint2*>(out + row * ldo + 4 * v) = float4_to_bf16x4(c0, c1, c2, c3);     // pad columns (< ld): don't-care
    }
}

// part [128][F] (rows 0..L-1: hi frames, L..2L-1: mid, 2L..3L-1: lo) -> dWq [F][L]
__global__ void dec_wgrad_fold_kernel(const float* __restrict__ part, float* __restrict__ dWq, int F, int L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * L) return;
    const int o = i / L, k = i - o * L;
    dWq[i] = (part[(int64_t)k * F + o] + part[(int64_t)(L + k) * F + o]) + part[(int64_t)(2 * L + k) * F + o];
}

// Decoder weight W [F][L] (ConvTranspose1d F -> 1, per-TENSOR symmetric quantiser: ch_out_idx = 1) as GEMM operands:
//   Wc [128][F]: row k < L = code[.,k]            (forward: frames[k] = sum_o code[o,k] * act_code[o])
//   WT [F][128]: [code | code | code | 0]         (dgrad on the split [hi ; mid ; lo] framed gradient)
//   s1[k] = dw * da, s0[k] = dw * min_a * sum_o code[o,k]   (k < L; zero beyond), dgs[o] = dw
__global__ void __launch_bounds__(256) dec_prep_kernel(const float* __restrict__ W, const float* wmin, const float* wmax, const float* amin,
                                                       const float* amax, __nv_bfloat16* __restrict__ Wc, __nv_bfloat16* __restrict__ WT,
                                                       float* __restrict__ s1, float* __restrict__ s0, float* __restrict__ dgs, int F, int L) {
    // grid: ceil(F / 8) CTAs of 8 filters each; CTA 0 also produces the epilogue constants
    __shared__ float part[256];
    const WQ wq = make_wq(__ldg(wmin), __ldg(wmax), 8);
    const int o0 = blockIdx.x * 8;
    for (int i = threadIdx.x; i < 8 * 128; i += blockDim.x) {
        const int o = o0 + (i >> 7), j = i & 127;
        if (o >= F) continue;
        const int k = j % L;
        const float c = j < 3 * L ? wq_code(wq, __ldg(W + (int64_t)o * L + k)) : 0.f;
        WT[(int64_t)o * 128 + j] = __float2bfloat16_rn(c);
    }
    for (int i = threadIdx.x; i < 8 * 128; i += blockDim.x) {
        const int k = i >> 3, o = o0 + (i & 7);
        if (o >= F) continue;
        Wc[(int64_t)k * F + o] = __float2bfloat16_rn(k < L ? wq_code(wq, __ldg(W + (int64_t)o * L + k)) : 0.f);
    }
    if (threadIdx.x < 8 && o0 + threadIdx.x < F) dgs[o0 + threadIdx.x] = wq.delta;
    if (blockIdx.x != 0) return;
    const ActQF qa = load_actqf(amin, amax, 8);
    const int nparts = 256 / L;                          // L <= 64
    const int k = threadIdx.x % L, pt = threadIdx.x / L;
    float cs = 0.f;                                      // integer sums < 2^24: exact in any order
    if (pt < nparts)
        for (int o = pt; o < F; o += nparts) cs += wq_code(wq, __ldg(W + (int64_t)o * L + k));
    part[threadIdx.x] = pt < nparts ? cs : 0.f;
    __syncthreads();
    if (threadIdx.x < 128) {
        const int kk = threadIdx.x;
        float tot = 0.f;
        if (kk < L)
            for (int q = 0; q < nparts; ++q) tot += part[q * L + kk];
        s1[kk] = kk < L ? __fmul_rn(wq.delta, qa.delta) : 0.f;
        s0[kk] = kk < L ? __fmul_rn(__fmul_rn(wq.delta, qa.mn), tot) : 0.f;
    }
}

// Encoder-type weight W [N][Kr] (Conv1d C -> N with L taps, Kr = C*L; per-output-channel quantiser) -> Wc [N][KP] codes
// (zero beyond Kr), s1[o] = dw[o] * da, s0[o] = dw[o] * min_a * sum_k code[o,k]
__global__ void __launch_bounds__(32) enc_prep_kernel(const float* __restrict__ W, const float* __restrict__ wmin, const float* __restrict__ wmax,
                                                       const float* amin, const float* amax, __nv_bfloat16* __restrict__ Wc,
                                                       float* __restrict__ s1, float* __restrict__ s0, int N, int Kr, int KP) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= N) return;
    const WQ wq = make_wq(__ldg(wmin + o), __ldg(wmax + o), 8);
    const ActQF qa = load_actqf(amin, amax, 8);
    float cs = 0.f;
    for (int k = 0; k < KP; ++k) {
        const float c = k < Kr ? wq_code(wq, __ldg(W + (int64_t)o * Kr + k)) : 0.f;
        cs += c;
        Wc[(int64_t)o * KP + k] = __float2bfloat16_rn(c);
    }
    s1[o] = __fmul_rn(wq.delta, qa.delta);
    s0[o] = __fmul_rn(__fmul_rn(wq.delta, qa.mn), cs);
}

}  // namespace fqss

using namespace fqss;

extern "C" {

int fqss_frames_split(const float* g, int64_t ldg, void* out_bf16, int64_t ldo, int64_t R, int M, int L, int H, int zero_rows,
                      double* rowsum, void* stream) {
    FQSS_REQUIRE(g && out_bf16 && R > 0 && R < 65536 && M > 0 && L > 0 && 3 * L <= 128 && H > 0 && ldo >= M && ldg >= (int64_t)(M - 1) * H + L, -1,
                 "frames_split: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    FQSS_PROF("frames_split", s);
    if (rowsum) cudaMemsetAsync(rowsum, 0, (size_t)3 * L * sizeof(double), s);
    frames_split_kernel<<<dim3((M + ET_THREADS - 1) / ET_THREADS, (unsigned)(zero_rows ? 128 : 3 * L), (unsigned)R), ET_THREADS, 0, s>>>(
        g, ldg, (__nv_bfloat16*)out_bf16, ldo, M, L, H, rowsum);
    return check_launch("frames_split");
}

int fqss_frames_encode(const float* x, int64_t ldx, void* out_bf16, int64_t ldo, int64_t R, int C, int M, int L, int H, int KP,
                       const float* rmin, const float* rmax, void* stream) {
    FQSS_REQUIRE(x && out_bf16 && rmin && rmax && R > 0 && R < 65536 && C > 0 && M > 0 && L > 0 && KP >= C * L && KP <= 1024 && H > 0 &&
                     ldo >= M && ldx >= (int64_t)(M - 1) * H + L, -1, "frames_encode: bad argument");
    FQSS_PROF("frames_encode", stream);
    frames_encode_kernel<<<dim3((M + ET_THREADS - 1) / ET_THREADS, (unsigned)KP, (unsigned)R), ET_THREADS, 0, (cudaStream_t)stream>>>(
        x, ldx, (__nv_bfloat16*)out_bf16, ldo, C, M, L, H, KP, rmin, rmax);
    return check_launch("frames_encode");
}

int fqss_edge_dec_prep(const float* W, const float* wmin, const float* wmax, const float* amin, const float* amax, void* Wc,
                       void* WT, float* s1, float* s0, float* dgs, int F, int L, void* stream) {
    FQSS_REQUIRE(W && wmin && wmax && amin && amax && Wc && WT && s1 && s0 && dgs && F > 0 && L > 0 && 3 * L <= 128, -1,
                 "edge_dec_prep: bad argument");
    FQSS_PROF("edge_prep", stream);
    dec_prep_kernel<<<(F + 7) / 8, 256, 0, (cudaStream_t)stream>>>(W, wmin, wmax, amin, amax, (__nv_bfloat16*)Wc, (__nv_bfloat16*)WT, s1, s0, dgs, F, L);
    return check_launch("edge_dec_prep");
}

int fqss_edge_enc_prep(const float* W, const float* wmin, const float* wmax, const float* amin, const float* amax, void* Wc,
                       float* s1, float* s0, int N, int Kr, int KP, void* stream) {
    FQSS_REQUIRE(W && wmin && wmax && amin && amax && Wc && s1 && s0 && N > 0 && Kr > 0 && KP >= Kr, -1, "edge_enc_prep: bad argument");
    FQSS_PROF("edge_prep", stream);
    enc_prep_kernel<<<(N + 31) / 32, 32, 0, (cudaStream_t)stream>>>(W, wmin, wmax, amin, amax, (__nv_bfloat16*)Wc, s1, s0, N, Kr, KP);
    return check_launch("edge_enc_prep");
}

int fqss_sub_fq_codes(const float* a, int64_t lda, const float* b, int64_t ldb, void* out_bf16, int64_t ldo, int64_t rows, int M,
                      const float* rmin, const float* rmax, void* stream) {
    FQSS_REQUIRE(a && b && out_bf16 && rmin && rmax && rows > 0 && M > 0, -1, "sub_fq_codes: null argument");
    FQSS_REQUIRE(lda >= ((M + 3) & ~3) && ldb >= ((M + 3) & ~3) && ldo >= ((M + 3) & ~3) && lda % 4 == 0 && ldb % 4 == 0 && ldo % 4 == 0 &&
                     aligned16(a) && aligned16(b) && (((uintptr_t)out_bf16) & 7) == 0, -2, "sub_fq_codes: rows must be pitched to 4 elements and aligned");
    FQSS_PROF("sub_fq_codes", stream);
    const int nq = (M + 3) >> 2;
    int gx = (nq + ET_THREADS - 1) / ET_THREADS;
    if (gx > 4) gx = 4;
    sub_fq_codes_kernel<<<dim3(gx, (unsigned)rows), ET_THREADS, 0, (cudaStream_t)stream>>>(a, lda, b, ldb, (__nv_bfloat16*)out_bf16, ldo, M,
                                                                                         rmin, rmax);
    return check_launch("sub_fq_codes");
}

int fqss_dec_wgrad_fold(const float* part, float* dWq, int F, int L, void* stream) {
    FQSS_REQUIRE(part && dWq && F > 0 && L > 0 && 3 * L <= 128, -1, "dec_wgrad_fold: bad argument");
    FQSS_PROF("dec_wgrad_fold", stream);
    dec_wgrad_fold_kernel<<<(F * L + 255) / 256, 256, 0, (cudaStream_t)stream>>>(part, dWq, F, L);
    return check_launch("dec_wgrad_fold");
}

}  // extern "C"
