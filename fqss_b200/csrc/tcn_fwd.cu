// tcn_fwd.cu -- forward of the fused TCN ConvBlock (convtasnetq.py:11-42 after quantize_model).
//
//   K1  tcgen05 GEMM  x_op -> y1 = W1q x + b1            epilogue: gLN statistics of a1 = FQ1(PReLU(y1))
//   K2  this file     y1 -> [PReLU,FQ1] -> [gLN1,FQ2] -> depthwise dilated 3-tap FIR + bias -> y3
//                     one CTA per (sample, channel) row; the whole row (M <= 8K frames) is staged in
//                     shared memory so the dilation halo costs no extra HBM traffic; epilogue: gLN
//                     statistics of a3 = FQ3(PReLU(y3)).  HBM bytes: read y1 + write y3 = 8 B/element.
//   K3a this file     y3 -> [PReLU,FQ3] -> [gLN2,FQ4] -> bf16 operand (integer code)      6 B/element
//   K3  tcgen05 GEMM  a4_op -> res_y / skip_y, x_out = FQ(x + FQ(res)), skip_out = FQ(skip_in + FQ(skip))
//
// Only pre-activation tensors (y1, y3, res_y, skip_y) and the 128-wide block outputs are stored;
// every quantised activation is recomputed on load by its consumer (and again in backward).
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

// ---------------------------------------------------------------------------------------------
// K2: depthwise kernel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROW_THREADS) tcn_dw_fwd_kernel(const fqss_tcn_block p) {
    extern __shared__ float row[];                   // a2[0..M)
    __shared__ double sh[2 * 32];
    const int64_t r = blockIdx.x;
    const int b = (int)(r / p.Chid), c = (int)(r % p.Chid);
    const int M = p.M;
    Hidden1 h = load_hidden1(p, b, c);
    const float* y1 = p.y1 + r * p.ld;
    const int nvec = (M + 3) >> 2;
    for (int v = threadIdx.x; v < nvec; v += ROW_THREADS) {
        float4 y = ldg4_stream(y1 + 4 * v);
        float4 a;
        a.x = hidden1_a2(h, y.x); a.y = hidden1_a2(h, y.y); a.z = hidden1_a2(h, y.z); a.w = hidden1_a2(h, y.w);
        *reinterpret_cast<float4*>(row + 4 * v) = a;  // pad columns hold garbage; never read below (index < M)
    }
    __syncthreads();
    const float w0 = __ldg(p.wdw + c * 3), w1 = __ldg(p.wdw + c * 3 + 1), w2 = __ldg(p.wdw + c * 3 + 2);
    const float bias = __ldg(p.bdw + c);
    const float slope3 = __ldg(p.slope3);
    ActQF q3;
    if (p.quant) q3 = load_actqf(p.q3.rmin, p.q3.rmax, 8);
    const int d = p.dil;
    float* y3 = p.y3 + r * p.ld;
    float s = 0.f, ss = 0.f;
    for (int v = threadIdx.x; v < nvec; v += ROW_THREADS) {
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int m = 4 * v + k;
            float acc = 0.f;
            if (m < M) {
                const float xl = (m - d >= 0) ? row[m - d] : 0.f;
                const float xr = (m + d < M) ? row[m + d] : 0.f;
                acc = fmaf(w0, xl, acc);
                acc = fmaf(w1, row[m], acc);
                acc = fmaf(w2, xr, acc);
                acc += bias;
                float z = acc > 0.f ? acc : slope3 * acc;
                float a3 = p.quant ? actqf_fq_approx(q3, z) : z;
                s += a3;
                ss = fmaf(a3, a3, ss);
            }
            o[k] = acc;
        }
        stg4(y3 + 4 * v, make_float4(o[0], o[1], o[2], o[3]));
    }
    double vv[2] = {(double)s, (double)ss};
    block_sum<2>(vv, sh);
    if (threadIdx.x == 0) {
        atomicAdd(p.stats3 + 2 * b, vv[0]);
        atomicAdd(p.stats3 + 2 * b + 1, vv[1]);
    }
}

// ---------------------------------------------------------------------------------------------
// K3a: hidden quantiser  y3 -> a4 operand (bf16 code / value)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROW_THREADS) tcn_hidden_fq_kernel(const fqss_tcn_block p) {
    const int64_t r = blockIdx.x;
    const int b = (int)(r / p.Chid), c = (int)(r % p.Chid);
    Hidden3 h = load_hidden3(p, b, c);
    const float* y3 = p.y3 + r * p.ld;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.a4_op) + r * p.ld;
    const int nvec = (p.M + 3) >> 2;
    for (int v = threadIdx.x; v < nvec; v += ROW_THREADS) {
        float4 y = ldg4_stream(y3 + 4 * v);
        const float o0 = hidden3_op(h, y.x), o1 = hidden3_op(h, y.y), o2 = hidden3_op(h, y.z), o3 = hidden3_op(h, y.w);
        __nv_bfloat162 lo = __floats2bfloat162_rn(o0, o1);
        __nv_bfloat162 hi = __floats2bfloat162_rn(o2, o3);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        if (!p.split) {
            *reinterpret_cast<uint2*>(out + 4 * v) = pk;
        } else {
            // [hi ; lo] pair: sample b owns rows [2*b*Chid, 2*(b+1)*Chid); residual = value - bf16(value)
            __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(p.a4_op) + ((int64_t)b * 2 * p.Chid + c) * p.ld;
            __nv_bfloat16* ol = oh + (int64_t)p.Chid * p.ld;
            *reinterpret_cast<uint2*>(oh + 4 * v) = pk;
            __nv_bfloat162 rl = __floats2bfloat162_rn(o0 - __low2float(lo), o1 - __high2float(lo));
            __nv_bfloat162 rh = __floats2bfloat162_rn(o2 - __low2float(hi), o3 - __high2float(hi));
            pk.x = *reinterpret_cast<uint32_t*>(&rl);
            pk.y = *reinterpret_cast<uint32_t*>(&rh);
            *reinterpret_cast<uint2*>(ol + 4 * v) = pk;
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

// values -> GEMM operand of the first block: integer code w.r.t. the producer's quantiser, or the value itself
__global__ void __launch_bounds__(ROW_THREADS) tcn_encode_kernel(const float* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ out,
                                                                int64_t ldo, int M, const float* rmin, const float* rmax) {
    const int64_t r = blockIdx.x;
    ActQF q;
    const bool quant = rmin != nullptr;
    if (quant) q = load_actqf(rmin, rmax, 8);
    for (int m = threadIdx.x; m < M; m += ROW_THREADS) {
        float v = x[r * ldx + m];
        out[r * ldo + m] = __float2bfloat16_rn(quant ? actqf_code(q, v) : v);
    }
}

static int validate_block(const fqss_tcn_block* p, const char* who) {
    FQSS_REQUIRE(p, -1, "%s: null block", who);
    FQSS_REQUIRE(p->B > 0 && p->M > 0 && p->dil >= 1 && p->Cio > 0 && p->Chid > 0, -1, "%s: bad shape", who);
    FQSS_REQUIRE(p->Cio % 64 == 0 && p->Chid % 128 == 0 && p->Cio % 128 == 0, -1,
                 "%s: fused path needs Cio %% 128 == 0 and Chid %% 128 == 0 (got %d, %d)", who, p->Cio, p->Chid);
    FQSS_REQUIRE(p->ld >= p->M && p->ld % 8 == 0, -2, "%s: ld must be >= M and a multiple of 8", who);
    FQSS_REQUIRE((size_t)p->ld * sizeof(float) * 3 <= 200 * 1024, -1, "%s: row too long for shared-memory staging (M=%d)", who, p->M);
    FQSS_REQUIRE(p->Wc1 && p->Wc2 && p->s1_1 && p->s0_1 && p->s1_2 && p->s0_2 && p->wdw && p->bdw, -1, "%s: block not prepared", who);
    FQSS_REQUIRE(p->slope1 && p->slope3 && p->gn1_w && p->gn1_b && p->gn2_w && p->gn2_b, -1, "%s: missing layer parameters", who);
    FQSS_REQUIRE(p->x_op && p->x_in && p->y1 && p->y3 && p->stats1 && p->stats3 && p->a4_op && p->skip_out, -1,
                 "%s: missing activation buffers", who);
    FQSS_REQUIRE(!p->split || !p->quant, -1, "%s: split operands are a float-model (quant == 0) feature", who);
    if (p->has_res) FQSS_REQUIRE(p->x_out && p->x_out_op, -1, "%s: missing residual buffers", who);
    if (!p->first_block) FQSS_REQUIRE(p->skip_in, -1, "%s: missing skip_in", who);
    if (p->quant) {
        const fqss_qrange* qs[] = {&p->q1, &p->q2, &p->q3, &p->q4, &p->qskip};
        for (auto q : qs) FQSS_REQUIRE(q->rmin && q->rmax, -1, "%s: missing quantiser range", who);
        if (p->has_res) FQSS_REQUIRE(p->qres.rmin && p->qadd.rmin, -1, "%s: missing residual quantisers", who);
        if (!p->first_block) FQSS_REQUIRE(p->qadds.rmin, -1, "%s: missing skip-sum quantiser", who);
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
    tcn_prep_kernel<<<N, 128, 0, (cudaStream_t)stream>>>(W, wmin, wmax, bias, amin, amax, (__nv_bfloat16*)Wc, (__nv_bfloat16*)WcT, s1,
                                                         s0, dws, K, Ntot, n_off, split);
    return check_launch("tcn_prep");
}

int fqss_tcn_encode(const float* x, int64_t ldx, void* out_bf16, int64_t ldo, int64_t rows, int M, const float* rmin,
                    const float* rmax, void* stream) {
    FQSS_REQUIRE(x && out_bf16 && rows > 0 && M > 0 && ldx >= M && ldo >= M, -1, "tcn_encode: bad argument");
    tcn_encode_kernel<<<(unsigned)rows, ROW_THREADS, 0, (cudaStream_t)stream>>>(x, ldx, (__nv_bfloat16*)out_bf16, ldo, M, rmin, rmax);
    return check_launch("tcn_encode");
}

int fqss_split_bf16(const float* x, int64_t ldx, void* out_bf16, int64_t ldo, int64_t rows, int cols, int C, int layout, void* stream) {
    FQSS_REQUIRE(x && out_bf16 && rows > 0 && cols > 0 && ldx >= cols && (layout == 0 || (layout == 1 && C > 0 && rows % C == 0 && ldo >= cols)),
                 -1, "split_bf16: bad argument");
    split_bf16_kernel<<<(unsigned)rows, ROW_THREADS, 0, (cudaStream_t)stream>>>(x, ldx, (__nv_bfloat16*)out_bf16, ldo, cols, C, layout);
    return check_launch("split_bf16");
}

int fqss_tcn_block_fwd(const fqss_tcn_block* p, void* stream) {
    int rc = validate_block(p, "tcn_block_fwd");
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int rows = p->B * p->Chid;
    cudaMemsetAsync(p->stats1, 0, (size_t)p->B * 2 * sizeof(double), s);
    cudaMemsetAsync(p->stats3, 0, (size_t)p->B * 2 * sizeof(double), s);
    // K1
    tcg::Args a{};
    a.B = p->B; a.M = p->M; a.K = p->Cio; a.N = p->Chid; a.ld = p->ld; a.s1 = p->s1_1; a.s0 = p->s0_1; a.quant = p->quant;
    if (p->split) { a.K = 3 * p->Cio; a.a_rows = 2 * p->Cio; }
    a.out_f32 = p->y1; a.slope = p->slope1; a.q1_min = p->q1.rmin; a.q1_max = p->q1.rmax; a.stats = p->stats1;
    rc = tcg::run(tcg::EPI_EXPAND, p->x_op, p->Wc1, a, s);
    if (rc) return rc;
    // K2
    static bool cfg = false;
    const size_t smem = (size_t)p->ld * sizeof(float);
    if (!cfg) {
        cudaFuncSetAttribute(tcn_dw_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cfg = true;
    }
    tcn_dw_fwd_kernel<<<rows, ROW_THREADS, smem, s>>>(*p);
    // K3a
    tcn_hidden_fq_kernel<<<rows, ROW_THREADS, 0, s>>>(*p);
    rc = check_launch("tcn_block_fwd(K2/K3a)");
    if (rc) return rc;
    // K3
    tcg::Args k{};
    k.B = p->B; k.M = p->M; k.K = p->Chid; k.N = p->has_res ? 2 * p->Cio : p->Cio; k.ld = p->ld;
    if (p->split) { k.K = 3 * p->Chid; k.a_rows = 2 * p->Chid; k.split = 1; }
    k.s1 = p->s1_2; k.s0 = p->s0_2; k.quant = p->quant;
    k.n_res = p->has_res ? p->Cio : 0; k.first_block = p->first_block;
    k.res_y = p->res_y; k.skip_y = p->skip_y; k.x_in = p->x_in; k.x_out = p->x_out; k.x_out_op = (__nv_bfloat16*)p->x_out_op;
    k.skip_in = p->skip_in; k.skip_out = p->skip_out;
    k.qres_min = p->qres.rmin; k.qres_max = p->qres.rmax; k.qskip_min = p->qskip.rmin; k.qskip_max = p->qskip.rmax;
    k.qadd_min = p->qadd.rmin; k.qadd_max = p->qadd.rmax; k.qadds_min = p->qadds.rmin; k.qadds_max = p->qadds.rmax;
    return tcg::run(tcg::EPI_RESSKIP, p->a4_op, p->Wc2, k, s);
}

}  // extern "C"
