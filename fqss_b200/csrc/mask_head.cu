// mask_head.cu -- backward of the quantised mask head (convtasnetq.py:97-99 mask_net[1..2] = Conv1dNlQ(1x1, ReLU) and
// :203 MulQ):   y = conv1x1(x) ; mask = FQ_m(relu(y)) ; masked[b,s,f,m] = FQ_p(mask[b,s,f,m] * feats[b,f,m])
//
// Forward runs in the epilogue of the tcgen05 GEMM (gemm_tc.cu, EPI_RELU_MUL with quant).  This kernel is the whole
// elementwise backward in ONE pass over the saved pre-activation y:
//   g_z    = STE_p(g)                      (FQ_p: mask on the rounded code, range sums sD_p / sZ_p)
//   g_mask = g_z * feats ; g_feats[b,f,m] = sum_s g_z * mask          (both speakers of a filter row in one CTA: no atomics)
//   g_r    = STE_m(g_mask)                 (FQ_m range sums)
//   g_y    = y > 0 ? g_r : 0               (ReLU)
//   dY     = bf16(dws[o] * g_y)            (pre-scaled operand of the dgrad / wgrad GEMMs) ; db[o] += g_y
// instead of MulQ backward + ReLU/FQ backward + row-scale (three passes, 3.1 GB at batch 32).  Algorithmic bytes per
// element of masked: 4 (g) + 4 (y) + 2 (dY) + (4 + 4) / S (feats, g_feats).
#include "fqss_common.cuh"
#include "tcn_common.cuh"

namespace fqss {

int num_sms();

constexpr int MH_THREADS = 256;
constexpr int MH_QUADS = 4;
constexpr int MH_CHUNK = MH_THREADS * 4 * MH_QUADS;

struct MaskHeadBwd {
    const float* g; int64_t ldg;        // [B][S*C][ldg] gradient w.r.t. masked
    const float* y;                     // [B][S*C][ld]  saved pre-activation
    const float* feats;                 // [B][C][ld]
    const float* dws;                   // [S*C] per-output-channel weight step (dY is pre-scaled by it)
    const float* qm_min; const float* qm_max; const float* qp_min; const float* qp_max;
    __nv_bfloat16* dY;                  // [B][S*C][ld]
    float* g_feats;                     // [B][C][ld]
    double* acc;                        // [4 + S*C]: sD_m, sZ_m, sD_p, sZ_p, then db[o]
    int B, C, S, M; int64_t ld;
};

__device__ __forceinline__ float mh_fq_bwd(const ActQF& q, float z, float g, float& sD, float& sZ) {
    const float t = actqf_t(q, z);
    const bool in = actqf_inside(q, t);
    const float c = actqf_unbias(actqf_biased(q, t));
    sD = fmaf(g, in ? __fsub_rn(c, t) : c, sD);
    sZ += in ? 0.f : g;
    return in ? exact_div(__fmul_rn(g, q.delta), q.delta, q.inv) : 0.f;
}

__global__ void __launch_bounds__(MH_THREADS) mask_head_bwd_kernel(const MaskHeadBwd p) {
    __shared__ double sh[6 * 32];
    const int64_t row2 = blockIdx.x;                      // (b, f)
    const int b = (int)(row2 / p.C), f = (int)(row2 % p.C);
    const int64_t chunk0 = (int64_t)blockIdx.y * MH_CHUNK;
    const ActQF qm = load_actqf(p.qm_min, p.qm_max, 8), qp = load_actqf(p.qp_min, p.qp_max, 8);
    float s[4] = {0.f, 0.f, 0.f, 0.f};                    // sD_m, sZ_m, sD_p, sZ_p
    float dbs[2] = {0.f, 0.f};                            // bias-gradient partials of the (up to two) speaker rows
    const bool vec = ((p.ld | p.ldg) & 3) == 0;
    for (int qd = 0; qd < MH_QUADS; ++qd) {
        const int64_t c0 = chunk0 + ((int64_t)qd * MH_THREADS + threadIdx.x) * 4;
        if (c0 >= p.M) break;
        const int nv = (int)min((int64_t)4, (int64_t)p.M - c0);
        float ft[4], gf[4] = {0.f, 0.f, 0.f, 0.f};
        if (vec) {
            const float4 v = ldg4(p.feats + row2 * p.ld + c0);
            ft[0] = v.x; ft[1] = v.y; ft[2] = v.z; ft[3] = v.w;
        } else {
            for (int k = 0; k < 4; ++k) ft[k] = k < nv ? p.feats[row2 * p.ld + c0 + k] : 0.f;
        }
        for (int sp = 0; sp < p.S; ++sp) {
            const int o = sp * p.C + f;
            const int64_t row1 = (int64_t)b * p.S * p.C + o;
            const float sc = __ldg(p.dws + o);
            float yv[4], gv[4], out[4];
            if (vec) {
                const float4 a = ldg4_stream(p.y + row1 * p.ld + c0), c = ldg4_stream(p.g + row1 * p.ldg + c0);
                yv[0] = a.x; yv[1] = a.y; yv[2] = a.z; yv[3] = a.w;
                gv[0] = c.x; gv[1] = c.y; gv[2] = c.z; gv[3] = c.w;
            } else {
                for (int k = 0; k < 4; ++k) {
                    yv[k] = k < nv ? p.y[row1 * p.ld + c0 + k] : 0.f;
                    gv[k] = k < nv ? p.g[row1 * p.ldg + c0 + k] : 0.f;
                }
            }
            float dsum = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool valid = k < nv;
                const float yk = valid ? yv[k] : 0.f, gk = valid ? gv[k] : 0.f, fk = valid ? ft[k] : 0.f;
                const float r = fmaxf(yk, 0.f);
                const float vm = actqf_fq(qm, r);
                float dD = 0.f, dZ = 0.f;
                const float gz = mh_fq_bwd(qp, __fmul_rn(vm, fk), gk, dD, dZ);
                if (valid) { s[2] += dD; s[3] += dZ; }
                gf[k] = fmaf(gz, vm, gf[k]);
                dD = 0.f; dZ = 0.f;
                const float gr = mh_fq_bwd(qm, r, gz * fk, dD, dZ);
                if (valid) { s[0] += dD; s[1] += dZ; }
                const float gy = (valid && yk > 0.f) ? gr : 0.f;
                dsum += gy;
                out[k] = gy * sc;
            }
            if (sp < 2) dbs[sp] += dsum; else atomicAdd(p.acc + 4 + o, (double)dsum);
            __nv_bfloat16* dst = p.dY + row1 * p.ld + c0;
            if (vec) {
                *reinterpret_cast<uint2*>(dst) = float4_to_bf16x4(out[0], out[1], out[2], out[3]);     // pad columns (< ld): don't-care
            } else {
                for (int k = 0; k < nv; ++k) dst[k] = __float2bfloat16_rn(out[k]);
            }
        }
        float* gd = p.g_feats + row2 * p.ld + c0;
        if (vec) {
            stg4(gd, make_float4(gf[0], gf[1], gf[2], gf[3]));
        } else {
            for (int k = 0; k < nv; ++k) gd[k] = gf[k];
        }
    }
    // block sums: 4 range sums + 2 bias partials
    float v6[6] = {s[0], s[1], s[2], s[3], dbs[0], dbs[1]};
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 6; ++i) v6[i] = warp_sum(v6[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 6; ++i) sh[i * 32 + wid] = (double)v6[i];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double x = lane < MH_THREADS / 32 ? sh[i * 32 + lane] : 0.0;
            x = warp_sum(x);
            if (lane == 0) {
                if (i < 4) atomicAdd(p.acc + i, x);
                else if (i - 4 < p.S) atomicAdd(p.acc + 4 + (i - 4) * p.C + f, x);
            }
        }
    }
}

// acc -> range gradients {g_min_m, g_max_m, g_min_p, g_max_p} and the fp32 bias gradient
__global__ void mask_head_finalize_kernel(const double* __restrict__ acc, float* __restrict__ g_q, float* __restrict__ g_bias, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2) {
        const double sD = acc[2 * i], sZ = acc[2 * i + 1];
        g_q[2 * i] = (float)(sZ - sD / 255.0);
        g_q[2 * i + 1] = (float)(sD / 255.0);
    }
    if (g_bias && i < N) g_bias[i] = (float)acc[4 + i];
}

}  // namespace fqss

using namespace fqss;

extern "C" {

size_t fqss_mask_head_ws_bytes(int N) { return (size_t)(4 + N) * sizeof(double); }

int fqss_mask_head_bwd(const float* g, int64_t ldg, const float* y, const float* feats, const float* dws, const float* qm_min,
                       const float* qm_max, const float* qp_min, const float* qp_max, void* dY_bf16, float* g_feats, float* g_q,
                       float* g_bias, double* db_f64, int B, int C, int S, int M, int64_t ld, void* ws, size_t ws_bytes, void* stream) {
    FQSS_REQUIRE(g && y && feats && dws && qm_min && qm_max && qp_min && qp_max && dY_bf16 && g_feats && g_q && ws, -1,
                 "mask_head_bwd: null argument");
    FQSS_REQUIRE(B > 0 && C > 0 && S > 0 && M > 0 && ld >= M && ldg >= M, -1, "mask_head_bwd: bad shape");
    FQSS_REQUIRE(ws_bytes >= fqss_mask_head_ws_bytes(S * C), -3, "mask_head_bwd: workspace too small");
    FQSS_REQUIRE(aligned16(g) && aligned16(y) && aligned16(feats) && aligned16(dY_bf16) && aligned16(g_feats), -2,
                 "mask_head_bwd: buffers must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    MaskHeadBwd p;
    p.g = g; p.ldg = ldg; p.y = y; p.feats = feats; p.dws = dws; p.qm_min = qm_min; p.qm_max = qm_max; p.qp_min = qp_min; p.qp_max = qp_max;
    p.dY = (__nv_bfloat16*)dY_bf16; p.g_feats = g_feats; p.acc = (double*)ws; p.B = B; p.C = C; p.S = S; p.M = M; p.ld = ld;
    FQSS_PROFN("mask_head_bwd", s, 2);
    cudaMemsetAsync(ws, 0, fqss_mask_head_ws_bytes(S * C), s);
    dim3 grid((unsigned)((int64_t)B * C), (unsigned)((M + MH_CHUNK - 1) / MH_CHUNK));
    mask_head_bwd_kernel<<<grid, MH_THREADS, 0, s>>>(p);
    int rc = check_launch("mask_head_bwd");
    if (rc) return rc;
    const int N = S * C;
    mask_head_finalize_kernel<<<(N + 255) / 256, 256, 0, s>>>((const double*)ws, g_q, g_bias, N);
    rc = check_launch("mask_head_bwd(finalize)");
    if (rc) return rc;
    if (db_f64) cudaMemcpyAsync(db_f64, (const double*)ws + 4, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, s);
    return 0;
}

}  // extern "C"
