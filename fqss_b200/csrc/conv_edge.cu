// conv_edge.cu -- the filterbank convolutions at the two ends of the separator, tiled for B200:
//   * analysis conv  (encoder, RQB re-encoder, and the decoder's input gradient):
//         y[b,o,m] = sum_{c,k} w[o,c,k] x[b,c,8m+k]                      qat_layers.py:1030,1189
//   * synthesis conv (decoder, and the re-encoder's input gradient):   overlap-add of 16-tap frames
//         y[b,8q+r] = sum_c ( w[c,r] x[b,c,q] + w[c,r+8] x[b,c,q-1] )    qat_layers.py:1332,1194
//   * their weight gradient  gw[o,c,k] = sum_{b,m} gy[b,o,m] x[b,c,8m+k]
// Geometry of the recipe only: kernel 16, stride 8 (50 % overlapped frames; configs/convtasnet_2spks_8k.yaml).
// Other geometries fall back to the generic kernels in conv.cu.
//
// All three are HBM-bound on the 512-channel tensor ([B,512,M] fp32, read or written exactly once):
// the contraction is tiny (K*Cin = 16..32) so it runs on the FFMA pipe out of registers / shared memory,
// with frames on lanes so that every global access is a coalesced 128-bit (or 64-bit) row segment.
#include "fqss_common.cuh"

namespace fqss {

constexpr int EK = 16, ES = 8;          // taps, hop

// ---------------------------------------------------------------------------------------------
// analysis conv: CTA = 128 frames x 64 output channels, thread = 4 frames x 8 channels
// ---------------------------------------------------------------------------------------------
constexpr int AF = 128, AO = 64, AX = 132;      // AX: row stride of the de-interleaved window (bank-conflict free)
constexpr int A_MAXCIN = 4;

__global__ void __launch_bounds__(256) analysis_fwd_kernel(const float* __restrict__ x, int64_t ldx, int T, const float* __restrict__ w,
                                                          float* __restrict__ y, int64_t ldy, int Cin, int Co, int Mo) {
    __shared__ __align__(16) float xs[A_MAXCIN][ES][AX];      // xs[c][r][q] = x[c][8(m0+q)+r], q in [0,129]
    __shared__ __align__(16) float ws[A_MAXCIN * EK][AO];      // ws[j][o]
    const int b = blockIdx.z, o0 = blockIdx.y * AO, m0 = blockIdx.x * AF;
    const int tid = threadIdx.x;
    const int CK = Cin * EK;
    for (int i = tid; i < CK * AO; i += 256) {
        const int o = i / CK, j = i - o * CK;
        ws[j][o] = (o0 + o < Co) ? __ldg(w + (int64_t)(o0 + o) * CK + j) : 0.f;
    }
    for (int c = 0; c < Cin; ++c) {
        const float* xr = x + ((int64_t)b * Cin + c) * ldx;
        const int t0 = m0 * ES;
        for (int i = tid; i < (AF + 2) * ES; i += 256) {
            const int t = t0 + i;
            xs[c][i & 7][i >> 3] = (t < T) ? __ldg(xr + t) : 0.f;
        }
    }
    __syncthreads();
    const int fg = tid & 31, og = tid >> 5;
    float acc[4][8];
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int o = 0; o < 8; ++o) acc[f][o] = 0.f;
    for (int c = 0; c < Cin; ++c) {
#pragma unroll
        for (int r = 0; r < ES; ++r) {
            const float4 xa = *reinterpret_cast<const float4*>(&xs[c][r][4 * fg]);
            const float x4 = xs[c][r][4 * fg + 4];
            const float xv[5] = {xa.x, xa.y, xa.z, xa.w, x4};
#pragma unroll
            for (int h = 0; h < 2; ++h) {                       // tap k = r + 8h reads frame q + h
                const float4 wa = *reinterpret_cast<const float4*>(&ws[c * EK + r + ES * h][8 * og]);
                const float4 wb = *reinterpret_cast<const float4*>(&ws[c * EK + r + ES * h][8 * og + 4]);
                const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int f = 0; f < 4; ++f)
#pragma unroll
                    for (int o = 0; o < 8; ++o) acc[f][o] = fmaf(wv[o], xv[f + h], acc[f][o]);
            }
        }
    }
    const int m = m0 + 4 * fg;
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        const int oc = o0 + 8 * og + o;
        if (oc >= Co) continue;
        float* yr = y + ((int64_t)b * Co + oc) * ldy + m;
        if (m + 3 < Mo || (m < Mo && m + 3 < ldy && (ldy & 3) == 0)) {
            stg4(yr, make_float4(acc[0][o], acc[1][o], acc[2][o], acc[3][o]));   // pad columns may take garbage
        } else {
#pragma unroll
            for (int f = 0; f < 4; ++f)
                if (m + f < Mo) yr[f] = acc[f][o];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// synthesis conv: CTA = (sample, 63 owned hops q); 8 warps split the channels, lanes own 2 frames.
//   Z[f][k] = sum_c w[c*wstride + k] x[b,c,f]  for the 64 frames f = q0-1 .. q0+62, then
//   y[8q+r] = Z[q][r] + Z[q-1][r+8]  for q = q0 .. q0+62.
// ---------------------------------------------------------------------------------------------
constexpr int SQ = 63;

__global__ void __launch_bounds__(256) synthesis_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                                                           int64_t wstride, float* __restrict__ y, int64_t ldy, int Ci, int M, int T) {
    extern __shared__ __align__(16) float sm[];
    float* wsm = sm;                         // [Ci][16]
    float* zs = sm + (size_t)Ci * EK;        // [8 warps][64 frames][17]
    const int b = blockIdx.y, q0 = blockIdx.x * SQ;
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    for (int i = tid; i < Ci * EK; i += 256) wsm[i] = __ldg(w + (int64_t)(i >> 4) * wstride + (i & 15));
    __syncthreads();
    const int f0 = q0 - 1 + 2 * lane;        // this lane's frames f0, f0+1
    const bool v0 = f0 >= 0 && f0 < M, v1 = f0 + 1 >= 0 && f0 + 1 < M;
    const int cper = Ci >> 3;
    const int cb = wp * cper;
    float a0[EK], a1[EK];
#pragma unroll
    for (int k = 0; k < EK; ++k) a0[k] = a1[k] = 0.f;
    const float* xb = x + ((int64_t)b * Ci + cb) * ldx;
#pragma unroll 8
    for (int c = 0; c < cper; ++c) {
        const float* xr = xb + (int64_t)c * ldx;
        const float x0 = v0 ? __ldg(xr + f0) : 0.f;
        const float x1 = v1 ? __ldg(xr + f0 + 1) : 0.f;
        const float4* wr = reinterpret_cast<const float4*>(wsm + (size_t)(cb + c) * EK);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const float4 wv = wr[g];
            a0[4 * g] = fmaf(wv.x, x0, a0[4 * g]);         a1[4 * g] = fmaf(wv.x, x1, a1[4 * g]);
            a0[4 * g + 1] = fmaf(wv.y, x0, a0[4 * g + 1]); a1[4 * g + 1] = fmaf(wv.y, x1, a1[4 * g + 1]);
            a0[4 * g + 2] = fmaf(wv.z, x0, a0[4 * g + 2]); a1[4 * g + 2] = fmaf(wv.z, x1, a1[4 * g + 2]);
            a0[4 * g + 3] = fmaf(wv.w, x0, a0[4 * g + 3]); a1[4 * g + 3] = fmaf(wv.w, x1, a1[4 * g + 3]);
        }
    }
    float* zw = zs + (size_t)wp * 64 * 17;
#pragma unroll
    for (int k = 0; k < EK; ++k) {
        zw[(2 * lane) * 17 + k] = a0[k];
        zw[(2 * lane + 1) * 17 + k] = a1[k];
    }
    __syncthreads();
    // 63 hops x 8 samples; frame index inside the tile: q - (q0 - 1) = jq + 1
    for (int i = tid; i < SQ * ES; i += 256) {
        const int jq = i >> 3, r = i & 7;
        const int t = (q0 + jq) * ES + r;
        if (t >= T) continue;
        float s = 0.f;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) s += zs[((size_t)ww * 64 + jq + 1) * 17 + r] + zs[((size_t)ww * 64 + jq) * 17 + r + ES];
        y[(int64_t)b * ldy + t] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// weight gradient: gw[o, c*16+k] = sum_{b,m} g[b,o,m] * xs[b,c,8m+k]
//   CTA = (frame chunk, 8*R rows o, sample b); warp = R rows, a lane owns 4 consecutive frames per trip.  The signal
//   window is de-interleaved in shared memory once per CTA (xw[c][r][q] = x[c][8(m0+q)+r]) and shared by all rows, so
//   one trip costs per lane: R 128-bit loads of g, 8*CIN x (LDS.128 + LDS.32) of the window and 64*R*CIN FMAs --
//   the kernel sits on the FMA pipe (16*CIN FMAs per element of g) instead of on shared-memory bandwidth.
// ---------------------------------------------------------------------------------------------
template <int CIN, int R, int CHUNK>
__global__ void __launch_bounds__(256) edge_wgrad_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ x, int64_t ldx,
                                                        int T, int Co, int Mo, double* __restrict__ acc) {
    constexpr int WX = CHUNK + 4;                          // row stride of the window (16-byte aligned rows)
    extern __shared__ __align__(16) float xw[];            // [CIN][8][WX]
    const int b = blockIdx.z, o0 = blockIdx.y * (8 * R), m0 = blockIdx.x * CHUNK;
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int nfr = min(CHUNK, Mo - m0);
    for (int c = 0; c < CIN; ++c) {
        const float* xr = x + ((int64_t)b * CIN + c) * ldx;
        const int t0 = m0 * ES;
        for (int i = tid; i < (CHUNK + 4) * ES; i += 256) {
            const int t = t0 + i;
            xw[((size_t)c * ES + (i & 7)) * WX + (i >> 3)] = (t < T) ? __ldg(xr + t) : 0.f;
        }
    }
    __syncthreads();
    const int ob = o0 + R * wp;
    const float* gr[R];
    bool vr[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
        vr[j] = ob + j < Co;
        gr[j] = g + ((int64_t)b * Co + (vr[j] ? ob + j : 0)) * ldg + m0;
    }
    const bool vec_ok = ((ldg & 3) == 0) && ((reinterpret_cast<uintptr_t>(g) & 15) == 0) && ((m0 & 3) == 0);
    float sa[R][CIN * EK];
#pragma unroll
    for (int j = 0; j < R; ++j)
#pragma unroll
        for (int k = 0; k < CIN * EK; ++k) sa[j][k] = 0.f;
    for (int q = 4 * lane; q < nfr; q += 128) {
        float gv[R][4];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            if (vr[j] && vec_ok && q + 3 < nfr) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(gr[j] + q));
                gv[j][0] = t.x; gv[j][1] = t.y; gv[j][2] = t.z; gv[j][3] = t.w;
            } else {
#pragma unroll
                for (int f = 0; f < 4; ++f) gv[j][f] = (vr[j] && q + f < nfr) ? __ldg(gr[j] + q + f) : 0.f;
            }
        }
#pragma unroll
        for (int c = 0; c < CIN; ++c)
#pragma unroll
            for (int r = 0; r < ES; ++r) {
                const float* row = xw + ((size_t)c * ES + r) * WX + q;
                const float4 xa = *reinterpret_cast<const float4*>(row);
                const float xv[5] = {xa.x, xa.y, xa.z, xa.w, row[4]};
#pragma unroll
                for (int j = 0; j < R; ++j)
#pragma unroll
                    for (int f = 0; f < 4; ++f) {
                        sa[j][c * EK + r] = fmaf(gv[j][f], xv[f], sa[j][c * EK + r]);                   // tap r reads frame q+f
                        sa[j][c * EK + r + ES] = fmaf(gv[j][f], xv[f + 1], sa[j][c * EK + r + ES]);     // tap r+8 reads frame q+f+1
                    }
            }
    }
#pragma unroll
    for (int j = 0; j < R; ++j)
#pragma unroll
        for (int k = 0; k < CIN * EK; ++k) {
            const float t = warp_sum(sa[j][k]);
            if (lane == 0 && vr[j]) atomicAdd(acc + (int64_t)(ob + j) * (CIN * EK) + k, (double)t);
        }
}

// ---------------------------------------------------------------------------------------------
// host-side dispatch (called from the C ABI functions in conv.cu)
// ---------------------------------------------------------------------------------------------
bool edge_geometry_ok(int K, int stride) { return K == EK && stride == ES; }

int edge_analysis_fwd(const float* x, int64_t ldx, int T, const float* w, float* y, int64_t ldy, int B, int Cin, int Co, int Mo,
                      cudaStream_t s) {
    if (Cin > A_MAXCIN) return 1;      // not handled here
    dim3 grid((Mo + AF - 1) / AF, (Co + AO - 1) / AO, B);
    analysis_fwd_kernel<<<grid, 256, 0, s>>>(x, ldx, T, w, y, ldy, Cin, Co, Mo);
    return 0;
}

int edge_synthesis_fwd(const float* x, int64_t ldx, const float* w, int64_t wstride, float* y, int64_t ldy, int B, int Ci, int M, int T,
                       cudaStream_t s) {
    if (Ci % 8 != 0) return 1;
    const size_t smem = ((size_t)Ci * EK + (size_t)8 * 64 * 17) * sizeof(float);
    if (smem > 200 * 1024) return 1;
    static bool cfg = false;
    if (!cfg) {
        cudaFuncSetAttribute(synthesis_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cfg = true;
    }
    dim3 grid((M + 1 + SQ - 1) / SQ, B);      // hops q = 0 .. M
    synthesis_fwd_kernel<<<grid, 256, smem, s>>>(x, ldx, w, wstride, y, ldy, Ci, M, T);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// weight gradient, rows on lanes (the fast path; needs 16-byte aligned rows of g):
//   a lane owns ONE output row o and its 16*CIN tap sums; all lanes of a warp walk the same frames, so the signal
//   window is read as 128-bit BROADCAST loads (one wavefront) and needs no de-interleaving; g reaches the lanes
//   through a cp.async double-buffered shared tile [32 rows][128 frames] (coalesced 16-byte copies in, conflict-free
//   128-bit reads out with a row stride of 132 floats).  Per 4 frames and lane: 1 + 10*CIN LDS.128 and 64*CIN FMAs --
//   the FMA pipe is the bound (16*CIN FMAs per element of g), shared memory and HBM stay below it.
//   CTA = (1024-frame chunk, 32 rows, sample); 8 warps split every 128-frame tile 16 frames each.
// ---------------------------------------------------------------------------------------------
constexpr int WL_CH = 1024, WL_FT = 128, WL_GS = WL_FT + 4;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc),
                 "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int CIN>
__global__ void __launch_bounds__(256) edge_wgrad_lanes_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ x,
                                                              int64_t ldx, int T, int Co, int Mo, double* __restrict__ acc) {
    extern __shared__ __align__(16) float sm[];
    constexpr int XW = ES * WL_CH + 8;                     // signal samples under one chunk (+ the second half of the last frame)
    float* xs = sm;                                        // [CIN][XW]
    float* gs = sm + CIN * XW;                             // [2][32][WL_GS]
    const int b = blockIdx.z, o0 = blockIdx.y * 32, m0 = blockIdx.x * WL_CH;
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int nfr = min(WL_CH, Mo - m0);
    const int ntiles = (nfr + WL_FT - 1) / WL_FT;
    auto issue_tile = [&](int t, int buf) {                // 32 rows x 32 quads = 1024 16-byte copies, 4 per thread
        float* dst = gs + buf * 32 * WL_GS;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = tid + 256 * k, row = i >> 5, qd = i & 31;
            const int f = t * WL_FT + 4 * qd;               // first frame of the quad inside the chunk
            int nval = nfr - f;
            nval = nval < 0 ? 0 : (nval > 4 ? 4 : nval);
            if (o0 + row >= Co) nval = 0;
            const float* src = g + ((int64_t)b * Co + (o0 + row < Co ? o0 + row : 0)) * ldg + m0 + (nval ? f : 0);
            cp_async16(dst + row * WL_GS + 4 * qd, src, 4 * nval);      // bytes beyond src-size are zero-filled
        }
        cp_async_commit();
    };
    issue_tile(0, 0);
    for (int c = 0; c < CIN; ++c) {
        const float* xr = x + ((int64_t)b * CIN + c) * ldx;
        const int t0 = m0 * ES;
        for (int i = tid; i < XW; i += 256) xs[c * XW + i] = (t0 + i < T) ? __ldg(xr + t0 + i) : 0.f;
    }
    float sa[CIN * EK];
#pragma unroll
    for (int k = 0; k < CIN * EK; ++k) sa[k] = 0.f;
    for (int t = 0; t < ntiles; ++t) {
        if (t + 1 < ntiles) {
            issue_tile(t + 1, (t + 1) & 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();                                   // tile t (and, first time round, the window) visible to all
        const float* gt = gs + (t & 1) * 32 * WL_GS + lane * WL_GS + 16 * wp;
        const int fbase = t * WL_FT + 16 * wp;               // first of this warp's 16 frames inside the chunk
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 g4 = *reinterpret_cast<const float4*>(gt + 4 * q);
            const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
            for (int c = 0; c < CIN; ++c) {
                const float4* xw = reinterpret_cast<const float4*>(xs + c * XW + ES * (fbase + 4 * q));      // 40 samples
                float xv[40];
#pragma unroll
                for (int k = 0; k < 10; ++k) {
                    const float4 v = xw[k];
                    xv[4 * k] = v.x; xv[4 * k + 1] = v.y; xv[4 * k + 2] = v.z; xv[4 * k + 3] = v.w;
                }
#pragma unroll
                for (int f = 0; f < 4; ++f)
#pragma unroll
                    for (int k = 0; k < EK; ++k) sa[c * EK + k] = fmaf(gv[f], xv[ES * f + k], sa[c * EK + k]);
            }
        }
        __syncthreads();                                   // everyone is done with buffer t&1 before it is refilled
    }
    // cross-warp reduction through shared memory (the g buffers are free now): red[wp][k][lane]
    float* red = gs;
#pragma unroll
    for (int k = 0; k < CIN * EK; ++k) red[(wp * CIN * EK + k) * 32 + lane] = sa[k];
    __syncthreads();
    for (int i = tid; i < CIN * EK * 32; i += 256) {
        const int k = i >> 5, row = i & 31;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[(w * CIN * EK + k) * 32 + row];
        if (o0 + row < Co) atomicAdd(acc + (int64_t)(o0 + row) * (CIN * EK) + k, (double)s);
    }
}

template <int CIN, int R, int CHUNK>
static void edge_wgrad_launch(const float* g, int64_t ldg, const float* x, int64_t ldx, int T, int B, int Co, int Mo, double* acc,
                              cudaStream_t s) {
    constexpr size_t smem = (size_t)CIN * ES * (CHUNK + 4) * sizeof(float);
    static bool cfg = false;
    if (!cfg) {
        cudaFuncSetAttribute(edge_wgrad_kernel<CIN, R, CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cfg = true;
    }
    dim3 grid((Mo + CHUNK - 1) / CHUNK, (Co + 8 * R - 1) / (8 * R), B);
    edge_wgrad_kernel<CIN, R, CHUNK><<<grid, 256, smem, s>>>(g, ldg, x, ldx, T, Co, Mo, acc);
}

int edge_wgrad(const float* g, int64_t ldg, const float* x, int64_t ldx, int T, int B, int Cin, int Co, int Mo, double* acc, cudaStream_t s) {
    if (Cin != 1 && Cin != 2) return 1;
    if ((ldg & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        const size_t smem = ((size_t)Cin * (ES * WL_CH + 8) + 2 * 32 * WL_GS) * sizeof(float);
        static bool cfg = false;
        if (!cfg) {
            cudaFuncSetAttribute(edge_wgrad_lanes_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
            cudaFuncSetAttribute(edge_wgrad_lanes_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
            cfg = true;
        }
        dim3 grid((Mo + WL_CH - 1) / WL_CH, (Co + 31) / 32, B);
        if (Cin == 1) edge_wgrad_lanes_kernel<1><<<grid, 256, smem, s>>>(g, ldg, x, ldx, T, Co, Mo, acc);
        else edge_wgrad_lanes_kernel<2><<<grid, 256, smem, s>>>(g, ldg, x, ldx, T, Co, Mo, acc);
        return 0;
    }
    if (Cin == 1) edge_wgrad_launch<1, 4, 2048>(g, ldg, x, ldx, T, B, Co, Mo, acc, s);
    else edge_wgrad_launch<2, 2, 1024>(g, ldg, x, ldx, T, B, Co, Mo, acc, s);
    return 0;
}

}  // namespace fqss
