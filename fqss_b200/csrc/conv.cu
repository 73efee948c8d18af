// conv.cu -- the convolutions of the separator on the fp32 FFMA path ("rel 1e-3" tier):
//   1x1 convs as batched SGEMM (fwd / dgrad / wgrad), the depthwise dilated conv of the TCN block,
//   the strided analysis conv of the encoder / RQB re-encoder and the transposed synthesis conv of
//   the decoder.  Reference call sites: F.conv1d / F.conv_transpose1d in qat_layers.py:138,203,
//   1030,1189,1194,1332.  The tensor-core (tcgen05) GEMMs of the fused TCN path live in tcn_*.cu.
#include "fqss_common.cuh"

namespace fqss {

int num_sms();

// tiled kernels for the recipe's filterbank geometry (kernel 16, hop 8): conv_edge.cu.  Return 1 when a shape
// is not handled there (the generic kernels below then run).
bool edge_geometry_ok(int K, int stride);
int edge_analysis_fwd(const float* x, int64_t ldx, int T, const float* w, float* y, int64_t ldy, int B, int Cin, int Co, int Mo,
                      cudaStream_t s);
int edge_synthesis_fwd(const float* x, int64_t ldx, const float* w, int64_t wstride, float* y, int64_t ldy, int B, int Ci, int M, int T,
                       cudaStream_t s);
int edge_wgrad(const float* g, int64_t ldg, const float* x, int64_t ldx, int T, int B, int Cin, int Co, int Mo, double* acc, cudaStream_t s);

// =============================================================================================
// batched SGEMM, 128x128x8 tiles, 256 threads, 8x8 register tile per thread
//   MODE 0 fwd  : C[o,m] = sum_i W[o,i]   X_b[i,m]  (+bias[o])
//   MODE 1 dgrad: C[i,m] = sum_o W[o,i]   GY_b[o,m]
//   MODE 2 wgrad: C[o,i] += sum_m GY_b[o,m] X_b[i,m]   (split over b and m-chunks, fp32 atomics)
// =============================================================================================
constexpr int GB = 128, GK = 8, GPAD = 4;

struct GemmArgs {
    const float* A; int64_t lda; int64_t batchA;   // W (modes 0/1) or GY (mode 2)
    const float* Bm; int64_t ldb; int64_t batchB;  // X or GY
    float* C; int64_t ldc; int64_t batchC;
    const float* bias;
    int R, Cn, K;        // C is R x Cn, reduction K
    int kchunk;          // mode 2: m-chunk length per block
};

template <int MODE>
__global__ void __launch_bounds__(256) sgemm_kernel(const GemmArgs g) {
    __shared__ __align__(16) float As[GK][GB + GPAD];
    __shared__ __align__(16) float Bs[GK][GB + GPAD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int r0 = blockIdx.y * GB, c0 = blockIdx.x * GB;
    int b, kbeg, kend;
    if (MODE == 2) {
        const int nchunk = (g.K + g.kchunk - 1) / g.kchunk;
        b = blockIdx.z / nchunk;
        kbeg = (blockIdx.z % nchunk) * g.kchunk;
        kend = min(g.K, kbeg + g.kchunk);
    } else {
        b = blockIdx.z; kbeg = 0; kend = g.K;
    }
    const float* A = g.A + (MODE == 2 ? (int64_t)b * g.batchA : 0);
    const float* Bm = g.Bm + (int64_t)b * g.batchB;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int k0 = kbeg; k0 < kend; k0 += GK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int idx = tid + j * 256;
            // ---- A tile -> As[k][r]
            if (MODE == 1) {                       // A(r,k) = W[k*lda + r] : r contiguous
                int k = idx >> 7, r = idx & 127;
                float v = 0.f;
                if (r0 + r < g.R && k0 + k < kend) v = __ldg(A + (int64_t)(k0 + k) * g.lda + r0 + r);
                As[k][r] = v;
            } else {                               // A(r,k) = A[r*lda + k] : k contiguous
                int r = idx >> 3, k = idx & 7;
                float v = 0.f;
                if (r0 + r < g.R && k0 + k < kend) v = __ldg(A + (int64_t)(r0 + r) * g.lda + k0 + k);
                As[k][r] = v;
            }
            // ---- B tile -> Bs[k][c]
            if (MODE == 2) {                       // B(k,c) = X[c*ldb + k] : k contiguous
                int c = idx >> 3, k = idx & 7;
                float v = 0.f;
                if (c0 + c < g.Cn && k0 + k < kend) v = __ldg(Bm + (int64_t)(c0 + c) * g.ldb + k0 + k);
                Bs[k][c] = v;
            } else {                               // B(k,c) = X[k*ldb + c] : c contiguous
                int k = idx >> 7, c = idx & 127;
                float v = 0.f;
                if (c0 + c < g.Cn && k0 + k < kend) v = __ldg(Bm + (int64_t)(k0 + k) * g.ldb + c0 + c);
                Bs[k][c] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* C = g.C + (MODE == 2 ? 0 : (int64_t)b * g.batchC);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = r0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
        if (r >= g.R) continue;
        const float bv = (MODE == 0 && g.bias) ? __ldg(g.bias + r) : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
            if (c >= g.Cn) continue;
            if (MODE == 2) atomicAdd(C + (int64_t)r * g.ldc + c, acc[i][j]);
            else C[(int64_t)r * g.ldc + c] = acc[i][j] + bv;
        }
    }
}

// row sums (bias gradients): out[r % C] += sum_m x[r, m]
__global__ void __launch_bounds__(256) rowsum_kernel(const float* __restrict__ x, int64_t cols, int64_t ld, int C,
                                                    double* __restrict__ acc) {
    __shared__ double sh[32];
    const int64_t row = blockIdx.x;
    double s = 0.0;
    for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) s += (double)x[row * ld + c];
    double v[1] = {s};
    block_sum<1>(v, sh);
    if (threadIdx.x == 0) atomicAdd(acc + (row % C), v[0]);
}

__global__ void f64_store_kernel(const double* __restrict__ a, float* __restrict__ o, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = (float)a[i];
}

// =============================================================================================
// depthwise dilated conv ("same" zero padding), K taps
// =============================================================================================
constexpr int DW_MAXK = 7;

__global__ void __launch_bounds__(256) dwconv_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ y, int64_t ldy,
                                                        int C, int M, int K, int dil) {
    const int64_t row = blockIdx.x;
    const int c = (int)(row % C);
    const int m = blockIdx.y * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float* xr = x + row * ldx;
    const int half = (K - 1) / 2;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {                      // ATen accumulates taps in order, fp32
        int j = m + (k - half) * dil;
        float xv = (j >= 0 && j < M) ? __ldg(xr + j) : 0.f;
        acc = fmaf(__ldg(w + c * K + k), xv, acc);
    }
    y[row * ldy + m] = acc + (bias ? __ldg(bias + c) : 0.f);
}

// K = 3 on 16-byte aligned rows: one thread = 4 consecutive frames (128-bit loads of the centre quad and, for dil % 4 == 0,
// of both tap quads; 128-bit store).  Same tap order and fp32 FMA chain as the scalar kernel, so the results are bit-identical.
__global__ void __launch_bounds__(256) dwconv3_fwd_vec_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ y, int64_t ldy,
                                                             int C, int M, int dil) {
    const int64_t row = blockIdx.x;
    const int c = (int)(row % C);
    const int m0 = 4 * (blockIdx.y * blockDim.x + threadIdx.x);
    if (m0 >= M) return;
    const float* xr = x + row * ldx;
    const float w0 = __ldg(w + c * 3), w1 = __ldg(w + c * 3 + 1), w2 = __ldg(w + c * 3 + 2);
    const float bb = bias ? __ldg(bias + c) : 0.f;
    const bool full = m0 + 3 < M, al = (dil & 3) == 0;
    float xc[4], xl[4], xq[4];
    auto quad = [&](float (&o)[4], int j0, bool vec) {
        if (vec) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(xr + j0));
            o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = j0 + i;
                o[i] = (j >= 0 && j < M) ? __ldg(xr + j) : 0.f;
            }
        }
    };
    quad(xc, m0, full);
    quad(xl, m0 - dil, al && m0 - dil >= 0 && m0 - dil + 3 < M);
    quad(xq, m0 + dil, al && m0 + dil + 3 < M);
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = fmaf(w2, xq[i], fmaf(w1, xc[i], fmaf(w0, xl[i], 0.f))) + bb;
    float* yr = y + row * ldy + m0;
    if (full) {
        *reinterpret_cast<float4*>(yr) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (m0 + i < M) yr[i] = o[i];
    }
}

// gx[j] = sum_k w[k] gy[j - (k-half)*dil];  per-row partial dW / dbias -> fp64 accumulators
__global__ void __launch_bounds__(256) dwconv_bwd_kernel(const float* __restrict__ gy, int64_t ldgy, const float* __restrict__ x,
                                                        int64_t ldx, const float* __restrict__ w, float* __restrict__ gx,
                                                        int64_t ldgx, int C, int M, int K, int dil,
                                                        double* __restrict__ acc) {
    __shared__ double sh[(DW_MAXK + 1) * 32];
    const int64_t row = blockIdx.x;
    const int c = (int)(row % C);
    const int half = (K - 1) / 2;
    const float* gr = gy + row * ldgy;
    const float* xr = x + row * ldx;
    float pw[DW_MAXK + 1];
#pragma unroll
    for (int k = 0; k <= DW_MAXK; ++k) pw[k] = 0.f;
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        float g0 = __ldg(gr + m);
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < DW_MAXK; ++k) {
            if (k < K) {
                int off = (k - half) * dil;
                int jg = m - off;                      // gx[m] += w[k] * gy[m - off]
                if (gx && jg >= 0 && jg < M) a = fmaf(__ldg(w + c * K + k), __ldg(gr + jg), a);
                int jx = m + off;                      // dW[k] += gy[m] * x[m + off]
                if (jx >= 0 && jx < M) pw[k] = fmaf(g0, __ldg(xr + jx), pw[k]);
            }
        }
        pw[DW_MAXK] += g0;
        if (gx) gx[row * ldgx + m] = a;
    }
    double v[DW_MAXK + 1];
#pragma unroll
    for (int k = 0; k <= DW_MAXK; ++k) v[k] = (double)pw[k];
    block_sum<DW_MAXK + 1>(v, sh);
    if (threadIdx.x == 0) {
        for (int k = 0; k < K; ++k) atomicAdd(acc + c * (DW_MAXK + 1) + k, v[k]);
        atomicAdd(acc + c * (DW_MAXK + 1) + DW_MAXK, v[DW_MAXK]);
    }
}

// K = 3 on 16-byte aligned rows: one thread = 4 consecutive frames per trip, 128-bit accesses (see dwconv3_fwd_vec_kernel);
// gx keeps the scalar kernel's tap order (bit-identical), the dW / dbias partial sums are fp32 per thread, fp64 beyond
__global__ void __launch_bounds__(256) dwconv3_bwd_vec_kernel(const float* __restrict__ gy, int64_t ldgy, const float* __restrict__ x,
                                                             int64_t ldx, const float* __restrict__ w, float* __restrict__ gx,
                                                             int64_t ldgx, int C, int M, int dil, double* __restrict__ acc) {
    __shared__ double sh[(DW_MAXK + 1) * 32];
    const int64_t row = blockIdx.x;
    const int c = (int)(row % C);
    const float* gr = gy + row * ldgy;
    const float* xr = x + row * ldx;
    const float w0 = __ldg(w + c * 3), w1 = __ldg(w + c * 3 + 1), w2 = __ldg(w + c * 3 + 2);
    const bool al = (dil & 3) == 0;
    auto quad = [&](const float* src, float (&o)[4], int j0, bool vec) {
        if (vec) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(src + j0));
            o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = j0 + i;
                o[i] = (j >= 0 && j < M) ? __ldg(src + j) : 0.f;
            }
        }
    };
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, pb = 0.f;
    const int nq = (M + 3) >> 2;
    for (int v = threadIdx.x; v < nq; v += blockDim.x) {
        const int m0 = 4 * v;
        const bool full = m0 + 3 < M;
        const bool vl = al && m0 - dil >= 0 && m0 - dil + 3 < M, vr = al && m0 + dil + 3 < M;
        float gc[4], gl[4], gq[4], xc[4], xl[4], xq[4];
        quad(gr, gc, m0, full);
        quad(xr, xc, m0, full);
        quad(xr, xl, m0 - dil, vl);
        quad(xr, xq, m0 + dil, vr);
#pragma unroll
        for (int i = 0; i < 4; ++i) {          // dW[k] += gy[m] * x[m + (k-1) dil] (gy is zero beyond M by the guarded load)
            p0 = fmaf(gc[i], xl[i], p0);
            p1 = fmaf(gc[i], xc[i], p1);
            p2 = fmaf(gc[i], xq[i], p2);
            pb += gc[i];
        }
        if (gx) {
            quad(gr, gl, m0 - dil, vl);
            quad(gr, gq, m0 + dil, vr);
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = fmaf(w2, gl[i], fmaf(w1, gc[i], fmaf(w0, gq[i], 0.f)));
            float* gxr = gx + row * ldgx + m0;
            if (full) {
                *reinterpret_cast<float4*>(gxr) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (m0 + i < M) gxr[i] = o[i];
            }
        }
    }
    double v[DW_MAXK + 1];
#pragma unroll
    for (int k = 0; k <= DW_MAXK; ++k) v[k] = 0.0;
    v[0] = (double)p0; v[1] = (double)p1; v[2] = (double)p2; v[DW_MAXK] = (double)pb;
    block_sum<DW_MAXK + 1>(v, sh);
    if (threadIdx.x == 0) {
        for (int k = 0; k < 3; ++k) atomicAdd(acc + c * (DW_MAXK + 1) + k, v[k]);
        atomicAdd(acc + c * (DW_MAXK + 1) + DW_MAXK, v[DW_MAXK]);
    }
}

__global__ void dwconv_finalize_kernel(const double* __restrict__ acc, float* gw, float* gbias, int C, int K) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (gw) for (int k = 0; k < K; ++k) gw[c * K + k] = (float)acc[c * (DW_MAXK + 1) + k];
    if (gbias) gbias[c] = (float)acc[c * (DW_MAXK + 1) + DW_MAXK];
}

// =============================================================================================
// strided analysis conv (encoder): y[b,o,m] = sum_{c,k} w[o,c,k] x[b,c,m*stride+k]
//   block = 128 frames x SC_OT output channels; weights broadcast from shared memory
// =============================================================================================
constexpr int SC_OT = 32;
constexpr int SC_MAXCK = 96;     // Cin*K <= 96 (2 x 16 in the speech recipe, 4 x 20 in the music model)

__global__ void __launch_bounds__(128) sconv_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                                                       float* __restrict__ y, int64_t ldy, int Cin, int Co, int Mo, int K,
                                                       int stride) {
    __shared__ float ws[SC_OT][SC_MAXCK];
    const int b = blockIdx.z, o0 = blockIdx.y * SC_OT;
    const int m = blockIdx.x * 128 + threadIdx.x;
    const int CK = Cin * K;
    for (int i = threadIdx.x; i < SC_OT * CK; i += 128) {
        int o = i / CK, j = i - o * CK;
        ws[o][j] = (o0 + o < Co) ? __ldg(w + (int64_t)(o0 + o) * CK + j) : 0.f;
    }
    __syncthreads();
    if (m >= Mo) return;
    float acc[SC_OT];
#pragma unroll
    for (int o = 0; o < SC_OT; ++o) acc[o] = 0.f;
    for (int c = 0; c < Cin; ++c) {
        const float* xr = x + ((int64_t)b * Cin + c) * ldx + (int64_t)m * stride;
        for (int k = 0; k < K; ++k) {
            float xv = __ldg(xr + k);
#pragma unroll
            for (int o = 0; o < SC_OT; ++o) acc[o] = fmaf(ws[o][c * K + k], xv, acc[o]);
        }
    }
#pragma unroll
    for (int o = 0; o < SC_OT; ++o)
        if (o0 + o < Co) y[((int64_t)b * Co + o0 + o) * ldy + m] = acc[o];
}

// transposed synthesis (overlap-add) to one channel:
//   y[b,t] = sum_c sum_{m,k : m*stride+k=t} w[c*wstride + k] x[b,c,m]
// one thread per output sample t; K <= 2*stride in the recipe but any K works.
__global__ void __launch_bounds__(256) tconv_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                                                       int64_t wstride, float* __restrict__ y, int64_t ldy, int Ci, int M,
                                                       int K, int stride, int T) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    int m_hi = min(t / stride, M - 1);
    int m_lo = max(0, (t - K + stride) / stride);      // smallest m with t - m*stride <= K-1
    float acc = 0.f;
    for (int c = 0; c < Ci; ++c) {
        const float* xr = x + ((int64_t)b * Ci + c) * ldx;
        const float* wr = w + (int64_t)c * wstride;
        for (int m = m_lo; m <= m_hi; ++m) acc = fmaf(__ldg(wr + (t - m * stride)), __ldg(xr + m), acc);
    }
    y[(int64_t)b * ldy + t] = acc;
}

// weight gradient shared by both: gw[o*CK + c*K + k] = sum_{b,m} gy[b,o,m] * x[b,c,m*stride+k]
//   grid (chunks of 1024 frames, Co, B); 32 partial sums per thread, block reduce, fp64 atomics
__global__ void __launch_bounds__(256) sconv_gw_kernel(const float* __restrict__ gy, int64_t ldgy, const float* __restrict__ x,
                                                      int64_t ldx, int Cin, int Co, int Mo, int K, int stride,
                                                      double* __restrict__ acc) {
    __shared__ double sh[32];
    const int b = blockIdx.z, o = blockIdx.y;
    const int CK = Cin * K;
    const float* gr = gy + ((int64_t)b * Co + o) * ldgy;
    const int mbeg = blockIdx.x * 1024, mend = min(Mo, mbeg + 1024);
    for (int j = 0; j < CK; ++j) {
        const int c = j / K, k = j - c * K;
        const float* xr = x + ((int64_t)b * Cin + c) * ldx + k;
        float s = 0.f;
        for (int m = mbeg + threadIdx.x; m < mend; m += blockDim.x) s = fmaf(__ldg(gr + m), __ldg(xr + (int64_t)m * stride), s);
        double v[1] = {(double)s};
        block_sum<1>(v, sh);
        if (threadIdx.x == 0) atomicAdd(acc + (int64_t)o * CK + j, v[0]);
        __syncthreads();
    }
}

// Tiled variant for geometries the kernel-16 / hop-8 kernels of conv_edge.cu do not cover (the music encoder: Cin = 4,
// K = 20, hop 10): one CTA = GW_OT output channels x all Cin*K taps over GW_MF frames.  The input span of the chunk (all
// channels) and the GW_OT gradient rows are staged in shared memory once; thread (tap j, group q) owns the sums of tap j
// for GW_OT / 4 output channels over the whole chunk -- no cross-thread reduction, one fp64 atomic per sum.  Per frame and
// thread: one LDS of x (consecutive taps -> consecutive banks) and GW_OT / 4 broadcast LDS of gy.
constexpr int GW_OT = 8, GW_MF = 256;

__global__ void __launch_bounds__(4 * SC_MAXCK) sconv_gw_tile_kernel(const float* __restrict__ gy, int64_t ldgy, const float* __restrict__ x,
                                                                     int64_t ldx, int T, int Cin, int Co, int Mo, int K, int stride,
                                                                     double* __restrict__ acc) {
    extern __shared__ float gsm[];                     // [Cin][span] input, then [GW_OT][GW_MF] gradient
    const int b = blockIdx.z, o0 = blockIdx.y * GW_OT, m0 = blockIdx.x * GW_MF;
    const int CK = Cin * K, nm = min(GW_MF, Mo - m0);
    const int span = (GW_MF - 1) * stride + K;
    float* xs = gsm;
    float* gs = gsm + (size_t)Cin * span;
    const int64_t t0 = (int64_t)m0 * stride;
    for (int i = threadIdx.x; i < Cin * span; i += blockDim.x) {
        const int c = i / span, t = i - c * span;
        xs[i] = (t0 + t < T) ? __ldg(x + ((int64_t)b * Cin + c) * ldx + t0 + t) : 0.f;
    }
    for (int i = threadIdx.x; i < GW_OT * GW_MF; i += blockDim.x) {
        const int o = i / GW_MF, m = i - o * GW_MF;
        gs[i] = (o0 + o < Co && m < nm) ? __ldg(gy + ((int64_t)b * Co + o0 + o) * ldgy + m0 + m) : 0.f;
    }
    __syncthreads();
    const int j = threadIdx.x % CK, q = threadIdx.x / CK;          // blockDim.x == 4 * CK
    const int c = j / K, k = j - c * K;
    const float* xr = xs + (size_t)c * span + k;
    const float* g0 = gs + (size_t)(2 * q) * GW_MF;
    const float* g1 = g0 + GW_MF;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
    for (int m = 0; m < GW_MF; ++m) {                               // frames beyond nm carry zero gradient
        const float xv = xr[m * stride];
        s0 = fmaf(g0[m], xv, s0);
        s1 = fmaf(g1[m], xv, s1);
    }
    const int oa = o0 + 2 * q, ob = oa + 1;
    if (oa < Co) atomicAdd(acc + (int64_t)oa * CK + j, (double)s0);
    if (ob < Co) atomicAdd(acc + (int64_t)ob * CK + j, (double)s1);
}

}  // namespace fqss

using namespace fqss;

extern "C" {

int fqss_conv1x1_fwd(const float* x, int64_t ldx, const float* w, const float* bias, float* y, int64_t ldy, int B, int Ci,
                     int Co, int M, void* stream) {
    FQSS_REQUIRE(x && w && y && B > 0 && Ci > 0 && Co > 0 && M > 0 && ldx >= M && ldy >= M, -1, "conv1x1_fwd: bad argument");
    GemmArgs g{w, Ci, 0, x, ldx, (int64_t)Ci * ldx, y, ldy, (int64_t)Co * ldy, bias, Co, M, Ci, 0};
    dim3 grid((M + GB - 1) / GB, (Co + GB - 1) / GB, B);
    FQSS_PROF("conv1x1_fwd(fp32)", stream);
    sgemm_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(g);
    return check_launch("conv1x1_fwd");
}

int fqss_conv1x1_dgrad(const float* gy, int64_t ldgy, const float* w, float* gx, int64_t ldgx, int B, int Ci, int Co, int M,
                       void* stream) {
    FQSS_REQUIRE(gy && w && gx && B > 0 && Ci > 0 && Co > 0 && M > 0 && ldgy >= M && ldgx >= M, -1, "conv1x1_dgrad: bad argument");
    GemmArgs g{w, Ci, 0, gy, ldgy, (int64_t)Co * ldgy, gx, ldgx, (int64_t)Ci * ldgx, nullptr, Ci, M, Co, 0};
    dim3 grid((M + GB - 1) / GB, (Ci + GB - 1) / GB, B);
    FQSS_PROF("conv1x1_dgrad(fp32)", stream);
    sgemm_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(g);
    return check_launch("conv1x1_dgrad");
}

int fqss_conv1x1_wgrad(const float* gy, int64_t ldgy, const float* x, int64_t ldx, float* gw, float* gbias, int B, int Ci,
                       int Co, int M, void* ws, size_t ws_bytes, void* stream) {
    FQSS_REQUIRE(gy && x && gw && B > 0 && Ci > 0 && Co > 0 && M > 0 && ldgy >= M && ldx >= M, -1, "conv1x1_wgrad: bad argument");
    FQSS_REQUIRE(!gbias || (ws && ws_bytes >= (size_t)Co * sizeof(double)), -3, "conv1x1_wgrad: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    FQSS_PROFN("conv1x1_wgrad(fp32)", s, gbias ? 3 : 1);
    cudaMemsetAsync(gw, 0, (size_t)Co * Ci * sizeof(float), s);
    const int kchunk = 1024;
    const int nchunk = (M + kchunk - 1) / kchunk;
    GemmArgs g{gy, ldgy, (int64_t)Co * ldgy, x, ldx, (int64_t)Ci * ldx, gw, Ci, 0, nullptr, Co, Ci, M, kchunk};
    dim3 grid((Ci + GB - 1) / GB, (Co + GB - 1) / GB, B * nchunk);
    sgemm_kernel<2><<<grid, 256, 0, s>>>(g);
    if (gbias) {
        cudaMemsetAsync(ws, 0, (size_t)Co * sizeof(double), s);
        rowsum_kernel<<<B * Co, 256, 0, s>>>(gy, M, ldgy, Co, (double*)ws);
        f64_store_kernel<<<(Co + 255) / 256, 256, 0, s>>>((const double*)ws, gbias, Co);
    }
    return check_launch("conv1x1_wgrad");
}

int fqss_dwconv_fwd(const float* x, int64_t ldx, const float* w, const float* bias, float* y, int64_t ldy, int B, int C, int M,
                    int K, int dil, void* stream) {
    FQSS_REQUIRE(x && w && y && B > 0 && C > 0 && M > 0 && K >= 1 && K <= DW_MAXK && (K & 1) && dil >= 1 && ldx >= M && ldy >= M,
                 -1, "dwconv_fwd: bad argument (odd K <= %d)", DW_MAXK);
    FQSS_PROF("dwconv_fwd(layer)", stream);
    if (K == 3 && ldx % 4 == 0 && ldy % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0) {
        dim3 grid(B * C, ((M + 3) / 4 + 255) / 256);
        dwconv3_fwd_vec_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, ldx, w, bias, y, ldy, C, M, dil);
    } else {
        dim3 grid(B * C, (M + 255) / 256);
        dwconv_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, ldx, w, bias, y, ldy, C, M, K, dil);
    }
    return check_launch("dwconv_fwd");
}

int fqss_dwconv_bwd(const float* gy, int64_t ldgy, const float* x, int64_t ldx, const float* w, float* gx, int64_t ldgx,
                    float* gw, float* gbias, int B, int C, int M, int K, int dil, void* ws, size_t ws_bytes, void* stream) {
    FQSS_REQUIRE(gy && x && w && B > 0 && C > 0 && M > 0 && K >= 1 && K <= DW_MAXK && (K & 1) && dil >= 1, -1,
                 "dwconv_bwd: bad argument");
    size_t need = (size_t)C * (DW_MAXK + 1) * sizeof(double);
    FQSS_REQUIRE(ws && ws_bytes >= need, -3, "dwconv_bwd: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    FQSS_PROFN("dwconv_bwd(layer)", s, 2);
    cudaMemsetAsync(ws, 0, need, s);
    const bool a16 = (((uintptr_t)gy | (uintptr_t)x | (uintptr_t)gx) & 15) == 0 && ldgy % 4 == 0 && ldx % 4 == 0 && (!gx || ldgx % 4 == 0);
    if (K == 3 && a16) dwconv3_bwd_vec_kernel<<<B * C, 256, 0, s>>>(gy, ldgy, x, ldx, w, gx, ldgx, C, M, dil, (double*)ws);
    else dwconv_bwd_kernel<<<B * C, 256, 0, s>>>(gy, ldgy, x, ldx, w, gx, ldgx, C, M, K, dil, (double*)ws);
    dwconv_finalize_kernel<<<(C + 127) / 128, 128, 0, s>>>((const double*)ws, gw, gbias, C, K);
    return check_launch("dwconv_bwd");
}

int fqss_sconv_fwd(const float* x, int64_t ldx, const float* w, float* y, int64_t ldy, int B, int Cin, int Co, int T, int K,
                   int stride, void* stream) {
    FQSS_REQUIRE(x && w && y && B > 0 && Cin > 0 && Co > 0 && T >= K && K > 0 && stride > 0 && Cin * K <= SC_MAXCK, -1,
                 "sconv_fwd: bad argument (Cin*K <= %d)", SC_MAXCK);
    const int Mo = (T - K) / stride + 1;
    FQSS_REQUIRE(ldx >= T && ldy >= Mo, -1, "sconv_fwd: bad pitch");
    FQSS_PROF("sconv_fwd", stream);
    if (edge_geometry_ok(K, stride) && aligned16(y) && edge_analysis_fwd(x, ldx, T, w, y, ldy, B, Cin, Co, Mo, (cudaStream_t)stream) == 0)
        return check_launch("sconv_fwd(tiled)");
    dim3 grid((Mo + 127) / 128, (Co + SC_OT - 1) / SC_OT, B);
    sconv_fwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x, ldx, w, y, ldy, Cin, Co, Mo, K, stride);
    return check_launch("sconv_fwd");
}

int fqss_sconv_bwd(const float* gy, int64_t ldgy, const float* x, int64_t ldx, const float* w, float* gx, int64_t ldgx,
                   float* gw, int B, int Cin, int Co, int T, int K, int stride, void* ws, size_t ws_bytes, void* stream) {
    FQSS_REQUIRE(gy && x && w && B > 0 && Cin > 0 && Co > 0 && T >= K && K > 0 && stride > 0, -1, "sconv_bwd: bad argument");
    const int Mo = (T - K) / stride + 1;
    cudaStream_t s = (cudaStream_t)stream;
    if (gx) {
        FQSS_PROF("sconv_bwd(dx: synthesis)", s);
        // gx[b,c,t] = sum_o sum_{m,k} w[o,c,k] gy[b,o,m]: the overlap-add kernel with weight row stride Cin*K
        FQSS_REQUIRE(Cin == 1, -1, "sconv_bwd: input gradient implemented for Cin == 1 (RQB re-encoder)");
        // samples beyond the last frame's support receive no gradient
        const int Tc = (Mo - 1) * stride + K;
        if (Tc < T) cudaMemset2DAsync(gx + Tc, (size_t)ldgx * sizeof(float), 0, (size_t)(T - Tc) * sizeof(float), (size_t)B, s);
        if (!(edge_geometry_ok(K, stride) && edge_synthesis_fwd(gy, ldgy, w, (int64_t)Cin * K, gx, ldgx, B, Co, Mo, Tc, s) == 0)) {
            dim3 grid((T + 255) / 256, B);
            tconv_fwd_kernel<<<grid, 256, 0, s>>>(gy, ldgy, w, (int64_t)Cin * K, gx, ldgx, Co, Mo, K, stride, T);
        }
    }
    if (gw) {
        size_t need = (size_t)Co * Cin * K * sizeof(double);
        FQSS_REQUIRE(ws && ws_bytes >= need, -3, "sconv_bwd: workspace too small");
        FQSS_PROFN("sconv_bwd(dw: edge_wgrad)", s, 2);
        cudaMemsetAsync(ws, 0, need, s);
        if (!(edge_geometry_ok(K, stride) && edge_wgrad(gy, ldgy, x, ldx, T, B, Cin, Co, Mo, (double*)ws, s) == 0)) {
            const int CK = Cin * K;
            const size_t smem = ((size_t)Cin * ((GW_MF - 1) * stride + K) + (size_t)GW_OT * GW_MF) * sizeof(float);
            if (CK <= SC_MAXCK && smem <= 160 * 1024 && (int64_t)B * Mo >= 4096) {
                static bool cfg = false;
                if (!cfg) { cudaFuncSetAttribute(sconv_gw_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); cfg = true; }
                dim3 grid((Mo + GW_MF - 1) / GW_MF, (Co + GW_OT - 1) / GW_OT, B);
                sconv_gw_tile_kernel<<<grid, 4 * CK, smem, s>>>(gy, ldgy, x, ldx, T, Cin, Co, Mo, K, stride, (double*)ws);
            } else {
                dim3 grid((Mo + 1023) / 1024, Co, B);
                sconv_gw_kernel<<<grid, 256, 0, s>>>(gy, ldgy, x, ldx, Cin, Co, Mo, K, stride, (double*)ws);
            }
        }
        int n = Co * Cin * K;
        f64_store_kernel<<<(n + 255) / 256, 256, 0, s>>>((const double*)ws, gw, n);
    }
    return check_launch("sconv_bwd");
}

int fqss_tconv_fwd(const float* x, int64_t ldx, const float* w, float* y, int64_t ldy, int B, int Ci, int M, int K, int stride,
                   void* stream) {
    FQSS_REQUIRE(x && w && y && B > 0 && Ci > 0 && M > 0 && K > 0 && stride > 0 && ldx >= M, -1, "tconv_fwd: bad argument");
    const int T = (M - 1) * stride + K;
    FQSS_REQUIRE(ldy >= T, -1, "tconv_fwd: bad output pitch");
    FQSS_PROF("tconv_fwd", stream);
    if (edge_geometry_ok(K, stride) && edge_synthesis_fwd(x, ldx, w, (int64_t)K, y, ldy, B, Ci, M, T, (cudaStream_t)stream) == 0)
        return check_launch("tconv_fwd(tiled)");
    dim3 grid((T + 255) / 256, B);
    tconv_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, ldx, w, (int64_t)K, y, ldy, Ci, M, K, stride, T);
    return check_launch("tconv_fwd");
}

int fqss_tconv_bwd(const float* gy, int64_t ldgy, const float* x, int64_t ldx, const float* w, float* gx, int64_t ldgx,
                   float* gw, int B, int Ci, int M, int K, int stride, void* ws, size_t ws_bytes, void* stream) {
    FQSS_REQUIRE(gy && x && w && B > 0 && Ci > 0 && M > 0 && K > 0 && stride > 0 && K <= SC_MAXCK, -1, "tconv_bwd: bad argument");
    const int T = (M - 1) * stride + K;
    cudaStream_t s = (cudaStream_t)stream;
    if (gx) {   // gx[b,c,m] = sum_k w[c,k] gy[b, m*stride+k]  == analysis conv with Cin=1, Co=Ci
        FQSS_PROF("tconv_bwd(dx: analysis)", s);
        if (!(edge_geometry_ok(K, stride) && aligned16(gx) && edge_analysis_fwd(gy, ldgy, T, w, gx, ldgx, B, 1, Ci, M, s) == 0)) {
            dim3 grid((M + 127) / 128, (Ci + SC_OT - 1) / SC_OT, B);
            sconv_fwd_kernel<<<grid, 128, 0, s>>>(gy, ldgy, w, gx, ldgx, 1, Ci, M, K, stride);
        }
    }
    if (gw) {   // gw[c,k] = sum_{b,m} x[b,c,m] gy[b, m*stride+k]
        size_t need = (size_t)Ci * K * sizeof(double);
        FQSS_REQUIRE(ws && ws_bytes >= need, -3, "tconv_bwd: workspace too small");
        FQSS_PROFN("tconv_bwd(dw: edge_wgrad)", s, 2);
        cudaMemsetAsync(ws, 0, need, s);
        if (!(edge_geometry_ok(K, stride) && edge_wgrad(x, ldx, gy, ldgy, T, B, 1, Ci, M, (double*)ws, s) == 0)) {
            dim3 grid((M + 1023) / 1024, Ci, B);
            sconv_gw_kernel<<<grid, 256, 0, s>>>(x, ldx, gy, ldgy, 1, Ci, M, K, stride, (double*)ws);
        }
        int n = Ci * K;
        f64_store_kernel<<<(n + 255) / 256, 256, 0, s>>>((const double*)ws, gw, n);
    }
    (void)T;
    return check_launch("tconv_bwd");
}

}  // extern "C"
