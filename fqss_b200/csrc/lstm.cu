// lstm.cu -- the recurrence of the quantised LSTM (LSTMQ, reference qat_layers.py:571-613: nn.LSTM evaluated with
// fake-quantised weights; used by DPTNetQ's improved transformer layer, dptnetq.py:57-97) as two persistent kernels that
// exploit what fake quantisation leaves of the recurrent weight: W_hh[j, :] = delta_j * code[j, :] with 8-bit integer codes.
//
//   forward   one CTA = 4 or 8 sequences x one direction, one thread per gate row j (4H threads).  The thread keeps ITS ROW OF
//             CODES PACKED IN REGISTERS (H/4 words) for all T steps, so the recurrent matrix-vector product reads no weight
//             from memory at all: per step  pre[n, j] = gx[t, n, j] + delta_j * sum_k code[j, k] h[n, k]  with h in shared
//             memory (broadcast float4 reads), then the gate nonlinearities, c = f c + i g, h = o tanh(c).  The activated
//             gates and c are saved for backward.  gx = x W_ih^T + b_ih + b_hh for all steps is one batched GEMM outside.
//   backward  the same decomposition run in reverse: per step the element-wise gate derivatives give the pre-activation
//             gradients dG[t] (stored: they are the gradient of gx and the operand of the batched weight-gradient GEMMs),
//             and dh_{t-1} = dG[t] (delta * code) is accumulated from the TRANSPOSED codes, again register-resident
//             (thread (q, k) holds column k of the q-th quarter of the rows).
//
// The codes and steps are produced inside the kernels by the same device functions as the weight quantiser kernel
// (make_wq / wq_code, fqss_common.cuh), so the weights used here are bit-identical to FakeQuantWeight's output.
#include "fqss_common.cuh"

namespace fqss {


__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float s8f(uint32_t w, int i) { return (float)(int)(int8_t)(w >> (8 * i)); }

template <int H, int LNB>
__global__ void __launch_bounds__(4 * H) lstm_rec_fwd_kernel(const float* __restrict__ gx, const float* __restrict__ whh0,
                                                             const float* __restrict__ whh1, const float* wmin0, const float* wmin1,
                                                             const float* wmax0, const float* wmax1, float* __restrict__ out,
                                                             float* __restrict__ gates, float* __restrict__ cseq, int T, int N, int D) {
    constexpr int G = 4 * H;
    __shared__ __align__(16) float hs[LNB][H];
    __shared__ float ga[LNB][G];
    const int d = blockIdx.y, n0 = blockIdx.x * LNB, j = threadIdx.x;
    const float* W = d ? whh1 : whh0;
    const WQ wq = make_wq(__ldg((d ? wmin1 : wmin0) + j), __ldg((d ? wmax1 : wmax0) + j), 8);
    uint32_t wp[H / 4];
#pragma unroll
    for (int kq = 0; kq < H / 4; ++kq) {
        uint32_t w = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) w |= ((uint32_t)(int)wq_code(wq, __ldg(W + (size_t)j * H + 4 * kq + i)) & 0xffu) << (8 * i);
        wp[kq] = w;
    }
    const float sc = wq.delta;
    for (int i = j; i < LNB * H; i += G) (&hs[0][0])[i] = 0.f;
    __syncthreads();
    const int pn = j / H, pk = j % H;          // this thread owns the cell states (pn + 4r, pk), r < LNB/4; its gate row is of type j / H
    float c_reg[LNB / 4];
#pragma unroll
    for (int r = 0; r < LNB / 4; ++r) c_reg[r] = 0.f;
    for (int tt = 0; tt < T; ++tt) {
        const int t = d ? T - 1 - tt : tt;
        const size_t row0 = ((size_t)d * T + t) * N + n0;
        float pre[LNB];
#pragma unroll
        for (int n = 0; n < LNB; ++n) pre[n] = (n0 + n < N) ? __ldg(gx + (row0 + n) * G + j) : 0.f;
        float acc[LNB];
#pragma unroll
        for (int n = 0; n < LNB; ++n) acc[n] = 0.f;
#pragma unroll
        for (int kq = 0; kq < H / 4; ++kq) {
            uint32_t w = wp[kq];
            asm volatile("" : "+r"(w));      // keep the int8 -> float decode inside the step loop (hoisted, it would need 4x the registers)
            const float w0 = s8f(w, 0), w1 = s8f(w, 1), w2 = s8f(w, 2), w3 = s8f(w, 3);
#pragma unroll
            for (int n = 0; n < LNB; ++n) {
                const float4 hv = *reinterpret_cast<const float4*>(&hs[n][4 * kq]);
                acc[n] = fmaf(w0, hv.x, fmaf(w1, hv.y, fmaf(w2, hv.z, fmaf(w3, hv.w, acc[n]))));
            }
        }
#pragma unroll
        for (int n = 0; n < LNB; ++n) {
            const float p = fmaf(sc, acc[n], pre[n]);
            const float a = (pn == 2) ? tanhf(p) : sigmoidf_(p);
            ga[n][j] = a;
            if (n0 + n < N) gates[(row0 + n) * G + j] = a;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < LNB / 4; ++r) {
            const int n = pn + 4 * r;
            const float ig = ga[n][pk], fg = ga[n][H + pk], gg = ga[n][2 * H + pk], og = ga[n][3 * H + pk];
            const float c = fg * c_reg[r] + ig * gg;
            c_reg[r] = c;
            const float h = og * tanhf(c);
            hs[n][pk] = h;
            if (n0 + n < N) {
                out[((size_t)t * N + n0 + n) * (D * H) + d * H + pk] = h;
                cseq[(row0 + n) * H + pk] = c;
            }
        }
        __syncthreads();
    }
}

template <int H, int LNB>
__global__ void __launch_bounds__(4 * H) lstm_rec_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ gates,
                                                             const float* __restrict__ cseq, const float* __restrict__ whh0,
                                                             const float* __restrict__ whh1, const float* wmin0, const float* wmin1,
                                                             const float* wmax0, const float* wmax1, float* __restrict__ dG, int T, int N, int D) {
    constexpr int G = 4 * H;
    __shared__ __align__(16) float dgs[LNB][G];        // pre-activation gradients x row step (operand of the transposed product)
    __shared__ float part[4][LNB][H];                   // dh_{t-1} as four partial sums (one per quarter of the gate rows)
    __shared__ float scl[G];
    const int d = blockIdx.y, n0 = blockIdx.x * LNB, tid = threadIdx.x;
    const int q = tid / H, k = tid % H;
    const float* W = d ? whh1 : whh0;
    const float* wmn = d ? wmin1 : wmin0;
    const float* wmx = d ? wmax1 : wmax0;
    uint32_t wT[H / 4];
#pragma unroll
    for (int jq = 0; jq < H / 4; ++jq) {
        uint32_t w = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int jrow = q * H + 4 * jq + i;
            const WQ wqj = make_wq(__ldg(wmn + jrow), __ldg(wmx + jrow), 8);
            w |= ((uint32_t)(int)wq_code(wqj, __ldg(W + (size_t)jrow * H + k)) & 0xffu) << (8 * i);
        }
        wT[jq] = w;
    }
    scl[tid] = make_wq(__ldg(wmn + tid), __ldg(wmx + tid), 8).delta;
    for (int i = tid; i < 4 * LNB * H; i += G) (&part[0][0][0])[i] = 0.f;
    __syncthreads();
    float dc_reg[LNB / 4];
#pragma unroll
    for (int r = 0; r < LNB / 4; ++r) dc_reg[r] = 0.f;
    for (int tt = T - 1; tt >= 0; --tt) {
        const int t = d ? T - 1 - tt : tt;
        const int tprev = d ? t + 1 : t - 1;
#pragma unroll
        for (int r = 0; r < LNB / 4; ++r) {
            const int n = q + 4 * r, nn = n0 + n;
            float dpi = 0.f, dpf = 0.f, dpg = 0.f, dpo = 0.f;
            if (nn < N) {
                const size_t row = ((size_t)d * T + t) * N + nn;
                const float dh = __ldg(dout + ((size_t)t * N + nn) * (D * H) + d * H + k) + part[0][n][k] + part[1][n][k] + part[2][n][k] +
                                 part[3][n][k];
                const float ig = __ldg(gates + row * G + k), fg = __ldg(gates + row * G + H + k), gg = __ldg(gates + row * G + 2 * H + k),
                            og = __ldg(gates + row * G + 3 * H + k);
                const float c = __ldg(cseq + row * H + k);
                const float cprev = tt > 0 ? __ldg(cseq + (((size_t)d * T + tprev) * N + nn) * H + k) : 0.f;
                const float tc = tanhf(c);
                const float dct = dc_reg[r] + dh * og * (1.f - tc * tc);
                dpi = dct * gg * ig * (1.f - ig);
                dpf = dct * cprev * fg * (1.f - fg);
                dpg = dct * ig * (1.f - gg * gg);
                dpo = dh * tc * og * (1.f - og);
                dc_reg[r] = dct * fg;
                dG[row * G + k] = dpi;
                dG[row * G + H + k] = dpf;
                dG[row * G + 2 * H + k] = dpg;
                dG[row * G + 3 * H + k] = dpo;
            }
            dgs[n][k] = dpi * scl[k];
            dgs[n][H + k] = dpf * scl[H + k];
            dgs[n][2 * H + k] = dpg * scl[2 * H + k];
            dgs[n][3 * H + k] = dpo * scl[3 * H + k];
        }
        __syncthreads();
        float acc[LNB];
#pragma unroll
        for (int n = 0; n < LNB; ++n) acc[n] = 0.f;
#pragma unroll
        for (int jq = 0; jq < H / 4; ++jq) {
            uint32_t w = wT[jq];
            asm volatile("" : "+r"(w));
            const float w0 = s8f(w, 0), w1 = s8f(w, 1), w2 = s8f(w, 2), w3 = s8f(w, 3);
#pragma unroll
            for (int n = 0; n < LNB; ++n) {
                const float4 dv = *reinterpret_cast<const float4*>(&dgs[n][q * H + 4 * jq]);
                acc[n] = fmaf(w0, dv.x, fmaf(w1, dv.y, fmaf(w2, dv.z, fmaf(w3, dv.w, acc[n]))));
            }
        }
#pragma unroll
        for (int n = 0; n < LNB; ++n) part[q][n][k] = acc[n];
        __syncthreads();
    }
}

}  // namespace fqss

using namespace fqss;

extern "C" {

int fqss_lstm_rec_fwd(const float* gx, const float* whh0, const float* whh1, const float* wmin0, const float* wmin1, const float* wmax0,
                      const float* wmax1, float* out, float* gates, float* cseq, int T, int N, int H, int D, void* stream) {
    FQSS_REQUIRE(gx && whh0 && wmin0 && wmax0 && out && gates && cseq && T > 0 && N > 0 && (D == 1 || (D == 2 && whh1 && wmin1 && wmax1)), -1,
                 "lstm_rec_fwd: bad argument");
    FQSS_REQUIRE(H == 32 || H == 64 || H == 128, -1, "lstm_rec_fwd: hidden size %d has no kernel (32, 64, 128)", H);
    cudaStream_t s = (cudaStream_t)stream;
    // sequences per CTA: 8, or 4 when that still leaves SMs idle (the recurrence is latency-bound: more, lighter CTAs win)
    const int nb = ((N + 7) / 8) * D < 100 ? 4 : 8;
    const dim3 grid((N + nb - 1) / nb, D);
    FQSS_PROF("lstm_rec_fwd", s);
#define FQSS_LSTM_F(HH, NB) lstm_rec_fwd_kernel<HH, NB><<<grid, 4 * HH, 0, s>>>(gx, whh0, whh1, wmin0, wmin1, wmax0, wmax1, out, gates, cseq, T, N, D)
    if (H == 32) { if (nb == 4) FQSS_LSTM_F(32, 4); else FQSS_LSTM_F(32, 8); }
    else if (H == 64) { if (nb == 4) FQSS_LSTM_F(64, 4); else FQSS_LSTM_F(64, 8); }
    else { if (nb == 4) FQSS_LSTM_F(128, 4); else FQSS_LSTM_F(128, 8); }
#undef FQSS_LSTM_F
    return check_launch("lstm_rec_fwd");
}

int fqss_lstm_rec_bwd(const float* dout, const float* gates, const float* cseq, const float* whh0, const float* whh1, const float* wmin0,
                      const float* wmin1, const float* wmax0, const float* wmax1, float* dG, int T, int N, int H, int D, void* stream) {
    FQSS_REQUIRE(dout && gates && cseq && whh0 && wmin0 && wmax0 && dG && T > 0 && N > 0 && (D == 1 || (D == 2 && whh1 && wmin1 && wmax1)), -1,
                 "lstm_rec_bwd: bad argument");
    FQSS_REQUIRE(H == 32 || H == 64 || H == 128, -1, "lstm_rec_bwd: hidden size %d has no kernel (32, 64, 128)", H);
    cudaStream_t s = (cudaStream_t)stream;
    const int nb = ((N + 7) / 8) * D < 100 ? 4 : 8;
    const dim3 grid((N + nb - 1) / nb, D);
    FQSS_PROF("lstm_rec_bwd", s);
#define FQSS_LSTM_B(HH, NB) lstm_rec_bwd_kernel<HH, NB><<<grid, 4 * HH, 0, s>>>(dout, gates, cseq, whh0, whh1, wmin0, wmin1, wmax0, wmax1, dG, T, N, D)
    if (H == 32) { if (nb == 4) FQSS_LSTM_B(32, 4); else FQSS_LSTM_B(32, 8); }
    else if (H == 64) { if (nb == 4) FQSS_LSTM_B(64, 4); else FQSS_LSTM_B(64, 8); }
    else { if (nb == 4) FQSS_LSTM_B(128, 4); else FQSS_LSTM_B(128, 8); }
#undef FQSS_LSTM_B
    return check_launch("lstm_rec_bwd");
}

}  // extern "C"
