// music.cu -- the pieces ConvTasNetMusicQ needs beyond the speech path (SURVEY.md 8f rank 1,
// quantization/qat/models/convtasnetq_music.py:10-50, 236-275; process.py:16-36 with normalize=False):
//
//   fqss_split_ex       FQSS splitter for multi-channel input, with or without the peak normalisation
//   fqss_cln_fwd/bwd    channel-wise LayerNorm (cLN): nn.LayerNorm(C) over the channel axis of an NCL tensor, per frame
//   fqss_ola_fwd/bwd    overlap-and-add of the per-frame outputs of the Linear decoder (index_add of frames, hop < length)
//
// All three are small HBM passes over tensors that occur once per forward; threads run along the contiguous frame /
// sample axis so every access is coalesced.
#include "fqss_common.cuh"

namespace fqss {

int num_sms();

static inline int grid_1d(int64_t items, int threads) {
    int64_t need = (items + threads - 1) / threads;
    int64_t cap = (int64_t)num_sms() * 16;
    return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

// ---------------------------------------------------------------------------------------------
// splitter (process.py:16-36).  normalize: x /= peak, threshold = 1; otherwise threshold = peak.  Every op is a separately
// rounded fp32 operation in the reference's order:
//   delta = threshold / 128;  q = clip(floor(x / delta), -128, 127) * delta;  x <- 2 * (x - q) * threshold / delta - threshold
// out[b, s*C + c, t] = s-th quantised part of channel c (torch.cat over dim 1).
// ---------------------------------------------------------------------------------------------
__global__ void split_ex_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ peak, float* __restrict__ y,
                                int64_t ldy, int B, int C, int T, int n_split, int n_bits, int normalize) {
    const float pk = __ldg(peak);
    const float half = (float)(1 << (n_bits - 1));
    const float thr = normalize ? 1.f : pk;
    const float step = __fdiv_rn(thr, half);
    const float lo = -half, hi = half - 1.f;
    const int64_t total = (int64_t)B * C * T;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / T;
        const int t = (int)(i - row * T);
        const int b = (int)(row / C), c = (int)(row - (int64_t)b * C);
        float v = x[row * ldx + t];
        if (normalize) v = __fdiv_rn(v, pk);
        for (int s = 0; s < n_split; ++s) {
            const float qv = __fmul_rn(fminf(fmaxf(floorf(__fdiv_rn(v, step)), lo), hi), step);
            y[((int64_t)b * n_split * C + (int64_t)s * C + c) * ldy + t] = qv;
            v = __fsub_rn(__fdiv_rn(__fmul_rn(__fmul_rn(2.f, __fsub_rn(v, qv)), thr), step), thr);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// cLN forward: per (sample, frame) mean / rstd over the C channels, y = (x - mean) * rstd * gamma_c + beta_c.
// One thread per frame (coalesced along frames), three sweeps over the channel axis (mean, centred variance, output):
// the second and third sweeps hit L2.  mean / rstd [B, M] are saved for backward.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) cln_fwd_kernel(const float* __restrict__ x, int64_t ld, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float eps, float* __restrict__ y, int64_t ldy,
                                                     float* __restrict__ mean, float* __restrict__ rstd, int C, int M) {
    const int b = blockIdx.y;
    const int m = blockIdx.x * 128 + threadIdx.x;
    if (m >= M) return;
    const float* xb = x + (int64_t)b * C * ld + m;
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += __ldg(xb + (int64_t)c * ld);
    const float mu = s / (float)C;
    float v = 0.f;
    for (int c = 0; c < C; ++c) {
        const float d = __ldg(xb + (int64_t)c * ld) - mu;
        v = fmaf(d, d, v);
    }
    const float r = 1.f / sqrtf(v / (float)C + eps);
    float* yb = y + (int64_t)b * C * ldy + m;
    for (int c = 0; c < C; ++c) yb[(int64_t)c * ldy] = fmaf((__ldg(xb + (int64_t)c * ld) - mu) * r, __ldg(gamma + c), __ldg(beta + c));
    mean[(int64_t)b * M + m] = mu;
    rstd[(int64_t)b * M + m] = r;
}

// cLN backward, input gradient: per frame  gx_c = rstd * (gamma_c g_c - m1 - xhat_c m2),  m1 = mean_c(gamma g), m2 = mean_c(gamma g xhat)
__global__ void __launch_bounds__(128) cln_bwd_dx_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ x, int64_t ld,
                                                        const float* __restrict__ gamma, const float* __restrict__ mean,
                                                        const float* __restrict__ rstd, float* __restrict__ gx, int64_t ldgx, int C,
                                                        int M) {
    const int b = blockIdx.y;
    const int m = blockIdx.x * 128 + threadIdx.x;
    if (m >= M) return;
    const float mu = mean[(int64_t)b * M + m], r = rstd[(int64_t)b * M + m];
    const float* xb = x + (int64_t)b * C * ld + m;
    const float* gb = g + (int64_t)b * C * ldg + m;
    float s1 = 0.f, s2 = 0.f;
    for (int c = 0; c < C; ++c) {
        const float gg = __ldg(gb + (int64_t)c * ldg) * __ldg(gamma + c);
        const float xh = (__ldg(xb + (int64_t)c * ld) - mu) * r;
        s1 += gg;
        s2 = fmaf(gg, xh, s2);
    }
    const float m1 = s1 / (float)C, m2 = s2 / (float)C;
    float* ob = gx + (int64_t)b * C * ldgx + m;
    for (int c = 0; c < C; ++c) {
        const float gg = __ldg(gb + (int64_t)c * ldg) * __ldg(gamma + c);
        const float xh = (__ldg(xb + (int64_t)c * ld) - mu) * r;
        ob[(int64_t)c * ldgx] = r * (gg - m1 - xh * m2);
    }
}

// cLN backward, affine gradients: one CTA per (sample, channel) row; dgamma_c += sum_m g xhat, dbeta_c += sum_m g (fp64 atomics)
__global__ void __launch_bounds__(256) cln_bwd_affine_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ x,
                                                            int64_t ld, const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            double* __restrict__ acc, int C, int M) {
    __shared__ double sh[2 * 32];
    const int64_t row = blockIdx.x;
    const int b = (int)(row / C), c = (int)(row % C);
    const float* gr = g + row * ldg;
    const float* xr = x + row * ld;
    const float* mu = mean + (int64_t)b * M;
    const float* rs = rstd + (int64_t)b * M;
    float sg = 0.f, sb = 0.f;
    for (int m = threadIdx.x; m < M; m += 256) {
        const float gg = __ldg(gr + m);
        sb += gg;
        sg = fmaf(gg, (__ldg(xr + m) - __ldg(mu + m)) * __ldg(rs + m), sg);
    }
    double v[2] = {(double)sg, (double)sb};
    block_sum<2>(v, sh);
    if (threadIdx.x == 0) {
        atomicAdd(acc + 2 * c, v[0]);
        atomicAdd(acc + 2 * c + 1, v[1]);
    }
}

__global__ void cln_affine_store_kernel(const double* __restrict__ acc, float* __restrict__ ggamma, float* __restrict__ gbeta, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) {
        ggamma[c] = (float)acc[2 * c];
        gbeta[c] = (float)acc[2 * c + 1];
    }
}

// ---------------------------------------------------------------------------------------------
// overlap-and-add (convtasnetq_music.py:10-30): frames of length L with hop H from rows [R, A*L, K] (row = (r, a, j), column
// = frame k) to out[r, a, t], t = k*H + j; contributions are added in increasing frame order, as index_add_ does.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ola_fwd_kernel(const float* __restrict__ y, int64_t ldy, float* __restrict__ out, int64_t ldo,
                                                     int A, int L, int H, int K, int T) {
    const int64_t ra = blockIdx.y;                 // (r, a)
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= T) return;
    const int k_hi = min(t / H, K - 1);
    const int k_lo = max(0, (t - L + H) / H);      // smallest k with t - k*H <= L-1
    const int64_t r = ra / A;
    const int a = (int)(ra - r * A);
    const float* yr = y + (r * A * L + (int64_t)a * L) * ldy;
    float acc = 0.f;
    for (int k = k_lo; k <= k_hi; ++k) acc = __fadd_rn(acc, __ldg(yr + (int64_t)(t - k * H) * ldy + k));
    out[ra * ldo + t] = acc;
}

__global__ void __launch_bounds__(256) ola_bwd_kernel(const float* __restrict__ go, int64_t ldo, float* __restrict__ gy, int64_t ldy,
                                                     int A, int L, int H, int K) {
    const int64_t row = blockIdx.y;                // (r, a, j)
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= K) return;
    const int64_t ra = row / L;
    const int j = (int)(row - ra * L);
    gy[row * ldy + k] = __ldg(go + ra * ldo + (int64_t)k * H + j);
}

}  // namespace fqss

using namespace fqss;

extern "C" {

int fqss_split_ex(const float* x, int64_t ldx, const float* peak, float* y, int64_t ldy, int B, int C, int T, int n_split, int n_bits,
                  int normalize, void* stream) {
    FQSS_REQUIRE(x && peak && y && B > 0 && C > 0 && T > 0 && n_split >= 1 && ldx >= T && ldy >= T, -1, "split_ex: bad argument");
    FQSS_PROF("split", stream);
    split_ex_kernel<<<grid_1d((int64_t)B * C * T, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, peak, y, ldy, B, C, T, n_split, n_bits,
                                                                                       normalize);
    return check_launch("split_ex");
}

int fqss_cln_fwd(const float* x, int64_t ld, const float* gamma, const float* beta, float eps, float* y, int64_t ldy, float* mean,
                 float* rstd, int B, int C, int M, void* stream) {
    FQSS_REQUIRE(x && gamma && beta && y && mean && rstd && B > 0 && C > 0 && M > 0 && ld >= M && ldy >= M, -1, "cln_fwd: bad argument");
    FQSS_PROF("cln_fwd", stream);
    cln_fwd_kernel<<<dim3((M + 127) / 128, B), 128, 0, (cudaStream_t)stream>>>(x, ld, gamma, beta, eps, y, ldy, mean, rstd, C, M);
    return check_launch("cln_fwd");
}

int fqss_cln_bwd(const float* g, int64_t ldg, const float* x, int64_t ld, const float* gamma, const float* mean, const float* rstd,
                 float* gx, int64_t ldgx, float* ggamma, float* gbeta, int B, int C, int M, void* ws, size_t ws_bytes, void* stream) {
    FQSS_REQUIRE(g && x && gamma && mean && rstd && B > 0 && C > 0 && M > 0 && ldg >= M && ld >= M, -1, "cln_bwd: bad argument");
    FQSS_REQUIRE(!gx || ldgx >= M, -1, "cln_bwd: bad pitch");
    cudaStream_t s = (cudaStream_t)stream;
    if (gx) {
        FQSS_PROF("cln_bwd(dx)", s);
        cln_bwd_dx_kernel<<<dim3((M + 127) / 128, B), 128, 0, s>>>(g, ldg, x, ld, gamma, mean, rstd, gx, ldgx, C, M);
    }
    if (ggamma || gbeta) {
        FQSS_REQUIRE(ggamma && gbeta, -1, "cln_bwd: the two affine gradients come together");
        const size_t need = (size_t)2 * C * sizeof(double);
        FQSS_REQUIRE(ws && ws_bytes >= need, -3, "cln_bwd: workspace too small");
        FQSS_PROFN("cln_bwd(affine)", s, 2);
        cudaMemsetAsync(ws, 0, need, s);
        cln_bwd_affine_kernel<<<(unsigned)((int64_t)B * C), 256, 0, s>>>(g, ldg, x, ld, mean, rstd, (double*)ws, C, M);
        cln_affine_store_kernel<<<(C + 127) / 128, 128, 0, s>>>((const double*)ws, ggamma, gbeta, C);
    }
    return check_launch("cln_bwd");
}

int fqss_ola_fwd(const float* y, int64_t ldy, float* out, int64_t ldo, int64_t R, int A, int L, int H, int K, void* stream) {
    FQSS_REQUIRE(y && out && R > 0 && A > 0 && L > 0 && H > 0 && K > 0 && ldy >= K, -1, "ola_fwd: bad argument");
    const int T = (K - 1) * H + L;
    FQSS_REQUIRE(ldo >= T, -1, "ola_fwd: bad output pitch");
    FQSS_PROF("ola_fwd", stream);
    ola_fwd_kernel<<<dim3((T + 255) / 256, (unsigned)(R * A)), 256, 0, (cudaStream_t)stream>>>(y, ldy, out, ldo, A, L, H, K, T);
    return check_launch("ola_fwd");
}

int fqss_ola_bwd(const float* gout, int64_t ldo, float* gy, int64_t ldy, int64_t R, int A, int L, int H, int K, void* stream) {
    FQSS_REQUIRE(gout && gy && R > 0 && A > 0 && L > 0 && H > 0 && K > 0 && ldy >= K, -1, "ola_bwd: bad argument");
    FQSS_REQUIRE(ldo >= (K - 1) * H + L, -1, "ola_bwd: bad pitch");
    FQSS_PROF("ola_bwd", stream);
    ola_bwd_kernel<<<dim3((K + 255) / 256, (unsigned)(R * A * L)), 256, 0, (cudaStream_t)stream>>>(gout, ldo, gy, ldy, A, L, H, K);
    return check_launch("ola_bwd");
}

}  // extern "C"
