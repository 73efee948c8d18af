// loss.cu -- FQSS knowledge-distillation SI-SDR loss as a fused segmented reduction, n_src = 2.
//
// Reference: System.common_step (train_env/asteroid_librimix/mysystem.py:124-146), PairwiseWSDR
// (wsdr.py:46-95) and asteroid 0.6 PITLossWrapper(pit_from="pw_mtx") with the factorial search.
//
//   rho[b,i,j] = |alpha t_j|^2 / (|e_i - alpha t_j|^2 + eps), alpha = <e_i,t_j>/(|t_j|^2+eps), zero-mean signals
//   sdr_x[b]   = min_perm mean_j -10 log10(rho_x[b,perm(j),j] + eps)            x in {teacher, student} vs target
//   w[b]       = 10^((sdr_f[b]-sdr_q[b])/10)                                      (no gradient)
//   kd = mean_b max_perm mean_j w[b] rho(est,fest)[b,perm(j),j] ;  task = same with target, w = 1
//   loss = -10 log10((1-lambda) task + lambda kd + eps)
//
// Three HBM passes over the 6 signals: means (fp64 atomics), 18 centred inner products (fp64),
// and the analytic gradient  g_est[b,i,:] = sum_j a_ij t_j + b_ij f_j + c_i e_i  (all centred).
// fp64 sufficient statistics avoid the cancellation of |e - alpha t|^2 at high SI-SDR.
#include "fqss_common.cuh"

namespace fqss {

int num_sms();

constexpr int LS_THREADS = 256;
constexpr int LS_CHUNK = 2048;      // samples of time per block
constexpr int NSTAT = 18;
constexpr double LEPS = 1e-8;

// stats layout per sample (double):
//  [0..5]   means: e0 e1 f0 f1 t0 t1
//  [6..23]  centred sums: ee0 ee1 ff0 ff1 tt0 tt1 | et00 et01 et10 et11 | ft00 ft01 ft10 ft11 | ef00 ef01 ef10 ef11
constexpr int LS_STRIDE = 24;
// coefficient layout per sample (float): a[2][2], b[2][2], c[2], then the 6 means as float
constexpr int LC_STRIDE = 16;

__global__ void __launch_bounds__(LS_THREADS) loss_mean_kernel(const float* __restrict__ est, int64_t lde,
                                                              const float* __restrict__ fest, int64_t ldf,
                                                              const float* __restrict__ tgt, int64_t ldt, int T,
                                                              double* __restrict__ st) {
    __shared__ double sh[6 * 32];
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * LS_CHUNK;
    const int t1 = min(T, t0 + LS_CHUNK);
    float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int t = t0 + threadIdx.x; t < t1; t += LS_THREADS) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            s[i] += __ldg(est + ((int64_t)b * 2 + i) * lde + t);
            s[2 + i] += __ldg(fest + ((int64_t)b * 2 + i) * ldf + t);
            s[4 + i] += __ldg(tgt + ((int64_t)b * 2 + i) * ldt + t);
        }
    }
    double v[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) v[i] = (double)s[i];
    block_sum<6>(v, sh);
    if (threadIdx.x == 0)
        for (int i = 0; i < 6; ++i) atomicAdd(st + b * LS_STRIDE + i, v[i]);
}

__global__ void __launch_bounds__(LS_THREADS) loss_dot_kernel(const float* __restrict__ est, int64_t lde,
                                                             const float* __restrict__ fest, int64_t ldf,
                                                             const float* __restrict__ tgt, int64_t ldt, int T,
                                                             double* __restrict__ st) {
    __shared__ double sh[NSTAT * 32];
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * LS_CHUNK;
    const int t1 = min(T, t0 + LS_CHUNK);
    float mu[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) mu[i] = (float)(st[b * LS_STRIDE + i] / (double)T);
    float s[NSTAT];
#pragma unroll
    for (int i = 0; i < NSTAT; ++i) s[i] = 0.f;
    for (int t = t0 + threadIdx.x; t < t1; t += LS_THREADS) {   // <= 8 terms per thread in fp32, then fp64
        float e[2], f[2], g[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            e[i] = __ldg(est + ((int64_t)b * 2 + i) * lde + t) - mu[i];
            f[i] = __ldg(fest + ((int64_t)b * 2 + i) * ldf + t) - mu[2 + i];
            g[i] = __ldg(tgt + ((int64_t)b * 2 + i) * ldt + t) - mu[4 + i];
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            s[i] = fmaf(e[i], e[i], s[i]);
            s[2 + i] = fmaf(f[i], f[i], s[2 + i]);
            s[4 + i] = fmaf(g[i], g[i], s[4 + i]);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                s[6 + 2 * i + j] = fmaf(e[i], g[j], s[6 + 2 * i + j]);
                s[10 + 2 * i + j] = fmaf(f[i], g[j], s[10 + 2 * i + j]);
                s[14 + 2 * i + j] = fmaf(e[i], f[j], s[14 + 2 * i + j]);
            }
        }
    }
    double v[NSTAT];
#pragma unroll
    for (int i = 0; i < NSTAT; ++i) v[i] = (double)s[i];
    block_sum<NSTAT>(v, sh);
    if (threadIdx.x == 0)
        for (int i = 0; i < NSTAT; ++i) atomicAdd(st + b * LS_STRIDE + 6 + i, v[i]);
}

struct Rho {
    double rho, alpha, inv_noise, tt, et_scale;   // pieces needed by the gradient
};

// rho of estimate (energy ee) against target (energy tt) with inner product d
__device__ __forceinline__ Rho make_rho(double ee, double tt, double d) {
    Rho r;
    double Et = tt + LEPS;
    r.alpha = d / Et;
    double P = r.alpha * r.alpha * tt;
    double N = ee - 2.0 * r.alpha * d + r.alpha * r.alpha * tt;
    N = N > 0.0 ? N : 0.0;
    r.inv_noise = 1.0 / (N + LEPS);
    r.rho = P * r.inv_noise;
    r.tt = tt;
    r.et_scale = 1.0 / Et;
    return r;
}

// d rho / d e = ce * e + ct * t   (e, t centred).  See DESIGN.md "loss gradient".
__device__ __forceinline__ void rho_grad(const Rho& r, double d, double& ce, double& ct) {
    // P = alpha^2 tt ; dP/de = 2 alpha tt / Et * t
    // N = |e - alpha t|^2 ; dN/de = 2 (e - alpha t) - 2 <n,t> / Et * t ,  <n,t> = d - alpha tt
    double nt = d - r.alpha * r.tt;
    double dP_t = 2.0 * r.alpha * r.tt * r.et_scale;
    double dN_e = 2.0;
    double dN_t = -2.0 * r.alpha - 2.0 * nt * r.et_scale;
    double P = r.rho / r.inv_noise;
    ce = -P * r.inv_noise * r.inv_noise * dN_e;
    ct = dP_t * r.inv_noise - P * r.inv_noise * r.inv_noise * dN_t;
}

// one block, thread b handles sample b (B <= 1024)
// means_io: NULL = single-process semantics.  phase 0: write the LOCAL means {kd, task, val} there and stop (the caller
// averages them over the data-parallel ranks); phase 1: take {kd, task, val} from there -- the means over the GLOBAL batch --
// for the logarithm, while the per-sample coefficients keep 1/(2 B_local), so that the mean over ranks of the per-rank
// gradients is the gradient of the global-batch loss (mysystem.py:145 under DDP).
__global__ void loss_finalize_kernel(const double* __restrict__ st, int B, int T, float kd_lambda, float* __restrict__ out,
                                     float* __restrict__ coef, int want_grad, double* __restrict__ means_io, int phase) {
    __shared__ double sh[3 * 32];
    const int b = threadIdx.x;
    double kd_b = 0.0, task_b = 0.0, val_b = 0.0;
    double w = 1.0;
    int pk = 0, pt = 0;
    Rho rt[2][2], rf[2][2];
    const double* s = st + (int64_t)b * LS_STRIDE;
    if (b < B) {
        double neg_f[2][2], neg_q[2][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                rt[i][j] = make_rho(s[6 + i], s[10 + j], s[12 + 2 * i + j]);          // est_i vs tgt_j
                rf[i][j] = make_rho(s[6 + i], s[8 + j], s[20 + 2 * i + j]);           // est_i vs fest_j
                Rho t = make_rho(s[8 + i], s[10 + j], s[16 + 2 * i + j]);             // fest_i vs tgt_j
                neg_f[i][j] = -10.0 * log10(t.rho + LEPS);
                neg_q[i][j] = -10.0 * log10(rt[i][j].rho + LEPS);
            }
        // PIT over the 2 permutations: identity uses (0,0),(1,1); swap uses (1,0),(0,1) as [est, tgt]
        double sdr_f = fmin(0.5 * (neg_f[0][0] + neg_f[1][1]), 0.5 * (neg_f[1][0] + neg_f[0][1]));
        double sdr_q = fmin(0.5 * (neg_q[0][0] + neg_q[1][1]), 0.5 * (neg_q[1][0] + neg_q[0][1]));
        val_b = sdr_q;
        w = pow(10.0, (sdr_f - sdr_q) / 10.0);
        double k_id = 0.5 * (rf[0][0].rho + rf[1][1].rho), k_sw = 0.5 * (rf[1][0].rho + rf[0][1].rho);
        double t_id = 0.5 * (rt[0][0].rho + rt[1][1].rho), t_sw = 0.5 * (rt[1][0].rho + rt[0][1].rho);
        // torch.min returns the first index on ties -> identity wins ties (losses are -rho)
        pk = (k_sw > k_id) ? 1 : 0;
        pt = (t_sw > t_id) ? 1 : 0;
        kd_b = w * (pk ? k_sw : k_id);
        task_b = pt ? t_sw : t_id;
    }
    double v[3] = {kd_b, task_b, val_b};
    block_sum<3>(v, sh);
    __shared__ double tot[3];
    if (threadIdx.x == 0) {
        tot[0] = v[0] / B; tot[1] = v[1] / B; tot[2] = v[2] / B;
        if (means_io && phase == 0) { means_io[0] = tot[0]; means_io[1] = tot[1]; means_io[2] = tot[2]; }
        if (means_io && phase == 1) { tot[0] = means_io[0]; tot[1] = means_io[1]; tot[2] = means_io[2]; }
    }
    __syncthreads();
    if (means_io && phase == 0) return;
    const double kd = tot[0], task = tot[1];
    const double A = (1.0 - (double)kd_lambda) * task + (double)kd_lambda * kd + LEPS;
    if (threadIdx.x == 0) {
        out[0] = (float)(-10.0 * log10(A));
        out[1] = (float)(-10.0 * log10(kd + LEPS));
        out[2] = (float)tot[2];
    }
    if (!want_grad || b >= B) return;
    // dL/dA = -10 / (ln10 A); dA/drho_t[i,j] = (1-l)/(B*2) on the winning task permutation,
    // dA/drho_f[i,j] = l*w/(B*2) on the winning kd permutation
    const double dLdA = -10.0 / (log(10.0) * A);
    const double gt = dLdA * (1.0 - (double)kd_lambda) / (2.0 * B);
    const double gf = dLdA * (double)kd_lambda * w / (2.0 * B);
    double a[2][2] = {{0, 0}, {0, 0}}, bb[2][2] = {{0, 0}, {0, 0}}, c[2] = {0, 0};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        int it = pt ? 1 - j : j;      // estimate index paired with target j
        double ce, ct;
        rho_grad(rt[it][j], s[12 + 2 * it + j], ce, ct);
        c[it] += gt * ce;
        a[it][j] += gt * ct;
        int ik = pk ? 1 - j : j;
        rho_grad(rf[ik][j], s[20 + 2 * ik + j], ce, ct);
        c[ik] += gf * ce;
        bb[ik][j] += gf * ct;
    }
    float* co = coef + (int64_t)b * LC_STRIDE;
    co[0] = (float)a[0][0]; co[1] = (float)a[0][1]; co[2] = (float)a[1][0]; co[3] = (float)a[1][1];
    co[4] = (float)bb[0][0]; co[5] = (float)bb[0][1]; co[6] = (float)bb[1][0]; co[7] = (float)bb[1][1];
    co[8] = (float)c[0]; co[9] = (float)c[1];
#pragma unroll
    for (int i = 0; i < 6; ++i) co[10 + i] = (float)(s[i] / (double)T);
}

__global__ void __launch_bounds__(LS_THREADS) loss_grad_kernel(const float* __restrict__ est, int64_t lde,
                                                              const float* __restrict__ fest, int64_t ldf,
                                                              const float* __restrict__ tgt, int64_t ldt, int T,
                                                              const float* __restrict__ coef, float* __restrict__ gest,
                                                              int64_t ldg) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * LS_THREADS + threadIdx.x;
    if (t >= T) return;
    const float* co = coef + (int64_t)b * LC_STRIDE;
    float e[2], f[2], g[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        e[i] = __ldg(est + ((int64_t)b * 2 + i) * lde + t) - co[10 + i];
        f[i] = __ldg(fest + ((int64_t)b * 2 + i) * ldf + t) - co[12 + i];
        g[i] = __ldg(tgt + ((int64_t)b * 2 + i) * ldt + t) - co[14 + i];
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float v = co[8 + i] * e[i] + co[2 * i] * g[0] + co[2 * i + 1] * g[1] + co[4 + 2 * i] * f[0] + co[4 + 2 * i + 1] * f[1];
        gest[((int64_t)b * 2 + i) * ldg + t] = v;
    }
}

}  // namespace fqss

using namespace fqss;

extern "C" int fqss_kd_loss(const float* est, int64_t lde, const float* fest, int64_t ldf, const float* tgt, int64_t ldt, int B,
                            int T, float kd_lambda, float* out, float* gest, int64_t ldg, void* ws, size_t ws_bytes,
                            void* stream) {
    FQSS_REQUIRE(est && fest && tgt && out && B > 0 && B <= 1024 && T > 0, -1, "kd_loss: bad argument (1 <= B <= 1024)");
    FQSS_REQUIRE(lde >= T && ldf >= T && ldt >= T && (!gest || ldg >= T), -1, "kd_loss: bad pitch");
    size_t st_bytes = (size_t)B * LS_STRIDE * sizeof(double);
    size_t need = st_bytes + (size_t)B * LC_STRIDE * sizeof(float);
    FQSS_REQUIRE(ws && ws_bytes >= need, -3, "kd_loss: workspace too small (%zu < %zu)", ws_bytes, need);
    cudaStream_t s = (cudaStream_t)stream;
    double* st = (double*)ws;
    float* coef = (float*)((char*)ws + st_bytes);
    FQSS_PROFN("kd_loss", s, gest ? 4 : 3);
    cudaMemsetAsync(st, 0, st_bytes, s);
    dim3 grid((T + LS_CHUNK - 1) / LS_CHUNK, B);
    loss_mean_kernel<<<grid, LS_THREADS, 0, s>>>(est, lde, fest, ldf, tgt, ldt, T, st);
    loss_dot_kernel<<<grid, LS_THREADS, 0, s>>>(est, lde, fest, ldf, tgt, ldt, T, st);
    int threads = ((B + 31) / 32) * 32;
    loss_finalize_kernel<<<1, threads, 0, s>>>(st, B, T, kd_lambda, out, coef, gest != nullptr, nullptr, 0);
    if (gest) {
        dim3 g2((T + LS_THREADS - 1) / LS_THREADS, B);
        loss_grad_kernel<<<g2, LS_THREADS, 0, s>>>(est, lde, fest, ldf, tgt, ldt, T, coef, gest, ldg);
    }
    return check_launch("kd_loss");
}

// Data-parallel variant with the loss of the GLOBAL batch: phase 0 = statistics + the local means {kd, task, val} (fp64, on the
// device) into `means`; the caller averages `means` over the ranks (one 3-double all-reduce); phase 1 = loss, logged values
// and dL/dest from those means.  `ws` must be the same, untouched buffer in both phases (it carries the statistics).
extern "C" int fqss_kd_loss_dp(const float* est, int64_t lde, const float* fest, int64_t ldf, const float* tgt, int64_t ldt, int B,
                               int T, float kd_lambda, float* out, float* gest, int64_t ldg, void* ws, size_t ws_bytes,
                               double* means, int phase, void* stream) {
    FQSS_REQUIRE(est && fest && tgt && means && B > 0 && B <= 1024 && T > 0 && (phase == 0 || phase == 1), -1,
                 "kd_loss_dp: bad argument (1 <= B <= 1024, phase 0 | 1)");
    FQSS_REQUIRE(phase == 0 || out, -1, "kd_loss_dp: phase 1 needs the output vector");
    FQSS_REQUIRE(lde >= T && ldf >= T && ldt >= T && (!gest || ldg >= T), -1, "kd_loss_dp: bad pitch");
    size_t st_bytes = (size_t)B * LS_STRIDE * sizeof(double);
    size_t need = st_bytes + (size_t)B * LC_STRIDE * sizeof(float);
    FQSS_REQUIRE(ws && ws_bytes >= need, -3, "kd_loss_dp: workspace too small (%zu < %zu)", ws_bytes, need);
    cudaStream_t s = (cudaStream_t)stream;
    double* st = (double*)ws;
    float* coef = (float*)((char*)ws + st_bytes);
    int threads = ((B + 31) / 32) * 32;
    if (phase == 0) {
        FQSS_PROFN("kd_loss", s, 3);
        cudaMemsetAsync(st, 0, st_bytes, s);
        dim3 grid((T + LS_CHUNK - 1) / LS_CHUNK, B);
        loss_mean_kernel<<<grid, LS_THREADS, 0, s>>>(est, lde, fest, ldf, tgt, ldt, T, st);
        loss_dot_kernel<<<grid, LS_THREADS, 0, s>>>(est, lde, fest, ldf, tgt, ldt, T, st);
        loss_finalize_kernel<<<1, threads, 0, s>>>(st, B, T, kd_lambda, nullptr, coef, 0, means, 0);
    } else {
        FQSS_PROFN("kd_loss", s, gest ? 2 : 1);
        loss_finalize_kernel<<<1, threads, 0, s>>>(st, B, T, kd_lambda, out, coef, gest != nullptr, means, 1);
        if (gest) {
            dim3 g2((T + LS_THREADS - 1) / LS_THREADS, B);
            loss_grad_kernel<<<g2, LS_THREADS, 0, s>>>(est, lde, fest, ldf, tgt, ldt, T, coef, gest, ldg);
        }
    }
    return check_launch("kd_loss_dp");
}

// The two HBM passes and the gradient pass as separate entry points: the sufficient statistics (means + 18 centred
// inner products per sample, fp64) and "apply per-sample coefficients" are all a pairwise SI-SDR matrix and its
// gradient need (PairwiseWSDR as a standalone module, wsdr.py:46-95); the O(B) algebra in between is the caller's.
extern "C" int fqss_loss_stats(const float* est, int64_t lde, const float* fest, int64_t ldf, const float* tgt, int64_t ldt, int B,
                               int T, double* stats, void* stream) {
    FQSS_REQUIRE(est && fest && tgt && stats && B > 0 && T > 0, -1, "loss_stats: bad argument");
    FQSS_REQUIRE(lde >= T && ldf >= T && ldt >= T, -1, "loss_stats: bad pitch");
    cudaStream_t s = (cudaStream_t)stream;
    FQSS_PROFN("loss_stats", s, 2);
    cudaMemsetAsync(stats, 0, (size_t)B * LS_STRIDE * sizeof(double), s);
    dim3 grid((T + LS_CHUNK - 1) / LS_CHUNK, B);
    loss_mean_kernel<<<grid, LS_THREADS, 0, s>>>(est, lde, fest, ldf, tgt, ldt, T, stats);
    loss_dot_kernel<<<grid, LS_THREADS, 0, s>>>(est, lde, fest, ldf, tgt, ldt, T, stats);
    return check_launch("loss_stats");
}

extern "C" int fqss_loss_grad_apply(const float* est, int64_t lde, const float* fest, int64_t ldf, const float* tgt, int64_t ldt,
                                    int B, int T, const float* coef, float* gest, int64_t ldg, void* stream) {
    FQSS_REQUIRE(est && fest && tgt && coef && gest && B > 0 && T > 0, -1, "loss_grad_apply: bad argument");
    FQSS_REQUIRE(lde >= T && ldf >= T && ldt >= T && ldg >= T, -1, "loss_grad_apply: bad pitch");
    FQSS_PROF("loss_grad_apply", stream);
    dim3 g2((T + LS_THREADS - 1) / LS_THREADS, B);
    loss_grad_kernel<<<g2, LS_THREADS, 0, (cudaStream_t)stream>>>(est, lde, fest, ldf, tgt, ldt, T, coef, gest, ldg);
    return check_launch("loss_grad_apply");
}
