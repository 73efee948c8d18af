// music_loss.cu -- the KD training loss of the music recipe (train_env/tasnet_musdbhq/musdbhq_train.py:87-109) as a fused
// segmented reduction + one gradient pass.  Per sample i of the batch (all sources / channels / samples of that item):
//
//   sdr_i  = 10 log10((sum fw^2 + eps) / (sum (fw - s)^2 + eps))        new-SDR of the float teacher   (process.py:70-75,
//   sdrq_i = 10 log10((sum w^2  + eps) / (sum (w  - s)^2 + eps))        ... and of the student          called as (ref=est, sig=src))
//   a_i    = 10^((sdr_i - sdrq_i) / 10)                                  no gradient
//   kd     = mean_i a_i * L1(w_i, fw_i) ;  task = L1(w, s) ;  loss = (1 - lambda) task + lambda kd       (nn.L1Loss = mean |.|)
//
// with eps = 1e-7, and for lambda = 0 just task.  dL/dw = (1-lambda) sign(w - s)/(B n) + lambda a_i sign(w - fw)/(B n)
// (n elements per item, torch's sign(0) = 0).  Tensors are row tensors [B*R rows][T] with a row pitch; R rows per item.
#include "fqss_common.cuh"

namespace fqss {

int num_sms();

constexpr int ML_THREADS = 256;
constexpr int ML_CHUNK = 4096;
constexpr int ML_NS = 6;            // per item: fw^2, (fw-s)^2, w^2, (w-s)^2, |w-fw|, |w-s|
constexpr double ML_EPS = 1e-7;

__global__ void __launch_bounds__(ML_THREADS) music_loss_stats_kernel(const float* __restrict__ w, int64_t ldw,
                                                                     const float* __restrict__ fw, int64_t ldf,
                                                                     const float* __restrict__ s, int64_t lds, int R, int T,
                                                                     double* __restrict__ st) {
    __shared__ double sh[ML_NS * 32];
    const int64_t row = blockIdx.y;
    const int item = (int)(row / R);
    const int t0 = blockIdx.x * ML_CHUNK, t1 = min(T, t0 + ML_CHUNK);
    float a[ML_NS] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int t = t0 + threadIdx.x; t < t1; t += ML_THREADS) {
        const float x = __ldg(w + row * ldw + t), f = fw ? __ldg(fw + row * ldf + t) : 0.f, y = __ldg(s + row * lds + t);
        const float dfs = f - y, dws = x - y, dwf = x - f;
        a[0] = fmaf(f, f, a[0]);
        a[1] = fmaf(dfs, dfs, a[1]);
        a[2] = fmaf(x, x, a[2]);
        a[3] = fmaf(dws, dws, a[3]);
        a[4] += fabsf(dwf);
        a[5] += fabsf(dws);
    }
    double v[ML_NS];
#pragma unroll
    for (int i = 0; i < ML_NS; ++i) v[i] = (double)a[i];
    block_sum<ML_NS>(v, sh);
    if (threadIdx.x == 0)
        for (int i = 0; i < ML_NS; ++i) atomicAdd(st + (int64_t)item * ML_NS + i, v[i]);
}

// one block: per-item weights a_i -> coef[i] (float), out = {loss, kd, task}
__global__ void music_loss_finalize_kernel(const double* __restrict__ st, int B, double n_per_item, float lambda, int use_kd,
                                           float* __restrict__ coef, float* __restrict__ out) {
    __shared__ double sh[2 * 32];
    double kd = 0.0, task = 0.0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const double* q = st + (int64_t)i * ML_NS;
        float a = 0.f;
        if (use_kd) {
            // float32 ratios, double log10, float32 dB values and float32 power: the precision ladder of the python code
            const float rf = (float)((q[0] + ML_EPS) / (q[1] + ML_EPS)), rq = (float)((q[2] + ML_EPS) / (q[3] + ML_EPS));
            const float sdr = (float)(10.0 * log10((double)rf)), sdrq = (float)(10.0 * log10((double)rq));
            a = powf(10.f, (sdr - sdrq) / 10.f);
            kd += (double)a * (q[4] / n_per_item);
        }
        coef[i] = a;
        task += q[5];
    }
    double v[2] = {kd, task};
    block_sum<2>(v, sh);
    if (threadIdx.x == 0) {
        const double kdm = v[0] / B, taskm = v[1] / (n_per_item * B);
        out[1] = (float)kdm;
        out[2] = (float)taskm;
        out[0] = use_kd ? (float)((1.0 - (double)lambda) * taskm + (double)lambda * kdm) : (float)taskm;
    }
}

__global__ void __launch_bounds__(ML_THREADS) music_loss_grad_kernel(const float* __restrict__ w, int64_t ldw,
                                                                    const float* __restrict__ fw, int64_t ldf,
                                                                    const float* __restrict__ s, int64_t lds, int R, int T,
                                                                    const float* __restrict__ coef, float c_task, float c_kd,
                                                                    float* __restrict__ g, int64_t ldg) {
    const int64_t row = blockIdx.y;
    const int item = (int)(row / R);
    const float ck = fw ? c_kd * __ldg(coef + item) : 0.f;
    const int t0 = blockIdx.x * ML_CHUNK, t1 = min(T, t0 + ML_CHUNK);
    for (int t = t0 + threadIdx.x; t < t1; t += ML_THREADS) {
        const float x = __ldg(w + row * ldw + t), y = __ldg(s + row * lds + t);
        const float d1 = x - y;
        float r = c_task * (d1 > 0.f ? 1.f : (d1 < 0.f ? -1.f : 0.f));
        if (fw) {
            const float d2 = x - __ldg(fw + row * ldf + t);
            r = fmaf(ck, d2 > 0.f ? 1.f : (d2 < 0.f ? -1.f : 0.f), r);
        }
        g[row * ldg + t] = r;
    }
}

}  // namespace fqss

using namespace fqss;

extern "C" {

size_t fqss_music_loss_ws_bytes(int B) { return (size_t)B * ML_NS * sizeof(double) + (size_t)B * sizeof(float) + 256; }

int fqss_music_kd_loss(const float* wavs, int64_t ldw, const float* fwavs, int64_t ldf, const float* sources, int64_t lds, int B,
                       int R, int T, float kd_lambda, float* out, float* g, int64_t ldg, void* ws, size_t ws_bytes, void* stream) {
    FQSS_REQUIRE(wavs && sources && out && ws, -1, "music_kd_loss: null argument");
    FQSS_REQUIRE(B > 0 && R > 0 && T > 0 && ldw >= T && lds >= T && (!fwavs || ldf >= T) && (!g || ldg >= T), -1, "music_kd_loss: bad shape");
    FQSS_REQUIRE(ws_bytes >= fqss_music_loss_ws_bytes(B), -3, "music_kd_loss: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    double* st = (double*)ws;
    float* coef = (float*)(st + (size_t)B * ML_NS);
    const int use_kd = (fwavs != nullptr && kd_lambda > 0.f) ? 1 : 0;
    const float* fw = use_kd ? fwavs : nullptr;
    FQSS_PROFN("music_kd_loss", s, g ? 3 : 2);
    cudaMemsetAsync(st, 0, (size_t)B * ML_NS * sizeof(double), s);
    dim3 grid((unsigned)((T + ML_CHUNK - 1) / ML_CHUNK), (unsigned)((int64_t)B * R));
    music_loss_stats_kernel<<<grid, ML_THREADS, 0, s>>>(wavs, ldw, fw, ldf, sources, lds, R, T, st);
    int rc = check_launch("music_kd_loss(stats)");
    if (rc) return rc;
    const double n_item = (double)R * (double)T;
    music_loss_finalize_kernel<<<1, 256, 0, s>>>(st, B, n_item, kd_lambda, use_kd, coef, out);
    rc = check_launch("music_kd_loss(finalize)");
    if (rc) return rc;
    if (g) {
        const float lam = use_kd ? kd_lambda : 0.f;
        const float c_task = (float)((1.0 - (double)lam) / (n_item * B)), c_kd = (float)((double)lam / (n_item * B));
        music_loss_grad_kernel<<<grid, ML_THREADS, 0, s>>>(wavs, ldw, fw, ldf, sources, lds, R, T, coef, c_task, c_kd, g, ldg);
        rc = check_launch("music_kd_loss(grad)");
    }
    return rc;
}

}  // extern "C"
