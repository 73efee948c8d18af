// fq_ops.cu -- quantiser kernels (Q2-Q5), observers, FQSS splitter / reconstructor (P1),
// flat-arena helpers (D1) and the error plumbing of the C ABI.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "fqss_common.cuh"

namespace fqss {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
        return -4;
    }
    return 0;
}

static int g_num_sms = 0;
int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

// =============================================================================================
// Q2: standalone activation fake-quant.  HBM-bound: 8 B/elem fwd (+1 with codes), 12 B/elem bwd.
// Grid = multiple of the SM count, 4 x 128-bit loads in flight per thread.
// =============================================================================================
constexpr int FQ_THREADS = 256;
constexpr int FQ_UNROLL = 4;

template <bool WITH_CODE>
__global__ void __launch_bounds__(FQ_THREADS) fq_act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                               uint8_t* __restrict__ code, int64_t n,
                                                               const float* __restrict__ rmin,
                                                               const float* __restrict__ rmax, int n_bits) {
    const ActQ q = load_actq(rmin, rmax, n_bits);
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * FQ_THREADS;
    int64_t i = (int64_t)blockIdx.x * FQ_THREADS + threadIdx.x;
    for (; i + (FQ_UNROLL - 1) * stride < n4; i += FQ_UNROLL * stride) {
        float4 v[FQ_UNROLL];
#pragma unroll
        for (int u = 0; u < FQ_UNROLL; ++u) v[u] = ldg4_stream(x + 4 * (i + u * stride));
#pragma unroll
        for (int u = 0; u < FQ_UNROLL; ++u) {
            float c0 = actq_code(q, v[u].x), c1 = actq_code(q, v[u].y), c2 = actq_code(q, v[u].z),
                  c3 = actq_code(q, v[u].w);
            float4 o = make_float4(actq_decode(q, c0), actq_decode(q, c1), actq_decode(q, c2), actq_decode(q, c3));
            stg4(y + 4 * (i + u * stride), o);
            if (WITH_CODE) {
                uchar4 cc = make_uchar4((unsigned char)c0, (unsigned char)c1, (unsigned char)c2, (unsigned char)c3);
                *reinterpret_cast<uchar4*>(code + 4 * (i + u * stride)) = cc;
            }
        }
    }
    for (; i < n4; i += stride) {
        float4 v = ldg4_stream(x + 4 * i);
        float c0 = actq_code(q, v.x), c1 = actq_code(q, v.y), c2 = actq_code(q, v.z), c3 = actq_code(q, v.w);
        stg4(y + 4 * i, make_float4(actq_decode(q, c0), actq_decode(q, c1), actq_decode(q, c2), actq_decode(q, c3)));
        if (WITH_CODE)
            *reinterpret_cast<uchar4*>(code + 4 * i) =
                make_uchar4((unsigned char)c0, (unsigned char)c1, (unsigned char)c2, (unsigned char)c3);
    }
    // tail (n % 4)
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        int64_t j = (n4 << 2) + threadIdx.x;
        float c = actq_code(q, x[j]);
        y[j] = actq_decode(q, c);
        if (WITH_CODE) code[j] = (unsigned char)c;
    }
}

// range-gradient finalisation shared by every activation-FQ backward:
//   acc[0] = sum g*D, acc[1] = sum g*Z  ->  g_max = sD/levels, g_min = sZ - sD/levels
__global__ void actq_finalize_kernel(const double* __restrict__ acc, float* g_rmin, float* g_rmax, int n_bits) {
    double levels = (double)((1 << n_bits) - 1);
    double sD = acc[0], sZ = acc[1];
    if (g_rmax) *g_rmax = (float)(sD / levels);
    if (g_rmin) *g_rmin = (float)(sZ - sD / levels);
}

__global__ void __launch_bounds__(FQ_THREADS) fq_act_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                               float* __restrict__ gx, int64_t n,
                                                               const float* __restrict__ rmin,
                                                               const float* __restrict__ rmax, int n_bits,
                                                               double* __restrict__ acc) {
    __shared__ double sh[2 * 32];
    const ActQ q = load_actq(rmin, rmax, n_bits);
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * FQ_THREADS;
    float sD = 0.f, sZ = 0.f;
    double dD = 0.0, dZ = 0.0;
    int64_t i = (int64_t)blockIdx.x * FQ_THREADS + threadIdx.x;
    for (; i + stride < n4; i += 2 * stride) {
        float4 a0 = ldg4_stream(x + 4 * i), b0 = ldg4_stream(g + 4 * i);
        float4 a1 = ldg4_stream(x + 4 * (i + stride)), b1 = ldg4_stream(g + 4 * (i + stride));
        float4 o0, o1;
        o0.x = actq_bwd(q, a0.x, b0.x, sD, sZ); o0.y = actq_bwd(q, a0.y, b0.y, sD, sZ);
        o0.z = actq_bwd(q, a0.z, b0.z, sD, sZ); o0.w = actq_bwd(q, a0.w, b0.w, sD, sZ);
        o1.x = actq_bwd(q, a1.x, b1.x, sD, sZ); o1.y = actq_bwd(q, a1.y, b1.y, sD, sZ);
        o1.z = actq_bwd(q, a1.z, b1.z, sD, sZ); o1.w = actq_bwd(q, a1.w, b1.w, sD, sZ);
        stg4(gx + 4 * i, o0);
        stg4(gx + 4 * (i + stride), o1);
        dD += sD; dZ += sZ; sD = 0.f; sZ = 0.f;      // fp32 partials stay short, long sums in fp64
    }
    for (; i < n4; i += stride) {
        float4 a0 = ldg4_stream(x + 4 * i), b0 = ldg4_stream(g + 4 * i);
        float4 o0;
        o0.x = actq_bwd(q, a0.x, b0.x, sD, sZ); o0.y = actq_bwd(q, a0.y, b0.y, sD, sZ);
        o0.z = actq_bwd(q, a0.z, b0.z, sD, sZ); o0.w = actq_bwd(q, a0.w, b0.w, sD, sZ);
        stg4(gx + 4 * i, o0);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        int64_t j = (n4 << 2) + threadIdx.x;
        gx[j] = actq_bwd(q, x[j], g[j], sD, sZ);
    }
    double v[2] = {dD + sD, dZ + sZ};
    block_sum<2>(v, sh);
    if (threadIdx.x == 0) {
        atomicAdd(acc + 0, v[0]);
        atomicAdd(acc + 1, v[1]);
    }
}

// =============================================================================================
// Q3/Q5: weight fake-quant (tiny tensors; one CTA per output channel, deterministic reductions)
// =============================================================================================
__global__ void fq_weight_fwd_kernel(const float* __restrict__ w, float* __restrict__ wq, int8_t* __restrict__ code,
                                     int outer, int ch, int inner, const float* __restrict__ rmin,
                                     const float* __restrict__ rmax, int n_bits) {
    const int c = blockIdx.x;
    const WQ q = make_wq(rmin[c], rmax[c], n_bits);
    const int per = outer * inner;
    for (int j = threadIdx.x; j < per; j += blockDim.x) {
        int o = j / inner, i = j - o * inner;
        int64_t idx = ((int64_t)o * ch + c) * inner + i;
        float cc = wq_code(q, w[idx]);
        if (wq) wq[idx] = __fmul_rn(q.delta, cc);          // qat_quant.py:135
        if (code) code[idx] = (int8_t)cc;
    }
}

__global__ void fq_weight_bwd_kernel(const float* __restrict__ g, const float* __restrict__ w, float* __restrict__ gw,
                                     float* __restrict__ g_rmin, float* __restrict__ g_rmax, int outer, int ch,
                                     int inner, const float* __restrict__ rmin, const float* __restrict__ rmax,
                                     int n_bits) {
    __shared__ double sh[32];
    const int c = blockIdx.x;
    const float mn = rmin[c], mx = rmax[c];
    const WQ q = make_wq(mn, mx, n_bits);
    const int per = outer * inner;
    double acc = 0.0;
    for (int j = threadIdx.x; j < per; j += blockDim.x) {
        int o = j / inner, i = j - o * inner;
        int64_t idx = ((int64_t)o * ch + c) * inner + i;
        float u = __fdiv_rn(w[idx], q.delta);
        float X = rintf(u);
        bool in = (X >= q.lo) && (X <= q.hi);
        float cc = fminf(fmaxf(X, q.lo), q.hi);
        float gi = g[idx];
        if (gw) gw[idx] = in ? __fdiv_rn(__fmul_rn(gi, q.delta), q.delta) : 0.f;
        acc += (double)gi * (double)(in ? __fsub_rn(X, u) : cc);
    }
    double v[1] = {acc};
    block_sum<1>(v, sh);
    if (threadIdx.x == 0) {
        // d delta / d a = 2/levels ; a = maximum(|min|,|max|), ties split evenly (torch.maximum)
        double ga = 2.0 * v[0] / (double)((1 << n_bits) - 1);
        float amn = fabsf(mn), amx = fabsf(mx);
        double fmx = amx > amn ? 1.0 : (amx == amn ? 0.5 : 0.0);
        double fmn = 1.0 - fmx;
        float smx = (mx > 0.f) - (mx < 0.f), smn = (mn > 0.f) - (mn < 0.f);
        if (g_rmax) g_rmax[c] = (float)(ga * fmx * smx);
        if (g_rmin) g_rmin[c] = (float)(ga * fmn * smn);
    }
}

constexpr int WQ_BATCH = 48;
struct WqBatch {
    fqss_wq_item it[WQ_BATCH];
};

__global__ void fq_weight_bwd_batch_kernel(const __grid_constant__ WqBatch b) {
    __shared__ double sh[32];
    const fqss_wq_item& t = b.it[blockIdx.y];
    const int c = blockIdx.x;
    if (c >= t.ch) return;
    const float mn = t.rmin[c], mx = t.rmax[c];
    const WQ q = make_wq(mn, mx, t.n_bits);
    const int per = t.outer * t.inner;
    double acc = 0.0;
    for (int j = threadIdx.x; j < per; j += blockDim.x) {
        int o = j / t.inner, i = j - o * t.inner;
        int64_t idx = ((int64_t)o * t.ch + c) * t.inner + i;
        float u = __fdiv_rn(t.w[idx], q.delta);
        float X = rintf(u);
        bool in = (X >= q.lo) && (X <= q.hi);
        float cc = fminf(fmaxf(X, q.lo), q.hi);
        float gi = t.g[idx];
        if (t.out) t.out[idx] = in ? __fdiv_rn(__fmul_rn(gi, q.delta), q.delta) : 0.f;
        acc += (double)gi * (double)(in ? __fsub_rn(X, u) : cc);
    }
    double v[1] = {acc};
    block_sum<1>(v, sh);
    if (threadIdx.x == 0) {
        double ga = 2.0 * v[0] / (double)((1 << t.n_bits) - 1);
        float amn = fabsf(mn), amx = fabsf(mx);
        double fmx = amx > amn ? 1.0 : (amx == amn ? 0.5 : 0.0);
        double fmn = 1.0 - fmx;
        float smx = (mx > 0.f) - (mx < 0.f), smn = (mn > 0.f) - (mn < 0.f);
        if (t.g_rmax) t.g_rmax[c] = (float)(ga * fmx * smx);
        if (t.g_rmin) t.g_rmin[c] = (float)(ga * fmn * smn);
    }
}

__global__ void fq_weight_fwd_batch_kernel(const __grid_constant__ WqBatch b) {
    const fqss_wq_item& t = b.it[blockIdx.y];
    const int c = blockIdx.x;
    if (c >= t.ch) return;
    const WQ q = make_wq(t.rmin[c], t.rmax[c], t.n_bits);
    const int per = t.outer * t.inner;
    for (int j = threadIdx.x; j < per; j += blockDim.x) {
        int o = j / t.inner, i = j - o * t.inner;
        int64_t idx = ((int64_t)o * t.ch + c) * t.inner + i;
        t.out[idx] = __fmul_rn(q.delta, wq_code(q, t.w[idx]));
    }
}

__global__ void weight_observe_kernel(const float* __restrict__ w, int outer, int ch, int inner,
                                      float* __restrict__ rmin, float* __restrict__ rmax) {
    __shared__ float smn[32], smx[32];
    const int c = blockIdx.x;
    const int per = outer * inner;
    float lo = INFINITY, hi = -INFINITY;
    for (int j = threadIdx.x; j < per; j += blockDim.x) {
        int o = j / inner, i = j - o * inner;
        float v = w[((int64_t)o * ch + c) * inner + i];
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
    lo = warp_min(lo);
    hi = warp_max(hi);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { smn[wid] = lo; smx[wid] = hi; }
    __syncthreads();
    if (wid == 0) {
        int nw = (blockDim.x + 31) >> 5;
        lo = lane < nw ? smn[lane] : INFINITY;
        hi = lane < nw ? smx[lane] : -INFINITY;
        lo = warp_min(lo);
        hi = warp_max(hi);
        if (lane == 0) { rmin[c] = lo; rmax[c] = hi; }
    }
}

// =============================================================================================
// Q4: activation observer -- two-stage min/max + EMA, all on device (the reference does
// x.min()/x.max() and a Python-side EMA; no host sync here)
// =============================================================================================
__global__ void __launch_bounds__(256) minmax_partial_kernel(const float* __restrict__ x, int64_t rows, int64_t cols,
                                                            int64_t ld, float* __restrict__ part, int absmode) {
    __shared__ float smn[32], smx[32];
    float lo = INFINITY, hi = -INFINITY;
    const int64_t total = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / cols, c = i - r * cols;
        float v = x[r * ld + c];
        if (absmode) v = fabsf(v);
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
    lo = warp_min(lo);
    hi = warp_max(hi);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { smn[wid] = lo; smx[wid] = hi; }
    __syncthreads();
    if (wid == 0) {
        lo = lane < 8 ? smn[lane] : INFINITY;
        hi = lane < 8 ? smx[lane] : -INFINITY;
        lo = warp_min(lo);
        hi = warp_max(hi);
        if (lane == 0) { part[2 * blockIdx.x] = lo; part[2 * blockIdx.x + 1] = hi; }
    }
}

// mode 0: EMA into (rmin,rmax) ; mode 1: rmax[0] = max  (absmax for the splitter)
__global__ void minmax_final_kernel(const float* __restrict__ part, int nparts, float* rmin, float* rmax, float alpha,
                                    float om, int mode) {
    float lo = INFINITY, hi = -INFINITY;
    for (int i = threadIdx.x; i < nparts; i += 32) {
        lo = fminf(lo, part[2 * i]);
        hi = fmaxf(hi, part[2 * i + 1]);
    }
    lo = warp_min(lo);
    hi = warp_max(hi);
    if (threadIdx.x == 0) {
        if (mode == 0) {
            // om = float32(1 - alpha) evaluated in double on the host, as python does (qat_quant.py:231)
            *rmin = __fadd_rn(__fmul_rn(alpha, *rmin), __fmul_rn(om, lo));
            *rmax = __fadd_rn(__fmul_rn(alpha, *rmax), __fmul_rn(om, hi));
        } else {
            *rmax = hi;
        }
    }
}

// =============================================================================================
// P1: splitter / reconstructor
// =============================================================================================
__global__ void split_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ peak,
                             float* __restrict__ y, int64_t ldy, int B, int T, int n_split, int n_bits) {
    const float pk = __ldg(peak);
    const float half = (float)(1 << (n_bits - 1));       // 128
    const float step = 1.f / half;                        // delta = threshold / 2^(bits-1), threshold = 1
    const float lo = -half, hi = half - 1.f;
    const int64_t total = (int64_t)B * T;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int b = (int)(i / T);
        int t = (int)(i - (int64_t)b * T);
        float v = __fdiv_rn(x[(int64_t)b * ldx + t], pk);                               // process.py:23
        for (int s = 0; s < n_split; ++s) {
            float qv = __fmul_rn(fminf(fmaxf(floorf(__fdiv_rn(v, step)), lo), hi), step);   // process.py:10-14
            y[((int64_t)b * n_split + s) * ldy + t] = qv;
            // x = 2*(x - x_q)*threshold/delta - threshold   (left-to-right, process.py:33)
            v = __fsub_rn(__fdiv_rn(__fmul_rn(__fmul_rn(2.f, __fsub_rn(v, qv)), 1.f), step), 1.f);
        }
    }
}

__global__ void combine_kernel(const float* __restrict__ parts, int64_t part_stride, int64_t ld, float* __restrict__ y,
                               int64_t ldy, int64_t rows, int T, int n_comb, int n_bits) {
    const float base = 0.5f / (float)(1 << (n_bits - 1));
    const int64_t total = rows * T;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / T;
        int t = (int)(i - r * T);
        float acc = parts[r * ld + t];
        float sc = base;
        for (int k = 1; k < n_comb; ++k) {
            acc = __fadd_rn(acc, __fmul_rn(parts[k * part_stride + r * ld + t], sc));   // process.py:46
            sc *= base;
        }
        y[r * ldy + t] = acc;
    }
}

// =============================================================================================
// D1: flat gradient arena helpers
// =============================================================================================
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ acc) {
    __shared__ double sh[32];
    double s = 0.0;
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = ldg4(g + 4 * i);
        s += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        float v = g[(n4 << 2) + threadIdx.x];
        s += (double)v * v;
    }
    double v[1] = {s};
    block_sum<1>(v, sh);
    if (threadIdx.x == 0) atomicAdd(acc, v[0]);
}

__global__ void f64_to_f32_kernel(const double* __restrict__ a, float* __restrict__ o, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = (float)a[i];
}

__global__ void __launch_bounds__(256) scale_clip_kernel(float* __restrict__ g, int64_t n, const float* __restrict__ sumsq,
                                                        float pre_scale, float max_norm) {
    // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
    float total = sqrtf(__ldg(sumsq)) * pre_scale;
    float coef = fminf(max_norm / (total + 1e-6f), 1.f) * pre_scale;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        g[i] *= coef;
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                  float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                                  float bc1, float bc2_sqrt) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float gi = g[i];
        float mi = m[i] = b1 * m[i] + (1.f - b1) * gi;
        float vi = v[i] = b2 * v[i] + (1.f - b2) * gi * gi;
        float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] -= (lr / bc1) * (mi / denom);
    }
}

// Adam with the step count on the DEVICE (graph-capturable: nothing about the step is baked into the launch)
__global__ void __launch_bounds__(256) adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                      float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                                      const int* __restrict__ step, const float* __restrict__ lr_dev) {
    if (lr_dev) lr = __ldg(lr_dev);          // learning rate read from the device: a scheduler can change it between graph replays
    const float t = (float)(*step + 1);
    const float bc1 = 1.f - powf(b1, t);
    const float bc2_sqrt = sqrtf(1.f - powf(b2, t));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float gi = g[i];
        float mi = m[i] = b1 * m[i] + (1.f - b1) * gi;
        float vi = v[i] = b2 * v[i] + (1.f - b2) * gi * gi;
        float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] -= (lr / bc1) * (mi / denom);
    }
}
__global__ void incr_kernel(int* c) { *c += 1; }

}  // namespace fqss

// =============================================================================================
// C ABI
// =============================================================================================
using namespace fqss;

extern "C" {

int fqss_abi_version(void) { return FQSS_ABI_VERSION; }
const char* fqss_last_error(void) { return g_err; }

size_t fqss_ws_bytes(int64_t rows) {
    // fp64 accumulators (64) + per-row partial sums (8 floats per row) + min/max partials
    return (size_t)(4096 + 8 * sizeof(float) * (rows > 0 ? rows : 0) + 2 * sizeof(float) * 4096);
}

static inline int grid_for(int64_t work_items, int threads, int per_sm) {
    int64_t need = (work_items + threads - 1) / threads;
    int64_t cap = (int64_t)num_sms() * per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

int fqss_fq_act_fwd(const float* x, float* y, uint8_t* code, int64_t n, const float* rmin, const float* rmax,
                    int n_bits, void* stream) {
    FQSS_REQUIRE(x && y && rmin && rmax && n >= 0, -1, "fq_act_fwd: null pointer or negative size");
    FQSS_REQUIRE(n_bits >= 2 && n_bits <= 8, -1, "fq_act_fwd: n_bits=%d unsupported (2..8)", n_bits);
    FQSS_REQUIRE(aligned16(x) && aligned16(y) && (!code || (reinterpret_cast<uintptr_t>(code) & 3u) == 0), -2,
                 "fq_act_fwd: pointers must be 16-byte aligned");
    if (n == 0) return 0;
    int grid = grid_for((n >> 2) / FQ_UNROLL + 1, FQ_THREADS, 8);
    cudaStream_t s = (cudaStream_t)stream;
    FQSS_PROF("fq_act_fwd", s);
    if (code)
        fq_act_fwd_kernel<true><<<grid, FQ_THREADS, 0, s>>>(x, y, code, n, rmin, rmax, n_bits);
    else
        fq_act_fwd_kernel<false><<<grid, FQ_THREADS, 0, s>>>(x, y, nullptr, n, rmin, rmax, n_bits);
    return check_launch("fq_act_fwd");
}

int fqss_fq_act_bwd(const float* g, const float* x, float* gx, float* g_rmin, float* g_rmax, int64_t n,
                    const float* rmin, const float* rmax, int n_bits, void* ws, size_t ws_bytes, void* stream) {
    FQSS_REQUIRE(g && x && gx && rmin && rmax && n >= 0, -1, "fq_act_bwd: null pointer or negative size");
    FQSS_REQUIRE(n_bits >= 2 && n_bits <= 8, -1, "fq_act_bwd: n_bits=%d unsupported", n_bits);
    FQSS_REQUIRE(aligned16(x) && aligned16(g) && aligned16(gx), -2, "fq_act_bwd: pointers must be 16-byte aligned");
    FQSS_REQUIRE(ws && ws_bytes >= 64, -3, "fq_act_bwd: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    double* acc = (double*)ws;
    FQSS_PROFN("fq_act_bwd", s, 2);
    cudaMemsetAsync(acc, 0, 2 * sizeof(double), s);
    if (n > 0) {
        int grid = grid_for((n >> 2) / 2 + 1, FQ_THREADS, 8);
        fq_act_bwd_kernel<<<grid, FQ_THREADS, 0, s>>>(g, x, gx, n, rmin, rmax, n_bits, acc);
    }
    actq_finalize_kernel<<<1, 1, 0, s>>>(acc, g_rmin, g_rmax, n_bits);
    return check_launch("fq_act_bwd");
}

int fqss_fq_weight_fwd(const float* w, float* wq, int8_t* code, int outer, int ch, int inner, const float* rmin,
                       const float* rmax, int n_bits, void* stream) {
    FQSS_REQUIRE(w && (wq || code) && rmin && rmax, -1, "fq_weight_fwd: null pointer");
    FQSS_REQUIRE(outer > 0 && ch > 0 && inner > 0 && n_bits >= 2 && n_bits <= 8, -1, "fq_weight_fwd: bad shape/bits");
    FQSS_PROF("fq_weight_fwd", stream);
    fq_weight_fwd_kernel<<<ch, 128, 0, (cudaStream_t)stream>>>(w, wq, code, outer, ch, inner, rmin, rmax, n_bits);
    return check_launch("fq_weight_fwd");
}

int fqss_fq_weight_bwd(const float* g, const float* w, float* gw, float* g_rmin, float* g_rmax, int outer, int ch,
                       int inner, const float* rmin, const float* rmax, int n_bits, void* stream) {
    FQSS_REQUIRE(g && w && rmin && rmax, -1, "fq_weight_bwd: null pointer");
    FQSS_REQUIRE(outer > 0 && ch > 0 && inner > 0 && n_bits >= 2 && n_bits <= 8, -1, "fq_weight_bwd: bad shape/bits");
    FQSS_PROF("fq_weight_bwd", stream);
    fq_weight_bwd_kernel<<<ch, 128, 0, (cudaStream_t)stream>>>(g, w, gw, g_rmin, g_rmax, outer, ch, inner, rmin, rmax,
                                                               n_bits);
    return check_launch("fq_weight_bwd");
}

static int wq_batch(const fqss_wq_item* items, int n, void* stream, bool bwd) {
    const char* who = bwd ? "fq_weight_bwd_batch" : "fq_weight_fwd_batch";
    FQSS_REQUIRE(items && n > 0, -1, "%s: empty batch", who);
    cudaStream_t s = (cudaStream_t)stream;
    for (int i0 = 0; i0 < n; i0 += WQ_BATCH) {
        WqBatch b;
        const int m = n - i0 < WQ_BATCH ? n - i0 : WQ_BATCH;
        int maxch = 0;
        for (int i = 0; i < m; ++i) {
            const fqss_wq_item& t = items[i0 + i];
            FQSS_REQUIRE(t.w && t.rmin && t.rmax && (bwd ? t.g != nullptr : t.out != nullptr), -1, "%s: null pointer in item %d", who, i0 + i);
            FQSS_REQUIRE(t.outer > 0 && t.ch > 0 && t.inner > 0 && t.n_bits >= 2 && t.n_bits <= 8, -1, "%s: bad shape/bits in item %d", who, i0 + i);
            b.it[i] = t;
            if (t.ch > maxch) maxch = t.ch;
        }
        FQSS_PROF(bwd ? "fq_weight_bwd(batch)" : "fq_weight_fwd(batch)", s);
        if (bwd) fq_weight_bwd_batch_kernel<<<dim3(maxch, m), 128, 0, s>>>(b);
        else fq_weight_fwd_batch_kernel<<<dim3(maxch, m), 128, 0, s>>>(b);
    }
    return check_launch(who);
}

int fqss_fq_weight_fwd_batch(const fqss_wq_item* items, int n, void* stream) { return wq_batch(items, n, stream, false); }
int fqss_fq_weight_bwd_batch(const fqss_wq_item* items, int n, void* stream) { return wq_batch(items, n, stream, true); }

int fqss_weight_observe(const float* w, int outer, int ch, int inner, float* rmin, float* rmax, void* stream) {
    FQSS_REQUIRE(w && rmin && rmax && outer > 0 && ch > 0 && inner > 0, -1, "weight_observe: bad argument");
    FQSS_PROF("weight_observe", stream);
    weight_observe_kernel<<<ch, 128, 0, (cudaStream_t)stream>>>(w, outer, ch, inner, rmin, rmax);
    return check_launch("weight_observe");
}

static int minmax_common(const float* x, int64_t rows, int64_t cols, int64_t ld, float* rmin, float* rmax, double alpha,
                         int mode, void* ws, size_t ws_bytes, void* stream, const char* who) {
    FQSS_REQUIRE(x && rmax && rows > 0 && cols > 0 && ld >= cols, -1, "%s: bad argument", who);
    int grid = grid_for(rows * cols, 256, 4);
    FQSS_REQUIRE(ws && ws_bytes >= (size_t)grid * 2 * sizeof(float), -3, "%s: workspace too small", who);
    cudaStream_t s = (cudaStream_t)stream;
    FQSS_PROFN(who, s, 2);
    minmax_partial_kernel<<<grid, 256, 0, s>>>(x, rows, cols, ld, (float*)ws, mode);
    minmax_final_kernel<<<1, 32, 0, s>>>((const float*)ws, grid, rmin, rmax, (float)alpha, (float)(1.0 - alpha), mode);
    return check_launch(who);
}

int fqss_act_observe(const float* x, int64_t rows, int64_t cols, int64_t ld, float* rmin, float* rmax, double alpha,
                     void* ws, size_t ws_bytes, void* stream) {
    FQSS_REQUIRE(rmin, -1, "act_observe: null rmin");
    return minmax_common(x, rows, cols, ld, rmin, rmax, alpha, 0, ws, ws_bytes, stream, "act_observe");
}

int fqss_absmax(const float* x, int64_t rows, int64_t cols, int64_t ld, float* peak, void* ws, size_t ws_bytes,
                void* stream) {
    return minmax_common(x, rows, cols, ld, nullptr, peak, 0.0, 1, ws, ws_bytes, stream, "absmax");
}

int fqss_split(const float* x, int64_t ldx, const float* peak, float* y, int64_t ldy, int B, int T, int n_split,
               int n_bits, void* stream) {
    FQSS_REQUIRE(x && peak && y && B > 0 && T > 0 && n_split >= 1 && ldx >= T && ldy >= T, -1, "split: bad argument");
    FQSS_PROF("split", stream);
    split_kernel<<<grid_for((int64_t)B * T, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, ldx, peak, y, ldy, B, T, n_split,
                                                                                   n_bits);
    return check_launch("split");
}

int fqss_combine(const float* parts, int64_t part_stride, int64_t ld, float* y, int64_t ldy, int64_t rows, int T,
                 int n_comb, int n_bits, void* stream) {
    FQSS_REQUIRE(parts && y && rows > 0 && T > 0 && n_comb >= 1, -1, "combine: bad argument");
    FQSS_PROF("combine", stream);
    combine_kernel<<<grid_for(rows * T, 256, 8), 256, 0, (cudaStream_t)stream>>>(parts, part_stride, ld, y, ldy, rows, T,
                                                                                 n_comb, n_bits);
    return check_launch("combine");
}

// gather: many small gradient tensors -> their slices of the flat arena, one launch per 448 tensors (the table travels
// as a kernel parameter; nothing is allocated).  CTA (x, y) copies elements [4096 x, 4096 (x+1)) of tensor y.
constexpr int GATHER_BATCH = 448;
struct GatherBatch {
    fqss_gather_item it[GATHER_BATCH];
};
__global__ void __launch_bounds__(256) arena_gather_kernel(const __grid_constant__ GatherBatch b, float* __restrict__ dst) {
    const fqss_gather_item& t = b.it[blockIdx.y];
    const int64_t lo = (int64_t)blockIdx.x * 4096;
    if (lo >= t.numel) return;
    const int64_t hi = lo + 4096 < t.numel ? lo + 4096 : t.numel;
    float* out = dst + t.offset;
    if (t.src == nullptr) {
        for (int64_t i = lo + threadIdx.x; i < hi; i += 256) out[i] = 0.f;
    } else {
        for (int64_t i = lo + threadIdx.x; i < hi; i += 256) out[i] = __ldg(t.src + i);
    }
}

int fqss_arena_gather(const fqss_gather_item* items, int n, float* dst, void* stream) {
    FQSS_REQUIRE(items && n > 0 && dst, -1, "arena_gather: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    for (int i0 = 0; i0 < n; i0 += GATHER_BATCH) {
        GatherBatch b;
        const int m = n - i0 < GATHER_BATCH ? n - i0 : GATHER_BATCH;
        int64_t maxn = 1;
        for (int i = 0; i < m; ++i) {
            b.it[i] = items[i0 + i];
            FQSS_REQUIRE(b.it[i].numel >= 0 && b.it[i].offset >= 0, -1, "arena_gather: bad item %d", i0 + i);
            if (b.it[i].numel > maxn) maxn = b.it[i].numel;
        }
        FQSS_PROF("arena_gather", s);
        arena_gather_kernel<<<dim3((unsigned)((maxn + 4095) / 4096), m), 256, 0, s>>>(b, dst);
    }
    return check_launch("arena_gather");
}

int fqss_arena_sumsq(const float* g, int64_t n, float* sumsq, void* ws, size_t ws_bytes, void* stream) {
    FQSS_REQUIRE(g && sumsq && n >= 0 && aligned16(g), -1, "arena_sumsq: bad argument");
    FQSS_REQUIRE(ws && ws_bytes >= 8, -3, "arena_sumsq: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    FQSS_PROFN("arena_sumsq", s, 2);
    cudaMemsetAsync(ws, 0, sizeof(double), s);
    if (n > 0) sumsq_kernel<<<grid_for(n >> 2, 256, 4), 256, 0, s>>>(g, n, (double*)ws);
    f64_to_f32_kernel<<<1, 32, 0, s>>>((const double*)ws, sumsq, 1);
    return check_launch("arena_sumsq");
}

int fqss_arena_scale_clip(float* g, int64_t n, const float* sumsq, float pre_scale, float max_norm, void* stream) {
    FQSS_REQUIRE(g && sumsq && n >= 0, -1, "arena_scale_clip: bad argument");
    if (n == 0) return 0;
    FQSS_PROF("arena_scale_clip", stream);
    scale_clip_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(g, n, sumsq, pre_scale, max_norm);
    return check_launch("arena_scale_clip");
}

int fqss_arena_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                    float eps, int step, void* stream) {
    FQSS_REQUIRE(p && g && m && v && n >= 0 && step >= 1, -1, "arena_adam: bad argument");
    if (n == 0) return 0;
    float bc1 = 1.f - powf(beta1, (float)step);
    float bc2s = sqrtf(1.f - powf(beta2, (float)step));
    FQSS_PROF("arena_adam", stream);
    adam_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, bc2s);
    return check_launch("arena_adam");
}

int fqss_arena_adam_dev(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                        int* step_dev, const float* lr_dev, void* stream) {
    FQSS_REQUIRE(p && g && m && v && n >= 0 && step_dev, -1, "arena_adam_dev: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    FQSS_PROFN("arena_adam", s, 2);
    if (n > 0) adam_dev_kernel<<<grid_for(n, 256, 8), 256, 0, s>>>(p, g, m, v, n, lr, beta1, beta2, eps, step_dev, lr_dev);
    incr_kernel<<<1, 1, 0, s>>>(step_dev);
    return check_launch("arena_adam_dev");
}

}  // extern "C"
