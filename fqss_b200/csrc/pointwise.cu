// pointwise.cu -- "op -> nonlinearity -> fake-quant" layers as single HBM passes, forward and
// backward (qat_layers.py AddQ :62, MulQ :86, GroupNormQ :438, NlQ :511 and the nl+FQ tails of
// Conv1dQ :124 / Conv1dNlQ :188), plus the gLN statistics pass.
//
// Row tensors: [rows x cols], row pitch ld.  One CTA handles one 1024-column chunk of one row with
// 128-bit accesses (4 columns per thread); a scalar twin covers layouts that are not 16-byte
// aligned.  The pre-quant value z is recomputed from the layer inputs in backward, so nothing but
// the inputs is saved between forward and backward.
#include "fqss_common.cuh"

namespace fqss {

int num_sms();

// float per-thread partials -> warp sums in fp32 -> cross-warp sums in fp64 (valid in thread 0); sh: 5*32 doubles
__device__ __forceinline__ void block_sum_fd5(const float (&s)[5], double (&v)[5], double* sh) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    float w[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) w[i] = warp_sum(s[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 5; ++i) sh[i * 32 + wid] = (double)w[i];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            double x = lane < nw ? sh[i * 32 + lane] : 0.0;
            v[i] = warp_sum(x);
        }
    }
}

constexpr int PW_THREADS = 256;
constexpr int PW_QUADS = 4;                               // frame quads per thread
constexpr int PW_CHUNK = PW_THREADS * 4 * PW_QUADS;       // elements of one row handled by one CTA

struct RowCtx {
    float slope;          // PReLU
    float scale, shift;   // gLN: z = x*scale + shift
    float mu, rstd;       // gLN backward
    float gamma;
};

// Per-row gLN constants: fp64 mean / variance by ONE thread, broadcast through shared memory (contains a barrier).
__device__ __forceinline__ void gln_row_consts(const fqss_pw_desc& d, int64_t row, RowCtx& c, float* sh5) {
    if (threadIdx.x == 0) {
        const int64_t b = row / d.C;
        const int ch = (int)(row - b * d.C);
        const double N = (double)d.C * (double)d.cols;
        const double mean = d.stats[2 * b] / N;
        double var = d.stats[2 * b + 1] / N - mean * mean;
        var = var > 0.0 ? var : 0.0;
        const float mu = (float)mean, rstd = (float)(1.0 / sqrt(var + (double)d.eps));
        const float gamma = __ldg(d.gamma + ch);
        const float scale = __fmul_rn(rstd, gamma);                          // ATen group_norm: scale = rstd*gamma
        sh5[0] = mu; sh5[1] = rstd; sh5[2] = gamma; sh5[3] = scale;
        sh5[4] = __fadd_rn(__fmul_rn(-scale, mu), __ldg(d.beta + ch));       //                bias = -scale*mean + beta
    }
    __syncthreads();
    c.mu = sh5[0]; c.rstd = sh5[1]; c.gamma = sh5[2]; c.scale = sh5[3]; c.shift = sh5[4];
}

template <int KIND>
__device__ __forceinline__ float pw_z(float a, float b, const RowCtx& c) {
    if (KIND == FQSS_PW_IDENT) return a;
    if (KIND == FQSS_PW_PRELU) return a > 0.f ? a : __fmul_rn(c.slope, a);
    if (KIND == FQSS_PW_RELU) return fmaxf(a, 0.f);
    if (KIND == FQSS_PW_ADD) return __fadd_rn(a, b);
    if (KIND == FQSS_PW_SUB) return __fsub_rn(a, b);
    if (KIND == FQSS_PW_MUL) return __fmul_rn(a, b);
    return __fadd_rn(__fmul_rn(a, c.scale), c.shift);   // GLN
}

template <int KIND>
__device__ __forceinline__ int64_t x2_row(const fqss_pw_desc& d, int64_t row) {
    if (KIND == FQSS_PW_MUL && d.bcast > 1) {
        int64_t per = (int64_t)d.bcast * d.C;
        return (row / per) * d.C + row % d.C;
    }
    return row;
}

template <int KIND>
__host__ __device__ constexpr bool pw_binary() { return KIND == FQSS_PW_ADD || KIND == FQSS_PW_SUB || KIND == FQSS_PW_MUL; }

// backward through the quantiser, bit-identical to the reference's autograd chain ((g*delta)*mask)/delta, with the
// division-free exact quotient; accumulates sD += g*(in ? X - t : clip(X)), sZ += g*(1 - in)
__device__ __forceinline__ float pw_fq_bwd(const ActQF& q, float z, float g, float& sD, float& sZ) {
    const float t = actqf_t(q, z);
    const bool in = actqf_inside(q, t);
    const float c = actqf_unbias(actqf_biased(q, t));
    sD = fmaf(g, in ? __fsub_rn(c, t) : c, sD);
    sZ += in ? 0.f : g;
    return in ? exact_div(__fmul_rn(g, q.delta), q.delta, q.inv) : 0.f;
}

__device__ __forceinline__ void ld_quad(const float* p, int nv, bool vec, float (&a)[4]) {
    if (vec) {
        const float4 v = ldg4(p);
        a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = k < nv ? p[k] : 0.f;
    }
}
__device__ __forceinline__ void st_quad(float* p, int nv, bool vec, const float (&a)[4]) {
    if (vec) {
        stg4(p, make_float4(a[0], a[1], a[2], a[3]));      // pad columns (< ld) may be written: don't-care
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < nv) p[k] = a[k];
    }
}

// ---------------------------------------------------------------------------------------------
// forward: one CTA = PW_CHUNK consecutive elements of one row (4 quads per thread)
// ---------------------------------------------------------------------------------------------
template <int KIND, bool VEC>
__global__ void __launch_bounds__(PW_THREADS) pw_fwd_kernel(const fqss_pw_desc d) {
    __shared__ float sh5[5];
    const int64_t row = blockIdx.x;
    const int64_t chunk0 = (int64_t)blockIdx.y * PW_CHUNK;
    const float* r1 = d.x1 + row * d.ld1;
    const float* r2 = pw_binary<KIND>() ? d.x2 + x2_row<KIND>(d, row) * d.ld2 : nullptr;
    float* ry = d.y + row * d.ldy;
    // the CTA's data is requested BEFORE the row constants and the quantiser constants are derived (a barrier, an fp64
    // division / square root and dependent loads): the CTA lives for 16 elements per thread, so "constants, then data"
    // in sequence exposes two memory round trips per CTA
    float a[PW_QUADS][4], b[PW_QUADS][4];
#pragma unroll
    for (int qd = 0; qd < PW_QUADS; ++qd) {
        const int64_t c0 = chunk0 + ((int64_t)qd * PW_THREADS + threadIdx.x) * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) a[qd][k] = b[qd][k] = 0.f;
        if (c0 < d.cols) {
            const int nv = (int)min((int64_t)4, d.cols - c0);
            ld_quad(r1 + c0, nv, VEC, a[qd]);
            if (pw_binary<KIND>()) ld_quad(r2 + c0, nv, VEC, b[qd]);
        }
    }
    RowCtx rc;
    if (KIND == FQSS_PW_PRELU) rc.slope = __ldg(d.slope);
    if (KIND == FQSS_PW_GLN) gln_row_consts(d, row, rc, sh5);
    ActQF q;
    if (d.quant) q = load_actqf(d.rmin, d.rmax, d.n_bits);
#pragma unroll
    for (int qd = 0; qd < PW_QUADS; ++qd) {
        const int64_t c0 = chunk0 + ((int64_t)qd * PW_THREADS + threadIdx.x) * 4;
        if (c0 >= d.cols) break;
        const int nv = (int)min((int64_t)4, d.cols - c0);
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float z = pw_z<KIND>(a[qd][k], b[qd][k], rc);
            o[k] = d.quant ? actqf_fq(q, z) : z;
        }
        st_quad(ry + c0, nv, VEC, o);
    }
}

// ---------------------------------------------------------------------------------------------
// backward.  acc: fp64 accumulators {sum g*D, sum g*Z, slope grad}; rowacc: per-row {sum g_n,
// sum g_n*xhat} for gLN (phase A).  PHASE: 0 = single-pass kinds, 1 = gLN phase A (sums only),
// 2 = gLN phase B (writes gx1, no sums).
// ---------------------------------------------------------------------------------------------
template <int KIND, bool VEC, int PHASE>
__global__ void __launch_bounds__(PW_THREADS) pw_bwd_kernel(const fqss_pw_desc d, const fqss_pw_grads o,
                                                           double* __restrict__ acc, double* __restrict__ rowacc,
                                                           const double* __restrict__ samp) {
    __shared__ double sh[5 * 32];
    __shared__ float sh5[5];
    const int64_t row = blockIdx.x;
    const int64_t chunk0 = (int64_t)blockIdx.y * PW_CHUNK;
    RowCtx rc;
    if (KIND == FQSS_PW_PRELU) rc.slope = __ldg(d.slope);
    if (KIND == FQSS_PW_GLN) gln_row_consts(d, row, rc, sh5);
    ActQF q;
    if (d.quant) q = load_actqf(d.rmin, d.rmax, d.n_bits);
    float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};       // sD, sZ, slope, r1, r2
    float S1 = 0.f, S2 = 0.f, invN = 0.f;
    if (KIND == FQSS_PW_GLN && PHASE == 2) {
        const int64_t bs = row / d.C;
        invN = __fdividef(1.f, (float)d.C * (float)d.cols);
        S1 = (float)samp[2 * bs];
        S2 = (float)samp[2 * bs + 1];
    }
    const float* r1 = d.x1 + row * d.ld1;
    const float* r2 = pw_binary<KIND>() ? d.x2 + x2_row<KIND>(d, row) * d.ld2 : nullptr;
    const float* rg = o.g + row * o.ldg;
#pragma unroll
    for (int qd = 0; qd < PW_QUADS; ++qd) {
        const int64_t c0 = chunk0 + ((int64_t)qd * PW_THREADS + threadIdx.x) * 4;
        if (c0 >= d.cols) break;
        const int nv = (int)min((int64_t)4, d.cols - c0);
        float a[4], b[4] = {0.f, 0.f, 0.f, 0.f}, g[4], g1[4], g2[4];
        ld_quad(r1 + c0, nv, VEC, a);
        ld_quad(rg + c0, nv, VEC, g);
        if (pw_binary<KIND>()) ld_quad(r2 + c0, nv, VEC, b);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool valid = k < nv;
            if (!valid) { a[k] = 0.f; b[k] = 0.f; }          // pad columns may hold NaN garbage
            const float z = pw_z<KIND>(a[k], b[k], rc);
            const float gk = valid ? g[k] : 0.f;
            float dsD = 0.f, dsZ = 0.f;
            float gz = d.quant ? pw_fq_bwd(q, z, gk, dsD, dsZ) : gk;
            if (valid && PHASE != 2) { s[0] += dsD; s[1] += dsZ; }
            if (!valid) gz = 0.f;
            if (KIND == FQSS_PW_IDENT || KIND == FQSS_PW_ADD) g1[k] = gz;
            if (KIND == FQSS_PW_SUB) { g1[k] = gz; g2[k] = -gz; }
            if (KIND == FQSS_PW_PRELU) {
                g1[k] = a[k] > 0.f ? gz : rc.slope * gz;
                s[2] += a[k] > 0.f ? 0.f : a[k] * gz;
            }
            if (KIND == FQSS_PW_RELU) g1[k] = a[k] > 0.f ? gz : 0.f;
            if (KIND == FQSS_PW_MUL) { g1[k] = gz * b[k]; g2[k] = gz * a[k]; }
            if (KIND == FQSS_PW_GLN) {
                const float xh = (a[k] - rc.mu) * rc.rstd;
                if (PHASE == 1) { s[3] += gz; s[4] += gz * xh; }
                if (PHASE == 2) g1[k] = rc.rstd * (rc.gamma * gz - (S1 + xh * S2) * invN);
            }
        }
        if (PHASE != 1) {
            if (o.gx1) st_quad(o.gx1 + row * o.ldg1 + c0, nv, VEC, g1);
            if (KIND == FQSS_PW_SUB && o.gx2) st_quad(o.gx2 + row * o.ldg2 + c0, nv, VEC, g2);
        }
    }
    if (PHASE == 2) return;
    double v[5];
    block_sum_fd5(s, v, sh);
    if (threadIdx.x == 0) {
        if (d.quant) { atomicAdd(acc + 0, v[0]); atomicAdd(acc + 1, v[1]); }
        if (KIND == FQSS_PW_PRELU) atomicAdd(acc + 2, v[2]);
        if (KIND == FQSS_PW_GLN) { atomicAdd(rowacc + 2 * row, v[3]); atomicAdd(rowacc + 2 * row + 1, v[4]); }
    }
}

// MulQ backward: grid over the rows of x2 ([B,C]); loops the `bcast` sources so that gx2 (sum over
// sources) needs no atomics.
template <bool VEC>
__global__ void __launch_bounds__(PW_THREADS) pw_mul_bwd_kernel(const fqss_pw_desc d, const fqss_pw_grads o,
                                                               double* __restrict__ acc) {
    __shared__ double sh[5 * 32];
    const int64_t row2 = blockIdx.x;                         // (b, c)
    const int64_t chunk0 = (int64_t)blockIdx.y * PW_CHUNK;
    ActQF q;
    if (d.quant) q = load_actqf(d.rmin, d.rmax, d.n_bits);
    float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    const int64_t bs = row2 / d.C, ch = row2 - bs * d.C;
#pragma unroll
    for (int qd = 0; qd < PW_QUADS; ++qd) {
        const int64_t c0 = chunk0 + ((int64_t)qd * PW_THREADS + threadIdx.x) * 4;
        if (c0 >= d.cols) break;
        const int nv = (int)min((int64_t)4, d.cols - c0);
        float b[4], s2[4] = {0.f, 0.f, 0.f, 0.f};
        ld_quad(d.x2 + row2 * d.ld2 + c0, nv, VEC, b);
        for (int sidx = 0; sidx < d.bcast; ++sidx) {
            const int64_t row1 = (bs * d.bcast + sidx) * d.C + ch;
            float a[4], g[4], g1[4];
            ld_quad(d.x1 + row1 * d.ld1 + c0, nv, VEC, a);
            ld_quad(o.g + row1 * o.ldg + c0, nv, VEC, g);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool valid = k < nv;
                const float ak = valid ? a[k] : 0.f, bk = valid ? b[k] : 0.f;
                const float gk = valid ? g[k] : 0.f;
                const float z = __fmul_rn(ak, bk);
                float dsD = 0.f, dsZ = 0.f;
                float gz = d.quant ? pw_fq_bwd(q, z, gk, dsD, dsZ) : gk;
                if (valid) { s[0] += dsD; s[1] += dsZ; } else gz = 0.f;
                g1[k] = gz * bk;
                s2[k] += gz * ak;
            }
            if (o.gx1) st_quad(o.gx1 + row1 * o.ldg1 + c0, nv, VEC, g1);
        }
        if (o.gx2) st_quad(o.gx2 + row2 * o.ldg2 + c0, nv, VEC, s2);
    }
    double v[5];
    block_sum_fd5(s, v, sh);
    if (threadIdx.x == 0 && d.quant) { atomicAdd(acc + 0, v[0]); atomicAdd(acc + 1, v[1]); }
}

// gLN backward reduction between the phases: per-channel dgamma/dbeta (sum over samples) and
// per-sample {S1 = sum_c gamma_c r1, S2 = sum_c gamma_c r2}.
__global__ void gln_reduce_kernel(const double* __restrict__ rowacc, int B, int C, const float* __restrict__ gamma,
                                  float* __restrict__ g_gamma, float* __restrict__ g_beta, double* __restrict__ samp) {
    __shared__ double sh[2 * 32];
    if ((int)blockIdx.x < B) {          // per-sample sums
        const int b = blockIdx.x;
        double s1 = 0.0, s2 = 0.0;
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            double gm = (double)gamma[c];
            s1 += gm * rowacc[2 * ((int64_t)b * C + c)];
            s2 += gm * rowacc[2 * ((int64_t)b * C + c) + 1];
        }
        double v[2] = {s1, s2};
        block_sum<2>(v, sh);
        if (threadIdx.x == 0) { samp[2 * b] = v[0]; samp[2 * b + 1] = v[1]; }
    } else {                            // per-channel sums: remaining blocks stride over channels
        const int nb = gridDim.x - B;
        for (int c = (blockIdx.x - B) * blockDim.x + threadIdx.x; c < C; c += nb * blockDim.x) {
            double gb = 0.0, gg = 0.0;
            for (int b = 0; b < B; ++b) {
                gb += rowacc[2 * ((int64_t)b * C + c)];
                gg += rowacc[2 * ((int64_t)b * C + c) + 1];
            }
            if (g_beta) g_beta[c] = (float)gb;
            if (g_gamma) g_gamma[c] = (float)gg;
        }
    }
}

__global__ void pw_finalize_kernel(const double* __restrict__ acc, float* g_rmin, float* g_rmax, float* g_slope,
                                   int n_bits, int quant) {
    if (quant) {
        double levels = (double)((1 << n_bits) - 1);
        if (g_rmax) *g_rmax = (float)(acc[0] / levels);
        if (g_rmin) *g_rmin = (float)(acc[1] - acc[0] / levels);
    }
    if (g_slope) *g_slope = (float)acc[2];
}

// ---------------------------------------------------------------------------------------------
// gLN statistics: per-sample sum / sum of squares (fp32 per thread over <= 4 values, fp64 beyond)
// ---------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(PW_THREADS) gln_stats_kernel(const float* __restrict__ x, int64_t cols, int64_t ld, int C,
                                                              double* __restrict__ stats) {
    __shared__ double sh[5 * 32];
    const int64_t row = blockIdx.x;
    const int64_t chunk0 = (int64_t)blockIdx.y * PW_CHUNK;
    float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int qd = 0; qd < PW_QUADS; ++qd) {
        const int64_t c0 = chunk0 + ((int64_t)qd * PW_THREADS + threadIdx.x) * 4;
        if (c0 >= cols) break;
        const int nv = (int)min((int64_t)4, cols - c0);
        float a[4];
        ld_quad(x + row * ld + c0, nv, VEC, a);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < nv) { s[0] += a[k]; s[1] = fmaf(a[k], a[k], s[1]); }
    }
    double v[5];
    block_sum_fd5(s, v, sh);
    if (threadIdx.x == 0) {
        atomicAdd(stats + 2 * (row / C), v[0]);
        atomicAdd(stats + 2 * (row / C) + 1, v[1]);
    }
}

static bool vec_ok(const fqss_pw_desc& d) {
    bool ok = aligned16(d.x1) && aligned16(d.y) && (d.ld1 % 4 == 0) && (d.ldy % 4 == 0);
    if (d.x2) ok = ok && aligned16(d.x2) && (d.ld2 % 4 == 0);
    return ok;
}

template <int KIND>
static void launch_fwd(const fqss_pw_desc& d, dim3 grid, cudaStream_t s) {
    if (vec_ok(d)) pw_fwd_kernel<KIND, true><<<grid, PW_THREADS, 0, s>>>(d);
    else pw_fwd_kernel<KIND, false><<<grid, PW_THREADS, 0, s>>>(d);
}

template <int KIND, int PHASE>
static void launch_bwd(const fqss_pw_desc& d, const fqss_pw_grads& o, bool vec, dim3 grid, cudaStream_t s, double* acc,
                       double* rowacc, const double* samp) {
    if (vec) pw_bwd_kernel<KIND, true, PHASE><<<grid, PW_THREADS, 0, s>>>(d, o, acc, rowacc, samp);
    else pw_bwd_kernel<KIND, false, PHASE><<<grid, PW_THREADS, 0, s>>>(d, o, acc, rowacc, samp);
}

static int validate(const fqss_pw_desc* d, const char* who) {
    FQSS_REQUIRE(d, -1, "%s: null descriptor", who);
    FQSS_REQUIRE(d->kind >= FQSS_PW_IDENT && d->kind <= FQSS_PW_GLN, -1, "%s: unknown kind %d", who, d->kind);
    FQSS_REQUIRE(d->rows > 0 && d->cols > 0 && d->x1 && d->ld1 >= d->cols, -1, "%s: bad shape", who);
    FQSS_REQUIRE(d->rows <= 0x7fffffffLL, -1, "%s: too many rows", who);
    if (d->quant) {
        FQSS_REQUIRE(d->rmin && d->rmax, -1, "%s: quant=1 needs ranges", who);
        FQSS_REQUIRE(d->n_bits >= 2 && d->n_bits <= 8, -1, "%s: n_bits=%d unsupported", who, d->n_bits);
    }
    if (d->kind == FQSS_PW_ADD || d->kind == FQSS_PW_SUB || d->kind == FQSS_PW_MUL)
        FQSS_REQUIRE(d->x2 && d->ld2 >= d->cols, -1, "%s: binary kind needs x2", who);
    if (d->kind == FQSS_PW_MUL) FQSS_REQUIRE(d->bcast >= 1 && d->C >= 1 && d->rows % ((int64_t)d->bcast * d->C) == 0, -1,
                                             "%s: MUL needs rows == B*bcast*C", who);
    if (d->kind == FQSS_PW_PRELU) FQSS_REQUIRE(d->slope, -1, "%s: PRELU needs slope", who);
    if (d->kind == FQSS_PW_GLN)
        FQSS_REQUIRE(d->gamma && d->beta && d->stats && d->C >= 1 && d->rows % d->C == 0, -1, "%s: GLN needs gamma/beta/stats/C",
                     who);
    return 0;
}

}  // namespace fqss

using namespace fqss;

extern "C" {

int fqss_pw_fwd(const fqss_pw_desc* dp, void* stream) {
    int rc = validate(dp, "pw_fwd");
    if (rc) return rc;
    FQSS_REQUIRE(dp->y && dp->ldy >= dp->cols, -1, "pw_fwd: bad output");
    const fqss_pw_desc& d = *dp;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid((unsigned)d.rows, (unsigned)((d.cols + PW_CHUNK - 1) / PW_CHUNK));
    static const char* const names[] = {"pw_fwd(ident)", "pw_fwd(prelu)", "pw_fwd(relu)", "pw_fwd(add)", "pw_fwd(sub)", "pw_fwd(mul)", "pw_fwd(gln)"};
    FQSS_PROF(names[d.kind], s);
    switch (d.kind) {
        case FQSS_PW_IDENT: launch_fwd<FQSS_PW_IDENT>(d, grid, s); break;
        case FQSS_PW_PRELU: launch_fwd<FQSS_PW_PRELU>(d, grid, s); break;
        case FQSS_PW_RELU: launch_fwd<FQSS_PW_RELU>(d, grid, s); break;
        case FQSS_PW_ADD: launch_fwd<FQSS_PW_ADD>(d, grid, s); break;
        case FQSS_PW_SUB: launch_fwd<FQSS_PW_SUB>(d, grid, s); break;
        case FQSS_PW_MUL: launch_fwd<FQSS_PW_MUL>(d, grid, s); break;
        default: launch_fwd<FQSS_PW_GLN>(d, grid, s); break;
    }
    return check_launch("pw_fwd");
}

int fqss_pw_bwd(const fqss_pw_desc* dp, const fqss_pw_grads* op, void* ws, size_t ws_bytes, void* stream) {
    int rc = validate(dp, "pw_bwd");
    if (rc) return rc;
    FQSS_REQUIRE(op && op->g && op->ldg >= dp->cols, -1, "pw_bwd: bad gradient descriptor");
    const fqss_pw_desc& d = *dp;
    const fqss_pw_grads& o = *op;
    const int64_t B = d.kind == FQSS_PW_GLN ? d.rows / d.C : 0;
    size_t need = 64 + (d.kind == FQSS_PW_GLN ? (size_t)(2 * d.rows + 2 * B) * sizeof(double) : 0);
    FQSS_REQUIRE(ws && ws_bytes >= need, -3, "pw_bwd: workspace too small (%zu < %zu)", ws_bytes, need);
    cudaStream_t s = (cudaStream_t)stream;
    double* acc = (double*)ws;
    double* rowacc = acc + 8;
    double* samp = rowacc + 2 * d.rows;
    static const char* const names[] = {"pw_bwd(ident)", "pw_bwd(prelu)", "pw_bwd(relu)", "pw_bwd(add)", "pw_bwd(sub)", "pw_bwd(mul)", "pw_bwd(gln)"};
    FQSS_PROFN(names[d.kind], s, d.kind == FQSS_PW_GLN ? 4 : 2);
    cudaMemsetAsync(ws, 0, need, s);
    bool vec = aligned16(d.x1) && aligned16(o.g) && d.ld1 % 4 == 0 && o.ldg % 4 == 0;
    if (o.gx1) vec = vec && aligned16(o.gx1) && o.ldg1 % 4 == 0;
    if (d.x2) vec = vec && aligned16(d.x2) && d.ld2 % 4 == 0;
    if (o.gx2) vec = vec && aligned16(o.gx2) && o.ldg2 % 4 == 0;
    const unsigned chunks = (unsigned)((d.cols + PW_CHUNK - 1) / PW_CHUNK);
    dim3 grid((unsigned)d.rows, chunks);
    switch (d.kind) {
        case FQSS_PW_IDENT: launch_bwd<FQSS_PW_IDENT, 0>(d, o, vec, grid, s, acc, rowacc, samp); break;
        case FQSS_PW_PRELU: launch_bwd<FQSS_PW_PRELU, 0>(d, o, vec, grid, s, acc, rowacc, samp); break;
        case FQSS_PW_RELU: launch_bwd<FQSS_PW_RELU, 0>(d, o, vec, grid, s, acc, rowacc, samp); break;
        case FQSS_PW_ADD: launch_bwd<FQSS_PW_ADD, 0>(d, o, vec, grid, s, acc, rowacc, samp); break;
        case FQSS_PW_SUB: launch_bwd<FQSS_PW_SUB, 0>(d, o, vec, grid, s, acc, rowacc, samp); break;
        case FQSS_PW_MUL: {
            dim3 g2((unsigned)(d.rows / d.bcast), chunks);
            if (vec) pw_mul_bwd_kernel<true><<<g2, PW_THREADS, 0, s>>>(d, o, acc);
            else pw_mul_bwd_kernel<false><<<g2, PW_THREADS, 0, s>>>(d, o, acc);
            break;
        }
        default: {
            launch_bwd<FQSS_PW_GLN, 1>(d, o, vec, grid, s, acc, rowacc, samp);
            int nb = (int)B + (d.C + 255) / 256;
            gln_reduce_kernel<<<nb, 256, 0, s>>>(rowacc, (int)B, d.C, d.gamma, o.g_gamma, o.g_beta, samp);
            launch_bwd<FQSS_PW_GLN, 2>(d, o, vec, grid, s, acc, rowacc, samp);
            break;
        }
    }
    pw_finalize_kernel<<<1, 1, 0, s>>>(acc, o.g_rmin, o.g_rmax, d.kind == FQSS_PW_PRELU ? o.g_slope : nullptr, d.n_bits,
                                       d.quant);
    return check_launch("pw_bwd");
}

int fqss_gln_stats(const float* x, int64_t rows, int64_t cols, int64_t ld, int C, double* stats, void* stream) {
    FQSS_REQUIRE(x && stats && rows > 0 && cols > 0 && ld >= cols && C > 0 && rows % C == 0, -1, "gln_stats: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    FQSS_PROF("gln_stats", s);
    cudaMemsetAsync(stats, 0, (size_t)(rows / C) * 2 * sizeof(double), s);
    dim3 grid((unsigned)rows, (unsigned)((cols + PW_CHUNK - 1) / PW_CHUNK));
    if (aligned16(x) && ld % 4 == 0) gln_stats_kernel<true><<<grid, PW_THREADS, 0, s>>>(x, cols, ld, C, stats);
    else gln_stats_kernel<false><<<grid, PW_THREADS, 0, s>>>(x, cols, ld, C, stats);
    return check_launch("gln_stats");
}

}  // extern "C"
