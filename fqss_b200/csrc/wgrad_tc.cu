// wgrad_tc.cu -- weight gradients of the 1x1 convolutions as a split-K tcgen05 GEMM.
//
//   G[o, i] = sum_{b, m} dY[b, o, m] * X[b, i, m]          (reduction over ALL frames of ALL samples)
//
// Both operands are K-major (frames are contiguous in the NCL layout), so TMA boxes of
// {64 frames x rows} with 128B swizzle feed the MMA directly.  The output is tiny (<= 256x512) and
// the reduction is long (B*M ~ 128K), so the work is split along K: every CTA owns a contiguous range
// of 64-frame chunks, keeps a [MT*128] x [NW] fp32 accumulator in TMEM (MT*NW <= 512 columns) for the
// whole range and writes ONE partial tile at the end; after a grid-wide barrier (the grid is co-resident) the same kernel
// reduces the partials deterministically (split order), every CTA one slice, and applies the de-quantisation affine:
//   dWq[o,i] = (da * G[o,i]) / dws[o] + min_a * db[o]
// (X holds integer codes c with x = da*c + min_a; dY was pre-scaled by dws[o] for the dgrad GEMM).
#include <cuda.h>

#include <cstdlib>

#include "fqss_common.cuh"
#include "tc_common.cuh"
#include "wgrad_tc.cuh"

namespace fqss {

int num_sms();

namespace tcw {

using namespace tc;

constexpr int BKF = 64;                 // frames per k-chunk (128 B of bf16)
constexpr int NUM_THREADS = 256;        // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warps 4-7 epilogue

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

struct __align__(8) Bars {
    uint64_t full[8];
    uint64_t empty[8];
    uint64_t done;
    uint32_t tmem_base;
};

template <int MT, int NW>
__host__ __device__ constexpr int stage_bytes() { return (MT * 128 + NW) * BKF * 2; }
// as many stages as fit ~208 KB (one CTA per SM): the loop is latency-bound, not tensor-bound, so what counts is the number
// of bytes in flight per SM (ncu r01e: 2-3 stages of 64-80 KB left the TMA pipeline starved, dram 47 %)
template <int MT, int NW, int OCC = 1>
__host__ __device__ constexpr int num_stages() {
    return (208 * 1024 / OCC) / stage_bytes<MT, NW>() > 8 ? 8 : (208 * 1024 / OCC) / stage_bytes<MT, NW>();
}
template <int MT, int NW>
__host__ __device__ constexpr uint32_t tmem_cols() { return MT * NW <= 128 ? 128u : (MT * NW <= 256 ? 256u : 512u); }
template <int MT, int NW, int OCC = 1>
__host__ __device__ constexpr int smem_bytes() { return num_stages<MT, NW, OCC>() * stage_bytes<MT, NW>() + (int)sizeof(Bars) + 1024; }

struct KArgs {
    int B, M, O, I;           // O rows of dY, I rows of X
    int chunks_per_sample;    // ceil(M / 64)
    int nsplit;               // gridDim.x
    float* part;              // [nsplit][O][I]
    // fused finalisation (after a grid-wide barrier every CTA reduces its slice of the partial tiles in split order)
    unsigned int* ctr;        // zeroed arrival counter
    const float* amin;
    const float* amax;
    const float* dws;
    const double* db;
    float* dWq;
};

template <int MT, int NW, int OCC>
__global__ void __launch_bounds__(NUM_THREADS, OCC)
wgrad_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX, const KArgs p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by POINTER arithmetic on the shared array: the compiler keeps the address space, so reads of the
    // per-channel constants below are LDS (an integer round trip would turn every one of them into a generic load)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int ST = num_stages<MT, NW, OCC>();
    constexpr int SB = stage_bytes<MT, NW>();
    Bars* bar = reinterpret_cast<Bars*>(smem + ST * SB);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // blockIdx.y enumerates (o-group, i-chunk)
    const int n_ichunks = p.I / NW;
    const int og = blockIdx.y / n_ichunks, ic = blockIdx.y % n_ichunks;
    const int o0 = og * MT * 128, i0 = ic * NW;
    const int total = p.B * p.chunks_per_sample;
    const int per = (total + p.nsplit - 1) / p.nsplit;
    const int cbeg = blockIdx.x * per;
    const int cend = min(total, cbeg + per);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmY);
        tma_prefetch_desc(&tmX);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < ST; ++s) {
            mbar_init(&bar->full[s], 1);
            mbar_init(&bar->empty[s], 1);
        }
        mbar_init(&bar->done, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&bar->tmem_base, tmem_cols<MT, NW>());
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bar->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int c = cbeg; c < cend; ++c) {
                const int b = c / p.chunks_per_sample, f0 = (c % p.chunks_per_sample) * BKF;
                mbar_wait(&bar->empty[stage], phase ^ 1);
                uint8_t* sa = smem + stage * SB;
                mbar_expect_tx(&bar->full[stage], SB);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) tma_load_3d(sa + mt * 128 * BKF * 2, &tmY, &bar->full[stage], f0, o0 + mt * 128, b);
                tma_load_3d(sa + MT * 128 * BKF * 2, &tmX, &bar->full[stage], f0, i0, b);
                if (++stage == ST) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_bf16(128, NW, false, false);
            int stage = 0;
            uint32_t phase = 0;
            for (int c = cbeg; c < cend; ++c) {
                mbar_wait(&bar->full[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * SB);
                const uint32_t sb = sa + MT * 128 * BKF * 2;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                    for (int k = 0; k < BKF / 16; ++k) {
                        const uint64_t adesc = smem_desc_sw128(sa + mt * 128 * BKF * 2 + k * 32, 16, 1024);
                        const uint64_t bdesc = smem_desc_sw128(sb + k * 32, 16, 1024);
                        umma_bf16(tmem_base + (uint32_t)(mt * NW), adesc, bdesc, idesc, (c > cbeg || k > 0) ? 1u : 0u);
                    }
                }
                umma_commit(&bar->empty[stage]);
                if (++stage == ST) { stage = 0; phase ^= 1; }
            }
            umma_commit(&bar->done);
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        float* out = p.part + (int64_t)blockIdx.x * p.O * p.I;
        if (cend > cbeg) {
            mbar_wait(&bar->done, 0);
            tc_fence_after();
        }
#pragma unroll 1
        for (int mt = 0; mt < MT; ++mt) {
            const int o = o0 + mt * 128 + q * 32 + lane;
#pragma unroll 1
            for (int c0 = 0; c0 < NW; c0 += 32) {
                uint32_t v[32];
                if (cend > cbeg) {
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * NW + c0), v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0u;
                }
                float4* dst = reinterpret_cast<float4*>(out + (int64_t)o * p.I + i0 + c0);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                         __uint_as_float(v[4 * j + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols<MT, NW>());
    }
    // ---- grid-wide barrier (all CTAs are co-resident: the grid never exceeds OCC CTAs per SM), then the deterministic
    // reduction of the partial tiles, distributed over the whole grid:
    //   dWq[o,i] = (da * sum_s part[s][o][i]) / dws[o] + min_a * db[o]          (summed in split order s = 0, 1, ...)
    const unsigned int nctas = gridDim.x * gridDim.y;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(p.ctr, 1u);
        unsigned int seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.ctr) : "memory");
            if (seen < nctas) __nanosleep(100);
        } while (seen < nctas);
        __threadfence();
    }
    __syncthreads();
    {
        const int64_t n = (int64_t)p.O * p.I;
        const int64_t nq = n >> 2;                              // float4 quads (I is a multiple of 128)
        float da = 1.f, mn = 0.f;
        if (p.amin) {
            mn = __ldg(p.amin);
            da = __fdiv_rn(__fsub_rn(__ldg(p.amax), mn), 255.f);
        }
        const int64_t cta = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
        for (int64_t qd = cta * NUM_THREADS + threadIdx.x; qd < nq; qd += (int64_t)nctas * NUM_THREADS) {
            const int64_t e = qd << 2;
            const int o = (int)(e / p.I);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int sp = 0; sp < p.nsplit; ++sp) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(p.part + (int64_t)sp * n + e));
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            const float ds = __ldg(p.dws + o), zb = mn * (float)p.db[o];
            float4 r;
            r.x = (da * acc.x) / ds + zb;
            r.y = (da * acc.y) / ds + zb;
            r.z = (da * acc.z) / ds + zb;
            r.w = (da * acc.w) / ds + zb;
            *reinterpret_cast<float4*>(p.dWq + e) = r;
        }
    }
}

static int make_map(CUtensorMap* tm, const void* base, int B, int C, int M, int64_t ld, int box_rows) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return -999;
    cuuint64_t dims[3] = {(cuuint64_t)M, (cuuint64_t)C, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)C * (cuuint64_t)ld * 2};
    cuuint32_t box[3] = {(cuuint32_t)BKF, (cuuint32_t)box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

template <int MT, int NW, int OCC = 1>
static int launch(const CUtensorMap& ty, const CUtensorMap& tx, const KArgs& a, int gy, cudaStream_t s) {
    static bool configured = false;
    constexpr int smem = smem_bytes<MT, NW, OCC>();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_kernel<MT, NW, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            set_error("wgrad: cannot set %d B dynamic smem: %s", smem, cudaGetErrorString(e));
            return -4;
        }
        configured = true;
    }
    FQSS_PROF("wgrad", s);
    wgrad_kernel<MT, NW, OCC><<<dim3(a.nsplit, gy), NUM_THREADS, smem, s>>>(ty, tx, a);
    return check_launch("wgrad");
}

// tile shape per CTA: MT 128-row blocks of dY x NW rows of X.  Smaller tiles = smaller stages = more stages (bytes) in flight.
// FQSS_WGRAD_CFG (development knob): 0 = the round-1 shapes (largest tile that fits TMEM), 1 = <2,128> / <2,128>,
// 2 = <2,128> / <1,256>, 3 (default, measured best: 52 / 62 us vs 71 / 77 for cfg 0) = <1,128> everywhere with 6 stages,
// 4 = <1,128> with two CTAs per SM (3 stages each; slower).
static void pick_tile(int O, int I, int* MT, int* NW, int* OCC) {
    static const int cfg = getenv("FQSS_WGRAD_CFG") ? atoi(getenv("FQSS_WGRAD_CFG")) : 3;
    const int ot = O / 128;
    *OCC = 1;
    if (cfg == 4) {
        *NW = 128;
        *MT = 1;
        *OCC = 2;
        return;
    }
    if (cfg == 0) {
        *NW = I >= 256 ? 256 : 128;
        *MT = (ot >= 4 && *NW == 128) ? 4 : (ot >= 2 ? 2 : 1);
    } else if (cfg == 3) {
        *NW = 128;
        *MT = 1;
    } else if (cfg == 2 && I >= 256 && I > O) {
        *NW = 256;
        *MT = 1;
    } else {
        *NW = 128;
        *MT = ot >= 2 ? 2 : 1;
    }
}

int plan_splits(int B, int M, int O, int I) {
    int MT, NW, OCC;
    pick_tile(O, I, &MT, &NW, &OCC);
    const int gy = (O / 128 / MT) * (I / NW);
    int ns = OCC * num_sms() / gy;
    const int total = B * ((M + BKF - 1) / BKF);
    if (ns > total) ns = total;
    if (ns < 1) ns = 1;
    return ns;
}

size_t part_bytes(int B, int M, int O, int I) { return (size_t)plan_splits(B, M, O, I) * O * I * sizeof(float) + 256; }

int run(const void* dY, const void* X, int B, int M, int64_t ld, int O, int I, float* part, size_t part_cap, const float* amin,
        const float* amax, const float* dws, const double* db, float* dWq, cudaStream_t s) {
    FQSS_REQUIRE(dY && X && part && dws && db && dWq, -1, "wgrad: null argument");
    FQSS_REQUIRE(O % 128 == 0 && I % 128 == 0 && O >= 128 && I >= 128, -1, "wgrad: O and I must be multiples of 128 (O=%d I=%d)", O, I);
    FQSS_REQUIRE(ld >= M && ld % 8 == 0, -2, "wgrad: bad pitch");
    int MT, NW, OCC;
    pick_tile(O, I, &MT, &NW, &OCC);
    const int ot = O / 128;
    FQSS_REQUIRE(I % NW == 0 && ot % MT == 0, -1, "wgrad: O=%d I=%d do not tile by %dx%d", O, I, MT * 128, NW);
    const int gy = (ot / MT) * (I / NW);
    KArgs a;
    a.B = B; a.M = M; a.O = O; a.I = I;
    a.chunks_per_sample = (M + BKF - 1) / BKF;
    a.nsplit = plan_splits(B, M, O, I);
    a.part = part;
    FQSS_REQUIRE(part_cap >= (size_t)a.nsplit * O * I * sizeof(float) + 256, -3, "wgrad: partial buffer too small");
    a.ctr = reinterpret_cast<unsigned int*>(part + (size_t)a.nsplit * O * I);
    a.amin = amin; a.amax = amax; a.dws = dws; a.db = db; a.dWq = dWq;
    FQSS_REQUIRE(a.nsplit * gy <= OCC * num_sms(), -4, "wgrad: grid of %d CTAs would not be co-resident", a.nsplit * gy);
    if (cudaMemsetAsync(a.ctr, 0, 2 * sizeof(unsigned int), s) != cudaSuccess) return check_launch("wgrad(counter memset)");
    CUtensorMap ty, tx;
    int r = make_map(&ty, dY, B, O, M, ld, 128);
    FQSS_REQUIRE(r == 0, -4, "wgrad: cuTensorMapEncodeTiled(dY) failed (%d)", r);
    r = make_map(&tx, X, B, I, M, ld, NW);
    FQSS_REQUIRE(r == 0, -4, "wgrad: cuTensorMapEncodeTiled(X) failed (%d)", r);
    if (NW == 128) {
        if (MT == 4) r = launch<4, 128>(ty, tx, a, gy, s);
        else if (MT == 2) r = launch<2, 128>(ty, tx, a, gy, s);
        else if (OCC == 2) r = launch<1, 128, 2>(ty, tx, a, gy, s);
        else r = launch<1, 128>(ty, tx, a, gy, s);
    } else {
        if (MT == 2) r = launch<2, 256>(ty, tx, a, gy, s);
        else r = launch<1, 256>(ty, tx, a, gy, s);
    }
    return r;
}

}  // namespace tcw
}  // namespace fqss

using namespace fqss;

extern "C" {

size_t fqss_wgrad_codes_ws_bytes(int B, int M, int O, int I) { return tcw::part_bytes(B, M, O, I) + 256; }

// Weight gradient of a code-operand 1x1 conv (see fqss.h).
int fqss_wgrad_codes(const void* dY_bf16, const void* x_op_bf16, int B, int M, int64_t ld, int O, int I, const float* amin,
                     const float* amax, const float* dws, const double* db, float* dWq, void* ws, size_t ws_bytes, void* stream) {
    FQSS_REQUIRE(ws && ws_bytes >= fqss_wgrad_codes_ws_bytes(B, M, O, I), -3, "wgrad_codes: workspace too small");
    float* part = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    return tcw::run(dY_bf16, x_op_bf16, B, M, ld, O, I, part, ws_bytes - 256, amin, amax, dws, db, dWq, (cudaStream_t)stream);
}

}  // extern "C"
