// gemm_tc.cu -- the 1x1 convolutions of the fused TCN path as persistent, warp-specialised
// tcgen05 GEMMs (north_star item (c)).
//
//   D[frame, o] = sum_k A[frame, k] * Wc[o, k]          (frames on the 128 TMEM lanes, channels on columns)
//
// * A = activations [B][K][Mp] in bf16, read by TMA straight from the NCL layout as an MN-major operand
//   (64-frame x 64-channel boxes, 128B swizzle); frames beyond M are zero-filled by TMA.
// * Wc = weights [N][K] in bf16, K-major.  In the quantised model both operands are INTEGER CODES
//   (activations 0..255, weights -128..127): every product and every partial sum (< 2^24) is exact in
//   the fp32 accumulator, so the result is independent of accumulation order and differs from the
//   reference's fp32 conv only by the final affine  y = s1[o]*acc + s0[o]  (two roundings).
// * Epilogue (8 warps): tcgen05.ld -> registers -> fused tail (affine, PReLU, fake-quant, residual /
//   skip adds, gLN statistics) -> coalesced global stores: one warp-store covers 32 consecutive frames
//   of one channel row, so no shared-memory transpose is needed.
// * Roles: warp 0 TMA producer, warp 1 MMA issuer (one elected lane), warp 2 TMEM allocator,
//   warps 4-11 epilogue.  4-stage smem ring (A 16 KB + B up to 32 KB per stage), 2 TMEM accumulator
//   stages so the epilogue of tile i overlaps the MMAs of tile i+1.  Grid = one CTA per SM.
#include <cuda.h>
#include <stdlib.h>

#include "fqss_common.cuh"
#include "tc_common.cuh"
#include <type_traits>

#include "gemm_tc.cuh"
#include "tcn_common.cuh"

namespace fqss {

int num_sms();

namespace tcg {

using namespace tc;

constexpr int BM = 128;          // frames per tile (TMEM lanes)
constexpr int BK = 64;           // channels per k-chunk (128 B of bf16 along K for the weights)
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;       // 16 KB: two 64-frame boxes of 64 rows x 128 B
constexpr int EPI_WARPS = 16;      // 4 per TMEM lane quarter: the fused tails are latency-bound, not issue-bound
constexpr int NUM_THREADS = 128 + EPI_WARPS * 32;
constexpr int MAXN = 1024;
constexpr int STAT_MAXB = 256;     // samples whose gLN statistics are accumulated in shared memory (EPI_EXPAND)
constexpr int STAT_PRIVB = 64;     // up to this many samples every epilogue warp owns a PRIVATE row of the accumulators: plain
                                   // read-modify-write instead of fp64 shared-memory atomics, which are CAS loops and, with the 16
                                   // warps of a CTA finishing the same tile together, took ~10 % of the expand GEMM's stall samples
constexpr int STAT_SMEM = (2 * STAT_MAXB > EPI_WARPS * 2 * STAT_PRIVB ? 2 * STAT_MAXB : EPI_WARPS * 2 * STAT_PRIVB) * 8;

struct __align__(8) Barriers {
    uint64_t full[STAGES];
    uint64_t empty[STAGES];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
};

template <int NT>
__host__ __device__ constexpr int stage_bytes() { return A_STAGE_BYTES + NT * BK * 2; }

template <int NT>
__host__ __device__ constexpr int smem_bytes() { return STAGES * stage_bytes<NT>() + 2 * MAXN * 4 + (int)sizeof(Barriers) + STAT_SMEM + 1024; }

__device__ __forceinline__ float prelu(float y, float a) { return y > 0.f ? y : a * y; }

// LDC > 0: the row pitch is the compile-time constant LDC (the recipe's 4 s / 8 kHz segments give M = 3999, pitch
// 4000): the per-column addresses of the epilogue become immediates of ONE base pointer per tensor instead of a
// 64-bit pointer bump per column and tensor (12 of the ~64 instructions per output of the res/skip tail).
template <int NT, int EPI, int LDC>
__global__ void __launch_bounds__(NUM_THREADS, 1)
pw_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Args p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by POINTER arithmetic on the shared array: the compiler keeps the address space, so reads of the
    // per-channel constants below are LDS (an integer round trip would turn every one of them into a generic load)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* s1s = reinterpret_cast<float*>(smem + STAGES * stage_bytes<NT>());
    float* s0s = s1s + MAXN;
    Barriers* bar = reinterpret_cast<Barriers*>(s0s + MAXN);
    double* stat_sm = (EPI == EPI_EXPAND && p.B <= STAT_MAXB) ? reinterpret_cast<double*>(bar + 1) : nullptr;
    const bool stat_priv = stat_sm && p.B <= STAT_PRIVB;
    if (stat_sm) {
        for (int i = threadIdx.x; i < 2 * p.B * (stat_priv ? EPI_WARPS : 1); i += NUM_THREADS) stat_sm[i] = 0.0;
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = (p.M + BM - 1) / BM;
    const int nt = p.N / NT;
    const int num_tiles = p.B * mt * nt;
    const int kchunks = p.K / BK;
    const int a_rows = p.a_rows > 0 ? p.a_rows : p.K;

    for (int i = threadIdx.x; i < p.N; i += NUM_THREADS) {
        s1s[i] = p.s1 ? __ldg(p.s1 + i) : 1.f;
        s0s[i] = p.s0 ? __ldg(p.s0 + i) : 0.f;
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bar->full[s], 1);
            mbar_init(&bar->empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bar->tmem_full[a], 1);
            mbar_init(&bar->tmem_empty[a], EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&bar->tmem_base, 2 * NT);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bar->tmem_base;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int n_idx = t % nt, mrow = t / nt;
                const int b = mrow / mt, m0 = (mrow % mt) * BM;
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&bar->empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * stage_bytes<NT>();
                    uint8_t* sb = sa + A_STAGE_BYTES;
                    mbar_expect_tx(&bar->full[stage], stage_bytes<NT>());
                    const int ak = (kc * BK) % a_rows;
                    tma_load_3d(sa, &tmA, &bar->full[stage], m0, ak, b);
                    tma_load_3d(sa + A_STAGE_BYTES / 2, &tmA, &bar->full[stage], m0 + 64, ak, b);
                    tma_load_2d(sb, &tmB, &bar->full[stage], kc * BK, n_idx * NT);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_bf16(BM, NT, /*A MN-major*/ true, /*B K-major*/ false);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                mbar_wait(&bar->tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * NT);
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&bar->full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * stage_bytes<NT>());
                    const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // A (MN-major, SW128): 16 K-rows = 2048 B per UMMA_K step; LBO = next 64-frame box, SBO = 8 rows
                        const uint64_t adesc = smem_desc_sw128(sa + k * 2048, A_STAGE_BYTES / 2, 1024);
                        // B (K-major, SW128): 32 B per UMMA_K step inside the 128 B row; SBO = 8 rows
                        const uint64_t bdesc = smem_desc_sw128(sb + k * 32, 16, 1024);
                        umma_bf16(tmem_d, adesc, bdesc, idesc, (kc | k) ? 1u : 0u);
                    }
                    umma_commit(&bar->empty[stage]);
                    if (kc == kchunks - 1) umma_commit(&bar->tmem_full[acc]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue =================
        // One warp = 32 consecutive frames (TMEM lanes) x CW-column chunks of its column group.  The chunk loop is
        // software-pipelined: while chunk c is computed, the TMEM load and the global loads (residual / skip-sum /
        // addend) of chunk c+1 are already in flight.  Pointers advance by one row pitch per column, so a warp
        // store is one coalesced 128 B row segment.
        const int q = warp & 3;                 // TMEM lane quarter this warp may touch
        const int h = (warp - 4) >> 2;          // column group
        constexpr int COLS = NT / (EPI_WARPS / 4);
        constexpr bool HAS_PRE = (EPI == EPI_RESSKIP || EPI == EPI_ADD || EPI == EPI_RELU_MUL);
        constexpr int CW = HAS_PRE ? 8 : 16;    // columns per chunk (two chunks live in registers)
        static_assert(COLS % (2 * CW) == 0, "chunk pipeline needs an even number of chunks");
        ActQF q1, qres, qskip, qadd, qadds, qm, qp;
        float slope = 0.f;
        if (EPI == EPI_RELU_MUL && p.quant) {
            qm = load_actqf(p.qm_min, p.qm_max, 8);
            qp = load_actqf(p.qp_min, p.qp_max, 8);
        }
        if (EPI == EPI_EXPAND) {
            slope = __ldg(p.slope);
            if (p.quant) q1 = load_actqf(p.q1_min, p.q1_max, 8);
        }
        if (EPI == EPI_RESSKIP && p.quant) {
            if (p.N > p.n_res) qskip = load_actqf(p.qskip_min, p.qskip_max, 8);      // no skip columns: skip-less block
            if (p.n_res) {
                qres = load_actqf(p.qres_min, p.qres_max, 8);
                qadd = load_actqf(p.qadd_min, p.qadd_max, 8);
            }
            if (!p.first_block && p.N > p.n_res) qadds = load_actqf(p.qadds_min, p.qadds_max, 8);
        }
        const int64_t ld = LDC > 0 ? (int64_t)LDC : p.ld;
        int acc = 0;
        uint32_t acc_phase = 0;
        int fold_b = -1;
        float fold_rstd = 1.f, fold_nrm = 0.f;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            const int n_idx = t % nt, mrow = t / nt;
            const int b = mrow / mt, m0 = (mrow % mt) * BM;
            const int m = m0 + q * 32 + lane;
            const bool valid = m < p.M;
            float st_s = 0.f, st_ss = 0.f;
            unsigned st_c = 0u, st_cc = 0u;
            if (EPI == EPI_RESSKIP && p.fold_stats && b != fold_b) {      // per-sample constants of the folded gLN
                const double mean = __ldg(p.fold_stats + 2 * b) / p.n_elems;
                double var = __ldg(p.fold_stats + 2 * b + 1) / p.n_elems - mean * mean;
                var = var > 0.0 ? var : 0.0;
                fold_rstd = (float)(1.0 / sqrt(var + (double)GLN_EPS));
                fold_nrm = -fold_rstd * (float)mean;
                fold_b = b;
            }
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NT + h * COLS);
            const int obase = n_idx * NT + h * COLS;      // first output channel of this warp's column group

            // global operands of the tail for the chunk starting at column c0 of the group
            auto load_pre = [&](float (&pre)[CW], int c0) {
                if (!HAS_PRE) return;
                const int o0 = obase + c0;
                const float* src = nullptr;
                if (EPI == EPI_ADD) {
                    src = p.addend + ((int64_t)b * p.N + o0) * ld + m;
                } else if (EPI == EPI_RELU_MUL) {
                    src = p.addend + ((int64_t)b * p.mul_C + (o0 % p.mul_C)) * ld + m;      // mul_C % 32 == 0
                } else {
                    const bool is_res = o0 < p.n_res;
                    const int n_loc = is_res ? p.n_res : p.N - p.n_res;
                    const int64_t i0 = ((int64_t)b * n_loc + (is_res ? o0 : o0 - p.n_res)) * ld + m;
                    src = is_res ? p.x_in + i0 : (p.first_block ? nullptr : p.skip_in + i0);
                }
#pragma unroll
                for (int j = 0; j < CW; ++j) pre[j] = (valid && src) ? __ldg(src + j * ld) : 0.f;
            };

            auto compute = [&](const uint32_t (&v)[CW], const float (&pre)[CW], int c0) {
                if (!valid) return;
                const int o0 = obase + c0;
                const float* s1c = s1s + o0;
                const float* s0c = s0s + o0;
                if (EPI == EPI_STORE) {
                    if (p.n_store > 0 && o0 >= p.n_store) return;        // warp-uniform: columns beyond the kept ones
                    const int64_t base = ((int64_t)b * (p.n_store > 0 ? p.n_store : p.N) + o0) * ld + m;
                    float* of = p.out_f32 ? p.out_f32 + base : nullptr;
                    __nv_bfloat16* ob = p.out_bf16 ? p.out_bf16 + base : nullptr;
#pragma unroll
                    for (int j = 0; j < CW; ++j) {
                        const float y = fmaf(__uint_as_float(v[j]), s1c[j], s0c[j]);
                        if (of) of[j * ld] = y;
                        if (ob) ob[j * ld] = __float2bfloat16_rn(y);
                    }
                } else if (EPI == EPI_BF16) {
                    __nv_bfloat16* ob = p.out_bf16 + ((int64_t)b * p.N + o0) * ld + m;
#pragma unroll
                    for (int j = 0; j < CW; ++j) ob[j * ld] = __float2bfloat16_rn(fmaf(__uint_as_float(v[j]), s1c[j], s0c[j]));
                } else if (EPI == EPI_EXPAND) {
                    float* of = p.out_f32 + ((int64_t)b * p.N + o0) * ld + m;
                    if (p.quant && !p.out_f32) {
                        // inference: y1 is only needed by backward -- emit the FQ1 codes and the statistics, nothing else
                        uint8_t* oc = p.code1 + ((int64_t)b * p.N + o0) * ld + m;
#pragma unroll
                        for (int j = 0; j < CW; ++j) {
                            const float y = fmaf(__uint_as_float(v[j]), s1c[j], s0c[j]);
                            const unsigned c = code_u8(actqf_t(q1, prelu_f(y, slope)));
                            oc[j * ld] = (uint8_t)c;
                            st_c += c;
                            st_cc += c * c;
                        }
                    } else if (p.quant) {
                        // gLN statistics of a1 = delta1 * code + min1 from INTEGER code sums (exact adds; expanded once per
                        // tile): prelu, (z - min) / delta, one saturating conversion, IADD + IMAD per element
                        uint8_t* oc = p.code1 ? p.code1 + ((int64_t)b * p.N + o0) * ld + m : nullptr;
                        if (oc) {
                            // the EXACT code of FQ1 (same arithmetic as every later consumer), stored for the depthwise
                            // kernel and for backward: the hidden tensor is re-read as 1 B/frame instead of 4
#pragma unroll
                            for (int j = 0; j < CW; ++j) {
                                const float y = fmaf(__uint_as_float(v[j]), s1c[j], s0c[j]);
                                of[j * ld] = y;
                                const unsigned c = code_u8(actqf_t(q1, prelu_f(y, slope)));
                                oc[j * ld] = (uint8_t)c;
                                st_c += c;
                                st_cc += c * c;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < CW; ++j) {
                                const float y = fmaf(__uint_as_float(v[j]), s1c[j], s0c[j]);
                                of[j * ld] = y;
                                const unsigned c = code_u8(__fmul_rn(__fadd_rn(prelu(y, slope), -q1.mn), q1.inv));
                                st_c += c;
                                st_cc += c * c;
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < CW; ++j) {
                            const float y = fmaf(__uint_as_float(v[j]), s1c[j], s0c[j]);
                            of[j * ld] = y;
                            const float a = prelu(y, slope);
                            st_s += a;
                            st_ss = fmaf(a, a, st_ss);
                        }
                    }
                } else if (EPI == EPI_ADD || EPI == EPI_RELU_MUL) {
                    float* of = p.out_f32 + ((int64_t)b * p.N + o0) * ld + m;
                    if (EPI == EPI_RELU_MUL && p.quant) {
                        // mask head of the quantised model: FQ_m(relu(y)) * features -> FQ_p, both quantisers exact
                        float* ys = p.y_save ? p.y_save + ((int64_t)b * p.N + o0) * ld + m : nullptr;
                        __nv_bfloat16* oc = p.out_bf16 ? p.out_bf16 + ((int64_t)b * p.N + o0) * ld + m : nullptr;
#pragma unroll
                        for (int j = 0; j < CW; ++j) {
                            const float y = fmaf(__uint_as_float(v[j]), s1c[j], s0c[j]);
                            if (ys) ys[j * ld] = y;
                            const float vm = actqf_fq(qm, fmaxf(y, 0.f));
                            const float c = actqf_code(qp, __fmul_rn(vm, pre[j]));
                            of[j * ld] = actqf_decode(qp, c);
                            if (oc) oc[j * ld] = __float2bfloat16_rn(c);      // the decoder GEMM's operand: the code itself
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < CW; ++j) {
                            const float y = fmaf(__uint_as_float(v[j]), s1c[j], s0c[j]);
                            of[j * ld] = (EPI == EPI_ADD) ? y + pre[j] : fmaxf(y, 0.f) * pre[j];
                        }
                    }
                } else if (EPI == EPI_RESSKIP) {
                    const bool is_res = o0 < p.n_res;           // warp-uniform: a chunk never straddles the two convs
                    const int n_loc = is_res ? p.n_res : p.N - p.n_res;
                    const int64_t i0 = ((int64_t)b * n_loc + (is_res ? o0 : o0 - p.n_res)) * ld + m;
                    if (is_res) {                               // residual conv -> FQ -> (x + res) -> FQ
                        float* ry = p.res_y ? p.res_y + i0 : nullptr;
                        float* xo = p.x_out + i0;
                        // CTA-uniform switches resolved once per chunk (see the skip branch below)
                        using T = std::true_type;
                        using F = std::false_type;
                        const bool save = ry != nullptr;
                        if (p.quant) {
                            __nv_bfloat16* xop = p.x_out_op + i0;
                            auto cols = [&](auto save_tag) {
                                constexpr bool SAVE = decltype(save_tag)::value;
#pragma unroll
                                for (int j = 0; j < CW; ++j) {
                                    const float y = fmaf(__uint_as_float(v[j]), s1c[j], s0c[j]);
                                    if (SAVE) ry[j * ld] = y;
                                    const float z = __fadd_rn(pre[j], actqf_fq(qres, y));
                                    const float c = actqf_code(qadd, z);
                                    xo[j * ld] = actqf_decode(qadd, c);
                                    xop[j * ld] = __float2bfloat16_rn(c);
                                }
                            };
                            if (save) cols(T{}); else cols(F{});
                        } else {
                            __nv_bfloat16* xop = p.x_out_op + (p.split ? ((int64_t)b * 2 * p.n_res + o0) * ld + m : i0);
                            const int64_t lo_off = (int64_t)p.n_res * ld;
                            auto cols = [&](auto save_tag, auto fold_tag, auto split_tag) {
                                constexpr bool SAVE = decltype(save_tag)::value, FOLD = decltype(fold_tag)::value, SPLIT = decltype(split_tag)::value;
#pragma unroll
                                for (int j = 0; j < CW; ++j) {
                                    const float y = FOLD ? fmaf(__uint_as_float(v[j]), fold_rstd, fmaf(fold_nrm, s1c[j], s0c[j]))
                                                         : fmaf(__uint_as_float(v[j]), s1c[j], s0c[j]);
                                    if (SAVE) ry[j * ld] = y;
                                    const float z = __fadd_rn(pre[j], y);
                                    xo[j * ld] = z;
                                    const __nv_bfloat16 hi = __float2bfloat16_rn(z);
                                    xop[j * ld] = hi;
                                    if (SPLIT) xop[j * ld + lo_off] = __float2bfloat16_rn(z - __bfloat162float(hi));
                                }
                            };
                            const bool fold = p.fold_stats != nullptr, split = p.split != 0;
                            if (save) {
                                if (fold) { if (split) cols(T{}, T{}, T{}); else cols(T{}, T{}, F{}); }
                                else { if (split) cols(T{}, F{}, T{}); else cols(T{}, F{}, F{}); }
                            } else {
                                if (fold) { if (split) cols(F{}, T{}, T{}); else cols(F{}, T{}, F{}); }
                                else { if (split) cols(F{}, F{}, T{}); else cols(F{}, F{}, F{}); }
                            }
                        }
                    } else {                                    // skip conv -> FQ -> (skip_sum + skip) -> FQ
                        float* sy = p.skip_y ? p.skip_y + i0 : nullptr;
                        float* so = p.skip_out + i0;
                        // The CTA-uniform switches are resolved ONCE per chunk, outside the column loop: with them inside, every
                        // column was its own basic block (three uniform branches per element in the SASS) and the quantiser chains
                        // of the CW columns -- division, conversion, FMA, each a ~20-cycle dependent chain -- could not overlap
                        auto cols = [&](auto q_tag, auto first_tag, auto save_tag, auto fold_tag) {
                            constexpr bool Q = decltype(q_tag)::value, FIRST = decltype(first_tag)::value;
                            constexpr bool SAVE = decltype(save_tag)::value, FOLD = decltype(fold_tag)::value;
#pragma unroll
                            for (int j = 0; j < CW; ++j) {
                                const float y = FOLD ? fmaf(__uint_as_float(v[j]), fold_rstd, fmaf(fold_nrm, s1c[j], s0c[j]))
                                                     : fmaf(__uint_as_float(v[j]), s1c[j], s0c[j]);
                                if (SAVE) sy[j * ld] = y;
                                const float sk = Q ? actqf_fq(qskip, y) : y;
                                if (FIRST) {
                                    so[j * ld] = sk;
                                } else {
                                    const float z = __fadd_rn(pre[j], sk);
                                    so[j * ld] = Q ? actqf_fq(qadds, z) : z;
                                }
                            }
                        };
                        using T = std::true_type;
                        using F = std::false_type;
                        const bool save = sy != nullptr;
                        if (p.quant) {
                            if (p.first_block) { if (save) cols(T{}, T{}, T{}, F{}); else cols(T{}, T{}, F{}, F{}); }
                            else { if (save) cols(T{}, F{}, T{}, F{}); else cols(T{}, F{}, F{}, F{}); }
                        } else if (p.fold_stats) {
                            if (p.first_block) { if (save) cols(F{}, T{}, T{}, T{}); else cols(F{}, T{}, F{}, T{}); }
                            else { if (save) cols(F{}, F{}, T{}, T{}); else cols(F{}, F{}, F{}, T{}); }
                        } else {
                            if (p.first_block) { if (save) cols(F{}, T{}, T{}, F{}); else cols(F{}, T{}, F{}, F{}); }
                            else { if (save) cols(F{}, F{}, T{}, F{}); else cols(F{}, F{}, F{}, F{}); }
                        }
                    }
                }
            };

            uint32_t va[CW], vb[CW];
            float pa[CW], pb[CW];
            load_pre(pa, 0);                          // independent of the accumulator: issue before waiting for the MMAs
            mbar_wait(&bar->tmem_full[acc], acc_phase);
            tc_fence_after();
            tmem_ld_cols<CW>(tbase, va);
#pragma unroll 1
            for (int c0 = 0; c0 < COLS; c0 += 2 * CW) {
                tmem_ld_wait();
                tmem_ld_cols<CW>(tbase + (uint32_t)(c0 + CW), vb);
                load_pre(pb, c0 + CW);
                compute(va, pa, c0);
                tmem_ld_wait();
                if (c0 + 2 * CW < COLS) {
                    tmem_ld_cols<CW>(tbase + (uint32_t)(c0 + 2 * CW), va);
                    load_pre(pa, c0 + 2 * CW);
                }
                compute(vb, pb, c0 + CW);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar->tmem_empty[acc]);
            if (EPI == EPI_EXPAND) {
                double ds, dss;
                if (p.quant) {
                    // per lane <= 64 columns: sum c <= 16 320, sum c^2 <= 4.2e6; per warp x32: both fit 32 bits
                    const unsigned wc = __reduce_add_sync(0xffffffffu, st_c), wcc = __reduce_add_sync(0xffffffffu, st_cc);
                    const unsigned nvalid = __popc(__ballot_sync(0xffffffffu, valid)) * COLS;
                    const double dl = (double)q1.delta, mn = (double)q1.mn, n = (double)nvalid;
                    ds = dl * (double)wc + n * mn;
                    dss = dl * dl * (double)wcc + 2.0 * dl * mn * (double)wc + n * mn * mn;
                } else {
                    ds = (double)warp_sum(st_s);
                    dss = (double)warp_sum(st_ss);
                }
                if (lane == 0) {
                    if (stat_priv) {
                        double* row = stat_sm + (warp - 4) * 2 * p.B + 2 * b;
                        row[0] += ds;
                        row[1] += dss;
                    } else if (stat_sm) {
                        atomicAdd(stat_sm + 2 * b, ds);
                        atomicAdd(stat_sm + 2 * b + 1, dss);
                    } else {
                        atomicAdd(p.stats + 2 * b, ds);
                        atomicAdd(p.stats + 2 * b + 1, dss);
                    }
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (EPI == EPI_EXPAND && stat_sm) {
        // one pair of global atomics per (CTA, sample) instead of one per (warp, tile): the per-sample accumulators
        // are otherwise hit by every epilogue warp of every CTA working on the same sample at the same time
        for (int i = threadIdx.x; i < 2 * p.B; i += NUM_THREADS) {
            double v = stat_sm[i];
            if (stat_priv)
                for (int w = 1; w < EPI_WARPS; ++w) v += stat_sm[w * 2 * p.B + i];      // fixed order
            if (v != 0.0) atomicAdd(p.stats + i, v);
        }
    }
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * NT);
    }
    if (EPI == EPI_EXPAND && p.rc) {
        RowConstJob job;
        job.stats = p.stats; job.rc = p.rc; job.B = p.B; job.n_elems = p.n_elems;
        job.qa_min = p.q1_min; job.qa_max = p.q1_max; job.qb_min = p.q2_min; job.qb_max = p.q2_max;
        job.qc_min = p.q3_min; job.qc_max = p.q3_max;
        rowconst_last_cta(job, gridDim.x, (int)threadIdx.x < 2 * p.B || !stat_sm);
    }
}

// ---------------------------------------------------------------------------------------------
// host: tensor maps + launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// resolved through the runtime so that the library has no link-time dependency on libcuda
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

static int make_act_map(CUtensorMap* tm, const void* base, int B, int C, int M, int64_t ld) {
    cuuint64_t dims[3] = {(cuuint64_t)M, (cuuint64_t)C, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)C * (cuuint64_t)ld * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)BK, 1};
    cuuint32_t es[3] = {1, 1, 1};
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return -999;
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

static int make_w_map(CUtensorMap* tm, const void* base, int N, int K, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return -999;
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

constexpr int LD_HOT = 4000;      // pitch of the recipe's segments (M = 3999): specialised epilogue addressing

template <int NT, int EPI, int LDC>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const Args& a, cudaStream_t s) {
    static bool configured = false;
    constexpr int smem = smem_bytes<NT>();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(pw_gemm_kernel<NT, EPI, LDC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            set_error("pw_gemm: cannot set %d B dynamic smem: %s", smem, cudaGetErrorString(e));
            return -4;
        }
        configured = true;
    }
    const int mt = (a.M + BM - 1) / BM;
    const int tiles = a.B * mt * (a.N / NT);
    int grid = num_sms();
    if (grid > tiles) grid = tiles;
    static const char* const names[] = {"gemm_store", "gemm_expand", "gemm_resskip", "gemm_dgrad_bf16", "gemm_dgrad_add", "gemm_relu_mul"};
    FQSS_PROF(a.prof ? a.prof : (a.a_rows > 0 && a.a_rows < a.K) ? (EPI == EPI_EXPAND ? "gemm_expand(split3)" : EPI == EPI_RESSKIP ? "gemm_resskip(split3)" : EPI == EPI_RELU_MUL ? "gemm_relu_mul(split3)" : "gemm_store(split3)") : names[EPI], s);
    pw_gemm_kernel<NT, EPI, LDC><<<grid, NUM_THREADS, smem, s>>>(ta, tb, a);
    return check_launch("pw_gemm");
}

int run(int epi, const void* act_bf16, const void* w_bf16, const Args& a, cudaStream_t s) {
    FQSS_REQUIRE(act_bf16 && w_bf16, -1, "pw_gemm: null operand");
    FQSS_REQUIRE(a.B > 0 && a.M > 0 && a.K >= BK && a.K % BK == 0 && a.N >= 128 && a.N % 128 == 0 && a.N <= MAXN, -1,
                 "pw_gemm: unsupported shape B=%d M=%d K=%d N=%d (K %% 64 == 0, N %% 128 == 0, N <= %d)", a.B, a.M, a.K, a.N, MAXN);
    FQSS_REQUIRE(a.ld >= a.M && a.ld % 8 == 0, -2, "pw_gemm: row pitch must be a multiple of 8 elements (TMA 16 B strides)");
    FQSS_REQUIRE(aligned16(act_bf16) && aligned16(w_bf16), -2, "pw_gemm: operands must be 16-byte aligned");
    FQSS_REQUIRE(a.a_rows >= 0 && a.a_rows <= a.K && a.a_rows % BK == 0, -1, "pw_gemm: a_rows=%d must be a multiple of %d and <= K", a.a_rows, BK);
    CUtensorMap ta, tb;
    int r = make_act_map(&ta, act_bf16, a.B, a.a_rows > 0 ? a.a_rows : a.K, a.M, a.ld);
    FQSS_REQUIRE(r == 0, -4, "pw_gemm: cuTensorMapEncodeTiled(activations) failed (%d)", r);
    const bool wide = (a.N % 256 == 0);
    r = make_w_map(&tb, w_bf16, a.N, a.K, wide ? 256 : 128);
    FQSS_REQUIRE(r == 0, -4, "pw_gemm: cuTensorMapEncodeTiled(weights) failed (%d)", r);
    static const bool ld_spec = !(getenv("FQSS_GEMM_LDSPEC") && atoi(getenv("FQSS_GEMM_LDSPEC")) == 0);
    const bool hot = ld_spec && a.ld == LD_HOT;
#define FQSS_GEMM_CASE(E)                                                                                      \
    case E:                                                                                                    \
        if (hot && E != EPI_EXPAND) /* measured: the expand tail is faster with pointer bumps */              \
            return wide ? launch<256, E, LD_HOT>(ta, tb, a, s) : launch<128, E, LD_HOT>(ta, tb, a, s);         \
        return wide ? launch<256, E, 0>(ta, tb, a, s) : launch<128, E, 0>(ta, tb, a, s);
    switch (epi) {
        FQSS_GEMM_CASE(EPI_STORE)
        FQSS_GEMM_CASE(EPI_EXPAND)
        FQSS_GEMM_CASE(EPI_RESSKIP)
        FQSS_GEMM_CASE(EPI_BF16)
        FQSS_GEMM_CASE(EPI_ADD)
        FQSS_GEMM_CASE(EPI_RELU_MUL)
    }
#undef FQSS_GEMM_CASE
    set_error("pw_gemm: unknown epilogue %d", epi);
    return -1;
}

}  // namespace tcg
}  // namespace fqss

using namespace fqss;

extern "C" {

// Generic entry (also the isolated test vehicle of the tensor path):
//   out[b,o,m] = s1[o] * sum_k act[b,k,m] * w[o,k] + s0[o]     (+ addend[b,o,m] when given)
// act, w: bf16.  out_f32 and/or out_bf16 may be given.
int fqss_pw_gemm(const void* act_bf16, const void* w_bf16, const float* s1, const float* s0, float* out_f32, void* out_bf16,
                 const float* addend, int B, int K, int N, int M, int64_t ld, int a_rows, void* stream) {
    return fqss_pw_gemm_ex(act_bf16, w_bf16, s1, s0, out_f32, out_bf16, addend, 0, B, K, N, M, ld, a_rows, stream);
}

// mul_C > 0: out_f32[b,o,m] = relu(s1*acc + s0) * addend[b, o % mul_C, m]   (mask head of the float model)
int fqss_pw_gemm_ex(const void* act_bf16, const void* w_bf16, const float* s1, const float* s0, float* out_f32, void* out_bf16,
                    const float* addend, int mul_C, int B, int K, int N, int M, int64_t ld, int a_rows, void* stream) {
    tcg::Args a{};
    a.B = B; a.M = M; a.K = K; a.N = N; a.ld = ld; a.s1 = s1; a.s0 = s0; a.quant = 0; a.a_rows = a_rows;
    a.out_f32 = out_f32; a.out_bf16 = (__nv_bfloat16*)out_bf16; a.addend = addend;
    int epi = tcg::EPI_STORE;
    if (mul_C > 0) {
        FQSS_REQUIRE(out_f32 && addend && N % mul_C == 0 && mul_C % 32 == 0, -1,
                     "pw_gemm_ex: relu-mul needs an fp32 output, a multiplicand, N %% mul_C == 0 and mul_C %% 32 == 0");
        a.mul_C = mul_C;
        epi = tcg::EPI_RELU_MUL;
    } else if (addend) {
        FQSS_REQUIRE(out_f32, -1, "pw_gemm: addend needs an fp32 output");
        epi = tcg::EPI_ADD;
    } else if (!out_f32) {
        FQSS_REQUIRE(out_bf16, -1, "pw_gemm: no output given");
        epi = tcg::EPI_BF16;
    }
    return tcg::run(epi, act_bf16, w_bf16, a, (cudaStream_t)stream);
}

// fqss_pw_gemm with only the first n_store output channels stored (out_f32: [B][n_store][ld]); see fqss.h
int fqss_pw_gemm_nstore(const void* act_bf16, const void* w_bf16, const float* s1, const float* s0, float* out_f32, int n_store,
                        int B, int K, int N, int M, int64_t ld, void* stream) {
    FQSS_REQUIRE(out_f32 && n_store > 0 && n_store <= N && n_store % 16 == 0, -1, "pw_gemm_nstore: n_store must be a multiple of 16 in (0, N]");
    tcg::Args a{};
    a.B = B; a.M = M; a.K = K; a.N = N; a.ld = ld; a.s1 = s1; a.s0 = s0; a.quant = 0; a.a_rows = 0;
    a.out_f32 = out_f32; a.n_store = n_store; a.prof = "gemm_frames";
    return tcg::run(tcg::EPI_STORE, act_bf16, w_bf16, a, (cudaStream_t)stream);
}

// Mask head of the quantised model in the GEMM epilogue (see fqss.h).
int fqss_mask_head_fwd(const void* x_op_bf16, const void* w_bf16, const float* s1, const float* s0, const float* feats, int C,
                       const float* qm_min, const float* qm_max, const float* qp_min, const float* qp_max, float* y_save,
                       float* masked, void* masked_codes_bf16, int B, int K, int N, int M, int64_t ld, void* stream) {
    FQSS_REQUIRE(feats && masked && qm_min && qm_max && qp_min && qp_max, -1, "mask_head_fwd: null argument");
    FQSS_REQUIRE(C > 0 && N % C == 0 && C % 32 == 0, -1, "mask_head_fwd: N=%d must be a multiple of C=%d, C a multiple of 32", N, C);
    tcg::Args a{};
    a.B = B; a.M = M; a.K = K; a.N = N; a.ld = ld; a.s1 = s1; a.s0 = s0; a.quant = 1; a.a_rows = 0;
    a.out_f32 = masked; a.out_bf16 = (__nv_bfloat16*)masked_codes_bf16; a.addend = feats; a.mul_C = C; a.y_save = y_save;
    a.qm_min = qm_min; a.qm_max = qm_max; a.qp_min = qp_min; a.qp_max = qp_max;
    return tcg::run(tcg::EPI_RELU_MUL, x_op_bf16, w_bf16, a, (cudaStream_t)stream);
}

}  // extern "C"
