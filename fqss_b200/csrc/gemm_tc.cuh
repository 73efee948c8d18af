// gemm_tc.cuh -- interface of the tcgen05 1x1-conv GEMM (gemm_tc.cu) for the other translation units.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fqss {
namespace tcg {

enum { EPI_STORE = 0, EPI_EXPAND = 1, EPI_RESSKIP = 2, EPI_BF16 = 3, EPI_ADD = 4, EPI_RELU_MUL = 5 };

struct Args {
    int B, M, K, N;              // batch, valid frames, reduction channels, output channels
    int a_rows;                  // activation rows per sample (0: = K); reduction index k reads row k % a_rows
    int split;                   // EPI_RESSKIP, float model: x_out_op is a [hi ; lo] bf16 pair ([B][2*n_res][ld])
    int64_t ld;                  // row pitch (elements) of every [.,.,Mp] activation tensor involved
    const float* s1;             // [N] per-output-channel scale   (delta_w[o]*delta_a, or 1)
    const float* s0;             // [N] per-output-channel offset  (delta_w[o]*min_a*R[o] + bias[o])
    int quant;                   // 0: float model (no fake-quant in the tails)
    // EPI_STORE / EPI_EXPAND / EPI_ADD: fp32 output [B][N][ld]
    float* out_f32;
    // EPI_BF16 (and optional for STORE): bf16 output [B][N][ld]
    __nv_bfloat16* out_bf16;
    // EPI_ADD: addend [B][N][ld].  EPI_RELU_MUL: multiplicand [B][mul_C][ld], out = relu(y) * addend[b][o % mul_C][m]
    const float* addend;
    int mul_C;
    // EPI_RELU_MUL, quantised model (the mask head: Conv1dNlQ(ReLU) + MulQ, convtasnetq.py:97-99,203): out =
    // FQ_p(FQ_m(relu(y)) * addend); y_save (may be NULL) keeps the pre-activation for backward
    const float* qm_min; const float* qm_max; const float* qp_min; const float* qp_max;
    float* y_save;
    // EPI_STORE: > 0 limits the stored output channels to [0, n_store) (a multiple of 16; the decoder GEMM computes 128
    // columns for the tensor core's shape and keeps the 16 taps)
    int n_store;
    const char* prof;      // profiler class of the launch (NULL: named after the epilogue)
    // EPI_EXPAND: gLN statistics of FQ(PReLU(y)) -> stats[2*B] (double, pre-zeroed)
    const float* slope;
    const float* q1_min; const float* q1_max;
    double* stats;
    uint8_t* code1;              // EPI_EXPAND, quantised: exact 8-bit codes of FQ1(PReLU(y)) [B][N][ld] (may be NULL)
    // EPI_EXPAND: the last CTA turns the finished statistics into the row constants rc[RC_HDR + 2*B] (tcn_common.cuh)
    float* rc; double n_elems;
    const float* q2_min; const float* q2_max; const float* q3_min; const float* q3_max;
    // EPI_RESSKIP: columns [0,Nres) = residual conv, [Nres,N) = skip conv
    int n_res;                   // 128, or 0 for the last block (no residual path)
    int first_block;             // 1: skip accumulator starts here (no adds-quantiser)
    float* res_y; float* skip_y;                 // pre-quant conv outputs (saved for backward) [B][128][ld]
    const float* x_in;                           // block input (fake-quantised values) [B][128][ld]
    float* x_out; __nv_bfloat16* x_out_op;       // block output: values and GEMM operand (codes, or values when !quant)
    const float* skip_in; float* skip_out;       // running skip sum [B][128][ld]
    const float* qres_min; const float* qres_max;
    const float* qskip_min; const float* qskip_max;
    const float* qadd_min; const float* qadd_max;
    const float* qadds_min; const float* qadds_max;
    // EPI_RESSKIP, float model with the second gLN folded into this conv (fqss_tcn_prep_fold): the operand is
    // a3 = PReLU(y3), s1 = u, s0 = v, and y = rstd_b * (acc - mu_b * u[o]) + v[o] with {mu_b, rstd_b} from these
    // finished statistics ({sum, sum of squares} per sample over n_elems elements).  NULL: plain affine.
    const double* fold_stats;
};


// epi: EPI_*; act_bf16 [B][K][ld], w_bf16 [N][K].  Returns 0 or a negative fqss error code.
int run(int epi, const void* act_bf16, const void* w_bf16, const Args& a, cudaStream_t s);

}  // namespace tcg
}  // namespace fqss
