/* fqss.h -- C ABI of libfqss_sm100.so: the B200 (sm_100a) kernels behind the FQSS
 * `quantization/qat` operator API (ssi-research/FQSS).
 *
 * The reference has no native code; every entry point below replaces a chain of ATen library
 * calls issued by a reference Python function (cited as file:line under the reference tree).
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; all tensor pointers are DEVICE pointers owned by the caller
 *     (PyTorch caching allocator); the library never allocates, frees or retains device memory.
 *   - activations are "row tensors": `rows` rows of `cols` fp32 values, row r starting at
 *     base + r*ld (ld >= cols, in elements).  A [B,C,M] tensor has rows=B*C, cols=M.  The fast
 *     (128-bit) path needs ld % 4 == 0 and a 16-byte aligned base; other layouts run a scalar path.
 *   - quantiser ranges (min_range / max_range Parameters) are read on the device; no host syncs.
 *   - every call is asynchronous and enqueues on `stream` (a cudaStream_t passed as void*).
 *   - return 0 on success; <0 on error: -1 bad argument, -2 misaligned, -3 workspace too small,
 *     -4 CUDA launch error.  fqss_last_error() returns a thread-local message.
 *   - `ws` is caller-provided scratch (device); required size from fqss_ws_bytes().
 */
#ifndef FQSS_H_
#define FQSS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FQSS_ABI_VERSION 21

int fqss_abi_version(void);
const char* fqss_last_error(void);
/* scratch bytes sufficient for ANY single call below on tensors of up to `rows` rows. */
size_t fqss_ws_bytes(int64_t rows);

/* ---------------------------------------------------------------------------------------------
 * Q2  activation fake-quant, standalone  (qat_quant.py:136-147 linear_quantize sym=False;
 *     module GradientActivationFakeQuantize.forward :227-242)
 *   y = delta*clip(rint((x-min)/delta),0,2^b-1)+min ; code (optional, may be NULL) = the clipped integer
 *   bwd: gx = ((g*delta)*mask)/delta ; g_rmin/g_rmax = range gradients (overwritten, 1 float each)
 * ------------------------------------------------------------------------------------------- */
int fqss_fq_act_fwd(const float* x, float* y, uint8_t* code, int64_t n,
                    const float* rmin, const float* rmax, int n_bits, void* stream);
int fqss_fq_act_bwd(const float* g, const float* x, float* gx, float* g_rmin, float* g_rmax, int64_t n,
                    const float* rmin, const float* rmax, int n_bits, void* ws, size_t ws_bytes, void* stream);

/* Q3  weight fake-quant, symmetric signed per-channel (qat_quant.py:126-135; module :350-381).
 *     w viewed as [outer, ch, inner]; ranges have `ch` entries (ch_out_idx=0: outer=1; =1: outer=d0). */
int fqss_fq_weight_fwd(const float* w, float* wq, int8_t* code, int outer, int ch, int inner,
                       const float* rmin, const float* rmax, int n_bits, void* stream);
int fqss_fq_weight_bwd(const float* g, const float* w, float* gw, float* g_rmin, float* g_rmax,
                       int outer, int ch, int inner, const float* rmin, const float* rmax, int n_bits,
                       void* stream);
/* first-call observer: max_range = amax, min_range = amin over non-channel dims (qat_quant.py:373-375) */
int fqss_weight_observe(const float* w, int outer, int ch, int inner, float* rmin, float* rmax, void* stream);

/* Batched forms of the two calls above: one launch covers up to 48 weight tensors (the QAT step has ~100 of them, each
 * a few thousand elements -- launch-bound one by one).  `items` is a HOST array. */
typedef struct fqss_wq_item {
    const float* g;      /* bwd: dL/d(fake-quantised weight); fwd: unused (NULL)             */
    const float* w;      /* raw weight                                                        */
    float* out;          /* bwd: dL/dw (may be NULL); fwd: fake-quantised weight              */
    float* g_rmin;       /* bwd only: range gradients, `ch` entries each                      */
    float* g_rmax;
    const float* rmin;
    const float* rmax;
    int32_t outer, ch, inner, n_bits;
} fqss_wq_item;
int fqss_fq_weight_fwd_batch(const fqss_wq_item* items, int n, void* stream);
int fqss_fq_weight_bwd_batch(const fqss_wq_item* items, int n, void* stream);

/* Q4  activation observer: min <- a*min + (1-a)*x.min(), max likewise (qat_quant.py:228-232) */
int fqss_act_observe(const float* x, int64_t rows, int64_t cols, int64_t ld, float* rmin, float* rmax,
                     double alpha, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * L1  fused "op -> nonlinearity -> fake-quant" pointwise layers (qat_layers.py: AddQ :62, MulQ :86,
 *     GroupNormQ :438, NlQ :511, and the nl+FQ tail of Conv1dQ :124 / Conv1dNlQ :188)
 * ------------------------------------------------------------------------------------------- */
enum {
    FQSS_PW_IDENT = 0, /* z = x1                         (tail of Conv1dQ)            */
    FQSS_PW_PRELU = 1, /* z = x1>0 ? x1 : slope*x1       (Conv1dNlQ / NlQ with PReLU)  */
    FQSS_PW_RELU = 2,  /* z = max(x1,0)                                                */
    FQSS_PW_ADD = 3,   /* z = x1 + x2                    (AddQ)                        */
    FQSS_PW_SUB = 4,   /* z = x1 - x2                    (RQB: Y - Y_q)                */
    FQSS_PW_MUL = 5,   /* z = x1 * x2, x2 broadcast over `bcast` sources (MulQ)        */
    FQSS_PW_GLN = 6    /* z = gLN(x1) = (x1-mu_b)*rstd_b*gamma_c + beta_c (GroupNormQ) */
};

typedef struct fqss_pw_desc {
    int32_t kind;    /* FQSS_PW_*                                                       */
    int32_t quant;   /* 1: y = FQ(z) with (rmin,rmax,n_bits); 0: y = z (observer/float) */
    int32_t n_bits;
    int32_t C;       /* channels per sample: channel of row r is r % C, sample r / C    */
    int32_t bcast;   /* MUL: x1 has rows [B,bcast,C], x2 has rows [B,C]                 */
    int32_t _pad;
    int64_t rows, cols;
    const float* x1; int64_t ld1;
    const float* x2; int64_t ld2;
    float* y;        int64_t ldy;
    const float* slope;   /* PRELU: 1 element                                           */
    const float* gamma;   /* GLN: C                                                     */
    const float* beta;    /* GLN: C                                                     */
    const double* stats;  /* GLN: per sample {sum, sum of squares} from fqss_gln_stats   */
    float eps;
    float _pad2;
    const float* rmin; const float* rmax;
} fqss_pw_desc;

int fqss_pw_fwd(const fqss_pw_desc* d, void* stream);

/* backward of fqss_pw_fwd.  g = dL/dy (rows x cols, ld ldg).  Outputs (each may be NULL if unused):
 *   gx1 (ld ldg1), gx2 (SUB/MUL only; MUL: rows/bcast rows), g_rmin/g_rmax (1), g_slope (1),
 *   g_gamma/g_beta (C).  Small outputs are OVERWRITTEN. */
typedef struct fqss_pw_grads {
    const float* g; int64_t ldg;
    float* gx1; int64_t ldg1;
    float* gx2; int64_t ldg2;
    float* g_rmin; float* g_rmax;
    float* g_slope;
    float* g_gamma; float* g_beta;
} fqss_pw_grads;

int fqss_pw_bwd(const fqss_pw_desc* d, const fqss_pw_grads* o, void* ws, size_t ws_bytes, void* stream);

/* per-sample {sum, sumsq} over C*cols elements -> stats[2*B] (double), for gLN (GroupNorm(1,C)) */
int fqss_gln_stats(const float* x, int64_t rows, int64_t cols, int64_t ld, int C, double* stats, void* stream);

/* ---------------------------------------------------------------------------------------------
 * L1/L2  convolutions of the separator (F.conv1d / F.conv_transpose1d call sites in
 *        qat_layers.py:138,203,1030,1189,1194,1332).  Weights are the already fake-quantised values.
 * ------------------------------------------------------------------------------------------- */
/* 1x1 conv:  y[b,o,m] = bias[o] + sum_i w[o,i] x[b,i,m]      (fp32 FFMA path, "rel 1e-3" tier) */
int fqss_conv1x1_fwd(const float* x, int64_t ldx, const float* w, const float* bias, float* y, int64_t ldy,
                     int B, int Ci, int Co, int M, void* stream);
/* gx[b,i,m] = sum_o w[o,i] gy[b,o,m] */
int fqss_conv1x1_dgrad(const float* gy, int64_t ldgy, const float* w, float* gx, int64_t ldgx,
                       int B, int Ci, int Co, int M, void* stream);
/* gw[o,i] = sum_{b,m} gy[b,o,m] x[b,i,m] ; gbias[o] = sum_{b,m} gy[b,o,m]  (gbias may be NULL) */
int fqss_conv1x1_wgrad(const float* gy, int64_t ldgy, const float* x, int64_t ldx, float* gw, float* gbias,
                       int B, int Ci, int Co, int M, void* ws, size_t ws_bytes, void* stream);

/* depthwise k-tap dilated conv, zero padding = dil*(k-1)/2 ("same"): convtasnetq.py:28-29 */
int fqss_dwconv_fwd(const float* x, int64_t ldx, const float* w, const float* bias, float* y, int64_t ldy,
                    int B, int C, int M, int K, int dil, void* stream);
int fqss_dwconv_bwd(const float* gy, int64_t ldgy, const float* x, int64_t ldx, const float* w,
                    float* gx, int64_t ldgx, float* gw, float* gbias,
                    int B, int C, int M, int K, int dil, void* ws, size_t ws_bytes, void* stream);

/* encoder-type strided conv, no padding, no bias: y[b,o,m] = sum_{c,k} w[o,c,k] x[b,c,m*stride+k] */
int fqss_sconv_fwd(const float* x, int64_t ldx, const float* w, float* y, int64_t ldy,
                   int B, int Cin, int Co, int T, int K, int stride, void* stream);
/* gx (may be NULL) [B,Cin,T]; gw (may be NULL) [Co,Cin,K] */
int fqss_sconv_bwd(const float* gy, int64_t ldgy, const float* x, int64_t ldx, const float* w,
                   float* gx, int64_t ldgx, float* gw, int B, int Cin, int Co, int T, int K, int stride,
                   void* ws, size_t ws_bytes, void* stream);

/* decoder-type transposed conv to ONE output channel (overlap-add), weight [Ci,1,K]:
 *   y[b,t] = sum_{c,m,k: m*stride+k=t} w[c,k] x[b,c,m] */
int fqss_tconv_fwd(const float* x, int64_t ldx, const float* w, float* y, int64_t ldy,
                   int B, int Ci, int M, int K, int stride, void* stream);
int fqss_tconv_bwd(const float* gy, int64_t ldgy, const float* x, int64_t ldx, const float* w,
                   float* gx, int64_t ldgx, float* gw, int B, int Ci, int M, int K, int stride,
                   void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Tensor-core (tcgen05 / TMEM / TMA) 1x1 convolution used by the fused TCN path:
 *   out[b,o,m] = s1[o] * sum_k act[b,k,m] * w[o,k] + s0[o]  (+ addend[b,o,m])
 * act [B][a_rows][ld] and w [N][K] are bf16.  With integer fake-quant codes as operands the accumulation is
 * exact (order independent); s1/s0 carry the de-quantisation affine and the bias.  K % 64 == 0,
 * N % 128 == 0, ld % 8 == 0.  Outputs: out_f32 and/or out_bf16 ([B][N][ld]); addend needs out_f32.  s1 / s0 may be
 * NULL (= all ones / all zeros).
 * a_rows = 0 means a_rows = K.  a_rows < K (a_rows % 64 == 0) makes reduction index k read activation row
 * k % a_rows: with act = [hi ; lo] (a_rows = 2C, x = hi + lo in bf16 pairs) and w = [w_hi | w_hi | w_lo]
 * (K = 3C) one launch computes the three-term split product hi*w_hi + lo*w_hi + hi*w_lo, i.e. an fp32-grade
 * (2^-16) contraction on the bf16 tensor pipe -- the float teacher's 1x1 convolutions.
 * ------------------------------------------------------------------------------------------- */
int fqss_pw_gemm(const void* act_bf16, const void* w_bf16, const float* s1, const float* s0, float* out_f32,
                 void* out_bf16, const float* addend, int B, int K, int N, int M, int64_t ld, int a_rows, void* stream);

/* as fqss_pw_gemm; mul_C > 0 selects the float mask-head epilogue (convtasnetq.py:97-99,204-205):
 *   out_f32[b,o,m] = relu(s1[o]*acc + s0[o]) * addend[b, o % mul_C, m],  addend = encoder features [B][mul_C][ld] */
int fqss_pw_gemm_ex(const void* act_bf16, const void* w_bf16, const float* s1, const float* s0, float* out_f32,
                    void* out_bf16, const float* addend, int mul_C, int B, int K, int N, int M, int64_t ld, int a_rows,
                    void* stream);

/* fp32 [rows][ldw] -> bf16 split pair for the float (teacher) GEMMs.
 *   layout 0 (weights [N][K]):   out [N][3K]  = [hi | hi | lo]      (rows = N, cols = K)
 *   layout 1 (activations [B][C][M]): out [B][2C][ldo] = per sample [hi rows ; lo rows]  (rows = B*C, cols = M) */
int fqss_split_bf16(const float* x, int64_t ldx, void* out_bf16, int64_t ldo, int64_t rows, int cols, int C, int layout,
                    void* stream);

/* ---------------------------------------------------------------------------------------------
 * Code-operand 1x1 convolution outside the ConvBlocks (bottleneck conv, mask conv: convtasnetq.py:67-70,
 * 97-99 as Conv1dQ / Conv1dNlQ, qat_layers.py:137-141, 202-207): the input is the output of an 8-bit
 * activation quantiser, so it is re-encoded to integer codes (fqss_tcn_encode), the weights are prepared by
 * fqss_tcn_prep and the product runs on fqss_pw_gemm.  Backward:
 *   fqss_rowscale_bf16: dY[b,o,m] = bf16(dws[o] * gy[b,o,m]), db[o] = sum gy (fp64; may be NULL)
 *   input gradient    : fqss_pw_gemm(dY, WcT, ones, zeros) -- gx[b,i,m] = sum_o code_w[o,i] dY[b,o,m]
 *   fqss_wgrad_codes  : dWq[o,i] = (delta_a * sum_{b,m} dY[b,o,m] code_a[b,i,m]) / dws[o] + min_a * db[o]
 *                       (gradient w.r.t. the FAKE-QUANTISED weight; feed it to fqss_fq_weight_bwd)
 * ------------------------------------------------------------------------------------------- */
int fqss_rowscale_bf16(const float* g, int64_t ldg, void* out_bf16, int64_t ldo, int64_t rows, int M, int C, const float* scale,
                       double* rowsum, void* stream);
size_t fqss_wgrad_codes_ws_bytes(int B, int M, int O, int I);
int fqss_wgrad_codes(const void* dY_bf16, const void* x_op_bf16, int B, int M, int64_t ld, int O, int I, const float* amin,
                     const float* amax, const float* dws, const double* db, float* dWq, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * M1  fused ConvBlock of the TCN (convtasnetq.py:11-42 after quantize_model :270-277), forward and
 *     backward, for the steady state (observers off).  Four forward launches per block (the last CTA of K1 / K2
 *     turns the finished gLN statistics into per-sample constants for the next kernel; stats1 / stats3 hold 2*B sums
 *     plus one slot used as its arrival counter = 2*B + 1 doubles):
 *       K1 expand GEMM (+bias, PReLU/FQ statistics for the first gLN)          x_op  -> y1, stats1
 *       K2 depthwise kernel (PReLU+FQ, gLN+FQ on load; 3-tap dilated FIR; stats) y1   -> y3, stats3
 *       K3a hidden quantiser (PReLU+FQ, gLN+FQ -> bf16 operand)                  y3   -> a4_op
 *       K3 res/skip GEMM (+bias, FQ, residual add + FQ, skip accumulate + FQ)    a4_op-> res_y, skip_y, x_out, skip_out
 *     Tensors: [B][C][ld] with ld % 8 == 0; *_op tensors are bf16 GEMM operands holding the integer
 *     fake-quant codes (quant=1) or the real values (quant=0, the float teacher).  Every q* entry is
 *     {min_range, max_range} device pointers of a GradientActivationFakeQuantize.
 * ------------------------------------------------------------------------------------------- */
typedef struct fqss_qrange { const float* rmin; const float* rmax; } fqss_qrange;

typedef struct fqss_tcn_block {
    int32_t B, M, dil, quant, first_block, has_res, Cio, Chid;
    int32_t split;   /* quant == 0 only: *_op tensors are bf16 [hi ; lo] pairs with 2x the rows, Wc1/Wc2 are
                        [N][3K] = [hi | hi | lo] (fp32-grade float model; forward / inference only).
                        2: additionally the second gLN is folded into the res/skip conv (Wc2, s1_2 = u, s0_2 = v
                        from fqss_tcn_prep_fold): K2 writes a3 = PReLU(y3) straight into a4_op, K3a is skipped and
                        y3 / rc3 are not touched                                                              */
    int32_t no_skip; /* 1: the block has no skip conv and no skip sum (ConvTasNetMusicQ, convtasnetq_music.py:141-199: 1x1 -> PReLU ->
                        gLN -> depthwise -> PReLU -> gLN -> 1x1, plus the residual): the second GEMM is the residual conv alone
                        (Wc2 [Cio][Chid]); needs has_res = 1; skip_in / skip_y / skip_out / qskip / qadds and, in backward,
                        g_skip_out / g_skip_in are ignored (may be NULL); dY2 is [B][Cio][ld], dW2q [Cio][Chid].  A bias-free
                        depthwise conv passes a zero vector as bdw                                                     */
    int64_t ld;
    /* prepared by fqss_tcn_prep (per step) */
    const void* Wc1;  const void* Wc1T; const float* s1_1; const float* s0_1; const float* dws1;   /* expand [Chid,Cio] */
    const void* Wc2;  const void* Wc2T; const float* s1_2; const float* s0_2; const float* dws2;   /* res|skip [2*Cio,Chid] */
    const float* wdw; const float* bdw;                                                          /* depthwise [Chid,3], [Chid] */
    /* layer parameters */
    const float* slope1; const float* slope3;
    const float* gn1_w; const float* gn1_b; const float* gn2_w; const float* gn2_b;
    fqss_qrange q_in, q1, q2, q3, q4, qres, qskip, qadd, qadds;
    /* activations */
    const void* x_op; const float* x_in; const float* skip_in;
    /* y1 / y3 (and res_y / skip_y) are the pre-activations backward re-reads; in quantised INFERENCE they may be NULL:
     * forward then skips those stores (the 8-bit codes carry the data on) and backward refuses to run */
    float* y1; double* stats1; float* y3; double* stats3; void* a4_op;
    float* res_y; float* skip_y; float* x_out; void* x_out_op; float* skip_out;
    /* row constants written by forward right after the statistics are complete and re-read by backward:
     * 16 + 2*B floats each = {min, delta, 1/delta, levels} of up to three quantisers, the two exact clipping thresholds of
     * the first one (z_lo, z_hi), two pad words, then {mean, rstd} per sample
     * (rc1: q1, q2, q3 and gLN1 from stats1; rc3: q3, q4 and gLN2 from stats3) */
    float* rc1; float* rc3;
    /* quantised model: the 8-bit codes of a1 = FQ1(PReLU(y1)) and a3 = FQ3(PReLU(y3)), one byte per frame
     * ([B][Chid][ld]); both REQUIRED.  The expand GEMM's epilogue writes code1 and the depthwise kernel reads it
     * (1 B/frame instead of re-deriving the code from y1); the depthwise kernel writes code3 and the hidden quantiser
     * (code3 -> FQ4 code, 3 B/frame) reads it.  Both are re-read by the backward stages that need only the codes
     * (FQ2/FQ4 masks, gLN sums, tap gradients); y1 / y3 are re-read only where the continuous value matters. */
    uint8_t* code1; uint8_t* code3;
} fqss_tcn_block;

/* Weight preparation for one 1x1 conv of the fused path: fake-quantise W [N][K] per output channel
 * (wmin/wmax NULL: float model, Wc = bf16(W)) and fold the input quantiser (amin/amax NULL: identity)
 * and the bias into the GEMM epilogue constants:
 *   Wc  [N][K] bf16 written at row offset n_off of a [Ntot][K] matrix, WcT [K][Ntot] bf16 (for dgrad),
 *   s1[n_off+o] = dw[o]*da, s0[n_off+o] = dw[o]*min_a*sum_k code[o,k] + bias[o], dws[n_off+o] = dw[o].
 * split != 0 (float model only): Wc is [Ntot][3K] = [hi | hi | lo] bf16 pairs of W, WcT may be NULL. */
int fqss_tcn_prep(const float* W, const float* wmin, const float* wmax, const float* bias, const float* amin,
                  const float* amax, void* Wc, void* WcT, float* s1, float* s0, float* dws, int N, int K, int Ntot,
                  int n_off, int split, void* stream);

/* Float (teacher) model, inference only: fold the block's second gLN (convtasnetq.py:31, GroupNorm(1, Chid)) into the
 * res/skip 1x1 conv that consumes it (convtasnetq.py:33-34).  With a3 = PReLU(y3) and per-sample {mu, rstd}:
 *   conv(gLN(a3))[o] = rstd * (sum_c (W[o,c] gamma[c]) a3[c] - mu * u[o]) + v[o]
 *   u[o] = sum_c W[o,c] gamma[c],  v[o] = bias[o] + sum_c W[o,c] beta[c]        (fp64 sums)
 * Wc [Ntot][3K] = [hi | hi | lo] bf16 split of W*gamma at row offset n_off; u, v [Ntot] go into s1_2 / s0_2 of a
 * block run with split = 2 (the depthwise kernel then writes a3 as the GEMM operand and the normalisation pass over
 * the hidden tensor disappears). */
int fqss_tcn_prep_fold(const float* W, const float* bias, const float* gamma, const float* beta, void* Wc, float* u, float* v,
                       int N, int K, int Ntot, int n_off, void* stream);

/* Batched fqss_tcn_prep: `items` is a HOST array; one launch covers up to 32 convolutions. */
typedef struct fqss_prep_item {
    const float* W; const float* wmin; const float* wmax; const float* bias; const float* amin; const float* amax;
    void* Wc; void* WcT; float* s1; float* s0; float* dws;
    int32_t N, K, Ntot, n_off, split, _pad;
} fqss_prep_item;
int fqss_tcn_prep_batch(const fqss_prep_item* items, int n, void* stream);

/* fp32 values -> bf16 GEMM operand of the first block (codes w.r.t. {rmin,rmax}; NULL ranges: plain cast) */
int fqss_tcn_encode(const float* x, int64_t ldx, void* out_bf16, int64_t ldo, int64_t rows, int M, const float* rmin,
                    const float* rmax, void* stream);

int fqss_tcn_block_fwd(const fqss_tcn_block* blk, void* stream);

/* Backward of one block.  In: g_x_out, g_skip_out (fp32 [B][Cio][ld]).  Out: g_x_in, g_skip_in (may alias the
 * inputs).  Scratch (caller-owned, reusable across blocks): dY2 bf16 [B][2Cio][ld], g_hid_a bf16
 * [B][Chid][ld], dY1 bf16 [B][Chid][ld], g_xd fp32 [B][Cio][ld], wpart fp32 (fqss_tcn_ws_bytes); g_hid_b (bf16
 * [B][Chid][ld]) may be NULL: the gLN2 and depthwise backward stages run as ONE kernel that keeps g_y3 in shared
 * memory (the two-kernel variant, FQSS_SPLIT_P2D=1 in the environment, is a development A/B path).  Parameter
 * gradients are WRITTEN (not accumulated): dW1q [Chid][Cio], db1, dW2q [2Cio][Chid], db2 (w.r.t. the
 * FAKE-QUANTISED weights: run fqss_fq_weight_bwd on them), dwdw [Chid][3] (same), dbdw, g_gn*, slopes,
 * and 2 floats {g_min,g_max} per activation quantiser in g_q (order: q1,q2,q3,q4,qres,qskip,qadd,qadds). */
typedef struct fqss_tcn_block_grads {
    const float* g_x_out; const float* g_skip_out; float* g_x_in; float* g_skip_in;
    void* dY2; void* g_hid_a; void* g_hid_b; void* dY1; float* g_xd;
    float* dW1q; float* db1; float* dW2q; float* db2; float* dwdw; float* dbdw;
    float* g_gn1_w; float* g_gn1_b; float* g_gn2_w; float* g_gn2_b; float* g_slope1; float* g_slope3;
    float* g_q;        /* 16 floats */
    void* ws; size_t ws_bytes;
} fqss_tcn_block_grads;

size_t fqss_tcn_ws_bytes(int B, int Cio, int Chid);
/* 1 (default): the first weight-gradient GEMM of fqss_tcn_block_bwd runs on a library-owned side stream next to the
 * gLN2 row-sum kernel (fork / join through events on the caller's stream; capturable); 0: everything on the caller's
 * stream (what per-kernel timing wants); -1: back to the environment default.  Returns the previous setting. */
int fqss_set_wgrad_overlap(int on);
/* 1: the second weight-gradient GEMM and the finalise kernel of fqss_tcn_block_bwd run on the library's side stream too, so the
 * next block's backward starts right after this block's last dgrad GEMM (scratch reuse ordered by events, accumulator block
 * double-buffered).  The caller MUST then call fqss_tcn_bwd_join(stream) after the last block and before anything reads the
 * parameter gradients.  Default 0 (every call returns with all its work ordered on `stream`).  Returns the previous setting. */
int fqss_set_bwd_tail_side(int on);
int fqss_tcn_bwd_join(void* stream);
int fqss_tcn_block_bwd(const fqss_tcn_block* blk, const fqss_tcn_block_grads* g, void* stream);

/* ---------------------------------------------------------------------------------------------
 * P1  FQSS splitter / reconstructor (process.py:10-52)
 * ------------------------------------------------------------------------------------------- */
/* peak[0] = max |x| over the whole batch (one scalar, process.py:23) */
int fqss_absmax(const float* x, int64_t rows, int64_t cols, int64_t ld, float* peak,
                void* ws, size_t ws_bytes, void* stream);
/* y[b, s, t], s < n_split: successive 8-bit floor quantisations of x[b,t]/peak */
int fqss_split(const float* x, int64_t ldx, const float* peak, float* y, int64_t ldy,
               int B, int T, int n_split, int n_bits, void* stream);
/* y[r,t] = sum_i parts[i][r,t] * (0.5*2^-(n_bits-1))^i ; parts stacked with stride part_stride */
int fqss_combine(const float* parts, int64_t part_stride, int64_t ld, float* y, int64_t ldy,
                 int64_t rows, int T, int n_comb, int n_bits, void* stream);

/* ---------------------------------------------------------------------------------------------
 * S1-S3  FQSS KD SI-SDR loss (mysystem.py:124-151, wsdr.py:46-95, asteroid PIT), n_src == 2
 *   est, fest, tgt: [B,2,T] (ld = row pitch).  out[0]=loss, out[1]=kd_loss(logged), out[2]=val_loss
 *   (mean_b PIT neg-SI-SDR dB of est vs tgt).  gest (may be NULL): dL/dest, same layout as est.
 * ------------------------------------------------------------------------------------------- */
int fqss_kd_loss(const float* est, int64_t lde, const float* fest, int64_t ldf, const float* tgt, int64_t ldt,
                 int B, int T, float kd_lambda, float* out, float* gest, int64_t ldg,
                 void* ws, size_t ws_bytes, void* stream);

/* KD training loss of the music recipe (train_env/tasnet_musdbhq/musdbhq_train.py:87-109; new-SDR process.py:70-75;
 * loss_fn = nn.L1Loss, :249).  wavs / fwavs / sources: [B items][R rows per item][T] row tensors (row pitches ld*);
 * per item a = 10^((nsdr(fwavs,src) - nsdr(wavs,src))/10) (no gradient), loss = (1-lambda) L1(wavs,src) +
 * lambda mean_i a_i L1(wavs_i, fwavs_i); fwavs == NULL or lambda == 0: loss = L1(wavs, src) (:109).
 * out[0] = loss, out[1] = kd term, out[2] = task term; g (may be NULL) = dL/dwavs with pitch ldg. */
size_t fqss_music_loss_ws_bytes(int B);
int fqss_music_kd_loss(const float* wavs, int64_t ldw, const float* fwavs, int64_t ldf, const float* sources, int64_t lds,
                       int B, int R, int T, float kd_lambda, float* out, float* g, int64_t ldg, void* ws, size_t ws_bytes,
                       void* stream);

/* fqss_kd_loss under data parallelism with the loss of the GLOBAL batch (mysystem.py:145 takes
 * -10 log10 of batch MEANS, so the mean of per-rank gradients is not the gradient of the global-batch loss).
 *   phase 0: statistics + the local means {kd, task, val} (3 doubles on the device) into `means`;
 *   the caller averages `means` over the ranks (one all-reduce of 3 doubles; equal shards);
 *   phase 1: out[0..2] from the averaged means; dL/dest with per-sample weights 1/(2 B_local), so that the mean over
 *            ranks of the per-rank gradients (DDP, asteroid_librimix_trainer.py:125-135) is the global-batch gradient.
 * `ws` carries the statistics from phase 0 to phase 1 and must not be touched in between. */
int fqss_kd_loss_dp(const float* est, int64_t lde, const float* fest, int64_t ldf, const float* tgt, int64_t ldt,
                    int B, int T, float kd_lambda, float* out, float* gest, int64_t ldg,
                    void* ws, size_t ws_bytes, double* means, int phase, void* stream);

/* The passes of fqss_kd_loss as separate entry points (PairwiseWSDR as a standalone module, wsdr.py:46-95):
 *   fqss_loss_stats: per sample 24 doubles = SUMS of e0 e1 f0 f1 t0 t1 (divide by T for the means), then the centred
 *     inner products ee0 ee1 ff0 ff1 tt0 tt1 | et00 et01 et10 et11 | ft00 .. ft11 | ef00 .. ef11  (xy_ij = <x_i, y_j>)
 *   fqss_loss_grad_apply: gest[b,i,:] = c_i e_i + sum_j a_ij t_j + b_ij f_j on the CENTRED signals, coef per sample =
 *     16 floats {a00 a01 a10 a11 | b00 b01 b10 b11 | c0 c1 | means e0 e1 f0 f1 t0 t1} */
int fqss_loss_stats(const float* est, int64_t lde, const float* fest, int64_t ldf, const float* tgt, int64_t ldt, int B, int T,
                    double* stats, void* stream);
int fqss_loss_grad_apply(const float* est, int64_t lde, const float* fest, int64_t ldf, const float* tgt, int64_t ldt, int B,
                         int T, const float* coef, float* gest, int64_t ldg, void* stream);

/* ---------------------------------------------------------------------------------------------
 * ConvTasNetMusicQ: the pieces beyond the speech path
 *   fqss_split_ex: process.preprocess for C input channels, normalize = 1 (x / peak, threshold 1; the speech recipe) or 0
 *     (threshold = peak; convtasnetq_music.py:233-234); y[b, s*C + c, :] = s-th part of channel c (process.py:16-36)
 *   fqss_cln_fwd/bwd: channel-wise LayerNorm = nn.LayerNorm(C) over the channel axis of [B, C, M] per frame
 *     (ChannelWiseLayerNorm, convtasnetq_music.py:32-50; LayerNormQ, qat_layers.py:455-468); mean / rstd: [B, M]
 *   fqss_ola_fwd/bwd: overlap_and_add (convtasnetq_music.py:10-30) of rows [R, A*L, K] (Linear-decoder outputs, frame k in
 *     column k) to [R, A, (K-1)*H + L]
 * ------------------------------------------------------------------------------------------- */
int fqss_split_ex(const float* x, int64_t ldx, const float* peak, float* y, int64_t ldy, int B, int C, int T, int n_split,
                  int n_bits, int normalize, void* stream);
int fqss_cln_fwd(const float* x, int64_t ld, const float* gamma, const float* beta, float eps, float* y, int64_t ldy,
                 float* mean, float* rstd, int B, int C, int M, void* stream);
int fqss_cln_bwd(const float* g, int64_t ldg, const float* x, int64_t ld, const float* gamma, const float* mean,
                 const float* rstd, float* gx, int64_t ldgx, float* ggamma, float* gbeta, int B, int C, int M,
                 void* ws, size_t ws_bytes, void* stream);
int fqss_ola_fwd(const float* y, int64_t ldy, float* out, int64_t ldo, int64_t R, int A, int L, int H, int K, void* stream);
int fqss_ola_bwd(const float* gout, int64_t ldo, float* gy, int64_t ldy, int64_t R, int A, int L, int H, int K, void* stream);

/* ---------------------------------------------------------------------------------------------
 * M2  mask head of the quantised separator as ONE GEMM + ONE backward pass (models/convtasnetq.py:97-99: mask_net[1..2] =
 *     Conv1dNlQ(1x1 bn->S*F, ReLU), qat_layers.py:188-212; :203 `self.mul(masks, feats)` = MulQ, qat_layers.py:86-96):
 *       y = s1[o]*sum_k x_op[b,k,m]*w[o,k] + s0[o] ;  mask = FQ_m(relu(y)) ;  masked[b,o,m] = FQ_p(mask * feats[b, o % C, m])
 *     fwd: tcgen05 GEMM on integer-code operands (as fqss_pw_gemm), both quantisers exact in the epilogue; y_save (may be
 *          NULL: inference) keeps the pre-activation for backward.  N = S*C output channels, speaker-major.
 *     bwd: one pass over (g, y_save, feats): dY = bf16(dws[o]*dL/dy) (operand of the dgrad / wgrad GEMMs), g_feats (summed
 *          over the S speakers), g_q = {g_min_m, g_max_m, g_min_p, g_max_p}, bias gradient (g_bias fp32 [N], may be NULL;
 *          db_f64 [N] the same sums in fp64 for fqss_wgrad_codes, may be NULL).
 * ------------------------------------------------------------------------------------------- */
int fqss_mask_head_fwd(const void* x_op_bf16, const void* w_bf16, const float* s1, const float* s0, const float* feats, int C,
                       const float* qm_min, const float* qm_max, const float* qp_min, const float* qp_max, float* y_save,
                       float* masked, void* masked_codes_bf16 /* may be NULL: the FQ_p codes as the decoder GEMM's operand */,
                       int B, int K, int N, int M, int64_t ld, void* stream);
size_t fqss_mask_head_ws_bytes(int N);
int fqss_mask_head_bwd(const float* g, int64_t ldg, const float* y, const float* feats, const float* dws, const float* qm_min,
                       const float* qm_max, const float* qp_min, const float* qp_max, void* dY_bf16, float* g_feats, float* g_q,
                       float* g_bias, double* db_f64, int B, int C, int S, int M, int64_t ld, void* ws, size_t ws_bytes,
                       void* stream);

/* ---------------------------------------------------------------------------------------------
 * L2  the filterbank edges on the tcgen05 GEMM (Conv1dEncoderQ qat_layers.py:993-1046, ConvTr1dDecoderQ :1305-1361,
 *     ResidualErrorBlock :1105-1220).  The transposed conv to one channel is an overlap-add of per-frame tap vectors,
 *       frames[r,k,m] = sum_o w[o,k] Y[r,o,m]   (a GEMM over the filters, taps as output channels; fqss_pw_gemm_nstore keeps
 *       the 16 real columns of the 128 the tensor core computes),   y = fqss_ola_fwd(frames);
 *     its gradients use the framed output gradient as a three-term bf16 operand (fqss_frames_split: rows 0..L-1 hi, L..2L-1
 *     mid, 2L..3L-1 lo -- together the fp32 value --, zero up to 128; rowsum[3L] = fp64 sums of those rows; 3L <= 128): dgrad =
 *     fqss_pw_gemm(frames_split, [c | c | c | 0]) and wgrad =
 *     fqss_wgrad_codes(frames_split, Y codes) folded by fqss_dec_wgrad_fold (hi + mid + lo rows, transposed to [F][L]).
 *     fqss_frames_encode frames a signal on an 8-bit grid into integer codes [R][KP][ld] (encoder-side operand, KP >= C*L rows);
 *     fqss_sub_fq_codes is the RQB's FQ(Y - Yq) (qat_layers.py:1195) producing the codes the second decode consumes.
 * ------------------------------------------------------------------------------------------- */
int fqss_pw_gemm_nstore(const void* act_bf16, const void* w_bf16, const float* s1, const float* s0, float* out_f32, int n_store,
                        int B, int K, int N, int M, int64_t ld, void* stream);
int fqss_frames_split(const float* g, int64_t ldg, void* out_bf16, int64_t ldo, int64_t R, int M, int L, int H,
                      int zero_rows /* 0: only rows < 3L are written (the caller keeps the others zero) */, double* rowsum,
                      void* stream);
int fqss_frames_encode(const float* x, int64_t ldx, void* out_bf16, int64_t ldo, int64_t R, int C, int M, int L, int H, int KP,
                       const float* rmin, const float* rmax, void* stream);
int fqss_sub_fq_codes(const float* a, int64_t lda, const float* b, int64_t ldb, void* out_bf16, int64_t ldo, int64_t rows, int M,
                      const float* rmin, const float* rmax, void* stream);
int fqss_dec_wgrad_fold(const float* part, float* dWq, int F, int L, void* stream);
/* operand preparation (one tiny launch each).  Decoder weight W [F][L] with its per-tensor quantiser and the quantiser
 * {amin, amax} of the incoming codes: Wc [128][F] bf16 (row k < L = code[.,k]), WT [F][128] bf16 = [code | code | 0], s1/s0
 * [128] (forward affine; zero beyond L), dgs [F] = weight step (dgrad scale).  Encoder-type weight W [N][Kr] (Kr = C*L) with
 * per-output-channel ranges: Wc [N][KP] bf16 codes (zero beyond Kr), s1/s0 [N]. */
int fqss_edge_dec_prep(const float* W, const float* wmin, const float* wmax, const float* amin, const float* amax, void* Wc,
                       void* WT, float* s1, float* s0, float* dgs, int F, int L, void* stream);
int fqss_edge_enc_prep(const float* W, const float* wmin, const float* wmax, const float* amin, const float* amax, void* Wc,
                       float* s1, float* s0, int N, int Kr, int KP, void* stream);

/* ---------------------------------------------------------------------------------------------
 * R1  recurrence of the quantised LSTM (LSTMQ, qat_layers.py:571-613: nn.LSTM evaluated with fake-quantised weights, zero
 *     initial state; DPTNetQ's improved transformer layer, models/dptnetq.py:57-97).  One layer, D = 1 or 2 directions
 *     (direction 1 runs the sequence backwards), hidden size H in {32, 64, 128}, gate order i, f, g, o as torch.
 *       gx    [D][T][N][4H]  input projections x W_ih^T + b_ih + b_hh of every step (a batched GEMM outside)
 *       whh*  [4H][H] raw recurrent weights of direction 0 / 1, wmin* / wmax* [4H] the ranges of their per-row 8-bit
 *             quantisers: the kernels derive the integer codes and steps themselves (same arithmetic as fqss_fq_weight_fwd)
 *             and keep the codes register-resident for all T steps
 *       out   [T][N][D*H]    hidden states (direction d in columns d*H .. d*H+H-1)
 *       gates [D][T][N][4H]  activated gates, cseq [D][T][N][H] cell states (saved for backward)
 *     bwd: dout [T][N][D*H] -> dG [D][T][N][4H], the gradient of gx (= of the pre-activations); the weight gradient is
 *          dG^T h_{t-1} summed over steps, a batched GEMM outside.
 * ------------------------------------------------------------------------------------------- */
int fqss_lstm_rec_fwd(const float* gx, const float* whh0, const float* whh1, const float* wmin0, const float* wmin1,
                      const float* wmax0, const float* wmax1, float* out, float* gates, float* cseq, int T, int N, int H, int D,
                      void* stream);
int fqss_lstm_rec_bwd(const float* dout, const float* gates, const float* cseq, const float* whh0, const float* whh1,
                      const float* wmin0, const float* wmin1, const float* wmax0, const float* wmax1, float* dG, int T, int N,
                      int H, int D, void* stream);

/* ---------------------------------------------------------------------------------------------
 * R2  attention core of MultiheadAttentionQ (qat_layers.py:926-939: between the quantiser of q / sqrt(d) and the quantiser of
 *     the head outputs):  o = softmax(q k^T) v  per (batch, head) for small heads (HD in {8, 16, 32}), scores never
 *     materialised.  q, o, dO, dq [BH][Lq][HD]; k, v, dk, dv [BH][Lk][HD]; lse [BH][Lq] (row log-sum-exp, saved by fwd);
 *     delta_ws [BH][Lq] scratch.  One head's K, V (or Q, dO) must fit shared memory: fqss_attn_smem_bytes <= 200 KB.
 * ------------------------------------------------------------------------------------------- */
size_t fqss_attn_smem_bytes(int Lq, int Lk, int HD);
int fqss_attn_fwd(const float* q, const float* k, const float* v, float* o, float* lse, int BH, int Lq, int Lk, int HD, void* stream);
int fqss_attn_bwd(const float* q, const float* k, const float* v, const float* o, const float* dO, const float* lse, float* dq,
                  float* dk, float* dv, float* delta_ws, int BH, int Lq, int Lk, int HD, void* stream);

/* ---------------------------------------------------------------------------------------------
 * X1  export-time quantisers (qat_quant.py:15-72: TorchWeightFakeQuantize, TorchActivationFakeQuantize,
 *     TorchDymActivationFakeQuantize; installed by qat_utils.py:334-349 replace_*_quantizer).  The reference evaluates
 *     torch.fake_quantize_per_tensor_affine / _per_channel_affine on (scale, zero-point) pairs derived from the learned
 *     ranges; these entry points restate that arithmetic:
 *         inv = 1.0f / scale;  q = nearbyint(x * inv) + zero_point;  y = (clamp(q, qmin, qmax) - zero_point) * scale
 *     mask (optional, 1 byte/elem) = qmin <= q <= qmax (the op's straight-through mask); code (optional, int32) = the
 *     clamped q, i.e. the integer a deployment toolchain stores.  A zero_point outside [qmin, qmax] is refused with
 *     ATen's own message (the reference's export of a range that excludes zero fails the same way).
 *     channel form: x viewed as [outer][ch][inner], scales[ch], zero-points 0.
 * ------------------------------------------------------------------------------------------- */
int fqss_fq_affine_tensor(const float* x, float* y, uint8_t* mask, int32_t* code, int64_t n, float scale, int zero_point,
                          int qmin, int qmax, void* stream);
int fqss_fq_affine_channel(const float* x, float* y, uint8_t* mask, int32_t* code, int outer, int ch, int inner,
                           const float* scales, int qmin, int qmax, void* stream);
int fqss_fq_affine_bwd(const float* g, const uint8_t* mask, float* gx, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * D1  flat gradient arena helpers for the data-parallel exchange (asteroid_librimix_trainer.py:125-135)
 *   sumsq[0] = sum g^2 (for the global-norm clip, gradient_clip_val=5.0)
 *   scale_clip: g *= pre_scale * min(1, max_norm / (sqrt(sumsq*pre_scale^2) + 1e-6))
 * ------------------------------------------------------------------------------------------- */
/* gather the per-parameter gradient tensors into the flat arena (the buffer the all-reduce runs on -- the role of DDP's
 * gradient bucket in the reference, pl.Trainer(strategy="ddp"), asteroid_librimix_trainer.py:125-135): item i copies
 * numel floats from src (NULL: zeros -- a parameter that received no gradient) to dst + offset.  `items` is a HOST array. */
typedef struct fqss_gather_item { const float* src; int64_t offset; int64_t numel; } fqss_gather_item;
int fqss_arena_gather(const fqss_gather_item* items, int n, float* dst, void* stream);
int fqss_arena_sumsq(const float* g, int64_t n, float* sumsq, void* ws, size_t ws_bytes, void* stream);
int fqss_arena_scale_clip(float* g, int64_t n, const float* sumsq, float pre_scale, float max_norm, void* stream);
/* fused Adam step over the flat arena (torch.optim.Adam semantics, weight_decay=0, amsgrad=False) */
int fqss_arena_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                    float eps, int step, void* stream);
/* same step with the step count kept on the device: uses t = *step_dev + 1 for the bias corrections, then increments
 * *step_dev -- nothing step-dependent is baked into the launch, so a captured CUDA graph of the step can be replayed */
int fqss_arena_adam_dev(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                        float eps, int* step_dev, const float* lr_dev /* may be NULL; else the learning rate is read from
                        the device at run time (schedulers such as half_lr keep working under a captured graph) */,
                        void* stream);

/* ---------------------------------------------------------------------------------------------
 * Measurement support (nothing like it in the reference): launch accounting and per-kernel-class
 * CUDA-event timing on the launching stream, used by bench.py's `roofline` / `gpu_launches`.
 *   fqss_launch_count: kernels launched by this library in the process so far.
 *   fqss_prof_enable(1): start bracketing every launch with events; (0): stop and resolve.
 *   fqss_prof_read: per class name, summed milliseconds, timed scopes and kernels (synchronises the device).
 * ------------------------------------------------------------------------------------------- */
int64_t fqss_launch_count(void);
int fqss_prof_enable(int on);
int fqss_prof_reset(void);
int fqss_prof_nslots(void);
int fqss_prof_read(int slot, char* name, int name_cap, double* total_ms, int64_t* scopes, int64_t* kernels);

#ifdef __cplusplus
}
#endif
#endif /* FQSS_H_ */
