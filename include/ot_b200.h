/*
 * ot_b200.h -- C ABI of libot_b200.so: the sm_100a kernels behind the per-op handlers of the
 * custom ONNX executor of gebegebegebe/onnx-transformer.
 *
 * The reference has no native layer: every op is a one-node onnxruntime session built in
 *   onnx_optimized_inference.py:18-57   (execute_node)
 * and the fault hooks are numpy code in
 *   onnx_optimized_inference.py:59-204  and  inject_utils/layers.py:7-142.
 * Each entry point below replaces one op chain of that executor (SURVEY.md section 8a rows are cited per
 * function).  All pointers are DEVICE pointers borrowed from the caller (torch CUDA tensors in the Python
 * host); nothing is allocated, no global state is kept except a TMA-descriptor cache keyed by
 * (pointer, shape); `stream` is a cudaStream_t passed as void*.  Return value: 0 on success, a negative
 * OT_E* code otherwise.  There is no CPU fallback: without a CUDA device every compute entry point
 * returns OT_ENODEV.
 */
#ifndef OT_B200_H_
#define OT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OT_OK 0
#define OT_EINVAL (-1)  /* bad shape / alignment / argument */
#define OT_ECUDA (-2)   /* CUDA runtime or driver error (see ot_last_error) */
#define OT_ENODEV (-3)  /* no sm_100 device */

/* Fault descriptor: the reference's inject_parameters (parallelized_inject_onnx_transformer.py:837-858)
 * with its np.random draws (inject_utils/layers.py:73, onnx_optimized_inference.py:63,118,160,167,
 * layers.py:21,27) made explicit so that a trial is replayable. */
enum OtFaultMode {
  OT_FAULT_NONE = 0,
  OT_FAULT_INPUT = 1,          /* INPUT / INPUT16: bit flip in the int8 A operand (Round_k_out0)      */
  OT_FAULT_WEIGHT = 2,         /* WEIGHT / WEIGHT16: bit flip in the int8 B operand                   */
  OT_FAULT_RANDOM_BITFLIP = 3, /* flip bit `bit` (0 = LSB .. 31) of one fp32 output element, NaN -> 0 */
  OT_FAULT_RANDOM = 4,         /* replace one fp32 output element by `value_bits`, NaN -> 0           */
  OT_FAULT_ACC_BITFLIP = 5,    /* flip bit `bit` of one int32 accumulator (north-star epilogue hook)  */
  OT_FAULT_OUT_Q8_BITFLIP = 6  /* flip bit `bit` (0..7, wrap as flip_int8_bit, inject_utils/layers.py:61-68) of one
                                  requantized int8 output element (OT_OUT_Q8 only; north-star epilogue hook) */
};

typedef struct OtFault {
  int32_t mode;         /* enum OtFaultMode */
  int32_t bit;          /* 0..7 for operand faults (0..3 for packed int4 weights: flip_int4_bit, layers.py:48-59),
                           0..31 for output / accumulator faults */
  int64_t flat_index;   /* row-major flat index into the faulty tensor: A [M,K], B [N,K] or out [M,N] */
  int32_t window_start; /* INPUT16: first affected output column; WEIGHT16: first affected output row */
  int32_t window_len;   /* number of affected columns / rows; <= 0 means the whole row / column      */
  uint32_t value_bits;  /* OT_FAULT_RANDOM: replacement fp32 bit pattern                              */
  int32_t reserved;
} OtFault;

enum OtGemmOut {
  OT_OUT_I32 = 0, /* raw int32 accumulators (ONNX MatMulInteger)                                        */
  OT_OUT_F32 = 1, /* fp32: fl(fl(float(acc)*row_scale[m])*col_scale[n]) + bias[n], ReLU, + residual      */
  OT_OUT_Q8 = 2,  /* OT_OUT_F32 followed by the per-row abs-max requant over groups of `quant_group` cols */
  OT_OUT_QLINEAR = 3 /* ONNX QLinearMatMul: int8 saturate(rint(fl(fl(float(acc')*a_scale)*b_scale) / y_scale) + y_zp)
                        (internal to ot_qlinear_matmul) */
};

/* ---- library / device ---------------------------------------------------------------------------- */
int ot_version(void);
const char* ot_last_error(void);
int ot_device_ok(void); /* 1 when the current device is sm_100 */
/* Programmatic dependent launch: when enabled (default off) every kernel of this library is launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization and waits (griddepcontrol.wait) before touching upstream data, so its
 * launch latency and prologue overlap the predecessor's tail.  Process-wide switch; CUDA-graph capturable. */
int ot_set_pdl(int enable);
/* Profiling aid: device timeline buffer (u64[1 + 4*capacity], zeroed by the caller; NULL disables).  Block 0 of every
 * kernel appends {kernel id, t_start, t_after_dependency_wait, t_end} in %globaltimer nanoseconds. */
int ot_set_timeline(unsigned long long* buf, unsigned int capacity);
/* Number of kernels launched by this library in this process (bench.py's gpu_launches counter). */
int64_t ot_launch_count(void);

/* ---- a8+a9+a10 (+ReLU, +residual, +fault): quantized linear / MatMulInteger ----------------------
 * Replaces the chain  Round -> Mul(scale) -> [Transpose] -> MatMul -> Add(bias) [-> Relu] [-> Abs ..
 * Round]  of quant_linear.py:111-119 as exported (SURVEY.md App. A "Linear").
 * A: int8 [M,K] row-major (lda bytes per row), W: int8 [N,K] row-major (ldw) -- both K-major.
 * tcgen05.mma kind::i8, operands staged by TMA, int32 accumulators in TMEM.
 * out_kind OT_OUT_I32: out = int32 [M,N] (ldo elements per row).
 * out_kind OT_OUT_F32: out = fp32  [M,N]; row_scale [M] / col_scale [N] / bias [N] / residual [M,N]
 *                      (ldr) may each be NULL (treated as 1 / 1 / 0 / 0).
 * out_kind OT_OUT_Q8 : out = int8 [M,N]; out_scale fp32 [M, N/quant_group]; each group of
 *                      quant_group columns is quantized as s = max(max|y|,1e-5)/127, q = rint(y/s).
 * fault may be NULL. */
int ot_linear_w8a8(const int8_t* A, int64_t lda, const int8_t* W, int64_t ldw, int M, int N, int K,
                   const float* row_scale, const float* col_scale, const float* bias,
                   const float* residual, int64_t ldr, int relu,
                   int out_kind, void* out, int64_t ldo, float* out_scale, int quant_group,
                   const OtFault* fault, void* stream);

/* Batched fault trials (one fault per batch unit): unit u owns output rows [u*rows_per_unit, (u+1)*rows_per_unit);
 * unit_fault_dev[u] (DEVICE int32) is the index of its fault in faults_dev (DEVICE array) or -1; every fault's flat_index /
 * window is relative to its own unit, i.e. addresses the one-sentence tensors the reference's trials see
 * (parallelized_inject_onnx_transformer.py runs batch 1).  A WEIGHT fault perturbs the rows of its unit only. */
int ot_linear_w8a8_mf(const int8_t* A, int64_t lda, const int8_t* W, int64_t ldw, int M, int N, int K,
                      const float* row_scale, const float* col_scale, const float* bias,
                      const float* residual, int64_t ldr, int relu,
                      int out_kind, void* out, int64_t ldo, float* out_scale, int quant_group,
                      const OtFault* faults_dev, const int32_t* unit_fault_dev, int rows_per_unit, void* stream);

/* Same contract, W holds int4 values packed two per byte (low nibble = even k), [N, K/2] bytes:
 * Brevitas 4-bit weights are unpacked to int8 in shared memory ahead of the MMA (config #4). */
int ot_linear_w4a8(const int8_t* A, int64_t lda, const uint8_t* W4, int64_t ldw, int M, int N, int K,
                   const float* row_scale, const float* col_scale, const float* bias,
                   const float* residual, int64_t ldr, int relu,
                   int out_kind, void* out, int64_t ldo, float* out_scale, int quant_group,
                   const OtFault* fault, void* stream);

/* ot_linear_w8a8_mf for packed int4 weights (cfg4 fault trials; WEIGHT faults flip one of the 4 bits: flip_int4_bit,
 * inject_utils/layers.py:48-59). */
int ot_linear_w4a8_mf(const int8_t* A, int64_t lda, const uint8_t* W4, int64_t ldw, int M, int N, int K,
                      const float* row_scale, const float* col_scale, const float* bias,
                      const float* residual, int64_t ldr, int relu,
                      int out_kind, void* out, int64_t ldo, float* out_scale, int quant_group,
                      const OtFault* faults_dev, const int32_t* unit_fault_dev, int rows_per_unit, void* stream);

/* ONNX-spec MatMulInteger (the dialect-B / QCDQ graphs of inject_operations.py):
 *   out[m,n] = sum_k (A[m,k] - a_zp[m]) * (W[n,k] - b_zp[n])          int32, wrap-around
 * computed as the int8 tensor-core GEMM with the zero-point correction applied to the accumulator in the epilogue:
 *   acc - a_zp[m]*b_colsum[n] - b_zp[n]*a_rowsum[m] + K*a_zp[m]*b_zp[n].
 * W is K-major [N,K] (the transposed ONNX B operand).  a_zp [M] / b_zp [N] are int32 DEVICE arrays or NULL (= 0);
 * a_rowsum [M] (required with b_zp) and b_colsum [N] (required with a_zp) come from ot_rowsum_i8.  uint8 operands are passed as
 * int8 (x ^ 0x80) with their zero point lowered by 128.  Parity: ONNX operator spec only (no reference artefact). */
int ot_matmul_integer(const int8_t* A, int64_t lda, const int8_t* W, int64_t ldw, int M, int N, int K,
                      const int32_t* a_zp, const int32_t* b_zp, const int32_t* a_rowsum, const int32_t* b_colsum,
                      int32_t* out, int64_t ldo, const OtFault* fault, void* stream);

/* ONNX-spec QLinearMatMul: out int8 [M,N] = saturate(rint(fl(fl(float(acc') * a_scale[m]) * b_scale[n]) / y_scale) + y_zp), acc' as in
 * ot_matmul_integer; a_scale [M] / b_scale [N] fp32 DEVICE arrays (NULL = 1). */
int ot_qlinear_matmul(const int8_t* A, int64_t lda, const int8_t* W, int64_t ldw, int M, int N, int K,
                      const float* a_scale, const float* b_scale, const int32_t* a_zp, const int32_t* b_zp,
                      const int32_t* a_rowsum, const int32_t* b_colsum, float y_scale, int y_zp,
                      int8_t* out, int64_t ldo, void* stream);

/* Row sums of an int8 matrix [rows, cols] (pitch ld) -> int32 [rows]: the operand sums of the zero-point correction. */
int ot_rowsum_i8(const int8_t* X, int64_t ld, int64_t rows, int cols, int32_t* out, void* stream);

/* a6+a7+a8..a10 in one kernel: A = RowQuant(LayerNorm(x)) is produced in shared memory by the GEMM's own warps
 * (layer_norm.py:12-15 + quant_linear.py:31-43, op for op as ot_layernorm_quant) and fed to the MMA without a round trip
 * through HBM; the rest as ot_linear_w8a8 (no fault hook: a fault trial uses the two separate entry points).  K must be 512. */
int ot_ln_linear_w8a8(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps,
                      const int8_t* W, int64_t ldw, int M, int N, int K,
                      const float* col_scale, const float* bias, const float* residual, int64_t ldr, int relu,
                      int out_kind, void* out, int64_t ldo, float* out_scale, int quant_group, void* stream);

/* Standalone nibble unpack: W4 [rows, cols/2] -> int8 [rows, cols] (sign-extended). */
int ot_unpack_int4(const uint8_t* W4, int8_t* W8, int64_t rows, int64_t cols, void* stream);

/* Inverse of ot_unpack_int4: int8 values in [-8, 7] -> packed nibbles (low nibble = even k). */
int ot_pack_int4(const int8_t* W8, uint8_t* W4, int64_t rows, int64_t cols, void* stream);

/* ---- a6+a7: LayerNorm (+ per-token quant) ---------------------------------------------------------
 * layer_norm.py:12-15: a*(x-mean)/(sqrt(sum((x-mean)^2)/(n-1)) + eps) + b, then the RowQuant of
 * quant_linear.py:31-43 when q_out != NULL.  x fp32 [rows, n]; y_out (fp32, optional) [rows,n];
 * q_out int8 [rows,n] + s_out fp32 [rows] (optional pair). */
int ot_layernorm_quant(const float* x, const float* gamma, const float* beta, int64_t rows, int n, float eps,
                       float* y_out, int8_t* q_out, float* s_out, void* stream);

/* ---- a7 / a10: per-row abs-max quantization over groups of `group` columns ------------------------
 * x fp32 [rows, n] (ldx), q int8 [rows, n], s fp32 [rows, n/group]; optional xhat fp32 = q*s. */
int ot_rowquant(const float* x, int64_t ldx, int64_t rows, int n, int group,
                int8_t* q, float* s, float* xhat, void* stream);

/* ---- a17/a18 residual Add -------------------------------------------------------------------------*/
int ot_residual_add(const float* a, const float* b, float* out, int64_t n, void* stream);

/* ---- a20: embedding * sqrt(d) + positional encoding (embeddings.py:12-13, positional_encodings.py:14-25)
 * Row r reads token ids[r*ids_stride] and sits at position pos0 + r % seq_len of its sentence (rows laid
 * out [B, seq_len]).  When pos_dev != NULL (CUDA-graph replay of the greedy loop) the step t = *pos_dev is
 * read from device memory: token = ids[r*ids_stride + t] (ids = the ys buffer [B, max_len]), position = t +
 * r % seq_len. */
int ot_embed_pe(const int64_t* ids, int64_t ids_stride, const float* table, const float* pe, int64_t rows,
                int seq_len, int d, int pos0, const int32_t* pos_dev, float scale, float* out, void* stream);

/* ---- a11..a17: fused quantized attention ------------------------------------------------------------
 * attention.py:23-36 on int8 Q/K/V with per-token scales (SURVEY.md App. A "Attention"):
 *   S = fl(fl(float(sum_d q*k) * sq[i]) * sk[j]) / 8 ; mask -> -1e9 ; P = softmax ; pq = rint(127 P) ;
 *   C[i,:] = sum_j (pq/127) * (sv[j]*v[j,:])   (fp32: the V scale sits on the contraction axis).
 * Query token (b,i) is the int8 row q + (b*Tq + i)*ldq with scale sq[(b*Tq + i)*sq_stride]; key token (b,j)
 * is k/v + (b*Tk_cap + j)*ldk with scales sk/sv[(b*Tk_cap + j)*skv_stride] -- i.e. k/v may be a persistent
 * KV cache of capacity Tk_cap.  When k_new != NULL the Tq newest keys (positions Tk-Tq .. Tk-1) are read from
 * k_new/v_new/sk_new/sv_new (row (b*Tq+i)*ld_new, scale (b*Tq+i)*snew_stride) and appended to the cache.
 * mask_kind 0: none; 1: key padding mask u8 key_mask[b*mask_stride + j] (1 = keep; the reference's
 * global_in_1/global_in_2 [B,1,S]); 2: causal, key j visible iff j <= q_pos0 + i (global_in_3).
 * When step_dev != NULL (CUDA-graph replay of the greedy loop) the step t = *step_dev is read on the device:
 * Tk = t + Tq and q_pos0 = t; the launch is sized for Tk_cap.
 * Outputs (any subset): ctx fp32 [B*Tq, 512] (ld_ctx) = merged heads; ctx_q int8 [B*Tq,512] + ctx_s [B*Tq] =
 * its per-token RowQuant (input of the O projection); probs_q u8 [B,8,Tq,Tk] = rint(127 P), the reference's
 * Round_k_out0 fault target.
 * fault->reserved selects the faulty tensor: 0 q (FirstMatMul INPUT), 1 k (FirstMatMul WEIGHT), 2 P
 * (SecondMatMul INPUT), 3 v (SecondMatMul WEIGHT), 4 scores output, 5 context output (RANDOM / RANDOM_BITFLIP);
 * flat_index addresses that tensor in the reference's layout ([B,T,512] for q/k/v, [B,8,Tq,Tk] / [B,8,Tq,64]
 * for the 4-D tensors); window_* restrict INPUT16 / WEIGHT16 as in onnx_optimized_inference.py:111-179. */
int ot_attention_q8(const int8_t* q, int64_t ldq, const float* sq, int64_t sq_stride,
                    int8_t* k, int8_t* v, int64_t ldk, float* sk, float* sv, int64_t skv_stride,
                    const int8_t* k_new, const int8_t* v_new, int64_t ld_new, const float* sk_new,
                    const float* sv_new, int64_t snew_stride,
                    int B, int H, int Tq, int Tk, int Tk_cap, int mask_kind, const uint8_t* key_mask,
                    int64_t mask_stride, int q_pos0, const int32_t* step_dev,
                    float* ctx, int64_t ld_ctx, int8_t* ctx_q, float* ctx_s, uint8_t* probs_q,
                    const OtFault* fault, void* stream);

/* ot_attention_q8 with batched fault trials: unit = sentence b; unit_fault_dev[b] / faults_dev as for ot_linear_w8a8_mf. */
int ot_attention_q8_mf(const int8_t* q, int64_t ldq, const float* sq, int64_t sq_stride,
                       int8_t* k, int8_t* v, int64_t ldk, float* sk, float* sv, int64_t skv_stride,
                       const int8_t* k_new, const int8_t* v_new, int64_t ld_new, const float* sk_new,
                       const float* sv_new, int64_t snew_stride,
                       int B, int H, int Tq, int Tk, int Tk_cap, int mask_kind, const uint8_t* key_mask,
                       int64_t mask_stride, int q_pos0, const int32_t* step_dev,
                       float* ctx, int64_t ld_ctx, int8_t* ctx_q, float* ctx_s, uint8_t* probs_q,
                       const OtFault* fault, const OtFault* faults_dev, const int32_t* unit_fault_dev, void* stream);

/* ---- a21: generator Linear(512 -> vocab) + log_softmax + arg-max (generator.py:14-15) ---------------
 * h fp32 [rows, d] (ldh), Wg fp32 [vocab, d], bg [vocab].  next_ids int64 [rows] = first arg-max (torch.max);
 * scratch_logits fp32 [rows, vocab] is required workspace (holds the logits on return); optional logp fp32
 * [rows, vocab] (log-probabilities) and margin fp32 [rows] (top1 - top2 logit, the parity filter of the
 * north star: "identical wherever the top-2 margin exceeds the tolerance"). */
int ot_generator_argmax(const float* h, int64_t ldh, const float* Wg, const float* bg, int rows, int d, int vocab,
                        int64_t* next_ids, float* scratch_logits, float* logp, float* margin, void* stream);

/* ---- greedy-loop glue ( parallelized_inject_onnx_transformer.py:753-758 ) ---------------------------
 * ys[b, *step + 1] = next_ids[b]; (*step)++ -- device-side so that one CUDA graph replays every step. */
int ot_append_token(int64_t* ys, int64_t ld_ys, const int64_t* next_ids, int B, int32_t* step_dev, void* stream);

/* ---- persistent greedy decoder: the fault-free greedy loop of parallelized_inject_onnx_transformer.py:616-758
 * (batched as batch_output.py:659-672) as ONE kernel launch for any number of steps.  One CTA per SM stays resident;
 * the ~70 dependent ops of a step are phases separated by a grid-wide barrier (see csrc/ot_decoder.cu).  The plan is an
 * opaque device-resident block (ot_decoder_plan_size bytes, 256-byte aligned) holding the TMA descriptors and every
 * pointer of the decoder; it is built once per (model, workspace) by ot_decoder_plan_build:
 *   layer_ptrs: n_layers x 28 device pointers, per layer
 *       ln1_g ln1_b ln2_g ln2_b ln3_g ln3_b | qkv_w qkv_sw qkv_b | o_w o_sw o_b | cq_w cq_sw cq_b | co_w co_sw co_b |
 *       w1_w w1_sw w1_b | w2_w w2_sw w2_b | kc vc skc svc          (int8 weights [N,K], fp32 scales/biases, KV cache)
 *   ws_ptrs: 24 device pointers
 *       x xq sx acc cq cs hq sh rowmax ckv sckv mask fin_g fin_b houtT gen_wt gen_b gen_pv gen_pi tgt_lut pe ys bar trace
 *       (acc: int32 [4*64*2048]; rowmax: u32 [n_layers*64]; houtT: fp32 [512*64]; gen_wt: the generator weight re-laid out
 *        as [ceil(vocab/32)][512][32]; gen_pv/gen_pi: [ceil(vocab/32)*64]; trace: u64 [256] or NULL)
 * B <= 64 sentences, S <= 96 source tokens, cap <= 96 cache positions.  ot_decoder_run executes greedy steps
 * t0 .. t0+n_steps-1: reads ys[:, t0], writes ys[:, t0+1 .. t0+n_steps] and KV-cache positions t0 .. t0+n_steps-1.
 * Results are bit-identical to stepping with the per-op entry points above. */
int ot_decoder_plan_size(void);
int ot_decoder_plan_build(void* plan_dev, int n_layers, int B, int S, int cap, int vocab, int64_t ys_ld,
                          const void* const* layer_ptrs, const void* const* ws_ptrs);
int ot_decoder_run(const void* plan_dev, unsigned int* bar_dev, int t0, int n_steps, void* stream);

/* ---- cluster-resident greedy decoder (csrc/ot_cdecoder.cu): the same fault-free greedy loop, same results bit for bit, but
 * the batch is cut into groups of `spc` <= 8 sentences and every group is decoded by ONE thread-block cluster of 8 CTAs that
 * exchanges rows <-> column slices through distributed shared memory (st.shared::cluster + barrier.cluster) instead of L2 +
 * grid barriers.  Clusters are independent: any B, no co-residency requirement.  The plan is an opaque device-resident block
 * (ot_cdecoder_plan_size bytes, 256-byte aligned):
 *   layer_ptrs: n_layers x 28 device pointers, per layer, in the order of ot_decoder_plan_build
 *   ws_ptrs: 12 device pointers: ckv sckv mask fin_g fin_b gen_w4 gen_b tgt_lut pe ys trace(u64[256] or NULL) gen_w16(or NULL)
 *       (gen_w16, optional: blob for the tensor-core screening generator = 1024-byte header whose first float is max_v ||w_v||_2 of
 *        the generator weight, followed by its fp16 copy as [8 * 576][512] row-major, zero rows past the vocabulary; NULL = exact
 *        fp32 generator only.  Tokens are identical either way: screened entries are re-evaluated with the exact fp32 chain.)
 *       (gen_w4: the generator weight re-laid out as [ceil(vocab/32)][128][32][4] = tile / k-quad / entry / k, zero padded)
 * S <= 96 source tokens, cap <= 96 cache positions, vocab <= 6144.  ot_cdecoder_run executes greedy steps t0 .. t0+n_steps-1:
 * reads ys[:, t0], writes ys[:, t0+1 .. t0+n_steps] and the self-attention KV-cache positions t0 .. t0+n_steps-1. */
int ot_cdecoder_plan_size(void);
int ot_cdecoder_plan_build(void* plan_dev, int n_layers, int B, int S, int cap, int vocab, int spc, int64_t ys_ld,
                           const void* const* layer_ptrs, const void* const* ws_ptrs);
int ot_cdecoder_run(const void* plan_dev, int B, int spc, int t0, int n_steps, void* stream);
/* Number of 8-CTA decoder clusters the current device keeps resident at once (15 on a 148-SM B200).  A greedy step takes the same
 * time whether one or all of them are in use: callers that choose their own batch (fault-trial campaigns) get the most out of a
 * launch at 8 x *n_out sentences; one cluster more and the launch runs in two waves. */
int ot_cdecoder_max_clusters(int* n_out);

/* ---- elementwise / shape op family: one CUDA handler per remaining ONNX op name of SURVEY.md 8a -----
 * Unary  (op: 0 Abs 1 Relu 2 Sqrt 3 Round 4 Neg 5 Exp 6 Identity), fp32, n elements. */
int ot_unary_f32(int op, const float* x, float* y, int64_t n, void* stream);
/* Binary with numpy broadcasting over up to 4 dims (op: 0 Add 1 Sub 2 Mul 3 Div 4 Max 5 Min);
 * shapes given as 4 extents each, right-aligned, stride 0 where an extent is 1. */
int ot_binary_f32(int op, const float* a, const int64_t a_shape[4], const float* b, const int64_t b_shape[4],
                  float* out, const int64_t out_shape[4], void* stream);
/* Clip(x, lo, hi). */
int ot_clip_f32(const float* x, float lo, float hi, float* y, int64_t n, void* stream);
/* ReduceMax / ReduceMean over the last axis (op 0 max, 1 mean): x [rows, n] -> y [rows]. */
int ot_reduce_last_f32(int op, const float* x, int64_t rows, int n, float* y, void* stream);
/* Softmax over the last axis. */
int ot_softmax_f32(const float* x, int64_t rows, int n, float* y, void* stream);
/* Where(cond != 0, a_scalar, x) with cond (u8) broadcast over 4 dims like ot_binary_f32. */
int ot_where_f32(const uint8_t* cond, const int64_t c_shape[4], float a_scalar, const float* x,
                 float* out, const int64_t out_shape[4], void* stream);
/* Equal(x, scalar) on int64 -> u8. */
int ot_equal_i64(const int64_t* x, int64_t scalar, uint8_t* out, int64_t n, void* stream);
/* Cast between element kinds (0 f32, 1 i64, 2 u8 read numerically / written as bool 0|1, 3 i8, 4 i32; destination only: 5 = numeric u8). */
int ot_cast(int src_kind, const void* src, int dst_kind, void* dst, int64_t n, void* stream);
/* 4-D permutation copy of 4-byte elements (Transpose). */
int ot_transpose4_b32(const void* x, const int64_t shape[4], const int perm[4], void* y, void* stream);
/* Batched fp32 MatMul C[b] = A[b] (MxK) * B[b] (KxN), row-major, batch strides in elements (0 = shared).
 * The executor's float MatMul handler (operands that are not int8-backed, e.g. P.V and fault deltas). */
int ot_matmul_f32(const float* A, const float* B, float* C, int batch, int M, int N, int K,
                  int64_t strideA, int64_t strideB, int64_t strideC, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OT_B200_H_ */
