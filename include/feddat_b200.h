/* feddat_b200.h -- C ABI of libfeddat_sm100.so
 *
 * Drop-in boundary for the FedDAT per-client local-training hot path.  The
 * reference (HaokunChen245/FedDAT) is pure Python/PyTorch and has no FFI of its own; each entry
 * point below replaces the PyTorch-eager arithmetic of one reference function, cited per function.
 * The host side that binds these (ctypes) lives in feddat_b200/_lib.py; INTEGRATION.md shows the
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise
 *   - all buffers are owned by the caller; the library allocates nothing persistent
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*), no internal sync,
 *     CUDA-graph capturable
 *   - return 0 on success, a negative FEDDAT_ERR_* otherwise; feddat_last_error() returns a
 *     thread-local message.  Never throws, never exits.
 *   - sm_100a only.  Unsupported shapes are an explicit error: there is NO CPU / Triton fallback.
 */
#ifndef FEDDAT_B200_H_
#define FEDDAT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FEDDAT_OK 0
#define FEDDAT_ERR_INVALID (-1)
#define FEDDAT_ERR_UNSUPPORTED (-2)
#define FEDDAT_ERR_CUDA (-3)
#define FEDDAT_ERR_NO_DEVICE (-4)

#define FEDDAT_DTYPE_BF16 0
#define FEDDAT_DTYPE_F32 1 /* MKD head only: the DAT operator itself is bf16 */

#define FEDDAT_ACT_RELU 0 /* reference default: adapter.py:24 */
#define FEDDAT_ACT_GELU 1 /* opt-in (BASELINE.json north_star wording) */

const char* feddat_last_error(void);
int feddat_abi_version(void);

/* ---------------------------------------------------------------------------------------------
 * DAT bottleneck forward.  Replaces Adapter.forward (src/modeling/models/adapter.py:124-163):
 *   single mode (:125-131)  Y = Res + up(act(down(X)))                        r_total = r, scale 1
 *   gating mode (:133-146)  Y = Res + 1.0 * (0.5 * A0(X) + 0.5 * A2(X))       r_total = 2r, scale .5
 * with the active branches concatenated along the bottleneck dimension:
 *   Wd_cat [r_total, d] bf16 row-major   (= rows of adapter_i_down.weight, stacked)
 *   bd_cat [r_total]    fp32
 *   Wu_cat [d, r_total] bf16 row-major   (= columns of adapter_i_up.weight, side by side)
 *   bu_cat [d]          fp32             (= sum of the active branches' up biases)
 *   Y = Res + branch_scale * ( act(X Wd_cat^T + bd_cat) Wu_cat^T + bu_cat )
 * X, Res, Y: [M, d] bf16 row-major contiguous, 16-byte aligned; Res may alias X; Y must not.
 * d must be 768; r_total a multiple of 16 in [16, 256].
 */
int feddat_dat_fwd(const void* X, const void* Res, void* Y, const void* Wd_cat,
                   const float* bd_cat, const void* Wu_cat, const float* bu_cat,
                   void* H_out /* NULL, or [M, r_total] bf16: the hidden act(X Wd_cat^T + bd_cat), saved for a
                                  backward that does not recompute it (what torch autograd would save) */,
                   int64_t M, int d, int r_total, float branch_scale, int act, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * DAT bottleneck backward (autograd of the above; the reference relies on torch autograd of
 * adapter.py:124-163 with trainability toggled at adapter.py:71-85).
 *
 * Stage 1, feddat_dat_bwd_dgrad: recomputes the hidden, then
 *     dP = (dY Wu_cat) * act'(X Wd_cat^T + bd_cat) * branch_scale
 *     dX = dP Wd_cat  (+ dY when add_dy != 0, i.e. the residual input IS X: adaptered_output.py:78)
 *   and, for the trainable slice [r_lo, r_hi) of the bottleneck, writes the bf16 hidden
 *   H_t [M, r_hi-r_lo] and dP_t [M, r_hi-r_lo] for stage 2 (pass NULL/NULL when nothing trains).
 *   dX may be NULL (first adapter site: nothing trainable upstream needs it) only if H_t != NULL.
 *   WuT_cat [r_total, d] and WdT_cat [d, r_total] are the transposes of Wu_cat / Wd_cat (bf16).
 *   Two modes.  RECOMPUTE (H_in == NULL; any activation): as above, H_t and dP_t written with row stride
 *   ld_t.  SAVED (H_in = the forward's H_out, ReLU only): P is not recomputed -- X, Wd_cat, bd_cat are
 *   unused (may be NULL), relu'(P) = (H_in > 0), H_t must be NULL (stage 2 reads the slice of H_in
 *   itself), dP_t [.., ld_t] is written for the trainable slice.
 *
 * Stage 2, feddat_dat_bwd_wgrad: fp32, OVERWRITES its outputs (no zero-initialisation needed), bit-wise
 * reproducible (fixed summation order across the row splits):
 *     dWu [d, r_t] = branch_scale * dY^T H_t                        dbu [d]   = scale * sum_m dY
 *     dWd [r_t, d] = dP_t^T X                                       dbd [r_t] = sum_m dP_t
 *   (dP_t already carries branch_scale.)  r_t must be a multiple of 16, <= 128; wider trainable
 *   slices are covered by one call per 128 columns: ld_ht is the row stride (elements) of H_t/dP_t,
 *   ld_dwu the row stride of dWu, so a call can address a column slice of wider arrays.  dbu / dbd
 *   may be NULL (skipped).  `workspace`: feddat_dat_wgrad_workspace_bytes() bytes of device memory,
 *   16-byte aligned, zero-initialised ONCE by the caller (the kernel leaves its counters at zero), not
 *   shared by launches that may run concurrently.
 */
int feddat_dat_bwd_dgrad(const void* X, const void* dY, void* dX, const void* Wd_cat,
                         const float* bd_cat, const void* WuT_cat, const void* WdT_cat,
                         const void* H_in, void* H_t, void* dP_t, int ld_t, int r_lo, int r_hi,
                         int64_t M, int d, int r_total, float branch_scale, int act, int add_dy,
                         int dtype, void* stream);

int feddat_dat_bwd_wgrad(const void* X, const void* dY, const void* H_t, const void* dP_t,
                         float* dWu, float* dbu, float* dWd, float* dbd, int64_t M, int d, int r_t,
                         int ld_ht, int ld_dwu, float branch_scale, int dtype, void* workspace,
                         size_t ws_bytes, void* stream);
size_t feddat_dat_wgrad_workspace_bytes(void);

/* ---------------------------------------------------------------------------------------------
 * Grouped launches: up to two independent row groups -- each with its own activations, packed weights,
 * bottleneck width and scale -- in ONE kernel launch.  This is the MKD schedule of
 * TaskTrainer.train_step (task_trainer.py:280-330) at one adapter site when passes A/C (gating pair,
 * r_total = 2r, scale .5) and B (adapter_1 alone, r_total = r, scale 1) are row-stacked: one launch per
 * site and direction instead of two, filling the machine without recomputing the hidden per column split.
 * A group uses the fields of the single-group call of the same direction (forward: X, Res, Y, Wd_cat,
 * bd_cat, Wu_cat, bu_cat, H_out; backward dgrad: X, dY, dX, Wd_cat, bd_cat, WuT_cat, WdT_cat, H_in, H_t,
 * dP_t, ld_t, r_lo, r_hi, add_dy) with the same meaning and constraints; unused fields are NULL / 0.  Both
 * groups of a backward launch must be in the same mode (saved / recompute).  Groups too large for one wave
 * of CTAs (more than one 128-row tile per SM in total) are launched one by one by the library.
 * `groups` is a HOST array.
 */
typedef struct FeddatDatGroup {
  const void* X;
  const void* Res;      /* forward: residual input */
  void* Y;              /* forward: output */
  const void* dY;       /* backward */
  void* dX;             /* backward (nullable) */
  const void* Wd_cat;
  const float* bd_cat;
  const void* Wu_cat;
  const float* bu_cat;
  const void* WuT_cat;
  const void* WdT_cat;
  void* H_out;          /* forward: hidden to save (nullable) */
  const void* H_in;     /* backward, saved mode */
  void* H_t;
  void* dP_t;
  int ld_t, r_lo, r_hi;
  int64_t M;
  int r_total;
  float branch_scale;
  int add_dy;
} FeddatDatGroup;
int feddat_dat_fwd_grouped(const FeddatDatGroup* groups, int n_groups, int d, int act, int dtype, void* stream);
int feddat_dat_bwd_dgrad_grouped(const FeddatDatGroup* groups, int n_groups, int d, int act, int dtype,
                                 void* stream);
/* Weight gradients of 1 .. 24 groups in one launch (same fields as feddat_dat_bwd_wgrad).  The launch has
 * groups x 6 column chunks x row splits CTAs, never more than one per SM: two groups = one adapter site of the batched
 * MKD schedule with 12 row splits and the two-stage reduction; 24 groups = BOTH row groups of ALL twelve ViLT sites
 * at the end of a backward pass (the deferred form, task_trainer.py:311-328 needs the gradients only at
 * optimizer.step()): one CTA per (group, chunk) contracts over all rows, no splits, no workspace traffic. */
typedef struct FeddatWgradGroup {
  const void *X, *dY, *H_t, *dP_t;
  float *dWu, *dbu, *dWd, *dbd;
  int64_t M;
  int r_t, ld_ht, ld_dwu;
  float branch_scale;
} FeddatWgradGroup;
int feddat_dat_bwd_wgrad_grouped(const FeddatWgradGroup* groups, int n_groups, int d, int dtype,
                                 void* workspace, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Packing of the fp32 master weights of the active branches into the bf16 operands above.
 *   down_w[i] : [r, d] fp32, down_b[i] : [r], up_w[i] : [d, r], up_b[i] : [d]   (nn.Linear layout,
 *   adapter.py:35,41), n_branch in {1, 2}.  Outputs: Wd_cat [n*r, d], WdT_cat [d, n*r],
 *   Wu_cat [d, n*r], WuT_cat [n*r, d] (bf16), bd_cat [n*r], bu_cat [d] (fp32).  Any output may be
 *   NULL to skip it.
 */
int feddat_pack_weights(const float* const* down_w, const float* const* down_b,
                        const float* const* up_w, const float* const* up_b, int n_branch, int r,
                        int d, void* Wd_cat, void* WdT_cat, void* Wu_cat, void* WuT_cat,
                        float* bd_cat, float* bu_cat, void* stream);

/* The same packing for several operand sets ("jobs") in ONE launch: a train step packs every adapter site
 * in both of its modes (gating pair; adapter_1 alone) once, when the weights change, instead of once per
 * forward.  `jobs` is a HOST array (copied into the kernel parameters; more than 24 jobs run as several
 * launches).  Per job: up_w[b] may point at a column slice of a wider [d, ld_up] matrix (ld_up = its row
 * stride in elements, 0 means r) and down_w[b] at a row slice, so a bottleneck wider than one launch
 * covers is packed segment by segment without copies; bu_cat = bu_src[0] + bu_src[1] (either may be NULL:
 * the up biases ride on ONE segment only).
 */
typedef struct FeddatPackJob {
  const float* down_w[2];
  const float* down_b[2];
  const float* up_w[2];
  const float* bu_src[2];
  int n_branch, r, ld_up;
  void *Wd_cat, *WdT_cat, *Wu_cat, *WuT_cat;
  float *bd_cat, *bu_cat;
} FeddatPackJob;
int feddat_pack_weights_batched(const FeddatPackJob* jobs, int n_jobs, int d, void* stream);

/* ---------------------------------------------------------------------------------------------
 * MKD head.  Replaces kl_loss (src/train/visionlanguage_tasks/task_trainer.py:506-516) plus the
 * ViLT task loss BCEWithLogits('mean') * C (task_trainer.py:299,319; train_vqa_crossvqa.py:237)
 * and the (task + kl) / 2 combination (task_trainer.py:300-301,320-321), forward and backward in
 * one pass:
 *     kl   = T^2 / batchmean_div * sum_rows KL( softmax(teacher/T) || softmax(logits/T) )
 *     task = task_scale * sum_{rows,c} bce_with_logits(logits, target)    (target may be NULL)
 *     loss_out[0] = kl_weight * kl + task_weight * task, loss_out[1] = kl, loss_out[2] = task
 *     dlogits = d loss_out[0] / d logits          (dlogits may be NULL: forward only)
 * logits/teacher/target/dlogits: [rows, C] fp32 row-major.  loss_out: 3 floats, device memory,
 * overwritten.  For the reference's ViLT path: batchmean_div = rows, task_scale = 1/rows
 * (mean over rows*C, times C), kl_weight = task_weight = 0.5.
 * row_ws: [rows, 2] floats of device scratch (per-row kl / task sums; the final sum over rows runs in a fixed
 * order, so loss_out is bit-wise reproducible).
 */
int feddat_mkd_loss(const float* logits, const float* teacher, const float* target,
                    float* loss_out, float* dlogits, int64_t rows, int C, float temp,
                    float kl_weight, float task_weight, float task_scale, int64_t batchmean_div,
                    float* row_ws, void* stream);

/* MKD head of the ALBEF path: kl_loss (task_trainer.py:506-516, the C > 3000 branch) over the decoder logits
 * plus the answer loss of BertLMHeadModel.forward with reduction='none' (src/modeling/models/xbert.py:1287-1297)
 * weighted as ALBEF.forward does (src/modeling/models/albef_model.py:142-143), value and gradient in one pass:
 *     kl   = T^2 / n_seq * sum_{s, p < La-1} KL( softmax(teacher[s,p]/T) || softmax(logits[s,p]/T) )
 *     task = sum_s seq_weight[s] * sum_{p < La-1, labels[s,p+1] != -100} CE(logits[s,p], labels[s,p+1])
 *     loss_out = {kl_weight * kl + task_weight * task, kl, task};   dlogits = d loss_out[0] / d logits
 * logits / dlogits: the UNSHIFTED prediction scores [n_seq, La, C] (position La-1 gets a zero gradient);
 * teacher: [n_seq, La_teacher, C] with La_teacher = La (unshifted) or La - 1 (the reference's returned
 * logits[:, :-1] copy); dtype FEDDAT_DTYPE_BF16 or FEDDAT_DTYPE_F32 for all three; labels [n_seq, La] int64;
 * seq_weight [n_seq] fp32 = answer weight / image batch size; row_ws: [n_seq * La, 2] floats of scratch.
 */
int feddat_mkd_ce_loss(const void* logits, const void* teacher, const int64_t* labels, const float* seq_weight,
                       float* loss_out, void* dlogits, int64_t n_seq, int La, int La_teacher, int C, float temp,
                       float kl_weight, float task_weight, int dtype, float* row_ws, void* stream);

/* ---------------------------------------------------------------------------------------------
 * FedAvg of the flat communicated buffer.  Replaces get_average_net (src/train/main.py:50-65):
 *     out[i] = sum_c (clients[c][i] * weights[c]) / sum(weights)      in client order, fp32
 * i.e. the reference's `temp += net[key] * num / total` with weights = nums (HOST floats,
 * n_clients <= 64), bit-identical to the PyTorch expression.  Used for clients that share a GPU;
 * across GPUs the flat buffer goes through one NCCL allreduce.  All buffers 16-byte aligned.
 */
int feddat_fedavg(const float* const* clients /* host array of device ptrs */,
                  const float* weights /* host */, int n_clients,
                  float total_weight /* <= 0: sum(weights); > 0: the global total when this call
                                        only sums the clients of one rank before the allreduce */,
                  float* out, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused (residual add +) LayerNorm over d = 768 for the frozen transformer block around each DAT
 * site (HF ViltLayer.forward: layernorm_before / first residual + layernorm_after; the tensors the
 * reference's Adaptered_ViltOutput.forward, src/modeling/adaptered_output.py:73-79, receives).
 *   s = bf16(x + res) (res, sum_out both NULL: s = x), y = bf16((s - mean) rstd w + b); mean, rstd: fp32 [M]
 *   s2 = bf16(s + bias2) (bias2, sum2_out both NULL: skipped): the residual stream with the next dense layer's
 *   bias pre-added (that layer then is one GEMM with a beta = 1 epilogue)
 *   dx = rstd (g - mean_c g - xhat mean_c(g xhat)) (+ dsum),  g = dy w,  xhat = (s - mean) rstd
 * The affine parameters are FROZEN on this path (main.py:138-139): no weight / bias gradients.
 * x, res, y, sum_out, dy, dsum, s, dx: [M, 768] bf16 contiguous; weight, bias: [768] bf16.
 */
int feddat_ln_fwd(const void* x, const void* res, const void* weight, const void* bias, void* y,
                  void* sum_out, const void* bias2, void* sum2_out, float* mean, float* rstd, int64_t M,
                  int d, float eps, int dtype, void* stream);
int feddat_ln_bwd(const void* dy, const void* dsum, const void* s, const void* weight,
                  const float* mean, const float* rstd, void* dx, int64_t M, int d, int dtype,
                  void* stream);

/* Exact (erf) GELU of the frozen ViLT intermediate layer (HF ViltIntermediate), bf16 in / out, n a
 * multiple of 8:  y = 0.5 x (1 + erf(x / sqrt 2));  dx = dy * d/dx of that. */
int feddat_gelu_fwd(const void* x, void* y, int64_t n, int dtype, void* stream);
int feddat_gelu_bwd(const void* dy, const void* x, void* dx, int64_t n, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The frozen MLP of the block around each ViLT DAT site with the exact GELU fused into the GEMM epilogue
 * (HF ViltIntermediate: dense 768 -> 3072 + GELU, feeding the ViltOutput.dense that the reference wraps,
 * src/modeling/adaptered_output.py:67-79):
 *   feddat_mlp_fc1_gelu_fwd    pre = A W^T + bias,  act = gelu(pre)                 (both written, bf16)
 *   feddat_mlp_fc2_dgelu_bwd   dpre = (dY W2T^T) * gelu'(pre)
 * A, dY: [M, K] bf16 row-major; W, W2T: [N, K] bf16 row-major (K contiguous: W is nn.Linear(K, N).weight, W2T the
 * TRANSPOSE of ViltOutput.dense.weight [K, N], made once -- the backbone is frozen); bias [N] FP32 (a widened copy
 * of the frozen bias, made once); pre, act, dpre: [M, N] bf16.  N a multiple of 256, K a multiple of 64, all
 * pointers 16-byte aligned.  Every output is rounded ONCE, from the fp32 accumulator: act = bf16(gelu(acc + b)),
 * dpre = bf16(acc * gelu'(pre)) -- one rounding fewer than the unfused sequence, which rounds the GEMM output first.
 */
int feddat_mlp_fc1_gelu_fwd(const void* A, const void* W, const void* bias, void* pre_out, void* act_out, int64_t M,
                            int N, int K, int dtype, void* stream);
int feddat_mlp_fc2_dgelu_bwd(const void* dY, const void* W2T, const void* pre, void* dpre_out, int64_t M, int N, int K,
                             int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Self-attention of the frozen block around each ViLT DAT site for SHORT sequences (HF ViltSelfAttention, the
 * backbone of src/modeling/vilt.py:19,127; S <= 256 keys, head dimension 64, no mask, no dropout):
 *   feddat_attn_fwd   O = softmax(Q K^T * scale) V,  LSE = log sum exp of the scaled scores (natural log)
 * Q, K, V, O: bf16, element (b, s, h, d) at ((b * S + s) * ld + h * 64 + d) -- the [B * S, H * 64] projection outputs
 * themselves (ld = 768) or column slices of a fused q/k/v projection (ld = 2304); LSE: [B, H, S] fp32.
 * Token strides in elements, multiples of 8; bases 16-byte aligned.
 */
int feddat_attn_fwd(const void* Q, const void* K, const void* V, void* O, void* LSE, int B, int S, int H, int D,
                    int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, float scale, int dtype, void* stream);
/*   feddat_attn_bwd   dQ, dK, dV of the same attention from dO, the forward's inputs, its output O and LSE (S <= 192):
 *                     a row-statistics pass (delta = dO . O) into the caller's workspace, then ONE fused launch.  All nine tensors in the token layout above, each with its own token stride, so the
 *                     three gradients may be column slices of one [B * S, 3 * H * 64] tensor. */
size_t feddat_attn_bwd_workspace_bytes(int B, int H);   /* per-row statistics (delta, scaled logsumexp), fp32 */
int feddat_attn_bwd(const void* dO, const void* Q, const void* K, const void* V, const void* O, const void* LSE, void* dQ,
                    void* dK, void* dV, int B, int S, int H, int D, int64_t lddo, int64_t ldq, int64_t ldk, int64_t ldv,
                    int64_t ldo, int64_t lddq, int64_t lddk, int64_t lddv, float scale, void* workspace, size_t ws_bytes,
                    int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Patch cut + cast for the ViLT patch embedding (HF ViltPatchEmbeddings: Conv2d(C, 768, kernel ps, stride ps)):
 * pixel_values [B, C, H, W] (fp32 or bf16, in_dtype) -> patches [B (H/ps) (W/ps), C ps ps] bf16, row = (b, patch row,
 * patch column), column = (channel, y, x); the embedding is then one GEMM against the weight's [768, C ps ps] view.
 */
int feddat_patchify(const void* pixel_values, void* patches, int B, int C, int H, int W, int ps, int in_dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * AdamW over the tensors of one optimizer step (torch.optim.AdamW as built by the reference's create_optimizer,
 * task_trainer.py:477-504; arithmetic of torch's fused kernel, fp32).  Every tensor carries its own device-side step
 * counter (fp32 scalar, incremented by this call) and a pointer to its group's learning rate on the device, so the
 * call can live inside a CUDA graph; tensors whose gradient is absent are simply not listed.
 */
typedef struct {
  void* param;            /* fp32 [numel], updated in place */
  const void* grad;       /* fp32 [numel] */
  void* exp_avg;          /* fp32 [numel] */
  void* exp_avg_sq;       /* fp32 [numel] */
  float* step;            /* device scalar: steps taken so far */
  const float* lr;        /* device scalar */
  float weight_decay;
  int64_t numel;
} FeddatAdamwTensor;
int feddat_adamw_step(const FeddatAdamwTensor* tensors, int n, float beta1, float beta2, float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FEDDAT_B200_H_ */
