/*
 * univtg_b200 — C ABI of the B200-native UniVTG hot path (cross-modal encoder + heads).
 *
 * Drop-in boundary: the reference reaches this path through ONE Python plugin call,
 *     importlib.import_module('model.' + opt.model_id).build_model(opt) -> (model, criterion)
 *     (reference main/config.py:341-342), then model(**model_inputs) (main/inference_mr.py:101,
 *     main/train_vlp_ddp.py:56) and criterion(outputs, targets) (main/train_vlp_ddp.py:57).
 * The reference has no FFI of its own (it is pure Python on torch); the functions below are what a
 * ctypes binding for that plugin binds (univtg_b200/_lib.py).  Plain pointers and
 * sizes only: no torch types.  All pointers are DEVICE pointers unless stated otherwise, all tensors are
 * row-major/contiguous, `stream` is a cudaStream_t passed as void*.  Every function returns 0 on success,
 * non-zero on failure; univtg_last_error() then describes the failure (thread local).
 * Nothing here allocates device memory: the caller owns `packed` and `workspace`.
 */
#ifndef UNIVTG_B200_H_
#define UNIVTG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UNIVTG_ABI_VERSION 2
#define UNIVTG_ADAMW_SCRATCH_FLOATS 2048

/* Model hyper-parameters: the fields of `args` that reference model/univtg.py:409-450 (build_model),
 * model/transformer_encoder_droppath.py:141-152 (build_transformer) and model/position_encoding.py:113-126 read. */
typedef struct univtg_config {
  int32_t hidden_dim;       /* args.hidden_dim  (d)            */
  int32_t nheads;           /* args.nheads      (H), dh = d/H  */
  int32_t dim_feedforward;  /* args.dim_feedforward            */
  int32_t enc_layers;       /* args.enc_layers                 */
  int32_t n_input_proj;     /* args.n_input_proj in {1,2,3}    */
  int32_t v_feat_dim;       /* args.v_feat_dim (already +2 TEF)*/
  int32_t t_feat_dim;       /* args.t_feat_dim                 */
  int32_t operand_format;   /* 0 = fp16 MMA operands (default), 1 = bf16; accumulation/LN/softmax/heads are fp32 */
} univtg_config;

/* Problem shape of one batch (reference Model.forward arguments, model/univtg.py:105). */
typedef struct univtg_shape {
  int32_t batch;  /* B   */
  int32_t l_vid;  /* L_v */
  int32_t l_txt;  /* L_t */
  int32_t training; /* 1: keep the activations backward needs (larger workspace) */
} univtg_shape;

/* Train-mode randomness generated inside the kernels (Philox4x32-10 keyed by (seed, mask index); csrc/philox.cuh): input
 * dropout of every LinearLayer (reference model/univtg.py:394,401) and DropPath (transformer_encoder_droppath.py:154-167).
 * The same struct passed to univtg_forward_train and univtg_backward reproduces the same draws; explicit multiplier tensors
 * (droppath_scale / drop_masks arguments) take precedence where given. */
typedef struct univtg_rng {
  uint64_t seed;        /* one value per training forward */
  float input_dropout;  /* args.input_dropout (p of nn.Dropout); 0 = off */
  float droppath;       /* args.droppath (drop probability); 0 = off */
} univtg_rng;

typedef struct univtg_plan univtg_plan; /* opaque; host memory only (tensor maps + pointer table) */

const char* univtg_last_error(void);
int univtg_abi_version(void);

/* Number of fp32 parameter tensors univtg_pack_weights expects, in this order (reference state_dict names):
 *   for i < n_input_proj: input_vid_proj.i.{LayerNorm.weight, LayerNorm.bias, net.1.weight, net.1.bias}
 *   for i < n_input_proj: input_txt_proj.i.{...same...}
 *   token_type_embeddings.weight
 *   for l < enc_layers: transformer.encoder.layers.l.{self_attn.in_proj_weight, self_attn.in_proj_bias,
 *        self_attn.out_proj.weight, self_attn.out_proj.bias, linear1.weight, linear1.bias, linear2.weight,
 *        linear2.bias, norm1.weight, norm1.bias, norm2.weight, norm2.bias}
 *   span_embed.layers.{0,1,2}.{weight,bias}; class_embed.layers.{0,1,2}.{weight,bias}; weightedpool.weight */
int univtg_num_params(const univtg_config* cfg);
/* Bytes of the packed-weight buffer (16-bit K-padded GEMM operands + fp32 vectors). */
size_t univtg_packed_bytes(const univtg_config* cfg);
/* Convert/re-layout the fp32 parameters into `packed` (device). `params`: HOST array of device pointers. */
int univtg_pack_weights(const univtg_config* cfg, const float* const* params, int32_t n_params, void* packed, void* stream);

/* Workspace bytes for one (config, shape). */
size_t univtg_workspace_bytes(const univtg_config* cfg, const univtg_shape* shape);
/* Zero the few regions of a workspace that kernels rely on reading as zeros (the separator rows of the conv-head buffers that
 * implement Conv1d's zero padding, reference model/univtg.py:375-377); everything else is written before it is read.  Call it
 * when a workspace buffer is used for the first time or handed over from another shape (workspaces may be pooled and shared
 * between shapes: size them for the largest shape).  training_ws: 0 = workspace of univtg_plan_create (univtg_workspace_bytes),
 * 1 = training workspace (univtg_train_workspace_bytes). */
int univtg_prepare_workspace(const univtg_config* cfg, const univtg_shape* shape, void* workspace, int32_t training_ws, void* stream);
/* Build a plan: tensor maps over `packed` and `workspace` (both must stay alive and must not move).
 * `dim_t`: device fp32 [hidden_dim], the sine-embedding denominators temperature**(2*(j//2)/d)
 * (reference model/position_encoding.py:75) evaluated by the caller.  Calls univtg_prepare_workspace on `stream`. */
int univtg_plan_create(const univtg_config* cfg, const univtg_shape* shape, const void* packed, void* workspace,
                       const float* dim_t, void* stream, univtg_plan** out);
void univtg_plan_destroy(univtg_plan* plan);

/* Element type of the src_txt / src_vid pointers the forward entry points (and univtg_backward) receive for this plan:
 * 0 = f32 (what the reference collate produces, default), 1 = fp16, 2 = bf16 (packed feature shards, univtg_b200/data.py: the CLIP /
 * SlowFast features are stored as 16-bit on disk, so the H2D copy and the first LayerNorm's read halve).  Masks stay f32. */
int univtg_plan_set_input_format(univtg_plan* plan, int32_t fmt);

/* Model.forward (reference model/univtg.py:105-155).
 *   src_txt [B,Lt,Dt] f32, src_txt_mask [B,Lt] f32 (1 = valid), src_vid [B,Lv,Dv] f32, src_vid_mask [B,Lv] f32
 *   droppath_scale: NULL (eval) or [2*enc_layers, B] f32 per-sample residual-branch scales
 *                   floor(keep + u)/keep in the reference's draw order (transformer_encoder_droppath.py:154-167)
 * outputs: pred_logits [B,Lv,1], pred_spans [B,Lv,2], vid_mem_proj [B,Lv,d], txt_mem_proj [B,1,d],
 *          saliency_scores [B,Lv]  (all f32) */
int univtg_forward(univtg_plan* plan, const float* src_txt, const float* src_txt_mask, const float* src_vid,
                   const float* src_vid_mask, const float* droppath_scale, float* pred_logits, float* pred_spans,
                   float* vid_mem_proj, float* txt_mem_proj, float* saliency_scores, void* stream);

/* ---- training (reference main/train_vlp_ddp.py:56-64: model(**inputs); criterion(outputs, targets); losses.backward()) ---- */

/* Bytes of the training workspace (saved activations + backward scratch) for one (config, shape). */
size_t univtg_train_workspace_bytes(const univtg_config* cfg, const univtg_shape* shape);
/* Model.forward in training mode: same outputs as univtg_forward, keeps what backward needs in `train_ws`.

 *   droppath_scale: NULL or [2*enc_layers, B] (see univtg_forward)
 *   drop_masks: NULL or HOST array of 2*n_input_proj device pointers (video layers, then text layers): fp32 [rows, din_i]
 *               input-dropout multipliers (0 or 1/(1-p)) drawn by the caller in the reference's order; entries may be NULL.
 *   rng: NULL or in-kernel randomness for whichever of the two is not given explicitly (mask index = position in drop_masks). */
int univtg_forward_train(univtg_plan* plan, void* train_ws, const float* src_txt, const float* src_txt_mask,
                         const float* src_vid, const float* src_vid_mask, const float* droppath_scale,
                         const float* const* drop_masks, const univtg_rng* rng, float* pred_logits, float* pred_spans,
                         float* vid_mem_proj, float* txt_mem_proj, float* saliency_scores, void* stream);
/* The multipliers univtg_forward_train draws for `rng`: mask `mask_index` ([rows, cols = din] row-major) and the DropPath
 * scales [n_sites = 2*enc_layers, batch].  Parity tests hand them to the oracle. */
int univtg_dropout_mask(const univtg_rng* rng, int32_t mask_index, size_t rows, size_t cols, float* out, void* stream);
int univtg_droppath_scales(const univtg_rng* rng, int32_t n_sites, int32_t batch, float* out, void* stream);
/* Backward of the last univtg_forward_train on (plan, train_ws).  g_*: upstream gradients of pred_logits [B,Lv,1],
 * pred_spans [B,Lv,2], vid_mem_proj [B,Lv,d], txt_mem_proj [B,1,d] (NULL = zero).  grads: HOST array of device pointers,
 * one ZERO-FILLED fp32 tensor per parameter in univtg_pack_weights order and in the parameter's own layout.
 * grad_scale: power-of-two loss scale S > 0.  Gradient GEMM operands share the plan's 16-bit format (one tcgen05.mma takes
 * A and B in one format); with fp16 operands the intermediate gradients are carried multiplied by S so they do not
 * underflow, and every parameter gradient is multiplied by 1/S where it is written (the results are unscaled).  Use 1 for
 * bf16 plans. */
int univtg_backward(univtg_plan* plan, void* train_ws, const float* src_txt, const float* src_vid, const float* droppath_scale,
                    const float* const* drop_masks, const univtg_rng* rng, const float* g_logits, const float* g_spans,
                    const float* g_vid_mem_proj, const float* g_txt_mem_proj, float grad_scale, float* const* grads,
                    int32_t n_grads, void* stream);

/* Gradient-exchange overlap (the reference relies on DistributedDataParallel's bucketed all-reduce overlapping backward,
 * main/train_vlp_ddp.py:272-275).  univtg_backward finalises parameter gradients in n = enc_layers + 3 stages:
 *   stage 0: conv heads + pooling weight; stage 1 + k: encoder layer enc_layers-1-k; stage n-2: projector weights / biases, the later
 *   projector layers' LayerNorm terms, token-type embedding; stage n-1: LayerNorm terms of the first projector layers (a few KB: the
 *   large input-projection gradients start their exchange before the backward's last kernels run).
 * univtg_backward_stages writes, per stage, two half-open parameter-index ranges {first0, last0, first1, last1}
 * (univtg_pack_weights order; an empty second range is 0,0) and returns n (ranges == NULL: just returns n).
 * univtg_plan_set_grad_events installs n cudaEvent_t handles; univtg_backward records event k on its stream as soon as stage
 * k's gradients are final, so a communication stream can wait on it and reduce that slice while the backward continues.
 * n = 0 removes them.  (enc_layers <= 16, so n <= 19.) */
int univtg_backward_stages(const univtg_config* cfg, int32_t* ranges, int32_t max_stages);
/* The GEMM launches of univtg_backward are persistent grids of one CTA per SM.  When a collective (NCCL) runs beside the backward
 * its CTAs occupy some SMs; give the backward the number of SMs that are left (0 = all) so that its grids stay single-wave. */
int univtg_plan_set_backward_sm_budget(univtg_plan* plan, int32_t num_sms);
int univtg_plan_set_grad_events(univtg_plan* plan, void* const* events, int32_t n);

/* SetCriterion for model_id=univtg (reference model/univtg.py:195-282): losses5 = {loss_b, loss_g, loss_f, loss_s_inter,
 * loss_s_intra}.  targets as main/dataset.py:1078-1098 builds them (all f32 except saliency_pos_idx = saliency_pos_labels[:,0],
 * int64, NULL when absent).  `scratch` (univtg_loss_scratch_bytes) carries per-loss gradients to univtg_loss_backward. */
size_t univtg_loss_scratch_bytes(int32_t B, int32_t Lv);
int univtg_loss_forward(const float* pred_logits, const float* pred_spans, const float* vid_mem_proj, const float* txt_mem_proj,
                        const float* timestamp, const float* timestamp_mask, const float* timestamp_window,
                        const float* span_labels_nn, const float* saliency_scores, const int64_t* saliency_pos_idx, int32_t B,
                        int32_t Lv, int32_t d, float eos_coef, float temperature, float* losses5, void* scratch, void* stream);
/* w5: device fp32 [5] = dL/d(loss_k).  Writes the gradients of the four model outputs. */
int univtg_loss_backward(const float* w5, const float* vid_mem_proj, const float* txt_mem_proj, const int64_t* saliency_pos_idx,
                         int32_t B, int32_t Lv, int32_t d, const void* scratch, float* d_logits, float* d_spans,
                         float* d_vid_mem_proj, float* d_txt_mem_proj, void* stream);

/* Number of kernels one univtg_forward launches (for bench accounting). */
int univtg_forward_num_launches(const univtg_plan* plan);
/* Kernels this library has launched since it was loaded (every launch of every entry point; memsets / memcpys not counted). */
int64_t univtg_launch_count(void);

/* Optional per-launch CUDA-event timeline of univtg_forward (bench / profiling only; adds event records to the stream).
 * read_profile returns the number of launches of the last forward and fills ms[i] / kinds[i]
 * (kind 0 = bandwidth-bound row kernel, 1 = tcgen05 GEMM, 2 = attention); it synchronises on the last event. */
int univtg_plan_set_profiling(univtg_plan* plan, int32_t enable);
int univtg_plan_read_profile(univtg_plan* plan, float* ms, int32_t* kinds, int32_t cap);

/* ---- single operators (unit tests / profiling; same kernels the plan uses) ---- */

/* C[M,N] = act(A*B^T + bias) * alpha.  a: [M,K] (a_mn=0) or [K,M] (a_mn=1); b: [N,K] (b_mn=0) or [K,N] (b_mn=1),
 * 16-bit operands in `fmt`; K and the leading dimensions must be multiples of 8 elements.  out32 [M,N] f32 and/or
 * out16 [M,N] 16-bit.  bn: tile width, multiple of 16 in [32,256] (multiple of 64 when b_mn); ksplit>1 accumulates
 * atomically into a pre-zeroed out32. */
int univtg_op_gemm(const void* a, const void* b, int32_t M, int32_t N, int32_t K, int32_t a_mn, int32_t b_mn, int32_t fmt,
                   int32_t bn, int32_t ksplit, const float* bias, int32_t act, float alpha, float* out32, void* out16,
                   void* stream);
/* Same GEMM launched as 2-CTA clusters: vertically adjacent tiles share their B tile through TMA multicast (half the
 * L2 -> SM traffic of B).  bn multiple of 32 (of 128 when b_mn). */
int univtg_op_gemm_cluster(const void* a, const void* b, int32_t M, int32_t N, int32_t K, int32_t a_mn, int32_t b_mn,
                           int32_t fmt, int32_t bn, int32_t ksplit, const float* bias, int32_t act, float alpha, float* out32,
                           void* out16, void* stream);
/* Profiling aid: when `buf` (device, >= 148*8 uint64) is non-NULL every following GEMM launch stamps %globaltimer per CTA:
 * [0] entry, [1] setup done, [2] all TMA issued, [3] first stage landed, [4] last MMA issued, [5] accumulator ready,
 * [6] epilogue done, [7] exit.  Pass NULL to switch it off. */
int univtg_debug_gemm_timeline(void* buf);
/* Host-only: the tile width / split-K factor the GEMM launcher's cost model picks for a grouped launch (num <= 4 problems of
 * M x N with kblocks 64-wide k-blocks each; step 16 for K-major B, 64 for MN-major B).  Testing / tuning aid. */
int univtg_debug_choose_tile(const int32_t* Ms, const int32_t* Ns, const int32_t* kblocks, int32_t num, int32_t num_sms, int32_t step,
                             int32_t max_split, int32_t* bn, int32_t* ksplit);
/* tcgen05.ld rate probe with the GEMM epilogue's access pattern: out_ns[block] = ns per 16-column step.  Profiling aid only. */
int univtg_debug_tmem_ld_rate(int32_t iters, int32_t mode, int32_t blocks, float* out_ns, float* sink, void* stream);
/* tcgen05.mma issue-rate probe (M=128, N=n, K=16 from resident smem): out_ns[block] = ns per MMA.  Profiling aid only. */
int univtg_debug_mma_rate(int32_t n, int32_t iters, int32_t per_commit, int32_t kstep_bytes, int32_t blocks, float* out_ns,
                          void* stream);
/* Parameter update of the reference's training loop (main/train_vlp_ddp.py:66-68 = main/train_mr.py:64-66; optimizer built at
 * main/config.py:350 as torch.optim.AdamW(lr, weight_decay)) over ONE flat fp32 buffer of n floats (n % 4 == 0, 16-byte
 * aligned; the plugin lays every parameter and its gradient out at the same offsets):
 *   total_norm = ||grads||_2;  if max_grad_norm > 0: g *= min(1, max_grad_norm / (total_norm + 1e-6))   (clip_grad_norm_)
 *   p *= 1 - lr*wd;  m = m + (1-beta1)(g - m);  v = beta2 v + (1-beta2) g^2;
 *   p -= lr/(1-beta1^step) * m / (sqrt(v)/sqrt(1-beta2^step) + eps)                                      (AdamW, step >= 1)
 * scratch3: device fp32 [UNIVTG_ADAMW_SCRATCH_FLOATS] (per-block partial sums of squares live behind the first four floats: the
 * norm is accumulated in a fixed order, without atomics, so the update is bit-reproducible and identical on every data-parallel
 * rank); on return [1] holds total_norm (what clip_grad_norm_ returns) and [2] is 1.0 when total_norm was not
 * finite - then NOTHING was updated (the skipped step of dynamic loss scaling; the fp16 gradient operands of univtg_backward can
 * overflow when grad_scale is too large), else 0.0.  write_clipped_grads != 0 also stores the clipped gradients back
 * (clip_grad_norm_ scales .grad in place).
 * cfg + packed (both non-NULL; the flat buffers must then hold exactly the config's parameters in univtg_pack_weights order, each
 * padded to a multiple of 4 floats): the kernel also refreshes the 16-bit copies of the GEMM weight matrices inside `packed` from the
 * values it has just computed, so no separate re-packing pass re-reads the weights; call univtg_pack_vectors afterwards for the
 * fp32 vectors (LayerNorm terms, biases, token-type rows - a few hundred KB). */
int univtg_adamw_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, size_t n, float lr, float beta1,
                      float beta2, float eps, float weight_decay, int32_t step, float max_grad_norm,
                      int32_t write_clipped_grads, float* scratch3, const univtg_config* cfg, void* packed, void* stream);
/* univtg_pack_weights restricted to the fp32 vectors and the two tiny last-conv tensors (everything that is not a 16-bit matrix). */
int univtg_pack_vectors(const univtg_config* cfg, const float* const* params, int32_t n_params, void* packed, void* stream);

/* HOST function (no device work): assemble one padded batch from a memory-mapped 16-bit feature shard into (pinned) staging
 * buffers - what the reference's per-sample loads + pad_sequences_1d collate do (main/dataset.py:644-696, 1037-1052,
 * utils/tensor_utils.py:6-53).  src_vid [rows, Dv] / src_txt [rows, Dt]: the shard's 16-bit matrices; sample b copies vid_len[b]
 * rows starting at vid_row0[b] (txt likewise) into dst_vid [B, Lv, Dv] / dst_txt [B, Lt, Dt], zero-fills the rest and writes the
 * float masks dst_vmask [B, Lv] / dst_tmask [B, Lt] (1 = valid).  threads > 1 uses a persistent worker pool. */
int univtg_host_assemble_batch(void* dst_vid, void* dst_txt, float* dst_vmask, float* dst_tmask, const void* src_vid,
                               const void* src_txt, const int64_t* vid_row0, const int64_t* txt_row0, const int32_t* vid_len,
                               const int32_t* txt_len, int32_t B, int32_t Lv, int32_t Lt, int32_t Dv, int32_t Dt, int32_t threads);

/* Direct variant: page-lock the shard's memory mapping once (univtg_host_register; enable = 0 undoes it) and let the copy engines
 * pull every sample's rows straight from it into the device batch on `stream` - no CPU staging copy.  mask_stage: pinned host
 * scratch of B * (Lv + Lt) floats that must stay untouched until the stream has passed this call. */
int univtg_host_register(void* base, size_t bytes, int32_t enable);
int univtg_h2d_gather_batch(void* dev_vid, void* dev_txt, float* dev_vmask, float* dev_tmask, float* mask_stage, const void* src_vid,
                            const void* src_txt, const int64_t* vid_row0, const int64_t* txt_row0, const int32_t* vid_len,
                            const int32_t* txt_len, int32_t B, int32_t Lv, int32_t Lt, int32_t Dv, int32_t Dt, void* stream);

/* Post-forward decode of the reference's MR evaluation loop, on the device.
 * univtg_decode_mr = main/inference_mr.py:112-120,146-157 (and main_gradio.py:100-106 with duration == NULL):
 *   score = pred_logits[b,l] (0 where timestamp_mask[b,l] == 0); (st, ed) = (timestamp + pred_spans)[b,l] * duration[b], clamped to
 *   [0, duration[b]] (no scaling / clamping when duration == NULL); rows [st, ed, score] sorted by score, descending, ties in clip
 *   order (Python's stable sort) when sort != 0.  windows [B,Lv,3] f32 receives the rows, windows_r4 [B,Lv,3] f64 (optional) the
 *   same numbers rounded like float(f"{e:.4f}") (exact), order [B,Lv] (optional) the source clip index of every row.  Lv <= 4096.
 * univtg_temporal_nms = utils/temporal_nms.py:25-74 as called by main/inference_mr.py:31-40: per sample, the first
 *   min(n, max_before_nms) rows of windows [B,n,3] f64 (sorted by score) go through greedy NMS with the reference's
 *   intersection / convex-hull "IoU" > nms_thd test in IEEE double; out [B,max_after_nms,3] f64, counts [B] rows kept. */
int univtg_decode_mr(const float* pred_logits, const float* pred_spans, const float* timestamp, const float* timestamp_mask,
                     const float* duration, int32_t B, int32_t Lv, int32_t sort, float* windows, double* windows_r4, int32_t* order,
                     void* stream);
int univtg_temporal_nms(const double* windows, int32_t B, int32_t n, int32_t max_before_nms, double nms_thd, int32_t max_after_nms,
                        double* out, int32_t* counts, void* stream);

/* LayerNorm rows: in [rows,d] f32 -> out32 [rows,d] f32 and/or out16 [rows,ld16] 16-bit (zero padded). */
int univtg_op_layernorm(const float* in, int32_t rows, int32_t d, const float* gamma, const float* beta, float eps,
                        int32_t fmt, float* out32, void* out16, int32_t ld16, void* stream);
/* Attention core.  qkv: [B*L, 3d] 16-bit, column blocks Q | K | V (heads are dh-wide sub-blocks), d = H*dh; scores are
 * scaled by 1/sqrt(dh); key_mask [B,L] f32 (1 = valid key); out [B*L,d] 16-bit; lse [B,H,L] f32 or NULL.
 * impl: 0 = tcgen05 (dh in {64,128}), 1 = SIMT (any dh). */
int univtg_op_attention(const void* qkv, const float* key_mask, void* out, float* lse, int32_t B, int32_t L, int32_t H,
                        int32_t dh, int32_t fmt, int32_t impl, void* stream);

/* Attention core backward.  qkv as above; dO [B*L,d] 16-bit gradient of `out`; O = forward output (16-bit, fmt_act);
 * (same 16-bit format as qkv); lse from the forward; delta_ws [B,H,L] f32 scratch; dqkv32 [B*L,3d] f32 receives
 * dQ | dK | dV.  impl: 0 tcgen05, 1 SIMT. */
int univtg_op_attention_bwd(const void* qkv, const void* dO, const void* O, const float* key_mask, const float* lse,
                            float* delta_ws, float* dqkv32, int32_t B, int32_t L, int32_t H, int32_t dh, int32_t fmt_act,
                            int32_t impl, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNIVTG_B200_H_ */
