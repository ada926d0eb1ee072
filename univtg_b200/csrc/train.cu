// Training path of the C-ABI: forward that keeps what backward needs, the backward pass (reference
// main/train_vlp_ddp.py:56-64: outputs = model(...); losses.backward()) and the criterion entry points.
// GEMM descriptors (tensor maps) are built per call here; shapes vary per training batch anyway (collate pads to the batch max).
#include <stdlib.h>

#include "plan.h"

namespace {

struct Mat16 {  // row-major 16-bit matrix view
  const uint16_t* p;
  int rows, cols, ld;
};

// C[M,N] = sum_k A(m,k) B(n,k).  a_mn: A is stored [K rows, M cols] (else [M rows, K cols]); same for B.
int setup_gemm(GemmProblem& p, Mat16 A, int a_mn, Mat16 B, int b_mn, int M, int N, int K, int bn) {
  init_problem(p);
  p.M = M;
  p.N = N;
  p.a_mn = a_mn;
  p.b_mn = b_mn;
  p.kblk_per_tap = (K + 63) / 64;
  int rc = 0;
  if (!a_mn) {
    rc |= make_tmap_2d(&p.tm_a, A.p, (uint64_t)A.rows, (uint64_t)A.cols, (uint64_t)A.ld, GEMM_BM, 64);
  } else {
    rc |= make_tmap_2d(&p.tm_a, A.p, (uint64_t)A.rows, (uint64_t)A.cols, (uint64_t)A.ld, 64, 64);
    p.ca = OperandCoord{0, 1, 0, 0, 0, 0, 0, 1};
  }
  if (!b_mn) {
    rc |= make_tmap_2d(&p.tm_b, B.p, (uint64_t)B.rows, (uint64_t)B.cols, (uint64_t)B.ld, (uint32_t)bn, 64);
    p.b_box_rows = bn;
  } else {
    rc |= make_tmap_b_mn(p, B.p, (uint64_t)B.rows, (uint64_t)B.cols, (uint64_t)B.ld, bn);
    p.cb = OperandCoord{0, 1, 0, 0, 0, 0, 0, 1};
  }
  return rc;
}

// Tile width + split-K factor for one grouped launch (cost model: choose_tile, gemm.cu).  K in elements; step 64 when a B operand
// is MN-major; max_split = 1 for launches whose epilogue cannot accumulate.
struct MNK {
  int M, N, K;
};
inline TileChoice tile_for(int sms, int step, int max_split, MNK a, MNK b = MNK{0, 0, 0}, MNK c3 = MNK{0, 0, 0}) {
  const int Ms[3] = {a.M, b.M, c3.M}, Ns[3] = {a.N, b.N, c3.N}, kb[3] = {(a.K + 63) / 64, (b.K + 63) / 64, (c3.K + 63) / 64};
  const int num = c3.M > 0 ? 3 : (b.M > 0 ? 2 : 1);
  return choose_tile(Ms, Ns, kb, num, sms, step, max_split);
}
inline int bn_for(int sms, int step, MNK a, MNK b = MNK{0, 0, 0}) { return tile_for(sms, step, 1, a, b).bn; }

struct TrainWs {
  // ---- saved by the forward ----
  uint16_t *a_vid[3], *a_txt[3];
  float *pmean_v[3], *prstd_v[3], *pmean_t[3], *prstd_t[3];
  float *p_vid32[3], *p_txt32[3];  // output of projector layer i (input of LayerNorm i+1)
  float *txtproj32, *pool_alpha, *pos, *key_mask, *pool_logits;
  float* dp_scale;  // [2 * enc_layers, B] DropPath scales drawn in-kernel by the forward (univtg_rng), reused by the backward
  uint16_t *xin16[17], *xpos16[17];  // operands of layer l's in-projections (index enc_layers: unused tail)
  uint16_t *qkv16[16], *attn16[16], *x1_16[16], *h16[16];
  float *lse[16], *y1[16], *mean1[16], *rstd1[16], *y2[16], *mean2[16], *rstd2[16];
  uint16_t* dgelu16[16];  // GELU'(pre-activation) of the FFN, written by FFN1's forward epilogue beside h16
  float *x32, *x1_32;
  uint16_t *hA, *h1, *hc2, *hs2, *br16;
  float *pred_logits, *pred_spans, *vid_mem_proj, *txt_mem_proj;  // copies of the outputs the backward needs
  // ---- backward scratch ----
  float *dx, *dy, *dqkv32, *delta, *dz, *dxt_pool, *dA_v, *dA_t, *wtap;
  uint16_t *dbr16, *dhpre16, *dO16, *dqkv16, *dhc2, *dhs2, *dh1, *dxv16, *dxt16;
  size_t total;
};

TrainWs make_train_ws(const univtg_config& c, const univtg_shape& s, const PackedLayout& L, uint8_t* base) {
  TrainWs w;
  memset(&w, 0, sizeof(w));
  Cursor cur;
  const size_t d = c.hidden_dim, ff = c.dim_feedforward, H = c.nheads;
  const size_t B = s.batch, Lv = s.l_vid, Lt = s.l_txt, Lc = Lv + Lt;
  const size_t M = B * Lc, Mv = B * Lv, Mt = B * Lt, Mh = B * (Lv + 1);
  auto take16 = [&](size_t elems) { return reinterpret_cast<uint16_t*>(base + cur.take(elems * 2)); };
  auto take32 = [&](size_t elems) { return reinterpret_cast<float*>(base + cur.take(elems * 4)); };
  size_t max_din_v = 0, max_din_t = 0;
  for (int i = 0; i < c.n_input_proj; ++i) {
    w.a_vid[i] = take16(Mv * L.vid[i].kpad);
    w.a_txt[i] = take16(Mt * L.txt[i].kpad);
    w.pmean_v[i] = take32(Mv);
    w.prstd_v[i] = take32(Mv);
    w.pmean_t[i] = take32(Mt);
    w.prstd_t[i] = take32(Mt);
    w.p_vid32[i] = take32(Mv * d);
    w.p_txt32[i] = take32(Mt * d);
    if ((size_t)L.vid[i].kpad > max_din_v) max_din_v = L.vid[i].kpad;
    if ((size_t)L.txt[i].kpad > max_din_t) max_din_t = L.txt[i].kpad;
  }
  w.txtproj32 = take32(Mt * d);
  w.pool_alpha = take32(B * Lt);
  w.pool_logits = take32(B * Lt);
  w.pos = take32(Mv * d);
  w.key_mask = take32(B * Lc);
  w.dp_scale = take32((size_t)2 * c.enc_layers * B);
  for (int l = 0; l <= c.enc_layers; ++l) {
    w.xin16[l] = take16(M * d);
    w.xpos16[l] = take16(M * d);
  }
  for (int l = 0; l < c.enc_layers; ++l) {
    w.qkv16[l] = take16(M * 3 * d);
    w.attn16[l] = take16(M * d);
    w.x1_16[l] = take16(M * d);
    w.h16[l] = take16(M * ff);
    w.lse[l] = take32(B * H * Lc);
    w.y1[l] = take32(M * d);
    w.mean1[l] = take32(M);
    w.rstd1[l] = take32(M);
    w.dgelu16[l] = take16(M * ff);
    w.y2[l] = take32(M * d);
    w.mean2[l] = take32(M);
    w.rstd2[l] = take32(M);
  }
  w.x32 = take32(M * d);
  w.x1_32 = take32(M * d);
  w.hA = take16((Mh + 2) * d);
  w.h1 = take16((Mh + 2) * 2 * d);
  w.hc2 = take16((Mh + 2) * d);
  w.hs2 = take16((Mh + 2) * d);
  w.br16 = take16(M * d);
  w.pred_logits = take32(Mv);
  w.pred_spans = take32(Mv * 2);
  w.vid_mem_proj = take32(Mv * d);
  w.txt_mem_proj = take32(B * d);
  w.dx = take32(M * d);
  w.dy = take32(M * d);
  w.dqkv32 = take32(M * 3 * d);
  w.delta = take32(B * H * Lc);
  w.dz = take32((Mh + 2) * 4);
  w.dxt_pool = take32(Mt * d);
  w.dA_v = take32(Mv * max_din_v);
  w.dA_t = take32(Mt * max_din_t);
  {  // tap-major planes of one conv weight gradient; also the K-padded copy of a projector weight gradient whose width is not a multiple of 8
    size_t n = (size_t)3 * d * d;
    if ((size_t)d * max_din_v > n) n = (size_t)d * max_din_v;
    if ((size_t)d * max_din_t > n) n = (size_t)d * max_din_t;
    w.wtap = take32(n);
  }
  w.dbr16 = take16(M * d);
  w.dhpre16 = take16(M * ff);
  w.dO16 = take16(M * d);
  w.dqkv16 = take16(M * 3 * d);
  w.dhc2 = take16((Mh + 2) * d);
  w.dhs2 = take16((Mh + 2) * d);
  w.dh1 = take16((Mh + 2) * 2 * d);
  w.dxv16 = take16(Mv * d);
  w.dxt16 = take16(Mt * d);
  w.total = cur.off;
  return w;
}

// Gradient operands use the plan's 16-bit format (one tcgen05.mma takes A and B in ONE format).  With fp16 they would
// underflow, so the whole backward runs on gradients multiplied by a power-of-two loss scale S: the upstream output
// gradients are scaled on entry, every intermediate stays scaled, and each PARAMETER gradient is multiplied by 1/S where it
// is written (GEMM alpha, column-sum / LayerNorm / head kernels).  bf16 plans simply use S = 1.

}  // namespace

extern "C" {

// Buffers whose untouched rows must read as zeros (conv-head layout [B*(Lv+1)+2, C]: the separator rows that implement the
// Conv1d zero padding, model/univtg.py:375-377).  Every other byte of a workspace is written by a kernel before it is read,
// so a pooled workspace only needs this when it is handed to a different shape (or for the first time).
int univtg_prepare_workspace(const univtg_config* cfg, const univtg_shape* shape, void* workspace, int32_t training_ws, void* stream) {
  if (!check_cfg(cfg) || !check_shape(shape) || !workspace) {
    if (workspace == nullptr) set_error("univtg_prepare_workspace: null workspace");
    return 1;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const PackedLayout L = make_layout(*cfg);
  const size_t d = cfg->hidden_dim, Mh = (size_t)shape->batch * (shape->l_vid + 1);
  uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
  cudaError_t e = cudaSuccess;
  auto zero = [&](void* p, size_t bytes) {
    if (e == cudaSuccess) e = cudaMemsetAsync(p, 0, bytes, st);
  };
  if (training_ws) {
    const TrainWs T = make_train_ws(*cfg, *shape, L, base);
    zero(T.hA, (Mh + 2) * d * 2);
    zero(T.h1, (Mh + 2) * 2 * d * 2);
    zero(T.hc2, (Mh + 2) * d * 2);
    zero(T.hs2, (Mh + 2) * d * 2);
    zero(T.dhc2, (Mh + 2) * d * 2);
    zero(T.dhs2, (Mh + 2) * d * 2);
    zero(T.dh1, (Mh + 2) * 2 * d * 2);
  } else {
    const WsLayout w = make_ws(*cfg, *shape, L);
    zero(base + w.hA, (Mh + 2) * d * 2);
    zero(base + w.h1, (Mh + 2) * 2 * d * 2);
    zero(base + w.hc2, (Mh + 2) * d * 2);
    zero(base + w.hs2, (Mh + 2) * d * 2);
  }
  if (e != cudaSuccess) {
    set_error("univtg_prepare_workspace: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

size_t univtg_train_workspace_bytes(const univtg_config* cfg, const univtg_shape* shape) {
  if (!check_cfg(cfg) || !check_shape(shape)) return 0;
  return make_train_ws(*cfg, *shape, make_layout(*cfg), nullptr).total;
}

// Training forward.  `ws` = training workspace (univtg_train_workspace_bytes, zero-filled once by the caller).
// drop_masks: HOST array of 2*n_input_proj device pointers (video layers, then text layers): fp32 [rows, din_i] input-dropout
// multipliers (0 or 1/(1-p)) or NULL entries / NULL array when dropout is off.
int univtg_forward_train(univtg_plan* P, void* ws, const float* src_txt, const float* src_txt_mask, const float* src_vid,
                         const float* src_vid_mask, const float* droppath_scale, const float* const* drop_masks,
                         const univtg_rng* rng, float* pred_logits, float* pred_spans, float* vid_mem_proj, float* txt_mem_proj,
                         float* saliency_scores, void* stream) {
  if (!P || !ws || !src_txt || !src_txt_mask || !src_vid || !src_vid_mask || !pred_logits || !pred_spans || !vid_mem_proj ||
      !txt_mem_proj || !saliency_scores) {
    set_error("univtg_forward_train: null argument");
    return 1;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const univtg_config& c = P->cfg;
  const PackedLayout& Lw = P->lay;
  const uint8_t* pk = P->packed;
  auto F32 = [&](size_t off) { return reinterpret_cast<const float*>(pk + off); };
  auto W16 = [&](size_t off) { return reinterpret_cast<const uint16_t*>(pk + off); };
  const TrainWs T = make_train_ws(c, P->shp, Lw, reinterpret_cast<uint8_t*>(ws));
  const int d = P->d, ff = P->ff, fmt = c.operand_format, M = P->M, Mv = P->Mv, Mt = P->Mt, Mh = P->Mh, L = P->L, Lv = P->Lv;
  const int sms = P->num_sms;
  int rc = 0;
  GemmGroup g;

  // train-mode randomness: explicit tensors (the caller drew them, e.g. with the reference's torch calls) win over `rng`
  prof_begin(P, st);
  const bool dp_rng = droppath_scale == nullptr && rng != nullptr && rng->droppath > 0.f;
  const bool drop_rng = drop_masks == nullptr && rng != nullptr && rng->input_dropout > 0.f;
  rc = launch_sine_pos(src_vid_mask, src_txt_mask, P->dim_t, T.pos, T.key_mask, P->B, Lv, P->Lt, d, st, dp_rng ? T.dp_scale : nullptr,
                       2 * c.enc_layers, rng ? rng->seed : 0ull, rng ? 1.0f - rng->droppath : 1.f);
  if (rc) return rc;
  if (dp_rng) droppath_scale = T.dp_scale;

  // ---- input projectors ----
  for (int i = 0; i < c.n_input_proj; ++i) {
    for (int s = 0; s < 2; ++s) {
      const ProjPacked& pp = s == 0 ? Lw.vid[i] : Lw.txt[i];
      LnArgs a;
      memset(&a, 0, sizeof(a));
      a.in = i == 0 ? (s == 0 ? src_vid : src_txt) : (s == 0 ? T.p_vid32[i - 1] : T.p_txt32[i - 1]);
      if (i == 0 && P->in_fmt != 0) {
        a.in16 = reinterpret_cast<const uint16_t*>(a.in);
        a.in_fmt = P->in_fmt - 1;
      }
      a.ld_in = pp.din;
      a.rows = s == 0 ? Mv : Mt;
      a.d = pp.din;
      a.gamma = F32(pp.ln_w);
      a.beta = F32(pp.ln_b);
      a.eps = 1e-5f;
      a.fmt = fmt;
      a.out16 = s == 0 ? T.a_vid[i] : T.a_txt[i];
      a.ld16 = pp.kpad;
      a.mul32 = drop_masks ? drop_masks[s * c.n_input_proj + i] : nullptr;
      if (drop_rng) a.drop = make_drop_spec(rng->seed, (unsigned int)(s * c.n_input_proj + i), rng->input_dropout);
      a.mean_out = s == 0 ? T.pmean_v[i] : T.pmean_t[i];
      a.rstd_out = s == 0 ? T.prstd_v[i] : T.prstd_t[i];
      rc = launch_layernorm(a, st);
      if (rc) return rc;
    }
    memset(&g, 0, sizeof(g));
    g.num = 2;
    g.fmt = fmt;
    const bool last = (i == c.n_input_proj - 1);
    rc |= setup_linear(g.p[0], T.a_vid[i], Mv, Lw.vid[i].kpad, Lw.vid[i].kpad, W16(Lw.vid[i].w16), d, Lw.vid[i].kpad, P->bn_proj[i]);
    rc |= setup_linear(g.p[1], T.a_txt[i], Mt, Lw.txt[i].kpad, Lw.txt[i].kpad, W16(Lw.txt[i].w16), d, Lw.txt[i].kpad, P->bn_proj[i]);
    if (rc) return rc;
    g.p[0].bias = F32(Lw.vid[i].bias);
    g.p[1].bias = F32(Lw.txt[i].bias);
    if (!last) {
      g.p[0].act = g.p[1].act = ACT_RELU;
      g.p[0].out32 = T.p_vid32[i];
      g.p[1].out32 = T.p_txt32[i];
      g.p[0].ld32 = g.p[1].ld32 = d;
    } else {
      g.p[0].rps_in = Lv;
      g.p[0].rps_out = L;
      g.p[1].rps_in = P->Lt;
      g.p[1].rps_out = L;
      g.p[1].row_off = Lv;
      for (int s = 0; s < 2; ++s) {
        g.p[s].out32 = T.x32;
        g.p[s].ld32 = d;
        g.p[s].out16 = T.xin16[0];
        g.p[s].out16p = T.xpos16[0];
        g.p[s].ld16 = d;
        g.p[s].ld32_id = d;
      }
      g.p[0].addtab = T.pos;
      g.p[0].ld_addtab = d;
      g.p[0].out32_id = vid_mem_proj;
      g.p[1].out32_id = T.txtproj32;
    }
    rc = gemm_launch(P, g, P->bn_proj[i], sms, st);
    if (rc) return rc;
  }

  // ---- encoder layers ----
  for (int l = 0; l < c.enc_layers; ++l) {
    const LayerPacked& lp = Lw.layer[l];
    memset(&g, 0, sizeof(g));
    g.num = 2;
    g.fmt = fmt;
    rc |= setup_linear(g.p[0], T.xpos16[l], M, d, d, W16(lp.w_in), 2 * d, d, P->bn_qkv);
    rc |= setup_linear(g.p[1], T.xin16[l], M, d, d, W16(lp.w_in) + (size_t)2 * d * d, d, d, P->bn_qkv);
    if (rc) return rc;
    g.p[0].bias = F32(lp.b_in);
    g.p[0].out16 = T.qkv16[l];
    g.p[0].ld16 = 3 * d;
    g.p[1].bias = F32(lp.b_in) + 2 * d;
    g.p[1].out16 = T.qkv16[l] + 2 * d;
    g.p[1].ld16 = 3 * d;
    rc = gemm_launch(P, g, P->bn_qkv, sms, st);
    if (rc) return rc;
    {
      AttnArgs a;
      memset(&a, 0, sizeof(a));
      a.key_mask = T.key_mask;
      a.out = T.attn16[l];
      a.lse = T.lse[l];
      a.scale = 1.0f / sqrtf((float)P->dh);
      a.B = P->B;
      a.L = L;
      a.H = P->H;
      a.dh = P->dh;
      a.d = d;
      a.fmt = fmt;
      if (P->dh == 64 || P->dh == 128) {
        if (make_tmap_2d(&a.tm_qkv, T.qkv16[l], (uint64_t)M, (uint64_t)3 * d, (uint64_t)3 * d, 128, 64)) return 1;
        prof_mark(P, st, 3);
        rc = launch_attention(a, st);
        prof_mark(P, st, 2);
      } else {
        rc = launch_attention_simt(a, T.qkv16[l], st);
      }
      if (rc) return rc;
    }
    memset(&g, 0, sizeof(g));
    g.num = 1;
    g.fmt = fmt;
    rc = setup_linear(g.p[0], T.attn16[l], M, d, d, W16(lp.w_out), d, d, P->bn_out);
    if (rc) return rc;
    g.p[0].bias = F32(lp.b_out);
    g.p[0].rps_in = L;
    g.p[0].rps_out = L;
    g.p[0].row_scale = droppath_scale ? droppath_scale + (size_t)(2 * l) * P->B : nullptr;
    g.p[0].out16 = T.br16;
    g.p[0].ld16 = d;
    rc = gemm_launch(P, g, P->bn_out, sms, st);
    if (rc) return rc;
    {
      LnArgs a;
      memset(&a, 0, sizeof(a));
      a.in = T.x32;
      a.ld_in = d;
      a.add16 = T.br16;
      a.ld_add16 = d;
      a.sum_out = T.y1[l];
      a.rows = M;
      a.d = d;
      a.gamma = F32(lp.n1w);
      a.beta = F32(lp.n1b);
      a.eps = 1e-5f;
      a.fmt = fmt;
      a.out32 = T.x1_32;
      a.out16 = T.x1_16[l];
      a.ld16 = d;
      a.mean_out = T.mean1[l];
      a.rstd_out = T.rstd1[l];
      rc = launch_layernorm(a, st);
      if (rc) return rc;
    }
    memset(&g, 0, sizeof(g));
    g.num = 1;
    g.fmt = fmt;
    rc = setup_linear(g.p[0], T.x1_16[l], M, d, d, W16(lp.w1), ff, d, P->bn_ffn1);
    if (rc) return rc;
    g.p[0].bias = F32(lp.b1);
    g.p[0].act = ACT_GELU;
    g.p[0].out16 = T.h16[l];
    g.p[0].ld16 = ff;
    g.p[0].dact16 = T.dgelu16[l];
    g.p[0].ld_dact = ff;
    rc = gemm_launch(P, g, P->bn_ffn1, sms, st);
    if (rc) return rc;
    memset(&g, 0, sizeof(g));
    g.num = 1;
    g.fmt = fmt;
    rc = setup_linear(g.p[0], T.h16[l], M, ff, ff, W16(lp.w2), d, ff, P->bn_ffn2);
    if (rc) return rc;
    g.p[0].bias = F32(lp.b2);
    g.p[0].rps_in = L;
    g.p[0].rps_out = L;
    g.p[0].row_scale = droppath_scale ? droppath_scale + (size_t)(2 * l + 1) * P->B : nullptr;
    g.p[0].out16 = T.br16;
    g.p[0].ld16 = d;
    rc = gemm_launch(P, g, P->bn_ffn2, sms, st);
    if (rc) return rc;
    {
      LnArgs a;
      memset(&a, 0, sizeof(a));
      a.in = T.x1_32;
      a.ld_in = d;
      a.add16 = T.br16;
      a.ld_add16 = d;
      a.sum_out = T.y2[l];
      a.rows = M;
      a.d = d;
      a.gamma = F32(lp.n2w);
      a.beta = F32(lp.n2b);
      a.eps = 1e-5f;
      a.fmt = fmt;
      a.L = L;
      a.Lv = Lv;
      a.out32 = T.x32;
      a.out16 = T.xin16[l + 1];
      a.out16p = T.xpos16[l + 1];
      a.ld16 = d;
      a.pos = T.pos;
      a.mean_out = T.mean2[l];
      a.rstd_out = T.rstd2[l];
      if (l == c.enc_layers - 1) a.outc = T.hA;
      rc = launch_layernorm(a, st);
      if (rc) return rc;
    }
  }

  // ---- heads ----
  auto conv_problem = [&](GemmProblem& p, const uint16_t* A, int lda, const uint16_t* W, int N, const float* bias, uint16_t* out,
                          int ldo, int bn) -> int {
    init_problem(p);
    p.M = Mh;
    p.N = N;
    p.taps = 3;
    p.kblk_per_tap = d / 64;
    p.ca = OperandCoord{0, 0, 0, 1, 0, 1, 1, 0};
    p.cb = OperandCoord{0, 0, d, 1, 0, 1, 0, 0};
    int r = make_tmap_2d(&p.tm_a, A, (uint64_t)Mh + 2, (uint64_t)d, (uint64_t)lda, GEMM_BM, 64);
    r |= make_tmap_2d(&p.tm_b, W, (uint64_t)N, (uint64_t)3 * d, (uint64_t)3 * d, (uint32_t)bn, 64);
    p.b_box_rows = bn;
    p.bias = bias;
    p.act = ACT_RELU;
    p.rps_in = Lv + 1;
    p.rps_out = Lv + 1;
    p.row_off = 1;
    p.zero_sep = 1;
    p.out16 = out;
    p.ld16 = ldo;
    return r;
  };
  memset(&g, 0, sizeof(g));
  g.num = 1;
  g.fmt = fmt;
  rc = conv_problem(g.p[0], T.hA, d, W16(Lw.conv1_w), 2 * d, F32(Lw.conv1_b), T.h1, 2 * d, P->bn_conv1);
  if (rc) return rc;
  rc = gemm_launch(P, g, P->bn_conv1, sms, st);
  if (rc) return rc;
  memset(&g, 0, sizeof(g));
  g.num = 2;
  g.fmt = fmt;
  rc |= conv_problem(g.p[0], T.h1, 2 * d, W16(Lw.conv2c_w), d, F32(Lw.conv2c_b), T.hc2, d, P->bn_conv2);
  rc |= conv_problem(g.p[1], T.h1 + d, 2 * d, W16(Lw.conv2s_w), d, F32(Lw.conv2s_b), T.hs2, d, P->bn_conv2);
  if (rc) return rc;
  rc = gemm_launch(P, g, P->bn_conv2, sms, st);
  if (rc) return rc;
  {
    HeadFinalArgs a;
    a.h_cls = T.hc2;
    a.h_span = T.hs2;
    a.w_cls = F32(Lw.conv3c_w);
    a.w_span = F32(Lw.conv3s_w);
    a.b_cls = F32(Lw.conv3c_b);
    a.b_span = F32(Lw.conv3s_b);
    a.pred_logits = pred_logits;
    a.pred_spans = pred_spans;
    a.B = P->B;
    a.Lv = Lv;
    a.d = d;
    a.fmt = fmt;
    rc = launch_conv_head_final(a, st);
    if (rc) return rc;
  }
  {
    PoolSalArgs a;
    a.x_txt = T.txtproj32;
    a.x_vid = vid_mem_proj;
    a.txt_mask = src_txt_mask;
    a.vid_mask = src_vid_mask;
    a.w = F32(Lw.pool_w);
    a.pooled = txt_mem_proj;
    a.saliency = saliency_scores;
    a.alpha_out = T.pool_alpha;
    a.logits_ws = T.pool_logits;
    a.B = P->B;
    a.Lt = P->Lt;
    a.Lv = Lv;
    a.d = d;
    rc = launch_pool_saliency(a, st);
    if (rc) return rc;
  }
  // keep the small outputs the backward needs (the caller owns the returned tensors and may free them)
  cudaMemcpyAsync(T.pred_logits, pred_logits, (size_t)Mv * 4, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(T.pred_spans, pred_spans, (size_t)Mv * 8, cudaMemcpyDeviceToDevice, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("univtg_forward_train: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

// Backward of univtg_forward_train.  g_*: upstream gradients of the four differentiable outputs (NULL = zero).
// grads: HOST array of device pointers, one fp32 gradient tensor per parameter in univtg_pack_weights order; the tensors must
// be zero-filled by the caller (several are accumulated atomically); they are written in the parameters' native layouts.
// drop_masks / droppath_scale: the same arrays that were passed to the forward.
int univtg_backward(univtg_plan* P, void* ws, const float* src_txt, const float* src_vid, const float* droppath_scale,
                    const float* const* drop_masks, const univtg_rng* rng, const float* g_logits, const float* g_spans,
                    const float* g_vid_mem_proj, const float* g_txt_mem_proj, float grad_scale, float* const* grads,
                    int32_t n_grads, void* stream) {
  if (!P || !ws || !grads || !src_txt || !src_vid || !(grad_scale > 0.f)) {
    set_error("univtg_backward: null argument");
    return 1;
  }
  const univtg_config& c = P->cfg;
  if (n_grads != univtg_num_params(&c)) {
    set_error("univtg_backward: expected %d gradient tensors, got %d", univtg_num_params(&c), n_grads);
    return 1;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const PackedLayout& Lw = P->lay;
  const uint8_t* pk = P->packed;
  auto F32 = [&](size_t off) { return reinterpret_cast<const float*>(pk + off); };
  auto W16 = [&](size_t off) { return reinterpret_cast<const uint16_t*>(pk + off); };
  const TrainWs T = make_train_ws(c, P->shp, Lw, reinterpret_cast<uint8_t*>(ws));
  if (droppath_scale == nullptr && rng != nullptr && rng->droppath > 0.f) droppath_scale = T.dp_scale;  // drawn by the forward
  const bool drop_rng = drop_masks == nullptr && rng != nullptr && rng->input_dropout > 0.f;
  const int d = P->d, ff = P->ff, fmt = c.operand_format, M = P->M, Mv = P->Mv, Mt = P->Mt, Mh = P->Mh, L = P->L, Lv = P->Lv,
            Lt = P->Lt, B = P->B;
  // persistent GEMM grids assume every CTA is resident at once; when a gradient all-reduce runs beside the backward its CTAs
  // hold some SMs, and a 148-CTA grid would wait for them (a second wave): launch on the SMs that are left
  const int sms = (P->num_sms_bwd > 0 && P->num_sms_bwd < P->num_sms) ? P->num_sms_bwd : P->num_sms;
  const int FMT_G = fmt;                 // gradient operand format == activation operand format
  const float GS = grad_scale;           // loss scale carried by every intermediate gradient
  const float INV = 1.0f / grad_scale;   // applied wherever a parameter gradient is written

  int rc = 0;
  GemmGroup g;
  // parameter-gradient index map (univtg_pack_weights order)
  const int np = c.n_input_proj;
  auto G_vid = [&](int i, int k) { return grads[4 * i + k]; };
  auto G_txt = [&](int i, int k) { return grads[4 * np + 4 * i + k]; };
  float* G_type = grads[8 * np];
  auto G_layer = [&](int l, int k) { return grads[8 * np + 1 + 12 * l + k]; };
  const int hb = 8 * np + 1 + 12 * c.enc_layers;
  auto G_span = [&](int k) { return grads[hb + k]; };
  auto G_cls = [&](int k) { return grads[hb + 6 + k]; };
  float* G_pool = grads[hb + 12];

  cudaMemsetAsync(T.dx, 0, (size_t)M * d * 4, st);

  // ================================================ heads ================================================
  if (g_logits && g_spans) {
    HeadFinalBwdArgs a;
    memset(&a, 0, sizeof(a));
    a.g_logits = g_logits;
    a.g_spans = g_spans;
    a.pred_logits = T.pred_logits;
    a.pred_spans = T.pred_spans;
    a.h_cls = T.hc2;
    a.h_span = T.hs2;
    a.w_cls = F32(Lw.conv3c_w);
    a.w_span = F32(Lw.conv3s_w);
    a.dz = T.dz;
    a.dh_cls = T.dhc2;
    a.dh_span = T.dhs2;
    a.gw_cls = G_cls(4);
    a.gb_cls = G_cls(5);
    a.gw_span = G_span(4);
    a.gb_span = G_span(5);
    a.cs_cls = G_cls(3);   // bias gradient of class_embed.layers.1 = column sums of d(hidden 2)
    a.cs_span = G_span(3);
    a.B = B;
    a.Lv = Lv;
    a.d = d;
    a.fmt_act = fmt;
    a.fmt_grad = FMT_G;
    a.in_scale = GS;
    a.pgrad_scale = INV;
    rc = launch_head_final_bwd(a, st);
    if (rc) return rc;

    // conv layout helpers: buffer row = logical row + 1
    // dgrad of a k=3 conv: dX[m] = sum_t' dY[m + t' - 1] W[:, :, 2 - t']  (A = dY K-major with row shift, B = packed W MN-major)
    auto conv_dgrad = [&](GemmProblem& p, const uint16_t* dY, int ldy, int Kc /*out channels*/, const uint16_t* Wp /*[Kc, 3*Cin]*/,
                          int Cin, int bnn) -> int {
      init_problem(p);
      p.M = Mh;
      p.N = Cin;
      p.taps = 3;
      p.kblk_per_tap = Kc / 64;
      p.b_mn = 1;
      p.a_fmt = FMT_G;
      p.b_fmt = fmt;
      p.ca = OperandCoord{0, 0, 0, 1, 0, 1, 1, 0};              // rows m0 + t', cols k
      p.cb = OperandCoord{2 * Cin, 1, -Cin, 0, 0, 0, 0, 1};     // cols n0 + (2 - t') * Cin, rows k (out channel)
      int r = make_tmap_2d(&p.tm_a, dY, (uint64_t)Mh + 2, (uint64_t)Kc, (uint64_t)ldy, GEMM_BM, 64);
      r |= make_tmap_b_mn(p, Wp, (uint64_t)Kc, (uint64_t)3 * Cin, (uint64_t)3 * Cin, bnn);
      return r;
    };
    // wgrad of one tap: dW[n, c, t] = sum_m dY[m, n] X[m + t - 1, c]  -> written with column stride 3 into [N, C, 3]
    const TileChoice t_cw = tile_for(sms, 64, 8, MNK{d, d, Mh}, MNK{d, d, Mh}, MNK{d, d, Mh});
    auto conv_wgrad = [&](GemmProblem& p, const uint16_t* dY, int ldy, int Nc, const uint16_t* X, int ldx, int Cin, int t,
                          float* gw) -> int {
      init_problem(p);
      p.M = Nc;
      p.N = Cin;
      p.a_mn = 1;
      p.b_mn = 1;
      p.a_fmt = FMT_G;
      p.b_fmt = fmt;
      p.kblk_per_tap = (Mh + 63) / 64;
      p.ca = OperandCoord{0, 1, 0, 0, 1, 0, 0, 1};      // cols m0 (out channel), rows 1 + k
      p.cb = OperandCoord{0, 1, 0, 0, t, 0, 0, 1};      // cols n0 (in channel), rows t + k
      int r = make_tmap_2d(&p.tm_a, dY, (uint64_t)Mh + 2, (uint64_t)Nc, (uint64_t)ldy, 64, 64);
      r |= make_tmap_b_mn(p, X, (uint64_t)Mh + 2, (uint64_t)Cin, (uint64_t)ldx, t_cw.bn);
      (void)gw;  // written by launch_tap_interleave from the tap-major planes (256-bit stores here instead of stride-3 scalars)
      p.out32 = T.wtap + (size_t)t * Nc * Cin;
      p.ld32 = Cin;
      p.alpha = INV;
      p.ksplit = t_cw.ksplit;
      return r;
    };
    const int bn_c2d = bn_for(sms, 64, MNK{Mh, d, 3 * d}, MNK{Mh, d, 3 * d}), bn_c1d = bn_for(sms, 64, MNK{Mh, d, 6 * d});
    // (splitting the k-blocks of the layer-1 dgrad - 76 tiles of 96 k-blocks - over idle SMs saves ~8 us but reduces into the stream
    // gradient with atomics, which makes every gradient upstream of the heads order-dependent in its last bits: not taken)
    const int bn_cw = t_cw.bn;
    const int bn = bn_c2d;
    // ---- conv layer 2 (two heads): dgrad -> dh1 [Mh+2, 2d] (class cols [0,d), span cols [d,2d)), ReLU mask of h1 ----
    memset(&g, 0, sizeof(g));
    g.num = 2;
    g.fmt = fmt;
    rc |= conv_dgrad(g.p[0], T.dhc2, d, d, W16(Lw.conv2c_w), d, bn);
    rc |= conv_dgrad(g.p[1], T.dhs2, d, d, W16(Lw.conv2s_w), d, bn);
    if (rc) return rc;
    for (int s = 0; s < 2; ++s) {
      GemmProblem& p = g.p[s];
      p.rps_in = Lv + 1;
      p.rps_out = Lv + 1;
      p.row_off = 1;
      p.zero_sep = 1;
      p.mask16 = T.h1 + s * d;
      p.ld_mask = 2 * d;
      p.out16 = T.dh1 + s * d;
      p.ld16 = 2 * d;
      p.out_fmt = FMT_G;
      p.colsum = s == 0 ? G_cls(1) : G_span(1);  // bias gradient of conv layer 0 (a separate column-sum pass measured 30 us/step slower)
      p.colsum_scale = INV;
    }
    rc = gemm_launch(P, g, bn_c2d, sms, st);
    if (rc) return rc;
    // wgrad conv layer 2: 2 heads x 3 taps
    for (int s = 0; s < 2; ++s) {
      memset(&g, 0, sizeof(g));
      g.num = 3;
      g.fmt = fmt;
      for (int t = 0; t < 3; ++t)
        rc |= conv_wgrad(g.p[t], s == 0 ? T.dhc2 : T.dhs2, d, d, T.h1 + s * d, 2 * d, d, t, s == 0 ? G_cls(2) : G_span(2));
      if (rc) return rc;
      if (t_cw.ksplit > 1) cudaMemsetAsync(T.wtap, 0, (size_t)3 * d * d * 4, st);  // split-K accumulates into the planes
      rc = gemm_launch(P, g, bn_cw, sms, st);
      if (rc) return rc;
      rc = launch_tap_interleave(T.wtap, s == 0 ? G_cls(2) : G_span(2), d, d, st);
      if (rc) return rc;
    }
    // ---- conv layer 1 (fused N = 2d): dgrad -> stream gradient of the video rows ----
    memset(&g, 0, sizeof(g));
    g.num = 1;
    g.fmt = fmt;
    rc = conv_dgrad(g.p[0], T.dh1, 2 * d, 2 * d, W16(Lw.conv1_w), d, bn_c1d);
    if (rc) return rc;
    g.p[0].rps_in = Lv + 1;
    g.p[0].rps_out = L;
    g.p[0].row_off = 0;
    g.p[0].skip_sep = 1;
    g.p[0].out32 = T.dx;
    g.p[0].ld32 = d;
    rc = gemm_launch(P, g, bn_c1d, sms, st);
    if (rc) return rc;
    // wgrad conv layer 1: class rows [0,d) and span rows [d,2d) of the fused weight
    for (int s = 0; s < 2; ++s) {
      memset(&g, 0, sizeof(g));
      g.num = 3;
      g.fmt = fmt;
      for (int t = 0; t < 3; ++t) rc |= conv_wgrad(g.p[t], T.dh1 + s * d, 2 * d, d, T.hA, d, d, t, s == 0 ? G_cls(0) : G_span(0));
      if (rc) return rc;
      if (t_cw.ksplit > 1) cudaMemsetAsync(T.wtap, 0, (size_t)3 * d * d * 4, st);
      rc = gemm_launch(P, g, bn_cw, sms, st);
      if (rc) return rc;
      rc = launch_tap_interleave(T.wtap, s == 0 ? G_cls(0) : G_span(0), d, d, st);
      if (rc) return rc;
    }
  }

  // stage events: gradients of a parameter group are final once the stream reaches this point (univtg_backward_stages)
  auto stage_done = [&](int k) {
    if (P->n_grad_events > 0 && k < P->n_grad_events) cudaEventRecord(P->grad_events[k], st);
  };
  // attention pooling of the text tokens: needs only the loss gradient of the pooled vector, so it runs before the first stage
  // event and its weight gradient travels with the heads' slice
  if (g_txt_mem_proj) {
    PoolBwdArgs a;
    a.x_txt = T.txtproj32;
    a.alpha = T.pool_alpha;
    a.w = F32(Lw.pool_w);
    a.g_pooled = g_txt_mem_proj;
    a.dx_txt = T.dxt_pool;
    a.gw = G_pool;
    a.out_scale = GS;
    a.B = B;
    a.Lt = Lt;
    a.d = d;
    rc = launch_pool_bwd(a, st);
    if (rc) return rc;
  }
  stage_done(0);  // conv heads + weightedpool.weight

  // ================================================ encoder ================================================
  for (int l = c.enc_layers - 1; l >= 0; --l) {
    const LayerPacked& lp = Lw.layer[l];
    const float* s1 = droppath_scale ? droppath_scale + (size_t)(2 * l) * B : nullptr;
    const float* s2 = droppath_scale ? droppath_scale + (size_t)(2 * l + 1) * B : nullptr;
    const int bn_dff = bn_for(sms, 64, MNK{M, ff, d}), bn_dd1 = bn_for(sms, 64, MNK{M, d, ff}), bn_ddo = bn_for(sms, 64, MNK{M, d, d}),
              bn_ddq = bn_for(sms, 64, MNK{M, d, 3 * d});
    const TileChoice t_wo = tile_for(sms, 64, 16, MNK{d, d, M}), t_wq = tile_for(sms, 64, 16, MNK{2 * d, d, M}, MNK{d, d, M}),
                     t_wf = tile_for(sms, 64, 16, MNK{d, ff, M}, MNK{ff, d, M});
    const int bn_wo = t_wo.bn, bn_wq = t_wq.bn;
    // ---- LN2 backward: dx (grad of the layer output) -> dy (grad of x1 + s2 * F), branch operand s2 * dy ----
    {
      LnBwdArgs a;
      memset(&a, 0, sizeof(a));
      a.dout = T.dx;
      a.ld_dout = d;
      a.y = T.y2[l];
      a.ld_y = d;
      a.mean = T.mean2[l];
      a.rstd = T.rstd2[l];
      a.gamma = F32(lp.n2w);
      a.rows = M;
      a.d = d;
      a.row_scale = s2;
      a.L = L;
      a.dy32 = T.dy;
      a.dbr16 = T.dbr16;
      a.ld16 = d;
      a.fmt16 = FMT_G;
      a.dgamma = G_layer(l, 10);
      a.dbeta = G_layer(l, 11);
      a.colsum = G_layer(l, 7);  // linear2.bias
      a.pgrad_scale = INV;
      rc = launch_layernorm_bwd(a, st);
      if (rc) return rc;
    }
    // ---- FFN2: dgrad -> d(hpre) = (dF W2) * gelu'(hpre) (saved 16-bit derivative);  wgrad dW2 = dF^T h ----
    // ---- FFN1: dgrad d(x1) = dhpre W1 + dy (residual);          wgrad dW1 = dhpre^T x1 ----
    // (sharing one launch between a data-gradient GEMM and the weight-gradient GEMM that reads the same dY was measured: fewer
    //  launches and 10 % less event-timed GEMM time, but the pipelined step got 50 us SLOWER - profiles/README.md round 2 - so the
    //  launches stay separate and are chained by programmatic dependent launch)
    auto ffn2_dgrad = [&](GemmProblem& p, int bnn) -> int {
      int r = setup_gemm(p, Mat16{T.dbr16, M, d, d}, 0, Mat16{W16(lp.w2), d, ff, ff}, 1, M, ff, d, bnn);
      p.a_fmt = FMT_G;
      p.b_fmt = fmt;
      p.mask16 = T.dgelu16[l];  // d(hpre) = (dF W2) * GELU'(hpre), the derivative saved by the forward as a 16-bit operand
      p.ld_mask = ff;
      p.mask_mul = 1;
      p.out16 = T.dhpre16;
      p.ld16 = ff;
      p.out_fmt = FMT_G;
      p.colsum = G_layer(l, 5);  // linear1.bias
      p.colsum_scale = INV;
      return r;
    };
    auto ffn2_wgrad = [&](GemmProblem& p, int bnn, int ks) -> int {
      int r = setup_gemm(p, Mat16{T.dbr16, M, d, d}, 1, Mat16{T.h16[l], M, ff, ff}, 1, d, ff, M, bnn);
      p.a_fmt = FMT_G;
      p.b_fmt = fmt;
      p.out32 = G_layer(l, 6);  // linear2.weight [d, ff]
      p.ld32 = ff;
      p.alpha = INV;
      p.ksplit = ks;
      return r;
    };
    auto ffn1_dgrad = [&](GemmProblem& p, int bnn) -> int {
      int r = setup_gemm(p, Mat16{T.dhpre16, M, ff, ff}, 0, Mat16{W16(lp.w1), ff, d, d}, 1, M, d, ff, bnn);
      p.a_fmt = FMT_G;
      p.b_fmt = fmt;
      p.resid = T.dy;
      p.ld_resid = d;
      p.out32 = T.dx;
      p.ld32 = d;
      return r;
    };
    auto ffn1_wgrad = [&](GemmProblem& p, int bnn, int ks) -> int {
      int r = setup_gemm(p, Mat16{T.dhpre16, M, ff, ff}, 1, Mat16{T.x1_16[l], M, d, d}, 1, ff, d, M, bnn);
      p.a_fmt = FMT_G;
      p.b_fmt = fmt;
      p.out32 = G_layer(l, 4);  // linear1.weight [ff, d]
      p.ld32 = d;
      p.alpha = INV;
      p.ksplit = ks;
      return r;
    };
    {
      memset(&g, 0, sizeof(g));
      g.num = 1;
      g.fmt = fmt;
      rc = ffn2_dgrad(g.p[0], bn_dff);
      if (rc) return rc;
      rc = gemm_launch(P, g, bn_dff, sms, st);
      if (rc) return rc;

      memset(&g, 0, sizeof(g));
      g.num = 2;
      g.fmt = fmt;
      rc |= ffn2_wgrad(g.p[0], t_wf.bn, t_wf.ksplit);
      rc |= ffn1_wgrad(g.p[1], t_wf.bn, t_wf.ksplit);
      if (rc) return rc;
      rc = gemm_launch(P, g, t_wf.bn, sms, st);
      if (rc) return rc;
      memset(&g, 0, sizeof(g));
      g.num = 1;
      g.fmt = fmt;
      rc = ffn1_dgrad(g.p[0], bn_dd1);
      if (rc) return rc;
      rc = gemm_launch(P, g, bn_dd1, sms, st);
      if (rc) return rc;
    }
    // ---- LN1 backward ----
    {
      LnBwdArgs a;
      memset(&a, 0, sizeof(a));
      a.dout = T.dx;
      a.ld_dout = d;
      a.y = T.y1[l];
      a.ld_y = d;
      a.mean = T.mean1[l];
      a.rstd = T.rstd1[l];
      a.gamma = F32(lp.n1w);
      a.rows = M;
      a.d = d;
      a.row_scale = s1;
      a.L = L;
      a.dy32 = T.dy;
      a.dbr16 = T.dbr16;
      a.ld16 = d;
      a.fmt16 = FMT_G;
      a.dgamma = G_layer(l, 8);
      a.dbeta = G_layer(l, 9);
      a.colsum = G_layer(l, 3);  // out_proj.bias
      a.pgrad_scale = INV;
      rc = launch_layernorm_bwd(a, st);
      if (rc) return rc;
    }
    // ---- out-proj: dgrad -> dO (16-bit); wgrad dWo = dA^T attn ----
    auto out_dgrad = [&](GemmProblem& p, int bnn) -> int {
      int r = setup_gemm(p, Mat16{T.dbr16, M, d, d}, 0, Mat16{W16(lp.w_out), d, d, d}, 1, M, d, d, bnn);
      p.a_fmt = FMT_G;
      p.b_fmt = fmt;
      p.out16 = T.dO16;
      p.ld16 = d;
      p.out_fmt = FMT_G;
      return r;
    };
    auto out_wgrad = [&](GemmProblem& p, int bnn, int ks) -> int {
      int r = setup_gemm(p, Mat16{T.dbr16, M, d, d}, 1, Mat16{T.attn16[l], M, d, d}, 1, d, d, M, bnn);
      p.a_fmt = FMT_G;
      p.b_fmt = fmt;
      p.out32 = G_layer(l, 2);
      p.ld32 = d;
      p.alpha = INV;
      p.ksplit = ks;
      return r;
    };
    {
      memset(&g, 0, sizeof(g));
      g.num = 1;
      g.fmt = fmt;
      rc = out_dgrad(g.p[0], bn_ddo);
      if (rc) return rc;
      rc = gemm_launch(P, g, bn_ddo, sms, st);
      if (rc) return rc;
      memset(&g, 0, sizeof(g));
      g.num = 1;
      g.fmt = fmt;
      rc = out_wgrad(g.p[0], bn_wo, t_wo.ksplit);
      if (rc) return rc;
      rc = gemm_launch(P, g, bn_wo, sms, st);
      if (rc) return rc;
    }
    // ---- attention core backward -> dqkv32 -> dqkv16 (+ in_proj_bias gradient) ----
    rc = launch_attn_delta(T.dO16, FMT_G, T.attn16[l], fmt, T.delta, B, L, P->H, P->dh, st);
    if (rc) return rc;
    bool fused16 = false;
    {
      AttnBwdArgs a;
      memset(&a, 0, sizeof(a));
      a.qkv = T.qkv16[l];
      a.dO = T.dO16;
      a.key_mask = T.key_mask;
      a.lse = T.lse[l];
      a.delta = T.delta;
      a.dqkv32 = T.dqkv32;
      a.scale = 1.0f / sqrtf((float)P->dh);
      a.B = B;
      a.L = L;
      a.H = P->H;
      a.dh = P->dh;
      a.d = d;
      a.fmt_act = fmt;
      a.fmt_grad = FMT_G;
      const bool tc = (P->dh == 64 || P->dh == 128);
      const int num_kv = (L + 127) / 128;
      a.dq_atomic = (!tc || num_kv > 1) ? 1 : 0;
      fused16 = !a.dq_atomic;  // one key tile: the kernel emits the 16-bit operands and the in_proj_bias gradient itself
      if (fused16) a.dqkv16 = T.dqkv16;  // (the kernel could also accumulate the column sums; a separate pass measured faster)
      if (a.dq_atomic) cudaMemsetAsync(T.dqkv32, 0, (size_t)M * 3 * d * 4, st);
      if (tc) {
        if (make_tmap_2d(&a.tm_qkv, T.qkv16[l], (uint64_t)M, (uint64_t)3 * d, (uint64_t)3 * d, 128, 64)) return 1;
        if (make_tmap_2d(&a.tm_do, T.dO16, (uint64_t)M, (uint64_t)d, (uint64_t)d, 128, 64)) return 1;
        prof_mark(P, st, 3);
        rc = launch_attention_bwd(a, st);
        prof_mark(P, st, 2);
      } else {
        rc = launch_attention_bwd_simt(a, st);
      }
      if (rc) return rc;
    }
    if (!fused16) rc = launch_cvt16_colsum(T.dqkv32, 3 * d, T.dqkv16, 3 * d, M, 3 * d, FMT_G, G_layer(l, 1), INV, st);
    else rc = launch_colsum16(T.dqkv16, 3 * d, M, 3 * d, FMT_G, G_layer(l, 1), INV, st);  // in_proj_bias gradient
    if (rc) return rc;
    // ---- in-projections: dgrad dx = dy + [dq|dk|dv] [Wq;Wk;Wv]; wgrad dWqk = [dq|dk]^T (x+pos), dWv = dv^T x ----
    auto qkv_dgrad = [&](GemmProblem& p, int bnn) -> int {
      int r = setup_gemm(p, Mat16{T.dqkv16, M, 3 * d, 3 * d}, 0, Mat16{W16(lp.w_in), 3 * d, d, d}, 1, M, d, 3 * d, bnn);
      p.a_fmt = FMT_G;
      p.b_fmt = fmt;
      p.resid = T.dy;
      p.ld_resid = d;
      p.out32 = T.dx;
      p.ld32 = d;
      return r;
    };
    auto qkv_wgrads = [&](GemmProblem& pqk, GemmProblem& pv, int bnn, int ks) -> int {
      int r = setup_gemm(pqk, Mat16{T.dqkv16, M, 2 * d, 3 * d}, 1, Mat16{T.xpos16[l], M, d, d}, 1, 2 * d, d, M, bnn);
      r |= setup_gemm(pv, Mat16{T.dqkv16 + 2 * d, M, d, 3 * d}, 1, Mat16{T.xin16[l], M, d, d}, 1, d, d, M, bnn);
      pqk.a_fmt = pv.a_fmt = FMT_G;
      pqk.b_fmt = pv.b_fmt = fmt;
      pqk.out32 = G_layer(l, 0);
      pqk.ld32 = d;
      pv.out32 = G_layer(l, 0) + (size_t)2 * d * d;
      pv.ld32 = d;
      pqk.alpha = pv.alpha = INV;
      pqk.ksplit = pv.ksplit = ks;
      return r;
    };
    {
      memset(&g, 0, sizeof(g));
      g.num = 1;
      g.fmt = fmt;
      rc = qkv_dgrad(g.p[0], bn_ddq);
      if (rc) return rc;
      rc = gemm_launch(P, g, bn_ddq, sms, st);
      if (rc) return rc;
      memset(&g, 0, sizeof(g));
      g.num = 2;
      g.fmt = fmt;
      rc = qkv_wgrads(g.p[0], g.p[1], bn_wq, t_wq.ksplit);
      if (rc) return rc;
      rc = gemm_launch(P, g, bn_wq, sms, st);
      if (rc) return rc;
    }
    stage_done(1 + (c.enc_layers - 1 - l));  // encoder layer l
  }

  // ================================================ projectors ================================================
  // gradient w.r.t. the projected tokens = stream gradient rows + direct (saliency-loss) gradients (dxt_pool: written up front)
  // column sums = bias gradient of the last projector layer AND the token-type embedding rows
  rc = launch_stream_gather(T.dx, L, 0, g_vid_mem_proj, GS, T.dxv16, G_type + d, INV, B, Lv, d, FMT_G, st);
  if (rc) return rc;
  rc = launch_stream_gather(T.dx, L, Lv, g_txt_mem_proj ? T.dxt_pool : nullptr, 1.0f, T.dxt16, G_type, INV, B, Lt, d, FMT_G, st);
  if (rc) return rc;
  cudaMemcpyAsync(G_vid(np - 1, 3), G_type + d, (size_t)d * 4, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(G_txt(np - 1, 3), G_type, (size_t)d * 4, cudaMemcpyDeviceToDevice, st);
  for (int i = np - 1; i >= 0; --i) {
    // wgrad: dW_i = dOut^T a_i   (video + text in one launch)
    memset(&g, 0, sizeof(g));
    g.num = 2;
    g.fmt = fmt;
    const int kpv = Lw.vid[i].kpad, kpt = Lw.txt[i].kpad, dinv = Lw.vid[i].din, dint = Lw.txt[i].din;
    // a width like 2818 would force scalar epilogue stores: write rows padded to kpad with 256-bit stores, then a pitched copy
    const bool pad_v = (dinv % 8) != 0;
    const int nv = pad_v ? kpv : dinv;  // the operand a_vid[i] is zero beyond dinv
    const TileChoice t_pw = tile_for(sms, 64, 16, MNK{d, nv, Mv}, MNK{d, dint, Mt});
    const int bn = t_pw.bn;
    const int bn_pd = bn_for(sms, 64, MNK{Mv, kpv, d}, MNK{Mt, kpt, d});
    rc |= setup_gemm(g.p[0], Mat16{T.dxv16, Mv, d, d}, 1, Mat16{T.a_vid[i], Mv, kpv, kpv}, 1, d, nv, Mv, bn);
    rc |= setup_gemm(g.p[1], Mat16{T.dxt16, Mt, d, d}, 1, Mat16{T.a_txt[i], Mt, kpt, kpt}, 1, d, dint, Mt, bn);
    if (rc) return rc;
    g.p[0].a_fmt = g.p[1].a_fmt = FMT_G;
    g.p[0].b_fmt = g.p[1].b_fmt = fmt;
    g.p[0].out32 = pad_v ? T.wtap : G_vid(i, 2);
    g.p[0].ld32 = nv;
    g.p[1].out32 = G_txt(i, 2);
    g.p[1].ld32 = dint;
    g.p[0].alpha = g.p[1].alpha = INV;
    g.p[0].ksplit = g.p[1].ksplit = t_pw.ksplit;
    if (pad_v && t_pw.ksplit > 1) cudaMemsetAsync(T.wtap, 0, (size_t)d * kpv * 4, st);  // split-K accumulates into the padded copy
    rc = gemm_launch(P, g, bn, sms, st);
    if (rc) return rc;
    if (pad_v)
      cudaMemcpy2DAsync(G_vid(i, 2), (size_t)dinv * 4, T.wtap, (size_t)kpv * 4, (size_t)dinv * 4, (size_t)d, cudaMemcpyDeviceToDevice, st);
    // every projector weight, the later layers' LayerNorm terms and all biases are final here; what follows only produces the
    // first layer's LayerNorm terms - the 11.5 MB video weight gradient need not wait for it to start its exchange
    if (i == 0) stage_done(c.enc_layers + 1);
    // dgrad: dA_i = dOut W_i  (fp32, [rows, kpad_i]: N is padded to the packed weight's K so the epilogue stays on its
    // 128-bit path even for the 2818-wide video features; the padded columns are zeros).  The input-dropout mask is applied
    // by the LayerNorm backward when it loads dA.
    memset(&g, 0, sizeof(g));
    g.num = 2;
    g.fmt = fmt;
    rc |= setup_gemm(g.p[0], Mat16{T.dxv16, Mv, d, d}, 0, Mat16{W16(Lw.vid[i].w16), d, kpv, kpv}, 1, Mv, kpv, d, bn_pd);
    rc |= setup_gemm(g.p[1], Mat16{T.dxt16, Mt, d, d}, 0, Mat16{W16(Lw.txt[i].w16), d, kpt, kpt}, 1, Mt, kpt, d, bn_pd);
    if (rc) return rc;
    g.p[0].a_fmt = g.p[1].a_fmt = FMT_G;
    g.p[0].b_fmt = g.p[1].b_fmt = fmt;
    g.p[0].out32 = T.dA_v;
    g.p[0].ld32 = kpv;
    g.p[1].out32 = T.dA_t;
    g.p[1].ld32 = kpt;
    rc = gemm_launch(P, g, bn_pd, sms, st);
    if (rc) return rc;
    // LayerNorm_i backward: parameter gradients; for i > 0 also the gradient of the previous layer's ReLU output
    for (int s = 0; s < 2; ++s) {
      LnBwdArgs a;
      memset(&a, 0, sizeof(a));
      const ProjPacked& pp = s == 0 ? Lw.vid[i] : Lw.txt[i];
      a.dout = s == 0 ? T.dA_v : T.dA_t;
      a.ld_dout = pp.kpad;
      a.dout_mul = drop_masks ? drop_masks[s * np + i] : nullptr;
      if (drop_rng) a.drop = make_drop_spec(rng->seed, (unsigned int)(s * np + i), rng->input_dropout);
      a.y = i == 0 ? (s == 0 ? src_vid : src_txt) : (s == 0 ? T.p_vid32[i - 1] : T.p_txt32[i - 1]);
      if (i == 0 && P->in_fmt != 0) {
        a.y16 = reinterpret_cast<const uint16_t*>(a.y);
        a.y_fmt = P->in_fmt - 1;
      }
      a.ld_y = pp.din;
      a.mean = s == 0 ? T.pmean_v[i] : T.pmean_t[i];
      a.rstd = s == 0 ? T.prstd_v[i] : T.prstd_t[i];
      a.gamma = F32(pp.ln_w);
      a.rows = s == 0 ? Mv : Mt;
      a.d = pp.din;
      a.dgamma = s == 0 ? G_vid(i, 0) : G_txt(i, 0);
      a.dbeta = s == 0 ? G_vid(i, 1) : G_txt(i, 1);
      a.pgrad_scale = INV;
      if (i > 0) {
        a.relu_mask_y = 1;
        a.dbr16 = s == 0 ? T.dxv16 : T.dxt16;
        a.ld16 = d;
        a.fmt16 = FMT_G;
        a.colsum = s == 0 ? G_vid(i - 1, 3) : G_txt(i - 1, 3);
      }
      rc = launch_layernorm_bwd(a, st);
      if (rc) return rc;
    }
  }
  stage_done(c.enc_layers + 2);  // LayerNorm terms of the first projector layers
  prof_mark(P, st, 3);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("univtg_backward: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

int univtg_backward_stages(const univtg_config* cfg, int32_t* ranges, int32_t max_stages) {
  const int n_params = univtg_num_params(cfg);
  if (n_params < 0) return -1;
  const int np = cfg->n_input_proj, nl = cfg->enc_layers;
  const int n = nl + 3;
  if (ranges == nullptr) return n;
  if (max_stages < n) {
    set_error("univtg_backward_stages: need room for %d stages", n);
    return -1;
  }
  const int hb = 8 * np + 1 + 12 * nl;
  auto put = [&](int k, int a0, int a1, int b0, int b1) {
    ranges[4 * k + 0] = a0;
    ranges[4 * k + 1] = a1;
    ranges[4 * k + 2] = b0;
    ranges[4 * k + 3] = b1;
  };
  put(0, hb, hb + 13, 0, 0);  // span_embed + class_embed, weightedpool.weight
  for (int k = 0; k < nl; ++k) {
    const int l = nl - 1 - k;
    put(1 + k, 8 * np + 1 + 12 * l, 8 * np + 1 + 12 * (l + 1), 0, 0);
  }
  // projector parameters are [ln.weight, ln.bias, W, b] per layer, video layers first, then text, then token_type_embeddings
  put(nl + 1, 2, 4 * np, 4 * np + 2, 8 * np + 1);  // everything but the first layers' LayerNorm terms
  put(nl + 2, 0, 2, 4 * np, 4 * np + 2);           // those (final only after the last LayerNorm backward)
  return n;
}

int univtg_plan_set_backward_sm_budget(univtg_plan* plan, int32_t num_sms) {
  if (!plan || num_sms < 0) {
    set_error("univtg_plan_set_backward_sm_budget: bad argument");
    return 1;
  }
  plan->num_sms_bwd = num_sms;
  return 0;
}

int univtg_plan_set_grad_events(univtg_plan* plan, void* const* events, int32_t n) {
  if (!plan || n < 0 || n > 24 || (n > 0 && !events)) {
    set_error("univtg_plan_set_grad_events: bad argument");
    return 1;
  }
  if (n > 0 && n != plan->cfg.enc_layers + 3) {
    set_error("univtg_plan_set_grad_events: expected %d events (univtg_backward_stages), got %d", plan->cfg.enc_layers + 3, n);
    return 1;
  }
  plan->n_grad_events = n;
  for (int i = 0; i < n; ++i) plan->grad_events[i] = reinterpret_cast<cudaEvent_t>(events[i]);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// criterion (reference SetCriterion.forward, model/univtg.py:338-351, losses 'spans' + 'labels' + 'saliency')
// scratch: device buffer of univtg_loss_scratch_bytes(B, Lv); holds the per-loss gradients between forward and backward.
// ------------------------------------------------------------------------------------------------
namespace {
struct LossScratch {
  float *g_spans_b, *g_spans_g, *g_logits_f, *cos_in, *vnorm, *tnorm, *sim, *g_cos_in, *g_sim;
  size_t total;
};
LossScratch make_loss_scratch(int B, int Lv, uint8_t* base) {
  LossScratch s;
  Cursor cur;
  auto take32 = [&](size_t n) { return reinterpret_cast<float*>(base + cur.take(n * 4)); };
  const size_t n = (size_t)B * Lv;
  s.g_spans_b = take32(2 * n);
  s.g_spans_g = take32(2 * n);
  s.g_logits_f = take32(n);
  s.cos_in = take32(n);
  s.vnorm = take32(n);
  s.tnorm = take32(B);
  s.sim = take32((size_t)B * B);
  s.g_cos_in = take32(n);
  s.g_sim = take32((size_t)B * B);
  s.total = cur.off;
  return s;
}
}  // namespace

int univtg_op_attention_bwd(const void* qkv, const void* dO, const void* O, const float* key_mask, const float* lse,
                            float* delta_ws, float* dqkv32, int32_t B, int32_t L, int32_t H, int32_t dh, int32_t fmt_act,
                            int32_t impl, void* stream) {
  if (!qkv || !dO || !O || !key_mask || !lse || !delta_ws || !dqkv32) {
    set_error("univtg_op_attention_bwd: null argument");
    return 1;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int d = H * dh, M = B * L;
  int rc = launch_attn_delta(reinterpret_cast<const uint16_t*>(dO), fmt_act, reinterpret_cast<const uint16_t*>(O), fmt_act,
                             delta_ws, B, L, H, dh, st);
  if (rc) return rc;
  AttnBwdArgs a;
  memset(&a, 0, sizeof(a));
  a.qkv = reinterpret_cast<const uint16_t*>(qkv);
  a.dO = reinterpret_cast<const uint16_t*>(dO);
  a.key_mask = key_mask;
  a.lse = lse;
  a.delta = delta_ws;
  a.dqkv32 = dqkv32;
  a.scale = 1.0f / sqrtf((float)dh);
  a.B = B;
  a.L = L;
  a.H = H;
  a.dh = dh;
  a.d = d;
  a.fmt_act = fmt_act;
  a.fmt_grad = fmt_act;
  const bool tc = impl == 0;
  a.dq_atomic = (!tc || (L + 127) / 128 > 1) ? 1 : 0;
  if (a.dq_atomic) cudaMemsetAsync(dqkv32, 0, (size_t)M * 3 * d * 4, st);
  if (!tc) return launch_attention_bwd_simt(a, st);
  if (make_tmap_2d(&a.tm_qkv, qkv, (uint64_t)M, (uint64_t)3 * d, (uint64_t)3 * d, 128, 64)) return 1;
  if (make_tmap_2d(&a.tm_do, dO, (uint64_t)M, (uint64_t)d, (uint64_t)d, 128, 64)) return 1;
  return launch_attention_bwd(a, st);
}

int univtg_dropout_mask(const univtg_rng* rng, int32_t mask_index, size_t rows, size_t cols, float* out, void* stream) {
  if (!rng || !out || mask_index < 0 || cols == 0) {
    set_error("univtg_dropout_mask: bad argument");
    return 1;
  }
  return launch_dropout_mask(make_drop_spec(rng->seed, (unsigned int)mask_index, rng->input_dropout), rows * cols, cols, out,
                             (cudaStream_t)stream);
}

int univtg_droppath_scales(const univtg_rng* rng, int32_t n_sites, int32_t batch, float* out, void* stream) {
  if (!rng || !out || n_sites < 1 || batch < 1) {
    set_error("univtg_droppath_scales: bad argument");
    return 1;
  }
  return launch_droppath_scales(rng->seed, n_sites * batch, 1.0f - rng->droppath, out, (cudaStream_t)stream);
}

size_t univtg_loss_scratch_bytes(int32_t B, int32_t Lv) { return make_loss_scratch(B, Lv, nullptr).total; }

int univtg_loss_forward(const float* pred_logits, const float* pred_spans, const float* vid_mem_proj, const float* txt_mem_proj,
                        const float* timestamp, const float* timestamp_mask, const float* timestamp_window,
                        const float* span_labels_nn, const float* saliency_scores, const int64_t* saliency_pos_idx, int32_t B,
                        int32_t Lv, int32_t d, float eos_coef, float temperature, float* losses5, void* scratch, void* stream) {
  if (!pred_logits || !pred_spans || !vid_mem_proj || !txt_mem_proj || !timestamp_mask || !timestamp_window || !saliency_scores ||
      !losses5 || !scratch || ((timestamp == nullptr) != (span_labels_nn == nullptr))) {
    set_error("univtg_loss_forward: null argument");
    return 1;
  }
  if (B > 256) {
    set_error("univtg_loss_forward: batch %d > 256 not supported by the single-block reduction", B);
    return 1;
  }
  const LossScratch s = make_loss_scratch(B, Lv, reinterpret_cast<uint8_t*>(scratch));
  LossArgs a;
  a.pred_logits = pred_logits;
  a.pred_spans = pred_spans;
  a.xv = vid_mem_proj;
  a.xt = txt_mem_proj;
  a.timestamp = timestamp;
  a.tmask = timestamp_mask;
  a.window = timestamp_window;
  a.span_gt = span_labels_nn;
  a.sal = saliency_scores;
  a.pos_idx = saliency_pos_idx;
  a.eos_coef = eos_coef;
  a.temperature = temperature;
  a.B = B;
  a.Lv = Lv;
  a.d = d;
  a.losses = losses5;
  a.g_spans_b = s.g_spans_b;
  a.g_spans_g = s.g_spans_g;
  a.g_logits_f = s.g_logits_f;
  a.cos_in = s.cos_in;
  a.vnorm = s.vnorm;
  a.tnorm = s.tnorm;
  a.sim = s.sim;
  a.g_cos_in = s.g_cos_in;
  a.g_sim = s.g_sim;
  return launch_loss_forward(a, (cudaStream_t)stream);
}

int univtg_loss_backward(const float* w5, const float* vid_mem_proj, const float* txt_mem_proj, const int64_t* saliency_pos_idx,
                         int32_t B, int32_t Lv, int32_t d, const void* scratch, float* d_logits, float* d_spans,
                         float* d_vid_mem_proj, float* d_txt_mem_proj, void* stream) {
  if (!w5 || !vid_mem_proj || !txt_mem_proj || !scratch || !d_logits || !d_spans || !d_vid_mem_proj || !d_txt_mem_proj) {
    set_error("univtg_loss_backward: null argument");
    return 1;
  }
  const LossScratch s = make_loss_scratch(B, Lv, const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(scratch)));
  LossBwdArgs a;
  a.w = w5;
  a.g_spans_b = s.g_spans_b;
  a.g_spans_g = s.g_spans_g;
  a.g_logits_f = s.g_logits_f;
  a.cos_in = s.cos_in;
  a.vnorm = s.vnorm;
  a.tnorm = s.tnorm;
  a.sim = s.sim;
  a.g_cos_in = s.g_cos_in;
  a.g_sim = s.g_sim;
  a.xv = vid_mem_proj;
  a.xt = txt_mem_proj;
  a.pos_idx = saliency_pos_idx;
  a.B = B;
  a.Lv = Lv;
  a.d = d;
  a.d_logits = d_logits;
  a.d_spans = d_spans;
  a.d_xv = d_vid_mem_proj;
  a.d_xt = d_txt_mem_proj;
  return launch_loss_backward(a, (cudaStream_t)stream);
}

}  // extern "C"
