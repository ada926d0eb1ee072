// C-ABI layer (include/univtg_b200.h): weight packing, plan construction (tensor maps + launch descriptors)
// and the forward orchestration of reference Model.forward (model/univtg.py:105-155).
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <vector>

#include "plan.h"
#include "ptx.cuh"

extern "C" {

const char* univtg_last_error(void) { return uv::last_error(); }
int univtg_abi_version(void) { return UNIVTG_ABI_VERSION; }

int univtg_num_params(const univtg_config* cfg) {
  if (!check_cfg(cfg)) return -1;
  return 8 * cfg->n_input_proj + 1 + 12 * cfg->enc_layers + 12 + 1;
}

size_t univtg_packed_bytes(const univtg_config* cfg) {
  if (!check_cfg(cfg)) return 0;
  return make_layout(*cfg).total;
}

// mode 0: everything; mode 1: only the fp32 vectors / small fp32 tensors (the 16-bit matrices are kept current by univtg_adamw_step)
static int pack_impl(const univtg_config* cfg, const float* const* params, int32_t n_params, void* packed, void* stream, int mode) {
  if (!check_cfg(cfg)) return 1;
  const int expect = univtg_num_params(cfg);
  if (n_params != expect || !params || !packed) {
    set_error("univtg_pack_weights: expected %d parameter tensors, got %d", expect, n_params);
    return 1;
  }
  for (int i = 0; i < n_params; ++i)
    if (!params[i]) {
      set_error("univtg_pack_weights: parameter %d is null", i);
      return 1;
    }
  const PackedLayout L = make_layout(*cfg);
  const int d = cfg->hidden_dim, ff = cfg->dim_feedforward;
  Packer pk;
  pk.base = reinterpret_cast<uint8_t*>(packed);
  pk.fmt = cfg->operand_format;
  pk.st = (cudaStream_t)stream;
  pk.tab.n = 0;
  pk.skip_matrices = mode == 1;
  int idx = 0;
  const int type_idx = 8 * cfg->n_input_proj;  // token_type_embeddings.weight [2, d]
  const float* type_emb = params[type_idx];
  for (int s = 0; s < 2; ++s) {
    const ProjPacked* pp = s == 0 ? L.vid : L.txt;
    for (int i = 0; i < cfg->n_input_proj; ++i) {
      pk.vec(params[idx + 0], pp[i].ln_w, pp[i].din);
      pk.vec(params[idx + 1], pp[i].ln_b, pp[i].din);
      pk.rows(params[idx + 2], pp[i].w16, d, pp[i].din, pp[i].kpad);
      const bool last = (i == cfg->n_input_proj - 1);
      // token_type_embeddings: index 1 for video tokens, 0 for text tokens (model/univtg.py:114-115)
      pk.vec(params[idx + 3], pp[i].bias, d, last ? type_emb + (s == 0 ? d : 0) : nullptr);
      idx += 4;
    }
  }
  idx += 1;  // type embedding consumed above
  for (int l = 0; l < cfg->enc_layers; ++l) {
    const LayerPacked& lp = L.layer[l];
    pk.rows(params[idx + 0], lp.w_in, 3 * d, d, d);
    pk.vec(params[idx + 1], lp.b_in, 3 * d);
    pk.rows(params[idx + 2], lp.w_out, d, d, d);
    pk.vec(params[idx + 3], lp.b_out, d);
    pk.rows(params[idx + 4], lp.w1, ff, d, d);
    pk.vec(params[idx + 5], lp.b1, ff);
    pk.rows(params[idx + 6], lp.w2, d, ff, ff);
    pk.vec(params[idx + 7], lp.b2, d);
    pk.vec(params[idx + 8], lp.n1w, d);
    pk.vec(params[idx + 9], lp.n1b, d);
    pk.vec(params[idx + 10], lp.n2w, d);
    pk.vec(params[idx + 11], lp.n2b, d);
    idx += 12;
  }
  // span_embed.layers.{0,1,2}, class_embed.layers.{0,1,2}
  const float* const* sp = params + idx;
  const float* const* cl = params + idx + 6;
  // fused first conv: rows [0,d) = class_embed.layers.0, rows [d,2d) = span_embed.layers.0
  pk.conv(cl[0], L.conv1_w, d, d);
  pk.conv(sp[0], L.conv1_w + (size_t)d * 3 * d * 2, d, d);
  pk.vec(cl[1], L.conv1_b, d);
  pk.vec(sp[1], L.conv1_b + (size_t)d * 4, d);
  pk.conv(cl[2], L.conv2c_w, d, d);
  pk.vec(cl[3], L.conv2c_b, d);
  pk.conv(sp[2], L.conv2s_w, d, d);
  pk.vec(sp[3], L.conv2s_b, d);
  pk.conv_f32(cl[4], L.conv3c_w, 1, d);
  pk.vec(cl[5], L.conv3c_b, 1);
  pk.conv_f32(sp[4], L.conv3s_w, 2, d);
  pk.vec(sp[5], L.conv3s_b, 2);
  idx += 12;
  pk.vec(params[idx], L.pool_w, d);
  pk.flush();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("univtg_pack_weights: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

int univtg_pack_weights(const univtg_config* cfg, const float* const* params, int32_t n_params, void* packed, void* stream) {
  return pack_impl(cfg, params, n_params, packed, stream, 0);
}
int univtg_pack_vectors(const univtg_config* cfg, const float* const* params, int32_t n_params, void* packed, void* stream) {
  return pack_impl(cfg, params, n_params, packed, stream, 1);
}

// Segments of the flat parameter buffer (univtg_pack_weights order, every tensor padded to a multiple of 4 floats) that are GEMM
// weight matrices, with their place in `packed`.
static int make_pack_segments(const univtg_config& c, void* packed, PackSegTable& t) {
  const PackedLayout L = make_layout(c);
  uint8_t* base = reinterpret_cast<uint8_t*>(packed);
  const long long d = c.hidden_dim, ff = c.dim_feedforward;
  t.n = 0;
  t.fmt = c.operand_format;
  long long off = 0;  // floats
  auto skip = [&](long long numel) { off += (numel + 3) / 4 * 4; };
  auto seg = [&](long long numel, size_t dst, int kind, int rows, int cols, int ld) {
    if (t.n >= kMaxPackSegs) return 1;
    PackSeg& s = t.s[t.n++];
    s.start4 = off / 4;
    s.end4 = (off + numel + 3) / 4;
    s.dst = base + dst;
    s.kind = kind;
    s.rows = rows;
    s.cols = cols;
    s.ld = ld;
    skip(numel);
    return 0;
  };
  int rc = 0;
  for (int sdx = 0; sdx < 2; ++sdx) {
    const ProjPacked* pp = sdx == 0 ? L.vid : L.txt;
    for (int i = 0; i < c.n_input_proj; ++i) {
      skip(pp[i].din);
      skip(pp[i].din);
      rc |= seg(d * pp[i].din, pp[i].w16, 0, (int)d, pp[i].din, pp[i].kpad);
      skip(d);
    }
  }
  skip(2 * d);  // token_type_embeddings
  for (int l = 0; l < c.enc_layers; ++l) {
    const LayerPacked& lp = L.layer[l];
    rc |= seg(3 * d * d, lp.w_in, 0, (int)(3 * d), (int)d, (int)d);
    skip(3 * d);
    rc |= seg(d * d, lp.w_out, 0, (int)d, (int)d, (int)d);
    skip(d);
    rc |= seg(ff * d, lp.w1, 0, (int)ff, (int)d, (int)d);
    skip(ff);
    rc |= seg(d * ff, lp.w2, 0, (int)d, (int)ff, (int)ff);
    skip(d);
    skip(d); skip(d); skip(d); skip(d);
  }
  // span_embed.layers.{0,1,2} then class_embed.layers.{0,1,2}; fused first conv: rows [0,d) class, [d,2d) span
  rc |= seg(d * d * 3, L.conv1_w + (size_t)d * 3 * d * 2, 1, (int)d, (int)d, 0);
  skip(d);
  rc |= seg(d * d * 3, L.conv2s_w, 1, (int)d, (int)d, 0);
  skip(d);
  skip(2 * d * 3);
  skip(2);
  rc |= seg(d * d * 3, L.conv1_w, 1, (int)d, (int)d, 0);
  skip(d);
  rc |= seg(d * d * 3, L.conv2c_w, 1, (int)d, (int)d, 0);
  skip(d);
  skip(1 * d * 3);
  skip(1);
  skip(d);  // weightedpool.weight
  if (rc) {
    set_error("univtg_adamw_step: too many weight matrices for the pack-segment table");
    return -1;
  }
  return (int)(off);  // total floats of the flat buffer
}

int univtg_adamw_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, size_t n, float lr, float beta1, float beta2,
                      float eps, float weight_decay, int32_t step, float max_grad_norm, int32_t write_clipped_grads,
                      float* scratch3, const univtg_config* cfg, void* packed, void* stream) {
  if (cfg == nullptr || packed == nullptr)
    return uv::adamw_step_impl(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, max_grad_norm,
                               write_clipped_grads, scratch3, nullptr, stream);
  if (!check_cfg(cfg)) return 1;
  PackSegTable t;
  const int total = make_pack_segments(*cfg, packed, t);
  if (total < 0) return 1;
  if ((size_t)total != n) {
    set_error("univtg_adamw_step: flat buffer has %zu floats, the config's parameters need %d", n, total);
    return 1;
  }
  return uv::adamw_step_impl(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, max_grad_norm,
                             write_clipped_grads, scratch3, &t, stream);
}

}  // extern "C"

extern "C" {

size_t univtg_workspace_bytes(const univtg_config* cfg, const univtg_shape* shape) {
  if (!check_cfg(cfg) || !check_shape(shape)) return 0;
  return make_ws(*cfg, *shape, make_layout(*cfg)).total;
}

int univtg_plan_create(const univtg_config* cfg, const univtg_shape* shape, const void* packed, void* workspace,
                       const float* dim_t, void* stream, univtg_plan** out) {
  if (!check_cfg(cfg) || !check_shape(shape)) return 1;
  if (!packed || !workspace || !dim_t || !out) {
    set_error("univtg_plan_create: null argument");
    return 1;
  }
  univtg_plan* P = new (std::nothrow) univtg_plan;
  if (!P) {
    set_error("out of host memory");
    return 1;
  }
  memset(static_cast<void*>(P), 0, sizeof(*P));
  P->cfg = *cfg;
  P->shp = *shape;
  P->lay = make_layout(*cfg);
  P->packed = reinterpret_cast<const uint8_t*>(packed);
  P->ws = reinterpret_cast<uint8_t*>(workspace);
  P->dim_t = dim_t;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&P->num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (P->num_sms <= 0) P->num_sms = 148;
  const int d = cfg->hidden_dim, ff = cfg->dim_feedforward;
  P->B = shape->batch;
  P->Lv = shape->l_vid;
  P->Lt = shape->l_txt;
  P->L = P->Lv + P->Lt;
  P->d = d;
  P->ff = ff;
  P->H = cfg->nheads;
  P->dh = d / cfg->nheads;
  P->M = P->B * P->L;
  P->Mv = P->B * P->Lv;
  P->Mt = P->B * P->Lt;
  P->Mh = P->B * (P->Lv + 1);
  const WsLayout w = make_ws(*cfg, *shape, P->lay);
  if (univtg_prepare_workspace(cfg, shape, workspace, 0, stream) != 0) {  // zero rows of the conv-head buffers
    delete P;
    return 1;
  }
  uint8_t* ws = P->ws;
  for (int i = 0; i < cfg->n_input_proj; ++i) {
    P->a_vid[i] = reinterpret_cast<uint16_t*>(ws + w.a_vid[i]);
    P->a_txt[i] = reinterpret_cast<uint16_t*>(ws + w.a_txt[i]);
  }
  P->p_vid32 = reinterpret_cast<float*>(ws + w.p_vid32);
  P->p_txt32 = reinterpret_cast<float*>(ws + w.p_txt32);
  P->txtproj32 = reinterpret_cast<float*>(ws + w.txtproj32);
  P->pos = reinterpret_cast<float*>(ws + w.pos);
  P->key_mask = reinterpret_cast<float*>(ws + w.key_mask);
  P->pool_logits = reinterpret_cast<float*>(ws + w.pool_logits);
  P->x32 = reinterpret_cast<float*>(ws + w.x32);
  P->br16 = reinterpret_cast<uint16_t*>(ws + w.br16);
  P->x16 = reinterpret_cast<uint16_t*>(ws + w.x16);
  P->xpos16 = reinterpret_cast<uint16_t*>(ws + w.xpos16);
  P->qkv16 = reinterpret_cast<uint16_t*>(ws + w.qkv16);
  P->attn16 = reinterpret_cast<uint16_t*>(ws + w.attn16);
  P->h16 = reinterpret_cast<uint16_t*>(ws + w.h16);
  P->hA = reinterpret_cast<uint16_t*>(ws + w.hA);
  P->h1 = reinterpret_cast<uint16_t*>(ws + w.h1);
  P->hc2 = reinterpret_cast<uint16_t*>(ws + w.hc2);
  P->hs2 = reinterpret_cast<uint16_t*>(ws + w.hs2);

  const PackedLayout& Lw = P->lay;
  const uint8_t* pk = P->packed;
  auto W16 = [&](size_t off) { return reinterpret_cast<const uint16_t*>(pk + off); };
  auto F32 = [&](size_t off) { return reinterpret_cast<const float*>(pk + off); };
  const int fmt = cfg->operand_format;
  int rc = 0;
  // tile width: 256 when the N extent has at least one full 256 tile, else 128
  P->bn_main = (d % 256 == 0) ? 256 : 128;
  {
    // per-launch tile widths: fill the SMs with as little wave quantisation as possible
    const int sms = P->num_sms;
    auto bn2 = [&](int M0, int N0, int K0, int M1, int N1, int K1) {  // K in elements
      const int Ms[2] = {M0, M1}, Ns[2] = {N0, N1}, kb[2] = {(K0 + 63) / 64, (K1 + 63) / 64};
      return choose_bn(Ms, Ns, kb, M1 > 0 ? 2 : 1, sms, 16);
    };
    for (int i = 0; i < cfg->n_input_proj; ++i) P->bn_proj[i] = bn2(P->Mv, d, Lw.vid[i].kpad, P->Mt, d, Lw.txt[i].kpad);
    P->bn_qkv = bn2(P->M, 2 * d, d, P->M, d, d);
    P->bn_out = bn2(P->M, d, d, 0, 0, 0);
    P->bn_ffn1 = bn2(P->M, ff, d, 0, 0, 0);
    P->bn_ffn2 = bn2(P->M, d, ff, 0, 0, 0);
    P->bn_conv1 = bn2(P->Mh, 2 * d, 3 * d, 0, 0, 0);
    P->bn_conv2 = bn2(P->Mh, d, 3 * d, P->Mh, d, 3 * d);
  }

  // ---- input projectors: one grouped launch per projector depth (video + text problems) ----
  for (int i = 0; i < cfg->n_input_proj && !rc; ++i) {
    GemmGroup& g = P->g_proj[i];
    memset(&g, 0, sizeof(g));
    g.num = 2;
    g.fmt = fmt;
    const bool last = (i == cfg->n_input_proj - 1);
    const int bn = P->bn_proj[i];
    GemmProblem& pv = g.p[0];
    GemmProblem& pt = g.p[1];
    rc |= setup_linear(pv, P->a_vid[i], P->Mv, Lw.vid[i].kpad, Lw.vid[i].kpad, W16(Lw.vid[i].w16), d, Lw.vid[i].kpad, bn);
    rc |= setup_linear(pt, P->a_txt[i], P->Mt, Lw.txt[i].kpad, Lw.txt[i].kpad, W16(Lw.txt[i].w16), d, Lw.txt[i].kpad, bn);
    pv.bias = F32(Lw.vid[i].bias);
    pt.bias = F32(Lw.txt[i].bias);
    if (!last) {
      pv.act = pt.act = ACT_RELU;
      pv.out32 = P->p_vid32;
      pt.out32 = P->p_txt32;
      pv.ld32 = pt.ld32 = d;
    } else {
      // video tokens -> stream rows b*L + l; text tokens -> rows b*L + Lv + l   (cat on the sequence axis, univtg.py:119)
      pv.rps_in = P->Lv;
      pv.rps_out = P->L;
      pv.row_off = 0;
      pt.rps_in = P->Lt;
      pt.rps_out = P->L;
      pt.row_off = P->Lv;
      pv.out32 = pt.out32 = P->x32;
      pv.ld32 = pt.ld32 = d;
      pv.out16 = pt.out16 = P->x16;
      pv.out16p = pt.out16p = P->xpos16;
      pv.ld16 = pt.ld16 = d;
      pv.addtab = P->pos;
      pv.ld_addtab = d;
      pv.out32_id = nullptr;  // vid_mem_proj: set per call
      pv.ld32_id = d;
      pt.out32_id = P->txtproj32;
      pt.ld32_id = d;
    }
  }
  // ---- encoder layers ----
  const float qscale = 1.0f / sqrtf((float)P->dh);
  for (int l = 0; l < cfg->enc_layers && !rc; ++l) {
    const LayerPacked& lp = Lw.layer[l];
    {
      GemmGroup& g = P->g_qkv[l];
      memset(&g, 0, sizeof(g));
      g.num = 2;
      g.fmt = fmt;
      // q = k = x + pos -> columns [0, 2d) of qkv16; v = x -> columns [2d, 3d)   (in_proj rows: Wq, Wk, Wv)
      rc |= setup_linear(g.p[0], P->xpos16, P->M, d, d, W16(lp.w_in), 2 * d, d, P->bn_qkv);
      rc |= setup_linear(g.p[1], P->x16, P->M, d, d, W16(lp.w_in) + (size_t)2 * d * d, d, d, P->bn_qkv);
      g.p[0].bias = F32(lp.b_in);
      g.p[0].out16 = P->qkv16;
      g.p[0].ld16 = 3 * d;
      g.p[1].bias = F32(lp.b_in) + 2 * d;
      g.p[1].out16 = P->qkv16 + 2 * d;
      g.p[1].ld16 = 3 * d;
    }
    {
      AttnArgs& a = P->attn[l];
      memset(&a, 0, sizeof(a));
      a.key_mask = P->key_mask;
      a.out = P->attn16;
      a.lse = nullptr;
      a.scale = qscale;  // torch MHA: q * dh**-0.5 before q k^T
      a.B = P->B;
      a.L = P->L;
      a.H = P->H;
      a.dh = P->dh;
      a.d = d;
      a.fmt = fmt;
      if (P->dh == 64 || P->dh == 128)
        rc |= make_tmap_2d(&a.tm_qkv, P->qkv16, (uint64_t)P->M, (uint64_t)3 * d, (uint64_t)3 * d, 128, 64);
    }
    {
      GemmGroup& g = P->g_out[l];
      memset(&g, 0, sizeof(g));
      g.num = 1;
      g.fmt = fmt;
      rc |= setup_linear(g.p[0], P->attn16, P->M, d, d, W16(lp.w_out), d, d, P->bn_out);
      g.p[0].bias = F32(lp.b_out);
      g.p[0].rps_in = P->L;
      g.p[0].rps_out = P->L;
      g.p[0].out16 = P->br16;  // DropPath-scaled branch; the LayerNorm kernel adds it to the fp32 residual stream
      g.p[0].ld16 = d;
    }
    {
      GemmGroup& g = P->g_ffn1[l];
      memset(&g, 0, sizeof(g));
      g.num = 1;
      g.fmt = fmt;
      rc |= setup_linear(g.p[0], P->x16, P->M, d, d, W16(lp.w1), ff, d, P->bn_ffn1);
      g.p[0].bias = F32(lp.b1);
      g.p[0].act = ACT_GELU;
      g.p[0].out16 = P->h16;
      g.p[0].ld16 = ff;
    }
    {
      GemmGroup& g = P->g_ffn2[l];
      memset(&g, 0, sizeof(g));
      g.num = 1;
      g.fmt = fmt;
      rc |= setup_linear(g.p[0], P->h16, P->M, ff, ff, W16(lp.w2), d, ff, P->bn_ffn2);
      g.p[0].bias = F32(lp.b2);
      g.p[0].rps_in = P->L;
      g.p[0].rps_out = P->L;
      g.p[0].out16 = P->br16;
      g.p[0].ld16 = d;
    }
  }
  // ---- conv heads (k=3, pad=1) as 3-tap GEMMs over the separated layout ----
  if (!rc) {
    auto conv_problem = [&](GemmProblem& p, const uint16_t* A, int lda, const uint16_t* W, int N, const float* bias,
                            uint16_t* out, int ldo, int bn) -> int {
      init_problem(p);
      p.M = P->Mh;
      p.N = N;
      p.taps = 3;
      p.kblk_per_tap = d / 64;
      // A tile row for tap t: buffer row m0 + t  (buffer row = logical row + 1)
      p.ca = OperandCoord{0, 0, 0, 1, 0, 1, 1, 0};
      // W2 [N, 3d]: column tap*d + k
      p.cb = OperandCoord{0, 0, d, 1, 0, 1, 0, 0};
      int r = make_tmap_2d(&p.tm_a, A, (uint64_t)P->Mh + 2, (uint64_t)d, (uint64_t)lda, GEMM_BM, 64);
      r |= make_tmap_2d(&p.tm_b, W, (uint64_t)N, (uint64_t)3 * d, (uint64_t)3 * d, (uint32_t)bn, 64);
      p.b_box_rows = bn;
      p.bias = bias;
      p.act = ACT_RELU;
      p.rps_in = P->Lv + 1;
      p.rps_out = P->Lv + 1;
      p.row_off = 1;
      p.zero_sep = 1;
      p.out16 = out;
      p.ld16 = ldo;
      return r;
    };
    memset(&P->g_conv1, 0, sizeof(GemmGroup));
    P->g_conv1.num = 1;
    P->g_conv1.fmt = fmt;
    rc |= conv_problem(P->g_conv1.p[0], P->hA, d, W16(Lw.conv1_w), 2 * d, F32(Lw.conv1_b), P->h1, 2 * d, P->bn_conv1);
    memset(&P->g_conv2, 0, sizeof(GemmGroup));
    P->g_conv2.num = 2;
    P->g_conv2.fmt = fmt;
    rc |= conv_problem(P->g_conv2.p[0], P->h1, 2 * d, W16(Lw.conv2c_w), d, F32(Lw.conv2c_b), P->hc2, d, P->bn_conv2);
    rc |= conv_problem(P->g_conv2.p[1], P->h1 + d, 2 * d, W16(Lw.conv2s_w), d, F32(Lw.conv2s_b), P->hs2, d, P->bn_conv2);
  }
  if (rc) {
    delete P;
    return 1;
  }
  P->launches = 1 + 3 * cfg->n_input_proj + 7 * cfg->enc_layers + 6;
  *out = P;
  return 0;
}

void univtg_plan_destroy(univtg_plan* plan) {
  if (!plan) return;
  for (int i = 0; i < kMaxMarks; ++i)
    if (plan->marks[i]) cudaEventDestroy(plan->marks[i]);
  delete plan;
}

int univtg_plan_set_input_format(univtg_plan* plan, int32_t fmt) {
  if (!plan || fmt < 0 || fmt > 2) {
    set_error("univtg_plan_set_input_format: format must be 0 (f32), 1 (fp16) or 2 (bf16)");
    return 1;
  }
  plan->in_fmt = fmt;
  return 0;
}

int univtg_plan_set_profiling(univtg_plan* plan, int32_t enable) {
  if (!plan) return 1;
  plan->profiling = enable ? 1 : 0;
  plan->n_marks = 0;
  return 0;
}

int univtg_plan_read_profile(univtg_plan* plan, float* ms, int32_t* kinds, int32_t cap) {
  if (!plan || !ms || !kinds) return -1;
  if (plan->n_marks < 2) return 0;
  cudaError_t e = cudaEventSynchronize(plan->marks[plan->n_marks - 1]);
  if (e != cudaSuccess) {
    set_error("profile sync: %s", cudaGetErrorString(e));
    return -1;
  }
  int n = 0;
  for (int i = 1; i < plan->n_marks && n < cap; ++i, ++n) {
    cudaEventElapsedTime(&ms[n], plan->marks[i - 1], plan->marks[i]);
    kinds[n] = plan->mark_kind[i];
  }
  return n;
}

int univtg_forward_num_launches(const univtg_plan* plan) { return plan ? plan->launches : -1; }
int64_t univtg_launch_count(void) { return (int64_t)*uv::launch_counter(); }

int univtg_forward(univtg_plan* P, const float* src_txt, const float* src_txt_mask, const float* src_vid,
                   const float* src_vid_mask, const float* droppath_scale, float* pred_logits, float* pred_spans,
                   float* vid_mem_proj, float* txt_mem_proj, float* saliency_scores, void* stream) {
  if (!P || !src_txt || !src_txt_mask || !src_vid || !src_vid_mask || !pred_logits || !pred_spans || !vid_mem_proj ||
      !txt_mem_proj || !saliency_scores) {
    set_error("univtg_forward: null argument");
    return 1;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const univtg_config& c = P->cfg;
  const PackedLayout& Lw = P->lay;
  const uint8_t* pk = P->packed;
  auto F32 = [&](size_t off) { return reinterpret_cast<const float*>(pk + off); };
  const int d = P->d, fmt = c.operand_format;
  int rc = 0;
  prof_begin(P, st);

  rc = launch_sine_pos(src_vid_mask, src_txt_mask, P->dim_t, P->pos, P->key_mask, P->B, P->Lv, P->Lt, d, st);
  if (rc) return rc;
  prof_mark(P, st, 0);

  // ---- input projectors (LinearLayer: LN -> Dropout(eval: identity) -> Linear -> ReLU) ----
  for (int i = 0; i < c.n_input_proj; ++i) {
    for (int s = 0; s < 2; ++s) {
      const ProjPacked& pp = s == 0 ? Lw.vid[i] : Lw.txt[i];
      LnArgs a;
      memset(&a, 0, sizeof(a));
      a.in = i == 0 ? (s == 0 ? src_vid : src_txt) : (s == 0 ? P->p_vid32 : P->p_txt32);
      if (i == 0 && P->in_fmt != 0) {
        a.in16 = reinterpret_cast<const uint16_t*>(a.in);
        a.in_fmt = P->in_fmt - 1;
      }
      a.ld_in = pp.din;
      a.rows = s == 0 ? P->Mv : P->Mt;
      a.d = pp.din;
      a.gamma = F32(pp.ln_w);
      a.beta = F32(pp.ln_b);
      a.eps = 1e-5f;
      a.fmt = fmt;
      a.out16 = s == 0 ? P->a_vid[i] : P->a_txt[i];
      a.ld16 = pp.kpad;
      rc = launch_layernorm(a, st);
      if (rc) return rc;
      prof_mark(P, st, 0);
    }
    GemmGroup g = P->g_proj[i];
    if (i == c.n_input_proj - 1) g.p[0].out32_id = vid_mem_proj;
    rc = launch_gemm_group(g, P->bn_proj[i], P->num_sms, st);
    if (rc) return rc;
    prof_mark(P, st, 1);
  }

  // ---- encoder layers (post-norm; TransformerEncoderLayer.forward_post) ----
  for (int l = 0; l < c.enc_layers; ++l) {
    const LayerPacked& lp = Lw.layer[l];
    rc = launch_gemm_group(P->g_qkv[l], P->bn_qkv, P->num_sms, st);
    if (rc) return rc;
    prof_mark(P, st, 1);
    if (P->dh == 64 || P->dh == 128) rc = launch_attention(P->attn[l], st);
    else rc = launch_attention_simt(P->attn[l], P->qkv16, st);
    if (rc) return rc;
    prof_mark(P, st, 2);
    {
      GemmGroup g = P->g_out[l];
      g.p[0].row_scale = droppath_scale ? droppath_scale + (size_t)(2 * l) * P->B : nullptr;
      rc = launch_gemm_group(g, P->bn_out, P->num_sms, st);
      if (rc) return rc;
      prof_mark(P, st, 1);
    }
    {
      LnArgs a;
      memset(&a, 0, sizeof(a));
      a.in = P->x32;
      a.ld_in = d;
      a.add16 = P->br16;
      a.ld_add16 = d;
      a.rows = P->M;
      a.d = d;
      a.gamma = F32(lp.n1w);
      a.beta = F32(lp.n1b);
      a.eps = 1e-5f;
      a.fmt = fmt;
      a.out32 = P->x32;
      a.out16 = P->x16;
      a.ld16 = d;
      rc = launch_layernorm(a, st);
      if (rc) return rc;
      prof_mark(P, st, 0);
    }
    rc = launch_gemm_group(P->g_ffn1[l], P->bn_ffn1, P->num_sms, st);
    if (rc) return rc;
    prof_mark(P, st, 1);
    {
      GemmGroup g = P->g_ffn2[l];
      g.p[0].row_scale = droppath_scale ? droppath_scale + (size_t)(2 * l + 1) * P->B : nullptr;
      rc = launch_gemm_group(g, P->bn_ffn2, P->num_sms, st);
      if (rc) return rc;
      prof_mark(P, st, 1);
    }
    {
      LnArgs a;
      memset(&a, 0, sizeof(a));
      a.in = P->x32;
      a.ld_in = d;
      a.add16 = P->br16;
      a.ld_add16 = d;
      a.rows = P->M;
      a.d = d;
      a.gamma = F32(lp.n2w);
      a.beta = F32(lp.n2b);
      a.eps = 1e-5f;
      a.fmt = fmt;
      a.L = P->L;
      a.Lv = P->Lv;
      a.out32 = P->x32;
      a.out16 = P->x16;
      a.out16p = P->xpos16;
      a.ld16 = d;
      a.pos = P->pos;
      if (l == c.enc_layers - 1) a.outc = P->hA;  // vid_mem = memory[:, :Lv] feeds the conv heads
      rc = launch_layernorm(a, st);
      if (rc) return rc;
      prof_mark(P, st, 0);
    }
  }

  // ---- heads ----
  rc = launch_gemm_group(P->g_conv1, P->bn_conv1, P->num_sms, st);
  if (rc) return rc;
  prof_mark(P, st, 1);
  rc = launch_gemm_group(P->g_conv2, P->bn_conv2, P->num_sms, st);
  if (rc) return rc;
  prof_mark(P, st, 1);
  {
    HeadFinalArgs a;
    a.h_cls = P->hc2;
    a.h_span = P->hs2;
    a.w_cls = F32(Lw.conv3c_w);
    a.w_span = F32(Lw.conv3s_w);
    a.b_cls = F32(Lw.conv3c_b);
    a.b_span = F32(Lw.conv3s_b);
    a.pred_logits = pred_logits;
    a.pred_spans = pred_spans;
    a.B = P->B;
    a.Lv = P->Lv;
    a.d = d;
    a.fmt = fmt;
    rc = launch_conv_head_final(a, st);
    if (rc) return rc;
    prof_mark(P, st, 0);
  }
  {
    PoolSalArgs a;
    a.x_txt = P->txtproj32;
    a.x_vid = vid_mem_proj;
    a.txt_mask = src_txt_mask;
    a.vid_mask = src_vid_mask;
    a.w = F32(Lw.pool_w);
    a.pooled = txt_mem_proj;
    a.saliency = saliency_scores;
    a.alpha_out = nullptr;
    a.logits_ws = P->pool_logits;
    a.B = P->B;
    a.Lt = P->Lt;
    a.Lv = P->Lv;
    a.d = d;
    rc = launch_pool_saliency(a, st);
    if (rc) return rc;
    prof_mark(P, st, 0);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// single operators
// ------------------------------------------------------------------------------------------------
static int op_gemm_impl(const void* a, const void* b, int32_t M, int32_t N, int32_t K, int32_t a_mn, int32_t b_mn, int32_t fmt,
                        int32_t bn, int32_t ksplit, const float* bias, int32_t act, float alpha, float* out32, void* out16,
                        int32_t cluster, void* stream) {
  if (!a || !b || M < 1 || N < 1 || K < 1 || bn < 32 || bn > 256 || bn % 16 != 0) {
    set_error("univtg_op_gemm: bad argument");
    return 1;
  }
  GemmGroup g;
  memset(&g, 0, sizeof(g));
  g.num = 1;
  g.fmt = fmt;
  g.cluster = cluster;
  GemmProblem& p = g.p[0];
  init_problem(p);
  p.M = M;
  p.N = N;
  p.a_mn = a_mn;
  p.b_mn = b_mn;
  p.kblk_per_tap = (K + 63) / 64;
  p.ksplit = ksplit < 1 ? 1 : ksplit;
  int rc = 0;
  if (!a_mn) {
    rc |= make_tmap_2d(&p.tm_a, a, (uint64_t)M, (uint64_t)K, (uint64_t)K, GEMM_BM, 64);
  } else {
    rc |= make_tmap_2d(&p.tm_a, a, (uint64_t)K, (uint64_t)M, (uint64_t)M, 64, 64);
    p.ca = OperandCoord{0, 1, 0, 0, 0, 0, 0, 1};  // c0 = m0, c1 = k
  }
  if (!b_mn) {
    rc |= make_tmap_2d(&p.tm_b, b, (uint64_t)N, (uint64_t)K, (uint64_t)K, (uint32_t)(cluster == 2 ? bn / 2 : bn), 64);
    p.b_box_rows = cluster == 2 ? bn / 2 : bn;
  } else {
    rc |= make_tmap_b_mn(p, b, (uint64_t)K, (uint64_t)N, (uint64_t)N, bn, cluster != 2);
    p.cb = OperandCoord{0, 1, 0, 0, 0, 0, 0, 1};
  }
  if (rc) return rc;
  p.bias = bias;
  p.act = act;
  p.alpha = alpha;
  p.out32 = out32;
  p.ld32 = N;
  p.out16 = reinterpret_cast<uint16_t*>(out16);
  p.ld16 = N;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return launch_gemm_group(g, bn, sms, (cudaStream_t)stream);
}

int univtg_op_gemm(const void* a, const void* b, int32_t M, int32_t N, int32_t K, int32_t a_mn, int32_t b_mn, int32_t fmt,
                   int32_t bn, int32_t ksplit, const float* bias, int32_t act, float alpha, float* out32, void* out16,
                   void* stream) {
  return op_gemm_impl(a, b, M, N, K, a_mn, b_mn, fmt, bn, ksplit, bias, act, alpha, out32, out16, 1, stream);
}

int univtg_op_gemm_cluster(const void* a, const void* b, int32_t M, int32_t N, int32_t K, int32_t a_mn, int32_t b_mn,
                           int32_t fmt, int32_t bn, int32_t ksplit, const float* bias, int32_t act, float alpha, float* out32,
                           void* out16, void* stream) {
  return op_gemm_impl(a, b, M, N, K, a_mn, b_mn, fmt, bn, ksplit, bias, act, alpha, out32, out16, 2, stream);
}

int univtg_debug_mma_rate(int32_t n, int32_t iters, int32_t per_commit, int32_t kstep_bytes, int32_t blocks, float* out_ns,
                          void* stream) {
  // kstep_bytes: bit 30 selects an MN-major A operand, bit 29 an MN-major B operand (probe-only encoding)
  return uv::debug_mma_rate(n, iters, per_commit, kstep_bytes & 0xffff, blocks, out_ns, reinterpret_cast<cudaStream_t>(stream),
                            (kstep_bytes >> 30) & 1, (kstep_bytes >> 29) & 1);
}

int univtg_debug_tmem_ld_rate(int32_t iters, int32_t mode, int32_t blocks, float* out_ns, float* sink, void* stream) {
  return uv::debug_tmem_ld_rate(iters, mode, blocks, out_ns, sink, reinterpret_cast<cudaStream_t>(stream));
}

int univtg_debug_choose_tile(const int32_t* Ms, const int32_t* Ns, const int32_t* kblocks, int32_t num, int32_t num_sms, int32_t step,
                              int32_t max_split, int32_t* bn, int32_t* ksplit) {
  if (!Ms || !Ns || !kblocks || !bn || !ksplit || num < 1 || num > GEMM_MAX_GROUP || num_sms < 1 || (step != 16 && step != 64) || max_split < 1) {
    set_error("univtg_debug_choose_tile: bad argument");
    return 1;
  }
  const uv::TileChoice t = uv::choose_tile(Ms, Ns, kblocks, num, num_sms, step, max_split);
  *bn = t.bn;
  *ksplit = t.ksplit;
  return 0;
}

int univtg_debug_gemm_timeline(void* buf) {
  uv::set_gemm_timeline_buffer(reinterpret_cast<unsigned long long*>(buf));
  return 0;
}

int univtg_op_layernorm(const float* in, int32_t rows, int32_t d, const float* gamma, const float* beta, float eps,
                        int32_t fmt, float* out32, void* out16, int32_t ld16, void* stream) {
  if (!in || !gamma || !beta || rows < 1 || d < 1) {
    set_error("univtg_op_layernorm: bad argument");
    return 1;
  }
  LnArgs a;
  memset(&a, 0, sizeof(a));
  a.in = in;
  a.ld_in = d;
  a.rows = rows;
  a.d = d;
  a.gamma = gamma;
  a.beta = beta;
  a.eps = eps;
  a.fmt = fmt;
  a.out32 = out32;
  a.out16 = reinterpret_cast<uint16_t*>(out16);
  a.ld16 = out16 ? ld16 : d;
  return launch_layernorm(a, (cudaStream_t)stream);
}

int univtg_op_attention(const void* qkv, const float* key_mask, void* out, float* lse, int32_t B, int32_t L, int32_t H,
                        int32_t dh, int32_t fmt, int32_t impl, void* stream) {
  if (!qkv || !key_mask || !out) {
    set_error("univtg_op_attention: null argument");
    return 1;
  }
  AttnArgs a;
  memset(&a, 0, sizeof(a));
  const int d = H * dh;
  a.key_mask = key_mask;
  a.out = reinterpret_cast<uint16_t*>(out);
  a.lse = lse;
  a.scale = 1.0f / sqrtf((float)dh);
  a.B = B;
  a.L = L;
  a.H = H;
  a.dh = dh;
  a.d = d;
  a.fmt = fmt;
  if (impl == 1) return launch_attention_simt(a, reinterpret_cast<const uint16_t*>(qkv), (cudaStream_t)stream);
  if (make_tmap_2d(&a.tm_qkv, qkv, (uint64_t)B * L, (uint64_t)3 * d, (uint64_t)3 * d, 128, 64)) return 1;
  return launch_attention(a, (cudaStream_t)stream);
}

}  // extern "C"
