// Shared between api.cu (inference orchestration) and train.cu (training forward/backward): packed-weight layout,
// workspace layout and the plan object behind the opaque univtg_plan handle.
#pragma once
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "../../include/univtg_b200.h"
#include "backward.h"
#include "kernels.h"
#include "loss.h"
#include "ptx.cuh"
#include "rowops.h"

using namespace uv;

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int pad64(int k) { return (k + 63) / 64 * 64; }

// ------------------------------------------------------------------------------------------------
// packed-weight layout
// ------------------------------------------------------------------------------------------------
struct ProjPacked {
  size_t ln_w, ln_b;  // fp32 [din]
  size_t w16;         // 16-bit [d, kpad]
  size_t bias;        // fp32 [d]  (last layer: linear bias + token-type embedding row)
  int din, kpad;
};
struct LayerPacked {
  size_t w_in;   // 16-bit [3d, d]  (rows: Wq, Wk, Wv)
  size_t b_in;   // fp32 [3d]
  size_t w_out;  // 16-bit [d, d]
  size_t b_out;
  size_t w1, b1;  // [ff, d], [ff]
  size_t w2, b2;  // [d, ff], [d]
  size_t n1w, n1b, n2w, n2b;
};
struct PackedLayout {
  ProjPacked vid[3], txt[3];
  LayerPacked layer[16];
  size_t conv1_w, conv1_b;                     // fused first conv of both heads: 16-bit [2d, 3d] (rows: class, span), fp32 [2d]
  size_t conv2c_w, conv2c_b, conv2s_w, conv2s_b;  // 16-bit [d, 3d], fp32 [d]
  size_t conv3c_w, conv3c_b, conv3s_w, conv3s_b;  // fp32 [3][d], [1], [2][3][d], [2]
  size_t pool_w;                                  // fp32 [d]
  size_t total;
};

struct Cursor {
  size_t off = 0;
  size_t take(size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  }
};

bool check_cfg(const univtg_config* c) {
  if (!c) {
    set_error("null config");
    return false;
  }
  if (c->hidden_dim <= 0 || c->hidden_dim % 64 != 0) {
    set_error("hidden_dim %d must be a positive multiple of 64", c->hidden_dim);
    return false;
  }
  if (c->nheads <= 0 || c->hidden_dim % c->nheads != 0) {
    set_error("nheads %d must divide hidden_dim %d", c->nheads, c->hidden_dim);
    return false;
  }
  if (c->dim_feedforward <= 0 || c->dim_feedforward % 64 != 0) {
    set_error("dim_feedforward %d must be a positive multiple of 64", c->dim_feedforward);
    return false;
  }
  if (c->enc_layers < 1 || c->enc_layers > 16) {
    set_error("enc_layers %d out of range [1,16]", c->enc_layers);
    return false;
  }
  if (c->n_input_proj < 1 || c->n_input_proj > 3) {
    set_error("n_input_proj %d out of range [1,3]", c->n_input_proj);
    return false;
  }
  if (c->v_feat_dim <= 0 || c->t_feat_dim <= 0) {
    set_error("feature dims must be positive");
    return false;
  }
  if (c->operand_format != 0 && c->operand_format != 1) {
    set_error("operand_format %d must be 0 (fp16) or 1 (bf16)", c->operand_format);
    return false;
  }
  return true;
}

PackedLayout make_layout(const univtg_config& c) {
  PackedLayout L;
  memset(&L, 0, sizeof(L));
  Cursor cur;
  const int d = c.hidden_dim, ff = c.dim_feedforward;
  for (int s = 0; s < 2; ++s) {
    ProjPacked* pp = s == 0 ? L.vid : L.txt;
    int din = s == 0 ? c.v_feat_dim : c.t_feat_dim;
    for (int i = 0; i < c.n_input_proj; ++i) {
      pp[i].din = din;
      pp[i].kpad = pad64(din);
      pp[i].ln_w = cur.take((size_t)din * 4);
      pp[i].ln_b = cur.take((size_t)din * 4);
      pp[i].w16 = cur.take((size_t)d * pp[i].kpad * 2);
      pp[i].bias = cur.take((size_t)d * 4);
      din = d;
    }
  }
  for (int l = 0; l < c.enc_layers; ++l) {
    LayerPacked& lp = L.layer[l];
    lp.w_in = cur.take((size_t)3 * d * d * 2);
    lp.b_in = cur.take((size_t)3 * d * 4);
    lp.w_out = cur.take((size_t)d * d * 2);
    lp.b_out = cur.take((size_t)d * 4);
    lp.w1 = cur.take((size_t)ff * d * 2);
    lp.b1 = cur.take((size_t)ff * 4);
    lp.w2 = cur.take((size_t)d * ff * 2);
    lp.b2 = cur.take((size_t)d * 4);
    lp.n1w = cur.take((size_t)d * 4);
    lp.n1b = cur.take((size_t)d * 4);
    lp.n2w = cur.take((size_t)d * 4);
    lp.n2b = cur.take((size_t)d * 4);
  }
  L.conv1_w = cur.take((size_t)2 * d * 3 * d * 2);
  L.conv1_b = cur.take((size_t)2 * d * 4);
  L.conv2c_w = cur.take((size_t)d * 3 * d * 2);
  L.conv2c_b = cur.take((size_t)d * 4);
  L.conv2s_w = cur.take((size_t)d * 3 * d * 2);
  L.conv2s_b = cur.take((size_t)d * 4);
  L.conv3c_w = cur.take((size_t)3 * d * 4);
  L.conv3c_b = cur.take(4);
  L.conv3s_w = cur.take((size_t)2 * 3 * d * 4);
  L.conv3s_b = cur.take(8);
  L.pool_w = cur.take((size_t)d * 4);
  L.total = cur.off;
  return L;
}

// ------------------------------------------------------------------------------------------------
// pack kernels
// ------------------------------------------------------------------------------------------------
// All (re)packing work of one univtg_pack_weights call is described by a task table and executed by a handful of launches
// (the table travels as a kernel parameter, <= 4 KB per launch) instead of ~75 tiny kernels.
struct PackTask {
  const float* src;
  const float* add;  // kind 3: optional second vector added element-wise
  void* dst;
  int kind;          // 0: rows -> 16-bit [rows, ld] zero-padded, 1: conv [N,C,3] -> 16-bit [N, 3C], 2: conv -> fp32 [N,3,C], 3: vector copy(+add)
  int rows, cols, ld;
  int blk0;          // first block of this task in the launch (blocks are dealt out in proportion to the task's size)
};
constexpr int kPackTasksPerLaunch = 64;
constexpr int kPackItemsPerBlock = 256 * 8;  // work items per block (an item = 4 output elements, or one conv (n, c) pair)
struct PackTable {
  int n, fmt;
  PackTask t[kPackTasksPerLaunch];
};

inline size_t pack_task_items(const PackTask& k) {
  if (k.kind == 0) return ((size_t)k.rows * k.ld + 3) / 4;  // ld % 4 == 0 for every packed matrix (K padded to 64)
  if (k.kind == 1 || k.kind == 2) return (size_t)k.rows * k.cols;
  return ((size_t)k.rows + 3) / 4;
}

__global__ void __launch_bounds__(256) pack_multi_kernel(const __grid_constant__ PackTable tab) {
  pdl_prologue();
  int ti = 0;
  while (ti + 1 < tab.n && (int)blockIdx.x >= tab.t[ti + 1].blk0) ++ti;
  const PackTask& k = tab.t[ti];
  const int fmt = tab.fmt;
  const size_t first = (size_t)(blockIdx.x - k.blk0) * kPackItemsPerBlock;
  if (k.kind == 0) {
    const size_t items = ((size_t)k.rows * k.ld) / 4;
    const bool dense = k.ld == k.cols && (k.cols & 3) == 0 && (((uintptr_t)k.src) & 15) == 0;
    uint2* dst2 = reinterpret_cast<uint2*>(k.dst);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const size_t it = first + u * 256 + threadIdx.x;
      if (it >= items) break;
      float4 v;
      if (dense) {
        v = __ldg(reinterpret_cast<const float4*>(k.src) + it);
      } else {
        const int r = (int)((it * 4) / k.ld), c = (int)((it * 4) % k.ld);  // 4 consecutive columns of one row (ld % 4 == 0)
        const float* row = k.src + (size_t)r * k.cols;
        v.x = c + 0 < k.cols ? __ldg(row + c + 0) : 0.f;
        v.y = c + 1 < k.cols ? __ldg(row + c + 1) : 0.f;
        v.z = c + 2 < k.cols ? __ldg(row + c + 2) : 0.f;
        v.w = c + 3 < k.cols ? __ldg(row + c + 3) : 0.f;
      }
      dst2[it] = make_uint2(cvt16x2(v.x, v.y, fmt), cvt16x2(v.z, v.w, fmt));
    }
  } else if (k.kind == 1 || k.kind == 2) {
    // item = one (n, c) pair: three consecutive source floats (taps), scattered to the three tap planes of row n
    const int C = k.cols;
    const size_t items = (size_t)k.rows * C;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const size_t it = first + u * 256 + threadIdx.x;
      if (it >= items) break;
      const int n = (int)(it / C), c = (int)(it % C);
      const float* sp = k.src + it * 3;
      const float v0 = __ldg(sp), v1 = __ldg(sp + 1), v2 = __ldg(sp + 2);
      const size_t o = (size_t)n * 3 * C + c;
      if (k.kind == 1) {
        uint16_t* d16 = reinterpret_cast<uint16_t*>(k.dst);
        d16[o] = cvt16(v0, fmt);
        d16[o + C] = cvt16(v1, fmt);
        d16[o + 2 * C] = cvt16(v2, fmt);
      } else {
        float* d32 = reinterpret_cast<float*>(k.dst);
        d32[o] = v0;
        d32[o + C] = v1;
        d32[o + 2 * C] = v2;
      }
    }
  } else {
    float* dst = reinterpret_cast<float*>(k.dst);
    for (int u = 0; u < 8 * 4; ++u) {
      const size_t i = first * 4 + (size_t)u * 256 + threadIdx.x;
      if (i >= (size_t)k.rows) break;
      dst[i] = k.src[i] + (k.add ? k.add[i] : 0.f);
    }
  }
}

struct Packer {
  uint8_t* base;
  int fmt;
  cudaStream_t st;
  PackTable tab;
  bool skip_matrices = false;  // only the fp32 vectors / small tensors (kinds 2, 3)
  void push(const PackTask& t) {
    if (skip_matrices && (t.kind == 0 || t.kind == 1)) return;
    if (tab.n == kPackTasksPerLaunch) flush();
    tab.t[tab.n++] = t;
  }
  void flush() {
    if (tab.n == 0) return;
    tab.fmt = fmt;
    int blocks = 0;
    for (int i = 0; i < tab.n; ++i) {
      tab.t[i].blk0 = blocks;
      blocks += (int)((pack_task_items(tab.t[i]) + kPackItemsPerBlock - 1) / kPackItemsPerBlock);
    }
    if (blocks > 0) launch_k(pack_multi_kernel, dim3(blocks), dim3(256), 0, st, tab);
    tab.n = 0;
  }
  void rows(const float* src, size_t off, int rows_, int cols, int ld) { push(PackTask{src, nullptr, base + off, 0, rows_, cols, ld, 0}); }
  void conv(const float* src, size_t off, int N, int C) { push(PackTask{src, nullptr, base + off, 1, N, C, 0, 0}); }
  void conv_f32(const float* src, size_t off, int N, int C) { push(PackTask{src, nullptr, base + off, 2, N, C, 0, 0}); }
  void vec(const float* src, size_t off, int n, const float* add = nullptr) { push(PackTask{src, add, base + off, 3, n, 0, 0, 0}); }
};

}  // namespace

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
constexpr int kMaxMarks = 640;
struct univtg_plan {
  univtg_config cfg;
  univtg_shape shp;
  PackedLayout lay;
  const uint8_t* packed;
  uint8_t* ws;
  const float* dim_t;
  int num_sms;
  int in_fmt;       // src_vid / src_txt element type: 0 f32 (reference collate), 1 fp16, 2 bf16 (packed feature shards)
  int num_sms_bwd;  // SM budget of the backward's GEMM launches (0: num_sms); see univtg_plan_set_backward_sm_budget
  int B, Lv, Lt, L, d, ff, H, dh, M, Mv, Mt, Mh;
  // workspace pointers
  uint16_t *a_vid[3], *a_txt[3];  // LN'd 16-bit projector inputs
  float *p_vid32, *p_txt32;       // fp32 projector hidden (between projector layers)
  float* txtproj32;               // [Mt, d] projected text tokens (+type embedding)
  float* pos;                     // [Mv, d]
  float* key_mask;                // [B, L]
  float* pool_logits;             // [B, Lt]
  float* x32;                     // fp32 residual stream (LayerNorm works in place on it)
  uint16_t *x16, *xpos16, *qkv16, *attn16, *h16, *br16;  // br16: DropPath-scaled residual branch (out-proj / FFN2 output)
  uint16_t *hA, *h1, *hc2, *hs2;  // conv-head buffers (separated layout)
  // launch descriptors
  GemmGroup g_proj[3];
  GemmGroup g_qkv[16], g_out[16], g_ffn1[16], g_ffn2[16];
  GemmGroup g_conv1, g_conv2;
  AttnArgs attn[16];
  int bn_proj[3], bn_main;  // tile widths chosen per launch (choose_bn)
  int bn_qkv, bn_out, bn_ffn1, bn_ffn2, bn_conv1, bn_conv2;
  int launches;
  // optional "gradients of stage k are final" events recorded by univtg_backward (gradient-exchange overlap)
  int n_grad_events;
  cudaEvent_t grad_events[24];
  // optional per-launch CUDA-event timeline (bench / profiling only)
  int profiling;
  int n_marks;
  cudaEvent_t marks[kMaxMarks];
  int mark_kind[kMaxMarks];  // kind of the interval that ENDS at mark i (i >= 1): 0 row kernel, 1 tcgen05 GEMM, 2 attention, 3 other work
};

namespace {
inline void prof_begin(univtg_plan* P, cudaStream_t st) {
  if (!P->profiling) return;
  P->n_marks = 0;
  if (!P->marks[0]) cudaEventCreate(&P->marks[0]);
  cudaEventRecord(P->marks[0], st);
  P->mark_kind[0] = -1;
  P->n_marks = 1;
}
inline void prof_mark(univtg_plan* P, cudaStream_t st, int kind) {
  if (!P->profiling || P->n_marks >= kMaxMarks) return;
  const int i = P->n_marks;
  if (!P->marks[i]) cudaEventCreate(&P->marks[i]);
  cudaEventRecord(P->marks[i], st);
  P->mark_kind[i] = kind;
  P->n_marks = i + 1;
}
// training path: a GEMM / attention launch bracketed by two marks, so that the interval ending at the second mark is that kernel
// alone (the interval ending at the first one - kind 3 - collects whatever ran since the previous mark)
inline int gemm_launch(univtg_plan* P, GemmGroup& g, int bn, int sms, cudaStream_t st) {
  prof_mark(P, st, 3);
  const int rc = launch_gemm_group(g, bn, sms, st);
  prof_mark(P, st, 1);
  return rc;
}
}  // namespace

namespace {

void init_problem(GemmProblem& p) {
  memset(&p, 0, sizeof(p));
  p.taps = 1;
  p.ksplit = 1;
  p.alpha = 1.f;
  p.a_fmt = p.b_fmt = p.out_fmt = -1;
  p.colsum_scale = 1.f;
  // default coordinate rules: K-major A [M,K] and B [N,K]
  p.ca = OperandCoord{0, 0, 0, 1, 0, 1, 0, 0};
  p.cb = OperandCoord{0, 0, 0, 1, 0, 1, 0, 0};
}

// K-major linear problem: A [M, K] (pitch lda), W [N, K] (pitch ldw), K multiple of 64.
int setup_linear(GemmProblem& p, const uint16_t* A, int M, int K, int lda, const uint16_t* W, int N, int ldw, int bn) {
  init_problem(p);
  p.M = M;
  p.N = N;
  p.kblk_per_tap = K / 64;
  if (K % 64 != 0) {
    set_error("setup_linear: K %d not a multiple of 64", K);
    return 1;
  }
  if (make_tmap_2d(&p.tm_a, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GEMM_BM, 64)) return 1;
  if (make_tmap_2d(&p.tm_b, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, (uint32_t)bn, 64)) return 1;
  p.b_box_rows = bn;
  return 0;
}

}  // namespace


namespace {

struct WsLayout {
  size_t a_vid[3], a_txt[3], p_vid32, p_txt32, txtproj32, pos, key_mask, pool_logits, x32, br16, x16, xpos16, qkv16, attn16, h16, hA,
      h1, hc2, hs2, total;
};

WsLayout make_ws(const univtg_config& c, const univtg_shape& s, const PackedLayout& L) {
  WsLayout w;
  memset(&w, 0, sizeof(w));
  Cursor cur;
  const size_t d = c.hidden_dim, ff = c.dim_feedforward;
  const size_t B = s.batch, Lv = s.l_vid, Lt = s.l_txt, Lc = Lv + Lt;
  const size_t M = B * Lc, Mv = B * Lv, Mt = B * Lt, Mh = B * (Lv + 1);
  for (int i = 0; i < c.n_input_proj; ++i) {
    w.a_vid[i] = cur.take(Mv * L.vid[i].kpad * 2);
    w.a_txt[i] = cur.take(Mt * L.txt[i].kpad * 2);
  }
  w.p_vid32 = cur.take(Mv * d * 4);
  w.p_txt32 = cur.take(Mt * d * 4);
  w.txtproj32 = cur.take(Mt * d * 4);
  w.pos = cur.take(Mv * d * 4);
  w.key_mask = cur.take(B * Lc * 4);
  w.pool_logits = cur.take(B * Lt * 4);
  w.x32 = cur.take(M * d * 4);
  w.br16 = cur.take(M * d * 2);
  w.x16 = cur.take(M * d * 2);
  w.xpos16 = cur.take(M * d * 2);
  w.qkv16 = cur.take(M * 3 * d * 2);
  w.attn16 = cur.take(M * d * 2);
  w.h16 = cur.take(M * ff * 2);
  w.hA = cur.take((Mh + 2) * d * 2);
  w.h1 = cur.take((Mh + 2) * 2 * d * 2);
  w.hc2 = cur.take((Mh + 2) * d * 2);
  w.hs2 = cur.take((Mh + 2) * d * 2);
  w.total = cur.off;
  return w;
}

bool check_shape(const univtg_shape* s) {
  if (!s || s->batch < 1 || s->l_vid < 1 || s->l_txt < 1) {
    set_error("bad shape");
    return false;
  }
  return true;
}

}  // namespace

