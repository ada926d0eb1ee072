// Persistent, warp-specialised tcgen05 GEMM for sm_100a with a fused, run-time configured epilogue.
//
//   C[M,N] = epilogue( sum_k A[m,k] * B[n,k] ),  A/B 16-bit (fp16 or bf16), fp32 accumulation in TMEM.
//
// This one kernel serves every dense contraction of the UniVTG hot path (SURVEY.md §2.2 rows K1-K3, K5, K7,
// K9-K11 and their backward passes): input projectors (model/univtg.py:399-406), the three attention
// in-projections + out-projection (torch MHA called at model/transformer_encoder_droppath.py:118), the FFN
// (:122) and the k=3 Conv1d heads (model/univtg.py:378-382) expressed as three row-shifted K segments.
//
// Roles (384 threads, 1 CTA / SM, persistent over a static round-robin tile schedule):
//   warp 0 lane 0 : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier expect_tx)
//   warp 1 lane 0 : MMA issuer     (tcgen05.mma cta_group::1, M=128, N=BN, K=16 x4 per 64-wide k-block)
//   warp 2        : TMEM allocator (2 accumulator stages of BN fp32 columns)
//   warps 4..11   : epilogue       (tcgen05.ld 32x32b -> swizzled smem tile -> full-line 128-bit global loads/stores)
#include <stdarg.h>
#include <stdio.h>

#include "kernels.h"
#include "ptx.cuh"

namespace uv {

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = GEMM_BM * 128;          // 128 rows x 64 x 2 B
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;  // multiple of 1024
  static constexpr int kEpiFloats = 8 * 1024 + 2 * 8 * 32;  // per-warp 32x32 fp32 staging tiles + per-row (out row, scale)
  static constexpr int kSmemBytes = 512 /*align slack*/ + kStages * kStageBytes + kEpiFloats * 4 + 192;
  static constexpr uint32_t kTmemCols = 2 * BN;          // 256 or 512 (power of two)
};

struct TileInfo {
  int p, m_blk, n_blk, kb0, kb1, split;
};

__device__ __forceinline__ bool decode_tile(const GemmGroup& g, int bn, int t, TileInfo& ti) {
  for (int p = 0; p < g.num; ++p) {
    const GemmProblem& pr = g.p[p];
    const int tm = (pr.M + GEMM_BM - 1) / GEMM_BM;
    const int tn = (pr.N + bn - 1) / bn;
    const int cnt = tm * tn * pr.ksplit;
    if (t < cnt) {
      ti.p = p;
      ti.n_blk = t % tn;
      const int rest = t / tn;
      ti.m_blk = rest % tm;
      ti.split = rest / tm;
      const int total_kb = pr.taps * pr.kblk_per_tap;
      const int per = (total_kb + pr.ksplit - 1) / pr.ksplit;
      ti.kb0 = ti.split * per;
      ti.kb1 = min(total_kb, ti.kb0 + per);
      return true;
    }
    t -= cnt;
  }
  return false;
}

__device__ __forceinline__ void stamp(unsigned long long* dbg, int slot) {
  if (dbg != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    dbg[(size_t)blockIdx.x * 8 + slot] = t;
  }
}

template <int BN>
__global__ void __launch_bounds__(384, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmGroup g) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t align_off = (1024u - (raw_addr & 1023u)) & 1023u;  // 0 when the runtime honours the 1024 B alignment
  if (align_off > 512u) __trap();                                  // the allocation carries only 512 B of slack
  uint8_t* smem = smem_raw + align_off;

  uint8_t* stage_base = smem;
  float* epi_buf = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_buf + Cfg::kEpiFloats);
  uint64_t* full_bar = bars;                        // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;        // [kStages]
  uint64_t* tmem_full = bars + 2 * Cfg::kStages;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;             // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) stamp(g.dbg, 0);  // kernel entry

  if (warp == 0 && lane == 0) {
    for (int p = 0; p < g.num; ++p) {
      tma_prefetch_desc(&g.p[p].tm_a);
      tma_prefetch_desc(&g.p[p].tm_b);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  if (threadIdx.x == 0) stamp(g.dbg, 1);  // setup done (barriers, TMEM)

  if (warp == 0) {
    // ======================================= TMA producer =======================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      TileInfo ti;
      for (int t = blockIdx.x; decode_tile(g, BN, t, ti); t += gridDim.x) {
        const GemmProblem& pr = g.p[ti.p];
        const int m0 = ti.m_blk * GEMM_BM;
        const int n0 = ti.n_blk * BN;
        for (int kb = ti.kb0; kb < ti.kb1; ++kb) {
          const int tap = kb / pr.kblk_per_tap;
          const int kk = (kb - tap * pr.kblk_per_tap) * GEMM_BK;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = stage_base + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          const int a0 = pr.ca.base0 + m0 * pr.ca.mn0s + tap * pr.ca.tap0 + kk * pr.ca.k0s;
          const int a1 = pr.ca.base1 + m0 * pr.ca.mn1s + tap * pr.ca.tap1 + kk * pr.ca.k1s;
          if (!pr.a_mn) {
            tma_load_2d(sa, &pr.tm_a, &full_bar[stage], a0, a1);
          } else {
#pragma unroll
            for (int j = 0; j < GEMM_BM / 64; ++j) tma_load_2d(sa + j * 8192, &pr.tm_a, &full_bar[stage], a0 + 64 * j, a1);
          }
          const int b0 = pr.cb.base0 + n0 * pr.cb.mn0s + tap * pr.cb.tap0 + kk * pr.cb.k0s;
          const int b1 = pr.cb.base1 + n0 * pr.cb.mn1s + tap * pr.cb.tap1 + kk * pr.cb.k1s;
          if (!pr.b_mn) {
            tma_load_2d(sb, &pr.tm_b, &full_bar[stage], b0, b1);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &pr.tm_b, &full_bar[stage], b0 + 64 * j, b1);
          }
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      stamp(g.dbg, 2);  // all TMA loads issued
    }
  } else if (warp == 1) {
    // ======================================== MMA issuer ========================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      TileInfo ti;
      for (int t = blockIdx.x; decode_tile(g, BN, t, ti); t += gridDim.x) {
        const GemmProblem& pr = g.p[ti.p];
        const uint32_t idesc = make_idesc_f16_ab(GEMM_BM, BN, pr.a_fmt < 0 ? g.fmt : pr.a_fmt, pr.b_fmt < 0 ? g.fmt : pr.b_fmt, pr.a_mn, pr.b_mn);
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = ti.kb0; kb < ti.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          if (kb == ti.kb0 && t == (int)blockIdx.x) stamp(g.dbg, 3);  // first operand stage landed
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // K-major: advance 16 elements (32 B) inside the 128 B swizzle span.
            // MN-major: advance 16 k-rows = two 1024 B swizzle atoms.
            const uint64_t da = pr.a_mn ? make_smem_desc_sw128(sa + k * 2048, 8192, 1024)
                                        : make_smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t db = pr.b_mn ? make_smem_desc_sw128(sb + k * 2048, 8192, 1024)
                                        : make_smem_desc_sw128(sb + k * 32, 16, 1024);
            umma_f16_ss(d_tmem, da, db, idesc, (kb > ti.kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[as]);  // accumulator complete -> epilogue
        stamp(g.dbg, 4);              // last MMA of the tile issued
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ========================================= epilogue =========================================
    // 8 warps: warp w may only touch TMEM lanes [32*(w%4), +32); the two warps of a lane quarter split the columns.
    // Per 32x32 step: tcgen05.ld (thread = row) -> XOR-swizzled 4 KB smem tile -> "coalesced domain" where 8 lanes cover
    // one 128 B row segment (4 rows per 128-bit instruction).  All math, residual/aux loads and stores happen in the
    // coalesced domain, so every global access is a full 128 B line (the LSU wavefront count is what bounds this phase).
    const int wq = warp & 3;
    const int half = (warp - 4) >> 2;
    const int ew = warp - 4;
    float* stg = epi_buf + ew * 1024;                 // [32 rows][32 cols] fp32, quad q of row r at physical quad q ^ (r & 7)
    int* orow_s = reinterpret_cast<int*>(epi_buf + 8 * 1024) + ew * 32;   // output row per tile row (-1: skip)
    float* rsc_s = epi_buf + 8 * 1024 + 8 * 32 + ew * 32;                 // alpha * row_scale per tile row
    const int fmt = g.fmt;
    const int rsub = lane >> 3;  // row inside a group of 4
    const int cq = lane & 7;     // 4-column quad inside the 32-column step
    int as = 0;
    uint32_t aphase = 0;
    TileInfo ti;
    for (int t = blockIdx.x; decode_tile(g, BN, t, ti); t += gridDim.x) {
      const GemmProblem& pr = g.p[ti.p];
      // hoist the problem description into registers (the struct lives in the constant bank)
      const int pM = pr.M, pN = pr.N, rps_in = pr.rps_in;
      const int act = pr.act;
      const float* __restrict__ bias = (ti.split == 0) ? pr.bias : nullptr;
      const float* __restrict__ resid = pr.resid;
      const float* __restrict__ addtab = pr.addtab;
      float* __restrict__ out32 = pr.out32;
      float* __restrict__ out32_id = pr.out32_id;
      uint16_t* __restrict__ out16 = pr.out16;
      uint16_t* __restrict__ out16p = pr.out16p;
      float* __restrict__ pre32 = pr.pre32;
      const float* __restrict__ aux32 = pr.aux32;
      const uint16_t* __restrict__ mask16 = pr.mask16;
      float* __restrict__ colsum = pr.colsum;
      const int ld_resid = pr.ld_resid, ld_addtab = pr.ld_addtab, ld32 = pr.ld32, ld32_id = pr.ld32_id, ld16 = pr.ld16;
      const int ld_pre = pr.ld_pre, ld_aux = pr.ld_aux, ld_mask = pr.ld_mask, aux_mode = pr.aux_mode;
      const bool atomic = (pr.accumulate != 0) || (pr.ksplit > 1);
      const bool vec = pr.vec_ok != 0;
      const int ofmt = pr.out_fmt < 0 ? fmt : pr.out_fmt;
      const int cs32 = pr.cs32 > 1 ? pr.cs32 : 1;

      const int m0 = ti.m_blk * GEMM_BM + wq * 32;
      const int n_base = ti.n_blk * BN + half * (BN / 2);
      // ---- per-row bookkeeping (thread = row) ----
      __syncwarp();
      {
        const int m = m0 + lane;
        int b = 0, l = m;
        if (rps_in > 0) {
          b = m / rps_in;
          l = m - b * rps_in;
        }
        const bool is_sep = (rps_in > 0) && (l == rps_in - 1);
        const bool valid = (m < pM) && !(pr.skip_sep && is_sep);
        float rsc = pr.alpha;
        if (pr.row_scale != nullptr && m < pM) rsc *= pr.row_scale[b];
        if (pr.zero_sep && is_sep) rsc = 0.f;
        orow_s[lane] = valid ? ((rps_in > 0 ? b * pr.rps_out + l : m) + pr.row_off) : -1;
        rsc_s[lane] = rsc;
      }
      __syncwarp();

      mbar_wait(&tmem_full[as], aphase);
      if (ew == 0 && lane == 0) stamp(g.dbg, 5);  // accumulator ready
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(wq * 32) << 16) + as * BN + half * (BN / 2);
      int nsteps = (pN - n_base + 31) / 32;  // 32-column steps this warp owns in this tile (warp-uniform)
      nsteps = nsteps < 0 ? 0 : (nsteps > BN / 64 ? BN / 64 : nsteps);

      uint32_t r[32];
      if (nsteps > 0) tmem_ld_32x32b_x32(t_addr, r);
      for (int c = 0; c < nsteps; ++c) {
        const int n0 = n_base + c * 32;
        const int n = n0 + 4 * cq;  // this lane's 4 columns
        // residual rows of this step: issue the loads before waiting on TMEM
        float4 rv[8];
        if (vec && resid != nullptr && n < pN) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int orr = orow_s[rsub + 4 * i];
            rv[i] = orr >= 0 ? *reinterpret_cast<const float4*>(resid + (size_t)orr * ld_resid + n) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias != nullptr && n < pN) {
          if (vec) bv = __ldg(reinterpret_cast<const float4*>(bias + n));
          else {
            bv.x = __ldg(bias + n);
            if (n + 1 < pN) bv.y = __ldg(bias + n + 1);
            if (n + 2 < pN) bv.z = __ldg(bias + n + 2);
            if (n + 3 < pN) bv.w = __ldg(bias + n + 3);
          }
        }
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(stg + lane * 32 + 4 * (q ^ (lane & 7))) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        __syncwarp();
        if (c + 1 < nsteps) tmem_ld_32x32b_x32(t_addr + (c + 1) * 32, r);  // in flight while this step is processed

        float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < pN) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = rsub + 4 * i;
            const int orr = orow_s[rr];
            if (orr < 0) continue;
            float4 v = *reinterpret_cast<const float4*>(stg + rr * 32 + 4 * (cq ^ (rr & 7)));
            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
            const int mrow = m0 + rr;
            if (pre32 != nullptr) *reinterpret_cast<float4*>(pre32 + (size_t)orr * ld_pre + n) = v;  // pre-activation (training)
            if (act == ACT_GELU) {
              v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
            } else if (act == ACT_RELU) {
              v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
            }
            const float rs = rsc_s[rr];
            v.x *= rs; v.y *= rs; v.z *= rs; v.w *= rs;
            if (vec) {
              if (resid != nullptr) {
                v.x += rv[i].x; v.y += rv[i].y; v.z += rv[i].z; v.w += rv[i].w;
              }
              if (aux32 != nullptr) {
                const float4 av = *reinterpret_cast<const float4*>(aux32 + (size_t)orr * ld_aux + n);
                if (aux_mode == 1) {
                  v.x *= gelu_erf_grad(av.x); v.y *= gelu_erf_grad(av.y); v.z *= gelu_erf_grad(av.z); v.w *= gelu_erf_grad(av.w);
                } else {
                  v.x *= av.x; v.y *= av.y; v.z *= av.z; v.w *= av.w;
                }
              }
              if (mask16 != nullptr) {
                const uint2 mk = *reinterpret_cast<const uint2*>(mask16 + (size_t)orr * ld_mask + n);
                if (!pos16((uint16_t)(mk.x & 0xffff))) v.x = 0.f;
                if (!pos16((uint16_t)(mk.x >> 16))) v.y = 0.f;
                if (!pos16((uint16_t)(mk.y & 0xffff))) v.z = 0.f;
                if (!pos16((uint16_t)(mk.y >> 16))) v.w = 0.f;
              }
              csum.x += v.x; csum.y += v.y; csum.z += v.z; csum.w += v.w;
              if (out32 != nullptr) {
                float* dst = out32 + (size_t)orr * ld32 + n;
                if (atomic) {
                  atomicAdd(dst, v.x); atomicAdd(dst + 1, v.y); atomicAdd(dst + 2, v.z); atomicAdd(dst + 3, v.w);
                } else {
                  *reinterpret_cast<float4*>(dst) = v;
                }
              }
              if (out32_id != nullptr) *reinterpret_cast<float4*>(out32_id + (size_t)mrow * ld32_id + n) = v;
              if (out16 != nullptr)
                *reinterpret_cast<uint2*>(out16 + (size_t)orr * ld16 + n) = make_uint2(cvt16x2(v.x, v.y, ofmt), cvt16x2(v.z, v.w, ofmt));
              if (out16p != nullptr) {
                float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (addtab != nullptr) pv = *reinterpret_cast<const float4*>(addtab + (size_t)mrow * ld_addtab + n);
                *reinterpret_cast<uint2*>(out16p + (size_t)orr * ld16 + n) =
                    make_uint2(cvt16x2(v.x + pv.x, v.y + pv.y, ofmt), cvt16x2(v.z + pv.z, v.w + pv.w, ofmt));
              }
            } else {
              // unaligned leading dimensions / ragged N (e.g. the [d, 2818] projector weight gradient): scalar accesses
              const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (n + e >= pN) break;
                float x = vv[e];
                if (resid != nullptr) x += resid[(size_t)orr * ld_resid + n + e];
                if (aux32 != nullptr) {
                  const float av = aux32[(size_t)orr * ld_aux + n + e];
                  x *= (aux_mode == 1) ? gelu_erf_grad(av) : av;
                }
                if (mask16 != nullptr && !pos16(mask16[(size_t)orr * ld_mask + n + e])) x = 0.f;
                (&csum.x)[e] += x;
                if (out32 != nullptr) {
                  float* dst = out32 + (size_t)orr * ld32 + (size_t)(n + e) * cs32;
                  if (atomic) atomicAdd(dst, x);
                  else *dst = x;
                }
                if (out32_id != nullptr) out32_id[(size_t)mrow * ld32_id + n + e] = x;
                if (out16 != nullptr) out16[(size_t)orr * ld16 + n + e] = cvt16(x, ofmt);
                if (out16p != nullptr)
                  out16p[(size_t)orr * ld16 + n + e] = cvt16(x + (addtab ? addtab[(size_t)mrow * ld_addtab + n + e] : 0.f), ofmt);
              }
            }
          }
        }
        if (colsum != nullptr) {  // warp-uniform; lanes cq, cq+8, cq+16, cq+24 hold partial sums of the same 4 columns
#pragma unroll
          for (int o = 8; o < 32; o <<= 1) {
            csum.x += __shfl_xor_sync(0xffffffffu, csum.x, o);
            csum.y += __shfl_xor_sync(0xffffffffu, csum.y, o);
            csum.z += __shfl_xor_sync(0xffffffffu, csum.z, o);
            csum.w += __shfl_xor_sync(0xffffffffu, csum.w, o);
          }
          if (rsub == 0 && n < pN) {
            const float cs = pr.colsum_scale;
            atomicAdd(colsum + n, csum.x * cs);
            if (n + 1 < pN) atomicAdd(colsum + n + 1, csum.y * cs);
            if (n + 2 < pN) atomicAdd(colsum + n + 2, csum.z * cs);
            if (n + 3 < pN) atomicAdd(colsum + n + 3, csum.w * cs);
          }
        }
        __syncwarp();  // staging tile is free again
      }
      // release the accumulator stage
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      if (ew == 7 && lane == 0) stamp(g.dbg, 6);  // epilogue of the tile done (last warp)
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
  if (threadIdx.x == 0) stamp(g.dbg, 7);  // exit
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static unsigned long long* g_timeline = nullptr;
void set_gemm_timeline_buffer(unsigned long long* buf) { g_timeline = buf; }
static thread_local char g_err[512] = "";
const char* last_error() { return g_err; }
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (cudaError %d)", (int)e);
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows,
                 uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return 1;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld_elems * 2) % 16 != 0) {
    set_error("tensor map: base %p / pitch %llu B not 16-byte aligned", base, (unsigned long long)(ld_elems * 2));
    return 2;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rows %llu cols %llu ld %llu box %ux%u)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_rows, box_cols);
    return 3;
  }
  return 0;
}

template <int BN>
static int launch_bn(GemmGroup& g, int num_sms, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm, smem=%d): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  int total = 0;
  for (int p = 0; p < g.num; ++p) {
    const GemmProblem& pr = g.p[p];
    if (pr.ksplit < 1 || pr.taps < 1 || pr.kblk_per_tap < 1 || pr.ksplit > pr.taps * pr.kblk_per_tap) {
      set_error("gemm problem %d: bad k configuration (taps %d kblk %d ksplit %d)", p, pr.taps, pr.kblk_per_tap, pr.ksplit);
      return (int)cudaErrorInvalidValue;
    }
    const int total_kb = pr.taps * pr.kblk_per_tap;
    const int per = (total_kb + pr.ksplit - 1) / pr.ksplit;
    if ((pr.ksplit - 1) * per >= total_kb) {
      set_error("gemm problem %d: ksplit %d leaves an empty split for %d k-blocks", p, pr.ksplit, total_kb);
      return (int)cudaErrorInvalidValue;
    }
    total += ((pr.M + GEMM_BM - 1) / GEMM_BM) * ((pr.N + BN - 1) / BN) * pr.ksplit;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    GemmProblem& w = g.p[p];
    w.vec_ok = (pr.N % 4 == 0) && (pr.cs32 <= 1) && al16(pr.bias) && (!pr.aux32 || (al16(pr.aux32) && pr.ld_aux % 4 == 0)) &&
               (!pr.pre32 || (al16(pr.pre32) && pr.ld_pre % 4 == 0)) &&
               (!pr.mask16 || ((reinterpret_cast<uintptr_t>(pr.mask16) & 7) == 0 && pr.ld_mask % 4 == 0)) && (!pr.resid || (al16(pr.resid) && pr.ld_resid % 4 == 0)) &&
               (!pr.addtab || (al16(pr.addtab) && pr.ld_addtab % 4 == 0)) && (!pr.out32 || (al16(pr.out32) && pr.ld32 % 4 == 0)) &&
               (!pr.out32_id || (al16(pr.out32_id) && pr.ld32_id % 4 == 0)) &&
               ((!pr.out16 && !pr.out16p) || (pr.ld16 % 4 == 0 && (reinterpret_cast<uintptr_t>(pr.out16) & 7) == 0 &&
                                              (reinterpret_cast<uintptr_t>(pr.out16p) & 7) == 0));
  }
  if (total == 0) return 0;
  const int grid = total < num_sms ? total : num_sms;
  g.dbg = g_timeline;
  gemm_tcgen05_kernel<BN><<<grid, 384, Cfg::kSmemBytes, stream>>>(g);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("gemm launch failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

int launch_gemm_group(GemmGroup& g, int bn, int num_sms, cudaStream_t stream) {
  if (g.num < 1 || g.num > GEMM_MAX_GROUP) {
    set_error("gemm group size %d out of range", g.num);
    return (int)cudaErrorInvalidValue;
  }
  if (bn == 256) return launch_bn<256>(g, num_sms, stream);
  if (bn == 128) return launch_bn<128>(g, num_sms, stream);
  set_error("unsupported BN %d", bn);
  return (int)cudaErrorInvalidValue;
}

}  // namespace uv
