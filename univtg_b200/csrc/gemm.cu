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
//   warps 4..11   : epilogue       (tcgen05.ld 32x32b -> smem transpose -> 128-bit coalesced global loads/stores)
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
  static constexpr int kEpiFloats = 8 * 32 * 16;         // per-warp 32x16 fp32 transpose buffers (XOR-swizzled quads)
  static constexpr int kSmemBytes = 1024 /*align slack*/ + kStages * kStageBytes + kEpiFloats * 4 + 8 * 32 * 8 + 256;
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

template <int BN>
__global__ void __launch_bounds__(384, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmGroup g) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint8_t* stage_base = smem;
  float* epi_buf = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes);
  int* s_orow = reinterpret_cast<int*>(epi_buf + Cfg::kEpiFloats);        // [8][32]
  float* s_rscale = reinterpret_cast<float*>(s_orow + 8 * 32);            // [8][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_rscale + 8 * 32);
  uint64_t* full_bar = bars;                        // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;        // [kStages]
  uint64_t* tmem_full = bars + 2 * Cfg::kStages;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;             // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

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
        const uint32_t idesc = make_idesc_f16(GEMM_BM, BN, g.fmt, pr.a_mn, pr.b_mn);
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = ti.kb0; kb < ti.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
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
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ========================================= epilogue =========================================
    // 8 warps: warp w may only touch TMEM lanes [32*(w%4), +32); the two warps of a lane quarter split the columns.
    const int wq = warp & 3;
    const int half = (warp - 4) >> 2;
    const int ew = warp - 4;
    float* buf = epi_buf + ew * (32 * 16);
    int* orow_s = s_orow + ew * 32;
    float* rscale_s = s_rscale + ew * 32;
    const int fmt = g.fmt;
    const int rsub = lane >> 2;  // row inside a group of 8
    const int cq = lane & 3;     // which 4-column quad of the 16-column chunk
    int as = 0;
    uint32_t aphase = 0;
    TileInfo ti;
    for (int t = blockIdx.x; decode_tile(g, BN, t, ti); t += gridDim.x) {
      const GemmProblem& pr = g.p[ti.p];
      // hoist the problem description into registers (the struct lives in the constant bank)
      const int pM = pr.M, pN = pr.N, rps_in = pr.rps_in, rps_out = pr.rps_out, row_off = pr.row_off;
      const int act = pr.act;
      const float* __restrict__ bias = (ti.split == 0) ? pr.bias : nullptr;
      const float* __restrict__ resid = pr.resid;
      const float* __restrict__ addtab = pr.addtab;
      float* __restrict__ out32 = pr.out32;
      float* __restrict__ out32_id = pr.out32_id;
      uint16_t* __restrict__ out16 = pr.out16;
      uint16_t* __restrict__ out16p = pr.out16p;
      const int ld_resid = pr.ld_resid, ld_addtab = pr.ld_addtab, ld32 = pr.ld32, ld32_id = pr.ld32_id, ld16 = pr.ld16;
      const bool atomic = (pr.accumulate != 0) || (pr.ksplit > 1);
      const bool vec = pr.vec_ok != 0;

      const int m0 = ti.m_blk * GEMM_BM + wq * 32;
      const int n_base = ti.n_blk * BN + half * (BN / 2);
      // ---- per-row bookkeeping (thread = row) ----
      {
        const int m = m0 + lane;
        int b = 0, l = m;
        if (rps_in > 0) {
          b = m / rps_in;
          l = m - b * rps_in;
        }
        const bool valid = m < pM;
        const bool sep = pr.zero_sep && (l == rps_in - 1);
        float rsc = pr.alpha;
        if (pr.row_scale != nullptr && valid) rsc *= pr.row_scale[b];
        if (sep) rsc = 0.f;
        orow_s[lane] = valid ? ((rps_in > 0 ? b * rps_out + l : m) + row_off) : -1;
        rscale_s[lane] = rsc;
      }
      __syncwarp();

      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(wq * 32) << 16) + as * BN + half * (BN / 2);

      for (int c = 0; c < BN / 32; ++c) {
        const int n0 = n_base + c * 16;
        if (n0 >= pN) break;  // warp-uniform
        uint32_t r[16];
        tmem_ld_32x32b_x16(t_addr + c * 16, r);
        tmem_ld_wait();
        // ---- transpose through smem: thread = row writes 16 columns as 4 quads; afterwards 4 lanes cover one row.
        //      quad q of row r lives at physical quad q ^ ((r >> 1) & 3): conflict-free for both access patterns ----
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(buf + lane * 16 + 4 * (q ^ ((lane >> 1) & 3))) =
              make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        __syncwarp();
        const int n = n0 + 4 * cq;
        if (n < pN) {
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias != nullptr) {
            if (vec) bv = __ldg(reinterpret_cast<const float4*>(bias + n));
            else {
              bv.x = __ldg(bias + n);
              if (n + 1 < pN) bv.y = __ldg(bias + n + 1);
              if (n + 2 < pN) bv.z = __ldg(bias + n + 2);
              if (n + 3 < pN) bv.w = __ldg(bias + n + 3);
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = rsub + 8 * i;
            const int orr = orow_s[rr];
            if (orr < 0) continue;
            float4 v = *reinterpret_cast<const float4*>(buf + rr * 16 + 4 * (cq ^ ((rr >> 1) & 3)));
            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
            if (act == ACT_RELU) {
              v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
            } else if (act == ACT_GELU) {
              v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
            }
            const float rs = rscale_s[rr];
            v.x *= rs; v.y *= rs; v.z *= rs; v.w *= rs;
            const int mrow = m0 + rr;
            if (vec) {
              if (resid != nullptr) {
                const float4 rv = *reinterpret_cast<const float4*>(resid + (size_t)orr * ld_resid + n);
                v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w;
              }
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
                *reinterpret_cast<uint2*>(out16 + (size_t)orr * ld16 + n) = make_uint2(cvt16x2(v.x, v.y, fmt), cvt16x2(v.z, v.w, fmt));
              if (out16p != nullptr) {
                float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (addtab != nullptr) pv = *reinterpret_cast<const float4*>(addtab + (size_t)mrow * ld_addtab + n);
                *reinterpret_cast<uint2*>(out16p + (size_t)orr * ld16 + n) =
                    make_uint2(cvt16x2(v.x + pv.x, v.y + pv.y, fmt), cvt16x2(v.z + pv.z, v.w + pv.w, fmt));
              }
            } else {
              // unaligned leading dimensions (e.g. the [d, 2818] projector weight gradient): scalar stores
              const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (n + e >= pN) break;
                float x = vv[e];
                if (resid != nullptr) x += resid[(size_t)orr * ld_resid + n + e];
                if (out32 != nullptr) {
                  float* dst = out32 + (size_t)orr * ld32 + n + e;
                  if (atomic) atomicAdd(dst, x);
                  else *dst = x;
                }
                if (out32_id != nullptr) out32_id[(size_t)mrow * ld32_id + n + e] = x;
                if (out16 != nullptr) out16[(size_t)orr * ld16 + n + e] = cvt16(x, fmt);
                if (out16p != nullptr)
                  out16p[(size_t)orr * ld16 + n + e] = cvt16(x + (addtab ? addtab[(size_t)mrow * ld_addtab + n + e] : 0.f), fmt);
              }
            }
          }
        }
        __syncwarp();
      }
      // release the accumulator stage
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
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
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
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
    w.vec_ok = (pr.N % 4 == 0) && al16(pr.bias) && (!pr.resid || (al16(pr.resid) && pr.ld_resid % 4 == 0)) &&
               (!pr.addtab || (al16(pr.addtab) && pr.ld_addtab % 4 == 0)) && (!pr.out32 || (al16(pr.out32) && pr.ld32 % 4 == 0)) &&
               (!pr.out32_id || (al16(pr.out32_id) && pr.ld32_id % 4 == 0)) &&
               ((!pr.out16 && !pr.out16p) || (pr.ld16 % 4 == 0 && (reinterpret_cast<uintptr_t>(pr.out16) & 7) == 0 &&
                                              (reinterpret_cast<uintptr_t>(pr.out16p) & 7) == 0));
  }
  if (total == 0) return 0;
  const int grid = total < num_sms ? total : num_sms;
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
