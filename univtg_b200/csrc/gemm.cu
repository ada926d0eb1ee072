// Persistent, warp-specialised tcgen05 GEMM for sm_100a with a fused, run-time configured epilogue.
//
//   C[M,N] = epilogue( sum_k A[m,k] * B[n,k] ),  A/B 16-bit (fp16 or bf16), fp32 accumulation in TMEM.
//
// This one kernel serves every dense contraction of the UniVTG hot path (SURVEY.md §2.2 rows K1-K3, K5, K7,
// K9-K11 and their backward passes): input projectors (model/univtg.py:399-406), the three attention
// in-projections + out-projection (torch MHA called at model/transformer_encoder_droppath.py:118), the FFN
// (:122) and the k=3 Conv1d heads (model/univtg.py:378-382) expressed as three row-shifted K segments.
//
// Roles (256 threads, 1 CTA / SM, persistent over a static round-robin tile schedule):
//   warp 0 lane 0 : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier expect_tx)
//   warp 1 lane 0 : MMA issuer     (tcgen05.mma cta_group::1, M=128, N=BN, K=16 x4 per 64-wide k-block)
//   warp 2        : TMEM allocator (2 accumulator stages of BN fp32 columns)
//   warps 4..7    : epilogue       (tcgen05.ld 32x32b -> smem transpose -> coalesced global stores)
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
  static constexpr int kEpiFloats = 4 * 32 * 33;         // per-warp 32x32 transpose buffers (padded)
  static constexpr int kSmemBytes = 1024 /*align slack*/ + kStages * kStageBytes + kEpiFloats * 4 + 4 * 32 * 8 + 256;
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
__global__ void __launch_bounds__(256, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmGroup g) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint8_t* stage_base = smem;
  float* epi_buf = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes);
  int* s_orow = reinterpret_cast<int*>(epi_buf + Cfg::kEpiFloats);        // [4][32]
  float* s_rscale = reinterpret_cast<float*>(s_orow + 4 * 32);            // [4][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_rscale + 4 * 32);
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
      mbar_init(&tmem_empty[s], 4);
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
    const int wq = warp & 3;  // TMEM lane quarter this warp may access
    float* buf = epi_buf + wq * (32 * 33);
    int* orow_s = s_orow + wq * 32;
    float* rscale_s = s_rscale + wq * 32;
    const int fmt = g.fmt;
    int as = 0;
    uint32_t aphase = 0;
    TileInfo ti;
    for (int t = blockIdx.x; decode_tile(g, BN, t, ti); t += gridDim.x) {
      const GemmProblem& pr = g.p[ti.p];
      const int m0 = ti.m_blk * GEMM_BM + wq * 32;
      const int n_base = ti.n_blk * BN;
      // ---- per-row bookkeeping (thread = row) ----
      const int m = m0 + lane;
      int b = 0, l = m;
      if (pr.rps_in > 0) {
        b = m / pr.rps_in;
        l = m - b * pr.rps_in;
      }
      const bool valid = m < pr.M;
      const bool sep = pr.zero_sep && (l == pr.rps_in - 1);
      const int orow = (pr.rps_in > 0 ? b * pr.rps_out + l : m) + pr.row_off;
      float rsc = pr.alpha;
      if (pr.row_scale != nullptr && valid) rsc *= pr.row_scale[b];
      if (sep) rsc = 0.f;
      orow_s[lane] = valid ? orow : -1;
      rscale_s[lane] = rsc;
      __syncwarp();

      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(wq * 32) << 16) + as * BN;
      const bool add_bias = (pr.bias != nullptr) && (ti.split == 0);
      const bool atomic = (pr.accumulate != 0) || (pr.ksplit > 1);

      for (int c = 0; c < BN / 32; ++c) {
        const int n0 = n_base + c * 32;
        if (n0 >= pr.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_addr + c * 32, r);
        tmem_ld_wait();
        if (pr.out16t != nullptr) {
          // transposed store (V^T for the attention PV product): thread = row, registers = columns
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = n0 + j;
              if (n < pr.N) {
                float v = __uint_as_float(r[j]);
                if (add_bias) v += __ldg(pr.bias + n);
                if (pr.act == ACT_RELU) v = fmaxf(v, 0.f);
                else if (pr.act == ACT_GELU) v = gelu_erf(v);
                v *= rsc;
                pr.out16t[((size_t)b * pr.N + n) * pr.ldt + l] = cvt16(v, fmt);
              }
            }
          }
          continue;
        }
        // ---- transpose through smem: afterwards lane = column, loop index = row ----
#pragma unroll
        for (int j = 0; j < 32; ++j) buf[lane * 33 + j] = __uint_as_float(r[j]);
        __syncwarp();
        const int n = n0 + lane;
        const bool ncol = n < pr.N;
        const float bias_v = (add_bias && ncol) ? __ldg(pr.bias + n) : 0.f;
#pragma unroll 4
        for (int rr = 0; rr < 32; ++rr) {
          const int orr = orow_s[rr];
          if (orr < 0 || !ncol) continue;
          float v = buf[rr * 33 + lane] + bias_v;
          if (pr.act == ACT_RELU) v = fmaxf(v, 0.f);
          else if (pr.act == ACT_GELU) v = gelu_erf(v);
          v *= rscale_s[rr];
          if (pr.resid != nullptr) v += pr.resid[(size_t)orr * pr.ld_resid + n];
          if (pr.out32 != nullptr) {
            float* dst = pr.out32 + (size_t)orr * pr.ld32 + n;
            if (atomic) atomicAdd(dst, v);
            else *dst = v;
          }
          const int mrow = m0 + rr;
          if (pr.out32_id != nullptr) pr.out32_id[(size_t)mrow * pr.ld32_id + n] = v;
          if (pr.out16 != nullptr) pr.out16[(size_t)orr * pr.ld16 + n] = cvt16(v, fmt);
          if (pr.out16p != nullptr) {
            const float pv = (pr.addtab != nullptr) ? pr.addtab[(size_t)mrow * pr.ld_addtab + n] : 0.f;
            pr.out16p[(size_t)orr * pr.ld16 + n] = cvt16(v + pv, fmt);
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
static int launch_bn(const GemmGroup& g, int num_sms, cudaStream_t stream) {
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
  }
  if (total == 0) return 0;
  const int grid = total < num_sms ? total : num_sms;
  gemm_tcgen05_kernel<BN><<<grid, 256, Cfg::kSmemBytes, stream>>>(g);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("gemm launch failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

int launch_gemm_group(const GemmGroup& g, int bn, int num_sms, cudaStream_t stream) {
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
