// Attention backward on tcgen05: per (batch, head, 128-key tile) CTA, loop over 128-query tiles.
//   S  = Q K^T                (recomputed)          P  = exp(scale * S - lse)  (key padding -> 0)
//   dP = dO V^T                                     dZ = P o (dP - delta),  dS = scale * dZ
//   dV += P^T dO    dK += dS^T Q    dQ_i = dS K     (dQ accumulated over key tiles in fp32 global memory)
// Five MMAs per (query tile, key tile); every operand is a 128B-swizzled 16-bit tile in shared memory that is read
// either K-major or MN-major, so no transposed copies are ever made:
//   S : A = Q  [q x dh]  K-major      B = K  [kv x dh] K-major
//   dP: A = dO [q x dh]  K-major      B = V  [kv x dh] K-major
//   dV: A = P  [q x kv]  MN-major (M = kv)   B = dO [q x dh] MN-major (N = dh)   contraction over q
//   dK: A = dS [q x kv]  MN-major            B = Q  [q x dh] MN-major
//   dQ: A = dS [q x kv]  K-major             B = K  [kv x dh] MN-major (N = dh)  contraction over kv
// TMEM (512 columns): [0,128) S then dQ_i, [128,256) dP, [256,256+dh) dV, [384,384+dh) dK.
// Activations (Q, K, V, P) are fp16/bf16 per plan; gradients (dO, dS) are bf16.
// warp 0 lane 0: TMA + MMA issue; warps 1..4: softmax / gradient math (thread = tile row).
#include <math.h>

#include "backward.h"
#include "kernels.h"
#include "ptx.cuh"

namespace uv {

template <int DH>
struct AttnBwdCfg {
  static constexpr int kTile = 128 * DH * 2;  // Q, dO, K, V tiles: DH/64 boxes of [128 rows x 64]
  static constexpr int kPS = 32768;           // P and dS tiles [128 x 128] 16-bit: 2 boxes of [128 rows x 64]
  static constexpr int kSmemBytes = 1024 + 4 * kTile + 2 * kPS + 128 * 4 + 256;
  static constexpr uint32_t kTmemCols = 512;
};

// 32 fp32 accumulator columns of this thread's row -> 16-bit at dqkv16[elem_off ..] (single-key-tile fast path: the values are
// final, so they go straight to the operand buffer of the in-projection dgrad / wgrad GEMMs; the in_proj_bias gradient is a
// separate column-sum pass over that buffer - accumulating it here with a transposing warp reduction measured slower).
__device__ __forceinline__ void store16_colsum(const AttnBwdArgs& a, const uint32_t (&r)[32], bool valid, size_t elem_off, int col0,
                                               int lane) {
  (void)col0;
  (void)lane;
  if (valid) {
    uint16_t* dst = a.dqkv16 + elem_off;
#pragma unroll
    for (int e = 0; e < 32; e += 8)
      *reinterpret_cast<uint4*>(dst + e) =
          make_uint4(cvt16x2(__uint_as_float(r[e]), __uint_as_float(r[e + 1]), a.fmt_grad),
                     cvt16x2(__uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]), a.fmt_grad),
                     cvt16x2(__uint_as_float(r[e + 4]), __uint_as_float(r[e + 5]), a.fmt_grad),
                     cvt16x2(__uint_as_float(r[e + 6]), __uint_as_float(r[e + 7]), a.fmt_grad));
  }
}

template <int DH>
__global__ void __launch_bounds__(160, 1) attention_bwd_tcgen05_kernel(const __grid_constant__ AttnBwdArgs a) {
  using Cfg = AttnBwdCfg<DH>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + Cfg::kTile;
  uint8_t* sK = sdO + Cfg::kTile;
  uint8_t* sV = sK + Cfg::kTile;
  uint8_t* sP = sV + Cfg::kTile;
  uint8_t* sdS = sP + Cfg::kPS;
  float* s_bias = reinterpret_cast<float*>(sdS + Cfg::kPS);  // [128] 0 / -inf per key of this CTA's key tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 128);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;
  uint64_t* sdp_full = bars + 2;
  uint64_t* pds_full = bars + 3;
  uint64_t* it_done = bars + 4;
  uint64_t* dq_read = bars + 5;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x;  // key tile
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int L = a.L;
  const int num_q = (L + 127) / 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tm_qkv);
    tma_prefetch_desc(&a.tm_do);
    mbar_init(kv_full, 1);
    mbar_init(qdo_full, 1);
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, 4);
    mbar_init(it_done, 1);
    mbar_init(dq_read, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // Only now (this CTA owns its TMEM columns) may the next grid be scheduled: a dependent CTA that grabbed TMEM first and
  // then blocked in griddepcontrol.wait could starve a CTA of this grid sharing its SM.
  pdl_launch_dependents();
  pdl_wait();  // setup done; everything below reads the previous kernels' outputs
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t tm_s = tmem_base, tm_dp = tmem_base + 128, tm_dv = tmem_base + 256, tm_dk = tmem_base + 384;

  if (warp == 0) {
    if (lane == 0) {
      const int fa = a.fmt_act, fg = a.fmt_grad;
      const uint32_t id_s = make_idesc_f16_ab(128, 128, fa, fa, 0, 0);
      const uint32_t id_dp = make_idesc_f16_ab(128, 128, fg, fa, 0, 0);
      const uint32_t id_dv = make_idesc_f16_ab(128, DH, fa, fg, 1, 1);
      const uint32_t id_dk = make_idesc_f16_ab(128, DH, fg, fa, 1, 1);
      const uint32_t id_dq = make_idesc_f16_ab(128, DH, fg, fa, 0, 1);
      mbar_arrive_expect_tx(kv_full, 2 * Cfg::kTile);
#pragma unroll
      for (int kb = 0; kb < DH / 64; ++kb) {
        tma_load_2d(sK + kb * 16384, &a.tm_qkv, kv_full, a.d + h * DH + kb * 64, b * L + j * 128);
        tma_load_2d(sV + kb * 16384, &a.tm_qkv, kv_full, 2 * a.d + h * DH + kb * 64, b * L + j * 128);
      }
      for (int i = 0; i < num_q; ++i) {
        const uint32_t ph = i & 1;
        if (i > 0) mbar_wait(it_done, ph ^ 1);  // MMAs of the previous query tile retired: Q, dO, P, dS tiles are free
        mbar_arrive_expect_tx(qdo_full, 2 * Cfg::kTile);
#pragma unroll
        for (int kb = 0; kb < DH / 64; ++kb) {
          tma_load_2d(sQ + kb * 16384, &a.tm_qkv, qdo_full, h * DH + kb * 64, b * L + i * 128);
          tma_load_2d(sdO + kb * 16384, &a.tm_do, qdo_full, h * DH + kb * 64, b * L + i * 128);
        }
        if (i == 0) mbar_wait(kv_full, 0);
        mbar_wait(qdo_full, ph);
        if (i > 0) mbar_wait(dq_read, ph ^ 1);  // dQ_{i-1} has been drained from the S columns
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) {
          const uint32_t off = (ks / 4) * 16384 + (ks % 4) * 32;
          umma_f16_ss(tm_s, make_smem_desc_sw128(smem_u32(sQ) + off, 16, 1024), make_smem_desc_sw128(smem_u32(sK) + off, 16, 1024),
                      id_s, ks > 0 ? 1u : 0u);
        }
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) {
          const uint32_t off = (ks / 4) * 16384 + (ks % 4) * 32;
          umma_f16_ss(tm_dp, make_smem_desc_sw128(smem_u32(sdO) + off, 16, 1024),
                      make_smem_desc_sw128(smem_u32(sV) + off, 16, 1024), id_dp, ks > 0 ? 1u : 0u);
        }
        umma_commit(sdp_full);
        mbar_wait(pds_full, ph);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {  // contraction over the 128 query rows, 16 per step = two 1024 B atoms
          const uint32_t offk = ks * 2048;
          umma_f16_ss(tm_dv, make_smem_desc_sw128(smem_u32(sP) + offk, 16384, 1024),
                      make_smem_desc_sw128(smem_u32(sdO) + offk, 16384, 1024), id_dv, (i > 0 || ks > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t offk = ks * 2048;
          umma_f16_ss(tm_dk, make_smem_desc_sw128(smem_u32(sdS) + offk, 16384, 1024),
                      make_smem_desc_sw128(smem_u32(sQ) + offk, 16384, 1024), id_dk, (i > 0 || ks > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {  // contraction over the 128 keys
          const uint32_t offa = (ks / 4) * 16384 + (ks % 4) * 32;
          umma_f16_ss(tm_s, make_smem_desc_sw128(smem_u32(sdS) + offa, 16, 1024),
                      make_smem_desc_sw128(smem_u32(sK) + ks * 2048, 16384, 1024), id_dq, ks > 0 ? 1u : 0u);
        }
        umma_commit(it_done);
      }
    }
  } else {
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const int tid = threadIdx.x - 32;
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    const float c_log2e = 1.4426950408889634f;
    const float sc2 = a.scale * c_log2e;
    {
      const int key = j * 128 + tid;
      s_bias[tid] = (key < L && a.key_mask[(size_t)b * L + key] != 0.f) ? 0.f : -INFINITY;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const size_t ld3 = (size_t)3 * a.d;
    for (int i = 0; i < num_q; ++i) {
      const uint32_t ph = i & 1;
      const int qi = i * 128 + row;
      const bool qvalid = qi < L;
      float lse2 = 0.f, dlt = 0.f;
      if (qvalid) {
        lse2 = a.lse[((size_t)b * a.H + h) * L + qi] * c_log2e;
        dlt = a.delta[((size_t)b * a.H + h) * L + qi];
      }
      mbar_wait(sdp_full, ph);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t rs[32], rp[32];
        tmem_ld_32x32b_x32(tm_s + lane_addr + c * 32, rs);
        tmem_ld_32x32b_x32(tm_dp + lane_addr + c * 32, rp);
        tmem_ld_wait();
        uint32_t pk[16], dk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float p0 = 0.f, p1 = 0.f;
          if (qvalid) {
            p0 = exp2f(__uint_as_float(rs[e]) * sc2 + s_bias[c * 32 + e] - lse2);
            p1 = exp2f(__uint_as_float(rs[e + 1]) * sc2 + s_bias[c * 32 + e + 1] - lse2);
          }
          const float d0 = p0 * (__uint_as_float(rp[e]) - dlt) * a.scale;
          const float d1 = p1 * (__uint_as_float(rp[e + 1]) - dlt) * a.scale;
          pk[e / 2] = cvt16x2(p0, p1, a.fmt_act);
          dk[e / 2] = cvt16x2(p0 == 0.f ? 0.f : d0, p1 == 0.f ? 0.f : d1, a.fmt_grad);
        }
        uint8_t* prow = sP + (c / 2) * 16384 + row * 128;
        uint8_t* drow = sdS + (c / 2) * 16384 + row * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = ((c & 1) * 4 + q) ^ (row & 7);
          *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
          *reinterpret_cast<uint4*>(drow + chunk * 16) = make_uint4(dk[q * 4], dk[q * 4 + 1], dk[q * 4 + 2], dk[q * 4 + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      // dQ_i (this thread's query row) -> fp32 global
      mbar_wait(it_done, ph);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < DH / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tm_s + lane_addr + c * 32, r);
        tmem_ld_wait();
        if (a.dqkv16 != nullptr) {  // single key tile: final values, straight to the 16-bit GEMM operand (+ bias-gradient sums)
          store16_colsum(a, r, qvalid, ((size_t)b * L + qi) * ld3 + h * DH + c * 32, h * DH + c * 32, lane);
        } else if (qvalid) {
          float* dst = a.dqkv32 + ((size_t)b * L + qi) * ld3 + h * DH + c * 32;
          if (a.dq_atomic) {
#pragma unroll
            for (int e = 0; e < 32; ++e) atomicAdd(dst + e, __uint_as_float(r[e]));
          } else {
#pragma unroll
            for (int e = 0; e < 32; e += 4)
              *reinterpret_cast<float4*>(dst + e) = make_float4(__uint_as_float(r[e]), __uint_as_float(r[e + 1]),
                                                                __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_read);
    }
    // dV, dK of this key tile (thread = key row).  tcgen05.ld is warp-collective: loads are unconditional, stores predicated.
    const int kv = j * 128 + row;
    const bool kvalid = kv < L;
    float* dk_dst = a.dqkv32 + ((size_t)b * L + (kvalid ? kv : 0)) * ld3 + a.d + h * DH;
    float* dv_dst = a.dqkv32 + ((size_t)b * L + (kvalid ? kv : 0)) * ld3 + 2 * a.d + h * DH;
#pragma unroll
    for (int c = 0; c < DH / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tm_dv + lane_addr + c * 32, r);
      tmem_ld_wait();
      if (a.dqkv16 != nullptr) {
        store16_colsum(a, r, kvalid, ((size_t)b * L + (kvalid ? kv : 0)) * ld3 + 2 * a.d + h * DH + c * 32, 2 * a.d + h * DH + c * 32, lane);
      } else if (kvalid) {
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          *reinterpret_cast<float4*>(dv_dst + c * 32 + e) =
              make_float4(__uint_as_float(r[e]), __uint_as_float(r[e + 1]), __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
      }
      tmem_ld_32x32b_x32(tm_dk + lane_addr + c * 32, r);
      tmem_ld_wait();
      if (a.dqkv16 != nullptr) {
        store16_colsum(a, r, kvalid, ((size_t)b * L + (kvalid ? kv : 0)) * ld3 + a.d + h * DH + c * 32, a.d + h * DH + c * 32, lane);
      } else if (kvalid) {
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          *reinterpret_cast<float4*>(dk_dst + c * 32 + e) =
              make_float4(__uint_as_float(r[e]), __uint_as_float(r[e + 1]), __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

template <int DH>
static int launch_bwd_tc(const AttnBwdArgs& a, cudaStream_t stream) {
  using Cfg = AttnBwdCfg<DH>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_bwd_tcgen05_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attention_bwd): %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  dim3 grid((a.L + 127) / 128, a.H, a.B);
  launch_k(attention_bwd_tcgen05_kernel<DH>, dim3(grid), dim3(160), Cfg::kSmemBytes, stream, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) set_error("attention_bwd launch failed: %s", cudaGetErrorString(e));
  return (int)e;
}

int launch_attention_bwd(const AttnBwdArgs& a, cudaStream_t stream) {
  if (a.dh == 128) return launch_bwd_tc<128>(a, stream);
  if (a.dh == 64) return launch_bwd_tc<64>(a, stream);
  set_error("launch_attention_bwd: tensor-core path needs dh in {64,128}, got %d", a.dh);
  return (int)cudaErrorInvalidValue;
}

}  // namespace uv
