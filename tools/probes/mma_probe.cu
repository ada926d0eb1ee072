// Stand-alone microbenchmark (not product code): issue rate of tcgen05.mma.cta_group::1.kind::f16 (M = 128, K = 16) as a
// function of N, of the number of independent TMEM accumulators, of the commit cadence and of HOW the instruction is issued:
//   style 0: `if (lane == 0) { loop }`  - a divergent single-lane region (what gemm.cu did in round 1; ptxas wraps every
//            UTCHMMA / UTCBAR in an ELECT + BRA.U.ANY "waterfall" because it cannot prove the operands warp-uniform)
//   style 1: the whole warp runs the loop convergently and each MMA / commit is guarded by elect.sync (CUTLASS's pattern)
// Operands stay resident in shared memory (no TMA in the loop).  Prints one JSON object per configuration:
//   ns and SM cycles per MMA (globaltimer and clock64 of the issuing warp), so the SM clock under load is visible too.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I univtg_b200/csrc tools/probes/mma_probe.cu -o tools/probes/mma_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#include "ptx.cuh"

using namespace uv;

__global__ void __launch_bounds__(128, 1) mma_probe_kernel(int n, int n_acc, int per_commit, int iters, int style, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (threadIdx.x == 0) {
    mbar_init(&bar, style == 2 ? (uint32_t)iters : 8u);  // style 2: one phase collects every commit - the issuer never drains the pipe
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc<512>(&holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = holder;
  const uint32_t sa = smem_u32(smem), sb = sa + 16384;
  const uint32_t idesc = make_idesc_f16_ab(128, n, 0, 0, 0, 0);
  const uint64_t da0 = make_smem_desc_sw128(sa, 16, 1024), db0 = make_smem_desc_sw128(sb, 16, 1024);
  const int acc_stride = 512 / n_acc;  // columns between accumulators (n <= acc_stride is the caller's job)
  unsigned long long t0 = 0, t1 = 0;
  long long c0 = 0, c1 = 0;
  if (style == 0) {
    if (threadIdx.x == 0) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      c0 = clock64();
      uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        const uint32_t d = tmem + (uint32_t)((it % n_acc) * acc_stride);
        for (int k = 0; k < per_commit; ++k) {
          const uint64_t off = (uint64_t)((k & 3) * 2);  // 32 B inside the swizzle span, in 16-byte units
          umma_f16_ss(d, da0 + off, db0 + off, idesc, 1u);
        }
        umma_commit(&bar);
        if ((it & 7) == 7) {
          mbar_wait(&bar, phase);
          phase ^= 1;
        }
      }
      c1 = clock64();
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    }
  } else {
    if (threadIdx.x < 32) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      c0 = clock64();
      uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        const uint32_t d = tmem + (uint32_t)((it % n_acc) * acc_stride);
        for (int k = 0; k < per_commit; ++k) {
          const uint64_t off = (uint64_t)((k & 3) * 2);
          if (elect_one()) umma_f16_ss(d, da0 + off, db0 + off, idesc, 1u);
        }
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
        if (style == 1 && (it & 7) == 7) {
          mbar_wait(&bar, phase);
          phase ^= 1;
        }
      }
      if (style == 2) mbar_wait(&bar, 0);
      c1 = clock64();
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    }
  }
  if (threadIdx.x == 0) {
    const float total = (float)((long long)iters * per_commit);
    out[2 * blockIdx.x] = (float)(t1 - t0) / total;
    out[2 * blockIdx.x + 1] = (float)(c1 - c0) / total;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

static float median(std::vector<float> v) {
  std::sort(v.begin(), v.end());
  return v[v.size() / 2];
}

int main() {
  cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024);
  float* d_out;
  cudaMalloc(&d_out, 2 * 148 * sizeof(float));
  std::vector<float> h(2 * 148);
  printf("[\n");
  bool first = true;
  for (int style = 0; style < 3; ++style)
    for (int blocks : {148})
      for (int n : {64, 128, 192, 256})
        for (int n_acc : {1, 2})
          for (int pc : {1, 2, 4, 8, 16, 64}) {
            if (n * n_acc > 512) continue;
            const int iters = 4096 / pc * 8;  // 32768 MMAs (style 2: <= 32768 pending arrivals, below the mbarrier count limit)
            for (int rep = 0; rep < 3; ++rep) mma_probe_kernel<<<blocks, 128, 50 * 1024>>>(n, n_acc, pc, iters, style, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
              printf("{\"error\": \"%s\"}]\n", cudaGetErrorString(e));
              return 1;
            }
            cudaMemcpy(h.data(), d_out, 2 * blocks * sizeof(float), cudaMemcpyDeviceToHost);
            std::vector<float> ns, cyc;
            for (int b = 0; b < blocks; ++b) {
              ns.push_back(h[2 * b]);
              cyc.push_back(h[2 * b + 1]);
            }
            printf("%s{\"style\": %d, \"blocks\": %d, \"n\": %d, \"n_acc\": %d, \"per_commit\": %d, \"ns_per_mma\": %.1f, \"cyc_per_mma\": %.1f}",
                   first ? "" : ",\n", style, blocks, n, n_acc, pc, median(ns), median(cyc));
            first = false;
          }
  printf("\n]\n");
  return 0;
}
