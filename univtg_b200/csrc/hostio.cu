// Host-side batch assembly for the packed feature shards (univtg_b200/data.py).  No device code here:
// the reference does this work per sample in Python (main/dataset.py:644-696 loads, utils/tensor_utils.py:6-53 pad_sequences_1d);
// at > 50 k pairs/s per GPU the collate of one batch (13.5 MB of fp16 rows at the cfg2 shape) has ~0.5 ms, so it is a plain
// multi-threaded gather: per sample one memcpy of its video rows and one of its query rows out of the memory-mapped shard into
// the pinned staging buffers, zero fill of the padding, float masks.
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/univtg_b200.h"
#include "kernels.h"

namespace {

struct AssembleJob {
  uint16_t *dst_vid, *dst_txt;
  float *dst_vmask, *dst_tmask;
  const uint16_t *src_vid, *src_txt;
  const int64_t *vid_row0, *txt_row0;  // first shard row of every sample's video / query
  const int32_t *vid_len, *txt_len;    // rows to copy (already clipped to Lv / Lt)
  int B, Lv, Lt, Dv, Dt;
};

void assemble_range(const AssembleJob& j, int b0, int b1) {
  for (int b = b0; b < b1; ++b) {
    const int nv = j.vid_len[b], nt = j.txt_len[b];
    uint16_t* dv = j.dst_vid + (size_t)b * j.Lv * j.Dv;
    uint16_t* dt = j.dst_txt + (size_t)b * j.Lt * j.Dt;
    memcpy(dv, j.src_vid + (size_t)j.vid_row0[b] * j.Dv, (size_t)nv * j.Dv * 2);
    if (nv < j.Lv) memset(dv + (size_t)nv * j.Dv, 0, (size_t)(j.Lv - nv) * j.Dv * 2);
    memcpy(dt, j.src_txt + (size_t)j.txt_row0[b] * j.Dt, (size_t)nt * j.Dt * 2);
    if (nt < j.Lt) memset(dt + (size_t)nt * j.Dt, 0, (size_t)(j.Lt - nt) * j.Dt * 2);
    float* mv = j.dst_vmask + (size_t)b * j.Lv;
    float* mt = j.dst_tmask + (size_t)b * j.Lt;
    for (int l = 0; l < j.Lv; ++l) mv[l] = l < nv ? 1.f : 0.f;
    for (int l = 0; l < j.Lt; ++l) mt[l] = l < nt ? 1.f : 0.f;
  }
}

// A small persistent pool: workers sleep on a condition variable between batches (thread creation per batch would cost more
// than the copy itself).
class Pool {
 public:
  explicit Pool(int n) : stop_(false), gen_(0), pending_(0) {
    for (int i = 0; i < n; ++i) workers_.emplace_back([this, i] { run(i); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      ++gen_;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int size() const { return (int)workers_.size(); }
  void run_job(const AssembleJob& job) {
    std::unique_lock<std::mutex> lk(mu_);
    job_ = job;
    next_.store(0);
    pending_ = (int)workers_.size();
    ++gen_;
    cv_.notify_all();
    done_.wait(lk, [this] { return pending_ == 0; });
  }

 private:
  void run(int) {
    uint64_t seen = 0;
    for (;;) {
      AssembleJob job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        job = job_;
      }
      for (;;) {  // samples are dealt out one at a time: rows differ in length
        const int b = next_.fetch_add(1);
        if (b >= job.B) break;
        assemble_range(job, b, b + 1);
      }
      std::lock_guard<std::mutex> lk(mu_);
      if (--pending_ == 0) done_.notify_all();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  bool stop_;
  uint64_t gen_;
  int pending_;
  std::atomic<int> next_;
  AssembleJob job_;
};

Pool* g_pool = nullptr;
std::mutex g_pool_mu;

}  // namespace

extern "C" int univtg_host_assemble_batch(void* dst_vid, void* dst_txt, float* dst_vmask, float* dst_tmask, const void* src_vid,
                                          const void* src_txt, const int64_t* vid_row0, const int64_t* txt_row0, const int32_t* vid_len,
                                          const int32_t* txt_len, int32_t B, int32_t Lv, int32_t Lt, int32_t Dv, int32_t Dt,
                                          int32_t threads) {
  if (!dst_vid || !dst_txt || !dst_vmask || !dst_tmask || !src_vid || !src_txt || !vid_row0 || !txt_row0 || !vid_len || !txt_len ||
      B < 1 || Lv < 1 || Lt < 1 || Dv < 1 || Dt < 1) {
    uv::set_error("univtg_host_assemble_batch: bad argument");
    return 1;
  }
  for (int b = 0; b < B; ++b)
    if (vid_len[b] < 0 || vid_len[b] > Lv || txt_len[b] < 0 || txt_len[b] > Lt) {
      uv::set_error("univtg_host_assemble_batch: sample %d has %d / %d rows for a [%d, %d] batch", b, vid_len[b], txt_len[b], Lv, Lt);
      return 1;
    }
  AssembleJob job{reinterpret_cast<uint16_t*>(dst_vid), reinterpret_cast<uint16_t*>(dst_txt), dst_vmask, dst_tmask,
                  reinterpret_cast<const uint16_t*>(src_vid), reinterpret_cast<const uint16_t*>(src_txt), vid_row0, txt_row0, vid_len,
                  txt_len, B, Lv, Lt, Dv, Dt};
  if (threads <= 1) {
    assemble_range(job, 0, B);
    return 0;
  }
  std::lock_guard<std::mutex> lk(g_pool_mu);  // one batch at a time per process (the loader has one producer thread)
  if (g_pool == nullptr || g_pool->size() != threads) {
    delete g_pool;
    g_pool = new Pool(threads);
  }
  g_pool->run_job(job);
  return 0;
}

// Direct path: the shard's memory mapping is registered with the driver (page-locked, univtg_host_register), so every sample's rows
// are DMA'd straight from the page cache into the device batch - no staging copy on the CPU at all.  Per batch: two async copies
// per sample, memsets for the padding, one small copy for the masks (built on the host into `mask_stage`, pinned).
extern "C" int univtg_host_register(void* base, size_t bytes, int32_t enable) {
  cudaError_t e;
  if (enable) {
    e = cudaHostRegister(base, bytes, cudaHostRegisterPortable | cudaHostRegisterReadOnly);
    if (e != cudaSuccess) {
      cudaGetLastError();
      e = cudaHostRegister(base, bytes, cudaHostRegisterPortable);  // writable (copy-on-write) mappings
    }
  } else {
    e = cudaHostUnregister(base);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    uv::set_error("univtg_host_register(%p, %zu, %d): %s", base, bytes, enable, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

extern "C" int univtg_h2d_gather_batch(void* dev_vid, void* dev_txt, float* dev_vmask, float* dev_tmask, float* mask_stage,
                                       const void* src_vid, const void* src_txt, const int64_t* vid_row0, const int64_t* txt_row0,
                                       const int32_t* vid_len, const int32_t* txt_len, int32_t B, int32_t Lv, int32_t Lt, int32_t Dv,
                                       int32_t Dt, void* stream) {
  if (!dev_vid || !dev_txt || !dev_vmask || !dev_tmask || !mask_stage || !src_vid || !src_txt || !vid_row0 || !txt_row0 || !vid_len ||
      !txt_len || B < 1) {
    uv::set_error("univtg_h2d_gather_batch: bad argument");
    return 1;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const uint16_t* sv = reinterpret_cast<const uint16_t*>(src_vid);
  const uint16_t* sx = reinterpret_cast<const uint16_t*>(src_txt);
  uint16_t* dv = reinterpret_cast<uint16_t*>(dev_vid);
  uint16_t* dx = reinterpret_cast<uint16_t*>(dev_txt);
  cudaError_t e = cudaSuccess;
  for (int b = 0; b < B && e == cudaSuccess; ++b) {
    const int nv = vid_len[b], nt = txt_len[b];
    if (nv < 0 || nv > Lv || nt < 0 || nt > Lt) {
      uv::set_error("univtg_h2d_gather_batch: sample %d has %d / %d rows for a [%d, %d] batch", b, nv, nt, Lv, Lt);
      return 1;
    }
    e = cudaMemcpyAsync(dv + (size_t)b * Lv * Dv, sv + (size_t)vid_row0[b] * Dv, (size_t)nv * Dv * 2, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && nv < Lv) e = cudaMemsetAsync(dv + ((size_t)b * Lv + nv) * Dv, 0, (size_t)(Lv - nv) * Dv * 2, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dx + (size_t)b * Lt * Dt, sx + (size_t)txt_row0[b] * Dt, (size_t)nt * Dt * 2, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && nt < Lt) e = cudaMemsetAsync(dx + ((size_t)b * Lt + nt) * Dt, 0, (size_t)(Lt - nt) * Dt * 2, st);
    float* mv = mask_stage + (size_t)b * Lv;
    float* mt = mask_stage + (size_t)B * Lv + (size_t)b * Lt;
    for (int l = 0; l < Lv; ++l) mv[l] = l < nv ? 1.f : 0.f;
    for (int l = 0; l < Lt; ++l) mt[l] = l < nt ? 1.f : 0.f;
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(dev_vmask, mask_stage, (size_t)B * Lv * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dev_tmask, mask_stage + (size_t)B * Lv, (size_t)B * Lt * 4, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) {
    uv::set_error("univtg_h2d_gather_batch: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}
