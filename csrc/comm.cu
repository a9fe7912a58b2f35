// NVLink / NVSwitch communication kernels over symmetric (peer-mapped) memory.  No NCCL.
// They replace the reference's collectives: the gradient pmean of src/training/xmap_train_functions.py:84
// (a reduce-scatter is all ZeRO-1 needs) and the parameter all-gather XLA inserts for
// main_zero.py:452-456; src/utils/pod_test.py:17-34 is the health probe ztb_p2p_copy serves.
//
//   ztb_signal            release-store a step counter into every peer's flag pad
//   ztb_allgather_pull    every rank pulls the other ranks' freshly updated bf16 parameter chunks
//                         straight out of their flat parameter buffers (16-byte ld.global over
//                         NVLink, all peers in flight at once -- NVSwitch gives each peer full
//                         bandwidth, so a flat all-to-all pull beats a ring) after acquiring the
//                         owners' "shard ready" flags
//   ztb_reduce_scatter    standalone pull reduce-scatter (bandwidth sweeps / tests; the training
//                         path uses the variant fused with AdamW in optim.cu)
//   ztb_p2p_copy          plain peer copy for the bandwidth probe
//
// Flag protocol: flags live in symmetric memory, one int per (kind, source rank).  A producer
// finishes its kernel, then ztb_signal (same stream) does __threadfence_system() +
// st.release.sys into each peer's pad; consumers spin with ld.acquire.sys at kernel start.  Values
// are monotonically increasing step numbers, so pads never need resetting.
#include <stdio.h>

#include "common.cuh"
#include "launchers.h"

namespace ztb {

constexpr int kMaxRanks = 8;
struct PtrList {
  void* p[kMaxRanks];
};

ZTB_DEVINL int ld_acquire_sys_i32(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
ZTB_DEVINL unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Failure detection: a peer that died (or a protocol bug) must not hang the GPU until the job's
// wall-clock limit.  Every cross-rank spin gives up after kFlagTimeoutNs and traps, which surfaces as a
// CUDA error on the host ("unspecified launch failure") instead of a silent hang.
constexpr unsigned long long kFlagTimeoutNs = 120ull * 1000000000ull;
ZTB_DEVINL void st_release_sys_i32(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// 1-warp spin-wait on a row of flags.  Cross-rank waits that overlap with compute run here (one
// tiny CTA) instead of inside the big kernels, so a stalled peer can never pin SM resources that
// this rank's own backward needs to make progress.
// `vptr` (optional): the step counter lives on the device (CUDA-graph replays cannot bake a new value into the launch);
// the awaited / published value is then *vptr + value.
__global__ void wait_flags_kernel(const int* flags, int n, int value, const int* vptr) {
  if (vptr != nullptr) value += *vptr;
  if ((int)threadIdx.x < n) {
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys_i32(flags + threadIdx.x) < value) {
      __nanosleep(128);
      if (globaltimer_ns() - t0 > kFlagTimeoutNs) {
        printf("[ztb] rank flag wait timed out: slot %d holds %d, expected >= %d (peer %d stalled or dead)\n", (int)threadIdx.x,
               ld_acquire_sys_i32(flags + threadIdx.x), value, (int)threadIdx.x);
        __trap();
      }
    }
  }
}

__global__ void signal_kernel(PtrList pads, int n, int slot, int value, const int* vptr) {
  if (vptr != nullptr) value += *vptr;
  if ((int)threadIdx.x < n) {
    __threadfence_system();
    st_release_sys_i32(reinterpret_cast<int*>(pads.p[threadIdx.x]) + slot, value);
  }
}

ZTB_DEVINL void wait_flags(const int* flags, int n, int value) {
  if (flags != nullptr) {
    if ((int)threadIdx.x < n) {
      while (ld_acquire_sys_i32(flags + threadIdx.x) < value) __nanosleep(64);
    }
    __syncthreads();
  }
}

// segs[i]: flat_off = element offset of a chunk (same offset on source and destination),
// len = elements, decay = source rank, tile_begin = prefix.  16-byte vectors, 4 in flight/thread.
constexpr int kCopyThreads = 256, kCopyTile = kCopyThreads * 8 * 4;  // 8192 elements (16 KB) per tile
__global__ void __launch_bounds__(kCopyThreads) allgather_pull_kernel(bf16* __restrict__ dst, PtrList src,
                                                                      const ZtbSegment* __restrict__ segs, int nseg,
                                                                      long long total_tiles, const int* flags, int nflags,
                                                                      int wait_value) {
  wait_flags(flags, nflags, wait_value);
  for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    int lo = 0, hi = nseg - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (segs[mid].tile_begin <= t) lo = mid; else hi = mid - 1;
    }
    const ZtbSegment sg = segs[lo];
    const bf16* s = reinterpret_cast<const bf16*>(src.p[sg.decay]) + sg.flat_off;
    bf16* d = dst + sg.flat_off;
    long long base = (t - sg.tile_begin) * kCopyTile;
    uint4 v[4];
    long long e[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      e[k] = base + ((long long)k * kCopyThreads + threadIdx.x) * 8;
      if (e[k] < sg.len) v[k] = ld_nc_v4(s + e[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (e[k] < sg.len) st_v4(d + e[k], v[k]);
  }
}

// out[shard_off + i] = sum_p grads[p][flat_off + i]  (fp32 accumulate, bf16 wire format)
__global__ void __launch_bounds__(256) reduce_scatter_kernel(bf16* __restrict__ out, PtrList grads, int n_peers,
                                                             const ZtbSegment* __restrict__ segs, int nseg,
                                                             long long total_tiles, const int* flags, int wait_value) {
  wait_flags(flags, n_peers, wait_value);
  for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    int lo = 0, hi = nseg - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (segs[mid].tile_begin <= t) lo = mid; else hi = mid - 1;
    }
    const ZtbSegment sg = segs[lo];
    long long e0 = (t - sg.tile_begin) * 2048 + (long long)threadIdx.x * 8;  // segment table packed with tile = 2048
    if (e0 >= sg.len) continue;
    float g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = 0; p < n_peers; ++p) {
      uint4 u = ld_nc_v4(reinterpret_cast<const bf16*>(grads.p[p]) + sg.flat_off + e0);
      float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
      g[0] += a.x; g[1] += a.y; g[2] += b.x; g[3] += b.y; g[4] += c.x; g[5] += c.y; g[6] += d.x; g[7] += d.y;
    }
    st_v4(out + sg.shard_off + e0, make_uint4(pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]), pack_bf16(g[4], g[5]), pack_bf16(g[6], g[7])));
  }
}

__global__ void __launch_bounds__(256) p2p_copy_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n16) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = ld_nc_v4(src + i);
}

}  // namespace ztb

using namespace ztb;

// The flag kernels can be resident for a long time (a wait spins until the slowest peer arrives) on an SM that the
// backward pass needs for a 225 KB tcgen05 CTA, and an SM cannot change its L1 / shared split while a CTA is resident:
// make them prefer the maximum shared-memory carveout so they never pin an SM in an incompatible configuration.
static void prefer_max_carveout_once() {
  static bool done = false;
  if (done) return;
  cudaFuncSetAttribute(wait_flags_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(signal_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(allgather_pull_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(reduce_scatter_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  done = true;
}

int ztb_wait_flags(const int* flags, int n, int value, cudaStream_t st, const int* vptr) {
  prefer_max_carveout_once();
  wait_flags_kernel<<<1, 32, 0, st>>>(flags, n, value, vptr);
  ZTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int ztb_signal(void* const* pads, int n, int slot, int value, cudaStream_t st, const int* vptr) {
  if (n < 1 || n > kMaxRanks) { ztb_set_error("ztb_signal", "bad rank count", __FILE__, __LINE__); return -1; }
  PtrList pl;
  for (int i = 0; i < kMaxRanks; ++i) pl.p[i] = i < n ? pads[i] : nullptr;
  prefer_max_carveout_once();
  signal_kernel<<<1, 32, 0, st>>>(pl, n, slot, value, vptr);
  ZTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int ztb_allgather_pull(void* dst, void* const* srcs, int world, const ZtbSegment* segs, int nseg, long long total_tiles,
                       const int* flags, int wait_value, int sms, cudaStream_t st) {
  if (total_tiles <= 0) return 0;
  PtrList pl;
  for (int i = 0; i < kMaxRanks; ++i) pl.p[i] = i < world ? srcs[i] : nullptr;
  long long want = total_tiles < (long long)sms * 4 ? total_tiles : (long long)sms * 4;
  prefer_max_carveout_once();
  allgather_pull_kernel<<<(int)want, kCopyThreads, 0, st>>>((bf16*)dst, pl, segs, nseg, total_tiles, flags, world, wait_value);
  ZTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int ztb_reduce_scatter(void* out, void* const* grads, int n_peers, const ZtbSegment* segs, int nseg, long long total_tiles,
                       const int* flags, int wait_value, int sms, cudaStream_t st) {
  if (total_tiles <= 0) return 0;
  PtrList pl;
  for (int i = 0; i < kMaxRanks; ++i) pl.p[i] = i < n_peers ? grads[i] : nullptr;
  long long want = total_tiles < (long long)sms * 8 ? total_tiles : (long long)sms * 8;
  prefer_max_carveout_once();
  reduce_scatter_kernel<<<(int)want, 256, 0, st>>>((bf16*)out, pl, n_peers, segs, nseg, total_tiles, flags, wait_value);
  ZTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int ztb_p2p_copy(void* dst, const void* src, size_t bytes, int sms, cudaStream_t st) {
  p2p_copy_kernel<<<sms * 4, 256, 0, st>>>((uint4*)dst, (const uint4*)src, bytes / 16);
  ZTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
