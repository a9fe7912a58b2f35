// Kernel launch helper with programmatic dependent launch (PDL).
//
// Every kernel of the training step is launched with cudaLaunchAttributeProgrammaticStreamSerialization: the grid may be
// scheduled while its predecessor in the stream is still draining, runs its prologue (barrier init, TMEM allocation,
// tensor-map prefetch, index math), and blocks in `griddepcontrol.wait` until the predecessor has completed and its
// memory is visible -- that hides the ~2.5 us launch latency that otherwise sits between two dependent kernels
// (~1100 launches per 1.3B step).  A kernel lets ITS successor start early with `griddepcontrol.launch_dependents`.
// Rules every participating kernel follows: no global-memory access before pdl_wait(); pdl_wait() is executed by every
// thread that later touches global memory.  Kernels launched without the attribute (torch's own, the communication
// stream's) are unaffected: for them both instructions are no-ops / the dependency stays a full stream dependency.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_set>
#include <utility>

// An SM cannot change its L1 / shared-memory split while any CTA is resident.  The background optimizer CTAs and the flag
// waits live for milliseconds next to the backward pass, whose tcgen05 kernels need the maximum (228 KB) carveout, so THOSE
// long-lived kernels prefer the maximum carveout (force = true): they never pin an SM in a small-shared configuration.
// Doing the same for every kernel (ZTB_CARVEOUT_ALL=1) was measured within +-1 ms of this default on the 1.3B step
// (profiles/bench_r2_ab_call{3,10}.jsonl); it stays off so the streaming kernels keep their L1.
inline bool ztb_carveout_all() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("ZTB_CARVEOUT_ALL"); v = e ? (atoi(e) != 0) : 0; }
  return v != 0;
}
inline void ztb_prefer_max_smem(const void* fn, bool force = false) {
  if (!force && !ztb_carveout_all()) return;
  static std::mutex mu;
  static std::unordered_set<const void*> done;
  std::lock_guard<std::mutex> lock(mu);
  if (done.insert(fn).second) cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

inline bool ztb_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ZTB_PDL");
    v = e ? (atoi(e) != 0) : 1;   // measured -3 ms per 1.3B step alone, -8 ms together with the bf16 dQ path (profiles/bench_r2_ab.jsonl)
  }
  return v != 0;
}

template <typename... P, typename... A>
inline cudaError_t ztb_launch(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  ztb_prefer_max_smem(reinterpret_cast<const void*>(kernel));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = ztb_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(std::forward<A>(args))...);
}

#ifdef __CUDACC__
// wait for the predecessor grid (complete + memory visible); then allow the successor to be scheduled
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_wait(); pdl_launch_dependents(); }
#endif
