// Sharded fused AdamW, optionally fused BEHIND a reduce-scatter over NVLink peer memory.
//
// One launch processes a list of segments (this rank's chunk of every parameter region of a
// bucket).  For every element of the chunk the kernel
//   1. sums the gradient over all peers' flat gradient buffers (peer pointers mapped through
//      symmetric memory: plain ld.global on NVLink addresses; n_peers == 1 is the local case),
//   2. applies the 1/(world*accum) scale and the elementwise clamp (optax.clip, reference
//      main_zero.py:161),
//   3. runs AdamW on the fp32 master weight / moments (reference main_zero.py:162-167),
//   4. writes the updated weight in bf16 straight into the local flat parameter buffer, which is
//      the source of the following all-gather.
// So the gradient is read exactly once (from every peer), never written back, and there is no
// separate unscale / cast / optimizer pass: 2*n_peers + 12 bytes read and 14 written per element.
#include "common.cuh"
#include "launchers.h"

namespace ztb {

constexpr int kAdamThreads = 128;  // small CTAs: <= 72 registers x 128 threads = 9 K registers, no shared memory, so one of them fits
                                   // next to a resident tcgen05 GEMM / attention CTA (those cap themselves at 48-54 K registers)
constexpr int kAdamVec = 8;                       // elements per thread per tile
constexpr int kAdamTile = kAdamThreads * kAdamVec;  // 1024
constexpr int kMaxPeers = 8;

struct PeerPtrs {
  const void* g[kMaxPeers];
};

// NVLS: one instruction returns the sum over every rank's copy of 8 bf16 values (the NVSwitch reduces; fp32
// accumulation, bf16 result); the store is replicated by the switch into every rank's buffer.
ZTB_DEVINL uint4 multimem_ld_reduce_bf16x8(const bf16* mc) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(mc) : "memory");
  return r;
}
ZTB_DEVINL void multimem_st_bf16x8(bf16* mc, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1, %2, %3, %4};"
               ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

ZTB_DEVINL uint32_t ld_acquire_sys(const int* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

ZTB_DEVINL void add_bf16x8(float* g, uint4 u) {
  float2 x0 = unpack_bf16(u.x), x1 = unpack_bf16(u.y), x2 = unpack_bf16(u.z), x3 = unpack_bf16(u.w);
  g[0] += x0.x; g[1] += x0.y; g[2] += x1.x; g[3] += x1.y; g[4] += x2.x; g[5] += x2.y; g[6] += x3.x; g[7] += x3.y;
}

// Two launch shapes of the same kernel:
//   wide        many CTAs, the whole GPU at HBM / NVLink speed (nothing else to run: the last buckets of a step);
//   background  at most one CTA per SM (grid <= SM count): co-resident with the GEMM and attention CTAs of the backward
//               pass that is still running, it soaks up their spare memory bandwidth instead of taking SMs away.
// Every CTA iteration covers kUnroll tiles: the gradient fetches (local, peer, or switch-reduced) of all of them are
// issued before the first is consumed -- a switch reduction has microseconds of latency, so the number of fetches in
// flight per SM, not the instruction rate, sets the throughput.  kMulti: sum over several peers' buffers with plain
// loads (fp32 accumulators, so fewer tiles in flight); otherwise one fetch per vector (local buffer or multicast).
template <bool kGradFp32, bool kMulti>
__global__ void __launch_bounds__(kAdamThreads, 7) adamw_shard_kernel(float* __restrict__ master, float* __restrict__ m,
                                                                      float* __restrict__ v, bf16* __restrict__ flat_param,
                                                                      PeerPtrs peers, int n_peers,
                                                                      const ZtbSegment* __restrict__ segs, int nseg,
                                                                      long long total_tiles, ZtbAdamArgs a,
                                                                      const int* __restrict__ wait_flags, int wait_value,
                                                                      const bf16* __restrict__ mc_grad, bf16* __restrict__ mc_param) {
  constexpr int kUnroll = (kGradFp32 || kMulti) ? 2 : 4;
  constexpr int kWords = kGradFp32 ? 8 : (kMulti ? 8 : 4);   // 32-bit registers holding one vector's gradient
  if (wait_flags != nullptr) {
    // every peer publishes "my gradients for this bucket are final" with a release store into OUR
    // signal pad; acquire it before touching peer memory.
    if (threadIdx.x < n_peers) {
      while ((int)ld_acquire_sys(wait_flags + threadIdx.x) < wait_value) __nanosleep(64);
    }
    __syncthreads();
  }
  const float gscale = a.scale_mul != nullptr ? a.grad_scale * __ldg(a.scale_mul) : a.grad_scale;
  const float lr = a.dyn != nullptr ? __ldg(a.dyn) : a.lr;
  const float inv_bc1 = a.dyn != nullptr ? __ldg(a.dyn + 1) : a.inv_bc1, inv_bc2 = a.dyn != nullptr ? __ldg(a.dyn + 2) : a.inv_bc2;
  for (long long t0 = (long long)blockIdx.x * kUnroll; t0 < total_tiles; t0 += (long long)gridDim.x * kUnroll) {
    int sgi[kUnroll];          // segment of tile t0 + k, or -1: nothing to do for this slot
    int e0[kUnroll];           // this thread's first element inside that segment's chunk (chunks are < 2^31 elements)
    uint32_t gw[kUnroll][kWords];
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
      const long long t = t0 + k;
      sgi[k] = -1; e0[k] = 0;
      if (t >= total_tiles) continue;
      int lo = 0, hi = nseg - 1;  // last segment with tile_begin <= t
      while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (segs[mid].tile_begin <= t) lo = mid; else hi = mid - 1;
      }
      const long long e = (t - segs[lo].tile_begin) * kAdamTile + (long long)threadIdx.x * kAdamVec;
      if (e >= segs[lo].len) continue;  // len is a multiple of 8 (chunks are 64-element aligned)
      sgi[k] = lo; e0[k] = (int)e;
    }
    // ---- gradient fetches of all kUnroll tiles
    if constexpr (kGradFp32) {
#pragma unroll
      for (int k = 0; k < kUnroll; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) gw[k][j] = 0u;   // +0.0f
      for (int p = 0; p < n_peers; ++p) {
        float4 x[kUnroll], y[kUnroll];
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) if (sgi[k] >= 0) {
          const float4* gp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(peers.g[p]) + segs[sgi[k]].flat_off + e0[k]);
          x[k] = gp[0]; y[k] = gp[1];
        }
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) if (sgi[k] >= 0) {
          float* g = reinterpret_cast<float*>(gw[k]);
          g[0] += x[k].x; g[1] += x[k].y; g[2] += x[k].z; g[3] += x[k].w; g[4] += y[k].x; g[5] += y[k].y; g[6] += y[k].z; g[7] += y[k].w;
        }
      }
    } else if constexpr (kMulti) {
#pragma unroll
      for (int k = 0; k < kUnroll; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) gw[k][j] = 0u;
      for (int p = 0; p < n_peers; ++p) {
        uint4 u[kUnroll];
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) if (sgi[k] >= 0) u[k] = ld_nc_v4(reinterpret_cast<const bf16*>(peers.g[p]) + segs[sgi[k]].flat_off + e0[k]);
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) if (sgi[k] >= 0) add_bf16x8(reinterpret_cast<float*>(gw[k]), u[k]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < kUnroll; ++k) if (sgi[k] >= 0) {
        const long long fo = segs[sgi[k]].flat_off + e0[k];
        uint4 u = mc_grad != nullptr ? multimem_ld_reduce_bf16x8(mc_grad + fo) : ld_nc_v4(reinterpret_cast<const bf16*>(peers.g[0]) + fo);
        gw[k][0] = u.x; gw[k][1] = u.y; gw[k][2] = u.z; gw[k][3] = u.w;
      }
    }
    // ---- AdamW on each tile
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
      if (sgi[k] < 0) continue;
      const ZtbSegment sg = segs[sgi[k]];
      const long long fo = sg.flat_off + e0[k], so = sg.shard_off + e0[k];
      const float wd = sg.decay ? a.wd : 0.f;
      float g[8];
      if constexpr (kGradFp32 || kMulti) {
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = __uint_as_float(gw[k][j]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) { float2 f = unpack_bf16(gw[k][j]); g[2 * j] = f.x; g[2 * j + 1] = f.y; }
      }
      float4* pm = reinterpret_cast<float4*>(master + so);
      float4* mm = reinterpret_cast<float4*>(m + so);
      float4* vv = reinterpret_cast<float4*>(v + so);
      float4 P[2] = {pm[0], pm[1]}, Mo[2] = {mm[0], mm[1]}, Vo[2] = {vv[0], vv[1]};
      float* pf = reinterpret_cast<float*>(P);
      float* mf = reinterpret_cast<float*>(Mo);
      float* vf = reinterpret_cast<float*>(Vo);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float gg = g[j] * gscale;
        if (a.clip > 0.f) gg = fminf(fmaxf(gg, -a.clip), a.clip);
        float mk = a.beta1 * mf[j] + a.omb1 * gg;
        float vk = a.beta2 * vf[j] + a.omb2 * gg * gg;
        float upd = (mk * inv_bc1) / (sqrtf(vk * inv_bc2) + a.eps) + wd * pf[j];
        pf[j] -= lr * upd;
        mf[j] = mk;
        vf[j] = vk;
      }
      pm[0] = P[0]; pm[1] = P[1]; mm[0] = Mo[0]; mm[1] = Mo[1]; vv[0] = Vo[0]; vv[1] = Vo[1];
      uint4 o = make_uint4(pack_bf16(pf[0], pf[1]), pack_bf16(pf[2], pf[3]), pack_bf16(pf[4], pf[5]), pack_bf16(pf[6], pf[7]));
      if (mc_param != nullptr) multimem_st_bf16x8(mc_param + fo, o);   // the all-gather: pushed to every rank by the switch
      else st_v4(flat_param + fo, o);
    }
  }
}

}  // namespace ztb

using namespace ztb;

// max_ctas > 0: the background shape (callers pass the SM count or less); 0: the wide shape.
int ztb_adamw_shard(float* master, float* m, float* v, void* flat_param, const void* const* grads, int n_peers,
                    int grad_is_fp32, const ZtbSegment* segs, int nseg, long long total_tiles, ZtbAdamArgs a,
                    const int* wait_flags, int wait_value, int sms, int max_ctas, const void* mc_grad_, void* mc_param_,
                    cudaStream_t st) {
  const bf16* mc_grad = (const bf16*)mc_grad_;
  bf16* mc_param = (bf16*)mc_param_;
  if (grad_is_fp32 && mc_grad != nullptr) { ztb_set_error("adamw_shard", "multicast reduce needs bf16 gradients", __FILE__, __LINE__); return -1; }
  if (n_peers < 1 || n_peers > kMaxPeers) { ztb_set_error("adamw_shard", "n_peers out of range", __FILE__, __LINE__); return -1; }
  if (total_tiles <= 0) return 0;
  PeerPtrs pp;
  for (int i = 0; i < kMaxPeers; ++i) pp.g[i] = i < n_peers ? grads[i] : nullptr;
  const bool multi = !grad_is_fp32 && mc_grad == nullptr && n_peers > 1;
  const int unroll = (grad_is_fp32 || multi) ? 2 : 4;
  long long units = (total_tiles + unroll - 1) / unroll;
  long long cap = max_ctas > 0 ? (long long)max_ctas : (long long)sms * 16;
  int grid = (int)(units < cap ? units : cap);
  // Shared-memory carveout: an SM cannot change its L1 / shared split while a CTA is resident.  These CTAs use no shared
  // memory, so by default they could pin an SM they land on (between two kernels of the backward pass) in a small-shared
  // configuration for their whole lifetime, and the next 225 KB tcgen05 CTA would have to wait for them.  Preferring the
  // maximum carveout makes an update CTA always compatible with the GEMM / attention CTAs it is meant to sit next to.
  static bool carveout_set = false;
  if (!carveout_set) {
    cudaFuncSetAttribute(adamw_shard_kernel<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(adamw_shard_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(adamw_shard_kernel<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    carveout_set = true;
  }
#define ZTB_ADAM_LAUNCH(FP32, MULTI) adamw_shard_kernel<FP32, MULTI><<<grid, kAdamThreads, 0, st>>>( \
    master, m, v, (bf16*)flat_param, pp, n_peers, segs, nseg, total_tiles, a, wait_flags, wait_value, mc_grad, mc_param)
  if (grad_is_fp32) ZTB_ADAM_LAUNCH(true, false);
  else if (multi) ZTB_ADAM_LAUNCH(false, true);
  else ZTB_ADAM_LAUNCH(false, false);
#undef ZTB_ADAM_LAUNCH
  ZTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
