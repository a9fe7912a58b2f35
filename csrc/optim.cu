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

constexpr int kAdamThreads = 128;  // small CTAs (8 K registers) co-reside with the attention / elementwise kernels of the
                                   // overlapping backward pass instead of waiting for a free SM
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

// kThreads == kAdamThreads: the wide form, many small CTAs filling the GPU.
// kThreads == 1024: the "side" form -- a handful of CTAs (launched as clusters of 2, one CTA per SM) that own a
// few SMs for the whole update while persistent GEMMs, sized to leave exactly those SMs free, keep the rest.
// A 1024-thread CTA behaves as 8 virtual 128-thread CTAs.
template <bool kGradFp32, int kThreads>
__global__ void __launch_bounds__(kThreads) adamw_shard_kernel(float* __restrict__ master, float* __restrict__ m,
                                                                   float* __restrict__ v, bf16* __restrict__ flat_param,
                                                                   PeerPtrs peers, int n_peers,
                                                                   const ZtbSegment* __restrict__ segs, int nseg,
                                                                   long long total_tiles, ZtbAdamArgs a,
                                                                   const int* __restrict__ wait_flags, int wait_value,
                                                                   const bf16* __restrict__ mc_grad, bf16* __restrict__ mc_param) {
  if (wait_flags != nullptr) {
    // every peer publishes "my gradients for this bucket are final" with a release store into OUR
    // signal pad; acquire it before touching peer memory.
    if (threadIdx.x < n_peers) {
      while ((int)ld_acquire_sys(wait_flags + threadIdx.x) < wait_value) __nanosleep(64);
    }
    __syncthreads();
  }
  const float gscale = a.scale_mul != nullptr ? a.grad_scale * __ldg(a.scale_mul) : a.grad_scale;
  constexpr int kSub = kThreads / kAdamThreads;
  const int vlane = threadIdx.x % kAdamThreads;
  const long long vblock = (long long)blockIdx.x * kSub + threadIdx.x / kAdamThreads, vgrid = (long long)gridDim.x * kSub;
  for (long long t = vblock; t < total_tiles; t += vgrid) {
    int lo = 0, hi = nseg - 1;  // last segment with tile_begin <= t
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (segs[mid].tile_begin <= t) lo = mid; else hi = mid - 1;
    }
    const ZtbSegment sg = segs[lo];
    long long e0 = (t - sg.tile_begin) * kAdamTile + (long long)vlane * kAdamVec;
    if (e0 >= sg.len) continue;  // len is a multiple of 8 (chunks are 64-element aligned)
    long long fo = sg.flat_off + e0, so = sg.shard_off + e0;
    float g[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] = 0.f;
    if (!kGradFp32 && mc_grad != nullptr) {
      uint4 u = multimem_ld_reduce_bf16x8(mc_grad + fo);
      float2 x0 = unpack_bf16(u.x), x1 = unpack_bf16(u.y), x2 = unpack_bf16(u.z), x3 = unpack_bf16(u.w);
      g[0] = x0.x; g[1] = x0.y; g[2] = x1.x; g[3] = x1.y; g[4] = x2.x; g[5] = x2.y; g[6] = x3.x; g[7] = x3.y;
    } else
    for (int p = 0; p < n_peers; ++p) {
      if constexpr (kGradFp32) {
        const float4* gp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(peers.g[p]) + fo);
        float4 x = gp[0], y = gp[1];
        g[0] += x.x; g[1] += x.y; g[2] += x.z; g[3] += x.w; g[4] += y.x; g[5] += y.y; g[6] += y.z; g[7] += y.w;
      } else {
        uint4 u = ld_nc_v4(reinterpret_cast<const bf16*>(peers.g[p]) + fo);
        float2 x0 = unpack_bf16(u.x), x1 = unpack_bf16(u.y), x2 = unpack_bf16(u.z), x3 = unpack_bf16(u.w);
        g[0] += x0.x; g[1] += x0.y; g[2] += x1.x; g[3] += x1.y; g[4] += x2.x; g[5] += x2.y; g[6] += x3.x; g[7] += x3.y;
      }
    }
    float4* pm = reinterpret_cast<float4*>(master + so);
    float4* mm = reinterpret_cast<float4*>(m + so);
    float4* vv = reinterpret_cast<float4*>(v + so);
    float4 P[2] = {pm[0], pm[1]}, Mo[2] = {mm[0], mm[1]}, Vo[2] = {vv[0], vv[1]};
    float* pf = reinterpret_cast<float*>(P);
    float* mf = reinterpret_cast<float*>(Mo);
    float* vf = reinterpret_cast<float*>(Vo);
    const float wd = sg.decay ? a.wd : 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float gg = g[k] * gscale;
      if (a.clip > 0.f) gg = fminf(fmaxf(gg, -a.clip), a.clip);
      float mk = a.beta1 * mf[k] + (1.f - a.beta1) * gg;
      float vk = a.beta2 * vf[k] + (1.f - a.beta2) * gg * gg;
      float upd = (mk / a.bc1) / (sqrtf(vk / a.bc2) + a.eps) + wd * pf[k];
      pf[k] -= a.lr * upd;
      mf[k] = mk;
      vf[k] = vk;
    }
    pm[0] = P[0]; pm[1] = P[1]; mm[0] = Mo[0]; mm[1] = Mo[1]; vv[0] = Vo[0]; vv[1] = Vo[1];
    uint4 o = make_uint4(pack_bf16(pf[0], pf[1]), pack_bf16(pf[2], pf[3]), pack_bf16(pf[4], pf[5]), pack_bf16(pf[6], pf[7]));
    if (mc_param != nullptr) multimem_st_bf16x8(mc_param + fo, o);   // the all-gather: pushed to every rank by the switch
    else st_v4(flat_param + fo, o);
  }
}

}  // namespace ztb

using namespace ztb;

constexpr int kSideThreads = 1024;
constexpr int kSideSmem = 120 * 1024;   // more than half an SM's shared memory: one side CTA per SM

template <bool kGradFp32>
static int launch_side(float* master, float* m, float* v, bf16* flat_param, const PeerPtrs& pp, int n_peers, const ZtbSegment* segs,
                       int nseg, long long total_tiles, const ZtbAdamArgs& a, const int* wait_flags, int wait_value, const bf16* mc_grad, bf16* mc_param,
                       int ctas, cudaStream_t st) {
  auto kern = adamw_shard_kernel<kGradFp32, kSideThreads>;
  static bool configured = false;
  if (!configured) {
    ZTB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSideSmem));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)ctas, 1, 1);
  cfg.blockDim = dim3(kSideThreads, 1, 1);
  cfg.dynamicSmemBytes = kSideSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;   // a pair lands on one TPC: the same granularity the 2-CTA GEMM needs
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  ZTB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, master, m, v, flat_param, pp, n_peers, segs, nseg, total_tiles, a, wait_flags, wait_value, mc_grad, mc_param));
  return 0;
}

// side_sms > 0: run on exactly that many SMs (rounded down to an even number, >= 2) next to concurrently running
// persistent kernels instead of spreading over the whole GPU.
int ztb_adamw_shard(float* master, float* m, float* v, void* flat_param, const void* const* grads, int n_peers,
                    int grad_is_fp32, const ZtbSegment* segs, int nseg, long long total_tiles, ZtbAdamArgs a,
                    const int* wait_flags, int wait_value, int sms, int side_sms, const void* mc_grad_, void* mc_param_,
                    cudaStream_t st) {
  const bf16* mc_grad = (const bf16*)mc_grad_;
  bf16* mc_param = (bf16*)mc_param_;
  if (grad_is_fp32 && mc_grad != nullptr) { ztb_set_error("adamw_shard", "multicast reduce needs bf16 gradients", __FILE__, __LINE__); return -1; }
  if (n_peers < 1 || n_peers > kMaxPeers) { ztb_set_error("adamw_shard", "n_peers out of range", __FILE__, __LINE__); return -1; }
  if (total_tiles <= 0) return 0;
  PeerPtrs pp;
  for (int i = 0; i < kMaxPeers; ++i) pp.g[i] = i < n_peers ? grads[i] : nullptr;
  if (side_sms > 0) {
    int ctas = side_sms < 2 ? 2 : (side_sms & ~1);
    return grad_is_fp32 ? launch_side<true>(master, m, v, (bf16*)flat_param, pp, n_peers, segs, nseg, total_tiles, a, wait_flags, wait_value, mc_grad, mc_param, ctas, st)
                        : launch_side<false>(master, m, v, (bf16*)flat_param, pp, n_peers, segs, nseg, total_tiles, a, wait_flags, wait_value, mc_grad, mc_param, ctas, st);
  }
  long long want = total_tiles < (long long)sms * 16 ? total_tiles : (long long)sms * 16;
  int grid = (int)want;
  if (grad_is_fp32)
    adamw_shard_kernel<true, kAdamThreads><<<grid, kAdamThreads, 0, st>>>(master, m, v, (bf16*)flat_param, pp, n_peers, segs, nseg, total_tiles, a, wait_flags, wait_value, mc_grad, mc_param);
  else
    adamw_shard_kernel<false, kAdamThreads><<<grid, kAdamThreads, 0, st>>>(master, m, v, (bf16*)flat_param, pp, n_peers, segs, nseg, total_tiles, a, wait_flags, wait_value, mc_grad, mc_param);
  ZTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
