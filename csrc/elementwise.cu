// Memory-bound sm_100a kernels: LayerNorm (bias-free) fwd/bwd, dropout+residual fwd/bwd,
// embedding gather / scatter-add, GELU backward, cross-entropy fwd/bwd over bf16 logits.
// All are single-pass over HBM with 16-byte vector accesses; row reductions stay in
// registers + warp shuffles.  Reference semantics: SURVEY 2.4-A rows K2, K3, K9, K11 -- embedding gather
// src/models/GPT.py:75-83, bias-free LayerNorm GPT.py:42,49,98, dropout + residual layers.py:77,191 and
// GPT.py:43-49, tanh-GELU layers.py:68, shifted cross entropy GPT.py:105-111 + src/utils/losses.py:9-23.
#include "common.cuh"
#include "launch.cuh"
#include "launchers.h"

namespace ztb {

struct LnBwdExtra;
// Last-block epilogue of LayerNorm backward (shared by the fast and the generic kernel): see layernorm_bwd_kernel.
ZTB_DEVINL void ln_bwd_finalize(float* dw, int D, const LnBwdExtra& ex, bool* is_last);

// ============================================================== LayerNorm
// One warp per row, the whole row in registers: D = 256 * NV elements, lane l owns
// 16-byte vectors v*32 + l  (v < NV).  Stats in fp32.
template <int NV>
__global__ void __launch_bounds__(128) layernorm_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                            bf16* __restrict__ y, float* __restrict__ mean,
                                                            float* __restrict__ rstd, int M, float eps) {
  pdl_enter();
  constexpr int D = NV * 256;
  int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= M) return;
  int lane = threadIdx.x & 31;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)row * D);
  float v[NV * 8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    uint4 u = ld_nc_v4(xr + i * 32 + lane);
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    v[i * 8 + 0] = a.x; v[i * 8 + 1] = a.y; v[i * 8 + 2] = b.x; v[i * 8 + 3] = b.y;
    v[i * 8 + 4] = c.x; v[i * 8 + 5] = c.y; v[i * 8 + 6] = d.x; v[i * 8 + 7] = d.y;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[i * 8 + k];
  }
  float mu = warp_sum(s) * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < NV * 8; ++k) { float d = v[k] - mu; q += d * d; }
  float rs = rsqrtf(warp_sum(q) * (1.f / D) + eps);
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  uint4* yr = reinterpret_cast<uint4*>(y + (size_t)row * D);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    uint4 g = ld_v4(wr + i * 32 + lane);
    float2 g0 = unpack_bf16(g.x), g1 = unpack_bf16(g.y), g2 = unpack_bf16(g.z), g3 = unpack_bf16(g.w);
    uint4 o;
    o.x = pack_bf16((v[i * 8 + 0] - mu) * rs * g0.x, (v[i * 8 + 1] - mu) * rs * g0.y);
    o.y = pack_bf16((v[i * 8 + 2] - mu) * rs * g1.x, (v[i * 8 + 3] - mu) * rs * g1.y);
    o.z = pack_bf16((v[i * 8 + 4] - mu) * rs * g2.x, (v[i * 8 + 5] - mu) * rs * g2.y);
    o.w = pack_bf16((v[i * 8 + 6] - mu) * rs * g3.x, (v[i * 8 + 7] - mu) * rs * g3.y);
    st_v4(yr + i * 32 + lane, o);
  }
}

// Backward: dx = rstd * (dy*w - mean(dy*w) - xhat * mean(dy*w*xhat));  dw += sum_rows dy*xhat.
// Persistent grid: each warp strides over rows and keeps its dw partial in registers; one smem reduction + fp32 atomics
// per block into a small persistent scratch vector, and the LAST block to finish folds the scratch into the weight's
// gradient buffer (bf16 or fp32, overwrite or accumulate) and leaves scratch + ticket zeroed for the next launch -- no
// memset, no separate add/copy kernel (those were 3 launches per LayerNorm backward).
// Optional second output `dmask`: dropout-masked copy of dx (x is the output of a dropout+residual epilogue, whose
// backward needs mask(dx)/keep as the operand of its dgrad/wgrad GEMMs): same Philox mapping as dropout_add_kernel.
struct LnBwdExtra {
  void* grad_out;        // [D] gradient of the LayerNorm weight
  int grad_fp32;         // dtype of grad_out
  int grad_accumulate;   // += instead of =
  unsigned int* ticket;  // finished-block counter (zero on entry, zero on exit)
  float* partials;       // optional [gridDim.x, D]: deterministic mode -- per-block partial sums, added up in block order by
                         // the last block instead of meeting through fp32 atomics (whose order changes run to run)
  bf16* dmask;           // optional [M, D]
  const unsigned char* maskbits;   // optional [M, D / 8] keep bits from the forward epilogue (else the mask is regenerated with Philox)
  float p, inv_keep;
  const uint64_t* rng;
  uint32_t op_id;
};

ZTB_DEVINL void ln_bwd_finalize(float* dw, int D, const LnBwdExtra& ex, bool* is_last) {
  if (ex.ticket == nullptr) return;       // caller reads the fp32 scratch itself
  __threadfence();                        // this block's atomics are performed before its ticket
  __syncthreads();
  if (threadIdx.x == 0) *is_last = atomicAdd(ex.ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!*is_last) return;
  __threadfence();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float s;
    if (ex.partials != nullptr) {
      s = 0.f;
      for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(ex.partials + (size_t)b * D + c);
    } else {
      s = __ldcg(dw + c);
      dw[c] = 0.f;
    }
    if (ex.grad_fp32) {
      float* g = reinterpret_cast<float*>(ex.grad_out) + c;
      *g = ex.grad_accumulate ? *g + s : s;
    } else {
      bf16* g = reinterpret_cast<bf16*>(ex.grad_out) + c;
      *g = __float2bfloat16(ex.grad_accumulate ? __bfloat162float(*g) + s : s);
    }
  }
  if (threadIdx.x == 0) *ex.ticket = 0u;
}

// DROP: 0 = no masked copy, 1 = keep bits from the forward epilogue, 2 = mask regenerated with Philox.  A template parameter,
// not a run-time branch: with the mask code in the loop body ptxas no longer batched the row's 16 loads ahead of the math and the
// plain kernel ran at half its speed (118 us vs 56 us per call at the 1.3B shape).
template <int NV, int DROP>
__global__ void __launch_bounds__(128) layernorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                            const bf16* __restrict__ w, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, bf16* __restrict__ dx,
                                                            float* __restrict__ dw, int M, const bf16* __restrict__ dres,
                                                            const LnBwdExtra ex) {
  pdl_enter();
  constexpr int D = NV * 256;
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[NV * 8];
#pragma unroll
  for (int k = 0; k < NV * 8; ++k) acc[k] = 0.f;
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  uint64_t stream = 0;
  uint32_t thr = 0;
  Philox ph(0);
  if constexpr (DROP == 2) { ph = Philox(ex.rng[0]); stream = ex.rng[1] * 65536ull + ex.op_id; thr = dropout_threshold(ex.p); }
  for (int row = blockIdx.x * 4 + warp; row < M; row += gridDim.x * 4) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)row * D);
    const uint4* gr = reinterpret_cast<const uint4*>(dy + (size_t)row * D);
    float mu = mean[row], rs = rstd[row];
    // Two passes over the row with the packed bf16 x / dy held in registers in between (2 x 8*NV of them).  Re-loading them
    // in the second pass instead (192 instead of 253 registers) was measured at 118 us vs 56 us per call at D = 2048: with
    // 8 warps per SM the second round of load latency is fully exposed (profiles/ln_bench_r2_call2.jsonl).
    uint4 px[NV], pg[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      px[i] = ld_nc_v4(xr + i * 32 + lane);
      pg[i] = ld_nc_v4(gr + i * 32 + lane);
      uint4 uw = ld_v4(wr + i * 32 + lane);
      uint32_t xs[4] = {px[i].x, px[i].y, px[i].z, px[i].w}, gs[4] = {pg[i].x, pg[i].y, pg[i].z, pg[i].w};
      uint32_t ws[4] = {uw.x, uw.y, uw.z, uw.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 xf = unpack_bf16(xs[k]), gf = unpack_bf16(gs[k]), wf = unpack_bf16(ws[k]);
        int j = i * 8 + 2 * k;
        float xh0 = (xf.x - mu) * rs, xh1 = (xf.y - mu) * rs;
        acc[j] += gf.x * xh0; acc[j + 1] += gf.y * xh1;
        float gw0 = gf.x * wf.x, gw1 = gf.y * wf.y;
        s1 += gw0 + gw1;
        s2 += gw0 * xh0 + gw1 * xh1;
      }
    }
    s1 = warp_sum(s1) * (1.f / D);
    s2 = warp_sum(s2) * (1.f / D);
    uint4* dr = reinterpret_cast<uint4*>(dx + (size_t)row * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      uint4 uw = ld_v4(wr + i * 32 + lane);
      uint4 ur = make_uint4(0, 0, 0, 0);  // gradient arriving through the residual branch (fused add)
      if (dres != nullptr) ur = ld_nc_v4(reinterpret_cast<const uint4*>(dres + (size_t)row * D) + i * 32 + lane);
      uint32_t xs[4] = {px[i].x, px[i].y, px[i].z, px[i].w}, gs[4] = {pg[i].x, pg[i].y, pg[i].z, pg[i].w};
      uint32_t ws[4] = {uw.x, uw.y, uw.z, uw.w}, rr[4] = {ur.x, ur.y, ur.z, ur.w}, os[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 xf = unpack_bf16(xs[k]), gf = unpack_bf16(gs[k]), wf = unpack_bf16(ws[k]), rf = unpack_bf16(rr[k]);
        float xh0 = (xf.x - mu) * rs, xh1 = (xf.y - mu) * rs;
        os[k] = pack_bf16(rs * (gf.x * wf.x - s1 - xh0 * s2) + rf.x, rs * (gf.y * wf.y - s1 - xh1 * s2) + rf.y);
      }
      st_v4(dr + i * 32 + lane, make_uint4(os[0], os[1], os[2], os[3]));
      if constexpr (DROP != 0) {  // masked copy of the (bf16-rounded) dx, exactly what dropout_add_kernel<false> would produce from it
        const size_t vi = (size_t)row * (D / 8) + i * 32 + lane;
        uint32_t keep8, ms[4];
        if constexpr (DROP == 1) {
          keep8 = ld_nc_u8(ex.maskbits + vi);   // 8 keep bits recorded by the forward GEMM epilogue: no RNG work here (asm load: may be hoisted above the stores, batching with the row's other loads)
        } else {
          uint4 r0 = ph.rounds<kPhiloxRounds>(2 * vi, stream), r1 = ph.rounds<kPhiloxRounds>(2 * vi + 1, stream);
          uint32_t rb[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
          keep8 = 0;
#pragma unroll
          for (int k = 0; k < 8; ++k) keep8 |= (rb[k] >= thr ? 1u : 0u) << k;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float2 df = unpack_bf16(os[k]);
          ms[k] = pack_bf16((keep8 >> (2 * k)) & 1u ? df.x * ex.inv_keep : 0.f, (keep8 >> (2 * k + 1)) & 1u ? df.y * ex.inv_keep : 0.f);
        }
        st_v4(reinterpret_cast<uint4*>(ex.dmask + (size_t)row * D) + i * 32 + lane, make_uint4(ms[0], ms[1], ms[2], ms[3]));
      }
    }
  }
  __shared__ float red[4][NV * 256];
  __shared__ bool is_last;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) red[warp][(i * 32 + lane) * 8 + k] = acc[i * 8 + k];
  __syncthreads();
  if (ex.partials != nullptr) {
    for (int c = threadIdx.x; c < D; c += 128) ex.partials[(size_t)blockIdx.x * D + c] = red[0][c] + red[1][c] + red[2][c] + red[3][c];
  } else {
    for (int c = threadIdx.x; c < D; c += 128) atomicAdd(dw + c, red[0][c] + red[1][c] + red[2][c] + red[3][c]);
  }
  ln_bwd_finalize(dw, D, ex, &is_last);
}

// Generic-D fallback (D not a multiple of 256, e.g. the 64-wide test model): one block per row.
__global__ void layernorm_fwd_generic(const bf16* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ y,
                                      float* __restrict__ mean, float* __restrict__ rstd, int M, int D, float eps) {
  __shared__ float sh[2][4];
  int row = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bf16* xr = x + (size_t)row * D;
  float s = 0.f, q = 0.f;
  for (int c = threadIdx.x; c < D; c += 128) { float v = __bfloat162float(xr[c]); s += v; q += v * v; }
  s = warp_sum(s); q = warp_sum(q);
  if (lane == 0) { sh[0][warp] = s; sh[1][warp] = q; }
  __syncthreads();
  s = sh[0][0] + sh[0][1] + sh[0][2] + sh[0][3];
  q = sh[1][0] + sh[1][1] + sh[1][2] + sh[1][3];
  float mu = s / D, rs = rsqrtf(fmaxf(q / D - mu * mu, 0.f) + eps);
  if (threadIdx.x == 0) { mean[row] = mu; rstd[row] = rs; }
  for (int c = threadIdx.x; c < D; c += 128)
    y[(size_t)row * D + c] = __float2bfloat16((__bfloat162float(xr[c]) - mu) * rs * __bfloat162float(w[c]));
}
__global__ void layernorm_bwd_generic(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ w,
                                      const float* __restrict__ mean, const float* __restrict__ rstd,
                                      bf16* __restrict__ dx, float* __restrict__ dw, int M, int D, const bf16* __restrict__ dres,
                                      const LnBwdExtra ex) {
  __shared__ float sh[2][4];
  __shared__ bool is_last;
  int row = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float mu = mean[row], rs = rstd[row], s1 = 0.f, s2 = 0.f;
  for (int c = threadIdx.x; c < D; c += 128) {
    float xh = (__bfloat162float(x[(size_t)row * D + c]) - mu) * rs;
    float g = __bfloat162float(dy[(size_t)row * D + c]);
    float gw = g * __bfloat162float(w[c]);
    s1 += gw; s2 += gw * xh;
    atomicAdd(dw + c, g * xh);
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if (lane == 0) { sh[0][warp] = s1; sh[1][warp] = s2; }
  __syncthreads();
  s1 = (sh[0][0] + sh[0][1] + sh[0][2] + sh[0][3]) / D;
  s2 = (sh[1][0] + sh[1][1] + sh[1][2] + sh[1][3]) / D;
  for (int c = threadIdx.x; c < D; c += 128) {
    float xh = (__bfloat162float(x[(size_t)row * D + c]) - mu) * rs;
    float gw = __bfloat162float(dy[(size_t)row * D + c]) * __bfloat162float(w[c]);
    float extra = dres ? __bfloat162float(dres[(size_t)row * D + c]) : 0.f;
    bf16 o = __float2bfloat16(rs * (gw - s1 - xh * s2) + extra);
    dx[(size_t)row * D + c] = o;
    if (ex.dmask != nullptr) {   // same mask as dropout_add_kernel: element e -> Philox counter e / 4, word e % 4
      const size_t e = (size_t)row * D + c;
      bool keep;
      if (ex.maskbits != nullptr) {
        keep = (ex.maskbits[e >> 3] >> (e & 7)) & 1u;
      } else {
        Philox ph(ex.rng[0]);
        uint4 r4 = ph.rounds<kPhiloxRounds>(e >> 2, ex.rng[1] * 65536ull + ex.op_id);
        uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
        keep = rw[e & 3] >= dropout_threshold(ex.p);
      }
      ex.dmask[e] = __float2bfloat16(keep ? __bfloat162float(o) * ex.inv_keep : 0.f);
    }
  }
  ln_bwd_finalize(dw, D, ex, &is_last);
}

// ============================================================== dropout + residual
// fwd: out = res + keep(idx) * y / (1-p);  bwd: dy = keep(idx) * dout / (1-p)  (dres = dout).
// idx = flat element index; Philox counter idx/4, component idx%4 -- the same mapping the
// GEMM epilogue uses, so a fused forward and this backward agree on the mask.
template <bool kHasRes>
__global__ void __launch_bounds__(256) dropout_add_kernel(const bf16* __restrict__ y, const bf16* __restrict__ res,
                                                          bf16* __restrict__ out, size_t n8, float p, float inv_keep,
                                                          const uint64_t* __restrict__ rng, uint32_t op_id) {
  pdl_enter();
  uint64_t seed = rng[0], stream = rng[1] * 65536ull + op_id;
  Philox ph(seed);
  uint32_t thr = dropout_threshold(p);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    uint4 uy = ld_nc_v4(reinterpret_cast<const uint4*>(y) + i);
    uint4 r0 = ph.rounds<kPhiloxRounds>(2 * i, stream), r1 = ph.rounds<kPhiloxRounds>(2 * i + 1, stream);
    uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    uint32_t ys[4] = {uy.x, uy.y, uy.z, uy.w};
    uint32_t rs[4] = {0, 0, 0, 0};
    if (kHasRes) { uint4 ur = ld_nc_v4(reinterpret_cast<const uint4*>(res) + i); rs[0] = ur.x; rs[1] = ur.y; rs[2] = ur.z; rs[3] = ur.w; }
    uint32_t os[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 yf = unpack_bf16(ys[k]);
      float2 rf = kHasRes ? unpack_bf16(rs[k]) : make_float2(0.f, 0.f);
      float a = (rr[2 * k] >= thr) ? yf.x * inv_keep : 0.f;
      float b = (rr[2 * k + 1] >= thr) ? yf.y * inv_keep : 0.f;
      os[k] = pack_bf16(rf.x + a, rf.y + b);
    }
    st_v4(reinterpret_cast<uint4*>(out) + i, make_uint4(os[0], os[1], os[2], os[3]));
  }
}

// ============================================================== embedding
__global__ void __launch_bounds__(128) embedding_fwd_kernel(const int* __restrict__ idx, const bf16* __restrict__ wte,
                                                            bf16* __restrict__ out, int M, int D8, int V) {
  pdl_enter();
  int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= M) return;
  int t = idx[row];
  t = t < 0 ? 0 : (t >= V ? V - 1 : t);
  const uint4* src = reinterpret_cast<const uint4*>(wte) + (size_t)t * D8;
  uint4* dst = reinterpret_cast<uint4*>(out) + (size_t)row * D8;
  for (int c = threadIdx.x & 31; c < D8; c += 32) st_v4(dst + c, ld_v4(src + c));
}
// scatter-add into the (bf16 or fp32) gradient of the tied embedding: red.global.add on bf16x2
template <typename G>
__global__ void __launch_bounds__(128) embedding_bwd_kernel(const int* __restrict__ idx, const bf16* __restrict__ dout,
                                                            G* __restrict__ dwte, int M, int D, int V) {
  int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= M) return;
  int t = idx[row];
  if (t < 0 || t >= V) return;
  const bf162* src = reinterpret_cast<const bf162*>(dout + (size_t)row * D);
  for (int c = threadIdx.x & 31; c < D / 2; c += 32) {
    bf162 v = src[c];
    if constexpr (sizeof(G) == 2) {
      atomicAdd(reinterpret_cast<bf162*>(dwte + (size_t)t * D) + c, v);
    } else {
      float2 f = __bfloat1622float2(v);
      atomicAdd(reinterpret_cast<float*>(dwte + (size_t)t * D) + 2 * c, f.x);
      atomicAdd(reinterpret_cast<float*>(dwte + (size_t)t * D) + 2 * c + 1, f.y);
    }
  }
}

// Deterministic scatter-add for real corpora: the token ids are sorted first (host side: one torch.sort of M ints), every
// run of equal ids is cut into chunks of kEmbChunk rows, one warp sums a chunk in fp32 and adds it to the gradient row once.
// A token that occurs once per chunk-run (the common case) is a plain read-modify-write; only tokens with more than
// kEmbChunk occurrences meet through atomics, with one rounding per 32 occurrences instead of one per occurrence.
// (The plain kernel above rounds to bf16 on every occurrence: a frequent token's later increments vanish once the running
// sum is ~256x the addend.)
constexpr int kEmbChunk = 32;
template <typename G>
__global__ void __launch_bounds__(128) embedding_bwd_sorted_kernel(const int* __restrict__ sorted_tok, const int* __restrict__ order,
                                                                   const bf16* __restrict__ dout, G* __restrict__ dwte, int M, int D, int V) {
  const int p = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (p >= M) return;
  const int tok = sorted_tok[p];
  if (tok < 0 || tok >= V) return;
  int lo = 0, hi = p;                       // first position holding `tok`
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (sorted_tok[mid] < tok) lo = mid + 1; else hi = mid;
  }
  if ((p - lo) % kEmbChunk != 0) return;    // not a chunk start
  int end = p + 1;
  while (end < M && end < p + kEmbChunk && sorted_tok[end] == tok) ++end;
  const bool shared_row = p != lo || (end < M && sorted_tok[end] == tok);   // another chunk adds into the same row
  G* grow = dwte + (size_t)tok * D;
  for (int c = lane; c < D / 8; c += 32) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int q = p; q < end; ++q) {
      uint4 u = ld_nc_v4(dout + (size_t)order[q] * D + c * 8);
      float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), cc = unpack_bf16(u.z), d = unpack_bf16(u.w);
      acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y; acc[4] += cc.x; acc[5] += cc.y; acc[6] += d.x; acc[7] += d.y;
    }
    if constexpr (sizeof(G) == 2) {
      bf16* g = reinterpret_cast<bf16*>(grow) + c * 8;
      if (shared_row) {
#pragma unroll
        for (int k = 0; k < 4; ++k) atomicAdd(reinterpret_cast<bf162*>(g) + k, __floats2bfloat162_rn(acc[2 * k], acc[2 * k + 1]));
      } else {
        uint4 o = ld_v4(g);
        float2 a = unpack_bf16(o.x), b = unpack_bf16(o.y), cc = unpack_bf16(o.z), d = unpack_bf16(o.w);
        st_v4(g, make_uint4(pack_bf16(a.x + acc[0], a.y + acc[1]), pack_bf16(b.x + acc[2], b.y + acc[3]),
                            pack_bf16(cc.x + acc[4], cc.y + acc[5]), pack_bf16(d.x + acc[6], d.y + acc[7])));
      }
    } else {
      float* g = reinterpret_cast<float*>(grow) + c * 8;
      if (shared_row) {
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(g + k, acc[k]);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] += acc[k];
      }
    }
  }
}

// ============================================================== GELU backward
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const bf16* __restrict__ dg, const bf16* __restrict__ z,
                                                       bf16* __restrict__ dz, size_t n8) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    uint4 a = ld_nc_v4(reinterpret_cast<const uint4*>(dg) + i), b = ld_nc_v4(reinterpret_cast<const uint4*>(z) + i);
    uint32_t as[4] = {a.x, a.y, a.z, a.w}, bs[4] = {b.x, b.y, b.z, b.w}, os[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 g = unpack_bf16(as[k]), zz = unpack_bf16(bs[k]);
      os[k] = pack_bf16(g.x * gelu_tanh_grad(zz.x), g.y * gelu_tanh_grad(zz.y));
    }
    st_v4(reinterpret_cast<uint4*>(dz) + i, make_uint4(os[0], os[1], os[2], os[3]));
  }
}
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const bf16* __restrict__ z, bf16* __restrict__ g, size_t n8) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    uint4 b = ld_nc_v4(reinterpret_cast<const uint4*>(z) + i);
    uint32_t bs[4] = {b.x, b.y, b.z, b.w}, os[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 zz = unpack_bf16(bs[k]);
      os[k] = pack_bf16(gelu_tanh(zz.x), gelu_tanh(zz.y));
    }
    st_v4(reinterpret_cast<uint4*>(g) + i, make_uint4(os[0], os[1], os[2], os[3]));
  }
}

// ============================================================== cross entropy
// fwd: one block per row, online (max, sum-exp) in a single pass over the bf16 logits.
// Writes lse[row] and nll[row] = lse - logit[label] (0 when label < 0 = ignored row).
__global__ void __launch_bounds__(256) ce_fwd_kernel(const bf16* __restrict__ logits, const int* __restrict__ labels,
                                                     float* __restrict__ lse, float* __restrict__ nll, int V, size_t ld) {
  __shared__ float shm[8], shs[8];
  pdl_enter();
  int row = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bf16* lr = logits + (size_t)row * ld;
  float mx = -INFINITY, sm = 0.f;
  int V8 = V / 8;
  const float L2E = 1.4426950408889634f;
  for (int i = threadIdx.x; i < V8; i += 256) {
    uint4 u = ld_nc_v4(reinterpret_cast<const uint4*>(lr) + i);
    uint32_t us[4] = {u.x, u.y, u.z, u.w};
    float f[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) { float2 t = unpack_bf16(us[k]); f[2 * k] = t.x; f[2 * k + 1] = t.y; }
    float m2 = f[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) m2 = fmaxf(m2, f[k]);
    if (m2 > mx) { sm *= exp2f((mx - m2) * L2E); mx = m2; }
#pragma unroll
    for (int k = 0; k < 8; ++k) sm += exp2f((f[k] - mx) * L2E);
  }
  for (int c = V8 * 8 + threadIdx.x; c < V; c += 256) {
    float f = __bfloat162float(lr[c]);
    if (f > mx) { sm *= exp2f((mx - f) * L2E); mx = f; }
    sm += exp2f((f - mx) * L2E);
  }
  float wm = warp_max(mx);
  sm *= (mx == -INFINITY) ? 0.f : exp2f((mx - wm) * L2E);
  sm = warp_sum(sm);
  if (lane == 0) { shm[warp] = wm; shs[warp] = sm; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float M = shm[0];
    for (int k = 1; k < 8; ++k) M = fmaxf(M, shm[k]);
    float S = 0.f;
    for (int k = 0; k < 8; ++k) S += shs[k] * exp2f((shm[k] - M) * L2E);
    float l = M + logf(S);
    lse[row] = l;
    int lab = labels[row];
    nll[row] = (lab >= 0 && lab < V) ? l - __bfloat162float(lr[lab]) : 0.f;
  }
}
// bwd (in place): logits <- (softmax - onehot) * gscale[0] * inv_count ; ignored rows -> 0.
__global__ void __launch_bounds__(256) ce_bwd_kernel(bf16* __restrict__ logits, const int* __restrict__ labels,
                                                     const float* __restrict__ lse, const float* __restrict__ gscale,
                                                     float inv_count, int V, size_t ld) {
  pdl_enter();
  int row = blockIdx.x;
  bf16* lr = logits + (size_t)row * ld;
  int lab = labels[row];
  float l = lse[row];
  float g = (lab >= 0 && lab < V) ? gscale[0] * inv_count : 0.f;
  const float L2E = 1.4426950408889634f;
  int V8 = V / 8;
  for (int i = threadIdx.x; i < V8; i += 256) {
    uint4 u = ld_v4(reinterpret_cast<const uint4*>(lr) + i);
    uint32_t us[4] = {u.x, u.y, u.z, u.w}, os[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 t = unpack_bf16(us[k]);
      int c = i * 8 + 2 * k;
      float a = exp2f((t.x - l) * L2E) - (c == lab ? 1.f : 0.f);
      float b = exp2f((t.y - l) * L2E) - (c + 1 == lab ? 1.f : 0.f);
      os[k] = pack_bf16(a * g, b * g);
    }
    st_v4(reinterpret_cast<uint4*>(lr) + i, make_uint4(os[0], os[1], os[2], os[3]));
  }
  for (int c = V8 * 8 + threadIdx.x; c < V; c += 256) {
    float t = __bfloat162float(lr[c]);
    lr[c] = __float2bfloat16((exp2f((t - l) * L2E) - (c == lab ? 1.f : 0.f)) * g);
  }
}

}  // namespace ztb

using namespace ztb;

static int grid_for(size_t n, int block, int max_blocks) {
  size_t g = (n + block - 1) / block;
  return (int)(g < (size_t)max_blocks ? (g ? g : 1) : max_blocks);
}

#define LN_DISPATCH2(NV, KERNEL, X, GRID, ...)                                                 \
  switch (NV) {                                                                                \
    case 1: ztb_launch(KERNEL<1, X>, dim3(GRID), dim3(128), 0, st, __VA_ARGS__); break;        \
    case 2: ztb_launch(KERNEL<2, X>, dim3(GRID), dim3(128), 0, st, __VA_ARGS__); break;        \
    case 3: ztb_launch(KERNEL<3, X>, dim3(GRID), dim3(128), 0, st, __VA_ARGS__); break;        \
    case 4: ztb_launch(KERNEL<4, X>, dim3(GRID), dim3(128), 0, st, __VA_ARGS__); break;        \
    case 6: ztb_launch(KERNEL<6, X>, dim3(GRID), dim3(128), 0, st, __VA_ARGS__); break;        \
    case 8: ztb_launch(KERNEL<8, X>, dim3(GRID), dim3(128), 0, st, __VA_ARGS__); break;        \
  }

#define LN_DISPATCH(NV, KERNEL, GRID, ...)                                                     \
  switch (NV) {                                                                                \
    case 1: ztb_launch(KERNEL<1>, dim3(GRID), dim3(128), 0, st, __VA_ARGS__); break;           \
    case 2: ztb_launch(KERNEL<2>, dim3(GRID), dim3(128), 0, st, __VA_ARGS__); break;           \
    case 3: ztb_launch(KERNEL<3>, dim3(GRID), dim3(128), 0, st, __VA_ARGS__); break;           \
    case 4: ztb_launch(KERNEL<4>, dim3(GRID), dim3(128), 0, st, __VA_ARGS__); break;           \
    case 6: ztb_launch(KERNEL<6>, dim3(GRID), dim3(128), 0, st, __VA_ARGS__); break;           \
    case 8: ztb_launch(KERNEL<8>, dim3(GRID), dim3(128), 0, st, __VA_ARGS__); break;           \
  }

static bool ln_fast(int D) {
  int nv = D / 256;
  return D % 256 == 0 && (nv == 1 || nv == 2 || nv == 3 || nv == 4 || nv == 6 || nv == 8);
}

int ztb_layernorm_fwd(const void* x, const void* w, void* y, float* mean, float* rstd, int M, int D, float eps, cudaStream_t st) {
  if (ln_fast(D)) {
    int nv = D / 256, grid = (M + 3) / 4;
    LN_DISPATCH(nv, layernorm_fwd_kernel, grid, (const bf16*)x, (const bf16*)w, (bf16*)y, mean, rstd, M, eps)
  } else {
    ztb_prefer_max_smem((const void*)layernorm_fwd_generic);
    layernorm_fwd_generic<<<M, 128, 0, st>>>((const bf16*)x, (const bf16*)w, (bf16*)y, mean, rstd, M, D, eps);
  }
  ZTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int ztb_layernorm_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd, void* dx,
                      float* dw, int M, int D, int sms, cudaStream_t st, const void* dres, const ZtbLnBwdExtra* extra) {
  LnBwdExtra ex{};
  if (extra != nullptr) {
    ex.grad_out = extra->grad_out; ex.grad_fp32 = extra->grad_fp32; ex.grad_accumulate = extra->grad_accumulate;
    ex.ticket = extra->ticket; ex.dmask = (bf16*)extra->dmask; ex.p = extra->p;
    ex.partials = ln_fast(D) ? extra->partials : nullptr;   // the generic (odd-D) kernel keeps its per-element atomics
    ex.inv_keep = extra->p < 1.f ? 1.f / (1.f - extra->p) : 0.f; ex.rng = extra->rng; ex.op_id = extra->op_id;
    ex.maskbits = extra->maskbits;
  }
  if (ex.ticket == nullptr) ZTB_CUDA_CHECK(cudaMemsetAsync(dw, 0, sizeof(float) * D, st));   // one-shot scratch: caller owns dw
  if (ln_fast(D)) {
    int nv = D / 256, grid = grid_for((M + 3) / 4, 1, sms * 4);
    if (ex.dmask == nullptr) {
      LN_DISPATCH2(nv, layernorm_bwd_kernel, 0, grid, (const bf16*)dy, (const bf16*)x, (const bf16*)w, mean, rstd, (bf16*)dx, dw, M, (const bf16*)dres, ex)
    } else if (ex.maskbits != nullptr) {
      LN_DISPATCH2(nv, layernorm_bwd_kernel, 1, grid, (const bf16*)dy, (const bf16*)x, (const bf16*)w, mean, rstd, (bf16*)dx, dw, M, (const bf16*)dres, ex)
    } else {
      LN_DISPATCH2(nv, layernorm_bwd_kernel, 2, grid, (const bf16*)dy, (const bf16*)x, (const bf16*)w, mean, rstd, (bf16*)dx, dw, M, (const bf16*)dres, ex)
    }
  } else {
    ztb_prefer_max_smem((const void*)layernorm_bwd_generic);
    layernorm_bwd_generic<<<M, 128, 0, st>>>((const bf16*)dy, (const bf16*)x, (const bf16*)w, mean, rstd, (bf16*)dx, dw, M, D, (const bf16*)dres, ex);
  }
  ZTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int ztb_layernorm_bwd_grid(int M, int D, int sms) { return ln_fast(D) ? grid_for((M + 3) / 4, 1, sms * 4) : M; }

int ztb_dropout_add(const void* y, const void* res, void* out, size_t n, float p, const uint64_t* rng, uint32_t op_id,
                    int sms, cudaStream_t st) {
  if (n % 8) { ztb_set_error("dropout_add", "numel must be a multiple of 8", __FILE__, __LINE__); return -1; }
  size_t n8 = n / 8;
  int grid = grid_for(n8, 256, sms * 8);
  float inv_keep = p < 1.f ? 1.f / (1.f - p) : 0.f;
  if (res)
    ztb_launch(dropout_add_kernel<true>, dim3(grid), dim3(256), 0, st, (const bf16*)y, (const bf16*)res, (bf16*)out, n8, p, inv_keep, rng, op_id);
  else
    ztb_launch(dropout_add_kernel<false>, dim3(grid), dim3(256), 0, st, (const bf16*)y, (const bf16*)nullptr, (bf16*)out, n8, p, inv_keep, rng, op_id);
  ZTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int ztb_embedding_fwd(const int* idx, const void* wte, void* out, int M, int D, int V, cudaStream_t st) {
  ztb_launch(embedding_fwd_kernel, dim3((M + 3) / 4), dim3(128), 0, st, idx, (const bf16*)wte, (bf16*)out, M, D / 8, V);
  ZTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
int ztb_embedding_bwd(const int* idx, const void* dout, void* dwte, int grad_is_fp32, int M, int D, int V, cudaStream_t st) {
  ztb_prefer_max_smem((const void*)embedding_bwd_kernel<float>);
  ztb_prefer_max_smem((const void*)embedding_bwd_kernel<bf16>);
  if (grad_is_fp32)
    embedding_bwd_kernel<float><<<(M + 3) / 4, 128, 0, st>>>(idx, (const bf16*)dout, (float*)dwte, M, D, V);
  else
    embedding_bwd_kernel<bf16><<<(M + 3) / 4, 128, 0, st>>>(idx, (const bf16*)dout, (bf16*)dwte, M, D, V);
  ZTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int ztb_embedding_bwd_sorted(const int* sorted_tok, const int* order, const void* dout, void* dwte, int grad_is_fp32, int M, int D, int V,
                             cudaStream_t st) {
  ztb_prefer_max_smem((const void*)embedding_bwd_sorted_kernel<float>);
  ztb_prefer_max_smem((const void*)embedding_bwd_sorted_kernel<bf16>);
  if (grad_is_fp32)
    embedding_bwd_sorted_kernel<float><<<(M + 3) / 4, 128, 0, st>>>(sorted_tok, order, (const bf16*)dout, (float*)dwte, M, D, V);
  else
    embedding_bwd_sorted_kernel<bf16><<<(M + 3) / 4, 128, 0, st>>>(sorted_tok, order, (const bf16*)dout, (bf16*)dwte, M, D, V);
  ZTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int ztb_gelu_bwd(const void* dg, const void* z, void* dz, size_t n, int sms, cudaStream_t st) {
  ztb_prefer_max_smem((const void*)gelu_bwd_kernel);
  gelu_bwd_kernel<<<grid_for(n / 8, 256, sms * 8), 256, 0, st>>>((const bf16*)dg, (const bf16*)z, (bf16*)dz, n / 8);
  ZTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
int ztb_gelu_fwd(const void* z, void* g, size_t n, int sms, cudaStream_t st) {
  ztb_prefer_max_smem((const void*)gelu_fwd_kernel);
  gelu_fwd_kernel<<<grid_for(n / 8, 256, sms * 8), 256, 0, st>>>((const bf16*)z, (bf16*)g, n / 8);
  ZTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int ztb_ce_fwd(const void* logits, const int* labels, float* lse, float* nll, int M, int V, size_t ld, cudaStream_t st) {
  ztb_launch(ce_fwd_kernel, dim3(M), dim3(256), 0, st, (const bf16*)logits, labels, lse, nll, V, ld);
  ZTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
int ztb_ce_bwd(void* logits, const int* labels, const float* lse, const float* gscale, float inv_count, int M, int V,
               size_t ld, cudaStream_t st) {
  ztb_launch(ce_bwd_kernel, dim3(M), dim3(256), 0, st, (bf16*)logits, labels, lse, gscale, inv_count, V, ld);
  ZTB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
