// Common device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM PTX
// wrappers, Philox RNG, vector load/store, small math.  Everything here is inline
// PTX / plain CUDA -- no CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define ZTB_DEVINL __device__ __forceinline__

#define ZTB_CUDA_CHECK(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ztb_set_error(#expr, cudaGetErrorString(_e), __FILE__, __LINE__);                        \
      return (int)_e;                                                                          \
    }                                                                                          \
  } while (0)

// Host-side error slot read by the Python bindings (defined in host_util.cu).
void ztb_set_error(const char* expr, const char* msg, const char* file, int line);
const char* ztb_last_error();

namespace ztb {

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

// ------------------------------------------------------------------ misc
ZTB_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
ZTB_DEVINL uint32_t lane_id() { return threadIdx.x & 31; }

ZTB_DEVINL bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

ZTB_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
ZTB_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

ZTB_DEVINL uint32_t pack_bf16(float lo, float hi) {
  bf162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
ZTB_DEVINL float2 unpack_bf16(uint32_t u) {
  bf162 v = *reinterpret_cast<bf162*>(&u);
  return __bfloat1622float2(v);
}

// tanh-approximate GELU (flax nn.gelu default; reference layers.py:68) and its derivative.
ZTB_DEVINL float gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float u = k0 * x * (1.f + k1 * x * x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  return 0.5f * x * (1.f + t);
}
ZTB_DEVINL float gelu_tanh_grad(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float x2 = x * x;
  float u = k0 * x * (1.f + k1 * x2);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  float du = k0 * (1.f + 3.f * k1 * x2);
  return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * du;
}

// gelu'(z) for a bf16 PAIR, evaluated in packed fp16 (HFMA2 / tanh.approx.f16x2): 11 packed instructions for two values
// instead of ~15 fp32 instructions each.  Used where the evaluation sits in a GEMM epilogue whose instruction count limits
// the kernel (dgrad + GELU' at K = 2048: epilogue issue time ~ mainloop time per tile).  z is clamped to +-16 first
// (gelu' is 0 / 1 to 30 digits there; keeps x^2 and the cubic far from the fp16 range).  Max absolute error against the
// fp64 formula over bf16 inputs in [-12, 12]: 2.1e-3, i.e. about half of the bf16 rounding (3.9e-3 at 1.0) that the product
// receives anyway.  ZTB_GELU_GRAD_F16=0 at compile time restores the fp32 evaluation.
#ifndef ZTB_GELU_GRAD_F16
#define ZTB_GELU_GRAD_F16 1
#endif
ZTB_DEVINL float2 gelu_tanh_grad_pair(uint32_t z_bf16x2) {
#if ZTB_GELU_GRAD_F16
  const float zlo = __uint_as_float(z_bf16x2 << 16), zhi = __uint_as_float(z_bf16x2 & 0xffff0000u);
  uint32_t h, g;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(zhi), "f"(zlo));
  asm("{\n\t"
      ".reg .b32 x, x2, pp, u, t, nt, du, a, b, c, one, half, k0, k1, k3, lo, hi;\n\t"
      "mov.b32 one, 0x3C003C00;\n\t"
      "mov.b32 half, 0x38003800;\n\t"
      "mov.b32 k0, 0x3A623A62;\n\t"      // 0.79788
      "mov.b32 k1, 0x29B929B9;\n\t"      // 0.044715
      "mov.b32 k3, 0x2EDA2EDA;\n\t"      // 3 k0 k1
      "mov.b32 lo, 0xCC00CC00;\n\t"      // -16
      "mov.b32 hi, 0x4C004C00;\n\t"      // +16
      "max.f16x2 x, %1, lo;\n\t"         // clamp to [-16, 16]
      "min.f16x2 x, x, hi;\n\t"
      "mul.f16x2 x2, x, x;\n\t"
      "fma.rn.f16x2 pp, x2, k1, one;\n\t"     // 1 + k1 x^2
      "mul.f16x2 u, x, pp;\n\t"
      "mul.f16x2 u, u, k0;\n\t"               // k0 x (1 + k1 x^2)
      "tanh.approx.f16x2 t, u;\n\t"
      "fma.rn.f16x2 du, x2, k3, k0;\n\t"      // k0 (1 + 3 k1 x^2)
      "fma.rn.f16x2 a, t, half, half;\n\t"    // 0.5 (1 + t)
      "neg.f16x2 nt, t;\n\t"
      "fma.rn.f16x2 b, nt, t, one;\n\t"       // 1 - t^2
      "mul.f16x2 c, x, du;\n\t"
      "mul.f16x2 c, c, half;\n\t"             // 0.5 x du
      "fma.rn.f16x2 %0, c, b, a;\n\t"
      "}" : "=r"(g) : "r"(h));
  float glo, ghi;
  asm("{\n\t.reg .b16 l, hh;\n\tmov.b32 {l, hh}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, hh;\n\t}" : "=f"(glo), "=f"(ghi) : "r"(g));
  return make_float2(glo, ghi);
#else
  return make_float2(gelu_tanh_grad(__uint_as_float(z_bf16x2 << 16)), gelu_tanh_grad(__uint_as_float(z_bf16x2 & 0xffff0000u)));
#endif
}

// ------------------------------------------------------------------ Philox4x32
// Dropout masks use the 7-round variant: Philox4x32-7 is the smallest round count Random123 documents as
// Crush-resistant (10 is its conservative default); one call per 16 attention probabilities sits in the softmax
// critical path, so the rounds are not free.  Every mask site (GEMM epilogue, dropout_add, LayerNorm-backward mask,
// attention forward / backward) uses the same constant, so a mask regenerated in backward is the forward's mask.
constexpr int kPhiloxRounds = 7;
// Counter-based RNG: dropout masks are regenerated in backward from (seed, offset)
// instead of being stored (per (step, rank, micro-step, layer, op) streams).
struct Philox {
  uint32_t k0, k1;
  ZTB_DEVINL Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  ZTB_DEVINL uint4 operator()(uint64_t ctr, uint64_t stream) const { return rounds<10>(ctr, stream); }
  template <int R>
  ZTB_DEVINL uint4 rounds(uint64_t ctr, uint64_t stream) const {
    uint4 c = make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)stream, (uint32_t)(stream >> 32));
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
      c = make_uint4(hi1 ^ c.y ^ a, lo1, hi0 ^ c.w ^ b, lo0);
      a += 0x9E3779B9u;
      b += 0xBB67AE85u;
    }
    return c;
  }
};
// keep-threshold compare on raw 32-bit words: keep iff r >= p * 2^32
ZTB_DEVINL uint32_t dropout_threshold(float p) { return (uint32_t)fminf(p * 4294967296.f, 4294967295.f); }

// ------------------------------------------------------------------ mbarrier
ZTB_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
ZTB_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
ZTB_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
ZTB_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
ZTB_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
ZTB_DEVINL void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
ZTB_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
ZTB_DEVINL void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

ZTB_DEVINL void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------ TMA
ZTB_DEVINL void prefetch_tmap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// 2D tile load global -> shared, completion on an mbarrier (c0 = innermost coordinate)
ZTB_DEVINL void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
ZTB_DEVINL void tma_load_2d_hint(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
ZTB_DEVINL void tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// reduce-add store (fp32/bf16 element type taken from the tensor map)
ZTB_DEVINL void tma_reduce_add_2d(const void* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
ZTB_DEVINL void tma_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// 16-byte store to shared memory through the shared window (pointers derived from the aligned dynamic smem base
// lose their address space, and the compiler would otherwise emit generic ST.E.128)
ZTB_DEVINL void sts_v4(void* p, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int N>
ZTB_DEVINL void tma_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
ZTB_DEVINL void tma_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// L2 eviction-priority policies for TMA cache hints
ZTB_DEVINL uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
ZTB_DEVINL uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// ------------------------------------------------------------------ tcgen05 / TMEM
template <uint32_t kCols>
ZTB_DEVINL void tmem_alloc(uint32_t* smem_slot) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
ZTB_DEVINL void tmem_dealloc(uint32_t taddr) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
ZTB_DEVINL void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
ZTB_DEVINL void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; single thread issues.
ZTB_DEVINL void umma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from TMEM (used by attention: P stays in tensor memory)
ZTB_DEVINL void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when complete
ZTB_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Shared-memory matrix descriptor (see cute/arch/mma_sm100_desc.hpp SmemDescriptor for the bit
// layout): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout [61,64)
constexpr uint64_t kSwizzle128B = 2, kSwizzle64B = 4, kSwizzle32B = 6, kSwizzleNone = 0;
ZTB_DEVINL uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}
// Instruction descriptor for kind::f16 (InstrDescriptor in mma_sm100_desc.hpp):
// c_format F32 (1<<4) | a_format (bits 7-9) | b_format (bits 10-12) | a_major bit15 | b_major bit16 |
// N>>3 [17,23) | M>>4 [24,29).  format: 0 = f16, 1 = bf16.  major: 0 = K-major, 1 = MN-major.
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn_major, uint32_t b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// TMEM -> registers: 32 lanes x 32-bit, N consecutive columns; thread i of the warp gets lane
// (taddr.lane + i).  A warp may only touch lanes [32*(warp_id%4), +32).
ZTB_DEVINL void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
ZTB_DEVINL void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
ZTB_DEVINL void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
ZTB_DEVINL void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
ZTB_DEVINL void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ global memory
ZTB_DEVINL uint4 ld_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
ZTB_DEVINL uint32_t ld_nc_u8(const void* p) {
  uint32_t r;
  asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
// read-only path WITH L1 allocation: for strided per-thread row reads where the other 16-byte pieces of
// the same 128-byte line are consumed by the following loads of the same thread
ZTB_DEVINL uint4 ld_nc_l1_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
ZTB_DEVINL uint4 ld_v4(const void* p) { return *reinterpret_cast<const uint4*>(p); }
ZTB_DEVINL void st_v4(void* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }
ZTB_DEVINL void st_na_v4(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

}  // namespace ztb
