// ALiBi-biased causal flash attention on tcgen05 / TMEM (forward).
//
// Reference semantics (src/models/layers.py:145-178): S = Q K^T * scale + slope_h * (j - i), causal
// mask, fp32 softmax, dropout on the probabilities, O = P V.  The reference materialises the
// (b, h, T, T) fp32 score tensor; here nothing of size T^2 ever leaves the SM.
//
// One CTA per (batch, head, 128-query block); KV blocks of 128 keys are streamed through a
// 2-stage TMA ring.  Warp roles: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (one thread),
// warps 2-5 = softmax (one query row per thread = one TMEM lane, so row max / row sum need no
// shuffles).  Per KV block:
//     S_j   = Q K_j^T            tcgen05.mma  SS  (both operands K-major)      -> TMEM (2 buffers)
//     P_j   = exp2(S_j*c + alibi - m)  (registers; lazy rescale: m only moves when the row max
//             grows by more than 2^8, FA4-style)                               -> smem bf16 (SW128)
//     O    += P_j V_j            tcgen05.mma  SS  (V consumed MN-major)         -> TMEM
// S_{j+1} is issued before P_j is consumed so the tensor core overlaps the softmax.  Dropout keeps
// an element iff an 8-bit Philox lane >= round(256 p) (16 decisions per Philox call) and the
// 1/(1-p') factor is folded into the final 1/l normalisation.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "launch.cuh"
#include "launchers.h"

namespace ztb {

constexpr int ATT_BM = 128, ATT_BN = 128;
constexpr int ATT_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 softmax (2 column halves x 4 lane quarters)

ZTB_DEVINL float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int HD>
struct AttnSmem {
  static constexpr int CH = (HD + 63) / 64;        // 64-column (128-byte) chunks per row (head_dim 96 is staged as 128)
  static constexpr int TILE = 128 * 128 * CH;      // bytes of a [128 x HD] bf16 tile
  static constexpr int Q_OFF = 0;
  static constexpr int K_OFF = TILE;               // 2 stages
  static constexpr int V_OFF = 3 * TILE;           // 2 stages
  static constexpr int P_OFF = 5 * TILE;           // [128 x 128] bf16 = 32 KB (also O staging)
  static constexpr int BAR_OFF = 5 * TILE + 32768;
  static constexpr int TOTAL = BAR_OFF + 128 + 3072 + 1024;  // barriers + row-max/row-sum exchange
};

struct AttnParams {
  int B, T, H;
  float scale_log2;          // softmax scale * log2(e)
  const float* slopes;       // [H] or nullptr
  float* lse;                // [B, H, T] natural-log LSE (for backward)
  bf16* out;                 // [B*T, H*hd] (direct stores for a partial 64-column chunk, head_dim 96)
  uint32_t drop_thr8;        // keep iff byte >= thr (0 = no dropout)
  float inv_keep;            // 1 / (1 - thr/256)
  const uint64_t* rng;
  uint32_t op_id;
};

ZTB_DEVINL void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// Four keep decisions at once: bit 7 of byte k of the result is set iff byte k of `w` >= byte k of `t4` (unsigned);
// 4 integer instructions for 4 probabilities instead of extract + compare per element.
ZTB_DEVINL uint32_t bytes_ge(uint32_t w, uint32_t t4) {
  const uint32_t H = 0x80808080u;
  const uint32_t d = (w | H) - (t4 & ~H);            // per byte (no borrow across bytes): bit 7 = low7(w) >= low7(t)
  return ((w & ~t4) | (~(w ^ t4) & d)) & H;
}
// expand the bit-7 flags of bytes (lo, lo + 1) into a mask over a packed bf16x2 word (PRMT sign-replicate mode)
// (inline PTX: the __byte_perm intrinsic only honours 3 selector bits per nibble; bit 3 = "replicate the selected byte's
// sign bit" is a feature of the prmt instruction's default mode)
ZTB_DEVINL uint32_t prmt_b32(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
ZTB_DEVINL uint32_t pair_mask01(uint32_t flags) { return prmt_b32(flags, 0u, 0x9988u); }
ZTB_DEVINL uint32_t pair_mask23(uint32_t flags) { return prmt_b32(flags, 0u, 0xbbaau); }

// 16 keep/drop decisions for elements [16*g, 16*g+16) of row `row_id` from one Philox call
ZTB_DEVINL uint4 attn_dropout_bits(const Philox& ph, uint64_t row_id, uint32_t g, uint64_t stream) {
  return ph.rounds<kPhiloxRounds>(row_id * 4096ull + g, stream);
}

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO, const AttnParams p) {
  using L = AttnSmem<HD>;
  constexpr int CH = L::CH;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smQ = smem + L::Q_OFF;
  uint8_t* smK = smem + L::K_OFF;
  uint8_t* smV = smem + L::V_OFF;
  uint8_t* smP = smem + L::P_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* q_full = bars;          // 1
  uint64_t* k_full = bars + 1;      // 2
  uint64_t* v_full = bars + 3;      // 2
  uint64_t* kv_empty = bars + 5;    // 2
  uint64_t* s_full = bars + 7;      // 2
  uint64_t* s_empty = bars + 9;     // 2
  uint64_t* p_ready = bars + 11;    // 1
  uint64_t* pv_done = bars + 12;    // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = gridDim.x - 1 - blockIdx.x;  // heavy (long) rows first
  const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H;
  const int D = p.H * HD;
  const int nblk = qb + 1;                    // causal: KV blocks 0..qb
  const int row0 = b * p.T + qb * ATT_BM;     // first global row of this query block
  const int colQ = h * HD, colK = D + h * HD, colV = 2 * D + h * HD;

  if (warp == 0 && lane == 0) { prefetch_tmap(&tmQKV); prefetch_tmap(&tmO); }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&kv_empty[i], 1);
                                  mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 256); }
    mbar_init(p_ready, 256);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_enter();   // prologue above overlapped the previous kernel's tail; global memory is first touched below
  const uint32_t tS = tmem_base;          // 2 x 128 columns
  const uint32_t tO = tmem_base + 256;    // HD columns

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, L::TILE);
#pragma unroll
      for (int c = 0; c < CH; ++c) tma_load_2d(smQ + c * 16384, &tmQKV, q_full, colQ + c * 64, row0);
      for (int j = 0; j < nblk; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        const int krow = b * p.T + j * ATT_BN;
        mbar_expect_tx(&k_full[st], L::TILE);
#pragma unroll
        for (int c = 0; c < CH; ++c) tma_load_2d(smK + st * L::TILE + c * 16384, &tmQKV, &k_full[st], colK + c * 64, krow);
        mbar_expect_tx(&v_full[st], L::TILE);
#pragma unroll
        for (int c = 0; c < CH; ++c) tma_load_2d(smV + st * L::TILE + c * 16384, &tmQKV, &v_full[st], colV + c * 64, krow);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(ATT_BM, HD, 0, 1);
      const uint32_t q_addr = smem_u32(smQ), p_addr = smem_u32(smP);
      auto issue_s = [&](int j) {
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1);
        mbar_wait(&s_empty[st], ((j >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t k_addr = smem_u32(smK + st * L::TILE);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_f16_ss(tS + st * 128, make_smem_desc(q_addr + off, 16, 1024, kSwizzle128B),
                      make_smem_desc(k_addr + off, 16, 1024, kSwizzle128B), idesc_s, k != 0);
        }
        umma_commit(&s_full[st]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < nblk; ++j) {
        const int st = j & 1;
        if (j + 1 < nblk) issue_s(j + 1);
        mbar_wait(p_ready, j & 1);
        mbar_wait(&v_full[st], (j >> 1) & 1);
        tc_fence_after_sync();
        const uint32_t v_addr = smem_u32(smV + st * L::TILE);
#pragma unroll
        for (int k = 0; k < ATT_BN / 16; ++k) {
          uint64_t da = make_smem_desc(p_addr + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024, kSwizzle128B);
          uint64_t db = make_smem_desc(v_addr + k * 2048, 16384, 1024, kSwizzle128B);
          umma_f16_ss(tO, da, db, idesc_o, (j | k) != 0);
        }
        umma_commit(&kv_empty[st]);
        umma_commit(pv_done);
      }
    }
  } else {
    // ===================================================== softmax warps
    // thread = (query row r = TMEM lane, column half hf): two warps share each lane quarter and
    // split the 128 key columns, so every SMSP runs two softmax warps.
    const int sw = warp - 2;
    const int quarter = warp & 3;
    const int hf = sw >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int qi = qb * ATT_BM + r;
    constexpr int SPLIT = (HD == 64) ? 32 : 64;   // output columns: half 0 owns [0, SPLIT), half 1 [SPLIT, HD)
    constexpr int FULLCH = HD / 64;               // 64-column chunks that leave through {64 x 128} TMA boxes
    const int hc0 = hf ? SPLIT : 0, hcn = hf ? HD - SPLIT : SPLIT;
    const float slope2 = p.slopes ? p.slopes[h] * 1.4426950408889634f : 0.f;
    const float kk = slope2 / p.scale_log2;     // ALiBi slope in raw-score units
    const bool use_drop = p.drop_thr8 != 0;
    uint64_t seed = 0, stream = 0;
    if (use_drop) { seed = p.rng[0]; stream = p.rng[1] * 65536ull + p.op_id; }
    Philox ph(seed);
    const uint64_t row_id = (uint64_t)bh * p.T + qi;
    float* xch = reinterpret_cast<float*>(smem + L::BAR_OFF + 128);  // [2 parity][2 half][128] row-max exchange
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < nblk; ++j) {
      const int st = j & 1;
      mbar_wait(&s_full[st], (j >> 1) & 1);
      tc_fence_after_sync();
      float s[64];
      {
        uint32_t* sr = reinterpret_cast<uint32_t*>(s);
        tmem_ld_32x32b_x32(tS + lane_addr + st * 128 + hf * 64, sr);
        tmem_ld_32x32b_x32(tS + lane_addr + st * 128 + hf * 64 + 32, sr + 32);
        tmem_wait_ld();
      }
      tc_fence_before_sync();
      mbar_arrive(&s_empty[st]);
      // Scores in log2 units: v_c = scale2 * u_c + base with u_c = S_c + kk * c (kk = slope / scale), so the
      // first pass is one FMA (immediate column index) + max per element and the second pass one FMA +
      // ex2 + add; the causal mask is only evaluated on the diagonal block.  4-way partial max / sum
      // break the serial dependency chains.
      const int kv0 = j * ATT_BN + hf * 64;
      const float base = slope2 * (float)(kv0 - qi);
      const int cmax = r - hf * 64;          // columns > cmax are in the future (diagonal block only)
      float um[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (j == qb) {
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          float u = fmaf(kk, (float)c, s[c]);
          u = c > cmax ? -INFINITY : u;
          s[c] = u;
          um[c & 3] = fmaxf(um[c & 3], u);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          float u = fmaf(kk, (float)c, s[c]);
          s[c] = u;
          um[c & 3] = fmaxf(um[c & 3], u);
        }
      }
      float mx = fmaf(fmaxf(fmaxf(um[0], um[1]), fmaxf(um[2], um[3])), p.scale_log2, base);
      xch[(j & 1) * 256 + hf * 128 + r] = mx;
      named_bar_sync(3, 256);
      mx = fmaxf(mx, xch[(j & 1) * 256 + (hf ^ 1) * 128 + r]);
      float alpha = 1.f;
      if (mx > m + 8.f) { alpha = ex2_ftz(m - mx); m = mx; }  // lazy rescale; first block: alpha = 0
      const float off = base - m;
      float sm[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 64; ++c) { s[c] = ex2_ftz(fmaf(s[c], p.scale_log2, off)); sm[c & 3] += s[c]; }
      l = l * alpha + ((sm[0] + sm[1]) + (sm[2] + sm[3]));
      uint32_t pk[32];
      if (use_drop) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 rb = attn_dropout_bits(ph, row_id, (uint32_t)(j * 8 + hf * 4 + g), stream);
          uint32_t w[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            uint32_t b0 = (w[e >> 2] >> ((e & 3) * 8)) & 0xFF, b1 = (w[(e + 1) >> 2] >> (((e + 1) & 3) * 8)) & 0xFF;
            float a0 = b0 >= p.drop_thr8 ? s[g * 16 + e] : 0.f, a1 = b1 >= p.drop_thr8 ? s[g * 16 + e + 1] : 0.f;
            pk[g * 8 + (e >> 1)] = pack_bf16(a0, a1);
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) pk[c] = pack_bf16(s[2 * c], s[2 * c + 1]);
      }
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);
        tc_fence_after_sync();
        if (__any_sync(0xffffffffu, alpha != 1.f)) {  // this thread rescales its half of the O columns
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (c * 32 >= hcn) break;
            uint32_t o[32];
            const uint32_t ta = tO + lane_addr + hc0 + c * 32;
            tmem_ld_32x32b_x32(ta, o);
            tmem_wait_ld();
#pragma unroll
            for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
            tmem_st_32x32b_x32(ta, o);
          }
          tmem_wait_st();
        }
      }
      uint8_t* prow = smP + hf * 16384 + r * 128;
#pragma unroll
      for (int c16 = 0; c16 < 8; ++c16)
        sts_v4(prow + ((c16 ^ (r & 7)) << 4), make_uint4(pk[c16 * 4], pk[c16 * 4 + 1], pk[c16 * 4 + 2], pk[c16 * 4 + 3]));
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(p_ready);
    }
    // ---- epilogue: combine the two partial row sums, O / l -> bf16 -> staging (P region) -> TMA store
    xch[512 + hf * 128 + r] = l;
    mbar_wait(pv_done, (nblk - 1) & 1);
    tc_fence_after_sync();
    named_bar_sync(3, 256);
    l += xch[512 + (hf ^ 1) * 128 + r];
    const float inv = (use_drop ? p.inv_keep : 1.f) / l;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (c * 32 >= hcn) break;
      uint32_t o[32];
      const int col = hc0 + c * 32;          // first O column of this 32-column piece
      tmem_ld_32x32b_x32(tO + lane_addr + col, o);
      tmem_wait_ld();
      uint4 o4[4];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4)
        o4[q4] = make_uint4(pack_bf16(__uint_as_float(o[q4 * 8 + 0]) * inv, __uint_as_float(o[q4 * 8 + 1]) * inv),
                            pack_bf16(__uint_as_float(o[q4 * 8 + 2]) * inv, __uint_as_float(o[q4 * 8 + 3]) * inv),
                            pack_bf16(__uint_as_float(o[q4 * 8 + 4]) * inv, __uint_as_float(o[q4 * 8 + 5]) * inv),
                            pack_bf16(__uint_as_float(o[q4 * 8 + 6]) * inv, __uint_as_float(o[q4 * 8 + 7]) * inv));
      if ((col >> 6) < FULLCH) {
        uint8_t* orow = smP + (col >> 6) * 16384 + r * 128;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) sts_v4(orow + (((((col & 63) >> 3) + q4) ^ (r & 7)) << 4), o4[q4]);
      } else {  // partial chunk (head_dim 96): straight to global, 64 contiguous bytes per row
        bf16* gp = p.out + (size_t)(row0 + r) * D + colQ + col;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) st_v4(gp + q4 * 8, o4[q4]);
      }
    }
    if (hf == 0) p.lse[(size_t)bh * p.T + qi] = (m + log2f(l)) * 0.6931471805599453f;
    fence_proxy_async_smem();
    named_bar_sync(1, 256);
    if (threadIdx.x == 64) {
#pragma unroll
      for (int c = 0; c < FULLCH; ++c) tma_store_2d(&tmO, smP + c * 16384, colQ + c * 64, row0);
      tma_commit_group();
      tma_wait_group<0>();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}


// =====================================================================================
// Forward, two-CTAs-per-SM variant (head_dim 128).  Profiling the 1-CTA/SM kernel showed ~40 % of
// its time in exposed per-CTA prologue/epilogue latency (TMEM alloc, Q load, O store) and role
// hand-offs: a second co-resident CTA fills those bubbles.  To fit two CTAs (<= 113 KB smem, 256 TMEM
// columns, <= 170 registers) the KV block shrinks to 64 keys and each of the 4 softmax warps owns
// full 64-column rows again (no cross-warp row-max exchange).
constexpr int ATT3_THREADS = 192;
constexpr int ATT3_BN = 64;
struct Attn3Smem {
  static constexpr int Q_OFF = 0;                  // [128 x 128] bf16 = 32 KB (also O staging)
  static constexpr int K_OFF = 32768;              // 2 stages x [64 x 128] = 2 x 16 KB
  static constexpr int V_OFF = 65536;              // 2 stages x 16 KB
  static constexpr int P_OFF = 98304;              // [128 x 64] bf16 = 16 KB
  static constexpr int BAR_OFF = 114688;
  static constexpr int TOTAL = 115712;             // 113 KB (multiple of 1 KB): two CTAs + 2 x 1 KB reserved = 228 KB
};

__global__ void __launch_bounds__(ATT3_THREADS, 2)
attn_fwd3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 const __grid_constant__ CUtensorMap tmO, const AttnParams p) {
  constexpr int HD = 128;
  using L = Attn3Smem;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // SWIZZLE_128B tiles need a 1 KB-aligned base
  uint8_t* smQ = smem + L::Q_OFF;
  uint8_t* smK = smem + L::K_OFF;
  uint8_t* smV = smem + L::V_OFF;
  uint8_t* smP = smem + L::P_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;      // 2
  uint64_t* v_full = bars + 3;      // 2
  uint64_t* kv_empty = bars + 5;    // 2
  uint64_t* s_full = bars + 7;      // 2
  uint64_t* s_empty = bars + 9;     // 2 (count 128)
  uint64_t* p_ready = bars + 11;    // count 128
  uint64_t* pv_done = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = gridDim.x - 1 - blockIdx.x;
  const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H;
  const int D = p.H * HD;
  const int nblk = 2 * qb + 2;                       // 64-key blocks 0 .. 2qb+1
  const int row0 = b * p.T + qb * ATT_BM;
  const int colQ = h * HD, colK = D + h * HD, colV = 2 * D + h * HD;

  if (warp == 0 && lane == 0) { prefetch_tmap(&tmQ); prefetch_tmap(&tmKV); prefetch_tmap(&tmO); }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&kv_empty[i], 1);
                                  mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 128); }
    mbar_init(p_ready, 128);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_enter();   // prologue above overlapped the previous kernel's tail; global memory is first touched below
  const uint32_t tS = tmem_base;          // 2 x 64 columns
  const uint32_t tO = tmem_base + 128;    // 128 columns

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 32768);
#pragma unroll
      for (int c = 0; c < 2; ++c) tma_load_2d(smQ + c * 16384, &tmQ, q_full, colQ + c * 64, row0);
      for (int j = 0; j < nblk; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        const int krow = b * p.T + j * ATT3_BN;
        mbar_expect_tx(&k_full[st], 16384);
#pragma unroll
        for (int c = 0; c < 2; ++c) tma_load_2d(smK + st * 16384 + c * 8192, &tmKV, &k_full[st], colK + c * 64, krow);
        mbar_expect_tx(&v_full[st], 16384);
#pragma unroll
        for (int c = 0; c < 2; ++c) tma_load_2d(smV + st * 16384 + c * 8192, &tmKV, &v_full[st], colV + c * 64, krow);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT3_BN, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(ATT_BM, HD, 0, 1);
      const uint32_t q_addr = smem_u32(smQ), p_addr = smem_u32(smP);
      auto issue_s = [&](int j) {
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1);
        mbar_wait(&s_empty[st], ((j >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t k_addr = smem_u32(smK + st * 16384);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_f16_ss(tS + st * 64, make_smem_desc(q_addr + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024, kSwizzle128B),
                      make_smem_desc(k_addr + (k >> 2) * 8192 + (k & 3) * 32, 16, 1024, kSwizzle128B), idesc_s, k != 0);
        umma_commit(&s_full[st]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < nblk; ++j) {
        const int st = j & 1;
        if (j + 1 < nblk) issue_s(j + 1);
        mbar_wait(p_ready, j & 1);
        mbar_wait(&v_full[st], (j >> 1) & 1);
        tc_fence_after_sync();
        const uint32_t v_addr = smem_u32(smV + st * 16384);
#pragma unroll
        for (int k = 0; k < ATT3_BN / 16; ++k)
          umma_f16_ss(tO, make_smem_desc(p_addr + k * 32, 16, 1024, kSwizzle128B),
                      make_smem_desc(v_addr + k * 2048, 8192, 1024, kSwizzle128B), idesc_o, (j | k) != 0);
        umma_commit(&kv_empty[st]);
        umma_commit(pv_done);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int qi = qb * ATT_BM + r;
    const float slope2 = p.slopes ? p.slopes[h] * 1.4426950408889634f : 0.f;
    const float kk = slope2 / p.scale_log2;
    const bool use_drop = p.drop_thr8 != 0;
    uint64_t seed = 0, stream = 0;
    if (use_drop) { seed = p.rng[0]; stream = p.rng[1] * 65536ull + p.op_id; }
    Philox ph(seed);
    const uint64_t row_id = (uint64_t)bh * p.T + qi;
    const uint32_t thr4 = p.drop_thr8 * 0x01010101u;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < nblk; ++j) {
      const int st = j & 1;
      mbar_wait(&s_full[st], (j >> 1) & 1);
      tc_fence_after_sync();
      float s[64];
      {
        uint32_t* sr = reinterpret_cast<uint32_t*>(s);
        tmem_ld_32x32b_x32(tS + lane_addr + st * 64, sr);
        tmem_ld_32x32b_x32(tS + lane_addr + st * 64 + 32, sr + 32);
        tmem_wait_ld();
      }
      tc_fence_before_sync();
      mbar_arrive(&s_empty[st]);
      const int kv0 = j * ATT3_BN;
      const float base = slope2 * (float)(kv0 - qi);
      const int cmax = qi - kv0;                     // columns > cmax are in the future
      float um[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (j >= 2 * qb) {
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          float u = fmaf(kk, (float)c, s[c]);
          u = c > cmax ? -INFINITY : u;
          s[c] = u;
          um[c & 3] = fmaxf(um[c & 3], u);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          float u = fmaf(kk, (float)c, s[c]);
          s[c] = u;
          um[c & 3] = fmaxf(um[c & 3], u);
        }
      }
      const float mx = fmaf(fmaxf(fmaxf(um[0], um[1]), fmaxf(um[2], um[3])), p.scale_log2, base);
      float alpha = 1.f;
      if (mx > m + 8.f) { alpha = ex2_ftz(m - mx); m = mx; }   // m stays finite: block 0 always has column 0 visible
      const float off = base - m;
      float sm[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 64; ++c) { s[c] = ex2_ftz(fmaf(s[c], p.scale_log2, off)); sm[c & 3] += s[c]; }
      l = l * alpha + ((sm[0] + sm[1]) + (sm[2] + sm[3]));
      uint32_t pk[32];
      if (use_drop) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 rb = attn_dropout_bits(ph, row_id, (uint32_t)(j * 4 + g), stream);
          uint32_t w[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {     // word k decides elements 16 g + 4 k .. + 3 (byte b <-> element 4 k + b, as in backward)
            const uint32_t keep = bytes_ge(w[k], thr4);
            pk[g * 8 + 2 * k] = pack_bf16(s[g * 16 + 4 * k], s[g * 16 + 4 * k + 1]) & pair_mask01(keep);
            pk[g * 8 + 2 * k + 1] = pack_bf16(s[g * 16 + 4 * k + 2], s[g * 16 + 4 * k + 3]) & pair_mask23(keep);
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) pk[c] = pack_bf16(s[2 * c], s[2 * c + 1]);
      }
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);
        tc_fence_after_sync();
        if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tO + lane_addr + c * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
            tmem_st_32x32b_x32(tO + lane_addr + c * 32, o);
          }
          tmem_wait_st();
        }
      }
      uint8_t* prow = smP + r * 128;
#pragma unroll
      for (int c16 = 0; c16 < 8; ++c16)
        sts_v4(prow + ((c16 ^ (r & 7)) << 4), make_uint4(pk[c16 * 4], pk[c16 * 4 + 1], pk[c16 * 4 + 2], pk[c16 * 4 + 3]));
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(p_ready);
    }
    // ---- epilogue: O / l -> bf16 -> staging (Q region: every S MMA has retired) -> TMA store; LSE
    mbar_wait(pv_done, (nblk - 1) & 1);
    tc_fence_after_sync();
    const float inv = (use_drop ? p.inv_keep : 1.f) / l;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(tO + lane_addr + c * 32, o);
      tmem_wait_ld();
      uint8_t* orow = smQ + (c >> 1) * 16384 + r * 128;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint4 o4 = make_uint4(pack_bf16(__uint_as_float(o[q4 * 8 + 0]) * inv, __uint_as_float(o[q4 * 8 + 1]) * inv),
                              pack_bf16(__uint_as_float(o[q4 * 8 + 2]) * inv, __uint_as_float(o[q4 * 8 + 3]) * inv),
                              pack_bf16(__uint_as_float(o[q4 * 8 + 4]) * inv, __uint_as_float(o[q4 * 8 + 5]) * inv),
                              pack_bf16(__uint_as_float(o[q4 * 8 + 6]) * inv, __uint_as_float(o[q4 * 8 + 7]) * inv));
        sts_v4(orow + ((((c & 1) * 4 + q4) ^ (r & 7)) << 4), o4);
      }
    }
    p.lse[(size_t)bh * p.T + qi] = (m + log2f(l)) * 0.6931471805599453f;
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (threadIdx.x == 64) {
#pragma unroll
      for (int c = 0; c < 2; ++c) tma_store_2d(&tmO, smQ + c * 16384, colQ + c * 64, row0);
      tma_commit_group();
      tma_wait_group<0>();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<256>(tmem_base);
  }
}

// =====================================================================================
// Backward.  One CTA per (batch, head, 128-key block j); K_j and V_j stay resident in smem and the
// CTA walks the query blocks i >= j.  Per (i, j) pair, five tcgen05 GEMMs:
//     S  = Q_i K_j^T                    -> TMEM[0:128)         (softmax recomputed from the saved LSE)
//     dP = dO_i V_j^T                   -> TMEM[128:256)
//     dV += P~^T dO_i   (P~ = dropout(P) from smem, consumed MN-major)   -> TMEM[256:256+HD)
//     dK += dS^T Q_i    (dS = P o (D o dP / keep - delta) * scale)        -> TMEM[384:384+HD)
//     dQ_i += dS K_j    -> TMEM[0:HD) (reuses the S columns) -> fp32 smem -> TMA reduce-add
// Softmax threads own one query row each (same mapping and Philox stream as forward).
template <int HD>
struct AttnBwdSmem {
  static constexpr int CH = (HD + 63) / 64;
  static constexpr int TILE = 128 * 128 * CH;
  static constexpr int K_OFF = 0, V_OFF = TILE, Q_OFF = 2 * TILE /* 2 stages */, DO_OFF = 4 * TILE;
  static constexpr int P_OFF = 5 * TILE;            // 32 KB: P~ (bf16) / dQ staging (fp32) / dV staging
  static constexpr int DS_OFF = 5 * TILE + 32768;   // 32 KB: dS (bf16) / dK staging
  static constexpr int BAR_OFF = 5 * TILE + 65536;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};

struct AttnBwdParams {
  int B, T, H;
  float scale_log2, scale;
  const float* slopes;
  const float* lse;      // [B,H,T]
  const float* delta;    // [B,H,T] rowsum(dO o O)
  bf16* dqkv;            // [B*T, 3*H*hd] (direct stores for a partial 64-column chunk)
  unsigned long long* dbg;  // optional: per-phase cycle counters of softmax thread 0 (bench/attn_timing.py)
  int dq_bf16;           // head_dim 128 only: dQ partial sums meet in the bf16 dqkv buffer (bf16 TMA reduce-add, Q columns zeroed by
                         // the launcher) instead of an fp32 scratch + convert pass: half the reduce traffic, no zero-fill of 4 B/elt,
                         // no convert kernel; costs one bf16 rounding per key block (<= T/128) instead of one at the end
  uint32_t drop_thr8;
  float inv_keep;
  const uint64_t* rng;
  uint32_t op_id;
};

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmDQKV, const __grid_constant__ CUtensorMap tmDQacc, const AttnBwdParams p) {
  using L = AttnBwdSmem<HD>;
  constexpr int CH = L::CH;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smK = smem + L::K_OFF;
  uint8_t* smV = smem + L::V_OFF;
  uint8_t* smQ = smem + L::Q_OFF;
  uint8_t* smDO = smem + L::DO_OFF;
  uint8_t* smP = smem + L::P_OFF;
  uint8_t* smDS = smem + L::DS_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* kv_full = bars;        // TMA -> MMA (once)
  uint64_t* q_full = bars + 1;     // TMA -> MMA, stage 0 (stage 1 = bars + 14)
  uint64_t* do_full = bars + 2;    // TMA -> MMA
  uint64_t* q_empty = bars + 3;    // MMA commit (after dK) -> TMA, stage 0 (stage 1 = bars + 15)
  uint64_t* s_full = bars + 4;     // MMA -> softmax
  uint64_t* dp_full = bars + 5;    // MMA -> softmax
  uint64_t* s_empty = bars + 6;    // softmax(256) -> MMA: S columns read, S of the next iteration may be issued
  uint64_t* p_ready = bars + 7;    // softmax(128) -> MMA
  uint64_t* ds_ready = bars + 8;   // softmax(128) -> MMA
  uint64_t* pdv_done = bars + 9;   // MMA commit -> softmax (P smem free)
  uint64_t* dq_full = bars + 10;   // MMA commit -> softmax
  uint64_t* dq_drained = bars + 11;  // softmax(256) -> MMA: dQ drained, the dP/dQ columns are free
  uint64_t* do_empty = bars + 12;  // MMA commit (after dV) -> TMA
  uint64_t* q_full1 = bars + 14;
  uint64_t* q_empty1 = bars + 15;
  uint64_t* dk_done = bars + 16;   // MMA commit after the last dK: every MMA has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jb = blockIdx.x;                   // key block (low j = most work, scheduled first)
  const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H;
  const int D = p.H * HD;
  const int nq = p.T / ATT_BM;
  const int niter = nq - jb;                   // query blocks jb..nq-1
  const int krow = b * p.T + jb * ATT_BN;
  const int colQ = h * HD, colK = D + h * HD, colV = 2 * D + h * HD;

  if (warp == 0 && lane == 0) { prefetch_tmap(&tmQKV); prefetch_tmap(&tmDO); prefetch_tmap(&tmDQKV); prefetch_tmap(&tmDQacc); }
  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1); mbar_init(q_full, 1); mbar_init(do_full, 1); mbar_init(q_empty, 1); mbar_init(do_empty, 1);
    mbar_init(q_full1, 1); mbar_init(q_empty1, 1); mbar_init(dk_done, 1);
    mbar_init(s_full, 1); mbar_init(dp_full, 1); mbar_init(s_empty, 256); mbar_init(p_ready, 256);
    mbar_init(ds_ready, 256); mbar_init(pdv_done, 1); mbar_init(dq_full, 1); mbar_init(dq_drained, 256);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_enter();   // prologue above overlapped the previous kernel's tail; global memory is first touched below
  const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 384;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * L::TILE);
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        tma_load_2d(smK + c * 16384, &tmQKV, kv_full, colK + c * 64, krow);
        tma_load_2d(smV + c * 16384, &tmQKV, kv_full, colV + c * 64, krow);
      }
      for (int it = 0; it < niter; ++it) {
        const int qrow = b * p.T + (jb + it) * ATT_BM;
        {                                     // Q is double-buffered: stage it&1 was released by dK of iteration it-2
          uint64_t* qe = (it & 1) ? q_empty1 : q_empty;
          uint64_t* qf = (it & 1) ? q_full1 : q_full;
          mbar_wait(qe, ((it >> 1) & 1) ^ 1);
          mbar_expect_tx(qf, L::TILE);
#pragma unroll
          for (int c = 0; c < CH; ++c) tma_load_2d(smQ + (it & 1) * L::TILE + c * 16384, &tmQKV, qf, colQ + c * 64, qrow);
        }
        mbar_wait(do_empty, (it & 1) ^ 1);   // dO (single stage) is released by the dV MMA of the previous iteration
        mbar_expect_tx(do_full, L::TILE);
#pragma unroll
        for (int c = 0; c < CH; ++c) tma_load_2d(smDO + c * 16384, &tmDO, do_full, colQ + c * 64, qrow);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_t = make_idesc_bf16(128, HD, 1, 1);   // dV, dK: both operands MN-major
      constexpr uint32_t idesc_q = make_idesc_bf16(128, HD, 0, 1);   // dQ: dS K-major, K_j MN-major
      const uint32_t k_addr = smem_u32(smK), v_addr = smem_u32(smV), q_addr = smem_u32(smQ), do_addr = smem_u32(smDO);
      const uint32_t p_addr = smem_u32(smP), ds_addr = smem_u32(smDS);
      mbar_wait(kv_full, 0);
      // S_i = Q_i K^T is issued at the END of iteration i-1 (Q is double-buffered and dQ no longer shares
      // the S columns), so it runs on the tensor core while the softmax warps drain dQ_{i-1}.
      auto issue_s = [&](int it) {
        mbar_wait((it & 1) ? q_full1 : q_full, (it >> 1) & 1);
        if (it > 0) mbar_wait(s_empty, (it - 1) & 1);
        tc_fence_after_sync();
        const uint32_t qa = q_addr + (it & 1) * L::TILE;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_f16_ss(tS, make_smem_desc(qa + off, 16, 1024, kSwizzle128B), make_smem_desc(k_addr + off, 16, 1024, kSwizzle128B), idesc_s, k != 0);
        }
        umma_commit(s_full);
      };
      issue_s(0);
      for (int it = 0; it < niter; ++it) {
        const uint32_t ph = it & 1;
        const uint32_t qa = q_addr + (it & 1) * L::TILE;
        // dP = dO V^T  (its TMEM columns also receive dQ: wait until the previous dQ has been drained)
        mbar_wait(do_full, ph);
        if (it > 0) mbar_wait(dq_drained, (it - 1) & 1);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_f16_ss(tDP, make_smem_desc(do_addr + off, 16, 1024, kSwizzle128B), make_smem_desc(v_addr + off, 16, 1024, kSwizzle128B), idesc_s, k != 0);
        }
        umma_commit(dp_full);
        // dV += P~^T dO
        mbar_wait(p_ready, ph);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16_ss(tDV, make_smem_desc(p_addr + k * 2048, 16384, 1024, kSwizzle128B),
                      make_smem_desc(do_addr + k * 2048, 16384, 1024, kSwizzle128B), idesc_t, (it | k) != 0);
        umma_commit(pdv_done);
        umma_commit(do_empty);
        // dK += dS^T Q ;  dQ = dS K  (dQ lands in the dP columns: dP was consumed to build dS)
        mbar_wait(ds_ready, ph);
        tc_fence_after_sync();
        // dQ first: the softmax warps are waiting for it; dK and the next S then run under the dQ drain
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16_ss(tDP, make_smem_desc(ds_addr + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024, kSwizzle128B),
                      make_smem_desc(k_addr + k * 2048, 16384, 1024, kSwizzle128B), idesc_q, k != 0);
        umma_commit(dq_full);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16_ss(tDK, make_smem_desc(ds_addr + k * 2048, 16384, 1024, kSwizzle128B),
                      make_smem_desc(qa + k * 2048, 16384, 1024, kSwizzle128B), idesc_t, (it | k) != 0);
        umma_commit((it & 1) ? q_empty1 : q_empty);
        umma_commit(dk_done);
        if (it + 1 < niter) issue_s(it + 1);
      }
    }
  } else {
    // thread = (query row r = TMEM lane, column half hf) as in the forward kernel
    const int sw = warp - 2;
    const int quarter = warp & 3;
    const int hf = sw >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int etid = threadIdx.x - 64;
    constexpr int SPLIT = (HD == 64) ? 32 : 64;
    constexpr int FULLCH = HD / 64;
    const int hc0 = hf ? SPLIT : 0, hcn = hf ? HD - SPLIT : SPLIT;
    const float slope2 = p.slopes ? p.slopes[h] * 1.4426950408889634f : 0.f;
    const float kk = slope2 / p.scale_log2;
    const bool use_drop = p.drop_thr8 != 0;
    const float keep_scale = use_drop ? p.inv_keep : 1.f;
    uint64_t seed = 0, stream = 0;
    if (use_drop) { seed = p.rng[0]; stream = p.rng[1] * 65536ull + p.op_id; }
    Philox ph(seed);
    const int kv0 = jb * ATT_BN + hf * 64;
    const int cmax = r - hf * 64;
    uint8_t* stage = hf == 0 ? smP : smDS;   // dQ staging: half 0 -> P buffer, half 1 -> dS buffer
    // the per-row statistics of iteration it + 1 are fetched during iteration it (an L2 round trip is ~700 cycles,
    // which would otherwise sit in front of every P phase)
    float lse_nxt = p.lse[(size_t)bh * p.T + jb * ATT_BM + r];
    float delta_nxt = p.delta[(size_t)bh * p.T + jb * ATT_BM + r];
    for (int it = 0; it < niter; ++it) {
      const uint32_t phs = it & 1;
      const int qb = jb + it;
      const int qi = qb * ATT_BM + r;
      const size_t ridx = (size_t)bh * p.T + qi;
      const float lse2 = lse_nxt * 1.4426950408889634f;
      const float delta = delta_nxt;
      if (it + 1 < niter) {
        lse_nxt = p.lse[ridx + ATT_BM];
        delta_nxt = p.delta[ridx + ATT_BM];
      }
      const float kscale = keep_scale * p.scale, ndelta = -delta * p.scale;
      const uint64_t row_id = (uint64_t)bh * p.T + qi;
      // ---- P = exp2(S*c + alibi - lse)
      long long tk0 = clock64();
      mbar_wait(s_full, phs);
      long long tk1 = clock64();
      tc_fence_after_sync();
      float s[64];
      {
        uint32_t* sr = reinterpret_cast<uint32_t*>(s);
        tmem_ld_32x32b_x32(tS + lane_addr + hf * 64, sr);
        tmem_ld_32x32b_x32(tS + lane_addr + hf * 64 + 32, sr + 32);
        tmem_wait_ld();
      }
      tc_fence_before_sync();
      mbar_arrive(s_empty);
      const float base = slope2 * (float)(kv0 - qi) - lse2;
      if (it == 0) {  // diagonal block: causal mask
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          float v = ex2_ftz(fmaf(fmaf(kk, (float)c, s[c]), p.scale_log2, base));
          s[c] = c > cmax ? 0.f : v;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 64; ++c) s[c] = ex2_ftz(fmaf(fmaf(kk, (float)c, s[c]), p.scale_log2, base));
      }
      // Dropout decision folded into the sign of P: s[c] = +P (kept) or -P (dropped).  The packed P~ for the
      // dV MMA is then max(s, 0) (the 1/keep factor moves to the dV epilogue) and the dS pass reads the
      // decision back from the sign -- 3 instructions per element (byte-permute, compare, signed select)
      // instead of building and re-testing bit masks.
      if (use_drop) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 rb = attn_dropout_bits(ph, row_id, (uint32_t)(jb * 8 + hf * 4 + g), stream);
          uint32_t w[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const uint32_t byte = __byte_perm(w[e >> 2], 0u, 0x4440u | (uint32_t)(e & 3));
            const float v = s[g * 16 + e];
            s[g * 16 + e] = byte >= p.drop_thr8 ? v : -v;
          }
        }
      }
      // ---- P~ -> smem chunk hf.  The buffers doubled as dQ staging in the previous iteration: make sure
      // those bulk reduce-adds have finished reading smem (the wait is free by now -- it was issued
      // a whole S/P phase ago).
      if (it > 0) {
        if (etid == 0) tma_wait_group_read<0>();
        named_bar_sync(1, 256);
      }
      {
        uint8_t* prow = smP + hf * 16384 + r * 128;
#pragma unroll
        for (int c16 = 0; c16 < 8; ++c16) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = c16 * 8 + e * 2;
            pk[e] = pack_bf16(fmaxf(s[c], 0.f), fmaxf(s[c + 1], 0.f));
          }
          sts_v4(prow + ((c16 ^ (r & 7)) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
        }
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(p_ready);
      // ---- dS = P o (D o dP / keep - delta) * scale  -> smem chunk hf
      long long tk2 = clock64();
      mbar_wait(dp_full, phs);
      long long tk3 = clock64();
      tc_fence_after_sync();
      {
        uint8_t* drow = smDS + hf * 16384 + r * 128;
#pragma unroll
        for (int c32 = 0; c32 < 2; ++c32) {
          uint32_t dpv[32];
          tmem_ld_32x32b_x32(tDP + lane_addr + hf * 64 + c32 * 32, dpv);
          tmem_wait_ld();
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int cc = q4 * 8 + e * 2;
              // dS = P * (D o dP / keep - delta) * scale, constants folded; D is the sign of s (see above)
              const float s0 = s[c32 * 32 + cc], s1 = s[c32 * 32 + cc + 1];
              float d0 = s0 > 0.f ? fmaf(__uint_as_float(dpv[cc]), kscale, ndelta) : ndelta;
              float d1 = s1 > 0.f ? fmaf(__uint_as_float(dpv[cc + 1]), kscale, ndelta) : ndelta;
              pk[e] = pack_bf16(fabsf(s0) * d0, fabsf(s1) * d1);
            }
            const int c16 = c32 * 4 + q4;
            sts_v4(drow + ((c16 ^ (r & 7)) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(ds_ready);
      // ---- drain dQ: each half moves HD/2 fp32 columns TMEM -> its staging buffer -> TMA reduce-add.
      // P buffer is free once the dV MMA retired, dS buffer once dQ and dK retired (dq_full, dk_done).
      long long tk4 = clock64();
      mbar_wait(pdv_done, phs);
      mbar_wait(dq_full, phs);
      long long tk5 = clock64();
      tc_fence_after_sync();
      const int qrow0 = b * p.T + qb * ATT_BM;
      const bool dq16 = HD == 128 && p.dq_bf16;
      if (dq16) {
        // bf16 path: this half's 64 columns become one 128-byte row of its staging buffer (64-column bf16 box)
#pragma unroll
        for (int c32 = 0; c32 < 2; ++c32) {
          uint32_t o[32];
          tmem_ld_32x32b_x32(tDP + lane_addr + hc0 + c32 * 32, o);
          tmem_wait_ld();
          if (hf && c32 == 0) mbar_wait(dk_done, phs);   // half 1 stages into the dS buffer, which the dK MMA is still reading
          uint8_t* orow = stage + r * 128;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            sts_v4(orow + (((c32 * 4 + q4) ^ (r & 7)) << 4),
                   make_uint4(pack_bf16(__uint_as_float(o[q4 * 8 + 0]), __uint_as_float(o[q4 * 8 + 1])),
                              pack_bf16(__uint_as_float(o[q4 * 8 + 2]), __uint_as_float(o[q4 * 8 + 3])),
                              pack_bf16(__uint_as_float(o[q4 * 8 + 4]), __uint_as_float(o[q4 * 8 + 5])),
                              pack_bf16(__uint_as_float(o[q4 * 8 + 6]), __uint_as_float(o[q4 * 8 + 7]))));
        }
      } else {
#pragma unroll
        for (int c32 = 0; c32 < 2; ++c32) {
          if (c32 * 32 >= hcn) break;
          uint32_t o[32];
          tmem_ld_32x32b_x32(tDP + lane_addr + hc0 + c32 * 32, o);
          tmem_wait_ld();
          if (hf && c32 == 0) mbar_wait(dk_done, phs);   // half 1 stages into the dS buffer, which the dK MMA is still reading
          uint8_t* orow = stage + c32 * 16384 + r * 128;
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4)
            sts_v4(orow + ((q4 ^ (r & 7)) << 4), make_uint4(o[q4 * 4], o[q4 * 4 + 1], o[q4 * 4 + 2], o[q4 * 4 + 3]));
        }
      }
      tc_fence_before_sync();
      mbar_arrive(dq_drained);
      fence_proxy_async_smem();
      named_bar_sync(2, 256);
      if (etid == 0) {
        if (dq16) {
          tma_reduce_add_2d(&tmDQKV, smP, colQ, qrow0);
          tma_reduce_add_2d(&tmDQKV, smDS, colQ + 64, qrow0);
        } else {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh)
#pragma unroll
            for (int c32 = 0; c32 < 2; ++c32)
              if (c32 * 32 < (hh ? HD - SPLIT : SPLIT))
                tma_reduce_add_2d(&tmDQacc, (hh == 0 ? smP : smDS) + c32 * 16384, colQ + (hh ? SPLIT : 0) + c32 * 32, qrow0);
        }
        tma_commit_group();
      }
      if (p.dbg != nullptr && etid == 0) {
        long long tk6 = clock64();
        atomicAdd(p.dbg + 0, (unsigned long long)(tk1 - tk0));   // wait S
        atomicAdd(p.dbg + 1, (unsigned long long)(tk2 - tk1));   // P phase (ld S, exp, dropout, st.shared)
        atomicAdd(p.dbg + 2, (unsigned long long)(tk3 - tk2));   // wait dP
        atomicAdd(p.dbg + 3, (unsigned long long)(tk4 - tk3));   // dS phase
        atomicAdd(p.dbg + 4, (unsigned long long)(tk5 - tk4));   // wait dV done + dQ ready
        atomicAdd(p.dbg + 5, (unsigned long long)(tk6 - tk5));   // dQ drain
        atomicAdd(p.dbg + 6, 1ull);
      }
    }
    // ---- epilogue: dV, dK (TMEM) -> bf16 -> smem -> TMA store into the V / K slots of dqkv
    mbar_wait(dk_done, (niter - 1) & 1);  // last commit covers every MMA
    tc_fence_after_sync();
    if (etid == 0) tma_wait_group_read<0>();  // last dQ staging reads
    named_bar_sync(1, 256);
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const uint32_t tsrc = which == 0 ? tDV : tDK;
      uint8_t* stg = which == 0 ? smP : smDS;
      const int gcol0 = which == 0 ? colV : colK;
      const float osc = which == 0 ? keep_scale : 1.f;   // dV = (D o P)^T dO / keep: the 1/keep factor is applied here
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c * 32 >= hcn) break;
        uint32_t o[32];
        const int col = hc0 + c * 32;
        tmem_ld_32x32b_x32(tsrc + lane_addr + col, o);
        tmem_wait_ld();
        uint4 o4[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          o4[q4] = make_uint4(pack_bf16(__uint_as_float(o[q4 * 8 + 0]) * osc, __uint_as_float(o[q4 * 8 + 1]) * osc),
                              pack_bf16(__uint_as_float(o[q4 * 8 + 2]) * osc, __uint_as_float(o[q4 * 8 + 3]) * osc),
                              pack_bf16(__uint_as_float(o[q4 * 8 + 4]) * osc, __uint_as_float(o[q4 * 8 + 5]) * osc),
                              pack_bf16(__uint_as_float(o[q4 * 8 + 6]) * osc, __uint_as_float(o[q4 * 8 + 7]) * osc));
        if ((col >> 6) < FULLCH) {
          uint8_t* orow = stg + (col >> 6) * 16384 + r * 128;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) sts_v4(orow + (((((col & 63) >> 3) + q4) ^ (r & 7)) << 4), o4[q4]);
        } else {
          bf16* gp = p.dqkv + (size_t)(krow + r) * (3 * D) + gcol0 + col;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) st_v4(gp + q4 * 8, o4[q4]);
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(2, 256);
      if (etid == 0) {
        const int col = which == 0 ? colV : colK;
#pragma unroll
        for (int c = 0; c < FULLCH; ++c) tma_store_2d(&tmDQKV, stg + c * 16384, col + c * 64, krow);
        tma_commit_group();
      }
    }
    if (etid == 0) tma_wait_group<0>();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

// delta[b,h,t] = sum_d dO[b,t,h,d] * O[b,t,h,d].  16-byte loads; hd/8 lanes cooperate on one (b,t,h) row.
template <int HD>
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout,
                                                         float* __restrict__ delta, int B, int T, int H) {
  pdl_enter();
  constexpr int LPR = HD / 8;                       // lanes per row: 8, 12 or 16
  constexpr int RPW = 32 / LPR;                     // rows per warp pass: 4, 2, 2
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, li = lane % LPR;
  const long long nrows = (long long)B * T * H;
  long long row = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW + sub;
  float acc = 0.f;
  const bool active = sub < RPW && row < nrows;
  if (active) {
    uint4 a = ld_nc_v4(o + row * HD + li * 8), d = ld_nc_v4(dout + row * HD + li * 8);
    uint32_t as[4] = {a.x, a.y, a.z, a.w}, ds[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 x = unpack_bf16(as[k]), y = unpack_bf16(ds[k]);
      acc += x.x * y.x + x.y * y.y;
    }
  }
  // reduce over the LPR lanes of the row (LPR = 12 is not a power of two: go through 16-lane groups)
  if constexpr (LPR == 12) {
    // lanes [0,12) -> row 0, [12,24) -> row 1: shuffle-sum via explicit gathers
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 12; ++k) tot += __shfl_sync(0xffffffffu, acc, sub * 12 + k);
    acc = tot;
  } else {
#pragma unroll
    for (int off = LPR / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  }
  if (active && li == 0) {
    int h = (int)(row % H);
    long long bt = row / H;
    int t = (int)(bt % T), b = (int)(bt / T);
    delta[((size_t)b * H + h) * T + t] = acc;
  }
}

// dqkv[:, 0:D] = bf16(dq_acc)
__global__ void __launch_bounds__(256) attn_dq_convert_kernel(const float* __restrict__ acc, bf16* __restrict__ dqkv, size_t rows, int D8) {
  pdl_enter();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < rows * D8; i += (size_t)gridDim.x * blockDim.x) {
    size_t rrow = i / D8;
    int c = (int)(i % D8);
    const float4* src = reinterpret_cast<const float4*>(acc + (rrow * D8 + c) * 8);
    float4 a = src[0], b2 = src[1];
    st_v4(dqkv + rrow * (size_t)D8 * 24 + (size_t)c * 8, make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b2.x, b2.y), pack_bf16(b2.z, b2.w)));
  }
}

}  // namespace ztb

using namespace ztb;

int ztb_make_tmap_2d(CUtensorMap* tm, const void* ptr, int elem_bytes, long long rows, long long cols, long long ld,
                     int box_cols, int box_rows);

static int ztb_attn_variant() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("ZTB_ATTN_FWD"); v = e ? atoi(e) : 3; }
  return v;  // 1 = one CTA/SM (column-split softmax; the kernel head_dim 64 / 96 always use), 3 = two CTAs/SM (64-key blocks; head_dim 128 default)
}

template <int HD>
static int launch_fwd(const CUtensorMap& tq, const CUtensorMap& to, const AttnParams& p, cudaStream_t st) {
  auto kern = attn_fwd_kernel<HD>;
  static bool configured = false;
  if (!configured) {
    ZTB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<HD>::TOTAL));
    configured = true;
  }
  dim3 grid(p.T / ATT_BM, p.B * p.H);
  ZTB_CUDA_CHECK(ztb_launch(kern, grid, dim3(ATT_THREADS), AttnSmem<HD>::TOTAL, st, tq, to, p));
  return 0;
}

// qkv: [B*T, 3*H*hd] bf16 (q | k | v, heads contiguous inside each); out: [B*T, H*hd]; lse: [B, H, T]
int ztb_attn_fwd(const void* qkv, void* out, float* lse, const float* slopes, int B, int T, int H, int hd, float scale,
                 float dropout_p, const uint64_t* rng, uint32_t op_id, cudaStream_t st) {
  if (T % 128 != 0 || (hd != 64 && hd != 96 && hd != 128)) {
    ztb_set_error("ztb_attn_fwd", "needs T % 128 == 0 and head_dim in {64, 96, 128}", __FILE__, __LINE__);
    return -1;
  }
  const long long D = (long long)H * hd;
  CUtensorMap tq, to;
  int rc = ztb_make_tmap_2d(&tq, qkv, 2, (long long)B * T, 3 * D, 3 * D, 64, 128);
  if (rc) return rc;
  rc = ztb_make_tmap_2d(&to, out, 2, (long long)B * T, D, D, 64, 128);
  if (rc) return rc;
  AttnParams p;
  p.B = B; p.T = T; p.H = H;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.slopes = slopes; p.lse = lse;
  uint32_t thr = (uint32_t)(dropout_p * 256.f + 0.5f);
  if (thr > 255) thr = 255;
  p.drop_thr8 = dropout_p > 0.f ? (thr ? thr : 1) : 0;
  p.inv_keep = 256.f / (256.f - (float)p.drop_thr8);
  p.rng = rng; p.op_id = op_id;
  if (p.drop_thr8 && !rng) { ztb_set_error("ztb_attn_fwd", "dropout needs rng", __FILE__, __LINE__); return -1; }
  p.out = (bf16*)out;
  if (hd == 128 && ztb_attn_variant() == 3) {
    CUtensorMap tkv;
    rc = ztb_make_tmap_2d(&tkv, qkv, 2, (long long)B * T, 3 * D, 3 * D, 64, 64);
    if (rc) return rc;
    static bool configured = false;
    if (!configured) {
      ZTB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn3Smem::TOTAL));
      configured = true;
    }
    dim3 grid(T / ATT_BM, B * H);
    ZTB_CUDA_CHECK(ztb_launch(attn_fwd3_kernel, grid, dim3(ATT3_THREADS), Attn3Smem::TOTAL, st, tq, tkv, to, p));
    return 0;
  }
  return hd == 128 ? launch_fwd<128>(tq, to, p, st) : hd == 96 ? launch_fwd<96>(tq, to, p, st) : launch_fwd<64>(tq, to, p, st);
}


template <int HD>
static int launch_bwd(const CUtensorMap& tq, const CUtensorMap& tdo, const CUtensorMap& tdqkv, const CUtensorMap& tacc,
                      const AttnBwdParams& p, cudaStream_t st) {
  auto kern = attn_bwd_kernel<HD>;
  static bool configured = false;
  if (!configured) {
    ZTB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnBwdSmem<HD>::TOTAL));
    configured = true;
  }
  dim3 grid(p.T / ATT_BN, p.B * p.H);
  ZTB_CUDA_CHECK(ztb_launch(kern, grid, dim3(ATT_THREADS), AttnBwdSmem<HD>::TOTAL, st, tq, tdo, tdqkv, tacc, p));
  return 0;
}

// dqkv: [B*T, 3*D] bf16 (written in full); dq_acc: [B*T, D] fp32 scratch, must be zero on entry -- or nullptr with
// head_dim 128: dQ is then accumulated directly in dqkv in bf16 (see AttnBwdParams::dq_bf16); delta: [B,H,T] fp32 scratch.
int ztb_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* dq_acc, float* delta,
                 const float* slopes, int B, int T, int H, int hd, float scale, float dropout_p, const uint64_t* rng,
                 uint32_t op_id, int sms, cudaStream_t st, unsigned long long* dbg, int delta_ready) {
  if (T % 128 != 0 || (hd != 64 && hd != 96 && hd != 128)) {
    ztb_set_error("ztb_attn_bwd", "needs T % 128 == 0 and head_dim in {64, 96, 128}", __FILE__, __LINE__);
    return -1;
  }
  const long long D = (long long)H * hd, rows = (long long)B * T;
  if (!delta_ready) {
    const long long nr = rows * H;
    const int rpw = 32 / (hd / 8);
    const unsigned blocks = (unsigned)((nr + 8 * rpw - 1) / (8 * rpw));
    auto kd = hd == 128 ? attn_delta_kernel<128> : hd == 96 ? attn_delta_kernel<96> : attn_delta_kernel<64>;
    ZTB_CUDA_CHECK(ztb_launch(kd, dim3(blocks), dim3(256), 0, st, (const bf16*)out, (const bf16*)dout, delta, B, T, H));
  }
  CUtensorMap tq, tdo, tdqkv, tacc;
  int rc = ztb_make_tmap_2d(&tq, qkv, 2, rows, 3 * D, 3 * D, 64, 128);
  if (rc) return rc;
  rc = ztb_make_tmap_2d(&tdo, dout, 2, rows, D, D, 64, 128);
  if (rc) return rc;
  rc = ztb_make_tmap_2d(&tdqkv, dqkv, 2, rows, 3 * D, 3 * D, 64, 128);
  if (rc) return rc;
  if (dq_acc != nullptr) {
    rc = ztb_make_tmap_2d(&tacc, dq_acc, 4, rows, D, D, 32, 128);
    if (rc) return rc;
  } else {
    tacc = tdqkv;   // unused by the kernel in bf16-dQ mode
  }
  AttnBwdParams p;
  p.B = B; p.T = T; p.H = H;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.slopes = slopes; p.lse = lse; p.delta = delta;
  uint32_t thr = (uint32_t)(dropout_p * 256.f + 0.5f);
  if (thr > 255) thr = 255;
  p.drop_thr8 = dropout_p > 0.f ? (thr ? thr : 1) : 0;
  p.inv_keep = 256.f / (256.f - (float)p.drop_thr8);
  p.rng = rng; p.op_id = op_id;
  p.dqkv = (bf16*)dqkv;
  p.dbg = dbg;
  p.dq_bf16 = (hd == 128 && dq_acc == nullptr) ? 1 : 0;
  if (p.dq_bf16)   // the Q columns of dqkv are the reduce-add target: zero them (strided region: a 2-D memset, not a kernel)
    ZTB_CUDA_CHECK(cudaMemset2DAsync(dqkv, (size_t)3 * D * 2, 0, (size_t)D * 2, (size_t)rows, st));
  else if (dq_acc == nullptr) { ztb_set_error("ztb_attn_bwd", "dq_acc scratch is required unless head_dim == 128 (bf16 dQ accumulation)", __FILE__, __LINE__); return -1; }
  rc = hd == 128 ? launch_bwd<128>(tq, tdo, tdqkv, tacc, p, st) : hd == 96 ? launch_bwd<96>(tq, tdo, tdqkv, tacc, p, st) : launch_bwd<64>(tq, tdo, tdqkv, tacc, p, st);
  if (rc) return rc;
  if (p.dq_bf16) return 0;   // dQ is already in dqkv
  size_t n = (size_t)rows * (D / 8);
  size_t g = (n + 255) / 256;
  if (g > (size_t)sms * 8) g = (size_t)sms * 8;
  ZTB_CUDA_CHECK(ztb_launch(attn_dq_convert_kernel, dim3((unsigned)g), dim3(256), 0, st, (const float*)dq_acc, (bf16*)dqkv, (size_t)rows, (int)(D / 8)));
  return 0;
}
