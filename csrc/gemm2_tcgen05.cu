// 2-CTA (cta_group::2) variant of the tcgen05 GEMM: a pair of CTAs on one TPC computes a 256x256
// tile.  Each CTA stages 128 rows of A and 128 rows (N) of B per 64-wide K block -- 32 KB instead of
// the 48 KB a lone 128x256 CTA needs for the same MMA work -- so the L2 -> SM traffic per FLOP drops
// by a third and the smem ring deepens to 5 stages (the rest of the 227 KB is the 64 KB epilogue staging).  The leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256) which reads both CTAs' smem and writes both CTAs' TMEM;
// completion is multicast to both CTAs' mbarriers.  Everything else (operand majors, epilogues,
// TMEM double buffering, TMA store / reduce-add) matches gemm_tcgen05.cu.
// Reference call sites served: the q/k/v, attention-output and MLP projections and their autodiff transposes
// (src/models/layers.py:60-77,112-143,180-189), which XLA lowered to MXU dots.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "launch.cuh"
#include "launchers.h"

namespace ztb {
namespace g2 {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 5;
constexpr int TILE_M = 256;                  // rows per CTA pair
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_STAGE_BYTES = 128 * BK * 2;  // 16 KB: this CTA's half of the 256 B rows
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int EPI_WARPS = 8;                 // two warps per TMEM lane quarter, alternating 64-column chunks
constexpr int WBUF_BYTES = 32 * 128;         // one warp's staging buffer: 32 rows x 128 B
constexpr int CSTAGE_BYTES = EPI_WARPS * 2 * WBUF_BYTES;  // 64 KB: double-buffered per warp
constexpr int GEMM_SMEM = STAGES * STAGE_BYTES + CSTAGE_BYTES + 256 + 1024;
constexpr int GEMM_THREADS = 128 + EPI_WARPS * 32;
// Register cap: 12 warps x 128 registers = 48 K of the SM's 64 K, so a 128-thread CTA of the overlapped ZeRO-1 update
// kernel (<= 64 registers per thread, no shared memory) can be co-resident with a GEMM CTA on every SM.
constexpr int GEMM_MAX_REGS = 128;
constexpr int GROUP_M = 8;                   // in 256-row tiles
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address

struct GemmParams {
  int M, N, K;
  int epi, c_fp32, accumulate;
  const bf16* aux_in;
  long long ld_aux;
  int has_aux_out;
  float p, inv_keep;
  const uint64_t* rng;
  uint32_t op_id;
  int tiles_m, tiles_n, num_tiles;
  // Tail splitting: the first `full_tiles` tiles (a multiple of the pair count) are whole-K work units;
  // each of the remaining `tail` tiles is cut into `split` K-ranges whose partial sums meet in C through
  // bulk reduce-adds (C pre-initialised by the launcher).  split == 1: every tile is one unit.
  int full_tiles, tail, split;
  int* chain;   // split > 1: [tail][16] per-(tile, CTA, epilogue warp) count of K-ranges already reduced into C
  // Output paths that skip the shared-memory staging (each epilogue thread owns 64 contiguous bytes of one row per
  // 32-column half): bit 0 = the auxiliary output (pre-activation Z), bit 1 = a non-accumulating bf16 C.
  int direct;
  bf16* c_raw; bf16* aux_raw; long long ldc_raw, ld_aux_raw;
  float* rowdot_out; int rowdot_T, rowdot_hd, rowdot_H;   // ROWDOT epilogue (see ZtbGemmArgs)
  unsigned int* mask_out; int mask_ld;                     // DROPOUT_ADD: keep bits, mask_ld = words per row
  int raster_n;   // 1: walk tiles in groups of GROUP_M column blocks (vocabulary-sized N: the A panel stays in L2)
                  // 2: all row blocks of one column block, column blocks rotated by nb_shift (all-gather variant: the
                  //    column blocks this rank owns come first)
  int nb_shift;
};

struct Unit { int tile, kb0, kb1, partial, part; };
// Units are dealt round-robin: unit u of a pair is pair + i * num_pairs.  Tail units are ordered K-range
// major, so the pairs working concurrently on tiles that share an A/B panel read the same K range (L2 reuse).
ZTB_DEVINL bool get_unit(int u, const GemmParams& p, int kblocks, Unit& w) {
  if (u < p.full_tiles) { w.tile = u; w.kb0 = 0; w.kb1 = kblocks; w.partial = 0; w.part = 0; return true; }
  const int v = u - p.full_tiles;
  if (v >= p.tail * p.split) return false;
  const int part = v / p.tail;
  w.tile = p.full_tiles + v % p.tail;
  w.kb0 = (int)((long long)kblocks * part / p.split);
  w.kb1 = (int)((long long)kblocks * (part + 1) / p.split);
  w.partial = p.split > 1;
  w.part = part;
  return true;
}

ZTB_DEVINL void decode_tile(int t, const GemmParams& p, int& mb, int& nb) {
  if (p.raster_n == 2) {
    mb = t % p.tiles_m;
    nb = t / p.tiles_m + p.nb_shift;
    if (nb >= p.tiles_n) nb -= p.tiles_n;
    return;
  }
  if (p.raster_n) {
    int per_group = GROUP_M * p.tiles_m;
    int g = t / per_group, r = t % per_group;
    int gn = min(GROUP_M, p.tiles_n - g * GROUP_M);
    nb = g * GROUP_M + r % gn;
    mb = r / gn;
    return;
  }
  int per_group = GROUP_M * p.tiles_n;
  int g = t / per_group, r = t % per_group;
  int gm = min(GROUP_M, p.tiles_m - g * GROUP_M);
  mb = g * GROUP_M + r % gm;
  nb = r / gm;
}

ZTB_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
ZTB_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
ZTB_DEVINL void tma_load_2d_2sm(void* smem_dst, const void* tmap, uint64_t* leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(leader_bar) & kPeerMask), "r"(c0), "r"(c1)
      : "memory");
}
ZTB_DEVINL void umma_f16_ss_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
ZTB_DEVINL void umma_commit_2sm(uint64_t* bar) {  // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
ZTB_DEVINL int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
ZTB_DEVINL void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
ZTB_DEVINL void mbar_arrive_cta0(uint64_t* bar) {  // arrive on CTA 0's copy of `bar`
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(smem_u32(bar)));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// All-gather fused into the GEMM (ZeRO-1 parameter all-gather + the first GEMM that consumes the gathered weight;
// SURVEY 2.4-B X3, reference main_zero.py:452-456,500 where XLA inserts an fp32 all-gather in front of the step).
// B is a weight whose row chunks are owned by different ranks.  One spare warp of EVERY CTA is a copy engine driver: an
// elected thread streams the remote chunks of the weight's whole ZeRO-1 bucket from the owners' symmetric buffers into
// the local flat parameter buffer with bulk async copies (peer global -> shared ring -> local global; the TMA unit
// moves the bytes, no register round trip), B's own chunks first.  Meanwhile the tensor cores work through the N blocks
// this rank owns; the TMA producers of the other N blocks acquire a per-owner "B chunk landed" counter first.
constexpr int AG_TILE_ELEMS = 4096;                 // 8 KB per bulk copy
constexpr int AG_TILE_BYTES = AG_TILE_ELEMS * 2;
constexpr int AG_SLOTS_MAX = 8;                     // gather-ring slots: 4 (AG = 1: five operand stages) or 8 (AG = 2: four stages, 7 x 8 KB copies in flight per CTA)
constexpr int kAgMaxSegs = 32;                     // pull-list entries staged in shared memory (8 ranks x 4 matrices = 28)
constexpr int STAGES_AG = 5;                        // the AG variant keeps all five operand stages (a 4-stage ring cost ~15 %) and takes the
                                                    // gather ring's 32 KB from the epilogue staging instead: one output buffer per warp

struct AgParams {
  const void* peer[8];          // flat bf16 parameter buffer of every rank (peer-mapped), index = rank
  bf16* local;                  // this rank's flat parameter buffer
  const ZtbSegment* segs;       // pull list: flat_off, len, decay = source rank, tile_begin (AG_TILE_ELEMS units); B's segments first
  int nseg;
  int n_b_segs;                 // the first n_b_segs segments are B's chunks (one per remote rank)
  long long total_tiles;
  const int* shard_ready;       // my flag row: [rank] >= epoch once that rank's chunk is final
  int epoch;
  const int* epoch_ptr;         // non-null: wait for *epoch_ptr + epoch (device-side step counter, CUDA-graph replays)
  unsigned int* done;           // [8] per-owner count of landed B tiles (zeroed before launch)
  unsigned int b_tiles_per_rank;  // counter value that means "this owner's B chunk has landed" (cumulative counters: base + tiles)
  int rows_per_rank;            // B rows owned by each rank
  int world, rank;
};

ZTB_DEVINL int ld_acquire_sys_s32(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
ZTB_DEVINL unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// 1-D bulk async copies (the TMA unit without a tensor map): global -> shared with mbarrier completion, shared -> global
ZTB_DEVINL void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
ZTB_DEVINL void bulk_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}

// L2 evict-first variants for the part of the bucket this GEMM does not consume (out-proj / fc_in / fc_out chunks): they
// stream through without displacing the A panel the tiles re-read from L2.
ZTB_DEVINL uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
ZTB_DEVINL void bulk_load_1d_hint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
ZTB_DEVINL void bulk_store_1d_hint(void* gdst, const void* smem_src, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes), "l"(pol) : "memory");
}

template <bool A_MN, bool B_MN, int EPI, int AG>   // AG: 0 = plain, 1 / 2 = all-gather variant (see AG_SLOTS_MAX)
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(GEMM_MAX_REGS)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAux,
                     const __grid_constant__ CUtensorMap tmAuxIn, const GemmParams p, const __grid_constant__ AgParams ag) {
  // Epilogues that read a second operand (residual, pre-activation Z, attention output O) fetch it with per-warp TMA
  // loads into the warp's first staging buffer.  Round 1 / early round 2 read it with per-thread 16-byte global loads,
  // one ROW per lane: 32 different cache lines per load instruction = 32 L1 wavefronts each, ~4000 L1/shared-pipeline
  // cycles per tile on the same SRAM datapath the tensor core reads its operands from (GELU-backward dgrad: 1.05 PFLOP/s
  // against 1.34 for the same GEMM without the read).  A TMA load costs that pipeline nothing; reading the row back from
  // swizzled shared memory is 4 conflict-free wavefronts per instruction.
  constexpr bool kAuxTma = EPI == ZTB_EPI_DROPOUT_ADD || EPI == ZTB_EPI_GELU_BWD || EPI == ZTB_EPI_ROWDOT;
  constexpr int NST = AG == 2 ? STAGES - 1 : (AG ? STAGES_AG : STAGES);
  constexpr int AG_SLOTS = AG == 2 ? 8 : 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + NST * A_STAGE_BYTES;
  uint8_t* smC = smem + NST * STAGE_BYTES;
  uint8_t* smG = smC + CSTAGE_BYTES / 2;     // AG only: AG_SLOTS x 8 KB gather ring (the second half of the epilogue staging area)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + CSTAGE_BYTES);
  uint64_t* full = bars;                     // [NST]  leader's copy collects both CTAs' TMA bytes
  uint64_t* empty = bars + STAGES;           // [NST]  multicast commit -> both CTAs
  uint64_t* tfull = bars + 2 * STAGES;       // [2]       multicast commit -> both CTAs
  uint64_t* tempty = bars + 2 * STAGES + 2;  // [2]       leader's copy: 8 epilogue warps (both CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint64_t* gfull = bars + 2 * STAGES + 5;   // [AG_SLOTS] gather ring: bulk load landed
  uint64_t* abar = bars + 2 * STAGES + 5 + AG_SLOTS_MAX;   // [EPI_WARPS] per-warp "auxiliary operand chunk landed"
  static_assert(AG_SLOTS * AG_TILE_BYTES == CSTAGE_BYTES / 2 + (STAGES - NST) * STAGE_BYTES || !AG, "the gather ring takes half of the epilogue staging area (+ one operand stage for AG = 2)");
  static_assert((2 * STAGES + 5 + AG_SLOTS_MAX + EPI_WARPS) * 8 <= 256, "barrier block");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int kblocks = (p.K + BK - 1) / BK;

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB); prefetch_tmap(&tmC);
    if (p.has_aux_out) prefetch_tmap(&tmAux);
    if constexpr (kAuxTma) prefetch_tmap(&tmAuxIn);
  }
  if (warp == 9 && lane == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * EPI_WARPS); }
    if constexpr (AG) { for (int i = 0; i < AG_SLOTS; ++i) mbar_init(&gfull[i], 1); }
    if constexpr (kAuxTma) { for (int i = 0; i < EPI_WARPS; ++i) mbar_init(&abar[i], 1); }
    fence_barrier_init();
  }
  if (warp == 10) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_enter();   // everything above overlapped the previous kernel's tail; no global memory has been touched yet

  // Control warps take the HIGHEST warp ids: the SMSP arbiter favours high warp ids, so the TMA
  // producer and the MMA issuer are never starved by the ALU-heavy epilogue warps (0-7).
  if (warp == 8) {
    // ===================================================== TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      Unit w;
      [[maybe_unused]] unsigned owners_ready = AG ? (1u << ag.rank) : 0u;
      for (int u = pair; get_unit(u, p, kblocks, w); u += num_pairs) {
        int mb, nb; decode_tile(w.tile, p, mb, nb);
        const int m0 = mb * TILE_M + cta_rank * BM;        // this CTA's A rows
        const int nrow0 = nb * BN + cta_rank * 128;        // this CTA's B rows
        if constexpr (AG) {
          // rows [nrow0, nrow0 + 128) of B: every owner other than this rank must have landed in the local buffer.  An owner
          // is checked until it has been seen complete once (a ~1 us acquire per tile was 20+ us of producer stalls per GEMM).
          int r0 = nrow0 / ag.rows_per_rank, r1 = (nrow0 + 127) / ag.rows_per_rank;
          if (r0 >= ag.world) r0 = ag.world - 1;
          if (r1 >= ag.world) r1 = ag.world - 1;
          bool fresh = false;
          for (int o = r0; o <= r1; ++o)
            if (!((owners_ready >> o) & 1u)) {
              while ((int)((unsigned)ld_acquire_gpu(reinterpret_cast<const int*>(ag.done + o)) - ag.b_tiles_per_rank) < 0) __nanosleep(64);
              owners_ready |= 1u << o;
              fresh = true;
            }
          if (fresh) asm volatile("fence.proxy.async.global;" ::: "memory");   // the gathered rows were written by bulk copies of other CTAs
        }
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (leader) mbar_expect_tx(&full[stage], 2 * STAGE_BYTES);
          uint8_t* a = smA + stage * A_STAGE_BYTES;
          uint8_t* b = smB + stage * B_STAGE_BYTES;
          if constexpr (!A_MN) {
            tma_load_2d_2sm(a, &tmA, &full[stage], kb * BK, m0);
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_2d_2sm(a + j * 8192, &tmA, &full[stage], m0 + j * 64, kb * BK);
          }
          if constexpr (!B_MN) {
            tma_load_2d_2sm(b, &tmB, &full[stage], kb * BK, nrow0);
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_2d_2sm(b + j * 8192, &tmB, &full[stage], nrow0 + j * 64, kb * BK);
          }
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    // ===================================================== MMA issuer (leader CTA, single thread)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BN, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      Unit w;
      for (int u = pair; get_unit(u, p, kblocks, w); u += num_pairs) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smA + stage * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(smB + stage * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            uint64_t da = A_MN ? make_smem_desc(a_addr + k * 2048, 8192, 1024, kSwizzle128B)
                               : make_smem_desc(a_addr + k * 32, 16, 1024, kSwizzle128B);
            uint64_t db = B_MN ? make_smem_desc(b_addr + k * 2048, 8192, 1024, kSwizzle128B)
                               : make_smem_desc(b_addr + k * 32, 16, 1024, kSwizzle128B);
            umma_f16_ss_2sm(d_tmem, da, db, idesc, (kb != w.kb0 || k != 0) ? 1u : 0u);
          }
          umma_commit_2sm(&empty[stage]);
          if (kb == w.kb1 - 1) umma_commit_2sm(&tfull[acc]);
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
        acc ^= 1; if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (AG && warp == 11) {
    // ===================================================== all-gather driver (one thread per CTA, bulk copies over NVLink)
    if constexpr (AG) {
      // The pull list is consulted three times per 8 KB copy (issue, store, publish): from global memory every lookup is a
      // chain of dependent ~0.6 us loads and the single driver thread, not NVLink, limited the gather (measured: fused GEMM
      // = plain GEMM + 0.8 x the standalone pull time at n = 2).  Stage the list in shared memory once.
      __shared__ ZtbSegment seg_s[kAgMaxSegs];
      const bool segs_staged = ag.nseg <= kAgMaxSegs;
      if (segs_staged) {
        const int words = ag.nseg * (int)(sizeof(ZtbSegment) / 8);
        const unsigned long long* gsrc = reinterpret_cast<const unsigned long long*>(ag.segs);
        unsigned long long* sdst = reinterpret_cast<unsigned long long*>(seg_s);
        for (int i = lane; i < words; i += 32) sdst[i] = gsrc[i];
      }
      __syncwarp();
      const ZtbSegment* seg_tab = segs_staged ? seg_s : ag.segs;
      if (lane == 0) {
        unsigned seen = 1u << ag.rank;
        const uint64_t evict_first = l2_evict_first_policy();
        const int want_epoch = ag.epoch + (ag.epoch_ptr != nullptr ? *ag.epoch_ptr : 0);
        // Ring of AG_SLOTS 8 KB slots.  Copy i uses slot i % AG_SLOTS: bulk-load peer -> slot (mbarrier), bulk-store
        // slot -> local buffer (bulk group).  Loads run AG_SLOTS - 1 ahead of the stores; a slot is reloaded once the
        // store that read it has finished reading shared memory; a B tile is published to the TMA producers once its
        // store has completed (bulk groups complete in order, so wait_group<k> tells how far completion got).
        long long issued = 0, stored = 0, published = 0;
        const long long first = blockIdx.x, step = gridDim.x;
        const long long mine = ag.total_tiles > first ? (ag.total_tiles - first + step - 1) / step : 0;
        auto tile_info = [&](long long i, const bf16*& src, bf16*& dst, uint32_t& bytes, int& owner, bool& is_b) {
          const long long t = first + i * step;
          int lo = 0, hi = ag.nseg - 1;
          while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (seg_tab[mid].tile_begin <= t) lo = mid; else hi = mid - 1;
          }
          const ZtbSegment sg = seg_tab[lo];
          const long long e0 = (t - sg.tile_begin) * AG_TILE_ELEMS;
          const long long left = sg.len - e0;
          bytes = (uint32_t)((left < AG_TILE_ELEMS ? left : AG_TILE_ELEMS) * 2);
          owner = sg.decay; is_b = lo < ag.n_b_segs;
          src = reinterpret_cast<const bf16*>(ag.peer[owner]) + sg.flat_off + e0;
          dst = ag.local + sg.flat_off + e0;
        };
        auto issue_load = [&](long long i) {
          const bf16* src; bf16* dst; uint32_t bytes; int owner; bool is_b;
          tile_info(i, src, dst, bytes, owner, is_b);
          if (!((seen >> owner) & 1u)) {   // the owner has published this step's chunk?
            const unsigned long long t0 = globaltimer_ns();
            while (ld_acquire_sys_s32(ag.shard_ready + owner) < want_epoch) {
              __nanosleep(64);
              if (globaltimer_ns() - t0 > 120ull * 1000000000ull) __trap();   // owner never published its shard: fail loudly
            }
            seen |= 1u << owner;
          }
          const int slot = (int)(i % AG_SLOTS);
          mbar_expect_tx(&gfull[slot], bytes);
          if (is_b) bulk_load_1d(smG + slot * AG_TILE_BYTES, src, bytes, &gfull[slot]);
          else bulk_load_1d_hint(smG + slot * AG_TILE_BYTES, src, bytes, &gfull[slot], evict_first);
        };
        auto publish = [&](long long i) {   // store i has completed
          const bf16* src; bf16* dst; uint32_t bytes; int owner; bool is_b;
          tile_info(i, src, dst, bytes, owner, is_b);
          if (is_b) {
            asm volatile("fence.proxy.async.global;" ::: "memory");
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ag.done + owner) : "memory");
          }
        };
        while (issued < mine && issued < AG_SLOTS - 1) issue_load(issued++);
        for (; stored < mine; ++stored) {
          const int slot = (int)(stored % AG_SLOTS);
          mbar_wait(&gfull[slot], (uint32_t)((stored / AG_SLOTS) & 1));
          const bf16* src; bf16* dst; uint32_t bytes; int owner; bool is_b;
          tile_info(stored, src, dst, bytes, owner, is_b);
          if (is_b) bulk_store_1d(dst, smG + slot * AG_TILE_BYTES, bytes);
          else bulk_store_1d_hint(dst, smG + slot * AG_TILE_BYTES, bytes, evict_first);
          tma_commit_group();
          if (issued < mine) {
            // copy `issued` reuses the slot that store `stored - 1` read: only the store just issued may still be reading
            tma_wait_group_read<1>();
            issue_load(issued++);
          }
          // stores complete in order: all but the newest AG_SLOTS - 1 have finished
          tma_wait_group<AG_SLOTS - 1>();
          for (; published + (AG_SLOTS - 1) <= stored; ++published) publish(published);
        }
        tma_wait_group<0>();
        for (; published < mine; ++published) publish(published);
      }
    }
  } else if (warp < 8) {
    // ===================================================== epilogue (8 warps per CTA, its 128 rows)
    // A warp drains two 64-column chunks of its 32 accumulator rows per tile, each in two 32-column halves
    // (one tcgen05.ld.x32 = 32 live fp32 values per thread), so the epilogue stays well below 128 registers and
    // the CTA leaves room for a small co-resident CTA of another kernel (the overlapped ZeRO-1 update).
    const int e = warp;                     // 0..7
    const int ew = e & 3;                   // TMEM lane quarter (== warp % 4)
    const int chalf = e >> 2;               // this warp handles 64-column chunks chalf, chalf + 2
    const int row = ew * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
    uint8_t* wbuf = smC + e * (AG ? 1 : 2) * WBUF_BYTES;   // this warp's two staging buffers (one in the all-gather variant)
    uint32_t thr = 0;
    uint64_t stream = 0;
    Philox ph(0);
    if constexpr (EPI == ZTB_EPI_DROPOUT_ADD) {
      thr = dropout_threshold(p.p);
      ph = Philox(p.rng[0]);
      stream = p.rng[1] * 65536ull + p.op_id;
    }
    auto arrive_tempty = [&](int a) { if (leader) mbar_arrive(&tempty[a]); else mbar_arrive_cta0(&tempty[a]); };
    int acc = 0; uint32_t acc_phase = 0;
    int cbuf = kAuxTma ? 1 : 0;             // aux epilogues: buffer 0 receives the auxiliary operand, buffer 1 stages the output
    uint32_t aphase = 0;
    auto chunk_of = [&](int ci) { return EPI == ZTB_EPI_ROWDOT ? chalf * 2 + ci : chalf + 2 * ci; };
    // TMA load of this warp's 32 x 64 auxiliary chunk for (unit, ci) into buffer 0 (rows / columns outside the tensor arrive
    // as zeros, so the consumers need no bounds checks)
    auto issue_aux = [&](const Unit& uu, int ci) {
      if (lane == 0) {
        int mb2, nb2; decode_tile(uu.tile, p, mb2, nb2);
        fence_proxy_async_smem();          // the generic-proxy reads of the previous chunk are complete (syncwarp) before the async write
        mbar_expect_tx(&abar[e], WBUF_BYTES);
        tma_load_2d(wbuf, &tmAuxIn, &abar[e], nb2 * BN + chunk_of(ci) * 64, mb2 * TILE_M + (int)cta_rank * BM + ew * 32);
      }
    };
    Unit w;
    if constexpr (kAuxTma) {
      if (get_unit(pair, p, kblocks, w)) issue_aux(w, 0);
    }
    for (int u = pair; get_unit(u, p, kblocks, w); u += num_pairs) {
      int mb, nb; decode_tile(w.tile, p, mb, nb);
      const int m0 = mb * TILE_M + cta_rank * BM, n0 = nb * BN;
      // K-ranges of a split tile meet in C in K order: range 0 stores (or reduce-adds when the GEMM accumulates),
      // every later range reduce-adds after the same region of the previous range has landed, so the bf16/fp32 sum
      // is reproducible run to run and C needs no zero-fill.  The previous range was finished by a different CTA
      // pair, in the same or the previous round -- every CTA is resident, so this cannot deadlock.
      const bool reduce_c = p.accumulate || w.part > 0;
      int* chain = w.partial ? p.chain + ((w.tile - p.full_tiles) * 16 + (int)cta_rank * 8 + e) : nullptr;
      const int grow = m0 + row;
      const bool direct_c = (p.direct & 2) && !p.c_fp32 && !reduce_c && !w.partial;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after_sync();
      if (chain != nullptr && w.part > 0) {
        if (lane == 0) {
          while (ld_acquire_gpu(chain) < w.part) __nanosleep(32);
          asm volatile("fence.proxy.async.global;" ::: "memory");
        }
        __syncwarp();
      }
      const uint32_t t_acc = tmem_base + lane_addr + acc * BN;
      // One staging buffer holds 128 B per row: 64 bf16 columns or 32 fp32 columns.  Buffers alternate; before
      // overwriting one, the issuing thread waits until at most one bulk store (the other buffer's) is still
      // reading smem.  Each epilogue warp owns its 32 rows of the staging buffer and issues its own bulk stores
      // ({64|32} x 32-row boxes), so the only synchronisation is warp-local.
      auto begin_buffer = [&]() {
        if (lane == 0) { if constexpr (kAuxTma || AG) tma_wait_group_read<0>(); else tma_wait_group_read<1>(); }
        __syncwarp();
      };
      auto end_buffer = [&](const CUtensorMap* tm, int col, bool reduce) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const uint8_t* src = wbuf + cbuf * WBUF_BYTES;
          if (reduce) tma_reduce_add_2d(tm, src, col, m0 + ew * 32);
          else tma_store_2d(tm, src, col, m0 + ew * 32);
          tma_commit_group();
        }
        if constexpr (!kAuxTma && !AG) cbuf ^= 1;
      };
      // this thread's 16-byte piece q of half h of its auxiliary row (buffer 0, same 128-byte swizzle as the staging rows)
      auto aux_piece = [&](int h, int q) {
        uint4 u;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                     : "r"(smem_u32(wbuf + lane * 128 + (((h * 4 + q) ^ (row & 7)) << 4))));
        return u;
      };
      // 16 packed bf16x2 words (32 columns) -> half h of this thread's swizzled 128-byte row
      auto stage_packed_half = [&](const uint32_t* pk, int h) {
        uint8_t* rowp = wbuf + cbuf * WBUF_BYTES + lane * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sts_v4(rowp + (((h * 4 + j) ^ (row & 7)) << 4), make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]));
      };
      // no staging: this thread owns 32 consecutive columns of one row = 64 contiguous bytes
      auto store_direct_half = [&](bf16* base, long long ld, const uint32_t* pk, int col) {
        if (grow >= p.M) return;
        bf16* dst = base + (size_t)grow * ld + col;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (col + 8 * j < p.N) st_v4(dst + 8 * j, make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]));
      };
      float rowdot = 0.f;
#pragma unroll 1
      for (int ci = 0; ci < 2; ++ci) {
        // ROWDOT: a warp takes two ADJACENT chunks (= one 128-wide head, or two 64-wide heads) so a head's row sum never
        // leaves the thread; every other epilogue interleaves the two warps of a lane quarter chunk by chunk
        const int c = chunk_of(ci);
        const int gcol = n0 + c * 64;
        if constexpr (kAuxTma) { mbar_wait(&abar[e], aphase); aphase ^= 1; }
        uint32_t zp[32];   // GELU with a staged auxiliary output: the packed pre-activation waits for the second buffer
        if (!p.c_fp32 && !direct_c) begin_buffer();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_acc + c * 64 + h * 32, r);
          tmem_wait_ld();
          if (ci == 1 && h == 1) {  // this warp's last TMEM read of the accumulator: hand it back to the MMA warp
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) arrive_tempty(acc);
          }
          const int hcol = gcol + h * 32;
          float v[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
          if constexpr (EPI == ZTB_EPI_GELU) {
            if (p.has_aux_out) {  // pre-activation Z
              uint32_t* zq = zp + h * 16;
#pragma unroll
              for (int k = 0; k < 16; ++k) zq[k] = pack_bf16(v[2 * k], v[2 * k + 1]);
              if (p.direct & 1) store_direct_half(p.aux_raw, p.ld_aux_raw, zq, hcol);
            }
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = gelu_tanh(v[k]);
          } else if constexpr (EPI == ZTB_EPI_ROWDOT) {
            // delta partial: dot of the bf16-ROUNDED output row (what attention backward will read as dO) with aux_in (O)
            const bool in_rows = grow < p.M;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 u = aux_piece(h, q);
              uint32_t us[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float2 of = unpack_bf16(us[k]);
                float2 df = unpack_bf16(pack_bf16(v[q * 8 + 2 * k], v[q * 8 + 2 * k + 1]));
                rowdot = fmaf(df.x, of.x, fmaf(df.y, of.y, rowdot));
              }
            }
            if (h == 1 && (p.rowdot_hd == 64 || ci == 1)) {   // a head's columns are complete
              const int head = hcol / p.rowdot_hd;
              if (in_rows && head < p.rowdot_H) {
                const int bb = grow / p.rowdot_T, tt = grow - bb * p.rowdot_T;
                p.rowdot_out[((size_t)bb * p.rowdot_H + head) * p.rowdot_T + tt] = rowdot;
              }
              rowdot = 0.f;
            }
          } else if constexpr (EPI == ZTB_EPI_DROPOUT_ADD || EPI == ZTB_EPI_GELU_BWD) {
            uint4 au[4];
            uint32_t keep_bits = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) au[q] = aux_piece(h, q);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t us[4] = {au[q].x, au[q].y, au[q].z, au[q].w};
              if constexpr (EPI == ZTB_EPI_DROPOUT_ADD) {
                size_t idx = (size_t)grow * p.N + hcol + q * 8;
                uint4 r0 = ph.rounds<kPhiloxRounds>(idx >> 2, stream), r1 = ph.rounds<kPhiloxRounds>((idx >> 2) + 1, stream);
                uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  float2 rf = unpack_bf16(us[k]);
                  const bool k0 = rr[2 * k] >= thr, k1 = rr[2 * k + 1] >= thr;
                  keep_bits |= (k0 ? 1u : 0u) << (q * 8 + 2 * k) | (k1 ? 1u : 0u) << (q * 8 + 2 * k + 1);
                  float a = k0 ? v[q * 8 + 2 * k] * p.inv_keep : 0.f;
                  float b = k1 ? v[q * 8 + 2 * k + 1] * p.inv_keep : 0.f;
                  v[q * 8 + 2 * k] = rf.x + a;
                  v[q * 8 + 2 * k + 1] = rf.y + b;
                }
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float2 gr = gelu_tanh_grad_pair(us[k]);   // packed fp16 evaluation (common.cuh)
                  v[q * 8 + 2 * k] *= gr.x;
                  v[q * 8 + 2 * k + 1] *= gr.y;
                }
              }
            }
            if constexpr (EPI == ZTB_EPI_DROPOUT_ADD) {
              if (p.mask_out != nullptr && grow < p.M && hcol < p.N) p.mask_out[(size_t)grow * p.mask_ld + (hcol >> 5)] = keep_bits;
            }
          }
          if (p.c_fp32) {   // 32 fp32 columns fill a 128-byte row: one buffer and one bulk store per half
            begin_buffer();
            uint8_t* rowp = wbuf + cbuf * WBUF_BYTES + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts_v4(rowp + ((j ^ (row & 7)) << 4), make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                                                                __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3])));
            end_buffer(&tmC, hcol, reduce_c);
          } else {
            uint32_t pk[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) pk[k] = pack_bf16(v[2 * k], v[2 * k + 1]);
            if (direct_c) store_direct_half(p.c_raw, p.ldc_raw, pk, hcol);
            else stage_packed_half(pk, h);
          }
        }
        if constexpr (kAuxTma) {
          __syncwarp();                    // every lane has consumed its auxiliary row: buffer 0 may be refilled
          Unit wn;
          if (ci == 0) issue_aux(w, 1);
          else if (get_unit(u + num_pairs, p, kblocks, wn)) issue_aux(wn, 0);
        }
        if (!p.c_fp32 && !direct_c) end_buffer(&tmC, gcol, reduce_c);
        if constexpr (EPI == ZTB_EPI_GELU) {
          if (p.has_aux_out && !(p.direct & 1)) {   // the pre-activation leaves through its own tensor map
            begin_buffer();
            stage_packed_half(zp, 0);
            stage_packed_half(zp + 16, 1);
            end_buffer(&tmAux, gcol, false);
          }
        }
      }
      if (chain != nullptr && lane == 0) {
        if (w.part + 1 < p.split) {
          tma_wait_group<0>();            // this range's stores / reduce-adds have been performed ...
          asm volatile("fence.proxy.async.global;" ::: "memory");
          st_release_gpu(chain, w.part + 1);  // ... the next range of this region may go
        } else {
          *chain = 0;                     // last range: leave the counter ready for the next launch that uses this slot
        }
      }
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) tma_wait_group<0>();
  }

  tc_fence_before_sync();
  cluster_sync_all();   // nobody may exit (or free TMEM) while the peer still reads this CTA's smem / barriers
  if (warp == 10) {
    tc_fence_after_sync();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

}  // namespace g2
}  // namespace ztb

using namespace ztb;

int ztb_make_tmap_2d(CUtensorMap* tm, const void* ptr, int elem_bytes, long long rows, long long cols, long long ld,
                     int box_cols, int box_rows);

template <bool A_MN, bool B_MN, int EPI, int AG = 0>
static int launch2(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const CUtensorMap& x, const CUtensorMap& xin,
                   const g2::GemmParams& p, int grid, cudaStream_t st, const g2::AgParams* agp = nullptr) {
  auto kern = g2::gemm2_tcgen05_kernel<A_MN, B_MN, EPI, AG>;
  static bool configured = false;
  if (!configured) {
    ZTB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g2::GEMM_SMEM));
    configured = true;
  }
  g2::AgParams ag{};
  if (agp) ag = *agp;
  ZTB_CUDA_CHECK(ztb_launch(kern, dim3(grid), dim3(g2::GEMM_THREADS), g2::GEMM_SMEM, st, a, b, c, x, xin, p, ag));  // cluster dims are compile-time (2,1,1)
  return 0;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// K-split ordering counters: one slot of a small per-device ring per launch.  The kernel leaves every counter it
// used at zero again (the last K-range resets it), so the ring is zeroed once, at allocation.
static int* chain_slot(int* err) {
  constexpr int kSlotInts = 128 * 16, kSlots = 64, kMaxDev = 16;
  static std::mutex mu;
  static int* ring[kMaxDev] = {nullptr};
  static unsigned next_slot[kMaxDev] = {0};
  *err = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) { *err = -1; return nullptr; }
  std::lock_guard<std::mutex> lock(mu);
  if (ring[dev] == nullptr) {
    if (cudaMalloc(&ring[dev], sizeof(int) * kSlotInts * kSlots) != cudaSuccess ||
        cudaMemset(ring[dev], 0, sizeof(int) * kSlotInts * kSlots) != cudaSuccess) { ring[dev] = nullptr; *err = -1; return nullptr; }
  }
  return ring[dev] + (size_t)(next_slot[dev]++ % kSlots) * kSlotInts;
}

// Same contract as ztb_gemm (without the all-gather variant); M is processed in 256-row tiles.
// Returns -2 (no error message) for an (operand major, epilogue) combination this kernel is not instantiated for;
// the caller then uses the general 1-CTA kernel.
int ztb_gemm_2cta(const ZtbGemmArgs* g, int sms, cudaStream_t st) {
  using namespace g2;
  if (g->M <= 0 || g->N <= 0 || g->K <= 0) return 0;
  const int major = (g->a_mn_major ? 2 : 0) | (g->b_mn_major ? 1 : 0);
  const bool supported = g->epilogue == ZTB_EPI_NONE || (major == 0 && (g->epilogue == ZTB_EPI_GELU || g->epilogue == ZTB_EPI_DROPOUT_ADD)) ||
                         (major == 1 && (g->epilogue == ZTB_EPI_GELU_BWD || g->epilogue == ZTB_EPI_ROWDOT));
  if (!supported) return -2;
  if (g->ag != nullptr && (major != 0 || g->epilogue != ZTB_EPI_NONE || g->accumulate || g->c_is_fp32)) {
    ztb_set_error("ztb_gemm_2cta", "all-gather fusion is implemented for the plain K-major forward GEMM", __FILE__, __LINE__);
    return -1;
  }
  if ((g->lda % 8) || (g->ldb % 8) || (g->ldc % (g->c_is_fp32 ? 4 : 8)) || (g->K % 8) || (g->N % 8)) {
    ztb_set_error("ztb_gemm_2cta", "leading dimensions / K / N must be multiples of 8 elements", __FILE__, __LINE__);
    return -1;
  }
  CUtensorMap ta, tb, tc, tx;
  int rc;
  rc = g->a_mn_major ? ztb_make_tmap_2d(&ta, g->A, 2, g->K, g->M, g->lda, 64, 64)
                     : ztb_make_tmap_2d(&ta, g->A, 2, g->M, g->K, g->lda, BK, BM);
  if (rc) return rc;
  rc = g->b_mn_major ? ztb_make_tmap_2d(&tb, g->B, 2, g->K, g->N, g->ldb, 64, 64)
                     : ztb_make_tmap_2d(&tb, g->B, 2, g->N, g->K, g->ldb, BK, 128);
  if (rc) return rc;
  rc = g->c_is_fp32 ? ztb_make_tmap_2d(&tc, g->C, 4, g->M, g->N, g->ldc, 32, 32)
                    : ztb_make_tmap_2d(&tc, g->C, 2, g->M, g->N, g->ldc, 64, 32);
  if (rc) return rc;
  tx = tc;
  CUtensorMap tin = tc;   // auxiliary INPUT operand (residual / Z / O): 64-column x 32-row boxes, one per epilogue warp and chunk
  if (g->aux_in != nullptr && (g->epilogue == ZTB_EPI_DROPOUT_ADD || g->epilogue == ZTB_EPI_GELU_BWD || g->epilogue == ZTB_EPI_ROWDOT)) {
    if (g->ld_aux % 8) { ztb_set_error("ztb_gemm_2cta", "aux_in row stride must be a multiple of 8 elements", __FILE__, __LINE__); return -1; }
    rc = ztb_make_tmap_2d(&tin, g->aux_in, 2, g->M, g->N, g->ld_aux, 64, 32);
    if (rc) return rc;
  }
  GemmParams p;
  p.M = g->M; p.N = g->N; p.K = g->K;
  p.epi = g->epilogue; p.c_fp32 = g->c_is_fp32; p.accumulate = g->accumulate;
  p.aux_in = (const bf16*)g->aux_in; p.ld_aux = g->ld_aux;
  p.has_aux_out = (g->epilogue == ZTB_EPI_GELU && g->aux_out != nullptr) ? 1 : 0;
  if (p.has_aux_out) {
    if (g->c_is_fp32 || g->accumulate) { ztb_set_error("ztb_gemm_2cta", "GELU epilogue needs a bf16 non-accumulating C", __FILE__, __LINE__); return -1; }
    rc = ztb_make_tmap_2d(&tx, g->aux_out, 2, g->M, g->N, g->ld_aux, 64, 32);
    if (rc) return rc;
  }
  if ((g->epilogue == ZTB_EPI_DROPOUT_ADD || g->epilogue == ZTB_EPI_GELU_BWD || g->epilogue == ZTB_EPI_ROWDOT) && g->aux_in == nullptr) {
    ztb_set_error("ztb_gemm_2cta", "epilogue needs aux_in", __FILE__, __LINE__); return -1;
  }
  p.mask_out = g->epilogue == ZTB_EPI_DROPOUT_ADD ? g->mask_out : nullptr; p.mask_ld = g->N / 32;
  if (p.mask_out != nullptr && g->N % 32) { ztb_set_error("ztb_gemm_2cta", "mask_out needs N % 32 == 0", __FILE__, __LINE__); return -1; }
  p.rowdot_out = g->rowdot_out; p.rowdot_T = g->rowdot_T; p.rowdot_hd = g->rowdot_hd; p.rowdot_H = g->rowdot_hd > 0 ? g->N / g->rowdot_hd : 0;
  if (g->epilogue == ZTB_EPI_ROWDOT && (g->rowdot_out == nullptr || g->rowdot_T <= 0 || (g->rowdot_hd != 64 && g->rowdot_hd != 128) ||
                                         g->N % g->rowdot_hd || g->M % g->rowdot_T || g->c_is_fp32 || g->accumulate)) {
    ztb_set_error("ztb_gemm_2cta", "ROWDOT epilogue: needs rowdot_out, T dividing M, head_dim 64 or 128 dividing N, bf16 non-accumulating C", __FILE__, __LINE__);
    return -1;
  }
  static const int direct_default = env_int("ZTB_GEMM_DIRECT", 0);
  p.direct = g->direct_store >= 0 ? g->direct_store : direct_default;
  p.c_raw = (bf16*)g->C; p.ldc_raw = g->ldc; p.aux_raw = (bf16*)g->aux_out; p.ld_aux_raw = g->ld_aux;
  p.p = g->p; p.inv_keep = g->p < 1.f ? 1.f / (1.f - g->p) : 0.f;
  p.rng = g->rng; p.op_id = g->op_id;
  p.tiles_m = (g->M + TILE_M - 1) / TILE_M; p.tiles_n = (g->N + BN - 1) / BN; p.num_tiles = p.tiles_m * p.tiles_n;
  // Raster: groups of GROUP_M row blocks sweep all column blocks (the B operand streams once per group, the A group stays
  // hot) -- unless B is the larger operand (more column blocks than row blocks: the LM head, 197 vs 64): then groups of
  // column blocks sweep all row blocks, so the activation panel (67 MB, L2-resident) is what gets re-read and the
  // 206 MB embedding streams through exactly once.
  static const int raster_env = env_int("ZTB_GEMM_RASTER_N", -1);
  p.raster_n = g->raster_n >= 0 ? g->raster_n : (raster_env >= 0 ? raster_env : (p.tiles_n > p.tiles_m ? 1 : 0));
  int pairs = sms / 2;
  // Wave quantisation: num_tiles is rarely a multiple of the 74 CTA pairs (weight gradients: 64 / 192 / 256
  // tiles).  Cut each tile of the last, partial wave into `split` K-ranges so that it costs
  // ceil(tail * split / pairs) / split of a tile time instead of a whole one.
  const int kblocks = (g->K + BK - 1) / BK;
  p.full_tiles = p.num_tiles / pairs * pairs; p.tail = p.num_tiles - p.full_tiles; p.split = 1;
  p.nb_shift = 0;
  if (g->ag != nullptr) {
    // all-gather variant: every CTA takes part in the gather (even with fewer tiles than CTA pairs), the column blocks
    // this rank owns are computed first, whole tiles only
    const ZtbAgArgs* a = g->ag;
    AgParams ag{};
    for (int i = 0; i < 8; ++i) ag.peer[i] = i < a->world ? a->peer[i] : nullptr;
    ag.local = (bf16*)a->local; ag.segs = a->segs; ag.nseg = a->nseg; ag.n_b_segs = a->n_b_segs; ag.total_tiles = a->total_tiles;
    ag.shard_ready = a->shard_ready; ag.epoch = a->epoch; ag.epoch_ptr = a->epoch_ptr; ag.done = a->done; ag.b_tiles_per_rank = a->b_tiles_per_rank;
    ag.rows_per_rank = a->rows_per_rank; ag.world = a->world; ag.rank = a->rank;
    if (a->done_base < 0) ZTB_CUDA_CHECK(cudaMemsetAsync(a->done, 0, 8 * sizeof(unsigned), st));
    else ag.b_tiles_per_rank = (unsigned)a->done_base + a->b_tiles_per_rank;   // counters are cumulative: the value that means "landed" for THIS launch
    p.raster_n = 2;
    p.nb_shift = ((a->rank * a->rows_per_rank) / BN) % p.tiles_n;
    p.full_tiles = p.num_tiles; p.tail = 0; p.split = 1; p.chain = nullptr; p.direct = 0;
    static const int ag_variant = [] { const char* e = getenv("ZTB_AG_VARIANT"); return e ? atoi(e) : 1; }();   // measured over NVLink at n = 2: variant 1 0.310 ms, variant 2 0.337 ms (plain GEMM 0.260)
    if (ag_variant == 1) return launch2<false, false, ZTB_EPI_NONE, 1>(ta, tb, tc, tx, tin, p, pairs * 2, st, &ag);
    return launch2<false, false, ZTB_EPI_NONE, 2>(ta, tb, tc, tx, tin, p, pairs * 2, st, &ag);
  }
  if (p.tail > 0 && g->epilogue == ZTB_EPI_NONE && !g->no_split) {
    float best = 1.f;
    for (int sp = 2; sp <= 8 && kblocks / sp >= 16; ++sp) {
      float cost = (float)((p.tail * sp + pairs - 1) / pairs) / sp + 0.03f * (sp - 1);  // + per-unit epilogue/ramp overhead (measured: 8-way split of a 64-tile problem gains nothing)
      if (cost < best - 1e-6f) { best = cost; p.split = sp; }
    }
  }
  p.chain = nullptr;
  if (p.split > 1) {
    int err = 0;
    p.chain = chain_slot(&err);
    if (err) { ztb_set_error("ztb_gemm_2cta", "could not allocate the K-split ordering counters", __FILE__, __LINE__); return -1; }
  }
  const int units = p.full_tiles + p.tail * p.split;
  if (units < pairs) pairs = units;
  int grid = pairs * 2;
  switch (major * 8 + g->epilogue) {
    case 0 * 8 + ZTB_EPI_NONE: return launch2<false, false, ZTB_EPI_NONE>(ta, tb, tc, tx, tin, p, grid, st);
    case 0 * 8 + ZTB_EPI_GELU: return launch2<false, false, ZTB_EPI_GELU>(ta, tb, tc, tx, tin, p, grid, st);
    case 0 * 8 + ZTB_EPI_DROPOUT_ADD: return launch2<false, false, ZTB_EPI_DROPOUT_ADD>(ta, tb, tc, tx, tin, p, grid, st);
    case 1 * 8 + ZTB_EPI_NONE: return launch2<false, true, ZTB_EPI_NONE>(ta, tb, tc, tx, tin, p, grid, st);
    case 1 * 8 + ZTB_EPI_GELU_BWD: return launch2<false, true, ZTB_EPI_GELU_BWD>(ta, tb, tc, tx, tin, p, grid, st);
    case 1 * 8 + ZTB_EPI_ROWDOT: return launch2<false, true, ZTB_EPI_ROWDOT>(ta, tb, tc, tx, tin, p, grid, st);
    case 2 * 8 + ZTB_EPI_NONE: return launch2<true, false, ZTB_EPI_NONE>(ta, tb, tc, tx, tin, p, grid, st);
    case 3 * 8 + ZTB_EPI_NONE: return launch2<true, true, ZTB_EPI_NONE>(ta, tb, tc, tx, tin, p, grid, st);
  }
  return -2;
}
