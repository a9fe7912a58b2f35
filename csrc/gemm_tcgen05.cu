// bf16 GEMM on the 5th-generation tensor cores:  C[M,N] (+)= A[M,K] * B[N,K]^T.
//
//   * persistent, warp-specialised: warps 0-3 = epilogue (TMEM -> registers -> swizzled smem -> TMA store /
//     TMA reduce-add), warp 4 = TMA producer, warp 5 = tcgen05.mma issuer (one thread), warp 6 = TMEM
//     allocator; in the all-gather-fused variant warps 6-7 also pull the remote weight chunks over NVLink.
//     (Control warps take the high warp ids: the SMSP arbiter favours them over the ALU-heavy epilogue.)
//   * 128x256 output tile per CTA, K-block 64, 4-stage smem ring fed by cp.async.bulk.tensor
//     (SWIZZLE_128B), fp32 accumulators in TMEM, double-buffered (2 x 256 columns) so the
//     epilogue of tile i overlaps the MMAs of tile i+1.
//   * operands may be K-major (row-major A[M,K], B[N,K]: forward y = x W^T) or MN-major
//     (A stored [K,M] / B stored [K,N]: dgrad dX = dY W and wgrad dW = dY^T X) -- selected by the
//     UMMA instruction-descriptor major bits and the smem descriptor strides, no transposes.
//   * fused epilogues: none | tanh-GELU (also emits the pre-activation)
//     | dropout+residual | GELU-backward | accumulate-into-C (wgrad beta=1) in bf16 or fp32.
//   Reference call sites: layers.py:60-77,112-143,180-189 (projections), GPT.py:100 (tied LM head), and the
//   implicit parameter all-gather of main_zero.py:452-456 (out_shardings=(None, ...)) for the AG variant.
#include <cuda.h>
#include <stdio.h>

#include "common.cuh"
#include "launch.cuh"
#include "launchers.h"

namespace ztb {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int CBUF_BYTES = BM * 128;  // 128 rows x 128 B (64 bf16 or 32 fp32 columns)
constexpr int GEMM_SMEM = STAGES * STAGE_BYTES + 2 * CBUF_BYTES + 256 /*barriers*/ + 1024 /*align slack*/;
constexpr int GEMM_THREADS = 256;
constexpr int GROUP_M = 16;

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
  int n_major_raster;   // AG variant: walk all M tiles of one N block before the next N block
  int nb_shift;         // AG variant: logical N block 0 maps to this rank's own rows
};

// All-gather fused into the GEMM (ZeRO-1 parameter all-gather + first consuming GEMM).
// B is a weight whose row chunks are owned by different ranks; warp 3 of every CTA pulls the remote
// chunks of the whole bucket from the owners' symmetric parameter buffers over NVLink into the local
// buffer while the tensor cores already work on the locally owned N blocks; the TMA producer acquires
// a per-owner completion counter before it loads a B tile that contains gathered rows.
struct AgParams {
  const void* peer[8];          // flat bf16 parameter buffer of every rank (peer-mapped), index = rank
  bf16* local;                  // this rank's flat parameter buffer
  const ZtbSegment* segs;       // pull list: flat_off, len, decay = source rank, tile_begin; B's segments first
  int nseg;
  int n_b_segs;                 // the first n_b_segs segments are B's chunks (one per remote rank)
  long long total_tiles;
  const int* shard_ready;       // my flag row: [rank] >= epoch once that rank's chunk is final
  int epoch;
  const int* epoch_ptr;         // non-null: wait for *epoch_ptr + epoch (device-side step counter, CUDA-graph replays)
  unsigned int* done;           // [8] per-owner count of finished B tiles (zeroed before launch)
  unsigned int b_tiles_per_rank;
  int rows_per_rank;            // B rows owned by each rank
  int world, rank;
};

constexpr int AG_TILE = 8192;   // elements per gather tile (16 KB) -- matches COPY_TILE

ZTB_DEVINL unsigned ld_acquire_gpu_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
ZTB_DEVINL int ld_acquire_sys_s32(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

ZTB_DEVINL void decode_tile(int t, const GemmParams& p, int& mb, int& nb) {
  if (p.n_major_raster) {
    int nl = t / p.tiles_m;
    mb = t % p.tiles_m;
    nb = nl + p.nb_shift;
    if (nb >= p.tiles_n) nb -= p.tiles_n;
    return;
  }
  int per_group = GROUP_M * p.tiles_n;
  int g = t / per_group, r = t % per_group;
  int gm = min(GROUP_M, p.tiles_m - g * GROUP_M);
  mb = g * GROUP_M + r % gm;
  nb = r / gm;
}

template <bool A_MN, bool B_MN, bool AG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAux,
                    const GemmParams p, const __grid_constant__ AgParams ag) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + STAGES * A_STAGE_BYTES;
  uint8_t* smC = smem + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smC + 2 * CBUF_BYTES);
  uint64_t* full = bars;              // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + STAGES;    // [STAGES]  MMA -> TMA
  uint64_t* tfull = bars + 2 * STAGES;      // [2] MMA -> epilogue
  uint64_t* tempty = bars + 2 * STAGES + 2;  // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = (p.K + BK - 1) / BK;

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB); prefetch_tmap(&tmC);
    if (p.has_aux_out) prefetch_tmap(&tmAux);
  }
  if (warp == 5 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 6) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_enter();

  // Control warps take the highest warp ids (the SMSP arbiter favours them over the epilogue warps).
  if (warp == 4) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        int mb, nb; decode_tile(t, p, mb, nb);
        if constexpr (AG) {
          // rows [nb*BN, nb*BN+BN) of B: every owner other than this rank must have been gathered
          const int r0 = (nb * BN) / ag.rows_per_rank;
          int r1 = (nb * BN + BN - 1) / ag.rows_per_rank;
          if (r1 >= ag.world) r1 = ag.world - 1;
          for (int o = r0; o <= r1; ++o)
            if (o != ag.rank)
              while ((int)(ld_acquire_gpu_u32(ag.done + o) - ag.b_tiles_per_rank) < 0) __nanosleep(32);
          fence_proxy_async_all();  // generic-proxy stores of the gather warps -> async-proxy TMA reads
        }
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          uint8_t* a = smA + stage * A_STAGE_BYTES;
          uint8_t* b = smB + stage * B_STAGE_BYTES;
          if constexpr (!A_MN) {
            tma_load_2d(a, &tmA, &full[stage], kb * BK, mb * BM);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(a + j * 8192, &tmA, &full[stage], mb * BM + j * 64, kb * BK);
          }
          if constexpr (!B_MN) {
            tma_load_2d(b, &tmB, &full[stage], kb * BK, nb * BN);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(b + j * 8192, &tmB, &full[stage], nb * BN + j * 64, kb * BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    // ===================================================== MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
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
            umma_f16_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);                    // smem slot free once these MMAs retire
          if (kb == kblocks - 1) umma_commit(&tfull[acc]);  // accumulator complete
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        acc ^= 1; if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (AG && warp >= 6) {
    // ===================================================== all-gather warp (NVLink peer pulls)
    if constexpr (AG) {
      unsigned seen = 1u << ag.rank;
      const int want_epoch = ag.epoch + (ag.epoch_ptr != nullptr ? *ag.epoch_ptr : 0);
      // two gather warps per CTA (warp 6 is free once it has allocated TMEM)
      for (long long t = blockIdx.x * 2 + (warp - 6); t < ag.total_tiles; t += gridDim.x * 2) {
        int lo = 0, hi = ag.nseg - 1;
        while (lo < hi) {
          int mid = (lo + hi + 1) >> 1;
          if (ag.segs[mid].tile_begin <= t) lo = mid; else hi = mid - 1;
        }
        const ZtbSegment sg = ag.segs[lo];
        const int src = sg.decay;
        if (!((seen >> src) & 1u)) {   // the owner has published this step's chunk?
          if (lane == 0) {
            unsigned long long t0;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            while (ld_acquire_sys_s32(ag.shard_ready + src) < want_epoch) {
              __nanosleep(64);
              unsigned long long t1;
              asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
              if (t1 - t0 > 120ull * 1000000000ull) __trap();  // owner never published its shard: fail loudly
            }
          }
          __syncwarp();
          seen |= 1u << src;
        }
        const bf16* sp = reinterpret_cast<const bf16*>(ag.peer[src]) + sg.flat_off;
        bf16* dp = ag.local + sg.flat_off;
        const long long base = (t - sg.tile_begin) * AG_TILE;
#pragma unroll 1
        for (int it = 0; it < 4; ++it) {       // 4 x (8 x 32 lanes x 16 B) = 16 KB
          uint4 v[8];
          long long e[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            e[k] = base + ((long long)(it * 8 + k) * 32 + lane) * 8;
            if (e[k] < sg.len) v[k] = ld_nc_v4(sp + e[k]);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (e[k] < sg.len) st_v4(dp + e[k], v[k]);
        }
        if (lo < ag.n_b_segs) {                // a tile of B: publish it to the TMA producers
          __threadfence();
          __syncwarp();
          if (lane == 0) atomicAdd(ag.done + src, 1u);
        }
      }
    }
  } else if (warp < 4) {
    // ===================================================== epilogue (4 warps, 128 rows)
    const int ew = warp;                    // == warp % 4 -> TMEM lane quarter
    const int row = ew * 32 + lane;         // row inside the tile
    const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
    const uint32_t thr = dropout_threshold(p.p);
    uint64_t seed = 0, stream = 0;
    if (p.epi == ZTB_EPI_DROPOUT_ADD) { seed = p.rng[0]; stream = p.rng[1] * 65536ull + p.op_id; }
    Philox ph(seed);
    int acc = 0; uint32_t acc_phase = 0;
    int cbuf = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      int mb, nb; decode_tile(t, p, mb, nb);
      const int m0 = mb * BM, n0 = nb * BN;
      const int grow = m0 + row;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t t_acc = tmem_base + lane_addr + acc * BN;
      // One staging buffer holds 128 B per row: 64 bf16 columns or 32 fp32 columns.  Buffers
      // alternate; before overwriting one, the issuing thread waits until at most one bulk store
      // (the other buffer's) is still reading smem.
      // Each epilogue warp owns its 32 rows of the staging buffer and issues its own bulk stores
      // ({64|32} x 32-row boxes), so the only synchronisation is warp-local.
      auto begin_buffer = [&]() {
        if (lane == 0) tma_wait_group_read<1>();
        __syncwarp();
      };
      auto end_buffer = [&](const CUtensorMap* tm, int col) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const uint8_t* src = smC + cbuf * CBUF_BYTES + ew * 4096;
          if (p.accumulate) tma_reduce_add_2d(tm, src, col, m0 + ew * 32);
          else tma_store_2d(tm, src, col, m0 + ew * 32);
          tma_commit_group();
        }
        cbuf ^= 1;
      };
      auto stage_bf16 = [&](const float* v) {  // 64 values -> this thread's swizzled 128-byte row
        uint8_t* rowp = smC + cbuf * CBUF_BYTES + row * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 o = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                               pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
          sts_v4(rowp + ((j ^ (row & 7)) << 4), o);
        }
      };
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c) {
        uint32_t r[64];
        tmem_ld_32x32b_x32(t_acc + c * 64, r);
        tmem_ld_32x32b_x32(t_acc + c * 64 + 32, r + 32);
        tmem_wait_ld();
        if (c == BN / 64 - 1) {  // last TMEM read of this accumulator: hand it back to the MMA warp
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        const int gcol = n0 + c * 64;
        float v[64];
#pragma unroll
        for (int k = 0; k < 64; ++k) v[k] = __uint_as_float(r[k]);
        if (p.epi == ZTB_EPI_GELU) {
          if (p.has_aux_out) {  // pre-activation Z leaves through its own tensor map
            begin_buffer();
            stage_bf16(v);
            end_buffer(&tmAux, gcol);
          }
#pragma unroll
          for (int k = 0; k < 64; ++k) v[k] = gelu_tanh(v[k]);
        } else if (p.epi == ZTB_EPI_DROPOUT_ADD || p.epi == ZTB_EPI_GELU_BWD) {
          const bool in_rows = grow < p.M;
          const bf16* ap = p.aux_in + (size_t)grow * p.ld_aux + gcol;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            uint4 u = make_uint4(0, 0, 0, 0);
            if (in_rows && gcol + q * 8 < p.N) u = ld_nc_l1_v4(ap + q * 8);
            uint32_t us[4] = {u.x, u.y, u.z, u.w};
            if (p.epi == ZTB_EPI_DROPOUT_ADD) {
              size_t idx = (size_t)grow * p.N + gcol + q * 8;
              uint4 r0 = ph.rounds<kPhiloxRounds>(idx >> 2, stream), r1 = ph.rounds<kPhiloxRounds>((idx >> 2) + 1, stream);
              uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float2 rf = unpack_bf16(us[k]);
                float a = rr[2 * k] >= thr ? v[q * 8 + 2 * k] * p.inv_keep : 0.f;
                float b = rr[2 * k + 1] >= thr ? v[q * 8 + 2 * k + 1] * p.inv_keep : 0.f;
                v[q * 8 + 2 * k] = rf.x + a;
                v[q * 8 + 2 * k + 1] = rf.y + b;
              }
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float2 zf = unpack_bf16(us[k]);
                v[q * 8 + 2 * k] *= gelu_tanh_grad(zf.x);
                v[q * 8 + 2 * k + 1] *= gelu_tanh_grad(zf.y);
              }
            }
          }
        }
        if (p.c_fp32) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            begin_buffer();
            uint8_t* rowp = smC + cbuf * CBUF_BYTES + row * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float* s4 = v + h * 32 + 4 * j;
              uint4 o = make_uint4(__float_as_uint(s4[0]), __float_as_uint(s4[1]), __float_as_uint(s4[2]), __float_as_uint(s4[3]));
              sts_v4(rowp + ((j ^ (row & 7)) << 4), o);
            }
            end_buffer(&tmC, gcol + h * 32);
          }
        } else {
          begin_buffer();
          stage_bf16(v);
          end_buffer(&tmC, gcol);
        }
      }
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) tma_wait_group<0>();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 6) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace ztb

using namespace ztb;

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2-D row-major tensor [rows, cols] (cols contiguous, row stride ld elements) -> tiled map with a
// {box_cols, box_rows} box and 128-byte swizzle.
int ztb_make_tmap_2d(CUtensorMap* tm, const void* ptr, int elem_bytes, long long rows, long long cols, long long ld,
                     int box_cols, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { ztb_set_error("cuTensorMapEncodeTiled", "driver entry point not found", __FILE__, __LINE__); return -1; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(tm, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    static char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (%d): ptr=%p rows=%lld cols=%lld ld=%lld box=%dx%d", (int)r, ptr, rows, cols, ld, box_cols, box_rows);
    ztb_set_error("cuTensorMapEncodeTiled", buf, __FILE__, __LINE__);
    return -1;
  }
  return 0;
}

template <bool A_MN, bool B_MN, bool AG = false>
static int launch(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const CUtensorMap& x, const GemmParams& p,
                  int grid, cudaStream_t st, const AgParams* agp = nullptr) {
  auto kern = gemm_tcgen05_kernel<A_MN, B_MN, AG>;
  AgParams ag{};
  if (agp) ag = *agp;
  static bool configured = false;
  if (!configured) {
    ZTB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    configured = true;
  }
  ZTB_CUDA_CHECK(ztb_launch(kern, dim3(grid), dim3(GEMM_THREADS), GEMM_SMEM, st, a, b, c, x, p, ag));
  return 0;
}

int ztb_gemm(const ZtbGemmArgs* g, int sms, cudaStream_t st) {
  if (g->M <= 0 || g->N <= 0 || g->K <= 0) return 0;
  if (g->two_cta) {
    const int rc2 = ztb_gemm_2cta(g, sms, st);
    if (rc2 != -2) return rc2;   // -2: that (operand major, epilogue) pair only exists in this general kernel
  }
  if (g->epilogue == ZTB_EPI_ROWDOT) {
    ztb_set_error("ztb_gemm", "the ROWDOT epilogue exists only in the CTA-pair kernel (two_cta, dgrad operand major)", __FILE__, __LINE__);
    return -1;
  }
  if ((g->lda % 8) || (g->ldb % 8) || (g->ldc % (g->c_is_fp32 ? 4 : 8)) || (g->K % 8) || (g->N % 8)) {
    ztb_set_error("ztb_gemm", "leading dimensions / K / N must be multiples of 8 elements (16-byte TMA strides)", __FILE__, __LINE__);
    return -1;
  }
  CUtensorMap ta, tb, tc, tx;
  int rc;
  // K-major A: [M,K] box {64 k, 128 m}.  MN-major A: stored [K,M], box {64 m, 64 k}.
  rc = g->a_mn_major ? ztb_make_tmap_2d(&ta, g->A, 2, g->K, g->M, g->lda, 64, 64)
                     : ztb_make_tmap_2d(&ta, g->A, 2, g->M, g->K, g->lda, BK, BM);
  if (rc) return rc;
  rc = g->b_mn_major ? ztb_make_tmap_2d(&tb, g->B, 2, g->K, g->N, g->ldb, 64, 64)
                     : ztb_make_tmap_2d(&tb, g->B, 2, g->N, g->K, g->ldb, BK, BN);
  if (rc) return rc;
  rc = g->c_is_fp32 ? ztb_make_tmap_2d(&tc, g->C, 4, g->M, g->N, g->ldc, 32, 32)
                    : ztb_make_tmap_2d(&tc, g->C, 2, g->M, g->N, g->ldc, 64, 32);
  if (rc) return rc;
  tx = tc;
  GemmParams p;
  p.M = g->M; p.N = g->N; p.K = g->K;
  p.epi = g->epilogue; p.c_fp32 = g->c_is_fp32; p.accumulate = g->accumulate;
  p.aux_in = (const bf16*)g->aux_in; p.ld_aux = g->ld_aux;
  p.has_aux_out = (g->epilogue == ZTB_EPI_GELU && g->aux_out != nullptr) ? 1 : 0;
  if (p.has_aux_out) {
    if (g->c_is_fp32 || g->accumulate) { ztb_set_error("ztb_gemm", "GELU epilogue needs a bf16 non-accumulating C", __FILE__, __LINE__); return -1; }
    rc = ztb_make_tmap_2d(&tx, g->aux_out, 2, g->M, g->N, g->ld_aux, 64, 32);
    if (rc) return rc;
  }
  if ((g->epilogue == ZTB_EPI_DROPOUT_ADD || g->epilogue == ZTB_EPI_GELU_BWD) && g->aux_in == nullptr) {
    ztb_set_error("ztb_gemm", "epilogue needs aux_in", __FILE__, __LINE__); return -1;
  }
  p.p = g->p; p.inv_keep = g->p < 1.f ? 1.f / (1.f - g->p) : 0.f;
  p.rng = g->rng; p.op_id = g->op_id;
  p.tiles_m = (g->M + BM - 1) / BM; p.tiles_n = (g->N + BN - 1) / BN; p.num_tiles = p.tiles_m * p.tiles_n;
  p.n_major_raster = 0; p.nb_shift = 0;
  int grid = p.num_tiles < sms ? p.num_tiles : sms;
  if (g->ag) {
    const ZtbAgArgs* a = g->ag;
    if (g->a_mn_major || g->b_mn_major || g->epilogue != ZTB_EPI_NONE || g->accumulate) {
      ztb_set_error("ztb_gemm", "all-gather fusion is implemented for the plain K-major forward GEMM", __FILE__, __LINE__);
      return -1;
    }
    AgParams ag{};
    for (int i = 0; i < 8; ++i) ag.peer[i] = i < a->world ? a->peer[i] : nullptr;
    ag.local = (bf16*)a->local; ag.segs = a->segs; ag.nseg = a->nseg; ag.n_b_segs = a->n_b_segs; ag.total_tiles = a->total_tiles;
    ag.shard_ready = a->shard_ready; ag.epoch = a->epoch; ag.epoch_ptr = a->epoch_ptr; ag.done = a->done; ag.b_tiles_per_rank = a->b_tiles_per_rank;
    ag.rows_per_rank = a->rows_per_rank; ag.world = a->world; ag.rank = a->rank;
    if (a->done_base < 0) ZTB_CUDA_CHECK(cudaMemsetAsync(a->done, 0, 8 * sizeof(unsigned), st));
    else ag.b_tiles_per_rank = (unsigned)a->done_base + a->b_tiles_per_rank;   // counters are cumulative: the value that means "landed" for THIS launch
    p.n_major_raster = 1;
    p.nb_shift = ((a->rank * a->rows_per_rank) / BN) % p.tiles_n;
    grid = sms;  // every SM contributes a gather warp even if there are fewer tiles than SMs
    return launch<false, false, true>(ta, tb, tc, tx, p, grid, st, &ag);
  }
  if (g->a_mn_major && g->b_mn_major) return launch<true, true>(ta, tb, tc, tx, p, grid, st);
  if (g->a_mn_major) return launch<true, false>(ta, tb, tc, tx, p, grid, st);
  if (g->b_mn_major) return launch<false, true>(ta, tb, tc, tx, p, grid, st);
  return launch<false, false>(ta, tb, tc, tx, p, grid, st);
}
