// C launcher prototypes shared by the .cu translation units and bindings.cpp.
// Every launcher returns 0 on success (or a cudaError_t / -1, message in ztb_last_error()).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

void ztb_set_error(const char* expr, const char* msg, const char* file, int line);
const char* ztb_last_error();

// ---- elementwise.cu
int ztb_layernorm_fwd(const void* x, const void* w, void* y, float* mean, float* rstd, int M, int D, float eps, cudaStream_t st);
// Optional extras of LayerNorm backward.  With `ticket` set, `dw` is a persistent zeroed fp32 scratch vector: the last
// block folds it into `grad_out` (overwrite or accumulate) and re-zeroes scratch and ticket.  `dmask`: second output,
// dropout-masked copy of dx (mask of the dropout+residual op that produced the LayerNorm input).
struct ZtbLnBwdExtra {
  void* grad_out; int grad_fp32; int grad_accumulate; unsigned int* ticket;
  float* partials;   // optional [ztb_layernorm_bwd_grid(M, D, sms), D]: order-fixed (bitwise reproducible) weight-gradient reduction
  void* dmask; float p; const uint64_t* rng; uint32_t op_id;
  const unsigned char* maskbits;   // optional [M, D / 8]: keep bits recorded by the forward GEMM epilogue; replaces the Philox recomputation
};
int ztb_layernorm_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd, void* dx,
                      float* dw, int M, int D, int sms, cudaStream_t st, const void* dres = nullptr,
                      const ZtbLnBwdExtra* extra = nullptr);
int ztb_layernorm_bwd_grid(int M, int D, int sms);
int ztb_dropout_add(const void* y, const void* res, void* out, size_t n, float p, const uint64_t* rng, uint32_t op_id,
                    int sms, cudaStream_t st);
int ztb_embedding_fwd(const int* idx, const void* wte, void* out, int M, int D, int V, cudaStream_t st);
int ztb_embedding_bwd(const int* idx, const void* dout, void* dwte, int grad_is_fp32, int M, int D, int V, cudaStream_t st);
int ztb_embedding_bwd_sorted(const int* sorted_tok, const int* order, const void* dout, void* dwte, int grad_is_fp32, int M, int D, int V,
                             cudaStream_t st);
int ztb_gelu_bwd(const void* dg, const void* z, void* dz, size_t n, int sms, cudaStream_t st);
int ztb_gelu_fwd(const void* z, void* g, size_t n, int sms, cudaStream_t st);
int ztb_ce_fwd(const void* logits, const int* labels, float* lse, float* nll, int M, int V, size_t ld, cudaStream_t st);
int ztb_ce_bwd(void* logits, const int* labels, const float* lse, const float* gscale, float inv_count, int M, int V,
               size_t ld, cudaStream_t st);

// ---- optim.cu
struct ZtbSegment {      // one parameter region's chunk owned by this rank
  long long flat_off;    // element offset of the chunk in the flat param/grad buffers
  long long shard_off;   // element offset in the rank's fp32 master/m/v shard
  long long len;         // elements in the chunk
  long long tile_begin;  // exclusive prefix sum of tiles over segments
  int decay;             // weight decay applies
  int pad;
};
struct ZtbAdamArgs {
  float lr, beta1, beta2, eps, wd, clip, grad_scale;
  float inv_bc1, inv_bc2;   // 1 / (1 - beta^step): bias corrections, reciprocal (the kernel multiplies)
  float omb1, omb2;         // 1 - beta1, 1 - beta2
  const float* scale_mul;   // optional device scalar multiplied into grad_scale (global-norm clip coefficient)
  const float* dyn;         // optional device [lr, inv_bc1, inv_bc2]: overrides the by-value fields (CUDA-graph replays)
};
// grads[p] = flat grad buffer of peer p (p = 0 is local); n_peers == 1 -> plain fused AdamW.
int ztb_adamw_shard(float* master, float* m, float* v, void* flat_param, const void* const* grads, int n_peers,
                    int grad_is_fp32, const ZtbSegment* segs, int nseg, long long total_tiles, ZtbAdamArgs a,
                    const int* wait_flags, int wait_value, int sms, int max_ctas, const void* mc_grad, void* mc_param,
                    cudaStream_t st);

// ---- gemm_tcgen05.cu
enum ZtbEpilogue { ZTB_EPI_NONE = 0, ZTB_EPI_GELU = 1, ZTB_EPI_DROPOUT_ADD = 2, ZTB_EPI_GELU_BWD = 3, ZTB_EPI_ROWDOT = 4 };
struct ZtbAgArgs {              // all-gather fused into the GEMM (see AgParams in gemm_tcgen05.cu)
  const void* peer[8]; void* local;
  const ZtbSegment* segs; int nseg; int n_b_segs; long long total_tiles;
  const int* shard_ready; int epoch; const int* epoch_ptr;   // epoch_ptr != null: the awaited value is *epoch_ptr + epoch
  unsigned int* done; unsigned int b_tiles_per_rank; int rows_per_rank; int world, rank;
  long long done_base;   // >= 0: the per-owner counters are cumulative and stood at done_base before this launch (no memset); < 0: zero them first
};
struct ZtbGemmArgs {
  const void* A; const void* B; void* C;   // C[M,N] = A[M,K] * B[N,K]^T  (bf16, K contiguous in both)
  int M, N, K;
  long long lda, ldb, ldc;                  // row strides in elements
  int epilogue;
  void* aux_out;                            // GELU: pre-activation Z[M,N] (optional)
  const void* aux_in;                       // DROPOUT_ADD: residual[M,N]; GELU_BWD: Z[M,N]
  long long ld_aux;
  float p;                                  // dropout probability
  const uint64_t* rng; uint32_t op_id;
  int accumulate;                           // C += result (fp32 C when c_is_fp32)
  int c_is_fp32;
  int a_mn_major, b_mn_major;               // operand stored with M (resp. N) contiguous: A[K,M] / B[K,N]
  const ZtbAgArgs* ag;                      // non-null: gather B's remote row chunks inside the kernel
  int two_cta;                              // use the cta_group::2 (256x256 per CTA pair) kernel
  int no_split;                             // two_cta: disable K-splitting of the last partial wave
  int direct_store;                         // two_cta: -1 default (ZTB_GEMM_DIRECT), else bit 0 aux output / bit 1 C skip smem staging
  int raster_n;                             // two_cta: -1 auto, 0 row-block groups, 1 column-block groups (vocabulary-sized N)
  // ROWDOT epilogue (two_cta, dgrad major): besides C, emit rowdot_out[b, h, t] = sum over head h's columns of C[m, :] * aux_in[m, :]
  // for m = b * T + t -- the attention-backward `delta = rowsum(dO o O)` computed where dO is produced.
  float* rowdot_out; int rowdot_T; int rowdot_hd;
  // DROPOUT_ADD epilogue (two_cta): optionally record the keep decisions, 1 bit per output element, row-major, (N / 32) 32-bit
  // words per row (bit k of word w of row m <-> column 32 w + k) -- backward then needs no Philox recomputation.
  unsigned int* mask_out;
};
int ztb_gemm_2cta(const ZtbGemmArgs* args, int sms, cudaStream_t st);
int ztb_gemm(const ZtbGemmArgs* args, int sms, cudaStream_t st);

// ---- comm.cu (NVLink peer-memory kernels; all pointers are device addresses, peers mapped via symmetric memory)
int ztb_signal(void* const* pads, int n, int slot, int value, cudaStream_t st, const int* value_ptr = nullptr);
int ztb_wait_flags(const int* flags, int n, int value, cudaStream_t st, const int* value_ptr = nullptr);
int ztb_allgather_pull(void* dst, void* const* srcs, int world, const ZtbSegment* segs, int nseg, long long total_tiles,
                       const int* flags, int wait_value, int sms, cudaStream_t st);
int ztb_reduce_scatter(void* out, void* const* grads, int n_peers, const ZtbSegment* segs, int nseg, long long total_tiles,
                       const int* flags, int wait_value, int sms, cudaStream_t st);
int ztb_p2p_copy(void* dst, const void* src, size_t bytes, int sms, cudaStream_t st);

// ---- attention.cu (tcgen05 ALiBi causal flash attention)
int ztb_attn_fwd(const void* qkv, void* out, float* lse, const float* slopes, int B, int T, int H, int hd, float scale,
                 float dropout_p, const uint64_t* rng, uint32_t op_id, cudaStream_t st);
// delta_ready != 0: `delta` already holds rowsum(dO o O) (computed by the out-projection dgrad GEMM's ROWDOT epilogue)
int ztb_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* dq_acc, float* delta,
                 const float* slopes, int B, int T, int H, int hd, float scale, float dropout_p, const uint64_t* rng,
                 uint32_t op_id, int sms, cudaStream_t st, unsigned long long* dbg = nullptr, int delta_ready = 0);
