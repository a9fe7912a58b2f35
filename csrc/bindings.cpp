// Python bindings (pybind11 through torch/extension.h).  Thin: argument checks, output
// allocation, current-stream lookup; all device code lives in the .cu files.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <vector>

#include "launchers.h"

namespace {

void check(int rc, const char* what) {
  if (rc != 0) throw std::runtime_error(std::string(what) + " failed: " + ztb_last_error());
}
cudaStream_t stream() { return at::cuda::getCurrentCUDAStream().stream(); }
int num_sms() {
  static int n = 0;
  if (!n) n = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  return n;
}
void need(bool c, const char* msg) {
  if (!c) throw std::invalid_argument(msg);
}
bool is_bf16_cuda(const at::Tensor& t) { return t.is_cuda() && t.scalar_type() == at::kBFloat16; }

std::vector<at::Tensor> layernorm_fwd(const at::Tensor& x, const at::Tensor& w, double eps) {
  need(is_bf16_cuda(x) && is_bf16_cuda(w) && x.is_contiguous() && w.is_contiguous(), "layernorm_fwd: bf16 contiguous CUDA tensors");
  c10::cuda::CUDAGuard g(x.device());
  int D = x.size(-1);
  int M = x.numel() / D;
  need(D % 8 == 0 && w.numel() == D, "layernorm_fwd: bad D");
  auto y = at::empty_like(x);
  auto mean = at::empty({M}, x.options().dtype(at::kFloat));
  auto rstd = at::empty({M}, x.options().dtype(at::kFloat));
  check(ztb_layernorm_fwd(x.data_ptr(), w.data_ptr(), y.data_ptr(), mean.data_ptr<float>(), rstd.data_ptr<float>(), M, D, (float)eps, stream()), "layernorm_fwd");
  return {y, mean, rstd};
}

// Returns {dx, dw, dmask}.  With `grad_out` (the weight's gradient buffer, bf16 or fp32) the kernel's last block writes /
// accumulates the weight gradient there and dw is an empty tensor; `scratch` (fp32 [D]) and `ticket` (int32 [1]) are
// persistent, zero on entry and zero again on exit.  drop_p > 0: also return dmask = dropout_mask(dx) / (1 - p).
std::vector<at::Tensor> layernorm_bwd(const at::Tensor& dy, const at::Tensor& x, const at::Tensor& w, const at::Tensor& mean,
                                      const at::Tensor& rstd, const c10::optional<at::Tensor>& dres,
                                      const c10::optional<at::Tensor>& grad_out, bool accumulate,
                                      const c10::optional<at::Tensor>& scratch, const c10::optional<at::Tensor>& ticket,
                                      double drop_p, const c10::optional<at::Tensor>& rng, int64_t op_id, bool deterministic,
                                      const c10::optional<at::Tensor>& drop_mask) {
  need(is_bf16_cuda(dy) && is_bf16_cuda(x) && dy.is_contiguous() && x.is_contiguous(), "layernorm_bwd: bf16 contiguous");
  c10::cuda::CUDAGuard g(x.device());
  int D = x.size(-1);
  int M = x.numel() / D;
  auto dx = at::empty_like(x);
  const void* dr = nullptr;
  if (dres.has_value()) { need(is_bf16_cuda(*dres) && dres->is_contiguous() && dres->numel() == x.numel(), "layernorm_bwd: bad dres"); dr = dres->data_ptr(); }
  ZtbLnBwdExtra ex{};
  at::Tensor dw, dmask, partials;
  float* dwp = nullptr;
  if (grad_out.has_value()) {
    need(scratch.has_value() && ticket.has_value(), "layernorm_bwd: grad_out needs scratch + ticket");
    need(grad_out->is_cuda() && grad_out->is_contiguous() && grad_out->numel() == D &&
         (grad_out->scalar_type() == at::kBFloat16 || grad_out->scalar_type() == at::kFloat), "layernorm_bwd: grad_out must be a contiguous bf16/fp32 [D] CUDA tensor");
    need(scratch->is_cuda() && scratch->scalar_type() == at::kFloat && scratch->numel() >= D && scratch->is_contiguous(), "layernorm_bwd: scratch fp32 [D]");
    need(ticket->is_cuda() && ticket->element_size() == 4 && ticket->numel() >= 1, "layernorm_bwd: ticket int32 [1]");
    ex.grad_out = grad_out->data_ptr(); ex.grad_fp32 = grad_out->scalar_type() == at::kFloat; ex.grad_accumulate = accumulate;
    ex.ticket = reinterpret_cast<unsigned int*>(ticket->data_ptr());
    dwp = scratch->data_ptr<float>();
    dw = at::empty({0}, x.options().dtype(at::kFloat));
    if (deterministic) {
      partials = at::empty({(int64_t)ztb_layernorm_bwd_grid(M, D, num_sms()), (int64_t)D}, x.options().dtype(at::kFloat));
      ex.partials = partials.data_ptr<float>();
    }
  } else {
    dw = at::empty({D}, x.options().dtype(at::kFloat));
    dwp = dw.data_ptr<float>();
  }
  if (drop_p > 0.0) {
    dmask = at::empty_like(x);
    ex.dmask = dmask.data_ptr(); ex.p = (float)drop_p; ex.op_id = (uint32_t)op_id;
    if (drop_mask.has_value()) {
      need(drop_mask->is_cuda() && drop_mask->is_contiguous() && drop_mask->element_size() == 4 && D % 32 == 0 &&
           drop_mask->numel() == (int64_t)M * (D / 32), "layernorm_bwd: drop_mask must be the contiguous 32-bit [M, D/32] keep-bit tensor of the forward epilogue");
      ex.maskbits = reinterpret_cast<const unsigned char*>(drop_mask->data_ptr());
    } else {
      need(rng.has_value() && rng->is_cuda() && rng->scalar_type() == at::kLong && rng->numel() >= 2, "layernorm_bwd: dropout output needs rng (CUDA int64[2]) or drop_mask");
      ex.rng = (const uint64_t*)rng->data_ptr<int64_t>();
    }
  } else {
    dmask = at::empty({0}, x.options());
  }
  check(ztb_layernorm_bwd(dy.data_ptr(), x.data_ptr(), w.data_ptr(), mean.data_ptr<float>(), rstd.data_ptr<float>(), dx.data_ptr(),
                          dwp, M, D, num_sms(), stream(), dr, &ex), "layernorm_bwd");
  return {dx, dw, dmask};
}

at::Tensor dropout_add(const at::Tensor& y, const c10::optional<at::Tensor>& res, double p, const at::Tensor& rng, int64_t op_id) {
  need(is_bf16_cuda(y) && y.is_contiguous(), "dropout_add: y bf16 contiguous");
  need(rng.is_cuda() && rng.scalar_type() == at::kLong && rng.numel() >= 2, "dropout_add: rng must be a CUDA int64[2] tensor");
  c10::cuda::CUDAGuard g(y.device());
  auto out = at::empty_like(y);
  const void* r = nullptr;
  if (res.has_value()) { need(is_bf16_cuda(*res) && res->is_contiguous() && res->numel() == y.numel(), "dropout_add: bad residual"); r = res->data_ptr(); }
  check(ztb_dropout_add(y.data_ptr(), r, out.data_ptr(), (size_t)y.numel(), (float)p, (const uint64_t*)rng.data_ptr<int64_t>(), (uint32_t)op_id, num_sms(), stream()), "dropout_add");
  return out;
}

at::Tensor embedding_fwd(const at::Tensor& idx, const at::Tensor& wte) {
  need(idx.is_cuda() && idx.scalar_type() == at::kInt && idx.is_contiguous(), "embedding_fwd: idx must be int32 CUDA contiguous");
  need(is_bf16_cuda(wte) && wte.is_contiguous() && wte.size(1) % 8 == 0, "embedding_fwd: wte bf16 contiguous, D % 8 == 0");
  c10::cuda::CUDAGuard g(wte.device());
  auto sizes = idx.sizes().vec();
  sizes.push_back(wte.size(1));
  auto out = at::empty(sizes, wte.options());
  check(ztb_embedding_fwd(idx.data_ptr<int>(), wte.data_ptr(), out.data_ptr(), (int)idx.numel(), (int)wte.size(1), (int)wte.size(0), stream()), "embedding_fwd");
  return out;
}

void embedding_bwd_(const at::Tensor& idx, const at::Tensor& dout, at::Tensor& dwte) {
  need(idx.is_cuda() && idx.scalar_type() == at::kInt && idx.is_contiguous(), "embedding_bwd: idx int32");
  need(is_bf16_cuda(dout) && dout.is_contiguous() && dwte.is_cuda() && dwte.is_contiguous(), "embedding_bwd: tensors");
  need(dwte.scalar_type() == at::kBFloat16 || dwte.scalar_type() == at::kFloat, "embedding_bwd: grad dtype");
  c10::cuda::CUDAGuard g(dout.device());
  check(ztb_embedding_bwd(idx.data_ptr<int>(), dout.data_ptr(), dwte.data_ptr(), dwte.scalar_type() == at::kFloat, (int)idx.numel(),
                          (int)dwte.size(1), (int)dwte.size(0), stream()), "embedding_bwd");
}

// sorted_tok / order: the token ids in ascending order and the positions they came from (int32, from torch.sort)
void embedding_bwd_sorted_(const at::Tensor& sorted_tok, const at::Tensor& order, const at::Tensor& dout, at::Tensor& dwte) {
  need(sorted_tok.is_cuda() && sorted_tok.scalar_type() == at::kInt && sorted_tok.is_contiguous() && order.is_cuda() &&
       order.scalar_type() == at::kInt && order.is_contiguous() && order.numel() == sorted_tok.numel(), "embedding_bwd_sorted: int32 index tensors");
  need(is_bf16_cuda(dout) && dout.is_contiguous() && dwte.is_cuda() && dwte.is_contiguous() && dwte.size(1) % 8 == 0, "embedding_bwd_sorted: tensors");
  need(dwte.scalar_type() == at::kBFloat16 || dwte.scalar_type() == at::kFloat, "embedding_bwd_sorted: grad dtype");
  c10::cuda::CUDAGuard g(dout.device());
  check(ztb_embedding_bwd_sorted(sorted_tok.data_ptr<int>(), order.data_ptr<int>(), dout.data_ptr(), dwte.data_ptr(), dwte.scalar_type() == at::kFloat,
                                 (int)sorted_tok.numel(), (int)dwte.size(1), (int)dwte.size(0), stream()), "embedding_bwd_sorted");
}

at::Tensor gelu_bwd(const at::Tensor& dg, const at::Tensor& z) {
  need(is_bf16_cuda(dg) && is_bf16_cuda(z) && dg.is_contiguous() && z.is_contiguous() && dg.numel() % 8 == 0, "gelu_bwd: bf16 contiguous, numel % 8 == 0");
  c10::cuda::CUDAGuard g(z.device());
  auto dz = at::empty_like(z);
  check(ztb_gelu_bwd(dg.data_ptr(), z.data_ptr(), dz.data_ptr(), (size_t)z.numel(), num_sms(), stream()), "gelu_bwd");
  return dz;
}
at::Tensor gelu_fwd(const at::Tensor& z) {
  need(is_bf16_cuda(z) && z.is_contiguous() && z.numel() % 8 == 0, "gelu_fwd: bf16 contiguous, numel % 8 == 0");
  c10::cuda::CUDAGuard g(z.device());
  auto o = at::empty_like(z);
  check(ztb_gelu_fwd(z.data_ptr(), o.data_ptr(), (size_t)z.numel(), num_sms(), stream()), "gelu_fwd");
  return o;
}

std::vector<at::Tensor> ce_fwd(const at::Tensor& logits, const at::Tensor& labels) {
  need(is_bf16_cuda(logits) && logits.dim() == 2 && logits.stride(1) == 1, "ce_fwd: logits must be 2-D bf16 with contiguous rows");
  need(labels.is_cuda() && labels.scalar_type() == at::kInt && labels.numel() == logits.size(0), "ce_fwd: labels int32 [M]");
  need(logits.stride(0) % 8 == 0, "ce_fwd: row stride must be a multiple of 8");
  c10::cuda::CUDAGuard g(logits.device());
  int M = logits.size(0), V = logits.size(1);
  auto lse = at::empty({M}, logits.options().dtype(at::kFloat));
  auto nll = at::empty({M}, logits.options().dtype(at::kFloat));
  check(ztb_ce_fwd(logits.data_ptr(), labels.data_ptr<int>(), lse.data_ptr<float>(), nll.data_ptr<float>(), M, V, (size_t)logits.stride(0), stream()), "ce_fwd");
  return {lse, nll};
}
void ce_bwd_(at::Tensor& logits, const at::Tensor& labels, const at::Tensor& lse, const at::Tensor& gscale, double inv_count) {
  need(is_bf16_cuda(logits) && logits.dim() == 2 && logits.stride(1) == 1, "ce_bwd: logits");
  need(gscale.is_cuda() && gscale.scalar_type() == at::kFloat && gscale.numel() == 1, "ce_bwd: gscale fp32 scalar tensor");
  c10::cuda::CUDAGuard g(logits.device());
  check(ztb_ce_bwd(logits.data_ptr(), labels.data_ptr<int>(), lse.data_ptr<float>(), gscale.data_ptr<float>(), (float)inv_count,
                   (int)logits.size(0), (int)logits.size(1), (size_t)logits.stride(0), stream()), "ce_bwd");
}

// C[M,N] (+)= op(A) * op(B)^T ; see launchers.h ZtbGemmArgs.  a_mn: A given as [K,M]; b_mn: B given as [K,N].
at::Tensor gemm(const at::Tensor& A, const at::Tensor& B, c10::optional<at::Tensor> C_, int64_t epilogue,
                const c10::optional<at::Tensor>& aux_in, c10::optional<at::Tensor> aux_out, double p,
                const c10::optional<at::Tensor>& rng, int64_t op_id, bool accumulate, bool a_mn, bool b_mn, bool two_cta, bool no_split,
                int64_t direct_store, int64_t raster_n, c10::optional<at::Tensor> rowdot_out, int64_t rowdot_T, int64_t rowdot_hd,
                c10::optional<at::Tensor> mask_out) {
  need(is_bf16_cuda(A) && is_bf16_cuda(B) && A.dim() == 2 && B.dim() == 2 && A.stride(1) == 1 && B.stride(1) == 1, "gemm: A, B must be 2-D bf16 CUDA with unit inner stride");
  c10::cuda::CUDAGuard g(A.device());
  ZtbGemmArgs a{};
  a.M = a_mn ? A.size(1) : A.size(0);
  a.K = a_mn ? A.size(0) : A.size(1);
  a.N = b_mn ? B.size(1) : B.size(0);
  need((b_mn ? B.size(0) : B.size(1)) == a.K, "gemm: K mismatch");
  at::Tensor C;
  if (C_.has_value()) {
    C = *C_;
    need(C.is_cuda() && C.dim() == 2 && C.size(0) == a.M && C.size(1) == a.N && C.stride(1) == 1, "gemm: bad C");
    need(C.scalar_type() == at::kBFloat16 || C.scalar_type() == at::kFloat, "gemm: C dtype");
  } else {
    need(!accumulate, "gemm: accumulate needs an explicit C");
    C = at::empty({a.M, a.N}, A.options());
  }
  a.A = A.data_ptr(); a.B = B.data_ptr(); a.C = C.data_ptr();
  a.lda = A.stride(0); a.ldb = B.stride(0); a.ldc = C.stride(0);
  a.epilogue = (int)epilogue;
  a.accumulate = accumulate; a.c_is_fp32 = C.scalar_type() == at::kFloat;
  a.a_mn_major = a_mn; a.b_mn_major = b_mn; a.two_cta = two_cta; a.no_split = no_split;
  a.direct_store = (int)direct_store; a.raster_n = (int)raster_n;
  a.p = (float)p; a.op_id = (uint32_t)op_id;
  if (aux_in.has_value()) {
    need(is_bf16_cuda(*aux_in) && aux_in->dim() == 2 && aux_in->size(0) == a.M && aux_in->size(1) == a.N && aux_in->stride(1) == 1, "gemm: bad aux_in");
    a.aux_in = aux_in->data_ptr(); a.ld_aux = aux_in->stride(0);
  }
  if (aux_out.has_value()) {
    need(is_bf16_cuda(*aux_out) && aux_out->dim() == 2 && aux_out->size(0) == a.M && aux_out->size(1) == a.N && aux_out->stride(1) == 1, "gemm: bad aux_out");
    a.aux_out = aux_out->data_ptr(); a.ld_aux = aux_out->stride(0);
  }
  if (rng.has_value()) {
    need(rng->is_cuda() && rng->scalar_type() == at::kLong && rng->numel() >= 2, "gemm: rng must be CUDA int64[2]");
    a.rng = (const uint64_t*)rng->data_ptr<int64_t>();
  }
  need(epilogue != ZTB_EPI_DROPOUT_ADD || a.rng != nullptr, "gemm: dropout epilogue needs rng");
  if (rowdot_out.has_value()) {
    need(rowdot_out->is_cuda() && rowdot_out->scalar_type() == at::kFloat && rowdot_out->is_contiguous() && rowdot_hd > 0 &&
         rowdot_out->numel() == (int64_t)a.M * (a.N / rowdot_hd), "gemm: rowdot_out must be fp32 contiguous with M * N / head_dim elements");
    a.rowdot_out = rowdot_out->data_ptr<float>(); a.rowdot_T = (int)rowdot_T; a.rowdot_hd = (int)rowdot_hd;
  }
  need(epilogue != ZTB_EPI_ROWDOT || (a.rowdot_out != nullptr && two_cta), "gemm: ROWDOT epilogue needs rowdot_out and two_cta");
  if (mask_out.has_value()) {
    need(two_cta && epilogue == ZTB_EPI_DROPOUT_ADD && a.N % 32 == 0 && mask_out->is_cuda() && mask_out->is_contiguous() && mask_out->element_size() == 4 &&
         mask_out->numel() == (int64_t)a.M * (a.N / 32), "gemm: mask_out must be a contiguous 32-bit [M, N/32] CUDA tensor (two_cta, dropout epilogue, N % 32 == 0)");
    a.mask_out = reinterpret_cast<unsigned int*>(mask_out->data_ptr());
  }
  check(ztb_gemm(&a, num_sms(), stream()), "gemm");
  return C;
}

// Forward GEMM C = A * B^T with the ZeRO-1 all-gather of B's bucket fused into the kernel.
at::Tensor gemm_ag(const at::Tensor& A, const at::Tensor& B, std::vector<int64_t> peer_ptrs, at::Tensor& local_flat, const at::Tensor& segs,
                   int64_t nseg, int64_t n_b_segs, int64_t total_tiles, int64_t shard_ready_ptr, int64_t epoch, at::Tensor& done,
                   int64_t b_tiles_per_rank, int64_t rows_per_rank, int64_t world, int64_t rank, bool two_cta, int64_t epoch_ptr, int64_t done_base) {
  need(is_bf16_cuda(A) && is_bf16_cuda(B) && A.dim() == 2 && B.dim() == 2 && A.stride(1) == 1 && B.stride(1) == 1, "gemm_ag: A, B bf16 2-D");
  need(A.size(1) == B.size(1), "gemm_ag: K mismatch");
  need(done.is_cuda() && done.numel() >= 8 && done.element_size() == 4, "gemm_ag: done must be an 8-element 32-bit CUDA tensor");
  need(is_bf16_cuda(local_flat), "gemm_ag: local flat buffer");
  c10::cuda::CUDAGuard g(A.device());
  auto C = at::empty({A.size(0), B.size(0)}, A.options());
  ZtbAgArgs ag{};
  for (size_t i = 0; i < peer_ptrs.size() && i < 8; ++i) ag.peer[i] = reinterpret_cast<const void*>(peer_ptrs[i]);
  ag.local = local_flat.data_ptr();
  ag.segs = reinterpret_cast<const ZtbSegment*>(segs.data_ptr()); ag.nseg = (int)nseg; ag.n_b_segs = (int)n_b_segs; ag.total_tiles = total_tiles;
  ag.shard_ready = reinterpret_cast<const int*>(shard_ready_ptr); ag.epoch = (int)epoch; ag.epoch_ptr = reinterpret_cast<const int*>(epoch_ptr);
  ag.done = reinterpret_cast<unsigned int*>(done.data_ptr()); ag.b_tiles_per_rank = (unsigned)b_tiles_per_rank;
  ag.rows_per_rank = (int)rows_per_rank; ag.world = (int)world; ag.rank = (int)rank; ag.done_base = done_base;
  ZtbGemmArgs a{};
  a.A = A.data_ptr(); a.B = B.data_ptr(); a.C = C.data_ptr();
  a.M = A.size(0); a.K = A.size(1); a.N = B.size(0);
  a.lda = A.stride(0); a.ldb = B.stride(0); a.ldc = C.stride(0);
  a.ag = &ag;
  a.direct_store = -1; a.raster_n = -1; a.two_cta = two_cta;
  check(ztb_gemm(&a, num_sms(), stream()), "gemm_ag");
  return C;
}

// Fused (reduce-scatter +) AdamW over this rank's segments.  grad_ptrs: device addresses of every
// peer's flat gradient buffer (first = local).  segs: uint8 CUDA tensor of packed ZtbSegment.
void adamw_shard(at::Tensor& master, at::Tensor& m, at::Tensor& v, at::Tensor& flat_param, std::vector<int64_t> grad_ptrs,
                 bool grad_is_fp32, const at::Tensor& segs, int64_t nseg, int64_t total_tiles, double lr, double beta1, double beta2,
                 double eps, double wd, double clip, double grad_scale, int64_t step, int64_t wait_flags_ptr, int64_t wait_value, int64_t max_ctas, int64_t scale_mul_ptr, int64_t mc_grad_ptr, int64_t mc_param_ptr,
                 int64_t dyn_ptr) {
  need(master.is_cuda() && master.scalar_type() == at::kFloat && m.scalar_type() == at::kFloat && v.scalar_type() == at::kFloat, "adamw_shard: fp32 state");
  need(is_bf16_cuda(flat_param), "adamw_shard: bf16 flat_param");
  need(segs.is_cuda() && segs.scalar_type() == at::kByte && segs.numel() >= (int64_t)(nseg * sizeof(ZtbSegment)), "adamw_shard: segs");
  c10::cuda::CUDAGuard g(master.device());
  std::vector<const void*> gp;
  for (auto x : grad_ptrs) gp.push_back(reinterpret_cast<const void*>(x));
  ZtbAdamArgs a;
  a.lr = (float)lr; a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = (float)eps; a.wd = (float)wd; a.clip = (float)clip;
  a.grad_scale = (float)grad_scale;
  a.scale_mul = reinterpret_cast<const float*>(scale_mul_ptr);
  a.dyn = reinterpret_cast<const float*>(dyn_ptr);
  a.inv_bc1 = (float)(1.0 / (1.0 - std::pow(beta1, (double)step)));
  a.inv_bc2 = (float)(1.0 / (1.0 - std::pow(beta2, (double)step)));
  a.omb1 = (float)(1.0 - beta1); a.omb2 = (float)(1.0 - beta2);
  check(ztb_adamw_shard(master.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(), flat_param.data_ptr(), gp.data(), (int)gp.size(),
                        grad_is_fp32, reinterpret_cast<const ZtbSegment*>(segs.data_ptr()), (int)nseg, total_tiles, a,
                        reinterpret_cast<const int*>(wait_flags_ptr), (int)wait_value, num_sms(), (int)max_ctas, reinterpret_cast<const void*>(mc_grad_ptr),
                        reinterpret_cast<void*>(mc_param_ptr), stream()), "adamw_shard");
}

int64_t segment_struct_size() { return sizeof(ZtbSegment); }

// qkv: (B, T, 3*H*hd) bf16 packed [q | k | v]; returns (out (B,T,H*hd) bf16, lse (B,H,T) fp32)
std::vector<at::Tensor> attention_fwd(const at::Tensor& qkv, const c10::optional<at::Tensor>& slopes, int64_t H, double scale,
                                      double p, const c10::optional<at::Tensor>& rng, int64_t op_id) {
  need(is_bf16_cuda(qkv) && qkv.dim() == 3 && qkv.is_contiguous(), "attention_fwd: qkv must be (B,T,3D) bf16 contiguous");
  c10::cuda::CUDAGuard g(qkv.device());
  int B = qkv.size(0), T = qkv.size(1), D = qkv.size(2) / 3, hd = D / H;
  auto out = at::empty({B, T, D}, qkv.options());
  auto lse = at::empty({B, (int)H, T}, qkv.options().dtype(at::kFloat));
  const float* sl = nullptr;
  if (slopes.has_value()) { need(slopes->is_cuda() && slopes->scalar_type() == at::kFloat && slopes->numel() == H, "attention_fwd: slopes fp32 [H]"); sl = slopes->data_ptr<float>(); }
  const uint64_t* r = rng.has_value() ? (const uint64_t*)rng->data_ptr<int64_t>() : nullptr;
  check(ztb_attn_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr<float>(), sl, B, T, (int)H, hd, (float)scale, (float)p, r, (uint32_t)op_id, stream()), "attention_fwd");
  return {out, lse};
}

at::Tensor attention_bwd(const at::Tensor& qkv, const at::Tensor& out, const at::Tensor& dout, const at::Tensor& lse,
                         const c10::optional<at::Tensor>& slopes, int64_t H, double scale, double p,
                         const c10::optional<at::Tensor>& rng, int64_t op_id, const c10::optional<at::Tensor>& dbg, bool dq_bf16,
                         const c10::optional<at::Tensor>& delta_in) {
  need(is_bf16_cuda(qkv) && qkv.dim() == 3 && qkv.is_contiguous() && is_bf16_cuda(out) && out.is_contiguous() && is_bf16_cuda(dout) && dout.is_contiguous(),
       "attention_bwd: bf16 contiguous tensors");
  c10::cuda::CUDAGuard g(qkv.device());
  int B = qkv.size(0), T = qkv.size(1), D = qkv.size(2) / 3, hd = D / H;
  auto dqkv = at::empty_like(qkv);
  const bool dq16 = dq_bf16 && hd == 128;
  at::Tensor acc;
  if (!dq16) acc = at::zeros({(int64_t)B * T, D}, qkv.options().dtype(at::kFloat));
  at::Tensor delta;
  if (delta_in.has_value()) {
    need(delta_in->is_cuda() && delta_in->scalar_type() == at::kFloat && delta_in->is_contiguous() && delta_in->numel() == (int64_t)B * H * T,
         "attention_bwd: delta must be fp32 contiguous [B, H, T]");
    delta = *delta_in;
  } else {
    delta = at::empty({B, (int)H, T}, qkv.options().dtype(at::kFloat));
  }
  const float* sl = slopes.has_value() ? slopes->data_ptr<float>() : nullptr;
  const uint64_t* r = rng.has_value() ? (const uint64_t*)rng->data_ptr<int64_t>() : nullptr;
  check(ztb_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr<float>(), dqkv.data_ptr(), dq16 ? nullptr : acc.data_ptr<float>(),
                     delta.data_ptr<float>(), sl, B, T, (int)H, hd, (float)scale, (float)p, r, (uint32_t)op_id, num_sms(), stream(),
                     dbg.has_value() ? reinterpret_cast<unsigned long long*>(dbg->data_ptr<int64_t>()) : nullptr, delta_in.has_value() ? 1 : 0), "attention_bwd");
  return dqkv;
}

std::vector<void*> to_ptrs(const std::vector<int64_t>& v) {
  std::vector<void*> o;
  for (auto x : v) o.push_back(reinterpret_cast<void*>(x));
  return o;
}

// release-store `value` into slot `slot` of every listed flag pad (device addresses, peers included)
void signal(std::vector<int64_t> pad_ptrs, int64_t slot, int64_t value, int64_t value_ptr) {
  auto p = to_ptrs(pad_ptrs);
  check(ztb_signal(p.data(), (int)p.size(), (int)slot, (int)value, stream(), reinterpret_cast<const int*>(value_ptr)), "signal");
}

void wait_flags(int64_t flags_ptr, int64_t n, int64_t value, int64_t value_ptr) {
  check(ztb_wait_flags(reinterpret_cast<const int*>(flags_ptr), (int)n, (int)value, stream(), reinterpret_cast<const int*>(value_ptr)), "wait_flags");
}

void allgather_pull(at::Tensor& flat_param, std::vector<int64_t> src_ptrs, const at::Tensor& segs, int64_t nseg, int64_t total_tiles,
                    int64_t flags_ptr, int64_t wait_value) {
  need(is_bf16_cuda(flat_param), "allgather_pull: bf16 flat_param");
  c10::cuda::CUDAGuard g(flat_param.device());
  auto p = to_ptrs(src_ptrs);
  check(ztb_allgather_pull(flat_param.data_ptr(), p.data(), (int)p.size(), reinterpret_cast<const ZtbSegment*>(segs.data_ptr()), (int)nseg,
                           total_tiles, reinterpret_cast<const int*>(flags_ptr), (int)wait_value, num_sms(), stream()), "allgather_pull");
}

void reduce_scatter(at::Tensor& out, std::vector<int64_t> grad_ptrs, const at::Tensor& segs, int64_t nseg, int64_t total_tiles,
                    int64_t flags_ptr, int64_t wait_value) {
  need(is_bf16_cuda(out), "reduce_scatter: bf16 out");
  c10::cuda::CUDAGuard g(out.device());
  auto p = to_ptrs(grad_ptrs);
  check(ztb_reduce_scatter(out.data_ptr(), p.data(), (int)p.size(), reinterpret_cast<const ZtbSegment*>(segs.data_ptr()), (int)nseg,
                           total_tiles, reinterpret_cast<const int*>(flags_ptr), (int)wait_value, num_sms(), stream()), "reduce_scatter");
}

void p2p_copy(int64_t dst_ptr, int64_t src_ptr, int64_t bytes) {
  check(ztb_p2p_copy(reinterpret_cast<void*>(dst_ptr), reinterpret_cast<const void*>(src_ptr), (size_t)bytes, num_sms(), stream()), "p2p_copy");
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("layernorm_fwd", &layernorm_fwd);
  m.def("layernorm_bwd", &layernorm_bwd, py::arg("dy"), py::arg("x"), py::arg("w"), py::arg("mean"), py::arg("rstd"), py::arg("dres") = py::none(),
        py::arg("grad_out") = py::none(), py::arg("accumulate") = false, py::arg("scratch") = py::none(), py::arg("ticket") = py::none(),
        py::arg("drop_p") = 0.0, py::arg("rng") = py::none(), py::arg("op_id") = 0, py::arg("deterministic") = false, py::arg("drop_mask") = py::none());
  m.def("dropout_add", &dropout_add);
  m.def("embedding_fwd", &embedding_fwd);
  m.def("embedding_bwd_", &embedding_bwd_);
  m.def("embedding_bwd_sorted_", &embedding_bwd_sorted_);
  m.def("gelu_fwd", &gelu_fwd);
  m.def("gelu_bwd", &gelu_bwd);
  m.def("ce_fwd", &ce_fwd);
  m.def("ce_bwd_", &ce_bwd_);
  m.def("gemm", &gemm, py::arg("A"), py::arg("B"), py::arg("C") = py::none(), py::arg("epilogue") = 0, py::arg("aux_in") = py::none(),
        py::arg("aux_out") = py::none(), py::arg("p") = 0.0, py::arg("rng") = py::none(), py::arg("op_id") = 0,
        py::arg("accumulate") = false, py::arg("a_mn") = false, py::arg("b_mn") = false, py::arg("two_cta") = false, py::arg("no_split") = false,
        py::arg("direct_store") = -1, py::arg("raster_n") = -1, py::arg("rowdot_out") = py::none(), py::arg("rowdot_T") = 0, py::arg("rowdot_hd") = 0, py::arg("mask_out") = py::none());
  m.def("gemm_ag", &gemm_ag, py::arg("A"), py::arg("B"), py::arg("peer_ptrs"), py::arg("local_flat"), py::arg("segs"), py::arg("nseg"),
        py::arg("n_b_segs"), py::arg("total_tiles"), py::arg("shard_ready_ptr"), py::arg("epoch"), py::arg("done"), py::arg("b_tiles_per_rank"),
        py::arg("rows_per_rank"), py::arg("world"), py::arg("rank"), py::arg("two_cta") = false, py::arg("epoch_ptr") = 0, py::arg("done_base") = -1);
  m.def("adamw_shard", &adamw_shard, py::arg("master"), py::arg("m"), py::arg("v"), py::arg("flat_param"), py::arg("grad_ptrs"),
        py::arg("grad_is_fp32"), py::arg("segs"), py::arg("nseg"), py::arg("total_tiles"), py::arg("lr"), py::arg("beta1"), py::arg("beta2"),
        py::arg("eps"), py::arg("wd"), py::arg("clip"), py::arg("grad_scale"), py::arg("step"), py::arg("wait_flags_ptr") = 0,
        py::arg("wait_value") = 0, py::arg("max_ctas") = 0, py::arg("scale_mul_ptr") = 0, py::arg("mc_grad_ptr") = 0, py::arg("mc_param_ptr") = 0, py::arg("dyn_ptr") = 0);
  m.def("segment_struct_size", &segment_struct_size);
  m.def("attention_fwd", &attention_fwd);
  m.def("attention_bwd", &attention_bwd, py::arg("qkv"), py::arg("out"), py::arg("dout"), py::arg("lse"), py::arg("slopes"), py::arg("H"),
        py::arg("scale"), py::arg("p"), py::arg("rng"), py::arg("op_id"), py::arg("dbg") = py::none(), py::arg("dq_bf16") = false, py::arg("delta") = py::none());
  m.def("signal", &signal, py::arg("pad_ptrs"), py::arg("slot"), py::arg("value"), py::arg("value_ptr") = 0);
  m.def("wait_flags", &wait_flags, py::arg("flags_ptr"), py::arg("n"), py::arg("value"), py::arg("value_ptr") = 0);
  m.def("allgather_pull", &allgather_pull);
  m.def("reduce_scatter", &reduce_scatter);
  m.def("p2p_copy", &p2p_copy);
}
