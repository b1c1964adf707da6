/* nvit_b200 C ABI — the drop-in boundary of the B200-native nViT training hot path.
 *
 * The reference (slobodaapl/nvit) is pure PyTorch and has no FFI of its own; every entry point below replaces a
 * group of eager PyTorch ops at the cited lines of /root/reference/nvit/{model,train}.py.  A maintainer binds them
 * with ctypes; nvit_b200/_lib.py is exactly that binding.
 *
 * Conventions
 *  - Plain pointers and sizes only.  All pointers are DEVICE pointers owned by the caller (PyTorch); kernels never
 *    allocate, free or retain them.  `stream` is a cudaStream_t passed as void*.
 *  - "bf16" buffers are raw uint16 bfloat16; "f32" are float.  Matrices are row-major with explicit leading dims
 *    (in elements) where given, dense otherwise.
 *  - Return 0 on success, negative on error (never throws, never exits); nvit_last_error() gives the message of the
 *    last failure on the calling thread.
 *  - Entry points are re-entrant and stream-ordered.
 *  - Tuning switches that select between equivalent kernel variants (tile mode, tile order) and the measurement-only
 *    hooks are NOT part of this header: see include/nvit_b200_tuning.h.
 */
#ifndef NVIT_B200_H_
#define NVIT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ------------------------------------------------------------------------------------------------- */
const char* nvit_last_error(void);
int nvit_version(void);
int nvit_sm_count(void);
/* Size persistent grids for at most n SMs (0 = all): leaves SMs free for concurrent NCCL kernels in data-parallel runs. */
int nvit_set_sm_budget(int n);
/* Programmatic dependent launch: with on != 0 every kernel of the library is launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization, so its on-chip set-up (barriers, TMEM allocation, descriptor
 * prefetch) and the launch latency overlap the tail of the previous kernel in the stream; each kernel executes
 * griddepcontrol.wait before its first global-memory access, so results are those of plain stream order.  Process-wide;
 * takes effect for launches (and graph captures) made after the call.  Default off. */
int nvit_set_pdl(int on);
/* TMA tensor maps are encoded once per distinct (base, type, shape, pitch, box) and kept in a process-wide table (the
 * engine's buffers are static, so a training step encodes none after its first run).  Counters since load: maps encoded
 * by the driver / maps served from the table.  Host pointers; either may be NULL. */
int nvit_tmap_cache_stats(int64_t* encodes, int64_t* hits);

/* ---- GEMM: nn.Linear / nn.Conv2d-as-GEMM forward, dgrad and wgrad (model.py:99-101,130,148,155,226-228,259,262,
 *      286-304,329-332,341-344 and their autograd backward) ------------------------------------------------------
 *   C[M,N] (+)= A · B^T, bf16 operands, fp32 accumulation in tensor memory (tcgen05).
 *   a_mn_major = 0: A stored [M,K] (lda = row pitch); 1: A stored [K,M] (M contiguous).  Likewise B ([N,K] / [K,N]).
 *   out_f32: C is float (else bf16).  C2_bf16 (optional, only with out_f32): second bf16 copy of the result.
 *   accumulate: C += result (fp32 only).  splits > 1: split the K range across CTAs, partial sums added with TMA
 *   reduce-add; splits <= 0 picks the split count that fills the SMs (fp32 outputs without epilogue only).
 *   Epilogue, in this order: + bias[N]; * colscale[N]*colscale_mul; + rowadd[row % rowadd_period, N]  (any may be NULL).
 *   swiglu_half F > 0: B holds [2F, K] (u rows then v rows, torch.chunk order of model.py:153); N must equal F.  The
 *     kernel computes both halves of a column pair in one CTA and writes C[M,F] = (u*cs_u) * silu(v*cs_v) as bf16
 *     (colscale has 2F entries or is NULL), and, if C2_bf16 != NULL, the raw bf16 product [M,2F] (ldc2) for backward.
 */
int nvit_gemm_bf16(const void* A, const void* B, void* C, void* C2_bf16, int64_t M, int64_t N, int64_t K, int64_t lda,
                   int64_t ldb, int64_t ldc, int64_t ldc2, int a_mn_major, int b_mn_major, int out_f32, int accumulate,
                   int splits, const float* bias, const float* colscale, float colscale_mul, const float* rowadd,
                   int64_t rowadd_period, int64_t swiglu_half, void* stream);


/* ---- casts / reductions ------------------------------------------------------------------------------------- */
/* autocast's weight/activation casts (torch.autocast around model.py:905 of train.py) */
int nvit_cast_f32_to_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);
/* out[0] += sum(x^2): global gradient norm of clip_grad_norm_ (train.py:938) */
int nvit_sumsq_f32(const float* x, int64_t n, float* out_accum, void* stream);
/* The same with a run-to-run and rank-to-rank reproducible result (no floating-point atomics): data-parallel replicas then
 * derive bit-identical clip coefficients from their bit-identical all-reduced gradients.  workspace: workspace_floats >= 2
 * device floats, workspace[0] zero before the FIRST use (the kernel leaves it zero); at most workspace_floats - 1 CTAs run. */
int nvit_sumsq_f32_det(const float* x, int64_t n, float* out_accum, float* workspace, int64_t workspace_floats, void* stream);
/* out[N] += column sums of a bf16 [M,N] matrix (bias gradients of nn.Linear when config.bias) */
int nvit_colsum_bf16(const void* x_bf16, int64_t M, int64_t N, int64_t ldx, float* out_accum, void* stream);
/* dpos[T,C] += sum_b dx[b,t,c]; dbias[C] += sum_{b,t} dx  (autograd of `+ pos_embed` and conv bias, model.py:407-415) */
int nvit_pos_bias_grad(const float* dx, int64_t B, int64_t T, int64_t C, float* dpos, float* dbias_accum, void* stream);

/* ---- normalized residual update (model.py:134-142, 159-167, 265-273; norm_skip model.py:84-87, 450-452) -------
 *   lr  = |alpha * alpha_mul|            (alpha_mul = 0.05 / base_scale)
 *   o   = N( N(h) + lr * (N(x) - N(h)) )                     N(v) = v / ||v||_2 over C, no epsilon
 *   out = h0 ? N( o * skip[0] + h0 ) : o
 *   fwd writes out as fp32 and/or bf16 (either may be NULL).
 *   bwd recomputes the forward from (h, x, h0) and, given g = dL/dout, writes dx (bf16), dh (fp32, += if
 *   dh_accumulate), dh0 (fp32, written), and accumulates dalpha[C] (w.r.t. the STORED alpha) and dskip[1].
 */
int nvit_residual_fwd(const float* h, const void* x_bf16, const float* alpha, float alpha_mul, const float* h0,
                      const float* skip, float* out_f32, void* out_bf16, int64_t M, int64_t C, void* stream);
int nvit_residual_bwd(const float* g, const float* h, const void* x_bf16, const float* alpha, float alpha_mul,
                      const float* h0, const float* skip, float* dh, int dh_accumulate, void* dx_bf16, float* dh0,
                      float* dalpha_accum, float* dskip_accum, int64_t M, int64_t C, void* stream);
/* Form of nvit_residual_bwd: 0 = each warp holds its row of every stream in registers (direct global loads);
 * 1 = the rows of the next iterations arrive through per-warp rings of 1-D bulk copies (cp.async.bulk + mbarrier) in
 * shared memory, one persistent CTA per SM; 2 (default) = whichever was measured faster for the variant and width
 * (staged for C >= 512 except the skip form without accumulation).  Same arithmetic, same results.  Process-wide. */
int nvit_residual_bwd_staged(int mode);

/* ---- original-ViT branch (config.use_nvit = False; BASELINE config 4) -------------------------------------------------
 *   add_rmsnorm : t = h (+ x_bf16, may be NULL);  y = t * rsqrt(mean(t^2) + eps) * w      (RMSNorm, model.py:172-184, applied
 *                 to the plain residual sum, model.py:95-96, 132-133, 145-146; cross-attention norms :221-223)
 *                 bwd: dh (+)= dt, dx = dt (bf16), dw[C] +=          (recomputes t from h and x)
 *   add_skipnorm: out = N((h + x) * skip[0] + h0)      (second residual add model.py:157-158 + norm_skip :84-87, 450-452)
 *                 bwd: dh (written, fp32) = dx (bf16) = d(h+x), dh0 (written), dskip[1] +=
 */
int nvit_add_rmsnorm_fwd(const float* h, const void* x_bf16, const float* w, float eps, float* y_f32, void* y_bf16,
                         int64_t M, int64_t C, void* stream);
int nvit_add_rmsnorm_bwd(const float* dy, const float* h, const void* x_bf16, const float* w, float eps, float* dh,
                         int dh_accumulate, void* dx_bf16, float* dw_accum, int64_t M, int64_t C, void* stream);
int nvit_add_skipnorm_fwd(const float* h, const void* x_bf16, const float* h0, const float* skip, float* out_f32,
                          void* out_bf16, int64_t M, int64_t C, void* stream);
int nvit_add_skipnorm_bwd(const float* g, const float* h, const void* x_bf16, const float* h0, const float* skip, float* dh,
                          void* dx_bf16, float* dh0, float* dskip_accum, int64_t M, int64_t C, void* stream);

/* ---- suv-scaled SiLU gate, unfused form (model.py:150-154, 259-261).  uv [M,2F] bf16: u = cols [0,F), v = [F,2F).
 *   x = (u*su) * silu(v*sv),  s = suv * suv_mul (suv NULL -> 1).  bwd: duv bf16 [M,2F], dsuv[2F] += (w.r.t. stored suv).
 */
int nvit_swiglu_fwd(const void* uv_bf16, const float* suv, float suv_mul, void* x_bf16, int64_t M, int64_t F, void* stream);
int nvit_swiglu_bwd(const void* dx_bf16, const void* uv_bf16, const float* suv, float suv_mul, void* duv_bf16,
                    float* dsuv_accum, int64_t M, int64_t F, void* stream);

/* ---- unit-norm QK attention (model.py:104-127, 231-258) ------------------------------------------------------
 *   q,k,v: bf16, token-major [B*T, ld*] with head h at columns [h*D, (h+1)*D).   D must be 64.
 *   qh = s * N_D(q), kh = s * N_D(k) with s = sqk * sqk_mul ([H*D]; sqk NULL -> no normalization, s = 1),
 *   out[B*T, H*D] = softmax(scale * qh kh^T) v  (non-causal).  lse [B,H,T] (fp32) is saved for backward.
 *   bwd: dq, dk, dv (bf16, same layouts) and dsqk[H*D] += (w.r.t. stored sqk).
 *   inv_q / inv_k (both or neither; need sqk): q and k are ALREADY qh / kh (written by nvit_gemm_qknorm) and
 *   inv_*[token * ld_inv_* + h] = 1 / ||x|| of the raw projection: the kernels skip their normalisation pass; dq / dk are
 *   still the gradients w.r.t. the raw projections.
 */
int nvit_attention_fwd(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                       const float* sqk, float sqk_mul, float scale, void* out, int64_t ldo, float* lse, int64_t B,
                       int64_t H, int64_t T, int64_t D, const float* inv_q, const float* inv_k, int64_t ld_inv_q,
                       int64_t ld_inv_k, void* stream);
int nvit_attention_bwd(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                       const float* sqk, float sqk_mul, float scale, const void* out, const void* dout, int64_t ldo,
                       const float* lse, void* dq, void* dk, void* dv, int64_t lddq, int64_t lddk, int64_t lddv,
                       float* dsqk_accum, int64_t B, int64_t H, int64_t T, int64_t D, const float* inv_q, const float* inv_k,
                       int64_t ld_inv_q, int64_t ld_inv_k, void* stream);

/* q/k/v projection with the unit-norm treatment of nViT in the epilogue (model.py:99-119, 226-249):
 *   C[M,N] = A[M,K] B[N,K]^T (+ bias);  for columns < norm_cols every 64-column head h of a row becomes
 *   scale[col % scale_period] * scale_mul * x / ||x||  and  inv_out[row * ld_inv + col / 64] = 1 / ||x||  (0 for a zero row);
 * columns >= norm_cols (the value projection) are stored as they are.  The attention kernels then take the normalised q/k
 * and the inverse norms (inv_q / inv_k arguments) and skip their own normalisation pass. */
int nvit_gemm_qknorm(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc,
                     const float* bias, const float* scale, float scale_mul, int64_t scale_period, int64_t norm_cols, float* inv_out,
                     int64_t ld_inv, void* stream);

/* Gate backward fused behind the mlp_c_proj / out_proj dgrad GEMM (model.py:152-155, 260-262 backward):
 *   dx = dY[M,K] W[K,F]            (W = the projection weight as it lies, [K, F] bf16; dx stays on chip, fp32)
 *   d_uv[:, :F]  = dx * silu(v sv) * su          d_uv[:, F:] = dx * (u su) * silu'(v sv) * sv
 * with u|v = uv_raw[M, 2F] (bf16, the c_fc / proj output saved by the forward pass) and su|sv = suv * suv_mul (suv NULL:
 * ones).  F must be a multiple of 64.  dL/dsuv is NOT produced here: see nvit_rowdot_div. */
int nvit_gemm_gate_bwd(const void* dY, const void* W, const void* uv_raw, const float* suv, float suv_mul, void* d_uv,
                       int64_t M, int64_t F, int64_t K, int64_t ld_dy, int64_t ld_w, int64_t ld_uv, int64_t ld_duv, void* stream);
/* out[r] = sum_k w[r,k] dw[r,k] / div[r] (0 where div == 0), overwriting out.  With w = c_fc.weight, dw = its gradient and
 * div = suv this is dL/dsuv (uv = suv mul (h W^T) makes dL/dsuv[c] = W[c,:].dW[c,:] / suv[c]): a 19 MB read instead of a
 * column reduction over [M, 8C] (model.py:148-151 backward). */
int nvit_rowdot_div(const float* w, const float* dw, const float* div, float* out, int64_t rows, int64_t cols, void* stream);

/* ---- patch embedding operand (model.py:286-304, 407-408; reconstruction target model.py:460-463) --------------
 *   out[(b,i,j), (c,kh,kw)] = reflect_pad(img, pad)[b, c, i*stride + kh, j*stride + kw]   as bf16, [B*g*g, ch*ksize^2]
 */
int nvit_im2col_bf16(const float* img, void* out_bf16, int64_t B, int64_t ch, int64_t S, int64_t ksize, int64_t stride,
                     int64_t pad, void* stream);
/* The same operand straight from uint8 HWC images ([B, S, S, ch], what the data loader holds before torchvision's ToTensor),
 * with ToTensor + Normalize folded in (train.py:266-273, 1081-1092): value = pixel * scale + shift, e.g. scale = 1/(255 std),
 * shift = -mean/std.  One byte per sample crosses PCIe instead of four. */
int nvit_im2col_u8(const void* img_u8_nhwc, void* out_bf16, int64_t B, int64_t ch, int64_t S, int64_t ksize, int64_t stride,
                   int64_t pad, float scale, float shift, void* stream);
/* AutoAugment on the device, before the patch gather (replaces kornia.augmentation.auto.AutoAugment(dataset) in the
 * DataLoader workers, train.py:1081-1092 / 262-273).  src, dst: uint8 [B, S, S, 3], distinct buffers.  Per image b two
 * operations are applied in order: ops[b*2 + k] with parameters params[(b*2 + k)*8 ..]:
 *   0 identity
 *   1 affine      p0..p5 = m00 m01 ox m10 m11 oy: source pixel = M (x - c, y - c) + o, c = (S-1)/2 (already inside o);
 *                 nearest neighbour (ties to even), zero fill         (ShearX/Y, TranslateX/Y, Rotate)
 *   2 brightness  p0 = ratio, p1 = 1 - ratio: trunc(clamp(ratio * x + (1 - ratio) * 0))
 *   3 color       blend with the grey image trunc(0.2989 r + 0.587 g + 0.114 b)
 *   4 contrast    blend with the mean of the grey image
 *   5 sharpness   blend with the 3x3 smoothed image ([1 1 1; 1 5 1; 1 1 1] / 13, rounded; on the border the image itself)
 *   6 posterize   p0 = byte mask (256 - 2^(8 - bits))       7 solarize  p0 = threshold: x >= p0 ? 255 - x : x
 *   8 autocontrast (per channel: (x - min) * ((1 / (max - min)) * 255))   9 equalize (per channel histogram)   10 invert
 * Codes outside 0..10 act as identity.  uint8 semantics of every operation are those of torchvision's tensor kernels
 * (bit-exact; the affine map differs from grid_sample on exact ties only).  One CTA per image with the intermediate
 * image in shared memory: 3 S^2 + 13 KB must fit 227 KB (S <= 267).  ops / params are device pointers. */
int nvit_augment_u8(const void* src_u8_nhwc, void* dst_u8_nhwc, const int32_t* ops, const float* params, int64_t B, int64_t S,
                    int64_t ch, void* stream);

/* ---- classifier head (model.py:455-456, 466-468): mean over T -> LayerNorm(eps) -> (GEMM) ------------------- */
int nvit_pool_ln_fwd(const float* h, const float* gamma, const float* beta, float eps, void* y_bf16, float* xhat,
                     float* rstd, int64_t B, int64_t T, int64_t C, void* stream);
/* dy bf16 [B,C] (grad wrt LayerNorm output) -> dh[B,T,C] (fp32, written), dgamma/dbeta += */
int nvit_pool_ln_bwd(const void* dy_bf16, const float* gamma, const float* xhat, const float* rstd, float* dh,
                     float* dgamma_accum, float* dbeta_accum, int64_t B, int64_t T, int64_t C, void* stream);
/* logits[B,N] = raw[B,N] * sz[N] * sz_mul (model.py:466-468): the head GEMM runs once and this scales its fp32 output */
int nvit_head_scale_fwd(const float* raw, const float* sz, float sz_mul, float* logits, int64_t B, int64_t N, void* stream);
/* logits = raw * sz_eff (model.py:466-468);  bwd: draw_bf16 = dlogits * sz_eff, dsz[N] += sum_b dlogits*raw*sz_mul */
int nvit_head_scale_bwd(const float* dlogits, const float* raw, const float* sz, float sz_mul, void* draw_bf16,
                        float* dsz_accum, int64_t B, int64_t N, int64_t ld_draw, void* stream);
/* softmax cross-entropy, mean over batch (train.py:906): loss[0] += mean CE; dlogits = (softmax - onehot)*gscale/B.
 * A target outside [0, N) contributes no loss and a zero gradient row (never an out-of-bounds read); the mean divides by B. */
int nvit_cross_entropy(const float* logits, const int64_t* target, float* loss, float* dlogits, float gscale,
                       int64_t B, int64_t N, void* stream);
/* reconstruction loss (model.py:459-464): out[0] += sum (tanh(pred) - target)^2 * inv_count */
int nvit_tanh_mse(const void* pred_bf16, const void* target_bf16, int64_t n, float inv_count, float* out_accum,
                  void* stream);

/* ---- Kohonen maps (BASELINE config 5; /root/reference/nvit/kohonen.py, model.py:417-445, 482-561) ---------------------
 * The token-to-node distance matrix is a GEMM: dots[M,G] = x[M,C] nodes[G,C]^T through nvit_gemm_bf16 on bf16 hi/lo
 * splits (x_hi n_hi + x_hi n_lo + x_lo n_hi, fp32 accumulate: ~2^-17 relative, replacing torch.cdist kohonen.py:110). */
/* x fp32 -> hi = bf16(x) (optional), lo = bf16(x - hi) */
int nvit_split_bf16(const float* x, void* hi_bf16_or_null, void* lo_bf16, int64_t n, void* stream);
/* per node: squared norm, bf16 hi/lo GEMM operands and a copy of the table before the in-forward update */
int nvit_som_prepare(const float* nodes, int64_t G, int64_t C, float* node_sq, void* hi_bf16, void* lo_bf16, float* snapshot,
                     void* stream);
/* KohonenMap.forward (kohonen.py:100-119): idx[r] = argmin_g (node_sq[g] - 2 dots[r,g]) (lowest index on ties),
 * repr[r] = nodes[idx[r]] (fp32 + bf16), optional int64 copy of idx, one-hot bf16 [M,G] and counts[g] += #tokens */
int nvit_som_select(const float* dots, const float* node_sq, const float* nodes, int64_t M, int64_t G, int64_t C, int32_t* idx,
                    int64_t* idx64_or_null, void* onehot_bf16_or_null, float* counts_or_null, float* repr32, void* repr_bf16,
                    void* stream);
/* kohonen.py:149-155: out[r] = mean(x[r*run : (r+1)*run]) */
int nvit_som_pool(const float* x, int64_t rows, int64_t run, float* out, void* stream);
/* KohonenMap.update_nodes (kohonen.py:121-165): for i < steps, in order: nodes += s_i (pooled[i] - nodes),
 * s_i[g] = coef * exp(-torus_dist2(g, bmu[i]) / (2 sigma^2)); coef = lr * alpha is read from device memory */
int nvit_som_update(float* nodes, const float* pooled, const int32_t* bmu, int64_t steps, int64_t grid_rows, int64_t grid_cols,
                    int64_t C, const float* coef_dev, float sigma, void* stream);
/* consistency + the two quantization (Huber) losses (model.py:437, 441-442, 491-500): sums3 += {sum cos, sum huber_l,
 * sum huber_g}; with weights3 (device: consistency, local q., global q. weight x incoming gradient) the gradients are
 * ADDED to d_repr_l, d_repr_g, d_x_l, d_x_g (fp32 [M,C]) */
int nvit_som_pair_losses(const float* repr_l, const float* repr_g, const float* x_l, const float* x_g, int64_t M, int64_t C,
                         float* sums3_or_null, const float* weights3_or_null, float* d_repr_l, float* d_repr_g, float* d_x_l,
                         float* d_x_g, void* stream);
/* map smoothness (model.py:503-561) from the unit histogram: loss += sum_g counts[g] sum_k ||n_g - n_nb(g,k)|| / (8 M);
 * with a device weight the gradient is added to gnodes */
int nvit_som_smoothness(const float* nodes, const float* counts, int64_t side, int64_t C, int64_t M, float* loss_accum,
                        const float* weight_or_null, float* gnodes_or_null, void* stream);
/* d/dpred of weight * mean((tanh(pred) - target)^2), bf16 (reconstruction head backward, train.py:925-926) */
int nvit_tanh_mse_bwd(const void* pred_bf16, const void* target_bf16, int64_t n, float inv_count, const float* weight_dev,
                      void* dpred_bf16, void* stream);

/* ---- optimizer tail ------------------------------------------------------------------------------------------
 * AdamW over one flat fp32 buffer (torch.optim.AdamW semantics, model.py:369-385; clip train.py:935-938).
 * Elements [0,n_decay) get weight decay.  gnorm_sq (device scalar, may be NULL) holds sum(g^2); the clip coefficient
 * min(1, max_norm / (sqrt(gnorm_sq) + 1e-6)) is applied to g on the fly.  step is the 1-based step count.
 * dev_lr_step (optional device float[2] = {learning rate, step count}) overrides lr/step so that a captured CUDA graph
 * of the training step can be replayed while the schedule advances.
 */
int nvit_adamw_flat(float* p, const float* g, float* m, float* v, int64_t n, int64_t n_decay, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int64_t step, const float* gnorm_sq, float max_norm,
                    const float* dev_lr_step, void* stream);
/* The whole optimizer tail in ONE pass (train.py:935-946 clip + AdamW + zero_grad, train.py:461-480
 * normalize_matrices, and the bf16 weight casts of the next autocast forward): for every element p, g, m, v are read once,
 * updated with nvit_adamw_flat's arithmetic, and - for the tensors normalize_matrices touches - the updated row / column
 * is L2-normalised before it is written back as fp32 and as the bf16 GEMM operand; g is zeroed if zero_grad != 0.
 * p, g, m, v are the flat buffers (same layout); w16 the flat bf16 operand buffer.  table_dev: n_segments entries of
 * 8 x int64 {element offset, rows, cols, kind, bf16 element offset or -1, weight-decay flag, first_unit, 0}; kind 0 = plain
 * (unit = 8192 elements), 1 = normalise every row over cols (unit = 32 rows), 2 = normalise every column over rows (unit =
 * 128 columns); first_unit = running sum of units.  unit_counter_zeroed: device uint32 that must be 0 at launch (units
 * are claimed dynamically).  The other arguments are those of nvit_adamw_flat. */
int nvit_adamw_norm_fused(float* p, float* g, float* m, float* v, void* w16_bf16, const int64_t* table_dev, int64_t n_segments,
                          int64_t total_units, float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                          const float* gnorm_sq, float max_norm, const float* dev_lr_step, uint32_t* unit_counter_zeroed,
                          int zero_grad, void* stream);
/* Trainer.normalize_matrices (train.py:461-480) as ONE launch over a device table of n_tensors entries, each
 * 6 x int64: {w_f32 ptr, w_bf16 ptr or 0, rows, cols, axis, first_unit}.  axis = 1 normalizes every row over cols,
 * axis = 0 every column over rows (the reference's norm dims).  Units: 8 rows (axis 1) or 128 columns (axis 0) each;
 * first_unit is the running sum of units before the tensor; total_units the grand total.  The optional bf16 pointer
 * receives the GEMM operand copy in the same pass.
 */
int nvit_weight_norm_multi(const int64_t* table_dev, int64_t n_tensors, int64_t total_units, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NVIT_B200_H_ */
