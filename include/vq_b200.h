/*
 * vq_b200.h -- C ABI of the B200-native vector-quantiser hot path (libvq_b200.so).
 *
 * The reference (hongrui16/VQ-VAE-GAN-Diffusion) is pure Python and has no FFI layer; its boundary for this
 * path is the class network/vqvae/submodule/codebook.py::CodeBook (codebook.py:13-111).  This header is the
 * C-ABI a binding for that class calls (see INTEGRATION.md for the ctypes stub a maintainer would add to the
 * reference).  Each entry point names the reference lines it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the caller owns all memory;
 *   - no entry point allocates, frees or synchronises the device; all work is enqueued on `stream`;
 *   - return value 0 = OK, otherwise a negative VQ_E_* code or a positive cudaError_t;
 *     vq_last_error() gives a thread-local message;
 *   - the library is re-entrant (forward runs on the caller's thread, backward on PyTorch's autograd thread);
 *   - D (latent_dim) must be 256 for the NCHW CodeBook entry points -- the value of every reference config
 *     (configs/*.yml) -- else VQ_E_UNSUPPORTED; the row-major searches (vq_argmin_rows) take any 1 <= D <= 512;
 *   - N = B*HW latents in (b, h, w) order, z is contiguous NCHW (B, D, HW).
 *   - there is NO CPU fallback: on a device that is not sm_100 every compute entry point fails with
 *     VQ_E_DEVICE.
 */
#ifndef VQ_B200_H_
#define VQ_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* vq_stream_t; /* == cudaStream_t */

#define VQ_OK             0
#define VQ_E_INVALID     -1   /* bad argument (null pointer, negative size, misaligned pointer) */
#define VQ_E_UNSUPPORTED -2   /* D != 256, K < 1, ... */
#define VQ_E_WORKSPACE   -3   /* workspace too small */
#define VQ_E_DEVICE      -4   /* current device is not sm_100 */

/* stats[] slots written by vq_argmin / vq_forward (uint64 each, overwritten) */
#define VQ_STAT_TIE_ROWS      0  /* rows whose minimal fp32 distance is attained by >= 2 codes */
#define VQ_STAT_RERANK_ROWS   1  /* rows that needed the exact-fp32 re-rank (>= 2 candidates within the margin) */
#define VQ_STAT_FALLBACK_ROWS 2  /* rows resolved by the full fp32 row scan (candidate list overflow) */
#define VQ_STAT_CANDIDATES    3  /* candidates that survived the margin filter, summed over the rows the exact stage
                                    decided (rows resolved by the full scan are counted in FALLBACK_ROWS only) */
#define VQ_STAT_COUNT         4

/* fp32 distance recipes of the reference's nearest-code searches (vq_argmin_rows) */
#define VQ_RECIPE_EXPANDED 0  /* |x|^2 + |e|^2 - 2 x.e   codebook.py:70-79, diffusion_gaussian2d.py:334-339 */
#define VQ_RECIPE_DIFFSQ   1  /* sum_d (x_d - e_d)^2     continous_vq_diffusion/v_vq_diffusion.py:114-123 */
#define VQ_RECIPE_CDIST_NORMALIZED 2  /* cdist(normalize(x), normalize(T)): vqDiffusion/submodule/diffusion_gaussian3d.py:543-570 */

int         vq_abi_version(void);
const char* vq_last_error(void);

/* 0 if the CURRENT device can run the kernels (compute capability 10.0), else VQ_E_DEVICE. */
int vq_device_check(void);

/* Padded codebook rows: K rounded up to the code-tile size (256).  E_h holds K_pad*D fp16, e_norm2 K_pad floats. */
int vq_padded_codes(int K);

/* Bytes of scratch vq_argmin / vq_forward need for N latents. */
int vq_workspace_bytes(int64_t N, int K, int D, size_t* out_host);

/*
 * Derived codebook state; call again whenever the weight changed.
 * Replaces: torch.sum(self.codebook.weight**2, dim=1)  (codebook.py:74) and the operand conversion the
 * reference's sgemm does implicitly.
 *   E            (K, D) fp32  nn.Embedding weight (codebook.py:40)
 *   E_h          (K_pad, D) fp16 tensor-core operand: E * 2^s with s chosen so max|E * 2^s| is in [2^14, 2^15)
 *                (an exact scaling); rows >= K zero
 *   e_norm2      (K_pad) fp32 |e_k|^2 in the oracle's canonical order, rows >= K = +inf
 *   cb_scalars   (4) fp32: [0] max_k |e_k|^2, [1] max |E|, [2] 2^-s, [3] reserved
 */
int vq_prepare_codebook(const float* E, int K, int D, void* E_h, float* e_norm2, float* cb_scalars,
                        vq_stream_t stream);

/*
 * Tokeniser mode: indices only.
 * Replaces codebook.py:62-82 as reached from VQVAE.encode (network/vqvae/vqvae.py:139-146) by
 * VQTransformer.encode_to_z (network/vqTransformer/vqTransformer.py:64-81) and VQDiffusion.encode_to_z
 * (network/vqDiffusion/vqDiffusion.py:140-156), which discard z_q and the loss.
 *   idx    (N) int64 -- torch.argmin semantics (first minimum) of the fp32 distance formula
 *   stats  (VQ_STAT_COUNT) uint64, optional (may be NULL)
 */
int vq_argmin(const float* z_nchw, int64_t B, int64_t HW, int D,
              const float* E, const void* E_h, const float* e_norm2, const float* cb_scalars, int K,
              int64_t* idx, unsigned long long* stats,
              void* workspace, size_t workspace_bytes, vq_stream_t stream);

/*
 * Tokeniser mode with narrow indices (the token stream fed to the stage-2 models,
 * network/vqTransformer/vqTransformer.py:79,117-141, never needs 64 bits): as vq_argmin, but idx holds
 * idx_bits-wide integers -- 64 (int64), 32 (int32) or 16 (uint16, K <= 65536).  Opt-in: the reference's own dtype
 * is int64.
 */
int vq_argmin_narrow(const float* z_nchw, int64_t B, int64_t HW, int D,
                     const float* E, const void* E_h, const float* e_norm2, const float* cb_scalars, int K,
                     void* idx, int idx_bits, unsigned long long* stats,
                     void* workspace, size_t workspace_bytes, vq_stream_t stream);

/*
 * Nearest-code search over ROW-MAJOR vectors, first-minimum argmin (torch.argmin), for callers
 * that hold (N, D) rows instead of an NCHW latent grid.  `recipe` selects the reference's fp32 formula:
 *   VQ_RECIPE_EXPANDED  replaces GaussianDiffusion2D.gaussian_to_indices
 *                       (network/vqDiffusion/submodule/diffusion_gaussian2d.py:322-347: gaussian_flat (B*L, gaussian_dim)
 *                       against gaussian_lookup_table (K, gaussian_dim)) -- the CodeBook's own formula;
 *   VQ_RECIPE_DIFFSQ    replaces the search at the end of V_VQDiffusion.sample
 *                       (network/continous_vq_diffusion/v_vq_diffusion.py:114-123: sum((x - e)^2) over a broadcast
 *                       (B, L, K, D) difference, which this path never materialises).
 *   VQ_RECIPE_CDIST_NORMALIZED  replaces VQGaussianDiffusion3DWrapper.gaussian_to_indices
 *                       (network/vqDiffusion/submodule/diffusion_gaussian3d.py:543-570): both sides L2-normalised
 *                       (F.normalize), torch.cdist, argmin.  The query rows are normalised inside the call; E / E_h /
 *                       e_norm2 must describe the NORMALISED table (vq_normalize_rows, then vq_prepare_codebook).
 * Any width 1 <= D <= 512 (gaussian_dim = 96 in configs/*.yml, 512 in the 3D wrapper's own runs): the contraction is
 * padded to a multiple of 64 inside the kernels, nothing is padded in memory; vq_workspace_bytes and vq_prepare_codebook
 * take the same D.
 *   x_rows  (N, D) fp32, contiguous
 *   idx     (N) integers of idx_bits (64 / 32 / 16)
 */
int vq_argmin_rows(const float* x_rows, int64_t N, int D,
                   const float* E, const void* E_h, const float* e_norm2, const float* cb_scalars, int K,
                   int recipe, void* idx, int idx_bits, unsigned long long* stats,
                   void* workspace, size_t workspace_bytes, vq_stream_t stream);

/* F.normalize(x, p=2, dim=-1) of row-major vectors: out = x / max(|x|_2, 1e-12) (diffusion_gaussian3d.py:560-563), the norm
 * summed in the oracle's canonical order.  1 <= D <= 512. */
int vq_normalize_rows(const float* x_rows, int64_t N, int D, float* out_rows, vq_stream_t stream);

/*
 * Training forward.  Replaces CodeBook.forward, codebook.py:47-111.
 *   zq_nhwc (N, D) fp32 -- fl(z + fl(e - z)) in NHWC memory (the caller exposes it as the NCHW view the
 *                          reference returns, codebook.py:109); may be NULL when nobody reads z_q (a caller that folds
 *                          post_quant_conv into a lookup, see postconv.py): saves the 4 D bytes per latent it costs
 *   idx     (N) int64
 *   loss    (1) fp32    -- mean((e-z)^2 + beta*mean((e-z)^2)), codebook.py:96-103
 *   hist    (K) int64   -- bincount(idx, minlength=K); optional (may be NULL); overwritten
 */
int vq_forward(const float* z_nchw, int64_t B, int64_t HW, int D,
               const float* E, const void* E_h, const float* e_norm2, const float* cb_scalars, int K,
               float beta, float* zq_nhwc, int64_t* idx, float* loss, int64_t* hist,
               unsigned long long* stats, void* workspace, size_t workspace_bytes, vq_stream_t stream);

/*
 * Backward of CodeBook.forward as autograd derives it:
 *   grad_z = g_out + g_loss * 2 (z - e) / (n_global * D)              (contiguous NCHW)
 *   grad_E[idx[n]] += g_loss * beta * 2 (e - z) / (n_global * D)      (dense (K, D), zeroed here first)
 *   gout          upstream gradient on z_q, logical (B, D, HW); element strides gout_strides_host[3] =
 *                 {b, d, hw}; may be NULL (no upstream gradient)
 *   g_loss        upstream gradient on the scalar loss, as a host value; if g_loss_dev is non-NULL the value is
 *                 read from that device scalar instead (lets an autograd backward run without a host sync)
 *   n_global      latents the loss mean ran over; pass N on one device, the global N when the batch is sharded
 *   grad_z_nchw   may be NULL (z does not require grad); grad_E may be NULL (frozen codebook)
 */
int vq_backward(const float* gout, const int64_t* gout_strides_host, float g_loss, const float* g_loss_dev,
                const float* z_nchw, const int64_t* idx, const float* E,
                int64_t B, int64_t HW, int D, int K, float beta, int64_t n_global,
                float* grad_z_nchw, float* grad_E, vq_stream_t stream);

/*
 * vq_backward with two extras (new in this build; the reference trains on one device with autograd's own kernels):
 *   grad_E_scale   factor applied to the codebook gradient only -- 1/W when W data-parallel ranks SUM their
 *                  gradients, so that the sum is the gradient of the mean loss over the global batch while grad_z and
 *                  the returned loss stay the local-mean quantities DDP expects (see dist.py);
 *   deterministic  non-zero: the scatter-add of grad_E (embedding_dense_backward in the reference, whose CUDA kernel
 *                  is itself order-dependent) runs in 64-bit fixed point -- integer addition is associative, so the
 *                  result is bit-reproducible from run to run whatever order the atomics land in; needs `workspace`
 *                  of vq_backward_workspace_bytes(K, D) bytes (256-byte aligned).  0: red.global.add.v4.f32, no
 *                  workspace needed (may be NULL);
 *   code_diff_sum  optional (K, D): the per-code sums of (e - z) a vq_forward_ex call accumulated (possibly summed over
 *                  data-parallel ranks since).  When given, grad_E = g_loss * beta * 2 / (n_global * D) * grad_E_scale *
 *                  code_diff_sum is all that is left of the codebook gradient (no scatter-add, `deterministic` ignored).
 */
int vq_backward_workspace_bytes(int K, int D, size_t* out_host);
int vq_backward_ex(const float* gout, const int64_t* gout_strides_host, float g_loss, const float* g_loss_dev,
                   const float* z_nchw, const int64_t* idx, const float* E,
                   int64_t B, int64_t HW, int D, int K, float beta, int64_t n_global,
                   float grad_E_scale, int deterministic, const float* code_diff_sum, float* grad_z_nchw, float* grad_E,
                   void* workspace, size_t workspace_bytes, vq_stream_t stream);

/*
 * vq_forward that also accumulates, per code, the sum of (e - z) over the latents it won (`code_diff_sum`, (K, D) fp32,
 * zeroed here; may be NULL = plain vq_forward).  That sum is the codebook gradient up to a scalar (vq_backward_ex), so the
 * backward's scatter-add disappears and -- the point -- a data-parallel wrapper can all-reduce it right after the forward,
 * while the rest of the step runs, instead of after the backward (dist.py).
 */
int vq_forward_ex(const float* z_nchw, int64_t B, int64_t HW, int D,
                  const float* E, const void* E_h, const float* e_norm2, const float* cb_scalars, int K,
                  float beta, float* zq_nhwc, int64_t* idx, float* loss, int64_t* hist, float* code_diff_sum,
                  unsigned long long* stats, void* workspace, size_t workspace_bytes, vq_stream_t stream);

/*
 * quant_conv folded into the quantiser (encoder side): the reference runs
 * `quant_x = self.quant_conv(encoded_images)` -- Conv2d(256, 256, 1), network/vqvae/vqvae.py:83,128 -- right before the CodeBook.
 *
 * vq_prepare_quant_conv: the convolution weight W (256 output x 256 input channels, fp32 row-major = Conv2d.weight) ->
 *   w_img      256 KiB, 16-byte aligned: hi / lo fp16 operand images of W (split precision, see csrc/vq_qconv.cuh)
 *   w_scalars  4 floats ([0] = inverse operand scale)
 * Re-run after every change of W.
 *
 * vq_forward_qconv: vq_forward on z = W h + bias, the convolution computed with fp32 accuracy on the tensor cores inside the
 * operand-preparation kernel (no separate convolution pass, no re-read of z):
 *   h_nchw   (B, 256, HW) fp32 in, 16-byte aligned, HW % 128 == 0 (else VQ_E_UNSUPPORTED: run the convolution separately)
 *   bias     (256) fp32 or NULL
 *   z_nchw   (B, 256, HW) fp32 OUT: the convolution's output, |z - conv_fp32(h)| <= 1e-5 max|z|; everything else is computed
 *            from exactly these values as vq_forward would (indices / z_q / histogram bit-exact given z_nchw)
 * All other arguments as vq_forward.
 */
int vq_prepare_quant_conv(const float* W, void* w_img, float* w_scalars, vq_stream_t stream);
int vq_forward_qconv(const float* h_nchw, int64_t B, int64_t HW, int D, const void* w_img, const float* w_scalars,
                     const float* bias, float* z_nchw, const float* E, const void* E_h, const float* e_norm2,
                     const float* cb_scalars, int K, float beta, float* zq_nhwc, int64_t* idx, float* loss, int64_t* hist,
                     unsigned long long* stats, void* workspace, size_t workspace_bytes, vq_stream_t stream);

/*
 * SUM all-reduce, in place, of a symmetric fp32 buffer over NVLink with the NVSwitch doing the additions (NVLS: multimem.ld_reduce /
 * multimem.st on the buffer's multicast address) -- the data-parallel exchange of the codebook gradient / usage histogram (dist.py;
 * new in this build, the reference is single-device).  One process per GPU; every rank calls it with the same n_floats.
 *   multicast_ptr     multicast address of the buffer (torch.distributed._symmetric_memory handle.multicast_ptr); the caller's
 *                     prior work on `stream` must have produced this rank's contribution in its own copy of the buffer
 *   signal_pads_dev   DEVICE array of `world` pointers to the ranks' signal pads (handle.signal_pad_ptrs_dev), zero when idle;
 *                     uses the first `world` words of each pad
 *   n_floats          multiple of 4 * world
 *   local_sync        two 32-bit words of this rank's own device memory, zero before the first call (the kernel re-arms them):
 *                     only one CTA per rank synchronises with the other GPUs, the rest of the grid through these words
 * The device-side waits spin with a bound (a rank that never arrives traps instead of hanging the box).
 */
int vq_allreduce_multimem(void* multicast_ptr, void* const* signal_pads_dev, int rank, int world, int64_t n_floats,
                          unsigned int* local_sync, vq_stream_t stream);

/*
 * Tail of the data-parallel exchange buffer (dist.py) in one launch:
 *   tail (2K + 2) fp32 = [hist & 0xffff (K) | hist >> 16 (K) | loss | 1]   -- count words that are exact in fp32 and stay exact
 * when summed over ranks.  hist (K) int64 and loss (1) fp32 are the outputs of vq_forward.
 */
int vq_pack_stats(const int64_t* hist, const float* loss, int K, float* tail, vq_stream_t stream);

/*
 * Index -> embedding lookup in NCHW layout (the decode side: codebook(indices).reshape(B,h,w,D).permute(0,3,1,2),
 * worker/vqganVqvaeWorker.py:459, network/vqTransformer/vqTransformer.py:98).
 *   out_nchw (B, D, HW) fp32
 */
int vq_embed_nchw(const int64_t* idx, const float* E, int64_t B, int64_t HW, int D, int K,
                  float* out_nchw, vq_stream_t stream);

/*
 * Token-stream formats after the tokeniser.
 *
 * vq_index_to_log_onehot replaces index_to_log_onehot(x, num_classes)
 * (network/vq_diffusion/vq_diffusion.py:29-35, network/vqDiffusion/submodule/diffusion_vq_official.py:53-60):
 *   out (B, num_classes, L) fp32 = log(clamp(one_hot(idx (B, L) int64), min = clamp_min)), class axis second;
 *   the reference's clamp_min is 1e-30.  Indices outside [0, num_classes) leave their column at log(clamp_min)
 *   (F.one_hot raises for them: the Python mirror checks before calling).
 *
 * vq_mask_replace replaces the arithmetic of VQTransformer.forward's input corruption
 * (network/vqTransformer/vqTransformer.py:117-141): out (B, L + 1) int64 with out[:, 0] = sos_token and
 *   out[:, 1 + l] = m * indices + (1 - m) * random_indices,  m = (int64) round(mask)   (mask: the fp32 Bernoulli draw).
 *   The random draws themselves stay with the caller (torch's generator), so the stream is the reference's.
 */
int vq_index_to_log_onehot(const int64_t* idx, int64_t B, int64_t L, int num_classes, float clamp_min, float* out,
                           vq_stream_t stream);
/* vq_log_onehot_to_index replaces log_onehot_to_index(log_x) = log_x.argmax(1) (network/vq_diffusion/vq_diffusion.py:37-38):
 *   log_x (B, num_classes, L) fp32 contiguous -> out (B, L) int64, first maximum, a NaN counts as the maximum (torch.argmax). */
int vq_log_onehot_to_index(const float* log_x, int64_t B, int64_t L, int num_classes, int64_t* out, vq_stream_t stream);
int vq_mask_replace(const int64_t* indices, const float* mask, const int64_t* random_indices, int64_t sos_token,
                    int64_t B, int64_t L, int64_t* out, vq_stream_t stream);

/* Number of kernel launches the last call on this thread enqueued (for bench.py's gpu_launches). */
int vq_last_launch_count(void);

/*
 * Timing of the distance-GEMM kernel for the roofline report.  vq_profile_enable(1) makes vq_argmin / vq_forward
 * calls ON THIS THREAD bracket that kernel with CUDA events on the caller's stream (up to 512 calls);
 * vq_profile_collect, called after the caller synchronised, returns the elapsed milliseconds per call and resets.
 */
int vq_profile_enable(int on);
int vq_profile_collect(float* ms_host, int cap, int* n_host);

/* Debug: dump the approximate (fp16 tensor-core) scores e2[k] - 2 z.e for all (n, k); N*K_pad floats. */
int vq_debug_scores(const float* z_nchw, int64_t B, int64_t HW, int D,
                    const void* E_h, const float* e_norm2, const float* cb_scalars, int K,
                    float* scores, void* workspace, size_t workspace_bytes, vq_stream_t stream);

/* Debug: while stamps_dev is non-NULL, vq_argmin / vq_forward calls on this thread run an instrumented GEMM kernel whose
 * CTA 0 records clock64() stamps for its first `tiles` code tiles, 8 int64 per tile: MMA wait start, MMA wait end,
 * MMA issued, epilogue(group 0) woke, released, done, epilogue(group 1) released, done. */
int vq_debug_timeline(long long* stamps_dev, int tiles);

#ifdef __cplusplus
}
#endif
#endif /* VQ_B200_H_ */
