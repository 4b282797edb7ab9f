/*
 * dgtta.h — C ABI of libdgtta_sm100.so: the B200 (sm_100a) implementation of DG-TTA's
 * input-transform hot path (MIND-SSC descriptor, GIN augmentation, affine trilinear sampling).
 *
 * Conventions:
 *   - plain C symbols, raw pointers and sizes, no torch / C++ types;
 *   - every tensor is contiguous NCDHW float32; "dev" pointers are CUDA device pointers of the
 *     current device, "host" pointers are ordinary host memory read before the call returns;
 *   - the library never allocates device memory: outputs and workspaces are caller-owned
 *     (query *_workspace_bytes first); it only enqueues kernels on `stream` and never syncs;
 *   - return value 0 = ok, >0 = cudaError_t from a launch, <0 = DGTTA_E* argument error;
 *     dgtta_last_error() gives the message of the last failure on the calling thread.
 *
 * Each entry point names the reference interface (multimodallearning/DG-TTA, file:line) it replaces.
 */
#ifndef DGTTA_H_
#define DGTTA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st *dgtta_stream_t; /* == cudaStream_t */

#define DGTTA_ABI_VERSION 1

#define DGTTA_EINVAL (-1)     /* bad shape / parameter */
#define DGTTA_ENULL (-2)      /* required pointer is NULL */
#define DGTTA_EWORKSPACE (-3) /* workspace too small */
#define DGTTA_EUNSUPPORTED (-4)

/* MIND noise source (dg_tta/mind.py:150-152) */
#define DGTTA_NOISE_NONE 0   /* randn_weighting ignored: E = I(p+s1) - I(p+s2)            */
#define DGTTA_NOISE_TENSOR 1 /* noise_dev holds the [B,12,D,H,W] normal field              */
#define DGTTA_NOISE_PHILOX 2 /* regenerate torch's CUDA randn stream on the device from (seed, offset) */

/* sampler modes (torch.nn.functional.grid_sample arguments used at tta.py:549,573; torch_utils.py:59,71) */
#define DGTTA_INTERP_TRILINEAR 0
#define DGTTA_INTERP_NEAREST 1
#define DGTTA_PAD_ZEROS 0
#define DGTTA_PAD_BORDER 1

int dgtta_abi_version(void);
const char *dgtta_last_error(void);
/* kernels launched by this library since the process started (diagnostics; bench.py's gpu_launches) */
uint64_t dgtta_launch_count(void);
/* Loads every kernel of the library into the current CUDA context (CUDA otherwise loads each kernel lazily at its
 * first launch, a few ms apiece).  Call once per device after the context exists; 0 = ok. */
int dgtta_preload_kernels(void);

/* ---------------------------------------------------------------------------------------------
 * MIND-SSC descriptor.  Replaces MIND3D.forward (dg_tta/mind.py:142-164) including the shift
 * kernels built in MIND3D.__init__ (:98-140), smooth/filter1D (:5-43) and therefore mind_hook
 * (:167-168).
 *   img_dev   [B,1,D,H,W]        out_dev [B,12,D,H,W]
 *   in_scale_dev: NULL, or [B,2] floats (a_b, c_b): the kernel reads I = (img * a_b) * c_b — the
 *                 deferred Frobenius re-normalisation of GIN (gin.py:228) when GIN feeds MIND.
 *   taps_host [ntaps] Gaussian taps (mind.py:27-37), ntaps odd, 1..9
 *   Output range: [0, 1].  The epilogue evaluates exp(-m/v) as ex2.approx.ftz(m * (-log2 e * rcp.approx.ftz(v))): results
 *                 below 2^-126 flush to exactly 0 where the reference returns a denormal (both are < 1.2e-38; inside the
 *                 1e-5 tolerance), and every voxel has at least one channel equal to 1.
 *   noise_mode/noise_dev/philox_*: see DGTTA_NOISE_*.  For PHILOX, noise_dev is a caller-owned scratch
 *                 of B*12*D*H*W floats that the call fills with the field torch.randn_like(edge_selection)
 *                 (mind.py:150) draws for the device generator state (seed, offset) before the draw
 *                 (dgtta_philox_normal_fill), and the caller advances the generator by
 *                 dgtta_mind_philox_offset_increment() afterwards.
 *   workspace: dgtta_mind_workspace_bytes(B,D,H,W) bytes, 16-byte aligned.
 * Two launches: a speculative fused stencil pass that also reduces the statistics of the
 * per-voxel variance, then a fix-up pass that recomputes only the tiles in which the global
 * clamp of mind.py:158-160 is active (none for ordinary images).
 * ------------------------------------------------------------------------------------------- */
size_t dgtta_mind_workspace_bytes(int B, int D, int H, int W);
int dgtta_mind_ssc_fwd(const float *img_dev, float *out_dev, const float *in_scale_dev, int B, int D, int H,
                       int W, int delta, const float *taps_host, int ntaps, float randn_weighting,
                       int noise_mode, const float *noise_dev, uint64_t philox_seed, uint64_t philox_offset,
                       void *workspace_dev, size_t workspace_bytes, dgtta_stream_t stream);
/* how far torch.randn_like(edge_selection) advances the CUDA generator offset for this shape
 * (ATen/native/cuda/DistributionTemplates.h calc_execution_policy); sm_count/max_threads_per_sm
 * are the device properties torch uses. */
uint64_t dgtta_mind_philox_offset_increment(int B, int D, int H, int W, int sm_count, int max_threads_per_sm);

/* Host-side note (dg_tta_b200/mind.py): where torch's own launch split or a stream capture makes the stream
 * irreproducible from host-known (seed, offset) — >= 2^31 elements, offset % 4 != 0, capture without the graph-safe
 * entry below — the Python layer draws the field with torch.randn itself (the reference's own draw: identical values)
 * and passes it as DGTTA_NOISE_TENSOR.  That is still this library's MIND kernel; only the noise generator is torch's.
 *
 * The N(0,1) field torch.randn(numel elements, device="cuda") writes for generator state (seed, offset):
 * Philox4x32-10 + Box-Muller in torch's element order (ATen/native/cuda/DistributionTemplates.h), the
 * draw behind mind.py:150.  offset must be a multiple of 4 and numel < 2^31 (torch's single-launch case);
 * sm_count / max_threads_per_sm are the device properties torch derives its grid from.
 * dgtta_philox_normal_offset_increment: how far that draw advances the generator offset. */
int dgtta_philox_normal_fill(float *out_dev, uint64_t numel, uint64_t philox_seed, uint64_t philox_offset,
                             int sm_count, int max_threads_per_sm, dgtta_stream_t stream);
uint64_t dgtta_philox_normal_offset_increment(uint64_t numel, int sm_count, int max_threads_per_sm);
/* CUDA-graph form of the same fill: the generator state {seed, offset} (two uint64) is read from device memory when
 * the kernel runs — as torch feeds Philox under capture — so one captured launch draws a fresh field per replay.  The
 * caller rewrites the two words before each replay (a captured memcpy from pinned memory) and advances the
 * generator by dgtta_philox_normal_offset_increment.  Same stream of values as dgtta_philox_normal_fill. */
int dgtta_philox_normal_fill_graphsafe(float *out_dev, uint64_t numel, const uint64_t *seed_offset_dev, int sm_count,
                                       int max_threads_per_sm, dgtta_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * GIN augmentation.  Replaces GINGroupConv.forward (dg_tta/gin.py:168-230) and the per-layer
 * GradlessGCReplayNonlinBlock.forward (gin.py:59-122) for 5-D input; the random draws stay with
 * the caller (reference order: alphas on the device generator, then per layer randint/randn/randn
 * on the CPU generator).
 *   x_dev [B,Cin,D,H,W]   out_dev same shape
 *   params_host: for layer L = 0..n_layer-1: ker_L [cout_L*B, cin_L, k,k,k] then shift_L [cout_L*B]
 *   ksizes_host [n_layer] each 1 or 3;   alphas_dev [B]
 *   scale_out_dev: NULL -> out = mixed * (1/(||mixed_b||+1e-5)) * ||x_b|| (gin.py:228);
 *                  non-NULL -> out = mixed (unscaled) and scale_out_dev[b] = {1/(||mixed_b||+1e-5), ||x_b||}
 *                  for a consumer that applies it on load (dgtta_mind_ssc_fwd in_scale_dev).
 * ------------------------------------------------------------------------------------------- */
size_t dgtta_gin_workspace_bytes(int B, int D, int H, int W, int in_channels, int n_layer, int interm_channels);
int dgtta_gin_fwd(const float *x_dev, float *out_dev, const float *params_host, const int *ksizes_host,
                  const float *alphas_dev, int B, int D, int H, int W, int in_channels, int n_layer,
                  int interm_channels, float *scale_out_dev, void *workspace_dev, size_t workspace_bytes,
                  dgtta_stream_t stream);

/* One GIN layer on its own.  Replaces GradlessGCReplayNonlinBlock.forward (dg_tta/gin.py:59-122, 3-D
 * branch): grouped conv (groups=B, zero padding k//2) + shift + optional leaky_relu(0.01).
 *   x_dev [B,cin,D,H,W] -> out_dev [B,cout,D,H,W]; ker_host [cout*B,cin,k,k,k]; shift_host [cout*B]
 *   workspace: (cout*B*cin*k^3 + cout*B) * 4 bytes. */
int dgtta_gin_layer_fwd(const float *x_dev, float *out_dev, const float *ker_host, const float *shift_host, int B,
                        int cin, int cout, int k, int D, int H, int W, int use_act, void *workspace_dev,
                        size_t workspace_bytes, dgtta_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Affine view warp.  Replaces the F.affine_grid + F.grid_sample pairs at dg_tta/tta/tta.py:
 * 523-532 + 548-551 (image, border), :571-575 (prediction, zeros, differentiable w.r.t. input)
 * and dg_tta/tta/torch_utils.py:55-73 (patch crop; nearest for labels), align_corners=False.
 * No grid tensor exists: coordinates come from theta in-kernel.
 *   in_dev [B,C,Di,Hi,Wi]  theta_dev [B,3,4]  out_dev [B,C,Do,Ho,Wo]
 * bwd_input: grad_in_dev [B,C,Di,Hi,Wi] is overwritten with the adjoint of the trilinear forward.  Default: scatter with
 * red.global.add.f32 like torch's grid_sample backward (summation order, hence the last bits, vary from run to run).
 * Environment DGTTA_SAMPLE_BWD_DETERMINISTIC=1 (zeros padding): gather over the source voxels instead — every element
 * written once, bit-identical from run to run, ~2.4x the scatter's time; affines that magnify more than ~2.5x fall
 * back to the scatter (decided on the device).
 * ------------------------------------------------------------------------------------------- */
int dgtta_affine_sample_fwd(const float *in_dev, const float *theta_dev, float *out_dev, int B, int C, int Di,
                            int Hi, int Wi, int Do, int Ho, int Wo, int interp, int padding,
                            dgtta_stream_t stream);
int dgtta_affine_sample_bwd_input(const float *grad_out_dev, const float *theta_dev, float *grad_in_dev, int B,
                                  int C, int Di, int Hi, int Wi, int Do, int Ho, int Wo, int padding,
                                  dgtta_stream_t stream);
/* Label crop.  Replaces the nearest-mode F.grid_sample of the one-hot label channels plus get_argmaxed_segs
 * (dg_tta/tta/torch_utils.py:71-73, 79-82): out[b,0,p] = 0 (background) where the labels sampled at p sum to < 1 or
 * p maps outside the volume, else 1 + argmax_l onehot[b,l,nearest(p)] (lowest index on ties, like torch.argmax).
 *   onehot_dev [B,L,Di,Hi,Wi] float32   theta_dev [B,3,4]   out_dev [B,1,Do,Ho,Wo] int64 */
int dgtta_affine_label_argmax(const float *onehot_dev, const float *theta_dev, long long *out_dev, int B, int L, int Di,
                              int Hi, int Wi, int Do, int Ho, int Wo, dgtta_stream_t stream);

/* Integer label path of the same crop.  Nearest sampling selects one source voxel and
 * get_argmaxed_segs is a per-voxel function, so the crop commutes with it: build the label map of a volume once
 * (dgtta_label_map_from_onehot: map[b,v] = 0 where the L channels sum to < 1, else 1 + argmax_l, the rule above) and
 * crop from the 2 B/voxel map (dgtta_affine_label_gather; 0 outside the volume) instead of reading the L x 4 B/voxel
 * one-hot volume (nnunet_utils.py:191-195 builds it with L = 104) on every call.
 *   onehot_dev [B,L,V] float32 -> map_dev [B,V] int16 (L <= 32766);   map_dev [B,Di,Hi,Wi] -> out_dev [B,1,Do,Ho,Wo] int64 */
int dgtta_label_map_from_onehot(const float *onehot_dev, short *map_dev, int B, int L, long long V, dgtta_stream_t stream);
int dgtta_affine_label_gather(const short *map_dev, const float *theta_dev, long long *out_dev, int B, int Di, int Hi, int Wi,
                              int Do, int Ho, int Wo, dgtta_stream_t stream);

/* Image crop of get_batch in one gather.  Replaces `grid_sample(vol - vol.min(), zeros) + vol.min()`
 * (dg_tta/tta/torch_utils.py:58-62): out = sum_k w_k (in[k] - shift[b]) + shift[b] over the in-bounds corners.
 * dgtta_volume_min computes shift = min(in) (NaN-propagating like torch.min) into out_dev[0];
 * workspace: dgtta_volume_min_workspace_bytes() bytes. */
int dgtta_affine_crop_shifted_fwd(const float *in_dev, const float *theta_dev, const float *shift_dev, float *out_dev, int B,
                                  int C, int Di, int Hi, int Wi, int Do, int Ho, int Wo, dgtta_stream_t stream);
size_t dgtta_volume_min_workspace_bytes(void);
int dgtta_volume_min(const float *in_dev, long long numel, float *out_dev, void *workspace_dev, size_t workspace_bytes,
                     dgtta_stream_t stream);


/* ---------------------------------------------------------------------------------------------
 * Consistency-loss reductions of the TTA step.  Replace the elementwise chain of dg_tta/tta/tta.py:263-268
 * (common-content mask, channel softmax of both branches) and the per-(sample, class) sums inside
 * soft_dice_loss (dg_tta/tta/torch_utils.py:94-95):
 *     sums[b,c,0] = sum_v 2 sm_a sm_b        sums[b,c,1] = sum_v (sm_a + sm_b)^2
 * with sm_x = softmax_c(target_x) * [sum_c target_a > 0][sum_c target_b > 0].  The caller finishes
 * dice = (sums0/V) / (0.5 sums1/V) and loss = 1 - mean(dice[:, 1:]) on the [B,C] scalars.
 *   target_a_dev, target_b_dev [B,C,V] float32 (V = D*H*W)   sums_dev [B,C,2] float64 (overwritten)
 * bwd: grad_a_dev [B,C,V] = d loss / d target_a for grad_sums_dev [B,C,2] float32 = d loss / d sums
 * (the mask is piecewise constant: no gradient through it, as in torch).  C <= 128. */
int dgtta_consistency_sums_fwd(const float *target_a_dev, const float *target_b_dev, double *sums_dev, int B, int C,
                               long long V, dgtta_stream_t stream);
int dgtta_consistency_sums_bwd(const float *target_a_dev, const float *target_b_dev, const float *grad_sums_dev,
                               float *grad_a_dev, int B, int C, long long V, dgtta_stream_t stream);

/* The same reductions with the inverse warps of the two branches fused in (dg_tta/tta/tta.py:571-575 feeding :263-268):
 *     target_x = grid_sample(logits_x, affine_grid(theta_x, size), zeros padding, align_corners=False),   x in {a, b}
 * is evaluated per output voxel inside the reduction, so the warped logits are never written or re-read — forward or
 * backward.  logits_*_dev [B,C,D,H,W], theta_*_dev [B,3,4] (the inverse affines R^-1), C <= 16 (otherwise warp with
 * dgtta_affine_sample_fwd and call dgtta_consistency_sums_*).  bwd: grad_logits_a_dev [B,C,D,H,W] is overwritten with
 * d loss / d logits_a (adjoint of branch a's trilinear gather applied to d loss / d target_a; float atomics). */
int dgtta_consistency_warp_sums_fwd(const float *logits_a_dev, const float *logits_b_dev, const float *theta_a_dev,
                                    const float *theta_b_dev, double *sums_dev, int B, int C, int D, int H, int W,
                                    dgtta_stream_t stream);
int dgtta_consistency_warp_sums_bwd(const float *logits_a_dev, const float *logits_b_dev, const float *theta_a_dev,
                                    const float *theta_b_dev, const float *grad_sums_dev, float *grad_logits_a_dev, int B, int C,
                                    int D, int H, int W, dgtta_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * MultiRes low-resolution simulation.  Replaces the two skimage.transform.resize(x, shape, order, mode='edge',
 * anti_aliasing=False) calls per channel of augment_discrete_linear_downsampling_scipy
 * (dg_tta/pretraining/discrete_downsampling.py:29-33; installed by nnUNetTrainer_GIN_MIND_MultiRes.py:57-69 with
 * order_downsample = 0, order_upsample = 3).  skimage delegates to scipy.ndimage.zoom(..., mode='nearest',
 * grid_mode=True) and clips to the input's range; orders 0 (nearest), 1 (linear) and 3 (cubic B-spline with the
 * 12-sample edge pre-padding and prefilter of scipy) are built.
 *   in_dev [N,Di,Hi,Wi] -> out_dev [N,Do,Ho,Wo]; the N volumes share the geometry (channels of one sample).
 *   workspace: dgtta_resize_edge_workspace_bytes(...) bytes, 256-byte aligned. */
size_t dgtta_resize_edge_workspace_bytes(int N, int Di, int Hi, int Wi, int Do, int Ho, int Wo, int order);
int dgtta_resize_edge(const float *in_dev, float *out_dev, int N, int Di, int Hi, int Wi, int Do, int Ho, int Wo, int order,
                      void *workspace_dev, size_t workspace_bytes, dgtta_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DGTTA_H_ */
