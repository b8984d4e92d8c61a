/* gof_b200.h -- C ABI of libgof_b200.so, the B200 (sm_100a) GOF Gaussian rasterizer.
 *
 * This is the drop-in boundary for the hot path of W-Ted/F3D-Gaus: every entry point
 * below replaces one function of the reference's native layer
 * (RAST = src/gaussian-splatting/submodules/diff-gof-rasterization):
 *
 *   gof_forward          <- CudaRasterizer::Rasterizer::forward   RAST/cuda_rasterizer/rasterizer.h:30-59
 *   gof_forward_batch    <- the per-frame render loops around it  visualize.py:293-306,387-402
 *                           (bound by RasterizeGaussiansCUDA,      RAST/rasterize_points.cu:36-122)
 *   gof_integrate        <- CudaRasterizer::Rasterizer::integrate RAST/cuda_rasterizer/rasterizer.h:93-123
 *   gof_backward         <- CudaRasterizer::Rasterizer::backward  RAST/cuda_rasterizer/rasterizer.h:61-91
 *                           (bound by RasterizeGaussiansBackwardCUDA, RAST/rasterize_points.cu:124-211)
 *   gof_preprocess_backward <- BACKWARD::preprocess             RAST/cuda_rasterizer/backward.cu:957-1033
 *   gof_mark_visible     <- CudaRasterizer::Rasterizer::markVisible RAST/cuda_rasterizer/rasterizer.h:24-29
 *                           (bound by markVisible,                 RAST/rasterize_points.cu:213-232)
 *   gof_state_sizes      <- required<GeometryState/ImageState/BinningState>()
 *                                                                  RAST/cuda_rasterizer/rasterizer_impl.h:80-87
 *   gof_state_get        <- (test accessor) GeometryState/ImageState/BinningState::fromChunk
 *                                                                  RAST/cuda_rasterizer/rasterizer_impl.cu:188-243
 *   gof_render_epilogue  <- the torch post-processing of render_predicted_more_v2_gof
 *                                                                  src/gaussian_renderer/__init__.py:881-909,1043-1053
 *   gof_render_epilogue_backward_batch <- torch.autograd through those same ops
 *   gof_set_frame_sink   <- the .cpu() of rgb / depth / alpha after every frame       visualize.py:304-306,396-398
 *   gof_predictor_head   <- GaussianSplatPredictor_gtunet.forward after the UNet      src/gaussian_predictor.py:954-1008
 *   gof_pack_gather      <- (no counterpart; the multi-GPU exchange step: pack + all-gather over NVLink peer memory)
 *
 * Conventions (same as the reference unless noted):
 *   - plain pointers and sizes only; all array pointers are DEVICE pointers, float32 unless noted;
 *   - "not provided" inputs are NULL (the reference passes data_ptr()==nullptr of an empty tensor);
 *   - viewmatrix/projmatrix are 16 floats read as m[0..15] with p' = (m0 x+m4 y+m8 z+m12, ...)
 *     (RAST/cuda_rasterizer/auxiliary.h:86-115);
 *   - rotations are (r,x,y,z) and are NOT normalised (forward.cu:138,172);
 *   - the three state blobs are opaque, caller-owned byte buffers that must not move between
 *     forward and backward (rasterizer_impl.cu:444-446);
 *   - every function returns 0 on success or a negative GOF_E* code; gof_last_error() gives the
 *     message.  No C++ exception crosses this boundary;
 *   - kernels are launched on the caller's `stream` (the reference uses the legacy default
 *     stream only; passing 0 reproduces that);
 *   - a GofContext belongs to one device and is not thread-safe (it owns the pinned mailbox and the
 *     backward accumulator): use one context per calling thread, as the reference is used under the GIL.
 */
#ifndef GOF_B200_H_
#define GOF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GOF_OK            0
#define GOF_EINVAL       -1   /* bad argument (shape, NULL, unsupported combination) */
#define GOF_ECUDA        -2   /* a CUDA runtime call failed */
#define GOF_ENOMEM       -3   /* a state blob / allocation callback was too small */
#define GOF_EOVERFLOW    -4   /* sync-free mode: num_rendered exceeded the binning capacity */

#define GOF_OUTPUT_CHANNELS 9   /* rgb(3) normal(3) depth alpha distortion: auxiliary.h:21-24 */
#define GOF_MAX_VIEWS      64   /* views per gof_forward_batch call */
#define GOF_SINK_CHANNELS   5   /* frame sink: rgb(3), median depth, alpha -- what visualize.py:304-306 reads back */
#define GOF_SINK_CHW        0   /* sink layout [V,5,H,W] */
#define GOF_SINK_HWC        1   /* sink layout [V,H,W,5] (channels last: 320-byte contiguous runs per tile row) */

typedef struct GofContext GofContext;   /* per-device handle (scratch, pinned mailbox) */
typedef void* gof_stream_t;             /* cudaStream_t */

/* Scalar settings: GaussianRasterizationSettings_GOF (RAST/diff_gof_rasterization/__init__.py:168-182). */
typedef struct GofParams {
	int32_t P;              /* number of Gaussians */
	int32_t D;              /* active SH degree (sh_degree) */
	int32_t M;              /* SH coefficients per channel = sh.size(1), 0 if no SH */
	int32_t W, H;           /* image_width, image_height */
	float tan_fovx, tan_fovy;
	float kernel_size;
	float scale_modifier;
	int32_t prefiltered;
	int32_t debug;          /* !=0: synchronise + check after every stage (auxiliary.h:204-211) */
	int32_t flags;          /* GOF_FLAG_* */
} GofParams;

/* GofParams.flags.
 * GOF_FLAG_EXACT_BLEND: evaluate the per-contributor depth mapping and normal normalisation with
 *   the reference's IEEE double divide / double sqrt / float divides, which makes all nine output
 *   channels bit-identical to the reference's sm_100a build.  Without it those two quantities use
 *   float32 reciprocal arithmetic (<= 4 ulp): rgb, median depth, alpha, T and the contributor
 *   counts stay bit-identical, normals and distortion agree to ~1e-6 (north-star bar: 1e-4). */
#define GOF_FLAG_EXACT_BLEND 1
/* GOF_FLAG_SAVE_CONTRIB (forward): the blend also records, per pixel, WHICH records of its tile list it blended (one
 *   bit per list position, 32 bytes per duplicate in the binning blob).  A backward on that state then walks exactly
 *   those pairs instead of re-running the conic sweep over the whole list and re-evaluating the survivors that did
 *   not contribute.  Set it for forwards whose result will be differentiated (the Python layer does when an input
 *   requires grad); the backward finds out from the state itself and works either way, with identical results. */
#define GOF_FLAG_SAVE_CONTRIB 2

/* Per-Gaussian inputs + camera (argument list of Rasterizer::forward). */
typedef struct GofInputs {
	const float* background;             /* [3] */
	const float* means3D;                /* [P,3] */
	const float* shs;                    /* [P,M,3] or NULL */
	const float* colors_precomp;         /* [P,3]   or NULL */
	const float* opacities;              /* [P] */
	const float* scales;                 /* [P,3]   or NULL (then cov3D_precomp) */
	const float* rotations;              /* [P,4]   or NULL */
	const float* cov3D_precomp;          /* [P,6]   or NULL */
	const float* view2gaussian_precomp;  /* [P,10]  or NULL */
	const float* viewmatrix;             /* [16] */
	const float* projmatrix;             /* [16] */
	const float* campos;                 /* [3] */
} GofInputs;

/* Gradient outputs of Rasterizer::backward, in the order RasterizeGaussiansBackwardCUDA
 * returns them (rasterize_points.cu:210).  All are fully written (no pre-zeroing needed). */
typedef struct GofGrads {
	float* dL_dmeans2D;        /* [P,3] */
	float* dL_dcolors;         /* [P,3] */
	float* dL_dopacity;        /* [P,1] */
	float* dL_dmeans3D;        /* [P,3] */
	float* dL_dcov3D;          /* [P,6]  (identically 0, as in the reference) */
	float* dL_dsh;             /* [P,M,3] or NULL when M==0 */
	float* dL_dscales;         /* [P,3] */
	float* dL_drotations;      /* [P,4] */
	float* dL_dview2gaussian;  /* [P,10] */
} GofGrads;

/* Binning-blob allocation callback (the reference's resizeFunctional, rasterize_points.cu:28-34):
 * called once per forward, after num_rendered is known, with the number of bytes needed;
 * must return a device pointer to at least that many bytes (256-byte aligned) or NULL. */
typedef void* (*GofAllocFn)(void* user, size_t bytes);

const char* gof_last_error(void);
const char* gof_version(void);

int gof_context_create(int device, GofContext** out);
void gof_context_destroy(GofContext* ctx);

/* Optional per-stage timing with CUDA events on the caller's stream (used by bench.py for the
 * roofline line; no reference counterpart).  gof_profile_read synchronises the device and sums the
 * stage durations (ms) of all calls since the previous read:
 *   fwd_ms[5] = preprocess, scan, num_rendered hand-off, binning (duplicate+sort+ranges/gather), blend
 *   bwd_ms[3] = accumulator clear, blend backward, preprocess backward */
int gof_profile_enable(GofContext* ctx, int on);
int gof_profile_read(GofContext* ctx, double* fwd_ms, int64_t* fwd_calls, double* bwd_ms, int64_t* bwd_calls);

/* Byte sizes of the three state blobs.  binning_bytes is for `num_rendered` duplicates
 * (pass an upper bound for the sync-free mode). */
int gof_state_sizes(int32_t P, int32_t W, int32_t H, int64_t num_rendered,
                    size_t* geom_bytes, size_t* img_bytes, size_t* binning_bytes);
/* Same for a batch of V views (gof_forward_batch); num_rendered is the batch total. */
int gof_state_sizes_batch(int32_t P, int32_t W, int32_t H, int32_t V, int64_t num_rendered,
                          size_t* geom_bytes, size_t* img_bytes, size_t* binning_bytes);

/* Forward: preprocess -> tile binning -> per-tile front-to-back GOF blend.
 *   geom, img      : caller-allocated blobs of at least gof_state_sizes() bytes.
 *   binning        : if non-NULL, a caller-allocated blob of `binning_bytes` bytes is used and no
 *                    host synchronisation happens (num_rendered is then read back lazily, see
 *                    gof_num_rendered); GOF_EOVERFLOW is reported by gof_num_rendered if it was
 *                    too small (outputs are then invalid).
 *                    If NULL, `alloc`(`alloc_user`, bytes) supplies the blob, as in the reference
 *                    (rasterizer_impl.cu:336-340).  After the first call of a context the blob is
 *                    requested speculatively (1.25x the previous num_rendered) and all kernels are
 *                    enqueued BEFORE the host waits for R, so the GPU does not idle behind the round
 *                    trip; if the guess was too small the callback is invoked a second time with the
 *                    exact size.  The blob must be passed to gof_backward at the same address and with
 *                    the same byte size (its layout is a function of the two).
 *   out_color      : [9,H,W], radii: [P] int32.  Both fully written.
 *   num_rendered   : host pointer; receives R in the callback mode, -1 in the sync-free mode.
 *   binning_out    : host pointer; receives the binning blob actually used. */
int gof_forward(GofContext* ctx, const GofParams* prm, const GofInputs* in,
                void* geom, size_t geom_bytes, void* img, size_t img_bytes,
                void* binning, size_t binning_bytes, GofAllocFn alloc, void* alloc_user,
                float* out_color, int32_t* radii,
                int32_t* num_rendered, void** binning_out, gof_stream_t stream);

/* Batched forward: V views of ONE Gaussian set in a single pass of the pipeline -- what the render
 * loops of the reference do one frame at a time (`for th in views: for bb in scenes: render(...)`,
 * visualize.py:293-306,387-402).  in->viewmatrix / projmatrix / campos point to V consecutive cameras
 * ([V,16], [V,16], [V,3]); in->background to [3] (bg_stride 0) or [V,3] (bg_stride 3).  Every kernel
 * covers the whole batch (grid.y/z = view), the tile lists of all views are binned together and the
 * num_rendered hand-off happens once.  out_color is [V,9,H,W], radii [V,P], num_rendered a host
 * array [V]; state blobs are sized by gof_state_sizes_batch.  Frame v of the batch is bit-identical
 * to gof_forward with camera v.  view2gaussian_precomp is per view and needs V == 1. */
int gof_forward_batch(GofContext* ctx, const GofParams* prm, const GofInputs* in, int32_t V, int32_t bg_stride,
                      void* geom, size_t geom_bytes, void* img, size_t img_bytes,
                      void* binning, size_t binning_bytes, GofAllocFn alloc, void* alloc_user,
                      float* out_color, int32_t* radii,
                      int32_t* num_rendered, void** binning_out, gof_stream_t stream);

/* Point integration for GOF mesh extraction (Rasterizer::integrate, rasterizer_impl.cu:530-792, bound by
 * IntegrateGaussiansToPointsCUDA, rasterize_points.cu:234-343): renders the Gaussians with five rays per
 * pixel and, for each of the PN query points (points3D [PN,3]) that projects into the image, accumulates
 * alpha along the ray through the point up to the point's depth.
 *   out_color [9,H,W]: rgb, 0,0,0, max depth, alpha, number of query points in the pixel;
 *   out_alpha_integrated [PN] (1 for points outside the view), out_color_integrated [PN,3].
 * The binning blob and the scratch are obtained through `alloc` (one call, after one stream sync). */
int gof_integrate(GofContext* ctx, const GofParams* prm, const GofInputs* in, int32_t PN, const float* points3D,
                  void* geom, size_t geom_bytes, void* img, size_t img_bytes, GofAllocFn alloc, void* alloc_user,
                  float* out_color, int32_t* radii, float* out_alpha_integrated, float* out_color_integrated,
                  int32_t* num_rendered, gof_stream_t stream);

/* Frame sink for the NEXT gof_forward / gof_forward_batch call on `ctx` (one-shot: the call consumes it).
 * The blend then ALSO stores rgb, median depth and alpha of every frame at `sink` (16-byte aligned, V*5*H*W
 * floats; layout GOF_SINK_CHW = [V,5,H,W] or GOF_SINK_HWC = [V,H,W,5]) from inside the kernel, as whole tile rows.
 * `sink` is device memory or pinned host memory (cudaHostAlloc / cudaHostRegister); with pinned host memory the
 * frames arrive on the host as posted PCIe writes while other tiles still blend, i.e. the `.cpu()` of the render
 * loops (visualize.py:304-306, 396-398) costs no time of its own; channels-last gives the longest PCIe bursts.
 * The frames are complete when the work enqueued on the call's stream has finished.  NULL clears a pending sink. */
int gof_set_frame_sink(GofContext* ctx, void* sink, size_t sink_bytes, int32_t layout);

/* Sync-free mode: blocks on `stream` and returns the per-view R ([V]) of the last forward that used
 * `geom` (or GOF_EOVERFLOW if the binning blob was too small for it). */
int gof_num_rendered(GofContext* ctx, const void* geom, int32_t P, int32_t V, gof_stream_t stream, int32_t* num_rendered);

/* Same without blocking: enqueues on `stream` a copy of the batch's mailbox into `host_dst`, 4 + V int32 of PINNED host
 * memory: {R_total, overflow flag, longest tile list, 0, R_view[0..V-1]}.  Valid once the work enqueued on `stream` up
 * to this call has completed (record an event behind it); overflow != 0 means the binning blob was too small and every
 * output of the batch holds NaN.  This is what lets a streaming loop (gaussian_renderer.SceneStreamer) check batch k
 * while batch k+1 is already running. */
#define GOF_MAILBOX_HEAD 4
int gof_num_rendered_async(const void* geom, int32_t P, int32_t V, int32_t* host_dst, gof_stream_t stream);

/* Backward: replays the blend back-to-front and produces the reference's 9 gradient tensors.
 * `binning` / `binning_bytes`: the blob the forward returned and its size in bytes (binningBuffer.numel()); the
 * blob's internal layout is derived from exactly these two values, in forward and backward alike. */
int gof_backward(GofContext* ctx, const GofParams* prm, const GofInputs* in,
                 int32_t num_rendered, const int32_t* radii,
                 const void* geom, const void* binning, size_t binning_bytes, const void* img,
                 const float* dL_dout_color /* [9,H,W] */, const GofGrads* grads,
                 gof_stream_t stream);

/* Backward of a gof_forward_batch call: V views in one pass.  dL_dout_color is [V,9,H,W], radii [V,P];
 * num_rendered is the batch total; cameras / background as in gof_forward_batch.  The gradient outputs have
 * the single-frame shapes and hold the SUM over the V views (what autograd would accumulate if the frames
 * had been rendered one by one). */
int gof_backward_batch(GofContext* ctx, const GofParams* prm, const GofInputs* in, int32_t V, int32_t bg_stride,
                       int64_t num_rendered, const int32_t* radii,
                       const void* geom, const void* binning, size_t binning_bytes, const void* img,
                       const float* dL_dout_color /* [V,9,H,W] */, const GofGrads* grads, gof_stream_t stream);

/* Stage entry: the per-Gaussian backward alone (BACKWARD::preprocess, backward.cu:957-1033):
 * from dL/dview2gaussian [P,10] and dL/dcolor [P,3] (may be NULL) produce dL/dmeans3D, dL/dscales,
 * dL/drotations, dL/dsh (and copy the two inputs through to grads->dL_dview2gaussian/dL_dcolors;
 * dL_dopacity/dL_dmeans2D/dL_dcov3D are zero-filled).  `geom` is the forward's geometry blob (for the
 * SH clamp flags).  This map is ill-conditioned at F3D-Gaus scales (the reference's own outputs move
 * by ~10% run to run under its unordered float atomics), so parity of this stage is tested on
 * identical inputs rather than end to end. */
int gof_preprocess_backward(GofContext* ctx, const GofParams* prm, const GofInputs* in, const int32_t* radii,
                            const void* geom, const float* dL_dview2gaussian_in, const float* dL_dcolors_in,
                            const GofGrads* grads, gof_stream_t stream);

int gof_mark_visible(int32_t P, const float* means3D, const float* viewmatrix,
                     const float* projmatrix, uint8_t* present /* [P] bool */, gof_stream_t stream);

/* Scalar settings of the predictor output head (cfg['model'] of the reference + the call's squre_clip). */
typedef struct GofHeadParams {
	int32_t BV;             /* images in the call: batch * views (x.reshape(B*N_views, ...), gaussian_predictor.py:913) */
	int32_t H, W;           /* training_resolution */
	int32_t C;              /* channels of the network output: [3 +] 1 + 3 + 4 + 3 [+ 9] (get_splits_and_inits, :683-728) */
	int32_t with_offset;    /* network_with_offset: the first 3 channels are the xyz offset */
	int32_t sh_degree;      /* max_sh_degree: 0 or 1 (the reference asserts 1, :993) */
	int32_t isotropic;      /* cfg['model']['isotropic']: scaling channel 0 replicated (:951-952) */
	float squre_clip;       /* x/y clamp of the world position, active when < 10 (:968-970) */
} GofHeadParams;

/* Predictor output head: everything GaussianSplatPredictor_gtunet.forward does after the UNet
 * (src/gaussian_predictor.py:954-1008) in one kernel: channel split, ray * depth + offset, view_to_world (row-vector
 * convention) and homogeneous divide, squre_clip, sigmoid / exp / normalize, rotation and degree-1 SH to the world
 * frame, NCHW -> point-list layout.  All pointers are device memory, float32, contiguous:
 *   net [BV,C,H,W]; depth [BV,1,H,W]; const_offset [BV,1,H,W] or NULL (origin_distances, :915-917,:872);
 *   ray_x [W], ray_y [H]: the x / y rows of the module's ray_dirs buffer (init_ray_dirs, :657-681);
 *   view_to_world [BV,16]; quat [BV,4] (source_cv2wT_quat); sh_transform [BV,9] = sh_to_v @ V2W[:3,:3] @ v_to_sh
 *   (:824-832; NULL = derive it in the kernel from the module's constant v_to_sh matrix, :649-655);
 *   outputs xyz [BV*N,3], opacity [BV*N,1], scaling [BV*N,3], rotation [BV*N,4], features_dc [BV*N,1,3],
 *   features_rest [BV*N,3,3] (NULL when sh_degree == 0) -- i.e. [B, V*N, .] after multi_view_union (:796-800). */
int gof_predictor_head(const GofHeadParams* prm, const float* net, const float* depth, const float* const_offset,
                       const float* ray_x, const float* ray_y, const float* view_to_world, const float* quat,
                       const float* sh_transform, float* xyz, float* opacity, float* scaling, float* rotation,
                       float* features_dc, float* features_rest, gof_stream_t stream);

/* Fused epilogue of render_predicted_more_v2_gof: from out_color[9,H,W] and the camera produce
 *   normal_world[3,H,W] = R_c2w * normalize(out_color[3:6])   and
 *   depth_normal[3,H,W] = normalize(cross(dP/dy, dP/dx)) of the back-projected median depth
 *                         (border pixels 0), as depth_to_normal() does.
 * viewmatrix is the same 16 floats passed to gof_forward; fovx/fovy in radians. */
int gof_render_epilogue(const float* out_color, const float* viewmatrix, int32_t W, int32_t H,
                        float fovx, float fovy, float* normal_world, float* depth_normal,
                        gof_stream_t stream);
/* Same for V frames: out_color [V,9,H,W], viewmatrix [V,16], outputs [V,3,H,W] (either may be NULL). */
int gof_render_epilogue_batch(const float* out_color, const float* viewmatrix, int32_t V, int32_t W, int32_t H,
                              float fovx, float fovy, float* normal_world, float* depth_normal,
                              gof_stream_t stream);

/* Backward of the fused epilogue (what autograd does through the reference's torch ops, in one kernel and without
 * atomics): from dL/dnormal_world [V,3,H,W] and dL/ddepth_normal [V,3,H,W] (either may be NULL = zero) produce
 * dL/dout_color [V,9,H,W]: channels 3..5 (through the normalisation and rotation) and 6 (median depth, through the
 * finite differences of the back-projected points); the other channels are written as zeros. */
int gof_render_epilogue_backward_batch(const float* out_color, const float* viewmatrix, int32_t V, int32_t W, int32_t H,
                                       float fovx, float fovy, const float* dL_dnormal_world, const float* dL_ddepth_normal,
                                       float* dL_dout_color, gof_stream_t stream);

/* Fused pack + all-gather of rendered frames over NVLink peer memory (the scene-sharded runner's one exchange
 * step; no reference counterpart -- the reference is single-GPU).  raster: this rank's [frames,9,H*W] output;
 * peer_ptrs_dev: DEVICE array of `world` pointers to every rank's gather buffer [total_frames,5,H*W] (a symmetric
 * allocation, e.g. torch symmetric memory); multicast_ptr: the allocation's NVSwitch multicast address or NULL;
 * dst_frame0: index of this rank's first frame in the gathered order.  Writes rgb, median depth, alpha of every
 * local frame into all ranks' buffers; the caller issues the symmetric-memory barrier that publishes them. */
int gof_pack_gather(const float* raster, int32_t frames, int64_t pixels, const int64_t* peer_ptrs_dev, int32_t world,
                    void* multicast_ptr, int64_t dst_frame0, gof_stream_t stream);

/* Test accessor: the per-(view, Gaussian) gradient accumulators of the LAST gof_backward[_batch] call on `ctx`,
 * [V,P,20] float32 = {dL/dview2gaussian[10], dL/dcolor[3], dL/dopacity, dL/dmean2D (x, y, |x|+|y|), pad[3]} -- what
 * the backward blend handed to the per-Gaussian backward, before the views are summed.  Copies them to `dst` (device
 * pointer) and returns the byte size, or a negative error; dst == NULL only returns the size. */
int64_t gof_backward_accumulators(GofContext* ctx, void* dst, int64_t dst_bytes, gof_stream_t stream);

/* Test accessor: copy one named array of the opaque state into dst (device pointer).
 * Names: depths[P] f32, means2D[P,2] f32, conic_opacity[P,4] f32, view2gaussian[P,10] f32,
 * rgb[P,3] f32, clamped[P,3] u8, tiles_touched[P] u32, point_offsets[P] u32,
 * final_T[4,H,W] f32, n_contrib[2,H,W] u32, ranges[T,2] u32, point_list[R] u32,
 * point_list_keys[R] u64.  Returns the byte size (>=0) or a negative error. */
int64_t gof_state_get(const char* name, int32_t P, int32_t W, int32_t H, int64_t num_rendered,
                      const void* geom, const void* binning, size_t binning_bytes, const void* img,
                      void* dst, int64_t dst_bytes, gof_stream_t stream);
/* Same on the state of a V-view batch: per-Gaussian arrays are [V,P,...], per-pixel [V,...,H,W],
 * ranges [V*T,2] (offsets into the batch's list), point_list / point_list_keys [R_total]. */
int64_t gof_state_get_batch(const char* name, int32_t P, int32_t W, int32_t H, int32_t V, int64_t num_rendered,
                            const void* geom, const void* binning, size_t binning_bytes, const void* img,
                            void* dst, int64_t dst_bytes, gof_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GOF_B200_H_ */
