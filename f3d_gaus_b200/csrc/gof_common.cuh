// gof_common.cuh -- shared declarations of libgof_b200 (sm_100a only).
#pragma once
#include <cstdint>
#include <cstddef>
#include <cuda_runtime.h>
#include "../../include/gof_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libgof_b200 is written for sm_100a (B200) only"
#endif

namespace gof {

// Behaviour-defining constants of the reference (auxiliary.h:18-37, config.h:15-17).
constexpr int TILE_X = 16;
constexpr int TILE_Y = 16;
constexpr int TILE_PIX = TILE_X * TILE_Y;
constexpr int OUT_CH = GOF_OUTPUT_CHANNELS;
constexpr int CH_DEPTH = 6, CH_ALPHA = 7, CH_DIST = 8;
constexpr int SINK_CH = GOF_SINK_CHANNELS;      // frame sink: rgb, median depth, alpha

// Per-(view, Gaussian) "geometry record" written by the preprocess, 64 B:
//  [0..5] Sigma_v (xx,xy,xz,yy,yz,zz)  [6..8] B  [9] C      (view2gaussian, forward.cu:268-277)
//  [10] tau: conservative reject threshold on the ray-minimum value (blend_math.cuh)
//  [11] w = opacity*coef (conic_opacity.w)  [12..14] rgb  [15] Gaussian index (int bits)
constexpr int REC_FLOATS = 16;
constexpr int REC_TAU = 10, REC_W = 11, REC_RGB = 12, REC_ID = 15;

// Per-duplicate "slab record", 80 B (five float4), written in tile order by the tile sort so that a
// tile's sorted Gaussians are one contiguous run that TMA bulk copies can stream.  80 B = 5 x 16 B
// also makes the lane-private record reads of the blend's pass 2 spread over all eight 16-byte
// bank groups of shared memory (5 j mod 8 is a permutation).
//  float4 #0  c0  c1  c2  c3    tile-local conic pre-test  g(x,y) = c0 + x(c1 + c3 x + c4 y) + y(c2 + c5 y)
//  float4 #1  c4  c5  w   Sxx
//  float4 #2  Sxy Sxz Syy Syz
//  float4 #3  Szz Bx  By  Bz
//  float4 #4  C   r   g   b
// (the Gaussian index of a slab entry is point_list[entry]; only the backward needs it)
constexpr int SLAB_FLOATS = 20;
constexpr int SLAB_BYTES = 80;

constexpr size_t ALIGN = 256;
__host__ __device__ inline size_t align_up(size_t x, size_t a = ALIGN) { return (x + a - 1) / a * a; }

constexpr int MAILBOX_HEAD = GOF_MAILBOX_HEAD;   // mailbox ints: {R_total, overflow, max tile count, contributor masks saved, R_view[0..V-1]}

// ---- opaque state layouts (our own; the reference's are rasterizer_impl.cu:188-243) -------
// All per-Gaussian arrays are [V, P] (view-major); V = 1 for the single-frame entry points.
struct GeomState {
	float* depths;          // [V,P]
	float2* means2D;        // [V,P]
	float4* conic_opacity;  // [V,P]
	float* rec;             // [V,P,16]
	uint32_t* tiles_touched;// [V,P]
	ushort4* rect;          // [V,P]  tile rectangle (x0,y0,x1,y1), valid where tiles_touched > 0
	uint8_t* clamped;       // [V,P,3]
	int32_t* mailbox;       // [MAILBOX_HEAD + V]
	size_t total;
	static GeomState carve(char* base, size_t P, size_t V);
};
struct ImgState {
	float* final_T;         // [V,4,N]  T, dist1, dist2, distortion_raw  (forward.cu:591-594)
	uint32_t* n_contrib;    // [V,2,N]  last_contributor, max_contributor (forward.cu:596-597)
	uint2* ranges;          // [V*T]    (0,0) for untouched tiles, offsets into the whole batch's list
	uint32_t* tile_counts;  // [V*T]
	uint32_t* tile_cursor;  // [V*T]
	uint32_t* tile_order;   // [V*T]    tiles by decreasing list length (launch order of the blend CTAs)
	size_t total;
	static ImgState carve(char* base, size_t N, size_t T, size_t V);
};
struct BinState {
	uint64_t* entries;         // [R]  (depth bits << 32) | Gaussian index, bucketed by tile, then sorted in place
	uint32_t* point_list;      // [R]  sorted Gaussian indices (the reference's point_list)
	float* slab;               // [R,20] tile-ordered slab records
	uint8_t* block_mask;       // [R]  bit b: the record can pass the conic test in 8x4 pixel block b of its tile
	float* bwd_rec;            // [R,8]  tile-ordered backward record {mean2D.xy, conic.xyz, Gaussian id, -, -}: what the backward
	                           //  blend needs per pair besides the slab, so that it streams instead of gathering (training forwards)
	uint32_t* contrib;         // [slots][4][256]  per pixel: which records of the tile list it BLENDED (training forwards
	                           //  only, GOF_FLAG_SAVE_CONTRIB); slot of chunk c of tile gt = (ranges[gt].x >> 7) + gt + c
	size_t total;
	static size_t contrib_slots(size_t R, size_t VT) { return (R >> 7) + VT + 1; }
	static BinState carve(char* base, size_t R, size_t VT);
};
constexpr int CONTRIB_SLOT_WORDS = 4 * TILE_PIX;   // 4 KB per 128-record chunk of a tile
constexpr int BWD_REC_FLOATS = 8;
constexpr int BWD_REC_BYTES = 32;

// ---- error plumbing ---------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define GOF_CUDA_CHECK(expr)                                                             \
	do {                                                                                 \
		cudaError_t _e = (expr);                                                         \
		if (_e != cudaSuccess) {                                                         \
			gof::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
			return GOF_ECUDA;                                                            \
		}                                                                                \
	} while (0)

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// The stages of a frame are a chain of dependent kernels on one stream.  Each stage is launched with the
// programmatic-stream-serialization attribute: its CTAs may be scheduled as soon as every CTA of the previous
// stage has executed pdl_trigger() (first thing it does) and SM resources free up, and they block in pdl_wait()
// (first thing THEY do, before touching any global memory) until the previous stage has completed and flushed.
// What overlaps is the launch latency and CTA ramp-up of stage k+1 with the tail of stage k -- 2-4 us per edge,
// which matters for the single-frame call (five ~5-25 us stages in front of the blend).  The blend itself is launched
// WITHOUT it: its CTAs would be placed while the sort's still occupy the SMs, and for a single frame (256 CTAs on 148
// SMs) that placement is less even than a launch into the drained GPU (measured: 5% slower per frame).
// GOF_PDL_MASK selects the edges (PdlEdge bits), GOF_NO_PDL=1 disables all.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
enum PdlEdge { PDL_SCAN = 1, PDL_SCATTER = 2, PDL_SORT = 4, PDL_BLEND = 8, PDL_PRE_BWD = 16 };
int pdl_mask();                 // edges on which PDL is used: GOF_PDL_MASK (default: every edge but the blend's), 0 with GOF_NO_PDL=1
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chained(int edge, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = grid;
	cfg.blockDim = block;
	cfg.dynamicSmemBytes = smem;
	cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = (pdl_mask() & edge) ? 1 : 0;
	return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Camera-dependent scalars shared by the stages.
struct Frame {
	int P, V, W, H;
	dim3 grid;          // tiles in x, y
	int T;              // tiles per view
	float focal_x, focal_y;
};

// ---- stage launchers (one .cu each) ---------------------------------------------------------
int launch_preprocess(const GofParams& prm, const GofInputs& in, const Frame& f, const GeomState& g,
                      const ImgState& im, int32_t* radii, cudaStream_t s);
// save_contrib: recorded in mailbox[3] -- tells the backward that the forward left per-pixel contributor masks
int launch_tile_scan(const Frame& f, const GeomState& g, const ImgState& im, int64_t capacity, cudaStream_t s, int save_contrib = 0,
                     int32_t* host_mail = nullptr, int32_t host_seq = 0);   // host_mail: mapped pinned mailbox + sequence word
// ray_pad: 0 for the blend (rays through pixel centres), 0.5 for point integration (conic.cuh)
// for_backward: also write BinState::bwd_rec (a training forward, GOF_FLAG_SAVE_CONTRIB)
int launch_binning(const Frame& f, const GeomState& g, const ImgState& im, const BinState& b, int64_t capacity,
                   cudaStream_t s, float ray_pad = 0.0f, int for_backward = 0);
// point integration (integrate.cu)
struct IntegrateScratch {
	uint32_t* used_mask;      // [T][words][256]  per-pixel contributed bits over the tile's list
	int words;                // 32-bit words per pixel = 4 * ceil(max tile count / 128)
	float2* points2D;         // [PN]
	float* point_depths;      // [PN]
	uint32_t* point_tile;     // [PN]  tile of the point, 0xffffffff if culled
	uint32_t* point_counts;   // [T]
	uint32_t* point_cursor;   // [T]
	uint2* point_ranges;      // [T]
	uint32_t* point_list;     // [PN]  bucketed by tile
	int32_t* point_mailbox;   // [MAILBOX_HEAD + 1]
	size_t total;
	static IntegrateScratch carve(char* base, size_t PN, size_t T, size_t max_count);
};
int launch_integrate(const GofParams& prm, const GofInputs& in, const Frame& f, const GeomState& g, const ImgState& im,
                     const BinState& b, const IntegrateScratch& sc, int PN, const float* points3D, float* out_color,
                     float* out_alpha_integrated, float* out_color_integrated, cudaStream_t s);
int launch_tile_scan_raw(int T, const uint32_t* counts, uint2* ranges, uint32_t* cursor, int32_t* mailbox, cudaStream_t s);
int launch_render_fwd(const GofParams& prm, const Frame& f, const GeomState& g, const ImgState& im, const BinState& b,
                      const float* background, int bg_stride, float* out_color, float* sink, int sink_hwc, cudaStream_t s);
int launch_predictor_head(const GofHeadParams& prm, const float* net, const float* depth, const float* const_offset,
                          const float* ray_x, const float* ray_y, const float* view_to_world, const float* quat,
                          const float* sh_transform, float* xyz, float* opacity, float* scaling, float* rotation,
                          float* features_dc, float* features_rest, cudaStream_t s);
int launch_render_bwd(const GofParams& prm, const Frame& f, const GeomState& g, const ImgState& im,
                      const BinState& b, const float* background, int bg_stride, const float* dL_dpix, float* gacc,
                      cudaStream_t s);
int launch_preprocess_bwd(const GofParams& prm, const GofInputs& in, int V, const GeomState& g,
                          const int32_t* radii, const float* gacc, const GofGrads& grads,
                          cudaStream_t s);
// test accessors that need kernels
int launch_extract(const char* what, const Frame& f, const GeomState& g, const ImgState& im, const BinState& b,
                   int64_t R, void* dst, cudaStream_t s);

// Per-Gaussian packed gradient accumulator written by the backward blend (atomics) and
// unpacked by the backward preprocess: [0..9] dL/dview2gaussian, [10..12] dL/dcolor,
// [13] dL/dopacity, [14..16] dL/dmean2D (x, y, |x|+|y|), [17..19] pad.
constexpr int GACC_FLOATS = 20;

}  // namespace gof
