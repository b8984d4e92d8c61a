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

// One per-Gaussian "blend record": everything the per-tile blend needs, 64 B so that a
// tile's sorted slab is a contiguous run of 64-byte records that TMA bulk copies can stream.
//  [0..5] Sigma_v (xx,xy,xz,yy,yz,zz)  [6..8] B  [9] C      (view2gaussian, forward.cu:268-277)
//  [10] tau: conservative reject threshold on the ray-minimum value (see blend_math.cuh)
//  [11] w = opacity*coef (conic_opacity.w)  [12..14] rgb  [15] Gaussian index (int bits)
// The first three float4 are all the float32 pre-test needs; the fourth only feeds
// contributing pairs.
constexpr int REC_FLOATS = 16;
constexpr int REC_BYTES = 64;
constexpr int REC_TAU = 10, REC_W = 11, REC_RGB = 12, REC_ID = 15;

constexpr size_t ALIGN = 256;
__host__ __device__ inline size_t align_up(size_t x, size_t a = ALIGN) { return (x + a - 1) / a * a; }

// ---- opaque state layouts (our own; the reference's are rasterizer_impl.cu:188-243) -------
struct GeomState {
	float* depths;          // [P]
	float2* means2D;        // [P]
	float4* conic_opacity;  // [P]
	float* rec;             // [P,16]
	uint32_t* tiles_touched;// [P]
	uint32_t* point_offsets;// [P]  inclusive scan of tiles_touched
	uint8_t* clamped;       // [P,3]
	int32_t* mailbox;       // [4]: {num_rendered, overflow flag, -, -}
	char* scan_temp; size_t scan_temp_bytes;
	size_t total;
	static GeomState carve(char* base, size_t P);
};
struct ImgState {
	float* final_T;         // [4,N]  T, dist1, dist2, distortion_raw  (forward.cu:591-594)
	uint32_t* n_contrib;    // [2,N]  last_contributor, max_contributor (forward.cu:596-597)
	uint2* ranges;          // [T]
	size_t total;
	static ImgState carve(char* base, size_t N, size_t T);
};
struct BinState {
	uint64_t* keys_unsorted;   // [R]
	uint64_t* keys;            // [R]
	uint32_t* vals_unsorted;   // [R]
	uint32_t* point_list;      // [R]
	float* slab;               // [R,16] tile-ordered blend records
	char* sort_temp; size_t sort_temp_bytes;
	size_t total;
	static BinState carve(char* base, size_t R);
};

size_t scan_temp_bytes(size_t P);
size_t sort_temp_bytes(size_t R);

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

// ---- stage launchers (one .cu each) ---------------------------------------------------------
int launch_preprocess(const GofParams& prm, const GofInputs& in, float focal_x, float focal_y,
                      dim3 tile_grid, const GeomState& g, int32_t* radii, cudaStream_t s);
int launch_scan(const GeomState& g, int P, cudaStream_t s);
int launch_binning(const GofParams& prm, dim3 tile_grid, const GeomState& g, const ImgState& im,
                   const BinState& b, const int32_t* radii, int R, cudaStream_t s);
int launch_render_fwd(const GofParams& prm, dim3 tile_grid, float focal_x, float focal_y,
                      const ImgState& im, const BinState& b, const float* background,
                      float* out_color, cudaStream_t s);
int launch_render_bwd(const GofParams& prm, dim3 tile_grid, float focal_x, float focal_y,
                      const GeomState& g, const ImgState& im, const BinState& b,
                      const float* background, const float* dL_dpix, float* gacc, cudaStream_t s);
int launch_preprocess_bwd(const GofParams& prm, const GofInputs& in, const GeomState& g,
                          const int32_t* radii, const float* gacc, const GofGrads& grads,
                          cudaStream_t s);

// Per-Gaussian packed gradient accumulator written by the backward blend (atomics) and
// unpacked by the backward preprocess: [0..9] dL/dview2gaussian, [10..12] dL/dcolor,
// [13] dL/dopacity, [14..16] dL/dmean2D (x, y, |x|+|y|), [17..19] pad.
constexpr int GACC_FLOATS = 20;

}  // namespace gof
