// oracle/ref_harness.cu -- TEST INFRASTRUCTURE, not product code.
//
// A thin extern "C" driver around the UNMODIFIED reference rasterizer
// (CudaRasterizer::Rasterizer::{forward,backward,integrate,markVisible},
// RAST/cuda_rasterizer/rasterizer.h:20-123), compiled for sm_100a from the
// sources where they lie under /root/reference by oracle/Makefile into
// oracle/_ref/libgof_ref.so.  It replaces the reference's torch glue
// (RAST/rasterize_points.cu:36-211) with plain cudaMalloc-backed state blobs so
// that it can be driven through ctypes without torch headers.
//
// Used only by tests/, __graft_entry__.smoke() and bench.py (reference arm).
// Nothing in f3d_gaus_b200/ may load this library.
//
// All pointers are DEVICE pointers unless noted.  The reference launches on the
// legacy default stream; callers synchronise with torch's default stream
// (which is the same stream).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <cuda_runtime.h>
#include "rasterizer_impl.h"   // reference header (state carving), found via -I

namespace {

struct Blob {
	char* ptr = nullptr;
	size_t cap = 0;
	size_t size = 0;
	char* resize(size_t n) {
		if (n > cap) {
			if (ptr) cudaFree(ptr);
			size_t want = n + n / 4 + 256;
			if (cudaMalloc(&ptr, want) != cudaSuccess) { ptr = nullptr; cap = 0; throw std::runtime_error("cudaMalloc failed"); }
			cap = want;
		}
		size = n;
		return ptr;
	}
	~Blob() { if (ptr) cudaFree(ptr); }
};

thread_local std::string g_err;

}  // namespace

struct RefState {
	Blob geom, binning, img, point, point_binning;
	int P = 0, R = 0, W = 0, H = 0;
};

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

RefState* ref_state_create() { return new RefState(); }
void ref_state_destroy(RefState* s) { delete s; }

// Returns num_rendered (R) or -1 on error.
int ref_forward(RefState* s, int P, int D, int M,
	const float* background, int W, int H,
	const float* means3D, const float* shs, const float* colors_precomp,
	const float* opacities, const float* scales, float scale_modifier,
	const float* rotations, const float* cov3D_precomp, const float* view2gaussian_precomp,
	const float* viewmatrix, const float* projmatrix, const float* cam_pos,
	float tan_fovx, float tan_fovy, float kernel_size, const float* subpixel_offset,
	int prefiltered, float* out_color, int* radii, int debug)
{
	try {
		std::function<char*(size_t)> g = [s](size_t n) { return s->geom.resize(n); };
		std::function<char*(size_t)> b = [s](size_t n) { return s->binning.resize(n); };
		std::function<char*(size_t)> i = [s](size_t n) { return s->img.resize(n); };
		int R = CudaRasterizer::Rasterizer::forward(g, b, i, P, D, M, background, W, H,
			means3D, shs, colors_precomp, opacities, scales, scale_modifier, rotations,
			cov3D_precomp, view2gaussian_precomp, viewmatrix, projmatrix, cam_pos,
			tan_fovx, tan_fovy, kernel_size, subpixel_offset, prefiltered != 0,
			out_color, radii, debug != 0);
		s->P = P; s->R = R; s->W = W; s->H = H;
		return R;
	} catch (const std::exception& e) {
		g_err = e.what();
		return -1;
	}
}

int ref_backward(RefState* s, int P, int D, int M, int R,
	const float* background, int W, int H,
	const float* means3D, const float* shs, const float* colors_precomp,
	const float* view2gaussian_precomp, const float* scales, float scale_modifier,
	const float* rotations, const float* cov3D_precomp,
	const float* viewmatrix, const float* projmatrix, const float* campos,
	float tan_fovx, float tan_fovy, float kernel_size, const float* subpixel_offset,
	const int* radii, const float* dL_dpix,
	float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
	float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale,
	float* dL_drot, float* dL_dview2gaussian, int debug)
{
	try {
		CudaRasterizer::Rasterizer::backward(P, D, M, R, background, W, H, means3D, shs,
			colors_precomp, view2gaussian_precomp, scales, scale_modifier, rotations,
			cov3D_precomp, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, kernel_size,
			subpixel_offset, radii, s->geom.ptr, s->binning.ptr, s->img.ptr, dL_dpix,
			dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh,
			dL_dscale, dL_drot, dL_dview2gaussian, debug != 0);
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		return -1;
	}
}

// Rasterizer::integrate (rasterizer.h:93-123).  Returns num_rendered or -1.
int ref_integrate(RefState* s, int PN, int P, int D, int M,
	const float* background, int W, int H, const float* points3D,
	const float* means3D, const float* shs, const float* colors_precomp,
	const float* opacities, const float* scales, float scale_modifier,
	const float* rotations, const float* cov3D_precomp, const float* view2gaussian_precomp,
	const float* viewmatrix, const float* projmatrix, const float* cam_pos,
	float tan_fovx, float tan_fovy, float kernel_size, const float* subpixel_offset,
	int prefiltered, float* out_color, int* radii, float* out_alpha_integrated, float* out_color_integrated, int debug)
{
	try {
		std::function<char*(size_t)> g = [s](size_t n) { return s->geom.resize(n); };
		std::function<char*(size_t)> b = [s](size_t n) { return s->binning.resize(n); };
		std::function<char*(size_t)> i = [s](size_t n) { return s->img.resize(n); };
		std::function<char*(size_t)> pt = [s](size_t n) { return s->point.resize(n); };
		std::function<char*(size_t)> pb = [s](size_t n) { return s->point_binning.resize(n); };
		int R = CudaRasterizer::Rasterizer::integrate(g, b, i, pt, pb, PN, P, D, M, background, W, H, points3D,
			means3D, shs, colors_precomp, opacities, scales, scale_modifier, rotations,
			cov3D_precomp, view2gaussian_precomp, viewmatrix, projmatrix, cam_pos,
			tan_fovx, tan_fovy, kernel_size, subpixel_offset, prefiltered != 0,
			out_color, radii, out_alpha_integrated, out_color_integrated, debug != 0);
		s->P = P; s->R = R; s->W = W; s->H = H;
		return R;
	} catch (const std::exception& e) {
		g_err = e.what();
		return -1;
	}
}

int ref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, unsigned char* present)
{
	CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, (bool*)present);
	return 0;
}

// Decode one named array of the reference's opaque state blobs
// (layout: RAST/cuda_rasterizer/rasterizer_impl.cu:188-243) into a caller
// device buffer.  Returns the number of bytes the array holds, or -1.
// If dst == nullptr only the size is returned.
long long ref_state_get(RefState* s, const char* name, void* dst, long long dst_bytes)
{
	using namespace CudaRasterizer;
	const size_t P = s->P, R = s->R, N = (size_t)s->W * s->H;
	char* gp = s->geom.ptr; char* bp = s->binning.ptr; char* ip = s->img.ptr;
	if (!gp || !ip) { g_err = "no forward state"; return -1; }
	GeometryState geo = GeometryState::fromChunk(gp, P);
	ImageState im = ImageState::fromChunk(ip, N);
	BinningState bin{};
	if (bp) bin = BinningState::fromChunk(bp, R);
	const void* src = nullptr; size_t bytes = 0;
	std::string n(name);
	const size_t T = (size_t)((s->W + 15) / 16) * ((s->H + 15) / 16);
	if (n == "depths") { src = geo.depths; bytes = P * 4; }
	else if (n == "clamped") { src = geo.clamped; bytes = P * 3; }
	else if (n == "internal_radii") { src = geo.internal_radii; bytes = P * 4; }
	else if (n == "means2D") { src = geo.means2D; bytes = P * 8; }
	else if (n == "cov3D") { src = geo.cov3D; bytes = P * 24; }
	else if (n == "view2gaussian") { src = geo.view2gaussian; bytes = P * 40; }
	else if (n == "conic_opacity") { src = geo.conic_opacity; bytes = P * 16; }
	else if (n == "rgb") { src = geo.rgb; bytes = P * 12; }
	else if (n == "tiles_touched") { src = geo.tiles_touched; bytes = P * 4; }
	else if (n == "point_offsets") { src = geo.point_offsets; bytes = P * 4; }
	else if (n == "final_T") { src = im.accum_alpha; bytes = N * 16; }
	else if (n == "n_contrib") { src = im.n_contrib; bytes = N * 8; }
	else if (n == "ranges") { src = im.ranges; bytes = T * 8; }
	else if (n == "point_list") { src = bin.point_list; bytes = R * 4; }
	else if (n == "point_list_unsorted") { src = bin.point_list_unsorted; bytes = R * 4; }
	else if (n == "point_list_keys") { src = bin.point_list_keys; bytes = R * 8; }
	else if (n == "point_list_keys_unsorted") { src = bin.point_list_keys_unsorted; bytes = R * 8; }
	else { g_err = "unknown state array: " + n; return -1; }
	if (dst) {
		if ((long long)bytes > dst_bytes) { g_err = "dst too small"; return -1; }
		if (bytes && cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToDevice) != cudaSuccess) { g_err = "memcpy failed"; return -1; }
	}
	return (long long)bytes;
}

}  // extern "C"
