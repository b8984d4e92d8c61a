// epilogue.cu -- the post-processing of render_predicted_more_v2_gof as ONE kernel, forward and backward.
//
// Replaces the torch op sequence of the reference's L2 wrapper (src/gaussian_renderer/__init__.py:881-909,
// 1043-1053): F.normalize of the rendered normal, a 4x4 inverse(), the view->world rotation, the back-projection of
// the median depth through inverse intrinsics (meshgrid, two matmuls), finite differences, cross product and a second
// F.normalize -- about twenty small launches per frame, which rival the rasterizer itself at 256x256.
//
//   normal_world[3,H,W] = R_c2w * (n / max(|n|, 1e-12)),          n = out_color[3:6]
//   depth_normal[3,H,W] = c / max(|c|, 1e-12),  c = dP/dy x dP/dx, P(x,y) = d(x,y) * ray(x,y) + o   (0 on the border)
//       ray(x,y) = R_c2w * ((x - W/2)/fx, (y - H/2)/fy, 1),  d = out_color[6] (median depth),
//       dP/dy = P(x,y+1) - P(x,y-1),  dP/dx = P(x+1,y) - P(x-1,y)
// with R_c2w | o the inverse of the column-vector world->view matrix A | t, A[r][c] = vm[4c+r], t[r] = vm[12+r].
//
// The backward produces dL/d out_color[9,H,W] from dL/d normal_world and dL/d depth_normal (channels 3..5 and 6;
// the other channels are written as zeros so the caller needs no memset).  It is a GATHER: pixel q's depth enters the
// normals of its four neighbours, so the thread of q re-evaluates those four cross products and sums
//   dL/dd(q) = ray(q) . [ Tdx(x,y-1) - Tdx(x,y+1) + Tdy(x-1,y) - Tdy(x+1,y) ],
//   Tdx(c) = dy_c x G_c,  Tdy(c) = G_c x dx_c,  G_c = dL/dc through the normalisation,
// which is deterministic (no atomics) and touches each depth a handful of times out of L1.
// The finite differences are formed from the depth DIFFERENCE times the centre ray plus the depth sum times the
// constant per-pixel ray increment (center_diffs): the camera origin and the common part of the two rays cancel
// exactly instead of being added to both points and subtracted again in float32 (the reference loses
// ~|P|/|dP| * 2^-24 ~ 1e-4 relative there).
#include "gof_common.cuh"
#include <math.h>

namespace gof {

namespace {

__device__ __forceinline__ void load_cam(const float* __restrict__ vm, float* s_Ai)
{
	float A[3][3];
	for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) A[r][c] = vm[4 * c + r];
	const float c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
	const float c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2];
	const float c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
	const float det = A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02;
	const float id = 1.0f / det;
	s_Ai[0] = c00 * id; s_Ai[3] = c01 * id; s_Ai[6] = c02 * id;
	s_Ai[1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * id;
	s_Ai[4] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * id;
	s_Ai[7] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * id;
	s_Ai[2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * id;
	s_Ai[5] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * id;
	s_Ai[8] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * id;
}

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 cross3(const F3& a, const F3& b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
__device__ __forceinline__ float dot3(const F3& a, const F3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// world-space ray of pixel (px, py): R_c2w * ((px - W/2)/fx, (py - H/2)/fy, 1)
__device__ __forceinline__ F3 pixel_ray_world(const float* Ai, int px, int py, int W, int H, float fx, float fy)
{
	const float cx = (px - W / 2.f) / fx, cy = (py - H / 2.f) / fy;
	return { Ai[0] * cx + Ai[1] * cy + Ai[2], Ai[3] * cx + Ai[4] * cy + Ai[5], Ai[6] * cx + Ai[7] * cy + Ai[8] };
}

// Finite differences at an interior centre pixel (x, y): dx along image rows (y), dy along columns (x).
// With rc = ray(x, y), ey = R_c2w[:,1]/fy and ex = R_c2w[:,0]/fx (the rays of neighbouring pixels differ by exactly
// these constant vectors):
//     dx = d(x,y+1) (rc + ey) - d(x,y-1) (rc - ey) = (d_dn - d_up) rc + (d_dn + d_up) ey
//     dy = d(x+1,y) (rc + ex) - d(x-1,y) (rc - ex) = (d_rt - d_lf) rc + (d_rt + d_lf) ex
// The depth difference is formed first, so neither the camera origin nor the common part of the two rays is added and
// subtracted again in float32.  The cross product is expanded likewise,
//     c = dx x dy = ty sx (rc x ex) + sy tx (ey x rc) + sy sx (ey x ex),    (the rc x rc term is identically zero)
// so that at silhouette corners, where both differences are ~ depth * rc and nearly parallel, nothing cancels.
__device__ __forceinline__ void center_diffs(const float* __restrict__ depth, const float* Ai, int x, int y, int W, int H,
                                             float fx, float fy, F3& dx, F3& dy, F3& c)
{
	const float dd = depth[(size_t)(y + 1) * W + x], du = depth[(size_t)(y - 1) * W + x];
	const float dr = depth[(size_t)y * W + x + 1], dl = depth[(size_t)y * W + x - 1];
	const F3 rc = pixel_ray_world(Ai, x, y, W, H, fx, fy);
	const float ify = 1.0f / fy, ifx = 1.0f / fx;
	const float sy = (dd + du) * ify, sx = (dr + dl) * ifx, ty = dd - du, tx = dr - dl;
	dx = { fmaf(ty, rc.x, sy * Ai[1]), fmaf(ty, rc.y, sy * Ai[4]), fmaf(ty, rc.z, sy * Ai[7]) };
	dy = { fmaf(tx, rc.x, sx * Ai[0]), fmaf(tx, rc.y, sx * Ai[3]), fmaf(tx, rc.z, sx * Ai[6]) };
	const F3 ex = { Ai[0], Ai[3], Ai[6] }, ey = { Ai[1], Ai[4], Ai[7] };
	const F3 a = cross3(rc, ex), b = cross3(ey, rc), e = cross3(ey, ex);
	const float ka = ty * sx, kb = sy * tx, ke = sy * sx;
	c = { fmaf(ka, a.x, fmaf(kb, b.x, ke * e.x)), fmaf(ka, a.y, fmaf(kb, b.y, ke * e.y)), fmaf(ka, a.z, fmaf(kb, b.z, ke * e.z)) };
}

__global__ void __launch_bounds__(128)
epilogue_fwd_kernel(const float* __restrict__ out_color_all, const float* __restrict__ vm_all, int W, int H,
                    float fx, float fy, float* __restrict__ normal_world_all, float* __restrict__ depth_normal_all)
{
	const size_t N = (size_t)W * H;
	const float* out_color = out_color_all + (size_t)blockIdx.z * OUT_CH * N;
	float* normal_world = normal_world_all ? normal_world_all + (size_t)blockIdx.z * 3 * N : nullptr;
	float* depth_normal = depth_normal_all ? depth_normal_all + (size_t)blockIdx.z * 3 * N : nullptr;
	__shared__ float s_Ai[9];
	if (threadIdx.x == 0) load_cam(vm_all + 16 * blockIdx.z, s_Ai);
	__syncthreads();
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	const int y = blockIdx.y;
	if (x >= W || y >= H) return;
	const size_t pid = (size_t)y * W + x;

	if (normal_world) {
		float n0 = out_color[3 * N + pid], n1 = out_color[4 * N + pid], n2 = out_color[5 * N + pid];
		const float nrm = fmaxf(sqrtf(n0 * n0 + n1 * n1 + n2 * n2), 1e-12f);
		n0 /= nrm; n1 /= nrm; n2 /= nrm;
		normal_world[0 * N + pid] = s_Ai[0] * n0 + s_Ai[1] * n1 + s_Ai[2] * n2;
		normal_world[1 * N + pid] = s_Ai[3] * n0 + s_Ai[4] * n1 + s_Ai[5] * n2;
		normal_world[2 * N + pid] = s_Ai[6] * n0 + s_Ai[7] * n1 + s_Ai[8] * n2;
	}
	if (depth_normal) {
		F3 o = { 0.f, 0.f, 0.f };
		if (x >= 1 && x < W - 1 && y >= 1 && y < H - 1) {
			F3 dx, dy, c;
			center_diffs(out_color + CH_DEPTH * N, s_Ai, x, y, W, H, fx, fy, dx, dy, c);
			const float nrm = fmaxf(sqrtf(dot3(c, c)), 1e-12f);
			o = { c.x / nrm, c.y / nrm, c.z / nrm };
		}
		depth_normal[0 * N + pid] = o.x;
		depth_normal[1 * N + pid] = o.y;
		depth_normal[2 * N + pid] = o.z;
	}
}

// dL/dv of v / max(|v|, eps) given dL/d(normalised) = g
__device__ __forceinline__ F3 normalize_backward(const F3& v, const F3& g)
{
	const float len = sqrtf(dot3(v, v));
	if (!(len > 1e-12f)) return { g.x * 1e12f, g.y * 1e12f, g.z * 1e12f };
	const float inv = 1.0f / len;
	const F3 n = { v.x * inv, v.y * inv, v.z * inv };
	const float ng = dot3(n, g);
	return { (g.x - n.x * ng) * inv, (g.y - n.y * ng) * inv, (g.z - n.z * ng) * inv };
}

// Contribution terms of an interior centre pixel c: Tdx = dy x G, Tdy = G x dx  (zero for border centres).
__device__ __forceinline__ void center_terms(const float* __restrict__ depth, const float* __restrict__ g_dn, size_t N,
                                             const float* Ai, int x, int y, int W, int H, float fx, float fy, F3& Tdx, F3& Tdy)
{
	Tdx = { 0.f, 0.f, 0.f };
	Tdy = { 0.f, 0.f, 0.f };
	if (!(x >= 1 && x < W - 1 && y >= 1 && y < H - 1)) return;
	F3 dx, dy, c;
	center_diffs(depth, Ai, x, y, W, H, fx, fy, dx, dy, c);
	const size_t pid = (size_t)y * W + x;
	const F3 g = { g_dn[pid], g_dn[N + pid], g_dn[2 * N + pid] };
	const F3 G = normalize_backward(c, g);
	Tdx = cross3(dy, G);
	Tdy = cross3(G, dx);
}

__global__ void __launch_bounds__(128)
epilogue_bwd_kernel(const float* __restrict__ out_color_all, const float* __restrict__ vm_all, int W, int H, float fx, float fy,
                    const float* __restrict__ g_nw_all, const float* __restrict__ g_dn_all, float* __restrict__ dL_dout_all)
{
	const size_t N = (size_t)W * H;
	const float* out_color = out_color_all + (size_t)blockIdx.z * OUT_CH * N;
	const float* g_nw = g_nw_all ? g_nw_all + (size_t)blockIdx.z * 3 * N : nullptr;
	const float* g_dn = g_dn_all ? g_dn_all + (size_t)blockIdx.z * 3 * N : nullptr;
	float* dL_dout = dL_dout_all + (size_t)blockIdx.z * OUT_CH * N;
	__shared__ float s_Ai[9];
	if (threadIdx.x == 0) load_cam(vm_all + 16 * blockIdx.z, s_Ai);
	__syncthreads();
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	const int y = blockIdx.y;
	if (x >= W || y >= H) return;
	const size_t pid = (size_t)y * W + x;

	F3 dn = { 0.f, 0.f, 0.f };
	if (g_nw) {
		const F3 g = { g_nw[pid], g_nw[N + pid], g_nw[2 * N + pid] };
		// through the rotation: R_c2w^T g, then through the normalisation
		const F3 gv = { s_Ai[0] * g.x + s_Ai[3] * g.y + s_Ai[6] * g.z, s_Ai[1] * g.x + s_Ai[4] * g.y + s_Ai[7] * g.z,
		                s_Ai[2] * g.x + s_Ai[5] * g.y + s_Ai[8] * g.z };
		const F3 n = { out_color[3 * N + pid], out_color[4 * N + pid], out_color[5 * N + pid] };
		dn = normalize_backward(n, gv);
	}
	float dd = 0.f;
	if (g_dn) {
		const float* depth = out_color + CH_DEPTH * N;
		F3 a, b, acc = { 0.f, 0.f, 0.f };
		center_terms(depth, g_dn, N, s_Ai, x, y - 1, W, H, fx, fy, a, b);     // q is the centre's (x, y+1) point: +Tdx
		acc = { acc.x + a.x, acc.y + a.y, acc.z + a.z };
		center_terms(depth, g_dn, N, s_Ai, x, y + 1, W, H, fx, fy, a, b);     // q is its (x, y-1) point: -Tdx
		acc = { acc.x - a.x, acc.y - a.y, acc.z - a.z };
		center_terms(depth, g_dn, N, s_Ai, x - 1, y, W, H, fx, fy, a, b);     // q is its (x+1, y) point: +Tdy
		acc = { acc.x + b.x, acc.y + b.y, acc.z + b.z };
		center_terms(depth, g_dn, N, s_Ai, x + 1, y, W, H, fx, fy, a, b);     // q is its (x-1, y) point: -Tdy
		acc = { acc.x - b.x, acc.y - b.y, acc.z - b.z };
		dd = dot3(acc, pixel_ray_world(s_Ai, x, y, W, H, fx, fy));
	}
	dL_dout[0 * N + pid] = 0.f;
	dL_dout[1 * N + pid] = 0.f;
	dL_dout[2 * N + pid] = 0.f;
	dL_dout[3 * N + pid] = dn.x;
	dL_dout[4 * N + pid] = dn.y;
	dL_dout[5 * N + pid] = dn.z;
	dL_dout[CH_DEPTH * N + pid] = dd;
	dL_dout[CH_ALPHA * N + pid] = 0.f;
	dL_dout[CH_DIST * N + pid] = 0.f;
}

}  // namespace
}  // namespace gof

using namespace gof;

extern "C" {

int gof_render_epilogue_batch(const float* out_color, const float* viewmatrix, int32_t V, int32_t W, int32_t H,
                              float fovx, float fovy, float* normal_world, float* depth_normal, gof_stream_t stream)
{
	if (!out_color || !viewmatrix || W <= 0 || H <= 0 || V <= 0) { set_error("gof_render_epilogue: bad argument"); return GOF_EINVAL; }
	const float fx = W / (2.f * tanf(fovx / 2.f));
	const float fy = H / (2.f * tanf(fovy / 2.f));
	dim3 grid((W + 127) / 128, H, V);
	epilogue_fwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(out_color, viewmatrix, W, H, fx, fy, normal_world, depth_normal);
	GOF_CUDA_CHECK(cudaGetLastError());
	return GOF_OK;
}

int gof_render_epilogue(const float* out_color, const float* viewmatrix, int32_t W, int32_t H,
                        float fovx, float fovy, float* normal_world, float* depth_normal, gof_stream_t stream)
{
	return gof_render_epilogue_batch(out_color, viewmatrix, 1, W, H, fovx, fovy, normal_world, depth_normal, stream);
}

int gof_render_epilogue_backward_batch(const float* out_color, const float* viewmatrix, int32_t V, int32_t W, int32_t H,
                                       float fovx, float fovy, const float* dL_dnormal_world, const float* dL_ddepth_normal,
                                       float* dL_dout_color, gof_stream_t stream)
{
	if (!out_color || !viewmatrix || !dL_dout_color || W <= 0 || H <= 0 || V <= 0) {
		set_error("gof_render_epilogue_backward: bad argument");
		return GOF_EINVAL;
	}
	const float fx = W / (2.f * tanf(fovx / 2.f));
	const float fy = H / (2.f * tanf(fovy / 2.f));
	dim3 grid((W + 127) / 128, H, V);
	epilogue_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(out_color, viewmatrix, W, H, fx, fy, dL_dnormal_world,
	                                                            dL_ddepth_normal, dL_dout_color);
	GOF_CUDA_CHECK(cudaGetLastError());
	return GOF_OK;
}

}  // extern "C"
