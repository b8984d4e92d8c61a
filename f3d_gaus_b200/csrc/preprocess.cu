// preprocess.cu -- per-Gaussian projection / covariance / view-to-Gaussian preprocess (K1).
//
// Replaces preprocessCUDA<3> and its helpers (RAST/cuda_rasterizer/forward.cu:20-404,
// auxiliary.h:59-74,177-202).  Behavioural contract: the float32
// state this kernel stores (depths, means2D, conic_opacity, rgb, the 10-float view2gaussian
// quadric) feeds an ill-conditioned blend, so every expression below keeps the reference's
// evaluation order and precision islands (which products are summed in which order, where
// double is used) -- the arithmetic is the specification.  What is ours: the memory system.
//   * AoS inputs (xyz/scale: 12 B, rot: 16 B, SH: 12*M B per Gaussian) are fetched with
//     block-cooperative 128-bit loads through shared memory (xyz, scale) or directly as
//     float4 (rot, SH) instead of per-thread scalar strided loads;
//   * outputs are written as SoA plus one packed 64-byte "blend record" per Gaussian
//     (gof_common.cuh) so the binning stage can build TMA-streamable tile slabs.
#include "gof_common.cuh"
#include <cstdio>

namespace gof {

namespace {

constexpr int PRE_THREADS = 256;
constexpr int HIST_TILES = 4096;      // shared-memory tile histogram up to this many tiles per view (16 KB)

__device__ const float kSH_C0 = 0.28209479177387814f;
__device__ const float kSH_C1 = 0.4886025119029199f;
__device__ const float kSH_C2[] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                   -1.0925484305920792f, 0.5462742152960396f};
__device__ const float kSH_C3[] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                   0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                   -0.5900435899266435f};

struct V3 { float x, y, z; };
// Column-major 3x3 (c[col][row]), the storage convention of the reference's matrix library.
struct M3 { float c[3][3]; };
struct M4 { float c[4][4]; };

__device__ __forceinline__ M3 m3_mul(const M3& a, const M3& b)
{
	M3 r;
#pragma unroll
	for (int i = 0; i < 3; i++)
#pragma unroll
		for (int j = 0; j < 3; j++)
			r.c[i][j] = a.c[0][j] * b.c[i][0] + a.c[1][j] * b.c[i][1] + a.c[2][j] * b.c[i][2];
	return r;
}
__device__ __forceinline__ M3 m3_t(const M3& a)
{
	M3 r;
#pragma unroll
	for (int i = 0; i < 3; i++)
#pragma unroll
		for (int j = 0; j < 3; j++) r.c[i][j] = a.c[j][i];
	return r;
}
// column-vector product M*v
__device__ __forceinline__ V3 m3_mulv(const M3& m, const V3& v)
{
	return { m.c[0][0] * v.x + m.c[1][0] * v.y + m.c[2][0] * v.z,
	         m.c[0][1] * v.x + m.c[1][1] * v.y + m.c[2][1] * v.z,
	         m.c[0][2] * v.x + m.c[1][2] * v.y + m.c[2][2] * v.z };
}
// row-vector product v*M
__device__ __forceinline__ V3 m3_vmul(const V3& v, const M3& m)
{
	return { m.c[0][0] * v.x + m.c[0][1] * v.y + m.c[0][2] * v.z,
	         m.c[1][0] * v.x + m.c[1][1] * v.y + m.c[1][2] * v.z,
	         m.c[2][0] * v.x + m.c[2][1] * v.y + m.c[2][2] * v.z };
}
// 4x4 product, column by column: r[i] = a[0]*b[i][0] + a[1]*b[i][1] + a[2]*b[i][2] + a[3]*b[i][3]
__device__ __forceinline__ M4 m4_mul(const M4& a, const M4& b)
{
	M4 r;
#pragma unroll
	for (int i = 0; i < 4; i++)
#pragma unroll
		for (int j = 0; j < 4; j++)
			r.c[i][j] = a.c[0][j] * b.c[i][0] + a.c[1][j] * b.c[i][1] + a.c[2][j] * b.c[i][2] + a.c[3][j] * b.c[i][3];
	return r;
}

// Rotation matrix of the (un-normalised) quaternion (r,x,y,z), in the element order the
// reference feeds to its column-major constructor (forward.cu:145-149,179-183).
__device__ __forceinline__ M3 quat_to_m3(float r, float x, float y, float z)
{
	M3 R;
	R.c[0][0] = 1.f - 2.f * (y * y + z * z); R.c[0][1] = 2.f * (x * y - r * z);       R.c[0][2] = 2.f * (x * z + r * y);
	R.c[1][0] = 2.f * (x * y + r * z);       R.c[1][1] = 1.f - 2.f * (x * x + z * z); R.c[1][2] = 2.f * (y * z - r * x);
	R.c[2][0] = 2.f * (x * z - r * y);       R.c[2][1] = 2.f * (y * z + r * x);       R.c[2][2] = 1.f - 2.f * (x * x + y * y);
	return R;
}

__device__ __forceinline__ float ndc_to_pix(float v, int S)
{
	return ((v + 1.0) * S - 1.0) * 0.5;   // evaluated in double (auxiliary.h:59-62)
}

// 3D covariance from scale/rotation (forward.cu:129-163): Sigma = (S R)^T (S R).
__device__ __forceinline__ void cov3d_from_scale_rot(const V3& scale, float mod, const float4& rot, float* cov3D)
{
	M3 S;
#pragma unroll
	for (int i = 0; i < 3; i++)
#pragma unroll
		for (int j = 0; j < 3; j++) S.c[i][j] = (i == j) ? 1.0f : 0.0f;
	S.c[0][0] = mod * scale.x;
	S.c[1][1] = mod * scale.y;
	S.c[2][2] = mod * scale.z;
	M3 R = quat_to_m3(rot.x, rot.y, rot.z, rot.w);
	M3 M = m3_mul(S, R);
	M3 Sigma = m3_mul(m3_t(M), M);
	cov3D[0] = Sigma.c[0][0];
	cov3D[1] = Sigma.c[0][1];
	cov3D[2] = Sigma.c[0][2];
	cov3D[3] = Sigma.c[1][1];
	cov3D[4] = Sigma.c[1][2];
	cov3D[5] = Sigma.c[2][2];
}

// EWA 2D covariance + the kernel-size opacity coefficient (forward.cu:74-124).
__device__ __forceinline__ float4 cov2d_ewa(const V3& mean, float focal_x, float focal_y, float tan_fovx,
                                            float tan_fovy, float kernel_size, const float* cov3D, const float* vm)
{
	V3 t = { vm[0] * mean.x + vm[4] * mean.y + vm[8] * mean.z + vm[12],
	         vm[1] * mean.x + vm[5] * mean.y + vm[9] * mean.z + vm[13],
	         vm[2] * mean.x + vm[6] * mean.y + vm[10] * mean.z + vm[14] };
	const float limx = 1.3f * tan_fovx;
	const float limy = 1.3f * tan_fovy;
	const float txtz = t.x / t.z;
	const float tytz = t.y / t.z;
	t.x = min(limx, max(-limx, txtz)) * t.z;
	t.y = min(limy, max(-limy, tytz)) * t.z;

	M3 J;
	J.c[0][0] = focal_x / t.z; J.c[0][1] = 0.0f;          J.c[0][2] = -(focal_x * t.x) / (t.z * t.z);
	J.c[1][0] = 0.0f;          J.c[1][1] = focal_y / t.z; J.c[1][2] = -(focal_y * t.y) / (t.z * t.z);
	J.c[2][0] = 0;             J.c[2][1] = 0;             J.c[2][2] = 0;
	M3 Wm;
	Wm.c[0][0] = vm[0]; Wm.c[0][1] = vm[4]; Wm.c[0][2] = vm[8];
	Wm.c[1][0] = vm[1]; Wm.c[1][1] = vm[5]; Wm.c[1][2] = vm[9];
	Wm.c[2][0] = vm[2]; Wm.c[2][1] = vm[6]; Wm.c[2][2] = vm[10];
	M3 T = m3_mul(Wm, J);
	M3 Vrk;
	Vrk.c[0][0] = cov3D[0]; Vrk.c[0][1] = cov3D[1]; Vrk.c[0][2] = cov3D[2];
	Vrk.c[1][0] = cov3D[1]; Vrk.c[1][1] = cov3D[3]; Vrk.c[1][2] = cov3D[4];
	Vrk.c[2][0] = cov3D[2]; Vrk.c[2][1] = cov3D[4]; Vrk.c[2][2] = cov3D[5];
	M3 cov = m3_mul(m3_mul(m3_t(T), m3_t(Vrk)), T);

	const float det_0 = max(1e-6, cov.c[0][0] * cov.c[1][1] - cov.c[0][1] * cov.c[0][1]);
	const float det_1 = max(1e-6, (cov.c[0][0] + kernel_size) * (cov.c[1][1] + kernel_size) - cov.c[0][1] * cov.c[0][1]);
	float coef = sqrt(det_0 / (det_1 + 1e-6) + 1e-6);
	if (det_0 <= 1e-6 || det_1 <= 1e-6) coef = 0.0f;
	cov.c[0][0] += kernel_size;
	cov.c[1][1] += kernel_size;
	return { float(cov.c[0][0]), float(cov.c[0][1]), float(cov.c[1][1]), float(coef) };
}

// The 10-float view-space quadric of the Gaussian (forward.cu:168-279): with the camera
// centre t2 and axes R^T expressed in the Gaussian's frame and Sinv = 1/(s^2+1e-7) in double,
//   Sigma_v = R Sinv R^T (6 unique), B = t2^T Sinv R^T, C = t2^T Sinv t2.
// scale_modifier is deliberately NOT applied (reference behaviour).
__device__ __forceinline__ void view2gaussian_quadric(const V3& scale, const V3& mean, const float4& rot,
                                                      const float* vm, float* out)
{
	M3 R = quat_to_m3(rot.x, rot.y, rot.z, rot.w);
	M4 G2W;
	G2W.c[0][0] = R.c[0][0]; G2W.c[0][1] = R.c[1][0]; G2W.c[0][2] = R.c[2][0]; G2W.c[0][3] = 0.0f;
	G2W.c[1][0] = R.c[0][1]; G2W.c[1][1] = R.c[1][1]; G2W.c[1][2] = R.c[2][1]; G2W.c[1][3] = 0.0f;
	G2W.c[2][0] = R.c[0][2]; G2W.c[2][1] = R.c[1][2]; G2W.c[2][2] = R.c[2][2]; G2W.c[2][3] = 0.0f;
	G2W.c[3][0] = mean.x;    G2W.c[3][1] = mean.y;    G2W.c[3][2] = mean.z;    G2W.c[3][3] = 1.0f;
	M4 W2V;
#pragma unroll
	for (int i = 0; i < 4; i++)
#pragma unroll
		for (int j = 0; j < 4; j++) W2V.c[i][j] = vm[4 * i + j];
	M4 G2V = m4_mul(W2V, G2W);

	M3 Rt;
	Rt.c[0][0] = G2V.c[0][0]; Rt.c[0][1] = G2V.c[1][0]; Rt.c[0][2] = G2V.c[2][0];
	Rt.c[1][0] = G2V.c[0][1]; Rt.c[1][1] = G2V.c[1][1]; Rt.c[1][2] = G2V.c[2][1];
	Rt.c[2][0] = G2V.c[0][2]; Rt.c[2][1] = G2V.c[1][2]; Rt.c[2][2] = G2V.c[2][2];
	V3 t = { G2V.c[3][0], G2V.c[3][1], G2V.c[3][2] };
	M3 nRt;
#pragma unroll
	for (int i = 0; i < 3; i++)
#pragma unroll
		for (int j = 0; j < 3; j++) nRt.c[i][j] = -Rt.c[i][j];
	V3 t2 = m3_mulv(nRt, t);

	double3 Sinv = { 1.0f / ((double)scale.x * scale.x + 1e-7), 1.0f / ((double)scale.y * scale.y + 1e-7),
	                 1.0f / ((double)scale.z * scale.z + 1e-7) };
	double C = t2.x * t2.x * Sinv.x + t2.y * t2.y * Sinv.y + t2.z * t2.z * Sinv.z;
	M3 SR;
	SR.c[0][0] = Sinv.x * Rt.c[0][0]; SR.c[0][1] = Sinv.y * Rt.c[0][1]; SR.c[0][2] = Sinv.z * Rt.c[0][2];
	SR.c[1][0] = Sinv.x * Rt.c[1][0]; SR.c[1][1] = Sinv.y * Rt.c[1][1]; SR.c[1][2] = Sinv.z * Rt.c[1][2];
	SR.c[2][0] = Sinv.x * Rt.c[2][0]; SR.c[2][1] = Sinv.y * Rt.c[2][1]; SR.c[2][2] = Sinv.z * Rt.c[2][2];
	V3 B = m3_vmul(t2, SR);
	M3 Sigma = m3_mul(m3_t(Rt), SR);
	out[0] = Sigma.c[0][0];
	out[1] = Sigma.c[0][1];
	out[2] = Sigma.c[0][2];
	out[3] = Sigma.c[1][1];
	out[4] = Sigma.c[1][2];
	out[5] = Sigma.c[2][2];
	out[6] = B.x;
	out[7] = B.y;
	out[8] = B.z;
	out[9] = C;
}

// SH -> RGB, degrees 0..3 (forward.cu:20-71).  lo holds coefficients 0..3 (already in
// registers, [k][channel]); hi points at this Gaussian's [M,3] block in global memory and is
// only dereferenced for coefficients >= 4.
// The roundings are pinned with explicit intrinsics to what the reference's sm_100a build
// evaluates (read off its SASS): c0 = C0*sh0 rounded, then one fused multiply-add per further
// coefficient, with polynomial factors such as 3xx-yy themselves fused (fma(xx,3,-yy)).
__device__ __forceinline__ V3 sh_to_rgb(int deg, const V3& pos, const V3& campos, const float (*lo)[3],
                                        const float* __restrict__ hi, uint8_t* clamped3)
{
	V3 dir = { pos.x - campos.x, pos.y - campos.y, pos.z - campos.z };
	float len = sqrt(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
	const float x = dir.x / len, y = dir.y / len, z = dir.z / len;
	float r[3];
#pragma unroll
	for (int ch = 0; ch < 3; ch++) r[ch] = __fmul_rn(lo[0][ch], kSH_C0);
	if (deg > 0) {
		const float a = __fmul_rn(y, kSH_C1), b = __fmul_rn(z, kSH_C1), c = __fmul_rn(x, kSH_C1);
#pragma unroll
		for (int ch = 0; ch < 3; ch++) {
			r[ch] = __fmaf_rn(-a, lo[1][ch], r[ch]);
			r[ch] = __fmaf_rn(b, lo[2][ch], r[ch]);
			r[ch] = __fmaf_rn(-c, lo[3][ch], r[ch]);
		}
		if (deg > 1) {
			const float xx = __fmul_rn(x, x), yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
			const float xy = __fmul_rn(x, y), yz = __fmul_rn(y, z), xz = __fmul_rn(x, z);
			const float zz2 = __fadd_rn(zz, zz);
			const float xx_yy = __fsub_rn(xx, yy);
			float t[12];
			t[0] = __fmul_rn(xy, kSH_C2[0]);
			t[1] = __fmul_rn(yz, kSH_C2[1]);
			t[2] = __fmul_rn(__fsub_rn(__fsub_rn(zz2, xx), yy), kSH_C2[2]);
			t[3] = __fmul_rn(xz, kSH_C2[3]);
			t[4] = __fmul_rn(xx_yy, kSH_C2[4]);
			int nt = 5;
			if (deg > 2) {
				const float p4 = __fsub_rn(__fmaf_rn(zz, 4.0f, -xx), yy);          // 4zz - xx - yy
				t[5] = __fmul_rn(__fmul_rn(y, kSH_C3[0]), __fmaf_rn(xx, 3.0f, -yy));
				t[6] = __fmul_rn(z, __fmul_rn(xy, kSH_C3[1]));
				t[7] = __fmul_rn(__fmul_rn(y, kSH_C3[2]), p4);
				t[8] = __fmul_rn(__fmul_rn(z, kSH_C3[3]), __fmaf_rn(yy, -3.0f, __fmaf_rn(xx, -3.0f, zz2)));
				t[9] = __fmul_rn(__fmul_rn(x, kSH_C3[4]), p4);
				t[10] = __fmul_rn(__fmul_rn(z, kSH_C3[5]), xx_yy);
				t[11] = __fmul_rn(__fmul_rn(x, kSH_C3[6]), __fmaf_rn(yy, -3.0f, xx));
				nt = 12;
			}
#pragma unroll
			for (int k = 0; k < 12; k++)
				if (k < nt) {
#pragma unroll
					for (int ch = 0; ch < 3; ch++) r[ch] = __fmaf_rn(t[k], hi[3 * (4 + k) + ch], r[ch]);
				}
		}
	}
	float res[3];
#pragma unroll
	for (int ch = 0; ch < 3; ch++) {
		const float v = __fadd_rn(r[ch], 0.5f);
		clamped3[ch] = (v < 0);
		res[ch] = (v < 0) ? 0.0f : v;
	}
	return { res[0], res[1], res[2] };
}

// Block-cooperative copy of `n` floats starting at src (global) into dst (shared), using
// 128-bit loads over the 16-byte aligned interior.
__device__ __forceinline__ void stage_floats(const float* __restrict__ src, float* dst, int n)
{
	const uintptr_t a = reinterpret_cast<uintptr_t>(src);
	int head = (int)(((16 - (a & 15)) & 15) >> 2);
	if (head > n) head = n;
	if ((a & 3) != 0) head = n;  // not even float aligned: scalar only
	for (int i = threadIdx.x; i < head; i += blockDim.x) dst[i] = src[i];
	const int nvec = (n - head) >> 2;
	const float4* s4 = reinterpret_cast<const float4*>(src + head);
	for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
		float4 v = __ldg(s4 + i);
		float* d = dst + head + 4 * i;
		d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
	}
	for (int i = head + 4 * nvec + threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

__device__ __forceinline__ void preprocess_one(const int idx, int P, int D, int M, const float* s_xyz, const float* s_scl,
	const float* s_vm, const float* s_pm, const float* s_cam, const bool has_scales, const float scale_modifier,
	const float* __restrict__ rotations, const float* __restrict__ opacities, const float* __restrict__ shs,
	const float* __restrict__ cov3D_precomp, const float* __restrict__ colors_precomp, const float* __restrict__ v2g_precomp,
	const int W, const int H, const float tan_fovx, const float tan_fovy, const float focal_x, const float focal_y,
	const float kernel_size, int* __restrict__ radii, float2* __restrict__ means2D, float* __restrict__ depths,
	float* __restrict__ rec, float4* __restrict__ conic_opacity, uint8_t* __restrict__ clamped, const dim3 grid,
	uint32_t* __restrict__ tiles_touched, ushort4* __restrict__ rects, uint32_t* s_hist, uint32_t* __restrict__ tile_counts,
	bool prefiltered);

#ifndef GOF_PRE_MIN_CTAS
#define GOF_PRE_MIN_CTAS 4        // 64 registers, 4 CTAs per SM: 40 us against 44 us at 3 (8 views x 65 536 Gaussians)
#endif
__global__ void __launch_bounds__(PRE_THREADS, GOF_PRE_MIN_CTAS)
preprocess_kernel(int P, int D, int M,
	const float* __restrict__ means3D, const float* __restrict__ scales, const float scale_modifier,
	const float* __restrict__ rotations, const float* __restrict__ opacities, const float* __restrict__ shs,
	const float* __restrict__ cov3D_precomp, const float* __restrict__ colors_precomp,
	const float* __restrict__ v2g_precomp, const float* __restrict__ viewmatrices,
	const float* __restrict__ projmatrices, const float* __restrict__ cam_positions,
	const int W, const int H, const float tan_fovx, const float tan_fovy,
	const float focal_x, const float focal_y, const float kernel_size,
	int* __restrict__ radii_all, float2* __restrict__ means2D_all, float* __restrict__ depths_all,
	float* __restrict__ rec_all, float4* __restrict__ conic_opacity_all, uint8_t* __restrict__ clamped_all,
	const dim3 grid, uint32_t* __restrict__ tiles_touched_all, ushort4* __restrict__ rect_all,
	uint32_t* __restrict__ tile_counts_all, bool prefiltered)
{
	__shared__ float s_xyz[PRE_THREADS * 3];
	__shared__ float s_scl[PRE_THREADS * 3];
	__shared__ float s_vm[16], s_pm[16], s_cam[3];
	// Tile histogram of this block's duplicates: counted in shared memory, flushed with one global
	// atomic per touched tile (grids above HIST_TILES tiles count straight into global memory).
	__shared__ uint32_t s_hist[HIST_TILES];
	pdl_trigger();                 // the tile scan may be scheduled behind this grid (it waits for our completion)
	const int T = grid.x * grid.y;
	const bool use_hist = T <= HIST_TILES;
	if (use_hist) for (int i = threadIdx.x; i < T; i += PRE_THREADS) s_hist[i] = 0;

	// blockIdx.y = view of the batch: per-view camera, per-view slice of every state array
	const int view = blockIdx.y;
	const size_t voff = (size_t)view * P;
	const float* viewmatrix = viewmatrices + 16 * view;
	const float* projmatrix = projmatrices + 16 * view;
	const float* cam_pos = cam_positions + 3 * view;
	int* radii = radii_all + voff;
	float2* means2D = means2D_all + voff;
	float* depths = depths_all + voff;
	float* rec = rec_all + voff * REC_FLOATS;
	float4* conic_opacity = conic_opacity_all + voff;
	uint8_t* clamped = clamped_all + voff * 3;
	uint32_t* tiles_touched = tiles_touched_all + voff;
	ushort4* rects = rect_all + voff;
	uint32_t* tile_counts = tile_counts_all + (size_t)view * grid.x * grid.y;

	const int base = blockIdx.x * PRE_THREADS;
	const int cnt = min(PRE_THREADS, P - base);
	stage_floats(means3D + (size_t)base * 3, s_xyz, cnt * 3);
	if (scales) stage_floats(scales + (size_t)base * 3, s_scl, cnt * 3);
	if (threadIdx.x < 16) { s_vm[threadIdx.x] = viewmatrix[threadIdx.x]; s_pm[threadIdx.x] = projmatrix[threadIdx.x]; }
	if (threadIdx.x < 3) s_cam[threadIdx.x] = cam_pos[threadIdx.x];
	__syncthreads();

	const int idx = base + threadIdx.x;
	if (idx < P) preprocess_one(idx, P, D, M, s_xyz, s_scl, s_vm, s_pm, s_cam, scales != nullptr, scale_modifier, rotations, opacities, shs,
	                            cov3D_precomp, colors_precomp, v2g_precomp, W, H, tan_fovx, tan_fovy, focal_x, focal_y, kernel_size,
	                            radii, means2D, depths, rec, conic_opacity, clamped, grid, tiles_touched, rects,
	                            use_hist ? s_hist : nullptr, tile_counts, prefiltered);
	if (use_hist) {
		__syncthreads();
		for (int i = threadIdx.x; i < T; i += PRE_THREADS) {
			const uint32_t c = s_hist[i];
			if (c) atomicAdd(&tile_counts[i], c);
		}
	}
}

// One Gaussian of one view (the body of the reference's preprocessCUDA, forward.cu:283-404).
__device__ __forceinline__ void preprocess_one(const int idx, int P, int D, int M, const float* s_xyz, const float* s_scl,
	const float* s_vm, const float* s_pm, const float* s_cam, const bool has_scales, const float scale_modifier,
	const float* __restrict__ rotations, const float* __restrict__ opacities, const float* __restrict__ shs,
	const float* __restrict__ cov3D_precomp, const float* __restrict__ colors_precomp, const float* __restrict__ v2g_precomp,
	const int W, const int H, const float tan_fovx, const float tan_fovy, const float focal_x, const float focal_y,
	const float kernel_size, int* __restrict__ radii, float2* __restrict__ means2D, float* __restrict__ depths,
	float* __restrict__ rec, float4* __restrict__ conic_opacity, uint8_t* __restrict__ clamped, const dim3 grid,
	uint32_t* __restrict__ tiles_touched, ushort4* __restrict__ rects, uint32_t* s_hist, uint32_t* __restrict__ tile_counts,
	bool prefiltered)
{
	// Not visible until proven otherwise (forward.cu:317-320).
	radii[idx] = 0;
	tiles_touched[idx] = 0;

	const V3 p_orig = { s_xyz[3 * threadIdx.x], s_xyz[3 * threadIdx.x + 1], s_xyz[3 * threadIdx.x + 2] };
	const float* vm = s_vm;
	const float* pm = s_pm;

	// Near-plane cull only (auxiliary.h:177-202).
	float4 p_hom = { pm[0] * p_orig.x + pm[4] * p_orig.y + pm[8] * p_orig.z + pm[12],
	                 pm[1] * p_orig.x + pm[5] * p_orig.y + pm[9] * p_orig.z + pm[13],
	                 pm[2] * p_orig.x + pm[6] * p_orig.y + pm[10] * p_orig.z + pm[14],
	                 pm[3] * p_orig.x + pm[7] * p_orig.y + pm[11] * p_orig.z + pm[15] };
	float p_w = 1.0f / (p_hom.w + 0.0000001f);
	V3 p_proj = { p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w };
	V3 p_view = { vm[0] * p_orig.x + vm[4] * p_orig.y + vm[8] * p_orig.z + vm[12],
	              vm[1] * p_orig.x + vm[5] * p_orig.y + vm[9] * p_orig.z + vm[13],
	              vm[2] * p_orig.x + vm[6] * p_orig.y + vm[10] * p_orig.z + vm[14] };
	if (p_view.z <= 0.2f) {
		if (prefiltered) {
			printf("Point is filtered although prefiltered is set. This shouldn't happen!");
			__trap();
		}
		return;
	}

	V3 scale = { 0.f, 0.f, 0.f };
	float4 rot = { 0.f, 0.f, 0.f, 0.f };
	if (has_scales) scale = { s_scl[3 * threadIdx.x], s_scl[3 * threadIdx.x + 1], s_scl[3 * threadIdx.x + 2] };
	if (rotations) {
		if ((reinterpret_cast<uintptr_t>(rotations) & 15) == 0) rot = __ldg(reinterpret_cast<const float4*>(rotations) + idx);
		else rot = { rotations[4 * idx], rotations[4 * idx + 1], rotations[4 * idx + 2], rotations[4 * idx + 3] };
	}

	float cov3D[6];
	if (cov3D_precomp != nullptr) {
#pragma unroll
		for (int k = 0; k < 6; k++) cov3D[k] = cov3D_precomp[(size_t)idx * 6 + k];
	} else {
		cov3d_from_scale_rot(scale, scale_modifier, rot, cov3D);
	}

	float4 cov = cov2d_ewa(p_orig, focal_x, focal_y, tan_fovx, tan_fovy, kernel_size, cov3D, vm);

	float det = (cov.x * cov.z - cov.y * cov.y);
	if (det == 0.0f) return;
	float det_inv = 1.f / det;
	float3 conic = { cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv };

	float mid = 0.5f * (cov.x + cov.z);
	float lambda1 = mid + sqrt(max(0.1f, mid * mid - det));
	float lambda2 = mid - sqrt(max(0.1f, mid * mid - det));
	float my_radius = ceil(3.f * sqrt(max(lambda1, lambda2)));
	float2 point_image = { ndc_to_pix(p_proj.x, W), ndc_to_pix(p_proj.y, H) };

	// Tile rectangle (auxiliary.h:64-74): truncation toward zero, then clamp to the grid.
	const int max_radius = (int)my_radius;
	uint2 rect_min = { min(grid.x, max((int)0, (int)((point_image.x - max_radius) / TILE_X))),
	                   min(grid.y, max((int)0, (int)((point_image.y - max_radius) / TILE_Y))) };
	uint2 rect_max = { min(grid.x, max((int)0, (int)((point_image.x + max_radius + TILE_X - 1) / TILE_X))),
	                   min(grid.y, max((int)0, (int)((point_image.y + max_radius + TILE_Y - 1) / TILE_Y))) };
	if ((rect_max.x - rect_min.x) * (rect_max.y - rect_min.y) == 0) return;

	float* r = rec + (size_t)idx * REC_FLOATS;
	float rgb0, rgb1, rgb2;
	if (colors_precomp == nullptr) {
		const V3 campos = { s_cam[0], s_cam[1], s_cam[2] };
		const float* sh = shs + (size_t)idx * M * 3;
		float lo[4][3] = { { 0.f, 0.f, 0.f }, { 0.f, 0.f, 0.f }, { 0.f, 0.f, 0.f }, { 0.f, 0.f, 0.f } };
		if ((M & 3) == 0 && (reinterpret_cast<uintptr_t>(shs) & 15) == 0) {
			// 12 floats = coefficients 0..3 as three 128-bit loads
			const float4* s4 = reinterpret_cast<const float4*>(sh);
			const float4 a = __ldg(s4), b = __ldg(s4 + 1), d = __ldg(s4 + 2);
			lo[0][0] = a.x; lo[0][1] = a.y; lo[0][2] = a.z; lo[1][0] = a.w;
			lo[1][1] = b.x; lo[1][2] = b.y; lo[2][0] = b.z; lo[2][1] = b.w;
			lo[2][2] = d.x; lo[3][0] = d.y; lo[3][1] = d.z; lo[3][2] = d.w;
		} else {
			const int nlo = min(M, 4);
#pragma unroll
			for (int k = 0; k < 4; k++)
				if (k < nlo) { lo[k][0] = sh[3 * k]; lo[k][1] = sh[3 * k + 1]; lo[k][2] = sh[3 * k + 2]; }
		}
		uint8_t cl[3];
		const V3 c = sh_to_rgb(D, p_orig, campos, lo, sh, cl);
		clamped[3 * (size_t)idx + 0] = cl[0];
		clamped[3 * (size_t)idx + 1] = cl[1];
		clamped[3 * (size_t)idx + 2] = cl[2];
		rgb0 = c.x; rgb1 = c.y; rgb2 = c.z;
	} else {
		rgb0 = colors_precomp[3 * (size_t)idx + 0];
		rgb1 = colors_precomp[3 * (size_t)idx + 1];
		rgb2 = colors_precomp[3 * (size_t)idx + 2];
	}

	depths[idx] = p_view.z;
	radii[idx] = my_radius;
	means2D[idx] = point_image;
	const float w = opacities[idx] * cov.w;
	conic_opacity[idx] = { conic.x, conic.y, conic.z, w };
	tiles_touched[idx] = (rect_max.y - rect_min.y) * (rect_max.x - rect_min.x);
	// Tile histogram for the bucketed binning (binning.cu): one count per (tile, Gaussian) duplicate.
	rects[idx] = make_ushort4((unsigned short)rect_min.x, (unsigned short)rect_min.y, (unsigned short)rect_max.x,
	                          (unsigned short)rect_max.y);
	uint32_t* hist = s_hist ? s_hist : tile_counts;
	for (uint32_t y = rect_min.y; y < rect_max.y; y++)
		for (uint32_t x = rect_min.x; x < rect_max.x; x++) atomicAdd(&hist[y * grid.x + x], 1u);

	float q[10];
	if (v2g_precomp == nullptr) {
		view2gaussian_quadric(scale, p_orig, rot, vm, q);
	} else {
#pragma unroll
		for (int k = 0; k < 10; k++) q[k] = v2g_precomp[(size_t)idx * 10 + k];
	}

	// Conservative reject threshold for the blend's float32 pre-test (render_fwd.cu):
	// a pair can only pass the reference's `alpha >= 1/255` test if the ray-minimum value
	// mv satisfies  w*exp(-mv/2) >= 1/255  <=>  mv <= 2*ln(255 w).  tau is that bound plus
	// a safety margin; -FLT_MAX when w < 1/255 can never contribute (power is clamped to <= 0).
	float tau;
	if (w < 1.0f / 255.0f) tau = -3.0e38f;
	else tau = 2.0f * logf(255.0f * w) * 1.00001f + 2e-3f;
	if (!(w == w)) tau = 3.0e38f;  // NaN opacity: never pre-reject, let the exact path decide

	float4* r4 = reinterpret_cast<float4*>(r);
	r4[0] = { q[0], q[1], q[2], q[3] };
	r4[1] = { q[4], q[5], q[6], q[7] };
	r4[2] = { q[8], q[9], tau, w };
	r4[3] = { rgb0, rgb1, rgb2, __int_as_float(idx) };
}

__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ vm,
                                    const float* __restrict__ pm, uint8_t* __restrict__ present)
{
	int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= P) return;
	float x = means3D[3 * idx], y = means3D[3 * idx + 1], z = means3D[3 * idx + 2];
	float vz = vm[2] * x + vm[6] * y + vm[10] * z + vm[14];
	present[idx] = !(vz <= 0.2f);
}

}  // namespace

int launch_preprocess(const GofParams& prm, const GofInputs& in, const Frame& f, const GeomState& g,
                      const ImgState& im, int32_t* radii, cudaStream_t s)
{
	const int P = prm.P;
	if (f.grid.x > 65535u || f.grid.y > 65535u) { set_error("image too large: more than 65535 tiles along one axis"); return GOF_EINVAL; }
	GOF_CUDA_CHECK(cudaMemsetAsync(im.tile_counts, 0, (size_t)f.V * f.T * sizeof(uint32_t), s));
	dim3 blocks((P + PRE_THREADS - 1) / PRE_THREADS, f.V);
	preprocess_kernel<<<blocks, PRE_THREADS, 0, s>>>(
		P, prm.D, prm.M, in.means3D, in.scales, prm.scale_modifier, in.rotations, in.opacities, in.shs,
		in.cov3D_precomp, in.colors_precomp, in.view2gaussian_precomp, in.viewmatrix, in.projmatrix,
		in.campos, prm.W, prm.H, prm.tan_fovx, prm.tan_fovy, f.focal_x, f.focal_y, prm.kernel_size,
		radii, g.means2D, g.depths, g.rec, g.conic_opacity, g.clamped, f.grid, g.tiles_touched, g.rect,
		im.tile_counts, prm.prefiltered != 0);
	GOF_CUDA_CHECK(cudaGetLastError());
	return GOF_OK;
}

}  // namespace gof

extern "C" int gof_mark_visible(int32_t P, const float* means3D, const float* viewmatrix,
                                const float* projmatrix, uint8_t* present, gof_stream_t stream)
{
	if (P <= 0) return GOF_OK;
	gof::mark_visible_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, means3D, viewmatrix, projmatrix, present);
	GOF_CUDA_CHECK(cudaGetLastError());
	return GOF_OK;
}
