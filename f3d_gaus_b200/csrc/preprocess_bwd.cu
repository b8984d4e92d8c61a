// preprocess_bwd.cu -- per-Gaussian backward (K10) fused with gradient unpacking/zero-fill.
//
// Replaces BACKWARD::preprocess / preprocessCUDA<3> / computeView2Gaussian_backward / the SH
// backward (RAST/cuda_rasterizer/backward.cu:20-139,381-631,957-1033) and the ten
// torch::zeros fills of RasterizeGaussiansBackwardCUDA (rasterize_points.cu:161-170): every
// element of the nine output tensors is written here exactly once, so the caller can hand in
// uninitialised memory.
//
// Reference behaviours kept on purpose: dL/dmean3D, dL/dscale, dL/drot are ASSIGNED from the
// view2gaussian path (not accumulated); the EWA/cov2D path has no gradient (computeCov2DCUDA
// is disabled, backward.cu:991-1007) so dL/dcov3D == 0; scale_modifier is ignored; Gaussians
// with radii <= 0 get zero gradients.
#include "gof_common.cuh"

namespace gof {

namespace {

__device__ const float kSH_C0 = 0.28209479177387814f;
__device__ const float kSH_C1 = 0.4886025119029199f;
__device__ const float kSH_C2[] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                   -1.0925484305920792f, 0.5462742152960396f};
__device__ const float kSH_C3[] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                   0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                   -0.5900435899266435f};

struct V3 { float x, y, z; };
// d(normalize(v))/dv applied to dv (auxiliary.h:150-161)
__device__ __forceinline__ V3 dnormvdv(V3 v, V3 dv)
{
	float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
	float invsum32 = 1.0f / sqrt(sum2 * sum2 * sum2);
	V3 r;
	r.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
	r.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
	r.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
	return r;
}

// Backward of the 10-float quadric w.r.t. mean / scale / rotation (backward.cu:381-587).
//
// Evaluated in DOUBLE from the float32 inputs.  The map is catastrophically ill-conditioned at F3D-Gaus scales:
// dL/dSinv_k = Rt_k^T dSigma Rt_k + t2_k (dB . Rt_k) + dC t2_k^2 is, pair by pair, Sinv-derivative of a perfect
// square (t2_k + t* Rt_k.r)^2 ~ (3 sigma)^2 assembled from terms of size t2_k^2 ~ 60 -- a cancellation of ~1e5..1e6 --
// and the reference, which evaluates it in float32 (with double only for Sinv), is 3e-2 away from the exact value of
// its own formula.  This kernel is a 15 us HBM-bound per-Gaussian pass, so double arithmetic is free here and
// removes the rounding error of the map altogether: what remains is the float32 noise already present in the
// accumulated dL/dview2gaussian it is applied to (tests: e_ours <= e_ref against the float64 oracle).
struct D3 { double x, y, z; };
struct DM3 { double c[3][3]; };   // column-major: c[col][row]

__device__ __forceinline__ DM3 dm3_mul(const DM3& a, const DM3& b)
{
	DM3 r;
#pragma unroll
	for (int i = 0; i < 3; i++)
#pragma unroll
		for (int j = 0; j < 3; j++)
			r.c[i][j] = a.c[0][j] * b.c[i][0] + a.c[1][j] * b.c[i][1] + a.c[2][j] * b.c[i][2];
	return r;
}
__device__ __forceinline__ DM3 dm3_t(const DM3& a)
{
	DM3 r;
#pragma unroll
	for (int i = 0; i < 3; i++)
#pragma unroll
		for (int j = 0; j < 3; j++) r.c[i][j] = a.c[j][i];
	return r;
}

__device__ void quadric_backward(const V3 scale, const V3 mean, const float4 rot, const float* vm,
                                 const float* dq, D3& dL_dmean, D3& dL_dscale, double* dL_drot)
{
	const double r = rot.x, x = rot.y, y = rot.z, z = rot.w;
	DM3 R;
	R.c[0][0] = 1. - 2. * (y * y + z * z); R.c[0][1] = 2. * (x * y - r * z);      R.c[0][2] = 2. * (x * z + r * y);
	R.c[1][0] = 2. * (x * y + r * z);      R.c[1][1] = 1. - 2. * (x * x + z * z); R.c[1][2] = 2. * (y * z - r * x);
	R.c[2][0] = 2. * (x * z - r * y);      R.c[2][1] = 2. * (y * z + r * x);      R.c[2][2] = 1. - 2. * (x * x + y * y);

	// G2V = W2V * G2W with G2W = [R^T-layout | mean]; only the 3x4 part is needed.
	// G2W column c (c<3) = (R[0][c], R[1][c], R[2][c], 0), column 3 = (mean, 1).
	double G2V[4][3];
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int j = 0; j < 3; j++)
			G2V[c][j] = (double)vm[0 + j] * R.c[0][c] + (double)vm[4 + j] * R.c[1][c] + (double)vm[8 + j] * R.c[2][c];
#pragma unroll
	for (int j = 0; j < 3; j++)
		G2V[3][j] = (double)vm[0 + j] * mean.x + (double)vm[4 + j] * mean.y + (double)vm[8 + j] * mean.z + (double)vm[12 + j];

	DM3 Rt;   // Rt[c][r] = G2V[r][c]
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int rr = 0; rr < 3; rr++) Rt.c[c][rr] = G2V[rr][c];
	const double ta[3] = { G2V[3][0], G2V[3][1], G2V[3][2] };
	double t2a[3];
#pragma unroll
	for (int rr = 0; rr < 3; rr++) t2a[rr] = -(Rt.c[0][rr] * ta[0] + Rt.c[1][rr] * ta[1] + Rt.c[2][rr] * ta[2]);

	const double sc[3] = { scale.x, scale.y, scale.z };
	double Sinv[3];
#pragma unroll
	for (int k = 0; k < 3; k++) Sinv[k] = 1.0 / (sc[k] * sc[k] + 1e-7);
	DM3 SR;
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int rr = 0; rr < 3; rr++) SR.c[c][rr] = Sinv[rr] * Rt.c[c][rr];

	DM3 dSig;
	dSig.c[0][0] = dq[0];        dSig.c[0][1] = 0.5 * dq[1]; dSig.c[0][2] = 0.5 * dq[2];
	dSig.c[1][0] = 0.5 * dq[1];  dSig.c[1][1] = dq[3];       dSig.c[1][2] = 0.5 * dq[4];
	dSig.c[2][0] = 0.5 * dq[2];  dSig.c[2][1] = 0.5 * dq[4]; dSig.c[2][2] = dq[5];
	const double dB[3] = { dq[6], dq[7], dq[8] };
	const double dC = dq[9];

	// dL/dSR = Rt * dSigma + outer(t2, dB)   (outer: column i = t2 * dB[i])
	DM3 dSR = dm3_mul(Rt, dSig);
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int rr = 0; rr < 3; rr++) dSR.c[c][rr] += t2a[rr] * dB[c];
	// dL/dRt = (dSigma * SR^T)^T + diag(Sinv) applied row-wise to dSR
	DM3 dRt = dm3_t(dm3_mul(dSig, dm3_t(SR)));
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int rr = 0; rr < 3; rr++) dRt.c[c][rr] += Sinv[rr] * dSR.c[c][rr];

	double dSinv[3], dt2[3];
#pragma unroll
	for (int rr = 0; rr < 3; rr++) {
		dSinv[rr] = dSR.c[0][rr] * Rt.c[0][rr] + dSR.c[1][rr] * Rt.c[1][rr] + dSR.c[2][rr] * Rt.c[2][rr] + dC * t2a[rr] * t2a[rr];
		dt2[rr] = 2 * t2a[rr] * Sinv[rr] * dC + dB[0] * SR.c[0][rr] + dB[1] * SR.c[1][rr] + dB[2] * SR.c[2][rr];
	}

	dL_dscale.x = -2 / sc[0] * Sinv[0] * dSinv[0];
	dL_dscale.y = -2 / sc[1] * Sinv[1] * dSinv[1];
	dL_dscale.z = -2 / sc[2] * Sinv[2] * dSinv[2];

	// Back through V2G = [G2V_R^T | -G2V_R^T t] to G2V, then through G2V = W2V * G2W.
	DM3 dG2V_R = dm3_t(dRt);
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int rr = 0; rr < 3; rr++) dG2V_R.c[c][rr] += -dt2[c] * ta[rr];
	// dL/dG2V_t = (-dt2) as a row vector times G2V_R^T (= Rt)
	double dG2V_t[3];
#pragma unroll
	for (int c = 0; c < 3; c++)
		dG2V_t[c] = Rt.c[c][0] * -dt2[0] + Rt.c[c][1] * -dt2[1] + Rt.c[c][2] * -dt2[2];

	// dL/dG2W = W2V^T * dL/dG2V (4x4, last row of dG2V is 0): element [c][r] = sum_k W2V[r][k] dG2V[c][k]
	double dG2W[4][3];
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int rr = 0; rr < 3; rr++)
			dG2W[c][rr] = (double)vm[4 * rr + 0] * dG2V_R.c[c][0] + (double)vm[4 * rr + 1] * dG2V_R.c[c][1] + (double)vm[4 * rr + 2] * dG2V_R.c[c][2];
#pragma unroll
	for (int rr = 0; rr < 3; rr++)
		dG2W[3][rr] = (double)vm[4 * rr + 0] * dG2V_t[0] + (double)vm[4 * rr + 1] * dG2V_t[1] + (double)vm[4 * rr + 2] * dG2V_t[2];

	dL_dmean = { dG2W[3][0], dG2W[3][1], dG2W[3][2] };

	// quaternion gradient from dL/dMt = the 3x3 block of dL/dG2W (backward.cu:575-586)
#define MT(a, b) dG2W[a][b]
	dL_drot[0] = 2 * z * (MT(0, 1) - MT(1, 0)) + 2 * y * (MT(2, 0) - MT(0, 2)) + 2 * x * (MT(1, 2) - MT(2, 1));
	dL_drot[1] = 2 * y * (MT(1, 0) + MT(0, 1)) + 2 * z * (MT(2, 0) + MT(0, 2)) + 2 * r * (MT(1, 2) - MT(2, 1)) - 4 * x * (MT(2, 2) + MT(1, 1));
	dL_drot[2] = 2 * x * (MT(1, 0) + MT(0, 1)) + 2 * r * (MT(2, 0) - MT(0, 2)) + 2 * z * (MT(1, 2) + MT(2, 1)) - 4 * y * (MT(2, 2) + MT(0, 0));
	dL_drot[3] = 2 * r * (MT(0, 1) - MT(1, 0)) + 2 * x * (MT(2, 0) + MT(0, 2)) + 2 * y * (MT(1, 2) + MT(2, 1)) - 4 * z * (MT(1, 1) + MT(0, 0));
#undef MT
}

// SH backward (backward.cu:20-139): ADDS this view's dL/dsh of the Gaussian (the caller zero-initialises the
// block; with one view that is the reference's plain assignment), returns the mean gradient contribution
// through the view direction.
__device__ V3 sh_backward(int deg, int M, const V3 pos, const V3 campos, const float* sh, const uint8_t* clamped3,
                          const float* dL_dcolor3, float* dL_dsh)
{
	const V3 dir_orig = { pos.x - campos.x, pos.y - campos.y, pos.z - campos.z };
	const float len = sqrt(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
	const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;
	float dRGB[3];
#pragma unroll
	for (int ch = 0; ch < 3; ch++) dRGB[ch] = dL_dcolor3[ch] * (clamped3[ch] ? 0.f : 1.f);
	float ddx = 0, ddy = 0, ddz = 0;   // dL/ddir
#define SHV(k, ch) sh[3 * (k) + (ch)]
#define DSH(k, ch) dL_dsh[3 * (k) + (ch)]
	for (int ch = 0; ch < 3; ch++) {
		const float g = dRGB[ch];
		float dx = 0, dy = 0, dz = 0;   // dRGB_ch/d(x,y,z)
		DSH(0, ch) += kSH_C0 * g;
		if (deg > 0) {
			DSH(1, ch) += -kSH_C1 * y * g;
			DSH(2, ch) += kSH_C1 * z * g;
			DSH(3, ch) += -kSH_C1 * x * g;
			dx = -kSH_C1 * SHV(3, ch);
			dy = -kSH_C1 * SHV(1, ch);
			dz = kSH_C1 * SHV(2, ch);
			if (deg > 1) {
				const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
				DSH(4, ch) += kSH_C2[0] * xy * g;
				DSH(5, ch) += kSH_C2[1] * yz * g;
				DSH(6, ch) += kSH_C2[2] * (2.f * zz - xx - yy) * g;
				DSH(7, ch) += kSH_C2[3] * xz * g;
				DSH(8, ch) += kSH_C2[4] * (xx - yy) * g;
				dx += kSH_C2[0] * y * SHV(4, ch) + kSH_C2[2] * 2.f * -x * SHV(6, ch) + kSH_C2[3] * z * SHV(7, ch) + kSH_C2[4] * 2.f * x * SHV(8, ch);
				dy += kSH_C2[0] * x * SHV(4, ch) + kSH_C2[1] * z * SHV(5, ch) + kSH_C2[2] * 2.f * -y * SHV(6, ch) + kSH_C2[4] * 2.f * -y * SHV(8, ch);
				dz += kSH_C2[1] * y * SHV(5, ch) + kSH_C2[2] * 2.f * 2.f * z * SHV(6, ch) + kSH_C2[3] * x * SHV(7, ch);
				if (deg > 2) {
					DSH(9, ch) += kSH_C3[0] * y * (3.f * xx - yy) * g;
					DSH(10, ch) += kSH_C3[1] * xy * z * g;
					DSH(11, ch) += kSH_C3[2] * y * (4.f * zz - xx - yy) * g;
					DSH(12, ch) += kSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy) * g;
					DSH(13, ch) += kSH_C3[4] * x * (4.f * zz - xx - yy) * g;
					DSH(14, ch) += kSH_C3[5] * z * (xx - yy) * g;
					DSH(15, ch) += kSH_C3[6] * x * (xx - 3.f * yy) * g;
					dx += kSH_C3[0] * SHV(9, ch) * 3.f * 2.f * xy + kSH_C3[1] * SHV(10, ch) * yz + kSH_C3[2] * SHV(11, ch) * -2.f * xy +
					      kSH_C3[3] * SHV(12, ch) * -3.f * 2.f * xz + kSH_C3[4] * SHV(13, ch) * (-3.f * xx + 4.f * zz - yy) +
					      kSH_C3[5] * SHV(14, ch) * 2.f * xz + kSH_C3[6] * SHV(15, ch) * 3.f * (xx - yy);
					dy += kSH_C3[0] * SHV(9, ch) * 3.f * (xx - yy) + kSH_C3[1] * SHV(10, ch) * xz + kSH_C3[2] * SHV(11, ch) * (-3.f * yy + 4.f * zz - xx) +
					      kSH_C3[3] * SHV(12, ch) * -3.f * 2.f * yz + kSH_C3[4] * SHV(13, ch) * -2.f * xy +
					      kSH_C3[5] * SHV(14, ch) * -2.f * yz + kSH_C3[6] * SHV(15, ch) * -3.f * 2.f * xy;
					dz += kSH_C3[1] * SHV(10, ch) * xy + kSH_C3[2] * SHV(11, ch) * 4.f * 2.f * yz + kSH_C3[3] * SHV(12, ch) * 3.f * (2.f * zz - xx - yy) +
					      kSH_C3[4] * SHV(13, ch) * 4.f * 2.f * xz + kSH_C3[5] * SHV(14, ch) * (xx - yy);
				}
			}
		}
		ddx += dx * g; ddy += dy * g; ddz += dz * g;
	}
#undef SHV
#undef DSH
	// coefficients above the active degree keep the zero the caller initialised them with
	return dnormvdv(dir_orig, V3{ ddx, ddy, ddz });
}

// One thread per Gaussian; loops over the V views of a batch and SUMS their gradients (V = 1: the reference's
// per-frame backward).  Per view: gacc[v], radii[v], clamped[v], viewmatrix[v], campos[v].
// 128 threads per CTA: the double-precision quadric chain needs ~220 registers per thread, so a 256-thread CTA would
// be alone on its SM; 128-thread CTAs sit two per SM and spread 65,536 Gaussians over 512 CTAs instead of 256.
#ifndef GOF_PRE_BWD_MIN_CTAS
#define GOF_PRE_BWD_MIN_CTAS 2
#endif
constexpr int PRE_BWD_THREADS = 128;
__global__ void __launch_bounds__(PRE_BWD_THREADS, GOF_PRE_BWD_MIN_CTAS)
preprocess_bwd_kernel(int P, int V, int D, int M, const float* __restrict__ means3D, const int* __restrict__ radii_all,
                      const float* __restrict__ shs, const uint8_t* __restrict__ clamped_all,
                      const float* __restrict__ scales, const float* __restrict__ rotations,
                      const float* __restrict__ viewmatrices, const float* __restrict__ campos_all,
                      const float* __restrict__ gacc_all,
                      float* __restrict__ dL_dmeans2D, float* __restrict__ dL_dcolors, float* __restrict__ dL_dopacity,
                      float* __restrict__ dL_dmeans3D, float* __restrict__ dL_dcov3D, float* __restrict__ dL_dsh,
                      float* __restrict__ dL_dscales, float* __restrict__ dL_drot, float* __restrict__ dL_dv2g)
{
	pdl_trigger();
	pdl_wait();                    // the accumulators come from the backward blend
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= P) return;

	const V3 mean = { means3D[3 * (size_t)idx], means3D[3 * (size_t)idx + 1], means3D[3 * (size_t)idx + 2] };
	const bool has_geom = scales != nullptr && rotations != nullptr;
	V3 scale = { 0.f, 0.f, 0.f };
	float4 rot = { 0.f, 0.f, 0.f, 0.f };
	if (has_geom) {
		scale = { scales[3 * (size_t)idx], scales[3 * (size_t)idx + 1], scales[3 * (size_t)idx + 2] };
		rot = { rotations[4 * (size_t)idx], rotations[4 * (size_t)idx + 1], rotations[4 * (size_t)idx + 2],
		        rotations[4 * (size_t)idx + 3] };
	}
	float* dsh = (dL_dsh != nullptr && M > 0) ? dL_dsh + (size_t)idx * M * 3 : nullptr;
	if (dsh) for (int k = 0; k < 3 * M; k++) dsh[k] = 0.0f;

	float sq[10] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 }, scol[3] = { 0, 0, 0 }, sop = 0.f, sm2[3] = { 0, 0, 0 };
	// mean / scale / rotation gradients are summed over the views in double (each view's term is exact to ~1e-16)
	D3 dmean = { 0., 0., 0. }, dscale = { 0., 0., 0. };
	double drot[4] = { 0., 0., 0., 0. };
	for (int v = 0; v < V; v++) {
		const float4* ga = reinterpret_cast<const float4*>(gacc_all + ((size_t)v * P + idx) * GACC_FLOATS);
		const float4 g0 = ga[0], g1 = ga[1], g2 = ga[2], g3 = ga[3], g4 = ga[4];
		const float dq[10] = { g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w, g2.x, g2.y };
		const float dcol[3] = { g2.z, g2.w, g3.x };
#pragma unroll
		for (int k = 0; k < 10; k++) sq[k] += dq[k];
		scol[0] += dcol[0]; scol[1] += dcol[1]; scol[2] += dcol[2];
		sop += g3.y;
		sm2[0] += g3.z; sm2[1] += g3.w; sm2[2] += g4.x;
		const bool visible = radii_all[(size_t)v * P + idx] > 0;
		if (!visible) continue;
		if (has_geom) {
			D3 dm, ds;
			double dr[4];
			quadric_backward(scale, mean, rot, viewmatrices + 16 * v, dq, dm, ds, dr);
			dmean.x += dm.x; dmean.y += dm.y; dmean.z += dm.z;
			dscale.x += ds.x; dscale.y += ds.y; dscale.z += ds.z;
			drot[0] += dr[0]; drot[1] += dr[1]; drot[2] += dr[2]; drot[3] += dr[3];
		}
		if (dsh && shs != nullptr) {
			const V3 cam = { campos_all[3 * v], campos_all[3 * v + 1], campos_all[3 * v + 2] };
			const V3 add = sh_backward(D, M, mean, cam, shs + (size_t)idx * M * 3, clamped_all + 3 * ((size_t)v * P + idx), dcol, dsh);
			dmean.x += add.x; dmean.y += add.y; dmean.z += add.z;
		}
	}

#pragma unroll
	for (int k = 0; k < 10; k++) dL_dv2g[(size_t)idx * 10 + k] = sq[k];
	dL_dcolors[3 * (size_t)idx + 0] = scol[0];
	dL_dcolors[3 * (size_t)idx + 1] = scol[1];
	dL_dcolors[3 * (size_t)idx + 2] = scol[2];
	dL_dopacity[idx] = sop;
	dL_dmeans2D[3 * (size_t)idx + 0] = sm2[0];
	dL_dmeans2D[3 * (size_t)idx + 1] = sm2[1];
	dL_dmeans2D[3 * (size_t)idx + 2] = sm2[2];
#pragma unroll
	for (int k = 0; k < 6; k++) dL_dcov3D[(size_t)idx * 6 + k] = 0.0f;
	dL_dmeans3D[3 * (size_t)idx + 0] = (float)dmean.x;
	dL_dmeans3D[3 * (size_t)idx + 1] = (float)dmean.y;
	dL_dmeans3D[3 * (size_t)idx + 2] = (float)dmean.z;
	dL_dscales[3 * (size_t)idx + 0] = (float)dscale.x;
	dL_dscales[3 * (size_t)idx + 1] = (float)dscale.y;
	dL_dscales[3 * (size_t)idx + 2] = (float)dscale.z;
	dL_drot[4 * (size_t)idx + 0] = (float)drot[0];
	dL_drot[4 * (size_t)idx + 1] = (float)drot[1];
	dL_drot[4 * (size_t)idx + 2] = (float)drot[2];
	dL_drot[4 * (size_t)idx + 3] = (float)drot[3];
}

}  // namespace

int launch_preprocess_bwd(const GofParams& prm, const GofInputs& in, int V, const GeomState& g,
                          const int32_t* radii, const float* gacc, const GofGrads& grads,
                          cudaStream_t s)
{
	const int P = prm.P;
	GOF_CUDA_CHECK(launch_chained(PDL_PRE_BWD, preprocess_bwd_kernel, dim3((P + PRE_BWD_THREADS - 1) / PRE_BWD_THREADS), dim3(PRE_BWD_THREADS), 0, s, P, V, prm.D, prm.M, in.means3D, radii,
		in.shs, g.clamped, in.scales, in.rotations, in.viewmatrix, in.campos, gacc, grads.dL_dmeans2D, grads.dL_dcolors,
		grads.dL_dopacity, grads.dL_dmeans3D, grads.dL_dcov3D, grads.dL_dsh, grads.dL_dscales,
		grads.dL_drotations, grads.dL_dview2gaussian));
	return GOF_OK;
}

}  // namespace gof
