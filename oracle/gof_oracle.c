/* gof_oracle.c -- CPU restatement of the reference GOF rasterizer's algorithm.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under f3d_gaus_b200/ may import, link or execute this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it, as the checker
 * (and as the reported CPU baseline), never as the product.
 *
 * What it follows (RAST = src/gaussian-splatting/submodules/diff-gof-rasterization):
 *   oracle_preprocess          RAST/cuda_rasterizer/forward.cu:20-404, auxiliary.h:59-74,177-202
 *   oracle_binning             RAST/cuda_rasterizer/rasterizer_impl.cu:35-50,70-111,149-171,332-373
 *   oracle_render_forward      RAST/cuda_rasterizer/forward.cu:409-612
 *   oracle_render_backward     RAST/cuda_rasterizer/backward.cu:634-955
 *   oracle_integrate           RAST/cuda_rasterizer/forward.cu:722-766,803-1218, rasterizer_impl.cu:113-144,530-792
 *   oracle_preprocess_backward RAST/cuda_rasterizer/backward.cu:20-139,381-631
 *
 * Pinning: checked against tests/golden/*.npz, which hold the outputs of the UNMODIFIED reference
 * CUDA build (oracle/_ref/libgof_ref.so) run on a B200 (tests/golden/make_golden.py).  Integer work
 * (keys, stable sort, ranges, offsets) is bit-exact given the same float state.  Float work is
 * IEEE float/double evaluated in the reference's operation order but without reproducing nvcc's
 * FMA contraction or the GPU's approximate ex2, so float results agree to a few ulp per stage --
 * tests feed each stage the golden state of the previous one and compare within the north-star
 * tolerances (1e-4 abs forward, 1e-3 rel backward).  The bit-exact float checker is oracle/_ref.
 *
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).  All pointers are host pointers.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BLOCK_X 16
#define BLOCK_Y 16
#define NEAR_PLANE 0.2
#define FAR_PLANE 100.0

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                              -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                              -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

int oracle_num_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

/* ---- tiny column-major matrix helpers (the reference's GLM convention: m[col][row]) ---- */
typedef struct { float c[3][3]; } m3;

static m3 m3_mul(m3 a, m3 b)
{
	m3 r;
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++)
			r.c[i][j] = a.c[0][j] * b.c[i][0] + a.c[1][j] * b.c[i][1] + a.c[2][j] * b.c[i][2];
	return r;
}
static m3 m3_t(m3 a)
{
	m3 r;
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) r.c[i][j] = a.c[j][i];
	return r;
}
static m3 quat_m3(const float* q)
{
	const float r = q[0], x = q[1], y = q[2], z = q[3];
	m3 R;
	R.c[0][0] = 1.f - 2.f * (y * y + z * z); R.c[0][1] = 2.f * (x * y - r * z); R.c[0][2] = 2.f * (x * z + r * y);
	R.c[1][0] = 2.f * (x * y + r * z); R.c[1][1] = 1.f - 2.f * (x * x + z * z); R.c[1][2] = 2.f * (y * z - r * x);
	R.c[2][0] = 2.f * (x * z - r * y); R.c[2][1] = 2.f * (y * z + r * x); R.c[2][2] = 1.f - 2.f * (x * x + y * y);
	return R;
}

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

static void get_rect(float px, float py, int max_radius, int gx, int gy, int* x0, int* y0, int* x1, int* y1)
{
	/* auxiliary.h:64-74: float division, C truncation toward zero, clamp to [0, grid] */
	*x0 = imin(gx, imax(0, (int)((px - max_radius) / BLOCK_X)));
	*y0 = imin(gy, imax(0, (int)((py - max_radius) / BLOCK_Y)));
	*x1 = imin(gx, imax(0, (int)((px + max_radius + BLOCK_X - 1) / BLOCK_X)));
	*y1 = imin(gy, imax(0, (int)((py + max_radius + BLOCK_Y - 1) / BLOCK_Y)));
}

/* Gaussian-to-view rotation block Rt[c][r] and translation of G2V = W2V * G2W (forward.cu:185-221). */
static void g2v(const float* q, const float* mean, const float* vm, m3* Rt, float* t)
{
	m3 R = quat_m3(q);
	float G2V[4][3];
	for (int c = 0; c < 3; c++)
		for (int j = 0; j < 3; j++)
			G2V[c][j] = vm[0 + j] * R.c[0][c] + vm[4 + j] * R.c[1][c] + vm[8 + j] * R.c[2][c] + vm[12 + j] * 0.0f;
	for (int j = 0; j < 3; j++)
		G2V[3][j] = vm[0 + j] * mean[0] + vm[4 + j] * mean[1] + vm[8 + j] * mean[2] + vm[12 + j];
	for (int c = 0; c < 3; c++)
		for (int r = 0; r < 3; r++) Rt->c[c][r] = G2V[r][c];
	t[0] = G2V[3][0]; t[1] = G2V[3][1]; t[2] = G2V[3][2];
}

/* ------------------------------------------------------------------ preprocess -------- */
/* Outputs are only defined where radii[i] > 0 (the reference leaves the rest unwritten). */
void oracle_preprocess(int P, int D, int M, const float* means3D, const float* scales, float scale_modifier,
                       const float* rotations, const float* opacities, const float* shs, const float* cov3D_precomp,
                       const float* colors_precomp, const float* v2g_precomp, const float* vm, const float* pm,
                       const float* campos, int W, int H, float tan_fovx, float tan_fovy, float kernel_size,
                       int32_t* radii, float* means2D, float* depths, float* v2g, float* rgb, float* conic_opacity,
                       uint32_t* tiles_touched, uint8_t* clamped)
{
	const float focal_y = H / (2.0f * tan_fovy);
	const float focal_x = W / (2.0f * tan_fovx);
	const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
#pragma omp parallel for schedule(static)
	for (int idx = 0; idx < P; idx++) {
		radii[idx] = 0;
		tiles_touched[idx] = 0;
		const float px = means3D[3 * idx], py = means3D[3 * idx + 1], pz = means3D[3 * idx + 2];
		float hom[4];
		for (int k = 0; k < 4; k++) hom[k] = pm[k] * px + pm[4 + k] * py + pm[8 + k] * pz + pm[12 + k];
		const float p_w = 1.0f / (hom[3] + 0.0000001f);
		const float projx = hom[0] * p_w, projy = hom[1] * p_w;
		float view[3];
		for (int k = 0; k < 3; k++) view[k] = vm[k] * px + vm[4 + k] * py + vm[8 + k] * pz + vm[12 + k];
		if (view[2] <= 0.2f) continue;   /* near-plane cull only (auxiliary.h:192) */

		float cov3D[6];
		if (cov3D_precomp) {
			for (int k = 0; k < 6; k++) cov3D[k] = cov3D_precomp[6 * idx + k];
		} else {
			m3 S;
			memset(&S, 0, sizeof(S));
			S.c[0][0] = scale_modifier * scales[3 * idx];
			S.c[1][1] = scale_modifier * scales[3 * idx + 1];
			S.c[2][2] = scale_modifier * scales[3 * idx + 2];
			m3 Mm = m3_mul(S, quat_m3(rotations + 4 * idx));
			m3 Sig = m3_mul(m3_t(Mm), Mm);
			cov3D[0] = Sig.c[0][0]; cov3D[1] = Sig.c[0][1]; cov3D[2] = Sig.c[0][2];
			cov3D[3] = Sig.c[1][1]; cov3D[4] = Sig.c[1][2]; cov3D[5] = Sig.c[2][2];
		}

		/* EWA cov2D + opacity coefficient (forward.cu:74-124) */
		float t[3] = { view[0], view[1], view[2] };
		const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
		const float txtz = t[0] / t[2], tytz = t[1] / t[2];
		t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
		t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
		m3 J;
		memset(&J, 0, sizeof(J));
		J.c[0][0] = focal_x / t[2]; J.c[0][2] = -(focal_x * t[0]) / (t[2] * t[2]);
		J.c[1][1] = focal_y / t[2]; J.c[1][2] = -(focal_y * t[1]) / (t[2] * t[2]);
		m3 Wm;
		for (int c = 0; c < 3; c++)
			for (int r = 0; r < 3; r++) Wm.c[c][r] = vm[4 * r + c];
		m3 T = m3_mul(Wm, J);
		m3 Vrk;
		Vrk.c[0][0] = cov3D[0]; Vrk.c[0][1] = cov3D[1]; Vrk.c[0][2] = cov3D[2];
		Vrk.c[1][0] = cov3D[1]; Vrk.c[1][1] = cov3D[3]; Vrk.c[1][2] = cov3D[4];
		Vrk.c[2][0] = cov3D[2]; Vrk.c[2][1] = cov3D[4]; Vrk.c[2][2] = cov3D[5];
		m3 cov = m3_mul(m3_mul(m3_t(T), m3_t(Vrk)), T);
		const float c00 = cov.c[0][0], c01 = cov.c[0][1], c11 = cov.c[1][1];
		const float det_0 = (float)fmax(1e-6, (double)(c00 * c11 - c01 * c01));
		const float det_1 = (float)fmax(1e-6, (double)((c00 + kernel_size) * (c11 + kernel_size) - c01 * c01));
		float coef = (float)sqrt((double)det_0 / ((double)det_1 + 1e-6) + 1e-6);
		if ((double)det_0 <= 1e-6 || (double)det_1 <= 1e-6) coef = 0.0f;
		const float cx = c00 + kernel_size, cy = c01, cz = c11 + kernel_size;

		const float det = cx * cz - cy * cy;
		if (det == 0.0f) continue;
		const float det_inv = 1.f / det;
		const float conic[3] = { cz * det_inv, -cy * det_inv, cx * det_inv };
		const float mid = 0.5f * (cx + cz);
		const float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
		const float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
		const float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
		const float pix_x = (float)((((double)projx + 1.0) * W - 1.0) * 0.5);   /* ndc2Pix in double */
		const float pix_y = (float)((((double)projy + 1.0) * H - 1.0) * 0.5);
		int x0, y0, x1, y1;
		get_rect(pix_x, pix_y, (int)my_radius, gx, gy, &x0, &y0, &x1, &y1);
		if ((x1 - x0) * (y1 - y0) == 0) continue;

		if (!colors_precomp) {
			const float* sh = shs + (size_t)idx * M * 3;
			float dir[3] = { px - campos[0], py - campos[1], pz - campos[2] };
			const float len = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
			const float x = dir[0] / len, y = dir[1] / len, z = dir[2] / len;
			for (int ch = 0; ch < 3; ch++) {
#define SHV(k) sh[3 * (k) + ch]
				float result = SH_C0 * SHV(0);
				if (D > 0) {
					result = result - SH_C1 * y * SHV(1) + SH_C1 * z * SHV(2) - SH_C1 * x * SHV(3);
					if (D > 1) {
						const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
						result = result + SH_C2[0] * xy * SHV(4) + SH_C2[1] * yz * SHV(5) +
						         SH_C2[2] * (2.0f * zz - xx - yy) * SHV(6) + SH_C2[3] * xz * SHV(7) +
						         SH_C2[4] * (xx - yy) * SHV(8);
						if (D > 2) {
							result = result + SH_C3[0] * y * (3.0f * xx - yy) * SHV(9) + SH_C3[1] * xy * z * SHV(10) +
							         SH_C3[2] * y * (4.0f * zz - xx - yy) * SHV(11) +
							         SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * SHV(12) +
							         SH_C3[4] * x * (4.0f * zz - xx - yy) * SHV(13) + SH_C3[5] * z * (xx - yy) * SHV(14) +
							         SH_C3[6] * x * (xx - 3.0f * yy) * SHV(15);
						}
					}
				}
#undef SHV
				result += 0.5f;
				clamped[3 * idx + ch] = (result < 0);
				rgb[3 * idx + ch] = result < 0 ? 0.0f : result;
			}
		}

		depths[idx] = view[2];
		radii[idx] = (int32_t)my_radius;
		means2D[2 * idx] = pix_x;
		means2D[2 * idx + 1] = pix_y;
		conic_opacity[4 * idx + 0] = conic[0];
		conic_opacity[4 * idx + 1] = conic[1];
		conic_opacity[4 * idx + 2] = conic[2];
		conic_opacity[4 * idx + 3] = opacities[idx] * coef;
		tiles_touched[idx] = (uint32_t)((y1 - y0) * (x1 - x0));

		if (!v2g_precomp) {
			/* view2gaussian quadric (forward.cu:168-279); scale_modifier deliberately not applied */
			m3 Rt;
			float tt[3];
			g2v(rotations + 4 * idx, means3D + 3 * idx, vm, &Rt, tt);
			float t2[3];
			for (int r = 0; r < 3; r++) t2[r] = -Rt.c[0][r] * tt[0] + -Rt.c[1][r] * tt[1] + -Rt.c[2][r] * tt[2];
			double Sinv[3];
			for (int k = 0; k < 3; k++) Sinv[k] = 1.0 / ((double)scales[3 * idx + k] * scales[3 * idx + k] + 1e-7);
			const double C = (double)(t2[0] * t2[0]) * Sinv[0] + (double)(t2[1] * t2[1]) * Sinv[1] + (double)(t2[2] * t2[2]) * Sinv[2];
			m3 SR;
			for (int c = 0; c < 3; c++)
				for (int r = 0; r < 3; r++) SR.c[c][r] = (float)(Sinv[r] * (double)Rt.c[c][r]);
			float B[3];
			for (int c = 0; c < 3; c++) B[c] = SR.c[c][0] * t2[0] + SR.c[c][1] * t2[1] + SR.c[c][2] * t2[2];
			m3 Sig = m3_mul(m3_t(Rt), SR);
			float* o = v2g + 10 * (size_t)idx;
			o[0] = Sig.c[0][0]; o[1] = Sig.c[0][1]; o[2] = Sig.c[0][2]; o[3] = Sig.c[1][1]; o[4] = Sig.c[1][2];
			o[5] = Sig.c[2][2]; o[6] = B[0]; o[7] = B[1]; o[8] = B[2]; o[9] = (float)C;
		}
	}
}

/* ------------------------------------------------------------------ binning ----------- */
static uint32_t higher_msb(uint32_t n)
{
	uint32_t msb = sizeof(n) * 4, step = msb;
	while (step > 1) {
		step /= 2;
		if (n >> msb) msb += step; else msb -= step;
	}
	if (n >> msb) msb++;
	return msb;
}

/* point_offsets = inclusive scan of tiles_touched; returns R (rasterizer_impl.cu:332-336). */
int64_t oracle_scan(int P, const uint32_t* tiles_touched, uint32_t* point_offsets)
{
	uint32_t acc = 0;
	for (int i = 0; i < P; i++) { acc += tiles_touched[i]; point_offsets[i] = acc; }
	return P > 0 ? (int64_t)acc : 0;
}

typedef struct { uint64_t key; uint32_t val; } kv_t;

/* Stable LSD radix sort on bits [0, nbits) -- what cub::DeviceRadixSort::SortPairs guarantees. */
static void radix_sort_pairs(kv_t* a, kv_t* tmp, size_t n, int nbits)
{
	for (int shift = 0; shift < nbits; shift += 8) {
		const int bits = imin(8, nbits - shift);
		const uint64_t mask = ((uint64_t)1 << bits) - 1;
		size_t count[257];
		memset(count, 0, sizeof(count));
		for (size_t i = 0; i < n; i++) count[((a[i].key >> shift) & mask) + 1]++;
		for (int b = 0; b < 256; b++) count[b + 1] += count[b];
		for (size_t i = 0; i < n; i++) tmp[count[(a[i].key >> shift) & mask]++] = a[i];
		memcpy(a, tmp, n * sizeof(kv_t));
	}
}

/* duplicateWithKeys + SortPairs + identifyTileRanges.  keys/point_list have R entries, ranges
 * has 2*T uint32 (x=start, y=end; (0,0) for untouched tiles).  keys_unsorted/vals_unsorted may be
 * NULL.  Returns 0. */
int oracle_binning(int P, int W, int H, const float* means2D, const float* depths, const int32_t* radii,
                   const uint32_t* point_offsets, int64_t R, uint64_t* keys_unsorted, uint32_t* vals_unsorted,
                   uint64_t* keys, uint32_t* point_list, uint32_t* ranges)
{
	const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
	const int T = gx * gy;
	memset(ranges, 0, (size_t)T * 2 * sizeof(uint32_t));
	if (R <= 0) return 0;
	kv_t* a = (kv_t*)malloc((size_t)R * sizeof(kv_t));
	kv_t* tmp = (kv_t*)malloc((size_t)R * sizeof(kv_t));
	if (!a || !tmp) { free(a); free(tmp); return -1; }
	for (int idx = 0; idx < P; idx++) {
		if (radii[idx] <= 0) continue;
		size_t off = idx == 0 ? 0 : point_offsets[idx - 1];
		int x0, y0, x1, y1;
		get_rect(means2D[2 * idx], means2D[2 * idx + 1], radii[idx], gx, gy, &x0, &y0, &x1, &y1);
		uint32_t dbits;
		memcpy(&dbits, &depths[idx], 4);
		for (int y = y0; y < y1; y++)
			for (int x = x0; x < x1; x++) {
				a[off].key = ((uint64_t)(uint32_t)(y * gx + x) << 32) | dbits;
				a[off].val = (uint32_t)idx;
				off++;
			}
	}
	if (keys_unsorted) for (int64_t i = 0; i < R; i++) keys_unsorted[i] = a[i].key;
	if (vals_unsorted) for (int64_t i = 0; i < R; i++) vals_unsorted[i] = a[i].val;
	radix_sort_pairs(a, tmp, (size_t)R, 32 + (int)higher_msb((uint32_t)T));
	for (int64_t i = 0; i < R; i++) { keys[i] = a[i].key; point_list[i] = a[i].val; }
	for (int64_t i = 0; i < R; i++) {
		const uint32_t cur = (uint32_t)(keys[i] >> 32);
		if (i == 0) ranges[2 * cur] = 0;
		else {
			const uint32_t prev = (uint32_t)(keys[i - 1] >> 32);
			if (cur != prev) { ranges[2 * prev + 1] = (uint32_t)i; ranges[2 * cur] = (uint32_t)i; }
		}
		if (i == R - 1) ranges[2 * cur + 1] = (uint32_t)R;
	}
	free(a);
	free(tmp);
	return 0;
}

/* ------------------------------------------------------------------ forward blend ----- */
typedef struct { float n[3]; double AA, BB; float t; float alpha; float G; } pair_t;

/* Shared pair evaluation (forward.cu:502-535 == backward.cu:776-804).  Returns 0 if skipped. */
static int eval_pair(const float* q, float w, float rx, float ry, pair_t* p)
{
	/* These five float32 values feed a difference of two ~6e5 terms: one ulp moves alpha by
	 * percents.  The reference's sm_100a build evaluates each 3-term sum as
	 * third + fma(first_a, first_b, round(second_a * second_b)) (read off its SASS), which is
	 * restated here with fmaf so that this stage agrees with the GPU to expf accuracy. */
	p->n[0] = q[2] + fmaf(q[0], rx, q[1] * ry);
	p->n[1] = q[4] + fmaf(q[1], rx, q[3] * ry);
	p->n[2] = q[5] + fmaf(q[4], ry, q[2] * rx);
	p->AA = (double)(fmaf(p->n[0], rx, p->n[1] * ry) + p->n[2]);
	const float bb = q[8] + fmaf(q[6], rx, q[7] * ry);
	p->BB = (double)(bb + bb);
	const float CC = q[9];
	p->t = (float)(-p->BB / (2 * p->AA));
	if ((double)p->t <= NEAR_PLANE) return 0;
	const double min_value = -(p->BB / p->AA) * (p->BB / 4.) + (double)CC;
	float power = (float)(-0.5 * min_value);
	if (power > 0.0f) power = 0.0f;
	p->G = expf(power);
	p->alpha = fminf(0.99f, w * p->G);
	if (p->alpha < 1.0f / 255.0f) return 0;
	return 1;
}

static inline float pixel_ray(int p, int S, float focal)
{
	const float pf = (float)p + 0.5f;
	return (float)(((double)pf - S / 2.) / (double)focal);
}

void oracle_render_forward(int W, int H, float tan_fovx, float tan_fovy, const uint32_t* ranges,
                           const uint32_t* point_list, const float* v2g, const float* conic_opacity,
                           const float* features, const float* bg, float* out_color, float* final_T,
                           uint32_t* n_contrib)
{
	const float focal_y = H / (2.0f * tan_fovy);
	const float focal_x = W / (2.0f * tan_fovx);
	const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
	const size_t N = (size_t)W * H;
#pragma omp parallel for schedule(dynamic, 1)
	for (int tile = 0; tile < gx * gy; tile++) {
		const int tx = tile % gx, ty = tile / gx;
		const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
		for (int ly = 0; ly < BLOCK_Y; ly++)
			for (int lx = 0; lx < BLOCK_X; lx++) {
				const int px = tx * BLOCK_X + lx, py = ty * BLOCK_Y + ly;
				if (px >= W || py >= H) continue;
				const size_t pix = (size_t)W * py + px;
				const float rx = pixel_ray(px, W, focal_x), ry = pixel_ray(py, H, focal_y);
				float T = 1.0f, C[8] = { 0 }, dist1 = 0, dist2 = 0, distortion = 0;
				uint32_t contributor = 0, last = 0, maxc = 0xFFFFFFFFu;
				for (uint32_t k = r0; k < r1; k++) {
					contributor++;
					const uint32_t id = point_list[k];
					pair_t p;
					if (!eval_pair(v2g + 10 * (size_t)id, conic_opacity[4 * (size_t)id + 3], rx, ry, &p)) continue;
					const float test_T = T * (1 - p.alpha);
					if (test_T < 0.0001f) break;   /* done = true */
					const double td = p.t;
					const float m = (float)((FAR_PLANE * td - FAR_PLANE * NEAR_PLANE) / ((FAR_PLANE - NEAR_PLANE) * td));
					const float len = (float)sqrt((double)(p.n[0] * p.n[0] + p.n[1] * p.n[1] + p.n[2] * p.n[2]) + 1e-7);
					const float nn[3] = { -p.n[0] / len, -p.n[1] / len, -p.n[2] / len };
					const float A = 1 - T;
					const float err = m * m * A + dist2 - 2 * m * dist1;
					distortion += err * p.alpha * T;
					dist1 += m * p.alpha * T;
					dist2 += m * m * p.alpha * T;
					for (int ch = 0; ch < 3; ch++) C[ch] += features[3 * (size_t)id + ch] * p.alpha * T;
					for (int ch = 0; ch < 3; ch++) C[3 + ch] += nn[ch] * p.alpha * T;
					if (T > 0.5f) { C[6] = p.t; maxc = contributor; }
					C[7] += p.alpha * T;
					T = test_T;
					last = contributor;
				}
				final_T[pix] = T;
				final_T[pix + N] = dist1;
				final_T[pix + 2 * N] = dist2;
				final_T[pix + 3 * N] = distortion;
				n_contrib[pix] = last;
				n_contrib[pix + N] = maxc;
				for (int ch = 0; ch < 3; ch++) out_color[ch * N + pix] = C[ch] + T * bg[ch];
				for (int ch = 3; ch < 8; ch++) out_color[ch * N + pix] = C[ch];
				out_color[8 * N + pix] = (float)((double)distortion / ((double)((1 - T) * (1 - T)) + 1e-7));
			}
	}
}

/* ------------------------------------------------------------------ point integration -- */
/* Rasterizer::integrate (rasterizer_impl.cu:530-792): preprocessPointsCUDA (forward.cu:722-766), createWithKeys
 * (rasterizer_impl.cu:113-144) + stable sort by (tile, depth), integrateCUDA (forward.cu:803-1218).
 * The per-ray float32 quadric uses the roundings of the reference's sm_100a build (tools/sass_symexec.py; the
 * fused/unfused pattern differs per ray) restated with fmaf, as eval_pair does for the blend.
 * Control flow follows the kernel, including its 256-points-per-batch re-scan (forward.cu:1029-1206): a tile
 * iterates while any of its pixels overflowed the batch, and pixels that did not overflow re-examine the tile's
 * LAST point in every further round (point_counter_last = point_counter - 1), which only affects the
 * "number of projected points" channel 8.
 * Inputs: the Gaussian state of oracle_preprocess + oracle_binning.  Outputs: out_color[9,H,W] (zero-initialised
 * by the caller, as the glue does), final_T[N], n_contrib[N], out_alpha[PN] (initialised to 1), out_rgb[PN,3]. */
#define MAX_CONTRIBUTED 1024   /* MAX_NUM_CONTRIBUTORS * 4 */
#define MAX_PROJECTED 256

typedef struct { float AA, BB, CC; } rayquad;
static rayquad ray_quadric(const float* v, float rx, float ry, int variant)
{
	float n0, n1, n2, bb;
	if (variant == 0) {
		n0 = v[2] + fmaf(rx, v[0], ry * v[1]);
		n1 = v[4] + fmaf(rx, v[1], ry * v[3]);
		n2 = v[5] + fmaf(ry, v[4], rx * v[2]);
		bb = v[8] + fmaf(rx, v[6], ry * v[7]);
	} else {
		n0 = v[2] + (rx * v[0] + ry * v[1]);
		n1 = (variant == 1) ? v[4] + fmaf(rx, v[1], ry * v[3]) : v[4] + (rx * v[1] + ry * v[3]);
		n2 = v[5] + fmaf(rx, v[2], ry * v[4]);
		bb = v[8] + (rx * v[6] + ry * v[7]);
	}
	rayquad q;
	q.AA = n2 + fmaf(rx, n0, ry * n1);
	q.BB = bb + bb;
	q.CC = v[9];
	return q;
}

int oracle_integrate(int P, int PN, int W, int H, float tan_fovx, float tan_fovy, const float* viewmatrix,
                     const float* points3D, const uint32_t* ranges, const uint32_t* point_list, const float* v2g,
                     const float* conic_opacity, const float* features, const float* bg, float* out_color,
                     float* final_T, uint32_t* n_contrib, float* out_alpha, float* out_rgb)
{
	(void)P;
	const float focal_y = H / (2.0f * tan_fovy);
	const float focal_x = W / (2.0f * tan_fovx);
	const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
	const int T = gx * gy;
	const size_t N = (size_t)W * H;
	/* ---- points: project, bin by tile, stable sort by (tile, depth) ---- */
	float* p2d = (float*)malloc(sizeof(float) * 2 * (size_t)(PN > 0 ? PN : 1));
	float* pdepth = (float*)malloc(sizeof(float) * (size_t)(PN > 0 ? PN : 1));
	kv_t* kv = (kv_t*)malloc(sizeof(kv_t) * (size_t)(PN > 0 ? PN : 1));
	kv_t* tmp = (kv_t*)malloc(sizeof(kv_t) * (size_t)(PN > 0 ? PN : 1));
	uint32_t* pranges = (uint32_t*)calloc((size_t)2 * T, sizeof(uint32_t));
	size_t np = 0;
	const float* vm = viewmatrix;
	for (int i = 0; i < PN; i++) {
		const float x = points3D[3 * i], y = points3D[3 * i + 1], z = points3D[3 * i + 2];
		const float vx = vm[0] * x + vm[4] * y + vm[8] * z + vm[12];
		const float vy = vm[1] * x + vm[5] * y + vm[9] * z + vm[13];
		const float vz = vm[2] * x + vm[6] * y + vm[10] * z + vm[14];
		if (vz <= 0.2f) continue;
		const float px = (float)(focal_x * vx / (vz + 0.0000001f) + W / 2.);
		const float py = (float)(focal_y * vy / (vz + 0.0000001f) + H / 2.);
		if (px < 0 || px >= W || py < 0 || py >= H) continue;
		p2d[2 * i] = px; p2d[2 * i + 1] = py; pdepth[i] = vz;
		const int tx = imin(gx - 1, imax(0, (int)(px / BLOCK_X))), ty = imin(gy - 1, imax(0, (int)(py / BLOCK_Y)));
		uint32_t dbits;
		memcpy(&dbits, &vz, 4);
		kv[np].key = ((uint64_t)(uint32_t)(ty * gx + tx) << 32) | dbits;
		kv[np].val = (uint32_t)i;
		np++;
	}
	radix_sort_pairs(kv, tmp, np, 64);
	for (size_t i = 0; i < np; i++) {
		const uint32_t t = (uint32_t)(kv[i].key >> 32);
		if (i == 0 || (uint32_t)(kv[i - 1].key >> 32) != t) pranges[2 * t] = (uint32_t)i;
		pranges[2 * t + 1] = (uint32_t)i + 1;
	}
#pragma omp parallel for schedule(dynamic, 1)
	for (int tile = 0; tile < T; tile++) {
		const int tx = tile % gx, ty = tile / gx;
		const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
		const uint32_t p0 = pranges[2 * tile], p1 = pranges[2 * tile + 1];
		/* per-pixel state of the tile */
		uint16_t* contributed = (uint16_t*)malloc(sizeof(uint16_t) * MAX_CONTRIBUTED * BLOCK_X * BLOCK_Y);
		uint32_t ncontrib[BLOCK_X * BLOCK_Y], lastc[BLOCK_X * BLOCK_Y], counter_last[BLOCK_X * BLOCK_Y];
		int total_proj[BLOCK_X * BLOCK_Y], pdone[BLOCK_X * BLOCK_Y];
		float pix_rgb[BLOCK_X * BLOCK_Y][3];
		/* ---- phase 1: five rays per pixel ---- */
		for (int l = 0; l < BLOCK_X * BLOCK_Y; l++) {
			const int lx = l % BLOCK_X, ly = l / BLOCK_X;
			const int px = tx * BLOCK_X + lx, py = ty * BLOCK_Y + ly;
			ncontrib[l] = 0; lastc[l] = 0; counter_last[l] = 0; total_proj[l] = 0;
			pdone[l] = !(px < W && py < H);
			if (pdone[l]) continue;
			const size_t pix = (size_t)W * py + px;
			const float pfx = (float)px + 0.5f, pfy = (float)py + 0.5f;
			static const float offx[5] = { 0.0f, -0.5f, 0.5f, -0.5f, 0.5f }, offy[5] = { 0.0f, -0.5f, -0.5f, 0.5f, 0.5f };
			static const int variant[5] = { 0, 1, 2, 1, 2 };
			float rxs[5], rys[5], Ts[5] = { 1, 1, 1, 1, 1 };
			for (int k = 0; k < 5; k++) {
				rxs[k] = (float)(((double)(pfx + offx[k]) - W / 2.) / (double)focal_x);
				rys[k] = (float)(((double)(pfy + offy[k]) - H / 2.) / (double)focal_y);
			}
			float C0 = 0, C1 = 0, C2 = 0, Cdepth = 0, Calpha = 0;
			uint32_t contributor = 0;
			int done = 0;
			for (uint32_t kk = r0; kk < r1 && !done; kk++) {
				contributor++;
				const uint32_t id = point_list[kk];
				const float* v = v2g + 10 * (size_t)id;
				const float w = conic_opacity[4 * (size_t)id + 3];
				int used = 0;
				for (int k = 0; k < 5; k++) {
					const rayquad q = ray_quadric(v, rxs[k], rys[k], variant[k]);
					const float t = -q.BB / (2 * q.AA);
					if ((double)t <= NEAR_PLANE) continue;
					const double min_value = fma((double)(-q.BB / q.AA), (double)q.BB * 0.25, (double)q.CC);
					float power = (float)(min_value * -0.5);
					if (power > 0.0f) power = 0.0f;
					const float alpha = fminf(0.99f, w * expf(power));
					if (alpha < 1.0f / 255.0f) continue;
					const float test_T = Ts[k] * (1 - alpha);
					if (test_T < 0.0001f) continue;
					if (k == 0) {
						C0 = fmaf(Ts[0], alpha * features[3 * (size_t)id + 0], C0);
						C1 = fmaf(Ts[0], alpha * features[3 * (size_t)id + 1], C1);
						C2 = fmaf(Ts[0], alpha * features[3 * (size_t)id + 2], C2);
					}
					if (t > Cdepth) Cdepth = t;
					if (k == 0) Calpha = fmaf(Ts[0], alpha, Calpha);
					Ts[k] = test_T;
					used = 1;
				}
				if (used) {
					lastc[l] = contributor;
					contributed[(size_t)l * MAX_CONTRIBUTED + ncontrib[l]] = (uint16_t)contributor;
					ncontrib[l]++;
					if (ncontrib[l] >= MAX_CONTRIBUTED) done = 1;
				}
			}
			final_T[pix] = Ts[0];
			n_contrib[pix] = lastc[l];
			pix_rgb[l][0] = C0 + Ts[0] * bg[0]; pix_rgb[l][1] = C1 + Ts[0] * bg[1]; pix_rgb[l][2] = C2 + Ts[0] * bg[2];
			for (int ch = 0; ch < 3; ch++) out_color[ch * N + pix] = pix_rgb[l][ch];
			out_color[6 * N + pix] = Cdepth;
			out_color[7 * N + pix] = Calpha;
		}
		/* ---- phase 2: rounds of at most 256 projected points per pixel, block-wide loop ---- */
		for (;;) {
			int all_done = 1;
			for (int l = 0; l < BLOCK_X * BLOCK_Y; l++) all_done &= pdone[l];
			if (all_done) break;
			for (int l = 0; l < BLOCK_X * BLOCK_Y; l++) {
				const int lx = l % BLOCK_X, ly = l / BLOCK_X;
				const int px = tx * BLOCK_X + lx, py = ty * BLOCK_Y + ly;
				if (!(px < W && py < H)) continue;
				const float pfx = (float)px + 0.5f, pfy = (float)py + 0.5f;
				int ids[MAX_PROJECTED], nproj = 0, exceeded = 0;
				float xs[MAX_PROJECTED], ys[MAX_PROJECTED], ds[MAX_PROJECTED];
				uint32_t counter = 0;
				for (uint32_t kk = p0; kk < p1; kk++) {
					counter++;
					if (counter <= counter_last[l]) continue;
					const uint32_t pid = kv[kk].val;
					const float x = p2d[2 * pid], y = p2d[2 * pid + 1];
					if (((double)x >= (double)pfx - 0.5) && ((double)x < (double)pfx + 0.5) && ((double)y >= (double)pfy - 0.5) && ((double)y < (double)pfy + 0.5)) {
						if (nproj >= MAX_PROJECTED) { exceeded = 1; break; }
						ids[nproj] = (int)pid; xs[nproj] = x; ys[nproj] = y; ds[nproj] = pdepth[pid]; nproj++;
					}
				}
				counter_last[l] = counter - 1;        /* uint32 wrap when the tile has no points, as in the kernel */
				pdone[l] = !exceeded;
				total_proj[l] += nproj;
				float acc[MAX_PROJECTED], Tp[MAX_PROJECTED];
				for (int k = 0; k < nproj; k++) { acc[k] = 0; Tp[k] = 1; }
				uint32_t iterated = 0;
				uint32_t second = 0;
				for (uint32_t kk = r0; kk < r1; kk++) {
					iterated++;
					if (iterated > lastc[l]) break;
					if (second >= ncontrib[l] || iterated != (uint32_t)contributed[(size_t)l * MAX_CONTRIBUTED + second]) continue;
					second++;
					const uint32_t id = point_list[kk];
					const float* v = v2g + 10 * (size_t)id;
					const float w = conic_opacity[4 * (size_t)id + 3];
					for (int k = 0; k < nproj; k++) {
						const float rx = (float)(((double)xs[k] - W / 2.) / (double)focal_x);
						const float ry = (float)(((double)ys[k] - H / 2.) / (double)focal_y);
						const rayquad q = ray_quadric(v, rx, ry, 0);
						float t = -q.BB / (2 * q.AA);
						if (t > ds[k]) t = ds[k];
						const float power = (q.CC + fmaf(q.BB, t, t * (q.AA * t))) * -0.5f;
						const float alpha = fminf(0.99f, w * expf(power));
						if (alpha < 1.0f / 255.0f) continue;
						acc[k] = fmaf(alpha, Tp[k], acc[k]);
						Tp[k] = Tp[k] * (1 - alpha);
					}
				}
				for (int k = 0; k < nproj; k++) {
					out_alpha[ids[k]] = acc[k];
					for (int ch = 0; ch < 3; ch++) out_rgb[3 * (size_t)ids[k] + ch] = pix_rgb[l][ch];
				}
			}
		}
		for (int l = 0; l < BLOCK_X * BLOCK_Y; l++) {
			const int lx = l % BLOCK_X, ly = l / BLOCK_X;
			const int px = tx * BLOCK_X + lx, py = ty * BLOCK_Y + ly;
			if (px < W && py < H) out_color[8 * N + (size_t)W * py + px] = (float)total_proj[l];
		}
		free(contributed);
	}
	free(p2d); free(pdepth); free(kv); free(tmp); free(pranges);
	return (int)np;
}

/* ------------------------------------------------------------------ backward blend ---- */
/* Accumulates into zero-initialised dL_dmean2D[P,3], dL_dopacity[P], dL_dcolors[P,3],
 * dL_dv2g[P,10] (backward.cu:634-955).  Serial over tiles per thread with private
 * accumulation order = tile order, so the result is deterministic. */
void oracle_render_backward(int P, int W, int H, float tan_fovx, float tan_fovy, const uint32_t* ranges,
                            const uint32_t* point_list, const float* v2g, const float* conic_opacity,
                            const float* means2D, const float* features, const float* bg, const float* final_Ts,
                            const uint32_t* n_contrib, const float* dL_dpixels, float* dL_dmean2D,
                            float* dL_dopacity, float* dL_dcolors, float* dL_dv2g)
{
	(void)P;
	const float focal_y = H / (2.0f * tan_fovy);
	const float focal_x = W / (2.0f * tan_fovx);
	const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
	const size_t N = (size_t)W * H;
	const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
	for (int tile = 0; tile < gx * gy; tile++) {
		const int tx = tile % gx, ty = tile / gx;
		const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
		for (int ly = 0; ly < BLOCK_Y; ly++)
			for (int lx = 0; lx < BLOCK_X; lx++) {
				const int px = tx * BLOCK_X + lx, py = ty * BLOCK_Y + ly;
				if (px >= W || py >= H) continue;
				const size_t pix = (size_t)W * py + px;
				const float rx = pixel_ray(px, W, focal_x), ry = pixel_ray(py, H, focal_y);
				const float T_final = final_Ts[pix];
				float T = T_final;
				const float final_D = final_Ts[pix + N];
				const float final_A = 1 - T_final;
				const float dL_dreg = dL_dpixels[8 * N + pix];
				const uint32_t last_contributor = n_contrib[pix];
				const int32_t max_contributor = (int32_t)n_contrib[pix + N];
				float accum_rec[3] = { 0 }, accum_normal_rec[3] = { 0 }, last_color[3] = { 0 }, last_normal[3] = { 0 };
				float dL_dpixel[3], dL_dnormal2D[3];
				for (int i = 0; i < 3; i++) { dL_dpixel[i] = dL_dpixels[i * N + pix]; dL_dnormal2D[i] = dL_dpixels[(3 + i) * N + pix]; }
				const float dL_dmax_depth = dL_dpixels[6 * N + pix];
				float last_alpha = 0;
				float bg_dot = 0;
				for (int i = 0; i < 3; i++) bg_dot += bg[i] * dL_dpixel[i];
				uint32_t contributor = r1 - r0;
				for (uint32_t kk = r1; kk > r0; kk--) {
					contributor--;
					if (contributor >= last_contributor) continue;
					const uint32_t id = point_list[kk - 1];
					const float* q = v2g + 10 * (size_t)id;
					const float* con = conic_opacity + 4 * (size_t)id;
					pair_t p;
					if (!eval_pair(q, con[3], rx, ry, &p)) continue;
					const float dx = (float)((double)means2D[2 * (size_t)id] - ((double)((float)px + 0.5f) - 0.5));
					const float dy = (float)((double)means2D[2 * (size_t)id + 1] - ((double)((float)py + 0.5f) - 0.5));
					const double td = p.t;
					const float m = (float)((FAR_PLANE * td - FAR_PLANE * NEAR_PLANE) / ((FAR_PLANE - NEAR_PLANE) * td));
					const float dm_dt = (float)((FAR_PLANE * NEAR_PLANE) / ((FAR_PLANE - NEAR_PLANE) * td * td));
					const float len = (float)sqrt((double)(p.n[0] * p.n[0] + p.n[1] * p.n[1] + p.n[2] * p.n[2]) + 1e-7);
					const float nn[3] = { -p.n[0] / len, -p.n[1] / len, -p.n[2] / len };
					T = T / (1.f - p.alpha);
					const float weight = p.alpha * T;
					float dL_dalpha = 0.0f;
					for (int ch = 0; ch < 3; ch++) {
						const float c = features[3 * (size_t)id + ch];
						accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
						last_color[ch] = c;
						dL_dalpha += (c - accum_rec[ch]) * dL_dpixel[ch];
						dL_dcolors[3 * (size_t)id + ch] += weight * dL_dpixel[ch];
					}
					/* distortion: dL_dweight is computed then detached in the reference (backward.cu:846-858) */
					const float dL_dmax_t = 2.0f * (T * p.alpha) * (m * final_A - final_D) * dL_dreg * dm_dt;
					float dL_dnn[3];
					for (int ch = 0; ch < 3; ch++) {
						accum_normal_rec[ch] = last_alpha * last_normal[ch] + (1.f - last_alpha) * accum_normal_rec[ch];
						last_normal[ch] = nn[ch];
						dL_dalpha += (nn[ch] - accum_normal_rec[ch]) * dL_dnormal2D[ch];
						dL_dnn[ch] = p.alpha * T * dL_dnormal2D[ch];
					}
					float dL_dlength = dL_dnn[0] * p.n[0] + dL_dnn[1] * p.n[1] + dL_dnn[2] * p.n[2];
					dL_dlength *= 1.f / (len * len);
					float dL_dn[3];
					for (int k = 0; k < 3; k++) dL_dn[k] = (-dL_dnn[k] + dL_dlength * p.n[k]) / len;
					float dL_dt = dL_dmax_t;
					if ((int64_t)contributor == (int64_t)max_contributor - 1 && max_contributor != -1) dL_dt += dL_dmax_depth;
					dL_dalpha *= T;
					last_alpha = p.alpha;
					dL_dalpha += (-T_final / (1.f - p.alpha)) * bg_dot;

					const float dL_dG = con[3] * dL_dalpha;
					const float gdx = p.G * dx, gdy = p.G * dy;
					const float dG_ddelx = -gdx * con[0] - gdy * con[1];
					const float dG_ddely = -gdy * con[2] - gdx * con[1];
					const float gmx = dL_dG * dG_ddelx * ddelx_dx, gmy = dL_dG * dG_ddely * ddely_dy;
					dL_dmean2D[3 * (size_t)id + 0] += gmx;
					dL_dmean2D[3 * (size_t)id + 1] += gmy;
					dL_dmean2D[3 * (size_t)id + 2] += fabsf(gmx) + fabsf(gmy);
					dL_dopacity[id] += p.G * dL_dalpha;

					const float dL_dmin_value = dL_dG * p.G * -0.5f;
					double dL_dA = dL_dmin_value * (p.BB / p.AA) * (p.BB / p.AA) / 4.f;
					double dL_dB = dL_dmin_value * -p.BB / (2 * p.AA);
					const double dL_dC = dL_dmin_value * 1.0f;
					dL_dA += dL_dt * p.BB / (2 * p.AA * p.AA);
					dL_dB += dL_dt * -1.f / (2 * p.AA);
					dL_dn[0] += dL_dA * rx;
					dL_dn[1] += dL_dA * ry;
					dL_dn[2] += dL_dA;
					float* o = dL_dv2g + 10 * (size_t)id;
					o[0] += dL_dn[0] * rx;
					o[1] += dL_dn[0] * ry + dL_dn[1] * rx;
					o[2] += dL_dn[0] + dL_dn[2] * rx;
					o[3] += dL_dn[1] * ry;
					o[4] += dL_dn[1] + dL_dn[2] * ry;
					o[5] += dL_dn[2];
					o[6] += dL_dB * 2 * rx;
					o[7] += dL_dB * 2 * ry;
					o[8] += dL_dB * 2;
					o[9] += dL_dC;
				}
			}
	}
}

/* ------------------------------------------------------------------ backward preprocess */
/* `real` is float in libgof_oracle.so (the reference's arithmetic) and double in
 * libgof_oracle_f64.so (oracle/Makefile: -DORACLE_REAL=double): the same formulas evaluated
 * without float32 rounding.  dL/dscale, dL/drot, dL/dmean are catastrophically ill-conditioned
 * functions of dL/dview2gaussian at F3D-Gaus scales, so tests measure both the reference and the
 * CUDA kernel against this float64 evaluation instead of against each other. */
#ifndef ORACLE_REAL
#define ORACLE_REAL float
#endif
typedef ORACLE_REAL real;
typedef struct { real c[3][3]; } m3r;

int oracle_real_bytes(void) { return (int)sizeof(real); }

static m3r m3r_mul(m3r a, m3r b)
{
	m3r r;
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++)
			r.c[i][j] = a.c[0][j] * b.c[i][0] + a.c[1][j] * b.c[i][1] + a.c[2][j] * b.c[i][2];
	return r;
}
static m3r m3r_t(m3r a)
{
	m3r r;
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) r.c[i][j] = a.c[j][i];
	return r;
}
static void g2v_r(const float* q, const float* mean, const float* vm, m3r* Rt, real* t)
{
	const real r = q[0], x = q[1], y = q[2], z = q[3];
	m3r R;
	R.c[0][0] = 1 - 2 * (y * y + z * z); R.c[0][1] = 2 * (x * y - r * z); R.c[0][2] = 2 * (x * z + r * y);
	R.c[1][0] = 2 * (x * y + r * z); R.c[1][1] = 1 - 2 * (x * x + z * z); R.c[1][2] = 2 * (y * z - r * x);
	R.c[2][0] = 2 * (x * z - r * y); R.c[2][1] = 2 * (y * z + r * x); R.c[2][2] = 1 - 2 * (x * x + y * y);
	real G2V[4][3];
	for (int c = 0; c < 3; c++)
		for (int j = 0; j < 3; j++)
			G2V[c][j] = (real)vm[0 + j] * R.c[0][c] + (real)vm[4 + j] * R.c[1][c] + (real)vm[8 + j] * R.c[2][c];
	for (int j = 0; j < 3; j++)
		G2V[3][j] = (real)vm[0 + j] * mean[0] + (real)vm[4 + j] * mean[1] + (real)vm[8 + j] * mean[2] + vm[12 + j];
	for (int c = 0; c < 3; c++)
		for (int rr = 0; rr < 3; rr++) Rt->c[c][rr] = G2V[rr][c];
	t[0] = G2V[3][0]; t[1] = G2V[3][1]; t[2] = G2V[3][2];
}

static void dnormvdv(const real* v, const real* dv, real* out)
{
	const real sum2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
	const real inv = (real)1 / (real)sqrt((double)(sum2 * sum2 * sum2));
	out[0] = ((+sum2 - v[0] * v[0]) * dv[0] - v[1] * v[0] * dv[1] - v[2] * v[0] * dv[2]) * inv;
	out[1] = (-v[0] * v[1] * dv[0] + (sum2 - v[1] * v[1]) * dv[1] - v[2] * v[1] * dv[2]) * inv;
	out[2] = (-v[0] * v[2] * dv[0] - v[1] * v[2] * dv[1] + (sum2 - v[2] * v[2]) * dv[2]) * inv;
}
/* All outputs zero-initialised by the caller ([P,*]); dL_dcolor is the blend's colour gradient.
 * dL_dmeans/dL_dscale/dL_drot are ASSIGNED for visible Gaussians (backward.cu:494-497,570-573,585). */
void oracle_preprocess_backward(int P, int D, int M, const float* means3D, const int32_t* radii, const float* shs,
                                const uint8_t* clamped, const float* scales, const float* rotations, const float* vm,
                                const float* campos, const float* dL_dv2g, const float* dL_dcolor, real* dL_dmeans,
                                real* dL_dsh, real* dL_dscale, real* dL_drot)
{
#pragma omp parallel for schedule(static)
	for (int idx = 0; idx < P; idx++) {
		if (!(radii[idx] > 0)) continue;
		const float* dq = dL_dv2g + 10 * (size_t)idx;
		const float* q = rotations + 4 * idx;
		const real r = q[0], x = q[1], y = q[2], z = q[3];
		m3r Rt;
		real t[3];
		g2v_r(q, means3D + 3 * idx, vm, &Rt, t);
		real t2[3];
		for (int k = 0; k < 3; k++) t2[k] = -(Rt.c[0][k] * t[0] + Rt.c[1][k] * t[1] + Rt.c[2][k] * t[2]);
		double Sinv[3];
		for (int k = 0; k < 3; k++) Sinv[k] = 1.0 / ((double)scales[3 * idx + k] * scales[3 * idx + k] + 1e-7);
		m3r SR;
		for (int c = 0; c < 3; c++)
			for (int k = 0; k < 3; k++) SR.c[c][k] = (real)(Sinv[k] * (double)Rt.c[c][k]);
		m3r dSig;
		dSig.c[0][0] = dq[0]; dSig.c[0][1] = (real)0.5 * dq[1]; dSig.c[0][2] = (real)0.5 * dq[2];
		dSig.c[1][0] = (real)0.5 * dq[1]; dSig.c[1][1] = dq[3]; dSig.c[1][2] = (real)0.5 * dq[4];
		dSig.c[2][0] = (real)0.5 * dq[2]; dSig.c[2][1] = (real)0.5 * dq[4]; dSig.c[2][2] = dq[5];
		const real dB[3] = { dq[6], dq[7], dq[8] };
		const real dC = dq[9];
		m3r dSR = m3r_mul(Rt, dSig);
		for (int c = 0; c < 3; c++)
			for (int k = 0; k < 3; k++) dSR.c[c][k] += t2[k] * dB[c];
		m3r dRt = m3r_t(m3r_mul(dSig, m3r_t(SR)));
		for (int c = 0; c < 3; c++)
			for (int k = 0; k < 3; k++) dRt.c[c][k] += (real)(Sinv[k] * (double)dSR.c[c][k]);
		real dSinv[3], dt2[3];
		for (int k = 0; k < 3; k++) {
			dSinv[k] = dSR.c[0][k] * Rt.c[0][k] + dSR.c[1][k] * Rt.c[1][k] + dSR.c[2][k] * Rt.c[2][k];
			dt2[k] = (real)(2 * t2[k] * Sinv[k] * dC + dB[0] * SR.c[0][k] + dB[1] * SR.c[1][k] + dB[2] * SR.c[2][k]);
			dSinv[k] += dC * t2[k] * t2[k];
			dL_dscale[3 * idx + k] = (real)(-2 / scales[3 * idx + k] * Sinv[k] * dSinv[k]);
		}
		m3r dG2V_R = m3r_t(dRt);
		for (int c = 0; c < 3; c++)
			for (int k = 0; k < 3; k++) dG2V_R.c[c][k] += -dt2[c] * t[k];
		real dG2V_t[3];
		for (int c = 0; c < 3; c++) dG2V_t[c] = Rt.c[c][0] * -dt2[0] + Rt.c[c][1] * -dt2[1] + Rt.c[c][2] * -dt2[2];
		real dG2W[4][3];
		for (int c = 0; c < 3; c++)
			for (int k = 0; k < 3; k++)
				dG2W[c][k] = vm[4 * k + 0] * dG2V_R.c[c][0] + vm[4 * k + 1] * dG2V_R.c[c][1] + vm[4 * k + 2] * dG2V_R.c[c][2];
		for (int k = 0; k < 3; k++)
			dG2W[3][k] = vm[4 * k + 0] * dG2V_t[0] + vm[4 * k + 1] * dG2V_t[1] + vm[4 * k + 2] * dG2V_t[2];
		real dmean[3] = { dG2W[3][0], dG2W[3][1], dG2W[3][2] };
#define MT(a, b) dG2W[a][b]
		dL_drot[4 * idx + 0] = 2 * z * (MT(0, 1) - MT(1, 0)) + 2 * y * (MT(2, 0) - MT(0, 2)) + 2 * x * (MT(1, 2) - MT(2, 1));
		dL_drot[4 * idx + 1] = 2 * y * (MT(1, 0) + MT(0, 1)) + 2 * z * (MT(2, 0) + MT(0, 2)) + 2 * r * (MT(1, 2) - MT(2, 1)) - 4 * x * (MT(2, 2) + MT(1, 1));
		dL_drot[4 * idx + 2] = 2 * x * (MT(1, 0) + MT(0, 1)) + 2 * r * (MT(2, 0) - MT(0, 2)) + 2 * z * (MT(1, 2) + MT(2, 1)) - 4 * y * (MT(2, 2) + MT(0, 0));
		dL_drot[4 * idx + 3] = 2 * r * (MT(0, 1) - MT(1, 0)) + 2 * x * (MT(2, 0) + MT(0, 2)) + 2 * y * (MT(1, 2) + MT(2, 1)) - 4 * z * (MT(1, 1) + MT(0, 0));
#undef MT

		if (shs) {
			const float* sh = shs + (size_t)idx * M * 3;
			real* dsh = dL_dsh + (size_t)idx * M * 3;
			const real dir_orig[3] = { (real)means3D[3 * idx] - campos[0], (real)means3D[3 * idx + 1] - campos[1], (real)means3D[3 * idx + 2] - campos[2] };
			const real len = (real)sqrt((double)(dir_orig[0] * dir_orig[0] + dir_orig[1] * dir_orig[1] + dir_orig[2] * dir_orig[2]));
			const real dx_ = dir_orig[0] / len, dy_ = dir_orig[1] / len, dz_ = dir_orig[2] / len;
			const real X = dx_, Y = dy_, Z = dz_;
			real ddir[3] = { 0, 0, 0 };
			for (int ch = 0; ch < 3; ch++) {
				const real g = dL_dcolor[3 * idx + ch] * (clamped[3 * idx + ch] ? 0.f : 1.f);
				real ddx = 0, ddy = 0, ddz = 0;
#define SHV(k) sh[3 * (k) + ch]
#define DSH(k) dsh[3 * (k) + ch]
				DSH(0) = SH_C0 * g;
				if (D > 0) {
					DSH(1) = -SH_C1 * Y * g; DSH(2) = SH_C1 * Z * g; DSH(3) = -SH_C1 * X * g;
					ddx = -SH_C1 * SHV(3); ddy = -SH_C1 * SHV(1); ddz = SH_C1 * SHV(2);
					if (D > 1) {
						const real xx = X * X, yy = Y * Y, zz = Z * Z, xy = X * Y, yz = Y * Z, xz = X * Z;
						DSH(4) = SH_C2[0] * xy * g; DSH(5) = SH_C2[1] * yz * g; DSH(6) = SH_C2[2] * (2.f * zz - xx - yy) * g;
						DSH(7) = SH_C2[3] * xz * g; DSH(8) = SH_C2[4] * (xx - yy) * g;
						ddx += SH_C2[0] * Y * SHV(4) + SH_C2[2] * 2.f * -X * SHV(6) + SH_C2[3] * Z * SHV(7) + SH_C2[4] * 2.f * X * SHV(8);
						ddy += SH_C2[0] * X * SHV(4) + SH_C2[1] * Z * SHV(5) + SH_C2[2] * 2.f * -Y * SHV(6) + SH_C2[4] * 2.f * -Y * SHV(8);
						ddz += SH_C2[1] * Y * SHV(5) + SH_C2[2] * 2.f * 2.f * Z * SHV(6) + SH_C2[3] * X * SHV(7);
						if (D > 2) {
							DSH(9) = SH_C3[0] * Y * (3.f * xx - yy) * g; DSH(10) = SH_C3[1] * xy * Z * g;
							DSH(11) = SH_C3[2] * Y * (4.f * zz - xx - yy) * g;
							DSH(12) = SH_C3[3] * Z * (2.f * zz - 3.f * xx - 3.f * yy) * g;
							DSH(13) = SH_C3[4] * X * (4.f * zz - xx - yy) * g; DSH(14) = SH_C3[5] * Z * (xx - yy) * g;
							DSH(15) = SH_C3[6] * X * (xx - 3.f * yy) * g;
							ddx += SH_C3[0] * SHV(9) * 3.f * 2.f * xy + SH_C3[1] * SHV(10) * yz + SH_C3[2] * SHV(11) * -2.f * xy +
							       SH_C3[3] * SHV(12) * -3.f * 2.f * xz + SH_C3[4] * SHV(13) * (-3.f * xx + 4.f * zz - yy) +
							       SH_C3[5] * SHV(14) * 2.f * xz + SH_C3[6] * SHV(15) * 3.f * (xx - yy);
							ddy += SH_C3[0] * SHV(9) * 3.f * (xx - yy) + SH_C3[1] * SHV(10) * xz +
							       SH_C3[2] * SHV(11) * (-3.f * yy + 4.f * zz - xx) + SH_C3[3] * SHV(12) * -3.f * 2.f * yz +
							       SH_C3[4] * SHV(13) * -2.f * xy + SH_C3[5] * SHV(14) * -2.f * yz + SH_C3[6] * SHV(15) * -3.f * 2.f * xy;
							ddz += SH_C3[1] * SHV(10) * xy + SH_C3[2] * SHV(11) * 4.f * 2.f * yz +
							       SH_C3[3] * SHV(12) * 3.f * (2.f * zz - xx - yy) + SH_C3[4] * SHV(13) * 4.f * 2.f * xz +
							       SH_C3[5] * SHV(14) * (xx - yy);
						}
					}
				}
#undef SHV
#undef DSH
				ddir[0] += ddx * g; ddir[1] += ddy * g; ddir[2] += ddz * g;
			}
			real add[3];
			dnormvdv(dir_orig, ddir, add);
			dmean[0] += add[0]; dmean[1] += add[1]; dmean[2] += add[2];
		}
		dL_dmeans[3 * idx] = dmean[0];
		dL_dmeans[3 * idx + 1] = dmean[1];
		dL_dmeans[3 * idx + 2] = dmean[2];
	}
}
