// render_fwd.cu -- per-tile front-to-back GOF alpha compositing (K8).
//
// Replaces renderCUDA<3> forward (RAST/cuda_rasterizer/forward.cu:409-612).  Same outputs:
// out_color[9,H,W] (rgb, view-space normal, median depth, alpha, normalised distortion),
// final_T[4,H,W] = (T, dist1, dist2, distortion_raw), n_contrib[2,H,W] = (last, max contributor).
//
// B200 design:
//   * One CTA per 16x16 tile (the tile size is part of the binning contract), 8 warps, each
//     warp owning an 8x4 pixel block so that a Gaussian's footprint diverges fewer warps than
//     the reference's 16x2 strips.
//   * The tile's sorted Gaussians arrive as a contiguous slab of 64-byte records (binning.cu);
//     one elected thread streams it into a 4-stage shared-memory ring with TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx), so the 256 pixel threads never issue a global
//     load in the loop and the next chunks land while the current one is blended.
//   * Each (pixel, Gaussian) pair first takes the conservative float32 pre-test of
//     blend_math.cuh; only survivors pay for the FP64 ray-minimum and expf.  Pre-tests run
//     4 records at a time for ILP, survivors are then blended in order.
//   * Accumulation order and every rounding of the contributing path follow the reference's
//     sm_100a build, so the forward outputs are bit-identical to it.
#include "blend_math.cuh"

namespace gof {

namespace {

constexpr int CHUNK = 128;                 // records per pipeline stage (8 KB)
constexpr int STAGES = 4;
constexpr int CHUNK_BYTES = CHUNK * REC_BYTES;

struct PixState {
	float T;
	float C[8];
	float dist1, dist2, distortion;
	uint32_t last_contributor, max_contributor;
};

// Blend one surviving pair into the pixel state (forward.cu:536-578).  Returns true when the
// pixel saturates (test_T < 1e-4) -- the pair is then NOT blended, as in the reference.
__device__ __forceinline__ bool blend_pair(PixState& s, const PairGeom& g, float t, float alpha,
                                           const float4& d, uint32_t contributor)
{
	const float T = s.T;
	const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
	if (test_T < 0.0001f) return true;

	// 2DGS NDC depth mapping, in double: (far*t - far*near) / ((far - near) * t)
	const double td = t;
	const float m = (float)(fma(td, 100.0, -(100.0 * 0.2)) / ((100.0 - 0.2) * td));

	const float len2 = __fmaf_rn(g.n2, g.n2, __fmaf_rn(g.n0, g.n0, __fmul_rn(g.n1, g.n1)));
	const float length = (float)sqrt((double)len2 + 1e-7);
	const float nn0 = __fdiv_rn(g.n0, length);
	const float nn1 = __fdiv_rn(g.n1, length);
	const float nn2 = __fdiv_rn(g.n2, length);

	const float A1 = __fsub_rn(1.0f, T);
	const float m2 = __fmul_rn(m, m);
	const float err = __fmaf_rn(-s.dist1, __fadd_rn(m, m), __fmaf_rn(A1, m2, s.dist2));
	s.distortion = __fmaf_rn(T, __fmul_rn(alpha, err), s.distortion);
	s.dist1 = __fmaf_rn(T, __fmul_rn(alpha, m), s.dist1);
	s.dist2 = __fmaf_rn(T, __fmul_rn(alpha, m2), s.dist2);

	s.C[0] = __fmaf_rn(T, __fmul_rn(alpha, d.x), s.C[0]);
	s.C[1] = __fmaf_rn(T, __fmul_rn(alpha, d.y), s.C[1]);
	s.C[2] = __fmaf_rn(T, __fmul_rn(alpha, d.z), s.C[2]);
	// view-space normal is -n/|n|
	s.C[3] = __fmaf_rn(-T, __fmul_rn(alpha, nn0), s.C[3]);
	s.C[4] = __fmaf_rn(-T, __fmul_rn(alpha, nn1), s.C[4]);
	s.C[5] = __fmaf_rn(-T, __fmul_rn(alpha, nn2), s.C[5]);
	if (T > 0.5f) {           // median depth: last Gaussian seen while T > 0.5
		s.C[6] = t;
		s.max_contributor = contributor;
	}
	s.C[7] = __fmaf_rn(T, alpha, s.C[7]);
	s.T = test_T;
	s.last_contributor = contributor;
	return false;
}

__global__ void __launch_bounds__(TILE_PIX)
render_fwd_kernel(const uint2* __restrict__ ranges, const float* __restrict__ slab, int W, int H,
                  float focal_x, float focal_y, const float* __restrict__ bg_color,
                  float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color)
{
	__shared__ __align__(128) float4 s_rec[STAGES][CHUNK * 4];
	__shared__ __align__(8) uint64_t s_full[STAGES];

	const int tid = threadIdx.x;
	const int warp = tid >> 5, lane = tid & 31;
	const uint32_t px = blockIdx.x * TILE_X + (warp & 1) * 8 + (lane & 7);
	const uint32_t py = blockIdx.y * TILE_Y + (warp >> 1) * 4 + (lane >> 3);
	const bool inside = px < (uint32_t)W && py < (uint32_t)H;
	const uint32_t pix_id = W * py + px;
	const float rx = pixel_ray(px, W, focal_x);
	const float ry = pixel_ray(py, H, focal_y);

	const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];
	const int n = (int)(range.y - range.x);
	const int nchunks = (n + CHUNK - 1) / CHUNK;
	const float* tile_slab = slab + (size_t)range.x * REC_FLOATS;

	if (tid == 0) {
#pragma unroll
		for (int s = 0; s < STAGES; s++) mbar_init(&s_full[s], 1);
		mbar_fence_init();
	}
	__syncthreads();

	auto issue = [&](int c) {
		const int s = c % STAGES;
		const int cnt = min(CHUNK, n - c * CHUNK);
		const uint32_t bytes = (uint32_t)cnt * REC_BYTES;
		mbar_arrive_expect_tx(&s_full[s], bytes);
		tma_bulk_g2s(&s_rec[s][0], tile_slab + (size_t)c * CHUNK * REC_FLOATS, bytes, &s_full[s]);
	};
	if (tid == 0) {
		const int pre = min(STAGES, nchunks);
		for (int c = 0; c < pre; c++) issue(c);
	}

	PixState st;
	st.T = 1.0f;
#pragma unroll
	for (int k = 0; k < 8; k++) st.C[k] = 0.0f;
	st.dist1 = st.dist2 = st.distortion = 0.0f;
	st.last_contributor = 0;
	st.max_contributor = 0xFFFFFFFFu;   // uint(-1), forward.cu:464
	bool done = !inside;

	int c = 0;
	for (; c < nchunks; c++) {
		const int s = c % STAGES;
		mbar_wait(&s_full[s], (uint32_t)((c / STAGES) & 1));
		const int cnt = min(CHUNK, n - c * CHUNK);
		const float4* rec = &s_rec[s][0];
		const uint32_t base = (uint32_t)c * CHUNK;

		if (!done) {
			for (int j0 = 0; j0 < cnt && !done; j0 += 4) {
				// pre-test 4 records (independent -> ILP), collect survivors in a bit mask
				uint32_t mask = 0;
#pragma unroll
				for (int k = 0; k < 4; k++) {
					const int j = j0 + k;
					if (j < cnt) {
						const float4 a = rec[4 * j + 0], b = rec[4 * j + 1], cc = rec[4 * j + 2];
						const PairGeom g = pair_geom(a, b, cc, rx, ry);
						if (!pair_pretest_reject(g, cc.y, cc.z)) mask |= 1u << k;
					}
				}
				while (mask) {
					const int k = __ffs(mask) - 1;
					mask &= mask - 1;
					const int j = j0 + k;
					const float4 a = rec[4 * j + 0], b = rec[4 * j + 1], cc = rec[4 * j + 2];
					const PairGeom g = pair_geom(a, b, cc, rx, ry);
					float t, alpha, G;
					if (!pair_alpha_exact(g, cc.y, cc.w, t, alpha, G)) continue;
					const float4 d = rec[4 * j + 3];
					if (blend_pair(st, g, t, alpha, d, base + j + 1)) { done = true; break; }
				}
			}
		}
		// Everyone is past this stage: vote for early exit, then refill the stage.
		const int num_done = __syncthreads_count(done);
		if (num_done == TILE_PIX) break;
		if (tid == 0 && c + STAGES < nchunks) issue(c + STAGES);
	}
	// Drain copies that were issued but never consumed (early exit) before the CTA retires.
	if (tid == 0 && c < nchunks) {
		const int issued = min(nchunks, c + STAGES);
		for (int cc = c + 1; cc < issued; cc++) mbar_wait(&s_full[cc % STAGES], (uint32_t)((cc / STAGES) & 1));
	}

	if (inside) {
		const size_t N = (size_t)W * H;
		const float T = st.T;
		const float om = __fsub_rn(1.0f, T);
		const float dnorm = (float)((double)st.distortion / ((double)__fmul_rn(om, om) + 1e-7));
		final_T[pix_id] = T;
		final_T[pix_id + N] = st.dist1;
		final_T[pix_id + 2 * N] = st.dist2;
		final_T[pix_id + 3 * N] = st.distortion;
		n_contrib[pix_id] = st.last_contributor;
		n_contrib[pix_id + N] = st.max_contributor;
#pragma unroll
		for (int ch = 0; ch < 3; ch++) out_color[ch * N + pix_id] = __fmaf_rn(T, bg_color[ch], st.C[ch]);
#pragma unroll
		for (int ch = 3; ch < 8; ch++) out_color[ch * N + pix_id] = st.C[ch];
		out_color[CH_DIST * N + pix_id] = dnorm;
	}
}

}  // namespace

int launch_render_fwd(const GofParams& prm, dim3 tile_grid, float focal_x, float focal_y,
                      const ImgState& im, const BinState& b, const float* background,
                      float* out_color, cudaStream_t s)
{
	render_fwd_kernel<<<tile_grid, TILE_PIX, 0, s>>>(im.ranges, b.slab, prm.W, prm.H, focal_x, focal_y,
	                                                background, im.final_T, im.n_contrib, out_color);
	GOF_CUDA_CHECK(cudaGetLastError());
	return GOF_OK;
}

}  // namespace gof
