// render_fwd.cu -- per-tile front-to-back GOF alpha compositing (K8).
//
// Replaces renderCUDA<3> forward (RAST/cuda_rasterizer/forward.cu:409-612).  Same outputs:
// out_color[9,H,W] (rgb, view-space normal, median depth, alpha, normalised distortion),
// final_T[4,H,W] = (T, dist1, dist2, distortion_raw), n_contrib[2,H,W] = (last, max contributor).
//
// B200 design:
//   * One CTA per 16x16 tile (the tile size is part of the binning contract): 8 consumer warps,
//     each owning an 8x4 pixel block, plus one producer warp.
//   * The tile's sorted Gaussians arrive as a contiguous slab of 80-byte records (binning.cu).
//     The producer streams it into a 4-stage shared-memory ring with TMA bulk copies
//     (cp.async.bulk, completion on a "full" mbarrier per stage); each consumer warp releases a
//     stage by arriving on its "empty" mbarrier.  There is no CTA-wide barrier in the loop: warps
//     drift up to three chunks apart, so a warp that has many contributors in one chunk does not
//     stall the other seven (the reference, and our first version, synchronise every batch).
//   * Per 128-record chunk, two passes.  Pass 1 is a branch-free sweep: every pixel evaluates the
//     tile-local CONIC pre-test (conic.cuh: a quadratic in the pixel's tile coordinates, five FMAs,
//     coefficients broadcast from shared memory) for the 128 records and keeps the survivors as a
//     128-bit mask.  Pass 2 is lane-private: each pixel walks ITS OWN survivors in list order
//     (float32 quadric, the tight float32 pre-test, exact FP64 ray minimum, expf, blend).  A warp
//     spends pass-2 iterations equal to its busiest pixel's survivor count over the whole chunk
//     (~11% of the list) instead of running the expensive path for every record any pixel touches.
//   * Rounding: alpha, T, rgb, median depth, alpha channel and the contributor counters follow
//     the reference's sm_100a build operation by operation and are bit-identical to it in both
//     modes.  With GOF_FLAG_EXACT_BLEND the depth mapping and the normal normalisation also use
//     its IEEE double divide / double sqrt / float divides (all nine channels bit-identical);
//     without it they use float32 reciprocal arithmetic (normals, distortion within ~1e-6).
#include "blend_math.cuh"
#include "conic.cuh"

namespace gof {

namespace {

#ifndef GOF_FWD_STAGES
#define GOF_FWD_STAGES 4
#endif
#ifndef GOF_FWD_MIN_CTAS
#define GOF_FWD_MIN_CTAS 3
#endif
constexpr int CHUNK = 128;                 // records per pipeline stage (10 KB)
constexpr int STAGES = GOF_FWD_STAGES;

struct PixState {
	float T;
	float C[8];
	float dist1, dist2, distortion;
	uint32_t last_contributor, max_contributor;
};

// Blend one surviving pair into the pixel state (forward.cu:536-578).  Returns true when the
// pixel saturates (test_T < 1e-4) -- the pair is then NOT blended, as in the reference.
template <bool EXACT>
__device__ __forceinline__ bool blend_pair(PixState& s, const PairGeom& g, float t, float alpha,
                                           const float4& d, uint32_t contributor)
{
	const float T = s.T;
	const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
	if (test_T < 0.0001f) return true;

	float m, nn0, nn1, nn2;
	const float len2 = __fmaf_rn(g.n2, g.n2, __fmaf_rn(g.n0, g.n0, __fmul_rn(g.n1, g.n1)));
	if (EXACT) {
		// 2DGS NDC depth mapping, in double: (far*t - far*near) / ((far - near) * t)
		const double td = t;
		m = (float)(fma(td, 100.0, -(100.0 * 0.2)) / ((100.0 - 0.2) * td));
		const float length = (float)sqrt((double)len2 + 1e-7);
		nn0 = __fdiv_rn(g.n0, length);
		nn1 = __fdiv_rn(g.n1, length);
		nn2 = __fdiv_rn(g.n2, length);
	} else {
		// same quantities with float32 reciprocals: far/(far-near) - (far*near/(far-near)) / t
		m = __fmaf_rn(-(float)((100.0 * 0.2) / (100.0 - 0.2)), rcp_approx(t), (float)(100.0 / (100.0 - 0.2)));
		const float inv_len = rsqrt_approx(len2 + 1e-7f);
		nn0 = g.n0 * inv_len;
		nn1 = g.n1 * inv_len;
		nn2 = g.n2 * inv_len;
	}

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

constexpr int CONSUMER_WARPS = TILE_PIX / 32;          // 8
constexpr int FWD_THREADS = TILE_PIX + 32;             // + 1 producer warp
constexpr int REC_F4 = SLAB_FLOATS / 4;                // float4 per slab record (6)

template <bool EXACT>
__global__ void __launch_bounds__(FWD_THREADS, GOF_FWD_MIN_CTAS)
render_fwd_kernel(const uint2* __restrict__ ranges, const float* __restrict__ slab, int W, int H,
                  float focal_x, float focal_y, const float* __restrict__ bg_colors, int bg_stride,
                  float* __restrict__ final_T_all, uint32_t* __restrict__ n_contrib_all, float* __restrict__ out_color_all,
                  const int32_t* __restrict__ mailbox)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	float4 (*s_rec)[CHUNK * REC_F4] = reinterpret_cast<float4 (*)[CHUNK * REC_F4]>(smem_raw);
	uint64_t* s_full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * CHUNK * SLAB_BYTES);
	uint64_t* s_empty = s_full + STAGES;

	const int tid = threadIdx.x;
	const int warp = tid >> 5, lane = tid & 31;
	const int view = blockIdx.z;
	const size_t N = (size_t)W * H;

	const uint2 range = ranges[((size_t)view * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x];
	// sync-free mode: if the binning blob was too small nothing was binned -- blend empty lists (the
	// caller sees the overflow flag in the mailbox and re-runs the batch with a larger blob)
	const int n = mailbox[1] ? 0 : (int)(range.y - range.x);
	const int nchunks = (n + CHUNK - 1) / CHUNK;
	const float* tile_slab = slab + (size_t)range.x * SLAB_FLOATS;

	if (tid == 0) {
#pragma unroll
		for (int s = 0; s < STAGES; s++) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], CONSUMER_WARPS); }
		mbar_fence_init();
	}
	__syncthreads();

	if (warp == CONSUMER_WARPS) {
		// ---------------- producer warp: one elected lane streams the slab ----------------------
		if (lane == 0) {
			for (int c = 0; c < nchunks; c++) {
				const int s = c % STAGES;
				if (c >= STAGES) mbar_wait_backoff(&s_empty[s], (uint32_t)(((c / STAGES) - 1) & 1));
				const int cnt = min(CHUNK, n - c * CHUNK);
				const uint32_t bytes = (uint32_t)cnt * SLAB_BYTES;
				mbar_arrive_expect_tx(&s_full[s], bytes);
				tma_bulk_g2s(&s_rec[s][0], tile_slab + (size_t)c * CHUNK * SLAB_FLOATS, bytes, &s_full[s]);
			}
		}
		return;
	}

	// -------------------- consumer warps: 8x4 pixel block each --------------------------------
	const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);   // tile-local pixel
	const uint32_t px = blockIdx.x * TILE_X + lx;
	const uint32_t py = blockIdx.y * TILE_Y + ly;
	const bool inside = px < (uint32_t)W && py < (uint32_t)H;
	const uint32_t pix_id = W * py + px;
	const float rx = pixel_ray(px, W, focal_x);
	const float ry = pixel_ray(py, H, focal_y);
	const float fx = (float)lx, fy = (float)ly;

	PixState st;
	st.T = 1.0f;
#pragma unroll
	for (int k = 0; k < 8; k++) st.C[k] = 0.0f;
	st.dist1 = st.dist2 = st.distortion = 0.0f;
	st.last_contributor = 0;
	st.max_contributor = 0xFFFFFFFFu;   // uint(-1), forward.cu:464
	bool done = !inside;
	bool warp_done = __all_sync(0xffffffffu, done);

	const uint32_t rec_base = smem_u32(smem_raw);
	for (int c = 0; c < nchunks; c++) {
		const int s = c % STAGES;
		mbar_wait(&s_full[s], (uint32_t)((c / STAGES) & 1));
		if (!warp_done) {
			const int cnt = min(CHUNK, n - c * CHUNK);
			const uint32_t rec = rec_base + (uint32_t)s * (CHUNK * SLAB_BYTES);   // shared-window address of the stage
			const uint32_t base = (uint32_t)c * CHUNK;
			// ---- pass 1: branch-free conic sweep over the chunk (broadcast reads, 5 FMAs per pair) ----
			uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;   // survivors among records [0,32) [32,64) [64,96) [96,128)
#pragma unroll 1
			for (int w = 0; w < CHUNK / 32; w++) {
				const int valid = cnt - 32 * w;
				if (valid <= 0) break;
				uint32_t bits = 0;
				const uint32_t rw = rec + (uint32_t)w * (32 * SLAB_BYTES);
#pragma unroll
				for (int jj = 0; jj < 32; jj++) {
					const float4 k0 = lds128(rw + jj * SLAB_BYTES);
					const float2 k1 = lds64(rw + jj * SLAB_BYTES + 16);
					if (!conic_reject(k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, fx, fy)) bits |= 1u << jj;
				}
				if (valid < 32) bits &= (1u << valid) - 1u;   // stale records beyond the list end
				if (w == 0) m0 = bits; else if (w == 1) m1 = bits; else if (w == 2) m2 = bits; else m3 = bits;
			}
			if (done) { m0 = 0; m1 = 0; m2 = 0; m3 = 0; }
			// ---- pass 2: each pixel blends its own survivors, in list order.  ONE loop over the whole
			// chunk: the warp iterates max-over-lanes(survivors in 128 records) times, m0 is the word being
			// consumed, m1..m3 shift down when it runs empty.
			// `cur` = the word being consumed, `wsel` its index; m1..m3 stay loop-invariant.
			uint32_t cur = m0, wsel = 0;
			for (;;) {
				if (cur == 0) {   // next non-empty word of this lane, or leave the loop
					if (wsel < 1 && m1 != 0) { cur = m1; wsel = 1; }
					else if (wsel < 2 && m2 != 0) { cur = m2; wsel = 2; }
					else if (wsel < 3 && m3 != 0) { cur = m3; wsel = 3; }
					else break;
				}
				const uint32_t j = 32u * wsel + (uint32_t)__ffs((int)cur) - 1u;
				cur &= cur - 1u;
				const uint32_t r = rec + j * SLAB_BYTES;
				const float4 k1 = lds128(r + 16), k2 = lds128(r + 32), k3 = lds128(r + 48), k4 = lds128(r + 64);
				const PairGeom g = pair_geom(k1, k2, k3, rx, ry);
				float t, alpha, G;
				if (pair_alpha_exact(g, k4.x, k1.z, t, alpha, G)) {
					const float4 d = make_float4(k4.y, k4.z, k4.w, 0.0f);
					if (blend_pair<EXACT>(st, g, t, alpha, d, base + j + 1)) { done = true; break; }
				}
			}
			warp_done = __all_sync(0xffffffffu, done);
		}
		__syncwarp();
		if (lane == 0) mbar_arrive(&s_empty[s]);   // this warp is finished with stage s
	}

	if (inside) {
		float* final_T = final_T_all + (size_t)view * 4 * N;
		uint32_t* n_contrib = n_contrib_all + (size_t)view * 2 * N;
		float* out_color = out_color_all + (size_t)view * OUT_CH * N;
		const float* bg_color = bg_colors + (size_t)view * bg_stride;
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

int launch_render_fwd(const GofParams& prm, const Frame& f, const GeomState& g, const ImgState& im, const BinState& b,
                      const float* background, int bg_stride, float* out_color, cudaStream_t s)
{
	const dim3 grid(f.grid.x, f.grid.y, f.V);
	const size_t smem = (size_t)STAGES * CHUNK * SLAB_BYTES + 2 * STAGES * sizeof(uint64_t);
	static bool attr_set = false;
	if (!attr_set) {
		GOF_CUDA_CHECK(cudaFuncSetAttribute(render_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		GOF_CUDA_CHECK(cudaFuncSetAttribute(render_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		attr_set = true;
	}
	if (prm.flags & GOF_FLAG_EXACT_BLEND)
		render_fwd_kernel<true><<<grid, FWD_THREADS, smem, s>>>(im.ranges, b.slab, prm.W, prm.H, f.focal_x, f.focal_y,
		                                                     background, bg_stride, im.final_T, im.n_contrib, out_color, g.mailbox);
	else
		render_fwd_kernel<false><<<grid, FWD_THREADS, smem, s>>>(im.ranges, b.slab, prm.W, prm.H, f.focal_x, f.focal_y,
		                                                      background, bg_stride, im.final_T, im.n_contrib, out_color, g.mailbox);
	GOF_CUDA_CHECK(cudaGetLastError());
	return GOF_OK;
}

}  // namespace gof
