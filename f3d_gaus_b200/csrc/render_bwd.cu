// render_bwd.cu -- backward of the per-tile GOF blend (K9).
//
// Replaces renderCUDA<3> backward (RAST/cuda_rasterizer/backward.cu:634-955): back-to-front
// replay of each pixel's contributors, producing per-Gaussian gradients w.r.t. colour,
// opacity, the 2-D mean (densification statistics) and the 10-float view2gaussian quadric.
// Reference behaviours kept on purpose: the alpha channel (7) has no gradient,
// the distortion gradient flows only through the mapped depth (dL_dweight is detached), the
// power/alpha clamps are not gated.
//
// B200 design (same skeleton as the forward blend, render_fwd.cu):
//   * one CTA per tile: 8 consumer warps (8x4 pixels each) + one producer warp; the tile's
//     slab is streamed by TMA bulk copies into a shared-memory ring, walked from the tile's deepest
//     contributor (block-max of last_contributor) towards the front -- chunks behind it are never
//     loaded; full/empty mbarriers per stage, no CTA-wide barrier in the loop;
//   * per 128-record chunk, pass 1 is the branch-free conic sweep (conic.cuh) that leaves each
//     pixel a 128-bit survivor mask, pass 2 is lane-private: each pixel walks ITS survivors from
//     the back, re-evaluates the exact alpha (blend_math.cuh -- the contributing set therefore
//     equals the forward's by construction) and forms the 17 partial gradients of the pair;
//   * the reference issues 17 scalar global atomics per contributing (pixel, Gaussian) pair; here a
//     pair issues four 128-bit vector reductions (red.global.add.v4.f32) + one scalar into a packed
//     80-byte per-Gaussian accumulator that the backward preprocess unpacks.  (Shared-memory
//     accumulation was measured and rejected: float atomic-add has no native shared-memory form on
//     sm_100a -- it compiles to a compare-and-swap spin loop, ATOMS.CAST.SPIN -- and was slower.)
#include "blend_math.cuh"
#include "conic.cuh"

namespace gof {

namespace {

#ifndef GOF_BWD_MIN_CTAS
#define GOF_BWD_MIN_CTAS 2
#endif
constexpr int CHUNK = 128;
constexpr int STAGES = 3;
constexpr int CONSUMER_WARPS = TILE_PIX / 32;
constexpr int BWD_THREADS = TILE_PIX + 32;
constexpr int STAGE_REC_BYTES = CHUNK * SLAB_BYTES;      // 10 KB
constexpr size_t BWD_SMEM = (size_t)STAGES * STAGE_REC_BYTES + 2 * STAGES * sizeof(uint64_t) + 64;

__device__ __forceinline__ void red_global_v4(float* addr, float4 v)
{
	asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// 1/x for a float x to ~44 bits: one MUFU.RCP (<= 1 ulp) + one Newton step in double (2 DFMA instead of an IEEE
// double division; only used for gradient VALUES, never for alpha / the contributing set).
#ifndef GOF_BWD_F32_QUADRIC
#define GOF_BWD_F32_QUADRIC 1     // dL/dA, dL/dB in float32 (0: double, as the reference; no measurable accuracy difference)
#endif
#ifndef GOF_BWD_IEEE_RCP
#define GOF_BWD_IEEE_RCP 0        // 1: the correctly rounded float reciprocal / sqrt sequences (A/B switch)
#endif
__device__ __forceinline__ float rcp_f32(float x) { return GOF_BWD_IEEE_RCP ? __frcp_rn(x) : rcp_approx(x); }
#if GOF_BWD_IEEE_RCP || !GOF_BWD_F32_QUADRIC
__device__ __forceinline__ double rcp_refined(float x)
{
	const double r = (double)rcp_f32(x);
	return fma(r, fma(-(double)x, r, 1.0), r);
}
#endif

__global__ void __launch_bounds__(BWD_THREADS, GOF_BWD_MIN_CTAS)
render_bwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order, int tiles_per_view, int tiles_x,
                  const float* __restrict__ slab, const uint32_t* __restrict__ point_list, int P, int W, int H,
                  float focal_x, float focal_y, const float* __restrict__ bg_colors, int bg_stride,
                  const float2* __restrict__ means2D_all, const float4* __restrict__ conic_opacity_all,
                  const float* __restrict__ final_Ts_all, const uint32_t* __restrict__ n_contrib_all,
                  const float* __restrict__ dL_dpixels_all, float* __restrict__ gacc_all)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	// layout: [STAGES] record stages | full[STAGES] | empty[STAGES] | s_max[8]
	const uint32_t rec_base = smem_u32(smem_raw);
	uint64_t* s_full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * STAGE_REC_BYTES);
	uint64_t* s_empty = s_full + STAGES;
	uint32_t* s_max = reinterpret_cast<uint32_t*>(s_empty + STAGES);

	const int tid = threadIdx.x;
	const int warp = tid >> 5, lane = tid & 31;
	pdl_trigger();                 // the per-Gaussian backward may be scheduled behind this grid
	// CTAs are launched longest-list-first: blockIdx.x -> (view, tile) through tile_order (render_fwd.cu)
	const uint32_t gt = tile_order[blockIdx.x];
	const int view = (int)(gt / (uint32_t)tiles_per_view);
	const int tile = (int)(gt - (uint32_t)view * (uint32_t)tiles_per_view);
	const int tile_y = tile / tiles_x, tile_x = tile - tile_y * tiles_x;
	const bool consumer = warp < CONSUMER_WARPS;
	const int lx = (warp & 1) * 8 + (lane & 7), ly = ((warp >> 1) & 3) * 4 + (lane >> 3);
	const uint32_t px = tile_x * TILE_X + lx;
	const uint32_t py = tile_y * TILE_Y + ly;
	const bool inside = consumer && px < (uint32_t)W && py < (uint32_t)H;
	const uint32_t pix_id = W * py + px;
	const size_t N = (size_t)W * H;

	const float* bg_color = bg_colors + (size_t)view * bg_stride;
	const float2* means2D = means2D_all + (size_t)view * P;
	const float4* conic_opacity = conic_opacity_all + (size_t)view * P;
	const float* final_Ts = final_Ts_all + (size_t)view * 4 * N;
	const uint32_t* n_contrib = n_contrib_all + (size_t)view * 2 * N;
	const float* dL_dpixels = dL_dpixels_all + (size_t)view * OUT_CH * N;
	float* gacc = gacc_all + (size_t)view * P * GACC_FLOATS;

	const uint2 range = ranges[gt];
	const int n = (int)(range.y - range.x);
	const float* tile_slab = slab + (size_t)range.x * SLAB_FLOATS;
	const uint32_t* tile_ids = point_list + range.x;

	const uint32_t last_contributor = inside ? n_contrib[pix_id] : 0;
	const uint32_t max_contributor = inside ? n_contrib[pix_id + N] : 0;

	// Deepest record any pixel of this tile blended; init barriers.
	const uint32_t wmax = __reduce_max_sync(0xffffffffu, last_contributor);
	if (consumer && lane == 0) s_max[warp] = wmax;
	if (tid == 0) {
#pragma unroll
		for (int s = 0; s < STAGES; s++) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], CONSUMER_WARPS); }
		mbar_fence_init();
	}
	__syncthreads();
	uint32_t tile_last = 0;
#pragma unroll
	for (int k = 0; k < CONSUMER_WARPS; k++) tile_last = max(tile_last, s_max[k]);
	const int m = min((int)tile_last, n);           // records [0, m) may contribute
	const int nchunks = (m + CHUNK - 1) / CHUNK;     // walked from chunk nchunks-1 down to 0
	// the i-th chunk in walk order is chunk (nchunks-1-i); it lives in stage i % STAGES

	if (!consumer) {
		// ------------- producer warp: one elected lane streams the slab, deepest chunk first ---------
		if (lane == 0) {
			for (int i = 0; i < nchunks; i++) {
				const int s = i % STAGES;
				if (i >= STAGES) mbar_wait_backoff(&s_empty[s], (uint32_t)(((i / STAGES) - 1) & 1));
				const int c = nchunks - 1 - i;
				const int cnt = min(CHUNK, m - c * CHUNK);
				const uint32_t bytes = (uint32_t)cnt * SLAB_BYTES;
				mbar_arrive_expect_tx(&s_full[s], bytes);
				tma_bulk_g2s(smem_raw + (size_t)s * STAGE_REC_BYTES, tile_slab + (size_t)c * CHUNK * SLAB_FLOATS, bytes, &s_full[s]);
			}
		}
		return;
	}

	// ------------- consumer warps -----------------------------------------------------------------
	const float rx = pixel_ray(px, W, focal_x);
	const float ry = pixel_ray(py, H, focal_y);
	const float fx = (float)lx, fy = (float)ly;

	// Per-pixel state (backward.cu:690-735).
	const float T_final = inside ? final_Ts[pix_id] : 0;
	float T = T_final;
	const float final_D = inside ? final_Ts[pix_id + N] : 0;
	const float final_A = 1 - T_final;
	const float dL_dreg = inside ? dL_dpixels[CH_DIST * N + pix_id] : 0;
	float accum_rec[3] = { 0, 0, 0 }, accum_normal_rec[3] = { 0, 0, 0 };
	float dL_dpixel[3] = { 0, 0, 0 }, dL_dnormal2D[3] = { 0, 0, 0 };
	float dL_dmax_depth = 0;
	if (inside) {
#pragma unroll
		for (int i = 0; i < 3; i++) {
			dL_dpixel[i] = dL_dpixels[i * N + pix_id];
			dL_dnormal2D[i] = dL_dpixels[(3 + i) * N + pix_id];
		}
		dL_dmax_depth = dL_dpixels[CH_DEPTH * N + pix_id];
	}
	float last_alpha = 0;
	float last_color[3] = { 0, 0, 0 }, last_normal[3] = { 0, 0, 0 };
	const float bg_dot_dpixel = bg_color[0] * dL_dpixel[0] + bg_color[1] * dL_dpixel[1] + bg_color[2] * dL_dpixel[2];
	const float ddelx_dx = 0.5 * W;
	const float ddely_dy = 0.5 * H;

	for (int i = 0; i < nchunks; i++) {
		const int c = nchunks - 1 - i;
		const int s = i % STAGES;
		mbar_wait(&s_full[s], (uint32_t)((i / STAGES) & 1));
		const int cnt = min(CHUNK, m - c * CHUNK);
		const uint32_t rec = rec_base + s * STAGE_REC_BYTES;
		const uint32_t base = (uint32_t)c * CHUNK;

		// ---- pass 1: conic sweep; only records this pixel blended (index < last_contributor) ----
		uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
		if (base < last_contributor) {
#pragma unroll 1
			for (int w = 0; w < CHUNK / 32; w++) {
				const int valid = min(cnt, (int)(last_contributor - base)) - 32 * w;
				if (valid <= 0) break;
				uint32_t bits = 0;
				const uint32_t rw = rec + (uint32_t)w * (32 * SLAB_BYTES);
#pragma unroll
				for (int jj = 0; jj < 32; jj++) {
					const float4 k0 = lds128(rw + jj * SLAB_BYTES);
					const float2 k1 = lds64(rw + jj * SLAB_BYTES + 16);
					if (!conic_reject(k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, fx, fy)) bits |= 1u << jj;
				}
				if (valid < 32) bits &= (1u << valid) - 1u;
				if (w == 0) m0 = bits; else if (w == 1) m1 = bits; else if (w == 2) m2 = bits; else m3 = bits;
			}
		}
		// ---- pass 2: this pixel's survivors, back to front; m3 is the word being consumed ----------
		uint32_t jbase = 96;
		while ((m0 | m1 | m2 | m3) != 0) {
			if (m3 == 0) { m3 = m2; m2 = m1; m1 = m0; m0 = 0; jbase -= 32; }
			if (m3 != 0) {
				const uint32_t bit = 31u - (uint32_t)__clz((int)m3);
				m3 &= ~(1u << bit);
				const uint32_t j = jbase + bit;
				const uint32_t r = rec + j * SLAB_BYTES;
				const float4 k1 = lds128(r + 16), k2 = lds128(r + 32), k3 = lds128(r + 48), k4 = lds128(r + 64);
				// the Gaussian's id and its 2-D mean / conic (two dependent L2 round trips) are requested before the exact
				// evaluation so that they are in flight during its double division; nearly every survivor needs them
				const uint32_t contributor = base + j;       // 0-based position in the tile list
				const int gid = (int)__ldg(&tile_ids[contributor]);
				const float2 xy = __ldg(&means2D[gid]);
				const float4 con = __ldg(&conic_opacity[gid]);
				const PairGeom g = pair_geom(k1, k2, k3, rx, ry);
				const float w = k1.z;
				float t, alpha, G;
				double u;                                       // -BB/AA, from the forward's own double division
				if (pair_alpha_exact(g, k4.x, w, t, alpha, G, u)) {
					// Gradient arithmetic: the reference spends ~8 double and ~8 float divisions per pair here
					// (backward.cu:843-925).  The contributing SET and alpha, T are bit-exact (shared with the
					// forward); the gradient VALUES only have to meet the 1e-3 relative bar, so each group of
					// divisions by the same quantity is one reciprocal and multiplies (differences ~1e-7).
#if GOF_BWD_IEEE_RCP
					const double inv_t = rcp_refined(t);
					const float mapped = (float)((100.0 / (100.0 - 0.2)) - ((100.0 * 0.2) / (100.0 - 0.2)) * inv_t);
					const float dmax_t_dd = (float)(((100.0 * 0.2) / (100.0 - 0.2)) * inv_t * inv_t);
#else
					// mapped depth far/(far-near) - (far*near/(far-near))/t and its derivative: well conditioned, float32
					const float inv_t = rcp_approx(t);
					const float kfn = (float)((100.0 * 0.2) / (100.0 - 0.2));
					const float mapped = __fmaf_rn(-kfn, inv_t, (float)(100.0 / (100.0 - 0.2)));
					const float dmax_t_dd = kfn * inv_t * inv_t;
#endif
					const float len2 = g.n0 * g.n0 + g.n1 * g.n1 + g.n2 * g.n2 + 1e-7f;
					const float inv_len = GOF_BWD_IEEE_RCP ? __frcp_rn(sqrtf(len2)) : rsqrt_approx(len2);
					const float nn[3] = { -g.n0 * inv_len, -g.n1 * inv_len, -g.n2 * inv_len };
					const float nraw[3] = { g.n0, g.n1, g.n2 };
					const float inv_1ma = rcp_f32(1.f - alpha);
					float* dst = gacc + (size_t)gid * GACC_FLOATS;
					float gcol[3];

					T = T * inv_1ma;
					const float weight = alpha * T;
					float dL_dalpha = 0.0f;
					const float col[3] = { k4.y, k4.z, k4.w };
#pragma unroll
					for (int ch = 0; ch < 3; ch++) {
						accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
						last_color[ch] = col[ch];
						dL_dalpha += (col[ch] - accum_rec[ch]) * dL_dpixel[ch];
						gcol[ch] = weight * dL_dpixel[ch];
					}
					// distortion: only through the mapped depth (weights detached)
					const float dL_dmax_t = 2.0f * weight * (mapped * final_A - final_D) * dL_dreg * dmax_t_dd;

					float dL_dnn[3];
#pragma unroll
					for (int ch = 0; ch < 3; ch++) {
						accum_normal_rec[ch] = last_alpha * last_normal[ch] + (1.f - last_alpha) * accum_normal_rec[ch];
						last_normal[ch] = nn[ch];
						dL_dalpha += (nn[ch] - accum_normal_rec[ch]) * dL_dnormal2D[ch];
						dL_dnn[ch] = weight * dL_dnormal2D[ch];
					}
					float dL_dlength = dL_dnn[0] * nraw[0] + dL_dnn[1] * nraw[1] + dL_dnn[2] * nraw[2];
					dL_dlength *= inv_len * inv_len;
					float dL_dn[3] = { (-dL_dnn[0] + dL_dlength * nraw[0]) * inv_len,
					                   (-dL_dnn[1] + dL_dlength * nraw[1]) * inv_len,
					                   (-dL_dnn[2] + dL_dlength * nraw[2]) * inv_len };

					float dL_dt = dL_dmax_t;
					if (contributor == max_contributor - 1) dL_dt += dL_dmax_depth;

					dL_dalpha *= T;
					last_alpha = alpha;
					dL_dalpha += (-T_final * inv_1ma) * bg_dot_dpixel;

					const float dL_dG = w * dL_dalpha;
					const float dx = xy.x - (float)px, dy = xy.y - (float)py;
					const float gdx = G * dx, gdy = G * dy;
					const float dG_ddelx = -gdx * con.x - gdy * con.y;
					const float dG_ddely = -gdy * con.z - gdx * con.y;
					const float gmx = dL_dG * dG_ddelx * ddelx_dx;
					const float gmy = dL_dG * dG_ddely * ddely_dy;

					const float dL_dmin_value = dL_dG * G * -0.5f;
#if GOF_BWD_F32_QUADRIC
					const float ba = (float)(-u);                  // BB / AA (from the forward's double division)
					const float half_inv_AA = 0.5f * rcp_approx(g.AA);
					const float dL_dA = __fmaf_rn(dL_dt * ba, half_inv_AA, dL_dmin_value * ba * ba * 0.25f);
					const float dL_dB = __fmaf_rn(-dL_dt, half_inv_AA, dL_dmin_value * ba * -0.5f);
					const float dL_dC = dL_dmin_value;
#else
					const double ba = -u;                          // BB / AA
					const double half_inv_AA = 0.5 * rcp_refined(g.AA);
					double dL_dA = (double)dL_dmin_value * ba * ba * 0.25;
					double dL_dB = (double)dL_dmin_value * ba * -0.5;
					const double dL_dC = dL_dmin_value;
					dL_dA += (double)dL_dt * ba * half_inv_AA;
					dL_dB -= (double)dL_dt * half_inv_AA;
#endif
					dL_dn[0] += dL_dA * rx;
					dL_dn[1] += dL_dA * ry;
					dL_dn[2] += dL_dA;

#ifdef GOF_BWD_DIAG_ONE_RED
					// DIAGNOSTIC build only (tools/build_variant.sh): one reduction per pair instead of five, everything
					// still computed -- separates the arithmetic from the L2-reduction issue rate.  Results are WRONG.
					red_global_v4(dst + 0, make_float4(dL_dn[0] * rx + dL_dn[1] + dL_dn[2] * ry + (float)(dL_dB * 2) + gcol[2],
					                                   dL_dn[0] * ry + dL_dn[1] * rx + dL_dn[2] + (float)dL_dC + G * dL_dalpha,
					                                   dL_dn[0] + dL_dn[2] * rx + (float)(dL_dB * 2 * rx) + gcol[0] + gmx,
					                                   dL_dn[1] * ry + (float)(dL_dB * 2 * ry) + gcol[1] + gmy + fabsf(gmx) + fabsf(gmy)));
#else
					red_global_v4(dst + 0, make_float4(dL_dn[0] * rx, dL_dn[0] * ry + dL_dn[1] * rx, dL_dn[0] + dL_dn[2] * rx, dL_dn[1] * ry));
					red_global_v4(dst + 4, make_float4(dL_dn[1] + dL_dn[2] * ry, dL_dn[2], (float)(dL_dB * 2 * rx), (float)(dL_dB * 2 * ry)));
					red_global_v4(dst + 8, make_float4((float)(dL_dB * 2), (float)dL_dC, gcol[0], gcol[1]));
					red_global_v4(dst + 12, make_float4(gcol[2], G * dL_dalpha, gmx, gmy));
					atomicAdd(dst + 16, fabsf(gmx) + fabsf(gmy));
#endif
				}
			}
		}
		__syncwarp();
		if (lane == 0) mbar_arrive(&s_empty[s]);   // this warp is finished with stage s
	}
}

}  // namespace

int launch_render_bwd(const GofParams& prm, const Frame& f, const GeomState& g, const ImgState& im,
                      const BinState& b, const float* background, int bg_stride, const float* dL_dpix, float* gacc,
                      cudaStream_t s)
{
	const dim3 grid((unsigned)(f.T * f.V), 1, 1);
	GOF_CUDA_CHECK(cudaFuncSetAttribute(render_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
	render_bwd_kernel<<<grid, BWD_THREADS, BWD_SMEM, s>>>(im.ranges, im.tile_order, f.T, (int)f.grid.x, b.slab, b.point_list, f.P, prm.W, prm.H, f.focal_x, f.focal_y, background, bg_stride,
	                                                     g.means2D, g.conic_opacity, im.final_T, im.n_contrib, dL_dpix, gacc);
	GOF_CUDA_CHECK(cudaGetLastError());
	return GOF_OK;
}

}  // namespace gof
