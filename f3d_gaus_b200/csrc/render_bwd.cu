// render_bwd.cu -- backward of the per-tile GOF blend (K9).
//
// Replaces renderCUDA<3> backward (RAST/cuda_rasterizer/backward.cu:634-955): back-to-front
// replay of each pixel's contributors, producing per-Gaussian gradients w.r.t. colour,
// opacity, the 2-D mean (densification statistics) and the 10-float view2gaussian quadric.
// Reference behaviours kept on purpose: the alpha channel (7) has no gradient,
// the distortion gradient flows only through the mapped depth (dL_dweight is detached), the
// power/alpha clamps are not gated.
//
// B200 design:
//   * same TMA-streamed slab as the forward, walked from the tile's deepest contributor
//     (block-max of last_contributor) towards the front; chunks behind it are never loaded;
//   * same float32 pre-test + exact alpha as the forward (blend_math.cuh), so the set of
//     contributing pairs is identical to the forward's by construction;
//   * the reference issues 17 scalar global atomics per contributing (pixel, Gaussian) pair.
//     Here all lanes of a warp visit the same Gaussian in lock-step, the 17 partial gradients
//     are summed across the warp with shuffles, and one lane issues five 128-bit vector
//     reductions (red.global.add.v4.f32) into a packed 80-byte per-Gaussian accumulator.
//     Warps with <= 2 contributing lanes skip the shuffle tree and reduce directly.
#include "blend_math.cuh"
#include "conic.cuh"

namespace gof {

namespace {

constexpr int CHUNK = 128;
constexpr int STAGES = 3;
constexpr int REC_F4 = SLAB_FLOATS / 4;

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d)
{
	asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ float warp_sum(float v)
{
	v += __shfl_xor_sync(0xffffffffu, v, 16);
	v += __shfl_xor_sync(0xffffffffu, v, 8);
	v += __shfl_xor_sync(0xffffffffu, v, 4);
	v += __shfl_xor_sync(0xffffffffu, v, 2);
	v += __shfl_xor_sync(0xffffffffu, v, 1);
	return v;
}

__global__ void __launch_bounds__(TILE_PIX)
render_bwd_kernel(const uint2* __restrict__ ranges, const float* __restrict__ slab, int P, int W, int H,
                  float focal_x, float focal_y, const float* __restrict__ bg_colors, int bg_stride,
                  const float2* __restrict__ means2D_all, const float4* __restrict__ conic_opacity_all,
                  const float* __restrict__ final_Ts_all, const uint32_t* __restrict__ n_contrib_all,
                  const float* __restrict__ dL_dpixels_all, float* __restrict__ gacc_all)
{
	__shared__ __align__(128) float4 s_rec[STAGES][CHUNK * REC_F4];
	__shared__ __align__(8) uint64_t s_full[STAGES];
	__shared__ uint32_t s_max[TILE_PIX / 32];

	const int tid = threadIdx.x;
	const int warp = tid >> 5, lane = tid & 31;
	const int view = blockIdx.z;
	const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);
	const uint32_t px = blockIdx.x * TILE_X + lx;
	const uint32_t py = blockIdx.y * TILE_Y + ly;
	const bool inside = px < (uint32_t)W && py < (uint32_t)H;
	const uint32_t pix_id = W * py + px;
	const size_t N = (size_t)W * H;
	const float rx = pixel_ray(px, W, focal_x);
	const float ry = pixel_ray(py, H, focal_y);
	const float fx = (float)lx, fy = (float)ly;

	const float* bg_color = bg_colors + (size_t)view * bg_stride;
	const float2* means2D = means2D_all + (size_t)view * P;
	const float4* conic_opacity = conic_opacity_all + (size_t)view * P;
	const float* final_Ts = final_Ts_all + (size_t)view * 4 * N;
	const uint32_t* n_contrib = n_contrib_all + (size_t)view * 2 * N;
	const float* dL_dpixels = dL_dpixels_all + (size_t)view * OUT_CH * N;
	float* gacc = gacc_all + (size_t)view * P * GACC_FLOATS;

	const uint2 range = ranges[((size_t)view * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x];
	const int n = (int)(range.y - range.x);
	const float* tile_slab = slab + (size_t)range.x * SLAB_FLOATS;

	const uint32_t last_contributor = inside ? n_contrib[pix_id] : 0;
	const uint32_t max_contributor = inside ? n_contrib[pix_id + N] : 0;

	// Deepest record any pixel of this tile blended.
	uint32_t wmax = __reduce_max_sync(0xffffffffu, last_contributor);
	if (lane == 0) s_max[warp] = wmax;
	if (tid == 0) {
#pragma unroll
		for (int s = 0; s < STAGES; s++) mbar_init(&s_full[s], 1);
		mbar_fence_init();
	}
	__syncthreads();
	uint32_t tile_last = 0;
#pragma unroll
	for (int k = 0; k < TILE_PIX / 32; k++) tile_last = max(tile_last, s_max[k]);
	const int m = min((int)tile_last, n);          // records [0, m) may contribute
	const int nchunks = (m + CHUNK - 1) / CHUNK;    // walked from chunk nchunks-1 down to 0

	// i-th chunk in walk order is chunk (nchunks-1-i); it lives in stage i % STAGES.
	auto issue = [&](int i) {
		const int c = nchunks - 1 - i;
		const int s = i % STAGES;
		const int cnt = min(CHUNK, m - c * CHUNK);
		const uint32_t bytes = (uint32_t)cnt * SLAB_BYTES;
		mbar_arrive_expect_tx(&s_full[s], bytes);
		tma_bulk_g2s(&s_rec[s][0], tile_slab + (size_t)c * CHUNK * SLAB_FLOATS, bytes, &s_full[s]);
	};
	if (tid == 0) {
		const int pre = min(STAGES, nchunks);
		for (int i = 0; i < pre; i++) issue(i);
	}

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
		const float4* rec = &s_rec[s][0];
		const uint32_t base = (uint32_t)c * CHUNK;

		// walk the chunk back to front, 4 records per pre-test group
		for (int j1 = cnt; j1 > 0; j1 -= 4) {
			uint32_t mask = 0;
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const int j = j1 - 1 - k;
				if (j >= 0 && base + j < last_contributor) {
					const float4 k0 = rec[REC_F4 * j];
					const float2 k1 = *reinterpret_cast<const float2*>(&rec[REC_F4 * j + 1]);
					if (!conic_reject(k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, fx, fy)) mask |= 1u << k;
				}
			}
			uint32_t wmask = __reduce_or_sync(0xffffffffu, mask);
			while (wmask) {   // warp-uniform loop over records some lane must evaluate exactly
				const int k = __ffs(wmask) - 1;
				wmask &= wmask - 1;
				const int j = j1 - 1 - k;
				const float4 k1 = rec[REC_F4 * j + 1], a = rec[REC_F4 * j + 2], b = rec[REC_F4 * j + 3], cc = rec[REC_F4 * j + 4];
				const float4 k5 = rec[REC_F4 * j + 5];
				const float4 d = make_float4(cc.z, cc.w, k5.x, k5.y);   // rgb, id
				const PairGeom g = pair_geom(a, b, cc, rx, ry);
				float t = 0, alpha = 0, G = 0;
				bool contrib = false;
				if ((mask & (1u << k)) && !pair_pretest_reject(g, cc.y, k1.z)) contrib = pair_alpha_exact(g, cc.y, k1.w, t, alpha, G);
				const uint32_t cmask = __ballot_sync(0xffffffffu, contrib);
				if (cmask == 0) continue;

				float gv[17];
#pragma unroll
				for (int q = 0; q < 17; q++) gv[q] = 0.0f;
				const int gid = __float_as_int(d.w);
				if (contrib) {
					const uint32_t contributor = base + j;   // 0-based position in the tile list
					const double td = t;
					const float mapped = (float)(fma(td, 100.0, -(100.0 * 0.2)) / ((100.0 - 0.2) * td));
					const float dmax_t_dd = (float)((100.0 * 0.2) / ((100.0 - 0.2) * td * td));
					const float length = (float)sqrt((double)(g.n0 * g.n0 + g.n1 * g.n1 + g.n2 * g.n2) + 1e-7);
					const float nn[3] = { -g.n0 / length, -g.n1 / length, -g.n2 / length };
					const float nraw[3] = { g.n0, g.n1, g.n2 };

					T = T / (1.f - alpha);
					const float weight = alpha * T;
					float dL_dalpha = 0.0f;
					const float col[3] = { d.x, d.y, d.z };
#pragma unroll
					for (int ch = 0; ch < 3; ch++) {
						accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
						last_color[ch] = col[ch];
						dL_dalpha += (col[ch] - accum_rec[ch]) * dL_dpixel[ch];
						gv[10 + ch] = weight * dL_dpixel[ch];
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
					dL_dlength *= 1.f / (length * length);
					float dL_dn[3] = { (-dL_dnn[0] + dL_dlength * nraw[0]) / length,
					                   (-dL_dnn[1] + dL_dlength * nraw[1]) / length,
					                   (-dL_dnn[2] + dL_dlength * nraw[2]) / length };

					float dL_dt = dL_dmax_t;
					if (contributor == max_contributor - 1) dL_dt += dL_dmax_depth;

					dL_dalpha *= T;
					last_alpha = alpha;
					dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;

					const float w = k1.w;
					const float dL_dG = w * dL_dalpha;
					const float2 xy = means2D[gid];
					const float4 con = conic_opacity[gid];
					const float dx = xy.x - (float)px, dy = xy.y - (float)py;
					const float gdx = G * dx, gdy = G * dy;
					const float dG_ddelx = -gdx * con.x - gdy * con.y;
					const float dG_ddely = -gdy * con.z - gdx * con.y;
					const float gmx = dL_dG * dG_ddelx * ddelx_dx;
					const float gmy = dL_dG * dG_ddely * ddely_dy;
					gv[14] = gmx;
					gv[15] = gmy;
					gv[16] = fabsf(gmx) + fabsf(gmy);
					gv[13] = G * dL_dalpha;

					const float dL_dmin_value = dL_dG * G * -0.5f;
					const double AA = g.AA, BB = g.BB;
					double dL_dA = dL_dmin_value * (BB / AA) * (BB / AA) / 4.f;
					double dL_dB = dL_dmin_value * -BB / (2 * AA);
					const double dL_dC = dL_dmin_value * 1.0f;
					dL_dA += dL_dt * BB / (2 * AA * AA);
					dL_dB += dL_dt * -1.f / (2 * AA);
					dL_dn[0] += dL_dA * rx;
					dL_dn[1] += dL_dA * ry;
					dL_dn[2] += dL_dA;

					gv[0] = dL_dn[0] * rx;
					gv[1] = dL_dn[0] * ry + dL_dn[1] * rx;
					gv[2] = dL_dn[0] + dL_dn[2] * rx;
					gv[3] = dL_dn[1] * ry;
					gv[4] = dL_dn[1] + dL_dn[2] * ry;
					gv[5] = dL_dn[2];
					gv[6] = dL_dB * 2 * rx;
					gv[7] = dL_dB * 2 * ry;
					gv[8] = dL_dB * 2;
					gv[9] = dL_dC;
				}

				float* dst = gacc + (size_t)gid * GACC_FLOATS;
				if (__popc(cmask) <= 2) {
					if (contrib) {
						red_add_v4(dst + 0, gv[0], gv[1], gv[2], gv[3]);
						red_add_v4(dst + 4, gv[4], gv[5], gv[6], gv[7]);
						red_add_v4(dst + 8, gv[8], gv[9], gv[10], gv[11]);
						red_add_v4(dst + 12, gv[12], gv[13], gv[14], gv[15]);
						atomicAdd(dst + 16, gv[16]);
					}
				} else {
#pragma unroll
					for (int q = 0; q < 17; q++) gv[q] = warp_sum(gv[q]);
					if (lane == 0) {
						red_add_v4(dst + 0, gv[0], gv[1], gv[2], gv[3]);
						red_add_v4(dst + 4, gv[4], gv[5], gv[6], gv[7]);
						red_add_v4(dst + 8, gv[8], gv[9], gv[10], gv[11]);
						red_add_v4(dst + 12, gv[12], gv[13], gv[14], gv[15]);
						atomicAdd(dst + 16, gv[16]);
					}
				}
			}
		}
		__syncthreads();   // everyone is done with stage s
		if (tid == 0 && i + STAGES < nchunks) issue(i + STAGES);
	}
}

}  // namespace

int launch_render_bwd(const GofParams& prm, const Frame& f, const GeomState& g, const ImgState& im,
                      const BinState& b, const float* background, const float* dL_dpix, float* gacc, cudaStream_t s)
{
	const dim3 grid(f.grid.x, f.grid.y, f.V);
	render_bwd_kernel<<<grid, TILE_PIX, 0, s>>>(im.ranges, b.slab, f.P, prm.W, prm.H, f.focal_x, f.focal_y, background, 0,
	                                           g.means2D, g.conic_opacity, im.final_T, im.n_contrib, dL_dpix, gacc);
	GOF_CUDA_CHECK(cudaGetLastError());
	return GOF_OK;
}

}  // namespace gof
