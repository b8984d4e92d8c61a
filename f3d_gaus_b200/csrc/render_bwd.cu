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
constexpr int NW = CHUNK / 32;
constexpr int STAGES = 4;            // power of two: the ring is addressed by list position (chunk c lives in stage c % 4)
constexpr int QUEUE_ROWS = 2 * NW;   // survivor words of the two chunks a warp works on (see LaneQueueDown)
constexpr int CONSUMER_WARPS = TILE_PIX / 32;
constexpr int BWD_THREADS = TILE_PIX + 32;
constexpr int STAGE_SLAB_BYTES = CHUNK * SLAB_BYTES;     // 10 KB of slab records ...
constexpr int STAGE_REC_BYTES = STAGE_SLAB_BYTES + CHUNK * BWD_REC_BYTES;   // ... + 4 KB of backward records per stage
constexpr size_t BWD_SMEM = (size_t)STAGES * STAGE_REC_BYTES + 2 * STAGES * sizeof(uint64_t) + 64 +
                            (size_t)QUEUE_ROWS * TILE_PIX * 4;      // ring | barriers | s_max | survivor queues

// Survivors of one lane (= pixel) in the two chunks its warp is working on, walked from the BACK of the tile list: eight
// 32-record words in shared memory, one private column per thread (word w of the list -> row w % 8), the word being
// consumed (`cur`, bits not popped yet) and its index `p` in registers.  Mirror image of render_fwd.cu's LaneQueue.
// Invariant: cur != 0 unless nothing is queued at or above p_low.
struct LaneQueueDown {
	uint32_t cur, p, col;
	__device__ __forceinline__ uint32_t row(uint32_t w) const { return col + (w % QUEUE_ROWS) * (TILE_PIX * 4); }
	__device__ __forceinline__ void store_chunk(int c, uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) const
	{
		const uint32_t w = (uint32_t)c * NW;
		asm volatile("st.shared.u32 [%0], %1;" ::"r"(row(w + 0)), "r"(m0) : "memory");
		asm volatile("st.shared.u32 [%0], %1;" ::"r"(row(w + 1)), "r"(m1) : "memory");
		asm volatile("st.shared.u32 [%0], %1;" ::"r"(row(w + 2)), "r"(m2) : "memory");
		asm volatile("st.shared.u32 [%0], %1;" ::"r"(row(w + 3)), "r"(m3) : "memory");
	}
	__device__ __forceinline__ void normalise(uint32_t p_low)       // skip exhausted words, down to word p_low
	{
		while (cur == 0 && p > p_low) {
			p--;
			asm volatile("ld.shared.u32 %0, [%1];" : "=r"(cur) : "r"(row(p)) : "memory");
		}
	}
	__device__ __forceinline__ bool pop(uint32_t& j, uint32_t p_low) // deepest queued survivor: its position in the tile list
	{
		if (cur == 0) return false;
		const uint32_t bit = 31u - (uint32_t)__clz((int)cur);
		j = (p << 5) + bit;
		cur &= ~(1u << bit);
		normalise(p_low);
		return true;
	}
};

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
                  const float* __restrict__ dL_dpixels_all, float* __restrict__ gacc_all,
                  const uint32_t* __restrict__ contrib, const int32_t* __restrict__ mailbox, const float* __restrict__ bwd_rec)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	// layout: [STAGES] record stages | full[STAGES] | empty[STAGES] | s_max[8]
	const uint32_t rec_base = smem_u32(smem_raw);
	uint64_t* s_full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * STAGE_REC_BYTES);
	uint64_t* s_empty = s_full + STAGES;
	uint32_t* s_max = reinterpret_cast<uint32_t*>(s_empty + STAGES);
	uint32_t* s_queue = s_max + 16;                                  // [QUEUE_ROWS][256]

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
	// Did the forward run with GOF_FLAG_SAVE_CONTRIB (recorded in the mailbox by the tile scan)?  Then it left (a) per-pixel
	// contributor masks: pass 1 is a 16-byte load per pixel and chunk instead of the conic sweep, and every queued record
	// contributes; (b) tile-ordered backward records {mean2D, conic, id}: pass 2 reads them from the ring instead of
	// chasing point_list -> means2D / conic_opacity through L2 (two dependent round trips per pair, the top stall).
	const bool have_masks = mailbox[3] != 0;

	const uint32_t last_contributor = inside ? n_contrib[pix_id] : 0;
	const uint32_t max_contributor = inside ? n_contrib[pix_id + N] : 0;
	const uint32_t* tile_contrib = contrib + ((size_t)(range.x >> 7) + gt) * CONTRIB_SLOT_WORDS + (tid & (TILE_PIX - 1));

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
	// the i-th chunk in walk order is chunk c = nchunks-1-i; it lives in stage c % STAGES (every run of STAGES consecutive
	// chunks uses every stage once, so a stage's k-th use is walk index i with i / STAGES == k: the barrier parities)

	if (!consumer) {
		// ------------- producer warp: one elected lane streams the slab, deepest chunk first ---------
		if (lane == 0) {
			for (int i = 0; i < nchunks; i++) {
				const int c = nchunks - 1 - i;
				const int s = c % STAGES;
				if (i >= STAGES) mbar_wait_backoff(&s_empty[s], (uint32_t)(((i / STAGES) - 1) & 1));
				const int cnt = min(CHUNK, m - c * CHUNK);
				const uint32_t bytes = (uint32_t)cnt * SLAB_BYTES, aux_bytes = have_masks ? (uint32_t)cnt * BWD_REC_BYTES : 0u;
				mbar_arrive_expect_tx(&s_full[s], bytes + aux_bytes);
				tma_bulk_g2s(smem_raw + (size_t)s * STAGE_REC_BYTES, tile_slab + (size_t)c * CHUNK * SLAB_FLOATS, bytes, &s_full[s]);
				if (have_masks)      // the training forward also left tile-ordered backward records: stream them alongside
					tma_bulk_g2s(smem_raw + (size_t)s * STAGE_REC_BYTES + STAGE_SLAB_BYTES,
					             bwd_rec + ((size_t)range.x + (size_t)c * CHUNK) * BWD_REC_FLOATS, aux_bytes, &s_full[s]);
			}
		}
		return;
	}

	// ------------- consumer warps -----------------------------------------------------------------
	const float rx = pixel_ray(px, W, focal_x);
	const float ry = pixel_ray(py, H, focal_y);

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

	// ---- pass 1 of chunk c: the lane's contributor candidates among the chunk's 128 records, into the queue rows.
	// With masks: four words loaded from the forward's contributor masks (requested one chunk ahead, see request_masks);
	// without: the conic sweep over the records this pixel can have blended (index < last_contributor).
	uint32_t nx0 = 0, nx1 = 0, nx2 = 0, nx3 = 0;
	auto request_masks = [&](int c) {
		// volatile: the loads must be ISSUED here, a whole chunk ahead of their use (the compiler would otherwise be
		// free to sink them to their use and expose the L2 latency once per chunk)
		const uint32_t* w = tile_contrib + (size_t)c * CONTRIB_SLOT_WORDS;
		asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(nx0) : "l"(w));
		asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(nx1) : "l"(w + TILE_PIX));
		asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(nx2) : "l"(w + 2 * TILE_PIX));
		asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(nx3) : "l"(w + 3 * TILE_PIX));
	};
	LaneQueueDown q;
	q.cur = 0;
	q.p = (uint32_t)max(nchunks, 1) * NW - 1u;
	q.col = smem_u32(s_queue) + (uint32_t)tid * 4u;
	auto stage_of = [&](int c) { return rec_base + (uint32_t)(c % STAGES) * STAGE_REC_BYTES; };
	auto sweep = [&](int c) {                 // chunk c must be the next one in walk order that has not been swept
		const int i = nchunks - 1 - c;
		uint32_t m0 = nx0, m1 = nx1, m2 = nx2, m3 = nx3;
		if (have_masks && c > 0) request_masks(c - 1);
		mbar_wait(&s_full[c % STAGES], (uint32_t)((i / STAGES) & 1));
		const uint32_t base = (uint32_t)c * CHUNK;
		if (!have_masks) {
			m0 = m1 = m2 = m3 = 0;
			if (base < last_contributor) {
				const int cnt = min(CHUNK, m - c * CHUNK);
				const float fx = (float)lx, fy = (float)ly;
				const uint32_t rec = stage_of(c);
#pragma unroll 1
				for (int w = 0; w < NW; w++) {
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
		}
		q.store_chunk(c, m0, m1, m2, m3);
		return m3;
	};
	if (have_masks && nchunks > 0) request_masks(nchunks - 1);
	if (nchunks > 0) q.cur = sweep(nchunks - 1);
	if (nchunks > 1) sweep(nchunks - 2);
	q.normalise((uint32_t)max(nchunks - 2, 0) * NW);

	// ---- pass 2: each pixel walks its own contributors, back to front.  As in the forward blend the lanes of a warp are
	// not synchronised at chunk boundaries: a lane queues the candidates of two chunks, ca (the deepest chunk some lane of
	// the warp still needs) and ca-1, and moves on to ca-1 as soon as it has nothing left in ca; the warp advances (lets
	// go of the stage of ca, sweeps chunk ca-2) when no lane has anything left in ca.
	for (int ca = nchunks - 1; ca >= 0; ca--) {
		const uint32_t p_low = (uint32_t)max(ca - 1, 0) * NW;       // lowest word queued
		const uint32_t p_a = (uint32_t)ca * NW;                     // first word of chunk ca
		while (__any_sync(0xffffffffu, q.cur != 0 && q.p >= p_a)) {
			uint32_t contributor;                                   // 0-based position in the tile list
			if (q.pop(contributor, p_low)) {
				const uint32_t j = contributor & (CHUNK - 1);
				const uint32_t rec = stage_of((int)(contributor >> 7));
				const uint32_t r = rec + j * SLAB_BYTES;
				const float4 k1 = lds128(r + 16), k2 = lds128(r + 32), k3 = lds128(r + 48), k4 = lds128(r + 64);
				// the Gaussian's id and its 2-D mean / conic: from the tile-ordered backward record in the ring when the
				// forward left one; otherwise gathered (two dependent L2 round trips, requested before the exact evaluation
				// so that they are in flight during its double division)
				int gid;
				float2 xy;
				float4 con;
				if (have_masks) {
					const uint32_t ar = rec + STAGE_SLAB_BYTES + j * BWD_REC_BYTES;
					const float4 x0 = lds128(ar);
					const float2 x1 = lds64(ar + 16);
					xy = make_float2(x0.x, x0.y);
					con = make_float4(x0.z, x0.w, x1.x, 0.0f);
					gid = __float_as_int(x1.y);
				} else {
					gid = (int)__ldg(&tile_ids[contributor]);
					xy = __ldg(&means2D[gid]);
					con = __ldg(&conic_opacity[gid]);
				}
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
		if (lane == 0) mbar_arrive(&s_empty[ca % STAGES]);   // this warp is finished with the stage of chunk ca
		if (ca >= 2) {
			sweep(ca - 2);
			q.normalise((uint32_t)(ca - 2) * NW);              // a lane that had run dry picks up the new words
		}
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
	                                                     g.means2D, g.conic_opacity, im.final_T, im.n_contrib, dL_dpix, gacc, b.contrib, g.mailbox, b.bwd_rec);
	GOF_CUDA_CHECK(cudaGetLastError());
	return GOF_OK;
}

}  // namespace gof
