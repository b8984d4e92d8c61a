// integrate.cu -- GOF point integration (the mesh-extraction query): K15-K18.
//
// Replaces preprocessPointsCUDA / createWithKeys / the second scan+sort+ranges / integrateCUDA of
// the reference (RAST/cuda_rasterizer/forward.cu:722-766,803-1218; rasterizer_impl.cu:113-144,
// 530-792).  For a set of 3-D query points (the tetrahedra vertices of GOF mesh extraction) and one
// camera it returns, per point, the alpha accumulated along the ray through the point up to the
// point's depth, plus the pixel colour; and per pixel an rgb / max-depth / alpha image.
//
// Reference algorithm, per pixel (integrateCUDA):
//   phase 1  walk the tile's sorted Gaussians with FIVE rays (pixel centre + the four corners,
//            +-0.5 px); each ray keeps its own transmittance; a Gaussian that contributes to any ray
//            is appended to the pixel's `contributed_ids` (a 1024-entry uint16 array in LOCAL
//            memory, 2 KB per thread + 8 KB stack in its sm_100a build); colour/alpha from the centre
//            ray, depth = max t over all rays;
//   phase 2  every pixel scans ALL query points of its tile for those inside its pixel square (256 at
//            a time), then re-walks its contributed Gaussians once per batch, accumulating
//            alpha along the exact ray through each point with t clamped to the point's depth.
// All arithmetic is float32 except the double product in min_value (forward.cu:918).
//
// B200 design:
//   * phase 1 (integrate_pixels_kernel) has the forward blend's skeleton: TMA-streamed slab,
//     8 consumer warps + producer, conic sweep (evaluated at the five ray positions; the slab's conic
//     coefficients are built with the +-0.5 px ray box, conic.cuh) and a lane-private exact pass.
//     The contributed set is a BIT MASK over the tile's list, written to HBM once per chunk
//     ([tile][word][pixel], 4 B per 32 records per pixel) instead of a per-thread local array;
//   * phase 2 (integrate_points_kernel) is POINT-parallel: query points are bucketed by tile (count /
//     scan / scatter -- their order inside a tile does not influence any output, so the reference's
//     sort by depth is not needed), one thread per point looks up its pixel's bit mask and walks only
//     those records, again streaming the tile's slab through the TMA ring.  The reference does
//     256 x (points in tile) point-in-pixel tests per tile; here each point finds its pixel directly.
// Differences kept out on purpose: the reference's re-scan quirk (a block that needs a second
// 256-point batch re-counts the tile's last point; it only affects the `number of projected points`
// debug channel 8) and its 256-points-per-batch limit do not exist here.
#include "blend_math.cuh"
#include "conic.cuh"

namespace gof {

template <typename T>
static void take2(char*& p, T*& ptr, size_t count)
{
	p = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(p)));
	ptr = reinterpret_cast<T*>(p);
	p += count * sizeof(T);
}

IntegrateScratch IntegrateScratch::carve(char* base, size_t PN, size_t T, size_t max_count)
{
	IntegrateScratch s;
	char* p = base;
	s.words = (int)(4 * ((max_count + 127) / 128));
	take2(p, s.used_mask, T * (size_t)s.words * TILE_PIX);
	take2(p, s.points2D, PN);
	take2(p, s.point_depths, PN);
	take2(p, s.point_tile, PN);
	take2(p, s.point_counts, T);
	take2(p, s.point_cursor, T);
	take2(p, s.point_ranges, T);
	take2(p, s.point_list, PN);
	take2(p, s.point_mailbox, MAILBOX_HEAD + 1);
	s.total = align_up((size_t)(p - base)) + ALIGN;
	return s;
}

namespace {

constexpr int CHUNK = 128;
constexpr int STAGES = 3;
constexpr int CONSUMER_WARPS = TILE_PIX / 32;
constexpr int INT_THREADS = TILE_PIX + 32;
constexpr int STAGE_BYTES = CHUNK * SLAB_BYTES;
constexpr size_t INT_SMEM = (size_t)STAGES * STAGE_BYTES + 2 * STAGES * sizeof(uint64_t) + 16;
constexpr int MAX_CONTRIBUTED = 1024;   // MAX_NUM_CONTRIBUTORS * 4 (auxiliary.h:26, forward.cu:879)

// Quadric coefficients of one (ray, Gaussian) pair (forward.cu:906-913) in the roundings of the reference's
// sm_100a build.  Its compiler shares products between the five rays of a pixel (rays 1/3 share rx, 1/2 and
// 3/4 share ry), so the fused/unfused pattern differs per ray; read off its SASS with a symbolic executor:
//   ray 0 (and the per-point ray of phase 2):
//       n0 = v2 + fma(rx,v0, ry*v1)   n1 = v4 + fma(rx,v1, ry*v3)   n2 = v5 + fma(ry,v4, rx*v2)   bb = v8 + fma(rx,v6, ry*v7)
//   rays 1,3 (x - 0.5):  n0 = v2 + (rx*v0 + ry*v1)   n1 = v4 + fma(rx,v1, ry*v3)   n2 = v5 + fma(rx,v2, ry*v4)   bb = v8 + (rx*v6 + ry*v7)
//   rays 2,4 (x + 0.5):  n0 = v2 + (rx*v0 + ry*v1)   n1 = v4 + (rx*v1 + ry*v3)     n2 = v5 + fma(rx,v2, ry*v4)   bb = v8 + (rx*v6 + ry*v7)
//   all:  AA = n2 + fma(rx,n0, ry*n1),  BB = bb + bb.
// These float32 values feed the same ill-conditioned difference as in the blend: one ulp moves
// alpha by percents at F3D-Gaus scales, so they are pinned with explicit round-to-nearest intrinsics.
struct RayQuad { float AA, BB, CC; };
template <int VARIANT>   // 0: ray 0 / points, 1: rays 1 and 3, 2: rays 2 and 4
__device__ __forceinline__ RayQuad ray_quadric(const float4& k1, const float4& k2, const float4& k3, float C, float rx, float ry)
{
	const float v0 = k1.w, v1 = k2.x, v2 = k2.y, v3 = k2.z, v4 = k2.w, v5 = k3.x, v6 = k3.y, v7 = k3.z, v8 = k3.w;
	float n0, n1, n2, bb;
	if (VARIANT == 0) {
		n0 = __fadd_rn(v2, __fmaf_rn(rx, v0, __fmul_rn(ry, v1)));
		n1 = __fadd_rn(v4, __fmaf_rn(rx, v1, __fmul_rn(ry, v3)));
		n2 = __fadd_rn(v5, __fmaf_rn(ry, v4, __fmul_rn(rx, v2)));
		bb = __fadd_rn(v8, __fmaf_rn(rx, v6, __fmul_rn(ry, v7)));
	} else {
		n0 = __fadd_rn(v2, __fadd_rn(__fmul_rn(rx, v0), __fmul_rn(ry, v1)));
		n1 = (VARIANT == 1) ? __fadd_rn(v4, __fmaf_rn(rx, v1, __fmul_rn(ry, v3)))
		                    : __fadd_rn(v4, __fadd_rn(__fmul_rn(rx, v1), __fmul_rn(ry, v3)));
		n2 = __fadd_rn(v5, __fmaf_rn(rx, v2, __fmul_rn(ry, v4)));
		bb = __fadd_rn(v8, __fadd_rn(__fmul_rn(rx, v6), __fmul_rn(ry, v7)));
	}
	RayQuad q;
	q.AA = __fadd_rn(n2, __fmaf_rn(rx, n0, __fmul_rn(ry, n1)));
	q.BB = __fadd_rn(bb, bb);
	q.CC = C;
	return q;
}

// One ray of phase 1 (forward.cu:915-962).  Returns true if the Gaussian contributed through this ray.
template <int K>
__device__ __forceinline__ bool integrate_ray(const float4& k1, const float4& k2, const float4& k3, const float4& k4,
                                              float rx, float ry, float& T, float& C0, float& C1, float& C2,
                                              float& Cdepth, float& Calpha)
{
	const RayQuad q = ray_quadric<(K == 0) ? 0 : ((K == 1 || K == 3) ? 1 : 2)>(k1, k2, k3, k4.x, rx, ry);
	const float t = __fdiv_rn(-q.BB, __fadd_rn(q.AA, q.AA));
	if (t <= 0.2) return false;
	const float u = __fdiv_rn(-q.BB, q.AA);
	const double min_value = fma((double)u, (double)q.BB * 0.25, (double)q.CC);
	float power = (float)(min_value * -0.5);
	if (power > 0.0f) power = 0.0f;
	const float alpha = min(0.99f, __fmul_rn(k1.z, expf(power)));
	if (alpha < 1.0f / 255.0f) return false;
	const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
	if (test_T < 0.0001f) return false;
	if (K == 0) {
		C0 = __fmaf_rn(T, __fmul_rn(alpha, k4.y), C0);
		C1 = __fmaf_rn(T, __fmul_rn(alpha, k4.z), C1);
		C2 = __fmaf_rn(T, __fmul_rn(alpha, k4.w), C2);
	}
	if (t > Cdepth) Cdepth = t;
	if (K == 0) Calpha = __fmaf_rn(T, alpha, Calpha);
	T = test_T;
	return true;
}

// the producer side shared by both kernels: stream chunks [0, nchunks) of a tile's slab, in order
__device__ __forceinline__ void stream_slab(const float* tile_slab, int n, int nchunks, unsigned char* smem_raw,
                                            uint64_t* s_full, uint64_t* s_empty)
{
	for (int c = 0; c < nchunks; c++) {
		const int s = c % STAGES;
		if (c >= STAGES) mbar_wait_backoff(&s_empty[s], (uint32_t)(((c / STAGES) - 1) & 1));
		const int cnt = min(CHUNK, n - c * CHUNK);
		const uint32_t bytes = (uint32_t)cnt * SLAB_BYTES;
		mbar_arrive_expect_tx(&s_full[s], bytes);
		tma_bulk_g2s(smem_raw + (size_t)s * STAGE_BYTES, tile_slab + (size_t)c * CHUNK * SLAB_FLOATS, bytes, &s_full[s]);
	}
}

// ---------------------------------------------------------------------------------- phase 1 --
__global__ void __launch_bounds__(INT_THREADS)
integrate_pixels_kernel(const uint2* __restrict__ ranges, const float* __restrict__ slab, int W, int H,
                        float focal_x, float focal_y, const float* __restrict__ bg_color,
                        float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
                        uint32_t* __restrict__ used_mask, int words)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const uint32_t rec_base = smem_u32(smem_raw);
	uint64_t* s_full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * STAGE_BYTES);
	uint64_t* s_empty = s_full + STAGES;

	const int tid = threadIdx.x;
	const int warp = tid >> 5, lane = tid & 31;
	const int tile = blockIdx.y * gridDim.x + blockIdx.x;
	const uint2 range = ranges[tile];
	const int n = (int)(range.y - range.x);
	const int nchunks = (n + CHUNK - 1) / CHUNK;
	const float* tile_slab = slab + (size_t)range.x * SLAB_FLOATS;

	if (tid == 0) {
#pragma unroll
		for (int s = 0; s < STAGES; s++) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], CONSUMER_WARPS); }
		mbar_fence_init();
	}
	__syncthreads();
	if (warp == CONSUMER_WARPS) {
		if (lane == 0) stream_slab(tile_slab, n, nchunks, smem_raw, s_full, s_empty);
		return;
	}

	const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);
	const int lpix = ly * TILE_X + lx;                  // pixel index inside the tile
	const uint32_t px = blockIdx.x * TILE_X + lx, py = blockIdx.y * TILE_Y + ly;
	const bool inside = px < (uint32_t)W && py < (uint32_t)H;
	const uint32_t pix_id = W * py + px;
	const size_t N = (size_t)W * H;
	const float pfx = (float)px + 0.5f, pfy = (float)py + 0.5f;
	// the five rays: centre, then the corners (-,-) (+,-) (-,+) (+,+)   (forward.cu:883-884)
	const float offx[5] = { 0.0f, -0.5f, 0.5f, -0.5f, 0.5f }, offy[5] = { 0.0f, -0.5f, -0.5f, 0.5f, 0.5f };
	float rxs[5], rys[5];
#pragma unroll
	for (int k = 0; k < 5; k++) {
		rxs[k] = (pfx + offx[k] - W / 2.) / focal_x;
		rys[k] = (pfy + offy[k] - H / 2.) / focal_y;
	}
	const float fx = (float)lx, fy = (float)ly;

	float Ts[5] = { 1.0f, 1.0f, 1.0f, 1.0f, 1.0f };
	float C0 = 0.f, C1 = 0.f, C2 = 0.f, Cdepth = 0.f, Calpha = 0.f;
	uint32_t last_contributor = 0, n_local = 0;
	bool done = !inside;
	uint32_t* my_mask = used_mask + ((size_t)tile * words) * TILE_PIX + lpix;   // + word * TILE_PIX

	for (int c = 0; c < nchunks; c++) {
		const int s = c % STAGES;
		mbar_wait(&s_full[s], (uint32_t)((c / STAGES) & 1));
		const int cnt = min(CHUNK, n - c * CHUNK);
		const uint32_t rec = rec_base + (uint32_t)s * STAGE_BYTES;
		const uint32_t base = (uint32_t)c * CHUNK;
		uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
		if (!done) {
#pragma unroll 1
			for (int w = 0; w < CHUNK / 32; w++) {
				const int valid = cnt - 32 * w;
				if (valid <= 0) break;
				uint32_t bits = 0;
				const uint32_t rw = rec + (uint32_t)w * (32 * SLAB_BYTES);
#pragma unroll 8
				for (int jj = 0; jj < 32; jj++) {
					const float4 k0 = lds128(rw + jj * SLAB_BYTES);
					const float2 k1 = lds64(rw + jj * SLAB_BYTES + 16);
					// a pair can contribute through any of the five rays
					bool rej = conic_reject(k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, fx, fy);
					rej = rej && conic_reject(k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, fx - 0.5f, fy - 0.5f);
					rej = rej && conic_reject(k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, fx + 0.5f, fy - 0.5f);
					rej = rej && conic_reject(k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, fx - 0.5f, fy + 0.5f);
					rej = rej && conic_reject(k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, fx + 0.5f, fy + 0.5f);
					if (!rej) bits |= 1u << jj;
				}
				if (valid < 32) bits &= (1u << valid) - 1u;
				if (w == 0) m0 = bits; else if (w == 1) m1 = bits; else if (w == 2) m2 = bits; else m3 = bits;
			}
		}
		uint32_t u0 = 0, u1 = 0, u2 = 0, u3 = 0;   // records that contributed through at least one ray
		uint32_t jbase = 0;
		while ((m0 | m1 | m2 | m3) != 0) {
			if (m0 == 0) { m0 = m1; m1 = m2; m2 = m3; m3 = 0; jbase += 32; }
			if (m0 != 0) {
				const uint32_t bit = (uint32_t)__ffs((int)m0) - 1u;
				m0 &= m0 - 1u;
				const uint32_t j = jbase + bit;
				const uint32_t r = rec + j * SLAB_BYTES;
				const float4 k1 = lds128(r + 16), k2 = lds128(r + 32), k3 = lds128(r + 48), k4 = lds128(r + 64);
				bool used = integrate_ray<0>(k1, k2, k3, k4, rxs[0], rys[0], Ts[0], C0, C1, C2, Cdepth, Calpha);
				used |= integrate_ray<1>(k1, k2, k3, k4, rxs[1], rys[1], Ts[1], C0, C1, C2, Cdepth, Calpha);
				used |= integrate_ray<2>(k1, k2, k3, k4, rxs[2], rys[2], Ts[2], C0, C1, C2, Cdepth, Calpha);
				used |= integrate_ray<3>(k1, k2, k3, k4, rxs[3], rys[3], Ts[3], C0, C1, C2, Cdepth, Calpha);
				used |= integrate_ray<4>(k1, k2, k3, k4, rxs[4], rys[4], Ts[4], C0, C1, C2, Cdepth, Calpha);
				if (used) {
					last_contributor = base + j + 1;
					const uint32_t ub = 1u << (j & 31);
					if (j < 32) u0 |= ub; else if (j < 64) u1 |= ub; else if (j < 96) u2 |= ub; else u3 |= ub;
					n_local++;
					if (n_local >= MAX_CONTRIBUTED) { done = true; m0 = 0; m1 = 0; m2 = 0; m3 = 0; }   // forward.cu:986-990
				}
			}
		}
		if (inside) {
			uint32_t* dst = my_mask + (size_t)(4 * c) * TILE_PIX;
			dst[0] = u0; dst[TILE_PIX] = u1; dst[2 * TILE_PIX] = u2; dst[3 * TILE_PIX] = u3;
		}
		__syncwarp();
		if (lane == 0) mbar_arrive(&s_empty[s]);
	}

	if (inside) {
		final_T[pix_id] = Ts[0];
		n_contrib[pix_id] = last_contributor;
		out_color[0 * N + pix_id] = C0 + Ts[0] * bg_color[0];
		out_color[1 * N + pix_id] = C1 + Ts[0] * bg_color[1];
		out_color[2 * N + pix_id] = C2 + Ts[0] * bg_color[2];
		out_color[3 * N + pix_id] = 0.0f;
		out_color[4 * N + pix_id] = 0.0f;
		out_color[5 * N + pix_id] = 0.0f;
		out_color[CH_DEPTH * N + pix_id] = Cdepth;
		out_color[CH_ALPHA * N + pix_id] = Calpha;
		out_color[CH_DIST * N + pix_id] = 0.0f;      // number of projected points: counted by phase 2
	}
}

// ------------------------------------------------------------------------ point binning ------
__global__ void preprocess_points_kernel(int PN, const float* __restrict__ points3D, const float* __restrict__ vm,
                                         const float* __restrict__ pm, int W, int H, float focal_x, float focal_y,
                                         dim3 grid, float2* __restrict__ points2D, float* __restrict__ depths,
                                         uint32_t* __restrict__ point_tile, uint32_t* __restrict__ counts,
                                         float* __restrict__ out_alpha, float* __restrict__ out_rgb)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= PN) return;
	// defaults of the reference's glue for points that are never reached (rasterize_points.cu:268-269)
	out_alpha[idx] = 1.0f;
	out_rgb[3 * (size_t)idx] = 0.0f; out_rgb[3 * (size_t)idx + 1] = 0.0f; out_rgb[3 * (size_t)idx + 2] = 0.0f;
	point_tile[idx] = 0xffffffffu;
	const float x = points3D[3 * (size_t)idx], y = points3D[3 * (size_t)idx + 1], z = points3D[3 * (size_t)idx + 2];
	const float3 p_view = { vm[0] * x + vm[4] * y + vm[8] * z + vm[12], vm[1] * x + vm[5] * y + vm[9] * z + vm[13],
	                        vm[2] * x + vm[6] * y + vm[10] * z + vm[14] };
	if (p_view.z <= 0.2f) return;                                           // auxiliary.h:190-201
	const float2 pi = { (float)(focal_x * p_view.x / (p_view.z + 0.0000001f) + W / 2.), (float)(focal_y * p_view.y / (p_view.z + 0.0000001f) + H / 2.) };
	if (pi.x < 0 || pi.x >= W || pi.y < 0 || pi.y >= H) return;             // forward.cu:758-760
	depths[idx] = p_view.z;
	points2D[idx] = pi;
	const int tx = min((int)grid.x - 1, max(0, (int)(pi.x / TILE_X)));       // rasterizer_impl.cu:134-135
	const int ty = min((int)grid.y - 1, max(0, (int)(pi.y / TILE_Y)));
	const uint32_t t = ty * grid.x + tx;
	point_tile[idx] = t;
	atomicAdd(&counts[t], 1u);
}

__global__ void scatter_points_kernel(int PN, const uint32_t* __restrict__ point_tile, uint32_t* __restrict__ cursor,
                                      uint32_t* __restrict__ point_list)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= PN) return;
	const uint32_t t = point_tile[idx];
	if (t == 0xffffffffu) return;
	point_list[atomicAdd(&cursor[t], 1u)] = (uint32_t)idx;
}

// ---------------------------------------------------------------------------------- phase 2 --
// One CTA per tile; its points are processed 256 at a time, one thread per point.
__global__ void __launch_bounds__(INT_THREADS)
integrate_points_kernel(const uint2* __restrict__ ranges, const float* __restrict__ slab, const uint2* __restrict__ point_ranges,
                        const uint32_t* __restrict__ point_list, const float2* __restrict__ points2D,
                        const float* __restrict__ point_depths, int W, int H, float focal_x, float focal_y,
                        const uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
                        const uint32_t* __restrict__ used_mask, int words, float* __restrict__ out_alpha,
                        float* __restrict__ out_rgb)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const uint32_t rec_base = smem_u32(smem_raw);
	uint64_t* s_full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * STAGE_BYTES);
	uint64_t* s_empty = s_full + STAGES;

	const int tid = threadIdx.x;
	const int warp = tid >> 5, lane = tid & 31;
	const int tile = blockIdx.y * gridDim.x + blockIdx.x;
	const uint2 range = ranges[tile];
	const int n = (int)(range.y - range.x);
	const int nchunks = (n + CHUNK - 1) / CHUNK;
	const float* tile_slab = slab + (size_t)range.x * SLAB_FLOATS;
	const uint2 prange = point_ranges[tile];
	const int npts = (int)(prange.y - prange.x);
	if (npts <= 0) return;
	const int ngroups = (npts + TILE_PIX - 1) / TILE_PIX;
	const size_t N = (size_t)W * H;

	if (tid == 0) {
#pragma unroll
		for (int s = 0; s < STAGES; s++) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], CONSUMER_WARPS); }
		mbar_fence_init();
	}
	__syncthreads();
	// The slab is streamed once per group of 256 points; chunk numbering continues across groups so that
	// the ring's phase bookkeeping stays monotonic.
	if (warp == CONSUMER_WARPS) {
		if (lane == 0) {
			for (int gidx = 0; gidx < ngroups; gidx++)
				for (int c = 0; c < nchunks; c++) {
					const int cc = gidx * nchunks + c;
					const int s = cc % STAGES;
					if (cc >= STAGES) mbar_wait_backoff(&s_empty[s], (uint32_t)(((cc / STAGES) - 1) & 1));
					const int cnt = min(CHUNK, n - c * CHUNK);
					const uint32_t bytes = (uint32_t)cnt * SLAB_BYTES;
					mbar_arrive_expect_tx(&s_full[s], bytes);
					tma_bulk_g2s(smem_raw + (size_t)s * STAGE_BYTES, tile_slab + (size_t)c * CHUNK * SLAB_FLOATS, bytes, &s_full[s]);
				}
		}
		return;
	}

	for (int gidx = 0; gidx < ngroups; gidx++) {
		const int pi = gidx * TILE_PIX + tid;
		const bool have = pi < npts;
		uint32_t pid = 0, last = 0;
		float rx = 0.f, ry = 0.f, ray_depth = 0.f;
		const uint32_t* my_mask = used_mask;
		if (have) {
			pid = point_list[prange.x + pi];
			const float2 xy = points2D[pid];
			ray_depth = point_depths[pid];
			// the pixel whose square [px, px+1) x [py, py+1) holds the point (forward.cu:1070-1071)
			const int px = min(W - 1, (int)floorf(xy.x)), py = min(H - 1, (int)floorf(xy.y));
			const int lpix = (py - blockIdx.y * TILE_Y) * TILE_X + (px - blockIdx.x * TILE_X);
			const size_t pix_id = (size_t)py * W + px;
			rx = (xy.x - W / 2.) / focal_x;
			ry = (xy.y - H / 2.) / focal_y;
			last = n_contrib[pix_id];
			my_mask = used_mask + ((size_t)tile * words) * TILE_PIX + lpix;
			out_rgb[3 * (size_t)pid + 0] = out_color[0 * N + pix_id];
			out_rgb[3 * (size_t)pid + 1] = out_color[1 * N + pix_id];
			out_rgb[3 * (size_t)pid + 2] = out_color[2 * N + pix_id];
			atomicAdd(&out_color[CH_DIST * N + pix_id], 1.0f);
		}
		float T = 1.0f, acc = 0.0f;
		for (int c = 0; c < nchunks; c++) {
			const int cc = gidx * nchunks + c;
			const int s = cc % STAGES;
			mbar_wait(&s_full[s], (uint32_t)((cc / STAGES) & 1));
			const uint32_t rec = rec_base + (uint32_t)s * STAGE_BYTES;
			const uint32_t base = (uint32_t)c * CHUNK;
			uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
			if (have && base < last) {
				const uint32_t* src = my_mask + (size_t)(4 * c) * TILE_PIX;
				m0 = src[0]; m1 = src[TILE_PIX]; m2 = src[2 * TILE_PIX]; m3 = src[3 * TILE_PIX];
			}
			uint32_t jbase = 0;
			while ((m0 | m1 | m2 | m3) != 0) {
				if (m0 == 0) { m0 = m1; m1 = m2; m2 = m3; m3 = 0; jbase += 32; }
				if (m0 != 0) {
					const uint32_t bit = (uint32_t)__ffs((int)m0) - 1u;
					m0 &= m0 - 1u;
					const uint32_t r = rec + (jbase + bit) * SLAB_BYTES;
					const float4 k1 = lds128(r + 16), k2 = lds128(r + 32), k3 = lds128(r + 48);
					const float C = lds32(r + 64);
					const RayQuad q = ray_quadric<0>(k1, k2, k3, C, rx, ry);
					float t = __fdiv_rn(-q.BB, __fadd_rn(q.AA, q.AA));
					if (t > ray_depth) t = ray_depth;
					// power = -0.5 (AA t t + BB t + CC) as the reference's build rounds it (forward.cu:1171)
					const float power = __fmul_rn(__fadd_rn(q.CC, __fmaf_rn(q.BB, t, __fmul_rn(t, __fmul_rn(q.AA, t)))), -0.5f);
					const float alpha = min(0.99f, __fmul_rn(k1.z, expf(power)));
					if (alpha < 1.0f / 255.0f) continue;
					acc = __fmaf_rn(alpha, T, acc);
					T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
				}
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(&s_empty[s]);
		}
		if (have) out_alpha[pid] = acc;
	}
}

}  // namespace

int launch_integrate(const GofParams& prm, const GofInputs& in, const Frame& f, const GeomState& g, const ImgState& im,
                     const BinState& b, const IntegrateScratch& sc, int PN, const float* points3D, float* out_color,
                     float* out_alpha_integrated, float* out_color_integrated, cudaStream_t s)
{
	const dim3 grid(f.grid.x, f.grid.y, 1);
	GOF_CUDA_CHECK(cudaFuncSetAttribute(integrate_pixels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)INT_SMEM));
	GOF_CUDA_CHECK(cudaFuncSetAttribute(integrate_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)INT_SMEM));
	integrate_pixels_kernel<<<grid, INT_THREADS, INT_SMEM, s>>>(im.ranges, b.slab, prm.W, prm.H, f.focal_x, f.focal_y,
	                                                           in.background, im.final_T, im.n_contrib, out_color,
	                                                           sc.used_mask, sc.words);
	GOF_CUDA_CHECK(cudaGetLastError());
	if (PN <= 0) return GOF_OK;
	GOF_CUDA_CHECK(cudaMemsetAsync(sc.point_counts, 0, (size_t)f.T * sizeof(uint32_t), s));
	preprocess_points_kernel<<<(PN + 255) / 256, 256, 0, s>>>(PN, points3D, in.viewmatrix, in.projmatrix, prm.W, prm.H,
	                                                         f.focal_x, f.focal_y, f.grid, sc.points2D, sc.point_depths,
	                                                         sc.point_tile, sc.point_counts, out_alpha_integrated,
	                                                         out_color_integrated);
	GOF_CUDA_CHECK(cudaGetLastError());
	int rc;
	if ((rc = launch_tile_scan_raw(f.T, sc.point_counts, sc.point_ranges, sc.point_cursor, sc.point_mailbox, s)) != GOF_OK) return rc;
	scatter_points_kernel<<<(PN + 255) / 256, 256, 0, s>>>(PN, sc.point_tile, sc.point_cursor, sc.point_list);
	GOF_CUDA_CHECK(cudaGetLastError());
	integrate_points_kernel<<<grid, INT_THREADS, INT_SMEM, s>>>(im.ranges, b.slab, sc.point_ranges, sc.point_list, sc.points2D,
	                                                           sc.point_depths, prm.W, prm.H, f.focal_x, f.focal_y, im.n_contrib,
	                                                           out_color, sc.used_mask, sc.words, out_alpha_integrated,
	                                                           out_color_integrated);
	GOF_CUDA_CHECK(cudaGetLastError());
	return GOF_OK;
}

}  // namespace gof
