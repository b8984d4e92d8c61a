// render_fwd.cu -- per-tile front-to-back GOF alpha compositing (K8).
//
// Replaces renderCUDA<3> forward (RAST/cuda_rasterizer/forward.cu:409-612).  Same outputs:
// out_color[9,H,W] (rgb, view-space normal, median depth, alpha, normalised distortion),
// final_T[4,H,W] = (T, dist1, dist2, distortion_raw), n_contrib[2,H,W] = (last, max contributor).
//
// B200 design (render_fwd_kernel, the throughput kernel of batched launches; one-frame launches take
// render_fwd_split_kernel further down):
//   * One CTA per 16x16 tile (the tile size is part of the binning contract): 8 warps, each owning an 8x4 pixel
//     block.  64 registers per thread, four CTAs = 32 warps per SM.
//   * The tile's sorted Gaussians arrive as a contiguous slab of 80-byte records (binning.cu), streamed into a 4-stage
//     shared-memory ring with TMA bulk copies (cp.async.bulk, completion on a "full" mbarrier per stage).  There is
//     no producer warp: the warp that is the last to let go of a stage requests the chunk that reuses it.  No
//     CTA-wide barrier in the loop: warps drift up to three chunks apart, so a warp that has many contributors in one
//     chunk does not stall the other seven (the reference, and our first version, synchronise every batch).
//   * Per 128-record chunk, two passes.  Pass 1 is the sweep: every pixel evaluates the tile-local CONIC pre-test
//     (conic.cuh: a quadratic in the pixel's tile coordinates, five FMAs, coefficients broadcast from shared memory)
//     for the chunk's records whose block mask reaches this warp's 8x4 block, and keeps the survivors as a 128-bit
//     mask.  Pass 2 is lane-private: each pixel walks ITS OWN survivors in list order (float32 quadric, exact FP64 ray
//     minimum, expf, blend), two per trip, from a queue that spans two chunks so that the lanes of a warp are not
//     synchronised at chunk boundaries.  A warp spends pass-2 iterations equal to its busiest pixel's survivor count
//     (~11% of the list) instead of running the expensive path for every record any pixel touches.
//   * Rounding: alpha, T, rgb, median depth, alpha channel and the contributor counters follow
//     the reference's sm_100a build operation by operation and are bit-identical to it in both
//     modes.  With GOF_FLAG_EXACT_BLEND the depth mapping and the normal normalisation also use
//     its IEEE double divide / double sqrt / float divides (all nine channels bit-identical);
//     without it they use float32 reciprocal arithmetic (normals, distortion within ~1e-6).
#include <cstdlib>
#include "blend_math.cuh"
#include "conic.cuh"

namespace gof {

namespace {

#ifndef GOF_FWD_STAGES
#define GOF_FWD_STAGES 4          // power of two (ring addressing by list position); a warp reads two chunks and sweeps a third
#endif
#ifndef GOF_FWD_MIN_CTAS
#define GOF_FWD_MIN_CTAS 4        // inference kernels: 64 registers (12 bytes of spills) for four CTAs = 32 warps per SM: 386 -> 366 us per
#endif                            // 8-view launch, 512^2 batches 7.80 -> 7.21 ms (round 1, with a ninth producer warp: 56 registers, slower).
                                  // The training instantiations (MASK) keep three: their second queue does not fit four times.
#ifndef GOF_FWD_CHUNK
#define GOF_FWD_CHUNK 128
#endif
#ifndef GOF_FWD_SWEEP_ILP
#define GOF_FWD_SWEEP_ILP 4       // records per trip of the conic sweep (A/B on B200: 2: 392 us, 3: 385, 4: 387, 6: 391 per 8-view launch; one frame: 4 saves 4 us)
#endif
#ifndef GOF_FWD_SPLIT_MAX_TILES
#define GOF_FWD_SPLIT_MAX_TILES 444 // launches of at most this many tiles take render_fwd_split_kernel: as long as the tile kernel's
                                   // CTAs fit in ONE wave (148 SMs x 3) the launch is latency-bound.  Measured per frame, one-frame
                                   // calls, tile / split kernel: 256^2 (256 tiles) 167 / 154 us, 320^2 (400) 168 / 160, 384^2 (576)
                                   // 164 / 182, 448^2 (784) 176 / 209, 512^2 (1024) 198 / 243
#endif
#ifndef GOF_FWD_FOLD_W
#define GOF_FWD_FOLD_W 1          // fast blend: normals / distortion accumulate with w = alpha*T formed once
#endif
constexpr int CHUNK = GOF_FWD_CHUNK;       // records per pipeline stage (10 KB at 128)
constexpr int STAGES = GOF_FWD_STAGES;
constexpr int NW = CHUNK / 32;             // 32-record words per chunk
constexpr int SWEEP_ILP = GOF_FWD_SWEEP_ILP;

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
	if (EXACT || !GOF_FWD_FOLD_W) {
		s.distortion = __fmaf_rn(T, __fmul_rn(alpha, err), s.distortion);
		s.dist1 = __fmaf_rn(T, __fmul_rn(alpha, m), s.dist1);
		s.dist2 = __fmaf_rn(T, __fmul_rn(alpha, m2), s.dist2);
	} else {
		// the relaxed channels (normals, distortion) share one weight w = alpha*T (one rounding instead of two)
		const float w = __fmul_rn(alpha, T);
		s.distortion = __fmaf_rn(w, err, s.distortion);
		s.dist1 = __fmaf_rn(w, m, s.dist1);
		s.dist2 = __fmaf_rn(w, m2, s.dist2);
		s.C[3] = __fmaf_rn(-w, nn0, s.C[3]);
		s.C[4] = __fmaf_rn(-w, nn1, s.C[4]);
		s.C[5] = __fmaf_rn(-w, nn2, s.C[5]);
	}

	s.C[0] = __fmaf_rn(T, __fmul_rn(alpha, d.x), s.C[0]);
	s.C[1] = __fmaf_rn(T, __fmul_rn(alpha, d.y), s.C[1]);
	s.C[2] = __fmaf_rn(T, __fmul_rn(alpha, d.z), s.C[2]);
	if (EXACT || !GOF_FWD_FOLD_W) {
		// view-space normal is -n/|n|
		s.C[3] = __fmaf_rn(-T, __fmul_rn(alpha, nn0), s.C[3]);
		s.C[4] = __fmaf_rn(-T, __fmul_rn(alpha, nn1), s.C[4]);
		s.C[5] = __fmaf_rn(-T, __fmul_rn(alpha, nn2), s.C[5]);
	}
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
constexpr int FWD_THREADS = TILE_PIX;                  // every warp blends; there is no dedicated producer warp

// Survivors of one lane (= pixel) in TWO consecutive 128-record chunks, ca and ca+1: eight 32-record words.  The words
// live in shared memory, one private column per thread (word w of the tile list -> row w % 8, so the chunk that enters
// the window overwrites the rows of the chunk that left it); registers hold only the word being consumed (`cur`, the
// bits not popped yet) and its index in the list (`p`).  Invariant ("normalised"): cur != 0 unless nothing is queued.
constexpr int QUEUE_ROWS = 2 * NW;
template <int ROWS, int COLS>
struct LaneQueueT {
	uint32_t cur, p;
	uint32_t col;                                   // shared-window address of this thread's column (row 0)
	__device__ __forceinline__ uint32_t row(uint32_t w) const { return col + (w % ROWS) * (COLS * 4); }
	__device__ __forceinline__ void store_chunk(int c, const uint32_t (&m)[NW]) const
	{
#pragma unroll
		for (int k = 0; k < NW; k++)
			asm volatile("st.shared.u32 [%0], %1;" ::"r"(row((uint32_t)c * NW + k)), "r"(m[k]) : "memory");
	}
	// skip exhausted words; p_end = words queued so far (warp-uniform)
	__device__ __forceinline__ void normalise(uint32_t p_end)
	{
		while (cur == 0 && p + 1 < p_end) {
			p++;
			asm volatile("ld.shared.u32 %0, [%1];" : "=r"(cur) : "r"(row(p)) : "memory");
		}
	}
	// true if the next survivor lies in a word below `limit` (= it belongs to the window's first chunk)
	__device__ __forceinline__ bool pending_below(uint32_t limit) const { return cur != 0 && p < limit; }
	// pop the next survivor in list order: its position in the tile list; false when nothing is queued
	__device__ __forceinline__ bool pop(uint32_t& j, uint32_t p_end)
	{
		if (cur == 0) return false;
		j = (p << 5) + (uint32_t)__ffs((int)cur) - 1u;
		cur &= cur - 1u;
		normalise(p_end);
		return true;
	}
};
using LaneQueue = LaneQueueT<QUEUE_ROWS, FWD_THREADS>;

// MASK: also record, per pixel, which records of the tile list were blended (GOF_FLAG_SAVE_CONTRIB): one bit per list
// position, kept for the window's two chunks in shared memory ([chunk parity][word][thread], lane-private columns) and
// written out as whole 128-record blocks when the warp lets go of a chunk -- contrib[slot][word][thread], coalesced.
template <bool EXACT, bool SINK, bool MASK>
__global__ void __launch_bounds__(FWD_THREADS, MASK ? 3 : GOF_FWD_MIN_CTAS)
render_fwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order, int T, int tiles_x,
                  const float* __restrict__ slab, int W, int H,
                  float focal_x, float focal_y, const float* __restrict__ bg_colors, int bg_stride,
                  float* __restrict__ final_T_all, uint32_t* __restrict__ n_contrib_all, float* __restrict__ out_color_all,
                  const int32_t* __restrict__ mailbox, const uint8_t* __restrict__ block_mask, float* __restrict__ sink_all, int sink_hwc,
                  uint32_t* __restrict__ contrib)
{
	static_assert(NW == 4, "the survivor queues are written for 128-record chunks");
	extern __shared__ __align__(128) unsigned char smem_raw[];
	uint64_t* s_full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * CHUNK * SLAB_BYTES);
	uint32_t* s_released = reinterpret_cast<uint32_t*>(s_full + STAGES);      // per stage: warps that have let go of it (monotonic)
	uint32_t* s_queue = reinterpret_cast<uint32_t*>(s_full + 2 * STAGES);     // [QUEUE_ROWS][256] survivor words (LaneQueue)
	uint32_t* s_blended = s_queue + QUEUE_ROWS * FWD_THREADS;                 // MASK: [QUEUE_ROWS][256] blended bits, same layout
	float* s_out = reinterpret_cast<float*>(s_blended + (MASK ? QUEUE_ROWS * FWD_THREADS : 0));   // [SINK_CH * 256] tile of the frame sink

	const int tid = threadIdx.x;
	const int warp = tid >> 5, lane = tid & 31;
	pdl_trigger();
	pdl_wait();                    // ranges / slab / block masks come from the binning stages
	// CTAs are launched longest-list-first: blockIdx.x -> (view, tile) through tile_order
	const uint32_t gt = tile_order[blockIdx.x];
	const int view = (int)(gt / (uint32_t)T);
	const int tile = (int)(gt - (uint32_t)view * (uint32_t)T);
	const int tile_y = tile / tiles_x, tile_x = tile - tile_y * tiles_x;
	const size_t N = (size_t)W * H;

	const uint2 range = ranges[gt];
	// sync-free mode: if the binning blob was too small nothing was binned -- no list is walked and every
	// output of the batch is POISONED with NaN, so that a caller who never reads the overflow flag
	// (gof_num_rendered / BatchWorkspace.finish) cannot mistake the batch for background-only frames
	const bool overflow = mailbox[1] != 0;
	const int n = overflow ? 0 : (int)(range.y - range.x);
	const int nchunks = (n + CHUNK - 1) / CHUNK;
	const float* tile_slab = slab + (size_t)range.x * SLAB_FLOATS;

	// The tile's slab is streamed through a STAGES-deep ring by TMA bulk copies.  There is no producer warp (it would
	// hold a ninth warp's registers for one busy lane): the first STAGES chunks are requested here, and the warp that is
	// the LAST to let go of a stage requests the chunk that reuses it (release_stage).
	auto request_chunk = [&](int c) {
		const int s = c % STAGES;
		const uint32_t bytes = (uint32_t)min(CHUNK, n - c * CHUNK) * SLAB_BYTES;
		mbar_arrive_expect_tx(&s_full[s], bytes);
		tma_bulk_g2s(smem_raw + (size_t)s * (CHUNK * SLAB_BYTES), tile_slab + (size_t)c * CHUNK * SLAB_FLOATS, bytes, &s_full[s]);
	};
	if (tid == 0) {
#pragma unroll
		for (int s = 0; s < STAGES; s++) { mbar_init(&s_full[s], 1); s_released[s] = 0; }
		mbar_fence_init();
		for (int c = 0; c < min(STAGES, nchunks); c++) request_chunk(c);
	}
	__syncthreads();
	auto release_stage = [&](int c) {              // this warp will not read chunk c's stage again
		__syncwarp();
		if (lane == 0) {
			__threadfence_block();
			const uint32_t before = atomicAdd(&s_released[c % STAGES], 1u);
			if ((before % CONSUMER_WARPS) == CONSUMER_WARPS - 1 && c + STAGES < nchunks) {
				__threadfence_block();
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the warps' reads of the stage precede the bulk write
				request_chunk(c + STAGES);
			}
		}
	};

	// -------------------- every warp: an 8x4 pixel block ----------------------------------------
	const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);   // tile-local pixel
	const uint32_t px = tile_x * TILE_X + lx;
	const uint32_t py = tile_y * TILE_Y + ly;
	const bool inside = px < (uint32_t)W && py < (uint32_t)H;
	const uint32_t pix_id = W * py + px;
	const float rx = pixel_ray(px, W, focal_x);
	const float ry = pixel_ray(py, H, focal_y);

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
	const uint8_t* tile_bm = block_mask + range.x;

	// ---- pass 1 of chunk c: conic sweep over the chunk's records that can touch this warp's 8x4 block (bit `warp` of
	// the record's block mask, ~1/3 of them); coefficients are warp-broadcast shared-memory reads.  Leaves the lane's
	// survivors among the chunk's 128 records in q.
	auto conic_sweep = [&](int c, uint32_t (&m)[NW]) {
		const int s = c % STAGES;
		const int cnt = min(CHUNK, n - c * CHUNK);
		// this lane's share of the chunk's block masks (records lane, 32+lane, 64+lane, 96+lane): read from
		// L2 before waiting for the TMA stage so that the latency overlaps it
		uint32_t bm[NW];
#pragma unroll
		for (int k = 0; k < NW; k++) bm[k] = 0;
		if (!warp_done) {
			const uint8_t* p = tile_bm + c * CHUNK + lane;
#pragma unroll
			for (int k = 0; k < NW; k++)
				if (32 * k + lane < cnt) bm[k] = __ldg(p + 32 * k);
		}
		mbar_wait(&s_full[s], (uint32_t)((c / STAGES) & 1));
#pragma unroll
		for (int k = 0; k < NW; k++) m[k] = 0;       // survivors among records [32k, 32k+32)
		if (!warp_done) {
			const float fx = (float)lx, fy = (float)ly;
			const uint32_t rec = rec_base + (uint32_t)s * (CHUNK * SLAB_BYTES);   // shared-window address of the stage
#pragma unroll 1
			for (int w = 0; w < NW; w++) {
				const int valid = cnt - 32 * w;
				if (valid <= 0) break;
				uint32_t bmw = bm[0];
#pragma unroll
				for (int k = 1; k < NW; k++) bmw = (w == k) ? bm[k] : bmw;
				uint32_t rel = __ballot_sync(0xffffffffu, (bmw >> warp) & 1u);
				uint32_t bits = 0;
				const uint32_t rw = rec + (uint32_t)w * (32 * SLAB_BYTES);
				while (rel != 0) {                       // warp-uniform; SWEEP_ILP records per trip for ILP
					int j[SWEEP_ILP];
					j[0] = __ffs((int)rel) - 1;
					rel &= rel - 1;
#pragma unroll
					for (int k = 1; k < SWEEP_ILP; k++) {
						j[k] = rel ? __ffs((int)rel) - 1 : j[0];
						rel &= rel - 1;                  // (0 & -1 == 0 when the record does not exist)
					}
					float4 a[SWEEP_ILP];
					float2 b[SWEEP_ILP];
#pragma unroll
					for (int k = 0; k < SWEEP_ILP; k++) {
						a[k] = lds128(rw + j[k] * SLAB_BYTES);
						b[k] = lds64(rw + j[k] * SLAB_BYTES + 16);
					}
#pragma unroll
					for (int k = 0; k < SWEEP_ILP; k++)
						if (!conic_reject(a[k].x, a[k].y, a[k].z, a[k].w, b[k].x, b[k].y, fx, fy)) bits |= 1u << j[k];
				}
#pragma unroll
				for (int k = 0; k < NW; k++)
					if (w == k) m[k] = bits;
			}
			if (done) {
#pragma unroll
				for (int k = 0; k < NW; k++) m[k] = 0;
			}
		}
	};

	// ---- pass 2: each pixel blends its own survivors, in list order (float32 quadric, exact FP64 ray minimum, expf,
	// blend), two per trip so that the two evaluations' dependency chains interleave.
	// The lanes of a warp are NOT synchronised at chunk boundaries.  A lane queues the survivors of TWO chunks: ca (the
	// oldest chunk some lane of the warp still needs) and ca+1.  A pixel with few survivors in chunk ca moves on to
	// those of ca+1 while its heavier neighbours are still in ca; the warp advances (lets go of the stage of ca, sweeps
	// chunk ca+2 into the queue) when no lane has anything left in ca.  Per-pixel survivor counts per chunk fluctuate
	// around a spatially smooth load: in per-chunk lock-step 14 of 32 lanes were active in this loop; the one-chunk
	// lookahead absorbs the fluctuation (oracle model: 19 of 32; the persistent part of the imbalance and saturated
	// pixels cap it at 21).
	LaneQueue q;
	q.cur = 0;
	q.p = 0;
	q.col = smem_u32(s_queue) + (uint32_t)tid * 4u;
	// MASK: this thread's column of blended bits sits QUEUE_ROWS rows behind its queue column
	constexpr uint32_t BLENDED_OFF = QUEUE_ROWS * FWD_THREADS * 4;
	if (MASK) {
#pragma unroll
		for (int w = 0; w < QUEUE_ROWS; w++) s_blended[w * FWD_THREADS + tid] = 0;
	}
	auto mark_blended = [&](uint32_t j) {               // list position j was blended by this pixel
		const uint32_t a = q.row(j >> 5) + BLENDED_OFF;
		uint32_t v;
		asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
		v |= 1u << (j & 31u);
		asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
	};
	// contributor-mask block of chunk c of this tile (gof_common.cuh: BinState::contrib)
	uint32_t* const tile_contrib = MASK ? contrib + ((size_t)(range.x >> 7) + gt) * CONTRIB_SLOT_WORDS + tid : nullptr;
	auto flush_blended = [&](int c) {                   // the warp is done with chunk c: write its four words, clear the rows
#pragma unroll
		for (int k = 0; k < NW; k++) {
			const uint32_t a = q.row((uint32_t)c * NW + k) + BLENDED_OFF;
			uint32_t v;
			asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
			tile_contrib[(size_t)c * CONTRIB_SLOT_WORDS + k * TILE_PIX] = v;
			asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(0u) : "memory");
		}
	};
	{
		uint32_t m[NW];
		if (nchunks > 0) { conic_sweep(0, m); q.store_chunk(0, m); q.cur = m[0]; }
		if (nchunks > 1) { conic_sweep(1, m); q.store_chunk(1, m); }
		q.normalise((uint32_t)min(nchunks, 2) * NW);
	}
	// the ring holds STAGES * CHUNK consecutive list positions: record j sits at rec_base + 80 (j mod 512)
	static_assert((STAGES & (STAGES - 1)) == 0, "ring addressing by list position needs a power-of-two ring");
	constexpr uint32_t RING_MASK = STAGES * CHUNK - 1;
	for (int ca = 0; ca < nchunks; ca++) {
		const uint32_t p_end = (uint32_t)min(nchunks, ca + 2) * NW;
		const uint32_t p_b = (uint32_t)(ca + 1) * NW;              // first word of chunk ca+1
		while (__any_sync(0xffffffffu, q.pending_below(p_b))) {
			uint32_t ja, jb;
			if (!q.pop(ja, p_end)) continue;                       // nothing queued: wait for the warp to advance
			const bool hb = q.pop(jb, p_end);
			if (!hb) jb = ja;
			const uint32_t ra = rec_base + (ja & RING_MASK) * SLAB_BYTES;
			const uint32_t rb = rec_base + (jb & RING_MASK) * SLAB_BYTES;
			const float4 a1 = lds128(ra + 16), a2 = lds128(ra + 32), a3 = lds128(ra + 48), a4 = lds128(ra + 64);
			const float4 b1 = lds128(rb + 16), b2 = lds128(rb + 32), b3 = lds128(rb + 48), b4 = lds128(rb + 64);
			const PairGeom ga = pair_geom(a1, a2, a3, rx, ry);
			const PairGeom gb = pair_geom(b1, b2, b3, rx, ry);
			float ta, alpha_a, tb, alpha_b;
			const bool oka = pair_alpha_eval(ga, a4.x, a1.z, ta, alpha_a);
			const bool okb = pair_alpha_eval(gb, b4.x, b1.z, tb, alpha_b) && hb;
			// contributor ids are 1-based list positions (forward.cu:494-496)
			bool saturated = false;
			if (oka) {
				saturated = blend_pair<EXACT>(st, ga, ta, alpha_a, make_float4(a4.y, a4.z, a4.w, 0.0f), ja + 1u);
				if (MASK && !saturated) mark_blended(ja);
			}
			if (okb && !saturated) {
				saturated = blend_pair<EXACT>(st, gb, tb, alpha_b, make_float4(b4.y, b4.z, b4.w, 0.0f), jb + 1u);
				if (MASK && !saturated) mark_blended(jb);
			}
			if (saturated) {
				done = true;                                        // the pixel is saturated: it has no survivors any more
				q.cur = 0;
				q.p = 0xffffff00u;                                  // beyond every p_end: normalise() never reloads
			}
		}
		if (MASK) flush_blended(ca);
		release_stage(ca);
		warp_done = __all_sync(0xffffffffu, done);
		if (ca + 2 < nchunks) {
			uint32_t m[NW];
			conic_sweep(ca + 2, m);
			q.store_chunk(ca + 2, m);
			q.normalise((uint32_t)(ca + 3) * NW);                  // a lane that had run dry picks up the new words
		}
	}

	if (overflow) {
		const float poison = __int_as_float(0x7fc00000);
#pragma unroll
		for (int k = 0; k < 8; k++) st.C[k] = poison;
		st.distortion = poison;
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
	if constexpr (SINK) {
		// Frame sink = rgb, median depth, alpha (what the render loops read back, visualize.py:304-306), usually
		// pinned host memory mapped into the device address space: the stores are posted PCIe writes that drain
		// while the other tiles still blend.  PCIe efficiency is set by the length of the contiguous runs, so the
		// tile is transposed through shared memory and written as whole tile rows with 16-byte stores: 64-byte
		// runs for [V,5,H,W], 320-byte runs for the channels-last layout [V,H,W,5].
		float* sink = sink_all + (size_t)view * SINK_CH * N;
		const bool whole = (W % 4 == 0) && (tile_x * TILE_X + TILE_X <= W) && (tile_y * TILE_Y + TILE_Y <= H);   // CTA-uniform
		float v[SINK_CH];
		{
			const float* bg_color = bg_colors + (size_t)view * bg_stride;
#pragma unroll
			for (int ch = 0; ch < 3; ch++) v[ch] = __fmaf_rn(st.T, bg_color[ch], st.C[ch]);
			v[3] = st.C[CH_DEPTH];
			v[4] = st.C[CH_ALPHA];
		}
		if (!whole) {
			if (inside) {
#pragma unroll
				for (int ch = 0; ch < SINK_CH; ch++) sink[sink_hwc ? (size_t)pix_id * SINK_CH + ch : (size_t)ch * N + pix_id] = v[ch];
			}
			return;
		}
		const int lp = ly * TILE_X + lx;
#pragma unroll
		for (int ch = 0; ch < SINK_CH; ch++) s_out[sink_hwc ? lp * SINK_CH + ch : ch * TILE_PIX + lp] = v[ch];
		asm volatile("bar.sync 1, %0;" :: "n"(TILE_PIX) : "memory");    // the 8 consumer warps (the producer has left)
		const size_t row0 = (size_t)(tile_y * TILE_Y) * W + (size_t)tile_x * TILE_X;   // first pixel of the tile
		for (int i = tid; i < SINK_CH * TILE_PIX / 4; i += TILE_PIX) {
			const float4 q = reinterpret_cast<const float4*>(s_out)[i];
			size_t at;
			if (sink_hwc) {
				const int row = i / (TILE_X * SINK_CH / 4), c4 = i - row * (TILE_X * SINK_CH / 4);
				at = (row0 + (size_t)row * W) * SINK_CH + (size_t)c4 * 4;
			} else {
				const int ch = i / (TILE_PIX / 4), r = i - ch * (TILE_PIX / 4), row = r / (TILE_X / 4), c4 = r - row * (TILE_X / 4);
				at = (size_t)ch * N + row0 + (size_t)row * W + (size_t)c4 * 4;
			}
			*reinterpret_cast<float4*>(sink + at) = q;
		}
	}
}

// ------------------------------------------------------------------------------------------------------------------
// Small launches (the one-frame-per-call API): render_fwd_split_kernel.
//
// A launch of one 256^2 frame is 256 tiles on 148 SMs, and its time is the time of the longest tile's warps: every warp
// walks the whole list alone -- conic sweep of chunk c+2, then the lane-private blend of chunks c, c+1 -- at ~0.2
// instructions per cycle, while three quarters of the SM's issue slots idle (ncu, one frame: issue active 49 % on the
// SMs that are busy, SMs busy 56 % of the launch).  The throughput kernel above has nothing to offer here; this variant
// spends the idle issue slots on thread-level parallelism INSIDE a tile without touching the arithmetic:
//   * a CTA owns HALF a tile (16x8 pixels = four 8x4 blocks), so a frame is 512 CTAs and the longest tile is spread over
//     two SMs; each half streams the tile's slab through its own TMA ring (the second read comes from L2);
//   * per 8x4 block TWO warps: a SWEEPER that runs the conic pre-test up to four chunks ahead and leaves the
//     per-pixel survivor words in the shared-memory queue, and a BLENDER that only walks survivors (pass 2 of the
//     kernel above, same code, same order, same roundings).  The sweep (~1/3 of the longest warp's path) leaves the
//     critical path; the two meet through two monotonic counters in shared memory (chunks swept / chunks consumed):
//     writer = data stores, __syncwarp, fence, volatile counter store; reader = volatile counter load, fence, data
//     loads (compute-sanitizer's racecheck only models barriers and reports these hand-offs as hazards; memcheck is clean).
// Results are bit-identical to render_fwd_kernel (tests/test_gpu_batch.py compares one-frame calls with batched ones).
constexpr int SPLIT_BLENDERS = 4;                       // 8x4 blocks per half tile, one blender warp each
#ifndef GOF_SPLIT_SWEEPERS
#define GOF_SPLIT_SWEEPERS 2
#endif
#ifndef GOF_SPLIT_MIN_CTAS
#define GOF_SPLIT_MIN_CTAS 4
#endif
constexpr int SPLIT_SWEEPERS = GOF_SPLIT_SWEEPERS;      // sweeper warps, each serving two of the blocks (or one, with four sweepers)
constexpr int SPLIT_THREADS = (SPLIT_BLENDERS + SPLIT_SWEEPERS) * 32;   // 192: four CTAs per SM, so that the 512 CTAs of one
                                                        // 256^2 frame are resident at once (592 slots)
constexpr int SPLIT_PIX = SPLIT_BLENDERS * 32;          // 128 pixels per CTA
constexpr int SPLIT_AHEAD = 4;                          // chunks of survivor words held per block (window of 2 + 2 swept ahead)
constexpr int SPLIT_QROWS = SPLIT_AHEAD * NW;
static_assert(SPLIT_AHEAD <= STAGES, "a swept chunk's records must still be in the ring when the blender reaches it");
constexpr int SPLIT_BLOCKS_PER_SWEEPER = SPLIT_BLENDERS / SPLIT_SWEEPERS;
static_assert(SPLIT_BLOCKS_PER_SWEEPER == 1 || SPLIT_BLOCKS_PER_SWEEPER == 2, "a sweeper serves one or two blocks");
using SplitQueue = LaneQueueT<SPLIT_QROWS, SPLIT_PIX>;

__device__ __forceinline__ uint32_t lds_volatile_u32(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
	return v;
}
__device__ __forceinline__ void sts_volatile_u32(uint32_t* p, uint32_t v)
{
	asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

template <bool EXACT>
__global__ void __launch_bounds__(SPLIT_THREADS, GOF_SPLIT_MIN_CTAS)
render_fwd_split_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order, int T, int tiles_x,
                        const float* __restrict__ slab, int W, int H,
                        float focal_x, float focal_y, const float* __restrict__ bg_colors, int bg_stride,
                        float* __restrict__ final_T_all, uint32_t* __restrict__ n_contrib_all, float* __restrict__ out_color_all,
                        const int32_t* __restrict__ mailbox, const uint8_t* __restrict__ block_mask)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	uint64_t* s_full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * CHUNK * SLAB_BYTES);
	uint32_t* s_released = reinterpret_cast<uint32_t*>(s_full + STAGES);      // per stage: blenders that have let go of it (monotonic)
	uint32_t* s_swept = s_released + STAGES;                                   // per block: chunks whose survivor words are queued
	uint32_t* s_consumed = s_swept + SPLIT_BLENDERS;                           // per block: chunks the blender has left behind
	uint32_t* s_wdone = s_consumed + SPLIT_BLENDERS;                           // per block: every pixel saturated, nothing left to sweep
	uint32_t* s_queue = s_wdone + SPLIT_BLENDERS;                              // [SPLIT_QROWS][128] survivor words

	const int tid = threadIdx.x;
	const int warp = tid >> 5, lane = tid & 31;
	const bool sweeper = warp >= SPLIT_BLENDERS;
	int bw = warp & (SPLIT_BLENDERS - 1);                  // blender: its block; sweeper: set per sweep
	pdl_trigger();
	pdl_wait();
	const uint32_t gt = tile_order[blockIdx.x >> 1];       // the two halves of a tile are neighbours in launch order
	const int half = (int)(blockIdx.x & 1u);
	const int view = (int)(gt / (uint32_t)T);
	const int tile = (int)(gt - (uint32_t)view * (uint32_t)T);
	const int tile_y = tile / tiles_x, tile_x = tile - tile_y * tiles_x;
	const size_t N = (size_t)W * H;

	const uint2 range = ranges[gt];
	const bool overflow = mailbox[1] != 0;                 // see render_fwd_kernel
	const int n = overflow ? 0 : (int)(range.y - range.x);
	const int nchunks = (n + CHUNK - 1) / CHUNK;
	const float* tile_slab = slab + (size_t)range.x * SLAB_FLOATS;

	auto request_chunk = [&](int c) {
		const int s = c % STAGES;
		const uint32_t bytes = (uint32_t)min(CHUNK, n - c * CHUNK) * SLAB_BYTES;
		mbar_arrive_expect_tx(&s_full[s], bytes);
		tma_bulk_g2s(smem_raw + (size_t)s * (CHUNK * SLAB_BYTES), tile_slab + (size_t)c * CHUNK * SLAB_FLOATS, bytes, &s_full[s]);
	};
	if (tid == 0) {
#pragma unroll
		for (int s = 0; s < STAGES; s++) { mbar_init(&s_full[s], 1); s_released[s] = 0; }
#pragma unroll
		for (int k = 0; k < SPLIT_BLENDERS; k++) { s_swept[k] = 0; s_consumed[k] = 0; s_wdone[k] = 0; }
		mbar_fence_init();
		for (int c = 0; c < min(STAGES, nchunks); c++) request_chunk(c);
	}
	__syncthreads();

	const uint32_t rec_base = smem_u32(smem_raw);
	SplitQueue q;
	q.cur = 0;
	q.p = 0;

	if (sweeper) {
		// ---- conic sweep (pass 1 of render_fwd_kernel) for the 32 pixels of a block, up to SPLIT_AHEAD chunks in front
		// of the block's blender.  A sweeper warp serves two blocks and always sweeps the one that is further behind.
		const uint8_t* tile_bm = block_mask + range.x;
		const int b0 = SPLIT_BLOCKS_PER_SWEEPER * (warp - SPLIT_BLENDERS);
		int nx0 = 0, nx1 = SPLIT_BLOCKS_PER_SWEEPER == 2 ? 0 : nchunks;                              // next chunk to sweep for blocks b0, b0 + 1
		// block k of this sweeper: done -> nchunks; else its next chunk if the queue rows are free, -1 if not yet
		auto eligible = [&](int k, int& nx) -> int {
			if (nx >= nchunks) return -1;
			if (lds_volatile_u32(&s_wdone[b0 + k])) {      // every pixel of the block is saturated: nothing left to sweep
				nx = nchunks;
				if (lane == 0) sts_volatile_u32(&s_swept[b0 + k], (uint32_t)nchunks);   // (the blender only lets go of the stages)
				return -1;
			}
			// the rows of chunk c reuse those of chunk c - AHEAD: the blender must have left that one behind
			return ((int)lds_volatile_u32(&s_consumed[b0 + k]) + SPLIT_AHEAD > nx) ? nx : -1;
		};
		while (nx0 < nchunks || nx1 < nchunks) {
			const int e0 = eligible(0, nx0), e1 = eligible(1, nx1);
			if (e0 < 0 && e1 < 0) { __nanosleep(20); continue; }
			const int pick = (e0 >= 0 && (e1 < 0 || e0 <= e1)) ? 0 : 1;
			__threadfence_block();                         // acquire side of the blender's `consumed` release: its reads of the rows come first
			bw = b0 + pick;
			const int c = pick ? e1 : e0;
			const int blk = half * SPLIT_BLENDERS + bw;        // 8x4 block of the tile = bit of the records' block masks
			const float fx = (float)((blk & 1) * 8 + (lane & 7)), fy = (float)((blk >> 1) * 4 + (lane >> 3));
			q.col = smem_u32(s_queue) + (uint32_t)(bw * 32 + lane) * 4u;
			const int s = c % STAGES;
			const int cnt = min(CHUNK, n - c * CHUNK);
			uint32_t bm[NW];
			{
				const uint8_t* p = tile_bm + c * CHUNK + lane;
#pragma unroll
				for (int k = 0; k < NW; k++) bm[k] = (32 * k + lane < cnt) ? (uint32_t)__ldg(p + 32 * k) : 0u;
			}
			mbar_wait(&s_full[s], (uint32_t)((c / STAGES) & 1));
			uint32_t m[NW];
			const uint32_t rec = rec_base + (uint32_t)s * (CHUNK * SLAB_BYTES);
#pragma unroll
			for (int w = 0; w < NW; w++) {
				uint32_t rel = __ballot_sync(0xffffffffu, (bm[w] >> blk) & 1u);
				uint32_t bits = 0;
				const uint32_t rw = rec + (uint32_t)w * (32 * SLAB_BYTES);
				while (rel != 0) {
					int j[SWEEP_ILP];
					j[0] = __ffs((int)rel) - 1;
					rel &= rel - 1;
#pragma unroll
					for (int k = 1; k < SWEEP_ILP; k++) {
						j[k] = rel ? __ffs((int)rel) - 1 : j[0];
						rel &= rel - 1;
					}
					float4 a[SWEEP_ILP];
					float2 b[SWEEP_ILP];
#pragma unroll
					for (int k = 0; k < SWEEP_ILP; k++) {
						a[k] = lds128(rw + j[k] * SLAB_BYTES);
						b[k] = lds64(rw + j[k] * SLAB_BYTES + 16);
					}
#pragma unroll
					for (int k = 0; k < SWEEP_ILP; k++)
						if (!conic_reject(a[k].x, a[k].y, a[k].z, a[k].w, b[k].x, b[k].y, fx, fy)) bits |= 1u << j[k];
				}
				m[w] = bits;
			}
			q.store_chunk(c, m);
			__syncwarp();
			if (lane == 0) {
				__threadfence_block();
				sts_volatile_u32(&s_swept[bw], (uint32_t)(c + 1));
			}
			if (pick) nx1 = c + 1; else nx0 = c + 1;
		}
		return;
	}

	const int blk = half * SPLIT_BLENDERS + bw;            // 8x4 block of the tile
	const int lx = (blk & 1) * 8 + (lane & 7), ly = (blk >> 1) * 4 + (lane >> 3);
	q.col = smem_u32(s_queue) + (uint32_t)(bw * 32 + lane) * 4u;

	// -------------------- blender: pass 2 of render_fwd_kernel over the queued survivor words --------------------
	const uint32_t px = tile_x * TILE_X + lx;
	const uint32_t py = tile_y * TILE_Y + ly;
	const bool inside = px < (uint32_t)W && py < (uint32_t)H;
	const uint32_t pix_id = W * py + px;
	const float rx = pixel_ray(px, W, focal_x);
	const float ry = pixel_ray(py, H, focal_y);

	PixState st;
	st.T = 1.0f;
#pragma unroll
	for (int k = 0; k < 8; k++) st.C[k] = 0.0f;
	st.dist1 = st.dist2 = st.distortion = 0.0f;
	st.last_contributor = 0;
	st.max_contributor = 0xFFFFFFFFu;
	bool done = !inside;
	if (__all_sync(0xffffffffu, done) && lane == 0) sts_volatile_u32(&s_wdone[bw], 1u);

	auto release_stage = [&](int c) {
		__syncwarp();
		if (lane == 0) {
			__threadfence_block();
			const uint32_t before = atomicAdd(&s_released[c % STAGES], 1u);
			if ((before % SPLIT_BLENDERS) == SPLIT_BLENDERS - 1 && c + STAGES < nchunks) {
				__threadfence_block();
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
				request_chunk(c + STAGES);
			}
			sts_volatile_u32(&s_consumed[bw], (uint32_t)(c + 1));
		}
	};
	// chunks [0, k) are swept and their records have landed (the blender observes the stage's barrier itself, so the
	// TMA writes are ordered before its reads without relying on the sweeper's observation)
	auto wait_chunk = [&](int c) {
		while ((int)lds_volatile_u32(&s_swept[bw]) <= c) __nanosleep(20);
		__threadfence_block();
		mbar_wait(&s_full[c % STAGES], (uint32_t)((c / STAGES) & 1));
	};
	if (nchunks > 0) {
		wait_chunk(0);
		if (nchunks > 1) wait_chunk(1);
		if (done) {                                    // pixel outside the image: never reads the queue
			q.p = 0xffffff00u;
		} else {
			asm volatile("ld.shared.u32 %0, [%1];" : "=r"(q.cur) : "r"(q.row(0)) : "memory");
			q.normalise((uint32_t)min(nchunks, 2) * NW);
		}
	}
	constexpr uint32_t RING_MASK = STAGES * CHUNK - 1;
	for (int ca = 0; ca < nchunks; ca++) {
		const uint32_t p_end = (uint32_t)min(nchunks, ca + 2) * NW;
		const uint32_t p_b = (uint32_t)(ca + 1) * NW;
		while (__any_sync(0xffffffffu, q.pending_below(p_b))) {
			uint32_t ja, jb;
			if (!q.pop(ja, p_end)) continue;
			const bool hb = q.pop(jb, p_end);
			if (!hb) jb = ja;
			const uint32_t ra = rec_base + (ja & RING_MASK) * SLAB_BYTES;
			const uint32_t rb = rec_base + (jb & RING_MASK) * SLAB_BYTES;
			const float4 a1 = lds128(ra + 16), a2 = lds128(ra + 32), a3 = lds128(ra + 48), a4 = lds128(ra + 64);
			const float4 b1 = lds128(rb + 16), b2 = lds128(rb + 32), b3 = lds128(rb + 48), b4 = lds128(rb + 64);
			const PairGeom ga = pair_geom(a1, a2, a3, rx, ry);
			const PairGeom gb = pair_geom(b1, b2, b3, rx, ry);
			float ta, alpha_a, tb, alpha_b;
			const bool oka = pair_alpha_eval(ga, a4.x, a1.z, ta, alpha_a);
			const bool okb = pair_alpha_eval(gb, b4.x, b1.z, tb, alpha_b) && hb;
			bool saturated = false;
			if (oka) saturated = blend_pair<EXACT>(st, ga, ta, alpha_a, make_float4(a4.y, a4.z, a4.w, 0.0f), ja + 1u);
			if (okb && !saturated) saturated = blend_pair<EXACT>(st, gb, tb, alpha_b, make_float4(b4.y, b4.z, b4.w, 0.0f), jb + 1u);
			if (saturated) {
				done = true;
				q.cur = 0;
				q.p = 0xffffff00u;
			}
		}
		release_stage(ca);
		if (__all_sync(0xffffffffu, done)) {
			if (lane == 0) sts_volatile_u32(&s_wdone[bw], 1u);
		}
		if (ca + 2 < nchunks) {
			wait_chunk(ca + 2);
			q.normalise((uint32_t)(ca + 3) * NW);
		}
	}

	if (overflow) {
		const float poison = __int_as_float(0x7fc00000);
#pragma unroll
		for (int k = 0; k < 8; k++) st.C[k] = poison;
		st.distortion = poison;
	}
	if (inside) {
		float* final_T = final_T_all + (size_t)view * 4 * N;
		uint32_t* n_contrib = n_contrib_all + (size_t)view * 2 * N;
		float* out_color = out_color_all + (size_t)view * OUT_CH * N;
		const float* bg_color = bg_colors + (size_t)view * bg_stride;
		const float T_ = st.T;
		const float om = __fsub_rn(1.0f, T_);
		const float dnorm = (float)((double)st.distortion / ((double)__fmul_rn(om, om) + 1e-7));
		final_T[pix_id] = T_;
		final_T[pix_id + N] = st.dist1;
		final_T[pix_id + 2 * N] = st.dist2;
		final_T[pix_id + 3 * N] = st.distortion;
		n_contrib[pix_id] = st.last_contributor;
		n_contrib[pix_id + N] = st.max_contributor;
#pragma unroll
		for (int ch = 0; ch < 3; ch++) out_color[ch * N + pix_id] = __fmaf_rn(T_, bg_color[ch], st.C[ch]);
#pragma unroll
		for (int ch = 3; ch < 8; ch++) out_color[ch * N + pix_id] = st.C[ch];
		out_color[CH_DIST * N + pix_id] = dnorm;
	}
}

}  // namespace

int launch_render_fwd(const GofParams& prm, const Frame& f, const GeomState& g, const ImgState& im, const BinState& b,
                      const float* background, int bg_stride, float* out_color, float* sink, int sink_hwc, cudaStream_t s)
{
	const dim3 grid((unsigned)(f.T * f.V), 1, 1);
	const bool mask = (prm.flags & GOF_FLAG_SAVE_CONTRIB) != 0;
	const size_t smem = (size_t)STAGES * CHUNK * SLAB_BYTES + 2 * STAGES * sizeof(uint64_t) + (size_t)QUEUE_ROWS * FWD_THREADS * 4 * (mask ? 2 : 1) +
	                    (sink ? (size_t)SINK_CH * TILE_PIX * sizeof(float) : 0);   // ring | full barriers | release counters | queues (| blended bits) | sink tile
	auto launch = [&](auto kernel) {
		// per device and per function; cheap enough to set on every launch (one process may drive several GPUs)
		if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return;
		launch_chained(PDL_BLEND, kernel, grid, dim3(FWD_THREADS), smem, s, im.ranges, im.tile_order, f.T, (int)f.grid.x, b.slab, prm.W, prm.H,
		               f.focal_x, f.focal_y, background, bg_stride, im.final_T, im.n_contrib, out_color, g.mailbox, b.block_mask, sink,
		               sink_hwc, b.contrib);
	};
	const bool exact = (prm.flags & GOF_FLAG_EXACT_BLEND) != 0;
	// small launches (the one-frame-per-call API): the latency variant, two CTAs per tile (render_fwd_split_kernel)
	const char* split_env = getenv("GOF_FWD_SPLIT_MAX_TILES");       // read per launch: the tests switch between the two kernels
	const int split_max = split_env ? atoi(split_env) : GOF_FWD_SPLIT_MAX_TILES;
	if (!mask && !sink && (int)grid.x <= split_max) {
		const size_t ssm = (size_t)STAGES * CHUNK * SLAB_BYTES + STAGES * sizeof(uint64_t) + (STAGES + 3 * SPLIT_BLENDERS) * sizeof(uint32_t) +
		                   (size_t)SPLIT_QROWS * SPLIT_PIX * 4;
		auto launch_split = [&](auto kernel) {
			if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm) != cudaSuccess) return;
			launch_chained(PDL_BLEND, kernel, dim3(grid.x * 2), dim3(SPLIT_THREADS), ssm, s, im.ranges, im.tile_order, f.T, (int)f.grid.x, b.slab,
			               prm.W, prm.H, f.focal_x, f.focal_y, background, bg_stride, im.final_T, im.n_contrib, out_color, g.mailbox, b.block_mask);
		};
		if (exact) launch_split(render_fwd_split_kernel<true>); else launch_split(render_fwd_split_kernel<false>);
		GOF_CUDA_CHECK(cudaGetLastError());
		return GOF_OK;
	}
	if (mask) {
		if (sink) { if (exact) launch(render_fwd_kernel<true, true, true>); else launch(render_fwd_kernel<false, true, true>); }
		else      { if (exact) launch(render_fwd_kernel<true, false, true>); else launch(render_fwd_kernel<false, false, true>); }
	} else {
		if (sink) { if (exact) launch(render_fwd_kernel<true, true, false>); else launch(render_fwd_kernel<false, true, false>); }
		else      { if (exact) launch(render_fwd_kernel<true, false, false>); else launch(render_fwd_kernel<false, false, false>); }
	}
	GOF_CUDA_CHECK(cudaGetLastError());
	return GOF_OK;
}

}  // namespace gof
