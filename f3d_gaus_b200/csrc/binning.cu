// binning.cu -- tile binning for a batch of V views: tile-count scan (K2/K6/K7), bucket scatter
// (K4) and per-tile sort fused with the slab gather (K5 + the tile-ordered record stream).
//
// Replaces cub::DeviceScan::InclusiveSum + duplicateWithKeys + cub::DeviceRadixSort::SortPairs +
// identifyTileRanges of the reference (RAST/cuda_rasterizer/rasterizer_impl.cu:70-111,149-171,
// 332-373).  The reference's contract is the ORDER of each tile's list: the 64-bit key
// (tile << 32 | float_bits(depth)) sorted ascending by a stable sort, ties keeping the emission
// order, which is ascending Gaussian index.  That order is reproduced exactly, but not by a
// global radix sort over 41..43 key bits:
//   1. the preprocess already counted, per tile, how many Gaussians touch it (one atomic per
//      duplicate); tile_scan_kernel turns the counts into ranges[tile] = [first, last+1) -- the
//      result identifyTileRanges extracts from the sorted keys -- plus the batch's num_rendered;
//   2. scatter_kernel drops (float_bits(depth) << 32 | index) of every duplicate into its tile's
//      bucket (slot claimed with an atomic; arrival order is irrelevant because)
//   3. tile_sort_gather_kernel sorts each bucket by that 64-bit value: depth first, index second.
//      All values of a bucket are distinct, so the result is THE stable order of the reference.
//      A bucket is ~700 entries at F3D-Gaus sizes: one CTA sorts it in shared memory with a
//      bitonic network (buckets above the shared-memory capacity are sorted in place in global
//      memory by the same network).  The same CTA then writes point_list and the tile-ordered
//      80-byte slab records, including the tile-local conic pre-test coefficients (conic.cuh).
// HBM/L2 traffic per duplicate: 8 B written + 8 B read for the bucket entry, 64 B gathered,
// 84 B written -- versus ~13 passes over 12-byte pairs for the 6-pass LSD sort it replaces, and
// three launches instead of twelve.
#include "gof_common.cuh"
#include "conic.cuh"

namespace gof {

template <typename T>
static void take(char*& p, T*& ptr, size_t count)
{
	p = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(p)));
	ptr = reinterpret_cast<T*>(p);
	p += count * sizeof(T);
}

GeomState GeomState::carve(char* base, size_t P, size_t V)
{
	GeomState g;
	char* p = base;
	const size_t n = P * V;
	take(p, g.depths, n);
	take(p, g.means2D, n);
	take(p, g.conic_opacity, n);
	take(p, g.rec, n * REC_FLOATS);
	take(p, g.tiles_touched, n);
	take(p, g.rect, n);
	take(p, g.clamped, n * 3);
	take(p, g.mailbox, MAILBOX_HEAD + V);
	g.total = align_up((size_t)(p - base)) + ALIGN;
	return g;
}

ImgState ImgState::carve(char* base, size_t N, size_t T, size_t V)
{
	ImgState im;
	char* p = base;
	take(p, im.final_T, 4 * N * V);
	take(p, im.n_contrib, 2 * N * V);
	take(p, im.ranges, T * V);
	take(p, im.tile_counts, T * V);
	take(p, im.tile_cursor, T * V);
	take(p, im.tile_order, T * V);
	im.total = align_up((size_t)(p - base)) + ALIGN;
	return im;
}

BinState BinState::carve(char* base, size_t R, size_t VT)
{
	BinState b;
	char* p = base;
	take(p, b.entries, R);
	take(p, b.point_list, R);
	take(p, b.slab, R * SLAB_FLOATS);
	take(p, b.block_mask, R);
	take(p, b.bwd_rec, R * BWD_REC_FLOATS);
	take(p, b.contrib, contrib_slots(R, VT) * CONTRIB_SLOT_WORDS);
	b.total = align_up((size_t)(p - base)) + ALIGN;
	return b;
}

namespace {

constexpr int SCAN_THREADS = 1024;

// One CTA: exclusive scan of the V*T tile counts.  ranges[t] = (start, start+count), (0,0) for empty
// tiles (the reference memsets ranges and only touched tiles are written, rasterizer_impl.cu:365).
// mailbox = {R_total, overflow, max count, 0, R_view[0..V-1]}.
__global__ void __launch_bounds__(SCAN_THREADS)
tile_scan_kernel(int VT, int T, int V, const uint32_t* __restrict__ counts, uint2* __restrict__ ranges,
                 uint32_t* __restrict__ cursor, int32_t* __restrict__ mailbox, long long capacity,
                 uint32_t* __restrict__ order, int save_contrib, volatile int32_t* host_mail, int32_t host_seq)
{
	__shared__ uint32_t s_part[SCAN_THREADS];
	__shared__ uint32_t s_max[SCAN_THREADS / 32];
	__shared__ uint32_t s_bucket[SCAN_THREADS], s_bmax;
	pdl_trigger();
	pdl_wait();                    // the tile counts come from the preprocess
	const int tid = threadIdx.x;
	const int per = (VT + SCAN_THREADS - 1) / SCAN_THREADS;
	const int lo = min(VT, tid * per), hi = min(VT, lo + per);
	uint32_t sum = 0, mx = 0;
	for (int i = lo; i < hi; i++) { const uint32_t c = counts[i]; sum += c; mx = max(mx, c); }
	s_part[tid] = sum;
	mx = __reduce_max_sync(0xffffffffu, mx);
	if ((tid & 31) == 0) s_max[tid >> 5] = mx;
	__syncthreads();
	// inclusive scan over the 1024 partials: shuffle scan inside each warp, then over the 32 warp totals
	{
		const int lane = tid & 31, wid = tid >> 5;
		uint32_t v = sum;
#pragma unroll
		for (int off = 1; off < 32; off <<= 1) {
			const uint32_t u = __shfl_up_sync(0xffffffffu, v, off);
			if (lane >= off) v += u;
		}
		__shared__ uint32_t s_wsum[SCAN_THREADS / 32];
		if (lane == 31) s_wsum[wid] = v;
		__syncthreads();
		if (wid == 0) {
			uint32_t t = s_wsum[lane];
#pragma unroll
			for (int off = 1; off < 32; off <<= 1) {
				const uint32_t u = __shfl_up_sync(0xffffffffu, t, off);
				if (lane >= off) t += u;
			}
			s_wsum[lane] = t;
		}
		__syncthreads();
		s_part[tid] = v + (wid ? s_wsum[wid - 1] : 0u);
		__syncthreads();
	}
	uint32_t run = s_part[tid] - sum;   // exclusive prefix of this thread's chunk
	for (int i = lo; i < hi; i++) {
		const uint32_t c = counts[i];
		ranges[i] = c ? make_uint2(run, run + c) : make_uint2(0u, 0u);
		cursor[i] = run;
		run += c;
		// a view's last tile: R of that view = end of its last tile - start of its first
		if ((i + 1) % T == 0) mailbox[MAILBOX_HEAD + i / T] = (int32_t)run;   // inclusive prefix; differenced below
	}
	__syncthreads();
	if (tid == 0) {
		const uint32_t total = s_part[SCAN_THREADS - 1];
		uint32_t m = 0;
		for (int k = 0; k < SCAN_THREADS / 32; k++) m = max(m, s_max[k]);
		const int32_t over = ((long long)total > capacity) ? 1 : 0;
		mailbox[0] = (int32_t)total;
		mailbox[1] = over;
		mailbox[2] = (int32_t)m;
		mailbox[3] = save_contrib;
		int32_t prev = 0;
		for (int v = 0; v < V; v++) {   // per-view R from the inclusive prefixes written above
			const int32_t inc = mailbox[MAILBOX_HEAD + v];
			mailbox[MAILBOX_HEAD + v] = inc - prev;
			if (host_mail) host_mail[MAILBOX_HEAD + v] = inc - prev;
			prev = inc;
		}
		if (host_mail) {
			// num_rendered hand-off without a copy in the stream: the mailbox goes straight to mapped pinned host memory
			// (posted PCIe writes) and the sequence number is released behind it; the host polls the sequence word.
			host_mail[0] = (int32_t)total;
			host_mail[1] = over;
			host_mail[2] = (int32_t)m;
			host_mail[3] = save_contrib;
			__threadfence_system();
			host_mail[MAILBOX_HEAD + GOF_MAX_VIEWS] = host_seq;
		}
	}
	if (order != nullptr) {
		// Launch order of the blend CTAs: longest tile lists first (counting sort over 1024 length classes, one per
		// thread), so that the last wave of a launch is made of short tiles instead of whatever the row-major order leaves.
		if (tid == 0) {
			uint32_t m = 0;
			for (int k = 0; k < SCAN_THREADS / 32; k++) m = max(m, s_max[k]);
			s_bmax = m;
		}
		s_bucket[tid] = 0;
		__syncthreads();
		const float scale = (float)SCAN_THREADS / (float)(s_bmax + 1u);      // class = floor(count * scale), monotone in count
		auto bucket_of = [&](uint32_t c) { return (uint32_t)(SCAN_THREADS - 1) - min((uint32_t)(SCAN_THREADS - 1), (uint32_t)((float)c * scale)); };
		for (int i = lo; i < hi; i++) atomicAdd(&s_bucket[bucket_of(counts[i])], 1u);
		__syncthreads();
		{	// exclusive scan of the class sizes (same two-level shuffle scan as above)
			const int lane = tid & 31, wid = tid >> 5;
			const uint32_t mine = s_bucket[tid];
			uint32_t v = mine;
#pragma unroll
			for (int off = 1; off < 32; off <<= 1) {
				const uint32_t u = __shfl_up_sync(0xffffffffu, v, off);
				if (lane >= off) v += u;
			}
			__shared__ uint32_t s_bsum[SCAN_THREADS / 32];
			if (lane == 31) s_bsum[wid] = v;
			__syncthreads();
			if (wid == 0) {
				uint32_t t = s_bsum[lane];
#pragma unroll
				for (int off = 1; off < 32; off <<= 1) {
					const uint32_t u = __shfl_up_sync(0xffffffffu, t, off);
					if (lane >= off) t += u;
				}
				s_bsum[lane] = t;
			}
			__syncthreads();
			s_bucket[tid] = v - mine + (wid ? s_bsum[wid - 1] : 0u);
		}
		__syncthreads();
		for (int i = lo; i < hi; i++) order[atomicAdd(&s_bucket[bucket_of(counts[i])], 1u)] = (uint32_t)i;
		__syncthreads();
	}
}

// One thread per (view, Gaussian): claim a slot in every touched tile's bucket.  Slots are claimed
// per BLOCK: duplicates are counted in a shared-memory histogram, one global atomic per touched tile
// reserves the block's share of the bucket, and threads then take their slots from the shared
// counters (grids above SCATTER_TILES tiles claim straight from the global cursors).
constexpr int SCATTER_TILES = 4096;
__global__ void __launch_bounds__(256)
scatter_kernel(int P, int T, dim3 grid, const uint32_t* __restrict__ tiles_touched, const ushort4* __restrict__ rect,
               const float* __restrict__ depths, uint32_t* __restrict__ cursor, uint64_t* __restrict__ entries,
               const int32_t* __restrict__ mailbox)
{
	__shared__ uint32_t s_cur[SCATTER_TILES];
	pdl_trigger();
	pdl_wait();                    // mailbox / cursors come from the tile scan
	if (mailbox[1]) return;   // binning capacity exceeded (sync-free mode): outputs are invalid, write nothing
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	const int view = blockIdx.y;
	const size_t g = (size_t)view * P + idx;
	const bool active = idx < P && tiles_touched[g] != 0;
	ushort4 r = make_ushort4(0, 0, 0, 0);
	uint64_t e = 0;
	if (active) {
		r = rect[g];
		e = ((uint64_t)__float_as_uint(depths[g]) << 32) | (uint32_t)idx;
	}
	uint32_t* cur = cursor + (size_t)view * T;
	if (T <= SCATTER_TILES) {
		for (int i = threadIdx.x; i < T; i += 256) s_cur[i] = 0;
		__syncthreads();
		for (uint32_t y = r.y; y < r.w; y++)
			for (uint32_t x = r.x; x < r.z; x++) atomicAdd(&s_cur[y * grid.x + x], 1u);
		__syncthreads();
		for (int i = threadIdx.x; i < T; i += 256) {
			const uint32_t c = s_cur[i];
			if (c) s_cur[i] = atomicAdd(&cur[i], c);      // first slot of this block's share
		}
		__syncthreads();
		for (uint32_t y = r.y; y < r.w; y++)
			for (uint32_t x = r.x; x < r.z; x++) entries[atomicAdd(&s_cur[y * grid.x + x], 1u)] = e;
	} else {
		for (uint32_t y = r.y; y < r.w; y++)
			for (uint32_t x = r.x; x < r.z; x++) entries[atomicAdd(&cur[y * grid.x + x], 1u)] = e;
	}
}

constexpr int SORT_THREADS = 256;
constexpr int SORT_CAP = 4096;   // bucket entries sorted in shared memory (32 KB)

// Bitonic network with the "flip" first stage of every merge, so that every compare-exchange puts
// the smaller value at the lower index.  Entries at index >= n are virtual +inf and never move,
// which makes the network valid for any n (no padding writes).
__device__ __forceinline__ void bitonic_sort(uint64_t* a, int n, int npad)
{
	for (int k = 2; k <= npad; k <<= 1) {
		for (int j = k >> 1; j > 0; j >>= 1) {
			for (int t = threadIdx.x; t < (npad >> 1); t += SORT_THREADS) {
				// t-th compare-exchange of this step: i = lower index, l = partner
				const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
				const int l = (j == (k >> 1)) ? (i ^ (k - 1)) : (i | j);
				if (l < n) {
					const uint64_t x = a[i], y = a[l];
					if (x > y) { a[i] = y; a[l] = x; }
				}
			}
			__syncthreads();
		}
	}
}

// The same network with the keys in REGISTERS: thread t holds elements t*E .. t*E+E-1 (E = npad / 256).
// Compare-exchanges whose partner lives in the same thread are register swaps, partners in the same warp are
// reached with shuffles, and only strides of 32*E elements or more go through shared memory -- for 1024 keys
// that is 6 of the 55 steps (12 barriers instead of 55, no shared-memory traffic in between).
// s: shared scratch of 256*E keys; on entry s[i] = key i for i < n (natural order), on exit sorted.
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int mask)
{
	const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, mask);
	const uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), mask);
	return ((uint64_t)hi << 32) | lo;
}

template <int E>
__device__ __forceinline__ void bitonic_sort_regs(uint64_t* s, int n)
{
	constexpr int NP = SORT_THREADS * E;
	const int t = threadIdx.x;
	uint64_t v[E], p[E];
#pragma unroll
	for (int r = 0; r < E; r++) {
		const int e = t * E + r;
		v[r] = e < n ? s[e] : ~0ull;          // virtual +inf padding
	}
	__syncthreads();
#pragma unroll
	for (int k = 2; k <= NP; k <<= 1) {
		// ---- first step of the merge: partner = e ^ (k-1) ("flip"); e is the lower index iff bit k/2 of e is 0
		if (k <= E) {
#pragma unroll
			for (int r = 0; r < E; r++)
				if ((r & (k >> 1)) == 0) {
					const uint64_t a = v[r], b = v[r ^ (k - 1)];
					v[r] = a < b ? a : b;
					v[r ^ (k - 1)] = a < b ? b : a;
				}
		} else {
			const int tmask = k / E - 1;               // partner thread = t ^ tmask, partner register = E-1-r
			if (k <= 32 * E) {
#pragma unroll
				for (int r = 0; r < E; r++) p[r] = shfl_xor_u64(v[E - 1 - r], tmask);
			} else {
#pragma unroll
				for (int r = 0; r < E; r++) s[r * SORT_THREADS + t] = v[r];
				__syncthreads();
#pragma unroll
				for (int r = 0; r < E; r++) p[r] = s[(E - 1 - r) * SORT_THREADS + (t ^ tmask)];
				__syncthreads();
			}
			const bool lower = (t & ((k >> 1) / E)) == 0;
#pragma unroll
			for (int r = 0; r < E; r++) v[r] = (lower == (v[r] < p[r])) ? v[r] : p[r];
		}
		// ---- remaining steps: partner = e ^ j; e is the lower index iff bit j of e is 0
#pragma unroll
		for (int j = k >> 2; j > 0; j >>= 1) {
			if (j < E) {
#pragma unroll
				for (int r = 0; r < E; r++)
					if ((r & j) == 0) {
						const uint64_t a = v[r], b = v[r | j];
						v[r] = a < b ? a : b;
						v[r | j] = a < b ? b : a;
					}
			} else {
				const int tmask = j / E;
				if (j < 32 * E) {
#pragma unroll
					for (int r = 0; r < E; r++) p[r] = shfl_xor_u64(v[r], tmask);
				} else {
#pragma unroll
					for (int r = 0; r < E; r++) s[r * SORT_THREADS + t] = v[r];
					__syncthreads();
#pragma unroll
					for (int r = 0; r < E; r++) p[r] = s[r * SORT_THREADS + (t ^ tmask)];
					__syncthreads();
				}
				const bool lower = (t & tmask) == 0;
#pragma unroll
				for (int r = 0; r < E; r++) v[r] = (lower == (v[r] < p[r])) ? v[r] : p[r];
			}
		}
	}
#pragma unroll
	for (int r = 0; r < E; r++) s[t * E + r] = v[r];
	__syncthreads();
}

#ifndef GOF_SORT_MIN_CTAS
#define GOF_SORT_MIN_CTAS 4      // 64 registers (36 bytes of spills): 92.7 -> 90.5 us per 8-view batch
#endif

// One CTA per tile of the batch: sort the bucket, write point_list and the slab.
__global__ void __launch_bounds__(SORT_THREADS, GOF_SORT_MIN_CTAS)
tile_sort_gather_kernel(int P, int T, dim3 grid, int W, int H, float focal_x, float focal_y, float ray_pad,
                        const uint2* __restrict__ ranges, uint64_t* __restrict__ entries,
                        const float* __restrict__ rec_all, uint32_t* __restrict__ point_list,
                        float* __restrict__ slab, uint8_t* __restrict__ block_mask, const int32_t* __restrict__ mailbox,
                        const uint32_t* __restrict__ tile_order, float* __restrict__ bwd_rec,
                        const float2* __restrict__ means2D_all, const float4* __restrict__ conic_opacity_all)
{
	__shared__ uint64_t s_e[SORT_CAP];
	pdl_trigger();
	pdl_wait();                    // bucket entries come from the scatter
	if (mailbox[1]) return;
	const int gt = (int)tile_order[blockIdx.x];   // global tile index (view * T + tile), largest buckets first
	const uint2 range = ranges[gt];
	const int n = (int)(range.y - range.x);
	if (n <= 0) return;
	const int view = gt / T, tile = gt - view * T;
	const int ty = tile / grid.x, tx = tile - ty * grid.x;
	uint64_t* bucket = entries + range.x;

	uint64_t* a;
	if (n <= SORT_CAP) {
		for (int i = threadIdx.x; i < n; i += SORT_THREADS) s_e[i] = bucket[i];
		a = s_e;
		__syncthreads();
		if (n <= SORT_THREADS) bitonic_sort_regs<1>(s_e, n);
		else if (n <= 2 * SORT_THREADS) bitonic_sort_regs<2>(s_e, n);
		else if (n <= 4 * SORT_THREADS) bitonic_sort_regs<4>(s_e, n);
		else if (n <= 8 * SORT_THREADS) bitonic_sort_regs<8>(s_e, n);
		else {
			int npad = 1;
			while (npad < n) npad <<= 1;
			bitonic_sort(a, n, npad);
		}
	} else {
		a = bucket;   // oversized bucket: the plain network, in place in global memory
		int npad = 1;
		while (npad < n) npad <<= 1;
		__syncthreads();
		bitonic_sort(a, n, npad);
	}

	const TileRays tr = tile_rays(tx, ty, W, H, focal_x, focal_y, (double)ray_pad);
	const float* rec = rec_all + (size_t)view * P * REC_FLOATS;
	for (int i = threadIdx.x; i < n; i += SORT_THREADS) {
		const uint64_t e = a[i];
		const uint32_t id = (uint32_t)e;
		if (n <= SORT_CAP) bucket[i] = e;        // keep the sorted keys (test accessor: point_list_keys)
		point_list[range.x + i] = id;
		const float4* src = reinterpret_cast<const float4*>(rec + (size_t)id * REC_FLOATS);
		const float4 q0 = __ldg(src), q1 = __ldg(src + 1), q2 = __ldg(src + 2), q3 = __ldg(src + 3);
		float c[6];
		conic_coefficients(q0, q1, q2, tr, c);
		block_mask[range.x + i] = (uint8_t)conic_block_mask(c, (double)ray_pad);
		float4* dst = reinterpret_cast<float4*>(slab + (size_t)(range.x + i) * SLAB_FLOATS);
		dst[0] = make_float4(c[0], c[1], c[2], c[3]);
		dst[1] = make_float4(c[4], c[5], q2.w, q0.x);          // c4 c5 w Sxx
		dst[2] = make_float4(q0.y, q0.z, q0.w, q1.x);          // Sxy Sxz Syy Syz
		dst[3] = make_float4(q1.y, q1.z, q1.w, q2.x);          // Szz Bx By Bz
		dst[4] = make_float4(q2.y, q3.x, q3.y, q3.z);          // C r g b
		if (bwd_rec != nullptr) {
			// what the backward blend needs per pair besides the slab record (backward.cu:897-909): 2-D mean, 2-D conic, id
			const float2 xy = __ldg(&means2D_all[(size_t)view * P + id]);
			const float4 co = __ldg(&conic_opacity_all[(size_t)view * P + id]);
			float4* bd = reinterpret_cast<float4*>(bwd_rec + (size_t)(range.x + i) * BWD_REC_FLOATS);
			bd[0] = make_float4(xy.x, xy.y, co.x, co.y);
			bd[1] = make_float4(co.z, __uint_as_float(id), 0.0f, 0.0f);
		}
	}
}

// ---- test accessors --------------------------------------------------------------------------
__global__ void extract_rec_kernel(size_t n, const float* __restrict__ rec, float* __restrict__ v2g, float* __restrict__ rgb)
{
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	if (v2g) for (int k = 0; k < 10; k++) v2g[i * 10 + k] = rec[i * REC_FLOATS + k];
	if (rgb) for (int k = 0; k < 3; k++) rgb[i * 3 + k] = rec[i * REC_FLOATS + REC_RGB + k];
}

// point_list_keys as the reference stores them: (tile << 32) | float_bits(depth), per view tile ids.
__global__ void extract_keys_kernel(int VT, int T, const uint2* __restrict__ ranges, const uint64_t* __restrict__ entries,
                                    uint64_t* __restrict__ keys)
{
	const int gt = blockIdx.x;
	const uint2 r = ranges[gt];
	const uint64_t tile = (uint64_t)(gt % T);
	for (uint32_t i = r.x + threadIdx.x; i < r.y; i += blockDim.x) keys[i] = (tile << 32) | (entries[i] >> 32);
}

// Inclusive scan of tiles_touched per view (the reference's point_offsets); single CTA per view.
__global__ void __launch_bounds__(1024)
extract_offsets_kernel(int P, const uint32_t* __restrict__ tiles_touched, uint32_t* __restrict__ offsets)
{
	__shared__ uint32_t s_part[1024];
	const int tid = threadIdx.x;
	const uint32_t* tt = tiles_touched + (size_t)blockIdx.x * P;
	uint32_t* out = offsets + (size_t)blockIdx.x * P;
	const int per = (P + 1023) / 1024;
	const int lo = min(P, tid * per), hi = min(P, lo + per);
	uint32_t sum = 0;
	for (int i = lo; i < hi; i++) sum += tt[i];
	s_part[tid] = sum;
	__syncthreads();
	for (int off = 1; off < 1024; off <<= 1) {
		const uint32_t v = (tid >= off) ? s_part[tid - off] : 0;
		__syncthreads();
		s_part[tid] += v;
		__syncthreads();
	}
	uint32_t run = s_part[tid] - sum;
	for (int i = lo; i < hi; i++) { run += tt[i]; out[i] = run; }
}

}  // namespace

int launch_tile_scan(const Frame& f, const GeomState& g, const ImgState& im, int64_t capacity, cudaStream_t s, int save_contrib,
                     int32_t* host_mail, int32_t host_seq)
{
	GOF_CUDA_CHECK(launch_chained(PDL_SCAN, tile_scan_kernel, dim3(1), dim3(SCAN_THREADS), 0, s, f.V * f.T, f.T, f.V, im.tile_counts, im.ranges,
	                              im.tile_cursor, g.mailbox, (long long)capacity, im.tile_order, save_contrib,
	                              (volatile int32_t*)host_mail, host_seq));
	return GOF_OK;
}

int launch_tile_scan_raw(int T, const uint32_t* counts, uint2* ranges, uint32_t* cursor, int32_t* mailbox, cudaStream_t s)
{
	GOF_CUDA_CHECK(launch_chained(PDL_SCAN, tile_scan_kernel, dim3(1), dim3(SCAN_THREADS), 0, s, T, T, 1, counts, ranges, cursor, mailbox,
	                              (long long)1 << 40, (uint32_t*)nullptr, 0, (volatile int32_t*)nullptr, 0));
	return GOF_OK;
}

int launch_binning(const Frame& f, const GeomState& g, const ImgState& im, const BinState& b, int64_t capacity,
                   cudaStream_t s, float ray_pad, int for_backward)
{
	if (capacity <= 0) return GOF_OK;
	dim3 blocks((f.P + 255) / 256, f.V);
	GOF_CUDA_CHECK(launch_chained(PDL_SCATTER, scatter_kernel, blocks, dim3(256), 0, s, f.P, f.T, f.grid, g.tiles_touched, g.rect, g.depths,
	                              im.tile_cursor, b.entries, g.mailbox));
	GOF_CUDA_CHECK(launch_chained(PDL_SORT, tile_sort_gather_kernel, dim3(f.V * f.T), dim3(SORT_THREADS), 0, s, f.P, f.T, f.grid, f.W, f.H,
	                              f.focal_x, f.focal_y, ray_pad, im.ranges, b.entries, g.rec, b.point_list, b.slab, b.block_mask,
	                              g.mailbox, im.tile_order, for_backward ? b.bwd_rec : (float*)nullptr, g.means2D, g.conic_opacity));
	return GOF_OK;
}

int launch_extract(const char* what, const Frame& f, const GeomState& g, const ImgState& im, const BinState& b,
                   int64_t R, void* dst, cudaStream_t s)
{
	const size_t n = (size_t)f.P * f.V;
	const char w = what[0];
	if (w == 'v' || w == 'r') {          // view2gaussian / rgb
		if (n) extract_rec_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, g.rec, w == 'v' ? (float*)dst : nullptr,
		                                                                      w == 'r' ? (float*)dst : nullptr);
	} else if (w == 'k') {               // point_list_keys
		if (R > 0) extract_keys_kernel<<<f.V * f.T, 128, 0, s>>>(f.V * f.T, f.T, im.ranges, b.entries, (uint64_t*)dst);
	} else if (w == 'o') {               // point_offsets
		if (n) extract_offsets_kernel<<<f.V, 1024, 0, s>>>(f.P, g.tiles_touched, (uint32_t*)dst);
	}
	GOF_CUDA_CHECK(cudaGetLastError());
	return GOF_OK;
}

}  // namespace gof
