// binning.cu -- tile binning: offsets scan (K2), key/value duplication (K4), stable radix
// sort (K5), tile ranges (K6/K7) and the tile-ordered slab gather.
//
// Replaces cub::DeviceScan::InclusiveSum + duplicateWithKeys + cub::DeviceRadixSort::SortPairs
// + identifyTileRanges of the reference (RAST/cuda_rasterizer/rasterizer_impl.cu:70-111,
// 149-171,332-373).  Integer contract (bit-exact): key = (tile_id << 32) | float_bits(depth),
// value = Gaussian index, emitted y-major per Gaussian in index order; stable sort on the low
// 32+bit bits (bit = "higher MSB" of the tile count, rasterizer_impl.cu:35-50); ranges[tile] =
// [first,last+1), (0,0) for untouched tiles.
//
// On top of the reference's outputs this stage writes the tile-ordered *slab*: the 64-byte
// blend record of every sorted duplicate, contiguous per tile, so that the blend kernels can
// stream a tile's Gaussians with TMA bulk copies instead of gathering through point_list.
#include "gof_common.cuh"
#include <cub/cub.cuh>

namespace gof {

size_t scan_temp_bytes(size_t P)
{
	size_t bytes = 0;
	cub::DeviceScan::InclusiveSum(nullptr, bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)P);
	return bytes;
}

size_t sort_temp_bytes(size_t R)
{
	size_t bytes = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, bytes, (uint64_t*)nullptr, (uint64_t*)nullptr,
	                                (uint32_t*)nullptr, (uint32_t*)nullptr, (int)R);
	return bytes;
}

template <typename T>
static void take(char*& p, T*& ptr, size_t count)
{
	p = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(p)));
	ptr = reinterpret_cast<T*>(p);
	p += count * sizeof(T);
}

GeomState GeomState::carve(char* base, size_t P)
{
	GeomState g;
	char* p = base;
	take(p, g.depths, P);
	take(p, g.means2D, P);
	take(p, g.conic_opacity, P);
	take(p, g.rec, P * REC_FLOATS);
	take(p, g.tiles_touched, P);
	take(p, g.point_offsets, P);
	take(p, g.clamped, P * 3);
	take(p, g.mailbox, 4);
	g.scan_temp_bytes = gof::scan_temp_bytes(P);
	take(p, g.scan_temp, g.scan_temp_bytes);
	g.total = align_up((size_t)(p - base)) + ALIGN;
	return g;
}

ImgState ImgState::carve(char* base, size_t N, size_t T)
{
	ImgState im;
	char* p = base;
	take(p, im.final_T, 4 * N);
	take(p, im.n_contrib, 2 * N);
	take(p, im.ranges, T);
	im.total = align_up((size_t)(p - base)) + ALIGN;
	return im;
}

BinState BinState::carve(char* base, size_t R)
{
	BinState b;
	char* p = base;
	take(p, b.keys_unsorted, R);
	take(p, b.keys, R);
	take(p, b.vals_unsorted, R);
	take(p, b.point_list, R);
	take(p, b.slab, R * REC_FLOATS);
	b.sort_temp_bytes = gof::sort_temp_bytes(R);
	take(p, b.sort_temp, b.sort_temp_bytes);
	b.total = align_up((size_t)(p - base)) + ALIGN;
	return b;
}

namespace {

// Smallest b with (n >> b) == 0 found by the reference's halving search
// (rasterizer_impl.cu:35-50); restated, it returns floor(log2 n) + 1 for n >= 1.
uint32_t higher_msb(uint32_t n)
{
	uint32_t msb = sizeof(n) * 4, step = msb;
	while (step > 1) {
		step /= 2;
		if (n >> msb) msb += step; else msb -= step;
	}
	if (n >> msb) msb++;
	return msb;
}

__global__ void duplicate_kernel(int P, const float2* __restrict__ means2D, const float* __restrict__ depths,
                                 const uint32_t* __restrict__ offsets, uint64_t* __restrict__ keys,
                                 uint32_t* __restrict__ vals, const int* __restrict__ radii, dim3 grid)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= P) return;
	const int rad = radii[idx];
	if (rad <= 0) return;
	uint32_t off = (idx == 0) ? 0 : offsets[idx - 1];
	const float2 p = means2D[idx];
	const uint32_t x0 = min(grid.x, max((int)0, (int)((p.x - rad) / TILE_X)));
	const uint32_t y0 = min(grid.y, max((int)0, (int)((p.y - rad) / TILE_Y)));
	const uint32_t x1 = min(grid.x, max((int)0, (int)((p.x + rad + TILE_X - 1) / TILE_X)));
	const uint32_t y1 = min(grid.y, max((int)0, (int)((p.y + rad + TILE_Y - 1) / TILE_Y)));
	const uint64_t dbits = __float_as_uint(depths[idx]);
	for (uint32_t y = y0; y < y1; y++)
		for (uint32_t x = x0; x < x1; x++) {
			keys[off] = ((uint64_t)(y * grid.x + x) << 32) | dbits;
			vals[off] = idx;
			off++;
		}
}

// One thread per sorted duplicate: tile-range boundaries + gather of the blend record.
__global__ void ranges_gather_kernel(int L, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ point_list,
                                     const float* __restrict__ rec, uint2* __restrict__ ranges, float* __restrict__ slab)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= L) return;
	const uint32_t tile = keys[idx] >> 32;
	if (idx == 0) ranges[tile].x = 0;
	else {
		const uint32_t prev = keys[idx - 1] >> 32;
		if (tile != prev) { ranges[prev].y = idx; ranges[tile].x = idx; }
	}
	if (idx == L - 1) ranges[tile].y = L;

	const uint32_t id = point_list[idx];
	const float4* src = reinterpret_cast<const float4*>(rec + (size_t)id * REC_FLOATS);
	float4* dst = reinterpret_cast<float4*>(slab + (size_t)idx * REC_FLOATS);
	float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2), d = __ldg(src + 3);
	dst[0] = a; dst[1] = b; dst[2] = c; dst[3] = d;
}

}  // namespace

int launch_scan(const GeomState& g, int P, cudaStream_t s)
{
	size_t bytes = g.scan_temp_bytes;
	GOF_CUDA_CHECK(cub::DeviceScan::InclusiveSum(g.scan_temp, bytes, g.tiles_touched, g.point_offsets, P, s));
	return GOF_OK;
}

int launch_binning(const GofParams& prm, dim3 tile_grid, const GeomState& g, const ImgState& im,
                   const BinState& b, const int32_t* radii, int R, cudaStream_t s)
{
	const int P = prm.P;
	const int T = tile_grid.x * tile_grid.y;
	GOF_CUDA_CHECK(cudaMemsetAsync(im.ranges, 0, (size_t)T * sizeof(uint2), s));
	if (R <= 0) return GOF_OK;
	duplicate_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, g.means2D, g.depths, g.point_offsets, b.keys_unsorted,
	                                                 b.vals_unsorted, radii, tile_grid);
	GOF_CUDA_CHECK(cudaGetLastError());
	const int bit = higher_msb(T);
	size_t bytes = b.sort_temp_bytes;
	GOF_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(b.sort_temp, bytes, b.keys_unsorted, b.keys, b.vals_unsorted,
	                                               b.point_list, R, 0, 32 + bit, s));
	ranges_gather_kernel<<<(R + 255) / 256, 256, 0, s>>>(R, b.keys, b.point_list, g.rec, im.ranges, b.slab);
	GOF_CUDA_CHECK(cudaGetLastError());
	return GOF_OK;
}

}  // namespace gof
