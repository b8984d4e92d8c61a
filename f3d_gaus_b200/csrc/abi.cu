// abi.cu -- the extern "C" boundary of libgof_b200.so (see include/gof_b200.h) and the
// stage orchestration that replaces CudaRasterizer::Rasterizer::forward/backward
// (RAST/cuda_rasterizer/rasterizer_impl.cu:247-526).
#include "gof_common.cuh"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <atomic>
#include <vector>

namespace gof {

static thread_local char g_err[512] = "";

int pdl_mask()
{
	static const int mask = [] {
		const char* off = getenv("GOF_NO_PDL");
		if (off && off[0] && off[0] != '0') return 0;
		const char* m = getenv("GOF_PDL_MASK");
		return m && m[0] ? atoi(m) : (PDL_SCAN | PDL_SCATTER | PDL_SORT | PDL_PRE_BWD);
	}();
	return mask;
}

void set_error(const char* fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

}  // namespace gof

struct GofContext {
	int device = 0;
	int32_t* pinned = nullptr;      // host mailbox for num_rendered (callback mode): MAILBOX_HEAD + GOF_MAX_VIEWS ints + a sequence word
	int32_t* pinned_dev = nullptr;  // the same memory in the device's address space (the tile scan writes it directly)
	int32_t handoff_seq = 0;        // sequence number of the last direct hand-off
	float* gacc = nullptr;          // backward gradient accumulator, grown on demand
	size_t gacc_floats = 0;
	size_t gacc_used = 0;           // floats written by the last backward (P * V * GACC_FLOATS)
	int64_t spec_capacity = 0;      // callback mode: binning capacity to allocate speculatively (1.25 x the last R)
	cudaEvent_t handoff = nullptr;  // completion of the mailbox copy
	float* sink = nullptr;          // frame sink of the NEXT forward call (device-visible address), gof_set_frame_sink
	size_t sink_bytes = 0;
	int sink_layout = 0;
	// optional per-stage CUDA-event timing (gof_profile_*): one event per stage boundary
	bool profiling = false;
	std::vector<cudaEvent_t> pool;          // reusable events
	std::vector<std::vector<cudaEvent_t>> calls[2];   // [0] forward calls, [1] backward calls
};

static cudaEvent_t prof_event(GofContext* c, cudaStream_t s)
{
	cudaEvent_t e;
	if (!c->pool.empty()) { e = c->pool.back(); c->pool.pop_back(); }
	else cudaEventCreate(&e);
	cudaEventRecord(e, s);
	return e;
}
#define GOF_PROF_MARK(ctx, vec, s) do { if ((ctx)->profiling) (vec).push_back(prof_event((ctx), (s))); } while (0)

using namespace gof;

#define GOF_STAGE_CHECK(prm, s)                                   \
	do {                                                          \
		if ((prm)->debug) {                                       \
			GOF_CUDA_CHECK(cudaStreamSynchronize(s));             \
			GOF_CUDA_CHECK(cudaGetLastError());                   \
		}                                                         \
	} while (0)

extern "C" {

const char* gof_last_error(void) { return g_err; }
const char* gof_version(void) { return "gof_b200 0.1 (sm_100a)"; }

int gof_context_create(int device, GofContext** out)
{
	if (!out) { set_error("gof_context_create: out is NULL"); return GOF_EINVAL; }
	GOF_CUDA_CHECK(cudaSetDevice(device));
	GofContext* c = new GofContext();
	c->device = device;
	cudaError_t e = cudaHostAlloc(&c->pinned, (MAILBOX_HEAD + GOF_MAX_VIEWS + 1) * sizeof(int32_t), cudaHostAllocMapped);
	if (e != cudaSuccess) { delete c; set_error("cudaHostAlloc failed: %s", cudaGetErrorString(e)); return GOF_ECUDA; }
	memset(c->pinned, 0, (MAILBOX_HEAD + GOF_MAX_VIEWS + 1) * sizeof(int32_t));
	if (cudaHostGetDevicePointer(&c->pinned_dev, c->pinned, 0) != cudaSuccess) { c->pinned_dev = nullptr; cudaGetLastError(); }   // fall back to a copy
	*out = c;
	return GOF_OK;
}

void gof_context_destroy(GofContext* c)
{
	if (!c) return;
	if (c->pinned) cudaFreeHost(c->pinned);
	if (c->gacc) cudaFree(c->gacc);
	if (c->handoff) cudaEventDestroy(c->handoff);
	for (cudaEvent_t e : c->pool) cudaEventDestroy(e);
	for (int k = 0; k < 2; k++) for (auto& m : c->calls[k]) for (cudaEvent_t e : m) cudaEventDestroy(e);
	delete c;
}

int gof_profile_enable(GofContext* ctx, int on)
{
	if (!ctx) { set_error("gof_profile_enable: NULL context"); return GOF_EINVAL; }
	ctx->profiling = on != 0;
	return GOF_OK;
}

// Sums the CUDA-event durations (ms) of every profiled call since the last read.
//   fwd_ms[5]: preprocess, scan, num_rendered hand-off (D2H + sync), binning, blend
//   bwd_ms[3]: accumulator clear, blend backward, preprocess backward
int gof_profile_read(GofContext* ctx, double* fwd_ms, int64_t* fwd_calls, double* bwd_ms, int64_t* bwd_calls)
{
	if (!ctx) { set_error("gof_profile_read: NULL context"); return GOF_EINVAL; }
	GOF_CUDA_CHECK(cudaDeviceSynchronize());
	for (int k = 0; k < 2; k++) {
		const int nst = k == 0 ? 5 : 3;
		double* out = k == 0 ? fwd_ms : bwd_ms;
		if (out) for (int i = 0; i < nst; i++) out[i] = 0.0;
		int64_t n = 0;
		for (auto& marks : ctx->calls[k]) {
			if ((int)marks.size() == nst + 1) {
				n++;
				for (int i = 0; i < nst; i++) {
					float ms = 0.f;
					cudaEventElapsedTime(&ms, marks[i], marks[i + 1]);
					if (out) out[i] += ms;
				}
			}
			for (cudaEvent_t e : marks) ctx->pool.push_back(e);
		}
		ctx->calls[k].clear();
		if (k == 0 && fwd_calls) *fwd_calls = n;
		if (k == 1 && bwd_calls) *bwd_calls = n;
	}
	return GOF_OK;
}

static Frame make_frame(const GofParams* prm, int V)
{
	Frame f;
	f.P = prm->P; f.V = V; f.W = prm->W; f.H = prm->H;
	f.grid = dim3((prm->W + TILE_X - 1) / TILE_X, (prm->H + TILE_Y - 1) / TILE_Y, 1);
	f.T = (int)(f.grid.x * f.grid.y);
	f.focal_y = prm->H / (2.0f * prm->tan_fovy);
	f.focal_x = prm->W / (2.0f * prm->tan_fovx);
	return f;
}

int gof_state_sizes_batch(int32_t P, int32_t W, int32_t H, int32_t V, int64_t num_rendered,
                          size_t* geom_bytes, size_t* img_bytes, size_t* binning_bytes)
{
	if (P < 0 || W <= 0 || H <= 0 || V <= 0 || num_rendered < 0) { set_error("gof_state_sizes: bad sizes"); return GOF_EINVAL; }
	const size_t T = (size_t)((W + TILE_X - 1) / TILE_X) * ((H + TILE_Y - 1) / TILE_Y);
	if (geom_bytes) *geom_bytes = GeomState::carve(nullptr, (size_t)P, (size_t)V).total;
	if (img_bytes) *img_bytes = ImgState::carve(nullptr, (size_t)W * H, T, (size_t)V).total;
	if (binning_bytes) *binning_bytes = BinState::carve(nullptr, (size_t)num_rendered, T * (size_t)V).total;
	return GOF_OK;
}

int gof_state_sizes(int32_t P, int32_t W, int32_t H, int64_t num_rendered,
                    size_t* geom_bytes, size_t* img_bytes, size_t* binning_bytes)
{
	return gof_state_sizes_batch(P, W, H, 1, num_rendered, geom_bytes, img_bytes, binning_bytes);
}

static char* align_base(const void* p) { return reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(p))); }

// Largest R whose binning layout fits in `bytes` (host arithmetic only).
static int64_t binning_capacity(size_t bytes, size_t VT)
{
	int64_t lo = 0, hi = (int64_t)1 << 31;
	while (lo + 1 < hi) {
		const int64_t mid = (lo + hi) / 2;
		if (BinState::carve(nullptr, (size_t)mid, VT).total <= bytes) lo = mid; else hi = mid;
	}
	return lo;
}

// The binning blob's layout is a pure function of (address, byte size, tiles in the batch): the capacity it is carved
// for is the largest R that fits.  Forward, backward and the test accessor all derive it this way, so a blob that was
// allocated speculatively (capacity > num_rendered) is decoded identically by every later call that is handed
// the same tensor -- no per-process side table, nothing to go stale between a forward and its backward.
static int64_t blob_capacity(const void* blob, size_t bytes, size_t VT)
{
	if (!blob) return 0;
	const size_t skew = (size_t)(align_base(blob) - (const char*)blob);
	if (bytes <= skew || BinState::carve(nullptr, 0, VT).total > bytes - skew) return 0;
	return binning_capacity(bytes - skew, VT);
}

int gof_set_frame_sink(GofContext* ctx, void* sink, size_t sink_bytes, int32_t layout)
{
	if (!ctx) { set_error("gof_set_frame_sink: ctx is NULL"); return GOF_EINVAL; }
	ctx->sink = nullptr;
	ctx->sink_bytes = 0;
	if (!sink) return GOF_OK;
	if (layout != GOF_SINK_CHW && layout != GOF_SINK_HWC) { set_error("gof_set_frame_sink: unknown layout %d", layout); return GOF_EINVAL; }
	if ((uintptr_t)sink % 16 != 0) { set_error("gof_set_frame_sink: the sink must be 16-byte aligned"); return GOF_EINVAL; }
	cudaPointerAttributes at;
	GOF_CUDA_CHECK(cudaPointerGetAttributes(&at, sink));
	if (at.type == cudaMemoryTypeUnregistered || at.devicePointer == nullptr) {
		set_error("gof_set_frame_sink: %p is pageable host memory; the sink must be device memory or pinned (page-locked) host memory", sink);
		return GOF_EINVAL;
	}
	ctx->sink = static_cast<float*>(at.devicePointer);   // pinned host memory: its address in the device's space
	ctx->sink_bytes = sink_bytes;
	ctx->sink_layout = layout;
	return GOF_OK;
}

int gof_forward_batch(GofContext* ctx, const GofParams* prm, const GofInputs* in, int32_t V, int32_t bg_stride,
                      void* geom, size_t geom_bytes, void* img, size_t img_bytes,
                      void* binning, size_t binning_bytes, GofAllocFn alloc, void* alloc_user,
                      float* out_color, int32_t* radii,
                      int32_t* num_rendered, void** binning_out, gof_stream_t stream)
{
	if (!ctx || !prm || !in) { set_error("gof_forward: NULL argument"); return GOF_EINVAL; }
	// one-shot frame sink (gof_set_frame_sink): consumed by this call whatever its outcome
	float* const sink = ctx->sink;
	const size_t sink_bytes = ctx->sink_bytes;
	const int sink_hwc = ctx->sink_layout == GOF_SINK_HWC;
	ctx->sink = nullptr;
	ctx->sink_bytes = 0;
	cudaStream_t s = (cudaStream_t)stream;
	const int P = prm->P, W = prm->W, H = prm->H;
	if (W <= 0 || H <= 0 || P < 0 || V <= 0) { set_error("gof_forward: bad sizes P=%d W=%d H=%d V=%d", P, W, H, V); return GOF_EINVAL; }
	if (V > GOF_MAX_VIEWS) { set_error("gof_forward: at most %d views per batch", GOF_MAX_VIEWS); return GOF_EINVAL; }
	if (!out_color) { set_error("gof_forward: out_color is NULL"); return GOF_EINVAL; }
	if (num_rendered) for (int v = 0; v < V; v++) num_rendered[v] = 0;
	if (binning_out) *binning_out = binning;
	const size_t N = (size_t)W * H;
	if (sink && sink_bytes < (size_t)V * SINK_CH * N * sizeof(float)) {
		set_error("gof_forward: frame sink too small (%zu < %zu)", sink_bytes, (size_t)V * SINK_CH * N * sizeof(float));
		return GOF_ENOMEM;
	}
	if (P == 0) {   // rasterize_points.cu:85: nothing is launched, outputs stay zero
		GOF_CUDA_CHECK(cudaMemsetAsync(out_color, 0, (size_t)V * N * OUT_CH * sizeof(float), s));
		if (sink) GOF_CUDA_CHECK(cudaMemsetAsync(sink, 0, (size_t)V * SINK_CH * N * sizeof(float), s));
		return GOF_OK;
	}
	if (!in->means3D || !in->opacities || !in->viewmatrix || !in->projmatrix || !in->campos || !in->background || !radii) {
		set_error("gof_forward: a required input pointer is NULL");
		return GOF_EINVAL;
	}
	if ((in->shs == nullptr) == (in->colors_precomp == nullptr)) {
		set_error("gof_forward: provide exactly one of shs / colors_precomp");
		return GOF_EINVAL;
	}
	if (((in->scales == nullptr) || (in->rotations == nullptr)) == (in->cov3D_precomp == nullptr)) {
		set_error("gof_forward: provide exactly one of (scales, rotations) / cov3D_precomp");
		return GOF_EINVAL;
	}
	if (in->cov3D_precomp && !in->view2gaussian_precomp) {
		set_error("gof_forward: cov3D_precomp needs view2gaussian_precomp (no scales/rotations to build the quadric)");
		return GOF_EINVAL;
	}
	if (in->view2gaussian_precomp && V > 1) {
		set_error("gof_forward_batch: view2gaussian_precomp is per view; use V == 1");
		return GOF_EINVAL;
	}
	if (in->shs && prm->M < (prm->D + 1) * (prm->D + 1)) {
		set_error("gof_forward: sh has %d coefficients, degree %d needs %d", prm->M, prm->D, (prm->D + 1) * (prm->D + 1));
		return GOF_EINVAL;
	}

	const Frame f = make_frame(prm, V);
	const size_t VT = (size_t)f.T * (size_t)V;
	const int save_contrib = (prm->flags & GOF_FLAG_SAVE_CONTRIB) ? 1 : 0;
	GeomState g = GeomState::carve(align_base(geom), (size_t)P, (size_t)V);
	ImgState im = ImgState::carve(align_base(img), N, (size_t)f.T, (size_t)V);
	if (!geom || g.total > geom_bytes) { set_error("gof_forward: geom blob too small (%zu < %zu)", geom_bytes, g.total); return GOF_ENOMEM; }
	if (!img || im.total > img_bytes) { set_error("gof_forward: img blob too small (%zu < %zu)", img_bytes, im.total); return GOF_ENOMEM; }

	int rc;
	std::vector<cudaEvent_t> marks;
	GOF_PROF_MARK(ctx, marks, s);
	if ((rc = launch_preprocess(*prm, *in, f, g, im, radii, s)) != GOF_OK) return rc;
	GOF_STAGE_CHECK(prm, s);
	GOF_PROF_MARK(ctx, marks, s);

	int64_t capacity;
	BinState b;
	if (binning != nullptr) {
		// sync-free mode: capacity is whatever fits in the caller's blob; R stays on the device
		capacity = blob_capacity(binning, binning_bytes, VT);
		if ((rc = launch_tile_scan(f, g, im, capacity, s, save_contrib)) != GOF_OK) return rc;
		GOF_STAGE_CHECK(prm, s);
		GOF_PROF_MARK(ctx, marks, s);
		b = BinState::carve(align_base(binning), (size_t)capacity, VT);
		if (num_rendered) for (int v = 0; v < V; v++) num_rendered[v] = -1;
	} else {
		if (!alloc) { set_error("gof_forward: neither a binning blob nor an allocation callback"); return GOF_EINVAL; }
		if (!ctx->handoff) GOF_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->handoff, cudaEventDisableTiming));
		bool done = false;
		if (ctx->spec_capacity > 0 && !prm->debug) {
			// Speculative hand-off: allocate the binning blob for 1.25x the previous call's num_rendered and enqueue
			// EVERYTHING before waiting for R, so the GPU never idles behind the host round trip of the reference's
			// protocol (rasterizer_impl.cu:336-340).  The host still returns R as soon as the scan has finished.
			const size_t need = BinState::carve(nullptr, (size_t)ctx->spec_capacity, VT).total;
			void* blob = alloc(alloc_user, need);
			if (!blob) { set_error("gof_forward: binning allocation callback returned NULL for %zu bytes", need); return GOF_ENOMEM; }
			capacity = blob_capacity(blob, need, VT);      // the layout every later call derives from (blob, need)
			b = BinState::carve(align_base(blob), (size_t)capacity, VT);
			// num_rendered reaches the host without a copy in the stream (a copy between the scan and the scatter costs
			// ~10 us of GPU time per frame and breaks their programmatic dependency): the scan stores the mailbox into
			// mapped pinned memory and releases a sequence number behind it, the host polls that word.
			const bool direct = ctx->pinned_dev != nullptr && !ctx->profiling;
			const int32_t seq = ++ctx->handoff_seq;
			if ((rc = launch_tile_scan(f, g, im, capacity, s, save_contrib, direct ? ctx->pinned_dev : nullptr, seq)) != GOF_OK) return rc;
			GOF_PROF_MARK(ctx, marks, s);
			if (!direct) {
				GOF_CUDA_CHECK(cudaMemcpyAsync(ctx->pinned, g.mailbox, (MAILBOX_HEAD + V) * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
				GOF_CUDA_CHECK(cudaEventRecord(ctx->handoff, s));
			}
			GOF_PROF_MARK(ctx, marks, s);
			if ((rc = launch_binning(f, g, im, b, capacity, s, 0.0f, save_contrib)) != GOF_OK) return rc;
			GOF_PROF_MARK(ctx, marks, s);
			if ((rc = launch_render_fwd(*prm, f, g, im, b, in->background, bg_stride, out_color, sink, sink_hwc, s)) != GOF_OK) return rc;
			GOF_PROF_MARK(ctx, marks, s);
			if (direct) {
				// the event (behind every kernel of the call) only ends the wait if the scan never ran (a failed launch)
				GOF_CUDA_CHECK(cudaEventRecord(ctx->handoff, s));
				volatile int32_t* seq_word = ctx->pinned + MAILBOX_HEAD + GOF_MAX_VIEWS;
				for (uint32_t spins = 1; *seq_word != seq; spins++) {
					if ((spins & 0x3ffu) == 0) {
						const cudaError_t q = cudaEventQuery(ctx->handoff);
						if (q == cudaErrorNotReady) continue;
						if (q != cudaSuccess || *seq_word != seq) {
							set_error("gof_forward: the tile scan did not deliver num_rendered (%s)", cudaGetErrorString(q == cudaSuccess ? cudaGetLastError() : q));
							return GOF_ECUDA;
						}
					}
				}
				std::atomic_thread_fence(std::memory_order_acquire);
			} else {
				GOF_CUDA_CHECK(cudaEventSynchronize(ctx->handoff));
			}
			const int64_t R = ctx->pinned[0];
			ctx->spec_capacity = R + R / 4 + 4096;
			if (!ctx->pinned[1]) {
				if (num_rendered) for (int v = 0; v < V; v++) num_rendered[v] = ctx->pinned[MAILBOX_HEAD + v];
				if (binning_out) *binning_out = blob;
				if (ctx->profiling) ctx->calls[0].push_back(std::move(marks));
				return GOF_OK;
			}
			// the guess was too small: every kernel after the scan degenerated to empty lists; redo exactly
			while (marks.size() > 2) { ctx->pool.push_back(marks.back()); marks.pop_back(); }
			done = false;
		}
		if (!done) {
			if ((rc = launch_tile_scan(f, g, im, (int64_t)1 << 40, s, save_contrib)) != GOF_OK) return rc;
			GOF_STAGE_CHECK(prm, s);
			GOF_PROF_MARK(ctx, marks, s);
			GOF_CUDA_CHECK(cudaMemcpyAsync(ctx->pinned, g.mailbox, (MAILBOX_HEAD + V) * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
			GOF_CUDA_CHECK(cudaStreamSynchronize(s));
			const int64_t R = ctx->pinned[0];
			ctx->spec_capacity = R + R / 4 + 4096;
			if (num_rendered) for (int v = 0; v < V; v++) num_rendered[v] = ctx->pinned[MAILBOX_HEAD + v];
			const size_t need = BinState::carve(nullptr, (size_t)R, VT).total;
			void* blob = alloc(alloc_user, need);
			if (!blob) { set_error("gof_forward: binning allocation callback returned NULL for %zu bytes", need); return GOF_ENOMEM; }
			if (binning_out) *binning_out = blob;
			capacity = blob_capacity(blob, need, VT);
			b = BinState::carve(align_base(blob), (size_t)capacity, VT);
		}
	}

	GOF_PROF_MARK(ctx, marks, s);   // after the num_rendered hand-off
	if ((rc = launch_binning(f, g, im, b, capacity, s, 0.0f, save_contrib)) != GOF_OK) return rc;
	GOF_STAGE_CHECK(prm, s);
	GOF_PROF_MARK(ctx, marks, s);
	if ((rc = launch_render_fwd(*prm, f, g, im, b, in->background, bg_stride, out_color, sink, sink_hwc, s)) != GOF_OK) return rc;
	GOF_STAGE_CHECK(prm, s);
	GOF_PROF_MARK(ctx, marks, s);
	if (ctx->profiling) ctx->calls[0].push_back(std::move(marks));
	return GOF_OK;
}

int gof_forward(GofContext* ctx, const GofParams* prm, const GofInputs* in,
                void* geom, size_t geom_bytes, void* img, size_t img_bytes,
                void* binning, size_t binning_bytes, GofAllocFn alloc, void* alloc_user,
                float* out_color, int32_t* radii,
                int32_t* num_rendered, void** binning_out, gof_stream_t stream)
{
	return gof_forward_batch(ctx, prm, in, 1, 0, geom, geom_bytes, img, img_bytes, binning, binning_bytes, alloc, alloc_user,
	                         out_color, radii, num_rendered, binning_out, stream);
}

int gof_integrate(GofContext* ctx, const GofParams* prm, const GofInputs* in, int32_t PN, const float* points3D,
                  void* geom, size_t geom_bytes, void* img, size_t img_bytes, GofAllocFn alloc, void* alloc_user,
                  float* out_color, int32_t* radii, float* out_alpha_integrated, float* out_color_integrated,
                  int32_t* num_rendered, gof_stream_t stream)
{
	if (!ctx || !prm || !in || !alloc) { set_error("gof_integrate: NULL argument"); return GOF_EINVAL; }
	cudaStream_t s = (cudaStream_t)stream;
	const int P = prm->P, W = prm->W, H = prm->H;
	if (W <= 0 || H <= 0 || P <= 0 || PN < 0) { set_error("gof_integrate: bad sizes P=%d PN=%d W=%d H=%d", P, PN, W, H); return GOF_EINVAL; }
	if (!out_color || !radii || (PN > 0 && (!points3D || !out_alpha_integrated || !out_color_integrated))) {
		set_error("gof_integrate: NULL output / points pointer");
		return GOF_EINVAL;
	}
	if (!in->means3D || !in->opacities || !in->viewmatrix || !in->projmatrix || !in->campos || !in->background) {
		set_error("gof_integrate: a required input pointer is NULL");
		return GOF_EINVAL;
	}
	if ((in->shs == nullptr) == (in->colors_precomp == nullptr)) { set_error("gof_integrate: provide exactly one of shs / colors_precomp"); return GOF_EINVAL; }
	if (((in->scales == nullptr) || (in->rotations == nullptr)) == (in->cov3D_precomp == nullptr)) {
		set_error("gof_integrate: provide exactly one of (scales, rotations) / cov3D_precomp");
		return GOF_EINVAL;
	}
	if (in->cov3D_precomp && !in->view2gaussian_precomp) { set_error("gof_integrate: cov3D_precomp needs view2gaussian_precomp"); return GOF_EINVAL; }
	const Frame f = make_frame(prm, 1);
	const size_t N = (size_t)W * H;
	GeomState g = GeomState::carve(align_base(geom), (size_t)P, 1);
	ImgState im = ImgState::carve(align_base(img), N, (size_t)f.T, 1);
	if (!geom || g.total > geom_bytes) { set_error("gof_integrate: geom blob too small"); return GOF_ENOMEM; }
	if (!img || im.total > img_bytes) { set_error("gof_integrate: img blob too small"); return GOF_ENOMEM; }
	int rc;
	if ((rc = launch_preprocess(*prm, *in, f, g, im, radii, s)) != GOF_OK) return rc;
	GOF_STAGE_CHECK(prm, s);
	if ((rc = launch_tile_scan(f, g, im, (int64_t)1 << 40, s)) != GOF_OK) return rc;
	GOF_CUDA_CHECK(cudaMemcpyAsync(ctx->pinned, g.mailbox, (MAILBOX_HEAD + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
	GOF_CUDA_CHECK(cudaStreamSynchronize(s));
	const int64_t R = ctx->pinned[0];
	const size_t max_count = (size_t)ctx->pinned[2];
	if (num_rendered) *num_rendered = (int32_t)R;
	const size_t bin_bytes = BinState::carve(nullptr, (size_t)R, (size_t)f.T).total;
	const size_t scr_bytes = IntegrateScratch::carve(nullptr, (size_t)PN, (size_t)f.T, max_count).total;
	char* blob = (char*)alloc(alloc_user, bin_bytes + scr_bytes + 2 * ALIGN);
	if (!blob) { set_error("gof_integrate: allocation callback returned NULL for %zu bytes", bin_bytes + scr_bytes); return GOF_ENOMEM; }
	BinState b = BinState::carve(align_base(blob), (size_t)R, (size_t)f.T);
	IntegrateScratch sc = IntegrateScratch::carve(align_base(align_base(blob) + bin_bytes), (size_t)PN, (size_t)f.T, max_count);
	if ((rc = launch_binning(f, g, im, b, R, s, 0.5f)) != GOF_OK) return rc;
	GOF_STAGE_CHECK(prm, s);
	if ((rc = launch_integrate(*prm, *in, f, g, im, b, sc, PN, points3D, out_color, out_alpha_integrated, out_color_integrated, s)) != GOF_OK) return rc;
	GOF_STAGE_CHECK(prm, s);
	return GOF_OK;
}

int gof_num_rendered(GofContext* ctx, const void* geom, int32_t P, int32_t V, gof_stream_t stream, int32_t* num_rendered)
{
	if (!ctx || !geom || !num_rendered || P <= 0 || V <= 0 || V > GOF_MAX_VIEWS) { set_error("gof_num_rendered: bad argument"); return GOF_EINVAL; }
	cudaStream_t s = (cudaStream_t)stream;
	GeomState g = GeomState::carve(align_base(geom), (size_t)P, (size_t)V);
	GOF_CUDA_CHECK(cudaMemcpyAsync(ctx->pinned, g.mailbox, (MAILBOX_HEAD + V) * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
	GOF_CUDA_CHECK(cudaStreamSynchronize(s));
	for (int v = 0; v < V; v++) num_rendered[v] = ctx->pinned[MAILBOX_HEAD + v];
	if (ctx->pinned[1]) { set_error("num_rendered=%d exceeded the binning capacity", ctx->pinned[0]); return GOF_EOVERFLOW; }
	return GOF_OK;
}

int gof_num_rendered_async(const void* geom, int32_t P, int32_t V, int32_t* host_dst, gof_stream_t stream)
{
	if (!geom || !host_dst || P <= 0 || V <= 0 || V > GOF_MAX_VIEWS) { set_error("gof_num_rendered_async: bad argument"); return GOF_EINVAL; }
	GeomState g = GeomState::carve(align_base(geom), (size_t)P, (size_t)V);
	GOF_CUDA_CHECK(cudaMemcpyAsync(host_dst, g.mailbox, (MAILBOX_HEAD + V) * sizeof(int32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
	return GOF_OK;
}

static int ensure_gacc(GofContext* ctx, size_t need, cudaStream_t s);

int gof_backward_batch(GofContext* ctx, const GofParams* prm, const GofInputs* in, int32_t V, int32_t bg_stride,
                       int64_t num_rendered, const int32_t* radii,
                       const void* geom, const void* binning, size_t binning_bytes, const void* img,
                       const float* dL_dout_color, const GofGrads* gr, gof_stream_t stream)
{
	if (!ctx || !prm || !in || !gr) { set_error("gof_backward: NULL argument"); return GOF_EINVAL; }
	cudaStream_t s = (cudaStream_t)stream;
	const int P = prm->P, W = prm->W, H = prm->H;
	if (P == 0) return GOF_OK;
	if (V <= 0 || V > GOF_MAX_VIEWS) { set_error("gof_backward: bad view count %d", V); return GOF_EINVAL; }
	if (!geom || !img || !radii || !dL_dout_color) { set_error("gof_backward: NULL state/gradient pointer"); return GOF_EINVAL; }
	if (!gr->dL_dmeans2D || !gr->dL_dcolors || !gr->dL_dopacity || !gr->dL_dmeans3D || !gr->dL_dcov3D ||
	    !gr->dL_dscales || !gr->dL_drotations || !gr->dL_dview2gaussian || (prm->M > 0 && !gr->dL_dsh)) {
		set_error("gof_backward: a gradient output pointer is NULL");
		return GOF_EINVAL;
	}
	const Frame f = make_frame(prm, V);
	GeomState g = GeomState::carve(align_base(geom), (size_t)P, (size_t)V);
	ImgState im = ImgState::carve(align_base(img), (size_t)W * H, (size_t)f.T, (size_t)V);
	BinState b{};
	if (num_rendered > 0) {
		if (!binning) { set_error("gof_backward: binning blob is NULL but num_rendered=%lld", (long long)num_rendered); return GOF_EINVAL; }
		const int64_t cap = blob_capacity(binning, binning_bytes, (size_t)f.T * (size_t)V);
		if (cap < num_rendered) {
			set_error("gof_backward: binning blob of %zu bytes holds %lld duplicates, num_rendered=%lld", binning_bytes, (long long)cap, (long long)num_rendered);
			return GOF_ENOMEM;
		}
		b = BinState::carve(align_base(binning), (size_t)cap, (size_t)f.T * (size_t)V);
	}

	const size_t need = (size_t)P * V * GACC_FLOATS;
	int rc;
	if ((rc = ensure_gacc(ctx, need, s)) != GOF_OK) return rc;
	ctx->gacc_used = need;
	std::vector<cudaEvent_t> marks;
	GOF_PROF_MARK(ctx, marks, s);
	GOF_CUDA_CHECK(cudaMemsetAsync(ctx->gacc, 0, need * sizeof(float), s));
	GOF_PROF_MARK(ctx, marks, s);
	if (num_rendered > 0) {
		if ((rc = launch_render_bwd(*prm, f, g, im, b, in->background, bg_stride, dL_dout_color, ctx->gacc, s)) != GOF_OK) return rc;
		GOF_STAGE_CHECK(prm, s);
	}
	GOF_PROF_MARK(ctx, marks, s);
	if ((rc = launch_preprocess_bwd(*prm, *in, V, g, radii, ctx->gacc, *gr, s)) != GOF_OK) return rc;
	GOF_STAGE_CHECK(prm, s);
	GOF_PROF_MARK(ctx, marks, s);
	if (ctx->profiling) ctx->calls[1].push_back(std::move(marks));
	return GOF_OK;
}

int gof_backward(GofContext* ctx, const GofParams* prm, const GofInputs* in,
                 int32_t num_rendered, const int32_t* radii,
                 const void* geom, const void* binning, size_t binning_bytes, const void* img,
                 const float* dL_dout_color, const GofGrads* gr, gof_stream_t stream)
{
	return gof_backward_batch(ctx, prm, in, 1, 0, num_rendered, radii, geom, binning, binning_bytes, img, dL_dout_color, gr, stream);
}

// ---- stage entry: per-Gaussian backward (K10) on its own -------------------------------------
__global__ void pack_gacc_kernel(int P, const float* __restrict__ dv2g, const float* __restrict__ dcol, float* __restrict__ gacc)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P) return;
	float* g = gacc + (size_t)i * GACC_FLOATS;
	for (int k = 0; k < 10; k++) g[k] = dv2g[(size_t)i * 10 + k];
	for (int k = 0; k < 3; k++) g[10 + k] = dcol ? dcol[(size_t)i * 3 + k] : 0.0f;
	for (int k = 13; k < GACC_FLOATS; k++) g[k] = 0.0f;
}

static int ensure_gacc(GofContext* ctx, size_t need, cudaStream_t s)
{
	if (ctx->gacc_floats < need) {
		if (ctx->gacc) { GOF_CUDA_CHECK(cudaStreamSynchronize(s)); cudaFree(ctx->gacc); ctx->gacc = nullptr; ctx->gacc_floats = 0; }
		GOF_CUDA_CHECK(cudaMalloc(&ctx->gacc, need * sizeof(float)));
		ctx->gacc_floats = need;
	}
	return GOF_OK;
}

int gof_preprocess_backward(GofContext* ctx, const GofParams* prm, const GofInputs* in, const int32_t* radii,
                            const void* geom, const float* dL_dview2gaussian_in, const float* dL_dcolors_in,
                            const GofGrads* gr, gof_stream_t stream)
{
	if (!ctx || !prm || !in || !gr || !geom || !radii || !dL_dview2gaussian_in) { set_error("gof_preprocess_backward: NULL argument"); return GOF_EINVAL; }
	cudaStream_t s = (cudaStream_t)stream;
	const int P = prm->P;
	if (P <= 0) return GOF_OK;
	GeomState g = GeomState::carve(align_base(geom), (size_t)P, 1);
	int rc;
	if ((rc = ensure_gacc(ctx, (size_t)P * GACC_FLOATS, s)) != GOF_OK) return rc;
	pack_gacc_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, dL_dview2gaussian_in, dL_dcolors_in, ctx->gacc);
	GOF_CUDA_CHECK(cudaGetLastError());
	return launch_preprocess_bwd(*prm, *in, 1, g, radii, ctx->gacc, *gr, s);
}

// ---- test accessors --------------------------------------------------------------------------
int64_t gof_backward_accumulators(GofContext* ctx, void* dst, int64_t dst_bytes, gof_stream_t stream)
{
	if (!ctx) { set_error("gof_backward_accumulators: NULL context"); return GOF_EINVAL; }
	const int64_t bytes = (int64_t)(ctx->gacc_used * sizeof(float));
	if (dst && bytes) {
		if (bytes > dst_bytes) { set_error("gof_backward_accumulators: dst too small (%lld < %lld)", (long long)dst_bytes, (long long)bytes); return GOF_ENOMEM; }
		GOF_CUDA_CHECK(cudaMemcpyAsync(dst, ctx->gacc, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
	}
	return bytes;
}

int64_t gof_state_get_batch(const char* name, int32_t P, int32_t W, int32_t H, int32_t V, int64_t R,
                            const void* geom, const void* binning, size_t binning_bytes, const void* img,
                            void* dst, int64_t dst_bytes, gof_stream_t stream)
{
	if (!name) { set_error("gof_state_get: NULL name"); return GOF_EINVAL; }
	if (P < 0 || W <= 0 || H <= 0 || V <= 0 || R < 0) { set_error("gof_state_get: bad sizes"); return GOF_EINVAL; }
	cudaStream_t s = (cudaStream_t)stream;
	GofParams prm{};
	prm.P = P; prm.W = W; prm.H = H; prm.tan_fovx = prm.tan_fovy = 1.0f;
	const Frame f = make_frame(&prm, V);
	const size_t N = (size_t)W * H, n = (size_t)P * V, VT = (size_t)V * f.T;
	GeomState g = GeomState::carve(align_base(geom), (size_t)P, (size_t)V);
	ImgState im = ImgState::carve(align_base(img), N, (size_t)f.T, (size_t)V);
	const int64_t cap = blob_capacity(binning, binning_bytes, VT);
	if (cap < R) { set_error("gof_state_get: binning blob of %zu bytes holds %lld duplicates, num_rendered=%lld", binning_bytes, (long long)cap, (long long)R); return GOF_ENOMEM; }
	BinState b = BinState::carve(align_base(binning), (size_t)cap, VT);
	const std::string nm(name);
	const void* src = nullptr;
	size_t bytes = 0;
	bool kernel = false;
	if (nm == "depths") { src = g.depths; bytes = n * 4; }
	else if (nm == "means2D") { src = g.means2D; bytes = n * 8; }
	else if (nm == "conic_opacity") { src = g.conic_opacity; bytes = n * 16; }
	else if (nm == "clamped") { src = g.clamped; bytes = n * 3; }
	else if (nm == "tiles_touched") { src = g.tiles_touched; bytes = n * 4; }
	else if (nm == "final_T") { src = im.final_T; bytes = (size_t)V * N * 16; }
	else if (nm == "n_contrib") { src = im.n_contrib; bytes = (size_t)V * N * 8; }
	else if (nm == "ranges") { src = im.ranges; bytes = VT * 8; }
	else if (nm == "point_list") { src = b.point_list; bytes = (size_t)R * 4; }
	else if (nm == "point_offsets") { kernel = true; bytes = n * 4; }
	else if (nm == "point_list_keys") { kernel = true; bytes = (size_t)R * 8; }
	else if (nm == "view2gaussian") { kernel = true; bytes = n * 40; }
	else if (nm == "rgb") { kernel = true; bytes = n * 12; }
	else { set_error("gof_state_get: unknown array '%s'", name); return GOF_EINVAL; }
	if (dst && bytes) {
		if ((int64_t)bytes > dst_bytes) { set_error("gof_state_get: dst too small"); return GOF_ENOMEM; }
		if (kernel) {
			const char* what = nm == "point_offsets" ? "o" : nm == "point_list_keys" ? "k" : nm == "view2gaussian" ? "v" : "r";
			const int rc = launch_extract(what, f, g, im, b, R, dst, s);
			if (rc != GOF_OK) return rc;
		} else {
			GOF_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
		}
	}
	return (int64_t)bytes;
}

int64_t gof_state_get(const char* name, int32_t P, int32_t W, int32_t H, int64_t R,
                      const void* geom, const void* binning, size_t binning_bytes, const void* img,
                      void* dst, int64_t dst_bytes, gof_stream_t stream)
{
	return gof_state_get_batch(name, P, W, H, 1, R, geom, binning, binning_bytes, img, dst, dst_bytes, stream);
}

// ---- pack + all-gather over peer memory -----------------------------------------------------------
// The path's one exchange step: every rank needs the rgb / median depth / alpha of all frames.
// Instead of packing the five consumed channels and then calling an all-gather, ONE kernel reads this rank's
// raster output [F,9,N] and stores the packed [F,5,N] block straight into the gather buffer of EVERY rank over
// NVLink (peer pointers of a symmetric allocation), or into all of them at once through the NVSwitch multicast
// address (multimem.st) when the allocation has one.  A symmetric-memory barrier afterwards makes the frames visible.
__device__ __forceinline__ void multimem_st_v4(float* addr, float4 v)
{
	asm volatile("multimem.st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void pack_gather_kernel(const float* __restrict__ raster, int F, size_t N, const long long* __restrict__ peer_ptrs,
                                   int world, float* multicast, size_t dst_frame0)
{
	const int f = blockIdx.y;
	const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;   // N % 4 == 0 is checked by the launcher
	if (i4 >= N || f >= F) return;
	const int src_ch[5] = { 0, 1, 2, CH_DEPTH, CH_ALPHA };
#pragma unroll
	for (int c = 0; c < 5; c++) {
		const float4 v = *reinterpret_cast<const float4*>(raster + ((size_t)f * OUT_CH + src_ch[c]) * N + i4);
		const size_t off = ((dst_frame0 + f) * 5 + c) * N + i4;
		if (multicast != nullptr) {
			multimem_st_v4(multicast + off, v);
		} else {
			for (int r = 0; r < world; r++) *reinterpret_cast<float4*>(reinterpret_cast<float*>(peer_ptrs[r]) + off) = v;
		}
	}
}

int gof_pack_gather(const float* raster, int32_t frames, int64_t pixels, const int64_t* peer_ptrs_dev, int32_t world,
                    void* multicast_ptr, int64_t dst_frame0, gof_stream_t stream)
{
	if (!raster || frames <= 0 || pixels <= 0 || world <= 0 || (!peer_ptrs_dev && !multicast_ptr)) { set_error("gof_pack_gather: bad argument"); return GOF_EINVAL; }
	if (pixels % 4 != 0) { set_error("gof_pack_gather: pixel count must be a multiple of 4"); return GOF_EINVAL; }
	dim3 grid((unsigned)((pixels / 4 + 255) / 256), frames);
	pack_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(raster, frames, (size_t)pixels, (const long long*)peer_ptrs_dev, world,
	                                                          (float*)multicast_ptr, (size_t)dst_frame0);
	GOF_CUDA_CHECK(cudaGetLastError());
	return GOF_OK;
}

int gof_predictor_head(const GofHeadParams* prm, const float* net, const float* depth, const float* const_offset,
                       const float* ray_x, const float* ray_y, const float* view_to_world, const float* quat,
                       const float* sh_transform, float* xyz, float* opacity, float* scaling, float* rotation,
                       float* features_dc, float* features_rest, gof_stream_t stream)
{
	if (!prm) { set_error("gof_predictor_head: prm is NULL"); return GOF_EINVAL; }
	if (prm->BV < 0 || prm->H <= 0 || prm->W <= 0) { set_error("gof_predictor_head: bad sizes BV=%d H=%d W=%d", prm->BV, prm->H, prm->W); return GOF_EINVAL; }
	if (prm->sh_degree != 0 && prm->sh_degree != 1) { set_error("gof_predictor_head: only SH degree 0 or 1 (gaussian_predictor.py:993), got %d", prm->sh_degree); return GOF_EINVAL; }
	const int need = (prm->with_offset ? 3 : 0) + 11 + (prm->sh_degree > 0 ? 9 : 0);
	if (prm->C != need) { set_error("gof_predictor_head: the network output has %d channels, this configuration needs %d", prm->C, need); return GOF_EINVAL; }
	if (prm->BV == 0) return GOF_OK;
	if (!net || !depth || !ray_x || !ray_y || !view_to_world || !quat || !xyz || !opacity || !scaling || !rotation || !features_dc ||
	    (prm->sh_degree > 0 && !features_rest)) {
		set_error("gof_predictor_head: a required pointer is NULL");
		return GOF_EINVAL;
	}
	if (((uintptr_t)xyz | (uintptr_t)scaling | (uintptr_t)rotation | (uintptr_t)features_dc | (uintptr_t)features_rest) % 16 != 0) {
		set_error("gof_predictor_head: output arrays must be 16-byte aligned");
		return GOF_EINVAL;
	}
	return launch_predictor_head(*prm, net, depth, const_offset, ray_x, ray_y, view_to_world, quat, sh_transform, xyz, opacity,
	                             scaling, rotation, features_dc, features_rest, (cudaStream_t)stream);
}

}  // extern "C"
