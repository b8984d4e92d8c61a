// predictor_head.cu -- the predictor's OUTPUT HEAD fused into one kernel.
//
// Replaces everything GaussianSplatPredictor_gtunet.forward does after the UNet
// (src/gaussian_predictor.py:954-1008, ~25 elementwise / permute / bmm / cat launches plus make_contiguous):
// channel split, ray * depth + offset (:857-881), row-vector view_to_world transform and homogeneous divide
// (:959-966), squre_clip (:968-970), sigmoid / exp / normalize (:636-638,:975-977), rotation to world
// (quaternion_raw_multiply, :45-63,:839-855), degree-1 SH rotation (:821-837), and the NCHW -> [B, V*N, .]
// point-list layout (flatten_vector :788-791, multi_view_union :796-800) that the rasterizer reads.
//
// One thread per pixel.  Reads: the C channel planes of the network output and the depth plane (coalesced
// 128-byte rows per warp and channel).  Writes: six point-list arrays; the 3- and 9-float records are staged
// through shared memory and stored as 16-byte vectors so every store instruction covers a contiguous run.
// Pure HBM streaming: (C + 1) * 4 bytes in, (14 + 9 * sh) * 4 bytes out per pixel.
//
// Arithmetic follows torch's float32 operation sequence (separate multiply / add roundings where torch runs
// separate kernels); expf / IEEE division like torch's CUDA sigmoid / exp / div kernels.
#include "gof_common.cuh"

namespace gof {
namespace {

constexpr int HEAD_THREADS = 256;

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

template <bool SH>
__global__ void __launch_bounds__(HEAD_THREADS)
predictor_head_kernel(GofHeadParams prm, const float* __restrict__ net, const float* __restrict__ depth,
                      const float* __restrict__ const_offset, const float* __restrict__ ray_x,
                      const float* __restrict__ ray_y, const float* __restrict__ view_to_world,
                      const float* __restrict__ quat, const float* __restrict__ sh_transform,
                      float* __restrict__ xyz, float* __restrict__ opacity, float* __restrict__ scaling,
                      float* __restrict__ rotation, float* __restrict__ features_dc, float* __restrict__ features_rest)
{
	__shared__ __align__(16) float s_xyz[HEAD_THREADS * 3];
	__shared__ __align__(16) float s_scl[HEAD_THREADS * 3];
	__shared__ __align__(16) float s_dc[HEAD_THREADS * 3];
	__shared__ __align__(16) float s_rest[SH ? HEAD_THREADS * 9 : 4];

	const int t = threadIdx.x;
	const size_t N = (size_t)prm.H * prm.W;
	const size_t total = (size_t)prm.BV * N;
	const size_t g0 = (size_t)blockIdx.x * HEAD_THREADS;
	const size_t g = g0 + t;
	const bool live = g < total;

	float q_out[4] = {0.f, 0.f, 0.f, 0.f};
	float op = 0.f;
	if (live) {
		const int bv = (int)(g / N);
		const size_t pix = g - (size_t)bv * N;
		const int py = (int)(pix / prm.W), px = (int)(pix - (size_t)py * prm.W);
		const float* ch = net + (size_t)bv * prm.C * N + pix;     // channel c of this pixel: ch[c * N]
		int at = 0;
		float off[3] = {0.f, 0.f, 0.f};
		if (prm.with_offset) {
#pragma unroll
			for (int c = 0; c < 3; c++) off[c] = __ldg(ch + (size_t)c * N);
			at = 3;
		}
		const float raw_op = __ldg(ch + (size_t)at * N);
		float raw_s[3], raw_q[4], dc[3];
#pragma unroll
		for (int c = 0; c < 3; c++) raw_s[c] = __ldg(ch + (size_t)(at + 1 + (prm.isotropic ? 0 : c)) * N);
#pragma unroll
		for (int c = 0; c < 4; c++) raw_q[c] = __ldg(ch + (size_t)(at + 4 + c) * N);
#pragma unroll
		for (int c = 0; c < 3; c++) dc[c] = __ldg(ch + (size_t)(at + 8 + c) * N);
		float sh[9];
		if (SH) {
#pragma unroll
			for (int c = 0; c < 9; c++) sh[c] = __ldg(ch + (size_t)(at + 11 + c) * N);
		}
		float d = __ldg(depth + g);
		if (const_offset) d = __fadd_rn(d, __ldg(const_offset + g));          // :872

		// position: ray * depth + offset (:878), then [p,1] @ view_to_world and the homogeneous divide (:959-966)
		const float pcx = __fadd_rn(__fmul_rn(__ldg(ray_x + px), d), off[0]);
		const float pcy = __fadd_rn(__fmul_rn(__ldg(ray_y + py), d), off[1]);
		const float pcz = __fadd_rn(d, off[2]);                                 // ray z is exactly 1
		const float* M = view_to_world + (size_t)bv * 16;
		float hom[4];
#pragma unroll
		for (int j = 0; j < 4; j++)
			hom[j] = fmaf(pcx, __ldg(M + j), fmaf(pcy, __ldg(M + 4 + j), fmaf(pcz, __ldg(M + 8 + j), __ldg(M + 12 + j))));
		const float w = __fadd_rn(hom[3], 1e-10f);
		float p[3];
#pragma unroll
		for (int j = 0; j < 3; j++) p[j] = __fdiv_rn(hom[j], w);
		if (prm.squre_clip < 10.0f) {                                           // :968-970
			p[0] = clampf(p[0], -prm.squre_clip, prm.squre_clip);
			p[1] = clampf(p[1], -prm.squre_clip, prm.squre_clip);
		}
#pragma unroll
		for (int j = 0; j < 3; j++) s_xyz[t * 3 + j] = p[j];

		op = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-raw_op)));                    // torch.sigmoid
#pragma unroll
		for (int j = 0; j < 3; j++) s_scl[t * 3 + j] = expf(raw_s[j]);           // torch.exp
#pragma unroll
		for (int j = 0; j < 3; j++) s_dc[t * 3 + j] = dc[j];

		// F.normalize(dim=1): x / max(||x||_2, 1e-12), then Mq (x) q with quaternion_raw_multiply (:45-63)
		const float n2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(raw_q[0], raw_q[0]), __fmul_rn(raw_q[1], raw_q[1])),
		                                     __fmul_rn(raw_q[2], raw_q[2])), __fmul_rn(raw_q[3], raw_q[3]));
		const float nrm = fmaxf(__fsqrt_rn(n2), 1e-12f);
		const float bw = __fdiv_rn(raw_q[0], nrm), bx = __fdiv_rn(raw_q[1], nrm), by = __fdiv_rn(raw_q[2], nrm), bz = __fdiv_rn(raw_q[3], nrm);
		const float* A = quat + (size_t)bv * 4;
		const float aw = __ldg(A), ax = __ldg(A + 1), ay = __ldg(A + 2), az = __ldg(A + 3);
#define GOF_M(a, b) __fmul_rn((a), (b))
		q_out[0] = __fsub_rn(__fsub_rn(__fsub_rn(GOF_M(aw, bw), GOF_M(ax, bx)), GOF_M(ay, by)), GOF_M(az, bz));
		q_out[1] = __fsub_rn(__fadd_rn(__fadd_rn(GOF_M(aw, bx), GOF_M(ax, bw)), GOF_M(ay, bz)), GOF_M(az, by));
		q_out[2] = __fadd_rn(__fadd_rn(__fsub_rn(GOF_M(aw, by), GOF_M(ax, bz)), GOF_M(ay, bw)), GOF_M(az, bx));
		q_out[3] = __fadd_rn(__fsub_rn(__fadd_rn(GOF_M(aw, bz), GOF_M(ax, by)), GOF_M(ay, bx)), GOF_M(az, bw));
#undef GOF_M

		if (SH) {
			// transform_SHs (:821-837): out[sh', rgb] = sum_sh in[sh, rgb] * T[sh, sh'], in[sh, rgb] = channel 3*sh + rgb
			float Tm[9];
			if (sh_transform) {
				const float* T = sh_transform + (size_t)bv * 9;
#pragma unroll
				for (int k = 0; k < 9; k++) Tm[k] = __ldg(T + k);
			} else {
				// the module's own basis change (init_sh_transform_matrices, :649-655): v_to_sh = [[0,0,-1],[-1,0,0],[0,1,0]],
				// sh_to_v its transpose => T[i][l] = a_i * R[j_i][j_l] * a_l with a = (-1,+1,-1), j = (1,2,0); exact (signs only)
#pragma unroll
				for (int i = 0; i < 3; i++)
#pragma unroll
					for (int l = 0; l < 3; l++) {
						const float r = __ldg(M + ((i + 1) % 3) * 4 + ((l + 1) % 3));
						Tm[i * 3 + l] = ((i == 1) != (l == 1)) ? -r : r;
					}
			}
#pragma unroll
			for (int so = 0; so < 3; so++)
#pragma unroll
				for (int rgb = 0; rgb < 3; rgb++)
					s_rest[t * 9 + so * 3 + rgb] = fmaf(sh[6 + rgb], Tm[6 + so], fmaf(sh[3 + rgb], Tm[3 + so], __fmul_rn(sh[rgb], Tm[so])));
		}
	}
	__syncthreads();

	if (live) {
		opacity[g] = op;
		reinterpret_cast<float4*>(rotation)[g] = make_float4(q_out[0], q_out[1], q_out[2], q_out[3]);
	}
	if (g0 + HEAD_THREADS <= total) {
		// full block: the block's records are one contiguous, 16-byte aligned run per array
		float4* dx = reinterpret_cast<float4*>(xyz + g0 * 3);
		float4* ds = reinterpret_cast<float4*>(scaling + g0 * 3);
		float4* dd = reinterpret_cast<float4*>(features_dc + g0 * 3);
		if (t < HEAD_THREADS * 3 / 4) {
			dx[t] = reinterpret_cast<const float4*>(s_xyz)[t];
			ds[t] = reinterpret_cast<const float4*>(s_scl)[t];
			dd[t] = reinterpret_cast<const float4*>(s_dc)[t];
		}
		if (SH) {
			float4* dr = reinterpret_cast<float4*>(features_rest + g0 * 9);
			for (int i = t; i < HEAD_THREADS * 9 / 4; i += HEAD_THREADS) dr[i] = reinterpret_cast<const float4*>(s_rest)[i];
		}
	} else if (live) {
#pragma unroll
		for (int j = 0; j < 3; j++) {
			xyz[g * 3 + j] = s_xyz[t * 3 + j];
			scaling[g * 3 + j] = s_scl[t * 3 + j];
			features_dc[g * 3 + j] = s_dc[t * 3 + j];
		}
		if (SH) {
#pragma unroll
			for (int j = 0; j < 9; j++) features_rest[g * 9 + j] = s_rest[t * 9 + j];
		}
	}
}

}  // namespace

int launch_predictor_head(const GofHeadParams& prm, const float* net, const float* depth, const float* const_offset,
                          const float* ray_x, const float* ray_y, const float* view_to_world, const float* quat,
                          const float* sh_transform, float* xyz, float* opacity, float* scaling, float* rotation,
                          float* features_dc, float* features_rest, cudaStream_t s)
{
	const size_t total = (size_t)prm.BV * prm.H * prm.W;
	const unsigned blocks = (unsigned)((total + HEAD_THREADS - 1) / HEAD_THREADS);
	if (blocks == 0) return GOF_OK;
	if (prm.sh_degree > 0)
		predictor_head_kernel<true><<<blocks, HEAD_THREADS, 0, s>>>(prm, net, depth, const_offset, ray_x, ray_y, view_to_world, quat,
		                                                             sh_transform, xyz, opacity, scaling, rotation, features_dc, features_rest);
	else
		predictor_head_kernel<false><<<blocks, HEAD_THREADS, 0, s>>>(prm, net, depth, const_offset, ray_x, ray_y, view_to_world, quat,
		                                                              sh_transform, xyz, opacity, scaling, rotation, features_dc, features_rest);
	GOF_CUDA_CHECK(cudaGetLastError());
	return GOF_OK;
}

}  // namespace gof
