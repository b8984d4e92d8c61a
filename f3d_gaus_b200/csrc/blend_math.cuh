// blend_math.cuh -- device helpers shared by the forward and backward tile blends:
//   * mbarrier + TMA bulk-copy (cp.async.bulk, SASS UBLKCP) primitives for the slab pipeline;
//   * the exact GOF ray-minimum evaluation of one (pixel, Gaussian) pair (the cheap conservative
//     pre-test that precedes it is the tile-local conic of conic.cuh).
//
// Arithmetic contract.  The reference evaluates, per pair (forward.cu:502-535):
//     n  = Sigma_v * (rx, ry, 1)            float32
//     AA = (rx, ry, 1) . n                  float32, then widened to double
//     BB = 2 * (B . (rx, ry, 1))            float32, then widened to double
//     t  = float(-BB / (2 AA)),  skip if t <= 0.2
//     mv = -(BB/AA) * (BB/4) + C            double
//     power = min(0, float(-0.5 * mv)),  alpha = min(0.99, w * expf(power)),  skip if < 1/255
// mv is a difference of two ~(t/s)^2 ~ 6e5 terms, so one ulp of the float32 n/AA/BB moves
// alpha by percents: those float32 values must be reproduced to the bit.  They are written
// with explicit round-to-nearest intrinsics in the association the reference's build uses
// (n_k = S_kz + fma(S_kx, rx, S_ky*ry) etc.; established from its sm_100a SASS), so the
// result does not depend on this compiler's contraction choices.
//
// The contribution threshold.  alpha >= 1/255 needs  w*exp(-mv/2) >= 1/255, i.e.
// mv <= 2 ln(255 w) =: tau0.  The preprocess stores tau = tau0 * 1.00001 + 2e-3 per Gaussian
// (-FLT_MAX when w < 1/255, since power <= 0 caps alpha at w; +FLT_MAX for a NaN opacity): a pair
// whose ray minimum exceeds tau is skipped by the reference too -- by its alpha test or by its t
// test, neither of which has a side effect other than `continue`.  The conic pre-test is built
// from this tau plus a bound of the reference's own float32 evaluation error.
#pragma once
#include "gof_common.cuh"

namespace gof {

// ---------------------------------------------------------------- async-copy primitives --
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
	return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
	uint32_t ok;
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		"selp.u32 %0, 1, 0, p;\n\t}"
		: "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
	while (!mbar_try_wait(bar, parity)) {}
}
// Producer-side wait: the elected lane polls rarely so that it does not take issue slots from the
// consumer warps of its scheduler.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity)
{
	while (!mbar_try_wait(bar, parity)) __nanosleep(1000);
}
// Shared-memory loads by 32-bit shared-window address (no generic-address arithmetic in the loops).
__device__ __forceinline__ float4 lds128(uint32_t addr)
{
	float4 v;
	asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
	return v;
}
__device__ __forceinline__ float2 lds64(uint32_t addr)
{
	float2 v;
	asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
	return v;
}
__device__ __forceinline__ float lds32(uint32_t addr)
{
	float v;
	asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
	return v;
}
// TMA bulk copy global -> shared, completion signalled on `bar` (bytes % 16 == 0, 16-B aligned).
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------- pixel rays ----
// ray = ((px + 0.5 - W/2.) / fx, (py + 0.5 - H/2.) / fy): numerator and quotient in double,
// narrowed to float (forward.cu:440,448).
__device__ __forceinline__ float pixel_ray(uint32_t p, int S, float focal)
{
	const float pf = (float)p + 0.5f;
	return (pf - S / 2.) / focal;
}

// ------------------------------------------------------------------ pair evaluation -----
struct PairGeom {
	float n0, n1, n2;   // Sigma_v * ray
	float AA, BB;       // float32 quadratic coefficients
};

// k1 = (c4, c5, w, Sxx)  k2 = (Sxy, Sxz, Syy, Syz)  k3 = (Szz, Bx, By, Bz)   (slab record, gof_common.cuh)
__device__ __forceinline__ PairGeom pair_geom(const float4& k1, const float4& k2, const float4& k3, float rx, float ry)
{
	const float Sxx = k1.w, Sxy = k2.x, Sxz = k2.y, Syy = k2.z, Syz = k2.w, Szz = k3.x, Bx = k3.y, By = k3.z, Bz = k3.w;
	PairGeom g;
	g.n0 = __fadd_rn(Sxz, __fmaf_rn(Sxx, rx, __fmul_rn(Sxy, ry)));
	g.n1 = __fadd_rn(Syz, __fmaf_rn(Sxy, rx, __fmul_rn(Syy, ry)));
	g.n2 = __fadd_rn(Szz, __fmaf_rn(Syz, ry, __fmul_rn(Sxz, rx)));
	g.AA = __fadd_rn(__fmaf_rn(g.n0, rx, __fmul_rn(g.n1, ry)), g.n2);
	const float bb = __fadd_rn(Bz, __fmaf_rn(Bx, rx, __fmul_rn(By, ry)));
	g.BB = __fadd_rn(bb, bb);
	return g;
}

__device__ __forceinline__ float rcp_approx(float x)
{
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));   // one MUFU.RCP, <= 1 ulp
	return r;
}

__device__ __forceinline__ float rsqrt_approx(float x)
{
	float r;
	asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));  // one MUFU.RSQ, no denormal fix-up (x >= 1e-7 here)
	return r;
}

// Branch-free form of pair_alpha_exact for callers that evaluate two independent pairs per iteration (the
// two dependency chains -- double division, exp -- then interleave): same results, validity returned.
__device__ __forceinline__ bool pair_alpha_eval(const PairGeom& g, float C, float w, float& t, float& alpha)
{
	const double AA = g.AA;
	const double BB = g.BB;
	const double u = (-BB) / AA;
	t = (float)(u * 0.5);
	const double mv = fma(u, BB * 0.25, (double)C);
	float power = (float)(mv * -0.5);
	power = (power > 0.0f) ? 0.0f : power;
	alpha = min(0.99f, __fmul_rn(w, expf(power)));
	return !(t <= __uint_as_float(0x3E4CCCCCu)) && !(alpha < 1.0f / 255.0f);
}

// Exact alpha of the pair.  Returns false if the reference `continue`s (t <= near plane or
// alpha < 1/255).  Outputs t, alpha and G = exp(power) (the backward needs G).
__device__ __forceinline__ bool pair_alpha_exact(const PairGeom& g, float C, float w, float& t, float& alpha, float& G, double& u)
{
	const double AA = g.AA;
	const double BB = g.BB;
	// -BB/(2AA) == (-BB/AA) * 0.5 exactly (power-of-two scaling), so one IEEE division serves
	// both t and the BB/AA factor of the ray minimum.
	u = (-BB) / AA;
	t = (float)(u * 0.5);
	// The reference compares the float t with the DOUBLE constant 0.2 (NEAR_PLANE, auxiliary.h:26).  0.2 is not a
	// float: the floats around it are 0.19999998807907104 (0x3E4CCCCC) and 0.20000000298023224 (0x3E4CCCCD), so
	// for a float t:  t <= 0.2 (double)  <=>  t <= 0x3E4CCCCC.  NaN fails both forms alike.
	if (t <= __uint_as_float(0x3E4CCCCCu)) return false;
	const double mv = fma(u, BB * 0.25, (double)C);
	float power = (float)(mv * -0.5);
	if (power > 0.0f) power = 0.0f;
	G = expf(power);
	alpha = min(0.99f, __fmul_rn(w, G));
	if (alpha < 1.0f / 255.0f) return false;
	return true;
}
__device__ __forceinline__ bool pair_alpha_exact(const PairGeom& g, float C, float w, float& t, float& alpha, float& G)
{
	double u;
	return pair_alpha_exact(g, C, w, t, alpha, G, u);
}

}  // namespace gof
