// conic.cuh -- the tile-local conic pre-test of the blend kernels.
//
// For one (Gaussian, tile) duplicate the reference evaluates, at every pixel of the tile
// (forward.cu:502-535), mv = C - BB^2/(4 AA) from float32 AA = r^T Sigma r and BB = 2 b.r with the
// pixel ray r = (rx, ry, 1), and the pair can only contribute if mv <= tau (tau: record field,
// see blend_math.cuh).  With AA > 0 that is
//        f(r) = (b.r)^2 - (C - tau') (r^T Sigma r) >= 0,
// a CONIC in the ray, hence a quadratic polynomial in the tile-local pixel coordinates
// (x, y) in [0,15]^2 because rx = x/fx + const, ry = y/fy + const.  Its six coefficients are
// computed here once per duplicate, in double (the difference of the two ~1e9 terms is resolved
// before anything is rounded to float32), normalised by AA at the tile centre, and the blend
// evaluates g(x,y) = c0 + x (c1 + c3 x + c4 y) + y (c2 + c5 y) with five FMAs per pixel and skips
// the pair iff g < 0.
//
// Soundness (a pair the reference blends is never skipped).  Let mv_ref be the value the
// reference computes and mv* the exact value of the same expression at the exact affine ray.
//   (1) mv_ref <= tau is necessary for contributing (blend_math.cuh).
//   (2) |mv_ref - mv*| <= E, with E a forward error bound of the reference's float32 evaluation
//       (3 roundings per n_k, 3 more for AA, 4 for BB, 1 ulp for each ray component):
//           |dAA| <= 10u Abar,  |dBB| <= 5u Bbar,  u = 2^-24,
//           Abar = sum_k N_k rho_k, N_k = sum_j |S_kj| rho_j, Bbar = 2 sum_j |B_j| rho_j  (rho = max |ray| in the tile)
//           |dq|  <= (bmax dBB + dBB^2/4)/AAlo + Q' dAA/(AAlo - dAA),   Q' = (bmax + dBB/2)^2/AAlo
//       with AAlo / bmax rigorous bounds of AA / |b.r| over the tile's ray box (q = BB^2/(4AA));
//       E = 1.25x that.  tau' = tau + E, so mv_ref <= tau  =>  mv* <= tau'  =>  f >= 0.
//   (3) the float32 evaluation of g differs from its exact value by at most eta (5 FMAs on
//       coefficients rounded once, |x|,|y| <= 15), which is added to c0.
// Whenever a bound cannot be established (AAlo <= 0, relative perturbations not small, non-finite
// values) the coefficients are set to "never skip" (c0 = +big, others 0); NaNs make `g < 0` false.
#pragma once
#include "gof_common.cuh"

namespace gof {

struct TileRays {
	double ax, bx, ay, by;     // rx = ax * x + bx, ry = ay * y + by for tile-local pixel x, y in [0, 15]
	double rcx, rcy, hx, hy;   // centre and half-widths of the tile's ray box
	double rhox, rhoy;         // max |rx|, max |ry| over the tile
	double m;                  // max |x|, |y| the polynomial is evaluated at (15, or 15.5 with sub-pixel rays)
	double rel_extra;          // extra relative error of q in the consumer's arithmetic (0 for the blend)
};

// pad = 0: rays through pixel centres only (the blend).  pad = 0.5: any ray through the tile's pixels
// (point integration: corner rays and query points at sub-pixel positions; x, y in [-0.5, 15.5]); that
// consumer also forms BB/AA with a float32 division, one more rounding of q (rel_extra).
__device__ __forceinline__ TileRays tile_rays(int tx, int ty, int W, int H, float focal_x, float focal_y, double pad = 0.0)
{
	TileRays t;
	t.ax = 1.0 / (double)focal_x;
	t.ay = 1.0 / (double)focal_y;
	t.bx = ((double)(tx * TILE_X) + 0.5 - W / 2.) / (double)focal_x;
	t.by = ((double)(ty * TILE_Y) + 0.5 - H / 2.) / (double)focal_y;
	const double rxl = t.bx - pad * t.ax, ryl = t.by - pad * t.ay;
	const double rxh = t.bx + (TILE_X - 1 + pad) * t.ax, ryh = t.by + (TILE_Y - 1 + pad) * t.ay;
	t.rcx = 0.5 * (rxl + rxh);
	t.rcy = 0.5 * (ryl + ryh);
	t.hx = 0.5 * (rxh - rxl);
	t.hy = 0.5 * (ryh - ryl);
	t.rhox = fmax(fabs(rxl), fabs(rxh));
	t.rhoy = fmax(fabs(ryl), fabs(ryh));
	t.m = 15.0 + pad;
	t.rel_extra = pad > 0.0 ? 24.0 * 5.9604644775390625e-08 : 0.0;   // float32 BB/AA division; float32 AA t^2 + BB t + C (4 q-sized terms)
	return t;
}

__device__ __forceinline__ float conic_rcp(float x)
{
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));   // one MUFU.RCP, within 2^-22
	return r;
}

// q0 = (Sxx, Sxy, Sxz, Syy), q1 = (Syz, Szz, Bx, By), q2 = (Bz, C, tau, w)
__device__ __forceinline__ void conic_coefficients(const float4& q0, const float4& q1, const float4& q2,
                                                   const TileRays& t, float* c)
{
	const double u = 5.9604644775390625e-08;   // 2^-24
	const double Sxx = q0.x, Sxy = q0.y, Sxz = q0.z, Syy = q0.w, Syz = q1.x, Szz = q1.y;
	const double Bx = q1.z, By = q1.w, Bz = q2.x, C = q2.y, tau = q2.z;

	// AA at the tile centre and a lower bound over the tile's ray box
	const double n0c = Sxx * t.rcx + Sxy * t.rcy + Sxz;
	const double n1c = Sxy * t.rcx + Syy * t.rcy + Syz;
	const double n2c = Sxz * t.rcx + Syz * t.rcy + Szz;
	const double AAc = n0c * t.rcx + n1c * t.rcy + n2c;
	const double AAlo = AAc - 2.0 * (t.hx * fabs(n0c) + t.hy * fabs(n1c))
	                    - (fabs(Sxx) * t.hx * t.hx + 2.0 * fabs(Sxy) * t.hx * t.hy + fabs(Syy) * t.hy * t.hy);
	// magnitudes that scale the reference's float32 rounding errors
	const double N0 = fabs(Sxx) * t.rhox + fabs(Sxy) * t.rhoy + fabs(Sxz);
	const double N1 = fabs(Sxy) * t.rhox + fabs(Syy) * t.rhoy + fabs(Syz);
	const double N2 = fabs(Sxz) * t.rhox + fabs(Syz) * t.rhoy + fabs(Szz);
	const double Abar = N0 * t.rhox + N1 * t.rhoy + N2;
	const double Bbar = 2.0 * (fabs(Bx) * t.rhox + fabs(By) * t.rhoy + fabs(Bz));
	const double bmax = fabs(Bx * t.rcx + By * t.rcy + Bz) + t.hx * fabs(Bx) + t.hy * fabs(By);

	bool ok = (AAlo > 0.0) && (AAc > 0.0) && (tau < 1.0e37);
	const double dAA = 10.0 * u * Abar, dBB = 5.0 * u * Bbar;
	ok = ok && (dAA < 0.5 * AAlo);
	ok = ok && (AAlo > 1.0e-30) && (AAc < 1.0e30);          // float32 reciprocals below stay finite and normal
	const double AAl = ok ? AAlo : 1.0;
	// |q_ref - q*| for q = BB^2/(4 AA), |BB| <= 2 bmax, AA >= AAl, perturbed by (dBB, dAA): exact, not first order.
	// The bounds only have to be UPPER bounds, so the divisions are one float32 reciprocal of AAl rounded up by
	// 2^-20 (MUFU.RCP is within 2^-22) and 1/(1-x) <= 1 + 2x for x = dAA/AAl < 0.5.
	const double rAAl = (double)conic_rcp((float)AAl) * (1.0 + 9.5367431640625e-07);
	const double E_bb = (bmax * dBB + 0.25 * dBB * dBB) * rAAl;
	const double Qp = (bmax + 0.5 * dBB) * (bmax + 0.5 * dBB) * rAAl;
	const double xa = (ok ? dAA : 0.0) * rAAl;
	const double E_aa = Qp * xa * (1.0 + 2.0 * xa);
	const double E = 1.25 * (E_bb + E_aa) + t.rel_extra * Qp;
	const double K = C - (tau + E);

	// polynomial expansion in tile-local pixel coordinates
	const double p = Bx * t.ax, q = By * t.ay, r0 = Bx * t.bx + By * t.by + Bz;
	const double a_xx = Sxx * t.ax * t.ax, a_xy = 2.0 * Sxy * t.ax * t.ay, a_yy = Syy * t.ay * t.ay;
	const double a_x = 2.0 * t.ax * (Sxx * t.bx + Sxy * t.by + Sxz);
	const double a_y = 2.0 * t.ay * (Syy * t.by + Sxy * t.bx + Syz);
	const double a_0 = Sxx * t.bx * t.bx + 2.0 * Sxy * t.bx * t.by + Syy * t.by * t.by + 2.0 * Sxz * t.bx + 2.0 * Syz * t.by + Szz;
	// any positive common scale leaves the sign of g unchanged: 1/AAc only normalises the magnitudes for float32
	const double inv = (double)conic_rcp((float)(ok ? AAc : 1.0));
	const double c0 = (r0 * r0 - K * a_0) * inv;
	const double c1 = (2.0 * p * r0 - K * a_x) * inv;
	const double c2 = (2.0 * q * r0 - K * a_y) * inv;
	const double c3 = (p * p - K * a_xx) * inv;
	const double c4 = (2.0 * p * q - K * a_xy) * inv;
	const double c5 = (q * q - K * a_yy) * inv;
	const double m = t.m;
	// float32 evaluation error of g (+ the double rounding of the expansion itself, which works on
	// terms of magnitude (r0^2 + |K| a_0)/AAc and their x, y analogues)
	const double eta = 8.0 * u * (fabs(c0) + m * (fabs(c1) + fabs(c2)) + m * m * (fabs(c3) + fabs(c4) + fabs(c5)))
	                 + 1.0e-13 * inv * ((r0 * r0 + fabs(K) * fabs(a_0)) + m * (fabs(2.0 * p * r0) + fabs(K * a_x) + fabs(2.0 * q * r0) + fabs(K * a_y))
	                                    + m * m * (p * p + q * q + fabs(2.0 * p * q) + fabs(K) * (fabs(a_xx) + fabs(a_xy) + fabs(a_yy))));
	const double c0e = c0 + eta;
	const double big = 1.0e30;
	ok = ok && (fabs(c0e) < big) && (fabs(c1) < big) && (fabs(c2) < big) && (fabs(c3) < big) && (fabs(c4) < big) && (fabs(c5) < big);
	if (tau < -1.0e37) {
		// w < 1/255: alpha <= w can never reach 1/255 (power <= 0) -- always skip
		c[0] = -1.0f;
		c[1] = c[2] = c[3] = c[4] = c[5] = 0.0f;
	} else if (ok) {
		// round the constant term up, so that rounding to float32 cannot lower it
		c[0] = __double2float_ru(c0e);
		c[1] = (float)c1; c[2] = (float)c2; c[3] = (float)c3; c[4] = (float)c4; c[5] = (float)c5;
	} else {
		c[0] = 3.0e38f;
		c[1] = c[2] = c[3] = c[4] = c[5] = 0.0f;
	}
}

// Which of the tile's eight 8x4 pixel blocks (block b: x in [8(b&1), 8(b&1)+7], y in [4(b>>1), 4(b>>1)+3], the
// pixels of consumer warp b of the blend kernels) can contain a pixel that passes the conic test.  For a block
// with centre (cx, cy) and half-widths (hx, hy), any quadratic obeys
//     g(x, y) <= g(c) + |g_x(c)| hx + |g_y(c)| hy + max(c3,0) hx^2 + |c4| hx hy + max(c5,0) hy^2 =: U,
// and the blend's float32 evaluation exceeds the exact polynomial by at most eta (see conic_coefficients), so
// U + eta < 0 proves that the whole block rejects the record.  `pad` widens the block for sub-pixel rays.
__device__ __forceinline__ uint32_t conic_block_mask(const float* c, double pad_d)
{
	// Evaluated in float32: U is built from ~30 roundings of partial sums that are all below 3 * mag, i.e. within
	// 90 u mag of its exact value; together with the blend's own evaluation error (16 u mag) the margin is 144 u mag.
	const float c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], c4 = c[4], c5 = c[5];
	if (!(c0 < 1.0e30f)) return 0xffu;                 // "never skip" records (and NaNs) stay relevant everywhere
	const float pad = (float)pad_d;
	const float m = 15.0f + pad;
	const float mag = fabsf(c0) + m * (fabsf(c1) + fabsf(c2)) + m * m * (fabsf(c3) + fabsf(c4) + fabsf(c5));
	const float eta = 144.0f * 5.9604644775390625e-08f * mag;
	const float hx = 3.5f + pad, hy = 1.5f + pad;
	const float quad = fmaxf(c3, 0.0f) * hx * hx + fabsf(c4) * hx * hy + fmaxf(c5, 0.0f) * hy * hy + eta;
	uint32_t mask = 0;
#pragma unroll
	for (int b = 0; b < 8; b++) {
		const float cx = 8.0f * (b & 1) + 3.5f, cy = 4.0f * (b >> 1) + 1.5f;
		const float gc = c0 + cx * (c1 + c3 * cx + c4 * cy) + cy * (c2 + c5 * cy);
		const float gx = c1 + 2.0f * c3 * cx + c4 * cy, gy = c2 + c4 * cx + 2.0f * c5 * cy;
		const float U = gc + fabsf(gx) * hx + fabsf(gy) * hy + quad;
		if (!(U < 0.0f)) mask |= 1u << b;
	}
	return mask;
}

// g(x, y) < 0  =>  the reference skips the pair.
__device__ __forceinline__ bool conic_reject(float c0, float c1, float c2, float c3, float c4, float c5, float x, float y)
{
	const float t1 = __fmaf_rn(c3, x, __fmaf_rn(c4, y, c1));
	const float t2 = __fmaf_rn(c5, y, c2);
	const float g = __fmaf_rn(y, t2, __fmaf_rn(x, t1, c0));
	return g < 0.0f;
}

}  // namespace gof
