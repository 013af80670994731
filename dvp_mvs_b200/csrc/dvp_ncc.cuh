// dvp_ncc.cuh — bilateral-NCC photometric cost with the reference-side work hoisted per pixel.
//
// Replaces ComputeHomography / ComputeCorrespondingPoint / ComputeBilateralWeight / ComputeBilateralNCCOld
// (reference APD.cu:679-748, 776-781, 1023-1113).
//
// What changed and why it is result-identical:
//  * The reference recomputes, for every (pixel, hypothesis, view), the 36 bilateral weights w_k, the 36
//    reference samples r_k and the three reference-side sums.  They depend only on (pixel, radius).  Here
//    they are computed ONCE per pixel per kernel (RefPatch::prepare), kept in shared memory as
//    (w_k, w_k*r_k) pairs and reused by every NCC of that pixel (23*S per propagation sweep, 62*S in the
//    cost-profile pass).  The float sequence per accumulator is unchanged: row sums, then patch sums,
//    in the reference's loop order (outer loop = x offset, inner = y offset), with the same
//    multiply/FMA split the reference's SASS shows (t = w*r rounded once; sum_ref += t; sum_ref_ref =
//    fma(r, t, .); sum_ref_src = fma(s, t, .); u = w*s; sum_src += u; sum_src_src = fma(s, u, .)).
//  * Reference-image samples are taken at integer texel centres, where the texture unit returns the texel
//    exactly (clamp addressing), so they are read from a linear copy of image 0 instead of through TEX.
//    Source-image samples keep going through the texture unit (hardware bilinear, 8-bit fractions) at the
//    same coordinates fma(x, 1/z, 0.5).
//  * Per-view constants of the homography are precomputed (ViewConst).
//  * If the reference patch variance is below kMinVar every NCC of that pixel is cost_max = 2.0
//    (APD.cu:1101-1104), so sampling is skipped altogether.
#pragma once
#include "dvp_common.cuh"

namespace dvp {

constexpr float kCostMax = 2.0f;
constexpr float kMinVar = 1e-5f;

// sqrt.approx(i*i + j*j) of the 36 offsets of the radius-5 / increment-2 patch, filled on the device with the
// very instruction RefPatch::weight uses (launch_fill_sd_table); read by the "recompute w" form of the NCC.
static __constant__ float c_sd_r5[kHoistSamples];

// reference APD.cu:709-738 with R_relative / t_relative hoisted into ViewConst.
__device__ __forceinline__ void compute_homography(const dvp_camera& ref, const ViewConst& vc, const float4 pl, float* H) {
	H[0] = vc.R_rel[0] - vc.t_rel[0] * pl.x / pl.w;
	H[1] = vc.R_rel[1] - vc.t_rel[0] * pl.y / pl.w;
	H[2] = vc.R_rel[2] - vc.t_rel[0] * pl.z / pl.w;
	H[3] = vc.R_rel[3] - vc.t_rel[1] * pl.x / pl.w;
	H[4] = vc.R_rel[4] - vc.t_rel[1] * pl.y / pl.w;
	H[5] = vc.R_rel[5] - vc.t_rel[1] * pl.z / pl.w;
	H[6] = vc.R_rel[6] - vc.t_rel[2] * pl.x / pl.w;
	H[7] = vc.R_rel[7] - vc.t_rel[2] * pl.y / pl.w;
	H[8] = vc.R_rel[8] - vc.t_rel[2] * pl.z / pl.w;

	float tmp[9];
	tmp[0] = H[0] / ref.K[0];
	tmp[1] = H[1] / ref.K[4];
	tmp[2] = -H[0] * ref.K[2] / ref.K[0] - H[1] * ref.K[5] / ref.K[4] + H[2];
	tmp[3] = H[3] / ref.K[0];
	tmp[4] = H[4] / ref.K[4];
	tmp[5] = -H[3] * ref.K[2] / ref.K[0] - H[4] * ref.K[5] / ref.K[4] + H[5];
	tmp[6] = H[6] / ref.K[0];
	tmp[7] = H[7] / ref.K[4];
	tmp[8] = -H[6] * ref.K[2] / ref.K[0] - H[7] * ref.K[5] / ref.K[4] + H[8];

	H[0] = vc.sK[0] * tmp[0] + vc.sK[2] * tmp[6];
	H[1] = vc.sK[0] * tmp[1] + vc.sK[2] * tmp[7];
	H[2] = vc.sK[0] * tmp[2] + vc.sK[2] * tmp[8];
	H[3] = vc.sK[4] * tmp[3] + vc.sK[5] * tmp[6];
	H[4] = vc.sK[4] * tmp[4] + vc.sK[5] * tmp[7];
	H[5] = vc.sK[4] * tmp[5] + vc.sK[5] * tmp[8];
	H[6] = vc.sK[8] * tmp[6];
	H[7] = vc.sK[8] * tmp[7];
	H[8] = vc.sK[8] * tmp[8];
}

// Per-pixel hoisted reference-side patch state.  The 36 (w, w*r) pairs live in shared memory at
// wt[k * stride + lane_slot]; the normalised reference moments live in registers.
struct RefPatch {
	int radius, inc, n;   // samples per axis n = number of i in [-radius, radius] step inc
	float inv_w;          // 1 / sum(w)      (MUFU.RCP, as the reference)
	float mean_ref;       // inv_w * sum(w r)
	float var_ref;        // fma(inv_w, sum(w r r), -mean_ref^2)
	bool degenerate;      // var_ref < kMinVar  -> every cost is kCostMax
	bool hoisted;         // n == kHoistAxis: shared-memory table valid; otherwise the slow path recomputes weights
	// "recompute w" form (RW): the table holds only the 36 reference samples r; w = weight(...) is re-evaluated
	// per use (one ex2) from these.  Halves the table, which hands the freed shared memory to the L1/TEX cache.
	float center, rcp_s, rcp_c;
	bool sd_const;        // radius 5, increment 2: spatial distances come from c_sd_r5

	__device__ __forceinline__ static float ref_pixel(const KArgs& a, int x, int y) {
		// texture clamp addressing at texel centres == clamped integer read
		x = min(max(x, 0), a.W - 1);
		y = min(max(y, 0), a.H - 1);
		return __ldg(a.ref_img + (size_t)y * a.W + x);
	}
	// reference APD.cu:776-781 as compiled: ex2(log2e * fma(-sqrt(i*i + j*j), 1/(2 ss^2), -(|r - rc| * 1/(2 sc^2))))
	__device__ __forceinline__ static float spatial_dist(int i, int j) {
		const float fi = (float)i, fj = (float)j;
		return sqrt_approx(__fmaf_rn(fj, fj, __fmul_rn(fi, fi)));
	}
	__device__ __forceinline__ static float weight_sd(float sd, float pix, float center, float rcp_s, float rcp_c) {
		const float cd = __fmul_rn(fabsf(__fadd_rn(pix, -center)), rcp_c);
		const float e = __fmaf_rn(-sd, rcp_s, -cd);
		return ex2_approx(__fmul_rn(e, 1.4426950216293334961f));
	}
	__device__ __forceinline__ static float weight(int i, int j, float pix, float center, float rcp_s, float rcp_c) {
		return weight_sd(spatial_dist(i, j), pix, center, rcp_s, rcp_c);
	}
	__device__ __forceinline__ static void sigma_rcps(const dvp_params& p, float& rcp_s, float& rcp_c) {
		rcp_s = rcp_approx(__fmul_rn(p.sigma_spatial, __fadd_rn(p.sigma_spatial, p.sigma_spatial)));
		rcp_c = rcp_approx(__fmul_rn(p.sigma_color, __fadd_rn(p.sigma_color, p.sigma_color)));
	}

	template <bool RW = false>
	__device__ __forceinline__ void prepare(const KArgs& a, int px, int py, int rad, float2* wt, int stride) {
		radius = rad;
		inc = a.prm.strong_increment;
		if (a.prm.use_radius) inc = DVP_MAX(2, (int)(2.0 * rad / 5.0));
		n = (rad >= 0) ? (2 * rad) / inc + 1 : 0;
		hoisted = (n == kHoistAxis) && wt != nullptr;   // no table given: every NCC recomputes its weights (general path)
		sigma_rcps(a.prm, rcp_s, rcp_c);
		center = ref_pixel(a, px, py);
		sd_const = (rad == 5 && inc == 2);
		float s_w = 0.f, s_r = 0.f, s_rr = 0.f;
		int k = 0;
		for (int i = -rad; i <= rad; i += inc) {
			float r_w = 0.f, r_r = 0.f, r_rr = 0.f;
			for (int j = -rad; j <= rad; j += inc) {
				const float pix = ref_pixel(a, px + i, py + j);
				const float w = weight(i, j, pix, center, rcp_s, rcp_c);
				const float t = __fmul_rn(pix, w);
				r_w = __fadd_rn(w, r_w);
				r_r = __fadd_rn(t, r_r);
				r_rr = __fmaf_rn(pix, t, r_rr);
				if (hoisted) {
					if (RW) reinterpret_cast<float*>(wt)[k * stride] = pix;
					else wt[k * stride] = make_float2(w, t);
				}
				++k;
			}
			s_w = __fadd_rn(r_w, s_w);
			s_r = __fadd_rn(r_r, s_r);
			s_rr = __fadd_rn(r_rr, s_rr);
		}
		inv_w = rcp_approx(s_w);
		mean_ref = __fmul_rn(inv_w, s_r);
		var_ref = __fmaf_rn(inv_w, s_rr, -__fmul_rn(mean_ref, mean_ref));
		degenerate = (var_ref < kMinVar);
	}
};

// final NCC formula shared by both paths (reference APD.cu:1091-1109 as compiled)
__device__ __forceinline__ float ncc_finish(const RefPatch& rp, float s_s, float s_ss, float s_rs) {
	const float mean_src = __fmul_rn(rp.inv_w, s_s);
	const float var_src = __fmaf_rn(rp.inv_w, s_ss, -__fmul_rn(mean_src, mean_src));
	const float e_rs = __fmul_rn(rp.inv_w, s_rs);
	if (rp.var_ref < kMinVar || var_src < kMinVar) return kCostMax;
	const float covar = __fmaf_rn(-rp.mean_ref, mean_src, e_rs);
	const float den = sqrt_approx(__fmul_rn(rp.var_ref, var_src));
	const float c = __fmaf_rn(-covar, rcp_approx(den), 1.0f);
	return fmaxf(0.0f, fminf(kCostMax, c));
}

// One bilateral NCC of pixel (px,py) against source view `v` (0-based) under plane hypothesis `pl`.
// RB = patch rows whose 6 texture fetches are issued back to back before any is consumed (RB*6 fetches in
// flight per thread).  The propagation sweep is shared-memory-limited to 12 warps/SM, has registers to
// spare and is latency-bound on TEX (ncu: long-scoreboard stalls), so it uses RB = 6; the 24-warp kernels use 2.
template <int RB, bool RW = false>
__device__ __forceinline__ float ncc_cost(const KArgs& a, const ViewConst& vc, cudaTextureObject_t src, int px, int py,
                                          const float4 pl, const RefPatch& rp, const float2* wt, int stride) {
	float H[9];
	compute_homography(a.ref, vc, pl, H);
	const float fpx = (float)px, fpy = (float)py;
	{   // centre must project inside the source image (APD.cu:1038-1041); expression shape as compiled
		const float z = __fadd_rn(H[8], __fmaf_rn(fpx, H[6], __fmul_rn(fpy, H[7])));
		const float rz = rcp_approx(z);
		const float x = __fmul_rn(__fadd_rn(H[2], __fmaf_rn(fpx, H[0], __fmul_rn(fpy, H[1]))), rz);
		const float y = __fmul_rn(__fadd_rn(H[5], __fmaf_rn(fpx, H[3], __fmul_rn(fpy, H[4]))), rz);
		if (x >= (float)a.W || x < 0.0f || y >= (float)a.H || y < 0.0f) return kCostMax;
	}
	if (rp.degenerate) return kCostMax;

	float s_s = 0.f, s_ss = 0.f, s_rs = 0.f;
	DVP_COUNT(a, rp.n * rp.n);
	if (rp.hoisted) {
		static_assert(kHoistAxis % RB == 0, "row batch must divide the patch height");
#pragma unroll 1
		for (int ii = 0; ii < kHoistAxis; ii += RB) {
			float sv[RB * kHoistAxis];
#pragma unroll
			for (int r = 0; r < RB; ++r) {
				const float xf = (float)(px - rp.radius + (ii + r) * rp.inc);
				const float hx = __fmul_rn(H[0], xf), hy = __fmul_rn(H[3], xf), hz = __fmul_rn(H[6], xf);
#pragma unroll
				for (int jj = 0; jj < kHoistAxis; ++jj) {
					const float yf = (float)(py - rp.radius + jj * rp.inc);
					const float z = __fadd_rn(H[8], __fmaf_rn(H[7], yf, hz));
					const float x = __fadd_rn(H[2], __fmaf_rn(H[1], yf, hx));
					const float y = __fadd_rn(H[5], __fmaf_rn(H[4], yf, hy));
					const float rz = rcp_approx(z);
					sv[r * kHoistAxis + jj] = tex2D<float>(src, __fmaf_rn(x, rz, 0.5f), __fmaf_rn(y, rz, 0.5f));
				}
			}
#pragma unroll
			for (int r = 0; r < RB; ++r) {
				float r_s = 0.f, r_ss = 0.f, r_rs = 0.f;
#pragma unroll
				for (int jj = 0; jj < kHoistAxis; ++jj) {
					float2 w_t;
					if (RW) {
						const int k = (ii + r) * kHoistAxis + jj;
						const float pix = reinterpret_cast<const float*>(wt)[k * stride];
						const float sd = rp.sd_const ? c_sd_r5[k] : RefPatch::spatial_dist(-rp.radius + (ii + r) * rp.inc, -rp.radius + jj * rp.inc);
						w_t.x = RefPatch::weight_sd(sd, pix, rp.center, rp.rcp_s, rp.rcp_c);
						w_t.y = __fmul_rn(pix, w_t.x);
					} else {
						w_t = wt[((ii + r) * kHoistAxis + jj) * stride];
					}
					const float s = sv[r * kHoistAxis + jj];
					const float u = __fmul_rn(s, w_t.x);
					r_rs = __fmaf_rn(s, w_t.y, r_rs);
					r_ss = __fmaf_rn(s, u, r_ss);
					r_s = __fadd_rn(u, r_s);
				}
				s_s = __fadd_rn(r_s, s_s);
				s_ss = __fadd_rn(r_ss, s_ss);
				s_rs = __fadd_rn(r_rs, s_rs);
			}
		}
	} else {
		// general radius (not a multiple of 5): recompute the reference side per call, like the reference
		float rcp_s, rcp_c; RefPatch::sigma_rcps(a.prm, rcp_s, rcp_c);
		const float center = RefPatch::ref_pixel(a, px, py);
		for (int i = -rp.radius; i <= rp.radius; i += rp.inc) {
			const float xf = (float)(px + i);
			const float hx = __fmul_rn(H[0], xf), hy = __fmul_rn(H[3], xf), hz = __fmul_rn(H[6], xf);
			float r_s = 0.f, r_ss = 0.f, r_rs = 0.f;
			for (int j = -rp.radius; j <= rp.radius; j += rp.inc) {
				const float yf = (float)(py + j);
				const float z = __fadd_rn(H[8], __fmaf_rn(H[7], yf, hz));
				const float x = __fadd_rn(H[2], __fmaf_rn(H[1], yf, hx));
				const float y = __fadd_rn(H[5], __fmaf_rn(H[4], yf, hy));
				const float rz = rcp_approx(z);
				const float s = tex2D<float>(src, __fmaf_rn(x, rz, 0.5f), __fmaf_rn(y, rz, 0.5f));
				const float pix = RefPatch::ref_pixel(a, px + i, py + j);
				const float w = RefPatch::weight(i, j, pix, center, rcp_s, rcp_c);
				const float t = __fmul_rn(pix, w);
				const float u = __fmul_rn(s, w);
				r_rs = __fmaf_rn(s, t, r_rs);
				r_ss = __fmaf_rn(s, u, r_ss);
				r_s = __fadd_rn(u, r_s);
			}
			s_s = __fadd_rn(r_s, s_s);
			s_ss = __fadd_rn(r_ss, s_ss);
			s_rs = __fadd_rn(r_rs, s_rs);
		}
	}
	return ncc_finish(rp, s_s, s_ss, s_rs);
}

// (An out-of-line, __noinline__ variant of ncc_cost was tried to shrink the code: it halves throughput,
// because the source texture handle stops being warp-uniform inside the callee and every fetch turns into the
// divergent-handle loop; the call sites stay inlined.)
// Forward-backward reprojection error against the neighbour depth map, clamped at 3 px.
// reference ComputeGeomConsistencyCost, APD.cu:1218-1256 (same expression shapes).
__device__ __forceinline__ float geom_cost(const KArgs& a, const ViewConst& vc, cudaTextureObject_t depth_tex, int px, int py, const float4 pl) {
	const float max_cost = 3.0f;
	float depth = depth_from_plane(a.ref, pl, px, py);
	float3 fwd = point_to_world((float)px, (float)py, depth, a.ref.K, a.ref.R, a.ref.c);
	float2 src_pt; float src_d;
	project_on_camera(fwd, vc.sK, vc.sR, vc.st, src_pt, src_d);
	const float src_depth = tex2D<float>(depth_tex, (int)src_pt.x + 0.5f, (int)src_pt.y + 0.5f);
	DVP_COUNT(a, 1);
	if (src_depth == 0.0f) return max_cost;
	float3 back = point_to_world(src_pt.x, src_pt.y, src_depth, vc.sK, vc.sR, vc.sc);
	float2 bpt; float ref_d;
	project_on_camera(back, a.ref.K, a.ref.R, a.ref.t, bpt, ref_d);
	const float dc = px - bpt.x;
	const float dr = py - bpt.y;
	const float cost = sqrt(dc * dc + dr * dr);
	return min(max_cost, cost);
}

}  // namespace dvp
