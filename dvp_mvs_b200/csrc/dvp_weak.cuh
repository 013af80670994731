// dvp_weak.cuh — device code of the adaptive-patch-deformation (WEAK pixel) path:
// Bresenham edge test, point-in-triangle, deformable bilateral NCC (reference APD.cu:244-317, 835-1021).
// WEAK pixels are a minority of a view and their work is irregular (data-dependent RNG draws, RANSAC),
// so this path keeps the reference's one-thread-per-pixel formulation; what changes is where it runs from:
// kernels are launched over the compact list of WEAK pixels instead of the whole image, parameters come from
// the kernel argument block, and the 25.6 KB/thread edge-test cache of GenNeighbours shrinks to 6.4 KB.
#pragma once
#include "dvp_strong.cuh"

namespace dvp {

// reference BresenhamLine, APD.cu:267-311: does the segment B->A cross an edge pixel within max(H,W)/30 steps?
// BATCH = iterations advanced per round of edge reads (1 = the reference's loop shape).
constexpr int kCoarseWalkMin = 4;   // shorter walks keep the plain loop (the set-up of the closed form would cost more than it saves)
template <int BATCH = 1>
__device__ __forceinline__ bool bresenham_crosses_edge(const KArgs& a, int Ax, int Ay, int Bx, int By) {
	const int width = a.W;
	const int max_step = (int)(DVP_MAX(a.H, a.W) / 30.0);
	int x0 = Bx, y0 = By;
	const int x1 = Ax, y1 = Ay;
	const int ABx = Ax - Bx, ABy = Ay - By;
	if (ABx * ABx + ABy * ABy > 9 * max_step * max_step) return false;
	if (a.edge[x0 + y0 * width] || a.edge[x1 + y1 * width]) return false;
	const int dx = abs(x1 - x0), sx = x0 < x1 ? 1 : -1;
	const int dy = abs(y1 - y0), sy = y0 < y1 ? 1 : -1;
	int erro = (dx > dy ? dx : dy) / 2;
	int step = 0;
	bool tagx = true, tagy = true;
#ifndef DVP_NO_COARSE_WALK
	// Walks longer than a few steps skip ahead on a distance map instead of reading the edge map once per step.  The loop
	// below visits exactly n = min(max(dx, dy) + 1, max_step) positions (it oversteps the end point by one) and position k
	// (1-based; k = 0 is the start) is
	//     x-major (dx > dy):  (x0 + sx k,            y0 + sy floor((k dy + dx - 1 - e0) / dx))
	//     y-major (dy > dx):  (x0 + sx min(k, floor((k dx + dy - 1 + e0) / dy)),   y0 + sy k)          e0 = max(dx, dy) / 2
	//     diagonal:           (x0 + sx k, y0 + sy k)
	// (checked against the loop in tests/test_host_logic.py).  Consecutive positions differ by at most one pixel per axis.
	// a.edge_dist holds the chessboard distance d to the nearest edge pixel: standing on position k with d > 0, positions
	// k+1 .. k+d-1 lie within chessboard distance d-1 and cannot be edge pixels, so the walk continues at k+d.  The answer is
	// the loop's; in an edge-free region a 200-step walk takes a handful of reads.
	{
		const int M = dx > dy ? dx : dy;
		const int n_total = M + 1 < max_step ? M + 1 : max_step;
		if (M > 0 && n_total > kCoarseWalkMin) {
			const bool xmaj = dx > dy, diag = dx == dy;
			const int mn = dx > dy ? dy : dx, e0 = M / 2;
			const int c = xmaj ? M - 1 - e0 : e0 - 1 + M;
			const float rcp_m = __frcp_rn((float)M);
			int k = 0, d = a.edge_dist[x0 + y0 * width];    // > 0: the start pixel is no edge pixel (tested above)
			for (;;) {
				k += d;
				if (k > n_total) return false;
				int mi = k;                                   // how far the minor axis has moved after k steps
				if (!diag) {
					const int n = k * mn + c;                 // < 2^24: exact in float; the estimate is off by at most one
					int q = (int)(__fmul_rn((float)n, rcp_m));
					const int r = n - q * M;
					q += (r >= M) - (r < 0);
					mi = xmaj ? q : (k < q ? k : q);
				}
				const int x = dy > dx ? x0 + sx * mi : x0 + sx * k, y = dy > dx ? y0 + sy * k : y0 + sy * mi;
				d = 1;                                        // the walk may overstep the map by one pixel: nothing to test there
				if (x >= 0 && x < a.W && y >= 0 && y < a.H) {
					d = a.edge_dist[x + y * width];
					if (d == 0) return true;
				}
			}
		}
	}
#endif
	// The reference tests one pixel per iteration and returns at the first edge pixel.  Only the yes/no answer
	// leaves this function, so BATCH iterations are advanced at a time (pure integer state, no memory) and their
	// edge reads are issued together: a hit in any of them is a hit, and iterations the reference would not
	// have run (end of the segment, step limit) are masked out exactly as its loop condition does.
	// Measured on B200: BATCH 4 helps K9 (long clear walks between anchors, 8.2 -> 6.7 ms) and hurts K4 (most walks
	// end within a few pixels and the lanes diverge, 36 -> 62 ms), so K4 keeps BATCH 1.
	if (BATCH == 1) {   // the reference's loop as it stands
		while (tagx || tagy) {
			if (x0 == x1) tagx = false;
			if (y0 == y1) tagy = false;
			const int e2 = erro;
			if (e2 > -dx) { erro -= dy; x0 += sx; }
			if (e2 < dy) { erro += dx; y0 += sy; }
			// the walk may step one pixel past the end point; stay inside the map (the reference reads whatever is there)
			if (x0 >= 0 && x0 < a.W && y0 >= 0 && y0 < a.H) { if (a.edge[x0 + y0 * width]) return true; }
			step += 1;
			if (step >= max_step) break;
		}
		return false;
	}
	bool go = true;   // the reference's first iteration always runs (tagx, tagy start true)
	while (go) {
		int idx[BATCH];
#pragma unroll
		for (int u = 0; u < BATCH; ++u) {
			idx[u] = -1;
			if (go) {
				if (x0 == x1) tagx = false;
				if (y0 == y1) tagy = false;
				const int e2 = erro;
				if (e2 > -dx) { erro -= dy; x0 += sx; }
				if (e2 < dy) { erro += dx; y0 += sy; }
				// the walk may step one pixel past the end point; stay inside the map (the reference reads whatever is there)
				if (x0 >= 0 && x0 < a.W && y0 >= 0 && y0 < a.H) idx[u] = x0 + y0 * width;
				step += 1;
				go = (step < max_step) && (tagx || tagy);
			}
		}
		uint8_t hit = 0;
#pragma unroll
		for (int u = 0; u < BATCH; ++u) hit |= idx[u] >= 0 ? a.edge[idx[u]] : (uint8_t)0;
		if (hit) return true;
	}
	return false;
}

// reference PointinTriangle, APD.cu:244-265
__device__ __forceinline__ bool point_in_triangle(short2 A, short2 B, short2 C, int Px, int Py) {
	const float2 AB = make_float2(B.x - A.x, B.y - A.y);
	const float2 BC = make_float2(C.x - B.x, C.y - B.y);
	const float2 CA = make_float2(A.x - C.x, A.y - C.y);
	const float AB_ = sqrt(AB.x * AB.x + AB.y * AB.y);
	const float BC_ = sqrt(BC.x * BC.x + BC.y * BC.y);
	const float CA_ = sqrt(CA.x * CA.x + CA.y * CA.y);
	if (AB_ <= 2 || BC_ <= 2 || CA_ <= 2) return false;
	if (!(AB_ + BC_ > CA_ && BC_ + CA_ > AB_ && AB_ + CA_ > BC_)) return false;
	const float2 PA = make_float2(A.x - Px, A.y - Py);
	const float2 PB = make_float2(B.x - Px, B.y - Py);
	const float2 PC = make_float2(C.x - Px, C.y - Py);
	const float t1 = PA.x * PB.y - PA.y * PB.x;
	const float t2 = PB.x * PC.y - PB.y * PC.x;
	const float t3 = PC.x * PA.y - PC.y * PA.x;
	return t1 * t2 >= 0 && t1 * t3 >= 0;
}

// colour-only bilateral weight (ComputeBilateralWeight_YZL, APD.cu:783-788) as compiled:
// ex2(log2e * (|pix - centre| * -(1/(2 sc^2))))
__device__ __forceinline__ float weight_colour(float pix, float center, float rcp_c) {
	const float cd = fabsf(__fadd_rn(pix, -center));
	return ex2_approx(__fmul_rn(__fmul_rn(cd, -rcp_c), 1.4426950216293334961f));
}

// projective warp of an integer pixel, both products varying (APD.cu:741-748 as compiled inside NCCNew):
// num = H2 + fma(H0, x, H1*y)
__device__ __forceinline__ void warp_point(const float* H, float xf, float yf, float& u, float& v) {
	const float z = __fadd_rn(H[8], __fmaf_rn(H[6], xf, __fmul_rn(H[7], yf)));
	const float x = __fadd_rn(H[2], __fmaf_rn(H[0], xf, __fmul_rn(H[1], yf)));
	const float y = __fadd_rn(H[5], __fmaf_rn(H[3], xf, __fmul_rn(H[4], yf)));
	const float rz = rcp_approx(z);
	u = __fmul_rn(x, rz);
	v = __fmul_rn(y, rz);
}
__device__ __forceinline__ float sample_src_warped(const float* H, cudaTextureObject_t src, float xf, float yf) {
	const float z = __fadd_rn(H[8], __fmaf_rn(H[6], xf, __fmul_rn(H[7], yf)));
	const float x = __fadd_rn(H[2], __fmaf_rn(H[0], xf, __fmul_rn(H[1], yf)));
	const float y = __fadd_rn(H[5], __fmaf_rn(H[3], xf, __fmul_rn(H[4], yf)));
	const float rz = rcp_approx(z);
	return tex2D<float>(src, __fmaf_rn(x, rz, 0.5f), __fmaf_rn(y, rz, 0.5f));
}

__device__ __forceinline__ float ncc_tail(float s_w, float s_r, float s_rr, float s_s, float s_ss, float s_rs) {
	const float inv = rcp_approx(s_w);
	const float mr = __fmul_rn(inv, s_r), ms = __fmul_rn(inv, s_s);
	const float var_r = __fmaf_rn(inv, s_rr, -__fmul_rn(mr, mr));
	const float var_s = __fmaf_rn(inv, s_ss, -__fmul_rn(ms, ms));
	const float e_rs = __fmul_rn(inv, s_rs);
	if (var_r < kMinVar || var_s < kMinVar) return kCostMax;
	const float covar = __fmaf_rn(-mr, ms, e_rs);
	const float den = sqrt_approx(__fmul_rn(var_r, var_s));
	return fmaxf(0.0f, fminf(kCostMax, __fmaf_rn(-covar, rcp_approx(den), 1.0f)));
}

// reference ComputeBilateralNCCNew, APD.cu:835-1021: 0.25 * centre patch (6x6 at the adaptive radius, colour-only
// weights) + 0.75 * mean over the <= 11 anchor pixels of a 9-sample NCC whose offsets are the anchor's
// visibility-aware `candidate` offsets for this view (fallback: the +-5 ring).
// Tried on B200 and rejected (bench workload, K10 + K11 = 53.8 ms; all bit-exact): scoring the hypotheses of a pixel in batches
// with the reference side formed once per (anchor, view) — 12 % fewer instructions, but 2.2 KB of stack per thread and 2.3 GB
// of local-memory DRAM traffic per launch (64.4 ms); a two-deep software pipeline over the anchors with the centre patch read a
// row at a time — long-scoreboard stalls 4.3 -> 3.4 per issue, but 200 KB of SASS and instruction-fetch stalls (61.0 ms;
// __noinline__: 61.8 ms).  ncu (profiles/r02_top_kernels_ncu.txt): TEX data pipe 54 %, LSU data pipe 41 %, issue 45 %, 24 warps.
__device__ __forceinline__ float ncc_new(const KArgs& a, int px, int py, int v /*0-based view*/, const float4 pl) {
	const int W = a.W, H_ = a.H;
	const ViewConst& vc = a.views[v];
	const cudaTextureObject_t src = a.tex_img[v + 1];
	float H[9];
	compute_homography(a.ref, vc, pl, H);
	{
		float u, w; warp_point(H, (float)px, (float)py, u, w);
		if (u >= (float)W || u < 0.0f || w >= (float)H_ || w < 0.0f) return kCostMax;
	}
	const int center = px + py * W;
	float rcp_s, rcp_c; RefPatch::sigma_rcps(a.prm, rcp_s, rcp_c);
	const float ref_center_pix = RefPatch::ref_pixel(a, px, py);
	const short2* nb = a.neighbours + (size_t)a.neighbours_map[center] * DVP_NEIGHBOUR_NUM;
	float center_cost = 0.0f, strong_cost = 0.0f;
	int strong_count = 0;
	for (int k = 0; k < DVP_NEIGHBOUR_NUM; ++k) {
		const short2 np = nb[k];
		if (np.x == -1 || np.y == -1) continue;
		{
			float u, w; warp_point(H, (float)np.x, (float)np.y, u, w);
			if (u < 0 || w < 0 || u >= (float)W || w >= (float)H_) {
				if (k != 0) {
					if (is_set(a.selected[np.x + np.y * W], v)) { strong_cost += kCostMax; strong_count++; }
					continue;
				}
				return kCostMax;
			}
		}
		float s_r = 0.f, s_rr = 0.f, s_s = 0.f, s_ss = 0.f, s_rs = 0.f, s_w = 0.f;
		if (k == 0) {
			int radius = a.prm.strong_radius, inc = a.prm.strong_increment;
			if (a.prm.use_radius) { radius = a.radius[center]; inc = DVP_MAX(2, (int)(2.0 * radius / 5.0)); }
			// radius 0 (what K9 stores for every pixel with a fit plane): ONE sample, the pixel itself.  Its weight is
			// ex2(-0) = 1, the sum of weights 1, rcp.approx(1) = 1, so the reference-side variance is fma(1, r*r, -(r*r)) = 0
			// < 1e-5 whatever the source holds: the reference returns cost_max.  Decided without the load and the fetch.
			if (radius == 0 && np.x == px && np.y == py) { center_cost = kCostMax; continue; }
			DVP_COUNT(a, radius >= 0 ? ((2 * radius) / inc + 1) * ((2 * radius) / inc + 1) : 0);
			for (int i = -radius; i <= radius; i += inc) {
				const float xf = (float)(np.x + i);
				const float hx = __fmul_rn(H[0], xf), hy = __fmul_rn(H[3], xf), hz = __fmul_rn(H[6], xf);
				float r_r = 0.f, r_rr = 0.f, r_s = 0.f, r_ss = 0.f, r_rs = 0.f, r_w = 0.f;
				for (int j = -radius; j <= radius; j += inc) {
					const float yf = (float)(np.y + j);
					const float ref_pix = RefPatch::ref_pixel(a, np.x + i, np.y + j);
					const float z = __fadd_rn(H[8], __fmaf_rn(H[7], yf, hz));
					const float x = __fadd_rn(H[2], __fmaf_rn(H[1], yf, hx));
					const float y = __fadd_rn(H[5], __fmaf_rn(H[4], yf, hy));
					const float rz = rcp_approx(z);
					const float src_pix = tex2D<float>(src, __fmaf_rn(x, rz, 0.5f), __fmaf_rn(y, rz, 0.5f));
					const float w = weight_colour(ref_pix, ref_center_pix, rcp_c);
					const float t = __fmul_rn(ref_pix, w), u = __fmul_rn(src_pix, w);
					r_r = __fadd_rn(t, r_r); r_rr = __fmaf_rn(ref_pix, t, r_rr);
					r_s = __fadd_rn(u, r_s); r_ss = __fmaf_rn(src_pix, u, r_ss);
					r_rs = __fmaf_rn(src_pix, t, r_rs);
					r_w = __fadd_rn(w, r_w);
				}
				s_r = __fadd_rn(r_r, s_r); s_rr = __fadd_rn(r_rr, s_rr); s_s = __fadd_rn(r_s, s_s);
				s_ss = __fadd_rn(r_ss, s_ss); s_rs = __fadd_rn(r_rs, s_rs); s_w = __fadd_rn(r_w, s_w);
			}
		} else if (is_set(a.selected[np.x + np.y * W], v) == 1) {
			const int nei_center = np.x + np.y * W;
			// the anchor's 8 visibility-aware offsets for this view: one 32-byte record, two 16-byte loads
			const uint4* cand4 = reinterpret_cast<const uint4*>(a.candidate + ((size_t)nei_center * DVP_NUM_IMAGES + v) * DVP_LAB_BOUNDARY_NUM);
			const uint4 c_lo = __ldg(cand4), c_hi = __ldg(cand4 + 1);
			const uint32_t cw[8] = {c_lo.x, c_lo.y, c_lo.z, c_lo.w, c_hi.x, c_hi.y, c_hi.z, c_hi.w};
			const int def_i[9] = {-5, -5, -5, 0, 0, 5, 5, 5, 0}, def_j[9] = {-5, 0, 5, -5, 5, -5, 0, 5, 0};
			// all nine reference reads and texture fetches are issued before the first is consumed; the sums are
			// then folded strictly in sample order (the reference's rounding sequence)
			float ref_pix[9], src_pix[9];
			DVP_COUNT(a, 9);
#pragma unroll
			for (int q = 0; q < 9; q++) {
				int i = 0, j = 0;
				if (q != 8) { i = (int)(short)(cw[q] & 0xffffu); j = (int)(short)(cw[q] >> 16); }
				if (i == 0 && j == 0) { i = def_i[q]; j = def_j[q]; }
				const int rx = np.x + i, ry = np.y + j;
				ref_pix[q] = RefPatch::ref_pixel(a, rx, ry);
				src_pix[q] = sample_src_warped(H, src, (float)rx, (float)ry);
			}
#pragma unroll
			for (int q = 0; q < 9; q++) {
				const float w = weight_colour(ref_pix[q], ref_center_pix, rcp_c);
				// each "row" is one sample: row = fma(x, y, 0), total += row -> product rounded, then added (SASS of the reference).
				// The reference's `0 + t`, `0 + u`, `0 + w` are not issued: w = ex2(.) > 0, the grey values are >= +0 and every
				// product is already flushed (FTZ), so none of the three can be -0 or denormal and 0 + x == x bit for bit.
				const float t = __fmul_rn(ref_pix[q], w), u = __fmul_rn(src_pix[q], w);
				s_r = __fadd_rn(t, s_r);
				s_rr = __fadd_rn(__fmul_rn(ref_pix[q], t), s_rr);
				s_s = __fadd_rn(u, s_s);
				s_ss = __fadd_rn(__fmul_rn(src_pix[q], u), s_ss);
				s_rs = __fadd_rn(__fmul_rn(src_pix[q], t), s_rs);
				s_w = __fadd_rn(w, s_w);
			}
		}
		const float temp_cost = ncc_tail(s_w, s_r, s_rr, s_s, s_ss, s_rs);
		if (k == 0) center_cost = temp_cost;
		else { strong_cost += temp_cost; strong_count++; }
	}
	float cost;
	if (strong_count == 0) cost = center_cost;
	else {
		strong_cost /= strong_count;
		strong_cost = DVP_MIN(strong_cost, kCostMax);
		cost = 0.25 * center_cost + 0.75 * strong_cost;
	}
	return cost;
}


}  // namespace dvp
