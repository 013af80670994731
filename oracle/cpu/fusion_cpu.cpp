// oracle/cpu/fusion_cpu.cpp — TEST INFRASTRUCTURE ONLY.
// CPU restatement of the reference's depth-map fusion: RunFusion, the "ETH version" that
// main() calls (APD.cpp:1809-1960, main.cpp:514), with Get3DPointonWorld (APD.cpp:501-525), ProjectCamera
// (APD.cpp:536-546) and GetAngle (APD.cpp:1797-1806).  Files are replaced by arrays: what the reference reads with
// ReadBinMat / imread and rescales with RescaleImageAndCamera / RescaleMatToTargetSize (APD.cpp:1841-1873) is the
// caller's input here (camera and colour image already at the depth map's size, weak map already at that size).
//
// PARITY UNPINNED: APD.cpp cannot be built here (OpenCV + Boost), the reference ships no fusion fixtures, and its own
// host build is `-O3 -ffast-math -march=native` (CMakeLists.txt:31), i.e. its float results depend on the build
// machine.  This restatement fixes the arithmetic the source text asks for under strict IEEE evaluation:
//   * float expressions evaluated left to right in float, no contraction (-ffp-contract=off);
//   * `sqrt(pow(a,2)+pow(b,2))` in double (std::pow(float,int) promotes to double), rounded to float on assignment;
//   * `exp(float)`, `fabs(float)` resolve to the float overloads (curand_kernel.h, included by APD.cpp:16, pulls in
//     <math.h>, whose libstdc++ wrapper exports them to the global namespace); `acosf` is called by name;
//   * `int(x + 0.5f)` truncates toward zero; a NaN or out-of-range value converts to INT_MIN on x86 (cvttss2si),
//     which fails the bounds test — stated explicitly here.
//
// Three entry points:
//   fusion_cpu_run        — the reference's loop as written (views in order, pixels in raster order).
//   fusion_cpu_candidates — the mask-independent part of one view's inner loop: for (pixel, source) the source cell the
//                           pixel would claim and its exp(-index) term, or -1.
//   fusion_cpu_resolve    — the mask-dependent part, sequential, from given candidates.  run == candidates + resolve
//                           (checked in the CPU suite); the CUDA path is compared stage by stage against these.
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "dvp_mvs.h"

namespace {

struct F3 { float x, y, z; };

// APD.cpp:501-525
F3 point_on_world(int x, int y, float depth, const dvp_camera& cam) {
	F3 p, t;
	p.x = depth * (x - cam.K[2]) / cam.K[0];
	p.y = depth * (y - cam.K[5]) / cam.K[4];
	p.z = depth;
	t.x = cam.R[0] * p.x + cam.R[3] * p.y + cam.R[6] * p.z;
	t.y = cam.R[1] * p.x + cam.R[4] * p.y + cam.R[7] * p.z;
	t.z = cam.R[2] * p.x + cam.R[5] * p.y + cam.R[8] * p.z;
	F3 c;
	c.x = -(cam.R[0] * cam.t[0] + cam.R[3] * cam.t[1] + cam.R[6] * cam.t[2]);
	c.y = -(cam.R[1] * cam.t[0] + cam.R[4] * cam.t[1] + cam.R[7] * cam.t[2]);
	c.z = -(cam.R[2] * cam.t[0] + cam.R[5] * cam.t[1] + cam.R[8] * cam.t[2]);
	p.x = t.x + c.x;
	p.y = t.y + c.y;
	p.z = t.z + c.z;
	return p;
}

// APD.cpp:536-546
void project(const F3& X, const dvp_camera& cam, float& px, float& py, float& depth) {
	F3 t;
	t.x = cam.R[0] * X.x + cam.R[1] * X.y + cam.R[2] * X.z + cam.t[0];
	t.y = cam.R[3] * X.x + cam.R[4] * X.y + cam.R[5] * X.z + cam.t[1];
	t.z = cam.R[6] * X.x + cam.R[7] * X.y + cam.R[8] * X.z + cam.t[2];
	depth = cam.K[6] * t.x + cam.K[7] * t.y + cam.K[8] * t.z;
	px = (cam.K[0] * t.x + cam.K[1] * t.y + cam.K[2] * t.z) / depth;
	py = (cam.K[3] * t.x + cam.K[4] * t.y + cam.K[5] * t.z) / depth;
}

// APD.cpp:1797-1806
float get_angle(const float* a, const float* b) {
	const float dot = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
	const float angle = acosf(dot);
	if (angle != angle) return 0.0f;
	return angle;
}

// int(v) as x86 evaluates it: truncation, INT_MIN for NaN and for values outside int's range
int to_int(float v) {
	if (!(v > -2147483648.0f && v < 2147483648.0f)) return INT_MIN;
	return (int)v;
}

// The mask-independent part of one (pixel, source) evaluation, shared by the three fusion variants
// (APD.cpp:1901-1921, 2067-2085, 2231-2251): the source cell the pixel projects to and, if that cell holds a depth, the
// three consistency measures.  Returns the cell (src_r * src_w + src_c) or -1 when the reference `continue`s / skips.
int measures(const dvp_fusion_view* views, int ref, int j, int r, int c, float ref_depth, const F3& X, float* reproj_error, float* relative_depth_diff, float* angle) {
	const dvp_fusion_view& rv = views[ref];
	const dvp_fusion_view& sv = views[rv.src_views[j]];
	float px, py, proj_depth;
	project(X, sv.camera, px, py, proj_depth);
	const int src_r = to_int(py + 0.5f);
	const int src_c = to_int(px + 0.5f);
	if (!(src_c >= 0 && src_c < sv.width && src_r >= 0 && src_r < sv.height)) return -1;
	const size_t cell = (size_t)src_r * sv.width + src_c;
	const float src_depth = sv.depth[cell];
	if (src_depth <= 0.0) return -1;
	const F3 Y = point_on_world(src_c, src_r, src_depth, sv.camera);
	float qx, qy;
	project(Y, rv.camera, qx, qy, proj_depth);
	const double dx = (double)(c - qx), dy = (double)(r - qy);
	*reproj_error = (float)std::sqrt(dx * dx + dy * dy);
	*relative_depth_diff = std::fabs(proj_depth - ref_depth) / ref_depth;
	*angle = get_angle(rv.normal + 3 * ((size_t)r * rv.width + c), sv.normal + 3 * cell);
	return (int)cell;
}

// The mask-independent half of APD.cpp:1899-1931 for pixel (r, c) of view `ref` and its j-th source.
// Returns the source cell and the term exp(-tmp_index), or -1.
int candidate(const dvp_fusion_view* views, int ref, int j, int r, int c, float ref_depth, const F3& X, float* term) {
	float reproj_error, relative_depth_diff, angle;
	const int cell = measures(views, ref, j, r, c, ref_depth, X, &reproj_error, &relative_depth_diff, &angle);
	if (cell < 0) return -1;
	if (reproj_error < 2.0f && relative_depth_diff < 0.01f && angle < 0.174533f) {
		const float tmp_index = reproj_error + 200 * relative_depth_diff + angle * 10;
		*term = expf(-tmp_index);
		return cell;
	}
	return -1;
}

// APD.cpp:1933-1954 once the consistent sources are known: accept test, colour average, mask writes.
// `cells[j]` / `terms[j]` hold the unmasked consistent sources (or -1).  Returns the used-source bit mask (0 = no point).
uint32_t accept(const dvp_fusion_view* views, int ref, int r, int c, const int32_t* cells, const float* terms, uint8_t* const* masks, float* color) {
	const dvp_fusion_view& rv = views[ref];
	int num_consistent = 0;
	float dynamic_consistency = 0.0f;
	for (int j = 0; j < rv.num_src; ++j)
		if (cells[j] >= 0) { dynamic_consistency += terms[j]; num_consistent++; }
	const size_t p = (size_t)r * rv.width + c;
	const float factor = (rv.weak[p] == DVP_WEAK ? 0.45f : 0.3f);
	if (!(num_consistent >= 1 && (dynamic_consistency > factor * num_consistent))) return 0;
	color[0] = (float)rv.image[3 * p + 0]; color[1] = (float)rv.image[3 * p + 1]; color[2] = (float)rv.image[3 * p + 2];
	uint32_t used = 0;
	for (int j = 0; j < rv.num_src; ++j) {
		if (cells[j] < 0) continue;
		const int s = rv.src_views[j];
		masks[s][cells[j]] = 1;
		const uint8_t* col = views[s].image + 3 * (size_t)cells[j];
		color[0] += col[0]; color[1] += col[1]; color[2] += col[2];
		used |= 1u << j;
	}
	color[0] /= (num_consistent + 1); color[1] /= (num_consistent + 1); color[2] /= (num_consistent + 1);
	return used;
}

bool skipped(const dvp_fusion_view& rv, size_t p, uint8_t* const* masks, int ref) {
	if (rv.block && rv.block[p] < 128) return true;   // APD.cpp:1884-1886
	if (masks[ref][p] == 1) return true;                // APD.cpp:1888-1890
	return rv.depth[p] <= 0.0;                          // APD.cpp:1892-1894
}

}  // namespace

extern "C" {

// RunFusion, APD.cpp:1875-1957.  masks[v]: height*width bytes per view, zeroed by the caller (APD.cpp:1867).
// points: [capacity][6] floats (coord xyz, colour in the image's channel order), in the reference's push order.
// Returns the number of points the reference would hold (may exceed capacity; the excess is not stored).
long long fusion_cpu_run(int num_views, const dvp_fusion_view* views, uint8_t* const* masks, float* points, long long capacity) {
	long long n = 0;
	std::vector<int32_t> cells;
	std::vector<float> terms;
	for (int i = 0; i < num_views; ++i) {
		const dvp_fusion_view& rv = views[i];
		cells.assign(rv.num_src, -1);
		terms.assign(rv.num_src, 0.0f);
		for (int r = 0; r < rv.height; ++r)
			for (int c = 0; c < rv.width; ++c) {
				const size_t p = (size_t)r * rv.width + c;
				if (skipped(rv, p, masks, i)) continue;
				const float ref_depth = rv.depth[p];
				const F3 X = point_on_world(c, r, ref_depth, rv.camera);
				for (int j = 0; j < rv.num_src; ++j) {
					cells[j] = candidate(views, i, j, r, c, ref_depth, X, &terms[j]);
					if (cells[j] >= 0 && masks[rv.src_views[j]][cells[j]] == 1) cells[j] = -1;   // APD.cpp:1911-1912
				}
				float color[3];
				if (accept(views, i, r, c, cells.data(), terms.data(), masks, color)) {
					if (n < capacity) {
						float* o = points + 6 * n;
						o[0] = X.x; o[1] = X.y; o[2] = X.z; o[3] = color[0]; o[4] = color[1]; o[5] = color[2];
					}
					++n;
				}
			}
	}
	return n;
}

// cells / terms: [height*width][num_src] of view `ref`.  Pixels the reference skips on their own depth or block mask
// get -1 everywhere; the pixel's own fusion mask is NOT looked at here (it is mask state, see resolve).
void fusion_cpu_candidates(const dvp_fusion_view* views, int ref, int32_t* cells, float* terms) {
	const dvp_fusion_view& rv = views[ref];
	const int S = rv.num_src;
	for (int r = 0; r < rv.height; ++r)
		for (int c = 0; c < rv.width; ++c) {
			const size_t p = (size_t)r * rv.width + c;
			for (int j = 0; j < S; ++j) { cells[p * S + j] = -1; terms[p * S + j] = 0.0f; }
			if ((rv.block && rv.block[p] < 128) || rv.depth[p] <= 0.0) continue;
			const float ref_depth = rv.depth[p];
			const F3 X = point_on_world(c, r, ref_depth, rv.camera);
			for (int j = 0; j < S; ++j) cells[p * S + j] = candidate(views, ref, j, r, c, ref_depth, X, &terms[p * S + j]);
		}
}

// The sequential greedy pass of view `ref` over given candidates: updates masks, writes used[p] (bit j = source j
// contributed; 0 = no point) and appends points.  Returns the number of points of this view.
long long fusion_cpu_resolve(const dvp_fusion_view* views, int ref, const int32_t* cells, const float* terms, uint8_t* const* masks, uint32_t* used, float* points, long long capacity) {
	const dvp_fusion_view& rv = views[ref];
	const int S = rv.num_src;
	long long n = 0;
	std::vector<int32_t> live(S);
	for (int r = 0; r < rv.height; ++r)
		for (int c = 0; c < rv.width; ++c) {
			const size_t p = (size_t)r * rv.width + c;
			used[p] = 0;
			if (skipped(rv, p, masks, ref)) continue;
			for (int j = 0; j < S; ++j) {
				live[j] = cells[p * S + j];
				if (live[j] >= 0 && masks[rv.src_views[j]][live[j]] == 1) live[j] = -1;
			}
			float color[3];
			used[p] = accept(views, ref, r, c, live.data(), terms + p * S, masks, color);
			if (used[p]) {
				if (n < capacity) {
					const F3 X = point_on_world(c, r, rv.depth[p], rv.camera);
					float* o = points + 6 * n;
					o[0] = X.x; o[1] = X.y; o[2] = X.z; o[3] = color[0]; o[4] = color[1]; o[5] = color[2];
				}
				++n;
			}
		}
	return n;
}

// RunFusion_TAT_Intermediate (mode 1, APD.cpp:1962-2130) and RunFusion_TAT_advanced (mode 2, APD.cpp:2132-2279): neither
// is called by main() (main.cpp:514 calls RunFusion).  A point needs k >= 2 sources within k-scaled limits; an emitted
// pixel masks ITSELF (APD.cpp:2121, 2270) and masked pixels are skipped when they are looked at as a source, so within a
// view nothing depends on the visiting order — except through `diff`: the vector of per-source measures is declared
// once per view (APD.cpp:2052, 2216), so a source that is not evaluated for a pixel (projects outside, hits a masked or
// empty cell) keeps the measures, and in mode 1 the colour cell, of the last pixel in raster order that did evaluate it.
// Reproduced as written.  used[p] (may be NULL): bit j = source j counted at the accepting k.
long long fusion_cpu_run_tat(int mode, int num_views, const dvp_fusion_view* views, uint8_t* const* masks, float* points, long long capacity, uint32_t* const* used_out) {
	const float dist_base = 0.25f;
	const float depth_base = mode == 1 ? 1.0f / 3500.0f : 1.0f / 3000.0f;
	const float angle_base = 0.06981317007977318f, angle_grad = 0.05235987755982988f;
	struct CostData { float dist, depth, angle; int cell; bool use; };
	long long n = 0;
	for (int i = 0; i < num_views; ++i) {
		const dvp_fusion_view& rv = views[i];
		const int num_ngb = rv.num_src;
		std::vector<CostData> diff(num_ngb, CostData{FLT_MAX, FLT_MAX, FLT_MAX, -1, false});
		for (int r = 0; r < rv.height; ++r)
			for (int c = 0; c < rv.width; ++c) {
				const size_t p = (size_t)r * rv.width + c;
				if (used_out) used_out[i][p] = 0;
				if (rv.block && rv.block[p] < 128) continue;
				const float ref_depth = rv.depth[p];
				if (ref_depth <= 0.0) continue;
				const F3 X = point_on_world(c, r, ref_depth, rv.camera);
				for (int j = 0; j < num_ngb; ++j) {
					float e, d, a;
					const int cell = measures(views, i, j, r, c, ref_depth, X, &e, &d, &a);
					// the mask test stands between the bounds test and the depth test (APD.cpp:2075-2079); both skip
					if (cell < 0 || masks[rv.src_views[j]][cell] == 1) continue;
					diff[j].dist = e; diff[j].depth = d; diff[j].angle = a; diff[j].cell = cell;
				}
				for (int k = 2; k <= num_ngb; ++k) {
					int count = 0;
					for (int j = 0; j < num_ngb; ++j) {
						diff[j].use = false;
						const bool ok = mode == 1 ? (diff[j].dist < k * dist_base && diff[j].depth < k * depth_base && diff[j].angle < (k * angle_grad + angle_base))
						                          : (diff[j].dist < k * dist_base && diff[j].depth < k * depth_base);
						if (ok) { count++; diff[j].use = true; }
					}
					if (count >= k) {
						float color[3] = {(float)rv.image[3 * p], (float)rv.image[3 * p + 1], (float)rv.image[3 * p + 2]};
						uint32_t used = 0;
						for (int j = 0; j < num_ngb; ++j)
							if (diff[j].use) {
								used |= 1u << j;
								if (mode == 1) {
									const uint8_t* col = views[rv.src_views[j]].image + 3 * (size_t)diff[j].cell;
									color[0] += (float)col[0]; color[1] += (float)col[1]; color[2] += (float)col[2];
								}
							}
						if (mode == 1) { color[0] /= (count + 1.0f); color[1] /= (count + 1.0f); color[2] /= (count + 1.0f); }
						if (n < capacity) {
							float* o = points + 6 * n;
							o[0] = X.x; o[1] = X.y; o[2] = X.z; o[3] = color[0]; o[4] = color[1]; o[5] = color[2];
						}
						++n;
						if (used_out) used_out[i][p] = used;
						masks[i][p] = 1;
						break;
					}
				}
			}
	}
	return n;
}

// ExportPointCloud's colour conversion (APD.cpp:868-871): static_cast<uchar>(float).
void fusion_cpu_ply_colors(const float* points, long long n, uint8_t* bgr) {
	for (long long i = 0; i < n; ++i)
		for (int k = 0; k < 3; ++k) bgr[3 * i + k] = (uint8_t)points[6 * i + 3 + k];
}

}  // extern "C"
