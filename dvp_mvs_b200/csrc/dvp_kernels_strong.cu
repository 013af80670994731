// dvp_kernels_strong.cu — kernels K1, K6, K7/K8, K12, K13/K14, K15, K16 of the PatchMatch sequence
// (reference APD.cu:1258-1309, 2010-2737, 3127-3328, 3892-4139), rewritten for sm_100a.
#include "dvp_strong.cuh"
#include "dvp_launch.h"
#include <cfloat>
#include <mutex>

namespace dvp {

__constant__ int c_dir[8][2] = {{0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {1, 1}, {-1, 1}, {1, -1}};

// ------------------------------------------------------------------------------------------------------
// per-view constants.  Same expressions as the preamble of ComputeHomography (APD.cu:681-707) and the
// R_c product of GenerateRandomNormal_YZL (APD.cu:540-542) so that rounding/contraction is unchanged.
__global__ void k_setup_views(const dvp_camera* cams, ViewConst* views, int S) {
	const int v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= S) return;
	const dvp_camera ref_camera = cams[0];
	const dvp_camera src_camera = cams[v + 1];
	ViewConst vc;
	float ref_C[3], src_C[3];
	ref_C[0] = -(ref_camera.R[0] * ref_camera.t[0] + ref_camera.R[3] * ref_camera.t[1] + ref_camera.R[6] * ref_camera.t[2]);
	ref_C[1] = -(ref_camera.R[1] * ref_camera.t[0] + ref_camera.R[4] * ref_camera.t[1] + ref_camera.R[7] * ref_camera.t[2]);
	ref_C[2] = -(ref_camera.R[2] * ref_camera.t[0] + ref_camera.R[5] * ref_camera.t[1] + ref_camera.R[8] * ref_camera.t[2]);
	src_C[0] = -(src_camera.R[0] * src_camera.t[0] + src_camera.R[3] * src_camera.t[1] + src_camera.R[6] * src_camera.t[2]);
	src_C[1] = -(src_camera.R[1] * src_camera.t[0] + src_camera.R[4] * src_camera.t[1] + src_camera.R[7] * src_camera.t[2]);
	src_C[2] = -(src_camera.R[2] * src_camera.t[0] + src_camera.R[5] * src_camera.t[1] + src_camera.R[8] * src_camera.t[2]);
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j)
			vc.R_rel[i * 3 + j] = src_camera.R[i * 3 + 0] * ref_camera.R[j * 3 + 0] + src_camera.R[i * 3 + 1] * ref_camera.R[j * 3 + 1] +
			                      src_camera.R[i * 3 + 2] * ref_camera.R[j * 3 + 2];
	float C_rel[3];
	C_rel[0] = (ref_C[0] - src_C[0]);
	C_rel[1] = (ref_C[1] - src_C[1]);
	C_rel[2] = (ref_C[2] - src_C[2]);
	for (int i = 0; i < 3; ++i)
		vc.t_rel[i] = src_camera.R[i * 3 + 0] * C_rel[0] + src_camera.R[i * 3 + 1] * C_rel[1] + src_camera.R[i * 3 + 2] * C_rel[2];
	for (int i = 0; i < 9; ++i) { vc.sK[i] = src_camera.K[i]; vc.sR[i] = src_camera.R[i]; }
	for (int i = 0; i < 3; ++i) { vc.st[i] = src_camera.t[i]; vc.sc[i] = src_camera.c[i]; }
	// R_c = ref.R * transpose(src.R), accumulated from zero like matMul3x3 (APD.cu:3-12)
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j) {
			float acc = 0;
			for (int k = 0; k < 3; ++k) acc += ref_camera.R[i * 3 + k] * src_camera.R[j * 3 + k];
			vc.R_c[i * 3 + j] = acc;
		}
	vc.pad = 0.f;
	views[v] = vc;
}

// ------------------------------------------------------------------------------------------------------
// K1: curand_init(seed, subsequence = y, offset = x) for every pixel (APD.cu:1258-1271).  The reference
// pays a full skip-ahead per pixel; the state at (x, y) is the state at (x-1, y) advanced by one draw, so
// each thread jumps to the start of a 32-pixel run once and then steps.
__global__ void __launch_bounds__(128) k_init_rng(const __grid_constant__ KArgs a, unsigned long long seed) {
	const int seg_per_row = (a.W + 31) / 32;
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= seg_per_row * a.H) return;
	const int y = t / seg_per_row;
	const int x0 = (t - y * seg_per_row) * 32;
	Rng r;
	curand_init(seed, (unsigned long long)y, (unsigned long long)x0, &r.st);
	const int x1 = min(x0 + 32, a.W);
	for (int x = x0; x < x1; ++x) {
		r.store(a.rng, a.N, y * a.W + x);
		(void)curand(&r.st);
	}
}

// ------------------------------------------------------------------------------------------------------
// K6 RandomInitialization (APD.cu:1273-1309) with ComputeMultiViewInitialCost[andSelectedViews]
// (APD.cu:1115-1194).
__global__ void __launch_bounds__(256, 3) k_random_init(const __grid_constant__ KArgs a) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	constexpr int T = 256;   // compile-time stride: shared-memory offsets become immediates
	const int tid = threadIdx.y * blockDim.x + threadIdx.x;
	float2* wt = reinterpret_cast<float2*>(smem_raw) + tid;
	int x, y; full_grid_pixel(x, y);
	if (x >= a.W || y >= a.H) return;
	const int center = y * a.W + x;
	RefPatch rp;
	rp.prepare(a, x, y, a.prm.use_radius ? a.radius[center] : a.prm.strong_radius, wt, T);

	if (a.prm.state == DVP_FIRST_INIT) {
		float4 pl = a.planes[center];
		if (pl.w > a.prm.depth_max || pl.w < a.prm.depth_min) {
			Rng rng; rng.load(a.rng, a.N, center);
			const float depth = rng.uniform() * (a.prm.depth_max - a.prm.depth_min) + a.prm.depth_min;
			pl = random_normal(a, x, y, rng, depth, a.selected[center]);
			pl.w = get_distance2origin(a.ref, x, y, depth, pl);
			a.planes[center] = pl;
			rng.store(a.rng, a.N, center);
		}  // else: left as (world normal, depth) — bug B20, reproduced
		// cost over all views; keep the top_k cheapest valid views
		float cv[kMaxImages];     // sorted copy (local memory; K6 runs once per pass)
		float cv_copy[kMaxImages];
		int cost_count = 0, num_valid = 0;
		for (int v = 0; v < a.S; ++v) {
			const float c = ncc_cost<kWideRB>(a, a.views[v], a.tex_img[v + 1], x, y, pl, rp, wt, T);
			cv[v] = c; cv_copy[v] = c; cost_count++;
			if (c < kCostMax) num_valid++;
		}
		for (int i = 1; i < cost_count; i++) {  // insertion sort, ascending (sort_small, APD.cu:114-123)
			const float tmp = cv[i];
			int j;
			for (j = i; j >= 1 && tmp < cv[j - 1]; j--) cv[j] = cv[j - 1];
			cv[j] = tmp;
		}
		uint32_t sel = 0;
		const int top_k = min(num_valid, a.prm.top_k);
		float out_cost = kCostMax;
		if (top_k > 0) {
			float cost = 0.0f;
			for (int i = 0; i < top_k; ++i) cost += cv[i];
			const float thr = cv[top_k - 1];
			for (int i = 0; i < a.S; ++i)
				if (cv_copy[i] <= thr) sel |= (1u << i);
			out_cost = cost / top_k;
		}
		a.selected[center] = sel;
		a.costs[center] = out_cost;
	} else {
		float4 pl = a.planes[center];
		pl = normal_to_refcam(a.ref, pl);
		const float depth = pl.w;
		pl.w = get_distance2origin(a.ref, x, y, depth, pl);
		a.planes[center] = pl;
		uint32_t sel = a.selected[center];
		int cost_count = 0;
		float cost = 0.0f;
		for (int v = 0; v < a.S; ++v) {
			if (is_set(sel, v)) {
				const float c = ncc_cost<kWideRB>(a, a.views[v], a.tex_img[v + 1], x, y, pl, rp, wt, T);
				if (c < kCostMax) { cost_count++; cost += c; }
				else unset_bit_ref(&sel, v);  // B1
			}
		}
		a.selected[center] = sel;
		a.costs[center] = (cost_count == 0) ? kCostMax : cost / cost_count;
	}
}

// ------------------------------------------------------------------------------------------------------
// K7/K8: red-black propagation sweep for non-WEAK pixels (edge-adaptive branch, params.use_edge).
// shared memory per thread: 36 float2 (w, w r) + 9*S floats (8 direction cost vectors + 1 spare) + 8 u16 ladder offsets.
// FORCE_D4 (parity instrumentation, never launched by dvp_run; see dvp_debug_race_explain in include/dvp_mvs.h): direction
// 4 — the one whose ladder reads pixels of the colour being written — takes its candidate from ladder offset
// `force.m`; the candidate's plane is assembled component by component from a `before` and an `after` snapshot for each of
// the three reads the reference makes of it (scoring, APD.cu:2084/2133; depth test and copy at acceptance, APD.cu:2559-2563),
// the reference's SASS reading planes with 32-bit loads.  Outputs go to shadow buffers: the state is left untouched.
__device__ __forceinline__ float4 mix_planes(const float4 b, const float4 a, unsigned mask) {
	return make_float4((mask & 1) ? a.x : b.x, (mask & 2) ? a.y : b.y, (mask & 4) ? a.z : b.z, (mask & 8) ? a.w : b.w);
}
template <bool FORCE_D4>
__global__ void __launch_bounds__(kSweepThreads, kSweepMinBlocks) k_strong_sweep(const __grid_constant__ KArgs a, int iter, int red, int yy_limit, const D4Force force) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	constexpr int T = kSweepThreads;   // compile-time stride: shared-memory offsets become immediates
	const int tid = threadIdx.y * blockDim.x + threadIdx.x;
	float2* wt = kSweepRW ? reinterpret_cast<float2*>(reinterpret_cast<float*>(smem_raw) + tid)   // RW: a column of 36 floats (r)
	                      : reinterpret_cast<float2*>(smem_raw) + tid;                              // else 36 (w, w r) pairs
	constexpr size_t kTable = (size_t)kHoistSamples * T * (kSweepRW ? sizeof(float) : sizeof(float2));
	float* cost_arr = reinterpret_cast<float*>(smem_raw + kTable) + tid;
	// winning ladder offset per direction, in pixels along the direction (step * step_len <= 21 * max(H,W)/660 < 2^16)
	uint16_t* pos_arr = reinterpret_cast<uint16_t*>(smem_raw + kTable + (size_t)9 * a.S * T * sizeof(float)) + tid;
	const int S = a.S, W = a.W, H = a.H;

	int x = blockIdx.x * blockDim.x + threadIdx.x;
	int y;
	if (FORCE_D4 && force.pixel_list) {   // instrumentation: a list of pixels of this colour instead of the whole half grid
		const int i = blockIdx.x * T + tid;
		if (i >= force.list_count) return;
		const int c = force.pixel_list[i];
		x = c % W; y = c / W;
	} else {
		const int yy = blockIdx.y * blockDim.y + threadIdx.y;
		y = 2 * yy + ((x & 1) ^ red);  // black: even x -> even row; red: the other colour (APD.cu:3129-3136)
		if (x >= W || y >= H || yy >= yy_limit) return;
	}
	const int center = y * W + x;
	if (a.weak[center] == DVP_WEAK) return;
	// where results go: the state itself, or (instrumentation) shadow buffers
	float4* const o_planes = FORCE_D4 ? force.out_planes : a.planes;
	float* const o_costs = FORCE_D4 ? force.out_costs : a.costs;
	uint32_t* const o_selected = FORCE_D4 ? force.out_selected : a.selected;
	uint8_t* const o_view_weight = FORCE_D4 ? force.out_view_weight : a.view_weight;
	uint32_t* const o_rng = FORCE_D4 ? force.out_rng : a.rng;

	RefPatch rp;
	rp.prepare<kSweepRW>(a, x, y, a.prm.use_radius ? a.radius[center] : a.prm.strong_radius, wt, T);

	// `float cost_array[8][32] = {2.0f}` : element [0][0] is 2, everything else 0 (bug B2, reproduced)
	for (int s = 0; s < 8; ++s)
		for (int v = 0; v < S; ++v) cost_arr[(s * S + v) * T] = 0.0f;
	cost_arr[0] = 2.0f;
	uint32_t flag = 0;

	const bool on_edge = a.edge[center] != 0;
	const short2* edge_neigh = a.edge_neigh + (size_t)center * DVP_EDGE_NEIGH_NUM;
	const float max_edge_dist = DVP_MAX(H, W) / 30.0f;
	const int min_step_len = 2;
	const float good_threshold = 0.8f * expf((iter) * (iter) / (-90.0f));
	const float bad_threshold = 1.2f;

#pragma unroll 1
	for (int d = 0; d < 8; ++d) {
		const int dx = c_dir[d][0], dy = c_dir[d][1];
		const int sx = 5 * dx, sy = 5 * dy;
		int fx = 0, fy = 0;
		if (d > 4) { if (d % 2) fx = dx; else fy = dy; }  // colour fix on directions 5,6,7 only (B6: 4 is racy)
		if (FORCE_D4 && d == 4) {
			const int tx = x + sx - force.m, ty = y + sy - force.m;
			if (tx >= 0 && ty >= 0) {
				flag |= 1u << d;
				pos_arr[d * T] = (uint16_t)force.m;
				const float4 pb = force.before[tx + ty * W], pa = force.after[tx + ty * W];
				for (int v = 0; v < S; ++v) {
					const float4 pl = mix_planes(pb, pa, (unsigned)(force.ncc_masks >> (4 * (v & 15))) & 15u);
					cost_arr[(d * S + v) * T] = ncc_cost<kSweepRB, kSweepRW>(a, a.views[v], a.tex_img[v + 1], x, y, pl, rp, wt, T);
				}
			}
			continue;
		}
		// ---- edge-adaptive ladder (APD.cu:2053-2087) ----
		{
			const short2 edge_pt = edge_neigh[d];
			const int ex = edge_pt.x - x, ey = edge_pt.y - y;
			float dist = (float)sqrt((double)(ex * ex) + (double)(ey * ey));
			if (d >= 4) dist = (float)((double)dist / sqrt(2.0));
			if (on_edge) {
				dist = 11 * min_step_len;
			} else if (edge_pt.y == -1 || dist >= max_edge_dist) {  // `!edge_pt.x == -1` is always false (B5)
				dist = max_edge_dist;
				if (d >= 4) dist = (float)((double)dist / sqrt(2.0));
			}
			const int step_num = DVP_MIN(DVP_MAX(11, (int)(1.0f * dist / min_step_len)), 22);
			int step_len = DVP_MAX((int)(1.0f * dist / step_num), min_step_len);
			if (d < 4 && step_len % 2 == 1) step_len -= 1;
			int min_pos = -1, min_k = 0;
			float min_cost = FLT_MAX;
			for (int step = 0; step < step_num; ++step) {
				const int tx = x + sx + step * step_len * dx + fx, ty = y + sy + step * step_len * dy + fy;
				if (!(tx >= 0 && ty >= 0 && tx < W && ty < H)) continue;
				const int tc = tx + ty * W;
				const float c = a.costs[tc];
				if (min_cost > c) { min_pos = tc; min_k = step * step_len; min_cost = c; }
			}
			if (min_cost < FLT_MAX) {
				flag |= 1u << d;
				pos_arr[d * T] = (uint16_t)min_k;
				const float4 pl = a.planes[min_pos];
				for (int v = 0; v < S; ++v)
					cost_arr[(d * S + v) * T] = ncc_cost<kSweepRB, kSweepRW>(a, a.views[v], a.tex_img[v + 1], x, y, pl, rp, wt, T);
			}
		}
		// ---- fixed 11 x 2 px ladder for non-edge pixels; keep the better of the two (APD.cu:2090-2140) ----
		if (!on_edge) {
			const bool has_before = (flag >> d) & 1;
			int min_pos = -1, min_k = 0;
			float min_cost = FLT_MAX;
			for (int step = 0; step < 11; ++step) {
				const int tx = x + sx + step * min_step_len * dx + fx, ty = y + sy + step * min_step_len * dy + fy;
				if (!(tx >= 0 && ty >= 0 && tx < W && ty < H)) continue;
				const int tc = tx + ty * W;
				const float c = a.costs[tc];
				if (min_cost > c) { min_pos = tc; min_k = step * min_step_len; min_cost = c; }
			}
			// Same winner as the adaptive ladder (always the case within 24 px of an edge, where both ladders are
			// 11 x 2 px): the plane is the one just scored, every c1 equals its c0, the tallies tie and nothing is
			// replaced (APD.cu:2126) — so the S NCCs are not recomputed.
			const bool same_winner = has_before && min_k == (int)pos_arr[d * T];
			if (min_cost < FLT_MAX && !same_winner) {
				flag |= 1u << d;
				const float4 pl = a.planes[min_pos];
				int good0 = 0, good1 = 0, bad0 = 0, bad1 = 0;
				for (int v = 0; v < S; ++v) {
					const float c1 = ncc_cost<kSweepRB, kSweepRW>(a, a.views[v], a.tex_img[v + 1], x, y, pl, rp, wt, T);
					cost_arr[(8 * S + v) * T] = c1;
					const float c0 = cost_arr[(d * S + v) * T];
					if (c0 < good_threshold) good0++;
					if (c0 > bad_threshold) bad0++;
					if (c1 < good_threshold) good1++;
					if (c1 > bad_threshold) bad1++;
				}
				if (!has_before || good1 > good0 || (good1 == good0 && bad1 < bad0)) {
					pos_arr[d * T] = (uint16_t)min_k;
					for (int v = 0; v < S; ++v) cost_arr[(d * S + v) * T] = cost_arr[(8 * S + v) * T];
				}
			}
		}
	}

	// ---- multi-hypothesis joint view selection (APD.cu:2462-2530) ----
	Rng rng; rng.load(a.rng, a.N, center);
	ViewWeights vw; vw.clear();
	{
		// priors from the 4-neighbours, guarded by flag[0], flag[2], flag[4], flag[6] (B16, reproduced;
		// `selected` has one padded row on each side so the border reads are defined)
		const int npos[4] = {center - W, center + W, center - 1, center + 1};
		uint32_t nsel[4];
#pragma unroll
		for (int i = 0; i < 4; ++i) nsel[i] = ((flag >> (2 * i)) & 1) ? a.selected[npos[i]] : 0u;
		const float cost_threshold = 0.8 * expf((iter) * (iter) / (-90.0f));
		float* probs = cost_arr + (size_t)8 * S * T;  // spare slot
		for (int i = 0; i < S; i++) {
			float prior = 0.0f;
#pragma unroll
			for (int k = 0; k < 4; ++k)
				if ((flag >> (2 * k)) & 1) prior += is_set(nsel[k], i) ? 0.9f : 0.1f;
			float count = 0;
			int count_false = 0;
			float tmpw = 0;
			for (int j = 0; j < 8; j++) {
				const float c = cost_arr[(j * S + i) * T];
				if (c < cost_threshold) {
					tmpw += expf(c * c / (-0.18f));
					count++;
				}
				if (c > 1.2f) count_false++;
			}
			float prob = 0.0f;
			if (count > 2 && count_false < 3) prob = tmpw / count;
			else if (count_false < 3) prob = expf(cost_threshold * cost_threshold / (-0.32f));
			prob = prob * prior;
			probs[i * T] = prob;
		}
		// TransformPDFToCDF (APD.cu:356-370); all-zero probabilities give NaN here (B18, reproduced)
		float prob_sum = 0.0f;
		for (int i = 0; i < S; ++i) prob_sum += probs[i * T];
		const float inv_prob_sum = 1.0f / prob_sum;
		float cum_prob = 0.0f;
		for (int i = 0; i < S; ++i) {
			const float prob = probs[i * T] * inv_prob_sum;
			cum_prob += prob;
			probs[i * T] = cum_prob;
		}
		for (int sample = 0; sample < 15; ++sample) {
			const float rand_prob = rng.uniform() - FLT_EPSILON;
			for (int image_id = 0; image_id < S; ++image_id) {
				if (probs[image_id * T] > rand_prob) { vw.inc(image_id); break; }
			}
		}
	}
	vw.store(o_view_weight + (size_t)center * DVP_MAX_IMAGES);

	uint32_t temp_selected = 0;
	float weight_norm = 0;
	for (int i = 0; i < S; ++i) {
		const int wv = vw.get(i);
		if (wv > 0) { temp_selected |= 1u << i; weight_norm += wv; }
	}

	// ---- aggregated candidate costs and the current plane (APD.cu:2532-2567) ----
	int min_cost_idx = 0;
	float min_final = 0.f;
	for (int i = 0; i < 8; ++i) {
		float fc = 0.0f;
		for (int j = 0; j < S; ++j) {
			const int wv = vw.get(j);
			if (wv > 0) fc += wv * cost_arr[(i * S + j) * T];
		}
		fc /= weight_norm;
		if (i == 0 || fc <= min_final) { min_final = fc; min_cost_idx = i; }  // FindMinCostIndex: '<=' keeps the last
	}

	float4 plane_now = a.planes[center];
	float cost_now = 0.0f;
	for (int v = 0; v < S; ++v) {
		const int wv = vw.get(v);
		if (wv > 0) {  // zero-weight views contribute exactly 0 in the reference
			const float c = ncc_cost<kSweepRB, kSweepRW>(a, a.views[v], a.tex_img[v + 1], x, y, plane_now, rp, wt, T);
			cost_now += wv * c;
		}
	}
	cost_now /= weight_norm;
	const float cost_stored = cost_now;  // costs[center] = cost_now (APD.cu:2554)
	float depth_now = depth_from_plane(a.ref, plane_now, x, y);
	uint32_t sel_now = a.selected[center];

	if ((flag >> min_cost_idx) & 1) {
		int cx, cy;
		{
			const int dx = c_dir[min_cost_idx][0], dy = c_dir[min_cost_idx][1], k = pos_arr[min_cost_idx * T];
			int fx = 0, fy = 0;
			if (min_cost_idx > 4) { if (min_cost_idx % 2) fx = dx; else fy = dy; }
			cx = x + 5 * dx + k * dx + fx; cy = y + 5 * dy + k * dy + fy;
		}
		float4 cand = a.planes[cx + cy * W];
		float4 cand_for_depth = cand;   // the reference reads the winner's plane twice here: for the depth test and for the copy
		if (FORCE_D4 && min_cost_idx == 4) {
			const float4 b = force.before[cx + cy * W], n = force.after[cx + cy * W];
			cand_for_depth = mix_planes(b, n, force.dep_mask);
			cand = mix_planes(b, n, force.acc_mask);
		}
		const float depth_before = depth_from_plane(a.ref, cand_for_depth, x, y);
		if (depth_before >= a.prm.depth_min && depth_before <= a.prm.depth_max && min_final < cost_now) {
			depth_now = depth_before;
			plane_now = cand;
			cost_now = min_final;
			sel_now = temp_selected;
			o_selected[center] = temp_selected;
		} else if (FORCE_D4) {
			o_selected[center] = sel_now;   // shadow buffers get every output of a processed pixel, changed or not
		}
	} else if (FORCE_D4) {
		o_selected[center] = sel_now;
	}

	const float4 plane_before_refine_write = FORCE_D4 ? a.planes[center] : plane_now;
	refine_strong(a, x, y, &plane_now, &depth_now, &cost_now, rng, vw, weight_norm, sel_now, rp, wt, T);
	rng.store(o_rng, a.N, center);

	if (a.prm.state == DVP_REFINE_INIT) {
		if (cost_now < cost_stored - 0.1) {
			o_costs[center] = cost_now;
			o_planes[center] = plane_now;
		} else {
			o_costs[center] = cost_stored;
			if (FORCE_D4) o_planes[center] = plane_before_refine_write;
		}
	} else {
		o_costs[center] = cost_now;
		o_planes[center] = plane_now;
	}
}

// ------------------------------------------------------------------------------------------------------
// K7/K8 as dvp_run launches them: two kernels per colour.
//   k_sweep_score   the 8-direction candidate search and the 16 x S (minus duplicates) NCCs that score the candidates — 70 %
//                   of the sweep's texture fetches, no RNG, no refinement state: 256-thread blocks, 3 per SM (24 warps / SM,
//                   against 12 for the single kernel, whose per-thread cost columns and RNG / refinement registers cap it).
//                   Candidate costs, the winning ladder offsets and the direction flags go to a scratch area in HBM
//                   ((9 S + 9) words per pixel of the colour, written and read once, coalesced).
//   k_sweep_update  view selection, the current plane's cost, acceptance, refinement, write-back: 128-thread blocks, 4 per SM.
// Results are those of the reference kernel for ONE outcome of its direction-4 race, the same on every run: every pixel
// of the colour scores its candidates before any pixel of the colour is rewritten, and direction 4's winner is kept as it
// was scored (the reference reads it again at acceptance, APD.cu:2559-2563).  tests/test_gpu_parity.py checks that this is
// an outcome the reference can produce (dvp_debug_race_explain) and that race-free launches equal the reference bit for bit.
struct SweepScratch {
	float* cost;        // [(9 S)][slots]: rows d * S + v = cost of direction d's candidate in view v; rows 8 S + v = spare (second ladder, then the CDF)
	uint32_t* flag;     // [slots] bit d: direction d has a candidate
	uint4* pos;         // [slots] eight 16-bit ladder offsets of the winners
	float4* d4_plane;   // [slots] plane of direction 4's winner as scored
	int slots;          // W * ceil(H / 2): pixel (x, y) of the colour <-> slot (y >> 1) * W + x
};

__global__ void __launch_bounds__(256, 3) k_sweep_score(const __grid_constant__ KArgs a, int iter, int red, int yy_limit, const SweepScratch sc) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	constexpr int T = 256;
	const int tid = threadIdx.y * blockDim.x + threadIdx.x;
	float2* wt = reinterpret_cast<float2*>(smem_raw) + tid;
	const int S = a.S, W = a.W, H = a.H;
	int x, y; sweep_pixel(red, x, y);
	const int yy = y >> 1;
	if (x < 0 || y < 0 || x >= W || y >= H || yy >= yy_limit) return;
	const int center = y * W + x;
	if (a.weak[center] == DVP_WEAK) return;
	const size_t slot = (size_t)yy * W + x;
	float* const cost = sc.cost + slot;
	const size_t row = (size_t)sc.slots;
#define SCOST(i) cost[(size_t)(i) * row]

	RefPatch rp;
	rp.prepare<false>(a, x, y, a.prm.use_radius ? a.radius[center] : a.prm.strong_radius, wt, T);

	// `float cost_array[8][32] = {2.0f}` : element [0][0] is 2, everything else 0 (bug B2, reproduced)
	for (int i = 0; i < 8 * S; ++i) SCOST(i) = 0.0f;
	SCOST(0) = 2.0f;
	uint32_t flag = 0;
	uint32_t posw[4] = {0, 0, 0, 0};   // eight 16-bit offsets
	auto set_pos = [&](int d, int k) { posw[d >> 1] = (posw[d >> 1] & ~(0xffffu << (16 * (d & 1)))) | ((uint32_t)k << (16 * (d & 1))); };
	auto get_pos = [&](int d) -> int { return (int)((posw[d >> 1] >> (16 * (d & 1))) & 0xffffu); };
	float4 d4_plane = make_float4(0.f, 0.f, 0.f, 0.f);

	const bool on_edge = a.edge[center] != 0;
	const short2* edge_neigh = a.edge_neigh + (size_t)center * DVP_EDGE_NEIGH_NUM;
	const float max_edge_dist = DVP_MAX(H, W) / 30.0f;
	const int min_step_len = 2;
	const float good_threshold = 0.8f * expf((iter) * (iter) / (-90.0f));
	const float bad_threshold = 1.2f;

#pragma unroll 1
	for (int d = 0; d < 8; ++d) {
		const int dx = c_dir[d][0], dy = c_dir[d][1];
		const int sx = 5 * dx, sy = 5 * dy;
		int fx = 0, fy = 0;
		if (d > 4) { if (d % 2) fx = dx; else fy = dy; }  // colour fix on directions 5,6,7 only (B6)
		// ---- edge-adaptive ladder (APD.cu:2053-2087) ----
		{
			const short2 edge_pt = edge_neigh[d];
			const int ex = edge_pt.x - x, ey = edge_pt.y - y;
			float dist = (float)sqrt((double)(ex * ex) + (double)(ey * ey));
			if (d >= 4) dist = (float)((double)dist / sqrt(2.0));
			if (on_edge) {
				dist = 11 * min_step_len;
			} else if (edge_pt.y == -1 || dist >= max_edge_dist) {  // `!edge_pt.x == -1` is always false (B5)
				dist = max_edge_dist;
				if (d >= 4) dist = (float)((double)dist / sqrt(2.0));
			}
			const int step_num = DVP_MIN(DVP_MAX(11, (int)(1.0f * dist / min_step_len)), 22);
			int step_len = DVP_MAX((int)(1.0f * dist / step_num), min_step_len);
			if (d < 4 && step_len % 2 == 1) step_len -= 1;
			int min_pos = -1, min_k = 0;
			float min_cost = FLT_MAX;
#pragma unroll 4   // four cost reads of the ladder in flight
			for (int step = 0; step < step_num; ++step) {
				const int tx = x + sx + step * step_len * dx + fx, ty = y + sy + step * step_len * dy + fy;
				if (!(tx >= 0 && ty >= 0 && tx < W && ty < H)) continue;
				const int tc = tx + ty * W;
				const float c = a.costs[tc];
				if (min_cost > c) { min_pos = tc; min_k = step * step_len; min_cost = c; }
			}
			if (min_cost < FLT_MAX) {
				flag |= 1u << d;
				set_pos(d, min_k);
				const float4 pl = a.planes[min_pos];
				if (d == 4) d4_plane = pl;
				for (int v = 0; v < S; ++v)
					SCOST(d * S + v) = ncc_cost<kWideRB>(a, a.views[v], a.tex_img[v + 1], x, y, pl, rp, wt, T);
			}
		}
		// ---- fixed 11 x 2 px ladder for non-edge pixels; keep the better of the two (APD.cu:2090-2140) ----
		if (!on_edge) {
			const bool has_before = (flag >> d) & 1;
			int min_pos = -1, min_k = 0;
			float min_cost = FLT_MAX;
			for (int step = 0; step < 11; ++step) {
				const int tx = x + sx + step * min_step_len * dx + fx, ty = y + sy + step * min_step_len * dy + fy;
				if (!(tx >= 0 && ty >= 0 && tx < W && ty < H)) continue;
				const int tc = tx + ty * W;
				const float c = a.costs[tc];
				if (min_cost > c) { min_pos = tc; min_k = step * min_step_len; min_cost = c; }
			}
			// same winner as the adaptive ladder: same plane, same costs, nothing is replaced (APD.cu:2126) — not rescored
			const bool same_winner = has_before && min_k == get_pos(d);
			if (min_cost < FLT_MAX && !same_winner) {
				flag |= 1u << d;
				const float4 pl = a.planes[min_pos];
				int good0 = 0, good1 = 0, bad0 = 0, bad1 = 0;
				for (int v = 0; v < S; ++v) {
					const float c0 = SCOST(d * S + v);   // read ahead of the NCC: its latency hides behind the fetches
					const float c1 = ncc_cost<kWideRB>(a, a.views[v], a.tex_img[v + 1], x, y, pl, rp, wt, T);
					SCOST(8 * S + v) = c1;
					if (c0 < good_threshold) good0++;
					if (c0 > bad_threshold) bad0++;
					if (c1 < good_threshold) good1++;
					if (c1 > bad_threshold) bad1++;
				}
				if (!has_before || good1 > good0 || (good1 == good0 && bad1 < bad0)) {
					set_pos(d, min_k);
					if (d == 4) d4_plane = pl;
					for (int v = 0; v < S; ++v) SCOST(d * S + v) = SCOST(8 * S + v);
				}
			}
		}
	}
	sc.flag[slot] = flag;
	sc.pos[slot] = make_uint4(posw[0], posw[1], posw[2], posw[3]);
	sc.d4_plane[slot] = d4_plane;
#undef SCOST
}

#ifndef DVP_UPDATE_MIN_BLOCKS
#define DVP_UPDATE_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(128, DVP_UPDATE_MIN_BLOCKS) k_sweep_update(const __grid_constant__ KArgs a, int iter, int red, int yy_limit, const SweepScratch sc) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	constexpr int T = 128;
	const int tid = threadIdx.y * blockDim.x + threadIdx.x;
	float2* wt = reinterpret_cast<float2*>(smem_raw) + tid;
	const int S = a.S, W = a.W, H = a.H;
	int x, y; sweep_pixel(red, x, y);
	const int yy = y >> 1;
	if (x < 0 || y < 0 || x >= W || y >= H || yy >= yy_limit) return;
	const int center = y * W + x;
	if (a.weak[center] == DVP_WEAK) return;
	const size_t slot = (size_t)yy * W + x;
	float* const cost = sc.cost + slot;
	const size_t row = (size_t)sc.slots;
#define SCOST(i) cost[(size_t)(i) * row]

	RefPatch rp;
	rp.prepare<false>(a, x, y, a.prm.use_radius ? a.radius[center] : a.prm.strong_radius, wt, T);
	const uint32_t flag = sc.flag[slot];
	const uint4 posq = sc.pos[slot];
	const uint32_t posw[4] = {posq.x, posq.y, posq.z, posq.w};

	// ---- multi-hypothesis joint view selection (APD.cu:2462-2530) ----
	Rng rng; rng.load(a.rng, a.N, center);
	ViewWeights vw; vw.clear();
	{
		// priors from the 4-neighbours, guarded by flag[0], flag[2], flag[4], flag[6] (B16, reproduced;
		// `selected` has one padded row on each side so the border reads are defined)
		const int npos[4] = {center - W, center + W, center - 1, center + 1};
		uint32_t nsel[4];
#pragma unroll
		for (int i = 0; i < 4; ++i) nsel[i] = ((flag >> (2 * i)) & 1) ? a.selected[npos[i]] : 0u;
		const float cost_threshold = 0.8 * expf((iter) * (iter) / (-90.0f));
		for (int i = 0; i < S; i++) {
			float prior = 0.0f;
#pragma unroll
			for (int k = 0; k < 4; ++k)
				if ((flag >> (2 * k)) & 1) prior += is_set(nsel[k], i) ? 0.9f : 0.1f;
			float count = 0;
			int count_false = 0;
			float tmpw = 0;
			for (int j = 0; j < 8; j++) {
				const float c = SCOST(j * S + i);
				if (c < cost_threshold) {
					tmpw += expf(c * c / (-0.18f));
					count++;
				}
				if (c > 1.2f) count_false++;
			}
			float prob = 0.0f;
			if (count > 2 && count_false < 3) prob = tmpw / count;
			else if (count_false < 3) prob = expf(cost_threshold * cost_threshold / (-0.32f));
			prob = prob * prior;
			SCOST(8 * S + i) = prob;
		}
		// TransformPDFToCDF (APD.cu:356-370); all-zero probabilities give NaN here (B18, reproduced)
		float prob_sum = 0.0f;
		for (int i = 0; i < S; ++i) prob_sum += SCOST(8 * S + i);
		const float inv_prob_sum = 1.0f / prob_sum;
		float cum_prob = 0.0f;
		for (int i = 0; i < S; ++i) {
			const float prob = SCOST(8 * S + i) * inv_prob_sum;
			cum_prob += prob;
			SCOST(8 * S + i) = cum_prob;
		}
		if (S <= 8) {
			// the CDF in registers: the 15 draws below would otherwise walk it in the scratch area, a dependent global load per
			// comparison (8 % of this kernel's stall samples).  First index whose CDF exceeds the draw, as the loop below.
			float cdf[8];
#pragma unroll
			for (int i = 0; i < 8; ++i) cdf[i] = i < S ? SCOST(8 * S + i) : 0.0f;
			for (int sample = 0; sample < 15; ++sample) {
				const float rand_prob = rng.uniform() - FLT_EPSILON;
				int image_id = -1;
#pragma unroll
				for (int i = 7; i >= 0; --i) if (i < S && cdf[i] > rand_prob) image_id = i;
				if (image_id >= 0) vw.inc(image_id);
			}
		} else {
			for (int sample = 0; sample < 15; ++sample) {
				const float rand_prob = rng.uniform() - FLT_EPSILON;
				for (int image_id = 0; image_id < S; ++image_id) {
					if (SCOST(8 * S + image_id) > rand_prob) { vw.inc(image_id); break; }
				}
			}
		}
	}
	vw.store(a.view_weight + (size_t)center * DVP_MAX_IMAGES);

	uint32_t temp_selected = 0;
	float weight_norm = 0;
	for (int i = 0; i < S; ++i) {
		const int wv = vw.get(i);
		if (wv > 0) { temp_selected |= 1u << i; weight_norm += wv; }
	}

	// ---- aggregated candidate costs and the current plane (APD.cu:2532-2567) ----
	int min_cost_idx = 0;
	float min_final = 0.f;
	for (int i = 0; i < 8; ++i) {
		float fc = 0.0f;
		for (int j = 0; j < S; ++j) {
			const int wv = vw.get(j);
			if (wv > 0) fc += wv * SCOST(i * S + j);
		}
		fc /= weight_norm;
		if (i == 0 || fc <= min_final) { min_final = fc; min_cost_idx = i; }  // FindMinCostIndex: '<=' keeps the last
	}

	float4 plane_now = a.planes[center];
	float cost_now = 0.0f;
	for (int v = 0; v < S; ++v) {
		const int wv = vw.get(v);
		if (wv > 0) {  // zero-weight views contribute exactly 0 in the reference
			const float c = ncc_cost<kSweepRB>(a, a.views[v], a.tex_img[v + 1], x, y, plane_now, rp, wt, T);
			cost_now += wv * c;
		}
	}
	cost_now /= weight_norm;
	const float cost_stored = cost_now;  // costs[center] = cost_now (APD.cu:2554)
	float depth_now = depth_from_plane(a.ref, plane_now, x, y);
	uint32_t sel_now = a.selected[center];

	if ((flag >> min_cost_idx) & 1) {
		float4 cand;
		if (min_cost_idx == 4) {
			cand = sc.d4_plane[slot];   // the same-colour candidate, as it was scored (see the header of this pair of kernels)
		} else {
			const int dx = c_dir[min_cost_idx][0], dy = c_dir[min_cost_idx][1];
			const int k = (int)((posw[min_cost_idx >> 1] >> (16 * (min_cost_idx & 1))) & 0xffffu);
			int fx = 0, fy = 0;
			if (min_cost_idx > 4) { if (min_cost_idx % 2) fx = dx; else fy = dy; }
			const int cx = x + 5 * dx + k * dx + fx, cy = y + 5 * dy + k * dy + fy;
			cand = a.planes[cx + cy * W];   // a pixel of the other colour: not written by this launch
		}
		const float depth_before = depth_from_plane(a.ref, cand, x, y);
		if (depth_before >= a.prm.depth_min && depth_before <= a.prm.depth_max && min_final < cost_now) {
			depth_now = depth_before;
			plane_now = cand;
			cost_now = min_final;
			sel_now = temp_selected;
			a.selected[center] = temp_selected;
		}
	}

	refine_strong(a, x, y, &plane_now, &depth_now, &cost_now, rng, vw, weight_norm, sel_now, rp, wt, T);
	rng.store(a.rng, a.N, center);

	if (a.prm.state == DVP_REFINE_INIT) {
		if (cost_now < cost_stored - 0.1) {
			a.costs[center] = cost_now;
			a.planes[center] = plane_now;
		} else {
			a.costs[center] = cost_stored;
		}
	} else {
		a.costs[center] = cost_now;
		a.planes[center] = plane_now;
	}
#undef SCOST
}

// ---- instrumentation: which pixels of an observed sweep result does a forced choice reproduce? -----------------------
// expected rand is the canonical [N][6] layout; the shadow rng is the engine's 6 SoA planes
__device__ __forceinline__ bool pixel_equal(const KArgs& a, int c, const float4* pl_a, const float* co_a, const uint32_t* se_a, const uint8_t* vw_a, const uint32_t* rng_soa,
                                            const float4* pl_e, const float* co_e, const uint32_t* se_e, const uint8_t* vw_e, const uint32_t* rng_aos) {
	const uint4 p0 = reinterpret_cast<const uint4*>(pl_a)[c], p1 = reinterpret_cast<const uint4*>(pl_e)[c];
	bool ok = p0.x == p1.x && p0.y == p1.y && p0.z == p1.z && p0.w == p1.w;
	ok = ok && __float_as_uint(co_a[c]) == __float_as_uint(co_e[c]) && se_a[c] == se_e[c];
	const uint4* va = reinterpret_cast<const uint4*>(vw_a + (size_t)c * DVP_MAX_IMAGES);
	const uint4* ve = reinterpret_cast<const uint4*>(vw_e + (size_t)c * DVP_MAX_IMAGES);
	for (int k = 0; k < 2; ++k) { const uint4 u = va[k], v = ve[k]; ok = ok && u.x == v.x && u.y == v.y && u.z == v.z && u.w == v.w; }
	for (int k = 0; k < 6; ++k) ok = ok && rng_soa[(size_t)k * a.N + c] == rng_aos[(size_t)c * 6 + k];
	return ok;
}
__device__ __forceinline__ bool sweep_processes(const KArgs& a, int c, int red, int yy_limit) {
	const int x = c % a.W, y = c / a.W;
	if (((x + y) & 1) != red) return false;
	if ((y >> 1) >= yy_limit) return false;
	return a.weak[c] != DVP_WEAK;
}
// pixels the launch does not process must equal the pre-launch state; processed pixels start unexplained
__global__ void k_explain_init(const __grid_constant__ KArgs a, int red, int yy_limit, const RaceExpected e, uint8_t* explained) {
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= a.N) return;
	if (sweep_processes(a, c, red, yy_limit)) { explained[c] = 0; return; }
	explained[c] = pixel_equal(a, c, a.planes, a.costs, a.selected, a.view_weight, a.rng, e.planes, e.costs, e.selected, e.view_weight, e.rand) ? 1 : 0;
}
__global__ void k_explain_compare(const __grid_constant__ KArgs a, int red, int yy_limit, const D4Force f, const RaceExpected e, uint8_t* explained) {
	int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (f.pixel_list) { if (c >= f.list_count) return; c = f.pixel_list[c]; }
	else if (c >= a.N || !sweep_processes(a, c, red, yy_limit)) return;
	if (explained[c]) return;
	if (pixel_equal(a, c, f.out_planes, f.out_costs, f.out_selected, f.out_view_weight, f.out_rng, e.planes, e.costs, e.selected, e.view_weight, e.rand)) explained[c] = 1;
}
__global__ void k_explain_collect(const __grid_constant__ KArgs a, int red, int yy_limit, const uint8_t* explained, int* list, int* count, int cap) {
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= a.N || explained[c] || !sweep_processes(a, c, red, yy_limit)) return;
	const int i = atomicAdd(count, 1);
	if (i < cap) list[i] = c;
}

// ------------------------------------------------------------------------------------------------------
// K12 GetDepthandNormal (APD.cu:3167-3182): plane offset -> depth, camera normal -> world normal.
__global__ void __launch_bounds__(256) k_depth_normal(const __grid_constant__ KArgs a) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= a.W || y >= a.H) return;
	const int center = y * a.W + x;
	float4 pl = a.planes[center];
	pl.w = depth_from_plane(a.ref, pl, x, y);
	a.planes[center] = normal_to_world(a.ref, pl);
}

// ------------------------------------------------------------------------------------------------------
// K13/K14 CheckerboardFilterStrong (APD.cu:3184-3294): median of the own depth and up to 20 STRONG
// neighbours at fixed opposite-colour offsets.
__global__ void __launch_bounds__(256) k_filter(const __grid_constant__ KArgs a, int red, int yy_limit) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	const int yy = blockIdx.y * blockDim.y + threadIdx.y;
	const int y = 2 * yy + ((x & 1) ^ red);
	const int W = a.W, H = a.H;
	if (x >= W || y >= H || yy >= yy_limit) return;
	const int center = y * W + x;
	if (a.weak[center] == DVP_WEAK) return;
	if (a.costs[center] < 0.001f) return;
	float f[21];
	int n = 0;
	f[n++] = a.planes[center].w;
	// (dx, dy, guard) in the reference's order
	auto take = [&](bool ok, int dx, int dy) {
		if (ok) {
			const int q = center + dy * W + dx;
			if (a.weak[q] == DVP_STRONG) f[n++] = a.planes[q].w;
		}
	};
	take(y > 0, 0, -1); take(y > 2, 0, -3); take(y > 4, 0, -5);
	take(y < H - 1, 0, 1); take(y < H - 3, 0, 3); take(y < H - 5, 0, 5);
	take(x > 0, -1, 0); take(x > 2, -3, 0); take(x > 4, -5, 0);
	take(x < W - 1, 1, 0); take(x < W - 3, 3, 0); take(x < W - 5, 5, 0);
	take(y > 0 && x < W - 2, 2, -1); take(y < H - 1 && x < W - 2, 2, 1);
	take(y > 0 && x > 1, -2, -1); take(y < H - 1 && x > 1, -2, 1);
	take(x > 0 && y > 2, -1, -2); take(x < W - 1 && y > 2, 1, -2);
	take(x > 0 && y < H - 2, -1, 2); take(x < W - 1 && y < H - 2, 1, 2);
	for (int i = 1; i < n; i++) {  // insertion sort (sort_small)
		const float tmp = f[i];
		int j;
		for (j = i; j >= 1 && tmp < f[j - 1]; j--) f[j] = f[j - 1];
		f[j] = tmp;
	}
	const int m = n / 2;
	a.planes[center].w = (n % 2 == 0) ? (f[m - 1] + f[m]) / 2 : f[m];
}

// ------------------------------------------------------------------------------------------------------
// Shared front end of K15/K16: plane at rest -> reference camera, weighted cost of the current depth,
// mean baseline of the selected views (APD.cu:3918-3960, 4067-4108).
struct ProfileCtx {
	float4 plane;        // normal in ref-camera frame, w = depth
	float depth;
	float weight_normal;
	float base_line;
	float cost_now;
	int valid;
};

__device__ __forceinline__ float profile_cost_sum(const KArgs& a, int x, int y, const float4 pl, uint32_t sel, const ViewWeights& vw,
                                                  const RefPatch& rp, const float2* wt, int T, bool k16_form) {
	// K15: temp = ncc (+ geom_factor*geom); p_cost += temp * w          (APD.cu:3976-3986)
	// K16: temp_cost += ncc * w; temp_cost += geom_factor * geom * w     (APD.cu:4121-4129)
	float acc = 0.0f;
	for (int v = 0; v < a.S; ++v) {
		if (!is_set(sel, v)) continue;
		const int wv = vw.get(v);
		if (wv == 0) continue;   // a selected view the last sweep never drew: its (finite) cost is multiplied by 0 and added — exactly +0
		if (k16_form) {
			acc += (ncc_cost<kWideRB>(a, a.views[v], a.tex_img[v + 1], x, y, pl, rp, wt, T) * wv);
			if (a.prm.geom_consistency) acc += (a.prm.geom_factor * geom_cost(a, a.views[v], a.tex_depth[v + 1], x, y, pl) * wv);
		} else {
			float temp_cost = 0.0f;
			temp_cost += ncc_cost<kWideRB>(a, a.views[v], a.tex_img[v + 1], x, y, pl, rp, wt, T);
			if (a.prm.geom_consistency) temp_cost += a.prm.geom_factor * geom_cost(a, a.views[v], a.tex_depth[v + 1], x, y, pl);
			acc += (temp_cost * wv);
		}
	}
	return acc;
}

// DepthToWeak's peak analysis (APD.cu:3999-4016) looks for local minima on indices 2..58 of the 61-entry profile, so it
// reads entries 1..59 and entry [min_peak] with min_peak in {0, 2..58}: the last entry is never read, and entry 0 only
// through `abs(min_peak - 30) > weak_peak_radius || p_costs[0] > 0.5f` when no peak exists — which the first operand
// already decides for every weak_peak_radius < 30 (the reference's schedules use 2, 4, 6).  Dead entries are not
// evaluated: 2 of 61 hypotheses, i.e. 2 * S_sel NCCs per pixel.
__device__ __forceinline__ int profile_first_live(const KArgs& a) { return a.prm.weak_peak_radius < 30 ? 1 : 0; }

template <bool WITH_COST>   // DepthToWeak never uses the cost of the current depth (APD.cu:3957 is dead there): K15 skips those NCCs
__device__ __forceinline__ void profile_front(const KArgs& a, int x, int y, int center, uint32_t sel, const ViewWeights& vw,
                                              const RefPatch& rp, const float2* wt, int T, ProfileCtx& pc) {
	pc.cost_now = 0.0f; pc.base_line = 0; pc.valid = 0; pc.weight_normal = 0.0f;
	// GetDistance2Origin(origin_depth) as the reference's LocalRefine compiles it (APD.cu:4083-4084): the
	// product depth*n.z is loop-invariant over the views and gets hoisted, so it is rounded on its own and
	// ADDED to fma(X0, n.x, X1*n.y) instead of being fused (SASS of the reference build, LocalRefine+0x0720).
	// DepthToWeak never uses the cost computed here, so only K16 observes the difference.
	float w_front;
	{
		const float rcp_k0 = rcp_approx(a.ref.K[0]), rcp_k4 = rcp_approx(a.ref.K[4]);
		const float x0 = __fmul_rn(__fmul_rn(pc.depth, __fadd_rn((float)x, -a.ref.K[2])), rcp_k0);
		const float x1 = __fmul_rn(__fmul_rn(pc.depth, __fadd_rn((float)y, -a.ref.K[5])), rcp_k4);
		const float dz = __fmul_rn(pc.depth, pc.plane.z);
		w_front = -__fadd_rn(dz, __fmaf_rn(x0, pc.plane.x, __fmul_rn(x1, pc.plane.y)));
	}
	for (int v = 0; v < a.S; ++v) {
		if (!is_set(sel, v)) continue;
		const int wv = vw.get(v);
		if (WITH_COST && wv != 0) {   // zero-weight view: contributes exactly +0
			float4 t = pc.plane;
			t.w = w_front;
			float temp_cost = ncc_cost<kWideRB>(a, a.views[v], a.tex_img[v + 1], x, y, t, rp, wt, T);
			if (a.prm.geom_consistency) temp_cost += a.prm.geom_factor * geom_cost(a, a.views[v], a.tex_depth[v + 1], x, y, t);
			pc.cost_now += (temp_cost * wv);
		}
		pc.weight_normal += wv;
		const dvp_camera& sc = a.cams[v + 1];
		float c_dist[3];
		c_dist[0] = a.ref.c[0] - sc.c[0];
		c_dist[1] = a.ref.c[1] - sc.c[1];
		c_dist[2] = a.ref.c[2] - sc.c[2];
		const double temp_val = c_dist[0] * c_dist[0] + c_dist[1] * c_dist[1] + c_dist[2] * c_dist[2];
		pc.base_line += sqrtf(temp_val);
		pc.valid++;
	}
}

// K15 DepthToWeak (APD.cu:3892-4051): 61-step disparity cost profile -> STRONG / WEAK / UNKNOWN.
__global__ void __launch_bounds__(256, 3) k_depth_to_weak(const __grid_constant__ KArgs a) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	constexpr int T = 256;   // compile-time stride: shared-memory offsets become immediates
	const int tid = threadIdx.y * blockDim.x + threadIdx.x;
	float2* wt = reinterpret_cast<float2*>(smem_raw) + tid;
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= a.W || y >= a.H) return;
	const int min_margin = 6;
	const int center = x + y * a.W;
	if (a.prm.use_radius && a.radius[center] == 0) a.radius[center] = a.prm.strong_radius;
	if (x < min_margin || y < min_margin || x >= a.W - min_margin || y >= a.H - min_margin) { a.weak[center] = DVP_UNKNOWN; return; }
	ProfileCtx pc;
	pc.plane = normal_to_refcam(a.ref, a.planes[center]);
	pc.depth = pc.plane.w;
	if (pc.depth == 0) { a.weak[center] = DVP_UNKNOWN; return; }
	const uint32_t sel = a.selected[center];
	ViewWeights vw; vw.load(a.view_weight + (size_t)center * DVP_MAX_IMAGES);
	RefPatch rp;
	rp.prepare(a, x, y, a.prm.use_radius ? a.radius[center] : a.prm.strong_radius, wt, T);
	profile_front<false>(a, x, y, center, sel, vw, rp, wt, T, pc);
	if (pc.valid == 0) { a.weak[center] = DVP_UNKNOWN; return; }
	pc.cost_now /= pc.weight_normal;
	pc.base_line /= pc.valid;

	const float disp = a.ref.K[0] * pc.base_line / pc.depth;
	const int radius = 30;
	const int p_costs_size = 2 * radius + 1;
	float p_costs[p_costs_size];  // 244 B of local memory per thread, touched ~240 times per ~8000 texture samples
	const int idx_lo = profile_first_live(a);
	p_costs[0] = 2.0f; p_costs[p_costs_size - 1] = 2.0f;
	for (int idx = idx_lo; idx < p_costs_size - 1; ++idx) {
		const int p_disp = idx - radius;
		const float p_depth = a.ref.K[0] * pc.base_line / (disp + p_disp);
		if (p_depth < a.prm.depth_min || p_depth > a.prm.depth_max) { p_costs[idx] = 2.0f; continue; }
		float4 t = pc.plane;
		t.w = get_distance2origin(a.ref, x, y, p_depth, t);
		float p_cost = profile_cost_sum(a, x, y, t, sel, vw, rp, wt, T, false);
		p_cost /= pc.weight_normal;
		p_costs[idx] = DVP_MIN(2.0f, p_cost);
	}
	// local minima of the profile (APD.cu:3999-4016)
	uint64_t is_peak = 0;
	int peak_count = 0, min_peak = 0;
	float min_cost = 2.0f;
	for (int i = 2; i < p_costs_size - 2; ++i) {
		if (p_costs[i - 1] > p_costs[i] && p_costs[i + 1] > p_costs[i]) {
			is_peak |= 1ull << i;
			peak_count++;
			if (p_costs[i] < min_cost) { min_peak = i; min_cost = p_costs[i]; }
		}
	}
	if (abs(min_peak - radius) > a.prm.weak_peak_radius || p_costs[min_peak] > 0.5f) { a.weak[center] = DVP_WEAK; return; }
	if (peak_count == 1) { a.weak[center] = (p_costs[min_peak] <= 0.15f) ? DVP_STRONG : DVP_WEAK; return; }
	float var = 0.0f;
	for (int i = 2; i < p_costs_size - 2; ++i) {
		if (((is_peak >> i) & 1) && i != min_peak) {
			const float dist = p_costs[i] - min_cost;
			var += dist * dist;
		}
	}
	var = sqrtf(var);
	var /= (peak_count - 1);
	a.weak[center] = (var > 0.2f) ? DVP_STRONG : DVP_WEAK;
}

// K16 LocalRefine (APD.cu:4053-4139): 11-step disparity scan, keep the best depth if it improves by > 0.1.
__device__ __forceinline__ void local_refine_pixel(const KArgs& a, int x, int y, int center, float2* wt) {
	constexpr int T = 256;
	ProfileCtx pc;
	pc.plane = normal_to_refcam(a.ref, a.planes[center]);
	pc.depth = pc.plane.w;
	if (pc.depth == 0) return;
	const uint32_t sel = a.selected[center];
	if (sel == 0) return;  // valid_neighbour == 0
	ViewWeights vw; vw.load(a.view_weight + (size_t)center * DVP_MAX_IMAGES);
	RefPatch rp;
	rp.prepare(a, x, y, a.prm.use_radius ? a.radius[center] : a.prm.strong_radius, wt, T);
	profile_front<true>(a, x, y, center, sel, vw, rp, wt, T, pc);
	if (pc.weight_normal == 0 || pc.valid == 0) return;
	pc.cost_now /= pc.weight_normal;
	pc.base_line /= pc.valid;
	const float disp = a.ref.K[0] * pc.base_line / pc.depth;
	const int radius = 5;
	float min_cost = 2.0f;
	float best_depth = pc.depth;
	for (int p_disp = -radius; p_disp <= radius; ++p_disp) {
		const float p_depth = a.ref.K[0] * pc.base_line / (disp + p_disp);
		if (p_depth < a.prm.depth_min || p_depth > a.prm.depth_max) continue;
		float4 t = pc.plane;
		t.w = get_distance2origin(a.ref, x, y, p_depth, t);
		float temp_cost = profile_cost_sum(a, x, y, t, sel, vw, rp, wt, T, true);
		temp_cost /= pc.weight_normal;
		if (temp_cost < min_cost) { min_cost = temp_cost; best_depth = p_depth; }
	}
	if (pc.cost_now - min_cost > 0.1) a.planes[center].w = best_depth;
}
__global__ void __launch_bounds__(256, 3) k_local_refine(const __grid_constant__ KArgs a) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int tid = threadIdx.y * blockDim.x + threadIdx.x;
	float2* wt = reinterpret_cast<float2*>(smem_raw) + tid;
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= a.W || y >= a.H) return;
	local_refine_pixel(a, x, y, x + y * a.W, wt);
}

// K15 + K16 in one kernel (what dvp_run launches).  LocalRefine's 11 disparity hypotheses (-5..5) are the middle
// of DepthToWeak's 61 (-30..30): same plane, same depth formula, same views, and DepthToWeak changes nothing that
// LocalRefine reads — so the per-view NCC and reprojection costs of those 11 are computed once and folded twice, in
// DepthToWeak's order (temp = ncc + f*geom; sum += temp*w, APD.cu:3976-3986) and in LocalRefine's
// (sum += ncc*w; sum += f*geom*w, APD.cu:4121-4129).  12*S_sel of the 73*S_sel NCCs of the two kernels remain as
// S_sel (the cost of the current depth, which only LocalRefine uses).  The 6-pixel margin DepthToWeak skips is
// refined by the stand-alone LocalRefine code.  Bit-exact against K15 followed by K16 (tests/test_gpu_parity.py).
#ifndef DVP_K15_MIN_BLOCKS
#define DVP_K15_MIN_BLOCKS 3
#endif
__global__ void __launch_bounds__(256, DVP_K15_MIN_BLOCKS) k_depth_to_weak_refine(const __grid_constant__ KArgs a) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	constexpr int T = 256;
	const int tid = threadIdx.y * blockDim.x + threadIdx.x;
	float2* wt = reinterpret_cast<float2*>(smem_raw) + tid;
	int x, y; full_grid_pixel(x, y);
	if (x >= a.W || y >= a.H) return;
	const int min_margin = 6;
	const int center = x + y * a.W;
	if (a.prm.use_radius && a.radius[center] == 0) a.radius[center] = a.prm.strong_radius;
	if (x < min_margin || y < min_margin || x >= a.W - min_margin || y >= a.H - min_margin) {
		a.weak[center] = DVP_UNKNOWN;
		local_refine_pixel(a, x, y, center, wt);
		return;
	}
	ProfileCtx pc;
	pc.plane = normal_to_refcam(a.ref, a.planes[center]);
	pc.depth = pc.plane.w;
	if (pc.depth == 0) { a.weak[center] = DVP_UNKNOWN; return; }   // LocalRefine returns here too
	const uint32_t sel = a.selected[center];
	ViewWeights vw; vw.load(a.view_weight + (size_t)center * DVP_MAX_IMAGES);
	RefPatch rp;
	rp.prepare(a, x, y, a.prm.use_radius ? a.radius[center] : a.prm.strong_radius, wt, T);
	profile_front<true>(a, x, y, center, sel, vw, rp, wt, T, pc);
	if (pc.valid == 0) { a.weak[center] = DVP_UNKNOWN; return; }   // no selected view: LocalRefine returns as well (sel == 0)
	const bool refine = (pc.weight_normal != 0);
	pc.cost_now /= pc.weight_normal;
	pc.base_line /= pc.valid;

	const float disp = a.ref.K[0] * pc.base_line / pc.depth;
	const int radius = 30;
	const int p_costs_size = 2 * radius + 1;
	float p_costs[p_costs_size];
	float refine_min = 2.0f, refine_depth = pc.depth;   // LocalRefine's running minimum over p_disp = -5..5, in its order
	const int idx_lo = profile_first_live(a);
	p_costs[0] = 2.0f; p_costs[p_costs_size - 1] = 2.0f;
	for (int idx = idx_lo; idx < p_costs_size - 1; ++idx) {
		const int p_disp = idx - radius;
		const float p_depth = a.ref.K[0] * pc.base_line / (disp + p_disp);
		if (p_depth < a.prm.depth_min || p_depth > a.prm.depth_max) { p_costs[idx] = 2.0f; continue; }
		float4 t = pc.plane;
		t.w = get_distance2origin(a.ref, x, y, p_depth, t);
		float p_cost;
		if (refine && p_disp >= -5 && p_disp <= 5) {
			float acc15 = 0.0f, acc16 = 0.0f;
			for (int v = 0; v < a.S; ++v) {
				if (!is_set(sel, v)) continue;
				if (vw.get(v) == 0) continue;   // adds exactly +0 to both folds
				const float wvf = (float)vw.get(v);
				const float ncc = ncc_cost<kWideRB>(a, a.views[v], a.tex_img[v + 1], x, y, t, rp, wt, T);
				float temp = __fadd_rn(0.0f, ncc);
				acc16 = __fmaf_rn(ncc, wvf, acc16);
				if (a.prm.geom_consistency) {
					const float g = geom_cost(a, a.views[v], a.tex_depth[v + 1], x, y, t);
					temp = __fmaf_rn(a.prm.geom_factor, g, temp);
					acc16 = __fmaf_rn(__fmul_rn(a.prm.geom_factor, g), wvf, acc16);
				}
				acc15 = __fmaf_rn(temp, wvf, acc15);
			}
			p_cost = acc15;
			float temp_cost = acc16;
			temp_cost /= pc.weight_normal;
			if (temp_cost < refine_min) { refine_min = temp_cost; refine_depth = p_depth; }
		} else {
			p_cost = profile_cost_sum(a, x, y, t, sel, vw, rp, wt, T, false);
		}
		p_cost /= pc.weight_normal;
		p_costs[idx] = DVP_MIN(2.0f, p_cost);
	}
	// LocalRefine's decision; DepthToWeak reads neither planes nor costs after this point
	if (refine && pc.cost_now - refine_min > 0.1) a.planes[center].w = refine_depth;
	// local minima of the profile (APD.cu:3999-4016)
	uint64_t is_peak = 0;
	int peak_count = 0, min_peak = 0;
	float min_cost = 2.0f;
	for (int i = 2; i < p_costs_size - 2; ++i) {
		if (p_costs[i - 1] > p_costs[i] && p_costs[i + 1] > p_costs[i]) {
			is_peak |= 1ull << i;
			peak_count++;
			if (p_costs[i] < min_cost) { min_peak = i; min_cost = p_costs[i]; }
		}
	}
	if (abs(min_peak - radius) > a.prm.weak_peak_radius || p_costs[min_peak] > 0.5f) { a.weak[center] = DVP_WEAK; return; }
	if (peak_count == 1) { a.weak[center] = (p_costs[min_peak] <= 0.15f) ? DVP_STRONG : DVP_WEAK; return; }
	float var = 0.0f;
	for (int i = 2; i < p_costs_size - 2; ++i) {
		if (((is_peak >> i) & 1) && i != min_peak) {
			const float dist = p_costs[i] - min_cost;
			var += dist * dist;
		}
	}
	var = sqrtf(var);
	var /= (peak_count - 1);
	a.weak[center] = (var > 0.2f) ? DVP_STRONG : DVP_WEAK;
}

// ------------------------------------------------------------------------------------------------------
// launchers
static inline dim3 full_grid(const KArgs& a, dim3 b) { return dim3((a.W + b.x - 1) / b.x, (a.H + b.y - 1) / b.y, 1); }
static inline int ref_half_rows(int H) { return (((H / 2) + 15) / 16) * 16; }  // rows-of-pairs covered by the reference's half grid

__global__ void k_fill_sd_table(float* out) {
	const int k = threadIdx.x;
	if (k < kHoistSamples) out[k] = RefPatch::spatial_dist(-5 + 2 * (k / kHoistAxis), -5 + 2 * (k % kHoistAxis));
}
cudaError_t launch_fill_sd_table(cudaStream_t st) {
	static std::mutex mu;
	static bool ready[256] = {false};
	int dev = 0; cudaGetDevice(&dev);
	std::lock_guard<std::mutex> lock(mu);
	if (dev >= 0 && dev < 256 && ready[dev]) return cudaSuccess;
	float* tmp = nullptr;
	cudaError_t e = cudaMalloc((void**)&tmp, kHoistSamples * sizeof(float));
	if (e != cudaSuccess) return e;
	k_fill_sd_table<<<1, 64, 0, st>>>(tmp);
	e = cudaMemcpyToSymbolAsync(c_sd_r5, tmp, kHoistSamples * sizeof(float), 0, cudaMemcpyDeviceToDevice, st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	cudaFree(tmp);
	if (e == cudaSuccess && dev >= 0 && dev < 256) ready[dev] = true;
	return e;
}
cudaError_t launch_setup_views(const dvp_camera* cams, ViewConst* views, int S, cudaStream_t st) {
	k_setup_views<<<1, 32, 0, st>>>(cams, views, S);
	return cudaGetLastError();
}
cudaError_t launch_init_rng(const KArgs& a, unsigned long long seed, cudaStream_t st) {
	const int n = ((a.W + 31) / 32) * a.H;
	k_init_rng<<<(n + 127) / 128, 128, 0, st>>>(a, seed);
	return cudaGetLastError();
}
cudaError_t launch_random_init(const KArgs& a, cudaStream_t st) {
	dim3 b(32, 8);
	k_random_init<<<full_grid(a, b), b, patch_smem_bytes(256), st>>>(a);
	return cudaGetLastError();
}
size_t sweep_scratch_bytes(int W, int H, int S) {
	const size_t slots = (size_t)W * ((H + 1) / 2);
	const size_t strong = slots * ((size_t)9 * S * sizeof(float) + sizeof(uint32_t) + sizeof(uint4) + sizeof(float4));
	const size_t weak = (slots + 32) * (size_t)8 * S * sizeof(float);   // k_weak_score: 8 S rows of the colour's WEAK count rounded up to 32 (<= slots + 31)
	return (strong > weak ? strong : weak) + 256;
}
cudaError_t launch_strong_sweep(const KArgs& a, int iter, int red, void* scratch, cudaStream_t st) {
	const int yy_limit = ref_half_rows(a.H);
#ifdef DVP_SWEEP_FUSED   // the single-kernel form (kept for A/B measurements and as the body of the forced instrumentation variant)
	(void)scratch;
	dim3 b(kSweepBlockX, kSweepThreads / kSweepBlockX);
	dim3 g((a.W + b.x - 1) / b.x, (yy_limit + b.y - 1) / b.y, 1);
	k_strong_sweep<false><<<g, b, sweep_smem_bytes(kSweepThreads, a.S), st>>>(a, iter, red, yy_limit, D4Force{});
	return cudaGetLastError();
#else
	if (!scratch) return cudaErrorInvalidValue;
	SweepScratch sc;
	sc.slots = a.W * ((a.H + 1) / 2);
	char* base = reinterpret_cast<char*>(scratch);
	sc.d4_plane = reinterpret_cast<float4*>(base);          base += (size_t)sc.slots * sizeof(float4);
	sc.pos = reinterpret_cast<uint4*>(base);                base += (size_t)sc.slots * sizeof(uint4);
	sc.cost = reinterpret_cast<float*>(base);               base += (size_t)sc.slots * 9 * a.S * sizeof(float);
	sc.flag = reinterpret_cast<uint32_t*>(base);
#if DVP_QUAD_SWEEP == 2
	// diamond rows r = 0 .. H / 2 + 1 (ya = 2 r - 2 + red from -2 to H: the first and last image rows are reached through a quad's lower / upper pixel), 8 diamonds of a
	// row per warp, k = 0 .. (W + 1) / 4 + 1 so that xa = 4 k - 2 (r & 1) reaches W - 1
	const int gx = ((a.W + 1) / 4 + 2 + 7) / 8, rows = a.H / 2 + 2;
#else
	const int gx = (a.W + 31) / 32, rows = yy_limit;
#endif
	{
		dim3 b(32, 8), g(gx, (rows + 7) / 8, 1);
		k_sweep_score<<<g, b, patch_smem_bytes(256), st>>>(a, iter, red, yy_limit, sc);
	}
	{
		dim3 b(32, 4), g(gx, (rows + 3) / 4, 1);
		k_sweep_update<<<g, b, patch_smem_bytes(128), st>>>(a, iter, red, yy_limit, sc);
	}
	return cudaGetLastError();
#endif
}
cudaError_t launch_strong_sweep_forced(const KArgs& a, int iter, int red, const D4Force& force, cudaStream_t st) {
	dim3 b(kSweepBlockX, kSweepThreads / kSweepBlockX);
	const int yy_limit = ref_half_rows(a.H);
	dim3 g((a.W + b.x - 1) / b.x, (yy_limit + b.y - 1) / b.y, 1);
	if (force.pixel_list) g = dim3((force.list_count + kSweepThreads - 1) / kSweepThreads, 1, 1);
	if (g.x == 0) return cudaSuccess;
	k_strong_sweep<true><<<g, b, sweep_smem_bytes(kSweepThreads, a.S), st>>>(a, iter, red, yy_limit, force);
	return cudaGetLastError();
}
cudaError_t launch_explain_init(const KArgs& a, int red, const RaceExpected& e, uint8_t* explained, cudaStream_t st) {
	k_explain_init<<<(a.N + 255) / 256, 256, 0, st>>>(a, red, ref_half_rows(a.H), e, explained);
	return cudaGetLastError();
}
cudaError_t launch_explain_compare(const KArgs& a, int red, const D4Force& f, const RaceExpected& e, uint8_t* explained, cudaStream_t st) {
	const int n = f.pixel_list ? f.list_count : a.N;
	if (n == 0) return cudaSuccess;
	k_explain_compare<<<(n + 255) / 256, 256, 0, st>>>(a, red, ref_half_rows(a.H), f, e, explained);
	return cudaGetLastError();
}
cudaError_t launch_explain_collect(const KArgs& a, int red, const uint8_t* explained, int* list, int* count, int cap, cudaStream_t st) {
	k_explain_collect<<<(a.N + 255) / 256, 256, 0, st>>>(a, red, ref_half_rows(a.H), explained, list, count, cap);
	return cudaGetLastError();
}
cudaError_t launch_depth_normal(const KArgs& a, cudaStream_t st) {
	dim3 b(32, 8);
	k_depth_normal<<<full_grid(a, b), b, 0, st>>>(a);
	return cudaGetLastError();
}
cudaError_t launch_filter(const KArgs& a, int red, cudaStream_t st) {
	dim3 b(32, 8);
	const int yy_limit = ref_half_rows(a.H);
	dim3 g((a.W + 31) / 32, (yy_limit + b.y - 1) / b.y, 1);
	k_filter<<<g, b, 0, st>>>(a, red, yy_limit);
	return cudaGetLastError();
}
cudaError_t launch_depth_to_weak(const KArgs& a, cudaStream_t st) {
	dim3 b(32, 8);
	k_depth_to_weak<<<full_grid(a, b), b, patch_smem_bytes(256), st>>>(a);
	return cudaGetLastError();
}
cudaError_t launch_depth_to_weak_refine(const KArgs& a, cudaStream_t st) {
	dim3 b(32, 8);
	k_depth_to_weak_refine<<<full_grid(a, b), b, patch_smem_bytes(256), st>>>(a);
	return cudaGetLastError();
}
cudaError_t launch_local_refine(const KArgs& a, cudaStream_t st) {
	dim3 b(32, 8);
	k_local_refine<<<full_grid(a, b), b, patch_smem_bytes(256), st>>>(a);
	return cudaGetLastError();
}
cudaError_t configure_strong_kernels(int S) {
	cudaError_t e;
	if ((e = cudaFuncSetAttribute(k_random_init, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)patch_smem_bytes(256)))) return e;
	if ((e = cudaFuncSetAttribute(k_depth_to_weak, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)patch_smem_bytes(256)))) return e;
	if ((e = cudaFuncSetAttribute(k_local_refine, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)patch_smem_bytes(256)))) return e;
	if ((e = cudaFuncSetAttribute(k_depth_to_weak_refine, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)patch_smem_bytes(256)))) return e;
	if ((e = cudaFuncSetAttribute(k_strong_sweep<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem_bytes(kSweepThreads, S)))) return e;
	if ((e = cudaFuncSetAttribute(k_sweep_score, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)patch_smem_bytes(256)))) return e;
	if ((e = cudaFuncSetAttribute(k_sweep_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)patch_smem_bytes(128)))) return e;
	if ((e = cudaFuncSetAttribute(k_strong_sweep<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem_bytes(kSweepThreads, S)))) return e;
#ifdef DVP_SWEEP_CARVEOUT
	if ((e = cudaFuncSetAttribute(k_strong_sweep<false>, cudaFuncAttributePreferredSharedMemoryCarveout, DVP_SWEEP_CARVEOUT))) return e;
#endif
	return cudaSuccess;
}

}  // namespace dvp
