// dvp_kernels_weak.cu — the adaptive-patch-deformation (WEAK pixel) path: K2 part (a) candidate offsets,
// K4 GenNeighbours, K9 RANSACToGetFitPlane, K10/K11 weak propagation
// (reference APD.cu:3746-3794, 3330-3711, 4195-4404, 2739-3125, 835-1021, 1897-2008).
#include "dvp_weak.cuh"
#include "dvp_launch.h"
#include <cfloat>
#include <mutex>

namespace dvp {

cudaError_t launch_edge_inform_prep(const KArgs& a, cudaStream_t st);  // dvp_kernels_prep.cu

__constant__ int c_dirw[8][2] = {{0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {1, 1}, {-1, 1}, {1, -1}};
// 30-degree sector of every offset of the 11x11 window, filled once on the device with the same double
// atan2 the reference evaluates 120 times per pixel per view (calculateAngle / getRegion, APD.cu:797-821)
__device__ int8_t g_sector[11][11];
// the same table, sector-major: the offsets of each sector in the reference's scan order (i outer, j inner)
__device__ int g_sector_cnt[12];
__device__ char2 g_sector_off[12][24];

__global__ void k_fill_sector_table() {
	const int i = (int)threadIdx.x - 5, j = (int)threadIdx.y - 5;
	double angle = atan2((double)j, (double)i) * (180.0 / 3.14159265358979323846);
	if (angle < 0) angle += 360.0;
	int region = -1;
	for (int r = 0; r < 12; ++r)
		if (angle >= 30.0 * r && angle < 30.0 * (r + 1)) region = r;
	g_sector[threadIdx.x][threadIdx.y] = (int8_t)region;
	__syncthreads();
	if (threadIdx.x == 0 && threadIdx.y == 0) {
		for (int r = 0; r < 12; ++r) g_sector_cnt[r] = 0;
		for (int ii = -5; ii <= 5; ++ii)
			for (int jj = -5; jj <= 5; ++jj) {
				if (ii == 0 && jj == 0) continue;
				const int r = g_sector[ii + 5][jj + 5];
				if (r >= 0 && g_sector_cnt[r] < 24) g_sector_off[r][g_sector_cnt[r]++] = make_char2((signed char)ii, (signed char)jj);
			}
	}
}

// ------------------------------------------------------------------------------------------------------
// K2 part (a) (APD.cu:3746-3794): per source view, the best (largest colour weight) pixel that sees the view
// in each of 12 sectors of the 11x11 window; the 8 best sectors' offsets -> candidate[pixel][view][8].
// Sectors without any visible pixel hold uninitialised stack data in the reference; here they
// rank last and yield the offset (0,0), which the deformable NCC replaces by its default ring offset.
// `only` (dvp_run): the records are read for one purpose — the deformable NCC looks up the offsets of a WEAK pixel's ANCHORS
// (ncc_new) — and the anchors are known after K4.  Nothing the records depend on (reference image, selected views) changes
// between K2 and K5, so dvp_run evaluates them after K4 for the pixels some anchor list names (k_mark_anchors) and skips the
// tiles without any: the same records where they are read, 1 % of the work at the bench workload.  only == nullptr (stage
// stepping, the parity tests): every pixel, as the reference.
__global__ void __launch_bounds__(256) k_candidate(const __grid_constant__ KArgs a, const uint8_t* __restrict__ only) {
	// 32x8 pixel tile + 5 px halo of the reference image and of the selected-view masks, staged once in shared
	// memory and reused by the 120 window reads of every pixel for every view (weak_radius is 5 in every schedule
	// of the reference; other radii fall back to global reads)
	constexpr int R = 5, TW = 32 + 2 * R, TH = 8 + 2 * R;
	__shared__ float s_img[TH][TW];
	__shared__ uint32_t s_sel[TH][TW];
	const int W = a.W, H = a.H;
	bool wanted = true;
	if (only) {
		const int qx = blockIdx.x * 32 + threadIdx.x, qy = blockIdx.y * 8 + threadIdx.y;
		wanted = qx < W && qy < H && only[qx + qy * W] != 0;
		if (!__syncthreads_or(wanted)) return;   // no anchor in this tile
	}
	const int x0 = blockIdx.x * 32 - R, y0 = blockIdx.y * 8 - R;
	for (int i = threadIdx.y * 32 + threadIdx.x; i < TW * TH; i += 256) {
		const int ty = i / TW, tx = i - ty * TW;
		const int gx = x0 + tx, gy = y0 + ty;
		const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
		s_img[ty][tx] = in ? a.ref_img[(size_t)gy * W + gx] : 0.f;
		s_sel[ty][tx] = in ? a.selected[gx + gy * W] : 0u;   // out-of-image pixels are skipped by the reference: no view set
	}
	__syncthreads();
	const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
	if (x >= W || y >= H || !wanted) return;
	const int center = x + y * W;
	float rcp_s, rcp_c; RefPatch::sigma_rcps(a.prm, rcp_s, rcp_c);
	const int radius = a.prm.weak_radius;
	const bool tiled = (radius == R);
	const int lx = threadIdx.x + R, ly = threadIdx.y + R;
	const float ref_center_pix = s_img[ly][lx];
	for (int v = 0; v < a.S; ++v) {
		// sector by sector (static register indexing): the winner of a sector is its largest weight, first in scan
		// order on ties — what the reference's stable descending bubble sort leaves in regions[r][0]
		float best_w[12]; int best_ij[12];   // (i, j) packed; -1 = empty sector
#pragma unroll
		for (int r = 0; r < 12; ++r) {
			float bw = 0.f; int bij = -1;
			const int cnt = tiled ? g_sector_cnt[r] : 0;
			for (int k = 0; k < cnt; ++k) {
				const char2 o = g_sector_off[r][k];
				const int i = o.x, j = o.y;
				const int rx = x + i, ry = y + j;
				if (!(rx >= 0 && rx < W && ry >= 0 && ry < H)) continue;
				if (is_set(s_sel[ly + j][lx + i], v) != 1) continue;
				const float w = weight_colour(s_img[ly + j][lx + i], ref_center_pix, rcp_c);
				if (bij < 0 || w > bw) { bw = w; bij = ((i + 5) << 4) | (j + 5); }
			}
			best_w[r] = bw; best_ij[r] = bij;
		}
		if (!tiled) {   // weak_radius != 5: generic window, global reads (never the case in the reference's schedules)
			for (int i = -radius; i <= radius; i++)
				for (int j = -radius; j <= radius; j++) {
					if (i == 0 && j == 0) continue;
					const int rx = x + i, ry = y + j;
					if (!(rx >= 0 && rx < W && ry >= 0 && ry < H)) continue;
					if (is_set(a.selected[rx + ry * W], v) != 1) continue;
					double angle = atan2((double)j, (double)i) * (180.0 / 3.14159265358979323846);
					if (angle < 0) angle += 360.0;
					const int r = (int)(angle / 30.0);
					if (r < 0 || r > 11) continue;
					const float w = weight_colour(RefPatch::ref_pixel(a, rx, ry), ref_center_pix, rcp_c);
#pragma unroll
					for (int rr = 0; rr < 12; ++rr)
						if (rr == r && (best_ij[rr] < 0 || w > best_w[rr])) { best_w[rr] = w; best_ij[rr] = ((i + 64) << 8) | (j + 64) | (1 << 20); }
				}
		}
		// stable descending sort of the 12 sector winners (bubbleSort, APD.cu:823-833); empty sectors last
#pragma unroll
		for (int p = 0; p < 11; ++p)
#pragma unroll
			for (int q = 0; q < 11 - p; ++q) {
				const bool swap = best_ij[q + 1] >= 0 && (best_ij[q] < 0 || best_w[q] < best_w[q + 1]);
				if (swap) {
					const float tw = best_w[q]; best_w[q] = best_w[q + 1]; best_w[q + 1] = tw;
					const int ti = best_ij[q]; best_ij[q] = best_ij[q + 1]; best_ij[q + 1] = ti;
				}
			}
		short2* cand = a.candidate + ((size_t)center * DVP_NUM_IMAGES + v) * DVP_LAB_BOUNDARY_NUM;
		short2 outc[DVP_LAB_BOUNDARY_NUM];
#pragma unroll
		for (int k = 0; k < DVP_LAB_BOUNDARY_NUM; ++k) {
			const int b = best_ij[k];
			if (b < 0) outc[k] = make_short2(0, 0);
			else if (b & (1 << 20)) outc[k] = make_short2((short)(((b >> 8) & 0xff) - 64), (short)((b & 0xff) - 64));
			else outc[k] = make_short2((short)((b >> 4) - 5), (short)((b & 15) - 5));
		}
		// one 32-byte record per (pixel, view): two 16-byte stores
		uint4* c4 = reinterpret_cast<uint4*>(cand);
		const uint32_t* o32 = reinterpret_cast<const uint32_t*>(outc);
		c4[0] = make_uint4(o32[0], o32[1], o32[2], o32[3]);
		c4[1] = make_uint4(o32[4], o32[5], o32[6], o32[7]);
	}
}

// ------------------------------------------------------------------------------------------------------
// K4 GenNeighbours (APD.cu:3330-3711), one thread per WEAK pixel.
constexpr int kMaxPts = 160;

__device__ __forceinline__ void normalize2(float2* v) {
	const float n2 = v->x * v->x + v->y * v->y;
	const float inv = rsqrtf(n2);
	v->x *= inv; v->y *= inv;
}

#ifndef DVP_K4_MIN_BLOCKS
#define DVP_K4_MIN_BLOCKS (768 / DVP_K4_THREADS)   // 24 warps per SM (80 registers)
#endif
__global__ void __launch_bounds__(kK4Threads, DVP_K4_MIN_BLOCKS) k_gen_neighbours(const __grid_constant__ KArgs a, const int* weak_list) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= a.weak_count) return;
	const int center = weak_list[t];
	const int W = a.W, H = a.H;
	if (a.weak[center] != DVP_WEAK) return;   // demoted by K2 since the list was built
	const int px = center % W, py = center / W;
	const int min_margin = 6;
	const float depth_diff = a.prm.depth_max - a.prm.depth_min;
	const int nm = a.neighbours_map[center];
	short2* neighbours = a.neighbours + (size_t)nm * DVP_NEIGHBOUR_NUM;
	Rng rng; rng.load(a.rng, a.N, center);
	for (int i = 0; i < DVP_NEIGHBOUR_NUM; ++i) neighbours[i] = make_short2(-1, -1);
	neighbours[0] = make_short2((short)px, (short)py);
	// Slots 0..31 belong to the 8 x 4 search directions (valid or not: `dir_valid`, one bit each); slots 32..extend_index are
	// appended by the label-boundary extension and are all valid.  Nothing above extend_index is ever read, so only the
	// direction slots are initialised (the reference initialises, and later scans, all 160: 1.4 KB of stack writes per pixel).
	constexpr int kDirSlots = 32;
	// Anchor slots as packed 32-bit words (x in the low half, y in the high half; (-1, -1) = 0xffffffff), 16-byte aligned: the
	// two duplicate scans below compare four slots per 128-bit local load without a branch per slot.  They were the top two
	// stall sites of this kernel (ncu, profiles/r02_final_kernels_ncu.txt: 19 % of the samples): a dependent 16-bit load,
	// compare and branch per slot.
	__align__(16) uint32_t strong_points[kMaxPts];
	auto pack_pt = [](short2 p) -> uint32_t { return (uint32_t)(uint16_t)p.x | ((uint32_t)(uint16_t)p.y << 16); };
	auto unpack_pt = [](uint32_t v) -> short2 { return make_short2((short)(v & 0xffffu), (short)(v >> 16)); };
	uint32_t dir_valid = 0;
	for (int i = 0; i < kDirSlots; ++i) strong_points[i] = 0xffffffffu;
	int origin_direction_index = -1;
	int strong_point_size = 0;

	const int rotate_time = a.prm.rotate_time;
	const float angle = 45.0f / rotate_time;
	const float cos_angle = cos(angle * 3.14159265358979323846 / 180.f);
	const float sin_angle = sin(angle * 3.14159265358979323846 / 180.f);
	const float threshhold = cos((angle / 2.0f) * 3.14159265358979323846 / 180.0f);
	const int shift_range = DVP_MAX((int)(tan((angle / 2.0f) * 3.14159265358979323846 / 180.0f) * 20), 1);
	const FastMod shift_mod((unsigned int)shift_range);   // x % shift_range for the 4 draws of every try: a multiply instead of the generic 32-bit division
	const float ransac_threshold = a.prm.ransac_threshold;

	bool edge_limit = false;
	if (a.prm.use_limit) {
		edge_limit = true;
		if (a.prm.use_edge) {
			const float complex_val = a.complex_[nm];
			const float rand_prob = rng.uniform() - FLT_EPSILON;
			if (rand_prob < complex_val) edge_limit = false;
		}
	}

	// (Flattening the three nested loops below into one loop whose iteration is one rung of the lane's own current direction —
	// so that a lane starts its next direction as soon as its own search ends — was measured on B200, bit-identical, and is
	// SLOWER: 14.2 -> 17.3 ms at 3111x2073, 23.5 -> 25.9 on derived states.  Lanes of a tile walk the same direction at the
	// same radius, so their probe gathers land next to each other; once the lanes drift apart the gathers scatter, and that
	// costs more than the idle lanes did.  profiles/r02_k4_restructure.txt)
	for (int ox = -1; ox <= 1; ++ox) {
		for (int oy = -1; oy <= 1; ++oy) {
			if (ox == 0 && oy == 0) continue;
			float2 origin_direction = make_float2(ox, oy);
			normalize2(&origin_direction);
			origin_direction_index++;
			for (int rotate_iter = 0; rotate_iter < rotate_time; ++rotate_iter) {
				const int dir_index = origin_direction_index * 4 + rotate_iter;
				for (int radius = 2; radius <= 4096; radius = DVP_MIN(radius * 2, radius + 25)) {
					const float2 test_pt = make_float2(px + origin_direction.x * radius, py + origin_direction.y * radius);
					if (test_pt.x < 0 || test_pt.y < 0 || test_pt.x >= W || test_pt.y >= H) break;
					// The reference makes up to four tries per radius, each: 4 RNG draws -> a probe pixel -> its state and,
					// if it is not STRONG, its nearest STRONG pixel (two dependent gathers) -> tests.  Tries only interact
					// through the RNG and the early exit, so all four probes are drawn first, their eight gathers are
					// issued together, and the tests then run in the reference's order; if try j succeeds, the RNG is
					// rolled back to its state after try j (the draws the reference never made).  The cheap angle test
					// goes before the duplicate scan (both merely skip the try, neither has side effects).
					short2 probe[4]; int probe_idx[4]; Rng::Snap after[4];
#pragma unroll
					for (int radius_iter = 0; radius_iter < 4; ++radius_iter) {
						// (cond ? 1 : -1) * curand() % range : unsigned arithmetic, two draws per shift, left operand first
						const unsigned int c0 = rng.next(); const unsigned int c1 = rng.next();
						const int rand_x_shift = (int)shift_mod.mod((unsigned int)(c0 % 2 == 0 ? 1 : -1) * c1);
						const unsigned int c2 = rng.next(); const unsigned int c3 = rng.next();
						const int rand_y_shift = (int)shift_mod.mod((unsigned int)(c2 % 2 == 0 ? 1 : -1) * c3);
						float2 direction = make_float2(origin_direction.x * 20 + rand_x_shift, origin_direction.y * 20 + rand_y_shift);
						normalize2(&direction);
						const short2 np = make_short2(px + direction.x * radius, py + direction.y * radius);
						const bool outside = np.x < min_margin || np.y < min_margin || np.x >= W - min_margin || np.y >= H - min_margin;
						probe[radius_iter] = np;
						probe_idx[radius_iter] = outside ? -1 : np.x + np.y * W;
						after[radius_iter] = rng.snap();
					}
					uint8_t probe_state[4]; short2 probe_ns[4];
#pragma unroll
					for (int radius_iter = 0; radius_iter < 4; ++radius_iter) {
						probe_state[radius_iter] = DVP_STRONG; probe_ns[radius_iter] = make_short2(-1, -1);
						if (probe_idx[radius_iter] >= 0) {
							probe_state[radius_iter] = a.weak[probe_idx[radius_iter]];
							probe_ns[radius_iter] = a.nearest_strong[probe_idx[radius_iter]];
						}
					}
#pragma unroll
					for (int radius_iter = 0; radius_iter < 4; ++radius_iter) {
						if ((dir_valid >> dir_index) & 1) continue;   // an earlier try of this batch succeeded: the reference has left the loop
						if (probe_idx[radius_iter] < 0) continue;
						short2 np = probe[radius_iter];
						if (probe_state[radius_iter] != DVP_STRONG) {
							np = probe_ns[radius_iter];
							if (np.x == -1 || np.y == -1) continue;
						}
						float2 test_direction = make_float2(np.x - px, np.y - py);
						normalize2(&test_direction);
						const float cosv = test_direction.x * origin_direction.x + test_direction.y * origin_direction.y;
						if (!(cosv > threshhold)) continue;
						// the reference scans slots 0 .. dir_index - 1; slots from dir_index on still hold (-1, -1), which no probe can equal
						// (both coordinates of np are >= 0 here), so all 32 direction slots are compared
						bool has_same_pt = false;
						{
							const uint32_t np32 = pack_pt(np);
							const uint4* sp4 = reinterpret_cast<const uint4*>(strong_points);
#pragma unroll
							for (int c = 0; c < kDirSlots / 4; ++c) {
								const uint4 v = sp4[c];
								has_same_pt |= (v.x == np32) | (v.y == np32) | (v.z == np32) | (v.w == np32);
							}
						}
						if (has_same_pt) continue;
						if (!edge_limit || !bresenham_crosses_edge(a, px, py, np.x, np.y)) {
							strong_points[dir_index] = pack_pt(np);
							dir_valid |= 1u << dir_index;
							strong_point_size++;
							rng.restore(after[radius_iter]);
						}
					}
					if ((dir_valid >> dir_index) & 1) break;
				}
				float2 rotated;
				rotated.x = origin_direction.x * cos_angle - origin_direction.y * sin_angle;
				rotated.y = origin_direction.x * sin_angle + origin_direction.y * cos_angle;
				normalize2(&rotated);
				origin_direction = rotated;
			}
		}
	}

	int extend_index = 31;
	const int my_label = a.label[center];
	if (a.prm.use_label && my_label > 0) {
		// {1, 0.5} etc. are narrowed to int in the reference's `const int dir[16][2]` (APD.cu:3462): 0.5 -> 0
		const int dir[16][2] = {{0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {1, 1}, {-1, 1}, {1, -1},
		                        {1, 0}, {0, 1}, {0, 1}, {-1, 0}, {-1, 0}, {0, -1}, {0, -1}, {1, 0}};
		const short2* lab_bound = a.label_boundary + (size_t)nm * DVP_LAB_BOUNDARY_NUM;
		float bound_dist[16] = {0};
		int dir_step[16] = {0};
		for (int i = 0; i < 8; ++i) {
			const short2 bp = lab_bound[i];
			float dist = 0.0f;
			if (bp.x != -1 && bp.y != -1) {
				const int ddx = px - bp.x, ddy = py - bp.y;
				dist = (float)sqrt((double)(ddx * ddx) + (double)(ddy * ddy));
				if (i >= 4) dist = (float)((double)dist / sqrt(2.0));
			}
			bound_dist[i] = dist;
			if (i % 2 == 1) {
				const float opposite_dist = bound_dist[i - 1];
				const int step = DVP_MIN(1, DVP_MAX(4 * rotate_time - 1, (int)(4 * rotate_time * dist / (dist + opposite_dist))));  // B13: always 1
				dir_step[i - 1] = 4 * rotate_time - step;
				dir_step[i] = step;
			}
		}
		dir_step[8] = (dir_step[3] + dir_step[5]) / 2;   bound_dist[8] = (bound_dist[3] + bound_dist[5]) / 2;
		dir_step[9] = (dir_step[1] + dir_step[5]) / 2;   bound_dist[9] = (bound_dist[1] + bound_dist[5]) / 2;
		dir_step[10] = (dir_step[1] + dir_step[6]) / 2;  bound_dist[10] = (bound_dist[1] + bound_dist[6]) / 2;
		dir_step[11] = (dir_step[2] + dir_step[6]) / 2;  bound_dist[11] = (bound_dist[2] + bound_dist[6]) / 2;
		dir_step[12] = (dir_step[2] + dir_step[4]) / 2;  bound_dist[12] = (bound_dist[2] + bound_dist[4]) / 2;
		dir_step[13] = (dir_step[4] + dir_step[0]) / 2;  bound_dist[13] = (bound_dist[4] + bound_dist[0]) / 2;
		dir_step[14] = (dir_step[7] + dir_step[0]) / 2;  bound_dist[14] = (bound_dist[7] + bound_dist[0]) / 2;
		dir_step[15] = (dir_step[7] + dir_step[3]) / 2;  bound_dist[15] = (bound_dist[7] + bound_dist[3]) / 2;
		for (int i = 0; i < 16; ++i) {
			const float dist = bound_dist[i];
			const int gap_num = dir_step[i] + 1;
			const int step_len = DVP_MAX(1, (int)floor(1.0 * dist / gap_num));
			for (int step = 1; step <= dir_step[i]; ++step) {
				short2 np = make_short2(px + step * step_len * dir[i][0], py + step * step_len * dir[i][1]);
				if (np.x < min_margin || np.y < min_margin || np.x >= W - min_margin || np.y >= H - min_margin) continue;
				int npc = np.x + np.y * W;
				if (a.weak[npc] != DVP_STRONG) {
					np = a.nearest_strong[npc];
					if (np.x == -1 || np.y == -1) continue;
					npc = np.x + np.y * W;
				}
				bool has_same_pt = false;   // any of slots 0 .. extend_index (slots above it are not initialised: masked)
				{
					const uint32_t np32 = pack_pt(np);
					const uint4* sp4 = reinterpret_cast<const uint4*>(strong_points);
					const int n_slots = extend_index + 1;
					for (int c = 0; 4 * c < n_slots; ++c) {
						const uint4 v = sp4[c];
						const int left = n_slots - 4 * c;   // >= 1
						has_same_pt |= (v.x == np32) | ((left > 1) & (v.y == np32)) | ((left > 2) & (v.z == np32)) | ((left > 3) & (v.w == np32));
					}
				}
				if (has_same_pt) continue;
				if (extend_index + 1 >= kMaxPts) continue;  // cannot happen with rotate_time <= 4 (31 + 128 slots)
				extend_index++;
				strong_points[extend_index] = pack_pt(np);
				strong_point_size++;
			}
		}
	}

	uint8_t* weak_reliable = a.weak_reliable + center;
	if (strong_point_size <= 3) { *weak_reliable = 0; rng.store(a.rng, a.N, center); return; }

	float4 best_plane = make_float4(0, 0, 0, 0);
	bool has_valid_plane = false;
	short2 valid_pts[kMaxPts];
	float3 valid_3d[kMaxPts];
	float2 valid_factor[kMaxPts];   // ((x - cx) / fx, (y - cy) / fy) of every anchor: the reference recomputes them for each anchor in each of its 200 RANSAC tries
	int valid_count = 0;
	float X[3];
	get_3d_point(a.ref, px, py, a.planes[center].w, X);
	const float center_z = X[2];
	for (int i = 0; i < DVP_NEIGHBOUR_NUM - 1; ++i) valid_pts[i] = make_short2(-1, -1);   // read back below even when fewer anchors survive
	for (int i = 0; i <= extend_index; ++i) {
		if (i >= kDirSlots || ((dir_valid >> i) & 1)) {
			const short2 sp = unpack_pt(strong_points[i]);
			const int spc = sp.x + sp.y * W;
			valid_pts[valid_count] = sp;
			get_3d_point(a.ref, sp.x, sp.y, a.planes[spc].w, X);
			valid_3d[valid_count] = make_float3(X[0], X[1], X[2]);
			valid_factor[valid_count] = make_float2((sp.x - a.ref.K[2]) / a.ref.K[0], (sp.y - a.ref.K[5]) / a.ref.K[4]);
			valid_count++;
		}
	}
	auto anchor_normal = [&](int idx) -> float3 {   // strong_points_valid_normals[idx], recomputed on demand
		const short2 sp = valid_pts[idx];
		const float4 n4 = normal_to_refcam(a.ref, a.planes[sp.x + sp.y * W]);
		return make_float3(n4.x, n4.y, n4.z);
	};
	{
		int iteration = 300, max_iter = 200;
		float min_cost = FLT_MAX;
		int max_count = 3;
		// edge_test[160][160] of the reference, 2 bits per pair (0 unknown, 1 crosses an edge, 2 clear): 6.4 KB instead of 25.6 KB
		// (indexed with the actual number of anchors, so a typical pixel touches ~2 KB of it)
		uint32_t edge_test[kMaxPts * kMaxPts / 16];
		const int et_words = (valid_count * valid_count + 15) / 16;
		for (int i = 0; i < et_words; ++i) edge_test[i] = 0;
		auto et_get = [&](int r, int c) -> int { const int b = r * valid_count + c; return (edge_test[b >> 4] >> ((b & 15) * 2)) & 3; };
		auto et_set = [&](int r, int c, int val) { const int b = r * valid_count + c; edge_test[b >> 4] = (edge_test[b >> 4] & ~(3u << ((b & 15) * 2))) | ((uint32_t)val << ((b & 15) * 2)); };
		auto crossing = [&](int i0, int i1) -> int {
			int e = et_get(i0, i1);
			if (e == 0) {
				e = bresenham_crosses_edge(a, valid_pts[i0].x, valid_pts[i0].y, valid_pts[i1].x, valid_pts[i1].y) ? 1 : 2;
				et_set(i0, i1, e); et_set(i1, i0, e);
			}
			return e;
		};
		// The reference's loop (APD.cu:3569-3678) makes 200 tries — `iteration` starts at 300 and falls by one per
		// evaluated triple, so `max_iter` alone ends it — and a try is: three draws, the triple's tests (distinct,
		// pixel inside the triangle, no edge between the corners, a usable normal, a proper plane), and for the ~25 % of
		// the triples that pass, an inlier count over all anchors and the running-best update.  As one loop a warp pays
		// the inlier count in nearly every iteration with a quarter of its lanes (some lane's triple passes almost every
		// time).  It is split where the data flow splits: the tests consume the RNG and fill the edge-test cache in try
		// order but never read the running best; the running best never feeds back into the tests.  Pass A (in two steps, below)
		// runs the 200 tries' draws and tests and records the triples that pass (with the one bit the tests hand on:
		// is_strong_plane); pass B walks the record in try order.  Same arithmetic on the same operands in the same order per pixel.
#ifndef DVP_K4_FUSED_RANSAC
		uint32_t passed[200];   // a | b << 8 | c << 16 | is_strong_plane << 24 (indices < kMaxPts = 160)
		int n_passed = 0;
		const FastMod count_mod((unsigned int)valid_count);   // x % valid_count for the 600 draws: exact (tests/test_host_logic.py)
		const bool label_rule = a.prm.use_label && my_label > 0;
		auto plane_of = [&](int ia, int ib, int ic, float4& cross_vec) -> bool {
			const float3 A = valid_3d[ia], B = valid_3d[ib], C = valid_3d[ic];
			const float3 A_C = make_float3(A.x - C.x, A.y - C.y, A.z - C.z);
			const float3 B_C = make_float3(B.x - C.x, B.y - C.y, B.z - C.z);
			cross_vec.x = A_C.y * B_C.z - B_C.y * A_C.z;
			cross_vec.y = -(A_C.x * B_C.z - B_C.x * A_C.z);
			cross_vec.z = A_C.x * B_C.y - B_C.x * A_C.y;
			if ((cross_vec.x == 0 && cross_vec.y == 0 && cross_vec.z == 0) || isnan(cross_vec.x) || isnan(cross_vec.y) || isnan(cross_vec.z)) return false;
			normalize3(&cross_vec);
			cross_vec.w = -(cross_vec.x * A.x + cross_vec.y * A.y + cross_vec.z * A.z);
			return true;
		};
		// pass A1: the draws and the two tests without side effects (a quarter of the triples survive)
		for (int tries = 0; tries < max_iter; ++tries) {
			const int a_index = (int)count_mod.mod(rng.next());
			const int b_index = (int)count_mod.mod(rng.next());
			const int c_index = (int)count_mod.mod(rng.next());
			if (a_index == b_index || b_index == c_index || a_index == c_index) continue;
			if (!point_in_triangle(valid_pts[a_index], valid_pts[b_index], valid_pts[c_index], px, py)) continue;
			passed[n_passed++] = (uint32_t)a_index | ((uint32_t)b_index << 8) | ((uint32_t)c_index << 16);
		}
		// pass A2: the remaining tests in try order (the edge-test cache is filled by whichever try asks first), compacting in place
		const int n_inside = n_passed;
		n_passed = 0;
		for (int q = 0; q < n_inside; ++q) {
			const int a_index = (int)(passed[q] & 255), b_index = (int)((passed[q] >> 8) & 255), c_index = (int)((passed[q] >> 16) & 255);
			if (edge_limit) {
				const int e_ab = crossing(a_index, b_index);
				const int e_bc = crossing(b_index, c_index);
				const int e_ca = crossing(c_index, a_index);
				if (e_ab == 1 || e_bc == 1 || e_ca == 1) continue;
			}
			const float3 AN = anchor_normal(a_index);   // B8: the reference compares normal A with itself
			const float nn = AN.x * AN.x + AN.y * AN.y + AN.z * AN.z;
			if (nn < 0.9f || nn < 0.9f || nn < 0.9f) continue;
			float4 cross_vec;
			if (!plane_of(a_index, b_index, c_index, cross_vec)) continue;
			bool is_strong_plane = true;
			const float dn = fabs(AN.x * cross_vec.x + AN.y * cross_vec.y + AN.z * cross_vec.z);
			if (label_rule && dn < 0.9f && dn < 0.9f && dn < 0.9f) is_strong_plane = false;
			passed[n_passed++] = (uint32_t)a_index | ((uint32_t)b_index << 8) | ((uint32_t)c_index << 16) | ((uint32_t)is_strong_plane << 24);
		}
		(void)iteration; (void)min_cost;
		bool has_strong_plane = false;
		const float factor_x = __fmul_rn(__fadd_rn((float)px, -a.ref.K[2]), rcp_approx(a.ref.K[0]));
		const float factor_y = __fmul_rn(__fadd_rn((float)py, -a.ref.K[5]), rcp_approx(a.ref.K[4]));
		const float rcp_depth_diff = rcp_approx(depth_diff);
		float best_center_distance = FLT_MAX;
		for (int q = 0; q < n_passed; ++q) {
			const uint32_t rec = passed[q];
			const bool is_strong_plane = (rec >> 24) & 1;
			if (has_strong_plane && !is_strong_plane) continue;
			float4 cross_vec;
			plane_of((int)(rec & 255), (int)((rec >> 8) & 255), (int)((rec >> 16) & 255), cross_vec);
			// |fit_depth - z| with fit_depth = -w / (x fx + y fy + z) exactly as the reference build evaluates it (SASS of
			// GenNeighbours and of the single-loop form here): the y product is rounded, the x product fused onto it, z added,
			// and the quotient fused with the subtraction — fma(-w, rcp(den), -z).  Written out because under
			// --use_fast_math the compiler picks the order of the two products per context.
			auto depth_gap = [&](float fx, float fy, float z) -> float {
				const float den = __fadd_rn(cross_vec.z, __fmaf_rn(cross_vec.x, fx, __fmul_rn(cross_vec.y, fy)));
				return fabsf(__fmaf_rn(-cross_vec.w, rcp_approx(den), -z));
			};
			int temp_count = 0;
			for (int si = 0; si < valid_count; ++si) {
				const float distance = depth_gap(valid_factor[si].x, valid_factor[si].y, valid_3d[si].z);
				if (__fmul_rn(distance, rcp_depth_diff) < ransac_threshold) temp_count++;   // (the reference also sums the distances; the sum is never read)
			}
			if (temp_count < 6) continue;
			if (temp_count > max_count || (!has_strong_plane && is_strong_plane)) {
				if (!has_strong_plane && is_strong_plane) has_strong_plane = true;
				best_center_distance = depth_gap(factor_x, factor_y, center_z);
				best_plane = cross_vec;
				max_count = temp_count;
				has_valid_plane = true;
			} else if (temp_count == max_count) {
				const float center_distance = depth_gap(factor_x, factor_y, center_z);
				if (center_distance < best_center_distance) { best_plane = cross_vec; max_count = temp_count; best_center_distance = center_distance; }
			}
		}
#else
		bool has_strong_plane = false;
		while (iteration > 0 && max_iter > 0) {
			max_iter--;
			const int a_index = rng.next() % valid_count;
			const int b_index = rng.next() % valid_count;
			const int c_index = rng.next() % valid_count;
			if (a_index == b_index || b_index == c_index || a_index == c_index) continue;
			if (!point_in_triangle(valid_pts[a_index], valid_pts[b_index], valid_pts[c_index], px, py)) continue;
			if (edge_limit) {
				const int e_ab = crossing(a_index, b_index);
				const int e_bc = crossing(b_index, c_index);
				const int e_ca = crossing(c_index, a_index);
				if (e_ab == 1 || e_bc == 1 || e_ca == 1) continue;
			}
			const float3 AN = anchor_normal(a_index);   // B8: the reference compares normal A with itself
			const float nn = AN.x * AN.x + AN.y * AN.y + AN.z * AN.z;
			if (nn < 0.9f || nn < 0.9f || nn < 0.9f) continue;
			const float3 A = valid_3d[a_index], B = valid_3d[b_index], C = valid_3d[c_index];
			const float3 A_C = make_float3(A.x - C.x, A.y - C.y, A.z - C.z);
			const float3 B_C = make_float3(B.x - C.x, B.y - C.y, B.z - C.z);
			float4 cross_vec;
			cross_vec.x = A_C.y * B_C.z - B_C.y * A_C.z;
			cross_vec.y = -(A_C.x * B_C.z - B_C.x * A_C.z);
			cross_vec.z = A_C.x * B_C.y - B_C.x * A_C.y;
			if ((cross_vec.x == 0 && cross_vec.y == 0 && cross_vec.z == 0) || isnan(cross_vec.x) || isnan(cross_vec.y) || isnan(cross_vec.z)) continue;
			iteration--;
			normalize3(&cross_vec);
			cross_vec.w = -(cross_vec.x * A.x + cross_vec.y * A.y + cross_vec.z * A.z);
			bool is_strong_plane = true;
			const float dn = fabs(AN.x * cross_vec.x + AN.y * cross_vec.y + AN.z * cross_vec.z);
			if (a.prm.use_label && my_label > 0 && dn < 0.9f && dn < 0.9f && dn < 0.9f) is_strong_plane = false;
			if (has_strong_plane && !is_strong_plane) continue;
			int temp_count = 0;
			for (int si = 0; si < valid_count; ++si) {
				const float factor_x = valid_factor[si].x, factor_y = valid_factor[si].y;
				const float fit_depth = -cross_vec.w / (cross_vec.x * factor_x + cross_vec.y * factor_y + cross_vec.z);
				const float distance = fabs(fit_depth - valid_3d[si].z);
				if (distance / depth_diff < ransac_threshold) temp_count++;   // (the reference also sums the distances; the sum is never read)
			}
			if (temp_count < 6) continue;
			if (temp_count > max_count || (!has_strong_plane && is_strong_plane)) {
				if (!has_strong_plane && is_strong_plane) has_strong_plane = true;
				const float factor_x = (px - a.ref.K[2]) / a.ref.K[0];
				const float factor_y = (py - a.ref.K[5]) / a.ref.K[4];
				const float fit_depth = -cross_vec.w / (cross_vec.x * factor_x + cross_vec.y * factor_y + cross_vec.z);
				const float center_distance = fabs(fit_depth - center_z);
				best_plane = cross_vec;
				max_count = temp_count;
				min_cost = center_distance;
				has_valid_plane = true;
			} else if (temp_count == max_count) {
				const float factor_x = (px - a.ref.K[2]) / a.ref.K[0];
				const float factor_y = (py - a.ref.K[5]) / a.ref.K[4];
				const float fit_depth = -cross_vec.w / (cross_vec.x * factor_x + cross_vec.y * factor_y + cross_vec.z);
				const float center_distance = fabs(fit_depth - center_z);
				if (center_distance < min_cost) { best_plane = cross_vec; max_count = temp_count; min_cost = center_distance; }
			}
		}
#endif
	}
	rng.store(a.rng, a.N, center);
	if (!has_valid_plane) { *weak_reliable = 0; return; }

	// Distance of every anchor to the winning plane, outliers last, the DVP_NEIGHBOUR_NUM - 1 best kept (APD.cu:3680-3708).
	// The reference insertion-sorts all anchors (sort_small_weighted, APD.cu:125-138: ~n^2 / 4 dependent moves through local
	// memory, 9 % of this kernel's stall samples) and then reads the first 11.  A stable ascending sort's first 11 are kept
	// here in a sorted register list instead — an anchor goes in front of the first kept one that is strictly heavier, i.e.
	// behind its equals, as the insertion sort leaves it.  NaN weights make the reference's loop order-dependent (`<` is false
	// both ways, a NaN stops every later element): those pixels take the reference's loop.
	constexpr int K = DVP_NEIGHBOUR_NUM - 1;
	float top_w[K]; uint32_t top_pt[K];
#pragma unroll
	for (int j = 0; j < K; ++j) { top_w[j] = __int_as_float(0x7f800000); top_pt[j] = 0xffffffffu; }   // +inf: behind every anchor (outliers weigh FLT_MAX)
	float weight[kMaxPts];
	bool any_nan = false;
	{
		const float rcp_dd = rcp_approx(depth_diff);
		for (int i = 0; i < valid_count; ++i) {
			// |fit_depth - z| in the rounding order of the reference build (see depth_gap above)
			const float den = __fadd_rn(best_plane.z, __fmaf_rn(best_plane.x, valid_factor[i].x, __fmul_rn(best_plane.y, valid_factor[i].y)));
			const float distance = fabsf(__fmaf_rn(-best_plane.w, rcp_approx(den), -valid_3d[i].z));
			const bool outlier = __fmul_rn(distance, rcp_dd) >= ransac_threshold;
			const float w = outlier ? FLT_MAX : distance;
			if (outlier) valid_pts[i] = make_short2(-1, -1);
			weight[i] = w;
			any_nan |= (w != w);
			const uint32_t pt = outlier ? 0xffffffffu : pack_pt(valid_pts[i]);
			if (w < top_w[K - 1]) {
#pragma unroll
				for (int j = K - 1; j >= 1; --j) {
					const bool shift = w < top_w[j - 1], here = w < top_w[j];
					top_pt[j] = shift ? top_pt[j - 1] : (here ? pt : top_pt[j]);
					top_w[j] = shift ? top_w[j - 1] : (here ? w : top_w[j]);
				}
				if (w < top_w[0]) { top_w[0] = w; top_pt[0] = pt; }
			}
		}
	}
	if (!any_nan) {
#pragma unroll
		for (int i = 1; i < DVP_NEIGHBOUR_NUM; ++i) neighbours[i] = unpack_pt(top_pt[i - 1]);
	} else {
		for (int i = 1; i < valid_count; i++) {   // sort_small_weighted, APD.cu:125-138
			const short2 tmp = valid_pts[i];
			const float tmp_w = weight[i];
			int j;
			for (j = i; j >= 1 && tmp_w < weight[j - 1]; j--) { valid_pts[j] = valid_pts[j - 1]; weight[j] = weight[j - 1]; }
			valid_pts[j] = tmp; weight[j] = tmp_w;
		}
		for (int i = 1; i < DVP_NEIGHBOUR_NUM; ++i) neighbours[i] = valid_pts[i - 1];
	}
	*weak_reliable = 1;
}

// ------------------------------------------------------------------------------------------------------
// K9 RANSACToGetFitPlane (APD.cu:4195-4404).
// Non-WEAK pixels: fit plane = current plane (APD.cu:4204-4207), a streaming copy over the image.
__global__ void __launch_bounds__(256) k_fit_copy_nonweak(const __grid_constant__ KArgs a) {
	const int center = blockIdx.x * blockDim.x + threadIdx.x;
	if (center < a.N && a.weak[center] != DVP_WEAK) a.fit_planes[center] = a.planes[center];
}
// WEAK pixels: one thread each over the tile-ordered WEAK list (a block = one 8x8-pixel tile, so the anchors' planes and
// the edge walks of a block overlap in the L1).  Pixels demoted since the list was built fall through to the copy rule.
__global__ void __launch_bounds__(64) k_ransac_fit(const __grid_constant__ KArgs a, const int* weak_list) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= a.weak_count) return;
	const int center = weak_list[t];
	const int W = a.W;
	if (a.weak[center] != DVP_WEAK) { a.fit_planes[center] = a.planes[center]; return; }
	const int px = center % W, py = center / W;
	Rng rng; rng.load(a.rng, a.N, center);
	const int nm = a.neighbours_map[center];
	bool edge_limit = false;
	if (a.prm.use_limit) {
		edge_limit = true;
		if (a.prm.use_edge) {
			const float complex_val = a.complex_[nm];
			const float rand_prob = rng.uniform() - FLT_EPSILON;
			if (rand_prob < complex_val) edge_limit = false;
		}
	}
	constexpr int M = DVP_NEIGHBOUR_NUM - 1;
	short2 strong_points[M];
	float3 pts3d[M], normals[M];
	int strong_count = 0;
	float X[3];
	const short2* nb = a.neighbours + (size_t)nm * DVP_NEIGHBOUR_NUM;
	for (int i = 1; i < DVP_NEIGHBOUR_NUM; ++i) {
		const short2 tp = nb[i];
		if (tp.x == -1 || tp.y == -1) continue;
		strong_points[strong_count] = tp;
		const int tc = tp.x + tp.y * W;
		const float4 pl = a.planes[tc];
		const float depth = depth_from_plane(a.ref, pl, tp.x, tp.y);
		get_3d_point(a.ref, tp.x, tp.y, depth, X);
		pts3d[strong_count] = make_float3(X[0], X[1], X[2]);
		normals[strong_count] = make_float3(pl.x, pl.y, pl.z);
		strong_count++;
	}
	if (strong_count < 3) { a.fit_planes[center] = a.planes[center]; rng.store(a.rng, a.N, center); return; }
	int iteration = 50;
	float min_cost = FLT_MAX;
	float4 best_plane = make_float4(0, 0, 0, 0);
	bool has_best_plane = false;
	uint8_t edge_test[M][M];
	for (int i = 0; i < M; ++i) for (int j = 0; j < M; ++j) edge_test[i][j] = 0;
	auto crossing = [&](int i0, int i1) -> int {
		if (edge_test[i0][i1] == 0)
			edge_test[i0][i1] = edge_test[i1][i0] = bresenham_crosses_edge<4>(a, strong_points[i0].x, strong_points[i0].y, strong_points[i1].x, strong_points[i1].y) ? 1 : 2;
		return edge_test[i0][i1];
	};
	while (iteration--) {
		const int a_index = rng.next() % strong_count;
		const int b_index = rng.next() % strong_count;
		const int c_index = rng.next() % strong_count;
		if (a_index == b_index || b_index == c_index || a_index == c_index) continue;
		const float3 AN = normals[a_index], BN = normals[b_index], CN = normals[c_index];
		if (AN.x * BN.x + AN.y * BN.y + AN.z * BN.z < 0.9f || AN.x * CN.x + AN.y * CN.y + AN.z * CN.z < 0.9f ||
		    BN.x * CN.x + BN.y * CN.y + BN.z * CN.z < 0.9f) continue;
		if (!point_in_triangle(strong_points[a_index], strong_points[b_index], strong_points[c_index], px, py)) continue;
		if (edge_limit) {
			const int e_ab = crossing(a_index, b_index);
			const int e_bc = crossing(b_index, c_index);
			const int e_ca = crossing(c_index, a_index);
			if (e_ab == 1 || e_bc == 1 || e_ca == 1) continue;
		}
		const float3 A = pts3d[a_index], B = pts3d[b_index], C = pts3d[c_index];
		const float3 A_C = make_float3(A.x - C.x, A.y - C.y, A.z - C.z);
		const float3 B_C = make_float3(B.x - C.x, B.y - C.y, B.z - C.z);
		float4 cross_vec;
		cross_vec.x = A_C.y * B_C.z - B_C.y * A_C.z;
		cross_vec.y = -(A_C.x * B_C.z - B_C.x * A_C.z);
		cross_vec.z = A_C.x * B_C.y - B_C.x * A_C.y;
		if ((cross_vec.x == 0 && cross_vec.y == 0 && cross_vec.z == 0) || isnan(cross_vec.x) || isnan(cross_vec.y) || isnan(cross_vec.z)) continue;
		normalize3(&cross_vec);
		cross_vec.w = -(cross_vec.x * A.x + cross_vec.y * A.y + cross_vec.z * A.z);
		float temp_cost = 0.0f;
		for (int si = 0; si < strong_count; ++si) {
			if (si == a_index || si == b_index || si == c_index) continue;
			const float3 tp = pts3d[si];
			const short2 tpix = strong_points[si];
			const float factor_x = (tpix.x - a.ref.K[2]) / a.ref.K[0];
			const float factor_y = (tpix.y - a.ref.K[5]) / a.ref.K[4];
			const float fit_depth = -cross_vec.w / (cross_vec.x * factor_x + cross_vec.y * factor_y + cross_vec.z);
			temp_cost += fabs(fit_depth - tp.z);
		}
		if (temp_cost < min_cost) { min_cost = temp_cost; best_plane = cross_vec; has_best_plane = true; }
	}
	rng.store(a.rng, a.N, center);
	if (has_best_plane) {
		const float depth = depth_from_plane(a.ref, a.planes[center], px, py);
		const float4 vd = get_view_direction(a.ref, px, py, depth);
		const float dot_product = best_plane.x * vd.x + best_plane.y * vd.y + best_plane.z * vd.z;
		if (dot_product > 0) { best_plane.x = -best_plane.x; best_plane.y = -best_plane.y; best_plane.z = -best_plane.z; best_plane.w = -best_plane.w; }
		a.fit_planes[center] = best_plane;
		if (a.prm.use_radius) {
			// The reference derives the adaptive radius from strong_points[use_a/b/c_index] with all three indices
			// still -1: A == B == C whatever that memory holds, the triangle is degenerate, its area
			// and hence the radius are 0; no later clamp can raise it, so the stored value is always 0.
			a.radius[center] = 0;
		}
	} else {
		a.fit_planes[center] = make_float4(0, 0, 0, 0);
		if (a.prm.use_radius) a.radius[center] = a.prm.strong_radius;
	}
}

// ------------------------------------------------------------------------------------------------------
// K10/K11 CheckerboardPropagationWeak + PlaneHypothesisRefinementWeak (APD.cu:2739-3089, 1897-2008).
// `reject_at`: see weighted_cost (dvp_strong.cuh) — the terms are non-negative (NCC in [0, 2], reprojection error in [0, 3],
// geom_factor >= 0), so the views left over are skipped once the scaled partial sum has reached the cost to beat.
__device__ __forceinline__ float weak_weighted_cost(const KArgs& a, int px, int py, const float4 pl, const ViewWeights& vw, float weight_norm, float reject_at) {
	float temp_cost = 0.0f;
	const bool can_reject = !(a.prm.geom_factor < 0.0f);
	for (int j = 0; j < a.S; ++j) {
		const int wv = vw.get(j);
		if (wv > 0) {
			const float c = ncc_new(a, px, py, j, pl);
			if (a.prm.geom_consistency) temp_cost += wv * (c + a.prm.geom_factor * geom_cost(a, a.views[j], a.tex_depth[j + 1], px, py, pl));
			else temp_cost += wv * c;
#ifdef DVP_EARLY_REJECT   // measured on B200: 7 % slower (see weighted_cost); off
			if (can_reject && !(temp_cost / weight_norm < reject_at)) break;
#endif
		}
	}
	temp_cost /= weight_norm;
	return temp_cost;
}

// K10/K11, part 1 — the propagation hypotheses, scored coherently.  A WEAK pixel tries the planes of its first eight anchors
// (APD.cu:2771-2797): 8 x S deformable NCCs, more than half of the 15 x S the kernel evaluates per pixel, and they need nothing
// but the pixel, the anchor's plane and the view.  Thread-per-pixel, every lane of a warp walks its own anchors: the reference
// pixel reads and the source fetches of the 32 lanes are scattered (ncu: TEX and LSU data pipes half busy each behind
// dependent loads, 207 Gfetch/s against the 1 142 the texture unit sustains when the four lanes of a quad fetch neighbouring
// texels, profiles/r01_tex_coherence_ubench.txt).  Here EIGHT ADJACENT LANES take the eight hypotheses of ONE pixel: they walk
// the same anchors at the same sample offsets, so their anchor / selected-view / offset / reference-pixel loads are one
// broadcast transaction instead of eight, and their source fetches — one reference position warped by eight neighbouring
// planes — land a few texels apart, which is what the texture unit's quad path wants.  Every lane still folds its own NCC in
// the reference's sample order: no cross-lane arithmetic, same bits.  Costs go to the sweep scratch area (free between K8 and
// the next K7), rows (i * S + v) of `stride` floats indexed by the pixel's position in the colour list.
#ifndef DVP_WEAK_SCORE_MIN_BLOCKS
#define DVP_WEAK_SCORE_MIN_BLOCKS 4   // 64 registers, 32 warps per SM: 3 % faster than 3 blocks at 80
#endif
__global__ void __launch_bounds__(256, DVP_WEAK_SCORE_MIN_BLOCKS) k_weak_score(const __grid_constant__ KArgs a, const int* colour_list, int count, float* __restrict__ scored, int stride) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	const int slot = t >> 3, i = t & 7;
	if (slot >= count) return;
	const int center = colour_list[slot];
	if (a.weak[center] != DVP_WEAK) return;   // demoted to UNKNOWN by K2/K5 since the list was built
	const int W = a.W, S = a.S;
	const int px = center % W, py = center / W;
	const short2 np = a.neighbours[(size_t)a.neighbours_map[center] * DVP_NEIGHBOUR_NUM + i + 1];
	const bool active = !(np.x == -1 || np.y == -1 || a.weak[np.x + np.y * W] != DVP_STRONG);   // else: the consumer applies the same test and never reads this row
	if (!active) return;
	const float4 pl = a.planes[np.x + np.y * W];   // a STRONG pixel's plane: K10 / K11 write WEAK pixels only
	for (int v = 0; v < S; ++v) scored[(size_t)(i * S + v) * stride + slot] = ncc_new(a, px, py, v, pl);
	// Tried and rejected: splitting the group's common work over its lanes (lane g reads offset word g and reference pixel g,
	// forms w, w r, r w r once; 32 shuffles per anchor hand them round; helpers without a hypothesis stay to do their share) —
	// bit-exact, 7 % slower (K10 + K11 40.3 against 37.6 ms): the broadcast loads it saves were already one transaction, and the
	// shuffles plus the predicated source side cost more issue slots than eight redundant weight evaluations.
}

// One thread per WEAK pixel of ONE checkerboard colour: `colour_list` (built at upload by a device prefix sum)
// holds the pixels the reference's half grid reaches for this colour (APD.cu:3093-3106), so warps are dense.
#ifndef DVP_WEAK_MIN_BLOCKS
#define DVP_WEAK_MIN_BLOCKS (768 / DVP_WEAK_THREADS)   // 24 warps per SM at 80 registers: measured 7 % faster than 16 warps at 127
#endif
__global__ void __launch_bounds__(kWeakThreads, DVP_WEAK_MIN_BLOCKS) k_weak_sweep(const __grid_constant__ KArgs a, const int* colour_list, int count, int iter, const float* __restrict__ scored, int stride) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= count) return;
	const int center = colour_list[t];
	const int W = a.W, S = a.S;
	const int px = center % W, py = center / W;
	if (a.weak[center] != DVP_WEAK) return;   // demoted to UNKNOWN by K2/K5 since the list was built
	const int nm = a.neighbours_map[center];
	const short2* nb = a.neighbours + (size_t)nm * DVP_NEIGHBOUR_NUM;

	// the reference's cost_array[8][32] (local memory; a shared-memory column per thread was measured slower: it
	// takes its 9*S*4 bytes per thread from the L1 this kernel's scattered fetches live on)
	float cost_array[8][DVP_MAX_IMAGES];
#define COST(i, j) cost_array[i][j]
	for (int i = 0; i < 8; ++i) for (int j = 0; j < S; ++j) COST(i, j) = 0.0f;
	COST(0, 0) = 2.0f;   // B2
	bool flag[8];
	int positions[8];
	for (int i = 0; i < 8; ++i) {
		flag[i] = false; positions[i] = 0;
		const short2 np = nb[i + 1];
		if (np.x == -1 || np.y == -1 || a.weak[np.x + np.y * W] != DVP_STRONG) continue;
		positions[i] = np.x + np.y * W;
		flag[i] = true;
		if (scored) {   // k_weak_score has been here
			for (int v = 0; v < S; ++v) COST(i, v) = scored[(size_t)(i * S + v) * stride + t];
		} else {
			const float4 pl = a.planes[positions[i]];
			for (int v = 0; v < S; ++v) COST(i, v) = ncc_new(a, px, py, v, pl);
		}
	}
	Rng rng; rng.load(a.rng, a.N, center);
	ViewWeights vw; vw.clear();
	float probs[DVP_MAX_IMAGES];
#define PROB(i) probs[i]
	{
		const float cost_threshold = 0.8 * expf((iter) * (iter) / (-90.0f));
		for (int i = 0; i < S; i++) {
			float prior = 0.0f;
			for (int k = 0; k < 8; ++k) {
				const short2 np = nb[k + 1];
				if (np.x == -1 || np.y == -1) continue;
				prior += is_set(a.selected[np.x + np.y * W], i) ? 0.9f : 0.1f;
			}
			float count = 0; int count_false = 0; float tmpw = 0;
			for (int j = 0; j < 8; j++) {
				const float c = COST(j, i);
				if (c < cost_threshold) { tmpw += expf(c * c / (-0.18f)); count++; }
				if (c > 1.2f) count_false++;
			}
			float prob = 0.0f;
			if (count > 2 && count_false < 3) prob = tmpw / count;
			else if (count_false < 3) prob = expf(cost_threshold * cost_threshold / (-0.32f));
			PROB(i) = prob * prior;
		}
		float prob_sum = 0.0f;
		for (int i = 0; i < S; ++i) prob_sum += PROB(i);
		const float inv_prob_sum = 1.0f / prob_sum;
		float cum_prob = 0.0f;
		for (int i = 0; i < S; ++i) { const float prob = PROB(i) * inv_prob_sum; cum_prob += prob; PROB(i) = cum_prob; }
		for (int sample = 0; sample < 15; ++sample) {
			const float rand_prob = rng.uniform() - FLT_EPSILON;
			for (int image_id = 0; image_id < S; ++image_id)
				if (PROB(image_id) > rand_prob) { vw.inc(image_id); break; }
		}
	}
	vw.store(a.view_weight + (size_t)center * DVP_MAX_IMAGES);
	uint32_t temp_selected = 0;
	float weight_norm = 0;
	for (int i = 0; i < S; ++i) { const int wv = vw.get(i); if (wv > 0) { temp_selected |= 1u << i; weight_norm += wv; } }

	int min_cost_idx = 0; float min_final = 0.f;
	for (int i = 0; i < 8; ++i) {
		float fc = 0.0f;
		for (int j = 0; j < S; ++j) {
			const int wv = vw.get(j);
			if (wv > 0) {
				if (a.prm.geom_consistency) {
					if (flag[i]) fc += wv * (COST(i, j) + a.prm.geom_factor * geom_cost(a, a.views[j], a.tex_depth[j + 1], px, py, a.planes[positions[i]]));
					else fc += wv * (COST(i, j) + a.prm.geom_factor * 3.0f);
				} else fc += wv * COST(i, j);
			}
		}
		fc /= weight_norm;
		if (i == 0 || fc <= min_final) { min_final = fc; min_cost_idx = i; }
	}
	float4 plane_now = a.planes[center];
	float cost_now = 0.0f;
	for (int i = 0; i < S; ++i) {
		const int wv = vw.get(i);
		if (wv == 0) continue;   // the reference multiplies by 0 here
		const float c = ncc_new(a, px, py, i, plane_now);
		if (a.prm.geom_consistency) cost_now += wv * (c + a.prm.geom_factor * geom_cost(a, a.views[i], a.tex_depth[i + 1], px, py, plane_now));
		else cost_now += wv * c;
	}
	cost_now /= weight_norm;
	const float cost_stored = cost_now;
	float depth_now = depth_from_plane(a.ref, plane_now, px, py);
	if (flag[min_cost_idx]) {
		const float4 winner = a.planes[positions[min_cost_idx]];   // a STRONG anchor's plane: not written by this kernel
		const float depth_before = depth_from_plane(a.ref, winner, px, py);
		if (depth_before >= a.prm.depth_min && depth_before <= a.prm.depth_max && min_final < cost_now) {
			depth_now = depth_before; plane_now = winner; cost_now = min_final;
			a.selected[center] = temp_selected;
		}
	}
	const uint32_t sel_now = a.selected[center];
	// ---- PlaneHypothesisRefinementWeak ----
	{
		const float depth_min = a.prm.depth_min, depth_max = a.prm.depth_max;
		bool skip_all = false;
		const float4 fit = a.fit_planes[center];
		if (fit.x == 0 && fit.y == 0 && fit.z == 0) skip_all = true;   // `return` before the random refinement (APD.cu:1923-1925)
		if (!skip_all) {
			const float depth_before = depth_from_plane(a.ref, fit, px, py);
			if (depth_before >= depth_min && depth_before <= depth_max) {   // else rejected whatever it costs
				const float temp_cost = weak_weighted_cost(a, px, py, fit, vw, weight_norm, cost_now);
				if (temp_cost < cost_now) { depth_now = depth_before; plane_now = fit; cost_now = temp_cost; }
			}
			const float depth_rand = rng.uniform() * (depth_max - depth_min) + depth_min;
			const float4 plane_rand = random_normal(a, px, py, rng, depth_now, sel_now);
			float depth_perturbed = depth_now;
			const float depth_min_perturbed = (1 - 0.02f) * depth_perturbed;
			const float depth_max_perturbed = (1 + 0.02f) * depth_perturbed;
			depth_perturbed = rng.uniform() * (depth_max_perturbed - depth_min_perturbed) + depth_min_perturbed;
			const float perturbation = 0.02f * 3.14159265358979323846;
			const float4 plane_pert = perturbed_normal(a, px, py, plane_now, rng, perturbation);
			(void)perturbed_normal(a, px, py, plane_now, rng, perturbation);
			const float depth0 = depth_now; const float4 plane0 = plane_now;
			for (int i = 0; i < 6; ++i) {
				if (i == 4) continue;   // identical to hypothesis 3, can never be accepted after it
				float d; float4 tpl;
				switch (i) {
				case 0: d = depth_rand; tpl = plane0; break;
				case 1: d = depth0; tpl = plane_rand; break;
				case 2: d = depth_rand; tpl = plane_rand; break;
				case 3: d = depth0; tpl = plane_pert; break;
				default: d = depth_perturbed; tpl = plane0; break;
				}
				tpl.w = get_distance2origin(a.ref, px, py, d, tpl);
				const float db = depth_from_plane(a.ref, tpl, px, py);
				if (!(db >= depth_min && db <= depth_max)) continue;
				const float tc = weak_weighted_cost(a, px, py, tpl, vw, weight_norm, cost_now);
				if (tc < cost_now) { depth_now = db; plane_now = tpl; cost_now = tc; }
			}
		}
	}
	rng.store(a.rng, a.N, center);
	float4 plane_final = plane_now;
	if (a.prm.state == DVP_REFINE_INIT) {
		if (cost_now < cost_stored - 0.1) a.planes[center] = plane_now;
		else plane_final = a.planes[center];
	} else {
		a.planes[center] = plane_now;
	}
	// "update cost with old method" (APD.cu:3072-3088): rescore the stored plane with the plain NCC at strong_radius
	{
		// No shared-memory table here: this kernel's fetches are scattered (each lane walks its own anchors), its hit rate
		// lives on the L1/TEX cache, and a 64-thread block's table would take 18 KB of it per block (147 KB per SM).  The
		// <= S NCCs of this rescoring recompute their 36 weights through ncc_cost's general path instead
		// (37.2 -> 34.5 ms per iteration on the bench workload, same bits).
		RefPatch rp;
		rp.prepare(a, px, py, a.prm.strong_radius, nullptr, 0);
		const float2* wt = nullptr;
		float c2 = 0.0f;
		for (int i = 0; i < S; ++i) {
			const int wv = vw.get(i);
			if (wv == 0) continue;
			c2 += wv * ncc_cost<1>(a, a.views[i], a.tex_img[i + 1], px, py, plane_final, rp, wt, 0);
		}
		c2 /= weight_norm;
		a.costs[center] = c2;
	}
}
#undef COST
#undef PROB

// ------------------------------------------------------------------------------------------------------
// every pixel some WEAK pixel's anchor list names (entries 1..10; entry 0 is the pixel itself, whose record nobody reads)
__global__ void __launch_bounds__(256) k_mark_anchors(const __grid_constant__ KArgs a, const int* weak_list, uint8_t* flag) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	const int wi = t / (DVP_NEIGHBOUR_NUM - 1), k = t % (DVP_NEIGHBOUR_NUM - 1) + 1;
	if (wi >= a.weak_count) return;
	const int center = weak_list[wi];
	if (a.weak[center] != DVP_WEAK) return;   // demoted before K4 wrote its list: the list is stale and has no reader
	const short2 np = a.neighbours[(size_t)a.neighbours_map[center] * DVP_NEIGHBOUR_NUM + k];
	if (np.x >= 0 && np.y >= 0 && np.x < a.W && np.y < a.H) flag[np.x + np.y * a.W] = 1;
}

cudaError_t launch_edge_inform(const KArgs& a, bool with_candidates, cudaStream_t st) {
	// part (a) (candidate offsets) is only ever read by the deformable NCC of WEAK pixels; with no WEAK pixel in
	// the view it has no reader and is skipped (the reference computes it regardless: 31 % of its pass at S=4).
	if (a.weak_count > 0 && with_candidates) {
		dim3 b(32, 8);   // the sector table was filled by configure_weak_kernels at upload
		dim3 g((a.W + 31) / 32, (a.H + 7) / 8, 1);
		k_candidate<<<g, b, 0, st>>>(a, nullptr);
	}
	return launch_edge_inform_prep(a, st);
}
cudaError_t launch_gen_neighbours(const KArgs& a, const int* weak_list, uint8_t* anchor_flag, cudaStream_t st) {
	if (a.weak_count == 0) return cudaSuccess;
	k_gen_neighbours<<<(a.weak_count + kK4Threads - 1) / kK4Threads, kK4Threads, 0, st>>>(a, weak_list);
	if (anchor_flag) {   // dvp_run: K2's candidate records, now that the anchors are known (see k_candidate)
		cudaError_t e = cudaMemsetAsync(anchor_flag, 0, (size_t)a.N, st);
		if (e != cudaSuccess) return e;
		const long long pairs = (long long)a.weak_count * (DVP_NEIGHBOUR_NUM - 1);
		k_mark_anchors<<<(int)((pairs + 255) / 256), 256, 0, st>>>(a, weak_list, anchor_flag);
		dim3 b(32, 8);
		dim3 g((a.W + 31) / 32, (a.H + 7) / 8, 1);
		k_candidate<<<g, b, 0, st>>>(a, anchor_flag);
	}
	return cudaGetLastError();
}
cudaError_t launch_ransac_fit(const KArgs& a, const int* weak_list, cudaStream_t st) {
	k_fit_copy_nonweak<<<(a.N + 255) / 256, 256, 0, st>>>(a);
	if (a.weak_count > 0) k_ransac_fit<<<(a.weak_count + 63) / 64, 64, 0, st>>>(a, weak_list);
	return cudaGetLastError();
}
cudaError_t launch_weak_sweep(const KArgs& a, const int* colour_list, int count, int iter, int red, void* scratch, cudaStream_t st) {
	(void)red;
	if (count == 0) return cudaSuccess;
#ifdef DVP_WEAK_NO_SCORE_KERNEL   // A/B: everything in the thread-per-pixel kernel
	scratch = nullptr;
#endif
	float* scored = static_cast<float*>(scratch);   // the K7 / K8 scratch area: >= (9 S + 9) words per pixel of a colour, idle now
	const int stride = (count + 31) & ~31;
	if (scored) k_weak_score<<<(int)(((size_t)count * 8 + 255) / 256), 256, 0, st>>>(a, colour_list, count, scored, stride);
	k_weak_sweep<<<(count + kWeakThreads - 1) / kWeakThreads, kWeakThreads, 0, st>>>(a, colour_list, count, iter, scored, stride);
	return cudaGetLastError();
}
// One-time, per-device fill of the sector table, completed before the flag is set so that contexts on other streams
// (or host threads) of the same device can never launch k_candidate ahead of it.
cudaError_t configure_weak_kernels(int S) {
	(void)S;
	static std::mutex mu;
	static bool table_ready[256] = {false};
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	std::lock_guard<std::mutex> lock(mu);
	if (dev >= 0 && dev < 256 && table_ready[dev]) return cudaSuccess;
	k_fill_sector_table<<<1, dim3(11, 11)>>>();
	if ((e = cudaGetLastError()) != cudaSuccess) return e;
	if ((e = cudaDeviceSynchronize()) != cudaSuccess) return e;
	if (dev >= 0 && dev < 256) table_ready[dev] = true;
	return cudaSuccess;
}

}  // namespace dvp
