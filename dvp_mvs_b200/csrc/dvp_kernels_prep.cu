// dvp_kernels_prep.cu — K2 GenEdgeInform, K3 FindNearestStrongPoint, K5 NeigbourUpdate and the RNG
// exchange helpers (reference APD.cu:3731-3890, 4159-4193, 3713-3729).
#include "dvp_common.cuh"
#include "dvp_launch.h"
#include <cub/cub.cuh>

namespace dvp {

__constant__ int c_dir8[8][2] = {{0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {1, 1}, {-1, 1}, {1, -1}};

// ------------------------------------------------------------------------------------------------------
// K2 part (b): nearest edge pixel along each of 8 rays (APD.cu:3803-3824).
// The reference walks an unbounded ray from every pixel (O(N * max(W,H)) reads).  The answer for pixel q
// along direction d is "the closest edge pixel strictly ahead of q on q's line", so one backwards sweep
// per line produces all answers on that line in O(length): 8N reads in total.  One thread per (direction, line).
__global__ void __launch_bounds__(128) k_edge_neigh_sweep(const __grid_constant__ KArgs a) {
	const int W = a.W, H = a.H;
	const int n_diag = W + H - 1;
	const int counts[8] = {W, W, H, H, n_diag, n_diag, n_diag, n_diag};
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	int d = 0;
	while (d < 8 && t >= counts[d]) { t -= counts[d]; ++d; }
	if (d >= 8) return;
	const int dx = c_dir8[d][0], dy = c_dir8[d][1];
	// far end of the line in direction +d
	int x, y;
	if (dx == 0) { x = t; y = (dy > 0) ? H - 1 : 0; }
	else if (dy == 0) { y = t; x = (dx > 0) ? W - 1 : 0; }
	else {
		const int ex = (dx > 0) ? W - 1 : 0, ey = (dy > 0) ? H - 1 : 0;
		if (t < W) { x = t; y = ey; }
		else { const int k = t - W; x = ex; y = (ey == 0) ? k + 1 : k; }
	}
	short2 last = make_short2(-1, -1);
	while (x >= 0 && x < W && y >= 0 && y < H) {
		const int q = y * W + x;
		a.edge_neigh[(size_t)q * DVP_EDGE_NEIGH_NUM + d] = last;
		if (a.edge[q]) last = make_short2((short)x, (short)y);
		x -= dx; y -= dy;
	}
}

// K2 parts (c), (d), (e): per-pixel edge density, label-region boundaries, demotions (APD.cu:3826-3889).
__global__ void __launch_bounds__(256) k_edge_inform_pixel(const __grid_constant__ KArgs a) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	const int W = a.W, H = a.H;
	if (x >= W || y >= H) return;
	const int center = x + y * W;
	uint8_t state = a.weak[center];
	if (a.prm.use_edge) {
		if (state == DVP_WEAK) {
			const int radius = a.prm.strong_radius;
			int edge_pix = 0, tot_pix = 0;
			for (int i = -radius; i <= radius; i++)
				for (int j = -radius; j <= radius; j++) {
					const int nx = x + i, ny = y + j;
					if (nx < 0 || nx >= W || ny < 0 || ny >= H) continue;
					if (a.edge[ny * W + nx]) edge_pix++;
					tot_pix++;
				}
			const float density = 1.0f * edge_pix / tot_pix;
			a.complex_[a.neighbours_map[center]] = 1.0f / (1.0f + exp(-25.0 * (density - 0.35)));
		}
		if (a.prm.state == DVP_REFINE_INIT && a.prm.use_detail && a.edge[center]) {
			if (state != DVP_STRONG) { state = DVP_UNKNOWN; a.weak[center] = DVP_UNKNOWN; }
		}
	}
	if (a.prm.use_label && state == DVP_WEAK) {
		// the label-boundary rays of part (d) are produced by k_label_boundary_sweep (launched before this kernel)
		if (a.prm.state == DVP_REFINE_INIT && a.prm.use_detail && a.label[center] == 0) {
			if (state != DVP_STRONG) a.weak[center] = DVP_UNKNOWN;
		}
	}
}

// K2 part (d) (APD.cu:3855-3885): for every WEAK pixel with a positive label, the LAST pixel carrying the same
// label along each of 8 rays before the ray leaves the image or meets a label of -1.  The reference walks every
// ray from every WEAK pixel (rays are unbounded: thousands of reads per pixel inside a large region).  Walking
// each image line ONCE from its far end, "last same-label pixel ahead of p" is the first position at which that
// label was met since the last -1 barrier; a 12-entry (label, position) table per line holds it.  If a segment
// ever carries more than 12 distinct labels, pixels whose label is not in the table fall back to the reference's
// forward walk, so the result is exact in every case.
__global__ void __launch_bounds__(128) k_label_boundary_sweep(const __grid_constant__ KArgs a) {
	const int W = a.W, H = a.H;
	const int n_diag = W + H - 1;
	const int counts[8] = {W, W, H, H, n_diag, n_diag, n_diag, n_diag};
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	int d = 0;
	while (d < 8 && t >= counts[d]) { t -= counts[d]; ++d; }
	if (d >= 8) return;
	const int dx = c_dir8[d][0], dy = c_dir8[d][1];
	int x, y;
	if (dx == 0) { x = t; y = (dy > 0) ? H - 1 : 0; }
	else if (dy == 0) { y = t; x = (dx > 0) ? W - 1 : 0; }
	else {
		const int ex = (dx > 0) ? W - 1 : 0, ey = (dy > 0) ? H - 1 : 0;
		if (t < W) { x = t; y = ey; }
		else { const int k = t - W; x = ex; y = (ey == 0) ? k + 1 : k; }
	}
	constexpr int CAP = 12;
	int lab[CAP]; short2 pos[CAP];
	int n = 0;
	bool overflow = false;
	const bool demote_edges = a.prm.use_edge && a.prm.state == DVP_REFINE_INIT && a.prm.use_detail;
	while (x >= 0 && x < W && y >= 0 && y < H) {
		const int q = y * W + x;
		const int L = a.label[q];
		int hit = -1;
		for (int k = 0; k < n; ++k) if (lab[k] == L) { hit = k; break; }
		// same condition as the reference thread: still WEAK after its own edge demotion (APD.cu:3847-3855)
		if (L > 0 && a.weak[q] == DVP_WEAK && !(demote_edges && a.edge[q])) {
			short2 out = make_short2(-1, -1);
			if (hit >= 0) out = pos[hit];
			else if (overflow) {   // exact fallback: the reference's forward walk
				int nx = x + dx, ny = y + dy;
				while (nx >= 0 && nx < W && ny >= 0 && ny < H) {
					const int nl = a.label[nx + ny * W];
					if (nl == L) out = make_short2((short)nx, (short)ny);
					else if (nl == -1) break;
					nx += dx; ny += dy;
				}
			}
			a.label_boundary[(size_t)a.neighbours_map[q] * DVP_LAB_BOUNDARY_NUM + d] = out;
		}
		if (L == -1) { n = 0; overflow = false; }
		else if (hit < 0) {
			if (n < CAP) { lab[n] = L; pos[n] = make_short2((short)x, (short)y); ++n; }
			else overflow = true;
		}
		x -= dx; y -= dy;
	}
}

// ------------------------------------------------------------------------------------------------------
// K3 FindNearestStrongPoint (APD.cu:4159-4193): first STRONG pixel on growing square rings, radius <= 100,
// scan order x-major then y inside each ring (the tie order is part of the result).
// The reference reads every pixel of every ring (up to 40 401 reads per WEAK pixel).  Here two pointer maps are
// built by line sweeps — next STRONG pixel at or to the right in the row, at or below in the column — and each
// ring is answered with four look-ups that reproduce the reference's scan order exactly:
//   left column (all dy ascending), then the top/bottom rows interleaved by ascending dx (top first), then the
//   right column.
constexpr short kNone = 32767;
__global__ void __launch_bounds__(128) k_next_strong_sweeps(const __grid_constant__ KArgs a, short* next_right, short* next_down) {
	const int W = a.W, H = a.H;
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t < W) {            // column t, bottom to top (coalesced across threads)
		short last = kNone;
		for (int y = H - 1; y >= 0; --y) {
			const int q = y * W + t;
			if (a.weak[q] == DVP_STRONG) last = (short)y;
			next_down[q] = last;
		}
	} else if (t < W + H) { // row t - W, right to left
		const int y = t - W;
		short last = kNone;
		for (int x = W - 1; x >= 0; --x) {
			const int q = y * W + x;
			if (a.weak[q] == DVP_STRONG) last = (short)x;
			next_right[q] = last;
		}
	}
}
__global__ void __launch_bounds__(256) k_nearest_strong(const __grid_constant__ KArgs a, const short* next_right, const short* next_down) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	const int W = a.W, H = a.H;
	if (x >= W || y >= H) return;
	const int center = x + y * W;
	short2 out = make_short2(-1, -1);
	if (a.weak[center] == DVP_WEAK) {
		const int max_radius = 100;
		for (int r = 0; r <= max_radius; ++r) {
			const int y_lo = max(y - r, 0), y_hi = min(y + r, H - 1);
			// dx = -r : whole column, dy ascending
			if (x - r >= 0) {
				const int yn = next_down[y_lo * W + (x - r)];
				if (yn <= y_hi) { out = make_short2((short)(x - r), (short)yn); break; }
			}
			if (r > 0) {
				// -r < dx < r : dy = -r, then dy = +r, for ascending dx
				const int x_lo = max(x - r + 1, 0), x_hi = min(x + r - 1, W - 1);
				if (x_lo <= x_hi) {
					int xt = kNone, xb = kNone;
					if (y - r >= 0) xt = next_right[(y - r) * W + x_lo];
					if (y + r < H) xb = next_right[(y + r) * W + x_lo];
					if (xt > x_hi) xt = kNone;
					if (xb > x_hi) xb = kNone;
					if (xt != kNone || xb != kNone) {
						if (xt <= xb) out = make_short2((short)xt, (short)(y - r));   // same dx: the top row comes first
						else out = make_short2((short)xb, (short)(y + r));
						break;
					}
				}
				// dx = +r : whole column
				if (x + r < W) {
					const int yn = next_down[y_lo * W + (x + r)];
					if (yn <= y_hi) { out = make_short2((short)(x + r), (short)yn); break; }
				}
			}
		}
	}
	a.nearest_strong[center] = out;
}

// K5 NeigbourUpdate (APD.cu:3713-3729)
__global__ void __launch_bounds__(256) k_neighbour_update(const __grid_constant__ KArgs a) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.N) return;
	if (a.weak[i] != DVP_WEAK) return;
	if (a.weak_reliable[i] != 1) a.weak[i] = DVP_UNKNOWN;
}

__global__ void k_rng_export(const __grid_constant__ KArgs a, uint32_t* dst) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.N) return;
	for (int k = 0; k < 6; ++k) dst[(size_t)6 * i + k] = a.rng[(size_t)k * a.N + i];
}
__global__ void k_rng_import(const __grid_constant__ KArgs a, const uint32_t* src) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.N) return;
	for (int k = 0; k < 6; ++k) a.rng[(size_t)k * a.N + i] = src[(size_t)6 * i + k];
}

__global__ void k_fill_i32(int32_t* dst, int32_t v, int n) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) dst[i] = v;
}
// Chessboard (L-infinity) distance from every pixel to the nearest edge pixel, capped at 255, for the edge walks of K4 / K9
// (bresenham_crosses_edge, dvp_weak.cuh): a walk standing on a pixel at distance d knows its next d - 1 steps are edge-free.
// Built from a summed-area table of the edge map: "is the (2r+1)^2 box around p edge-free" is four reads, the largest such
// r is found by bisection.
__global__ void __launch_bounds__(128) k_edge_sat_rows(const uint8_t* __restrict__ edge, int W, int H, int* __restrict__ sat) {
	// one warp per image row: inclusive prefix count of edge pixels, stored at sat[(y + 1) * (W + 1) + x + 1]; row 0 and column 0 are zero
	const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (warp > H) return;
	int* out = sat + (size_t)warp * (W + 1);
	if (warp == 0) { for (int x = lane; x <= W; x += 32) out[x] = 0; return; }
	const uint8_t* row = edge + (size_t)(warp - 1) * W;
	int carry = 0;
	if (lane == 0) out[0] = 0;
	for (int x0 = 0; x0 < W; x0 += 32) {
		const int x = x0 + lane;
		int v = (x < W && row[x]) ? 1 : 0;
		for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
		if (x < W) out[x + 1] = carry + v;
		carry += __shfl_sync(0xffffffffu, v, 31);
	}
}
__global__ void __launch_bounds__(128) k_edge_sat_cols(int W, int H, int* __restrict__ sat) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	if (x > W) return;
	int acc = 0;
	for (int y = 0; y <= H; ++y) { acc += sat[(size_t)y * (W + 1) + x]; sat[(size_t)y * (W + 1) + x] = acc; }
}
__global__ void __launch_bounds__(256) k_edge_distance(const uint8_t* __restrict__ edge, int W, int H, const int* __restrict__ sat, uint8_t* __restrict__ dist) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= W || y >= H) return;
	if (edge[(size_t)y * W + x]) { dist[(size_t)y * W + x] = 0; return; }
	auto clear = [&](int r) -> bool {   // no edge pixel with max(|dx|, |dy|) <= r (pixels outside the image hold no edge)
		const int x1 = max(x - r, 0), x2 = min(x + r, W - 1) + 1, y1 = max(y - r, 0), y2 = min(y + r, H - 1) + 1;
		const int* a = sat + (size_t)y1 * (W + 1); const int* b = sat + (size_t)y2 * (W + 1);
		return b[x2] - b[x1] - a[x2] + a[x1] == 0;
	};
	int lo = 0, hi = 254;          // clear(0) holds (the pixel itself is no edge); find the largest r <= 254 with clear(r)
	if (clear(hi)) lo = hi;
	else while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (clear(mid)) lo = mid; else hi = mid; }
	dist[(size_t)y * W + x] = (uint8_t)(lo + 1);   // the nearest edge pixel is lo + 1 away (or farther: capped at 255)
}
cudaError_t launch_edge_distance(const uint8_t* edge, int W, int H, int* sat, uint8_t* dist, cudaStream_t st) {
	k_edge_sat_rows<<<((H + 1) * 32 + 127) / 128, 128, 0, st>>>(edge, W, H, sat);
	k_edge_sat_cols<<<(W + 1 + 127) / 128, 128, 0, st>>>(W, H, sat);
	dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8);
	k_edge_distance<<<g, b, 0, st>>>(edge, W, H, sat, dist);
	return cudaGetLastError();
}

cudaError_t launch_fill_i32(int32_t* dst, int32_t v, int n, cudaStream_t st) {
	k_fill_i32<<<(n + 255) / 256, 256, 0, st>>>(dst, v, n);
	return cudaGetLastError();
}


// ------------------------------------------------------------------------------------------------------
// WEAK-pixel indexing on the device: neighbours_map[p] = rank of p among the WEAK pixels in raster order
// (what InuputInitialization builds on the host, APD.cpp:1182-1193) and its inverse, the compact WEAK list.
constexpr int kScanBlock = 1024;
// colour < 0: every WEAK pixel; colour 0/1: WEAK pixels with (x + y) % 2 == colour that the reference's half grid reaches
__device__ __forceinline__ bool weak_pred(const uint8_t* weak, int p, int n, int W, int colour, int yy_limit) {
	if (p >= n || weak[p] != DVP_WEAK) return false;
	if (colour < 0) return true;
	const int y = p / W, x = p - y * W;
	return ((x + y) & 1) == colour && (y >> 1) < yy_limit;
}
__global__ void __launch_bounds__(256) k_weak_count_blocks(const uint8_t* weak, int n, int W, int colour, int yy_limit, int* block_sums) {
	__shared__ int s[8];
	const int base = blockIdx.x * kScanBlock;
	int c = 0;
	for (int i = threadIdx.x; i < kScanBlock; i += 256) { if (weak_pred(weak, base + i, n, W, colour, yy_limit)) ++c; }
	for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
	if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
	__syncthreads();
	if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += s[w]; block_sums[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(1024) k_weak_scan_blocks(int* block_sums, int nblocks, int* total) {
	// exclusive scan of the per-block counts by one CTA (nblocks <= ~1e6 / 1024 chunks handled serially per thread)
	__shared__ int s[1024];
	const int per = (nblocks + 1023) / 1024;
	const int lo = threadIdx.x * per, hi = min(lo + per, nblocks);
	int sum = 0;
	for (int i = lo; i < hi; ++i) sum += block_sums[i];
	s[threadIdx.x] = sum;
	__syncthreads();
	for (int o = 1; o < 1024; o <<= 1) {
		const int v = (threadIdx.x >= o) ? s[threadIdx.x - o] : 0;
		__syncthreads();
		s[threadIdx.x] += v;
		__syncthreads();
	}
	int run = s[threadIdx.x] - sum;
	for (int i = lo; i < hi; ++i) { const int c = block_sums[i]; block_sums[i] = run; run += c; }
	if (threadIdx.x == 1023) *total = s[1023];
}
__global__ void __launch_bounds__(256) k_weak_index(const uint8_t* weak, int n, int W, int colour, int yy_limit, const int* block_offsets, int* nmap, int* weak_list) {
	// one CTA per 1024-pixel chunk; thread t owns 4 consecutive pixels so that raster order is preserved
	__shared__ int s[256];
	const int base = blockIdx.x * kScanBlock + threadIdx.x * 4;
	int f[4], c = 0;
#pragma unroll
	for (int k = 0; k < 4; ++k) { f[k] = weak_pred(weak, base + k, n, W, colour, yy_limit) ? 1 : 0; c += f[k]; }
	s[threadIdx.x] = c;
	__syncthreads();
	for (int o = 1; o < 256; o <<= 1) {
		const int v = (threadIdx.x >= o) ? s[threadIdx.x - o] : 0;
		__syncthreads();
		s[threadIdx.x] += v;
		__syncthreads();
	}
	int run = block_offsets[blockIdx.x] + s[threadIdx.x] - c;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const int p = base + k;
		if (p < n) {
			if (nmap) nmap[p] = f[k] ? run : 0;
			if (f[k]) { if (weak_list) weak_list[run] = p; ++run; }
		}
	}
}
__global__ void k_reset_unknown_radius(const uint8_t* weak, int32_t* radius, int32_t strong_radius, int n) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && weak[i] == DVP_UNKNOWN) radius[i] = strong_radius;   // APD.cpp:1663-1666
}
cudaError_t launch_weak_count(const uint8_t* weak, int n, int W, int colour, int yy_limit, int* block_sums, int* total, cudaStream_t st) {
	const int nb = (n + kScanBlock - 1) / kScanBlock;
	k_weak_count_blocks<<<nb, 256, 0, st>>>(weak, n, W, colour, yy_limit, block_sums);
	k_weak_scan_blocks<<<1, 1024, 0, st>>>(block_sums, nb, total);
	return cudaGetLastError();
}
cudaError_t launch_weak_index(const uint8_t* weak, int n, int W, int colour, int yy_limit, const int* block_offsets, int* nmap, int* weak_list, cudaStream_t st) {
	const int nb = (n + kScanBlock - 1) / kScanBlock;
	k_weak_index<<<nb, 256, 0, st>>>(weak, n, W, colour, yy_limit, block_offsets, nmap, weak_list);
	return cudaGetLastError();
}
// Reorder a raster-ordered pixel list into tiles of tile_w x tile_h pixels (stable: raster order inside a tile).  The
// WEAK kernels take 64 consecutive list entries per block; with compact tiles the anchors and patches a block touches
// overlap, which is what their scattered fetches need from the L1.  Results do not depend on the order (one thread
// owns one pixel and same-colour WEAK pixels never read each other).
__global__ void __launch_bounds__(256) k_tile_keys(const int* __restrict__ list, int count, int W, int tile_w, int tile_h, int tiles_x, int* __restrict__ keys) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	const int p = list[i];
	const int y = p / W, x = p - y * W;
	keys[i] = (y / tile_h) * tiles_x + (x / tile_w);
}
size_t tile_order_temp_bytes(int count) {
	size_t bytes = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int*)nullptr, (int*)nullptr, (const int*)nullptr, (int*)nullptr, count);
	return bytes;
}
cudaError_t launch_tile_order(int* list, int count, int W, int H, int tile_w, int tile_h, int* keys_in, int* keys_out, int* vals_out,
                              void* temp, size_t temp_bytes, cudaStream_t st) {
	if (count <= 1) return cudaSuccess;
	const int tiles_x = (W + tile_w - 1) / tile_w, tiles_y = (H + tile_h - 1) / tile_h;
	int bits = 1;
	while ((1ll << bits) < (long long)tiles_x * tiles_y) ++bits;
	k_tile_keys<<<(count + 255) / 256, 256, 0, st>>>(list, count, W, tile_w, tile_h, tiles_x, keys_in);
	cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, (const int*)keys_in, keys_out, (const int*)list, vals_out, count, 0, bits, st);
	if (e != cudaSuccess) return e;
	return cudaMemcpyAsync(list, vals_out, (size_t)count * sizeof(int), cudaMemcpyDeviceToDevice, st);
}
cudaError_t launch_reset_unknown_radius(const uint8_t* weak, int32_t* radius, int32_t strong_radius, int n, cudaStream_t st) {
	k_reset_unknown_radius<<<(n + 255) / 256, 256, 0, st>>>(weak, radius, strong_radius, n);
	return cudaGetLastError();
}

cudaError_t launch_edge_inform_prep(const KArgs& a, cudaStream_t st) {
	if (a.prm.use_edge) {
		const int total = 2 * a.W + 2 * a.H + 4 * (a.W + a.H - 1);
		k_edge_neigh_sweep<<<(total + 127) / 128, 128, 0, st>>>(a);
	}
	if (a.prm.use_label && a.weak_count > 0) {
		const int total = 2 * a.W + 2 * a.H + 4 * (a.W + a.H - 1);
		k_label_boundary_sweep<<<(total + 127) / 128, 128, 0, st>>>(a);   // reads the pixel states before the demotions below
	}
	dim3 b(32, 8);
	dim3 g((a.W + 31) / 32, (a.H + 7) / 8, 1);
	k_edge_inform_pixel<<<g, b, 0, st>>>(a);
	return cudaGetLastError();
}
cudaError_t launch_nearest_strong(const KArgs& a, short* next_right, short* next_down, cudaStream_t st) {
	dim3 b(32, 8);
	dim3 g((a.W + 31) / 32, (a.H + 7) / 8, 1);
	if (a.weak_count > 0) k_next_strong_sweeps<<<(a.W + a.H + 127) / 128, 128, 0, st>>>(a, next_right, next_down);
	k_nearest_strong<<<g, b, 0, st>>>(a, next_right, next_down);   // non-WEAK pixels only store (-1,-1)
	return cudaGetLastError();
}
cudaError_t launch_neighbour_update(const KArgs& a, cudaStream_t st) {
	k_neighbour_update<<<(a.N + 255) / 256, 256, 0, st>>>(a);
	return cudaGetLastError();
}
cudaError_t launch_rng_export(const KArgs& a, uint32_t* dst, cudaStream_t st) {
	k_rng_export<<<(a.N + 255) / 256, 256, 0, st>>>(a, dst);
	return cudaGetLastError();
}
cudaError_t launch_rng_import(const KArgs& a, const uint32_t* src, cudaStream_t st) {
	k_rng_import<<<(a.N + 255) / 256, 256, 0, st>>>(a, src);
	return cudaGetLastError();
}

}  // namespace dvp
