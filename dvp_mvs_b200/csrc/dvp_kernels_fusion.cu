// dvp_kernels_fusion.cu — depth-map fusion on the device (reference RunFusion, the "ETH version"
// main() calls, APD.cpp:1809-1960 with Get3DPointonWorld APD.cpp:501-525, ProjectCamera APD.cpp:536-546, GetAngle
// APD.cpp:1797-1806, ExportPointCloud APD.cpp:842-882).
//
// The reference is one sequential loop: views in order, pixels in raster order; a pixel that is accepted masks the
// source pixels that agreed with it, and a masked source pixel is ignored by every later pixel — so the result
// depends on the visiting order.  That order's result is reproduced exactly, in parallel, in three stages per view:
//
//   candidates  everything that does not depend on the masks: per (pixel, source) the source cell the pixel would
//               claim and its exp(-index) term.  One thread per pixel, 2*S projections, the only floating-point stage.
//   resolve     deterministic reservations.  Round: (A) every undecided pixel drops the cells that are masked by now
//               and writes its raster index into each remaining cell with atomicMin; (B) a pixel that holds ALL its
//               cells has no earlier undecided competitor, so what it sees is what the sequential loop would show
//               it: it decides (accept test on the surviving terms, in source order), masks its cells if accepted and
//               releases its reservations.  The smallest undecided pixel always decides, so rounds terminate; a pixel
//               that never shares a cell decides in round 1.  Integer work only: bit-exact by construction.
//   emit        accepted pixels → exclusive scan → points written in raster order (the reference's push order).
//
// This file is compiled WITHOUT --use_fast_math and with -fmad=false (csrc/Makefile): the reference's fusion is host
// C++, so +, -, *, / and sqrt are IEEE-rounded one at a time here as in the CPU restatement (oracle/cpu/fusion_cpu.cpp).
// exp and acos are evaluated in double and rounded to float (libm's expf / acosf differ from that by at most one ulp).
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <stdint.h>
#include <climits>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/dvp_mvs.h"

namespace dvp_fuse {

constexpr int kMaxSrc = 32;
constexpr unsigned kFree = 0xFFFFFFFFu;
enum : uint8_t { ST_DECIDED = 0, ST_ACTIVE = 1 };

struct ViewDev {
	dvp_camera cam;
	int w, h;
	const float* depth;
	const float* normal;
	const uint8_t* image;
	const uint8_t* weak;
	const uint8_t* block;
	uint8_t* mask;       // fusion mask, APD.cpp:1867
	unsigned* resv;      // reservation word per pixel: raster index of the smallest undecided claimant, or kFree
};

struct RefArgs {
	int ref, S, N;
	int src[kMaxSrc];
};

struct F3 { float x, y, z; };

// APD.cpp:501-525
__device__ __forceinline__ F3 point_on_world(int x, int y, float depth, const dvp_camera& cam) {
	F3 p, t, c;
	p.x = depth * ((float)x - cam.K[2]) / cam.K[0];
	p.y = depth * ((float)y - cam.K[5]) / cam.K[4];
	p.z = depth;
	t.x = cam.R[0] * p.x + cam.R[3] * p.y + cam.R[6] * p.z;
	t.y = cam.R[1] * p.x + cam.R[4] * p.y + cam.R[7] * p.z;
	t.z = cam.R[2] * p.x + cam.R[5] * p.y + cam.R[8] * p.z;
	c.x = -(cam.R[0] * cam.t[0] + cam.R[3] * cam.t[1] + cam.R[6] * cam.t[2]);
	c.y = -(cam.R[1] * cam.t[0] + cam.R[4] * cam.t[1] + cam.R[7] * cam.t[2]);
	c.z = -(cam.R[2] * cam.t[0] + cam.R[5] * cam.t[1] + cam.R[8] * cam.t[2]);
	p.x = t.x + c.x;
	p.y = t.y + c.y;
	p.z = t.z + c.z;
	return p;
}

// APD.cpp:536-546
__device__ __forceinline__ void project(const F3& X, const dvp_camera& cam, float& px, float& py, float& depth) {
	F3 t;
	t.x = cam.R[0] * X.x + cam.R[1] * X.y + cam.R[2] * X.z + cam.t[0];
	t.y = cam.R[3] * X.x + cam.R[4] * X.y + cam.R[5] * X.z + cam.t[1];
	t.z = cam.R[6] * X.x + cam.R[7] * X.y + cam.R[8] * X.z + cam.t[2];
	depth = cam.K[6] * t.x + cam.K[7] * t.y + cam.K[8] * t.z;
	px = (cam.K[0] * t.x + cam.K[1] * t.y + cam.K[2] * t.z) / depth;
	py = (cam.K[3] * t.x + cam.K[4] * t.y + cam.K[5] * t.z) / depth;
}

// int(v) as the host evaluates it: truncation; NaN and out-of-range values become INT_MIN (x86 cvttss2si), which fails
// every bounds test.  (The device's own conversion would turn NaN into 0.)
__device__ __forceinline__ int to_int(float v) {
	if (!(v > -2147483648.0f && v < 2147483648.0f)) return INT_MIN;
	return (int)v;
}

// The mask-independent part of one (pixel, source) evaluation, shared by the three fusion variants (APD.cpp:1901-1921,
// 2067-2085, 2231-2251): the source cell the pixel projects to and, if that cell holds a depth, the three measures.
// Returns the cell or -1 where the reference skips the source.
__device__ __forceinline__ int measures(const ViewDev& rv, const ViewDev& sv, int r, int c, float ref_depth, const F3& X,
		float n0, float n1, float n2, float& reproj_error, float& relative_depth_diff, float& angle) {
	float px, py, proj_depth;
	project(X, sv.cam, px, py, proj_depth);
	const int src_r = to_int(py + 0.5f), src_c = to_int(px + 0.5f);
	if (!(src_c >= 0 && src_c < sv.w && src_r >= 0 && src_r < sv.h)) return -1;
	const int q = src_r * sv.w + src_c;
	const float src_depth = sv.depth[q];
	if (src_depth <= 0.0f) return -1;
	const F3 Y = point_on_world(src_c, src_r, src_depth, sv.cam);
	float qx, qy;
	project(Y, rv.cam, qx, qy, proj_depth);
	const double dx = (double)((float)c - qx), dy = (double)((float)r - qy);
	reproj_error = (float)sqrt(dx * dx + dy * dy);
	relative_depth_diff = fabsf(proj_depth - ref_depth) / ref_depth;
	const float dot = n0 * sv.normal[3 * (size_t)q] + n1 * sv.normal[3 * (size_t)q + 1] + n2 * sv.normal[3 * (size_t)q + 2];
	angle = (float)acos((double)dot);
	if (angle != angle) angle = 0.0f;                                  // APD.cpp:1802-1803
	return q;
}

// ---- stage 1: candidates ------------------------------------------------------------------------------------------
// cells / terms are [S][N] (one coalesced plane per source).  live[p]: bit j = cell j is a candidate.
__global__ void __launch_bounds__(256) k_fuse_candidates(const __grid_constant__ RefArgs a, const ViewDev* __restrict__ views,
		int* __restrict__ cells, float* __restrict__ terms, unsigned* __restrict__ live, unsigned* __restrict__ used,
		uint8_t* __restrict__ state, int* __restrict__ list, int* __restrict__ count) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	bool active = false;
	if (p < a.N) {
		const ViewDev& rv = views[a.ref];
		const int r = p / rv.w, c = p - r * rv.w;
		unsigned bits = 0;
		const float ref_depth = rv.depth[p];
		const bool skip = (rv.block && rv.block[p] < 128) || ref_depth <= 0.0f;   // APD.cpp:1884-1886, 1892-1894
		F3 X = {0.f, 0.f, 0.f};
		float n0 = 0.f, n1 = 0.f, n2 = 0.f;
		if (!skip) {
			X = point_on_world(c, r, ref_depth, rv.cam);
			n0 = rv.normal[3 * (size_t)p]; n1 = rv.normal[3 * (size_t)p + 1]; n2 = rv.normal[3 * (size_t)p + 2];
		}
		for (int j = 0; j < a.S; ++j) {
			int cell = -1;
			float term = 0.0f;
			if (!skip) {
				float reproj_error, relative_depth_diff, angle;
				const int q = measures(rv, views[a.src[j]], r, c, ref_depth, X, n0, n1, n2, reproj_error, relative_depth_diff, angle);
				if (q >= 0 && reproj_error < 2.0f && relative_depth_diff < 0.01f && angle < 0.174533f) {
					const float tmp_index = reproj_error + 200.0f * relative_depth_diff + angle * 10.0f;
					term = (float)exp((double)(-tmp_index));
					cell = q;
					bits |= 1u << j;
				}
			}
			cells[(size_t)j * a.N + p] = cell;
			terms[(size_t)j * a.N + p] = term;
		}
		live[p] = bits;
		used[p] = 0;
		active = bits != 0 && rv.mask[p] != 1;    // a pixel masked by an earlier view is skipped, APD.cpp:1888-1890
		state[p] = active ? ST_ACTIVE : ST_DECIDED;
	}
	// warp-aggregated append to the undecided list (its order does not matter: priority is the pixel index)
	const unsigned m = __ballot_sync(0xffffffffu, active);
	if (m) {
		const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
		int base = 0;
		if (lane == leader) base = atomicAdd(count, __popc(m));
		base = __shfl_sync(0xffffffffu, base, leader);
		if (active) list[base + __popc(m & ((1u << lane) - 1u))] = p;
	}
}

// ---- stage 2: resolve ----------------------------------------------------------------------------------------------
// Mutable words (mask, resv, state, live) are read with ld.cg: the single-block tail loops over rounds inside one kernel.
__device__ __forceinline__ void reserve_one(const RefArgs& a, const ViewDev* views, const int* cells, unsigned* live, int p) {
	unsigned bits = __ldcg(&live[p]);
	const unsigned before = bits;
	for (unsigned rest = bits; rest;) {
		const int j = __ffs(rest) - 1;
		rest &= rest - 1;
		const ViewDev& sv = views[a.src[j]];
		const int cell = cells[(size_t)j * a.N + p];
		if (__ldcg(&sv.mask[cell]) == 1) bits &= ~(1u << j);           // claimed by an earlier pixel: APD.cpp:1911-1912
		else atomicMin(&sv.resv[cell], (unsigned)p);
	}
	if (bits != before) __stcg(&live[p], bits);
}

// returns true if p is still undecided after this round
__device__ __forceinline__ bool decide_one(const RefArgs& a, const ViewDev* views, const int* cells, const float* terms,
		unsigned* live, unsigned* used, uint8_t* state, int p) {
	const unsigned bits = __ldcg(&live[p]);
	for (unsigned rest = bits; rest;) {
		const int j = __ffs(rest) - 1;
		rest &= rest - 1;
		if (__ldcg(&views[a.src[j]].resv[cells[(size_t)j * a.N + p]]) != (unsigned)p) return true;   // an earlier pixel is still undecided
	}
	// APD.cpp:1925-1934: the surviving sources, in source order
	int num_consistent = 0;
	float dynamic_consistency = 0.0f;
	for (unsigned rest = bits; rest;) {
		const int j = __ffs(rest) - 1;
		rest &= rest - 1;
		dynamic_consistency += terms[(size_t)j * a.N + p];
		num_consistent++;
	}
	const float factor = views[a.ref].weak[p] == DVP_WEAK ? 0.45f : 0.3f;
	const bool accepted = num_consistent >= 1 && dynamic_consistency > factor * (float)num_consistent;
	for (unsigned rest = bits; rest;) {
		const int j = __ffs(rest) - 1;
		rest &= rest - 1;
		const ViewDev& sv = views[a.src[j]];
		const int cell = cells[(size_t)j * a.N + p];
		if (accepted) __stcg(&sv.mask[cell], (uint8_t)1);             // APD.cpp:1942
		__stcg(&sv.resv[cell], kFree);
	}
	used[p] = accepted ? bits : 0u;
	__stcg(&state[p], (uint8_t)ST_DECIDED);
	return false;
}

__global__ void __launch_bounds__(256) k_fuse_reserve(const __grid_constant__ RefArgs a, const ViewDev* __restrict__ views,
		const int* __restrict__ cells, unsigned* live, const uint8_t* state, const int* __restrict__ list, int count) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	const int p = list[i];
	if (__ldcg(&state[p]) == ST_ACTIVE) reserve_one(a, views, cells, live, p);
}

__global__ void __launch_bounds__(256) k_fuse_decide(const __grid_constant__ RefArgs a, const ViewDev* __restrict__ views,
		const int* __restrict__ cells, const float* __restrict__ terms, unsigned* live, unsigned* used, uint8_t* state,
		const int* __restrict__ list, int count) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	const int p = list[i];
	if (__ldcg(&state[p]) == ST_ACTIVE) decide_one(a, views, cells, terms, live, used, state, p);
}

__global__ void __launch_bounds__(256) k_fuse_compact(const uint8_t* __restrict__ state, const int* __restrict__ list, int count,
		int* __restrict__ out, int* __restrict__ out_count) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	int p = -1;
	bool keep = false;
	if (i < count) { p = list[i]; keep = state[p] == ST_ACTIVE; }
	const unsigned m = __ballot_sync(0xffffffffu, keep);
	if (m) {
		const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
		int base = 0;
		if (lane == leader) base = atomicAdd(out_count, __popc(m));
		base = __shfl_sync(0xffffffffu, base, leader);
		if (keep) out[base + __popc(m & ((1u << lane) - 1u))] = p;
	}
}

// the last few undecided pixels (long dependency chains): one block runs the rounds back to back
constexpr int kTailThreads = 1024;
__global__ void __launch_bounds__(kTailThreads) k_fuse_tail(const __grid_constant__ RefArgs a, const ViewDev* __restrict__ views,
		const int* __restrict__ cells, const float* __restrict__ terms, unsigned* live, unsigned* used, uint8_t* state,
		const int* __restrict__ list, int count, int* __restrict__ rounds_out) {
	__shared__ int remaining;
	int rounds = 0;
	for (;;) {
		if (threadIdx.x == 0) remaining = 0;
		for (int i = threadIdx.x; i < count; i += kTailThreads) {
			const int p = list[i];
			if (__ldcg(&state[p]) == ST_ACTIVE) reserve_one(a, views, cells, live, p);
		}
		__threadfence();
		__syncthreads();
		int mine = 0;
		for (int i = threadIdx.x; i < count; i += kTailThreads) {
			const int p = list[i];
			if (__ldcg(&state[p]) == ST_ACTIVE && decide_one(a, views, cells, terms, live, used, state, p)) ++mine;
		}
		if (mine) atomicAdd(&remaining, mine);
		__threadfence();
		__syncthreads();
		++rounds;
		const int left = remaining;
		__syncthreads();
		if (left == 0) break;
		if (rounds > count) { rounds = -1; break; }   // cannot happen: every round decides its smallest pixel
	}
	if (threadIdx.x == 0) *rounds_out = rounds;
}

// ---- stage 3: emit --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fuse_flags(int n, const unsigned* __restrict__ used, int* __restrict__ flags) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p < n) flags[p] = used[p] != 0u;
}

// APD.cpp:1935-1953: the point is the reference pixel's own 3-D point; the colour is the mean over the reference pixel
// and the sources that agreed, summed in source order.
__global__ void __launch_bounds__(256) k_fuse_emit(const __grid_constant__ RefArgs a, const ViewDev* __restrict__ views,
		const int* __restrict__ cells, const unsigned* __restrict__ used, const int* __restrict__ offs, float* __restrict__ points) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= a.N) return;
	const unsigned bits = used[p];
	if (!bits) return;
	const ViewDev& rv = views[a.ref];
	const int r = p / rv.w, c = p - r * rv.w;
	const F3 X = point_on_world(c, r, rv.depth[p], rv.cam);
	float col0 = (float)rv.image[3 * (size_t)p], col1 = (float)rv.image[3 * (size_t)p + 1], col2 = (float)rv.image[3 * (size_t)p + 2];
	int num_consistent = 0;
	for (unsigned rest = bits; rest;) {
		const int j = __ffs(rest) - 1;
		rest &= rest - 1;
		const uint8_t* s = views[a.src[j]].image + 3 * (size_t)cells[(size_t)j * a.N + p];
		col0 += (float)s[0]; col1 += (float)s[1]; col2 += (float)s[2];
		num_consistent++;
	}
	const float div = (float)(num_consistent + 1);
	float* o = points + 6 * (size_t)offs[p];
	o[0] = X.x; o[1] = X.y; o[2] = X.z;
	o[3] = col0 / div; o[4] = col1 / div; o[5] = col2 / div;
}

// ---- the two Tanks-and-Temples variants (RunFusion_TAT_Intermediate APD.cpp:1962-2130, RunFusion_TAT_advanced
// APD.cpp:2132-2279; mode 1 / 2) ---------------------------------------------------------------------------------------
// An emitted pixel masks itself and masks are only looked at on source views, so inside one view no pixel depends on
// another through the masks.  The one sequential thing is the reference's `diff` vector, declared once per view
// (APD.cpp:2052, 2216): a source that is not evaluated for a pixel keeps the measures of the last pixel in raster order
// that did evaluate it.  That is a running "last evaluated pixel" per source: a max-scan.
//   k_fuse_tat_measures   per (pixel, source): measures + cell if evaluated; last = pixel index or -1
//   cub inclusive max-scan of `last` per source -> carry
//   k_fuse_tat_decide     per pixel: measures at carry, the k = 2..S test, own mask, used bits
//   k_fuse_tat_emit       points in raster order (colour: mode 1 mean with the counted sources' cells, mode 2 own colour)
struct MaxOp { __host__ __device__ __forceinline__ int operator()(int x, int y) const { return x > y ? x : y; } };

__device__ __forceinline__ bool tat_skipped(const ViewDev& rv, int p) {
	return (rv.block && rv.block[p] < 128) || rv.depth[p] <= 0.0f;     // APD.cpp:2055-2061
}

__global__ void __launch_bounds__(256) k_fuse_tat_measures(const __grid_constant__ RefArgs a, const ViewDev* __restrict__ views,
		int* __restrict__ cells, float* __restrict__ m_dist, float* __restrict__ m_depth, float* __restrict__ m_angle, int* __restrict__ last) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= a.N) return;
	const ViewDev& rv = views[a.ref];
	const int r = p / rv.w, c = p - r * rv.w;
	const bool skip = tat_skipped(rv, p);
	const float ref_depth = rv.depth[p];
	F3 X = {0.f, 0.f, 0.f};
	float n0 = 0.f, n1 = 0.f, n2 = 0.f;
	if (!skip) {
		X = point_on_world(c, r, ref_depth, rv.cam);
		n0 = rv.normal[3 * (size_t)p]; n1 = rv.normal[3 * (size_t)p + 1]; n2 = rv.normal[3 * (size_t)p + 2];
	}
	for (int j = 0; j < a.S; ++j) {
		int q = -1;
		float e = 0.f, d = 0.f, g = 0.f;
		if (!skip) {
			const ViewDev& sv = views[a.src[j]];
			q = measures(rv, sv, r, c, ref_depth, X, n0, n1, n2, e, d, g);
			if (q >= 0 && sv.mask[q] == 1) q = -1;                      // APD.cpp:2075-2076
		}
		const size_t at = (size_t)j * a.N + p;
		cells[at] = q; m_dist[at] = e; m_depth[at] = d; m_angle[at] = g;
		last[at] = q >= 0 ? p : -1;
	}
}

__global__ void __launch_bounds__(256) k_fuse_tat_decide(const __grid_constant__ RefArgs a, int mode, const ViewDev* __restrict__ views,
		const float* __restrict__ m_dist, const float* __restrict__ m_depth, const float* __restrict__ m_angle,
		const int* __restrict__ carry, unsigned* __restrict__ used, int* __restrict__ flags) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= a.N) return;
	const ViewDev& rv = views[a.ref];
	unsigned bits = 0;
	if (!tat_skipped(rv, p)) {
		const float dist_base = 0.25f;
		const float depth_base = mode == 1 ? 1.0f / 3500.0f : 1.0f / 3000.0f;
		const float angle_base = 0.06981317007977318f, angle_grad = 0.05235987755982988f;
		for (int k = 2; k <= a.S && !bits; ++k) {
			int count = 0;
			unsigned ok_bits = 0;
			const float lim_dist = (float)k * dist_base, lim_depth = (float)k * depth_base, lim_angle = (float)k * angle_grad + angle_base;
			for (int j = 0; j < a.S; ++j) {
				const int q = carry[(size_t)j * a.N + p];
				if (q < 0) continue;                                        // still FLT_MAX: fails every test
				const size_t at = (size_t)j * a.N + q;
				const bool ok = mode == 1 ? (m_dist[at] < lim_dist && m_depth[at] < lim_depth && m_angle[at] < lim_angle)
				                          : (m_dist[at] < lim_dist && m_depth[at] < lim_depth);
				if (ok) { count++; ok_bits |= 1u << j; }
			}
			if (count >= k) bits = ok_bits;
		}
		if (bits) rv.mask[p] = 1;                                           // APD.cpp:2121, 2270
	}
	used[p] = bits;
	flags[p] = bits != 0u;
}

__global__ void __launch_bounds__(256) k_fuse_tat_emit(const __grid_constant__ RefArgs a, int mode, const ViewDev* __restrict__ views,
		const int* __restrict__ cells, const int* __restrict__ carry, const unsigned* __restrict__ used, const int* __restrict__ offs, float* __restrict__ points) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= a.N) return;
	const unsigned bits = used[p];
	if (!bits) return;
	const ViewDev& rv = views[a.ref];
	const int r = p / rv.w, c = p - r * rv.w;
	const F3 X = point_on_world(c, r, rv.depth[p], rv.cam);
	float col0 = (float)rv.image[3 * (size_t)p], col1 = (float)rv.image[3 * (size_t)p + 1], col2 = (float)rv.image[3 * (size_t)p + 2];
	if (mode == 1) {                                                        // APD.cpp:2106-2116
		int count = 0;
		for (unsigned rest = bits; rest;) {
			const int j = __ffs(rest) - 1;
			rest &= rest - 1;
			const int q = carry[(size_t)j * a.N + p];
			const uint8_t* s = views[a.src[j]].image + 3 * (size_t)cells[(size_t)j * a.N + q];
			col0 += (float)s[0]; col1 += (float)s[1]; col2 += (float)s[2];
			count++;
		}
		const float div = (float)count + 1.0f;
		col0 = col0 / div; col1 = col1 / div; col2 = col2 / div;
	}
	float* o = points + 6 * (size_t)offs[p];
	o[0] = X.x; o[1] = X.y; o[2] = X.z; o[3] = col0; o[4] = col1; o[5] = col2;
}

// (world normal, depth) plane map -> the depth and normal maps ProcessProblem writes (main.cpp:300-306)
__global__ void __launch_bounds__(256) k_fuse_split_planes(int n, const float4* __restrict__ planes, float* __restrict__ depth, float* __restrict__ normal) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n) return;
	const float4 v = planes[p];
	depth[p] = v.w;
	normal[3 * (size_t)p] = v.x; normal[3 * (size_t)p + 1] = v.y; normal[3 * (size_t)p + 2] = v.z;
}

}  // namespace dvp_fuse

// =====================================================================================================================
using namespace dvp_fuse;

struct dvp_fusion {
	int device = 0;
	int V = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	struct HostView { bool set = false; int w = 0, h = 0, num_src = 0; int src[kMaxSrc]; ViewDev d; };
	std::vector<HostView> views;
	ViewDev* d_views = nullptr;
	// scratch of the view being fused, grown on demand
	size_t cap_n = 0, cap_ns = 0;
	int* cells = nullptr; float* terms = nullptr;
	unsigned* live = nullptr; unsigned* used = nullptr; uint8_t* state = nullptr;
	int* list[2] = {nullptr, nullptr};
	int* counters = nullptr;       // [0] list length, [1] tail rounds
	int* flags = nullptr; int* offs = nullptr;
	void* scan_temp = nullptr; size_t scan_temp_bytes = 0;
	float* d_points = nullptr;
	int mode = 0;                  // 0 RunFusion (ETH), 1 RunFusion_TAT_Intermediate, 2 RunFusion_TAT_advanced
	size_t cap_tat = 0;            // elements of the four [S][N] planes below (modes 1 and 2 only)
	float* m_depth = nullptr; float* m_angle = nullptr; int* last = nullptr; int* carry = nullptr;
	std::vector<float> points;     // 6 floats per point, the reference's order
	int last_view = -1, last_rounds = 0;
	int last_err = 0;
};

#define FCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { f->last_err = (int)e_; \
	fprintf(stderr, "[dvp_fusion] %s failed: %s (%s:%d)\n", #call, cudaGetErrorString(e_), __FILE__, __LINE__); return DVP_ERR_CUDA; } } while (0)

namespace {

void free_view(dvp_fusion::HostView& v) {
	cudaFree((void*)v.d.depth); cudaFree((void*)v.d.normal); cudaFree((void*)v.d.image); cudaFree((void*)v.d.weak);
	cudaFree((void*)v.d.block); cudaFree(v.d.mask); cudaFree(v.d.resv);
	memset(&v.d, 0, sizeof(v.d));
	v.set = false;
}

template <typename T> cudaError_t upload(const T** dst, const T* src, size_t count, cudaStream_t st) {
	T* d = nullptr;
	cudaError_t e = cudaMalloc((void**)&d, count * sizeof(T));
	if (e != cudaSuccess) return e;
	*dst = d;
	return cudaMemcpyAsync(d, src, count * sizeof(T), cudaMemcpyDefault, st);   // host or device source
}

int grow(dvp_fusion* f, size_t n, size_t s) {
	if (n > f->cap_n) {
		cudaFree(f->live); cudaFree(f->used); cudaFree(f->state); cudaFree(f->list[0]); cudaFree(f->list[1]);
		cudaFree(f->flags); cudaFree(f->offs); cudaFree(f->d_points); cudaFree(f->scan_temp);
		f->cap_n = 0;
		FCK(cudaMalloc((void**)&f->live, n * 4)); FCK(cudaMalloc((void**)&f->used, n * 4)); FCK(cudaMalloc((void**)&f->state, n));
		FCK(cudaMalloc((void**)&f->list[0], n * 4)); FCK(cudaMalloc((void**)&f->list[1], n * 4));
		FCK(cudaMalloc((void**)&f->flags, n * 4)); FCK(cudaMalloc((void**)&f->offs, n * 4));
		FCK(cudaMalloc((void**)&f->d_points, n * 6 * sizeof(float)));
		size_t sum_bytes = 0, max_bytes = 0;   // one temporary serves the exclusive sum and the max-scan of modes 1 / 2
		FCK(cub::DeviceScan::ExclusiveSum(nullptr, sum_bytes, (const int*)nullptr, (int*)nullptr, (int)n));
		FCK(cub::DeviceScan::InclusiveScan(nullptr, max_bytes, (const int*)nullptr, (int*)nullptr, MaxOp(), (int)n));
		f->scan_temp_bytes = (sum_bytes > max_bytes ? sum_bytes : max_bytes) + 1;
		FCK(cudaMalloc(&f->scan_temp, f->scan_temp_bytes));
		f->cap_n = n;
	}
	if (n * s > f->cap_ns) {
		cudaFree(f->cells); cudaFree(f->terms);
		f->cap_ns = 0;
		FCK(cudaMalloc((void**)&f->cells, n * s * 4)); FCK(cudaMalloc((void**)&f->terms, n * s * 4));
		f->cap_ns = n * s;
	}
	if (f->mode != 0 && n * s > f->cap_tat) {
		cudaFree(f->m_depth); cudaFree(f->m_angle); cudaFree(f->last); cudaFree(f->carry);
		f->cap_tat = 0;
		FCK(cudaMalloc((void**)&f->m_depth, n * s * 4)); FCK(cudaMalloc((void**)&f->m_angle, n * s * 4));
		FCK(cudaMalloc((void**)&f->last, n * s * 4)); FCK(cudaMalloc((void**)&f->carry, n * s * 4));
		f->cap_tat = n * s;
	}
	return DVP_OK;
}

}  // namespace

extern "C" {

dvp_fusion* dvp_fusion_create(int device, int num_views) {
	if (num_views <= 0) return nullptr;
	if (cudaSetDevice(device) != cudaSuccess) return nullptr;
	dvp_fusion* f = new dvp_fusion();
	f->device = device; f->V = num_views;
	f->views.resize(num_views);
	for (auto& v : f->views) memset(&v.d, 0, sizeof(v.d));
	bool ok = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking) == cudaSuccess
		&& cudaEventCreate(&f->ev0) == cudaSuccess && cudaEventCreate(&f->ev1) == cudaSuccess
		&& cudaMalloc((void**)&f->d_views, sizeof(ViewDev) * num_views) == cudaSuccess
		&& cudaMemset(f->d_views, 0, sizeof(ViewDev) * num_views) == cudaSuccess
		&& cudaMalloc((void**)&f->counters, 4 * sizeof(int)) == cudaSuccess;
	if (!ok) { dvp_fusion_destroy(f); return nullptr; }
	return f;
}

void dvp_fusion_destroy(dvp_fusion* f) {
	if (!f) return;
	cudaSetDevice(f->device);
	if (f->stream) cudaStreamSynchronize(f->stream);
	for (auto& v : f->views) free_view(v);
	cudaFree(f->d_views); cudaFree(f->cells); cudaFree(f->terms); cudaFree(f->live); cudaFree(f->used); cudaFree(f->state);
	cudaFree(f->list[0]); cudaFree(f->list[1]); cudaFree(f->counters); cudaFree(f->flags); cudaFree(f->offs);
	cudaFree(f->scan_temp); cudaFree(f->d_points);
	cudaFree(f->m_depth); cudaFree(f->m_angle); cudaFree(f->last); cudaFree(f->carry);
	if (f->ev0) cudaEventDestroy(f->ev0);
	if (f->ev1) cudaEventDestroy(f->ev1);
	if (f->stream) cudaStreamDestroy(f->stream);
	delete f;
}

// `planes` != NULL: depth and normal come from a [h][w][4] (world normal, depth) plane map instead of v->depth / v->normal
static int set_view_impl(dvp_fusion* f, int view, const dvp_fusion_view* v, const float* planes) {
	if (!f || !v || view < 0 || view >= f->V) return DVP_ERR_ARG;
	if (v->width <= 0 || v->height <= 0 || !v->image) return DVP_ERR_ARG;
	if (!planes && (!v->depth || !v->normal)) return DVP_ERR_ARG;
	if ((long long)v->width * v->height > 0x7fffffffLL / 8) return DVP_ERR_ARG;
	if (v->num_src < 0 || v->num_src > kMaxSrc || (v->num_src > 0 && !v->src_views)) return DVP_ERR_ARG;
	for (int j = 0; j < v->num_src; ++j)
		if (v->src_views[j] < 0 || v->src_views[j] >= f->V || v->src_views[j] == view) return DVP_ERR_ARG;
	FCK(cudaSetDevice(f->device));
	dvp_fusion::HostView& hv = f->views[view];
	FCK(cudaStreamSynchronize(f->stream));
	free_view(hv);
	const size_t n = (size_t)v->width * v->height;
	hv.w = v->width; hv.h = v->height; hv.num_src = v->num_src;
	for (int j = 0; j < v->num_src; ++j) hv.src[j] = v->src_views[j];
	hv.d.cam = v->camera; hv.d.w = v->width; hv.d.h = v->height;
	if (planes) {
		const float* d_planes = nullptr;
		float* d_depth = nullptr; float* d_normal = nullptr;
		FCK(upload(&d_planes, planes, 4 * n, f->stream));
		cudaError_t e = cudaMalloc((void**)&d_depth, n * sizeof(float));
		if (e == cudaSuccess) e = cudaMalloc((void**)&d_normal, 3 * n * sizeof(float));
		hv.d.depth = d_depth; hv.d.normal = d_normal;
		if (e == cudaSuccess) {
			k_fuse_split_planes<<<(unsigned)((n + 255) / 256), 256, 0, f->stream>>>((int)n, (const float4*)d_planes, d_depth, d_normal);
			e = cudaGetLastError();
		}
		if (e == cudaSuccess) e = cudaStreamSynchronize(f->stream);
		cudaFree((void*)d_planes);
		FCK(e);
	} else {
		FCK(upload(&hv.d.depth, v->depth, n, f->stream));
		FCK(upload(&hv.d.normal, v->normal, 3 * n, f->stream));
	}
	FCK(upload(&hv.d.image, v->image, 3 * n, f->stream));
	if (v->weak) FCK(upload(&hv.d.weak, v->weak, n, f->stream));
	if (v->block) FCK(upload(&hv.d.block, v->block, n, f->stream));
	FCK(cudaMalloc((void**)&hv.d.mask, n));
	FCK(cudaMalloc((void**)&hv.d.resv, n * sizeof(unsigned)));
	FCK(cudaMemsetAsync(hv.d.mask, 0, n, f->stream));
	FCK(cudaMemsetAsync(hv.d.resv, 0xFF, n * sizeof(unsigned), f->stream));
	FCK(cudaMemcpyAsync(f->d_views + view, &hv.d, sizeof(ViewDev), cudaMemcpyHostToDevice, f->stream));
	FCK(cudaStreamSynchronize(f->stream));   // the caller's buffers may go away after the call
	hv.set = true;
	return DVP_OK;
}

int dvp_fusion_set_view(dvp_fusion* f, int view, const dvp_fusion_view* v) { return set_view_impl(f, view, v, nullptr); }

int dvp_fusion_set_view_planes(dvp_fusion* f, int view, const dvp_fusion_view* v, const float* planes) {
	if (!planes) return DVP_ERR_ARG;
	return set_view_impl(f, view, v, planes);
}

int dvp_fusion_set_mode(dvp_fusion* f, int mode) {
	if (!f || mode < 0 || mode > 2) return DVP_ERR_ARG;
	f->mode = mode;
	return DVP_OK;
}

int dvp_fusion_reset(dvp_fusion* f) {
	if (!f) return DVP_ERR_ARG;
	FCK(cudaSetDevice(f->device));
	for (auto& v : f->views)
		if (v.set) {
			FCK(cudaMemsetAsync(v.d.mask, 0, (size_t)v.w * v.h, f->stream));
			FCK(cudaMemsetAsync(v.d.resv, 0xFF, (size_t)v.w * v.h * sizeof(unsigned), f->stream));
		}
	FCK(cudaStreamSynchronize(f->stream));
	f->points.clear();
	f->last_view = -1; f->last_rounds = 0;
	return DVP_OK;
}

// the end of a view: how many points the emit kernel wrote, copy them behind the ones already held
static int finish_view(dvp_fusion* f, int view, int n, int rounds) {
	cudaStream_t st = f->stream;
	int tail[2] = {0, 0};
	FCK(cudaMemcpyAsync(&tail[0], f->offs + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
	FCK(cudaMemcpyAsync(&tail[1], f->flags + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
	FCK(cudaEventRecord(f->ev1, st));
	FCK(cudaStreamSynchronize(st));
	const size_t emitted = (size_t)tail[0] + (size_t)tail[1];
	const size_t at = f->points.size();
	f->points.resize(at + 6 * emitted);
	if (emitted) FCK(cudaMemcpy(f->points.data() + at, f->d_points, 6 * emitted * sizeof(float), cudaMemcpyDeviceToHost));
	f->last_view = view; f->last_rounds = rounds;
	return DVP_OK;
}

int dvp_fusion_run_view(dvp_fusion* f, int view, float* device_ms) {
	if (!f || view < 0 || view >= f->V) return DVP_ERR_ARG;
	dvp_fusion::HostView& hv = f->views[view];
	if (!hv.set) return DVP_ERR_STATE;
	for (int j = 0; j < hv.num_src; ++j)
		if (!f->views[hv.src[j]].set) return DVP_ERR_STATE;
	FCK(cudaSetDevice(f->device));
	const int n = hv.w * hv.h, S = hv.num_src;
	{ int r = grow(f, (size_t)n, (size_t)(S > 0 ? S : 1)); if (r) return r; }
	RefArgs a;
	memset(&a, 0, sizeof(a));
	a.ref = view; a.S = S; a.N = n;
	for (int j = 0; j < S; ++j) a.src[j] = hv.src[j];
	cudaStream_t st = f->stream;
	const int blocks = (n + 255) / 256;
	FCK(cudaEventRecord(f->ev0, st));
	if (f->mode != 0) {
		k_fuse_tat_measures<<<blocks, 256, 0, st>>>(a, f->d_views, f->cells, f->terms, f->m_depth, f->m_angle, f->last);
		FCK(cudaGetLastError());
		for (int j = 0; j < S; ++j)
			FCK(cub::DeviceScan::InclusiveScan(f->scan_temp, f->scan_temp_bytes, (const int*)(f->last + (size_t)j * n), f->carry + (size_t)j * n, MaxOp(), n, st));
		k_fuse_tat_decide<<<blocks, 256, 0, st>>>(a, f->mode, f->d_views, f->terms, f->m_depth, f->m_angle, f->carry, f->used, f->flags);
		FCK(cub::DeviceScan::ExclusiveSum(f->scan_temp, f->scan_temp_bytes, (const int*)f->flags, f->offs, n, st));
		k_fuse_tat_emit<<<blocks, 256, 0, st>>>(a, f->mode, f->d_views, f->cells, f->carry, f->used, f->offs, f->d_points);
		FCK(cudaGetLastError());
		{ const int r = finish_view(f, view, n, 0); if (r) return r; }
		if (device_ms) FCK(cudaEventElapsedTime(device_ms, f->ev0, f->ev1));
		return DVP_OK;
	}
	if (!hv.d.weak) return DVP_ERR_STATE;   // RunFusion reads weak.bin (APD.cpp:1933)
	FCK(cudaMemsetAsync(f->counters, 0, 4 * sizeof(int), st));
	k_fuse_candidates<<<blocks, 256, 0, st>>>(a, f->d_views, f->cells, f->terms, f->live, f->used, f->state, f->list[0], f->counters);
	FCK(cudaGetLastError());
	int count = 0;
	FCK(cudaMemcpyAsync(&count, f->counters, sizeof(int), cudaMemcpyDeviceToHost, st));
	FCK(cudaStreamSynchronize(st));
	int rounds = 0, cur = 0;
	const int kRoundsPerSync = 4, kTail = 4 * kTailThreads;
	while (count > 0) {
		if (count <= kTail) {
			k_fuse_tail<<<1, kTailThreads, 0, st>>>(a, f->d_views, f->cells, f->terms, f->live, f->used, f->state, f->list[cur], count, f->counters + 1);
			FCK(cudaGetLastError());
			int tail_rounds = 0;
			FCK(cudaMemcpyAsync(&tail_rounds, f->counters + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
			FCK(cudaStreamSynchronize(st));
			if (tail_rounds < 0) return DVP_ERR_STATE;
			rounds += tail_rounds;
			break;
		}
		const int g = (count + 255) / 256;
		for (int k = 0; k < kRoundsPerSync; ++k) {
			k_fuse_reserve<<<g, 256, 0, st>>>(a, f->d_views, f->cells, f->live, f->state, f->list[cur], count);
			k_fuse_decide<<<g, 256, 0, st>>>(a, f->d_views, f->cells, f->terms, f->live, f->used, f->state, f->list[cur], count);
		}
		rounds += kRoundsPerSync;
		FCK(cudaMemsetAsync(f->counters, 0, sizeof(int), st));
		k_fuse_compact<<<g, 256, 0, st>>>(f->state, f->list[cur], count, f->list[cur ^ 1], f->counters);
		FCK(cudaGetLastError());
		const int before = count;
		FCK(cudaMemcpyAsync(&count, f->counters, sizeof(int), cudaMemcpyDeviceToHost, st));
		FCK(cudaStreamSynchronize(st));
		if (count >= before) return DVP_ERR_STATE;   // cannot happen: every round decides its smallest pixel
		cur ^= 1;
	}
	k_fuse_flags<<<blocks, 256, 0, st>>>(n, f->used, f->flags);
	FCK(cub::DeviceScan::ExclusiveSum(f->scan_temp, f->scan_temp_bytes, (const int*)f->flags, f->offs, n, st));
	k_fuse_emit<<<blocks, 256, 0, st>>>(a, f->d_views, f->cells, f->used, f->offs, f->d_points);
	FCK(cudaGetLastError());
	{ const int r = finish_view(f, view, n, rounds); if (r) return r; }
	if (device_ms) FCK(cudaEventElapsedTime(device_ms, f->ev0, f->ev1));
	return DVP_OK;
}

int dvp_fusion_run(dvp_fusion* f, long long* num_points, float* device_ms) {
	if (!f) return DVP_ERR_ARG;
	int r = dvp_fusion_reset(f);
	if (r) return r;
	float total = 0.0f;
	for (int v = 0; v < f->V; ++v) {
		float ms = 0.0f;
		r = dvp_fusion_run_view(f, v, &ms);
		if (r) return r;
		total += ms;
	}
	if (num_points) *num_points = (long long)(f->points.size() / 6);
	if (device_ms) *device_ms = total;
	return DVP_OK;
}

long long dvp_fusion_num_points(dvp_fusion* f) { return f ? (long long)(f->points.size() / 6) : -1; }

int dvp_fusion_get_points(dvp_fusion* f, float* dst, long long first, long long count) {
	if (!f || !dst || first < 0 || count < 0 || (size_t)(first + count) * 6 > f->points.size()) return DVP_ERR_ARG;
	memcpy(dst, f->points.data() + 6 * first, (size_t)count * 6 * sizeof(float));
	return DVP_OK;
}

int dvp_fusion_get_mask(dvp_fusion* f, int view, uint8_t* dst) {
	if (!f || !dst || view < 0 || view >= f->V) return DVP_ERR_ARG;
	const dvp_fusion::HostView& hv = f->views[view];
	if (!hv.set) return DVP_ERR_STATE;
	FCK(cudaSetDevice(f->device));
	FCK(cudaStreamSynchronize(f->stream));
	FCK(cudaMemcpy(dst, hv.d.mask, (size_t)hv.w * hv.h, cudaMemcpyDeviceToHost));
	return DVP_OK;
}

int dvp_fusion_last_view_index(dvp_fusion* f) { return (f && f->last_view >= 0) ? f->last_view : DVP_ERR_STATE; }

int dvp_fusion_last_view(dvp_fusion* f, int32_t* cells, float* terms, uint32_t* used, int* rounds) {
	if (!f) return DVP_ERR_ARG;
	if (f->last_view < 0) return DVP_ERR_STATE;
	if (f->mode != 0 && (cells || terms)) return DVP_ERR_STATE;   // candidates are a stage of RunFusion (mode 0) only
	const dvp_fusion::HostView& hv = f->views[f->last_view];
	const size_t n = (size_t)hv.w * hv.h, S = (size_t)hv.num_src;
	FCK(cudaSetDevice(f->device));
	FCK(cudaStreamSynchronize(f->stream));
	if (cells || terms) {   // device layout is [S][N]; the exchange layout is [N][S]
		std::vector<int32_t> tmp(n * (S ? S : 1));
		if (cells) {
			FCK(cudaMemcpy(tmp.data(), f->cells, n * S * 4, cudaMemcpyDeviceToHost));
			for (size_t j = 0; j < S; ++j) for (size_t p = 0; p < n; ++p) cells[p * S + j] = tmp[j * n + p];
		}
		if (terms) {
			FCK(cudaMemcpy(tmp.data(), f->terms, n * S * 4, cudaMemcpyDeviceToHost));
			const float* t = (const float*)tmp.data();
			for (size_t j = 0; j < S; ++j) for (size_t p = 0; p < n; ++p) terms[p * S + j] = t[j * n + p];
		}
	}
	if (used) FCK(cudaMemcpy(used, f->used, n * 4, cudaMemcpyDeviceToHost));
	if (rounds) *rounds = f->last_rounds;
	return DVP_OK;
}

int dvp_fusion_write_ply(dvp_fusion* f, const char* path) {
	if (!f || !path) return DVP_ERR_ARG;
	FILE* out = fopen(path, "wb");
	if (!out) return DVP_ERR_ARG;
	const long long n = (long long)(f->points.size() / 6);
	fprintf(out, "ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
		"property uchar diffuse_blue\nproperty uchar diffuse_green\nproperty uchar diffuse_red\nend_header\n", (int)n);
	std::vector<unsigned char> rec(15 * 4096);
	for (long long i = 0; i < n;) {
		const long long m = (n - i < 4096) ? n - i : 4096;
		for (long long k = 0; k < m; ++k) {
			const float* p = f->points.data() + 6 * (i + k);
			memcpy(&rec[15 * k], p, 12);
			for (int c = 0; c < 3; ++c) rec[15 * k + 12 + c] = (unsigned char)p[3 + c];   // static_cast<uchar>, APD.cpp:868-871
		}
		fwrite(rec.data(), 15, (size_t)m, out);
		i += m;
	}
	fclose(out);
	return DVP_OK;
}

}  // extern "C"
