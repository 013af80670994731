// dvp_kernels_post.cu — what ProcessProblem does to the maps right after RunPatchMatch, on the device
// (reference main.cpp:297-363 with Connect / Label_Seek / Label_Update, APD.cpp:138-346):
//   (1) depths outside [depth_min, depth_max] become 0 and their pixel state UNKNOWN (main.cpp:300-306);
//   (2) per source view, the pixels that do NOT select the view are grouped into 4-connected regions and every
//       region smaller than 20 * (8 / scale)^2 pixels gets the view's bit set (main.cpp:323-363).
// The reference labels regions with a sequential two-pass scan plus an O(L^2) label-merge list on the host
// (one D2H + H2D round trip of every map per view per pass).  Here: a lock-free union-find over the pixel grid
// (atomicMin on parent links), one pass to flatten, a warp-aggregated size count and one pass to apply — the
// maps never leave HBM.  Region sizes, and therefore the result, do not depend on label numbering.
#include "dvp_common.cuh"
#include "dvp_launch.h"
#include "dvp_unionfind.cuh"
#include <cstring>

namespace dvp {

__global__ void __launch_bounds__(256) k_vis_init(const uint32_t* __restrict__ selected, int view, int n, int* __restrict__ parent, int* __restrict__ count) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n) return;
	parent[p] = ((selected[p] >> view) & 1u) ? -1 : p;
	count[p] = 0;
}

// one thread per pixel: join with the left and upper neighbour (the two links Connect looks at, APD.cpp:250-275)
__global__ void __launch_bounds__(256) k_vis_merge(int W, int H, int* parent) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= W || y >= H) return;
	const int p = y * W + x;
	if (parent[p] < 0) return;
	if (x > 0 && parent[p - 1] >= 0) uf_union(parent, p, p - 1);
	if (y > 0 && parent[p - W] >= 0) uf_union(parent, p, p - W);
}

// flatten (parent[p] = root) and count pixels per root; lanes of a warp that share a root add once
__global__ void __launch_bounds__(256) k_vis_count(int n, int* parent, int* count) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	int root = -1;
	if (p < n && parent[p] >= 0) {
		root = uf_find(parent, p);
		parent[p] = root;   // a concurrent walker sees the old link or the root; both are ancestors
	}
	const unsigned peers = __match_any_sync(0xffffffffu, root);
	if (root >= 0 && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&count[root], __popc(peers));
}

__global__ void __launch_bounds__(256) k_vis_apply(int n, int view, int threshold, const int* __restrict__ parent, const int* __restrict__ count, uint32_t* selected) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n) return;
	const int root = parent[p];
	if (root >= 0 && count[root] < threshold) selected[p] |= 1u << view;
}

// main.cpp:300-306: depth out of range -> depth 0, state UNKNOWN
__global__ void __launch_bounds__(256) k_invalidate_depth(int n, float depth_min, float depth_max, float4* planes, uint8_t* weak) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n) return;
	const float d = planes[p].w;
	if (d < depth_min || d > depth_max) { planes[p].w = 0.0f; weak[p] = DVP_UNKNOWN; }
}

// RescaleMatToTargetSize (APD.cpp:1773-1796): nearest-neighbour resampling with the two scale factors swapped
// (row index divided by scale_x, column index by scale_y — SURVEY B10, reproduced).  Host float arithmetic is
// IEEE, so the divisions are __fdiv_rn here (the library is built with --use_fast_math).  Target pixels whose
// source falls outside are left uninitialised by the reference; they are defined as zero here.
template <typename T>
__global__ void __launch_bounds__(256) k_rescale(const T* __restrict__ src, int sw, int sh, T* __restrict__ dst, int dw, int dh) {
	const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y * blockDim.y + threadIdx.y;
	if (c >= dw || r >= dh) return;
	const float scale_x = __fdiv_rn((float)dw, (float)sw);
	const float scale_y = __fdiv_rn((float)dh, (float)sh);
	const int o_r = (int)__fdiv_rn((float)r, scale_x);
	const int o_c = (int)__fdiv_rn((float)c, scale_y);
	T v; memset(&v, 0, sizeof(T));
	if (o_r >= 0 && o_c >= 0 && o_r < sh && o_c < sw) v = src[(size_t)o_r * sw + o_c];
	dst[(size_t)r * dw + c] = v;
}

cudaError_t launch_rescale(const void* src, int sw, int sh, void* dst, int dw, int dh, int elem_bytes, cudaStream_t st) {
	const dim3 b(32, 8), g((dw + 31) / 32, (dh + 7) / 8);
	switch (elem_bytes) {
	case 1: k_rescale<uint8_t><<<g, b, 0, st>>>((const uint8_t*)src, sw, sh, (uint8_t*)dst, dw, dh); break;
	case 4: k_rescale<uint32_t><<<g, b, 0, st>>>((const uint32_t*)src, sw, sh, (uint32_t*)dst, dw, dh); break;
	case 16: k_rescale<uint4><<<g, b, 0, st>>>((const uint4*)src, sw, sh, (uint4*)dst, dw, dh); break;
	default: return cudaErrorInvalidValue;
	}
	return cudaGetLastError();
}

// depth map of a finished pass: the w component of the plane map (what ProcessProblem writes to depths.dmb)
__global__ void __launch_bounds__(256) k_extract_depth(int n, const float4* __restrict__ planes, float* __restrict__ depth) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p < n) depth[p] = planes[p].w;
}
cudaError_t launch_extract_depth(const float4* planes, float* depth, int n, cudaStream_t st) {
	k_extract_depth<<<(n + 255) / 256, 256, 0, st>>>(n, planes, depth);
	return cudaGetLastError();
}

cudaError_t launch_invalidate_depth(const KArgs& a, cudaStream_t st) {
	k_invalidate_depth<<<(a.N + 255) / 256, 256, 0, st>>>(a.N, a.prm.depth_min, a.prm.depth_max, a.planes, a.weak);
	return cudaGetLastError();
}

cudaError_t launch_restore_visibility(const KArgs& a, int scale_size, int* parent, int* count, cudaStream_t st) {
	const int k = scale_size > 0 ? 8 / scale_size : 0;
	const int threshold = 20 * k * k;
	const int blocks = (a.N + 255) / 256;
	const dim3 b2(32, 8), g2((a.W + 31) / 32, (a.H + 7) / 8);
	for (int v = 0; v < a.S; ++v) {
		k_vis_init<<<blocks, 256, 0, st>>>(a.selected, v, a.N, parent, count);
		k_vis_merge<<<g2, b2, 0, st>>>(a.W, a.H, parent);
		k_vis_count<<<blocks, 256, 0, st>>>(a.N, parent, count);
		k_vis_apply<<<blocks, 256, 0, st>>>(a.N, v, threshold, parent, count, a.selected);
	}
	return cudaGetLastError();
}

}  // namespace dvp
