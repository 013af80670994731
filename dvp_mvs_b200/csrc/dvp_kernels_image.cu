// dvp_kernels_image.cu — image preparation on the device:
//   * the float image pyramid InuputInitialization and GetProblemEdges build with cv::resize(.., INTER_LINEAR) from the
//     full-resolution grey image (reference APD.cpp:1119-1140, main.cpp:203-209);
//   * (label half of row N4, further down) EdgeSegment(scale, image, mode 1): APD.cpp:348-402, 437-499.
// cv::resize is third-party code (OpenCV, not under the reference tree; the reference asks for "OpenCV >= 3.3").  What is
// reproduced here is its GENERIC bilinear path for CV_32F (imgproc/src/resize.cpp: resizeGeneric_ with HResizeLinear /
// VResizeLinear): per destination column fx = (float)((dx + 0.5) * scale_x - 0.5) with scale_x = 1 / (dst_w / src_w) in
// double, sx = floor(fx), fx -= sx, taps clamped at the borders; a row pass D = S[sx] * (1 - fx) + S[sx + 1] * fx, then a
// column pass of the same shape, every product and sum rounded to float on its own (the baseline build has no FMA).
// That is bit-exact with OpenCV 4.13 when its IPP back end is off (tests/golden/resize_f32.npz, tools/make_image_golden.py);
// an IPP-enabled build of OpenCV answers up to 0.015 grey levels differently — the tolerance any caller comparing against
// "whatever cv::resize gave" has to allow.
#include "dvp_common.cuh"
#include "dvp_launch.h"
#include "dvp_unionfind.cuh"
#include <algorithm>

namespace dvp {

namespace {

__device__ __forceinline__ void linear_tap(int d, double scale, int n_src, bool zero_at_borders, int& s0, int& s1, float& f) {
	const float v = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
	int s = (int)floorf(v);
	f = __fsub_rn(v, (float)s);
	if (zero_at_borders) {
		if (s < 0) { f = 0.f; s = 0; }
		if (s >= n_src - 1) { f = 0.f; s = n_src - 1; }
	}
	s0 = min(max(s, 0), n_src - 1);
	s1 = min(max(s + 1, 0), n_src - 1);
}

__global__ void __launch_bounds__(256) k_resize_linear_f32(const float* __restrict__ src, int sw, int sh, float* __restrict__ dst, int dw, int dh,
                                                            double scale_x, double scale_y) {
	const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y * blockDim.y + threadIdx.y;
	if (dx >= dw || dy >= dh) return;
	int x0, x1, y0, y1; float fx, fy;
	linear_tap(dx, scale_x, sw, true, x0, x1, fx);     // resize.cpp zeroes fx where the second tap leaves the row
	linear_tap(dy, scale_y, sh, false, y0, y1, fy);    // rows are clipped instead (VResize gets the same row twice)
	const float a0 = __fsub_rn(1.0f, fx), b0 = __fsub_rn(1.0f, fy);
	const float* r0 = src + (size_t)y0 * sw; const float* r1 = src + (size_t)y1 * sw;
	const float h0 = __fadd_rn(__fmul_rn(r0[x0], a0), __fmul_rn(r0[x1], fx));
	const float h1 = __fadd_rn(__fmul_rn(r1[x0], a0), __fmul_rn(r1[x1], fx));
	dst[(size_t)dy * dw + dx] = __fadd_rn(__fmul_rn(h0, b0), __fmul_rn(h1, fy));
}

__global__ void __launch_bounds__(256) k_u8_to_f32(const uint8_t* __restrict__ src, size_t n, float* __restrict__ dst) {
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = (float)src[i];
}

}  // namespace

cudaError_t launch_resize_linear_f32(const float* src, int sw, int sh, float* dst, int dw, int dh, cudaStream_t st) {
	const double scale_x = 1.0 / ((double)dw / (double)sw), scale_y = 1.0 / ((double)dh / (double)sh);
	dim3 b(32, 8), g((dw + 31) / 32, (dh + 7) / 8);
	k_resize_linear_f32<<<g, b, 0, st>>>(src, sw, sh, dst, dw, dh, scale_x, scale_y);
	return cudaGetLastError();
}
cudaError_t launch_u8_to_f32(const uint8_t* src, size_t n, float* dst, cudaStream_t st) {
	k_u8_to_f32<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 16), 256, 0, st>>>(src, n, dst);
	return cudaGetLastError();
}

}  // namespace dvp

using namespace dvp;

extern "C" {

int dvp_resize_linear_f32(int device, const float* src, int src_w, int src_h, float* dst, int dst_w, int dst_h) {
	if (!src || !dst || src_w <= 0 || src_h <= 0 || dst_w <= 0 || dst_h <= 0) return DVP_ERR_ARG;
	if (cudaSetDevice(device) != cudaSuccess) return DVP_ERR_CUDA;
	const size_t ns = (size_t)src_w * src_h, nd = (size_t)dst_w * dst_h;
	float *d_src = nullptr, *d_dst = nullptr;
	cudaError_t e = cudaMalloc((void**)&d_src, ns * 4);
	if (e == cudaSuccess) e = cudaMalloc((void**)&d_dst, nd * 4);
	if (e == cudaSuccess) e = cudaMemcpy(d_src, src, ns * 4, cudaMemcpyDefault);
	if (e == cudaSuccess) e = launch_resize_linear_f32(d_src, src_w, src_h, d_dst, dst_w, dst_h, 0);
	if (e == cudaSuccess) e = cudaMemcpy(dst, d_dst, nd * 4, cudaMemcpyDefault);
	cudaFree(d_src); cudaFree(d_dst);
	return e == cudaSuccess ? DVP_OK : DVP_ERR_CUDA;
}

}  // extern "C"

// =====================================================================================================================
// Row N4, label half: EdgeSegment(scale, image, mode 1, use_canny = false) — reference APD.cpp:348-402, 437-499 — as
// GetProblemEdges calls it on the full-resolution 8-bit image (main.cpp:229-241).  Its result is `label_cuda`, which K2 and
// K4 only ever test for `> 0`, `== 0`, `== -1` and equality with the centre's label (APD.cu:3461, 3629, 3857-3886): the
// device path reproduces the reference's PARTITION and the three classes, not its label numbering.
//   two 8-bit cv::resize halvings -> Roberts (APD.cpp:120-136) -> threshold > 4 -> connected regions (Connect / Label_Update,
//   APD.cpp:195-346) -> for every region of >= weak_tex_num pixels: its 4-neighbour border -> cv::HoughLinesP -> cv::line
//   onto the edge image -> 8-bit cv::resize to the level size -> threshold -> border clean-up (APD.cpp:452-463) ->
//   connected regions -> regions of <= weak_tex_num pixels become -1.
// cv::resize / cv::HoughLinesP / cv::line are OpenCV's (third party); restated from the published algorithms exactly as
// oracle/cpu/label_cpu.cpp and hough_cpu.cpp restate them (pinned against OpenCV 4.13 golden label maps).
// Everything is data-parallel except HoughLinesP, whose points vote in the pseudo-random order of cv::RNG((uint64)-1) and
// remove each other's votes: one CTA per region follows that order, its threads owning one of the 180 angles each
// (votes, maximum, vote removal) while its first warp walks the candidate line 32 pixels per step; regions run concurrently.
namespace dvp {
namespace {

// cv::resize(src, dst, Size(dw, dh), 0, 0, INTER_LINEAR) for CV_8UC1 (imgproc/src/resize.cpp): fixed-point, 11-bit
// coefficients; when both ratios are exactly 2 OpenCV takes the 2x2 INTER_AREA path instead
__global__ void __launch_bounds__(256) k_resize8u(const uint8_t* __restrict__ src, int sw, int sh, uint8_t* __restrict__ dst, int dw, int dh,
                                                   double scale_x, double scale_y, int area2x2) {
	const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y * blockDim.y + threadIdx.y;
	if (dx >= dw || dy >= dh) return;
	if (area2x2) {
		const uint8_t* p = src + (size_t)(2 * dy) * sw + 2 * dx;
		dst[(size_t)dy * dw + dx] = (uint8_t)((p[0] + p[1] + p[sw] + p[sw + 1] + 2) >> 2);
		return;
	}
	float fx = (float)__dadd_rn(__dmul_rn((double)dx + 0.5, scale_x), -0.5);
	int sx = (int)floorf(fx);
	fx = __fsub_rn(fx, (float)sx);
	if (sx < 0) { fx = 0.f; sx = 0; }
	if (sx >= sw - 1) { fx = 0.f; sx = sw - 1; }
	const int a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, fx), 2048.f)), a1 = __float2int_rn(__fmul_rn(fx, 2048.f));
	float fy = (float)__dadd_rn(__dmul_rn((double)dy + 0.5, scale_y), -0.5);
	const int sy = (int)floorf(fy);
	fy = __fsub_rn(fy, (float)sy);
	const int b0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, fy), 2048.f)), b1 = __float2int_rn(__fmul_rn(fy, 2048.f));
	const int sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
	const int y0 = min(max(sy, 0), sh - 1), y1 = min(max(sy + 1, 0), sh - 1);
	const uint8_t* r0 = src + (size_t)y0 * sw; const uint8_t* r1 = src + (size_t)y1 * sw;
	const int h0 = r0[sx] * a0 + r0[sx1] * a1, h1 = r1[sx] * a0 + r1[sx1] * a1;
	dst[(size_t)dy * dw + dx] = (uint8_t)((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
}

// Roberts (APD.cpp:120-136: (uchar)sqrt(t1^2 + t2^2), 50 / 50 on the image border) followed by cv::threshold(> 4 -> 255)
__global__ void __launch_bounds__(256) k_roberts_threshold(const uint8_t* __restrict__ src, int w, int h, uint8_t* __restrict__ dst) {
	const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
	if (j >= w || i >= h) return;
	int t1 = 50, t2 = 50;
	if (i > 0 && i < h - 1 && j > 0 && j < w - 1) {
		t1 = (int)src[(size_t)i * w + j] - (int)src[(size_t)(i + 1) * w + j + 1];
		t2 = (int)src[(size_t)(i + 1) * w + j] - (int)src[(size_t)i * w + j + 1];
	}
	const uint8_t v = (uint8_t)(int)sqrt((double)(t1 * t1 + t2 * t2));
	dst[(size_t)i * w + j] = v > 4 ? 255 : 0;
}
__global__ void __launch_bounds__(256) k_threshold4(uint8_t* img, size_t n) {
	const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p < n) img[p] = img[p] > 4 ? 255 : 0;
}

// regions = 4-connected components of the pixels that are not 255 (Connect joins two pixels when both are 0; after the
// threshold every non-255 pixel is 0)
__global__ void __launch_bounds__(256) k_lab_init(const uint8_t* __restrict__ img, int n, int* __restrict__ parent, int* __restrict__ count) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n) return;
	parent[p] = img[p] == 255 ? -1 : p;
	count[p] = 0;
}
__global__ void __launch_bounds__(256) k_lab_link(int W, int H, int* parent) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= W || y >= H) return;
	const int p = y * W + x;
	if (parent[p] < 0) return;
	if (x > 0 && parent[p - 1] >= 0) uf_union(parent, p, p - 1);
	if (y > 0 && parent[p - W] >= 0) uf_union(parent, p, p - W);
}
// flatten (parent[p] = root) and count region sizes at the roots
__global__ void __launch_bounds__(256) k_lab_count(int n, int* __restrict__ parent, int* __restrict__ count) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n || parent[p] < 0) return;
	const int r = uf_find(parent, p);
	parent[p] = r;
	atomicAdd(&count[r], 1);
}
// regions of at least weak_tex_num pixels get a slot (order irrelevant: regions are processed independently)
__global__ void __launch_bounds__(256) k_lab_big_regions(int n, const int* __restrict__ parent, const int* __restrict__ count, int weak_tex_num,
                                                          int* __restrict__ region_of_root, int* __restrict__ num_regions) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n) return;
	region_of_root[p] = -1;
	if (parent[p] == p && count[p] >= weak_tex_num) region_of_root[p] = atomicAdd(num_regions, 1);
}
// border pixels of every big region (APD.cpp:378-393): pixels outside the region with a 4-neighbour inside it.  One
// 64-bit key (region << 32 | pixel) per (region, border pixel); sorting the keys yields, per region, its border in raster
// order — the order in which HoughLinesP collects its points.
__global__ void __launch_bounds__(256) k_lab_border_keys(int W, int H, const int* __restrict__ parent, const int* __restrict__ region_of_root,
                                                          unsigned long long* __restrict__ keys, int* __restrict__ num_keys, int cap) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= W || y >= H) return;
	const int p = y * W + x;
	const int own = parent[p] >= 0 ? region_of_root[parent[p]] : -1;
	int seen[4]; int ns = 0;
	auto visit = [&](int q) {
		if (parent[q] < 0) return;
		const int r = region_of_root[parent[q]];
		if (r < 0 || r == own) return;
		for (int k = 0; k < ns; ++k) if (seen[k] == r) return;
		seen[ns++] = r;
		const int slot = atomicAdd(num_keys, 1);
		if (slot < cap) keys[slot] = ((unsigned long long)r << 32) | (unsigned)p;
	};
	if (x > 0) visit(p - 1);
	if (x < W - 1) visit(p + 1);
	if (y > 0) visit(p - W);
	if (y < H - 1) visit(p + W);
}
__global__ void __launch_bounds__(256) k_lab_region_starts(const unsigned long long* __restrict__ keys, int m, int* __restrict__ start, int num_regions) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	const int r = (int)(keys[i] >> 32);
	if (i == 0 || (int)(keys[i - 1] >> 32) != r) start[r] = i;
	if (i == m - 1) start[num_regions] = m;
}

// cv::RNG (multiply-with-carry), core/operations.hpp
struct CvRng {
	unsigned long long state;
	__device__ unsigned next() { state = (unsigned long long)(unsigned)state * 4164903690ull + (unsigned)(state >> 32); return (unsigned)state; }
	__device__ int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

constexpr int kHoughThreads = 192;     // >= numangle (180 for theta = pi / 180)
constexpr int kHoughMaxLinePixels = 4096;

// cv::HoughLinesP(img_weak, lines, 1, CV_PI / 180, threshold, min_line_length, max_line_gap) (imgproc/src/hough.cpp,
// HoughLinesProbabilistic) for the border image of one region per CTA; regions are dealt to the CTAs round robin.
// slab per CTA: accum [numangle][numrho] ints, mask [h][w] bytes (both all-zero between regions).
__global__ void __launch_bounds__(kHoughThreads) k_hough_regions(int W, int H, int num_regions, const int* __restrict__ start, unsigned long long* __restrict__ keys,
                                                                   int numangle, int numrho, const float* __restrict__ trigtab, int threshold, int line_length, int line_gap,
                                                                   int* __restrict__ accum_slabs, uint8_t* __restrict__ mask_slabs, int4* __restrict__ lines, int* __restrict__ num_lines, int max_lines) {
	__shared__ int s_red[kHoughThreads / 32];
	__shared__ int s_ctl[8];                       // broadcast scratch: [0] go / skip, [1] x, [2] y, [3] max key, [4] number of listed pixels, [5] good line
	__shared__ int s_px[kHoughMaxLinePixels];      // pixels of the current line whose votes have to be removed
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	int* accum = accum_slabs + (size_t)blockIdx.x * numangle * numrho;
	uint8_t* mask = mask_slabs + (size_t)blockIdx.x * W * H;
	const float t_cos = tid < numangle ? trigtab[2 * tid] : 0.f, t_sin = tid < numangle ? trigtab[2 * tid + 1] : 0.f;
	const int half_rho = (numrho - 1) / 2;
	const int shift = 16;
	for (int region = blockIdx.x; region < num_regions; region += gridDim.x) {
		const int first = start[region], npts = start[region + 1] - first;
		unsigned* loc = reinterpret_cast<unsigned*>(keys + first);   // low words of the 64-bit keys: loc[2 * i] = pixel index (little endian)
		for (int i = tid; i < npts; i += kHoughThreads) mask[loc[2 * i]] = 1;
		__syncthreads();
		CvRng rng; rng.state = 0xffffffffffffffffull;
		for (int count = npts; count > 0; count--) {
			if (tid == 0) {
				const int idx = rng.uniform(0, count);
				const unsigned pt = loc[2 * idx];
				loc[2 * idx] = loc[2 * (count - 1)];
				s_ctl[0] = mask[pt] ? 1 : 0;
				s_ctl[1] = (int)(pt % (unsigned)W); s_ctl[2] = (int)(pt / (unsigned)W);
			}
			__syncthreads();
			const int go = s_ctl[0], j = s_ctl[1], i = s_ctl[2];
			if (!go) { __syncthreads(); continue; }
			// votes: thread n owns angle n
			int key = 0;
			if (tid < numangle) {
				const int r = __float2int_rn(__fadd_rn(__fmul_rn((float)j, t_cos), __fmul_rn((float)i, t_sin))) + half_rho;
				const int val = ++accum[(size_t)tid * numrho + r];
				key = (val << 8) | (255 - tid);        // maximum vote count, smallest angle among equals (`if (max_val < val)`)
			}
			for (int o = 16; o > 0; o >>= 1) key = max(key, __shfl_xor_sync(0xffffffffu, key, o));
			if (lane == 0) s_red[warp] = key;
			__syncthreads();
			if (tid == 0) {
				int k = s_red[0];
				for (int w = 1; w < kHoughThreads / 32; ++w) k = max(k, s_red[w]);
				s_ctl[3] = k;
			}
			__syncthreads();
			const int max_val = s_ctl[3] >> 8, max_n = 255 - (s_ctl[3] & 255);
			if (max_val < threshold) { __syncthreads(); continue; }
			// the candidate line through (j, i) at angle max_n, in 16.16 fixed point
			const float a = -trigtab[2 * max_n + 1], b = trigtab[2 * max_n];
			int x0 = j, y0 = i, dx0, dy0, xflag;
			if (fabsf(a) > fabsf(b)) {
				xflag = 1; dx0 = a > 0 ? 1 : -1;
				dy0 = __float2int_rn(__fdiv_rn(__fmul_rn(b, 65536.f), fabsf(a)));
				y0 = (y0 << shift) + (1 << (shift - 1));
			} else {
				xflag = 0; dy0 = b > 0 ? 1 : -1;
				dx0 = __float2int_rn(__fdiv_rn(__fmul_rn(a, 65536.f), fabsf(b)));
				x0 = (x0 << shift) + (1 << (shift - 1));
			}
			// first warp: walk both directions 32 pixels per step; every lane holds the same ballots, so the gap logic runs
			// redundantly (and identically) in all lanes
			int end_t[2] = {0, 0};
			if (warp == 0) {
				for (int k = 0; k < 2; ++k) {
					const int dx = k ? -dx0 : dx0, dy = k ? -dy0 : dy0;
					int gap = 0, last = 0; bool done = false;
					for (int base = 0; !done; base += 32) {
						const int t = base + lane;
						const int x = x0 + t * dx, y = y0 + t * dy;
						const int j1 = xflag ? x : x >> shift, i1 = xflag ? y >> shift : y;
						const bool inb = j1 >= 0 && j1 < W && i1 >= 0 && i1 < H;
						const unsigned oob = __ballot_sync(0xffffffffu, !inb);
						const unsigned set = __ballot_sync(0xffffffffu, inb && mask[(size_t)i1 * W + j1]);
						for (int bit = 0; bit < 32; ++bit) {
							if ((oob >> bit) & 1) { done = true; break; }
							if ((set >> bit) & 1) { gap = 0; last = base + bit; }
							else if (++gap > line_gap) { done = true; break; }
						}
					}
					end_t[k] = last;
				}
			}
			// line ends, good-line test (hough.cpp: either extent >= lineLength)
			int ex[2], ey[2];
			for (int k = 0; k < 2; ++k) {
				const int dx = k ? -dx0 : dx0, dy = k ? -dy0 : dy0;
				const int x = x0 + end_t[k] * dx, y = y0 + end_t[k] * dy;
				ex[k] = xflag ? x : x >> shift; ey[k] = xflag ? y >> shift : y;
			}
			if (warp == 0) {
				const bool good = abs(ex[1] - ex[0]) >= line_length || abs(ey[1] - ey[0]) >= line_length;
				// second walk: every set pixel between the start and each end is cleared; on a good line its votes go too
				int listed = 0;
				for (int k = 0; k < 2; ++k) {
					const int dx = k ? -dx0 : dx0, dy = k ? -dy0 : dy0;
					for (int base = 0; base <= end_t[k]; base += 32) {
						const int t = base + lane;
						bool is_set = false; int pix = 0;
						if (t <= end_t[k]) {
							const int x = x0 + t * dx, y = y0 + t * dy;
							const int j1 = xflag ? x : x >> shift, i1 = xflag ? y >> shift : y;
							pix = i1 * W + j1;
							is_set = mask[pix] != 0;
						}
						const unsigned set = __ballot_sync(0xffffffffu, is_set);
						if (is_set) {
							mask[pix] = 0;
							const int slot = listed + __popc(set & ((1u << lane) - 1));
							if (good && slot < kHoughMaxLinePixels) s_px[slot] = pix;
						}
						listed += __popc(set);
						__syncwarp();
					}
				}
				if (lane == 0) {
					s_ctl[4] = good ? min(listed, kHoughMaxLinePixels) : 0;
					s_ctl[5] = good ? 1 : 0;
					if (good) {
						const int slot = atomicAdd(num_lines, 1);
						if (slot < max_lines) lines[slot] = make_int4(ex[0], ey[0], ex[1], ey[1]);
					}
				}
			}
			__syncthreads();
			const int nlist = s_ctl[4];
			if (tid < numangle)
				for (int q = 0; q < nlist; ++q) {
					const int pix = s_px[q];
					const int jj = pix % W, ii = pix / W;
					const int r = __float2int_rn(__fadd_rn(__fmul_rn((float)jj, t_cos), __fmul_rn((float)ii, t_sin))) + half_rho;
					accum[(size_t)tid * numrho + r]--;
				}
			__syncthreads();
		}
		// leave the slab clean for the next region: votes that were never removed, mask pixels that never became a line
		for (size_t q = tid; q < (size_t)numangle * numrho; q += kHoughThreads) accum[q] = 0;
		for (int q = tid; q < npts; q += kHoughThreads) mask[loc[2 * q]] = 0;
		__syncthreads();
	}
}

// cv::line(img, p0, p1, 255, 1) (imgproc/src/drawing.cpp, LineIterator, 8-connected, left to right); one thread per line
__global__ void __launch_bounds__(128) k_draw_lines(uint8_t* img, int w, int h, const int4* __restrict__ lines, int n) {
	const int q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= n) return;
	int x0 = lines[q].x, y0 = lines[q].y; const int x1 = lines[q].z, y1 = lines[q].w;
	if (x0 < 0 || x0 >= w || x1 < 0 || x1 >= w || y0 < 0 || y0 >= h || y1 < 0 || y1 >= h) return;
	int dx = x1 - x0, dy = y1 - y0;
	if (dx < 0) { dx = -dx; dy = -dy; x0 = x1; y0 = y1; }
	int major_x = 1, major_y = 0, minor_x = 0, minor_y = dy < 0 ? -1 : 1;
	if (dy < 0) dy = -dy;
	if (dy > dx) { const int t = dx; dx = dy; dy = t; major_x = 0; major_y = minor_y; minor_x = 1; minor_y = 0; }
	int err = dx - (dy + dy);
	const int plus_delta = dx + dx, minus_delta = -(dy + dy);
	int x = x0, y = y0;
	for (int i = 0; i <= dx; ++i) {
		img[(size_t)y * w + x] = 255;
		const bool both = err < 0;
		err += minus_delta + (both ? plus_delta : 0);
		x += major_x + (both ? minor_x : 0);
		y += major_y + (both ? minor_y : 0);
	}
}

// final classes (APD.cpp:486-492): 0 on edge pixels, -1 in regions of <= weak_tex_num pixels, a positive id (root + 1) elsewhere
__global__ void __launch_bounds__(256) k_lab_final(int n, const int* __restrict__ parent, const int* __restrict__ count, int weak_tex_num, int32_t* __restrict__ labels) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n) return;
	const int r = parent[p];
	labels[p] = r < 0 ? 0 : (count[r] <= weak_tex_num ? -1 : r + 1);
}

}  // namespace
}  // namespace dvp

#include <cub/device/device_radix_sort.cuh>
#include <cmath>
#include <vector>
#include <cstdio>

namespace dvp {
namespace {

cudaError_t resize8u(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh, cudaStream_t st) {
	const double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
	const int iscale_x = (int)std::lround(scale_x), iscale_y = (int)std::lround(scale_y);
	const bool area_fast = std::abs(scale_x - iscale_x) < 2.220446049250313e-16 && std::abs(scale_y - iscale_y) < 2.220446049250313e-16;
	dim3 b(32, 8), g((dw + 31) / 32, (dh + 7) / 8);
	k_resize8u<<<g, b, 0, st>>>(src, sw, sh, dst, dw, dh, scale_x, scale_y, (area_fast && iscale_x == 2 && iscale_y == 2) ? 1 : 0);
	return cudaGetLastError();
}

// connected regions of the non-255 pixels: parent[p] = root pixel (or -1), count[root] = size
cudaError_t label_regions(const uint8_t* img, int W, int H, int* parent, int* count, cudaStream_t st) {
	const int n = W * H;
	dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8);
	k_lab_init<<<(n + 255) / 256, 256, 0, st>>>(img, n, parent, count);
	k_lab_link<<<g, b, 0, st>>>(W, H, parent);
	k_lab_count<<<(n + 255) / 256, 256, 0, st>>>(n, parent, count);
	return cudaGetLastError();
}

struct DeviceBuf {   // frees on scope exit
	void* p = nullptr;
	~DeviceBuf() { cudaFree(p); }
	cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
	template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};

}  // namespace
}  // namespace dvp

extern "C" {

// Level size of the label map: round(cols / 2^scale) in float arithmetic (APD.cpp:441-443)
int dvp_label_size(int cols, int rows, int scale, int* new_cols, int* new_rows) {
	if (!new_cols || !new_rows || cols <= 0 || rows <= 0 || scale < 0 || scale > 8) return DVP_ERR_ARG;
	const float factor = 1.0f / (float)(1 << scale);
	*new_cols = (int)std::round(cols * factor);
	*new_rows = (int)std::round(rows * factor);
	return DVP_OK;
}

int dvp_label_segment(int device, const uint8_t* image, int cols, int rows, int scale, int32_t* labels, uint8_t* edge_small, float* device_ms) {
	if (!image || !labels || cols < 16 || rows < 16 || scale < 0 || scale > 8 || (long long)cols * rows > 0x7fffffffLL) return DVP_ERR_ARG;
	if (cudaSetDevice(device) != cudaSuccess) return DVP_ERR_CUDA;
	const int weak_tex_num = (int)(1.0 * rows * cols / (1024 << scale << scale));
	const int w1 = cols / 2, h1 = rows / 2, w2 = w1 / 2, h2 = h1 / 2;
	int new_cols = 0, new_rows = 0;
	dvp_label_size(cols, rows, scale, &new_cols, &new_rows);
	if (w2 < 3 || h2 < 3 || new_cols < 3 || new_rows < 3) return DVP_ERR_ARG;
	const size_t n0 = (size_t)cols * rows, n1 = (size_t)w1 * h1, n2 = (size_t)w2 * h2, nl = (size_t)new_cols * new_rows;
	const size_t nmax = n2 > nl ? n2 : nl;
	const int m = w2 < h2 ? w2 : h2;
	const int houthr = (int)(m / 30.0), min_line_length = (int)(m / 30.0), max_line_gap = (int)(m / 30.0);
	// HoughLinesP set-up as OpenCV computes it (hough.cpp): angle count, rho range, cos / sin table in double, stored as float
	const float theta = (float)(3.14159265358979323846 / 180), rho = 1.0f, irho = 1 / rho;
	int numangle = (int)std::floor((3.14159265358979323846 - 0.0) / theta) + 1;
	if (numangle > 1 && std::fabs(3.14159265358979323846 - (numangle - 1) * theta) < theta / 2) --numangle;
	const int numrho = (int)std::nearbyint(((w2 + h2) * 2 + 1) / rho);
	if (numangle > kHoughThreads) return DVP_ERR_UNSUPPORTED;
	std::vector<float> trig((size_t)numangle * 2);
	for (int k = 0; k < numangle; k++) {
		trig[k * 2] = (float)(std::cos((double)k * theta) * irho);
		trig[k * 2 + 1] = (float)(std::sin((double)k * theta) * irho);
	}
	const int key_cap = (int)std::min<size_t>(4 * n2, (size_t)1 << 28);
	const int max_lines = 1 << 18;
	const int hough_ctas = 148;
	DeviceBuf d_src, d_down1, d_down2, d_dst, d_up, d_parent, d_count, d_region, d_scalars, d_keys, d_keys_alt, d_start, d_sort_tmp, d_trig, d_accum, d_mask, d_lines, d_labels;
	cudaStream_t st = 0;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	int rc = DVP_ERR_CUDA;
	do {
		if (d_src.alloc(n0) || d_down1.alloc(n1) || d_down2.alloc(n2) || d_dst.alloc(n2) || d_up.alloc(nl)) break;
		if (d_parent.alloc(nmax * 4) || d_count.alloc(nmax * 4) || d_region.alloc(n2 * 4) || d_scalars.alloc(16)) break;
		if (d_keys.alloc((size_t)key_cap * 8) || d_keys_alt.alloc((size_t)key_cap * 8) || d_trig.alloc(trig.size() * 4)) break;
		if (d_lines.alloc((size_t)max_lines * 16) || d_labels.alloc(nl * 4)) break;
		if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) break;
		if (cudaMemcpy(d_src.p, image, n0, cudaMemcpyDefault) != cudaSuccess) break;
		if (cudaMemcpy(d_trig.p, trig.data(), trig.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) break;
		if (cudaEventRecord(e0, st) != cudaSuccess) break;
		// two halvings, Roberts + threshold, regions of the quarter-size edge image
		if (resize8u(d_src.as<uint8_t>(), cols, rows, d_down1.as<uint8_t>(), w1, h1, st)) break;
		if (resize8u(d_down1.as<uint8_t>(), w1, h1, d_down2.as<uint8_t>(), w2, h2, st)) break;
		{ dim3 b(32, 8), g((w2 + 31) / 32, (h2 + 7) / 8); k_roberts_threshold<<<g, b, 0, st>>>(d_down2.as<uint8_t>(), w2, h2, d_dst.as<uint8_t>()); }
		if (label_regions(d_dst.as<uint8_t>(), w2, h2, d_parent.as<int>(), d_count.as<int>(), st)) break;
		int* scal = d_scalars.as<int>();   // [0] regions, [1] border keys, [2] lines
		if (cudaMemsetAsync(scal, 0, 16, st) != cudaSuccess) break;
		k_lab_big_regions<<<(int)((n2 + 255) / 256), 256, 0, st>>>((int)n2, d_parent.as<int>(), d_count.as<int>(), weak_tex_num, d_region.as<int>(), scal);
		{ dim3 b(32, 8), g((w2 + 31) / 32, (h2 + 7) / 8);
		  k_lab_border_keys<<<g, b, 0, st>>>(w2, h2, d_parent.as<int>(), d_region.as<int>(), d_keys.as<unsigned long long>(), scal + 1, key_cap); }
		int h_scal[4] = {0, 0, 0, 0};
		if (cudaMemcpyAsync(h_scal, scal, 16, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) break;
		const int num_regions = h_scal[0], num_keys = h_scal[1];
		if (num_keys > key_cap) { rc = DVP_ERR_UNSUPPORTED; break; }
		if (num_regions > 0 && num_keys > 0) {
			// per region, its border pixels in raster order: sort (region, pixel)
			size_t tmp_bytes = 0;
			unsigned long long* k_in = d_keys.as<unsigned long long>(); unsigned long long* k_out = d_keys_alt.as<unsigned long long>();
			if (cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, k_in, k_out, num_keys, 0, 64, st) != cudaSuccess) break;
			if (d_sort_tmp.alloc(tmp_bytes)) break;
			if (cub::DeviceRadixSort::SortKeys(d_sort_tmp.p, tmp_bytes, k_in, k_out, num_keys, 0, 64, st) != cudaSuccess) break;
			if (d_start.alloc((size_t)(num_regions + 1) * 4)) break;
			if (cudaMemsetAsync(d_start.p, 0, (size_t)(num_regions + 1) * 4, st) != cudaSuccess) break;
			k_lab_region_starts<<<(num_keys + 255) / 256, 256, 0, st>>>(k_out, num_keys, d_start.as<int>(), num_regions);
			// HoughLinesP per region; every CTA has its own accumulator and mask, kept all-zero between regions
			const int ctas = num_regions < hough_ctas ? num_regions : hough_ctas;
			if (d_accum.alloc((size_t)ctas * numangle * numrho * 4) || d_mask.alloc((size_t)ctas * n2)) break;
			if (cudaMemsetAsync(d_accum.p, 0, (size_t)ctas * numangle * numrho * 4, st) != cudaSuccess || cudaMemsetAsync(d_mask.p, 0, (size_t)ctas * n2, st) != cudaSuccess) break;
			k_hough_regions<<<ctas, kHoughThreads, 0, st>>>(w2, h2, num_regions, d_start.as<int>(), k_out, numangle, numrho, d_trig.as<float>(), houthr, min_line_length, max_line_gap,
			                                                d_accum.as<int>(), d_mask.as<uint8_t>(), d_lines.as<int4>(), scal + 2, max_lines);
			if (cudaMemcpyAsync(h_scal, scal, 16, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) break;
			const int num_lines = h_scal[2] < max_lines ? h_scal[2] : max_lines;
			if (num_lines > 0) k_draw_lines<<<(num_lines + 127) / 128, 128, 0, st>>>(d_dst.as<uint8_t>(), w2, h2, d_lines.as<int4>(), num_lines);
		}
		if (edge_small && cudaMemcpyAsync(edge_small, d_dst.p, n2, cudaMemcpyDefault, st) != cudaSuccess) break;
		// up to the level size, threshold, border clean-up (APD.cpp:452-463), regions, classes
		if (resize8u(d_dst.as<uint8_t>(), w2, h2, d_up.as<uint8_t>(), new_cols, new_rows, st)) break;
		k_threshold4<<<(int)((nl + 255) / 256), 256, 0, st>>>(d_up.as<uint8_t>(), nl);
		if (launch_border_cleanup(d_up.as<uint8_t>(), new_cols, new_rows, st)) break;
		if (label_regions(d_up.as<uint8_t>(), new_cols, new_rows, d_parent.as<int>(), d_count.as<int>(), st)) break;
		k_lab_final<<<(int)((nl + 255) / 256), 256, 0, st>>>((int)nl, d_parent.as<int>(), d_count.as<int>(), weak_tex_num, d_labels.as<int32_t>());
		if (cudaEventRecord(e1, st) != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess) break;
		if (cudaGetLastError() != cudaSuccess) break;
		if (cudaMemcpy(labels, d_labels.p, nl * 4, cudaMemcpyDefault) != cudaSuccess) break;
		if (device_ms && cudaEventElapsedTime(device_ms, e0, e1) != cudaSuccess) break;
		rc = DVP_OK;
	} while (0);
	if (rc == DVP_ERR_CUDA) fprintf(stderr, "[dvp] dvp_label_segment failed: %s\n", cudaGetErrorString(cudaGetLastError()));
	if (e0) cudaEventDestroy(e0);
	if (e1) cudaEventDestroy(e1);
	return rc;
}

}  // extern "C"
