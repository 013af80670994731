// dvp_kernels_edge.cu — the depth-edge prior on the device (edge half): what GetProblemEdges
// computes per view and pyramid level with EdgeSegment(scale, image, mode 0, use_canny = true) and the hot path then
// consumes as `edge_cuda` (reference main.cpp:193-226, APD.cpp:348-466):
//   histogram "median" of the 8-bit level image -> Canny thresholds (APD.cpp:405-432), cv::Canny(.., 3, L2gradient=true)
//   (OpenCV imgproc/src/canny.cpp: 3x3 Sobel with replicated borders, squared L2 magnitude, fixed-point non-maximum
//   suppression, 8-neighbour hysteresis), then the border clean-up of APD.cpp:452-463.  Integer work: bit-exact.
// Kernels:
//   k_edge_histogram     256 bins in shared memory per block, merged with global atomics
//   k_edge_thresholds    one thread: the reference's float-accumulated median and the squared Canny thresholds
//   k_edge_nms           32x8 tile + 2-pixel halo of the image in shared memory -> Sobel + magnitude for the tile and a
//                        1-pixel ring -> suppression -> map {0 weak candidate, 1 no edge, 2 strong}
//   k_edge_link/_seed/_apply  hysteresis as connected components: candidates are joined with their 8-neighbours by a
//                        lock-free union-find; a component is an edge iff it holds a strong pixel (what the reference's
//                        stack-based flood computes, independent of visiting order)
//   k_edge_border_cols/_rows  APD.cpp:452-463, columns first, then rows
#include "dvp_common.cuh"
#include "dvp_launch.h"
#include "dvp_unionfind.cuh"
#include <cstdio>

namespace dvp {

namespace {

constexpr int kTileW = 32, kTileH = 8;

__global__ void __launch_bounds__(256) k_edge_histogram(const uint8_t* __restrict__ img, size_t n, unsigned* __restrict__ hist) {
	__shared__ unsigned s[256];
	s[threadIdx.x] = 0;
	__syncthreads();
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) atomicAdd(&s[img[i]], 1u);
	__syncthreads();
	if (s[threadIdx.x]) atomicAdd(&hist[threadIdx.x], s[threadIdx.x]);
}

// APD.cpp:405-432 + the threshold preparation of cv::Canny for L2gradient (swap, square when positive).
// thr = {threshold1, threshold2, low, high}.  The reference counts in float (increments stop at 2^24) and adds the
// float bin to an int running sum: reproduced with explicitly rounded float operations.
__global__ void k_edge_thresholds(const unsigned* __restrict__ hist, int rows, int cols, int* __restrict__ thr) {
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	const int half = rows * cols / 2;
	int median_val = -1, temp_sum = 0;
	for (int i = 0; i < 255; i++) {
		const float bin = __uint2float_rn(min(hist[i], 16777216u));
		temp_sum = __float2int_rz(__fadd_rn(__int2float_rn(temp_sum), bin));
		if (temp_sum > half) { median_val = i; break; }
	}
	const int t1 = __float2int_rz(__fmul_rn(__fsub_rn(1.0f, 0.67f), __int2float_rn(median_val)));
	const int t2 = median_val;
	int low = t1, high = t2;
	if (low > high) { const int t = low; low = high; high = t; }
	if (low > 0) low *= low;
	if (high > 0) high *= high;
	thr[0] = t1; thr[1] = t2; thr[2] = low; thr[3] = high;
}

__global__ void __launch_bounds__(kTileW * kTileH) k_edge_nms(const uint8_t* __restrict__ img, int W, int H, const int* __restrict__ thr, uint8_t* __restrict__ map) {
	__shared__ uint8_t s_img[kTileH + 4][kTileW + 4];
	__shared__ int s_mag[kTileH + 2][kTileW + 2];
	__shared__ short2 s_g[kTileH + 2][kTileW + 2];
	const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
	const int tid = threadIdx.y * kTileW + threadIdx.x;
	for (int i = tid; i < (kTileH + 4) * (kTileW + 4); i += kTileW * kTileH) {
		const int ly = i / (kTileW + 4), lx = i - ly * (kTileW + 4);
		const int gy = min(max(y0 + ly - 2, 0), H - 1), gx = min(max(x0 + lx - 2, 0), W - 1);   // BORDER_REPLICATE
		s_img[ly][lx] = img[(size_t)gy * W + gx];
	}
	__syncthreads();
	for (int i = tid; i < (kTileH + 2) * (kTileW + 2); i += kTileW * kTileH) {
		const int ly = i / (kTileW + 2), lx = i - ly * (kTileW + 2);
		const int gy = y0 + ly - 1, gx = x0 + lx - 1;
		int dx = 0, dy = 0, m = 0;
		if (gy >= 0 && gy < H && gx >= 0 && gx < W) {          // outside the image the magnitude frame is zero
			const int a = s_img[ly][lx], b = s_img[ly][lx + 1], c = s_img[ly][lx + 2];
			const int d = s_img[ly + 1][lx], f = s_img[ly + 1][lx + 2];
			const int g = s_img[ly + 2][lx], h = s_img[ly + 2][lx + 1], k = s_img[ly + 2][lx + 2];
			dx = (c + 2 * f + k) - (a + 2 * d + g);
			dy = (g + 2 * h + k) - (a + 2 * b + c);
			m = dx * dx + dy * dy;
		}
		s_mag[ly][lx] = m;
		s_g[ly][lx] = make_short2((short)dx, (short)dy);
	}
	__syncthreads();
	const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
	if (x >= W || y >= H) return;
	const int ly = threadIdx.y + 1, lx = threadIdx.x + 1;
	const int low = thr[2], high = thr[3];
	const int m = s_mag[ly][lx];
	bool candidate = false;
	if (m > low) {
		const int xs = s_g[ly][lx].x, ys = s_g[ly][lx].y;
		const int ax = abs(xs), ay = abs(ys) << 15;
		const int tg22x = ax * 13573;                              // (int)(tan(22.5 deg) * 2^15 + 0.5)
		if (ay < tg22x) candidate = m > s_mag[ly][lx - 1] && m >= s_mag[ly][lx + 1];
		else {
			const int tg67x = tg22x + (ax << 16);
			if (ay > tg67x) candidate = m > s_mag[ly - 1][lx] && m >= s_mag[ly + 1][lx];
			else {
				const int s = (xs ^ ys) < 0 ? -1 : 1;
				candidate = m > s_mag[ly - 1][lx - s] && m > s_mag[ly + 1][lx + s];
			}
		}
	}
	map[(size_t)y * W + x] = !candidate ? 1 : (m > high ? 2 : 0);
}

__global__ void __launch_bounds__(256) k_edge_uf_init(const uint8_t* __restrict__ map, int n, int* __restrict__ parent, int* __restrict__ flag) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n) return;
	parent[p] = map[p] != 1 ? p : -1;
	flag[p] = 0;
}

// join with the four neighbours already "behind" the pixel in raster order (the other four are joined from their side)
__global__ void __launch_bounds__(256) k_edge_link(int W, int H, int* parent) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= W || y >= H) return;
	const int p = y * W + x;
	if (parent[p] < 0) return;
	if (x > 0 && parent[p - 1] >= 0) uf_union(parent, p, p - 1);
	if (y > 0) {
		if (parent[p - W] >= 0) uf_union(parent, p, p - W);
		if (x > 0 && parent[p - W - 1] >= 0) uf_union(parent, p, p - W - 1);
		if (x < W - 1 && parent[p - W + 1] >= 0) uf_union(parent, p, p - W + 1);
	}
}

__global__ void __launch_bounds__(256) k_edge_seed(const uint8_t* __restrict__ map, int n, const int* __restrict__ parent, int* __restrict__ flag) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p < n && map[p] == 2) flag[uf_find(parent, p)] = 1;
}

// the final map, already through cv::threshold(> 4 -> 255) (APD.cpp:446)
__global__ void __launch_bounds__(256) k_edge_apply(int n, const int* __restrict__ parent, const int* __restrict__ flag, uint8_t* __restrict__ edge) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n) return;
	edge[p] = (parent[p] >= 0 && flag[uf_find(parent, p)]) ? 255 : 0;
}

// APD.cpp:452-457
__global__ void __launch_bounds__(256) k_edge_border_cols(int W, int H, uint8_t* edge) {
	const int y = blockIdx.x * blockDim.x + threadIdx.x;
	if (y >= H) return;
	uint8_t* row = edge + (size_t)y * W;
	if (row[1] == 0) row[0] = 0;
	if (row[W - 2] == 0) row[W - 1] = 0;
}
// APD.cpp:458-463
__global__ void __launch_bounds__(256) k_edge_border_rows(int W, int H, uint8_t* edge) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	if (x >= W) return;
	if (edge[(size_t)W + x] == 0) edge[x] = 0;
	if (edge[(size_t)(H - 2) * W + x] == 0) edge[(size_t)(H - 1) * W + x] = 0;
}

// GetProblemEdges: scaled_image_float.convertTo(src_img, CV_8UC1) (main.cpp:209) = saturate_cast<uchar>(cvRound(v)),
// round half to even
__global__ void __launch_bounds__(256) k_edge_to_u8(const float* __restrict__ img, int n, uint8_t* __restrict__ out) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n) return;
	const int v = __float2int_rn(img[p]);
	out[p] = (uint8_t)min(max(v, 0), 255);
}

}  // namespace

size_t edge_scratch_bytes(int W, int H) {
	const size_t n = (size_t)W * H;
	return 4 * sizeof(int) /*thresholds*/ + 256 * sizeof(unsigned) /*histogram*/ + 2 * n * sizeof(int) /*parent, flag*/ + n /*map*/ + 64;
}

// d_img, d_edge: [H][W] u8 on the device; scratch: edge_scratch_bytes(W, H) bytes; *d_thr (d_thr may be null) points at
// the 4 ints {threshold1, threshold2, low, high} the run leaves at the start of the scratch.
cudaError_t launch_edge_segment(const uint8_t* d_img, int W, int H, uint8_t* d_edge, void* scratch, int** d_thr, cudaStream_t st) {
	const size_t n = (size_t)W * H;
	char* base = (char*)scratch;
	int* thr = (int*)base;                       base += 4 * sizeof(int);
	unsigned* hist = (unsigned*)base;            base += 256 * sizeof(unsigned);
	int* parent = (int*)base;                    base += n * sizeof(int);
	int* flag = (int*)base;                      base += n * sizeof(int);
	uint8_t* map = (uint8_t*)base;
	if (d_thr) *d_thr = thr;
	cudaError_t e = cudaMemsetAsync(hist, 0, 256 * sizeof(unsigned), st);
	if (e != cudaSuccess) return e;
	const int blocks = (int)((n + 255) / 256);
	k_edge_histogram<<<min(blocks, 148 * 8), 256, 0, st>>>(d_img, n, hist);
	k_edge_thresholds<<<1, 32, 0, st>>>(hist, H, W, thr);
	const dim3 tb(kTileW, kTileH), tg((W + kTileW - 1) / kTileW, (H + kTileH - 1) / kTileH);
	k_edge_nms<<<tg, tb, 0, st>>>(d_img, W, H, thr, map);
	k_edge_uf_init<<<blocks, 256, 0, st>>>(map, (int)n, parent, flag);
	k_edge_link<<<tg, tb, 0, st>>>(W, H, parent);
	k_edge_seed<<<blocks, 256, 0, st>>>(map, (int)n, parent, flag);
	k_edge_apply<<<blocks, 256, 0, st>>>((int)n, parent, flag, d_edge);
	k_edge_border_cols<<<(H + 255) / 256, 256, 0, st>>>(W, H, d_edge);
	k_edge_border_rows<<<(W + 255) / 256, 256, 0, st>>>(W, H, d_edge);
	return cudaGetLastError();
}

// APD.cpp:452-463 on a 0 / 255 image (shared with the label half, dvp_kernels_image.cu)
cudaError_t launch_border_cleanup(uint8_t* d_img, int W, int H, cudaStream_t st) {
	k_edge_border_cols<<<(H + 255) / 256, 256, 0, st>>>(W, H, d_img);
	k_edge_border_rows<<<(W + 255) / 256, 256, 0, st>>>(W, H, d_img);
	return cudaGetLastError();
}

cudaError_t launch_edge_to_u8(const float* d_img, int n, uint8_t* d_out, cudaStream_t st) {
	k_edge_to_u8<<<(n + 255) / 256, 256, 0, st>>>(d_img, n, d_out);
	return cudaGetLastError();
}

}  // namespace dvp

extern "C" int dvp_edge_segment(int device, const uint8_t* image, int width, int height, uint8_t* edge, int32_t* thresholds, float* device_ms) {
	if (!image || !edge || width < 3 || height < 3 || (long long)width * height > 0x7fffffffLL) return DVP_ERR_ARG;
	if (cudaSetDevice(device) != cudaSuccess) return DVP_ERR_CUDA;
	const size_t n = (size_t)width * height;
	uint8_t* d_img = nullptr; uint8_t* d_edge = nullptr; void* scratch = nullptr;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	int* d_thr = nullptr;
	int rc = DVP_ERR_CUDA;
	do {
		if (cudaMalloc((void**)&d_img, n) != cudaSuccess || cudaMalloc((void**)&d_edge, n) != cudaSuccess) break;
		if (cudaMalloc(&scratch, dvp::edge_scratch_bytes(width, height)) != cudaSuccess) break;
		if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) break;
		if (cudaMemcpy(d_img, image, n, cudaMemcpyDefault) != cudaSuccess) break;        // host or device source
		if (cudaEventRecord(e0, 0) != cudaSuccess) break;
		if (dvp::launch_edge_segment(d_img, width, height, d_edge, scratch, &d_thr, 0) != cudaSuccess) break;
		if (cudaEventRecord(e1, 0) != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess) break;
		if (cudaMemcpy(edge, d_edge, n, cudaMemcpyDefault) != cudaSuccess) break;
		if (thresholds && cudaMemcpy(thresholds, d_thr, 2 * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) break;
		if (device_ms && cudaEventElapsedTime(device_ms, e0, e1) != cudaSuccess) break;
		rc = DVP_OK;
	} while (0);
	if (rc != DVP_OK) fprintf(stderr, "[dvp] dvp_edge_segment failed: %s\n", cudaGetErrorString(cudaGetLastError()));
	if (e0) cudaEventDestroy(e0);
	if (e1) cudaEventDestroy(e1);
	cudaFree(d_img); cudaFree(d_edge); cudaFree(scratch);
	return rc;
}
