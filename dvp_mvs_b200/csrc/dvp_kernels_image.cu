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
