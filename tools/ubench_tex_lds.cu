// tools/ubench_tex_lds.cu — does shared-memory traffic cost texture throughput?  (design input for the NCC core)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/ubench_tex_lds tools/ubench_tex_lds.cu && gpurun_out/ubench_tex_lds
// The NCC inner loop of the product takes, per source sample, one bilinear fetch through the texture unit and one LDS.64
// of the hoisted (w, w r) pair from a per-thread column of shared memory (dvp_ncc.cuh).  Texture fetches and shared-memory
// loads both go through the SM's L1TEX unit.  Every variant below issues the SAME coherent fetch pattern (32 adjacent
// pixels sample the same patch offset, 12 fetches in flight per thread, as the wide kernels do) and differs only in the
// shared-memory loads that accompany the fetches:
//   0  none                     4  one LDS.64 per fetch + the product's 5 FP32 operations per sample
//   1  one LDS.32 per fetch     5  one LDS.128 per two fetches (the bytes of variant 2 in half the instructions)
//   2  one LDS.64 per fetch     6  as 4 with the pairs held in registers instead (no shared-memory load)
//   3  two LDS.64 per fetch
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int NS = 36;      // samples per NCC
constexpr int RB = 12;      // fetches in flight per thread (two patch rows)

template <int MODE, int T, int MINB>
__global__ void __launch_bounds__(T, MINB) k_mix(cudaTextureObject_t tex, int W, int H, int reps, float* out) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	float2* wt = reinterpret_cast<float2*>(smem_raw) + threadIdx.x;            // column per thread: wt[k * T], conflict-free
	float4* wq = reinterpret_cast<float4*>(smem_raw) + threadIdx.x;            // the same bytes as 18 float4 per thread
	const int lane = threadIdx.x & 31;
	const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int wpr = (W + 31) / 32;
	const int wy = (warp / wpr) % H, wx = (warp % wpr) * 32;
	for (int k = 0; k < NS; ++k) wt[k * T] = make_float2(1.0f + 0.001f * k + 1e-6f * threadIdx.x, 0.5f + 0.002f * k);
	float reg_w[NS], reg_t[NS];
	if (MODE == 6) {
#pragma unroll
		for (int k = 0; k < NS; ++k) { reg_w[k] = wt[k * T].x; reg_t[k] = wt[k * T].y; }
	}
	__syncthreads();
	float s0 = 0.f, s1 = 0.f, s2 = 0.f;
	for (int r = 0; r < reps; ++r) {
		const float shift = 0.37f * (r & 63);   // a different hypothesis: slightly different mapping
		const float x = (wx + lane) * 1.01f + shift, y = wy * 0.99f + shift;
#pragma unroll
		for (int b = 0; b < NS / RB; ++b) {
			float v[RB];
#pragma unroll
			for (int q = 0; q < RB; ++q) {
				const int k = b * RB + q;
				v[q] = tex2D<float>(tex, x + (float)(2 * (k / 6) - 5) + 0.5f, y + (float)(2 * (k % 6) - 5) + 0.5f);
			}
#pragma unroll
			for (int q = 0; q < RB; ++q) {
				const int k = b * RB + q;
				const float s = v[q];
				if (MODE == 0) { s0 += s; }
				else if (MODE == 1) { s0 = fmaf(s, reinterpret_cast<const float*>(smem_raw)[k * T + threadIdx.x], s0); }
				else if (MODE == 2) { const float2 w = wt[k * T]; s0 = fmaf(s, w.x, s0); s1 = fmaf(s, w.y, s1); }
				else if (MODE == 3) { const float2 w = wt[k * T]; const float2 w2 = wt[((k + 7) % NS) * T]; s0 = fmaf(s, w.x, s0); s1 = fmaf(s, w.y, s1); s2 = fmaf(w2.x, w2.y, s2); }
				else if (MODE == 4) { const float2 w = wt[k * T]; const float u = s * w.x; s2 = fmaf(s, w.y, s2); s1 = fmaf(s, u, s1); s0 += u; }
				else if (MODE == 5) { if ((q & 1) == 0) { const float4 w = wq[(k / 2) * T]; s0 = fmaf(s, w.x, s0); s1 = fmaf(v[q + 1], w.z, s1); s2 = fmaf(s, w.y, s2); s2 = fmaf(v[q + 1], w.w, s2); } }
				else { const float u = s * reg_w[k]; s2 = fmaf(s, reg_t[k], s2); s1 = fmaf(s, u, s1); s0 += u; }
			}
		}
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2;
}

template <int MODE, int T = 256, int MINB = 3>   // 256 threads, 3 blocks per SM: k_depth_to_weak_refine's shape
static void run(const char* name, cudaTextureObject_t tex, int W, int H, int reps, float* out, int nthreads) {
	const int smem = NS * T * (int)sizeof(float2);   // 73 728 B, the product's table
	CK(cudaFuncSetAttribute(k_mix<MODE, T, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
	cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	for (int i = 0; i < 2; ++i) k_mix<MODE, T, MINB><<<nthreads / T, T, smem>>>(tex, W, H, reps, out);
	CK(cudaEventRecord(e0));
	for (int i = 0; i < 3; ++i) k_mix<MODE, T, MINB><<<nthreads / T, T, smem>>>(tex, W, H, reps, out);
	CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
	CK(cudaGetLastError());
	float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 3;
	const double fetches = (double)nthreads * reps * NS;
	printf("%-64s %8.3f ms  %8.1f Gfetch/s\n", name, ms, fetches / ms / 1e6);
}

int main() {
	const int W = 3111, H = 2073;
	std::vector<float> img((size_t)W * H);
	for (size_t i = 0; i < img.size(); ++i) img[i] = (float)(rand() % 256);
	cudaArray_t arr; cudaChannelFormatDesc cd = cudaCreateChannelDesc<float>();
	CK(cudaMallocArray(&arr, &cd, W, H));
	CK(cudaMemcpy2DToArray(arr, 0, 0, img.data(), W * 4, W * 4, H, cudaMemcpyHostToDevice));
	cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
	cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModeLinear;
	td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
	cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
	const int nthreads = ((W + 31) / 32) * 32 * 1024;   // 1024 rows of warps
	float* out; CK(cudaMalloc(&out, (size_t)nthreads * 4));
	const int reps = 16;
	run<0>("0 fetch only", tex, W, H, reps, out, nthreads);
	run<1>("1 + LDS.32 per fetch", tex, W, H, reps, out, nthreads);
	run<2>("2 + LDS.64 per fetch", tex, W, H, reps, out, nthreads);
	run<3>("3 + two LDS.64 per fetch", tex, W, H, reps, out, nthreads);
	run<5>("5 + LDS.128 per two fetches", tex, W, H, reps, out, nthreads);
	run<4>("4 + LDS.64 per fetch + the NCC's 5 FP32 operations", tex, W, H, reps, out, nthreads);
	run<6, 256, 1>("6 pairs in registers + the NCC's 5 FP32 operations, 8 warps / SM", tex, W, H, reps, out, nthreads);
	run<6, 128, 3>("6 the same in 128-thread blocks, 12 warps / SM", tex, W, H, reps, out, nthreads);
	run<4, 128, 4>("4 in 128-thread blocks, 16 warps / SM", tex, W, H, reps, out, nthreads);
	run<0, 128, 2>("0 fetch only, 8 warps / SM", tex, W, H, reps, out, nthreads);
	return 0;
}
