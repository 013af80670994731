// tools/ubench_tex.cu — texture-unit throughput versus lane coherence (design input for the WEAK sweep).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/ubench_tex tools/ubench_tex.cu && gpurun_out/ubench_tex
// Variants (each thread issues the same number of bilinear fp32 fetches, 9 in flight):
//   A  coherent        : lanes = 32 adjacent pixels, all sampling the same patch offset (the STRONG sweep's shape)
//   B  lane-per-anchor : every lane walks the 9 samples of its OWN anchor (anchors scattered within +-R px of the pixel)
//   C  9-lanes-per-anchor : lanes l -> (anchor l/9, sample l%9); 3.5 anchors per instruction
//   D  as B, but the 32 lanes' anchors are the anchors of 3 pixels (11 anchors each) -> what a unit-per-lane warp sees
//   E  quad-coherent: the 4 lanes of a quad sample 4 adjacent pixels of one anchor, the 8 quads of a warp are scattered
//   F  pair-coherent: 2 adjacent lanes share an anchor neighbourhood, the 16 pairs are scattered
//   G  quad lanes `gap` pixels apart along x (gap = spread argument); H: the same in a 2x2 arrangement (x and y)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int MODE>
__global__ void __launch_bounds__(128) k_tex(cudaTextureObject_t tex, int W, int H, int reps, int spread, float* out) {
	const int lane = threadIdx.x & 31;
	const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int wpr = (W + 31) / 32;
	const int wy = (warp / wpr) % H, wx = (warp % wpr) * 32;
	float acc = 0.f;
	const int di[9] = {-5, -5, -5, 0, 0, 5, 5, 5, 0}, dj[9] = {-5, 0, 5, -5, 5, -5, 0, 5, 0};
	for (int r = 0; r < reps; ++r) {
		const float shift = 0.37f * r;   // a different hypothesis: slightly different mapping
		float v[9];
		if (MODE == 0) {
			const float x = (wx + lane) * 1.01f + shift, y = wy * 0.99f + shift;
#pragma unroll
			for (int q = 0; q < 9; ++q) v[q] = tex2D<float>(tex, x + di[q] + 0.5f, y + dj[q] + 0.5f);
		} else if (MODE == 4 || MODE == 5) {
			const int grp = (MODE == 4) ? 4 : 2;
			const unsigned key = hash(warp * 32 + lane / grp);
			const int ax = wx + (lane / grp) * grp + (lane % grp) + (int)(key % (2 * spread + 1)) - spread;
			const int ay = wy + (int)((key >> 12) % (2 * spread + 1)) - spread;
			const float x = ax * 1.01f + shift, y = ay * 0.99f + shift;
#pragma unroll
			for (int q = 0; q < 9; ++q) v[q] = tex2D<float>(tex, x + di[q] + 0.5f, y + dj[q] + 0.5f);
		} else if (MODE == 6 || MODE == 7) {
			const int gap = spread;   // reused argument: distance between the quad's lanes
			const unsigned key = hash(warp * 32 + lane / 4);
			const int l4 = lane % 4;
			const int ax = wx + (int)(key % 301) - 150 + (MODE == 6 ? l4 * gap : (l4 & 1) * gap);
			const int ay = wy + (int)((key >> 12) % 301) - 150 + (MODE == 6 ? 0 : (l4 >> 1) * gap);
			const float x = ax * 1.01f + shift, y = ay * 0.99f + shift;
#pragma unroll
			for (int q = 0; q < 9; ++q) v[q] = tex2D<float>(tex, x + di[q] + 0.5f, y + dj[q] + 0.5f);
		} else if (MODE == 1 || MODE == 3) {
			// own anchor per lane; MODE 3: anchors of pixel (wx + lane/11) -> neighbouring lanes' anchors belong to the same pixel
			const unsigned key = (MODE == 1) ? hash(warp * 32 + lane) : hash((warp * 3 + lane / 11) * 16 + lane % 11);
			const int ax = wx + (MODE == 3 ? lane / 11 : lane) + (int)(key % (2 * spread + 1)) - spread;
			const int ay = wy + (int)((key >> 12) % (2 * spread + 1)) - spread;
			const float x = ax * 1.01f + shift, y = ay * 0.99f + shift;
#pragma unroll
			for (int q = 0; q < 9; ++q) v[q] = tex2D<float>(tex, x + di[q] + 0.5f, y + dj[q] + 0.5f);
		} else {
			// 9 lanes per anchor: 9 instructions cover 32 anchors x 9 samples = 288 (anchor, sample) pairs
#pragma unroll
			for (int q = 0; q < 9; ++q) {
				const int g = q * 32 + lane;          // flat (anchor, sample)
				const int an = g / 9, sq = g - an * 9;
				const unsigned key = hash(warp * 32 + an);
				const int ax = wx + an + (int)(key % (2 * spread + 1)) - spread;
				const int ay = wy + (int)((key >> 12) % (2 * spread + 1)) - spread;
				const float x = ax * 1.01f + shift, y = ay * 0.99f + shift;
				v[q] = tex2D<float>(tex, x + di[sq] + 0.5f, y + dj[sq] + 0.5f);
			}
		}
#pragma unroll
		for (int q = 0; q < 9; ++q) acc += v[q];
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
static void run(const char* name, cudaTextureObject_t tex, int W, int H, int reps, int spread, float* out, int nthreads) {
	cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	for (int i = 0; i < 2; ++i) k_tex<MODE><<<nthreads / 128, 128>>>(tex, W, H, reps, spread, out);
	CK(cudaEventRecord(e0));
	for (int i = 0; i < 3; ++i) k_tex<MODE><<<nthreads / 128, 128>>>(tex, W, H, reps, spread, out);
	CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
	float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 3;
	const double fetches = (double)nthreads * reps * 9;
	printf("%-32s arg=%4d  %8.3f ms  %8.1f Gfetch/s\n", name, spread, ms, fetches / ms / 1e6);
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
	const int reps = 64;
	run<0>("A coherent", tex, W, H, reps, 0, out, nthreads);
	for (int spread : {8, 40, 150}) {
		run<1>("B lane-per-anchor", tex, W, H, reps, spread, out, nthreads);
		run<3>("D lane-per-anchor (3 px)", tex, W, H, reps, spread, out, nthreads);
		run<2>("C 9-lanes-per-anchor", tex, W, H, reps, spread, out, nthreads);
		run<4>("E quad-coherent", tex, W, H, reps, spread, out, nthreads);
		run<5>("F pair-coherent", tex, W, H, reps, spread, out, nthreads);
	}
	for (int gap : {1, 2, 3, 4, 5, 8, 16}) {
		run<6>("G quad, lanes gap px apart (x)", tex, W, H, reps, gap, out, nthreads);
		run<7>("H quad, 2x2 lanes gap px apart", tex, W, H, reps, gap, out, nthreads);
	}
	return 0;
}
