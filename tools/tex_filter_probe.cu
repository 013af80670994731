// tools/tex_filter_probe.cu — what arithmetic does the B200 texture unit use for a bilinear fp32 fetch when the
// texels are NOT 8-bit-valued?  (tests/golden/tex_probe.npz pinned the coordinate quantisation and the four 8-bit
// weights on integer-valued texels, where every candidate arithmetic gives the same float.)  A software sampler that
// is to replace tex2D bit for bit needs the answer for arbitrary floats.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tex_filter_probe tools/tex_filter_probe.cu
//   gpurun_out/tex_filter_probe gpurun_out/tex_filter_probe.bin       (analysed by tools/tex_filter_fit.py)
// File layout: int32 W, H, n; float image[H*W]; float xy[n*2]; float result[n].
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void k_probe(cudaTextureObject_t tex, const float2* xy, float* out, int n) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = tex2D<float>(tex, xy[i].x, xy[i].y);
}

static uint64_t s_state = 0x9E3779B97F4A7C15ull;
static inline uint32_t rnd() { s_state ^= s_state << 13; s_state ^= s_state >> 7; s_state ^= s_state << 17; return (uint32_t)(s_state >> 32); }
static inline double urand() { return rnd() / 4294967296.0; }

int main(int argc, char** argv) {
	const char* path = argc > 1 ? argv[1] : "gpurun_out/tex_filter_probe.bin";
	const int W = 96, H = 64, n = 1 << 18;
	std::vector<float> img((size_t)W * H);
	for (int y = 0; y < H; ++y)
		for (int x = 0; x < W; ++x) {
			float v;
			if (x < 32) v = (float)(urand() * 255.0);                                   // grey levels with fractions (resized images)
			else if (x < 64) v = (float)std::ldexp(urand() + 0.5, (int)(rnd() % 24) - 8);  // wide dynamic range
			else v = (float)(rnd() % 256);                                               // 8-bit-valued control region
			img[(size_t)y * W + x] = v;
		}
	std::vector<float> xy((size_t)n * 2);
	for (int i = 0; i < n; ++i) {
		const int region = i % 3;
		xy[2 * i] = (float)(region * 32 + 1 + urand() * 29.5);   // footprints stay inside one region
		xy[2 * i + 1] = (float)(1 + urand() * (H - 2.5));
		if (i % 97 == 0) { xy[2 * i] = (float)(urand() * (W + 4) - 2); xy[2 * i + 1] = (float)(urand() * (H + 4) - 2); }   // borders, clamp
	}
	cudaArray_t arr; cudaChannelFormatDesc cd = cudaCreateChannelDesc<float>();
	CK(cudaMallocArray(&arr, &cd, W, H));
	CK(cudaMemcpy2DToArray(arr, 0, 0, img.data(), W * 4, W * 4, H, cudaMemcpyHostToDevice));
	cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
	cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModeLinear;
	td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
	cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
	float2* d_xy; float* d_out;
	CK(cudaMalloc(&d_xy, (size_t)n * 8)); CK(cudaMalloc(&d_out, (size_t)n * 4));
	CK(cudaMemcpy(d_xy, xy.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
	k_probe<<<(n + 255) / 256, 256>>>(tex, d_xy, d_out, n);
	CK(cudaDeviceSynchronize());
	std::vector<float> out(n);
	CK(cudaMemcpy(out.data(), d_out, (size_t)n * 4, cudaMemcpyDeviceToHost));
	FILE* f = fopen(path, "wb");
	if (!f) { printf("cannot open %s\n", path); return 1; }
	const int32_t hdr[3] = {W, H, n};
	fwrite(hdr, 4, 3, f); fwrite(img.data(), 4, img.size(), f); fwrite(xy.data(), 4, xy.size(), f); fwrite(out.data(), 4, out.size(), f);
	fclose(f);
	printf("wrote %s: %d fetches of a %dx%d float texture\n", path, n, W, H);
	return 0;
}
