// tools/ubench_sample.cu — can the bilinear source fetch of the NCC inner loop leave the texture unit?
// (VERDICT r01 item 4 / DESIGN "the TEX roof itself").  Same arithmetic around the fetch as dvp_ncc.cuh's hoisted loop
// (homography apply, rcp, (w, w r) table in shared memory, row-then-total sums), five ways of getting the sample:
//   T  tex2D<float>, hardware bilinear (what the product does)
//   Q  ONE LDG.128 from a "quad" image (every texel stores its clamped 2x2 footprint, 16 B) + software filter
//   L  four LDG.32 from the plain linear image + software filter
//   S  source tile staged in shared memory by TMA (cp.async.bulk.tensor.2d, double-buffered, one tile per hypothesis
//      per block; best case: the whole block shares one homography) + four LDS.32 + software filter
//   H  hybrid: patch rows alternate between T and Q (two pipes in flight)
// The software filter is the texture unit's (8 fractional bits, weights derived from one rounded product, see
// oracle/cpu/apd_cpu.cpp:142-160) written without conversion instructions: float->int through FADD.RM with the
// 1.5*2^23 constant, int->float through the mantissa trick (F2I/I2F share the 16-lane/clk XU pipe with the rcp).
// On an 8-bit-valued image all five give bit-identical results (checked), so the table is apples to apples.
// Two access shapes: "coherent" = NCC of 32x8 adjacent pixels (STRONG kernels), "scattered" = every lane walks the
// 9-sample +-5 ring of its own anchor, anchors +-150 px apart (WEAK sweep).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o gpurun_out/ubench_sample tools/ubench_sample.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include <cuda.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

enum { V_TEX = 0, V_QUAD = 1, V_LIN = 2, V_SMEM = 3, V_HYB = 4 };
constexpr int TILE_W = 64, TILE_H = 24;   // floats; 6 KB per buffer

struct Args {
	cudaTextureObject_t tex;
	const float4* quad;   // [(H+1)][(W+1)] : entry (j+1, i+1) = T(j,i) T(j,i+1) T(j+1,i) T(j+1,i+1), indices clamped
	const float* lin;     // [H][W]
	int W, H;
};

__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// texel index and 8-bit fraction of one texture coordinate (u = x + 0.5 convention of tex2D)
__device__ __forceinline__ void split_coord(float u, int& i, int& A) {
	const float t = __fmaf_rn(__fadd_rn(u, -0.5f), 256.0f, 0.5f);        // exact for |u| < 2^14
	const int q = __float_as_int(__fadd_rd(t, 12582912.0f)) - 0x4B400000;   // floor(t)
	i = q >> 8; A = q & 255;
}
__device__ __forceinline__ float i2f_small(int w) { return __fadd_rn(__int_as_float(w | 0x4B000000), -8388608.0f); }
__device__ __forceinline__ float blend(int A, int B, float t00, float t10, float t01, float t11) {
	const int w11 = (A * B + 128) >> 8, w10 = A - w11, w01 = B - w11, w00 = 256 - A - B + w11;
	float acc = __fmul_rn(i2f_small(w00), t00);
	acc = __fmaf_rn(i2f_small(w10), t10, acc);
	acc = __fmaf_rn(i2f_small(w01), t01, acc);
	acc = __fmaf_rn(i2f_small(w11), t11, acc);
	return __fmul_rn(acc, 0.00390625f);
}
__device__ __forceinline__ float sample_quad(const Args& a, float u, float v) {
	int i, j, A, B; split_coord(u, i, A); split_coord(v, j, B);
	i = min(max(i, -1), a.W - 1); j = min(max(j, -1), a.H - 1);
	const float4 q = __ldg(a.quad + (size_t)(j + 1) * (a.W + 1) + (i + 1));
	return blend(A, B, q.x, q.y, q.z, q.w);
}
__device__ __forceinline__ float sample_lin(const Args& a, float u, float v) {
	int i, j, A, B; split_coord(u, i, A); split_coord(v, j, B);
	const int i0 = min(max(i, 0), a.W - 1), i1 = min(max(i + 1, 0), a.W - 1);
	const int j0 = min(max(j, 0), a.H - 1), j1 = min(max(j + 1, 0), a.H - 1);
	const float* r0 = a.lin + (size_t)j0 * a.W; const float* r1 = a.lin + (size_t)j1 * a.W;
	return blend(A, B, __ldg(r0 + i0), __ldg(r0 + i1), __ldg(r1 + i0), __ldg(r1 + i1));
}
__device__ __forceinline__ float sample_smem(const Args& a, const float* tile, int ox, int oy, float u, float v) {
	int i, j, A, B; split_coord(u, i, A); split_coord(v, j, B);
	const int li = i - ox, lj = j - oy;
	if (li >= 0 && li < TILE_W - 1 && lj >= 0 && lj < TILE_H - 1) {
		const float* p = tile + lj * TILE_W + li;
		return blend(A, B, p[0], p[1], p[TILE_W], p[TILE_W + 1]);
	}
	return sample_lin(a, u, v);   // outside the staged tile (or clamped at the image border)
}

__device__ __forceinline__ void make_h(int rep, float* H) {
	const int r = rep & 15;
	H[0] = 1.01f; H[1] = 0.004f; H[2] = 2.5f + 0.37f * r;
	H[3] = -0.003f; H[4] = 0.99f; H[5] = -1.5f + 0.21f * r;
	H[6] = 1e-6f; H[7] = -2e-6f; H[8] = 1.0f;
}

// ---- TMA plumbing (raw PTX; SASS: UTMALDG) ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
	             ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

// ---- coherent shape: one NCC (36 samples) per thread per hypothesis, 32x8 adjacent pixels per block -------------
template <int V>
__global__ void __launch_bounds__(256, 3) k_coherent(const Args a, const __grid_constant__ CUtensorMap tmap, int reps, float* out) {
	extern __shared__ __align__(1024) unsigned char smem_raw[];
	constexpr int T = 256;
	const int tid = threadIdx.y * 32 + threadIdx.x;
	// S only: [2 tiles, 128-byte aligned for TMA][2 mbarriers, padded to 128 B][(w, w r) table]
	float* tiles = reinterpret_cast<float*>(smem_raw);
	uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + 2 * TILE_W * TILE_H * 4);
	float2* wt = reinterpret_cast<float2*>(smem_raw + (V == V_SMEM ? 2 * TILE_W * TILE_H * 4 + 128 : 0)) + tid;
	const int px = 64 + blockIdx.x * 32 + threadIdx.x, py = 64 + blockIdx.y * 8 + threadIdx.y;
	for (int k = 0; k < 36; ++k) { const float w = 1.0f / (1.0f + (k % 7)); wt[k * T] = make_float2(w, w * (float)((px + k) & 15)); }
	const int bx0 = 64 + blockIdx.x * 32 - 5, by0 = 64 + blockIdx.y * 8 - 5;     // top-left of the block's reference footprint
	auto origin = [&](int rep, int& ox, int& oy) {
		float H[9]; make_h(rep, H);
		// H is close to the identity with positive scale: the top-left corner maps to the smallest coordinates
		const float z = H[8] + H[6] * bx0 + H[7] * by0;
		ox = ((int)floorf((H[2] + H[0] * bx0 + H[1] * by0) / z) - 2) & ~3;   // TMA: the innermost box coordinate has to be 16-byte aligned (a misaligned one traps as an illegal instruction)
		oy = (int)floorf((H[5] + H[3] * bx0 + H[4] * (by0 + 17)) / z) - 2;   // H[3] < 0: the right edge is lower; conservative
		oy = min(oy, (int)floorf((H[5] + H[3] * (bx0 + 41) + H[4] * by0) / z) - 2);
	};
	if (V == V_SMEM) {
		if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
		__syncthreads();
		if (tid == 0) { int ox, oy; origin(0, ox, oy); mbar_expect_tx(&bar[0], TILE_W * TILE_H * 4); tma_load_2d(tiles, &tmap, ox, oy, &bar[0]); }
	}
	float total = 0.f;
#pragma unroll 1
	for (int rep = 0; rep < reps; ++rep) {
		float H[9]; make_h(rep, H);
		int ox = 0, oy = 0;
		const float* tile = nullptr;
		if (V == V_SMEM) {
			if (tid == 0 && rep + 1 < reps) {
				int nx, ny; origin(rep + 1, nx, ny);
				mbar_expect_tx(&bar[(rep + 1) & 1], TILE_W * TILE_H * 4);
				tma_load_2d(tiles + ((rep + 1) & 1) * TILE_W * TILE_H, &tmap, nx, ny, &bar[(rep + 1) & 1]);
			}
			origin(rep, ox, oy);
			mbar_wait(&bar[rep & 1], (rep >> 1) & 1);
			tile = tiles + (rep & 1) * TILE_W * TILE_H;
		}
		float s_s = 0.f, s_ss = 0.f, s_rs = 0.f;
#pragma unroll 1
		for (int ii = 0; ii < 6; ii += 2) {
			float sv[12];
#pragma unroll
			for (int r = 0; r < 2; ++r) {
				const float xf = (float)(px - 5 + (ii + r) * 2);
				const float hx = __fmul_rn(H[0], xf), hy = __fmul_rn(H[3], xf), hz = __fmul_rn(H[6], xf);
#pragma unroll
				for (int jj = 0; jj < 6; ++jj) {
					const float yf = (float)(py - 5 + jj * 2);
					const float z = __fadd_rn(H[8], __fmaf_rn(H[7], yf, hz));
					const float x = __fadd_rn(H[2], __fmaf_rn(H[1], yf, hx));
					const float y = __fadd_rn(H[5], __fmaf_rn(H[4], yf, hy));
					const float rz = rcp_approx(z);
					const float u = __fmaf_rn(x, rz, 0.5f), v = __fmaf_rn(y, rz, 0.5f);
					float s;
					if (V == V_TEX) s = tex2D<float>(a.tex, u, v);
					else if (V == V_QUAD) s = sample_quad(a, u, v);
					else if (V == V_LIN) s = sample_lin(a, u, v);
					else if (V == V_SMEM) s = sample_smem(a, tile, ox, oy, u, v);
					else s = (r == 0) ? tex2D<float>(a.tex, u, v) : sample_quad(a, u, v);
					sv[r * 6 + jj] = s;
				}
			}
#pragma unroll
			for (int r = 0; r < 2; ++r) {
				float r_s = 0.f, r_ss = 0.f, r_rs = 0.f;
#pragma unroll
				for (int jj = 0; jj < 6; ++jj) {
					const float2 w_t = wt[((ii + r) * 6 + jj) * T];
					const float s = sv[r * 6 + jj];
					const float u = __fmul_rn(s, w_t.x);
					r_rs = __fmaf_rn(s, w_t.y, r_rs);
					r_ss = __fmaf_rn(s, u, r_ss);
					r_s = __fadd_rn(u, r_s);
				}
				s_s = __fadd_rn(r_s, s_s); s_ss = __fadd_rn(r_ss, s_ss); s_rs = __fadd_rn(r_rs, s_rs);
			}
		}
		total += s_s + 1e-3f * s_ss + 1e-3f * s_rs;
		if (V == V_SMEM) __syncthreads();   // everyone is done with this buffer before it is refilled two hypotheses later
	}
	out[(blockIdx.y * gridDim.x + blockIdx.x) * T + tid] = total;
}

// ---- scattered shape: every lane samples the +-5 ring of its own anchor (9 fetches in flight) ------------------
__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
template <int V>
__global__ void __launch_bounds__(256) k_scattered(const Args a, int reps, int spread, float* out) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	const int wpr = (a.W - 400) / 32;
	const int warp = t >> 5, lane = t & 31;
	const int wy = 200 + (warp / wpr) % (a.H - 400), wx = 200 + (warp % wpr) * 32;
	const int di[9] = {-5, -5, -5, 0, 0, 5, 5, 5, 0}, dj[9] = {-5, 0, 5, -5, 5, -5, 0, 5, 0};
	float total = 0.f;
#pragma unroll 1
	for (int rep = 0; rep < reps; ++rep) {
		float H[9]; make_h(rep, H);
		const unsigned key = hash(t * 64 + (rep & 7));          // 8 anchors per pixel, revisited every 8 hypotheses
		const int ax = wx + lane + (int)(key % (2 * spread + 1)) - spread;
		const int ay = wy + (int)((key >> 12) % (2 * spread + 1)) - spread;
		float sv[9];
#pragma unroll
		for (int q = 0; q < 9; ++q) {
			const float xf = (float)(ax + di[q]), yf = (float)(ay + dj[q]);
			const float z = __fadd_rn(H[8], __fmaf_rn(H[6], xf, __fmul_rn(H[7], yf)));
			const float x = __fadd_rn(H[2], __fmaf_rn(H[0], xf, __fmul_rn(H[1], yf)));
			const float y = __fadd_rn(H[5], __fmaf_rn(H[3], xf, __fmul_rn(H[4], yf)));
			const float rz = rcp_approx(z);
			const float u = __fmaf_rn(x, rz, 0.5f), v = __fmaf_rn(y, rz, 0.5f);
			if (V == V_TEX) sv[q] = tex2D<float>(a.tex, u, v);
			else if (V == V_QUAD) sv[q] = sample_quad(a, u, v);
			else if (V == V_LIN) sv[q] = sample_lin(a, u, v);
			else sv[q] = (q & 1) ? sample_quad(a, u, v) : tex2D<float>(a.tex, u, v);
		}
#pragma unroll
		for (int q = 0; q < 9; ++q) total = __fmaf_rn(sv[q], 1.0f / (1 + q), total);
	}
	out[t] = total;
}

static CUtensorMap make_tmap(const float* d_img, int W, int H) {
	CUtensorMap m; memset(&m, 0, sizeof(m));
	typedef CUresult (*Encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
	                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
	void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
	CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
	const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
	const cuuint64_t strides[1] = {(cuuint64_t)W * 4};     // bytes, dimension 1; must be a multiple of 16 -> W % 4 == 0
	const cuuint32_t box[2] = {TILE_W, TILE_H};
	const cuuint32_t estr[2] = {1, 1};
	CUresult r = ((Encode)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)d_img, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
	                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
	return m;
}

static std::vector<float> g_ref;
template <typename F>
static void timed(const char* shape, const char* name, double samples, size_t n_out, float* d_out, F launch) {
	cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	for (int i = 0; i < 2; ++i) launch();
	CK(cudaDeviceSynchronize());
	CK(cudaEventRecord(e0));
	for (int i = 0; i < 3; ++i) launch();
	CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
	CK(cudaGetLastError());
	float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 3;
	std::vector<float> h(n_out);
	CK(cudaMemcpy(h.data(), d_out, n_out * 4, cudaMemcpyDeviceToHost));
	size_t bad = 0;
	if (g_ref.size() == n_out) { for (size_t i = 0; i < n_out; ++i) bad += (memcmp(&h[i], &g_ref[i], 4) != 0); }
	else g_ref = h;
	printf("%-10s %-34s %8.3f ms  %8.1f Gsample/s   %s\n", shape, name, ms, samples / ms / 1e6,
	       bad ? "RESULTS DIFFER FROM tex2D" : "bit-identical to tex2D");
	if (bad) printf("           (%zu of %zu outputs differ)\n", bad, n_out);
}

int main(int argc, char** argv) {
	const char* only = argc > 1 ? argv[1] : "TQLHSX";   // which variants to run: T Q L H S (coherent), X (scattered)
	auto want = [&](char c) { return strchr(only, c) != nullptr; };
	const int W = 3112, H = 2073;   // TMA needs a 16-byte row pitch
	std::vector<float> img((size_t)W * H);
	srand(7);
	for (auto& v : img) v = (float)(rand() % 256);
	std::vector<float4> quad((size_t)(W + 1) * (H + 1));
	auto T = [&](int j, int i) { i = i < 0 ? 0 : (i >= W ? W - 1 : i); j = j < 0 ? 0 : (j >= H ? H - 1 : j); return img[(size_t)j * W + i]; };
	for (int j = -1; j < H; ++j)
		for (int i = -1; i < W; ++i) quad[(size_t)(j + 1) * (W + 1) + (i + 1)] = make_float4(T(j, i), T(j, i + 1), T(j + 1, i), T(j + 1, i + 1));
	cudaArray_t arr; cudaChannelFormatDesc cd = cudaCreateChannelDesc<float>();
	CK(cudaMallocArray(&arr, &cd, W, H));
	CK(cudaMemcpy2DToArray(arr, 0, 0, img.data(), W * 4, W * 4, H, cudaMemcpyHostToDevice));
	cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
	cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModeLinear;
	td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
	Args a; CK(cudaCreateTextureObject(&a.tex, &rd, &td, nullptr));
	float* d_lin; float4* d_quad;
	CK(cudaMalloc(&d_lin, img.size() * 4)); CK(cudaMemcpy(d_lin, img.data(), img.size() * 4, cudaMemcpyHostToDevice));
	CK(cudaMalloc(&d_quad, quad.size() * 16)); CK(cudaMemcpy(d_quad, quad.data(), quad.size() * 16, cudaMemcpyHostToDevice));
	a.quad = d_quad; a.lin = d_lin; a.W = W; a.H = H;
	const CUtensorMap tmap = make_tmap(d_lin, W, H);

	// coherent: 92 x 232 blocks of 32x8 pixels = 5.5 Mpix, 16 hypotheses each
	const dim3 grid(92, 232), block(32, 8);
	const int reps = 16;
	const size_t n_coh = (size_t)grid.x * grid.y * 256;
	float* d_out; CK(cudaMalloc(&d_out, n_coh * 4));
	const double samples = (double)n_coh * reps * 36;
	const size_t sm_wt = 36 * 256 * 8, sm_tiles = 2 * TILE_W * TILE_H * 4 + 128;
	CK(cudaFuncSetAttribute(k_coherent<V_TEX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_wt));
	CK(cudaFuncSetAttribute(k_coherent<V_QUAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_wt));
	CK(cudaFuncSetAttribute(k_coherent<V_LIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_wt));
	CK(cudaFuncSetAttribute(k_coherent<V_HYB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_wt));
	CK(cudaFuncSetAttribute(k_coherent<V_SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sm_wt + sm_tiles)));
	printf("# tools/ubench_sample.cu on B200 (variants %s): NCC inner loop with five source-sample paths (Gsample/s = bilinear samples per second)\n", only);
	timed("coherent", "T tex2D", samples, n_coh, d_out, [&] { k_coherent<V_TEX><<<grid, block, sm_wt>>>(a, tmap, reps, d_out); });   // always: the checksum reference
	if (want('Q')) timed("coherent", "Q LDG.128 quad image + sw filter", samples, n_coh, d_out, [&] { k_coherent<V_QUAD><<<grid, block, sm_wt>>>(a, tmap, reps, d_out); });
	if (want('L')) timed("coherent", "L 4 x LDG.32 + sw filter", samples, n_coh, d_out, [&] { k_coherent<V_LIN><<<grid, block, sm_wt>>>(a, tmap, reps, d_out); });
	if (want('H')) timed("coherent", "H rows alternate T / Q", samples, n_coh, d_out, [&] { k_coherent<V_HYB><<<grid, block, sm_wt>>>(a, tmap, reps, d_out); });
	if (want('S')) timed("coherent", "S TMA tile in smem + 4 x LDS", samples, n_coh, d_out, [&] { k_coherent<V_SMEM><<<grid, block, sm_wt + sm_tiles>>>(a, tmap, reps, d_out); });
	if (!want('X')) return 0;

	// scattered: one anchor ring per lane and hypothesis
	const int n_sc = 148 * 256 * 64, sreps = 64;
	float* d_out2; CK(cudaMalloc(&d_out2, (size_t)n_sc * 4));
	const double ssamples = (double)n_sc * sreps * 9;
	for (int spread : {8, 40, 150}) {
		char nm[64];
		g_ref.clear();
		snprintf(nm, sizeof nm, "T tex2D                (+-%d px)", spread);
		timed("scattered", nm, ssamples, n_sc, d_out2, [&] { k_scattered<V_TEX><<<n_sc / 256, 256>>>(a, sreps, spread, d_out2); });
		snprintf(nm, sizeof nm, "Q LDG.128 + sw filter  (+-%d px)", spread);
		timed("scattered", nm, ssamples, n_sc, d_out2, [&] { k_scattered<V_QUAD><<<n_sc / 256, 256>>>(a, sreps, spread, d_out2); });
		snprintf(nm, sizeof nm, "L 4 x LDG.32 + filter  (+-%d px)", spread);
		timed("scattered", nm, ssamples, n_sc, d_out2, [&] { k_scattered<V_LIN><<<n_sc / 256, 256>>>(a, sreps, spread, d_out2); });
		snprintf(nm, sizeof nm, "H samples alternate T/Q (+-%d px)", spread);
		timed("scattered", nm, ssamples, n_sc, d_out2, [&] { k_scattered<V_HYB><<<n_sc / 256, 256>>>(a, sreps, spread, d_out2); });
	}
	return 0;
}
