// dvp_common.cuh — shared device-side types and small math for the B200-native PatchMatch engine.
//
// Parity note (applies to every file in csrc/): results must match the reference kernels bit for bit
// on race-free stages.  The reference is built with `--use_fast_math` (CMakeLists.txt:21), so this
// library is built with the same flag and arithmetic that decides anything is written either in the
// same expression shape as the reference (so nvcc contracts/approximates identically) or with explicit
// intrinsics (__fmaf_rn / __fmul_rn / __fadd_rn, rcp/sqrt/ex2.approx.ftz via inline PTX) wherever the
// computation was restructured (hoisting, warp cooperation).  The explicit forms were read off the
// SASS of the reference build (oracle/_ref).
#pragma once
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <stdint.h>
#include "../../include/dvp_mvs.h"

#ifndef DVP_MIN
#define DVP_MIN(a, b) ((a) > (b) ? (b) : (a))  // same definition as OpenCV's MIN (cvdef.h) used by the reference
#define DVP_MAX(a, b) ((a) < (b) ? (b) : (a))
#endif

namespace dvp {

constexpr int kMaxImages = DVP_MAX_IMAGES;
constexpr int kHoistAxis = 6;                        // samples per patch axis when radius is a multiple of 5
constexpr int kHoistSamples = kHoistAxis * kHoistAxis;

// Per-source-view constants hoisted out of ComputeHomography (reference APD.cu:679-707 recomputes them on
// every NCC call).  Filled on the device by k_setup_views so FMA contraction matches the reference.
struct ViewConst {
	float R_rel[9];   // src.R * ref.R^T
	float t_rel[3];   // src.R * (C_ref - C_src)
	float sK[9];      // source intrinsics
	float sR[9];      // source rotation
	float st[3];      // source translation
	float sc[3];      // source centre
	float R_c[9];     // ref.R * src.R^T   (GenerateRandomNormal_YZL, APD.cu:541-542)
	float pad;
};

// Everything a kernel needs, passed by value as a __grid_constant__ parameter (the reference re-reads
// helper->params->x through two pointer hops inside its inner loops).
struct KArgs {
	int W, H, S;       // S = number of source views
	int N;
	dvp_params prm;
	dvp_camera ref;    // reference camera (cameras[0])
	// textures
	const cudaTextureObject_t* tex_img;    // [1+S] device array of texture objects (linear filter, clamp)
	const cudaTextureObject_t* tex_depth;  // [1+S] or null
	const float* ref_img;                  // [N] pitched-linear copy of image 0 (exact texel reads)
	const dvp_camera* cams;                // [1+S]
	const ViewConst* views;                // [S]  (index = src_idx-1)
	float4* planes;
	float4* fit_planes;
	float* costs;
	uint32_t* selected;   // padded allocation: index -W-1 .. N+W valid
	uint8_t* weak;
	int32_t* radius;
	uint8_t* view_weight; // [N][32]
	uint32_t* rng;        // SoA: 6 planes of N words {d, v0..v4}
	uint8_t* edge;
	short2* edge_neigh;   // [N][8]
	int32_t* label;
	short2* candidate;    // [N][4][8]
	short2* nearest_strong;
	uint8_t* weak_reliable;
	int32_t* neighbours_map;
	short2* neighbours;     // [weak_count][12]
	short2* label_boundary; // [weak_count][8]
	float* complex_;        // [weak_count]
	float* scratch;         // per-pixel spill area for large S (cost arrays)
	int weak_count;
	const uint8_t* edge_dist;     // [N] chessboard distance to the nearest edge pixel, capped at 255 (0 on edge pixels): lets the edge walks of K4 / K9 skip ahead
	unsigned long long* fetch_counter;   // [kFetchSlots] texture-fetch tally of the instrumented build (-DDVP_COUNT_FETCHES), else null
};

// Instrumented build only (libdvp_mvs_count.so, `make count`): every texture fetch site adds its fetches to one of
// kFetchSlots counters so that bench.py can report MEASURED fetches per launch next to the measured texture roof.
// The production build compiles this to nothing.
constexpr int kFetchSlots = 4096;
#ifdef DVP_COUNT_FETCHES
#define DVP_COUNT(a, n) do { if ((a).fetch_counter) atomicAdd((a).fetch_counter + ((blockIdx.x * 131u + blockIdx.y * 17u + threadIdx.x) & (kFetchSlots - 1)), (unsigned long long)(n)); } while (0)
#else
#define DVP_COUNT(a, n) do { } while (0)
#endif

// ---- approximate MUFU ops exactly as the reference build emits them -----------------------------------
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// ---- RNG: cuRAND XORWOW, stored as 6 SoA planes --------------------------------------------------------
struct Rng {
	curandStateXORWOW_t st;
	__device__ __forceinline__ void load(const uint32_t* rng, int N, int idx) {
		st.d = rng[idx];
		st.v[0] = rng[N + idx]; st.v[1] = rng[2 * N + idx]; st.v[2] = rng[3 * N + idx];
		st.v[3] = rng[4 * N + idx]; st.v[4] = rng[5 * N + idx];
		st.boxmuller_flag = 0; st.boxmuller_flag_double = 0; st.boxmuller_extra = 0.f; st.boxmuller_extra_double = 0.0;
	}
	__device__ __forceinline__ void store(uint32_t* rng, int N, int idx) const {
		rng[idx] = st.d;
		rng[N + idx] = st.v[0]; rng[2 * N + idx] = st.v[1]; rng[3 * N + idx] = st.v[2];
		rng[4 * N + idx] = st.v[3]; rng[5 * N + idx] = st.v[4];
	}
	// the six live words, for callers that draw ahead speculatively and roll back (K4's direction search)
	struct Snap { uint32_t d, v0, v1, v2, v3, v4; };
	__device__ __forceinline__ Snap snap() const { return Snap{st.d, st.v[0], st.v[1], st.v[2], st.v[3], st.v[4]}; }
	__device__ __forceinline__ void restore(const Snap& s) { st.d = s.d; st.v[0] = s.v0; st.v[1] = s.v1; st.v[2] = s.v2; st.v[3] = s.v3; st.v[4] = s.v4; }
	__device__ __forceinline__ float uniform() { return curand_uniform(&st); }
	__device__ __forceinline__ unsigned int next() { return curand(&st); }
};

// ---- small geometry helpers; expression shapes follow the reference so contraction is identical ------
// reference APD.cu:372-377
__device__ __forceinline__ void get_3d_point(const dvp_camera& cam, int px, int py, float depth, float* X) {
	X[0] = depth * (px - cam.K[2]) / cam.K[0];
	X[1] = depth * (py - cam.K[5]) / cam.K[4];
	X[2] = depth;
}
// reference APD.cu:386-398
__device__ __forceinline__ float4 get_view_direction(const dvp_camera& cam, int px, int py, float depth) {
	float X[3];
	get_3d_point(cam, px, py, depth, X);
	float norm = sqrt(X[0] * X[0] + X[1] * X[1] + X[2] * X[2]);
	float4 v;
	v.x = X[0] / norm; v.y = X[1] / norm; v.z = X[2] / norm; v.w = 0;
	return v;
}
// reference APD.cu:400-405
__device__ __forceinline__ float get_distance2origin(const dvp_camera& cam, int px, int py, float depth, const float4 n) {
	float X[3];
	get_3d_point(cam, px, py, depth, X);
	return -(n.x * X[0] + n.y * X[1] + n.z * X[2]);
}
// reference APD.cu:419-422
__device__ __forceinline__ float depth_from_plane(const dvp_camera& cam, const float4 pl, int px, int py) {
	return -pl.w * cam.K[0] / ((px - cam.K[2]) * pl.x + (cam.K[0] / cam.K[4]) * (py - cam.K[5]) * pl.y + cam.K[0] * pl.z);
}
// reference APD.cu:750-758 (camera -> world: R^T n)
__device__ __forceinline__ float4 normal_to_world(const dvp_camera& cam, float4 pl) {
	float4 o;
	o.x = cam.R[0] * pl.x + cam.R[3] * pl.y + cam.R[6] * pl.z;
	o.y = cam.R[1] * pl.x + cam.R[4] * pl.y + cam.R[7] * pl.z;
	o.z = cam.R[2] * pl.x + cam.R[5] * pl.y + cam.R[8] * pl.z;
	o.w = pl.w;
	return o;
}
// reference APD.cu:760-768 (world -> reference camera: R n)
__device__ __forceinline__ float4 normal_to_refcam(const dvp_camera& cam, float4 pl) {
	float4 o;
	o.x = cam.R[0] * pl.x + cam.R[1] * pl.y + cam.R[2] * pl.z;
	o.y = cam.R[3] * pl.x + cam.R[4] * pl.y + cam.R[5] * pl.z;
	o.z = cam.R[6] * pl.x + cam.R[7] * pl.y + cam.R[8] * pl.z;
	o.w = pl.w;
	return o;
}
// reference APD.cu:331-338
__device__ __forceinline__ void normalize3(float4* v) {
	const float n2 = v->x * v->x + v->y * v->y + v->z * v->z;
	const float inv = rsqrtf(n2);
	v->x *= inv; v->y *= inv; v->z *= inv;
}
// reference APD.cu:467-487
__device__ __forceinline__ float3 point_to_world(float x, float y, float depth, const float* K, const float* R, const float* c) {
	float3 pX, t;
	pX.x = depth * (x - K[2]) / K[0];
	pX.y = depth * (y - K[5]) / K[4];
	pX.z = depth;
	t.x = R[0] * pX.x + R[3] * pX.y + R[6] * pX.z;
	t.y = R[1] * pX.x + R[4] * pX.y + R[7] * pX.z;
	t.z = R[2] * pX.x + R[5] * pX.y + R[8] * pX.z;
	pX.x = t.x + c[0]; pX.y = t.y + c[1]; pX.z = t.z + c[2];
	return pX;
}
// reference APD.cu:489-499
__device__ __forceinline__ void project_on_camera(const float3 P, const float* K, const float* R, const float* t, float2& pt, float& depth) {
	float3 tmp;
	tmp.x = R[0] * P.x + R[1] * P.y + R[2] * P.z + t[0];
	tmp.y = R[3] * P.x + R[4] * P.y + R[5] * P.z + t[1];
	tmp.z = R[6] * P.x + R[7] * P.y + R[8] * P.z + t[2];
	depth = K[6] * tmp.x + K[7] * tmp.y + K[8] * tmp.z;
	pt.x = (K[0] * tmp.x + K[1] * tmp.y + K[2] * tmp.z) / depth;
	pt.y = (K[3] * tmp.x + K[4] * tmp.y + K[5] * tmp.z) / depth;
}

// x % m for a divisor fixed over a loop, exact for every 32-bit x and m >= 1: q = floor(x * floor((2^32 - 1) / m) / 2^32)
// underestimates floor(x / m) by at most 2 (x M / 2^32 > x / m - 1 - x / 2^32), so two conditional subtractions finish it
// (tests/test_host_logic.py checks the arithmetic against % on adversarial and random operands).
struct FastMod {
	unsigned int m, M;
	__host__ __device__ explicit FastMod(unsigned int m_) : m(m_), M(0xFFFFFFFFu / m_) {}
	__device__ __forceinline__ unsigned int mod(unsigned int x) const {
		unsigned int r = x - __umulhi(x, M) * m;
		if (r >= m) r -= m;
		if (r >= m) r -= m;
		return r;
	}
};

__device__ __forceinline__ int is_set(uint32_t v, int n) { return (v >> n) & 1; }
// reference APD.cu:186-189 — clears bit n AND every lower bit (bug B1, reproduced on purpose)
__device__ __forceinline__ void unset_bit_ref(uint32_t* v, int n) { (*v) &= (uint32_t)(0xFFFFFFFEu << n); }

// ---- thread -> pixel maps of the NCC kernels ----
// The texture unit retires the four lanes of a quad in one clock only when their four bilinear footprints fall in a
// small texel window (profiles/r01_tex_coherence_ubench.txt).  With one image row per warp a quad is 4 x 1 pixels —
// on a checkerboard colour a 4 x 2 zigzag — whose footprints span 3 s + 2 texels at source scale s.  These maps give
// every quad a compact pixel set instead; the work per pixel is unchanged, so results are bit-identical (same SHA-1
// over every output buffer of a pass, tools/ab_variants.py).  Measured on B200 (profiles/r02_quad_maps.txt):
//   full-grid kernels, 2 x 2 quads: fused K15+K16 51.1 -> 49.0 ms at 3111x2073, 205.2 -> 197.1 ms at 6221x4146; K6 1.25 -> 1.17
//     (a 16 x 2 and an 8 x 4 tile per warp measure the same)                                          -> default 1
//   checkerboard kernels: both compact maps LOSE 4 % (K7 113.7 -> 118.9 ms at 6221x4146): the scratch rows a warp
//     writes and reads stop being one 128-byte line                                                   -> default 0
#ifndef DVP_QUAD_FULL
#define DVP_QUAD_FULL 1     // full-grid kernels (K6, K15+K16): 0 = a row of 32 pixels per warp, 1 = a 16 x 2 tile per warp (quads 2 x 2), 2 = an 8 x 4 tile
#endif
#ifndef DVP_QUAD_SWEEP
#define DVP_QUAD_SWEEP 0    // checkerboard kernels (K7/K8): 0 = zigzag row, 1 = 2 x 2 in (x, y / 2), 2 = diamonds (3 x 3 pixels per quad)
#endif
// 32 x (T / 32) pixel block, T = 128 or 256 threads: warp w owns the 16 x 2 tile (w & 1, w >> 1), lane l the pixel
// (2 (l >> 2) + (l & 1), (l >> 1) & 1) of it.
__device__ __forceinline__ void quad_tile_xy(int tid, int& lx, int& ly) {
	const int lane = tid & 31, w = tid >> 5;
#if DVP_QUAD_FULL == 2   // 8 x 4 tile per warp (quads 4 x 2), warps 4 x 2 in the block
	const int q = lane >> 2;
	lx = (w & 3) * 8 + 2 * (q & 3) + (lane & 1);
	ly = (w >> 2) * 4 + 2 * (q >> 2) + ((lane >> 1) & 1);
#else
	lx = (w & 1) * 16 + 2 * (lane >> 2) + (lane & 1);
	ly = (w >> 1) * 2 + ((lane >> 1) & 1);
#endif
}
__device__ __forceinline__ void full_grid_pixel(int& x, int& y) {
#if DVP_QUAD_FULL
	int lx, ly; quad_tile_xy(threadIdx.y * blockDim.x + threadIdx.x, lx, ly);
	x = blockIdx.x * 32 + lx; y = blockIdx.y * (blockDim.x * blockDim.y / 32) + ly;
#else
	x = blockIdx.x * blockDim.x + threadIdx.x; y = blockIdx.y * blockDim.y + threadIdx.y;
#endif
}
// Pixel of colour `red` ((x + y) & 1 == red) owned by this thread of a checkerboard launch; y = 2 yy + ((x & 1) ^ red).
// Mode 2 tiles the colour's lattice with diamonds {(xa, ya), (xa + 1, ya - 1), (xa + 1, ya + 1), (xa + 2, ya)} laid like
// bricks: diamond row r sits at ya = 2 r - 2 + red (row 0 reaches y = 0 through its lower pixel) and starts at xa = 4 k - 2 (r & 1); every pixel of the colour belongs to
// exactly one diamond (tests/test_host_logic.py enumerates it).  x or y may come out negative: callers bounds-check.
__device__ __forceinline__ void sweep_pixel(int red, int& x, int& y) {
	const int tid = threadIdx.y * blockDim.x + threadIdx.x;
#if DVP_QUAD_SWEEP == 2
	const int rows = blockDim.x * blockDim.y / 32;
	const int k = blockIdx.x * 8 + ((tid >> 2) & 7), r = blockIdx.y * rows + (tid >> 5), q = tid & 3;
	x = 4 * k - 2 * (r & 1) + ((q + 1) >> 1);
	y = 2 * r - 2 + red + ((q == 1) ? -1 : (q == 2) ? 1 : 0);
#elif DVP_QUAD_SWEEP == 1
	int lx, ly; quad_tile_xy(tid, lx, ly);
	x = blockIdx.x * 32 + lx;
	y = 2 * (blockIdx.y * (blockDim.x * blockDim.y / 32) + ly) + ((x & 1) ^ red);
#else
	x = blockIdx.x * blockDim.x + threadIdx.x;
	y = 2 * (blockIdx.y * blockDim.y + threadIdx.y) + ((x & 1) ^ red);
#endif
}

}  // namespace dvp
