// dvp_launch.h — host-callable launchers of every stage kernel (one per reference kernel, see dvp_stage).
#pragma once
#include "dvp_common.cuh"

namespace dvp {

#ifndef DVP_SWEEP_THREADS
#define DVP_SWEEP_THREADS 128
#endif
#ifndef DVP_SWEEP_MIN_BLOCKS
#define DVP_SWEEP_MIN_BLOCKS 3
#endif
#ifndef DVP_SWEEP_BX
#define DVP_SWEEP_BX 32
#endif
constexpr int kSweepBlockX = DVP_SWEEP_BX;        // block width in pixels; height = threads / width row pairs (2 image rows each)
constexpr int kSweepThreads = DVP_SWEEP_THREADS;  // threads per block of the propagation sweep
constexpr int kSweepMinBlocks = DVP_SWEEP_MIN_BLOCKS;  // resident blocks per SM the register budget is cut for
#ifndef DVP_SWEEP_RW
#define DVP_SWEEP_RW 0
#endif
constexpr bool kSweepRW = DVP_SWEEP_RW != 0;  // sweep keeps only the reference samples in shared memory and re-evaluates the bilateral weight per use
#ifndef DVP_SWEEP_RB
#define DVP_SWEEP_RB 3
#endif
constexpr int kSweepRB = DVP_SWEEP_RB;         // patch rows of texture fetches in flight per thread in the sweep (all 36 samples)
#ifndef DVP_WEAK_THREADS
#define DVP_WEAK_THREADS 256
#endif
constexpr int kWeakThreads = DVP_WEAK_THREADS;   // WEAK sweep block = one tile of the per-colour list: 16 x (threads / 8) pixels
#ifndef DVP_K4_THREADS
#define DVP_K4_THREADS 64
#endif
#ifdef DVP_WEAK_NO_SCORE_KERNEL
constexpr int kWeakSweepKernels = 1;
#else
constexpr int kWeakSweepKernels = 2;             // kernels per K10 / K11 launch: k_weak_score + k_weak_sweep
#endif
constexpr int kK4Threads = DVP_K4_THREADS;       // K4 block = one (threads / 8) x 8 pixel tile of the WEAK list
#ifndef DVP_WIDE_RB
#define DVP_WIDE_RB 2
#endif
constexpr int kWideRB = DVP_WIDE_RB;          // ... in the 256-thread, 24-warp/SM kernels (K6, K15, K16)

// shared memory: 36 (w, w*r) pairs per thread
inline size_t patch_smem_bytes(int threads) { return (size_t)kHoistSamples * threads * sizeof(float2); }
// + 9 cost vectors of S floats + 8 16-bit ladder offsets per thread
inline size_t sweep_table_bytes(int threads) { return kSweepRW ? patch_smem_bytes(threads) / 2 : patch_smem_bytes(threads); }
inline size_t sweep_smem_bytes(int threads, int S) { return sweep_table_bytes(threads) + (size_t)(9 * S + 4) * threads * 4; }

cudaError_t configure_strong_kernels(int S);
cudaError_t configure_weak_kernels(int S);

cudaError_t launch_fill_sd_table(cudaStream_t st);
cudaError_t launch_setup_views(const dvp_camera* cams, ViewConst* views, int S, cudaStream_t st);
cudaError_t launch_init_rng(const KArgs& a, unsigned long long seed, cudaStream_t st);          // K1
cudaError_t launch_edge_inform(const KArgs& a, bool with_candidates, cudaStream_t st);                                 // K2
cudaError_t launch_nearest_strong(const KArgs& a, short* next_right, short* next_down, cudaStream_t st);  // K3
cudaError_t launch_gen_neighbours(const KArgs& a, const int* weak_list, uint8_t* anchor_flag, cudaStream_t st);        // K4
cudaError_t launch_neighbour_update(const KArgs& a, cudaStream_t st);                            // K5
cudaError_t launch_random_init(const KArgs& a, cudaStream_t st);                                 // K6
size_t sweep_scratch_bytes(int W, int H, int S);                                                 // candidate costs / winners between the two kernels of a sweep
cudaError_t launch_strong_sweep(const KArgs& a, int iter, int red, void* scratch, cudaStream_t st);   // K7 / K8
// parity instrumentation: direction 4's candidate forced to ladder offset m, planes from snapshots (see k_strong_sweep)
struct D4Force {
	int m = 0;                                               // ladder offset of direction 4's candidate: pixel (x - 5 - m, y - 5 - m)
	const float4* before = nullptr; const float4* after = nullptr;   // the candidate's plane before / after its own update
	unsigned long long ncc_masks = 0;                        // 4 bits per source view (views 0..15): bit c set = component c of the plane scored for
	                                                         // that view comes from `after` (the reference re-reads the plane, 32 bits at a time, for every view)
	unsigned dep_mask = 0, acc_mask = 0;                     // the same for the depth-test read and for the copy at acceptance
	const int* pixel_list = nullptr; int list_count = 0;     // null: the whole half grid of the colour
	float4* out_planes = nullptr; float* out_costs = nullptr; uint32_t* out_selected = nullptr; uint8_t* out_view_weight = nullptr; uint32_t* out_rng = nullptr;
};
struct RaceExpected { const float4* planes; const float* costs; const uint32_t* selected; const uint8_t* view_weight; const uint32_t* rand; };
cudaError_t launch_strong_sweep_forced(const KArgs& a, int iter, int red, const D4Force& force, cudaStream_t st);
cudaError_t launch_explain_init(const KArgs& a, int red, const RaceExpected& e, uint8_t* explained, cudaStream_t st);
cudaError_t launch_explain_compare(const KArgs& a, int red, const D4Force& f, const RaceExpected& e, uint8_t* explained, cudaStream_t st);
cudaError_t launch_explain_collect(const KArgs& a, int red, const uint8_t* explained, int* list, int* count, int cap, cudaStream_t st);
cudaError_t launch_ransac_fit(const KArgs& a, const int* weak_list, cudaStream_t st);                                  // K9
cudaError_t launch_weak_sweep(const KArgs& a, const int* colour_list, int count, int iter, int red, void* scratch, cudaStream_t st);  // K10 / K11
cudaError_t launch_depth_normal(const KArgs& a, cudaStream_t st);                                // K12
cudaError_t launch_filter(const KArgs& a, int red, cudaStream_t st);                             // K13 / K14
cudaError_t launch_depth_to_weak(const KArgs& a, cudaStream_t st);                               // K15
cudaError_t launch_local_refine(const KArgs& a, cudaStream_t st);                                // K16
cudaError_t launch_depth_to_weak_refine(const KArgs& a, cudaStream_t st);                        // K15 + K16 fused (dvp_run)

cudaError_t launch_fill_i32(int32_t* dst, int32_t v, int n, cudaStream_t st);
cudaError_t launch_edge_distance(const uint8_t* edge, int W, int H, int* sat, uint8_t* dist, cudaStream_t st);   // chessboard distance to the nearest edge, capped at 255
// WEAK-pixel indexing (device-side replacement of the host loop APD.cpp:1182-1193)
constexpr int kWeakScanBlock = 1024;
cudaError_t launch_weak_count(const uint8_t* weak, int n, int W, int colour, int yy_limit, int* block_sums, int* total, cudaStream_t st);
cudaError_t launch_weak_index(const uint8_t* weak, int n, int W, int colour, int yy_limit, const int* block_offsets, int* nmap, int* weak_list, cudaStream_t st);
size_t tile_order_temp_bytes(int count);
cudaError_t launch_tile_order(int* list, int count, int W, int H, int tile_w, int tile_h, int* keys_in, int* keys_out, int* vals_out,
                              void* temp, size_t temp_bytes, cudaStream_t st);
cudaError_t launch_reset_unknown_radius(const uint8_t* weak, int32_t* radius, int32_t strong_radius, int n, cudaStream_t st);

// post-pass on the resident maps (main.cpp:297-363): dvp_kernels_post.cu
cudaError_t launch_invalidate_depth(const KArgs& a, cudaStream_t st);
cudaError_t launch_rescale(const void* src, int sw, int sh, void* dst, int dw, int dh, int elem_bytes, cudaStream_t st);
cudaError_t launch_extract_depth(const float4* planes, float* depth, int n, cudaStream_t st);
size_t edge_scratch_bytes(int W, int H);
cudaError_t launch_edge_segment(const uint8_t* d_img, int W, int H, uint8_t* d_edge, void* scratch, int** d_thr, cudaStream_t st);
cudaError_t launch_edge_to_u8(const float* d_img, int n, uint8_t* d_out, cudaStream_t st);
cudaError_t launch_border_cleanup(uint8_t* d_img, int W, int H, cudaStream_t st);
cudaError_t launch_restore_visibility(const KArgs& a, int scale_size, int* parent, int* count, cudaStream_t st);

// image preparation (dvp_kernels_image.cu): cv::resize(INTER_LINEAR) of float images, 8-bit -> float
cudaError_t launch_resize_linear_f32(const float* src, int sw, int sh, float* dst, int dw, int dh, cudaStream_t st);
cudaError_t launch_u8_to_f32(const uint8_t* src, size_t n, float* dst, cudaStream_t st);

// canonical RNG exchange format <-> SoA planes
cudaError_t launch_rng_export(const KArgs& a, uint32_t* dst_aos, cudaStream_t st);
cudaError_t launch_rng_import(const KArgs& a, const uint32_t* src_aos, cudaStream_t st);

}  // namespace dvp
