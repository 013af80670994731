// dvp_api.cu — host side of the flat C ABI declared in include/dvp_mvs.h: context, device memory and
// texture set-up (replaces APD::CudaSpaceInitialization / SetDataPassHelperInCuda / ~APD, reference
// APD.cpp:989-1043, 1497-1613, 1670-1704) and the RunPatchMatch launch sequence (reference APD.cu:4406-4532).
//
// Differences from the reference's host code, none of which changes results:
//  * one context = one device + one non-blocking stream; no cudaDeviceSynchronize between kernels
//    (the reference synchronises the whole device after each of its 11 + 5*iters launches);
//  * parameters travel as a __grid_constant__ kernel argument instead of a device-resident pointer bundle;
//  * buffers are allocated once per context and reused across uploads (multi-pass / multi-view reuse);
//  * `selected_views` carries one zeroed padding row on each side so the reference's out-of-bounds
//    4-neighbour reads at the image border are defined (they read 0).
#include "dvp_common.cuh"
#include "dvp_launch.h"
#include <cstdio>
#include <cstring>
#include <vector>
#include <algorithm>
#include <new>

namespace dvp { cudaError_t launch_edge_inform(const KArgs& a, bool with_candidates, cudaStream_t st); }

using namespace dvp;

struct dvp_ctx {
	int device = 0;
	int W = 0, H = 0, S = 0, N = 0;
	dvp_params prm;
	cudaStream_t stream = nullptr;
	// dvp_upload_overlapped: the large maps travel on a second stream while K1..K5 already run
	cudaStream_t copy_stream = nullptr;
	cudaEvent_t ev_main_idle = nullptr, ev_planes = nullptr, ev_maps = nullptr;
	bool copies_pending = false;
	bool uploaded = false;
	unsigned long long seed = 0;
	int weak_count = 0;
	int weak_capacity = 0;
	int last_err = 0;
	// textures
	cudaArray_t img_arr[DVP_MAX_IMAGES] = {nullptr};
	cudaArray_t dep_arr[DVP_MAX_IMAGES] = {nullptr};
	cudaTextureObject_t img_tex[DVP_MAX_IMAGES] = {0};
	cudaTextureObject_t dep_tex[DVP_MAX_IMAGES] = {0};
	cudaTextureObject_t* d_img_tex = nullptr;
	cudaTextureObject_t* d_dep_tex = nullptr;
	float* ref_img = nullptr;
	dvp_camera* cams = nullptr;
	dvp_camera ref_cam;
	ViewConst* views = nullptr;
	float4* planes = nullptr;
	float4* fit_planes = nullptr;
	float* costs = nullptr;
	uint32_t* selected_alloc = nullptr;
	uint32_t* selected = nullptr;
	uint8_t* weak = nullptr;
	int32_t* radius = nullptr;
	uint8_t* view_weight = nullptr;
	uint32_t* rng = nullptr;
	uint8_t* edge = nullptr;
	int* edge_sat = nullptr;          // summed-area table of `edge` and the chessboard distance map built from it (edge walks of K4 / K9)
	uint8_t* edge_dist = nullptr;
	short2* edge_neigh = nullptr;
	int32_t* label = nullptr;
	short2* candidate = nullptr;
	short2* nearest_strong = nullptr;
	uint8_t* weak_reliable = nullptr;
	uint8_t* anchor_flag = nullptr;                 // [N] pixels named by some WEAK pixel's anchor list (dvp_run: K2's candidate records are evaluated for these only)
	bool lazy_candidates = false;                   // inside dvp_run
	bool last_run_lazy = false;                     // ... and whether the last dvp_run was (launch counts of dvp_last_run_times)
	int32_t* neighbours_map = nullptr;
	short2* neighbours = nullptr;
	short2* label_boundary = nullptr;
	float* complex_ = nullptr;
	int* weak_list = nullptr;   // pixel index of every WEAK pixel, in neighbours_map order
	short* next_right = nullptr; // K3 pointer maps: next STRONG pixel in the row / column
	short* next_down = nullptr;
	int* scan_blocks = nullptr; // per-1024-pixel WEAK counts / offsets
	int* scan_total = nullptr;  // [3]: all WEAK, black WEAK, red WEAK
	int* colour_list[2] = {nullptr, nullptr};  // WEAK pixels of one checkerboard colour (dense warps in the weak sweep)
	int colour_count[2] = {0, 0};
	int* scan_blocks_c[2] = {nullptr, nullptr};
	int* sort_keys[2] = {nullptr, nullptr};   // tile reordering of the WEAK lists (sized by weak_capacity)
	int* sort_vals = nullptr;
	void* sort_temp = nullptr;
	size_t sort_temp_bytes = 0;
	void* sweep_scratch = nullptr;                  // K7 / K8: candidate costs and winners between k_sweep_score and k_sweep_update (allocated on first use)
	unsigned long long* fetch_counter = nullptr;   // instrumented build only (DVP_COUNT_FETCHES)
	int* vis_parent = nullptr;  // union-find links / region sizes of dvp_restore_visibility (allocated on first use)
	int* vis_count = nullptr;
	// host staging
	std::vector<int32_t> h_i32;
	std::vector<uint8_t> h_u8;
	std::vector<int32_t> h_list;
	// timing
	static const int kMaxLaunch = 16 + 5 * 64;
	cudaEvent_t ev[2 * (16 + 5 * 64) + 2] = {nullptr};
	int ev_stage[16 + 5 * 64] = {0};
	int n_timed = 0;
	bool timed_valid = false;
};

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->last_err = (int)e_; \
	fprintf(stderr, "[dvp] %s failed: %s (%s:%d)\n", #call, cudaGetErrorString(e_), __FILE__, __LINE__); return DVP_ERR_CUDA; } } while (0)

namespace {

template <typename T> cudaError_t zalloc(T** p, size_t count) {
	if (count == 0) count = 1;
	cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
	if (e != cudaSuccess) return e;
	return cudaMemset(*p, 0, count * sizeof(T));
}

KArgs make_args(const dvp_ctx* c) {
	KArgs a;
	memset(&a, 0, sizeof(a));
	a.W = c->W; a.H = c->H; a.S = c->S; a.N = c->N;
	a.prm = c->prm; a.ref = c->ref_cam;
	a.tex_img = c->d_img_tex; a.tex_depth = c->d_dep_tex; a.ref_img = c->ref_img;
	a.cams = c->cams; a.views = c->views;
	a.planes = c->planes; a.fit_planes = c->fit_planes; a.costs = c->costs; a.selected = c->selected;
	a.weak = c->weak; a.radius = c->radius; a.view_weight = c->view_weight; a.rng = c->rng;
	a.edge = c->edge; a.edge_neigh = c->edge_neigh; a.label = c->label; a.candidate = c->candidate;
	a.nearest_strong = c->nearest_strong; a.weak_reliable = c->weak_reliable; a.neighbours_map = c->neighbours_map;
	a.neighbours = c->neighbours; a.label_boundary = c->label_boundary; a.complex_ = c->complex_;
	a.scratch = nullptr; a.weak_count = c->weak_count; a.fetch_counter = c->fetch_counter;
	a.edge_dist = c->edge_dist;
	return a;
}

int fill_texture(dvp_ctx* ctx, cudaArray_t arr, const float* src, cudaMemcpyKind kind, cudaStream_t stream) {
	CK(cudaMemcpy2DToArrayAsync(arr, 0, 0, src, ctx->W * sizeof(float), ctx->W * sizeof(float), ctx->H, kind, stream));
	return DVP_OK;
}

// array + texture object, created once per context and image slot (no data yet)
int ensure_texture(dvp_ctx* ctx, cudaArray_t* arr, cudaTextureObject_t* tex) {
	cudaChannelFormatDesc desc = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
	if (!*arr) CK(cudaMallocArray(arr, &desc, ctx->W, ctx->H));
	if (!*tex) {
		// hardware bilinear filter, unnormalised coordinates, clamp addressing: what the reference's texture
		// objects do in effect (it asks for wrap, which degrades to clamp with unnormalised coordinates)
		cudaResourceDesc res; memset(&res, 0, sizeof(res));
		res.resType = cudaResourceTypeArray; res.res.array.array = *arr;
		cudaTextureDesc td; memset(&td, 0, sizeof(td));
		td.addressMode[0] = cudaAddressModeClamp; td.addressMode[1] = cudaAddressModeClamp;
		td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
		CK(cudaCreateTextureObject(tex, &res, &td, nullptr));
	}
	return DVP_OK;
}

struct BufDesc { void* ptr; size_t bytes; };
BufDesc buf_desc(dvp_ctx* c, int id) {
	const size_t N = (size_t)c->N, wc = (size_t)c->weak_count;
	switch (id) {
	case DVP_BUF_PLANES: return {c->planes, N * 16};
	case DVP_BUF_COSTS: return {c->costs, N * 4};
	case DVP_BUF_SELECTED: return {c->selected, N * 4};
	case DVP_BUF_WEAK: return {c->weak, N};
	case DVP_BUF_RADIUS: return {c->radius, N * 4};
	case DVP_BUF_VIEW_WEIGHT: return {c->view_weight, N * DVP_MAX_IMAGES};
	case DVP_BUF_RAND: return {c->rng, N * 24};
	case DVP_BUF_FIT_PLANES: return {c->fit_planes, N * 16};
	case DVP_BUF_EDGE_NEIGH: return {c->edge_neigh, N * DVP_EDGE_NEIGH_NUM * 4};
	case DVP_BUF_CANDIDATE: return {c->candidate, N * DVP_LAB_BOUNDARY_NUM * DVP_NUM_IMAGES * 4};
	case DVP_BUF_NEAREST_STRONG: return {c->nearest_strong, N * 4};
	case DVP_BUF_WEAK_RELIABLE: return {c->weak_reliable, N};
	case DVP_BUF_NEIGHBOURS_MAP: return {c->neighbours_map, N * 4};
	case DVP_BUF_NEIGHBOURS: return {c->neighbours, wc * DVP_NEIGHBOUR_NUM * 4};
	case DVP_BUF_LABEL_BOUNDARY: return {c->label_boundary, wc * DVP_LAB_BOUNDARY_NUM * 4};
	case DVP_BUF_COMPLEX: return {c->complex_, wc * 4};
	default: return {nullptr, 0};
	}
}

// kernels one stage launches (K2 and K3 are several kernels; WEAK-only work is skipped when no pixel is WEAK)
int stage_kernel_count(const dvp_ctx* c, int stage) {
	const bool weak = c->weak_count > 0;
	switch (stage) {
	case DVP_K2_GEN_EDGE_INFORM: return ((weak && !c->lazy_candidates) ? 1 : 0) + (c->prm.use_edge ? 1 : 0) + ((c->prm.use_label && weak) ? 1 : 0) + 1;
	case DVP_K3_FIND_NEAREST_STRONG: return (weak ? 1 : 0) + 1;
	case DVP_K4_GEN_NEIGHBOURS: return weak ? (c->lazy_candidates ? 3 : 1) : 0;   // + k_mark_anchors + k_candidate for the anchors
	case DVP_K7_BLACK_STRONG: case DVP_K8_RED_STRONG: return 2;   // k_sweep_score + k_sweep_update
	case DVP_K9_RANSAC_FIT_PLANE: return 1 + (weak ? 1 : 0);
	case DVP_K10_BLACK_WEAK: return c->colour_count[0] > 0 ? kWeakSweepKernels : 0;   // k_weak_score + k_weak_sweep
	case DVP_K11_RED_WEAK: return c->colour_count[1] > 0 ? kWeakSweepKernels : 0;
	default: return 1;
	}
}

cudaError_t launch_stage(dvp_ctx* c, const KArgs& a, int stage, int iter) {
	cudaStream_t st = c->stream;
	switch (stage) {
	case DVP_K1_INIT_RANDOM_STATES: return launch_init_rng(a, c->seed, st);
	case DVP_K2_GEN_EDGE_INFORM: return launch_edge_inform(a, !c->lazy_candidates, st);
	case DVP_K3_FIND_NEAREST_STRONG: return launch_nearest_strong(a, c->next_right, c->next_down, st);
	case DVP_K4_GEN_NEIGHBOURS: return launch_gen_neighbours(a, c->weak_list, c->lazy_candidates ? c->anchor_flag : nullptr, st);
	case DVP_K5_NEIGHBOUR_UPDATE: return launch_neighbour_update(a, st);
	case DVP_K6_RANDOM_INITIALIZATION: return launch_random_init(a, st);
	case DVP_K7_BLACK_STRONG:
	case DVP_K8_RED_STRONG: {
		if (!c->sweep_scratch) { cudaError_t e = cudaMalloc(&c->sweep_scratch, sweep_scratch_bytes(c->W, c->H, c->S)); if (e != cudaSuccess) return e; }
		return launch_strong_sweep(a, iter, stage == DVP_K8_RED_STRONG ? 1 : 0, c->sweep_scratch, st);
	}
	case DVP_K9_RANSAC_FIT_PLANE: return launch_ransac_fit(a, c->weak_list, st);
	case DVP_K10_BLACK_WEAK:
	case DVP_K11_RED_WEAK: {   // the anchors' hypotheses are scored into the (idle) K7 / K8 scratch area first
		const int col = stage == DVP_K11_RED_WEAK ? 1 : 0;
		if (c->colour_count[col] > 0 && !c->sweep_scratch) { cudaError_t e = cudaMalloc(&c->sweep_scratch, sweep_scratch_bytes(c->W, c->H, c->S)); if (e != cudaSuccess) return e; }
		return launch_weak_sweep(a, c->colour_list[col], c->colour_count[col], iter, col, c->sweep_scratch, st);
	}
	case DVP_K12_DEPTH_NORMAL: return launch_depth_normal(a, st);
	case DVP_K13_BLACK_FILTER: return launch_filter(a, 0, st);
	case DVP_K14_RED_FILTER: return launch_filter(a, 1, st);
	case DVP_K15_DEPTH_TO_WEAK: return launch_depth_to_weak(a, st);
	case DVP_K16_LOCAL_REFINE: return launch_local_refine(a, st);
	case DVP_K15_K16_FUSED: return launch_depth_to_weak_refine(a, st);
	default: return cudaErrorInvalidValue;
	}
}

// One upload, with every image / depth map addressed on its own (the scene driver keeps them in separate
// per-view buffers) and the cameras possibly on the host while the maps are on the device.
struct UploadSrc {
	const float* images[DVP_MAX_IMAGES] = {nullptr};
	const float* depths[DVP_MAX_IMAGES] = {nullptr};
	bool have_depths = false;
	const dvp_camera* cameras = nullptr;
	bool cameras_on_device = false;
	const float* planes = nullptr;
	const uint32_t* selected_views = nullptr;
	const uint8_t* weak_info = nullptr;
	const uint8_t* edge = nullptr;
	const int32_t* label = nullptr;
	const int32_t* radius = nullptr;
	uint64_t seed = 0;
};

int upload_parts(dvp_ctx* ctx, const UploadSrc* in, const dvp_params* params, bool from_device, bool overlap = false);

// everything on the context stream is ordered after the copies of an overlapped upload
int join_copies(dvp_ctx* ctx) {
	if (ctx->copies_pending) {
		CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_maps, 0));
		ctx->copies_pending = false;
	}
	return DVP_OK;
}

int upload_common(dvp_ctx* ctx, const dvp_inputs* din, const dvp_params* params, bool from_device, bool overlap = false) {
	if (!ctx || !din || !din->images || !din->cameras || !din->planes) return DVP_ERR_ARG;
	UploadSrc u;
	const size_t N = (size_t)ctx->N;
	for (int i = 0; i <= ctx->S; ++i) {
		u.images[i] = din->images + (size_t)i * N;
		if (din->depths) u.depths[i] = din->depths + (size_t)i * N;
	}
	u.have_depths = din->depths != nullptr;
	u.cameras = din->cameras; u.cameras_on_device = from_device;
	u.planes = din->planes; u.selected_views = din->selected_views; u.weak_info = din->weak_info;
	u.edge = din->edge; u.label = din->label; u.radius = din->radius; u.seed = din->seed;
	return upload_parts(ctx, &u, params, from_device, overlap);
}

int upload_parts(dvp_ctx* ctx, const UploadSrc* in, const dvp_params* params, bool from_device, bool overlap) {
	if (!ctx || !in || !in->images[0] || !in->cameras || !in->planes) return DVP_ERR_ARG;
	CK(cudaSetDevice(ctx->device));
	{   // validate first: a rejected upload must leave the context as it was
		const dvp_params& q = params ? *params : ctx->prm;
		if (q.num_images != ctx->S + 1) return DVP_ERR_ARG;
		if (q.geom_consistency && !in->have_depths) return DVP_ERR_ARG;
		if (q.max_iterations < 0 || q.max_iterations > 64) return DVP_ERR_ARG;
		if (!q.use_edge) return DVP_ERR_UNSUPPORTED;  // the ACMH-style branch (APD.cu:2142-2460) is never enabled by main.cpp
		// GenNeighbours gives every one of its 8 origin directions 4 slots ("max is 4 from [1, 2, 4]", APD.cu:3375):
		// rotate_time is 1, 2 or 4 in every schedule (main.cpp:467, 500); more than 4 would overlap the slots of the next direction
		if (q.use_APD && (q.rotate_time < 1 || q.rotate_time > 4)) return DVP_ERR_ARG;
	}
	if (params) ctx->prm = *params;
	const cudaMemcpyKind kind = from_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
	const size_t N = (size_t)ctx->N;
	cudaStream_t st = ctx->stream;
	ctx->seed = in->seed;
	for (int i = 0; i <= ctx->S; ++i)
		if (!in->images[i] || (in->have_depths && !in->depths[i])) return DVP_ERR_ARG;
	{   // a previous overlapped upload that was never run: its copies must not overtake this one
		int r = join_copies(ctx);
		if (r) return r;
	}
	for (int i = 0; i <= ctx->S; ++i) {
		int r = ensure_texture(ctx, &ctx->img_arr[i], &ctx->img_tex[i]);
		if (!r && in->have_depths) r = ensure_texture(ctx, &ctx->dep_arr[i], &ctx->dep_tex[i]);
		if (r) return r;
	}
	CK(cudaMemcpyAsync(ctx->d_img_tex, ctx->img_tex, sizeof(ctx->img_tex), cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(ctx->d_dep_tex, ctx->dep_tex, sizeof(ctx->dep_tex), cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(ctx->ref_img, in->images[0], N * 4, kind, st));
	CK(cudaMemcpyAsync(ctx->cams, in->cameras, sizeof(dvp_camera) * (ctx->S + 1), in->cameras_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
	if (in->cameras_on_device) CK(cudaMemcpyAsync(&ctx->ref_cam, in->cameras, sizeof(dvp_camera), cudaMemcpyDeviceToHost, st));
	else ctx->ref_cam = in->cameras[0];
	CK(launch_setup_views(ctx->cams, ctx->views, ctx->S, st));
	CK(cudaMemsetAsync(ctx->fit_planes, 0, N * 16, st));
	if (in->selected_views) CK(cudaMemcpyAsync(ctx->selected, in->selected_views, N * 4, kind, st));
	else CK(cudaMemsetAsync(ctx->selected, 0, N * 4, st));
	if (in->edge) CK(cudaMemcpyAsync(ctx->edge, in->edge, N, kind, st)); else CK(cudaMemsetAsync(ctx->edge, 0, N, st));
	CK(launch_edge_distance(ctx->edge, ctx->W, ctx->H, ctx->edge_sat, ctx->edge_dist, st));
	if (in->label) CK(cudaMemcpyAsync(ctx->label, in->label, N * 4, kind, st)); else CK(cudaMemsetAsync(ctx->label, 0, N * 4, st));

	// pixel states, neighbours map, WEAK list, radius (APD.cpp:1169-1204, 1647-1667) — all on the device: the
	// reference's host loop over the state map becomes a prefix sum; only the WEAK count (4 bytes) comes back,
	// because three buffers are sized by it.
	int weak_count = 0;
	const bool have_weak = ctx->prm.use_APD && in->weak_info;
	if (have_weak) {
		CK(cudaMemcpyAsync(ctx->weak, in->weak_info, N, kind, st));
		const int yy_limit = (((ctx->H / 2) + 15) / 16) * 16;
		int counts[3] = {0, 0, 0};
		CK(launch_weak_count(ctx->weak, ctx->N, ctx->W, -1, yy_limit, ctx->scan_blocks, ctx->scan_total, st));
		for (int k = 0; k < 2; ++k) CK(launch_weak_count(ctx->weak, ctx->N, ctx->W, k, yy_limit, ctx->scan_blocks_c[k], ctx->scan_total + 1 + k, st));
		CK(cudaMemcpyAsync(counts, ctx->scan_total, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		weak_count = counts[0]; ctx->colour_count[0] = counts[1]; ctx->colour_count[1] = counts[2];
	} else {
		CK(cudaMemsetAsync(ctx->weak, DVP_STRONG, N, st));
		CK(cudaMemsetAsync(ctx->neighbours_map, 0, N * 4, st));
		ctx->colour_count[0] = ctx->colour_count[1] = 0;
	}
	ctx->weak_count = weak_count;
	if (weak_count > ctx->weak_capacity) {
		cudaFree(ctx->neighbours); cudaFree(ctx->label_boundary); cudaFree(ctx->complex_); cudaFree(ctx->weak_list);
		cudaFree(ctx->colour_list[0]); cudaFree(ctx->colour_list[1]);
		ctx->neighbours = nullptr; ctx->label_boundary = nullptr; ctx->complex_ = nullptr; ctx->weak_list = nullptr;
		ctx->colour_list[0] = ctx->colour_list[1] = nullptr;
		// grow by at least half: a multi-pass caller's WEAK counts creep up pass after pass, and every cudaFree of
		// maps this size synchronises the device
		const size_t cap = std::min((size_t)ctx->N, std::max((size_t)weak_count, (size_t)ctx->weak_capacity + (size_t)ctx->weak_capacity / 2));
		CK(zalloc(&ctx->weak_list, cap));
		CK(zalloc(&ctx->colour_list[0], cap));
		CK(zalloc(&ctx->colour_list[1], cap));
		CK(zalloc(&ctx->neighbours, cap * DVP_NEIGHBOUR_NUM));
		CK(zalloc(&ctx->label_boundary, cap * DVP_LAB_BOUNDARY_NUM));
		CK(zalloc(&ctx->complex_, cap));
		cudaFree(ctx->sort_keys[0]); cudaFree(ctx->sort_keys[1]); cudaFree(ctx->sort_vals); cudaFree(ctx->sort_temp);
		ctx->sort_keys[0] = ctx->sort_keys[1] = ctx->sort_vals = nullptr; ctx->sort_temp = nullptr;
		CK(zalloc(&ctx->sort_keys[0], cap)); CK(zalloc(&ctx->sort_keys[1], cap)); CK(zalloc(&ctx->sort_vals, cap));
		ctx->sort_temp_bytes = tile_order_temp_bytes((int)cap);
		CK(cudaMalloc(&ctx->sort_temp, ctx->sort_temp_bytes ? ctx->sort_temp_bytes : 1));
		ctx->weak_capacity = (int)cap;
	}
	if (have_weak) {
		const int yy_limit = (((ctx->H / 2) + 15) / 16) * 16;
		CK(launch_weak_index(ctx->weak, ctx->N, ctx->W, -1, yy_limit, ctx->scan_blocks, ctx->neighbours_map, weak_count > 0 ? ctx->weak_list : nullptr, st));
		if (weak_count > 0)
			for (int k = 0; k < 2; ++k)
				CK(launch_weak_index(ctx->weak, ctx->N, ctx->W, k, yy_limit, ctx->scan_blocks_c[k], nullptr, ctx->colour_list[k], st));
#ifndef DVP_NO_TILE_ORDER
		if (weak_count > 0) {
			// one block of each kernel = one compact tile: (threads/8) x 8 pixels (K4), 16 x (threads/8) pixels of one colour (K10/K11)
			CK(launch_tile_order(ctx->weak_list, weak_count, ctx->W, ctx->H, kK4Threads / 8, 8, ctx->sort_keys[0], ctx->sort_keys[1], ctx->sort_vals, ctx->sort_temp, ctx->sort_temp_bytes, st));
			for (int k = 0; k < 2; ++k)
				CK(launch_tile_order(ctx->colour_list[k], ctx->colour_count[k], ctx->W, ctx->H, 16, kWeakThreads / 8, ctx->sort_keys[0], ctx->sort_keys[1], ctx->sort_vals, ctx->sort_temp, ctx->sort_temp_bytes, st));
		}
#endif
		if (weak_count > 0)
			CK(cudaMemsetAsync(ctx->neighbours, 0xFF, (size_t)weak_count * DVP_NEIGHBOUR_NUM * sizeof(short2), st));  // (-1,-1): no anchor yet
	}
	// radius map; UNKNOWN pixels are reset to strong_radius (APD.cpp:1663-1666)
	if (in->radius) CK(cudaMemcpyAsync(ctx->radius, in->radius, N * 4, kind, st));
	else CK(launch_fill_i32(ctx->radius, ctx->prm.strong_radius, ctx->N, st));
	if (have_weak && in->radius) CK(launch_reset_unknown_radius(ctx->weak, ctx->radius, ctx->prm.strong_radius, ctx->N, st));
	// The large maps — plane hypotheses (first read by K4), the 1+S images (K6) and depth maps (K6/K7) — go to the
	// copy stream when the upload is overlapped; everything K1..K3 and K5 need stays on the context stream.  They are
	// enqueued LAST: the host-to-device DMA engine serves copies in issue order, whatever the stream.
	cudaStream_t cs = overlap ? ctx->copy_stream : st;
	if (overlap) {
		CK(cudaEventRecord(ctx->ev_main_idle, st));
		CK(cudaStreamWaitEvent(cs, ctx->ev_main_idle, 0));   // the buffers may still be in use by work queued on the context stream
	}
	CK(cudaMemcpyAsync(ctx->planes, in->planes, N * 16, kind, cs));
	if (overlap) CK(cudaEventRecord(ctx->ev_planes, cs));
	CK(launch_fill_sd_table(st));
	CK(configure_strong_kernels(ctx->S));
	CK(configure_weak_kernels(ctx->S));
	if (overlap) CK(cudaStreamSynchronize(st));   // the small maps are in place; what follows is not waited for
	for (int i = 0; i <= ctx->S; ++i) {
		int r = fill_texture(ctx, ctx->img_arr[i], in->images[i], kind, cs);
		if (!r && in->have_depths) r = fill_texture(ctx, ctx->dep_arr[i], in->depths[i], kind, cs);
		if (r) return r;
	}
	if (overlap) { CK(cudaEventRecord(ctx->ev_maps, cs)); ctx->copies_pending = true; }
	else CK(cudaStreamSynchronize(st));
	ctx->uploaded = true;
	ctx->timed_valid = false;
	return DVP_OK;
}

}  // namespace

extern "C" {

const char* dvp_version(void) { return "dvp_mvs_b200 0.1 sm_100a"; }

void dvp_default_params(dvp_params* p) {
	if (!p) return;
	// reference defaults, main.h:86-112
	p->max_iterations = 3; p->num_images = 5; p->sigma_spatial = 5.0f; p->sigma_color = 3.0f; p->top_k = 4;
	p->depth_min = 0.0f; p->depth_max = 1.0f; p->geom_consistency = 0; p->strong_radius = 5; p->strong_increment = 2;
	p->weak_radius = 5; p->weak_increment = 5; p->use_APD = 1; p->use_edge = 1; p->use_limit = 1; p->use_label = 1;
	p->use_detail = 0; p->use_radius = 1; p->weak_peak_radius = 2; p->rotate_time = 4; p->ransac_threshold = 0.005f;
	p->geom_factor = 0.2f; p->state = DVP_FIRST_INIT;
}

dvp_ctx* dvp_create(int device, int width, int height, int num_src, const dvp_params* params) {
	if (!params || width <= 0 || height <= 0 || width > 32767 || height > 32767 || num_src < 1 || num_src + 1 > DVP_MAX_IMAGES) return nullptr;
	if (cudaSetDevice(device) != cudaSuccess) return nullptr;
	dvp_ctx* c = new (std::nothrow) dvp_ctx();
	if (!c) return nullptr;
	c->device = device; c->W = width; c->H = height; c->S = num_src; c->N = width * height; c->prm = *params;
	const size_t N = (size_t)c->N;
	bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
	ok = ok && cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&c->ev_main_idle, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&c->ev_planes, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&c->ev_maps, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && zalloc(&c->d_img_tex, DVP_MAX_IMAGES) == cudaSuccess;
	ok = ok && zalloc(&c->d_dep_tex, DVP_MAX_IMAGES) == cudaSuccess;
	ok = ok && zalloc(&c->ref_img, N) == cudaSuccess;
	ok = ok && zalloc(&c->cams, (size_t)num_src + 1) == cudaSuccess;
	ok = ok && zalloc(&c->views, (size_t)num_src) == cudaSuccess;
	ok = ok && zalloc(&c->planes, N) == cudaSuccess;
	ok = ok && zalloc(&c->fit_planes, N) == cudaSuccess;
	ok = ok && zalloc(&c->costs, N) == cudaSuccess;
	ok = ok && zalloc(&c->selected_alloc, N + 2 * (size_t)width + 2) == cudaSuccess;
	ok = ok && zalloc(&c->weak, N) == cudaSuccess;
	ok = ok && zalloc(&c->radius, N) == cudaSuccess;
	ok = ok && zalloc(&c->view_weight, N * DVP_MAX_IMAGES) == cudaSuccess;
	ok = ok && zalloc(&c->rng, N * 6) == cudaSuccess;
	ok = ok && zalloc(&c->edge, N) == cudaSuccess;
	ok = ok && zalloc(&c->edge_sat, (size_t)(width + 1) * (height + 1)) == cudaSuccess;
	ok = ok && zalloc(&c->edge_dist, N) == cudaSuccess;
	ok = ok && zalloc(&c->edge_neigh, N * DVP_EDGE_NEIGH_NUM) == cudaSuccess;
	ok = ok && zalloc(&c->label, N) == cudaSuccess;
	const int cand_views = num_src > DVP_NUM_IMAGES ? num_src : DVP_NUM_IMAGES;
	ok = ok && zalloc(&c->candidate, (N + 1) * DVP_LAB_BOUNDARY_NUM * cand_views) == cudaSuccess;
	ok = ok && zalloc(&c->nearest_strong, N) == cudaSuccess;
	ok = ok && zalloc(&c->weak_reliable, N) == cudaSuccess;
	ok = ok && zalloc(&c->anchor_flag, N) == cudaSuccess;
	ok = ok && zalloc(&c->neighbours_map, N) == cudaSuccess;
	ok = ok && zalloc(&c->neighbours, 1) == cudaSuccess;
	ok = ok && zalloc(&c->label_boundary, 1) == cudaSuccess;
	ok = ok && zalloc(&c->complex_, 1) == cudaSuccess;
	ok = ok && zalloc(&c->scan_blocks, (N + kWeakScanBlock - 1) / kWeakScanBlock + 1) == cudaSuccess;
	ok = ok && zalloc(&c->scan_total, 4) == cudaSuccess;
	ok = ok && zalloc(&c->next_right, N) == cudaSuccess;
	ok = ok && zalloc(&c->next_down, N) == cudaSuccess;
	for (int k = 0; k < 2; ++k) ok = ok && zalloc(&c->scan_blocks_c[k], (N + kWeakScanBlock - 1) / kWeakScanBlock + 1) == cudaSuccess;
	for (size_t i = 0; ok && i < sizeof(c->ev) / sizeof(c->ev[0]); ++i) ok = cudaEventCreate(&c->ev[i]) == cudaSuccess;
	if (!ok) {
		fprintf(stderr, "[dvp] dvp_create: allocation failed: %s\n", cudaGetErrorString(cudaGetLastError()));
		dvp_destroy(c);
		return nullptr;
	}
	c->selected = c->selected_alloc + width + 1;
	return c;
}

void dvp_destroy(dvp_ctx* c) {
	if (!c) return;
	cudaSetDevice(c->device);
	if (c->stream) cudaStreamSynchronize(c->stream);
	if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
	if (c->ev_main_idle) cudaEventDestroy(c->ev_main_idle);
	if (c->ev_planes) cudaEventDestroy(c->ev_planes);
	if (c->ev_maps) cudaEventDestroy(c->ev_maps);
	for (int i = 0; i < DVP_MAX_IMAGES; ++i) {
		if (c->img_tex[i]) cudaDestroyTextureObject(c->img_tex[i]);
		if (c->dep_tex[i]) cudaDestroyTextureObject(c->dep_tex[i]);
		if (c->img_arr[i]) cudaFreeArray(c->img_arr[i]);
		if (c->dep_arr[i]) cudaFreeArray(c->dep_arr[i]);
	}
	cudaFree(c->d_img_tex); cudaFree(c->d_dep_tex); cudaFree(c->ref_img); cudaFree(c->cams); cudaFree(c->views);
	cudaFree(c->planes); cudaFree(c->fit_planes); cudaFree(c->costs); cudaFree(c->selected_alloc); cudaFree(c->weak);
	cudaFree(c->radius); cudaFree(c->view_weight); cudaFree(c->rng); cudaFree(c->edge); cudaFree(c->edge_sat); cudaFree(c->edge_dist); cudaFree(c->edge_neigh);
	cudaFree(c->label); cudaFree(c->candidate); cudaFree(c->nearest_strong); cudaFree(c->weak_reliable); cudaFree(c->anchor_flag);
	cudaFree(c->neighbours_map); cudaFree(c->neighbours); cudaFree(c->label_boundary); cudaFree(c->complex_); cudaFree(c->weak_list); cudaFree(c->scan_blocks); cudaFree(c->scan_total); cudaFree(c->next_right); cudaFree(c->next_down); for (int k = 0; k < 2; ++k) { cudaFree(c->scan_blocks_c[k]); cudaFree(c->colour_list[k]); }
	cudaFree(c->vis_parent); cudaFree(c->vis_count); cudaFree(c->fetch_counter); cudaFree(c->sweep_scratch);
	cudaFree(c->sort_keys[0]); cudaFree(c->sort_keys[1]); cudaFree(c->sort_vals); cudaFree(c->sort_temp);
	for (size_t i = 0; i < sizeof(c->ev) / sizeof(c->ev[0]); ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
	if (c->stream) cudaStreamDestroy(c->stream);
	delete c;
}

int dvp_upload(dvp_ctx* ctx, const dvp_inputs* in, const dvp_params* params) { return upload_common(ctx, in, params, false); }
int dvp_upload_device(dvp_ctx* ctx, const dvp_inputs* in, const dvp_params* params) { return upload_common(ctx, in, params, true); }
int dvp_upload_overlapped(dvp_ctx* ctx, const dvp_inputs* in, const dvp_params* params) { return upload_common(ctx, in, params, false, true); }

int dvp_run_stage(dvp_ctx* ctx, int stage, int iter) {
	if (!ctx) return DVP_ERR_ARG;
	if (!ctx->uploaded) return DVP_ERR_STATE;
	if (stage < 0 || stage > DVP_K15_K16_FUSED) return DVP_ERR_ARG;
	CK(cudaSetDevice(ctx->device));
	{ int r = join_copies(ctx); if (r) return r; }
	const KArgs a = make_args(ctx);
	cudaError_t e = launch_stage(ctx, a, stage, iter);
	if (e == cudaErrorNotSupported) return DVP_ERR_UNSUPPORTED;
	CK(e);
	CK(cudaStreamSynchronize(ctx->stream));
	return DVP_OK;
}

int dvp_run(dvp_ctx* ctx, int sync) {
	if (!ctx) return DVP_ERR_ARG;
	if (!ctx->uploaded) return DVP_ERR_STATE;
	if (ctx->prm.max_iterations > 64) return DVP_ERR_ARG;
	CK(cudaSetDevice(ctx->device));
	const KArgs a = make_args(ctx);
	cudaStream_t st = ctx->stream;
	int n = 0;
	ctx->timed_valid = false;
	auto go = [&](int stage, int iter) -> int {
		if (ctx->copies_pending) {   // overlapped upload: K1..K3 and K5 need none of the large maps
			if (stage == DVP_K4_GEN_NEIGHBOURS) CK(cudaStreamWaitEvent(st, ctx->ev_planes, 0));
			else if (stage >= DVP_K6_RANDOM_INITIALIZATION) { int jr = join_copies(ctx); if (jr) return jr; }
		}
		CK(cudaEventRecord(ctx->ev[2 * n], st));
		cudaError_t e = launch_stage(ctx, a, stage, iter);
		if (e == cudaErrorNotSupported) return DVP_ERR_UNSUPPORTED;
		CK(e);
		CK(cudaEventRecord(ctx->ev[2 * n + 1], st));
		ctx->ev_stage[n] = stage;
		++n;
		return DVP_OK;
	};
	int r;
#ifndef DVP_EAGER_CANDIDATES
	ctx->lazy_candidates = true;   // K2's candidate records after K4, for the anchors only; stage stepping (dvp_run_stage) keeps K2 whole
#endif
	ctx->last_run_lazy = ctx->lazy_candidates;
	struct Unset { dvp_ctx* c; ~Unset() { c->lazy_candidates = false; } } unset{ctx};
	for (int s = DVP_K1_INIT_RANDOM_STATES; s <= DVP_K6_RANDOM_INITIALIZATION; ++s) if ((r = go(s, 0))) return r;
	for (int it = 0; it < ctx->prm.max_iterations; ++it)
		for (int s = DVP_K7_BLACK_STRONG; s <= DVP_K11_RED_WEAK; ++s) if ((r = go(s, it))) return r;
	for (int s = DVP_K12_DEPTH_NORMAL; s <= DVP_K14_RED_FILTER; ++s) if ((r = go(s, 0))) return r;
	if ((r = go(DVP_K15_K16_FUSED, 0))) return r;   // K15 + K16 in one launch (timed under K15)
	ctx->n_timed = n;
	ctx->timed_valid = true;
	if (sync) CK(cudaStreamSynchronize(st));
	return DVP_OK;
}

int dvp_last_run_times(dvp_ctx* ctx, float* total_ms, float* per_stage_ms, int* launches) {
	if (!ctx) return DVP_ERR_ARG;
	if (!ctx->timed_valid) return DVP_ERR_STATE;
	CK(cudaSetDevice(ctx->device));
	CK(cudaStreamSynchronize(ctx->stream));
	if (per_stage_ms) for (int i = 0; i < DVP_STAGE_COUNT; ++i) per_stage_ms[i] = 0.f;
	float sum = 0.f;
	for (int i = 0; i < ctx->n_timed; ++i) {
		float ms = 0.f;
		CK(cudaEventElapsedTime(&ms, ctx->ev[2 * i], ctx->ev[2 * i + 1]));
		if (per_stage_ms) per_stage_ms[ctx->ev_stage[i] == DVP_K15_K16_FUSED ? DVP_K15_DEPTH_TO_WEAK : ctx->ev_stage[i]] += ms;
		sum += ms;
	}
	if (total_ms) {
		// wall time on the stream from the first launch to the end of the last one
		float ms = 0.f;
		CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[2 * (ctx->n_timed - 1) + 1]));
		*total_ms = ms;
	}
	(void)sum;
	if (launches) {
		int n = 0;
		const bool was = ctx->lazy_candidates;
		ctx->lazy_candidates = ctx->last_run_lazy;
		for (int i = 0; i < ctx->n_timed; ++i) n += stage_kernel_count(ctx, ctx->ev_stage[i]);
		ctx->lazy_candidates = was;
		*launches = n;
	}
	return DVP_OK;
}

int dvp_download(dvp_ctx* ctx, float* planes, uint8_t* weak_info, uint32_t* selected_views, int32_t* radius) {
	if (!ctx) return DVP_ERR_ARG;
	CK(cudaSetDevice(ctx->device));
	{ int r = join_copies(ctx); if (r) return r; }
	const size_t N = (size_t)ctx->N;
	cudaStream_t st = ctx->stream;
	// cudaMemcpyDefault: destinations may be host (pageable / pinned) or device memory (in-memory pass chaining, row N2)
	if (planes) CK(cudaMemcpyAsync(planes, ctx->planes, N * 16, cudaMemcpyDefault, st));
	if (weak_info) CK(cudaMemcpyAsync(weak_info, ctx->weak, N, cudaMemcpyDefault, st));
	if (selected_views) CK(cudaMemcpyAsync(selected_views, ctx->selected, N * 4, cudaMemcpyDefault, st));
	if (radius) CK(cudaMemcpyAsync(radius, ctx->radius, N * 4, cudaMemcpyDefault, st));
	CK(cudaStreamSynchronize(st));
	return DVP_OK;
}

size_t dvp_buffer_bytes(dvp_ctx* ctx, int buffer) { return ctx ? buf_desc(ctx, buffer).bytes : 0; }

int dvp_get_buffer(dvp_ctx* ctx, int buffer, void* dst, size_t bytes) {
	if (!ctx || !dst) return DVP_ERR_ARG;
	CK(cudaSetDevice(ctx->device));
	{ int r = join_copies(ctx); if (r) return r; }
	BufDesc b = buf_desc(ctx, buffer);
	if (!b.ptr || bytes != b.bytes) return DVP_ERR_ARG;
	if (bytes == 0) return DVP_OK;
	if (buffer == DVP_BUF_RAND) {
		uint32_t* tmp = nullptr;
		CK(cudaMalloc((void**)&tmp, bytes));
		const KArgs a = make_args(ctx);
		CK(launch_rng_export(a, tmp, ctx->stream));
		CK(cudaMemcpyAsync(dst, tmp, bytes, cudaMemcpyDeviceToHost, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
		cudaFree(tmp);
		return DVP_OK;
	}
	CK(cudaMemcpyAsync(dst, b.ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return DVP_OK;
}

int dvp_set_buffer(dvp_ctx* ctx, int buffer, const void* src, size_t bytes) {
	if (!ctx || !src) return DVP_ERR_ARG;
	CK(cudaSetDevice(ctx->device));
	{ int r = join_copies(ctx); if (r) return r; }
	BufDesc b = buf_desc(ctx, buffer);
	if (!b.ptr || bytes != b.bytes) return DVP_ERR_ARG;
	if (bytes == 0) return DVP_OK;
	if (buffer == DVP_BUF_RAND) {
		uint32_t* tmp = nullptr;
		CK(cudaMalloc((void**)&tmp, bytes));
		CK(cudaMemcpyAsync(tmp, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
		const KArgs a = make_args(ctx);
		CK(launch_rng_import(a, tmp, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
		cudaFree(tmp);
		return DVP_OK;
	}
	CK(cudaMemcpyAsync(b.ptr, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return DVP_OK;
}

int dvp_restore_visibility(dvp_ctx* ctx, int scale_size, float* device_ms) {
	if (!ctx || scale_size <= 0) return DVP_ERR_ARG;
	if (!ctx->uploaded) return DVP_ERR_STATE;
	CK(cudaSetDevice(ctx->device));
	{ int r = join_copies(ctx); if (r) return r; }
	if (!ctx->vis_parent) CK(zalloc(&ctx->vis_parent, (size_t)ctx->N));
	if (!ctx->vis_count) CK(zalloc(&ctx->vis_count, (size_t)ctx->N));
	const KArgs a = make_args(ctx);
	cudaStream_t st = ctx->stream;
	const int e0 = 2 * dvp_ctx::kMaxLaunch, e1 = e0 + 1;   // the two spare events after the per-launch pairs
	CK(cudaEventRecord(ctx->ev[e0], st));
	CK(launch_invalidate_depth(a, st));
	CK(launch_restore_visibility(a, scale_size, ctx->vis_parent, ctx->vis_count, st));
	CK(cudaEventRecord(ctx->ev[e1], st));
	CK(cudaStreamSynchronize(st));
	if (device_ms) CK(cudaEventElapsedTime(device_ms, ctx->ev[e0], ctx->ev[e1]));
	return DVP_OK;
}

long long dvp_debug_fetch_count(dvp_ctx* ctx, int reset) {
#ifdef DVP_COUNT_FETCHES
	if (!ctx) return DVP_ERR_ARG;
	if (cudaSetDevice(ctx->device) != cudaSuccess) return DVP_ERR_CUDA;
	if (!ctx->fetch_counter) {
		if (zalloc(&ctx->fetch_counter, (size_t)kFetchSlots) != cudaSuccess) return DVP_ERR_CUDA;
		return 0;
	}
	cudaStreamSynchronize(ctx->stream);
	std::vector<unsigned long long> h(kFetchSlots);
	if (cudaMemcpy(h.data(), ctx->fetch_counter, kFetchSlots * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return DVP_ERR_CUDA;
	unsigned long long sum = 0;
	for (unsigned long long v : h) sum += v;
	if (reset) cudaMemset(ctx->fetch_counter, 0, kFetchSlots * sizeof(unsigned long long));
	return (long long)sum;
#else
	(void)ctx; (void)reset;
	return DVP_ERR_UNSUPPORTED;   // only libdvp_mvs_count.so counts
#endif
}

int dvp_debug_race_explain(dvp_ctx* ctx, int iter, int red, const int32_t* offsets, int num_offsets, const float* planes_before, const float* planes_after,
                           const float* exp_planes, const float* exp_costs, const uint32_t* exp_selected, const uint8_t* exp_view_weight, const uint32_t* exp_rand,
                           int tear, uint8_t* explained, long long* stats) {
	if (!ctx || !offsets || num_offsets <= 0 || !planes_before || !planes_after || !exp_planes || !exp_costs || !exp_selected || !exp_view_weight || !exp_rand || !explained)
		return DVP_ERR_ARG;
	if (!ctx->uploaded) return DVP_ERR_STATE;
	CK(cudaSetDevice(ctx->device));
	{ int r = join_copies(ctx); if (r) return r; }
	const size_t N = (size_t)ctx->N;
	cudaStream_t st = ctx->stream;
	const KArgs a = make_args(ctx);
	// device copies of the snapshots and of the observed result, shadow outputs, bookkeeping
	struct Bufs {
		float4 *before = nullptr, *after = nullptr, *e_planes = nullptr, *o_planes = nullptr;
		float *e_costs = nullptr, *o_costs = nullptr;
		uint32_t *e_sel = nullptr, *o_sel = nullptr, *e_rand = nullptr, *o_rng = nullptr;
		uint8_t *e_vw = nullptr, *o_vw = nullptr, *expl = nullptr;
		int *list = nullptr, *count = nullptr;
		~Bufs() { cudaFree(before); cudaFree(after); cudaFree(e_planes); cudaFree(o_planes); cudaFree(e_costs); cudaFree(o_costs); cudaFree(e_sel); cudaFree(o_sel);
		          cudaFree(e_rand); cudaFree(o_rng); cudaFree(e_vw); cudaFree(o_vw); cudaFree(expl); cudaFree(list); cudaFree(count); }
	} b;
	const int list_cap = 1 << 16;
	CK(cudaMalloc((void**)&b.before, N * 16)); CK(cudaMalloc((void**)&b.after, N * 16)); CK(cudaMalloc((void**)&b.e_planes, N * 16)); CK(cudaMalloc((void**)&b.o_planes, N * 16));
	CK(cudaMalloc((void**)&b.e_costs, N * 4)); CK(cudaMalloc((void**)&b.o_costs, N * 4)); CK(cudaMalloc((void**)&b.e_sel, N * 4)); CK(cudaMalloc((void**)&b.o_sel, N * 4));
	CK(cudaMalloc((void**)&b.e_rand, N * 24)); CK(cudaMalloc((void**)&b.o_rng, N * 24)); CK(cudaMalloc((void**)&b.e_vw, N * DVP_MAX_IMAGES)); CK(cudaMalloc((void**)&b.o_vw, N * DVP_MAX_IMAGES));
	CK(cudaMalloc((void**)&b.expl, N)); CK(cudaMalloc((void**)&b.list, (size_t)list_cap * 4)); CK(cudaMalloc((void**)&b.count, 4));
	CK(cudaMemcpyAsync(b.before, planes_before, N * 16, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(b.after, planes_after, N * 16, cudaMemcpyDefault, st));
	CK(cudaMemcpyAsync(b.e_planes, exp_planes, N * 16, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(b.e_costs, exp_costs, N * 4, cudaMemcpyDefault, st));
	CK(cudaMemcpyAsync(b.e_sel, exp_selected, N * 4, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(b.e_vw, exp_view_weight, N * DVP_MAX_IMAGES, cudaMemcpyDefault, st));
	CK(cudaMemcpyAsync(b.e_rand, exp_rand, N * 24, cudaMemcpyDefault, st));
	const RaceExpected e{b.e_planes, b.e_costs, b.e_sel, b.e_vw, b.e_rand};
	CK(launch_explain_init(a, red ? 1 : 0, e, b.expl, st));
	D4Force f;
	f.before = b.before; f.after = b.after;
	f.out_planes = b.o_planes; f.out_costs = b.o_costs; f.out_selected = b.o_sel; f.out_view_weight = b.o_vw; f.out_rng = b.o_rng;
	long long launches = 0;
	const int S = ctx->S;
	auto all_views = [&](unsigned m4) { unsigned long long w = 0; for (int v = 0; v < S && v < 16; ++v) w |= (unsigned long long)(m4 & 15u) << (4 * v); return w; };
	auto try_choice = [&]() -> int {
		CK(launch_strong_sweep_forced(a, iter, red ? 1 : 0, f, st));
		CK(launch_explain_compare(a, red ? 1 : 0, f, e, b.expl, st));
		++launches;
		return DVP_OK;
	};
	// phase 1, whole colour: every ladder offset, each of the three reads entirely before or entirely after the update
	for (int k = 0; k < num_offsets; ++k) {
		if (offsets[k] < 0 || offsets[k] > 65535) return DVP_ERR_ARG;
		f.m = offsets[k];
		for (unsigned v = 0; v < 8; ++v) {
			f.ncc_masks = all_views((v & 1) ? 15u : 0u); f.dep_mask = (v & 2) ? 15u : 0u; f.acc_mask = (v & 4) ? 15u : 0u;
			const int r = try_choice(); if (r) return r;
		}
	}
	int left1 = 0, left2 = 0;
	CK(cudaMemsetAsync(b.count, 0, 4, st));
	CK(launch_explain_collect(a, red ? 1 : 0, b.expl, b.list, b.count, list_cap, st));
	CK(cudaMemcpyAsync(&left1, b.count, 4, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	left2 = left1;
	// phase 2, only the pixels still unexplained: reads torn between components.  The reference build loads a plane with
	// four 32-bit loads (SASS: 840 LD.E against 14 LD.E.128 in BlackPixelUpdateStrong) while its owner replaces it with wider
	// stores, and it loads the candidate's plane anew for every source view it scores (the NCC is an out-of-line call).
	if (tear && left1 > 0 && left1 <= list_cap) {
		f.pixel_list = b.list; f.list_count = left1;
		// (a) one mixture for all views x every mixture of the two acceptance reads
		for (int k = 0; k < num_offsets; ++k) {
			f.m = offsets[k];
			for (unsigned v = 0; v < 4096; ++v) {
				const unsigned n4 = v & 15u;
				f.ncc_masks = all_views(n4); f.dep_mask = (v >> 4) & 15u; f.acc_mask = (v >> 8) & 15u;
				const bool whole = (n4 == 0 || n4 == 15) && (f.dep_mask == 0 || f.dep_mask == 15) && (f.acc_mask == 0 || f.acc_mask == 15);
				if (whole) continue;   // done in phase 1
				const int r = try_choice(); if (r) return r;
			}
		}
		// (b) a mixture per view, the acceptance reads whole.  Up to two views: all 16 per view; more views: the mixtures four
		// loads in ascending or descending component order can see when one store lands between them
		static const unsigned ordered[8] = {0x0, 0xF, 0x8, 0xC, 0xE, 0x1, 0x3, 0x7};
		const int per_view = S <= 2 ? 16 : 8;
		if (S <= 4) {
			long long combos = 1;
			for (int v = 0; v < S; ++v) combos *= per_view;
			for (int k = 0; k < num_offsets; ++k) {
				f.m = offsets[k];
				for (long long cmb = 0; cmb < combos; ++cmb) {
					unsigned long long w = 0; long long rest = cmb; bool uniform = true; unsigned first = 0;
					for (int v = 0; v < S; ++v) {
						const unsigned m4 = S <= 2 ? (unsigned)(rest % per_view) : ordered[rest % per_view];
						rest /= per_view;
						if (v == 0) first = m4; else if (m4 != first) uniform = false;
						w |= (unsigned long long)m4 << (4 * v);
					}
					if (uniform) continue;   // the same mixture for every view: covered by (a)
					f.ncc_masks = w;
					for (unsigned v = 0; v < 4; ++v) {
						f.dep_mask = (v & 1) ? 15u : 0u; f.acc_mask = (v & 2) ? 15u : 0u;
						const int r = try_choice(); if (r) return r;
					}
				}
			}
		}
		f.pixel_list = nullptr; f.list_count = 0;
		CK(cudaMemsetAsync(b.count, 0, 4, st));
		CK(launch_explain_collect(a, red ? 1 : 0, b.expl, b.list, b.count, list_cap, st));
		CK(cudaMemcpyAsync(&left2, b.count, 4, cudaMemcpyDeviceToHost, st));
	}
	CK(cudaMemcpyAsync(explained, b.expl, N, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	if (stats) { stats[0] = left1; stats[1] = left2; stats[2] = launches; }
	return DVP_OK;
}

int dvp_weak_count(dvp_ctx* ctx) { return ctx ? ctx->weak_count : -1; }
int dvp_last_cuda_error(dvp_ctx* ctx) { return ctx ? ctx->last_err : 0; }
void* dvp_stream(dvp_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

}  // extern "C"

#include "dvp_scene.inc"
#include "dvp_farm.inc"
