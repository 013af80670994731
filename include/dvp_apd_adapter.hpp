// dvp_apd_adapter.hpp — header-only `APD` class with the reference's public surface (reference APD.h:94-115)
// on top of the flat C ABI in dvp_mvs.h, so that the reference's driver (main.cpp:267-419, ProcessProblem)
// compiles and runs unchanged against libdvp_mvs.so.
//
// How a maintainer uses it (INTEGRATION.md has the full recipe):
//   1. keep main.cpp, main.h and the host-only parts of APD.cpp (file I/O, EdgeSegment, fusion);
//   2. replace the `class APD { ... }` declaration in APD.h by `#include <dvp_apd_adapter.hpp>` and drop APD.cu
//      plus the APD:: member definitions of APD.cpp:984-1748 that are re-implemented here, EXCEPT
//      InuputInitialization / SupportInitialization (APD.cpp:1045-1495, 1615-1668), which stay as they are —
//      they only fill the host members declared below (same names, same types) from disk;
//   3. link with -ldvp_mvs.
//
// Requires the reference's own main.h (Camera, Problem, PatchMatchParams, PixelState, MAX_IMAGES ...) and OpenCV
// to be included first, exactly as the reference's APD.h does.  Error policy is the reference's: a failing
// call prints to stderr and exit(EXIT_FAILURE)s (APD.cpp:943-951); the library underneath never exits.
#ifndef DVP_APD_ADAPTER_HPP
#define DVP_APD_ADAPTER_HPP

#include "dvp_mvs.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifndef DVP_ADAPTER_SEED
#define DVP_ADAPTER_SEED() ((unsigned long long)std::chrono::steady_clock::now().time_since_epoch().count())  // the reference seeds with clock64()
#endif

class APD {
public:
	APD(const Problem& problem) : plane_hypotheses_host(nullptr), ctx_(nullptr), device_(0) {
		params_host = problem.params;   // APD.cpp:984-987
		this->problem = problem;
		cudaGetDevice(&device_);        // main() selected it with cudaSetDevice (main.cpp:430-434)
	}
	~APD() {
		delete[] plane_hypotheses_host;  // APD.cpp:990
		if (ctx_) dvp_destroy(ctx_);
	}

	// Host-side loaders: UNCHANGED reference code (APD.cpp:1045-1495 and 1615-1668) — they only touch the host
	// members below.  Declared here, defined by the maintainer's copy of APD.cpp.
	void InuputInitialization();
	void SupportInitialization();

	// Replaces APD.cpp:1497-1613: every cudaMalloc / texture / H2D copy now happens inside dvp_create + dvp_upload.
	void CudaSpaceInitialization() {
		dvp_params p = to_dvp(params_host);
		if (!ctx_) ctx_ = dvp_create(device_, width, height, num_images - 1, &p);
		if (!ctx_) die("dvp_create");
		const size_t N = (size_t)width * height;
		staged_images_.resize((size_t)num_images * N);
		for (int i = 0; i < num_images; ++i) copy_rows<float>(images[i], &staged_images_[(size_t)i * N]);
		const float* depth_ptr = nullptr;
		if (params_host.geom_consistency) {
			staged_depths_.resize((size_t)num_images * N);
			for (int i = 0; i < num_images; ++i) copy_rows<float>(depths[i], &staged_depths_[(size_t)i * N]);
			depth_ptr = staged_depths_.data();
		}
		static_assert(sizeof(dvp_camera) == sizeof(Camera), "Camera layout (main.h:58-67)");
		dvp_inputs in;
		std::memset(&in, 0, sizeof(in));
		in.images = staged_images_.data();
		in.depths = depth_ptr;
		in.cameras = reinterpret_cast<const dvp_camera*>(cameras.data());
		in.planes = reinterpret_cast<const float*>(plane_hypotheses_host);
		in.selected_views = selected_views_host.ptr<unsigned int>(0);
		in.weak_info = weak_info_host.ptr<uchar>(0);
		in.edge = edge_host.empty() ? nullptr : edge_host.ptr<uchar>(0);
		in.label = label_host.empty() ? nullptr : label_host.ptr<int>(0);   // reference bug B9: label_host may be empty
		in.radius = radius_host.empty() ? nullptr : radius_host.ptr<int>(0);
		in.seed = DVP_ADAPTER_SEED();
		if (dvp_upload(ctx_, &in, &p) != DVP_OK) die("dvp_upload");
	}
	// The pointer bundle no longer exists (kernel arguments carry it): kept for call-order compatibility (main.cpp:279).
	void SetDataPassHelperInCuda() {}

	// Replaces APD.cu:4406-4532.
	void RunPatchMatch() {
		if (dvp_run(ctx_, 1) != DVP_OK) die("dvp_run");
		if (dvp_download(ctx_, reinterpret_cast<float*>(plane_hypotheses_host), weak_info_host.ptr<uchar>(0),
		                 selected_views_host.ptr<unsigned int>(0),
		                 problem.params.use_radius ? radius_host.ptr<int>(0) : nullptr) != DVP_OK) die("dvp_download");
	}

	float4 GetPlaneHypothesis(int r, int c) { return plane_hypotheses_host[c + r * width]; }           // APD.cpp:1706
	int GetPixelSelectedViews(int r, int c) { return selected_views_host.at<int>(r, c); }                // APD.cpp:1710
	void SetPixelSelectedViews(int r, int c, int v) { selected_views_host.at<int>(r, c) = v; }           // APD.cpp:1714
	cv::Mat GetEdge() { return edge_host; }
	cv::Mat GetPixelStates() { return weak_info_host; }
	cv::Mat GetSelectedViews() { return selected_views_host; }
	cv::Mat GetRadiusMap() { return radius_host; }
	int GetWidth() { return width; }
	int GetHeight() { return height; }
	float GetDepthMin() { return params_host.depth_min; }
	float GetDepthMax() { return params_host.depth_max; }

	// Not in the reference (row N1, optional): what ProcessProblem does to the maps on the host right after
	// RunPatchMatch — depth range check + per-view visibility restoration (main.cpp:297-363) — run on the device
	// on the maps that are still resident, then mirrored into the host members.  A maintainer who calls this can
	// delete main.cpp:288-363 except the two lines that copy depth and normal out.
	void RestoreVisibilityOnDevice() {
		if (dvp_restore_visibility(ctx_, problem.scale_size, nullptr) != DVP_OK) die("dvp_restore_visibility");
		if (dvp_download(ctx_, reinterpret_cast<float*>(plane_hypotheses_host), weak_info_host.ptr<uchar>(0),
		                 selected_views_host.ptr<unsigned int>(0), nullptr) != DVP_OK) die("dvp_download");
	}

	// Not in the reference: costs never leave the device there. Fills a CV_32F map.
	void GetCostMap(cv::Mat& out) {
		out.create(height, width, CV_32FC1);
		if (dvp_get_buffer(ctx_, DVP_BUF_COSTS, out.ptr<float>(0), (size_t)width * height * 4) != DVP_OK) die("dvp_get_buffer");
	}

private:
	// ---- host members with the reference's names and types (APD.h:120-198); filled by the unchanged loaders ----
	int num_images;
	int width;
	int height;
	Problem problem;
	std::vector<cv::Mat> images;
	std::vector<cv::Mat> depths;
	std::vector<Camera> cameras;
	int weak_count;
	cv::Mat weak_info_host;
	cv::Mat neighbours_map_host;   // built by InuputInitialization; the library rebuilds it from weak_info
	float4* plane_hypotheses_host;
	cv::Mat edge_host;
	cv::Mat radius_host;
	cv::Mat label_host;
	PatchMatchParams params_host;
	cv::Mat selected_views_host;
	// ---- adapter state ----
	dvp_ctx* ctx_;
	int device_;
	std::vector<float> staged_images_, staged_depths_;

	static dvp_params to_dvp(const PatchMatchParams& s) {
		dvp_params d;
		d.max_iterations = s.max_iterations; d.num_images = s.num_images; d.sigma_spatial = s.sigma_spatial;
		d.sigma_color = s.sigma_color; d.top_k = s.top_k; d.depth_min = s.depth_min; d.depth_max = s.depth_max;
		d.geom_consistency = s.geom_consistency; d.strong_radius = s.strong_radius; d.strong_increment = s.strong_increment;
		d.weak_radius = s.weak_radius; d.weak_increment = s.weak_increment; d.use_APD = s.use_APD; d.use_edge = s.use_edge;
		d.use_limit = s.use_limit; d.use_label = s.use_label; d.use_detail = s.use_detail; d.use_radius = s.use_radius;
		d.weak_peak_radius = s.weak_peak_radius; d.rotate_time = s.rotate_time; d.ransac_threshold = s.ransac_threshold;
		d.geom_factor = s.geom_factor; d.state = (int)s.state;
		return d;
	}
	template <typename T> void copy_rows(const cv::Mat& m, T* dst) const {   // cv::Mat rows may be padded (step[0])
		for (int r = 0; r < height; ++r) std::memcpy(dst + (size_t)r * width, m.ptr<T>(r), (size_t)width * sizeof(T));
	}
	void die(const char* what) const {
		std::fprintf(stderr, "[APD/dvp] %s failed (cudaError %d)\n", what, ctx_ ? dvp_last_cuda_error(ctx_) : -1);
		std::exit(EXIT_FAILURE);
	}
};

// ---- rows N3 / N4 of the scope table: the two other device-side replacements a maintainer can switch to ---------------
#include <unordered_map>

// Replaces the fusing loop of RunFusion (APD.cpp:1875-1957).  Call it where that loop stands, with the vectors the
// loading loop above it (APD.cpp:1841-1873) filled; points are appended to PointCloud in the reference's order
// (views in index order, pixels in raster order).  `blocks` is only read when use_block is set.
inline void DvpRunFusionLoop(const std::vector<Problem>& problems, std::unordered_map<int, int>& imageIdToindexMap,
		const std::vector<cv::Mat>& images, const std::vector<Camera>& cameras, const std::vector<cv::Mat>& depths,
		const std::vector<cv::Mat>& normals, const std::vector<cv::Mat>& weaks, const std::vector<cv::Mat>& blocks,
		bool use_block, std::vector<PointList>& PointCloud) {
	const int num_images = (int)problems.size();
	int device = 0;
	cudaGetDevice(&device);
	dvp_fusion* f = dvp_fusion_create(device, num_images);
	if (!f) { std::fprintf(stderr, "[APD/dvp] dvp_fusion_create failed\n"); std::exit(EXIT_FAILURE); }
	auto pack = [](const cv::Mat& m, size_t elem, std::vector<unsigned char>& out) {   // cv::Mat rows may be padded
		out.resize((size_t)m.rows * m.cols * elem);
		for (int r = 0; r < m.rows; ++r) std::memcpy(&out[(size_t)r * m.cols * elem], m.ptr<unsigned char>(r), (size_t)m.cols * elem);
	};
	std::vector<unsigned char> depth, normal, image, weak, block;
	for (int i = 0; i < num_images; ++i) {
		const int ref_index = imageIdToindexMap[problems[i].ref_image_id];
		std::vector<int32_t> src;
		for (size_t j = 0; j < problems[i].src_image_ids.size(); ++j) src.push_back(imageIdToindexMap[problems[i].src_image_ids[j]]);
		pack(depths[ref_index], 4, depth); pack(normals[ref_index], 12, normal); pack(images[ref_index], 3, image); pack(weaks[ref_index], 1, weak);
		if (use_block) pack(blocks[ref_index], 1, block);
		dvp_fusion_view v;
		std::memset(&v, 0, sizeof(v));
		static_assert(sizeof(dvp_camera) == sizeof(Camera), "Camera layout (main.h:58-67)");
		std::memcpy(&v.camera, &cameras[ref_index], sizeof(Camera));
		v.width = depths[ref_index].cols; v.height = depths[ref_index].rows;
		v.depth = reinterpret_cast<const float*>(depth.data()); v.normal = reinterpret_cast<const float*>(normal.data());
		v.image = image.data(); v.weak = weak.data(); v.block = use_block ? block.data() : nullptr;
		v.num_src = (int32_t)src.size(); v.src_views = src.data();
		if (dvp_fusion_set_view(f, ref_index, &v) != DVP_OK) { std::fprintf(stderr, "[APD/dvp] dvp_fusion_set_view failed\n"); std::exit(EXIT_FAILURE); }
	}
	long long n = 0;
	if (dvp_fusion_run(f, &n, nullptr) != DVP_OK) { std::fprintf(stderr, "[APD/dvp] dvp_fusion_run failed\n"); std::exit(EXIT_FAILURE); }
	std::vector<float> pts((size_t)n * 6);
	if (n) dvp_fusion_get_points(f, pts.data(), 0, n);
	for (long long k = 0; k < n; ++k) {
		PointList point3D;
		point3D.coord = make_float3(pts[6 * k], pts[6 * k + 1], pts[6 * k + 2]);
		point3D.color = make_float3(pts[6 * k + 3], pts[6 * k + 4], pts[6 * k + 5]);
		PointCloud.emplace_back(point3D);
	}
	dvp_fusion_destroy(f);
}

// Replaces EdgeSegment(scale, src_img, 0, true) as GetProblemEdges calls it (main.cpp:218; APD.cpp:348-466, the
// use_canny branch): src_image is the 8-bit level image; returns the CV_8UC1 edge map (0 / 255).
inline cv::Mat DvpEdgeSegmentCanny(const cv::Mat& src_image) {
	int device = 0;
	cudaGetDevice(&device);
	std::vector<unsigned char> in((size_t)src_image.rows * src_image.cols), out(in.size());
	for (int r = 0; r < src_image.rows; ++r) std::memcpy(&in[(size_t)r * src_image.cols], src_image.ptr<unsigned char>(r), (size_t)src_image.cols);
	if (dvp_edge_segment(device, in.data(), src_image.cols, src_image.rows, out.data(), nullptr, nullptr) != DVP_OK) {
		std::fprintf(stderr, "[APD/dvp] dvp_edge_segment failed\n");
		std::exit(EXIT_FAILURE);
	}
	cv::Mat edge(src_image.rows, src_image.cols, CV_8UC1);
	for (int r = 0; r < edge.rows; ++r) std::memcpy(edge.ptr<unsigned char>(r), &out[(size_t)r * edge.cols], (size_t)edge.cols);
	return edge;
}

#endif  // DVP_APD_ADAPTER_HPP
