// oracle/ref_host.cu — TEST INFRASTRUCTURE ONLY.
// HOST functions of the reference, compiled FROM THE REFERENCE'S OWN LINES: oracle/Makefile cuts the line ranges below
// out of /root/reference/APD.cpp into oracle/_ref/src/*.inc at build time (nothing of them is committed) and this file
// wraps them in a C ABI.  They pin the CPU restatements of the callers' rows to the reference itself:
//   APD.cpp:120-346    Roberts, Label_Seek, Label_Update, Connect          -> row N1 (visibility restoration), N4 (labels)
//   APD.cpp:348-499    EdgeSegment: its own glue over cv::resize / Canny / HoughLinesP / line / threshold, which are the
//                      restated primitives of oracle/cpu (pinned against OpenCV 4.13 golden vectors)                  -> row N4, both halves
//   APD.cpp:501-546    Get3DPointonWorld, Get3DPoint, ProjectCamera          -> row N3
//   APD.cpp:548-692    ReadBinMat, writeDepthDmb, writeNormalDmb, WriteBinMat, ReadCamera  -> row N4 (on-disk formats)
//   APD.cpp:978-982, main.cpp:127-170   ToFormatIndex, GenerateSampleList (pair.txt)      -> row N4 (on-disk formats)
//   APD.cpp:1119-1140  the level-size / camera-rescale block of InuputInitialization              -> row N2 (image pyramid)
//   main.cpp:6-9, 282-363   setBit_YZL and ProcessProblem's post-pass: depth range check + visibility restoration        -> row N1
//   main.cpp:193-246, 248-265   GetProblemEdges (which image, at which scale, through which EdgeSegment mode, into which file) and ComputeRoundNum   -> rows N2 / N4
//   main.cpp:450-512   the rounds x passes x views loop of main(), ProcessProblem / GetProblemEdges recorded     -> row N2 (schedule)
//   APD.cpp:1147-1205, 1426-1493, 1615-1668   InuputInitialization's and SupportInitialization's assembly of a pass's inputs
//                      from the previous pass's files (depths.dmb, APD_normals.dmb, weak.bin, selected_views.bin, radius.bin)   -> row N2
//   APD.cpp:1773-1796  RescaleMatToTargetSize (swapped scale factors, B10)    -> row N2
//   APD.cpp:1797-1806  GetAngle                                               -> row N3
//   APD.cpp:1875-1957  the fusing loop of RunFusion (ETH version)             -> row N3
//   APD.cpp:1967-1971, 2028-2127   constants and fusing loop of RunFusion_TAT_Intermediate   -> row N3, mode 1
//   APD.cpp:2137-2138, 2195-2276   constants and fusing loop of RunFusion_TAT_advanced       -> row N3, mode 2
// The loops around them that cannot be cut out (they sit inside ProcessProblem between OpenCV and file calls,
// main.cpp:322-363) are restated here in a few lines each, citing the lines they follow.
// cv::Mat is the stub of oracle/stubs (rows, cols, at<T>(), clone()); OpenCV itself is not installed in this image.
// Built twice: -O2 (IEEE, one rounding per operation) and with the reference's own host flags
// `-O3 -ffast-math -march=native` (CMakeLists.txt:31) — the two differ where a decision sits within an ulp of a threshold.
#include "main.h"
#include <cstdint>
#include <cstring>
#include <vector>
#include <unordered_map>

#include <fstream>
#include <sstream>
#include <iomanip>
#include <iostream>
#include "_ref/src/apd_cpp_120_346.inc"
#include "_ref/src/apd_cpp_501_546.inc"
#include "_ref/src/apd_cpp_548_692.inc"
#include "_ref/src/apd_cpp_978_982.inc"
#include "_ref/src/main_cpp_127_170.inc"
#include "_ref/src/apd_cpp_1773_1796.inc"
#include "_ref/src/apd_cpp_1797_1806.inc"

struct refhost_view {   // == dvp_fusion_view (include/dvp_mvs.h)
	Camera camera;
	int32_t width, height;
	const float* depth;
	const float* normal;
	const uint8_t* image;
	const uint8_t* weak;
	const uint8_t* block;
	int32_t num_src;
	const int32_t* src_views;
};

extern "C" {

const char* refhost_flags(void) {
#ifdef __FAST_MATH__
	return "-O3 -ffast-math -march=native";
#else
	return "-O2";
#endif
}

// main.cpp:322-363 around the reference's Connect + Label_Update: per source view, the pixels that do not select the
// view are labelled; labels with fewer than 20 * (8 / scale)^2 pixels (and label 0) become "visible".
int refhost_restore_visibility(const uint32_t* selected_in, uint32_t* selected_out, int W, int H, int S, int scale_size) {
	if (!selected_in || !selected_out || W <= 0 || H <= 0 || S < 0 || S > 32 || scale_size <= 0) return 1;
	const size_t n = (size_t)W * H;
	std::memset(selected_out, 0, n * sizeof(uint32_t));
	for (int i = 0; i < S; ++i) {
		cv::Mat vis(H, W, CV_8UC1);
		for (int r = 0; r < H; ++r)
			for (int c = 0; c < W; ++c) vis.at<uchar>(r, c) = ((selected_in[(size_t)r * W + c] >> i) & 1) ? 255 : 0;   // main.cpp:310-318
		cv::Mat lab_mask(H, W, CV_32S);
		std::vector<int> label_cnt;
		Connect(vis, lab_mask, label_cnt);        // main.cpp:326
		Label_Update(lab_mask, label_cnt);        // main.cpp:327
		const int label_num = (int)label_cnt.size();
		std::vector<char> kept(label_num, 0);     // colors[j] != black  <=>  a region large enough to stay invisible (main.cpp:332-337)
		for (int j = 1; j < label_num; j++) kept[j] = !(label_cnt[j] < 20 * (8 / scale_size) * (8 / scale_size));
		for (int y = 0; y < H; y++)
			for (int x = 0; x < W; x++)
				if (!kept[lab_mask.at<int>(y, x)]) selected_out[(size_t)y * W + x] |= 1u << i;   // main.cpp:343-349, 354-361
	}
	return 0;
}

// Connect + Label_Update on a 0 / 255 image, as EdgeSegment's label mode calls them (APD.cpp:437-440): labels and counts out.
int refhost_connect_update(const uint8_t* image, int W, int H, int32_t* labels, int32_t* counts, int counts_cap, int* num_labels) {
	if (!image || !labels || W <= 0 || H <= 0) return 1;
	cv::Mat img(H, W, CV_8UC1), lab(H, W, CV_32S);
	std::memcpy(img.ptr<uchar>(0), image, (size_t)W * H);
	std::vector<int> cnt;
	Connect(img, lab, cnt);
	Label_Update(lab, cnt);
	std::memcpy(labels, lab.ptr<int>(0), (size_t)W * H * 4);
	if (num_labels) *num_labels = (int)cnt.size();
	if (counts) for (int i = 0; i < (int)cnt.size() && i < counts_cap; ++i) counts[i] = cnt[i];
	return 0;
}

int refhost_roberts(const uint8_t* image, int W, int H, uint8_t* out) {
	if (!image || !out || W <= 0 || H <= 0) return 1;
	cv::Mat img(H, W, CV_8UC1);
	std::memcpy(img.ptr<uchar>(0), image, (size_t)W * H);
	const cv::Mat r = Roberts(img);
	std::memcpy(out, r.ptr<uchar>(0), (size_t)W * H);
	return 0;
}

// RunFusion's fusing loop (APD.cpp:1875-1957) over views laid out as dvp_fusion_view; what precedes it in the
// reference (APD.cpp:1841-1873: file reads, RescaleImageAndCamera) is the caller's business here as it is in the product.
// points: [cap][6] floats; masks (may be NULL): concatenated per-view [h*w] masks after the loop.  Returns the number
// of points (which may exceed cap; only the first cap are written), or -1.
long long refhost_run_fusion(int num_views, const refhost_view* views, float* points, long long cap, uint8_t* masks_out) {
	if (num_views <= 0 || !views) return -1;
	const int num_images = num_views;
	std::vector<Problem> problems(num_views);
	std::vector<cv::Mat> images, depths, normals, masks, blocks, weaks;
	std::vector<Camera> cameras;
	std::unordered_map<int, int> imageIdToindexMap;
	bool use_block = false;
	for (int i = 0; i < num_views; ++i) use_block = use_block || views[i].block != nullptr;
	for (int i = 0; i < num_views; ++i) {
		const refhost_view& v = views[i];
		problems[i].index = i; problems[i].ref_image_id = i;
		for (int k = 0; k < v.num_src; ++k) problems[i].src_image_ids.push_back(v.src_views[k]);
		imageIdToindexMap.emplace(i, i);
		const size_t n = (size_t)v.width * v.height;
		cv::Mat image(v.height, v.width, CV_8UC3), depth(v.height, v.width, CV_32FC1), normal(v.height, v.width, CV_32FC3);
		cv::Mat mask(v.height, v.width, CV_8UC1), weak(v.height, v.width, CV_8UC1), block(v.height, v.width, CV_8UC1);
		std::memcpy(image.ptr<uchar>(0), v.image, n * 3);
		std::memcpy(depth.ptr<float>(0), v.depth, n * 4);
		std::memcpy(normal.ptr<float>(0), v.normal, n * 12);
		if (v.weak) std::memcpy(weak.ptr<uchar>(0), v.weak, n); else std::memset(weak.ptr<uchar>(0), STRONG, n);
		if (v.block) std::memcpy(block.ptr<uchar>(0), v.block, n); else std::memset(block.ptr<uchar>(0), 255, n);
		images.push_back(image); depths.push_back(depth); normals.push_back(normal); masks.push_back(mask); weaks.push_back(weak);
		blocks.push_back(block); cameras.push_back(v.camera);
	}
	std::vector<PointList> PointCloud;
	std::cout.setstate(std::ios_base::failbit);   // the loop announces every view on std::cout
	{
#include "_ref/src/apd_cpp_1875_1957.inc"
	}
	std::cout.clear();
	const long long n_pts = (long long)PointCloud.size();
	if (points)
		for (long long k = 0; k < n_pts && k < cap; ++k) {
			const PointList& p = PointCloud[(size_t)k];
			float* o = points + 6 * k;
			o[0] = p.coord.x; o[1] = p.coord.y; o[2] = p.coord.z; o[3] = p.color.x; o[4] = p.color.y; o[5] = p.color.z;
		}
	if (masks_out) {
		size_t off = 0;
		for (int i = 0; i < num_views; ++i) {
			const size_t n = (size_t)views[i].width * views[i].height;
			std::memcpy(masks_out + off, masks[i].ptr<uchar>(0), n);
			off += n;
		}
	}
	return n_pts;
}

}  // extern "C"

// The two Tanks-and-Temples variants (APD.cpp:1962-2279): same set-up, their own fusing loops and thresholds.  mode 1 =
// RunFusion_TAT_Intermediate, mode 2 = RunFusion_TAT_advanced.  Outputs as refhost_run_fusion.
namespace {
struct FusionInputs {
	std::vector<Problem> problems;
	std::vector<cv::Mat> images, depths, normals, masks, blocks, weaks;
	std::vector<Camera> cameras;
	std::unordered_map<int, int> imageIdToindexMap;
	bool use_block = false;
	FusionInputs(int num_views, const refhost_view* views) : problems(num_views) {
		for (int i = 0; i < num_views; ++i) use_block = use_block || views[i].block != nullptr;
		for (int i = 0; i < num_views; ++i) {
			const refhost_view& v = views[i];
			problems[i].index = i; problems[i].ref_image_id = i;
			for (int k = 0; k < v.num_src; ++k) problems[i].src_image_ids.push_back(v.src_views[k]);
			imageIdToindexMap.emplace(i, i);
			const size_t n = (size_t)v.width * v.height;
			cv::Mat image(v.height, v.width, CV_8UC3), depth(v.height, v.width, CV_32FC1), normal(v.height, v.width, CV_32FC3);
			cv::Mat mask(v.height, v.width, CV_8UC1), weak(v.height, v.width, CV_8UC1), block(v.height, v.width, CV_8UC1);
			std::memcpy(image.ptr<uchar>(0), v.image, n * 3);
			std::memcpy(depth.ptr<float>(0), v.depth, n * 4);
			std::memcpy(normal.ptr<float>(0), v.normal, n * 12);
			if (v.weak) std::memcpy(weak.ptr<uchar>(0), v.weak, n); else std::memset(weak.ptr<uchar>(0), STRONG, n);
			if (v.block) std::memcpy(block.ptr<uchar>(0), v.block, n); else std::memset(block.ptr<uchar>(0), 255, n);
			images.push_back(image); depths.push_back(depth); normals.push_back(normal); masks.push_back(mask); weaks.push_back(weak);
			blocks.push_back(block); cameras.push_back(v.camera);
		}
	}
};

long long export_points(const std::vector<PointList>& PointCloud, const FusionInputs& in, int num_views, const refhost_view* views, float* points, long long cap, uint8_t* masks_out) {
	const long long n_pts = (long long)PointCloud.size();
	if (points)
		for (long long k = 0; k < n_pts && k < cap; ++k) {
			const PointList& p = PointCloud[(size_t)k];
			float* o = points + 6 * k;
			o[0] = p.coord.x; o[1] = p.coord.y; o[2] = p.coord.z; o[3] = p.color.x; o[4] = p.color.y; o[5] = p.color.z;
		}
	if (masks_out) {
		size_t off = 0;
		for (int i = 0; i < num_views; ++i) {
			const size_t n = (size_t)views[i].width * views[i].height;
			std::memcpy(masks_out + off, in.masks[i].ptr<uchar>(0), n);
			off += n;
		}
	}
	return n_pts;
}

long long run_tat_intermediate(int num_views, const refhost_view* views, float* points, long long cap, uint8_t* masks_out) {
	FusionInputs in(num_views, views);
	const int num_images = num_views;
	auto& problems = in.problems; auto& images = in.images; auto& depths = in.depths; auto& normals = in.normals; auto& masks = in.masks;
	auto& blocks = in.blocks; auto& cameras = in.cameras; auto& imageIdToindexMap = in.imageIdToindexMap; const bool use_block = in.use_block;
	std::vector<PointList> PointCloud;
#include "_ref/src/apd_cpp_1967_1971.inc"
#include "_ref/src/apd_cpp_2028_2127.inc"
	return export_points(PointCloud, in, num_views, views, points, cap, masks_out);
}

long long run_tat_advanced(int num_views, const refhost_view* views, float* points, long long cap, uint8_t* masks_out) {
	FusionInputs in(num_views, views);
	const int num_images = num_views;
	auto& problems = in.problems; auto& images = in.images; auto& depths = in.depths; auto& normals = in.normals; auto& masks = in.masks;
	auto& blocks = in.blocks; auto& cameras = in.cameras; auto& imageIdToindexMap = in.imageIdToindexMap; const bool use_block = in.use_block;
	std::vector<PointList> PointCloud;
#include "_ref/src/apd_cpp_2137_2138.inc"
#include "_ref/src/apd_cpp_2195_2276.inc"
	return export_points(PointCloud, in, num_views, views, points, cap, masks_out);
}
}  // namespace

extern "C" long long refhost_run_fusion_tat(int mode, int num_views, const refhost_view* views, float* points, long long cap, uint8_t* masks_out) {
	if (num_views <= 0 || !views || (mode != 1 && mode != 2)) return -1;
	std::cout.setstate(std::ios_base::failbit);   // the loops announce every view on std::cout
	const long long n = mode == 1 ? run_tat_intermediate(num_views, views, points, cap, masks_out) : run_tat_advanced(num_views, views, points, cap, masks_out);
	std::cout.clear();
	return n;
}

// RescaleMatToTargetSize<TYPE> (APD.cpp:1773-1796) as InuputInitialization / SupportInitialization instantiate it between
// levels (APD.cpp:1162, 1178, 1437-1438, 1454, 1661): kind 0 uchar, 1 float, 2 cv::Vec3f, 3 unsigned int, 4 int.
// Pixels the reference leaves unwritten (its source index falls outside) come back as 0.
extern "C" int refhost_rescale(int kind, const void* src, int sw, int sh, void* dst, int dw, int dh) {
	if (!src || !dst || sw <= 0 || sh <= 0 || dw <= 0 || dh <= 0 || kind < 0 || kind > 4) return 1;
	const int types[5] = {CV_8UC1, CV_32FC1, CV_32FC3, CV_32SC1, CV_32SC1};
	cv::Mat in(sh, sw, types[kind]), out;
	std::memcpy(in.ptr<uchar>(0), src, (size_t)sw * sh * in.elem_size);
	out = in;
	switch (kind) {
	case 0: RescaleMatToTargetSize<uchar>(in, out, cv::Size2i(dw, dh)); break;
	case 1: RescaleMatToTargetSize<float>(in, out, cv::Size2i(dw, dh)); break;
	case 2: RescaleMatToTargetSize<cv::Vec3f>(in, out, cv::Size2i(dw, dh)); break;
	case 3: RescaleMatToTargetSize<unsigned int>(in, out, cv::Size2i(dw, dh)); break;
	default: RescaleMatToTargetSize<int>(in, out, cv::Size2i(dw, dh)); break;
	}
	if (out.rows != dh || out.cols != dw) return 2;
	std::memcpy(dst, out.ptr<uchar>(0), (size_t)dw * dh * out.elem_size);
	return 0;
}

// ---- the reference's own file readers / writers (row N4, on-disk formats) -----------------------------------------
// kind as in refhost_rescale (0 uchar, 1 float, 2 Vec3f, 3 unsigned int, 4 int).
extern "C" int refhost_write_binmat(const char* file, int kind, const void* src, int w, int h) {
	if (!file || !src || kind < 0 || kind > 4) return 1;
	const int types[5] = {CV_8UC1, CV_32FC1, CV_32FC3, CV_32SC1, CV_32SC1};
	cv::Mat m(h, w, types[kind]);
	std::memcpy(m.ptr<uchar>(0), src, (size_t)w * h * m.elem_size);
	return WriteBinMat(path(file), m) ? 0 : 2;
}
// -> rows, cols, OpenCV type code; data copied when `dst` is given (capacity in bytes)
extern "C" int refhost_read_binmat(const char* file, int* rows, int* cols, int* type, void* dst, long long cap) {
	if (!file) return 1;
	cv::Mat m;
	std::cerr.setstate(std::ios_base::failbit);
	const bool ok = ReadBinMat(path(file), m);
	std::cerr.clear();
	if (!ok) return 2;
	if (rows) *rows = m.rows; if (cols) *cols = m.cols; if (type) *type = m.type();
	const long long n = (long long)m.rows * m.cols * (long long)m.elem_size;
	if (dst) { if (n > cap) return 3; std::memcpy(dst, m.ptr<uchar>(0), (size_t)n); }
	return 0;
}
extern "C" int refhost_write_dmb(const char* file, int channels, const float* src, int w, int h) {
	if (!file || !src || (channels != 1 && channels != 3)) return 1;
	cv::Mat m(h, w, channels == 1 ? CV_32FC1 : CV_32FC3);
	std::memcpy(m.ptr<uchar>(0), src, (size_t)w * h * m.elem_size);
	if (channels == 1) return writeDepthDmb(path(file), cv::Mat_<float>(m));
	return writeNormalDmb(path(file), cv::Mat_<cv::Vec3f>(m));
}
extern "C" int refhost_read_camera(const char* file, Camera* cam) {
	if (!file || !cam) return 1;
	std::memset(cam, 0, sizeof(Camera));
	return ReadCamera(path(file), *cam) ? 0 : 2;
}
// pair.txt: -> number of problems; ref ids and source lists (cap_src entries per problem, -1 padded) when given
extern "C" int refhost_read_pairs(const char* dense_folder, int* ref_ids, int* num_src, int* src_ids, int cap_problems, int cap_src) {
	if (!dense_folder) return -1;
	std::vector<Problem> problems;
	GenerateSampleList(path(dense_folder), problems);
	const int n = (int)problems.size();
	for (int i = 0; i < n && i < cap_problems; ++i) {
		if (ref_ids) ref_ids[i] = problems[i].ref_image_id;
		const int ns = (int)problems[i].src_image_ids.size();
		if (num_src) num_src[i] = ns;
		if (src_ids) for (int k = 0; k < cap_src; ++k) src_ids[(size_t)i * cap_src + k] = k < ns ? problems[i].src_image_ids[k] : -1;
	}
	return n;
}

// InuputInitialization's "scale images" block (APD.cpp:1119-1143): level size = round(size / scale_size) in float, the
// intrinsics scaled by the ratios of the ROUNDED sizes.  cv::resize is stood in for by an allocation of the target size
// (the pixels are row N2's image pyramid, pinned against OpenCV itself in tests/test_image.py); what is checked here is
// the size and camera arithmetic.
namespace cv {
enum { INTER_LINEAR = 1, IMREAD_GRAYSCALE = 0 };
// cv::resize(.., INTER_LINEAR) for CV_32FC1: OpenCV's generic path (imgproc/src/resize.cpp, resizeGeneric_ with HResizeLinear /
// VResizeLinear) as oracle/image_oracle.py restates it and tests/golden/resize_f32.npz pins it against OpenCV 4.13 without IPP:
// fx = (float)((dx + 0.5) * scale - 0.5) with scale = 1 / (dst / src) in double, taps clamped (columns: weight zeroed at the
// borders; rows: clipped), a row pass then a column pass, every product and sum rounded to float (this file is built with
// -ffp-contract=off).  The level-size block above only needs the size of the result.
inline void resize(const Mat& src_in, Mat& dst, Size dsize, double, double, int) {
	const Mat src = src_in;
	const int sw = src.cols, sh = src.rows, dw = dsize.width, dh = dsize.height;
	Mat out(dh, dw, CV_32FC1);
	if (src.type() != CV_32FC1 || sw <= 0 || sh <= 0) { dst = out; return; }
	std::vector<int> x0(dw), x1(dw); std::vector<float> fx(dw);
	const double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
	for (int dx = 0; dx < dw; ++dx) {
		float f = (float)((dx + 0.5) * scale_x - 0.5);
		int s = (int)std::floor(f); f -= (float)s;
		if (s < 0) { f = 0.f; s = 0; }
		if (s >= sw - 1) { f = 0.f; s = sw - 1; }
		x0[dx] = s; x1[dx] = s + 1 < sw ? s + 1 : sw - 1; fx[dx] = f;
	}
	std::vector<float> rows((size_t)sh * dw);
	for (int y = 0; y < sh; ++y) {
		const float* r = src.ptr<float>(y);
		for (int dx = 0; dx < dw; ++dx) { const float a0 = 1.0f - fx[dx]; const float p0 = r[x0[dx]] * a0, p1 = r[x1[dx]] * fx[dx]; rows[(size_t)y * dw + dx] = p0 + p1; }
	}
	for (int dy = 0; dy < dh; ++dy) {
		float f = (float)((dy + 0.5) * scale_y - 0.5);
		const int s = (int)std::floor(f); f -= (float)s;
		const int y0 = s < 0 ? 0 : (s > sh - 1 ? sh - 1 : s), y1 = s + 1 < 0 ? 0 : (s + 1 > sh - 1 ? sh - 1 : s + 1);
		const float b0 = 1.0f - f;
		float* o = out.ptr<float>(dy);
		for (int dx = 0; dx < dw; ++dx) { const float p0 = rows[(size_t)y0 * dw + dx] * b0, p1 = rows[(size_t)y1 * dw + dx] * f; o[dx] = p0 + p1; }
	}
	dst = out;
}
}
extern "C" int refhost_level_camera(const Camera* full, int full_w, int full_h, int scale_size, Camera* out, int* out_w, int* out_h) {
	if (!full || !out || full_w <= 0 || full_h <= 0 || scale_size < 1) return 1;
	struct { int scale_size; } problem = {scale_size};
	int num_images = 1, width = full_w, height = full_h;
	std::vector<cv::Mat> images(1, cv::Mat(full_h, full_w, CV_32FC1));
	std::vector<Camera> cameras(1, *full);
	cameras[0].width = full_w; cameras[0].height = full_h;   // APD.cpp:1087-1088, 1103-1104
	std::cout.setstate(std::ios_base::failbit);
	{
#include "_ref/src/apd_cpp_1119_1143.inc"
	}
	std::cout.clear();
	*out = cameras[0];
	if (out_w) *out_w = width; if (out_h) *out_h = height;
	return 0;
}

// main()'s schedule (main.cpp:450-512), compiled as it stands: ProcessProblem and GetProblemEdges are recorders here, so
// the loop yields, call by call, which view runs at which scale with which parameters — what dvp_scene_pass_params and the
// scene driver's order restate.  A record = 4 ints (view, iteration, scale_size, edges_requested) + the PatchMatchParams.
namespace {
struct ScheduleRecord { int view, iteration, scale_size, edges; PatchMatchParams params; };
std::vector<ScheduleRecord>* g_schedule = nullptr;
int g_edges_seen = 0;
void GetProblemEdges(const Problem&) { g_edges_seen = 1; }
void ProcessProblem(const Problem& p) {
	if (g_schedule) g_schedule->push_back({p.ref_image_id, p.iteration, p.scale_size, g_edges_seen, p.params});
	g_edges_seen = 0;
}
}  // namespace
// out: [cap][27] ints / floats-as-bits in the order of dvp_params after the 4 record ints; returns the number of calls
extern "C" int refhost_schedule(int round_num, int num_problems, int32_t* out, int cap) {
	if (round_num < 2 || num_problems < 1) return -1;
	std::vector<Problem> problems(num_problems);
	for (int i = 0; i < num_problems; ++i) { problems[i].index = i; problems[i].ref_image_id = i; }
	std::vector<ScheduleRecord> log;
	g_schedule = &log; g_edges_seen = 0;
	int iteration_index = 0;
	bool flag = true;
	std::cout.setstate(std::ios_base::failbit);
#include "_ref/src/main_cpp_450_512.inc"
	std::cout.clear();
	g_schedule = nullptr;
	(void)iteration_index;
	for (int k = 0; k < (int)log.size() && k < cap; ++k) {
		const ScheduleRecord& r = log[k];
		const PatchMatchParams& q = r.params;
		int32_t* o = out + (size_t)k * 27;
		auto fbits = [](float f) { int32_t b; std::memcpy(&b, &f, 4); return b; };
		o[0] = r.view; o[1] = r.iteration; o[2] = r.scale_size; o[3] = r.edges;
		o[4] = q.max_iterations; o[5] = q.num_images; o[6] = fbits(q.sigma_spatial); o[7] = fbits(q.sigma_color); o[8] = q.top_k;
		o[9] = fbits(q.depth_min); o[10] = fbits(q.depth_max); o[11] = q.geom_consistency; o[12] = q.strong_radius; o[13] = q.strong_increment;
		o[14] = q.weak_radius; o[15] = q.weak_increment; o[16] = q.use_APD; o[17] = q.use_edge; o[18] = q.use_limit; o[19] = q.use_label;
		o[20] = q.use_detail; o[21] = q.use_radius; o[22] = q.weak_peak_radius; o[23] = q.rotate_time; o[24] = fbits(q.ransac_threshold);
		o[25] = fbits(q.geom_factor); o[26] = (int32_t)q.state;
	}
	return (int)log.size();
}

// ProcessProblem after RunPatchMatch (main.cpp:282-363), compiled as it stands against an object that answers the APD
// getters it calls: out-of-range depths are zeroed and their pixels become UNKNOWN; per source view the pixels that do not
// select the view are labelled (Connect + Label_Update) and regions below 20 * (8 / scale)^2 pixels are given the view.
// (The reference draws the kept regions' display colours with rand(): a region whose three draws all give 0 would count as
// small — probability 2^-24 per region, and the same for anyone who runs the reference; srand(1) fixes the draws here.)
#include "_ref/src/main_cpp_6_9.inc"
namespace {
struct PostPassAPD {
	int w, h; float dmin, dmax;
	const float4* planes; cv::Mat states; const uint32_t* sel_in; uint32_t* sel_out;
	int GetWidth() const { return w; }
	int GetHeight() const { return h; }
	float GetDepthMin() const { return dmin; }
	float GetDepthMax() const { return dmax; }
	cv::Mat GetPixelStates() const { return states; }
	float4 GetPlaneHypothesis(int r, int c) const { return planes[(size_t)r * w + c]; }
	unsigned int GetPixelSelectedViews(int r, int c) const { return sel_in[(size_t)r * w + c]; }
	void SetPixelSelectedViews(int r, int c, unsigned int v) { sel_out[(size_t)r * w + c] = v; }
};
}  // namespace
extern "C" int refhost_post_pass(const float* planes /*[h][w][4]*/, const uint8_t* states_in, const uint32_t* selected_in, int w, int h, int num_src, int scale_size,
                                 float depth_min, float depth_max, float* depth_out, uint8_t* states_out, uint32_t* selected_out) {
	if (!planes || !states_in || !selected_in || !depth_out || !states_out || !selected_out || w <= 0 || h <= 0 || num_src < 0 || num_src > 32 || scale_size < 1) return 1;
	struct { std::vector<int> src_image_ids; int scale_size; } problem;
	problem.src_image_ids.assign(num_src, 0); problem.scale_size = scale_size;
	PostPassAPD APD{w, h, depth_min, depth_max, reinterpret_cast<const float4*>(planes), cv::Mat(h, w, CV_8UC1), selected_in, selected_out};
	std::memcpy(APD.states.ptr<uchar>(0), states_in, (size_t)w * h);
	srand(1);
#include "_ref/src/main_cpp_282_363.inc"
	std::memcpy(depth_out, depth.ptr<float>(0), (size_t)w * h * 4);
	std::memcpy(states_out, pixel_states.ptr<uchar>(0), (size_t)w * h);
	return 0;
}

// ---- EdgeSegment (APD.cpp:348-499) from the reference's own lines ---------------------------------------------------
// OpenCV is not in this image: the five OpenCV calls EdgeSegment makes are forwarded to the restatements of oracle/cpu
// (libapd_cpu.so), each of which is pinned against real OpenCV 4.13 output (tests/golden/edge_canny.npz, label_segment.npz,
// tests/test_labels_oracle.py).  Everything ELSE in the function — thresholds from the histogram median, weak_tex_num, the
// region loop, border extraction, line drawing order, the border clean-up, the final labelling rule — is the reference's code.
extern "C" {
void edge_cpu_canny(const uint8_t* img, int cols, int rows, double low, double high, uint8_t* dst);
void label_cpu_resize8u(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh);
void label_cpu_line(uint8_t* img, int w, int h, int x0, int y0, int x1, int y1, int value);
int hough_cpu_lines_p(const uint8_t* image, int width, int height, float rho, float theta, int threshold, int line_length, int line_gap, int* lines, int max_lines);
}
namespace cv {
inline void resize8(const Mat& src_in, Mat& dst, Size dsize) {
	const Mat src = src_in;   // src and dst may be the same object (APD.cpp:361, 439, 445)
	Mat out(dsize.height, dsize.width, CV_8UC1);
	if (src.cols == dsize.width && src.rows == dsize.height) out = src;   // cv::resize to the same size copies
	else label_cpu_resize8u(src.ptr<uchar>(0), src.cols, src.rows, out.ptr<uchar>(0), dsize.width, dsize.height);
	dst = out;
}
#define DVP_REF_HOST_RESIZE8 1
inline void threshold(const Mat& src, Mat& dst, double thresh, double maxval, int) {
	Mat out = src;
	for (size_t i = 0; i < out.storage.size(); ++i) out.storage[i] = out.storage[i] > thresh ? (unsigned char)maxval : 0;
	dst = out;
}
inline void Canny(const Mat& image, Mat& edges, double t1, double t2, int, bool) {
	Mat out(image.rows, image.cols, CV_8UC1);
	edge_cpu_canny(image.ptr<uchar>(0), image.cols, image.rows, t1, t2, out.ptr<uchar>(0));
	edges = out;
}
inline void HoughLinesP(const Mat& image, std::vector<Vec4i>& lines, double rho, double theta, int threshold, double min_len, double max_gap) {
	std::vector<int> buf(4 * 65536);
	const int n = hough_cpu_lines_p(image.ptr<uchar>(0), image.cols, image.rows, (float)rho, (float)theta, threshold, (int)min_len, (int)max_gap, buf.data(), 65536);
	lines.clear();
	for (int i = 0; i < n && i < 65536; ++i) lines.push_back(Vec4i{{buf[4 * i], buf[4 * i + 1], buf[4 * i + 2], buf[4 * i + 3]}});
}
inline void line(Mat& img, Point a, Point b, const Scalar& color, int) {
	label_cpu_line(img.ptr<uchar>(0), img.cols, img.rows, a.x, a.y, b.x, b.y, (int)color.val[0]);
}
}  // namespace cv
namespace edge_segment_ref {
// the cut lines call cv::resize(src, dst, Size, 0, 0, INTER_LINEAR) on 8-bit images: route it to the 8-bit restatement
// (the float stand-in above serves the level-size block only)
namespace cv_shadow {}
#define resize(src, dst, size, fx, fy, interp) resize8(src, dst, size)
#include "_ref/src/apd_cpp_348_499.inc"
#undef resize
}  // namespace edge_segment_ref

// mode 0 (use_canny = true): edge map [rows][cols] u8.  mode 1 (use_canny = false): label map at the level size, int32.
extern "C" int refhost_edge_segment(int scale, const uint8_t* image, int cols, int rows, int mode, int use_canny, void* out, int* out_cols, int* out_rows) {
	if (!image || !out || cols < 4 || rows < 4 || (mode != 0 && mode != 1)) return 1;
	cv::Mat src(rows, cols, CV_8UC1);
	std::memcpy(src.ptr<uchar>(0), image, (size_t)cols * rows);
	srand(1);
	const cv::Mat r = edge_segment_ref::EdgeSegment(scale, src, mode, use_canny != 0);
	if (out_cols) *out_cols = r.cols; if (out_rows) *out_rows = r.rows;
	std::memcpy(out, r.ptr<uchar>(0), (size_t)r.cols * r.rows * r.elem_size);
	return 0;
}

// ---- how a pass's inputs are assembled from the previous pass's files (row N2) ------------------------------------------
// InuputInitialization (APD.cpp:1147-1205: depth maps for the geometric term, pixel states; 1426-1493: plane hypotheses from
// depths.dmb + APD_normals.dmb, selected views) and SupportInitialization (APD.cpp:1615-1668: edge map, label map, radius map)
// compiled from the reference's lines as members of a class that has exactly the members those lines touch.  The files are
// real files in a temporary dense folder, read by the reference's own ReadBinMat.
namespace assembly_ref {
class APD {
public:
	Problem problem;
	PatchMatchParams params_host;
	int width = 0, height = 0, weak_count = 0, num_images = 0;
	std::vector<cv::Mat> depths;
	std::vector<Camera> cameras;
	cv::Mat weak_info_host, neighbours_map_host, selected_views_host, edge_host, label_host, radius_host;
	float4* plane_hypotheses_host = nullptr;
	~APD() { delete[] plane_hypotheses_host; }
	void Assemble();
	void SupportInitialization();
};
void APD::Assemble() {
#include "_ref/src/apd_cpp_1147_1205.inc"
	plane_hypotheses_host = new float4[(size_t)width * height]();   // APD.cpp:1209 (the FIRST_INIT prior that follows is out of scope)
#include "_ref/src/apd_cpp_1426_1493.inc"
}
#include "_ref/src/apd_cpp_1615_1668.inc"
}  // namespace assembly_ref

// outputs (any may be NULL): depths [(1+S)][h][w] f32 (geom only), weak [h][w] u8, planes [h][w][4] f32 (x, y, z, w),
// selected [h][w] u32, radius [h][w] i32, edge [eh][ew] u8 as read (its size in edge_wh), weak_count
// flags: state, geom_consistency, use_APD, use_edge, use_limit, use_label, use_radius, strong_radius (PatchMatchParams fields)
extern "C" int refhost_assemble(const char* dense_folder, int ref_id, int num_src, const int* src_ids, int scale_size, const int* flags,
                                int width, int height, float* depths, uint8_t* weak, float* planes, uint32_t* selected, int32_t* radius,
                                uint8_t* edge, int* edge_wh, int* weak_count) {
	if (!dense_folder || !flags || width <= 0 || height <= 0 || num_src < 0) return 1;
	PatchMatchParams prm;
	prm.state = (RunState)flags[0]; prm.geom_consistency = flags[1] != 0; prm.use_APD = flags[2] != 0; prm.use_edge = flags[3] != 0;
	prm.use_limit = flags[4] != 0; prm.use_label = flags[5] != 0; prm.use_radius = flags[6] != 0; prm.strong_radius = flags[7];
	const PatchMatchParams* params = &prm;
	assembly_ref::APD a;
	a.problem.index = 0; a.problem.ref_image_id = ref_id; a.problem.scale_size = scale_size; a.problem.params = *params;
	for (int k = 0; k < num_src; ++k) a.problem.src_image_ids.push_back(src_ids[k]);
	a.problem.dense_folder = path(dense_folder);
	a.problem.result_folder = path(dense_folder) / path("APD") / path(ToFormatIndex(ref_id));
	a.params_host = *params;
	a.width = width; a.height = height; a.num_images = num_src + 1;
	a.cameras.resize(num_src + 1);
	for (auto& c : a.cameras) { c.width = width; c.height = height; }
	std::cout.setstate(std::ios_base::failbit); std::cerr.setstate(std::ios_base::failbit);
	a.Assemble();
	a.SupportInitialization();
	std::cout.clear(); std::cerr.clear();
	const size_t n = (size_t)width * height;
	if (depths && params->geom_consistency)
		for (size_t k = 0; k < a.depths.size(); ++k) {
			if (a.depths[k].cols != width || a.depths[k].rows != height) return 2;
			std::memcpy(depths + k * n, a.depths[k].ptr<float>(0), n * 4);
		}
	if (weak) { if (a.weak_info_host.cols != width || a.weak_info_host.rows != height) return 3; std::memcpy(weak, a.weak_info_host.ptr<uchar>(0), n); }
	if (planes) std::memcpy(planes, a.plane_hypotheses_host, n * 16);
	if (selected) { if (a.selected_views_host.cols != width || a.selected_views_host.rows != height) return 4; std::memcpy(selected, a.selected_views_host.ptr<uchar>(0), n * 4); }
	if (radius && params->use_radius) { if (a.radius_host.cols != width || a.radius_host.rows != height) return 5; std::memcpy(radius, a.radius_host.ptr<uchar>(0), n * 4); }
	if (edge_wh) { edge_wh[0] = a.edge_host.cols; edge_wh[1] = a.edge_host.rows; }
	if (edge && !a.edge_host.empty() && a.edge_host.cols == width && a.edge_host.rows == height) std::memcpy(edge, a.edge_host.ptr<uchar>(0), n);
	if (weak_count) *weak_count = a.weak_count;
	return 0;
}

// ---- GetProblemEdges (main.cpp:193-246) and ComputeRoundNum (main.cpp:248-265) from the reference's own lines --------------
// cv::imread hands back the image the caller supplied; cv::imwrite is a no-op (show_medium_result is switched off).
namespace cv {
static Mat g_imread_image;
inline Mat imread(const std::string&, int = 1) { return g_imread_image; }
inline bool imwrite(const std::string&, const Mat&) { return true; }
}
namespace problem_edges_ref {
cv::Mat EdgeSegment(const int scale, const cv::Mat& srcImage, int mode = 0, bool useCanny = false) { return edge_segment_ref::EdgeSegment(scale, srcImage, mode, useCanny); }
#include "_ref/src/main_cpp_193_246.inc"
#include "_ref/src/main_cpp_248_265.inc"
}  // namespace problem_edges_ref
#include <chrono>
extern "C" int refhost_get_problem_edges(const char* dense_folder, int ref_id, int scale_size, const uint8_t* image, int cols, int rows) {
	if (!dense_folder || !image || cols < 16 || rows < 16 || scale_size < 1) return 1;
	cv::g_imread_image = cv::Mat(rows, cols, CV_8UC1);
	std::memcpy(cv::g_imread_image.ptr<uchar>(0), image, (size_t)cols * rows);
	Problem problem;
	problem.index = 0; problem.ref_image_id = ref_id; problem.scale_size = scale_size; problem.show_medium_result = false;
	problem.dense_folder = path(dense_folder);
	problem.result_folder = path(dense_folder) / path("APD") / path(ToFormatIndex(ref_id));
	srand(1);
	std::cout.setstate(std::ios_base::failbit);
	problem_edges_ref::GetProblemEdges(problem);
	std::cout.clear();
	return 0;
}
// ComputeRoundNum over `n` problems whose images all have the given size (it reads the first image's size through cv::imread)
extern "C" int refhost_compute_round_num(const char* dense_folder, int n, int cols, int rows) {
	cv::g_imread_image = cv::Mat(rows, cols, CV_8UC1);
	std::vector<Problem> problems((size_t)(n > 0 ? n : 0));
	for (int i = 0; i < n; ++i) { problems[i].index = i; problems[i].ref_image_id = i; problems[i].dense_folder = path(dense_folder ? dense_folder : "."); }
	std::cout.setstate(std::ios_base::failbit);
	const int r = problem_edges_ref::ComputeRoundNum(problems);
	std::cout.clear();
	return r;
}
