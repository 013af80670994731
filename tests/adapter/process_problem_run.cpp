// The reference's ProcessProblem call sequence (main.cpp:273-321) EXECUTED through include/dvp_apd_adapter.hpp:
// `APD` is the adapter class, Problem / PatchMatchParams / Camera come from the reference's own main.h, cv::Mat is the
// stub of oracle/stubs (it has storage), and the two loaders that stay the maintainer's code — InuputInitialization /
// SupportInitialization, which in the reference read jpg / cam.txt / .dmb files through OpenCV — are replaced here by
// loaders that fill THE SAME host members from raw arrays a test wrote.  Built by __graft_entry__.build() where the
// reference headers exist (tests/adapter/_build/, ships with the tree); run by tests/test_adapter.py on the GPU.
//   process_problem_run <dir>      reads <dir>/meta.txt + raw maps, writes <dir>/out_{planes,states,selected,radius}.bin
#include "main.h"           // the reference's, through -I/root/reference
#include <cstdio>
#include <string>
static unsigned long long g_seed = 0;
#define DVP_ADAPTER_SEED() (g_seed)
#include "dvp_apd_adapter.hpp"

static std::string g_dir;
static int g_w, g_h, g_s, g_has_radius, g_has_weak, g_has_selected, g_has_depths;
static float g_dmin, g_dmax;

template <typename T> static bool read_raw(const std::string& name, T* dst, size_t count) {
	FILE* f = fopen((g_dir + "/" + name).c_str(), "rb");
	if (!f) return false;
	const size_t got = fread(dst, sizeof(T), count, f);
	fclose(f);
	return got == count;
}
template <typename T> static bool write_raw(const std::string& name, const T* src, size_t count) {
	FILE* f = fopen((g_dir + "/" + name).c_str(), "wb");
	if (!f) return false;
	const size_t put = fwrite(src, sizeof(T), count, f);
	fclose(f);
	return put == count;
}
#define NEED(x) do { if (!(x)) { fprintf(stderr, "process_problem_run: %s failed\n", #x); exit(3); } } while (0)

// fills what APD.cpp:1045-1495 fills: images, depths, cameras, sizes, depth range, plane hypotheses, selected views, states
void APD::InuputInitialization() {
	width = g_w; height = g_h; num_images = g_s + 1;
	const size_t N = (size_t)width * height;
	images.clear(); depths.clear(); cameras.resize(num_images);
	std::vector<float> buf((size_t)num_images * N);
	NEED(read_raw("images.f32", buf.data(), buf.size()));
	for (int i = 0; i < num_images; ++i) { cv::Mat m(height, width, CV_32FC1); memcpy(m.ptr<float>(0), &buf[(size_t)i * N], N * 4); images.push_back(m); }
	if (g_has_depths) {
		NEED(read_raw("depths.f32", buf.data(), buf.size()));
		for (int i = 0; i < num_images; ++i) { cv::Mat m(height, width, CV_32FC1); memcpy(m.ptr<float>(0), &buf[(size_t)i * N], N * 4); depths.push_back(m); }
	}
	NEED(read_raw("cameras.bin", reinterpret_cast<unsigned char*>(cameras.data()), sizeof(Camera) * num_images));
	params_host.depth_min = g_dmin; params_host.depth_max = g_dmax; params_host.num_images = num_images;   // APD.cpp:1109-1112
	plane_hypotheses_host = new float4[N];                                                                    // APD.cpp:1208
	NEED(read_raw("planes.f32", reinterpret_cast<float*>(plane_hypotheses_host), N * 4));
	selected_views_host = cv::Mat(height, width, CV_32SC1);                                                  // zeros (APD.cpp:1425)
	if (g_has_selected) NEED(read_raw("selected.u32", selected_views_host.ptr<unsigned int>(0), N));
	weak_info_host = cv::Mat(height, width, CV_8UC1);
	if (g_has_weak) NEED(read_raw("weak.u8", weak_info_host.ptr<uchar>(0), N));
	else memset(weak_info_host.ptr<uchar>(0), STRONG, N);                                                     // APD.cpp:1196-1204
}
// fills what APD.cpp:1615-1668 fills: edge, label and radius maps
void APD::SupportInitialization() {
	const size_t N = (size_t)width * height;
	edge_host = cv::Mat(height, width, CV_8UC1);
	NEED(read_raw("edge.u8", edge_host.ptr<uchar>(0), N));
	label_host = cv::Mat(height, width, CV_32SC1);
	NEED(read_raw("label.i32", label_host.ptr<int>(0), N));
	radius_host = cv::Mat(height, width, CV_32SC1);
	if (g_has_radius) NEED(read_raw("radius.i32", radius_host.ptr<int>(0), N));
	else for (size_t i = 0; i < N; ++i) radius_host.ptr<int>(0)[i] = params_host.strong_radius;             // APD.cpp:1649-1654
}

int main(int argc, char** argv) {
	if (argc < 2) { fprintf(stderr, "usage: process_problem_run <dir>\n"); return 2; }
	g_dir = argv[1];
	Problem problem;
	int state, geom, use_apd, iters, use_detail, rotate_time, weak_peak_radius;
	float ransac;
	{
		FILE* f = fopen((g_dir + "/meta.txt").c_str(), "r");
		if (!f) return 2;
		NEED(fscanf(f, "%d %d %d %d %d %d %d %llu %f %f %d %d %d %d %d %d %f %d", &g_w, &g_h, &g_s, &iters, &state, &geom, &use_apd, &g_seed, &g_dmin, &g_dmax,
		            &g_has_radius, &g_has_weak, &g_has_selected, &g_has_depths, &use_detail, &rotate_time, &ransac, &weak_peak_radius) == 18);
		fclose(f);
	}
	problem.index = 0; problem.ref_image_id = 0; problem.scale_size = 2; problem.iteration = 0;
	for (int i = 0; i < g_s; ++i) problem.src_image_ids.push_back(i + 1);
	problem.params.max_iterations = iters; problem.params.state = (RunState)state; problem.params.geom_consistency = geom != 0;
	problem.params.use_APD = use_apd != 0; problem.params.use_detail = use_detail != 0; problem.params.rotate_time = rotate_time;
	problem.params.ransac_threshold = ransac; problem.params.weak_peak_radius = weak_peak_radius;

	// ---- main.cpp:273-280, verbatim order ----
	APD APD(problem);
	float depth_min = APD.GetDepthMin(), depth_max = APD.GetDepthMax();
	(void)depth_min; (void)depth_max;
	APD.InuputInitialization();
	APD.SupportInitialization();
	APD.CudaSpaceInitialization();
	APD.SetDataPassHelperInCuda();
	APD.RunPatchMatch();
	// ---- main.cpp:281-321: the getters ----
	const int width = APD.GetWidth(), height = APD.GetHeight();
	const size_t N = (size_t)width * height;
	std::vector<float4> planes(N);
	std::vector<unsigned int> views(N);
	for (int r = 0; r < height; ++r)
		for (int c = 0; c < width; ++c) {
			planes[(size_t)r * width + c] = APD.GetPlaneHypothesis(r, c);
			views[(size_t)r * width + c] = (unsigned int)APD.GetPixelSelectedViews(r, c);
		}
	cv::Mat pixel_states = APD.GetPixelStates(), selected = APD.GetSelectedViews(), radius = APD.GetRadiusMap();
	NEED(memcmp(selected.ptr<unsigned int>(0), views.data(), N * 4) == 0);
	NEED(write_raw("out_planes.bin", reinterpret_cast<const float*>(planes.data()), N * 4));
	NEED(write_raw("out_states.bin", pixel_states.ptr<uchar>(0), N));
	NEED(write_raw("out_selected.bin", views.data(), N));
	NEED(write_raw("out_radius.bin", radius.ptr<int>(0), N));
	printf("process_problem_run: %dx%d S=%d depth range [%g, %g]\n", width, height, g_s, APD.GetDepthMin(), APD.GetDepthMax());
	return 0;
}
