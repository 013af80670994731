// Compile-only check of include/dvp_apd_adapter.hpp: the call sequence of the reference's ProcessProblem
// (main.cpp:273-376) written against the adapter, with the reference's own main.h providing Problem / Camera /
// PatchMatchParams.  OpenCV/Boost come from the oracle stubs (neither is installed here).
#include "main.h"           // the reference's, through -I/root/reference
#include "dvp_apd_adapter.hpp"

// the two loaders stay the maintainer's unchanged reference code; dummies here
void APD::InuputInitialization() { width = 8; height = 8; num_images = 2; }
void APD::SupportInitialization() {}

int process_problem_like(const Problem& problem) {
	APD apd(problem);
	float depth_min = apd.GetDepthMin(), depth_max = apd.GetDepthMax();
	(void)depth_min; (void)depth_max;
	apd.InuputInitialization();
	apd.SupportInitialization();
	apd.CudaSpaceInitialization();
	apd.SetDataPassHelperInCuda();
	apd.RunPatchMatch();
	int acc = 0;
	for (int r = 0; r < apd.GetHeight(); ++r)
		for (int c = 0; c < apd.GetWidth(); ++c) {
			float4 plane = apd.GetPlaneHypothesis(r, c);
			acc += plane.w > 0 ? apd.GetPixelSelectedViews(r, c) : 0;
			apd.SetPixelSelectedViews(r, c, apd.GetPixelSelectedViews(r, c));
		}
	cv::Mat states = apd.GetPixelStates(), sel = apd.GetSelectedViews(), rad = apd.GetRadiusMap(), edge = apd.GetEdge();
	return acc + states.rows + sel.rows + rad.rows + edge.rows;
}

// the fusing loop of RunFusion (APD.cpp:1875-1957) and the edge prior of GetProblemEdges (main.cpp:218) on the device
int fusion_and_edges_like(const std::vector<Problem>& problems) {
	std::unordered_map<int, int> imageIdToindexMap;
	std::vector<cv::Mat> images, depths, normals, weaks, blocks;
	std::vector<Camera> cameras;
	for (size_t i = 0; i < problems.size(); ++i) {
		imageIdToindexMap.emplace(problems[i].ref_image_id, (int)i);
		images.emplace_back(cv::Mat(8, 8, CV_8UC1)); depths.emplace_back(cv::Mat(8, 8, CV_32FC1));
		normals.emplace_back(cv::Mat(8, 8, CV_32FC3)); weaks.emplace_back(cv::Mat(8, 8, CV_8UC1));
		cameras.emplace_back(Camera());
	}
	std::vector<PointList> PointCloud;
	DvpRunFusionLoop(problems, imageIdToindexMap, images, cameras, depths, normals, weaks, blocks, false, PointCloud);
	cv::Mat edge = DvpEdgeSegmentCanny(cv::Mat(8, 8, CV_8UC1));
	return (int)PointCloud.size() + edge.rows;
}
