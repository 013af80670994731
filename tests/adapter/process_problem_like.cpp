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
