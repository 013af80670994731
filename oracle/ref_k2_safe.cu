// ref_k2_safe.cu — TEST INFRASTRUCTURE ONLY (oracle/_ref/libapd_ref_k2.so, libapd_ref_k2_O1.so).
//
// ptxas 12.9 miscompiles the reference kernel GenEdgeInform (APD.cu:3731) for sm_100a at every
// optimisation level above -O0 (measured on B200, profiles/r02_reference_k2_miscompile.md):
//   -O3 (default) / -O2 : inside the per-source-view loop it emits `LDL.64 R0, [R1+0x3c0]` (SASS offset
//        +0x7750), overwriting R1 — the local-memory frame base every later LDL/STL uses — with a field
//        of the uninitialised regions[2][0]; the kernel faults with an illegal address (even for S = 1);
//   -O1 : runs, but the two horizontal rays of part (b) never find an edge (edge_neigh[2], [3] == (-1,-1)
//        for every pixel), while a brute-force evaluation of the loop and the -O0 build agree;
//   -O0 : agrees with the brute-force definition on all eight rays.
// (The source reads the uninitialised array `Point regions[12][20]`, SURVEY B17, which is probably what
// trips the optimiser.)  The same, unmodified APD.cu is therefore built a second time with
// `-Xptxas -O0` into libapd_ref_k2.so, which exports ONE launcher for kernel K2; oracle/ref_harness.cu
// calls it instead of its own copy.  Every other reference kernel runs from the build with the
// reference's own flags.  libapd_ref_k2_O1.so (same source, -Xptxas -O1) exists only so that
// `bench.py --impl reference` can time K2 at a realistic optimisation level (its horizontal rays are
// wrong, its amount of work is not); it is selected with DVP_REF_K2_LIB and never used for parity.
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <vector>
#include <string>
#include <iostream>
#include <fstream>
#include <sstream>
#include <cstdio>
#include <opencv2/opencv.hpp>
#include <boost/filesystem.hpp>
#define private public
#include <APD.cu>   // the unmodified reference translation unit, resolved through -I/root/reference
#undef private

// symbols APD.cu expects from APD.cpp (which needs a real OpenCV and cannot be compiled here)
void CudaSafeCall(const cudaError_t error, const std::string& file, const int line) {
	if (error != cudaSuccess) { fprintf(stderr, "[ref-k2] CUDA error %s at %s:%d\n", cudaGetErrorString(error), file.c_str(), line); exit(EXIT_FAILURE); }
}
void CudaCheckError(const char* file, const int line) { CudaSafeCall(cudaGetLastError(), file, line); }
bool WriteBinMat(const path&, const cv::Mat&) { return true; }
APD::APD(const Problem& p) : plane_hypotheses_host(nullptr) { params_host = p.params; problem = p; }
APD::~APD() {}

extern "C" int ref_k2_safe_launch(void* helper_dev, int width, int height) {
	dim3 grid((width + 15) / 16, (height + 15) / 16, 1), block(16, 16, 1);  // APD.cu:4412-4419
	// unoptimised code keeps every callee out of line; give the ABI call stack room (default limit is 1 KB)
	static bool stack_set = false;
	if (!stack_set) { cudaDeviceSetLimit(cudaLimitStackSize, 16 * 1024); stack_set = true; }
	GenEdgeInform<<<grid, block>>>(reinterpret_cast<DataPassHelper*>(helper_dev));
	return (int)cudaGetLastError();
}
