// dvp_kernels_weak.cu — the adaptive-patch-deformation (WEAK pixel) path: K2 part (a) candidate offsets,
// K4 GenNeighbours, K9 RANSACToGetFitPlane, K10/K11 weak propagation
// (reference APD.cu:3746-3794, 3330-3711, 4195-4404, 2739-3125, 835-1021, 1897-2008).
#include "dvp_strong.cuh"
#include "dvp_launch.h"
#include <cfloat>

namespace dvp {

cudaError_t launch_edge_inform_prep(const KArgs& a, cudaStream_t st);  // dvp_kernels_prep.cu

// ------------------------------------------------------------------------------------------------------
// K9 RANSACToGetFitPlane (APD.cu:4195-4404).  Non-WEAK pixels: fit plane := current plane.
__global__ void __launch_bounds__(256) k_ransac_fit_nonweak(const __grid_constant__ KArgs a) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.N) return;
	if (a.weak[i] != DVP_WEAK) a.fit_planes[i] = a.planes[i];
}

cudaError_t launch_edge_inform(const KArgs& a, cudaStream_t st) {
	// part (a) (candidate offsets) is only ever read by the deformable NCC of WEAK pixels; with no WEAK
	// pixel in the view it has no reader and is skipped (the reference computes it regardless).
	if (a.weak_count > 0) return cudaErrorNotSupported;
	return launch_edge_inform_prep(a, st);
}
cudaError_t launch_gen_neighbours(const KArgs& a, cudaStream_t st) {
	if (a.weak_count > 0) return cudaErrorNotSupported;
	return cudaSuccess;
}
cudaError_t launch_ransac_fit(const KArgs& a, cudaStream_t st) {
	if (a.weak_count > 0) return cudaErrorNotSupported;
	k_ransac_fit_nonweak<<<(a.N + 255) / 256, 256, 0, st>>>(a);
	return cudaGetLastError();
}
cudaError_t launch_weak_sweep(const KArgs& a, int iter, int red, cudaStream_t st) {
	if (a.weak_count > 0) return cudaErrorNotSupported;
	return cudaSuccess;
}
cudaError_t configure_weak_kernels(int S) { return cudaSuccess; }

}  // namespace dvp
