// The reference's main() (main.cpp:421-520) written against the dvp_scene_* entry points (row N2): what replaces the
// rounds x passes x ProcessProblem loop and its file round trips.  Built and run by tests/test_adapter.py with a
// host-only scene (device -1): it prints the schedule the library would execute — the run itself needs a GPU.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "dvp_mvs.h"

int main(int argc, char** argv) {
	const int device = argc > 1 ? std::atoi(argv[1]) : -1;
	const int cols = 6221, rows = 4146;                  // ETH3D
	// ComputeRoundNum (main.cpp:248-264): halve the larger side until it is <= 800
	int max_size = cols > rows ? cols : rows, round_num = 1;
	while (max_size > 800) { max_size /= 2; round_num++; }
	const int num_views = 3, num_levels = round_num - 1; // rounds 0 .. round_num-2 run at scale 2^(round_num-1-i)
	dvp_scene* sc = dvp_scene_create(device, num_views, num_levels);
	if (!sc) { std::fprintf(stderr, "dvp_scene_create failed\n"); return 1; }
	std::printf("round_num %d -> %d pyramid levels\n", round_num, num_levels);
	for (int level = 0; level < num_levels; ++level) {
		int w = 0, h = 0;
		if (dvp_scene_level_size(sc, cols, rows, level, &w, &h) != DVP_OK) return 2;
		for (int pass = 0; pass < 4; ++pass) {
			dvp_params p;
			if (dvp_scene_pass_params(sc, level, pass, &p) != DVP_OK) return 3;
			std::printf("level %d %dx%d pass %d: state %d use_APD %d geom %d weak_peak_radius %d rotate_time %d ransac %.5f use_detail %d\n",
			            level, w, h, pass, p.state, p.use_APD, p.geom_consistency, p.weak_peak_radius, p.rotate_time, p.ransac_threshold, p.use_detail);
		}
	}
	if (device >= 0) {
		// with a GPU: per view  dvp_scene_set_view(cam.txt, pair.txt) ; per level dvp_scene_set_level(resized grey image,
		// edges_<level>.dmb, labels) ; dvp_scene_set_initial_planes(prior) ; then the whole of main.cpp:452-511 is
		float ms = 0.f;
		const int rc = dvp_scene_run(sc, /*seed=*/1, &ms);
		std::printf("dvp_scene_run -> %d (%.1f ms)\n", rc, ms);   // DVP_ERR_STATE here: no views were configured
	} else {
		if (dvp_scene_run(sc, 1, nullptr) != DVP_ERR_STATE) return 4;   // a host-only scene refuses to run
	}
	dvp_scene_destroy(sc);
	return 0;
}
