// The whole of the reference's main() (main.cpp:421-520) as a host C++ program on the C ABI — every device-side row in
// one place: GetProblemEdges' edge prior (row N4, main.cpp:443-447), the rounds x passes x views schedule with its
// visibility restoration (rows N1 + N2, main.cpp:449-511) and RunFusion (row N3, main.cpp:514), ending in the PLY the
// reference writes.  What a dense folder provides (cams, pair.txt, the image pyramid, label maps, the FIRST_INIT prior)
// is read from flat binary files that tests/test_adapter.py dumps from a synthetic scene:
//   <dir>/meta.txt                      num_views num_levels full_w full_h iterations seed
//   <dir>/view<v>.cam                   dvp_camera (112 bytes), then int32 num_src, then int32 src[num_src]
//   <dir>/view<v>.gray                  uint8 [full_h][full_w]: the grey image.  If present it is ALL a view needs besides camera
//                                       and prior: pyramid, edge maps and label maps are built on the device (dvp_scene_set_image);
//   <dir>/view<v>_level<l>.image|.label otherwise: float32 / int32 [h][w] at the level's size (dvp_scene_level_size)
//   <dir>/view<v>.planes                float32 [h0][w0][4]
//   <dir>/view<v>.color                 uint8 [h][w][3] at the finest level
// usage: pipeline_main_like <dir> <out.ply> [device]
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "dvp_mvs.h"

static bool read_file(const std::string& path, void* dst, size_t bytes) {
	FILE* f = std::fopen(path.c_str(), "rb");
	if (!f) { std::fprintf(stderr, "cannot open %s\n", path.c_str()); return false; }
	const size_t got = std::fread(dst, 1, bytes, f);
	std::fclose(f);
	if (got != bytes) { std::fprintf(stderr, "%s: %zu of %zu bytes\n", path.c_str(), got, bytes); return false; }
	return true;
}

#define CHECK(call) do { const int rc_ = (call); if (rc_ != DVP_OK) { std::fprintf(stderr, "%s -> %d\n", #call, rc_); return 10; } } while (0)

int main(int argc, char** argv) {
	if (argc < 3) { std::fprintf(stderr, "usage: %s <dir> <out.ply> [device]\n", argv[0]); return 1; }
	const std::string dir = argv[1];
	const int device = argc > 3 ? std::atoi(argv[3]) : 0;
	int num_views = 0, num_levels = 0, full_w = 0, full_h = 0, iterations = 3;
	unsigned long long seed = 0;
	{
		FILE* f = std::fopen((dir + "/meta.txt").c_str(), "r");
		if (!f || std::fscanf(f, "%d %d %d %d %d %llu", &num_views, &num_levels, &full_w, &full_h, &iterations, &seed) != 6) return 2;
		std::fclose(f);
	}
	dvp_scene* sc = dvp_scene_create(device, num_views, num_levels);
	if (!sc) return 3;
	CHECK(dvp_scene_set_max_iterations(sc, iterations));
	std::vector<std::vector<unsigned char>> colors(num_views);
	for (int v = 0; v < num_views; ++v) {
		const std::string base = dir + "/view" + std::to_string(v);
		std::vector<unsigned char> cam(sizeof(dvp_camera) + 4 + 4 * DVP_MAX_IMAGES, 0);
		FILE* f = std::fopen((base + ".cam").c_str(), "rb");
		if (!f) return 4;
		const size_t got = std::fread(cam.data(), 1, cam.size(), f);
		std::fclose(f);
		if (got < sizeof(dvp_camera) + 4) return 4;
		const dvp_camera* camera = reinterpret_cast<const dvp_camera*>(cam.data());
		const int num_src = *reinterpret_cast<const int*>(cam.data() + sizeof(dvp_camera));
		const int* src = reinterpret_cast<const int*>(cam.data() + sizeof(dvp_camera) + 4);
		CHECK(dvp_scene_set_view(sc, v, camera, full_w, full_h, num_src, src));                    // cam.txt, pair.txt
		int w = 0, h = 0;
		bool from_image = false;
		if (FILE* g = std::fopen((base + ".gray").c_str(), "rb")) {
			std::fclose(g);
			std::vector<unsigned char> gray((size_t)full_w * full_h);
			if (!read_file(base + ".gray", gray.data(), gray.size())) return 5;
			CHECK(dvp_scene_set_image(sc, v, gray.data(), /*edges + labels*/ 3));                      // InuputInitialization's pyramid + GetProblemEdges
			from_image = true;
		}
		for (int level = 0; level < num_levels; ++level) {
			CHECK(dvp_scene_level_size(sc, full_w, full_h, level, &w, &h));
			if (!from_image) {
				std::vector<float> image((size_t)w * h);
				std::vector<int32_t> label((size_t)w * h);
				if (!read_file(base + "_level" + std::to_string(level) + ".image", image.data(), image.size() * 4)) return 5;
				if (!read_file(base + "_level" + std::to_string(level) + ".label", label.data(), label.size() * 4)) return 5;
				CHECK(dvp_scene_set_level(sc, v, level, image.data(), /*edge=*/nullptr, label.data()));
				CHECK(dvp_scene_compute_edges(sc, v, level));                                         // GetProblemEdges, main.cpp:443-447
			}
			if (level == 0) {
				std::vector<float> planes((size_t)w * h * 4);
				if (!read_file(base + ".planes", planes.data(), planes.size() * 4)) return 6;
				CHECK(dvp_scene_set_initial_planes(sc, v, planes.data()));
			}
		}
		colors[v].resize((size_t)w * h * 3);                                                       // (w, h) = finest level here
		if (!read_file(base + ".color", colors[v].data(), colors[v].size())) return 7;
	}
	float schedule_ms = 0.f, fusion_ms = 0.f;
	CHECK(dvp_scene_run(sc, seed, &schedule_ms));                                                 // main.cpp:449-511
	dvp_fusion* fu = dvp_fusion_create(device, num_views);
	if (!fu) return 8;
	std::vector<const uint8_t*> images(num_views);
	for (int v = 0; v < num_views; ++v) images[v] = colors[v].data();
	CHECK(dvp_scene_fuse_views(sc, fu, images.data(), nullptr));                                  // depths.dmb / normals / weak.bin stay in HBM
	long long points = 0;
	CHECK(dvp_fusion_run(fu, &points, &fusion_ms));                                               // RunFusion, main.cpp:514
	CHECK(dvp_fusion_write_ply(fu, argv[2]));
	std::printf("views %d levels %d schedule %.3f ms fusion %.3f ms points %lld\n", num_views, num_levels, schedule_ms, fusion_ms, points);
	dvp_fusion_destroy(fu);
	dvp_scene_destroy(sc);
	return 0;
}
