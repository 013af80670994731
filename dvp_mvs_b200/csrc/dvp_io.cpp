// dvp_io.cpp — the reference's on-disk exchange formats (second half), host code only:
//   ReadBinMat / WriteBinMat   APD.cpp:548-573, 630-648   .bin / .dmb written with WriteBinMat: int32 version = 1, rows,
//                                                          cols, OpenCV type code, then rows * cols * elemSize bytes
//   writeDepthDmb / writeNormalDmb  APD.cpp:575-628        int32 type = 1, h, w, channels, then h * w * channels floats
//   ReadCamera                 APD.cpp:651-692             cams/<id>_cam.txt ("extrinsic" 4x4, "intrinsic" 3x3, depth range)
//   GenerateSampleList         main.cpp:127-170            pair.txt (view ids, scored source lists; score <= 0 dropped)
// With these a host program on the C ABI (tests/adapter/pipeline_main_like.cpp) can exchange every non-image file of a
// reference dense folder with the reference itself, without OpenCV or Boost.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/dvp_mvs.h"

namespace {

// bytes per element of an OpenCV type code: depth = type & 7, channels = (type >> 3) + 1
size_t cv_elem_size(int type) {
	static const size_t depth_bytes[8] = {1, 1, 2, 2, 4, 4, 8, 2};   // 8U 8S 16U 16S 32S 32F 64F 16F
	return depth_bytes[type & 7] * (size_t)((type >> 3) + 1);
}

}  // namespace

extern "C" {

int dvp_io_binmat_header(const char* path, int32_t* rows, int32_t* cols, int32_t* cv_type) {
	if (!path || !rows || !cols || !cv_type) return DVP_ERR_ARG;
	FILE* f = std::fopen(path, "rb");
	if (!f) return DVP_ERR_STATE;
	int32_t h[4] = {0, 0, 0, 0};
	const size_t got = std::fread(h, sizeof(int32_t), 4, f);
	std::fclose(f);
	if (got != 4 || h[0] != 1 || h[1] < 0 || h[2] < 0) return DVP_ERR_UNSUPPORTED;   // "Version error" (APD.cpp:562-566)
	*rows = h[1]; *cols = h[2]; *cv_type = h[3];
	return DVP_OK;
}

int dvp_io_read_binmat(const char* path, void* data, size_t bytes) {
	int32_t rows, cols, type;
	const int rc = dvp_io_binmat_header(path, &rows, &cols, &type);
	if (rc != DVP_OK) return rc;
	const size_t elem = cv_elem_size(type);
	if ((size_t)rows != 0 && (size_t)cols > SIZE_MAX / elem / (size_t)rows) return DVP_ERR_UNSUPPORTED;   // header sizes whose product overflows
	if (!data || bytes != (size_t)rows * cols * elem) return DVP_ERR_ARG;
	FILE* f = std::fopen(path, "rb");
	if (!f) return DVP_ERR_STATE;
	std::fseek(f, 16, SEEK_SET);
	const size_t got = std::fread(data, 1, bytes, f);
	std::fclose(f);
	return got == bytes ? DVP_OK : DVP_ERR_STATE;
}

int dvp_io_write_binmat(const char* path, int32_t rows, int32_t cols, int32_t cv_type, const void* data) {
	if (!path || !data || rows < 0 || cols < 0) return DVP_ERR_ARG;
	FILE* f = std::fopen(path, "wb");
	if (!f) return DVP_ERR_STATE;
	const int32_t h[4] = {1, rows, cols, cv_type};
	const size_t bytes = (size_t)rows * cols * cv_elem_size(cv_type);
	const bool ok = std::fwrite(h, sizeof(int32_t), 4, f) == 4 && std::fwrite(data, 1, bytes, f) == bytes;
	std::fclose(f);
	return ok ? DVP_OK : DVP_ERR_STATE;
}

int dvp_io_write_dmb(const char* path, int32_t rows, int32_t cols, int32_t channels, const float* data) {
	if (!path || !data || rows < 0 || cols < 0 || channels < 1) return DVP_ERR_ARG;
	FILE* f = std::fopen(path, "wb");
	if (!f) return DVP_ERR_STATE;
	const int32_t h[4] = {1, rows, cols, channels};
	const size_t count = (size_t)rows * cols * channels;
	const bool ok = std::fwrite(h, sizeof(int32_t), 4, f) == 4 && std::fwrite(data, sizeof(float), count, f) == count;
	std::fclose(f);
	return ok ? DVP_OK : DVP_ERR_STATE;
}

int dvp_io_read_camera(const char* path, dvp_camera* cam) {
	if (!path || !cam) return DVP_ERR_ARG;
	std::ifstream in(path);
	if (!in.good()) return DVP_ERR_STATE;
	std::memset(cam, 0, sizeof(*cam));
	std::string word;
	in >> word;                                                    // "extrinsic"
	for (int i = 0; i < 3; ++i) in >> cam->R[3 * i + 0] >> cam->R[3 * i + 1] >> cam->R[3 * i + 2] >> cam->t[i];
	float last_row[4];
	in >> last_row[0] >> last_row[1] >> last_row[2] >> last_row[3];
	in >> word;                                                    // "intrinsic"
	for (int i = 0; i < 3; ++i) in >> cam->K[3 * i + 0] >> cam->K[3 * i + 1] >> cam->K[3 * i + 2];
	for (int j = 0; j < 3; ++j)                                    // camera centre, accumulated in double (APD.cpp:676)
		cam->c[j] = -float(double(cam->R[0 + j]) * double(cam->t[0]) + double(cam->R[3 + j]) * double(cam->t[1]) + double(cam->R[6 + j]) * double(cam->t[2]));
	float interval = 0.f, depth_num = 0.f;
	in >> cam->depth_min >> interval >> depth_num >> cam->depth_max; // the TAT & ETH layout (APD.cpp:679-683)
	return in.fail() ? DVP_ERR_UNSUPPORTED : DVP_OK;
}

int dvp_io_read_pairs(const char* path, int32_t max_views, int32_t* num_views, int32_t* ref_ids, int32_t* num_src, int32_t* src_ids) {
	if (!path || !num_views) return DVP_ERR_ARG;
	std::ifstream file(path);
	if (!file.good()) return DVP_ERR_STATE;
	std::string line;
	std::getline(file, line);
	int n = 0;
	{ std::stringstream iss(line); iss >> n; }
	if (n < 0) return DVP_ERR_UNSUPPORTED;                         // not a pair.txt
	*num_views = n;
	if (!ref_ids || !num_src || !src_ids) return DVP_OK;           // count only
	if (n > max_views) return DVP_ERR_ARG;
	for (int i = 0; i < n; ++i) {
		std::getline(file, line);
		{ std::stringstream iss(line); iss >> ref_ids[i]; }
		std::getline(file, line);
		std::stringstream iss(line);
		int listed = 0, kept = 0;
		iss >> listed;
		for (int j = 0; j < listed; ++j) {
			int id = 0; float score = 0.f;
			iss >> id >> score;
			if (score <= 0.0f) continue;                           // main.cpp:163-165
			// a context takes the reference view + at most DVP_MAX_IMAGES - 1 sources (main.h:39 MAX_IMAGES): a view listing more
			// usable sources is reported, not silently cut (the reference would overrun its fixed-size arrays with it)
			if (kept >= DVP_MAX_IMAGES - 1) return DVP_ERR_UNSUPPORTED;
			src_ids[(size_t)i * DVP_MAX_IMAGES + kept++] = id;
		}
		num_src[i] = kept;
	}
	return DVP_OK;
}

}  // extern "C"
