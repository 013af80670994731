// oracle/cpu/label_cpu.cpp — TEST INFRASTRUCTURE ONLY.
// Groundwork for the label half of row N4 (no device path uses it yet): CPU restatement of
// EdgeSegment(scale, image, mode 1, use_canny = false), reference APD.cpp:348-402 + 437-499, as GetProblemEdges calls it
// on the full-resolution 8-bit image to produce labels_<scale>.dmb (main.cpp:234) — the `label_cuda` input of the hot path:
//   two cv::resize halvings -> Roberts (APD.cpp:120-136) -> threshold -> Connect / Label_Update -> for every region of at
//   least weak_tex_num pixels: its 4-neighbour border -> cv::HoughLinesP -> cv::line onto the edge image -> cv::resize to
//   the level size -> threshold -> border clean-up -> Connect / Label_Update -> regions of <= weak_tex_num pixels = -1.
// The OpenCV calls (third party, not under /root/reference) are restated from the published algorithms: 8-bit bilinear
// cv::resize (fixed-point, 11-bit coefficients; exact 2x2 averaging when both ratios are exactly 2), cv::line (8-connected
// Bresenham), cv::HoughLinesP (hough_cpu.cpp).  PINNED against OpenCV 4.13.0: tools/make_label_golden.py ->
// tests/golden/label_segment.npz, and piecewise against cv2 in tests/test_edges.py where cv2 is importable.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ccl_ref.hpp"

extern "C" int hough_cpu_lines_p(const uint8_t* image, int width, int height, float rho, float theta, int threshold, int line_length, int line_gap, int* lines, int max_lines);

namespace {

int cv_round_f(float v) { return (int)std::nearbyintf(v); }   // saturate_cast<short>(float): round half to even

// cv::resize(src, dst, Size(dw, dh), 0, 0, INTER_LINEAR) for CV_8UC1 (modules/imgproc/src/resize.cpp)
void resize8u(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) {
	const double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
	const int iscale_x = (int)std::lround(scale_x), iscale_y = (int)std::lround(scale_y);
	const bool area_fast = std::abs(scale_x - iscale_x) < 2.220446049250313e-16 && std::abs(scale_y - iscale_y) < 2.220446049250313e-16;
	if (area_fast && iscale_x == 2 && iscale_y == 2) {   // INTER_LINEAR is replaced by the 2x2 INTER_AREA fast path
		for (int y = 0; y < dh; ++y)
			for (int x = 0; x < dw; ++x) {
				const uint8_t* p = src + (size_t)(2 * y) * sw + 2 * x;
				dst[(size_t)y * dw + x] = (uint8_t)((p[0] + p[1] + p[sw] + p[sw + 1] + 2) >> 2);
			}
		return;
	}
	std::vector<int> xofs(dw), yofs(dh);
	std::vector<int> alpha(2 * (size_t)dw), beta(2 * (size_t)dh);
	for (int dx = 0; dx < dw; ++dx) {
		float fx = (float)((dx + 0.5) * scale_x - 0.5);
		int sx = (int)std::floor(fx);
		fx -= sx;
		if (sx < 0) { fx = 0; sx = 0; }
		if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
		xofs[dx] = sx;
		alpha[2 * dx] = cv_round_f((1.f - fx) * 2048); alpha[2 * dx + 1] = cv_round_f(fx * 2048);
	}
	for (int dy = 0; dy < dh; ++dy) {   // no clamping of the coefficients in y: rows are clipped when they are fetched
		float fy = (float)((dy + 0.5) * scale_y - 0.5);
		const int sy = (int)std::floor(fy);
		fy -= sy;
		yofs[dy] = sy;
		beta[2 * dy] = cv_round_f((1.f - fy) * 2048); beta[2 * dy + 1] = cv_round_f(fy * 2048);
	}
	std::vector<int> row0(dw), row1(dw);
	auto hrow = [&](int sy, std::vector<int>& out) {
		sy = sy < 0 ? 0 : (sy >= sh ? sh - 1 : sy);
		const uint8_t* S = src + (size_t)sy * sw;
		for (int dx = 0; dx < dw; ++dx) {
			const int sx = xofs[dx], sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
			out[dx] = S[sx] * alpha[2 * dx] + S[sx1] * alpha[2 * dx + 1];
		}
	};
	for (int dy = 0; dy < dh; ++dy) {
		hrow(yofs[dy], row0); hrow(yofs[dy] + 1, row1);
		const int b0 = beta[2 * dy], b1 = beta[2 * dy + 1];
		for (int dx = 0; dx < dw; ++dx)
			dst[(size_t)dy * dw + dx] = (uint8_t)((((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2);
	}
}

// cv::line(img, p0, p1, 255, 1): thickness 1, LINE_8, both ends inside the image (modules/imgproc/src/drawing.cpp,
// LineIterator with leftToRight = true)
void line8(uint8_t* img, int w, int h, int x0, int y0, int x1, int y1, uint8_t value) {
	if (x0 < 0 || x0 >= w || x1 < 0 || x1 >= w || y0 < 0 || y0 >= h || y1 < 0 || y1 >= h) return;   // callers pass mask pixels
	int dx = x1 - x0, dy = y1 - y0;
	if (dx < 0) { dx = -dx; dy = -dy; x0 = x1; y0 = y1; }           // start from the left end
	int major_x = 1, major_y = 0, minor_x = 0, minor_y = dy < 0 ? -1 : 1;
	if (dy < 0) dy = -dy;
	if (dy > dx) { const int t = dx; dx = dy; dy = t; major_x = 0; major_y = minor_y; minor_x = 1; minor_y = 0; }
	int err = dx - (dy + dy);
	const int plus_delta = dx + dx, minus_delta = -(dy + dy);
	int x = x0, y = y0;
	for (int i = 0; i <= dx; ++i) {
		img[(size_t)y * w + x] = value;
		const bool both = err < 0;
		err += minus_delta + (both ? plus_delta : 0);
		x += major_x + (both ? minor_x : 0);
		y += major_y + (both ? minor_y : 0);
	}
}

// Roberts, APD.cpp:120-136 ((uchar)sqrt(int): truncation, modulo 256 beyond 255 as x86 converts)
void roberts(const uint8_t* src, int w, int h, uint8_t* dst) {
	for (int i = 0; i < h; i++)
		for (int j = 0; j < w; j++) {
			int t1, t2;
			if (i > 0 && i < h - 1 && j > 0 && j < w - 1) {
				t1 = src[(size_t)i * w + j] - src[(size_t)(i + 1) * w + j + 1];
				t2 = src[(size_t)(i + 1) * w + j] - src[(size_t)i * w + j + 1];
			} else t1 = t2 = 50;
			dst[(size_t)i * w + j] = (uint8_t)(int)std::sqrt((double)(t1 * t1 + t2 * t2));
		}
}

void threshold4(std::vector<uint8_t>& img) { for (auto& v : img) v = v > 4 ? 255 : 0; }   // cv::threshold(.., robthr = 4, 255, BINARY)

}  // namespace

extern "C" {

void label_cpu_resize8u(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) { resize8u(src, sw, sh, dst, dw, dh); }
void label_cpu_line(uint8_t* img, int w, int h, int x0, int y0, int x1, int y1, int value) { line8(img, w, h, x0, y0, x1, y1, (uint8_t)value); }
void label_cpu_roberts(const uint8_t* src, int w, int h, uint8_t* dst) { roberts(src, w, h, dst); }
// Connect + Label_Update on a 0 / 255 image: labels [h][w] int32; returns the number of label slots (label_cnt.size())
int label_cpu_connect(const uint8_t* img, int w, int h, int32_t* labels, int32_t* counts, int max_counts) {
	std::vector<int> lab, cnt;
	ccl_ref::connect_ref(img, h, w, lab, cnt);
	ccl_ref::label_update_ref(lab, h, w, cnt);
	for (size_t i = 0; i < lab.size(); ++i) labels[i] = lab[i];
	for (size_t i = 0; i < cnt.size() && (int)i < max_counts; ++i) counts[i] = cnt[i];
	return (int)cnt.size();
}

// Level size of the label map: round(cols / 2^scale) in float arithmetic (APD.cpp:441-443)
void label_cpu_size(int cols, int rows, int scale, int* new_cols, int* new_rows) {
	const float factor = 1.0f / (float)(1 << scale);
	*new_cols = (int)std::round(cols * factor);
	*new_rows = (int)std::round(rows * factor);
}

// EdgeSegment(scale, src, 1, false).  labels: [new_rows][new_cols] int32 (0 boundary, -1 small region, > 0 region id).
// edge_small (may be NULL): the quarter-size edge image after the Hough lines were drawn ([rows/4... ] see code), for
// stage-wise pinning.  Returns 0, or -1 on bad arguments.
int label_cpu_segment(const uint8_t* src, int cols, int rows, int scale, int32_t* labels, uint8_t* edge_small) {
	if (!src || !labels || cols < 16 || rows < 16 || scale < 0 || scale > 8) return -1;
	const int weak_tex_num = (int)(1.0 * rows * cols / (1024 << scale << scale));
	const int w1 = cols / 2, h1 = rows / 2, w2 = w1 / 2, h2 = h1 / 2;
	std::vector<uint8_t> down1((size_t)w1 * h1), down2((size_t)w2 * h2), dst((size_t)w2 * h2);
	resize8u(src, cols, rows, down1.data(), w1, h1);
	resize8u(down1.data(), w1, h1, down2.data(), w2, h2);
	const int m = w2 < h2 ? w2 : h2;
	const int houthr = (int)(m / 30.0), min_line_length = (int)(m / 30.0), max_line_gap = (int)(m / 30.0);
	roberts(down2.data(), w2, h2, dst.data());
	threshold4(dst);
	std::vector<int> lab0, cnt0;
	ccl_ref::connect_ref(dst.data(), h2, w2, lab0, cnt0);
	ccl_ref::label_update_ref(lab0, h2, w2, cnt0);
	std::vector<uint8_t> img_weak((size_t)w2 * h2);
	std::vector<int> lines(4 * 65536);
	for (size_t k = 1; k < cnt0.size(); k++) {
		if (cnt0[k] < weak_tex_num) continue;
		const int weak_index = (int)k;
		std::fill(img_weak.begin(), img_weak.end(), 0);
		for (int y = 0; y < h2; y++)
			for (int x = 0; x < w2; x++) {
				if (lab0[(size_t)y * w2 + x] == weak_index) continue;
				bool border = false;
				if (x > 0 && lab0[(size_t)y * w2 + x - 1] == weak_index) border = true;
				if (x < w2 - 1 && lab0[(size_t)y * w2 + x + 1] == weak_index) border = true;
				if (y > 0 && lab0[(size_t)(y - 1) * w2 + x] == weak_index) border = true;
				if (y < h2 - 1 && lab0[(size_t)(y + 1) * w2 + x] == weak_index) border = true;
				if (border) img_weak[(size_t)y * w2 + x] = 255;
			}
		const int n = hough_cpu_lines_p(img_weak.data(), w2, h2, 1.0f, (float)(3.14159265358979323846 / 180), houthr, min_line_length, max_line_gap, lines.data(), 65536);
		for (int i = 0; i < n && i < 65536; i++) line8(dst.data(), w2, h2, lines[4 * i], lines[4 * i + 1], lines[4 * i + 2], lines[4 * i + 3], 255);
	}
	if (edge_small) memcpy(edge_small, dst.data(), dst.size());
	int new_cols, new_rows;
	label_cpu_size(cols, rows, scale, &new_cols, &new_rows);
	std::vector<uint8_t> up((size_t)new_cols * new_rows);
	resize8u(dst.data(), w2, h2, up.data(), new_cols, new_rows);
	threshold4(up);
	for (int y = 0; y < new_rows; y++) {                                                   // APD.cpp:452-457
		if (up[(size_t)y * new_cols + 1] == 0) up[(size_t)y * new_cols] = 0;
		if (up[(size_t)y * new_cols + new_cols - 2] == 0) up[(size_t)y * new_cols + new_cols - 1] = 0;
	}
	for (int x = 0; x < new_cols; x++) {                                                   // APD.cpp:458-463
		if (up[(size_t)new_cols + x] == 0) up[x] = 0;
		if (up[(size_t)(new_rows - 2) * new_cols + x] == 0) up[(size_t)(new_rows - 1) * new_cols + x] = 0;
	}
	std::vector<int> lab, cnt;
	ccl_ref::connect_ref(up.data(), new_rows, new_cols, lab, cnt);
	ccl_ref::label_update_ref(lab, new_rows, new_cols, cnt);
	for (size_t i = 0; i < lab.size(); ++i) {                                               // APD.cpp:488-491
		const int label = lab[i];
		labels[i] = (cnt[label] <= weak_tex_num && label != 0) ? -1 : label;
	}
	return 0;
}

}  // extern "C"
