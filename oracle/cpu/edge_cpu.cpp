// oracle/cpu/edge_cpu.cpp — TEST INFRASTRUCTURE ONLY.
// CPU restatement of the depth-edge prior the hot path consumes as `edge_cuda` (edge half):
// EdgeSegment(scale, image, mode 0, use_canny = true), reference APD.cpp:348-466, called from GetProblemEdges
// (main.cpp:218) on the 8-bit image of the pyramid level:
//   1. grey-level histogram (in float, as the reference keeps it) -> "median" = first level < 255 whose cumulative count
//      exceeds rows*cols/2, else -1 (APD.cpp:405-428);
//   2. threshold1 = (1 - 0.67f) * median, threshold2 = median, both truncated to int (APD.cpp:430-432);
//   3. cv::Canny(src, dst, threshold1, threshold2, 3, L2gradient = true) (APD.cpp:433);
//   4. cv::resize to the image's own size (a copy), cv::threshold(> 4 -> 255) (APD.cpp:437-446): no-ops on a 0/255 map;
//   5. border clean-up: a border pixel whose inner neighbour is 0 becomes 0, columns first, then rows (APD.cpp:452-463).
// Step 3 lives in a third-party dependency that is not under /root/reference: OpenCV (the reference asks for >= 3.3,
// CMakeLists.txt:8-10; not vendored, no pinned version).  Its published algorithm (modules/imgproc/src/canny.cpp, 4.x)
// is restated here: 3x3 Sobel with replicated borders into 16-bit gradients, squared L2 magnitude in int, thresholds
// squared, non-maximum suppression with the fixed-point tan(22.5 deg) test, hysteresis over 8-neighbours.
// PINNED against OpenCV 4.13.0 (the Python cv2 of this container): tools/make_edge_golden.py ->
// tests/golden/edge_canny.npz, and exhaustively on random images in the CPU suite where cv2 is importable.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

// APD.cpp:405-432
void median_thresholds(const uint8_t* img, int rows, int cols, int* t1, int* t2) {
	float histogram[256] = {0};
	for (size_t i = 0; i < (size_t)rows * cols; ++i) histogram[img[i]]++;
	const int half = rows * cols / 2;
	int median_val = -1, temp_sum = 0;
	for (int i = 0; i < 255; i++) {
		temp_sum = temp_sum + histogram[i];   // int + float -> float -> int, as written
		if (temp_sum > half) { median_val = i; break; }
	}
	const float sigma = 0.67;
	*t1 = (1 - sigma) * median_val;
	*t2 = median_val;
}

// cv::Canny(8-bit single channel, aperture 3, L2gradient = true)
void canny_l2(const uint8_t* src, int rows, int cols, double low_thresh, double high_thresh, uint8_t* dst) {
	if (low_thresh > high_thresh) std::swap(low_thresh, high_thresh);
	low_thresh = std::min(32767.0, low_thresh);
	high_thresh = std::min(32767.0, high_thresh);
	if (low_thresh > 0) low_thresh *= low_thresh;
	if (high_thresh > 0) high_thresh *= high_thresh;
	const int low = (int)std::floor(low_thresh), high = (int)std::floor(high_thresh);
	const size_t n = (size_t)rows * cols;
	std::vector<int16_t> dx(n), dy(n);
	auto at = [&](int y, int x) -> int {
		y = y < 0 ? 0 : (y >= rows ? rows - 1 : y);
		x = x < 0 ? 0 : (x >= cols ? cols - 1 : x);
		return src[(size_t)y * cols + x];
	};
	for (int y = 0; y < rows; ++y)
		for (int x = 0; x < cols; ++x) {
			dx[(size_t)y * cols + x] = (int16_t)((at(y - 1, x + 1) + 2 * at(y, x + 1) + at(y + 1, x + 1)) - (at(y - 1, x - 1) + 2 * at(y, x - 1) + at(y + 1, x - 1)));
			dy[(size_t)y * cols + x] = (int16_t)((at(y + 1, x - 1) + 2 * at(y + 1, x) + at(y + 1, x + 1)) - (at(y - 1, x - 1) + 2 * at(y - 1, x) + at(y - 1, x + 1)));
		}
	// magnitude with a zero frame, map with a frame of 1 ("cannot be an edge")
	const int mw = cols + 2;
	std::vector<int> mag((size_t)(rows + 2) * mw, 0);
	for (int y = 0; y < rows; ++y)
		for (int x = 0; x < cols; ++x) {
			const int gx = dx[(size_t)y * cols + x], gy = dy[(size_t)y * cols + x];
			mag[(size_t)(y + 1) * mw + x + 1] = gx * gx + gy * gy;
		}
	std::vector<uint8_t> map((size_t)(rows + 2) * mw, 1);
	std::vector<size_t> stack;
	const int TG22 = (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5);
	for (int y = 0; y < rows; ++y)
		for (int x = 0; x < cols; ++x) {
			const size_t c = (size_t)(y + 1) * mw + x + 1;
			const int m = mag[c];
			bool candidate = false;
			if (m > low) {
				const int xs = dx[(size_t)y * cols + x], ys = dy[(size_t)y * cols + x];
				const int ax = std::abs(xs), ay = std::abs(ys) << 15;
				const int tg22x = ax * TG22;
				if (ay < tg22x) candidate = m > mag[c - 1] && m >= mag[c + 1];
				else {
					const int tg67x = tg22x + (ax << 16);
					if (ay > tg67x) candidate = m > mag[c - mw] && m >= mag[c + mw];
					else {
						const int s = (xs ^ ys) < 0 ? -1 : 1;
						candidate = m > mag[c - mw - s] && m > mag[c + mw + s];
					}
				}
			}
			if (!candidate) map[c] = 1;
			else if (m > high) { map[c] = 2; stack.push_back(c); }
			else map[c] = 0;
		}
	while (!stack.empty()) {
		const size_t c = stack.back();
		stack.pop_back();
		const ptrdiff_t nb[8] = {-mw - 1, -mw, -mw + 1, -1, 1, mw - 1, mw, mw + 1};
		for (ptrdiff_t d : nb)
			if (map[c + d] == 0) { map[c + d] = 2; stack.push_back(c + d); }
	}
	for (int y = 0; y < rows; ++y)
		for (int x = 0; x < cols; ++x) dst[(size_t)y * cols + x] = map[(size_t)(y + 1) * mw + x + 1] == 2 ? 255 : 0;
}

}  // namespace

extern "C" {

void edge_cpu_thresholds(const uint8_t* img, int cols, int rows, int* t1, int* t2) { median_thresholds(img, rows, cols, t1, t2); }

void edge_cpu_canny(const uint8_t* img, int cols, int rows, double low, double high, uint8_t* dst) { canny_l2(img, rows, cols, low, high, dst); }

// EdgeSegment(scale, img, 0, true); `canny_out` (may be NULL) receives the map before the border clean-up
int edge_cpu_segment(const uint8_t* img, int cols, int rows, uint8_t* edge, uint8_t* canny_out) {
	if (!img || !edge || cols < 3 || rows < 3) return -1;
	int t1, t2;
	median_thresholds(img, rows, cols, &t1, &t2);
	canny_l2(img, rows, cols, (double)t1, (double)t2, edge);
	if (canny_out) memcpy(canny_out, edge, (size_t)rows * cols);
	for (size_t i = 0; i < (size_t)rows * cols; ++i) edge[i] = edge[i] > 4 ? 255 : 0;   // APD.cpp:446
	for (int y = 0; y < rows; y++) {                                                       // APD.cpp:452-457
		if (edge[(size_t)y * cols + 1] == 0) edge[(size_t)y * cols] = 0;
		if (edge[(size_t)y * cols + cols - 2] == 0) edge[(size_t)y * cols + cols - 1] = 0;
	}
	for (int x = 0; x < cols; x++) {                                                       // APD.cpp:458-463
		if (edge[(size_t)1 * cols + x] == 0) edge[x] = 0;
		if (edge[(size_t)(rows - 2) * cols + x] == 0) edge[(size_t)(rows - 1) * cols + x] = 0;
	}
	return 0;
}

}  // extern "C"
