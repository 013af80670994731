// oracle/cpu/hough_cpu.cpp — TEST INFRASTRUCTURE ONLY.
// Groundwork for the label half of row N4 (NOT yet used by any device path): CPU restatement of the two OpenCV routines
// EdgeSegment's Roberts + Hough branch calls on the region borders (reference APD.cpp:396-400):
//   cv::HoughLinesP(img, lines, 1, CV_PI / 180, threshold, minLineLength, maxLineGap)   and   cv::line(img, p0, p1, 255, 1)
// Both live in OpenCV (third party, not under /root/reference); restated from the published algorithm (modules/imgproc/
// src/hough.cpp HoughLinesProbabilistic: points visited in the order of cv::RNG seeded with (uint64)-1, votes over
// numangle x numrho, fixed-point line walk with gap / length tests; modules/imgproc/src/drawing.cpp: 8-connected
// Bresenham with clipping).  Pinned against OpenCV 4.13.0 in tests/test_edges.py where cv2 is importable.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

struct Rng {   // cv::RNG (multiply-with-carry), modules/core/include/opencv2/core/operations.hpp
	uint64_t state;
	explicit Rng(uint64_t s) : state(s ? s : 0xffffffffu) {}
	unsigned next() { state = (uint64_t)(unsigned)state * 4164903690U + (unsigned)(state >> 32); return (unsigned)state; }
	int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

int cv_round(double v) { return (int)std::nearbyint(v); }   // cvRound: round half to even (SSE2 cvtsd2si)

}  // namespace

extern "C" {

// lines: [max_lines][4] ints (x0, y0, x1, y1); returns the number of lines found (may exceed max_lines; excess not stored)
int hough_cpu_lines_p(const uint8_t* image, int width, int height, float rho, float theta, int threshold, int line_length, int line_gap, int* lines, int max_lines) {
	const float irho = 1 / rho;
	Rng rng((uint64_t)-1);
	int numangle = (int)std::floor((3.14159265358979323846 - 0.0) / theta) + 1;
	if (numangle > 1 && std::fabs(3.14159265358979323846 - (numangle - 1) * theta) < theta / 2) --numangle;
	const int numrho = cv_round(((width + height) * 2 + 1) / rho);
	std::vector<int> accum((size_t)numangle * numrho, 0);
	std::vector<uint8_t> mask((size_t)width * height);
	std::vector<float> trigtab((size_t)numangle * 2);
	for (int n = 0; n < numangle; n++) {
		trigtab[n * 2] = (float)(std::cos((double)n * theta) * irho);
		trigtab[n * 2 + 1] = (float)(std::sin((double)n * theta) * irho);
	}
	const float* ttab = trigtab.data();
	struct Pt { int x, y; };
	std::vector<Pt> nzloc;
	for (int y = 0; y < height; y++)
		for (int x = 0; x < width; x++) {
			if (image[(size_t)y * width + x]) { mask[(size_t)y * width + x] = 1; nzloc.push_back({x, y}); }
			else mask[(size_t)y * width + x] = 0;
		}
	int found = 0;
	for (int count = (int)nzloc.size(); count > 0; count--) {
		const int idx = rng.uniform(0, count);
		int max_val = threshold - 1, max_n = 0;
		const Pt point = nzloc[idx];
		Pt line_end[2] = {{0, 0}, {0, 0}};
		const int i = point.y, j = point.x;
		const int shift = 16;
		nzloc[idx] = nzloc[count - 1];
		if (!mask[(size_t)i * width + j]) continue;
		int* adata = accum.data();
		for (int n = 0; n < numangle; n++, adata += numrho) {
			int r = cv_round(j * ttab[n * 2] + i * ttab[n * 2 + 1]);
			r += (numrho - 1) / 2;
			const int val = ++adata[r];
			if (max_val < val) { max_val = val; max_n = n; }
		}
		if (max_val < threshold) continue;
		const float a = -ttab[max_n * 2 + 1], b = ttab[max_n * 2];
		int x0 = j, y0 = i, dx0, dy0, xflag;
		if (std::fabs(a) > std::fabs(b)) {
			xflag = 1;
			dx0 = a > 0 ? 1 : -1;
			dy0 = cv_round(b * (1 << shift) / std::fabs(a));
			y0 = (y0 << shift) + (1 << (shift - 1));
		} else {
			xflag = 0;
			dy0 = b > 0 ? 1 : -1;
			dx0 = cv_round(a * (1 << shift) / std::fabs(b));
			x0 = (x0 << shift) + (1 << (shift - 1));
		}
		for (int k = 0; k < 2; k++) {
			int gap = 0, x = x0, y = y0, dx = dx0, dy = dy0;
			if (k > 0) { dx = -dx; dy = -dy; }
			for (;; x += dx, y += dy) {
				int i1, j1;
				if (xflag) { j1 = x; i1 = y >> shift; } else { j1 = x >> shift; i1 = y; }
				if (j1 < 0 || j1 >= width || i1 < 0 || i1 >= height) break;
				if (mask[(size_t)i1 * width + j1]) { gap = 0; line_end[k].y = i1; line_end[k].x = j1; }
				else if (++gap > line_gap) break;
			}
		}
		const bool good_line = std::abs(line_end[1].x - line_end[0].x) >= line_length || std::abs(line_end[1].y - line_end[0].y) >= line_length;
		for (int k = 0; k < 2; k++) {
			int x = x0, y = y0, dx = dx0, dy = dy0;
			if (k > 0) { dx = -dx; dy = -dy; }
			for (;; x += dx, y += dy) {
				int i1, j1;
				if (xflag) { j1 = x; i1 = y >> shift; } else { j1 = x >> shift; i1 = y; }
				uint8_t* m = &mask[(size_t)i1 * width + j1];
				if (*m) {
					if (good_line) {
						int* ad = accum.data();
						for (int n = 0; n < numangle; n++, ad += numrho) {
							int r = cv_round(j1 * ttab[n * 2] + i1 * ttab[n * 2 + 1]);
							r += (numrho - 1) / 2;
							ad[r]--;
						}
					}
					*m = 0;
				}
				if (i1 == line_end[k].y && j1 == line_end[k].x) break;
			}
		}
		if (good_line) {
			if (found < max_lines) { int* o = lines + 4 * found; o[0] = line_end[0].x; o[1] = line_end[0].y; o[2] = line_end[1].x; o[3] = line_end[1].y; }
			++found;
		}
	}
	return found;
}

}  // extern "C"
