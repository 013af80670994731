// oracle/cpu/visibility_cpu.cpp — TEST INFRASTRUCTURE ONLY.
// CPU restatement of the visibility-map restoration that follows RunPatchMatch in the reference
// (SURVEY §8f row N1): ProcessProblem, main.cpp:288-363, with Connect (APD.cpp:244-346), Label_Seek
// (APD.cpp:138-193) and Label_Update (APD.cpp:195-241).  For every source view the pixels whose selected-view
// bit is clear are grouped into 4-connected regions; regions smaller than 20*(8/scale)^2 pixels get the bit set.
//
// The restatement keeps the reference's sequential algorithm, including its two quirks:
//   * Connect records a merge by overwriting connection[larger] = smaller without looking up roots, so an
//     earlier link of `larger` can be lost;
//   * Label_Update repairs lost links from pixel adjacencies, but only looks at pairs whose left/upper pixel is
//     not in the last row / last column (loops run to rows-1 / cols-1).
// cpu_restore_visibility_cc is the same operation on exact 4-connected components; the tests measure where the
// two differ (only a link lost by the first quirk that the second quirk cannot see can make them differ).
#include <cstdint>
#include <cstring>
#include <vector>

#include "ccl_ref.hpp"

namespace {

using namespace ccl_ref;

// exact 4-connected components of the invisible pixels (union-find), sizes per pixel
void component_sizes_cc(const uint8_t* mask, int rows, int cols, std::vector<int>& size_of_pixel) {
	const size_t n = (size_t)rows * cols;
	std::vector<int> parent(n);
	for (size_t i = 0; i < n; ++i) parent[i] = (int)i;
	auto find = [&](int a) { while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; } return a; };
	auto unite = [&](int a, int b) { a = find(a); b = find(b); if (a != b) parent[a > b ? a : b] = a > b ? b : a; };
	for (int y = 0; y < rows; ++y)
		for (int x = 0; x < cols; ++x) {
			const size_t c = (size_t)y * cols + x;
			if (mask[c] != 0) continue;
			if (x > 0 && mask[c - 1] == 0) unite((int)c, (int)c - 1);
			if (y > 0 && mask[c - cols] == 0) unite((int)c, (int)(c - cols));
		}
	std::vector<int> cnt(n, 0);
	for (size_t c = 0; c < n; ++c) if (mask[c] == 0) cnt[find((int)c)]++;
	size_of_pixel.assign(n, 0);
	for (size_t c = 0; c < n; ++c) if (mask[c] == 0) size_of_pixel[c] = cnt[find((int)c)];
}

}  // namespace

extern "C" {

// main.cpp:288-363.  selected_in/out: [H*W] uint32 bit masks; scale_size in {1,2,4,8} (problem.scale_size).
// exact_cc = 0: the reference's Connect + Label_Update; 1: exact 4-connected components.
int cpu_restore_visibility(const uint32_t* selected_in, uint32_t* selected_out, int W, int H, int S, int scale_size, int exact_cc) {
	if (!selected_in || !selected_out || W <= 0 || H <= 0 || S < 0 || S > 32 || scale_size <= 0) return 1;
	const size_t n = (size_t)W * H;
	const int thr = 20 * (8 / scale_size) * (8 / scale_size);
	std::vector<uint8_t> mask(n);
	std::memset(selected_out, 0, n * sizeof(uint32_t));
	for (int i = 0; i < S; ++i) {
		for (size_t c = 0; c < n; ++c) mask[c] = ((selected_in[c] >> i) & 1) ? 255 : 0;
		if (exact_cc) {
			std::vector<int> sz;
			component_sizes_cc(mask.data(), H, W, sz);
			for (size_t c = 0; c < n; ++c)
				if (mask[c] == 255 || sz[c] < thr) selected_out[c] |= 1u << i;
		} else {
			std::vector<int> lab, cnt;
			connect_ref(mask.data(), H, W, lab, cnt);
			label_update_ref(lab, H, W, cnt);
			// label 0 and labels with fewer than thr pixels are painted "visible" (main.cpp:333-352)
			for (size_t c = 0; c < n; ++c) {
				const int l = lab[c];
				if (l == 0 || cnt[l] < thr) selected_out[c] |= 1u << i;
			}
		}
	}
	return 0;
}

}  // extern "C"
