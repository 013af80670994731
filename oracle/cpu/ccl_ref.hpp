// oracle/cpu/ccl_ref.hpp — TEST INFRASTRUCTURE ONLY.
// The reference's own connected-component labelling, restated once and shared by the visibility restoration (row N1,
// visibility_cpu.cpp) and the label half of EdgeSegment (row N4, label_cpu.cpp): Connect (APD.cpp:233-346), Label_Seek
// (APD.cpp:138-190) and Label_Update (APD.cpp:192-231), quirks included (see visibility_cpu.cpp's header).
#pragma once
#include <cstdint>
#include <vector>

namespace ccl_ref {

// reference Connect, APD.cpp:244-346 (mask: 255 = visible, 0 = invisible)
inline void connect_ref(const uint8_t* mask, int rows, int cols, std::vector<int>& label_mask, std::vector<int>& label_cnt) {
	label_mask.assign((size_t)rows * cols, 0);
	int cnt = 1;
	std::vector<int> connection;
	connection.push_back(0);
	for (int y = 0; y < rows; y++)
		for (int x = 0; x < cols; x++) {
			const size_t c = (size_t)y * cols + x;
			if (mask[c] == 255) { label_mask[c] = 0; continue; }
			const bool left = x > 0 && mask[c] == 0 && mask[c - 1] == 0;
			const bool up = y > 0 && mask[c] == 0 && mask[c - cols] == 0;
			if (left) label_mask[c] = label_mask[c - 1];
			if (up) label_mask[c] = label_mask[c - cols];
			if (!left && !up) {
				label_mask[c] = cnt;
				connection.push_back(cnt);
				cnt++;
			} else if (left && up) {
				const int left_label = label_mask[c - 1], up_label = label_mask[c - cols];
				if (left_label > up_label) { connection[left_label] = up_label; label_mask[c] = up_label; }
				else if (left_label < up_label) { connection[up_label] = left_label; label_mask[c] = left_label; }
			}
		}
	for (size_t i = 1; i < connection.size(); i++) {
		int cur = connection[i], pre = connection[cur];
		while (pre != cur) { cur = pre; pre = connection[pre]; }
		connection[i] = cur;
	}
	int label_num = 1;
	std::vector<int> mapping(connection.size(), 0);
	for (size_t i = 1; i < connection.size(); i++)
		if (connection[i] == (int)i) mapping[i] = label_num++;
	for (size_t i = 1; i < connection.size(); i++) connection[i] = mapping[connection[i]];
	label_cnt.assign(label_num, 0);
	for (size_t c = 0; c < label_mask.size(); c++) {
		label_mask[c] = connection[label_mask[c]];
		label_cnt[label_mask[c]]++;
	}
}

// reference Label_Seek, APD.cpp:138-193 (groups of mutually connected labels; the first member names the group)
inline void label_seek(int l1, int l2, std::vector<std::vector<int>>& groups) {
	int ind1 = -1, ind2 = -1;
	for (size_t y = 0; y < groups.size() && (ind1 < 0 || ind2 < 0); y++)
		for (size_t x = 0; x < groups[y].size(); x++) {
			if (ind1 >= 0 && ind2 >= 0) break;
			if (groups[y][x] == l1) ind1 = (int)y;   // a later match overwrites an earlier one, as in the reference
			if (groups[y][x] == l2) ind2 = (int)y;
		}
	if (ind1 < 0 && ind2 < 0) groups.push_back({l1, l2});
	else if (ind1 < 0) groups[ind2].push_back(l1);
	else if (ind2 < 0) groups[ind1].push_back(l2);
	else if (ind1 != ind2) {
		for (int v : groups[ind2]) {
			bool rep = false;
			for (int u : groups[ind1]) if (u == v) { rep = true; break; }
			if (!rep) groups[ind1].push_back(v);
		}
		groups.erase(groups.begin() + ind2);
	}
}

// reference Label_Update, APD.cpp:195-241
inline void label_update_ref(std::vector<int>& label_mask, int rows, int cols, std::vector<int>& label_cnt) {
	std::vector<std::vector<int>> groups;
	for (int i = 0; i < rows - 1; ++i)
		for (int j = 0; j < cols - 1; ++j) {
			const int center = label_mask[(size_t)i * cols + j];
			const int right = label_mask[(size_t)i * cols + j + 1];
			const int down = label_mask[(size_t)(i + 1) * cols + j];
			if (center != 0 && right != 0 && center != right) label_seek(center, right, groups);
			if (center != 0 && down != 0 && center != down) label_seek(center, down, groups);
		}
	std::vector<int> map(label_cnt.size());
	for (size_t i = 0; i < label_cnt.size(); ++i) { map[i] = (int)i; label_cnt[i] = 0; }
	for (auto& g : groups) for (int v : g) map[v] = g[0];
	for (size_t c = 0; c < label_mask.size(); ++c) {
		label_mask[c] = map[label_mask[c]];
		label_cnt[label_mask[c]]++;
	}
}

}  // namespace ccl_ref
