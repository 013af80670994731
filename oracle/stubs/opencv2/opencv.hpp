// Minimal stand-in for the OpenCV C++ headers (OpenCV is not installed in this image).
// TEST INFRASTRUCTURE ONLY: lets the reference's unmodified APD.cu / APD.h / main.h compile
// (oracle/_ref). Provides only the names those three files mention: uchar, MIN/MAX (same
// definitions as opencv2/core/cvdef.h), cv::Mat{rows,cols,ptr<T>()}, cv::Mat_, cv::Vec3f,
// cv::Point, cv::Size2i. Nothing here is used by the product library.
#ifndef DVP_ORACLE_OPENCV_STUB_HPP
#define DVP_ORACLE_OPENCV_STUB_HPP
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <cfloat>
#include <cassert>
#include <cmath>
#include <vector>
#include <string>

typedef unsigned char uchar;
#ifndef MIN
#define MIN(a, b) ((a) > (b) ? (b) : (a))
#endif
#ifndef MAX
#define MAX(a, b) ((a) < (b) ? (b) : (a))
#endif
#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_8UC1 0
#define CV_32SC1 4
#define CV_32FC1 5
#define CV_32FC3 21
#define CV_8UC3 16

#define CV_PI 3.1415926535897932384626433832795
namespace cv {
struct Scalar {   // EdgeSegment (APD.cpp:379, 399) initialises an image with Scalar(0) and draws with Scalar(255, 0, 0)
	double val[4];
	Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : val{a, b, c, d} {}
};
struct Vec4i {
	int val[4];
	int& operator[](int i) { return val[i]; }
	const int& operator[](int i) const { return val[i]; }
};
enum { THRESH_BINARY = 0 };
struct Point {
	int x, y;
	Point() : x(0), y(0) {}
	Point(int _x, int _y) : x(_x), y(_y) {}
};
struct Size2i {
	int width, height;
	Size2i() : width(0), height(0) {}
	Size2i(int w, int h) : width(w), height(h) {}
};
typedef Size2i Size;
struct Vec3f {
	float val[3];
	Vec3f() : val{0.f, 0.f, 0.f} {}
	Vec3f(float a, float b, float c) : val{a, b, c} {}
	float& operator[](int i) { return val[i]; }
	const float& operator[](int i) const { return val[i]; }
};
struct Vec3b {   // host code compiled from reference line ranges (oracle/ref_host.cu) reads colours through it
	unsigned char val[3];
	Vec3b() : val{0, 0, 0} {}
	Vec3b(int a, int b, int c) : val{(unsigned char)a, (unsigned char)b, (unsigned char)c} {}
	bool operator==(const Vec3b& o) const { return val[0] == o.val[0] && val[1] == o.val[1] && val[2] == o.val[2]; }
	bool operator!=(const Vec3b& o) const { return !(*this == o); }
	unsigned char& operator[](int i) { return val[i]; }
	const unsigned char& operator[](int i) const { return val[i]; }
};
// A non-owning-or-owning dense matrix: just enough for `rows`, `cols`, `ptr<T>(r)`.
class Mat {
public:
	int rows, cols;
	size_t elem_size;
	int type_;
	std::vector<unsigned char> storage;
	unsigned char* data;   // == storage.data(); `data` and `step` are what ReadBinMat / WriteBinMat touch (APD.cpp:548-650)
	size_t step;           // bytes per row
	Mat() : rows(0), cols(0), elem_size(1), type_(0), data(nullptr), step(0) {}
	Mat(int r, int c, int type) { create(r, c, type); }
	Mat(int r, int c, int type, const Scalar& s) { create(r, c, type); if (elem_size == 1) std::memset(storage.data(), (int)s.val[0], storage.size()); }
	Mat(const Mat& o) : rows(o.rows), cols(o.cols), elem_size(o.elem_size), type_(o.type_), storage(o.storage) { rebind(); }
	Mat& operator=(const Mat& o) {
		if (this != &o) { rows = o.rows; cols = o.cols; elem_size = o.elem_size; type_ = o.type_; storage = o.storage; }
		rebind();
		return *this;
	}
	void create(int r, int c, int type) {
		rows = r; cols = c; type_ = type;
		elem_size = (type == CV_8U) ? 1 : (type == CV_8UC3 ? 3 : (type == CV_32FC3 ? 12 : 4));
		storage.assign((size_t)r * c * elem_size, 0);
		rebind();
	}
	bool empty() const { return rows == 0 || cols == 0; }
	Size size() const { return Size(cols, rows); }
	// convertTo between CV_8UC1 and CV_32FC1 (GetProblemEdges, main.cpp:202, 209): to float exactly, to 8 bits with
	// cv::saturate_cast<uchar>(float) = clamp(cvRound(v)), cvRound rounding halves to even
	void convertTo(Mat& dst, int type) const {
		Mat out(rows, cols, type);
		const size_t n = (size_t)rows * cols;
		if (type_ == CV_8UC1 && type == CV_32FC1) { for (size_t i = 0; i < n; ++i) reinterpret_cast<float*>(out.storage.data())[i] = (float)storage[i]; }
		else if (type_ == CV_32FC1 && type == CV_8UC1) { for (size_t i = 0; i < n; ++i) { const long v = std::lrint(reinterpret_cast<const float*>(storage.data())[i]); out.storage[i] = (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v)); } }
		else out = *this;
		dst = out;
	}
	static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }          // create() zero-fills
	static Mat zeros(Size s, int type) { return Mat(s.height, s.width, type); }
	int type() const { return type_; }   // RescaleMatToTargetSize (APD.cpp:1781) builds its destination from it
	template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }
	template <typename T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
	Mat clone() const { return *this; }   // storage is a std::vector: copying the object is a deep copy
	template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(storage.data() + (size_t)r * cols * elem_size); }
	template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(storage.data() + (size_t)r * cols * elem_size); }
private:
	void rebind() { data = storage.empty() ? nullptr : storage.data(); step = (size_t)cols * elem_size; }
};
template <typename T> class Mat_ : public Mat {
public:
	Mat_() {}
	Mat_(const Mat& m) : Mat(m) {}
	Mat_(int r, int c) : Mat(r, c, sizeof(T) == 1 ? CV_8UC1 : (sizeof(T) == 3 ? CV_8UC3 : (sizeof(T) == 12 ? CV_32FC3 : CV_32FC1))) {}   // main.cpp:293
	Mat_(int r, int c, int type) : Mat(r, c, type) {}                                                                              // main.cpp:292
};
}  // namespace cv
#endif
