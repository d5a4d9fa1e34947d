// oracle/standin/l3d_standin_opencv.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Stand-ins for the OpenCV names the reference's Line3D++ sources mention.  On the matching / scoring / affinity
// path OpenCV only supplies the image SIZE (cv::Mat::cols / rows), cv::Vec4f for the given 2-D segments and the
// tick counter; everything else (undistortion, resizing, LSD, drawing, the 4-view SVD whose result the reference
// discards) is off the path and only DECLARED here: the shared object is loaded with lazy binding and those
// functions are never called.
#pragma once
#include <chrono>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32F 5
#define CV_32FC1 5
#define CV_64F 6
#define CV_64FC1 6
#define CV_RGB2GRAY 7
#define CV_BGR2GRAY 6
#define CV_AA 16

namespace cv {
template <typename T>
using vector = std::vector<T>;   // OpenCV 2.4 exported std::vector as cv::vector

template <typename T, int N>
struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(0); }
    Vec(T a, T b, T c, T d) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    T& operator()(int i) { return val[i]; }
    const T& operator()(int i) const { return val[i]; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
typedef Vec<float, 4> Vec4f;
typedef Vec<int, 4> Vec4i;

struct Size {
    int width = 0, height = 0;
    Size() {}
    Size(int w, int h) : width(w), height(h) {}
};
struct Point {
    int x = 0, y = 0;
    Point() {}
    template <typename A, typename B>
    Point(A a, B b) : x((int)a), y((int)b) {}
};
struct Point2f {
    float x = 0, y = 0;
    Point2f() {}
    Point2f(float a, float b) : x(a), y(b) {}
};
struct Scalar {
    double v[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; }
};
struct KeyPoint {
    Point2f pt;
};

class Mat;
struct MatExpr {
    operator Mat() const;
};
class Mat {
  public:
    int rows = 0, cols = 0;
    Mat() {}
    Mat(int r, int c, int /*type*/) : rows(r), cols(c) {}
    Mat(int r, int c, int /*type*/, const Scalar&) : rows(r), cols(c) {}
    int type() const;
    int channels() const;
    bool empty() const { return rows == 0 || cols == 0; }
    Mat clone() const;
    Mat row(int) const;
    Mat rowRange(int, int) const;
    Mat col(int) const;
    Mat t() const;
    void copyTo(Mat&) const;
    template <typename T> T& at(int, int = 0);
    template <typename T> const T& at(int, int = 0) const;
    Mat& operator=(const Scalar&);
    static MatExpr zeros(int, int, int);
    static MatExpr eye(int, int, int);
};
Mat operator*(const Mat&, const Mat&);
Mat operator*(double, const Mat&);
Mat operator-(const Mat&, const Mat&);
Mat operator/(const Mat&, double);
template <typename T>
class Mat_ : public Mat {
  public:
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, 0) {}
    static MatExpr zeros(int, int);
    static MatExpr eye(int, int);
    T& operator()(int, int);
};
template <typename T>
class Ptr {
  public:
    T* operator->() const;
};
class LineSegmentDetector {
  public:
    void detect(const Mat&, std::vector<Vec4f>&);
};
enum { LSD_REFINE_NONE = 0, LSD_REFINE_STD = 1, LSD_REFINE_ADV = 2 };
enum { INTER_LINEAR = 1, BORDER_CONSTANT = 0 };
Ptr<LineSegmentDetector> createLineSegmentDetectorPtr(int);
Ptr<LineSegmentDetector> createLineSegmentDetector(int);
void initUndistortRectifyMap(const Mat&, const Mat&, const Mat&, const Mat&, Size, int, Mat&, Mat&);
void remap(const Mat&, Mat&, const Mat&, const Mat&, int, int = 0);
void cvtColor(const Mat&, Mat&, int);
void resize(const Mat&, Mat&, Size, double = 0, double = 0);
void line(Mat&, Point, Point, const Scalar&, int = 1);
struct SVD {
    enum { MODIFY_A = 1, FULL_UV = 4 };
    static void compute(const Mat&, Mat&, Mat&, Mat&, int = 0);
};
// a monotonic tick counter (only differences divided by the frequency are used, for log lines)
inline int64_t getTickCount()
{
    return (int64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline double getTickFrequency() { return 1e9; }
}  // namespace cv
