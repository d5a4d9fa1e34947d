// oracle/standin/l3d_standin_offpath.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Definitions for the OpenCV / LineDescriptor / Converter functions the reference's sources CALL only off the
// matching -> scoring -> affinity -> 3-D line path (undistortion, resizing, LSD / EDLines detection, drawing, the
// 4-view SVD triangulation whose result the reference discards).  The harness never reaches them; each one
// aborts loudly if it ever is, instead of returning something made up.  (Python's ctypes loads libraries with
// RTLD_NOW, so the symbols must exist.)
#pragma once
#include <cstdio>
#include <cstdlib>
#define L3D_OFFPATH(name)                                                                                 \
    do {                                                                                                  \
        std::fprintf(stderr, "oracle/_ref: %s is off the tested path and has no stand-in\n", name);       \
        std::abort();                                                                                     \
    } while (0)

int LineDescriptor::GetLineDescriptor(cv::Mat&, ScaleLines&) { L3D_OFFPATH("LineDescriptor::GetLineDescriptor"); }
namespace ORB_SLAM2 {
cv::Mat Converter::toCvMat(const Eigen::Matrix<double, 3, 4>&) { L3D_OFFPATH("Converter::toCvMat"); }
cv::Mat Converter::toCvMat(const Eigen::Matrix3d&) { L3D_OFFPATH("Converter::toCvMat"); }
Eigen::Vector3d Converter::toVector3d(const cv::Mat&) { L3D_OFFPATH("Converter::toVector3d"); }
}  // namespace ORB_SLAM2
namespace cv {
MatExpr::operator Mat() const { L3D_OFFPATH("cv::MatExpr"); }
int Mat::type() const { L3D_OFFPATH("cv::Mat::type"); }
int Mat::channels() const { L3D_OFFPATH("cv::Mat::channels"); }
Mat Mat::clone() const { L3D_OFFPATH("cv::Mat::clone"); }
Mat Mat::row(int) const { L3D_OFFPATH("cv::Mat::row"); }
Mat Mat::rowRange(int, int) const { L3D_OFFPATH("cv::Mat::rowRange"); }
Mat Mat::col(int) const { L3D_OFFPATH("cv::Mat::col"); }
Mat Mat::t() const { L3D_OFFPATH("cv::Mat::t"); }
void Mat::copyTo(Mat&) const { L3D_OFFPATH("cv::Mat::copyTo"); }
Mat& Mat::operator=(const Scalar&) { L3D_OFFPATH("cv::Mat::operator="); }
template <typename T> T& Mat::at(int, int) { L3D_OFFPATH("cv::Mat::at"); }
template <typename T> const T& Mat::at(int, int) const { L3D_OFFPATH("cv::Mat::at"); }
template double& Mat::at<double>(int, int);
template float& Mat::at<float>(int, int);
MatExpr Mat::zeros(int, int, int) { L3D_OFFPATH("cv::Mat::zeros"); }
MatExpr Mat::eye(int, int, int) { L3D_OFFPATH("cv::Mat::eye"); }
template <typename T> MatExpr Mat_<T>::zeros(int, int) { L3D_OFFPATH("cv::Mat_::zeros"); }
template <typename T> MatExpr Mat_<T>::eye(int, int) { L3D_OFFPATH("cv::Mat_::eye"); }
template <typename T> T& Mat_<T>::operator()(int, int) { L3D_OFFPATH("cv::Mat_::operator()"); }
template class Mat_<double>;
Mat operator*(const Mat&, const Mat&) { L3D_OFFPATH("cv::operator*"); }
Mat operator*(double, const Mat&) { L3D_OFFPATH("cv::operator*"); }
Mat operator-(const Mat&, const Mat&) { L3D_OFFPATH("cv::operator-"); }
Mat operator/(const Mat&, double) { L3D_OFFPATH("cv::operator/"); }
template <typename T> T* Ptr<T>::operator->() const { L3D_OFFPATH("cv::Ptr"); }
template class Ptr<LineSegmentDetector>;
void LineSegmentDetector::detect(const Mat&, std::vector<Vec4f>&) { L3D_OFFPATH("cv::LineSegmentDetector"); }
Ptr<LineSegmentDetector> createLineSegmentDetectorPtr(int) { L3D_OFFPATH("cv::createLineSegmentDetectorPtr"); }
Ptr<LineSegmentDetector> createLineSegmentDetector(int) { L3D_OFFPATH("cv::createLineSegmentDetector"); }
void initUndistortRectifyMap(const Mat&, const Mat&, const Mat&, const Mat&, Size, int, Mat&, Mat&) { L3D_OFFPATH("cv::initUndistortRectifyMap"); }
void remap(const Mat&, Mat&, const Mat&, const Mat&, int, int) { L3D_OFFPATH("cv::remap"); }
void cvtColor(const Mat&, Mat&, int) { L3D_OFFPATH("cv::cvtColor"); }
void resize(const Mat&, Mat&, Size, double, double) { L3D_OFFPATH("cv::resize"); }
void line(Mat&, Point, Point, const Scalar&, int) { L3D_OFFPATH("cv::line"); }
void SVD::compute(const Mat&, Mat&, Mat&, Mat&, int) { L3D_OFFPATH("cv::SVD::compute"); }
}  // namespace cv
