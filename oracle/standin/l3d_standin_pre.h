// oracle/standin/l3d_standin_pre.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Force-included (-include) in front of the reference's Line3D++ sources when they are compiled, unmodified and
// from where they lie under /root/reference, into oracle/_ref/libref_line3d*.so (oracle/Makefile).
//   * float4 / float2 / int2: include/dataArray.h:49-65 defines them itself when L3DPP_CUDA is off.
//   * LineDescriptor.hh / Converter.h are included with quotes from include/line3D.h and would drag in EDLines,
//     OpenCV and g2o: their include guards are pre-defined on the command line (-DLINEDESCRIPTOR_HH_ -DCONVERTER_H)
//     and the few names line3D.cc mentions are declared here.  The code that uses them (segment detection, a
//     4-view triangulation whose result the reference discards) is never reached by the harness.
//   * L3D_REF_DETMATH: the libm calls on decision paths (expf, acos(float), acos(double), sin) are routed to the
//     deterministic functions of oracle/detmath.h -- the same the restatement and the CUDA kernels use -- so that
//     the comparison with the restatement can be bit for bit.  Without it the build calls glibc.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <list>
#include <map>
#include <math.h>
#include <queue>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "l3d_standin_boost.h"
#include "l3d_standin_eigen.h"
#include "l3d_standin_opencv.h"

// ---- what line3D.cc names from LineDescriptor.hh (src/line3D.cc:327-384) ----
struct OctaveSingleLine {
    float startPointX, startPointY, endPointX, endPointY;
    float sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
    float direction, salience, lineLength;
    unsigned int numOfPixels, octaveCount;
    std::vector<float> descriptor;
};
typedef std::vector<OctaveSingleLine> LinesVec;
typedef std::vector<LinesVec> ScaleLines;
class LineDescriptor {
  public:
    int GetLineDescriptor(cv::Mat& image, ScaleLines& keyLines);
};

// ---- what line3D.cc names from Converter.h (src/line3D.cc:2182-2221) ----
namespace ORB_SLAM2 {
class Converter {
  public:
    static cv::Mat toCvMat(const Eigen::Matrix<double, 3, 4>& m);
    static cv::Mat toCvMat(const Eigen::Matrix3d& m);
    static Eigen::Vector3d toVector3d(const cv::Mat& m);
};
}  // namespace ORB_SLAM2

#ifdef L3D_REF_DETMATH
#include "../detmath.h"
namespace l3d_ref {
inline float r_expf(float x) { return orc_expf(x); }
inline float r_acos(float x) { return orc_acosf(x); }
inline double r_acos(double x) { return orc_acos(x); }
inline double r_acos(int x) { return orc_acos((double)x); }
inline double r_sin(double x) { return orc_sin(x); }
}  // namespace l3d_ref
#define expf(x) ::l3d_ref::r_expf(x)
#define acos(x) ::l3d_ref::r_acos(x)
#define sin(x) ::l3d_ref::r_sin(x)
#endif
