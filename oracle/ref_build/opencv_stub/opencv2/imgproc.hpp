// TEST INFRASTRUCTURE (oracle/_ref build only): the imgproc entry points the vendored line_descriptor calls.
// Functional: GaussianBlur (CV_8U), Sobel (3x3 -> CV_16S), LineIterator::count, createLineSegmentDetector - each
// forwarded to the cv2-pinned C restatements in oracle/*.c.  The rest compiles and aborts when called.
#ifndef PLF_OPENCV_STUB_IMGPROC_HPP
#define PLF_OPENCV_STUB_IMGPROC_HPP
#include "opencv2/core.hpp"
extern "C" {
#include "oracle.h"
}
namespace cv {
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
enum { COLOR_BGR2GRAY = 6, COLOR_RGB2GRAY = 7, COLOR_GRAY2BGR = 8 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };
enum { THRESH_BINARY = 0, THRESH_BINARY_INV = 1, THRESH_TRUNC = 2, THRESH_TOZERO = 3, THRESH_TOZERO_INV = 4 };
enum { LSD_REFINE_NONE = 0, LSD_REFINE_STD = 1, LSD_REFINE_ADV = 2 };

void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void Sobel(InputArray src, OutputArray dst, int ddepth, int dx, int dy, int ksize = 3, double scale = 1, double delta = 0, int borderType = BORDER_DEFAULT);
inline void pyrDown(InputArray, OutputArray, const Size& = Size(), int = BORDER_DEFAULT) { plf_stub_abort("cv::pyrDown"); }
inline void resize(InputArray, OutputArray, Size, double = 0, double = 0, int = INTER_LINEAR) { plf_stub_abort("cv::resize"); }
inline void cvtColor(InputArray, OutputArray, int, int = 0) { plf_stub_abort("cv::cvtColor"); }
inline double threshold(InputArray, OutputArray, double, double, int) { plf_stub_abort("cv::threshold"); }
// drawing + the OpenCV-1 colour codes (src/keyFrame.cpp:86-127 plotKeyFrame, only has to compile: tests/test_shim_compile.py)
#define CV_GRAY2BGR 8
#define CV_BGRA2BGR 1
inline void circle(InputOutputArray, Point, int, const Scalar&, int = 1, int = 8, int = 0) { plf_stub_abort("cv::circle"); }
inline void line(InputOutputArray, Point, Point, const Scalar&, int = 1, int = 8, int = 0) { plf_stub_abort("cv::line"); }

class LineIterator {
 public:
  int count;
  // 8-connected: count = max(|dx|, |dy|) + 1 on the integer end points (verified against cv2.line pixel counts, SURVEY
  // Appendix B; the vendored caller clamps the end points into the image first, so OpenCV's clipLine is a no-op)
  LineIterator(const Mat&, Point pt1, Point pt2, int connectivity = 8, bool = false) {
    const int dx = std::abs(pt2.x - pt1.x), dy = std::abs(pt2.y - pt1.y);
    count = (connectivity == 8 ? std::max(dx, dy) : dx + dy) + 1;
  }
};

class LineSegmentDetector : public Algorithm {
 public:
  virtual void detect(InputArray image, OutputArray lines, OutputArray width = noArray(), OutputArray prec = noArray(), OutputArray nfa = noArray()) = 0;
  virtual ~LineSegmentDetector() {}
};
Ptr<LineSegmentDetector> createLineSegmentDetector(int refine = LSD_REFINE_STD, double scale = 0.8, double sigma_scale = 0.6, double quant = 2.0,
                                                   double ang_th = 22.5, double log_eps = 0, double density_th = 0.7, int n_bins = 1024);
}  // namespace cv
#endif
