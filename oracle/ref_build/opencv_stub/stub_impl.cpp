// TEST INFRASTRUCTURE (oracle/_ref build only): the four OpenCV primitives on the LBD / KeyLine path, forwarded to the
// cv2-pinned C restatements (oracle/lbd.c: orc_gaussian_blur_u8, orc_sobel3_i16; oracle/lsd.c: orc_lsd_detect).
#include "opencv2/imgproc.hpp"

namespace cv {
void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sigmaX, double, int) {
  Mat src = src_.getMat();
  if (src.type() != CV_8UC1 || !src.isContinuous() || ksize.width != ksize.height) plf_stub_abort("GaussianBlur (only CV_8UC1, square kernel)");
  Mat out(src.rows, src.cols, CV_8UC1);
  orc_gaussian_blur_u8(src.data, src.cols, src.rows, ksize.width, sigmaX, out.data);
  dst_.getMatRef() = out;
}
void Sobel(InputArray src_, OutputArray dst_, int ddepth, int dx, int dy, int ksize, double, double, int) {
  Mat src = src_.getMat();
  if (src.type() != CV_8UC1 || !src.isContinuous() || ksize != 3 || CV_MAT_DEPTH(ddepth) != CV_16S || dx + dy != 1) plf_stub_abort("Sobel (only CV_8UC1 -> CV_16S, 3x3, first order)");
  Mat gx(src.rows, src.cols, CV_16SC1), gy(src.rows, src.cols, CV_16SC1);
  orc_sobel3_i16(src.data, src.cols, src.rows, (int16_t*)gx.data, (int16_t*)gy.data);
  dst_.getMatRef() = dx == 1 ? gx : gy;
}

class LsdImpl : public LineSegmentDetector {
 public:
  int refine; double scale, sigma_scale, quant, ang_th; int n_bins;
  void detect(InputArray image, OutputArray lines, OutputArray, OutputArray, OutputArray) override {
    Mat img = image.getMat();
    if (refine != LSD_REFINE_NONE || img.type() != CV_8UC1 || !img.isContinuous() || !lines.v4f) plf_stub_abort("LineSegmentDetector (only refine 0, CV_8UC1, vector<Vec4f> output)");
    std::vector<float> segs(4 * 65536);
    const int n = orc_lsd_detect(img.data, img.cols, img.rows, scale, sigma_scale, quant, ang_th, n_bins, 0, 0, segs.data(), 65536);
    if (n < 0) plf_stub_abort("LineSegmentDetector capacity");
    lines.v4f->clear();
    for (int i = 0; i < n; ++i) lines.v4f->push_back(Vec4f(segs[4 * i], segs[4 * i + 1], segs[4 * i + 2], segs[4 * i + 3]));
  }
};
Ptr<LineSegmentDetector> createLineSegmentDetector(int refine, double scale, double sigma_scale, double quant, double ang_th, double, double, int n_bins) {
  LsdImpl* p = new LsdImpl();
  p->refine = refine; p->scale = scale; p->sigma_scale = sigma_scale; p->quant = quant; p->ang_th = ang_th; p->n_bins = n_bins;
  return Ptr<LineSegmentDetector>(p);
}
}  // namespace cv
