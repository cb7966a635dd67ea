// TEST INFRASTRUCTURE (oracle/_ref build only): plain-C entry points around the reference's own, unmodified
// 3rdparty/line_descriptor classes (compiled from /root/reference by oracle/ref_build/Makefile), so that the tests can
// pin oracle/lbd.c against the code the reference actually ships:
//   ref_keylines : LSDDetectorC::detect(image, keylines, scale, numOctaves, opts)      LSDDetector_custom.cpp:218-324
//   ref_lbd      : BinaryDescriptor::compute(image, keylines, descriptors, false)       binary_descriptor_custom.cpp:524-687, :1026-1372
// OpenCV itself is a stand-in here (oracle/ref_build/opencv_stub): its primitives on this path are the cv2-pinned
// restatements of oracle/*.c, so what these functions add over the oracle is exactly the vendored code.
#include <vector>

#include "precomp_custom.hpp"   // (brings in bitops_custom.hpp: the reference's own popcount distance, :83-96)
#include "line_descriptor_custom.hpp"
extern "C" {
#include "oracle.h"
}

using namespace cv;
using namespace cv::line_descriptor;

static void to_orc(const KeyLine& k, orc_keyline* o) {
  o->angle = k.angle; o->class_id = k.class_id; o->octave = k.octave; o->ptx = k.pt.x; o->pty = k.pt.y;
  o->response = k.response; o->size = k.size;
  o->startPointX = k.startPointX; o->startPointY = k.startPointY; o->endPointX = k.endPointX; o->endPointY = k.endPointY;
  o->sPointInOctaveX = k.sPointInOctaveX; o->sPointInOctaveY = k.sPointInOctaveY;
  o->ePointInOctaveX = k.ePointInOctaveX; o->ePointInOctaveY = k.ePointInOctaveY;
  o->lineLength = k.lineLength; o->numOfPixels = k.numOfPixels;
}
static KeyLine from_orc(const orc_keyline& o) {
  KeyLine k;
  k.angle = o.angle; k.class_id = o.class_id; k.octave = o.octave; k.pt = Point2f(o.ptx, o.pty);
  k.response = o.response; k.size = o.size;
  k.startPointX = o.startPointX; k.startPointY = o.startPointY; k.endPointX = o.endPointX; k.endPointY = o.endPointY;
  k.sPointInOctaveX = o.sPointInOctaveX; k.sPointInOctaveY = o.sPointInOctaveY;
  k.ePointInOctaveX = o.ePointInOctaveX; k.ePointInOctaveY = o.ePointInOctaveY;
  k.lineLength = o.lineLength; k.numOfPixels = o.numOfPixels;
  return k;
}

extern "C" int ref_keylines(const uint8_t* img, int w, int h, int scale_arg, int num_octaves, int refine, double scale,
                            double sigma_scale, double quant, double ang_th, double log_eps, double density_th, int n_bins,
                            double min_length, orc_keyline* out, int cap) {
  try {
    Mat image = Mat(h, w, CV_8UC1, (void*)img).clone();
    Ptr<LSDDetectorC> lsd = LSDDetectorC::createLSDDetectorC();
    LSDDetectorC::LSDOptions opts;
    opts.refine = refine; opts.scale = scale; opts.sigma_scale = sigma_scale; opts.quant = quant; opts.ang_th = ang_th;
    opts.log_eps = log_eps; opts.density_th = density_th; opts.n_bins = n_bins; opts.min_length = min_length;
    std::vector<KeyLine> kls;
    lsd->detect(image, kls, scale_arg, num_octaves, opts);
    if ((int)kls.size() > cap) return -2;
    for (size_t i = 0; i < kls.size(); ++i) to_orc(kls[i], out + i);
    return (int)kls.size();
  } catch (const std::exception& e) {
    fprintf(stderr, "ref_keylines: %s\n", e.what());
    return -1;
  }
}

extern "C" int ref_lbd(const uint8_t* img, int w, int h, const orc_keyline* kls, int n, uint8_t* desc) {
  try {
    Mat image = Mat(h, w, CV_8UC1, (void*)img).clone();
    std::vector<KeyLine> v;
    for (int i = 0; i < n; ++i) v.push_back(from_orc(kls[i]));
    Ptr<BinaryDescriptor> bd = BinaryDescriptor::createBinaryDescriptor();
    Mat d;
    bd->compute(image, v, d);
    if (d.rows != n || d.cols != 32 || d.type() != CV_8UC1) return -2;
    for (int i = 0; i < n; ++i) memcpy(desc + 32 * i, d.ptr(i), 32);
    return n;
  } catch (const std::exception& e) {
    fprintf(stderr, "ref_lbd: %s\n", e.what());
    return -1;
  }
}

// cv::line_descriptor::match(P, Q, codelb): the Hamming primitive of the reference (src/bitops_custom.hpp:83-96)
extern "C" int ref_hamming(const uint8_t* p, const uint8_t* q, int nbytes) {
  return cv::line_descriptor::match((UINT8*)p, (UINT8*)q, nbytes);
}
