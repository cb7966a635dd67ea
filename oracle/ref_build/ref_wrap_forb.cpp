// TEST INFRASTRUCTURE (oracle/_ref build only): DBoW2::FORB::distance (3rdparty/DBoW2/src/DBoW2/FORB.cpp:78-101), the
// reference's third Hamming primitive, compiled unmodified from /root/reference.
#include "FORB.h"

extern "C" double ref_forb_distance(const unsigned char* a, const unsigned char* b) {
  cv::Mat ma = cv::Mat(1, 32, CV_8UC1, (void*)a).clone(), mb = cv::Mat(1, 32, CV_8UC1, (void*)b).clone();
  return DBoW2::FORB::distance(ma, mb);
}
