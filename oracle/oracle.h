/* oracle/oracle.h — declarations of the CPU oracle (TEST INFRASTRUCTURE ONLY; see oracle/__init__.py). */
#ifndef PLF_ORACLE_H
#define PLF_ORACLE_H
#include <stdint.h>

/* Mirrors cv::line_descriptor::KeyLine (3rdparty/line_descriptor/include/line_descriptor/descriptor_custom.hpp:105-176)
 * and, field for field, plf_keyline in include/plslam_b200.h. */
typedef struct orc_keyline {
  float angle;
  int class_id;
  int octave;
  float ptx, pty;
  float response;
  float size;
  float startPointX, startPointY, endPointX, endPointY;
  float sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
  float lineLength;
  int numOfPixels;
} orc_keyline;

void orc_gaussian_kernel_q8(int ksize, double sigma, int* taps);
void orc_gaussian_blur_u8(const uint8_t* src, int w, int h, int ksize, double sigma, uint8_t* dst);
void orc_sobel3_i16(const uint8_t* src, int w, int h, int16_t* dx, int16_t* dy);
void orc_lbd_weights(double* gaussCoefL, double* gaussCoefG);
void orc_lbd_compute(const uint8_t* img, int w, int h, const orc_keyline* kls, int n, uint8_t* desc_bin,
                     float* desc_float);
int orc_keylines_from_segments(const float* segs, int m, int w, int h, double min_length, orc_keyline* out);
#endif
