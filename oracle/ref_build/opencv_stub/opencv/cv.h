// TEST INFRASTRUCTURE (oracle/_ref build only): <opencv/cv.h> as included by the reference's include/mapFeatures.h.
#include "opencv2/core.hpp"
#include "opencv2/imgproc.hpp"
