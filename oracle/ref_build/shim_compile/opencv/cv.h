// TEST INFRASTRUCTURE: <opencv/cv.h> of OpenCV 3 (include/keyFrame.h:26) -> the OpenCV stand-in of oracle/ref_build.
#include <opencv2/core.hpp>
#include <opencv2/imgproc.hpp>
