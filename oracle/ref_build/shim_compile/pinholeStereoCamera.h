// The forwarding header INTEGRATION.md section 1 describes: stvo-pl's header name -> the shim (cv::Mat images).
#pragma once
#include <opencv2/core.hpp>
#define STVO_SHIM_WITH_OPENCV
#include "stvo_shim.h"
