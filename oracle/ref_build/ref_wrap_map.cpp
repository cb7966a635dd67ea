// TEST INFRASTRUCTURE (oracle/_ref build only): plain-C entry point around the reference's own, unmodified
// PLSLAM::MapPoint (include/mapFeatures.h, src/mapFeatures.cpp, compiled from /root/reference), to pin
// oracle/mapfeatures.py / plf_median_descriptors against it:
//   ref_median_descriptor : MapPoint(...) for the first observation, addMapPointObservation for the others (each runs
//                           updateAverageDescDir, src/mapFeatures.cpp:40-93); reports which observation's descriptor ended
//                           up in med_desc and med_obs_dir.
// Eigen is a stand-in here (oracle/ref_build/eigen_stub) that zero-initialises fixed-size vectors; the reference's
// direction accumulator is uninitialised in a real build, so only max_idx is a statement about the reference.
#include "mapFeatures.h"

extern "C" int ref_median_descriptor(const unsigned char* desc, int n, const double* dirs, int* max_idx, double* med_dir) {
  if (n < 1) return -1;
  std::vector<cv::Mat> rows;
  for (int i = 0; i < n; ++i) rows.push_back(cv::Mat(1, 32, CV_8UC1, (void*)(desc + 32 * i)).clone());
  Eigen::Vector3d d0(dirs[0], dirs[1], dirs[2]);
  PLSLAM::MapPoint mp(0, Eigen::Vector3d(0, 0, 1), rows[0], 0, Eigen::Vector2d(0, 0), d0);
  for (int i = 1; i < n; ++i)
    mp.addMapPointObservation(rows[i], i, Eigen::Vector2d(0, 0), Eigen::Vector3d(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));
  *max_idx = -1;
  for (int i = 0; i < n; ++i)
    if (mp.med_desc.data == mp.desc_list[i].data) { *max_idx = i; break; }   // med_desc = desc_list[max_idx] shares its buffer
  for (int k = 0; k < 3; ++k) med_dir[k] = mp.med_obs_dir(k);
  return 0;
}
