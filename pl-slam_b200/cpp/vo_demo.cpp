// vo_demo — the VO part of the reference's frame loop (app/plslam_dataset.cpp:111-163) on top of the C++ shim.
// Usage: vo_demo <frames.bin> [orb_nfeatures lsd_nfeatures]
// frames.bin: int32 n, w, h; double fx, fy, cx, cy, b; then n x (left h*w bytes, right h*w bytes).
// Prints one line per frame: idx status n_stereo_pt n_stereo_ls n_inliers newKF Tfw(16 values, row-major).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "stvo_shim.h"

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s frames.bin [orb_nfeatures lsd_nfeatures]\n", argv[0]);
    return 1;
  }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) { std::perror("open"); return 1; }
  int32_t hdr[3];
  double cam[5];
  if (std::fread(hdr, 4, 3, f) != 3 || std::fread(cam, 8, 5, f) != 5) { std::fprintf(stderr, "bad header\n"); return 1; }
  const int n = hdr[0], w = hdr[1], h = hdr[2];
  std::vector<uint8_t> buf((size_t)2 * w * h);
  try {
    StVO::PinholeStereoCamera cam_pin(w, h, cam[0], cam[1], cam[2], cam[3], cam[4]);
    plf_params prm;
    plf_default_params(&prm);
    if (argc >= 4) { prm.orb_nfeatures = std::atoi(argv[2]); prm.lsd_nfeatures = std::atoi(argv[3]); }
    StVO::StereoFrameHandler* StVO_ = new StVO::StereoFrameHandler(&cam_pin, &prm);   // app:109
    for (int frame_counter = 0; frame_counter < n; ++frame_counter) {                   // app:111
      if (std::fread(buf.data(), 1, buf.size(), f) != buf.size()) { std::fprintf(stderr, "short read\n"); return 1; }
      plf::Image img_l{buf.data(), w, h, w}, img_r{buf.data() + (size_t)w * h, w, h, w};
      bool new_kf = false;
      StVO::StereoFrame* cur;
      if (frame_counter == 0) {
        StVO_->initialize(img_l, img_r, 0);                                             // app:115
        cur = StVO_->prev_frame;
      } else {
        StVO_->insertStereoPair(img_l, img_r, frame_counter);                           // app:127
        StVO_->optimizePose();                                                          // app:128
        cur = StVO_->curr_frame;
        if (StVO_->needNewKF()) {                                                       // app:135
          new_kf = true;
          StVO_->currFrameIsKF();                                                       // app:145
        }
      }
      const plf_frame_result& r = StVO_->last_result();
      std::printf("%d %d %zu %zu %d %d", frame_counter, r.status, cur->stereo_pt.size(), cur->stereo_ls.size(),
                  StVO_->n_inliers, new_kf ? 1 : 0);
      for (int i = 0; i < 16; ++i) std::printf(" %.17g", cur->Tfw.v[i]);
      std::printf("\n");
      if (frame_counter > 0) StVO_->updateFrame();                                      // app:159
    }
    delete StVO_;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 2;
  }
  std::fclose(f);
  return 0;
}
