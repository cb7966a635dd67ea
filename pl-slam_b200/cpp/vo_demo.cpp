// vo_demo — the VO part of the reference's frame loop (app/plslam_dataset.cpp:111-163) on top of the C++ shim.
// Usage: vo_demo <frames.bin> [orb_nfeatures lsd_nfeatures]
// frames.bin: int32 n, w, h; double fx, fy, cx, cy, b; then n x (left h*w bytes, right h*w bytes).
// Prints one line per frame: idx status n_stereo_pt n_stereo_ls n_inliers newKF Tfw(16 values, row-major).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "stvo_shim.h"

// The fast-matching branch of MapHandler::matchKF2KFPoints / matchKF2KFLines (src/mapHandler.cpp:251-271, :382-418)
// written against the shim: grid filled with grid.at(x, y).push_back(idx), GridWindow of +-ws cells, matchGrid.
// `vo_demo --grid-selftest` runs it on the features of one synthetic frame against themselves (every feature must then
// find itself) - a build-time check that the overloads instantiate and a usage example for INTEGRATION.md.
static int grid_selftest(StVO::StereoFrameHandler* h, const StVO::StereoFrame* fr, float nnr) {
  using namespace StVO;
  const double inv_w = fr->inv_width, inv_h = fr->inv_height;  // = GRID_COLS / width, GRID_ROWS / height
  std::vector<point_2d> pj_points;
  GridStructure grid(GRID_ROWS, GRID_COLS);
  for (size_t idx = 0; idx < fr->stereo_pt.size(); ++idx) {
    const PointFeature* pt = fr->stereo_pt[idx];
    pj_points.push_back(std::make_pair((int)(pt->pl[0] * inv_w), (int)(pt->pl[1] * inv_h)));
    grid.at((int)(pt->pl[0] * inv_w), (int)(pt->pl[1] * inv_h)).push_back((int)idx);
  }
  GridWindow w;
  w.width = std::make_pair(1, 1);
  w.height = std::make_pair(1, 1);
  std::vector<int> m12;
  const int np = matchGrid(h->ctx(), pj_points, fr->pdesc_l, grid, fr->pdesc_l, w, m12, nnr);
  int self = 0;
  for (size_t i = 0; i < m12.size(); ++i) self += m12[i] == (int)i;
  std::vector<line_2d> pj_lines;
  std::vector<std::pair<double, double>> directions(fr->stereo_ls.size());
  GridStructure lgrid(GRID_ROWS, GRID_COLS);
  std::list<point_2d> cells;
  for (size_t idx = 0; idx < fr->stereo_ls.size(); ++idx) {
    const LineFeature* ls = fr->stereo_ls[idx];
    std::pair<double, double>& v = directions[idx];
    v = std::make_pair((ls->epl[0] - ls->spl[0]) * inv_w, (ls->epl[1] - ls->spl[1]) * inv_h);
    normalize(v);
    getLineCoords(ls->spl[0] * inv_w, ls->spl[1] * inv_h, ls->epl[0] * inv_w, ls->epl[1] * inv_h, cells);
    for (const point_2d& p : cells) lgrid.at(p.first, p.second).push_back((int)idx);
    pj_lines.push_back(std::make_pair(std::make_pair((int)(ls->spl[0] * inv_w), (int)(ls->spl[1] * inv_h)),
                                      std::make_pair((int)(ls->epl[0] * inv_w), (int)(ls->epl[1] * inv_h))));
  }
  std::vector<int> l12;
  const int nl = matchGrid(h->ctx(), pj_lines, fr->ldesc_l, lgrid, fr->ldesc_l, directions, w, l12, nnr, 0.75);
  std::printf("grid-selftest points %zu matched %d self %d lines %zu matched %d\n", fr->stereo_pt.size(), np, self,
              fr->stereo_ls.size(), nl);
  return 0;
}

int main(int argc, char** argv) {
  bool grid_test = false;
  if (argc >= 2 && std::string(argv[argc - 1]) == "--grid-selftest") { grid_test = true; --argc; }
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
        if (grid_test) return grid_selftest(StVO_, cur, prm.min_ratio_12_p);
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
