// TEST INFRASTRUCTURE (tests/test_shim_compile.py): the reference's src/keyFrame.cpp, compiled UNMODIFIED against
// pl-slam_b200/cpp/stvo_shim.h, is linked with this driver, which checks what KeyFrame::KeyFrame (src/keyFrame.cpp:39-53,
// :63-77) promises: a deep copy of the frame (images, four descriptor blocks, safeCopy() of every feature) and the pose
// fields.  No GPU is involved: StereoFrame is a plain value holder.
#include <cstdio>
#include <cstdlib>

#include "keyFrame.h"

#define CHECK(c) do { if (!(c)) { std::printf("FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main() {
  StVO::PinholeStereoCamera cam(64, 48, 50.0, 50.0, 32.0, 24.0, 0.1);
  cv::Mat l(48, 64, CV_8UC1, cv::Scalar(7)), r(48, 64, CV_8UC1, cv::Scalar(9));
  StVO::StereoFrame* sf = new StVO::StereoFrame(l, r, 5, &cam);
  sf->Tfw = StVO::expmap_se3([] { StVO::Vector6d x; x(0) = 0.3; x(1) = -0.1; x(2) = 1.2; x(3) = 0.02; x(4) = -0.03; x(5) = 0.01; return x; }());
  sf->Tfw_cov = StVO::Matrix6d::Identity();
  for (int i = 0; i < 3; ++i) {
    StVO::PointFeature* p = new StVO::PointFeature(StVO::Vector3d(i, 2 * i, 5.0), StVO::Vector2d(10.0 + i, 20.0));
    p->pl = StVO::Vector2d(11.0 + i, 21.0);
    p->idx = i;
    sf->stereo_pt.push_back(p);
  }
  StVO::LineFeature* q = new StVO::LineFeature();
  q->spl = StVO::Vector2d(1, 2); q->epl = StVO::Vector2d(30, 40); q->idx = 4;
  sf->stereo_ls.push_back(q);
  sf->pdesc_l.rows = 3; sf->pdesc_l.data.assign(3 * 32, 0xAB);
  sf->ldesc_l.rows = 1; sf->ldesc_l.data.assign(32, 0xCD);

  PLSLAM::KeyFrame* kf = new PLSLAM::KeyFrame(sf, 3);
  CHECK(kf->kf_idx == 3);
  CHECK(kf->T_kf_w == sf->Tfw);
  CHECK(kf->xcov_kf_w == sf->Tfw_cov);
  StVO::Vector6d x = StVO::logmap_se3(sf->Tfw);
  for (int i = 0; i < 6; ++i) CHECK(kf->x_kf_w(i) == x(i));
  CHECK(std::abs(x(0) - 0.3) < 1e-12 && std::abs(x(5) - 0.01) < 1e-12);   // logmap(expmap(x)) == x
  StVO::StereoFrame* c = kf->stereo_frame;
  CHECK(c != sf && c->cam == sf->cam && c->frame_idx == 3);
  CHECK(c->stereo_pt.size() == 3 && c->stereo_ls.size() == 1);
  for (int i = 0; i < 3; ++i) {
    CHECK(c->stereo_pt[i] != sf->stereo_pt[i]);                       // safeCopy(): new objects ...
    CHECK(c->stereo_pt[i]->pl(0) == sf->stereo_pt[i]->pl(0) && c->stereo_pt[i]->idx == i && c->stereo_pt[i]->P(2) == 5.0);   // ... same values
  }
  CHECK(c->stereo_ls[0] != sf->stereo_ls[0] && c->stereo_ls[0]->epl(1) == 40.0 && c->stereo_ls[0]->idx == 4);
  CHECK(c->pdesc_l.rows == 3 && c->pdesc_l.data == sf->pdesc_l.data && c->pdesc_l.data.data() != sf->pdesc_l.data.data());
  CHECK(c->ldesc_l.rows == 1 && c->ldesc_l.row(0)[31] == 0xCD);
  CHECK(c->img_l.rows == 48 && c->img_l.cols == 64 && c->img_l.data[100] == 7 && c->img_r.data[100] == 9);
  CHECK(c->inv_width == STVO_GRID_COLS / 64.0 && c->inv_height == STVO_GRID_ROWS / 48.0);
  delete kf;          // ~KeyFrame deletes its own frame copy (src/keyFrame.cpp:80-83) ...
  CHECK(sf->stereo_pt[2]->idx == 2);   // ... and not the source frame's features
  delete sf;
  PLSLAM::KeyFrame* kf2 = nullptr;
  {
    StVO::StereoFrame* s2 = new StVO::StereoFrame(l, r, 9, &cam);
    kf2 = new PLSLAM::KeyFrame(s2);    // the one-argument overload (:39): kf_idx = -1
    delete s2;
  }
  CHECK(kf2->kf_idx == -1 && kf2->stereo_frame->stereo_pt.empty());
  delete kf2;
  std::printf("keyframe-shim ok\n");
  return 0;
}
