// gf2_tracker.cu — pyramidal Lucas-Kanade tracker (cv::calcOpticalFlowPyrLK replacement). Placeholder until the
// kernels land: every entry point reports GF2_ERR_UNSUPPORTED.
#include "gf2_common.h"
extern "C" {
int gf2_tracker_create(const gf2_tracker_cfg* cfg, gf2_tracker** out) { (void)cfg; (void)out; return gf2::fail(GF2_ERR_UNSUPPORTED, "tracker not built yet"); }
void gf2_tracker_destroy(gf2_tracker* h) { (void)h; }
int gf2_tracker_track(gf2_tracker* h, int n_streams, const uint8_t* prev, const uint8_t* cur, size_t stride, const int32_t* n_pts,
                      const float* prev_pts, float* cur_pts, uint8_t* status, float* err, int flags, int max_level) {
  (void)h; (void)n_streams; (void)prev; (void)cur; (void)stride; (void)n_pts; (void)prev_pts; (void)cur_pts; (void)status; (void)err; (void)flags; (void)max_level;
  return gf2::fail(GF2_ERR_UNSUPPORTED, "tracker not built yet");
}
int gf2_tracker_track_fb(gf2_tracker* h, int n_streams, const uint8_t* prev, const uint8_t* cur, size_t stride, const int32_t* n_pts,
                         const float* prev_pts, float* cur_pts, uint8_t* status, int max_level) {
  (void)h; (void)n_streams; (void)prev; (void)cur; (void)stride; (void)n_pts; (void)prev_pts; (void)cur_pts; (void)status; (void)max_level;
  return gf2::fail(GF2_ERR_UNSUPPORTED, "tracker not built yet");
}
int gf2_tracker_last_timing(gf2_tracker* h, double out[8]) { (void)h; (void)out; return gf2::fail(GF2_ERR_UNSUPPORTED, "tracker not built yet"); }
}
