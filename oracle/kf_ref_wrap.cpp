// oracle/kf_ref_wrap.cpp -- TEST INFRASTRUCTURE.  C entry points around the REFERENCE's KalmanFilter3D, whose header
// (PC/src/kf.hpp) is included as it lies under the reference tree and compiled against oracle/eigen_shim (Eigen is not
// installed here; see that file for what the stand-in does and does not pin).  Built by oracle/build_ref.py:build_kf
// into oracle/_ref/libkf_ref.so; mirrors what PC/src/kf.pyx:18-46 exposes (update / get_state / predict).
#include "kf.hpp"

extern "C" {
void *kfref_create() { return new KalmanFilter3D(); }
void kfref_destroy(void *p) { delete (KalmanFilter3D *)p; }
void kfref_update(void *p, const float *m3) { ((KalmanFilter3D *)p)->updatef(std::vector<float>{m3[0], m3[1], m3[2]}); }
void kfref_get_state(void *p, float *out3)
{
    std::vector<float> s = ((KalmanFilter3D *)p)->getStatef();
    out3[0] = s[0]; out3[1] = s[1]; out3[2] = s[2];
}
void kfref_predict(void *p, int n, float *out3)
{
    std::vector<float> s = ((KalmanFilter3D *)p)->predictf(n);
    out3[0] = s[0]; out3[1] = s[1]; out3[2] = s[2];
}
}
