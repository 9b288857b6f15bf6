// kf_host.cu -- peak tracking filter (SURVEY.md section 8f "next" #4), host code only.
//
// Restates the reference's KalmanFilter3D (PC/src/kf.hpp:36-165, wrapped by PC/src/kf.pyx:18-46
// as lib.kf.CyKF): a linear constant-velocity filter on (x, y, z) in float32 with
// A = [[I, I], [0, I]], Q = 0.1 I6, H = [I3 0], R = 0.1 I3, P0 = I6, x0 = 0.  It smooths the
// arg-max of the power map before it is turned into a steering offset (visual.py:64-72).
// The reference needs Eigen, absent from this image; the test suite compiles the reference's header unmodified
// against a minimal stand-in (oracle/eigen_shim) and this file agrees with it to 2e-6 relative over 300 updates
// (tests/test_tracking_and_capture.py); plain loops here.
// predict(N) keeps the reference's quirk: the transition applied in step i is A^(i+1)
// (kf.hpp:121-125 multiplies An by A after every step).
#include <string.h>

#include "bf_common.cuh"

namespace {

struct Kf {
    float A[6][6], Q[6][6], P[6][6], x[6];
    float R;
    Kf()
    {
        memset(this, 0, sizeof(*this));
        for (int i = 0; i < 6; i++) { A[i][i] = 1.0f; Q[i][i] = 0.1f; P[i][i] = 1.0f; }
        for (int i = 0; i < 3; i++) A[i][i + 3] = 1.0f;
        R = 0.1f;
    }
};

void matmul6(const float a[6][6], const float b[6][6], float out[6][6])
{
    float t[6][6];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            float s = 0.0f;
            for (int k = 0; k < 6; k++) s += a[i][k] * b[k][j];
            t[i][j] = s;
        }
    memcpy(out, t, sizeof(t));
}

void matvec6(const float a[6][6], const float v[6], float out[6])
{
    float t[6];
    for (int i = 0; i < 6; i++) {
        float s = 0.0f;
        for (int k = 0; k < 6; k++) s += a[i][k] * v[k];
        t[i] = s;
    }
    memcpy(out, t, sizeof(t));
}

// kf.hpp:86-101
void kf_update(Kf &f, const float m[3])
{
    float At[6][6], AP[6][6];
    matvec6(f.A, f.x, f.x);
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) At[i][j] = f.A[j][i];
    matmul6(f.A, f.P, AP);
    matmul6(AP, At, f.P);
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) f.P[i][j] += f.Q[i][j];
    // S = H P H^T + R = P[0:3][0:3] + R I ; inverse by cofactors (what Eigen does for 3x3)
    float S[3][3], Si[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) S[i][j] = f.P[i][j] + (i == j ? f.R : 0.0f);
    const float c00 = S[1][1] * S[2][2] - S[1][2] * S[2][1];
    const float c01 = S[1][2] * S[2][0] - S[1][0] * S[2][2];
    const float c02 = S[1][0] * S[2][1] - S[1][1] * S[2][0];
    const float inv_det = 1.0f / (S[0][0] * c00 + S[0][1] * c01 + S[0][2] * c02);
    Si[0][0] = c00 * inv_det;
    Si[1][0] = c01 * inv_det;
    Si[2][0] = c02 * inv_det;
    Si[0][1] = (S[0][2] * S[2][1] - S[0][1] * S[2][2]) * inv_det;
    Si[1][1] = (S[0][0] * S[2][2] - S[0][2] * S[2][0]) * inv_det;
    Si[2][1] = (S[0][1] * S[2][0] - S[0][0] * S[2][1]) * inv_det;
    Si[0][2] = (S[0][1] * S[1][2] - S[0][2] * S[1][1]) * inv_det;
    Si[1][2] = (S[0][2] * S[1][0] - S[0][0] * S[1][2]) * inv_det;
    Si[2][2] = (S[0][0] * S[1][1] - S[0][1] * S[1][0]) * inv_det;
    // K = P H^T S^-1 = P[:, 0:3] S^-1
    float K[6][3];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 3; j++) {
            float s = 0.0f;
            for (int k = 0; k < 3; k++) s += f.P[i][k] * Si[k][j];
            K[i][j] = s;
        }
    float y[3];
    for (int i = 0; i < 3; i++) y[i] = m[i] - f.x[i];
    for (int i = 0; i < 6; i++) {
        float s = 0.0f;
        for (int k = 0; k < 3; k++) s += K[i][k] * y[k];
        f.x[i] += s;
    }
    // P = (I - K H) P
    float IKH[6][6];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) IKH[i][j] = (i == j ? 1.0f : 0.0f) - (j < 3 ? K[i][j] : 0.0f);
    matmul6(IKH, f.P, f.P);
}

}  // namespace

extern "C" void *bf_kf_create(void) { return new Kf(); }

extern "C" void bf_kf_destroy(void *kf) { delete (Kf *)kf; }

extern "C" int bf_kf_update(void *kf, const float *measurement)
{
    bf::clear_error();
    if (!kf || !measurement) { bf::set_error(BF_ERR_ARG, "bf_kf_update: null argument"); return BF_ERR_ARG; }
    kf_update(*(Kf *)kf, measurement);
    return BF_OK;
}

extern "C" int bf_kf_get_state(void *kf, float *xyz)
{
    bf::clear_error();
    if (!kf || !xyz) { bf::set_error(BF_ERR_ARG, "bf_kf_get_state: null argument"); return BF_ERR_ARG; }
    memcpy(xyz, ((Kf *)kf)->x, 3 * sizeof(float));
    return BF_OK;
}

extern "C" int bf_kf_predict(void *kf, int n, float *xyz)
{
    bf::clear_error();
    if (!kf || !xyz || n < 0) { bf::set_error(BF_ERR_ARG, "bf_kf_predict: bad argument"); return BF_ERR_ARG; }
    const Kf &f = *(Kf *)kf;
    float An[6][6], xn[6];
    memcpy(An, f.A, sizeof(An));
    memcpy(xn, f.x, sizeof(xn));
    for (int i = 0; i < n; i++) {
        matvec6(An, xn, xn);
        matmul6(An, f.A, An);
    }
    memcpy(xyz, xn, 3 * sizeof(float));
    return BF_OK;
}
