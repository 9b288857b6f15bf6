"""float64 NumPy restatement of the reference's KalmanFilter3D (PC/src/kf.hpp:36-165).
TEST INFRASTRUCTURE ONLY.  Second check beside oracle/ref.py:RefKalman (the reference's own class compiled
against oracle/eigen_shim): the product (csrc/kf_host.cu, float32) is checked against this restatement to 1e-5."""
import numpy as np


class KalmanFilter3D:
    def __init__(self):
        self.A = np.eye(6)
        self.A[:3, 3:] = np.eye(3)               # kf.hpp:54-59
        self.Q = 0.1 * np.eye(6)                 # 61-66
        self.H = np.hstack([np.eye(3), np.zeros((3, 3))])   # 68-70
        self.R = 0.1 * np.eye(3)                 # 72-74
        self.P = np.eye(6)
        self.x = np.zeros(6)

    def update(self, m):                         # 86-101
        A, H = self.A, self.H
        self.x = A @ self.x
        self.P = A @ self.P @ A.T + self.Q
        S = H @ self.P @ H.T + self.R
        K = self.P @ H.T @ np.linalg.inv(S)
        y = np.asarray(m, float) - H @ self.x
        self.x = self.x + K @ y
        self.P = (np.eye(6) - K @ H) @ self.P

    def get_state(self):                         # 108-111
        return self.x[:3].copy()

    def predict(self, n):                        # 119-131: An grows every step
        An, xn = self.A.copy(), self.x.copy()
        for _ in range(n):
            xn = An @ xn
            An = An @ self.A
        return xn[:3]
