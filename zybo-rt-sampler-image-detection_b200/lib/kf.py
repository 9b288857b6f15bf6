"""Host mirror of the reference's `lib.kf` (PC/src/kf.pyx:18-46 over PC/src/kf.hpp:36-165):
CyKF().update([x, y, z]) / get_state() / predict(n), float32, same quirks (see csrc/kf_host.cu)."""
import numpy as np

from . import _native


class CyKF:
    def __init__(self):
        self._L = _native.lib()
        self._h = self._L.bf_kf_create()

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.bf_kf_destroy(self._h)
            self._h = None

    def update(self, meas):
        m = np.array([meas[0], meas[1], meas[2]], np.float32)
        _native.check(self._L.bf_kf_update(self._h, _native.ptr(m)))

    def get_state(self):
        out = np.zeros(3, np.float32)
        _native.check(self._L.bf_kf_get_state(self._h, _native.ptr(out)))
        return out

    def predict(self, n: int):
        out = np.zeros(3, np.float32)
        _native.check(self._L.bf_kf_predict(self._h, int(n), _native.ptr(out)))
        return out
