"""Loader for the REAL reference build in oracle/_ref/<cfg>/ (see build_ref.py).

TEST INFRASTRUCTURE ONLY.  Gives tests and bench.py's reference arm access to
  * libref.so  -- the reference's C kernels (sizes frozen per cfg by config.h)
  * lib.directions -- the cythonized reference delay generator
Nothing here reads /root/reference at run time; only the prebuilt artefacts.
"""
import ctypes
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "_ref")


def available(cfg):
    return os.path.exists(os.path.join(ROOT, cfg, "libref.so"))


def general(cfg):
    with open(os.path.join(ROOT, cfg, "config.json")) as f:
        return json.load(f)["general"]


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def host_has_avx512():
    """True when this host can run the x86-64-v4 build (libref_v4.so)."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    fl = set(line.split(":", 1)[1].split())
                    return {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"} <= fl
    except OSError:
        pass
    return False


def v4_available(cfg):
    return os.path.exists(os.path.join(ROOT, cfg, "libref_v4.so")) and host_has_avx512()


class RefC:
    """ctypes view of one oracle/_ref/<cfg>/libref.so (process-global tables,
    exactly like the reference).  lib="libref_v4.so" selects the AVX-512 build."""

    def __init__(self, cfg, lib="libref.so"):
        self.cfg = cfg
        self.g = general(cfg)
        self.N = self.g["N_SAMPLES"]
        self.M = self.g["N_MICROPHONES"]
        self.D = self.g["MAX_RES_X"] * self.g["MAX_RES_Y"]
        self.T = self.g["N_TAPS"]
        self.L = ctypes.CDLL(os.path.join(ROOT, cfg, lib))
        self._keep = []

    def _sig(self, s):
        s = np.ascontiguousarray(s, np.float32)
        assert s.shape == (self.M, self.N), (s.shape, self.M, self.N)
        return s

    def mimo_pad(self, signals, mic_ids, whole):
        s, a = self._sig(signals), np.ascontiguousarray(mic_ids, np.int32)
        w = np.ascontiguousarray(whole, np.int32).ravel()
        self.L.load_coefficients_pad(_p(w), ctypes.c_int(w.size))
        img = np.zeros(self.D, np.float32)
        self.L.mimo_pad(_p(s), _p(img), _p(a), ctypes.c_int(len(a)))
        self.L.unload_coefficients_pad()
        return img

    def miso_pad(self, signals, mic_ids, whole, offset):
        s, a = self._sig(signals), np.ascontiguousarray(mic_ids, np.int32)
        w = np.ascontiguousarray(whole, np.int32).ravel()
        self.L.load_coefficients_pad(_p(w), ctypes.c_int(w.size))
        out = np.zeros(self.N, np.float32)
        self.L.miso_pad(_p(s), _p(out), _p(a), ctypes.c_int(len(a)), ctypes.c_int(int(offset)))
        self.L.unload_coefficients_pad()
        return out

    def miso_pad2(self, signals, mic_ids, whole_by_mic):
        s, a = self._sig(signals), np.ascontiguousarray(mic_ids, np.int32)
        w = np.ascontiguousarray(whole_by_mic, np.int32).ravel()
        self.L.load_coefficients_pad2(_p(w), ctypes.c_int(w.size))
        out = np.zeros(self.N, np.float32)
        self.L.miso_pad2(_p(s), _p(out), _p(a), ctypes.c_int(len(a)), ctypes.c_int(0))
        self.L.unload_coefficients_pad2()
        return out

    def mimo_lerp(self, signals, mic_ids, delays_f32):
        s, a = self._sig(signals), np.ascontiguousarray(mic_ids, np.int32)
        d = np.ascontiguousarray(delays_f32, np.float32).ravel()
        self.L.load_coefficients_lerp(_p(d), ctypes.c_int(d.size))
        img = np.zeros(self.D, np.float32)
        self.L.mimo_lerp(_p(s), _p(img), _p(a), ctypes.c_int(len(a)))
        self.L.unload_coefficients_lerp()
        return img

    def miso_lerp(self, signals, mic_ids, delays_f32, offset):
        s, a = self._sig(signals), np.ascontiguousarray(mic_ids, np.int32)
        d = np.ascontiguousarray(delays_f32, np.float32).ravel()
        self.L.load_coefficients_lerp(_p(d), ctypes.c_int(d.size))
        out = np.zeros(self.N, np.float32)
        self.L.miso_lerp(_p(s), _p(out), _p(a), ctypes.c_int(len(a)), ctypes.c_int(int(offset)))
        self.L.unload_coefficients_lerp()
        return out

    def lerp_tables(self, delays_f32):
        """whole/weight arrays as load_coefficients_lerp leaves them in its globals."""
        d = np.ascontiguousarray(delays_f32, np.float32).ravel()
        self.L.load_coefficients_lerp(_p(d), ctypes.c_int(d.size))
        wp = ctypes.c_void_p.in_dll(self.L, "whole_samples_lerp").value
        fp = ctypes.c_void_p.in_dll(self.L, "fractional_samples_lerp").value
        whole = np.ctypeslib.as_array((ctypes.c_int32 * d.size).from_address(wp)).copy()
        weight = np.ctypeslib.as_array((ctypes.c_float * d.size).from_address(fp)).copy()
        self.L.unload_coefficients_lerp()
        return whole, weight

    def mimo_fir(self, signals, mic_ids, taps, lanes):
        s, a = self._sig(signals), np.ascontiguousarray(mic_ids, np.int32)
        h = np.ascontiguousarray(taps, np.float32).ravel()
        self.L.load_coefficients_convolve(_p(h), ctypes.c_int(h.size))
        img = np.zeros(self.D, np.float32)
        fn = self.L.mimo_convolve_vectorized if lanes else self.L.mimo_convolve_naive
        fn(_p(s), _p(img), _p(a), ctypes.c_int(len(a)))
        self.L.unload_coefficients_convolve()
        return img

    def hybrid_tables(self, delays_f32):
        d = np.ascontiguousarray(delays_f32, np.float32).ravel()
        self.L.load_coefficients_convolve_hybrid(_p(d), ctypes.c_int(d.size))
        wp = ctypes.c_void_p.in_dll(self.L, "whole_samples_convolve").value
        fp = ctypes.c_void_p.in_dll(self.L, "convolve_coefficients_fractional").value
        whole = np.ctypeslib.as_array((ctypes.c_int32 * d.size).from_address(wp)).copy()
        taps = np.ctypeslib.as_array((ctypes.c_float * (d.size * self.T)).from_address(fp)).copy()
        return whole, taps  # (the reference's unload double-frees; leak instead)

    def mimo_hybrid(self, signals, mic_ids, delays_f32):
        s, a = self._sig(signals), np.ascontiguousarray(mic_ids, np.int32)
        d = np.ascontiguousarray(delays_f32, np.float32).ravel()
        self.L.load_coefficients_convolve_hybrid(_p(d), ctypes.c_int(d.size))
        img = np.zeros(self.D, np.float32)
        self.L.mimo_convolve_hybrid(_p(s), _p(img), _p(a), ctypes.c_int(len(a)))
        return img


_dir_cache = {}


def directions(cfg):
    """The cythonized reference `directions` module for this cfg."""
    if cfg in _dir_cache:
        return _dir_cache[cfg]
    libdir = os.path.join(ROOT, cfg, "lib")
    cand = [f for f in os.listdir(libdir) if f.startswith("directions") and f.endswith(".so")]
    if not cand:
        raise FileNotFoundError("no directions extension in " + libdir)
    spec = importlib.util.spec_from_file_location("directions", os.path.join(libdir, cand[0]))
    mod = importlib.util.module_from_spec(spec)
    cwd = os.getcwd()
    os.chdir(os.path.join(ROOT, cfg))      # active_microphones() looks for unused_mics.npy in cwd
    try:
        spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
    _dir_cache[cfg] = mod
    return mod


def run_ref_wrappers(signals):
    """Run the reference's own Python wrappers (benchmark.pyx, built as lib.tests:
    mimo_pad_wrapper / mimo_lerp_wrapper) on `signals`, default cfg only.

    They import `lib.directions` by package name, which would collide with this
    repo's own `lib` package, so they run in a clean subprocess with
    cwd = oracle/_ref/default (exactly how PC/plot.py runs from PC/)."""
    import subprocess
    import tempfile
    base = os.path.join(ROOT, "default")
    with tempfile.TemporaryDirectory() as tmp:
        np.save(os.path.join(tmp, "sig.npy"), np.ascontiguousarray(signals, np.float32))
        code = (
            "import numpy as np, sys\n"
            "sys.path.insert(0, %r)\n"
            "from lib.tests import mimo_pad_wrapper, mimo_lerp_wrapper\n"
            "s = np.load(%r)\n"
            "np.savez(%r, pad=np.asarray(mimo_pad_wrapper(s)), lerp=np.asarray(mimo_lerp_wrapper(s)))\n"
        ) % (base, os.path.join(tmp, "sig.npy"), os.path.join(tmp, "out.npz"))
        subprocess.run([sys.executable, "-c", code], cwd=base, check=True,
                       stdout=subprocess.DEVNULL)
        z = np.load(os.path.join(tmp, "out.npz"))
        return z["pad"], z["lerp"]


class RefReceiver:
    """The reference's UDP receiver (PC/src/receiver.c, built as oracle/_ref/<cfg>/librecv.so) fed through a
    socketpair: receive_and_write_to_buffer() turns N_SAMPLES datagrams (receiver.h:51-59) into the
    [mic][sample] float buffer.  On the last odd row the reference reads one int PAST the payload
    (receiver.c:140, `row + COLUMNS - x` with x = 0); the message buffer handed to it here is followed
    by one extra int, `past_end`, so that read is defined and chosen by the test."""

    def __init__(self, cfg):
        self.g = general(cfg)
        self.N, self.M = self.g["N_SAMPLES"], self.g["N_MICROPHONES"]
        self.L = ctypes.CDLL(os.path.join(ROOT, cfg, "librecv.so"))

    @staticmethod
    def available(cfg):
        return os.path.exists(os.path.join(ROOT, cfg, "librecv.so"))

    def receive(self, stream, n_arrays, counter0=0, frequency=48828, version=2, past_end=0, ring=True):
        """stream: int32 [N_SAMPLES][N_MICROPHONES] payloads -> float32 [N_MICROPHONES][N_SAMPLES]."""
        import socket
        import struct
        import threading
        stream = np.ascontiguousarray(stream, np.int32)
        assert stream.shape == (self.N, self.M)
        a, b = socket.socketpair(socket.AF_UNIX, socket.SOCK_DGRAM)
        try:
            def feed():
                for i in range(self.N):
                    a.send(struct.pack("<Hbbi", frequency, n_arrays, version, counter0 + i) + stream[i].tobytes())
            th = threading.Thread(target=feed)
            th.start()
            msg = (ctypes.c_int32 * (2 + self.M + 1))()
            msg[2 + self.M] = past_end
            if ring:      # ring_buffer: int index; float data[4 * BUFFER_LENGTH]; int counter  (receiver.h:31-36)
                rb = (ctypes.c_float * (1 + 4 * self.M * self.N + 1))()
                rc = self.L.receive_and_write_to_buffer(ctypes.c_int(b.fileno()), rb, msg, ctypes.c_int(n_arrays))
                assert rc == 0
                out = np.frombuffer(rb, np.float32, self.M * self.N, 4).copy()
            else:
                buf = np.zeros(self.M * self.N, np.float32)
                self.L.receive_to_buffer(ctypes.c_int(b.fileno()), _p(buf), msg, ctypes.c_int(n_arrays))
                out = buf
            th.join()
        finally:
            a.close()
            b.close()
        return out.reshape(self.M, self.N)


class RefKalman:
    """The reference's KalmanFilter3D (PC/src/kf.hpp, compiled unmodified against oracle/eigen_shim -- Eigen itself
    is absent here -- behind oracle/kf_ref_wrap.cpp).  Same three calls as PC/src/kf.pyx:18-46."""

    PATH = os.path.join(ROOT, "default", "libkf_ref.so")

    @staticmethod
    def available():
        return os.path.exists(RefKalman.PATH)

    def __init__(self):
        self.L = ctypes.CDLL(RefKalman.PATH)
        self.L.kfref_create.restype = ctypes.c_void_p
        for f in (self.L.kfref_destroy, self.L.kfref_update, self.L.kfref_get_state, self.L.kfref_predict):
            f.restype = None
        self.h = ctypes.c_void_p(self.L.kfref_create())

    def __del__(self):
        if getattr(self, "h", None):
            self.L.kfref_destroy(self.h)
            self.h = None

    def update(self, m):
        a = np.ascontiguousarray(m, np.float32)
        self.L.kfref_update(self.h, _p(a))

    def get_state(self):
        out = np.zeros(3, np.float32)
        self.L.kfref_get_state(self.h, _p(out))
        return out

    def predict(self, n):
        out = np.zeros(3, np.float32)
        self.L.kfref_predict(self.h, ctypes.c_int(int(n)), _p(out))
        return out
