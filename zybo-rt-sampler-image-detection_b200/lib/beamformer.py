"""`lib.beamformer` -- the Python calling surface of the reference's Cython module
(PC/src/main.pyx), backed by the B200 library.

Kept: connect / disconnect / receive, the producer loops taking
(q: JoinableQueue, running: Value) -- uti_api, uti_api_with_miso, conv_api,
miso_api, just_miso_api, b, just_miso_loop -- and (q_steer, q_out, running) --
multi_pad, multi_lerp -- plus steer_cartesian_degree / stear_miso_beam.  Queue
payloads are what the reference's consumers expect: float32
C-contiguous (MAX_RES_X, MAX_RES_Y) maps, `(power_map, frame_nr)` tuples from `b`.

Not kept (outside the hot path): the UDP receiver, SysV shared
memory and PortAudio playback.  Their place is taken by a *data source*: any
callable filling a float32 (N_MICROPHONES, N_SAMPLES) array, registered with
connect(source=...) -- e.g. ArraySource (replays a record.py `.npy` capture) or a
ctypes pointer to the reference's own get_data().  The MISO audio child process
(api.c:491-543) becomes MisoBeam: steer() selects the table row, listen()
returns one beamformed block scaled like api.c:519-523.

CUDA is initialised lazily inside the process that runs the loop (the reference
forks its producers, main.pyx:702-721).
"""
import ctypes
import queue as _queue
import time

import numpy as np

from interface import config
from . import _native
from .directions import (active_microphones, calculate_coefficients, calculate_delay_miso,  # noqa: F401
                         calculate_delays, compute_convolve_h, whole_and_f32)

DTYPE_arr = np.float32

# ---------------------------------------------------------------------------
# data source (stands in for load()/get_data()/stop_receiving(), api.h:6-9)
# ---------------------------------------------------------------------------
_source = None
_source_c = None          # keeps the ctypes callback alive


class ArraySource:
    """Replays a (N_MICROPHONES, k*N_SAMPLES) float32 recording block by block
    (the format written by PC/record.py:28-46); wraps around at the end."""

    def __init__(self, recording):
        self.rec = np.ascontiguousarray(recording, dtype=np.float32)
        assert self.rec.shape[0] == config.N_MICROPHONES
        self.blocks = self.rec.shape[1] // config.N_SAMPLES
        assert self.blocks >= 1
        self.pos = 0

    def __call__(self, out):
        N = config.N_SAMPLES
        out[:, :] = self.rec[:, self.pos * N:(self.pos + 1) * N]
        self.pos = (self.pos + 1) % self.blocks


def connect(replay_mode: bool = False, verbose=True, source=None) -> None:
    """main.pyx:95-119.  `source(out)` fills a float32 (N_MICROPHONES, N_SAMPLES) array
    with the latest sample buffer; it replaces the forked UDP receiver."""
    global _source, _source_c
    assert isinstance(replay_mode, bool), "Replay mode must be either True or False"
    if source is None:
        raise RuntimeError("connect(): pass source=<callable filling (N_MICROPHONES, N_SAMPLES) "
                           "float32>; the UDP receiver of the reference is not part of this library")
    _native.configure_from(config)
    _source = source
    shape = (config.N_MICROPHONES, config.N_SAMPLES)

    def _fill(ptr):
        buf = np.ctypeslib.as_array(ptr, shape=shape)
        _source(buf)

    _source_c = _native.DATA_SOURCE_FN(_fill)
    _native.lib().bf_set_data_source(_source_c)
    if verbose:
        print("Data source registered.\nContinue your program!\n")


def disconnect() -> None:
    """main.pyx:122-130"""
    global _source, _source_c
    if _native._lib is not None:
        _native.lib().bf_set_data_source(ctypes.cast(None, _native.DATA_SOURCE_FN))
    _source = None
    _source_c = None


def receive(signals) -> None:
    """main.pyx:133-159: fill `signals` with the latest N_SAMPLES of every microphone."""
    assert signals.shape == (config.N_MICROPHONES, config.N_SAMPLES), "Arrays do not match shape"
    assert signals.dtype == np.float32, "Arrays dtype do not match"
    if _source is None:
        raise RuntimeError("receive(): not connected")
    _source(signals)


# ---------------------------------------------------------------------------
# table loading helpers (the prologue of every reference loop)
# ---------------------------------------------------------------------------
def _active():
    mics, n = active_microphones()
    return _native.i32(mics), int(n)


def _load_pad():
    _native.configure_from(config)
    whole, _ = whole_and_f32()                    # == calculate_coefficients()[0]
    w = _native.i32(whole)
    _native.lib().load_coefficients_pad(_native.ptr(w), int(w.size))
    _native.check()


def _load_lerp():
    _native.configure_from(config)
    _, d32 = whole_and_f32()                      # == float32(calculate_delays())
    _native.lib().load_coefficients_lerp(_native.ptr(d32), int(d32.size))
    _native.check()


def _new_map():
    return np.ascontiguousarray(np.zeros((config.MAX_RES_X, config.MAX_RES_Y), dtype=DTYPE_arr))


def _call_map(name, power_map, mics, n):
    getattr(_native.lib(), name)(_native.ptr(power_map), _native.ptr(mics), n)
    _native.check()


# ---------------------------------------------------------------------------
# MISO beam (replaces the forked audio child: load_miso/load_pa/steer/stop_miso)
# ---------------------------------------------------------------------------
class MisoBeam:
    """State of api.h:32-45's `Miso` record: steer offset + adaptive array."""

    def __init__(self):
        self.steer_offset = 0
        self.mics = None
        self.n = 1

    def load_pa(self, mics, n):
        self.mics, self.n = _native.i32(mics[:n]), int(n)

    def steer(self, offset):
        self.steer_offset = int(offset)

    def listen(self, scaled=True):
        """One block of beam audio: get_data -> miso_pad -> /n*MIC_GAIN (api.c:505-523)."""
        out = np.zeros(config.N_SAMPLES, dtype=DTYPE_arr)
        _native.lib().miso_steer_listen(_native.ptr(out), _native.ptr(self.mics), self.n,
                                        self.steer_offset)
        _native.check()
        if scaled:
            out /= np.float32(self.n)
            out *= np.float32(config.MIC_GAIN)
        return out


_miso = MisoBeam()


def load_miso():
    return 0


def load_pa(mics, n):
    _miso.load_pa(mics, n)


def steer(offset):
    _miso.steer(offset)


def stop_miso():
    pass


def _clamp_offset(offset, n):
    # steer_cartesian_degree(90, .) yields azimuth == MAX_RES_X: one row past the table
    # (SURVEY.md appendix A-16); clamp to the last row, keep the arithmetic otherwise.
    last = (config.MAX_RES_X * config.MAX_RES_Y - 1) * n
    return max(0, min(int(offset), last))


def steer_cartesian_degree(azimuth: float, elevation: float):
    """main.pyx:498-513"""
    assert -90 <= azimuth <= 90, "Invalid range"
    assert -90 <= elevation <= 90, "Invalid range"
    azimuth += 90
    azimuth /= 180
    azimuth = int(azimuth * config.MAX_RES_X)
    elevation += 90
    elevation /= 180
    elevation = int(elevation * config.MAX_RES_Y)
    _, n_active_mics = active_microphones()
    steer_offset = int(elevation * config.MAX_RES_X * n_active_mics + azimuth * n_active_mics)
    steer(_clamp_offset(steer_offset, n_active_mics))
    return steer_offset


def stear_miso_beam(azimuth: float, elevation: float):
    """main.pyx:515-528 (unit-square click -> table row)"""
    azimuth = int(azimuth * config.MAX_RES_X)
    elevation = int(elevation * config.MAX_RES_Y)
    _, n_active_mics = active_microphones()
    steer_offset = int(elevation * config.MAX_RES_X * n_active_mics + azimuth * n_active_mics)
    print(steer_offset)
    steer(_clamp_offset(steer_offset, n_active_mics))
    return steer_offset


# ---------------------------------------------------------------------------
# producer loops
# ---------------------------------------------------------------------------
def _loop_mimo_pad(q, running):
    """main.pyx:172-202"""
    power_framenr = 0
    _load_pad()
    mics, n = _active()
    power_map = _new_map()
    while running.value:
        try:
            _call_map("pad_mimo", power_map, mics, n)
            power_framenr += 1
            q.put((power_map, power_framenr))
        except Exception:  # noqa: BLE001
            break
    _native.lib().unload_coefficients_pad()


def api(q, running):
    """main.pyx:383-407"""
    _load_pad()
    mics, n = _active()
    mimo_arr = _new_map()
    while running.value:
        _call_map("pad_mimo", mimo_arr, mics, n)
        q.put(mimo_arr)
    _native.lib().unload_coefficients_pad()


def api_with_miso(q, running):
    """main.pyx:419-449"""
    _load_pad()
    mics, n = _active()
    mimo_arr = _new_map()
    load_miso()
    load_pa(mics, n)
    steer(0)
    steer_cartesian_degree(0, 0)
    while running.value:
        _call_map("pad_mimo", mimo_arr, mics, n)
        q.put(mimo_arr)
    stop_miso()
    _native.lib().unload_coefficients_pad()


def just_miso(q, running):
    """main.pyx:451-475"""
    _load_pad()
    mics, n = _active()
    load_miso()
    load_pa(mics, n)
    steer_cartesian_degree(0, 0)
    while running.value:
        time.sleep(0.1)
    stop_miso()
    _native.lib().unload_coefficients_pad()


def api_convolve(q, running):
    """main.pyx:477-495"""
    _native.configure_from(config)
    image = _new_map()
    h = _native.f32(compute_convolve_h())
    mics, n = _active()
    _native.lib().load_coefficients_convolve(_native.ptr(h), int(h.size))
    _native.check()
    while running.value:
        _call_map("convolve_mimo_vectorized", image, mics, n)
        q.put(image)
    _native.lib().unload_coefficients_convolve()


def api_miso(q, running):
    """main.pyx:531-549 (the reference reads an undefined global `steer_offset` here and
    would raise NameError; the offset of the module's MisoBeam is used instead)."""
    out = np.ascontiguousarray(np.zeros(config.N_SAMPLES, dtype=DTYPE_arr))
    _load_pad()
    mics, n = _active()
    while running.value:
        _native.lib().miso_steer_listen(_native.ptr(out), _native.ptr(mics), n, _miso.steer_offset)
        _native.check()
        q.put(out)


def _loop_miso(load, unload, q, running):
    load()
    mics, n = _active()
    load_miso()
    load_pa(mics, 64)                 # main.pyx:222 n_active_mics = 64
    steer_cartesian_degree(0, 0)
    while running.value:
        try:
            (x, y) = q.get()
            q.task_done()
            stear_miso_beam(x, y)
        except Exception as e:  # noqa: BLE001
            print(e)
    stop_miso()
    unload()


def _loop_miso_pad(q, running):
    """main.pyx:204-240"""
    _loop_miso(_load_pad, _native.lib().unload_coefficients_pad, q, running)


def _loop_miso_lerp(q, running):
    """main.pyx:243-277"""
    _loop_miso(_load_lerp, _native.lib().unload_coefficients_lerp, q, running)


def _loop_mimo_and_miso(load, unload, fn, miso_mics, q_steer, q_out, running):
    load()
    mics, n = _active()
    power_map = _new_map()
    load_miso()
    load_pa(mics, miso_mics if miso_mics else n)
    steer_cartesian_degree(0, 0)
    while running.value:
        try:
            _call_map(fn, power_map, mics, n)
            q_out.put(power_map)
            try:
                (x, y) = q_steer.get(block=False)
                q_steer.task_done()
                stear_miso_beam(x, y)
            except _queue.Empty:
                pass
            except Exception as e:  # noqa: BLE001
                print(e)
        except Exception:  # noqa: BLE001
            break
    stop_miso()
    unload()


def _loop_mimo_and_miso_pad(q_steer, q_out, running):
    """main.pyx:279-328"""
    _loop_mimo_and_miso(_load_pad, _native.lib().unload_coefficients_pad, "pad_mimo", 0,
                        q_steer, q_out, running)


def _loop_mimo_and_miso_lerp(q_steer, q_out, running):
    """main.pyx:330-380 (load_pa with 128 microphones, line 353)"""
    _loop_mimo_and_miso(_load_lerp, _native.lib().unload_coefficients_lerp, "lerp_mimo", 128,
                        q_steer, q_out, running)


# Web interface names (main.pyx:553-579, 809-820)
def uti_api(q, running):
    api(q, running)


def uti_api_with_miso(q, running):
    api_with_miso(q, running)


def conv_api(q, running):
    api_convolve(q, running)


def miso_api(q, running):
    api_miso(q, running)


def just_miso_api(q, running):
    just_miso(q, running)


def b(q, running):
    _loop_mimo_pad(q, running)


def just_miso_loop(q, running):
    while running.value:
        try:
            time.sleep(0.1)
        except KeyboardInterrupt:
            running.value = 0


def pure_miso_pad(q, running):
    _loop_miso_pad(q, running)


def pure_miso_lerp(q, running):
    _loop_miso_lerp(q, running)


def multi_pad(q_steer, q_out, running):
    _loop_mimo_and_miso_pad(q_steer, q_out, running)


def multi_lerp(q_steer, q_out, running):
    _loop_mimo_and_miso_lerp(q_steer, q_out, running)


def miso_listen(scaled=True):
    """One block of MISO audio from the module's beam (what the reference's audio child
    writes to its ring buffer each iteration, api.c:505-529)."""
    return _miso.listen(scaled)
