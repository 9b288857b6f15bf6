"""Seeded synthetic microphone-array buffers.

Used by bench.py, the tests and the golden-vector generator so that the CPU
reference and the CUDA path always see identical bytes.  NumPy only.
"""
import numpy as np


def plot_py_stimulus(n_mics, n_samples, fs=48828, frequency=8000):
    """The reference's only reproducible stimulus (PC/plot.py:8-25): a unit
    sine sampled at config.fs, identical on every microphone, float32."""
    time = np.arange(0, 1, 1 / fs)
    wave = (1 * np.sin(2 * np.pi * frequency * time + 0))[:n_samples]
    sig = np.repeat(wave, n_mics, axis=0).reshape((n_samples, n_mics)).T
    return np.ascontiguousarray(sig, dtype=np.float32)   # (PC/plot.py returns the transposed view)


def point_sources(delays, mic_ids, n_mics_total, n_samples, fs, sources, noise_sigma, seed,
                  t0=0):
    """Far-field point sources seen through a delay table.

    delays   float64 [D][n] (samples, >= 0) -- row d is the steering delay set
    mic_ids  int[n]   microphone id of table column m
    sources  iterable of (direction_index, frequency_hz, amplitude)
    Microphone m receives  A*sin(2*pi*f*(t - tau_m)/fs),  tau_m = max_m(delay[d0]) -
    delay[d0, m]  (fractional), so steering to d0 re-aligns the channels.
    White noise N(0, noise_sigma^2) is added on every channel.  Returns float32
    [n_mics_total][n_samples].
    """
    delays = np.asarray(delays, dtype=np.float64).reshape(-1, len(mic_ids))
    rng = np.random.default_rng(seed)
    t = np.arange(t0, t0 + n_samples, dtype=np.float64)
    sig = np.zeros((n_mics_total, n_samples), dtype=np.float64)
    for d0, f0, amp in sources:
        tau = delays[d0].max() - delays[d0]
        for m, mic in enumerate(mic_ids):
            sig[mic] += amp * np.sin(2.0 * np.pi * f0 * (t - tau[m]) / fs)
    sig += rng.normal(0.0, noise_sigma, size=sig.shape)
    return sig.astype(np.float32)


# The BASELINE.json workloads, as data.
C1 = dict(name="c1", grid=(20, 20), n_mics_total=64, sources=[(14 * 20 + 6, 4000.0, 0.1)],
          noise=0.01, seed=1234)
C3 = dict(name="c3", grid=(180, 180), n_mics_total=256,
          sources=[(40 * 180 + 100, 2000.0, 0.1), (90 * 180 + 30, 5000.0, 0.07),
                   (150 * 180 + 150, 9000.0, 0.05)],
          noise=0.01, seed=1236)
