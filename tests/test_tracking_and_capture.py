"""CPU (host-only code, no GPU call): the peak-tracking filter against its oracle and the
capture-file readers of the replay path."""
import struct

import numpy as np
import pytest

from oracle import kf_np


def test_kalman_filter_matches_oracle():
    from lib.kf import CyKF
    rng = np.random.default_rng(11)
    kf, ref = CyKF(), kf_np.KalmanFilter3D()
    assert np.array_equal(kf.get_state(), np.zeros(3, np.float32))
    pos = np.array([10.0, 5.0, 0.0])
    for k in range(200):
        pos = pos + np.array([0.3, -0.1, 0.0]) + (5.0 if k == 120 else 0.0)      # a jump at k = 120
        m = pos + rng.normal(0, 0.5, 3)
        kf.update(list(m))
        ref.update(m.astype(np.float32))
        assert np.allclose(kf.get_state(), ref.get_state(), rtol=1e-5, atol=1e-4), k
        if k % 37 == 0:
            for n in (0, 1, 2, 3):
                assert np.allclose(kf.predict(n), ref.predict(n), rtol=1e-4, atol=1e-3), (k, n)
    assert kf.get_state().dtype == np.float32
    # the known-answer sequence of the reference's own debug main (kf.hpp:170-176)
    kf = CyKF()
    for m in [(1, 1, 0), (2, 2, 0), (3, 4, 0)]:
        kf.update(m)
    assert np.allclose(kf.get_state(), [2.9707165, 3.8292835, 0.0], atol=1e-5)


def test_kalman_filter_matches_the_reference_class():
    """The product's filter (csrc/kf_host.cu behind lib.kf.CyKF) against the REFERENCE's own KalmanFilter3D --
    PC/src/kf.hpp compiled unmodified (oracle/build_ref.py:build_kf) against a minimal stand-in for Eigen, which
    this image lacks (oracle/eigen_shim: plain-loop fixed-size matrices).  Pinned: the reference's matrices, its
    update / predict equations and the quirk of predict() (the transition matrix is re-multiplied every step).
    Not pinned: Eigen's own evaluation order inside a 6x6 product, hence a float32 tolerance, not bit equality."""
    from oracle import ref
    if not ref.RefKalman.available():
        pytest.skip("oracle/_ref/default/libkf_ref.so not built")
    from lib.kf import CyKF
    rng = np.random.default_rng(12)
    kf, rk = CyKF(), ref.RefKalman()
    assert np.array_equal(rk.get_state(), np.zeros(3, np.float32))
    pos = np.array([-3.0, 8.0, 1.0])
    worst = 0.0
    for k in range(300):
        pos = pos + np.array([0.2, 0.05, -0.01]) + (4.0 if k == 150 else 0.0)
        m = (pos + rng.normal(0, 0.4, 3)).astype(np.float32)
        kf.update(list(m))
        rk.update(m)
        a, b = kf.get_state(), rk.get_state()
        worst = max(worst, float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0))))
        if k % 29 == 0:
            for n in (0, 1, 2, 3, 5):
                assert np.allclose(kf.predict(n), rk.predict(n), rtol=2e-5, atol=2e-5), (k, n)
    assert worst <= 2e-6, worst
    # the reference's own debug sequence (kf.hpp:170-176) through the reference's own class
    rk = ref.RefKalman()
    for m in [(1, 1, 0), (2, 2, 0), (3, 4, 0)]:
        rk.update(m)
    assert np.allclose(rk.get_state(), [2.9707165, 3.8292835, 0.0], atol=1e-5)


# ---- capture files -------------------------------------------------------------------------------
def _udp_frame(payload, src_port=21844, dst_port=21844):
    udp = struct.pack(">HHHH", src_port, dst_port, 8 + len(payload), 0) + payload
    ip = struct.pack(">BBHHHBBH4s4s", 0x45, 0, 20 + len(udp), 0, 0, 64, 17, 0, bytes([10, 0, 0, 1]), bytes([10, 0, 0, 2]))
    return b"\x02" * 6 + b"\x04" * 6 + b"\x08\x00" + ip + udp


def _datagram(counter, stream, frequency=48828, n_arrays=4, version=2):
    return struct.pack("<Hbbi", frequency, n_arrays, version, counter) + stream.astype("<i4").tobytes()


def _write_pcap(path, frames, ts):
    with open(path, "wb") as f:
        f.write(struct.pack("<IHHiIII", 0xA1B2C3D4, 2, 4, 0, 0, 65535, 1))
        for fr, t in zip(frames, ts):
            f.write(struct.pack("<IIII", int(t), int(round((t - int(t)) * 1e6)), len(fr), len(fr)) + fr)


def _block(btype, body):
    pad = (-len(body)) % 4
    total = 12 + len(body) + pad
    return struct.pack("<II", btype, total) + body + b"\0" * pad + struct.pack("<I", total)


def _write_pcapng(path, frames, ts):
    with open(path, "wb") as f:
        f.write(_block(0x0A0D0D0A, struct.pack("<IHHq", 0x1A2B3C4D, 1, 0, -1)))
        f.write(_block(1, struct.pack("<HHI", 1, 0, 65535)))                    # Ethernet, default usec resolution
        for fr, t in zip(frames, ts):
            us = int(round(t * 1e6))
            f.write(_block(6, struct.pack("<IIIII", 0, us >> 32, us & 0xFFFFFFFF, len(fr), len(fr)) + fr))
        f.write(_block(5, struct.pack("<IQ", 0, 0)))                             # a block type to skip


@pytest.mark.parametrize("fmt", ["pcap", "pcapng"])
def test_capture_reader(tmp_path, fmt):
    from lib import capture
    rng = np.random.default_rng(5)
    M, P = 256, 700
    streams = rng.integers(-2 ** 23, 2 ** 23, (P, M), dtype=np.int32)
    ts = 1700000000.25 + np.arange(P) / 48828.0
    frames = [_udp_frame(_datagram(1000 + i, streams[i])) for i in range(P)]
    frames.insert(5, _udp_frame(b"short"))                                    # foreign UDP traffic: skipped
    ts_all = np.insert(ts, 5, ts[5])
    path = str(tmp_path / ("cap." + fmt))
    (_write_pcap if fmt == "pcap" else _write_pcapng)(path, frames, ts_all)
    cap = capture.read_capture(path, n_microphones=M)
    assert cap.stream.shape == (P, M) and cap.stream.dtype == np.int32
    assert np.array_equal(cap.stream, streams)
    assert np.array_equal(cap.counter, 1000 + np.arange(P))
    assert cap.frequency == 48828 and cap.n_arrays == 4 and cap.protocol_version == 2
    assert np.allclose(cap.timestamps, ts, atol=2e-6)
    assert cap.dropped == 0
    # timestamp CSVs (main.pyx:753-767, 789-794) and frame alignment
    csv_udp, csv_vid = str(tmp_path / "udp.csv"), str(tmp_path / "video.csv")
    with open(csv_udp, "w") as f:
        f.write("packet_number,timestamp\n" + "".join("%d,%.6f\n" % (i, t) for i, t in enumerate(ts)))
    vts = ts[0] + np.array([0.0, 1 / 30, 2 / 30, 0.0123])
    with open(csv_vid, "w") as f:
        f.write("frame_number,timestamp\n" + "".join("%d,%.6f\n" % (i, t) for i, t in enumerate(vts)))
    assert np.allclose(capture.read_timestamps(csv_udp), ts, atol=1e-6)
    starts = capture.align_frames(capture.read_timestamps(csv_vid), cap.timestamps, n_samples=256)
    want = np.searchsorted(ts, capture.read_timestamps(csv_vid), side="left")
    assert np.array_equal(starts, np.minimum(want, P - 256))
    blocks = capture.blocks(cap.stream, 256)
    assert blocks.shape == (2, 256, M) and np.array_equal(blocks[1], streams[256:512])


@pytest.mark.parametrize("fmt", ["pcap", "pcapng"])
def test_capture_streaming_iterator(tmp_path, fmt):
    from lib import capture
    rng = np.random.default_rng(8)
    M, N, P = 64, 32, 32 * 7 + 5                                   # 7 whole blocks + a partial one
    streams = rng.integers(-1000, 1000, (P, M), dtype=np.int32)
    frames = [_udp_frame(_datagram(i, streams[i], n_arrays=1)) for i in range(P)]
    frames.insert(40, _udp_frame(b"x" * 100))
    ts = 5.0 + np.arange(len(frames)) * 1e-4
    path = str(tmp_path / ("s." + fmt))
    (_write_pcap if fmt == "pcap" else _write_pcapng)(path, frames, ts)
    chunks = list(capture.iter_capture_blocks(path, n_microphones=M, n_samples=N, chunk_blocks=3))
    assert [c[0].shape[0] for c in chunks] == [3, 3, 1]
    got = np.concatenate([c[0].reshape(-1, M) for c in chunks])
    assert np.array_equal(got, streams[:7 * N])
    assert np.array_equal(np.concatenate([c[1] for c in chunks]), np.arange(7 * N))
    whole = capture.read_capture(path, n_microphones=M)
    assert np.array_equal(capture.blocks(whole.stream, N), got.reshape(7, N, M))


def test_capture_detects_drops(tmp_path):
    from lib import capture
    M = 64
    streams = np.arange(10 * M, dtype=np.int32).reshape(10, M)
    counters = [0, 1, 2, 4, 5, 9, 10, 11, 12, 13]
    frames = [_udp_frame(_datagram(c, s, n_arrays=1)) for c, s in zip(counters, streams)]
    path = str(tmp_path / "d.pcap")
    _write_pcap(path, frames, np.arange(10) * 1e-3)
    cap = capture.read_capture(path, n_microphones=M)
    assert cap.dropped == 4 and cap.n_arrays == 1
    with pytest.raises(ValueError):
        capture.read_capture(path, n_microphones=256)           # payload size does not match
