"""Capture-file readers for batch replay (SURVEY.md 8f "next" #3).

The reference records a session as (PC/src/main.pyx:741-806)
  * a packet capture of the Zybo's UDP stream written by tshark (`record_udp`, pcapng by default,
    classic pcap with `-F pcap`) plus `udp_timestamps.csv` (packet_number, timestamp),
  * the webcam video plus `video_timestamps.csv` (frame_number, timestamp),
and `.npy` sample buffers (PC/record.py:28-46, handled by lib/replay.py).
One datagram = one sample instant (receiver.h:51-59): u16 frequency, i8 n_arrays, i8 protocol_ver,
i32 counter, i32 stream[N_MICROPHONES], little endian.  This module turns a capture into the
int32 `[packets][N_MICROPHONES]` payload matrix that `bf_ingest_dev` (csrc/ingest.cu) converts on
the device, and aligns video frames to packet indices.  Host-side I/O only: no arithmetic on the
samples happens here.
"""
import struct
from collections import namedtuple

import numpy as np

Capture = namedtuple("Capture", "stream counter timestamps frequency n_arrays protocol_version dropped")

_HDR = struct.Struct("<Hbbi")


def _udp_payload(frame, linktype):
    """UDP payload of an Ethernet (1) / raw-IP (101) / Linux cooked (113) frame, or None."""
    if linktype == 1:
        off = 14
        et = frame[12:14]
        if et == b"\x81\x00":            # 802.1Q
            et, off = frame[16:18], 18
        if et != b"\x08\x00":
            return None
    elif linktype == 113:
        if frame[14:16] != b"\x08\x00":
            return None
        off = 16
    elif linktype == 101:
        off = 0
    else:
        return None
    if len(frame) < off + 28 or frame[off] >> 4 != 4 or frame[off + 9] != 17:
        return None
    ihl = (frame[off] & 15) * 4
    ulen = struct.unpack(">H", frame[off + ihl + 4:off + ihl + 6])[0]
    return frame[off + ihl + 8:off + ihl + ulen]


def _iter_pcap(buf):
    magic = struct.unpack("<I", buf[:4])[0]
    if magic in (0xA1B2C3D4, 0xA1B23C4D):
        end, nano = "<", magic == 0xA1B23C4D
    elif magic in (0xD4C3B2A1, 0x4D3CB2A1):
        end, nano = ">", magic == 0x4D3CB2A1
    else:
        raise ValueError("not a pcap file")
    linktype = struct.unpack(end + "I", buf[20:24])[0]
    pos, div = 24, 1e9 if nano else 1e6
    while pos + 16 <= len(buf):
        sec, frac, incl, _ = struct.unpack(end + "IIII", buf[pos:pos + 16])
        yield sec + frac / div, buf[pos + 16:pos + 16 + incl], linktype
        pos += 16 + incl


def _iter_pcapng(buf):
    pos, end = 0, "<"
    links, res = [], []
    while pos + 12 <= len(buf):
        btype = struct.unpack(end + "I", buf[pos:pos + 4])[0]
        if btype == 0x0A0D0D0A:
            end = "<" if struct.unpack("<I", buf[pos + 8:pos + 12])[0] == 0x1A2B3C4D else ">"
            links, res = [], []
        total = struct.unpack(end + "I", buf[pos + 4:pos + 8])[0]
        body = buf[pos + 8:pos + total - 4]
        if btype == 1:                                               # interface description
            links.append(struct.unpack(end + "H", body[:2])[0])
            tsres, o = 1e-6, 8
            while o + 4 <= len(body):
                code, ln = struct.unpack(end + "HH", body[o:o + 4])
                if code == 0:
                    break
                if code == 9 and ln >= 1:                            # if_tsresol
                    v = body[o + 4]
                    tsres = 2.0 ** -(v & 0x7F) if v & 0x80 else 10.0 ** -v
                o += 4 + ln + ((-ln) % 4)
            res.append(tsres)
        elif btype == 6:                                             # enhanced packet
            iface, hi, lo, cap_len, _ = struct.unpack(end + "IIIII", body[:20])
            yield ((hi << 32) | lo) * res[iface], body[20:20 + cap_len], links[iface]
        elif btype == 3 and links:                                   # simple packet (no timestamp)
            yield 0.0, body[4:], links[0]
        pos += total if total >= 12 else len(buf)


def read_capture(path, n_microphones=256):
    """-> Capture(stream int32 [packets][n_microphones], counter, timestamps, header fields, dropped).

    Only UDP datagrams of exactly 8 + 4*n_microphones bytes are kept (other traffic is skipped);
    `dropped` is the number of sample instants missing according to the 32-bit counter."""
    with open(path, "rb") as f:
        buf = f.read()
    it = _iter_pcapng(buf) if buf[:4] == b"\x0a\x0d\x0d\x0a" else _iter_pcap(buf)
    size = 8 + 4 * n_microphones
    rows, counters, stamps, hdr, other = [], [], [], None, 0
    for ts, frame, link in it:
        pl = _udp_payload(frame, link)
        if pl is None:
            continue
        if len(pl) != size:
            other += len(pl) >= 8
            continue
        freq, n_arrays, ver, counter = _HDR.unpack_from(pl)
        hdr = hdr or (freq, n_arrays, ver)
        rows.append(np.frombuffer(pl, "<i4", n_microphones, 8))
        counters.append(counter)
        stamps.append(ts)
    if not rows:
        raise ValueError("%s: no %d-byte datagrams (N_MICROPHONES = %d); %d UDP datagrams of other sizes"
                         % (path, size, n_microphones, other))
    counter = np.array(counters, np.int64)
    gaps = np.diff(counter) & 0xFFFFFFFF
    # forward jumps of 2..2^31 are losses; a gap of 0 (duplicate) or >= 2^31 (a datagram that arrived
    # late, i.e. a backward step modulo 2^32) is reordering, not loss
    lost = gaps[(gaps > 1) & (gaps < (1 << 31))]
    dropped = int(np.sum(lost - 1))
    return Capture(np.ascontiguousarray(np.stack(rows)), counter, np.array(stamps), hdr[0], hdr[1], hdr[2], dropped)


def iter_capture_blocks(path, n_microphones=256, n_samples=256, chunk_blocks=64):
    """Stream a capture that does not fit in memory (1 h of the 256-channel stream is ~180 GB): memory-map the
    file and yield (stream int32 [k][n_samples][n_microphones], counter int64 [k*n_samples], timestamps) for
    successive chunks of up to `chunk_blocks` whole blocks; the tail that does not fill a block is dropped,
    like the reference's receiver drops a partial buffer.  Each chunk is what replay.signals_from_capture /
    bf_ingest_dev take."""
    import mmap
    size = 8 + 4 * n_microphones
    with open(path, "rb") as f, mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ) as buf:
        it = _iter_pcapng(buf) if buf[:4] == b"\x0a\x0d\x0d\x0a" else _iter_pcap(buf)
        want = chunk_blocks * n_samples
        rows = np.empty((want, n_microphones), np.int32)
        counters = np.empty(want, np.int64)
        stamps = np.empty(want, np.float64)
        fill = 0
        for ts, frame, link in it:
            pl = _udp_payload(frame, link)
            if pl is None or len(pl) != size:
                continue
            rows[fill] = np.frombuffer(pl, "<i4", n_microphones, 8)
            counters[fill] = _HDR.unpack_from(pl)[3]
            stamps[fill] = ts
            fill += 1
            if fill == want:
                yield rows.reshape(chunk_blocks, n_samples, n_microphones).copy(), counters.copy(), stamps.copy()
                fill = 0
        k = fill // n_samples
        if k:
            yield (rows[:k * n_samples].reshape(k, n_samples, n_microphones).copy(), counters[:k * n_samples].copy(),
                   stamps[:k * n_samples].copy())


def read_timestamps(csv_path):
    """Second column of the reference's timestamp CSVs (header row, then index,timestamp)."""
    return np.atleast_1d(np.loadtxt(csv_path, delimiter=",", skiprows=1, usecols=1, dtype=np.float64))


def align_frames(video_ts, packet_ts, n_samples=256):
    """First packet index of the N_SAMPLES window shown with each video frame: the first packet at or
    after the frame's timestamp, clipped so that the window stays inside the capture."""
    idx = np.searchsorted(packet_ts, video_ts, side="left")
    return np.minimum(idx, max(0, len(packet_ts) - n_samples)).astype(np.int64)


def blocks(stream, n_samples=256):
    """Whole consecutive blocks [k][n_samples][mics] as the ingest kernel takes them."""
    k = stream.shape[0] // n_samples
    return stream[:k * n_samples].reshape(k, n_samples, stream.shape[1])
