"""Batch replay of recordings into power-map video (BASELINE config C5).

A recording is the `.npy` layout written by the reference's PC/record.py:28-46 -- float32
(N_MICROPHONES, total_samples).  Video frame k looks at the N_SAMPLES window that starts at
sample floor(k * fs / fps); frames are independent (the beamformer zero-pads every block, no
history), so they shard over GPUs / ranks with no collective: rank r takes frames r, r+W, ....
"""
import numpy as np

from interface import config
from . import _native


def frame_starts(n_frames, fs=48828, fps=30, first=0):
    """floor(k*fs/fps) in integer arithmetic (hop = 1627.6 samples at 30 fps)."""
    k = np.arange(first, first + n_frames, dtype=np.int64)
    return (k * int(fs)) // int(fps)


def n_frames_in(total_samples, fs=48828, fps=30):
    """Number of complete N_SAMPLES windows available: frames k with floor(k*fs/fps) <= last start,
    i.e. k*fs < (last+1)*fps (same integer arithmetic as frame_starts)."""
    last = total_samples - config.N_SAMPLES
    return 0 if last < 0 else int(((last + 1) * int(fps) + int(fs) - 1) // int(fs))


def replay_dev(algo, d_recording, d_mic_ids, n, fps=30, fs=48828, chunk=64, rank=0, world=1, out=None):
    """Power maps of this rank's share of the frames of a device-resident recording.

    d_recording: torch CUDA float32 (N_MICROPHONES, samples).  Returns (frame_indices, maps) with
    maps a CUDA tensor (n_local_frames, D).  Tables must be loaded (load_coefficients_*)."""
    import torch
    L = _native.lib()
    M, N = config.N_MICROPHONES, config.N_SAMPLES
    D = config.MAX_RES_X * config.MAX_RES_Y
    total = n_frames_in(d_recording.shape[1], fs, fps)
    mine = np.arange(rank, total, world, dtype=np.int64)
    starts = torch.from_numpy((mine * int(fs)) // int(fps)).cuda()
    maps = out if out is not None else torch.empty((len(mine), D), dtype=torch.float32, device="cuda")
    frames = torch.empty((chunk, M, N), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for i in range(0, len(mine), chunk):
        c = min(chunk, len(mine) - i)
        _native.check(L.bf_window_dev(d_recording.data_ptr(), d_recording.shape[1], starts[i:].data_ptr(), c,
                                      frames.data_ptr(), stream))
        _native.check(L.bf_mimo_dev(algo, frames.data_ptr(), maps[i:].data_ptr(), c, d_mic_ids.data_ptr(), n,
                                    0, D, stream))
    return mine, maps


def video_dev(algo, d_recording, d_mic_ids, n, window=(640, 360), fps=30, fs=48828, chunk=64, rank=0, world=1,
              keep_heat=True, **heat_kw):
    """BASELINE config C5 end to end on the device: recording -> 30 fps power maps -> heat overlay
    (lib.visual.heatmaps_dev: colour map at `window` size, peak, entropy confidence) for this rank's share
    of the frames.  Returns dict(frames=indices, maps=[k][D], info=[k][48] u8, confidence=[k],
    heat=[k][H][W][3] u8 or None).  Nothing leaves the GPU; frames shard over ranks with no collective."""
    import torch
    from . import visual
    mine, maps = replay_dev(algo, d_recording, d_mic_ids, n, fps=fps, fs=fs, chunk=chunk, rank=rank, world=world)
    k = len(mine)
    W, H = window
    info = torch.empty((k, _native.HEAT_INFO_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    conf = torch.empty(k, dtype=torch.float64, device="cuda")
    heat = torch.empty((k if keep_heat else min(k, chunk), H, W, 3), dtype=torch.uint8, device="cuda")
    for i in range(0, k, chunk):
        c = min(chunk, k - i)
        res = visual.heatmaps_dev(maps[i:i + c], window=window, confidence=True, **heat_kw)
        info[i:i + c] = res["info"]
        conf[i:i + c] = res["confidence"]
        (heat[i:i + c] if keep_heat else heat[:c]).copy_(res["heat"])
    return {"frames": mine, "maps": maps, "info": info, "confidence": conf, "heat": heat if keep_heat else None}


def signals_from_blocks(blocks, n_arrays, quirk=True, zero_mask=None, norm=16777216.0, rows=8, cols=8):
    """int32 datagram payloads [k][N_SAMPLES][N_MICROPHONES] (lib.capture.blocks / iter_capture_blocks) ->
    device sample buffers float32 [k][N_MICROPHONES][N_SAMPLES]: the reference's receiver
    (receiver.c:94-151) on the device."""
    import torch
    L = _native.lib()
    M, N = config.N_MICROPHONES, config.N_SAMPLES
    if blocks.ndim != 3 or blocks.shape[1] != N or blocks.shape[2] != M:
        raise ValueError("payload blocks are %s, expected (k, N_SAMPLES=%d, N_MICROPHONES=%d)" % (blocks.shape, N, M))
    d_in = torch.from_numpy(np.ascontiguousarray(blocks, np.int32)).cuda()
    d_out = torch.zeros((blocks.shape[0], M, N), dtype=torch.float32, device="cuda")
    d_mask = torch.from_numpy(np.ascontiguousarray(zero_mask, np.uint8)).cuda() if zero_mask is not None else None
    _native.check(L.bf_ingest_dev(d_in.data_ptr(), d_out.data_ptr(), blocks.shape[0], int(n_arrays), rows, cols,
                                  float(norm), int(bool(quirk)), d_mask.data_ptr() if d_mask is not None else None,
                                  torch.cuda.current_stream().cuda_stream))
    return d_out


def signals_from_capture(cap, quirk=True, zero_mask=None, norm=16777216.0, rows=8, cols=8):
    """Packet capture (lib.capture.read_capture) -> device sample buffers: every whole block of N_SAMPLES
    datagrams becomes one float32 [N_MICROPHONES][N_SAMPLES] frame.  Returns a torch CUDA tensor
    [blocks][N_MICROPHONES][N_SAMPLES]."""
    from . import capture
    if cap.stream.shape[1] != config.N_MICROPHONES:
        raise ValueError("capture has %d channels per datagram, config.N_MICROPHONES is %d"
                         % (cap.stream.shape[1], config.N_MICROPHONES))
    return signals_from_blocks(capture.blocks(cap.stream, config.N_SAMPLES), cap.n_arrays, quirk, zero_mask, norm,
                               rows, cols)


def stream_video(h_stream, n_arrays, algo, d_mic_ids, n, fps=30, fs=48828, chunk_frames=256, window=(640, 360),
                 passes=1, quirk=True, overlay=True, first_frame=0, frame_step=1, keep_maps=False):
    """BASELINE config C5 as a STREAM: a stored recording in the wire format -- h_stream, pinned host int32
    [datagrams][N_MICROPHONES], what the Zybo sent (receiver.h:51-59 payloads; lib.capture gives exactly this) --
    is pushed through the GPU in chunks: pinned double-buffered H2D copy (copy stream) -> windowed wire-format
    conversion (bf_ingest_windows_dev: frame k = the N_SAMPLES datagrams from floor(k*fs/fps)) -> power maps
    (bf_mimo_dev) -> sensor-fusion overlay at `window` size + peak + entropy confidence (lib.visual), with the
    H2D copy of chunk c+1 overlapping the compute of chunk c.  Nothing of the recording is resident beyond two
    chunks.  `passes` replays the host buffer that many times (a recording longer than host memory allows:
    the bytes crossing PCIe are real, the content repeats).  Frames first_frame, first_frame + frame_step, ...
    are this caller's share (recordings / frames shard over ranks with no collective).

    Returns dict(frames, seconds (wall), frames_per_s, h2d_bytes, h2d_gb_per_s, stage_ms = device time per
    stage summed over chunks (h2d on the copy stream, the others on the compute stream), info, confidence,
    and with keep_maps the [frames][D] power maps of the last pass as a CUDA tensor)."""
    import time
    import torch
    from . import visual
    L = _native.lib()
    M, N = config.N_MICROPHONES, config.N_SAMPLES
    D = config.MAX_RES_X * config.MAX_RES_Y
    if h_stream.dtype != torch.int32 or h_stream.ndim != 2 or h_stream.shape[1] != M or not h_stream.is_pinned():
        raise ValueError("h_stream must be a pinned torch.int32 tensor [datagrams][N_MICROPHONES=%d]" % M)
    total = int(h_stream.shape[0])
    n_frames = n_frames_in(total, fs, fps)
    mine = np.arange(first_frame, n_frames, frame_step, dtype=np.int64)
    starts_all = (mine * int(fs)) // int(fps)
    chunks = [(i, min(i + chunk_frames, len(mine))) for i in range(0, len(mine), chunk_frames)]
    span = max(int(starts_all[b - 1] + N - starts_all[a]) for a, b in chunks) if chunks else 0
    d_slot = [torch.empty((span, M), dtype=torch.int32, device="cuda") for _ in range(2)]
    d_frames = torch.empty((chunk_frames, M, N), dtype=torch.float32, device="cuda")
    d_maps = torch.empty((chunk_frames, D), dtype=torch.float32, device="cuda")
    k_total = len(mine) * passes
    info = torch.empty((k_total, _native.HEAT_INFO_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    conf = torch.empty(k_total, dtype=torch.float64, device="cuda")
    heat_out = None
    all_maps = torch.empty((len(mine), D), dtype=torch.float32, device="cuda") if keep_maps else None
    # window starts relative to the first datagram of their chunk, for all chunks, in one upload
    local = starts_all.copy()
    for a, b in chunks:
        local[a:b] -= starts_all[a]
    d_local = torch.from_numpy(local).cuda()
    s_copy, s_run = torch.cuda.Stream(), torch.cuda.current_stream()
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    ev_used = [torch.cuda.Event() for _ in range(2)]
    timing = {k: [] for k in ("h2d", "ingest", "maps", "overlay")}
    mk = lambda: torch.cuda.Event(enable_timing=True)
    work = [(p, a, b) for p in range(passes) for a, b in chunks]
    h2d_bytes = 0

    def enqueue_copy(j):
        p, a, b = work[j]
        sl = j & 1
        base = int(starts_all[a])
        cnt = min(int(starts_all[b - 1] + N - base), total - base)
        with torch.cuda.stream(s_copy):
            if j >= 2:
                s_copy.wait_event(ev_used[sl])                   # the compute that read this slot has finished
            e0, e1 = mk(), mk()
            e0.record(s_copy)
            d_slot[sl][:cnt].copy_(h_stream[base:base + cnt], non_blocking=True)
            e1.record(s_copy)
            ev_copied[sl].record(s_copy)
        timing["h2d"].append((e0, e1))
        return cnt * M * 4

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if work:
        h2d_bytes += enqueue_copy(0)
    done = 0
    for j, (p, a, b) in enumerate(work):
        sl = j & 1
        if j + 1 < len(work):
            h2d_bytes += enqueue_copy(j + 1)                     # overlaps this chunk's compute
        c = b - a
        base = int(starts_all[a])
        cnt = min(int(starts_all[b - 1] + N - base), total - base)
        d_starts = d_local[a:b]
        s_run.wait_event(ev_copied[sl])
        st = s_run.cuda_stream
        e = [mk() for _ in range(4)]
        e[0].record(s_run)
        _native.check(L.bf_ingest_windows_dev(d_slot[sl].data_ptr(), cnt, d_starts.data_ptr(), d_frames.data_ptr(), c,
                                              int(n_arrays), 8, 8, 16777216.0, int(bool(quirk)), None, st))
        ev_used[sl].record(s_run)
        e[1].record(s_run)
        _native.check(L.bf_mimo_dev(algo, d_frames.data_ptr(), d_maps.data_ptr(), c, d_mic_ids.data_ptr(), n, 0, D, st))
        e[2].record(s_run)
        if keep_maps:
            all_maps[a:b] = d_maps[:c]
        if overlay:
            res = visual.heatmaps_dev(d_maps[:c], window=window, confidence=True,
                                      out=heat_out if (heat_out is not None and c == chunk_frames) else None)
            if c == chunk_frames:
                heat_out = res
            info[done:done + c] = res["info"]
            conf[done:done + c] = res["confidence"]
        e[3].record(s_run)
        timing["ingest"].append((e[0], e[1]))
        timing["maps"].append((e[1], e[2]))
        timing["overlay"].append((e[2], e[3]))
        done += c
    h_info, h_conf = info.cpu(), conf.cpu()                      # the per-frame results reach the host
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    stage_ms = {k: float(sum(a.elapsed_time(b) for a, b in v)) for k, v in timing.items()}
    return {"frames": done, "seconds": dt, "frames_per_s": done / dt if dt > 0 else 0.0, "h2d_bytes": h2d_bytes,
            "h2d_gb_per_s": h2d_bytes / dt / 1e9 if dt > 0 else 0.0, "stage_ms": stage_ms,
            "info": h_info, "confidence": h_conf, "recording_frames": len(mine), "passes": passes, "maps": all_maps}
