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
