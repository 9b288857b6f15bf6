"""Heat-map post-processing on the B200 -- host mirror of the reference's `lib.visual`
(PC/src/visual.py, compiled to lib/visual by the reference's build and imported as
`from lib.visual import calculate_heatmap, calculate_heatmap_fft`, PC/application/camera.py:6).

Same names, arguments and return values; the work runs in csrc/heatmap.cu (no CPU path: the
calls fail loudly without the CUDA library / a GPU).  The window drawing, webcam and Kalman
parts of the reference module are UI and out of scope.

    calculate_heatmap(image)                -> (heatmap uint8 [H][W][3], should_overlay)   visual.py:130-171
    calculate_heatmap_fft(image)            -> same, linear scale                           visual.py:173-205
    calculate_heatmap_with_detection(image) -> (box, heatmap, should_overlay)               visual.py:227-291
    find_power_center(image)                -> (center_x, center_y)                         visual.py:293-322
    generate_color_map()                    -> uint8 [256][3]                               visual.py:27-48
    heatmaps_dev(maps)                      -> batched, device-resident form (torch CUDA tensors)
"""
import ctypes

import numpy as np

from interface import config
from . import _native

WINDOW_DIMENSIONS = (1920, 1080)      # visual.py:9 (overrides the config values, as there)
POWER = 5                             # visual.py:13


def generate_color_map(name="jet"):
    if name != "jet":
        raise ValueError("only the reference's default colour map (jet) is built in; pass a table via lut=")
    lut = np.zeros((256, 3), np.uint8)
    _native.check(_native.lib().bf_jet_lut(_native.ptr(lut)))
    return lut


colors = None


def _lut(lut):
    global colors
    if lut is not None:
        lut = np.ascontiguousarray(lut, np.uint8)
        assert lut.shape == (256, 3)
        return lut
    if colors is None:
        colors = generate_color_map()
    return colors


def _run(images, threshold, amount, exponent, log_scale, window, lut, confidence=False):
    """images float32 [frames][X][Y] (host) -> (heat [frames][H][W][3], info records, confidence)."""
    images = np.ascontiguousarray(images, np.float32)
    frames, X, Y = images.shape
    W, H = window
    heat = np.empty((frames, H, W, 3), np.uint8)
    info = np.zeros(frames, _native.HEAT_INFO_DTYPE)
    conf = np.zeros(frames, np.float64) if confidence else None
    _native.check(_native.lib().bf_heatmap(
        _native.ptr(images), frames, X, Y, float(threshold), float(amount), int(exponent), int(log_scale),
        _native.ptr(_lut(lut)), int(W), int(H), _native.ptr(heat), _native.ptr(info),
        _native.ptr(conf) if confidence else None))
    return heat, info, conf


def _plane(image):
    image = np.asarray(image)
    if image.ndim == 3:
        image = image[..., 0]             # visual.py:145-146
    return image


def calculate_heatmap(image, threshold=1e-7, amount=0.5, exponent=POWER, lut=None, window=None):
    heat, info, _ = _run(_plane(image)[None], threshold, amount, exponent, 1, window or WINDOW_DIMENSIONS, lut)
    return heat[0], bool(info["overlay"][0])


def calculate_heatmap_fft(image, threshold=5e-8, lut=None, window=None):
    """visual.py:173-205: normalises `image` in place like the reference, gate max > threshold*1e6."""
    image_in = _plane(image)
    heat, info, _ = _run(image_in[None], threshold * 1000000, 0.5, 2, 0, window or WINDOW_DIMENSIONS, lut)
    if isinstance(image, np.ndarray) and image.ndim == 2 and image.dtype == np.float32:
        image /= np.max(image)            # the reference's visible side effect (visual.py:180)
    return heat[0], bool(info["overlay"][0])


def find_power_center(image, region_size=3):
    image = np.ascontiguousarray(_plane(image), np.float32)
    X, Y = image.shape
    _, info, _ = _run(image[None], 0.0, 0.5, POWER, 1, (X, Y), None)
    return float(info["center_col"][0]), float(info["center_row"][0])


def detection_box(peak_x, peak_y, window=None, box_size_ratio=0.1):
    """visual.py:268-281: peak (grid units) -> (centre_x, centre_y, x1, y1, x2, y2) in window pixels."""
    win = window or WINDOW_DIMENSIONS
    cx = win[0] - 1 - int(peak_x / (config.MAX_RES_X - 1) * win[0])
    cy = win[1] - 1 - int(peak_y / (config.MAX_RES_Y - 1) * win[1])
    bw, bh = int(win[0] * box_size_ratio), int(win[1] * box_size_ratio)
    return (cx, cy, max(0, cx - bw // 2), max(0, cy - bh // 2), min(win[0], cx + bw // 2), min(win[1], cy + bh // 2))


def calculate_heatmap_with_detection(image, threshold=1e-7, amount=0.5, exponent=POWER, box_size_ratio=0.1,
                                     region_size=3, lut=None, window=None):
    """Returns (box, heatmap, should_overlay).  `box` is the (centre_x, centre_y, x1, y1, x2, y2) tuple the
    reference draws into its `power_detection` image (visual.py:283-284), or None when nothing is overlaid;
    the drawing itself (cv2.rectangle / cv2.circle) is left to the UI."""
    win = window or WINDOW_DIMENSIONS
    heat, info, _ = _run(_plane(image)[None], threshold, amount, exponent, 1, win, lut)
    overlay = bool(info["overlay"][0])
    box = None
    if overlay:
        # visual.py:244: peak_y, peak_x = find_power_center(...)
        peak_y, peak_x = float(info["center_col"][0]), float(info["center_row"][0])
        box = detection_box(peak_x, peak_y, win, box_size_ratio)
    return box, heat[0], overlay


def heatmaps_dev(d_maps, window=None, threshold=1e-7, amount=0.5, exponent=POWER, log_scale=True, lut=None,
                 confidence=False, out=None):
    """Batched device-resident form: d_maps torch CUDA float32 [frames][X*Y] or [frames][X][Y]
    -> dict(heat=uint8 [frames][H][W][3], small=uint8 [frames][Y][X][3], index=int16 [frames][X][Y],
            info=uint8 [frames][48] (view with _native.HEAT_INFO_DTYPE after .cpu()), confidence=float64 [frames])."""
    import torch
    L = _native.lib()
    frames = d_maps.shape[0]
    X, Y = config.MAX_RES_X, config.MAX_RES_Y
    assert d_maps.is_cuda and d_maps.dtype == torch.float32 and d_maps[0].numel() == X * Y and d_maps.is_contiguous()
    W, H = window or WINDOW_DIMENSIONS
    dev = d_maps.device
    res = out or {}
    if "small" not in res:
        res["small"] = torch.empty((frames, Y, X, 3), dtype=torch.uint8, device=dev)
        res["index"] = torch.empty((frames, X, Y), dtype=torch.int16, device=dev)
        res["info"] = torch.empty((frames, _native.HEAT_INFO_DTYPE.itemsize), dtype=torch.uint8, device=dev)
        res["heat"] = torch.empty((frames, H, W, 3), dtype=torch.uint8, device=dev)
        if confidence:
            res["confidence"] = torch.empty(frames, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _native.check(L.bf_heatmap_dev(d_maps.data_ptr(), frames, X * Y, X, Y, float(threshold), float(amount),
                                   int(exponent), int(bool(log_scale)), _native.ptr(_lut(lut)),
                                   res["small"].data_ptr(), res["index"].data_ptr(), res["info"].data_ptr(), st))
    _native.check(L.bf_resize_linear_u8_dev(res["small"].data_ptr(), frames, Y, X, 3, res["heat"].data_ptr(), H, W, st))
    if confidence:
        _native.check(L.bf_entropy_dev(res["heat"].data_ptr(), frames, H * W * 3, res["confidence"].data_ptr(), st))
    return res
