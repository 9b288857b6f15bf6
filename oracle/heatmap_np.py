"""NumPy restatement of the reference's heat-map post-processing.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header): tests/, smoke() and the
CPU legs of bench.py may import this, the product path may not.

Follows /root/reference/PC/src/visual.py:
  * generate_color_map            27-48   (Matplotlib's "jet", reversed, truncated to u8)
  * calculate_heatmap             130-171 (clip, log10, -log10(min), /max, threshold, **exponent, LUT, flip)
  * calculate_heatmap_fft         173-205 (linear variant, fixed 0.5 / **2)
  * calculate_heatmap_with_detection 227-291 (adds the peak + box coordinates)
  * find_power_center             293-322 (5x5 Gaussian sigma 1, >= 95 % mask, cube-weighted centroid)
and /root/reference/PC/sensorfusion/decider.py:16-24 (get_entropy), 70-81 (focus_beam).

Third-party arithmetic restated here (the reference calls it through cv2 / Matplotlib):
  * cv2.resize(..., INTER_LINEAR) on uint8 (OpenCV 4.13 imgproc/resize.cpp: 11-bit fixed-point
    coefficients, x clamped at the table level, y clamped at row fetch): `resize_linear_u8`
    is bit-identical to cv2 4.13 on every size tried (tests/test_heatmap_oracle.py).
  * cv2.GaussianBlur(f32, (5,5), 1, 1), BORDER_REFLECT_101: `gaussian5` agrees to 2e-7 of the
    map maximum (the SIMD summation order of cv2 is not restated).
  * Matplotlib's jet: Matplotlib is absent from this image -> the LUT is restated from the
    published segment data, PARITY UNPINNED for the LUT; it is a caller-supplied table at the
    C ABI (bf_heatmap_dev lut argument).
The float32 steps use the same NumPy calls as the reference (np.log10 on float32 arrays, scalar
np.float32 ** int), so on one machine the restatement is bit-identical to the reference; NumPy's
SIMD log10/pow differ in the last ulp between CPUs, hence colour indices may differ by one step on
isolated pixels between machines (and between this file and the CUDA path).
"""
import numpy as np

POWER = 5                       # visual.py:13
WINDOW_DIMENSIONS = (1920, 1080)  # visual.py:9

_JET = {
    "red": ((0.00, 0, 0), (0.35, 0, 0), (0.66, 1, 1), (0.89, 1, 1), (1.00, 0.5, 0.5)),
    "green": ((0.000, 0, 0), (0.125, 0, 0), (0.375, 1, 1), (0.640, 1, 1), (0.910, 0, 0), (1.000, 0, 0)),
    "blue": ((0.00, 0.5, 0.5), (0.11, 1, 1), (0.34, 1, 1), (0.65, 0, 0), (1.00, 0, 0)),
}


def _segment_lut(data, n=256):
    a = np.array(data, dtype=float)
    x, y0, y1 = a[:, 0] * (n - 1), a[:, 1], a[:, 2]
    xind = (n - 1) * np.linspace(0, 1, n)
    ind = np.searchsorted(x, xind)[1:-1]
    dist = (xind[1:-1] - x[ind - 1]) / (x[ind] - x[ind - 1])
    lut = np.concatenate([[y1[0]], dist * (y0[ind] - y1[ind - 1]) + y1[ind - 1], [y0[-1]]])
    return np.clip(lut, 0.0, 1.0)


def generate_color_map():
    """visual.py:27-48: colors[i] = u8(jet(255 - i)[:3] * 255)."""
    rgb = np.stack([_segment_lut(_JET[c]) for c in ("red", "green", "blue")], axis=1)
    colors = np.empty((256, 3), np.uint8)
    for i in range(256):
        colors[i] = (rgb[255 - i] * 255).astype(np.uint8)
    return colors


def color_index(image, threshold=1e-7, amount=0.5, exponent=POWER, log_scale=True):
    """Index map [X][Y] (int16, -1 = not painted) + should_overlay, visual.py:143-166 / 173-200."""
    image = np.asarray(image, np.float32)
    X, Y = image.shape
    idx = np.full((X, Y), -1, np.int16)
    mx = np.max(image)
    if log_scale:
        safe = np.clip(image, 1e-12, None)
        if not (mx > threshold):
            return idx, False
        with np.errstate(invalid="ignore", divide="ignore"):
            img = np.log10(safe)
            img -= np.log10(np.min(safe))
            img /= np.max(img)
        overlay = True
    else:
        img = image / mx
        if not (mx > threshold):
            return idx, False
        overlay = False
    for x in range(X):
        for y in range(Y):
            p = img[x, y]
            if p >= amount:
                p -= amount
                p /= amount
                idx[x, y] = int(255 * p ** exponent)
                if not log_scale:
                    overlay = True
    return idx, overlay


def small_heatmap(idx, colors):
    """visual.py:166: small[Y-1-y, X-1-x] = colors[idx[x, y]]."""
    X, Y = idx.shape
    small = np.zeros((Y, X, 3), np.uint8)
    xs, ys = np.nonzero(idx >= 0)
    small[Y - 1 - ys, X - 1 - xs] = colors[idx[xs, ys]]
    return small


def _resize_coeffs(src, dst, clamp):
    scale = 1.0 / (dst / src)
    ofs = np.zeros(dst, np.int32)
    coef = np.zeros((dst, 2), np.int16)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if clamp:
            if s < 0:
                f, s = np.float32(0), 0
            if s >= src - 1:
                f, s = np.float32(0), src - 1
        coef[d, 0] = int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048))))
        coef[d, 1] = int(np.rint(np.float32(f * np.float32(2048))))
        ofs[d] = s
    return ofs, coef


def resize_tables(sh, sw, dh, dw):
    """(xofs, xcoef, yofs, ycoef) of cv2.resize INTER_LINEAR for 8-bit images."""
    xo, xc = _resize_coeffs(sw, dw, True)
    yo, yc = _resize_coeffs(sh, dh, False)
    return xo, xc, yo, yc


def resize_linear_u8(img, dsize):
    """cv2.resize(img, dsize=(W, H), interpolation=cv2.INTER_LINEAR) for uint8 [h][w][c]."""
    W, H = dsize
    squeeze = img.ndim == 2
    if squeeze:
        img = img[:, :, None]
    h, w, _ = img.shape
    if (h, w) == (H, W):
        return img[:, :, 0].copy() if squeeze else img.copy()
    xo, xc, yo, yc = resize_tables(h, w, H, W)
    S = img.astype(np.int32)
    x1 = np.minimum(xo + 1, w - 1)
    rows = S[:, xo, :] * xc[:, 0].astype(np.int32)[None, :, None] + \
        S[:, x1, :] * xc[:, 1].astype(np.int32)[None, :, None]
    y0 = np.clip(yo, 0, h - 1)
    y1 = np.clip(yo + 1, 0, h - 1)
    b0 = yc[:, 0].astype(np.int32)[:, None, None]
    b1 = yc[:, 1].astype(np.int32)[:, None, None]
    out = (((b0 * (rows[y0] >> 4)) >> 16) + ((b1 * (rows[y1] >> 4)) >> 16) + 2) >> 2
    out = out.astype(np.uint8)
    return out[:, :, 0] if squeeze else out


def calculate_heatmap(image, threshold=1e-7, amount=0.5, exponent=POWER, colors=None,
                      window=WINDOW_DIMENSIONS, log_scale=True):
    colors = generate_color_map() if colors is None else colors
    idx, overlay = color_index(image, threshold, amount, exponent, log_scale)
    return resize_linear_u8(small_heatmap(idx, colors), window), overlay


def gaussian_kernel5():
    x = np.arange(5) - 2.0
    k = np.exp(-(x * x) / 2.0)
    return (k / k.sum()).astype(np.float32)


def gaussian5(img):
    """cv2.GaussianBlur(f32, (5,5), sigmaX=1, sigmaY=1), reflect-101 border, separable f32."""
    k = gaussian_kernel5()
    p = np.pad(np.asarray(img, np.float32), 2, mode="reflect")
    h, w = img.shape
    r = (p[:, 2:2 + w] * k[2] + (p[:, 1:1 + w] + p[:, 3:3 + w]) * k[1]) + (p[:, 0:w] + p[:, 4:4 + w]) * k[0]
    r = r.astype(np.float32)
    o = (r[2:2 + h] * k[2] + (r[1:1 + h] + r[3:3 + h]) * k[1]) + (r[0:h] + r[4:4 + h]) * k[0]
    return o.astype(np.float32)


def find_power_center(image, smoothed=None):
    """visual.py:293-322 -> (centroid along axis 1, centroid along axis 0)."""
    sm = gaussian5(np.asarray(image, np.float32)) if smoothed is None else smoothed
    mx = np.max(sm)
    mask = sm >= mx * 0.95
    if np.sum(mask) > 0:
        yi, xi = np.indices(sm.shape)
        wgt = (sm ** 3) * mask
        tot = np.sum(wgt)
        if tot > 0:
            return float(np.sum(xi * wgt) / tot), float(np.sum(yi * wgt) / tot)
    pk = np.unravel_index(np.argmax(sm), sm.shape)
    return float(pk[1]), float(pk[0])


def detection_box(peak_x, peak_y, X, Y, window=WINDOW_DIMENSIONS, box_size_ratio=0.1):
    """visual.py:268-281 -> (centre_x, centre_y, x1, y1, x2, y2) in window pixels."""
    cx = window[0] - 1 - int(peak_x / (X - 1) * window[0])
    cy = window[1] - 1 - int(peak_y / (Y - 1) * window[1])
    bw, bh = int(window[0] * box_size_ratio), int(window[1] * box_size_ratio)
    return (cx, cy, max(0, cx - bw // 2), max(0, cy - bh // 2),
            min(window[0], cx + bw // 2), min(window[1], cy + bh // 2))


def get_entropy(heatmap):
    """decider.py:16-24."""
    s = np.sum(heatmap)
    hm = heatmap / s if s > 0 else np.zeros_like(heatmap)
    ent = -np.sum(hm * np.log(hm + 1e-12))
    return float(1 / (1 + ent))
