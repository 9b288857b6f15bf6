"""CPU: the heat-map oracle (oracle/heatmap_np.py) against the golden vectors made by the
reference's own visual.py / decider.py (oracle/gen_golden_heat.py) and, when cv2 is importable,
against OpenCV directly.  Also the host-only parts of the product (jet table, box arithmetic)."""
import numpy as np
import pytest

from oracle import heatmap_np as hn
from util import gold, sha

CASES = ["c1", "default", "c3"]
MAPS = ["map", "floor", "quiet", "flat"]


def maps_close(a, b, max_frac=2e-3, max_step=12):
    """Colour maps equal up to isolated one-step colour-index differences (NumPy's SIMD log10 / pow are
    1-2 ulp off a correctly rounded result and differ between CPUs, see oracle/heatmap_np.py)."""
    a, b = a.astype(int), b.astype(int)
    diff = np.abs(a - b).max(axis=-1)
    return (diff > 0).mean() <= max_frac and diff.max() <= max_step


@pytest.mark.parametrize("case", CASES)
def test_lut_and_small_maps(case):
    g = gold("heat_" + case)
    lut = hn.generate_color_map()
    assert np.array_equal(lut, g["lut"])
    for m in MAPS:
        idx, ov = hn.color_index(g["in_" + m])
        assert ov == bool(g["overlay_" + m]), m
        small = hn.small_heatmap(idx, lut)
        assert small.shape == g["small_" + m].shape
        assert maps_close(small, g["small_" + m]), m
    assert (hn.color_index(g["in_quiet"])[0] == -1).all()
    assert (hn.color_index(g["in_flat"])[0] == -1).all()          # 0/0 -> nothing painted, overlay still True


@pytest.mark.parametrize("case", CASES)
def test_resize_bit_exact_from_golden_small(case):
    g = gold("heat_" + case)
    for m in MAPS:
        small = g["small_" + m]
        assert sha(hn.resize_linear_u8(small, (640, 360))) == str(g["sha640_" + m]), m
    assert np.array_equal(hn.resize_linear_u8(g["small_map"], (640, 360)), g["heat640_map"])
    assert sha(hn.resize_linear_u8(g["small_map"], (1920, 1080))) == str(g["sha1920_map"])


@pytest.mark.parametrize("case", CASES)
def test_center_and_entropy(case):
    g = gold("heat_" + case)
    for m in MAPS:
        safe = np.clip(g["in_" + m], 1e-12, None)
        cx, cy = hn.find_power_center(safe)
        assert abs(cx - g["center_" + m][0]) < 1e-3 and abs(cy - g["center_" + m][1]) < 1e-3, m
        heat = hn.resize_linear_u8(g["small_" + m], (640, 360))
        assert abs(hn.get_entropy(heat) - float(g["entropy_" + m])) <= 1e-12 * max(1.0, float(g["entropy_" + m])), m


@pytest.mark.parametrize("case", CASES)
def test_linear_variant(case):
    g = gold("heat_" + case)
    lut = hn.generate_color_map()
    for m in MAPS:
        if "in_fft_" + m not in g:
            continue
        idx, ov = hn.color_index(g["in_fft_" + m], threshold=1e-13 * 1000000, amount=0.5, exponent=2, log_scale=False)
        assert ov == bool(g["fftoverlay_" + m]), m
        assert maps_close(hn.small_heatmap(idx, lut), g["fftsmall_" + m], max_frac=0.02), m


def test_against_opencv_directly():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for (h, w, H, W) in [(32, 57, 1080, 1920), (20, 20, 360, 640), (11, 11, 480, 720), (7, 11, 100, 333),
                         (180, 180, 360, 640), (57, 32, 270, 480), (40, 30, 10, 7)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(hn.resize_linear_u8(img, (W, H)), cv2.resize(img, (W, H), interpolation=cv2.INTER_LINEAR))
    gray = rng.integers(0, 256, (13, 17), dtype=np.uint8)
    assert np.array_equal(hn.resize_linear_u8(gray, (200, 100)), cv2.resize(gray, (200, 100), interpolation=cv2.INTER_LINEAR))
    assert np.array_equal(cv2.getGaussianKernel(5, 1.0, cv2.CV_32F).ravel(), hn.gaussian_kernel5())
    for shape in [(180, 180), (57, 32), (20, 20), (5, 9)]:
        img = (rng.random(shape) ** 4 * 1e-3).astype(np.float32)
        ref = cv2.GaussianBlur(img, (5, 5), sigmaX=1.0, sigmaY=1.0)
        assert np.abs(hn.gaussian5(img) - ref).max() <= 4e-7 * ref.max()


def test_product_host_parts():
    """jet table and box arithmetic of the product's lib.visual (host-only, no GPU call)."""
    from lib import visual
    from util import product_config
    assert np.array_equal(visual.generate_color_map(), hn.generate_color_map())
    product_config("c3")
    for peak in [(0.0, 0.0), (98.655, 40.441), (179.0, 179.0), (13.2, 170.9)]:
        assert visual.detection_box(*peak) == hn.detection_box(peak[0], peak[1], 180, 180)
        assert visual.detection_box(*peak, window=(640, 360)) == hn.detection_box(peak[0], peak[1], 180, 180, (640, 360))
    product_config("default")
