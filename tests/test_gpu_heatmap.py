"""GPU: heat-map post-processing (csrc/heatmap.cu through the C ABI) against the oracle
(oracle/heatmap_np.py) and the golden vectors produced by the reference's visual.py / decider.py.

Bars: the resize is bit-exact (integer work); colour indices are exact except isolated one-step
differences where a float32 log10 / pow result straddles a quantisation boundary (<= 0.2 % of the
pixels, |step| <= 1); centroids agree to 1e-3 grid cells; entropy confidence to 1e-12 relative."""
import numpy as np
import pytest

from oracle import heatmap_np as hn
from util import gold, product_config

pytestmark = pytest.mark.gpu

CASES = ["c1", "default", "c3"]
MAPS = ["map", "floor", "quiet", "flat"]


def maps_close(a, b, max_frac=2e-3, max_step=12):
    a, b = a.astype(int), b.astype(int)
    diff = np.abs(a - b).max(axis=-1)
    return (diff > 0).mean() <= max_frac and diff.max() <= max_step


@pytest.fixture(params=CASES)
def case(request):
    product_config(request.param)
    yield request.param, gold("heat_" + request.param)
    product_config("default")


def test_reference_entry_points_vs_golden(case):
    from lib import visual
    name, g = case
    X, Y = int(g["X"]), int(g["Y"])
    for m in MAPS:
        img = g["in_" + m]
        small, ov = visual.calculate_heatmap(img.copy(), window=(X, Y))
        assert ov == bool(g["overlay_" + m]), m
        assert small.shape == (Y, X, 3) and maps_close(small, g["small_" + m]), m
        cx, cy = visual.find_power_center(np.clip(img, 1e-12, None))
        assert abs(cx - g["center_" + m][0]) < 1e-3 and abs(cy - g["center_" + m][1]) < 1e-3, (m, cx, cy)
        heat, ov2 = visual.calculate_heatmap(img.copy(), window=(640, 360))
        assert heat.shape == (360, 640, 3) and ov2 == ov
        # the window-sized map equals the bit-exact resize of the small map this library produced
        assert np.array_equal(heat, hn.resize_linear_u8(small, (640, 360))), m
    box, heat, ov = visual.calculate_heatmap_with_detection(g["in_map"].copy(), window=(640, 360))
    want = hn.detection_box(g["center_map"][1], g["center_map"][0], X, Y, (640, 360))
    assert ov and max(abs(a - b) for a, b in zip(box, want)) <= 1
    box, _, ov = visual.calculate_heatmap_with_detection(g["in_quiet"].copy(), window=(640, 360))
    assert box is None and not ov


def test_linear_variant_vs_golden(case):
    from interface import config
    from lib import visual
    name, g = case
    if "in_fft_map" not in g:
        pytest.skip("grid smaller than 11x11")
    config.reload(MAX_RES_X=11, MAX_RES_Y=11)
    for m in MAPS:
        img = g["in_fft_" + m].copy()
        small, ov = visual.calculate_heatmap_fft(img, threshold=1e-13, window=(11, 11))
        assert ov == bool(g["fftoverlay_" + m]), m
        assert maps_close(small, g["fftsmall_" + m], max_frac=0.02), m


def test_device_batch_index_info_entropy(case):
    import torch
    from lib import _native, visual
    name, g = case
    X, Y = int(g["X"]), int(g["Y"])
    stack = np.stack([g["in_" + m] for m in MAPS] + [g["in_map"] * np.float32(0.37)])
    d = torch.from_numpy(stack.reshape(len(stack), -1)).cuda()
    res = visual.heatmaps_dev(d, window=(640, 360), confidence=True)
    torch.cuda.synchronize()
    info = res["info"].cpu().numpy().view(_native.HEAT_INFO_DTYPE).ravel()
    index = res["index"].cpu().numpy()
    small = res["small"].cpu().numpy()
    heat = res["heat"].cpu().numpy()
    conf = res["confidence"].cpu().numpy()
    lut = hn.generate_color_map()
    for k, img in enumerate(stack):
        want_idx, want_ov = hn.color_index(img)
        assert bool(info["overlay"][k]) == want_ov
        d_idx = np.abs(index[k].astype(int) - want_idx.astype(int))
        assert d_idx.max() <= 1 and (d_idx > 0).mean() <= 2e-3, (k, d_idx.max(), (d_idx > 0).mean())
        assert np.array_equal(small[k], hn.small_heatmap(index[k], lut))          # LUT + flipped store: exact
        assert np.array_equal(heat[k], hn.resize_linear_u8(small[k], (640, 360)))  # resize: exact
        assert int(info["painted"][k]) == int((index[k] >= 0).sum())
        assert info["max_power"][k] == img.max() and info["min_power"][k] == np.clip(img, 1e-12, None).min()
        cx, cy = hn.find_power_center(np.clip(img, 1e-12, None))
        assert abs(info["center_col"][k] - cx) < 1e-3 and abs(info["center_row"][k] - cy) < 1e-3
        sm = hn.gaussian5(np.clip(img, 1e-12, None))
        assert abs(info["smooth_max"][k] - sm.max()) <= 1e-6 * sm.max()
        want_conf = hn.get_entropy(heat[k])
        assert abs(conf[k] - want_conf) <= 1e-12 * max(1.0, want_conf), (k, conf[k], want_conf)
    # scaling a map by a constant leaves the log-scale colours (almost) alone
    assert maps_close(small[0], small[4], max_frac=5e-3)


def test_resize_bit_exact_random():
    import torch
    from lib import _native
    L = _native.lib()
    rng = np.random.default_rng(7)
    for (h, w, H, W, cn, frames) in [(32, 57, 1080, 1920, 3, 1), (180, 180, 360, 640, 3, 3), (11, 11, 480, 720, 3, 2),
                                      (7, 11, 100, 333, 3, 2), (13, 17, 101, 203, 1, 2), (20, 20, 77, 50, 4, 1),
                                      (40, 30, 10, 7, 3, 1), (20, 20, 20, 20, 3, 2)]:
        src = rng.integers(0, 256, (frames, h, w, cn), dtype=np.uint8)
        d_src = torch.from_numpy(src).cuda()
        d_dst = torch.empty((frames, H, W, cn), dtype=torch.uint8, device="cuda")
        _native.check(L.bf_resize_linear_u8_dev(d_src.data_ptr(), frames, h, w, cn, d_dst.data_ptr(), H, W, None))
        torch.cuda.synchronize()
        got = d_dst.cpu().numpy()
        for f in range(frames):
            assert np.array_equal(got[f], hn.resize_linear_u8(src[f], (W, H))), (h, w, H, W, cn, f)


def test_entropy_and_errors():
    import torch
    from lib import _native
    from sensorfusion.decider import sensorfusiondecider
    dec = sensorfusiondecider()
    rng = np.random.default_rng(3)
    for img in [rng.integers(0, 256, (360, 640, 3), dtype=np.uint8), np.zeros((50, 33, 3), np.uint8),
                (rng.random((77, 13)) > 0.97).astype(np.uint8) * 200]:
        want = hn.get_entropy(img)
        got = dec.get_entropy(img)
        assert abs(got - want) <= 1e-12 * max(1.0, want)
    assert dec.focus_beam(lambda h, v: None, (0, 0, 10, 10, 0.1)) == (-1, -1)
    seen = []
    assert dec.focus_beam(lambda h, v: seen.append((h, v)), (160, 90, 480, 270, 0.9)) == 0
    assert seen == [(0.0, 0.0)]
    L = _native.lib()
    big = torch.zeros((1, 300 * 300), dtype=torch.float32, device="cuda")
    small = torch.empty((1, 300, 300, 3), dtype=torch.uint8, device="cuda")
    info = torch.empty((1, 48), dtype=torch.uint8, device="cuda")
    rc = L.bf_heatmap_dev(big.data_ptr(), 1, 300 * 300, 300, 300, 1e-7, 0.5, 5, 1, None, small.data_ptr(), None,
                          info.data_ptr(), None)
    assert rc != 0 and b"shared-memory" in L.bf_last_error()


def test_replay_video_pipeline_matches_stagewise():
    """lib.replay.video_dev (recording -> maps -> overlay, all on the device) == the stages run one by one."""
    import torch
    from lib import _native, directions, replay, visual
    config = product_config("c1")
    _native.configure_from(config)
    L = _native.lib()
    M, N = config.N_MICROPHONES, config.N_SAMPLES
    mics, n = directions.active_microphones()
    mics = _native.i32(mics)
    whole = _native.i32(directions.calculate_delays().astype(int)).ravel()
    L.load_coefficients_pad(_native.ptr(whole), whole.size)
    _native.check()
    gen = torch.Generator(device="cuda").manual_seed(9)
    rec = 0.05 * torch.randn((M, 48828 // 2), generator=gen, device="cuda")
    rec[:, 3000:9000] += 0.2 * torch.sin(torch.arange(6000, device="cuda") * 0.6)[None, :]
    d_mics = torch.from_numpy(mics).cuda()
    out = replay.video_dev(_native.ALGO_PAD, rec, d_mics, n, window=(64, 36), chunk=5)
    idx, maps = replay.replay_dev(_native.ALGO_PAD, rec, d_mics, n)
    torch.cuda.synchronize()
    assert np.array_equal(out["frames"], idx) and torch.equal(out["maps"], maps) and len(idx) == 15
    ref = visual.heatmaps_dev(maps, window=(64, 36), confidence=True)
    torch.cuda.synchronize()
    assert torch.equal(out["heat"], ref["heat"]) and torch.equal(out["info"], ref["info"])
    assert torch.equal(out["confidence"], ref["confidence"])
    info = out["info"].cpu().numpy().view(_native.HEAT_INFO_DTYPE).ravel()
    assert info["overlay"].all() and np.all(info["max_power"] > 0)
    product_config("default")
