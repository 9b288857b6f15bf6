"""Golden vectors for the heat-map post-processing stage.

Runs the REFERENCE's own functions (PC/src/visual.py, PC/sensorfusion/decider.py, imported in
place from /root/reference with this container's cv2 4.13 / NumPy) on the power maps already
stored in tests/golden/{c1,default,c3}.npz and writes tests/golden/heat_<cfg>.npz.

    python oracle/gen_golden_heat.py <cfg>        (one clean interpreter per cfg: the reference
                                                   reads MAX_RES_X/Y from interface.config at import)

Matplotlib is absent here: plt.cm.get_cmap is stubbed with the jet table restated in
oracle/heatmap_np.py (LUT parity unpinned, see there); everything downstream of the LUT is the
reference's code.  TEST INFRASTRUCTURE ONLY.
"""
import hashlib
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_reference(cfg):
    sys.path.insert(0, ROOT)
    from oracle import heatmap_np as hn
    lut = hn.generate_color_map()
    # Matplotlib stub: cmap(j) -> RGBA floats such that u8(rgb * 255) == lut[255 - j]
    segs = np.stack([hn._segment_lut(hn._JET[c]) for c in ("red", "green", "blue")], axis=1)
    plt = types.ModuleType("matplotlib.pyplot")
    plt.cm = types.SimpleNamespace(get_cmap=lambda name="jet": (lambda j: (*segs[j], 1.0)))
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    work = os.path.join(HERE, "_ref", cfg)
    os.chdir(work)
    sys.path.insert(0, work)

    def imp(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    visual = imp("ref_visual", os.path.join(REF, "PC", "src", "visual.py"))
    decider = imp("ref_decider", os.path.join(REF, "PC", "sensorfusion", "decider.py"))
    assert np.array_equal(visual.colors, lut)
    return visual, decider


def main(cfg):
    visual, decider = load_reference(cfg)
    X, Y = visual.MAX_RES_X, visual.MAX_RES_Y
    g = np.load(os.path.join(GOLD, cfg + ".npz"))
    base = g["img_pad"].reshape(X, Y).astype(np.float32)
    rng = np.random.default_rng(4321)
    maps = {
        "map": base,
        # a second, flatter map: the same scene over a noise floor (exercises the log scale)
        "floor": (base + np.float32(0.2) * base.max() * rng.random((X, Y)).astype(np.float32)),
        "quiet": (base * np.float32(1e-9 / max(base.max(), 1e-30))).astype(np.float32),   # max < threshold
        "flat": np.full((X, Y), 3.0e-4, np.float32),                                        # 0/0 -> nothing painted
    }
    out = {"lut": visual.colors, "X": np.array(X), "Y": np.array(Y)}
    dec = decider.sensorfusiondecider()
    for name, m in maps.items():
        out["in_" + name] = m
        visual.WINDOW_DIMENSIONS = (X, Y)                 # identity resize -> the small map itself
        small, ov = visual.calculate_heatmap(m.copy())
        out["small_" + name] = small
        out["overlay_" + name] = np.array(bool(ov))
        visual.WINDOW_DIMENSIONS = (640, 360)
        mid, _ = visual.calculate_heatmap(m.copy())
        out["sha640_" + name] = np.array(sha(mid))
        if name == "map":
            out["heat640_map"] = mid
        visual.WINDOW_DIMENSIONS = (1920, 1080)
        big, _ = visual.calculate_heatmap(m.copy())
        out["sha1920_" + name] = np.array(sha(big))
        safe = np.clip(m, 1e-12, None)
        cx, cy = visual.find_power_center(safe)
        out["center_" + name] = np.array([cx, cy], np.float64)
        out["entropy_" + name] = np.array(dec.get_entropy(mid), np.float64)
        # the linear variant (calculate_heatmap_fft hard-codes an 11x11 grid: top-left crop)
        if X >= 11 and Y >= 11:
            visual.WINDOW_DIMENSIONS = (11, 11)
            crop = np.ascontiguousarray(m[:11, :11]).copy()
            out["in_fft_" + name] = crop.copy()
            fft_small, fov = visual.calculate_heatmap_fft(crop, threshold=1e-13)
            out["fftsmall_" + name] = fft_small
            out["fftoverlay_" + name] = np.array(bool(fov))
    np.savez_compressed(os.path.join(GOLD, "heat_%s.npz" % cfg), **out)
    print("heat_%s: %dx%d  overlay" % (cfg, X, Y), {k: bool(out["overlay_" + k]) for k in maps},
          "center(map)", out["center_map"], "entropy(map)", float(out["entropy_map"]))


if __name__ == "__main__":
    main(sys.argv[1])
