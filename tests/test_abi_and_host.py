"""CPU: the C-ABI library builds, loads and exports every symbol include/bf_b200.h declares;
host-side logic of the product (config, geometry scalars, steer offsets, data sources);
and the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from util import CASES, ROOT, bits_equal, gold, oracle_cfg, product_config


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "bf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:void|int|uint64_t|char)\s*\*?\s*(\w+)\s*\(", text, flags=re.M)
    return sorted(set(n for n in names if n != "bf_data_source_fn"))


def test_library_exports_every_declared_symbol():
    from lib import _native
    L = _native.lib()
    syms = _declared_symbols()
    assert len(syms) >= 50, syms
    for s in syms:
        assert hasattr(L, s), "missing export: " + s
    for must in ("mimo_pad", "mimo_lerp", "load_coefficients_pad", "load_coefficients_lerp", "miso_pad",
                 "miso_lerp", "pad_mimo", "lerp_mimo", "miso_steer_listen", "mimo_convolve_vectorized",
                 "mimo_convolve_hybrid", "bf_mimo_dev", "bf_miso_dev", "bf_generate_delays"):
        assert must in syms
    assert b"sm_100a" in L.bf_version()


def test_configure_roundtrip_and_validation():
    from lib import _native
    L = _native.lib()
    _native.configure(64, 256, 8, 20, 20, 128.0, -1)
    cfg = _native.BfConfig()
    assert L.bf_get_config(ctypes.byref(cfg)) == 0
    assert (cfg.n_microphones, cfg.n_samples, cfg.max_res_x, cfg.max_res_y) == (64, 256, 20, 20)
    with pytest.raises(RuntimeError):
        _native.configure(64, 4096, 8, 20, 20)          # N_SAMPLES > 1024 unsupported
    _native.configure(256, 256, 8, 57, 32)


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="box has a GPU")
def test_fails_loudly_without_gpu():
    from lib import _native
    L = _native.lib()
    _native.configure(256, 256, 8, 57, 32)
    w = np.zeros(57 * 32 * 256, np.int32)
    L.load_coefficients_pad(_native.ptr(w), w.size)
    assert L.bf_last_status() != 0
    assert b"no CPU fallback" in L.bf_last_error()
    with pytest.raises(RuntimeError):
        _native.check()
    sig = np.zeros((256, 256), np.float32)
    img = np.full(57 * 32, -1, np.float32)
    mics = np.arange(256, dtype=np.int32)
    L.mimo_pad(_native.ptr(sig), _native.ptr(img), _native.ptr(mics), 256)
    assert L.bf_last_status() != 0 and np.all(img == -1)      # nothing was computed on the CPU


@pytest.mark.parametrize("case", list(CASES))
def test_product_geometry_matches_reference(case, capsys):
    g = gold(case)
    config = product_config(case)
    from lib import directions
    mics, n = directions.active_microphones()
    assert np.array_equal(mics, g["mic_ids"]) and n == len(g["mic_ids"])
    assert bits_equal(directions.calc_r_prime(float(np.float32(0.02))), g["r_prime"])
    # scalar prologue of calculate_delays vs the oracle's restatement
    from oracle import directions_np as dn
    k, xs, ys, z2, d = directions._scan_scalars()
    k2, xs2, ys2, z22 = dn.scan_axes(oracle_cfg(case))
    assert k == k2 and z2 == z22 and bits_equal(xs, xs2) and bits_equal(ys, ys2)
    assert d == float(np.float32(0.02))
    product_config("default")


def test_tap_tables_match_reference_restatement():
    """get_h / get_h2 table builders (host NumPy in the reference too) vs oracle + golden sha."""
    from oracle import directions_np as dn
    from lib import directions
    from util import sha
    for case in ("ragged", "taps64"):
        g, cfg = gold(case), oracle_cfg(case)
        delays = dn.calculate_delays(cfg)
        taps = directions._get_h2_table(delays, cfg["N_TAPS"])
        assert sha(taps) == str(g["taps_sha"])
    d = np.array([0.0, 0.3, 2.75, 47.62])
    for x in d:
        assert bits_equal(directions.get_h(x), dn.get_h(x))
        assert bits_equal(directions.get_h2(x, 8), dn.get_h2(x, 8))
    frac = d - d.astype(int)
    tab = directions._get_h_table(frac)
    for i, x in enumerate(frac):
        assert bits_equal(tab[i], dn.get_h(x).astype(np.float32))


def test_steering_offsets_and_legacy_generators():
    config = product_config("default")
    from lib import beamformer, directions
    from oracle import directions_np as dn
    cfg = oracle_cfg("default")
    for az, el in ((0, 0), (-90, -90), (45.5, -10), (89.9, 89.9)):
        assert beamformer.steer_cartesian_degree(az, el) == dn.steer_offset_degree(cfg, az, el, 256)
    for x, y in ((0.5, 0.5), (0.0, 0.99), (0.31, 0.77)):
        assert beamformer.stear_miso_beam(x, y) == dn.steer_offset_unit(cfg, x, y, 256)
    # appendix A-16: azimuth 90 deg indexes one row past the table; the shim clamps
    raw = beamformer.steer_cartesian_degree(90, 90)
    assert raw > (57 * 32 - 1) * 256 and beamformer._miso.steer_offset == (57 * 32 - 1) * 256
    assert np.array_equal(directions.calculate_delay_miso(20.0, -35.0), dn.calculate_delay_miso(cfg, 20.0, -35.0))


def test_interface_config_reads_reference_schema(tmp_path):
    import json
    from interface import config
    assert config.BUFFER_LENGTH == config.N_SAMPLES * config.N_MICROPHONES
    assert config.NP_DTYPE is np.float32 and config.DTYPE is ctypes.c_int32
    data = {"general": {"N_MICROPHONES": 64, "N_SAMPLES": 128, "MAX_RES_X": 5, "MAX_RES_Y": 4, "COLUMNS": 8,
                        "ROWS": 8, "expression": {"BUFFER_LENGTH": "N_SAMPLES * N_MICROPHONES"}},
            "python": {"imports": ["numpy"], "expression": {"NP_DTYPE": "numpy.float32"}},
            "c": {"MIC_GAIN": 64, "expression": {}}}
    p = tmp_path / "config.json"
    p.write_text(json.dumps(data))
    old = config.CONFIG_PATH
    try:
        config.reload(str(p))
        assert config.BUFFER_LENGTH == 64 * 128 and config.MIC_GAIN == 64
    finally:
        config.reload(old)
    assert config.N_MICROPHONES == 256


def test_array_source_replays_record_format():
    config = product_config("default")
    from lib import beamformer
    rec = np.arange(256 * 512, dtype=np.float32).reshape(256, 512)
    src = beamformer.ArraySource(rec)
    buf = np.empty((256, 256), np.float32)
    src(buf)
    assert np.array_equal(buf, rec[:, :256])
    src(buf)
    assert np.array_equal(buf, rec[:, 256:])
    src(buf)
    assert np.array_equal(buf, rec[:, :256])
    with pytest.raises(RuntimeError):
        beamformer.connect()          # the UDP receiver is not part of this library


def test_synthetic_inputs_are_deterministic():
    from lib import synthetic
    g = gold("c1")
    sig = synthetic.point_sources(g["delays"].reshape(-1, 64), g["mic_ids"], 64, 256, 48828.0,
                                  synthetic.C1["sources"], synthetic.C1["noise"], synthetic.C1["seed"])
    assert np.allclose(sig, g["signals"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("cfg,M,N", [("default", 256, 256), ("c1", 64, 256)])
def test_shared_memory_record_layouts_match_the_reference_headers(cfg, M, N):
    """bf_layout_{miso,padata,ring_buffer} reproduce sizeof / offsetof of the reference's own
    structs (api.h:26-38, receiver.h:31-36) as compiled from its headers for this configuration
    (oracle/build_ref.py:build_record_layouts -> oracle/_ref/<cfg>/record_layouts.json), and bf_datagram_header
    is the head of receiver.h:51-59 `msg`."""
    import json
    path = os.path.join(ROOT, "oracle", "_ref", cfg, "record_layouts.json")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built")
    ref = json.load(open(path))
    from lib import _native
    L = _native.lib()
    lm = L.bf_layout_miso(M, N)
    assert [lm.size, *lm.off] == ref["Miso"]
    lp = L.bf_layout_padata(N)
    assert [lp.size, lp.off[0], lp.off[1]] == ref["paData"]
    lr = L.bf_layout_ring_buffer(M, N)
    assert [lr.size, lr.off[0], lr.off[1], lr.off[2]] == ref["ring_buffer"]
    import struct
    assert ref["msg"] == [8 + 4 * M, 0, 2, 3, 4, 8] and struct.calcsize("<Hbbi") == 8      # bf_datagram_header
