"""GPU: device-resident, batched and direction-sharded entry points (bf_mimo_dev, bf_miso_dev),
the MISO stream kernel, the producer loops of lib.beamformer, and size-independent properties at
the full BASELINE sizes.  PyTorch is used only to own device memory (tensor.data_ptr())."""
import numpy as np
import pytest

from util import bits_equal, gold, product_config, sha

pytestmark = pytest.mark.gpu


def _setup(case):
    config = product_config(case)
    from lib import _native as nat
    nat.configure_from(config)
    nat.lib().bf_set_kernel_options(0, 1)
    return config, nat, nat.lib()


def _torch():
    import torch
    return torch


def test_batched_and_sharded_maps_equal_single_calls():
    config, nat, L = _setup("c1")
    torch = _torch()
    from lib import directions
    g = gold("c1")
    mics = nat.i32(g["mic_ids"])
    D, n, N, M = 400, 64, 256, 64
    whole, d32 = directions.whole_and_f32()
    L.load_coefficients_pad(nat.ptr(whole), whole.size)
    L.load_coefficients_lerp(nat.ptr(d32), d32.size)
    nat.check()
    F = 5
    rng = np.random.default_rng(11)
    frames = rng.standard_normal((F, M, N)).astype(np.float32)
    frames[0] = g["signals"]
    d_sig = torch.from_numpy(frames).cuda()
    d_mics = torch.from_numpy(mics).cuda()
    for algo, name, key in ((nat.ALGO_PAD, "mimo_pad", "img_pad"), (nat.ALGO_LERP, "mimo_lerp", "img_lerp")):
        single = np.zeros((F, D), np.float32)
        for f in range(F):
            getattr(L, name)(nat.ptr(frames[f]), nat.ptr(single[f]), nat.ptr(mics), n)
            nat.check()
        assert bits_equal(single[0], g[key])
        d_img = torch.full((F, D), float("nan"), device="cuda")
        nat.check(L.bf_mimo_dev(algo, d_sig.data_ptr(), d_img.data_ptr(), F, d_mics.data_ptr(), n, 0, D, None))
        torch.cuda.synchronize()
        assert bits_equal(d_img.cpu().numpy(), single)
        # direction sharding: 3 uneven slices written into one image buffer
        d_img2 = torch.full((F, D), float("nan"), device="cuda")
        for lo, cnt in ((0, 133), (133, 134), (267, 133)):
            nat.check(L.bf_mimo_dev(algo, d_sig.data_ptr(), d_img2.data_ptr(), F, d_mics.data_ptr(), n, lo, cnt, None))
        torch.cuda.synchronize()
        assert bits_equal(d_img2.cpu().numpy(), single)


@pytest.mark.parametrize("algo_name", ["pad", "lerp"])
@pytest.mark.parametrize("n_use,blocks", [(64, 300), (200, 37), (5, 8)])
def test_miso_stream_vs_oracle(algo_name, n_use, blocks):
    from oracle import cpu
    torch = _torch()
    from lib import _native as nat
    L = nat.lib()
    M, N, X, Y = 256, 256, 6, 5
    nat.configure(M, N, 8, X, Y, 128.0)
    rng = np.random.default_rng(blocks + n_use)
    sig = rng.standard_normal((blocks, M, N)).astype(np.float32)
    mics = nat.i32(np.sort(rng.permutation(M)[:n_use]))
    whole = rng.integers(0, 60, (X * Y, n_use)).astype(np.int32)
    d32 = (rng.random((X * Y, n_use)) * 59.5).astype(np.float32)
    L.load_coefficients_pad(nat.ptr(whole), whole.size)
    L.load_coefficients_lerp(nat.ptr(d32), d32.size)
    nat.check()
    off = 17 * n_use
    algo = nat.ALGO_PAD if algo_name == "pad" else nat.ALGO_LERP
    d_sig, d_mics = torch.from_numpy(sig).cuda(), torch.from_numpy(mics).cuda()
    for scale in (0, 1):
        d_out = torch.full((blocks, N), float("nan"), device="cuda")
        nat.check(L.bf_miso_dev(algo, d_sig.data_ptr(), d_out.data_ptr(), blocks, d_mics.data_ptr(), n_use, off, scale, None))
        torch.cuda.synchronize()
        got = d_out.cpu().numpy()
        for b in (0, 1, blocks // 2, blocks - 1):
            ref = cpu.miso_pad(sig[b], mics, whole, off) if algo_name == "pad" else cpu.miso_lerp(sig[b], mics, d32, off)
            if scale:
                ref = cpu.miso_scale(ref, n_use, 128.0)
            assert bits_equal(got[b], ref), (b, scale)
    # the simple kernel gives the same stream
    L.bf_set_kernel_options(1, 1)
    d_out2 = torch.zeros((blocks, N), device="cuda")
    nat.check(L.bf_miso_dev(algo, d_sig.data_ptr(), d_out2.data_ptr(), blocks, d_mics.data_ptr(), n_use, off, 1, None))
    torch.cuda.synchronize()
    L.bf_set_kernel_options(0, 1)
    assert bits_equal(d_out2.cpu().numpy(), got)


@pytest.mark.parametrize("taps", [8, 32])
@pytest.mark.parametrize("n_use,blocks", [(64, 70), (37, 9)])
def test_miso_fir_hybrid_stream_vs_oracle(taps, n_use, blocks):
    """FIR (both accumulation orders) and hybrid MISO streams through the ring kernel == per-sample kernel ==
    oracle, bit for bit: fused (8 taps) and unfused (32 taps) chains, shuffled microphones, post-scale."""
    from oracle import cpu
    torch = _torch()
    from lib import _native as nat
    L = nat.lib()
    M, N, X, Y = 256, 256, 6, 5
    D = X * Y
    nat.configure(M, N, taps, X, Y, 128.0)
    rng = np.random.default_rng(blocks + n_use + taps)
    sig = rng.standard_normal((blocks, M, N)).astype(np.float32)
    mics = nat.i32(rng.permutation(M)[:n_use])
    h = (rng.standard_normal((D, n_use, taps)) / taps).astype(np.float32)
    d32 = (rng.random((D, n_use)) * 59.5).astype(np.float32)
    d32[17, :3] = [0.0, 5.0, 254.5]
    L.load_coefficients_convolve(nat.ptr(h), h.size)
    L.load_coefficients_convolve_hybrid(nat.ptr(d32), d32.size)
    nat.check()
    off = 17 * n_use
    d_sig, d_mics = torch.from_numpy(sig).cuda(), torch.from_numpy(mics).cuda()
    # offset units follow the reference: tap-table FLOATS for FIR (d*n*T), entries for hybrid (d*n)
    for algo, off_a, ref_fn in (
            (nat.ALGO_FIR_SEQ, off * taps, lambda b: cpu.miso_fir(sig[b], mics, h, off * taps, taps, 0)),
            (nat.ALGO_FIR_LANES, off * taps, lambda b: cpu.miso_fir(sig[b], mics, h, off * taps, taps, 1)),
            (nat.ALGO_HYBRID, off, lambda b: cpu.miso_hybrid(sig[b], mics, d32, off, taps))):
        outs = []
        for simple in (0, 1):
            L.bf_set_kernel_options(simple, 1)
            d_out = torch.full((blocks, N), float("nan"), device="cuda")
            nat.check(L.bf_miso_dev(algo, d_sig.data_ptr(), d_out.data_ptr(), blocks, d_mics.data_ptr(), n_use, off_a, 0, None))
            torch.cuda.synchronize()
            outs.append(d_out.cpu().numpy())
        L.bf_set_kernel_options(0, 1)
        assert bits_equal(outs[0], outs[1]), algo
        for b in (0, blocks // 2, blocks - 1):
            assert bits_equal(outs[0][b], ref_fn(b)), (algo, b)
        d_out = torch.zeros((blocks, N), device="cuda")
        nat.check(L.bf_miso_dev(algo, d_sig.data_ptr(), d_out.data_ptr(), blocks, d_mics.data_ptr(), n_use, off_a, 1, None))
        torch.cuda.synchronize()
        assert bits_equal(d_out.cpu().numpy()[1], cpu.miso_scale(ref_fn(1), n_use, 128.0)), algo


def test_full_size_c3_properties():
    """BASELINE config C3 (256 mics, 180x180 grid) at full size: golden image from the reference,
    tiled == simple on a slice, permutation/idempotence/scaling properties."""
    config, nat, L = _setup("c3")
    torch = _torch()
    from lib import directions
    g = gold("c3")
    sig, mics = np.ascontiguousarray(g["signals"]), nat.i32(g["mic_ids"])
    D, n = 32400, 256
    directions.load_pad_from_geometry()                # generator -> device table, no host trip
    img = np.zeros(D, np.float32)
    L.mimo_pad(nat.ptr(sig), nat.ptr(img), nat.ptr(mics), n)
    nat.check()
    assert bits_equal(img, g["img_pad"])
    img_again = np.zeros(D, np.float32)
    L.mimo_pad(nat.ptr(sig), nat.ptr(img_again), nat.ptr(mics), n)
    assert bits_equal(img, img_again)                   # idempotent / deterministic
    # the strongest source (2 kHz, A = 0.1) dominates: global maximum on its grid column +-3
    ix = int(np.argmax(img)) // 180
    assert abs(ix - 40) <= 3, ix
    # exact power-of-two scaling: signals*2 -> image*4, bit for bit
    img4 = np.zeros(D, np.float32)
    sig2 = (sig * np.float32(2)).astype(np.float32)
    L.mimo_pad(nat.ptr(sig2), nat.ptr(img4), nat.ptr(mics), n)
    assert bits_equal(img4, img * np.float32(4))
    # lerp at full size
    directions.load_lerp_from_geometry()
    imgl = np.zeros(D, np.float32)
    L.mimo_lerp(nat.ptr(sig), nat.ptr(imgl), nat.ptr(mics), n)
    nat.check()
    assert bits_equal(imgl, g["img_lerp"])
    # tiled == simple on a 2000-direction slice (device API)
    d_sig, d_mics = torch.from_numpy(sig[None]).cuda(), torch.from_numpy(mics).cuda()
    outs = []
    for simple in (0, 1):
        L.bf_set_kernel_options(simple, 1)
        d_img = torch.zeros((1, D), device="cuda")
        nat.check(L.bf_mimo_dev(nat.ALGO_LERP, d_sig.data_ptr(), d_img.data_ptr(), 1, d_mics.data_ptr(), n, 15000, 2003, None))
        torch.cuda.synchronize()
        outs.append(d_img.cpu().numpy()[0, 15000:17003])
    L.bf_set_kernel_options(0, 1)
    assert bits_equal(outs[0], outs[1]) and bits_equal(outs[0], g["img_lerp"][15000:17003])
    for i, d in enumerate(g["miso_dirs"]):
        out = np.zeros(256, np.float32)
        L.miso_lerp(nat.ptr(sig), nat.ptr(out), nat.ptr(mics), n, int(d) * n)
        assert bits_equal(out, g["miso_lerp"][i])


@pytest.mark.parametrize("target", ["b", "uti_api"])
def test_producer_loop_in_child_process(target, tmp_path):
    """The reference runs its loops in multiprocessing.Process children created by fork
    (main.pyx:702-721); CUDA must initialise lazily inside the child, so the scenario runs from a
    fresh interpreter that never touched CUDA (tests/_loop_child.py).  Payload contract: `b` puts
    (power_map, frame_nr), the api loops put the bare map (camera.py:83, visual.py:421)."""
    import os, subprocess, sys
    g = gold("c1")
    out = str(tmp_path / "loop.npz")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_loop_child.py")
    subprocess.run([sys.executable, script, target, out], check=True, timeout=280)
    z = np.load(out)
    maps, nrs = z["maps"], z["nrs"]
    if target == "b":
        assert list(nrs) == [1, 2, 3]
    else:
        assert list(nrs) == [-1, -1, -1]
    assert maps.shape == (3, 20, 20) and maps.dtype == np.float32 and z["contiguous"].all()
    assert bits_equal(maps[0].ravel(), g["img_pad"])


def test_miso_beam_listen_and_steering():
    config, nat, L = _setup("c1")
    from oracle import cpu
    from lib import beamformer, directions
    g = gold("c1")
    beamformer.connect(False, verbose=False, source=beamformer.ArraySource(g["signals"]))
    try:
        whole, _ = directions.whole_and_f32()
        L.load_coefficients_pad(nat.ptr(whole), whole.size)
        mics, n = directions.active_microphones()
        beamformer.load_pa(mics, n)
        off = beamformer.stear_miso_beam(14 / 20 + 0.01, 6 / 20 + 0.01)
        assert off == int(6 * 20 * 64 + 14 * 64)
        audio = beamformer.miso_listen(scaled=True)
        ref = cpu.miso_scale(cpu.miso_pad(g["signals"], mics, whole, off), n, 128.0)
        assert bits_equal(audio, ref)
    finally:
        beamformer.disconnect()


def test_miso_record_listen_on_the_reference_shm_layout():
    """one iteration of the reference's audio child (api.c:505-529) on a `Miso` record laid out
    as api.h:32-38 declares it (what load_pa() / steer() write into the SysV segment) and a `paData` record
    (api.h:26-30): beam of miso->signals at miso->steer_offset, post-scaled into pa->out, pa->can_read = 1."""
    import ctypes
    config, nat, L = _setup("c1")
    from oracle import cpu
    from lib import directions
    g = gold("c1")
    M, N = 64, 256
    whole, _ = directions.whole_and_f32()
    L.load_coefficients_pad(nat.ptr(whole), whole.size)
    nat.check()
    mics = nat.i32(g["mic_ids"])
    n = len(mics)
    off = (6 * 20 + 14) * n
    lm, lp = L.bf_layout_miso(M, N), L.bf_layout_padata(N)
    rec = np.zeros(lm.size, np.uint8)
    rec[lm.off[0]:lm.off[0] + 4] = np.array([off], np.int32).view(np.uint8)
    rec[lm.off[1]:lm.off[1] + M * N * 4] = np.ascontiguousarray(g["signals"], np.float32).view(np.uint8).ravel()
    rec[lm.off[2]:lm.off[2] + n * 4] = mics.view(np.uint8)
    rec[lm.off[3]:lm.off[3] + 4] = np.array([n], np.int32).view(np.uint8)
    pa = np.zeros(lp.size, np.uint8)
    nat.check(L.bf_miso_record_listen(nat.ptr(rec), nat.ptr(pa)))
    assert pa[:4].view(np.int32)[0] == 1
    got = pa[lp.off[1]:].view(np.float32)
    ref = cpu.miso_scale(cpu.miso_pad(g["signals"], mics, whole, off), n, 128.0)
    assert bits_equal(got, ref)


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("algo_name", ["pad", "lerp"])
def test_host_batch_replay_equals_per_buffer_calls(algo_name, pinned):
    """bf_mimo_host_batch (chunked, copies overlapped with the kernel on three streams) gives the
    same maps as one mimo_pad/mimo_lerp call per buffer; pageable and pinned host memory."""
    config, nat, L = _setup("c1")
    torch = _torch()
    from lib import directions
    g = gold("c1")
    mics = nat.i32(g["mic_ids"])
    D, n, N, M = 400, 64, 256, 64
    whole, d32 = directions.whole_and_f32()
    L.load_coefficients_pad(nat.ptr(whole), whole.size)
    L.load_coefficients_lerp(nat.ptr(d32), d32.size)
    nat.check()
    F = 37                                            # chunk schedule 4 + 29 + 4 (small first and last chunk, ragged)
    rng = np.random.default_rng(21)
    frames = rng.standard_normal((F, M, N)).astype(np.float32)
    frames[3] = g["signals"]
    algo = nat.ALGO_PAD if algo_name == "pad" else nat.ALGO_LERP
    name = "mimo_pad" if algo_name == "pad" else "mimo_lerp"
    single = np.zeros((F, D), np.float32)
    for f in range(F):
        getattr(L, name)(nat.ptr(frames[f]), nat.ptr(single[f]), nat.ptr(mics), n)
        nat.check()
    assert bits_equal(single[3], g["img_pad" if algo_name == "pad" else "img_lerp"])
    if pinned:
        h_in = torch.from_numpy(frames).pin_memory()
        h_out = torch.zeros((F, D), dtype=torch.float32).pin_memory()
        nat.check(L.bf_mimo_host_batch(algo, h_in.data_ptr(), h_out.data_ptr(), F, nat.ptr(mics), n))
        got = h_out.numpy()
    else:
        got = np.full((F, D), np.nan, np.float32)
        nat.check(L.bf_mimo_host_batch(algo, nat.ptr(frames), nat.ptr(got), F, nat.ptr(mics), n))
    assert bits_equal(got, single)


def test_host_batch_chunk_schedule_any_frame_count():
    """Every frame count goes through the chunk schedule of bf_mimo_host_batch (small first chunk, then up to
    32 frames) and still equals one device launch over the same frames."""
    config, nat, L = _setup("c1")
    torch = _torch()
    from lib import directions
    mics = nat.i32(gold("c1")["mic_ids"])
    D, n, N, M = 400, 64, 256, 64
    whole, _ = directions.whole_and_f32()
    L.load_coefficients_pad(nat.ptr(whole), whole.size)
    nat.check()
    rng = np.random.default_rng(5)
    frames = rng.standard_normal((131, M, N)).astype(np.float32)
    d_sig, d_mics = torch.from_numpy(frames).cuda(), torch.from_numpy(mics).cuda()
    d_ref = torch.zeros((131, D), device="cuda")
    nat.check(L.bf_mimo_dev(nat.ALGO_PAD, d_sig.data_ptr(), d_ref.data_ptr(), 131, d_mics.data_ptr(), n, 0, D, None))
    torch.cuda.synchronize()
    ref = d_ref.cpu().numpy()
    for F in (1, 2, 3, 4, 5, 8, 31, 32, 33, 47, 48, 49, 64, 65, 97, 131):
        got = np.full((F, D), np.nan, np.float32)
        nat.check(L.bf_mimo_host_batch(nat.ALGO_PAD, nat.ptr(frames), nat.ptr(got), F, nat.ptr(mics), n))
        assert bits_equal(got, ref[:F]), F


@pytest.mark.parametrize("quirk", [1, 0])
def test_wire_format_ingest_vs_oracle(quirk):
    """datagram payloads -> [mic][sample] float buffer on the device, bit-exact
    against the oracle's restatement of receiver.c:94-151 (with and without the reference's
    odd-row off-by-one), plus the get_data() channel mask; then straight into the beamformer."""
    from oracle import cpu
    torch = _torch()
    config, nat, L = _setup("default")
    M, N, frames = 256, 256, 3
    rng = np.random.default_rng(31 + quirk)
    payload = rng.integers(-(1 << 23), 1 << 23, (frames, N, M)).astype(np.int32)
    d_in = torch.from_numpy(payload).cuda()
    d_out = torch.full((frames, M, N), float("nan"), device="cuda")
    mask = np.zeros(256, np.uint8)
    mask[[0, 1, 4, 63, 200]] = 1
    d_mask = torch.from_numpy(mask).cuda()
    nat.check(L.bf_ingest_dev(d_in.data_ptr(), d_out.data_ptr(), frames, 4, 8, 8, 16777216.0, quirk, None, None))
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    for f in range(frames):
        assert bits_equal(got[f], cpu.ingest(payload[f], 4, quirk=bool(quirk)))
    nat.check(L.bf_ingest_dev(d_in.data_ptr(), d_out.data_ptr(), frames, 4, 8, 8, 16777216.0, quirk, d_mask.data_ptr(), None))
    torch.cuda.synchronize()
    got2 = d_out.cpu().numpy()
    assert not got2[:, mask == 1].any() and bits_equal(got2[:, mask == 0], got[:, mask == 0])
    # three arrays only (ACTIVE_ARRAYS = 3 in the stock config): channels 192.. stay untouched
    d_out3 = torch.full((1, M, N), 7.0, device="cuda")
    nat.check(L.bf_ingest_dev(d_in.data_ptr(), d_out3.data_ptr(), 1, 3, 8, 8, 16777216.0, quirk, None, None))
    torch.cuda.synchronize()
    o3 = d_out3.cpu().numpy()[0]
    assert np.all(o3[192:] == 7.0) and bits_equal(o3[:192], cpu.ingest(payload[0], 3, quirk=bool(quirk))[:192])
    assert L.bf_ingest_dev(d_in.data_ptr(), d_out.data_ptr(), 1, 4, 8, 8, 1000.0, 1, None, None) != 0   # not 2^k


def test_streaming_replay_of_a_wire_format_recording():
    """BASELINE config C5 as a stream (lib.replay.stream_video): a stored recording of datagram payloads in pinned
    host memory -> chunked double-buffered H2D -> windowed wire-format conversion (bf_ingest_windows_dev) ->
    power maps -> overlay.  At C1 size against the oracle chain: every 30 fps frame == orc_ingest of its
    256-datagram window (reference quirk included) followed by the oracle's mimo_pad, bit for bit, with chunk
    boundaries that cut between frames, two passes, and frames sharded 2 ways."""
    import ctypes
    from oracle import cpu
    config, nat, L = _setup("c1")
    torch = _torch()
    from lib import directions, replay
    g = gold("c1")
    mics = nat.i32(g["mic_ids"])
    D, n, N, M = 400, 64, 256, 64
    whole, _ = directions.whole_and_f32()
    L.load_coefficients_pad(nat.ptr(whole), whole.size)
    nat.check()
    rng = np.random.default_rng(77)
    total = 20000                                                    # 13 frames at 30 fps
    stream = rng.integers(-(1 << 23), 1 << 23, (total, M)).astype(np.int32)
    h_stream = torch.from_numpy(stream).pin_memory()
    d_mics = torch.from_numpy(mics).cuda()
    starts = replay.frame_starts(replay.n_frames_in(total))
    want = []
    for s0 in starts:
        sig = np.zeros((M, N), np.float32)
        win = np.ascontiguousarray(stream[s0:s0 + N])
        cpu.lib().orc_ingest(win.ctypes.data_as(ctypes.c_void_p), sig.ctypes.data_as(ctypes.c_void_p), N, M, 1, 8, 8,
                             ctypes.c_double(2.0 ** 24), 1)
        want.append(cpu.mimo_pad(sig, mics, whole, D))
    want = np.stack(want)
    res = replay.stream_video(h_stream, 1, nat.ALGO_PAD, d_mics, n, chunk_frames=5, passes=2, keep_maps=True,
                              window=(64, 36))
    assert res["frames"] == 2 * len(starts) and res["recording_frames"] == len(starts)
    assert bits_equal(res["maps"].cpu().numpy(), want)
    # every chunk copies the datagrams from its first window to the end of its last one, once per pass
    per_pass = sum(int(starts[min(a + 5, len(starts)) - 1] + N - starts[a]) for a in range(0, len(starts), 5))
    assert res["h2d_bytes"] == 2 * per_pass * M * 4 and res["confidence"].shape[0] == 2 * len(starts)
    assert torch.equal(res["confidence"][:len(starts)], res["confidence"][len(starts):])          # same content twice
    odd = replay.stream_video(h_stream, 1, nat.ALGO_PAD, d_mics, n, chunk_frames=4, first_frame=1, frame_step=2,
                              keep_maps=True, overlay=False)
    assert bits_equal(odd["maps"].cpu().numpy(), want[1::2])


def test_overlapping_gather_steps_on_one_gpu():
    """bf_gather_overlap: back-to-back steps of the fused gather kernel launched with programmatic stream
    serialisation (step i + 1 takes over SMs while step i runs its last tiles), world = 1 with the step flags on, a
    ring of 3 buffers, several frame counts (so that the last round of tiles is ragged): every step's maps == the
    plain launch, bit for bit; the setting is sticky until switched off and does not leak into other launches."""
    import ctypes
    config, nat, L = _setup("c1")
    torch = _torch()
    from lib import directions
    g = gold("c1")
    mics = nat.i32(g["mic_ids"])
    D, n, N, M = 400, 64, 256, 64
    whole, _ = directions.whole_and_f32()
    L.load_coefficients_pad(nat.ptr(whole), whole.size)
    nat.check()
    d_mics = torch.from_numpy(mics).cuda()
    vp = ctypes.c_void_p
    st = torch.cuda.current_stream().cuda_stream
    seq = 0
    flags = torch.zeros(8, dtype=torch.int64, device="cuda")
    timed_out = torch.zeros(1, dtype=torch.int32, device="cuda")
    fl = (vp * 1)(vp(flags.data_ptr()))
    for F, d0, dc, steps, depth in ((37, 0, 400, 9, 3), (150, 96, 250, 6, 3), (3, 8, 392, 12, 4)):
        gen = torch.Generator(device="cuda").manual_seed(F)
        sig = 0.1 * torch.randn((steps, F, M, N), generator=gen, device="cuda")
        ring = [torch.full((1, F, dc), float("nan"), device="cuda") for _ in range(depth)]
        got = []
        torch.cuda.synchronize()
        assert L.bf_gather_overlap(1) == 0
        for i in range(steps):
            seq += 1
            bufs = (vp * 1)(vp(ring[i % depth].data_ptr()))
            nat.check(L.bf_mimo_dev_gather_sync(nat.ALGO_PAD, sig[i].data_ptr(), F, d_mics.data_ptr(), n, d0, dc, 0, 1,
                                                bufs, dc, fl, max(0, seq - depth + 1), seq, timed_out.data_ptr(), st))
            if i >= depth - 1:                      # the buffer about to be reused: copy it out (ordinary stream order)
                got.append(ring[(i + 1) % depth][0].clone())
        assert L.bf_gather_overlap(0) == 1 and L.bf_gather_overlap(-1) == 0
        torch.cuda.synchronize()
        assert int(timed_out) == 0 and int(flags[0]) == seq
        # steps whose buffer was not reused are still in the ring
        first_kept = steps - (depth - 1)
        want = torch.zeros((steps, F, D), device="cuda")
        for i in range(steps):
            nat.check(L.bf_mimo_dev(nat.ALGO_PAD, sig[i].data_ptr(), want[i].data_ptr(), F, d_mics.data_ptr(), n, 0, D, None))
        torch.cuda.synchronize()
        for j, m in enumerate(got):                 # got[j] = step j (copied out before step j + depth reused its buffer)
            assert torch.equal(m, want[j][:, d0:d0 + dc]), (F, j)
        for i in range(max(first_kept, len(got)), steps):
            assert torch.equal(ring[i % depth][0], want[i][:, d0:d0 + dc]), (F, i)


def test_batch_replay_windows_and_maps():
    """BASELINE config C5 mechanics at C1 size: 30 fps windows at floor(k*fs/30) of a channel-major
    recording -> gather kernel -> batched maps == one mimo_pad call per NumPy-sliced window."""
    config, nat, L = _setup("c1")
    torch = _torch()
    from lib import directions, replay
    g = gold("c1")
    mics = nat.i32(g["mic_ids"])
    D, n, N, M = 400, 64, 256, 64
    whole, _ = directions.whole_and_f32()
    L.load_coefficients_pad(nat.ptr(whole), whole.size)
    nat.check()
    starts = replay.frame_starts(5)
    assert list(starts) == [0, 1627, 3255, 4882, 6510]          # floor(k*48828/30)
    rng = np.random.default_rng(8)
    rec = rng.standard_normal((M, 20000)).astype(np.float32)
    assert replay.n_frames_in(20000) == 13
    d_rec, d_mics = torch.from_numpy(rec).cuda(), torch.from_numpy(mics).cuda()
    got = {}
    for rank in range(2):                                         # frames shard over ranks, no collective
        idx, maps = replay.replay_dev(nat.ALGO_PAD, d_rec, d_mics, n, chunk=4, rank=rank, world=2)
        torch.cuda.synchronize()
        for k, mp in zip(idx, maps.cpu().numpy()):
            got[int(k)] = mp
    assert sorted(got) == list(range(13))
    for k in range(13):
        s = (k * 48828) // 30
        ref = np.zeros(D, np.float32)
        win = np.ascontiguousarray(rec[:, s:s + N])
        L.mimo_pad(nat.ptr(win), nat.ptr(ref), nat.ptr(mics), n)
        nat.check()
        assert bits_equal(got[k], ref), k


def test_capture_file_to_power_maps(tmp_path):
    """a pcap of the UDP stream -> device ingest -> power maps, bit-exact against the
    oracle chain (receiver restatement + mimo_pad) on the same datagrams."""
    import struct
    from oracle import cpu
    from lib import capture, replay
    from test_tracking_and_capture import _datagram, _udp_frame, _write_pcap
    torch = _torch()
    config, nat, L = _setup("default")
    from lib import directions
    M, N = 256, 256
    rng = np.random.default_rng(77)
    streams = rng.integers(-(1 << 20), 1 << 20, (2 * N + 17, M)).astype(np.int32)
    frames = [_udp_frame(_datagram(i, s)) for i, s in enumerate(streams)]
    path = str(tmp_path / "udp_capture.pcap")
    _write_pcap(path, frames, 1.0 + np.arange(len(frames)) / 48828.0)
    cap = capture.read_capture(path, n_microphones=M)
    assert cap.dropped == 0 and cap.stream.shape == streams.shape
    d_sig = replay.signals_from_capture(cap)
    assert tuple(d_sig.shape) == (2, M, N)
    mics, n = directions.active_microphones()
    mics = nat.i32(mics)
    whole = nat.i32(directions.calculate_delays().astype(int)).ravel()
    L.load_coefficients_pad(nat.ptr(whole), whole.size)
    nat.check()
    D = config.MAX_RES_X * config.MAX_RES_Y
    d_maps = torch.zeros((2, D), device="cuda")
    d_mics = torch.from_numpy(mics).cuda()
    nat.check(L.bf_mimo_dev(0, d_sig.data_ptr(), d_maps.data_ptr(), 2, d_mics.data_ptr(), n, 0, D, None))
    torch.cuda.synchronize()
    got_sig, got = d_sig.cpu().numpy(), d_maps.cpu().numpy()
    for b in range(2):
        want_sig = cpu.ingest(streams[b * N:(b + 1) * N], 4, quirk=True)
        assert bits_equal(got_sig[b], want_sig)
        assert bits_equal(got[b], cpu.mimo_pad(want_sig, mics, whole, D))
