#!/usr/bin/env python3
"""bench.py -- beam power maps/s (and DAS GMAC/s) of the time-domain delay-and-sum hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--algo pad|lerp] [--frames F] [--workload c3|c1|stock]

Workload (config.workload): BASELINE config C3 -- 4 arrays / 256 microphones, 256-sample
buffers, 180x180 direction grid (D = 32 400), time-domain DAS power map, synthetic 3-source
input.  C3 is the configuration BASELINE.json's metric "beam power maps/s" is
quoted on and it fits one GPU; configs[1] (MISO) is measured as the extra "miso" object.

A *step* is one pass of the hot path over one batch of F frames (F maps).  Inputs are resident
in HBM before the timed region; every step reads a different batch out of a pool larger than
the 126 MB L2 ("inputs larger than L2").  With N > 1 ranks (torchrun) the direction grid is
sharded: each rank computes D/N directions of all F maps and ONE in-place NCCL all-gather per
step assembles the direction-major maps on every rank -> total work fixed -> "strong".

value      whole-job maps/s, device-timed (CUDA events, max over ranks)
e2e        maps/s through the reference-facing C-ABI call mimo_pad()/mimo_lerp() with HOST
           buffers (H2D + D2H inside the timed region), each rank serving its own frames
roofline   dominant kernel (das_mimo_kernel) vs the measured HBM copy bandwidth
cpu_baseline  the reference's own C (oracle/_ref) or the oracle port on the host cores
--impl reference   the reference CPU implementation alone, same metric/config
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "zybo-rt-sampler-image-detection_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    "c3": dict(N_MICROPHONES=256, N_SAMPLES=256, MAX_RES_X=180, MAX_RES_Y=180, GEOMETRY_N_MICS=256,
               GEOMETRY_N_ARRAYS=4, ref_cfg="c3",
               desc="C3: 256 mics (4 arrays), 256-sample buffers @48.828 kHz, 180x180 grid, TD-DAS power map"),
    "c1": dict(N_MICROPHONES=64, N_SAMPLES=256, MAX_RES_X=20, MAX_RES_Y=20, GEOMETRY_N_MICS=64,
               GEOMETRY_N_ARRAYS=1, ref_cfg="c1",
               desc="C1: 64 mics (8x8), 256-sample buffers, 20x20 grid, TD-DAS power map"),
    "stock": dict(N_MICROPHONES=256, N_SAMPLES=256, MAX_RES_X=57, MAX_RES_Y=32, GEOMETRY_N_MICS=256,
                  GEOMETRY_N_ARRAYS=4, ref_cfg="default",
                  desc="stock config.json: 256 mics, 256 samples, 57x32 grid, TD-DAS power map"),
}


def _source_sha(files):
    """Short hash of the kernel sources a profile entry belongs to."""
    import hashlib
    h = hashlib.sha256()
    for f in files:
        with open(os.path.join(PKG, "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:12]


def _traffic(key):
    """Per-launch DRAM traffic (or another ncu-derived figure) recorded in profiles/traffic.json.
    Every entry names the kernel sources it was captured from and their hash; an entry whose sources have
    changed since the capture is STALE and is not reported (None) rather than printed as if it were current."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    e = json.load(open(path)).get(key)
    if not isinstance(e, dict):
        return None
    try:
        if _source_sha(e["sources"]) != e["source_sha"]:
            return None
    except (KeyError, OSError):
        return None
    return e["value"]


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------
# CPU side: the reference's own C (oracle/_ref) or the oracle port, one process per core
# ----------------------------------------------------------------------------------------
_BARRIER = None          # inherited by the forked workers (cannot be pickled)


def _cpu_worker(args):
    kind, ref_cfg, algo, sig, mics, table, D, maps, lib = args
    barrier = _BARRIER
    if kind == "reference":
        from oracle import ref
        R = ref.RefC(ref_cfg, lib)
        import ctypes
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        img = np.zeros(R.D, np.float32)
        if algo == "pad":
            R.L.load_coefficients_pad(p(table), ctypes.c_int(table.size))
            fn = R.L.mimo_pad
        else:
            R.L.load_coefficients_lerp(p(table), ctypes.c_int(table.size))
            fn = R.L.mimo_lerp
        call = lambda: fn(p(sig), p(img), p(mics), ctypes.c_int(len(mics)))
    else:
        from oracle import cpu
        if algo == "pad":
            call = lambda: cpu.mimo_pad(sig, mics, table, D)
        else:
            call = lambda: cpu.mimo_lerp(sig, mics, table, D)
    call()                                   # warm-up map (page-in, table load)
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(maps):
        call()
    return time.perf_counter() - t0


def cpu_setup(wl_name, algo):
    """Inputs of the CPU run, from the oracle's NumPy restatement of the delay generator."""
    from oracle import directions_np as dn, ref
    from lib import synthetic
    wl = WORKLOADS[wl_name]
    cfg = dn.cfg_with(N_MICROPHONES=wl["N_MICROPHONES"], MAX_RES_X=wl["MAX_RES_X"],
                      MAX_RES_Y=wl["MAX_RES_Y"], n_mics_geom=wl["GEOMETRY_N_MICS"],
                      n_arrays_geom=wl["GEOMETRY_N_ARRAYS"])
    mics, n = dn.active_microphones(cfg)
    delays = dn.calculate_delays(cfg).reshape(-1, n)
    D = delays.shape[0]
    src = synthetic.C3 if wl_name == "c3" else synthetic.C1
    sources = [s for s in src["sources"] if s[0] < D]
    sig = synthetic.point_sources(delays, mics, wl["N_MICROPHONES"], wl["N_SAMPLES"], 48828.0,
                                  sources, src["noise"], src["seed"])
    table = delays.astype(int).astype(np.int32) if algo == "pad" else np.float32(delays)
    kind = "reference" if ref.available(wl["ref_cfg"]) else "port"
    return kind, wl["ref_cfg"], sig, np.ascontiguousarray(mics, np.int32), np.ascontiguousarray(table), D, n


_REF_HANDLES = []        # the reference's shared objects, opened in the PARENT too (the workers are forked from it)


def reference_builds(ref_cfg, kind):
    """Builds of the reference's C available on this host: always the portable one (-mavx2 -mfma, what the
    container that has /root/reference compiled), plus the x86-64-v4 one where the CPU has AVX-512 -- the
    reference's own flags say -march=native, so on such a host that is what its authors would get."""
    if kind != "reference":
        return ["port"]
    import ctypes
    from oracle import ref
    libs = ["libref.so"] + (["libref_v4.so"] if ref.v4_available(ref_cfg) else [])
    for lib in libs:
        _REF_HANDLES.append(ctypes.CDLL(os.path.join(ref.ROOT, ref_cfg, lib)))
    return libs


def cpu_run(wl_name, algo, maps_per_core, cores=None, setup=None, lib="libref.so"):
    """maps/s of the CPU implementation with `cores` processes, each producing whole maps
    (the reference is single-threaded with process-global tables)."""
    kind, ref_cfg, sig, mics, table, D, n = setup or cpu_setup(wl_name, algo)
    cores = cores or os.cpu_count() or 1
    global _BARRIER
    ctx = mp.get_context("fork")
    _BARRIER = ctx.Barrier(cores)
    with ctx.Pool(cores) as pool:
        times = pool.map(_cpu_worker, [(kind, ref_cfg, algo, sig, mics, table, D, maps_per_core, lib)] * cores,
                         chunksize=1)
    wall = max(times)
    return dict(maps_per_s=cores * maps_per_core / wall, seconds=wall, cores=cores, kind=kind,
                maps=cores * maps_per_core, D=D, n=n, lib=lib)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = WORKLOADS[args.workload]
    setup = cpu_setup(args.workload, args.algo)
    D, n = setup[5], setup[6]
    N = wl["N_SAMPLES"]
    cores = os.cpu_count() or 1
    libs = reference_builds(setup[1], setup[0])
    # one step = one whole map per core (bounded sample of the F-frame step of the GPU arm); every available
    # build of the reference is timed, the line's value is the fastest
    builds, total = {}, None
    for lib in libs:
        t = 0.0
        for i in range(args.warmup + args.steps):
            r = cpu_run(args.workload, args.algo, 1, cores, setup, lib)
            if i >= args.warmup:
                t += r["seconds"]
        builds[lib] = cores * args.steps / t
        if total is None or t < total:
            total = t
    kind = r["kind"]
    maps = cores * args.steps
    value = maps / total
    line = {
        "impl": "reference", "metric": "beam power maps/s", "value": value, "unit": "maps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "gmac_per_s": value * D * n * N / 1e9,
        "config": {"workload": wl["desc"], "algo": args.algo, "directions": D, "mics": n, "samples": N},
        "cpu_baseline": {"value": value, "unit": "maps/s", "cores": cores, "kind": kind,
                         "sample": "%d whole maps per step (one per host core), %d steps" % (cores, args.steps),
                         "builds_maps_per_s": builds,
                         "builds_note": "libref.so = the reference's C with -O3 -mavx2 -mfma; libref_v4.so = -march=x86-64-v4 "
                                        "(what its -march=native gives on an AVX-512 host), timed only where the CPU has it"},
        "e2e": {"value": value, "unit": "maps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------
# clocks during the timed region
# ----------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}



def mvdr_c4(nat, L, torch, stream, bins=512, dirs=32768, K=64):
    """BASELINE config C4: frequency-domain MVDR, 256 mics, 1024-point FFT, bins 1..512, K = 64
    snapshots, 256 x 128 directions.  Useful flops of the steering contraction: 8*D*M^2*F
    (SURVEY.md 8d).  Tensor roofline = useful flops / steering-kernel time against the measured
    dense bf16 peak; the kernel issues 1.875x that as kind::f16 MMAs (3-pass two-term fp16 split, 20 of 32
    (k-chunk, row-quarter) items thanks to the triangular L^-1)."""
    import realtime_scripts.calc_r_prime as rp
    import realtime_scripts.config as cfg
    M, N, F = 256, 1024, bins
    res_x, res_y = 256, dirs // 256
    D = res_x * res_y
    pos_all, _ = rp.calc_r_prime(cfg.ELEMENT_DISTANCE)
    x_max = np.tan(np.deg2rad(cfg.VIEW_ANGLE / 2))
    xs = np.linspace(-x_max, x_max, res_x)
    ys = np.linspace(-x_max / cfg.ASPECT_RATIO, x_max / cfg.ASPECT_RATIO, res_y)
    act = np.arange(M, dtype=np.int32)
    p = nat.ptr
    mx, my = np.ascontiguousarray(pos_all[0]), np.ascontiguousarray(pos_all[1])
    nat.check(L.bf_fd_setup(M, N, 48828.0, 343.0, 1, 1 + F, p(xs), res_x, p(ys), res_y, 1.0, p(mx), p(my), p(act), M))
    gen = torch.Generator(device="cuda").manual_seed(1237)
    snaps = 0.05 * torch.randn((K, M, N), generator=gen, device="cuda")
    t = torch.arange(N, device="cuda")[None, None, :]
    m = torch.arange(M, device="cuda")[None, :, None]
    for f0, amp, slope in ((2000.0, 0.3, 0.011), (5000.0, 0.2, -0.023), (9000.0, 0.1, 0.005)):
        ph = 2 * np.pi * torch.rand((K, 1, 1), generator=gen, device="cuda")
        snaps += amp * torch.sin(2 * np.pi * f0 * (t + slope * m * 48.828) / 48828.0 + ph)
    snaps = snaps.float().contiguous()
    power = torch.zeros(D, device="cuda")
    times, stage = [], np.zeros(5, np.float32)
    for r in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        nat.check(L.bf_fd_mvdr_dev(snaps.data_ptr(), power.data_ptr(), K, 1e-2, stream))
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    L.bf_fd_mvdr_timings(nat.ptr(stage))
    ms = float(np.mean(times[1:]))
    useful = 8.0 * D * M * M * F
    peak_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peak_path))["bf16_tflops"]) if os.path.exists(peak_path) else 1590.0
    ach = useful / 1e12 / (float(stage[4]) * 1e-3)
    return {"workload": "C4: FD-MVDR, 256 mics, 1024-pt FFT, %d bins, K=%d snapshots, %d directions, loading 1e-2" % (F, K, D),
            "maps_per_s": 1e3 / ms, "ms_per_map": ms, "finite": bool(torch.isfinite(power).all()),
            "stage_ms": dict(zip(["fft_f64", "covariance_f64", "cholesky_f64", "tri_inverse_f64", "steering_tcgen05"],
                                 [float(x) for x in stage])),
            "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                         "kernel": "mvdr_tc_steer_kernel3<QUARTER> (tcgen05 kind::f16, 3-pass two-term fp16 split, N = 128)",
                         "kernel_ms": float(stage[4]), "useful_flops_per_launch": useful,
                         "issued_over_useful": 1.875, "traffic": None,
                         "tensor_pipe_active_pct_ncu": _traffic("mvdr_tc_steer_tensor_pipe_active_pct"),
                         "peak_source": "measured dense bf16 burst (MEASURED_PEAKS.json)"}}


def mvdr_c4_sharded(torch, dist, nat, L, rank, world, bins=512, dirs=32768, K=64):
    """BASELINE config C4 with the DIRECTIONS sharded over the ranks (FD path): every rank computes the
    spectra, covariance, factor and inverse in full and steers dirs / world directions on the tcgen05 kernel;
    the slices are exchanged with NVLink peer stores (lib.sharded.PeerGather, no NCCL on the data path).
    Whole-job maps/s = 1 / (slowest rank's time per map, device-timed)."""
    import realtime_scripts.calc_r_prime as rp
    import realtime_scripts.config as cfg
    from lib.sharded import PeerGather, fd_mvdr_sharded, fd_mvdr_sharded_bins
    M, N, F = 256, 1024, bins
    res_x, res_y = 256, dirs // 256
    D = res_x * res_y
    pos_all, _ = rp.calc_r_prime(cfg.ELEMENT_DISTANCE)
    x_max = np.tan(np.deg2rad(cfg.VIEW_ANGLE / 2))
    xs = np.linspace(-x_max, x_max, res_x)
    ys = np.linspace(-x_max / cfg.ASPECT_RATIO, x_max / cfg.ASPECT_RATIO, res_y)
    act = np.arange(M, dtype=np.int32)
    p = nat.ptr
    mx, my = np.ascontiguousarray(pos_all[0]), np.ascontiguousarray(pos_all[1])
    nat.check(L.bf_fd_setup(M, N, 48828.0, 343.0, 1, 1 + F, p(xs), res_x, p(ys), res_y, 1.0, p(mx), p(my), p(act), M))
    gen = torch.Generator(device="cuda").manual_seed(1237)          # same seed on every rank: same snapshots
    snaps = 0.05 * torch.randn((K, M, N), generator=gen, device="cuda")
    t = torch.arange(N, device="cuda")[None, None, :]
    m = torch.arange(M, device="cuda")[None, :, None]
    for f0, amp, slope in ((2000.0, 0.3, 0.011), (5000.0, 0.2, -0.023), (9000.0, 0.1, 0.005)):
        snaps += amp * torch.sin(2 * np.pi * f0 * (t + slope * m * 48.828) / 48828.0)
    snaps = snaps.float().contiguous()
    peer = PeerGather(D, 1, rank, world, dist, depth=2)
    times = []
    for i in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        a.record()
        if F % world == 0:
            fd_mvdr_sharded_bins(peer, i, snaps, K, 1e-2, F, dist)      # float64 stages by bins, steering by directions
        else:
            fd_mvdr_sharded(peer, i, snaps, K, 1e-2)
        peer.ready(i)
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    peer.check()
    power = peer.maps(3).reshape(-1)
    finite = bool(torch.isfinite(power).all())
    tt = torch.tensor([float(np.mean(times[1:]))], device="cuda", dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dist.barrier()
    peer.close()
    ms = float(tt[0])
    return {"workload": "C4 sharded: FD-MVDR, 256 mics, 1024-pt FFT, %d bins, K=%d, %d directions over %d GPUs "
                        "(%d each), slices exchanged by NVLink peer stores%s"
                        % (F, K, D, world, (D + world - 1) // world,
                           "; float64 stages sharded by bins, operand images all-gathered (NCCL)" if F % world == 0 else ""),
            "maps_per_s": 1e3 / ms, "ms_per_map": ms, "finite": finite, "n_gpus": world}


def replay_c5_sharded_local(torch, nat, algo, d_mics, n, D, M, N, rank):
    """C5 across GPUs, this rank's part: its own 20 s recording (recording id = rank, seed 1238 + id) replayed
    entirely on its GPU (recordings shard with no collective).  Returns (frames, ms); no collective in here."""
    from lib import replay
    seconds = 20
    samples = seconds * 48828
    gen = torch.Generator(device="cuda").manual_seed(1238 + rank)
    rec = torch.empty((M, samples), device="cuda")
    for i in range(0, M, 32):
        rec[i:i + 32].normal_(generator=gen)
    rec *= 0.05
    total = replay.n_frames_in(samples)
    maps = torch.empty((total, D), device="cuda")
    replay.replay_dev(algo, rec, d_mics, n, out=maps)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    replay.replay_dev(algo, rec, d_mics, n, out=maps)
    b.record()
    torch.cuda.synchronize()
    return total, a.elapsed_time(b)


def replay_c5_stream(torch, nat, algo, d_mics, n, M, rank, minutes=None):
    """BASELINE config C5 for real: one HOUR of the 256-channel stream as the Zybo sends it (int32 datagram
    payloads, 1 KiB per sample instant = 180 GB) goes host -> device through pinned double-buffered copies and on
    through wire-format conversion -> 30 fps windows -> C3 power maps -> 640x360 overlay + peak + confidence
    (lib.replay.stream_video), H2D inside the timing.  Host memory holds `minutes` of recording (BF_C5_MINUTES,
    default 4 = 11.7 GB pinned; seed 1238 + recording id = rank) and is replayed 60 / minutes times, so the
    bytes crossing PCIe are the full hour's; nothing but two chunks (2 x 0.43 GB) is resident on the GPU."""
    from lib import replay
    if minutes is None:
        minutes = float(os.environ.get("BF_C5_MINUTES", "4"))
    if minutes <= 0:
        return None
    fs = 48828
    total = int(minutes * 60 * fs)
    h = torch.empty((total, M), dtype=torch.int32, pin_memory=True)
    gen = torch.Generator(device="cuda").manual_seed(1238 + rank)
    piece = 1 << 20
    tone = torch.arange(M, device="cuda", dtype=torch.float32)[None, :] * 0.37
    for off in range(0, total, piece):
        c = min(piece, total - off)
        t = torch.arange(off, off + c, device="cuda", dtype=torch.float32)[:, None]
        x = 0.02 * torch.randn((c, M), generator=gen, device="cuda") + 0.05 * torch.sin(2 * np.pi * 3000.0 / fs * t + tone)
        h[off:off + c].copy_((x * 8388608.0).to(torch.int32))
    torch.cuda.synchronize()
    stream_minutes = float(os.environ.get("BF_C5_STREAM_MINUTES", "60"))     # tests shorten the hour
    passes = max(1, int(round(stream_minutes / minutes)))
    warm = replay.stream_video(h[:10 * fs], 4, algo, d_mics, n, chunk_frames=256)           # warm-up: 10 s of stream
    res = replay.stream_video(h, 4, algo, d_mics, n, chunk_frames=256, passes=passes)
    info = res["info"].numpy().view(nat.HEAT_INFO_DTYPE)
    out = {"frames": res["frames"], "seconds": res["seconds"], "frames_per_s": res["frames_per_s"],
           "x_realtime_at_30fps": res["frames_per_s"] / 30.0, "stream_seconds": res["frames"] / 30.0,
           "host_recording_minutes": minutes, "passes": passes, "h2d_gb": res["h2d_bytes"] / 1e9,
           "pcie_h2d_gb_per_s": res["h2d_gb_per_s"],
           "stage_ms": res["stage_ms"], "stage_note": "device time per stage summed over chunks; h2d runs on the copy "
                                                      "stream under the compute of the previous chunk",
           "overlay_fraction": float(np.mean(info["overlay"] != 0)), "mean_confidence": float(res["confidence"].mean()),
           "warmup_frames": warm["frames"]}
    del h
    return out


def replay_c5(args, nat, torch, algo, d_mics, n, D, M, N):
    """BASELINE config C5 on a bounded sample: a 20 s slice of a 256-channel recording resident in
    HBM (1.0 GB, channel-major like PC/record.py's .npy), power-map video at 30 fps: window gather
    (floor(k*fs/30)) + C3 power maps, 600 frames.  Frames are independent -> recordings / frames
    shard over GPUs with no collective; the 1 h recording of C5 is 180 such slices."""
    from lib import replay
    seconds = 20
    samples = seconds * 48828
    gen = torch.Generator(device="cuda").manual_seed(1238)
    rec = torch.empty((M, samples), device="cuda")
    for i in range(0, M, 32):
        rec[i:i + 32].normal_(generator=gen)
    rec *= 0.05
    total = replay.n_frames_in(samples)
    maps = torch.empty((total, D), device="cuda")
    replay.replay_dev(algo, rec, d_mics, n, out=maps)               # warm-up
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    replay.replay_dev(algo, rec, d_mics, n, out=maps)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    fps = total / (ms * 1e-3)
    # the same with the sensor-fusion overlay: heat map at the decider's 640x360 display size, peak and
    # entropy confidence per frame (lib.replay.video_dev), everything device-resident
    replay.video_dev(algo, rec, d_mics, n, keep_heat=False)
    torch.cuda.synchronize()
    a.record()
    out = replay.video_dev(algo, rec, d_mics, n, keep_heat=False)
    b.record()
    torch.cuda.synchronize()
    ms_v = a.elapsed_time(b)
    conf = float(out["confidence"].mean())
    del rec, out
    return {"workload": "C5 sample: %d s of a 256-mic recording (%.2f GB resident), 30 fps video, 180x180 maps, %d frames"
                        % (seconds, M * samples * 4 / 1e9, total),
            "frames_per_s": fps, "x_realtime_at_30fps": fps / 30.0, "ms": ms,
            "one_hour_recording_s_per_gpu": 108000 / fps,
            "with_overlay_640x360": {"frames_per_s": total / (ms_v * 1e-3), "ms": ms_v,
                                     "x_realtime_at_30fps": total / (ms_v * 1e-3) / 30.0,
                                     "mean_confidence": conf}}


def heatmap_c5(torch, d_maps, hbm_peak):
    """Post-processing of config C5's video on the device: a batch of 180x180 power maps
    resident in HBM -> log-scale colour index + jet LUT + peak centroid (one CTA per frame), cv2-exact
    bilinear resize to the sensor-fusion display (640x360) and to the application window (1920x1080),
    entropy confidence.  The resize is write-bound: out bytes / time against the HBM copy peak."""
    from lib import visual
    frames = d_maps.shape[0]
    out = {}
    for (W, H) in ((640, 360), (1920, 1080)):
        res = visual.heatmaps_dev(d_maps, window=(W, H), confidence=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        a.record()
        for _ in range(reps):
            visual.heatmaps_dev(d_maps, window=(W, H), confidence=True, out=res)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        out["%dx%d" % (W, H)] = {"frames_per_s": frames / (ms * 1e-3), "ms_per_batch": ms,
                                 "written_GB_per_s": frames * W * H * 3 / (ms * 1e-3) / 1e9,
                                 "frac_of_hbm_copy_peak": frames * W * H * 3 / (ms * 1e-3) / 1e9 / hbm_peak}
        del res
    out["workload"] = "%d maps of 180x180 (HBM-resident) -> heat overlay + peak + confidence per frame" % frames
    return out


def miso_c2(args, nat, L, config, directions, torch, stream, hbm_peak):
    """BASELINE config C2: MISO single-beam output, 64-mic 8x8 array, a continuous stream of
    256-sample blocks (2^16 blocks = 4.3 GB, HBM-resident, >> L2), pad and lerp delays, with the
    audio loop's /n*MIC_GAIN post-scale.  Per block: 64 rows x 1 KiB read + table row + 1 KiB
    written = 67 072 B for 16 384 MAC -> HBM-bound."""
    wl = WORKLOADS["c1"]
    config.reload(N_MICROPHONES=64, N_SAMPLES=256, MAX_RES_X=20, MAX_RES_Y=20, N_TAPS=8, SKIP_N_MICS=1,
                  GEOMETRY_N_MICS=64, GEOMETRY_N_ARRAYS=1)
    nat.configure_from(config)
    directions.load_pad_from_geometry()
    directions.load_lerp_from_geometry()
    mics, n = directions.active_microphones()
    d_mics = torch.from_numpy(nat.i32(mics)).cuda()
    M = N = 256
    M = 64
    blocks = 1 << 16
    gen = torch.Generator(device="cuda").manual_seed(1235)
    sig = torch.empty((blocks, M, N), device="cuda")
    for i in range(0, blocks, 8192):
        sig[i:i + 8192].normal_(generator=gen)
    out = torch.zeros((blocks, N), device="cuda")
    off = (14 * 20 + 6) * n
    res = {"workload": "C2: MISO steered beam, 64 mics, %d consecutive 256-sample blocks (%.1f GB resident, "
                       "inputs larger than L2), /n*MIC_GAIN post-scale" % (blocks, blocks * M * N * 4 / 1e9)}
    blk_bytes = n * N * 4 + 2 * n * 4 + N * 4
    for name, algo in (("pad", nat.ALGO_PAD), ("lerp", nat.ALGO_LERP)):
        ts = []
        for i in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            nat.check(L.bf_miso_dev(algo, sig.data_ptr(), out.data_ptr(), blocks, d_mics.data_ptr(), n, off, 1, stream))
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.mean(ts[3:]))
        gbs = blocks * blk_bytes / (ms * 1e-3) / 1e9
        res[name] = {"samples_per_s": blocks * N / (ms * 1e-3), "x_realtime": blocks * N / (ms * 1e-3) / 48828.0,
                     "gmac_per_s": blocks * n * N / (ms * 1e-3) / 1e9,
                     "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                  "frac": gbs / hbm_peak, "kernel": "miso_stream_kernel<%s>" % name,
                                  "kernel_ms": ms, "algorithmic_bytes_per_launch": blocks * blk_bytes,
                                  "traffic": _traffic("miso_stream_%s_%d" % (name, blocks))}}
    del sig, out
    return res

def _percentiles(ns):
    a = np.sort(np.asarray(ns, np.float64)) / 1e3
    return {"p50_us": float(a[len(a) // 2]), "p99_us": float(a[min(len(a) - 1, int(len(a) * 0.99))]),
            "min_us": float(a[0]), "max_us": float(a[-1]), "calls": int(len(a))}


def live_latency(nat, L, config, directions):
    """Per-call latency of the two LIVE surfaces of the reference, through the C ABI with pageable host buffers
    (what a maintainer gets by swapping the library under PC/src/main.pyx):
      mimo_pad(signals, image, adaptive_array, n)   one power map per 256-sample buffer -- the producer loops
                                                    (main.pyx:554-579) call it once per 5.24 ms buffer
      miso_steer_listen(out, adaptive_array, n, o)  one block of beam audio (api.c:491-543, miso_loop): get_data
                                                    from the registered source -> delay-and-sum -> host
    at the reference's stock configuration (256 mics, 57x32 grid) and at C3's 180x180 grid for mimo_pad.
    Wall clock per call (time.perf_counter_ns), after warm-up; budget = one buffer period = 5243 us."""
    from lib import beamformer, synthetic
    out = {"budget_us_per_buffer": 1e6 * 256 / 48828.0}
    for name, rx, ry in (("stock_57x32", 57, 32), ("c3_180x180", 180, 180)):
        config.reload(N_MICROPHONES=256, N_SAMPLES=256, MAX_RES_X=rx, MAX_RES_Y=ry, N_TAPS=8, SKIP_N_MICS=1,
                      GEOMETRY_N_MICS=256, GEOMETRY_N_ARRAYS=4)
        nat.configure_from(config)
        directions.load_pad_from_geometry()
        mics, n = directions.active_microphones()
        mics = nat.i32(mics)
        D = rx * ry
        sig = np.ascontiguousarray(synthetic.plot_py_stimulus(256, 256))
        img = np.zeros(D, np.float32)
        for _ in range(20):
            L.mimo_pad(nat.ptr(sig), nat.ptr(img), nat.ptr(mics), n)
        nat.check()
        ts = []
        for _ in range(300):
            t0 = time.perf_counter_ns()
            L.mimo_pad(nat.ptr(sig), nat.ptr(img), nat.ptr(mics), n)
            ts.append(time.perf_counter_ns() - t0)
        nat.check()
        out["mimo_pad_" + name] = _percentiles(ts)
        if name == "stock_57x32":
            beamformer.connect(False, verbose=False, source=beamformer.ArraySource(np.tile(sig, (1, 8))))
            try:
                beam = np.zeros(256, np.float32)
                off = (16 * 57 + 28) * n
                for _ in range(20):
                    L.miso_steer_listen(nat.ptr(beam), nat.ptr(mics), n, off)
                nat.check()
                ts = []
                for _ in range(500):
                    t0 = time.perf_counter_ns()
                    L.miso_steer_listen(nat.ptr(beam), nat.ptr(mics), n, off)
                    ts.append(time.perf_counter_ns() - t0)
                nat.check()
                out["miso_steer_listen_" + name] = _percentiles(ts)
            finally:
                beamformer.disconnect()
    return out


def fir_default(nat, L, config, torch, stream, clocks):
    """FIR (fractional-delay filter) power maps at the reference's stock configuration: 57x32 grid, 256
    mics, 256 samples, 8 taps (956 MFMA per map): mimo_convolve_naive's fused sequential
    chain and the AVX-ordered "vectorized" variant, 16 frames per launch, tiled kernel (das_fir.cu)."""
    config.reload(N_MICROPHONES=256, N_SAMPLES=256, MAX_RES_X=57, MAX_RES_Y=32, N_TAPS=8, SKIP_N_MICS=1,
                  GEOMETRY_N_MICS=256, GEOMETRY_N_ARRAYS=4)
    nat.configure_from(config)
    D, n, N, T, F = 57 * 32, 256, 256, 8, 16
    gen = torch.Generator(device="cuda").manual_seed(1239)
    taps = torch.randn((D, n, T), generator=gen, device="cuda") / T       # timing does not depend on the values
    sig = 0.1 * torch.randn((F, 256, N), generator=gen, device="cuda")
    nat.check(L.bf_load_table_dev(nat.ALGO_FIR_SEQ, taps.data_ptr(), taps.numel()))
    d_mics = torch.arange(n, dtype=torch.int32, device="cuda")
    img = torch.zeros((F, D), device="cuda")
    fp32 = 148 * 128 * ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
    res = {"workload": "stock config: 57x32 directions, 256 mics, 256 samples, 8 taps, %d frames per launch" % F}
    for name, algo in (("naive_sequential", nat.ALGO_FIR_SEQ), ("vectorized_lane_order", nat.ALGO_FIR_LANES)):
        ts = []
        for i in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            nat.check(L.bf_mimo_dev(algo, sig.data_ptr(), img.data_ptr(), F, d_mics.data_ptr(), n, 0, D, stream))
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.mean(ts[3:]))
        fma = F * D * n * N * T / (ms * 1e-3)
        res[name] = {"maps_per_s": F / (ms * 1e-3), "gfma_per_s": fma / 1e9, "kernel_ms": ms,
                     "fp32_frac_of_148x128_lanes": fma / fp32, "kernel": "das_fir_kernel"}
    return res


def e2e_sharded(torch, dist, nat, L, algo, d_pool, d_mics, n, D, F, M, N, rank, world, steps, bounds=None):
    """End to end through the DIRECTION-SHARDED path (the partition `value` measures at N > 1): every step the
    F frames of the batch go host -> device ONCE over PCIe, F / world of them on every rank (each GPU has its
    own link), are all-gathered over NVLink so that every rank holds the whole batch (one NCCL all-gather of the
    INPUT: the only exchange this path needs besides the maps), each rank computes its direction slice with the
    fused kernel + NVLink peer stores, and rank 0 copies the assembled [F][D] maps device -> host.  All copies
    are inside the timed region; H2D + input gather of step i+1 and D2H of step i-1 overlap the kernel of step i
    (three streams).  Returns maps/s (max over ranks) or None when peer memory is unavailable."""
    from lib.sharded import PeerGather, PeerInput
    try:
        peer = PeerGather(D, F, rank, world, dist, depth=4, consume_lag=1, bounds=bounds)
    except RuntimeError:
        return None
    frame_bytes = M * N * 4
    n_host = min(d_pool.shape[0], 3)
    split = F % world == 0                      # otherwise every rank copies the whole batch itself
    Fp = F // world if split else F
    # input all-gather: NCCL by default; BF_E2E_INPUT=p2p uses the copy engines over NVLink peer memory instead
    # (lib.sharded.PeerInput: no SMs needed).  Measured the same at 8 GPUs (83.1-83.6 k against 84.6 k maps/s): this
    # leg is bound by rank 0's host side (~25 driver calls per 1.2 ms step from Python), not by the gather.
    pin = None
    if split and os.environ.get("BF_E2E_INPUT", "nccl") == "p2p":
        try:
            pin = PeerInput((Fp, M, N), rank, world, dist, slots=2)
        except RuntimeError:
            pin = None
    h_in = torch.empty((n_host, Fp, M, N), dtype=torch.float32).pin_memory()
    h_in.copy_(d_pool[:n_host, rank * Fp:(rank + 1) * Fp].cpu() if split else d_pool[:n_host].cpu())
    h_out = torch.empty((2, F, D), dtype=torch.float32).pin_memory() if rank == 0 else None
    d_in = pin.views if pin else torch.empty((2, F, M, N), device="cuda")
    d_part = torch.empty((2, Fp, M, N), device="cuda") if split and not pin else None
    d_asm = torch.empty((2, F, D), device="cuda") if rank == 0 else None
    s_in, s_run, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_run = [torch.cuda.Event() for _ in range(2)]
    ev_asm = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]

    def run(count, base):
        for j in range(count):
            i = base + j
            sl = i & 1
            with torch.cuda.stream(s_in):
                if j >= 2:
                    s_in.wait_event(ev_run[sl])                  # the kernel that read this slot has finished
                    if pin:
                        peer.ready(i - 2, s_in.cuda_stream)      # ... on EVERY rank (its step flag is published)
                if pin:
                    pin.push(sl, h_in[i % n_host], s_in.cuda_stream)
                elif split:
                    d_part[sl].copy_(h_in[i % n_host], non_blocking=True)
                    dist.all_gather_into_tensor(d_in[sl], d_part[sl])        # NVLink; stream-ordered behind the copy
                else:
                    d_in[sl].copy_(h_in[i % n_host], non_blocking=True)
                ev_in[sl].record(s_in)
            with torch.cuda.stream(s_run):
                s_run.wait_event(ev_in[sl])
                if pin:
                    pin.wait(sl, s_run.cuda_stream)              # every rank's frames of this step have arrived
                peer.step(i, algo, d_in[sl], d_mics, n, s_run.cuda_stream)
                ev_run[sl].record(s_run)
                if rank == 0 and j >= 1:                         # consume step i-1 behind the launch of step i
                    pv = (i - 1) & 1
                    if j >= 3:
                        s_run.wait_event(ev_out[pv])             # its previous D2H has left the assembly buffer
                    view = peer.ready(i - 1, s_run.cuda_stream)
                    d_asm[pv].copy_(peer.assemble(view))
                    ev_asm[pv].record(s_run)
                    with torch.cuda.stream(s_out):
                        s_out.wait_event(ev_asm[pv])
                        h_out[pv].copy_(d_asm[pv], non_blocking=True)
                        ev_out[pv].record(s_out)
        if rank == 0:                                            # drain the last step
            i = base + count - 1
            pv = i & 1
            with torch.cuda.stream(s_run):
                view = peer.ready(i, s_run.cuda_stream)
                d_asm[pv].copy_(peer.assemble(view))
                ev_asm[pv].record(s_run)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_asm[pv])
                h_out[pv].copy_(d_asm[pv], non_blocking=True)
        for st in (s_in, s_run, s_out):
            st.synchronize()

    run(3, 0)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(steps, 3)
    dt = time.perf_counter() - t0
    peer.check()
    ok = None
    if rank == 0:                                                # the maps that reached the host == one-GPU maps
        last = 3 + steps - 1
        full = torch.zeros((F, D), device="cuda")
        nat.check(L.bf_mimo_dev_ex(algo, d_in[last & 1].data_ptr(), full.data_ptr(), F, d_mics.data_ptr(), n,
                                   0, D, D, 1, 0, None))
        torch.cuda.synchronize()
        ok = bool(torch.equal(full.cpu(), h_out[last & 1]))
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    if pin:
        pin.check()
        pin.close()
    peer.close()
    return {"value": steps * F / float(t[0]), "unit": "maps/s", "steps": steps,
            "h2d_bytes_per_step_per_rank": Fp * frame_bytes, "h2d_bytes_per_step": F * frame_bytes if split else world * F * frame_bytes,
            "input_gather": ("copy-engine all-gather of the F frames over NVLink peer memory (%d frames per rank over PCIe)" % Fp
                             if pin else "NCCL all-gather of the F frames over NVLink (%d frames per rank over PCIe)" % Fp)
            if split else None,
            "d2h_bytes_per_step_rank0": F * D * 4, "host_maps_bit_exact_vs_one_gpu": ok,
            "api": "host batch in (pinned, F/world frames per rank) -> input all-gather -> bf_mimo_dev_gather_sync (this "
                   "rank's direction slice, peer stores) -> assembled maps out to the host on rank 0; copies inside the "
                   "timed region"}


# ----------------------------------------------------------------------------------------
# the B200 arm
# ----------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--algo", default="pad", choices=["pad", "lerp"])
    ap.add_argument("--frames", type=int, default=128, help="maps per step")
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the MISO / e2e extras")
    ap.add_argument("--exact-sum", type=int, default=1)
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: fused kernel + NVLink peer stores (p2p) or kernel then NCCL all-gather (nccl)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]

    # ---- cpu_baseline first (forks worker processes: before CUDA is initialised) ----------
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu:
        setup = cpu_setup(args.workload, args.algo)
        cores = os.cpu_count() or 1
        libs = reference_builds(setup[1], setup[0])
        probe = cpu_run(args.workload, args.algo, 1, cores, setup, libs[0])
        maps_per_core = max(1, int(12.0 / len(libs) / max(probe["seconds"], 1e-3)))       # ~12 s of CPU work in all
        maps_per_core = min(maps_per_core, 200)
        runs = {lib: cpu_run(args.workload, args.algo, maps_per_core, cores, setup, lib) for lib in libs}
        best = max(runs, key=lambda k: runs[k]["maps_per_s"])
        r = runs[best]
        one = cpu_run(args.workload, args.algo, max(1, maps_per_core // 4), 1, setup, best)
        cpu_base = {"value": r["maps_per_s"], "unit": "maps/s", "cores": cores, "kind": r["kind"],
                    "builds_maps_per_s": {k: v["maps_per_s"] for k, v in runs.items()},
                    "sample": "%d whole %s maps (%d per core, one process per core), %.1f s"
                              % (r["maps"], args.workload.upper(), maps_per_core, r["seconds"]),
                    "one_core_maps_per_s": one["maps_per_s"],
                    "gmac_per_s": r["maps_per_s"] * r["D"] * r["n"] * wl["N_SAMPLES"] / 1e9}

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from interface import config
    config.reload(N_MICROPHONES=wl["N_MICROPHONES"], N_SAMPLES=wl["N_SAMPLES"], MAX_RES_X=wl["MAX_RES_X"],
                  MAX_RES_Y=wl["MAX_RES_Y"], N_TAPS=8, SKIP_N_MICS=1,
                  GEOMETRY_N_MICS=wl["GEOMETRY_N_MICS"], GEOMETRY_N_ARRAYS=wl["GEOMETRY_N_ARRAYS"])
    from lib import _native as nat, directions, synthetic
    L = nat.lib()
    nat.check(L.bf_set_device(local_rank))
    nat.configure_from(config)
    L.bf_set_kernel_options(0, args.exact_sum)
    algo = nat.ALGO_PAD if args.algo == "pad" else nat.ALGO_LERP

    # ---- tables: device generator -> device table, no host round trip ----------------------
    if args.algo == "pad":
        directions.load_pad_from_geometry()
    else:
        directions.load_lerp_from_geometry()
    mics, n = directions.active_microphones()
    mics = nat.i32(mics)
    M, N = config.N_MICROPHONES, config.N_SAMPLES
    D = config.MAX_RES_X * config.MAX_RES_Y
    F = args.frames

    # ---- synthetic inputs: pool of batches larger than L2 ----------------------------------
    delays = directions.calculate_delays().reshape(-1, n)
    src = synthetic.C3 if args.workload == "c3" else synthetic.C1
    sources = [s for s in src["sources"] if s[0] < D]
    base = synthetic.point_sources(delays, mics, M, N, 48828.0, sources, src["noise"], src["seed"])
    del delays
    frame_bytes = M * N * 4
    pool = max(2, int(np.ceil(160e6 / (F * frame_bytes))))
    gen = torch.Generator(device="cuda").manual_seed(1236)
    d_base = torch.from_numpy(base).cuda()
    d_pool = d_base[None, None] + 0.01 * torch.randn((pool, F, M, N), generator=gen, device="cuda")
    d_pool = d_pool.contiguous()
    d_mics = torch.from_numpy(mics).cuda()

    # ---- direction shard of this rank --------------------------------------------------------
    from lib.sharded import shard_bounds
    per, d_begin, d_count = shard_bounds(D, world, rank)
    pipe, peer = None, None
    gather_note = None
    shard_note = None
    if world > 1 and args.gather == "p2p":
        from lib.sharded import PeerGather, weighted_bounds
        bounds = None
        if os.environ.get("BF_SHARD_WEIGHTED", "1") != "0":
            # the GPUs of a node do not all run at the same clock under load and a step is as fast as its slowest
            # rank: time this GPU on a launch without tile rounding (all directions, 32 frames: contiguous unit
            # ranges, every CTA gets the same work), share the times, and size the slices by measured speed
            Fp = min(F, 32)
            probe = torch.zeros((Fp, D), device="cuda")
            ts = []
            for i in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                nat.check(L.bf_mimo_dev_ex(algo, d_pool[i % pool].data_ptr(), probe.data_ptr(), Fp, d_mics.data_ptr(), n,
                                           0, D, D, 1, 0, None))
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            tk = torch.zeros(world, device="cuda", dtype=torch.float64)
            tk[rank] = float(np.mean(ts[1:]))
            dist.all_reduce(tk, op=dist.ReduceOp.SUM)
            bounds = weighted_bounds(D, [1.0 / float(x) for x in tk])
            d_begin, d_count = bounds[rank]
            shard_note = {"probe_ms_all_directions_%d_frames" % Fp: [float(x) for x in tk],
                          "directions_per_rank": [c for _, c in bounds]}
            del probe
        try:
            # overlap: the steps of the timed loop are launched back to back from inputs already resident, each into
            # its own buffer of the ring -- step i + 1 may take over SMs while step i runs its last tiles
            overlap = os.environ.get("BF_GATHER_OVERLAP", "1") != "0" and args.algo == "pad"
            peer = PeerGather(D, F, rank, world, dist, depth=int(os.environ.get("BF_GATHER_DEPTH", "4" if overlap else "2")),
                              bounds=bounds, overlap=overlap)   # kernel stores into every rank's buffer
            d_maps = torch.zeros((F, D), device="cuda")
            fs, ds = D, 1
        except RuntimeError as e:               # raised on every rank together: fall back to the NCCL route
            peer, gather_note = None, str(e)
            per, d_begin, d_count = shard_bounds(D, world, rank)
    if world > 1 and peer is None:
        from lib.sharded import GatherPipeline
        pipe = GatherPipeline(D, F, rank, world, torch.device("cuda"), dist)   # direction-major, gather in place
        d_maps = pipe.bufs[0]
        fs, ds = 1, F
    else:
        d_maps = torch.zeros((F, D), device="cuda")
        fs, ds = D, 1
    stream = torch.cuda.current_stream().cuda_stream

    def step(i):
        sig = d_pool[i % pool]
        if peer:
            peer.step(i, algo, sig, d_mics, n, stream)
            return
        out = pipe.begin(i) if pipe else d_maps
        nat.check(L.bf_mimo_dev_ex(algo, sig.data_ptr(), out.data_ptr(), F, d_mics.data_ptr(), n,
                                   d_begin, d_count, fs, ds, 0, stream))
        if pipe:
            pipe.gather(i)          # on the comm stream, overlapping the next step's kernel

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    L.bf_kernel_launches(1)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter()
    per_step_events = not (peer and peer.overlap)    # an event between two launches would serialise them again
    if peer and os.environ.get("BF_RENDEZVOUS", "1") != "0":
        peer.rendezvous()             # all GPUs enter the timed region together (the host barrier leaves them skewed)
    ev_begin = torch.cuda.Event(enable_timing=True)
    ev_begin.record()
    for i in range(args.steps):
        if per_step_events:
            ev[i][0].record()
        step(args.warmup + i)
        if per_step_events:
            ev[i][1].record()
    if pipe:
        pipe.finish()                 # the last gathers are part of the timed region
    if peer:
        peer.ready(args.warmup + args.steps - 1)      # every rank's last slice has arrived
    ev_end = torch.cuda.Event(enable_timing=True)
    ev_end.record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = int(L.bf_kernel_launches(0))
    # whole timed region on the device (first start -> last stop)
    span_ms = ev_begin.elapsed_time(ev_end)
    dev_ms = sum(a.elapsed_time(b) for a, b in ev) if per_step_events else span_ms
    t = torch.tensor([dev_ms, span_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, span_ms = float(t[0]), float(t[1])
    clocks = sampler.stop() if rank == 0 else None
    value = F * args.steps / (span_ms * 1e-3)

    # ---- sharded result check: the gathered maps of the last step == all directions on one GPU
    gather_check = None
    if peer:
        peer.check()
        last = args.warmup + args.steps - 1
        full = torch.zeros((F, D), device="cuda")
        nat.check(L.bf_mimo_dev_ex(algo, d_pool[last % pool].data_ptr(), full.data_ptr(), F, d_mics.data_ptr(), n,
                                   0, D, D, 1, 0, stream))
        torch.cuda.synchronize()
        got = peer.maps(last)
        ok = torch.tensor([1 if torch.equal(full, got) else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        gather_check = "bit-exact on every rank" if int(ok) else "MISMATCH"
        d_maps = got.contiguous()
    if pipe and rank == 0:
        last = args.warmup + args.steps - 1
        full = torch.zeros((per * world, F), device="cuda")
        nat.check(L.bf_mimo_dev_ex(algo, d_pool[last % pool].data_ptr(), full.data_ptr(), F, d_mics.data_ptr(), n,
                                   0, D, 1, F, 0, stream))
        torch.cuda.synchronize()
        gather_check = "bit-exact" if torch.equal(full[:D], pipe.bufs[last & 1][:D]) else "MISMATCH"

    # ---- kernel-only time of the dominant kernel for the roofline (rank 0's slice) -----------
    kt = []
    for i in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sig = d_pool[(i + 3) % pool]
        a.record()
        nat.check(L.bf_mimo_dev_ex(algo, sig.data_ptr(), d_maps.data_ptr(), F, d_mics.data_ptr(), n,
                                   d_begin, d_count, fs, ds, 0, stream))
        b.record()
        torch.cuda.synchronize()
        kt.append(a.elapsed_time(b))
    k_ms = float(np.mean(kt))
    per_rank_kernel_ms = None
    if world > 1:                                    # every rank's slice kernel alone: which GPU sets the pace
        tk = torch.zeros(world, device="cuda", dtype=torch.float64)
        tk[rank] = k_ms
        dist.all_reduce(tk, op=dist.ReduceOp.SUM)
        per_rank_kernel_ms = [float(x) for x in tk]
    table_bytes = D * n * 4 * (2 if args.algo == "lerp" else 1)
    map_bytes = frame_bytes + table_bytes + D * 4                 # table every map
    launch_bytes = F * map_bytes * (d_count / D)
    hbm_peak, peak_src = peaks()
    achieved = launch_bytes / (k_ms * 1e-3) / 1e9
    traffic = _traffic("das_mimo_%s_F%d" % (args.algo, F))
    adds = F * d_count * n * N * (2 if args.algo == "lerp" else 1)
    fp32_peak = 148 * 128 * (clocks["sm_mhz"] or 1965.0) * 1e6 if clocks else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "kernel": "das_mimo_kernel<%s>" % args.algo, "kernel_ms": k_ms,
                "algorithmic_bytes_per_launch": launch_bytes, "per_rank_kernel_ms": per_rank_kernel_ms,
                "accounting": "table counted once per map (SURVEY 8d primary); per launch = F maps",
                "actual_bound": "fp32 issue (FADD2/FFMA2) -- dense maps are ~63 adds/byte, see DESIGN.md",
                "fp32_lane_ops_per_s": adds / (k_ms * 1e-3),
                "fp32_frac_of_148x128_lanes": (adds / (k_ms * 1e-3)) / fp32_peak if fp32_peak else None}

    # ---- the opt-in tolerance mode beside the configured one (exact_sum = 2: shared-sum kernel) -----------
    shared_sum = None
    if world == 1 and args.exact_sum != 2 and not args.no_extras:
        sig = d_pool[0]
        ref_maps = torch.zeros((F, D), device="cuda")
        nat.check(L.bf_mimo_dev_ex(algo, sig.data_ptr(), ref_maps.data_ptr(), F, d_mics.data_ptr(), n, 0, D, D, 1, 0, stream))
        L.bf_set_kernel_options(0, 2)
        tol_maps = torch.zeros((F, D), device="cuda")
        ts = []
        for i in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            nat.check(L.bf_mimo_dev_ex(algo, d_pool[i % pool].data_ptr(), tol_maps.data_ptr(), F, d_mics.data_ptr(), n,
                                       0, D, D, 1, 0, stream))
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        nat.check(L.bf_mimo_dev_ex(algo, sig.data_ptr(), tol_maps.data_ptr(), F, d_mics.data_ptr(), n, 0, D, D, 1, 0, stream))
        torch.cuda.synchronize()
        L.bf_set_kernel_options(0, args.exact_sum)
        rel = ((tol_maps.double() - ref_maps.double()).abs() / ref_maps.double().clamp_min(1e-300))
        ms = float(np.mean(ts[1:]))
        shared_sum = {"exact_sum": 2, "value": F / (ms * 1e-3), "unit": "maps/s", "kernel_ms": ms,
                      "max_pixel_rel_err_vs_configured_mode": float(rel.max()), "mean_pixel_rel_err": float(rel.mean()),
                      "tolerance": 1e-5,
                      "note": "opt-in: microphones whose delay is uniform over a group of 8 directions are summed once per "
                              "group (different rounding order); the headline stays on the configured exact_sum"}
        del ref_maps, tol_maps

    # ---- e2e: the reference-facing call with host buffers -------------------------------------
    e2e, miso, mvdr, replay, heat, fir, latency = None, None, None, None, None, None, None
    if not args.no_extras:
        # (1) the drop-in per-buffer call: mimo_pad(signals, image, adaptive_array, n), pageable host
        #     memory, one map per call, synchronous (what PC/src/main.pyx loops do per frame)
        h_sig = np.ascontiguousarray(base)
        h_img = np.zeros(D, np.float32)
        fn = L.mimo_pad if args.algo == "pad" else L.mimo_lerp
        for _ in range(5):
            fn(nat.ptr(h_sig), nat.ptr(h_img), nat.ptr(mics), n)
        nat.check()
        single_maps = 64
        barrier()
        t0 = time.perf_counter()
        for _ in range(single_maps):
            fn(nat.ptr(h_sig), nat.ptr(h_img), nat.ptr(mics), n)
        t_single = time.perf_counter() - t0
        nat.check()
        # (2) the batch replay call with host buffers: bf_mimo_host_batch(), pinned host memory,
        #     every step copies its F frames in and its F maps out (copies overlap the kernel)
        e2e_steps = max(3, min(args.steps, 20))
        h_pool = torch.empty((min(pool, 4), F, M, N), dtype=torch.float32).pin_memory()
        h_pool.copy_(d_pool[:h_pool.shape[0]].cpu())
        h_maps = torch.empty((F, D), dtype=torch.float32).pin_memory()
        for i in range(2):
            nat.check(L.bf_mimo_host_batch(algo, h_pool[i % h_pool.shape[0]].data_ptr(), h_maps.data_ptr(), F,
                                           nat.ptr(mics), n))
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            nat.check(L.bf_mimo_host_batch(algo, h_pool[i % h_pool.shape[0]].data_ptr(), h_maps.data_ptr(), F,
                                           nat.ptr(mics), n))
        t_e2e = time.perf_counter() - t0
        checksum = float(h_maps.sum())
        te = torch.tensor([t_e2e, t_single], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": world * e2e_steps * F / float(te[0]), "unit": "maps/s",
               "h2d_bytes_per_step": F * frame_bytes + n * 4, "d2h_bytes_per_step": F * D * 4,
               "api": "bf_mimo_host_batch(algo, signals, images, F, adaptive_array, n): pinned host buffers in, "
                      "host maps out, copies inside the timed region; each rank replays its own frames",
               "steps": e2e_steps, "ms_per_step": 1e3 * float(te[0]) / e2e_steps, "result_checksum": checksum,
               "single_call": {"value": world * single_maps / float(te[1]), "unit": "maps/s",
                               "api": "mimo_%s(signals, image, adaptive_array, n): the reference's per-buffer "
                                      "call, pageable host memory, synchronous" % args.algo,
                               "ms_per_map": 1e3 * float(te[1]) / single_maps}}
        del h_pool, h_maps
        if world > 1 and peer is not None:
            try:
                sh = e2e_sharded(torch, dist, nat, L, algo, d_pool, d_mics, n, D, F, M, N, rank, world, e2e_steps,
                                 bounds=peer.bounds)
            except Exception as ex:  # noqa: BLE001  (every rank reaches the collectives inside or none does)
                sh = {"error": str(ex)}
            e2e["partition"] = "replicas: each rank replays its own frames (recordings shard with no collective)"
            e2e["sharded"] = sh

        # ---- extra: BASELINE config C5 (bounded sample), then C2 and C4 -----------------------
        rest = os.environ.get("BF_EXTRAS") != "e2e"            # tools: BF_EXTRAS=e2e stops after the end-to-end legs
        if not rest:
            pass
        elif world > 1 and args.workload == "c3":
            frames_r, ms_r, err_r = 0, 0.0, None
            stream_r = None
            try:
                dist.barrier()
                stream_r = replay_c5_stream(torch, nat, algo, d_mics, n, M, rank)
                if stream_r is not None:
                    frames_r, ms_r = stream_r["frames"], 1e3 * stream_r["seconds"]
                else:
                    frames_r, ms_r = replay_c5_sharded_local(torch, nat, algo, d_mics, n, D, M, N, rank)
            except Exception as e:  # noqa: BLE001
                err_r = str(e)
            tr = torch.tensor([ms_r, float(frames_r), 0.0 if err_r is None else 1.0], device="cuda", dtype=torch.float64)
            tmax = tr.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)            # slowest rank, any failure
            dist.all_reduce(tr, op=dist.ReduceOp.SUM)
            if float(tmax[2]) > 0:
                replay = {"error": err_r or "a peer rank failed"}
            else:
                fps_all = float(tr[1]) / (float(tmax[0]) * 1e-3)
                if stream_r is not None:
                    replay = {"workload": "C5 sharded: %d one-hour recordings in the wire format (one per GPU, streamed from "
                                          "pinned host memory: %.0f min resident x %d passes each), 30 fps video, 180x180 maps, "
                                          "640x360 overlay, %d frames in total, no collective"
                                          % (world, stream_r["host_recording_minutes"], stream_r["passes"], int(tr[1])),
                              "frames_per_s": fps_all, "x_realtime_at_30fps": fps_all / 30.0, "ms": float(tmax[0]),
                              "rank0": stream_r}
                else:
                    replay = {"workload": "C5 sharded: %d recordings of 20 s (one per GPU, 1.0 GB each, resident), 30 fps "
                                          "video, 180x180 maps, %d frames in total, no collective" % (world, int(tr[1])),
                              "frames_per_s": fps_all, "x_realtime_at_30fps": fps_all / 30.0, "ms": float(tmax[0])}
        elif rank == 0 and args.workload == "c3":
            try:
                replay = replay_c5(args, nat, torch, algo, d_mics, n, D, M, N)
                replay["one_hour_stream"] = replay_c5_stream(torch, nat, algo, d_mics, n, M, 0)
            except Exception as e:  # noqa: BLE001
                replay = {"error": str(e)}
        mvdr_sh = None
        if rest and world > 1 and peer is not None:      # every rank: direction-sharded C4 (peer memory available)
            try:
                mvdr_sh = mvdr_c4_sharded(torch, dist, nat, L, rank, world)
            except Exception as ex:  # noqa: BLE001
                mvdr_sh = {"error": str(ex)}
        if rest and rank == 0 and args.workload == "c3":
            try:
                hm = d_maps if (world == 1 or peer) else d_maps[:D].t().contiguous()   # [F][D] maps of the last step
                heat = heatmap_c5(torch, hm, hbm_peak)
            except Exception as e:  # noqa: BLE001
                heat = {"error": str(e)}
        if rest and rank == 0:
            try:
                miso = miso_c2(args, nat, L, config, directions, torch, stream, hbm_peak)
            except Exception as e:  # noqa: BLE001
                miso = {"error": str(e)}
            try:
                mvdr = mvdr_c4(nat, L, torch, stream)
            except Exception as e:  # noqa: BLE001
                mvdr = {"error": str(e)}
            if mvdr_sh is not None and isinstance(mvdr, dict):
                mvdr["sharded"] = mvdr_sh
            try:
                fir = fir_default(nat, L, config, torch, stream, clocks)
            except Exception as e:  # noqa: BLE001
                fir = {"error": str(e)}
            try:
                latency = live_latency(nat, L, config, directions)
            except Exception as e:  # noqa: BLE001
                latency = {"error": str(e)}

    if rank == 0:
        line = {
            "metric": "beam power maps/s", "value": value, "unit": "maps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": span_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "gmac_per_s": value * D * n * N / 1e9,
            "config": {"workload": wl["desc"], "algo": args.algo, "frames_per_step": F, "directions": D,
                       "mics": n, "samples": N, "l2": "inputs larger than L2 (pool of %d batches, %.0f MB)"
                       % (pool, pool * F * frame_bytes / 1e6),
                       "parallelism": "directions sharded over %d rank(s)%s" % (
                           world, (", all-gather fused into the kernel: epilogue stores go to every rank's buffer over NVLink peer memory"
                            if peer else ", one in-place NCCL all-gather per step on a second stream") if world > 1 else ""),
                       "exact_sum": args.exact_sum},
            "sum_step_ms": dev_ms, "wall_s": t_wall, "gather_note": gather_note, "shard_weights": shard_note, "gpu_launches": launches, "clocks": clocks,
            "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu_base, "miso": miso, "mvdr": mvdr, "replay": replay, "heatmap": heat, "fir": fir,
            "latency": latency, "shared_sum_mode": shared_sum,
            # last on purpose: the tail of the line is what a truncated log keeps
            "e2e_sharded_maps_per_s": (e2e or {}).get("sharded", {}).get("value") if isinstance((e2e or {}).get("sharded"), dict) else None,
            "gather_check": gather_check,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
