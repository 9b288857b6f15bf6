"""Direction-sharded power maps over several GPUs (one process per GPU).

Every `img[d]` depends only on table row d and on the whole sample buffer
(pad_and_sum.c:114-131), so the direction grid is cut into `world` contiguous slices
with no halo.  Each rank computes its slice for a batch of F frames straight into a
direction-major buffer `maps[D_padded][F]` (bf_mimo_dev_ex with frame_stride = 1,
dir_stride = F): a rank's slice is then one contiguous block and ONE in-place
all-gather (NCCL over NVLink on GPUs; gloo in the CPU tests) assembles the maps on every
rank.  There is no other collective on this path.

PyTorch is used for device memory and the process group only.
"""
import numpy as np


def shard_bounds(n_directions, world, rank):
    """(per_rank, d_begin, d_count): equal slices of ceil(D/world); the last ranks may be
    short or empty.  The gather buffer holds per_rank*world rows."""
    per = (n_directions + world - 1) // world
    d_begin = min(rank * per, n_directions)
    d_count = max(0, min(per, n_directions - d_begin))
    return per, d_begin, d_count


def weighted_bounds(n_directions, weights, multiple=8):
    """Direction slices proportional to `weights` (one per rank, e.g. 1 / measured kernel time of that GPU), in
    multiples of `multiple` directions (the kernel's group size) except the last: -> list of (d_begin, d_count).
    On a node whose GPUs do not run at the same clock the step is as fast as its slowest rank; giving that rank
    fewer directions evens the ranks out."""
    w = np.asarray(weights, np.float64)
    w = w / w.sum()
    edges = np.round(np.cumsum(w) * n_directions / multiple).astype(np.int64) * multiple
    edges[-1] = n_directions
    edges = np.minimum(np.maximum.accumulate(edges), n_directions)
    out, begin = [], 0
    for e in edges:
        out.append((int(begin), int(e - begin)))
        begin = int(e)
    return out


def gather_layout(n_directions, world, bounds=None):
    """Slices and row stride of the fused gather buffers float [world][F][per]: -> (bounds, per).

    bounds: list of (d_begin, d_count) per rank tiling [0, D) in rank order (None: equal slices, shard_bounds).
    per: the largest slice rounded up to 8 directions -- a warp stores its group of 8 values as ONE 32-byte write
    per peer, which must not straddle two 32-byte sectors (equal slices of 4050 directions did: 100.2 k instead of
    103.4 k maps/s on 8 GPUs)."""
    if bounds is None:
        bounds = [shard_bounds(n_directions, world, r)[1:] for r in range(world)]
    bounds = [(int(b), int(c)) for b, c in bounds]
    if len(bounds) != world or bounds[0][0] != 0 or any(c < 0 for _, c in bounds) or \
            any(bounds[r][0] + bounds[r][1] != (bounds[r + 1][0] if r + 1 < world else n_directions)
                for r in range(world)):
        raise ValueError("PeerGather: bounds must tile [0, D) in rank order")
    per = (max(c for _, c in bounds) + 7) // 8 * 8
    return bounds, max(per, 8)


class ShardedMaps:
    """maps = ShardedMaps(D, frames, rank, world, device); maps.compute(fn); maps.gather()

    compute_slice(d_begin, d_count, out_rows) must fill out_rows[d_count][frames] (a view of
    the gather buffer) -- on the GPU that is one bf_mimo_dev_ex launch, see `gpu_compute`."""

    def __init__(self, n_directions, frames, rank, world, device, dist=None):
        import torch
        self.torch, self.dist = torch, dist
        self.D, self.F, self.rank, self.world = n_directions, frames, rank, world
        self.per, self.d_begin, self.d_count = shard_bounds(n_directions, world, rank)
        self.buf = torch.zeros((self.per * world, frames), dtype=torch.float32, device=device)

    def my_rows(self):
        return self.buf[self.rank * self.per:(self.rank + 1) * self.per]

    def gather(self):
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.buf, self.my_rows())
        return self.buf[:self.D]

    def maps(self):
        """[frames][D] view (transposed, non-contiguous) of the assembled maps."""
        return self.buf[:self.D].t()

    def gpu_compute(self, lib, algo, d_signals, d_mic_ids, n, stream=None):
        """One launch of the tiled kernel for this rank's slice of all frames."""
        if self.d_count == 0:
            return 0
        return lib.bf_mimo_dev_ex(algo, d_signals.data_ptr(), self.buf.data_ptr(), self.F,
                                  d_mic_ids.data_ptr(), n, self.d_begin, self.d_count, 1, self.F, 0,
                                  stream)


class GatherPipeline:
    """Two gather buffers and a communication stream: the all-gather of batch i runs on the comm
    stream while the kernel of batch i+1 fills the other buffer (CUDA only).

        buf = pipe.begin(i)            # waits until the gather that last used this buffer is done
        ... launch the slice kernel into pipe.my_rows(buf) on the current stream ...
        pipe.gather(i)                 # all-gather of buf on the comm stream, after the kernel
        pipe.finish()                  # current stream waits for every outstanding gather
    """

    def __init__(self, n_directions, frames, rank, world, device, dist):
        import torch
        self.torch, self.dist = torch, dist
        self.D, self.F, self.rank, self.world = n_directions, frames, rank, world
        self.per, self.d_begin, self.d_count = shard_bounds(n_directions, world, rank)
        self.bufs = [torch.zeros((self.per * world, frames), dtype=torch.float32, device=device) for _ in range(2)]
        self.comm = torch.cuda.Stream(device=device)
        self.done = [None, None]

    def my_rows(self, buf):
        return buf[self.rank * self.per:(self.rank + 1) * self.per]

    def begin(self, i):
        if self.done[i & 1] is not None:
            self.torch.cuda.current_stream().wait_event(self.done[i & 1])
        return self.bufs[i & 1]

    def gather(self, i):
        torch = self.torch
        buf = self.bufs[i & 1]
        ready = torch.cuda.Event()
        ready.record()
        self.comm.wait_event(ready)
        with torch.cuda.stream(self.comm):
            self.dist.all_gather_into_tensor(buf, self.my_rows(buf))
            ev = torch.cuda.Event()
            ev.record()
        self.done[i & 1] = ev
        return buf

    def finish(self):
        self.torch.cuda.current_stream().wait_stream(self.comm)


class _DevArray:
    """Device allocation owned by the library, visible to torch through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr):
        self.ptr = ptr
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 3}


class PeerGather:
    """Fused power maps + all-gather over NVLink peer memory (no collective on the data path).

    Every rank owns `depth` gather buffers float [world][F][per] and one flag array; the buffers of
    all ranks are mapped into every process through CUDA IPC (handles exchanged once with
    torch.distributed -- plumbing only).  step(i, ...) launches the tiled kernel, whose epilogue
    stores each finished value into slice `rank` of all `world` buffers, then publishes step i to
    every rank's flags.  Buffer i % depth is overwritten by step i + depth: the writers first wait
    (inside the kernel) until every rank has published step i + 1 + consume_lag, so whatever reads the
    maps of step i must be enqueued on the compute stream before step i + 1 + consume_lag is launched
    (consume_lag = 0: consume right after the step; 1: consume while the next step runs; needs
    depth >= consume_lag + 2).  maps(i) is the assembled [F][D] tensor of step i.
    """

    def __init__(self, n_directions, frames, rank, world, dist, depth=2, consume_lag=0, bounds=None, overlap=False):
        """bounds: optional list of (d_begin, d_count) per rank (weighted_bounds) instead of equal slices; the
        gather buffers are sized for the largest slice and maps() drops the padding.
        overlap: launch the steps so that step i + 1 may take over SMs while step i still runs its last tiles
        (bf_gather_overlap in bf_b200.h; pad only).  The caller promises that the inputs of a step are complete
        before the previous step is launched -- nothing that produces them is enqueued between two step() calls."""
        import ctypes
        if depth < consume_lag + 2:
            raise ValueError("PeerGather: depth must be at least consume_lag + 2")
        self.consume_lag = consume_lag
        self.overlap = bool(overlap)
        import torch
        from . import _native
        self.torch, self.nat, self.L = torch, _native, _native.lib()
        self.D, self.F, self.rank, self.world, self.depth = n_directions, frames, rank, world, depth
        self.bounds, self.per = gather_layout(n_directions, world, bounds)
        bounds = self.bounds
        self.d_begin, self.d_count = bounds[rank]
        L = self.L
        vp = ctypes.c_void_p
        n_buf = world * frames * self.per

        def alloc(nbytes):
            out = vp()
            _native.check(L.bf_dev_alloc(ctypes.c_size_t(nbytes), ctypes.byref(out)))
            return out.value
        # Every collective below is reached by every rank whatever fails locally: a failure is carried as
        # data and raised on all ranks together (the caller can then fall back to the NCCL route).
        self.own, self.own_flags, self._opened = [], None, []
        handles, err = None, None
        try:
            import os
            if os.environ.get("BF_PEER_GATHER_DISABLE") == str(rank) or os.environ.get("BF_PEER_GATHER_DISABLE") == "all":
                raise RuntimeError("disabled by BF_PEER_GATHER_DISABLE (test hook)")
            self.own = [alloc(n_buf * 4) for _ in range(depth)]
            self.own_flags = alloc(8 * 8)
            handles = []
            for ptr in self.own + [self.own_flags]:
                h = (ctypes.c_ubyte * 64)()
                _native.check(L.bf_ipc_export(vp(ptr), h))
                handles.append(bytes(h))
        except Exception as e:  # noqa: BLE001
            handles, err = None, "export: %s" % e
        everyone = [None] * world
        dist.all_gather_object(everyone, handles)
        self.bufs = [[None] * world for _ in range(depth)]      # [depth][rank] device pointers
        self.flags = [None] * world
        if err is None and any(h is None for h in everyone):
            err = "a peer could not export its buffers"
        if err is None:
            try:
                for r in range(world):
                    for k in range(depth + 1):
                        if r == rank:
                            ptr = (self.own + [self.own_flags])[k]
                        else:
                            out = vp()
                            hb = (ctypes.c_ubyte * 64).from_buffer_copy(everyone[r][k])
                            _native.check(L.bf_ipc_open(hb, ctypes.byref(out)))
                            ptr = out.value
                            self._opened.append(ptr)
                        if k < depth:
                            self.bufs[k][r] = ptr
                        else:
                            self.flags[r] = ptr
            except Exception as e:  # noqa: BLE001
                err = "open: %s" % e
        errs = [None] * world
        dist.all_gather_object(errs, err)
        bad = [(r, e) for r, e in enumerate(errs) if e]
        if bad:
            for ptr in self._opened:
                L.bf_ipc_close(vp(ptr))
            dist.barrier()
            for ptr in self.own + ([self.own_flags] if self.own_flags else []):
                L.bf_dev_free(vp(ptr))
            raise RuntimeError("PeerGather: peer memory unavailable (rank %d: %s)" % bad[0])
        self._buf_arrays = [(vp * world)(*[vp(x) for x in self.bufs[k]]) for k in range(depth)]
        self._flag_array = (vp * world)(*[vp(x) for x in self.flags])
        self.views = [torch.as_tensor(_DevArray(self.own[k], (world, frames, self.per), "<f4"), device="cuda")
                      for k in range(depth)]
        self.timed_out = torch.zeros(1, dtype=torch.int32, device="cuda")
        self._seq = 0
        self._seq_of = {}
        self.dist_barrier = dist.barrier
        dist.barrier()

    def step(self, i, algo, d_signals, d_mic_ids, n, stream=None):
        """Launch step i (i must increase by one per call): ONE kernel that computes the slice, stores it
        into every rank's buffer and publishes the step.  Before its first store into buffer i % depth it
        waits (inside the kernel) until every rank has published step i - depth + 1, i.e. has enqueued
        everything that reads that buffer's previous contents."""
        nat, L = self.nat, self.L
        st = stream if stream is not None else self.torch.cuda.current_stream().cuda_stream
        self._seq += 1
        seq = self._seq
        self._seq_of[i] = seq
        k = i % self.depth
        wait_seq = max(0, seq - self.depth + 1 + self.consume_lag)
        if self.overlap:
            L.bf_gather_overlap(1)
        try:
            nat.check(L.bf_mimo_dev_gather_sync(algo, d_signals.data_ptr(), self.F, d_mic_ids.data_ptr(), n,
                                                self.d_begin, self.d_count, self.rank, self.world,
                                                self._buf_arrays[k], self.per, self._flag_array, wait_seq, seq,
                                                self.timed_out.data_ptr(), st))
        finally:
            if self.overlap:
                L.bf_gather_overlap(0)

    def scatter(self, i, d_slice, stream=None):
        """Step i for a producer that does not store to the peers itself (the frequency-domain maps): d_slice is
        this rank's finished slice, CUDA float32 [F][d_count].  Waits (on the stream) until the buffer's previous
        contents have been consumed everywhere, stores the slice into every rank's buffer over NVLink and
        publishes the step -- the same protocol as step()."""
        nat, L = self.nat, self.L
        st = stream if stream is not None else self.torch.cuda.current_stream().cuda_stream
        self._seq += 1
        seq = self._seq
        self._seq_of[i] = seq
        k = i % self.depth
        wait_seq = max(0, seq - self.depth + 1 + self.consume_lag)
        if wait_seq > 0:
            nat.check(L.bf_gather_wait(self.flags[self.rank], self.world, wait_seq, self.timed_out.data_ptr(), st))
        if self.d_count > 0:
            assert tuple(d_slice.shape) == (self.F, self.d_count) and d_slice.is_contiguous()
            nat.check(L.bf_peer_scatter(d_slice.data_ptr(), self.d_count, self.F, self.rank, self.world,
                                        self._buf_arrays[k], self.per, st))
        nat.check(L.bf_gather_signal(self._flag_array, self.world, self.rank, seq, st))

    def rendezvous(self, stream=None):
        """Device-side barrier over the step flags: every rank publishes one (empty) step and its stream waits for
        everyone's.  After a host barrier the ranks still leave it tens to hundreds of microseconds apart; with this
        in front of a timed run all GPUs start their first step together."""
        st = stream if stream is not None else self.torch.cuda.current_stream().cuda_stream
        self._seq += 1
        self.nat.check(self.L.bf_gather_signal(self._flag_array, self.world, self.rank, self._seq, st))
        self.nat.check(self.L.bf_gather_wait(self.flags[self.rank], self.world, self._seq,
                                             self.timed_out.data_ptr(), st))

    def ready(self, i, stream=None):
        """Make the stream wait until every rank's slice of step i has arrived; returns the
        [world][F][per] view of that step."""
        st = stream if stream is not None else self.torch.cuda.current_stream().cuda_stream
        self.nat.check(self.L.bf_gather_wait(self.flags[self.rank], self.world, self._seq_of[i],
                                             self.timed_out.data_ptr(), st))
        return self.views[i % self.depth]

    def assemble(self, view):
        """[world][F][per] gather view -> [F][D] maps (a copy; drops the padding of unequal slices)."""
        return self.torch.cat([view[r, :, :c] for r, (_, c) in enumerate(self.bounds)], dim=1)

    def maps(self, i):
        """[F][D] tensor of step i (a copy: the gather layout is [rank][F][per]); waits for the step."""
        return self.assemble(self.ready(i))

    def check(self):
        if int(self.timed_out.item()):
            raise RuntimeError("PeerGather: a peer did not publish its step within the spin limit")

    def close(self):
        import ctypes
        self.torch.cuda.synchronize()
        for ptr in self._opened:
            self.L.bf_ipc_close(ctypes.c_void_p(ptr))
        self._opened = []
        self.dist_barrier()                       # peers have unmapped our buffers before they are freed
        self.views = []
        for ptr in self.own + [self.own_flags]:
            self.L.bf_dev_free(ctypes.c_void_p(ptr))
        self.own, self.own_flags = [], None


class PeerInput:
    """All-gather of the INPUT frames of a sharded step over NVLink, on the copy engines.

    Every rank brings F / world frames of a step over its own PCIe link; every rank needs all F.  An NCCL all-gather
    does that with a kernel, and a kernel has to wait for SMs: the persistent power-map kernel holds all of them, so
    the collective only ran in the gap between two steps (8 GPUs: 1.51 ms per step against 1.27 ms of kernel).
    Here every rank owns `slots` buffers float [world][Fp][M][N] (= [F][M][N], frames rank-major), mapped into
    every process through CUDA IPC like the gather buffers; push() copies this rank's part host -> own buffer, then
    own buffer -> the same place in every peer's buffer with cudaMemcpyAsync (bf_peer_copy) and publishes the step in
    every rank's arrival flags; wait() makes a stream wait until all parts of the step have arrived.  The caller
    must not push into a slot while a peer's kernel still reads it: wait for the step flags of the step that last
    used the slot first (PeerGather.ready(i, stream) on the copy stream)."""

    def __init__(self, part_shape, rank, world, dist, slots=2):
        import ctypes
        import torch
        from . import _native
        self.torch, self.nat, self.L = torch, _native, _native.lib()
        self.rank, self.world, self.slots = rank, world, slots
        self.part_shape = tuple(int(x) for x in part_shape)
        self.part_bytes = 4 * int(np.prod(self.part_shape))
        L, vp = self.L, ctypes.c_void_p

        def alloc(nbytes):
            out = vp()
            _native.check(L.bf_dev_alloc(ctypes.c_size_t(nbytes), ctypes.byref(out)))
            return out.value
        self.own, self.own_flags, self._opened = [], None, []
        handles, err = None, None
        try:
            self.own = [alloc(world * self.part_bytes) for _ in range(slots)]
            self.own_flags = alloc(8 * 8)
            handles = []
            for ptr in self.own + [self.own_flags]:
                h = (ctypes.c_ubyte * 64)()
                _native.check(L.bf_ipc_export(vp(ptr), h))
                handles.append(bytes(h))
        except Exception as e:  # noqa: BLE001  (carried as data: every rank reaches the collectives below)
            handles, err = None, "export: %s" % e
        everyone = [None] * world
        dist.all_gather_object(everyone, handles)
        self.bufs = [[None] * world for _ in range(slots)]
        self.flags = [None] * world
        if err is None and any(h is None for h in everyone):
            err = "a peer could not export its buffers"
        if err is None:
            try:
                for r in range(world):
                    for k in range(slots + 1):
                        if r == rank:
                            ptr = (self.own + [self.own_flags])[k]
                        else:
                            out = vp()
                            hb = (ctypes.c_ubyte * 64).from_buffer_copy(everyone[r][k])
                            _native.check(L.bf_ipc_open(hb, ctypes.byref(out)))
                            ptr = out.value
                            self._opened.append(ptr)
                        if k < slots:
                            self.bufs[k][r] = ptr
                        else:
                            self.flags[r] = ptr
            except Exception as e:  # noqa: BLE001
                err = "open: %s" % e
        errs = [None] * world
        dist.all_gather_object(errs, err)
        bad = [(r, e) for r, e in enumerate(errs) if e]
        if bad:
            for ptr in self._opened:
                L.bf_ipc_close(vp(ptr))
            dist.barrier()
            for ptr in self.own + ([self.own_flags] if self.own_flags else []):
                L.bf_dev_free(vp(ptr))
            raise RuntimeError("PeerInput: peer memory unavailable (rank %d: %s)" % bad[0])
        self._flag_array = (vp * world)(*[vp(x) for x in self.flags])
        full = (world * self.part_shape[0],) + self.part_shape[1:]
        self.views = [torch.as_tensor(_DevArray(self.own[k], full, "<f4"), device="cuda") for k in range(slots)]
        self.timed_out = torch.zeros(1, dtype=torch.int32, device="cuda")
        self._seq = 0
        self._seq_of_slot = [0] * slots
        self.dist_barrier = dist.barrier
        dist.barrier()

    def push(self, slot, h_part, stream=None):
        """h_part: pinned host (or device) float32 tensor of part_shape: this rank's frames of the step.  Enqueues,
        on `stream`: part -> own buffer, own buffer -> every peer's buffer (copy engines), arrival flag."""
        import ctypes
        torch, L, vp = self.torch, self.L, ctypes.c_void_p
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        Fp = self.part_shape[0]
        off = self.rank * self.part_bytes
        nat = self.nat
        nat.check(L.bf_peer_copy(vp(self.own[slot] + off), vp(h_part.data_ptr()), self.part_bytes, st))
        for r in range(self.world):
            if r != self.rank:
                nat.check(L.bf_peer_copy(vp(self.bufs[slot][r] + off), vp(self.own[slot] + off), self.part_bytes, st))
        self._seq += 1
        self._seq_of_slot[slot] = self._seq
        nat.check(L.bf_gather_signal(self._flag_array, self.world, self.rank, self._seq, st))
        return Fp

    def wait(self, slot, stream=None):
        """Make `stream` wait until every rank's part of the step last pushed into `slot` has arrived; returns the
        [F][...] tensor of the slot.  Every rank pushes the same steps into the same slots, so the sequence
        numbers agree."""
        st = stream if stream is not None else self.torch.cuda.current_stream().cuda_stream
        self.nat.check(self.L.bf_gather_wait(self.flags[self.rank], self.world, self._seq_of_slot[slot],
                                             self.timed_out.data_ptr(), st))
        return self.views[slot]

    def check(self):
        if int(self.timed_out.item()):
            raise RuntimeError("PeerInput: a peer's frames did not arrive within the spin limit")

    def close(self):
        import ctypes
        self.torch.cuda.synchronize()
        for ptr in self._opened:
            self.L.bf_ipc_close(ctypes.c_void_p(ptr))
        self._opened = []
        self.dist_barrier()
        self.views = []
        for ptr in self.own + [self.own_flags]:
            self.L.bf_dev_free(ctypes.c_void_p(ptr))
        self.own, self.own_flags = [], None


def assemble_peer_layout(buf, n_directions):
    """Gather-buffer layout of the fused path, [world][F][per] (slice r = rank r's directions), to maps
    [F][D].  Works on NumPy arrays and torch tensors alike."""
    world, F, per = buf.shape
    if hasattr(buf, "permute"):
        return buf.permute(1, 0, 2).reshape(F, world * per)[:, :n_directions]
    return buf.transpose(1, 0, 2).reshape(F, world * per)[:, :n_directions]


def assemble_reference(slices, n_directions):
    """NumPy model of the gather: list of per-rank [per][F] arrays -> [D][F]."""
    return np.concatenate(slices, axis=0)[:n_directions]


def fd_mvdr_sharded(peer, i, d_snapshots, K, loading, stream=None):
    """Step i of a direction-sharded MVDR map (FD path): every rank computes spectra, covariance,
    factor and inverse in full (0.5 % of the work) and steers only its slice of the directions; the slices are
    exchanged through `peer` (PeerGather(D, 1, ...)).  Returns nothing; peer.maps(i) is the assembled [1][D] map."""
    torch, nat, L = peer.torch, peer.nat, peer.L
    st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
    part = torch.empty((1, max(peer.d_count, 1)), dtype=torch.float32, device="cuda")
    if peer.d_count > 0:
        nat.check(L.bf_fd_mvdr_dev_slice(d_snapshots.data_ptr(), part.data_ptr(), K, loading, peer.d_begin,
                                         peer.d_count, st))
    peer.scatter(i, part[:, :peer.d_count].contiguous() if peer.d_count else part[:, :0], st)


def fd_mvdr_sharded_bins(peer, i, d_snapshots, K, loading, n_bins, dist, stream=None):
    """Step i of an MVDR map sharded TWICE: the float64 stages (spectra, covariance, Cholesky, inverse, operand
    images) by BINS -- rank r factors bins [r*B/W, (r+1)*B/W) and the 1 MB-per-bin operand images + per-bin scales
    are all-gathered over NVLink (two NCCL all-gathers, in place) -- and the steering contraction by DIRECTIONS as
    in fd_mvdr_sharded.  Needs n_bins divisible by the number of ranks (else use fd_mvdr_sharded).
    peer.maps(i) is the assembled [1][D] map."""
    import ctypes
    torch, nat, L = peer.torch, peer.nat, peer.L
    world, rank = peer.world, peer.rank
    if n_bins % world:
        raise ValueError("fd_mvdr_sharded_bins: %d bins do not divide over %d ranks" % (n_bins, world))
    st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
    fc = n_bins // world
    nat.check(L.bf_fd_mvdr_factor_dev(d_snapshots.data_ptr(), K, loading, rank * fc, fc, st))
    img, per_bin, scale = ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_void_p()
    nat.check(L.bf_fd_mvdr_operands(ctypes.byref(img), ctypes.byref(per_bin), ctypes.byref(scale)))
    t_img = torch.as_tensor(_DevArray(img.value, (n_bins, per_bin.value), "|u1"), device="cuda")
    t_scale = torch.as_tensor(_DevArray(scale.value, (n_bins,), "<f4"), device="cuda")
    dist.all_gather_into_tensor(t_img, t_img[rank * fc:(rank + 1) * fc])          # in place: slice r of the buffer
    dist.all_gather_into_tensor(t_scale, t_scale[rank * fc:(rank + 1) * fc])
    part = torch.empty((1, max(peer.d_count, 1)), dtype=torch.float32, device="cuda")
    if peer.d_count > 0:
        nat.check(L.bf_fd_mvdr_steer_dev(part.data_ptr(), peer.d_begin, peer.d_count, st))
    peer.scatter(i, part[:, :peer.d_count].contiguous() if peer.d_count else part[:, :0], st)


def fd_das_sharded(peer, i, d_signals, threshold=0.2, normalise=True, stream=None):
    """Step i of direction-sharded frequency-domain DAS for peer.F frames: every rank transforms all channels,
    steers its slice, scatters it; the assembled [F][D] power is normalised like the reference after the
    gather (beam_forming_algorithm.py:58-63).  Returns the normalised (or raw) [F][D] maps."""
    torch, nat, L = peer.torch, peer.nat, peer.L
    st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
    part = torch.empty((peer.F, max(peer.d_count, 1)), dtype=torch.float32, device="cuda")
    if peer.d_count > 0:
        part = torch.empty((peer.F, peer.d_count), dtype=torch.float32, device="cuda")
        nat.check(L.bf_fd_das_dev_slice(d_signals.data_ptr(), part.data_ptr(), peer.F, peer.d_begin, peer.d_count, st))
    peer.scatter(i, part, st)
    maps = peer.maps(i).contiguous()
    nat.check(L.bf_fd_normalise_dev(maps.data_ptr(), peer.F, float(threshold), 1 if normalise else 0, st))
    return maps
