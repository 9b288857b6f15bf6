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


def assemble_reference(slices, n_directions):
    """NumPy model of the gather: list of per-rank [per][F] arrays -> [D][F]."""
    return np.concatenate(slices, axis=0)[:n_directions]
