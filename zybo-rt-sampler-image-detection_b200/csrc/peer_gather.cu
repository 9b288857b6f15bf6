// peer_gather.cu -- plumbing of the fused "power maps + all-gather" path.
//
// Direction-sharded runs need every rank's slice of the maps on every rank.  Instead of a
// separate collective after the kernel, the tiled kernel stores each finished value into its
// own buffer AND into the same position of every peer's buffer (plain global stores to memory
// mapped through CUDA IPC; they travel over NVLink / NVSwitch tile by tile while the kernel
// keeps computing).  This file holds what surrounds that: device allocations that can be
// exported, IPC export / open, and the step flags (a rank publishes "my step s is complete" into
// every peer's flag array after its kernel; a one-thread kernel waits until all ranks did).
#include "bf_common.cuh"

namespace bf {

struct FlagPtrs { long long *p[8]; };

__global__ void gather_signal_kernel(const FlagPtrs flags, int n, int rank, long long step)
{
    // runs after the map kernel on the same stream: its peer stores are complete; publish
    __threadfence_system();
    if (threadIdx.x < n) {
        volatile long long *f = flags.p[threadIdx.x] + rank;
        *f = step;
    }
    __threadfence_system();
}

__global__ void gather_wait_kernel(const long long *flags, int world, long long step, long long spin_limit,
                                   int *timed_out)
{
    if (threadIdx.x >= world) return;
    const volatile long long *f = flags + threadIdx.x;
    long long spins = 0;
    while (*f < step) {
        __nanosleep(200);
        if (++spins > spin_limit) { *timed_out = 1; break; }       // never hang the GPU on a lost peer
    }
    __threadfence_system();
}

// Device-side all-gather of a finished slice: src float [frames][count] -> slice `rank` of every rank's
// gather buffer float [world][frames][per_rank] (own memory and, through CUDA IPC, the peers': plain global
// stores over NVLink).  Used where the producing kernel does not store to the peers itself (the frequency-
// domain maps); the time-domain kernel fuses these stores into its epilogue.
struct ScatterPtrs { float *p[8]; };
__global__ void peer_scatter_kernel(const float *__restrict__ src, long count, int frames, int rank, int world,
                                    const ScatterPtrs bufs, long per_rank)
{
    const long total = count * frames;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long f = i / count, d = i - f * count;
        const float v = src[i];
        const long off = ((long)rank * frames + f) * per_rank + d;
        for (int r = 0; r < world; r++) bufs.p[r][off] = v;
    }
}

}  // namespace bf

using namespace bf;

extern "C" int bf_peer_scatter(const float *d_src, long count, int frames, int rank, int world,
                               void *const *gather_bufs, long per_rank, void *stream)
{
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    if (!d_src || !gather_bufs || count < 0 || frames < 1 || world < 1 || world > 8 || rank < 0 || rank >= world ||
        per_rank < count) {
        set_error(BF_ERR_ARG, "bf_peer_scatter: bad arguments");
        return BF_ERR_ARG;
    }
    if (count == 0) return BF_OK;
    ScatterPtrs sp{};
    for (int r = 0; r < world; r++) {
        if (!gather_bufs[r]) { set_error(BF_ERR_ARG, "bf_peer_scatter: null buffer for rank %d", r); return BF_ERR_ARG; }
        sp.p[r] = (float *)gather_bufs[r];
    }
    const long total = count * frames;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    peer_scatter_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_src, count, frames, rank, world, sp, per_rank);
    BF_CHECK_LAUNCH();
    count_launch();
    return BF_OK;
}

// Copy between device buffers of this GPU and of a peer (opened with bf_ipc_open): cudaMemcpyAsync, i.e. the copy
// engines over NVLink -- no SMs, so it runs while a persistent kernel holds every SM (a collective kernel would wait
// for the gap between two launches).
extern "C" int bf_peer_copy(void *dst, const void *src, size_t bytes, void *stream)
{
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    if (!dst || !src) { set_error(BF_ERR_ARG, "bf_peer_copy: null pointer"); return BF_ERR_ARG; }
    if (bytes == 0) return BF_OK;
    BF_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return BF_OK;
}

extern "C" int bf_dev_alloc(size_t bytes, void **d_ptr)
{
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    if (!d_ptr || bytes == 0) { set_error(BF_ERR_ARG, "bf_dev_alloc: bad arguments"); return BF_ERR_ARG; }
    BF_CUDA(cudaMalloc(d_ptr, bytes));
    BF_CUDA(cudaMemset(*d_ptr, 0, bytes));
    return BF_OK;
}

extern "C" int bf_dev_free(void *d_ptr)
{
    clear_error();
    if (d_ptr) BF_CUDA(cudaFree(d_ptr));
    return BF_OK;
}

extern "C" int bf_ipc_export(void *d_ptr, unsigned char *handle64)
{
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!d_ptr || !handle64) { set_error(BF_ERR_ARG, "bf_ipc_export: null argument"); return BF_ERR_ARG; }
    cudaIpcMemHandle_t h;
    BF_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle64, &h, 64);
    return BF_OK;
}

extern "C" int bf_ipc_open(const unsigned char *handle64, void **d_ptr)
{
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    if (!d_ptr || !handle64) { set_error(BF_ERR_ARG, "bf_ipc_open: null argument"); return BF_ERR_ARG; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    BF_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return BF_OK;
}

extern "C" int bf_ipc_close(void *d_ptr)
{
    clear_error();
    if (d_ptr) BF_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return BF_OK;
}

extern "C" int bf_gather_signal(void *const *flag_arrays, int world, int rank, long long step, void *stream)
{
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    if (!flag_arrays || world < 1 || world > 8 || rank < 0 || rank >= world) {
        set_error(BF_ERR_ARG, "bf_gather_signal: bad arguments");
        return BF_ERR_ARG;
    }
    FlagPtrs fp{};
    for (int r = 0; r < world; r++) fp.p[r] = (long long *)flag_arrays[r];
    gather_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(fp, world, rank, step);
    BF_CHECK_LAUNCH();
    count_launch();
    return BF_OK;
}

extern "C" int bf_gather_wait(const void *d_my_flags, int world, long long step, int *d_timed_out, void *stream)
{
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    if (!d_my_flags || !d_timed_out || world < 1 || world > 8) {
        set_error(BF_ERR_ARG, "bf_gather_wait: bad arguments");
        return BF_ERR_ARG;
    }
    // ~200 ns per spin: give up after about 20 s
    gather_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const long long *)d_my_flags, world, step, 100000000LL,
                                                           d_timed_out);
    BF_CHECK_LAUNCH();
    count_launch();
    return BF_OK;
}
