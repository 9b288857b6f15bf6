// bf_common.cuh -- shared declarations of libbf_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/bf_b200.h"

namespace bf {

// ---------------------------------------------------------------------------
// error plumbing: part-1 functions return void, so status is kept per thread
// ---------------------------------------------------------------------------
void set_error(int status, const char *fmt, ...);
void clear_error();
int last_status();
void count_launch(int n = 1);

#define BF_CUDA(expr)                                                              \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) {                                                   \
            bf::set_error(BF_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, \
                          cudaGetErrorString(_e));                                 \
            return BF_ERR_CUDA;                                                    \
        }                                                                          \
    } while (0)

#define BF_CHECK_LAUNCH()                                                          \
    do {                                                                           \
        cudaError_t _e = cudaGetLastError();                                       \
        if (_e != cudaSuccess) {                                                   \
            bf::set_error(BF_ERR_CUDA, "%s:%d kernel launch -> %s", __FILE__,      \
                          __LINE__, cudaGetErrorString(_e));                       \
            return BF_ERR_CUDA;                                                    \
        }                                                                          \
    } while (0)

// ---------------------------------------------------------------------------
// device buffer with explicit ownership
// ---------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need);   // grows (never shrinks); returns BF_OK / BF_ERR_CUDA
    void release();
    template <class T> T *as() const { return (T *)p; }
};

// ---------------------------------------------------------------------------
// coefficient tables (one per algorithm per process, as in the reference:
// pad_and_sum.c:28-29, lerp_and_sum.c:32-33, convolve_and_sum.c:43,
// hybrid_convolve_and_sum.c:44-45)
// ---------------------------------------------------------------------------
struct GroupTable {            // device-friendly re-layout, built lazily per (n, slice)
    DevBuf offs;               // uint4 [groups][n]: 8 x u16 byte offsets (+ uniform flag)
    DevBuf wts;                // float [groups][n][8] (lerp weights), lerp only
    int n = -1, d_begin = -1, d_count = -1, pad = -1, n_samples = -1;
    int groups = 0;
    int R = 0;                 // directions per group the layout was built for
};

struct Tables {
    // PAD
    DevBuf pad_whole;  size_t pad_count = 0;        // int32 [count]
    DevBuf pad2_whole; size_t pad2_count = 0;       // int32 [count], indexed by mic id
    DevBuf trunc_whole; size_t trunc_count = 0;     // api.c:1004 load_coefficients2
    // LERP
    DevBuf lerp_whole, lerp_weight; size_t lerp_count = 0;
    // FIR
    DevBuf fir_taps;   size_t fir_count = 0;        // float [count] = [D][n][T]
    // HYBRID
    DevBuf hyb_whole, hyb_taps; size_t hyb_count = 0;
    // max integer delay of each table (drives the zero-pad width of the smem rows)
    int pad_max = 0, lerp_max = 0, trunc_max = 0;
    GroupTable g_pad, g_lerp, g_trunc;
    uint64_t version = 0;                            // bumped on every load
};

struct State {
    bf_config cfg;
    Tables tab;
    // staging for the host-pointer (part 1) entry points
    DevBuf d_signals, d_image, d_mic_ids, d_out, d_scratch, d_diff;
    void *h_pinned = nullptr; size_t h_pinned_bytes = 0;
    bf_data_source_fn source = nullptr;
    int simple_kernel = 0;
    int exact_sum = 1;
    int sm_count = 0;
    int device = -1;
    DevBuf d_done_counter;       // last-CTA tickets of the fused gather kernel (4: overlapping steps rotate)
    int gather_overlap = 0;      // bf_gather_overlap
};

State &state();
int ensure_device();     // lazy CUDA init (after fork), fills sm_count
int ensure_pinned(size_t bytes);

// launchers implemented in the .cu files --------------------------------------
int launch_max_abs_i32(const int *d, size_t count, int *h_max);   // bf_tables.cu

// where the power of (frame f, direction d) goes: img[f*frame_stride + (d-d_origin)*dir_stride]
struct ImgLayout {
    long frame_stride; long dir_stride; int d_origin;
    // fused all-gather: the same element is also stored into these buffers (other GPUs' memory
    // mapped through CUDA IPC / NVLink peer access), same layout as the primary one
    float *peers[7]; int n_peers;
    // step flags of the fused gather (all optional): before its first peer store a warp waits until
    // flags_local[r] >= wait_seq for every rank r; the last CTA to finish publishes signal_seq into
    // flags_all[r][flag_rank] of every rank
    long long *flags_local; long long wait_seq;
    long long *flags_all[8]; int world, flag_rank; long long signal_seq;
    unsigned int *done_counter; int *timed_out;
    // launch with programmatic stream serialisation: the kernel's CTAs may take over SMs while the previous kernel
    // of the stream is still running its last tiles (consecutive steps must not touch the same output locations and
    // the inputs must be complete before the PREVIOUS launch: see bf_gather_overlap in bf_b200.h)
    int overlap;
};
int mimo_tiled(int algo, const float *d_sig, float *d_img, int frames, const int *d_mics, int n,
               int d_begin, int d_count, ImgLayout lay, cudaStream_t st);   // das_mimo.cu
bool fir_tiled_supported(int algo, int N, int T);                           // das_fir.cu
int fir_tiled(int algo, const float *d_sig, float *d_img, int frames, const int *d_mics, int n,
              int d_begin, int d_count, ImgLayout lay, cudaStream_t st);    // das_fir.cu
int mimo_simple(int algo, const float *d_sig, float *d_img, int frames, const int *d_mics, int n,
                int d_begin, int d_count, ImgLayout lay, cudaStream_t st);  // das_simple.cu
int miso_run(int algo, const float *d_sig, float *d_out, int blocks, const int *d_mics, int n,
             int offset, int by_mic_id, int scale, cudaStream_t st); // das_miso.cu
int miso_simple(int algo, const float *d_sig, float *d_out, int blocks, const int *d_mics, int n,
                int offset, int by_mic_id, int scale, cudaStream_t st);  // das_simple.cu
int split_hybrid_host(const float *h_delays, size_t count, int *h_whole, float *h_taps, int T);
int delay_table_dev(double k, const double *d_xs, int res_x, const double *d_ys, int res_y,
                    double z2, const double *d_mx, const double *d_my, int n, double *d_f64,
                    int *d_i32, float *d_f32, cudaStream_t st);         // bf_tables.cu
int split_lerp_dev(const float *d_delays, size_t count, int *d_whole, float *d_weight,
                   cudaStream_t st);                                // bf_tables.cu
int single_delay(int kind, const float *h_signal, const float *h_coef, float h, int pad,
                 float *h_out);                                     // das_simple.cu

}  // namespace bf

// ---------------------------------------------------------------------------
// PTX helpers (mbarrier + bulk async copy, the sm_90+/sm_100 TMA engine's 1-D form)
// ---------------------------------------------------------------------------
#ifdef __CUDACC__
namespace bfptx {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
                 : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 16-byte asynchronous copy global -> shared (LDGSTS): no register staging
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// global -> shared bulk copy completing on an mbarrier (SASS: UBLKCP).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

}  // namespace bfptx
#endif
