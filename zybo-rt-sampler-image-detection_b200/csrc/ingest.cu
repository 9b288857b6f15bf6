// ingest.cu -- wire format -> sample buffer (SURVEY.md section 8f "next" #1).
//
// The Zybo sends one UDP datagram per sample instant (receiver.h:51-59: u16 frequency,
// i8 n_arrays, i8 protocol_ver, i32 counter, i32 stream[N_MICROPHONES]).  The reference's
// receiver (receiver.c:94-151) turns N_SAMPLES consecutive payloads into the [mic][sample]
// float buffer the beamformer reads: per 8x8 array, even rows in order, odd rows reversed
// (boustrophedon wiring), value = (float)((double)v / NORM_FACTOR).  The odd-row index of the
// reference is `row + COLUMNS - x` (receiver.c:140), one past the intended `COLUMNS-1-x`: it
// reads the first element of the next row (and one int past the payload on the very last
// row).  quirk != 0 reproduces it (out-of-payload reads give 0), quirk == 0 applies the fix.
//
// Doing this on the device lets recordings be stored as the raw int32 payloads and fuses the
// transpose + conversion in front of the beamformer; it is a pure HBM-bound transpose:
// 32x32 shared-memory tiles, coalesced on both sides.  Channels listed in `zero_mask` are
// cleared (the hard-coded 122-channel mask of get_data(), api.c:835-858).
#include <math.h>

#include "bf_common.cuh"

namespace bf {

// starts == NULL: frame f is the block of N consecutive datagrams f*N .. f*N+N-1 (what the receiver hands over);
// otherwise frame f is the N-sample WINDOW that begins at datagram starts[f] of a stream of `total` datagrams
// (batch replay: video frame k looks at floor(k*fs/fps), any offset); datagrams beyond `total` read as silence.
__global__ void ingest_kernel(const int *__restrict__ stream, float *__restrict__ out, int N,
                              int n_mics_total, int n_channels, int rows, int cols, double inv_norm,
                              int quirk, const unsigned char *__restrict__ zero_mask,
                              const long *__restrict__ starts, long total)
{
    __shared__ float tile[32][33];
    const int frame = blockIdx.z;
    const int s0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
    const long first = starts ? starts[frame] : (long)frame * N;
    const int *in = stream + (size_t)first * n_mics_total;
    const long avail = starts ? total - first : (long)N;          // datagrams readable from `first` on
    float *o = out + (size_t)frame * n_mics_total * N;
    // load: thread (x = channel within tile, y = step within tile, 4 passes of 8 steps)
    const int s = s0 + threadIdx.x;
    int src = -1;
    if (s < n_channels) {
        const int per = rows * cols;
        const int a = s / per, r = (s - a * per) / cols, x = s - a * per - r * cols;
        const int row = a * per + r * cols;
        src = (r & 1) ? row + (quirk ? cols - x : cols - 1 - x) : row + x;
        if (src >= n_mics_total) src = -1;              // past the payload: 0
    }
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int step = t0 + k;
        float v = 0.0f;
        if (step < N && step < avail && first + step >= 0 && src >= 0) v = __double2float_rn(__dmul_rn((double)in[(size_t)step * n_mics_total + src], inv_norm));
        tile[k][threadIdx.x] = v;
    }
    __syncthreads();
    // store: thread x = step within tile, rows = channels
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int sc = s0 + k, step = t0 + threadIdx.x;
        if (sc < n_channels && step < N) {
            const bool zero = zero_mask && zero_mask[sc];
            o[(size_t)sc * N + step] = zero ? 0.0f : tile[threadIdx.x][k];
        }
    }
}

int ingest_dev(const int *d_stream, float *d_out, int frames, int n_arrays, int rows, int cols, double norm,
               int quirk, const unsigned char *d_zero_mask, cudaStream_t st, const long *d_starts = nullptr,
               long total = 0)
{
    State &S = state();
    const int N = S.cfg.n_samples, M = S.cfg.n_microphones;
    const int n_channels = n_arrays * rows * cols;
    if (n_arrays < 1 || rows < 1 || cols < 1 || n_channels > M || !(norm > 0.0)) {
        set_error(BF_ERR_ARG, "ingest: %d arrays of %dx%d do not fit %d channels", n_arrays, rows, cols, M);
        return BF_ERR_ARG;
    }
    // norm is a power of two in the reference (2^24): multiplying by 1/norm is exact division
    dim3 grid((N + 31) / 32, (n_channels + 31) / 32, frames), block(32, 8);
    ingest_kernel<<<grid, block, 0, st>>>(d_stream, d_out, N, M, n_channels, rows, cols, 1.0 / norm, quirk,
                                          d_zero_mask, d_starts, total);
    BF_CHECK_LAUNCH();
    count_launch();
    return BF_OK;
}

}  // namespace bf

using namespace bf;

extern "C" int bf_ingest_dev(const int *d_stream, float *d_signals, int frames, int n_arrays, int rows, int cols,
                             double norm, int quirk, const unsigned char *d_zero_mask, void *stream)
{
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    if (!d_stream || !d_signals || frames < 1) { set_error(BF_ERR_ARG, "bf_ingest_dev: bad arguments"); return BF_ERR_ARG; }
    double frac = norm;
    int e = 0;
    frac = frexp(norm, &e);
    if (frac != 0.5) { set_error(BF_ERR_ARG, "bf_ingest_dev: NORM_FACTOR must be a power of two (got %g)", norm); return BF_ERR_ARG; }
    return ingest_dev(d_stream, d_signals, frames, n_arrays, rows, cols, norm, quirk, d_zero_mask, (cudaStream_t)stream);
}

extern "C" int bf_ingest_windows_dev(const int *d_stream, long total_datagrams, const long *d_starts, float *d_signals,
                                     int frames, int n_arrays, int rows, int cols, double norm, int quirk,
                                     const unsigned char *d_zero_mask, void *stream)
{
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    if (!d_stream || !d_starts || !d_signals || frames < 1 || total_datagrams < 1) {
        set_error(BF_ERR_ARG, "bf_ingest_windows_dev: bad arguments");
        return BF_ERR_ARG;
    }
    int e = 0;
    if (frexp(norm, &e) != 0.5) { set_error(BF_ERR_ARG, "bf_ingest_windows_dev: NORM_FACTOR must be a power of two (got %g)", norm); return BF_ERR_ARG; }
    return ingest_dev(d_stream, d_signals, frames, n_arrays, rows, cols, norm, quirk, d_zero_mask, (cudaStream_t)stream,
                      d_starts, total_datagrams);
}

// ---------------------------------------------------------------------------------------------
// window gather for batch replay (BASELINE config C5): a recording is stored channel-major,
// float [n_microphones][samples] (the .npy format of PC/record.py:28-46); every video frame k is
// the N_SAMPLES window starting at sample starts[k] (= floor(k*fs/fps)).  Output is the frame
// batch the beamformer kernels take: float [frames][n_microphones][n_samples].  Pure copy.
// ---------------------------------------------------------------------------------------------
namespace bf {
__global__ void window_kernel(const float *__restrict__ rec, long samples, const long *__restrict__ starts,
                              int n_mics, int N, float *__restrict__ out)
{
    const int f = blockIdx.y, m = blockIdx.x;
    const long s0 = starts[f];
    const float *src = rec + (size_t)m * samples + s0;
    float *dst = out + ((size_t)f * n_mics + m) * N;
    for (int t = threadIdx.x; t < N; t += blockDim.x) dst[t] = (s0 + t >= 0 && s0 + t < samples) ? src[t] : 0.0f;   // outside the recording: silence
}
}  // namespace bf

extern "C" int bf_window_dev(const float *d_recording, long samples, const long *d_starts, int frames,
                             float *d_frames, void *stream)
{
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    State &S = state();
    if (!d_recording || !d_starts || !d_frames || frames < 1 || samples < 1) { set_error(BF_ERR_ARG, "bf_window_dev: bad arguments"); return BF_ERR_ARG; }
    const int N = S.cfg.n_samples, M = S.cfg.n_microphones;
    bf::window_kernel<<<dim3(M, frames), N < 256 ? N : 256, 0, (cudaStream_t)stream>>>(d_recording, samples, d_starts, M, N, d_frames);
    BF_CHECK_LAUNCH();
    count_launch();
    return BF_OK;
}
