// fd_path.cu -- frequency-domain delay-and-sum (the reference's NumPy "FFT beamforming"
// backend, PC/application/realtime_scripts/):
//
//   X[f,m]   = rfft(signal[:,m])[lo:hi]                     beam_forming_algorithm.py:31-32
//   P[x,y]   = sum_f | sum_m X[f,m] * exp(-j k_f u[x,y,m]) |^2     :33-34, :52-56
//              k_f = 2*pi*f/c,  u = (x_s x_m + y_s y_m)/r_s       calc_phase_shift_cartesian.py:40-48
//   heat     = P / max(P), or all zeros when max(P) < threshold            :57-61
//
// The reference materialises exp(j phi) for every (bin, mic, direction) at import time
// (complex128, 94 x 256 x 13 x 13 = 65 MB; 34 GB at 512 bins x 32k directions).  Here the
// phasors are generated on the fly: the phase is reduced in float64 (it reaches ~70 rad at
// 18 kHz, where an fp32 argument alone would cost 4e-6 of accuracy) and sincospi is taken
// in fp32 on the reduced argument.
//
// Kernels:  fd_rfft_kernel   one CTA per (frame, channel): radix-2 Stockham real-input FFT
//                            in shared memory, writes only the bins [lo, hi)
//           fd_steer_kernel  one thread per direction, CTA tile of directions x all bins,
//                            spectrum staged in shared memory per bin
//           fd_norm_kernel   max-reduce + normalise / threshold
#include <math.h>

#include "bf_common.cuh"

namespace bf {

struct FdState {
    int n_mics = 0, n_active = 0, N = 0, lo = 0, hi = 0, D = 0;
    double fs = 0, c = 0;
    DevBuf u;          // double [D][n_active]  path term u = (xs*xm + ys*ym)/r
    DevBuf active;     // int [n_active]
    DevBuf spec;       // float2 [frames][F][n_active]
    DevBuf power;      // float [frames][D]
    DevBuf sig;        // staging for host signals
    DevBuf red;        // reductions
};
static FdState g_fd;
static uint64_t g_fd_generation = 0;       // bumped by every bf_fd_setup: keys caches derived from the geometry
uint64_t fd_geometry_generation() { return g_fd_generation; }

// ---- geometry: u[d][m], float64, the reference's operation order -----------------------
__global__ void fd_geometry_kernel(const double *__restrict__ xs, const double *__restrict__ ys,
                                   int res_y, double z2, const double *__restrict__ mx,
                                   const double *__restrict__ my, const int *__restrict__ active,
                                   int n_active, double *__restrict__ u)
{
    const int d = blockIdx.x;
    const double x = xs[d / res_y], y = ys[d % res_y];
    const double r = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), z2));
    for (int m = threadIdx.x; m < n_active; m += blockDim.x) {
        const int mic = active[m];
        const double dot = __dadd_rn(__dmul_rn(x, mx[mic]), __dmul_rn(y, my[mic]));
        u[(size_t)d * n_active + m] = __ddiv_rn(dot, r);
    }
}

// ---- real FFT of one channel (N = 2^p <= 2048), bins [lo, hi) ---------------------------
// Complex Stockham autosort radix-2 on the N real samples (imaginary part 0); fp32, twiddles
// from sincospif (exact argument k/N).  Output layout spec[frame][f - lo][m].
__global__ void fd_rfft_kernel(const float *__restrict__ sig, const int *__restrict__ active,
                               int n_active, int n_mics, int N, int logN, int lo, int hi,
                               float2 *__restrict__ spec)
{
    extern __shared__ float2 sh[];                  // 2 * N
    float2 *a = sh, *b = sh + N;
    const int m = blockIdx.x, frame = blockIdx.y;
    const float *row = sig + ((size_t)frame * n_mics + active[m]) * N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) a[i] = make_float2(row[i], 0.0f);
    __syncthreads();
    // Stockham: at stage s (half-size l = 2^s), n/2 butterflies
    int l = 1;
    for (int s = 0; s < logN; s++, l <<= 1) {
        const int half = N >> 1;
        for (int i = threadIdx.x; i < half; i += blockDim.x) {
            const int j = i / l, k = i - j * l;          // block j (of N/(2l)), position k < l
            const float2 x0 = a[j * l + k];
            const float2 x1 = a[j * l + k + half];
            float sn, cs;
            sincospif(-(float)k / (float)l, &sn, &cs);   // w = exp(-j*pi*k/l)
            const float2 t = make_float2(x1.x * cs - x1.y * sn, x1.x * sn + x1.y * cs);
            b[2 * j * l + k] = make_float2(x0.x + t.x, x0.y + t.y);
            b[2 * j * l + k + l] = make_float2(x0.x - t.x, x0.y - t.y);
        }
        __syncthreads();
        float2 *tmp = a; a = b; b = tmp;
    }
    const int F = hi - lo;
    for (int f = lo + threadIdx.x; f < hi; f += blockDim.x)
        spec[((size_t)frame * F + (f - lo)) * n_active + m] = a[f];
}

// ---- steering: P[d] = sum_f | sum_m X[f,m] exp(-j 2 pi f_hz u[d,m] / c) |^2 -------------
// f_hz of bin f follows the reference: linspace(0, int(fs/2), N/2+1)[f]
// (calc_phase_shift_cartesian.py:34) -- NOT f*fs/N.
template <int TD>
__global__ void __launch_bounds__(TD) fd_steer_kernel(const float2 *__restrict__ spec,
                                                      const double *__restrict__ u, int n_active,
                                                      int F, int lo, double bin_hz, double inv_c,
                                                      int D, float *__restrict__ power)
{
    extern __shared__ float2 sx[];                  // [n_active] spectrum of the current bin
    const int frame = blockIdx.y;
    const int d = blockIdx.x * TD + threadIdx.x;
    const bool ok = d < D;
    const double *ud = u + (size_t)(ok ? d : 0) * n_active;
    float total = 0.0f;
    for (int f = 0; f < F; f++) {
        __syncthreads();
        for (int m = threadIdx.x; m < n_active; m += TD)
            sx[m] = spec[((size_t)frame * F + f) * n_active + m];
        __syncthreads();
        const double turns_per_u = (double)(lo + f) * bin_hz * inv_c;     // f_hz / c  [1/m]
        float re = 0.0f, im = 0.0f;
        for (int m = 0; m < n_active; m++) {
            // phase = -2*pi*turns; reduce to (-0.5, 0.5] turns in fp64, sincospi in fp32
            const double turns = turns_per_u * ud[m];
            const float fr = (float)(turns - rint(turns));
            float sn, cs;
            sincospif(-2.0f * fr, &sn, &cs);
            const float2 x = sx[m];
            re = fmaf(x.x, cs, fmaf(-x.y, sn, re));
            im = fmaf(x.x, sn, fmaf(x.y, cs, im));
        }
        total += re * re + im * im;
    }
    if (ok) power[(size_t)frame * D + d] = total;
}

// ---- normalise: heat = P/max(P) or 0 when max < threshold (per frame) --------------------
__global__ void fd_norm_kernel(float *__restrict__ power, int D, float threshold, int normalise)
{
    __shared__ float red[32];
    float *p = power + (size_t)blockIdx.x * D;
    float mx = 0.0f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) mx = fmaxf(mx, p[i]);
    for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < (int)((blockDim.x + 31) / 32); w++) mx = fmaxf(mx, red[w]);
    if (!normalise) return;
    const bool quiet = mx < threshold;
    for (int i = threadIdx.x; i < D; i += blockDim.x) p[i] = quiet ? 0.0f : __fdiv_rn(p[i], mx);
}

// accessor for fd_mvdr.cu
struct FdGeom { int n_mics, n_active, N, lo, hi, D; double fs, c; const double *u; const int *active; };
int fd_geometry(FdGeom *g)
{
    if (g_fd.D == 0) { set_error(BF_ERR_NOT_LOADED, "fd: bf_fd_setup() has not been called"); return BF_ERR_NOT_LOADED; }
    g->n_mics = g_fd.n_mics; g->n_active = g_fd.n_active; g->N = g_fd.N; g->lo = g_fd.lo; g->hi = g_fd.hi;
    g->D = g_fd.D; g->fs = g_fd.fs; g->c = g_fd.c; g->u = g_fd.u.as<double>(); g->active = g_fd.active.as<int>();
    return BF_OK;
}

static int ilog2_exact(int n)
{
    int l = 0;
    while ((1 << l) < n) l++;
    return (1 << l) == n ? l : -1;
}

int fd_setup(int n_mics, int n_samples, double fs, double c, int lo_bin, int hi_bin,
             const double *x_scan, int res_x, const double *y_scan, int res_y, double z,
             const double *mic_x, const double *mic_y, const int *active, int n_active)
{
    int rc = ensure_device();
    if (rc) return rc;
    if (ilog2_exact(n_samples) < 1 || n_samples > 2048 || lo_bin < 0 || hi_bin > n_samples / 2 + 1 ||
        lo_bin >= hi_bin || n_active < 1 || n_active > n_mics || res_x < 1 || res_y < 1) {
        set_error(BF_ERR_CONFIG, "fd_setup: N=%d (power of two <= 2048) bins [%d,%d) mics %d/%d", n_samples,
                  lo_bin, hi_bin, n_active, n_mics);
        return BF_ERR_CONFIG;
    }
    FdState &G = g_fd;
    G.n_mics = n_mics; G.n_active = n_active; G.N = n_samples; G.lo = lo_bin; G.hi = hi_bin;
    G.D = res_x * res_y; G.fs = fs; G.c = c;
    g_fd_generation++;
    DevBuf tmp;
    const size_t nd = (size_t)res_x + res_y + 2 * (size_t)n_mics;
    if ((rc = tmp.ensure(nd * sizeof(double)))) return rc;
    double *d_xs = tmp.as<double>(), *d_ys = d_xs + res_x, *d_mx = d_ys + res_y, *d_my = d_mx + n_mics;
    if ((rc = G.active.ensure((size_t)n_active * sizeof(int)))) { tmp.release(); return rc; }
    if ((rc = G.u.ensure((size_t)G.D * n_active * sizeof(double)))) { tmp.release(); return rc; }
    cudaMemcpy(d_xs, x_scan, res_x * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(d_ys, y_scan, res_y * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(d_mx, mic_x, n_mics * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(d_my, mic_y, n_mics * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(G.active.p, active, n_active * sizeof(int), cudaMemcpyHostToDevice);
    const double z2 = z * z;                    // config.Z**2: plain Python float arithmetic
    fd_geometry_kernel<<<G.D, 128>>>(d_xs, d_ys, res_y, z2, d_mx, d_my, G.active.as<int>(), n_active,
                                     G.u.as<double>());
    cudaError_t e = cudaDeviceSynchronize();
    tmp.release();
    if (e != cudaSuccess) { set_error(BF_ERR_CUDA, "fd_setup: %s", cudaGetErrorString(e)); return BF_ERR_CUDA; }
    count_launch();
    return BF_OK;
}

// d_count < 0: the whole grid, normalised per the reference; otherwise the un-normalised power of directions
// [d_begin, d_begin + d_count) as float [frames][d_count] (direction-sharded runs normalise after the gather).
int fd_das_dev(const float *d_signals, float *d_heat, int frames, float threshold, int normalise,
               cudaStream_t st, int d_begin = 0, int d_count = -1)
{
    FdState &G = g_fd;
    if (G.D == 0) { set_error(BF_ERR_NOT_LOADED, "fd: bf_fd_setup() has not been called"); return BF_ERR_NOT_LOADED; }
    const bool slice = d_count >= 0;
    if (slice && (d_begin < 0 || d_count < 1 || d_begin + d_count > G.D)) {
        set_error(BF_ERR_ARG, "fd: direction slice [%d,+%d) outside the %d-direction grid", d_begin, d_count, G.D);
        return BF_ERR_ARG;
    }
    const int Dn = slice ? d_count : G.D;
    const double *u = G.u.as<double>() + (size_t)(slice ? d_begin : 0) * G.n_active;
    const int F = G.hi - G.lo;
    int rc = G.spec.ensure((size_t)frames * F * G.n_active * sizeof(float2));
    if (rc) return rc;
    const int threads = G.N / 2 < 256 ? (G.N / 2 < 32 ? 32 : G.N / 2) : 256;
    fd_rfft_kernel<<<dim3(G.n_active, frames), threads, 2 * G.N * sizeof(float2), st>>>(
        d_signals, G.active.as<int>(), G.n_active, G.n_mics, G.N, ilog2_exact(G.N), G.lo, G.hi,
        G.spec.as<float2>());
    BF_CHECK_LAUNCH();
    constexpr int TD = 64;
    // bin spacing of the reference's frequency axis: int(fs/2) / (N/2)
    const double bin_hz = (double)(int)((int)G.fs / 2) / (double)(G.N / 2);
    fd_steer_kernel<TD><<<dim3((Dn + TD - 1) / TD, frames), TD, G.n_active * sizeof(float2), st>>>(
        G.spec.as<float2>(), u, G.n_active, F, G.lo, bin_hz, 1.0 / G.c, Dn, d_heat);
    BF_CHECK_LAUNCH();
    if (!slice) {
        fd_norm_kernel<<<frames, 256, 0, st>>>(d_heat, G.D, threshold, normalise);
        BF_CHECK_LAUNCH();
    }
    count_launch(slice ? 2 : 3);
    return BF_OK;
}

int fd_normalise_dev(float *d_heat, int frames, float threshold, int normalise, cudaStream_t st)
{
    FdState &G = g_fd;
    if (G.D == 0) { set_error(BF_ERR_NOT_LOADED, "fd: bf_fd_setup() has not been called"); return BF_ERR_NOT_LOADED; }
    fd_norm_kernel<<<frames, 256, 0, st>>>(d_heat, G.D, threshold, normalise);
    BF_CHECK_LAUNCH();
    count_launch();
    return BF_OK;
}

int fd_das_host(const float *signals_mn, float *heat, int frames, float threshold, int normalise)
{
    FdState &G = g_fd;
    int rc = ensure_device();
    if (rc) return rc;
    if (G.D == 0) { set_error(BF_ERR_NOT_LOADED, "fd: bf_fd_setup() has not been called"); return BF_ERR_NOT_LOADED; }
    const size_t sb = (size_t)frames * G.n_mics * G.N * sizeof(float), hb = (size_t)frames * G.D * sizeof(float);
    if ((rc = G.sig.ensure(sb))) return rc;
    if ((rc = G.power.ensure(hb))) return rc;
    BF_CUDA(cudaMemcpy(G.sig.p, signals_mn, sb, cudaMemcpyHostToDevice));
    if ((rc = fd_das_dev(G.sig.as<float>(), G.power.as<float>(), frames, threshold, normalise, 0))) return rc;
    BF_CUDA(cudaMemcpy(heat, G.power.p, hb, cudaMemcpyDeviceToHost));
    return BF_OK;
}

}  // namespace bf

using namespace bf;

extern "C" {

int bf_fd_setup(int n_mics, int n_samples, double fs, double c, int lo_bin, int hi_bin,
                const double *x_scan, int res_x, const double *y_scan, int res_y, double z,
                const double *mic_x, const double *mic_y, const int *active, int n_active)
{
    clear_error();
    if (!x_scan || !y_scan || !mic_x || !mic_y || !active) { set_error(BF_ERR_ARG, "bf_fd_setup: null pointer"); return BF_ERR_ARG; }
    return fd_setup(n_mics, n_samples, fs, c, lo_bin, hi_bin, x_scan, res_x, y_scan, res_y, z, mic_x, mic_y,
                    active, n_active);
}

int bf_fd_das(const float *signals, float *heatmap, int frames, float threshold, int normalise)
{
    clear_error();
    if (!signals || !heatmap || frames < 1) { set_error(BF_ERR_ARG, "bf_fd_das: bad arguments"); return BF_ERR_ARG; }
    return fd_das_host(signals, heatmap, frames, threshold, normalise);
}

int bf_fd_das_dev(const float *d_signals, float *d_heatmap, int frames, float threshold, int normalise,
                  void *stream)
{
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    if (!d_signals || !d_heatmap || frames < 1) { set_error(BF_ERR_ARG, "bf_fd_das_dev: bad arguments"); return BF_ERR_ARG; }
    return fd_das_dev(d_signals, d_heatmap, frames, threshold, normalise, (cudaStream_t)stream);
}

int bf_fd_das_dev_slice(const float *d_signals, float *d_power, int frames, int d_begin, int d_count, void *stream)
{
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    if (!d_signals || !d_power || frames < 1) { set_error(BF_ERR_ARG, "bf_fd_das_dev_slice: bad arguments"); return BF_ERR_ARG; }
    return fd_das_dev(d_signals, d_power, frames, 0.0f, 0, (cudaStream_t)stream, d_begin, d_count);
}

int bf_fd_normalise_dev(float *d_heatmap, int frames, float threshold, int normalise, void *stream)
{
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    if (!d_heatmap || frames < 1) { set_error(BF_ERR_ARG, "bf_fd_normalise_dev: bad arguments"); return BF_ERR_ARG; }
    return fd_normalise_dev(d_heatmap, frames, threshold, normalise, (cudaStream_t)stream);
}

}  // extern "C"
