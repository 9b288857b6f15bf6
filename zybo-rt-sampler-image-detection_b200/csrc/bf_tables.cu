// bf_tables.cu -- coefficient-table kernels: delay generator, lerp split, reductions.
#include <math.h>

#include "bf_common.cuh"

namespace bf {

// ---------------------------------------------------------------------------
// max of an int table (drives the zero-pad width of the tiled kernel's rows)
// ---------------------------------------------------------------------------
__global__ void max_i32_kernel(const int *__restrict__ v, size_t count, int *__restrict__ out)
{
    int best = INT_MIN;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (size_t)gridDim.x * blockDim.x)
        best = max(best, v[i]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, s));
    if ((threadIdx.x & 31) == 0) atomicMax(out, best);
}

int launch_max_abs_i32(const int *d, size_t count, int *h_max)
{
    State &S = state();
    int rc = S.d_scratch.ensure(256);
    if (rc) return rc;
    int *d_out = S.d_scratch.as<int>();
    int init = INT_MIN;
    BF_CUDA(cudaMemcpy(d_out, &init, sizeof(int), cudaMemcpyHostToDevice));
    int blocks = (int)((count + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    if (blocks < 1) blocks = 1;
    max_i32_kernel<<<blocks, 256>>>(d, count, d_out);
    BF_CHECK_LAUNCH();
    count_launch();
    BF_CUDA(cudaMemcpy(h_max, d_out, sizeof(int), cudaMemcpyDeviceToHost));
    return BF_OK;
}

// ---------------------------------------------------------------------------
// load_coefficients_lerp split (lerp_and_sum.c:139-153), one thread per entry:
//   frac  = (float) modf((double)delay, &ip)
//   h     = (float)(1.0 - (double)frac)        <- stored "reversed" weight
//   whole = (int) ip
// modf == trunc + exact subtraction, so this is bit-identical to libm.
// ---------------------------------------------------------------------------
__global__ void split_lerp_kernel(const float *__restrict__ delays, size_t count,
                                  int *__restrict__ whole, float *__restrict__ weight)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double x = (double)delays[i];
    const double ip = trunc(x);
    const float fr = __double2float_rn(__dsub_rn(x, ip));
    weight[i] = __double2float_rn(__dsub_rn(1.0, (double)fr));
    whole[i] = (int)ip;
}

int split_lerp_dev(const float *d_delays, size_t count, int *d_whole, float *d_weight,
                   cudaStream_t st)
{
    if (count == 0) return BF_OK;
    split_lerp_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(d_delays, count, d_whole,
                                                                       d_weight);
    BF_CHECK_LAUNCH();
    count_launch();
    return BF_OK;
}

// ---------------------------------------------------------------------------
// hybrid split (hybrid_convolve_and_sum.c:124-180).  The integer part is split
// on the device like the lerp one; the windowed-sinc taps need double sin/cos
// whose last-ulp behaviour differs between CUDA's and the host's libm, and the
// reference evaluates them on the host at load time -- so do we (load time, not
// on the per-frame path), to stay bit-identical with the reference's tables.
// ---------------------------------------------------------------------------
static void hybrid_taps_host(float *h, double delay, int T)
{
    const double PI_REF = 3.14159265359, eps = 1e-9;
    const double tau = 0.5 - delay + eps;
    double total = 0.0;
    for (int i = 0; i < T; i++) {
        double x = (double)i - ((double)T - 1.0) / 2.0 - tau;
        double v = sin(x * PI_REF) / (x * PI_REF);
        double nn = (double)(i * 2 - T + 1);
        double win = 0.42 + 0.5 * cos(PI_REF * nn / ((double)(T - 1)) + eps) +
                     0.08 * cos(2.0 * PI_REF * nn / ((double)(T - 1) + eps));
        v = v * win;
        total = total + v;
        h[i] = (float)v;
    }
    const float ft = (float)total;
    for (int i = 0; i < T; i++) h[i] = h[i] / ft;
}

int split_hybrid_host(const float *h_delays, size_t count, int *h_whole, float *h_taps, int T)
{
    for (size_t i = 0; i < count; i++) {
        double ip;
        const double fr = 1.0 - modf((double)h_delays[i], &ip);
        h_whole[i] = (int)ip;
        hybrid_taps_host(h_taps + i * T, fr, T);
    }
    return BF_OK;
}

// ---------------------------------------------------------------------------
// delay-table generator (directions.pyx:105-124), one CTA per look direction.
//   r      = sqrt((xs*xs + ys*ys) + z2)
//   raw_m  = (k * (xs*x_m + ys*y_m)) / r
//   out_m  = raw_m - min_m raw_m
// every operation individually rounded to float64, in NumPy's evaluation order
// (no FMA contraction), so the table equals the reference's bit for bit.
// ---------------------------------------------------------------------------
__global__ void delay_table_kernel(double k, const double *__restrict__ xs,
                                   const double *__restrict__ ys, int res_y, double z2,
                                   const double *__restrict__ mx, const double *__restrict__ my,
                                   int n, double *__restrict__ out_f64, int *__restrict__ out_i32,
                                   float *__restrict__ out_f32)
{
    __shared__ double red[32];
    const int d = blockIdx.x;
    const double x = xs[d / res_y], y = ys[d % res_y];
    const double r = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), z2));
    double best = INFINITY;
    for (int m = threadIdx.x; m < n; m += blockDim.x) {
        const double dot = __dadd_rn(__dmul_rn(x, mx[m]), __dmul_rn(y, my[m]));
        const double raw = __ddiv_rn(__dmul_rn(k, dot), r);
        best = fmin(best, raw);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) best = fmin(best, __shfl_xor_sync(0xffffffffu, best, s));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    const int nw = (blockDim.x + 31) / 32;
    best = red[0];
    for (int w = 1; w < nw; w++) best = fmin(best, red[w]);
    for (int m = threadIdx.x; m < n; m += blockDim.x) {
        const double dot = __dadd_rn(__dmul_rn(x, mx[m]), __dmul_rn(y, my[m]));
        const double raw = __ddiv_rn(__dmul_rn(k, dot), r);
        const double v = __dsub_rn(raw, best);
        const size_t e = (size_t)d * n + m;
        if (out_f64) out_f64[e] = v;
        if (out_i32) out_i32[e] = (int)v;                 // astype(int): truncation
        if (out_f32) out_f32[e] = __double2float_rn(v);   // np.float32(...)
    }
}

int delay_table_dev(double k, const double *d_xs, int res_x, const double *d_ys, int res_y,
                    double z2, const double *d_mx, const double *d_my, int n, double *d_f64,
                    int *d_i32, float *d_f32, cudaStream_t st)
{
    const int D = res_x * res_y;
    int threads = n >= 256 ? 256 : ((n + 31) / 32) * 32;
    if (threads < 32) threads = 32;
    delay_table_kernel<<<D, threads, 0, st>>>(k, d_xs, d_ys, res_y, z2, d_mx, d_my, n, d_f64, d_i32,
                                              d_f32);
    BF_CHECK_LAUNCH();
    count_launch();
    return BF_OK;
}

}  // namespace bf
