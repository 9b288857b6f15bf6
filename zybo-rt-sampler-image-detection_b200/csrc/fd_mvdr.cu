// fd_mvdr.cu -- frequency-domain MVDR (Capon) power map: SURVEY.md section 8 row a20.
//
// NOT in the reference (it only has frequency-domain DAS, fd_path.cu); parity for this file
// is "unpinned": the oracle is the float64 NumPy restatement in tests/test_gpu_mvdr.py, which
// follows the reference's conventions for the pieces the reference does define -- real FFT
// along samples, no window, no scaling (beam_forming_algorithm.py:31-33) and the steering
// phasor a_f(d)[m] = exp(-j k_f u[d,m]) (calc_phase_shift_cartesian.py:44-48).  The one
// reference-pinned check is on the covariance: Re(a^H R a) averaged over snapshots must equal
// the DAS power of fd_path.cu (tested).
//
//   X_k[f,m]  = rfft(snapshot_k[:,m])[lo:hi]                      K snapshots
//   R_f       = 1/K sum_k x_k x_k^H  + delta * tr(R_f)/M * I      (diagonal loading)
//   R_f       = L_f L_f^H            (Cholesky)
//   P(d)      = sum_f 1 / Re(a^H R_f^-1 a) = sum_f 1 / || L_f^-1 a_f(d) ||^2
//
// Conditioning: cond(R_f) can reach M/delta, so the spectra, the covariance, the
// factorisation and the triangular inverse are kept in float64 (a few GFLOP); only the
// steering contraction Y = L^-1 A (8*D*M^2*F flops, 8.8 TFLOP at BASELINE config C4) runs in
// reduced precision.  This file holds the first correct version of that contraction on the
// CUDA cores (fp32); the tcgen05 version is the round-2 item.
#include <math.h>

#include "bf_common.cuh"

namespace bf {

// shared with fd_path.cu through accessor functions ------------------------------------------
struct FdGeom { int n_mics, n_active, N, lo, hi, D; double fs, c; const double *u; const int *active; };
int fd_geometry(FdGeom *g);     // fd_path.cu
int mvdr_steer_tc(const float2 *d_linv, const double *d_u, int M, int F, int lo, double bin_hz, double inv_c,
                  int D, float *d_power, cudaStream_t st);     // fd_tc.cu
int mvdr_tc_prepare(const float2 *d_linv, int M, int F, int f0, int fc, cudaStream_t st);      // fd_tc.cu
int mvdr_tc_steer_only(const double *d_u, int M, int F, int lo, double bin_hz, double inv_c, int D, float *d_power,
                       cudaStream_t st);                                                      // fd_tc.cu
int mvdr_tc_operands(void **d_image, size_t *bytes_per_bin, void **d_binscale, int F);         // fd_tc.cu

struct MvdrState {
    DevBuf spec;      // double2 [K][F][M]
    DevBuf cov;       // double2 [F][M][M]  (row-major, Hermitian), later overwritten by L (lower)
    DevBuf linv;      // float2  [F][M][M]  L^-1, row-major, lower triangular
    DevBuf sig, out;
    int K = 0;
};
static MvdrState g_mv;
static float g_stage_ms[5] = {0, 0, 0, 0, 0};     // fft, covariance+loading, cholesky, inverse, steering

// ---- float64 real FFT, bins [lo,hi): same Stockham scheme as fd_rfft_kernel -----------------
__global__ void mvdr_rfft64_kernel(const float *__restrict__ sig, const int *__restrict__ active,
                                   int n_active, int n_mics, int N, int logN, int lo, int hi,
                                   double2 *__restrict__ spec)
{
    extern __shared__ double2 shd[];
    double2 *a = shd, *b = shd + N;
    const int m = blockIdx.x, k = blockIdx.y;
    const float *row = sig + ((size_t)k * n_mics + active[m]) * N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) a[i] = make_double2((double)row[i], 0.0);
    __syncthreads();
    int l = 1;
    for (int s = 0; s < logN; s++, l <<= 1) {
        const int half = N >> 1;
        for (int i = threadIdx.x; i < half; i += blockDim.x) {
            const int j = i / l, kk = i - j * l;
            const double2 x0 = a[j * l + kk], x1 = a[j * l + kk + half];
            double sn, cs;
            sincospi(-(double)kk / (double)l, &sn, &cs);
            const double2 t = make_double2(x1.x * cs - x1.y * sn, x1.x * sn + x1.y * cs);
            b[2 * j * l + kk] = make_double2(x0.x + t.x, x0.y + t.y);
            b[2 * j * l + kk + l] = make_double2(x0.x - t.x, x0.y - t.y);
        }
        __syncthreads();
        double2 *tmp = a; a = b; b = tmp;
    }
    const int F = hi - lo;
    for (int f = lo + threadIdx.x; f < hi; f += blockDim.x)
        spec[((size_t)k * F + (f - lo)) * n_active + m] = a[f];
}

// ---- covariance R_f[i][j] = 1/K sum_k X_k[f,i] conj(X_k[f,j]) -------------------------------
// A 16x16-thread CTA computes a 64x64 tile of R_f, every thread a 4x4 block; the
// snapshots of the tile's rows and columns go through shared memory 16 at a time.  Same
// ascending-k order per element; explicit fma (the library is built with --fmad=false).
__global__ void __launch_bounds__(256) mvdr_cov_tiled_kernel(const double2 *__restrict__ spec, int K, int F, int M,
                                                             double2 *__restrict__ cov, int f0)
{
    constexpr int TB = 64, KC = 16;
    __shared__ double2 xi[KC][TB], xj[KC][TB];
    const int f = f0 + blockIdx.z;
    const int i0 = blockIdx.y * TB, j0 = blockIdx.x * TB;
    const int ty = threadIdx.y, tx = threadIdx.x, tid = ty * 16 + tx;
    double re[4][4], im[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) { re[a][b] = 0.0; im[a][b] = 0.0; }
    for (int k0 = 0; k0 < K; k0 += KC) {
        __syncthreads();
        for (int t = tid; t < KC * TB; t += 256) {
            const int kk = t / TB, c = t - kk * TB;
            const bool kv = k0 + kk < K;
            const size_t base = ((size_t)(k0 + kk) * F + f) * M;
            xi[kk][c] = (kv && i0 + c < M) ? spec[base + i0 + c] : make_double2(0.0, 0.0);
            xj[kk][c] = (kv && j0 + c < M) ? spec[base + j0 + c] : make_double2(0.0, 0.0);
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < KC; kk++) {
            double2 a[4], b[4];
#pragma unroll
            for (int q = 0; q < 4; q++) { a[q] = xi[kk][ty + 16 * q]; b[q] = xj[kk][tx + 16 * q]; }
#pragma unroll
            for (int p = 0; p < 4; p++)
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    re[p][q] = fma(a[p].x, b[q].x, fma(a[p].y, b[q].y, re[p][q]));      // a * conj(b)
                    im[p][q] = fma(a[p].y, b[q].x, fma(-a[p].x, b[q].y, im[p][q]));
                }
        }
    }
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int i = i0 + ty + 16 * p, j = j0 + tx + 16 * q;
            if (i < M && j < M) cov[((size_t)f * M + i) * M + j] = make_double2(re[p][q] / K, im[p][q] / K);
        }
}

// diagonal loading: R_ii += delta * tr(R)/M
__global__ void mvdr_load_kernel(double2 *__restrict__ cov, int M, double delta)
{
    __shared__ double red[32];
    double2 *R = cov + (size_t)blockIdx.x * M * M;
    double tr = 0.0;
    for (int i = threadIdx.x; i < M; i += blockDim.x) tr += R[(size_t)i * M + i].x;
    for (int s = 16; s > 0; s >>= 1) tr += __shfl_xor_sync(0xffffffffu, tr, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tr;
    __syncthreads();
    tr = 0.0;
    for (int w = 0; w < (int)((blockDim.x + 31) / 32); w++) tr += red[w];
    const double add = delta * tr / M;
    for (int i = threadIdx.x; i < M; i += blockDim.x) R[(size_t)i * M + i].x += add;
}

// ---- Cholesky R = L L^H, one CTA per bin; thread i owns row i of L.  L is stored TRANSPOSED in
//      the upper triangle of the row-major array (element L[i][k] lives at [k][i]) so that the
//      inner loop over k reads consecutive addresses across threads; R's own upper triangle
//      supplies R[i][j] = conj(R[j][i]) with the same access pattern. ------------------------------
__global__ void mvdr_chol_kernel(double2 *__restrict__ cov, int M, int *__restrict__ fail)
{
    double2 *R = cov + (size_t)blockIdx.x * M * M;
    __shared__ double2 colj[1024];      // L[j][0..j) of the pivot row (M <= 1024)
    __shared__ double pivot;
    for (int j = 0; j < M; j++) {
        for (int k = threadIdx.x; k < j; k += blockDim.x) colj[k] = R[(size_t)k * M + j];
        __syncthreads();
        for (int i = j + threadIdx.x; i < M; i += blockDim.x) {
            const double2 rji = R[(size_t)j * M + i];          // R[i][j] = conj(R[j][i])
            double sx = rji.x, sy = -rji.y;
            // s -= sum_k L[i][k] * conj(L[j][k]); four independent partial sums
            double px[4] = {0, 0, 0, 0}, py[4] = {0, 0, 0, 0};
            int k = 0;
            for (; k + 4 <= j; k += 4) {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const double2 a = R[(size_t)(k + u) * M + i], b = colj[k + u];
                    px[u] += a.x * b.x + a.y * b.y;
                    py[u] += a.y * b.x - a.x * b.y;
                }
            }
            for (; k < j; k++) {
                const double2 a = R[(size_t)k * M + i], b = colj[k];
                px[0] += a.x * b.x + a.y * b.y;
                py[0] += a.y * b.x - a.x * b.y;
            }
            sx -= (px[0] + px[1]) + (px[2] + px[3]);
            sy -= (py[0] + py[1]) + (py[2] + py[3]);
            if (i == j) {
                if (!(sx > 0.0)) { atomicExch(fail, 1); sx = 1.0; }
                pivot = sqrt(sx);
            }
            R[(size_t)j * M + i] = make_double2(sx, sy);       // un-normalised L[i][j] at [j][i]
        }
        __syncthreads();
        const double d = pivot;
        for (int i = j + threadIdx.x; i < M; i += blockDim.x) {
            const double2 v = R[(size_t)j * M + i];
            R[(size_t)j * M + i] = (i == j) ? make_double2(d, 0.0) : make_double2(v.x / d, v.y / d);
        }
        __syncthreads();
    }
}

// ---- Z = L^-1 (lower triangular), forward substitution row by row; thread c owns column c.
//      L[i][k] (stored at [k][i]) is the same address for every thread of a step (broadcast),
//      Z[k][c] is consecutive across threads.  Output float2 row-major [i][c]. --------------------
__global__ void mvdr_trinv_kernel(const double2 *__restrict__ chol, int M, float2 *__restrict__ linv,
                                  double2 *__restrict__ work)
{
    const double2 *Lt = chol + (size_t)blockIdx.x * M * M;
    double2 *Z = work + (size_t)blockIdx.x * M * M;          // float64 copy, row-major [i][c]
    float2 *Zf = linv + (size_t)blockIdx.x * M * M;
    __shared__ double2 rowi[1024];                           // L[i][0..i]
    for (int i = 0; i < M; i++) {
        for (int k = threadIdx.x; k <= i; k += blockDim.x) rowi[k] = Lt[(size_t)k * M + i];
        __syncthreads();
        const double dii = rowi[i].x;
        for (int c = threadIdx.x; c < M; c += blockDim.x) {
            double2 s = make_double2(0.0, 0.0);
            if (c <= i) {
                double px[2] = {c == i ? 1.0 : 0.0, 0.0}, py[2] = {0.0, 0.0};
                int k = c;
                for (; k + 2 <= i; k += 2) {                   // s = delta_ic - sum_k L[i][k] * Z[k][c]
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const double2 a = rowi[k + u], z = Z[(size_t)(k + u) * M + c];
                        px[u] -= a.x * z.x - a.y * z.y;
                        py[u] -= a.x * z.y + a.y * z.x;
                    }
                }
                for (; k < i; k++) {
                    const double2 a = rowi[k], z = Z[(size_t)k * M + c];
                    px[0] -= a.x * z.x - a.y * z.y;
                    py[0] -= a.x * z.y + a.y * z.x;
                }
                s.x = (px[0] + px[1]) / dii;
                s.y = (py[0] + py[1]) / dii;
            }
            Z[(size_t)i * M + c] = s;
            Zf[(size_t)i * M + c] = make_float2((float)s.x, (float)s.y);
        }
        __syncthreads();
    }
}

// ---- blocked versions (M <= 256, one thread per row / column) ---------------------------------
// The column-by-column kernels above re-read the whole factor for every column (M^3/3 x 16 B per
// bin from L2: 45 GB per C4 map, L2-bandwidth bound).  Here a thread streams its row of L (its
// column of Z) once per PANEL of NB columns (rows) and keeps NB complex accumulators in registers;
// the NB pivot rows are broadcast from shared memory.  Same arithmetic, different summation order.
static constexpr int kNB = 8;       // panel width
static constexpr int kKC = 64;      // k-chunk staged in shared memory
// The streamed operand (a thread's own row of L / column of Z, 16 bytes per k) comes from L2 at ~500 cycles a load;
// four loads in flight per thread (what the registers allow) leave the FP64 pipe idle most of the time.  Each thread
// therefore prefetches its own values through a private lane of a shared-memory ring with cp.async -- kSub k's per
// group, two groups in flight, no registers held and no barrier (a thread only reads what it copied itself).
static constexpr int kSub = 8;
static constexpr int kRingBytes = 2 * kSub * 256 * (int)sizeof(double2);      // 64 KB per CTA, dynamic
__device__ __forceinline__ void cp_async16_cg(void *dst_smem, const void *src_gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(bfptx::smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Register budget (measured, C4: 512 bins): __launch_bounds__(256, 1) lets ptxas keep the unrolled row loads in flight
// (128 registers, 2 CTAs per SM): Cholesky 1.75 ms, inverse 1.83 ms.  Capped at 64 registers for 4 CTAs per SM (all 512
// bins in one wave) they take 1.82 / 2.43 ms, at the 80-88 registers ptxas picks unprompted 2.08 / 2.31 ms: these loops
// are bound by load latency per thread, not by the number of resident CTAs.
__global__ void __launch_bounds__(256, 2) mvdr_chol_blocked_kernel(double2 *__restrict__ cov, int M, int *__restrict__ fail)
{
    double2 *R = cov + (size_t)blockIdx.x * M * M;
    extern __shared__ __align__(16) unsigned char ring_raw[];
    double2 *ring = (double2 *)ring_raw;       // [2][kSub][256]: this thread's prefetched L[i][k]
    __shared__ double2 pan[kNB][kKC];          // L[j0+c][k0 .. k0+KC)
    __shared__ double2 lrow[kNB];              // L[j0+c'][j] of the column being finished
    __shared__ double pivot;
    const int i = threadIdx.x;
    for (int j0 = 0; j0 < M; j0 += kNB) {
        const int nb = min(kNB, M - j0);
        const bool live = i >= j0 && i < M;
        double sx[kNB], sy[kNB];
#pragma unroll
        for (int c = 0; c < kNB; c++) {
            sx[c] = 0.0; sy[c] = 0.0;
            if (live && c < nb) {
                const double2 a = R[(size_t)(j0 + c) * M + i];       // A[i][j0+c] = conj(A[j0+c][i])
                sx[c] = a.x; sy[c] = -a.y;
            }
        }
        // s[c] -= sum_{k < j0} L[i][k] * conj(L[j0+c][k])
        for (int k0 = 0; k0 < j0; k0 += kKC) {
            const int kc = min(kKC, j0 - k0);
            __syncthreads();
            for (int t = threadIdx.x; t < kNB * kKC; t += blockDim.x) {
                const int c = t / kKC, kk = t - c * kKC;
                pan[c][kk] = (c < nb && kk < kc) ? R[(size_t)(k0 + kk) * M + (j0 + c)] : make_double2(0.0, 0.0);
            }
            __syncthreads();
            if (live) {
                // kc is a multiple of kSub (j0 and kKC are multiples of 8)
                const int nsub = kc / kSub;
                auto issue = [&](int sb) {
#pragma unroll
                    for (int kk = 0; kk < kSub; kk++)
                        cp_async16_cg(&ring[((sb & 1) * kSub + kk) * 256 + i], &R[(size_t)(k0 + sb * kSub + kk) * M + i]);
                    cp_async_commit();
                };
                issue(0);
                for (int sb = 0; sb < nsub; sb++) {
                    if (sb + 1 < nsub) { issue(sb + 1); cp_async_wait<1>(); }
                    else cp_async_wait<0>();
#pragma unroll
                    for (int kk = 0; kk < kSub; kk++) {
                        const double2 a = ring[((sb & 1) * kSub + kk) * 256 + i];
#pragma unroll
                        for (int c = 0; c < kNB; c++) {
                            const double2 b = pan[c][sb * kSub + kk];
                            sx[c] = fma(-a.x, b.x, fma(-a.y, b.y, sx[c]));
                            sy[c] = fma(-a.y, b.x, fma(a.x, b.y, sy[c]));
                        }
                    }
                }
            }
        }
        // factor the panel: column j = j0 + c
#pragma unroll
        for (int c = 0; c < kNB; c++) {
            if (c < nb) {
                const int j = j0 + c;
                __syncthreads();
                if (i == j) {
                    double d2 = sx[c];
                    if (!(d2 > 0.0)) { atomicExch(fail, 1); d2 = 1.0; }
                    pivot = sqrt(d2);
                }
                __syncthreads();
                const double d = pivot;
                double lx = 0.0, ly = 0.0;
                if (live && i >= j) {
                    lx = (i == j) ? d : sx[c] / d;
                    ly = (i == j) ? 0.0 : sy[c] / d;
                    R[(size_t)j * M + i] = make_double2(lx, ly);     // L[i][j] at [j][i]
                    if (i < j0 + nb) lrow[i - j0] = make_double2(lx, ly);
                }
                __syncthreads();
                if (live && i > j) {
#pragma unroll
                    for (int c2 = 0; c2 < kNB; c2++) {
                        if (c2 > c && c2 < nb) {
                            const double2 b = lrow[c2];              // L[j0+c2][j]
                            sx[c2] = fma(-lx, b.x, fma(-ly, b.y, sx[c2]));
                            sy[c2] = fma(-ly, b.x, fma(lx, b.y, sy[c2]));
                        }
                    }
                }
            }
        }
    }
}

// Z = L^-1, row panels of NB: thread c owns column c, acc[r] = sum_{k < i0} L[i0+r][k] Z[k][c]
__global__ void __launch_bounds__(256, 2) mvdr_trinv_blocked_kernel(const double2 *__restrict__ chol, int M,
                                                                  float2 *__restrict__ linv, double2 *__restrict__ work)
{
    const double2 *Lt = chol + (size_t)blockIdx.x * M * M;
    double2 *Z = work + (size_t)blockIdx.x * M * M;
    float2 *Zf = linv + (size_t)blockIdx.x * M * M;
    extern __shared__ __align__(16) unsigned char ring_raw[];
    double2 *ring = (double2 *)ring_raw;       // [2][kSub][256]: this thread's prefetched Z[k][c]
    __shared__ double2 pan[kNB][kKC];          // L[i0+r][k0 .. k0+KC)
    __shared__ double2 tri[kNB][kNB];          // L[i0+r][i0+r']
    const int c = threadIdx.x;
    const int cw = c & ~31;                    // first column of this warp: Z[k][c] = 0 for k < c
    for (int i0 = 0; i0 < M; i0 += kNB) {
        const int nb = min(kNB, M - i0);
        double ax[kNB], ay[kNB];
#pragma unroll
        for (int r = 0; r < kNB; r++) { ax[r] = 0.0; ay[r] = 0.0; }
        for (int k0 = 0; k0 < i0; k0 += kKC) {
            const int kc = min(kKC, i0 - k0);
            __syncthreads();
            for (int t = threadIdx.x; t < kNB * kKC; t += blockDim.x) {
                const int r = t / kKC, kk = t - r * kKC;
                pan[r][kk] = (r < nb && kk < kc) ? Lt[(size_t)(k0 + kk) * M + (i0 + r)] : make_double2(0.0, 0.0);
            }
            __syncthreads();
            if (c < M && cw < k0 + kc) {
                // Z[k][c] = 0 for k < cw (cw is a multiple of 32, hence of kSub): start at the warp's first column
                const int sb0 = max(0, cw - k0) / kSub, nsub = kc / kSub;
                auto issue = [&](int sb) {
#pragma unroll
                    for (int kk = 0; kk < kSub; kk++)
                        cp_async16_cg(&ring[((sb & 1) * kSub + kk) * 256 + c], &Z[(size_t)(k0 + sb * kSub + kk) * M + c]);
                    cp_async_commit();
                };
                issue(sb0);
                for (int sb = sb0; sb < nsub; sb++) {
                    if (sb + 1 < nsub) { issue(sb + 1); cp_async_wait<1>(); }
                    else cp_async_wait<0>();
#pragma unroll
                    for (int kk = 0; kk < kSub; kk++) {
                        const double2 z = ring[((sb & 1) * kSub + kk) * 256 + c];
#pragma unroll
                        for (int r = 0; r < kNB; r++) {
                            const double2 a = pan[r][sb * kSub + kk];
                            ax[r] = fma(a.x, z.x, fma(-a.y, z.y, ax[r]));
                            ay[r] = fma(a.x, z.y, fma(a.y, z.x, ay[r]));
                        }
                    }
                }
            }
        }
        __syncthreads();
        for (int t = threadIdx.x; t < kNB * kNB; t += blockDim.x) {
            const int r = t / kNB, r2 = t - r * kNB;
            tri[r][r2] = (r < nb && r2 <= r) ? Lt[(size_t)(i0 + r2) * M + (i0 + r)] : make_double2(0.0, 0.0);
        }
        __syncthreads();
        if (c < M) {
            double zx[kNB], zy[kNB];
#pragma unroll
            for (int r = 0; r < kNB; r++) {
                zx[r] = 0.0; zy[r] = 0.0;
                if (r < nb) {
                    const int i = i0 + r;
                    if (c <= i) {
                        double px = (c == i ? 1.0 : 0.0) - ax[r], py = -ay[r];
#pragma unroll
                        for (int r2 = 0; r2 < kNB; r2++) {
                            if (r2 < r) {
                                const double2 a = tri[r][r2];
                                px -= a.x * zx[r2] - a.y * zy[r2];
                                py -= a.x * zy[r2] + a.y * zx[r2];
                            }
                        }
                        const double dii = tri[r][r].x;
                        zx[r] = px / dii;
                        zy[r] = py / dii;
                    }
                    Z[(size_t)i * M + c] = make_double2(zx[r], zy[r]);
                    Zf[(size_t)i * M + c] = make_float2((float)zx[r], (float)zy[r]);
                }
            }
        }
    }
}

// ---- steering: P[d] = sum_f 1 / || L_f^-1 a_f(d) ||^2, CUDA-core fp32 --------------------------
// CTA = TD directions; per bin the TD x M phasors are generated once into shared memory
// ([m][d], float2), then rows of L^-1 stream through shared memory in chunks of RC rows.
template <int TD, int RC>
__global__ void __launch_bounds__(TD) mvdr_steer_kernel(const float2 *__restrict__ linv,
                                                        const double *__restrict__ u, int M, int F,
                                                        int lo, double bin_hz, double inv_c, int D,
                                                        float *__restrict__ power)
{
    extern __shared__ float2 shm[];
    float2 *ph = shm;                        // [M][TD]
    float2 *rows = shm + (size_t)M * TD;     // [RC][M]
    const int d = blockIdx.x * TD + threadIdx.x;
    const bool ok = d < D;
    const double *ud = u + (size_t)(ok ? d : 0) * M;
    float total = 0.0f;
    for (int f = 0; f < F; f++) {
        const double turns_per_u = (double)(lo + f) * bin_hz * inv_c;
        for (int m = 0; m < M; m++) {
            const double turns = turns_per_u * ud[m];
            const float fr = (float)(turns - rint(turns));
            float sn, cs;
            sincospif(-2.0f * fr, &sn, &cs);
            ph[(size_t)m * TD + threadIdx.x] = make_float2(cs, sn);
        }
        const float2 *Lf = linv + (size_t)f * M * M;
        float q = 0.0f;
        for (int i0 = 0; i0 < M; i0 += RC) {
            __syncthreads();
            const int rc = min(RC, M - i0);
            for (int e = threadIdx.x; e < rc * M; e += TD) rows[e] = Lf[(size_t)i0 * M + e];
            __syncthreads();
            for (int r = 0; r < rc; r++) {
                const int i = i0 + r;
                float yr = 0.0f, yi = 0.0f;
                const float2 *Lr = rows + (size_t)r * M;
                for (int j = 0; j <= i; j++) {                 // y_i = sum_{j<=i} Linv[i][j] a_j
                    const float2 l = Lr[j], a = ph[(size_t)j * TD + threadIdx.x];
                    yr = fmaf(l.x, a.x, fmaf(-l.y, a.y, yr));
                    yi = fmaf(l.x, a.y, fmaf(l.y, a.x, yi));
                }
                q = fmaf(yr, yr, fmaf(yi, yi, q));
            }
        }
        total += 1.0f / q;
        __syncthreads();
    }
    if (ok) power[d] = total;
}

static int ilog2e(int n) { int l = 0; while ((1 << l) < n) l++; return (1 << l) == n ? l : -1; }

// Stages 1-4 for bins [f0, f0 + fc) of the band: float64 FFT of every snapshot (all bins: it is one pass per
// channel), covariance + loading, Cholesky, triangular inverse -> S.linv[f].  Event i..i+4 time the stages.
static int mvdr_factor(const float *d_snap, int K, double delta, int f0, int fc, cudaStream_t st, cudaEvent_t *ev,
                       int *h_fail)
{
    FdGeom G;
    int rc = fd_geometry(&G);
    if (rc) return rc;
    const int M = G.n_active, F = G.hi - G.lo;
    if (M > 1024) { set_error(BF_ERR_CONFIG, "mvdr: at most 1024 microphones"); return BF_ERR_CONFIG; }
    if (f0 < 0 || fc < 1 || f0 + fc > F) { set_error(BF_ERR_ARG, "mvdr: bin range [%d,+%d) outside the %d-bin band", f0, fc, F); return BF_ERR_ARG; }
    MvdrState &S = g_mv;
    if ((rc = S.spec.ensure((size_t)K * F * M * sizeof(double2)))) return rc;
    if ((rc = S.cov.ensure((size_t)F * M * M * sizeof(double2)))) return rc;
    if ((rc = S.linv.ensure((size_t)F * M * M * sizeof(float2)))) return rc;
    static DevBuf work, fail, cov_copy;
    if ((rc = work.ensure((size_t)F * M * M * sizeof(double2)))) return rc;
    if ((rc = fail.ensure(sizeof(int)))) return rc;
    if ((rc = cov_copy.ensure((size_t)F * M * M * sizeof(double2)))) return rc;
    cudaMemsetAsync(fail.p, 0, sizeof(int), st);
    const size_t mm = (size_t)M * M;
    double2 *cov = S.cov.as<double2>() + (size_t)f0 * mm;
    cudaEventRecord(ev[0], st);
    const int threads = G.N / 2 < 256 ? (G.N / 2 < 32 ? 32 : G.N / 2) : 256;
    mvdr_rfft64_kernel<<<dim3(M, K), threads, 2 * G.N * sizeof(double2), st>>>(
        d_snap, G.active, M, G.n_mics, G.N, ilog2e(G.N), G.lo, G.hi, S.spec.as<double2>());
    cudaEventRecord(ev[1], st);
    mvdr_cov_tiled_kernel<<<dim3((M + 63) / 64, (M + 63) / 64, fc), dim3(16, 16), 0, st>>>(
        S.spec.as<double2>(), K, F, M, S.cov.as<double2>(), f0);
    mvdr_load_kernel<<<fc, 256, 0, st>>>(cov, M, delta);
    cudaEventRecord(ev[2], st);
    S.K = K;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error(BF_ERR_CUDA, "mvdr: %s", cudaGetErrorString(e)); return BF_ERR_CUDA; }
    // keep a copy of the loaded covariance for bf_fd_get_covariance (the factorisation overwrites it)
    cudaMemcpyAsync(cov_copy.as<double2>() + (size_t)f0 * mm, cov, (size_t)fc * mm * sizeof(double2), cudaMemcpyDeviceToDevice, st);
    // blocked float64 factorisation / inverse (one thread per row); BF_MVDR_BLOCKED=0 selects the
    // column-by-column kernels
    const bool blocked = M <= 256 && !(getenv("BF_MVDR_BLOCKED") && atoi(getenv("BF_MVDR_BLOCKED")) == 0);
    const int mt = (M + 31) / 32 * 32;
    float2 *linv = S.linv.as<float2>() + (size_t)f0 * mm;
    double2 *wk = work.as<double2>() + (size_t)f0 * mm;
    if (blocked) {
        BF_CUDA(cudaFuncSetAttribute(mvdr_chol_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingBytes));
        mvdr_chol_blocked_kernel<<<fc, mt, kRingBytes, st>>>(cov, M, fail.as<int>());
    }
    else mvdr_chol_kernel<<<fc, 256, 0, st>>>(cov, M, fail.as<int>());
    cudaEventRecord(ev[3], st);
    if (blocked) {
        BF_CUDA(cudaFuncSetAttribute(mvdr_trinv_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingBytes));
        mvdr_trinv_blocked_kernel<<<fc, mt, kRingBytes, st>>>(cov, M, linv, wk);
    }
    else mvdr_trinv_kernel<<<fc, 256, 0, st>>>(cov, M, linv, wk);
    cudaEventRecord(ev[4], st);
    // restore the covariance for inspection, then report a failed factorisation
    cudaMemcpyAsync(cov, cov_copy.as<double2>() + (size_t)f0 * mm, (size_t)fc * mm * sizeof(double2), cudaMemcpyDeviceToDevice, st);
    e = cudaMemcpyAsync(h_fail, fail.p, sizeof(int), cudaMemcpyDeviceToHost, st);
    count_launch(5);
    if (e != cudaSuccess) { set_error(BF_ERR_CUDA, "mvdr: %s", cudaGetErrorString(e)); return BF_ERR_CUDA; }
    return BF_OK;
}

// d_count < 0: all directions; otherwise only directions [d_begin, d_begin + d_count) are steered (power[d - d_begin]):
// the covariance, its factor and the inverse are computed in full on every rank of a direction-sharded run
// (0.5 % of the flops), the steering contraction -- all the rest -- is what shards.  (bf_fd_mvdr_factor_dev /
// bf_fd_mvdr_steer_dev split the two so that the float64 stages can be sharded by bins as well.)
int mvdr_dev(const float *d_snap, float *d_power, int K, double delta, cudaStream_t st, int d_begin = 0, int d_count = -1)
{
    FdGeom G;
    int rc = fd_geometry(&G);
    if (rc) return rc;
    if (d_count >= 0) {
        if (d_begin < 0 || d_count < 1 || d_begin + d_count > G.D) {
            set_error(BF_ERR_ARG, "mvdr: direction slice [%d,+%d) outside the %d-direction grid", d_begin, d_count, G.D);
            return BF_ERR_ARG;
        }
        G.u += (size_t)d_begin * G.n_active;
        G.D = d_count;
    }
    const int M = G.n_active, F = G.hi - G.lo;
    static cudaEvent_t ev[6] = {nullptr};
    if (!ev[0]) for (int i = 0; i < 6; i++) cudaEventCreate(&ev[i]);
    static int *h_fail = nullptr;
    if (!h_fail) { if (cudaMallocHost(&h_fail, sizeof(int)) != cudaSuccess) { set_error(BF_ERR_CUDA, "mvdr: pinned flag"); return BF_ERR_CUDA; } }
    *h_fail = 0;
    if ((rc = mvdr_factor(d_snap, K, delta, 0, F, st, ev, h_fail))) return rc;
    MvdrState &S = g_mv;
    const double bin_hz = (double)(int)((int)G.fs / 2) / (double)(G.N / 2);
    // steering contraction: tcgen05 tensor-core kernel (fd_tc.cu) for 256 microphones, CUDA-core
    // fp32 kernel otherwise (BF_MVDR_TC=0 forces the latter)
    const bool use_tc = M == 256 && !(getenv("BF_MVDR_TC") && atoi(getenv("BF_MVDR_TC")) == 0);
    if (use_tc) {
        if ((rc = mvdr_steer_tc(S.linv.as<float2>(), G.u, M, F, G.lo, bin_hz, 1.0 / G.c, G.D, d_power, st))) return rc;
    } else {
        constexpr int TD = 64, RC = 8;
        const size_t smem = ((size_t)M * TD + (size_t)RC * M) * sizeof(float2);
        cudaFuncSetAttribute(mvdr_steer_kernel<TD, RC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        mvdr_steer_kernel<TD, RC><<<(G.D + TD - 1) / TD, TD, smem, st>>>(S.linv.as<float2>(), G.u, M, F, G.lo,
                                                                         bin_hz, 1.0 / G.c, G.D, d_power);
    }
    cudaEventRecord(ev[5], st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) for (int i = 0; i < 5; i++) cudaEventElapsedTime(&g_stage_ms[i], ev[i], ev[i + 1]);
    count_launch(1);
    if (e != cudaSuccess) { set_error(BF_ERR_CUDA, "mvdr: %s", cudaGetErrorString(e)); return BF_ERR_CUDA; }
    if (*h_fail) { set_error(BF_ERR_CONFIG, "mvdr: covariance not positive definite (increase loading)"); return BF_ERR_CONFIG; }
    return BF_OK;
}

// ---- two-phase form for runs that shard the float64 stages by BINS (lib/sharded.py:fd_mvdr_sharded_bins) ----
int mvdr_factor_dev(const float *d_snap, int K, double delta, int f0, int fc, cudaStream_t st)
{
    FdGeom G;
    int rc = fd_geometry(&G);
    if (rc) return rc;
    const int M = G.n_active, F = G.hi - G.lo;
    if (M != 256) { set_error(BF_ERR_CONFIG, "bf_fd_mvdr_factor_dev: the operand images exist for 256 microphones only"); return BF_ERR_CONFIG; }
    static cudaEvent_t ev[5] = {nullptr};
    if (!ev[0]) for (int i = 0; i < 5; i++) cudaEventCreate(&ev[i]);
    static int *h_fail = nullptr;
    if (!h_fail) { if (cudaMallocHost(&h_fail, sizeof(int)) != cudaSuccess) { set_error(BF_ERR_CUDA, "mvdr: pinned flag"); return BF_ERR_CUDA; } }
    *h_fail = 0;
    if ((rc = mvdr_factor(d_snap, K, delta, f0, fc, st, ev, h_fail))) return rc;
    if ((rc = mvdr_tc_prepare(g_mv.linv.as<float2>(), M, F, f0, fc, st))) return rc;
    cudaError_t e = cudaStreamSynchronize(st);          // the failure flag is read here: one sync per map
    if (e != cudaSuccess) { set_error(BF_ERR_CUDA, "mvdr: %s", cudaGetErrorString(e)); return BF_ERR_CUDA; }
    if (*h_fail) { set_error(BF_ERR_CONFIG, "mvdr: covariance not positive definite (increase loading)"); return BF_ERR_CONFIG; }
    return BF_OK;
}

int mvdr_steer_dev(float *d_power, int d_begin, int d_count, cudaStream_t st)
{
    FdGeom G;
    int rc = fd_geometry(&G);
    if (rc) return rc;
    if (d_begin < 0 || d_count < 1 || d_begin + d_count > G.D) {
        set_error(BF_ERR_ARG, "mvdr: direction slice [%d,+%d) outside the %d-direction grid", d_begin, d_count, G.D);
        return BF_ERR_ARG;
    }
    const double bin_hz = (double)(int)((int)G.fs / 2) / (double)(G.N / 2);
    return mvdr_tc_steer_only(G.u + (size_t)d_begin * G.n_active, G.n_active, G.hi - G.lo, G.lo, bin_hz, 1.0 / G.c,
                              d_count, d_power, st);
}

}  // namespace bf

using namespace bf;

extern "C" {

int bf_fd_mvdr(const float *snapshots, float *power, int K, double loading)
{
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    FdGeom G;
    if ((rc = fd_geometry(&G))) return rc;
    if (!snapshots || !power || K < 1 || !(loading >= 0.0)) { set_error(BF_ERR_ARG, "bf_fd_mvdr: bad arguments"); return BF_ERR_ARG; }
    MvdrState &S = g_mv;
    const size_t sb = (size_t)K * G.n_mics * G.N * sizeof(float);
    if ((rc = S.sig.ensure(sb))) return rc;
    if ((rc = S.out.ensure((size_t)G.D * sizeof(float)))) return rc;
    BF_CUDA(cudaMemcpy(S.sig.p, snapshots, sb, cudaMemcpyHostToDevice));
    if ((rc = mvdr_dev(S.sig.as<float>(), S.out.as<float>(), K, loading, 0))) return rc;
    BF_CUDA(cudaMemcpy(power, S.out.p, (size_t)G.D * sizeof(float), cudaMemcpyDeviceToHost));
    return BF_OK;
}

int bf_fd_mvdr_dev(const float *d_snapshots, float *d_power, int K, double loading, void *stream)
{
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    if (!d_snapshots || !d_power || K < 1) { set_error(BF_ERR_ARG, "bf_fd_mvdr_dev: bad arguments"); return BF_ERR_ARG; }
    return mvdr_dev(d_snapshots, d_power, K, loading, (cudaStream_t)stream);
}

int bf_fd_mvdr_dev_slice(const float *d_snapshots, float *d_power, int K, double loading, int d_begin, int d_count,
                         void *stream)
{
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    if (!d_snapshots || !d_power || K < 1 || d_count < 1) { set_error(BF_ERR_ARG, "bf_fd_mvdr_dev_slice: bad arguments"); return BF_ERR_ARG; }
    return mvdr_dev(d_snapshots, d_power, K, loading, (cudaStream_t)stream, d_begin, d_count);
}

int bf_fd_mvdr_factor_dev(const float *d_snapshots, int K, double loading, int f_begin, int f_count, void *stream)
{
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    if (!d_snapshots || K < 1 || !(loading >= 0.0)) { set_error(BF_ERR_ARG, "bf_fd_mvdr_factor_dev: bad arguments"); return BF_ERR_ARG; }
    return mvdr_factor_dev(d_snapshots, K, loading, f_begin, f_count, (cudaStream_t)stream);
}

int bf_fd_mvdr_operands(void **d_image, size_t *bytes_per_bin, void **d_binscale)
{
    clear_error();
    FdGeom G;
    int rc = fd_geometry(&G);
    if (rc) return rc;
    return mvdr_tc_operands(d_image, bytes_per_bin, d_binscale, G.hi - G.lo);
}

int bf_fd_mvdr_steer_dev(float *d_power, int d_begin, int d_count, void *stream)
{
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    if (!d_power) { set_error(BF_ERR_ARG, "bf_fd_mvdr_steer_dev: bad arguments"); return BF_ERR_ARG; }
    return mvdr_steer_dev(d_power, d_begin, d_count, (cudaStream_t)stream);
}

// device time of each stage of the last MVDR call, milliseconds:
// [0] float64 FFT  [1] covariance + loading  [2] Cholesky  [3] triangular inverse  [4] steering
int bf_fd_mvdr_timings(float *ms5)
{
    if (!ms5) return BF_ERR_ARG;
    for (int i = 0; i < 5; i++) ms5[i] = g_stage_ms[i];
    return BF_OK;
}

// loaded covariance of the last bf_fd_mvdr call: HOST double [F][M][M][2] (re, im)
int bf_fd_get_covariance(double *cov, size_t count)
{
    clear_error();
    FdGeom G;
    int rc = fd_geometry(&G);
    if (rc) return rc;
    const size_t have = (size_t)(G.hi - G.lo) * G.n_active * G.n_active;
    if (!cov || count > have || g_mv.K == 0) { set_error(BF_ERR_NOT_LOADED, "no covariance (run bf_fd_mvdr first)"); return BF_ERR_NOT_LOADED; }
    BF_CUDA(cudaMemcpy(cov, g_mv.cov.p, count * sizeof(double2), cudaMemcpyDeviceToHost));
    return BF_OK;
}

}  // extern "C"
