// heatmap.cu -- power map -> colour overlay, peak and confidence (SURVEY.md section 8f "next" #2).
//
// The step right after the beamformer in the reference (PC/src/visual.py):
//   calculate_heatmap             130-171  clip(1e-12), log10, -log10(min), /max, >= amount,
//                                          ((p-amount)/amount)**exponent, jet LUT, flipped store,
//                                          cv2.resize(INTER_LINEAR) to the window
//   calculate_heatmap_fft         173-205  the linear variant (image/max, fixed 0.5 and **2)
//   find_power_center             293-322  5x5 Gaussian (sigma 1), >= 95 %-of-max mask,
//                                          cube-weighted centroid, arg-max fall-back
//   sensorfusiondecider.get_entropy  PC/sensorfusion/decider.py:16-24
// Doing it on the device removes the D2H of raw maps in batch replay (config C5) and lets one
// launch post-process a whole batch of frames.
//
// Arithmetic: every float32 step of the NumPy code is one IEEE float32 operation here
// (library built with --fmad=false); log10 and the integer powers are evaluated in float64
// and rounded once (= a correctly rounded float32 log10f / powf; NumPy's own SIMD versions are
// 1-2 ulp off that, so colour indices can differ by one step on isolated pixels -- see
// oracle/heatmap_np.py).  The resize is OpenCV's 11-bit fixed-point bilinear scheme, integer
// only and bit-identical to cv2.resize 4.13.  One CTA post-processes one frame with the map held
// in shared memory; reductions are fixed trees (deterministic).
#include <math.h>
#include <string.h>

#include <mutex>

#include "bf_common.cuh"

namespace bf {

static constexpr int kHeatThreads = 1024;

struct HeatParams {
    const float *maps;
    long frame_stride;
    int X, Y;
    float threshold, amount;
    int exponent, log_scale;
    const unsigned char *lut;     // [256][3]
    unsigned char *small;         // [frames][Y][X][3]
    short *index;                 // [frames][X][Y] or null
    bf_heat_info *info;           // [frames]
};

template <class T, class Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T identity, void *scratch)
{
    T *s = (T *)scratch;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) s[warp] = v;
    __syncthreads();
    T r = (threadIdx.x < nw) ? s[threadIdx.x] : identity;
    if (warp == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r = op(r, __shfl_xor_sync(0xffffffffu, r, o));
        if (lane == 0) s[0] = r;
    }
    __syncthreads();
    r = s[0];
    __syncthreads();
    return r;
}

__device__ __forceinline__ int reflect101(int i, int n)
{
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

// cv2.getGaussianKernel(5, 1.0, CV_32F)
#define BF_G0 0.054488684982061386f
#define BF_G1 0.24420134723186493f
#define BF_G2 0.40261995792388916f

__device__ __forceinline__ float tap5(float p0, float p1, float p2, float p3, float p4)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(p2, BF_G2), __fmul_rn(__fadd_rn(p1, p3), BF_G1)),
                     __fmul_rn(__fadd_rn(p0, p4), BF_G0));
}

// separable 5x5 blur of the clipped map at pixel (x, y): row pass along y (axis 1), then along x
__device__ __forceinline__ float smooth_at(const float *s_map, int X, int Y, int x, int y)
{
    int c[5];
#pragma unroll
    for (int j = 0; j < 5; j++) c[j] = reflect101(y + j - 2, Y);
    float r[5];
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const float *row = s_map + reflect101(x + j - 2, X) * Y;
        r[j] = tap5(fmaxf(row[c[0]], 1e-12f), fmaxf(row[c[1]], 1e-12f), fmaxf(row[c[2]], 1e-12f),
                    fmaxf(row[c[3]], 1e-12f), fmaxf(row[c[4]], 1e-12f));
    }
    return tap5(r[0], r[1], r[2], r[3], r[4]);
}

__device__ __forceinline__ float log10_f32(float v) { return (float)log10((double)v); }

__global__ void __launch_bounds__(kHeatThreads, 1) heat_kernel(const HeatParams p)
{
    extern __shared__ __align__(16) float s_map[];
    __shared__ double scratch[32];
    const int f = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const int X = p.X, Y = p.Y, P = X * Y;
    const float *map = p.maps + (long)f * p.frame_stride;
    unsigned char *small = p.small + (size_t)f * P * 3;
    short *index = p.index ? p.index + (size_t)f * P : nullptr;

    auto fmax_op = [](float a, float b) { return fmaxf(a, b); };
    auto fmin_op = [](float a, float b) { return fminf(a, b); };
    auto dadd_op = [](double a, double b) { return a + b; };
    auto iadd_op = [](int a, int b) { return a + b; };
    auto umax_op = [](unsigned long long a, unsigned long long b) { return a > b ? a : b; };

    // ---- 1. load, max of the raw map, min of the clipped map (visual.py:143-147) ----
    float vmax = -INFINITY, smin = INFINITY, lmax = -INFINITY;
    for (int i = tid; i < P; i += T) {
        const float v = map[i];
        s_map[i] = v;
        vmax = fmaxf(vmax, v);
        smin = fminf(smin, fmaxf(v, 1e-12f));
    }
    const float mx = block_reduce(vmax, fmax_op, -INFINITY, scratch);
    const float mn = block_reduce(smin, fmin_op, INFINITY, scratch);
    const bool gate = mx > p.threshold;

    // ---- 2. colour index (visual.py:149-166 / 176-200) ----
    float lmin = 0.0f, span = 0.0f;
    if (gate && p.log_scale) {
        lmin = log10_f32(mn);
        for (int i = tid; i < P; i += T) lmax = fmaxf(lmax, log10_f32(fmaxf(s_map[i], 1e-12f)));
        span = __fsub_rn(block_reduce(lmax, fmax_op, -INFINITY, scratch), lmin);   // np.max(img) after img -= log10(min)
    }
    int painted = 0;
    for (int i = tid; i < P; i += T) {
        int idx = -1;
        if (gate) {
            const float v = s_map[i];
            const float pl = p.log_scale ? __fdiv_rn(__fsub_rn(log10_f32(fmaxf(v, 1e-12f)), lmin), span)
                                         : __fdiv_rn(v, mx);
            if (pl >= p.amount) {
                const double q = (double)__fdiv_rn(__fsub_rn(pl, p.amount), p.amount);
                double r = 1.0;
                for (int e = 0; e < p.exponent; e++) r *= q;
                idx = (int)__fmul_rn(255.0f, (float)r);
                idx = idx < 0 ? 0 : (idx > 255 ? 255 : idx);
                painted++;
            }
        }
        const int x = i / Y, y = i - x * Y;
        unsigned char *o = small + ((size_t)(Y - 1 - y) * X + (X - 1 - x)) * 3;
        if (idx >= 0) {
            const unsigned char *c = p.lut + idx * 3;
            o[0] = c[0]; o[1] = c[1]; o[2] = c[2];
        } else {
            o[0] = 0; o[1] = 0; o[2] = 0;
        }
        if (index) index[i] = (short)idx;
    }
    const int n_painted = block_reduce(painted, iadd_op, 0, scratch);

    // ---- 3. find_power_center on the clipped map (visual.py:293-322) ----
    float sm_max = -INFINITY;
    unsigned long long best = 0ull;
    for (int i = tid; i < P; i += T) {
        const int x = i / Y, y = i - x * Y;
        const float s = smooth_at(s_map, X, Y, x, y);
        sm_max = fmaxf(sm_max, s);
        const unsigned long long key = ((unsigned long long)__float_as_uint(s) << 32) | (0xffffffffu - (unsigned)i);
        best = key > best ? key : best;
    }
    const float smax = block_reduce(sm_max, fmax_op, -INFINITY, scratch);
    best = block_reduce(best, umax_op, 0ull, scratch);
    const float thr = __fmul_rn(smax, 0.95f);
    double tot = 0.0, sc = 0.0, sr = 0.0;
    int cnt = 0;
    for (int i = tid; i < P; i += T) {
        const int x = i / Y, y = i - x * Y;
        const float s = smooth_at(s_map, X, Y, x, y);
        if (s >= thr) {
            const double w = (double)(float)((double)s * (double)s * (double)s);
            tot += w;
            sc += (double)y * w;
            sr += (double)x * w;
            cnt++;
        }
    }
    tot = block_reduce(tot, dadd_op, 0.0, scratch);
    sc = block_reduce(sc, dadd_op, 0.0, scratch);
    sr = block_reduce(sr, dadd_op, 0.0, scratch);
    cnt = block_reduce(cnt, iadd_op, 0, scratch);

    if (tid == 0) {
        bf_heat_info o;
        o.max_power = mx;
        o.min_power = mn;
        o.log_span = span;
        o.smooth_max = smax;
        const float tot32 = (float)tot;
        if (cnt > 0 && tot32 > 0.0f) {
            o.center_col = sc / (double)tot32;
            o.center_row = sr / (double)tot32;
            o.fallback = 0;
        } else {
            const int i = (int)(0xffffffffu - (unsigned)(best & 0xffffffffu));
            o.center_col = (double)(i % Y);
            o.center_row = (double)(i / Y);
            o.fallback = 1;
        }
        o.overlay = p.log_scale ? (gate ? 1 : 0) : ((gate && n_painted > 0) ? 1 : 0);
        o.painted = n_painted;
        o.reserved = 0;
        p.info[f] = o;
    }
}

// ---------------------------------------------------------------------------------------------
// cv2.resize(INTER_LINEAR) for 8-bit images (OpenCV imgproc/resize.cpp, generic path):
//   fx = (float)((dx + 0.5) * scale - 0.5), sx = floor(fx), fx -= sx, clamped to the image in x
//   (not in y: rows are clamped when fetched), coefficients = round(f * 2048) as int16,
//   horizontal pass in int32, vertical: (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2.
// A CTA row handles one SOURCE row pair (j, j+1) and every output row that interpolates between
// them (yofs == j: dst_h/src_h rows when enlarging): a thread owns 4 consecutive output bytes,
// does the horizontal pass for them once (8 values >> 4 kept in registers) and then only the
// two multiplies of the vertical pass per output row; one coalesced 32-bit store per row.
// ---------------------------------------------------------------------------------------------
template <int CN>
__global__ void __launch_bounds__(256) resize_rows_kernel(const unsigned char *__restrict__ src, int sh, int sw,
                                                          unsigned char *__restrict__ dst, int dh, int dw,
                                                          const int *__restrict__ xofs, const short2 *__restrict__ xco,
                                                          const short2 *__restrict__ yco, const int *__restrict__ ystart,
                                                          int aligned)
{
    const int f = blockIdx.z;
    const int j = (int)blockIdx.y - 1;
    const int ya = ystart[blockIdx.y], yb = ystart[blockIdx.y + 1];
    if (ya >= yb) return;
    const int wq = blockIdx.x * blockDim.x + threadIdx.x;
    const int row_bytes = dw * CN;
    if (wq * 4 >= row_bytes) return;
    const int y0 = j < 0 ? 0 : (j > sh - 1 ? sh - 1 : j);
    const int y1 = j + 1 > sh - 1 ? sh - 1 : j + 1;
    const unsigned char *r0 = src + ((size_t)f * sh + y0) * sw * CN;
    const unsigned char *r1 = src + ((size_t)f * sh + y1) * sw * CN;
    int h0[4], h1[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int bi = wq * 4 + k;
        h0[k] = 0; h1[k] = 0;
        if (bi < row_bytes) {
            const int px = bi / CN, ch = bi - px * CN;
            const int sx = xofs[px];
            const int x1 = sx + 1 > sw - 1 ? sw - 1 : sx + 1;
            const short2 a = xco[px];
            h0[k] = ((int)__ldg(r0 + sx * CN + ch) * a.x + (int)__ldg(r0 + x1 * CN + ch) * a.y) >> 4;
            h1[k] = ((int)__ldg(r1 + sx * CN + ch) * a.x + (int)__ldg(r1 + x1 * CN + ch) * a.y) >> 4;
        }
    }
    const bool full = aligned && wq * 4 + 3 < row_bytes;
    unsigned char *out = dst + ((size_t)f * dh + ya) * row_bytes + (size_t)wq * 4;
    for (int y = ya; y < yb; y++, out += row_bytes) {
        const short2 b = yco[y];
        unsigned int v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int t = ((((int)b.x * h0[k]) >> 16) + (((int)b.y * h1[k]) >> 16) + 2) >> 2;
            v[k] = (unsigned int)(t > 255 ? 255 : t);
        }
        if (full) {
            *(unsigned int *)out = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
        } else {
            for (int k = 0; k < 4 && wq * 4 + k < row_bytes; k++) out[k] = (unsigned char)v[k];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// get_entropy (decider.py:16-24): p = v / sum(v), H = -sum p ln(p + 1e-12), confidence 1/(1+H).
// An 8-bit image has 256 distinct values: histogram (zeros skipped, they contribute 0), then
// the sum over values in float64.
// ---------------------------------------------------------------------------------------------
// A frame is split over `parts` CTAs; each adds its (non-zero-value) counts to the frame's global
// histogram, the last one to finish (ticket counter) evaluates the sum in a fixed order and
// clears the workspace for the next call.  Integer counts: the result does not depend on timing.
__global__ void __launch_bounds__(512) entropy_kernel(const unsigned char *__restrict__ img, long bytes, long chunk,
                                                      unsigned int *__restrict__ ws, double *__restrict__ conf)
{
    __shared__ unsigned int hist[8][256];
    __shared__ double scratch[32];
    __shared__ int is_last;
    const int f = blockIdx.y;
    const long lo = (long)blockIdx.x * chunk;
    const long hi = lo + chunk < bytes ? lo + chunk : bytes;
    const unsigned char *p = img + (size_t)f * bytes;
    unsigned int *w = ws + (size_t)f * 260;
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&hist[0][0])[i] = 0u;
    __syncthreads();
    unsigned int *h = hist[(threadIdx.x >> 5) & 7];
    // 16-byte vector body between the aligned bounds, scalar head / tail
    long va = lo + ((16 - (((uintptr_t)(p + lo)) & 15)) & 15);
    if (va > hi) va = hi;
    const long nvec = (hi - va) / 16;
    for (long i = lo + threadIdx.x; i < va; i += blockDim.x) {
        const unsigned int v = p[i];
        if (v) atomicAdd(&h[v], 1u);
    }
    const uint4 *pv = (const uint4 *)(p + va);
    for (long i = threadIdx.x; i < nvec; i += blockDim.x) {
        const uint4 q = __ldg(pv + i);
        const unsigned int wd[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (wd[k] == 0u) continue;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const unsigned int v = (wd[k] >> (8 * j)) & 255u;
                if (v) atomicAdd(&h[v], 1u);
            }
        }
    }
    for (long i = va + nvec * 16 + threadIdx.x; i < hi; i += blockDim.x) {
        const unsigned int v = p[i];
        if (v) atomicAdd(&h[v], 1u);
    }
    __syncthreads();
    const int v = threadIdx.x;
    if (v > 0 && v < 256) {
        unsigned int c = 0;
        for (int k = 0; k < 8; k++) c += hist[k][v];
        if (c) atomicAdd(&w[v], c);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&w[256], 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    auto dadd_op = [](double a, double b) { return a + b; };
    double cnt = 0.0;
    if (v > 0 && v < 256) {
        cnt = (double)__ldcg(&w[v]);
        w[v] = 0u;
    }
    if (v == 0) w[256] = 0u;
    const double s = block_reduce(cnt * (double)v, dadd_op, 0.0, scratch);
    double term = 0.0;
    if (cnt > 0.0 && s > 0.0) {
        const double pr = (double)v / s;
        term = cnt * (pr * log(pr + 1e-12));
    }
    const double ent = -block_reduce(term, dadd_op, 0.0, scratch);
    if (threadIdx.x == 0) conf[f] = 1.0 / (1.0 + ent);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct HeatState {
    std::mutex mu;
    DevBuf lut; bool lut_set = false; bool lut_is_default = false;
    unsigned char lut_host[768];
    DevBuf xofs, xco, yco, ystart;
    int key[4] = {-1, -1, -1, -1};
    // staging of the host-pointer entry point
    DevBuf maps, small, info, big, conf;
    DevBuf ent_ws; int ent_frames = 0;      // [frames][260] u32, zero between calls
};
static HeatState &hs() { static HeatState s; return s; }

// Matplotlib's "jet" (LinearSegmentedColormap, N = 256, gamma 1), reversed and truncated to
// uint8 exactly as generate_color_map does (visual.py:27-48).
static void jet_lut(unsigned char *out)
{
    static const double red[][3] = {{0.00, 0, 0}, {0.35, 0, 0}, {0.66, 1, 1}, {0.89, 1, 1}, {1.00, 0.5, 0.5}};
    static const double green[][3] = {{0.000, 0, 0}, {0.125, 0, 0}, {0.375, 1, 1}, {0.640, 1, 1}, {0.910, 0, 0}, {1.000, 0, 0}};
    static const double blue[][3] = {{0.00, 0.5, 0.5}, {0.11, 1, 1}, {0.34, 1, 1}, {0.65, 0, 0}, {1.00, 0, 0}};
    const double (*seg[3])[3] = {red, green, blue};
    const int nseg[3] = {5, 6, 5};
    const int n = 256;
    const double step = 1.0 / (double)(n - 1);
    for (int c = 0; c < 3; c++) {
        double lut[256];
        const double (*d)[3] = seg[c];
        const int k = nseg[c];
        lut[0] = d[0][2];
        lut[n - 1] = d[k - 1][1];
        for (int i = 1; i < n - 1; i++) {
            const double xi = (double)(n - 1) * ((double)i * step);
            int ind = 0;
            while (ind < k && d[ind][0] * (double)(n - 1) < xi) ind++;       // searchsorted, side = left
            const double x0 = d[ind - 1][0] * (double)(n - 1), x1 = d[ind][0] * (double)(n - 1);
            const double dist = (xi - x0) / (x1 - x0);
            double v = dist * (d[ind][1] - d[ind - 1][2]) + d[ind - 1][2];
            lut[i] = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
        }
        for (int i = 0; i < n; i++) out[i * 3 + c] = (unsigned char)(lut[n - 1 - i] * 255.0);
    }
}

static int ensure_lut(const unsigned char *lut, cudaStream_t st)
{
    HeatState &H = hs();
    unsigned char tmp[768];
    if (!lut) {                                  // NULL = the default jet map, whatever was installed before
        if (H.lut_set && H.lut_is_default) return BF_OK;
        jet_lut(tmp);
        lut = tmp;
    }
    const bool is_default = (lut == tmp);
    if (H.lut_set && memcmp(lut, H.lut_host, 768) == 0) { H.lut_is_default = is_default; return BF_OK; }
    int rc = H.lut.ensure(768);
    if (rc) return rc;
    memcpy(H.lut_host, lut, 768);
    BF_CUDA(cudaMemcpyAsync(H.lut.p, H.lut_host, 768, cudaMemcpyHostToDevice, st));
    BF_CUDA(cudaStreamSynchronize(st));
    H.lut_set = true;
    H.lut_is_default = is_default;
    return BF_OK;
}

static void resize_coeffs(int src, int dst, bool clamp, std::vector<int> &ofs, std::vector<short> &co)
{
    ofs.resize(dst);
    co.resize((size_t)dst * 2);
    const double scale = 1.0 / ((double)dst / (double)src);
    for (int d = 0; d < dst; d++) {
        float f = (float)(((double)d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (clamp) {
            if (s < 0) { f = 0.0f; s = 0; }
            if (s >= src - 1) { f = 0.0f; s = src - 1; }
        }
        co[(size_t)d * 2] = (short)lrintf((1.0f - f) * 2048.0f);
        co[(size_t)d * 2 + 1] = (short)lrintf(f * 2048.0f);
        ofs[d] = s;
    }
}

static int ensure_resize_tables(int sh, int sw, int dh, int dw, cudaStream_t st)
{
    HeatState &H = hs();
    if (H.key[0] == sh && H.key[1] == sw && H.key[2] == dh && H.key[3] == dw) return BF_OK;
    std::vector<int> xo, yo;
    std::vector<short> xc, yc;
    resize_coeffs(sw, dw, true, xo, xc);
    resize_coeffs(sh, dh, false, yo, yc);
    int rc;
    // ystart[j + 1] = first output row whose source row pair is (j, j+1), j = -1 .. sh-1 (yofs is
    // non-decreasing and lies in [-1, sh-1]); ystart[sh + 1] = dh
    std::vector<int> ys((size_t)sh + 2, dh);
    for (int y = dh - 1; y >= 0; y--) {
        int j = yo[y] < -1 ? -1 : (yo[y] > sh - 1 ? sh - 1 : yo[y]);
        ys[(size_t)j + 1] = y;
    }
    for (int j = sh; j >= 0; j--)
        if (ys[(size_t)j] > ys[(size_t)j + 1]) ys[(size_t)j] = ys[(size_t)j + 1];
    ys[0] = 0;
    if ((rc = H.xofs.ensure(xo.size() * 4)) || (rc = H.xco.ensure(xc.size() * 2)) ||
        (rc = H.ystart.ensure(ys.size() * 4)) || (rc = H.yco.ensure(yc.size() * 2)))
        return rc;
    BF_CUDA(cudaMemcpyAsync(H.xofs.p, xo.data(), xo.size() * 4, cudaMemcpyHostToDevice, st));
    BF_CUDA(cudaMemcpyAsync(H.xco.p, xc.data(), xc.size() * 2, cudaMemcpyHostToDevice, st));
    BF_CUDA(cudaMemcpyAsync(H.ystart.p, ys.data(), ys.size() * 4, cudaMemcpyHostToDevice, st));
    BF_CUDA(cudaMemcpyAsync(H.yco.p, yc.data(), yc.size() * 2, cudaMemcpyHostToDevice, st));
    BF_CUDA(cudaStreamSynchronize(st));          // the vectors die here
    H.key[0] = sh; H.key[1] = sw; H.key[2] = dh; H.key[3] = dw;
    return BF_OK;
}

static int heat_dev(const float *d_maps, int frames, long frame_stride, int X, int Y, float threshold,
                    float amount, int exponent, int log_scale, const unsigned char *lut,
                    unsigned char *d_small, short *d_index, bf_heat_info *d_info, cudaStream_t st)
{
    if (!d_maps || !d_small || !d_info || frames < 1 || X < 3 || Y < 3 || exponent < 0 || exponent > 64 ||
        !(amount > 0.0f)) {
        set_error(BF_ERR_ARG, "bf_heatmap_dev: bad arguments (grid %dx%d, frames %d, exponent %d)", X, Y, frames, exponent);
        return BF_ERR_ARG;
    }
    const size_t smem = (size_t)X * Y * sizeof(float);
    if (smem > 200 * 1024) {
        set_error(BF_ERR_CONFIG, "bf_heatmap_dev: a %dx%d map does not fit the %d KB shared-memory tile", X, Y, 200);
        return BF_ERR_CONFIG;
    }
    int rc = ensure_lut(lut, st);
    if (rc) return rc;
    HeatParams hp{d_maps, frame_stride, X, Y, threshold, amount, exponent, log_scale,
                  hs().lut.as<unsigned char>(), d_small, d_index, d_info};
    BF_CUDA(cudaFuncSetAttribute(heat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    heat_kernel<<<frames, kHeatThreads, smem, st>>>(hp);
    BF_CHECK_LAUNCH();
    count_launch();
    return BF_OK;
}

static int resize_dev(const unsigned char *d_src, int frames, int sh, int sw, int cn, unsigned char *d_dst, int dh,
                      int dw, cudaStream_t st)
{
    if (!d_src || !d_dst || frames < 1 || sh < 1 || sw < 1 || dh < 1 || dw < 1 || cn < 1 || cn > 4 || sh > 65000) {
        set_error(BF_ERR_ARG, "bf_resize_linear_u8_dev: bad arguments");
        return BF_ERR_ARG;
    }
    if (sh == dh && sw == dw) {
        BF_CUDA(cudaMemcpyAsync(d_dst, d_src, (size_t)frames * sh * sw * cn, cudaMemcpyDeviceToDevice, st));
        return BF_OK;
    }
    int rc = ensure_resize_tables(sh, sw, dh, dw, st);
    if (rc) return rc;
    HeatState &H = hs();
    const int row_bytes = dw * cn;
    const int words = (row_bytes + 3) / 4;
    const int aligned = (row_bytes % 4 == 0) && (((uintptr_t)d_dst & 3) == 0);
    for (int f0 = 0; f0 < frames; f0 += 65535) {
        const int nf = frames - f0 < 65535 ? frames - f0 : 65535;
        dim3 grid((words + 255) / 256, sh + 1, nf);
        const unsigned char *s0 = d_src + (size_t)f0 * sh * sw * cn;
        unsigned char *o0 = d_dst + (size_t)f0 * dh * row_bytes;
        const int *xo = H.xofs.as<int>(), *ys = H.ystart.as<int>();
        const short2 *xc = H.xco.as<short2>(), *yc = H.yco.as<short2>();
        switch (cn) {
            case 1: resize_rows_kernel<1><<<grid, 256, 0, st>>>(s0, sh, sw, o0, dh, dw, xo, xc, yc, ys, aligned); break;
            case 2: resize_rows_kernel<2><<<grid, 256, 0, st>>>(s0, sh, sw, o0, dh, dw, xo, xc, yc, ys, aligned); break;
            case 3: resize_rows_kernel<3><<<grid, 256, 0, st>>>(s0, sh, sw, o0, dh, dw, xo, xc, yc, ys, aligned); break;
            default: resize_rows_kernel<4><<<grid, 256, 0, st>>>(s0, sh, sw, o0, dh, dw, xo, xc, yc, ys, aligned); break;
        }
        BF_CHECK_LAUNCH();
        count_launch();
    }
    return BF_OK;
}

static int entropy_dev(const unsigned char *d_img, int frames, long bytes, double *d_conf, cudaStream_t st)
{
    if (!d_img || !d_conf || frames < 1 || bytes < 1) {
        set_error(BF_ERR_ARG, "bf_entropy_dev: bad arguments");
        return BF_ERR_ARG;
    }
    HeatState &H = hs();
    if (frames > H.ent_frames) {
        int rc = H.ent_ws.ensure((size_t)frames * 260 * 4);
        if (rc) return rc;
        BF_CUDA(cudaMemsetAsync(H.ent_ws.p, 0, (size_t)frames * 260 * 4, st));
        H.ent_frames = frames;
    }
    long parts = (bytes + 65535) / 65536;
    parts = parts < 1 ? 1 : (parts > 64 ? 64 : parts);
    long chunk = ((bytes + parts - 1) / parts + 15) / 16 * 16;
    parts = (bytes + chunk - 1) / chunk;
    for (int f0 = 0; f0 < frames; f0 += 65535) {
        const int nf = frames - f0 < 65535 ? frames - f0 : 65535;
        entropy_kernel<<<dim3((unsigned)parts, nf), 512, 0, st>>>(d_img + (size_t)f0 * bytes, bytes, chunk,
                                                                  H.ent_ws.as<unsigned int>() + (size_t)f0 * 260,
                                                                  d_conf + f0);
    }
    BF_CHECK_LAUNCH();
    count_launch();
    return BF_OK;
}

}  // namespace bf

using namespace bf;

extern "C" int bf_jet_lut(unsigned char *lut768)
{
    clear_error();
    if (!lut768) { set_error(BF_ERR_ARG, "bf_jet_lut: null output"); return BF_ERR_ARG; }
    jet_lut(lut768);
    return BF_OK;
}

extern "C" int bf_heatmap_dev(const float *d_maps, int frames, long frame_stride, int res_x, int res_y,
                              float threshold, float amount, int exponent, int log_scale,
                              const unsigned char *lut, unsigned char *d_small, short *d_index,
                              bf_heat_info *d_info, void *stream)
{
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(hs().mu);
    return heat_dev(d_maps, frames, frame_stride, res_x, res_y, threshold, amount, exponent, log_scale, lut, d_small,
                    d_index, d_info, (cudaStream_t)stream);
}

extern "C" int bf_resize_linear_u8_dev(const unsigned char *d_src, int frames, int src_h, int src_w, int channels,
                                       unsigned char *d_dst, int dst_h, int dst_w, void *stream)
{
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(hs().mu);
    return resize_dev(d_src, frames, src_h, src_w, channels, d_dst, dst_h, dst_w, (cudaStream_t)stream);
}

extern "C" int bf_entropy_dev(const unsigned char *d_img, int frames, long bytes_per_frame, double *d_confidence,
                              void *stream)
{
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(hs().mu);
    return entropy_dev(d_img, frames, bytes_per_frame, d_confidence, (cudaStream_t)stream);
}

// Host-pointer form: maps in, window-sized heat maps + per-frame info (+ confidence) out.
extern "C" int bf_heatmap(const float *maps, int frames, int res_x, int res_y, float threshold, float amount,
                          int exponent, int log_scale, const unsigned char *lut, int out_w, int out_h,
                          unsigned char *heat_out, bf_heat_info *info_out, double *confidence_out)
{
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    if (!maps || !heat_out || !info_out || frames < 1 || out_w < 1 || out_h < 1) {
        set_error(BF_ERR_ARG, "bf_heatmap: bad arguments");
        return BF_ERR_ARG;
    }
    HeatState &H = hs();
    std::lock_guard<std::mutex> lk(H.mu);
    const size_t P = (size_t)res_x * res_y;
    const size_t big_bytes = (size_t)frames * out_w * out_h * 3;
    if ((rc = H.maps.ensure(frames * P * 4)) || (rc = H.small.ensure(frames * P * 3)) ||
        (rc = H.info.ensure(frames * sizeof(bf_heat_info))) || (rc = H.big.ensure(big_bytes)) ||
        (rc = H.conf.ensure(frames * sizeof(double))))
        return rc;
    cudaStream_t st = 0;
    BF_CUDA(cudaMemcpyAsync(H.maps.p, maps, frames * P * 4, cudaMemcpyHostToDevice, st));
    rc = heat_dev(H.maps.as<float>(), frames, (long)P, res_x, res_y, threshold, amount, exponent, log_scale, lut,
                  H.small.as<unsigned char>(), nullptr, H.info.as<bf_heat_info>(), st);
    if (rc) return rc;
    rc = resize_dev(H.small.as<unsigned char>(), frames, res_y, res_x, 3, H.big.as<unsigned char>(), out_h, out_w, st);
    if (rc) return rc;
    if (confidence_out) {
        rc = entropy_dev(H.big.as<unsigned char>(), frames, (long)out_w * out_h * 3, H.conf.as<double>(), st);
        if (rc) return rc;
        BF_CUDA(cudaMemcpyAsync(confidence_out, H.conf.p, frames * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    BF_CUDA(cudaMemcpyAsync(heat_out, H.big.p, big_bytes, cudaMemcpyDeviceToHost, st));
    BF_CUDA(cudaMemcpyAsync(info_out, H.info.p, frames * sizeof(bf_heat_info), cudaMemcpyDeviceToHost, st));
    BF_CUDA(cudaStreamSynchronize(st));
    return BF_OK;
}
