// das_simple.cu -- one-thread-per-sample delay-and-sum kernels.
//
// These are the general kernels: any N_SAMPLES <= 1024, every delay algorithm of
// the reference (pad, lerp, FIR in both accumulation orders, hybrid).  They are
// the product path for the FIR/hybrid power maps and for the single-call MISO and
// *_delay entry points, and the cross-check for the tiled TMA kernel.  One CTA
// steers one direction of one frame: thread t owns output sample t and walks the
// microphones in table order, so the arithmetic order per sample is exactly the
// reference's (algorithms/*.c, see per-function citations).
#include "bf_common.cuh"

namespace bf {

struct SimpleTab {
    const int *whole;       // pad / lerp / hybrid integer delays
    const float *weight;    // lerp weights (1 - frac)
    const float *taps;      // FIR taps [..][T]
    int T;                  // taps
    int fused;              // sequential FIR: fma (1) or mul+add (0)
};

// contribution of microphone row `row` to output sample t, accumulated into acc
// in the reference's rounding order.  e = flat table index (offset + m).
template <int ALGO>
__device__ __forceinline__ float accumulate(float acc, const float *__restrict__ row, int t, int N,
                                            const SimpleTab &tb, size_t e)
{
    if (ALGO == BF_ALGO_PAD) {
        // pad_and_sum.c:41-47: out[w+i] += s[i]
        const int i = t - tb.whole[e];
        if (i >= 0 && i < N) acc = __fadd_rn(acc, row[i]);
    } else if (ALGO == BF_ALGO_LERP) {
        // lerp_and_sum.c:50-56: out[w+i+1] += s[i] + h*(s[i+1]-s[i]), i < N-w-1
        const int i = t - tb.whole[e] - 1;
        if (i >= 0 && i + 1 < N) {
            const float a = row[i], b = row[i + 1];
            acc = __fadd_rn(acc, __fmaf_rn(tb.weight[e], __fsub_rn(b, a), a));
        }
    } else if (ALGO == BF_ALGO_FIR_SEQ) {
        // convolve_and_sum.c:197-211: out[i] += h[k]*padded[i+k], padded = zeros + s at T/2
        const float *h = tb.taps + e * tb.T;
        const int base = t - tb.T / 2;
        for (int k = 0; k < tb.T; k++) {
            const int i = base + k;
            const float p = (i >= 0 && i < N) ? row[i] : 0.0f;
            acc = tb.fused ? __fmaf_rn(h[k], p, acc) : __fadd_rn(acc, __fmul_rn(h[k], p));
        }
    } else if (ALGO == BF_ALGO_FIR_LANES) {
        // convolve_and_sum.c:158-192 + sum8 131-153 (AVX2 lane order)
        const float *h = tb.taps + e * tb.T;
        const int base = t - tb.T / 2;
        float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < tb.T; k += 8) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int i = base + k + j;
                const float p = (i >= 0 && i < N) ? row[i] : 0.0f;
                x[j] = __fmaf_rn(p, h[k + j], x[j]);
            }
        }
        const float q0 = __fadd_rn(x[0], x[4]), q1 = __fadd_rn(x[1], x[5]);
        const float q2 = __fadd_rn(x[2], x[6]), q3 = __fadd_rn(x[3], x[7]);
        acc = __fadd_rn(acc, __fadd_rn(__fadd_rn(q0, q2), __fadd_rn(q1, q3)));
    } else if (ALGO == BF_ALGO_HYBRID) {
        // hybrid_convolve_and_sum.c:51-64: out[w+i+1] += sum_k h[k]*padded[i+k], i < N-w-1
        const int i0 = t - tb.whole[e] - 1;
        if (i0 >= 0) {
            const float *h = tb.taps + e * tb.T;
            const int base = i0 - tb.T / 2;
            for (int k = 0; k < tb.T; k++) {
                const int i = base + k;
                const float p = (i >= 0 && i < N) ? row[i] : 0.0f;
                acc = tb.fused ? __fmaf_rn(h[k], p, acc) : __fadd_rn(acc, __fmul_rn(h[k], p));
            }
        }
    }
    return acc;
}

// ---- power map: grid (d_count, frames), block = N threads ---------------------
template <int ALGO>
__global__ void mimo_simple_kernel(const float *__restrict__ sig, float *__restrict__ img,
                                   const int *__restrict__ mic_ids, int n, int n_mics_total, int N,
                                   ImgLayout lay, int d_begin, SimpleTab tb, float fn)
{
    extern __shared__ float sq[];
    const int d = d_begin + blockIdx.x, f = blockIdx.y, t = threadIdx.x;
    const float *fs = sig + (size_t)f * n_mics_total * N;
    float acc = 0.0f;
    for (int m = 0; m < n; m++)
        acc = accumulate<ALGO>(acc, fs + (size_t)mic_ids[m] * N, t, N, tb, (size_t)d * n + m);
    // *_and_sum.c power epilogue: out/n, square, in-order sum, /N
    const float x = __fdiv_rn(acc, fn);
    sq[t] = __fmul_rn(x, x);
    __syncthreads();
    if (t == 0) {
        float s = 0.0f;
        for (int k = 0; k < N; k++) s = __fadd_rn(s, sq[k]);
        img[(long)f * lay.frame_stride + (long)(d - lay.d_origin) * lay.dir_stride] = __fdiv_rn(s, (float)N);
    }
}

// ---- MISO: grid (blocks), block = N threads -----------------------------------
// by_mic: table indexed by mic id (miso_pad2, pad_and_sum.c:77-92) instead of column.
template <int ALGO>
__global__ void miso_simple_kernel(const float *__restrict__ sig, float *__restrict__ out,
                                   const int *__restrict__ mic_ids, int n, int n_mics_total, int N,
                                   int offset, int by_mic, SimpleTab tb, int scale, float fn,
                                   float gain)
{
    const int b = blockIdx.x, t = threadIdx.x;
    const float *fs = sig + (size_t)b * n_mics_total * N;
    float acc = 0.0f;
    for (int m = 0; m < n; m++) {
        const int mic = mic_ids[m];
        const size_t e = by_mic ? (size_t)mic : (size_t)offset + m;
        acc = accumulate<ALGO>(acc, fs + (size_t)mic * N, t, N, tb, e);
    }
    if (scale) acc = __fmul_rn(__fdiv_rn(acc, fn), gain);      // api.c:519-523
    out[(size_t)b * N + t] = acc;
}

static int make_tab(int algo, SimpleTab &tb, size_t need, const char *who)
{
    State &S = state();
    Tables &T = S.tab;
    tb = SimpleTab{};
    tb.T = S.cfg.n_taps;
    tb.fused = S.cfg.fir_fused < 0 ? (S.cfg.n_taps <= 16) : (S.cfg.fir_fused != 0);
    size_t have = 0;
    switch (algo) {
        case BF_ALGO_PAD:   tb.whole = T.pad_whole.as<int>(); have = T.pad_count; break;
        case -1:            tb.whole = T.trunc_whole.as<int>(); have = T.trunc_count; break;
        case -2:            tb.whole = T.pad2_whole.as<int>(); have = T.pad2_count; break;
        case BF_ALGO_LERP:  tb.whole = T.lerp_whole.as<int>(); tb.weight = T.lerp_weight.as<float>();
                            have = T.lerp_count; break;
        case BF_ALGO_FIR_SEQ:
        case BF_ALGO_FIR_LANES:
            tb.taps = T.fir_taps.as<float>(); have = T.fir_count / (size_t)tb.T; break;
        case BF_ALGO_HYBRID: tb.whole = T.hyb_whole.as<int>(); tb.taps = T.hyb_taps.as<float>();
                            have = T.hyb_count; break;
        default: set_error(BF_ERR_ARG, "%s: bad algo %d", who, algo); return BF_ERR_ARG;
    }
    if (have < need || have == 0) {
        set_error(BF_ERR_NOT_LOADED, "%s: table for algo %d holds %zu entries, need %zu", who, algo,
                  have, need);
        return BF_ERR_NOT_LOADED;
    }
    if (algo == BF_ALGO_FIR_LANES && tb.T % 8 != 0) {
        set_error(BF_ERR_CONFIG, "vectorized FIR order needs N_TAPS %% 8 == 0 (got %d)", tb.T);
        return BF_ERR_CONFIG;
    }
    return BF_OK;
}

int mimo_simple(int algo, const float *d_sig, float *d_img, int frames, const int *d_mics, int n,
                int d_begin, int d_count, ImgLayout lay, cudaStream_t st)
{
    State &S = state();
    const int N = S.cfg.n_samples, D = S.cfg.max_res_x * S.cfg.max_res_y;
    if (N < 1 || N > 1024) { set_error(BF_ERR_CONFIG, "N_SAMPLES %d not in [1,1024]", N); return BF_ERR_CONFIG; }
    SimpleTab tb;
    int rc = make_tab(algo, tb, (size_t)D * n, "mimo");
    if (rc) return rc;
    dim3 grid(d_count, frames);
    const size_t sm = N * sizeof(float);
    const float fn = (float)n;
    const int M = S.cfg.n_microphones;
#define BF_GO(A) mimo_simple_kernel<A><<<grid, N, sm, st>>>(d_sig, d_img, d_mics, n, M, N, lay, d_begin, tb, fn)
    switch (algo) {
        case BF_ALGO_PAD: case -1: BF_GO(BF_ALGO_PAD); break;
        case BF_ALGO_LERP:      BF_GO(BF_ALGO_LERP); break;
        case BF_ALGO_FIR_SEQ:   BF_GO(BF_ALGO_FIR_SEQ); break;
        case BF_ALGO_FIR_LANES: BF_GO(BF_ALGO_FIR_LANES); break;
        case BF_ALGO_HYBRID:    BF_GO(BF_ALGO_HYBRID); break;
    }
#undef BF_GO
    BF_CHECK_LAUNCH();
    count_launch();
    return BF_OK;
}

int miso_simple(int algo, const float *d_sig, float *d_out, int blocks, const int *d_mics, int n,
                int offset, int by_mic_id, int scale, cudaStream_t st)
{
    State &S = state();
    const int N = S.cfg.n_samples;
    if (N < 1 || N > 1024) { set_error(BF_ERR_CONFIG, "N_SAMPLES %d not in [1,1024]", N); return BF_ERR_CONFIG; }
    SimpleTab tb;
    const int talgo = by_mic_id ? -2 : algo;
    // miso_pad2 indexes its table by microphone id (pad_and_sum.c:77-92): every id < N_MICROPHONES must be covered
    size_t need = by_mic_id ? (size_t)S.cfg.n_microphones : (size_t)offset + n;
    if (algo == BF_ALGO_FIR_SEQ || algo == BF_ALGO_FIR_LANES)
        need = (size_t)offset / S.cfg.n_taps + n;        // FIR offsets are in floats (d*n*T)
    int rc = make_tab(talgo, tb, need, "miso");
    if (rc) return rc;
    const float fn = (float)n, gain = S.cfg.mic_gain;
    const int M = S.cfg.n_microphones;
    // FIR tables are addressed per entry of T floats: offset (floats) -> entries
    int off_e = offset;
    if (algo == BF_ALGO_FIR_SEQ || algo == BF_ALGO_FIR_LANES) {
        if (offset % tb.T != 0) { set_error(BF_ERR_ARG, "FIR offset %d not a multiple of N_TAPS", offset); return BF_ERR_ARG; }
        off_e = offset / tb.T;
    }
#define BF_GO(A) miso_simple_kernel<A><<<blocks, N, 0, st>>>(d_sig, d_out, d_mics, n, M, N, off_e, by_mic_id, tb, scale, fn, gain)
    switch (algo) {
        case BF_ALGO_PAD:       BF_GO(BF_ALGO_PAD); break;
        case BF_ALGO_LERP:      BF_GO(BF_ALGO_LERP); break;
        case BF_ALGO_FIR_SEQ:   BF_GO(BF_ALGO_FIR_SEQ); break;
        case BF_ALGO_FIR_LANES: BF_GO(BF_ALGO_FIR_LANES); break;
        case BF_ALGO_HYBRID:    BF_GO(BF_ALGO_HYBRID); break;
        default: set_error(BF_ERR_ARG, "miso: bad algo %d", algo); return BF_ERR_ARG;
    }
#undef BF_GO
    BF_CHECK_LAUNCH();
    count_launch();
    return BF_OK;
}

// ---- single-row *_delay entry points (pad_delay, lerp_delay, convolve_delay_*) ----
// kind: 0 pad (coef unused, pad), 1 lerp (h, pad), 2 FIR seq add, 3 FIR lanes add,
//       4 FIR lanes overwrite (convolve_delay_vectorized), 5 hybrid add (coef, pad)
__global__ void single_delay_kernel(int kind, const float *__restrict__ row,
                                    const float *__restrict__ coef, float h, int pad,
                                    float *__restrict__ out, int N, int T, int fused)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    float acc = out[t];
    SimpleTab tb{};
    tb.T = T; tb.fused = fused; tb.taps = coef; tb.whole = &pad; tb.weight = &h;
    // the table pointers above are kernel-parameter addresses: usable because the
    // accumulate() helpers only read element 0 of them (e = 0).
    switch (kind) {
        case 0: acc = accumulate<BF_ALGO_PAD>(acc, row, t, N, tb, 0); break;
        case 1: acc = accumulate<BF_ALGO_LERP>(acc, row, t, N, tb, 0); break;
        case 2: acc = accumulate<BF_ALGO_FIR_SEQ>(acc, row, t, N, tb, 0); break;
        case 3: acc = accumulate<BF_ALGO_FIR_LANES>(acc, row, t, N, tb, 0); break;
        case 4: acc = accumulate<BF_ALGO_FIR_LANES>(0.0f, row, t, N, tb, 0); break;
        case 5: acc = accumulate<BF_ALGO_HYBRID>(acc, row, t, N, tb, 0); break;
    }
    out[t] = acc;
}

int single_delay(int kind, const float *h_signal, const float *h_coef, float h, int pad,
                 float *h_out)
{
    State &S = state();
    int rc = ensure_device();
    if (rc) return rc;
    const int N = S.cfg.n_samples, T = S.cfg.n_taps;
    const int fused = S.cfg.fir_fused < 0 ? (T <= 16) : (S.cfg.fir_fused != 0);
    if ((kind == 3 || kind == 4) && T % 8 != 0) {
        set_error(BF_ERR_CONFIG, "vectorized FIR order needs N_TAPS %% 8 == 0 (got %d)", T);
        return BF_ERR_CONFIG;
    }
    rc = S.d_scratch.ensure((size_t)(2 * N + T + 8) * sizeof(float));
    if (rc) return rc;
    float *d_row = S.d_scratch.as<float>(), *d_out = d_row + N, *d_coef = d_out + N;
    BF_CUDA(cudaMemcpy(d_row, h_signal, N * sizeof(float), cudaMemcpyHostToDevice));
    BF_CUDA(cudaMemcpy(d_out, h_out, N * sizeof(float), cudaMemcpyHostToDevice));
    if (h_coef) BF_CUDA(cudaMemcpy(d_coef, h_coef, T * sizeof(float), cudaMemcpyHostToDevice));
    single_delay_kernel<<<(N + 255) / 256, 256>>>(kind, d_row, d_coef, h, pad, d_out, N, T, fused);
    BF_CHECK_LAUNCH();
    count_launch();
    BF_CUDA(cudaMemcpy(h_out, d_out, N * sizeof(float), cudaMemcpyDeviceToHost));
    return BF_OK;
}

}  // namespace bf
