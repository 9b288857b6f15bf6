// das_fir.cu -- FIR (fractional-delay filter) delay-and-sum power maps, tiled TMA kernel.
//
// Replaces mimo_convolve_naive / mimo_convolve_vectorized (algorithms/convolve_and_sum.c:264-324
// over convolve_delay_naive_add 197-211 and convolve_delay_vectorized_add 158-192 + sum8 131-153):
//
//   out[t] (+)= sum_k h[d][m][k] * padded_m[t + k],  padded = T/2 zeros, the N samples, T/2 zeros
//   img[d] = 1/N * sum_t (out[t] / n)^2
//
// Every (direction, microphone) has its own T taps, so unlike the integer-delay kernels nothing
// is shared between directions except the samples.  Design (sm_100a):
//   * persistent grid, one CTA per SM = 1 producer warp + W consumer warps; a consumer warp owns
//     8 consecutive directions x all 256 samples; lane l holds the 8 CONSECUTIVE samples
//     8l..8l+7 of 4 direction PAIRS in 32 float2 accumulators
//   * per microphone and group of 8 taps the lane loads one 16-float window of the zero-padded
//     row (4 LDS.128, 32-byte aligned because T/2 and 8l are multiples of 4/8) and reuses it for
//     all 8 taps x 8 directions: 256 packed FMAs (fma.f32x2: the two halves are two directions,
//     the sample is the broadcast operand) per 4 window loads + 16 broadcast coefficient loads
//   * the producer warp streams, per ring stage, Mt microphone rows (1 KB bulk copies) and for
//     every consumer warp the coefficient block of its direction group for those microphones
//     (one bulk copy each) -- the taps are re-laid out once per table load as
//     [group][mic][tap][8 directions] so that block is contiguous
//   * the FMA chain per output sample runs in the reference's order (m, then k) with the
//     reference's contraction (fused for T <= 16, mul+add above, SURVEY 7.3 / oracle.c), or in
//     the AVX lane order + sum8 tree of the "vectorized" variant (T = 8)
//   * exact epilogue: out/n, square, in-order sum over t through a per-warp shared row
#include "bf_common.cuh"

namespace bf {

static constexpr int kFirStages = 4;
static constexpr int kFirMaxWarps = 15;
static constexpr int kFirScratch = 264;     // floats per direction row of the epilogue scratch

enum { kFirSeqFused = 0, kFirSeqUnfused = 1, kFirLanes = 2 };

struct FirParams {
    const float *sig;          // [frames][n_mics_total][256]
    float *img;
    const int *mic_ids;
    const float *coef;         // [groups][n][T][8]
    int n, n_mics_total, d_begin, d_count, frames;
    long img_fs, img_ds;
    int d_origin;
    int groups, tiles_per_frame, total_tiles;
    int W, Mt, T;
    int n_pow2;
    float fn, inv_n;
    unsigned zero;             // always 0; opaque to the compiler (see fir_step)
};

// taps [D][n][T] (reference layout, flat index (d*n + m)*T + k) -> [group][m][k][r]
__global__ void fir_relayout_kernel(const float *__restrict__ taps, float *__restrict__ out, int n, int T,
                                    int d_begin, int d_count, long total)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int r = (int)(idx & 7);
    long q = idx >> 3;
    const int k = (int)(q % T); q /= T;
    const int m = (int)(q % n);
    const long g = q / n;
    long dl = g * 8 + r;
    if (dl >= d_count) dl = d_count - 1;
    out[idx] = taps[((size_t)(d_begin + dl) * n + m) * T + k];
}

// Unfused multiply-add (the reference's mul + add for T >= 32).  The compiler contracts a packed
// multiply followed by a packed add into one FFMA2 even under --fmad=false, whether written with
// the __fmul2_rn/__fadd2_rn intrinsics, as inline mul.rn.f32x2 / add.rn.f32x2, or as
// fma(h, a, -0) + add (all three seen in SASS).  XOR-ing the product with a zero that is only
// known at run time (a kernel parameter) keeps the two roundings apart; the LOP3s run on the
// integer pipe, next to the FP32 pipe.
template <int MODE>
__device__ __forceinline__ float2 fir_step(float2 acc, float2 h, float2 a, unsigned zero)
{
    if (MODE == kFirSeqUnfused) {
        float2 pr = __fmul2_rn(h, a);
        pr.x = __uint_as_float(__float_as_uint(pr.x) ^ zero);
        pr.y = __uint_as_float(__float_as_uint(pr.y) ^ zero);
        return __fadd2_rn(acc, pr);
    }
    return __ffma2_rn(h, a, acc);
}

template <int MODE, bool EXACT>
__global__ void __launch_bounds__((kFirMaxWarps + 1) * 32, 1) das_fir_kernel(const FirParams p)
{
    constexpr int N = 256;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = p.W, T = p.T, TG = T >> 3, Mt = p.Mt;
    const int RS = N + T;                                   // row stride (floats): T/2 zeros each side
    const size_t rows_bytes = (size_t)Mt * RS * 4;
    const size_t coef_warp_floats = (size_t)Mt * T * 8;     // one warp's block in a stage
    const size_t stage_bytes = rows_bytes + (size_t)W * coef_warp_floats * 4;

    uint64_t *full = (uint64_t *)smem;
    uint64_t *empty = full + kFirStages;
    unsigned char *stages = smem + 128;
    float *scratch_all = (float *)(stages + kFirStages * stage_bytes);

    // zero the pad columns of every row once; the bulk copies only ever write the N samples
    {
        const int rows_total = kFirStages * Mt;
        for (int i = threadIdx.x; i < rows_total * T; i += blockDim.x) {
            const int row = i / T, c = i - row * T;
            const int s = row / Mt, r = row - s * Mt;
            float *rp = (float *)(stages + (size_t)s * stage_bytes) + (size_t)r * RS;
            rp[c < T / 2 ? c : N + c] = 0.0f;
        }
        if (threadIdx.x == 0) {
            for (int s = 0; s < kFirStages; s++) {
                bfptx::mbar_init(&full[s], 1);
                bfptx::mbar_init(&empty[s], W);
            }
            bfptx::fence_mbar_init();
        }
    }
    __syncthreads();

    const int nchunks = (p.n + Mt - 1) / Mt;

    if (warp == W) {
        // =================== producer warp ==================================
        int s = 0;
        uint32_t ph = 1;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int frame = tile / p.tiles_per_frame;
            const int g0 = (tile - frame * p.tiles_per_frame) * W;
            const float *fsig = p.sig + (size_t)frame * p.n_mics_total * N;
            for (int c = 0; c < nchunks; c++) {
                const int m0 = c * Mt;
                const int cnt = min(Mt, p.n - m0);
                bfptx::mbar_wait(&empty[s], ph);
                const uint32_t cbytes = (uint32_t)(cnt * T * 8 * 4);
                if (lane == 0)
                    bfptx::mbar_arrive_expect_tx(&full[s], (uint32_t)(cnt * N * 4) + (uint32_t)W * cbytes);
                __syncwarp();
                unsigned char *sb = stages + (size_t)s * stage_bytes;
                if (lane < cnt) {
                    const int mic = p.mic_ids[m0 + lane];
                    bfptx::bulk_g2s((float *)sb + (size_t)lane * RS + T / 2, fsig + (size_t)mic * N, N * 4, &full[s]);
                }
                if (lane < W) {
                    const int g = min(g0 + lane, p.groups - 1);
                    bfptx::bulk_g2s((float *)(sb + rows_bytes) + (size_t)lane * coef_warp_floats,
                                    p.coef + ((size_t)g * p.n + m0) * T * 8, cbytes, &full[s]);
                }
                if (++s == kFirStages) { s = 0; ph ^= 1; }
            }
        }
        return;
    }

    // ======================= consumer warps =================================
    float *scratch = scratch_all + (size_t)warp * 2 * kFirScratch;
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int frame = tile / p.tiles_per_frame;
        const int g = (tile - frame * p.tiles_per_frame) * W + warp;
        const bool active = g < p.groups;

        float2 acc[4][8];                                   // [direction pair][sample 8l + t]
#pragma unroll
        for (int rp = 0; rp < 4; rp++)
#pragma unroll
            for (int t = 0; t < 8; t++) acc[rp][t] = make_float2(0.f, 0.f);

        for (int c = 0; c < nchunks; c++) {
            const int cnt = min(Mt, p.n - c * Mt);
            bfptx::mbar_wait(&full[s], ph);
            if (active) {
                const unsigned char *sb = stages + (size_t)s * stage_bytes;
                const float *rowp = (const float *)sb + 8 * lane;
                const float4 *cb = (const float4 *)((const float *)(sb + rows_bytes) + (size_t)warp * coef_warp_floats);
                for (int mm = 0; mm < cnt; mm++, rowp += RS) {
                    for (int tg = 0; tg < TG; tg++, cb += 16) {
                        // window: padded[8l + 8tg .. + 15] (15 used)
                        const float4 *wp = (const float4 *)(rowp + 8 * tg);
                        const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
                        const float w[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w,
                                             w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
                        if (MODE != kFirLanes) {
                            float2 ws[15];
#pragma unroll
                            for (int i = 0; i < 15; i++) ws[i] = make_float2(w[i], w[i]);
#pragma unroll
                            for (int kk = 0; kk < 8; kk++) {
                                const float4 c0 = cb[kk * 2], c1 = cb[kk * 2 + 1];
                                const float2 h0 = make_float2(c0.x, c0.y), h1 = make_float2(c0.z, c0.w);
                                const float2 h2 = make_float2(c1.x, c1.y), h3 = make_float2(c1.z, c1.w);
#pragma unroll
                                for (int t = 0; t < 8; t++) {
                                    acc[0][t] = fir_step<MODE>(acc[0][t], h0, ws[t + kk], p.zero);
                                    acc[1][t] = fir_step<MODE>(acc[1][t], h1, ws[t + kk], p.zero);
                                    acc[2][t] = fir_step<MODE>(acc[2][t], h2, ws[t + kk], p.zero);
                                    acc[3][t] = fir_step<MODE>(acc[3][t], h3, ws[t + kk], p.zero);
                                }
                            }
                        } else {
                            // AVX order (T == 8): x[j] = fma(p[t+j], h[j], 0); ((x0+x4)+(x2+x6)) + ((x1+x5)+(x3+x7))
                            const float2 *cb2 = (const float2 *)cb;
#pragma unroll
                            for (int rp = 0; rp < 4; rp++) {
                                float2 h[8];
#pragma unroll
                                for (int j = 0; j < 8; j++) h[j] = cb2[j * 4 + rp];
#pragma unroll
                                for (int t = 0; t < 8; t++) {
                                    float2 x[8];
#pragma unroll
                                    for (int j = 0; j < 8; j++)
                                        x[j] = __ffma2_rn(make_float2(w[t + j], w[t + j]), h[j], make_float2(0.f, 0.f));
                                    const float2 q0 = __fadd2_rn(x[0], x[4]), q1 = __fadd2_rn(x[1], x[5]);
                                    const float2 q2 = __fadd2_rn(x[2], x[6]), q3 = __fadd2_rn(x[3], x[7]);
                                    acc[rp][t] = __fadd2_rn(acc[rp][t], __fadd2_rn(__fadd2_rn(q0, q2), __fadd2_rn(q1, q3)));
                                }
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) bfptx::mbar_arrive(&empty[s]);
            if (++s == kFirStages) { s = 0; ph ^= 1; }
        }

        if (!active) continue;

        // ---- epilogue: out/n, square, sum over t in order, /N (convolve_and_sum.c:281-290) ----
        float *img = p.img + (long)frame * p.img_fs + (long)(p.d_begin + g * 8 - p.d_origin) * p.img_ds;
        const int valid = min(8, p.d_count - g * 8);
#pragma unroll
        for (int rp = 0; rp < 4; rp++) {
            float sq0[8], sq1[8];
#pragma unroll
            for (int t = 0; t < 8; t++) {
                float x0 = acc[rp][t].x, x1 = acc[rp][t].y;
                if (p.n_pow2) { x0 = __fmul_rn(x0, p.inv_n); x1 = __fmul_rn(x1, p.inv_n); }
                else          { x0 = __fdiv_rn(x0, p.fn);    x1 = __fdiv_rn(x1, p.fn); }
                sq0[t] = __fmul_rn(x0, x0);
                sq1[t] = __fmul_rn(x1, x1);
            }
            if (EXACT) {
                float4 *s0 = (float4 *)(scratch + 8 * lane), *s1 = (float4 *)(scratch + kFirScratch + 8 * lane);
                s0[0] = make_float4(sq0[0], sq0[1], sq0[2], sq0[3]);
                s0[1] = make_float4(sq0[4], sq0[5], sq0[6], sq0[7]);
                s1[0] = make_float4(sq1[0], sq1[1], sq1[2], sq1[3]);
                s1[1] = make_float4(sq1[4], sq1[5], sq1[6], sq1[7]);
                __syncwarp();
                if (lane < 2) {
                    const float4 *sp = (const float4 *)(scratch + lane * kFirScratch);
                    float run = 0.0f;
#pragma unroll 8
                    for (int i = 0; i < N / 4; i++) {
                        const float4 v = sp[i];
                        run = __fadd_rn(run, v.x);
                        run = __fadd_rn(run, v.y);
                        run = __fadd_rn(run, v.z);
                        run = __fadd_rn(run, v.w);
                    }
                    const int r = rp * 2 + lane;
                    if (r < valid) img[(long)r * p.img_ds] = __fmul_rn(run, 1.0f / (float)N);
                }
                __syncwarp();
            } else {
                float t0 = 0.0f, t1 = 0.0f;
#pragma unroll
                for (int t = 0; t < 8; t++) { t0 += sq0[t]; t1 += sq1[t]; }
#pragma unroll
                for (int sh = 16; sh > 0; sh >>= 1) {
                    t0 += __shfl_xor_sync(0xffffffffu, t0, sh);
                    t1 += __shfl_xor_sync(0xffffffffu, t1, sh);
                }
                if (lane == 0) {
                    if (rp * 2 < valid) img[(long)(rp * 2) * p.img_ds] = __fmul_rn(t0, 1.0f / (float)N);
                    if (rp * 2 + 1 < valid) img[(long)(rp * 2 + 1) * p.img_ds] = __fmul_rn(t1, 1.0f / (float)N);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------
struct FirGroupTable {
    DevBuf coef;
    uint64_t version = ~0ull;
    int n = -1, d_begin = -1, d_count = -1, T = -1, groups = 0;
};
static FirGroupTable &fgt() { static FirGroupTable t; return t; }

bool fir_tiled_supported(int algo, int N, int T)
{
    if (N != 256 || T < 8 || (T & 7)) return false;
    if (algo == BF_ALGO_FIR_LANES) return T == 8;
    return algo == BF_ALGO_FIR_SEQ;
}

int fir_tiled(int algo, const float *d_sig, float *d_img, int frames, const int *d_mics, int n, int d_begin,
              int d_count, ImgLayout lay, cudaStream_t st)
{
    State &S = state();
    Tables &Tb = S.tab;
    const int N = S.cfg.n_samples, T = S.cfg.n_taps;
    const int D = S.cfg.max_res_x * S.cfg.max_res_y;
    if (!fir_tiled_supported(algo, N, T)) {
        set_error(BF_ERR_CONFIG, "fir_tiled: unsupported N_SAMPLES %d / N_TAPS %d", N, T);
        return BF_ERR_CONFIG;
    }
    if (Tb.fir_count < (size_t)D * n * T || Tb.fir_taps.p == nullptr) {
        set_error(BF_ERR_NOT_LOADED, "FIR table holds %zu floats, need D*n*T = %d*%d*%d", Tb.fir_count, D, n, T);
        return BF_ERR_NOT_LOADED;
    }
    if (((uintptr_t)d_sig & 15) != 0) {
        set_error(BF_ERR_ARG, "signal buffer must be 16-byte aligned for bulk copies");
        return BF_ERR_ARG;
    }
    FirGroupTable &G = fgt();
    const int groups = (d_count + 7) / 8;
    if (G.version != Tb.version || G.n != n || G.d_begin != d_begin || G.d_count != d_count || G.T != T) {
        const long total = (long)groups * n * T * 8;
        int rc = G.coef.ensure((size_t)total * 4);
        if (rc) return rc;
        fir_relayout_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(Tb.fir_taps.as<float>(), G.coef.as<float>(),
                                                                            n, T, d_begin, d_count, total);
        BF_CHECK_LAUNCH();
        count_launch();
        G.version = Tb.version; G.n = n; G.d_begin = d_begin; G.d_count = d_count; G.T = T; G.groups = groups;
    }

    FirParams fp{};
    fp.sig = d_sig; fp.img = d_img; fp.mic_ids = d_mics; fp.coef = G.coef.as<float>();
    fp.n = n; fp.n_mics_total = S.cfg.n_microphones; fp.d_begin = d_begin; fp.d_count = d_count; fp.frames = frames;
    fp.img_fs = lay.frame_stride; fp.img_ds = lay.dir_stride; fp.d_origin = lay.d_origin;
    fp.groups = groups; fp.T = T;
    fp.fn = (float)n; fp.inv_n = 1.0f / (float)n; fp.n_pow2 = (n & (n - 1)) == 0;

    const long total_groups = (long)groups * frames;
    int W = (int)((total_groups + S.sm_count - 1) / S.sm_count);
    W = W < 1 ? 1 : (W > kFirMaxWarps ? kFirMaxWarps : W);
    if (W > 4) W = ((W + 1 + 3) / 4 * 4 - 1) > kFirMaxWarps ? kFirMaxWarps : ((W + 1 + 3) / 4 * 4 - 1);
    fp.W = W;
    fp.tiles_per_frame = (groups + W - 1) / W;
    fp.total_tiles = fp.tiles_per_frame * frames;
    const int grid = fp.total_tiles < S.sm_count ? fp.total_tiles : S.sm_count;

    const size_t scratch_bytes = (size_t)W * 2 * kFirScratch * 4;
    const size_t budget = 227 * 1024 - 128 - scratch_bytes - 1024;
    const size_t per_mic = (size_t)(N + T) * 4 + (size_t)W * T * 8 * 4;
    int Mt = (int)(budget / kFirStages / per_mic);
    if (Mt < 1) {
        set_error(BF_ERR_CONFIG, "fir_tiled: N_TAPS %d does not fit the shared-memory ring", T);
        return BF_ERR_CONFIG;
    }
    Mt = Mt > 8 ? 8 : Mt;
    if (Mt > n) Mt = n;
    fp.Mt = Mt;
    const size_t smem = 128 + (size_t)kFirStages * Mt * per_mic + scratch_bytes;

    const bool exact = S.exact_sum != 0;
    const bool fused = S.cfg.fir_fused < 0 ? (T <= 16) : (S.cfg.fir_fused != 0);
    auto go = [&](auto kern) -> int {
        BF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, (W + 1) * 32, smem, st>>>(fp);
        BF_CHECK_LAUNCH();
        count_launch();
        return BF_OK;
    };
    if (algo == BF_ALGO_FIR_LANES) return exact ? go(das_fir_kernel<kFirLanes, true>) : go(das_fir_kernel<kFirLanes, false>);
    if (fused) return exact ? go(das_fir_kernel<kFirSeqFused, true>) : go(das_fir_kernel<kFirSeqFused, false>);
    return exact ? go(das_fir_kernel<kFirSeqUnfused, true>) : go(das_fir_kernel<kFirSeqUnfused, false>);
}

}  // namespace bf
