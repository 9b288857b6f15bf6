// das_miso.cu -- single-direction (MISO) beam output over a stream of sample blocks.
//
// Replaces miso_pad (algorithms/pad_and_sum.c:54-70), miso_lerp
// (algorithms/lerp_and_sum.c:67-92) and the post-scale of the audio loop
// (api.c:519-523) for a *batch* of consecutive N_SAMPLES blocks (BASELINE
// config C2: continuous 48.828 kHz stream).  Blocks are independent (the
// reference zero-pads every block, no history), so the batch is the grid.
//
// This path has 0.25 add per byte: it is HBM-bound.  Design: persistent CTAs, one
// per SM; a producer warp streams microphone rows with 1-D bulk TMA copies
// (cp.async.bulk -> SASS UBLKCP) into a ring of shared-memory stages of up to 32
// rows each; runs of consecutive microphone ids in the adaptive array are fetched
// with ONE copy (a fully populated array is a single 32 KiB copy per stage), which
// matters because every bulk copy carries a fixed issue cost.  N consumer threads
// each own one output sample and add the delayed rows in table order
// (bit-identical to the reference), then store the block with coalesced writes.
#include <string.h>

#include "bf_common.cuh"

namespace bf {

static constexpr int kMisoMaxStages = 16;   // barriers live in the first 256 bytes of smem

struct MisoParams {
    const float *sig;      // [blocks][n_mics_total][N]
    float *out;            // [blocks][N]
    const int *mic_ids;    // [n]
    const int *whole;      // table row (already offset), [n]
    const float *weight;   // lerp weights row, [n]
    int n, n_mics_total, N, blocks, Mt, stages;
    int copy_rows;         // max rows per bulk copy (run splitting)
    int scale;
    float fn, gain;
};

template <bool LERP>
__global__ void __launch_bounds__(512 + 32, 1) miso_stream_kernel(const MisoParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int N = p.N;
    const int cwarps = N >> 5;                      // consumer warps (N threads)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t *full = (uint64_t *)smem;              // [kMisoMaxStages]
    uint64_t *empty = full + kMisoMaxStages;        // [kMisoMaxStages]
    const int npad = (p.n + 3) & ~3;
    int *s_w = (int *)(smem + 256);                 // [npad]
    float *s_h = (float *)(s_w + npad);             // [npad]
    int *s_mic = (int *)(s_h + npad);               // [npad]
    const size_t tab_bytes = ((size_t)npad * 12 + 127) / 128 * 128;
    unsigned char *stages = smem + 256 + tab_bytes;
    const size_t stage_bytes = (size_t)p.Mt * N * 4;

    for (int m = threadIdx.x; m < npad; m += blockDim.x) {
        int w = m < p.n ? p.whole[m] : N;           // pad entries: delay N = no contribution
        s_w[m] = w < 0 ? 0 : w;
        s_h[m] = (LERP && m < p.n) ? p.weight[m] : 0.0f;
        s_mic[m] = m < p.n ? p.mic_ids[m] : 0;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; s++) {
            bfptx::mbar_init(&full[s], 1);
            bfptx::mbar_init(&empty[s], cwarps);
        }
        bfptx::fence_mbar_init();
    }
    __syncthreads();

    const int nchunks = (p.n + p.Mt - 1) / p.Mt;

    if (warp == cwarps) {
        // ---------------- producer warp ----------------
        int s = 0;
        uint32_t ph = 1;
        for (int b = blockIdx.x; b < p.blocks; b += gridDim.x) {
            const float *bs = p.sig + (size_t)b * p.n_mics_total * N;
            for (int c = 0; c < nchunks; c++) {
                const int m0 = c * p.Mt, cnt = min(p.Mt, p.n - m0);
                // lane r owns row r of the chunk; a lane is a run head when its microphone
                // does not directly follow the previous row's -> one copy per run
                const bool in = lane < cnt;
                const int mic = in ? s_mic[m0 + lane] : -2;
                const int prev = __shfl_up_sync(0xffffffffu, mic, 1);
                const bool head = in && (lane == 0 || mic != prev + 1 || (lane % p.copy_rows) == 0);
                const unsigned heads = __ballot_sync(0xffffffffu, head);
                bfptx::mbar_wait(&empty[s], ph);
                if (lane == 0) bfptx::mbar_arrive_expect_tx(&full[s], (uint32_t)(cnt * N * 4));
                __syncwarp();
                if (head) {
                    const unsigned rest = lane == 31 ? 0u : (heads >> (lane + 1));
                    const int next = rest ? lane + __ffs(rest) : cnt;
                    float *sb = (float *)(stages + (size_t)s * stage_bytes);
                    bfptx::bulk_g2s(sb + (size_t)lane * N, bs + (size_t)mic * N,
                                    (uint32_t)((next - lane) * N * 4), &full[s]);
                }
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
        return;
    }

    // ---------------- consumer threads: thread t owns output sample t ----------------
    const int t = threadIdx.x;
    int s = 0;
    uint32_t ph = 0;
    for (int b = blockIdx.x; b < p.blocks; b += gridDim.x) {
        float acc = 0.0f;
        for (int c = 0; c < nchunks; c++) {
            const int m0 = c * p.Mt, cnt = min(p.Mt, p.n - m0);
            bfptx::mbar_wait(&full[s], ph);
            const float *sb = (const float *)(stages + (size_t)s * stage_bytes);
            // 4 microphones per step: one 16-byte broadcast load of their delays (+ weights)
            for (int mm = 0; mm < cnt; mm += 4) {
                const int4 w4 = *(const int4 *)(s_w + m0 + mm);
                const int wv[4] = {w4.x, w4.y, w4.z, w4.w};
                float hv[4] = {0.f, 0.f, 0.f, 0.f};
                if (LERP) {
                    const float4 h4 = *(const float4 *)(s_h + m0 + mm);
                    hv[0] = h4.x; hv[1] = h4.y; hv[2] = h4.z; hv[3] = h4.w;
                }
                // branch-free: all loads of the step are issued first (index clamped into the
                // row), the adds are selected away where the reference's loop does not reach
                // (table padding beyond n carries delay N, i.e. never reaches)
                float v[4], u[4];
                int iv[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float *row = sb + (size_t)(mm + k) * N;
                    iv[k] = t - wv[k] - (LERP ? 1 : 0);      // pad_and_sum.c:43-46 / lerp_and_sum.c:52-55
                    const int ic = max(iv[k], 0);
                    v[k] = row[ic];
                    u[k] = LERP ? row[ic + 1] : 0.0f;
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float c = LERP ? __fmaf_rn(hv[k], __fsub_rn(u[k], v[k]), v[k]) : v[k];
                    const float sum = __fadd_rn(acc, c);
                    acc = iv[k] >= 0 ? sum : acc;
                }
            }
            __syncwarp();
            if (lane == 0) bfptx::mbar_arrive(&empty[s]);
            if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        if (p.scale) acc = __fmul_rn(__fdiv_rn(acc, p.fn), p.gain);   // api.c:519-523
        p.out[(size_t)b * N + t] = acc;
    }
}

// Alternative without staging: one CTA per block of samples, thread t reads its delayed
// sample of every microphone row straight from global memory (coalesced, 8 loads in flight
// per thread), table row in shared memory.  Same arithmetic order as above.
template <bool LERP>
__global__ void __launch_bounds__(512) miso_direct_kernel(const MisoParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int N = p.N, t = threadIdx.x;
    const int npad = (p.n + 7) & ~7;
    int *s_w = (int *)smem;
    float *s_h = (float *)(s_w + npad);
    int *s_mic = (int *)(s_h + npad);
    for (int m = t; m < npad; m += blockDim.x) {
        int w = m < p.n ? p.whole[m] : N;
        s_w[m] = w < 0 ? 0 : w;
        s_h[m] = (LERP && m < p.n) ? p.weight[m] : 0.0f;
        s_mic[m] = m < p.n ? p.mic_ids[m] : p.mic_ids[0];
    }
    __syncthreads();
    for (int b = blockIdx.x; b < p.blocks; b += gridDim.x) {
        const float *bs = p.sig + (size_t)b * p.n_mics_total * N;
        float acc = 0.0f;
        for (int m0 = 0; m0 < npad; m0 += 8) {
            float v[8], u[8];
            int iv[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float *row = bs + (size_t)s_mic[m0 + k] * N;
                iv[k] = t - s_w[m0 + k] - (LERP ? 1 : 0);
                const int ic = max(iv[k], 0);
                v[k] = __ldg(row + ic);
                u[k] = LERP ? __ldg(row + ic + 1) : 0.0f;
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float c = LERP ? __fmaf_rn(s_h[m0 + k], __fsub_rn(u[k], v[k]), v[k]) : v[k];
                const float sum = __fadd_rn(acc, c);
                acc = iv[k] >= 0 ? sum : acc;
            }
        }
        if (p.scale) acc = __fmul_rn(__fdiv_rn(acc, p.fn), p.gain);
        p.out[(size_t)b * N + t] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// FIR / hybrid MISO stream (miso_convolve_naive / _vectorized, convolve_and_sum.c:213-262, and
// miso_convolve_hybrid, hybrid_convolve_and_sum.c:66-90): same ring as above, but every row sits
// between T/2 zeros in shared memory (per-row bulk copies), so the tap loop needs no bounds checks;
// the steered direction's taps live in shared memory.  Thread t owns output sample t; the chain
// over (microphone, tap) runs in the reference's order and contraction.
// ---------------------------------------------------------------------------------------------
enum { kMisoFirFused = 0, kMisoFirUnfused = 1, kMisoFirLanes = 2, kMisoHybFused = 3, kMisoHybUnfused = 4 };

struct MisoFirParams {
    const float *sig; float *out; const int *mic_ids;
    const float *taps;     // [n][T] row of the steered direction (already offset)
    const int *whole;      // hybrid: integer delays [n] (already offset)
    int n, n_mics_total, N, T, blocks, Mt, stages, scale;
    float fn, gain;
};

template <int MODE>
__global__ void __launch_bounds__(512 + 32, 1) miso_fir_stream_kernel(const MisoFirParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int N = p.N, T = p.T, RS = N + T;
    const int cwarps = N >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t *full = (uint64_t *)smem;
    uint64_t *empty = full + kMisoMaxStages;
    float *s_taps = (float *)(smem + 256);                          // [n][T]
    int *s_w = (int *)(s_taps + (size_t)p.n * T);                   // [n]
    int *s_mic = s_w + p.n;                                         // [n]
    const size_t tab_bytes = (((size_t)p.n * T + 2 * (size_t)p.n) * 4 + 127) / 128 * 128;
    float *stages = (float *)(smem + 256 + tab_bytes);
    const size_t stage_floats = (size_t)p.Mt * RS;

    for (int i = threadIdx.x; i < p.n * T; i += blockDim.x) s_taps[i] = p.taps[i];
    for (int m = threadIdx.x; m < p.n; m += blockDim.x) {
        int w = p.whole ? p.whole[m] : 0;
        s_w[m] = w < 0 ? 0 : w;
        s_mic[m] = p.mic_ids[m];
    }
    for (size_t i = threadIdx.x; i < (size_t)p.stages * p.Mt * T; i += blockDim.x) {   // the zero borders, once
        const size_t row = i / T;
        const int c = (int)(i - row * T);
        stages[row * RS + (c < T / 2 ? c : N + c)] = 0.0f;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; s++) {
            bfptx::mbar_init(&full[s], 1);
            bfptx::mbar_init(&empty[s], cwarps);
        }
        bfptx::fence_mbar_init();
    }
    __syncthreads();
    const int nchunks = (p.n + p.Mt - 1) / p.Mt;

    if (warp == cwarps) {
        int s = 0;
        uint32_t ph = 1;
        for (int b = blockIdx.x; b < p.blocks; b += gridDim.x) {
            const float *bs = p.sig + (size_t)b * p.n_mics_total * N;
            for (int c = 0; c < nchunks; c++) {
                const int m0 = c * p.Mt, cnt = min(p.Mt, p.n - m0);
                bfptx::mbar_wait(&empty[s], ph);
                if (lane == 0) bfptx::mbar_arrive_expect_tx(&full[s], (uint32_t)(cnt * N * 4));
                __syncwarp();
                if (lane < cnt)
                    bfptx::bulk_g2s(stages + (size_t)s * stage_floats + (size_t)lane * RS + T / 2,
                                    bs + (size_t)s_mic[m0 + lane] * N, (uint32_t)(N * 4), &full[s]);
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
        return;
    }

    const int t = threadIdx.x;
    int s = 0;
    uint32_t ph = 0;
    for (int b = blockIdx.x; b < p.blocks; b += gridDim.x) {
        float acc = 0.0f;
        for (int c = 0; c < nchunks; c++) {
            const int m0 = c * p.Mt, cnt = min(p.Mt, p.n - m0);
            bfptx::mbar_wait(&full[s], ph);
            const float *sb = stages + (size_t)s * stage_floats;
            for (int mm = 0; mm < cnt; mm++) {
                const float *h = s_taps + (size_t)(m0 + mm) * T;
                // padded[j] = row[j - T/2]: the stored row starts T/2 floats into its slot
                if (MODE == kMisoFirFused || MODE == kMisoFirUnfused) {
                    const float *pr = sb + (size_t)mm * RS + t;              // padded[t + k]
                    for (int k = 0; k < T; k += 4) {
                        const float4 hk = *(const float4 *)(h + k);
                        const float hv[4] = {hk.x, hk.y, hk.z, hk.w};
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            acc = MODE == kMisoFirFused ? __fmaf_rn(hv[u], pr[k + u], acc)
                                                        : __fadd_rn(acc, __fmul_rn(hv[u], pr[k + u]));
                    }
                } else if (MODE == kMisoFirLanes) {
                    const float *pr = sb + (size_t)mm * RS + t;
                    float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    for (int k = 0; k < T; k += 8) {
                        const float4 h0 = *(const float4 *)(h + k), h1 = *(const float4 *)(h + k + 4);
                        const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
                        for (int j = 0; j < 8; j++) x[j] = __fmaf_rn(pr[k + j], hv[j], x[j]);
                    }
                    const float q0 = __fadd_rn(x[0], x[4]), q1 = __fadd_rn(x[1], x[5]);
                    const float q2 = __fadd_rn(x[2], x[6]), q3 = __fadd_rn(x[3], x[7]);
                    acc = __fadd_rn(acc, __fadd_rn(__fadd_rn(q0, q2), __fadd_rn(q1, q3)));
                } else {
                    // hybrid: out[w + i + 1] += sum_k h[k] padded[i + k], i = t - w - 1 >= 0
                    const int i0 = t - s_w[m0 + mm] - 1;
                    const float *pr = sb + (size_t)mm * RS + max(i0, 0);
                    float a2 = acc;
                    for (int k = 0; k < T; k += 4) {
                        const float4 hk = *(const float4 *)(h + k);
                        const float hv[4] = {hk.x, hk.y, hk.z, hk.w};
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            a2 = MODE == kMisoHybFused ? __fmaf_rn(hv[u], pr[k + u], a2)
                                                       : __fadd_rn(a2, __fmul_rn(hv[u], pr[k + u]));
                    }
                    acc = i0 >= 0 ? a2 : acc;
                }
            }
            __syncwarp();
            if (lane == 0) bfptx::mbar_arrive(&empty[s]);
            if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        if (p.scale) acc = __fmul_rn(__fdiv_rn(acc, p.fn), p.gain);
        p.out[(size_t)b * N + t] = acc;
    }
}

static int miso_fir_run(int algo, const float *d_sig, float *d_out, int blocks, const int *d_mics, int n,
                        int offset, int scale, cudaStream_t st)
{
    State &S = state();
    Tables &Tb = S.tab;
    const int N = S.cfg.n_samples, T = S.cfg.n_taps;
    MisoFirParams mp{};
    // offset units follow the reference: entries for the hybrid tables (hybrid_convolve_and_sum.c:80-82),
    // FLOATS of the tap table for the FIR ones (convolve_and_sum.c:227: coefficients + offset + m*N_TAPS)
    size_t have, off_e = (size_t)(offset < 0 ? 0 : offset);
    if (algo == BF_ALGO_HYBRID) {
        mp.taps = Tb.hyb_taps.as<float>(); mp.whole = Tb.hyb_whole.as<int>(); have = Tb.hyb_count;
    } else {
        if (offset % T != 0) { set_error(BF_ERR_ARG, "FIR offset %d not a multiple of N_TAPS", offset); return BF_ERR_ARG; }
        mp.taps = Tb.fir_taps.as<float>(); have = Tb.fir_count / (size_t)T;
        off_e /= (size_t)T;
    }
    if (offset < 0 || have < off_e + n || mp.taps == nullptr) {
        set_error(BF_ERR_NOT_LOADED, "miso: table holds %zu entries, need offset+n = %zu+%d", have, off_e, n);
        return BF_ERR_NOT_LOADED;
    }
    mp.taps += off_e * T;
    if (mp.whole) mp.whole += off_e;
    mp.sig = d_sig; mp.out = d_out; mp.mic_ids = d_mics;
    mp.n = n; mp.n_mics_total = S.cfg.n_microphones; mp.N = N; mp.T = T; mp.blocks = blocks;
    mp.scale = scale; mp.fn = (float)n; mp.gain = S.cfg.mic_gain;
    const size_t tab_bytes = (((size_t)n * T + 2 * (size_t)n) * 4 + 127) / 128 * 128;
    int ctas = 3;
    while (ctas > 1 && (N + 32) * ctas > 2048) ctas--;
    size_t budget = 0;
    for (; ctas >= 1; ctas--) {
        const size_t per = (size_t)(227 * 1024) / ctas;
        if (per > 2048 + 256 + tab_bytes + 2 * 4 * (size_t)(N + T) * 4) { budget = per - 2048 - 256 - tab_bytes; break; }
    }
    if (!budget) { set_error(BF_ERR_CONFIG, "miso FIR: tap table does not fit shared memory"); return BF_ERR_CONFIG; }
    int Mt = n < 32 ? n : 32;
    while (Mt > 1 && (size_t)Mt * (N + T) * 4 * 2 > budget) Mt--;
    mp.Mt = Mt; mp.stages = 2;
    const size_t smem = 256 + tab_bytes + (size_t)2 * Mt * (N + T) * 4;
    const int grid = blocks < S.sm_count * ctas ? blocks : S.sm_count * ctas;
    const bool fused = S.cfg.fir_fused < 0 ? (T <= 16) : (S.cfg.fir_fused != 0);
    auto go = [&](auto kern) -> int {
        BF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, N + 32, smem, st>>>(mp);
        BF_CHECK_LAUNCH();
        count_launch();
        return BF_OK;
    };
    if (algo == BF_ALGO_FIR_LANES) return go(miso_fir_stream_kernel<kMisoFirLanes>);
    if (algo == BF_ALGO_HYBRID) return fused ? go(miso_fir_stream_kernel<kMisoHybFused>) : go(miso_fir_stream_kernel<kMisoHybUnfused>);
    return fused ? go(miso_fir_stream_kernel<kMisoFirFused>) : go(miso_fir_stream_kernel<kMisoFirUnfused>);
}

int miso_run(int algo, const float *d_sig, float *d_out, int blocks, const int *d_mics, int n,
             int offset, int by_mic_id, int scale, cudaStream_t st)
{
    State &S = state();
    const int N = S.cfg.n_samples;
    const bool streamable = (algo == BF_ALGO_PAD || algo == BF_ALGO_LERP) && !by_mic_id &&
                            !S.simple_kernel && N % 32 == 0 && N >= 32 && N <= 512 &&
                            blocks >= 8 && (((uintptr_t)d_sig & 15) == 0);
    const bool fir_like = algo == BF_ALGO_FIR_SEQ || algo == BF_ALGO_FIR_LANES || algo == BF_ALGO_HYBRID;
    if (fir_like && !by_mic_id && !S.simple_kernel && N % 32 == 0 && N >= 32 && N <= 512 && blocks >= 8 &&
        (((uintptr_t)d_sig & 15) == 0) && S.cfg.n_taps % 8 == 0 && S.cfg.n_taps >= 8)
        return miso_fir_run(algo, d_sig, d_out, blocks, d_mics, n, offset, scale, st);
    if (!streamable)
        return miso_simple(algo, d_sig, d_out, blocks, d_mics, n, offset, by_mic_id, scale, st);

    Tables &T = S.tab;
    MisoParams mp{};
    size_t have;
    if (algo == BF_ALGO_PAD) { mp.whole = T.pad_whole.as<int>(); have = T.pad_count; }
    else { mp.whole = T.lerp_whole.as<int>(); mp.weight = T.lerp_weight.as<float>(); have = T.lerp_count; }
    if (offset < 0 || have < (size_t)offset + n || mp.whole == nullptr) {
        set_error(BF_ERR_NOT_LOADED, "miso: table holds %zu entries, need offset+n = %d+%d", have,
                  offset, n);
        return BF_ERR_NOT_LOADED;
    }
    mp.whole += offset;
    if (mp.weight) mp.weight += offset;
    mp.sig = d_sig; mp.out = d_out; mp.mic_ids = d_mics;
    mp.n = n; mp.n_mics_total = S.cfg.n_microphones; mp.N = N; mp.blocks = blocks;
    mp.scale = scale; mp.fn = (float)n; mp.gain = S.cfg.mic_gain;

    if (const char *e = getenv("BF_MISO_MODE")) {
        if (!strcmp(e, "direct") && N >= 2) {
            const int np8 = (n + 7) & ~7;
            const int per_sm = 2048 / N < 1 ? 1 : 2048 / N;
            int grid = S.sm_count * per_sm;
            if (grid > blocks) grid = blocks;
            if (algo == BF_ALGO_LERP) miso_direct_kernel<true><<<grid, N, (size_t)np8 * 12, st>>>(mp);
            else miso_direct_kernel<false><<<grid, N, (size_t)np8 * 12, st>>>(mp);
            BF_CHECK_LAUNCH();
            count_launch();
            return BF_OK;
        }
    }
    const int npad = (n + 3) & ~3;
    const size_t tab_bytes = ((size_t)npad * 12 + 127) / 128 * 128;
    // several CTAs per SM: the consumer is a chain of dependent shared-memory loads, so the
    // kernel needs many resident warps to keep enough bulk copies in flight (measured)
    // Measured on B200 (tools/miso_sweep.py, C2): contiguous microphone ids (one 32 KiB copy per
    // stage) -> 2 CTAs/SM x 2 stages: 6.9 TB/s pad, 6.8 TB/s lerp; scattered ids (1 KiB copies)
    // -> 3 CTAs/SM x 2 stages: 6.8 / 5.9 TB/s.
    // whether the microphone ids are mostly consecutive decides the launch shape; the ids are read back once per
    // (pointer, n).  While the stream is being captured into a CUDA graph a blocking copy is not allowed: an
    // unknown array is then treated as scattered (3 CTAs per SM, per-row copies: correct for any ids).
    static const int *s_key_ptr = nullptr; static int s_key_n = -1; static bool s_contig = false;
    bool contig = s_contig;
    if (s_key_ptr != d_mics || s_key_n != n) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
        if (cap == cudaStreamCaptureStatusNone) {
            std::vector<int> h(n);
            BF_CUDA(cudaMemcpyAsync(h.data(), d_mics, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
            BF_CUDA(cudaStreamSynchronize(st));
            int breaks = 0;
            for (int m = 1; m < n; m++) breaks += (h[m] != h[m - 1] + 1);
            s_contig = contig = breaks * 8 <= n;            // mostly runs of >= 8 rows
            s_key_ptr = d_mics; s_key_n = n;
        } else {
            contig = false;
        }
    }
    int ctas = contig ? 2 : 3;
    while (ctas > 1 && (N + 32) * ctas > 2048) ctas--;
    if (const char *e = getenv("BF_MISO_CTAS")) { int v = atoi(e); if (v >= 1 && v <= 4 && (N + 32) * v <= 2048) ctas = v; }
    const size_t budget = (size_t)(227 * 1024) / ctas - 1024 - 256 - tab_bytes - 1024;
    // stages of up to 32 rows (one producer lane per row, multiple of 4 rows), as many as fit
    int Mt = npad < 32 ? npad : 32;
    if (const char *e = getenv("BF_MISO_MT")) { int v = atoi(e) & ~3; if (v >= 4 && v <= Mt) Mt = v; }
    while (Mt > 4 && (size_t)Mt * N * 4 * 2 > budget) Mt -= 4;
    int stages = (int)(budget / ((size_t)Mt * N * 4));
    if (stages > 2) stages = 2;                // deeper rings measured no faster (consumer-paced)
    if (const char *e = getenv("BF_MISO_STAGES")) { int v = atoi(e); if (v >= 2 && v <= stages) stages = v; }
    mp.copy_rows = 32;
    if (const char *e = getenv("BF_MISO_COPY_ROWS")) { int v = atoi(e); if (v >= 1 && v <= 32) mp.copy_rows = v; }
    if (stages < 2) {
        set_error(BF_ERR_CONFIG, "miso: shared memory budget exceeded");
        return BF_ERR_CONFIG;
    }
    mp.Mt = Mt; mp.stages = stages;
    const size_t smem = 256 + tab_bytes + (size_t)stages * Mt * N * 4;
    const int grid = blocks < S.sm_count * ctas ? blocks : S.sm_count * ctas;
    if (algo == BF_ALGO_LERP) {
        BF_CUDA(cudaFuncSetAttribute(miso_stream_kernel<true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        miso_stream_kernel<true><<<grid, N + 32, smem, st>>>(mp);
    } else {
        BF_CUDA(cudaFuncSetAttribute(miso_stream_kernel<false>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        miso_stream_kernel<false><<<grid, N + 32, smem, st>>>(mp);
    }
    BF_CHECK_LAUNCH();
    count_launch();
    return BF_OK;
}

}  // namespace bf
