// das_miso.cu -- single-direction (MISO) beam output over a stream of sample blocks.
//
// Replaces miso_pad (algorithms/pad_and_sum.c:54-70), miso_lerp
// (algorithms/lerp_and_sum.c:67-92) and the post-scale of the audio loop
// (api.c:519-523) for a *batch* of consecutive N_SAMPLES blocks (BASELINE
// config C2: continuous 48.828 kHz stream).  Blocks are independent (the
// reference zero-pads every block, no history), so the batch is the grid.
//
// This path has 0.25 add per byte: it is HBM-bound.  Design: persistent CTAs, one
// per SM; a producer warp streams whole microphone rows (1 KiB each at N = 256)
// with 1-D bulk TMA copies into a 3-stage shared-memory ring (up to 64 KiB per
// stage -> ~190 KiB of loads in flight per SM); N consumer threads each own one
// output sample and add the delayed rows in table order (bit-identical to the
// reference), then store the block with coalesced writes.
#include "bf_common.cuh"

namespace bf {

int miso_simple(int algo, const float *d_sig, float *d_out, int blocks, const int *d_mics, int n,
                int offset, int by_mic_id, int scale, cudaStream_t st);

static constexpr int kMisoStages = 3;

struct MisoParams {
    const float *sig;      // [blocks][n_mics_total][N]
    float *out;            // [blocks][N]
    const int *mic_ids;    // [n]
    const int *whole;      // table row (already offset), [n]
    const float *weight;   // lerp weights row, [n]
    int n, n_mics_total, N, blocks, Mt;
    int scale;
    float fn, gain;
};

template <bool LERP>
__global__ void __launch_bounds__(512 + 32, 1) miso_stream_kernel(const MisoParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int N = p.N;
    const int nthreads_c = N;                       // consumer threads (multiple of 32)
    const int cwarps = nthreads_c >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t *full = (uint64_t *)smem;
    uint64_t *empty = full + kMisoStages;
    int *s_w = (int *)(smem + 128);                 // [n]
    float *s_h = (float *)(s_w + p.n);              // [n]
    int *s_mic = (int *)(s_h + p.n);                // [n]
    const size_t tab_bytes = ((size_t)p.n * 12 + 127) / 128 * 128;
    unsigned char *stages = smem + 128 + tab_bytes;
    const size_t stage_bytes = (size_t)p.Mt * N * 4;

    for (int m = threadIdx.x; m < p.n; m += blockDim.x) {
        int w = p.whole[m];
        s_w[m] = w < 0 ? 0 : w;
        s_h[m] = LERP ? p.weight[m] : 0.0f;
        s_mic[m] = p.mic_ids[m];
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < kMisoStages; s++) {
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
                bfptx::mbar_wait(&empty[s], ph);
                if (lane == 0) bfptx::mbar_arrive_expect_tx(&full[s], (uint32_t)(cnt * N * 4));
                __syncwarp();
                float *sb = (float *)(stages + (size_t)s * stage_bytes);
                for (int r = lane; r < cnt; r += 32)
                    bfptx::bulk_g2s(sb + (size_t)r * N, bs + (size_t)s_mic[m0 + r] * N, N * 4,
                                    &full[s]);
                if (++s == kMisoStages) { s = 0; ph ^= 1; }
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
#pragma unroll 4
            for (int mm = 0; mm < cnt; mm++) {
                const float *row = sb + (size_t)mm * N;
                if (LERP) {
                    const int i = t - s_w[m0 + mm] - 1;          // lerp_and_sum.c:52-55
                    if (i >= 0) {
                        const float a = row[i], bb = row[i + 1];
                        acc = __fadd_rn(acc, __fmaf_rn(s_h[m0 + mm], __fsub_rn(bb, a), a));
                    }
                } else {
                    const int i = t - s_w[m0 + mm];              // pad_and_sum.c:43-46
                    if (i >= 0) acc = __fadd_rn(acc, row[i]);
                }
            }
            __syncwarp();
            if (lane == 0) bfptx::mbar_arrive(&empty[s]);
            if (++s == kMisoStages) { s = 0; ph ^= 1; }
        }
        if (p.scale) acc = __fmul_rn(__fdiv_rn(acc, p.fn), p.gain);   // api.c:519-523
        p.out[(size_t)b * N + t] = acc;
    }
}

int miso_run(int algo, const float *d_sig, float *d_out, int blocks, const int *d_mics, int n,
             int offset, int by_mic_id, int scale, cudaStream_t st)
{
    State &S = state();
    const int N = S.cfg.n_samples;
    const bool streamable = (algo == BF_ALGO_PAD || algo == BF_ALGO_LERP) && !by_mic_id &&
                            !S.simple_kernel && N % 32 == 0 && N >= 32 && N <= 512 &&
                            blocks >= 8 && (((uintptr_t)d_sig & 15) == 0) && (N * 4) % 16 == 0;
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

    const size_t tab_bytes = ((size_t)n * 12 + 127) / 128 * 128;
    const size_t budget = 227 * 1024 - 128 - tab_bytes - 1024;
    int Mt = n;
    while (Mt > 1 && (size_t)Mt * N * 4 * kMisoStages > budget) Mt = (Mt + 1) / 2;
    if ((size_t)Mt * N * 4 * kMisoStages > budget) {
        set_error(BF_ERR_CONFIG, "miso: shared memory budget exceeded");
        return BF_ERR_CONFIG;
    }
    mp.Mt = Mt;
    const size_t smem = 128 + tab_bytes + (size_t)kMisoStages * Mt * N * 4;
    const int grid = blocks < S.sm_count ? blocks : S.sm_count;
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
