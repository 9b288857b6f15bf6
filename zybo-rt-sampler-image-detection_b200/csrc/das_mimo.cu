// das_mimo.cu -- time-domain delay-and-sum power maps (MIMO), tiled TMA kernel.
//
// Replaces mimo_pad (algorithms/pad_and_sum.c:100-143), mimo_lerp
// (algorithms/lerp_and_sum.c:103-136) and the inline trunc-and-sum copy
// (api.c:1014-1062) of the reference:
//
//   img[d] = 1/N * sum_t ( 1/n * sum_m  delayed_m,d[t] )^2
//
// Design (B200 / sm_100a):
//   * persistent grid, one CTA per SM; a CTA = 1 producer warp + W consumer warps (W <= 19 pad, 15 lerp)
//   * a consumer warp owns a GROUP of R = 8 consecutive directions and all N
//     samples of them in registers (lane l holds samples l, l+32, ...), so the
//     microphone sum runs in the reference's order (m = 0..n-1, fp32) and the
//     summed block out[t] is bit-identical to the CPU reference
//   * microphone rows are streamed through a 4-stage shared-memory ring by the
//     producer warp with 1-D bulk TMA copies (cp.async.bulk, SASS UBLKCP)
//     completing on mbarriers; each row is stored behind a block of zeros so a
//     delayed read is just an offset load (no predicates in the inner loop)
//   * consecutive directions mostly share their integer delay for a microphone
//     (the table varies slowly along the grid), so one shifted row load is
//     reused for all directions of the group with the same delay: shared-memory
//     traffic drops by ~R and the kernel becomes FP32-pipe bound; the adds are
//     issued as packed add.f32x2 / fma.f32x2 (FADD2/FFMA2: plain 3-register FADD
//     only reaches half the FP32 rate on this SM, measured)
//   * the table is re-laid out once per load into 16-byte "group entries" per
//     (group, microphone), classified as UNIFORM (one delay for all 8 directions),
//     TWO-RUN (delay changes once inside the group) or GENERAL; the next chunk's 32
//     entries (and lerp weights) of a warp arrive in its private, double-buffered
//     shared-memory slot by cp.async while the current chunk is processed
//   * latency is hidden by warps, not registers: pad runs 19 consumer warps at 96
//     registers (five per scheduler) without a register row prefetch -- measured
//     faster than 15 warps at 128 registers with one (profiles/r1_kernel_variants.md; the
//     prefetching loop is in the git history); a microphone is processed sample-pair-outer,
//     so only two row values are live next to the 64 accumulators
//   * GATHER instantiation (direction-sharded multi-GPU runs): the epilogue also stores every
//     value into the same position of the peer GPUs' buffers (CUDA-IPC mapped, NVLink P2P
//     stores) and the step flags of that exchange are handled inside the kernel -- the
//     all-gather is fused into the map kernel
//   * the epilogue reproduces out/n, square, in-order sum over t, /N exactly
//     (exact_sum) or uses a warp-shuffle tree
#include "bf_common.cuh"

namespace bf {

static constexpr int kR = 8;           // directions per warp
static constexpr int kStages = 4;      // smem ring depth
// Consumer warps per CTA (+1 producer).  pad: 19 + 1 = 640 threads at 96 registers -- five warps per
// scheduler hide the per-microphone latencies better than a register-hungry row prefetch does with four
// (measured: 10 302 vs 10 022 maps/s); lerp keeps 15 + 1 = 512 threads at 128 registers (measured: 5 772
// vs 5 008 maps/s with 19 + 1 at 96).
static constexpr int kMaxWarpsPad = 19, kMaxWarpsLerp = 15;
__host__ __device__ constexpr int max_warps(bool lerp) { return lerp ? kMaxWarpsLerp : kMaxWarpsPad; }
static constexpr int kScratchStride = 36;   // floats per direction row of the epilogue scratch (32 + pad)
// per-warp shared slot = two buffers; a buffer holds the staged entries of one chunk (uint4[32], + float[32][8]
// weights for lerp) and doubles as the epilogue scratch (8 rows x 36 floats) once its chunk is consumed
__host__ __device__ constexpr int slot_buf_bytes(bool lerp) { return lerp ? 1536 : 1152; }

enum { kGeneral = 0u, kUniform = 1u, kTwoRun = 2u };

// ---------------------------------------------------------------------------
// group-table builder
// ---------------------------------------------------------------------------
// One thread per (group, mic).  off = (P - w [- 1 for lerp]) * 4 bytes, so that
// row_base + off + 4*t addresses sample (t - w [- 1]) of a row stored behind P
// zeros.  Offsets are multiples of 4: bits 1:0 of the first u16 carry the kind.
//   GENERAL  x = o0|o1<<16, y = o2|o3<<16, z = o4|o5<<16, w = o6|o7<<16
//   UNIFORM  x = o0|1
//   TWO-RUN  x = o0|2 | ob<<16, y = split   (directions [0,split) use o0, the rest ob)
__global__ void build_groups_kernel(const int *__restrict__ whole, const float *__restrict__ weight,
                                    uint4 *__restrict__ offs, float *__restrict__ wts, int n,
                                    int d_begin, int d_count, int groups, int P, int w_hi,
                                    int lerp)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= groups * n) return;
    int g = idx / n, m = idx - g * n;
    uint32_t o[kR];
#pragma unroll
    for (int r = 0; r < kR; r++) {
        int dl = g * kR + r;
        if (dl >= d_count) dl = d_count - 1;            // tail group: replicate last direction
        size_t e = (size_t)(d_begin + dl) * n + m;
        int w = whole[e];
        w = w < 0 ? 0 : (w > w_hi ? w_hi : w);
        o[r] = (uint32_t)(P - w - (lerp ? 1 : 0)) * 4u;
        if (lerp) wts[(size_t)idx * kR + r] = weight[e];
    }
    int split = 1;
    while (split < kR && o[split] == o[0]) split++;
    bool two = split < kR;
    for (int r = split; r < kR; r++)
        if (o[r] != o[split < kR ? split : 0]) two = false;
    uint4 v;
    if (split == kR) {
        v = make_uint4(o[0] | kUniform, 0u, 0u, 0u);
    } else if (two) {
        v = make_uint4(o[0] | kTwoRun | (o[split] << 16), (uint32_t)split, 0u, 0u);
    } else {
        v.x = o[0] | (o[1] << 16);
        v.y = o[2] | (o[3] << 16);
        v.z = o[4] | (o[5] << 16);
        v.w = o[6] | (o[7] << 16);
    }
    offs[idx] = v;
}

static int build_groups(GroupTable &gt, const int *d_whole, const float *d_weight, int n,
                        int d_begin, int d_count, int P, int n_samples, bool lerp,
                        cudaStream_t st)
{
    int groups = (d_count + kR - 1) / kR;
    size_t entries = (size_t)groups * n;
    int rc = gt.offs.ensure(entries * sizeof(uint4) + 512);     // +512: prefetch overrun pad
    if (rc) return rc;
    if (lerp) {
        rc = gt.wts.ensure(entries * kR * sizeof(float) + 1024);
        if (rc) return rc;
    }
    int w_hi = lerp ? n_samples - 1 : n_samples;
    int threads = 256;
    int blocks = (int)((entries + threads - 1) / threads);
    build_groups_kernel<<<blocks, threads, 0, st>>>(d_whole, d_weight, gt.offs.as<uint4>(),
                                                    gt.wts.as<float>(), n, d_begin, d_count,
                                                    groups, P, w_hi, lerp ? 1 : 0);
    BF_CHECK_LAUNCH();
    count_launch();
    gt.n = n; gt.d_begin = d_begin; gt.d_count = d_count; gt.pad = P; gt.n_samples = n_samples;
    gt.groups = groups;
    return BF_OK;
}

// lerp needs the per-row first difference s[i+1]-s[i] (lerp_and_sum.c:54); it is
// computed once per frame in table-column order so that the main kernel can
// bulk-copy it like a signal row.  diff[f][m][N-1] = 0 (never read).
__global__ void diff_rows_kernel(const float *__restrict__ sig, const int *__restrict__ mic_ids,
                                 float *__restrict__ diff, int n, int N, int n_mics_total)
{
    int f = blockIdx.y, m = blockIdx.x;
    const float *row = sig + ((size_t)f * n_mics_total + mic_ids[m]) * N;
    float *out = diff + ((size_t)f * n + m) * N;
    for (int i = threadIdx.x; i < N; i += blockDim.x)
        out[i] = (i + 1 < N) ? __fsub_rn(row[i + 1], row[i]) : 0.0f;
}

// ---------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------
struct MimoParams {
    const float *sig;        // [frames][n_mics_total][N]
    const float *diff;       // [frames][n][N] (lerp only)
    float *img;              // output, see img_fs / img_ds
    const int *mic_ids;      // [n]
    const uint4 *offs;       // [groups][n]
    const float *wts;        // [groups][n][8]
    int n, n_mics_total, d_begin, d_count, frames;
    long img_fs, img_ds;     // output strides (frame, direction)
    int d_origin;
    float *peers[7];         // fused all-gather targets (peer GPUs), same layout as img
    int n_peers;
    long long *flags_local; long long wait_seq;      // see ImgLayout
    long long *flags_all[8]; int world, flag_rank; long long signal_seq;
    unsigned int *done_counter; int *timed_out;
    int groups;              // groups per frame
    int tiles_per_frame, total_tiles;
    int W;                   // consumer warps
    int Mt;                  // mic rows per stage (<= 32)
    int P;                   // zero floats in front of every row
    int n_pow2;              // n is a power of two -> exact reciprocal multiply
    float fn, inv_n;
};

// acc += a            (pad:  pad_and_sum.c:45)
// acc += fma(h, b, a) (lerp: lerp_and_sum.c:54, b = s[i+1]-s[i]); PACK selects the packed
// add.f32x2 / fma.f32x2 forms -- same IEEE rounding per element either way.
template <bool LERP, bool PACK>
__device__ __forceinline__ float2 das_accum(float2 acc, float2 a, float2 b, float h)
{
    if (PACK) {
        if (LERP) return __fadd2_rn(acc, __ffma2_rn(make_float2(h, h), b, a));
        return __fadd2_rn(acc, a);
    }
    if (LERP)
        return make_float2(__fadd_rn(acc.x, __fmaf_rn(h, b.x, a.x)),
                           __fadd_rn(acc.y, __fmaf_rn(h, b.y, a.y)));
    return make_float2(__fadd_rn(acc.x, a.x), __fadd_rn(acc.y, a.y));
}

// Sample-pair-outer form of one microphone (used by the shipped loops): only the two row values (and, for
// lerp, the two differences) of ONE sample pair are live at a time instead of a whole row, which is what
// lets the kernel run at 96 registers: acc[r][q] += row value (pad) or fma(h[r], diff, row value) (lerp),
// microphones in table order, so every accumulator sees the reference's sequence of roundings.
template <int J>
__device__ __forceinline__ float2 ld_pair(const char *p, int q)
{
    return make_float2(*(const float *)(p + q * 256), *(const float *)(p + q * 256 + 128));
}

template <int J, bool LERP, bool PACK, int S>
__device__ __forceinline__ void mic_pairs(float2 (&acc)[kR][J / 2], const char *pa, const char *pb,
                                          const uint32_t row_bytes, const float (&h)[kR])
{
#pragma unroll
    for (int q = 0; q < J / 2; q++) {
        const float2 a = ld_pair<J>(pa, q);
        const float2 d = LERP ? ld_pair<J>(pa + row_bytes, q) : a;
        float2 a2 = a, d2 = d;
        if (S < kR) {
            a2 = ld_pair<J>(pb, q);
            d2 = LERP ? ld_pair<J>(pb + row_bytes, q) : a2;
        }
#pragma unroll
        for (int r = 0; r < kR; r++)
            acc[r][q] = (r < S) ? das_accum<LERP, PACK>(acc[r][q], a, d, h[r])
                                : das_accum<LERP, PACK>(acc[r][q], a2, d2, h[r]);
    }
}

template <int J, bool LERP, bool PACK>
__device__ __forceinline__ void process_mic_q(float2 (&acc)[kR][J / 2], const uint2 e, const uint4 *efull,
                                              const char *rowp, const uint32_t row_bytes, const float *wrow)
{
    const uint32_t kind = e.x & 3u;
    const char *pa = rowp + (e.x & 0xfffcu);
    float h[kR] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (LERP) {
        const float4 h0 = *(const float4 *)wrow, h1 = *(const float4 *)(wrow + 4);
        h[0] = h0.x; h[1] = h0.y; h[2] = h0.z; h[3] = h0.w;
        h[4] = h1.x; h[5] = h1.y; h[6] = h1.z; h[7] = h1.w;
    }
    if (kind == kUniform) {
        mic_pairs<J, LERP, PACK, kR>(acc, pa, pa, row_bytes, h);
    } else if (kind == kTwoRun) {
        const char *pb = rowp + (e.x >> 16);
        switch (e.y) {
            case 1: mic_pairs<J, LERP, PACK, 1>(acc, pa, pb, row_bytes, h); break;
            case 2: mic_pairs<J, LERP, PACK, 2>(acc, pa, pb, row_bytes, h); break;
            case 3: mic_pairs<J, LERP, PACK, 3>(acc, pa, pb, row_bytes, h); break;
            case 4: mic_pairs<J, LERP, PACK, 4>(acc, pa, pb, row_bytes, h); break;
            case 5: mic_pairs<J, LERP, PACK, 5>(acc, pa, pb, row_bytes, h); break;
            case 6: mic_pairs<J, LERP, PACK, 6>(acc, pa, pb, row_bytes, h); break;
            default: mic_pairs<J, LERP, PACK, 7>(acc, pa, pb, row_bytes, h); break;
        }
    } else {
        // general (0.4 % of the C3 table): one direction at a time
        const uint4 ef = *efull;
#pragma unroll
        for (int r = 0; r < kR; r++) {
            const uint32_t wd = (r >> 1) == 0 ? ef.x : ((r >> 1) == 1 ? ef.y : ((r >> 1) == 2 ? ef.z : ef.w));
            const char *pr = rowp + ((r & 1) ? (wd >> 16) : (wd & 0xfffcu));
#pragma unroll
            for (int q = 0; q < J / 2; q++) {
                const float2 a = ld_pair<J>(pr, q);
                const float2 d = LERP ? ld_pair<J>(pr + row_bytes, q) : a;
                acc[r][q] = das_accum<LERP, PACK>(acc[r][q], a, d, h[r]);
            }
        }
    }
}

template <int J, bool LERP, bool EXACT, bool PACK, bool GATHER>
__global__ void __launch_bounds__((max_warps(LERP) + 1) * 32, 1) das_mimo_kernel(const MimoParams p)
{
    constexpr int N = J * 32;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = p.W;
    const int RS = p.P + N;                              // row stride in floats
    const int arrays = LERP ? 2 : 1;
    const size_t row_bytes = (size_t)RS * 4;
    const size_t mic_bytes = arrays * row_bytes;         // one microphone's slot in a stage
    const size_t stage_bytes = (size_t)p.Mt * mic_bytes;

    uint64_t *full = (uint64_t *)smem;                   // [kStages]
    uint64_t *empty = full + kStages;                    // [kStages]
    unsigned int *warps_done = (unsigned int *)(smem + 224);   // fused gather: consumer warps finished
    unsigned char *stages = smem + 256;
    float *scratch_all = (float *)(stages + kStages * stage_bytes);

    // ---- prologue: zero the pad columns once, init barriers ----------------
    {
        const int rows_total = kStages * p.Mt * arrays;
        for (int i = threadIdx.x; i < rows_total * p.P; i += blockDim.x) {
            int row = i / p.P, c = i - row * p.P;
            ((float *)(stages + (size_t)row * row_bytes))[c] = 0.0f;
        }
        if (threadIdx.x == 0) {
            for (int s = 0; s < kStages; s++) {
                bfptx::mbar_init(&full[s], 1);
                bfptx::mbar_init(&empty[s], W);
            }
            bfptx::fence_mbar_init();
            if (GATHER) *warps_done = 0u;
        }
    }
    __syncthreads();

    const int nchunks = (p.n + p.Mt - 1) / p.Mt;

    if (warp == W) {
        // =================== producer warp ==================================
        int s = 0;
        uint32_t ph = 1;                                  // fresh barrier: parity-1 wait passes
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int frame = tile / p.tiles_per_frame;
            const float *fsig = p.sig + (size_t)frame * p.n_mics_total * N;
            const float *fdiff = LERP ? p.diff + (size_t)frame * p.n * N : nullptr;
            for (int c = 0; c < nchunks; c++) {
                const int m0 = c * p.Mt;
                const int cnt = min(p.Mt, p.n - m0);
                bfptx::mbar_wait(&empty[s], ph);
                if (lane == 0)
                    bfptx::mbar_arrive_expect_tx(&full[s], (uint32_t)(cnt * arrays * N * 4));
                __syncwarp();
                unsigned char *sb = stages + (size_t)s * stage_bytes;
                for (int r = lane; r < cnt; r += 32) {
                    const int mic = p.mic_ids[m0 + r];
                    float *dst = (float *)(sb + (size_t)r * mic_bytes) + p.P;
                    bfptx::bulk_g2s(dst, fsig + (size_t)mic * N, N * 4, &full[s]);
                    if (LERP)
                        bfptx::bulk_g2s(dst + RS, fdiff + (size_t)(m0 + r) * N, N * 4, &full[s]);
                }
                if (++s == kStages) { s = 0; ph ^= 1; }
            }
        }
        return;
    }

    // ======================= consumer warps =================================
    // per-warp shared slot: two buffers (see slot_buf_bytes).  Entries (and lerp weights) of the next chunk
    // are copied global -> shared with cp.async while the current chunk is processed: no registers held.
    constexpr int BUF = slot_buf_bytes(LERP);
    unsigned char *slot = (unsigned char *)scratch_all + (size_t)warp * 2 * BUF;

    auto group_of = [&](int tile) {
        const int frame = tile / p.tiles_per_frame;
        return (tile - frame * p.tiles_per_frame) * W + warp;
    };
    auto prefetch_entries = [&](int tile, int c, unsigned char *dst) {
        if (tile >= p.total_tiles) return;
        const int g = group_of(tile);
        if (g >= p.groups) return;
        const int m = min(c * p.Mt + lane, p.n - 1);
        const size_t idx = (size_t)g * p.n + m;
        bfptx::cp_async16(dst + lane * 16, p.offs + idx);
        if (LERP) {
            const float *wp = p.wts + idx * kR;
            bfptx::cp_async16(dst + 512 + lane * 32, wp);
            bfptx::cp_async16(dst + 512 + lane * 32 + 16, wp + 4);
        }
    };
    uint32_t qc = 0;                                      // running chunk counter: buffer = qc & 1
    prefetch_entries(blockIdx.x, 0, slot);

    int s = 0;
    uint32_t ph = 0;
    bool peers_ready = false;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int frame = tile / p.tiles_per_frame;
        const int g = group_of(tile);
        const bool active = g < p.groups;

        float2 acc[kR][J / 2];
#pragma unroll
        for (int r = 0; r < kR; r++)
#pragma unroll
            for (int q = 0; q < J / 2; q++) acc[r][q] = make_float2(0.f, 0.f);

        for (int c = 0; c < nchunks; c++) {
            const int cnt = min(p.Mt, p.n - c * p.Mt);
            unsigned char *cur = slot + (qc & 1u) * BUF;
            const uint4 *ebuf = (const uint4 *)cur;
            const float *wbuf = (const float *)(cur + 512);
            bfptx::cp_async_wait_all();
            __syncwarp();
            if (c + 1 < nchunks) prefetch_entries(tile, c + 1, slot + ((qc + 1u) & 1u) * BUF);
            else prefetch_entries(tile + gridDim.x, 0, slot + ((qc + 1u) & 1u) * BUF);
            qc++;
            bfptx::mbar_wait(&full[s], ph);
            if (active) {
                const char *rowp = (const char *)(stages + (size_t)s * stage_bytes) + lane * 4;
                const uint2 *e2p = (const uint2 *)ebuf;          // first two words of entry i at e2p[2 * i]
                {
                    for (int mm = 0; mm < cnt; mm++, rowp += mic_bytes)
                        process_mic_q<J, LERP, PACK>(acc, e2p[2 * mm], ebuf + mm, rowp, (uint32_t)row_bytes, wbuf + mm * 8);
                }
            }
            __syncwarp();
            if (lane == 0) bfptx::mbar_arrive(&empty[s]);
            if (++s == kStages) { s = 0; ph ^= 1; }
        }

        if (!active) continue;

        // fused gather: the peers' buffers may be overwritten only after every rank published wait_seq
        if (GATHER && p.wait_seq > 0 && !peers_ready) {
            if (lane < p.world) {
                const volatile long long *f = p.flags_local + lane;
                long long spins = 0;
                while (*f < p.wait_seq) {
                    __nanosleep(100);
                    if (++spins > 100000000LL) { *p.timed_out = 1; break; }
                }
            }
            __threadfence_system();
            __syncwarp();
            peers_ready = true;
        }

        // ---- epilogue: out/n, square, sum over t, /N (pad_and_sum.c:122-131) ----
        float *img = p.img + (long)frame * p.img_fs + (long)(p.d_begin + g * kR - p.d_origin) * p.img_ds;
        const int valid = min(kR, p.d_count - g * kR);
        const long ds = p.img_ds;
        if (EXACT) {
            // the buffer of the chunk just consumed is free; the other one holds the next tile's first entries
            float *scratch = (float *)(slot + ((qc - 1u) & 1u) * BUF);
            float run = 0.0f;
#pragma unroll
            for (int q = 0; q < J; q++) {                 // 32 samples per round: t = 32 q + lane
#pragma unroll
                for (int r = 0; r < kR; r++) {
                    float x0 = (q & 1) ? acc[r][q >> 1].y : acc[r][q >> 1].x;
                    if (p.n_pow2) x0 = __fmul_rn(x0, p.inv_n);
                    else          x0 = __fdiv_rn(x0, p.fn);
                    scratch[r * kScratchStride + lane] = __fmul_rn(x0, x0);
                }
                __syncwarp();
                if (lane < kR) {
                    const float4 *s4 = (const float4 *)(scratch + lane * kScratchStride);
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const float4 v = s4[i];
                        run = __fadd_rn(run, v.x);
                        run = __fadd_rn(run, v.y);
                        run = __fadd_rn(run, v.z);
                        run = __fadd_rn(run, v.w);
                    }
                }
                __syncwarp();
            }
            if (lane < valid) {
                const float v = __fmul_rn(run, 1.0f / (float)N);
                img[lane * ds] = v;
                if (GATHER)
                    for (int q = 0; q < p.n_peers; q++) (p.peers[q] + (img - p.img))[lane * ds] = v;   // NVLink P2P store
            }
        } else {
            float tot[kR];
#pragma unroll
            for (int r = 0; r < kR; r++) {
                float t = 0.0f;
#pragma unroll
                for (int q = 0; q < J / 2; q++) {
                    float x0 = acc[r][q].x, x1 = acc[r][q].y;
                    if (p.n_pow2) { x0 = __fmul_rn(x0, p.inv_n); x1 = __fmul_rn(x1, p.inv_n); }
                    else          { x0 = __fdiv_rn(x0, p.fn);    x1 = __fdiv_rn(x1, p.fn); }
                    t = fmaf(x0, x0, t);
                    t = fmaf(x1, x1, t);
                }
#pragma unroll
                for (int sh = 16; sh > 0; sh >>= 1) t += __shfl_xor_sync(0xffffffffu, t, sh);
                tot[r] = t;
            }
            float mine = 0.0f;
#pragma unroll
            for (int r = 0; r < kR; r++)
                if (lane == r) mine = tot[r];
            if (lane < valid) {
                const float v = __fmul_rn(mine, 1.0f / (float)N);
                img[lane * ds] = v;
                if (GATHER)
                    for (int q = 0; q < p.n_peers; q++) (p.peers[q] + (img - p.img))[lane * ds] = v;
            }
            __syncwarp();
        }
    }

    // fused gather: publish "step complete" once every warp of every CTA has finished its stores
    if (GATHER && p.signal_seq > 0) {
        __threadfence_system();
        __syncwarp();
        if (lane == 0 && atomicAdd(warps_done, 1u) == (unsigned)W - 1u) {
            if (atomicAdd(p.done_counter, 1u) == gridDim.x - 1u) {
                *p.done_counter = 0u;
                __threadfence_system();
                for (int r = 0; r < p.world; r++) {
                    volatile long long *f = p.flags_all[r] + p.flag_rank;
                    *f = p.signal_seq;
                }
                __threadfence_system();
            }
        }
    }
}

// ---------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------
static int round_up(int v, int m) { return (v + m - 1) / m * m; }

template <int J>
static int launch_J(bool lerp, bool exact, const MimoParams &mp, int grid, size_t smem,
                    cudaStream_t st)
{
    auto go = [&](auto kern) -> int {
        BF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, (mp.W + 1) * 32, smem, st>>>(mp);
        BF_CHECK_LAUNCH();
        count_launch();
        return BF_OK;
    };
    static const bool pack = getenv("BF_MIMO_PACKED") ? atoi(getenv("BF_MIMO_PACKED")) != 0 : true;
    if (mp.n_peers > 0 || mp.wait_seq > 0 || mp.signal_seq > 0) {      // fused all-gather variant
        if (lerp) return exact ? go(das_mimo_kernel<J, true, true, true, true>) : go(das_mimo_kernel<J, true, false, true, true>);
        return exact ? go(das_mimo_kernel<J, false, true, true, true>) : go(das_mimo_kernel<J, false, false, true, true>);
    }
    if (pack) {
        if (lerp) return exact ? go(das_mimo_kernel<J, true, true, true, false>) : go(das_mimo_kernel<J, true, false, true, false>);
        return exact ? go(das_mimo_kernel<J, false, true, true, false>) : go(das_mimo_kernel<J, false, false, true, false>);
    }
    if (lerp) return exact ? go(das_mimo_kernel<J, true, true, false, false>) : go(das_mimo_kernel<J, true, false, false, false>);
    return exact ? go(das_mimo_kernel<J, false, true, false, false>) : go(das_mimo_kernel<J, false, false, false, false>);
}

int mimo_tiled(int algo, const float *d_sig, float *d_img, int frames, const int *d_mics, int n,
               int d_begin, int d_count, ImgLayout lay, cudaStream_t st)
{
    State &S = state();
    const int N = S.cfg.n_samples;
    const int D = S.cfg.max_res_x * S.cfg.max_res_y;
    const bool lerp = (algo == BF_ALGO_LERP);
    if (N != 64 && N != 128 && N != 256) {
        set_error(BF_ERR_CONFIG, "tiled kernel supports N_SAMPLES in {64,128,256}, got %d", N);
        return BF_ERR_CONFIG;
    }
    Tables &T = S.tab;
    const int *whole; const float *weight = nullptr; size_t count; int wmax; GroupTable *gt;
    if (algo == BF_ALGO_PAD)       { whole = T.pad_whole.as<int>(); count = T.pad_count; wmax = T.pad_max; gt = &T.g_pad; }
    else if (algo == BF_ALGO_LERP) { whole = T.lerp_whole.as<int>(); weight = T.lerp_weight.as<float>(); count = T.lerp_count; wmax = T.lerp_max; gt = &T.g_lerp; }
    else if (algo == -1)           { whole = T.trunc_whole.as<int>(); count = T.trunc_count; wmax = T.trunc_max; gt = &T.g_trunc; }
    else { set_error(BF_ERR_ARG, "mimo_tiled: bad algo %d", algo); return BF_ERR_ARG; }
    if (count < (size_t)D * n || whole == nullptr) {
        set_error(BF_ERR_NOT_LOADED, "coefficient table holds %zu entries, need D*n = %d*%d", count, D, n);
        return BF_ERR_NOT_LOADED;
    }
    // zero-pad width of the smem rows: covers the largest (clamped) delay in the table
    const int w_hi = lerp ? N - 1 : N;
    int wc = wmax < 0 ? 0 : (wmax > w_hi ? w_hi : wmax);
    const int P = round_up(wc + (lerp ? 1 : 0), 32);
    if (gt->n != n || gt->d_begin != d_begin || gt->d_count != d_count || gt->pad != P ||
        gt->n_samples != N) {
        int rc = build_groups(*gt, whole, weight, n, d_begin, d_count, P, N, lerp, st);
        if (rc) return rc;
    }
    const float *d_diff = nullptr;
    if (lerp) {
        int rc = S.d_diff.ensure((size_t)frames * n * N * sizeof(float));
        if (rc) return rc;
        diff_rows_kernel<<<dim3(n, frames), 128, 0, st>>>(d_sig, d_mics, S.d_diff.as<float>(), n, N,
                                                          S.cfg.n_microphones);
        BF_CHECK_LAUNCH();
        count_launch();
        d_diff = S.d_diff.as<float>();
    }

    MimoParams mp{};
    mp.sig = d_sig; mp.diff = d_diff; mp.img = d_img; mp.mic_ids = d_mics;
    mp.offs = gt->offs.as<uint4>(); mp.wts = gt->wts.as<float>();
    mp.n = n; mp.n_mics_total = S.cfg.n_microphones;
    mp.img_fs = lay.frame_stride; mp.img_ds = lay.dir_stride; mp.d_origin = lay.d_origin;
    mp.n_peers = lay.n_peers;
    mp.flags_local = lay.flags_local; mp.wait_seq = lay.wait_seq; mp.world = lay.world; mp.flag_rank = lay.flag_rank;
    mp.signal_seq = lay.signal_seq; mp.done_counter = lay.done_counter; mp.timed_out = lay.timed_out;
    for (int r = 0; r < lay.world && r < 8; r++) mp.flags_all[r] = lay.flags_all[r];
    for (int q = 0; q < lay.n_peers && q < 7; q++) mp.peers[q] = lay.peers[q];
    mp.d_begin = d_begin; mp.d_count = d_count; mp.frames = frames;
    mp.groups = gt->groups;
    mp.P = P;
    mp.fn = (float)n; mp.inv_n = 1.0f / (float)n; mp.n_pow2 = (n & (n - 1)) == 0;

    const long total_groups = (long)gt->groups * frames;
    int W = (int)((total_groups + S.sm_count - 1) / S.sm_count);
    const int wcap = max_warps(lerp);
    W = W < 1 ? 1 : (W > wcap ? wcap : W);
    if (W > 4) W = (round_up(W + 1, 4) - 1) > wcap ? wcap : (round_up(W + 1, 4) - 1);
    mp.W = W;
    mp.tiles_per_frame = (gt->groups + W - 1) / W;
    mp.total_tiles = mp.tiles_per_frame * frames;
    int grid = mp.total_tiles < S.sm_count ? mp.total_tiles : S.sm_count;

    // stage geometry: as many mic rows per stage as fit in the smem budget (<= 32: one entry per lane)
    const size_t row_bytes = (size_t)(P + N) * 4 * (lerp ? 2 : 1);
    const size_t scratch_bytes = (size_t)W * 2 * slot_buf_bytes(lerp);
    const size_t budget = 227 * 1024 - 256 - scratch_bytes - 1024;
    int Mt = 32;
    while (Mt > 1 && (size_t)Mt * row_bytes * kStages > budget) Mt >>= 1;
    if ((size_t)Mt * row_bytes * kStages > budget) {
        set_error(BF_ERR_CONFIG, "shared memory budget exceeded (row %zu bytes)", row_bytes);
        return BF_ERR_CONFIG;
    }
    if (Mt > n) Mt = n;
    mp.Mt = Mt;
    const size_t smem = 256 + (size_t)kStages * Mt * row_bytes + scratch_bytes;

    if (((uintptr_t)d_sig & 15) != 0) {
        set_error(BF_ERR_ARG, "signal buffer must be 16-byte aligned for bulk copies");
        return BF_ERR_ARG;
    }
    const bool exact = S.exact_sum != 0;
    switch (N) {
        case 64:  return launch_J<2>(lerp, exact, mp, grid, smem, st);
        case 128: return launch_J<4>(lerp, exact, mp, grid, smem, st);
        case 256: return launch_J<8>(lerp, exact, mp, grid, smem, st);
    }
    set_error(BF_ERR_CONFIG, "unsupported N_SAMPLES %d", N);
    return BF_ERR_CONFIG;
}

}  // namespace bf
