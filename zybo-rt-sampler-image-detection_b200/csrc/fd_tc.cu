// fd_tc.cu -- MVDR steering contraction on the 5th-generation tensor cores (tcgen05 + TMEM).
//
//   q_f(d) = || L_f^-1 a_f(d) ||^2 ,   P(d) = sum_f 1 / q_f(d)            (see fd_mvdr.cu)
//
// Real formulation of the complex product y = L^-1 a, a = c + j s:
//   [yr_i]   [ Lr_ij  -Li_ij ] [c_j]
//   [yi_i] = [ Li_ij   Lr_ij ] [s_j]       -> one real GEMM  Y[512] = Lblk[512 x 512] * [c;s][512]
// per (bin, direction).  MMA roles (tcgen05.mma D[M x N] += A[M x K] * B[N x K]^T, both K-major):
//   A operand  = the phasor tile: M = 128 directions (TMEM lanes), K = (cos_j, sin_j) pairs,
//                GENERATED on the fly by the CTA's threads (fp64 phase reduction + sincospif),
//                never read from memory
//   B operand  = Lblk rows: N = 256 rows per MMA (two N-tiles: rows of microphones i < 128 and
//                i >= 128), streamed from a pre-swizzled image with 1-D bulk TMA copies
//   D          = fp32 accumulators in TMEM: 128 lanes x 512 columns (all of TMEM)
// so a thread (= TMEM lane = direction) ends up owning all 512 Y values of its direction and
// q(d) is a private sum of squares: no cross-thread reduction.
//
// Generations living here (BF_MVDR_TC selects; 4 is the default):
//   1  mvdr_tc_steer_kernel   one CTA per (bin, 128 directions), generate -> copy -> MMA serially, kind::tf32
//   2  mvdr_tc_steer_kernel2  persistent, warp-specialised, double-buffered, kind::tf32 with a 3-pass tf32 split
//                             (x = hi + lo, hi = x with the low 13 mantissa bits cleared)  -- round 1
//   3  mvdr_tc_steer_kernel3  the same roles with kind::f16 and a two-term fp16 split, a converged MMA warp
//                             (descriptors in uniform registers) and fixed-point phases         -- round 2
//   4  mvdr_tc_steer_kernel3<QUARTER>  the same with N = 128 MMAs and quarter-granular triangular skipping
//                             (20 instead of 24 half-tile equivalents per unit)                 -- round 2, shipped
// "Issued" tensor flops are 3x the useful 8*M^2 per (bin, direction); the triangular structure of L^-1 lets
// N-tile 0 skip the second half of K (-25 %): issued = 2.25 x useful (generation 4: 20 / 32 quarters: 1.875 x).
//
// Shared-memory operand layout: K-major, 128-byte rows (32 tf32 = 16 microphones, or 64 halves = 32 microphones,
// per k-chunk), SWIZZLE_128B (16-byte chunk index XOR (row & 7)), 8-row groups of 1024 bytes
// (stride_byte_offset = 1024); one k-chunk = 4 MMA k-steps of 32 bytes (descriptor start + 32 B).
#include <math.h>
#include <cuda_fp16.h>

#include "bf_common.cuh"

namespace bf {

uint64_t fd_geometry_generation();          // fd_path.cu: bumped by every bf_fd_setup

static constexpr int kTcDirs = 128;        // directions per CTA (MMA M)
static constexpr int kTcKc = 32;           // K elements per chunk = 16 microphones (128-byte rows)
static constexpr int kTcMics = 256;        // this kernel is specialised for M = 256 microphones
static constexpr int kTcRows = 2 * kTcMics;                 // 512 Lblk rows
static constexpr int kTcChunks = 2 * kTcMics / kTcKc;       // 16 k-chunks
static constexpr size_t kTcPlaneB = (size_t)kTcRows * 128;  // bytes of one B plane per chunk (64 KiB)
static constexpr size_t kTcPlaneA = (size_t)kTcDirs * 128;  // bytes of one A plane per chunk (16 KiB)

__device__ __forceinline__ uint32_t swz128(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }

// ---- prep: L^-1 (float2 [M][M], lower triangular) -> pre-swizzled B-operand image --------------
// image[f][chunk][plane hi/lo][row n = 2i+part][32 k]  (k = 2*(j - 16*chunk) + {0: cos, 1: sin})
__global__ void mvdr_tc_prep_kernel(const float2 *__restrict__ linv, unsigned char *__restrict__ image)
{
    const int f = blockIdx.y, chunk = blockIdx.x;
    const float2 *L = linv + (size_t)f * kTcMics * kTcMics;
    unsigned char *img = image + ((size_t)f * kTcChunks + chunk) * 2 * kTcPlaneB;
    for (int e = threadIdx.x; e < kTcRows * kTcKc; e += blockDim.x) {
        const int n = e / kTcKc, k = e - n * kTcKc;
        const int i = n >> 1, part = n & 1;
        const int j = chunk * (kTcKc / 2) + (k >> 1), sc = k & 1;
        float v = 0.0f;
        if (j <= i) {
            const float2 l = L[(size_t)i * kTcMics + j];
            // part 0 (yr): cos -> Lr, sin -> -Li ; part 1 (yi): cos -> Li, sin -> Lr
            v = part == 0 ? (sc == 0 ? l.x : -l.y) : (sc == 0 ? l.y : l.x);
        }
        const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        const float lo = v - hi;
        const uint32_t off = swz128((uint32_t)n * 128u + (uint32_t)k * 4u);
        *(float *)(img + off) = hi;
        *(float *)(img + kTcPlaneB + off) = lo;
    }
}

// ---- tcgen05 helpers ------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr)
{
    // start address >> 4 | LBO (unused for swizzled K-major) | SBO = 1024 B | version 1 | SWIZZLE_128B
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(bfptx::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 columns (one fp32 per lane per column) -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// ---- the kernel: grid (direction tiles, bins), 128 threads ----------------------------------------
__global__ void __launch_bounds__(128, 1) mvdr_tc_steer_kernel(const unsigned char *__restrict__ image,
                                                               const double *__restrict__ u, int F, int lo,
                                                               double bin_hz, double inv_c, int D,
                                                               float *__restrict__ qout)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    // [0, 32K): A planes (hi, lo);  [32K, 160K): B planes (hi, lo);  then barriers + tmem address
    unsigned char *sA = smem;
    unsigned char *sB = smem + 2 * kTcPlaneA;
    uint64_t *bar_tma = (uint64_t *)(sB + 2 * kTcPlaneB);
    uint64_t *bar_mma = bar_tma + 1;
    uint32_t *tmem_slot = (uint32_t *)(bar_mma + 1);

    const int t = threadIdx.x, warp = t >> 5;
    const int f = blockIdx.y;
    const int d = blockIdx.x * kTcDirs + t;
    const double *ud = u + (size_t)(d < D ? d : D - 1) * kTcMics;
    const double turns_per_u = (double)(lo + f) * bin_hz * inv_c;

    if (t == 0) {
        bfptx::mbar_init(bar_tma, 1);
        bfptx::mbar_init(bar_mma, 1);
        bfptx::fence_mbar_init();
    }
    if (warp == 0) {      // allocate all 512 TMEM columns (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;"
                     ::"r"(bfptx::smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // instruction descriptor: D fp32, A/B tf32, both K-major, M = 128, N = 256
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    const unsigned char *img = image + (size_t)f * kTcChunks * 2 * kTcPlaneB;

    uint32_t phase = 0;
    for (int chunk = 0; chunk < kTcChunks; chunk++) {
        // N-tile 0 (microphone rows i < 128) only sees microphones j < 128: chunks 0..7
        const bool tile0 = chunk < kTcChunks / 2;
        if (chunk > 0) {                       // previous MMAs have finished reading sA / sB
            bfptx::mbar_wait(bar_mma, phase ^ 1);
            tc_fence_after();
        }
        if (t == 0) {
            const unsigned char *src = img + (size_t)chunk * 2 * kTcPlaneB;
            const uint32_t skip = tile0 ? 0u : (uint32_t)(kTcPlaneB / 2);        // rows 256..511 only
            const uint32_t bytes = (uint32_t)kTcPlaneB - skip;
            bfptx::mbar_arrive_expect_tx(bar_tma, 2 * bytes);
            bfptx::bulk_g2s(sB + skip, src + skip, bytes, bar_tma);
            bfptx::bulk_g2s(sB + kTcPlaneB + skip, src + kTcPlaneB + skip, bytes, bar_tma);
        }
        // generate the A chunk: row t = direction, 16 microphones -> (cos, sin) pairs, hi / lo planes
        {
            const int j0 = chunk * (kTcKc / 2);
#pragma unroll 4
            for (int c = 0; c < 8; c++) {                 // 16-byte chunk c holds microphones j0+2c, j0+2c+1
                float4 hi, lo;
                float sn, cs;
                double turns = turns_per_u * ud[j0 + 2 * c];
                sincospif(-2.0f * (float)(turns - rint(turns)), &sn, &cs);
                hi.x = __uint_as_float(__float_as_uint(cs) & 0xffffe000u); lo.x = cs - hi.x;
                hi.y = __uint_as_float(__float_as_uint(sn) & 0xffffe000u); lo.y = sn - hi.y;
                turns = turns_per_u * ud[j0 + 2 * c + 1];
                sincospif(-2.0f * (float)(turns - rint(turns)), &sn, &cs);
                hi.z = __uint_as_float(__float_as_uint(cs) & 0xffffe000u); lo.z = cs - hi.z;
                hi.w = __uint_as_float(__float_as_uint(sn) & 0xffffe000u); lo.w = sn - hi.w;
                const uint32_t off = swz128((uint32_t)t * 128u + (uint32_t)c * 16u);
                *(float4 *)(sA + off) = hi;
                *(float4 *)(sA + kTcPlaneA + off) = lo;
            }
        }
        bfptx::fence_proxy_async();            // generic-proxy smem writes -> visible to the async proxy
        __syncthreads();
        if (t == 0) {
            bfptx::mbar_wait(bar_tma, phase);
            tc_fence_after();
            const uint32_t a_hi = bfptx::smem_u32(sA), a_lo = a_hi + (uint32_t)kTcPlaneA;
            const uint32_t b_hi = bfptx::smem_u32(sB), b_lo = b_hi + (uint32_t)kTcPlaneB;
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
                if (nt == 0 && !tile0) continue;
                const uint32_t brow = (uint32_t)nt * 256u * 128u;          // 256 rows per N-tile
                const uint32_t dcol = tmem + (uint32_t)nt * 256u;
#pragma unroll
                for (int ks = 0; ks < 4; ks++) {
                    const uint32_t ko = (uint32_t)ks * 32u;                // 8 tf32 = 32 bytes per k-step
                    const uint32_t acc = (chunk > 0 || ks > 0) ? 1u : 0u;
                    umma_tf32(dcol, umma_desc_sw128(a_hi + ko), umma_desc_sw128(b_hi + brow + ko), idesc, acc);
                    umma_tf32(dcol, umma_desc_sw128(a_hi + ko), umma_desc_sw128(b_lo + brow + ko), idesc, 1u);
                    umma_tf32(dcol, umma_desc_sw128(a_lo + ko), umma_desc_sw128(b_hi + brow + ko), idesc, 1u);
                }
            }
            umma_commit(bar_mma);              // implies tcgen05.fence::before_thread_sync
        }
        phase ^= 1;
    }
    // ---- epilogue: thread t = TMEM lane t = direction: q = sum over the 512 columns of Y^2 --------
    bfptx::mbar_wait(bar_mma, phase ^ 1);
    tc_fence_after();
    float q = 0.0f;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < 512; c0 += 32) {
        float v[32];
        tmem_ld32(lane_base + (uint32_t)c0, v);
#pragma unroll
        for (int i = 0; i < 32; i++) q = fmaf(v[i], v[i], q);
    }
    if (d < D) qout[(size_t)f * D + d] = q;
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// ---- version 2: persistent, warp-specialised, double-buffered -------------------------------------
// Roles (448 threads): warps 0-3 epilogue (TMEM lane quarters 0-3), warp 4 MMA issuer, warp 5 bulk-TMA
// producer, warps 6-13 phasor generators (256 threads: 2 per direction).  Rings: A chunk buffers x2
// (hi/lo planes, 32 KiB each), B slots x2 (one N-tile of one k-chunk, hi/lo planes, 64 KiB each).
// N-tile 0 (microphone rows i < 128) finishes after chunk 7 and is read out of TMEM by the epilogue
// warps while N-tile 1 keeps the tensor pipe busy for chunks 8..15.
static constexpr int kV2GenWarps = 16;
static constexpr int kV2Threads = (6 + kV2GenWarps) * 32;       // 704
// k-chunks are visited in the order 0,8,1,9,...,7,15: chunks 8..15 only feed N-tile 1 (half the MMA
// work of chunks 0..7), interleaving them evens out the MMA time per generated A chunk.
__device__ __forceinline__ int v2_chunk(int i) { return (i & 1) ? kTcChunks / 2 + (i >> 1) : (i >> 1); }

// phase table for the generators: turns per bin index, reduced: phi = tau - rint(tau),
// tau = u * bin_hz / c, split as hi (multiple of 2^-12, so bin*hi is exact in fp32) + lo (fp32).
__global__ void mvdr_tc_phi_kernel(const double *__restrict__ u, size_t count, double bin_hz_over_c,
                                   float2 *__restrict__ phi)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double tau = u[i] * bin_hz_over_c;
    const double ph = tau - rint(tau);
    const double hi = rint(ph * 4096.0) * (1.0 / 4096.0);
    phi[i] = make_float2((float)hi, (float)(ph - hi));
}
static constexpr size_t kV2SlotB = (size_t)256 * 128 * 2;      // one N-tile, hi + lo planes: 64 KiB
static constexpr size_t kV2BufA = 2 * kTcPlaneA;               // hi + lo planes: 32 KiB

__global__ void __launch_bounds__(kV2Threads, 1) mvdr_tc_steer_kernel2(const unsigned char *__restrict__ image,
                                                                       const float2 *__restrict__ phi, int F, int lo,
                                                                       int D, int tiles, float *__restrict__ qout)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sA = smem;                                  // 2 x 32 KiB
    unsigned char *sB = smem + 2 * kV2BufA;                    // 2 x 64 KiB
    uint64_t *bars = (uint64_t *)(sB + 2 * kV2SlotB);
    uint64_t *a_full = bars, *a_empty = bars + 2, *b_full = bars + 4, *b_empty = bars + 6;
    uint64_t *t0_done = bars + 8, *acc_done = bars + 9, *acc_free = bars + 10;
    uint32_t *tmem_slot = (uint32_t *)(bars + 11);

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int units = F * tiles;

    if (t == 0) {
        for (int i = 0; i < 2; i++) {
            bfptx::mbar_init(&a_full[i], kV2GenWarps);   // one arrive per generator warp
            bfptx::mbar_init(&a_empty[i], 1);      // tcgen05.commit
            bfptx::mbar_init(&b_full[i], 1);       // expect_tx arrive
            bfptx::mbar_init(&b_empty[i], 1);      // tcgen05.commit
        }
        bfptx::mbar_init(t0_done, 1);
        bfptx::mbar_init(acc_done, 1);
        bfptx::mbar_init(acc_free, 4);             // one arrive per epilogue warp
        bfptx::fence_mbar_init();
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;"
                     ::"r"(bfptx::smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 5) {
        // ================= bulk-TMA producer: B slots ==================================================
        if (lane == 0) {
            uint32_t k = 0;                                      // global item counter
            for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
                const int f = unit / tiles;
                const unsigned char *img = image + (size_t)f * kTcChunks * 2 * kTcPlaneB;
                for (int ci = 0; ci < kTcChunks; ci++) {
                    const int chunk = v2_chunk(ci);
                    for (int nt = (chunk < kTcChunks / 2 ? 0 : 1); nt < 2; nt++, k++) {
                        const uint32_t slot = k & 1, ph = (k >> 1) & 1;
                        bfptx::mbar_wait(&b_empty[slot], ph ^ 1);
                        unsigned char *dst = sB + slot * kV2SlotB;
                        const unsigned char *src = img + (size_t)chunk * 2 * kTcPlaneB + (size_t)nt * 256 * 128;
                        bfptx::mbar_arrive_expect_tx(&b_full[slot], (uint32_t)kV2SlotB);
                        bfptx::bulk_g2s(dst, src, 256 * 128, &b_full[slot]);                         // hi rows
                        bfptx::bulk_g2s(dst + 256 * 128, src + kTcPlaneB, 256 * 128, &b_full[slot]);  // lo rows
                    }
                }
            }
        }
    } else if (warp == 4) {
        // ================= MMA issuer =====================================================================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
            uint32_t k = 0, g = 0, w = 0;                        // item, chunk, work-unit counters
            for (int unit = blockIdx.x; unit < units; unit += gridDim.x, w++) {
                bfptx::mbar_wait(acc_free, (w & 1) ^ 1);          // epilogue has drained TMEM
                tc_fence_after();
                for (int ci = 0; ci < kTcChunks; ci++, g++) {
                    const int chunk = v2_chunk(ci);
                    const uint32_t ab = g & 1, aph = (g >> 1) & 1;
                    bfptx::mbar_wait(&a_full[ab], aph);
                    const uint32_t a_hi = bfptx::smem_u32(sA + ab * kV2BufA), a_lo = a_hi + (uint32_t)kTcPlaneA;
                    for (int nt = (chunk < kTcChunks / 2 ? 0 : 1); nt < 2; nt++, k++) {
                        const uint32_t slot = k & 1, ph = (k >> 1) & 1;
                        bfptx::mbar_wait(&b_full[slot], ph);
                        tc_fence_after();
                        const uint32_t b_hi = bfptx::smem_u32(sB + slot * kV2SlotB), b_lo = b_hi + 256u * 128u;
                        const uint32_t dcol = tmem + (uint32_t)nt * 256u;
#pragma unroll
                        for (int ks = 0; ks < 4; ks++) {
                            const uint32_t ko = (uint32_t)ks * 32u;
                            const uint32_t acc = (ci > 0 || ks > 0) ? 1u : 0u;
                            umma_tf32(dcol, umma_desc_sw128(a_hi + ko), umma_desc_sw128(b_hi + ko), idesc, acc);
                            umma_tf32(dcol, umma_desc_sw128(a_hi + ko), umma_desc_sw128(b_lo + ko), idesc, 1u);
                            umma_tf32(dcol, umma_desc_sw128(a_lo + ko), umma_desc_sw128(b_hi + ko), idesc, 1u);
                        }
                        umma_commit(&b_empty[slot]);
                        if (nt == 0 && chunk == kTcChunks / 2 - 1) umma_commit(t0_done);
                    }
                    umma_commit(&a_empty[ab]);
                }
                umma_commit(acc_done);
            }
        }
    } else if (warp >= 6) {
        // ================= phasor generators: A chunk buffers ==========================================
        const int gt = t - 6 * 32;                               // 0..511
        const int row = gt & 127, quarter = gt >> 7;             // direction row, which 4 of the 16 microphones
        uint32_t g = 0;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
            const int f = unit / tiles, tile = unit - f * tiles;
            const int d = tile * kTcDirs + row;
            const float4 *pr = (const float4 *)(phi + (size_t)(d < D ? d : D - 1) * kTcMics + quarter * 4);
            const float bin = (float)(lo + f);
            float4 p0 = __ldg(pr + v2_chunk(0) * 8), p1 = __ldg(pr + v2_chunk(0) * 8 + 1);
            for (int ci = 0; ci < kTcChunks; ci++, g++) {
                const uint32_t ab = g & 1, aph = (g >> 1) & 1;
                const float4 c0 = p0, c1 = p1;
                if (ci + 1 < kTcChunks) {                         // prefetch the next chunk's phases
                    p0 = __ldg(pr + v2_chunk(ci + 1) * 8);
                    p1 = __ldg(pr + v2_chunk(ci + 1) * 8 + 1);
                }
                const float ph_hi[4] = {c0.x, c0.z, c1.x, c1.z}, ph_lo[4] = {c0.y, c0.w, c1.y, c1.w};
                float4 hi[2], lo4[2];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    float t1 = __fmul_rn(bin, ph_hi[i]);          // exact: bin < 2^10, ph_hi multiple of 2^-12
                    t1 = __fsub_rn(t1, rintf(t1));
                    const float fr = __fmaf_rn(bin, ph_lo[i], t1);
                    float sn, cs;
                    sincospif(-2.0f * fr, &sn, &cs);
                    const float ch = __uint_as_float(__float_as_uint(cs) & 0xffffe000u);
                    const float sh = __uint_as_float(__float_as_uint(sn) & 0xffffe000u);
                    if (i & 1) { hi[i >> 1].z = ch; hi[i >> 1].w = sh; lo4[i >> 1].z = cs - ch; lo4[i >> 1].w = sn - sh; }
                    else       { hi[i >> 1].x = ch; hi[i >> 1].y = sh; lo4[i >> 1].x = cs - ch; lo4[i >> 1].y = sn - sh; }
                }
                bfptx::mbar_wait(&a_empty[ab], aph ^ 1);
                unsigned char *dst = sA + ab * kV2BufA;
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    const uint32_t off = swz128((uint32_t)row * 128u + (uint32_t)(quarter * 2 + c) * 16u);
                    *(float4 *)(dst + off) = hi[c];
                    *(float4 *)(dst + kTcPlaneA + off) = lo4[c];
                }
                bfptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) bfptx::mbar_arrive(&a_full[ab]);
            }
        }
    } else {
        // ================= epilogue warps 0-3: q(d) = sum over 512 columns of Y^2 ====================
        uint32_t w = 0;
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x, w++) {
            const int f = unit / tiles, tile = unit - f * tiles;
            const int d = tile * kTcDirs + warp * 32 + lane;
            float q = 0.0f;
            bfptx::mbar_wait(t0_done, w & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < 256; c0 += 32) {
                float v[32];
                tmem_ld32(lane_base + (uint32_t)c0, v);
#pragma unroll
                for (int i = 0; i < 32; i++) q = fmaf(v[i], v[i], q);
            }
            bfptx::mbar_wait(acc_done, w & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 256; c0 < 512; c0 += 32) {
                float v[32];
                tmem_ld32(lane_base + (uint32_t)c0, v);
#pragma unroll
                for (int i = 0; i < 32; i++) q = fmaf(v[i], v[i], q);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) bfptx::mbar_arrive(acc_free);
            if (d < D) qout[(size_t)f * D + d] = q;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// ---- version 3: kind::f16 with a two-term fp16 split (round 2) -------------------------------------
// Same roles, rings and TMEM layout as version 2, but the operands are fp16 (hi + lo, 11 + 11 significant
// bits) and the MMAs are kind::f16, which run at twice the tf32 rate: three passes hi*hi + hi*lo + lo*hi cost
// 1.5 bf16-equivalent passes instead of 3 tf32 ones (= 6).  A 128-byte operand row now holds 64 halves = 32
// microphones, so a map needs 8 k-chunks instead of 16 and half the MMA instructions.
// Range: fp16 has 5 exponent bits, so both operands are pre-scaled by exact powers of two -- the phasors by
// 2^8 (|hi| <= 256, lo stays a normal number down to |x| = 2^-10) and the L^-1 of every bin so that its largest
// entry lies in [2^7, 2^8) (mvdr_tc_rowscale_kernel) -- and the epilogue multiplies q by binscale[f]^2 =
// 2^-2(e_f + 8) (exact).  Absolute representation error per operand element: <= 2^-25 of the bin's largest
// entry (phasors: of 1).
static constexpr int kV3Kc = 64;                                 // halves per 128-byte row = 32 microphones
static constexpr int kV3Chunks = 2 * kTcMics / kV3Kc;            // 8
static constexpr int kV3GenWarps = 16;
static constexpr int kV3Threads = (6 + kV3GenWarps) * 32;        // 704
__device__ __forceinline__ int v3_chunk(int i) { return (i & 1) ? kV3Chunks / 2 + (i >> 1) : (i >> 1); }
// QUARTER variant (BF_MVDR_TC=4): N = 128 MMAs, the 512 rows of L^-1 in four quarters of 64 microphones; k-chunk c
// (microphones 32c .. 32c+31) only feeds quarters t >= c / 2 (L^-1 is lower triangular): 4,4,3,3,2,2,1,1 = 20 of 32
// (chunk, quarter) items instead of the 24 half-tile equivalents of the N = 256 version (-17 % MMA work).  Chunks are
// visited 0,7,1,6,2,5,3,4 so that every pair of chunks carries 5 items.
__device__ __forceinline__ int v3q_chunk(int i) { return (i & 1) ? kV3Chunks - 1 - (i >> 1) : (i >> 1); }
template <bool QUARTER> __device__ __forceinline__ int v3_order(int i) { return QUARTER ? v3q_chunk(i) : v3_chunk(i); }

// per bin: exponent e with max_ij(|Lr|,|Li|) * 2^e in [2^7, 2^8); binscale = 2^-(e+8) undoes it and the 2^8 of the
// phasors.  One scale per bin (not per row) keeps the epilogue free of per-column loads; rows whose entries
// are up to 2^6 below the bin's largest still keep a normal-range lo part, and the absolute error of
// any element stays <= 2^-25 of the bin's largest entry (cond(R) <= M / loading bounds the spread of the rows).
__global__ void mvdr_tc_rowscale_kernel(const float2 *__restrict__ linv, int *__restrict__ expo,
                                        float *__restrict__ binscale, int f0)
{
    __shared__ float red[kTcMics];
    const int f = f0 + blockIdx.x, i = threadIdx.x;              // 256 threads, one per row
    const float2 *row = linv + ((size_t)f * kTcMics + i) * kTcMics;
    float mx = 0.0f;
    for (int j = 0; j <= i; j++) { const float2 l = row[j]; mx = fmaxf(mx, fmaxf(fabsf(l.x), fabsf(l.y))); }
    red[i] = mx;
    __syncthreads();
    for (int s = kTcMics / 2; s > 0; s >>= 1) {
        if (i < s) red[i] = fmaxf(red[i], red[i + s]);
        __syncthreads();
    }
    if (i == 0) {
        mx = red[0];
        int e = 0;
        if (mx > 0.0f && isfinite(mx)) { int ex; frexpf(mx, &ex); e = 8 - ex; }      // mx = m * 2^ex, m in [0.5, 1)
        expo[f] = e;
        binscale[f] = ldexpf(1.0f, -(e + 8));
    }
}

// image3[f][chunk][plane hi/lo][row n = 2i+part][64 halves]   (k = 2*(j - 32*chunk) + {0: cos, 1: sin})
__global__ void mvdr_tc_prep3_kernel(const float2 *__restrict__ linv, const int *__restrict__ expo,
                                     unsigned char *__restrict__ image, int f0)
{
    const int f = f0 + blockIdx.y, chunk = blockIdx.x;
    const float2 *L = linv + (size_t)f * kTcMics * kTcMics;
    unsigned char *img = image + ((size_t)f * kV3Chunks + chunk) * 2 * kTcPlaneB;
    for (int e = threadIdx.x; e < kTcRows * (kV3Kc / 2); e += blockDim.x) {       // one (row, microphone) pair per step
        const int n = e / (kV3Kc / 2), jj = e - n * (kV3Kc / 2);
        const int i = n >> 1, part = n & 1;
        const int j = chunk * (kV3Kc / 2) + jj;
        float vc = 0.0f, vs = 0.0f;                                            // multiplies cos_j / sin_j
        if (j <= i) {
            const float2 l = L[(size_t)i * kTcMics + j];
            const int ex = expo[f];
            // part 0 (yr): cos -> Lr, sin -> -Li ; part 1 (yi): cos -> Li, sin -> Lr
            vc = ldexpf(part == 0 ? l.x : l.y, ex);
            vs = ldexpf(part == 0 ? -l.y : l.x, ex);
        }
        const __half2 hi = __floats2half2_rn(vc, vs);
        const float2 hf = __half22float2(hi);
        const __half2 lo = __floats2half2_rn(vc - hf.x, vs - hf.y);
        const uint32_t off = swz128((uint32_t)n * 128u + (uint32_t)jj * 4u);
        *(__half2 *)(img + off) = hi;
        *(__half2 *)(img + kTcPlaneB + off) = lo;
    }
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Phase table of version 3: tau(d, m) = u[d][m] * bin_hz / c in TURNS per bin index, reduced modulo 1 and stored
// as 32-bit fixed point.  The phase of bin b is then the low 32 bits of b * fix -- an exact integer multiply that
// wraps modulo one turn by itself -- with 2^-33 turns of quantisation per bin index (6e-8 turns at bin 512).
// Layout [tile][chunk][i][quarter][row]: the 32 lanes of a generator warp (32 consecutive rows = directions,
// one quarter) read 128 consecutive bytes per load.
__global__ void mvdr_tc_phifix_kernel(const double *__restrict__ u, int D, double bin_hz_over_c, int tiles,
                                      uint32_t *__restrict__ fix)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)tiles * kTcMics * kTcDirs;
    if (idx >= total) return;
    const int row = (int)(idx % kTcDirs);
    const int quarter = (int)((idx / kTcDirs) % 4);
    const int i = (int)((idx / (kTcDirs * 4)) % 8);
    const int chunk = (int)((idx / (kTcDirs * 32)) % 8);
    const int tile = (int)(idx / ((size_t)kTcDirs * kTcMics));
    int d = tile * kTcDirs + row;
    if (d >= D) d = D - 1;
    const int m = chunk * 32 + quarter * 8 + i;
    const double tau = u[(size_t)d * kTcMics + m] * bin_hz_over_c;
    const double fr = tau - floor(tau);                              // [0, 1)
    fix[idx] = (uint32_t)(unsigned long long)llrint(fr * 4294967296.0);   // 2^32 wraps to 0
}

// (cos, -sin) of 2 pi x / 2^32 for a 32-bit fixed-point phase x (turns): the phasor exp(-j 2 pi x / 2^32).
// Quadrant by integer arithmetic, then on [-pi/4, pi/4] either fp32 polynomials (truncation error < 2e-9,
// FAST = false) or the SFU's MUFU.SIN / MUFU.COS (FAST = true: 2 special-function ops instead of 11 FMAs; on this
// reduced range their absolute error is ~2^-21.4, an order of magnitude below the map tolerance -- measured in
// tests/test_gpu_c4_size.py).
template <bool FAST>
__device__ __forceinline__ void phasor_fix(uint32_t x, float &cs, float &sn)
{
    const uint32_t q = (x + 0x20000000u) >> 30;                      // nearest quarter turn, 0..3 (4 wraps to 0)
    const int32_t r = (int32_t)(x - (q << 30));                      // [-2^29, 2^29)
    const float z = (float)r * 1.4629180792671596e-09f;              // 2 pi / 2^32
    float s, c;
    if (FAST) {
        s = __sinf(z);
        c = __cosf(z);
    } else {
        const float z2 = z * z;
        s = fmaf(z2, 2.7557319e-06f, -1.9841270e-04f);               // z - z^3/3! + z^5/5! - z^7/7! + z^9/9!
        s = fmaf(s, z2, 8.3333333e-03f);
        s = fmaf(s, z2, -1.6666667e-01f);
        s = fmaf(s * z2, z, z);
        c = fmaf(z2, -2.7557319e-07f, 2.4801587e-05f);               // 1 - z^2/2! + z^4/4! - z^6/6! + z^8/8! - z^10/10!
        c = fmaf(c, z2, -1.3888889e-03f);
        c = fmaf(c, z2, 4.1666667e-02f);
        c = fmaf(c, z2, -0.5f);
        c = fmaf(c, z2, 1.0f);
    }
    // angle = q * pi/2 + z:  cos, sin = (c, s), (-s, c), (-c, -s), (s, -c)
    const float a = (q & 1u) ? s : c, b = (q & 1u) ? c : s;
    const float co = ((q + 1u) & 2u) ? -a : a;
    const float si = (q & 2u) ? -b : b;
    cs = co;
    sn = -si;
}

// Issued by ALL lanes of the MMA warp: elect.sync picks one lane and only that lane's instruction is
// predicated on.  Keeping the warp converged (no `if (lane == 0)` region) lets the compiler hold the
// descriptors in uniform registers; inside a one-lane branch every UTCHMMA was preceded by R2UR moves and
// cost the issuing thread ~170 cycles (ncu, round 2), more than the MMA itself takes.
__device__ __forceinline__ void umma_f16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t *bar)
{
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(bfptx::smem_u32(bar)) : "memory");
}

template <bool QUARTER>
__global__ void __launch_bounds__(kV3Threads, 1) mvdr_tc_steer_kernel3(const unsigned char *__restrict__ image,
                                                                       const uint32_t *__restrict__ phifix,
                                                                       const float *__restrict__ colscale, int F, int lo,
                                                                       int D, int tiles, float *__restrict__ qout, int dbg)
{
    // dbg (BF_MVDR_DBG): timing experiments, results are wrong -- bit 0 generators skip the sincos, bit 1 the
    // producer skips the B copies; bit 2 (results valid): polynomial phasors instead of MUFU.SIN / MUFU.COS
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sA = smem;                                  // 2 x 32 KiB (hi + lo planes of 128 rows)
    unsigned char *sB = smem + 2 * kV2BufA;                    // 2 x 64 KiB (hi + lo planes of 256 rows)
    constexpr int NB = QUARTER ? 4 : 2;                       // B ring slots (32 KiB quarters or 64 KiB halves)
    constexpr uint32_t kSlotB = QUARTER ? (uint32_t)kV2SlotB / 2 : (uint32_t)kV2SlotB;
    constexpr uint32_t kRowsB = QUARTER ? 128u : 256u;        // rows per slot = MMA N
    uint64_t *bars = (uint64_t *)(sB + 2 * kV2SlotB);
    uint64_t *a_full = bars, *a_empty = bars + 2, *b_full = bars + 4, *b_empty = bars + 8;
    uint64_t *t0_done = bars + 12, *acc_done = bars + 13, *acc_free = bars + 14;
    uint32_t *tmem_slot = (uint32_t *)(bars + 15);

    const int t = threadIdx.x, lane = t & 31;
    // warp index as a warp-uniform value: the role dispatch below is then a uniform branch and the MMA warp's
    // addresses / descriptors can live in uniform registers
    const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);
    const int units = F * tiles;

    if (t == 0) {
        for (int i = 0; i < 2; i++) {
            bfptx::mbar_init(&a_full[i], kV3GenWarps);
            bfptx::mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < NB; i++) {
            bfptx::mbar_init(&b_full[i], 1);
            bfptx::mbar_init(&b_empty[i], 1);
        }
        bfptx::mbar_init(t0_done, 1);
        bfptx::mbar_init(acc_done, 1);
        bfptx::mbar_init(acc_free, 4);
        bfptx::fence_mbar_init();
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;"
                     ::"r"(bfptx::smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 5) {
        // ================= bulk-TMA producer: B slots ==================================================
        if (lane == 0) {
            uint32_t k = 0;
            for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
                const int f = unit / tiles;
                const unsigned char *img = image + (size_t)f * kV3Chunks * 2 * kTcPlaneB;
                for (int ci = 0; ci < kV3Chunks; ci++) {
                    const int chunk = v3_order<QUARTER>(ci);
                    const int nt0 = QUARTER ? (chunk >> 1) : (chunk < kV3Chunks / 2 ? 0 : 1);
                    for (int nt = nt0; nt < (QUARTER ? 4 : 2); nt++, k++) {
                        const uint32_t slot = k % NB, ph = (k / NB) & 1;
                        bfptx::mbar_wait(&b_empty[slot], ph ^ 1);
                        unsigned char *dst = sB + slot * kSlotB;
                        const unsigned char *src = img + (size_t)chunk * 2 * kTcPlaneB + (size_t)nt * kRowsB * 128;
                        if (dbg & 2) { bfptx::mbar_arrive(&b_full[slot]); continue; }
                        bfptx::mbar_arrive_expect_tx(&b_full[slot], kSlotB);
                        bfptx::bulk_g2s(dst, src, kRowsB * 128, &b_full[slot]);
                        bfptx::bulk_g2s(dst + kRowsB * 128, src + kTcPlaneB, kRowsB * 128, &b_full[slot]);
                    }
                }
            }
        }
    } else if (warp == 4) {
        // ================= MMA issuer =====================================================================
        {
            // D fp32, A / B fp16, both K-major, M = 128, N = 256
            const uint32_t idesc = (1u << 4) | ((kRowsB >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t a_base = bfptx::smem_u32(sA), b_base = bfptx::smem_u32(sB);
            uint32_t k = 0, g = 0, w = 0;
            for (int unit = blockIdx.x; unit < units; unit += gridDim.x, w++) {
                bfptx::mbar_wait(acc_free, (w & 1) ^ 1);
                tc_fence_after();
                uint32_t accm = 0u;                              // bit t set: N-tile t already holds a partial sum
                for (int ci = 0; ci < kV3Chunks; ci++, g++) {
                    const int chunk = v3_order<QUARTER>(ci);
                    const uint32_t ab = g & 1, aph = (g >> 1) & 1;
                    bfptx::mbar_wait(&a_full[ab], aph);
                    tc_fence_after();
                    const uint64_t da_hi = umma_desc_sw128(a_base + ab * (uint32_t)kV2BufA);
                    const uint64_t da_lo = umma_desc_sw128(a_base + ab * (uint32_t)kV2BufA + (uint32_t)kTcPlaneA);
                    const int nt0 = QUARTER ? (chunk >> 1) : (chunk < kV3Chunks / 2 ? 0 : 1);
                    for (int nt = nt0; nt < (QUARTER ? 4 : 2); nt++, k++) {
                        const uint32_t slot = k % NB, ph = (k / NB) & 1;
                        bfptx::mbar_wait(&b_full[slot], ph);
                        tc_fence_after();
                        const uint64_t db_hi = umma_desc_sw128(b_base + slot * kSlotB);
                        const uint64_t db_lo = umma_desc_sw128(b_base + slot * kSlotB + kRowsB * 128u);
                        const uint32_t dcol = tmem + (uint32_t)nt * kRowsB;
                        const uint32_t acc = (accm >> nt) & 1u;
#pragma unroll
                        for (int ks = 0; ks < 4; ks++) {
                            const uint64_t ko = (uint64_t)(ks * 2);          // 16 halves = 32 bytes = 2 descriptor units
                            umma_f16_elect(dcol, da_hi + ko, db_hi + ko, idesc, ks == 0 ? acc : 1u);
                            umma_f16_elect(dcol, da_hi + ko, db_lo + ko, idesc, 1u);
                            umma_f16_elect(dcol, da_lo + ko, db_hi + ko, idesc, 1u);
                        }
                        accm |= 1u << nt;
                        umma_commit_elect(&b_empty[slot]);
                        // columns 0..255 are final after the last item of k-chunk 3 that lands in them
                        if (chunk == kV3Chunks / 2 - 1 && nt == (QUARTER ? 1 : 0)) umma_commit_elect(t0_done);
                    }
                    umma_commit_elect(&a_empty[ab]);
                }
                umma_commit_elect(acc_done);
            }
        }
    } else if (warp >= 6) {
        // ================= phasor generators: A chunk buffers ==========================================
        const int gt = t - 6 * 32;                               // 0..511
        const int row = gt & 127, quarter = gt >> 7;             // direction row, which 8 of the chunk's 32 microphones
        uint32_t g = 0;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
            const int f = unit / tiles, tile = unit - f * tiles;
            // fixed-point phases of this thread's 8 microphones per chunk: [chunk][i][quarter][row]
            const uint32_t *pf = phifix + (size_t)tile * (kTcMics * kTcDirs) + quarter * kTcDirs + row;
            const uint32_t bin = (uint32_t)(lo + f);
            uint32_t p[8], pn[8];                                 // next chunk, the one after (loads two chunks ahead)
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = __ldg(pf + (v3_order<QUARTER>(0) * 8 + i) * (4 * kTcDirs));
#pragma unroll
            for (int i = 0; i < 8; i++) pn[i] = __ldg(pf + (v3_order<QUARTER>(1) * 8 + i) * (4 * kTcDirs));
            for (int ci = 0; ci < kV3Chunks; ci++, g++) {
                const uint32_t ab = g & 1, aph = (g >> 1) & 1;
                uint32_t c[8];
#pragma unroll
                for (int i = 0; i < 8; i++) { c[i] = p[i]; p[i] = pn[i]; }
                if (ci + 2 < kV3Chunks) {
#pragma unroll
                    for (int i = 0; i < 8; i++) pn[i] = __ldg(pf + (v3_order<QUARTER>(ci + 2) * 8 + i) * (4 * kTcDirs));
                }
                uint32_t hi[8], lo8[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    float sn, cs;
                    if (dbg & 1) { sn = 0.25f; cs = 0.75f; } else if (dbg & 4) phasor_fix<false>(bin * c[i], cs, sn); else phasor_fix<true>(bin * c[i], cs, sn);
                    cs *= 256.0f; sn *= 256.0f;
                    const __half2 h = __floats2half2_rn(cs, sn);
                    const float2 hf = __half22float2(h);
                    const __half2 l = __floats2half2_rn(cs - hf.x, sn - hf.y);
                    hi[i] = *(const uint32_t *)&h;
                    lo8[i] = *(const uint32_t *)&l;
                }
                bfptx::mbar_wait(&a_empty[ab], aph ^ 1);
                unsigned char *dst = sA + ab * kV2BufA;
#pragma unroll
                for (int cc = 0; cc < 2; cc++) {
                    const uint32_t off = swz128((uint32_t)row * 128u + (uint32_t)(quarter * 2 + cc) * 16u);
                    *(uint4 *)(dst + off) = make_uint4(hi[4 * cc], hi[4 * cc + 1], hi[4 * cc + 2], hi[4 * cc + 3]);
                    *(uint4 *)(dst + kTcPlaneA + off) = make_uint4(lo8[4 * cc], lo8[4 * cc + 1], lo8[4 * cc + 2], lo8[4 * cc + 3]);
                }
                bfptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) bfptx::mbar_arrive(&a_full[ab]);
            }
        }
    } else {
        // ================= epilogue warps 0-3: q(d) = sum over 512 columns of (Y * colscale)^2 ========
        uint32_t w = 0;
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x, w++) {
            const int f = unit / tiles, tile = unit - f * tiles;
            const int d = tile * kTcDirs + warp * 32 + lane;
            const float sc = __ldg(colscale + f);          // per-bin scale 2^-(e+8), applied once to q
            float q0 = 0.0f, q1 = 0.0f, q2 = 0.0f, q3 = 0.0f;
            bfptx::mbar_wait(t0_done, w & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < 512; c0 += 32) {
                if (c0 == 256) { bfptx::mbar_wait(acc_done, w & 1); tc_fence_after(); }
                float v[32];
                tmem_ld32(lane_base + (uint32_t)c0, v);
#pragma unroll
                for (int i = 0; i < 8; i++) {                 // four independent chains
                    q0 = fmaf(v[4 * i], v[4 * i], q0);
                    q1 = fmaf(v[4 * i + 1], v[4 * i + 1], q1);
                    q2 = fmaf(v[4 * i + 2], v[4 * i + 2], q2);
                    q3 = fmaf(v[4 * i + 3], v[4 * i + 3], q3);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) bfptx::mbar_arrive(acc_free);
            if (d < D) qout[(size_t)f * D + d] = ((q0 + q1) + (q2 + q3)) * (sc * sc);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// P[d] = sum_f 1/q[f][d]   (fixed order: deterministic)
__global__ void mvdr_tc_reduce_kernel(const float *__restrict__ q, int F, int D, float *__restrict__ power)
{
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    float p = 0.0f;
    for (int f = 0; f < F; f++) p += 1.0f / q[(size_t)f * D + d];
    power[d] = p;
}

static DevBuf g_image, g_q, g_phi;
static const uint32_t *g_phifix = nullptr;

// reduced phase table of the generators (mvdr_tc_phi_kernel), rebuilt when the geometry changes
static int ensure_phi(const double *d_u, int D, int M, double scale, cudaStream_t st)
{
    static const double *phi_key = nullptr; static int phi_D = 0; static double phi_scale = 0.0;
    static uint64_t phi_gen = ~0ull;
    if (phi_key != d_u || phi_D != D || phi_scale != scale || phi_gen != fd_geometry_generation()) {
        const size_t cnt = (size_t)D * M;
        int rc = g_phi.ensure(cnt * sizeof(float2));
        if (rc) return rc;
        mvdr_tc_phi_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(d_u, cnt, scale, g_phi.as<float2>());
        BF_CHECK_LAUNCH();
        phi_key = d_u; phi_D = D; phi_scale = scale; phi_gen = fd_geometry_generation();
    }
    return BF_OK;
}

static DevBuf g_expo, g_binscale;

// Operand images (fp16 hi / lo, pre-swizzled) and per-bin scales of bins [f0, f0 + fc) from their L^-1.
int mvdr_tc_prepare(const float2 *d_linv, int M, int F, int f0, int fc, cudaStream_t st)
{
    if (M != kTcMics) { set_error(BF_ERR_CONFIG, "tensor-core MVDR steering is specialised for 256 microphones (got %d)", M); return BF_ERR_CONFIG; }
    int rc = g_image.ensure((size_t)F * kTcChunks * 2 * kTcPlaneB);
    if (rc) return rc;
    if ((rc = g_expo.ensure((size_t)F * sizeof(int)))) return rc;
    if ((rc = g_binscale.ensure((size_t)F * sizeof(float)))) return rc;
    mvdr_tc_rowscale_kernel<<<fc, kTcMics, 0, st>>>(d_linv, g_expo.as<int>(), g_binscale.as<float>(), f0);
    BF_CHECK_LAUNCH();
    mvdr_tc_prep3_kernel<<<dim3(kV3Chunks, fc), 256, 0, st>>>(d_linv, g_expo.as<int>(), g_image.as<unsigned char>(), f0);
    BF_CHECK_LAUNCH();
    count_launch(2);
    return BF_OK;
}

// The buffers a bin-sharded run all-gathers: image = F x bytes_per_bin bytes, binscale = F floats.
int mvdr_tc_operands(void **d_image, size_t *bytes_per_bin, void **d_binscale, int F)
{
    int rc = g_image.ensure((size_t)F * kTcChunks * 2 * kTcPlaneB);
    if (rc) return rc;
    if ((rc = g_binscale.ensure((size_t)F * sizeof(float)))) return rc;
    if ((rc = g_expo.ensure((size_t)F * sizeof(int)))) return rc;
    if (d_image) *d_image = g_image.p;
    if (bytes_per_bin) *bytes_per_bin = (size_t)kV3Chunks * 2 * kTcPlaneB;
    if (d_binscale) *d_binscale = g_binscale.p;
    return BF_OK;
}

// Steering contraction over all F bins from the prepared operand images, directions [0, D) of d_u.
int mvdr_tc_steer_only(const double *d_u, int M, int F, int lo, double bin_hz, double inv_c, int D, float *d_power,
                       cudaStream_t st)
{
    if (M != kTcMics) { set_error(BF_ERR_CONFIG, "tensor-core MVDR steering is specialised for 256 microphones (got %d)", M); return BF_ERR_CONFIG; }
    if (!g_image.p || !g_binscale.p) { set_error(BF_ERR_NOT_LOADED, "mvdr: no operand images (run the factor stage first)"); return BF_ERR_NOT_LOADED; }
    int rc = g_q.ensure((size_t)F * D * sizeof(float));
    if (rc) return rc;
    const int version = getenv("BF_MVDR_TC") ? atoi(getenv("BF_MVDR_TC")) : 4;
    const int tiles = (D + kTcDirs - 1) / kTcDirs;
    const size_t smem = 2 * kV2BufA + 2 * kV2SlotB + 128;
    const bool quarter = version >= 4;
    BF_CUDA(cudaFuncSetAttribute(mvdr_tc_steer_kernel3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BF_CUDA(cudaFuncSetAttribute(mvdr_tc_steer_kernel3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int units = tiles * F;
    const int grid = units < state().sm_count ? units : state().sm_count;
    {   // fixed-point phase table, rebuilt when the geometry (or the direction slice) changes
        static DevBuf fixbuf;
        static uint64_t fix_gen = ~0ull; static double fix_scale = 0.0; static int fix_D = 0;
        static const double *fix_u = nullptr;                  // a direction slice starts at another row of u
        if (fix_gen != fd_geometry_generation() || fix_scale != bin_hz * inv_c || fix_D != D || fix_u != d_u) {
            const size_t cnt = (size_t)tiles * kTcMics * kTcDirs;
            if ((rc = fixbuf.ensure(cnt * sizeof(uint32_t)))) return rc;
            mvdr_tc_phifix_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(d_u, D, bin_hz * inv_c, tiles, fixbuf.as<uint32_t>());
            BF_CHECK_LAUNCH();
            fix_gen = fd_geometry_generation(); fix_scale = bin_hz * inv_c; fix_D = D; fix_u = d_u;
        }
        g_phifix = fixbuf.as<uint32_t>();
    }
    if (lo + F > 1024) { set_error(BF_ERR_CONFIG, "tensor-core MVDR: bin index must stay below 1024"); return BF_ERR_CONFIG; }
    const int dbg = getenv("BF_MVDR_DBG") ? atoi(getenv("BF_MVDR_DBG")) : 0;
    if (quarter)
        mvdr_tc_steer_kernel3<true><<<grid, kV3Threads, smem, st>>>(g_image.as<unsigned char>(), g_phifix,
                                                                   g_binscale.as<float>(), F, lo, D, tiles, g_q.as<float>(), dbg);
    else
        mvdr_tc_steer_kernel3<false><<<grid, kV3Threads, smem, st>>>(g_image.as<unsigned char>(), g_phifix,
                                                                    g_binscale.as<float>(), F, lo, D, tiles, g_q.as<float>(), dbg);
    BF_CHECK_LAUNCH();
    mvdr_tc_reduce_kernel<<<(D + 255) / 256, 256, 0, st>>>(g_q.as<float>(), F, D, d_power);
    BF_CHECK_LAUNCH();
    count_launch(3);
    return BF_OK;
}

int mvdr_steer_tc(const float2 *d_linv, const double *d_u, int M, int F, int lo, double bin_hz, double inv_c,
                  int D, float *d_power, cudaStream_t st)
{
    if (M != kTcMics) { set_error(BF_ERR_CONFIG, "tensor-core MVDR steering is specialised for 256 microphones (got %d)", M); return BF_ERR_CONFIG; }
    int rc = g_image.ensure((size_t)F * kTcChunks * 2 * kTcPlaneB);
    if (rc) return rc;
    if ((rc = g_q.ensure((size_t)F * D * sizeof(float)))) return rc;
    const int version = getenv("BF_MVDR_TC") ? atoi(getenv("BF_MVDR_TC")) : 4;
    if (version >= 3) {
        if ((rc = mvdr_tc_prepare(d_linv, M, F, 0, F, st))) return rc;
        return mvdr_tc_steer_only(d_u, M, F, lo, bin_hz, inv_c, D, d_power, st);
    }
    mvdr_tc_prep_kernel<<<dim3(kTcChunks, F), 256, 0, st>>>(d_linv, g_image.as<unsigned char>());
    BF_CHECK_LAUNCH();
    if (version == 1) {
        const size_t smem = 2 * kTcPlaneA + 2 * kTcPlaneB + 64;
        BF_CUDA(cudaFuncSetAttribute(mvdr_tc_steer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mvdr_tc_steer_kernel<<<dim3((D + kTcDirs - 1) / kTcDirs, F), 128, smem, st>>>(
            g_image.as<unsigned char>(), d_u, F, lo, bin_hz, inv_c, D, g_q.as<float>());
    } else {
        const int tiles = (D + kTcDirs - 1) / kTcDirs;
        const size_t smem = 2 * kV2BufA + 2 * kV2SlotB + 128;
        BF_CUDA(cudaFuncSetAttribute(mvdr_tc_steer_kernel2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int units = tiles * F;
        const int grid = units < state().sm_count ? units : state().sm_count;
        if ((rc = ensure_phi(d_u, D, M, bin_hz * inv_c, st))) return rc;
        DevBuf &phi = g_phi;
        if (lo + F > 1024) { set_error(BF_ERR_CONFIG, "tensor-core MVDR: bin index must stay below 1024"); return BF_ERR_CONFIG; }
        mvdr_tc_steer_kernel2<<<grid, kV2Threads, smem, st>>>(g_image.as<unsigned char>(), phi.as<float2>(), F, lo,
                                                             D, tiles, g_q.as<float>());
    }
    BF_CHECK_LAUNCH();
    mvdr_tc_reduce_kernel<<<(D + 255) / 256, 256, 0, st>>>(g_q.as<float>(), F, D, d_power);
    BF_CHECK_LAUNCH();
    count_launch(3);
    return BF_OK;
}

}  // namespace bf
