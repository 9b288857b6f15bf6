/*
 * bf_b200.h -- C ABI of libbf_b200.so, the B200 (sm_100a) beamformer.
 *
 * Part 1 re-declares, with identical names and signatures, the C functions
 * that the reference's Cython boundary binds for the beamforming hot path
 * (PC/src/main.pyx:45-90 `cdef extern`, PC/src/benchmark.pyx:28-55).  A build
 * of the reference that links this library instead of compiling
 * PC/src/algorithms/*.c gets the CUDA path with no source change (see
 * INTEGRATION.md).  Pointer arguments of part 1 are HOST pointers, exactly as
 * in the reference; each call copies in, runs on the current CUDA device and
 * copies the result back before returning.
 *
 * Part 2 is what the reference cannot express because its sizes are
 * compile-time macros (PC/src/config.h generated from config.json): run-time
 * configuration, error reporting, and device-resident / batched / sharded
 * entry points used by the host layer, bench.py and the multi-GPU path.
 *
 * Plain C: pointers and sizes only, no CUDA or torch types.  `stream` arguments
 * are a cudaStream_t passed as void* (NULL = default stream).
 *
 * All file:line citations are relative to /root/reference/PC/src/.
 */
#ifndef BF_B200_H
#define BF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ===================================================================== *
 *  Part 1 -- drop-in names                                              *
 * ===================================================================== */

/* ---- algorithms/pad_and_sum.h:5-13 ---------------------------------- */
void pad_delay(float *signal, float *out, int pos_pad);
void miso_pad(float *signals, float *out, int *adaptive_array, int n, int offset);
void miso_pad2(float *signals, float *out, int *adaptive_array, int n, int offset);
void mimo_pad(float *signals, float *image, int *adaptive_array, int n);
void load_coefficients_pad(int *whole_samples, int n);
void load_coefficients_pad2(int *whole_miso, int n);
void unload_coefficients_pad(void);
void unload_coefficients_pad2(void);

/* ---- algorithms/lerp_and_sum.h:4-12 --------------------------------- */
void lerp_delay(float *signal, float *out, float h, int pad);
void miso_lerp(float *signals, float *out, int *adaptive_array, int n, int offset);
void mimo_lerp(float *signals, float *image, int *adaptive_array, int n);
void load_coefficients_lerp(float *delays, int n);
void unload_coefficients_lerp(void);

/* ---- algorithms/convolve_and_sum.h:4-22 ------------------------------
 * (convolve_naive is declared there but never defined in the reference; it is
 * not exported here either.) */
void convolve_delay_naive_add(float *signal, float *h, float *out);
void convolve_delay_vectorized(float *signal, float *h, float *out);
void convolve_delay_vectorized_add(float *signal, float *h, float *out);
void convolve_delay_naive(float *signal, float *out, float *h);
void miso_convolve_naive(float *signals, float *out, int *adaptive_array, int n, int offset);
void mimo_convolve_naive(float *signals, float *image, int *adaptive_array, int n);
void miso_convolve_vectorized(float *signals, float *out, int *adaptive_array, int n, int offset);
void mimo_convolve_vectorized(float *signals, float *image, int *adaptive_array, int n);
void load_coefficients_convolve(float *h, int n);
void unload_coefficients_convolve(void);

/* ---- algorithms/hybrid_convolve_and_sum.h:4-12 ---------------------- */
void convolve_hybrid_delay_add(float *signal, float *h, int pad, float *out);
void miso_convolve_hybrid(float *signals, float *out, int *adaptive_array, int n, int offset);
void mimo_convolve_hybrid(float *signals, float *image, int *adaptive_array, int n);
void load_coefficients_convolve_hybrid(float *h, int n);
void unload_coefficients_convolve_hybrid(void);

/* ---- api.h:11-20 -------------------------------------------------------
 * In the reference these fetch the latest sample buffer with get_data()
 * (api.c:830-859, SysV shared memory fed by the UDP receiver, which stays host
 * C) and then call the kernels above.  Here the buffer comes from the callback
 * registered with bf_set_data_source(); without one they fail (bf_last_error). */
void pad_mimo(float *image, int *adaptive_array, int n);
void lerp_mimo(float *image, int *adaptive_array, int n);
void convolve_mimo_naive(float *image, int *adaptive_array, int n);
void convolve_mimo_vectorized(float *image, int *adaptive_array, int n);
void mimo_truncated(float *image, int *adaptive_array, int n);
void load_coefficients2(int *whole_samples, int n);
void miso_steer_listen(float *out, int *adaptive_array, int n, int steer_offset);

/* ===================================================================== *
 *  Part 2 -- run-time configuration and device-side entry points        *
 * ===================================================================== */

/* The config.json "general" constants the hot path reads as macros
 * (config.json:3-21, config.pxd:5-26). */
typedef struct bf_config {
    int n_microphones;      /* N_MICROPHONES: rows of the sample buffer        */
    int n_samples;          /* N_SAMPLES:     samples per row (block length)   */
    int n_taps;             /* N_TAPS                                          */
    int max_res_x;          /* MAX_RES_X                                       */
    int max_res_y;          /* MAX_RES_Y  (D = max_res_x * max_res_y)          */
    float mic_gain;         /* MIC_GAIN (api.c:519-523 MISO post-scale)        */
    int fir_fused;          /* -1 auto (n_taps <= 16), 0 mul+add, 1 fma: how the
                               sequential FIR tap loop rounds, see oracle.c     */
} bf_config;

enum {
    BF_OK = 0,
    BF_ERR_CONFIG = 1,      /* bad size / unsupported configuration            */
    BF_ERR_NOT_LOADED = 2,  /* coefficient table missing or too small          */
    BF_ERR_CUDA = 3,        /* CUDA runtime error (text in bf_last_error)      */
    BF_ERR_ARG = 4          /* bad pointer / range argument                    */
};

enum {                      /* delay algorithms                                */
    BF_ALGO_PAD = 0,        /* integer zero-pad delay      (pad_and_sum.c)     */
    BF_ALGO_LERP = 1,       /* integer + linear interp.    (lerp_and_sum.c)    */
    BF_ALGO_FIR_SEQ = 2,    /* N_TAPS FIR, sequential taps (convolve naive)    */
    BF_ALGO_FIR_LANES = 3,  /* N_TAPS FIR, 8-lane AVX order (convolve vector.) */
    BF_ALGO_HYBRID = 4      /* integer pad + FIR fraction  (hybrid_convolve)   */
};

int bf_configure(const bf_config *cfg);        /* default = stock config.json   */
int bf_get_config(bf_config *cfg);
const char *bf_last_error(void);               /* thread-local, never NULL      */
int bf_last_status(void);                      /* status of the last part-1 call */
int bf_device_count(void);
int bf_set_device(int ordinal);                /* device used by this thread's calls */
const char *bf_version(void);

/* Source of sample buffers for the api.h wrappers: called with a host pointer
 * to n_microphones*n_samples floats to fill (same contract as get_data()). */
typedef void (*bf_data_source_fn)(float *signals);
void bf_set_data_source(bf_data_source_fn fn);

/* Implementation selector for the power-map kernels: 0 = tiled TMA kernel
 * (default), 1 = simple one-block-per-direction kernel (any n_samples <= 1024).
 * exact_sum: 1 = reproduce the reference's in-order sum of squares bit for bit
 * (default), 0 = warp-shuffle tree (<= 1e-6 relative). */
int bf_set_kernel_options(int simple_kernel, int exact_sum);

/* ---- device-resident, batched, direction-sharded power maps ----------
 * d_signals  device float [frames][n_microphones][n_samples]
 * d_images   device float [frames][D]; only [d_begin, d_begin+d_count) written
 * d_mic_ids  device int [n]  (adaptive_array: mic id of table column m)
 * Uses the table loaded for `algo` by the matching load_coefficients_* (or
 * bf_load_table_dev).  Asynchronous on `stream`. */
int bf_mimo_dev(int algo, const float *d_signals, float *d_images, int frames,
                const int *d_mic_ids, int n, int d_begin, int d_count, void *stream);

/* Same, with an explicit output layout: the power of frame f, direction d is stored at
 *   d_images[f * frame_stride + (d - d_origin) * dir_stride]
 * bf_mimo_dev == (frame_stride = D, dir_stride = 1, d_origin = 0).  Direction-major maps
 * (frame_stride = 1, dir_stride = frames) make every rank's direction slice one contiguous
 * block, so a single in-place NCCL all-gather assembles the sharded map. */
int bf_mimo_dev_ex(int algo, const float *d_signals, float *d_images, int frames,
                   const int *d_mic_ids, int n, int d_begin, int d_count, long frame_stride,
                   long dir_stride, int d_origin, void *stream);

/* ---- device-resident MISO stream (BASELINE config C2) -----------------
 * d_signals device float [blocks][n_microphones][n_samples]
 * d_out     device float [blocks][n_samples]
 * offset    row offset into the table (= direction * n), as in miso_pad
 * scale     0: raw sum (miso_pad / miso_lerp); 1: out/n*mic_gain (api.c:519-523) */
int bf_miso_dev(int algo, const float *d_signals, float *d_out, int blocks,
                const int *d_mic_ids, int n, int offset, int scale, void *stream);

/* ---- batch replay with HOST buffers (BASELINE config C5: recorded streams) ---------
 * signals float [frames][n_microphones][n_samples], images float [frames][D], both HOST
 * pointers (pageable or pinned; pinned buffers are copied without staging).  Frames are
 * processed in chunks: the host-to-device copy of chunk k+1 and the device-to-host copy of
 * chunk k-1 overlap the kernel of chunk k (three streams).  Returns when `images` is
 * complete.  Same result as `frames` successive mimo_pad / mimo_lerp calls. */
int bf_mimo_host_batch(int algo, const float *signals, float *images, int frames,
                       const int *adaptive_array, int n);
/* The chunk sizes bf_mimo_host_batch uses for `frames` frames (host logic only, no device needed): small chunks
 * at both ends -- the first H2D copy and the last D2H copy are the two the call cannot overlap with a kernel --
 * and up to 32 frames in between.  Writes at most `capacity` sizes, returns the number of chunks (-1: bad argument). */
int bf_host_batch_schedule(int frames, int *chunk_frames, int capacity);

/* ---- tables straight from device memory -------------------------------
 * algo PAD: d_table = int32 [count]; LERP/HYBRID: float32 delays [count];
 * FIR_*: float32 taps [count].  Same semantics as load_coefficients_*. */
int bf_load_table_dev(int algo, const void *d_table, size_t count);

/* ---- delay-table generator (directions.pyx:90-124) --------------------
 * Host computes the O(X+Y+n) scalars exactly as the reference does; the device
 * evaluates the D*n table in float64 in the reference's operation order.
 *   k        (double)((float)fs / (float)c)
 *   x_scan   double [res_x], y_scan double [res_y]  (np.linspace values)
 *   z2       (double)powf(z, 2)
 *   mic_x/y  double [n] coordinates of the active microphones
 * Outputs (any may be NULL), all HOST pointers, row-major [res_x*res_y][n]:
 *   delays_f64, whole_i32 (trunc), delays_f32 (round to nearest).
 * If load_algo >= 0 the generated table is also installed on the device as if
 * load_coefficients_{pad,lerp,convolve_hybrid} had been called (no host trip). */
int bf_generate_delays(double k, const double *x_scan, int res_x, const double *y_scan,
                       int res_y, double z2, const double *mic_x, const double *mic_y, int n,
                       double *delays_f64, int *whole_i32, float *delays_f32, int load_algo);

/* The lerp split of load_coefficients_lerp (lerp_and_sum.c:139-153) as data:
 * copies the device-resident tables back (HOST pointers, may be NULL). */
int bf_get_lerp_tables(int *whole, float *weight, size_t count);
/* ... and the hybrid split (hybrid_convolve_and_sum.c:161-180). */
int bf_get_hybrid_tables(int *whole, float *taps, size_t count);

/* ---- frequency-domain delay-and-sum (PC/application/realtime_scripts/) ---------------
 * bf_fd_setup installs the scan grid and microphone geometry of
 * calc_phase_shift_cartesian.py:8-50 (host computes the O(X+Y+M) arrays exactly as the
 * reference; the (bin, mic, direction) phasor table is never materialised):
 *   lo_bin/hi_bin   threshold_freq_{lower,upper}_idx (bins [lo, hi) of the rfft)
 *   x_scan, y_scan  np.linspace scan axes;  z = config.Z
 *   mic_x, mic_y    r_prime_all coordinates of all n_mics microphones
 *   active          indices of the microphones used (phase_shift[:, active_mics])
 * bf_fd_das: signals HOST float [frames][n_mics][n_samples] (the (M, N) buffer the reference
 * transposes before the call, camera.py:72), heatmap HOST float [frames][res_x*res_y]:
 * rfft -> band limit -> steer -> |.|^2 -> sum over bins -> (normalise != 0) P/max(P), or all
 * zeros when max(P) < threshold (beam_forming_algorithm.py:50-63). */
int bf_fd_setup(int n_mics, int n_samples, double fs, double c, int lo_bin, int hi_bin,
                const double *x_scan, int res_x, const double *y_scan, int res_y, double z,
                const double *mic_x, const double *mic_y, const int *active, int n_active);
int bf_fd_das(const float *signals, float *heatmap, int frames, float threshold, int normalise);
int bf_fd_das_dev(const float *d_signals, float *d_heatmap, int frames, float threshold,
                  int normalise, void *stream);

/* ---- frequency-domain MVDR (Capon) map; geometry/band from bf_fd_setup ------------------
 * Not in the reference (parity unpinned, see csrc/fd_mvdr.cu).  snapshots: float
 * [K][n_mics][n_samples]; power: float [res_x*res_y] = sum over bins of
 * 1 / Re(a^H R_f^-1 a), R_f = 1/K sum_k x_k x_k^H + loading * tr(R_f)/M * I. */
int bf_fd_mvdr(const float *snapshots, float *power, int K, double loading);
int bf_fd_mvdr_dev(const float *d_snapshots, float *d_power, int K, double loading, void *stream);
/* Direction-sharded frequency-domain maps ("FD path: shard directions"): only directions
 * [d_begin, d_begin + d_count) of the grid are steered -- d_power is float [d_count] (MVDR) or
 * [frames][d_count] un-normalised (DAS).  Spectra, covariance, factor and inverse are computed in full on every
 * rank (0.5 % of the work).  bf_peer_scatter / PeerGather assemble the slices; bf_fd_normalise_dev applies the
 * reference's P / max(P) (or all zeros below the threshold) to an assembled [frames][D] map. */
int bf_fd_mvdr_dev_slice(const float *d_snapshots, float *d_power, int K, double loading, int d_begin,
                         int d_count, void *stream);
int bf_fd_das_dev_slice(const float *d_signals, float *d_power, int frames, int d_begin, int d_count,
                        void *stream);
int bf_fd_normalise_dev(float *d_heatmap, int frames, float threshold, int normalise, void *stream);
/* The same map in two phases, for runs that ALSO shard the float64 stages (31 % of a C4 map's time) by BINS:
 * bf_fd_mvdr_factor_dev computes spectra, covariance + loading, Cholesky, inverse and the tensor-core operand
 * images of bins [f_begin, f_begin + f_count) of the band; bf_fd_mvdr_operands exposes the image buffer
 * (bins x bytes_per_bin bytes) and the per-bin scales (bins floats) so that the ranks can all-gather them
 * (NCCL over NVLink, lib/sharded.py:fd_mvdr_sharded_bins); bf_fd_mvdr_steer_dev then steers directions
 * [d_begin, d_begin + d_count) over all bins.  256 microphones (the tcgen05 path) only. */
int bf_fd_mvdr_factor_dev(const float *d_snapshots, int K, double loading, int f_begin, int f_count, void *stream);
int bf_fd_mvdr_operands(void **d_image, size_t *bytes_per_bin, void **d_binscale);
int bf_fd_mvdr_steer_dev(float *d_power, int d_begin, int d_count, void *stream);
/* device milliseconds of the stages of the last MVDR call: FFT, covariance + loading, Cholesky,
 * triangular inverse, steering contraction */
int bf_fd_mvdr_timings(float *ms5);
/* loaded covariance of the last MVDR call: HOST double [bins][M][M][2] */
int bf_fd_get_covariance(double *cov, size_t count);

/* Wire format of one sample instant (receiver.h:51-59 `msg`): what the Zybo sends per UDP datagram
 * and what lib/capture.py extracts from pcap / pcapng files.  bf_ingest_dev takes the `stream`
 * payloads of n_samples consecutive datagrams with this 8-byte header stripped. */
typedef struct bf_datagram_header {
    uint16_t frequency;      /* sampling rate, Hz */
    int8_t n_arrays;         /* 8x8 arrays daisy-chained (1..4) */
    int8_t protocol_ver;
    int32_t counter;         /* sample counter: gaps = dropped datagrams */
    /* int32_t stream[N_MICROPHONES] follows */
} bf_datagram_header;

/* ---- the reference's shared-memory records with run-time sizes ---------------
 * api.h:26-30   typedef struct { int can_read; float out[N_SAMPLES]; } paData;
 * api.h:32-38   typedef struct { int steer_offset; float signals[BUFFER_LENGTH];
 *                                int adaptive_array[N_MICROPHONES]; int n; } Miso;
 * receiver.h:31-36  typedef struct { int index; float data[BUFFER_LENGTH * 4]; int counter; } ring_buffer;
 * (receiver.h:51-59 `msg` is bf_datagram_header + the payload.)  The reference fixes the array sizes with macros;
 * here they follow from the configured sizes: bf_layout_* return the record size and the byte offset of every
 * member in declaration order.  bf_miso_record_listen runs one iteration of the audio child's loop
 * (api.c:505-529) on a Miso record in host / SysV shared memory: the steered beam of miso->signals, and -- when
 * padata_record is not NULL -- the post-scale out / n * MIC_GAIN into pa->out with pa->can_read = 1. */
typedef struct bf_record_layout { size_t size; size_t off[4]; } bf_record_layout;
bf_record_layout bf_layout_miso(int n_microphones, int n_samples);        /* steer_offset, signals, adaptive_array, n */
bf_record_layout bf_layout_padata(int n_samples);                         /* can_read, out */
bf_record_layout bf_layout_ring_buffer(int n_microphones, int n_samples); /* index, data, counter */
int bf_miso_record_listen(const void *miso_record, void *padata_record);

/* ---- wire-format ingest (receiver.c:94-151), SURVEY 8f "next" #1 -------------------------
 * d_stream  device int32 [frames][n_samples][n_microphones]: the `stream` payload of
 *           n_samples consecutive datagrams (receiver.h:51-59), header stripped
 * d_signals device float [frames][n_microphones][n_samples] (channels >= n_arrays*rows*cols
 *           are left untouched, like the reference's ring buffer)
 * quirk     1 = the reference's odd-row index `row + COLUMNS - x`, 0 = corrected
 * d_zero_mask optional uint8[n_channels]: channels to clear (api.c:835-858), or NULL */
int bf_ingest_dev(const int *d_stream, float *d_signals, int frames, int n_arrays, int rows, int cols,
                  double norm, int quirk, const unsigned char *d_zero_mask, void *stream);
/* Batch replay of a stored stream (BASELINE config C5): frame f is the n_samples-datagram WINDOW that begins at
 * datagram d_starts[f] (device int64 [frames], e.g. floor(k*fs/fps)) of d_stream, device int32
 * [total_datagrams][n_microphones]; datagrams beyond the end read as silence.  Same conversion as
 * bf_ingest_dev -- the window gather and the wire-format conversion in one pass. */
int bf_ingest_windows_dev(const int *d_stream, long total_datagrams, const long *d_starts, float *d_signals,
                          int frames, int n_arrays, int rows, int cols, double norm, int quirk,
                          const unsigned char *d_zero_mask, void *stream);

/* ---- window gather for batch replay (BASELINE config C5) ---------------------------------
 * d_recording device float [n_microphones][samples] (PC/record.py .npy layout), d_starts device
 * int64 [frames] first sample of each frame's window; d_frames device float
 * [frames][n_microphones][n_samples] (zero beyond the end of the recording). */
int bf_window_dev(const float *d_recording, long samples, const long *d_starts, int frames,
                  float *d_frames, void *stream);

/* ---- fused power maps + all-gather over NVLink peer memory -----------------------
 * One process per GPU.  Every rank allocates a gather buffer float [world][frames][per_rank] and
 * a flag array int64 [world] with bf_dev_alloc, exports both (bf_ipc_export), exchanges the
 * 64-byte handles with its peers (any transport; lib/sharded.py uses torch.distributed) and opens
 * theirs (bf_ipc_open).  bf_mimo_dev_gather then computes this rank's direction slice and stores
 * every finished value into slice `rank` of ALL `world` buffers (its own and, through NVLink P2P
 * stores issued by the kernel's epilogue, its peers') -- no collective call on the data path.
 * bf_gather_signal publishes "step complete" into every rank's flag array after the kernel;
 * bf_gather_wait blocks the stream until all ranks published `step` (sets *d_timed_out after ~20 s
 * instead of hanging).  Only the tiled pad / lerp kernels (N_SAMPLES 64/128/256) store to peers. */
int bf_dev_alloc(size_t bytes, void **d_ptr);            /* cudaMalloc + zero fill (exportable) */
int bf_dev_free(void *d_ptr);
int bf_ipc_export(void *d_ptr, unsigned char *handle64);
int bf_ipc_open(const unsigned char *handle64, void **d_ptr);
int bf_ipc_close(void *d_ptr);
int bf_mimo_dev_gather(int algo, const float *d_signals, int frames, const int *d_mic_ids, int n,
                       int d_begin, int d_count, int rank, int world, void *const *gather_bufs,
                       long per_rank, void *stream);
/* Same, with the step flags handled inside the kernel (no extra launches): before its first peer
 * store a warp waits until every rank published wait_seq (0 = do not wait); the last CTA to finish
 * publishes signal_seq into every rank's flag array (0 = do not signal). */
int bf_mimo_dev_gather_sync(int algo, const float *d_signals, int frames, const int *d_mic_ids, int n,
                            int d_begin, int d_count, int rank, int world, void *const *gather_bufs,
                            long per_rank, void *const *flag_arrays, long long wait_seq, long long signal_seq,
                            int *d_timed_out, void *stream);
/* Asynchronous copy between device buffers of this GPU and of a peer (a pointer from bf_ipc_open) on the copy
 * engines (no SMs: it proceeds while a persistent kernel occupies the whole GPU).  Used by lib/sharded.py:PeerInput
 * for the all-gather of the INPUT frames of a sharded step; follow it with bf_gather_signal. */
int bf_peer_copy(void *dst, const void *src, size_t bytes, void *stream);

/* Launch plan of the tiled power-map kernel, host logic only (no device needed; for tests and tools): a launch
 * of `groups` direction groups (8 directions each) x `frames` frames on `sm_count` SMs runs `grid` CTAs of
 * `consumer_warps` + 1 warps; ranges = 1: every CTA owns a contiguous range of (frame, group) units, 0: whole tiles
 * handed out round-robin.  bf_mimo_walk lists the tiles CTA `block` processes, in order, as (frame, first group,
 * groups) triples (at most `capacity` of them are written) and returns their number, or -1 on a bad argument. */
int bf_mimo_plan(int lerp, int groups, int frames, int overlap, int sm_count, int *consumer_warps, int *grid,
                 int *ranges);
long bf_mimo_walk(int lerp, int groups, int frames, int overlap, int sm_count, int block, int *tiles, long capacity);

/* Overlapping steps (pad, bf_mimo_dev_gather_sync with signal_seq > 0): on = 1 launches every following step with
 * programmatic stream serialisation -- its CTAs take over SMs as the CTAs of the previous kernel on the stream exit,
 * so a run of back-to-back steps has no idle SMs in the last round of tiles, no launch gap and no prologue between
 * steps.  The caller promises (a) consecutive steps store to different buffers (a PeerGather of depth >= 2 does) and
 * (b) a step's inputs were complete before the PREVIOUS kernel on the stream was launched (nothing that produces
 * them is enqueued between two steps).  Work enqueued AFTER a step still sees it, and everything before it,
 * complete.  on = 0 (default): ordinary stream order.  on < 0: query.  Returns the previous setting. */
int bf_gather_overlap(int on);
int bf_gather_signal(void *const *flag_arrays /* host array of `world` device pointers */, int world, int rank,
                     long long step, void *stream);
int bf_gather_wait(const void *d_my_flags, int world, long long step, int *d_timed_out, void *stream);
/* Device-side all-gather of a finished slice: d_src float [frames][count] -> slice `rank` of every rank's gather
 * buffer float [world][frames][per_rank] (peer stores over NVLink).  For producers that do not store to the
 * peers themselves (the frequency-domain maps); follow it with bf_gather_signal. */
int bf_peer_scatter(const float *d_src, long count, int frames, int rank, int world, void *const *gather_bufs,
                    long per_rank, void *stream);

/* ---- heat-map post-processing (SURVEY 8f "next" #2) -----------------------------------------
 * The step after the beamformer: PC/src/visual.py:130-171 calculate_heatmap (log scale),
 * 173-205 calculate_heatmap_fft (linear), 293-322 find_power_center, 227-291
 * calculate_heatmap_with_detection; PC/sensorfusion/decider.py:16-24 get_entropy.
 * Maps are float [res_x][res_y] exactly as the queue payload. */
typedef struct bf_heat_info {
    float max_power;    /* np.max(image) */
    float min_power;    /* np.min(np.clip(image, 1e-12, None)) */
    float log_span;     /* np.max(img) after img -= log10(min) (log scale only) */
    float smooth_max;   /* max of the 5x5 Gaussian-blurred map */
    double center_col;  /* find_power_center()[0]: centroid along axis 1 (the res_y index) */
    double center_row;  /* find_power_center()[1]: centroid along axis 0 (the res_x index) */
    int overlay;        /* should_overlay */
    int painted;        /* pixels at or above `amount` */
    int fallback;       /* 1: the centroid fell back to the arg-max of the blurred map */
    int reserved;
} bf_heat_info;

/* Fills the 256x3 colour table of generate_color_map() (visual.py:27-48): Matplotlib's jet,
 * reversed, truncated to uint8.  Host only. */
int bf_jet_lut(unsigned char *lut768);

/* d_maps   device float [frames] maps, frame f at d_maps + f*frame_stride
 * lut      HOST uint8[256][3] colour table, or NULL for bf_jet_lut's
 * d_small  device uint8 [frames][res_y][res_x][3]: small_heatmap, stored flipped
 *          (small[res_y-1-y][res_x-1-x] = lut[index]) as visual.py:166 does
 * d_index  optional device int16 [frames][res_x][res_y]: colour index or -1 (not painted)
 * d_info   device bf_heat_info [frames]
 * log_scale 1 = calculate_heatmap, 0 = calculate_heatmap_fft (pass its threshold*1e6) */
int bf_heatmap_dev(const float *d_maps, int frames, long frame_stride, int res_x, int res_y,
                   float threshold, float amount, int exponent, int log_scale,
                   const unsigned char *lut, unsigned char *d_small, short *d_index,
                   bf_heat_info *d_info, void *stream);

/* cv2.resize(src, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR) for 8-bit images with 1..4
 * interleaved channels, bit-identical to OpenCV 4.13; a batch of `frames` images per call. */
int bf_resize_linear_u8_dev(const unsigned char *d_src, int frames, int src_h, int src_w, int channels,
                            unsigned char *d_dst, int dst_h, int dst_w, void *stream);

/* get_entropy (decider.py:16-24) of `frames` 8-bit images of bytes_per_frame bytes each:
 * d_confidence[f] = 1 / (1 + H(f)). */
int bf_entropy_dev(const unsigned char *d_img, int frames, long bytes_per_frame, double *d_confidence,
                   void *stream);

/* Host-pointer form of the whole stage: maps [frames][res_x][res_y] in; heat maps
 * uint8 [frames][out_h][out_w][3], info [frames] and (optional) confidence [frames] out. */
int bf_heatmap(const float *maps, int frames, int res_x, int res_y, float threshold, float amount,
               int exponent, int log_scale, const unsigned char *lut, int out_w, int out_h,
               unsigned char *heat_out, bf_heat_info *info_out, double *confidence_out);

/* ---- peak tracking filter (SURVEY 8f "next" #4), host only ----------------------------------
 * KalmanFilter3D of PC/src/kf.hpp:36-165 (lib.kf.CyKF, kf.pyx:18-46): updatef / getStatef /
 * predictf on an opaque handle. */
void *bf_kf_create(void);
void bf_kf_destroy(void *kf);
int bf_kf_update(void *kf, const float *measurement_xyz);
int bf_kf_get_state(void *kf, float *xyz);
int bf_kf_predict(void *kf, int n, float *xyz);

/* Counters for bench.py: kernels launched by this library since the last reset. */
uint64_t bf_kernel_launches(int reset);

#ifdef __cplusplus
}
#endif
#endif /* BF_B200_H */
