// bf_api.cu -- process state and the C ABI of libbf_b200.so (include/bf_b200.h).
//
// Part-1 entry points keep the reference's names, signatures and ownership rules
// (caller owns every pointer; load_* copies the table; one table per algorithm
// per process).  They take HOST pointers: stage through
// pinned memory, run on the current device, copy back, return.  CUDA is initialised
// lazily on first use so that the library can be loaded before fork() the way the
// reference's producer processes do (main.pyx:702-721).
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "bf_common.cuh"

namespace bf {

static thread_local char g_err[512] = "";
static thread_local int g_status = BF_OK;
static std::atomic<uint64_t> g_launches{0};

void set_error(int status, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    g_status = status;
    if (getenv("BF_VERBOSE")) fprintf(stderr, "[bf_b200] error %d: %s\n", status, g_err);
}
void clear_error() { g_err[0] = 0; g_status = BF_OK; }
int last_status() { return g_status; }
void count_launch(int n) { g_launches += (uint64_t)n; }

int DevBuf::ensure(size_t need)
{
    if (need <= bytes && p) return BF_OK;
    if (p) { cudaFree(p); p = nullptr; bytes = 0; }
    size_t want = need < 256 ? 256 : need;
    BF_CUDA(cudaMalloc(&p, want));
    bytes = want;
    return BF_OK;
}
void DevBuf::release()
{
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
}

State &state()
{
    static State S = [] {
        State s;
        // stock PC/src/config.json values
        s.cfg.n_microphones = 256; s.cfg.n_samples = 256; s.cfg.n_taps = 8;
        s.cfg.max_res_x = 57; s.cfg.max_res_y = 32; s.cfg.mic_gain = 128.0f; s.cfg.fir_fused = -1;
        return s;
    }();
    return S;
}

int ensure_device()
{
    State &S = state();
    if (S.sm_count > 0) return BF_OK;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error(BF_ERR_CUDA, "no CUDA device available (%s): this library has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return BF_ERR_CUDA;
    }
    int dev = 0;
    if (S.device >= 0) { BF_CUDA(cudaSetDevice(S.device)); dev = S.device; }
    else BF_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    BF_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) {
        set_error(BF_ERR_CUDA, "device %d is sm_%d%d; this build targets sm_100a only", dev,
                  prop.major, prop.minor);
        return BF_ERR_CUDA;
    }
    S.device = dev;
    S.sm_count = prop.multiProcessorCount;
    return BF_OK;
}

int ensure_pinned(size_t bytes)
{
    State &S = state();
    if (bytes <= S.h_pinned_bytes) return BF_OK;
    if (S.h_pinned) cudaFreeHost(S.h_pinned);
    S.h_pinned = nullptr; S.h_pinned_bytes = 0;
    BF_CUDA(cudaMallocHost(&S.h_pinned, bytes));
    S.h_pinned_bytes = bytes;
    return BF_OK;
}

static std::mutex g_mu;   // part-1 calls share global tables/staging, like the reference

// The device entry points (bf_*_dev) take a stream but share process-global scratch with every other call
// (lerp first differences, group tables and FIR re-layouts built lazily on the first caller's stream, the
// last-CTA ticket of the fused gather).  They are serialised here: the host side by g_mu, and the device
// side by ordering a call behind the previous one whenever it arrives on a DIFFERENT stream (event
// record on the old stream, wait on the new one).  Calls that stay on one stream pay nothing.
static cudaStream_t g_last_stream = nullptr;
static bool g_have_last_stream = false;
static cudaEvent_t g_order_event = nullptr;
static int order_behind_previous(cudaStream_t st)
{
    if (g_have_last_stream && g_last_stream != st) {
        if (!g_order_event) BF_CUDA(cudaEventCreateWithFlags(&g_order_event, cudaEventDisableTiming));
        if (cudaEventRecord(g_order_event, g_last_stream) == cudaSuccess) {
            BF_CUDA(cudaStreamWaitEvent(st, g_order_event, 0));
        } else {
            cudaGetLastError();                  // the previous stream is gone: its work has completed
        }
    }
    g_last_stream = st;
    g_have_last_stream = true;
    return BF_OK;
}

// ---- table installation --------------------------------------------------------
static int install_pad(DevBuf &buf, size_t &count, int &wmax, GroupTable *gt, const int *src,
                       size_t n, bool src_on_device)
{
    int rc = ensure_device();
    if (rc) return rc;
    if (n == 0) { count = 0; return BF_OK; }
    rc = buf.ensure(n * sizeof(int));
    if (rc) return rc;
    BF_CUDA(cudaMemcpy(buf.p, src, n * sizeof(int),
                       src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    count = n;
    rc = launch_max_abs_i32(buf.as<int>(), n, &wmax);
    if (rc) return rc;
    if (gt) gt->n = -1;            // invalidate the cached group layout
    state().tab.version++;
    return BF_OK;
}

static int install_lerp(const float *src, size_t n, bool src_on_device)
{
    State &S = state();
    int rc = ensure_device();
    if (rc) return rc;
    Tables &T = S.tab;
    if (n == 0) { T.lerp_count = 0; return BF_OK; }
    if ((rc = T.lerp_whole.ensure(n * sizeof(int)))) return rc;
    if ((rc = T.lerp_weight.ensure(n * sizeof(float)))) return rc;
    const float *d_src = src;
    if (!src_on_device) {
        if ((rc = S.d_scratch.ensure(n * sizeof(float) + 256))) return rc;
        BF_CUDA(cudaMemcpy(S.d_scratch.p, src, n * sizeof(float), cudaMemcpyHostToDevice));
        d_src = S.d_scratch.as<float>();
    }
    if ((rc = split_lerp_dev(d_src, n, T.lerp_whole.as<int>(), T.lerp_weight.as<float>(), 0))) return rc;
    BF_CUDA(cudaDeviceSynchronize());
    T.lerp_count = n;
    if ((rc = launch_max_abs_i32(T.lerp_whole.as<int>(), n, &T.lerp_max))) return rc;
    T.g_lerp.n = -1;
    T.version++;
    return BF_OK;
}

static int install_fir(const float *src, size_t n, bool src_on_device)
{
    State &S = state();
    int rc = ensure_device();
    if (rc) return rc;
    Tables &T = S.tab;
    if (n == 0) { T.fir_count = 0; return BF_OK; }
    if ((rc = T.fir_taps.ensure(n * sizeof(float)))) return rc;
    BF_CUDA(cudaMemcpy(T.fir_taps.p, src, n * sizeof(float),
                       src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    T.fir_count = n;
    T.version++;
    return BF_OK;
}

static int install_hybrid(const float *src, size_t n, bool src_on_device)
{
    State &S = state();
    int rc = ensure_device();
    if (rc) return rc;
    Tables &T = S.tab;
    const int taps = S.cfg.n_taps;
    if (n == 0) { T.hyb_count = 0; return BF_OK; }
    std::vector<float> h_d;
    const float *h_src = src;
    if (src_on_device) {
        h_d.resize(n);
        BF_CUDA(cudaMemcpy(h_d.data(), src, n * sizeof(float), cudaMemcpyDeviceToHost));
        h_src = h_d.data();
    }
    std::vector<int> whole(n);
    std::vector<float> tp(n * (size_t)taps);
    split_hybrid_host(h_src, n, whole.data(), tp.data(), taps);
    if ((rc = T.hyb_whole.ensure(n * sizeof(int)))) return rc;
    if ((rc = T.hyb_taps.ensure(tp.size() * sizeof(float)))) return rc;
    BF_CUDA(cudaMemcpy(T.hyb_whole.p, whole.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    BF_CUDA(cudaMemcpy(T.hyb_taps.p, tp.data(), tp.size() * sizeof(float), cudaMemcpyHostToDevice));
    T.hyb_count = n;
    T.version++;
    return BF_OK;
}

// ---- host-pointer MIMO / MISO ----------------------------------------------------
static int mimo_dispatch(int algo, const float *d_sig, float *d_img, int frames, const int *d_mics,
                         int n, int d_begin, int d_count, ImgLayout lay, cudaStream_t st)
{
    State &S = state();
    const int N = S.cfg.n_samples;
    if (lay.frame_stride == 0 && lay.dir_stride == 0)
        lay = ImgLayout{(long)S.cfg.max_res_x * S.cfg.max_res_y, 1, 0};
    const bool tiled_ok = (algo == BF_ALGO_PAD || algo == BF_ALGO_LERP || algo == -1) &&
                          !S.simple_kernel && (N == 64 || N == 128 || N == 256);
    if (tiled_ok) return mimo_tiled(algo, d_sig, d_img, frames, d_mics, n, d_begin, d_count, lay, st);
    if ((algo == BF_ALGO_FIR_SEQ || algo == BF_ALGO_FIR_LANES) && !S.simple_kernel &&
        fir_tiled_supported(algo, N, S.cfg.n_taps))
        return fir_tiled(algo, d_sig, d_img, frames, d_mics, n, d_begin, d_count, lay, st);
    return mimo_simple(algo, d_sig, d_img, frames, d_mics, n, d_begin, d_count, lay, st);
}

static int check_n(int n, const char *who)
{
    State &S = state();
    if (n <= 0 || n > S.cfg.n_microphones) {
        set_error(BF_ERR_ARG, "%s: n = %d outside (0, N_MICROPHONES = %d]", who, n, S.cfg.n_microphones);
        return BF_ERR_ARG;
    }
    return BF_OK;
}

static int upload_mics(const int *adaptive_array, int n)
{
    State &S = state();
    for (int m = 0; m < n; m++)
        if (adaptive_array[m] < 0 || adaptive_array[m] >= S.cfg.n_microphones) {
            set_error(BF_ERR_ARG, "adaptive_array[%d] = %d outside [0, N_MICROPHONES)", m, adaptive_array[m]);
            return BF_ERR_ARG;
        }
    int rc = S.d_mic_ids.ensure((size_t)n * sizeof(int));
    if (rc) return rc;
    BF_CUDA(cudaMemcpyAsync(S.d_mic_ids.p, adaptive_array, (size_t)n * sizeof(int),
                            cudaMemcpyHostToDevice, 0));
    return BF_OK;
}

static int host_mimo(int algo, const float *signals, float *image, const int *adaptive_array, int n)
{
    std::lock_guard<std::mutex> lk(g_mu);
    clear_error();
    State &S = state();
    int rc = ensure_device();
    if (rc) return rc;
    if ((rc = order_behind_previous(0))) return rc;
    if ((rc = check_n(n, "mimo"))) return rc;
    const size_t sig_b = (size_t)S.cfg.n_microphones * S.cfg.n_samples * sizeof(float);
    const int D = S.cfg.max_res_x * S.cfg.max_res_y;
    const size_t img_b = (size_t)D * sizeof(float);
    if ((rc = ensure_pinned(sig_b + img_b))) return rc;
    if ((rc = S.d_signals.ensure(sig_b))) return rc;
    if ((rc = S.d_image.ensure(img_b))) return rc;
    if ((rc = upload_mics(adaptive_array, n))) return rc;
    // pageable -> pinned -> device keeps the H2D copy asynchronous and at full PCIe rate
    memcpy(S.h_pinned, signals, sig_b);
    BF_CUDA(cudaMemcpyAsync(S.d_signals.p, S.h_pinned, sig_b, cudaMemcpyHostToDevice, 0));
    if ((rc = mimo_dispatch(algo, S.d_signals.as<float>(), S.d_image.as<float>(), 1,
                            S.d_mic_ids.as<int>(), n, 0, D, ImgLayout{0, 0, 0}, 0)))
        return rc;
    float *h_img = (float *)((char *)S.h_pinned + sig_b);
    BF_CUDA(cudaMemcpyAsync(h_img, S.d_image.p, img_b, cudaMemcpyDeviceToHost, 0));
    BF_CUDA(cudaStreamSynchronize(0));
    memcpy(image, h_img, img_b);
    return BF_OK;
}

static int host_miso(int algo, const float *signals, float *out, const int *adaptive_array, int n,
                     int offset, int by_mic)
{
    std::lock_guard<std::mutex> lk(g_mu);
    clear_error();
    State &S = state();
    int rc = ensure_device();
    if (rc) return rc;
    if ((rc = order_behind_previous(0))) return rc;
    if ((rc = check_n(n, "miso"))) return rc;
    const size_t sig_b = (size_t)S.cfg.n_microphones * S.cfg.n_samples * sizeof(float);
    const size_t out_b = (size_t)S.cfg.n_samples * sizeof(float);
    if ((rc = ensure_pinned(sig_b + out_b))) return rc;
    if ((rc = S.d_signals.ensure(sig_b))) return rc;
    if ((rc = S.d_out.ensure(out_b))) return rc;
    if ((rc = upload_mics(adaptive_array, n))) return rc;
    memcpy(S.h_pinned, signals, sig_b);
    BF_CUDA(cudaMemcpyAsync(S.d_signals.p, S.h_pinned, sig_b, cudaMemcpyHostToDevice, 0));
    if ((rc = miso_run(algo, S.d_signals.as<float>(), S.d_out.as<float>(), 1, S.d_mic_ids.as<int>(), n,
                       offset, by_mic, 0, 0)))
        return rc;
    float *h_out = (float *)((char *)S.h_pinned + sig_b);
    BF_CUDA(cudaMemcpyAsync(h_out, S.d_out.p, out_b, cudaMemcpyDeviceToHost, 0));
    BF_CUDA(cudaStreamSynchronize(0));
    memcpy(out, h_out, out_b);
    return BF_OK;
}

static int sourced(float **sig_out)
{
    State &S = state();
    if (!S.source) {
        set_error(BF_ERR_ARG, "no data source registered: call bf_set_data_source() (get_data() "
                              "of the reference's receiver stays host C, api.c:830-859)");
        return BF_ERR_ARG;
    }
    static std::vector<float> buf;
    buf.resize((size_t)S.cfg.n_microphones * S.cfg.n_samples);
    S.source(buf.data());
    *sig_out = buf.data();
    return BF_OK;
}

}  // namespace bf

using namespace bf;

extern "C" {

// =========================== part 2: configuration ==============================
int bf_configure(const bf_config *cfg)
{
    clear_error();
    if (!cfg) { set_error(BF_ERR_ARG, "bf_configure(NULL)"); return BF_ERR_ARG; }
    if (cfg->n_microphones < 1 || cfg->n_samples < 1 || cfg->n_samples > 1024 || cfg->n_taps < 1 ||
        cfg->max_res_x < 1 || cfg->max_res_y < 1) {
        set_error(BF_ERR_CONFIG, "bad sizes: mics %d samples %d (1..1024) taps %d grid %dx%d",
                  cfg->n_microphones, cfg->n_samples, cfg->n_taps, cfg->max_res_x, cfg->max_res_y);
        return BF_ERR_CONFIG;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    State &S = state();
    S.cfg = *cfg;
    S.tab.g_pad.n = S.tab.g_lerp.n = S.tab.g_trunc.n = -1;
    return BF_OK;
}
int bf_get_config(bf_config *cfg)
{
    if (!cfg) return BF_ERR_ARG;
    *cfg = state().cfg;
    return BF_OK;
}
const char *bf_last_error(void) { return g_err; }
int bf_last_status(void) { return g_status; }
int bf_device_count(void)
{
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) return 0;
    return c;
}
int bf_set_device(int ordinal)
{
    clear_error();
    State &S = state();
    if (S.sm_count > 0 && S.device != ordinal) {
        set_error(BF_ERR_CONFIG, "device already initialised as %d; one device per process", S.device);
        return BF_ERR_CONFIG;
    }
    S.device = ordinal;
    return ensure_device();
}
const char *bf_version(void) { return "bf_b200 0.1 (sm_100a)"; }
void bf_set_data_source(bf_data_source_fn fn) { state().source = fn; }
int bf_set_kernel_options(int simple_kernel, int exact_sum)
{
    State &S = state();
    if (simple_kernel >= 0) S.simple_kernel = simple_kernel;
    if (exact_sum >= 0) S.exact_sum = exact_sum;
    return BF_OK;
}
uint64_t bf_kernel_launches(int reset)
{
    uint64_t v = g_launches.load();
    if (reset) g_launches = 0;
    return v;
}

// =========================== part 2: device entry points ========================
int bf_mimo_dev(int algo, const float *d_signals, float *d_images, int frames, const int *d_mic_ids,
                int n, int d_begin, int d_count, void *stream)
{
    return bf_mimo_dev_ex(algo, d_signals, d_images, frames, d_mic_ids, n, d_begin, d_count, 0, 0, 0,
                          stream);
}

int bf_mimo_dev_ex(int algo, const float *d_signals, float *d_images, int frames,
                   const int *d_mic_ids, int n, int d_begin, int d_count, long frame_stride,
                   long dir_stride, int d_origin, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    clear_error();
    State &S = state();
    int rc = ensure_device();
    if (rc) return rc;
    if ((rc = order_behind_previous((cudaStream_t)stream))) return rc;
    if ((rc = check_n(n, "bf_mimo_dev"))) return rc;
    const int D = S.cfg.max_res_x * S.cfg.max_res_y;
    if (frames < 1 || d_begin < 0 || d_count < 1 || d_begin + d_count > D || !d_signals || !d_images || !d_mic_ids) {
        set_error(BF_ERR_ARG, "bf_mimo_dev: frames %d slice [%d,+%d) of D=%d", frames, d_begin, d_count, D);
        return BF_ERR_ARG;
    }
    return mimo_dispatch(algo, d_signals, d_images, frames, d_mic_ids, n, d_begin, d_count,
                         ImgLayout{frame_stride, dir_stride, d_origin}, (cudaStream_t)stream);
}

// Fused power maps + all-gather: this rank's direction slice, stored into the gather buffer of
// every rank (layout per buffer: float [world][frames][per_rank]; slice r belongs to rank r).
int bf_mimo_dev_gather(int algo, const float *d_signals, int frames, const int *d_mic_ids, int n,
                       int d_begin, int d_count, int rank, int world, void *const *gather_bufs,
                       long per_rank, void *stream)
{
    return bf_mimo_dev_gather_sync(algo, d_signals, frames, d_mic_ids, n, d_begin, d_count, rank, world, gather_bufs,
                                   per_rank, nullptr, 0, 0, nullptr, stream);
}

int bf_mimo_dev_gather_sync(int algo, const float *d_signals, int frames, const int *d_mic_ids, int n,
                            int d_begin, int d_count, int rank, int world, void *const *gather_bufs,
                            long per_rank, void *const *flag_arrays, long long wait_seq, long long signal_seq,
                            int *d_timed_out, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    clear_error();
    State &S = state();
    int rc = ensure_device();
    if (rc) return rc;
    if ((rc = order_behind_previous((cudaStream_t)stream))) return rc;
    if ((rc = check_n(n, "bf_mimo_dev_gather"))) return rc;
    const int D = S.cfg.max_res_x * S.cfg.max_res_y;
    const int N = S.cfg.n_samples;
    if (frames < 1 || d_begin < 0 || d_count < 0 || d_begin + d_count > D || !d_signals || !d_mic_ids ||
        !gather_bufs || world < 1 || world > 8 || rank < 0 || rank >= world || per_rank < d_count) {
        set_error(BF_ERR_ARG, "bf_mimo_dev_gather: frames %d slice [%d,+%d) of D=%d rank %d/%d per_rank %ld", frames,
                  d_begin, d_count, D, rank, world, per_rank);
        return BF_ERR_ARG;
    }
    if (!((algo == BF_ALGO_PAD || algo == BF_ALGO_LERP) && !S.simple_kernel && (N == 64 || N == 128 || N == 256))) {
        set_error(BF_ERR_CONFIG, "bf_mimo_dev_gather: only the tiled pad / lerp kernels store to peers");
        return BF_ERR_CONFIG;
    }
    if ((wait_seq > 0 || signal_seq > 0) && (!flag_arrays || !d_timed_out)) {
        set_error(BF_ERR_ARG, "bf_mimo_dev_gather_sync: flags requested without flag arrays");
        return BF_ERR_ARG;
    }
    if (d_count == 0) {
        // nothing to compute on this rank, but its peers still wait for its flag
        if (signal_seq > 0) return bf_gather_signal(flag_arrays, world, rank, signal_seq, stream);
        return BF_OK;
    }
    ImgLayout lay{};
    if (wait_seq > 0 || signal_seq > 0) {
        if ((rc = S.d_done_counter.ensure(4 * sizeof(unsigned int)))) return rc;
        static bool zeroed = false;
        if (!zeroed) { BF_CUDA(cudaMemset(S.d_done_counter.p, 0, 4 * sizeof(unsigned int))); zeroed = true; }
        lay.flags_local = (long long *)flag_arrays[rank];
        lay.wait_seq = wait_seq;
        lay.signal_seq = signal_seq;
        lay.world = world;
        lay.flag_rank = rank;
        // two steps can be in flight when they overlap: the ticket counter rotates with the step number
        lay.done_counter = S.d_done_counter.as<unsigned int>() + (signal_seq & 3);
        lay.overlap = (S.gather_overlap && algo == BF_ALGO_PAD && signal_seq > 0) ? 1 : 0;
        lay.timed_out = d_timed_out;
        for (int r = 0; r < world; r++) lay.flags_all[r] = (long long *)flag_arrays[r];
    }
    lay.frame_stride = per_rank;
    lay.dir_stride = 1;
    lay.d_origin = d_begin;
    const size_t slice = (size_t)rank * frames * per_rank;
    float *own = nullptr;
    for (int r = 0; r < world; r++) {
        if (!gather_bufs[r]) { set_error(BF_ERR_ARG, "bf_mimo_dev_gather: null buffer for rank %d", r); return BF_ERR_ARG; }
        float *dst = (float *)gather_bufs[r] + slice;
        if (r == rank) own = dst;
        else lay.peers[lay.n_peers++] = dst;
    }
    return mimo_tiled(algo, d_signals, own, frames, d_mic_ids, n, d_begin, d_count, lay, (cudaStream_t)stream);
}

int bf_gather_overlap(int on)
{
    std::lock_guard<std::mutex> lk(g_mu);
    State &S = state();
    const int was = S.gather_overlap;
    if (on >= 0) S.gather_overlap = on ? 1 : 0;
    return was;
}

int bf_miso_dev(int algo, const float *d_signals, float *d_out, int blocks, const int *d_mic_ids,
                int n, int offset, int scale, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    clear_error();
    int rc = ensure_device();
    if (rc) return rc;
    if ((rc = order_behind_previous((cudaStream_t)stream))) return rc;
    if ((rc = check_n(n, "bf_miso_dev"))) return rc;
    if (blocks < 1 || offset < 0 || !d_signals || !d_out || !d_mic_ids) {
        set_error(BF_ERR_ARG, "bf_miso_dev: blocks %d offset %d", blocks, offset);
        return BF_ERR_ARG;
    }
    return miso_run(algo, d_signals, d_out, blocks, d_mic_ids, n, offset, 0, scale, (cudaStream_t)stream);
}

// Chunk schedule of bf_mimo_host_batch: ramp up and down.  What the call cannot hide is the H2D copy of the FIRST
// chunk (nothing to run yet) and the D2H copy of the LAST one (nothing left to run): 4-frame chunks at both ends
// (1 MB in, 0.5 MB out at C3: ~70 us together instead of ~400 us for 16 / 32 frames), 12-frame chunks next to them
// on long batches, chunks of up to 32 frames in between.  Short launches are not inefficient: the kernel hands out
// equal unit ranges (TileWalk), so a 4-frame launch has no ragged last round.
static void host_batch_schedule(int frames, std::vector<int> &c_start, std::vector<int> &c_size)
{
    std::vector<int> head, tail;
    if (frames > 8 && frames < 64) { head = {4}; tail = {4}; }
    else if (frames >= 64) { head = {4, 12}; tail = {12, 4}; }
    int used = 0;
    for (int v : head) used += v;
    for (int v : tail) used += v;
    const int mid = frames - used;
    const int nmid = (mid + 31) / 32;
    int f = 0;
    auto push = [&](int take) { if (take > 0) { c_start.push_back(f); c_size.push_back(take); f += take; } };
    for (int v : head) push(v);
    for (int i = 0; i < nmid; i++) push(mid / nmid + (i < mid % nmid ? 1 : 0));
    for (int v : tail) push(v);
}

int bf_host_batch_schedule(int frames, int *chunk_frames, int capacity)
{
    if (frames < 1) return -1;
    std::vector<int> c_start, c_size;
    host_batch_schedule(frames, c_start, c_size);
    for (size_t i = 0; i < c_size.size() && (int)i < capacity; i++)
        if (chunk_frames) chunk_frames[i] = c_size[i];
    return (int)c_size.size();
}

int bf_mimo_host_batch(int algo, const float *signals, float *images, int frames,
                       const int *adaptive_array, int n)
{
    std::lock_guard<std::mutex> lk(g_mu);
    clear_error();
    State &S = state();
    int rc = ensure_device();
    if (rc) return rc;
    if ((rc = check_n(n, "bf_mimo_host_batch"))) return rc;
    if (frames < 1 || !signals || !images || !adaptive_array) {
        set_error(BF_ERR_ARG, "bf_mimo_host_batch: frames %d", frames);
        return BF_ERR_ARG;
    }
    const size_t sig_f = (size_t)S.cfg.n_microphones * S.cfg.n_samples;     // floats per frame
    const int D = S.cfg.max_res_x * S.cfg.max_res_y;
    std::vector<int> c_start, c_size;
    host_batch_schedule(frames, c_start, c_size);
    int chunk = 0;
    for (int v : c_size) chunk = v > chunk ? v : chunk;
    static cudaStream_t st_in = nullptr, st_run = nullptr, st_out = nullptr;
    static cudaEvent_t ev_in[2], ev_run[2], ev_out[2];
    static DevBuf d_sig[2], d_img[2];
    if (!st_in) {
        BF_CUDA(cudaStreamCreateWithFlags(&st_in, cudaStreamNonBlocking));
        BF_CUDA(cudaStreamCreateWithFlags(&st_run, cudaStreamNonBlocking));
        BF_CUDA(cudaStreamCreateWithFlags(&st_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            BF_CUDA(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
            BF_CUDA(cudaEventCreateWithFlags(&ev_run[i], cudaEventDisableTiming));
            BF_CUDA(cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming));
        }
    }
    auto pinned = [](const void *p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;
    };
    const bool in_pinned = pinned(signals), out_pinned = pinned(images);
    const size_t stage_in = in_pinned ? 0 : (size_t)chunk * sig_f * sizeof(float);
    const size_t stage_out = out_pinned ? 0 : (size_t)chunk * D * sizeof(float);
    if ((rc = ensure_pinned(2 * (stage_in + stage_out) + 256))) return rc;
    char *h_in[2] = {(char *)S.h_pinned, (char *)S.h_pinned + stage_in};
    char *h_out[2] = {(char *)S.h_pinned + 2 * stage_in, (char *)S.h_pinned + 2 * stage_in + stage_out};
    for (int i = 0; i < 2; i++) {
        if ((rc = d_sig[i].ensure((size_t)chunk * sig_f * sizeof(float)))) return rc;
        if ((rc = d_img[i].ensure((size_t)chunk * D * sizeof(float)))) return rc;
    }
    if ((rc = order_behind_previous(0))) return rc;
    if ((rc = upload_mics(adaptive_array, n))) return rc;
    BF_CUDA(cudaStreamSynchronize(0));
    if ((rc = order_behind_previous(st_run))) return rc;
    const int nchunks = (int)c_size.size();
    auto drain = [&](int k) -> int {          // wait for chunk k's results, un-stage them
        const int sl = k & 1, f0 = c_start[k], fc = c_size[k];
        BF_CUDA(cudaEventSynchronize(ev_out[sl]));
        if (!out_pinned) memcpy(images + (size_t)f0 * D, h_out[sl], (size_t)fc * D * sizeof(float));
        return BF_OK;
    };
    for (int k = 0; k < nchunks; k++) {
        const int sl = k & 1, f0 = c_start[k], fc = c_size[k];
        if (k >= 2 && (rc = drain(k - 2))) return rc;                  // slot sl is free again
        const float *src = signals + (size_t)f0 * sig_f;
        if (!in_pinned) { memcpy(h_in[sl], src, (size_t)fc * sig_f * sizeof(float)); src = (const float *)h_in[sl]; }
        BF_CUDA(cudaMemcpyAsync(d_sig[sl].p, src, (size_t)fc * sig_f * sizeof(float), cudaMemcpyHostToDevice, st_in));
        BF_CUDA(cudaEventRecord(ev_in[sl], st_in));
        BF_CUDA(cudaStreamWaitEvent(st_run, ev_in[sl], 0));
        if ((rc = mimo_dispatch(algo, d_sig[sl].as<float>(), d_img[sl].as<float>(), fc, S.d_mic_ids.as<int>(), n,
                                0, D, ImgLayout{0, 0, 0}, st_run)))
            return rc;
        BF_CUDA(cudaEventRecord(ev_run[sl], st_run));
        BF_CUDA(cudaStreamWaitEvent(st_out, ev_run[sl], 0));
        float *dst = out_pinned ? images + (size_t)f0 * D : (float *)h_out[sl];
        BF_CUDA(cudaMemcpyAsync(dst, d_img[sl].p, (size_t)fc * D * sizeof(float), cudaMemcpyDeviceToHost, st_out));
        BF_CUDA(cudaEventRecord(ev_out[sl], st_out));
        // the next H2D into the other slot must not overtake the kernel still reading it
        BF_CUDA(cudaStreamWaitEvent(st_in, ev_run[sl ^ 1], 0));
    }
    for (int k = nchunks > 2 ? nchunks - 2 : 0; k < nchunks; k++)
        if ((rc = drain(k))) return rc;
    return BF_OK;
}

int bf_load_table_dev(int algo, const void *d_table, size_t count)
{
    std::lock_guard<std::mutex> lk(g_mu);
    clear_error();
    Tables &T = state().tab;
    switch (algo) {
        case BF_ALGO_PAD: return install_pad(T.pad_whole, T.pad_count, T.pad_max, &T.g_pad, (const int *)d_table, count, true);
        case BF_ALGO_LERP: return install_lerp((const float *)d_table, count, true);
        case BF_ALGO_FIR_SEQ: case BF_ALGO_FIR_LANES: return install_fir((const float *)d_table, count, true);
        case BF_ALGO_HYBRID: return install_hybrid((const float *)d_table, count, true);
    }
    set_error(BF_ERR_ARG, "bf_load_table_dev: bad algo %d", algo);
    return BF_ERR_ARG;
}

int bf_generate_delays(double k, const double *x_scan, int res_x, const double *y_scan, int res_y,
                       double z2, const double *mic_x, const double *mic_y, int n,
                       double *delays_f64, int *whole_i32, float *delays_f32, int load_algo)
{
    std::lock_guard<std::mutex> lk(g_mu);
    clear_error();
    State &S = state();
    int rc = ensure_device();
    if (rc) return rc;
    if (res_x < 1 || res_y < 1 || n < 1 || !x_scan || !y_scan || !mic_x || !mic_y) {
        set_error(BF_ERR_ARG, "bf_generate_delays: bad arguments");
        return BF_ERR_ARG;
    }
    const size_t cnt = (size_t)res_x * res_y * n;
    DevBuf d_in, d_f64, d_i32, d_f32;
    const size_t in_cnt = (size_t)res_x + res_y + 2 * (size_t)n;
    if ((rc = d_in.ensure(in_cnt * sizeof(double)))) return rc;
    double *d_xs = d_in.as<double>(), *d_ys = d_xs + res_x, *d_mx = d_ys + res_y, *d_my = d_mx + n;
    auto fail = [&](int r) { d_in.release(); d_f64.release(); d_i32.release(); d_f32.release(); return r; };
#define BF_TRY(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { set_error(BF_ERR_CUDA, "%s -> %s", #x, cudaGetErrorString(_e)); return fail(BF_ERR_CUDA); } } while (0)
    BF_TRY(cudaMemcpy(d_xs, x_scan, res_x * sizeof(double), cudaMemcpyHostToDevice));
    BF_TRY(cudaMemcpy(d_ys, y_scan, res_y * sizeof(double), cudaMemcpyHostToDevice));
    BF_TRY(cudaMemcpy(d_mx, mic_x, n * sizeof(double), cudaMemcpyHostToDevice));
    BF_TRY(cudaMemcpy(d_my, mic_y, n * sizeof(double), cudaMemcpyHostToDevice));
    const bool want_i32 = whole_i32 || load_algo == BF_ALGO_PAD;
    const bool want_f32 = delays_f32 || load_algo == BF_ALGO_LERP || load_algo == BF_ALGO_HYBRID;
    if (delays_f64 && (rc = d_f64.ensure(cnt * sizeof(double)))) return fail(rc);
    if (want_i32 && (rc = d_i32.ensure(cnt * sizeof(int)))) return fail(rc);
    if (want_f32 && (rc = d_f32.ensure(cnt * sizeof(float)))) return fail(rc);
    if ((rc = delay_table_dev(k, d_xs, res_x, d_ys, res_y, z2, d_mx, d_my, n, d_f64.as<double>(),
                              d_i32.as<int>(), d_f32.as<float>(), 0)))
        return fail(rc);
    BF_TRY(cudaDeviceSynchronize());
    if (delays_f64) BF_TRY(cudaMemcpy(delays_f64, d_f64.p, cnt * sizeof(double), cudaMemcpyDeviceToHost));
    if (whole_i32) BF_TRY(cudaMemcpy(whole_i32, d_i32.p, cnt * sizeof(int), cudaMemcpyDeviceToHost));
    if (delays_f32) BF_TRY(cudaMemcpy(delays_f32, d_f32.p, cnt * sizeof(float), cudaMemcpyDeviceToHost));
#undef BF_TRY
    Tables &T = S.tab;
    if (load_algo == BF_ALGO_PAD)
        rc = install_pad(T.pad_whole, T.pad_count, T.pad_max, &T.g_pad, d_i32.as<int>(), cnt, true);
    else if (load_algo == BF_ALGO_LERP)
        rc = install_lerp(d_f32.as<float>(), cnt, true);
    else if (load_algo == BF_ALGO_HYBRID)
        rc = install_hybrid(d_f32.as<float>(), cnt, true);
    else if (load_algo >= 0) { set_error(BF_ERR_ARG, "bf_generate_delays: load_algo %d", load_algo); rc = BF_ERR_ARG; }
    return fail(rc);
}

int bf_get_lerp_tables(int *whole, float *weight, size_t count)
{
    clear_error();
    Tables &T = state().tab;
    if (count > T.lerp_count) { set_error(BF_ERR_NOT_LOADED, "lerp table holds %zu entries", T.lerp_count); return BF_ERR_NOT_LOADED; }
    if (whole) BF_CUDA(cudaMemcpy(whole, T.lerp_whole.p, count * sizeof(int), cudaMemcpyDeviceToHost));
    if (weight) BF_CUDA(cudaMemcpy(weight, T.lerp_weight.p, count * sizeof(float), cudaMemcpyDeviceToHost));
    return BF_OK;
}
int bf_get_hybrid_tables(int *whole, float *taps, size_t count)
{
    clear_error();
    State &S = state();
    Tables &T = S.tab;
    if (count > T.hyb_count) { set_error(BF_ERR_NOT_LOADED, "hybrid table holds %zu entries", T.hyb_count); return BF_ERR_NOT_LOADED; }
    if (whole) BF_CUDA(cudaMemcpy(whole, T.hyb_whole.p, count * sizeof(int), cudaMemcpyDeviceToHost));
    if (taps) BF_CUDA(cudaMemcpy(taps, T.hyb_taps.p, count * S.cfg.n_taps * sizeof(float), cudaMemcpyDeviceToHost));
    return BF_OK;
}

// =========================== part 1: drop-in names ==============================
// ---- pad_and_sum.h ----
void load_coefficients_pad(int *whole_samples, int n)
{
    std::lock_guard<std::mutex> lk(g_mu);
    clear_error();
    Tables &T = state().tab;
    install_pad(T.pad_whole, T.pad_count, T.pad_max, &T.g_pad, whole_samples, n < 0 ? 0 : (size_t)n, false);
}
void load_coefficients_pad2(int *whole_miso, int n)
{
    std::lock_guard<std::mutex> lk(g_mu);
    clear_error();
    Tables &T = state().tab;
    int dummy;
    install_pad(T.pad2_whole, T.pad2_count, dummy, nullptr, whole_miso, n < 0 ? 0 : (size_t)n, false);
}
void unload_coefficients_pad(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Tables &T = state().tab;
    T.pad_whole.release(); T.pad_count = 0; T.g_pad.offs.release(); T.g_pad.n = -1;
}
void unload_coefficients_pad2(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Tables &T = state().tab;
    T.pad2_whole.release(); T.pad2_count = 0;
}
void pad_delay(float *signal, float *out, int pos_pad)
{
    std::lock_guard<std::mutex> lk(g_mu);
    clear_error();
    single_delay(0, signal, nullptr, 0.0f, pos_pad, out);
}
void miso_pad(float *signals, float *out, int *adaptive_array, int n, int offset)
{
    host_miso(BF_ALGO_PAD, signals, out, adaptive_array, n, offset, 0);
}
void miso_pad2(float *signals, float *out, int *adaptive_array, int n, int offset)
{
    (void)offset;   // unused by the reference too (pad_and_sum.c:77-92)
    host_miso(BF_ALGO_PAD, signals, out, adaptive_array, n, 0, 1);
}
void mimo_pad(float *signals, float *image, int *adaptive_array, int n)
{
    host_mimo(BF_ALGO_PAD, signals, image, adaptive_array, n);
}

// ---- lerp_and_sum.h ----
void load_coefficients_lerp(float *delays, int n)
{
    std::lock_guard<std::mutex> lk(g_mu);
    clear_error();
    install_lerp(delays, n < 0 ? 0 : (size_t)n, false);
}
void unload_coefficients_lerp(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Tables &T = state().tab;
    T.lerp_whole.release(); T.lerp_weight.release(); T.lerp_count = 0;
    T.g_lerp.offs.release(); T.g_lerp.wts.release(); T.g_lerp.n = -1;
}
void lerp_delay(float *signal, float *out, float h, int pad)
{
    std::lock_guard<std::mutex> lk(g_mu);
    clear_error();
    single_delay(1, signal, nullptr, h, pad, out);
}
void miso_lerp(float *signals, float *out, int *adaptive_array, int n, int offset)
{
    host_miso(BF_ALGO_LERP, signals, out, adaptive_array, n, offset, 0);
}
void mimo_lerp(float *signals, float *image, int *adaptive_array, int n)
{
    host_mimo(BF_ALGO_LERP, signals, image, adaptive_array, n);
}

// ---- convolve_and_sum.h ----
void load_coefficients_convolve(float *h, int n)
{
    std::lock_guard<std::mutex> lk(g_mu);
    clear_error();
    install_fir(h, n < 0 ? 0 : (size_t)n, false);
}
void unload_coefficients_convolve(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Tables &T = state().tab;
    T.fir_taps.release(); T.fir_count = 0;
}
void convolve_delay_naive_add(float *signal, float *h, float *out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    clear_error();
    single_delay(2, signal, h, 0.0f, 0, out);
}
void convolve_delay_naive(float *signal, float *out, float *h)
{
    std::lock_guard<std::mutex> lk(g_mu);
    clear_error();
    single_delay(2, signal, h, 0.0f, 0, out);
}
void convolve_delay_vectorized(float *signal, float *h, float *out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    clear_error();
    single_delay(4, signal, h, 0.0f, 0, out);
}
void convolve_delay_vectorized_add(float *signal, float *h, float *out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    clear_error();
    single_delay(3, signal, h, 0.0f, 0, out);
}
void miso_convolve_naive(float *signals, float *out, int *adaptive_array, int n, int offset)
{
    host_miso(BF_ALGO_FIR_SEQ, signals, out, adaptive_array, n, offset, 0);
}
void miso_convolve_vectorized(float *signals, float *out, int *adaptive_array, int n, int offset)
{
    host_miso(BF_ALGO_FIR_LANES, signals, out, adaptive_array, n, offset, 0);
}
void mimo_convolve_naive(float *signals, float *image, int *adaptive_array, int n)
{
    host_mimo(BF_ALGO_FIR_SEQ, signals, image, adaptive_array, n);
}
void mimo_convolve_vectorized(float *signals, float *image, int *adaptive_array, int n)
{
    host_mimo(BF_ALGO_FIR_LANES, signals, image, adaptive_array, n);
}

// ---- hybrid_convolve_and_sum.h ----
void load_coefficients_convolve_hybrid(float *h, int n)
{
    std::lock_guard<std::mutex> lk(g_mu);
    clear_error();
    install_hybrid(h, n < 0 ? 0 : (size_t)n, false);
}
void unload_coefficients_convolve_hybrid(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Tables &T = state().tab;
    T.hyb_whole.release(); T.hyb_taps.release(); T.hyb_count = 0;
}
void convolve_hybrid_delay_add(float *signal, float *h, int pad, float *out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    clear_error();
    single_delay(5, signal, h, 0.0f, pad, out);
}
void miso_convolve_hybrid(float *signals, float *out, int *adaptive_array, int n, int offset)
{
    host_miso(BF_ALGO_HYBRID, signals, out, adaptive_array, n, offset, 0);
}
void mimo_convolve_hybrid(float *signals, float *image, int *adaptive_array, int n)
{
    host_mimo(BF_ALGO_HYBRID, signals, image, adaptive_array, n);
}

// ---- api.h wrappers: buffer from the registered source, then the kernel ----
void pad_mimo(float *image, int *adaptive_array, int n)
{
    float *sig;
    if (sourced(&sig) == BF_OK) host_mimo(BF_ALGO_PAD, sig, image, adaptive_array, n);
}
void lerp_mimo(float *image, int *adaptive_array, int n)
{
    float *sig;
    if (sourced(&sig) == BF_OK) host_mimo(BF_ALGO_LERP, sig, image, adaptive_array, n);
}
void convolve_mimo_naive(float *image, int *adaptive_array, int n)
{
    float *sig;
    if (sourced(&sig) == BF_OK) host_mimo(BF_ALGO_FIR_SEQ, sig, image, adaptive_array, n);
}
void convolve_mimo_vectorized(float *image, int *adaptive_array, int n)
{
    float *sig;
    if (sourced(&sig) == BF_OK) host_mimo(BF_ALGO_FIR_LANES, sig, image, adaptive_array, n);
}
void load_coefficients2(int *whole_samples, int n)
{
    std::lock_guard<std::mutex> lk(g_mu);
    clear_error();
    Tables &T = state().tab;
    install_pad(T.trunc_whole, T.trunc_count, T.trunc_max, &T.g_trunc, whole_samples, n < 0 ? 0 : (size_t)n, false);
}
void mimo_truncated(float *image, int *adaptive_array, int n)
{
    float *sig;
    if (sourced(&sig) == BF_OK) host_mimo(-1, sig, image, adaptive_array, n);
}
void miso_steer_listen(float *out, int *adaptive_array, int n, int steer_offset)
{
    float *sig;
    if (sourced(&sig) == BF_OK) host_miso(BF_ALGO_PAD, sig, out, adaptive_array, n, steer_offset, 0);
}

// ---- the reference's shared-memory records, with run-time sizes (api.h:26-38, receiver.h:31-36) ----
// All members are 4-byte types in declaration order, so the layouts are plain running sums.
bf_record_layout bf_layout_miso(int n_microphones, int n_samples)
{
    bf_record_layout l{};
    l.off[0] = 0;                                                   // int steer_offset
    l.off[1] = 4;                                                   // float signals[BUFFER_LENGTH]
    l.off[2] = l.off[1] + (size_t)n_microphones * n_samples * 4;    // int adaptive_array[N_MICROPHONES]
    l.off[3] = l.off[2] + (size_t)n_microphones * 4;                // int n
    l.size = l.off[3] + 4;
    return l;
}
bf_record_layout bf_layout_padata(int n_samples)
{
    bf_record_layout l{};
    l.off[0] = 0;                                                   // int can_read
    l.off[1] = 4;                                                   // float out[N_SAMPLES]
    l.size = 4 + (size_t)n_samples * 4;
    return l;
}
bf_record_layout bf_layout_ring_buffer(int n_microphones, int n_samples)
{
    bf_record_layout l{};
    l.off[0] = 0;                                                   // int index
    l.off[1] = 4;                                                   // float data[BUFFER_LENGTH * 4]
    l.off[2] = 4 + (size_t)n_microphones * n_samples * 4 * 4;       // int counter
    l.size = l.off[2] + 4;
    return l;
}

// One iteration of the reference's audio child (api.c:505-529, miso_loop) on a `Miso` record in host (shared)
// memory: miso_pad(miso->signals, out, miso->adaptive_array, miso->n, miso->steer_offset), then, when `pa` is
// not NULL, the post-scale out[i] / n * MIC_GAIN into pa->out and pa->can_read = 1.
int bf_miso_record_listen(const void *miso_record, void *padata_record)
{
    State &S = state();
    const bf_record_layout lm = bf_layout_miso(S.cfg.n_microphones, S.cfg.n_samples);
    if (!miso_record) { set_error(BF_ERR_ARG, "bf_miso_record_listen: null record"); return BF_ERR_ARG; }
    const char *m = (const char *)miso_record;
    const int steer_offset = *(const int *)(m + lm.off[0]);
    const float *signals = (const float *)(m + lm.off[1]);
    const int *adaptive = (const int *)(m + lm.off[2]);
    const int n = *(const int *)(m + lm.off[3]);
    std::vector<float> out((size_t)S.cfg.n_samples);
    int rc = host_miso(BF_ALGO_PAD, signals, out.data(), adaptive, n, steer_offset, 0);
    if (rc) return rc;
    if (padata_record) {
        const bf_record_layout lp = bf_layout_padata(S.cfg.n_samples);
        char *pa = (char *)padata_record;
        float *dst = (float *)(pa + lp.off[1]);
        const float fn = (float)n, gain = S.cfg.mic_gain;
        for (int i = 0; i < S.cfg.n_samples; i++) dst[i] = out[i] / fn * gain;          // api.c:519-523
        *(int *)(pa + lp.off[0]) = 1;
    }
    return BF_OK;
}

}  // extern "C"
