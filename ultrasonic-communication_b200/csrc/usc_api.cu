// usc_api.cu — the C-ABI of libusc.so (include/usc.h).  Thin: argument checks, table ownership,
// launches.  No torch types, no exceptions across the boundary, no CPU fallback.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <vector>

#include "../../include/usc.h"
#include "usc_kernels.cuh"
#include "usc_launch.h"
#include "usc_tables.h"

using namespace usc;

static_assert(sizeof(usc_history) == sizeof(history_rec), "usc_history layout");
static_assert(sizeof(usc_rx_result) == sizeof(rx_result_rec), "usc_rx_result layout");

struct usc_handle {
    usc_config cfg;
    int device;
    int num_sms;
    cudaStream_t stream;
    uint32_t bandwidth, bandwidth2, idx_left_zero;
    std::vector<float> hann, up, down, H_up, H_down;
    float *d_hann, *d_up, *d_down, *d_ud, *d_H_up, *d_H_down;
    float2 *d_op_pass = nullptr, *d_op_split = nullptr;    // tables of the warp-level FFT operators when cfg.n != 2048
    float* d_rs_taps = nullptr; uint32_t rs_up = 0;       // resampler polyphase table (usc_resample_i16_to_pcm)
    float2 *d_G_up, *d_G_down, *d_tw_split4096;          // overlap-save synchroniser: template spectra (2n points), W_4096 split table
    float2 *d_tw_pass, *d_tw_split, *d_tw_l0;      // d_tw_l0: W_{n/2}^(a d), [d][a], 32768- and 65536-point frames only
    float* d_fir_coeffs;                              // 256 floats of scratch for usc_arm_fir_f32_batch
    // I/Q path (usc_iq_init): carrier tables, baseband chirp and its conjugate, half-length Hann, FIR taps
    std::vector<float> iq_cos, iq_sin, iq_chirp, iq_hann;
    float *d_iq_cos, *d_iq_sin, *d_iq_chirp, *d_iq_conj, *d_iq_hann, *d_iq_taps;
    uint32_t iq_ntaps, iq_window;
    int32_t* d_sym_table;                             // synthetic generator: up/down symbol tables (2n int32)
    double sym_amp;
    double sym_iq[4];                                 // carrier, bw, sideband, phase of the cached table; carrier 0 = chirp_orth symbols
    float* d_work;                                    // grow-on-demand scratch (large FFTs, generic demod)
    size_t work_bytes;
    std::map<uint32_t, float2*> tw_cache;             // master twiddle tables by length
    std::map<uint32_t, std::vector<float>> tw_host;
    uint64_t launches;
    // A/B and debugging switches, read ONCE in usc_create (USC_FFT_GENERIC, USC_LONG_UNFUSED, USC_IQ_UNFUSED):
    // the hot path never looks at the process environment
    bool dbg_fft_generic = false, dbg_long_unfused = false, dbg_iq_unfused = false;
    // host-buffer paths (usc_*_host): three chunk pipelines ("lanes"), each with its own stream, a device input
    // arena and a device result arena; sized by usc_host_workspace or on first use, grown on demand
    size_t lane_frames;                                // preferred chunk, in units of 8 KB (one 2048-sample frame)
    size_t lane_in_bytes, lane_out_bytes;
    cudaStream_t lane_stream[3];
    cudaEvent_t lane_done[3];                          // kernel of the lane's current chunk has finished (K7 ordering)
    void *lane_in[3], *lane_out[3];
    // K7 host path: per-stream carried state on the device, per-chunk pinned staging of uart bytes / results
    usc_rx_state* rxh_state; size_t rxh_state_n;
    uint8_t* rxh_stage; size_t rxh_stage_bytes;        // pinned host memory
};

static inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? USC_OK : USC_ERR_CUDA_BASE - (int) e; }
/* Every entry point makes the handle's device current for the duration of the call and puts the caller's
   device back on every return path (a process may hold handles on several GPUs, or run torch on another one). */
struct device_guard {
    int prev = -1;
    bool switched = false;
    explicit device_guard(int want) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != want) switched = cudaSetDevice(want) == cudaSuccess;
    }
    ~device_guard() {
        if (switched) cudaSetDevice(prev);
    }
    device_guard(const device_guard&) = delete;
    device_guard& operator=(const device_guard&) = delete;
};
#define USC_ENTER(h) device_guard guard__((h) ? (h)->device : -1)
#define CK(expr)                                   \
    do {                                           \
        cudaError_t e__ = (expr);                  \
        if (e__ != cudaSuccess) return cuda_rc(e__); \
    } while (0)

static bool pow2(uint32_t n) { return n && !(n & (n - 1)); }

static int upload(const void* host, size_t bytes, void** dev) {
    CK(cudaMalloc(dev, bytes));
    CK(cudaMemcpy(*dev, host, bytes, cudaMemcpyHostToDevice));
    return USC_OK;
}

// master twiddle table of `len` entries on the device (cached per handle)
static int get_twiddles(usc_handle* h, uint32_t len, float2** out) {
    auto it = h->tw_cache.find(len);
    if (it != h->tw_cache.end()) { *out = it->second; return USC_OK; }
    std::vector<float> tw(2 * (size_t) len);
    usc_host_twiddles(tw.data(), len);
    void* d = nullptr;
    int rc = upload(tw.data(), tw.size() * sizeof(float), &d);
    if (rc) return rc;
    h->tw_cache[len] = (float2*) d;
    h->tw_host[len] = std::move(tw);
    *out = (float2*) d;
    return USC_OK;
}

// scratch owned by the handle; grows (with a stream sync) only when a call needs more than before
static int reserve_work(usc_handle* h, size_t bytes) {
    if (h->work_bytes >= bytes) return USC_OK;
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_work);
    h->d_work = nullptr;
    h->work_bytes = 0;
    CK(cudaMalloc((void**) &h->d_work, bytes));
    h->work_bytes = bytes;
    return USC_OK;
}

static int make_plan(usc_handle* h, uint32_t n_complex, uint32_t tw_len, fft_plan_dev* plan) {
    plan->n = n_complex;
    plan->nrad = usc_host_radices(n_complex, plan->rad);
    if (!plan->nrad) return USC_ERR_ARGUMENT;
    plan->tw_n = tw_len;
    float2* tw = nullptr;
    int rc = get_twiddles(h, tw_len, &tw);
    if (rc) return rc;
    plan->tw = tw;
    return USC_OK;
}

extern "C" {

void usc_default_config(usc_config* cfg) {
    memset(cfg, 0, sizeof *cfg);
    cfg->n = 2048;                    /* receiver/Inc/main.h:97 */
    cfg->fs = 78125.0f;               /* 80 MHz / 32 / 32 / 1, receiver/Src/dfsdm.c:59-61,69 */
    cfg->f0 = 16000.0f;               /* receiver/Inc/chirp.h:18 */
    cfg->f1 = 19000.0f;               /* receiver/Inc/chirp.h:19 */
    cfg->sweep_T = 0.0205f;           /* receiver/Inc/chirp.h:16 */
    cfg->chirp_variant = USC_CHIRP_R;
    cfg->window = USC_HANN_PERIODIC;
    cfg->snr_threshold = 2.0f;        /* receiver/Inc/main.h:98 */
}

const char* usc_error_string(int code) {
    if (code == USC_OK) return "ok";
    if (code == USC_ERR_ARGUMENT) return "argument error (ARM_MATH_ARGUMENT_ERROR)";
    if (code == USC_ERR_NOMEM) return "out of host memory";
    if (code <= USC_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t) (USC_ERR_CUDA_BASE - code));
    return "unknown error";
}

int usc_create(const usc_config* cfg, int device, usc_handle** out) {
    if (!cfg || !out) return USC_ERR_ARGUMENT;
    *out = nullptr;
    if (!pow2(cfg->n) || cfg->n < 32 || cfg->n > 65536) return USC_ERR_ARGUMENT;
    if (!(cfg->fs > 0.0f) || cfg->chirp_variant > USC_CHIRP_F || cfg->window > USC_HANN_SYMMETRIC) return USC_ERR_ARGUMENT;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));                    // no device -> error: there is no CPU fallback
    if (device < 0 || device >= ndev) return USC_ERR_ARGUMENT;
    device_guard guard__(device);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));      // before the handle exists: nothing to release on failure
    usc_handle* h = new (std::nothrow) usc_handle();
    if (!h) return USC_ERR_NOMEM;
    h->dbg_fft_generic = getenv("USC_FFT_GENERIC") != nullptr;
    h->dbg_long_unfused = getenv("USC_LONG_UNFUSED") != nullptr;
    h->dbg_iq_unfused = getenv("USC_IQ_UNFUSED") != nullptr;
    h->cfg = *cfg;
    h->device = device;
    h->stream = 0;
    h->launches = 0;
    h->d_hann = h->d_up = h->d_down = h->d_ud = h->d_H_up = h->d_H_down = nullptr;
    h->d_tw_pass = h->d_tw_split = h->d_tw_l0 = nullptr;
    h->d_G_up = h->d_G_down = h->d_tw_split4096 = nullptr;
    h->d_fir_coeffs = nullptr;
    h->d_work = nullptr;
    h->work_bytes = 0;
    h->d_sym_table = nullptr;
    h->sym_amp = 0.0;
    h->sym_iq[0] = h->sym_iq[1] = h->sym_iq[2] = h->sym_iq[3] = 0.0;
    h->d_iq_cos = h->d_iq_sin = h->d_iq_chirp = h->d_iq_conj = h->d_iq_hann = h->d_iq_taps = nullptr;
    h->iq_ntaps = h->iq_window = 0;
    h->lane_frames = 0;
    h->lane_in_bytes = h->lane_out_bytes = 0;
    for (int i = 0; i < 3; ++i) { h->lane_stream[i] = nullptr; h->lane_done[i] = nullptr; h->lane_in[i] = h->lane_out[i] = nullptr; }
    h->rxh_state = nullptr; h->rxh_state_n = 0; h->rxh_stage = nullptr; h->rxh_stage_bytes = 0;
    h->num_sms = prop.multiProcessorCount;
    const uint32_t n = cfg->n;
    h->bandwidth = usc_host_bandwidth(n, cfg->fs, cfg->f0, cfg->f1);        /* main.c:372 */
    h->bandwidth2 = h->bandwidth * 2;                                       /* main.c:373 */
    h->idx_left_zero = n - h->bandwidth2;                                   /* main.c:374 */

    const bool cplx = cfg->chirp_variant == USC_CHIRP_S;
    h->hann.resize(n);
    h->up.resize(cplx ? 2 * n : n);
    h->down.resize(cplx ? 2 * n : n);
    usc_host_hann(h->hann.data(), n, cfg->window);
    float phase = (cfg->chirp_variant <= USC_CHIRP_S) ? -90.0f : (float) (-3.14159265358979f / 2.0);
    usc_host_ref_chirp(cfg->chirp_variant, n, cfg->fs, cfg->f0, cfg->f1, cfg->sweep_T, phase, 1, h->up.data());
    usc_host_ref_chirp(cfg->chirp_variant, n, cfg->fs, cfg->f0, cfg->f1, cfg->sweep_T, phase, 0, h->down.data());

    int rc;
    if ((rc = upload(h->hann.data(), h->hann.size() * 4, (void**) &h->d_hann))) { usc_destroy(h); return rc; }
    if ((rc = upload(h->up.data(), h->up.size() * 4, (void**) &h->d_up))) { usc_destroy(h); return rc; }
    if ((rc = upload(h->down.data(), h->down.size() * 4, (void**) &h->d_down))) { usc_destroy(h); return rc; }
    if (!cplx) {
        std::vector<float> ud(2 * (size_t) n);
        for (uint32_t i = 0; i < n; ++i) { ud[2 * i] = h->up[i]; ud[2 * i + 1] = h->down[i]; }
        if ((rc = upload(ud.data(), ud.size() * 4, (void**) &h->d_ud))) { usc_destroy(h); return rc; }
    }
    if (cudaMalloc((void**) &h->d_fir_coeffs, 256 * sizeof(float)) != cudaSuccess) { usc_destroy(h); return USC_ERR_CUDA_BASE - (int) cudaErrorMemoryAllocation; }
    cudaError_t e = fft_generic_prepare();
    if (e != cudaSuccess) { usc_destroy(h); return cuda_rc(e); }

    // master table W_n (n entries) serves the n-point real FFT (n/2 complex) and its split stage
    float2* d_master = nullptr;
    if ((rc = get_twiddles(h, n, &d_master))) { usc_destroy(h); return rc; }
    if (n >= 2048) {
        // fused-kernel tables: pass twiddles W_1024^(a d) laid out [d][a] (identical values for every master
        // length: the double argument 2*pi*j/1024 is reproduced exactly); split table (cos, sin)(2 pi k / 2048)
        const std::vector<float>& tw = h->tw_host[n];
        std::vector<float> pass(2 * 1024);
        const uint32_t step = n / 1024;
        for (uint32_t d = 0; d < 32; ++d)
            for (uint32_t a = 0; a < 32; ++a) {
                const uint32_t j = a * d * step;                 // W_1024^(ad) = W_n^(ad * n/1024)
                pass[2 * (d * 32 + a)] = tw[2 * j];
                pass[2 * (d * 32 + a) + 1] = tw[2 * j + 1];
            }
        if ((rc = upload(pass.data(), pass.size() * 4, (void**) &h->d_tw_pass))) { usc_destroy(h); return rc; }
        if (n == 65536 || n == 32768) {
            // level-0 twiddles of the [R0, 32, 32] plan (R0 = n / 2048): W_{n/2}^(a d) = W_n^(2 a d), laid out [d][a] for coalescing
            const uint32_t r0 = n / 2048;
            std::vector<float> l0(2 * (size_t) r0 * 1024);
            for (uint32_t d = 0; d < r0; ++d)
                for (uint32_t a = 0; a < 1024; ++a) {
                    const uint32_t j = 2 * a * d;
                    l0[2 * (d * 1024 + a)] = tw[2 * j];
                    l0[2 * (d * 1024 + a) + 1] = tw[2 * j + 1];
                }
            if ((rc = upload(l0.data(), l0.size() * 4, (void**) &h->d_tw_l0))) { usc_destroy(h); return rc; }
        }
        if (n == 2048) {
            std::vector<float> split(2 * 1024);
            for (uint32_t k = 0; k < 1024; ++k) {
                split[2 * k] = tw[2 * k];
                split[2 * k + 1] = -tw[2 * k + 1];
            }
            if ((rc = upload(split.data(), split.size() * 4, (void**) &h->d_tw_split))) { usc_destroy(h); return rc; }
        }
    }
    if (cfg->chirp_variant == USC_CHIRP_T) {
        // init_ref_chirp of experiments/chirp_compression_time_domain/Src/chirp.c:52-75:
        // H = rfft(window * chirp), computed once on the device with the same canonical FFT.
        std::vector<float> wu(n), wd(n);
        for (uint32_t i = 0; i < n; ++i) { wu[i] = h->up[i] * h->hann[i]; wd[i] = h->down[i] * h->hann[i]; }
        if ((rc = upload(wu.data(), n * 4, (void**) &h->d_H_up))) { usc_destroy(h); return rc; }
        if ((rc = upload(wd.data(), n * 4, (void**) &h->d_H_down))) { usc_destroy(h); return rc; }
        fft_plan_dev plan;
        if ((rc = make_plan(h, n / 2, n, &plan))) { usc_destroy(h); return rc; }
        if ((e = launch_fft_generic(FFT_R2C, plan, h->d_H_up, h->d_H_up, 1, 0)) != cudaSuccess ||
            (e = launch_fft_generic(FFT_R2C, plan, h->d_H_down, h->d_H_down, 1, 0)) != cudaSuccess ||
            (e = cudaDeviceSynchronize()) != cudaSuccess) { usc_destroy(h); return cuda_rc(e); }
        h->H_up.resize(n);
        h->H_down.resize(n);
        cudaMemcpy(h->H_up.data(), h->d_H_up, n * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(h->H_down.data(), h->d_H_down, n * 4, cudaMemcpyDeviceToHost);
    }
    *out = h;
    return USC_OK;
}

void usc_destroy(usc_handle* h) {
    if (!h) return;
    device_guard guard__(h->device);
    cudaFree(h->d_hann); cudaFree(h->d_up); cudaFree(h->d_down); cudaFree(h->d_ud); cudaFree(h->d_H_up); cudaFree(h->d_H_down);
    cudaFree(h->d_G_up); cudaFree(h->d_G_down); cudaFree(h->d_tw_split4096);
    cudaFree(h->d_tw_pass); cudaFree(h->d_tw_split); cudaFree(h->d_tw_l0); cudaFree(h->d_rs_taps); cudaFree(h->d_op_pass); cudaFree(h->d_op_split); cudaFree(h->d_fir_coeffs); cudaFree(h->d_work);
    cudaFree(h->d_iq_cos); cudaFree(h->d_iq_sin); cudaFree(h->d_iq_chirp); cudaFree(h->d_iq_conj); cudaFree(h->d_iq_hann); cudaFree(h->d_iq_taps);
    cudaFree(h->d_sym_table);
    for (auto& kv : h->tw_cache) cudaFree(kv.second);
    for (int i = 0; i < 3; ++i) {
        cudaFree(h->lane_in[i]); cudaFree(h->lane_out[i]);
        if (h->lane_done[i]) cudaEventDestroy(h->lane_done[i]);
        if (h->lane_stream[i]) cudaStreamDestroy(h->lane_stream[i]);
    }
    cudaFree(h->rxh_state);
    if (h->rxh_stage) cudaFreeHost(h->rxh_stage);
    delete h;
}

int usc_set_stream(usc_handle* h, void* cuda_stream) {
    USC_ENTER(h);
    if (!h) return USC_ERR_ARGUMENT;
    h->stream = (cudaStream_t) cuda_stream;
    return USC_OK;
}

int usc_sync(usc_handle* h) {
    USC_ENTER(h);
    if (!h) return USC_ERR_ARGUMENT;
    CK(cudaStreamSynchronize(h->stream));
    return USC_OK;
}

int usc_get_geometry(const usc_handle* h, uint32_t* bandwidth, uint32_t* bandwidth2, uint32_t* idx_left_zero) {
    if (!h) return USC_ERR_ARGUMENT;
    if (bandwidth) *bandwidth = h->bandwidth;
    if (bandwidth2) *bandwidth2 = h->bandwidth2;
    if (idx_left_zero) *idx_left_zero = h->idx_left_zero;
    return USC_OK;
}

int usc_get_table(const usc_handle* h, const char* what, float* dst, size_t cap) {
    if (!h || !what) return USC_ERR_ARGUMENT;
    const std::vector<float>* t = nullptr;
    if (!strcmp(what, "hann")) t = &h->hann;
    else if (!strcmp(what, "up")) t = &h->up;
    else if (!strcmp(what, "down")) t = &h->down;
    else if (!strcmp(what, "H_up")) t = &h->H_up;
    else if (!strcmp(what, "H_down")) t = &h->H_down;
    else if (!strcmp(what, "twiddle")) {
        auto it = h->tw_host.find(h->cfg.n);
        if (it != h->tw_host.end()) t = &it->second;
    }
    if (!t || t->empty()) return USC_ERR_ARGUMENT;
    if (dst) {
        if (cap < t->size()) return USC_ERR_ARGUMENT;
        memcpy(dst, t->data(), t->size() * sizeof(float));
    }
    return (int) t->size();
}

uint64_t usc_launch_count(const usc_handle* h) { return h ? h->launches : 0; }

int usc_malloc(void** dptr, size_t bytes) {
    if (!dptr) return USC_ERR_ARGUMENT;
    CK(cudaMalloc(dptr, bytes));
    return USC_OK;
}
int usc_malloc_on(usc_handle* h, void** dptr, size_t bytes) {
    USC_ENTER(h);
    if (!h) return USC_ERR_ARGUMENT;
    return usc_malloc(dptr, bytes);
}

int usc_free(void* dptr) {
    CK(cudaFree(dptr));
    return USC_OK;
}
int usc_malloc_host(void** hptr, size_t bytes) {
    if (!hptr) return USC_ERR_ARGUMENT;
    CK(cudaMallocHost(hptr, bytes));
    return USC_OK;
}
int usc_free_host(void* hptr) {
    CK(cudaFreeHost(hptr));
    return USC_OK;
}
int usc_memcpy_h2d(usc_handle* h, void* dst, const void* src, size_t bytes) {
    USC_ENTER(h);
    if (!h) return USC_ERR_ARGUMENT;
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    return USC_OK;
}
int usc_memcpy_d2h(usc_handle* h, void* dst, const void* src, size_t bytes) {
    USC_ENTER(h);
    if (!h) return USC_ERR_ARGUMENT;
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    return USC_OK;
}
int usc_memset(usc_handle* h, void* dst, int value, size_t bytes) {
    USC_ENTER(h);
    if (!h) return USC_ERR_ARGUMENT;
    CK(cudaMemsetAsync(dst, value, bytes, h->stream));
    return USC_OK;
}

/* ---- batched CMSIS-shaped operators ---- */
/* launches run under the entry point's device_guard, so the handle's device is current here */
#define LAUNCHED(h, expr)     \
    do {                      \
        CK(expr);             \
        (h)->launches++;      \
    } while (0)

int usc_i32_to_f32(usc_handle* h, const int32_t* src, float* dst, size_t count) {
    USC_ENTER(h);
    if (!h || !src || !dst) return USC_ERR_ARGUMENT;
    if (!count) return USC_OK;
    LAUNCHED(h, launch_i32_to_f32(src, dst, count, h->stream));
    return USC_OK;
}
int usc_arm_mult_f32_batch(usc_handle* h, const float* a, size_t sa, const float* b, size_t sb, float* dst,
                           size_t sd, uint32_t block_size, uint32_t batch) {
    USC_ENTER(h);
    if (!h || !a || !b || !dst) return USC_ERR_ARGUMENT;
    if (!block_size || !batch) return USC_OK;
    LAUNCHED(h, launch_mult(a, sa, b, sb, dst, sd, block_size, batch, h->stream));
    return USC_OK;
}
int usc_arm_scale_f32_batch(usc_handle* h, const float* src, float scale, float* dst, uint32_t block_size,
                            uint32_t batch) {
    USC_ENTER(h);
    if (!h || !src || !dst) return USC_ERR_ARGUMENT;
    if (!block_size || !batch) return USC_OK;
    LAUNCHED(h, launch_scale(src, scale, dst, (size_t) block_size * batch, h->stream));
    return USC_OK;
}
int usc_arm_cmplx_mult_cmplx_f32_batch(usc_handle* h, const float* a, size_t sa, const float* b, size_t sb,
                                       float* dst, size_t sd, uint32_t num_samples, uint32_t batch) {
    USC_ENTER(h);
    if (!h || !a || !b || !dst) return USC_ERR_ARGUMENT;
    if (!num_samples || !batch) return USC_OK;
    LAUNCHED(h, launch_cmul(a, sa, b, sb, dst, sd, num_samples, batch, h->stream));
    return USC_OK;
}
int usc_arm_cmplx_mult_real_f32_batch(usc_handle* h, const float* cplx, size_t sc, const float* real, size_t sr,
                                      float* dst, size_t sd, uint32_t num_samples, uint32_t batch) {
    USC_ENTER(h);
    if (!h || !cplx || !real || !dst) return USC_ERR_ARGUMENT;
    if (!num_samples || !batch) return USC_OK;
    LAUNCHED(h, launch_cmul_real(cplx, sc, real, sr, dst, sd, num_samples, batch, h->stream));
    return USC_OK;
}
int usc_arm_cmplx_mag_f32_batch(usc_handle* h, const float* src, size_t ss, float* dst, size_t sd,
                                uint32_t num_samples, uint32_t batch) {
    USC_ENTER(h);
    if (!h || !src || !dst) return USC_ERR_ARGUMENT;
    if (!num_samples || !batch) return USC_OK;
    const cudaError_t e = launch_cmag(src, ss, dst, sd, num_samples, batch, h->stream);
    if (e == cudaErrorInvalidValue) return USC_ERR_ARGUMENT;      /* overlapping src/dst other than the in-place form */
    LAUNCHED(h, e);
    return USC_OK;
}
int usc_arm_max_f32_batch(usc_handle* h, const float* src, size_t ss, uint32_t block_size, float* result,
                          uint32_t* index, uint32_t batch) {
    USC_ENTER(h);
    if (!h || !src || !result || !block_size) return USC_ERR_ARGUMENT;
    if (!batch) return USC_OK;
    LAUNCHED(h, launch_max(src, ss, block_size, result, index, batch, h->stream));
    return USC_OK;
}
int usc_arm_mean_f32_batch(usc_handle* h, const float* src, size_t ss, uint32_t block_size, float* result,
                           uint32_t batch) {
    USC_ENTER(h);
    if (!h || !src || !result || !block_size) return USC_ERR_ARGUMENT;
    if (!batch) return USC_OK;
    LAUNCHED(h, launch_mean(src, ss, block_size, result, batch, h->stream));
    return USC_OK;
}
/* Tables of the warp-level FFT operators (k_fft_warp.cu): W_1024^(a d) as [d][a] and the 2048-point split
 * twiddles.  A handle configured for n = 2048 already owns them; other handles build them on first use. */
static int ensure_op_tables(usc_handle* h, const float2** pass, const float2** split) {
    if (h->d_tw_pass && h->d_tw_split) { *pass = h->d_tw_pass; *split = h->d_tw_split; return USC_OK; }
    if (!h->d_op_pass) {
        float2* master = nullptr;
        int rc = get_twiddles(h, 2048, &master);
        if (rc) return rc;
        const std::vector<float>& tw = h->tw_host[2048];
        std::vector<float> pass_h(2 * 1024), split_h(2 * 1024);
        for (uint32_t d = 0; d < 32; ++d)
            for (uint32_t a = 0; a < 32; ++a) {
                const uint32_t j = a * d * 2;                       /* W_1024^(ad) = W_2048^(2ad) */
                pass_h[2 * (d * 32 + a)] = tw[2 * j];
                pass_h[2 * (d * 32 + a) + 1] = tw[2 * j + 1];
            }
        for (uint32_t k = 0; k < 1024; ++k) { split_h[2 * k] = tw[2 * k]; split_h[2 * k + 1] = -tw[2 * k + 1]; }
        if ((rc = upload(pass_h.data(), pass_h.size() * 4, (void**) &h->d_op_pass))) return rc;
        if ((rc = upload(split_h.data(), split_h.size() * 4, (void**) &h->d_op_split))) return rc;
    }
    *pass = h->d_op_pass; *split = h->d_op_split;
    return USC_OK;
}

/* forward transform in place on an aligned scratch buffer: the warp-level operator where it exists, else the generic kernel */
static int fft_forward_inplace(usc_handle* h, int mode, uint32_t n_complex, const fft_plan_dev& plan, float* buf, uint32_t batch) {
    if (n_complex == 1024 && ((uintptr_t) buf & 15u) == 0 && !h->dbg_fft_generic) {
        const float2 *pass, *split;
        int rc = ensure_op_tables(h, &pass, &split);
        if (rc) return rc;
        LAUNCHED(h, launch_fft_warp(mode, buf, buf, batch, pass, split, h->num_sms, h->stream));
        return USC_OK;
    }
    LAUNCHED(h, launch_fft_generic(mode, plan, buf, buf, batch, h->stream));
    return USC_OK;
}

int usc_arm_rfft_fast_f32_batch(usc_handle* h, uint32_t fft_len, const float* in, float* out, uint8_t ifft_flag,
                                uint32_t batch) {
    USC_ENTER(h);
    if (h && in && out && fft_len == 2048 && batch && (((uintptr_t) in | (uintptr_t) out) & 15u) == 0 && !h->dbg_fft_generic) {
        /* the receiver's own length: two transforms per warp on the packed register core, one pass over HBM */
        const float2 *pass, *split;
        int rc = ensure_op_tables(h, &pass, &split);
        if (rc) return rc;
        LAUNCHED(h, launch_fft_warp(ifft_flag ? FFT_C2R : FFT_R2C, in, out, batch, pass, split, h->num_sms, h->stream));
        return USC_OK;
    }
    /* supported lengths: CMSIS's 32..4096 (arm_math.h:2242-2244 returns ARM_MATH_ARGUMENT_ERROR
     * otherwise) extended to 8192 while one transform fits shared memory */
    if (!h || !in || !out || !pow2(fft_len) || fft_len < 32 || fft_len > 65536) return USC_ERR_ARGUMENT;
    if (ifft_flag && fft_len > 16384) return USC_ERR_ARGUMENT;      /* inverse: shared-memory sizes only */
    if (!batch) return USC_OK;
    fft_plan_dev plan;
    int rc = make_plan(h, fft_len / 2, fft_len, &plan);
    if (rc) return rc;
    if (fft_len > 16384) {                                          /* 32768 / 65536: level 0 in global memory */
        if ((rc = reserve_work(h, (size_t) batch * fft_len * sizeof(float)))) return rc;
        CK(launch_fft_large(FFT_R2C, plan, in, out, h->d_work, batch, h->stream));
        h->launches += 3;
        return USC_OK;
    }
    LAUNCHED(h, launch_fft_generic(ifft_flag ? FFT_C2R : FFT_R2C, plan, in, out, batch, h->stream));
    return USC_OK;
}
int usc_arm_cfft_f32_batch(usc_handle* h, uint32_t fft_len, float* data, uint8_t ifft_flag, uint32_t batch) {
    USC_ENTER(h);
    if (h && data && fft_len == 2048 && batch && ((uintptr_t) data & 15u) == 0 && !h->dbg_fft_generic) {
        /* arm_cfft_sR_f32_len2048 (experiments/synchronization): [2, 32, 32] on the register core, one transform per warp pass */
        const float2 *pass, *split;
        float2* master = nullptr;
        int rc = ensure_op_tables(h, &pass, &split);
        if (!rc) rc = get_twiddles(h, 2048, &master);
        if (rc) return rc;
        LAUNCHED(h, launch_cfft2048_warp(ifft_flag != 0, data, batch, pass, master, h->num_sms, h->stream));
        return USC_OK;
    }
    if (h && data && fft_len == 1024 && batch && ((uintptr_t) data & 15u) == 0 && !h->dbg_fft_generic) {
        const float2 *pass, *split;
        int rc = ensure_op_tables(h, &pass, &split);
        if (rc) return rc;
        LAUNCHED(h, launch_fft_warp(ifft_flag ? FFT_C2C_INV : FFT_C2C_FWD, data, data, batch, pass, split, h->num_sms, h->stream));
        return USC_OK;
    }
    if (!h || !data || !pow2(fft_len) || fft_len < 16 || fft_len > 32768) return USC_ERR_ARGUMENT;
    if (ifft_flag && fft_len > 8192) return USC_ERR_ARGUMENT;
    if (!batch) return USC_OK;
    fft_plan_dev plan;
    int rc = make_plan(h, fft_len, fft_len, &plan);
    if (rc) return rc;
    if (fft_len > 8192) {
        if ((rc = reserve_work(h, (size_t) batch * fft_len * 2 * sizeof(float)))) return rc;
        CK(launch_fft_large(FFT_C2C_FWD, plan, data, data, h->d_work, batch, h->stream));
        h->launches += 3;
        return USC_OK;
    }
    LAUNCHED(h, launch_fft_generic(ifft_flag ? FFT_C2C_INV : FFT_C2C_FWD, plan, data, data, batch, h->stream));
    return USC_OK;
}
int usc_arm_fir_f32_batch(usc_handle* h, const float* coeffs_host, uint32_t num_taps, float* state,
                          const float* src, float* dst, uint32_t block_size, uint32_t batch) {
    USC_ENTER(h);
    if (!h || !coeffs_host || !state || !src || !dst || num_taps < 1 || num_taps > 256 || !block_size ||
        block_size > 8192)
        return USC_ERR_ARGUMENT;
    if (!batch) return USC_OK;
    CK(cudaMemcpyAsync(h->d_fir_coeffs, coeffs_host, num_taps * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    LAUNCHED(h, launch_fir(h->d_fir_coeffs, num_taps, state, src, dst, block_size, batch, h->stream));
    return USC_OK;
}

/* ---- fused stage-level operators ---- */
static void fill_common(const usc_handle* h, demod_params* p) {
    memset(p, 0, sizeof *p);
    p->chirp_up = (const float2*) h->d_up;
    p->chirp_down = (const float2*) h->d_down;
    p->chirp_ud = (const float2*) h->d_ud;
    p->hann = (const float2*) h->d_hann;
    p->tw_pass = h->d_tw_pass;
    p->tw_split = h->d_tw_split;
    { auto it = h->tw_cache.find(2048); p->tw_master = it == h->tw_cache.end() ? nullptr : it->second; }
    p->bandwidth2 = h->bandwidth2;
    p->idx_left_zero = h->idx_left_zero;
    p->fs_int = (int32_t) h->cfg.fs;
}

// Any frame length (32..65536): the same chain operator by operator on handle-owned scratch
// (cast, de-chirp, window, RFFT, magnitude, arg-max), once per hypothesis.  Used for the long-frame
// sweep (BASELINE config 5) and for geometries the fused 2048-point kernel does not cover.
static int demod_generic(usc_handle* h, const void* pcm, uint32_t pcm_format, size_t nframes, float* mag_up,
                         uint32_t* idx_up, float* mag_down, uint32_t* idx_down, uint8_t* bit) {
    const uint32_t n = h->cfg.n;
    if (nframes > 0xffffffffu || h->bandwidth2 == 0 || h->bandwidth2 > n / 2) return USC_ERR_ARGUMENT;
    const uint32_t B = (uint32_t) nframes;
    const size_t W = nframes * n;
    int rc = reserve_work(h, (2 * W + 2 * nframes) * sizeof(float));
    if (rc) return rc;
    float *fA = h->d_work, *fB = fA + W, *tmp_mag = fB + W;
    fft_plan_dev plan;
    if ((rc = make_plan(h, n / 2, n, &plan))) return rc;
    float* mags[2] = {mag_up ? mag_up : tmp_mag, mag_down ? mag_down : tmp_mag + nframes};
    uint32_t* idxs[2] = {idx_up, idx_down};
    const float* chirps[2] = {h->d_up, h->d_down};
    for (int hyp = 0; hyp < 2; ++hyp) {
        /* cast, de-chirp and window in one pass; RFFT; magnitude + windowed arg-max in one pass */
        LAUNCHED(h, launch_prep(pcm, pcm_format, chirps[hyp], h->d_hann, fA, n, W, h->stream));
        if (n > 16384) {
            CK(launch_fft_large(FFT_R2C, plan, fA, fA, fB, B, h->stream));
            h->launches += 3;
        } else {
            LAUNCHED(h, launch_fft_generic(FFT_R2C, plan, fA, fA, B, h->stream));
        }
        LAUNCHED(h, launch_mag_max(fA, n, h->bandwidth2, mags[hyp], idxs[hyp], B, h->stream));
    }
    if (bit) LAUNCHED(h, launch_decide(mags[0], mags[1], bit, nframes, h->stream));
    return USC_OK;
}

int usc_demod_frames(usc_handle* h, const void* pcm, uint32_t pcm_format, size_t nframes, float* mag_up,
                     uint32_t* idx_up, float* mag_down, uint32_t* idx_down, uint8_t* bit) {
    USC_ENTER(h);
    if (!h || !pcm || pcm_format > USC_PCM_I32) return USC_ERR_ARGUMENT;
    if (h->cfg.chirp_variant == USC_CHIRP_S) return USC_ERR_ARGUMENT;
    if (((uintptr_t) pcm & 15u) != 0) return USC_ERR_ARGUMENT;     /* frames are fetched by 16-byte-aligned bulk copies */
    if (!nframes) return USC_OK;
    if ((h->cfg.n == 4096 || h->cfg.n == 8192 || h->cfg.n == 16384) && !h->dbg_long_unfused && h->bandwidth2 > 0 && h->bandwidth2 <= 160u * (h->cfg.n / 2048u)) {
        /* long frames: one CTA per frame, level 0 from global memory, packed 1024-point cores (k_long.cu) */
        float2* master = nullptr;
        int rc = get_twiddles(h, h->cfg.n, &master);
        if (rc) return rc;
        LAUNCHED(h, launch_demod_long(pcm, pcm_format, nframes, h->cfg.n, (const float2*) h->d_ud, (const float2*) h->d_hann,
                                      master, h->d_tw_pass, h->bandwidth2, mag_up, idx_up, mag_down, idx_down, bit,
                                      h->num_sms, h->stream));
        return USC_OK;
    }
    if ((h->cfg.n == 65536 || h->cfg.n == 32768) && h->d_tw_l0 && h->bandwidth2 > 0 && h->bandwidth2 <= 160u * (h->cfg.n / 2048u) &&
        !h->dbg_long_unfused) {
        /* 32768 / 65536-point frames: a cluster of two / four CTAs per frame, sub-sequences in distributed shared memory (k_long.cu) */
        float2* master = nullptr;
        int rc = get_twiddles(h, h->cfg.n, &master);
        if (rc) return rc;
        LAUNCHED(h, launch_demod_long32(pcm, pcm_format, nframes, h->cfg.n, (const float2*) h->d_ud, (const float2*) h->d_hann, master,
                                        h->d_tw_pass, h->d_tw_l0, h->bandwidth2, mag_up, idx_up, mag_down, idx_down, bit,
                                        h->num_sms, h->stream));
        return USC_OK;
    }
    if (h->cfg.n != 2048 || h->bandwidth2 == 0 || h->bandwidth2 > 512)
        return demod_generic(h, pcm, pcm_format, nframes, mag_up, idx_up, mag_down, idx_down, bit);
    demod_params p;
    fill_common(h, &p);
    p.pcm = pcm;
    p.nframes = nframes;
    p.mag_up = mag_up; p.idx_up = idx_up; p.mag_down = mag_down; p.idx_down = idx_down; p.bit = bit;
    const bool want_up = mag_up || idx_up, want_down = mag_down || idx_down;
    if (!bit && want_up != want_down) {
        /* only one hypothesis asked for: dsp(pos, .., UP) or dsp(pos, .., DOWN) alone — frames go
         * through the packed core two at a time */
        p.updown = want_up ? 1 : 0;
        LAUNCHED(h, launch_demod2048_single(p, pcm_format, h->num_sms, h->stream));
        return USC_OK;
    }
    LAUNCHED(h, launch_demod2048(p, pcm_format, h->num_sms, h->stream));
    return USC_OK;
}

/* ---- host-buffer entry points: chunks flow through three lanes (H2D copy, kernel, D2H of the results) ---- */
static int ensure_lanes(usc_handle* h, size_t in_bytes, size_t out_bytes) {
    for (int i = 0; i < 3; ++i) {
        if (!h->lane_stream[i]) CK(cudaStreamCreateWithFlags(&h->lane_stream[i], cudaStreamNonBlocking));
        if (!h->lane_done[i]) CK(cudaEventCreateWithFlags(&h->lane_done[i], cudaEventDisableTiming));
    }
    if (in_bytes > h->lane_in_bytes) {
        for (int i = 0; i < 3; ++i) { CK(cudaStreamSynchronize(h->lane_stream[i])); cudaFree(h->lane_in[i]); h->lane_in[i] = nullptr; }
        h->lane_in_bytes = 0;
        for (int i = 0; i < 3; ++i) CK(cudaMalloc(&h->lane_in[i], in_bytes));
        h->lane_in_bytes = in_bytes;
    }
    if (out_bytes > h->lane_out_bytes) {
        for (int i = 0; i < 3; ++i) { CK(cudaStreamSynchronize(h->lane_stream[i])); cudaFree(h->lane_out[i]); h->lane_out[i] = nullptr; }
        h->lane_out_bytes = 0;
        for (int i = 0; i < 3; ++i) CK(cudaMalloc(&h->lane_out[i], out_bytes));
        h->lane_out_bytes = out_bytes;
    }
    return USC_OK;
}
/* the five per-frame result vectors of a chunk of `cap` frames inside a lane's result arena (256-byte aligned) */
struct frame_results {
    float *mu, *md; uint32_t *iu, *id; uint8_t* bit;
    static size_t bytes(size_t cap) { return 4 * ((cap * 4 + 255) & ~(size_t) 255) + ((cap + 255) & ~(size_t) 255); }
    frame_results(void* arena, size_t cap) {
        const size_t v = (cap * 4 + 255) & ~(size_t) 255;
        char* p = (char*) arena;
        mu = (float*) p; md = (float*) (p + v); iu = (uint32_t*) (p + 2 * v); id = (uint32_t*) (p + 3 * v); bit = (uint8_t*) (p + 4 * v);
    }
};
/* runs `body` with the handle's stream swapped for a lane stream (the device-pointer entry points launch on h->stream) */
struct stream_swap {
    usc_handle* h; cudaStream_t saved;
    stream_swap(usc_handle* h_, cudaStream_t st) : h(h_), saved(h_->stream) { h->stream = st; }
    ~stream_swap() { h->stream = saved; }
};
static int sync_lanes(usc_handle* h) {
    for (int l = 0; l < 3; ++l) CK(cudaStreamSynchronize(h->lane_stream[l]));
    return USC_OK;
}
static int copy_frame_results(const frame_results& r, size_t f0, size_t nf, float* mag_up, uint32_t* idx_up, float* mag_down,
                              uint32_t* idx_down, uint8_t* bit, cudaStream_t st) {
    if (mag_up) CK(cudaMemcpyAsync(mag_up + f0, r.mu, nf * 4, cudaMemcpyDeviceToHost, st));
    if (idx_up) CK(cudaMemcpyAsync(idx_up + f0, r.iu, nf * 4, cudaMemcpyDeviceToHost, st));
    if (mag_down) CK(cudaMemcpyAsync(mag_down + f0, r.md, nf * 4, cudaMemcpyDeviceToHost, st));
    if (idx_down) CK(cudaMemcpyAsync(idx_down + f0, r.id, nf * 4, cudaMemcpyDeviceToHost, st));
    if (bit) CK(cudaMemcpyAsync(bit + f0, r.bit, nf, cudaMemcpyDeviceToHost, st));
    return USC_OK;
}

int usc_host_workspace(usc_handle* h, size_t chunk_frames) {
    USC_ENTER(h);
    if (!h || !chunk_frames) return USC_ERR_ARGUMENT;
    h->lane_frames = chunk_frames;
    return ensure_lanes(h, chunk_frames * 8192, frame_results::bytes(chunk_frames));
}

int usc_demod_frames_host(usc_handle* h, const void* pcm_host, uint32_t pcm_format, size_t nframes,
                          float* mag_up, uint32_t* idx_up, float* mag_down, uint32_t* idx_down, uint8_t* bit) {
    USC_ENTER(h);
    if (!h || !pcm_host || pcm_format > USC_PCM_I32) return USC_ERR_ARGUMENT;
    const uint32_t n = h->cfg.n;
    if (h->cfg.chirp_variant == USC_CHIRP_S || h->bandwidth2 == 0) return USC_ERR_ARGUMENT;
    /* frame lengths with a fused kernel only: the operator chain of the other lengths shares one scratch arena */
    const bool fused = (n == 2048 && h->bandwidth2 <= 512) ||
                       (n >= 4096 && n <= 65536 && h->bandwidth2 <= 160u * (n / 2048u) && !h->dbg_long_unfused && (n <= 16384 || h->d_tw_l0));
    if (!fused) return USC_ERR_ARGUMENT;
    if (!nframes) return USC_OK;
    if (!h->lane_frames) h->lane_frames = 4096;
    size_t cf = h->lane_frames * 2048 / n;                       /* frames of n samples per chunk */
    if (!cf) cf = 1;
    int rc = ensure_lanes(h, cf * n * 4, frame_results::bytes(cf));
    if (rc) return rc;
    const char* src = (const char*) pcm_host;
    size_t chunk = 0;
    for (size_t f0 = 0; f0 < nframes; f0 += cf, ++chunk) {
        const int l = (int) (chunk % 3);
        const size_t nf = nframes - f0 < cf ? nframes - f0 : cf;
        cudaStream_t st = h->lane_stream[l];
        CK(cudaMemcpyAsync(h->lane_in[l], src + f0 * n * 4, nf * n * 4, cudaMemcpyHostToDevice, st));
        frame_results r(h->lane_out[l], cf);
        {
            stream_swap sw(h, st);
            rc = usc_demod_frames(h, h->lane_in[l], pcm_format, nf, r.mu, r.iu, r.md, r.id, r.bit);
        }
        if (rc) return rc;
        if ((rc = copy_frame_results(r, f0, nf, mag_up, idx_up, mag_down, idx_down, bit, st))) return rc;
    }
    return sync_lanes(h);
}

int usc_iq_demod_host(usc_handle* h, const void* pcm_host, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes,
                      size_t stream_stride, float* mag_up, uint32_t* idx_up, float* mag_down, uint32_t* idx_down,
                      uint8_t* bit) {
    USC_ENTER(h);
    if (!h || !pcm_host || pcm_format > USC_PCM_I32 || !h->d_iq_taps) return USC_ERR_ARGUMENT;
    const uint32_t n = h->cfg.n;
    if (n != 2048 || h->iq_window > 32 || h->iq_ntaps > 32 || !h->d_tw_pass || h->dbg_iq_unfused) return USC_ERR_ARGUMENT;   /* fused kernel only */
    if (stream_stride < (size_t) nframes * n) return USC_ERR_ARGUMENT;
    if (!nstreams || !nframes) return USC_OK;
    if (!h->lane_frames) h->lane_frames = 4096;
    /* the FIR state runs along a stream, so chunks are whole streams */
    size_t cs = h->lane_frames / nframes;
    if (!cs) cs = 1;
    const size_t row = (size_t) nframes * n * 4;                 /* bytes of one stream on the device (packed) */
    int rc = ensure_lanes(h, cs * row, frame_results::bytes(cs * nframes));
    if (rc) return rc;
    const char* src = (const char*) pcm_host;
    size_t chunk = 0;
    for (size_t s0 = 0; s0 < nstreams; s0 += cs, ++chunk) {
        const int l = (int) (chunk % 3);
        const size_t ns = nstreams - s0 < cs ? nstreams - s0 : cs;
        cudaStream_t st = h->lane_stream[l];
        CK(cudaMemcpy2DAsync(h->lane_in[l], row, src + s0 * stream_stride * 4, stream_stride * 4, row, ns, cudaMemcpyHostToDevice, st));
        frame_results r(h->lane_out[l], cs * nframes);
        {
            stream_swap sw(h, st);
            rc = usc_iq_demod(h, h->lane_in[l], pcm_format, (uint32_t) ns, nframes, (size_t) nframes * n, r.mu, r.iu, r.md, r.id, r.bit);
        }
        if (rc) return rc;
        if ((rc = copy_frame_results(r, s0 * nframes, ns * nframes, mag_up, idx_up, mag_down, idx_down, bit, st))) return rc;
    }
    return sync_lanes(h);
}

int usc_receiver_run_host(usc_handle* h, const void* pcm_host, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes,
                          size_t stream_stride, uint8_t* uart, uint32_t uart_cap, usc_rx_result* results) {
    USC_ENTER(h);
    if (!h || !pcm_host || pcm_format > USC_PCM_I32 || (!uart && !results)) return USC_ERR_ARGUMENT;
    if (h->cfg.n != 2048 || h->cfg.chirp_variant == USC_CHIRP_S || h->bandwidth2 == 0 || h->bandwidth2 > 160) return USC_ERR_ARGUMENT;
    if ((stream_stride & 1u) != 0 || stream_stride < (size_t) nframes * 2048) return USC_ERR_ARGUMENT;
    if (uart && !uart_cap) return USC_ERR_ARGUMENT;
    if (!nstreams || !nframes) return USC_OK;
    if (!h->lane_frames) h->lane_frames = 4096;
    /* The state machine runs along a stream and one warp serves a stream, so a chunk holds ALL streams and a slice of
     * time: cf new frames per stream behind the two frames of FIFO history the receiver keeps (main.c:659-668),
     * resumed through the carried per-stream state exactly as usc_receiver_run_chunk documents. */
    size_t cf = h->lane_frames / nstreams;
    if (cf < 6) cf = 6;
    if (cf > nframes) cf = nframes;
    const size_t nchunks = (nframes + cf - 1) / cf;
    const uint32_t cap = uart ? uart_cap : 0;
    const size_t dev_stride = (cf + 2) * 2048;                    /* samples per stream in a lane buffer */
    const size_t uart_bytes = ((size_t) nstreams * cap + 255) & ~(size_t) 255;
    const size_t out_bytes = uart_bytes + (size_t) nstreams * sizeof(usc_rx_result);
    int rc = ensure_lanes(h, (size_t) nstreams * dev_stride * 4, out_bytes);
    if (rc) return rc;
    if (h->rxh_state_n < nstreams) {
        if ((rc = sync_lanes(h))) return rc;
        cudaFree(h->rxh_state); h->rxh_state = nullptr; h->rxh_state_n = 0;
        CK(cudaMalloc((void**) &h->rxh_state, (size_t) nstreams * sizeof(usc_rx_state)));
        h->rxh_state_n = nstreams;
    }
    if (h->rxh_stage_bytes < nchunks * out_bytes) {
        if (h->rxh_stage) cudaFreeHost(h->rxh_stage);
        h->rxh_stage = nullptr; h->rxh_stage_bytes = 0;
        CK(cudaMallocHost((void**) &h->rxh_stage, nchunks * out_bytes));
        h->rxh_stage_bytes = nchunks * out_bytes;
    }
    CK(cudaMemsetAsync(h->rxh_state, 0, (size_t) nstreams * sizeof(usc_rx_state), h->lane_stream[0]));
    CK(cudaEventRecord(h->lane_done[2], h->lane_stream[0]));      /* "previous kernel" of the first chunk */
    const char* src = (const char*) pcm_host;
    const int esz = 4;
    for (size_t c = 0; c < nchunks; ++c) {
        const int l = (int) (c % 3), lprev = (int) ((c + 2) % 3);
        const size_t f0 = c * cf, nf = nframes - f0 < cf ? nframes - f0 : cf;
        const uint32_t carry = f0 >= 2 ? 2u : (uint32_t) f0;
        cudaStream_t st = h->lane_stream[l];
        CK(cudaMemcpy2DAsync(h->lane_in[l], dev_stride * esz, src + (f0 - carry) * 2048 * esz, stream_stride * esz,
                             (nf + carry) * 2048 * esz, nstreams, cudaMemcpyHostToDevice, st));
        CK(cudaStreamWaitEvent(st, h->lane_done[lprev], 0));      /* the carried state of chunk c-1 must be written */
        uint8_t* d_uart = cap ? (uint8_t*) h->lane_out[l] : nullptr;
        usc_rx_result* d_res = (usc_rx_result*) ((char*) h->lane_out[l] + uart_bytes);
        {
            stream_swap sw(h, st);
            rc = usc_receiver_run_chunk(h, h->lane_in[l], pcm_format, nstreams, (uint32_t) nf, dev_stride, carry, h->rxh_state, d_uart,
                                        cap, d_res);
        }
        if (rc) return rc;
        CK(cudaEventRecord(h->lane_done[l], st));
        CK(cudaMemcpyAsync(h->rxh_stage + c * out_bytes, h->lane_out[l], out_bytes, cudaMemcpyDeviceToHost, st));
    }
    if ((rc = sync_lanes(h))) return rc;
    /* stitch the chunks on the host: bytes are appended per stream in chunk order; the last chunk's record is the result */
    std::vector<uint32_t> filled(nstreams, 0);
    for (size_t c = 0; c < nchunks; ++c) {
        const uint8_t* cu = h->rxh_stage + c * out_bytes;
        const usc_rx_result* cr = (const usc_rx_result*) (cu + uart_bytes);
        for (uint32_t s = 0; s < nstreams; ++s) {
            const uint32_t nb = cr[s].nbytes;
            if (uart) {
                const uint32_t have = nb < cap ? nb : cap;
                for (uint32_t i = 0; i < have && filled[s] + i < cap; ++i) uart[(size_t) s * cap + filled[s] + i] = cu[(size_t) s * cap + i];
            }
            filled[s] += nb;
        }
        if (results && c + 1 == nchunks)
            for (uint32_t s = 0; s < nstreams; ++s) { results[s] = cr[s]; results[s].nbytes = filled[s]; }
    }
    return USC_OK;
}

int usc_dsp(usc_handle* h, const float* fifo, size_t fifo_stride, const uint32_t* sync_position,
            const float* mag_mean, int updown, usc_history* hist, uint32_t batch) {
    USC_ENTER(h);
    if (!h || !fifo || !sync_position || !mag_mean || !hist) return USC_ERR_ARGUMENT;
    if (h->cfg.n != 2048) return USC_ERR_ARGUMENT;
    if (h->cfg.chirp_variant == USC_CHIRP_S) {
        /* complex-FFT variant (experiments/synchronization/Src/main.c:161-213): both windows are real */
        if (h->bandwidth2 == 0 || h->bandwidth2 > 192) return USC_ERR_ARGUMENT;
        if (!batch) return USC_OK;
        demod_params q;
        fill_common(h, &q);
        q.pcm = fifo; q.nframes = batch; q.fifo_stride = fifo_stride; q.sync_position = sync_position;
        q.mag_mean = mag_mean; q.hist = (history_rec*) hist; q.updown = updown ? 1 : 0;
        LAUNCHED(h, launch_dsp2048c(q, h->num_sms, h->stream));
        return USC_OK;
    }
    if (h->bandwidth2 == 0 || h->bandwidth2 > 512) return USC_ERR_ARGUMENT;
    if (!batch) return USC_OK;
    demod_params p;
    fill_common(h, &p);
    p.pcm = fifo;
    p.nframes = batch;
    p.fifo_stride = fifo_stride;
    p.sync_position = sync_position;
    p.mag_mean = mag_mean;
    p.hist = (history_rec*) hist;
    p.updown = updown ? 1 : 0;
    LAUNCHED(h, launch_dsp2048(p, h->num_sms, h->stream));
    return USC_OK;
}

static int fill_rx(usc_handle* h, rx_launch* a, const void* pcm, uint32_t pcm_format, uint32_t nstreams,
                   uint32_t nframes, size_t stream_stride) {
    if (!h || !pcm || pcm_format > USC_PCM_I32) return USC_ERR_ARGUMENT;
    if (h->cfg.n != 2048 || h->cfg.chirp_variant == USC_CHIRP_S) return USC_ERR_ARGUMENT;
    if (h->bandwidth2 == 0 || h->bandwidth2 > 160) return USC_ERR_ARGUMENT;
    if (((uintptr_t) pcm & 7u) != 0 || (stream_stride & 1u) != 0 || stream_stride < (size_t) nframes * 2048) return USC_ERR_ARGUMENT;
    memset(a, 0, sizeof *a);
    a->pcm = pcm; a->pcm_format = pcm_format; a->nstreams = nstreams; a->nframes = nframes; a->stream_stride = stream_stride;
    a->up = (const float2*) h->d_up; a->down = (const float2*) h->d_down; a->hann = (const float2*) h->d_hann;
    a->tw_pass = h->d_tw_pass; a->tw_split = h->d_tw_split;
    a->bandwidth2 = h->bandwidth2; a->snr_threshold = h->cfg.snr_threshold;
    a->sync_add = 1;
    return USC_OK;
}

int usc_receiver_run(usc_handle* h, const void* pcm, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes,
                     size_t stream_stride, uint8_t* uart, uint32_t uart_cap, usc_rx_result* results) {
    USC_ENTER(h);
    rx_launch a;
    int rc = fill_rx(h, &a, pcm, pcm_format, nstreams, nframes, stream_stride);
    if (rc) return rc;
    if (!results && !uart) return USC_ERR_ARGUMENT;
    if (!nstreams || !nframes) return USC_OK;
    a.uart = uart; a.uart_cap = uart_cap; a.results = (rx_result_rec*) results;
    LAUNCHED(h, launch_receiver_run(a, h->num_sms, h->stream));
    return USC_OK;
}

static_assert(sizeof(usc_rx_state) == sizeof(rx_state_rec), "usc_rx_state layout");

int usc_receiver_run_chunk(usc_handle* h, const void* pcm, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes,
                           size_t stream_stride, uint32_t carry_frames, usc_rx_state* state, uint8_t* uart,
                           uint32_t uart_cap, usc_rx_result* results) {
    USC_ENTER(h);
    rx_launch a;
    if (carry_frames > 2) return USC_ERR_ARGUMENT;
    if (stream_stride < ((size_t) nframes + carry_frames) * 2048) return USC_ERR_ARGUMENT;
    int rc = fill_rx(h, &a, pcm, pcm_format, nstreams, nframes, stream_stride);
    if (rc) return rc;
    if (!state) return USC_ERR_ARGUMENT;
    if (!nstreams || !nframes) return USC_OK;
    a.uart = uart; a.uart_cap = uart_cap; a.results = (rx_result_rec*) results;
    a.carry = carry_frames; a.rx_state = (rx_state_rec*) state;
    LAUNCHED(h, launch_receiver_run(a, h->num_sms, h->stream));
    return USC_OK;
}

int usc_sync_search(usc_handle* h, const void* pcm, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes,
                    size_t stream_stride, uint32_t sync_add, float* mag, uint32_t* idx) {
    USC_ENTER(h);
    rx_launch a;
    int rc = fill_rx(h, &a, pcm, pcm_format, nstreams, nframes, stream_stride);
    if (rc) return rc;
    if (!mag || !idx || sync_add < 1 || sync_add > 64) return USC_ERR_ARGUMENT;
    if (!nstreams || !nframes) return USC_OK;
    a.sync_add = sync_add; a.ss_mag = mag; a.ss_idx = idx;
    LAUNCHED(h, launch_sync_search(a, h->num_sms, h->stream));
    return USC_OK;
}

int usc_scan4(usc_handle* h, const float* pcm2n, uint32_t batch, usc_scan_entry* out) {
    USC_ENTER(h);
    /* experiments/chirp_compression_freq_domain/Src/main.c:113-160, 245-251 on `batch` 2n-sample buffers */
    if (!h || !pcm2n || !out || h->cfg.chirp_variant == USC_CHIRP_S) return USC_ERR_ARGUMENT;
    const uint32_t n = h->cfg.n, bw8 = h->bandwidth * 8;
    if (n > 16384 || bw8 == 0 || bw8 > n / 2) return USC_ERR_ARGUMENT;
    if (!batch) return USC_OK;
    int rc = reserve_work(h, ((size_t) batch * n + 4 * (size_t) batch) * sizeof(float));
    if (rc) return rc;
    float* w = h->d_work;
    float *mr = w + (size_t) batch * n, *ml = mr + batch;
    uint32_t *ir = (uint32_t*) (ml + batch), *il = ir + batch;
    fft_plan_dev plan;
    if ((rc = make_plan(h, n / 2, n, &plan))) return rc;
    for (uint32_t i = 0; i < 4; ++i) {
        const float* src = pcm2n + (size_t) (n / 4) * i;                               /* main.c:246-249 */
        LAUNCHED(h, launch_mult(src, 2 * (size_t) n, h->d_down, 0, w, n, n, batch, h->stream));   /* mult_ref_chirp (down-chirp) */
        LAUNCHED(h, launch_mult(w, n, h->d_hann, 0, w, n, n, batch, h->stream));
        if ((rc = fft_forward_inplace(h, FFT_R2C, n / 2, plan, w, batch))) return rc;
        LAUNCHED(h, launch_pipeline_tail(w, n, batch, 0, h->stream));                  /* in-place magnitudes, upper half kept */
        LAUNCHED(h, launch_max(w, n, bw8, mr, ir, batch, h->stream));
        LAUNCHED(h, launch_max(w + (n - bw8), n, bw8, ml, il, batch, h->stream));
        LAUNCHED(h, launch_scan_pack(mr, ir, ml, il, bw8, (float*) out, i, batch, h->stream));
    }
    return USC_OK;
}

/* the generator's symbol tables (2n int32: up, down), cached per handle: chirp_orth symbols (carrier == 0) or the I/Q transmitter's */
static int ensure_symbol_table(usc_handle* h, double amp, double carrier = 0.0, double bw = 0.0, int sideband = 0, double phase = 0.0) {
    const uint32_t n = h->cfg.n;
    const double key[4] = {carrier, bw, (double) sideband, phase};
    if (!h->d_sym_table || h->sym_amp != amp || memcmp(key, h->sym_iq, sizeof key) != 0) {
        std::vector<int32_t> tab(2 * (size_t) n);
        if (carrier == 0.0) usc_host_symbol_tables(n, h->cfg.fs, h->cfg.f0, h->cfg.f1, amp, tab.data());
        else usc_host_iq_symbol_tables(n, h->cfg.fs, carrier, bw, sideband, phase, amp, tab.data());
        CK(cudaStreamSynchronize(h->stream));
        cudaFree(h->d_sym_table);
        h->d_sym_table = nullptr;
        int rc = upload(tab.data(), tab.size() * sizeof(int32_t), (void**) &h->d_sym_table);
        if (rc) return rc;
        h->sym_amp = amp;
        memcpy(h->sym_iq, key, sizeof key);
    }
    return USC_OK;
}

int usc_synth_streams(usc_handle* h, uint64_t seed, uint64_t first_stream, uint32_t nstreams, uint32_t nframes,
                      size_t stream_stride, uint32_t lead_in, uint32_t msg_bytes, uint32_t guard, double amp,
                      double noise_sigma, int32_t* pcm, uint32_t* offsets, uint8_t* messages) {
    USC_ENTER(h);
    if (!h || !pcm || !(amp >= 0.0) || !(noise_sigma >= 0.0) || amp + 8.0 * noise_sigma > 8.0e6) return USC_ERR_ARGUMENT;
    const uint32_t n = h->cfg.n;
    if (((uintptr_t) pcm & 7u) != 0 || (stream_stride & 1u) || stream_stride < (size_t) nframes * n) return USC_ERR_ARGUMENT;
    if (msg_bytes > 4096u || lead_in > 65536u || guard > 65536u || (size_t) nframes * (n / 2) > 0xffffffffu) return USC_ERR_ARGUMENT;
    if (!nstreams || !nframes) return USC_OK;
    int rc = ensure_symbol_table(h, amp);
    if (rc) return rc;
    LAUNCHED(h, launch_synth_streams(seed, first_stream, nstreams, nframes, stream_stride, n, lead_in, msg_bytes, guard,
                                     h->d_sym_table, usc_host_noise_gain(noise_sigma), pcm, offsets, messages, h->stream));
    return USC_OK;
}

int usc_synth_frames(usc_handle* h, uint64_t seed, uint64_t first_frame, size_t nframes, double amp, double noise_sigma,
                     int32_t* pcm, uint8_t* bits) {
    USC_ENTER(h);
    if (!h || !pcm || !(amp >= 0.0) || !(noise_sigma >= 0.0) || amp + 8.0 * noise_sigma > 8.0e6) return USC_ERR_ARGUMENT;
    if (((uintptr_t) pcm & 7u) != 0) return USC_ERR_ARGUMENT;
    if (!nframes) return USC_OK;
    const uint32_t n = h->cfg.n;
    int rc = ensure_symbol_table(h, amp);
    if (rc) return rc;
    LAUNCHED(h, launch_synth_frames(seed, first_frame, nframes, n, h->d_sym_table, usc_host_noise_gain(noise_sigma), pcm, bits,
                                    h->stream));
    return USC_OK;
}

int usc_synth_iq_frames(usc_handle* h, uint64_t seed, uint64_t first_frame, size_t nframes, double carrier_hz, double bw_hz,
                        int sideband, double phase_rad, double amp, double noise_sigma, int32_t* pcm, uint8_t* bits) {
    USC_ENTER(h);
    if (!h || !pcm || !(amp >= 0.0) || !(noise_sigma >= 0.0) || amp + 8.0 * noise_sigma > 8.0e6) return USC_ERR_ARGUMENT;
    if (!(carrier_hz > 0.0) || !(bw_hz >= 0.0) || (sideband != 1 && sideband != -1)) return USC_ERR_ARGUMENT;
    if (((uintptr_t) pcm & 7u) != 0) return USC_ERR_ARGUMENT;
    if (!nframes) return USC_OK;
    int rc = ensure_symbol_table(h, amp, carrier_hz, bw_hz, sideband, phase_rad);
    if (rc) return rc;
    LAUNCHED(h, launch_synth_frames(seed, first_frame, nframes, h->cfg.n, h->d_sym_table, usc_host_noise_gain(noise_sigma), pcm, bits,
                                    h->stream));
    return USC_OK;
}

int usc_resample_i16_to_pcm(usc_handle* h, const int16_t* in, size_t n_in, uint32_t up, uint32_t down, int32_t* out,
                            size_t n_out) {
    USC_ENTER(h);
    if (!h || !in || !out || !up || !down || up > 8192u) return USC_ERR_ARGUMENT;
    if (n_out > ((unsigned long long) n_in * up + down - 1) / down) return USC_ERR_ARGUMENT;
    if (!n_out) return USC_OK;
    const uint32_t ktaps = 32;
    if (!h->d_rs_taps || h->rs_up != up) {
        std::vector<float> taps((size_t) up * ktaps);
        usc_host_resample_taps(up, ktaps, taps.data());
        CK(cudaStreamSynchronize(h->stream));
        cudaFree(h->d_rs_taps);
        h->d_rs_taps = nullptr;
        int rc = upload(taps.data(), taps.size() * sizeof(float), (void**) &h->d_rs_taps);
        if (rc) return rc;
        h->rs_up = up;
    }
    LAUNCHED(h, launch_resample_i16(in, n_in, up, down, ktaps, h->d_rs_taps, out, n_out, h->stream));
    return USC_OK;
}

void usc_onoff_default_config(usc_onoff_config* c) {
    if (!c) return;
    c->f1_hz = 17000.0f; c->f2_hz = 18000.0f; c->magnitude_threshold = 3000.0f;
    c->high_frac = 0.1f; c->low_frac = 0.05f;
    c->frame_start = 3; c->frame_bit = 2; c->sync_threshold = 2; c->sampling_offset = 1;
}

void usc_fsk_default_config(usc_fsk_config* c) {
    if (!c) return;
    c->sof_bin = 340; c->eof_bin = 344; c->hex0_bin = 348; c->hex_step = 4; c->tolerance = 0; c->tq_n = 2;
    c->magnitude_threshold = 5000.0f;
}

static int band_common(usc_handle* h, const void* pcm, uint32_t pcm_format, size_t nframes, band_params* p) {
    if (!h || !pcm || pcm_format > USC_PCM_I32 || h->cfg.n != 2048 || !h->d_tw_pass || !h->d_tw_split) return USC_ERR_ARGUMENT;
    if (((uintptr_t) pcm & 15u) != 0) return USC_ERR_ARGUMENT;
    memset(p, 0, sizeof(*p));
    p->pcm = pcm; p->nframes = nframes;
    p->hann = (const float2*) h->d_hann; p->tw_pass = h->d_tw_pass; p->tw_split = h->d_tw_split;
    p->inv_sqrt_n = 1.0f / sqrtf(2048.0f);
    p->fs = h->cfg.fs;
    p->band_lo = 1; p->band_hi = 0;                       /* empty band */
    return USC_OK;
}

int usc_band_magnitudes(usc_handle* h, const void* pcm, uint32_t pcm_format, size_t nframes, float* mag) {
    USC_ENTER(h);
    band_params p;
    int rc = band_common(h, pcm, pcm_format, nframes, &p);
    if (rc) return rc;
    if (!mag) return USC_ERR_ARGUMENT;
    if (!nframes) return USC_OK;
    p.mag = mag;
    LAUNCHED(h, launch_band2048(p, pcm_format, h->num_sms, h->stream));
    return USC_OK;
}

int usc_onoff_detect(usc_handle* h, const void* pcm, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes,
                     const usc_onoff_config* cfg, uint16_t* strength, int8_t* level, uint8_t* chars, uint32_t cap,
                     uint32_t* nchars, uint32_t* sync_errors) {
    USC_ENTER(h);
    band_params p;
    const size_t F = (size_t) nstreams * nframes;
    int rc = band_common(h, pcm, pcm_format, F, &p);
    if (rc) return rc;
    if (!cfg || cfg->frame_bit == 0) return USC_ERR_ARGUMENT;
    if (!F) return USC_OK;
    /* chirp/Src/main.c:372-387: first bins at or above F1 and F1 + 2 (F2 - F1) on the float frequency axis */
    const float freq1 = cfg->f1_hz, freq2 = (float) ((double) cfg->f1_hz + 2.0 * ((double) cfg->f2_hz - (double) cfg->f1_hz));
    uint32_t b1 = 0, b2 = 0;
    for (uint32_t i = 0; i < 1024; ++i) {
        const float f = (float) i * h->cfg.fs / 2048.0f;
        if (b1 == 0 && f >= freq1) b1 = i;
        if (b2 == 0 && f >= freq2) b2 = i;
    }
    if (b1 == 0 || b2 < b1 || b2 >= 512) return USC_ERR_ARGUMENT;
    p.band_lo = b1; p.band_hi = b2;
    p.onoff_threshold = cfg->magnitude_threshold;
    p.thr_high = (uint16_t) ((float) (b2 - b1 + 1) * cfg->high_frac);
    p.thr_low = (uint16_t) ((float) (b2 - b1 + 1) * cfg->low_frac);
    const bool want_decode = chars || nchars || sync_errors;
    int8_t* lv = level;
    if (want_decode && !lv) {
        if ((rc = reserve_work(h, F))) return rc;
        lv = (int8_t*) h->d_work;
    }
    p.strength = strength; p.level = lv;
    LAUNCHED(h, launch_band2048(p, pcm_format, h->num_sms, h->stream));
    if (want_decode)
        LAUNCHED(h, launch_onoff_decode(lv, nstreams, nframes, cfg->frame_start, cfg->frame_bit, cfg->sync_threshold,
                                        cfg->sampling_offset, chars, cap, nchars, sync_errors, h->stream));
    return USC_OK;
}

int usc_fsk_detect(usc_handle* h, const void* pcm, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes,
                   const usc_fsk_config* cfg, uint8_t* code, float* magnitude, float* frequency, uint8_t* chars,
                   uint32_t cap, uint32_t* nchars, uint32_t* nsof, uint32_t* neof) {
    USC_ENTER(h);
    band_params p;
    const size_t F = (size_t) nstreams * nframes;
    int rc = band_common(h, pcm, pcm_format, F, &p);
    if (rc) return rc;
    if (!cfg || cfg->tq_n == 0) return USC_ERR_ARGUMENT;
    const uint32_t lo = cfg->sof_bin < cfg->eof_bin ? cfg->sof_bin : cfg->eof_bin;
    const uint32_t hi_hex = cfg->hex0_bin + 15u * cfg->hex_step;
    uint32_t hi = cfg->sof_bin > cfg->eof_bin ? cfg->sof_bin : cfg->eof_bin;
    if (hi_hex > hi) hi = hi_hex;
    if ((cfg->hex0_bin < lo ? cfg->hex0_bin : lo) < cfg->tolerance || hi + cfg->tolerance >= 512) return USC_ERR_ARGUMENT;
    if (!F) return USC_OK;
    p.sof_bin = cfg->sof_bin; p.eof_bin = cfg->eof_bin; p.hex0_bin = cfg->hex0_bin; p.hex_step = cfg->hex_step;
    p.tolerance = cfg->tolerance; p.fsk_threshold = cfg->magnitude_threshold;
    const bool want_parse = chars || nchars || nsof || neof;
    uint8_t* cd = code;
    if (!cd) {
        if ((rc = reserve_work(h, F))) return rc;
        cd = (uint8_t*) h->d_work;
    }
    p.code = cd; p.code_mag = magnitude; p.code_freq = frequency;
    LAUNCHED(h, launch_band2048(p, pcm_format, h->num_sms, h->stream));
    if (want_parse) LAUNCHED(h, launch_fsk_parse(cd, nstreams, nframes, cfg->tq_n, chars, cap, nchars, nsof, neof, h->stream));
    return USC_OK;
}

int usc_spectrum_analyzer(usc_handle* h, const void* pcm, uint32_t pcm_format, uint32_t nframes, float ac_coupling_hz,
                          float* mag, float* db, float* peak, uint32_t* peak_idx) {
    USC_ENTER(h);
    /* fft() of experiments/basic/Src/main.c:107-142: window, RFFT, magnitude/sqrt(N), AC coupling, dB, arg-max */
    if (!h || !pcm || pcm_format > USC_PCM_I32) return USC_ERR_ARGUMENT;
    const uint32_t n = h->cfg.n;
    if (n > 16384) return USC_ERR_ARGUMENT;
    if (!nframes) return USC_OK;
    const size_t W = (size_t) nframes * n;
    int rc = reserve_work(h, W * sizeof(float));
    if (rc) return rc;
    float* x = h->d_work;
    if (pcm_format == USC_PCM_I32) {
        LAUNCHED(h, launch_i32_to_f32((const int32_t*) pcm, x, W, h->stream));
        LAUNCHED(h, launch_mult(x, n, h->d_hann, 0, x, n, n, nframes, h->stream));
    } else {
        LAUNCHED(h, launch_mult((const float*) pcm, n, h->d_hann, 0, x, n, n, nframes, h->stream));
    }
    fft_plan_dev plan;
    if ((rc = make_plan(h, n / 2, n, &plan))) return rc;
    if ((rc = fft_forward_inplace(h, FFT_R2C, n / 2, plan, x, nframes))) return rc;
    /* fft_frequency[i] = i*fs/N < FFT_AC_COUPLING_HZ (main.c:127,250): count of leading bins forced to 1.0 */
    uint32_t ac_bins = 0;
    while (ac_bins < n / 2 && (float) ac_bins * h->cfg.fs / (float) n < ac_coupling_hz) ++ac_bins;
    LAUNCHED(h, launch_spectrum_tail(x, n, 1.0f / sqrtf((float) n), ac_bins, mag, db, peak, peak_idx, nframes, h->stream));
    return USC_OK;
}

int usc_iq_init(usc_handle* h, float carrier_hz, float bw_hz, const float* fir_coeffs_host, uint32_t num_taps,
                uint32_t window_bins) {
    USC_ENTER(h);
    if (!h || !fir_coeffs_host || num_taps < 1 || num_taps > 64) return USC_ERR_ARGUMENT;
    const uint32_t n = h->cfg.n, half = n / 2;
    if (n < 64 || n > 4096 || window_bins < 1 || window_bins > half / 2) return USC_ERR_ARGUMENT;
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_iq_cos); cudaFree(h->d_iq_sin); cudaFree(h->d_iq_chirp); cudaFree(h->d_iq_conj); cudaFree(h->d_iq_hann); cudaFree(h->d_iq_taps);
    h->d_iq_cos = h->d_iq_sin = h->d_iq_chirp = h->d_iq_conj = h->d_iq_hann = h->d_iq_taps = nullptr;
    h->iq_cos.resize(n); h->iq_sin.resize(n); h->iq_chirp.resize(n); h->iq_hann.resize(half);
    /* carrier: experiments/iq_modulation/Src/iq_modem.c:34-46 (degrees, float-accumulated t) */
    const float sweep_T = h->cfg.sweep_T, fs = h->cfg.fs;
    float dt = sweep_T / (sweep_T * fs), t = 0.0f;
    for (uint32_t i = 0; i < n; ++i) {
        float theta = (float) (360.0 * (double) carrier_hz * (double) t);
        usc_host_arm_sin_cos_f32(theta, &h->iq_sin[i], &h->iq_cos[i]);
        t = t + dt;
    }
    /* baseband chirp -bw/2..+bw/2 over one frame at fs/2 (simulation/IQ_modulation.ipynb cells 2-3, 10) */
    usc_host_ref_chirp(USC_CHIRP_S, half, fs / 2.0f, -bw_hz / 2.0f, bw_hz / 2.0f, (float) n / fs, 0.0f, 1, h->iq_chirp.data());
    std::vector<float> conj(n);
    for (uint32_t m = 0; m < half; ++m) { conj[2 * m] = h->iq_chirp[2 * m]; conj[2 * m + 1] = -h->iq_chirp[2 * m + 1]; }
    usc_host_hann(h->iq_hann.data(), half, USC_HANN_PERIODIC);
    int rc;
    if ((rc = upload(h->iq_cos.data(), n * 4, (void**) &h->d_iq_cos))) return rc;
    if ((rc = upload(h->iq_sin.data(), n * 4, (void**) &h->d_iq_sin))) return rc;
    if ((rc = upload(h->iq_chirp.data(), n * 4, (void**) &h->d_iq_chirp))) return rc;
    if ((rc = upload(conj.data(), n * 4, (void**) &h->d_iq_conj))) return rc;
    if ((rc = upload(h->iq_hann.data(), half * 4, (void**) &h->d_iq_hann))) return rc;
    if ((rc = upload(fir_coeffs_host, num_taps * 4, (void**) &h->d_iq_taps))) return rc;
    h->iq_ntaps = num_taps;
    h->iq_window = window_bins;
    return USC_OK;
}

int usc_iq_demod(usc_handle* h, const void* pcm, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes,
                 size_t stream_stride, float* mag_up, uint32_t* idx_up, float* mag_down, uint32_t* idx_down,
                 uint8_t* bit) {
    USC_ENTER(h);
    if (!h || !pcm || pcm_format > USC_PCM_I32 || !h->d_iq_taps) return USC_ERR_ARGUMENT;
    const uint32_t n = h->cfg.n, half = n / 2, W = h->iq_window;
    if (stream_stride < (size_t) nframes * n) return USC_ERR_ARGUMENT;
    const size_t F = (size_t) nstreams * nframes;
    if (!F) return USC_OK;
    if (F > 0xffffffffu) return USC_ERR_ARGUMENT;
    if (n == 2048 && W <= 32 && h->iq_ntaps <= 32 && h->d_tw_pass && !h->dbg_iq_unfused) {
        /* whole path in one kernel: mix, FIR, de-chirp both ways, Hann, FFT, windowed peaks (k_iq.cu) */
        LAUNCHED(h, launch_iq_fused(pcm, pcm_format, nstreams, nframes, stream_stride, h->d_iq_cos, h->d_iq_sin, h->d_iq_taps,
                                    h->iq_ntaps, h->d_iq_chirp, h->d_iq_hann, h->d_tw_pass, W, mag_up, idx_up, mag_down,
                                    idx_down, bit, h->num_sms, h->stream));
        return USC_OK;
    }
    /* scratch: R (F*n) | P (F*n) | mags (F*half) | 4 result vectors + 2 spare magnitude vectors */
    int rc = reserve_work(h, (2 * F * n + F * half + 8 * F) * sizeof(float));
    if (rc) return rc;
    float *R = h->d_work, *P = R + F * n, *M = P + F * n;
    float *mr = M + F * half, *ml = mr + F, *spare_u = ml + F, *spare_d = spare_u + F;
    uint32_t *ir = (uint32_t*) (spare_d + F), *il = ir + F;
    fft_plan_dev plan;
    if ((rc = make_plan(h, half, half, &plan))) return rc;
    LAUNCHED(h, launch_iq_frontend(pcm, pcm_format, nstreams, nframes, stream_stride, n, h->d_iq_cos, h->d_iq_sin,
                                   h->d_iq_taps, h->iq_ntaps, R, h->stream));
    if (n == 2048 && W <= 32 && h->d_tw_pass) {
        /* fused back end: both hypotheses in the packed 32x32 core, one warp per frame */
        LAUNCHED(h, launch_iq_backend(R, F, h->d_iq_chirp, h->d_iq_hann, h->d_tw_pass, W, mag_up, idx_up, mag_down, idx_down,
                                      bit, h->num_sms, h->stream));
        return USC_OK;
    }
    float* mags[2] = {mag_up ? mag_up : spare_u, mag_down ? mag_down : spare_d};
    uint32_t* idxs[2] = {idx_up, idx_down};
    const float* refs[2] = {h->d_iq_conj, h->d_iq_chirp};       /* up: R x conj(chirp); down: R x chirp */
    for (int hyp = 0; hyp < 2; ++hyp) {
        LAUNCHED(h, launch_cmul(R, n, refs[hyp], 0, P, n, half, (uint32_t) F, h->stream));
        LAUNCHED(h, launch_cmul_real(P, n, h->d_iq_hann, 0, P, n, half, (uint32_t) F, h->stream));
        if ((rc = fft_forward_inplace(h, FFT_C2C_FWD, half, plan, P, (uint32_t) F))) return rc;
        LAUNCHED(h, launch_cmag(P, n, M, half, half, (uint32_t) F, h->stream));
        LAUNCHED(h, launch_max(M, half, W, mr, ir, (uint32_t) F, h->stream));
        LAUNCHED(h, launch_max(M + (half - W), half, W, ml, il, (uint32_t) F, h->stream));
        LAUNCHED(h, launch_iq_pick(mr, ir, ml, il, half - W, mags[hyp], idxs[hyp], F, h->stream));
    }
    if (bit) LAUNCHED(h, launch_decide(mags[0], mags[1], bit, F, h->stream));
    return USC_OK;
}

int usc_pipeline(usc_handle* h, const float* frames, float* mags, int updown, uint32_t batch) {
    USC_ENTER(h);
    /* operator-by-operator form of receiver/Src/main.c:163-180 (full spectrum wanted, so nothing
     * to prune): mult, mult, rfft, mag into the lower half, zeros above (hazard H1 defined). */
    if (!h || !frames || !mags) return USC_ERR_ARGUMENT;
    if (h->cfg.chirp_variant == USC_CHIRP_S) return USC_ERR_ARGUMENT;
    if (!batch) return USC_OK;
    const uint32_t n = h->cfg.n;
    int rc;
    if ((rc = usc_arm_mult_f32_batch(h, frames, n, updown ? h->d_up : h->d_down, 0, mags, n, n, batch))) return rc;
    if ((rc = usc_arm_mult_f32_batch(h, mags, n, h->d_hann, 0, mags, n, n, batch))) return rc;
    if ((rc = usc_arm_rfft_fast_f32_batch(h, n, mags, mags, 0, batch))) return rc;
    LAUNCHED(h, launch_pipeline_tail(mags, n, batch, 1, h->stream));
    return USC_OK;
}

int usc_compress_chirp(usc_handle* h, const void* pcm, uint32_t pcm_format, size_t nframes, int use_up,
                       float* out_frames, float* max_val, uint32_t* max_idx) {
    USC_ENTER(h);
    if (!h || !pcm || pcm_format > USC_PCM_I32) return USC_ERR_ARGUMENT;
    if (h->cfg.n != 2048 || h->cfg.chirp_variant != USC_CHIRP_T || !h->d_H_up) return USC_ERR_ARGUMENT;
    if (((uintptr_t) pcm & 7u) != 0 || ((uintptr_t) out_frames & 7u) != 0) return USC_ERR_ARGUMENT;
    if (!nframes) return USC_OK;
    LAUNCHED(h, launch_compress2048(pcm, pcm_format, nframes, (const float2*) h->d_hann,
                                    (const float2*) (use_up ? h->d_H_up : h->d_H_down), h->d_tw_pass,
                                    h->d_tw_split, out_frames, max_val, max_idx, h->num_sms, h->stream));
    return USC_OK;
}

// tables of the overlap-save synchroniser, built on first use: G = rfft_4096(window * chirp, zero-padded) for both chirps
// (the canonical FFT on the device, as init_ref_chirp's H), the W_4096 split table and the W_2048 radix-2 twiddles
static int ensure_os_tables(usc_handle* h) {
    if (h->d_G_up) return USC_OK;
    const uint32_t n = h->cfg.n;
    fft_plan_dev plan;
    int rc = make_plan(h, n, 2 * n, &plan);
    if (rc) return rc;
    float2* tw2048 = nullptr;
    if ((rc = get_twiddles(h, n, &tw2048))) return rc;
    const std::vector<float>& tw = h->tw_host[2 * n];
    std::vector<float> split(2 * (size_t) n);
    for (uint32_t k = 0; k < n; ++k) {
        split[2 * k] = tw[2 * k];
        split[2 * k + 1] = -tw[2 * k + 1];
    }
    if ((rc = upload(split.data(), split.size() * 4, (void**) &h->d_tw_split4096))) return rc;
    std::vector<float> gu(2 * (size_t) n, 0.0f), gd(2 * (size_t) n, 0.0f);
    for (uint32_t i = 0; i < n; ++i) { gu[i] = h->up[i] * h->hann[i]; gd[i] = h->down[i] * h->hann[i]; }
    float *du = nullptr, *dd = nullptr;
    if ((rc = upload(gu.data(), gu.size() * 4, (void**) &du))) return rc;
    if ((rc = upload(gd.data(), gd.size() * 4, (void**) &dd))) { cudaFree(du); return rc; }
    cudaError_t e;
    if ((e = launch_fft_generic(FFT_R2C, plan, du, du, 1, h->stream)) != cudaSuccess ||
        (e = launch_fft_generic(FFT_R2C, plan, dd, dd, 1, h->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(h->stream)) != cudaSuccess) { cudaFree(du); cudaFree(dd); return cuda_rc(e); }
    h->d_G_up = (float2*) du;
    h->d_G_down = (float2*) dd;
    return USC_OK;
}

int usc_correlate_os(usc_handle* h, const void* pcm, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes,
                     size_t stream_stride, int use_up, float* out, float* max_val, uint32_t* max_idx) {
    USC_ENTER(h);
    if (!h || !pcm || pcm_format > USC_PCM_I32) return USC_ERR_ARGUMENT;
    if (h->cfg.n != 2048 || h->cfg.chirp_variant == USC_CHIRP_S || !h->d_tw_pass) return USC_ERR_ARGUMENT;
    if (((uintptr_t) pcm & 15u) != 0 || (stream_stride & 3u) != 0 || ((uintptr_t) out & 15u) != 0) return USC_ERR_ARGUMENT;
    if (nstreams && nframes >= 2 && stream_stride < (size_t) nframes * 2048) return USC_ERR_ARGUMENT;
    if (!nstreams || nframes < 2) return USC_OK;
    int rc = ensure_os_tables(h);
    if (rc) return rc;
    LAUNCHED(h, launch_correlate_os(pcm, pcm_format, nstreams, nframes, stream_stride, use_up ? h->d_G_up : h->d_G_down,
                                    h->d_tw_pass, h->tw_cache[2048], h->d_tw_split4096, out, max_val, max_idx, h->num_sms,
                                    h->stream));
    return USC_OK;
}

}  // extern "C"
