// k_receiver.cu — K4 (sliding-correlation sync search, optional synchronous addition) and K7 (the
// receiver's whole main-loop state machine) for N = 2048.
//
// K7 follows receiver/Src/main.c:417-580 literally, one WARP per stream:
//   FIFO        frames t-2, t-1, t of the stream (zeros before it starts)          main.c:659-668
//   IDLE/SYNCHRONIZING   4 x dsp(UP) at N/2 + turn*N/8 + i*N/4, decision every 2nd frame over the
//               8 offsets, noise floor mag_stat[12] (starts at 1e37), 3 in a row -> SYNCHRONIZED   :428-488
//   SYNCHRONIZED / DATA_RECEIVING   symbol_snr(UP), symbol_snr(DOWN), resync +-N/8, bits MSB first,
//               byte every 8 bits, '\n' at the end of a message                    :491-550, 233-273
// Every dsp() is the fused chain of K1 for one hypothesis (de-chirp, Hann, 2048-pt RFFT, magnitude,
// arg-max over [0, bandwidth2)); the control flow is warp-uniform (one stream per warp), the FFT
// uses all 32 lanes.  Hazards are defined as in the oracle: H1 left window = zeros, H3/H5 probes
// outside [0, 2N] give snr = -inf, H4 history starts zeroed.
#include "usc_kernels.cuh"
#include "usc_launch.h"
#include "usc_warpfft.cuh"

namespace usc {

constexpr int kRxWarps = 4;
constexpr int kRxNB = 5;
constexpr int kRxSmem = (4096 + kRxWarps * kTileFloat2) * (int) sizeof(float2);

struct rx_tables {
    const float2* up;       // shared-memory copies
    const float2* down;
    const float2* hann;
    const float2* tw;
};

// dsp() for one hypothesis on the FIFO of frame t at sync_position pos: returns mag_max (the right
// window always wins in the receiver variant, hazard H1) and its bin.
template <typename PCM>
__device__ __forceinline__ void dsp_fifo(const PCM* __restrict__ stream, int64_t nsamples, int64_t g0,
                                         const rx_tables& tb, bool up, float2* tile, const float2 (&ws)[kRxNB],
                                         int lane, uint32_t bw2, float& mag, uint32_t& idx) {
    using V2 = typename vec2<PCM>::type;
    const float2* chirp = up ? tb.up : tb.down;
    float re[32], im[32];
#pragma unroll
    for (int b = 0; b < 32; ++b) {
        const int m = lane + 32 * b;
        const int64_t g = g0 + 2 * m;                        // g0 is a multiple of N/8: pairs never straddle 0
        float x0 = 0.0f, x1 = 0.0f;
        if (g >= 0 && g + 1 < nsamples) {
            V2 raw = *reinterpret_cast<const V2*>(stream + g);
            x0 = pcm_to_float(raw.x);
            x1 = pcm_to_float(raw.y);
        }
        float2 c = chirp[m], w = tb.hann[m];
        re[b] = __fmul_rn(__fmul_rn(x0, c.x), w.x);
        im[b] = __fmul_rn(__fmul_rn(x1, c.y), w.y);
    }
    fft1024_warp(re, im, tile, tb.tw, lane);
    peak_window<kRxNB>(re, im, ws, lane, bw2, mag, idx);
}

struct rx_params {
    const void* pcm; uint32_t nstreams; uint32_t nframes; size_t stream_stride;
    const float2* up; const float2* down; const float2* hann; const float2* tw_pass; const float2* tw_split;
    uint32_t bandwidth2; float snr_threshold;
    uint8_t* uart; uint32_t uart_cap; rx_result_rec* results;
    // sync search
    uint32_t sync_add; float* ss_mag; uint32_t* ss_idx;
};

__device__ __forceinline__ void load_tables(const rx_params& p, float2* s_up, float2* s_down, float2* s_hann, float2* s_tw) {
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        s_up[i] = p.up[i];
        s_down[i] = p.down[i];
        s_hann[i] = p.hann[i];
        s_tw[i] = p.tw_pass[i];
    }
}

template <typename PCM>
__global__ void __launch_bounds__(kRxWarps * 32, 3) k_receiver_run(rx_params p) {
    extern __shared__ float2 s_rx[];
    float2 *s_up = s_rx, *s_down = s_rx + 1024, *s_hann = s_rx + 2048, *s_tw = s_rx + 3072;
    float2* s_tile_base = s_rx + 4096;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    load_tables(p, s_up, s_down, s_hann, s_tw);
    float2 ws[kRxNB];
#pragma unroll
    for (int d1 = 0; d1 < kRxNB; ++d1) ws[d1] = p.tw_split[lane + 32 * d1];
    __syncthreads();
    const rx_tables tb{s_up, s_down, s_hann, s_tw};
    float2* tile = s_tile_base + warp * kTileFloat2;
    const uint32_t N = 2048, offset = N / 8, shift = N / 4;
    const float thr = p.snr_threshold;
    const uint32_t bw2 = p.bandwidth2;

    const uint32_t nwarps = gridDim.x * kRxWarps;
    for (uint32_t s = blockIdx.x * kRxWarps + warp; s < p.nstreams; s += nwarps) {
        const PCM* stream = static_cast<const PCM*>(p.pcm) + (size_t) s * p.stream_stride;
        const int64_t nsamples = (int64_t) p.nframes * N;
        uint8_t* uart = p.uart ? p.uart + (size_t) s * p.uart_cap : nullptr;
        // state (main.c:311-339), identical in every lane
        uint32_t state = 0, turn = 0, sync_cnt = 0, pos = N / 2, max_idx = 0, msg = 0, msg_cnt = 0, nout = 0;
        int32_t lock_frame = -1;
        uint32_t lock_pos = 0;
        float mag_mean = 0.0f;
        float mag_stat[12], hmag[8], hmean[4];
#pragma unroll
        for (int i = 0; i < 12; ++i) mag_stat[i] = 1E37f;
#pragma unroll
        for (int i = 0; i < 8; ++i) hmag[i] = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i) hmean[i] = 0.0f;

        auto emit = [&](uint32_t c) {
            if (lane == 0 && uart && nout < p.uart_cap) uart[nout] = (uint8_t) c;
            nout++;
        };
        // symbol_snr (main.c:233-236): dsp into history slot `slot` (< 4) using the slot's own mag_mean
        auto symbol_snr = [&](int64_t fifo0, int64_t q, int slot, bool up) -> float {
            if (q < 0 || q > (int64_t) 2 * N) return -INFINITY;          // hazards H3/H5 defined
            float m;
            uint32_t k;
            dsp_fifo<PCM>(stream, nsamples, fifo0 + q, tb, up, tile, ws, lane, bw2, m, k);
            float mean = 0.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) mean = slot == i ? hmean[i] : mean;
#pragma unroll
            for (int i = 0; i < 4; ++i) hmag[i] = slot == i ? m : hmag[i];
            return __fdiv_rn(__fsub_rn(m, mean), mean);                   // main.c:229
        };
        // resync (main.c:243-273)
        auto resync = [&](int64_t fifo0, float snr, bool up) {
            const int64_t l = (int64_t) pos - offset, r = (int64_t) pos + offset;
            const float snr_l = symbol_snr(fifo0, l, 2, up);
            const float snr_r = symbol_snr(fifo0, r, 3, up);
            if ((snr > snr_l) && (snr > snr_r)) {
            } else if (snr_l >= snr_r) {
                if (l >= 0) pos = (uint32_t) l;
            } else if (snr_l < snr_r) {
                if (r <= (int64_t) 2 * N) pos = (uint32_t) r;
            }
        };

        for (uint32_t t = 0; t < p.nframes; ++t) {
            const int64_t fifo0 = ((int64_t) t - 2) * N;                  // stream index of fifo_queue[0]
            const uint32_t prev = state;
            if (state == 0 || state == 1) {
                if (state == 0) {                                         // IDLE (main.c:428-434)
                    sync_cnt = 0;
                    float sum = 0.0f;
#pragma unroll
                    for (int i = 4; i < 12; ++i) sum = __fadd_rn(sum, mag_stat[i]);
                    mag_mean = __fdiv_rn(sum, 8.0f);
                }
                for (uint32_t i = 0; i < 4; ++i) {                        // main.c:447-451
                    pos = N / 2 + turn * offset + shift * i;
                    float m;
                    uint32_t k;
                    dsp_fifo<PCM>(stream, nsamples, fifo0 + pos, tb, true, tile, ws, lane, bw2, m, k);
                    const uint32_t slot = i * 2 + turn;
#pragma unroll
                    for (int j = 0; j < 8; ++j) hmag[j] = slot == (uint32_t) j ? m : hmag[j];
#pragma unroll
                    for (int j = 0; j < 4; ++j) hmean[j] = slot == (uint32_t) j ? mag_mean : hmean[j];
                }
                turn ^= 1u;
                if (turn == 1u) {
#pragma unroll
                    for (int i = 10; i >= 0; --i) mag_stat[i + 1] = mag_stat[i];     // main.c:458-460
                    float mmm = 0.0f;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (hmag[i] > mmm) { mmm = hmag[i]; max_idx = i; }           // main.c:463-471
                    mag_stat[0] = mmm;
                    const float snr = __fdiv_rn(__fsub_rn(mmm, mag_mean), mag_mean);  // main.c:477
                    if (snr >= thr) {
                        state = 1;
                        if (++sync_cnt >= 3) {
                            state = 2;
                            pos = N / 2 + max_idx * offset;                           // main.c:483
                        }
                    } else {
                        state = 0;
                    }
                }
            } else {
                const float up = symbol_snr(fifo0, pos, 0, true);         // main.c:493-494 / 518-519
                const float down = symbol_snr(fifo0, pos, 1, false);
                if (up >= thr || down >= thr) {
                    const bool is_down = down > up;
                    if (state == 3) msg = ((msg << 1) + (is_down ? 0u : 1u)) & 0xffu;   // main.c:525,529
                    resync(fifo0, is_down ? down : up, !is_down);
                    if (state == 2) {
                        if (is_down) state = 3;                           // the delimiter (main.c:500)
                    } else if (++msg_cnt >= 8) {                          // main.c:532-537
                        emit(msg);
                        msg = 0;
                        msg_cnt = 0;
                    }
                } else {
                    if (state == 3) {                                     // end of message (main.c:539-549)
                        emit((uint32_t) '\n');
                        msg = 0;
                        msg_cnt = 0;
                    }
                    state = 0;
                }
            }
            if (prev != 2 && state == 2 && lock_frame < 0) {
                lock_frame = (int32_t) t;
                lock_pos = pos;
            }
        }
        if (lane == 0 && p.results) {
            rx_result_rec r;
            r.state = state; r.sync_position = pos; r.lock_frame = lock_frame; r.lock_position = lock_pos;
            r.nbytes = nout; r.frames_seen = p.nframes; r.turn = turn; r.sync_cnt = sync_cnt;
            p.results[s] = r;
        }
    }
}

// K4: the search grid of main.c:447-451 for every (stream, frame), after optional synchronous
// addition of sync_add frame-aligned FIFOs (oldest first).  One warp per (stream, frame).
template <typename PCM>
__global__ void __launch_bounds__(kRxWarps * 32, 3) k_sync_search(rx_params p) {
    extern __shared__ float2 s_rx[];
    float2 *s_up = s_rx, *s_down = s_rx + 1024, *s_hann = s_rx + 2048, *s_tw = s_rx + 3072;
    float2* s_tile_base = s_rx + 4096;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    load_tables(p, s_up, s_down, s_hann, s_tw);
    float2 ws[kRxNB];
#pragma unroll
    for (int d1 = 0; d1 < kRxNB; ++d1) ws[d1] = p.tw_split[lane + 32 * d1];
    __syncthreads();
    float2* tile = s_tile_base + warp * kTileFloat2;
    using V2 = typename vec2<PCM>::type;
    const uint32_t N = 2048, offset = N / 8, shift = N / 4;
    const size_t total = (size_t) p.nstreams * p.nframes;
    const size_t nwarps = (size_t) gridDim.x * kRxWarps;
    for (size_t w = (size_t) blockIdx.x * kRxWarps + warp; w < total; w += nwarps) {
        const uint32_t s = (uint32_t) (w / p.nframes), t = (uint32_t) (w - (size_t) s * p.nframes);
        const PCM* stream = static_cast<const PCM*>(p.pcm) + (size_t) s * p.stream_stride;
        const int64_t nsamples = (int64_t) p.nframes * N;
        for (uint32_t i = 0; i < 4; ++i) {
            const uint32_t pos = N / 2 + (t & 1u) * offset + shift * i;
            float re[32], im[32];
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                const int m = lane + 32 * b;
                float x0 = 0.0f, x1 = 0.0f;
                for (uint32_t j = p.sync_add; j-- > 0;) {                 // oldest FIFO first
                    const int64_t g = ((int64_t) t - (int64_t) j - 2) * N + pos + 2 * m;
                    float v0 = 0.0f, v1 = 0.0f;
                    if (g >= 0 && g + 1 < nsamples) {
                        V2 raw = *reinterpret_cast<const V2*>(stream + g);
                        v0 = pcm_to_float(raw.x);
                        v1 = pcm_to_float(raw.y);
                    }
                    if (j == p.sync_add - 1) { x0 = v0; x1 = v1; }
                    else { x0 = __fadd_rn(x0, v0); x1 = __fadd_rn(x1, v1); }
                }
                float2 c = s_up[m], wn = s_hann[m];
                re[b] = __fmul_rn(__fmul_rn(x0, c.x), wn.x);
                im[b] = __fmul_rn(__fmul_rn(x1, c.y), wn.y);
            }
            fft1024_warp(re, im, tile, s_tw, lane);
            float mag;
            uint32_t idx;
            peak_window<kRxNB>(re, im, ws, lane, p.bandwidth2, mag, idx);
            if (lane == 0) {
                p.ss_mag[w * 4 + i] = mag;
                p.ss_idx[w * 4 + i] = idx;
            }
        }
    }
}

static rx_params make_params(const rx_launch& a) {
    rx_params p{};
    p.pcm = a.pcm; p.nstreams = a.nstreams; p.nframes = a.nframes; p.stream_stride = a.stream_stride;
    p.up = a.up; p.down = a.down; p.hann = a.hann; p.tw_pass = a.tw_pass; p.tw_split = a.tw_split;
    p.bandwidth2 = a.bandwidth2; p.snr_threshold = a.snr_threshold;
    p.uart = a.uart; p.uart_cap = a.uart_cap; p.results = a.results;
    p.sync_add = a.sync_add < 1 ? 1 : a.sync_add; p.ss_mag = a.ss_mag; p.ss_idx = a.ss_idx;
    return p;
}

static cudaError_t rx_prepare() {
    static bool done = false;
    if (done) return cudaSuccess;
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_receiver_run<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRxSmem))) return e;
    if ((e = cudaFuncSetAttribute(k_receiver_run<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRxSmem))) return e;
    if ((e = cudaFuncSetAttribute(k_sync_search<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRxSmem))) return e;
    if ((e = cudaFuncSetAttribute(k_sync_search<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRxSmem))) return e;
    done = true;
    return cudaSuccess;
}

cudaError_t launch_receiver_run(const rx_launch& a, int num_sms, cudaStream_t st) {
    rx_params p = make_params(a);
    size_t ctas = ((size_t) a.nstreams + kRxWarps - 1) / kRxWarps;
    const size_t cap = (size_t) num_sms * 3;
    if (ctas > cap) ctas = cap;
    cudaError_t e = rx_prepare();
    if (e != cudaSuccess) return e;
    if (a.pcm_format == 1u) k_receiver_run<int32_t><<<(int) ctas, kRxWarps * 32, kRxSmem, st>>>(p);
    else k_receiver_run<float><<<(int) ctas, kRxWarps * 32, kRxSmem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_sync_search(const rx_launch& a, int num_sms, cudaStream_t st) {
    rx_params p = make_params(a);
    size_t ctas = ((size_t) a.nstreams * a.nframes + kRxWarps - 1) / kRxWarps;
    const size_t cap = (size_t) num_sms * 3 * 2;
    if (ctas > cap) ctas = cap;
    cudaError_t e = rx_prepare();
    if (e != cudaSuccess) return e;
    if (a.pcm_format == 1u) k_sync_search<int32_t><<<(int) ctas, kRxWarps * 32, kRxSmem, st>>>(p);
    else k_sync_search<float><<<(int) ctas, kRxWarps * 32, kRxSmem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace usc
