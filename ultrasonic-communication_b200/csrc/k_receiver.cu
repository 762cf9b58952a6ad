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

constexpr int kRxNB = 5;
#ifndef USC_RX_WARPS
#define USC_RX_WARPS 8                                  // K7 (12 = three warps per scheduler at 168 registers: 139 spilled words, 23.3 against 19.9 ms)
#endif
#ifndef USC_SS_WARPS
#define USC_SS_WARPS 8                                  // K4 without synchronous addition (12: gathered instead of TMA-staged windows)
#endif
// shared memory (float2 units): per-warp 8 KB tile | per-warp state (128 B) | TMEM slot | [MULTI] per-warp sum of the window union
template <int W> struct rx_smem {                       // W warps per CTA
    static constexpr int tile_f2 = kTileFloat2;         // padded 32 x 33 float2 tile per warp (8448 bytes)
    static constexpr int tile = 0, state = W * tile_f2, slot = state + W * 16, bar = slot + 2, sum = bar + W + (W & 1);
    static constexpr int bytes = sum * 8;
    // + per warp 14 KB for the union of a frame's four search windows (1.75 N samples): the staged PCM of K4 (filled by one
    // TMA bulk copy per work item, one item ahead), or its K-frame sum (synchronous addition)
    static constexpr int bytes_union = bytes + W * 1792 * 8;
};

// The four tables every dsp() reads — up chirp, down chirp, Hann, inter-pass twiddles — live in tensor memory, one row per
// lane (usc_tmem.cuh; K1 keeps them the same way): columns 4 b .. 4 b + 3 = (up[2m], down[2m], up[2m+1], down[2m+1]) of
// m = lane + 32 b | 128 + 2 b = Hann pair | 192 + 2 d = W_1024^(lane d).
constexpr int kRxTud = 0, kRxThann = 128, kRxTtw = 192, kRxTcols = 256;
struct rx_tables {
    uint32_t tq;            // this warp's TMEM lane quadrant
    float one;              // 1.0f read from a table (first butterfly stage as FMAs by 1.0, usc_arith.cuh)
};
enum rx_chirps : int { RX_UP_UP = 0, RX_UP_DOWN = 1, RX_DOWN_DOWN = 2 };   // de-chirp tables of the two halves of a packed pass

struct rx_params {
    const void* pcm; uint32_t nstreams; uint32_t nframes; size_t stream_stride;
    const float2* up; const float2* down; const float2* hann; const float2* tw_pass; const float2* tw_split;
    uint32_t bandwidth2; float snr_threshold;
    uint8_t* uart; uint32_t uart_cap; rx_result_rec* results;
    // sync search
    uint32_t sync_add; float* ss_mag; uint32_t* ss_idx;
    // chunked operation of K7: `carry` frames of history precede the nframes new ones; state is loaded / stored
    uint32_t carry; rx_state_rec* rx_state;
    uint32_t staged;        // K4: stream bases are 16-byte aligned, so window unions can arrive by TMA bulk copies
};

template <typename PCM>
__device__ __forceinline__ float2 load_pair(const PCM* __restrict__ stream, int64_t nsamples, int64_t g, int64_t gmin = 0) {
    using V2 = typename vec2<PCM>::type;
    if (g >= gmin && g + 1 < nsamples) {                  // g is even (window starts are multiples of N/8)
        const V2 raw = *reinterpret_cast<const V2*>(stream + g);
        return make_float2(pcm_to_float(raw.x), pcm_to_float(raw.y));
    }
    return make_float2(0.0f, 0.0f);                    // before the stream starts the FIFO holds zeros
}

// (x * c) * w on both halves, packed; the products meet FMAs by 1.0 in the first butterfly stage (usc_arith.cuh)
template <int MODE>
__device__ __forceinline__ void rx_front(float2 (&re)[32], float2 (&im)[32], uint32_t tq) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {                                     // table values of four rows per TMEM round trip
        uint32_t c[16], w[8];
        ldtm16_8(tq + kRxTud + 16 * g, c, tq + kRxThann + 8 * g, w);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int b = 4 * g + j;
            const float u0 = __uint_as_float(c[4 * j]), d0 = __uint_as_float(c[4 * j + 1]);
            const float u1 = __uint_as_float(c[4 * j + 2]), d1 = __uint_as_float(c[4 * j + 3]);
            const float2 c0 = MODE == RX_UP_UP ? bc2(u0) : (MODE == RX_UP_DOWN ? make_float2(u0, d0) : bc2(d0));
            const float2 c1 = MODE == RX_UP_UP ? bc2(u1) : (MODE == RX_UP_DOWN ? make_float2(u1, d1) : bc2(d1));
            re[b] = __fmul2_rn(__fmul2_rn(re[b], c0), bc2(__uint_as_float(w[2 * j])));
            im[b] = __fmul2_rn(__fmul2_rn(im[b], c1), bc2(__uint_as_float(w[2 * j + 1])));
        }
    }
}

// second half of dsp(): de-chirp, Hann, FFT, right-window peaks, on samples already in (re, im) = (x[2m], x[2m+1])
__device__ __forceinline__ void dsp_pair_tail(float2 (&re)[32], float2 (&im)[32], int chirps, const rx_tables& tb, float2* tile,
                                              const float2 (&ws)[kRxNB], int lane,
                                              uint32_t bw2, float& magA, uint32_t& idxA, float& magB, uint32_t& idxB) {
    if (chirps == RX_UP_UP) rx_front<RX_UP_UP>(re, im, tb.tq);        // warp-uniform
    else if (chirps == RX_UP_DOWN) rx_front<RX_UP_DOWN>(re, im, tb.tq);
    else rx_front<RX_DOWN_DOWN>(re, im, tb.tq);
    fft1024_pair_tm<true, true>(re, im, tile, tb.tq + kRxTtw, tb.one, lane);
    peak_window_pair<kRxNB>(re, im, ws, lane, bw2, magA, idxA, magB, idxB);
}

// TWO dsp() calls at once (halves .x / .y of the packed core): windows starting at stream samples gA
// and gB, de-chirped by chirpA / chirpB (the same table for two offsets of one hypothesis, the up
// and down tables for the two hypotheses of one position).  Returns the right-window peaks (the
// right window always wins in the receiver variant, hazard H1).  sync_add > 1 sums that many
// frame-aligned windows, oldest first, before the de-chirp (synchronous addition).
template <typename PCM, bool MULTI>
__device__ __forceinline__ void dsp_pair(const PCM* __restrict__ stream, int64_t nsamples, int64_t gA, int64_t gB,
                                         int chirps, const rx_tables& tb,
                                         uint32_t sync_add, float2* tile, const float2 (&ws)[kRxNB], int lane,
                                         uint32_t bw2, float& magA, uint32_t& idxA, float& magB, uint32_t& idxB,
                                         int64_t gmin = 0) {
    float2 re[32], im[32];
    if (!MULTI) {
        if (gA >= gmin && gB >= gmin && gA + 2048 <= nsamples && gB + 2048 <= nsamples) {    // both windows inside the stream
            using V2 = typename vec2<PCM>::type;
            const V2* pa = reinterpret_cast<const V2*>(stream + gA) + lane;
            const V2* pb = reinterpret_cast<const V2*>(stream + gB) + lane;
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                const V2 ra = pa[32 * b], rb = pb[32 * b];
                re[b] = make_float2(pcm_to_float(ra.x), pcm_to_float(rb.x));
                im[b] = make_float2(pcm_to_float(ra.y), pcm_to_float(rb.y));
            }
        } else {
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                const int m = lane + 32 * b;
                const float2 xa = load_pair<PCM>(stream, nsamples, gA + 2 * m, gmin), xb = load_pair<PCM>(stream, nsamples, gB + 2 * m, gmin);
                re[b] = make_float2(xa.x, xb.x);
                im[b] = make_float2(xa.y, xb.y);
            }
        }
    } else if (gA - (int64_t) (sync_add - 1) * 2048 >= 0 && gB - (int64_t) (sync_add - 1) * 2048 >= 0 &&
               gA + 2048 <= nsamples && gB + 2048 <= nsamples) {   // every summed window inside the stream: unchecked loads
        using V2 = typename vec2<PCM>::type;
        {
            const int64_t back = (int64_t) (sync_add - 1) * 2048;          // oldest FIFO first
            const V2* pa = reinterpret_cast<const V2*>(stream + gA - back) + lane;
            const V2* pb = reinterpret_cast<const V2*>(stream + gB - back) + lane;
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                const V2 ra = pa[32 * b], rb = pb[32 * b];
                re[b] = make_float2(pcm_to_float(ra.x), pcm_to_float(rb.x));
                im[b] = make_float2(pcm_to_float(ra.y), pcm_to_float(rb.y));
            }
        }
        for (uint32_t j = sync_add - 1; j-- > 0;) {
            const int64_t back = (int64_t) j * 2048;
            const V2* pa = reinterpret_cast<const V2*>(stream + gA - back) + lane;
            const V2* pb = reinterpret_cast<const V2*>(stream + gB - back) + lane;
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                const V2 ra = pa[32 * b], rb = pb[32 * b];
                re[b] = __fadd2_rn(re[b], make_float2(pcm_to_float(ra.x), pcm_to_float(rb.x)));
                im[b] = __fadd2_rn(im[b], make_float2(pcm_to_float(ra.y), pcm_to_float(rb.y)));
            }
        }
    } else {
        for (uint32_t j = sync_add; j-- > 0;) {        // oldest FIFO first
            const int64_t back = (int64_t) j * 2048;
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                const int m = lane + 32 * b;
                const float2 xa = load_pair<PCM>(stream, nsamples, gA - back + 2 * m);
                const float2 xb = load_pair<PCM>(stream, nsamples, gB - back + 2 * m);
                if (j == sync_add - 1) {
                    re[b] = make_float2(xa.x, xb.x);
                    im[b] = make_float2(xa.y, xb.y);
                } else {
                    re[b] = make_float2(__fadd_rn(re[b].x, xa.x), __fadd_rn(re[b].y, xb.x));
                    im[b] = make_float2(__fadd_rn(im[b].x, xa.y), __fadd_rn(im[b].y, xb.y));
                }
            }
        }
    }
    dsp_pair_tail(re, im, chirps, tb, tile, ws, lane, bw2, magA, idxA, magB, idxB);
}

// allocate the CTA's TMEM columns and fill every lane quadrant with the table rows (warp q fills quadrant q; warps q and
// q + 4 read it); ends with a CTA barrier
__device__ __forceinline__ rx_tables load_tables(const rx_params& p, uint32_t* s_tslot, int lane, int warp) {
    if (warp == 0) tmem_alloc<kRxTcols>(s_tslot);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t tq = tmem_quadrant(*s_tslot, warp);
    if (warp < 4) {
#pragma unroll 1
        for (int b0 = 0; b0 < 32; b0 += 4) {
            float2 u[4], d[4], w[4], z[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                u[j] = p.up[lane + 32 * (b0 + j)];
                d[j] = p.down[lane + 32 * (b0 + j)];
                w[j] = p.hann[lane + 32 * (b0 + j)];
                z[j] = p.tw_pass[(b0 + j) * 32 + lane];
            }
#pragma unroll
            for (int j = 0; j < 4; j += 2)
                sttm_f2x4(tq + kRxTud + 4 * (b0 + j), make_float2(u[j].x, d[j].x), make_float2(u[j].y, d[j].y),
                          make_float2(u[j + 1].x, d[j + 1].x), make_float2(u[j + 1].y, d[j + 1].y));
            sttm_f2x4(tq + kRxThann + 2 * b0, w[0], w[1], w[2], w[3]);
            sttm_f2x4(tq + kRxTtw + 2 * b0, z[0], z[1], z[2], z[3]);
        }
        sttm_wait();
    }
    const float one = p.tw_pass[lane].x;
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    return rx_tables{tq, one};
}
__device__ __forceinline__ void free_tables(uint32_t* s_tslot, int warp) {
    tmem_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<kRxTcols>(*s_tslot);
}

template <typename PCM, int W>
__global__ void __launch_bounds__(W * 32, 1) k_receiver_run(rx_params p) {
    using L = rx_smem<W>;
    constexpr int kRxWarps = W;
    extern __shared__ float2 s_rx[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2 ws[kRxNB];
#pragma unroll
    for (int d1 = 0; d1 < kRxNB; ++d1) ws[d1] = p.tw_split[lane + 32 * d1];
    const rx_tables tb = load_tables(p, reinterpret_cast<uint32_t*>(s_rx + L::slot), lane, warp);
    float2* tile = s_rx + L::tile + warp * L::tile_f2;
    const uint32_t N = 2048, offset = N / 8, shift = N / 4;
    const float thr = p.snr_threshold;
    const uint32_t bw2 = p.bandwidth2;

    const uint32_t nwarps = gridDim.x * kRxWarps;
    for (uint32_t s = blockIdx.x * kRxWarps + warp; s < p.nstreams; s += nwarps) {
        // sample 0 of `stream` is the first NEW frame; a resumed chunk has `carry` frames of history before it
        const PCM* stream = static_cast<const PCM*>(p.pcm) + (size_t) s * p.stream_stride + (size_t) p.carry * N;
        const int64_t nsamples = (int64_t) p.nframes * N, gmin = -(int64_t) p.carry * N;
        uint8_t* uart = p.uart ? p.uart + (size_t) s * p.uart_cap : nullptr;
        // state (main.c:311-339), identical in every lane
        uint32_t state = 0, turn = 0, sync_cnt = 0, pos = N / 2, max_idx = 0, msg = 0, msg_cnt = 0, nout = 0;
        int32_t lock_frame = -1;
        uint32_t lock_pos = 0, frames_before = 0;
        float mag_mean = 0.0f;
        const rx_state_rec* saved = p.rx_state && p.rx_state[s].magic == kRxStateMagic ? p.rx_state + s : nullptr;
        if (saved) {
            state = saved->state; turn = saved->turn; sync_cnt = saved->sync_cnt; pos = saved->pos; max_idx = saved->max_idx;
            msg = saved->msg; msg_cnt = saved->msg_cnt; lock_frame = saved->lock_frame; lock_pos = saved->lock_pos;
            frames_before = saved->frames_seen; mag_mean = saved->mag_mean;
        }
        // mag_stat[12] | history mag_max[8] | history mag_mean[4] live in shared memory (the packed core
        // needs the registers); every lane reads them (broadcast), lane 0 writes
        float* mag_stat = reinterpret_cast<float*>(s_rx + L::state) + warp * 32;
        float* hmag = mag_stat + 12;
        float* hmean = mag_stat + 20;
        __syncwarp();
        if (saved) { if (lane < 24) mag_stat[lane] = saved->stat[lane]; }
        else if (lane < 12) mag_stat[lane] = 1E37f;
        else if (lane < 24) mag_stat[lane] = 0.0f;
        __syncwarp();

        auto emit = [&](uint32_t c) {
            if (lane == 0 && uart && nout < p.uart_cap) uart[nout] = (uint8_t) c;
            nout++;
        };
        for (uint32_t t = 0; t < p.nframes; ++t) {
            const int64_t fifo0 = ((int64_t) t - 2) * N;                  // stream index of fifo_queue[0]
            if (t + 1 < p.nframes) {                                      // pull the next frame towards L2
                const char* nxt = reinterpret_cast<const char*>(stream + (size_t) (t + 1) * N);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + lane * 128));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + 4096 + lane * 128));
            }
            const uint32_t prev = state;
            const bool searching = state == 0 || state == 1;
            if (state == 0) {                                             // IDLE (main.c:428-434)
                sync_cnt = 0;
                float sum = 0.0f;
                for (int i = 4; i < 12; ++i) sum = __fadd_rn(sum, mag_stat[i]);
                mag_mean = __fdiv_rn(sum, 8.0f);
            }
            // Every frame costs two packed passes through ONE inlined copy of the chain:
            //   searching: offsets (0,1) then (2,3) of the grid, up-chirp            main.c:447-451
            //   locked:    (UP, DOWN) at sync_position, then the two resync probes    main.c:493-494, 243-249
            float up = 0.0f, down = 0.0f;
            bool is_down = false, symbol = false;
            for (int pass = 0; pass < 2; ++pass) {
                int64_t qa, qb;
                int chirps;
                bool a_ok = true, b_ok = true;
                if (searching) {
                    qa = N / 2 + turn * offset + shift * (2 * pass);
                    qb = qa + shift;
                    chirps = RX_UP_UP;
                } else if (pass == 0) {
                    qa = qb = pos;
                    chirps = RX_UP_DOWN;
                } else {
                    if (!symbol) break;                                   // neither SNR reached the threshold
                    qa = (int64_t) pos - offset;                          // resync (main.c:246-249); probes outside
                    qb = (int64_t) pos + offset;                          // [0, 2N] are hazards H3/H5: snr = -inf
                    a_ok = qa >= 0 && qa <= (int64_t) 2 * N;
                    b_ok = qb >= 0 && qb <= (int64_t) 2 * N;
                    chirps = is_down ? RX_DOWN_DOWN : RX_UP_UP;
                }
                float ma, mb;
                uint32_t ka, kb;
                dsp_pair<PCM, false>(stream, nsamples, fifo0 + (a_ok ? qa : (int64_t) pos), fifo0 + (b_ok ? qb : (int64_t) pos), chirps, tb,
                              1, tile, ws, lane, bw2, ma, ka, mb, kb, gmin);
                if (searching) {
                    const int sa = (int) (4 * pass + turn), sb = sa + 2;  // history[i*2 + turn]
                    __syncwarp();
                    if (lane == 0) {
                        hmag[sa] = ma;
                        hmag[sb] = mb;
                        if (sa < 4) hmean[sa] = mag_mean;
                        if (sb < 4) hmean[sb] = mag_mean;
                    }
                    __syncwarp();
                    pos = (uint32_t) qb;
                } else if (pass == 0) {
                    const float mean0 = hmean[0], mean1 = hmean[1];
                    __syncwarp();
                    if (lane == 0) { hmag[0] = ma; hmag[1] = mb; }
                    __syncwarp();
                    up = __fdiv_rn(__fsub_rn(ma, mean0), mean0);          // main.c:229
                    down = __fdiv_rn(__fsub_rn(mb, mean1), mean1);
                    symbol = up >= thr || down >= thr;
                    is_down = down > up;
                } else {
                    float snr_l = -INFINITY, snr_r = -INFINITY;
                    const float mean2 = hmean[2], mean3 = hmean[3];
                    __syncwarp();
                    if (lane == 0) {
                        if (a_ok) hmag[2] = ma;
                        if (b_ok) hmag[3] = mb;
                    }
                    __syncwarp();
                    if (a_ok) snr_l = __fdiv_rn(__fsub_rn(ma, mean2), mean2);
                    if (b_ok) snr_r = __fdiv_rn(__fsub_rn(mb, mean3), mean3);
                    const float snr = is_down ? down : up;
                    if ((snr > snr_l) && (snr > snr_r)) {                 // main.c:252-270
                    } else if (snr_l >= snr_r) {
                        if (qa >= 0) pos = (uint32_t) qa;
                    } else if (snr_l < snr_r) {
                        if (qb <= (int64_t) 2 * N) pos = (uint32_t) qb;
                    }
                }
            }
            if (searching) {
                turn ^= 1u;
                if (turn == 1u) {
                    float mmm = 0.0f;
                    for (int i = 0; i < 8; ++i)
                        if (hmag[i] > mmm) { mmm = hmag[i]; max_idx = i; }           // main.c:463-471
                    const float shifted = lane >= 1 && lane < 12 ? mag_stat[lane - 1] : mmm;
                    __syncwarp();
                    if (lane < 12) mag_stat[lane] = shifted;                         // main.c:458-460, 473
                    __syncwarp();
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
            } else if (symbol) {
                if (state == 3) msg = ((msg << 1) + (is_down ? 0u : 1u)) & 0xffu;     // main.c:525,529
                if (state == 2) {
                    if (is_down) state = 3;                               // the delimiter (main.c:500)
                } else if (++msg_cnt >= 8) {                              // main.c:532-537
                    emit(msg);
                    msg = 0;
                    msg_cnt = 0;
                }
            } else {
                if (state == 3) {                                         // end of message (main.c:539-549)
                    emit((uint32_t) '\n');
                    msg = 0;
                    msg_cnt = 0;
                }
                state = 0;
            }
            if (prev != 2 && state == 2 && lock_frame < 0) {
                lock_frame = (int32_t) (frames_before + t);
                lock_pos = pos;
            }
        }
        if (lane == 0 && p.results) {
            rx_result_rec r;
            r.state = state; r.sync_position = pos; r.lock_frame = lock_frame; r.lock_position = lock_pos;
            r.nbytes = nout; r.frames_seen = frames_before + p.nframes; r.turn = turn; r.sync_cnt = sync_cnt;
            p.results[s] = r;
        }
        if (p.rx_state) {
            __syncwarp();
            rx_state_rec* o = p.rx_state + s;
            if (lane < 24) o->stat[lane] = mag_stat[lane];
            if (lane == 0) {
                o->magic = kRxStateMagic; o->state = state; o->turn = turn; o->sync_cnt = sync_cnt; o->pos = pos; o->max_idx = max_idx;
                o->msg = msg; o->msg_cnt = msg_cnt; o->lock_frame = lock_frame; o->lock_pos = lock_pos;
                o->frames_seen = frames_before + p.nframes; o->mag_mean = mag_mean;
            }
            __syncwarp();
        }
    }
    free_tables(reinterpret_cast<uint32_t*>(s_rx + L::slot), warp);
}

// K4: the search grid of main.c:447-451 for every (stream, frame), after optional synchronous
// addition of sync_add frame-aligned FIFOs (oldest first).  One warp per (stream, frame), two packed
// passes of two offsets each.
template <typename PCM, bool MULTI, int W>
__global__ void __launch_bounds__(W * 32, 1) k_sync_search(rx_params p) {
    using L = rx_smem<W>;
    constexpr int kRxWarps = W;
    constexpr bool kUnion = W == 8;                                                // the 14 KB per-warp union area exists (8-warp forms)
    extern __shared__ float2 s_rx[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2 ws[kRxNB];
#pragma unroll
    for (int d1 = 0; d1 < kRxNB; ++d1) ws[d1] = p.tw_split[lane + 32 * d1];
    const rx_tables tb = load_tables(p, reinterpret_cast<uint32_t*>(s_rx + L::slot), lane, warp);
    float2* tile = s_rx + L::tile + warp * L::tile_f2;
    float2* sum = s_rx + L::sum + warp * 1792;                                     // 14 KB per warp: staged union (PCM) or its K-frame sum
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_rx + L::bar) + warp;
    const uint32_t N = 2048, offset = N / 8, shift = N / 4;
    const size_t total = (size_t) p.nstreams * p.nframes;
    const size_t nwarps = (size_t) gridDim.x * kRxWarps;
    const int64_t nsamples = (int64_t) p.nframes * N;
    using V2 = typename vec2<PCM>::type;
    const bool staged = kUnion && !MULTI && p.staged != 0;
    // the union of work item wi's four windows: [base, base + 3584) of its stream; `inside`: wholly within the stream
    auto item_union = [&](size_t wi, const PCM*& strm, int64_t& base) -> bool {
        const uint32_t si = (uint32_t) (wi / p.nframes), ti = (uint32_t) (wi - (size_t) si * p.nframes);
        strm = static_cast<const PCM*>(p.pcm) + (size_t) si * p.stream_stride;
        base = ((int64_t) ti - 2) * N + N / 2 + (ti & 1u) * offset;
        return base >= 0 && base + 3584 <= nsamples;
    };
    auto fetch_union = [&](size_t wi) {                                            // lane 0: one 14 KB bulk copy, if the union is inside
        const PCM* strm;
        int64_t base;
        if (wi < total && item_union(wi, strm, base)) {
            mbar_expect_tx(bar, 14336u);
            bulk_g2s(sum, strm + base, 14336u, bar);
        }
    };
    uint32_t parity = 0;
    if (staged) {
        if (lane == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            fetch_union((size_t) blockIdx.x * kRxWarps + warp);
        }
        __syncwarp();
    }
    for (size_t w = (size_t) blockIdx.x * kRxWarps + warp; w < total; w += nwarps) {
        const uint32_t s = (uint32_t) (w / p.nframes), t = (uint32_t) (w - (size_t) s * p.nframes);
        const PCM* stream = static_cast<const PCM*>(p.pcm) + (size_t) s * p.stream_stride;
        const int64_t fifo0 = ((int64_t) t - 2) * N;
        if (staged) {
            // ---- staged form: the union arrived by TMA while the previous item computed ----
            const PCM* strm;
            int64_t base;
            const bool inside = item_union(w, strm, base);
            if (inside) {
                mbar_wait(bar, parity);
                parity ^= 1u;
            } else if (lane == 0) {
                fetch_union(w + nwarps);                                           // this item gathers from global memory: the stage is free
            }
            for (uint32_t i = 0; i < 4; i += 2) {
                const uint32_t pa = N / 2 + (t & 1u) * offset + shift * i, pb = pa + shift;
                float ma, mb;
                uint32_t ka, kb;
                if (inside) {
                    float2 re[32], im[32];
                    const V2* sa = reinterpret_cast<const V2*>(sum) + (shift / 2) * i + lane;
                    const V2* sb = sa + shift / 2;
#pragma unroll
                    for (int b = 0; b < 32; ++b) {
                        const V2 ra = sa[32 * b], rb = sb[32 * b];
                        re[b] = make_float2(pcm_to_float(ra.x), pcm_to_float(rb.x));
                        im[b] = make_float2(pcm_to_float(ra.y), pcm_to_float(rb.y));
                    }
                    if (i == 2) {                                                  // the stage has been consumed: next item's union
                        __syncwarp();
                        if (lane == 0) fetch_union(w + nwarps);
                    }
                    dsp_pair_tail(re, im, RX_UP_UP, tb, tile, ws, lane, p.bandwidth2, ma, ka, mb, kb);
                } else {
                    dsp_pair<PCM, false>(stream, nsamples, fifo0 + pa, fifo0 + pb, RX_UP_UP, tb, 1, tile, ws, lane, p.bandwidth2,
                                         ma, ka, mb, kb);
                }
                if (lane == 0) {
                    p.ss_mag[w * 4 + i] = ma; p.ss_idx[w * 4 + i] = ka;
                    p.ss_mag[w * 4 + i + 1] = mb; p.ss_idx[w * 4 + i + 1] = kb;
                }
            }
            continue;
        }
        {   // pull the next work item's window union (1.75 N samples = 14 KB) towards L2 while this one computes
            const size_t wn = w + nwarps;
            if (wn < total) {
                const uint32_t sn = (uint32_t) (wn / p.nframes), tn = (uint32_t) (wn - (size_t) sn * p.nframes);
                const int64_t gn = ((int64_t) tn - 2) * N + N / 2 + (tn & 1u) * offset;
                if (gn >= 0) {
                    const char* nxt = reinterpret_cast<const char*>(static_cast<const PCM*>(p.pcm) + (size_t) sn * p.stream_stride + gn);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (q * 4096 + lane * 128 < 14336) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + q * 4096 + lane * 128));
                }
            }
        }
        if (MULTI) {
            // Synchronous addition: the four search windows overlap (hop N/4), so the K-frame sum is formed ONCE over
            // their union (1.75 N samples, oldest FIFO first as the per-window form does) and parked in shared memory.
            const int64_t base = fifo0 + N / 2 + (t & 1u) * offset;
            const int64_t oldest = base - (int64_t) (p.sync_add - 1) * N;
            const bool inside = oldest >= 0 && base + 3584 <= nsamples;
            __syncwarp();
            if (inside) {                                                  // two batches of 28 pairs per lane: 28 loads in flight
#pragma unroll 1
                for (uint32_t u0 = 0; u0 < 1792u; u0 += 896u) {
                    float2 acc[28];
                    const V2* src0 = reinterpret_cast<const V2*>(stream + oldest) + u0 + lane;
#pragma unroll
                    for (int q = 0; q < 28; ++q) {
                        const V2 r = src0[32 * q];
                        acc[q] = make_float2(pcm_to_float(r.x), pcm_to_float(r.y));
                    }
                    for (uint32_t j = p.sync_add - 1; j-- > 0;) {
                        const V2* srcj = reinterpret_cast<const V2*>(stream + base - (int64_t) j * N) + u0 + lane;
                        V2 r[28];
#pragma unroll
                        for (int q = 0; q < 28; ++q) r[q] = srcj[32 * q];
#pragma unroll
                        for (int q = 0; q < 28; ++q) acc[q] = __fadd2_rn(acc[q], make_float2(pcm_to_float(r[q].x), pcm_to_float(r[q].y)));
                    }
#pragma unroll
                    for (int q = 0; q < 28; ++q) sum[u0 + lane + 32 * q] = acc[q];
                }
            } else {
                for (uint32_t u = lane; u < 1792u; u += 32u) {             // stream edges: checked loads
                    float2 acc = load_pair<PCM>(stream, nsamples, oldest + 2 * u);
                    for (uint32_t j = p.sync_add - 1; j-- > 0;) {
                        const float2 r = load_pair<PCM>(stream, nsamples, base - (int64_t) j * N + 2 * u);
                        acc = make_float2(__fadd_rn(acc.x, r.x), __fadd_rn(acc.y, r.y));
                    }
                    sum[u] = acc;
                }
            }
            __syncwarp();
        }
        for (uint32_t i = 0; i < 4; i += 2) {
            const uint32_t pa = N / 2 + (t & 1u) * offset + shift * i, pb = pa + shift;
            float ma, mb;
            uint32_t ka, kb;
            if (MULTI) {
                float2 re[32], im[32];
                const float2* sa = sum + (shift / 2) * i;
                const float2* sb = sa + shift / 2;
#pragma unroll
                for (int b = 0; b < 32; ++b) {
                    const float2 xa = sa[lane + 32 * b], xb = sb[lane + 32 * b];
                    re[b] = make_float2(xa.x, xb.x);
                    im[b] = make_float2(xa.y, xb.y);
                }
                dsp_pair_tail(re, im, RX_UP_UP, tb, tile, ws, lane, p.bandwidth2, ma, ka, mb, kb);
            } else
            dsp_pair<PCM, false>(stream, nsamples, fifo0 + pa, fifo0 + pb, RX_UP_UP, tb, 1, tile, ws, lane, p.bandwidth2,
                          ma, ka, mb, kb);
            if (lane == 0) {
                p.ss_mag[w * 4 + i] = ma; p.ss_idx[w * 4 + i] = ka;
                p.ss_mag[w * 4 + i + 1] = mb; p.ss_idx[w * 4 + i + 1] = kb;
            }
        }
    }
    free_tables(reinterpret_cast<uint32_t*>(s_rx + L::slot), warp);
}

static rx_params make_params(const rx_launch& a) {
    rx_params p{};
    p.pcm = a.pcm; p.nstreams = a.nstreams; p.nframes = a.nframes; p.stream_stride = a.stream_stride;
    p.up = a.up; p.down = a.down; p.hann = a.hann; p.tw_pass = a.tw_pass; p.tw_split = a.tw_split;
    p.bandwidth2 = a.bandwidth2; p.snr_threshold = a.snr_threshold;
    p.uart = a.uart; p.uart_cap = a.uart_cap; p.results = a.results;
    p.sync_add = a.sync_add < 1 ? 1 : a.sync_add; p.ss_mag = a.ss_mag; p.ss_idx = a.ss_idx;
    p.carry = a.carry; p.rx_state = a.rx_state;
    return p;
}

constexpr int kRxW = USC_RX_WARPS, kSsW = USC_SS_WARPS;
static cudaError_t rx_prepare() {
    static per_device<bool> done_pd;
    bool& done = done_pd.get();
    if (done) return cudaSuccess;
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_receiver_run<int32_t, kRxW>, cudaFuncAttributeMaxDynamicSharedMemorySize, rx_smem<kRxW>::bytes))) return e;
    if ((e = cudaFuncSetAttribute(k_receiver_run<float, kRxW>, cudaFuncAttributeMaxDynamicSharedMemorySize, rx_smem<kRxW>::bytes))) return e;
    constexpr int ss_bytes = kSsW == 8 ? rx_smem<8>::bytes_union : rx_smem<kSsW>::bytes;
    if ((e = cudaFuncSetAttribute(k_sync_search<int32_t, false, kSsW>, cudaFuncAttributeMaxDynamicSharedMemorySize, ss_bytes))) return e;
    if ((e = cudaFuncSetAttribute(k_sync_search<float, false, kSsW>, cudaFuncAttributeMaxDynamicSharedMemorySize, ss_bytes))) return e;
    if ((e = cudaFuncSetAttribute(k_sync_search<int32_t, true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, rx_smem<8>::bytes_union))) return e;
    if ((e = cudaFuncSetAttribute(k_sync_search<float, true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, rx_smem<8>::bytes_union))) return e;
    done = true;
    return cudaSuccess;
}

cudaError_t launch_receiver_run(const rx_launch& a, int num_sms, cudaStream_t st) {
    rx_params p = make_params(a);
    size_t ctas = ((size_t) a.nstreams + kRxW - 1) / kRxW;
    const size_t cap = (size_t) num_sms;                 // persistent: one CTA per SM
    if (ctas > cap) ctas = cap;
    cudaError_t e = rx_prepare();
    if (e != cudaSuccess) return e;
    if (a.pcm_format == 1u) k_receiver_run<int32_t, kRxW><<<(int) ctas, kRxW * 32, rx_smem<kRxW>::bytes, st>>>(p);
    else k_receiver_run<float, kRxW><<<(int) ctas, kRxW * 32, rx_smem<kRxW>::bytes, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_sync_search(const rx_launch& a, int num_sms, cudaStream_t st) {
    rx_params p = make_params(a);
    const bool multi = p.sync_add > 1;
    const int w = multi ? 8 : kSsW;
    size_t ctas = ((size_t) a.nstreams * a.nframes + w - 1) / w;
    const size_t cap = (size_t) num_sms;                 // persistent: one CTA per SM
    if (ctas > cap) ctas = cap;
    cudaError_t e = rx_prepare();
    if (e != cudaSuccess) return e;
    p.staged = (reinterpret_cast<uintptr_t>(a.pcm) % 16u == 0 && a.stream_stride % 4u == 0) ? 1u : 0u;
    constexpr int ss_bytes = kSsW == 8 ? rx_smem<8>::bytes_union : rx_smem<kSsW>::bytes;
    if (a.pcm_format == 1u) {
        if (multi) k_sync_search<int32_t, true, 8><<<(int) ctas, 8 * 32, rx_smem<8>::bytes_union, st>>>(p);
        else k_sync_search<int32_t, false, kSsW><<<(int) ctas, kSsW * 32, ss_bytes, st>>>(p);
    } else {
        if (multi) k_sync_search<float, true, 8><<<(int) ctas, 8 * 32, rx_smem<8>::bytes_union, st>>>(p);
        else k_sync_search<float, false, kSsW><<<(int) ctas, kSsW * 32, ss_bytes, st>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace usc
