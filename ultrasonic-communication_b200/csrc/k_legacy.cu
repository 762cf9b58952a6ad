// k_legacy.cu — the two earlier detectors of the reference behind the same packed core.
// Both start like the spectrum analyser: (float) pcm x Hann -> 2048-point RFFT -> magnitude x 1/sqrt(N)
// (experiments/chirp/Src/main.c:200-216, experiments/ultracom/Src/main.c:115-128), then
//   on/off chirp  count of bins in [f1_idx, f2_idx] above a magnitude threshold -> HIGH / LOW / UNKNOWN
//                 (chirp/Src/main.c:237-286) -> decode() bit framing (:119-198)
//   FSK 18 tones  first tone bin (start-of-frame, end-of-frame, hex 0..F, in that order) above a threshold
//                 (ultracom/Src/main.c:130-168) -> parser() debounce + nibble pairing (:175-236)
// k_band2048_pair is the single-hypothesis pair kernel of k_demod.cu without the de-chirp: two frames
// ride in the f32x2 halves, PCM arrives by TMA bulk copies, bins k < 512 are split and magnituded
// (both detectors live below bin 512 at every sampling rate the reference used).  The per-stream
// framing state machines are sequential integer code: one thread per stream.
#include "usc_kernels.cuh"
#include "usc_launch.h"
#include "usc_warpfft.cuh"

namespace usc {

constexpr int kBandWarps = 8, kBandNB = 16;                         // bins [0, 512)
constexpr int kBandTabs = 2 * 8192;                                 // twiddles | Hann
constexpr int kBandWarpBytes = kTileFloat2 * 8 + 16384;             // padded tile (later: 2 x 512 magnitudes) + 2-frame PCM stage
constexpr int kBandBar = kBandTabs + kBandWarps * kBandWarpBytes;
constexpr int kBandSmem = kBandBar + kBandWarps * 8;

template <typename PCM>
__global__ void __launch_bounds__(kBandWarps * 32, 1) k_band2048_pair(band_params p) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    float2* s_tw = reinterpret_cast<float2*>(s_raw);
    float2* s_hann = reinterpret_cast<float2*>(s_raw + 8192);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wbase = s_raw + kBandTabs + warp * kBandWarpBytes;
    using V2 = typename vec2<PCM>::type;
    V2* xstage = reinterpret_cast<V2*>(wbase);
    float2* tile = reinterpret_cast<float2*>(wbase + 16384);
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_raw + kBandBar) + warp;

    const size_t npairs = (p.nframes + 1) / 2;
    const size_t nwarps = (size_t) gridDim.x * kBandWarps;
    size_t q = (size_t) blockIdx.x * kBandWarps + warp;
    const PCM* pcm = static_cast<const PCM*>(p.pcm);
    auto pair_bytes = [&](size_t pr) -> uint32_t { return 2 * pr + 1 < p.nframes ? 16384u : 8192u; };
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (q < npairs) {
            mbar_expect_tx(bar, pair_bytes(q));
            bulk_g2s(xstage, pcm + q * 4096, pair_bytes(q), bar);
        }
    }
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        s_tw[i] = p.tw_pass[i];
        s_hann[i] = p.hann[i];
    }
    float2 ws[kBandNB];
#pragma unroll
    for (int d1 = 0; d1 < kBandNB; ++d1) ws[d1] = p.tw_split[lane + 32 * d1];
    __syncthreads();

    uint32_t parity = 0;
    for (; q < npairs; q += nwarps) {
        const bool two = 2 * q + 1 < p.nframes;
        mbar_wait(bar, parity);
        parity ^= 1u;
        float2 re[32], im[32];                                        // (.x, .y) = (frame 2q, frame 2q+1)
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const int m = lane + 32 * b;
            const V2 ra = xstage[m];
            const V2 rb = two ? xstage[1024 + m] : ra;
            const float2 w = s_hann[m];
            // arm_mult_f32(fft_input, fft_window), packed; the first butterfly stage takes the products as FMAs by 1.0
            re[b] = __fmul2_rn(make_float2(pcm_to_float(ra.x), pcm_to_float(rb.x)), bc2(w.x));
            im[b] = __fmul2_rn(make_float2(pcm_to_float(ra.y), pcm_to_float(rb.y)), bc2(w.y));
        }
        __syncwarp();
        if (lane == 0 && q + nwarps < npairs) {
            mbar_expect_tx(bar, pair_bytes(q + nwarps));
            bulk_g2s(xstage, pcm + (q + nwarps) * 4096, pair_bytes(q + nwarps), bar);
        }
        fft_base2_prod<32>(re, im, s_tw[lane].x);
#pragma unroll
        for (int d = 1; d < 32; ++d) {
            const float2 w = s_tw[d * 32 + lane];
            float2 tr, ti;
            cmul2(re[d], im[d], w.x, w.y, tr, ti);
            re[d] = tr;
            im[d] = ti;
        }
#pragma unroll
        for (int d = 0; d < 32; ++d) tile[d * kTileStride + lane] = re[d];
        __syncwarp();
#pragma unroll
        for (int a = 0; a < 32; ++a) re[a] = tile[lane * kTileStride + a];
        __syncwarp();
#pragma unroll
        for (int d = 0; d < 32; ++d) tile[d * kTileStride + lane] = im[d];
        __syncwarp();
#pragma unroll
        for (int a = 0; a < 32; ++a) im[a] = tile[lane * kTileStride + a];
        __syncwarp();
        fft_base2<32>(re, im);
        float pa[kBandNB], pb[kBandNB];
        mag2_window_pair<kBandNB>(re, im, ws, lane, pa, pb);
        // magnitudes: arm_cmplx_mag_f32 then arm_scale_f32 by 1/sqrt(N); the tile now holds 2 x 512 of them
        float* row = reinterpret_cast<float*>(tile);
        uint32_t cnt_a = 0, cnt_b = 0;
#pragma unroll
        for (int d1 = 0; d1 < kBandNB; ++d1) {
            const uint32_t k = lane + 32u * d1;
            const float ma = __fmul_rn(__fsqrt_rn(pa[d1]), p.inv_sqrt_n), mb = __fmul_rn(__fsqrt_rn(pb[d1]), p.inv_sqrt_n);
            row[k] = ma;
            row[512 + k] = mb;
            const bool in_band = k >= p.band_lo && k <= p.band_hi;
            cnt_a += (in_band && ma > p.onoff_threshold) ? 1u : 0u;        // chirp/Src/main.c:238-242
            cnt_b += (in_band && mb > p.onoff_threshold) ? 1u : 0u;
        }
        __syncwarp();
        cnt_a = __reduce_add_sync(0xffffffffu, cnt_a);
        cnt_b = __reduce_add_sync(0xffffffffu, cnt_b);
        if (p.mag) {                                                   // optional: the 512 magnitudes per frame
            for (int k = lane; k < 512; k += 32) {
                p.mag[(2 * q) * 512 + k] = row[k];
                if (two) p.mag[(2 * q + 1) * 512 + k] = row[512 + k];
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h == 1 && !two) break;
                const uint32_t c = h ? cnt_b : cnt_a;
                if (p.strength) p.strength[2 * q + h] = (uint16_t) c;
                // chirp/Src/main.c:276-286: HIGH 1, LOW -1, UNKNOWN 0
                if (p.level) p.level[2 * q + h] = c >= p.thr_high ? (int8_t) 1 : (c <= p.thr_low ? (int8_t) -1 : (int8_t) 0);
            }
        }
        if (p.code) {
            // ultracom/Src/main.c:130-168: lanes 0..17 = start-of-frame, end-of-frame, hex 0..F; each scans its
            // tolerance window in ascending bin order; the first lane that found something wins
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h == 1 && !two) break;
                const float* r = row + 512 * h;
                uint32_t found = 0xffffffffu;
                if (lane < 18) {
                    const uint32_t centre = lane == 0 ? p.sof_bin : (lane == 1 ? p.eof_bin : p.hex0_bin + (uint32_t) (lane - 2) * p.hex_step);
                    for (uint32_t j = centre - p.tolerance; j <= centre + p.tolerance; ++j)
                        if (r[j] > p.fsk_threshold) { found = j; break; }
                }
                const uint32_t hit = __ballot_sync(0xffffffffu, found != 0xffffffffu);
                const int win = hit ? __ffs(hit) - 1 : 0;
                const uint32_t j = __shfl_sync(0xffffffffu, found, win);
                if (lane == 0) {
                    const size_t f = 2 * q + h;
                    if (!hit) {
                        p.code[f] = 0xFF;                              // NOT_FOUND_CODE; magnitude/frequency keep their last value in the
                        if (p.code_mag) p.code_mag[f] = 0.0f;          // firmware — defined as 0 here
                        if (p.code_freq) p.code_freq[f] = 0.0f;
                    } else {
                        p.code[f] = win == 0 ? 0xF0 : (win == 1 ? 0xF1 : (uint8_t) (win - 2));
                        if (p.code_mag) p.code_mag[f] = r[j];
                        // result->frequency = frequency[j + 1] with frequency[i] = (float) i * fs / (float) N (main.c:137, 314)
                        if (p.code_freq) p.code_freq[f] = __fdiv_rn(__fmul_rn((float) (j + 1), p.fs), 2048.0f);
                    }
                }
            }
        }
        __syncwarp();
    }
}

// decode() of experiments/chirp/Src/main.c:119-198, one thread per stream over its frames in order.
__global__ void k_onoff_decode(const int8_t* __restrict__ level, uint32_t nstreams, uint32_t nframes, uint32_t frame_start,
                               uint32_t frame_bit, uint32_t sync_threshold, uint32_t sampling_offset, uint8_t* chars, uint32_t cap,
                               uint32_t* nchars, uint32_t* sync_errors) {
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < nstreams; s += gridDim.x * blockDim.x) {
        uint32_t count = 0, n = 0, high_count = 0, bits = 0, out = 0, errs = 0;
        const uint32_t offset = frame_start + sampling_offset, max_length = offset + frame_bit * 8u;
        for (uint32_t t = 0; t < nframes; ++t) {
            const int lv = level[(size_t) s * nframes + t];
            const bool sampling_point = count == offset + frame_bit * n;
            if (lv > 0) {
                if (count < offset) high_count++;
                else if (sampling_point) { bits |= n < 8u ? (0x80u >> n) : 0u; n++; }
                count++;
            } else if (lv == 0) {
                if (sampling_point) n++;
                count++;
            } else if (count > 0) {
                count++;
                if (sampling_point) n++;
            }
            if (count >= frame_start && high_count < sync_threshold) {     // "Sync error!" (n and bits are NOT reset, as in the firmware)
                errs++;
                count = 0;
                high_count = 0;
            }
            if (count >= max_length) {
                count = 0; n = 0; high_count = 0;
                if (chars && out < cap) chars[(size_t) s * cap + out] = (uint8_t) bits;
                out++;
                bits = 0;
            }
        }
        if (nchars) nchars[s] = out;
        if (sync_errors) sync_errors[s] = errs;
    }
}

// parser() of experiments/ultracom/Src/main.c:175-236, one thread per stream.
__global__ void k_fsk_parse(const uint8_t* __restrict__ code, uint32_t nstreams, uint32_t nframes, uint32_t tq_n, uint8_t* chars,
                            uint32_t cap, uint32_t* nchars, uint32_t* nsof, uint32_t* neof) {
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < nstreams; s += gridDim.x * blockDim.x) {
        uint32_t state = 0 /* IDLE, 1 DATA_MSB, 2 DATA_LSB */, hex = 0xFF, cnt = 0, msb = 0, out = 0, sof = 0, eof = 0;
        for (uint32_t t = 0; t < nframes; ++t) {
            const uint32_t data = code[(size_t) s * nframes + t];
            bool output = false;
            if (data != hex) { cnt = 0; hex = data; }
            else if (cnt == tq_n) { }
            else if (++cnt == tq_n && hex != 0xFFu) output = true;
            if (!output) continue;
            if (hex == 0xF0u) { state = 1; sof++; }
            else if (hex == 0xF1u) { state = 0; eof++; }
            else if (state == 1) { msb = (hex << 4) & 0xffu; state = 2; }
            else if (state == 2) {
                if (chars && out < cap) chars[(size_t) s * cap + out] = (uint8_t) (msb + hex);
                out++;
                msb = 0;
                state = 1;
            }
        }
        if (nchars) nchars[s] = out;
        if (nsof) nsof[s] = sof;
        if (neof) neof[s] = eof;
    }
}

cudaError_t launch_band2048(const band_params& p, uint32_t pcm_format, int num_sms, cudaStream_t st) {
    size_t ctas = ((p.nframes + 1) / 2 + kBandWarps - 1) / kBandWarps;
    if (ctas > (size_t) num_sms) ctas = (size_t) num_sms;
    static per_device<bool> configured_pd;
    bool& configured = configured_pd.get();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_band2048_pair<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBandSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_band2048_pair<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBandSmem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    if (pcm_format == 1u) k_band2048_pair<int32_t><<<(int) ctas, kBandWarps * 32, kBandSmem, st>>>(p);
    else k_band2048_pair<float><<<(int) ctas, kBandWarps * 32, kBandSmem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_onoff_decode(const int8_t* level, uint32_t nstreams, uint32_t nframes, uint32_t frame_start, uint32_t frame_bit,
                                uint32_t sync_threshold, uint32_t sampling_offset, uint8_t* chars, uint32_t cap, uint32_t* nchars,
                                uint32_t* sync_errors, cudaStream_t st) {
    k_onoff_decode<<<(int) ((nstreams + 127) / 128), 128, 0, st>>>(level, nstreams, nframes, frame_start, frame_bit, sync_threshold,
                                                                    sampling_offset, chars, cap, nchars, sync_errors);
    return cudaGetLastError();
}

cudaError_t launch_fsk_parse(const uint8_t* code, uint32_t nstreams, uint32_t nframes, uint32_t tq_n, uint8_t* chars, uint32_t cap,
                             uint32_t* nchars, uint32_t* nsof, uint32_t* neof, cudaStream_t st) {
    k_fsk_parse<<<(int) ((nstreams + 127) / 128), 128, 0, st>>>(code, nstreams, nframes, tq_n, chars, cap, nchars, nsof, neof);
    return cudaGetLastError();
}

}  // namespace usc
