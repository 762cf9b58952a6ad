// k_synth.cu — counter-based synthetic receiver input on the device (SURVEY §8d "input generation",
// §8f row f2): frame f carries the transmitter's up- or down-chirp symbol (chirp_orth law of
// simulation/signal.py:45-53, as integer tables) plus noise, as int32 DFSDM-style words (x256).
// Everything is integer arithmetic on Philox-4x32-10 output keyed by (seed, frame, sample block), so
// any frame can be regenerated bit for bit on the CPU (the oracle has a twin) without storing the
// dataset: full-size runs are checked by re-creating sampled frames on the host.
//   bit[f]      = philox(seed; f, 0, 0xB175, 0).x & 1
//   noise[f][n] = ((u0 + u1 + u2 + u3 - 131070) * gain) / 65536      (four 16-bit uniforms: Irwin-Hall)
//   pcm[f][n]   = (table[bit][n] + noise) * 256
#include "usc_kernels.cuh"
#include "usc_launch.h"

namespace usc {

__host__ __device__ inline void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t) 0xD2511F53u * c0, p1 = (uint64_t) 0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t) (p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t) p1;
        const uint32_t n2 = (uint32_t) (p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t) p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void k_synth_frames(uint64_t seed, uint64_t first_frame, size_t nframes, uint32_t n, const int32_t* __restrict__ table,
                               int32_t gain, int32_t* __restrict__ pcm, uint8_t* __restrict__ bits) {
    const uint32_t k0 = (uint32_t) seed, k1 = (uint32_t) (seed >> 32);
    const size_t pairs = (size_t) n / 2;
    const size_t total = nframes * pairs;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t) gridDim.x * blockDim.x) {
        const size_t fl = i / pairs;
        const uint32_t blk = (uint32_t) (i - fl * pairs);
        const uint64_t f = first_frame + fl;
        uint32_t r[4];
        philox4x32_10(k0, k1, (uint32_t) f, (uint32_t) (f >> 32), 0xB175u, 0u, r);
        const uint32_t bit = r[0] & 1u;
        if (blk == 0 && bits) bits[fl] = (uint8_t) bit;
        philox4x32_10(k0, k1, (uint32_t) f, (uint32_t) (f >> 32), blk, 1u, r);
        const int32_t* tab = table + (size_t) (bit ? 0u : 1u) * n + 2 * blk;      // bit 1 = up symbol (table 0)
        int2 o;
        {
            const int32_t s = (int32_t) ((r[0] & 0xffffu) + (r[0] >> 16) + (r[1] & 0xffffu) + (r[1] >> 16)) - 131070;
            o.x = (tab[0] + (int32_t) (((int64_t) s * gain) / 65536)) * 256;
        }
        {
            const int32_t s = (int32_t) ((r[2] & 0xffffu) + (r[2] >> 16) + (r[3] & 0xffffu) + (r[3] >> 16)) - 131070;
            o.y = (tab[1] + (int32_t) (((int64_t) s * gain) / 65536)) * 256;
        }
        reinterpret_cast<int2*>(pcm)[i] = o;
    }
}

// Whole streams in the transmitter's frame format (usc_synth_streams).  Per stream g:
//   offset[g]  = philox(seed; g, 0x0FF5E7, 0).x mod n
//   msg[g][m]  = 0x20 + (byte (m & 3) of word ((m >> 2) & 3) of philox(seed; g, 0x4D5347, m >> 4)) mod 95
//   symbol k of the pattern: G for k < lead_in, H for the next 7, L, then the message bits MSB first, then G
//   pcm[g][i]  = (table[kind][(i - offset) mod n] + noise(seed; g, block i/2, 2)) * 256, silence before the offset
__host__ __device__ inline uint32_t stream_offset(uint32_t k0, uint32_t k1, uint64_t g, uint32_t n) {
    uint32_t r[4];
    philox4x32_10(k0, k1, (uint32_t) g, (uint32_t) (g >> 32), 0x0FF5E7u, 0u, r);
    return r[0] % n;
}
__host__ __device__ inline uint32_t stream_msg_byte(uint32_t k0, uint32_t k1, uint64_t g, uint32_t m) {
    uint32_t r[4];
    philox4x32_10(k0, k1, (uint32_t) g, (uint32_t) (g >> 32), 0x4D5347u, m >> 4, r);
    return 0x20u + ((r[(m >> 2) & 3u] >> (8u * (m & 3u))) & 0xffu) % 95u;
}

__global__ void k_synth_streams(uint64_t seed, uint64_t first_stream, uint32_t nstreams, uint32_t nframes, size_t stream_stride,
                                uint32_t n, uint32_t lead_in, uint32_t msg_bytes, uint32_t guard,
                                const int32_t* __restrict__ table, int32_t gain, int32_t* __restrict__ pcm,
                                uint32_t* __restrict__ offsets, uint8_t* __restrict__ messages) {
    const uint32_t k0 = (uint32_t) seed, k1 = (uint32_t) (seed >> 32);
    const uint32_t pattern = lead_in + 8u + 8u * msg_bytes + guard;
    const size_t pairs = (size_t) nframes * (n / 2);
    for (uint32_t s = blockIdx.y; s < nstreams; s += gridDim.y) {
        const uint64_t g = first_stream + s;
        const uint32_t off = stream_offset(k0, k1, g, n);
        if (blockIdx.x == 0) {
            if (threadIdx.x == 0 && offsets) offsets[s] = off;
            if (messages)
                for (uint32_t m = threadIdx.x; m < msg_bytes; m += blockDim.x) messages[(size_t) s * msg_bytes + m] = (uint8_t) stream_msg_byte(k0, k1, g, m);
        }
        int32_t* dst = pcm + (size_t) s * stream_stride;
        for (size_t blk = (size_t) blockIdx.x * blockDim.x + threadIdx.x; blk < pairs; blk += (size_t) gridDim.x * blockDim.x) {
            uint32_t r[4];
            philox4x32_10(k0, k1, (uint32_t) g, (uint32_t) (g >> 32), (uint32_t) blk, 2u, r);
            const int32_t s0 = (int32_t) ((r[0] & 0xffffu) + (r[0] >> 16) + (r[1] & 0xffffu) + (r[1] >> 16)) - 131070;
            const int32_t s1 = (int32_t) ((r[2] & 0xffffu) + (r[2] >> 16) + (r[3] & 0xffffu) + (r[3] >> 16)) - 131070;
            int32_t v[2] = {(int32_t) (((int64_t) s0 * gain) / 65536), (int32_t) (((int64_t) s1 * gain) / 65536)};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const size_t i = 2 * blk + e;
                if (i >= off) {
                    const size_t t = i - off;
                    const uint32_t k = (uint32_t) ((t / n) % pattern), tau = (uint32_t) (t % n);
                    int kind = 0;                                     // 0 G, 1 H, 2 L
                    if (k >= lead_in && k < lead_in + 7u) kind = 1;
                    else if (k == lead_in + 7u) kind = 2;
                    else if (k > lead_in + 7u && k < lead_in + 8u + 8u * msg_bytes) {
                        const uint32_t b = k - (lead_in + 8u);
                        kind = ((stream_msg_byte(k0, k1, g, b >> 3) >> (7u - (b & 7u))) & 1u) ? 1 : 2;
                    }
                    if (kind) v[e] += table[(size_t) (kind - 1) * n + tau];
                }
                v[e] *= 256;
            }
            reinterpret_cast<int2*>(dst)[blk] = make_int2(v[0], v[1]);
        }
    }
}

// Band-limited rate conversion by the exact ratio up/down (the transmitter's 44.1 kHz track at the receiver's
// 78.125 kHz: 3125/1764): output j sits at input position j*down/up = k0 + p/up, all in integers, and is the
// ktaps-tap polyphase FIR  y = sum_i x[k0 - ktaps/2 + 1 + i] * h[p][i]  (taps ascending, one FMA each; samples
// outside the track are zero).  Result rounded to nearest and stored as a x256 DFSDM-style word.
__global__ void k_resample_i16(const int16_t* __restrict__ in, size_t n_in, uint32_t up, uint32_t down, uint32_t ktaps,
                               const float* __restrict__ taps, int32_t* __restrict__ out, size_t n_out) {
    for (size_t j = (size_t) blockIdx.x * blockDim.x + threadIdx.x; j < n_out; j += (size_t) gridDim.x * blockDim.x) {
        const unsigned long long pos = (unsigned long long) j * down;
        const long long k0 = (long long) (pos / up);
        const uint32_t ph = (uint32_t) (pos % up);
        const float* h = taps + (size_t) ph * ktaps;
        const long long first = k0 - (long long) (ktaps / 2) + 1;
        float acc = 0.0f;
        for (uint32_t i = 0; i < ktaps; ++i) {
            const long long k = first + i;
            const float x = (k >= 0 && (size_t) k < n_in) ? (float) in[k] : 0.0f;
            acc = __fmaf_rn(x, h[i], acc);
        }
        out[j] = __float2int_rn(acc) * 256;
    }
}

cudaError_t launch_resample_i16(const int16_t* in, size_t n_in, uint32_t up, uint32_t down, uint32_t ktaps, const float* taps,
                                int32_t* out, size_t n_out, cudaStream_t st) {
    size_t b = (n_out + 255) / 256;
    if (b > 148u * 32u) b = 148u * 32u;
    k_resample_i16<<<(int) (b ? b : 1), 256, 0, st>>>(in, n_in, up, down, ktaps, taps, out, n_out);
    return cudaGetLastError();
}

cudaError_t launch_synth_streams(uint64_t seed, uint64_t first_stream, uint32_t nstreams, uint32_t nframes, size_t stream_stride,
                                 uint32_t n, uint32_t lead_in, uint32_t msg_bytes, uint32_t guard, const int32_t* table,
                                 int32_t gain, int32_t* pcm, uint32_t* offsets, uint8_t* messages, cudaStream_t st) {
    const size_t pairs = (size_t) nframes * (n / 2);
    size_t bx = (pairs + 255) / 256;
    if (bx > 64) bx = 64;
    const uint32_t by = nstreams < 8192u ? nstreams : 8192u;
    k_synth_streams<<<dim3((unsigned) bx, by), 256, 0, st>>>(seed, first_stream, nstreams, nframes, stream_stride, n, lead_in,
                                                              msg_bytes, guard, table, gain, pcm, offsets, messages);
    return cudaGetLastError();
}

cudaError_t launch_synth_frames(uint64_t seed, uint64_t first_frame, size_t nframes, uint32_t n, const int32_t* table,
                                int32_t gain, int32_t* pcm, uint8_t* bits, cudaStream_t st) {
    const size_t total = nframes * (n / 2);
    size_t b = (total + 255) / 256;
    if (b > 148u * 32u) b = 148u * 32u;
    k_synth_frames<<<(int) (b ? b : 1), 256, 0, st>>>(seed, first_frame, nframes, n, table, gain, pcm, bits);
    return cudaGetLastError();
}

}  // namespace usc
