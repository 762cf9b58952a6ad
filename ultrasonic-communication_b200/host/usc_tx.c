/* usc_tx.c — transmitter symbols, framing and WAV I/O (include/usc_tx.h). */
#include "usc_tx.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

uint32_t usc_tx_symbol_len(double fs, double T) { return (uint32_t) (T * fs); }      /* int(self.T * self.fs) */

int usc_tx_symbol(double fs, double f0, double f1, double T, double A, int kind, double *out, uint32_t cap) {
    const uint32_t n = usc_tx_symbol_len(fs, T);
    if (!out || n < 2 || cap < n || kind < 0 || kind > 2) return -1;
    if (kind == 0) {                                    /* silence(), signal.py:55-56 */
        memset(out, 0, sizeof(double) * n);
        return (int) n;
    }
    const double step = T / (double) (n - 1), k = (f1 - f0) / T;       /* numpy.linspace(0, T, n) */
    for (uint32_t i = 0; i < n; ++i) {
        const double t = i == n - 1 ? T : (double) i * step;
        const double f = kind == 1 ? f0 + k * t / 2.0 : f1 - k * t / 2.0;
        const double arg = (2.0 * M_PI * f * t) + (-M_PI / 2.0);
        out[i] = (cos(arg) + sin(arg)) * A;             /* chirp_orth(), signal.py:45-53 */
    }
    return (int) n;
}

int usc_tx_symbol_iq(double fs, double bw, double fc, double T, double A, double phase, int kind, double *out, uint32_t cap) {
    const uint32_t n = usc_tx_symbol_len(fs, T);
    if (!out || n < 2 || cap < n || kind < 0 || kind > 2) return -1;
    if (kind == 0) {
        memset(out, 0, sizeof(double) * n);
        return (int) n;
    }
    const double f0 = -bw / 2.0, f1 = bw / 2.0;                        /* cell 3 */
    const double step = T / (double) (n - 1), k = (f1 - f0) / T;       /* numpy.linspace(0, T, n) */
    for (uint32_t i = 0; i < n; ++i) {
        const double t = i == n - 1 ? T : (double) i * step;
        const double fb = kind == 1 ? f0 + k * t / 2.0 : f1 - k * t / 2.0;
        const double arg = (2.0 * M_PI * (fc + fb) * t) + phase;
        out[i] = cos(arg) * A;                                          /* chirp_iq(), cell 5 */
    }
    return (int) n;
}

size_t usc_tx_frame_len(double fs, double T, uint32_t msg_len, uint32_t guard) {
    return (size_t) usc_tx_symbol_len(fs, T) * (1u + 7u + 1u + 8u * (size_t) msg_len + guard);
}

long usc_tx_frame_i16(double fs, double f0, double f1, double T, double A, const uint8_t *msg, uint32_t msg_len,
                      uint32_t guard, int16_t *out, size_t cap) {
    const uint32_t n = usc_tx_symbol_len(fs, T);
    const size_t total = usc_tx_frame_len(fs, T, msg_len, guard);
    if (!out || (!msg && msg_len) || n < 2 || cap < total) return -1;
    double *sym = (double *) malloc(sizeof(double) * 2 * n);
    if (!sym) return -1;
    usc_tx_symbol(fs, f0, f1, T, A, 1, sym, n);
    usc_tx_symbol(fs, f0, f1, T, A, 2, sym + n, n);
    size_t w = 0;
    /* ChirpGenerator.ipynb cell 3: tone = G, then PREAMBLE (7 H), DELIMITER (L), DATA, GUARD */
#define PUT(kind)                                                                          \
    do {                                                                                   \
        for (uint32_t i_ = 0; i_ < n; ++i_)                                                \
            out[w + i_] = (kind) == 0 ? 0 : (int16_t) sym[((kind) - 1) * (size_t) n + i_]; \
        w += n;                                                                            \
    } while (0)
    PUT(0);
    for (int i = 0; i < 7; ++i) PUT(1);
    PUT(2);
    for (uint32_t m = 0; m < msg_len; ++m)
        for (int b = 0; b < 8; ++b) {                   /* ascii(): a & (0b10000000 >> i) */
            if (msg[m] & (0x80u >> b)) PUT(1);
            else PUT(2);
        }
    for (uint32_t g = 0; g < guard; ++g) PUT(0);
#undef PUT
    free(sym);
    return (long) w;
}

static void put_le(unsigned char *p, uint32_t v, int bytes) {
    for (int i = 0; i < bytes; ++i) p[i] = (unsigned char) (v >> (8 * i));
}

long usc_wav_write_i16(const char *path, uint32_t fs, const int16_t *x, size_t n) {
    if (!path || (!x && n) || n > 0x7fffffffu / 2) return -1;
    FILE *f = fopen(path, "wb");
    if (!f) return -2;
    unsigned char h[44];
    const uint32_t data = (uint32_t) (n * 2);
    memcpy(h, "RIFF", 4); put_le(h + 4, 36 + data, 4); memcpy(h + 8, "WAVEfmt ", 8);
    put_le(h + 16, 16, 4); put_le(h + 20, 1, 2); put_le(h + 22, 1, 2);          /* PCM, mono */
    put_le(h + 24, fs, 4); put_le(h + 28, fs * 2, 4); put_le(h + 32, 2, 2); put_le(h + 34, 16, 2);
    memcpy(h + 36, "data", 4); put_le(h + 40, data, 4);
    int bad = fwrite(h, 1, 44, f) != 44;
    for (size_t i = 0; i < n && !bad; ++i) {
        unsigned char s[2];
        put_le(s, (uint16_t) x[i], 2);
        bad = fwrite(s, 1, 2, f) != 2;
    }
    return fclose(f) || bad ? -2 : (long) n;
}

static uint32_t get_le(const unsigned char *p, int bytes) {
    uint32_t v = 0;
    for (int i = 0; i < bytes; ++i) v |= (uint32_t) p[i] << (8 * i);
    return v;
}

long usc_wav_read_i16(const char *path, uint32_t *fs, int16_t *x, size_t cap) {
    if (!path) return -1;
    FILE *f = fopen(path, "rb");
    if (!f) return -2;
    unsigned char h[12];
    if (fread(h, 1, 12, f) != 12 || memcmp(h, "RIFF", 4) || memcmp(h + 8, "WAVE", 4)) { fclose(f); return -3; }
    int have_fmt = 0;
    long nsamp = -3;
    for (;;) {                                          /* chunk walk: fmt, then data (others skipped) */
        unsigned char c[8];
        if (fread(c, 1, 8, f) != 8) break;
        const uint32_t len = get_le(c + 4, 4);
        if (!memcmp(c, "fmt ", 4)) {
            unsigned char m[16];
            if (len < 16 || fread(m, 1, 16, f) != 16) break;
            if (get_le(m, 2) != 1 || get_le(m + 2, 2) != 1 || get_le(m + 14, 2) != 16) break;   /* PCM, mono, 16 bit */
            if (fs) *fs = get_le(m + 4, 4);
            if (len > 16) fseek(f, (long) (len - 16 + (len & 1)), SEEK_CUR);
            have_fmt = 1;
        } else if (!memcmp(c, "data", 4)) {
            if (!have_fmt) break;
            nsamp = (long) (len / 2);
            for (long i = 0; i < nsamp && x && (size_t) i < cap; ++i) {
                unsigned char s[2];
                if (fread(s, 1, 2, f) != 2) { nsamp = -3; break; }
                x[i] = (int16_t) get_le(s, 2);
            }
            break;
        } else {
            fseek(f, (long) (len + (len & 1)), SEEK_CUR);
        }
    }
    fclose(f);
    return nsamp;
}
