/* usc_tables.c — see usc_tables.h.  Plain C, compiled with -ffp-contract=off: the float/double
 * expression types below follow the reference's C source literally, because the exact table
 * values decide bit-parity of everything downstream. */
#include "usc_tables.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#define SINE_TABLE_SIZE 512
static float s_sine[SINE_TABLE_SIZE + 1];
static int s_sine_ready;

/* The CMSIS table (arm_common_tables.c, sinTable_f32) holds sin(2*pi*i/512) as decimal literals
 * with 8 fractional digits; reproducing the literals — not the exact sines — is what makes the
 * Hann window agree with the device captures to the last printed digit. */
static void sine_table_init(void) {
    if (s_sine_ready) return;
    char lit[32];
    for (int i = 0; i <= SINE_TABLE_SIZE; ++i) {
        snprintf(lit, sizeof lit, "%.8f", sin(2.0 * M_PI * (double) i / (double) SINE_TABLE_SIZE));
        s_sine[i] = strtof(lit, NULL);
    }
    s_sine_ready = 1;
}

float usc_host_arm_cos_f32(float x) {
    sine_table_init();
    float turns = x * 0.159154943092f + 0.25f;
    int32_t whole = (int32_t) turns;
    if (turns < 0.0f) whole--;
    turns = turns - (float) whole;
    float pos = (float) SINE_TABLE_SIZE * turns;
    uint16_t idx = ((uint16_t) pos) & 0x1ff;
    float frac = pos - (float) idx;
    return (1.0f - frac) * s_sine[idx] + frac * s_sine[idx + 1];
}

/* CMSIS-DSP V1.4.5 arm_sin_cos_f32 (arm_math.h:4634-4637), degrees in: turns of |theta|, the 512-entry table read at
 * the sine position and a quarter turn later, cubic interpolation whose end-point slopes are the other function's
 * table values times 2 pi/512; the sine takes theta's sign.  Plain float operations in this order (the vendored
 * archive's code has no fused multiply-add). */
static float hermite512(float y0, float y1, float s0, float s1, float frac) {
    const float step = 0.0122718463030f;
    const float rise = y1 - y0;
    float acc = step * (s0 + s1) - 2 * rise;
    acc = frac * acc + (3 * rise - (s1 + 2 * s0) * step);
    acc = frac * acc + s0 * step;
    return frac * acc + y0;
}

void usc_host_arm_sin_cos_f32(float theta_deg, float *s, float *c) {
    sine_table_init();
    float turns = theta_deg * 0.00277777777778f;
    if (turns < 0.0f) turns = -turns;
    turns = turns - (float) (int32_t) turns;
    const float pos = (float) SINE_TABLE_SIZE * turns;
    const uint16_t is = ((uint16_t) pos) & 0x1ff;
    const uint16_t ic = (uint16_t) ((is + SINE_TABLE_SIZE / 4) & 0x1ff);
    const float frac = pos - (float) is;
    *c = hermite512(s_sine[ic], s_sine[ic + 1], -s_sine[is], -s_sine[is + 1], frac);
    const float sv = hermite512(s_sine[is], s_sine[is + 1], s_sine[ic], s_sine[ic + 1], frac);
    *s = theta_deg < 0.0f ? -sv : sv;
}

void usc_host_hann(float *w, uint32_t n, uint32_t kind) {
    /* periodic: `2.0f * M_PI / (float) NN` is a double expression; symmetric: `2.0f * PI /
     * (float)(PCM_SAMPLES - 1)` with CMSIS's float PI is a float expression. */
    float scale = kind == 0 ? (float) (2.0f * M_PI / (float) n)
                            : 2.0f * 3.14159265358979f / (float) (n - 1);
    for (uint32_t i = 0; i < n; ++i) w[i] = 0.5f - 0.5f * usc_host_arm_cos_f32((float) i * scale);
}

void usc_host_ref_chirp(uint32_t variant, uint32_t n, float fs, float f0, float f1, float sweep_T,
                        float phase, int up, float *out) {
    float t = 0.0f;
    if (variant <= 1u) {
        /* degrees; `/ 2.0` and `360.0 *` promote to double before the store to float */
        float slope = (float) (f1 - f0) / sweep_T;
        float dt = sweep_T / (sweep_T * fs);
        for (uint32_t i = 0; i < n; ++i) {
            double half = (double) (slope * t) / 2.0;
            float freq = up ? (float) ((double) f0 + half) : (float) ((double) f1 - half);
            float theta = (float) (360.0 * (double) freq * (double) t + (double) phase);
            t = t + dt;
            float sv, cv;
            usc_host_arm_sin_cos_f32(theta, &sv, &cv);
            if (variant == 0u) out[i] = sv * 1.0f;
            else { out[2 * i] = cv * 1.0f; out[2 * i + 1] = sv * 1.0f; }
        }
    } else {
        /* radians, float slope law without the /2, T = n/fs */
        const double two_pi = 2.0 * (double) 3.14159265358979f;
        float T = (float) n / fs;
        float slope = (f1 - f0) / T;
        float dt = T / (T * fs);
        for (uint32_t i = 0; i < n; ++i) {
            float freq = up ? f0 + slope * t : f1 - slope * t;
            double a = two_pi * (double) freq * (double) t;
            float arg = variant == 2u ? (float) (a + (double) phase) : (float) a;
            t = t + dt;
            out[i] = usc_host_arm_cos_f32(arg) * 1.0f;
        }
    }
}

void usc_host_twiddles(float *tw, uint32_t n) {
    for (uint32_t j = 0; j < n; ++j) {
        double a = 2.0 * M_PI * (double) j / (double) n;
        tw[2 * j] = (float) cos(a);
        tw[2 * j + 1] = (float) -sin(a);
    }
}

uint32_t usc_host_radices(uint32_t n, uint32_t *rad) {
    if (n < 2 || (n & (n - 1))) return 0;
    uint32_t rev[USC_MAX_RADICES], cnt = 0;
    while (n > 32) {
        if (cnt >= USC_MAX_RADICES - 1) return 0;
        rev[cnt++] = 32;
        n /= 32;
    }
    rev[cnt++] = n;
    for (uint32_t i = 0; i < cnt; ++i) rad[i] = rev[cnt - 1 - i];
    return cnt;
}

uint32_t usc_host_bandwidth(uint32_t n, float fs, float f0, float f1) {
    unsigned long span = (unsigned long) (int) (f1 - f0) * (unsigned long) n;
    return (uint32_t) ((float) span / fs);
}

void usc_host_symbol_tables(uint32_t n, float fs, float f0, float f1, double amp, int32_t *out) {
    const double T = (double) n / (double) fs, k = ((double) f1 - (double) f0) / T;
    for (int down = 0; down < 2; ++down)
        for (uint32_t i = 0; i < n; ++i) {
            const double t = T * (double) i / (double) (n - 1);
            const double f = down ? (double) f1 - k * t / 2.0 : (double) f0 + k * t / 2.0;
            const double arg = 2.0 * M_PI * f * t - M_PI / 2.0;
            out[(size_t) down * n + i] = (int32_t) llround(amp * (cos(arg) + sin(arg)));
        }
}

void usc_host_iq_symbol_tables(uint32_t n, float fs, double carrier, double bw, int sideband, double phase, double amp,
                               int32_t *out) {
    const double T = (double) n / (double) fs, k = bw / T, sgn = sideband < 0 ? -1.0 : 1.0;
    for (int down = 0; down < 2; ++down)
        for (uint32_t i = 0; i < n; ++i) {
            const double t = T * (double) i / (double) (n - 1);
            const double fb = down ? bw / 2.0 - k * t / 2.0 : -bw / 2.0 + k * t / 2.0;
            const double arg = (2.0 * M_PI * (carrier + sgn * fb) * t) + phase;
            out[(size_t) down * n + i] = (int32_t) llround(amp * cos(arg));
        }
}

int32_t usc_host_noise_gain(double sigma) {
    /* sum of four independent 16-bit uniforms: variance 4 * (65536^2 - 1) / 12 */
    const double unit = sqrt(4.0 * (65536.0 * 65536.0 - 1.0) / 12.0);
    return (int32_t) llround(sigma / unit * 65536.0);
}

void usc_host_resample_taps(uint32_t up, uint32_t ktaps, float *taps) {
    for (uint32_t p = 0; p < up; ++p)
        for (uint32_t i = 0; i < ktaps; ++i) {
            const double x = ((double) i - (double) (ktaps / 2) + 1.0) - (double) p / (double) up;
            const double s = x == 0.0 ? 1.0 : sin(M_PI * x) / (M_PI * x);
            const double w = 0.5 + 0.5 * cos(2.0 * M_PI * x / (double) ktaps);
            taps[(size_t) p * ktaps + i] = (float) (s * w);
        }
}
