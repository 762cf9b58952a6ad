/* usc_tables.h — host-side (plain C) builders of the constant tables the kernels consume.
 * Product code: replaces the init-time table generation of the reference firmware
 * (receiver/Src/main.c:372-374,390-393; receiver/Src/chirp.c:16-45 and the experiment variants). */
#ifndef USC_TABLES_H_
#define USC_TABLES_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define USC_MAX_RADICES 8

/* arm_cos_f32 (arm_math.h:5685): CMSIS-DSP V1.4.5 512-entry table + linear interpolation. */
float usc_host_arm_cos_f32(float x);
/* arm_sin_cos_f32 (arm_math.h:4634-4637), degrees. */
void usc_host_arm_sin_cos_f32(float theta_deg, float *s, float *c);
/* Hann window: kind 0 periodic (receiver/Src/main.c:99,390-393), 1 symmetric
 * (experiments/chirp_compression_time_domain/Src/chirp.c:13,63-65). */
void usc_host_hann(float *w, uint32_t n, uint32_t kind);
/* Reference chirp tables, variants R/S/T/F (see usc.h). out: n floats (2n for S). */
void usc_host_ref_chirp(uint32_t variant, uint32_t n, float fs, float f0, float f1, float sweep_T,
                        float phase, int up, float *out);
/* Master twiddle table: (cos, -sin)(2*pi*j/n), j < n, each rounded once from double. 2n floats. */
void usc_host_twiddles(float *tw, uint32_t n);
/* Radix list of the canonical mixed-radix plan. Returns the count, 0 if bad. */
uint32_t usc_host_radices(uint32_t n, uint32_t *rad);
/* Transmitter symbol tables for the synthetic generator: round(amp * (cos(arg) + sin(arg))) with the
 * chirp_orth law of simulation/signal.py:45-53 (t = linspace(0, T, n), T = n/fs, phase -pi/2), evaluated
 * in double.  out: 2n int32 — the up symbol then the down symbol. */
void usc_host_symbol_tables(uint32_t n, float fs, float f0, float f1, double amp, int32_t *out);
/* I/Q transmitter symbols (generator/ChirpGeneratorIQmodulation.ipynb cell 5, simulation/IQ_modulation.ipynb cell 4):
 * out[down*n + i] = round(amp * cos(2 pi (fc + sideband * fb(t_i)) t_i + phase)), fb = -bw/2 + k t/2 (up) or
 * +bw/2 - k t/2 (down), k = bw/T, t = linspace(0, T, n), T = n/fs */
void usc_host_iq_symbol_tables(uint32_t n, float fs, double carrier, double bw, int sideband, double phase, double amp,
                               int32_t *out);
/* noise gain of the integer generator for a target standard deviation (PCM units before the x256) */
int32_t usc_host_noise_gain(double sigma);
/* (F1 - F0) * NN / fs truncated to uint32 (receiver/Src/main.c:372). */
uint32_t usc_host_bandwidth(uint32_t n, float fs, float f0, float f1);

/* Polyphase table of the band-limited resampler (usc_resample_i16_to_pcm): `up` phases x `ktaps` taps,
 * h[p][i] = sinc(x) * (0.5 + 0.5 cos(2 pi x / ktaps)), x = (i - ktaps/2 + 1) - p/up, in double, rounded to float. */
void usc_host_resample_taps(uint32_t up, uint32_t ktaps, float *taps);

#ifdef __cplusplus
}
#endif
#endif
