/* analyser_host.c — the audio spectrum analyser of experiments/basic (fft(), Src/main.c:107-175) as a
 * plain-C host program over the C-ABI and the CSV wire formats: reads a captured <name>.raw
 * ("Index,Amplitude" rows, agent/README.md:5-11), runs cast -> Hann -> RFFT -> magnitude/sqrt(N) ->
 * AC coupling -> dB on the GPU, writes <out>.fft and <out>.flt the way the PC agent stores them and prints
 * the firmware's whole UART dump to stdout.  Links against libusc.so and libusc_wire.so only.
 *   usage: analyser_host <capture.raw> <fs_hz> <out_prefix> [mic]                                         */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "usc.h"
#include "usc_wire.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        int rc_ = (call);                                                             \
        if (rc_ != USC_OK) {                                                          \
            fprintf(stderr, "%s failed: %s\n", #call, usc_error_string(rc_));         \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

int main(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s <capture.raw> <fs_hz> <out_prefix> [mic]\n", argv[0]);
        return 2;
    }
    usc_config cfg;
    usc_default_config(&cfg);                                   /* PCM_SAMPLES 2048, periodic Hann */
    cfg.fs = (float) atof(argv[2]);
    const uint32_t n = cfg.n, half = n / 2;
    int32_t *pcm = (int32_t *) malloc(sizeof(int32_t) * n);
    int rows = usc_wire_read_raw(argv[1], n, pcm);
    if (rows != (int) n) {
        fprintf(stderr, "%s: expected %u rows, got %d\n", argv[1], n, rows);
        return 1;
    }
    usc_handle *h = NULL;
    CHECK(usc_create(&cfg, 0, &h));
    void *d_pcm = NULL, *d_x = NULL, *d_hann = NULL, *d_mag = NULL, *d_db = NULL;
    CHECK(usc_malloc(&d_pcm, n * 4)); CHECK(usc_malloc(&d_x, n * 4)); CHECK(usc_malloc(&d_hann, n * 4));
    CHECK(usc_malloc(&d_mag, half * 4)); CHECK(usc_malloc(&d_db, half * 4));
    float *hann = (float *) malloc(sizeof(float) * n), *win = (float *) malloc(sizeof(float) * n);
    float *mag = (float *) malloc(sizeof(float) * half), *db = (float *) malloc(sizeof(float) * half);
    if (usc_get_table(h, "hann", hann, n) < 0) return 1;
    CHECK(usc_memcpy_h2d(h, d_pcm, pcm, n * 4));
    CHECK(usc_memcpy_h2d(h, d_hann, hann, n * 4));
    /* fft_hanning[] (main.c:110-113): the windowed samples the firmware also dumps */
    CHECK(usc_i32_to_f32(h, (const int32_t *) d_pcm, (float *) d_x, n));
    CHECK(usc_arm_mult_f32_batch(h, (const float *) d_x, n, (const float *) d_hann, 0, (float *) d_x, n, n, 1));
    CHECK(usc_spectrum_analyzer(h, d_pcm, USC_PCM_I32, 1, 1000.0f /* FFT_AC_COUPLING_HZ */, (float *) d_mag, (float *) d_db, NULL, NULL));
    CHECK(usc_memcpy_d2h(h, win, d_x, n * 4));
    CHECK(usc_memcpy_d2h(h, mag, d_mag, half * 4));
    CHECK(usc_memcpy_d2h(h, db, d_db, half * 4));
    CHECK(usc_sync(h));
    char path[1024];
    snprintf(path, sizeof path, "%s.fft", argv[3]);
    if (usc_wire_write_fft(path, cfg.fs, n, mag, db) < 0) return 1;
    snprintf(path, sizeof path, "%s.flt", argv[3]);
    if (usc_wire_write_flt(path, win, n) < 0) return 1;
    if (usc_wire_write_dump(stdout, argc > 4 ? argv[4] : "M1", cfg.fs, n, mag, db, pcm, win) < 0) return 1;
    usc_free(d_pcm); usc_free(d_x); usc_free(d_hann); usc_free(d_mag); usc_free(d_db);
    free(pcm); free(hann); free(win); free(mag); free(db);
    usc_destroy(h);
    return 0;
}
