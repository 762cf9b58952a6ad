/*
 * shim_chain.c — a reference-style caller of the CMSIS names, linked against libusc_cmsis.so.
 * It performs the call sequence of the receiver's pipeline()+dsp() (receiver/Src/main.c:163-215) and
 * of compress_chirp() (experiments/chirp_compression_time_domain/Src/chirp.c:78-83) on one frame
 * read from a file, exactly as firmware code written against arm_math.h would, and prints the
 * results for the test to compare with the oracle.
 *
 * usage: shim_chain <frame.f32> <chirp.f32> <hann.f32> <H.f32>
 */
#include <stdio.h>
#include <stdlib.h>

#include "usc_cmsis_shim.h"

#define NN 2048

static int load(const char *path, float *dst, size_t n) {
    FILE *f = fopen(path, "rb");
    if (!f) return 0;
    size_t got = fread(dst, sizeof(float), n, f);
    fclose(f);
    return got == n;
}

int main(int argc, char **argv) {
    static float frame[NN], chirp[NN], hann[NN], H[NN], work[2 * NN], spectrum[NN];
    if (argc < 5 || !load(argv[1], frame, NN) || !load(argv[2], chirp, NN) || !load(argv[3], hann, NN) ||
        !load(argv[4], H, NN)) {
        fprintf(stderr, "usage: shim_chain frame chirp hann H (raw float32 x %d)\n", NN);
        return 2;
    }
    arm_rfft_fast_instance_f32 S;
    if (arm_rfft_fast_init_f32(&S, NN) != ARM_MATH_SUCCESS) return 3;
    if (arm_rfft_fast_init_f32(&S, 1000) != ARM_MATH_ARGUMENT_ERROR) return 4;
    arm_rfft_fast_init_f32(&S, NN);

    /* receiver chain: de-chirp, window, RFFT, magnitude, windowed arg-max */
    arm_copy_f32(frame, work, NN);
    arm_mult_f32(work, chirp, work, NN);
    arm_mult_f32(work, hann, work, NN);
    arm_rfft_fast_f32(&S, work, spectrum, 0);
    arm_cmplx_mag_f32(spectrum, work, NN / 2);
    float peak;
    uint32_t bin;
    arm_max_f32(work, 156, &peak, &bin);
    float mean;
    arm_mean_f32(work, 8, &mean);
    printf("receiver %a %u %a\n", peak, bin, mean);

    /* compression chain: window, RFFT in place, packed complex product, inverse RFFT in place, max */
    arm_copy_f32(frame, work, NN);
    arm_mult_f32(work, hann, work, NN);
    arm_rfft_fast_f32(&S, work, work, 0);
    arm_cmplx_mult_cmplx_f32(work, H, work, NN / 2);
    arm_rfft_fast_f32(&S, work, work, 1);
    arm_max_f32(work, NN, &peak, &bin);
    printf("compress %a %u\n", peak, bin);

    /* complex FFT round trip on the const-struct instance */
    for (int i = 0; i < NN; ++i) { work[2 * i] = frame[i]; work[2 * i + 1] = 0.0f; }
    arm_cfft_f32(&arm_cfft_sR_f32_len2048, work, 0, 1);
    arm_cmplx_mag_f32(work, spectrum, NN);
    arm_max_f32(spectrum, NN, &peak, &bin);
    printf("cfft %a %u\n", peak, bin);
    printf("status %d\n", usc_cmsis_last_status());
    return usc_cmsis_last_status() == 0 ? 0 : 5;
}
