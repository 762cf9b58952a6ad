/*
 * receiver_host.c — plain-C host driver over libusc.so: the reference firmware's receive path
 * (receiver/Src/main.c) for a batch of streams, written against include/usc.h only (no CUDA
 * headers).  It mirrors what main() does around the DSP chain:
 *
 *   init:    arm_rfft_fast_init_f32 + init_ref_chirp + Hann loop   (main.c:377-393)  -> usc_create
 *   ingest:  DMA buffer of int32 words                              (main.c:659-668)  -> usc_memcpy_h2d
 *   run:     the while(1) state machine                             (main.c:417-580)  -> usc_receiver_run
 *   output:  printf("%c", msg)                                       (main.c:533)      -> the uart bytes
 *
 * Usage: receiver_host <pcm.i32> <nstreams> <nframes>     (raw little-endian int32, stream-major)
 * Build: gcc -O2 -I include host/receiver_host.c -L. -lusc -Wl,-rpath,'$ORIGIN' (see build.py)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "usc.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc__ = (call);                                                           \
        if (rc__ != USC_OK) {                                                        \
            fprintf(stderr, "%s failed: %s\n", #call, usc_error_string(rc__));       \
            return 1;                                                                \
        }                                                                            \
    } while (0)

int main(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s <pcm.i32> <nstreams> <nframes>\n", argv[0]);
        return 2;
    }
    const uint32_t nstreams = (uint32_t) strtoul(argv[2], NULL, 10);
    const uint32_t nframes = (uint32_t) strtoul(argv[3], NULL, 10);
    const uint32_t uart_cap = 256;

    usc_config cfg;
    usc_default_config(&cfg);                 /* NN 2048, fs 78125, F0/F1 16/19 kHz, SNR_THRESHOLD 2.0 */
    usc_handle *h = NULL;
    CHECK(usc_create(&cfg, 0, &h));

    const size_t samples = (size_t) nstreams * nframes * cfg.n;
    int32_t *pcm_host = NULL;
    CHECK(usc_malloc_host((void **) &pcm_host, samples * sizeof(int32_t)));
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(pcm_host, sizeof(int32_t), samples, f) != samples) {
        fprintf(stderr, "cannot read %zu samples from %s\n", samples, argv[1]);
        return 1;
    }
    fclose(f);

    void *pcm_dev = NULL, *uart_dev = NULL, *res_dev = NULL;
    CHECK(usc_malloc(&pcm_dev, samples * sizeof(int32_t)));
    CHECK(usc_malloc(&uart_dev, (size_t) nstreams * uart_cap));
    CHECK(usc_malloc(&res_dev, (size_t) nstreams * sizeof(usc_rx_result)));
    CHECK(usc_memcpy_h2d(h, pcm_dev, pcm_host, samples * sizeof(int32_t)));
    CHECK(usc_receiver_run(h, pcm_dev, USC_PCM_I32, nstreams, nframes, (size_t) nframes * cfg.n,
                           (uint8_t *) uart_dev, uart_cap, (usc_rx_result *) res_dev));

    uint8_t *uart = (uint8_t *) malloc((size_t) nstreams * uart_cap);
    usc_rx_result *res = (usc_rx_result *) malloc((size_t) nstreams * sizeof(usc_rx_result));
    CHECK(usc_memcpy_d2h(h, uart, uart_dev, (size_t) nstreams * uart_cap));
    CHECK(usc_memcpy_d2h(h, res, res_dev, (size_t) nstreams * sizeof(usc_rx_result)));
    CHECK(usc_sync(h));

    for (uint32_t s = 0; s < nstreams; ++s) {
        uint32_t nb = res[s].nbytes < uart_cap ? res[s].nbytes : uart_cap;
        printf("stream %u: lock_frame=%d sync_position=%u state=%u bytes=%u: ", s, res[s].lock_frame,
               res[s].lock_position, res[s].state, res[s].nbytes);
        fwrite(uart + (size_t) s * uart_cap, 1, nb, stdout);
        if (!nb || uart[(size_t) s * uart_cap + nb - 1] != '\n') putchar('\n');
    }
    free(uart);
    free(res);
    usc_free(pcm_dev); usc_free(uart_dev); usc_free(res_dev);
    usc_free_host(pcm_host);
    usc_destroy(h);
    return 0;
}
