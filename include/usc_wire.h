/* usc_wire.h — the reference's UART / file wire formats, plain host C, no CUDA.
 *
 * The firmware prints its results as CSV text over the UART and a PC agent stores the sections as
 * files (agent/README.md:5-11: *.fft, *.raw, *.flt).  These readers and writers let the host driver
 * ingest real captures and emit logs the reference's notebooks can plot unchanged.
 *   analyser dump      experiments/basic/Src/main.c:144-172
 *   history tables     receiver/Src/main.c:276-300 (print_history, SIMPLE and DETAIL modes)
 * All functions return a count >= 0 or a negative error (-1 argument, -2 I/O, -3 format). */
#ifndef USC_WIRE_H
#define USC_WIRE_H
#include <stdint.h>
#include <stdio.h>
#include "usc.h"
#ifdef __cplusplus
extern "C" {
#endif

/* One whole analyser dump as the firmware prints it when the user button is pressed: header lines, the
 * "Frequency(Hz),Magnitude,Magnitude(dB)" table (n/2 rows "%.1f,%f,%f", frequency = i*fs/n), a blank
 * line, "Index,Amplitude" with the n raw words ("%lu,%ld"), "EORAW", "Index,Amplitude" with the n
 * windowed samples ("%lu,%f"), "EOFLT".  The peak line is derived from mag (first maximum). */
int usc_wire_write_dump(FILE *f, const char *mic, float fs, uint32_t n, const float *mag, const float *db,
                        const int32_t *pcm, const float *windowed);
/* Reads such a dump back (any of the output pointers may be NULL).  mic receives the text after
 * "MEMS mic: ".  Returns n/2 on success. */
int usc_wire_read_dump(FILE *f, uint32_t n, char *mic, size_t mic_cap, float *freq, float *mag, float *db,
                       int32_t *pcm, float *windowed);
/* The three per-section files of the PC agent (each starts with its CSV header line). */
int usc_wire_write_fft(const char *path, float fs, uint32_t n, const float *mag, const float *db);
int usc_wire_write_raw(const char *path, const int32_t *pcm, uint32_t n);
int usc_wire_write_flt(const char *path, const float *x, uint32_t n);
int usc_wire_read_fft(const char *path, uint32_t max_rows, float *freq, float *mag, float *db);   /* -> rows */
int usc_wire_read_raw(const char *path, uint32_t max_rows, int32_t *pcm);
int usc_wire_read_flt(const char *path, uint32_t max_rows, float *x);
/* print_history(): detail = 0 -> SIMPLE ("I => G" and "rank,snr" rows), 1 -> DETAIL (the 11-column
 * table; the two tick columns of the firmware are printed as 0).  States: 0 IDLE, 1 SYNCHRONIZING,
 * 2 SYNCHRONIZED, 3 DATA_RECEIVING. */
int usc_wire_write_history(FILE *f, int detail, uint32_t prev_state, uint32_t state, const usc_history *hist,
                           uint32_t num);
/* Parses the rows of a SIMPLE or DETAIL table written by the firmware or by the function above into
 * hist[] (fields the table does not carry are zeroed).  Returns the number of rows. */
int usc_wire_read_history(FILE *f, int detail, uint32_t *prev_state, uint32_t *state, usc_history *hist, uint32_t max_rows);

#ifdef __cplusplus
}
#endif
#endif
