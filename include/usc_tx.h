/* usc_tx.h — the transmitter side, plain host C except the resampler's device call.
 *
 * generator/ChirpGenerator.ipynb cells 1-3 build the audio the speaker plays: symbols H (up) / L (down) /
 * G (silence) of simulation/signal.py:45-56 (chirp_orth, silence) at fs = 44100, T = 0.0262 s, A = 20000,
 * framed as G + 7 x H + L + message bits (MSB first, H = 1) + 12 x G and written as a 16-bit WAV file
 * (Signal.play(), signal.py:122-124).  The receiver samples the same sound at 78125 Hz: the resampler
 * (usc.h: usc_resample_i16_to_pcm) renders the 44.1 kHz track at the receiver's rate, 3125/1764 exactly.
 * Returns: counts >= 0, or -1 argument, -2 I/O, -3 format. */
#ifndef USC_TX_H
#define USC_TX_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* int(T * fs): samples per symbol (1155 at the notebook's settings) */
uint32_t usc_tx_symbol_len(double fs, double T);
/* kind: 0 G, 1 H, 2 L.  out[i] = A (cos(arg) + sin(arg)), arg = 2 pi f(t) t - pi/2 on t = linspace(0, T, n). */
int usc_tx_symbol(double fs, double f0, double f1, double T, double A, int kind, double *out, uint32_t cap);
/* The I/Q transmitter's symbol, generator/ChirpGeneratorIQmodulation.ipynb cell 5 chirp_iq(): out[i] =
 * A cos(2 pi (fc + fb) t + phase), fb = -bw/2 + k t/2 (kind 1, up) or +bw/2 - k t/2 (kind 2, down), k = bw/T,
 * t = linspace(0, T, n); kind 0 = silence.  The notebook's tone (cell 9) is 100 x the up symbol as int16. */
int usc_tx_symbol_iq(double fs, double bw, double fc, double T, double A, double phase, int kind, double *out, uint32_t cap);
/* number of samples of the framed message: (1 + 7 + 1 + 8*msg_len + guard) symbols */
size_t usc_tx_frame_len(double fs, double T, uint32_t msg_len, uint32_t guard);
/* the whole tone as int16 (C truncation == numpy astype(int16)); guard = 12 in the notebook */
long usc_tx_frame_i16(double fs, double f0, double f1, double T, double A, const uint8_t *msg, uint32_t msg_len,
                      uint32_t guard, int16_t *out, size_t cap);
/* 16-bit mono PCM WAV, the layout scipy.io.wavfile.write produces (44-byte header) */
long usc_wav_write_i16(const char *path, uint32_t fs, const int16_t *x, size_t n);
long usc_wav_read_i16(const char *path, uint32_t *fs, int16_t *x, size_t cap);   /* -> samples in the file */

#ifdef __cplusplus
}
#endif
#endif
