/* usc_wire.c — readers and writers of the reference's UART / file CSV formats (include/usc_wire.h). */
#include "usc_wire.h"

#include <stdlib.h>
#include <string.h>

static const char STATE_NAME[4][16] = {"IDLE", "SYNCHRONIZING", "SYNCHRONIZED", "DATA_RECEIVING"};   /* receiver/Src/main.c:113 */
static const char STATE_CHAR[4] = {'I', 'G', 'S', 'R'};                                              /* :115 */

static void fft_rows(FILE *f, float fs, uint32_t n, const float *mag, const float *db) {
    fprintf(f, "Frequency(Hz),Magnitude,Magnitude(dB)\n");
    for (uint32_t i = 0; i < n / 2; ++i) {
        float fr = (float) i * fs / (float) n;                     /* fft_frequency[], basic/Src/main.c:248-250 */
        fprintf(f, "%.1f,%f,%f\n", (double) fr, (double) mag[i], (double) (db ? db[i] : 0.0f));
    }
}

int usc_wire_write_dump(FILE *f, const char *mic, float fs, uint32_t n, const float *mag, const float *db,
                        const int32_t *pcm, const float *windowed) {
    if (!f || !mag || !pcm || !windowed || n < 2) return -1;
    uint32_t imax = 0;
    for (uint32_t i = 1; i < n / 2; ++i)
        if (mag[i] > mag[imax]) imax = i;                          /* arm_max_f32: first maximum */
    fprintf(f, "\nMEMS mic: %s\n", mic ? mic : "");
    fprintf(f, "Frequency at max magnitude: %.1f, Max magnitude: %f\n", (double) ((float) imax * fs / (float) n),
            (double) mag[imax]);
    fft_rows(f, fs, n, mag, db);
    fprintf(f, "\n");
    fprintf(f, "Index,Amplitude\n");
    for (uint32_t i = 0; i < n; ++i) fprintf(f, "%lu,%ld\n", (unsigned long) i, (long) pcm[i]);
    fprintf(f, "EORAW\n");
    fprintf(f, "Index,Amplitude\n");
    for (uint32_t i = 0; i < n; ++i) fprintf(f, "%lu,%f\n", (unsigned long) i, (double) windowed[i]);
    fprintf(f, "EOFLT\n");
    return ferror(f) ? -2 : (int) (n / 2);
}

static int next_line(FILE *f, char *buf, size_t cap) {
    if (!fgets(buf, (int) cap, f)) return 0;
    size_t l = strlen(buf);
    while (l && (buf[l - 1] == '\n' || buf[l - 1] == '\r')) buf[--l] = 0;
    return 1;
}

int usc_wire_read_dump(FILE *f, uint32_t n, char *mic, size_t mic_cap, float *freq, float *mag, float *db,
                       int32_t *pcm, float *windowed) {
    if (!f || n < 2) return -1;
    char line[256];
    int seen = 0;
    while (next_line(f, line, sizeof line)) {                      /* header lines up to the FFT table */
        if (!strncmp(line, "MEMS mic: ", 10) && mic && mic_cap) {
            strncpy(mic, line + 10, mic_cap - 1);
            mic[mic_cap - 1] = 0;
        }
        if (!strcmp(line, "Frequency(Hz),Magnitude,Magnitude(dB)")) { seen = 1; break; }
    }
    if (!seen) return -3;
    for (uint32_t i = 0; i < n / 2; ++i) {
        float a, b, c;
        if (!next_line(f, line, sizeof line) || sscanf(line, "%f,%f,%f", &a, &b, &c) != 3) return -3;
        if (freq) freq[i] = a;
        if (mag) mag[i] = b;
        if (db) db[i] = c;
    }
    if (!next_line(f, line, sizeof line) || line[0] != 0) return -3;                      /* blank line */
    if (!next_line(f, line, sizeof line) || strcmp(line, "Index,Amplitude")) return -3;
    for (uint32_t i = 0; i < n; ++i) {
        unsigned long idx; long v;
        if (!next_line(f, line, sizeof line) || sscanf(line, "%lu,%ld", &idx, &v) != 2 || idx != i) return -3;
        if (pcm) pcm[i] = (int32_t) v;
    }
    if (!next_line(f, line, sizeof line) || strcmp(line, "EORAW")) return -3;
    if (!next_line(f, line, sizeof line) || strcmp(line, "Index,Amplitude")) return -3;
    for (uint32_t i = 0; i < n; ++i) {
        unsigned long idx; float v;
        if (!next_line(f, line, sizeof line) || sscanf(line, "%lu,%f", &idx, &v) != 2 || idx != i) return -3;
        if (windowed) windowed[i] = v;
    }
    if (!next_line(f, line, sizeof line) || strcmp(line, "EOFLT")) return -3;
    return (int) (n / 2);
}

int usc_wire_write_fft(const char *path, float fs, uint32_t n, const float *mag, const float *db) {
    if (!path || !mag) return -1;
    FILE *f = fopen(path, "w");
    if (!f) return -2;
    fft_rows(f, fs, n, mag, db);
    int bad = ferror(f);
    return fclose(f) || bad ? -2 : (int) (n / 2);
}

int usc_wire_write_raw(const char *path, const int32_t *pcm, uint32_t n) {
    if (!path || !pcm) return -1;
    FILE *f = fopen(path, "w");
    if (!f) return -2;
    fprintf(f, "Index,Amplitude\n");
    for (uint32_t i = 0; i < n; ++i) fprintf(f, "%lu,%ld\n", (unsigned long) i, (long) pcm[i]);
    int bad = ferror(f);
    return fclose(f) || bad ? -2 : (int) n;
}

int usc_wire_write_flt(const char *path, const float *x, uint32_t n) {
    if (!path || !x) return -1;
    FILE *f = fopen(path, "w");
    if (!f) return -2;
    fprintf(f, "Index,Amplitude\n");
    for (uint32_t i = 0; i < n; ++i) fprintf(f, "%lu,%f\n", (unsigned long) i, (double) x[i]);
    int bad = ferror(f);
    return fclose(f) || bad ? -2 : (int) n;
}

/* rows of "a,b[,c]" after one header line, until EOF, a blank line or an EO* marker */
static int read_rows(const char *path, uint32_t max_rows, int cols, float *c0, float *c1, float *c2, int32_t *i1) {
    if (!path) return -1;
    FILE *f = fopen(path, "r");
    if (!f) return -2;
    char line[256];
    if (!next_line(f, line, sizeof line)) { fclose(f); return -3; }
    uint32_t rows = 0;
    while (rows < max_rows && next_line(f, line, sizeof line)) {
        if (line[0] == 0 || !strncmp(line, "EO", 2)) break;
        double a, b, c = 0.0;
        int got = cols == 3 ? sscanf(line, "%lf,%lf,%lf", &a, &b, &c) : sscanf(line, "%lf,%lf", &a, &b);
        if (got != cols) { fclose(f); return -3; }
        if (c0) c0[rows] = (float) a;
        if (c1) c1[rows] = (float) b;
        if (c2) c2[rows] = (float) c;
        if (i1) i1[rows] = (int32_t) b;
        ++rows;
    }
    fclose(f);
    return (int) rows;
}

int usc_wire_read_fft(const char *path, uint32_t max_rows, float *freq, float *mag, float *db) {
    return read_rows(path, max_rows, 3, freq, mag, db, NULL);
}
int usc_wire_read_raw(const char *path, uint32_t max_rows, int32_t *pcm) { return read_rows(path, max_rows, 2, NULL, NULL, NULL, pcm); }
int usc_wire_read_flt(const char *path, uint32_t max_rows, float *x) { return read_rows(path, max_rows, 2, NULL, x, NULL, NULL); }

int usc_wire_write_history(FILE *f, int detail, uint32_t prev_state, uint32_t state, const usc_history *hist, uint32_t num) {
    if (!f || !hist || prev_state > 3 || state > 3) return -1;
    if (!detail) {                                                 /* receiver/Src/main.c:283-288 */
        fprintf(f, "%c => %c\n", STATE_CHAR[prev_state], STATE_CHAR[state]);
        for (uint32_t i = 0; i < num; ++i) fprintf(f, "%c,%6.1f\n", (char) hist[i].rank, (double) hist[i].snr);
    } else {                                                       /* :290-300 */
        fprintf(f, "\nstate: %s => %s\n", STATE_NAME[prev_state], STATE_NAME[state]);
        fprintf(f, "r,  freq,freq_l,freq_r, t_s, t_f,      max,    max_l,    max_r, mag_mean,   snr\n");
        for (uint32_t i = 0; i < num; ++i)
            fprintf(f, "%c,%6ld,%6ld,%6ld,%4lu,%4lu, %4.2e, %4.2e, %4.2e, %4.2e,%6.1f\n", (char) hist[i].rank,
                    (long) hist[i].max_freq, (long) hist[i].max_freq_left, (long) hist[i].max_freq_right, 0ul, 0ul,
                    (double) hist[i].mag_max, (double) hist[i].mag_max_left, (double) hist[i].mag_max_right,
                    (double) hist[i].mag_mean, (double) hist[i].snr);
    }
    return ferror(f) ? -2 : (int) num;
}

static int state_from_name(const char *s) {
    for (int i = 3; i >= 0; --i)                                   /* longest names first: SYNCHRONIZED before SYNCHRONIZING is irrelevant, exact match */
        if (!strcmp(s, STATE_NAME[i])) return i;
    return -1;
}

int usc_wire_read_history(FILE *f, int detail, uint32_t *prev_state, uint32_t *state, usc_history *hist, uint32_t max_rows) {
    if (!f || !hist) return -1;
    char line[512];
    int ps = -1, st = -1;
    while (next_line(f, line, sizeof line)) {
        if (!detail) {
            char a, b;
            if (sscanf(line, "%c => %c", &a, &b) == 2) {
                for (int i = 0; i < 4; ++i) { if (STATE_CHAR[i] == a) ps = i; if (STATE_CHAR[i] == b) st = i; }
                break;
            }
        } else if (!strncmp(line, "state: ", 7)) {
            char a[32], b[32];
            if (sscanf(line + 7, "%31s => %31s", a, b) != 2) return -3;
            ps = state_from_name(a); st = state_from_name(b);
            if (!next_line(f, line, sizeof line)) return -3;       /* column header */
            break;
        }
    }
    if (ps < 0 || st < 0) return -3;
    if (prev_state) *prev_state = (uint32_t) ps;
    if (state) *state = (uint32_t) st;
    uint32_t rows = 0;
    while (rows < max_rows) {
        long pos = ftell(f);
        if (!next_line(f, line, sizeof line) || line[0] == 0) break;
        usc_history h;
        memset(&h, 0, sizeof h);
        char r;
        if (!detail) {
            float snr;
            if (sscanf(line, "%c,%f", &r, &snr) != 2 || line[1] != ',') { fseek(f, pos, SEEK_SET); break; }
            h.snr = snr;
        } else {
            long fq, fl, fr; unsigned long ts, tf; float m, ml, mr, mm, snr;
            if (sscanf(line, "%c,%ld,%ld,%ld,%lu,%lu, %e, %e, %e, %e,%f", &r, &fq, &fl, &fr, &ts, &tf, &m, &ml, &mr, &mm, &snr) != 11) {
                fseek(f, pos, SEEK_SET);
                break;
            }
            h.max_freq = (int32_t) fq; h.max_freq_left = (int32_t) fl; h.max_freq_right = (int32_t) fr;
            h.mag_max = m; h.mag_max_left = ml; h.mag_max_right = mr; h.mag_mean = mm; h.snr = snr;
        }
        h.rank = (uint32_t) (unsigned char) r;
        hist[rows++] = h;
    }
    return (int) rows;
}
