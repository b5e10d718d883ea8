/*
 * orc_mix.c - CPU ORACLE (test infrastructure, see orc.h): spatializer stereo
 * mixdown, tuner (ring feed + autocorrelation analysis) and the per-block
 * controller.process fan-out.  Citations are /root/reference paths.
 */
#define _GNU_SOURCE
#include "orc.h"
#include "../include/gdg_tables.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================== */
/* spatializer/spatializer.go                                                */
/* ======================================================================== */
#define DEFAULT_SAMPLE_RATE 96000
#define EFFECTIVE_DISTANCE 0.215
#define HALF_EFFECTIVE_DISTANCE (0.5 * EFFECTIVE_DISTANCE)
#define GROUP_DELAY 6.3e-4

typedef struct { double azimuth, distance, level; } position_t;
struct orc_spatializer {
    double **buffers; int buffer_size;
    uint32_t input_count;
    uint32_t sample_rate;            /* stays DEFAULT_SAMPLE_RATE: SetSampleRate never stores it (spatializer.go:418-431) */
    position_t *pos;
};

/* spatializer.go:436-469 Create. */
orc_spatializer *orc_spatializer_create(uint32_t n)
{
    orc_spatializer *s = (orc_spatializer *)calloc(1, sizeof(*s));
    s->input_count = n; s->sample_rate = DEFAULT_SAMPLE_RATE;
    s->pos = (position_t *)calloc(n ? n : 1, sizeof(position_t));
    for (uint32_t i = 0; i < n; i++) s->pos[i].level = 1.0;
    s->buffer_size = (int)ceil((double)DEFAULT_SAMPLE_RATE * GROUP_DELAY);
    s->buffers = (double **)calloc(n ? n : 1, sizeof(double *));
    for (uint32_t i = 0; i < n; i++) s->buffers[i] = (double *)calloc(s->buffer_size, sizeof(double));
    return s;
}
void orc_spatializer_destroy(orc_spatializer *s)
{
    if (!s) return;
    for (uint32_t i = 0; i < s->input_count; i++) free(s->buffers[i]);
    free(s->buffers); free(s->pos); free(s);
}
/* spatializer.go:340-413 setters (note the `>` range check of the reference). */
int orc_spatializer_set_azimuth(orc_spatializer *s, uint32_t ch, double az)
{
    if (ch > s->input_count || ch >= s->input_count) return -1;
    s->pos[ch].azimuth = az; return 0;
}
int orc_spatializer_set_distance(orc_spatializer *s, uint32_t ch, double d)
{
    if (ch >= s->input_count) return -1;
    if (d < 0.0 || d > 10.0) return -2;
    s->pos[ch].distance = d; return 0;
}
int orc_spatializer_set_level(orc_spatializer *s, uint32_t ch, double l)
{
    if (ch >= s->input_count) return -1;
    if (l < 0.0 || l > 1.0) return -2;
    s->pos[ch].level = l; return 0;
}
/* spatializer.go:418-431 SetSampleRate: re-creates (zeroes) the histories only. */
void orc_spatializer_set_sample_rate(orc_spatializer *s, uint32_t rate)
{
    int bs = (int)ceil((double)rate * GROUP_DELAY);
    for (uint32_t i = 0; i < s->input_count; i++) {
        free(s->buffers[i]);
        s->buffers[i] = (double *)calloc(bs > 0 ? bs : 1, sizeof(double));
    }
    s->buffer_size = bs;
}

/* spatializer.go:140-335 Process. */
void orc_spatializer_process(orc_spatializer *s, const double *const *in, uint32_t n_in,
                             const double *aux, double *const *out, uint32_t n_out, size_t n)
{
    if (n_in != s->input_count || n_out != 2) return;
    double sr = (double)s->sample_rate;
    memset(out[0], 0, sizeof(double) * n);
    memset(out[1], 0, sizeof(double) * n);
    for (uint32_t c = 0; c < n_in; c++) {
        const double *x = in[c];
        position_t p = s->pos[c];
        double az = (M_PI / 180.0) * p.azimuth;
        const double *hist = s->buffers[c]; int hs = s->buffer_size;
        double sn = sin(az), cs = cos(az);
        double xp = p.distance * sn, yp = p.distance * cs;
        double xl = fabs(xp + HALF_EFFECTIVE_DISTANCE), xr = fabs(xp - HALF_EFFECTIVE_DISTANCE);
        double yd = fabs(yp), yd2 = yd * yd;
        double dl = sqrt(xl * xl + yd2);
        double pl = 1.0 / dl; if (pl > 1.0) pl = 1.0;
        double fl = p.level * pl;
        double dr = sqrt(xr * xr + yd2);
        double pr = 1.0 / dr; if (pr > 1.0) pr = 1.0;
        double fr = p.level * pr;
        double delay_time = (GROUP_DELAY / EFFECTIVE_DISTANCE) * (dl - dr);
        double ds = fabs(delay_time) * sr;
        double early = floor(ds), late = ceil(ds);
        int ei = (int)early, li = (int)late;
        if (ei >= hs) ei = hs - 1;
        if (li >= hs) li = hs - 1;
        for (size_t j = 0; j < n; j++) {
            if (delay_time == 0.0) {
                out[0][j] += fl * x[j];
                out[1][j] += fr * x[j];
            } else {
                long ie = (long)j - ei, il = (long)j - li;
                double se = ie >= 0 ? x[ie] : hist[hs + ie];
                double sl = il >= 0 ? x[il] : hist[hs + il];
                double we = 1.0 - (ds - early), wl = 1.0 - (late - ds);
                double dsmp = (we * se) + (wl * sl);
                if (delay_time > 0.0) { out[0][j] += fl * dsmp; out[1][j] += fr * x[j]; }
                else { out[0][j] += fl * x[j]; out[1][j] += fr * dsmp; }
            }
        }
    }
    if (aux)
        for (size_t j = 0; j < n; j++) { out[0][j] += aux[j]; out[1][j] += aux[j]; }
    for (uint32_t c = 0; c < n_in; c++) {
        double *hist = s->buffers[c]; long hs = s->buffer_size;
        long boundary = hs - (long)n;
        if (boundary >= 0) {
            memmove(hist, hist + n, sizeof(double) * boundary);
            memcpy(hist + boundary, in[c], sizeof(double) * n);
        } else memcpy(hist, in[c] + (-boundary), sizeof(double) * hs);
    }
}

/* ======================================================================== */
/* tuner/tuner.go                                                            */
/* ======================================================================== */
#define TUNER_NUM_SAMPLES 96000
static const char *g_note_names[GDG_TUNER_NOTE_COUNT] = GDG_TUNER_NOTE_NAMES_INIT;
static const double g_note_freqs[GDG_TUNER_NOTE_COUNT] = GDG_TUNER_NOTE_FREQS_INIT;

struct orc_tuner {
    orc_circular *buffer; uint32_t sample_rate;
    double *corr; double *fft; size_t fft_size;
};
orc_tuner *orc_tuner_create(void)
{
    orc_tuner *t = (orc_tuner *)calloc(1, sizeof(*t));
    t->buffer = orc_circular_create(TUNER_NUM_SAMPLES);
    return t;
}
void orc_tuner_destroy(orc_tuner *t)
{
    if (!t) return;
    orc_circular_destroy(t->buffer); free(t->corr); free(t->fft); free(t);
}
/* tuner.go:582-587 Process. */
void orc_tuner_process(orc_tuner *t, const double *samples, size_t n, uint32_t rate)
{
    orc_circular_enqueue(t->buffer, samples, n);
    t->sample_rate = rate;
}
/* tuner.go:379-577 Analyze: FFT autocorrelation, peak in the lag window of the
 * note range, parabolic refinement, nearest note. */
int orc_tuner_analyze(orc_tuner *t, orc_tuner_result *res)
{
    size_t n = orc_circular_length(t->buffer);
    uint64_t two_n = 2 * (uint64_t)n, fft_size;
    orc_fft_next_power_of_two(two_n, &fft_size, NULL);
    if (t->fft_size != fft_size) {
        free(t->corr); free(t->fft);
        t->corr = (double *)calloc(fft_size, sizeof(double));
        t->fft = (double *)calloc(2 * fft_size, sizeof(double));
        t->fft_size = fft_size;
    }
    double *corr = t->corr;
    if (orc_circular_retrieve(t->buffer, corr, n)) return -1;
    memset(corr + n, 0, sizeof(double) * (fft_size - n));
    if (orc_fft_real_fourier(corr, t->fft, fft_size, ORC_SCALING_DEFAULT)) return -2;
    for (size_t i = 0; i < fft_size; i++) {
        double re = t->fft[2 * i], im = t->fft[2 * i + 1];
        /* elem * conj(elem): real re*re + im*im, imag re*(-im) + im*re */
        t->fft[2 * i] = re * re + im * im;
        t->fft[2 * i + 1] = re * (-im) + im * re;
    }
    if (orc_fft_real_inverse_fourier(t->fft, corr, fft_size, ORC_SCALING_DEFAULT)) return -3;
    double low_freq = g_note_freqs[0], high_freq = g_note_freqs[GDG_TUNER_NOTE_COUNT - 1];
    double sr = (double)t->sample_rate;
    long low_idx = (long)((sr / high_freq) + 0.5);
    if (low_idx < 0 || (uint64_t)low_idx >= two_n) low_idx = 0;
    long high_idx = (long)((sr / low_freq) + 0.5);
    if (high_idx < 0 || (uint64_t)high_idx >= two_n) high_idx = (long)(two_n - 1);
    double max_val = -INFINITY; long max_idx = -1;
    for (long i = low_idx; i < high_idx; i++)
        if (corr[i] > max_val) { max_val = corr[i]; max_idx = i - low_idx; }
    long idx = low_idx + max_idx;
    long up = idx + 1; if (up > (long)n) up = (long)n;
    long down = idx - 1; if (down < 0) down = 0;
    double left = corr[down], right = corr[up];
    double shift = (0.5 * (right - left)) / ((2.0 * max_val) - (right + left));
    if (shift < -0.5) shift = -0.5; else if (shift > 0.5) shift = 0.5;
    double freq = sr / ((double)idx + shift);
    const char *note = "Unknown";
    double cents = INFINITY, cents_abs = INFINITY;
    for (int k = 0; k < GDG_TUNER_NOTE_COUNT; k++) {
        double dc = 1200.0 * log2(freq / g_note_freqs[k]);
        if (fabs(dc) < cents_abs) { note = g_note_names[k]; cents = dc; cents_abs = fabs(dc); }
    }
    res->cents = (isinf(cents) || isnan(cents)) ? 0 : (int8_t)cents;
    res->frequency = freq;
    res->note = note;
    return 0;
}

/* ======================================================================== */
/* level/level.go: level meter (peak programme meter with 2 s peak hold)      */
/* ======================================================================== */
typedef struct { double cur, peak; uint64_t counter; int enabled; } orc_meter_channel;
struct orc_meter { uint32_t count; int enabled; orc_meter_channel *ch; };

orc_meter *orc_meter_create(uint32_t channels)          /* level.go:318-399: all channels disabled, state zero */
{
    orc_meter *m = (orc_meter *)calloc(1, sizeof(*m));
    m->count = channels;
    m->ch = (orc_meter_channel *)calloc(channels ? channels : 1, sizeof(orc_meter_channel));
    return m;
}
void orc_meter_destroy(orc_meter *m) { if (m) { free(m->ch); free(m); } }
uint32_t orc_meter_channel_count(const orc_meter *m) { return m->count; }
int orc_meter_enabled(const orc_meter *m) { return m->enabled; }
void orc_meter_set_enabled(orc_meter *m, int value)      /* level.go:333-356 -> 214-237 */
{
    value = value != 0;
    if (value == m->enabled) return;
    for (uint32_t c = 0; c < m->count; c++) {
        orc_meter_channel *k = &m->ch[c];
        if (value != k->enabled) {
            if (!value) { k->cur = 0.0; k->peak = 0.0; k->counter = 0; }
            k->enabled = value;
        }
    }
    m->enabled = value;
}
/* level.go:147-208 */
static void meter_channel_process(orc_meter_channel *k, const double *buf, size_t n, uint32_t rate)
{
    if (!k->enabled) return;
    double cur = k->cur, peak = k->peak;
    uint64_t counter = k->counter;
    const double sr = (double)rate;
    const uint64_t hold = (uint64_t)(2.0 * sr);
    const double f = pow(10.0, -1.0 / (1.7 * sr));
    for (size_t i = 0; i < n; i++) {
        cur *= f;
        if (counter > hold) peak *= f; else counter++;
        const double a = fabs(buf[i]);
        if (a > cur) cur = a;
        if (a >= peak) { peak = a; counter = 0; }
    }
    k->cur = cur; k->peak = peak; k->counter = counter;
}
/* level.go:302-328: -1 when the number of buffers differs from the channel count; lens[c] samples in bufs[c] (NULL: none) */
int orc_meter_process(orc_meter *m, const double *const *bufs, const size_t *lens, uint32_t n_bufs, uint32_t rate)
{
    if (n_bufs != m->count) return -1;
    for (uint32_t c = 0; c < n_bufs; c++)
        if (bufs[c]) meter_channel_process(&m->ch[c], bufs[c], lens[c], rate);
    return 0;
}
static int32_t meter_db(double v)                       /* level.go:70-73, 95-137 */
{
    double l = 20.0 * log10(v);
    if (isnan(l) || l < -200.0) l = -200.0;
    return (int32_t)round(l);
}
int orc_meter_analyze(const orc_meter *m, uint32_t channel, int32_t *level, int32_t *peak)   /* level.go:242-255 */
{
    if (channel >= m->count) return -1;
    *level = meter_db(m->ch[channel].cur);
    *peak = meter_db(m->ch[channel].peak);
    return 0;
}
int orc_meter_state(const orc_meter *m, uint32_t channel, double *cur, double *peak, uint64_t *counter)
{
    if (channel >= m->count) return -1;
    *cur = m->ch[channel].cur; *peak = m->ch[channel].peak; *counter = m->ch[channel].counter;
    return 0;
}

/* ======================================================================== */
/* controller/controller.go:2648-2783 process (metronome = zeros, no meters)  */
/* ======================================================================== */
void orc_controller_process(orc_chain *const *chains, uint32_t n_ch, orc_tuner *tuner, int tuner_channel,
                            orc_spatializer *spat, const double *const *in, double *const *out,
                            uint32_t n_out, size_t n, uint32_t rate)
{
    if (tuner && tuner_channel >= 0 && (uint32_t)tuner_channel < n_ch)
        orc_tuner_process(tuner, in[tuner_channel], n, rate);
    if (n_out >= n_ch)
        for (uint32_t c = 0; c < n_ch; c++)
            orc_chain_process(chains[c], in[c], out[c], n, rate);
    if (spat && n_out >= n_ch + 2)
        orc_spatializer_process(spat, (const double *const *)out, n_ch, NULL, out + n_ch, 2, n);
}
