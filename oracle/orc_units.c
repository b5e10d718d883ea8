/*
 * orc_units.c - CPU ORACLE (test infrastructure, see orc.h): the 21 effects
 * units (effects/<unit>.go Process bodies) and the signal chain driver
 * (signal/signal.go).  Sequential per-sample loops exactly as the reference
 * runs them; citations are /root/reference paths.
 */
#define _GNU_SOURCE
#include "orc.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

orc_filter *orc__irbank_create_filter(const orc_irbank *b, const char *name, uint32_t rate);
orc_filter *orc__filter_add(const orc_filter *a, const orc_filter *b);

static __thread char g_err[512];
const char *orc_last_error(void) { return g_err; }
static int fail(const char *fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
    return -1;
}

/* ---- parameters: effects/effects.go:69-78, 96-384 ------------------------ */
enum { PT_INVALID = 0, PT_DISCRETE = 1, PT_NUMERIC = 2 };
typedef struct {
    char *name; int type; char *phys;
    int32_t min, max, num; int didx; int ndisc; char **disc;
} param_t;

enum {
#define GDG_UNIT(code, ident, name) UNIT_##ident = code,
#define GDG_NUM(...)
#define GDG_DISC(...)
#include "../include/gdg_unit_params.def"
#undef GDG_UNIT
#undef GDG_NUM
#undef GDG_DISC
    UNIT_COUNT
};
#define NUM_FILTERS 8
#define STRING_NONE "- NONE -"

typedef struct { double *buf; int n; } dbuf;

typedef struct {
    double *buffer; int n; int ptr; double feedback;
} allpass_t;

typedef struct unit {
    int type;
    param_t *params; int nparams;
    /* generic state (each unit uses the fields its Go struct has) */
    double envelope, coupling, previous_phase, phase, previous_polarity;
    uint32_t octave_register, on_hold_since, in_state_since;
    int gate_open, attenuated;
    double hcv[NUM_FILTERS], lcv[NUM_FILTERS];
    int n_hcv, n_lcv;                         /* len() of the Go slices (0 = nil) */
    dbuf delay_buf;                           /* "buffer" history of delay-type units */
    int delay_buf_valid;
    orc_prng prng; int prng_valid;
    /* oversampled units */
    orc_osd *os2, *os4; dbuf os_in, os_out;
    /* reverb */
    allpass_t allpass[3]; dbuf rv_delay; uint32_t rv_taps[4]; double rv_coeffs[4];
    uint32_t rv_rate; dbuf rv_front, rv_back, rv_line;
    /* cabinet */
    dbuf cab_buf;
    /* poweramp */
    uint32_t pa_rate; orc_irbank *bank; orc_filter *pa_filter;
} unit_t;

static void dbuf_resize_zero(dbuf *b, int n)
{
    free(b->buf); b->buf = (double *)calloc(n > 0 ? n : 1, sizeof(double)); b->n = n;
}

static void add_num(unit_t *u, const char *name, const char *phys, int32_t mn, int32_t mx, int32_t def)
{
    u->params = (param_t *)realloc(u->params, sizeof(param_t) * (u->nparams + 1));
    param_t *p = &u->params[u->nparams++];
    memset(p, 0, sizeof *p);
    p->name = strdup(name); p->type = PT_NUMERIC; p->phys = strdup(phys);
    p->min = mn; p->max = mx; p->num = def; p->didx = -1;
}
static void add_disc_list(unit_t *u, const char *name, int def_idx, char **vals, int n)
{
    u->params = (param_t *)realloc(u->params, sizeof(param_t) * (u->nparams + 1));
    param_t *p = &u->params[u->nparams++];
    memset(p, 0, sizeof *p);
    p->name = strdup(name); p->type = PT_DISCRETE; p->phys = strdup("");
    p->min = p->max = p->num = -1; p->didx = def_idx; p->ndisc = n; p->disc = vals;
}
static void add_disc(unit_t *u, const char *name, int def_idx, const char *joined)
{
    int n = 1;
    for (const char *c = joined; *c; c++) if (*c == '|') n++;
    char **vals = (char **)malloc(sizeof(char *) * n);
    const char *s = joined;
    for (int i = 0; i < n; i++) {
        const char *e = strchr(s, '|');
        size_t len = e ? (size_t)(e - s) : strlen(s);
        vals[i] = strndup(s, len);
        s += len + 1;
    }
    add_disc_list(u, name, def_idx, vals, n);
}

/* effects.go:254-317 getNumericValue / 177-227 getDiscreteValue (last match wins). */
static int32_t pnum(const unit_t *u, const char *name)
{
    int idx = -1;
    for (int i = 0; i < u->nparams; i++) if (!strcmp(u->params[i].name, name)) idx = i;
    if (idx < 0 || u->params[idx].type != PT_NUMERIC) return 0;
    return u->params[idx].num;
}
static const char *pdisc(const unit_t *u, const char *name)
{
    int idx = -1;
    for (int i = 0; i < u->nparams; i++) if (!strcmp(u->params[i].name, name)) idx = i;
    if (idx < 0 || u->params[idx].type != PT_DISCRETE) return "";
    return u->params[idx].disc[u->params[idx].didx];
}

/* effects.go:389-394 decibelsToFactor, 399-402 factorToDecibels, 421-438 signFloat. */
static double db2fac(int32_t db) { return pow(10.0, 0.05 * (double)db); }
static double fac2db(double f) { return 20.0 * log10(f); }
static double sign_float(double x) { return x < 0.0 ? -1.0 : (x > 0.0 ? 1.0 : 0.0); }
static double clamp1(double x) { return x < -1.0 ? -1.0 : (x > 1.0 ? 1.0 : x); }

#define MATH_TWO_PI (2.0 * M_PI)

/* shared by every follower unit: compressor.go:44-58 etc.; follow: 0 envelope, 1 level, 2 other */
static int follow_mode(const char *s) { return !strcmp(s, "envelope") ? 0 : (!strcmp(s, "level") ? 1 : 2); }
static inline double follow_step(int mode, double env, double abs_s, double d_inv, double d)
{
    if (mode == 0) { env *= d_inv; if (abs_s > env) env = abs_s; return env; }
    if (mode == 1) { double diff = abs_s - env; return env + diff * d; }
    return 1.0;
}

/* history shift shared by the delay-type units (e.g. chorus.go:119-129). */
static void history_push(double *buffer, int size, const double *in, int n)
{
    int boundary = size - n;
    if (boundary >= 0) {
        memmove(buffer, buffer + n, sizeof(double) * boundary);
        memcpy(buffer + boundary, in, sizeof(double) * n);
    } else {
        memcpy(buffer, in + (-boundary), sizeof(double) * size);
    }
}
/* x[idx] for idx in [-hsize, n).  DELIBERATE DIFFERENCE: flanger / phaser at 22050, 44100 and 88200 Hz with depth 100
 * ask for idx = -hsize - 1 (delay 0.002 fs = 44.1 / 88.2 / 176.4 samples against a history of round(0.002 fs)); the
 * reference panics there (flanger.go:80-84: buffer[-1]).  The oracle and the product read the OLDEST history sample
 * instead (index clamped to -hsize) and never touch memory outside the history. */
static inline double delayed(const double *in, const double *hist, int hsize, int idx)
{
    if (idx >= 0) return in[idx];
    if (idx < -hsize) idx = -hsize;
    return hsize > 0 ? hist[hsize + idx] : 0.0;
}

/* ---- signalgenerator.go:20-153 ------------------------------------------- */
static void signalgenerator_process(unit_t *u, const double *in, double *out, int n, uint32_t rate)
{
    double fac_in = (0.01 * (double)pnum(u, "input_amplitude")) * db2fac(pnum(u, "input_gain"));
    const char *type = pdisc(u, "signal_type");
    double fac_sig_gain = db2fac(pnum(u, "signal_gain"));
    double fac_sig = (0.01 * (double)pnum(u, "signal_amplitude")) * fac_sig_gain;
    double phase = u->phase;
    double inc = MATH_TWO_PI * ((double)pnum(u, "signal_frequency") / (double)rate);
    double two_over_pi = 2.0 / M_PI;
    int periodic = 1;
    if (!strcmp(type, "sine")) {
        for (int i = 0; i < n; i++) {
            double cp = fmod(phase + ((double)i * inc), MATH_TWO_PI);
            out[i] = (fac_in * in[i]) + (fac_sig * sin(cp));
        }
    } else if (!strcmp(type, "triangle")) {
        for (int i = 0; i < n; i++) {
            double cp = fmod(phase + ((double)i * inc), MATH_TWO_PI);
            double s = cp < M_PI ? (two_over_pi * cp) - 1.0 : 3.0 - (two_over_pi * cp);
            out[i] = (fac_in * in[i]) + (fac_sig * s);
        }
    } else if (!strcmp(type, "square")) {
        for (int i = 0; i < n; i++) {
            double cp = fmod(phase + ((double)i * inc), MATH_TWO_PI);
            out[i] = (fac_in * in[i]) + (fac_sig * sign_float(M_PI - cp));
        }
    } else if (!strcmp(type, "sawtooth")) {
        for (int i = 0; i < n; i++) {
            double cp = fmod(phase + ((double)i * inc), MATH_TWO_PI);
            double s = cp / M_PI;
            if (cp > M_PI) s -= 2.0;
            out[i] = (fac_in * in[i]) + (fac_sig * s);
        }
    } else if (!strcmp(type, "noise")) {
        periodic = 0;
        if (!u->prng_valid) { orc_prng_seed(&u->prng, 1337); u->prng_valid = 1; }
        for (int i = 0; i < n; i++) {
            double r = orc_prng_next_float(&u->prng);
            out[i] = (fac_in * in[i]) + (fac_sig * (1.0 - (2.0 * r)));
        }
    } else periodic = 0;
    if (periodic) { phase += (double)n * inc; phase = fmod(phase, MATH_TWO_PI); }
    u->phase = phase;
}

/* ---- noisegate.go:19-96 ---------------------------------------------------- */
static void noisegate_process(unit_t *u, const double *in, double *out, int n, uint32_t rate)
{
    int32_t lo = pnum(u, "threshold_open"), lc = pnum(u, "threshold_close"), ht = pnum(u, "hold_time");
    double fac_open = db2fac(lo), fac_close = db2fac(lc);
    if (lo < lc) {
        memcpy(out, in, sizeof(double) * n);
        u->gate_open = 1; u->on_hold_since = 0;
        return;
    }
    uint32_t hold_samples = (uint32_t)floor(((0.001 * (double)ht) * (double)rate) + 0.5);
    int open = u->gate_open; uint32_t since = u->on_hold_since;
    for (int i = 0; i < n; i++) {
        double amp = fabs(in[i]);
        if (amp > fac_open) open = 1;
        if (amp > fac_close) since = 0;
        if (since >= hold_samples) open = 0;
        out[i] = (open ? 1.0 : 0.0) * in[i];
        if (since < UINT32_MAX) since++;
    }
    u->gate_open = open; u->on_hold_since = since;
}

/* ---- bandpass.go:20-98 ----------------------------------------------------- */
static void bandpass_process(unit_t *u, const double *in, double *out, int n, uint32_t rate)
{
    int half = (int)(strtoul(pdisc(u, "filter_order"), NULL, 10) >> 1);
    int32_t fa = pnum(u, "frequency_1"), fb = pnum(u, "frequency_2");
    if (fa > fb) { int32_t t = fa; fa = fb; fb = t; }
    if (u->n_hcv != half) { memset(u->hcv, 0, sizeof u->hcv); u->n_hcv = half; }
    if (u->n_lcv != half) { memset(u->lcv, 0, sizeof u->lcv); u->n_lcv = half; }
    double m2pi = -MATH_TWO_PI / (double)rate;
    double hp_inv = 1.0 - exp(m2pi * (double)fa);
    double lp_inv = 1.0 - exp(m2pi * (double)fb);
    for (int i = 0; i < n; i++) {
        double pre = in[i];
        for (int j = 0; j < half; j++) {
            double hcv = u->hcv[j];
            double diff = pre - hcv;
            hcv += diff * hp_inv;
            u->hcv[j] = hcv;
            double lcv = u->lcv[j];
            diff -= lcv;
            double iv = lcv;
            lcv += diff * lp_inv;
            u->lcv[j] = lcv;
            pre = clamp1(iv);
        }
        out[i] = pre;
    }
}

/* ---- autowah.go:20-130 ------------------------------------------------------ */
static void autowah_process(unit_t *u, const double *in, double *out, int n, uint32_t rate)
{
    int mode = follow_mode(pdisc(u, "follow"));
    int32_t la = pnum(u, "level_1"), lb = pnum(u, "level_2"), fa = pnum(u, "frequency_1"), fb = pnum(u, "frequency_2");
    if (la > lb) { int32_t t = la; la = lb; lb = t; t = fa; fa = fb; fb = t; }
    double laf = la, lbf = lb, faf = fa, fbf = fb;
    double slope = (fbf - faf) / (lbf - laf);
    double sr = (double)rate;
    double d_inv = exp(-20.0 / sr), d = 1.0 - d_inv;
    double env = u->envelope;
    double gain_comp = pow(2.0, NUM_FILTERS);
    for (int i = 0; i < n; i++) {
        double s = in[i];
        env = follow_step(mode, env, fabs(s), d_inv, d);
        double level = fac2db(env);
        double freq;
        if (level <= laf) freq = faf;
        else if (level >= lbf) freq = fbf;
        else freq = faf + (slope * (level - laf));
        double k = 1.0 - exp(-freq / sr);                 /* no 2*pi here (autowah.go:95-96) */
        double lcv = s;
        for (int j = 0; j < NUM_FILTERS; j++) {
            double hcv = u->hcv[j];
            double diff = lcv - hcv;
            hcv += diff * k;
            u->hcv[j] = hcv;
            lcv = u->lcv[j];
            diff -= lcv;
            lcv += diff * k;
            u->lcv[j] = lcv;
        }
        out[i] = clamp1(gain_comp * lcv);
    }
    u->envelope = env;
}

/* ---- autoyoy.go:19-156 ------------------------------------------------------ */
static void autoyoy_process(unit_t *u, const double *in, double *out, int n, uint32_t rate)
{
    int mode = follow_mode(pdisc(u, "follow"));
    int32_t la = pnum(u, "level_1"), lb = pnum(u, "level_2");
    double depth_a = 0.0, depth_b = 0.01 * (double)pnum(u, "depth");
    if (la > lb) { int32_t t = la; la = lb; lb = t; double d = depth_a; depth_a = depth_b; depth_b = d; }
    double laf = la, lbf = lb;
    double slope = (depth_b - depth_a) / (lbf - laf);
    double sr = (double)rate, sr_inv = 1.0 / sr;
    double d_inv = exp(-20.0 * sr_inv), d = 1.0 - d_inv;
    int max_delay = (int)floor((0.01 * sr) + 0.5);
    if (!u->delay_buf_valid || u->delay_buf.n != max_delay) { dbuf_resize_zero(&u->delay_buf, max_delay); u->delay_buf_valid = 1; }
    double *hist = u->delay_buf.buf; int hs = max_delay;
    double env = u->envelope;
    for (int i = 0; i < n; i++) {
        double s = in[i];
        env = follow_step(mode, env, fabs(s), d_inv, d);
        double level = fac2db(env);
        double dfac;
        if (level <= laf) dfac = depth_a;
        else if (level >= lbf) dfac = depth_b;
        else dfac = depth_a + (slope * (level - laf));
        double ds = (0.01 * dfac) * sr;
        double early = floor(ds), late = ceil(ds);
        double se = delayed(in, hist, hs, i - (int)early);
        double sl = delayed(in, hist, hs, i - (int)late);
        double we = 1.0 - (ds - early), wl = 1.0 - (late - ds);
        double dsmp = (we * se) + (wl * sl);
        out[i] = (0.5 * s) + (0.5 * dsmp);
    }
    u->envelope = env;
    history_push(hist, hs, in, n);
}

/* ---- compressor.go:18-84 ---------------------------------------------------- */
static void compressor_process(unit_t *u, const double *in, double *out, int n, uint32_t rate)
{
    int mode = follow_mode(pdisc(u, "follow"));
    double limit = db2fac(pnum(u, "gain_limit")), target = db2fac(pnum(u, "target_level"));
    double d_inv = exp(-20.0 / (double)rate), d = 1.0 - d_inv;
    double env = u->envelope;
    for (int i = 0; i < n; i++) {
        double s = in[i];
        env = follow_step(mode, env, fabs(s), d_inv, d);
        double gain = target / env;
        if (gain > limit) gain = limit;
        out[i] = clamp1(gain * s);
    }
    u->envelope = env;
}

/* ---- octaver.go:21-139 ------------------------------------------------------ */
static void octaver_process(unit_t *u, const double *in, double *out, int n, uint32_t rate)
{
    int mode = follow_mode(pdisc(u, "follow"));
    double f_up = db2fac(pnum(u, "level_octave_up")), f_clean = db2fac(pnum(u, "level_clean"));
    double f_dist = db2fac(pnum(u, "level_dist")), f_d1 = db2fac(pnum(u, "level_octave_down_first"));
    double f_d2 = db2fac(pnum(u, "level_octave_down_second")), f_hyst = db2fac(pnum(u, "level_hysteresis"));
    double pol = u->previous_polarity; uint32_t reg = u->octave_register;
    double env = u->envelope, ccv = u->coupling;
    double d_inv = exp(-20.0 / (double)rate), d = 1.0 - d_inv;
    for (int i = 0; i < n; i++) {
        double s = in[i], sa = fabs(s);
        env = follow_step(mode, env, sa, d_inv, d);
        double square = s * s;
        double sign = sign_float(s);
        double hyst = env * f_hyst;
        if ((sign != 0.0) && (sign != pol) && (sa > hyst)) { reg = (reg + 1) & 0x7; pol = sign; }
        double first = (reg & 0x2) ? -1.0 : 1.0;
        double second = (reg & 0x4) ? -1.0 : 1.0;
        double pre = f_clean * s;
        if (env > 0.0001) pre += f_up * (square / env);
        pre += f_dist * (sign * env);
        pre += f_d1 * (first * env);
        pre += f_d2 * (second * env);
        ccv += (pre - ccv) * d;
        pre -= ccv;
        out[i] = clamp1(pre);
    }
    u->previous_polarity = pol; u->octave_register = reg; u->envelope = env; u->coupling = ccv;
}

/* ---- oversampling wrapper shared by excess/fuzz/overdrive/distortion
 *      (e.g. overdrive.go:83-144) ------------------------------------------- */
typedef void (*os_fn)(unit_t *, const double *, double *, int, uint32_t);
static void oversampled_process(unit_t *u, os_fn fn, const double *in, double *out, int n, uint32_t rate)
{
    const char *os = pdisc(u, "oversampling");
    int factor = 1;
    if (!strcmp(os, "2")) factor = 2; else if (!strcmp(os, "4")) factor = 4;
    if (factor > 1) {
        int ns = factor * n;
        if (u->os_in.n != ns) dbuf_resize_zero(&u->os_in, ns);
        if (u->os_out.n != ns) dbuf_resize_zero(&u->os_out, ns);
        orc_osd *o = factor == 4 ? u->os4 : u->os2;
        orc_osd_oversample(o, in, n, u->os_in.buf, ns);             /* errors ignored like the reference */
        fn(u, u->os_in.buf, u->os_out.buf, ns, (uint32_t)factor * rate);
        orc_osd_decimate(o, u->os_out.buf, ns, out, n);
    } else fn(u, in, out, n, rate);
}

/* excess.go:22-66 */
static void excess_os(unit_t *u, const double *in, double *out, int n, uint32_t rate)
{
    (void)rate;
    double g = db2fac(pnum(u, "gain")), lvl = db2fac(pnum(u, "level"));
    for (int i = 0; i < n; i++) {
        double pre = g * in[i];
        double ap = fabs(pre);
        int exceeded = ap > 1.0, negative = pre < 0.0;
        int32_t section = (int32_t)(0.5 * floor(ap + 1.0));
        int odd = (section % 2) != 0;
        int inverted = odd != (exceeded && negative);
        double ex = fmod(ap + 1.0, 2.0);
        if (exceeded) pre = inverted ? 1.0 - ex : ex - 1.0;
        out[i] = lvl * pre;
    }
}

/* fuzz.go:24-108 */
static void fuzz_os(unit_t *u, const double *in, double *out, int n, uint32_t rate)
{
    int mode = follow_mode(pdisc(u, "follow"));
    double bias = 0.01 * (double)pnum(u, "bias");
    double g = db2fac(pnum(u, "boost") + pnum(u, "gain"));
    double ff = 0.01 * (double)pnum(u, "fuzz"), ffi = 1.0 - ff;
    double lvl = db2fac(pnum(u, "level"));
    double env = u->envelope, ccv = u->coupling;
    double d_inv = exp(-20.0 / (double)rate), d = 1.0 - d_inv;
    for (int i = 0; i < n; i++) {
        double s = in[i];
        env = follow_step(mode, env, fabs(s), d_inv, d);
        double pre = clamp1(g * (s - (bias * env)));
        pre = (ff * pre) + (ffi * s);
        double diff = pre - ccv;
        ccv += diff * d;
        pre -= ccv;
        out[i] = lvl * clamp1(pre);
    }
    u->envelope = env; u->coupling = ccv;
}

/* overdrive.go:28-78 */
static void overdrive_os(unit_t *u, const double *in, double *out, int n, uint32_t rate)
{
    (void)rate;
    double g = db2fac(pnum(u, "boost") + pnum(u, "gain"));
    double drive = 0.01 * (double)pnum(u, "drive"), clean = 1.0 - drive;
    double lvl = db2fac(pnum(u, "level"));
    const char *valve = pdisc(u, "valve");
    int vt = !strcmp(valve, "ECC82 (12AU7)") ? 0 : (!strcmp(valve, "ECC83 (12AX7)") ? 1 : -1);
    for (int i = 0; i < n; i++) {
        double s = in[i], arg = g * s, dist = 0.0;
        if (vt == 0) dist = (2.0 / M_PI) * atan((0.25 * M_PI) * arg);
        else if (vt == 1) { double x = exp(-arg); dist = (2.0 / (1.0 + x)) - 1.0; }
        out[i] = lvl * ((drive * dist) + (clean * s));
    }
}

/* distortion.go:21-49 */
static void distortion_os(unit_t *u, const double *in, double *out, int n, uint32_t rate)
{
    (void)rate;
    double g = db2fac(pnum(u, "boost") + pnum(u, "gain")), lvl = db2fac(pnum(u, "level"));
    for (int i = 0; i < n; i++) out[i] = lvl * clamp1(g * in[i]);
}

/* ---- tonestack.go:19-100 ---------------------------------------------------- */
static void tonestack_process(unit_t *u, const double *in, double *out, int n, uint32_t rate)
{
    static const double freqs[5] = {20.0, 300.0, 3000.0, 6000.0, 20000.0};
    static const char *names[4] = {"low", "middle", "presence", "high"};
    double facs[4];
    for (int j = 0; j < 4; j++) facs[j] = db2fac(pnum(u, names[j]));
    if (u->n_hcv != 4) { memset(u->hcv, 0, sizeof u->hcv); u->n_hcv = 4; }
    if (u->n_lcv != 4) { memset(u->lcv, 0, sizeof u->lcv); u->n_lcv = 4; }
    double m2pi = -MATH_TWO_PI / (double)rate;
    for (int i = 0; i < n; i++) {
        double s = in[i], sum = 0.0;
        for (int j = 0; j < 4; j++) {
            double hcv = u->hcv[j], lcv = u->lcv[j];
            double hp_inv = 1.0 - exp(m2pi * freqs[j]);           /* recomputed per sample, tonestack.go:69-76 */
            double lp_inv = 1.0 - exp(m2pi * freqs[j + 1]);
            double diff = s - hcv;
            hcv += diff * hp_inv;
            diff -= lcv;
            double pre = lcv;
            lcv += diff * lp_inv;
            u->hcv[j] = hcv; u->lcv[j] = lcv;
            sum += facs[j] * pre;
        }
        out[i] = clamp1(sum);
    }
}

/* ---- chorus.go:19-130 -------------------------------------------------------- */
static void chorus_process(unit_t *u, const double *in, double *out, int n, uint32_t rate)
{
    double depth = 0.1 * (double)pnum(u, "depth");
    if (depth < 0.0) depth = 0.0; else if (depth > 10.0) depth = 10.0;
    double ang = (0.001 * M_PI) * (double)pnum(u, "speed");
    double sr = (double)rate;
    int max_delay = (int)floor((0.05 * sr) + 0.5);
    double prev = u->previous_phase;
    if (!u->delay_buf_valid || u->delay_buf.n != max_delay) { dbuf_resize_zero(&u->delay_buf, max_delay); u->delay_buf_valid = 1; }
    double *hist = u->delay_buf.buf; int hs = max_delay;
    for (int i = 0; i < n; i++) {
        double t = (double)i / sr;
        double zero_phase = fmod(prev + (ang * t), MATH_TWO_PI);
        double eff = 0.0;
        for (int j = 0; j < 5; j++) {
            double phase = fmod(zero_phase + ((0.4 * M_PI) * (double)j), MATH_TWO_PI);
            double offset = depth * sin(phase);
            double ds = (0.001 * (40.0 + offset)) * sr;
            double early = floor(ds), late = ceil(ds);
            double se = delayed(in, hist, hs, i - (int)early);
            double sl = delayed(in, hist, hs, i - (int)late);
            double we = 1.0 - (ds - early), wl = 1.0 - (late - ds);
            eff += 0.2 * ((we * se) + (wl * sl));
        }
        out[i] = (0.5 * in[i]) + (0.5 * eff);
    }
    /* phase advances by the HISTORY length, not the block length (chorus.go:114-117) */
    u->previous_phase = fmod(prev + (ang * ((double)hs / sr)), MATH_TWO_PI);
    history_push(hist, hs, in, n);
}

/* ---- flanger.go:19-117 / phaser.go:19-125 ------------------------------------ */
static void flanger_phaser_process(unit_t *u, const double *in, double *out, int n, uint32_t rate, int is_phaser)
{
    double depth = 0.01 * (double)pnum(u, "depth");
    if (depth < 0.0) depth = 0.0; else if (depth > 1.0) depth = 1.0;
    double ang = (0.02 * M_PI) * (double)pnum(u, "speed");
    double wa = 0.5, wb = 0.5;
    if (is_phaser) {
        double pr = (M_PI / 180.0) * (double)pnum(u, "phase");
        wb = 0.5 * sin(pr);
        wa = 1.0 - fabs(wb);
    }
    double sr = (double)rate, sr_inv = 1.0 / sr;
    int max_delay = (int)floor((0.002 * sr) + 0.5);
    double prev = u->previous_phase;
    if (!u->delay_buf_valid || u->delay_buf.n != max_delay) { dbuf_resize_zero(&u->delay_buf, max_delay); u->delay_buf_valid = 1; }
    double *hist = u->delay_buf.buf; int hs = max_delay;
    for (int i = 0; i < n; i++) {
        double t = (double)i * sr_inv;
        double phase = fmod(prev + (ang * t), MATH_TWO_PI);
        double offset = depth * sin(phase);
        double ds = (0.001 * (depth + offset)) * sr;
        double early = floor(ds), late = ceil(ds);
        double se = delayed(in, hist, hs, i - (int)early);
        double sl = delayed(in, hist, hs, i - (int)late);
        double we = 1.0 - (ds - early), wl = 1.0 - (late - ds);
        double dsmp = (we * se) + (wl * sl);
        out[i] = (wa * in[i]) + (wb * dsmp);
    }
    u->previous_phase = fmod(prev + (ang * ((double)hs * sr_inv)), MATH_TWO_PI);
    history_push(hist, hs, in, n);
}

/* ---- tremolo.go:15-65 -------------------------------------------------------- */
static void tremolo_process(unit_t *u, const double *in, double *out, int n, uint32_t rate)
{
    double period_f = (double)rate / (0.1 * (double)pnum(u, "frequency"));
    uint32_t period = (uint32_t)period_f;
    uint32_t un = (uint32_t)(period_f * (0.01 * (double)pnum(u, "phase")));
    uint32_t at = period - un;
    double fac = db2fac(pnum(u, "depth"));
    int att = u->attenuated; uint32_t since = u->in_state_since;
    for (int i = 0; i < n; i++) {
        double r = in[i];
        if (att && (since >= at)) { att = 0; since = 0; }
        else if (!att && (since >= un)) { att = 1; since = 0; }
        if (att) r *= fac;
        out[i] = r;
        since++;
    }
    u->attenuated = att; u->in_state_since = since;
}

/* ---- ringmodulator.go:18-47 --------------------------------------------------- */
static void ringmod_process(unit_t *u, const double *in, double *out, int n, uint32_t rate)
{
    double phase = u->phase;
    double frac = (MATH_TWO_PI * (double)pnum(u, "frequency")) / (double)rate;
    for (int i = 0; i < n; i++) {
        double cp = fmod(phase + ((double)i * frac), MATH_TWO_PI);
        out[i] = sin(cp) * in[i];
    }
    u->phase = fmod(phase + ((double)n * frac), MATH_TWO_PI);
}

/* ---- delay.go:17-91 ----------------------------------------------------------- */
static void delay_process(unit_t *u, const double *in, double *out, int n, uint32_t rate)
{
    int ds = (int)floor(((0.001 * (double)pnum(u, "delay_time")) * (double)rate) + 0.5);
    double fb = db2fac(pnum(u, "feedback")), lvl = db2fac(pnum(u, "level"));
    if (!u->delay_buf_valid || u->delay_buf.n != ds) { dbuf_resize_zero(&u->delay_buf, ds); u->delay_buf_valid = 1; }
    double *hist = u->delay_buf.buf;
    for (int i = 0; i < n; i++) {
        double dsmp = delayed(in, hist, ds, i - ds);
        out[i] = clamp1(lvl * (in[i] + (fb * dsmp)));
    }
    history_push(hist, ds, in, n);
}

/* ---- reverb.go:41-60 all-pass, 65-116 delay line, 179-338 Process ------------- */
static void reverb_process(unit_t *u, const double *in, double *out, int n, uint32_t rate)
{
    double wet = 0.01 * (double)pnum(u, "mix"), dry = 1.0 - wet;
    double sr = (double)rate;
    if (u->rv_rate != rate) {
        static const double ap_delays[3] = {0.04204, 0.01348, 0.00452};
        static const double taps_t[4] = {0.19196, 0.19996, 0.21596, 0.23204};
        static const double taps_c[4] = {0.1855, 0.18325, 0.17875, 0.17425};
        for (int i = 0; i < 3; i++) {
            int d = (int)round(ap_delays[i] * sr);
            free(u->allpass[i].buffer);
            u->allpass[i].buffer = (double *)calloc(d > 0 ? d : 1, sizeof(double));
            u->allpass[i].n = d; u->allpass[i].ptr = 0; u->allpass[i].feedback = 0.7;
        }
        uint32_t mx = 0;
        for (int i = 0; i < 4; i++) {
            u->rv_taps[i] = (uint32_t)round(taps_t[i] * sr);
            u->rv_coeffs[i] = taps_c[i];
            if (u->rv_taps[i] > mx) mx = u->rv_taps[i];
        }
        dbuf_resize_zero(&u->rv_delay, (int)mx);
        u->rv_rate = rate;
    }
    if (u->rv_front.n != n) dbuf_resize_zero(&u->rv_front, n);
    if (u->rv_back.n != n) dbuf_resize_zero(&u->rv_back, n);
    if (u->rv_line.n != n) dbuf_resize_zero(&u->rv_line, n);
    double *line = u->rv_line.buf, *front = u->rv_front.buf, *back = u->rv_back.buf;
    double *hist = u->rv_delay.buf; int hs = u->rv_delay.n;
    for (int i = 0; i < n; i++) {                         /* reverbDelayLine.process */
        double pre = 0.0;
        for (int j = 0; j < 4; j++) {
            int idx = i - (int)u->rv_taps[j];
            double cs = 0.0;
            if (idx >= 0) cs = in[idx];
            else if (idx >= -hs) cs = hist[hs + idx];
            pre += u->rv_coeffs[j] * cs;
        }
        line[i] = pre;
    }
    history_push(hist, hs, in, n);
    memcpy(front, line, sizeof(double) * n);
    for (int a = 0; a < 3; a++) {                         /* reverbAllpass.process */
        allpass_t *ap = &u->allpass[a];
        int w = ap->ptr;
        for (int i = 0; i < n; i++) {
            int r = (w + 1) % ap->n;
            double dsmp = ap->buffer[r];
            double pre = front[i] - (ap->feedback * dsmp);
            ap->buffer[w] = pre;
            back[i] = (ap->feedback * pre) + dsmp;
            w = r;
        }
        ap->ptr = w;
        double *t = front; front = back; back = t;
    }
    double half_wet = 0.5 * wet;
    for (int i = 0; i < n; i++)
        out[i] = clamp1((dry * in[i]) + (half_wet * (line[i] + front[i])));
    u->rv_front.buf = front; u->rv_back.buf = back;
}

/* ---- poweramp.go:25-127 compile, 186-216 Process ------------------------------ */
static orc_filter *poweramp_compile(unit_t *u, uint32_t rate)
{
    if (!u->bank) return NULL;
    uint32_t target = (uint32_t)strtoul(pdisc(u, "filter_order"), NULL, 10);
    orc_filter *filters[NUM_FILTERS] = {0};
    char pf[32], pl[32];
    for (int i = 0; i < NUM_FILTERS; i++) {
        snprintf(pf, sizeof pf, "filter_%d", i + 1);
        snprintf(pl, sizeof pl, "level_%d", i + 1);
        const char *name = pdisc(u, pf);
        int32_t level = pnum(u, pl);
        if (strcmp(name, STRING_NONE) != 0) {
            double fac = db2fac(level);
            orc_filter *flt = orc__irbank_create_filter(u->bank, name, rate);
            if (!flt) { for (int k = 0; k < i; k++) orc_filter_destroy(filters[k]); return NULL; }
            if (target > 0) { orc_filter *r = orc_filter_reduce(flt, target); orc_filter_destroy(flt); flt = r; }
            orc_filter *nrm = orc_filter_normalize(flt); orc_filter_destroy(flt);
            orc_filter *mul = orc_filter_multiply(nrm, fac); orc_filter_destroy(nrm);
            filters[i] = mul;
        }
    }
    orc_filter *comp = orc_filter_from_coefficients(NULL, 0, rate);          /* filter.Empty */
    for (int i = 0; i < NUM_FILTERS; i++) {
        orc_filter *s = orc__filter_add(comp, filters[i]);
        orc_filter_destroy(comp); orc_filter_destroy(filters[i]);
        comp = s;
    }
    return comp;
}
static void poweramp_recompile(unit_t *u)
{
    orc_filter *f = poweramp_compile(u, u->pa_rate);
    if (f) { orc_filter_destroy(u->pa_filter); u->pa_filter = f; }
}
static void poweramp_process(unit_t *u, const double *in, double *out, int n, uint32_t rate)
{
    if (rate != u->pa_rate) { u->pa_rate = rate; poweramp_recompile(u); }
    if (u->pa_filter) orc_filter_process(u->pa_filter, in, out, n);
    else memset(out, 0, sizeof(double) * n);
}

/* ---- cabinet.go:27-162 -------------------------------------------------------- */
static void cabinet_process(unit_t *u, const double *in, double *out, int n, uint32_t rate)
{
    static const double hp_f[3] = {300.0, 120.0, 80.0};
    static const double lp_f[4] = {3000.0, 4000.0, 5000.0, 6000.0};
    if (u->cab_buf.n != n) dbuf_resize_zero(&u->cab_buf, n);
    double *buf = u->cab_buf.buf;
    memcpy(buf, in, sizeof(double) * n);
    double m2pi = -MATH_TWO_PI / (double)rate;
    for (int i = 0; i < 3; i++) {
        double hcv = u->hcv[i];
        double k = 1.0 - exp(m2pi * hp_f[i]);
        for (int j = 0; j < n; j++) { double diff = buf[j] - hcv; buf[j] = diff; hcv += diff * k; }
        u->hcv[i] = hcv;
    }
    for (int i = 0; i < 4; i++) {
        double lcv = u->lcv[i];
        double k = 1.0 - exp(m2pi * lp_f[i]);
        for (int j = 0; j < n; j++) { double diff = buf[j] - lcv; buf[j] = lcv; lcv += diff * k; }
        u->lcv[i] = lcv;
    }
    for (int i = 0; i < n; i++) out[i] = clamp1(buf[i]);
}

/* ---- effects.go:443-516 CreateUnit -------------------------------------------- */
static unit_t *unit_create(int type)
{
    if (type < 0 || type >= UNIT_COUNT) return NULL;
    unit_t *u = (unit_t *)calloc(1, sizeof(*u));
    u->type = type;
#define GDG_UNIT(code, ident, name)
#define GDG_NUM(ident, name, phys, mn, mx, def) if (type == UNIT_##ident) add_num(u, name, phys, mn, mx, def);
#define GDG_DISC(ident, name, didx, vals) if (type == UNIT_##ident) add_disc(u, name, didx, vals);
#include "../include/gdg_unit_params.def"
#undef GDG_UNIT
#undef GDG_NUM
#undef GDG_DISC
    if (type == UNIT_EXCESS || type == UNIT_FUZZ || type == UNIT_OVERDRIVE || type == UNIT_DISTORTION) {
        u->os2 = orc_osd_create(2); u->os4 = orc_osd_create(4);
    }
    return u;
}
static void unit_destroy(unit_t *u)
{
    if (!u) return;
    for (int i = 0; i < u->nparams; i++) {
        param_t *p = &u->params[i];
        free(p->name); free(p->phys);
        for (int k = 0; k < p->ndisc; k++) free(p->disc[k]);
        free(p->disc);
    }
    free(u->params); free(u->delay_buf.buf); free(u->os_in.buf); free(u->os_out.buf);
    orc_osd_destroy(u->os2); orc_osd_destroy(u->os4);
    for (int i = 0; i < 3; i++) free(u->allpass[i].buffer);
    free(u->rv_delay.buf); free(u->rv_front.buf); free(u->rv_back.buf); free(u->rv_line.buf);
    free(u->cab_buf.buf); orc_filter_destroy(u->pa_filter); free(u);
}

/* poweramp.go:221-289 PreparePowerAmp: append filter_i / level_i parameters. */
static void prepare_poweramp(unit_t *u, orc_irbank *bank)
{
    if (!bank) return;
    size_t nn = orc_irbank_name_count(bank);
    for (int i = 0; i < NUM_FILTERS; i++) {
        char nm[32];
        char **vals = (char **)malloc(sizeof(char *) * (nn + 1));
        vals[0] = strdup(STRING_NONE);
        for (size_t k = 0; k < nn; k++) vals[k + 1] = strdup(orc_irbank_name(bank, k));
        snprintf(nm, sizeof nm, "filter_%d", i + 1);
        add_disc_list(u, nm, 0, vals, (int)nn + 1);
        snprintf(nm, sizeof nm, "level_%d", i + 1);
        add_num(u, nm, "dB", -60, 0, 0);
    }
    u->bank = bank;
}

static void unit_process(unit_t *u, const double *in, double *out, int n, uint32_t rate)
{
    switch (u->type) {
    case UNIT_SIGNALGENERATOR: signalgenerator_process(u, in, out, n, rate); break;
    case UNIT_NOISEGATE: noisegate_process(u, in, out, n, rate); break;
    case UNIT_BANDPASS: bandpass_process(u, in, out, n, rate); break;
    case UNIT_AUTOWAH: autowah_process(u, in, out, n, rate); break;
    case UNIT_AUTOYOY: autoyoy_process(u, in, out, n, rate); break;
    case UNIT_COMPRESSOR: compressor_process(u, in, out, n, rate); break;
    case UNIT_OCTAVER: octaver_process(u, in, out, n, rate); break;
    case UNIT_EXCESS: oversampled_process(u, excess_os, in, out, n, rate); break;
    case UNIT_FUZZ: oversampled_process(u, fuzz_os, in, out, n, rate); break;
    case UNIT_OVERDRIVE: oversampled_process(u, overdrive_os, in, out, n, rate); break;
    case UNIT_DISTORTION: oversampled_process(u, distortion_os, in, out, n, rate); break;
    case UNIT_TONESTACK: tonestack_process(u, in, out, n, rate); break;
    case UNIT_CHORUS: chorus_process(u, in, out, n, rate); break;
    case UNIT_FLANGER: flanger_phaser_process(u, in, out, n, rate, 0); break;
    case UNIT_PHASER: flanger_phaser_process(u, in, out, n, rate, 1); break;
    case UNIT_TREMOLO: tremolo_process(u, in, out, n, rate); break;
    case UNIT_RINGMODULATOR: ringmod_process(u, in, out, n, rate); break;
    case UNIT_DELAY: delay_process(u, in, out, n, rate); break;
    case UNIT_REVERB: reverb_process(u, in, out, n, rate); break;
    case UNIT_POWERAMP: poweramp_process(u, in, out, n, rate); break;
    case UNIT_CABINET: cabinet_process(u, in, out, n, rate); break;
    }
}

/* effects.go:127-172 setDiscreteValue, 232-249/266-317 setNumericValue (+ poweramp.go:132-181) */
static int unit_set_discrete(unit_t *u, const char *name, const char *value)
{
    int idx = -1;
    for (int i = 0; i < u->nparams; i++) if (!strcmp(u->params[i].name, name)) idx = i;
    if (idx < 0) return fail("Failed to set discrete value: Could not find parameter with name '%s'.", name);
    param_t *p = &u->params[idx];
    if (p->type != PT_DISCRETE) return fail("Failed to set discrete value: Parameter '%s' is not discrete.", name);
    int v = -1;
    for (int i = 0; i < p->ndisc; i++) if (!strcmp(p->disc[i], value)) v = i;
    if (v < 0) return fail("Failed to set discrete value: Value '%s' is not valid for parameter '%s'.", value, name);
    p->didx = v;
    if (u->type == UNIT_POWERAMP) poweramp_recompile(u);
    return 0;
}
static int unit_set_numeric(unit_t *u, const char *name, int32_t value)
{
    int idx = -1;
    for (int i = 0; i < u->nparams; i++) if (!strcmp(u->params[i].name, name)) idx = i;
    if (idx < 0) return fail("Failed to set numeric value: Could not find parameter with name '%s'.", name);
    param_t *p = &u->params[idx];
    if (p->type != PT_NUMERIC) return fail("Failed to set numeric value: Parameter '%s' is not numeric.", name);
    if (value < p->min || value > p->max)
        return fail("Failed to set numeric value: Parameter '%s' must be between '%d' and '%d' - got '%d'.", name, p->min, p->max, value);
    p->num = value;
    if (u->type == UNIT_POWERAMP) poweramp_recompile(u);
    return 0;
}

/* ======================================================================== */
/* signal/signal.go                                                          */
/* ======================================================================== */
typedef struct { unit_t *unit; int bypass; } slot_t;
struct orc_chain {
    double *buf_in, *buf_out; size_t buf_n;
    orc_irbank *bank;
    slot_t *slots; int n;
};

orc_chain *orc_chain_create(orc_irbank *bank)
{
    orc_chain *c = (orc_chain *)calloc(1, sizeof(*c));
    c->bank = bank;
    return c;
}
void orc_chain_destroy(orc_chain *c)
{
    if (!c) return;
    for (int i = 0; i < c->n; i++) unit_destroy(c->slots[i].unit);
    free(c->slots); free(c->buf_in); free(c->buf_out); free(c);
}
/* signal.go:52-86 AppendUnit: new units start BYPASSED. */
int orc_chain_append_unit(orc_chain *c, int type)
{
    unit_t *u = unit_create(type);
    if (!u) { fail("%s", "Failed to create effects unit."); return -1; }
    if (type == UNIT_POWERAMP) prepare_poweramp(u, c->bank);
    c->slots = (slot_t *)realloc(c->slots, sizeof(slot_t) * (c->n + 1));
    c->slots[c->n].unit = u; c->slots[c->n].bypass = 1;
    return c->n++;
}
int orc_chain_remove_unit(orc_chain *c, int id)
{
    if (id < 0 || id >= c->n) return fail("Cannot remove unit %d.", id);
    unit_destroy(c->slots[id].unit);
    memmove(&c->slots[id], &c->slots[id + 1], sizeof(slot_t) * (c->n - id - 1));
    c->n--;
    return 0;
}
int orc_chain_move_up(orc_chain *c, int id)
{
    if (id <= 0 || id >= c->n) return fail("Cannot move unit %d up.", id);
    slot_t t = c->slots[id]; c->slots[id] = c->slots[id - 1]; c->slots[id - 1] = t;
    return 0;
}
int orc_chain_move_down(orc_chain *c, int id)
{
    if (id < 0 || id >= c->n - 1) return fail("Cannot move unit %d down.", id);
    slot_t t = c->slots[id]; c->slots[id] = c->slots[id + 1]; c->slots[id + 1] = t;
    return 0;
}
int orc_chain_unit_type(orc_chain *c, int id)
{
    if (id < 0 || id >= c->n) { fail("Cannot get unit type: No unit %d.", id); return -1; }
    return c->slots[id].unit->type;
}
int orc_chain_set_bypass(orc_chain *c, int id, int bypass)
{
    if (id < 0 || id >= c->n) return fail("Cannot %s bypass: No unit %d.", bypass ? "enable" : "disable", id);
    c->slots[id].bypass = bypass != 0;
    return 0;
}
int orc_chain_get_bypass(orc_chain *c, int id, int *bypass)
{
    if (id < 0 || id >= c->n) return fail("Cannot get bypass value: No unit %d.", id);
    *bypass = c->slots[id].bypass;
    return 0;
}
int orc_chain_set_numeric(orc_chain *c, int id, const char *name, int32_t v)
{
    if (id < 0 || id >= c->n) return fail("Cannot set numeric value: No unit %d.", id);
    return unit_set_numeric(c->slots[id].unit, name, v);
}
int orc_chain_get_numeric(orc_chain *c, int id, const char *name, int32_t *v)
{
    if (id < 0 || id >= c->n) return fail("Cannot get numeric value: No unit %d.", id);
    unit_t *u = c->slots[id].unit; int idx = -1;
    for (int i = 0; i < u->nparams; i++) if (!strcmp(u->params[i].name, name)) idx = i;
    if (idx < 0) return fail("Failed to get numeric value: Could not find parameter with name '%s'.", name);
    if (u->params[idx].type != PT_NUMERIC) return fail("Failed to get numeric value: Parameter '%s' is not numeric.", name);
    *v = u->params[idx].num;
    return 0;
}
int orc_chain_set_discrete(orc_chain *c, int id, const char *name, const char *value)
{
    if (id < 0 || id >= c->n) return fail("Cannot set discrete value: No unit %d.", id);
    return unit_set_discrete(c->slots[id].unit, name, value);
}
int orc_chain_get_discrete(orc_chain *c, int id, const char *name, const char **value)
{
    if (id < 0 || id >= c->n) return fail("Cannot get discrete value: No unit %d.", id);
    unit_t *u = c->slots[id].unit; int idx = -1;
    for (int i = 0; i < u->nparams; i++) if (!strcmp(u->params[i].name, name)) idx = i;
    if (idx < 0) return fail("Failed to get discrete value: Could not find parameter with name '%s'.", name);
    if (u->params[idx].type != PT_DISCRETE) return fail("Failed to get discrete value: Parameter '%s' is not discrete.", name);
    *value = u->params[idx].disc[u->params[idx].didx];
    return 0;
}
int orc_chain_length(orc_chain *c) { return c->n; }

/* signal.go:361-414 Process: ping-pong through the non-bypassed slots. */
void orc_chain_process(orc_chain *c, const double *in, double *out, size_t n, uint32_t rate)
{
    if (c->buf_n != n) {
        free(c->buf_in); free(c->buf_out);
        c->buf_in = (double *)calloc(n ? n : 1, sizeof(double));
        c->buf_out = (double *)calloc(n ? n : 1, sizeof(double));
        c->buf_n = n;
    }
    double *bi = c->buf_in, *bo = c->buf_out;
    memcpy(bi, in, sizeof(double) * n);
    for (int s = 0; s < c->n; s++)
        if (!c->slots[s].bypass) {
            unit_process(c->slots[s].unit, bi, bo, (int)n, rate);
            double *t = bi; bi = bo; bo = t;
        }
    c->buf_in = bi; c->buf_out = bo;
    memcpy(out, bi, sizeof(double) * n);
}
