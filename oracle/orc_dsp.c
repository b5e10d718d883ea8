/*
 * orc_dsp.c - CPU ORACLE (test infrastructure, see orc.h): the DSP kernels
 * underneath the effects units - fft, resample, random, circular, PCM16 codec,
 * filter (FFT overlap-add FIR + IR algebra) and the oversampler/decimator.
 *
 * Restated from the reference's behaviour; citations are /root/reference paths.
 */
#define _GNU_SOURCE
#include "orc.h"
#include "../include/gdg_tables.h"

#include <complex.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef double complex cplx;

/* ======================================================================== */
/* fft/fft.go                                                                */
/* ======================================================================== */

/* fft.go:393-409 NextPowerOfTwo: smallest power of two >= value (value>=1). */
int orc_fft_next_power_of_two(uint64_t value, uint64_t *pow_out, uint32_t *exp_out)
{
    uint32_t digit = 0;
    uint64_t v = value;
    while (v) { digit++; v >>= 1; }          /* bits.Len64 */
    uint32_t e = digit - 1u;                   /* value==0 wraps like the Go code; never used */
    uint64_t p = (uint64_t)1 << (e & 63);
    if (p < value) { e++; p <<= 1; }
    if (pow_out) *pow_out = p;
    if (exp_out) *exp_out = e;
    return 0;
}

/* Twiddle tables W_n^k = exp(-2 pi i k / n), k in [0,n): fft.go:68-93,140-181. */
typedef struct { size_t n; cplx *w; size_t *perm; } fft_table;
static fft_table g_tables[40];

/* fft.go guards its lazily built global tables with a mutex (fft.go:140-240); so does this */
static pthread_mutex_t g_tables_mutex = PTHREAD_MUTEX_INITIALIZER;
static const fft_table *fft_tables(size_t n)
{
    uint32_t e = 0;
    while (((size_t)1 << e) < n) e++;
    fft_table *t = &g_tables[e];
    if (__atomic_load_n(&t->n, __ATOMIC_ACQUIRE) == n) return t;
    pthread_mutex_lock(&g_tables_mutex);
    if (t->n != n) {
        t->w = (cplx *)malloc(sizeof(cplx) * (n ? n : 1));
        t->perm = (size_t *)malloc(sizeof(size_t) * (n ? n : 1));
        for (size_t i = 0; i < n; i++) {
            double arg = (-2.0 * M_PI * (double)i) / (double)n;
            t->w[i] = cos(arg) + I * sin(arg);         /* cmplx.Exp(complex(0, arg)) */
            size_t r = 0;                               /* bit reversal: fft.go:98-135 */
            for (uint32_t b = 0; b < e; b++)
                if (i & ((size_t)1 << b)) r |= (size_t)1 << (e - 1 - b);
            t->perm[i] = r;
        }
        __atomic_store_n(&t->n, n, __ATOMIC_RELEASE);
    }
    pthread_mutex_unlock(&g_tables_mutex);
    return t;
}

/* fft.go:445-473 cooleyTukey: recursive radix-2 DIT (MODE_STANDARD). */
static void cooley_tukey(cplx *vec, size_t n)
{
    if (n <= 1) return;
    size_t h = n / 2;
    cplx *even = (cplx *)malloc(sizeof(cplx) * h), *odd = (cplx *)malloc(sizeof(cplx) * h);
    for (size_t i = 0; i < h; i++) { even[i] = vec[2 * i]; odd[i] = vec[2 * i + 1]; }
    cooley_tukey(even, h);
    cooley_tukey(odd, h);
    const fft_table *t = fft_tables(n);
    for (size_t i = 0; i < h; i++) {
        cplx p = t->w[i] * odd[i];
        vec[i] = even[i] + p;
        vec[h + i] = even[i] - p;
    }
    free(even); free(odd);
}

/* fft.go:528-607 permute + inplaceTransform: bit-reversal copy then log2(n)
 * butterfly passes using W[l] for the upper and W[l + n/2] for the lower output. */
static void inplace_transform(cplx *vec, size_t n)
{
    if (n <= 1) return;
    const fft_table *t = fft_tables(n);
    /* the reference keeps this buffer in the transform object and re-makes it only when the size changes (fft.go:528-551
     * scrapspace); one buffer per thread here - a malloc/free (mmap above 128 KB) per FFT is not part of the algorithm */
    static __thread cplx *scrap = NULL; static __thread size_t scrap_n = 0;
    if (scrap_n < n) { free(scrap); scrap = (cplx *)malloc(sizeof(cplx) * n); scrap_n = n; }
    memcpy(scrap, vec, sizeof(cplx) * n);
    for (size_t i = 0; i < n; i++) vec[i] = scrap[t->perm[i]];
    size_t stride = n;
    for (size_t size = 2; size <= n; size <<= 1) {
        stride >>= 1;
        size_t half = size >> 1;
        for (size_t off = 0; off < n; off += size)
            for (size_t k = 0; k < half; k++) {
                size_t i = off + k, j = i + half;
                cplx vi = vec[i], vj = vec[j];
                size_t l = k * stride;
                vec[i] = vi + t->w[l] * vj;
                vec[j] = vi + t->w[l + half * stride] * vj;
            }
    }
}

/* fft.go:612-667 Fourier. */
int orc_fft_fourier(double *v, size_t n, int scaling, int mode)
{
    cplx *vec = (cplx *)v;
    if (mode == ORC_MODE_STANDARD) cooley_tukey(vec, n);
    else if (mode == ORC_MODE_INPLACE) inplace_transform(vec, n);
    else return -1;
    if (scaling == ORC_SCALING_ORTHONORMAL) {
        double r = 1.0 / sqrt((double)n);
        for (size_t i = 0; i < n; i++) vec[i] *= r;
    }
    return 0;
}

/* fft.go:672-739 InverseFourier: swap re/im, forward transform, swap back, scale. */
int orc_fft_inverse_fourier(double *v, size_t n, int scaling, int mode)
{
    cplx *vec = (cplx *)v;
    double r = 0.0;
    if (scaling == ORC_SCALING_DEFAULT) r = 1.0 / (double)n;
    else if (scaling == ORC_SCALING_ORTHONORMAL) r = 1.0 / sqrt((double)n);
    if (mode != ORC_MODE_STANDARD && mode != ORC_MODE_INPLACE) return -1;
    for (size_t i = 0; i < n; i++) vec[i] = cimag(vec[i]) + I * creal(vec[i]);
    if (mode == ORC_MODE_STANDARD) cooley_tukey(vec, n); else inplace_transform(vec, n);
    for (size_t i = 0; i < n; i++) vec[i] = r * (cimag(vec[i]) + I * creal(vec[i]));
    return 0;
}

/* fft.go:744-856 RealFourier: pack even/odd into n/2 complex, FFT, split pass,
 * mirror conjugates, fix the centre element. */
int orc_fft_real_fourier(const double *in, double *o, size_t n, int scaling)
{
    cplx *out = (cplx *)o;
    if (n % 2 != 0) {
        if (n == 1) { out[0] = in[0]; return 0; }
        return -1;
    }
    size_t h = n / 2;
    for (size_t i = 0; i < h; i++) out[i] = in[2 * i] + I * in[2 * i + 1];
    orc_fft_fourier((double *)out, h, scaling, ORC_MODE_INPLACE);
    memcpy(out + h, out, sizeof(cplx) * h);
    const fft_table *t = fft_tables(n);
    for (size_t i = 0; i < h; i++) {
        size_t lo = h + i, hi = (i == 0) ? h : n - i;
        cplx low = out[lo], highc = conj(out[hi]);
        cplx coeff = I * t->w[i];
        out[i] = 0.5 * ((low + highc) - (coeff * (low - highc)));
    }
    for (size_t i = 1; i < h; i++) out[n - i] = conj(out[i]);
    cplx c = out[h], cc = conj(c);
    out[h] = 0.5 * ((c + cc) + (I * (c - cc)));
    if (scaling == ORC_SCALING_ORTHONORMAL)
        for (size_t i = 0; i < n; i++) out[i] *= M_SQRT1_2;
    return 0;
}

/* fft.go:863-990 RealInverseFourier (averages the conjugate-symmetric halves,
 * clobbers its input like the reference). */
int orc_fft_real_inverse_fourier(double *iv, double *out, size_t n, int scaling)
{
    cplx *in = (cplx *)iv;
    if (n % 2 != 0) {
        if (n == 1) { out[0] = creal(in[0]); return 0; }
        return -1;
    }
    size_t h = n / 2;
    for (size_t i = 1; i < h; i++) in[i] = 0.5 * (in[i] + conj(in[n - i]));
    double dc = creal(in[0]), ny = creal(in[h]);
    memcpy(in + h, in, sizeof(cplx) * h);
    const fft_table *t = fft_tables(n);
    for (size_t i = 0; i < h; i++) {
        size_t lo = h + i, hi = (i == 0) ? h : n - i;
        cplx low = in[lo], highc = conj(in[hi]);
        cplx even = low + highc;
        cplx odd = (low - highc) * conj(t->w[i]);
        in[i] = 0.5 * (even + (I * odd));
    }
    in[0] = 0.5 * (dc + ny) + I * (0.5 * (dc - ny));
    for (size_t i = h; i < n; i++) in[i] = 0.0;
    orc_fft_inverse_fourier((double *)in, h, scaling, ORC_MODE_INPLACE);
    for (size_t i = 0; i < h; i++) { out[2 * i] = creal(in[i]); out[2 * i + 1] = cimag(in[i]); }
    if (scaling == ORC_SCALING_ORTHONORMAL)
        for (size_t i = 0; i < n; i++) out[i] *= M_SQRT2;
    return 0;
}

/* fft.go:467-523 Shift (fftshift / ifftshift by element swaps). */
void orc_fft_shift(double *v, size_t n, int inverse)
{
    cplx *vec = (cplx *)v;
    size_t n_neg = n >> 1, n_pos = n_neg;
    int odd = (n & 1) != 0;
    if (odd) n_pos++;
    size_t a = 0, b = inverse ? n_neg : n_pos;
    while (b < n) { cplx t = vec[a]; vec[a] = vec[b]; vec[b] = t; a++; b++; }
    if (odd) {
        if (inverse) {
            long pb = (long)n - 1, pa = pb - 1;
            while (pa >= (long)n_pos) { cplx t = vec[pa]; vec[pa] = vec[pb]; vec[pb] = t; pa--; pb--; }
        } else {
            b = a + 1;
            while (b < n) { cplx t = vec[a]; vec[a] = vec[b]; vec[b] = t; a++; b++; }
        }
    }
}

/* ======================================================================== */
/* resample/resample.go                                                      */
/* ======================================================================== */

/* resample.go:10-31 lanczosKernel. */
static double lanczos_kernel(double x, double a)
{
    if (x == 0) return 1.0;
    if ((-a < x) && (x < a)) {
        double pix = M_PI * x;
        double pixa = pix / a;
        double pix2 = pix * pix;
        double prod = sin(pix) * sin(pixa);
        return (a * prod) / pix2;
    }
    return 0.0;
}

/* resample.go:36-66 lanczosInterpolate: 2a taps around x, out-of-range taps skipped. */
static double lanczos_interpolate(const double *s, long n, double x, int a)
{
    long idx = (long)floor(x);
    long lb = idx + 1 - a, ub = idx + 1 + a;
    double sum = 0.0;
    for (long i = lb; i < ub; i++)
        if (i >= 0 && i < n) {
            double diff = x - (double)i;
            sum += s[i] * lanczos_kernel(diff, (double)a);
        }
    return sum;
}

/* resample.go:72-103 Time: output length drops one sample when the ratio is integral. */
size_t orc_resample_time_length(size_t n, uint32_t source_rate, uint32_t target_rate)
{
    double expansion = (double)target_rate / (double)source_rate;
    double lf = (double)n * expansion;
    double fl = floor(lf);
    long len = (long)fl;
    if (fl == lf) len--;
    return len < 0 ? 0 : (size_t)len;
}

void orc_resample_time(const double *samples, size_t n, uint32_t source_rate, uint32_t target_rate, double *out)
{
    size_t len = orc_resample_time_length(n, source_rate, target_rate);
    double dx = (double)source_rate / (double)target_rate;
    for (size_t i = 0; i < len; i++)
        out[i] = lanczos_interpolate(samples, (long)n, (double)i * dx, 3);
}

/* resample.go:109-142 Frequency: Lanczos-resample real and imaginary parts of bins. */
void orc_resample_frequency(const double *bins, size_t nb, uint32_t nt, double *out)
{
    double *re = (double *)malloc(sizeof(double) * (nb ? nb : 1)), *im = (double *)malloc(sizeof(double) * (nb ? nb : 1));
    for (size_t i = 0; i < nb; i++) { re[i] = bins[2 * i]; im[i] = bins[2 * i + 1]; }
    double dx = (double)nb / (double)nt;
    for (uint32_t i = 0; i < nt; i++) {
        double x = (double)i * dx;
        out[2 * i] = lanczos_interpolate(re, (long)nb, x, 3);
        out[2 * i + 1] = lanczos_interpolate(im, (long)nb, x, 3);
    }
    free(re); free(im);
}

/* resample.go:148-176 Oversample: integer phases copy, others interpolate. */
void orc_resample_oversample(const double *source, size_t ns, double *target, size_t nt, uint32_t factor)
{
    double dx = 1.0 / (double)factor;
    for (size_t i = 0; i < nt; i++) {
        if (i % factor == 0) target[i] = source[i / factor];
        else target[i] = lanczos_interpolate(source, (long)ns, (double)i * dx, 3);
    }
}

/* ======================================================================== */
/* random/random.go                                                          */
/* ======================================================================== */

/* random.go:40-55 CreatePRNG (uint64 wrap-around arithmetic as in Go). */
void orc_prng_seed(orc_prng *g, uint64_t seed)
{
    g->a = 16807; g->b = 0; g->n = ((uint64_t)1 << 31) - 1;
    g->x = ((64979ull * seed) + 83ull) % g->n;
}

/* random.go:23-35 NextFloat. */
double orc_prng_next_float(orc_prng *g)
{
    g->x = ((g->a * g->x) + g->b) % g->n;
    return (double)g->x / (double)(g->n - 1);
}

/* ======================================================================== */
/* circular/circular.go                                                      */
/* ======================================================================== */
struct orc_circular { double *values; size_t n; size_t pointer; };

orc_circular *orc_circular_create(size_t size)
{
    orc_circular *c = (orc_circular *)calloc(1, sizeof(*c));
    c->values = (double *)calloc(size ? size : 1, sizeof(double));
    c->n = size;
    return c;
}
void orc_circular_destroy(orc_circular *c) { if (c) { free(c->values); free(c); } }

/* circular.go:33-71 Enqueue. */
void orc_circular_enqueue(orc_circular *c, const double *elems, size_t m)
{
    size_t n = c->n;
    if (m >= n) {
        memcpy(c->values, elems + (m - n), sizeof(double) * n);
        c->pointer = 0;
    } else {
        size_t ptr = c->pointer, inc = ptr + m;
        if (inc < n) {
            memcpy(c->values + ptr, elems, sizeof(double) * m);
            c->pointer = inc;
        } else {
            size_t head = inc - n, tail = n - ptr;
            memcpy(c->values + ptr, elems, sizeof(double) * tail);
            memcpy(c->values, elems + tail, sizeof(double) * head);
            c->pointer = head;
        }
    }
}
size_t orc_circular_length(const orc_circular *c) { return c->n; }

/* circular.go:85-105 Retrieve (oldest first). */
int orc_circular_retrieve(const orc_circular *c, double *buf, size_t m)
{
    if (m != c->n) return -1;
    size_t tail = c->n - c->pointer;
    memcpy(buf, c->values + c->pointer, sizeof(double) * tail);
    memcpy(buf + tail, c->values, sizeof(double) * c->pointer);
    return 0;
}

/* ======================================================================== */
/* wave/wave.go - PCM16 only                                                 */
/* ======================================================================== */

/* wave.go:400-428 bytesToSamplesLPCM16: scaling = 2 / (MaxInt16 - MinInt16). */
void orc_wave_pcm16_decode(const int16_t *pcm, size_t n, double *out)
{
    double scaling = 2.0 / (32767.0 - (-32768.0));
    for (size_t i = 0; i < n; i++) out[i] = scaling * (double)pcm[i];
}

/* wave.go:347-379 samplesToBytesLPCM16: clamp, truncate int32(32767.5 * x), clamp. */
void orc_wave_pcm16_encode(const double *samples, size_t n, int16_t *out)
{
    double scale = 0.5 * 65535.0;
    for (size_t i = 0; i < n; i++) {
        double s = samples[i];
        if (s < -1.0) s = -1.0; else if (s > 1.0) s = 1.0;
        int32_t tmp = (int32_t)(scale * s);
        if (tmp > 32767) tmp = 32767; else if (tmp < -32768) tmp = -32768;
        out[i] = (int16_t)tmp;
    }
}

/* all six sample formats of wave.go, on the data-chunk payload (little endian) */
size_t orc_wave_bytes_per_sample(int fmt)
{
    static const size_t b[6] = {1, 2, 3, 4, 4, 8};
    return (fmt >= 0 && fmt < 6) ? b[fmt] : 0;
}
int orc_wave_decode(int fmt, const unsigned char *d, size_t n, double *out)
{
    switch (fmt) {
    case ORC_FMT_PCM8: {                                     /* wave.go:316-342 */
        const double scale = 1.0 / 127.0;
        for (size_t i = 0; i < n; i++) {
            const int16_t temp = (int16_t)((int16_t)d[i] + (-128));
            const double r = scale * (double)temp;
            out[i] = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
        }
        return 0; }
    case ORC_FMT_PCM16: {                                    /* wave.go:400-428 */
        const double scaling = 2.0 / (32767.0 - (-32768.0));
        for (size_t i = 0; i < n; i++) { const int16_t v = (int16_t)((uint16_t)d[2 * i] | ((uint16_t)d[2 * i + 1] << 8)); out[i] = scaling * (double)v; }
        return 0; }
    case ORC_FMT_PCM24: {                                    /* wave.go:483-531 */
        const double scaling = 2.0 / (8388607.0 - (-8388608.0));
        for (size_t i = 0; i < n; i++) {
            const uint32_t w = (uint32_t)d[3 * i] | ((uint32_t)d[3 * i + 1] << 8) | ((uint32_t)d[3 * i + 2] << 16);
            int32_t v = (int32_t)w;
            if (w & 0x00800000u) v = -8388608 + (v & 0x007fffff);
            out[i] = scaling * (double)v;
        }
        return 0; }
    case ORC_FMT_PCM32: {                                    /* wave.go:589-617 */
        const double scaling = 2.0 / (2147483647.0 - (-2147483648.0));
        for (size_t i = 0; i < n; i++) {
            const uint32_t w = (uint32_t)d[4 * i] | ((uint32_t)d[4 * i + 1] << 8) | ((uint32_t)d[4 * i + 2] << 16) | ((uint32_t)d[4 * i + 3] << 24);
            out[i] = scaling * (double)(int32_t)w;
        }
        return 0; }
    case ORC_FMT_IEEE32:                                     /* wave.go:662-689 */
        for (size_t i = 0; i < n; i++) { float f; memcpy(&f, d + 4 * i, 4); out[i] = (double)f; }
        return 0;
    case ORC_FMT_IEEE64:                                     /* wave.go:714-732 */
        memcpy(out, d, 8 * n);
        return 0;
    }
    return -1;
}
static double wave_clamp(double s) { return s < -1.0 ? -1.0 : (s > 1.0 ? 1.0 : s); }
int orc_wave_encode(int fmt, const double *x, size_t n, unsigned char *o)
{
    switch (fmt) {
    case ORC_FMT_PCM8:                                       /* wave.go:275-311 */
        for (size_t i = 0; i < n; i++) {
            const int16_t temp = (int16_t)(127.0 * wave_clamp(x[i]));
            const int res = (int)temp - (-128);
            o[i] = res < 0 ? 0 : (res > 255 ? 255 : (unsigned char)res);
        }
        return 0;
    case ORC_FMT_PCM16: {                                    /* wave.go:347-395 */
        const double scale = 0.5 * 65535.0;
        for (size_t i = 0; i < n; i++) {
            int32_t t = (int32_t)(scale * wave_clamp(x[i]));
            if (t > 32767) t = 32767; else if (t < -32768) t = -32768;
            const uint16_t u = (uint16_t)(int16_t)t;
            o[2 * i] = (unsigned char)(u & 0xff); o[2 * i + 1] = (unsigned char)(u >> 8);
        }
        return 0; }
    case ORC_FMT_PCM24: {                                    /* wave.go:433-478 */
        const double scale = 0.5 * 16777215.0;
        for (size_t i = 0; i < n; i++) {
            int32_t t = (int32_t)(scale * wave_clamp(x[i]));
            if (t > 8388607) t = 8388607; else if (t < -8388608) t = -8388608;
            const uint32_t u = (uint32_t)t;
            o[3 * i] = (unsigned char)(u & 0xff); o[3 * i + 1] = (unsigned char)((u >> 8) & 0xff); o[3 * i + 2] = (unsigned char)((u >> 16) & 0xff);
        }
        return 0; }
    case ORC_FMT_PCM32: {                                    /* wave.go:536-584 */
        const double scale = 0.5 * 4294967295.0;
        for (size_t i = 0; i < n; i++) {
            int64_t t = (int64_t)(scale * wave_clamp(x[i]));
            if (t > 2147483647LL) t = 2147483647LL; else if (t < -2147483648LL) t = -2147483648LL;
            const uint32_t u = (uint32_t)(int32_t)t;
            o[4 * i] = (unsigned char)(u & 0xff); o[4 * i + 1] = (unsigned char)((u >> 8) & 0xff);
            o[4 * i + 2] = (unsigned char)((u >> 16) & 0xff); o[4 * i + 3] = (unsigned char)(u >> 24);
        }
        return 0; }
    case ORC_FMT_IEEE32:                                     /* wave.go:622-657 */
        for (size_t i = 0; i < n; i++) { const float f = (float)wave_clamp(x[i]); memcpy(o + 4 * i, &f, 4); }
        return 0;
    case ORC_FMT_IEEE64:                                     /* wave.go:694-709: no clamp */
        memcpy(o, x, 8 * n);
        return 0;
    }
    return -1;
}

/* ======================================================================== */
/* filter/filter.go                                                          */
/* ======================================================================== */
struct orc_filter {
    double *coeffs; size_t len;             /* impulseResponse.data */
    uint32_t sample_rate;
    double gain_compensation;
    /* Process state (filter.go:62-72) */
    size_t fft_size;                        /* len(filterComplex) */
    cplx *filter_complex, *filtered_complex;
    double *input_buffer, *output_buffer, *tail_buffer;
};

static orc_filter *filter_new(const double *coeffs, size_t n, uint32_t rate, double comp)
{
    orc_filter *f = (orc_filter *)calloc(1, sizeof(*f));
    f->coeffs = (double *)malloc(sizeof(double) * (n ? n : 1));
    if (n) memcpy(f->coeffs, coeffs, sizeof(double) * n);
    f->len = n; f->sample_rate = rate; f->gain_compensation = comp;
    return f;
}
orc_filter *orc_filter_from_coefficients(const double *coeffs, size_t n, uint32_t rate) { return filter_new(coeffs, n, rate, 0.0); }
void orc_filter_destroy(orc_filter *f)
{
    if (!f) return;
    free(f->coeffs); free(f->filter_complex); free(f->filtered_complex);
    free(f->input_buffer); free(f->output_buffer); free(f->tail_buffer); free(f);
}
size_t orc_filter_length(const orc_filter *f) { return f->len; }
const double *orc_filter_coefficients(const orc_filter *f) { return f->coeffs; }
void orc_filter_set_compensation(orc_filter *f, double g) { f->gain_compensation = g; }

/* filter.go:270-323 Multiply (the product carries gainCompensation 0). */
orc_filter *orc_filter_multiply(const orc_filter *f, double scalar)
{
    orc_filter *r = filter_new(f->coeffs, f->len, f->sample_rate, 0.0);
    for (size_t i = 0; i < r->len; i++) r->coeffs[i] = scalar * f->coeffs[i];
    return r;
}

/* filter.go:127-138 estimateGain + 328-336 Normalize. */
orc_filter *orc_filter_normalize(const orc_filter *f)
{
    double sum = 0.0;
    for (size_t i = 0; i < f->len; i++) sum += f->coeffs[i] * f->coeffs[i];
    double gain = sqrt(sum);
    return orc_filter_multiply(f, f->gain_compensation / gain);
}

/* filter.go:167-253 Add: zero-extend to the longer and sum (other may be NULL). */
static orc_filter *filter_add(const orc_filter *a, const orc_filter *b)
{
    if (!b) return filter_new(a->coeffs, a->len, a->sample_rate, a->gain_compensation);
    if (a->sample_rate != b->sample_rate) return NULL;
    size_t n = a->len > b->len ? a->len : b->len;
    double *c = (double *)calloc(n ? n : 1, sizeof(double));
    memcpy(c, a->coeffs, sizeof(double) * a->len);
    for (size_t i = 0; i < b->len; i++) c[i] += b->coeffs[i];
    orc_filter *r = filter_new(c, n, a->sample_rate, 0.0);
    free(c);
    return r;
}

/* filter.go:520-604 Reduce: FFT, Lanczos-resample the positive-frequency bins,
 * rebuild the conjugate half, inverse FFT, truncate to `order`. */
orc_filter *orc_filter_reduce(const orc_filter *f, uint32_t order)
{
    uint64_t n = f->len, n_src, n_tgt;
    if (n <= (uint64_t)order) return filter_new(f->coeffs, f->len, f->sample_rate, f->gain_compensation);
    orc_fft_next_power_of_two(n, &n_src, NULL);
    orc_fft_next_power_of_two(order, &n_tgt, NULL);
    double *padded = (double *)calloc(n_src, sizeof(double));
    memcpy(padded, f->coeffs, sizeof(double) * n);
    cplx *fr = (cplx *)calloc(n_src, sizeof(cplx));
    orc_fft_real_fourier(padded, (double *)fr, n_src, ORC_SCALING_DEFAULT);
    uint32_t pos_src = (uint32_t)(n_src >> 1) + 1;
    uint32_t half_tgt = (uint32_t)(n_tgt >> 1), pos_tgt = half_tgt + 1;
    cplx *fr_pos = (cplx *)calloc(pos_tgt, sizeof(cplx));
    orc_resample_frequency((double *)fr, pos_src, pos_tgt, (double *)fr_pos);
    cplx *fr_new = (cplx *)calloc(n_tgt, sizeof(cplx));
    memcpy(fr_new, fr_pos, sizeof(cplx) * (pos_tgt < n_tgt ? pos_tgt : n_tgt));
    for (uint32_t i = 1; i < half_tgt; i++) fr_new[n_tgt - i] = conj(fr_pos[i]);
    double *target = (double *)calloc(n_tgt, sizeof(double));
    orc_fft_real_inverse_fourier((double *)fr_new, target, n_tgt, ORC_SCALING_DEFAULT);
    orc_filter *r = filter_new(target, order, f->sample_rate, f->gain_compensation);
    free(padded); free(fr); free(fr_pos); free(fr_new); free(target);
    return r;
}

/* filter.go:342-515 Process: overlap-add fast convolution with the hard clamp to
 * [-1, 1] on every emitted sample; block = nextPow2(L), FFT = 2*block,
 * numBlocks = ceil(nextPow2(N) / block). */
int orc_filter_process(orc_filter *f, const double *in, double *out, size_t N)
{
    size_t L = f->len;
    if (L == 0) { memset(out, 0, sizeof(double) * N); return 0; }
    uint64_t npow, bs;
    orc_fft_next_power_of_two(N, &npow, NULL);
    orc_fft_next_power_of_two(L, &bs, NULL);
    uint64_t nb = npow / bs;
    if (npow % bs != 0) nb++;
    size_t F = (size_t)(bs << 1);
    for (uint64_t i = 0; i < nb; i++) {
        if (f->fft_size != F) {
            free(f->filter_complex); free(f->filtered_complex);
            free(f->input_buffer); free(f->output_buffer); free(f->tail_buffer);
            double *padded = (double *)calloc(F, sizeof(double));
            memcpy(padded, f->coeffs, sizeof(double) * L);
            f->filter_complex = (cplx *)calloc(F, sizeof(cplx));
            orc_fft_real_fourier(padded, (double *)f->filter_complex, F, ORC_SCALING_DEFAULT);
            free(padded);
            f->filtered_complex = (cplx *)calloc(F, sizeof(cplx));
            f->input_buffer = (double *)calloc(F, sizeof(double));
            f->output_buffer = (double *)calloc(F, sizeof(double));
            f->tail_buffer = (double *)calloc(F, sizeof(double));
            f->fft_size = F;
        }
        uint64_t lb = i * bs, ub = lb + bs;
        if (ub > N) ub = N;
        if (lb > ub) return -2;                 /* the reference panics here */
        size_t ns = (size_t)(ub - lb);
        memcpy(f->input_buffer, in + lb, sizeof(double) * ns);
        memset(f->input_buffer + ns, 0, sizeof(double) * (F - ns));
        orc_fft_real_fourier(f->input_buffer, (double *)f->filtered_complex, F, ORC_SCALING_DEFAULT);
        for (size_t k = 0; k < F; k++) f->filtered_complex[k] *= f->filter_complex[k];   /* hadamardComplex */
        orc_fft_real_inverse_fourier((double *)f->filtered_complex, f->output_buffer, F, ORC_SCALING_DEFAULT);
        for (size_t j = 0; j < F; j++) {
            double pre = f->output_buffer[j] + f->tail_buffer[j];
            if (j < ns) {
                if (pre > 1.0) out[lb + j] = 1.0;
                else if (pre < -1.0) out[lb + j] = -1.0;
                else out[lb + j] = pre;
            } else {
                f->tail_buffer[j - ns] = pre;
            }
        }
        memset(f->tail_buffer + (F - ns), 0, sizeof(double) * ns);
    }
    return 0;
}

/* ---- impulse-response bank: filter.go:25-33 g_sampleRates, 704-802 Import ---- */
typedef struct { char *name; uint32_t rate; double comp; double *data; size_t n; } ir_entry;
struct orc_irbank { ir_entry *e; size_t n, cap; };
static const uint32_t g_sample_rates[7] = {22050, 32000, 44100, 48000, 88200, 96000, 192000};

orc_irbank *orc_irbank_create(void) { return (orc_irbank *)calloc(1, sizeof(orc_irbank)); }
void orc_irbank_destroy(orc_irbank *b)
{
    if (!b) return;
    for (size_t i = 0; i < b->n; i++) { free(b->e[i].name); free(b->e[i].data); }
    free(b->e); free(b);
}
static void irbank_push(orc_irbank *b, const char *name, uint32_t rate, double comp, double *data, size_t n)
{
    if (b->n == b->cap) { b->cap = b->cap ? 2 * b->cap : 64; b->e = (ir_entry *)realloc(b->e, sizeof(ir_entry) * b->cap); }
    ir_entry *e = &b->e[b->n++];
    e->name = strdup(name); e->rate = rate; e->comp = comp; e->data = data; e->n = n;
}
int orc_irbank_add(orc_irbank *b, const char *name, int32_t comp_db, const double *samples, size_t n, uint32_t src_rate)
{
    double fac = pow(10.0, 0.05 * (double)comp_db);          /* filter.go:742-745 */
    for (int r = 0; r < 7; r++) {
        size_t len = orc_resample_time_length(n, src_rate, g_sample_rates[r]);
        double *d = (double *)malloc(sizeof(double) * (len ? len : 1));
        orc_resample_time(samples, n, src_rate, g_sample_rates[r], d);
        irbank_push(b, name, g_sample_rates[r], fac, d, len);
    }
    return 0;
}
int orc_irbank_add_raw(orc_irbank *b, const char *name, int32_t comp_db, const double *samples, size_t n, uint32_t rate)
{
    double *d = (double *)malloc(sizeof(double) * (n ? n : 1));
    memcpy(d, samples, sizeof(double) * n);
    irbank_push(b, name, rate, pow(10.0, 0.05 * (double)comp_db), d, n);
    return 0;
}
/* filter.go:681-699 Names: distinct names in first-seen order. */
size_t orc_irbank_name_count(const orc_irbank *b)
{
    size_t cnt = 0;
    for (size_t i = 0; i < b->n; i++) {
        int seen = 0;
        for (size_t j = 0; j < i; j++) if (!strcmp(b->e[j].name, b->e[i].name)) { seen = 1; break; }
        if (!seen) cnt++;
    }
    return cnt;
}
const char *orc_irbank_name(const orc_irbank *b, size_t k)
{
    size_t cnt = 0;
    for (size_t i = 0; i < b->n; i++) {
        int seen = 0;
        for (size_t j = 0; j < i; j++) if (!strcmp(b->e[j].name, b->e[i].name)) { seen = 1; break; }
        if (!seen) { if (cnt == k) return b->e[i].name; cnt++; }
    }
    return NULL;
}
/* filter.go:619-660 CreateFilter (first match of name & rate), used by orc_units.c */
orc_filter *orc__irbank_create_filter(const orc_irbank *b, const char *name, uint32_t rate)
{
    if (!b) return NULL;
    for (size_t i = 0; i < b->n; i++)
        if (!strcmp(b->e[i].name, name) && b->e[i].rate == rate)
            return filter_new(b->e[i].data, b->e[i].n, rate, b->e[i].comp);
    return NULL;
}
orc_filter *orc__filter_add(const orc_filter *a, const orc_filter *b) { return filter_add(a, b); }

/* ======================================================================== */
/* oversampling/oversampling.go                                              */
/* ======================================================================== */
#define LOOKAHEAD_ONE_SIDE 4
#define LOOKAHEAD_BOTH_SIDES 8
#define ATTENUATION_HALF_DECIBEL 0.9440608762859234

struct orc_osd {
    uint32_t factor;
    orc_filter *aa;
    double *pre; size_t pre_n;       /* bufferPreUpsampling */
    double *post; size_t post_n;     /* bufferPostUpsampling */
    double *dec; size_t dec_n;       /* bufferPreDecimation */
};

/* oversampling.go:189-532 CreateOversamplerDecimator. */
orc_osd *orc_osd_create(uint32_t factor)
{
    static const double half2[] = GDG_AA2_HALF_INIT;
    static const double half4[] = GDG_AA4_HALF_INIT;
    if (factor != 1 && factor != 2 && factor != 4) return NULL;
    orc_osd *o = (orc_osd *)calloc(1, sizeof(*o));
    o->factor = factor;
    if (factor > 1) {
        const double *h = factor == 2 ? half2 : half4;
        size_t L = factor == 2 ? GDG_AA2_TAPS : GDG_AA4_TAPS;
        double *c = (double *)malloc(sizeof(double) * L);
        for (size_t k = 0; k < L; k++) c[k] = h[k <= L - 1 - k ? k : L - 1 - k];
        o->aa = orc_filter_from_coefficients(c, L, 0);
        free(c);
    }
    return o;
}
void orc_osd_destroy(orc_osd *o)
{
    if (!o) return;
    orc_filter_destroy(o->aa); free(o->pre); free(o->post); free(o->dec); free(o);
}

/* oversampling.go:54-115 Oversample: keep the last 8 inputs, Lanczos-upsample
 * (N+7)f+1 points, emit [4f, 4f+Nf) => 4-sample look-ahead latency. */
int orc_osd_oversample(orc_osd *o, const double *in, size_t n_in, double *out, size_t n_out)
{
    size_t f = o->factor;
    if (f <= 1) { memcpy(out, in, sizeof(double) * (n_in < n_out ? n_in : n_out)); return 0; }
    if (n_out != n_in * f) return -1;
    size_t pre_n = n_in + LOOKAHEAD_BOTH_SIDES;
    if (o->pre_n != pre_n) { free(o->pre); o->pre = (double *)calloc(pre_n, sizeof(double)); o->pre_n = pre_n; }
    memmove(o->pre, o->pre + (pre_n - LOOKAHEAD_BOTH_SIDES), sizeof(double) * LOOKAHEAD_BOTH_SIDES);
    memcpy(o->pre + LOOKAHEAD_BOTH_SIDES, in, sizeof(double) * n_in);
    size_t post_n = ((pre_n - 1) * f) + 1;
    if (o->post_n != post_n) { free(o->post); o->post = (double *)calloc(post_n, sizeof(double)); o->post_n = post_n; }
    orc_resample_oversample(o->pre, pre_n, o->post, post_n, (uint32_t)f);
    memcpy(out, o->post + LOOKAHEAD_ONE_SIDE * f, sizeof(double) * n_out);
    return 0;
}

/* oversampling.go:126-184 Decimate: anti-alias FIR over all f*N samples (with its
 * clamp), keep every f-th, scale by -0.5 dB. */
int orc_osd_decimate(orc_osd *o, const double *in, size_t n_in, double *out, size_t n_out)
{
    size_t f = o->factor;
    if (f <= 1) { memcpy(out, in, sizeof(double) * (n_in < n_out ? n_in : n_out)); return 0; }
    if (o->dec_n != n_in) { free(o->dec); o->dec = (double *)calloc(n_in ? n_in : 1, sizeof(double)); o->dec_n = n_in; }
    int err = orc_filter_process(o->aa, in, o->dec, n_in);
    if (err) return err;
    for (size_t i = 0; i < n_out; i++) {
        size_t idx = f * i;
        out[i] = idx < n_in ? ATTENUATION_HALF_DECIBEL * o->dec[idx] : 0.0;
    }
    return 0;
}
