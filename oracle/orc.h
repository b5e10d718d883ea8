/*
 * orc.h - CPU ORACLE for the go-dsp-guitar effects-chain hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C, float64, single-threaded
 * restatement of the reference's Go arithmetic (same block partitioning, same
 * state carry, same clamp points, same quirks), written from the reference's
 * behaviour - every function cites the reference file:line it follows.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it; nothing under go-dsp-guitar_b200/ links, imports or calls it.
 *
 * Pinned against: fft/fft_test.go, oversampling/oversampling_test.go,
 * resample/resample_test.go, random/random_test.go, circular/circular_test.go and
 * tuner/tuner_test.go (+ tuner/samples/ WAV tails) - see tests/test_oracle_goldens.py.
 * Units, chain driver, IR algebra and spatializer have NO tests in the reference
 * ("parity unpinned by the reference").  For those rows the pin is
 * (i) OUTPUTS OF THE REFERENCE'S OWN SOURCES: tests/golden/go2py translates the .go
 * files mechanically to Python (its translation is validated on the reference's own
 * unit-test vectors), executes every unit / the chain driver / the power amp's IR
 * algebra / the spatializer / the level meter, and this oracle must reproduce the 435
 * recorded streams at <= 1e-12 (tests/test_reference_translated.py, vectors committed
 * under tests/golden/translated_reference_vectors.*); further
 * (ii) a second restatement written independently from the Go sources in Python
 * (tests/goref.py, parameter tables extracted from the Go Parameter{} literals),
 * diffed against this oracle at <= 1e-12 per unit / parameter corner / CHAIN12 /
 * Reduce-Normalize-Multiply-Add / filter.Import / spatializer, and (iii) analytic
 * closed forms (distortion = clamp, ECC83 = tanh(g x / 2), excess = triangle fold,
 * ...): tests/test_oracle_crosscheck.py, part of the CPU suite.
 *
 * The Go toolchain is absent from the build image, so the reference binary itself
 * can not be executed; math.Sin/Exp/Atan/Pow/Log10 are replaced by glibc libm
 * (<= 1-2 ulp apart; far inside the 1e-9 RMS budget).
 */
#ifndef ORC_H
#define ORC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- fft/fft.go ---------------------------------------------------------- */
#define ORC_SCALING_DEFAULT 0
#define ORC_SCALING_ORTHONORMAL 1
#define ORC_MODE_STANDARD 2
#define ORC_MODE_INPLACE 3
/* complex vectors are interleaved (re, im) doubles */
int orc_fft_next_power_of_two(uint64_t value, uint64_t *pow, uint32_t *exp);
int orc_fft_fourier(double *vec, size_t n, int scaling, int mode);          /* in place */
int orc_fft_inverse_fourier(double *vec, size_t n, int scaling, int mode);  /* in place */
int orc_fft_real_fourier(const double *in, double *out, size_t n, int scaling);
int orc_fft_real_inverse_fourier(double *in, double *out, size_t n, int scaling); /* clobbers in, like the reference */
void orc_fft_shift(double *vec, size_t n, int inverse);

/* ---- resample/resample.go ------------------------------------------------ */
size_t orc_resample_time_length(size_t n, uint32_t source_rate, uint32_t target_rate);
void orc_resample_time(const double *samples, size_t n, uint32_t source_rate, uint32_t target_rate, double *out);
void orc_resample_frequency(const double *bins, size_t n_bins, uint32_t n_target, double *out);
void orc_resample_oversample(const double *source, size_t n_source, double *target, size_t n_target, uint32_t factor);

/* ---- random/random.go ---------------------------------------------------- */
typedef struct { uint64_t a, b, n, x; } orc_prng;
void orc_prng_seed(orc_prng *g, uint64_t seed);
double orc_prng_next_float(orc_prng *g);

/* ---- circular/circular.go ------------------------------------------------ */
typedef struct orc_circular orc_circular;
orc_circular *orc_circular_create(size_t size);
void orc_circular_destroy(orc_circular *c);
void orc_circular_enqueue(orc_circular *c, const double *elems, size_t n);
size_t orc_circular_length(const orc_circular *c);
int orc_circular_retrieve(const orc_circular *c, double *buf, size_t n);

/* ---- wave/wave.go sample codecs --------------------------------------------- */
void orc_wave_pcm16_decode(const int16_t *pcm, size_t n, double *out);  /* wave.go:400-428 */
void orc_wave_pcm16_encode(const double *samples, size_t n, int16_t *out); /* wave.go:347-379 */
/* formats: bytesToSamples / samplesToBytes (wave.go:737-838): little-endian payload bytes */
enum { ORC_FMT_PCM8 = 0, ORC_FMT_PCM16, ORC_FMT_PCM24, ORC_FMT_PCM32, ORC_FMT_IEEE32, ORC_FMT_IEEE64 };
size_t orc_wave_bytes_per_sample(int fmt);
int orc_wave_decode(int fmt, const unsigned char *data, size_t n_samples, double *out);   /* wave.go:316-342,400-428,483-531,589-617,662-689,714-732 */
int orc_wave_encode(int fmt, const double *samples, size_t n_samples, unsigned char *out); /* wave.go:275-311,347-395,433-478,536-584,622-657,694-709 */

/* ---- filter/filter.go ---------------------------------------------------- */
typedef struct orc_filter orc_filter;
orc_filter *orc_filter_from_coefficients(const double *coeffs, size_t n, uint32_t sample_rate); /* filter.go:850 */
void orc_filter_destroy(orc_filter *f);
size_t orc_filter_length(const orc_filter *f);
const double *orc_filter_coefficients(const orc_filter *f);
int orc_filter_process(orc_filter *f, const double *in, double *out, size_t n);   /* filter.go:342-515 */
orc_filter *orc_filter_reduce(const orc_filter *f, uint32_t order);               /* filter.go:520-604 */
orc_filter *orc_filter_normalize(const orc_filter *f);                            /* filter.go:328-336 */
orc_filter *orc_filter_multiply(const orc_filter *f, double scalar);              /* filter.go:270-323 */
void orc_filter_set_compensation(orc_filter *f, double gain_compensation);

typedef struct orc_irbank orc_irbank;
orc_irbank *orc_irbank_create(void);
void orc_irbank_destroy(orc_irbank *b);
/* filter.Import body for one descriptor (filter.go:738-780): resample to the 7 rates */
int orc_irbank_add(orc_irbank *b, const char *name, int32_t compensation_db,
                   const double *samples, size_t n, uint32_t source_rate);
/* one already-resampled response at one rate (test helper for synthetic IRs) */
int orc_irbank_add_raw(orc_irbank *b, const char *name, int32_t compensation_db,
                       const double *samples, size_t n, uint32_t rate);
size_t orc_irbank_name_count(const orc_irbank *b);
const char *orc_irbank_name(const orc_irbank *b, size_t i);

/* ---- oversampling/oversampling.go ---------------------------------------- */
typedef struct orc_osd orc_osd;
orc_osd *orc_osd_create(uint32_t factor);
void orc_osd_destroy(orc_osd *o);
int orc_osd_oversample(orc_osd *o, const double *in, size_t n_in, double *out, size_t n_out);
int orc_osd_decimate(orc_osd *o, const double *in, size_t n_in, double *out, size_t n_out);

/* ---- effects/ (all units) + signal/signal.go -------------------------------------- */
typedef struct orc_chain orc_chain;
orc_chain *orc_chain_create(orc_irbank *bank);                 /* signal.go:419 */
void orc_chain_destroy(orc_chain *c);
int orc_chain_append_unit(orc_chain *c, int unit_type);        /* id or -1 */
int orc_chain_remove_unit(orc_chain *c, int id);
int orc_chain_move_up(orc_chain *c, int id);
int orc_chain_move_down(orc_chain *c, int id);
int orc_chain_unit_type(orc_chain *c, int id);
int orc_chain_set_bypass(orc_chain *c, int id, int bypass);
int orc_chain_get_bypass(orc_chain *c, int id, int *bypass);
int orc_chain_set_numeric(orc_chain *c, int id, const char *name, int32_t value);
int orc_chain_get_numeric(orc_chain *c, int id, const char *name, int32_t *value);
int orc_chain_set_discrete(orc_chain *c, int id, const char *name, const char *value);
int orc_chain_get_discrete(orc_chain *c, int id, const char *name, const char **value);
int orc_chain_length(orc_chain *c);
void orc_chain_process(orc_chain *c, const double *in, double *out, size_t n, uint32_t sample_rate);
const char *orc_last_error(void);

/* ---- spatializer/spatializer.go ------------------------------------------ */
typedef struct orc_spatializer orc_spatializer;
orc_spatializer *orc_spatializer_create(uint32_t input_channels);
void orc_spatializer_destroy(orc_spatializer *s);
int orc_spatializer_set_azimuth(orc_spatializer *s, uint32_t ch, double azimuth);
int orc_spatializer_set_distance(orc_spatializer *s, uint32_t ch, double distance);
int orc_spatializer_set_level(orc_spatializer *s, uint32_t ch, double level);
void orc_spatializer_set_sample_rate(orc_spatializer *s, uint32_t rate);
/* in: n_in pointers to n samples; aux may be NULL; out: 2 pointers */
void orc_spatializer_process(orc_spatializer *s, const double *const *in, uint32_t n_in,
                             const double *aux, double *const *out, uint32_t n_out, size_t n);

/* ---- tuner/tuner.go -------------------------------------------------------- */
typedef struct orc_tuner orc_tuner;
typedef struct { int8_t cents; double frequency; const char *note; } orc_tuner_result;
orc_tuner *orc_tuner_create(void);
void orc_tuner_destroy(orc_tuner *t);
void orc_tuner_process(orc_tuner *t, const double *samples, size_t n, uint32_t sample_rate);
int orc_tuner_analyze(orc_tuner *t, orc_tuner_result *res);

/* ---- level/level.go ---------------------------------------------------------- */
typedef struct orc_meter orc_meter;
orc_meter *orc_meter_create(uint32_t channels);
void orc_meter_destroy(orc_meter *m);
uint32_t orc_meter_channel_count(const orc_meter *m);
int orc_meter_enabled(const orc_meter *m);
void orc_meter_set_enabled(orc_meter *m, int value);
int orc_meter_process(orc_meter *m, const double *const *bufs, const size_t *lens, uint32_t n_bufs, uint32_t sample_rate);
int orc_meter_analyze(const orc_meter *m, uint32_t channel, int32_t *level, int32_t *peak);
int orc_meter_state(const orc_meter *m, uint32_t channel, double *cur, double *peak, uint64_t *counter);

/* ---- controller.process mirror (controller/controller.go:2648-2783) -------
 * One block over C chains: tuner feed (channel tuner_channel or -1), chains,
 * spatializer (may be NULL) into out[C], out[C+1]; no metronome (zeros).      */
void orc_controller_process(orc_chain *const *chains, uint32_t n_ch, orc_tuner *tuner, int tuner_channel,
                            orc_spatializer *spat, const double *const *in, double *const *out,
                            uint32_t n_out, size_t n, uint32_t sample_rate);

/* ---- measurement driver (orc_bench.c): the CPU baseline of bench.py --------------------------------- */
double orc_bench_controller(orc_chain *const *chains, uint32_t n_ch, orc_spatializer *spat, const double *x, size_t n, size_t blocks,
                            uint32_t rate, int threads, int pin);

#ifdef __cplusplus
}
#endif
#endif
