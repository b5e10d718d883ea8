/*
 * gdg_host.cpp - host runtime of libgdg: the C ABI of include/gdg.h.
 *
 * Mirrors the reference's control surface for the hot path - effects.Unit
 * parameter store (effects/effects.go:96-384), signal.Chain slot list
 * (signal/signal.go), filter.ImpulseResponses + IR algebra
 * (filter/filter.go:127-336, 520-802), spatializer and tuner control state - and
 * turns it into POD descriptors + device state for the sm_100a kernels
 * (gdg_kernels.cu through gdg_backend.h).  No audio sample is ever computed
 * here: the only arithmetic on this side is parameter resolution (dB -> factor,
 * RC coefficients, delay lengths) and impulse-response compilation (control
 * path, runs when a power-amp parameter or the sample rate changes).
 */
#include "../../include/gdg.h"
#include "../../include/gdg_consts.h"
#include "../../include/gdg_tables.h"
#include "gdg_backend.h"
#include "gdg_device.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <vector>

namespace {

thread_local std::string g_error;
int fail(const char *fmt, ...)
{
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_error = buf;
    return -1;
}
int backend_fail(const char *what) { return fail("%s: %s", what, gdgb_error()); }

/* The reference guards every unit's parameters and every chain's slot list with an RWMutex (effects/effects.go:98,
 * signal/signal.go:45): web-UI goroutines edit chains while the audio callback runs.  Here ONE reader/writer lock plays
 * both roles: every processing entry point holds it SHARED from "resolve descriptors" to "kernels enqueued", every
 * mutator (append / remove / move / bypass / set_*) and every getter that walks the parameter vectors holds it
 * EXCLUSIVE resp. SHARED.  A mutator therefore never sees a half-prepared block, RemoveUnit cannot free device state that
 * a not-yet-enqueued launch still references (it synchronises the stream after taking the lock), and parameter changes
 * take effect at the next block boundary, as in the reference. */
std::shared_mutex g_edit_mutex;
typedef std::unique_lock<std::shared_mutex> EditLock;
typedef std::shared_lock<std::shared_mutex> ProcessLock;

/* One process may drive several GPUs (gdg_use_device / gdg_cluster_*): every object lives on the device context that was
 * selected on the creating thread and every entry point works under its object's context. */
struct Use {
    int prev;
    explicit Use(int ctx) : prev(gdgb_current()) { if (ctx >= 0 && ctx != prev) gdgb_select(ctx); else gdgb_bind(); }
    ~Use() { if (gdgb_current() != prev) gdgb_select(prev); }
    Use(const Use &) = delete;
};

std::once_flag g_init_once;
int g_init_status = -1;
std::string g_init_error;

/* ------------------------------------------------------------------------ */
/* small host-side numerics used by the control path                         */
/* ------------------------------------------------------------------------ */
typedef std::complex<double> cplx;

uint64_t next_pow2(uint64_t v, uint32_t *exp_out = nullptr)     /* fft/fft.go:393-409 */
{
    uint64_t p = 1; uint32_t e = 0;
    while (p < v) { p <<= 1; e++; }
    if (exp_out) *exp_out = e;
    return p;
}

/* resample/resample.go:10-31, 36-66 */
double lanczos_kernel(double x, double a)
{
    if (x == 0) return 1.0;
    if (-a < x && x < a) {
        const double pix = GDG_MATH_PI * x;
        return (a * (std::sin(pix) * std::sin(pix / a))) / (pix * pix);
    }
    return 0.0;
}
/* output length of resample.Time (resample/resample.go:72-89) */
size_t resample_time_length(size_t n, uint32_t src, uint32_t dst)
{
    if (!src) return 0;
    const double lf = (double)n * ((double)dst / (double)src);
    const double fl = std::floor(lf);
    long len = (long)fl;
    if (fl == lf) len--;
    return len < 0 ? 0 : (size_t)len;
}

double db2fac(int32_t db) { return std::pow(10.0, 0.05 * (double)db); }     /* effects.go:389-394 */

/* ------------------------------------------------------------------------ */
/* impulse responses + IR algebra (filter/filter.go)                          */
/* ------------------------------------------------------------------------ */
const uint32_t k_sample_rates[7] = {22050, 32000, 44100, 48000, 88200, 96000, 192000};   /* filter.go:25-33 */

/* one response at one rate (filter.impulseResponseStruct): the coefficients live on the device, where filter.Import's
 * resampling produced them and where the power amp compiles them */
struct ImpulseResponse {
    std::string name; uint32_t rate; double comp; double *d_data = nullptr; size_t n = 0;
    ImpulseResponse(std::string nm, uint32_t r, double c, double *d, size_t len) : name(std::move(nm)), rate(r), comp(c), d_data(d), n(len) {}
    ImpulseResponse(const ImpulseResponse &) = delete;
    ImpulseResponse(ImpulseResponse &&o) noexcept : name(std::move(o.name)), rate(o.rate), comp(o.comp), d_data(o.d_data), n(o.n) { o.d_data = nullptr; }
    ~ImpulseResponse() { if (d_data) gdgb_free(d_data); }
};

/* a filter's coefficients on the device (the coefficient side of filter.filterStruct) */
struct DevFilter {
    double *d = nullptr; size_t n = 0; bool owned = false;
    DevFilter() {}
    DevFilter(const DevFilter &) = delete;
    DevFilter &operator=(const DevFilter &) = delete;
    void reset() { if (owned && d) gdgb_free(d); d = nullptr; n = 0; owned = false; }
    ~DevFilter() { reset(); }
};

} // namespace

struct gdg_irbank {
    int ctx = -1;                                          /* device context of the responses (set by the first add) */
    std::vector<ImpulseResponse> responses;
    uint64_t generation = 0;                               /* bumped by every add: compiled-filter cache key */
    std::vector<std::string> names() const                 /* filter.go:681-699 */
    {
        std::vector<std::string> out;
        for (auto &r : responses) if (std::find(out.begin(), out.end(), r.name) == out.end()) out.push_back(r.name);
        return out;
    }
    const ImpulseResponse *find(const std::string &name, uint32_t rate) const   /* CreateFilter filter.go:619-660 */
    {
        for (auto &r : responses) if (r.name == name && r.rate == rate) return &r;
        return nullptr;
    }
    std::vector<std::string> name_cache; uint64_t name_cache_generation = ~0ull;    /* gdg_irbank_name's strings: stable until the next add */
};

namespace {

/* ------------------------------------------------------------------------ */
/* compiled power-amp filters: device spectrum shared between chains           */
/* ------------------------------------------------------------------------ */
struct DeviceSpectrum { double *d_H = nullptr; int L = 0, F = 0, logM = 0, K = 1; ~DeviceSpectrum() { if (d_H) gdgb_free(d_H); } };
/* a compiled power-amp filter (the composite of poweramp.compile, poweramp.go:25-127): coefficients on the device, shared by
 * every chain whose recipe (bank, rate, target order, filter names and levels) is the same */
struct CompiledFilter { double *d = nullptr; int L = 0; std::string key; ~CompiledFilter() { if (d) gdgb_free(d); } };
std::mutex g_spec_mutex;
std::map<std::string, std::weak_ptr<DeviceSpectrum>> g_spec_cache;
std::map<std::string, std::weak_ptr<CompiledFilter>> g_filter_cache;

/* partition length of the convolution for a filter of L taps processed in blocks of n samples: filters of up to 8192 taps run
 * as ONE partition of nextPow2(L) taps whatever the block length (the reference's own block size, filter.go:372-388);
 * longer ones as a uniformly partitioned convolution whose partition = hop = block length (at most 8192), so that the
 * reference's default filter_order (no reduction: up to 19199 taps at 192 kHz) also works at JACK period sizes. */
uint64_t partition_length(size_t L, size_t n)
{
    const uint64_t B = next_pow2((uint64_t)L);
    if (2 * B <= GDG_MAX_FFT) return B;
    return std::min<uint64_t>((uint64_t)n, GDG_MAX_FFT / 2);
}
/* H of filter.Process (filter.go:395-401) for every partition, built on the device by the chain kernel's own FFT */
std::shared_ptr<DeviceSpectrum> make_spectrum(const CompiledFilter &flt, uint64_t B)
{
    const int L = flt.L;
    uint32_t e = 0;
    next_pow2(B, &e);
    const int K = std::max(1, (int)((L + (int)B - 1) / (int)B));
    const int F = (int)(2 * B);
    std::string key = flt.key;
    key.append("|B="); key.append(std::to_string(B));
    std::lock_guard<std::mutex> lk(g_spec_mutex);
    auto it = g_spec_cache.find(key);
    if (it != g_spec_cache.end()) if (auto sp = it->second.lock()) return sp;
    auto sp = std::make_shared<DeviceSpectrum>();
    sp->L = L; sp->F = F; sp->logM = (int)e; sp->K = K;     /* M = F/2 = B = 2^e */
    const size_t bins = (size_t)F / 2 + 1;
    sp->d_H = (double *)gdgb_alloc(2 * bins * (size_t)K * sizeof(double));
    if (!sp->d_H || gdgb_ir_partition_spectra(flt.d, L, (int)B, (int)e, K, sp->d_H, gdgb_stream())) return nullptr;
    g_spec_cache[key] = sp;
    return sp;
}

/* ------------------------------------------------------------------------ */
/* effects.Unit                                                               */
/* ------------------------------------------------------------------------ */
struct Param {
    std::string name; int32_t type; std::string phys;
    int32_t min, max, num; int32_t didx; std::vector<std::string> disc;
};

std::vector<std::string> split_bar(const char *s)
{
    std::vector<std::string> out; std::string cur;
    for (const char *c = s; *c; c++) { if (*c == '|') { out.push_back(cur); cur.clear(); } else cur.push_back(*c); }
    out.push_back(cur);
    return out;
}

struct Unit {
    int type;
    std::vector<Param> params;
    gdg_irbank *bank = nullptr;
    /* device state */
    double *d_st = nullptr; size_t st_words = 0; long st_key = -1;
    size_t os_last_n[2] = {0, 0};                          /* block length last seen by the x2 / x4 oversampler */
    /* cached descriptor */
    gdg_unit_desc desc; bool dirty = true; uint32_t cached_rate = 0; bool valid = false;
    /* power amp */
    uint32_t pa_rate = 0; bool pa_has_filter = false; std::shared_ptr<CompiledFilter> pa_filter; bool pa_new_filter = false;
    std::shared_ptr<DeviceSpectrum> pa_spec;

    explicit Unit(int t) : type(t) { memset(&desc, 0, sizeof desc); }
    ~Unit() { if (d_st) gdgb_free(d_st); }

    void add_num(const char *n, const char *ph, int32_t mn, int32_t mx, int32_t def)
    { params.push_back(Param{n, GDG_PARAMETER_TYPE_NUMERIC, ph, mn, mx, def, -1, {}}); }
    void add_disc(const char *n, int didx, std::vector<std::string> vals)
    { params.push_back(Param{n, GDG_PARAMETER_TYPE_DISCRETE, "", -1, -1, -1, didx, std::move(vals)}); }

    int find(const char *name) const     /* last match wins, like the reference's linear search (effects.go:130-140) */
    {
        int idx = -1;
        for (size_t i = 0; i < params.size(); i++) if (params[i].name == name) idx = (int)i;
        return idx;
    }
    int32_t num(const char *name) const { int i = find(name); return (i >= 0 && params[i].type == GDG_PARAMETER_TYPE_NUMERIC) ? params[i].num : 0; }
    std::string disc(const char *name) const
    { int i = find(name); return (i >= 0 && params[i].type == GDG_PARAMETER_TYPE_DISCRETE) ? params[i].disc[params[i].didx] : std::string(); }
    int disc_index(const char *name) const { int i = find(name); return i >= 0 ? params[i].didx : -1; }

    /* (re)allocate zeroed state when its shape key changes (the reference re-makes its slices the same way) */
    int ensure_state(size_t words, long key, size_t keep_words = 0)
    {
        if (d_st && st_key == key && st_words == words) return 0;
        double *fresh = (double *)gdgb_alloc(words * sizeof(double));
        if (!fresh) return backend_fail("device allocation for unit state failed");
        /* scalar fields of the Go struct (previousPhase, envelope) survive when only the history slice is re-made */
        if (d_st && keep_words && gdgb_copy(fresh, d_st, std::min(keep_words, std::min(words, st_words)) * sizeof(double)))
            return backend_fail("state copy failed");
        if (d_st) { gdgb_sync(); gdgb_free(d_st); }
        d_st = fresh; st_words = words; st_key = key;
        return 0;
    }
    int follow_mode() const { std::string f = disc("follow"); return f == "envelope" ? GDG_FOLLOW_ENVELOPE : (f == "level" ? GDG_FOLLOW_LEVEL : GDG_FOLLOW_OTHER); }
    int os_factor() const { std::string o = disc("oversampling"); return o == "2" ? 2 : (o == "4" ? 4 : 1); }

    int compile_poweramp(uint32_t rate);
    int prepare(uint32_t rate, size_t n);
};

/* effects/poweramp.go:25-127 compile - on the device: per selected response
 *   Reduce(filter_order)  filter.go:520-604   real FFT -> Lanczos resampling of the positive bins -> inverse -> truncate
 *   Normalize             filter.go:328-336   scale to compensation / ||h||_2
 *   Multiply(10^(level/20)), Add into the composite   filter.go:270-323, 167-253
 * enqueued on the library stream; identical recipes are compiled once and shared. */
int Unit::compile_poweramp(uint32_t rate)
{
    if (!bank) return -1;                                   /* "No impulse responses were loaded." */
    const uint32_t target = (uint32_t)strtoul(disc("filter_order").c_str(), nullptr, 10);
    struct Term { const ImpulseResponse *ir; double fac; size_t len; };
    std::vector<Term> terms;
    char head[96];
    snprintf(head, sizeof head, "bank=%p/%llu|rate=%u|order=%u|", (void *)bank, (unsigned long long)bank->generation, rate, target);
    std::string key(head);
    size_t L = 0;
    for (int i = 0; i < 8; i++) {
        char pf[32], pl[32];
        snprintf(pf, sizeof pf, "filter_%d", i + 1);
        snprintf(pl, sizeof pl, "level_%d", i + 1);
        const std::string name = disc(pf);
        if (name != "- NONE -") {
            const ImpulseResponse *ir = bank->find(name, rate);
            if (!ir) return -1;                             /* "Failed to load filter '%s' for sample rate '%d'." */
            const size_t len = (target > 0 && ir->n > (size_t)target) ? (size_t)target : ir->n;     /* Reduce is a no-op for n <= order */
            terms.push_back(Term{ir, db2fac(num(pl)), len});
            L = std::max(L, len);
            key.append(name); key.push_back('\x1f'); key.append(std::to_string(num(pl))); key.push_back(';');
        }
    }
    std::lock_guard<std::mutex> lk(g_spec_mutex);
    auto it = g_filter_cache.find(key);
    if (it != g_filter_cache.end()) if (auto sp = it->second.lock()) { pa_filter = sp; pa_has_filter = true; pa_new_filter = true; return 0; }
    auto sp = std::make_shared<CompiledFilter>();
    sp->L = (int)L; sp->key = key;
    sp->d = (double *)gdgb_alloc(std::max<size_t>(1, L) * sizeof(double));     /* filter.Empty: zeros */
    double *d_ss = (double *)gdgb_alloc(sizeof(double));
    if (!sp->d || !d_ss) { gdgb_free(d_ss); return backend_fail("filter allocation failed"); }
    std::vector<double *> temps{d_ss};
    void *st = gdgb_stream();
    int rc = 0;
    for (auto &t : terms) {
        const double *cur = t.ir->d_data;
        if (t.len != t.ir->n) {                             /* filter.Reduce */
            const int n_src = (int)next_pow2(t.ir->n), n_tgt = (int)next_pow2(target);
            const int pos_src = n_src / 2 + 1, pos_tgt = n_tgt / 2 + 1;
            double *work = (double *)gdgb_alloc((size_t)std::max(n_src, n_tgt) * sizeof(double));
            double *spec = (double *)gdgb_alloc(2 * (size_t)pos_src * sizeof(double));
            double *spec2 = (double *)gdgb_alloc(2 * (size_t)pos_tgt * sizeof(double));
            double *red = (double *)gdgb_alloc((size_t)target * sizeof(double));
            temps.push_back(work); temps.push_back(spec); temps.push_back(spec2); temps.push_back(red);
            if (!work || !spec || !spec2 || !red) { rc = backend_fail("filter allocation failed"); break; }
            if (gdgb_ir_real_spectrum(cur, (long)t.ir->n, n_src, work, spec, st) || gdgb_ir_resample_bins(spec, pos_src, spec2, pos_tgt, st) ||
                gdgb_ir_real_inverse(spec2, n_tgt, work, red, (long)target, st)) { rc = backend_fail("filter reduction failed"); break; }
            cur = red;
        }
        if (gdgb_ir_sum_squares(cur, (long)t.len, d_ss, st) || gdgb_ir_scale_add(sp->d, cur, (long)t.len, t.ir->comp, t.fac, d_ss, st)) { rc = backend_fail("filter compilation failed"); break; }
    }
    gdgb_sync();                                            /* the scratch buffers go away */
    for (double *p : temps) gdgb_free(p);
    if (rc) return -1;
    g_filter_cache[key] = sp;
    pa_filter = sp; pa_has_filter = true; pa_new_filter = true;
    return 0;
}

/* Resolve parameters -> descriptor (what each Process body recomputes per block in the reference). */
int Unit::prepare(uint32_t rate, size_t n)
{
    if (type == GDG_UNIT_POWERAMP && rate != pa_rate) {     /* poweramp.go:187-199 */
        pa_rate = rate;
        compile_poweramp(rate);
        dirty = true;
    }
    if (type == GDG_UNIT_POWERAMP && pa_spec && pa_has_filter && !pa_new_filter && pa_spec->F != (int)(2 * partition_length((size_t)pa_filter->L, n))) {
        /* a filter longer than 8192 taps is partitioned by the block length: a new block length re-partitions it.  DELIBERATE
         * DIFFERENCE: the overlap tail restarts here, the reference's single big FFT keeps it across block-length changes. */
        pa_new_filter = true; dirty = true;
    }
    if (type == GDG_UNIT_EXCESS || type == GDG_UNIT_FUZZ || type == GDG_UNIT_OVERDRIVE || type == GDG_UNIT_DISTORTION) {
        /* oversampling.go:74-84: the up-sampler's input buffer (N + 8, its first 8 cells = the last 8 inputs) is re-made,
         * i.e. zeroed, whenever the block length differs from the one it was made for; the anti-aliasing filter's
         * overlap tail does not depend on N and survives.  Each unit owns one x2 and one x4 oversampler
         * (e.g. overdrive.go:150-151), each with its own history. */
        const int f = os_factor();
        if (f > 1) {
            size_t &last = os_last_n[f == 2 ? 0 : 1];
            if (last && last != n && d_st && gdgb_zero(d_st + (f == 2 ? GDG_OS_ST_UP2 : GDG_OS_ST_UP4), 8 * sizeof(double)))
                return backend_fail("state reset failed");
            last = n;
        }
    }
    if (valid && !dirty && cached_rate == rate) return 0;
    gdg_unit_desc d; memset(&d, 0, sizeof d);
    d.type = type;
    const double sr = (double)rate;
    const double m2pi = -GDG_MATH_TWO_PI / sr;
    switch (type) {
    case GDG_UNIT_SIGNALGENERATOR: {
        const std::string t = disc("signal_type");
        d.i[0] = t == "sine" ? 0 : t == "triangle" ? 1 : t == "square" ? 2 : t == "sawtooth" ? 3 : 4;
        d.p[0] = (0.01 * (double)num("input_amplitude")) * db2fac(num("input_gain"));
        d.p[1] = (0.01 * (double)num("signal_amplitude")) * db2fac(num("signal_gain"));
        d.p[2] = GDG_MATH_TWO_PI * ((double)num("signal_frequency") / sr);
        const bool fresh = !d_st;
        if (ensure_state(2, 0)) return -1;
        if (fresh) {                                       /* random.CreatePRNG(1337): random.go:40-55 */
            const uint64_t x = ((64979ull * 1337ull) + 83ull) % 2147483647ull;
            double w[2] = {0.0, 0.0}; memcpy(&w[1], &x, 8);
            if (gdgb_upload(d_st, w, sizeof w)) return backend_fail("upload");
        }
        break; }
    case GDG_UNIT_NOISEGATE: {
        const int32_t lo = num("threshold_open"), lc = num("threshold_close");
        d.i[0] = lo < lc;
        d.i[1] = (int32_t)(uint32_t)std::floor(((0.001 * (double)num("hold_time")) * sr) + 0.5);
        d.p[0] = db2fac(lo); d.p[1] = db2fac(lc);
        if (ensure_state(2, 0)) return -1;
        break; }
    case GDG_UNIT_BANDPASS: {
        const int half = (int)(strtoul(disc("filter_order").c_str(), nullptr, 10) >> 1);
        int32_t fa = num("frequency_1"), fb = num("frequency_2");
        if (fa > fb) std::swap(fa, fb);
        d.i[0] = half;
        d.p[0] = 1.0 - std::exp(m2pi * (double)fa);
        d.p[1] = 1.0 - std::exp(m2pi * (double)fb);
        if (ensure_state(8, half)) return -1;              /* bandpass.go:47-54: re-made when the order changes */
        break; }
    case GDG_UNIT_AUTOWAH: {
        int32_t la = num("level_1"), lb = num("level_2"), fa = num("frequency_1"), fb = num("frequency_2");
        if (la > lb) { std::swap(la, lb); std::swap(fa, fb); }
        d.i[0] = follow_mode();
        d.p[0] = std::exp(-20.0 / sr); d.p[1] = 1.0 - d.p[0];
        d.p[2] = la; d.p[3] = lb; d.p[4] = fa; d.p[5] = fb;
        d.p[6] = ((double)fb - (double)fa) / ((double)lb - (double)la);
        d.p[7] = sr; d.p[8] = std::pow(2.0, 8.0);
        if (ensure_state(17, 0)) return -1;
        break; }
    case GDG_UNIT_AUTOYOY: {
        int32_t la = num("level_1"), lb = num("level_2");
        double da = 0.0, dbb = 0.01 * (double)num("depth");
        if (la > lb) { std::swap(la, lb); std::swap(da, dbb); }
        const double sr_inv = 1.0 / sr;
        const int hs = (int)std::floor((0.01 * sr) + 0.5);
        d.i[0] = hs; d.i[1] = follow_mode();
        d.p[0] = std::exp(-20.0 * sr_inv); d.p[1] = 1.0 - d.p[0];
        d.p[2] = la; d.p[3] = lb; d.p[4] = da; d.p[5] = dbb;
        d.p[6] = (dbb - da) / ((double)lb - (double)la);
        d.p[7] = sr;
        if (ensure_state(2 + (size_t)hs, hs, 1)) return -1;
        break; }
    case GDG_UNIT_COMPRESSOR:
        d.i[0] = follow_mode();
        d.p[0] = std::exp(-20.0 / sr); d.p[1] = 1.0 - d.p[0];
        d.p[2] = db2fac(num("target_level")); d.p[3] = db2fac(num("gain_limit"));
        if (ensure_state(1, 0)) return -1;
        break;
    case GDG_UNIT_OCTAVER:
        d.i[0] = follow_mode();
        d.p[0] = std::exp(-20.0 / sr); d.p[1] = 1.0 - d.p[0];
        d.p[2] = db2fac(num("level_octave_up")); d.p[3] = db2fac(num("level_clean")); d.p[4] = db2fac(num("level_dist"));
        d.p[5] = db2fac(num("level_octave_down_first")); d.p[6] = db2fac(num("level_octave_down_second"));
        d.p[7] = db2fac(num("level_hysteresis"));
        if (ensure_state(4, 0)) return -1;
        break;
    case GDG_UNIT_EXCESS: case GDG_UNIT_FUZZ: case GDG_UNIT_OVERDRIVE: case GDG_UNIT_DISTORTION: {
        const int f = os_factor();
        d.i[0] = f;
        d.p[1] = db2fac(num("level"));
        if (type == GDG_UNIT_EXCESS) d.p[0] = db2fac(num("gain"));
        else d.p[0] = db2fac(num("boost") + num("gain"));
        if (type == GDG_UNIT_FUZZ) {
            d.i[1] = follow_mode();
            d.p[2] = 0.01 * (double)num("bias");
            d.p[3] = 0.01 * (double)num("fuzz"); d.p[4] = 1.0 - d.p[3];
            const double sr_os = (double)((uint32_t)f * rate);           /* fuzz.go:162-163 */
            d.p[5] = std::exp(-20.0 / sr_os); d.p[6] = 1.0 - d.p[5];
        }
        if (type == GDG_UNIT_OVERDRIVE) {
            const std::string v = disc("valve");
            d.i[2] = v == "ECC82 (12AU7)" ? 0 : (v == "ECC83 (12AX7)" ? 1 : -1);
            d.p[2] = 0.01 * (double)num("drive"); d.p[3] = 1.0 - d.p[2];
        }
        if (ensure_state(GDG_OS_ST_WORDS, 0)) return -1;
        break; }
    case GDG_UNIT_TONESTACK: {
        static const double freqs[5] = {20.0, 300.0, 3000.0, 6000.0, 20000.0};
        static const char *names[4] = {"low", "middle", "presence", "high"};
        for (int j = 0; j < 4; j++) {
            d.p[j] = 1.0 - std::exp(m2pi * freqs[j]);
            d.p[4 + j] = 1.0 - std::exp(m2pi * freqs[j + 1]);
            d.p[8 + j] = db2fac(num(names[j]));
        }
        if (ensure_state(8, 0)) return -1;
        break; }
    case GDG_UNIT_CHORUS: {
        double depth = 0.1 * (double)num("depth");
        depth = depth < 0.0 ? 0.0 : (depth > 10.0 ? 10.0 : depth);
        const int hs = (int)std::floor((0.05 * sr) + 0.5);
        d.i[0] = hs; d.p[0] = depth; d.p[1] = GDG_MATH_PI_THOUSANDTH * (double)num("speed"); d.p[2] = sr;
        for (int j = 1; j < 5; j++) { d.p[2 + j] = std::cos(GDG_MATH_TWO_PI_FIFTH * (double)j); d.p[6 + j] = std::sin(GDG_MATH_TWO_PI_FIFTH * (double)j); }
        if (ensure_state(2 + (size_t)hs, hs, 1)) return -1;
        break; }
    case GDG_UNIT_FLANGER: case GDG_UNIT_PHASER: {
        double depth = 0.01 * (double)num("depth");
        depth = depth < 0.0 ? 0.0 : (depth > 1.0 ? 1.0 : depth);
        const int hs = (int)std::floor((0.002 * sr) + 0.5);
        d.i[0] = hs; d.p[0] = depth; d.p[1] = GDG_MATH_TWO_PI_HUNDREDTH * (double)num("speed");
        d.p[2] = sr; d.p[3] = 1.0 / sr; d.p[4] = 0.5; d.p[5] = 0.5;
        if (type == GDG_UNIT_PHASER) {
            d.p[5] = 0.5 * std::sin(GDG_MATH_DEGREE_TO_RADIANS * (double)num("phase"));
            d.p[4] = 1.0 - std::fabs(d.p[5]);
        }
        if (ensure_state(2 + (size_t)hs, hs, 1)) return -1;
        break; }
    case GDG_UNIT_TREMOLO: {
        const double period_f = sr / (0.1 * (double)num("frequency"));
        const uint32_t period = (uint32_t)period_f;
        const uint32_t un = (uint32_t)(period_f * (0.01 * (double)num("phase")));
        d.i[0] = (int32_t)(period - un); d.i[1] = (int32_t)un;
        d.p[0] = db2fac(num("depth"));
        if (ensure_state(2, 0)) return -1;
        break; }
    case GDG_UNIT_RINGMODULATOR:
        d.p[0] = (GDG_MATH_TWO_PI * (double)num("frequency")) / sr;
        if (ensure_state(1, 0)) return -1;
        break;
    case GDG_UNIT_DELAY: {
        const int ds = (int)std::floor(((0.001 * (double)num("delay_time")) * sr) + 0.5);
        d.i[0] = ds; d.p[0] = db2fac(num("feedback")); d.p[1] = db2fac(num("level"));
        if (ensure_state(2 + (size_t)ds, ds)) return -1;
        break; }
    case GDG_UNIT_REVERB: {
        static const double ap[3] = {0.04204, 0.01348, 0.00452};
        static const double tt[4] = {0.19196, 0.19996, 0.21596, 0.23204};
        static const double tc[4] = {0.1855, 0.18325, 0.17875, 0.17425};
        const double wet = 0.01 * (double)num("mix");
        d.p[0] = 1.0 - wet; d.p[1] = 0.5 * wet; d.p[6] = 0.7;
        int hs = 0; size_t total = GDG_RV_ST_DATA;
        for (int j = 0; j < 4; j++) { d.i[1 + j] = (int32_t)(uint32_t)std::round(tt[j] * sr); d.p[2 + j] = tc[j]; hs = std::max(hs, d.i[1 + j]); }
        d.i[0] = hs; total += (size_t)hs; d.i[8] = (int32_t)total;
        for (int a = 0; a < 3; a++) { d.i[5 + a] = (int)std::round(ap[a] * sr); total += (size_t)std::max(0, d.i[5 + a] - 1); }
        if (ensure_state(total, (long)rate)) return -1;   /* reverb.go:199-262: rebuilt when the rate changes */
        break; }
    case GDG_UNIT_POWERAMP: {
        if (!pa_has_filter) { d.i[0] = -1; break; }        /* poweramp.go:210-214: zeros */
        const int L = pa_filter->L;
        d.i[0] = L;
        if (L > 0) {
            if (pa_new_filter || !pa_spec) {
                pa_spec = make_spectrum(*pa_filter, partition_length((size_t)pa_filter->L, n));
                if (!pa_spec) return backend_fail("filter spectrum upload failed");
                /* a recompiled filter is a NEW filter object in the reference: fresh (zero) overlap tail */
                if (d_st) { gdgb_sync(); gdgb_free(d_st); d_st = nullptr; }
                const size_t words = (size_t)pa_spec->F + 2 + (pa_spec->K > 1 ? (size_t)pa_spec->K * (pa_spec->F + 2) : 0);
                if (ensure_state(words, pa_spec->F)) return -1;   /* tail[F], zero-run word, slot word, spectrum delay line */
                pa_new_filter = false;
            }
            d.i[3] = pa_spec->K; d.i[4] = pa_spec->F + 2;
            d.i[1] = pa_spec->F; d.i[2] = pa_spec->logM; d.aux = pa_spec->d_H;
        }
        break; }
    case GDG_UNIT_CABINET: {
        static const double f[7] = {300.0, 120.0, 80.0, 3000.0, 4000.0, 5000.0, 6000.0};
        for (int j = 0; j < 7; j++) d.p[j] = 1.0 - std::exp(m2pi * f[j]);
        if (ensure_state(7, 0)) return -1;
        break; }
    default: return fail("unknown unit type %d", type);
    }
    d.st = d_st;
    desc = d; dirty = false; cached_rate = rate; valid = true;
    return 1;       /* descriptor changed */
}

Unit *create_unit(int type, gdg_irbank *bank)
{
    if (type < 0 || type >= GDG_UNIT_COUNT) return nullptr;
    Unit *u = new Unit(type);
#define GDG_UNIT(code, ident, name)
#define GDG_NUM(ident, name, phys, mn, mx, def) if (type == GDG_UNIT_##ident) u->add_num(name, phys, mn, mx, def);
#define GDG_DISC(ident, name, didx, vals) if (type == GDG_UNIT_##ident) u->add_disc(name, didx, split_bar(vals));
#include "../../include/gdg_unit_params.def"
#undef GDG_UNIT
#undef GDG_NUM
#undef GDG_DISC
    if (type == GDG_UNIT_POWERAMP && bank) {               /* effects.PreparePowerAmp: poweramp.go:221-289 */
        std::vector<std::string> ext{"- NONE -"};
        for (auto &n : bank->names()) ext.push_back(n);
        for (int i = 0; i < 8; i++) {
            char nm[32];
            snprintf(nm, sizeof nm, "filter_%d", i + 1); u->add_disc(nm, 0, ext);
            snprintf(nm, sizeof nm, "level_%d", i + 1); u->add_num(nm, "dB", -60, 0, 0);
        }
        u->bank = bank;
    }
    return u;
}

const char *k_unit_names[GDG_UNIT_COUNT] = {
#define GDG_UNIT(code, ident, name) name,
#define GDG_NUM(...)
#define GDG_DISC(...)
#include "../../include/gdg_unit_params.def"
#undef GDG_UNIT
#undef GDG_NUM
#undef GDG_DISC
};

int ensure_init()
{
    std::call_once(g_init_once, [] {
        int dev = -1;
        if (const char *e = getenv("GDG_DEVICE")) dev = atoi(e);
        g_init_status = gdg_init(dev);
    });
    if (g_init_status) { g_error = g_init_error; return -1; }
    if (gdgb_bind()) return backend_fail("cudaSetDevice");      /* this OS thread may never have seen the device (cgo, UI threads) */
    return 0;
}

int check_block(size_t n)
{
    if (n == 0 || n > GDG_MAX_N || (n & (n - 1)))
        return fail("block length %zu is not supported: must be a power of two <= %d", n, GDG_MAX_N);
    return 0;
}

} // namespace

/* ---- signal.Chain ----------------------------------------------------------- */
struct Slot { std::unique_ptr<Unit> unit; bool bypass; };
/* The active units of a chain are cut into SEGMENTS, each one kernel launch ("round") per block:
 *   kind 0  block-resident chain kernel (k_chain): up to GDG_MAX_SLOTS units of any type;
 *   kind 1  oversampled-run kernel (k_os): a maximal run of consecutive excess / fuzz / overdrive / distortion units
 *           with oversampling "2" or "4" (up to GDG_OSK_MAX_UNITS per launch), streamed tile by tile at four CTAs per SM. */
enum { SEG_CHAIN = 0, SEG_OS = 1, SEG_KINDS = 2 };
#define GDG_OSK_MAX_UNITS_HOST 6                /* == GDG_OSK_MAX_UNITS (gdg_os.cuh) */
struct gdg_chain {
    int ctx = 0;                                           /* device context the chain (descriptors, unit state) lives on */
    gdg_irbank *bank;
    std::vector<Slot> slots;
    /* device descriptors: one per segment */
    std::vector<gdg_chain_desc *> d_desc; std::vector<gdg_chain_desc> h_desc; std::vector<int> seg_kind;
    bool topo_dirty = true;
    /* bumped whenever prepare() re-made the segment descriptors: every launch-table cache (the context's batch, the chain's own
     * one-chain batch) remembers the version it was built from, so a chain that is driven through gdg_chain_process AND through
     * the batch entry points can never leave one of them with stale tables */
    uint64_t version = 0;
    /* nothing was edited since the last prepare() at this rate and block length: the per-unit parameter resolution (string
     * look-ups) is skipped - with thousands of chains per block that walk is what bounds the host */
    bool clean = false; uint32_t clean_rate = 0; size_t clean_n = 0;
    double *d_in = nullptr, *d_out = nullptr; size_t d_n = 0;   /* single-channel staging for gdg_chain_process */
    struct Batch *self = nullptr;                               /* one-chain launch tables of gdg_chain_process */
    ~gdg_chain();

    static bool is_os(const Unit *u)
    {
        return (u->type == GDG_UNIT_EXCESS || u->type == GDG_UNIT_FUZZ || u->type == GDG_UNIT_OVERDRIVE || u->type == GDG_UNIT_DISTORTION) && u->os_factor() > 1;
    }
    /* bring the device descriptors up to date; returns number of segments or -1 */
    int prepare(uint32_t rate, size_t n)
    {
        if (clean && clean_rate == rate && clean_n == n && !h_desc.empty()) return (int)h_desc.size();
        const int r_ = prepare_units(rate, n);
        if (r_ > 0) { clean = true; clean_rate = rate; clean_n = n; }
        return r_;
    }
    int prepare_units(uint32_t rate, size_t n)
    {
        bool changed = topo_dirty;
        std::vector<Unit *> active;
        for (auto &s : slots) if (!s.bypass) {
            int r = s.unit->prepare(rate, n);
            if (r < 0) return -1;
            if (r > 0) changed = true;
            active.push_back(s.unit.get());
        }
        if (!changed && !h_desc.empty()) return (int)h_desc.size();
        static const bool split_os = !getenv("GDG_NO_OS_KERNEL");     /* development knob: everything through k_chain */
        std::vector<gdg_chain_desc> segs; std::vector<int> kinds;
        for (Unit *u : active) {
            const int kind = (split_os && is_os(u)) ? SEG_OS : SEG_CHAIN;
            const int cap = kind == SEG_OS ? GDG_OSK_MAX_UNITS_HOST : GDG_MAX_SLOTS;
            if (segs.empty() || kinds.back() != kind || segs.back().nslots >= cap) {
                gdg_chain_desc cd; memset(&cd, 0, sizeof cd);
                segs.push_back(cd); kinds.push_back(kind);
            }
            segs.back().slot[segs.back().nslots++] = u->desc;
        }
        if (segs.empty()) { gdg_chain_desc cd; memset(&cd, 0, sizeof cd); segs.push_back(cd); kinds.push_back(SEG_CHAIN); }   /* pass-through */
        version++;
        h_desc = segs; seg_kind = kinds;
        const size_t nseg = h_desc.size();
        while (d_desc.size() < nseg) {
            gdg_chain_desc *p = (gdg_chain_desc *)gdgb_alloc(sizeof(gdg_chain_desc));
            if (!p) return backend_fail("descriptor allocation failed");
            d_desc.push_back(p);
        }
        for (size_t g = 0; g < nseg; g++)
            if (gdgb_upload(d_desc[g], &h_desc[g], sizeof(gdg_chain_desc))) return backend_fail("descriptor upload failed");
        topo_dirty = false;
        return (int)nseg;
    }
};

namespace {

/* a batch = ordered list of chains; caches the device launch tables: per segment round and kernel kind the array of
 * descriptor pointers (NULL = this channel has no segment of that kind in that round), and the segment count per channel */
} // namespace
struct Batch {
    std::vector<gdg_chain *> chains;
    std::vector<std::vector<const gdg_chain_desc *>> h_ptrs;     /* [round * SEG_KINDS + kind][channel] */
    std::vector<const gdg_chain_desc **> d_ptrs;
    std::vector<bool> kind_used;                                 /* [round * SEG_KINDS + kind] */
    std::vector<int> max_units;                                  /* [round * SEG_KINDS + kind]: longest segment among the channels */
    std::vector<int32_t> h_nseg; int32_t *d_nseg = nullptr;
    std::vector<uint64_t> seen;                                  /* chains[c]->version the tables were built from */
    double *d_scratch = nullptr; size_t scratch_cap = 0;         /* float64 rows between the rounds of encoded calls */
    /* two staging slots: block k+1 can be submitted (its upload overlapping block k's kernels and download) before
     * block k is waited for; the synchronous entry points use whichever slot is free */
    struct Slot {
        double *d_in = nullptr, *d_out = nullptr; size_t cap = 0;    /* device staging for host-buffer calls */
        double *d_enc_in = nullptr, *d_enc_out = nullptr; size_t enc_cap = 0;   /* encoded rows of the full controller path */
        double *h_stage = nullptr, *h_stage2 = nullptr; size_t h_cap = 0;   /* pinned staging (in / out) for pageable callers */
        bool busy = false;                                           /* submitted, not yet waited for */
        std::vector<std::pair<size_t, size_t>> staged;               /* output rows still in h_stage2 */
        std::vector<void *> out; size_t ob = 0;
    } slot[2];
    ~Batch() {
        for (auto p : d_ptrs) gdgb_free((void *)p);
        gdgb_free(d_nseg); gdgb_free(d_scratch);
        for (auto &sl : slot) { if (sl.d_in) gdgb_free(sl.d_in); if (sl.d_out) gdgb_free(sl.d_out); if (sl.d_enc_in) gdgb_free(sl.d_enc_in); if (sl.d_enc_out) gdgb_free(sl.d_enc_out); if (sl.h_stage) gdgb_pinned_free(sl.h_stage); if (sl.h_stage2) gdgb_pinned_free(sl.h_stage2); }
    }
};
namespace {
std::mutex g_batch_mutex;
std::unique_ptr<Batch> g_batch_of[GDGB_MAX_CONTEXTS];       /* the batch cache of each device context */
#define g_batch (g_batch_of[gdgb_current()])
} // namespace
gdg_chain::~gdg_chain() { for (auto p : d_desc) gdgb_free(p); if (d_in) gdgb_free(d_in); if (d_out) gdgb_free(d_out); delete self; }
namespace {

/* bring descriptors + launch tables of batch `b` up to date; returns the number of segment rounds or -1 */
int prepare_tables(Batch &b, gdg_chain *const *chains, size_t C, uint32_t rate, size_t n, void *stream)
{
    size_t rounds = 1;
    std::vector<int32_t> nseg(C);
    for (size_t c = 0; c < C; c++) {
        const int r = chains[c]->prepare(rate, n);
        if (r < 0) return -1;
        nseg[c] = r; rounds = std::max(rounds, (size_t)r);
    }
    bool touched = b.seen.size() != C;
    for (size_t c = 0; c < C && !touched; c++) touched = chains[c]->version != b.seen[c];
    if (touched || b.h_ptrs.size() != rounds * SEG_KINDS) {          /* tables only change when some chain changed */
        std::vector<std::vector<const gdg_chain_desc *>> ptrs(rounds * SEG_KINDS, std::vector<const gdg_chain_desc *>(C, nullptr));
        std::vector<bool> used(rounds * SEG_KINDS, false);
        for (size_t r = 0; r < rounds; r++)
            for (size_t c = 0; c < C; c++) if ((int)r < nseg[c]) {
                const size_t t = r * SEG_KINDS + (size_t)chains[c]->seg_kind[r];
                ptrs[t][c] = chains[c]->d_desc[r]; used[t] = true;
            }
        std::vector<int> mu(rounds * SEG_KINDS, 0);
        for (size_t c = 0; c < C; c++)
            for (size_t r = 0; r < (size_t)nseg[c]; r++) {
                int &m = mu[r * SEG_KINDS + (size_t)chains[c]->seg_kind[r]];
                m = std::max(m, (int)chains[c]->h_desc[r].nslots);
            }
        b.max_units = mu;
        if (ptrs != b.h_ptrs || nseg != b.h_nseg) {
            gdgb_sync();                                              /* launches still reading the old tables */
            for (auto p : b.d_ptrs) gdgb_free((void *)p);
            b.d_ptrs.clear();
            for (size_t t = 0; t < ptrs.size(); t++) {
                const gdg_chain_desc **dp = (const gdg_chain_desc **)gdgb_alloc(C * sizeof(void *));
                if (!dp || gdgb_upload((void *)dp, ptrs[t].data(), C * sizeof(void *))) return backend_fail("descriptor table upload failed");
                b.d_ptrs.push_back(dp);
            }
            gdgb_free(b.d_nseg);
            b.d_nseg = (int32_t *)gdgb_alloc(C * sizeof(int32_t));
            if (!b.d_nseg || gdgb_upload(b.d_nseg, nseg.data(), C * sizeof(int32_t))) return backend_fail("segment table upload failed");
            b.h_ptrs = ptrs; b.kind_used = used; b.h_nseg = nseg;
        }
        /* descriptor / state uploads run on the library stream: order a foreign stream after them */
        if (stream != gdgb_stream() && gdgb_sync()) return backend_fail("synchronize failed");
        b.seen.resize(C);
        for (size_t c = 0; c < C; c++) b.seen[c] = chains[c]->version;
    }
    return (int)rounds;
}
int prepare_batch(gdg_chain *const *chains, size_t C, uint32_t rate, size_t n, void *stream)
{
    if (!g_batch || g_batch->chains.size() != C || !std::equal(g_batch->chains.begin(), g_batch->chains.end(), chains)) {
        if (g_batch && (g_batch->slot[0].busy || g_batch->slot[1].busy))
            return fail("a submitted block of another chain set is still in flight: gdg_controller_wait() for it first");
        gdgb_sync();
        g_batch.reset(new Batch());
        g_batch->chains.assign(chains, chains + C);
    }
    return prepare_tables(*g_batch, chains, C, rate, n, stream);
}
size_t fmt_bytes(int fmt) { static const size_t b[6] = {1, 2, 3, 4, 4, 8}; return (fmt >= 0 && fmt < 6) ? b[fmt] : 0; }

/* launch channels [c0, c0 + cnt) of prepared batch `b`; strides in bytes, samples in the given formats.  Round r of channel
 * c reads the caller's input rows when r == 0 and the float64 rows of the previous round otherwise, and writes the caller's
 * output rows when it is the channel's LAST segment, float64 rows otherwise.  With float64 I/O the in-between rows are the
 * output rows themselves; encoded calls with more than one round keep them in a scratch buffer.
 * d_side: optional float64 rows (stride n) that receive a copy of an encoded result. */
int launch_tables(Batch &b, size_t c0, size_t cnt, int rounds, const void *d_in, size_t in_stride, int in_fmt,
                  void *d_out, size_t out_stride, int out_fmt, size_t n, void *stream, double *d_side = nullptr)
{
    unsigned char *mid = (unsigned char *)d_out; size_t mid_stride = out_stride;
    if (rounds > 1 && out_fmt != GDG_FMT_IEEE64) {
        const size_t C = b.chains.size();
        if (b.scratch_cap < C * n) {
            gdgb_sync(); gdgb_free(b.d_scratch);
            b.d_scratch = (double *)gdgb_alloc(C * n * sizeof(double));
            if (!b.d_scratch) return backend_fail("scratch allocation failed");
            b.scratch_cap = C * n;
            if (stream != gdgb_stream() && gdgb_sync()) return backend_fail("synchronize failed");
        }
        mid = (unsigned char *)b.d_scratch; mid_stride = n * sizeof(double);
    }
    for (int r = 0; r < rounds; r++)
        for (int k = 0; k < SEG_KINDS; k++) {
            const size_t t = (size_t)r * SEG_KINDS + (size_t)k;
            if (!b.kind_used[t]) continue;
            if (gdgb_launch_segments(k, b.d_ptrs[t] + c0, b.d_nseg + c0, r, (int)cnt,
                                     (const unsigned char *)d_in + c0 * in_stride, (long)in_stride, in_fmt,
                                     (unsigned char *)d_out + c0 * out_stride, (long)out_stride, out_fmt,
                                     mid + c0 * mid_stride, (long)mid_stride, d_side ? d_side + c0 * n : nullptr, (long)n, (int)n,
                                     t < b.max_units.size() ? b.max_units[t] : 0, stream))
                return backend_fail("chain kernel launch failed");
        }
    return 0;
}
int launch_batch_fmt(size_t c0, size_t cnt, int rounds, const void *d_in, size_t in_stride, int in_fmt,
                     void *d_out, size_t out_stride, int out_fmt, size_t n, void *stream, double *d_side = nullptr)
{
    return launch_tables(*g_batch, c0, cnt, rounds, d_in, in_stride, in_fmt, d_out, out_stride, out_fmt, n, stream, d_side);
}
int launch_batch(size_t c0, size_t cnt, int rounds, const double *d_in, double *d_out, size_t stride, size_t n, void *stream)
{
    return launch_batch_fmt(c0, cnt, rounds, d_in, stride * sizeof(double), GDG_FMT_IEEE64, d_out, stride * sizeof(double), GDG_FMT_IEEE64, n, stream);
}
int run_batch_device(gdg_chain *const *chains, size_t C, const double *d_in, double *d_out, size_t stride, size_t n,
                     uint32_t rate, void *stream)
{
    if (check_block(n)) return -1;
    ProcessLock el(g_edit_mutex);
    std::lock_guard<std::mutex> lk(g_batch_mutex);
    const int rounds = prepare_batch(chains, C, rate, n, stream);
    if (rounds < 0) return -1;
    /* Two halves of the channels on two kernel streams: the tail of one half's kernel (CTAs of the last wave, the 1.7 waves of
     * the oversampled-run kernel) overlaps the next kernel of the other half (measured -3 % per CHAIN12 pass; 3 and 4 ways
     * are no better).  The caller's stream is forked and joined inside this call, so its ordering is unchanged. */
    static const long ways_env = getenv("GDG_WAYS") ? atol(getenv("GDG_WAYS")) : 0;    /* tuning knob */
    const size_t ways = ways_env > 0 ? std::min<size_t>((size_t)ways_env, 1 + GDGB_SIDE_STREAMS)
                                     : (C >= 2 * (size_t)std::max(1, gdgb_sm_count()) ? 2 : 1);
    if (ways > 1 && C >= 2 * ways) {
        const size_t per = (C + ways - 1) / ways;
        for (size_t w = 1; w < ways; w++) if (gdgb_order(stream, gdgb_side_stream((int)w - 1), 56 + (int)w)) return backend_fail("event failed");
        for (size_t w = 0; w < ways; w++) {
            const size_t c0 = w * per, cnt = std::min(per, C - c0);
            if (launch_batch(c0, cnt, rounds, d_in, d_out, stride, n, w ? gdgb_side_stream((int)w - 1) : stream)) return -1;
        }
        for (size_t w = 1; w < ways; w++) if (gdgb_order(gdgb_side_stream((int)w - 1), stream, 59 + (int)w)) return backend_fail("event failed");
        return 0;
    }
    return launch_batch(0, C, rounds, d_in, d_out, stride, n, stream);
}

/* Host-buffer pipeline: channels are cut into groups; H2D of group g+1, the kernel of group g and D2H of
 * group g-1 overlap on three streams.  `extra_out` rows (spatializer master) follow the channel rows. */
size_t pipeline_group_size(size_t C, bool blocking)
{
    static const long forced = getenv("GDG_PIPE_GROUP") ? atol(getenv("GDG_PIPE_GROUP")) : 0;   /* tuning knob */
    if (forced > 0) return (size_t)forced;
    /* An EVEN number of equal groups (they alternate between two kernel streams, block_submit), each at most four
     * CTA-per-SM waves of the chain kernel = one full wave of the oversampled-run kernel (smaller groups run k_os below its
     * occupancy).  A blocking call has nothing else in flight to hide its first upload and last download behind, so it cuts
     * finer (two waves).  Measured at 1024 channels, Msamples/s e2e submit/wait | blocking call:
     *   groups of 1024: 3639 | 1858,  592 + 432: 3820 | 2380,  2 x 512: 3940 | 2332,  4 x 256: 3598 | 2620. */
    const size_t sms = (size_t)std::max(1, gdgb_sm_count());
    if (C < 2 * sms) return C;
    const size_t target = (blocking ? 2 : 4) * sms;
    const size_t G = 2 * ((C + 2 * target - 1) / (2 * target));
    return (C + G - 1) / G;
}

bool contiguous(const double *const *p, size_t C, size_t n)
{
    for (size_t c = 1; c < C; c++) if (p[c] != p[0] + c * n) return false;
    return true;
}

int stage_buffers(Batch::Slot &b, size_t words)
{
    if (b.cap < words) {
        if (b.d_in) gdgb_free(b.d_in);
        if (b.d_out) gdgb_free(b.d_out);
        b.d_in = (double *)gdgb_alloc(words * sizeof(double));
        b.d_out = (double *)gdgb_alloc(words * sizeof(double));
        if (!b.d_in || !b.d_out) return backend_fail("device staging allocation failed");
        b.cap = words;
    }
    return 0;
}

/* host rows [c0, c0+cnt) (array of pointers) -> device planar, asynchronously on `stream_id` */
bool contiguous_bytes(const void *const *p, size_t C, size_t row_bytes)
{
    for (size_t c = 1; c < C; c++) if ((const char *)p[c] != (const char *)p[0] + c * row_bytes) return false;
    return true;
}
/* rows are `bytes` long (n samples of the batch's input format) */
int upload_rows(Batch::Slot &b, const void *const *in, size_t c0, size_t cnt, size_t bytes, void *d_dst, int stream_id)
{
    char *dst = (char *)d_dst, *hs = (char *)b.h_stage;
    if (contiguous_bytes(in + c0, cnt, bytes) && gdgb_is_pinned(in[c0]))
        return gdgb_h2d_async(dst + c0 * bytes, in[c0], cnt * bytes, stream_id) ? backend_fail("upload failed") : 0;
    /* pageable / scattered rows: gather into the library's pinned staging, then one copy */
    for (size_t c = c0; c < c0 + cnt; c++) memcpy(hs + c * bytes, in[c], bytes);
    return gdgb_h2d_async(dst + c0 * bytes, hs + c0 * bytes, cnt * bytes, stream_id) ? backend_fail("upload failed") : 0;
}
/* device planar rows -> host rows; returns 1 if the data still sits in the pinned staging (scatter after sync) */
int download_rows(Batch::Slot &b, void *const *out, size_t c0, size_t cnt, size_t bytes, const void *d_src, int stream_id)
{
    const char *src = (const char *)d_src;
    if (contiguous_bytes((const void *const *)out + c0, cnt, bytes) && gdgb_is_pinned(out[c0]))
        return gdgb_d2h_async(out[c0], src + c0 * bytes, cnt * bytes, stream_id) ? backend_fail("download failed") : 0;
    if (gdgb_d2h_async((char *)b.h_stage2 + c0 * bytes, src + c0 * bytes, cnt * bytes, stream_id)) return backend_fail("download failed");
    return 1;
}
int ensure_host_stage(Batch::Slot &b, size_t words)
{
    if (b.h_cap < words) {
        if (b.h_stage) gdgb_pinned_free(b.h_stage);
        if (b.h_stage2) gdgb_pinned_free(b.h_stage2);
        b.h_stage = (double *)gdgb_pinned_alloc(words * sizeof(double));
        b.h_stage2 = (double *)gdgb_pinned_alloc(words * sizeof(double));
        if (!b.h_stage || !b.h_stage2) return backend_fail("pinned staging allocation failed");
        b.h_cap = words;
    }
    return 0;
}

int upload_simple(const double *const *in, size_t C, size_t n, double *d_dst)
{
    if (contiguous(in, C, n)) return gdgb_upload(d_dst, in[0], C * n * sizeof(double)) ? backend_fail("upload failed") : 0;
    for (size_t c = 0; c < C; c++) if (gdgb_upload(d_dst + c * n, in[c], n * sizeof(double))) return backend_fail("upload failed");
    return 0;
}
int download_simple(double *const *out, size_t C, size_t n, const double *d_src)
{
    if (contiguous((const double *const *)out, C, n)) return gdgb_download(out[0], d_src, C * n * sizeof(double)) ? backend_fail("download failed") : 0;
    for (size_t c = 0; c < C; c++) if (gdgb_download(out[c], d_src + c * n, n * sizeof(double))) return backend_fail("download failed");
    return 0;
}

/* the whole host-buffer block: upload -> (tuner feed) -> chains -> (spatializer) -> download, pipelined */
int host_block(gdg_chain *const *chains, size_t C, gdg_tuner *tuner, int tuner_channel, gdg_spatializer *spat, struct gdg_meter *meter,
               const void *const *in, int in_fmt, void *const *out, int out_fmt, size_t n_out_rows, size_t n, uint32_t rate,
               const double *metronome = nullptr, bool metro_master = false);

} // namespace

/* ---- spatializer -------------------------------------------------------------- */
struct gdg_spatializer {
    int ctx = 0;
    /* spatializer.go:59: setters arrive from the UI thread while the audio thread processes.  A leaf lock (nothing else is
     * taken under it): setters hold it while they write a position, the processing call while it turns the positions into
     * the per-channel constants and enqueues its kernel, SetSampleRate while it re-makes the histories */
    std::mutex mu;
    uint32_t count; uint32_t sample_rate = 96000;         /* never updated: spatializer.go:418-431 does not store the rate */
    struct Pos { double az = 0, dist = 0, level = 1.0; };
    std::vector<Pos> pos;
    int hs; bool dirty = true; int flip = 0;
    gdg_spat_channel *d_ch = nullptr; double *d_hist[2] = {nullptr, nullptr};
    double *d_in = nullptr, *d_out = nullptr, *d_aux = nullptr; size_t cap_n = 0;
    double *d_part = nullptr; size_t part_n = 0;          /* partial sums of 32-channel groups */
    ~gdg_spatializer() { gdgb_free(d_part); gdgb_free(d_ch); gdgb_free(d_hist[0]); gdgb_free(d_hist[1]); gdgb_free(d_in); gdgb_free(d_out); gdgb_free(d_aux); }

    int alloc_hist(uint32_t rate)
    {
        hs = (int)std::ceil((double)rate * GDG_SPAT_GROUP_DELAY);
        for (int k = 0; k < 2; k++) {
            gdgb_free(d_hist[k]);
            d_hist[k] = (double *)gdgb_alloc(std::max<size_t>(1, (size_t)count * hs) * sizeof(double));
            if (!d_hist[k]) return backend_fail("spatializer history allocation failed");
        }
        flip = 0; dirty = true;
        return 0;
    }
    /* per-channel constants of spatializer.go:171-232 */
    int upload_channels()
    {
        std::vector<gdg_spat_channel> ch(count);
        const double sr = (double)sample_rate;
        for (uint32_t c = 0; c < count; c++) {
            const Pos &p = pos[c];
            const double az = GDG_MATH_DEGREE_TO_RADIANS * p.az;
            const double sn = std::sin(az), cs = std::cos(az);
            const double xp = p.dist * sn, yp = p.dist * cs;
            const double xl = std::fabs(xp + GDG_SPAT_HALF_EFFECTIVE_DISTANCE), xr = std::fabs(xp - GDG_SPAT_HALF_EFFECTIVE_DISTANCE);
            const double yd = std::fabs(yp), yd2 = yd * yd;
            const double dl = std::sqrt(xl * xl + yd2), dr = std::sqrt(xr * xr + yd2);
            double pl = 1.0 / dl; if (pl > 1.0) pl = 1.0;
            double pr = 1.0 / dr; if (pr > 1.0) pr = 1.0;
            gdg_spat_channel &o = ch[c];
            o.fac_left = p.level * pl; o.fac_right = p.level * pr;
            o.delay_time = GDG_SPAT_DELAY_PER_METRE * (dl - dr);
            const double ds = std::fabs(o.delay_time) * sr;
            const double early = std::floor(ds), late = std::ceil(ds);
            o.d_early = (int)early; if (o.d_early >= hs) o.d_early = hs - 1;
            o.d_late = (int)late; if (o.d_late >= hs) o.d_late = hs - 1;
            o.w_early = 1.0 - (ds - early); o.w_late = 1.0 - (late - ds);
        }
        if (!d_ch) d_ch = (gdg_spat_channel *)gdgb_alloc(std::max<size_t>(1, count) * sizeof(gdg_spat_channel));
        if (!d_ch || gdgb_upload(d_ch, ch.data(), count * sizeof(gdg_spat_channel))) return backend_fail("spatializer upload failed");
        dirty = false;
        return 0;
    }
};

/* ---- tuner ---------------------------------------------------------------------- */
struct gdg_tuner {
    int ctx = 0;
    static const int N = 96000;                           /* tuner.go:16 NUM_SAMPLES */
    double *d_ring = nullptr, *d_work = nullptr, *d_block = nullptr; size_t block_cap = 0;
    int pos = 0; uint32_t rate = 0;
    std::vector<double> corr;
    ~gdg_tuner() { gdgb_free(d_ring); gdgb_free(d_work); gdgb_free(d_block); }
};
namespace {
/* tuner.go:50-53 has two locks and so does this side: g_tuner_mutex (mutexBuffer) is a leaf lock around a ring's write
 * position - Process holds it while it enqueues the ring copy, Analyze while it reads the position and enqueues the
 * autocorrelation, so an analysis never pairs a new position with an old ring; g_analyze_mutex (mutexAnalyze) serialises
 * whole analyses (they share a tuner's work buffer).  Lock order: g_analyze_mutex -> g_batch_mutex -> g_tuner_mutex. */
std::mutex g_tuner_mutex, g_analyze_mutex;
}

/* ---- level meter ---------------------------------------------------------------- */
struct gdg_meter {
    int ctx = 0;
    std::vector<std::string> names;
    uint32_t count = 0;
    std::atomic<bool> enabled{false};                     /* level.go:383-386: created disabled; SetEnabled may come from the UI thread (level.go:333-356) */
    double *d_state = nullptr;                            /* [count][4]: cur, peak, counter, unused */
    double *d_stage = nullptr; size_t stage_cap = 0;      /* rows of the host entry point */
    ~gdg_meter() { gdgb_free(d_state); gdgb_free(d_stage); }
};

namespace {
int meter_rows(gdg_meter *m, uint32_t first, const double *d_rows, size_t stride, uint32_t rows, size_t n, uint32_t rate, void *stream)
{
    if (!m->enabled || !rows || !n) return 0;             /* level.go:156: disabled channels do nothing */
    const double sr = (double)rate;
    const long long hold = (long long)(uint64_t)(2.0 * sr);                  /* level.go:163 */
    const double f = std::pow(10.0, -1.0 / (1.7 * sr));                       /* level.go:164-165 */
    if (gdgb_launch_meter(d_rows, (long)stride, (int)rows, (long)n, f, hold, m->d_state + 4 * (size_t)first, stream ? stream : gdgb_stream()))
        return backend_fail("level meter launch failed");
    return 0;
}

/* Enqueue one block on staging slot `si` (caller holds g_edit_mutex shared and g_batch_mutex): uploads on stream 1, kernels on
 * stream 0, downloads on stream 2, one SM wave of channels per pipeline group, everything of a group enqueued before the next
 * group is touched (an event is recorded and waited for within the same iteration, so the small id pool never aliases).
 * Nothing here waits for the device.
 * in_fmt / out_fmt: sample formats of the host rows (tuner, spatializer and level meter need float64 rows).
 * metronome: host row of n float64 samples = the metronome's output buffer of this block (controller.go:2724-2738), or NULL;
 * it feeds meter channel 2C and, when `metro_master`, the spatializer's aux input (controller.go:2746-2761).
 * Event ids: slot*28 + {0..7 uploads, 8..15 kernels, 16 spatializer, 17 all downloads, 18 metronome upload}. */
int block_submit(int si, gdg_chain *const *chains, size_t C, gdg_tuner *tuner, int tuner_channel, gdg_spatializer *spat, gdg_meter *meter,
                 const void *const *in, int in_fmt, void *const *out, int out_fmt, size_t n_out_rows, size_t n, uint32_t rate,
                 const double *metronome = nullptr, bool metro_master = false, const double **partial_master = nullptr, bool blocking = false)
{
    /* partial_master != NULL (multi-device blocks): the spatializer's [2][n] result stays on the device (row stride n) and its
     * address is returned; event slot*28 + 16 marks it ready for the device that sums the partial masters */
    const int rounds = prepare_batch(chains, C, rate, n, gdgb_stream());
    if (rounds < 0) return -1;
    Batch::Slot &b = g_batch->slot[si];
    const int ev = si * 28;
    const size_t ib = n * fmt_bytes(in_fmt), ob = n * fmt_bytes(out_fmt);      /* bytes per host row */
    /* Encoded host rows are decoded / encoded inside the segment kernels.  Tuner, spatializer and meter read float64 rows:
     * the last segment stores a float64 copy of its result beside the encoded one (d_side), the tuner's single input row is
     * decoded on its own, and only a level meter (which reads EVERY input row) has the whole input payload converted first. */
    const bool aux = tuner || spat || meter;
    const bool enc_in = aux && in_fmt != GDG_FMT_IEEE64, enc_out = aux && out_fmt != GDG_FMT_IEEE64;
    const bool dec_in = enc_in && meter;
    const bool recode = enc_in || enc_out;
    const int kin = dec_in ? GDG_FMT_IEEE64 : in_fmt, kout = out_fmt;             /* formats the chain kernels see */
    const size_t kib = n * fmt_bytes(kin), kob = n * fmt_bytes(kout);
    if (stage_buffers(b, (C + 3) * n)) return -1;                             /* C channel rows + master L/R + metronome */
    if (recode && b.enc_cap < (C + 2) * n) {
        if (b.d_enc_in) gdgb_free(b.d_enc_in);
        if (b.d_enc_out) gdgb_free(b.d_enc_out);
        b.d_enc_in = (double *)gdgb_alloc((C + 2) * n * sizeof(double));
        b.d_enc_out = (double *)gdgb_alloc((C + 2) * n * sizeof(double));
        if (!b.d_enc_in || !b.d_enc_out) return backend_fail("device staging allocation failed");
        b.enc_cap = (C + 2) * n;
    }
    void *up_dst = enc_in ? (void *)b.d_enc_in : (void *)b.d_in;
    const void *k_src = (enc_in && !dec_in) ? (const void *)b.d_enc_in : (const void *)b.d_in;    /* rows the chain kernels read */
    void *k_dst = enc_out ? (void *)b.d_enc_out : (void *)b.d_out;                                /* ... and write */
    double *k_side = (enc_out && (spat || meter)) ? b.d_out : nullptr;
    const void *down_src = k_dst;
    const bool in_direct = contiguous_bytes(in, C, ib) && gdgb_is_pinned(in[0]);
    const bool out_direct = contiguous_bytes((const void *const *)out, n_out_rows, ob) && gdgb_is_pinned(out[0]);
    if ((!in_direct || !out_direct) && ensure_host_stage(b, (C + 2) * n)) return -1;
    const size_t per = pipeline_group_size(C, blocking);
    const int G = (int)((C + per - 1) / per);
    void *s0 = gdgb_stream_handle(0);
    b.staged.clear(); b.out.assign(out, out + n_out_rows); b.ob = ob;
    double *d_metro = nullptr;
    if (metronome && (spat || meter)) {                                        /* the metronome row rides on the upload stream */
        d_metro = b.d_in + (C + 2) * n;
        if (gdgb_h2d_async(d_metro, metronome, n * sizeof(double), 1) || gdgb_record(ev + 18, 1)) return backend_fail("metronome upload failed");
    }
    /* groups alternate between the kernel stream and a side stream (forked here, joined after the loop): the kernels of
     * neighbouring groups overlap each other's tails, see run_batch_device */
    void *side = G > 1 ? gdgb_side_stream(0) : s0;
    if (G > 1 && gdgb_order(s0, side, ev + 19)) return backend_fail("event failed");
    for (int g = 0; g < G; g++) {
        const size_t c0 = (size_t)g * per;
        if (c0 >= C) break;
        const size_t cnt = std::min(per, C - c0);
        void *ks = (g & 1) ? side : s0;                                         /* odd groups run beside the even ones */
        if (upload_rows(b, in, c0, cnt, ib, up_dst, 1)) return -1;
        if (gdgb_order(gdgb_stream_handle(1), ks, ev + g % 8)) return backend_fail("event failed");
        if (dec_in && gdgb_convert_rows((const char *)b.d_enc_in + c0 * ib, (long)ib, in_fmt, b.d_in + c0 * n, (long)(n * sizeof(double)), GDG_FMT_IEEE64, (int)cnt, (int)n, ks))
            return backend_fail("payload decode failed");
        if (tuner && tuner_channel >= 0 && (size_t)tuner_channel >= c0 && (size_t)tuner_channel < c0 + cnt) {
            double *row = b.d_in + (size_t)tuner_channel * n;
            if (enc_in && !dec_in && gdgb_convert_rows((const char *)b.d_enc_in + (size_t)tuner_channel * ib, (long)ib, in_fmt, row, (long)(n * sizeof(double)), GDG_FMT_IEEE64, 1, (int)n, ks))
                return backend_fail("payload decode failed");
            if (gdg_tuner_process_device(tuner, row, n, rate, ks)) return -1;                                    /* controller.go:2668-2672 */
        }
        if (meter && meter_rows(meter, (uint32_t)c0, b.d_in + c0 * n, n, (uint32_t)cnt, n, rate, ks)) return -1;      /* controller.go:2710-2714: inputs */
        if (launch_batch_fmt(c0, cnt, rounds, k_src, kib, kin, k_dst, kob, kout, n, ks, k_side)) return -1;      /* controller.go:2682-2705 */
        if (meter && meter_rows(meter, (uint32_t)(C + c0), b.d_out + c0 * n, n, (uint32_t)cnt, n, rate, ks)) return -1;   /* outputs */
        if (gdgb_order(ks, gdgb_stream_handle(2), ev + 8 + g % 8)) return backend_fail("event failed");
        const int r = download_rows(b, out, c0, cnt, ob, down_src, 2);
        if (r < 0) return -1;
        if (r == 1) b.staged.push_back({c0, cnt});
    }
    if (G > 1 && gdgb_order(side, s0, ev + 20)) return backend_fail("event failed");
    if (d_metro && gdgb_wait(0, ev + 18)) return backend_fail("stream wait failed");
    if (d_metro && meter && meter_rows(meter, (uint32_t)(2 * C), d_metro, n, 1, n, rate, s0)) return -1;        /* controller.go:2735-2738: the metronome port */
    if (spat) {                                                                                                 /* controller.go:2761 */
        if (gdg_spatializer_process_device(spat, b.d_out, n, (d_metro && metro_master) ? d_metro : nullptr, b.d_out + C * n, n, n, 0, s0)) return -1;
        if (meter && !partial_master && meter_rows(meter, (uint32_t)(2 * C + 1), b.d_out + C * n, n, 2, n, rate, s0)) return -1;    /* controller.go:2768-2770: master (a partial master is metered after the sum) */
        if (enc_out && !partial_master && gdgb_convert_rows(b.d_out + C * n, (long)(n * sizeof(double)), GDG_FMT_IEEE64, (char *)b.d_enc_out + C * ob, (long)ob, out_fmt, 2, (int)n, s0))
            return backend_fail("payload encode failed");
        if (gdgb_record(ev + 16, 0)) return backend_fail("event failed");
        if (partial_master) *partial_master = b.d_out + C * n;
        else {
            if (gdgb_wait(2, ev + 16)) return backend_fail("event failed");
            const int r = download_rows(b, out, C, 2, ob, down_src, 2);
            if (r < 0) return -1;
            if (r == 1) b.staged.push_back({C, 2});
        }
    }
    if (gdgb_record(ev + 17, 2)) return backend_fail("event record failed");
    b.busy = true;
    return 0;
}
/* the host waits for slot `si`'s downloads and scatters rows that went through the pinned staging */
int block_wait(int si)
{
    Batch::Slot &b = g_batch->slot[si];
    b.busy = false;
    if (gdgb_sync_event(si * 28 + 17)) return backend_fail("synchronize failed");
    for (auto &sg : b.staged)
        for (size_t c = sg.first; c < sg.first + sg.second; c++) memcpy(b.out[c], (char *)b.h_stage2 + c * b.ob, b.ob);
    b.staged.clear();
    return 0;
}
int free_slot()
{
    if (!g_batch) return 0;
    return !g_batch->slot[0].busy ? 0 : (!g_batch->slot[1].busy ? 1 : -1);
}
int host_block(gdg_chain *const *chains, size_t C, gdg_tuner *tuner, int tuner_channel, gdg_spatializer *spat, gdg_meter *meter,
               const void *const *in, int in_fmt, void *const *out, int out_fmt, size_t n_out_rows, size_t n, uint32_t rate,
               const double *metronome, bool metro_master)
{
    ProcessLock el(g_edit_mutex);
    std::lock_guard<std::mutex> lk(g_batch_mutex);
    const int si = free_slot();
    if (si < 0) return fail("two submitted blocks are in flight: gdg_controller_wait() for one first");
    if (block_submit(si, chains, C, tuner, tuner_channel, spat, meter, in, in_fmt, out, out_fmt, n_out_rows, n, rate, metronome, metro_master, nullptr, true)) {
        if (g_batch) { gdgb_sync(); g_batch->slot[si].busy = false; }
        return -1;
    }
    return block_wait(si);
}
} // namespace

/* ================================================================================ */
/* C ABI                                                                             */
/* ================================================================================ */
extern "C" {

const char *gdg_last_error(void) { return g_error.c_str(); }
const char *gdg_backend_name(void) { return gdgb_name(); }
const char *gdg_version(void) { return GDG_VERSION; }
uint64_t gdg_launch_count(void) { return gdgb_launch_count(); }
void *gdg_pinned_alloc(size_t bytes) { if (ensure_init()) return nullptr; return gdgb_pinned_alloc(bytes); }
void gdg_pinned_free(void *p) { gdgb_pinned_free(p); }
void *gdg_stream(void) { if (ensure_init()) return nullptr; return gdgb_stream(); }
int gdg_synchronize(void) { if (ensure_init()) return -1; return gdgb_sync() ? backend_fail("synchronize") : 0; }
int gdg_unit_type_count(void) { return GDG_UNIT_COUNT; }
const char *gdg_unit_type_name(int t) { return (t >= 0 && t < GDG_UNIT_COUNT) ? k_unit_names[t] : nullptr; }

int gdg_init(int device)
{
    static std::mutex m; static bool done = false; static int status = -1;
    std::lock_guard<std::mutex> lk(m);
    if (done) return status;
    done = true;
    if (gdgb_init(device)) { g_init_error = std::string("gdg_init: no usable CUDA device (there is no CPU fallback): ") + gdgb_error(); g_error = g_init_error; return status = -1; }
    /* constant tables */
    static const double h2[] = GDG_AA2_HALF_INIT, h4[] = GDG_AA4_HALF_INIT;
    std::vector<double> aa2(80, 0.0), aa4(160, 0.0);
    for (int k = 0; k < GDG_AA2_TAPS; k++) aa2[k] = h2[std::min(k, GDG_AA2_TAPS - 1 - k)];
    for (int k = 0; k < GDG_AA4_TAPS; k++) aa4[k] = h4[std::min(k, GDG_AA4_TAPS - 1 - k)];
    double lz2[6], lz4[18];
    for (int q = 0; q < 6; q++) lz2[q] = lanczos_kernel(2.0 - (double)q + 0.5, 3.0);
    for (int p = 1; p < 4; p++) for (int q = 0; q < 6; q++) lz4[(p - 1) * 6 + q] = lanczos_kernel(2.0 - (double)q + (double)p * 0.25, 3.0);
    std::vector<double> tw(GDG_MAX_FFT);                   /* GDG_MAX_FFT/2 complex */
    for (int k = 0; k < GDG_MAX_FFT / 2; k++) {
        const double arg = (GDG_MATH_MINUS_TWO_PI * (double)k) / (double)GDG_MAX_FFT;
        tw[2 * k] = std::cos(arg); tw[2 * k + 1] = std::sin(arg);
    }
    if (gdgb_set_tables(aa2.data(), aa4.data(), lz2, lz4, tw.data())) { g_init_error = std::string("gdg_init: table upload failed: ") + gdgb_error(); g_error = g_init_error; return status = -1; }
    g_init_status = 0;
    return status = 0;
}

/* ---- impulse responses ---- */
gdg_irbank *gdg_irbank_create(void) { return new gdg_irbank(); }
void gdg_irbank_destroy(gdg_irbank *b) { if (!b) return; Use use_ctx_(b->ctx); delete b; }
int gdg_irbank_add(gdg_irbank *b, const char *name, int32_t comp_db, const double *s, size_t n, uint32_t src_rate)
{
    if (ensure_init()) return -1;
    if (!b || !name || (!s && n)) return fail("gdg_irbank_add: bad argument");
    if (src_rate == 0) return fail("gdg_irbank_add: source rate must not be zero");
    if (b->ctx < 0) b->ctx = gdgb_current();
    Use use_ctx_(b->ctx);
    const double fac = std::pow(10.0, 0.05 * (double)comp_db);      /* filter.go:742-745 */
    /* filter.Import's resample.Time to the seven supported rates (filter.go:766-780) runs on the device; the responses stay there */
    double *d_src = (double *)gdgb_alloc(std::max<size_t>(1, n) * sizeof(double));
    if (!d_src || (n && gdgb_upload(d_src, s, n * sizeof(double)))) { gdgb_free(d_src); return backend_fail("impulse response upload failed"); }
    int rc = 0;
    for (uint32_t r : k_sample_rates) {
        const size_t len = resample_time_length(n, src_rate, r);
        double *d = (double *)gdgb_alloc(std::max<size_t>(1, len) * sizeof(double));
        if (!d || gdgb_resample_time(d_src, (long)n, (double)src_rate / (double)r, d, (long)len, gdgb_stream())) { gdgb_free(d); rc = backend_fail("impulse response resampling failed"); break; }
        b->responses.emplace_back(name, r, fac, d, len);
    }
    gdgb_sync(); gdgb_free(d_src);
    b->generation++;
    return rc;
}
int gdg_irbank_add_raw(gdg_irbank *b, const char *name, int32_t comp_db, const double *s, size_t n, uint32_t rate)
{
    if (ensure_init()) return -1;
    if (!b || !name || (!s && n)) return fail("gdg_irbank_add_raw: bad argument");
    if (b->ctx < 0) b->ctx = gdgb_current();
    Use use_ctx_(b->ctx);
    double *d = (double *)gdgb_alloc(std::max<size_t>(1, n) * sizeof(double));
    if (!d || (n && gdgb_upload(d, s, n * sizeof(double)))) { gdgb_free(d); return backend_fail("impulse response upload failed"); }
    gdgb_sync();
    b->responses.emplace_back(name, rate, std::pow(10.0, 0.05 * (double)comp_db), d, n);
    b->generation++;
    return 0;
}
size_t gdg_irbank_name_count(const gdg_irbank *b) { return b ? b->names().size() : 0; }
const char *gdg_irbank_name(const gdg_irbank *b, size_t i)
{
    if (!b) return nullptr;
    gdg_irbank *mb = const_cast<gdg_irbank *>(b);
    static std::mutex m;
    std::lock_guard<std::mutex> lk(m);
    if (mb->name_cache_generation != b->generation) { mb->name_cache = b->names(); mb->name_cache_generation = b->generation; }
    return i < mb->name_cache.size() ? mb->name_cache[i].c_str() : nullptr;
}

/* minimal readers for filter.Import: a JSON array of {"Name","Path","Compensation"} and mono RIFF/WAVE */
static bool read_file(const std::string &path, std::vector<unsigned char> &out)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    if (sz < 0 || sz > (1l << 31)) { fclose(f); return false; }                 /* a directory, or nothing an impulse response could be */
    out.resize((size_t)sz);
    const size_t got = sz > 0 ? fread(out.data(), 1, (size_t)sz, f) : 0;
    fclose(f);
    return got == out.size();
}
static bool json_string(const std::string &s, size_t &i, std::string &out)
{
    if (s[i] != '"') return false;
    out.clear();
    for (i++; i < s.size() && s[i] != '"'; i++) {
        if (s[i] == '\\' && i + 1 < s.size()) {
            i++;
            switch (s[i]) { case 'n': out.push_back('\n'); break; case 't': out.push_back('\t'); break;
                case 'u': { unsigned cp = (unsigned)strtoul(s.substr(i + 1, 4).c_str(), nullptr, 16); i += 4;
                    if (cp < 0x80) out.push_back((char)cp);
                    else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
                    else { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
                    break; }
                default: out.push_back(s[i]); }
        } else out.push_back(s[i]);
    }
    if (i >= s.size()) return false;
    i++;
    return true;
}
struct Descriptor { std::string name, path; int32_t comp = 0; };
static bool parse_index(const std::string &s, std::vector<Descriptor> &out)
{
    size_t i = 0;
    auto ws = [&] { while (i < s.size() && isspace((unsigned char)s[i])) i++; };
    ws(); if (i >= s.size() || s[i] != '[') return false;
    i++;
    for (;;) {
        ws(); if (i < s.size() && s[i] == ']') return true;
        if (i >= s.size() || s[i] != '{') return false;
        i++;
        Descriptor d;
        for (;;) {
            ws(); if (i < s.size() && s[i] == '}') { i++; break; }
            std::string key; if (i >= s.size() || !json_string(s, i, key)) return false;
            ws(); if (i >= s.size() || s[i] != ':') return false; i++; ws();
            if (i < s.size() && s[i] == '"') { std::string v; if (!json_string(s, i, v)) return false;
                if (!strcasecmp(key.c_str(), "Name")) d.name = v; else if (!strcasecmp(key.c_str(), "Path")) d.path = v; }
            else { char *end; const double v = strtod(s.c_str() + i, &end); if (end == s.c_str() + i) return false; i = (size_t)(end - s.c_str());
                if (!strcasecmp(key.c_str(), "Compensation")) d.comp = (int32_t)v; }
            ws(); if (i < s.size() && s[i] == ',') i++;
        }
        out.push_back(d);
        ws(); if (i < s.size() && s[i] == ',') i++;
    }
}
/* the six payload formats of wave.bytesToSamples (wave/wave.go:790-838): LPCM 8 / 16 / 24 / 32 bit (wave.go:316-342, 400-428,
 * 483-531, 589-617) and IEEE 32 / 64 bit (wave.go:662-689, 714-732); same arithmetic as the device codecs (gdg_interp.cuh) */
static bool decode_payload(int fmt, int bits, const unsigned char *p, size_t bytes, std::vector<double> &out)
{
    if (fmt == 1 && bits == 8) {
        out.resize(bytes);
        for (size_t i = 0; i < bytes; i++) { const double r = (1.0 / 127.0) * (double)((int)p[i] - 128); out[i] = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r); }
    } else if (fmt == 1 && bits == 16) {
        out.resize(bytes / 2);
        for (size_t i = 0; i < out.size(); i++) { int16_t v; memcpy(&v, p + 2 * i, 2); out[i] = GDG_PCM16_DECODE_SCALE * (double)v; }
    } else if (fmt == 1 && bits == 24) {
        out.resize(bytes / 3);
        for (size_t i = 0; i < out.size(); i++) {
            const uint32_t w = (uint32_t)p[3 * i] | ((uint32_t)p[3 * i + 1] << 8) | ((uint32_t)p[3 * i + 2] << 16);
            int32_t v = (int32_t)w;
            if (w & 0x00800000u) v = -8388608 + (v & 0x007fffff);
            out[i] = (2.0 / 16777215.0) * (double)v;
        }
    } else if (fmt == 1 && bits == 32) {
        out.resize(bytes / 4);
        for (size_t i = 0; i < out.size(); i++) { int32_t v; memcpy(&v, p + 4 * i, 4); out[i] = (2.0 / 4294967295.0) * (double)v; }
    } else if (fmt == 3 && bits == 32) {
        out.resize(bytes / 4);
        for (size_t i = 0; i < out.size(); i++) { float v; memcpy(&v, p + 4 * i, 4); out[i] = (double)v; }
    } else if (fmt == 3 && bits == 64) {
        out.resize(bytes / 8);
        if (!out.empty()) memcpy(out.data(), p, out.size() * 8);
    } else return false;
    return true;
}
static int parse_wav(const std::vector<unsigned char> &b, std::vector<double> &samples, uint32_t &rate, int &channels, std::string &err)
{
    if (b.size() < 12 || memcmp(b.data(), "RIFF", 4) || memcmp(b.data() + 8, "WAVE", 4)) { err = "not a RIFF/WAVE file"; return -1; }
    size_t pos = 12; int fmt = 0, bits = 0; channels = 0; rate = 0; bool have_fmt = false;
    while (pos + 8 <= b.size()) {
        uint32_t size; memcpy(&size, b.data() + pos + 4, 4);
        const unsigned char *payload = b.data() + pos + 8;
        const size_t avail = std::min<size_t>(size, b.size() - pos - 8);
        if (!memcmp(b.data() + pos, "fmt ", 4) && avail >= 16) {
            uint16_t f, c, bp; uint32_t r; memcpy(&f, payload, 2); memcpy(&c, payload + 2, 2); memcpy(&r, payload + 4, 4); memcpy(&bp, payload + 14, 2);
            if (f == 0xFFFE && avail >= 26) memcpy(&f, payload + 24, 2);       /* WAVE_FORMAT_EXTENSIBLE sub-format */
            fmt = f; channels = c; rate = r; bits = bp; have_fmt = true;
        } else if (!memcmp(b.data() + pos, "data", 4)) {
            if (!have_fmt) { err = "data chunk before fmt chunk"; return -1; }
            if (!decode_payload(fmt, bits, payload, avail, samples)) { err = "Unknown sample format or bit depth."; return -1; }
            if (channels > 1) {                                  /* interleaved frames: the caller rejects non-mono files, keep channel 0 */
                std::vector<double> c0(samples.size() / (size_t)channels);
                for (size_t i = 0; i < c0.size(); i++) c0[i] = samples[i * (size_t)channels];
                samples.swap(c0);
            }
            return 0;
        }
        pos += 8 + (size_t)size + (size & 1);
    }
    err = "no data chunk";
    return -1;
}

gdg_irbank *gdg_irbank_import(const char *index_json_path)
{
    std::vector<unsigned char> buf;
    if (!index_json_path || !read_file(index_json_path, buf)) { fail("Failed to read descriptor file: '%s'", index_json_path ? index_json_path : "(null)"); return nullptr; }
    std::vector<Descriptor> ds;
    if (!parse_index(std::string(buf.begin(), buf.end()), ds)) { fail("Failed to decode descriptor file: '%s'", index_json_path); return nullptr; }
    gdg_irbank *bank = new gdg_irbank();
    for (auto &d : ds) {
        std::vector<unsigned char> wav;
        if (!read_file(d.path, wav)) { printf("WARNING: During filter import: Could not read file '%s'. - Skipping.\n", d.path.c_str()); continue; }
        std::vector<double> s; uint32_t rate; int ch; std::string err;
        if (parse_wav(wav, s, rate, ch, err)) { printf("WARNING: During filter import (file '%s'): %s\n", d.path.c_str(), err.c_str()); continue; }
        if (ch != 1) { printf("WARNING: During filter import: File '%s' contains %d channels, expected: %d - Skipping.\n", d.path.c_str(), ch, 1); continue; }
        gdg_irbank_add(bank, d.name.c_str(), d.comp, s.data(), s.size(), rate);
    }
    return bank;
}

/* ---- chain ---- */
gdg_chain *gdg_chain_create(gdg_irbank *bank)
{
    if (ensure_init()) return nullptr;
    gdg_chain *c = new gdg_chain();
    c->bank = bank;
    c->ctx = (bank && bank->ctx >= 0) ? bank->ctx : gdgb_current();      /* a chain lives with its impulse responses */
    return c;
}
void gdg_chain_destroy(gdg_chain *c)
{
    if (!c) return;
    Use use_ctx_(c->ctx);
    EditLock el(g_edit_mutex);
    std::lock_guard<std::mutex> lk(g_batch_mutex);
    if (g_batch && std::find(g_batch->chains.begin(), g_batch->chains.end(), c) != g_batch->chains.end()) g_batch.reset();
    delete c;
}
#define CHAIN_CHECK(c) if (!(c)) return fail("chain is NULL"); Use use_ctx_((c)->ctx)
int gdg_chain_append_unit(gdg_chain *c, int type)
{
    EditLock lk(g_edit_mutex);
    if (c) c->clean = false;
    if (!c) { fail("chain is NULL"); return -1; }
    Use use_ctx_(c->ctx);
    Unit *u = create_unit(type, c->bank);
    if (!u) { fail("%s", "Failed to create effects unit."); return -1; }
    c->slots.push_back(Slot{std::unique_ptr<Unit>(u), true});          /* starts bypassed: signal.go:74 */
    c->topo_dirty = true;
    return (int)c->slots.size() - 1;
}
int gdg_chain_remove_unit(gdg_chain *c, int id)
{
    EditLock lk(g_edit_mutex);
    if (c) c->clean = false;
    CHAIN_CHECK(c);
    if (id < 0 || id >= (int)c->slots.size()) return fail("Cannot remove unit %d.", id);
    gdgb_sync();
    c->slots.erase(c->slots.begin() + id); c->topo_dirty = true;
    return 0;
}
int gdg_chain_move_up(gdg_chain *c, int id)
{
    EditLock lk(g_edit_mutex);
    if (c) c->clean = false;
    CHAIN_CHECK(c);
    if (id <= 0 || id >= (int)c->slots.size()) return fail("Cannot move unit %d up.", id);
    std::swap(c->slots[id], c->slots[id - 1]); c->topo_dirty = true;
    return 0;
}
int gdg_chain_move_down(gdg_chain *c, int id)
{
    EditLock lk(g_edit_mutex);
    if (c) c->clean = false;
    CHAIN_CHECK(c);
    if (id < 0 || id >= (int)c->slots.size() - 1) return fail("Cannot move unit %d down.", id);
    std::swap(c->slots[id], c->slots[id + 1]); c->topo_dirty = true;
    return 0;
}
int gdg_chain_unit_type(gdg_chain *c, int id)
{
    ProcessLock lk(g_edit_mutex);
    if (!c) { fail("chain is NULL"); return -1; }
    Use use_ctx_(c->ctx);
    if (id < 0 || id >= (int)c->slots.size()) { fail("Cannot get unit type: No unit %d.", id); return -1; }
    return c->slots[id].unit->type;
}
int gdg_chain_set_bypass(gdg_chain *c, int id, int bypass)
{
    EditLock lk(g_edit_mutex);
    if (c) c->clean = false;
    CHAIN_CHECK(c);
    if (id < 0 || id >= (int)c->slots.size()) return fail("Cannot %s bypass: No unit %d.", bypass ? "enable" : "disable", id);
    if (c->slots[id].bypass != (bypass != 0)) { c->slots[id].bypass = bypass != 0; c->topo_dirty = true; }
    return 0;
}
int gdg_chain_get_bypass(gdg_chain *c, int id, int *bypass)
{
    ProcessLock lk(g_edit_mutex);
    CHAIN_CHECK(c);
    if (id < 0 || id >= (int)c->slots.size()) return fail("Cannot get bypass value: No unit %d.", id);
    if (!bypass) return fail("gdg_chain_get_bypass: bad argument");
    *bypass = c->slots[id].bypass;
    return 0;
}
int gdg_chain_set_discrete_value(gdg_chain *c, int id, const char *name, const char *value)
{
    EditLock lk(g_edit_mutex);
    if (c) c->clean = false;
    CHAIN_CHECK(c);
    if (id < 0 || id >= (int)c->slots.size()) return fail("Cannot set discrete value: No unit %d.", id);
    if (!name || !value) return fail("gdg_chain_set_discrete_value: bad argument");
    Unit *u = c->slots[id].unit.get();
    const int idx = u->find(name);
    if (idx < 0) return fail("Failed to set discrete value: Could not find parameter with name '%s'.", name);
    Param &p = u->params[idx];
    if (p.type != GDG_PARAMETER_TYPE_DISCRETE) return fail("Failed to set discrete value: Parameter '%s' is not discrete.", name);
    int v = -1;
    for (size_t i = 0; i < p.disc.size(); i++) if (p.disc[i] == value) v = (int)i;
    if (v < 0) return fail("Failed to set discrete value: Value '%s' is not valid for parameter '%s'.", value, name);
    p.didx = v; u->dirty = true;
    if (u->type == GDG_UNIT_POWERAMP) u->compile_poweramp(u->pa_rate);     /* poweramp.go:132-154 */
    return 0;
}
int gdg_chain_get_discrete_value(gdg_chain *c, int id, const char *name, const char **value)
{
    ProcessLock lk(g_edit_mutex);
    CHAIN_CHECK(c);
    if (id < 0 || id >= (int)c->slots.size()) return fail("Cannot get discrete value: No unit %d.", id);
    if (!name || !value) return fail("gdg_chain_get_discrete_value: bad argument");
    Unit *u = c->slots[id].unit.get();
    const int idx = u->find(name);
    if (idx < 0) return fail("Failed to get discrete value: Could not find parameter with name '%s'.", name);
    if (u->params[idx].type != GDG_PARAMETER_TYPE_DISCRETE) return fail("Failed to get discrete value: Parameter '%s' is not discrete.", name);
    *value = u->params[idx].disc[u->params[idx].didx].c_str();
    return 0;
}
int gdg_chain_set_numeric_value(gdg_chain *c, int id, const char *name, int32_t value)
{
    EditLock lk(g_edit_mutex);
    if (c) c->clean = false;
    CHAIN_CHECK(c);
    if (id < 0 || id >= (int)c->slots.size()) return fail("Cannot set numeric value: No unit %d.", id);
    if (!name) return fail("gdg_chain_set_numeric_value: bad argument");
    Unit *u = c->slots[id].unit.get();
    const int idx = u->find(name);
    if (idx < 0) return fail("Failed to set numeric value: Could not find parameter with name '%s'.", name);
    Param &p = u->params[idx];
    if (p.type != GDG_PARAMETER_TYPE_NUMERIC) return fail("Failed to set numeric value: Parameter '%s' is not numeric.", name);
    if (value < p.min || value > p.max)
        return fail("Failed to set numeric value: Parameter '%s' must be between '%d' and '%d' - got '%d'.", name, p.min, p.max, value);
    p.num = value; u->dirty = true;
    if (u->type == GDG_UNIT_POWERAMP) u->compile_poweramp(u->pa_rate);     /* poweramp.go:159-181 */
    return 0;
}
int gdg_chain_get_numeric_value(gdg_chain *c, int id, const char *name, int32_t *value)
{
    ProcessLock lk(g_edit_mutex);
    CHAIN_CHECK(c);
    if (id < 0 || id >= (int)c->slots.size()) return fail("Cannot get numeric value: No unit %d.", id);
    if (!name || !value) return fail("gdg_chain_get_numeric_value: bad argument");
    Unit *u = c->slots[id].unit.get();
    const int idx = u->find(name);
    if (idx < 0) return fail("Failed to get numeric value: Could not find parameter with name '%s'.", name);
    if (u->params[idx].type != GDG_PARAMETER_TYPE_NUMERIC) return fail("Failed to get numeric value: Parameter '%s' is not numeric.", name);
    *value = u->params[idx].num;
    return 0;
}
int gdg_chain_length(gdg_chain *c) { ProcessLock lk(g_edit_mutex); return c ? (int)c->slots.size() : 0; }
int gdg_chain_parameter_count(gdg_chain *c, int id)
{
    ProcessLock lk(g_edit_mutex);
    if (!c) { fail("chain is NULL"); return -1; }
    Use use_ctx_(c->ctx);
    if (id < 0 || id >= (int)c->slots.size()) { fail("Cannot get parameters: No unit %d.", id); return -1; }
    return (int)c->slots[id].unit->params.size();
}
int gdg_chain_parameter(gdg_chain *c, int id, int index, gdg_parameter *out)
{
    ProcessLock lk(g_edit_mutex);
    CHAIN_CHECK(c);
    if (id < 0 || id >= (int)c->slots.size()) return fail("Cannot get parameters: No unit %d.", id);
    Unit *u = c->slots[id].unit.get();
    if (index < 0 || index >= (int)u->params.size()) return fail("Cannot get parameter %d of unit %d.", index, id);
    if (!out) return fail("gdg_chain_parameter: bad argument");
    const Param &p = u->params[index];
    out->name = p.name.c_str(); out->type = p.type; out->physical_unit = p.phys.c_str();
    out->minimum = p.min; out->maximum = p.max; out->numeric_value = p.num;
    out->discrete_value_index = p.didx; out->discrete_value_count = (int32_t)p.disc.size();
    return 0;
}
const char *gdg_chain_parameter_discrete_value(gdg_chain *c, int id, int index, int vi)
{
    ProcessLock lk(g_edit_mutex);
    if (!c || id < 0 || id >= (int)c->slots.size()) return nullptr;
    Unit *u = c->slots[id].unit.get();
    if (index < 0 || index >= (int)u->params.size()) return nullptr;
    const Param &p = u->params[index];
    return (vi >= 0 && vi < (int)p.disc.size()) ? p.disc[vi].c_str() : nullptr;
}

/* Filter.Coefficients() of a power amp's compiled composite filter (filter.go:258-265): for inspection and tests */
long gdg_chain_filter_coefficients(gdg_chain *c, int id, double *out, size_t capacity)
{
    if (ensure_init()) return -1;
    if (!c) { fail("chain is NULL"); return -1; }
    Use use_ctx_(c->ctx);
    ProcessLock lk(g_edit_mutex);
    if (id < 0 || id >= (int)c->slots.size()) { fail("Cannot get filter coefficients: No unit %d.", id); return -1; }
    Unit *u = c->slots[id].unit.get();
    if (u->type != GDG_UNIT_POWERAMP) { fail("Cannot get filter coefficients: Unit %d is not a power amp.", id); return -1; }
    if (!u->pa_has_filter || !u->pa_filter) return 0;
    const size_t L = (size_t)u->pa_filter->L;
    if (out && capacity >= L && L && gdgb_download(out, u->pa_filter->d, L * sizeof(double))) { backend_fail("download failed"); return -1; }
    return (long)L;
}

int gdg_process_batch_device(gdg_chain *const *chains, size_t C, const double *d_in, double *d_out, size_t stride,
                             size_t n, uint32_t rate, void *stream)
{
    if (ensure_init()) return -1;
    if (!chains || !C || !d_in || !d_out) return fail("gdg_process_batch_device: bad argument");
    if (stride < n) return fail("gdg_process_batch_device: stride < n");
    for (size_t c_ = 1; c_ < (size_t)C; c_++) if (chains[c_] && chains[c_]->ctx != chains[0]->ctx) return fail("the chains of one call must live on one device (use gdg_cluster_* to span devices)");
    Use use_ctx_(chains[0]->ctx);
    return run_batch_device(chains, C, d_in, d_out, stride, n, rate, stream ? stream : gdgb_stream());
}

int gdg_process_batch(gdg_chain *const *chains, size_t C, const double *const *in, double *const *out, size_t n, uint32_t rate)
{
    if (ensure_init()) return -1;
    if (!chains || !C || !in || !out) return fail("gdg_process_batch: bad argument");
    if (check_block(n)) return -1;
    for (size_t c_ = 1; c_ < (size_t)C; c_++) if (chains[c_] && chains[c_]->ctx != chains[0]->ctx) return fail("the chains of one call must live on one device (use gdg_cluster_* to span devices)");
    Use use_ctx_(chains[0]->ctx);
    return host_block(chains, C, nullptr, -1, nullptr, nullptr, (const void *const *)in, GDG_FMT_IEEE64, (void *const *)out, GDG_FMT_IEEE64, C, n, rate);
}

int gdg_chain_process(gdg_chain *c, const double *in, double *out, size_t n, uint32_t rate)
{
    if (ensure_init()) return -1;
    CHAIN_CHECK(c);
    if (!in || !out) return fail("gdg_chain_process: bad argument");
    if (check_block(n)) return -1;
    ProcessLock el(g_edit_mutex);
    if (c->d_n < n) {
        gdgb_free(c->d_in); gdgb_free(c->d_out);
        c->d_in = (double *)gdgb_alloc(n * sizeof(double)); c->d_out = (double *)gdgb_alloc(n * sizeof(double));
        if (!c->d_in || !c->d_out) return backend_fail("device staging allocation failed");
        c->d_n = n;
    }
    if (gdgb_upload(c->d_in, in, n * sizeof(double))) return backend_fail("upload failed");
    /* single chain (the JACK path calls Chain.Process per channel): its own launch tables, no batch cache churn */
    if (!c->self) { c->self = new Batch(); c->self->chains.assign(1, c); }
    gdg_chain *one[1] = {c};
    const int rounds = prepare_tables(*c->self, one, 1, rate, n, gdgb_stream());
    if (rounds < 0) return -1;
    if (launch_tables(*c->self, 0, 1, rounds, c->d_in, n * sizeof(double), GDG_FMT_IEEE64, c->d_out, n * sizeof(double), GDG_FMT_IEEE64, n, gdgb_stream()))
        return -1;
    return gdgb_download(out, c->d_out, n * sizeof(double)) ? backend_fail("download failed") : 0;
}

/* ---- spatializer ---- */
gdg_spatializer *gdg_spatializer_create(uint32_t n)
{
    if (ensure_init()) return nullptr;
    gdg_spatializer *s = new gdg_spatializer();
    s->ctx = gdgb_current();
    s->count = n; s->pos.resize(n);
    if (s->alloc_hist(96000)) { delete s; return nullptr; }
    return s;
}
void gdg_spatializer_destroy(gdg_spatializer *s) { if (!s) return; Use use_ctx_(s->ctx); gdgb_sync(); delete s; }
#define SPAT_CH(s, ch, what) if (!(s)) return fail("spatializer is NULL"); Use use_ctx_((s)->ctx); \
    if ((ch) >= (s)->count) return fail("Cannot " what " for channel %d: Only %d channels exist.", (int)(ch), (int)(s)->count); \
    std::lock_guard<std::mutex> spat_lock_((s)->mu)
int gdg_spatializer_set_azimuth(gdg_spatializer *s, uint32_t ch, double v) { SPAT_CH(s, ch, "set azimuth"); s->pos[ch].az = v; s->dirty = true; return 0; }
int gdg_spatializer_get_azimuth(gdg_spatializer *s, uint32_t ch, double *v) { SPAT_CH(s, ch, "get azimuth"); if (!v) return fail("gdg_spatializer_get_azimuth: bad argument"); *v = s->pos[ch].az; return 0; }
int gdg_spatializer_set_distance(gdg_spatializer *s, uint32_t ch, double v)
{
    SPAT_CH(s, ch, "set distance");
    if (v < 0.0 || v > 10.0) return fail("%s", "Failed to set distance: Value must be within [0, 10].");
    s->pos[ch].dist = v; s->dirty = true; return 0;
}
int gdg_spatializer_get_distance(gdg_spatializer *s, uint32_t ch, double *v) { SPAT_CH(s, ch, "get distance"); if (!v) return fail("gdg_spatializer_get_distance: bad argument"); *v = s->pos[ch].dist; return 0; }
int gdg_spatializer_set_level(gdg_spatializer *s, uint32_t ch, double v)
{
    SPAT_CH(s, ch, "set distance");                         /* sic: spatializer.go:393 reuses the "distance" wording */
    if (v < 0.0 || v > 1.0) return fail("%s", "Failed to set level: Value must be within [0, 1].");
    s->pos[ch].level = v; s->dirty = true; return 0;
}
int gdg_spatializer_get_level(gdg_spatializer *s, uint32_t ch, double *v) { SPAT_CH(s, ch, "get level"); if (!v) return fail("gdg_spatializer_get_level: bad argument"); *v = s->pos[ch].level; return 0; }
uint32_t gdg_spatializer_input_count(gdg_spatializer *s) { return s ? s->count : 0; }
uint32_t gdg_spatializer_output_count(gdg_spatializer *) { return 2; }
int gdg_spatializer_set_sample_rate(gdg_spatializer *s, uint32_t rate)
{
    if (!s) return fail("spatializer is NULL");
    Use use_ctx_(s->ctx);
    std::lock_guard<std::mutex> lk(s->mu);
    gdgb_sync();
    return s->alloc_hist(rate);
}
int gdg_spatializer_process_device(gdg_spatializer *s, const double *d_in, size_t stride, const double *d_aux,
                                   double *d_out, size_t out_stride, size_t n, int accumulate, void *stream)
{
    if (ensure_init()) return -1;
    if (!s || !d_in || !d_out) return fail("gdg_spatializer_process_device: bad argument");
    Use use_ctx_(s->ctx);
    if (n == 0 || n > (size_t)1 << 24) return fail("gdg_spatializer_process_device: bad block length");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->dirty) { if (s->upload_channels()) return -1; if (stream && stream != gdgb_stream() && gdgb_sync()) return backend_fail("synchronize failed"); }
    if (s->count > 32 && s->part_n < n) {
        gdgb_sync(); gdgb_free(s->d_part);
        s->d_part = (double *)gdgb_alloc(2 * ((size_t)(s->count + 31) / 32) * n * sizeof(double));
        if (!s->d_part) return backend_fail("spatializer scratch allocation failed");
        s->part_n = n;
        if (stream && stream != gdgb_stream() && gdgb_sync()) return backend_fail("synchronize failed");
    }
    if (gdgb_launch_spatializer(s->d_ch, (int)s->count, d_in, (long)stride, s->d_hist[s->flip], s->d_hist[s->flip ^ 1], s->hs,
                                d_aux, d_out, d_out + out_stride, (int)n, accumulate, s->d_part, stream ? stream : gdgb_stream()))
        return backend_fail("spatializer kernel launch failed");
    s->flip ^= 1;
    return 0;
}
int gdg_spatializer_process(gdg_spatializer *s, const double *const *in, uint32_t n_in, const double *aux,
                            double *const *out, uint32_t n_out, size_t n)
{
    if (ensure_init()) return -1;
    if (!s || !in || !out) return fail("gdg_spatializer_process: bad argument");
    Use use_ctx_(s->ctx);
    if (n_in != s->count || n_out != 2) return 0;           /* silent no-op like spatializer.go:149 */
    if (s->cap_n < n) {
        gdgb_free(s->d_in); gdgb_free(s->d_out); gdgb_free(s->d_aux);
        s->d_in = (double *)gdgb_alloc(std::max<size_t>(1, s->count) * n * sizeof(double));
        s->d_out = (double *)gdgb_alloc(2 * n * sizeof(double));
        s->d_aux = (double *)gdgb_alloc(n * sizeof(double));
        if (!s->d_in || !s->d_out || !s->d_aux) return backend_fail("spatializer staging allocation failed");
        s->cap_n = n;
    }
    if (upload_simple(in, s->count, n, s->d_in)) return -1;
    if (aux && gdgb_upload(s->d_aux, aux, n * sizeof(double))) return backend_fail("upload failed");
    if (gdg_spatializer_process_device(s, s->d_in, n, aux ? s->d_aux : nullptr, s->d_out, n, n, 0, nullptr)) return -1;
    return download_simple(out, 2, n, s->d_out);
}

/* ---- tuner ---- */
gdg_tuner *gdg_tuner_create(void)
{
    if (ensure_init()) return nullptr;
    gdg_tuner *t = new gdg_tuner();
    t->ctx = gdgb_current();
    t->d_ring = (double *)gdgb_alloc(gdg_tuner::N * sizeof(double));
    t->d_work = (double *)gdgb_alloc(2 * 262144 * sizeof(double));
    if (!t->d_ring || !t->d_work) { backend_fail("tuner allocation failed"); delete t; return nullptr; }
    return t;
}
void gdg_tuner_destroy(gdg_tuner *t) { if (!t) return; Use use_ctx_(t->ctx); gdgb_sync(); delete t; }
/* circular.Enqueue (circular/circular.go:33-71) on the device ring */
int gdg_tuner_process_device(gdg_tuner *t, const double *d, size_t n, uint32_t rate, void *stream)
{
    if (ensure_init()) return -1;
    if (!t || (!d && n)) return fail("gdg_tuner_process_device: bad argument");
    Use use_ctx_(t->ctx);
    const size_t N = gdg_tuner::N;
    void *st = stream ? stream : gdgb_stream();
    int rc = 0;
    std::lock_guard<std::mutex> lk(g_tuner_mutex);
    if (n >= N) { rc = gdgb_copy_on(t->d_ring, d + (n - N), N * sizeof(double), st); t->pos = 0; }
    else if (n > 0) {
        const size_t inc = (size_t)t->pos + n;
        if (inc < N) { rc = gdgb_copy_on(t->d_ring + t->pos, d, n * sizeof(double), st); t->pos = (int)inc; }
        else {
            const size_t head = inc - N, tail = N - (size_t)t->pos;
            rc = gdgb_copy_on(t->d_ring + t->pos, d, tail * sizeof(double), st);
            if (!rc && head) rc = gdgb_copy_on(t->d_ring, d + tail, head * sizeof(double), st);
            t->pos = (int)head;
        }
    }
    t->rate = rate;
    return rc ? backend_fail("tuner ring copy failed") : 0;
}
int gdg_tuner_process(gdg_tuner *t, const double *s, size_t n, uint32_t rate)
{
    if (ensure_init()) return -1;
    if (!t || (!s && n)) return fail("gdg_tuner_process: bad argument");
    Use use_ctx_(t->ctx);
    const size_t N = gdg_tuner::N;
    if (n > N) { s += n - N; n = N; }                       /* only the tail can survive the enqueue */
    if (t->block_cap < n) { gdgb_free(t->d_block); t->d_block = (double *)gdgb_alloc(std::max<size_t>(1, n) * sizeof(double)); t->block_cap = n; if (!t->d_block) return backend_fail("tuner staging allocation failed"); }
    if (n && gdgb_upload(t->d_block, s, n * sizeof(double))) return backend_fail("upload failed");
    return gdg_tuner_process_device(t, t->d_block, n == N ? N : n, rate, nullptr);
}
/* tuner.Analyze: tuner/tuner.go:379-577 - the 262144-point autocorrelation runs on the device (all tuners of a batch in
 * one set of launches); the scalar peak pick / note lookup below is the control-plane tail of the call. */
static int tuner_pick(gdg_tuner *t, uint32_t rate, gdg_tuner_result *res)
{
    static const char *names[GDG_TUNER_NOTE_COUNT] = GDG_TUNER_NOTE_NAMES_INIT;
    static const double freqs[GDG_TUNER_NOTE_COUNT] = GDG_TUNER_NOTE_FREQS_INIT;
    const long n = gdg_tuner::N;
    const uint64_t two_n = 2 * (uint64_t)n;
    const int fft_size = (int)next_pow2(two_n);
    const double sr = (double)rate;
    long low = (long)((sr / freqs[GDG_TUNER_NOTE_COUNT - 1]) + 0.5);
    if (low < 0 || (uint64_t)low >= two_n) low = 0;
    long high = (long)((sr / freqs[0]) + 0.5);
    if (high < 0 || (uint64_t)high >= two_n) high = (long)(two_n - 1);
    const long lo_fetch = std::max<long>(0, low - 1), hi_fetch = std::min<long>((long)fft_size, std::max(high, low) + 2);
    t->corr.assign((size_t)(hi_fetch - lo_fetch), 0.0);
    if (gdgb_download(t->corr.data(), t->d_work + lo_fetch, t->corr.size() * sizeof(double))) return backend_fail("download failed");
    auto corr = [&](long i) { return t->corr[(size_t)(i - lo_fetch)]; };
    double max_val = -INFINITY; long max_idx = -1;
    for (long i = low; i < high; i++) if (corr(i) > max_val) { max_val = corr(i); max_idx = i - low; }
    const long idx = low + max_idx;
    long up = idx + 1; if (up > n) up = n;
    long down = idx - 1; if (down < 0) down = 0;
    down = std::max(down, lo_fetch); up = std::min(up, hi_fetch - 1);
    const double left = corr(down), right = corr(up);
    double shift = (0.5 * (right - left)) / ((2.0 * max_val) - (right + left));
    if (shift < -0.5) shift = -0.5; else if (shift > 0.5) shift = 0.5;
    const double freq = sr / ((double)idx + shift);
    const char *note = "Unknown"; double cents = INFINITY, cents_abs = INFINITY;
    for (int k = 0; k < GDG_TUNER_NOTE_COUNT; k++) {
        const double dc = 1200.0 * std::log2(freq / freqs[k]);
        if (std::fabs(dc) < cents_abs) { note = names[k]; cents = dc; cents_abs = std::fabs(dc); }
    }
    res->cents = (std::isinf(cents) || std::isnan(cents)) ? 0 : (int8_t)cents;
    res->frequency = freq; res->note = note;
    return 0;
}
int gdg_tuner_analyze_batch(gdg_tuner *const *tuners, uint32_t count, gdg_tuner_result *results)
{
    if (ensure_init()) return -1;
    if ((!tuners || !results) && count) return fail("gdg_tuner_analyze_batch: bad argument");
    std::vector<const double *> rings(count); std::vector<double *> works(count); std::vector<int> pos(count); std::vector<uint32_t> rates(count);
    for (uint32_t i = 0; i < count; i++) {
        if (!tuners[i]) return fail("gdg_tuner_analyze_batch: tuner %u is NULL", i);
        if (tuners[i]->ctx != tuners[0]->ctx) return fail("gdg_tuner_analyze_batch: the tuners of one batch must live on one device");
    }
    Use use_ctx_(count ? tuners[0]->ctx : gdgb_current());
    const int fft_size = (int)next_pow2(2 * (uint64_t)gdg_tuner::N);
    std::lock_guard<std::mutex> analyses(g_analyze_mutex);
    {
        /* A block being enqueued feeds its tuner on one of the two kernel streams and joins them at its end (block_submit):
         * with the batch lock held no block is half enqueued, so the library stream - where the analysis runs - is ordered
         * after every ring copy enqueued so far, and later blocks fork from it, i.e. run after the analysis has read the ring */
        std::lock_guard<std::mutex> blocks(g_batch_mutex);
        std::lock_guard<std::mutex> lk(g_tuner_mutex);
        for (uint32_t i = 0; i < count; i++) { rings[i] = tuners[i]->d_ring; works[i] = tuners[i]->d_work; pos[i] = tuners[i]->pos; rates[i] = tuners[i]->rate; }
        if (count && gdgb_autocorrelate_batch(rings.data(), gdg_tuner::N, pos.data(), works.data(), (int)count, fft_size, gdgb_stream()))
            return backend_fail("autocorrelation failed");
    }
    for (uint32_t i = 0; i < count; i++) if (tuner_pick(tuners[i], rates[i], &results[i])) return -1;
    return 0;
}
int gdg_tuner_analyze(gdg_tuner *t, gdg_tuner_result *res)
{
    if (!t || !res) return fail("gdg_tuner_analyze: bad argument");
    return gdg_tuner_analyze_batch(&t, 1, res);
}

/* ---- encoded batch (wave/wave.go sample formats fused into the kernel's load / store) ---- */
size_t gdg_format_bytes_per_sample(int format) { return fmt_bytes(format); }
int gdg_process_batch_encoded(gdg_chain *const *chains, size_t C, const void *const *in, int in_format, void *const *out, int out_format,
                              size_t n, uint32_t rate)
{
    if (ensure_init()) return -1;
    if (!chains || !C || !in || !out) return fail("gdg_process_batch_encoded: bad argument");
    if (!fmt_bytes(in_format) || !fmt_bytes(out_format)) return fail("Unknown sample format: %#04x", fmt_bytes(in_format) ? out_format : in_format);
    if (check_block(n)) return -1;
    for (size_t c_ = 1; c_ < (size_t)C; c_++) if (chains[c_] && chains[c_]->ctx != chains[0]->ctx) return fail("the chains of one call must live on one device (use gdg_cluster_* to span devices)");
    Use use_ctx_(chains[0]->ctx);
    return host_block(chains, C, nullptr, -1, nullptr, nullptr, in, in_format, out, out_format, C, n, rate);
}
/* the same without waiting: ticket for gdg_controller_wait (two blocks in flight, like gdg_controller_submit) */
int gdg_process_batch_encoded_submit(gdg_chain *const *chains, size_t C, const void *const *in, int in_format, void *const *out,
                                     int out_format, size_t n, uint32_t rate)
{
    if (ensure_init()) return -1;
    if (!chains || !C || !in || !out) return fail("gdg_process_batch_encoded_submit: bad argument");
    if (!fmt_bytes(in_format) || !fmt_bytes(out_format)) return fail("Unknown sample format: %#04x", fmt_bytes(in_format) ? out_format : in_format);
    if (check_block(n)) return -1;
    for (size_t c_ = 1; c_ < (size_t)C; c_++) if (chains[c_] && chains[c_]->ctx != chains[0]->ctx) return fail("the chains of one call must live on one device (use gdg_cluster_* to span devices)");
    Use use_ctx_(chains[0]->ctx);
    ProcessLock el(g_edit_mutex);
    std::lock_guard<std::mutex> lk(g_batch_mutex);
    const int si = free_slot();
    if (si < 0) return fail("two submitted blocks are in flight: gdg_controller_wait() for one first");
    if (block_submit(si, chains, C, nullptr, -1, nullptr, nullptr, in, in_format, out, out_format, C, n, rate)) {
        if (g_batch) { gdgb_sync(); g_batch->slot[si].busy = false; }
        return -1;
    }
    return si;
}
int gdg_process_batch_device_encoded(gdg_chain *const *chains, size_t C, const void *d_in, int in_format, size_t in_stride_bytes,
                                     void *d_out, int out_format, size_t out_stride_bytes, size_t n, uint32_t rate, void *stream)
{
    if (ensure_init()) return -1;
    if (!chains || !C || !d_in || !d_out) return fail("gdg_process_batch_device_encoded: bad argument");
    if (!fmt_bytes(in_format) || !fmt_bytes(out_format)) return fail("Unknown sample format: %#04x", fmt_bytes(in_format) ? out_format : in_format);
    if (in_stride_bytes < n * fmt_bytes(in_format) || out_stride_bytes < n * fmt_bytes(out_format)) return fail("gdg_process_batch_device_encoded: stride too small");
    /* the float formats are read and written as float / double words: every row must start on a multiple of the sample size
     * (the integer formats go byte by byte and may sit anywhere) */
    {
        const size_t ia = in_format == GDG_FMT_IEEE64 ? 8 : (in_format == GDG_FMT_IEEE32 ? 4 : 1), oa = out_format == GDG_FMT_IEEE64 ? 8 : (out_format == GDG_FMT_IEEE32 ? 4 : 1);
        if ((uintptr_t)d_in % ia || in_stride_bytes % ia || (uintptr_t)d_out % oa || out_stride_bytes % oa)
            return fail("gdg_process_batch_device_encoded: rows of a float format must be aligned to the sample size");
    }
    if (check_block(n)) return -1;
    for (size_t c_ = 1; c_ < (size_t)C; c_++) if (chains[c_] && chains[c_]->ctx != chains[0]->ctx) return fail("the chains of one call must live on one device (use gdg_cluster_* to span devices)");
    Use use_ctx_(chains[0]->ctx);
    ProcessLock el(g_edit_mutex);
    std::lock_guard<std::mutex> lk(g_batch_mutex);
    void *st = stream ? stream : gdgb_stream();
    const int rounds = prepare_batch(chains, C, rate, n, st);
    if (rounds < 0) return -1;
    return launch_batch_fmt(0, C, rounds, d_in, in_stride_bytes, in_format, d_out, out_stride_bytes, out_format, n, st);
}

/* ---- resample.Time (resample/resample.go:72-103) ---- */
size_t gdg_resample_time_length(size_t n, uint32_t src, uint32_t dst) { return resample_time_length(n, src, dst); }
long gdg_resample_time(const double *in, size_t n, uint32_t src, uint32_t dst, double *out, size_t out_capacity)
{
    if (ensure_init()) return -1;
    if ((!in && n) || !out || !src || !dst) return fail("gdg_resample_time: bad argument");
    const size_t len = gdg_resample_time_length(n, src, dst);
    if (out_capacity < len) return fail("gdg_resample_time: output buffer holds %zu samples, %zu needed", out_capacity, len);
    double *d_in = (double *)gdgb_alloc(std::max<size_t>(1, n) * sizeof(double)), *d_out = (double *)gdgb_alloc(std::max<size_t>(1, len) * sizeof(double));
    int rc = (!d_in || !d_out) ? backend_fail("allocation failed") : 0;
    if (!rc && n && gdgb_upload(d_in, in, n * sizeof(double))) rc = backend_fail("upload failed");
    if (!rc && gdgb_resample_time(d_in, (long)n, (double)src / (double)dst, d_out, (long)len, gdgb_stream())) rc = backend_fail("resample kernel launch failed");
    if (!rc && gdgb_download(out, d_out, len * sizeof(double))) rc = backend_fail("download failed");
    gdgb_sync(); gdgb_free(d_in); gdgb_free(d_out);
    return rc ? -1 : (long)len;
}

/* ---- level.Meter ---- */
gdg_meter *gdg_meter_create(uint32_t channels, const char *const *names)
{
    if (ensure_init()) return nullptr;
    gdg_meter *m = new gdg_meter();
    m->ctx = gdgb_current();
    m->count = channels;
    for (uint32_t c = 0; c < channels; c++) m->names.push_back(names && names[c] ? names[c] : "");
    m->d_state = (double *)gdgb_alloc(std::max<size_t>(1, channels) * 4 * sizeof(double));
    if (!m->d_state) { backend_fail("level meter allocation failed"); delete m; return nullptr; }
    return m;
}
void gdg_meter_destroy(gdg_meter *m) { if (!m) return; Use use_ctx_(m->ctx); gdgb_sync(); delete m; }
uint32_t gdg_meter_channel_count(gdg_meter *m) { return m ? m->count : 0; }
const char *gdg_meter_channel_name(gdg_meter *m, uint32_t channel)
{
    if (!m) { fail("gdg_meter_channel_name: bad argument"); return nullptr; }
    if (channel >= m->count) { fail("Requested name of channel %u, but level meter only has %u channels.", channel, m->count); return nullptr; }
    return m->names[channel].c_str();
}
int gdg_meter_enabled(gdg_meter *m) { return m && m->enabled ? 1 : 0; }
int gdg_meter_set_enabled(gdg_meter *m, int value)
{
    if (ensure_init()) return -1;
    if (!m) return fail("gdg_meter_set_enabled: bad argument");
    Use use_ctx_(m->ctx);
    const bool v = value != 0;
    if (v != m->enabled) {
        if (!v && gdgb_zero(m->d_state, std::max<size_t>(1, m->count) * 4 * sizeof(double))) return backend_fail("level meter reset failed");   /* level.go:226-230 */
        m->enabled = v;
    }
    return 0;
}
int gdg_meter_process_device(gdg_meter *m, uint32_t first_channel, const double *d_rows, size_t stride, uint32_t rows, size_t n,
                             uint32_t rate, void *stream)
{
    if (ensure_init()) return -1;
    if (!m || (!d_rows && rows && n) || !rate) return fail("gdg_meter_process_device: bad argument");
    Use use_ctx_(m->ctx);
    if ((uint64_t)first_channel + rows > m->count) return fail("gdg_meter_process_device: channels %u..%u exceed the meter's %u channels", first_channel, first_channel + rows, m->count);
    return meter_rows(m, first_channel, d_rows, stride, rows, n, rate, stream);
}
int gdg_meter_process(gdg_meter *m, const double *const *bufs, uint32_t nbufs, size_t n, uint32_t rate)
{
    if (ensure_init()) return -1;
    if (!m || (!bufs && nbufs)) return fail("gdg_meter_process: bad argument");
    Use use_ctx_(m->ctx);
    if (nbufs != m->count) return fail("Number of input buffers (%u) does not match number of channels (%u) for this level meter.", nbufs, m->count);   /* level.go:311-312 */
    if (!m->enabled || !n || !nbufs) return 0;
    if (!rate) return fail("gdg_meter_process: bad sample rate");
    const size_t need = (size_t)nbufs * n;
    if (m->stage_cap < need) { gdgb_free(m->d_stage); m->d_stage = (double *)gdgb_alloc(need * sizeof(double)); m->stage_cap = m->d_stage ? need : 0; if (!m->d_stage) return backend_fail("level meter staging allocation failed"); }
    /* rows that are NULL (a nil buffer in the reference: zero samples) are skipped; runs of present rows go up together */
    uint32_t c = 0;
    while (c < nbufs) {
        if (!bufs[c]) { c++; continue; }
        uint32_t e = c;
        while (e < nbufs && bufs[e]) { if (gdgb_upload(m->d_stage + (size_t)e * n, bufs[e], n * sizeof(double))) return backend_fail("upload failed"); e++; }
        if (meter_rows(m, c, m->d_stage + (size_t)c * n, n, e - c, n, rate, nullptr)) return -1;
        c = e;
    }
    return 0;
}
int gdg_meter_state(gdg_meter *m, uint32_t channel, double *cur, double *peak, uint64_t *counter)
{
    if (ensure_init()) return -1;
    if (!m) return fail("gdg_meter_state: bad argument");
    Use use_ctx_(m->ctx);
    if (channel >= m->count) return fail("Requested analysis for channel %u, but level meter only has %u channels.", channel, m->count);
    double w[4];
    if (gdgb_sync_stream(0) || gdgb_download(w, m->d_state + 4 * (size_t)channel, sizeof(w))) return backend_fail("download failed");
    if (cur) *cur = w[0];
    if (peak) *peak = w[1];
    if (counter) memcpy(counter, &w[2], sizeof(uint64_t));
    return 0;
}
/* level.go:95-137: 20 log10, NaN / below -200 dB -> -200, round half away from zero */
int gdg_meter_analyze(gdg_meter *m, uint32_t channel, int32_t *level, int32_t *peak)
{
    double c = 0.0, p = 0.0;
    if (gdg_meter_state(m, channel, &c, &p, nullptr)) return -1;
    auto db = [](double v) { double l = 20.0 * std::log10(v); if (std::isnan(l) || l < -200.0) l = -200.0; return (int32_t)std::round(l); };
    if (level) *level = db(c);
    if (peak) *peak = db(p);
    return 0;
}

/* ---- controller.process mirror ---- */
namespace {
/* One block of controller.process (controller/controller.go:2648-2783) on HOST rows; `wait` selects the blocking call or the
 * ticketed submit.  Returns 0 / the ticket, or -1. */
int controller_block(const char *who, gdg_chain *const *chains, uint32_t C, gdg_tuner *tuner, int tuner_channel, gdg_spatializer *spat,
                     gdg_meter *meter, const void *const *in, int in_fmt, const double *metronome, int metro_master,
                     void *const *out, int out_fmt, uint32_t n_out, size_t n, uint32_t rate, bool wait)
{
    if (ensure_init()) return -1;
    if (!chains || !C || !in || !out) return fail("%s: bad argument", who);
    if (!fmt_bytes(in_fmt) || !fmt_bytes(out_fmt)) return fail("Unknown sample format: %#04x", fmt_bytes(in_fmt) ? out_fmt : in_fmt);
    if (check_block(n)) return -1;
    for (size_t c_ = 1; c_ < (size_t)C; c_++) if (chains[c_] && chains[c_]->ctx != chains[0]->ctx) return fail("the chains of one call must live on one device (use gdg_cluster_* to span devices)");
    Use use_ctx_(chains[0]->ctx);
    if (meter && meter->count != 2 * C + 3)
        return fail("%s: the level meter has %u channels, %u inputs need %u (inputs, outputs, metronome, master L/R: controller.go:2710-2771)", who, meter->count, C, 2 * C + 3);
    if (n_out < C) {
        /* controller.go:2666-2682: the tuner is fed whatever the number of outputs; the chains only run when nOut >= nIn */
        if (!wait) return fail("%s: %u output rows for %u channels", who, n_out, C);
        if (tuner && tuner_channel >= 0 && (uint32_t)tuner_channel < C && in_fmt == GDG_FMT_IEEE64)
            return gdg_tuner_process(tuner, (const double *)in[tuner_channel], n, rate);
        return 0;
    }
    /* spatializer (and the metronome's way into the master) only with nOut >= nIn + 2 + 1 (controller.go:2650, 2722); callers
     * that pass no metronome row get the master pair with nIn + 2 rows as before */
    const bool do_spat = spat && n_out >= C + 2 && spat->count == C;
    gdg_meter *m = meter && meter->enabled ? meter : nullptr;
    if (wait)
        return host_block(chains, C, tuner, tuner_channel, do_spat ? spat : nullptr, m, in, in_fmt, out, out_fmt, (size_t)C + (do_spat ? 2 : 0), n, rate,
                          metronome, metro_master != 0);
    ProcessLock el(g_edit_mutex);
    std::lock_guard<std::mutex> lk(g_batch_mutex);
    const int si = free_slot();
    if (si < 0) return fail("two submitted blocks are in flight: gdg_controller_wait() for one first");
    if (block_submit(si, chains, C, tuner, tuner_channel, do_spat ? spat : nullptr, m, in, in_fmt, out, out_fmt, (size_t)C + (do_spat ? 2 : 0), n, rate,
                     metronome, metro_master != 0)) {
        if (g_batch) { gdgb_sync(); g_batch->slot[si].busy = false; }
        return -1;
    }
    return si;
}
} // namespace

int gdg_controller_process_metered(gdg_chain *const *chains, uint32_t C, gdg_tuner *tuner, int tuner_channel, gdg_spatializer *spat,
                                   gdg_meter *meter, const double *const *in, double *const *out, uint32_t n_out, size_t n, uint32_t rate)
{
    return controller_block("gdg_controller_process", chains, C, tuner, tuner_channel, spat, meter, (const void *const *)in, GDG_FMT_IEEE64, nullptr, 0,
                            (void *const *)out, GDG_FMT_IEEE64, n_out, n, rate, true);
}
int gdg_controller_process_aux(gdg_chain *const *chains, uint32_t C, gdg_tuner *tuner, int tuner_channel, gdg_spatializer *spat,
                               gdg_meter *meter, const double *const *in, const double *metronome, int metronome_to_master,
                               double *const *out, uint32_t n_out, size_t n, uint32_t rate)
{
    return controller_block("gdg_controller_process_aux", chains, C, tuner, tuner_channel, spat, meter, (const void *const *)in, GDG_FMT_IEEE64, metronome,
                            metronome_to_master, (void *const *)out, GDG_FMT_IEEE64, n_out, n, rate, true);
}
/* the same block, submitted without waiting: returns a ticket (0 or 1) for gdg_controller_wait */
int gdg_controller_submit(gdg_chain *const *chains, uint32_t C, gdg_tuner *tuner, int tuner_channel, gdg_spatializer *spat,
                          gdg_meter *meter, const double *const *in, double *const *out, uint32_t n_out, size_t n, uint32_t rate)
{
    return controller_block("gdg_controller_submit", chains, C, tuner, tuner_channel, spat, meter, (const void *const *)in, GDG_FMT_IEEE64, nullptr, 0,
                            (void *const *)out, GDG_FMT_IEEE64, n_out, n, rate, false);
}
int gdg_controller_submit_aux(gdg_chain *const *chains, uint32_t C, gdg_tuner *tuner, int tuner_channel, gdg_spatializer *spat,
                              gdg_meter *meter, const double *const *in, const double *metronome, int metronome_to_master,
                              double *const *out, uint32_t n_out, size_t n, uint32_t rate)
{
    return controller_block("gdg_controller_submit_aux", chains, C, tuner, tuner_channel, spat, meter, (const void *const *)in, GDG_FMT_IEEE64, metronome,
                            metronome_to_master, (void *const *)out, GDG_FMT_IEEE64, n_out, n, rate, false);
}
/* gdg_controller_submit with encoded host rows (wave.go sample formats): the payload is decoded / encoded on the device,
 * tuner, spatializer and meter see float64 */
int gdg_controller_submit_encoded(gdg_chain *const *chains, uint32_t C, gdg_tuner *tuner, int tuner_channel, gdg_spatializer *spat,
                                  gdg_meter *meter, const void *const *in, int in_format, void *const *out, int out_format,
                                  uint32_t n_out, size_t n, uint32_t rate)
{
    return controller_block("gdg_controller_submit_encoded", chains, C, tuner, tuner_channel, spat, meter, in, in_format, nullptr, 0, out, out_format,
                            n_out, n, rate, false);
}
int gdg_controller_wait(int ticket)
{
    if (ensure_init()) return -1;
    std::lock_guard<std::mutex> lk(g_batch_mutex);
    if (ticket < 0 || ticket > 1 || !g_batch || !g_batch->slot[ticket].busy) return fail("gdg_controller_wait: ticket %d is not in flight", ticket);
    return block_wait(ticket);
}
int gdg_controller_process(gdg_chain *const *chains, uint32_t C, gdg_tuner *tuner, int tuner_channel, gdg_spatializer *spat,
                           const double *const *in, double *const *out, uint32_t n_out, size_t n, uint32_t rate)
{
    return gdg_controller_process_metered(chains, C, tuner, tuner_channel, spat, nullptr, in, out, n_out, n, rate);
}


/* ---- several GPUs in ONE process (controller.go:3339-3341: the reference host is one process, one worker per channel) ---- */
int gdg_device_count(void) { return gdgb_device_count(); }
int gdg_use_device(int device)
{
    if (ensure_init()) return -1;
    const int ctx = gdgb_context_open(device);
    if (ctx < 0) return backend_fail("gdg_use_device");
    if (gdgb_select(ctx)) return backend_fail("gdg_use_device");
    return ctx;
}
int gdg_current_device(void) { if (ensure_init()) return -1; return gdgb_context_device(gdgb_current()); }
} /* extern "C" */

struct gdg_cluster {
    std::vector<int> ctx;                                   /* member -> device context */
    /* on member 0, one set per block in flight: the summed [2][n] master, its encoded copy, the metronome row, the pinned
     * staging of the master rows */
    double *d_master[2] = {nullptr, nullptr}, *d_master_enc[2] = {nullptr, nullptr}, *d_metro[2] = {nullptr, nullptr}, *h_master[2] = {nullptr, nullptr};
    size_t cap_n = 0;
    struct Flight { bool busy = false; std::vector<int> members; void *out_l = nullptr, *out_r = nullptr; size_t ob = 0; bool master = false; } flight[2];
    void release()
    {
        for (int k = 0; k < 2; k++) {
            gdgb_free(d_master[k]); gdgb_free(d_master_enc[k]); gdgb_free(d_metro[k]); gdgb_pinned_free(h_master[k]);
            d_master[k] = d_master_enc[k] = d_metro[k] = h_master[k] = nullptr;
        }
        cap_n = 0;
    }
    ~gdg_cluster()
    {
        if (!ctx.empty()) { Use u(ctx[0]); gdgb_sync(); release(); }
    }
};

namespace {
/* One block of controller.process (controller.go:2648-2783) over the cluster: chains[c] live on the members' devices in
 * member order (contiguous channel ranges); spats[m] is member m's spatializer over ITS channels (or spats ==
 * NULL: no master); meters[m] is member m's level meter with 2 * cnt_m + 3 channels laid out like the one-device call's
 * (inputs, outputs, metronome port, master L / R - the last three are fed on member 0 only, which owns the summed master).
 * Each member runs tuner feed -> chains -> spatializer (-> meter) on its own streams; member 0 then sums the partial stereo
 * masters (+ the metronome row, spatializer.go:300-308) with one kernel that reads the other devices' buffers through peer
 * (NVLink) addresses and the master rows out[C], out[C+1] come back from member 0.  Host rows in any wave.go sample format.
 * Returns a ticket for gdg_cluster_wait(). */
int cluster_block(const char *who, gdg_cluster *cl, gdg_chain *const *chains, uint32_t C, gdg_tuner *tuner, int tuner_channel,
                  gdg_spatializer *const *spats, gdg_meter *const *meters, const void *const *in, int in_fmt, const double *metronome,
                  int metro_master, void *const *out, int out_fmt, uint32_t n_out, size_t n, uint32_t rate)
{
    if (ensure_init()) return -1;
    if (!cl || !chains || !C || !in || !out) return fail("%s: bad argument", who);
    if (!fmt_bytes(in_fmt) || !fmt_bytes(out_fmt)) return fail("Unknown sample format: %#04x", fmt_bytes(in_fmt) ? out_fmt : in_fmt);
    if (check_block(n)) return -1;
    if (n_out < C) return fail("%s: %u output rows for %u channels", who, n_out, C);
    const int M = (int)cl->ctx.size();
    std::vector<uint32_t> first(M + 1, C);
    {   /* channel ranges: the chains' device contexts must follow the member order */
        int m = 0; first[0] = 0;
        for (uint32_t c = 0; c < C; c++) {
            if (!chains[c]) return fail("%s: chain %u is NULL", who, c);
            while (m < M && chains[c]->ctx != cl->ctx[m]) { m++; if (m < M) first[m] = c; }
            if (m >= M) return fail("%s: chain %u does not live on a member device (or the chains are not grouped in member order)", who, c);
        }
        for (int k = m + 1; k <= M; k++) first[k] = C;
    }
    const bool master = spats && n_out >= C + 2;
    for (int m = 0; m < M; m++) {                           /* every check before anything is enqueued */
        const uint32_t cnt = first[m + 1] - first[m];
        gdg_spatializer *sp = master ? spats[m] : nullptr;
        if (cnt && sp && (sp->count != cnt || sp->ctx != cl->ctx[m]))
            return fail("%s: member %d's spatializer must cover its %u channels on its own device", who, m, cnt);
        gdg_meter *mt = meters ? meters[m] : nullptr;
        if (mt && (mt->count != 2 * cnt + 3 || mt->ctx != cl->ctx[m]))
            return fail("%s: member %d's level meter must have %u channels (its %u inputs, %u outputs, metronome, master L/R) on its own device", who, m, 2 * cnt + 3, cnt, cnt);
    }
    gdg_meter *meter0 = (meters && meters[0] && meters[0]->enabled) ? meters[0] : nullptr;
    const uint32_t cnt0 = first[1] - first[0];
    int si = -1;
    {
        std::lock_guard<std::mutex> lk(g_batch_mutex);
        for (int k = 0; k < 2 && si < 0; k++) {
            bool ok = !cl->flight[k].busy;
            for (int m = 0; m < M && ok; m++) { Use u(cl->ctx[m]); if (g_batch && g_batch->slot[k].busy) ok = false; }
            if (ok) si = k;
        }
    }
    if (si < 0) return fail("two submitted blocks are in flight: gdg_cluster_wait() for one first");
    gdg_cluster::Flight &fl = cl->flight[si];
    const size_t ob = n * fmt_bytes(out_fmt);
    fl.members.clear(); fl.master = master; fl.ob = ob;
    auto unwind = [&]() {
        for (int k : fl.members) { Use u2(cl->ctx[k]); gdgb_sync(); if (g_batch) g_batch->slot[si].busy = false; }
        fl.members.clear();
        return -1;
    };
    std::vector<const double *> parts; std::vector<long> strides; std::vector<int> part_ctx;
    for (int m = 0; m < M; m++) {
        const uint32_t c0 = first[m], cnt = first[m + 1] - first[m];
        if (!cnt) continue;
        Use u(cl->ctx[m]);
        ProcessLock el(g_edit_mutex);
        std::lock_guard<std::mutex> lk(g_batch_mutex);
        gdg_spatializer *sp = master ? spats[m] : nullptr;
        gdg_meter *mt = (meters && meters[m] && meters[m]->enabled) ? meters[m] : nullptr;
        gdg_tuner *tu = (tuner && tuner_channel >= (int)c0 && tuner_channel < (int)(c0 + cnt)) ? tuner : nullptr;
        const double *pm = nullptr;
        if (block_submit(si, chains + c0, cnt, tu, tu ? tuner_channel - (int)c0 : -1, sp, mt, in + c0, in_fmt, out + c0, out_fmt, cnt, n, rate,
                         nullptr, false, sp ? &pm : nullptr)) {
            gdgb_sync(); if (g_batch) g_batch->slot[si].busy = false;
            return unwind();
        }
        fl.members.push_back(m);
        if (pm) { parts.push_back(pm); strides.push_back((long)n); part_ctx.push_back(cl->ctx[m]); }
    }
    const bool metro = metronome && (master || meter0);
    if ((master && !parts.empty()) || metro) {
        Use u(cl->ctx[0]);
        if (cl->cap_n < n) {
            /* sized once for the longest block the library accepts: the buffers of a ticket still in flight (its master rows
             * are copied out of h_master in gdg_cluster_wait) are never re-made under it when the block length grows */
            const size_t cap = std::max<size_t>(n, GDG_MAX_N);
            gdgb_sync(); cl->release();
            bool ok = true;
            for (int k = 0; k < 2; k++) {
                cl->d_master[k] = (double *)gdgb_alloc(2 * cap * sizeof(double));
                cl->d_master_enc[k] = (double *)gdgb_alloc(2 * cap * sizeof(double));
                cl->d_metro[k] = (double *)gdgb_alloc(cap * sizeof(double));
                cl->h_master[k] = (double *)gdgb_pinned_alloc(2 * cap * sizeof(double));
                ok = ok && cl->d_master[k] && cl->d_master_enc[k] && cl->d_metro[k] && cl->h_master[k];
            }
            if (!ok) { cl->release(); backend_fail("cluster master allocation failed"); return unwind(); }
            cl->cap_n = cap;
        }
        const int ev = si * 28;
        void *s0 = gdgb_stream();
        double *d_metro = nullptr;
        if (metro) {                                        /* controller.go:2724-2738: the metronome's row of this block */
            d_metro = cl->d_metro[si];
            if (gdgb_h2d_async(d_metro, metronome, n * sizeof(double), 1) || gdgb_record(ev + 21, 1) || gdgb_wait(0, ev + 21)) { backend_fail("metronome upload failed"); return unwind(); }
            if (meter0 && meter_rows(meter0, 2 * cnt0, d_metro, n, 1, n, rate, s0)) return unwind();
        }
        if (master && !parts.empty()) {
            double *dm = cl->d_master[si];
            for (int pc : part_ctx) if (gdgb_wait_other(cl->ctx[0], pc, ev + 16)) { backend_fail("cross-device wait failed"); return unwind(); }
            if (gdgb_master_sum(parts.data(), strides.data(), (int)parts.size(), (d_metro && metro_master) ? d_metro : nullptr, dm, dm + n, (int)n, s0))
                { backend_fail("master sum failed"); return unwind(); }
            if (meter0 && meter_rows(meter0, 2 * cnt0 + 1, dm, n, 2, n, rate, s0)) return unwind();      /* controller.go:2768-2770 */
            const void *src = dm;
            if (out_fmt != GDG_FMT_IEEE64) {
                if (gdgb_convert_rows(dm, (long)(n * sizeof(double)), GDG_FMT_IEEE64, cl->d_master_enc[si], (long)ob, out_fmt, 2, (int)n, s0))
                    { backend_fail("payload encode failed"); return unwind(); }
                src = cl->d_master_enc[si];
            }
            if (gdgb_record(ev + 22, 0) || gdgb_wait(2, ev + 22)) { backend_fail("event failed"); return unwind(); }
            if (gdgb_d2h_async(cl->h_master[si], src, 2 * ob, 2) || gdgb_record(ev + 23, 2)) { backend_fail("master download failed"); return unwind(); }
            fl.out_l = out[C]; fl.out_r = out[C + 1];
        } else fl.master = false;
    } else fl.master = false;
    fl.busy = true;
    return si;
}
} // namespace

extern "C" {
gdg_cluster *gdg_cluster_create(const int *devices, int n_devices)
{
    if (ensure_init()) return nullptr;
    if (!devices || n_devices < 1 || n_devices > GDGB_MAX_CONTEXTS) { fail("gdg_cluster_create: bad argument"); return nullptr; }
    std::unique_ptr<gdg_cluster> cl(new gdg_cluster());
    for (int i = 0; i < n_devices; i++) {
        const int c = gdgb_context_open(devices[i]);
        if (c < 0) { backend_fail("gdg_cluster_create"); return nullptr; }
        if (std::find(cl->ctx.begin(), cl->ctx.end(), c) != cl->ctx.end()) { fail("gdg_cluster_create: device %d listed twice", devices[i]); return nullptr; }
        cl->ctx.push_back(c);
    }
    if (n_devices > 1 && gdgb_enable_peers(cl->ctx[0])) { backend_fail("gdg_cluster_create"); return nullptr; }   /* member 0 reads the others' partial masters */
    return cl.release();
}
void gdg_cluster_destroy(gdg_cluster *cl) { delete cl; }
int gdg_cluster_size(gdg_cluster *cl) { return cl ? (int)cl->ctx.size() : 0; }
/* objects created after this call live on member `member`'s device */
int gdg_cluster_select(gdg_cluster *cl, int member)
{
    if (!cl || member < 0 || member >= (int)cl->ctx.size()) return fail("gdg_cluster_select: no member %d", member);
    return gdgb_select(cl->ctx[member]) ? backend_fail("gdg_cluster_select") : 0;
}
int gdg_cluster_submit(gdg_cluster *cl, gdg_chain *const *chains, uint32_t C, gdg_tuner *tuner, int tuner_channel, gdg_spatializer *const *spats,
                       const double *const *in, double *const *out, uint32_t n_out, size_t n, uint32_t rate)
{
    return cluster_block("gdg_cluster_submit", cl, chains, C, tuner, tuner_channel, spats, nullptr, (const void *const *)in, GDG_FMT_IEEE64, nullptr, 0,
                         (void *const *)out, GDG_FMT_IEEE64, n_out, n, rate);
}
int gdg_cluster_submit_aux(gdg_cluster *cl, gdg_chain *const *chains, uint32_t C, gdg_tuner *tuner, int tuner_channel, gdg_spatializer *const *spats,
                           gdg_meter *const *meters, const double *const *in, const double *metronome, int metronome_to_master,
                           double *const *out, uint32_t n_out, size_t n, uint32_t rate)
{
    return cluster_block("gdg_cluster_submit_aux", cl, chains, C, tuner, tuner_channel, spats, meters, (const void *const *)in, GDG_FMT_IEEE64, metronome,
                         metronome_to_master, (void *const *)out, GDG_FMT_IEEE64, n_out, n, rate);
}
int gdg_cluster_submit_encoded(gdg_cluster *cl, gdg_chain *const *chains, uint32_t C, gdg_tuner *tuner, int tuner_channel, gdg_spatializer *const *spats,
                               gdg_meter *const *meters, const void *const *in, int in_format, void *const *out, int out_format,
                               uint32_t n_out, size_t n, uint32_t rate)
{
    return cluster_block("gdg_cluster_submit_encoded", cl, chains, C, tuner, tuner_channel, spats, meters, in, in_format, nullptr, 0, out, out_format,
                         n_out, n, rate);
}
int gdg_cluster_wait(gdg_cluster *cl, int ticket)
{
    if (ensure_init()) return -1;
    if (!cl || ticket < 0 || ticket > 1 || !cl->flight[ticket].busy) return fail("gdg_cluster_wait: ticket %d is not in flight", ticket);
    gdg_cluster::Flight &fl = cl->flight[ticket];
    int rc = 0;
    for (int m : fl.members) {
        Use u(cl->ctx[m]);
        std::lock_guard<std::mutex> lk(g_batch_mutex);
        if (g_batch && g_batch->slot[ticket].busy && block_wait(ticket)) rc = -1;
    }
    if (fl.master) {
        Use u(cl->ctx[0]);
        if (gdgb_sync_event(ticket * 28 + 23)) rc = backend_fail("synchronize failed");
        else { memcpy(fl.out_l, cl->h_master[ticket], fl.ob); memcpy(fl.out_r, (const char *)cl->h_master[ticket] + fl.ob, fl.ob); }
    }
    fl.busy = false;
    return rc;
}
int gdg_cluster_process(gdg_cluster *cl, gdg_chain *const *chains, uint32_t C, gdg_tuner *tuner, int tuner_channel, gdg_spatializer *const *spats,
                        const double *const *in, double *const *out, uint32_t n_out, size_t n, uint32_t rate)
{
    const int t = gdg_cluster_submit(cl, chains, C, tuner, tuner_channel, spats, in, out, n_out, n, rate);
    if (t < 0) return -1;
    return gdg_cluster_wait(cl, t);
}

} /* extern "C" */
