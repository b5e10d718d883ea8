/*
 * emu_backend.cpp - TEST-ONLY host build of the device side of libgdg.
 *
 * Compiles the very same stage source (go-dsp-guitar_b200/csrc/gdg_interp.cuh)
 * with GDG_HOST_EMU, where a "phase" is a loop over the CTA's thread ids, and
 * implements gdg_backend.h with plain host memory.  Linked with the unmodified
 * gdg_host.cpp it yields tests/emu/libgdg_emu.so, which lets the CPU test suite
 * (-m "not gpu") check the host logic AND the parallel reformulation of every
 * unit (scans, rounds, tiles, FFT) against the oracle without a GPU.
 *
 * It is never shipped, never loaded by the product package, and is not a
 * fallback: go-dsp-guitar_b200/libgdg.so contains only the CUDA backend and
 * refuses to start without a B200-class device.
 */
#define GDG_HOST_EMU 1
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../go-dsp-guitar_b200/csrc/gdg_backend.h"
#include "../../go-dsp-guitar_b200/csrc/gdg_interp.cuh"
#include "../../go-dsp-guitar_b200/csrc/gdg_os.cuh"

static thread_local std::string t_err;
static std::vector<double> g_tw;
static std::atomic<uint64_t> g_launches{0};

const char *gdgb_error(void) { return t_err.c_str(); }
const char *gdgb_name(void) { return "host-emulation (tests only)"; }
/* device contexts: the emulation has one memory space; contexts are bookkeeping only (tests of the multi-device host logic) */
static int g_nctx = 1, g_ctx_dev[GDGB_MAX_CONTEXTS] = {0};
static thread_local int t_cur = 0;
int gdgb_init(int) { return 0; }
int gdgb_device_count(void) { const char *e = getenv("GDG_EMU_DEVICES"); return e ? atoi(e) : 1; }
int gdgb_context_open(int device)
{
    if (device < 0 || device >= gdgb_device_count()) { t_err = "device index out of range"; return -1; }
    for (int i = 0; i < g_nctx; i++) if (g_ctx_dev[i] == device) return i;
    if (g_nctx >= GDGB_MAX_CONTEXTS) { t_err = "too many device contexts"; return -1; }
    g_ctx_dev[g_nctx] = device;
    return g_nctx++;
}
int gdgb_context_count(void) { return g_nctx; }
int gdgb_context_device(int ctx) { return (ctx >= 0 && ctx < g_nctx) ? g_ctx_dev[ctx] : -1; }
int gdgb_current(void) { return t_cur; }
int gdgb_select(int ctx) { if (ctx < 0 || ctx >= g_nctx) { t_err = "no such device context"; return -1; } t_cur = ctx; return 0; }
int gdgb_bind(void) { return 0; }
int gdgb_enable_peers(int) { return 0; }
int gdgb_wait_other(int, int, int) { return 0; }
int gdgb_master_sum(const double *const *parts, const long *strides, int count, const double *aux, double *outL, double *outR, int n, void *)
{
    for (int j = 0; j < n; j++) {
        double l = 0.0, r = 0.0;
        for (int d = 0; d < count; d++) { l += parts[d][j]; r += parts[d][strides[d] + j]; }
        if (aux) { l += aux[j]; r += aux[j]; }
        outL[j] = l; outR[j] = r;
    }
    g_launches++;
    return 0;
}
int gdgb_device(void) { return -1; }
int gdgb_sm_count(void) { return 148; }            /* the part being emulated: pipeline groups are sized like on a B200 */
void *gdgb_stream(void) { return nullptr; }
uint64_t gdgb_launch_count(void) { return g_launches.load(); }
void *gdgb_alloc(size_t bytes) { return calloc(1, bytes ? bytes : 8); }
void gdgb_free(void *p) { free(p); }
int gdgb_zero(void *p, size_t bytes) { memset(p, 0, bytes); return 0; }
int gdgb_upload(void *d, const void *s, size_t b) { memcpy(d, s, b); return 0; }
int gdgb_download(void *d, const void *s, size_t b) { memcpy(d, s, b); return 0; }
int gdgb_copy(void *d, const void *s, size_t b) { memmove(d, s, b); return 0; }
int gdgb_copy_on(void *d, const void *s, size_t b, void *) { memmove(d, s, b); return 0; }
int gdgb_sync(void) { return 0; }
void *gdgb_stream_handle(int) { return nullptr; }
void *gdgb_side_stream(int) { return nullptr; }
int gdgb_h2d_async(void *d, const void *s, size_t b, int) { memcpy(d, s, b); return 0; }
int gdgb_d2h_async(void *d, const void *s, size_t b, int) { memcpy(d, s, b); return 0; }
int gdgb_record(int, int) { return 0; }
int gdgb_order(void *, void *, int) { return 0; }
int gdgb_wait(int, int) { return 0; }
int gdgb_sync_stream(int) { return 0; }
int gdgb_sync_event(int) { return 0; }
void *gdgb_pinned_alloc(size_t bytes) { return malloc(bytes ? bytes : 8); }
void gdgb_pinned_free(void *p) { free(p); }
int gdgb_is_pinned(const void *) { return 0; }

int gdgb_set_tables(const double *aa2, const double *aa4, const double *lz2, const double *lz4, const double *tw)
{
    memcpy(c_aa2, aa2, sizeof(double) * 80);
    memcpy(c_aa4, aa4, sizeof(double) * 160);
    memcpy(c_lz2, lz2, sizeof(double) * 6);
    memcpy(c_lz4, lz4, sizeof(double) * 18);
    g_tw.assign(tw, tw + GDG_MAX_FFT);
    return 0;
}

int gdgb_launch_segments(int kind, const gdg_chain_desc *const *descs, const int32_t *nseg, int round, int channels,
                         const void *in, long in_stride, int in_fmt, void *out, long out_stride, int out_fmt,
                         void *mid, long mid_stride, double *side, long side_stride, int n, int, void *)
{
    std::vector<double> smem(kind == 0 ? GDG_SMEM_DOUBLES : GDG_OSK_SMEM_DOUBLES);
    for (int c = 0; c < channels; c++) {
        if (!descs[c]) continue;
        const bool first = round == 0, last = round == nseg[c] - 1;
        const unsigned char *src = first ? (const unsigned char *)in + (long)c * in_stride : (const unsigned char *)mid + (long)c * mid_stride;
        unsigned char *dst = last ? (unsigned char *)out + (long)c * out_stride : (unsigned char *)mid + (long)c * mid_stride;
        const int sf = first ? in_fmt : GDG_FMT_IEEE64, df = last ? out_fmt : GDG_FMT_IEEE64;
        double *sd = (last && side && out_fmt != GDG_FMT_IEEE64) ? side + (long)c * side_stride : nullptr;
        if (kind == 0) run_chain(smem.data(), descs[c], src, sf, dst, df, sd, n, g_tw.data());
        else run_os(smem.data(), descs[c], src, sf, dst, df, sd, n);
    }
    g_launches++;
    return 0;
}
int gdgb_launch_meter(const double *in, long stride, int channels, long n, double f, long long hold, double *state, void *)
{
    std::vector<double> red(GDG_METER_RED);
    for (int c = 0; c < channels; c++) meter_channel(red.data(), in + (long)c * stride, n, f, hold, state + 4 * (long)c);
    if (channels > 0 && n > 0) g_launches++;
    return 0;
}
int gdgb_convert_rows(const void *in, long in_stride, int in_fmt, void *out, long out_stride, int out_fmt, int rows, int n, void *)
{
    for (int r = 0; r < rows; r++)
        for (int k = 0; k < n; k++)
            encode_sample(out_fmt, (unsigned char *)out + (long)r * out_stride, k, decode_sample(in_fmt, (const unsigned char *)in + (long)r * in_stride, k));
    if (rows > 0 && n > 0) g_launches++;
    return 0;
}
int gdgb_resample_time(const double *in, long n, double dx, double *out, long out_len, void *)
{
    for (long i = 0; i < out_len; i++) out[i] = resample_time_body(in, n, dx, i);
    g_launches++;
    return 0;
}

int gdgb_launch_spatializer(const gdg_spat_channel *ch, int C, const double *in, long stride, const double *hist_in,
                            double *hist_out, int hs, const double *aux, double *outL, double *outR, int n, int accumulate, double *part, void *)
{
    if (C <= GDG_SPAT_GROUP || !part) {
        for (int j = 0; j < n; j++) spat_frame(ch, C, in, stride, hist_in, hs, aux, outL, outR, j, accumulate);
    } else {
        const int groups = (C + GDG_SPAT_GROUP - 1) / GDG_SPAT_GROUP;
        for (int g = 0; g < groups; g++) for (int j = 0; j < n; j++) spat_partial(ch, C, in, stride, hist_in, hs, part, n, g, j);
        for (int j = 0; j < n; j++) spat_combine(part, groups, aux, outL, outR, n, j, accumulate);
    }
    for (int c = 0; c < C; c++)
        for (int i = 0; i < hs; i++) {
            const long src = (long)i + n - hs;
            hist_out[(long)c * hs + i] = src >= 0 ? in[(long)c * stride + src] : hist_in[(long)c * hs + i + n];
        }
    g_launches += 2;
    return 0;
}

int gdgb_autocorrelate(const double *ring, int n_ring, int pos, double *w, int F, void *)
{
    const int M = F / 2;
    int logM = 0;
    while ((1 << logM) < M) logM++;
    for (int i = 0; i < F; i++) ac_load_body(ring, n_ring, pos, w, i);
    for (int h = M / 2; h >= 1; h >>= 1) for (int b = 0; b < M / 2; b++) ac_dif_body(w, h, b);
    for (int k = 0; k <= M / 2; k++) ac_power_body(w, M, logM, k);
    for (int h = 1; h <= M / 2; h <<= 1) for (int b = 0; b < M / 2; b++) ac_dit_body(w, h, b);
    for (int i = 0; i < F; i++) w[i] *= 1.0 / (double)M;
    return 0;
}

/* ---- impulse-response compilation (host loops over the same bodies) ---- */
static int ilog2(int v) { int e = 0; while ((1 << e) < v) e++; return e; }
int gdgb_ir_real_spectrum(const double *x, long n_x, int nfft, double *w, double *spec, void *)
{
    const int M = nfft / 2, logM = ilog2(M);
    for (long i = 0; i < nfft; i++) ir_pad_body(x, n_x, w, i);
    for (int h = M / 2; h >= 1; h >>= 1) for (int b = 0; b < M / 2; b++) ac_dif_body(w, h, b);
    for (int k = 0; k <= M; k++) ir_split_body(w, M, logM, spec, k);
    g_launches++;
    return 0;
}
int gdgb_ir_resample_bins(const double *in, int bins_in, double *out, int bins_out, void *)
{
    for (int i = 0; i < bins_out; i++) ir_resample_bins_body(in, bins_in, out, bins_out, i);
    g_launches++;
    return 0;
}
int gdgb_ir_real_inverse(const double *spec, int nfft, double *w, double *out, long n_keep, void *)
{
    const int M = nfft / 2, logM = ilog2(M);
    for (int k = 0; k < M; k++) ir_unsplit_body(spec, M, logM, w, k);
    for (int h = 1; h <= M / 2; h <<= 1) for (int b = 0; b < M / 2; b++) ac_dit_body(w, h, b);
    for (long i = 0; i < n_keep; i++) out[i] = (1.0 / (double)M) * w[i];
    g_launches++;
    return 0;
}
int gdgb_ir_sum_squares(const double *x, long n, double *ss, void *)
{
    double red[GDG_NT];
    for (int t = 0; t < GDG_NT; t++) { double s = 0.0; for (long i = t; i < n; i += GDG_NT) s += x[i] * x[i]; red[t] = s; }
    for (int h = GDG_NT / 2; h >= 1; h >>= 1) for (int t = 0; t < h; t++) red[t] += red[t + h];
    ss[0] = red[0];
    g_launches++;
    return 0;
}
int gdgb_ir_scale_add(double *acc, const double *x, long n, double comp, double fac, const double *ss, void *)
{
    for (long i = 0; i < n; i++) ir_scale_add_body(acc, x, comp, fac, ss, i);
    g_launches++;
    return 0;
}
int gdgb_ir_partition_spectra(const double *coeffs, int L, int B, int logB, int K, double *H, void *)
{
    std::vector<double> smem(GDG_SMEM_DOUBLES);
    for (int p = 0; p < K; p++) run_partition_spectrum(smem.data(), coeffs, L, B, logB, p, H + 2L * (long)p * (B + 1), g_tw.data());
    g_launches++;
    return 0;
}
int gdgb_autocorrelate_batch(const double *const *rings, int n_ring, const int *pos, double *const *works, int count, int F, void *st)
{
    for (int c = 0; c < count; c++) if (gdgb_autocorrelate(rings[c], n_ring, pos[c], works[c], F, st)) return -1;
    return 0;
}
