/*
 * gdg_interp.cuh - the fused chain-segment interpreter: one CTA walks the
 * non-bypassed slots of its channel's chain (signal/signal.go:361-414) with the
 * block resident in shared memory, plus the spatializer mix-down and the tuner
 * ring feed.  Compiles for sm_100a (gdg_kernels.cu) and, for tests only, as host
 * C++ (GDG_HOST_EMU).
 */
#ifndef GDG_INTERP_CUH
#define GDG_INTERP_CUH

#include "gdg_stages2.cuh"

/* --------------------------------------------------------------------------
 * Sample codecs of wave/wave.go (bytesToSamples / samplesToBytes, wave.go:275-732) as the load / store
 * prologue and epilogue of the chain kernel: formats GDG_FMT_* of gdg_device.h, little-endian payload.
 * Every step is an exactly rounded IEEE operation, so the bytes equal the reference's for equal float64 input.
 * -------------------------------------------------------------------------- */
GDG_DEVI int fmt_bytes(int fmt) { return fmt == GDG_FMT_PCM8 ? 1 : (fmt == GDG_FMT_PCM16 ? 2 : (fmt == GDG_FMT_PCM24 ? 3 : (fmt == GDG_FMT_IEEE64 ? 8 : 4))); }
GDG_DEVI double decode_sample(int fmt, const unsigned char *p, int k)
{
    switch (fmt) {
    case GDG_FMT_PCM8: {                                     /* wave.go:316-342 */
        const double r = (1.0 / 127.0) * (double)((int)p[k] - 128);
        return r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r); }
    case GDG_FMT_PCM16: {                                    /* wave.go:400-428 */
        const int v = (int)(short)((unsigned)p[2 * k] | ((unsigned)p[2 * k + 1] << 8));
        return (2.0 / 65535.0) * (double)v; }
    case GDG_FMT_PCM24: {                                    /* wave.go:483-531 */
        const unsigned w = (unsigned)p[3 * k] | ((unsigned)p[3 * k + 1] << 8) | ((unsigned)p[3 * k + 2] << 16);
        int v = (int)w;
        if (w & 0x00800000u) v = -8388608 + (v & 0x007fffff);
        return (2.0 / 16777215.0) * (double)v; }
    case GDG_FMT_PCM32: {                                    /* wave.go:589-617 */
        const unsigned w = (unsigned)p[4 * k] | ((unsigned)p[4 * k + 1] << 8) | ((unsigned)p[4 * k + 2] << 16) | ((unsigned)p[4 * k + 3] << 24);
        return (2.0 / 4294967295.0) * (double)(int)w; }
    case GDG_FMT_IEEE32: return (double)((const float *)p)[k];           /* wave.go:662-689 */
    default: return ((const double *)p)[k];                              /* wave.go:714-732 */
    }
}
GDG_DEVI void encode_sample(int fmt, unsigned char *p, int k, double s)
{
    if (fmt == GDG_FMT_IEEE64) { ((double *)p)[k] = s; return; }         /* wave.go:694-709: no clamp */
    s = s < -1.0 ? -1.0 : (s > 1.0 ? 1.0 : s);
    switch (fmt) {
    case GDG_FMT_PCM8: {                                     /* wave.go:275-311 */
        const int res = (int)(short)(127.0 * s) + 128;
        p[k] = (unsigned char)(res < 0 ? 0 : (res > 255 ? 255 : res)); break; }
    case GDG_FMT_PCM16: {                                    /* wave.go:347-395 */
        int t = (int)((0.5 * 65535.0) * s);
        t = t > 32767 ? 32767 : (t < -32768 ? -32768 : t);
        p[2 * k] = (unsigned char)(t & 0xff); p[2 * k + 1] = (unsigned char)((t >> 8) & 0xff); break; }
    case GDG_FMT_PCM24: {                                    /* wave.go:433-478 */
        int t = (int)((0.5 * 16777215.0) * s);
        t = t > 8388607 ? 8388607 : (t < -8388608 ? -8388608 : t);
        p[3 * k] = (unsigned char)(t & 0xff); p[3 * k + 1] = (unsigned char)((t >> 8) & 0xff); p[3 * k + 2] = (unsigned char)((t >> 16) & 0xff); break; }
    case GDG_FMT_PCM32: {                                    /* wave.go:536-584 */
        long long t = (long long)((0.5 * 4294967295.0) * s);
        t = t > 2147483647LL ? 2147483647LL : (t < -2147483648LL ? -2147483648LL : t);
        const unsigned u = (unsigned)(int)t;
        p[4 * k] = (unsigned char)(u & 0xff); p[4 * k + 1] = (unsigned char)((u >> 8) & 0xff);
        p[4 * k + 2] = (unsigned char)((u >> 16) & 0xff); p[4 * k + 3] = (unsigned char)(u >> 24); break; }
    default: ((float *)p)[k] = (float)s; break;              /* IEEE32, wave.go:622-657 */
    }
}

/* resample.Time (resample/resample.go:72-103): out[i] = Lanczos-3 interpolation of the input at i * src/dst */
GDG_DEVI double resample_time_body(const double *in, long n, double dx, long i)
{
    const double x = (double)i * dx;
    const long idx = (long)floor(x);
    double sum = 0.0;
    for (long t = idx - 2; t < idx + 4; t++)
        if (t >= 0 && t < n) {
            const double df = x - (double)t;
            double kv;
            if (df == 0.0) kv = 1.0;
            else if (-3.0 < df && df < 3.0) { const double pix = GDG_PI * df; kv = (3.0 * (sin(pix) * sin(pix / 3.0))) / (pix * pix); }
            else kv = 0.0;
            sum += in[t] * kv;
        }
    return sum;
}

/* --------------------------------------------------------------------------
 * level meter (level/level.go:147-208): per channel and sample
 *     cur *= f;  if counter > hold: peak *= f else counter++;
 *     if |s| > cur: cur = |s|;  if |s| >= peak: peak = |s|, counter = 0
 * with f = 10^(-1/(1.7 fs)) and hold = 2 fs samples.  One CTA per channel walks the buffer in
 * segments of GDG_METER_SEG <= hold samples (a thread owns 16 consecutive samples):
 *   cur  : max-times reduction  cur' = max(cur f^len, max_t e_t f^(len - end_t)),  e_t the thread's exact local run;
 *   peak : a segment shorter than the hold time can decay the peak only BEFORE its first reset, so
 *          j* = first j with |s_j| >= peak f^(decays up to j)   (min-reduction),
 *          after j* the peak is the running maximum: peak' = max_{j >= j*} |s_j|, counter' = len-1 - (last index of it);
 *          without a reset peak' = peak f^(decays), counter' advances to at most hold + 1.
 * Powers f^k are formed by squaring instead of k successive products: the state agrees with the reference to a
 * few 1e-15 relative (the dB read-out, level.go:95-137, is integer).  Segments longer than the hold time
 * (sample rates below GDG_METER_SEG/2 Hz) are walked by one thread with the literal loop.
 * state words: st[0] cur, st[1] peak, st[2] counter (int64).  red: 4*GDG_NT + 8 doubles of shared scratch.
 * -------------------------------------------------------------------------- */
#define GDG_METER_SPT 16
#define GDG_METER_SEG (GDG_NT * GDG_METER_SPT)
#define GDG_METER_RED (4 * GDG_NT + 8)
GDG_DEV void meter_segment(double *red, const double *x, int len, double f, long long hold, double *st)
{
    double *r_cur = red, *r_m = red + GDG_NT, *r_j = red + 2 * GDG_NT, *r_l = red + 3 * GDG_NT, *res = red + 4 * GDG_NT;
    GDG_REGA(double, v, GDG_METER_SPT);
    if ((long long)len > hold) {                                   /* not reachable for fs >= 4096 Hz */
        GDG_THREADS_BEGIN
            if (tid == 0) {
                double cur = st[0], peak = st[1]; long long cnt = st_geti(st, 2);
                for (int j = 0; j < len; j++) {
                    cur *= f;
                    if (cnt > hold) peak *= f; else cnt++;
                    const double a = fabs(x[j]);
                    if (a > cur) cur = a;
                    if (a >= peak) { peak = a; cnt = 0; }
                }
                st[0] = cur; st[1] = peak; st_seti(st, 2, cnt);
            }
        GDG_THREADS_END
        return;
    }
    const double c0 = st[0], p0 = st[1];
    const long long q0 = st_geti(st, 2);
    const int grace = q0 <= hold ? (int)(hold + 1 - q0 < (long long)len ? hold + 1 - q0 : (long long)len) : 0;   /* samples that only count */
    GDG_THREADS_BEGIN
        const int k0 = tid * GDG_METER_SPT, cnt = imax(0, imin(GDG_METER_SPT, len - k0));
        double e = 0.0, thr = p0 * pow_int(f, imax(0, k0 - grace));
        int jf = len;                                               /* first reset this chunk would cause, if none came earlier */
#pragma unroll
        for (int r = 0; r < GDG_METER_SPT; r++) {
            const int j = k0 + r;
            const double a = r < cnt ? fabs(x[j]) : 0.0;
            R(v)[r] = a;
            if (r < cnt) {
                e *= f; if (a > e) e = a;
                if (j >= grace) thr *= f;
                if (jf == len && a >= thr) jf = j;
            }
        }
        r_cur[tid] = cnt > 0 ? e * pow_int(f, len - (k0 + cnt)) : 0.0;
        r_j[tid] = (double)jf;
    GDG_THREADS_END
    GDG_THREADS_BEGIN
        if (tid < 32) {
            double m = 0.0, j = (double)len;
            for (int q = tid * (GDG_NT / 32); q < (tid + 1) * (GDG_NT / 32); q++) { if (r_cur[q] > m) m = r_cur[q]; if (r_j[q] < j) j = r_j[q]; }
            r_m[tid] = m; r_l[tid] = j;
        }
    GDG_THREADS_END
    GDG_THREADS_BEGIN
        if (tid == 0) {
            double m = 0.0, j = (double)len;
            for (int q = 0; q < 32; q++) { if (r_m[q] > m) m = r_m[q]; if (r_l[q] < j) j = r_l[q]; }
            res[1] = m; res[2] = j;
        }
    GDG_THREADS_END
    const int jstar = (int)res[2];
    const double curmax = res[1];
    GDG_THREADS_BEGIN
        const int k0 = tid * GDG_METER_SPT, cnt = imax(0, imin(GDG_METER_SPT, len - k0));
        double m = -1.0; int li = -1;
#pragma unroll
        for (int r = 0; r < GDG_METER_SPT; r++) {
            const int j = k0 + r;
            if (r < cnt && j >= jstar && R(v)[r] >= m) { m = R(v)[r]; li = j; }
        }
        r_cur[tid] = m; r_j[tid] = (double)li;
    GDG_THREADS_END
    GDG_THREADS_BEGIN
        if (tid < 32) {
            double m = -1.0, l = -1.0;
            for (int q = tid * (GDG_NT / 32); q < (tid + 1) * (GDG_NT / 32); q++) if (r_cur[q] > m || (r_cur[q] == m && r_j[q] > l)) { m = r_cur[q]; l = r_j[q]; }
            r_m[tid] = m; r_l[tid] = l;
        }
    GDG_THREADS_END
    GDG_THREADS_BEGIN
        if (tid == 0) {
            double m = -1.0, l = -1.0;
            for (int q = 0; q < 32; q++) if (r_m[q] > m || (r_m[q] == m && r_l[q] > l)) { m = r_m[q]; l = r_l[q]; }
            const double cdec = c0 * pow_int(f, len);
            st[0] = curmax > cdec ? curmax : cdec;
            if (jstar < len) { st[1] = m; st_seti(st, 2, (long long)(len - 1) - (long long)l); }
            else {
                st[1] = p0 * pow_int(f, len - grace);
                st_seti(st, 2, q0 <= hold ? q0 + (long long)grace : q0);
            }
        }
    GDG_THREADS_END
}

GDG_DEV void meter_channel(double *red, const double *x, long n, double f, long long hold, double *st)
{
    for (long o = 0; o < n; o += GDG_METER_SEG) {
        const long len = n - o < (long)GDG_METER_SEG ? n - o : (long)GDG_METER_SEG;
        meter_segment(red, x + o, (int)len, f, hold, st);
    }
}

GDG_DEVI void ctx_init(gdg_ctx &cx, int n)
{
    cx.cur_off = GDG_SMEM_R; cx.alt_off = GDG_SMEM_R + GDG_MAX_N; cx.n = n;
}

GDG_DEV void run_unit(gdg_ctx &cx, const gdg_unit_desc &d, const double *tw)
{
    switch (d.type) {
    case GDG_UNIT_SIGNALGENERATOR: stage_siggen(cx, d); break;
    case GDG_UNIT_NOISEGATE: stage_noisegate(cx, d); break;
    case GDG_UNIT_BANDPASS: stage_bandpass(cx, d); break;
    case GDG_UNIT_AUTOWAH: stage_autowah(cx, d); break;
    case GDG_UNIT_AUTOYOY: stage_autoyoy(cx, d); break;
    case GDG_UNIT_COMPRESSOR: stage_compressor(cx, d); break;
    case GDG_UNIT_OCTAVER: stage_octaver(cx, d); break;
    case GDG_UNIT_EXCESS: case GDG_UNIT_FUZZ: case GDG_UNIT_OVERDRIVE: case GDG_UNIT_DISTORTION:
        stage_oversampled(cx, d); break;
    case GDG_UNIT_TONESTACK: stage_tonestack(cx, d); break;
    case GDG_UNIT_CHORUS: stage_chorus(cx, d); break;
    case GDG_UNIT_FLANGER: case GDG_UNIT_PHASER: stage_flanger(cx, d); break;
    case GDG_UNIT_TREMOLO: stage_tremolo(cx, d); break;
    case GDG_UNIT_RINGMODULATOR: stage_ringmod(cx, d); break;
    case GDG_UNIT_DELAY: stage_delay(cx, d); break;
    case GDG_UNIT_REVERB: stage_reverb(cx, d); break;
    case GDG_UNIT_POWERAMP: stage_poweramp(cx, d, tw); break;
    case GDG_UNIT_CABINET: stage_cascade(cx, d, 7, 3); break;
    default: break;
    }
}

/* one channel, one block: load -> every slot -> store */
GDG_DEV void run_chain(double *smem, const gdg_chain_desc *cd, const void *in, int in_fmt, void *out, int out_fmt, double *side, int n, const double *tw)
{
#if !defined(__CUDACC__) || defined(GDG_HOST_EMU)
    gdg_smem = smem;
#else
    (void)smem;
#endif
    /* input rows (float64 or an encoded payload) come in as ONE bulk asynchronous copy (TMA, cp.async.bulk -> UBLKCP) into S, signalled through an
     * mbarrier, while the threads fetch the twiddles and the chain descriptor; then S is folded into the swizzled block
     * layout.  Needs 16-byte aligned rows of a multiple of 16 bytes; anything else takes the per-thread loop. */
    const unsigned in_bytes = (unsigned)(n * fmt_bytes(in_fmt));
    const bool bulk_in = GDG_HAVE_BULK && (in_bytes % 16) == 0 && (((unsigned long long)in) & 15ull) == 0;
    double *mbar = SM_MISC + 32;
    gdg_ctx &cx = *(gdg_ctx *)(gdg_smem + GDG_SMEM_CTX);
    GDG_THREADS_BEGIN
        if (tid == 0) { ctx_init(cx, n); if (bulk_in) bulk_load(mbar, SM_S, in, in_bytes); }
    GDG_THREADS_END
    GDG_THREADS_BEGIN
        if (!bulk_in) {
            if (in_fmt == GDG_FMT_IEEE64) { for (int k = tid; k < n; k += GDG_NT) SM_CUR(cx)[sw(k)] = ((const double *)in)[k]; }
            else { for (int k = tid; k < n; k += GDG_NT) SM_CUR(cx)[sw(k)] = decode_sample(in_fmt, (const unsigned char *)in, k); }
        }
        for (int k = tid; k < GDG_MAX_N / 8; k += GDG_NT) { SM_TWS[2 * k] = GDG_LDG(tw + 4 * k); SM_TWS[2 * k + 1] = GDG_LDG(tw + 4 * k + 1); }   /* exp(-2 pi i k / 8192) */
        /* the chain descriptor rides along with the input load: no global round trip between units */
        for (int k = tid; k < (int)(sizeof(gdg_chain_desc) / 4); k += GDG_NT) ((int *)(gdg_smem + GDG_SMEM_CHAIN))[k] = ((const int *)cd)[k];
        if (bulk_in) {
            bulk_wait(mbar);
            if (in_fmt == GDG_FMT_IEEE64) {
#pragma unroll 4
                for (int k = tid; k < n; k += GDG_NT) SM_CUR(cx)[sw(k)] = SM_S[k];
            } else {                                               /* encoded payload: decode out of shared memory */
#pragma unroll 4
                for (int k = tid; k < n; k += GDG_NT) SM_CUR(cx)[sw(k)] = decode_sample(in_fmt, (const unsigned char *)SM_S, k);
            }
        }
    GDG_THREADS_END
    const gdg_chain_desc *scd = (const gdg_chain_desc *)(gdg_smem + GDG_SMEM_CHAIN);
    const int ns = scd->nslots;
    for (int s = 0; s < ns; s++) {
        GDG_THREADS_BEGIN
            if (tid < (int)(sizeof(gdg_unit_desc) / 4)) ((int *)(gdg_smem + GDG_SMEM_DESC))[tid] = ((const int *)&scd->slot[s])[tid];
        GDG_THREADS_END
        run_unit(cx, GDG_DESC, tw);
    }
    GDG_THREADS_BEGIN
        if (out_fmt == GDG_FMT_IEEE64) { for (int k = tid; k < n; k += GDG_NT) ((double *)out)[k] = SM_CUR(cx)[sw(k)]; }
        else if (side == nullptr) { for (int k = tid; k < n; k += GDG_NT) encode_sample(out_fmt, (unsigned char *)out, k, SM_CUR(cx)[sw(k)]); }
        else { for (int k = tid; k < n; k += GDG_NT) { const double v = SM_CUR(cx)[sw(k)]; side[k] = v; encode_sample(out_fmt, (unsigned char *)out, k, v); } }
    GDG_THREADS_END
}

/* --------------------------------------------------------------------------
 * spatializer mix-down (spatializer/spatializer.go:140-335): each thread owns
 * frames; channels are accumulated in ascending order exactly like the
 * reference (0..C-1, then aux).  hist: [C][hs] linear history (oldest first).
 * Launched over frames; a second pass (same launch, after the loop) is not
 * possible because other CTAs still read the history -> the history update is
 * its own kernel (spat_push).
 * -------------------------------------------------------------------------- */
#define GDG_SPAT_GROUP 32          /* channels mixed by one thread before the partial sums are combined */
/* channels [c0, c1) of frame j accumulated in ascending order onto (l, r) */
GDG_DEVI void spat_accumulate(const gdg_spat_channel *ch, int c0, int c1, const double *in, long stride, const double *hist, int hs,
                              int j, double &l, double &r)
{
    for (int c = c0; c < c1; c++) {
        const gdg_spat_channel &p = ch[c];
        const double *x = in + (long)c * stride;
        const double cur = x[j];
        if (p.delay_time == 0.0) {
            l += p.fac_left * cur;
            r += p.fac_right * cur;
        } else {
            const int ie = j - p.d_early, il = j - p.d_late;
            const double *h = hist + (long)c * hs;
            const double se = ie >= 0 ? x[ie] : h[hs + ie];
            const double sl = il >= 0 ? x[il] : h[hs + il];
            const double dsmp = (p.w_early * se) + (p.w_late * sl);
            if (p.delay_time > 0.0) { l += p.fac_left * dsmp; r += p.fac_right * cur; }
            else { l += p.fac_left * cur; r += p.fac_right * dsmp; }
        }
    }
}
/* single pass (C <= GDG_SPAT_GROUP): exactly the reference's accumulation order (channels ascending, then aux) */
GDG_DEVI void spat_frame(const gdg_spat_channel *ch, int C, const double *in, long stride, const double *hist, int hs,
                         const double *aux, double *outL, double *outR, int j, int accumulate)
{
    double l = accumulate ? outL[j] : 0.0, r = accumulate ? outR[j] : 0.0;
    spat_accumulate(ch, 0, C, in, stride, hist, hs, j, l, r);
    if (aux) { l += aux[j]; r += aux[j]; }
    outL[j] = l; outR[j] = r;
}
/* two passes for many channels: group g -> part[(2 g + {0,1}) * n + j]; then the groups are summed in ascending order */
GDG_DEVI void spat_partial(const gdg_spat_channel *ch, int C, const double *in, long stride, const double *hist, int hs,
                           double *part, int n, int g, int j)
{
    double l = 0.0, r = 0.0;
    const int c0 = g * GDG_SPAT_GROUP, c1 = imin(C, c0 + GDG_SPAT_GROUP);
    spat_accumulate(ch, c0, c1, in, stride, hist, hs, j, l, r);
    part[(long)(2 * g) * n + j] = l; part[(long)(2 * g + 1) * n + j] = r;
}
GDG_DEVI void spat_combine(const double *part, int groups, const double *aux, double *outL, double *outR, int n, int j, int accumulate)
{
    double l = accumulate ? outL[j] : 0.0, r = accumulate ? outR[j] : 0.0;
    for (int g = 0; g < groups; g++) { l += part[(long)(2 * g) * n + j]; r += part[(long)(2 * g + 1) * n + j]; }
    if (aux) { l += aux[j]; r += aux[j]; }
    outL[j] = l; outR[j] = r;
}

/* --------------------------------------------------------------------------
 * tuner autocorrelation (tuner/tuner.go:379-446): real FFT of size F through a
 * complex FFT of size M = F/2 living in global memory; per-index bodies shared by
 * the CUDA kernels and the test-only host build.
 * -------------------------------------------------------------------------- */
#if defined(__CUDACC__) && !defined(GDG_HOST_EMU)
#define GDG_SINCOSPI(x, s, c) sincospi((x), (s), (c))
#else
#define GDG_SINCOSPI(x, s, c) do { *(s) = sin(GDG_PI * (x)); *(c) = cos(GDG_PI * (x)); } while (0)
#endif
GDG_DEVI void ac_load_body(const double *ring, int n_ring, int pos, double *w, int i)
{
    double v = 0.0;
    if (i < n_ring) { int p = pos + i; if (p >= n_ring) p -= n_ring; v = ring[p]; }     /* circular.Retrieve: oldest first */
    w[i] = v;
}
GDG_DEVI void ac_twiddle(int pos, int h, double &re, double &im)
{
    double s, c;
    GDG_SINCOSPI(-(double)pos / (double)h, &s, &c);            /* exp(-2 pi i pos / (2h)) */
    re = c; im = s;
}
/* one radix-2 butterfly on the cells i0, i1 = i0 + (row distance) of w; pos / h name its twiddle exp(-2 pi i pos / (2h)) */
GDG_DEVI void ac_dif_pair(double *w, int i0, int i1, int pos, int h)
{
    const double ar = w[2 * i0], ai = w[2 * i0 + 1], cr = w[2 * i1], ci = w[2 * i1 + 1];
    double tr, ti; ac_twiddle(pos, h, tr, ti);
    const double dr = ar - cr, di = ai - ci;
    w[2 * i0] = ar + cr; w[2 * i0 + 1] = ai + ci;
    w[2 * i1] = dr * tr - di * ti; w[2 * i1 + 1] = dr * ti + di * tr;
}
GDG_DEVI void ac_dit_pair(double *w, int i0, int i1, int pos, int h)
{
    const double ar = w[2 * i0], ai = w[2 * i0 + 1], cr = w[2 * i1], ci = w[2 * i1 + 1];
    double tr, ti; ac_twiddle(pos, h, tr, ti); ti = -ti;
    const double xr = cr * tr - ci * ti, xi = cr * ti + ci * tr;
    w[2 * i0] = ar + xr; w[2 * i0 + 1] = ai + xi;
    w[2 * i1] = ar - xr; w[2 * i1 + 1] = ai - xi;
}
GDG_DEVI void ac_dif_body(double *w, int h, int b)
{
    const int grp = b / h, pos = b - grp * h, i0 = grp * 2 * h + pos;
    ac_dif_pair(w, i0, i0 + h, pos, h);
}
GDG_DEVI void ac_dit_body(double *w, int h, int b)
{
    const int grp = b / h, pos = b - grp * h, i0 = grp * 2 * h + pos;
    ac_dit_pair(w, i0, i0 + h, pos, h);
}
/* split -> |X|^2 (tuner.go:437-440) -> inverse split, on bit-reversed storage */
GDG_DEVI void ac_power_body(double *w, int M, int logM, int k)
{
    if (k == 0) {
        const double zr = w[0], zi = w[1];
        const double x0 = zr + zi, xm = zr - zi;
        const double y0 = x0 * x0, ym = xm * xm;
        w[0] = 0.5 * (y0 + ym); w[1] = 0.5 * (y0 - ym);
        return;
    }
    const int k2 = M - k;
    const int p1 = brev(k, logM), p2 = brev(k2, logM);
    const double ar = w[2 * p1], ai = w[2 * p1 + 1], br = w[2 * p2], bi = w[2 * p2 + 1];
    double s, c;
    GDG_SINCOSPI(-(double)k / (double)M, &s, &c);               /* W_F^k, F = 2M */
    /* X[k] = 0.5((A + conj B) - jW (A - conj B)) */
    const double sr = ar + br, si = ai - bi, dr = ar - br, di = ai + bi;
    const double jr = -s, ji = c;                           /* jW */
    const double tr = jr * dr - ji * di, ti = jr * di + ji * dr;
    const double xr = 0.5 * (sr - tr), xi = 0.5 * (si - ti);
    const double pa = xr * xr + xi * xi;                    /* |X[k]|^2 ( = |X[M-k]|^2 by symmetry of a real input) */
    /* X[M-k] */
    const double s2r = br + ar, s2i = bi - ai, d2r = br - ar, d2i = bi + ai;
    const double j2r = jr, j2i = -ji;                       /* conj(jW) */
    const double t2r = j2r * d2r - j2i * d2i, t2i = j2r * d2i + j2i * d2r;
    const double x2r = 0.5 * (s2r - t2r), x2i = 0.5 * (s2i - t2i);
    const double pb = x2r * x2r + x2i * x2i;
    /* Z'[k] = 0.5((Ya + conj Yb) + j (Ya - conj Yb) conj(W)) with real Ya = pa, Yb = pb */
    const double er = pa + pb, od = pa - pb;                /* even real, odd real */
    /* o = od * conj(W) = od*(c - j s) ; j*o = od*(s + j c) */
    w[2 * p1] = 0.5 * (er + od * s); w[2 * p1 + 1] = 0.5 * (od * c);
    if (p2 != p1) {
        /* roles swapped: W_F^{M-k} = (-c, s); conj = (-c, -s); o2 = -od * (-c - j s) = od*(c + j s); j*o2 = od*(-s + j c) */
        w[2 * p2] = 0.5 * (er - od * s); w[2 * p2 + 1] = 0.5 * (od * c);
    }
}

/* --------------------------------------------------------------------------
 * Impulse-response compilation on the device: the per-index bodies of
 *   filter.Reduce      filter/filter.go:520-604  (RealFourier -> resample.Frequency -> RealInverseFourier -> truncate)
 *   filter.Normalize   filter/filter.go:127-138, 328-336 ; Multiply :270-323 ; Add :167-253
 *   the spectrum H of filter.Process (filter.go:395-401), per partition
 * Real transforms of size n go through a complex transform of M = n/2 points in global memory: the tuner's radix-2
 * DIF / DIT passes (ac_dif_body / ac_dit_body, natural -> bit-reversed -> natural) plus the split passes below.
 * -------------------------------------------------------------------------- */
/* zero-padded copy: w[i] = i < n_x ? x[i] : 0, i < n (n real = n/2 complex points, interleaved) */
GDG_DEVI void ir_pad_body(const double *x, long n_x, double *w, long i) { w[i] = i < n_x ? x[i] : 0.0; }
/* forward split: bins X[k], k = 0..M, in natural order from the bit-reversed half-size transform Z (fft.go:801-818, 832-838) */
GDG_DEVI void ir_split_body(const double *w, int M, int logM, double *spec, int k)
{
    if (k == 0 || k == M) {
        const double zr = w[0], zi = w[1];
        spec[2 * k] = k == 0 ? zr + zi : zr - zi; spec[2 * k + 1] = 0.0;
        return;
    }
    const int p1 = brev(k, logM), p2 = brev(M - k, logM);
    const double ar = w[2 * p1], ai = w[2 * p1 + 1], br = w[2 * p2], bi = w[2 * p2 + 1];
    double s, c;
    GDG_SINCOSPI(-(double)k / (double)M, &s, &c);               /* W_n^k, n = 2M */
    const double sr = ar + br, si = ai - bi, dr = ar - br, di = ai + bi;        /* A + conj B, A - conj B */
    const double jr = -s, ji = c;                                               /* j W */
    const double tr = jr * dr - ji * di, ti = jr * di + ji * dr;
    spec[2 * k] = 0.5 * (sr - tr); spec[2 * k + 1] = 0.5 * (si - ti);
}
/* resample.Frequency (resample.go:109-142): Lanczos-3 on the real and imaginary parts, bin i of bins_out at x = i bins_in / bins_out */
GDG_DEVI void ir_resample_bins_body(const double *spec_in, int bins_in, double *spec_out, int bins_out, int i)
{
    const double dx = (double)bins_in / (double)bins_out;
    const double x = (double)i * dx;
    const long idx = (long)floor(x);
    double re = 0.0, im = 0.0;
    for (long t = idx - 2; t < idx + 4; t++)
        if (t >= 0 && t < bins_in) {
            const double df = x - (double)t;
            double kv;
            if (df == 0.0) kv = 1.0;
            else if (-3.0 < df && df < 3.0) { const double pix = GDG_PI * df; kv = (3.0 * (sin(pix) * sin(pix / 3.0))) / (pix * pix); }
            else kv = 0.0;
            re += spec_in[2 * t] * kv; im += spec_in[2 * t + 1] * kv;
        }
    spec_out[2 * i] = re; spec_out[2 * i + 1] = im;
}
/* inverse split (fft.go:899-955): half-size complex input Z'[k] at its bit-reversed slot from the Hermitian half Y[0..M]
 * (only the real parts of Y[0] and Y[M] count, fft.go:908-913) */
GDG_DEVI void ir_unsplit_body(const double *spec, int M, int logM, double *w, int k)
{
    if (k == 0) { const double y0 = spec[0], ym = spec[2 * M]; w[0] = 0.5 * (y0 + ym); w[1] = 0.5 * (y0 - ym); return; }
    const int k2 = M - k;
    const double ar = spec[2 * k], ai = spec[2 * k + 1], br = spec[2 * k2], bi = spec[2 * k2 + 1];
    double s, c;
    GDG_SINCOSPI(-(double)k / (double)M, &s, &c);
    const double er = ar + br, ei = ai - bi, orr = ar - br, oi = ai + bi;       /* Ya + conj Yb, Ya - conj Yb */
    const double pr = orr * c + oi * s, pi = oi * c - orr * s;                  /* (Ya - conj Yb) conj(W) */
    const int p1 = brev(k, logM);
    w[2 * p1] = 0.5 * (er - pi); w[2 * p1 + 1] = 0.5 * (ei + pr);               /* 0.5 (even + j odd) */
}
/* acc[i] = acc[i] + fac * ((comp / sqrt(ss)) * x[i]): Normalize -> Multiply -> Add of one term, element by element in the
 * reference's order of operations (filter.go:328-336, 270-323, 209-215) */
GDG_DEVI void ir_scale_add_body(double *acc, const double *x, double comp, double fac, const double *ss, long i)
{
    const double f = comp / sqrt(ss[0]);
    acc[i] = acc[i] + (fac * (f * x[i]));
}

/* spectrum of partition p of a filter (filter.go:395-401 per 8192-tap / block-length partition): real FFT of size F = 2B of
 * h[pB .. pB + B) zero-padded, through the chain kernel's own shared-memory FFT (half-size complex transform, DIF, natural ->
 * bit-reversed) and a split pass; H: bins 0..B interleaved (re, im), bin k at index brev(k) (0 and B stay). One CTA per partition. */
GDG_DEV void run_partition_spectrum(double *smem, const double *coeffs, int L, int B, int logB, int p, double *H, const double *tw)
{
#if !defined(__CUDACC__) || defined(GDG_HOST_EMU)
    gdg_smem = smem;
#else
    (void)smem;
#endif
    gdg_ctx &cx = *(gdg_ctx *)(gdg_smem + GDG_SMEM_CTX);
    const int M = B, F = 2 * B, fstep = GDG_MAX_FFT / F;
    double *D = SM_R;
    GDG_THREADS_BEGIN
        if (tid == 0) ctx_init(cx, B);
        for (int k = tid; k < GDG_MAX_N / 8; k += GDG_NT) { SM_TWS[2 * k] = GDG_LDG(tw + 4 * k); SM_TWS[2 * k + 1] = GDG_LDG(tw + 4 * k + 1); }
        for (int c = tid; c < M; c += GDG_NT) {
            gdg_c2 v; v.re = 0.0; v.im = 0.0;
            const long i0 = (long)p * B + 2 * c;
            if (2 * c < B) { v.re = i0 < L ? coeffs[i0] : 0.0; v.im = (2 * c + 1 < B && i0 + 1 < L) ? coeffs[i0 + 1] : 0.0; }
            cst(D, cz(c), v);
        }
    GDG_THREADS_END
    fft_forward(cx, D, SM_TWS, logB, 1);
    GDG_THREADS_BEGIN
        for (int k = tid; k <= M; k += GDG_NT) {
            double xr, xi;
            if (k == 0 || k == M) { const gdg_c2 z0 = cld(D, cz(0)); xr = k == 0 ? z0.re + z0.im : z0.re - z0.im; xi = 0.0; }
            else {
                const gdg_c2 za = cld(D, cz(brev(k, logB))), zb = cld(D, cz(brev(M - k, logB)));
                const gdg_c2 wk = tw_full(tw, k * fstep);                         /* W_F^k */
                gdg_c2 jw; jw.re = -wk.im; jw.im = wk.re;
                const gdg_c2 s1 = cadd(za, cconj(zb)), d1 = csub(za, cconj(zb));
                const gdg_c2 t1 = cmul(jw, d1);
                xr = 0.5 * (s1.re - t1.re); xi = 0.5 * (s1.im - t1.im);
            }
            const int at = (k == 0 || k == M) ? k : brev(k, logB);           /* storage order of the chain kernel's transform (stage_poweramp) */
            H[2 * at] = xr; H[2 * at + 1] = xi;
        }
    GDG_THREADS_END
}

#endif
