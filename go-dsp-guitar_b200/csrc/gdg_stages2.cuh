/*
 * gdg_stages2.cuh - the heavier stages: oversampled nonlinear units, octaver,
 * auto_wah, reverb and the power_amp FFT convolution.  See gdg_stages.cuh for
 * the execution model.
 */
#ifndef GDG_STAGES2_CUH
#define GDG_STAGES2_CUH

#include "gdg_stages.cuh"

/* ==========================================================================
 * Oversampled nonlinear units: excess / fuzz / overdrive / distortion
 *   (effects/excess.go, fuzz.go, overdrive.go, distortion.go; the shared
 *    Oversample -> processOversampled -> Decimate wrapper e.g. overdrive.go:83-144)
 *
 *   i[0] factor (1, 2, 4)   i[1] follow mode (fuzz)   i[2] valve (overdrive: 0 ECC82, 1 ECC83, -1 none)
 *   p[0] gainFac p[1] levelFac
 *   fuzz:      p[2] biasFac p[3] fuzzFac p[4] 1-fuzzFac p[5] dInv (at the oversampled rate) p[6] d
 *   overdrive: p[2] driveFac p[3] cleanFac
 *
 * Up-sampling (oversampling.go:54-115 + resample.go:148-176): emitted sample
 * M = f*j + ph of the block equals in[j-4] for ph == 0 and
 * sum_{q<6} in[j-6+q] * L(2 - q + ph/f) otherwise (4-sample look-ahead latency;
 * the last 8 inputs are carried).  Decimation (oversampling.go:126-184 via
 * filter.go:342-515): out[i] = 0.94406 * clamp(sum_k h[k] z[f i - k]) - the
 * reference filters (and clamps) all f*N samples and keeps every f-th; only the
 * kept ones are computed here, as a direct POLYPHASE FIR: with k = f q + p,
 *   y[i] = sum_p sum_{q<39} h[f q + p] * z_p[i - q],   z_p[m] = z[f m - p]
 * (both designs are zero-padded to f*39 taps).  The oversampled signal is kept
 * de-interleaved per phase in shared memory (tile of <= 1024 base samples + a
 * halo of 41); warp w computes phase p = w mod f of GDG_FIR_R consecutive outputs
 * per thread with a register sliding window and the taps as constant-bank
 * immediates; the f partial sums are combined in phase order.
 * ========================================================================== */
#define GDG_FIR_Q 39                         /* taps per polyphase branch */
#define GDG_ZH 40                            /* halo length per phase (>= Q-1), multiple of 8 */
#define GDG_ZPS (((GDG_ZH + GDG_OS_TILE + 1) * 9) / 8 + 8)   /* doubles reserved per phase array >= pz(GDG_ZH + GDG_OS_TILE + 1) + 1 */
#define GDG_OS_Z2 GDG_OS_ST_Z2
GDG_DEVI int pz(int m) { return m + (m >> 3); }      /* one pad per 8: stride-8/9 access patterns spread over banks */
/* location of oversampled sample (tile-local base index jl, sub-phase ph) in the per-phase arrays */
GDG_DEVI int zloc(int f, int jl, int ph, int zps = GDG_ZPS) { return ph == 0 ? pz(GDG_ZH + jl) : (f - ph) * zps + pz(GDG_ZH + jl + 1); }

GDG_DEVI double nl_pointwise(int type, const gdg_unit_desc &d, double s)
{
    if (type == GDG_UNIT_DISTORTION) return d.p[1] * clamp1(d.p[0] * s);        /* distortion.go:36-47 */
    if (type == GDG_UNIT_OVERDRIVE) {                                           /* overdrive.go:59-76 */
        const double arg = d.p[0] * s;
        double dist = 0.0;
        if (d.i[2] == 0) dist = (2.0 / GDG_PI) * atan((0.25 * GDG_PI) * arg);
        else if (d.i[2] == 1) { const double e = exp(-arg); dist = (2.0 / (1.0 + e)) - 1.0; }
        return d.p[1] * ((d.p[2] * dist) + (d.p[3] * s));
    }
    /* excess.go:37-64: triangular wave folding */
    double pre = d.p[0] * s;
    const double ap = fabs(pre);
    const int exceeded = ap > 1.0, negative = pre < 0.0;
    const int section = (int)(0.5 * floor(ap + 1.0));
    const int odd = (section % 2) != 0;
    const int inverted = odd != (exceeded && negative);
    const double tt = ap + 1.0;
    const double ex = tt - 2.0 * floor(0.5 * tt);          /* == fmod(tt, 2.0) exactly (tt >= 1, all steps exact in binary64) */
    if (exceeded) pre = inverted ? 1.0 - ex : ex - 1.0;
    return d.p[1] * pre;
}

/* fuzz.go:52-105 over the F*nb oversampled samples of nb base samples (F == 1: the block itself): follower
 * scan, coupling-capacitor scan, in place.  Thread chunk = Sb consecutive base samples = F*Sb consecutive
 * oversampled samples, visited in time order (base jl, sub-phase ph). */
template <int F, int ZPS> GDG_DEVI int fz_loc(int jl, int ph) { return F == 1 ? sw(jl) : zloc(F, jl, ph, ZPS); }

/* NT: threads of the CTA; ZPS: doubles per polyphase array of z; d: the unit's descriptor (shared memory); st2: {envelope,
 * coupling-capacitor voltage} (device state or a shared-memory copy of it); ws: warp totals of the scans */
template <int F, int NT, int ZPS>
GDG_DEV void fuzz_run(const gdg_unit_desc &d, double *z, int nbase, double *st2, double *ws)
{
    const int Sb = (nbase + NT - 1) / NT, mode = d.i[1];
    const double gain = d.p[0], lvl = d.p[1], bias = d.p[2], ff = d.p[3], ffi = d.p[4], dinv = d.p[5], dd = d.p[6];
    double *st = st2; (void)ws;
    const double astep = follow_A(mode, dinv, dd), afull = pow_int(astep, F * Sb), cstep = 1.0 - dd, cfull = pow_int(cstep, F * Sb);
    const double inite[1] = {st[GDG_OS_ST_ENV]}, initc[1] = {st[GDG_OS_ST_CCV]};
    GDG_REGA(double, a, 1); GDG_REGA(double, b, 1); GDG_REGA(double, ein, 1); GDG_REGA(double, cin, 1);
    GDG_THREADS_BEGIN_N(NT)
        const int j0 = tid * Sb, j1 = imin(nbase, j0 + Sb);
        double e = 0.0;
#pragma unroll 1
        for (int jl = j0; jl < j1; jl++)
#pragma unroll
            for (int ph = 0; ph < F; ph++) e = follow_step(mode, e, fabs(z[fz_loc<F, ZPS>(jl, ph)]), dinv, dd);
        R(a)[0] = j0 < nbase ? chunk_A(astep, F * (j1 - j0), F * Sb, afull) : 1.0; R(b)[0] = j0 < nbase ? e : 0.0;
    GDG_THREADS_END_NOSYNC
    GDG_SCAN_N(1, NT, ws, mode == GDG_FOLLOW_ENVELOPE, a, b, inite, ein);
    GDG_THREADS_BEGIN_N(NT)
        const int j0 = tid * Sb, j1 = imin(nbase, j0 + Sb);
        double e = R(ein)[0], c = 0.0;
#pragma unroll 1
        for (int jl = j0; jl < j1; jl++)
#pragma unroll
            for (int ph = 0; ph < F; ph++) {
                const double s = z[fz_loc<F, ZPS>(jl, ph)];
                e = follow_step(mode, e, fabs(s), dinv, dd);
                double pre = clamp1(gain * (s - (bias * e)));
                pre = (ff * pre) + (ffi * s);
                c += (pre - c) * dd;
            }
        R(a)[0] = j0 < nbase ? chunk_A(cstep, F * (j1 - j0), F * Sb, cfull) : 1.0; R(b)[0] = j0 < nbase ? c : 0.0;
    GDG_THREADS_END_NOSYNC
    GDG_SCAN_N(1, NT, ws, 0, a, b, initc, cin);
    GDG_THREADS_BEGIN_N(NT)
        const int j0 = tid * Sb, j1 = imin(nbase, j0 + Sb);
        double e = R(ein)[0], c = R(cin)[0];
#pragma unroll 1
        for (int jl = j0; jl < j1; jl++)
#pragma unroll
            for (int ph = 0; ph < F; ph++) {
                const int ad = fz_loc<F, ZPS>(jl, ph);
                const double s = z[ad];
                e = follow_step(mode, e, fabs(s), dinv, dd);
                double pre = clamp1(gain * (s - (bias * e)));
                pre = (ff * pre) + (ffi * s);
                const double diff = pre - c;
                c += diff * dd;
                pre -= c;
                z[ad] = lvl * clamp1(pre);
            }
        if (j0 < nbase && j1 == nbase) { st[GDG_OS_ST_ENV] = e; st[GDG_OS_ST_CCV] = c; }
    GDG_THREADS_END
}

/* One polyphase branch (compile-time phase P) of RR consecutive outputs starting at i0: register sliding window
 * over z_p, taps h[F q + P] as constant-bank operands (uniform across the warp). */
template <int F, int P, int RR>
GDG_DEVI void fir_branch(const double *zp, int i0, double (&acc)[RR])
{
    const double *hh = F == 2 ? c_aa2 : c_aa4;
    double win[RR];
#pragma unroll
    for (int r = 0; r < RR; r++) { win[r] = zp[pz(GDG_ZH + i0 + r)]; acc[r] = 0.0; }
#pragma unroll
    for (int q = 0; q < GDG_FIR_Q; q++) {
        const double hq = hh[F * q + P];
#pragma unroll
        for (int r = 0; r < RR; r++) acc[r] += hq * win[r];
        if (q + 1 < GDG_FIR_Q) {
#pragma unroll
            for (int r = RR - 1; r > 0; r--) win[r] = win[r - 1];
            win[0] = zp[pz(GDG_ZH + i0 - (q + 1))];
        }
    }
}

/* up-sampled + shaped samples of base sample j (tile-local jl) into the per-phase arrays */
template <int F>
GDG_DEVI void os_produce(const gdg_unit_desc &d, int type, const double *x, const double *uph, double *z, int j, int jl)
{
    double xs[6];
    if (j >= 6) {
#pragma unroll
        for (int q = 0; q < 6; q++) xs[q] = x[sw(j - 6 + q)];
    } else {
#pragma unroll
        for (int q = 0; q < 6; q++) { const int idx = j - 6 + q; xs[q] = idx >= 0 ? x[sw(idx)] : uph[8 + idx]; }
    }
    const int m0 = pz(GDG_ZH + jl), m1 = pz(GDG_ZH + jl + 1);
#pragma unroll
    for (int ph = 0; ph < F; ph++) {
        double u;
        if (ph == 0) u = xs[2];                                   /* in[j-4] */
        else {
            u = 0.0;
#pragma unroll
            for (int q = 0; q < 6; q++) u += xs[q] * (F == 2 ? c_lz2[0][q] : c_lz4[ph - 1][q]);
        }
        if (type != GDG_UNIT_FUZZ) u = nl_pointwise(type, d, u);
        if (ph == 0) z[m0] = u; else z[(F - ph) * GDG_ZPS + m1] = u;
    }
}

/* finish a tile: y[t0 + i] = 0.944.. * clamp(sum of the f branch sums in branch order), i < cnt (oversampling.go:150-177) */
template <int F>
GDG_DEVI void os_combine(double *y, const double *part, int t0, int cnt, int tid)
{
#pragma unroll 1
    for (int i = tid; i < cnt; i += GDG_NT) {
        double acc = y[sw(t0 + i)];
#pragma unroll
        for (int p = 1; p < F; p++) acc += part[(p - 1) * GDG_OS_TILE + sw(i)];
        y[sw(t0 + i)] = 0.9440608762859234 * clamp1(acc);
    }
}

template <int F>
GDG_DEV void oversampled_tiles(gdg_ctx &cx, const gdg_unit_desc &d_global)
{
    const gdg_unit_desc &d = GDG_DESC; (void)d_global;
    constexpr int RR = (2 * F * 512) / GDG_NT;  /* outputs per thread: GDG_NT / F threads of one phase cover a tile of GDG_OS_TILE outputs */
    const int n = cx.n, type = d.type;
    double *x = SM_CUR(cx), *y = SM_ALT(cx), *z = SM_S, *st = d.st;
    double *uph = st + (F == 2 ? GDG_OS_ST_UP2 : GDG_OS_ST_UP4);
    double *zh = st + (F == 2 ? GDG_OS_ST_Z2 : GDG_OS_ST_Z4);
    double *part = z + F * GDG_ZPS;             /* partial sums of phases 1..F-1: (F-1) * GDG_OS_TILE */
    const int T = imin(n, GDG_OS_TILE);
    const int stateless = type != GDG_UNIT_FUZZ;
    /* carried oversampled history: per phase GDG_ZH + 1 samples; taps to shared memory */
    GDG_THREADS_BEGIN
        if (tid < F * 64) { const int p = tid >> 6, m = tid & 63; if (m <= GDG_ZH) z[p * GDG_ZPS + pz(m)] = zh[p * (GDG_ZH + 1) + m]; }
    GDG_THREADS_END
    for (int t0 = 0; t0 < n; t0 += T) {
        const int Tt = imin(T, n - t0);
        /* up-sample (+ stateless nonlinearity): two base samples per thread for a full tile.  The same phase lands the
         * halo stashed by the previous tile (cells m <= GDG_ZH; this tile writes cells m >= GDG_ZH, and (phase 0,
         * m = GDG_ZH) is this tile's own sample, which the landing skips) and finishes the PREVIOUS tile: partial sums
         * combined in phase order, clamp, -0.5 dB (its latency hides behind the up-sampling; one barrier fewer per tile) */
        GDG_THREADS_BEGIN
            if (t0 > 0 && tid < F * 64) { const int p = tid >> 6, m = tid & 63; if (m < GDG_ZH || (m == GDG_ZH && p > 0)) z[p * GDG_ZPS + pz(m)] = SM_SCANB[tid]; }
#pragma unroll 1
            for (int jl = tid; jl < Tt; jl += GDG_NT) os_produce<F>(d, type, x, uph, z, t0 + jl, jl);
            if (t0 > 0) os_combine<F>(y, part, t0 - T, T, tid);
        GDG_THREADS_END
        if (!stateless) fuzz_run<F, GDG_NT, GDG_ZPS>(d, z, Tt, st, SM_SEG);
        /* polyphase FIR partial sums: warp w computes phase w % F (taps warp-uniform) of RR outputs per thread; then the
         * last GDG_ZH + 1 cells of every phase are stashed for the next tile (z is read-only in this phase) */
        GDG_THREADS_BEGIN
            const int warp = tid >> 5, p = warp % F;
            const int i0 = ((warp / F) * 32 + (tid & 31)) * RR;
            if (i0 < Tt) {
                double acc[RR];
                const double *zp = z + p * GDG_ZPS;
                if (F == 2) { if (p == 0) fir_branch<F, 0, RR>(zp, i0, acc); else fir_branch<F, 1, RR>(zp, i0, acc); }
                else {
                    if (p == 0) fir_branch<F, 0, RR>(zp, i0, acc);
                    else if (p == 1) fir_branch<F, 1 % F, RR>(zp, i0, acc);
                    else if (p == 2) fir_branch<F, 2 % F, RR>(zp, i0, acc);
                    else fir_branch<F, 3 % F, RR>(zp, i0, acc);
                }
#pragma unroll
                for (int r = 0; r < RR; r++)
                    if (i0 + r < Tt) { if (p == 0) y[sw(t0 + i0 + r)] = acc[r]; else part[(p - 1) * GDG_OS_TILE + sw(i0 + r)] = acc[r]; }
            }
            if (t0 + Tt < n && tid < F * 64) { const int pp = tid >> 6, m = tid & 63; if (m <= GDG_ZH) SM_SCANB[tid] = z[pp * GDG_ZPS + pz(m + Tt)]; }
        GDG_THREADS_END
        if (t0 + Tt >= n) {
            /* last tile: finish it, and carry the oversampled history (entries Tt .. Tt + GDG_ZH) and the last 8 inputs
             * (oversampling.go:91-93) */
            GDG_THREADS_BEGIN
                os_combine<F>(y, part, t0, Tt, tid);
                if (tid < F * 64) { const int p = tid >> 6, m = tid & 63; if (m <= GDG_ZH) zh[p * (GDG_ZH + 1) + m] = z[p * GDG_ZPS + pz(m + Tt)]; }
                if (tid < 8) SM_MISC[8 + tid] = (n - 8 + tid) >= 0 ? x[sw(n - 8 + tid)] : uph[tid + n];
            GDG_THREADS_END
        }
    }
    GDG_THREADS_BEGIN
        if (tid < 8) uph[tid] = SM_MISC[8 + tid];
    GDG_THREADS_END
    swap_bufs(cx);
}

GDG_DEV void stage_oversampled(gdg_ctx &cx, const gdg_unit_desc &d_global)
{
    const gdg_unit_desc &d = GDG_DESC; (void)d_global;
    const int n = cx.n, f = d.i[0], type = d.type;
    double *x = SM_CUR(cx);
    if (f <= 1) {
        if (type == GDG_UNIT_FUZZ) { fuzz_run<1, GDG_NT, GDG_ZPS>(d, x, n, d.st, SM_SEG); return; }
        GDG_THREADS_BEGIN
#pragma unroll 1
            for (int k = tid; k < n; k += GDG_NT) x[sw(k)] = nl_pointwise(type, d, x[sw(k)]);
        GDG_THREADS_END
        return;
    }
    /* (a warp-specialised variant - FIR warps / produce warps, one barrier per 512 samples - was measured and was
     * not faster on B200: distortion 0.257 vs 0.253 ms, overdrive 0.455 vs 0.360 ms per 1024-channel block) */
    if (f == 2) oversampled_tiles<2>(cx, d); else oversampled_tiles<4>(cx, d);
}

/* ==========================================================================
 * octaver  (effects/octaver.go:21-139)
 *   i[0] follow; p[0] dInv p[1] d p[2] facUp p[3] facClean p[4] facDist p[5] facDown1 p[6] facDown2 p[7] facHyst
 *   st[0] previousPolarity st[1] octaveRegister (i64) st[2] envelope st[3] coupling
 * polarity after a sample = sign of the last "candidate" (nonzero sample above the
 * hysteresis level); the register counts sign changes between successive candidates.
 * ========================================================================== */
GDG_DEV void stage_octaver(gdg_ctx &cx, const gdg_unit_desc &d_global)
{
    const gdg_unit_desc &d = GDG_DESC; (void)d_global;
    const int n = cx.n, S = chunk_len(n), mode = d.i[0];
    const double dinv = d.p[0], dd = d.p[1], fUp = d.p[2], fClean = d.p[3], fDist = d.p[4], fD1 = d.p[5], fD2 = d.p[6], fH = d.p[7];
    double *x = SM_CUR(cx), *st = d.st;
    double *first = SM_S + GDG_NT, *last = SM_S + 2 * GDG_NT, *flips = SM_S + 3 * GDG_NT,
           *polin = SM_S + 4 * GDG_NT, *regin = SM_S + 5 * GDG_NT;
    const double astep = follow_A(mode, dinv, dd), afull = pow_int(astep, S), cstep = 1.0 - dd, cfull = pow_int(cstep, S);
    const double inite[1] = {st[2]}, initc[1] = {st[3]};
    GDG_REGA(double, a, 1); GDG_REGA(double, b, 1); GDG_REGA(double, ein, 1); GDG_REGA(double, cin, 1);
    GDG_THREADS_BEGIN
        const int k0 = tid * S, k1 = imin(n, k0 + S);
        double e = 0.0;
#pragma unroll 1
        for (int k = k0; k < k1; k++) e = follow_step(mode, e, fabs(x[sw(k)]), dinv, dd);
        R(a)[0] = k0 < n ? chunk_A(astep, k1 - k0, S, afull) : 1.0; R(b)[0] = k0 < n ? e : 0.0;
    GDG_THREADS_END_NOSYNC
    GDG_SCAN(1, cx, mode == GDG_FOLLOW_ENVELOPE, a, b, inite, ein);
    GDG_THREADS_BEGIN
        const int k0 = tid * S, k1 = imin(n, k0 + S);
        double e = R(ein)[0], fs = 0.0, ls = 0.0, fl = 0.0;
#pragma unroll 1
        for (int k = k0; k < k1; k++) {
            const double s = x[sw(k)], sa = fabs(s);
            e = follow_step(mode, e, sa, dinv, dd);
            const double sign = s < 0.0 ? -1.0 : (s > 0.0 ? 1.0 : 0.0);
            if (sign != 0.0 && sa > e * fH) {
                if (fs == 0.0) fs = sign; else if (sign != ls) fl += 1.0;
                ls = sign;
            }
        }
        first[tid] = fs; last[tid] = ls; flips[tid] = fl;
        if (k0 < n && k1 == n) SM_MISC[0] = e;
    GDG_THREADS_END
    GDG_THREADS_BEGIN
        if (tid == 0) {                                /* serial combine over the chunk summaries */
            double pol = st[0];
            long long reg = st_geti(st, 1);
            for (int t = 0; t < GDG_NT; t++) {
                polin[t] = pol; regin[t] = (double)reg;
                if (first[t] != 0.0) {
                    if (first[t] != pol) reg++;
                    reg += (long long)flips[t];
                    pol = last[t];
                }
            }
            SM_MISC[1] = pol; SM_MISC[2] = (double)(reg & 7);
        }
    GDG_THREADS_END
    /* pre-coupling signal -> alt, coupling zero-state response */
    GDG_THREADS_BEGIN
        const int k0 = tid * S, k1 = imin(n, k0 + S);
        double e = R(ein)[0], pol = polin[tid], c = 0.0;
        long long reg = (long long)regin[tid];
#pragma unroll 1
        for (int k = k0; k < k1; k++) {
            const double s = x[sw(k)], sa = fabs(s);
            e = follow_step(mode, e, sa, dinv, dd);
            const double sign = s < 0.0 ? -1.0 : (s > 0.0 ? 1.0 : 0.0);
            if ((sign != 0.0) && (sign != pol) && (sa > e * fH)) { reg = (reg + 1) & 7; pol = sign; }
            const double fd = (reg & 2) ? -1.0 : 1.0, sd = (reg & 4) ? -1.0 : 1.0;
            double pre = fClean * s;
            if (e > 0.0001) pre += fUp * ((s * s) / e);
            pre += fDist * (sign * e);
            pre += fD1 * (fd * e);
            pre += fD2 * (sd * e);
            SM_ALT(cx)[sw(k)] = pre;
            c += (pre - c) * dd;
        }
        R(a)[0] = k0 < n ? chunk_A(cstep, k1 - k0, S, cfull) : 1.0; R(b)[0] = k0 < n ? c : 0.0;
    GDG_THREADS_END_NOSYNC
    GDG_SCAN(1, cx, 0, a, b, initc, cin);
    GDG_THREADS_BEGIN
        const int k0 = tid * S, k1 = imin(n, k0 + S);
        double c = R(cin)[0];
#pragma unroll 1
        for (int k = k0; k < k1; k++) {
            double pre = SM_ALT(cx)[sw(k)];
            c += (pre - c) * dd;
            pre -= c;
            x[sw(k)] = clamp1(pre);
        }
        if (k0 < n && k1 == n) { st[0] = SM_MISC[1]; st_seti(st, 1, (long long)SM_MISC[2]); st[2] = SM_MISC[0]; st[3] = c; }
    GDG_THREADS_END
}

/* ==========================================================================
 * auto_wah  (effects/autowah.go:20-130)
 *   i[0] follow; p[0] dInv p[1] d p[2] levelA p[3] levelB p[4] freqA p[5] freqB p[6] slope p[7] sampleRate p[8] gainCompensation
 *   st[0] envelope st[1..8] hcv st[9..16] lcv
 * k[n] = 1 - exp(-f(env[n])/fs) is stored in alt; the 8 HP/LP pairs are affine
 * recurrences with that time-varying coefficient (the NEW lp value feeds the next pair).
 * ========================================================================== */
GDG_DEV void stage_autowah(gdg_ctx &cx, const gdg_unit_desc &d_global)
{
    const gdg_unit_desc &d = GDG_DESC; (void)d_global;
    const int n = cx.n, S = chunk_len(n), mode = d.i[0];
    const double dinv = d.p[0], dd = d.p[1];
    double *x = SM_CUR(cx), *kv = SM_ALT(cx), *st = d.st;
    const double astep = follow_A(mode, dinv, dd), afull = pow_int(astep, S);
    const double inite[1] = {st[0]};
    GDG_REGA(double, a, 1); GDG_REGA(double, b, 1); GDG_REGA(double, ein, 1); GDG_REGA(double, hin, 1); GDG_REGA(double, lin, 1);
    GDG_REG(double, aprod);
    GDG_THREADS_BEGIN
        const int k0 = tid * S, k1 = imin(n, k0 + S);
        double e = 0.0;
#pragma unroll 1
        for (int k = k0; k < k1; k++) e = follow_step(mode, e, fabs(x[sw(k)]), dinv, dd);
        R(a)[0] = k0 < n ? chunk_A(astep, k1 - k0, S, afull) : 1.0; R(b)[0] = k0 < n ? e : 0.0;
    GDG_THREADS_END_NOSYNC
    GDG_SCAN(1, cx, mode == GDG_FOLLOW_ENVELOPE, a, b, inite, ein);
    GDG_THREADS_BEGIN
        const int k0 = tid * S, k1 = imin(n, k0 + S);
        double e = R(ein)[0], ap = 1.0;
#pragma unroll 1
        for (int k = k0; k < k1; k++) {
            e = follow_step(mode, e, fabs(x[sw(k)]), dinv, dd);
            const double level = 20.0 * log10(e);
            double freq;
            if (level <= d.p[2]) freq = d.p[4];
            else if (level >= d.p[3]) freq = d.p[5];
            else freq = d.p[4] + (d.p[6] * (level - d.p[2]));
            const double kk = 1.0 - exp(-freq / d.p[7]);
            kv[sw(k)] = kk;
            ap *= (1.0 - kk);
        }
        R(aprod) = ap;                                  /* product of (1 - k) over the chunk: the A of every section */
        if (k0 < n && k1 == n) st[0] = e;
    GDG_THREADS_END_NOSYNC
    for (int j = 0; j < 8; j++) {
        const double inith[1] = {st[1 + j]}, initl[1] = {st[9 + j]};
        GDG_THREADS_BEGIN
            const int k0 = tid * S, k1 = imin(n, k0 + S);
            double h = 0.0;
#pragma unroll 1
            for (int k = k0; k < k1; k++) h += (x[sw(k)] - h) * kv[sw(k)];
            R(a)[0] = R(aprod); R(b)[0] = h;
        GDG_THREADS_END_NOSYNC
        GDG_SCAN(1, cx, 0, a, b, inith, hin);
        GDG_THREADS_BEGIN
            const int k0 = tid * S, k1 = imin(n, k0 + S);
            double h = R(hin)[0], l = 0.0;
#pragma unroll 1
            for (int k = k0; k < k1; k++) {
                const double kk = kv[sw(k)];
                double diff = x[sw(k)] - h;
                h += diff * kk;
                diff -= l;
                l += diff * kk;
            }
            R(a)[0] = R(aprod); R(b)[0] = l;
        GDG_THREADS_END_NOSYNC
        GDG_SCAN(1, cx, 0, a, b, initl, lin);
        GDG_THREADS_BEGIN
            const int k0 = tid * S, k1 = imin(n, k0 + S);
            double h = R(hin)[0], l = R(lin)[0];
#pragma unroll 1
            for (int k = k0; k < k1; k++) {
                const double kk = kv[sw(k)];
                double diff = x[sw(k)] - h;
                h += diff * kk;
                diff -= l;
                l += diff * kk;
                x[sw(k)] = (j == 7) ? clamp1(d.p[8] * l) : l;
            }
            if (k0 < n && k1 == n) { st[1 + j] = h; st[9 + j] = l; }
        GDG_THREADS_END_NOSYNC
    }
    GDG_THREADS_BEGIN
    GDG_THREADS_END
}

/* ==========================================================================
 * reverb  (effects/reverb.go:41-60 all-pass, 65-116 delay line, 179-338 Process)
 *   i[0] line ring size (max tap) i[1..4] taps i[5..7] all-pass delays D
 *   i[8] offset of the first all-pass ring inside st (line ring starts at GDG_RV_ST_DATA)
 *   p[0] dryFrac p[1] halfWetFrac p[2..5] tap coefficients p[6] all-pass feedback
 * All-pass (reverb.go:51-56: the read pointer runs one slot ahead of the write pointer in a ring of D):
 *   pre[i] = x[i] - g*pre[i-lag],  out[i] = g*pre[i] + pre[i-lag],  lag = D-1.
 * The block's pre sequence lives in shared memory (P): P[k] first holds the carried history
 * pre[k-lag] (k < lag, fetched from the global ring before the rounds start) and is overwritten
 * by pre[k]; rounds of <= lag samples are independent.  Only the last `lag` values go back to
 * the global ring (a true ring of `lag` slots with its write position).
 * out = clamp((dry x + hw line) + hw wet): the dry/line part is folded first, which differs from
 * reverb.go:318-319 "(dry x) + hw (line + wet)" by one rounding (~1e-16).
 * ========================================================================== */
GDG_DEV void stage_reverb(gdg_ctx &cx, const gdg_unit_desc &d_global)
{
    const gdg_unit_desc &d = GDG_DESC; (void)d_global;
    const int n = cx.n, hs = d.i[0];
    double *x = SM_CUR(cx), *w = SM_ALT(cx), *P = SM_S, *st = d.st;
    double *lring = st + GDG_RV_ST_DATA;
    const double g = d.p[6];
    /* delay line taps -> w.  The first (longest) all-pass's carried history - up to 63 KB from its ring in HBM - is requested
     * here with asynchronous copies into P, which nobody touches until the all-pass rounds start: its latency hides behind
     * the tap gathers. */
    GDG_THREADS_BEGIN
        {
            const int lag0 = d.i[5] - 1, nh0 = imin(lag0, n);
            const double *ring0 = st + d.i[8];
            const int w00 = (int)st_geti(st, GDG_RV_ST_APPTR + 0);
#pragma unroll 4
            for (int k = tid; k < nh0; k += GDG_NT) { int p = w00 + k; if (p >= lag0) p -= lag0; async_copy8(P + k, ring0 + p); }
        }
        const int pos = (int)st_geti(st, GDG_RV_ST_LINEPOS);
        const int t0 = d.i[1], t1 = d.i[2], t2 = d.i[3], t3 = d.i[4];
        const double c0 = d.p[2], c1 = d.p[3], c2 = d.p[4], c3 = d.p[5];
#pragma unroll 4
        for (int k = tid; k < n; k += GDG_NT) {
            /* taps beyond the ring (idx < -hs) read zero; loads are unconditional on clamped addresses */
            const int i0 = k - t0, i1 = k - t1, i2 = k - t2, i3 = k - t3;
            const double v0 = hist_or_block_nb(x, lring, hs, pos, i0 < -hs ? -hs : i0), v1 = hist_or_block_nb(x, lring, hs, pos, i1 < -hs ? -hs : i1);
            const double v2 = hist_or_block_nb(x, lring, hs, pos, i2 < -hs ? -hs : i2), v3 = hist_or_block_nb(x, lring, hs, pos, i3 < -hs ? -hs : i3);
            double pre = 0.0;
            pre += c0 * (i0 < -hs ? 0.0 : v0);
            pre += c1 * (i1 < -hs ? 0.0 : v1);
            pre += c2 * (i2 < -hs ? 0.0 : v2);
            pre += c3 * (i3 < -hs ? 0.0 : v3);
            w[sw(k)] = pre;
        }
    GDG_THREADS_END
    /* push the dry block into the line ring, then fold dry + line */
    GDG_THREADS_BEGIN
        const int pos = (int)st_geti(st, GDG_RV_ST_LINEPOS);
#pragma unroll 1
        for (int k = tid; k < n; k += GDG_NT) {
            const double xv = x[sw(k)];
            if (n >= hs) { if (k >= n - hs) lring[k - (n - hs)] = xv; }
            else { int p = pos + k; if (p >= hs) p -= hs; lring[p] = xv; }
            x[sw(k)] = (d.p[0] * xv) + (d.p[1] * w[sw(k)]);
        }
    GDG_THREADS_END
    int off = d.i[8];
    for (int a = 0; a < 3; a++) {
        const int D = d.i[5 + a], lag = D - 1;
        double *ring = st + off;
        off += lag > 0 ? lag : 0;
        if (lag <= 0) continue;                              /* not reachable: D >= 100 for every supported rate */
        const int w0 = (int)st_geti(st, GDG_RV_ST_APPTR + a);   /* ring slot of the oldest carried value */
        const int nh = imin(lag, n);
        GDG_THREADS_BEGIN
            if (a == 0 && tid == 0) { const int pos = (int)st_geti(st, GDG_RV_ST_LINEPOS); st_seti(st, GDG_RV_ST_LINEPOS, n >= hs ? 0 : (pos + n) % hs); }
            if (a == 0) async_wait();                            /* requested at the top of the stage; each thread asked for the cells it owns here */
            else {
#pragma unroll 4
                for (int k = tid; k < nh; k += GDG_NT) { int p = w0 + k; if (p >= lag) p -= lag; P[k] = GDG_LDST(ring + p); }
            }
        GDG_THREADS_END
        for (int r0 = 0; r0 < n; r0 += lag) {
            const int r1 = imin(n, r0 + lag);
            GDG_THREADS_BEGIN
#pragma unroll 1
                for (int k = r0 + tid; k < r1; k += GDG_NT) {
                    const double dl = k < lag ? P[k] : P[k - lag];
                    const double pre = w[sw(k)] - (g * dl);
                    P[k] = pre;
                    w[sw(k)] = (g * pre) + dl;
                }
            GDG_THREADS_END
        }
        /* carry the last `lag` pre values: ring slot (w0 + k) % lag <- pre[k] for k >= n - lag */
        GDG_THREADS_BEGIN
            for (int k = imax(0, n - lag) + tid; k < n; k += GDG_NT) ring[(w0 + k) % lag] = P[k];
            if (tid == 0) st_seti(st, GDG_RV_ST_APPTR + a, (w0 + n) % lag);
        GDG_THREADS_END
    }
    GDG_THREADS_BEGIN
#pragma unroll 1
        for (int k = tid; k < n; k += GDG_NT) x[sw(k)] = clamp1(x[sw(k)] + (d.p[1] * w[sw(k)]));
    GDG_THREADS_END
}

/* ==========================================================================
 * power_amp: FFT overlap-add FIR  (effects/poweramp.go:186-216 -> filter/filter.go:342-515)
 *   i[0] L (taps; 0 = empty filter -> zeros, -1 = no filter compiled -> zeros)
 *   i[1] F (real FFT size = 2*nextPow2(L), capped at GDG_MAX_FFT)   i[2] log2(F/2)
 *   i[3] K partitions (1 unless L > GDG_MAX_FFT/2)   i[4] word offset of the spectrum delay line in st
 *   aux: H_p, p < K: bins 0..F/2 of RealFourier(zero-padded partition p of h), interleaved (re, im), in the STORAGE order of the
 *        shared-memory transform: bin k at index brev(k) for 0 < k < F/2, bins 0 and F/2 at indices 0 and F/2
 *   st: tail[F], zero-run word, [K > 1: current slot word, then K x (F/2 + 1) complex input spectra]
 * Filters longer than 8192 taps run as a uniformly partitioned convolution (block = partition = 8192):
 * y_t = IFFT(sum_p X_(t-p) H_p) - the same linear convolution the reference evaluates with one
 * 2*nextPow2(L)-point FFT per 8192-sample hop (filter.go:372-455), then the same overlap-add and clamp.
 * Blocks of B = F/2 samples (filter.go:372-388): out = clamp(y[0:ns] + tail[0:ns]),
 * tail <- (y + tail)[ns:F] shifted, zero-filled.  All nb = ceil(n/B) block FFTs sit in
 * R side by side (2B doubles each) and are transformed together.
 * tw: global twiddle table exp(-2 pi i k / GDG_MAX_FFT), k in [0, GDG_MAX_FFT/2)
 * ========================================================================== */
struct gdg_c2 { double re, im; };
GDG_DEVI gdg_c2 cmul(gdg_c2 a, gdg_c2 b) { gdg_c2 r; r.re = a.re * b.re - a.im * b.im; r.im = a.re * b.im + a.im * b.re; return r; }
GDG_DEVI gdg_c2 cadd(gdg_c2 a, gdg_c2 b) { gdg_c2 r; r.re = a.re + b.re; r.im = a.im + b.im; return r; }
GDG_DEVI gdg_c2 csub(gdg_c2 a, gdg_c2 b) { gdg_c2 r; r.re = a.re - b.re; r.im = a.im - b.im; return r; }
GDG_DEVI gdg_c2 cconj(gdg_c2 a) { gdg_c2 r; r.re = a.re; r.im = -a.im; return r; }
GDG_DEVI gdg_c2 cld(const double *p, int i) { gdg_c2 r; r.re = p[2 * i]; r.im = p[2 * i + 1]; return r; }
GDG_DEVI void cst(double *p, int i, gdg_c2 v) { p[2 * i] = v.re; p[2 * i + 1] = v.im; }
GDG_DEVI gdg_c2 twd(const double *tw, int idx) { gdg_c2 r; r.re = GDG_LDG(tw + 2 * idx); r.im = GDG_LDG(tw + 2 * idx + 1); return r; }
GDG_DEVI int brev(int v, int bits)
{
    unsigned int r = (unsigned int)v;
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    r = ((r >> 2) & 0x33333333u) | ((r & 0x33333333u) << 2);
    r = ((r >> 4) & 0x0f0f0f0fu) | ((r & 0x0f0f0f0fu) << 4);
    r = ((r >> 8) & 0x00ff00ffu) | ((r & 0x00ff00ffu) << 8);
    r = (r >> 16) | (r << 16);
    return bits ? (int)(r >> (32 - bits)) : 0;
}

/* complex storage swizzle (16-byte units): every pass - unit stride, stride 8, large strides - touches each
 * 128-byte bank row at most 4 times per warp, the minimum for 32 x 16 B */
GDG_DEVI int cz(int i) { return i ^ ((i >> 3) & 7); }

/* W_16^idx times a (idx compile-time after unrolling); conj selects conj(W) */
GDG_DEVI gdg_c2 mul_w16(gdg_c2 a, int idx, int conj)
{
    const double h = 0.70710678118654752440;
    gdg_c2 r;
    switch (idx) {
    case 0: return a;
    case 4: if (!conj) { r.re = a.im; r.im = -a.re; } else { r.re = -a.im; r.im = a.re; } return r;
    case 2: if (!conj) { r.re = h * (a.re + a.im); r.im = h * (a.im - a.re); } else { r.re = h * (a.re - a.im); r.im = h * (a.im + a.re); } return r;
    case 6: if (!conj) { r.re = h * (a.im - a.re); r.im = -h * (a.re + a.im); } else { r.re = -h * (a.re + a.im); r.im = h * (a.re - a.im); } return r;
    default: {
        /* cos/sin(pi/8), cos/sin(3pi/8): W_16^1 = (c1,-s1), W_16^3 = (s1,-c1), W_16^5 = (-s1,-c1), W_16^7 = (-c1,-s1) */
        const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;
        gdg_c2 w;
        w.re = idx == 1 ? c1 : (idx == 3 ? s1 : (idx == 5 ? -s1 : -c1));
        w.im = idx == 1 ? -s1 : (idx == 3 ? -c1 : (idx == 5 ? -c1 : -s1));
        if (conj) w.im = -w.im;
        return cmul(a, w); }
    }
}
/* in-register radix-2 networks over RR points */
template <int RR>
GDG_DEVI void dif_net(gdg_c2 (&v)[RR])
{
#pragma unroll
    for (int half = RR / 2; half >= 1; half >>= 1)
#pragma unroll
        for (int g = 0; g < RR; g += 2 * half)
#pragma unroll
            for (int pos = 0; pos < half; pos++) {
                const gdg_c2 a = v[g + pos], c = v[g + pos + half];
                v[g + pos] = cadd(a, c);
                v[g + pos + half] = mul_w16(csub(a, c), pos * (8 / half), 0);
            }
}
template <int RR>
GDG_DEVI void dit_net_inv(gdg_c2 (&v)[RR])
{
#pragma unroll
    for (int half = 1; half <= RR / 2; half <<= 1)
#pragma unroll
        for (int g = 0; g < RR; g += 2 * half)
#pragma unroll
            for (int pos = 0; pos < half; pos++) {
                const gdg_c2 a = v[g + pos], t = mul_w16(v[g + pos + half], pos * (8 / half), 1);
                v[g + pos] = cadd(a, t);
                v[g + pos + half] = csub(a, t);
            }
}
GDG_DEVI gdg_c2 tw_full(const double *tw, int idx)          /* exp(-2 pi i idx / GDG_MAX_FFT), idx in [0, GDG_MAX_FFT) */
{
    gdg_c2 r;
    if (idx >= GDG_MAX_FFT / 2) { r = twd(tw, idx - GDG_MAX_FFT / 2); r.re = -r.re; r.im = -r.im; }
    else r = twd(tw, idx);
    return r;
}
template <int RR> GDG_DEVI int brev_small(int k)
{
    int r = 0;
#pragma unroll
    for (int bit = 1, rb = RR >> 1; bit < RR; bit <<= 1, rb >>= 1) if (k & bit) r |= rb;
    return r;
}
/* powers w^1 .. w^(RR-1) with multiplication depth <= log2(RR) */
template <int RR>
GDG_DEVI void tw_powers(gdg_c2 w1, gdg_c2 (&pw)[RR])
{
    pw[0].re = 1.0; pw[0].im = 0.0; pw[1] = w1;
#pragma unroll
    for (int q = 2; q < RR; q++) {
        const int hb = q >= 8 ? 8 : (q >= 4 ? 4 : 2);       /* highest power of two <= q */
        pw[q] = (q == hb) ? cmul(pw[hb / 2], pw[hb / 2]) : cmul(pw[hb], pw[q - hb]);
    }
}

/* One radix-RR pass of span L = 2^logL over nseg transforms of M = 2^logM complex points held (swizzled) in D.
 * Forward (inverse == 0): decimation in frequency: RR-point DFT of the points j + k L/RR, then slot k (which
 * holds output brev(k)) times W_L^(j brev(k)); natural in -> bit-reversed out over all passes.
 * Inverse: the mirror image with conjugate twiddles (bit-reversed in -> natural out), unnormalised. */
template <int RR>
GDG_DEV void fft_pass(double *D, const double *tws, int logM, int nseg, int logL, int inverse)
{
    constexpr int LOGR = RR == 16 ? 4 : (RR == 8 ? 3 : (RR == 4 ? 2 : 1));
    const int logS = logL - LOGR;                              /* stride L/RR */
    const int nb = nseg << (logM - LOGR);
    const int tshift = GDG_LOG_MAX_N - logL;                              /* W_L^j = tws[j << tshift]: shared table exp(-2 pi i k / 8192), k < 1024 */
    GDG_THREADS_BEGIN
#pragma unroll 1
        for (int b = tid; b < nb; b += GDG_NT) {
            const int seg = b >> (logM - LOGR), rb = b & ((1 << (logM - LOGR)) - 1);
            const int blk = rb >> logS, j = rb & ((1 << logS) - 1);
            const int base = (seg << logM) + (blk << logL) + j;
            gdg_c2 v[RR];
#pragma unroll
            for (int k = 0; k < RR; k++) v[k] = cld(D, cz(base + (k << logS)));
            gdg_c2 w1 = cld(tws, j << tshift);
            if (inverse) w1.im = -w1.im;
            if (!inverse) dif_net<RR>(v);
            if (RR == 16) {
                /* slot k holds output brev(k): powers w^1..w^15 from w, w^2, w^4, w^8 (4 complex registers) instead of a
                 * 16-entry table next to the 16 points - the radix-16 pass no longer spills at 128 registers */
                const gdg_c2 w2 = cmul(w1, w1), w4 = cmul(w2, w2), w8 = cmul(w4, w4);
#pragma unroll
                for (int k = 1; k < RR; k++) {
                    const int q = brev_small<RR>(k);
                    gdg_c2 t; t.re = 1.0; t.im = 0.0; bool first = true;
                    if (q & 8) { t = w8; first = false; }
                    if (q & 4) { t = first ? w4 : cmul(t, w4); first = false; }
                    if (q & 2) { t = first ? w2 : cmul(t, w2); first = false; }
                    if (q & 1) { t = first ? w1 : cmul(t, w1); first = false; }
                    v[k] = cmul(v[k], t);
                }
            } else {
                gdg_c2 pw[RR];
                tw_powers<RR>(w1, pw);
#pragma unroll
                for (int k = 1; k < RR; k++) v[k] = cmul(v[k], pw[brev_small<RR>(k)]);
            }
            if (inverse) dit_net_inv<RR>(v);
#pragma unroll
            for (int k = 0; k < RR; k++) cst(D, cz(base + (k << logS)), v[k]);
        }
    GDG_THREADS_END
}
GDG_DEVI void fft_pass_any(double *D, const double *tw, int logM, int nseg, int logL, int r, int inverse, gdg_ctx &cx)
{
    (void)cx;
    if (r == 4) fft_pass<16>(D, tw, logM, nseg, logL, inverse);
    else if (r == 3) fft_pass<8>(D, tw, logM, nseg, logL, inverse);
    else if (r == 2) fft_pass<4>(D, tw, logM, nseg, logL, inverse);
    else fft_pass<2>(D, tw, logM, nseg, logL, inverse);
}
/* radix plan: ceil(logM / 4) passes of radix 16 / 8 / 4 / 2, the larger radices first (13 -> 4+3+3+3, 12 -> 4+4+4) */
GDG_DEVI int fft_plan_radix(int logM, int np, int i) { return logM / np + (i < logM % np ? 1 : 0); }
GDG_DEV void fft_forward(gdg_ctx &cx, double *D, const double *tw, int logM, int nseg)
{
    const int np = (logM + 3) / 4;
    int logL = logM;
#pragma unroll 1
    for (int i = 0; i < np; i++) { const int r = fft_plan_radix(logM, np, i); fft_pass_any(D, tw, logM, nseg, logL, r, 0, cx); logL -= r; }
}
GDG_DEV void fft_inverse(gdg_ctx &cx, double *D, const double *tw, int logM, int nseg)
{
    const int np = (logM + 3) / 4;
    int logL = 0;
#pragma unroll 1
    for (int i = np - 1; i >= 0; i--) { const int r = fft_plan_radix(logM, np, i); logL += r; fft_pass_any(D, tw, logM, nseg, logL, r, 1, cx); }
}

/* power_amp state: tail[F] followed by one word = number of trailing tail cells known to be zero */
GDG_DEV void stage_poweramp(gdg_ctx &cx, const gdg_unit_desc &d_global, const double *tw)
{
    const gdg_unit_desc &d = GDG_DESC; (void)d_global;
    const int n = cx.n, L = d.i[0];
    double *x = SM_CUR(cx);
    if (L <= 0) {                                              /* filter.go:366-367 / poweramp.go:210-214 */
        GDG_THREADS_BEGIN
#pragma unroll 1
            for (int k = tid; k < n; k += GDG_NT) x[sw(k)] = 0.0;
        GDG_THREADS_END
        return;
    }
    const int F = d.i[1], M = F / 2, logM = d.i[2], B = M;
    const int nb = n <= B ? 1 : n / B;                           /* n and B are powers of two */
    const int ns = n <= B ? n : B;                               /* samples per block */
    const int total = nb * M;                                    /* complex points in flight (<= GDG_MAX_N) */
    double *D = SM_R, *tail = d.st;
    const double *tws = SM_TWS;                                  /* 1024 complex: exp(-2 pi i k / 8192), loaded by run_chain */
    /* split twiddles indexed by STORAGE (bit-reversed) position r = r_hi 2^s + r_lo: W_F^brev(r) = TA[r_lo] TB[r_hi] with
     * TA[x] = W_F^(brev_s(x) 2^(logM-s)), TB[y] = W_F^brev_(logM-s)(y): a warp reads consecutive TA entries and one TB entry */
    const int tsplit = imin(7, logM);
    double *wc = SM_SCANA, *wf = SM_SCANA + 256;
    double *tl = SM_S;                                           /* prefetched overlap tail (plain layout) */
    const double *H = d.aux;
    const int zlen = (int)st_geti(tail, F);                      /* trailing zeros guaranteed in tail[] */
    const int K = d.i[3] > 1 ? d.i[3] : 1;                       /* partitions */
    double *fdl = d.st + d.i[4];                                 /* K x (M + 1) complex input spectra, slot ring */
    const int slot = K > 1 ? (int)st_geti(tail, F + 1) : 0;
    const long hstride = 2L * (M + 1);
    const int fstep = GDG_MAX_FFT / F;                           /* W_F^k = tw[k * MAXFFT/F] */
    /* pack: block j -> complex points (x[j ns + 2c], x[j ns + 2c + 1]), c < ns/2, zero padded to M.  The destination
     * overlays the source, so the block first moves to S in plain order; the split twiddles are set up on the way. */
    GDG_THREADS_BEGIN
#pragma unroll 4
        for (int k = tid; k < n; k += GDG_NT) tl[k] = x[sw(k)];
        if (tid < (1 << tsplit)) cst(wc, tid, tw_full(tw, (brev(tid, tsplit) << (logM - tsplit)) * fstep));
        if (tid >= 128 && tid < 128 + (1 << (logM - tsplit))) cst(wf, tid - 128, tw_full(tw, brev(tid - 128, logM - tsplit) * fstep));
    GDG_THREADS_END
    GDG_THREADS_BEGIN
#pragma unroll 2
        for (int g = tid; g < total; g += GDG_NT) {
            const int j = g >> logM, c = g & (M - 1);
            gdg_c2 v; v.re = 0.0; v.im = 0.0;
            if (2 * c < ns) { v.re = tl[j * ns + 2 * c]; v.im = (2 * c + 1 < ns) ? tl[j * ns + 2 * c + 1] : 0.0; }
            cst(D, cz(g), v);
        }
    GDG_THREADS_END
    /* overlap cells [0, ns) of the tail are needed after the inverse transform: request them now - one bulk asynchronous
     * copy straight into S (nobody waits here; the transform's barriers separate this from the wait below); cells known
     * to be zero are not read */
    const int tl_live = imax(0, imin(ns, F - zlen));            /* cells of tail[0, ns) that may be non-zero */
    const bool tl_bulk = GDG_HAVE_BULK && tl_live > 0 && (tl_live & 1) == 0;
    double *mbar = SM_MISC + 32;
    GDG_THREADS_BEGIN
        if (tl_bulk) { if (tid == 0) bulk_load(mbar, tl, tail, (unsigned)(tl_live * 8)); }
        else {
#pragma unroll 4
            for (int q = tid; q < tl_live; q += GDG_NT) async_copy8(tl + q, tail + q);
        }
        for (int q = tl_live + tid; q < ns; q += GDG_NT) tl[q] = 0.0;
    GDG_THREADS_END_NOSYNC
    fft_forward(cx, D, tws, logM, nb);
    /* split (fft.go:801-818) * H (filter.go:459) * inverse split (fft.go:899-955), on bit-reversed storage */
    GDG_THREADS_BEGIN
        /* bins 0 and M (both real) of every block: one thread per block */
        if (tid < nb) {
            const int zb = tid * M;
            {
                const gdg_c2 z0 = cld(D, cz(zb));
                const double x0 = z0.re + z0.im, xm = z0.re - z0.im;     /* X[0], X[M] (real) */
                double y0 = x0 * GDG_LDG(H + 0), ym = xm * GDG_LDG(H + 2 * M);  /* real(dc), real(nyquist) of the product */
                if (K > 1) {
                    for (int p = 1; p < K; p++) {
                        const double *X = fdl + (long)((slot - p + K) % K) * hstride, *Hp = H + (long)p * hstride;
                        y0 += GDG_LDST(X + 0) * GDG_LDG(Hp + 0); ym += GDG_LDST(X + 2 * M) * GDG_LDG(Hp + 2 * M);
                    }
                    double *Xc = fdl + (long)slot * hstride;
                    Xc[0] = x0; Xc[1] = 0.0; Xc[2 * M] = xm; Xc[2 * M + 1] = 0.0;
                }
                gdg_c2 o; o.re = 0.5 * (y0 + ym); o.im = 0.5 * (y0 - ym);
                cst(D, cz(zb), o);
            }
        }
        /* the conjugate pairs (k, M - k) sit at storage positions (r, r ^ (2^j - 1)) inside the block [2^j, 2^(j+1)) of the
         * bit-reversed order: walking r instead of k makes both shared-memory accesses unit-stride across the warp
         * (walking k, they were 8-way bank-conflicted) */
#pragma unroll 2
        for (int b = tid; b < nb * (M / 2); b += GDG_NT) {
            const int seg = b / (M / 2), bb = b - seg * (M / 2);
            const int zb = seg * M;
            {
                int r = 1, r2 = 1;
                if (bb > 0) { const int j = ilog2_floor(bb) + 1; r = (1 << j) + (bb - (1 << (j - 1))); r2 = r ^ ((1 << j) - 1); }
                /* spectra (H, and the delay line of partitioned filters) are stored in the same storage order: coalesced */
                gdg_c2 ha, hb; ha.re = GDG_LDG(H + 2 * r); ha.im = GDG_LDG(H + 2 * r + 1);
                hb.re = GDG_LDG(H + 2 * r2); hb.im = GDG_LDG(H + 2 * r2 + 1);
                const int p1 = cz(zb + r), p2 = cz(zb + r2);
                const gdg_c2 za = cld(D, p1), zbv = cld(D, p2);
                const gdg_c2 wk = cmul(cld(wc, r & ((1 << tsplit) - 1)), cld(wf, r >> tsplit));   /* W_F^k, k = brev(r) */
                gdg_c2 jw; jw.re = -wk.im; jw.im = wk.re;                 /* j * W_F^k */
                const gdg_c2 s1 = cadd(za, cconj(zbv)), d1 = csub(za, cconj(zbv));
                const gdg_c2 t1 = cmul(jw, d1);
                gdg_c2 xa; xa.re = 0.5 * (s1.re - t1.re); xa.im = 0.5 * (s1.im - t1.im);
                const gdg_c2 s2 = cadd(zbv, cconj(za)), d2 = csub(zbv, cconj(za));
                const gdg_c2 t2 = cmul(cconj(jw), d2);                    /* j W_F^(M-k) = conj(j W_F^k) */
                gdg_c2 xb; xb.re = 0.5 * (s2.re - t2.re); xb.im = 0.5 * (s2.im - t2.im);
                gdg_c2 ya = cmul(xa, ha), yb = cmul(xb, hb);
                if (K > 1) {
                    for (int p = 1; p < K; p++) {
                        const double *X = fdl + (long)((slot - p + K) % K) * hstride, *Hp = H + (long)p * hstride;
                        gdg_c2 xp, hp;
                        xp.re = GDG_LDST(X + 2 * r); xp.im = GDG_LDST(X + 2 * r + 1); hp.re = GDG_LDG(Hp + 2 * r); hp.im = GDG_LDG(Hp + 2 * r + 1);
                        ya = cadd(ya, cmul(xp, hp));
                        xp.re = GDG_LDST(X + 2 * r2); xp.im = GDG_LDST(X + 2 * r2 + 1); hp.re = GDG_LDG(Hp + 2 * r2); hp.im = GDG_LDG(Hp + 2 * r2 + 1);
                        yb = cadd(yb, cmul(xp, hp));
                    }
                    double *Xc = fdl + (long)slot * hstride;
                    Xc[2 * r] = xa.re; Xc[2 * r + 1] = xa.im; Xc[2 * r2] = xb.re; Xc[2 * r2 + 1] = xb.im;
                }
                const gdg_c2 e1 = cadd(ya, cconj(yb)), o1 = cmul(csub(ya, cconj(yb)), cconj(wk));
                gdg_c2 ra; ra.re = 0.5 * (e1.re - o1.im); ra.im = 0.5 * (e1.im + o1.re);
                gdg_c2 wk2; wk2.re = -wk.re; wk2.im = wk.im;               /* W_F^(M-k) */
                const gdg_c2 e2 = cadd(yb, cconj(ya)), o2 = cmul(csub(yb, cconj(ya)), cconj(wk2));
                gdg_c2 rb; rb.re = 0.5 * (e2.re - o2.im); rb.im = 0.5 * (e2.im + o2.re);
                cst(D, p1, ra);
                if (p2 != p1) cst(D, p2, rb);
            }
        }
    GDG_THREADS_END
    fft_inverse(cx, D, tws, logM, nb);
    /* real sample r of block j = component (r & 1) of complex point r >> 1 */
#define GDG_YR(j, r) D[2 * cz((j) * M + ((r) >> 1)) + ((r) & 1)]
    const double scale = 1.0 / (double)M;
    /* overlap-add + clamp, staged in S (the output buffer overlays the FFT workspace); running tail staged in place (filter.go:473-503): for nb >= 2 the
     * tail beyond ns is zero after the first block, so only the single-block case adds the old tail there.
     * Cells of tail[] known to be zero (the zero-run word) are neither read nor rewritten. */
    GDG_THREADS_BEGIN
        if (tl_bulk) bulk_wait(mbar); else async_wait();           /* the tail cells are in S now */
#pragma unroll 2
        for (int k = tid; k < n; k += GDG_NT) {
            {
                const int j = k / ns, q = k - j * ns;
                double pre = scale * GDG_YR(j, q);
                if (j == 0) pre += tl[q];
                else { pre += scale * GDG_YR(j - 1, ns + q); if (j == 1 && ns + q < F - zlen) pre += tail[ns + q]; }
                tl[k] = clamp1(pre);                                 /* staged in S: cell k is read (j == 0) and written by this thread only */
            }
        }
#pragma unroll 1
        for (int q = tid; q < F - ns; q += GDG_NT) {
            double v = scale * GDG_YR(nb - 1, ns + q);
            if (nb == 1 && ns + q < F - zlen) v += tail[ns + q];
            GDG_YR(nb - 1, ns + q) = v;
        }
        if (tid == 0) { cx.cur_off = GDG_SMEM_R; cx.alt_off = GDG_SMEM_R + GDG_MAX_N; }   /* the output goes to the first half of R */
    GDG_THREADS_END
    GDG_THREADS_BEGIN
#pragma unroll 1
        for (int q = tid; q < F - ns; q += GDG_NT) tail[q] = GDG_YR(nb - 1, ns + q);
        for (int q = F - ns + tid; q < F - zlen; q += GDG_NT) tail[q] = 0.0;
        if (tid == 0) { st_seti(tail, F, ns); if (K > 1) st_seti(tail, F + 1, (slot + 1) % K); }
        double *out = SM_CUR(cx);
#pragma unroll 4
        for (int k = tid; k < n; k += GDG_NT) out[sw(k)] = tl[k];
    GDG_THREADS_END
#undef GDG_YR
}

#endif /* GDG_STAGES2_CUH */
