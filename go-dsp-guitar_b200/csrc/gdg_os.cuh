/*
 * gdg_os.cuh - the "oversampled run" kernel body: a maximal run of consecutive oversampled nonlinear units
 * (excess / fuzz / overdrive / distortion with oversampling "2" or "4") of one channel, streamed tile by tile.
 *
 * Why a second launch shape: inside the block-resident chain kernel (one 512-thread CTA per SM, 220 KB of
 * shared memory, 128 registers per thread) these units were half of the run time although their polyphase FIR already
 * sits at the FP64 pipe limit: the up-sampling / waveshaper / scan / combine phases are latency-bound and, with a
 * single CTA per SM, nothing else can use the pipe while they wait.  This body needs only a tile (512 base samples) of
 * its signal in shared memory (~50 KB), so four 256-thread CTAs share an SM: twice the resident warps, and the phases
 * of different CTAs overlap (one CTA's FIR fills the FP64 pipe while another one up-samples).  All units of the run are
 * applied to a tile before the next tile is touched; intermediate signals never leave shared memory.
 *
 * Arithmetic, state words (GDG_OS_ST_*) and carries are exactly those of stage_oversampled (gdg_stages2.cuh): the two
 * code paths are interchangeable block by block (a unit may move between them when the chain is edited).
 *
 * Two drivers share the unit tile (osk_unit_tile): run_os - one CTA per channel, every unit of the run on a tile before the
 * next tile (many channels: the channels fill the GPU) - and run_os_pipe at the end of this file - a thread-block cluster
 * per channel, one CTA per unit, tiles handed from CTA to CTA through distributed shared memory (few channels: the latency
 * of a block is what counts).
 *
 * Reference: effects/{excess,fuzz,overdrive,distortion}.go Process wrappers (e.g. overdrive.go:83-144),
 * oversampling/oversampling.go:54-184, resample/resample.go:148-176, filter/filter.go:342-515 (the decimator's FIR).
 */
#ifndef GDG_OS_CUH
#define GDG_OS_CUH

#include "gdg_stages2.cuh"

#ifndef GDG_OSK_NT
#define GDG_OSK_NT 256                          /* threads per CTA */
#endif
#ifndef GDG_OSK_CTAS
#define GDG_OSK_CTAS 4                          /* resident CTAs per SM the kernel is compiled for */
#endif
#ifndef GDG_OSK_TILE
#define GDG_OSK_TILE 512                        /* base-rate samples per tile */
#endif
#define GDG_OSK_MAX_UNITS 6                     /* units per run (longer runs are split by the host) */
#define GDG_OSK_ZPS (((GDG_ZH + GDG_OSK_TILE + 1) * 9) / 8 + 8)      /* doubles per polyphase array */
#define GDG_OSK_STASH (4 * (GDG_ZH + 1) + 8)    /* per unit: carried halo of 4 branches + the last 8 inputs */

/* shared-memory carve-up (offsets in doubles off gdg_smem).  The unit input / output buffers XA, XB are PLAIN ([8 history |
 * tile]): every access to them is unit-stride across the warp, and the up-sampler's six taps are immediate offsets off one
 * address.  The FIR writes its F branch sums as chunks of 8 consecutive outputs per thread (stride 8 across the warp), so
 * those go to PART in the one-pad-per-8 layout (pz, stride 9: conflict-free) and the combine pass - unit-stride again - sums
 * them into the plain output buffer. */
#define GDG_OSK_XLEN (8 + GDG_OSK_TILE)                                /* [8 history | tile] */
#define GDG_OSK_PLEN ((GDG_OSK_TILE * 9) / 8 + 2)
#define GDG_OSK_XA 0
#define GDG_OSK_XB (GDG_OSK_XA + GDG_OSK_XLEN)
#define GDG_OSK_Z (GDG_OSK_XB + GDG_OSK_XLEN)                          /* 4 polyphase arrays */
#define GDG_OSK_PART (GDG_OSK_Z + 4 * GDG_OSK_ZPS)                     /* branch sums 0..3 */
#define GDG_OSK_ST (GDG_OSK_PART + 4 * GDG_OSK_PLEN)                   /* per unit: halo + input history */
#define GDG_OSK_FZ (GDG_OSK_ST + GDG_OSK_MAX_UNITS * GDG_OSK_STASH)    /* per unit: fuzz envelope, coupling capacitor */
#define GDG_OSK_WS (GDG_OSK_FZ + 2 * GDG_OSK_MAX_UNITS)                /* scan warp totals */
#define GDG_OSK_DESC (GDG_OSK_WS + 2 * (GDG_OSK_NT / 32))              /* nslots + GDG_OSK_MAX_UNITS unit descriptors */
#define GDG_OSK_DESC_DOUBLES ((int)((8 + GDG_OSK_MAX_UNITS * sizeof(gdg_unit_desc) + 7) / 8))
#define GDG_OSK_SMEM_DOUBLES (GDG_OSK_DESC + GDG_OSK_DESC_DOUBLES)
#define GDG_OSK_SMEM_BYTES (GDG_OSK_SMEM_DOUBLES * 8)

/* waveshaper parameters in registers (the descriptor sits in shared memory, which the compiler must re-read after every
 * store to z otherwise) */
struct gdg_nl_par { double g, lv, a, b; int type, valve; };
GDG_DEVI gdg_nl_par nl_load(const gdg_unit_desc &d)
{
    gdg_nl_par p; p.g = d.p[0]; p.lv = d.p[1]; p.a = d.p[2]; p.b = d.p[3]; p.type = d.type; p.valve = d.i[2];
    return p;
}
template <int TYPE>
GDG_DEVI double nl_apply(const gdg_nl_par &p, double s)
{
    if (TYPE == GDG_UNIT_DISTORTION) return p.lv * clamp1(p.g * s);           /* distortion.go:36-47 */
    if (TYPE == GDG_UNIT_OVERDRIVE) {                                          /* overdrive.go:59-76 */
        const double arg = p.g * s;
        double dist = 0.0;
        if (p.valve == 0) dist = (2.0 / GDG_PI) * atan((0.25 * GDG_PI) * arg);
        else if (p.valve == 1) { const double e = exp(-arg); dist = (2.0 / (1.0 + e)) - 1.0; }
        return p.lv * ((p.a * dist) + (p.b * s));
    }
    double pre = p.g * s;                                                      /* excess.go:37-64 */
    const double ap = fabs(pre);
    const int exceeded = ap > 1.0, negative = pre < 0.0;
    const int section = (int)(0.5 * floor(ap + 1.0));
    const int odd = (section % 2) != 0;
    const int inverted = odd != (exceeded && negative);
    const double tt = ap + 1.0;
    const double ex = tt - 2.0 * floor(0.5 * tt);            /* == fmod(tt, 2.0) exactly (tt >= 1) */
    if (exceeded) pre = inverted ? 1.0 - ex : ex - 1.0;
    return p.lv * pre;
}

/* One polyphase branch (run-time branch index p) of 8 consecutive outputs starting at i0, i0 a multiple of 8:
 *   acc[r] = sum_q h[F q + p] z_p[i0 + r - q].
 * The window z_p[i0 + r - q], r < 8, lives in 8 registers used as a ring (slot (r - q) mod 8), so a step of 8 taps
 * needs no register moves and the tap loop stays ROLLED: 5 x 8 taps (both tap tables are zero-padded to 40 per branch),
 * ~90 instructions that stay in the instruction cache - the fully unrolled, branch-specialised form (4 x 5.6 KB) made
 * instruction fetch the top stall of this kernel.  Every z cell sits at a fixed offset from zb = zp + pz(GDG_ZH + i0)
 * (cell t of the window at t + floor(t / 8)): LDS with immediate offsets off a pointer that moves by 9 per step. */
template <int F>
GDG_DEVI void fir_branch_rolled(const double *zb, int p, double (&acc)[8])
{
    const double *hh = (F == 2 ? c_aa2 : c_aa4) + p;
    double win[8];
#pragma unroll
    for (int r = 0; r < 8; r++) { win[r] = zb[r]; acc[r] = 0.0; }
#pragma unroll 1
    for (int it = 0; it < 5; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const double hq = hh[F * (8 * it + j)];                  /* q = 8 it + j; q = 39 is the zero padding */
#pragma unroll
            for (int r = 0; r < 8; r++) acc[r] += hq * win[(r - j) & 7];
            if (j < 7 || it < 4) win[(7 - j) & 7] = zb[-9 * it - j - 2];   /* cell -(q+1): -(q+1) - ceil((q+1)/8) */
        }
    }
}

/* fuzz.go:52-105 on a tile of the oversampled signal, for this kernel's thread shape: a thread owns the TWO base samples
 * 2 tid and 2 tid + 1 (F oversampled samples each) and walks them as two INDEPENDENT recurrence chains (the second one
 * starts from the first one's chunk map applied to the carry) - twice the instruction-level parallelism of one chain of 2F
 * samples in the latency-bound follower / coupling-capacitor replays; all loads of z are issued before the chains start.
 * Same scans, same update formulas as fuzz_run. */
template <int MODE> GDG_DEVI double follow_step_t(double env, double abs_s, double d_inv, double d)
{
    if (MODE == GDG_FOLLOW_ENVELOPE) { env *= d_inv; if (abs_s > env) env = abs_s; return env; }
    if (MODE == GDG_FOLLOW_LEVEL) { const double diff = abs_s - env; return env + diff * d; }
    return 1.0;
}
template <int F, int NT, int ZPS, int MODE>
GDG_DEV void fuzz_run_pair(const gdg_unit_desc &d, double *z, int nbase, double *st, double *ws)
{
    constexpr int MP = MODE == GDG_FOLLOW_ENVELOPE;
    const double gain = d.p[0], lvl = d.p[1], bias = d.p[2], ff = d.p[3], ffi = d.p[4], dinv = d.p[5], dd = d.p[6];
    const double astep = follow_A(MODE, dinv, dd), aF = pow_int(astep, F), cstep = 1.0 - dd, cF = pow_int(cstep, F);
    const double inite[1] = {st[GDG_OS_ST_ENV]}, initc[1] = {st[GDG_OS_ST_CCV]};
    (void)ws;
    GDG_REGA(double, a, 1); GDG_REGA(double, b, 1); GDG_REGA(double, ein, 1); GDG_REGA(double, cin, 1);
    GDG_REGA(double, ea, 2); GDG_REGA(double, eb, 2); GDG_REGA(double, ca, 2); GDG_REGA(double, cb, 2);
    GDG_THREADS_BEGIN_N(NT)
        const int j0 = 2 * tid;
        double e[2] = {0.0, 0.0}, xv[2][F];
#pragma unroll
        for (int s = 0; s < 2; s++)
#pragma unroll
            for (int ph = 0; ph < F; ph++) xv[s][ph] = j0 + s < nbase ? z[zloc(F, j0 + s, ph, ZPS)] : 0.0;
#pragma unroll
        for (int ph = 0; ph < F; ph++)
#pragma unroll
            for (int s = 0; s < 2; s++) if (j0 + s < nbase) e[s] = follow_step_t<MODE>(e[s], fabs(xv[s][ph]), dinv, dd);
#pragma unroll
        for (int s = 0; s < 2; s++) { R(ea)[s] = j0 + s < nbase ? aF : 1.0; R(eb)[s] = j0 + s < nbase ? e[s] : 0.0; }
        R(a)[0] = R(ea)[1] * R(ea)[0]; R(b)[0] = scan_apply(MP, R(ea)[1], R(eb)[0], R(eb)[1]);
    GDG_THREADS_END_NOSYNC
    GDG_SCAN_N(1, NT, ws, MP, a, b, inite, ein);
    GDG_THREADS_BEGIN_N(NT)
        const int j0 = 2 * tid;
        double e[2], c[2] = {0.0, 0.0}, xv[2][F];
        e[0] = R(ein)[0]; e[1] = scan_apply(MP, R(ea)[0], R(ein)[0], R(eb)[0]);
#pragma unroll
        for (int s = 0; s < 2; s++)
#pragma unroll
            for (int ph = 0; ph < F; ph++) xv[s][ph] = j0 + s < nbase ? z[zloc(F, j0 + s, ph, ZPS)] : 0.0;
#pragma unroll
        for (int ph = 0; ph < F; ph++)
#pragma unroll
            for (int s = 0; s < 2; s++) if (j0 + s < nbase) {
                const double x = xv[s][ph];
                e[s] = follow_step_t<MODE>(e[s], fabs(x), dinv, dd);
                double pre = clamp1(gain * (x - (bias * e[s])));
                pre = (ff * pre) + (ffi * x);
                c[s] += (pre - c[s]) * dd;
            }
#pragma unroll
        for (int s = 0; s < 2; s++) { R(ca)[s] = j0 + s < nbase ? cF : 1.0; R(cb)[s] = j0 + s < nbase ? c[s] : 0.0; }
        R(a)[0] = R(ca)[1] * R(ca)[0]; R(b)[0] = (R(ca)[1] * R(cb)[0]) + R(cb)[1];
    GDG_THREADS_END_NOSYNC
    GDG_SCAN_N(1, NT, ws, 0, a, b, initc, cin);
    GDG_THREADS_BEGIN_N(NT)
        const int j0 = 2 * tid;
        double e[2], c[2], xv[2][F];
        e[0] = R(ein)[0]; e[1] = scan_apply(MP, R(ea)[0], R(ein)[0], R(eb)[0]);
        c[0] = R(cin)[0]; c[1] = (R(ca)[0] * R(cin)[0]) + R(cb)[0];
#pragma unroll
        for (int s = 0; s < 2; s++)
#pragma unroll
            for (int ph = 0; ph < F; ph++) xv[s][ph] = j0 + s < nbase ? z[zloc(F, j0 + s, ph, ZPS)] : 0.0;
#pragma unroll
        for (int ph = 0; ph < F; ph++)
#pragma unroll
            for (int s = 0; s < 2; s++) if (j0 + s < nbase) {
                const double x = xv[s][ph];
                e[s] = follow_step_t<MODE>(e[s], fabs(x), dinv, dd);
                double pre = clamp1(gain * (x - (bias * e[s])));
                pre = (ff * pre) + (ffi * x);
                const double diff = pre - c[s];
                c[s] += diff * dd;
                pre -= c[s];
                z[zloc(F, j0 + s, ph, ZPS)] = lvl * clamp1(pre);
            }
        if (j0 < nbase && j0 + 2 >= nbase) { const bool l = j0 + 1 < nbase; st[GDG_OS_ST_ENV] = l ? e[1] : e[0]; st[GDG_OS_ST_CCV] = l ? c[1] : c[0]; }
    GDG_THREADS_END
}

/* one unit of the run on one tile: xin = [8 history | Tt samples] -> xout (same layout); plain layout */
template <int F, int TYPE>
GDG_DEV void osk_unit_tile(const gdg_unit_desc &d, const double *xin, double *xout, double *stash, double *fz, int Tt, const double *next_hist)
{
    constexpr int NT = GDG_OSK_NT, ZPS = GDG_OSK_ZPS, T = GDG_OSK_TILE; (void)T;
    constexpr int RR = (T * F) / NT;                /* outputs per thread in the FIR phase: NT / (32 F) warps of 32 lanes per branch */
    double *z = gdg_smem + GDG_OSK_Z, *part = gdg_smem + GDG_OSK_PART, *ws = gdg_smem + GDG_OSK_WS;
    /* land the carried halo (cells m <= GDG_ZH; this tile writes cells m >= GDG_ZH and (branch 0, m = GDG_ZH) is its own
     * sample), up-sample (resample.go:148-176 with constant Lanczos weights) and shape */
    GDG_THREADS_BEGIN_N(NT)
        if (tid < F * 64) { const int p = tid >> 6, m = tid & 63; if (m < GDG_ZH || (m == GDG_ZH && p > 0)) z[p * ZPS + pz(m)] = stash[p * (GDG_ZH + 1) + m]; }
        const gdg_nl_par par = nl_load(d);
#pragma unroll (TYPE == GDG_UNIT_OVERDRIVE ? 1 : 2)
        for (int jl = tid; jl < Tt; jl += NT) {
            double xs[6];
#pragma unroll
            for (int q = 0; q < 6; q++) xs[q] = xin[jl + 2 + q];                /* in[j - 6 + q] */
            const int m0 = pz(GDG_ZH + jl), m1 = pz(GDG_ZH + jl + 1);
#pragma unroll
            for (int ph = 0; ph < F; ph++) {
                double u;
                if (ph == 0) u = xs[2];                                         /* in[j - 4] */
                else {
                    u = 0.0;
#pragma unroll
                    for (int q = 0; q < 6; q++) u += xs[q] * (F == 2 ? c_lz2[0][q] : c_lz4[ph - 1][q]);
                }
                if (TYPE != GDG_UNIT_FUZZ) u = nl_apply<TYPE>(par, u);
                if (ph == 0) z[m0] = u; else z[(F - ph) * ZPS + m1] = u;
            }
        }
    GDG_THREADS_END
    if (TYPE == GDG_UNIT_FUZZ) {
        static_assert(GDG_OSK_TILE <= 2 * GDG_OSK_NT, "fuzz_run_pair: two base samples per thread");
        const int mode = d.i[1];
        if (mode == GDG_FOLLOW_ENVELOPE) fuzz_run_pair<F, NT, ZPS, GDG_FOLLOW_ENVELOPE>(d, z, Tt, fz, ws);
        else if (mode == GDG_FOLLOW_LEVEL) fuzz_run_pair<F, NT, ZPS, GDG_FOLLOW_LEVEL>(d, z, Tt, fz, ws);
        else fuzz_run_pair<F, NT, ZPS, GDG_FOLLOW_OTHER>(d, z, Tt, fz, ws);
    }
    /* polyphase FIR branch sums (warp w: branch w % F), then stash the halo and the input history for the next tile */
    GDG_THREADS_BEGIN_N(NT)
        const int warp = tid >> 5, p = warp % F;
        const int i0 = ((warp / F) * 32 + (tid & 31)) * RR;
        if (i0 < Tt) {
            double acc[RR];
            const double *zp = z + p * ZPS;
            if constexpr (F == 2) { if (p == 0) fir_branch<F, 0, RR>(zp, i0, acc); else fir_branch<F, 1, RR>(zp, i0, acc); }
            else fir_branch_rolled<F>(zp + pz(GDG_ZH + i0), p, acc);   /* RR == 8: i0 is a multiple of 8 */
            double *dst = part + p * GDG_OSK_PLEN + pz(i0);
            if (RR == 8 && i0 + RR <= Tt) {
#pragma unroll
                for (int r = 0; r < RR; r++) dst[r] = acc[r];            /* 8 | i0: the 8 cells are contiguous */
            } else {
#pragma unroll
                for (int r = 0; r < RR; r++) if (i0 + r < Tt) part[p * GDG_OSK_PLEN + pz(i0 + r)] = acc[r];
            }
        }
        if (tid < F * 64) { const int pp = tid >> 6, m = tid & 63; if (m <= GDG_ZH) stash[pp * (GDG_ZH + 1) + m] = z[pp * ZPS + pz(m + Tt)]; }
        if (tid >= NT - 8) { const int k = tid - (NT - 8); stash[4 * (GDG_ZH + 1) + k] = xin[Tt + k]; }     /* last 8 of [history | tile] */
    GDG_THREADS_END
    /* combine in branch order, clamp, -0.5 dB (oversampling.go:150-177); the next unit's input history goes in front */
    GDG_THREADS_BEGIN_N(NT)
#pragma unroll 1
        for (int i = tid; i < Tt; i += NT) {
            const int pi = pz(i);
            double acc = part[pi];
#pragma unroll
            for (int p = 1; p < F; p++) acc += part[p * GDG_OSK_PLEN + pi];
            xout[8 + i] = 0.9440608762859234 * clamp1(acc);
        }
        if (next_hist && tid < 8) xout[tid] = next_hist[tid];
    GDG_THREADS_END
}

GDG_DEVI void osk_dispatch(const gdg_unit_desc &d, const double *xin, double *xout, double *sh, double *fz, int Tt, const double *nh)
{
    const int f = d.i[0];
    switch (d.type) {                           /* one instantiation per (factor, unit): only the running one is hot in the instruction cache */
    case GDG_UNIT_EXCESS: if (f == 2) osk_unit_tile<2, GDG_UNIT_EXCESS>(d, xin, xout, sh, fz, Tt, nh); else osk_unit_tile<4, GDG_UNIT_EXCESS>(d, xin, xout, sh, fz, Tt, nh); break;
    case GDG_UNIT_FUZZ: if (f == 2) osk_unit_tile<2, GDG_UNIT_FUZZ>(d, xin, xout, sh, fz, Tt, nh); else osk_unit_tile<4, GDG_UNIT_FUZZ>(d, xin, xout, sh, fz, Tt, nh); break;
    case GDG_UNIT_OVERDRIVE: if (f == 2) osk_unit_tile<2, GDG_UNIT_OVERDRIVE>(d, xin, xout, sh, fz, Tt, nh); else osk_unit_tile<4, GDG_UNIT_OVERDRIVE>(d, xin, xout, sh, fz, Tt, nh); break;
    default: if (f == 2) osk_unit_tile<2, GDG_UNIT_DISTORTION>(d, xin, xout, sh, fz, Tt, nh); else osk_unit_tile<4, GDG_UNIT_DISTORTION>(d, xin, xout, sh, fz, Tt, nh); break;
    }
}

/* one channel, one block: the whole run */
GDG_DEV void run_os(double *smem, const gdg_chain_desc *cd, const void *in, int in_fmt, void *out, int out_fmt, double *side, int n)
{
#if !defined(__CUDACC__) || defined(GDG_HOST_EMU)
    gdg_smem = smem;
#else
    (void)smem;
#endif
    constexpr int NT = GDG_OSK_NT, T = GDG_OSK_TILE;
    double *xa = gdg_smem + GDG_OSK_XA, *xb = gdg_smem + GDG_OSK_XB;
    double *stash = gdg_smem + GDG_OSK_ST, *fz = gdg_smem + GDG_OSK_FZ;
    const gdg_chain_desc *scd = (const gdg_chain_desc *)(gdg_smem + GDG_OSK_DESC);
    /* descriptor and carried state -> shared memory */
    GDG_THREADS_BEGIN_N(NT)
        const int words = (int)((8 + (size_t)imin(GDG_LDG(&cd->nslots), GDG_OSK_MAX_UNITS) * sizeof(gdg_unit_desc)) / 4);
        for (int k = tid; k < words; k += NT) ((int *)(gdg_smem + GDG_OSK_DESC))[k] = ((const int *)cd)[k];
    GDG_THREADS_END
    const int U = imin(scd->nslots, GDG_OSK_MAX_UNITS);
    GDG_THREADS_BEGIN_N(NT)
        for (int u = 0; u < U; u++) {
            const gdg_unit_desc &d = scd->slot[u];
            const int F = d.i[0];
            const double *st = d.st;
            double *sh = stash + u * GDG_OSK_STASH;
            const double *zh = st + (F == 2 ? GDG_OS_ST_Z2 : GDG_OS_ST_Z4), *uph = st + (F == 2 ? GDG_OS_ST_UP2 : GDG_OS_ST_UP4);
            if (tid < F * (GDG_ZH + 1)) sh[tid] = GDG_LDST(zh + tid);
            if (tid >= NT - 8) sh[4 * (GDG_ZH + 1) + tid - (NT - 8)] = GDG_LDST(uph + tid - (NT - 8));
            if (tid == 32) { fz[2 * u] = st[GDG_OS_ST_ENV]; fz[2 * u + 1] = st[GDG_OS_ST_CCV]; }
        }
    GDG_THREADS_END
    /* the input tile travels through registers: tile t+1 is requested from global memory while tile t is processed */
    GDG_REGA(double, nx, GDG_OSK_TILE / GDG_OSK_NT);
    GDG_THREADS_BEGIN_N(NT)
#pragma unroll
        for (int k = 0; k < T / NT; k++) { const int i = tid + k * NT; R(nx)[k] = i < n ? (in_fmt == GDG_FMT_IEEE64 ? ((const double *)in)[i] : decode_sample(in_fmt, (const unsigned char *)in, i)) : 0.0; }
    GDG_THREADS_END_NOSYNC
    for (int t0 = 0; t0 < n; t0 += T) {
        const int Tt = imin(T, n - t0);
        GDG_THREADS_BEGIN_N(NT)
#pragma unroll
            for (int k = 0; k < T / NT; k++) { const int i = tid + k * NT; if (i < Tt) xa[8 + i] = R(nx)[k]; }
            if (tid < 8) xa[tid] = stash[4 * (GDG_ZH + 1) + tid];
#pragma unroll
            for (int k = 0; k < T / NT; k++) {
                const int i = t0 + T + tid + k * NT;
                if (i < n) R(nx)[k] = in_fmt == GDG_FMT_IEEE64 ? ((const double *)in)[i] : decode_sample(in_fmt, (const unsigned char *)in, i);
            }
        GDG_THREADS_END
        double *xin = xa, *xout = xb;
        for (int u = 0; u < U; u++) {
            const gdg_unit_desc &d = scd->slot[u];
            double *sh = stash + u * GDG_OSK_STASH;
            const double *nh = u + 1 < U ? stash + (u + 1) * GDG_OSK_STASH + 4 * (GDG_ZH + 1) : nullptr;
            osk_dispatch(d, xin, xout, sh, fz + 2 * u, Tt, nh);
            double *t = xin; xin = xout; xout = t;
        }
        GDG_THREADS_BEGIN_N(NT)                                   /* xin holds the run's output tile now */
            if (out_fmt == GDG_FMT_IEEE64) { for (int i = tid; i < Tt; i += NT) ((double *)out)[t0 + i] = xin[8 + i]; }
            else if (side == nullptr) { for (int i = tid; i < Tt; i += NT) encode_sample(out_fmt, (unsigned char *)out, t0 + i, xin[8 + i]); }
            else { for (int i = tid; i < Tt; i += NT) { const double v = xin[8 + i]; side[t0 + i] = v; encode_sample(out_fmt, (unsigned char *)out, t0 + i, v); } }
        GDG_THREADS_END
    }
    /* carried state back to the units' state words */
    GDG_THREADS_BEGIN_N(NT)
        for (int u = 0; u < U; u++) {
            const gdg_unit_desc &d = scd->slot[u];
            const int F = d.i[0];
            double *st = d.st;
            const double *sh = stash + u * GDG_OSK_STASH;
            double *zh = st + (F == 2 ? GDG_OS_ST_Z2 : GDG_OS_ST_Z4), *uph = st + (F == 2 ? GDG_OS_ST_UP2 : GDG_OS_ST_UP4);
            if (tid < F * (GDG_ZH + 1)) zh[tid] = sh[tid];
            if (tid >= NT - 8) uph[tid - (NT - 8)] = sh[4 * (GDG_ZH + 1) + tid - (NT - 8)];
            if (tid == 32 && d.type == GDG_UNIT_FUZZ) { st[GDG_OS_ST_ENV] = fz[2 * u]; st[GDG_OS_ST_CCV] = fz[2 * u + 1]; }
        }
    GDG_THREADS_END
}

#if defined(__CUDACC__) && !defined(GDG_HOST_EMU)
/* --------------------------------------------------------------------------------------------------------------------
 * The same run as a PIPELINE over a thread-block cluster (few channels: one CTA per channel leaves the GPU idle and a block
 * costs the latency of U x ntiles unit-tiles in sequence).  CTA r of the cluster owns unit r of the run; in step s it
 * applies its unit to tile s - r and writes the result straight into the input buffer of CTA r + 1 - distributed shared
 * memory, no trip through L2 - then the cluster synchronises: U + ntiles - 1 steps instead of U x ntiles.  Every unit sees
 * exactly the tiles, carries and arithmetic of run_os (osk_unit_tile is shared), so the two are interchangeable block by
 * block.  Input buffers alternate between XA and XB (the producer of step s writes the buffer the consumer reads in step
 * s + 1); the last unit stages its output in the stash area of the unused units.
 * Launch: cluster of P >= U CTAs per channel (ranks >= U only take part in the barriers). */
#include <cooperative_groups.h>
GDG_DEV void run_os_pipe(const gdg_chain_desc *cd, const void *in, int in_fmt, void *out, int out_fmt, double *side, int n)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cl = cg::this_cluster();
    constexpr int NT = GDG_OSK_NT, T = GDG_OSK_TILE;
    const int r = (int)cl.block_rank(), P = (int)cl.num_blocks();
    const int tid = (int)threadIdx.x;
    const int U = imin(GDG_LDG(&cd->nslots), imin(GDG_OSK_MAX_UNITS, P));
    const bool mine = r < U;
    double *const xb0 = gdg_smem + GDG_OSK_XA, *const xb1 = gdg_smem + GDG_OSK_XB;
    double *xc = gdg_smem + GDG_OSK_ST + GDG_OSK_STASH;                 /* output staging of the last unit (stash slots 1.. are free here) */
    static_assert((GDG_OSK_MAX_UNITS - 1) * GDG_OSK_STASH >= GDG_OSK_XLEN, "staging buffer does not fit the unused stash slots");
    double *stash = gdg_smem + GDG_OSK_ST, *fz = gdg_smem + GDG_OSK_FZ;
    gdg_chain_desc *scd = (gdg_chain_desc *)(gdg_smem + GDG_OSK_DESC);
    if (mine) {                                                          /* my unit's descriptor -> slot 0 of the shared-memory copy */
        const int words = (int)(sizeof(gdg_unit_desc) / 4);
        for (int k = tid; k < words; k += NT) ((int *)&scd->slot[0])[k] = ((const int *)&cd->slot[r])[k];
    }
    __syncthreads();
    const gdg_unit_desc &d = scd->slot[0];
    if (mine) {
        const int F = d.i[0];
        const double *st = d.st;
        const double *zh = st + (F == 2 ? GDG_OS_ST_Z2 : GDG_OS_ST_Z4), *uph = st + (F == 2 ? GDG_OS_ST_UP2 : GDG_OS_ST_UP4);
        if (tid < F * (GDG_ZH + 1)) stash[tid] = GDG_LDST(zh + tid);
        if (tid >= NT - 8) stash[4 * (GDG_ZH + 1) + tid - (NT - 8)] = GDG_LDST(uph + tid - (NT - 8));
        if (tid == 32) { fz[0] = st[GDG_OS_ST_ENV]; fz[1] = st[GDG_OS_ST_CCV]; }
    }
    const int ntiles = (n + T - 1) / T;
    double nx[T / NT];
    if (r == 0) {
#pragma unroll
        for (int k = 0; k < T / NT; k++) { const int i = tid + k * NT; nx[k] = i < n ? (in_fmt == GDG_FMT_IEEE64 ? ((const double *)in)[i] : decode_sample(in_fmt, (const unsigned char *)in, i)) : 0.0; }
    }
    cl.sync();                                                           /* every CTA of the cluster runs: remote shared memory may be written */
    for (int s = 0; s < ntiles + U - 1; s++) {
        const int t = s - r;
        if (mine && t >= 0 && t < ntiles) {
            const int t0 = t * T, Tt = imin(T, n - t0);
            double *xin = (s & 1) ? xb1 : xb0;
            if (r == 0) {
#pragma unroll
                for (int k = 0; k < T / NT; k++) { const int i = tid + k * NT; if (i < Tt) xin[8 + i] = nx[k]; }
#pragma unroll
                for (int k = 0; k < T / NT; k++) {
                    const int i = t0 + T + tid + k * NT;
                    if (i < n) nx[k] = in_fmt == GDG_FMT_IEEE64 ? ((const double *)in)[i] : decode_sample(in_fmt, (const unsigned char *)in, i);
                }
            }
            if (tid < 8) xin[tid] = stash[4 * (GDG_ZH + 1) + tid];           /* my unit's last 8 inputs */
            __syncthreads();
            const bool last = r == U - 1;
            double *xout = last ? xc : cl.map_shared_rank((s & 1) ? xb0 : xb1, r + 1);
            osk_dispatch(d, xin, xout, stash, fz, Tt, nullptr);
            if (last) {
                if (out_fmt == GDG_FMT_IEEE64) { for (int i = tid; i < Tt; i += NT) ((double *)out)[t0 + i] = xc[8 + i]; }
                else if (side == nullptr) { for (int i = tid; i < Tt; i += NT) encode_sample(out_fmt, (unsigned char *)out, t0 + i, xc[8 + i]); }
                else { for (int i = tid; i < Tt; i += NT) { const double v = xc[8 + i]; side[t0 + i] = v; encode_sample(out_fmt, (unsigned char *)out, t0 + i, v); } }
            }
        }
        cl.sync();
    }
    if (mine) {
        const int F = d.i[0];
        double *st = d.st;
        double *zh = st + (F == 2 ? GDG_OS_ST_Z2 : GDG_OS_ST_Z4), *uph = st + (F == 2 ? GDG_OS_ST_UP2 : GDG_OS_ST_UP4);
        if (tid < F * (GDG_ZH + 1)) zh[tid] = stash[tid];
        if (tid >= NT - 8) uph[tid - (NT - 8)] = stash[4 * (GDG_ZH + 1) + tid - (NT - 8)];
        if (tid == 32 && d.type == GDG_UNIT_FUZZ) { st[GDG_OS_ST_ENV] = fz[0]; st[GDG_OS_ST_CCV] = fz[1]; }
    }
}
#endif

#endif /* GDG_OS_CUH */
