/*
 * gdg_stages.cuh - CTA-cooperative float64 "stages": one per effects unit.
 *
 * Execution model.  One CTA (GDG_NT threads) owns one channel's audio block
 * (n <= 8192 samples) which stays in shared memory across the whole fused chain
 * segment.  Every stage is a sequence of PHASES; a phase is a pure function of
 * (thread id, shared/global memory) and phases are separated by CTA barriers.
 * Thread-private values never live across a phase boundary - they go through
 * shared memory - so the very same source also compiles as plain C++ (macro
 * GDG_HOST_EMU: a phase becomes a loop over thread ids).  That build exists only
 * for tests/ (CPU-side check of the parallel formulation against the oracle);
 * the product library contains the CUDA build alone.
 *
 * How the reference's strictly sequential per-sample loops become parallel:
 *   - one-pole RC sections and "level" followers  v += (x - v) k  are affine
 *     recurrences: each thread owns a contiguous chunk, computes the chunk's
 *     zero-state response, a block-level scan yields every chunk's carry-in, and
 *     the chunk is then replayed with the reference's own update formula;
 *   - "envelope" followers  e = max(|x|, d e)  are max-plus recurrences: same scheme;
 *   - threshold state machines (noise gate, octaver register, tremolo) become
 *     "last event" / counting scans or closed forms in the sample index;
 *   - delay-line units are gathers from [history ring | block];
 *   - the reverb all-passes (lag D-1 recurrences) run in rounds of D-1 samples;
 *   - oversampled nonlinear units run tile-wise: polyphase Lanczos-3 up-sampling
 *     (constant weights), nonlinearity, direct polyphase FIR decimation that only
 *     evaluates the retained every-f-th output (+ clamp, -0.5 dB);
 *   - power_amp is FFT overlap-add convolution with the whole FFT in shared memory
 *     (real transform through a half-size complex transform, DIF forward / DIT
 *     inverse so no bit-reversal pass is needed).
 *
 * Reference citations are /root/reference paths (file:line).
 */
#ifndef GDG_STAGES_CUH
#define GDG_STAGES_CUH

#include <math.h>
#include <stdint.h>
#include "gdg_device.h"

#if defined(__CUDACC__) && !defined(GDG_HOST_EMU)
#define GDG_DEV __device__ __noinline__
#define GDG_DEVI __device__ __forceinline__
#define GDG_THREADS_BEGIN { const int tid = (int)threadIdx.x; (void)tid;
#define GDG_THREADS_BEGIN_N(NT_) { const int tid = (int)threadIdx.x; (void)tid;       /* CTA of NT_ threads */
#define GDG_THREADS_END } __syncthreads();
#define GDG_THREADS_END_NOSYNC }
#define GDG_CONST __constant__
#define GDG_LDG(p) __ldg(p)
/* coherent load from GLOBAL memory (unit state): tells the compiler the address space, so that the loads can be
 * hoisted over shared-memory stores and batched */
#define GDG_LDST(p) __ldcg(p)
/* c + a*b with BOTH roundings, as the reference computes it (no FMA contraction): used where the result feeds a
 * discontinuity (phase wrap, square / sawtooth edge) or is carried as state, so that it stays bit-identical */
#define GDG_MULADD_RN(a, b, c) __dadd_rn(__dmul_rn((a), (b)), (c))
#define GDG_SINCOS(x, s, c) sincos((x), (s), (c))
/* 8-byte global -> shared copy that does not pass through registers (LDGSTS): the issuing thread does not wait for it
 * until async_wait(); a thread only ever reads back cells it requested itself, so no barrier is involved */
static __device__ __forceinline__ void async_copy8(double *smem_dst, const double *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
static __device__ __forceinline__ void async_wait() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }
/* Bulk asynchronous copy global -> shared (TMA engine, UBLKCP) signalled through an mbarrier word in shared memory.
 * ONE thread calls bulk_load(); a __syncthreads() must separate that call from the first bulk_wait() of any other
 * thread (the barrier word is initialised by the issuing thread).  src, dst and bytes are multiples of 16. */
static __device__ __forceinline__ void bulk_load(double *mbar_word, double *smem_dst, const void *gsrc, unsigned bytes)
{
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(mbar_word);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes), "r"(mbar) : "memory");
}
static __device__ __forceinline__ void bulk_wait(double *mbar_word)
{
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(mbar_word);
    asm volatile("{\n\t.reg .pred p;\n\tGDG_BULK_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@!p bra GDG_BULK_WAIT;\n\t}" :: "r"(mbar) : "memory");
}
#define GDG_HAVE_BULK 1
/* thread-private value that stays in a register across phases */
#define GDG_REG(type, name) type name
#define GDG_REGA(type, name, cnt) type name[cnt]
#define R(name) name
#else
#define GDG_DEV static
#define GDG_DEVI static inline
#define GDG_THREADS_BEGIN for (int tid = 0; tid < GDG_NT; ++tid) {
#define GDG_THREADS_BEGIN_N(NT_) for (int tid = 0; tid < (NT_); ++tid) {
#define GDG_THREADS_END }
#define GDG_THREADS_END_NOSYNC }
#define GDG_CONST static
#define GDG_LDG(p) (*(p))
#define GDG_LDST(p) (*(p))
#define GDG_MULADD_RN(a, b, c) (((a) * (b)) + (c))          /* the emulation is built with -ffp-contract=off */
#define GDG_SINCOS(x, s, c) do { *(s) = sin(x); *(c) = cos(x); } while (0)
static inline void async_copy8(double *dst, const double *src) { *dst = *src; }
static inline void async_wait() {}
static inline void bulk_load(double *, double *dst, const void *src, unsigned bytes) { memcpy(dst, src, bytes); }
static inline void bulk_wait(double *) {}
#define GDG_HAVE_BULK 0
/* host emulation: one slot per emulated thread */
#define GDG_REG(type, name) type name##_r[GDG_NT]
#define GDG_REGA(type, name, cnt) type name##_r[GDG_NT][cnt]
#define R(name) name##_r[tid]
#endif

#define GDG_PI 3.14159265358979323846
#define GDG_TWO_PI (2.0 * GDG_PI)
#define GDG_NSEG (GDG_NT / 32)

/* ---- constant tables (filled by the host at start-up) --------------------- */
GDG_CONST double c_aa2[80];            /* 77-tap x2 anti-aliasing FIR (oversampling.go:239-317) */
GDG_CONST double c_aa4[160];           /* 155-tap x4 anti-aliasing FIR (oversampling.go:357-513) */
GDG_CONST double c_lz2[1][6];          /* Lanczos-3 weights of the fractional phases, x2 (resample.go:10-66) */
GDG_CONST double c_lz4[3][6];          /* ... x4: phase p/4, taps q: L(2 - q + p/4) */

/* ---- execution context: shared-memory carve-up (offsets in doubles) -----------
 * Everything is addressed off the CTA's one dynamic shared array so that the
 * compiler emits LDS/STS (not generic loads) in the separately compiled stages. */
#define GDG_SMEM_R 0                                         /* 2*GDG_MAX_N: bufA | bufB; also the FFT workspace */
#define GDG_SMEM_S (2 * GDG_MAX_N)                           /* GDG_MAX_N scratch */
#define GDG_SMEM_TWS (GDG_SMEM_S + GDG_MAX_N)                /* 1024 complex: exp(-2 pi i k / 8192), loaded once per launch */
#define GDG_SMEM_SCANB (GDG_SMEM_TWS + GDG_MAX_N / 4)                 /* 256 doubles staging */
#define GDG_SMEM_SCANA (GDG_SMEM_SCANB + 256)                /* 512 doubles: split twiddles / staging */
#define GDG_SMEM_SEG (GDG_SMEM_SCANA + 512)                  /* warp totals of the scans */
#define GDG_SMEM_MISC (GDG_SMEM_SEG + 3 * GDG_SCAN_LANES * GDG_NSEG)
#define GDG_SMEM_DESC (GDG_SMEM_MISC + 64)                    /* the running unit's descriptor (24 doubles) */
#define GDG_SMEM_CTX (GDG_SMEM_DESC + 24)                    /* gdg_ctx of the running block (2 doubles) */
#define GDG_SMEM_CHAIN (GDG_SMEM_CTX + 2)                    /* the whole chain descriptor, staged once per launch */
#define GDG_SMEM_DOUBLES (GDG_SMEM_CHAIN + (int)((sizeof(gdg_chain_desc) + 7) / 8))
#define GDG_SMEM_BYTES (GDG_SMEM_DOUBLES * 8)

#if defined(__CUDACC__) && !defined(GDG_HOST_EMU)
extern __shared__ __align__(16) double gdg_smem[];
#else
static thread_local double *gdg_smem;                        /* host emulation: set by run_chain (per thread: concurrent calls) */
#endif

struct gdg_ctx {
    int cur_off;      /* current block (swizzled layout): GDG_SMEM_R or GDG_SMEM_R + GDG_MAX_N */
    int alt_off;      /* the other half */
    int n;
};
#define SM_R (gdg_smem + GDG_SMEM_R)
#define SM_S (gdg_smem + GDG_SMEM_S)
#define SM_TWS (gdg_smem + GDG_SMEM_TWS)
#define SM_SCANB (gdg_smem + GDG_SMEM_SCANB)
#define SM_SCANA (gdg_smem + GDG_SMEM_SCANA)
#define SM_SEG (gdg_smem + GDG_SMEM_SEG)
#define SM_MISC (gdg_smem + GDG_SMEM_MISC)
/* the unit descriptor is staged in shared memory by run_chain: parameter reads inside the stages are LDS, not
 * global loads that the compiler may not hoist over stores */
#define GDG_DESC (*(const gdg_unit_desc *)(gdg_smem + GDG_SMEM_DESC))
#define SM_CUR(cx) (gdg_smem + (cx).cur_off)
#define SM_ALT(cx) (gdg_smem + (cx).alt_off)

/* XOR swizzle of the block buffers: a thread that walks its own 16-sample chunk
 * (stride-16 across the warp) and a warp that reads 32 consecutive samples are
 * both bank-conflict-free. */
GDG_DEVI int sw(int i) { return i ^ ((i >> 4) & 15); }

GDG_DEVI double clamp1(double x) { return x < -1.0 ? -1.0 : (x > 1.0 ? 1.0 : x); }
GDG_DEVI int imin(int a, int b) { return a < b ? a : b; }
GDG_DEVI int imax(int a, int b) { return a > b ? a : b; }
GDG_DEVI long long st_geti(const double *st, int k) { return ((const long long *)st)[k]; }
GDG_DEVI void st_seti(double *st, int k, long long v) { ((long long *)st)[k] = v; }
GDG_DEVI int chunk_len(int n) { return (n + GDG_NT - 1) / GDG_NT; }
GDG_DEVI int ilog2_floor(int v)                   /* floor(log2(v)), v >= 1 */
{
#if defined(__CUDACC__) && !defined(GDG_HOST_EMU)
    return 31 - __clz(v);
#else
    return 31 - __builtin_clz((unsigned)v);
#endif
}
/* the context lives in shared memory (a private copy passed by reference to the non-inlined stages would sit in
 * local memory, and there is next to no L1 left beside the 220 KB carve-out): one thread flips it between barriers.
 * Every stage ends with a barrier before it swaps, so no thread still reads the old orientation. */
GDG_DEVI void swap_bufs(gdg_ctx &cx)
{
    GDG_THREADS_BEGIN
        if (tid == 0) { const int t = cx.cur_off; cx.cur_off = cx.alt_off; cx.alt_off = t; }
    GDG_THREADS_END
}

GDG_DEVI double pow_int(double a, int k)          /* a^k by binary exponentiation (k >= 0) */
{
    double r = 1.0;
#pragma unroll 1
    while (k > 0) { if (k & 1) r *= a; a *= a; k >>= 1; }
    return r;
}

/* compressor.go:44-58 (shared by autowah/autoyoy/octaver/fuzz) */
GDG_DEVI double follow_step(int mode, double env, double abs_s, double d_inv, double d)
{
    if (mode == GDG_FOLLOW_ENVELOPE) { env *= d_inv; if (abs_s > env) env = abs_s; return env; }
    if (mode == GDG_FOLLOW_LEVEL) { double diff = abs_s - env; return env + diff * d; }
    return 1.0;
}
GDG_DEVI double follow_A(int mode, double d_inv, double d)
{
    return mode == GDG_FOLLOW_ENVELOPE ? d_inv : (mode == GDG_FOLLOW_LEVEL ? 1.0 - d : 0.0);
}

/* --------------------------------------------------------------------------
 * Block scan of per-chunk recurrence maps held in registers.
 *   chunk t maps its carry c to   affine: a_t c + b_t     maxplus: max(a_t c, b_t)
 * Every thread passes the map of its own chunk (identity a=1, b=0 for an empty
 * chunk; all maxplus operands are >= 0) and receives the carry entering its chunk;
 * chunk 0 receives init.  K independent recurrences are scanned together.
 * Device: inclusive warp scan with shuffles, warp totals through shared memory,
 * serial prefix over the <= 15 preceding warps.  Host emulation: plain serial scan.
 * -------------------------------------------------------------------------- */
GDG_DEVI double scan_apply(int maxplus, double a, double c, double b)
{
    const double v = a * c;
    if (maxplus) return v > b ? v : b;
    return v + b;
}

#if defined(__CUDACC__) && !defined(GDG_HOST_EMU)
template <int K, int NT = GDG_NT>
GDG_DEVI void scan_regs(double *ws, int maxplus, double (&a)[K], double (&b)[K], const double *init, double (&cin)[K])
{
    constexpr int NSEG = NT / 32;
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
#pragma unroll
        for (int j = 0; j < K; j++) {
            const double au = __shfl_up_sync(0xffffffffu, a[j], off), bu = __shfl_up_sync(0xffffffffu, b[j], off);
            if (lane >= off) { b[j] = scan_apply(maxplus, a[j], bu, b[j]); a[j] *= au; }
        }
    }
    /* ws: [K][NSEG][2] warp totals in shared memory */
    if (lane == 31) {
#pragma unroll
        for (int j = 0; j < K; j++) { ws[(j * NSEG + warp) * 2] = a[j]; ws[(j * NSEG + warp) * 2 + 1] = b[j]; }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < K; j++) {
        double c = init[j];
        for (int w = 0; w < warp; w++) c = scan_apply(maxplus, ws[(j * NSEG + w) * 2], c, ws[(j * NSEG + w) * 2 + 1]);
        double ae = __shfl_up_sync(0xffffffffu, a[j], 1), be = __shfl_up_sync(0xffffffffu, b[j], 1);
        if (lane == 0) { ae = 1.0; be = 0.0; }
        cin[j] = scan_apply(maxplus, ae, c, be);
    }
    __syncthreads();
}
#define GDG_SCAN(K, cx, maxplus, a, b, init, cin) scan_regs<K>(SM_SEG, maxplus, a, b, init, cin)
/* the same scan in a CTA of NT threads with its warp totals at `ws` */
#define GDG_SCAN_N(K, NT, ws, maxplus, a, b, init, cin) scan_regs<K, NT>(ws, maxplus, a, b, init, cin)
#else
template <int K>
static void scan_regs_host(int nt, int maxplus, double (*a)[K], double (*b)[K], const double *init, double (*cin)[K])
{
    for (int j = 0; j < K; j++) {
        double c = init[j];
        for (int t = 0; t < nt; t++) { cin[t][j] = c; c = scan_apply(maxplus, a[t][j], c, b[t][j]); }
    }
}
#define GDG_SCAN(K, cx, maxplus, a, b, init, cin) scan_regs_host<K>(GDG_NT, maxplus, a##_r, b##_r, init, cin##_r)
#define GDG_SCAN_N(K, NT, ws, maxplus, a, b, init, cin) scan_regs_host<K>(NT, maxplus, a##_r, b##_r, init, cin##_r)
#endif

GDG_DEVI double chunk_A(double a_step, int len, int S, double a_full) { return len == S ? a_full : pow_int(a_step, len); }

/* History rings of the delay-type units.  The reference keeps the last `hs`
 * inputs oldest-first and memmoves them every block (e.g. chorus.go:119-129);
 * here they are rings: logical index q (0 = oldest) lives at (pos + q) % hs. */
GDG_DEVI double ring_get(const double *ring, int hs, int pos, int q)
{
    int p = pos + q;
    if (p >= hs) p -= hs;
    return GDG_LDST(ring + p);
}
/* x[idx] for idx in [-hs, n): block sample or history.  An index below -hs (flanger / phaser at 22.05 / 44.1 / 88.2 kHz,
 * depth 100: delay 0.002 fs against a history of round(0.002 fs) samples - the reference panics there, flanger.go:80-84)
 * reads the OLDEST history sample: defined behaviour shared with the oracle, never outside the ring. */
GDG_DEVI double hist_or_block(const double *blk, const double *ring, int hs, int pos, int idx)
{
    return idx >= 0 ? blk[sw(idx)] : ring_get(ring, hs, pos, idx < -hs ? 0 : hs + idx);
}
/* same, for units that read mostly from the ring (delay, reverb line): both loads are issued unconditionally on
 * clamped (valid) addresses so that the global loads of an unrolled loop are batched instead of branched around */
GDG_DEVI double hist_or_block_nb(const double *blk, const double *ring, int hs, int pos, int idx)
{
    int p = pos + hs + (idx < 0 ? idx : -1);
    if (p >= hs) p -= hs;
    if (p < 0) p = 0;
    const double hv = hs > 0 ? GDG_LDST(ring + p) : 0.0;
    const double bv = blk[sw(idx < 0 ? 0 : idx)];
    return idx >= 0 ? bv : hv;
}
/* push the block into the ring (one phase), returns nothing; new pos written by thread 0 */
GDG_DEV void ring_push(gdg_ctx &cx, const double *blk, double *st, int hs)
{
    const int n = cx.n;
    double *ring = st + GDG_RING_ST_DATA;
    GDG_THREADS_BEGIN
        const int pos = (int)st_geti(st, GDG_RING_ST_POS);
        if (hs > 0) {
            if (n >= hs) {
#pragma unroll 1
                for (int q = tid; q < hs; q += GDG_NT) ring[q] = blk[sw(n - hs + q)];
            } else {
#pragma unroll 1
                for (int k = tid; k < n; k += GDG_NT) { int p = pos + k; if (p >= hs) p -= hs; ring[p] = blk[sw(k)]; }
            }
        }
    GDG_THREADS_END
    GDG_THREADS_BEGIN
        if (tid == 0 && hs > 0) {
            const int pos = (int)st_geti(st, GDG_RING_ST_POS);
            st_seti(st, GDG_RING_ST_POS, n >= hs ? 0 : (pos + n) % hs);
        }
    GDG_THREADS_END
}

/* ==========================================================================
 * noise_gate  (effects/noisegate.go:19-96)
 *   i[0] bypass-like flag (threshold_open < threshold_close), i[1] holdSamples
 *   p[0] facOpen p[1] facClose        st: [0] gateOpen (i64) [1] onHoldSince (i64)
 * "since" at sample k = k - (index of the last sample above the close threshold)
 * -> max-scan; gate = type of the last open/close event -> last-event (max) scan.
 * ========================================================================== */
GDG_DEV void stage_noisegate(gdg_ctx &cx, const gdg_unit_desc &d_global)
{
    const gdg_unit_desc &d = GDG_DESC; (void)d_global;
    const int n = cx.n, S = chunk_len(n);
    double *x = SM_CUR(cx), *st = d.st;
    if (d.i[0]) {                                   /* levelOpen < levelClose: copy, force open */
        GDG_THREADS_BEGIN
            if (tid == 0) { st_seti(st, 0, 1); st_seti(st, 1, 0); }
        GDG_THREADS_END
        return;
    }
    const double facOpen = d.p[0], facClose = d.p[1];
    const unsigned long long hold = (unsigned long long)(unsigned int)d.i[1];
    const unsigned long long since0 = (unsigned long long)st_geti(st, 1);
    const int open0 = (int)st_geti(st, 0);
    const double zero[1] = {0.0};
    GDG_REGA(double, a, 1); GDG_REGA(double, b, 1); GDG_REGA(double, lcin, 1); GDG_REGA(double, evin, 1);
    /* last close-threshold crossing per chunk (as index+1, 0 = none) */
    GDG_THREADS_BEGIN
        const int k0 = tid * S, k1 = imin(n, k0 + S);
        double last = 0.0;
#pragma unroll 1
        for (int k = k0; k < k1; k++) if (fabs(x[sw(k)]) > facClose) last = (double)(k + 1);
        R(a)[0] = 1.0; R(b)[0] = last;
    GDG_THREADS_END_NOSYNC
    GDG_SCAN(1, cx, 1, a, b, zero, lcin);
    /* gate events per chunk: 1 open, 2 close (close wins on the same sample) -> code of the last event */
    GDG_THREADS_BEGIN
        const int k0 = tid * S, k1 = imin(n, k0 + S);
        long long lc = (long long)R(lcin)[0];
        double ev = 0.0;
#pragma unroll 1
        for (int k = k0; k < k1; k++) {
            const double amp = fabs(x[sw(k)]);
            if (amp > facClose) lc = k + 1;
            unsigned long long since = lc > 0 ? (unsigned long long)(k + 1 - lc) : since0 + (unsigned long long)k;
            if (since > 0xffffffffull) since = 0xffffffffull;
            if (amp > facOpen) ev = 1.0;
            if (since >= hold) ev = 2.0;
        }
        R(a)[0] = 1.0; R(b)[0] = ev > 0.0 ? (double)((tid + 1) * 4) + ev : 0.0;
    GDG_THREADS_END_NOSYNC
    GDG_SCAN(1, cx, 1, a, b, zero, evin);
    GDG_THREADS_BEGIN
        const int k0 = tid * S, k1 = imin(n, k0 + S);
        long long lc = (long long)R(lcin)[0];
        const long long code = (long long)R(evin)[0];
        int open = code > 0 ? ((code & 3) == 1) : open0;
        unsigned long long since = 0;
#pragma unroll 1
        for (int k = k0; k < k1; k++) {
            const double s = x[sw(k)];
            const double amp = fabs(s);
            if (amp > facClose) lc = k + 1;
            since = lc > 0 ? (unsigned long long)(k + 1 - lc) : since0 + (unsigned long long)k;
            if (since > 0xffffffffull) since = 0xffffffffull;
            if (amp > facOpen) open = 1;
            if (since >= hold) open = 0;
            x[sw(k)] = (open ? 1.0 : 0.0) * s;
        }
        if (k0 < n && k1 == n) {
            if (since < 0xffffffffull) since++;
            st_seti(st, 0, open); st_seti(st, 1, (long long)since);
        }
    GDG_THREADS_END
}

/* ==========================================================================
 * compressor  (effects/compressor.go:18-84)
 *   i[0] follow mode; p[0] dInv p[1] d p[2] targetFac p[3] limitFac; st[0] envelope
 * ========================================================================== */
GDG_DEV void stage_compressor(gdg_ctx &cx, const gdg_unit_desc &d_global)
{
    const gdg_unit_desc &d = GDG_DESC; (void)d_global;
    const int n = cx.n, S = chunk_len(n), mode = d.i[0];
    const double dinv = d.p[0], dd = d.p[1], target = d.p[2], limit = d.p[3];
    double *x = SM_CUR(cx), *st = d.st;
    const double astep = follow_A(mode, dinv, dd), afull = pow_int(astep, S);
    const double init[1] = {st[0]};
    GDG_REGA(double, a, 1); GDG_REGA(double, b, 1); GDG_REGA(double, cin, 1);
    GDG_THREADS_BEGIN
        const int k0 = tid * S, k1 = imin(n, k0 + S);
        double e = 0.0;
#pragma unroll 1
        for (int k = k0; k < k1; k++) e = follow_step(mode, e, fabs(x[sw(k)]), dinv, dd);
        R(a)[0] = k0 < n ? chunk_A(astep, k1 - k0, S, afull) : 1.0; R(b)[0] = k0 < n ? e : 0.0;
    GDG_THREADS_END_NOSYNC
    GDG_SCAN(1, cx, mode == GDG_FOLLOW_ENVELOPE, a, b, init, cin);
    GDG_THREADS_BEGIN
        const int k0 = tid * S, k1 = imin(n, k0 + S);
        double e = R(cin)[0];
#pragma unroll 1
        for (int k = k0; k < k1; k++) {
            const double s = x[sw(k)];
            e = follow_step(mode, e, fabs(s), dinv, dd);
            double gain = target / e;
            if (gain > limit) gain = limit;
            x[sw(k)] = clamp1(gain * s);
        }
        if (k0 < n && k1 == n) st[0] = e;
    GDG_THREADS_END
}

/* ==========================================================================
 * tone_stack  (effects/tonestack.go:19-100): 4 bands HP(f_j) -> LP(f_{j+1}),
 * LP output taken BEFORE its update, weighted sum, clamp.
 *   p[0..3] HP k, p[4..7] LP k, p[8..11] band factors; st[0..3] hcv, st[4..7] lcv
 * ========================================================================== */
GDG_DEV void stage_tonestack(gdg_ctx &cx, const gdg_unit_desc &d_global)
{
    const gdg_unit_desc &d = GDG_DESC; (void)d_global;
    const int n = cx.n, S = chunk_len(n);
    double *x = SM_CUR(cx), *st = d.st;
    double ahf[4], alf[4], inith[4], initl[4];
    for (int j = 0; j < 4; j++) { ahf[j] = pow_int(1.0 - d.p[j], S); alf[j] = pow_int(1.0 - d.p[4 + j], S); inith[j] = st[j]; initl[j] = st[4 + j]; }
    GDG_REGA(double, a, 4); GDG_REGA(double, b, 4); GDG_REGA(double, hin, 4); GDG_REGA(double, lin, 4);
    GDG_THREADS_BEGIN
        const int k0 = tid * S, k1 = imin(n, k0 + S);
        double h[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
        for (int k = k0; k < k1; k++) {
            const double s = x[sw(k)];
            for (int j = 0; j < 4; j++) h[j] += (s - h[j]) * d.p[j];
        }
        for (int j = 0; j < 4; j++) { R(a)[j] = k0 < n ? chunk_A(1.0 - d.p[j], k1 - k0, S, ahf[j]) : 1.0; R(b)[j] = k0 < n ? h[j] : 0.0; }
    GDG_THREADS_END_NOSYNC
    GDG_SCAN(4, cx, 0, a, b, inith, hin);
    GDG_THREADS_BEGIN
        const int k0 = tid * S, k1 = imin(n, k0 + S);
        double h[4], l[4] = {0.0, 0.0, 0.0, 0.0};
        for (int j = 0; j < 4; j++) h[j] = R(hin)[j];
#pragma unroll 1
        for (int k = k0; k < k1; k++) {
            const double s = x[sw(k)];
            for (int j = 0; j < 4; j++) {
                double diff = s - h[j];
                h[j] += diff * d.p[j];
                diff -= l[j];
                l[j] += diff * d.p[4 + j];
            }
        }
        for (int j = 0; j < 4; j++) { R(a)[j] = k0 < n ? chunk_A(1.0 - d.p[4 + j], k1 - k0, S, alf[j]) : 1.0; R(b)[j] = k0 < n ? l[j] : 0.0; }
    GDG_THREADS_END_NOSYNC
    GDG_SCAN(4, cx, 0, a, b, initl, lin);
    GDG_THREADS_BEGIN
        const int k0 = tid * S, k1 = imin(n, k0 + S);
        double h[4], l[4];
        for (int j = 0; j < 4; j++) { h[j] = R(hin)[j]; l[j] = R(lin)[j]; }
#pragma unroll 1
        for (int k = k0; k < k1; k++) {
            const double s = x[sw(k)];
            double sum = 0.0;
            for (int j = 0; j < 4; j++) {
                double diff = s - h[j];
                h[j] += diff * d.p[j];
                diff -= l[j];
                const double pre = l[j];
                l[j] += diff * d.p[4 + j];
                sum += d.p[8 + j] * pre;
            }
            x[sw(k)] = clamp1(sum);
        }
        if (k0 < n && k1 == n) for (int j = 0; j < 4; j++) { st[j] = h[j]; st[4 + j] = l[j]; }
    GDG_THREADS_END
}

/* ==========================================================================
 * cabinet  (effects/cabinet.go:27-162): 3 cascaded one-pole high-passes
 * (out = x - v_old) then 4 cascaded low-passes (out = v_old), clamp.
 *   p[0..6] k of the 7 sections; st[0..6] cap voltages
 * Every pass is chunk-local; the replay of section s also accumulates the
 * zero-state response of section s+1, so a section costs one scan + one pass.
 * ========================================================================== */
GDG_DEVI double onepole_out(int highpass, double xin, double &v, double k)
{
    const double diff = xin - v;
    const double out = highpass ? diff : v;
    v += diff * k;
    return out;
}

/* The same cascade for full chunks whose length is a multiple of 4 (every block of >= 4 GDG_NT samples): a thread walks its
 * chunk as FOUR sub-chunks in parallel - four independent recurrence chains instead of one four times as long - because these
 * replays are bound by the dependent-issue latency of the FP64 pipe (8.3 cycles per step, tools/fp64_lat.cu), not by its
 * rate.  The sub-chunk maps are folded into the thread's map for the block scan; the carry into sub-chunk j is the carry into
 * j-1 pushed through j-1's map. */
GDG_DEV void stage_cascade_sub4(gdg_ctx &cx, const gdg_unit_desc &d_global, int nsec, int nhp)
{
    const gdg_unit_desc &d = GDG_DESC; (void)d_global;
    const int n = cx.n, S = chunk_len(n), Ls = S / 4;
    double *x = SM_CUR(cx), *st = d.st;
    GDG_REGA(double, a, 1); GDG_REGA(double, b, 1); GDG_REGA(double, cin, 1); GDG_REGA(double, bs, 4);
    GDG_THREADS_BEGIN
        const int k0 = tid * S;
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        const double k1 = d.p[0];
#pragma unroll 1
        for (int i = 0; i < Ls; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) v[j] += (x[sw(k0 + j * Ls + i)] - v[j]) * k1;
#pragma unroll
        for (int j = 0; j < 4; j++) R(bs)[j] = v[j];
    GDG_THREADS_END_NOSYNC
    for (int s = 0; s < nsec; s++) {
        const double astep = 1.0 - d.p[s], asub = pow_int(astep, Ls), afull = pow_int(asub, 4);
        const double init[1] = {st[s]};
        GDG_THREADS_BEGIN
            double B = R(bs)[0];
#pragma unroll
            for (int j = 1; j < 4; j++) B = (asub * B) + R(bs)[j];
            R(a)[0] = afull; R(b)[0] = B;
        GDG_THREADS_END_NOSYNC
        GDG_SCAN(1, cx, 0, a, b, init, cin);
        const int last = (s == nsec - 1);
        GDG_THREADS_BEGIN
            const int k0 = tid * S;
            double v[4], w[4] = {0.0, 0.0, 0.0, 0.0};
            v[0] = R(cin)[0];
#pragma unroll
            for (int j = 1; j < 4; j++) v[j] = (asub * v[j - 1]) + R(bs)[j - 1];
            const double kk = d.p[s], kn = last ? 0.0 : d.p[s + 1];
            const int hp = s < nhp;
#pragma unroll 1
            for (int i = 0; i < Ls; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int at = sw(k0 + j * Ls + i);
                    double y = onepole_out(hp, x[at], v[j], kk);
                    if (last) y = clamp1(y); else w[j] += (y - w[j]) * kn;
                    x[at] = y;
                }
#pragma unroll
            for (int j = 0; j < 4; j++) R(bs)[j] = w[j];
            if (k0 + S == n) st[s] = v[3];
        GDG_THREADS_END_NOSYNC
    }
    GDG_THREADS_BEGIN
    GDG_THREADS_END
}

GDG_DEV void stage_cascade(gdg_ctx &cx, const gdg_unit_desc &d_global, int nsec, int nhp)
{
    if (cx.n % (4 * GDG_NT) == 0) { stage_cascade_sub4(cx, d_global, nsec, nhp); return; }
    const gdg_unit_desc &d = GDG_DESC; (void)d_global;
    const int n = cx.n, S = chunk_len(n);
    double *x = SM_CUR(cx), *st = d.st;
    GDG_REGA(double, a, 1); GDG_REGA(double, b, 1); GDG_REGA(double, cin, 1);
    GDG_THREADS_BEGIN
        const int k0 = tid * S, k1 = imin(n, k0 + S);
        double v = 0.0;
#pragma unroll 1
        for (int k = k0; k < k1; k++) v += (x[sw(k)] - v) * d.p[0];
        R(b)[0] = k0 < n ? v : 0.0;
    GDG_THREADS_END_NOSYNC
    for (int s = 0; s < nsec; s++) {
        const double astep = 1.0 - d.p[s], afull = pow_int(astep, S);
        const double init[1] = {st[s]};
        GDG_THREADS_BEGIN
            const int k0 = tid * S, k1 = imin(n, k0 + S);
            R(a)[0] = k0 < n ? chunk_A(astep, k1 - k0, S, afull) : 1.0;
        GDG_THREADS_END_NOSYNC
        GDG_SCAN(1, cx, 0, a, b, init, cin);
        const int last = (s == nsec - 1);
        GDG_THREADS_BEGIN
            const int k0 = tid * S, k1 = imin(n, k0 + S);
            double v = R(cin)[0], w = 0.0;
            const double kk = d.p[s], kn = last ? 0.0 : d.p[s + 1];
#pragma unroll 1
            for (int k = k0; k < k1; k++) {
                double y = onepole_out(s < nhp, x[sw(k)], v, kk);
                if (last) y = clamp1(y); else w += (y - w) * kn;
                x[sw(k)] = y;
            }
            R(b)[0] = k0 < n ? w : 0.0;
            if (k0 < n && k1 == n) st[s] = v;
        GDG_THREADS_END_NOSYNC
    }
    GDG_THREADS_BEGIN
    GDG_THREADS_END
}

/* ==========================================================================
 * bandpass  (effects/bandpass.go:20-98): order/2 cascaded HP+LP pairs, the LP
 * output (pre-update value) is clamped between pairs.
 *   i[0] halfOrder; p[0] HP k, p[1] LP k; st[0..3] hcv, st[4..7] lcv
 * ========================================================================== */
GDG_DEV void stage_bandpass(gdg_ctx &cx, const gdg_unit_desc &d_global)
{
    const gdg_unit_desc &d = GDG_DESC; (void)d_global;
    const int n = cx.n, S = chunk_len(n), half = d.i[0];
    const double ka = d.p[0], kb = d.p[1];
    const double aaf = pow_int(1.0 - ka, S), abf = pow_int(1.0 - kb, S);
    double *x = SM_CUR(cx), *st = d.st;
    GDG_REGA(double, a, 1); GDG_REGA(double, b, 1); GDG_REGA(double, hin, 1); GDG_REGA(double, lin, 1);
    for (int j = 0; j < half; j++) {
        const double inith[1] = {st[j]}, initl[1] = {st[4 + j]};
        GDG_THREADS_BEGIN
            const int k0 = tid * S, k1 = imin(n, k0 + S);
            double h = 0.0;
#pragma unroll 1
            for (int k = k0; k < k1; k++) h += (x[sw(k)] - h) * ka;
            R(a)[0] = k0 < n ? chunk_A(1.0 - ka, k1 - k0, S, aaf) : 1.0; R(b)[0] = k0 < n ? h : 0.0;
        GDG_THREADS_END_NOSYNC
        GDG_SCAN(1, cx, 0, a, b, inith, hin);
        GDG_THREADS_BEGIN
            const int k0 = tid * S, k1 = imin(n, k0 + S);
            double h = R(hin)[0], l = 0.0;
#pragma unroll 1
            for (int k = k0; k < k1; k++) {
                double diff = x[sw(k)] - h;
                h += diff * ka;
                diff -= l;
                l += diff * kb;
            }
            R(a)[0] = k0 < n ? chunk_A(1.0 - kb, k1 - k0, S, abf) : 1.0; R(b)[0] = k0 < n ? l : 0.0;
        GDG_THREADS_END_NOSYNC
        GDG_SCAN(1, cx, 0, a, b, initl, lin);
        GDG_THREADS_BEGIN
            const int k0 = tid * S, k1 = imin(n, k0 + S);
            double h = R(hin)[0], l = R(lin)[0];
#pragma unroll 1
            for (int k = k0; k < k1; k++) {
                double diff = x[sw(k)] - h;
                h += diff * ka;
                diff -= l;
                const double iv = l;
                l += diff * kb;
                x[sw(k)] = clamp1(iv);
            }
            if (k0 < n && k1 == n) { st[j] = h; st[4 + j] = l; }
        GDG_THREADS_END_NOSYNC
    }
    GDG_THREADS_BEGIN
    GDG_THREADS_END
}

/* ==========================================================================
 * pointwise closed-form units
 * ========================================================================== */

/* tremolo (effects/tremolo.go:15-65): two-state counter machine in closed form.
 *   i[0] samplesAttenuated i[1] samplesUnattenuated (uint32 bit patterns); p[0] fac
 *   st[0] attenuated (i64) st[1] inStateSince (i64) */
GDG_DEVI void tremolo_at(long long i, int att0, unsigned long long s0, unsigned long long lenA, unsigned long long lenU,
                         int &att, unsigned long long &since)
{
    const unsigned long long lenCur = att0 ? lenA : lenU;
    const unsigned long long r0 = s0 >= lenCur ? 0ull : lenCur - s0;
    if ((unsigned long long)i < r0) { att = att0; since = s0 + (unsigned long long)i; return; }
    const unsigned long long dA = lenA > 1 ? lenA : 1, dU = lenU > 1 ? lenU : 1;
    const unsigned long long j = (unsigned long long)i - r0;
    const unsigned long long m = j % (dA + dU);
    const unsigned long long dFirst = att0 ? dU : dA;   /* first state after the toggle is !att0 */
    if (m < dFirst) { att = !att0; since = m; }
    else { att = att0; since = m - dFirst; }
}
GDG_DEV void stage_tremolo(gdg_ctx &cx, const gdg_unit_desc &d_global)
{
    const gdg_unit_desc &d = GDG_DESC; (void)d_global;
    const int n = cx.n;
    double *x = SM_CUR(cx), *st = d.st;
    const unsigned long long lenA = (unsigned int)d.i[0], lenU = (unsigned int)d.i[1];
    const double fac = d.p[0];
    GDG_THREADS_BEGIN
        const int att0 = (int)st_geti(st, 0);
        const unsigned long long s0 = (unsigned long long)st_geti(st, 1) & 0xffffffffull;
#pragma unroll 1
        for (int k = tid; k < n; k += GDG_NT) {
            int att; unsigned long long since;
            tremolo_at(k, att0, s0, lenA, lenU, att, since);
            if (att) x[sw(k)] *= fac;
        }
    GDG_THREADS_END
    GDG_THREADS_BEGIN
        if (tid == 0 && n > 0) {
            const int att0 = (int)st_geti(st, 0);
            const unsigned long long s0 = (unsigned long long)st_geti(st, 1) & 0xffffffffull;
            int att; unsigned long long since;
            tremolo_at(n - 1, att0, s0, lenA, lenU, att, since);
            st_seti(st, 0, att); st_seti(st, 1, (long long)((since + 1) & 0xffffffffull));
        }
    GDG_THREADS_END
}

/* ring_modulator (effects/ringmodulator.go:18-47): p[0] phaseFraction; st[0] phase */
GDG_DEV void stage_ringmod(gdg_ctx &cx, const gdg_unit_desc &d_global)
{
    const gdg_unit_desc &d = GDG_DESC; (void)d_global;
    const int n = cx.n;
    double *x = SM_CUR(cx), *st = d.st;
    const double frac = d.p[0];
    GDG_THREADS_BEGIN
        const double phase = st[0];
#pragma unroll 1
        for (int k = tid; k < n; k += GDG_NT) {
            const double cp = fmod(GDG_MULADD_RN((double)k, frac, phase), GDG_TWO_PI);
            x[sw(k)] = sin(cp) * x[sw(k)];
        }
    GDG_THREADS_END
    GDG_THREADS_BEGIN
        if (tid == 0) st[0] = fmod(GDG_MULADD_RN((double)n, frac, st[0]), GDG_TWO_PI);
    GDG_THREADS_END
}

/* signal_generator (effects/signalgenerator.go:20-153)
 *   i[0] type (0 sine 1 triangle 2 square 3 sawtooth 4 noise); p[0] facInput p[1] facSignal p[2] phaseIncrement
 *   st[0] phase, st[1] LCG state x (random/random.go:23-55; a = 16807, m = 2^31 - 1) */
GDG_DEVI unsigned long long lcg_pow(unsigned long long e)
{
    const unsigned long long m = 2147483647ull;
    unsigned long long r = 1, b = 16807ull;
    while (e) { if (e & 1) r = (r * b) % m; b = (b * b) % m; e >>= 1; }
    return r;
}
GDG_DEV void stage_siggen(gdg_ctx &cx, const gdg_unit_desc &d_global)
{
    const gdg_unit_desc &d = GDG_DESC; (void)d_global;
    const int n = cx.n, S = chunk_len(n), type = d.i[0];
    double *x = SM_CUR(cx), *st = d.st;
    const double fin = d.p[0], fsig = d.p[1], inc = d.p[2];
    if (type == 4) {
        GDG_THREADS_BEGIN
            const int k0 = tid * S, k1 = imin(n, k0 + S);
            const unsigned long long m = 2147483647ull;
            if (k0 < n) {
                unsigned long long xs = ((unsigned long long)st_geti(st, 1) * lcg_pow((unsigned long long)k0)) % m;
#pragma unroll 1
                for (int k = k0; k < k1; k++) {
                    xs = (16807ull * xs) % m;
                    const double r = (double)xs / (double)(m - 1);
                    x[sw(k)] = (fin * x[sw(k)]) + (fsig * (1.0 - (2.0 * r)));
                }
                if (k1 == n) SM_MISC[0] = (double)xs;
            }
        GDG_THREADS_END
        GDG_THREADS_BEGIN
            if (tid == 0 && n > 0) st_seti(st, 1, (long long)SM_MISC[0]);
        GDG_THREADS_END
        return;
    }
    GDG_THREADS_BEGIN
        const double phase = st[0];
        const double two_over_pi = 2.0 / GDG_PI;
#pragma unroll 1
        for (int k = tid; k < n; k += GDG_NT) {
            const double cp = fmod(GDG_MULADD_RN((double)k, inc, phase), GDG_TWO_PI);
            double sig;
            if (type == 0) sig = sin(cp);
            else if (type == 1) sig = cp < GDG_PI ? (two_over_pi * cp) - 1.0 : 3.0 - (two_over_pi * cp);
            else if (type == 2) { const double v = GDG_PI - cp; sig = v < 0.0 ? -1.0 : (v > 0.0 ? 1.0 : 0.0); }
            else { sig = cp / GDG_PI; if (cp > GDG_PI) sig -= 2.0; }
            x[sw(k)] = (fin * x[sw(k)]) + (fsig * sig);
        }
    GDG_THREADS_END
    GDG_THREADS_BEGIN
        if (tid == 0) { const double ph = GDG_MULADD_RN((double)n, inc, st[0]); st[0] = fmod(ph, GDG_TWO_PI); }
    GDG_THREADS_END
}

/* ==========================================================================
 * delay-line gathers
 * ========================================================================== */
GDG_DEVI double frac_delay_read(const double *blk, const double *ring, int hs, int pos, int i, double ds)
{
    /* e.g. chorus.go:82-108: weights 1-(ds-floor) and 1-(ceil-ds) (sum to 2 when ds is integral) */
    const double early = floor(ds), late = ceil(ds);
    const double se = hist_or_block(blk, ring, hs, pos, i - (int)early);
    const double sl = hist_or_block(blk, ring, hs, pos, i - (int)late);
    const double we = 1.0 - (ds - early), wl = 1.0 - (late - ds);
    return (we * se) + (wl * sl);
}

/* the same read with both candidate loads of each tap issued unconditionally (hist_or_block_nb): for units whose taps fall
 * mostly into the history ring and never before it (chorus: 40 +- 10 ms against a 50 ms ring) - the global loads of the ten
 * taps of a sample are batched instead of branched around */
GDG_DEVI double frac_delay_read_nb(const double *blk, const double *ring, int hs, int pos, int i, double ds)
{
    const double early = floor(ds), late = ceil(ds);
    const double se = hist_or_block_nb(blk, ring, hs, pos, i - (int)early);
    const double sl = hist_or_block_nb(blk, ring, hs, pos, i - (int)late);
    const double we = 1.0 - (ds - early), wl = 1.0 - (late - ds);
    return (we * se) + (wl * sl);
}

/* delay (effects/delay.go:17-91): i[0] delaySamples (= ring size); p[0] feedbackFac p[1] levelFac */
GDG_DEV void stage_delay(gdg_ctx &cx, const gdg_unit_desc &d_global)
{
    const gdg_unit_desc &d = GDG_DESC; (void)d_global;
    const int n = cx.n, ds = d.i[0];
    const double fb = d.p[0], lvl = d.p[1];
    double *x = SM_CUR(cx), *y = SM_ALT(cx), *st = d.st;
    GDG_THREADS_BEGIN
        const int pos = (int)st_geti(st, GDG_RING_ST_POS);
        const double *ring = st + GDG_RING_ST_DATA;
        #pragma unroll 4
        for (int k = tid; k < n; k += GDG_NT) {
            const double dl = hist_or_block_nb(x, ring, ds, pos, k - ds);
            y[sw(k)] = clamp1(lvl * (x[sw(k)] + (fb * dl)));
        }
    GDG_THREADS_END
    ring_push(cx, x, st, ds);
    swap_bufs(cx);
}

/* chorus (effects/chorus.go:19-130): i[0] ring size; p[0] depth p[1] angularSpeed p[2] sampleRate p[3..6] cos(0.4 pi j), p[7..10] sin(0.4 pi j), j = 1..4
 * st[0] previousPhase (advanced by the RING length per call, chorus.go:114-117) */
GDG_DEV void stage_chorus(gdg_ctx &cx, const gdg_unit_desc &d_global)
{
    const gdg_unit_desc &d = GDG_DESC; (void)d_global;
    const int n = cx.n, hs = d.i[0];
    const double depth = d.p[0], ang = d.p[1], sr = d.p[2];
    double *x = SM_CUR(cx), *y = SM_ALT(cx), *st = d.st;
    GDG_THREADS_BEGIN
        const int pos = (int)st_geti(st, GDG_RING_ST_POS);
        const double *ring = st + GDG_RING_ST_DATA;
        const double prev = st[0];
        /* the five voices sit at fixed offsets 0.4*pi*j from the zero phase: one sincos per sample and the
         * angle-addition formula instead of five fmod + sin (differs from chorus.go:75-79 by ~1 ulp of the
         * sine, i.e. ~1e-13 samples of delay - far inside the 1e-9 budget) */
        double cj[5], sj[5];                           /* cos / sin of the voice offsets 0.4*pi*j (p[3..10], host-computed) */
        cj[0] = 1.0; sj[0] = 0.0;
        for (int j = 1; j < 5; j++) { cj[j] = d.p[2 + j]; sj[j] = d.p[6 + j]; }
#pragma unroll 2
        for (int k = tid; k < n; k += GDG_NT) {
            const double t = (double)k / sr;
            const double zero_phase = fmod(GDG_MULADD_RN(ang, t, prev), GDG_TWO_PI);
            double s0, c0;
            GDG_SINCOS(zero_phase, &s0, &c0);          /* one argument reduction for both */
            double eff = 0.0;
            for (int j = 0; j < 5; j++) {
                const double offset = depth * ((s0 * cj[j]) + (c0 * sj[j]));
                const double dsmp = (0.001 * (40.0 + offset)) * sr;
                eff += 0.2 * frac_delay_read_nb(x, ring, hs, pos, k, dsmp);
            }
            y[sw(k)] = (0.5 * x[sw(k)]) + (0.5 * eff);
        }
    GDG_THREADS_END
    GDG_THREADS_BEGIN
        if (tid == 0) st[0] = fmod(GDG_MULADD_RN(ang, (double)hs / sr, st[0]), GDG_TWO_PI);
    GDG_THREADS_END
    ring_push(cx, x, st, hs);
    swap_bufs(cx);
}

/* flanger / phaser (effects/flanger.go:19-117, phaser.go:19-125)
 *   i[0] ring size; p[0] depth p[1] angularSpeed p[2] sampleRate p[3] 1/sampleRate p[4] dry weight p[5] wet weight */
GDG_DEV void stage_flanger(gdg_ctx &cx, const gdg_unit_desc &d_global)
{
    const gdg_unit_desc &d = GDG_DESC; (void)d_global;
    const int n = cx.n, hs = d.i[0];
    const double depth = d.p[0], ang = d.p[1], sr = d.p[2], sr_inv = d.p[3], wa = d.p[4], wb = d.p[5];
    double *x = SM_CUR(cx), *y = SM_ALT(cx), *st = d.st;
    GDG_THREADS_BEGIN
        const int pos = (int)st_geti(st, GDG_RING_ST_POS);
        const double *ring = st + GDG_RING_ST_DATA;
        const double prev = st[0];
#pragma unroll 1
        for (int k = tid; k < n; k += GDG_NT) {
            const double t = (double)k * sr_inv;
            const double phase = fmod(GDG_MULADD_RN(ang, t, prev), GDG_TWO_PI);
            const double offset = depth * sin(phase);
            const double dsmp = (0.001 * (depth + offset)) * sr;
            y[sw(k)] = (wa * x[sw(k)]) + (wb * frac_delay_read(x, ring, hs, pos, k, dsmp));
        }
    GDG_THREADS_END
    GDG_THREADS_BEGIN
        if (tid == 0) st[0] = fmod(GDG_MULADD_RN(ang, (double)hs * sr_inv, st[0]), GDG_TWO_PI);
    GDG_THREADS_END
    ring_push(cx, x, st, hs);
    swap_bufs(cx);
}

/* auto_yoy (effects/autoyoy.go:19-156)
 *   i[0] ring size i[1] follow; p[0] dInv p[1] d p[2] levelA p[3] levelB p[4] depthA p[5] depthB p[6] slope p[7] sampleRate
 *   st[0] envelope */
GDG_DEV void stage_autoyoy(gdg_ctx &cx, const gdg_unit_desc &d_global)
{
    const gdg_unit_desc &d = GDG_DESC; (void)d_global;
    const int n = cx.n, S = chunk_len(n), hs = d.i[0], mode = d.i[1];
    const double dinv = d.p[0], dd = d.p[1];
    double *x = SM_CUR(cx), *y = SM_ALT(cx), *st = d.st;
    const double astep = follow_A(mode, dinv, dd), afull = pow_int(astep, S);
    const double init[1] = {st[0]};
    GDG_REGA(double, a, 1); GDG_REGA(double, b, 1); GDG_REGA(double, cin, 1);
    GDG_THREADS_BEGIN
        const int k0 = tid * S, k1 = imin(n, k0 + S);
        double e = 0.0;
#pragma unroll 1
        for (int k = k0; k < k1; k++) e = follow_step(mode, e, fabs(x[sw(k)]), dinv, dd);
        R(a)[0] = k0 < n ? chunk_A(astep, k1 - k0, S, afull) : 1.0; R(b)[0] = k0 < n ? e : 0.0;
    GDG_THREADS_END_NOSYNC
    GDG_SCAN(1, cx, mode == GDG_FOLLOW_ENVELOPE, a, b, init, cin);
    GDG_THREADS_BEGIN
        const int k0 = tid * S, k1 = imin(n, k0 + S);
        const int pos = (int)st_geti(st, GDG_RING_ST_POS);
        const double *ring = st + GDG_RING_ST_DATA;
        double e = R(cin)[0];
#pragma unroll 1
        for (int k = k0; k < k1; k++) {
            const double s = x[sw(k)];
            e = follow_step(mode, e, fabs(s), dinv, dd);
            const double level = 20.0 * log10(e);
            double dfac;
            if (level <= d.p[2]) dfac = d.p[4];
            else if (level >= d.p[3]) dfac = d.p[5];
            else dfac = d.p[4] + (d.p[6] * (level - d.p[2]));
            const double dsmp = (0.01 * dfac) * d.p[7];
            y[sw(k)] = (0.5 * s) + (0.5 * frac_delay_read(x, ring, hs, pos, k, dsmp));
        }
        if (k0 < n && k1 == n) st[0] = e;
    GDG_THREADS_END
    ring_push(cx, x, st, hs);
    swap_bufs(cx);
}

#endif /* GDG_STAGES_CUH */
