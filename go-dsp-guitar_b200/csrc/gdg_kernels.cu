/*
 * gdg_kernels.cu - CUDA backend (sm_100a) of libgdg: kernel wrappers around the
 * CTA-cooperative stages (gdg_interp.cuh) and the device plumbing the host
 * runtime needs (gdg_backend.h).  There is no CPU path in this file: if no CUDA
 * device is usable, gdgb_init() fails and with it every constructor of the ABI.
 */
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "gdg_backend.h"
#include "gdg_interp.cuh"
#include "gdg_os.cuh"

static thread_local std::string t_err;
/* One context per CUDA device the process uses (gdg_backend.h: gdgb_context_open / gdgb_select): its streams, events and
 * tables.  Context 0 is the device of gdgb_init(); a thread works on the context it selected last (default 0). */
struct DevCtx {
    int device = -1, sms = 0;
    cudaStream_t stream = nullptr, copy[2] = {nullptr, nullptr}, side[GDGB_SIDE_STREAMS] = {};
    cudaEvent_t events[GDGB_EVENTS];
    double *d_tw = nullptr;
    int pipe_fit[3] = {-1, -1, -1};                        /* resident clusters of k_os_pipe for P = 2, 4, 8 (queried on first use) */
};
static DevCtx g_ctx[GDGB_MAX_CONTEXTS];
static int g_nctx = 0;
static thread_local int t_cur = 0;
/* CUDA's current device belongs to the OS thread, the library's contexts to the process: a thread that never called
 * cudaSetDevice (a Go host's cgo calls arrive on whatever thread the scheduler picked, a UI thread polls the tuner) would
 * launch context 0's kernels with device 0 current - wrong whenever context 0 is another device.  t_dev remembers what this
 * thread was bound to; gdgb_bind() (every entry point of the host runtime passes through it) binds it to its context's device. */
static thread_local int t_dev = -1;
static std::vector<double> g_tab_aa2, g_tab_aa4, g_tab_lz2, g_tab_lz4, g_tab_tw;      /* host copies: uploaded to every context */
#define CX (g_ctx[t_cur])
#define g_device (CX.device)
#define g_sms (CX.sms)
#define g_stream (CX.stream)
#define g_copy_streams (CX.copy)
#define g_events (CX.events)
#define g_d_tw (CX.d_tw)
static std::atomic<uint64_t> g_launches{0};

static int cu_fail(const char *what, cudaError_t e)
{
    t_err = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return -1;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cu_fail(#call, e_); } while (0)

/* ---- kernels ------------------------------------------------------------------- */
/* source / destination rows of channel c in segment round `round` (see gdgb_launch_segments) */
struct gdg_io { const unsigned char *src; unsigned char *dst; double *side; int src_fmt, dst_fmt; };
__device__ __forceinline__ gdg_io segment_io(int c, int round, int nseg, const unsigned char *in, long in_stride, int in_fmt,
                                             unsigned char *out, long out_stride, int out_fmt, unsigned char *mid, long mid_stride,
                                             double *side, long side_stride)
{
    gdg_io io;
    const bool first = round == 0, last = round == nseg - 1;
    io.src = first ? in + (long)c * in_stride : mid + (long)c * mid_stride;
    io.src_fmt = first ? in_fmt : GDG_FMT_IEEE64;
    io.dst = last ? out + (long)c * out_stride : mid + (long)c * mid_stride;
    io.dst_fmt = last ? out_fmt : GDG_FMT_IEEE64;
    io.side = (last && side != nullptr && out_fmt != GDG_FMT_IEEE64) ? side + (long)c * side_stride : nullptr;   /* float64 copy of an encoded result */
    return io;
}
__global__ void __launch_bounds__(GDG_NT, GDG_MIN_CTAS)
k_chain(const gdg_chain_desc *const *descs, const int32_t *nseg, int round, const unsigned char *in, long in_stride, int in_fmt,
        unsigned char *out, long out_stride, int out_fmt, unsigned char *mid, long mid_stride, double *side, long side_stride, int n,
        const double *tw)
{
    const int c = blockIdx.x;
    const gdg_chain_desc *cd = descs[c];
    if (cd == nullptr) return;                             /* this channel's segment of this round runs in the other kernel (or it has none) */
    const gdg_io io = segment_io(c, round, nseg[c], in, in_stride, in_fmt, out, out_stride, out_fmt, mid, mid_stride, side, side_stride);
    run_chain(gdg_smem, cd, io.src, io.src_fmt, io.dst, io.dst_fmt, io.side, n, tw);
}
/* a run of consecutive oversampled units, tile by tile, GDG_OSK_CTAS CTAs per SM (gdg_os.cuh) */
__global__ void __launch_bounds__(GDG_OSK_NT, GDG_OSK_CTAS)
k_os(const gdg_chain_desc *const *descs, const int32_t *nseg, int round, const unsigned char *in, long in_stride, int in_fmt,
     unsigned char *out, long out_stride, int out_fmt, unsigned char *mid, long mid_stride, double *side, long side_stride, int n)
{
    const int c = blockIdx.x;
    const gdg_chain_desc *cd = descs[c];
    if (cd == nullptr) return;
    const gdg_io io = segment_io(c, round, nseg[c], in, in_stride, in_fmt, out, out_stride, out_fmt, mid, mid_stride, side, side_stride);
    run_os(gdg_smem, cd, io.src, io.src_fmt, io.dst, io.dst_fmt, io.side, n);
}
/* the same run pipelined over a thread-block cluster of P CTAs per channel (gdg_os.cuh: run_os_pipe); grid = channels * P */
__global__ void __launch_bounds__(GDG_OSK_NT, GDG_OSK_CTAS)
k_os_pipe(const gdg_chain_desc *const *descs, const int32_t *nseg, int round, const unsigned char *in, long in_stride, int in_fmt,
          unsigned char *out, long out_stride, int out_fmt, unsigned char *mid, long mid_stride, double *side, long side_stride, int n, int P)
{
    const int c = blockIdx.x / P;
    const gdg_chain_desc *cd = descs[c];
    if (cd == nullptr) return;                             /* the whole cluster leaves */
    const gdg_io io = segment_io(c, round, nseg[c], in, in_stride, in_fmt, out, out_stride, out_fmt, mid, mid_stride, side, side_stride);
    run_os_pipe(cd, io.src, io.src_fmt, io.dst, io.dst_fmt, io.side, n);
}
/* level meter: CTA c meters row c of d_in into state words d_state + 4*c */
__global__ void __launch_bounds__(GDG_NT) k_meter(const double *in, long stride, long n, double f, long long hold, double *state)
{
    __shared__ double red[GDG_METER_RED];
    meter_channel(red, in + (long)blockIdx.x * stride, n, f, hold, state + 4 * (long)blockIdx.x);
}
/* sample-format conversion of planar rows (wave.go codecs): row r of `in` (in_fmt) -> row r of `out` (out_fmt) */
__global__ void k_convert_rows(const unsigned char *in, long in_stride, int in_fmt, unsigned char *out, long out_stride, int out_fmt, int n)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const long r = blockIdx.y;
    if (k < n) encode_sample(out_fmt, out + r * out_stride, k, decode_sample(in_fmt, in + r * in_stride, k));
}
__global__ void k_resample_time(const double *in, long n, double dx, double *out, long out_len)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < out_len) out[i] = resample_time_body(in, n, dx, i);
}

__global__ void k_spat(const gdg_spat_channel *ch, int C, const double *in, long stride, const double *hist, int hs,
                       const double *aux, double *outL, double *outR, int n, int accumulate)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) spat_frame(ch, C, in, stride, hist, hs, aux, outL, outR, j, accumulate);
}
__global__ void k_spat_partial(const gdg_spat_channel *ch, int C, const double *in, long stride, const double *hist, int hs, double *part, int n)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) spat_partial(ch, C, in, stride, hist, hs, part, n, blockIdx.y, j);
}
__global__ void k_spat_combine(const double *part, int groups, const double *aux, double *outL, double *outR, int n, int accumulate)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) spat_combine(part, groups, aux, outL, outR, n, j, accumulate);
}
/* history <- last hs samples of (history ++ block): spatializer.go:316-331 */
__global__ void k_spat_push(int C, const double *in, long stride, const double *hist_in, double *hist_out, int hs, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (i < hs && c < C) {
        const long src = (long)i + n - hs;                 /* index into the block; negative -> old history */
        hist_out[(long)c * hs + i] = src >= 0 ? in[(long)c * stride + src] : hist_in[(long)c * hs + i + n];
    }
}

/* --- tuner autocorrelation: real FFT of size F through a complex FFT of size M = F/2 in global memory --- */
__global__ void k_ac_load(const double *ring, int n_ring, int pos, double *w, int F)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < F) ac_load_body(ring, n_ring, pos, w, i);
}
__global__ void k_ac_dif(double *w, int M, int h)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < M / 2) ac_dif_body(w, h, b);
}
__global__ void k_ac_dit(double *w, int M, int h)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < M / 2) ac_dit_body(w, h, b);
}
__global__ void k_ac_power(double *w, int M, int logM)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k <= M / 2) ac_power_body(w, M, logM, k);
}
__global__ void k_scale(double *w, int n, double f)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) w[i] *= f;
}

/* --- impulse-response compilation (gdg_interp.cuh ir_* bodies) --- */
__global__ void k_ir_pad(const double *x, long n_x, double *w, long n)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ir_pad_body(x, n_x, w, i);
}
__global__ void k_ir_split(const double *w, int M, int logM, double *spec)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k <= M) ir_split_body(w, M, logM, spec, k);
}
__global__ void k_ir_resample_bins(const double *in, int bins_in, double *out, int bins_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < bins_out) ir_resample_bins_body(in, bins_in, out, bins_out, i);
}
__global__ void k_ir_unsplit(const double *spec, int M, int logM, double *w)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < M) ir_unsplit_body(spec, M, logM, w, k);
}
__global__ void k_ir_scale_copy(const double *w, double f, double *out, long n)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = f * w[i];
}
/* one CTA, fixed summation tree: the same bits from run to run */
__global__ void __launch_bounds__(GDG_NT) k_ir_sum_squares(const double *x, long n, double *ss)
{
    __shared__ double red[GDG_NT];
    double s = 0.0;
    for (long i = threadIdx.x; i < n; i += GDG_NT) s += x[i] * x[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int h = GDG_NT / 2; h >= 1; h >>= 1) { if ((int)threadIdx.x < h) red[threadIdx.x] += red[threadIdx.x + h]; __syncthreads(); }
    if (threadIdx.x == 0) ss[0] = red[0];
}
__global__ void k_ir_scale_add(double *acc, const double *x, long n, double comp, double fac, const double *ss)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ir_scale_add_body(acc, x, comp, fac, ss, i);
}
__global__ void __launch_bounds__(GDG_NT, GDG_MIN_CTAS)
k_ir_partition_spectra(const double *coeffs, int L, int B, int logB, double *H, const double *tw)
{
    run_partition_spectrum(gdg_smem, coeffs, L, B, logB, blockIdx.x, H + 2L * (long)blockIdx.x * (B + 1), tw);
}
/* stereo master of a multi-device block: out[r][j] = sum over devices d (ascending) of part[d][r * stride_d + j] (+ aux).
 * The partial masters live on the OTHER devices: the loads are peer (NVLink) loads - the reduction and its transfer are
 * one kernel (spatializer.go:140-335 sums channels 0..C-1; here: the per-device partial sums in device order). */
struct gdg_master_parts { const double *part[GDGB_MAX_CONTEXTS]; long stride[GDGB_MAX_CONTEXTS]; int count; };
__global__ void k_master_sum(gdg_master_parts p, const double *aux, double *outL, double *outR, int n)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double l = 0.0, r = 0.0;
    for (int d = 0; d < p.count; d++) { l += p.part[d][j]; r += p.part[d][p.stride[d] + j]; }
    if (aux) { l += aux[j]; r += aux[j]; }
    outL[j] = l; outR[j] = r;
}
/* batched tuner autocorrelation: blockIdx.y selects the tuner */
struct gdg_ac_batch { const double *ring[32]; double *work[32]; int pos[32]; };
__global__ void k_acb_load(gdg_ac_batch b, int n_ring, int F)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < F) ac_load_body(b.ring[blockIdx.y], n_ring, b.pos[blockIdx.y], b.work[blockIdx.y], i);
}
__global__ void k_acb_power(gdg_ac_batch b, int M, int logM)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k <= M / 2) ac_power_body(b.work[blockIdx.y], M, logM, k);
}
__global__ void k_acb_scale(gdg_ac_batch b, int n, double f)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b.work[blockIdx.y][i] *= f;
}

/* ---- radix-2 passes fused through shared memory ----------------------------------------------------------------------
 * The 2^17-point transform of tuner.Analyze ran as one launch per radix-2 stage (17 + 17 launches, each a full round trip of
 * the work buffer through HBM).  The same butterflies - same twiddles, same order, bit-identical results - now run as two
 * launches per direction (the IR compiler's single, smaller transforms keep one launch per stage: spread over more SMs
 * they finish sooner than 16 fused CTAs - measured, tools/recompile_bench.py): `k_acb_lo` does the lowest (up to 11) stages of a
 * contiguous block of <= 2048 points in shared memory, `k_acb_hi` up to 6 of the stages above on a tile of 2^S rows x
 * 2048 / 2^S contiguous columns (rows `hl` points apart).  blockIdx.y selects the transform of a batch. */
#define GDG_ACF_N 2048
#define GDG_ACF_NT 512
__global__ void __launch_bounds__(GDG_ACF_NT) k_acb_lo(gdg_ac_batch b, int N, int inverse)
{
    __shared__ double sm[2 * GDG_ACF_N];
    double *w = b.work[blockIdx.y] + 2L * (long)blockIdx.x * N;
    const int tid = (int)threadIdx.x;
    for (int i = tid; i < 2 * N; i += GDG_ACF_NT) sm[i] = w[i];
    __syncthreads();
    if (!inverse) { for (int h = N / 2; h >= 1; h >>= 1) { for (int t = tid; t < N / 2; t += GDG_ACF_NT) ac_dif_body(sm, h, t); __syncthreads(); } }
    else { for (int h = 1; h <= N / 2; h <<= 1) { for (int t = tid; t < N / 2; t += GDG_ACF_NT) ac_dit_body(sm, h, t); __syncthreads(); } }
    for (int i = tid; i < 2 * N; i += GDG_ACF_NT) w[i] = sm[i];
}
/* stages with half-distances h0, h0/2, .., hl = h0 >> (S - 1) (forward: in this order, inverse: reversed); hl >= 2048 */
__global__ void __launch_bounds__(GDG_ACF_NT) k_acb_hi(gdg_ac_batch b, int h0, int S, int inverse)
{
    __shared__ double sm[2 * GDG_ACF_N];
    const int tid = (int)threadIdx.x;
    const int Rr = 1 << S, T = GDG_ACF_N >> S, logT = 11 - S, hl = h0 >> (S - 1);
    const int tiles_per_hi = hl / T, hi = (int)blockIdx.x / tiles_per_hi, mid = (int)blockIdx.x - hi * tiles_per_hi;
    double *w = b.work[blockIdx.y] + 2L * ((long)hi * 2 * h0 + (long)mid * T);
    for (int e = tid; e < GDG_ACF_N; e += GDG_ACF_NT) {
        const int r = e >> logT, c = e & (T - 1);
        const double2 v = *(const double2 *)(w + 2L * ((long)r * hl + c));
        sm[2 * e] = v.x; sm[2 * e + 1] = v.y;
    }
    __syncthreads();
    for (int st = 0; st < S; st++) {
        const int hr = inverse ? (1 << st) : (Rr >> (st + 1));                 /* row half-distance of this stage */
        const int h = hr * hl;
        for (int t = tid; t < GDG_ACF_N / 2; t += GDG_ACF_NT) {
            const int c = t & (T - 1), q = t >> logT;
            const int grp = q / hr, rp = q - grp * hr, r0 = grp * 2 * hr + rp;
            const int l0 = r0 * T + c, l1 = l0 + hr * T, pos = rp * hl + mid * T + c;
            if (!inverse) ac_dif_pair(sm, l0, l1, pos, h); else ac_dit_pair(sm, l0, l1, pos, h);
        }
        __syncthreads();
    }
    for (int e = tid; e < GDG_ACF_N; e += GDG_ACF_NT) {
        const int r = e >> logT, c = e & (T - 1);
        *(double2 *)(w + 2L * ((long)r * hl + c)) = make_double2(sm[2 * e], sm[2 * e + 1]);
    }
}
/* all radix-2 stages of `count` M-point transforms (work buffers in b), forward (natural -> bit-reversed) or inverse */
static void ac_fft_launch(const gdg_ac_batch &b, int count, int M, int inverse, cudaStream_t s)
{
    int logM = 0; while ((1 << logM) < M) logM++;
    const int lo_bits = logM < 11 ? logM : 11, N = 1 << lo_bits;
    /* chunks of the upper stages, from the largest half-distance down */
    int chunk_h0[8], chunk_S[8], nch = 0;
    for (int top = logM - 1; top >= lo_bits;) { const int S = (top - lo_bits + 1) < 6 ? (top - lo_bits + 1) : 6; chunk_h0[nch] = 1 << top; chunk_S[nch] = S; nch++; top -= S; }
    if (!inverse) {
        for (int k = 0; k < nch; k++) { k_acb_hi<<<dim3(M / GDG_ACF_N, count), GDG_ACF_NT, 0, s>>>(b, chunk_h0[k], chunk_S[k], 0); g_launches++; }
        if (lo_bits > 0) { k_acb_lo<<<dim3(M / N, count), GDG_ACF_NT, 0, s>>>(b, N, 0); g_launches++; }
    } else {
        if (lo_bits > 0) { k_acb_lo<<<dim3(M / N, count), GDG_ACF_NT, 0, s>>>(b, N, 1); g_launches++; }
        for (int k = nch - 1; k >= 0; k--) { k_acb_hi<<<dim3(M / GDG_ACF_N, count), GDG_ACF_NT, 0, s>>>(b, chunk_h0[k], chunk_S[k], 1); g_launches++; }
    }
}

/* ---- backend API ---------------------------------------------------------------- */
const char *gdgb_error(void) { return t_err.c_str(); }
const char *gdgb_name(void) { return "cuda:sm_100a"; }
int gdgb_device(void) { return g_device; }
int gdgb_sm_count(void) { return g_sms; }
void *gdgb_stream(void) { return (void *)g_stream; }
uint64_t gdgb_launch_count(void) { return g_launches.load(); }

static int open_context(int slot, int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) return cu_fail("cudaGetDeviceCount", e);
    if (count == 0) { t_err = "no CUDA device"; return -1; }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= count) { t_err = "device index out of range"; return -1; }
    CU(cudaSetDevice(device));
    t_dev = device;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) { t_err = std::string("device '") + prop.name + "' is not a Blackwell (sm_100a) GPU"; return -1; }
    DevCtx &c = g_ctx[slot];
    c.device = device; c.sms = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) CU(cudaStreamCreateWithFlags(&c.copy[i], cudaStreamNonBlocking));
    for (int i = 0; i < GDGB_SIDE_STREAMS; i++) CU(cudaStreamCreateWithFlags(&c.side[i], cudaStreamNonBlocking));
    for (int i = 0; i < GDGB_EVENTS; i++) CU(cudaEventCreateWithFlags(&c.events[i], cudaEventDisableTiming));
    CU(cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, GDG_SMEM_BYTES));
    CU(cudaFuncSetAttribute(k_ir_partition_spectra, cudaFuncAttributeMaxDynamicSharedMemorySize, GDG_SMEM_BYTES));
    CU(cudaFuncSetAttribute(k_os, cudaFuncAttributeMaxDynamicSharedMemorySize, GDG_OSK_SMEM_BYTES));
    CU(cudaFuncSetAttribute(k_os, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(k_os_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, GDG_OSK_SMEM_BYTES));
    CU(cudaFuncSetAttribute(k_os_pipe, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    return 0;
}
static int upload_tables()
{
    if (g_tab_tw.empty()) return 0;
    CU(cudaMemcpyToSymbol(c_aa2, g_tab_aa2.data(), sizeof(double) * 80));
    CU(cudaMemcpyToSymbol(c_aa4, g_tab_aa4.data(), sizeof(double) * 160));
    CU(cudaMemcpyToSymbol(c_lz2, g_tab_lz2.data(), sizeof(double) * 6));
    CU(cudaMemcpyToSymbol(c_lz4, g_tab_lz4.data(), sizeof(double) * 18));
    if (!g_d_tw) CU(cudaMalloc((void **)&g_d_tw, sizeof(double) * GDG_MAX_FFT));
    CU(cudaMemcpy(g_d_tw, g_tab_tw.data(), sizeof(double) * GDG_MAX_FFT, cudaMemcpyHostToDevice));
    return 0;
}
int gdgb_init(int device)
{
    t_cur = 0;
    if (open_context(0, device)) return -1;
    if (g_nctx < 1) g_nctx = 1;
    return 0;
}
/* a further device of this process: returns its context index (the existing one if the device is open already) or -1 */
int gdgb_context_open(int device)
{
    for (int i = 0; i < g_nctx; i++) if (g_ctx[i].device == device) return i;
    if (g_nctx >= GDGB_MAX_CONTEXTS) { t_err = "too many device contexts"; return -1; }
    const int prev = t_cur, slot = g_nctx;
    t_cur = slot;
    if (open_context(slot, device) || upload_tables()) { t_cur = prev; t_dev = -1; gdgb_bind(); return -1; }
    g_nctx++;
    t_cur = prev;
    t_dev = -1; gdgb_bind();
    return slot;
}
int gdgb_context_count(void) { return g_nctx; }
int gdgb_context_device(int ctx) { return (ctx >= 0 && ctx < g_nctx) ? g_ctx[ctx].device : -1; }
int gdgb_current(void) { return t_cur; }
int gdgb_bind(void)
{
    const int d = CX.device;
    if (d >= 0 && t_dev != d) { CU(cudaSetDevice(d)); t_dev = d; }
    return 0;
}
int gdgb_select(int ctx)
{
    if (ctx < 0 || ctx >= g_nctx) { t_err = "no such device context"; return -1; }
    t_cur = ctx;
    CU(cudaSetDevice(g_ctx[ctx].device));
    t_dev = g_ctx[ctx].device;
    return 0;
}
int gdgb_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }
/* peer access from context `ctx` to every other open context (NVLink / NVSwitch P2P): kernels on `ctx` may then dereference
 * the other devices' buffers */
int gdgb_enable_peers(int ctx)
{
    const int prev = t_cur;
    if (gdgb_select(ctx)) return -1;
    int rc = 0;
    for (int i = 0; i < g_nctx && !rc; i++) {
        if (i == ctx) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, g_ctx[ctx].device, g_ctx[i].device);
        if (!can) { t_err = "peer access between the devices is not available"; rc = -1; break; }
        cudaError_t e = cudaDeviceEnablePeerAccess(g_ctx[i].device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) rc = cu_fail("cudaDeviceEnablePeerAccess", e);
    }
    gdgb_select(prev);
    return rc;
}
/* stream 0 of context `ctx` waits for event `event` of context `other` (events order work across devices) */
int gdgb_wait_other(int ctx, int other, int event)
{
    if (ctx < 0 || ctx >= g_nctx || other < 0 || other >= g_nctx) { t_err = "no such device context"; return -1; }
    CU(cudaStreamWaitEvent(g_ctx[ctx].stream, g_ctx[other].events[event % GDGB_EVENTS], 0));
    return 0;
}

void *gdgb_alloc(size_t bytes)
{
    void *p = nullptr;
    if (bytes == 0) bytes = 8;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { cu_fail("cudaMalloc", e); return nullptr; }
    e = cudaMemsetAsync(p, 0, bytes, g_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);   /* the buffer may next be touched from another stream (uploads
                                                                  * run on their own): the zero fill must have landed by then */
    if (e != cudaSuccess) { cu_fail("cudaMemsetAsync", e); cudaFree(p); return nullptr; }
    return p;
}
void gdgb_free(void *p) { if (p) cudaFree(p); }
int gdgb_zero(void *p, size_t bytes) { CU(cudaMemsetAsync(p, 0, bytes, g_stream)); return 0; }
int gdgb_upload(void *dst, const void *src, size_t bytes) { if (bytes) CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_stream)); return 0; }
int gdgb_download(void *dst, const void *src, size_t bytes)
{
    if (bytes) CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    return 0;
}
int gdgb_copy(void *dst, const void *src, size_t bytes) { if (bytes) CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, g_stream)); return 0; }
int gdgb_copy_on(void *dst, const void *src, size_t bytes, void *stream) { if (bytes) CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream)); return 0; }
int gdgb_sync(void) { if (g_stream) CU(cudaStreamSynchronize(g_stream)); return 0; }
static cudaStream_t stream_of(int s) { return s == 0 ? g_stream : g_copy_streams[(s - 1) & 1]; }
void *gdgb_stream_handle(int s) { return (void *)stream_of(s); }
void *gdgb_side_stream(int i) { return (void *)CX.side[i % GDGB_SIDE_STREAMS]; }
int gdgb_h2d_async(void *dst, const void *src, size_t bytes, int s) { if (bytes) CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream_of(s))); return 0; }
int gdgb_d2h_async(void *dst, const void *src, size_t bytes, int s) { if (bytes) CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream_of(s))); return 0; }
int gdgb_record(int ev, int s) { CU(cudaEventRecord(g_events[ev % GDGB_EVENTS], stream_of(s))); return 0; }
int gdgb_wait(int s, int ev) { CU(cudaStreamWaitEvent(stream_of(s), g_events[ev % GDGB_EVENTS], 0)); return 0; }
/* everything enqueued on `then` after this call runs after what `first` holds now (event id: one of the pool) */
int gdgb_order(void *first, void *then, int ev)
{
    CU(cudaEventRecord(g_events[ev % GDGB_EVENTS], (cudaStream_t)first));
    CU(cudaStreamWaitEvent((cudaStream_t)then, g_events[ev % GDGB_EVENTS], 0));
    return 0;
}
int gdgb_sync_stream(int s) { CU(cudaStreamSynchronize(stream_of(s))); return 0; }
int gdgb_sync_event(int ev) { CU(cudaEventSynchronize(g_events[ev % GDGB_EVENTS])); return 0; }
void *gdgb_pinned_alloc(size_t bytes) { void *p = nullptr; if (cudaMallocHost(&p, bytes ? bytes : 8) != cudaSuccess) return nullptr; return p; }
void gdgb_pinned_free(void *p) { if (p) cudaFreeHost(p); }
int gdgb_is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return 0; }
    return a.type == cudaMemoryTypeHost;
}

int gdgb_set_tables(const double *aa2, const double *aa4, const double *lz2, const double *lz4, const double *tw)
{
    g_tab_aa2.assign(aa2, aa2 + 80); g_tab_aa4.assign(aa4, aa4 + 160); g_tab_lz2.assign(lz2, lz2 + 6); g_tab_lz4.assign(lz4, lz4 + 18);
    g_tab_tw.assign(tw, tw + GDG_MAX_FFT);
    return upload_tables();
}

int gdgb_launch_segments(int kind, const gdg_chain_desc *const *d_descs, const int32_t *d_nseg, int round, int channels,
                         const void *d_in, long in_stride_bytes, int in_fmt, void *d_out, long out_stride_bytes, int out_fmt,
                         void *d_mid, long mid_stride_bytes, double *d_side, long side_stride, int n, int max_units, void *stream)
{
    /* few channels and runs of several units: one cluster of P CTAs per channel, the units of a run pipelined over the tiles
     * (all clusters resident at once) - otherwise one CTA per channel */
    static const bool no_pipe = getenv("GDG_NO_OS_PIPE") != nullptr;        /* development knob */
    const int P = max_units <= 2 ? 2 : (max_units <= 4 ? 4 : 8);
    if (kind == 1 && !no_pipe && max_units >= 2 && max_units <= GDG_OSK_MAX_UNITS && (long)channels * P <= (long)GDG_OSK_CTAS * g_sms) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(channels * P)); cfg.blockDim = dim3(GDG_OSK_NT); cfg.dynamicSmemBytes = GDG_OSK_SMEM_BYTES; cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)P; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        /* only when every cluster is resident at once (clusters live inside one GPC: fewer fit than CTAs / P) - a second
         * wave of clusters would double the latency this shape exists to cut */
        int &fit = CX.pipe_fit[P == 2 ? 0 : (P == 4 ? 1 : 2)];
        if (fit < 0) { int nc = 0; fit = cudaOccupancyMaxActiveClusters(&nc, k_os_pipe, &cfg) == cudaSuccess ? nc : 0; cudaGetLastError(); }
        if (channels <= fit) {
        CU(cudaLaunchKernelEx(&cfg, k_os_pipe, d_descs, d_nseg, round, (const unsigned char *)d_in, in_stride_bytes, in_fmt,
                              (unsigned char *)d_out, out_stride_bytes, out_fmt, (unsigned char *)d_mid, mid_stride_bytes, d_side, side_stride, n, P));
        g_launches++;
        return 0;
        }
    }
    if (kind == 0)
        k_chain<<<channels, GDG_NT, GDG_SMEM_BYTES, (cudaStream_t)stream>>>(d_descs, d_nseg, round, (const unsigned char *)d_in, in_stride_bytes, in_fmt,
                                                                          (unsigned char *)d_out, out_stride_bytes, out_fmt, (unsigned char *)d_mid, mid_stride_bytes, d_side, side_stride, n, g_d_tw);
    else
        k_os<<<channels, GDG_OSK_NT, GDG_OSK_SMEM_BYTES, (cudaStream_t)stream>>>(d_descs, d_nseg, round, (const unsigned char *)d_in, in_stride_bytes, in_fmt,
                                                                               (unsigned char *)d_out, out_stride_bytes, out_fmt, (unsigned char *)d_mid, mid_stride_bytes, d_side, side_stride, n);
    g_launches++;
    CU(cudaGetLastError());
    return 0;
}
int gdgb_launch_meter(const double *d_in, long stride, int channels, long n, double f, long long hold, double *d_state, void *stream)
{
    if (channels > 0 && n > 0) {
        k_meter<<<channels, GDG_NT, 0, (cudaStream_t)stream>>>(d_in, stride, n, f, hold, d_state);
        g_launches++;
        CU(cudaGetLastError());
    }
    return 0;
}
int gdgb_convert_rows(const void *d_in, long in_stride_bytes, int in_fmt, void *d_out, long out_stride_bytes, int out_fmt, int rows, int n, void *stream)
{
    if (rows > 0 && n > 0) {
        dim3 grid((unsigned)((n + 255) / 256), (unsigned)rows);
        k_convert_rows<<<grid, 256, 0, (cudaStream_t)stream>>>((const unsigned char *)d_in, in_stride_bytes, in_fmt, (unsigned char *)d_out, out_stride_bytes, out_fmt, n);
        g_launches++;
        CU(cudaGetLastError());
    }
    return 0;
}
int gdgb_resample_time(const double *d_in, long n, double dx, double *d_out, long out_len, void *stream)
{
    if (out_len > 0) {
        k_resample_time<<<(unsigned)((out_len + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_in, n, dx, d_out, out_len);
        g_launches++;
        CU(cudaGetLastError());
    }
    return 0;
}

int gdgb_launch_spatializer(const gdg_spat_channel *d_ch, int channels, const double *d_in, long stride,
                            const double *d_hist_in, double *d_hist_out, int hs, const double *d_aux,
                            double *d_outL, double *d_outR, int n, int accumulate, double *d_partials, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (channels <= GDG_SPAT_GROUP || !d_partials) {
        k_spat<<<(n + 255) / 256, 256, 0, s>>>(d_ch, channels, d_in, stride, d_hist_in, hs, d_aux, d_outL, d_outR, n, accumulate);
        g_launches++;
    } else {
        const int groups = (channels + GDG_SPAT_GROUP - 1) / GDG_SPAT_GROUP;
        dim3 grid((n + 127) / 128, groups);
        k_spat_partial<<<grid, 128, 0, s>>>(d_ch, channels, d_in, stride, d_hist_in, hs, d_partials, n);
        k_spat_combine<<<(n + 127) / 128, 128, 0, s>>>(d_partials, groups, d_aux, d_outL, d_outR, n, accumulate);
        g_launches += 2;
    }
    CU(cudaGetLastError());
    if (channels > 0 && hs > 0) {
        dim3 grid((hs + 127) / 128, channels);
        k_spat_push<<<grid, 128, 0, s>>>(channels, d_in, stride, d_hist_in, d_hist_out, hs, n);
        g_launches++;
        CU(cudaGetLastError());
    }
    return 0;
}

int gdgb_autocorrelate(const double *d_ring, int n_ring, int pos, double *d_work, int F, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    const int M = F / 2;
    int logM = 0;
    while ((1 << logM) < M) logM++;
    k_ac_load<<<(F + 255) / 256, 256, 0, s>>>(d_ring, n_ring, pos, d_work, F); g_launches++;
    for (int h = M / 2; h >= 1; h >>= 1) { k_ac_dif<<<(M / 2 + 255) / 256, 256, 0, s>>>(d_work, M, h); g_launches++; }
    k_ac_power<<<(M / 2 + 1 + 255) / 256, 256, 0, s>>>(d_work, M, logM); g_launches++;
    for (int h = 1; h <= M / 2; h <<= 1) { k_ac_dit<<<(M / 2 + 255) / 256, 256, 0, s>>>(d_work, M, h); g_launches++; }
    k_scale<<<(F + 255) / 256, 256, 0, s>>>(d_work, F, 1.0 / (double)M); g_launches++;
    CU(cudaGetLastError());
    return 0;
}

/* ---- impulse-response compilation ------------------------------------------------------ */
static int ilog2(int v) { int e = 0; while ((1 << e) < v) e++; return e; }
int gdgb_ir_real_spectrum(const double *d_x, long n_x, int nfft, double *d_work, double *d_spec, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    const int M = nfft / 2, logM = ilog2(M);
    k_ir_pad<<<(unsigned)((nfft + 255) / 256), 256, 0, s>>>(d_x, n_x, d_work, nfft); g_launches++;
    for (int h = M / 2; h >= 1; h >>= 1) { k_ac_dif<<<(M / 2 + 255) / 256, 256, 0, s>>>(d_work, M, h); g_launches++; }
    k_ir_split<<<(M + 1 + 255) / 256, 256, 0, s>>>(d_work, M, logM, d_spec); g_launches++;
    CU(cudaGetLastError());
    return 0;
}
int gdgb_ir_resample_bins(const double *d_spec_in, int bins_in, double *d_spec_out, int bins_out, void *stream)
{
    k_ir_resample_bins<<<(bins_out + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_spec_in, bins_in, d_spec_out, bins_out); g_launches++;
    CU(cudaGetLastError());
    return 0;
}
int gdgb_ir_real_inverse(const double *d_spec, int nfft, double *d_work, double *d_out, long n_keep, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    const int M = nfft / 2, logM = ilog2(M);
    k_ir_unsplit<<<(M + 255) / 256, 256, 0, s>>>(d_spec, M, logM, d_work); g_launches++;
    for (int h = 1; h <= M / 2; h <<= 1) { k_ac_dit<<<(M / 2 + 255) / 256, 256, 0, s>>>(d_work, M, h); g_launches++; }
    k_ir_scale_copy<<<(unsigned)((n_keep + 255) / 256), 256, 0, s>>>(d_work, 1.0 / (double)M, d_out, n_keep); g_launches++;
    CU(cudaGetLastError());
    return 0;
}
int gdgb_ir_sum_squares(const double *d_x, long n, double *d_ss, void *stream)
{
    k_ir_sum_squares<<<1, GDG_NT, 0, (cudaStream_t)stream>>>(d_x, n, d_ss); g_launches++;
    CU(cudaGetLastError());
    return 0;
}
int gdgb_ir_scale_add(double *d_acc, const double *d_x, long n, double comp, double fac, const double *d_ss, void *stream)
{
    if (n > 0) { k_ir_scale_add<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_acc, d_x, n, comp, fac, d_ss); g_launches++; }
    CU(cudaGetLastError());
    return 0;
}
int gdgb_ir_partition_spectra(const double *d_coeffs, int L, int B, int logB, int K, double *d_H, void *stream)
{
    k_ir_partition_spectra<<<K, GDG_NT, GDG_SMEM_BYTES, (cudaStream_t)stream>>>(d_coeffs, L, B, logB, d_H, g_d_tw); g_launches++;
    CU(cudaGetLastError());
    return 0;
}
int gdgb_autocorrelate_batch(const double *const *d_rings, int n_ring, const int *pos, double *const *d_works, int count, int F, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    const int M = F / 2, logM = ilog2(M);
    for (int c0 = 0; c0 < count; c0 += 32) {
        const int cnt = count - c0 < 32 ? count - c0 : 32;
        gdg_ac_batch b;
        for (int i = 0; i < cnt; i++) { b.ring[i] = d_rings[c0 + i]; b.work[i] = d_works[c0 + i]; b.pos[i] = pos[c0 + i]; }
        dim3 gF((F + 255) / 256, cnt), gP((M / 2 + 1 + 255) / 256, cnt);
        k_acb_load<<<gF, 256, 0, s>>>(b, n_ring, F); g_launches++;
        ac_fft_launch(b, cnt, M, 0, s);
        k_acb_power<<<gP, 256, 0, s>>>(b, M, logM); g_launches++;
        ac_fft_launch(b, cnt, M, 1, s);
        k_acb_scale<<<gF, 256, 0, s>>>(b, F, 1.0 / (double)M); g_launches++;
    }
    CU(cudaGetLastError());
    return 0;
}

int gdgb_master_sum(const double *const *d_parts, const long *strides, int count, const double *d_aux, double *d_outL, double *d_outR, int n, void *stream)
{
    if (count > GDGB_MAX_CONTEXTS) { t_err = "too many partial masters"; return -1; }
    gdg_master_parts p; p.count = count;
    for (int d = 0; d < count; d++) { p.part[d] = d_parts[d]; p.stride[d] = strides[d]; }
    k_master_sum<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p, d_aux, d_outL, d_outR, n); g_launches++;
    CU(cudaGetLastError());
    return 0;
}
