/*
 * gdg_backend.h - what the host runtime (gdg_host.cpp) needs from the device
 * side.  The product implements it with CUDA (gdg_kernels.cu); tests/emu/ builds
 * a host-only implementation of the same kernels' source for CPU-side checks of
 * the parallel formulation (never shipped, never a fallback: the CUDA backend
 * fails loudly when no device is present).
 */
#ifndef GDG_BACKEND_H
#define GDG_BACKEND_H

#include <stddef.h>
#include <stdint.h>
#include "gdg_device.h"

/* all calls return 0 on success; on failure the message is in gdgb_error() */
const char *gdgb_error(void);
const char *gdgb_name(void);                    /* "cuda:sm_100a" or "host-emulation (tests only)" */
int gdgb_init(int device);                      /* context 0: selects the device, uploads constant tables */
/* further devices of ONE process (the reference host is one process with a goroutine per channel, controller.go:3339-3341):
 * a context = device + streams + events + tables; every call works on the context the calling thread selected last */
#define GDGB_MAX_CONTEXTS 16
int gdgb_device_count(void);
int gdgb_context_open(int device);              /* -> context index (existing one if already open) or -1 */
int gdgb_context_count(void);
int gdgb_context_device(int ctx);
int gdgb_current(void);
int gdgb_select(int ctx);
int gdgb_bind(void);                            /* make the calling OS thread's CUDA device the one of its current context */
int gdgb_enable_peers(int ctx);                 /* peer (NVLink) access from ctx to every other open context */
int gdgb_wait_other(int ctx, int other, int event);   /* stream 0 of ctx waits for an event recorded on context `other` */
/* out[r][j] = sum_d parts[d][r * strides[d] + j] (+ aux): the partial stereo masters of the other devices are read through peer addresses */
int gdgb_master_sum(const double *const *d_parts, const long *strides, int count, const double *d_aux, double *d_outL, double *d_outR, int n, void *stream);
int gdgb_device(void);
int gdgb_sm_count(void);
void *gdgb_stream(void);                        /* the library's stream (cudaStream_t) */

void *gdgb_alloc(size_t bytes);                 /* zero-initialised device memory, NULL on failure */
void gdgb_free(void *p);
int gdgb_zero(void *p, size_t bytes);
int gdgb_upload(void *dst, const void *src, size_t bytes);      /* host -> device, stream ordered */
int gdgb_download(void *dst, const void *src, size_t bytes);    /* device -> host, synchronises */
int gdgb_copy(void *dst, const void *src, size_t bytes);        /* device -> device, library stream */
int gdgb_copy_on(void *dst, const void *src, size_t bytes, void *stream);
int gdgb_sync(void);
/* copy/compute overlap for the host-buffer entry points: stream 0 = compute (== gdgb_stream()), 1 = H2D, 2 = D2H;
 * a small pool of events orders them.  Copies are truly asynchronous only for page-locked host memory. */
#define GDGB_EVENTS 64
void *gdgb_stream_handle(int stream);
#define GDGB_SIDE_STREAMS 3
void *gdgb_side_stream(int i);                         /* kernel streams beside the caller's: channel ranges of one batch run concurrently */
int gdgb_h2d_async(void *dst, const void *src, size_t bytes, int stream);
int gdgb_d2h_async(void *dst, const void *src, size_t bytes, int stream);
int gdgb_record(int event, int stream);
int gdgb_wait(int stream, int event);
int gdgb_order(void *first, void *then, int event);   /* `then` continues after what `first` holds now */
int gdgb_sync_stream(int stream);
int gdgb_sync_event(int event);                 /* host waits for the last record of `event` */
void *gdgb_pinned_alloc(size_t bytes);
void gdgb_pinned_free(void *p);
int gdgb_is_pinned(const void *p);

/* constant tables: anti-aliasing FIRs, Lanczos phase weights; twiddles live in device memory */
int gdgb_set_tables(const double *aa2, const double *aa4, const double *lz2, const double *lz4,
                    const double *twiddles /* GDG_MAX_FFT/2 complex */);

/* One segment round of a batch: CTA c runs d_descs[c] (NULL -> nothing to do for this channel in this launch).
 * kind 0: block-resident chain kernel (any units), kind 1: oversampled-run kernel (gdg_os.cuh).
 * Channel c reads d_in (in_fmt, GDG_FMT_*) when round == 0 and the float64 rows d_mid otherwise; it writes d_out (out_fmt)
 * when round == d_nseg[c] - 1 (its last segment) and d_mid otherwise.  Row c of a buffer starts at base + c * stride.
 * d_side (may be NULL; row stride in doubles): when out_fmt is an encoded format, the last segment also stores the
 * float64 result there (what the spatializer and the level meter read, controller.go:2746-2770).
 * max_units (kind 1): the longest run among the channels of this launch - with few channels the backend spreads a run over
 * a cluster of CTAs, one unit each (gdg_os.cuh: run_os_pipe). */
int gdgb_launch_segments(int kind, const gdg_chain_desc *const *d_descs, const int32_t *d_nseg, int round, int channels,
                         const void *d_in, long in_stride_bytes, int in_fmt, void *d_out, long out_stride_bytes, int out_fmt,
                         void *d_mid, long mid_stride_bytes, double *d_side, long side_stride, int n, int max_units, void *stream);
/* level meter (level/level.go:147-208): row c = d_in + c*stride (n samples) updates the 4-word state d_state + 4*c
 * {cur, peak, counter (int64), unused}; f = decay factor per sample, hold = peak hold time in samples */
int gdgb_launch_meter(const double *d_in, long stride, int channels, long n, double f, long long hold, double *d_state, void *stream);
/* planar rows from one sample format (GDG_FMT_*) to another: the wave.go codecs outside the chain kernel */
int gdgb_convert_rows(const void *d_in, long in_stride_bytes, int in_fmt, void *d_out, long out_stride_bytes, int out_fmt, int rows, int n, void *stream);
/* resample.Time on device: d_out[i], i < out_len */
int gdgb_resample_time(const double *d_in, long n, double dx, double *d_out, long out_len, void *stream);

/* spatializer: out[2][n] (+)= sum over channels; hist_in/hist_out: [C][hs] double-buffered history */
int gdgb_launch_spatializer(const gdg_spat_channel *d_ch, int channels, const double *d_in, long stride,
                            const double *d_hist_in, double *d_hist_out, int hs, const double *d_aux,
                            double *d_outL, double *d_outR, int n, int accumulate, double *d_partials, void *stream);
/* d_partials: 2 * ceil(channels / 32) * n doubles of scratch, needed when channels > 32 */

/* tuner: autocorrelation of `ring` (n_ring samples, oldest at `pos`) zero-padded to fft_size (power of two):
 * d_work: 2*fft_size doubles; result corr[0..fft_size) left in d_work */
int gdgb_autocorrelate(const double *d_ring, int n_ring, int pos, double *d_work, int fft_size, void *stream);

/* ---- impulse-response compilation on the device (bodies in gdg_interp.cuh): everything is
 * stream-ordered, nothing waits for the device ---- */
/* bins 0..nfft/2 (interleaved re, im) of the real FFT of x zero-padded to nfft (power of two >= 2); d_work: nfft doubles */
int gdgb_ir_real_spectrum(const double *d_x, long n_x, int nfft, double *d_work, double *d_spec, void *stream);
/* resample.Frequency (resample.go:109-142) */
int gdgb_ir_resample_bins(const double *d_spec_in, int bins_in, double *d_spec_out, int bins_out, void *stream);
/* RealInverseFourier of the Hermitian half d_spec (bins 0..nfft/2), first n_keep samples to d_out; d_work: nfft doubles */
int gdgb_ir_real_inverse(const double *d_spec, int nfft, double *d_work, double *d_out, long n_keep, void *stream);
/* *d_ss = sum of x[i]^2 (estimateGain without the root, filter.go:127-138) */
int gdgb_ir_sum_squares(const double *d_x, long n, double *d_ss, void *stream);
/* acc[i] += fac * ((comp / sqrt(*d_ss)) * x[i]), i < n: Normalize -> Multiply -> Add of one filter (filter.go:328-336, 270-323, 167-253) */
int gdgb_ir_scale_add(double *d_acc, const double *d_x, long n, double comp, double fac, const double *d_ss, void *stream);
/* H_p[k], p < K, k = 0..B: spectra of the K partitions of B taps (real FFT size 2B) of the L-tap filter d_coeffs */
int gdgb_ir_partition_spectra(const double *d_coeffs, int L, int B, int logB, int K, double *d_H, void *stream);
/* the tuner's autocorrelation for `count` tuners in one batch of launches: ring / work / pos per tuner */
int gdgb_autocorrelate_batch(const double *const *d_rings, int n_ring, const int *pos, double *const *d_works, int count, int fft_size, void *stream);

uint64_t gdgb_launch_count(void);               /* kernels launched so far (bench "gpu_launches") */

#endif
