/*
 * orc_bench.c - CPU-baseline driver for bench.py (TEST / MEASUREMENT INFRASTRUCTURE, never linked by the product).
 *
 * Times the restated reference the way the reference runs: one worker per channel up to the number of host threads
 * (controller/controller.go:3339-3341 starts one processing goroutine per channel; controller.process, :2682-2705, hands
 * every channel's block to them and waits for all, then runs the spatializer on the calling thread, :2761), block after
 * block.  A persistent pthread pool with pinned threads and two barriers per block replaces the goroutines/channels, so
 * that the figure is reproducible from box to box (round 1 spawned Python threads per block and the FFT called
 * malloc/free per transform; its numbers moved by 3x).
 */
#define _GNU_SOURCE
#include "orc.h"

#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
    orc_chain *const *chains; uint32_t C; orc_spatializer *spat;
    const double *x; double *out; size_t n, blocks; uint32_t rate;
    int threads, pin; pthread_barrier_t bar; const int *cpus; int ncpus;
    double seconds;
} bench_t;
typedef struct { bench_t *b; int t; } worker_t;

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; }

static void *worker(void *arg)
{
    worker_t *w = (worker_t *)arg; bench_t *b = w->b;
    if (b->pin && b->ncpus > 0) {
        cpu_set_t set; CPU_ZERO(&set); CPU_SET(b->cpus[w->t % b->ncpus], &set);
        pthread_setaffinity_np(pthread_self(), sizeof set, &set);
    }
    const size_t total = b->n * b->blocks;
    double *const *rows = NULL;
    double **outp = (double **)malloc(sizeof(double *) * (b->C + 2));
    for (uint32_t c = 0; c < b->C + 2; c++) outp[c] = b->out + (size_t)c * b->n;
    rows = outp;
    pthread_barrier_wait(&b->bar);
    double t0 = 0.0;
    if (w->t == 0) t0 = now_s();
    for (size_t k = 0; k < b->blocks; k++) {
        for (uint32_t c = (uint32_t)w->t; c < b->C; c += (uint32_t)b->threads)
            orc_chain_process(b->chains[c], b->x + (size_t)c * total + k * b->n, outp[c], b->n, b->rate);
        pthread_barrier_wait(&b->bar);                       /* all channels of block k done */
        if (w->t == 0 && b->spat) orc_spatializer_process(b->spat, (const double *const *)rows, b->C, NULL, outp + b->C, 2, b->n);
        pthread_barrier_wait(&b->bar);                       /* the caller's thread has mixed block k */
    }
    if (w->t == 0) b->seconds = now_s() - t0;
    free(outp);
    return NULL;
}

/* `blocks` consecutive blocks of n samples for C channels (x: [C][blocks * n]) on `threads` workers; returns the wall time
 * in seconds (excluding thread start-up), < 0 on error.  pin != 0 binds worker t to the t-th CPU the process may use. */
double orc_bench_controller(orc_chain *const *chains, uint32_t C, orc_spatializer *spat, const double *x, size_t n, size_t blocks,
                            uint32_t rate, int threads, int pin)
{
    if (!chains || !C || !x || !n || !blocks || threads < 1) return -1.0;
    if ((uint32_t)threads > C) threads = (int)C;
    bench_t b; memset(&b, 0, sizeof b);
    b.chains = chains; b.C = C; b.spat = spat; b.x = x; b.n = n; b.blocks = blocks; b.rate = rate; b.threads = threads; b.pin = pin;
    b.out = (double *)calloc((size_t)(C + 2) * n, sizeof(double));
    cpu_set_t allowed; CPU_ZERO(&allowed);
    int *cpus = (int *)malloc(sizeof(int) * CPU_SETSIZE); int ncpus = 0;
    if (sched_getaffinity(0, sizeof allowed, &allowed) == 0)
        for (int i = 0; i < CPU_SETSIZE; i++) if (CPU_ISSET(i, &allowed)) cpus[ncpus++] = i;
    b.cpus = cpus; b.ncpus = ncpus;
    pthread_barrier_init(&b.bar, NULL, (unsigned)threads);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    worker_t *ws = (worker_t *)malloc(sizeof(worker_t) * (size_t)threads);
    int started = 0;
    for (int t = 0; t < threads; t++) { ws[t].b = &b; ws[t].t = t; if (pthread_create(&th[t], NULL, worker, &ws[t]) != 0) break; started++; }
    if (started != threads) { b.seconds = -1.0; /* cannot release the barrier safely */ exit(3); }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    pthread_barrier_destroy(&b.bar);
    free(th); free(ws); free(cpus); free(b.out);
    return b.seconds;
}
