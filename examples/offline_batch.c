/* Offline batch through the C ABI alone (no Python, no torch): what the cgo wrapper of INTEGRATION.md does, written in C.
 * C channels, each "overdrive (x4 oversampling) -> tone_stack -> delay -> power_amp(synthetic IR)", B blocks of 8192
 * samples at 96 kHz; blocks are submitted with two in flight (gdg_controller_submit / gdg_controller_wait); the stereo
 * master comes from the spatializer; a level meter rides along.
 *
 *   gcc -O2 -Iinclude examples/offline_batch.c -Lgo-dsp-guitar_b200 -lgdg -Wl,-rpath,$PWD/go-dsp-guitar_b200 -lm -o offline_batch
 *   ./offline_batch [channels] [blocks]          (needs a B200: there is no CPU fallback)                                   */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "gdg.h"

#define N 8192
#define FS 96000u

static int die(const char *what) { fprintf(stderr, "%s: %s\n", what, gdg_last_error()); return 1; }

static int unit_type(const char *name)
{
    for (int t = 0; t < gdg_unit_type_count(); t++) {
        const char *a = gdg_unit_type_name(t), *b = name;
        while (*a && *a == *b) { a++; b++; }
        if (!*a && !*b) return t;
    }
    return -1;
}

int main(int argc, char **argv)
{
    const uint32_t C = argc > 1 ? (uint32_t)atoi(argv[1]) : 64;
    const int blocks = argc > 2 ? atoi(argv[2]) : 8;
    if (gdg_init(-1)) return die("gdg_init");

    /* impulse-response bank with one synthetic response (a real host would call gdg_irbank_import("ir/index.json")) */
    gdg_irbank *bank = gdg_irbank_create();
    if (!bank) return die("gdg_irbank_create");
    enum { L = 2048 };
    static double ir[L];
    for (int k = 0; k < L; k++) ir[k] = (k % 7 == 0 ? 0.02 : -0.003) * exp(-6.0 * k / L);
    if (gdg_irbank_add_raw(bank, "demo", 0, ir, L, FS)) return die("gdg_irbank_add_raw");

    gdg_chain **chains = (gdg_chain **)calloc(C, sizeof(*chains));
    for (uint32_t c = 0; c < C; c++) {
        gdg_chain *ch = chains[c] = gdg_chain_create(bank);
        if (!ch) return die("gdg_chain_create");
        const int od = gdg_chain_append_unit(ch, unit_type("overdrive")), ts = gdg_chain_append_unit(ch, unit_type("tone_stack"));
        const int dl = gdg_chain_append_unit(ch, unit_type("delay")), pa = gdg_chain_append_unit(ch, unit_type("power_amp"));
        if (od < 0 || ts < 0 || dl < 0 || pa < 0) return die("gdg_chain_append_unit");
        if (gdg_chain_set_numeric_value(ch, od, "gain", 10 + (int32_t)(c % 20)) || gdg_chain_set_discrete_value(ch, od, "oversampling", "4") ||
            gdg_chain_set_numeric_value(ch, dl, "delay_time", 120) || gdg_chain_set_discrete_value(ch, pa, "filter_1", "demo"))
            return die("parameter");
        for (int id = od; id <= pa; id++) if (gdg_chain_set_bypass(ch, id, 0)) return die("gdg_chain_set_bypass");   /* units start bypassed */
    }
    gdg_spatializer *spat = gdg_spatializer_create(C);
    gdg_meter *meter = gdg_meter_create(2 * C + 3, NULL);
    if (!spat || !meter || gdg_meter_set_enabled(meter, 1)) return die("spatializer / meter");
    for (uint32_t c = 0; c < C; c++) gdg_spatializer_set_azimuth(spat, c, -90.0 + 180.0 * c / (C > 1 ? C - 1 : 1));

    /* two pinned slabs for the inputs and two for the outputs: block k uses slab k % 2 */
    double *in_slab[2], *out_slab[2];
    const double **in_rows[2]; double **out_rows[2];
    for (int s = 0; s < 2; s++) {
        in_slab[s] = (double *)gdg_pinned_alloc((size_t)C * N * sizeof(double));
        out_slab[s] = (double *)gdg_pinned_alloc((size_t)(C + 2) * N * sizeof(double));
        if (!in_slab[s] || !out_slab[s]) return die("gdg_pinned_alloc");
        in_rows[s] = (const double **)calloc(C, sizeof(double *)); out_rows[s] = (double **)calloc(C + 2, sizeof(double *));
        for (uint32_t c = 0; c < C; c++) in_rows[s][c] = in_slab[s] + (size_t)c * N;
        for (uint32_t c = 0; c < C + 2; c++) out_rows[s][c] = out_slab[s] + (size_t)c * N;
    }
    double energy = 0.0;
    int prev = -1, prev_slab = 0;
    for (int k = 0; k <= blocks; k++) {
        int t = -1;
        const int s = k % 2;
        if (k < blocks) {
            for (uint32_t c = 0; c < C; c++)                                  /* "read the next block of every file" */
                for (int i = 0; i < N; i++) {
                    const double tt = (double)(k * N + i) / FS;
                    in_slab[s][(size_t)c * N + i] = 0.4 * sin(2.0 * M_PI * (110.0 + 3.0 * c) * tt) * exp(-3.0 * fmod(tt, 0.5));
                }
            t = gdg_controller_submit(chains, C, NULL, -1, spat, meter, in_rows[s], out_rows[s], C + 2, N, FS);
            if (t < 0) return die("gdg_controller_submit");
        }
        if (prev >= 0) {
            if (gdg_controller_wait(prev)) return die("gdg_controller_wait");
            for (int i = 0; i < N; i++) {                                     /* "write the block": here just the master's energy */
                const double l = out_slab[prev_slab][(size_t)C * N + i], r = out_slab[prev_slab][(size_t)(C + 1) * N + i];
                energy += l * l + r * r;
            }
        }
        prev = t; prev_slab = s;
    }
    int32_t level = 0, peak = 0;
    if (gdg_meter_analyze(meter, 2 * C + 1, &level, &peak)) return die("gdg_meter_analyze");
    printf("%u channels x %d blocks of %d samples @ %u Hz on %s: master energy %.6f, master-left level %d dB (peak %d dB), %llu kernel launches\n",
           C, blocks, N, FS, gdg_backend_name(), energy, level, peak, (unsigned long long)gdg_launch_count());
    for (uint32_t c = 0; c < C; c++) gdg_chain_destroy(chains[c]);
    gdg_spatializer_destroy(spat); gdg_meter_destroy(meter); gdg_irbank_destroy(bank);
    for (int s = 0; s < 2; s++) { gdg_pinned_free(in_slab[s]); gdg_pinned_free(out_slab[s]); free(in_rows[s]); free(out_rows[s]); }
    free(chains);
    return 0;
}
