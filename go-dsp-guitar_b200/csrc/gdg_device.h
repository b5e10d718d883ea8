/*
 * gdg_device.h - POD descriptors shared by the host runtime (gdg_host.cpp) and
 * the sm_100a kernels (gdg_stages.cuh / gdg_kernels.cu).
 *
 * One CTA owns one channel for one audio block: the block (n <= 8192 float64
 * samples) lives in shared memory for the whole chain segment; every effects
 * unit is a CTA-cooperative "stage" that transforms it in place / ping-pong.
 * Unit parameters are snapshotted by the host into UnitDesc (the reference
 * re-reads them by string compare every block, e.g. effects/overdrive.go:29-35);
 * all DSP state that the reference keeps in its Go structs lives in a per-unit
 * device array of 8-byte words (UnitDesc::st) that the kernels update.
 */
#ifndef GDG_DEVICE_H
#define GDG_DEVICE_H

#include <stdint.h>

#ifndef GDG_NT
#define GDG_NT 512              /* threads per CTA                                   */
#endif
#ifndef GDG_LOG_MAX_N
#define GDG_LOG_MAX_N 13
#endif
#define GDG_MAX_N (1 << GDG_LOG_MAX_N)   /* max block length per CTA pass: 8192 (controller.go:36 BLOCK_SIZE) */
#define GDG_MAX_FFT (2 * GDG_MAX_N)      /* max real FFT size held in shared memory            */
#ifndef GDG_MIN_CTAS
#define GDG_MIN_CTAS 1          /* resident CTAs per SM the chain kernel is compiled for */
#endif
#define GDG_MAX_SLOTS 24        /* non-bypassed units per fused segment              */
#define GDG_SCAN_LANES 4        /* independent recurrences scanned together          */
#ifndef GDG_OS_TILE
#define GDG_OS_TILE (GDG_MAX_N / 8)      /* base-rate samples per oversampling tile            */
#endif
#define GDG_NP 12               /* double parameters per unit                        */
#define GDG_NI 10               /* int parameters per unit                           */

/* unit type codes == effects/effects.go:21-43 */
enum {
    GDG_UNIT_SIGNALGENERATOR = 0, GDG_UNIT_NOISEGATE, GDG_UNIT_BANDPASS, GDG_UNIT_AUTOWAH, GDG_UNIT_AUTOYOY,
    GDG_UNIT_COMPRESSOR, GDG_UNIT_OCTAVER, GDG_UNIT_EXCESS, GDG_UNIT_FUZZ, GDG_UNIT_OVERDRIVE,
    GDG_UNIT_DISTORTION, GDG_UNIT_TONESTACK, GDG_UNIT_CHORUS, GDG_UNIT_FLANGER, GDG_UNIT_PHASER,
    GDG_UNIT_TREMOLO, GDG_UNIT_RINGMODULATOR, GDG_UNIT_DELAY, GDG_UNIT_REVERB, GDG_UNIT_POWERAMP,
    GDG_UNIT_CABINET, GDG_UNIT_COUNT
};

/* sample formats of the encoded entry points == the (format, bit depth) pairs of wave/wave.go:737-838 */
enum { GDG_FMT_PCM8 = 0, GDG_FMT_PCM16, GDG_FMT_PCM24, GDG_FMT_PCM32, GDG_FMT_IEEE32, GDG_FMT_IEEE64 };

/* follower modes (compressor.go:44-58 and friends) */
enum { GDG_FOLLOW_ENVELOPE = 0, GDG_FOLLOW_LEVEL = 1, GDG_FOLLOW_OTHER = 2 };

typedef struct {
    int32_t type;
    int32_t i[GDG_NI];          /* resolved integer parameters (meaning per unit, see gdg_host.cpp) */
    double p[GDG_NP];           /* resolved float64 parameters                                      */
    double *st;                 /* device state words (float64 / int64 bit patterns)                */
    const double *aux;          /* read-only device data (power_amp: filter spectrum, interleaved)  */
} gdg_unit_desc;

typedef struct {
    int32_t nslots;
    int32_t pad;
    gdg_unit_desc slot[GDG_MAX_SLOTS];
} gdg_chain_desc;

/* ---- state-word layouts (indices into UnitDesc::st) ----------------------- */
/* oversampled units (excess/fuzz/overdrive/distortion) */
#define GDG_OS_ST_ENV 0
#define GDG_OS_ST_CCV 1
#define GDG_OS_ST_UP2 2            /* 8 words: last 8 inputs of the x2 up-sampler (oversampling.go:91-93) */
#define GDG_OS_ST_UP4 10           /* 8 words: x4                                                          */
#define GDG_OS_ST_Z2 18            /* 2 x 41 words: per polyphase branch, the last 41 oversampled samples
                                      entering the x2 decimator (its 77-tap history, de-interleaved)      */
#define GDG_OS_ST_Z4 100           /* 4 x 41 words: x4 (155-tap history)                                   */
#define GDG_OS_ST_WORDS 264

/* delay-type units: word 0 = phase / envelope, word 1 = ring write position, ring from word 2 */
#define GDG_RING_ST_POS 1
#define GDG_RING_ST_DATA 2

/* reverb: word 0 line-ring pos, words 1..3 all-pass write pointers; ring offsets in desc.i[] */
#define GDG_RV_ST_LINEPOS 0
#define GDG_RV_ST_APPTR 1
#define GDG_RV_ST_DATA 4

/* spatializer channel descriptor */
typedef struct {
    double fac_left, fac_right;
    double delay_time;             /* sign selects the delayed ear (spatializer.go:286-292) */
    double w_early, w_late;
    int32_t d_early, d_late;       /* clamped integer delays                                  */
} gdg_spat_channel;

#endif
