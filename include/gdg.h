/*
 * gdg.h - C ABI of libgdg: the B200-native effects-chain engine behind
 * go-dsp-guitar's signal.Chain / effects.Unit / spatializer.Spatializer /
 * tuner.Tuner interfaces.  Plain pointers and sizes only; this is what a cgo
 * wrapper binds.  Each entry point names the reference
 * interface it replaces (paths relative to the go-dsp-guitar tree).
 *
 * Conventions
 *   - every function that can fail returns 0 on success / a negative value on
 *     failure (or NULL for constructors); gdg_last_error() then returns the
 *     message, using the reference's own error strings where it has one
 *     (effects/effects.go:165-312, signal/signal.go:59-337);
 *   - audio buffers are float64, one contiguous array per channel (the
 *     reference's []float64); block lengths must be powers of two <= 8192
 *     (controller/controller.go:36 BLOCK_SIZE; filter.Process is only safe for
 *     powers of two, filter/filter.go:372-455) - other sizes are REJECTED, where
 *     the reference would panic or silently misbehave;
 *   - there is NO CPU fallback: without a CUDA device gdg_init() fails and every
 *     processing call fails;
 *   - threading (effects/effects.go:98, signal/signal.go:45: RWMutex per unit / chain): every mutator and getter of a
 *     chain may be called from any thread while another thread processes blocks - they serialise against the
 *     "resolve parameters + enqueue kernels" part of the processing calls, a change takes effect at the next block
 *     boundary.  Processing the SAME chain from two threads at once is not allowed (as in the reference);
 *     gdg_last_error() is per thread: fetch it on the thread that made the failing call (cgo: runtime.LockOSThread,
 *     see INTEGRATION.md).  The other objects keep the reference's own locks: a spatializer's setters / getters and
 *     SetSampleRate against its Process (spatializer/spatializer.go:59), a tuner's Analyze against Process and against
 *     other Analyze calls (tuner/tuner.go:50-53), a level meter's SetEnabled against Process (level/level.go:40-52);
 *   - any OS thread may call in: CUDA's "current device" is per thread, the library binds the calling thread to the device
 *     of the object it is handed (or of the context selected with gdg_use_device) on entry, so goroutines that migrate
 *     between threads need no cudaSetDevice of their own.
 */
#ifndef GDG_H
#define GDG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDG_VERSION "0.1.0"

/* unit type codes: effects/effects.go:21-43 (UNIT_SIGNALGENERATOR .. UNIT_CABINET) */
enum {
    GDG_SIGNALGENERATOR = 0, GDG_NOISEGATE, GDG_BANDPASS, GDG_AUTOWAH, GDG_AUTOYOY, GDG_COMPRESSOR, GDG_OCTAVER,
    GDG_EXCESS, GDG_FUZZ, GDG_OVERDRIVE, GDG_DISTORTION, GDG_TONESTACK, GDG_CHORUS, GDG_FLANGER, GDG_PHASER,
    GDG_TREMOLO, GDG_RINGMODULATOR, GDG_DELAY, GDG_REVERB, GDG_POWERAMP, GDG_CABINET
};
/* parameter types: effects/effects.go:12-16 */
enum { GDG_PARAMETER_TYPE_INVALID = 0, GDG_PARAMETER_TYPE_DISCRETE = 1, GDG_PARAMETER_TYPE_NUMERIC = 2 };

typedef struct gdg_irbank gdg_irbank;           /* filter.ImpulseResponses   filter/filter.go:79-82  */
typedef struct gdg_chain gdg_chain;             /* signal.Chain              signal/signal.go:21-36  */
typedef struct gdg_spatializer gdg_spatializer; /* spatializer.Spatializer   spatializer/spatializer.go:30-41 */
typedef struct gdg_tuner gdg_tuner;             /* tuner.Tuner               tuner/tuner.go:62-65    */
typedef struct gdg_meter gdg_meter;             /* level.Meter               level/level.go:48-65    */

/* ---- runtime ----------------------------------------------------------------- */
/* Select the CUDA device (one process per GPU) and upload the constant tables.  Called
 * implicitly (device = current CUDA device) by the first constructor if omitted. */
int gdg_init(int device);
const char *gdg_last_error(void);
const char *gdg_backend_name(void);
const char *gdg_version(void);
/* number of kernels launched so far (for bench.py's "gpu_launches") */
uint64_t gdg_launch_count(void);
/* page-locked host memory a caller can hand to gdg_process_batch for zero-staging copies */
void *gdg_pinned_alloc(size_t bytes);
void gdg_pinned_free(void *p);
/* the library's CUDA stream (cudaStream_t) for callers that own device buffers */
void *gdg_stream(void);
int gdg_synchronize(void);

/* effects.UnitTypes() / effects.ParameterTypes(): effects/effects.go:521-568 */
int gdg_unit_type_count(void);
const char *gdg_unit_type_name(int unit_type);

/* ---- impulse responses: filter.Import  filter/filter.go:704-802 ------------------ */
/* index.json ([{"Name","Path","Compensation"}...]); WAV paths are resolved relative to the
 * current working directory like the reference (os.ReadFile(wavePath)).  Unreadable or
 * non-mono files are skipped with a warning on stdout, like the reference. */
gdg_irbank *gdg_irbank_import(const char *index_json_path);
gdg_irbank *gdg_irbank_create(void);
void gdg_irbank_destroy(gdg_irbank *bank);
/* one descriptor's worth of filter.Import: resample to the 7 supported rates (filter/filter.go:25-33,766-780) */
int gdg_irbank_add(gdg_irbank *bank, const char *name, int32_t compensation_db,
                   const double *samples, size_t n, uint32_t source_rate);
/* one response at exactly one rate (what filter.FromCoefficients-style synthetic IRs need) */
int gdg_irbank_add_raw(gdg_irbank *bank, const char *name, int32_t compensation_db,
                       const double *samples, size_t n, uint32_t rate);
size_t gdg_irbank_name_count(const gdg_irbank *bank);         /* ImpulseResponses.Names() filter.go:681-699 */
const char *gdg_irbank_name(const gdg_irbank *bank, size_t i);

/* ---- signal.Chain  signal/signal.go ------------------------------------------------ */
gdg_chain *gdg_chain_create(gdg_irbank *bank);                /* signal.CreateChain            signal.go:419  */
void gdg_chain_destroy(gdg_chain *chain);
int gdg_chain_append_unit(gdg_chain *chain, int unit_type);   /* AppendUnit -> id | -1 (starts BYPASSED) signal.go:52-86 */
int gdg_chain_remove_unit(gdg_chain *chain, int id);          /* RemoveUnit                    signal.go:91-111 */
int gdg_chain_move_up(gdg_chain *chain, int id);              /* MoveUp                        signal.go:116-136 */
int gdg_chain_move_down(gdg_chain *chain, int id);            /* MoveDown                      signal.go:141-162 */
int gdg_chain_unit_type(gdg_chain *chain, int id);            /* UnitType -> type | -1         signal.go:167-185 */
int gdg_chain_set_bypass(gdg_chain *chain, int id, int bypass);           /* SetBypass        signal.go:190-218 */
int gdg_chain_get_bypass(gdg_chain *chain, int id, int *bypass);          /* GetBypass        signal.go:223-240 */
int gdg_chain_set_discrete_value(gdg_chain *chain, int id, const char *name, const char *value);   /* signal.go:245-263, effects.go:127-172 */
int gdg_chain_get_discrete_value(gdg_chain *chain, int id, const char *name, const char **value);  /* signal.go:268-286 */
int gdg_chain_set_numeric_value(gdg_chain *chain, int id, const char *name, int32_t value);        /* signal.go:291-308, effects.go:266-317 */
int gdg_chain_get_numeric_value(gdg_chain *chain, int id, const char *name, int32_t *value);       /* signal.go:313-330 */
int gdg_chain_length(gdg_chain *chain);                       /* Length                        signal.go:352-358 */
/* Parameters(id): effects.Parameter (effects/effects.go:69-78), flattened */
int gdg_chain_parameter_count(gdg_chain *chain, int id);
typedef struct {
    const char *name;
    int32_t type;                 /* GDG_PARAMETER_TYPE_* */
    const char *physical_unit;
    int32_t minimum, maximum, numeric_value;
    int32_t discrete_value_index;
    int32_t discrete_value_count;
} gdg_parameter;
int gdg_chain_parameter(gdg_chain *chain, int id, int index, gdg_parameter *out);
const char *gdg_chain_parameter_discrete_value(gdg_chain *chain, int id, int index, int value_index);

/* Filter.Coefficients() (filter/filter.go:258-265) of the composite filter a power amp compiled (effects/poweramp.go:25-127:
 * Reduce / Normalize / Multiply / Add run on the device).  Returns the number of coefficients (0 when no
 * filter has been compiled yet - the sample rate is only known at the first Process) and copies them when capacity suffices. */
long gdg_chain_filter_coefficients(gdg_chain *chain, int id, double *out, size_t capacity);

/* Chain.Process(in, out []float64, sampleRate uint32)  signal.go:361-414.
 * HOST buffers, synchronous: H2D, fused segment kernel(s), D2H.  The JACK-path drop-in. */
int gdg_chain_process(gdg_chain *chain, const double *in, double *out, size_t n, uint32_t sample_rate);

/* The per-block fan-out of controller.process (controller/controller.go:2682-2705): all chains,
 * one block.  HOST buffers (in[c], out[c] point to n doubles each; pinned and/or contiguous
 * buffers are detected and copied without staging). */
int gdg_process_batch(gdg_chain *const *chains, size_t channels, const double *const *in, double *const *out,
                      size_t n, uint32_t sample_rate);
/* Same with DEVICE-resident planar buffers: channel c at d_in + c*stride.  Asynchronous on `stream`
 * (NULL = gdg_stream()). */
int gdg_process_batch_device(gdg_chain *const *chains, size_t channels, const double *d_in, double *d_out,
                             size_t stride, size_t n, uint32_t sample_rate, void *stream);

/* ---- "next" rows of the path: the steps either side of the block loop ----------------
 * Sample formats == the (format, bit depth) pairs of wave.samplesToBytes / bytesToSamples (wave/wave.go:737-838). */
enum { GDG_FORMAT_PCM8 = 0, GDG_FORMAT_PCM16, GDG_FORMAT_PCM24, GDG_FORMAT_PCM32, GDG_FORMAT_IEEE32, GDG_FORMAT_IEEE64 };
size_t gdg_format_bytes_per_sample(int format);
/* gdg_process_batch with the sample codecs of wave/wave.go:275-732 fused into the kernel's load and store: row c of `in`
 * holds n little-endian samples of in_format (the data-chunk payload of one channel), row c of `out` receives n samples of
 * out_format.  Replaces wave.Channel.Floats() -> process -> Channel.WriteFloats()/File.Bytes() of the batch path
 * (controller/controller.go:2908-3143) for one block; chains of at most one fused segment (24 active units). */
int gdg_process_batch_encoded(gdg_chain *const *chains, size_t channels, const void *const *in, int in_format,
                              void *const *out, int out_format, size_t n, uint32_t sample_rate);
/* the same block submitted without waiting: returns a ticket for gdg_controller_wait() (see gdg_controller_submit below) */
int gdg_process_batch_encoded_submit(gdg_chain *const *chains, size_t channels, const void *const *in, int in_format,
                                     void *const *out, int out_format, size_t n, uint32_t sample_rate);
int gdg_process_batch_device_encoded(gdg_chain *const *chains, size_t channels, const void *d_in, int in_format, size_t in_stride_bytes,
                                     void *d_out, int out_format, size_t out_stride_bytes, size_t n, uint32_t sample_rate, void *stream);
/* resample.Time (resample/resample.go:72-103): the batch path's input rate conversion (controller.go:2991-3003).
 * Returns the number of samples written (== gdg_resample_time_length) or -1. */
size_t gdg_resample_time_length(size_t n, uint32_t source_rate, uint32_t target_rate);
long gdg_resample_time(const double *in, size_t n, uint32_t source_rate, uint32_t target_rate, double *out, size_t out_capacity);

/* ---- spatializer.Spatializer  spatializer/spatializer.go ---------------------------- */
gdg_spatializer *gdg_spatializer_create(uint32_t input_channels);          /* Create           :436-469 */
void gdg_spatializer_destroy(gdg_spatializer *s);
int gdg_spatializer_set_azimuth(gdg_spatializer *s, uint32_t channel, double azimuth);    /* :340-355 */
int gdg_spatializer_get_azimuth(gdg_spatializer *s, uint32_t channel, double *azimuth);   /* :63-78   */
int gdg_spatializer_set_distance(gdg_spatializer *s, uint32_t channel, double distance);  /* :360-383 */
int gdg_spatializer_get_distance(gdg_spatializer *s, uint32_t channel, double *distance); /* :83-98   */
int gdg_spatializer_set_level(gdg_spatializer *s, uint32_t channel, double level);        /* :388-413 */
int gdg_spatializer_get_level(gdg_spatializer *s, uint32_t channel, double *level);       /* :103-118 */
uint32_t gdg_spatializer_input_count(gdg_spatializer *s);                                 /* :123-125 */
uint32_t gdg_spatializer_output_count(gdg_spatializer *s);                                /* :130-132 */
int gdg_spatializer_set_sample_rate(gdg_spatializer *s, uint32_t rate);                   /* :418-431 */
/* Process(inputBuffers, auxInputBuffer|nil, outputBuffers[2])  :140-335  (HOST buffers) */
int gdg_spatializer_process(gdg_spatializer *s, const double *const *in, uint32_t n_in, const double *aux,
                            double *const *out, uint32_t n_out, size_t n);
/* DEVICE buffers.  `accumulate` != 0 adds onto d_out (used for the multi-GPU partial sums);
 * channel c at d_in + c*stride, outputs d_out (left) and d_out + out_stride (right). */
int gdg_spatializer_process_device(gdg_spatializer *s, const double *d_in, size_t stride, const double *d_aux,
                                   double *d_out, size_t out_stride, size_t n, int accumulate, void *stream);

/* ---- tuner.Tuner  tuner/tuner.go ---------------------------------------------------- */
gdg_tuner *gdg_tuner_create(void);                                                        /* Create  :592-607 */
void gdg_tuner_destroy(gdg_tuner *t);
int gdg_tuner_process(gdg_tuner *t, const double *samples, size_t n, uint32_t sample_rate);        /* Process :582-587 (HOST) */
int gdg_tuner_process_device(gdg_tuner *t, const double *d_samples, size_t n, uint32_t sample_rate, void *stream);
typedef struct { int8_t cents; double frequency; const char *note; } gdg_tuner_result;    /* tuner.Result :27-43 */
int gdg_tuner_analyze(gdg_tuner *t, gdg_tuner_result *result);                            /* Analyze :379-577 */
/* Analyze for `count` tuners at once (one per channel strip, polled together by the UI, controller.go:625-627): the 262144-point
 * autocorrelations run as ONE batch of launches, results[i] belongs to tuners[i] */
int gdg_tuner_analyze_batch(gdg_tuner *const *tuners, uint32_t count, gdg_tuner_result *results);

/* ---- level.Meter  level/level.go ------------------------------------------------------ */
/* CreateMeter(numChannels, names) :318-399 (names may be NULL: empty names); created disabled */
gdg_meter *gdg_meter_create(uint32_t channels, const char *const *names);
void gdg_meter_destroy(gdg_meter *m);
uint32_t gdg_meter_channel_count(gdg_meter *m);                                           /* ChannelCount :260-265 */
const char *gdg_meter_channel_name(gdg_meter *m, uint32_t channel);                       /* ChannelName  :270-287 (NULL + error when out of range) */
int gdg_meter_enabled(gdg_meter *m);                                                      /* Enabled      :292-297 */
int gdg_meter_set_enabled(gdg_meter *m, int value);                                       /* SetEnabled   :333-356 (disabling clears the state, :214-237) */
/* Process(buffers, sampleRate) :302-328 -> channelMeter.process :147-208.  HOST buffers, one per channel (a NULL row is a
 * nil buffer: no samples); fails when n_buffers != channel count; a disabled meter does nothing. */
int gdg_meter_process(gdg_meter *m, const double *const *buffers, uint32_t n_buffers, size_t n, uint32_t sample_rate);
/* DEVICE rows: `rows` consecutive meter channels from first_channel, row r at d_rows + r*stride */
int gdg_meter_process_device(gdg_meter *m, uint32_t first_channel, const double *d_rows, size_t stride, uint32_t rows, size_t n,
                             uint32_t sample_rate, void *stream);
int gdg_meter_analyze(gdg_meter *m, uint32_t channel, int32_t *level, int32_t *peak);    /* Analyze :242-255 -> :95-137 (dB, rounded) */
/* the raw state of one channel meter (currentValue, peakValue, sampleCounter: level.go:37-44); for tests and persistence */
int gdg_meter_state(gdg_meter *m, uint32_t channel, double *current_value, double *peak_value, uint64_t *sample_counter);

/* ---- controller.process mirror (controller/controller.go:2648-2783) ------------------- */
/* tuner feed (channel tuner_channel, or -1) -> chains -> spatializer into out[C], out[C+1].
 * HOST buffers; n_out must be >= channels (+2 when spat != NULL).  The metronome is
 * the caller's business (out of scope). */
int gdg_controller_process(gdg_chain *const *chains, uint32_t channels, gdg_tuner *tuner, int tuner_channel,
                           gdg_spatializer *spat, const double *const *in, double *const *out, uint32_t n_out,
                           size_t n, uint32_t sample_rate);
/* the same with the level meter of controller.go:2653-2781 fed on the device: meter channels [0,C) inputs, [C,2C) chain
 * outputs, 2C the metronome port (not fed here), 2C+1 / 2C+2 master left / right; meter may be NULL */
int gdg_controller_process_metered(gdg_chain *const *chains, uint32_t channels, gdg_tuner *tuner, int tuner_channel,
                                   gdg_spatializer *spat, gdg_meter *meter, const double *const *in, double *const *out,
                                   uint32_t n_out, size_t n, uint32_t sample_rate);
/* The same block WITHOUT waiting for it: returns a ticket (0 or 1, at most two blocks in flight) or -1.  The rows of `in`
 * must stay untouched and the rows of `out` are undefined until gdg_controller_wait(ticket) returns.  Blocks run in
 * submission order (unit state carries from one to the next), so a batch loop - controller.go:2991-3106 walks the file
 * block by block - submits block k+1 before it waits for block k: the upload of k+1 and the download of k overlap the
 * kernels.  Pinned, contiguous rows (gdg_pinned_alloc) make both copies truly asynchronous. */
int gdg_controller_submit(gdg_chain *const *chains, uint32_t channels, gdg_tuner *tuner, int tuner_channel,
                          gdg_spatializer *spat, gdg_meter *meter, const double *const *in, double *const *out,
                          uint32_t n_out, size_t n, uint32_t sample_rate);
int gdg_controller_wait(int ticket);
/* The same two calls WITH the metronome of controller.go:2722-2761: `metronome` is the host row (n float64 samples) that
 * metronome.Process just wrote (the reference hands it to the spatializer as auxInputBuffer, spatializer.go:300-308), or NULL
 * when there is no metronome.  It feeds level-meter channel 2C (controller.go:2735-2738) and - when metronome_to_master is
 * non-zero (controller.go:2746-2748 metrMasterOutput) - is added to both master channels.  The metronome generator itself
 * stays on the host (out of scope).  With the _submit call the row is like the rows of `in`: leave it untouched
 * until gdg_controller_wait (a pageable row has been staged when the call returns, a pinned one is read asynchronously). */
int gdg_controller_process_aux(gdg_chain *const *chains, uint32_t channels, gdg_tuner *tuner, int tuner_channel,
                               gdg_spatializer *spat, gdg_meter *meter, const double *const *in, const double *metronome,
                               int metronome_to_master, double *const *out, uint32_t n_out, size_t n, uint32_t sample_rate);
int gdg_controller_submit_aux(gdg_chain *const *chains, uint32_t channels, gdg_tuner *tuner, int tuner_channel,
                              gdg_spatializer *spat, gdg_meter *meter, const double *const *in, const double *metronome,
                              int metronome_to_master, double *const *out, uint32_t n_out, size_t n, uint32_t sample_rate);
/* gdg_controller_submit on ENCODED host rows (the data-chunk payload of wave/wave.go's formats, GDG_FORMAT_*): rows are
 * decoded to float64 on the device, tuner feed / chains / spatializer / meter run as above, the C (+2) output rows come
 * back encoded - the whole of controller.go:2908-3143 for one block without a float64 ever crossing PCIe. */
int gdg_controller_submit_encoded(gdg_chain *const *chains, uint32_t channels, gdg_tuner *tuner, int tuner_channel,
                                  gdg_spatializer *spat, gdg_meter *meter, const void *const *in, int in_format,
                                  void *const *out, int out_format, uint32_t n_out, size_t n, uint32_t sample_rate);

/* ---- several GPUs in ONE process ---------------------------------------------------------------------------
 * The reference host is one process with one worker per channel (controller/controller.go:3339-3341); a cgo host cannot
 * call torch.distributed.  Channels shard over devices in contiguous ranges; every object lives on the
 * device that was selected on the creating thread (gdg_use_device / gdg_cluster_select); the only exchange is the stereo
 * master: member 0 sums the members' partial masters with ONE kernel that reads the other devices' buffers through peer
 * (NVLink 5 / NVSwitch) addresses - no NCCL, no host round trip. */
typedef struct gdg_cluster gdg_cluster;
int gdg_device_count(void);
int gdg_use_device(int device);                   /* thread-local: objects created afterwards live on `device`; returns >= 0 */
int gdg_current_device(void);
gdg_cluster *gdg_cluster_create(const int *devices, int n_devices);
void gdg_cluster_destroy(gdg_cluster *cluster);
int gdg_cluster_size(gdg_cluster *cluster);
int gdg_cluster_select(gdg_cluster *cluster, int member);       /* == gdg_use_device(devices[member]) */
/* controller.process over the cluster (see gdg_controller_submit): chains[c] in member order, spats[m] = member m's
 * spatializer over its own channels (NULL array: no master); ticketed like gdg_controller_submit (two blocks in flight) */
int gdg_cluster_submit(gdg_cluster *cluster, gdg_chain *const *chains, uint32_t channels, gdg_tuner *tuner, int tuner_channel,
                       gdg_spatializer *const *spats, const double *const *in, double *const *out, uint32_t n_out, size_t n, uint32_t sample_rate);
int gdg_cluster_wait(gdg_cluster *cluster, int ticket);
/* The same block WITH the level meter and the metronome of controller.go:2710-2771.  meters[m] (array of gdg_cluster_size
 * entries, or NULL; entries may be NULL) is member m's level meter, created on member m's device with 2 * cnt_m + 3 channels
 * for its cnt_m channels: [0, cnt_m) inputs, [cnt_m, 2 cnt_m) chain outputs, then the metronome port and master left / right -
 * those three are fed on member 0 only (it owns the summed master); the reference's meter channel c of controller.go:2710-2714
 * is channel c - first_m of member m's meter, its output channel C + c is cnt_m + c - first_m.  `metronome`: see
 * gdg_controller_submit_aux; the row is added to the master inside the peer-memory sum (spatializer.go:300-308). */
int gdg_cluster_submit_aux(gdg_cluster *cluster, gdg_chain *const *chains, uint32_t channels, gdg_tuner *tuner, int tuner_channel,
                           gdg_spatializer *const *spats, gdg_meter *const *meters, const double *const *in, const double *metronome,
                           int metronome_to_master, double *const *out, uint32_t n_out, size_t n, uint32_t sample_rate);
/* ... and on ENCODED host rows (GDG_FORMAT_*, see gdg_controller_submit_encoded): every member decodes / encodes its own rows
 * inside its segment kernels, the master pair comes back from member 0 in out_format */
int gdg_cluster_submit_encoded(gdg_cluster *cluster, gdg_chain *const *chains, uint32_t channels, gdg_tuner *tuner, int tuner_channel,
                               gdg_spatializer *const *spats, gdg_meter *const *meters, const void *const *in, int in_format,
                               void *const *out, int out_format, uint32_t n_out, size_t n, uint32_t sample_rate);
int gdg_cluster_process(gdg_cluster *cluster, gdg_chain *const *chains, uint32_t channels, gdg_tuner *tuner, int tuner_channel,
                        gdg_spatializer *const *spats, const double *const *in, double *const *out, uint32_t n_out, size_t n, uint32_t sample_rate);

#ifdef __cplusplus
}
#endif
#endif
