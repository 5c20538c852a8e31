/*
 * groove_b200.h — C ABI of the B200-native block renderer for Groove's
 * per-sample synthesis + effects hot path.
 *
 * The reference (sowbug/groove, Rust) has no FFI; its "plugin API" is the
 * trait family the Orchestrator consumes.  Each entry point below replaces
 * one of those interfaces (paths relative to the reference tree):
 *
 *   gb_create / gb_destroy     Orchestrator::new_with            orchestration/src/orchestrator.rs:522-568
 *   gb_add_instrument/effect   Orchestrator::add_with_uvid +     settings/src/songs.rs:106-132,
 *                              *Params structs (derive(Params))  settings/src/{instruments,effects}.rs, proc-macros/src/params.rs:14-151
 *   gb_load_sample             Sampler/Drumkit::new_with(paths)  settings/src/instruments.rs:81-88
 *   gb_patch                   Orchestrator::patch               orchestration/src/orchestrator.rs:263-304
 *   gb_finalize                the author's "snapshot the walk"  orchestration/src/orchestrator.rs:357-359 (TODO)
 *   gb_push_events             HandlesMidi::handle_midi_message, orchestration/src/orchestrator.rs:710-754,
 *                              Controllable::control_set_param   proc-macros/src/control.rs:178-185,237-249
 *   gb_render_block            Orchestrator::tick(&mut [StereoSample]) -> frames done
 *                              = handle_work + gather_audio      orchestration/src/orchestrator.rs:856-877,367-470
 *   gb_render_pcm16            IOHelper::send_performance_to_file orchestration/src/helpers.rs:74-97
 *   gb_save_state/restore      (state carried across tick() calls; serde of Orchestrator)
 *
 * Conventions: every call returns 0 on success or a negative GB_E* code; the
 * message is available from gb_last_error().  Nothing throws across the
 * boundary.  One engine must be driven by one thread at a time (the
 * reference holds its Orchestrator under a Mutex: src/panels/legacy/audio_panel.rs:54-111).
 * All buffers crossing the boundary are HOST memory; audio is f64 interleaved
 * L,R — bit-compatible with the reference's [StereoSample] = [(f64,f64)].
 * There is no CPU fallback: without a CUDA device gb_create fails.
 */
#ifndef GROOVE_B200_H
#define GROOVE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GB_ABI_VERSION 4 /* 2: gb_link_control, GB_FX_SIGNAL_PASSTHROUGH, gb_stats grew (rest_kernel_*);
                            3: GB_INST_OSCILLATOR / GB_INST_ENVELOPE, gb_stats grew (solo_*, fm_*, idle_*, *_ctas),
                               gb_set_lookahead;
                            4: gb_stats grew (rest_tp_launches, rest_vr_launches, rest_vr16_launches), gb_bus_exchange_* */

/* ---- error codes --------------------------------------------------------- */
enum {
  GB_OK = 0,
  GB_EINVAL = -1,   /* bad argument / bad params */
  GB_ENOENT = -2,   /* unknown uid */
  GB_ESTATE = -3,   /* call not valid in this engine state (e.g. patch after finalize) */
  GB_ENODEV = -4,   /* no usable CUDA device */
  GB_ECUDA = -5,    /* CUDA runtime failure (message has the detail) */
  GB_ENOMEM = -6,
  GB_EGRAPH = -7    /* patch graph violation (cycle, effect expected, ...) */
};

/* ---- entity kinds -------------------------------------------------------- */
typedef enum {
  /* instruments (leaves of the patch graph) */
  GB_INST_WELSH = 1,     /* gb_welsh_params     — WelshSynth      settings/src/patches.rs:110-169 */
  GB_INST_FM = 2,        /* gb_fm_params        — FmSynth         settings/src/patches.rs:691-715 */
  GB_INST_SAMPLER = 3,   /* gb_sampler_params   — Sampler         settings/src/instruments.rs:34-36 */
  GB_INST_DRUMKIT = 4,   /* gb_drumkit_params   — Drumkit         settings/src/instruments.rs:34-36 */
  GB_INST_TOY_SOURCE = 5,/* gb_toy_source_params— ToyAudioSource  orchestration/src/orchestrator.rs:1415,1445-1668 */
  GB_INST_OSCILLATOR = 6,/* gb_oscillator_source_params — a bare Oscillator as a device ("oscillator": 58 of the reference's
                            project fixtures, e.g. projects/demos/effects/filter-*.json, projects/demos/instruments/oscillator-*.json):
                            free-running from frame 0 at a fixed frequency, mono on both channels, ignores MIDI */
  GB_INST_ENVELOPE = 7,  /* gb_envelope_source_params — a bare Envelope as a device (projects/demos/instruments/
                            envelope-adsr-linear.json): any note-on triggers it, note-off releases; output = its level on both channels */
  /* effects (inner nodes) */
  GB_FX_MIXER = 32,      /* no params           — Mixer (main-mixer is created by gb_create) */
  GB_FX_GAIN = 33,       /* gb_gain_params */
  GB_FX_LIMITER = 34,    /* gb_limiter_params */
  GB_FX_BITCRUSHER = 35, /* gb_bitcrusher_params */
  GB_FX_COMPRESSOR = 36, /* gb_compressor_params */
  GB_FX_DELAY = 37,      /* gb_delay_params */
  GB_FX_CHORUS = 38,     /* gb_chorus_params */
  GB_FX_REVERB = 39,     /* gb_reverb_params */
  GB_FX_LOW_PASS_12DB = 40,   /* gb_biquad_params {cutoff, q}         settings/src/effects.rs:38-55 */
  GB_FX_HIGH_PASS_12DB = 41,  /* {cutoff, q} */
  GB_FX_BAND_PASS_12DB = 42,  /* {cutoff, bandwidth} */
  GB_FX_BAND_STOP_12DB = 43,  /* {cutoff, bandwidth} */
  GB_FX_ALL_PASS_12DB = 44,   /* {cutoff, q} */
  GB_FX_PEAKING_EQ_12DB = 45, /* {cutoff, db-gain} */
  GB_FX_LOW_SHELF_12DB = 46,  /* {cutoff, db-gain} */
  GB_FX_HIGH_SHELF_12DB = 47, /* {cutoff, db-gain} */
  GB_FX_LOW_PASS_24DB = 48,   /* gb_lowpass24_params {cutoff, passband-ripple} */
  GB_FX_SIGNAL_PASSTHROUGH = 49 /* no params — SignalPassthroughController (settings/src/controllers.rs:110-111,181-187):
                                   sits in a patch chain, passes audio through unchanged, and is the SOURCE of control
                                   links (gb_link_control): the sidechain of projects/demos/controllers/sidechain.json */
} gb_kind;

/* uid of the main mixer every engine starts with (orchestrator.rs:104,543-546) */
#define GB_MAIN_MIXER 1u

/* ---- generator params ---------------------------------------------------- */
typedef enum {
  GB_WAVE_NONE = 0,
  GB_WAVE_SINE = 1,
  GB_WAVE_SQUARE = 2,
  GB_WAVE_PULSE_WIDTH = 3,
  GB_WAVE_TRIANGLE = 4,
  GB_WAVE_SAWTOOTH = 5,
  GB_WAVE_NOISE = 6,
  GB_WAVE_DEBUG_ZERO = 7,
  GB_WAVE_DEBUG_MAX = 8,
  GB_WAVE_DEBUG_MIN = 9
} gb_waveform;              /* settings/src/patches.rs:175-189 */

typedef struct {
  int32_t waveform;         /* gb_waveform */
  int32_t _pad;
  double pulse_width;       /* duty cycle for GB_WAVE_PULSE_WIDTH, 0..1 */
  double frequency;         /* Hz; used by LFOs. 0 = follow the MIDI note */
  double fixed_frequency;   /* Hz; >0 overrides note tracking (osc-2-track=false, patches.rs:93-101) */
  double frequency_tune;    /* Ratio; 2^((100*semis+cents)/1200), patches.rs:255-258 */
} gb_oscillator_params;

typedef struct {
  double attack;            /* seconds */
  double decay;             /* seconds */
  double sustain;           /* Normal 0..1 */
  double release;           /* seconds */
} gb_envelope_params;

typedef struct {
  double gain;              /* Normal 0..1 */
  double pan;               /* BipolarNormal -1..1 */
} gb_dca_params;            /* patches.rs:160-168 */

typedef enum {
  GB_LFO_NONE = 0,
  GB_LFO_AMPLITUDE = 1,
  GB_LFO_PITCH = 2,
  GB_LFO_PULSE_WIDTH = 3,
  GB_LFO_FILTER_CUTOFF = 4
} gb_lfo_routing;           /* patches.rs:269-294 */

/* WelshSynthParams/WelshVoiceParams as reconstructed at settings/src/patches.rs:110-169 */
typedef struct {
  gb_oscillator_params oscillator_1;
  gb_oscillator_params oscillator_2;
  int32_t oscillator_2_sync;
  int32_t lfo_routing;              /* gb_lfo_routing */
  double oscillator_mix;            /* Normal: share of oscillator 1 */
  gb_envelope_params amp_envelope;
  gb_oscillator_params lfo;
  double lfo_depth;                 /* Normal */
  double filter_cutoff_hz;          /* BiQuadFilterLowPass24dbParams.cutoff */
  double filter_passband_ripple;    /* BiQuadFilterLowPass24dbParams.passband_ripple */
  double filter_cutoff_start;       /* Normal (percent of the 25 Hz..20 kHz log range) */
  double filter_cutoff_end;         /* Normal (= filter-envelope-weight) */
  gb_envelope_params filter_envelope;
  gb_dca_params voice_dca;
  gb_dca_params dca;
  uint32_t voices;                  /* polyphony of this instrument (voice store size) */
  uint32_t _pad;
} gb_welsh_params;

typedef struct {
  double ratio;                     /* modulator Hz = ratio * carrier Hz */
  double depth;                     /* Normal */
  double beta;                      /* modulation index */
  gb_envelope_params carrier_envelope;
  gb_envelope_params modulator_envelope;
  gb_dca_params dca;
  uint32_t voices;
  uint32_t _pad;
} gb_fm_params;

typedef struct {
  double root_hz;                   /* frequency at which the sample plays at its native rate */
  uint32_t voices;
  uint32_t _pad;
} gb_sampler_params;                /* sample data arrives through gb_load_sample(key = 0) */

typedef struct {
  uint32_t _reserved;               /* drum sounds arrive through gb_load_sample(key = MIDI key) */
  uint32_t _pad;
} gb_drumkit_params;

typedef struct {
  double level_left;                /* constant DC source (groove-toys ToyAudioSource{level}) */
  double level_right;
} gb_toy_source_params;

typedef struct {
  gb_oscillator_params oscillator;  /* waveform, pulse_width, frequency (Hz, fixed); fixed_frequency / frequency_tune unused */
} gb_oscillator_source_params;

typedef struct {
  gb_envelope_params envelope;
} gb_envelope_source_params;

/* ---- effect params (kebab-case field names of the reference in comments) -- */
typedef struct { double ceiling; } gb_gain_params;                      /* "ceiling" */
typedef struct { double min, max; } gb_limiter_params;                  /* "min"/"max" (older: minimum/maximum) */
typedef struct { double bits; } gb_bitcrusher_params;                   /* "bits" (older: bits-to-crush) */
/* attack / release are accepted (the reference's CompressorParams carries them, projects/default.json5:56-61) and
   do not affect the audio: the compressor is a static threshold/ratio curve per sample (docs/ORACLE_SPEC.md §6). */
typedef struct { double threshold, ratio, attack, release; } gb_compressor_params;
typedef struct { double seconds; } gb_delay_params;                     /* "delay" (older) / "seconds" */
typedef struct { double voices, delay_seconds, wet_dry_mix; } gb_chorus_params;
typedef struct { double attenuation, seconds; } gb_reverb_params;
typedef struct { double cutoff; double param2; } gb_biquad_params;      /* param2 = q | bandwidth | db-gain */
typedef struct { double cutoff; double passband_ripple; } gb_lowpass24_params;

/* ---- control indices for GB_EV_CONTROL (flattened, as proc-macros/src/control.rs:210-226) */
enum {
  GB_CTL_GAIN_CEILING = 0,
  GB_CTL_LIMITER_MIN = 0, GB_CTL_LIMITER_MAX = 1,
  GB_CTL_BITCRUSHER_BITS = 0,
  GB_CTL_COMPRESSOR_THRESHOLD = 0, GB_CTL_COMPRESSOR_RATIO = 1,
  GB_CTL_FILTER_CUTOFF = 0, GB_CTL_FILTER_PARAM2 = 1,
  GB_CTL_CHORUS_WET_DRY_MIX = 2,
  GB_CTL_REVERB_ATTENUATION = 0,
  GB_CTL_INST_DCA_GAIN = 0, GB_CTL_INST_DCA_PAN = 1
};

/* ---- events --------------------------------------------------------------- */
typedef enum {
  GB_EV_NOTE_ON = 1,     /* a = MIDI key, b = velocity */
  GB_EV_NOTE_OFF = 2,    /* a = MIDI key */
  GB_EV_CONTROL = 3,     /* a = control index, value = ControlValue 0..1 (orchestration/src/lib.rs:43-46) */
  GB_EV_SET_PARAM = 4    /* a = control index, value = raw parameter value (no 0..1 mapping) */
} gb_event_type;

typedef struct {
  int64_t frame;         /* absolute frame at which the event takes effect */
  uint32_t uid;          /* target entity */
  uint32_t type;         /* gb_event_type */
  int32_t a;
  int32_t b;
  double value;
} gb_event;

/* ---- engine ---------------------------------------------------------------- */
typedef struct {
  uint32_t abi_version;  /* GB_ABI_VERSION */
  int32_t device;        /* CUDA device ordinal */
  double sample_rate;    /* Hz; SampleRate::DEFAULT = 44100 (src/lib.rs:30) */
  uint32_t max_block;    /* largest `frames` gb_render_block will be called with (0 = default 1<<16) */
  uint32_t flags;        /* reserved, 0 */
} gb_config;

typedef struct gb_engine gb_engine;

int gb_create(const gb_config* cfg, gb_engine** out);
void gb_destroy(gb_engine* e);
const char* gb_last_error(const gb_engine* e);   /* e may be NULL: error of the last failed gb_create */

int gb_add_instrument(gb_engine* e, int32_t kind, const void* params, size_t params_size, uint32_t* uid);
int gb_add_effect(gb_engine* e, int32_t kind, const void* params, size_t params_size, uint32_t* uid);
/* frames: interleaved when channels == 2; values in [-1,1). key: MIDI key for a drumkit, 0 for a sampler.
   Call before gb_finalize (the sample table is part of the frozen plan): GB_ESTATE afterwards. */
int gb_load_sample(gb_engine* e, uint32_t uid, uint8_t key, const double* frames, size_t n_frames,
                   int32_t channels, double sample_rate, double root_hz);
/* source -> destination; destination must be an effect (orchestrator.rs:263-304). */
int gb_patch(gb_engine* e, uint32_t src_uid, uint32_t dst_uid);
/* Freeze the graph: topological plan + device state allocation. */
int gb_finalize(gb_engine* e);
int gb_push_events(gb_engine* e, const gb_event* ev, size_t n);
/* Render `frames` frames from the current position into out_interleaved_lr (host, 2*frames doubles). */
int gb_render_block(gb_engine* e, double* out_interleaved_lr, size_t frames, size_t* frames_done);
/* Same, converted on the device to 16-bit PCM: (x * 32767.0) as i16, truncating, saturating. */
int gb_render_pcm16(gb_engine* e, int16_t* out_interleaved_lr, size_t frames, size_t* frames_done);
/* Render without any host copy (result stays in HBM); used to time the device path alone. */
int gb_render_device(gb_engine* e, size_t frames, size_t* frames_done);
/* Device address (double2[frames], L,R interleaved f64) of the most recent gb_render_device result; for
   in-HBM consumers such as a multi-GPU bus reduction.  Valid until the next render call. */
int gb_last_device_buffer(gb_engine* e, void** device_ptr, size_t* frames);
/* Copy the most recent device-resident render (<= frames of it) to the host. */
int gb_read_last(gb_engine* e, double* out_interleaved_lr, size_t frames);
int64_t gb_position(const gb_engine* e);         /* frames handed out so far */
/* Look-ahead for small caller buffers.  Orchestrator::tick is called with 64-frame buffers
   (orchestrator.rs:1696, audio_panel.rs:69); at that size a GPU call is all launch latency.  With a
   look-ahead of L frames gb_render_block renders L frames in one go whenever its host buffer runs dry and
   serves requests shorter than L from that buffer, so the device sees the same large chunks as an offline
   render.  The caller's side of the contract: every event with a frame below gb_position() + L has been
   pushed before the call (offline renders and the sequencer-driven Orchestrator know their events ahead;
   a live MIDI input accepts L frames of latency: later events apply from the next refill on).  The audio is
   the same as without look-ahead.  0 switches it off; changing it or gb_save_state needs an empty buffer
   (GB_ESTATE otherwise).  Only gb_render_block looks ahead. */
int gb_set_lookahead(gb_engine* e, size_t frames);

/* Control link from an audio-rate source: replaces Orchestrator::link_control_by_name
 * (orchestration/src/orchestrator.rs:207-234) for a SignalPassthroughController source.  Controllers do
 * their work once per caller buffer (handle_work, orchestrator.rs:631-708; 64 frames in groove-cli), so
 * at every absolute frame n > 0 that is a multiple of GB_CONTROL_PERIOD the target's parameter
 * `control_index` is set to the control value min(1, |(l + r) / 2|) of the source's output at frame
 * n - 1.  `source_uid` must be a GB_FX_SIGNAL_PASSTHROUGH node, `target_uid` a gain, limiter or
 * compressor (one link per target).  Call before gb_finalize.  A linked parameter is owned by its link:
 * GB_EV_CONTROL events for it are ignored from the first boundary on. */
#define GB_CONTROL_PERIOD 64
int gb_link_control(gb_engine* e, uint32_t source_uid, uint32_t target_uid, int32_t control_index);
/* Serialise all time-varying state.  Call with buf == NULL to query the size. */
int gb_save_state(gb_engine* e, void* buf, size_t* size);
int gb_restore_state(gb_engine* e, const void* buf, size_t size);

/* ---- measurement hooks (used by bench.py; not part of the render contract) - */
typedef struct {
  uint64_t kernel_launches;      /* kernels this library launched since the last reset */
  uint64_t voice_kernel_launches;
  double voice_kernel_ms;        /* CUDA-event time of the voice kernels since the last reset */
  double fx_kernel_ms;           /* ... of the effect / mix kernels */
  double render_ms;              /* CUDA-event time of whole render calls (first op to last op on the stream) */
  uint64_t voice_samples;        /* voice x frame units the voice kernels covered */
  uint64_t h2d_bytes;
  uint64_t d2h_bytes;
  uint64_t rest_kernel_launches; /* of the voice kernel launches: the resting-voice kernel (time-invariant stretches) */
  double rest_kernel_ms;         /* ... its share of voice_kernel_ms */
  uint64_t rest_voice_samples;   /* ... and of voice_samples */
  uint64_t sweep_kernel_launches; /* likewise for the sweeping-voice kernel (one moving envelope stage per chunk) */
  double sweep_kernel_ms;
  uint64_t sweep_voice_samples;
  uint64_t solo_kernel_launches;  /* the job-list kernel of solo voices (instruments with fewer voices than a CTA has warps) */
  double solo_kernel_ms;
  uint64_t solo_voice_samples;    /* non-idle (voice, sub-chunk) items x sub-chunk frames */
  uint64_t solo_jobs;
  uint64_t solo_class_items[4];   /* items by class: resting / sweeping (knots) / general / sweeping (exact per frame) */
  uint64_t fm_kernel_launches;
  double fm_kernel_ms;
  uint64_t idle_voice_samples;    /* of voice_samples: voices the host knew to be silent (never launched) */
  uint64_t rest_ctas, sweep_ctas; /* CTAs of the resting / sweeping kernel launches, summed */
  uint64_t fx_batched_nodes;      /* effect nodes that shared a launch with others of their kind and graph level */
  uint64_t rest_tp_launches;      /* of rest_kernel_launches: the time-parallel variant (CTAs of <= 8 voices) */
  uint64_t rest_vr_launches;      /* of rest_kernel_launches: all-resting chunks over voice ranges (CTAs across instrument boundaries) */
  uint64_t rest_vr16_launches;    /* of rest_vr_launches: with 16 frames per lane (chunks of a multiple of 512 frames) */
} gb_stats;
int gb_get_stats(gb_engine* e, gb_stats* out);
int gb_reset_stats(gb_engine* e);
int gb_set_timing(gb_engine* e, int32_t enabled); /* per-launch CUDA events on/off (default off) */
/* FP64/FP32 FMA microbenchmark on the engine's device: returns TFLOP/s. */
int gb_measure_fma_peak(gb_engine* e, int32_t fp64, double* tflops);

/* ---- multi-GPU bus exchange ------------------------------------------------
 * One process per GPU: independent tracks render on their own GPU and meet in ONE exchange step,
 * the stereo bus.  Every rank owns an exchange buffer; its CUDA IPC handle travels through whatever the host side
 * uses for rendezvous (torch.distributed in bench.py; a pipe in a Rust host); the root maps the peers' buffers and
 * sums them with one kernel whose loads cross NVLink (P2P) — transfer and mix in the same pass.  `stream` is a
 * cudaStream_t (NULL = the legacy default stream); nothing here synchronises: the caller orders publish ->
 * (its own cross-rank barrier) -> reduce.  Replaces the reference's single-process bus sum
 * (orchestration/src/orchestrator.rs:401-410 walks every track on one CPU thread). */
#define GB_IPC_HANDLE_BYTES 64
typedef struct gb_bus_exchange gb_bus_exchange;
int gb_bus_exchange_create(int32_t device, size_t frames, gb_bus_exchange** out);
void gb_bus_exchange_destroy(gb_bus_exchange* x);
int gb_bus_exchange_export(gb_bus_exchange* x, void* handle /* GB_IPC_HANDLE_BYTES */);
int gb_bus_exchange_open(gb_bus_exchange* x, const void* handles /* n x GB_IPC_HANDLE_BYTES, rank order */, int32_t n, int32_t self);
int gb_bus_exchange_publish(gb_bus_exchange* x, gb_engine* e, size_t frames, void* stream); /* after gb_render_device */
int gb_bus_exchange_reduce(gb_bus_exchange* x, size_t frames, void* stream);                /* root only */
int gb_bus_exchange_result(gb_bus_exchange* x, void** device_ptr, size_t* frames);

#ifdef __cplusplus
}
#endif
#endif /* GROOVE_B200_H */
