//! Raw bindings of `include/groove_b200.h` (ABI version 4) and a safe wrapper.
//!
//! Each entry point replaces one interface of the reference (paths relative to sowbug/groove):
//! `gb_render_block` = `Orchestrator::tick`'s `gather_audio` (orchestration/src/orchestrator.rs:367-470,856-877),
//! `gb_patch` = `Orchestrator::patch` (:263-304), `gb_push_events` = `handle_midi_message` /
//! `control_set_param_by_index` fan-out (:710-754; proc-macros/src/control.rs:237-249), and so on —
//! the full table is in INTEGRATION.md §1.
#![allow(non_camel_case_types)]

pub mod block_render;
pub use block_render::{BlockRender, Engine, Error, StereoSample};

pub mod ffi {
    use std::os::raw::{c_char, c_int, c_void};

    pub const GB_ABI_VERSION: u32 = 4;
    pub const GB_MAIN_MIXER: u32 = 1;
    pub const GB_CONTROL_PERIOD: i64 = 64;

    // error codes
    pub const GB_OK: c_int = 0;
    pub const GB_EINVAL: c_int = -1;
    pub const GB_ENOENT: c_int = -2;
    pub const GB_ESTATE: c_int = -3;
    pub const GB_ENODEV: c_int = -4;
    pub const GB_ECUDA: c_int = -5;
    pub const GB_ENOMEM: c_int = -6;
    pub const GB_EGRAPH: c_int = -7;

    // entity kinds (gb_kind)
    pub const GB_INST_WELSH: i32 = 1;
    pub const GB_INST_FM: i32 = 2;
    pub const GB_INST_SAMPLER: i32 = 3;
    pub const GB_INST_DRUMKIT: i32 = 4;
    pub const GB_INST_TOY_SOURCE: i32 = 5;
    pub const GB_INST_OSCILLATOR: i32 = 6;
    pub const GB_INST_ENVELOPE: i32 = 7;
    pub const GB_FX_MIXER: i32 = 32;
    pub const GB_FX_GAIN: i32 = 33;
    pub const GB_FX_LIMITER: i32 = 34;
    pub const GB_FX_BITCRUSHER: i32 = 35;
    pub const GB_FX_COMPRESSOR: i32 = 36;
    pub const GB_FX_DELAY: i32 = 37;
    pub const GB_FX_CHORUS: i32 = 38;
    pub const GB_FX_REVERB: i32 = 39;
    pub const GB_FX_LOW_PASS_12DB: i32 = 40;
    pub const GB_FX_HIGH_PASS_12DB: i32 = 41;
    pub const GB_FX_BAND_PASS_12DB: i32 = 42;
    pub const GB_FX_BAND_STOP_12DB: i32 = 43;
    pub const GB_FX_ALL_PASS_12DB: i32 = 44;
    pub const GB_FX_PEAKING_EQ_12DB: i32 = 45;
    pub const GB_FX_LOW_SHELF_12DB: i32 = 46;
    pub const GB_FX_HIGH_SHELF_12DB: i32 = 47;
    pub const GB_FX_LOW_PASS_24DB: i32 = 48;
    pub const GB_FX_SIGNAL_PASSTHROUGH: i32 = 49;

    // waveforms (gb_waveform), LFO routings (gb_lfo_routing), event types (gb_event_type)
    pub const GB_WAVE_NONE: i32 = 0;
    pub const GB_WAVE_SINE: i32 = 1;
    pub const GB_WAVE_SQUARE: i32 = 2;
    pub const GB_WAVE_PULSE_WIDTH: i32 = 3;
    pub const GB_WAVE_TRIANGLE: i32 = 4;
    pub const GB_WAVE_SAWTOOTH: i32 = 5;
    pub const GB_WAVE_NOISE: i32 = 6;
    pub const GB_WAVE_DEBUG_ZERO: i32 = 7;
    pub const GB_WAVE_DEBUG_MAX: i32 = 8;
    pub const GB_WAVE_DEBUG_MIN: i32 = 9;
    pub const GB_LFO_NONE: i32 = 0;
    pub const GB_LFO_AMPLITUDE: i32 = 1;
    pub const GB_LFO_PITCH: i32 = 2;
    pub const GB_LFO_PULSE_WIDTH: i32 = 3;
    pub const GB_LFO_FILTER_CUTOFF: i32 = 4;
    pub const GB_EV_NOTE_ON: u32 = 1;
    pub const GB_EV_NOTE_OFF: u32 = 2;
    pub const GB_EV_CONTROL: u32 = 3;
    pub const GB_EV_SET_PARAM: u32 = 4;

    #[repr(C)]
    pub struct gb_engine {
        _private: [u8; 0],
    }

    /// Opaque: one rank's side of the multi-GPU bus exchange (`GB_IPC_HANDLE_BYTES`-byte handles name the buffers).
    #[repr(C)]
    pub struct gb_bus_exchange {
        _private: [u8; 0],
    }
    pub const GB_IPC_HANDLE_BYTES: usize = 64;

    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct gb_oscillator_params {
        pub waveform: i32,
        pub _pad: i32,
        pub pulse_width: f64,
        pub frequency: f64,
        pub fixed_frequency: f64,
        pub frequency_tune: f64,
    }
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct gb_envelope_params {
        pub attack: f64,
        pub decay: f64,
        pub sustain: f64,
        pub release: f64,
    }
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct gb_dca_params {
        pub gain: f64,
        pub pan: f64,
    }
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct gb_welsh_params {
        pub oscillator_1: gb_oscillator_params,
        pub oscillator_2: gb_oscillator_params,
        pub oscillator_2_sync: i32,
        pub lfo_routing: i32,
        pub oscillator_mix: f64,
        pub amp_envelope: gb_envelope_params,
        pub lfo: gb_oscillator_params,
        pub lfo_depth: f64,
        pub filter_cutoff_hz: f64,
        pub filter_passband_ripple: f64,
        pub filter_cutoff_start: f64,
        pub filter_cutoff_end: f64,
        pub filter_envelope: gb_envelope_params,
        pub voice_dca: gb_dca_params,
        pub dca: gb_dca_params,
        pub voices: u32,
        pub _pad: u32,
    }
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct gb_fm_params {
        pub ratio: f64,
        pub depth: f64,
        pub beta: f64,
        pub carrier_envelope: gb_envelope_params,
        pub modulator_envelope: gb_envelope_params,
        pub dca: gb_dca_params,
        pub voices: u32,
        pub _pad: u32,
    }
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct gb_sampler_params {
        pub root_hz: f64,
        pub voices: u32,
        pub _pad: u32,
    }
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct gb_drumkit_params {
        pub _reserved: u32,
        pub _pad: u32,
    }
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct gb_toy_source_params {
        pub level_left: f64,
        pub level_right: f64,
    }
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct gb_oscillator_source_params {
        pub oscillator: gb_oscillator_params,
    }
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct gb_envelope_source_params {
        pub envelope: gb_envelope_params,
    }
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct gb_gain_params {
        pub ceiling: f64,
    }
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct gb_limiter_params {
        pub min: f64,
        pub max: f64,
    }
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct gb_bitcrusher_params {
        pub bits: f64,
    }
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct gb_compressor_params {
        pub threshold: f64,
        pub ratio: f64,
        pub attack: f64,
        pub release: f64,
    }
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct gb_delay_params {
        pub seconds: f64,
    }
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct gb_chorus_params {
        pub voices: f64,
        pub delay_seconds: f64,
        pub wet_dry_mix: f64,
    }
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct gb_reverb_params {
        pub attenuation: f64,
        pub seconds: f64,
    }
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct gb_biquad_params {
        pub cutoff: f64,
        pub param2: f64,
    }
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct gb_lowpass24_params {
        pub cutoff: f64,
        pub passband_ripple: f64,
    }
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct gb_event {
        pub frame: i64,
        pub uid: u32,
        pub type_: u32,
        pub a: i32,
        pub b: i32,
        pub value: f64,
    }
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct gb_config {
        pub abi_version: u32,
        pub device: i32,
        pub sample_rate: f64,
        pub max_block: u32,
        pub flags: u32,
    }
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct gb_stats {
        pub kernel_launches: u64,
        pub voice_kernel_launches: u64,
        pub voice_kernel_ms: f64,
        pub fx_kernel_ms: f64,
        pub render_ms: f64,
        pub voice_samples: u64,
        pub h2d_bytes: u64,
        pub d2h_bytes: u64,
        pub rest_kernel_launches: u64,
        pub rest_kernel_ms: f64,
        pub rest_voice_samples: u64,
        pub sweep_kernel_launches: u64,
        pub sweep_kernel_ms: f64,
        pub sweep_voice_samples: u64,
        pub solo_kernel_launches: u64,
        pub solo_kernel_ms: f64,
        pub solo_voice_samples: u64,
        pub solo_jobs: u64,
        pub solo_class_items: [u64; 4],
        pub fm_kernel_launches: u64,
        pub fm_kernel_ms: f64,
        pub idle_voice_samples: u64,
        pub rest_ctas: u64,
        pub sweep_ctas: u64,
        pub fx_batched_nodes: u64,
        pub rest_tp_launches: u64,
        pub rest_vr_launches: u64,
        pub rest_vr16_launches: u64,
    }

    extern "C" {
        pub fn gb_create(cfg: *const gb_config, out: *mut *mut gb_engine) -> c_int;
        pub fn gb_destroy(e: *mut gb_engine);
        pub fn gb_last_error(e: *const gb_engine) -> *const c_char;
        pub fn gb_add_instrument(e: *mut gb_engine, kind: i32, params: *const c_void, params_size: usize, uid: *mut u32) -> c_int;
        pub fn gb_add_effect(e: *mut gb_engine, kind: i32, params: *const c_void, params_size: usize, uid: *mut u32) -> c_int;
        pub fn gb_load_sample(e: *mut gb_engine, uid: u32, key: u8, frames: *const f64, n_frames: usize, channels: i32,
                              sample_rate: f64, root_hz: f64) -> c_int;
        pub fn gb_patch(e: *mut gb_engine, src_uid: u32, dst_uid: u32) -> c_int;
        pub fn gb_finalize(e: *mut gb_engine) -> c_int;
        pub fn gb_push_events(e: *mut gb_engine, ev: *const gb_event, n: usize) -> c_int;
        pub fn gb_render_block(e: *mut gb_engine, out_interleaved_lr: *mut f64, frames: usize, frames_done: *mut usize) -> c_int;
        pub fn gb_render_pcm16(e: *mut gb_engine, out_interleaved_lr: *mut i16, frames: usize, frames_done: *mut usize) -> c_int;
        pub fn gb_render_device(e: *mut gb_engine, frames: usize, frames_done: *mut usize) -> c_int;
        pub fn gb_last_device_buffer(e: *mut gb_engine, device_ptr: *mut *mut c_void, frames: *mut usize) -> c_int;
        pub fn gb_read_last(e: *mut gb_engine, out_interleaved_lr: *mut f64, frames: usize) -> c_int;
        pub fn gb_position(e: *const gb_engine) -> i64;
        pub fn gb_set_lookahead(e: *mut gb_engine, frames: usize) -> c_int;
        pub fn gb_link_control(e: *mut gb_engine, source_uid: u32, target_uid: u32, control_index: i32) -> c_int;
        pub fn gb_save_state(e: *mut gb_engine, buf: *mut c_void, size: *mut usize) -> c_int;
        pub fn gb_restore_state(e: *mut gb_engine, buf: *const c_void, size: usize) -> c_int;
        pub fn gb_get_stats(e: *mut gb_engine, out: *mut gb_stats) -> c_int;
        pub fn gb_reset_stats(e: *mut gb_engine) -> c_int;
        pub fn gb_set_timing(e: *mut gb_engine, enabled: i32) -> c_int;
        pub fn gb_measure_fma_peak(e: *mut gb_engine, fp64: i32, tflops: *mut f64) -> c_int;
        // multi-GPU bus exchange: one process per GPU, the root sums the ranks' CUDA IPC buffers over NVLink
        pub fn gb_bus_exchange_create(device: i32, frames: usize, out: *mut *mut gb_bus_exchange) -> c_int;
        pub fn gb_bus_exchange_destroy(x: *mut gb_bus_exchange);
        pub fn gb_bus_exchange_export(x: *mut gb_bus_exchange, handle: *mut c_void) -> c_int;
        pub fn gb_bus_exchange_open(x: *mut gb_bus_exchange, handles: *const c_void, n: i32, self_rank: i32) -> c_int;
        pub fn gb_bus_exchange_publish(x: *mut gb_bus_exchange, e: *mut gb_engine, frames: usize, stream: *mut c_void) -> c_int;
        pub fn gb_bus_exchange_reduce(x: *mut gb_bus_exchange, frames: usize, stream: *mut c_void) -> c_int;
        pub fn gb_bus_exchange_result(x: *mut gb_bus_exchange, device_ptr: *mut *mut c_void, frames: *mut usize) -> c_int;
    }
}
