//! Safe wrapper over the C ABI and the `BlockRender` adapter that stands behind
//! `Orchestrator::tick` (orchestration/src/orchestrator.rs:856-877): the Orchestrator keeps its
//! entities, patch cables, MIDI routing and controllers; only the per-frame graph walk
//! (`gather_audio`, :367-470) is replaced by one `gb_render_block` call per caller buffer.

use crate::ffi;
use std::collections::HashMap;
use std::ffi::CStr;
use std::fmt;
use std::os::raw::c_void;
use std::ptr;

/// The reference's `StereoSample` is `(Sample(f64), Sample(f64))` (orchestration/src/helpers.rs:78,
/// src/lib.rs:32): two f64, left then right.  This mirror has the same layout, so a
/// `&mut [ensnare::StereoSample]` can be reinterpreted as `&mut [StereoSample]` (or handed over as
/// `*mut f64` of length `2 * frames`) without a copy.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct StereoSample(pub f64, pub f64);

#[derive(Debug, Clone, PartialEq, Eq)]
pub struct Error {
    pub code: i32,
    pub message: String,
}
impl fmt::Display for Error {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        write!(f, "groove_b200 error {}: {}", self.code, self.message)
    }
}
impl std::error::Error for Error {}
pub type Result<T> = std::result::Result<T, Error>;

/// One engine = one Orchestrator-equivalent render graph living on one GPU.
/// Not `Sync`: one thread at a time, as the reference holds its Orchestrator under a Mutex
/// (src/panels/legacy/audio_panel.rs:54-111).
pub struct Engine {
    raw: *mut ffi::gb_engine,
    sample_rate: f64,
}
unsafe impl Send for Engine {}

impl Engine {
    /// `Orchestrator::new_with` (orchestration/src/orchestrator.rs:522-568); the main mixer exists from the start.
    pub fn new(sample_rate: f64, device: i32, max_block: u32) -> Result<Self> {
        let cfg = ffi::gb_config { abi_version: ffi::GB_ABI_VERSION, device, sample_rate, max_block, flags: 0 };
        let mut raw: *mut ffi::gb_engine = ptr::null_mut();
        let rc = unsafe { ffi::gb_create(&cfg, &mut raw) };
        if rc != ffi::GB_OK {
            return Err(Error { code: rc, message: last_error(ptr::null()) });
        }
        Ok(Engine { raw, sample_rate })
    }
    pub fn sample_rate(&self) -> f64 {
        self.sample_rate
    }
    fn check(&self, rc: i32) -> Result<()> {
        if rc == ffi::GB_OK {
            Ok(())
        } else {
            Err(Error { code: rc, message: last_error(self.raw) })
        }
    }
    fn add<T>(&mut self, instrument: bool, kind: i32, params: Option<&T>) -> Result<u32> {
        let mut uid = 0u32;
        let (p, n) = match params {
            Some(p) => (p as *const T as *const c_void, std::mem::size_of::<T>()),
            None => (ptr::null(), 0),
        };
        let rc = unsafe {
            if instrument {
                ffi::gb_add_instrument(self.raw, kind, p, n, &mut uid)
            } else {
                ffi::gb_add_effect(self.raw, kind, p, n, &mut uid)
            }
        };
        self.check(rc).map(|_| uid)
    }
    pub fn add_welsh(&mut self, p: &ffi::gb_welsh_params) -> Result<u32> {
        self.add(true, ffi::GB_INST_WELSH, Some(p))
    }
    pub fn add_fm(&mut self, p: &ffi::gb_fm_params) -> Result<u32> {
        self.add(true, ffi::GB_INST_FM, Some(p))
    }
    pub fn add_sampler(&mut self, p: &ffi::gb_sampler_params) -> Result<u32> {
        self.add(true, ffi::GB_INST_SAMPLER, Some(p))
    }
    pub fn add_drumkit(&mut self) -> Result<u32> {
        self.add(true, ffi::GB_INST_DRUMKIT, Some(&ffi::gb_drumkit_params::default()))
    }
    pub fn add_oscillator(&mut self, p: &ffi::gb_oscillator_source_params) -> Result<u32> {
        self.add(true, ffi::GB_INST_OSCILLATOR, Some(p))
    }
    pub fn add_envelope(&mut self, p: &ffi::gb_envelope_source_params) -> Result<u32> {
        self.add(true, ffi::GB_INST_ENVELOPE, Some(p))
    }
    /// Any effect kind with its params struct (`GB_FX_*`); `None` for mixer / signal-passthrough.
    pub fn add_effect<T>(&mut self, kind: i32, params: Option<&T>) -> Result<u32> {
        self.add(false, kind, params)
    }
    /// Decoded sample data for a sampler (`key` = 0) or one drum of a drumkit (`key` = MIDI key).
    pub fn load_sample(&mut self, uid: u32, key: u8, frames: &[f64], channels: i32, sample_rate: f64, root_hz: f64) -> Result<()> {
        let n = frames.len() / channels.max(1) as usize;
        let rc = unsafe { ffi::gb_load_sample(self.raw, uid, key, frames.as_ptr(), n, channels, sample_rate, root_hz) };
        self.check(rc)
    }
    /// `Orchestrator::patch(output_uid, input_uid)` (:263-304): the input device must be an effect.
    pub fn patch(&mut self, src: u32, dst: u32) -> Result<()> {
        let rc = unsafe { ffi::gb_patch(self.raw, src, dst) };
        self.check(rc)
    }
    /// `Orchestrator::link_control_by_name` for a SignalPassthroughController source (:207-234).
    pub fn link_control(&mut self, source: u32, target: u32, control_index: i32) -> Result<()> {
        let rc = unsafe { ffi::gb_link_control(self.raw, source, target, control_index) };
        self.check(rc)
    }
    pub fn finalize(&mut self) -> Result<()> {
        let rc = unsafe { ffi::gb_finalize(self.raw) };
        self.check(rc)
    }
    pub fn push_events(&mut self, events: &[ffi::gb_event]) -> Result<()> {
        let rc = unsafe { ffi::gb_push_events(self.raw, events.as_ptr(), events.len()) };
        self.check(rc)
    }
    /// `gather_audio(&mut [StereoSample])`: renders `out.len()` frames from the current position.
    pub fn render(&mut self, out: &mut [StereoSample]) -> Result<usize> {
        let mut done = 0usize;
        let rc = unsafe { ffi::gb_render_block(self.raw, out.as_mut_ptr() as *mut f64, out.len(), &mut done) };
        self.check(rc).map(|_| done)
    }
    /// `IOHelper::send_performance_to_file`'s conversion (orchestration/src/helpers.rs:74-97), done on the device.
    pub fn render_pcm16(&mut self, out_interleaved_lr: &mut [i16]) -> Result<usize> {
        let mut done = 0usize;
        let frames = out_interleaved_lr.len() / 2;
        let rc = unsafe { ffi::gb_render_pcm16(self.raw, out_interleaved_lr.as_mut_ptr(), frames, &mut done) };
        self.check(rc).map(|_| done)
    }
    /// Serve the Orchestrator's 64-frame `tick` buffers (orchestrator.rs:1696) from one big device render
    /// every `frames` frames; the sequencer knows its events that far ahead (include/groove_b200.h).
    pub fn set_lookahead(&mut self, frames: usize) -> Result<()> {
        let rc = unsafe { ffi::gb_set_lookahead(self.raw, frames) };
        self.check(rc)
    }
    pub fn position(&self) -> i64 {
        unsafe { ffi::gb_position(self.raw) }
    }
    pub fn save_state(&mut self) -> Result<Vec<u8>> {
        let mut size = 0usize;
        let rc = unsafe { ffi::gb_save_state(self.raw, ptr::null_mut(), &mut size) };
        self.check(rc)?;
        let mut buf = vec![0u8; size];
        let rc = unsafe { ffi::gb_save_state(self.raw, buf.as_mut_ptr() as *mut c_void, &mut size) };
        self.check(rc)?;
        buf.truncate(size);
        Ok(buf)
    }
    pub fn restore_state(&mut self, blob: &[u8]) -> Result<()> {
        let rc = unsafe { ffi::gb_restore_state(self.raw, blob.as_ptr() as *const c_void, blob.len()) };
        self.check(rc)
    }
    pub fn stats(&mut self) -> Result<ffi::gb_stats> {
        let mut st = ffi::gb_stats::default();
        let rc = unsafe { ffi::gb_get_stats(self.raw, &mut st) };
        self.check(rc).map(|_| st)
    }
}
impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { ffi::gb_destroy(self.raw) }
    }
}

fn last_error(raw: *const ffi::gb_engine) -> String {
    let p = unsafe { ffi::gb_last_error(raw) };
    if p.is_null() {
        String::new()
    } else {
        unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned()
    }
}

/// What the Orchestrator hands over when its composition changes: plain data it already has.
/// `#[derive(Params)]` structs of the reference are plain data (proc-macros/src/params.rs:14-151), so
/// filling these is a field-by-field copy; `settings/src/patches.rs:87-170` produces the Welsh ones.
pub enum EntitySpec {
    Welsh(ffi::gb_welsh_params),
    Fm(ffi::gb_fm_params),
    Sampler { params: ffi::gb_sampler_params, frames: Vec<f64>, channels: i32, sample_rate: f64, root_hz: f64 },
    Drumkit { drums: Vec<(u8, Vec<f64>, i32, f64)> },
    Oscillator(ffi::gb_oscillator_source_params),
    Envelope(ffi::gb_envelope_source_params),
    Mixer,
    SignalPassthrough,
    Gain(ffi::gb_gain_params),
    Limiter(ffi::gb_limiter_params),
    Bitcrusher(ffi::gb_bitcrusher_params),
    Compressor(ffi::gb_compressor_params),
    Delay(ffi::gb_delay_params),
    Chorus(ffi::gb_chorus_params),
    Reverb(ffi::gb_reverb_params),
    /// kind = one of `GB_FX_LOW_PASS_12DB ..= GB_FX_HIGH_SHELF_12DB`
    Biquad { kind: i32, params: ffi::gb_biquad_params },
    LowPass24(ffi::gb_lowpass24_params),
}

/// The reference's entity events, reduced to what reaches the render path
/// (`EntityEvent::Midi` / `EntityEvent::Control`: orchestration/src/orchestrator.rs:710-754).
pub enum EntityEvent {
    NoteOn { key: u8, velocity: u8 },
    NoteOff { key: u8 },
    /// `Controllable::control_set_param_by_index(index, ControlValue)`: value in 0..=1
    Control { index: i32, value: f64 },
}

/// The block-render entry point behind `Orchestrator::tick`.
pub struct BlockRender {
    engine: Engine,
    /// the Orchestrator's `Uid` (usize) -> engine uid
    uids: HashMap<usize, u32>,
    pending: Vec<ffi::gb_event>,
}

impl BlockRender {
    /// Mirror the Orchestrator's store, patch cables and control links once.
    /// `main_mixer_uid` is the Orchestrator's uid of `main-mixer` (:104,543-546); it maps to `GB_MAIN_MIXER`.
    pub fn snapshot(
        sample_rate: f64,
        device: i32,
        main_mixer_uid: usize,
        entities: &[(usize, EntitySpec)],
        patch_cables: &[(usize, usize)],
        passthrough_links: &[(usize, usize, i32)],
    ) -> Result<Self> {
        let mut engine = Engine::new(sample_rate, device, 0)?;
        let mut uids = HashMap::new();
        uids.insert(main_mixer_uid, ffi::GB_MAIN_MIXER);
        for (uid, spec) in entities {
            let id = match spec {
                EntitySpec::Welsh(p) => engine.add_welsh(p)?,
                EntitySpec::Fm(p) => engine.add_fm(p)?,
                EntitySpec::Sampler { params, frames, channels, sample_rate, root_hz } => {
                    let id = engine.add_sampler(params)?;
                    engine.load_sample(id, 0, frames, *channels, *sample_rate, *root_hz)?;
                    id
                }
                EntitySpec::Drumkit { drums } => {
                    let id = engine.add_drumkit()?;
                    for (key, frames, channels, sr) in drums {
                        engine.load_sample(id, *key, frames, *channels, *sr, 0.0)?;
                    }
                    id
                }
                EntitySpec::Oscillator(p) => engine.add_oscillator(p)?,
                EntitySpec::Envelope(p) => engine.add_envelope(p)?,
                EntitySpec::Mixer => engine.add_effect::<ffi::gb_gain_params>(ffi::GB_FX_MIXER, None)?,
                EntitySpec::SignalPassthrough => engine.add_effect::<ffi::gb_gain_params>(ffi::GB_FX_SIGNAL_PASSTHROUGH, None)?,
                EntitySpec::Gain(p) => engine.add_effect(ffi::GB_FX_GAIN, Some(p))?,
                EntitySpec::Limiter(p) => engine.add_effect(ffi::GB_FX_LIMITER, Some(p))?,
                EntitySpec::Bitcrusher(p) => engine.add_effect(ffi::GB_FX_BITCRUSHER, Some(p))?,
                EntitySpec::Compressor(p) => engine.add_effect(ffi::GB_FX_COMPRESSOR, Some(p))?,
                EntitySpec::Delay(p) => engine.add_effect(ffi::GB_FX_DELAY, Some(p))?,
                EntitySpec::Chorus(p) => engine.add_effect(ffi::GB_FX_CHORUS, Some(p))?,
                EntitySpec::Reverb(p) => engine.add_effect(ffi::GB_FX_REVERB, Some(p))?,
                EntitySpec::Biquad { kind, params } => engine.add_effect(*kind, Some(params))?,
                EntitySpec::LowPass24(p) => engine.add_effect(ffi::GB_FX_LOW_PASS_24DB, Some(p))?,
            };
            uids.insert(*uid, id);
        }
        let lookup = |uids: &HashMap<usize, u32>, u: usize| {
            uids.get(&u).copied().ok_or(Error { code: ffi::GB_ENOENT, message: format!("unknown entity uid {u}") })
        };
        for (src, dst) in patch_cables {
            engine.patch(lookup(&uids, *src)?, lookup(&uids, *dst)?)?;
        }
        for (src, dst, index) in passthrough_links {
            engine.link_control(lookup(&uids, *src)?, lookup(&uids, *dst)?, *index)?;
        }
        engine.finalize()?;
        Ok(BlockRender { engine, uids, pending: Vec::new() })
    }

    /// Called from `handle_work`'s callbacks instead of `entity.handle_midi_message` /
    /// `control_set_param_by_index`; `frame` is the first frame of the caller buffer being processed
    /// (the reference runs controllers once per buffer: :631-708).
    pub fn push(&mut self, frame: usize, target: usize, ev: EntityEvent) {
        let Some(&uid) = self.uids.get(&target) else { return };  // unknown target: the reference panics (:741-744)
        let (type_, a, b, value) = match ev {
            EntityEvent::NoteOn { key, velocity } => (ffi::GB_EV_NOTE_ON, key as i32, velocity as i32, 0.0),
            EntityEvent::NoteOff { key } => (ffi::GB_EV_NOTE_OFF, key as i32, 0, 0.0),
            EntityEvent::Control { index, value } => (ffi::GB_EV_CONTROL, index, 0, value),
        };
        self.pending.push(ffi::gb_event { frame: frame as i64, uid, type_, a, b, value });
    }

    /// Replaces `gather_audio` (:367-470): returns the frames rendered (always `samples.len()`).
    pub fn gather_audio(&mut self, samples: &mut [StereoSample]) -> Result<usize> {
        if !self.pending.is_empty() {
            self.engine.push_events(&self.pending)?;
            self.pending.clear();
        }
        self.engine.render(samples)
    }

    pub fn engine(&mut self) -> &mut Engine {
        &mut self.engine
    }
}

/// One rank's side of the multi-GPU bus mixdown (include/groove_b200.h, "multi-GPU bus exchange"): tracks are
/// independent until the main mixer (orchestration/src/orchestrator.rs:401-410), so a multi-GPU host runs one
/// `BlockRender` per GPU over a share of the tracks and the stereo buses meet once per render — the root GPU sums
/// them with one kernel whose loads of the other GPUs' buffers cross NVLink.  The host carries the 64-byte
/// handles between the processes over whatever channel it has.
pub struct BusExchange {
    raw: *mut ffi::gb_bus_exchange,
}
unsafe impl Send for BusExchange {}

impl BusExchange {
    pub fn new(device: i32, max_frames: usize) -> Result<Self> {
        let mut raw: *mut ffi::gb_bus_exchange = ptr::null_mut();
        let rc = unsafe { ffi::gb_bus_exchange_create(device, max_frames, &mut raw) };
        if rc != ffi::GB_OK {
            return Err(Error { code: rc, message: last_error(ptr::null()) });
        }
        Ok(BusExchange { raw })
    }
    fn check(&self, rc: i32) -> Result<()> {
        if rc == ffi::GB_OK { Ok(()) } else { Err(Error { code: rc, message: last_error(ptr::null()) }) }
    }
    /// This rank's handle, to be gathered from all ranks in rank order.
    pub fn export(&self) -> Result<[u8; ffi::GB_IPC_HANDLE_BYTES]> {
        let mut h = [0u8; ffi::GB_IPC_HANDLE_BYTES];
        self.check(unsafe { ffi::gb_bus_exchange_export(self.raw, h.as_mut_ptr() as *mut c_void) })?;
        Ok(h)
    }
    /// Root only: map the other ranks' buffers (`handles` = n x 64 bytes in rank order).
    pub fn open(&mut self, handles: &[u8], self_rank: i32) -> Result<()> {
        let n = (handles.len() / ffi::GB_IPC_HANDLE_BYTES) as i32;
        self.check(unsafe { ffi::gb_bus_exchange_open(self.raw, handles.as_ptr() as *const c_void, n, self_rank) })
    }
    /// After `Engine::render_device`: copy the rank's bus into its exchange buffer (enqueued on `stream`).
    pub fn publish(&mut self, engine: &mut Engine, frames: usize, stream: *mut c_void) -> Result<()> {
        self.check(unsafe { ffi::gb_bus_exchange_publish(self.raw, engine.raw, frames, stream) })
    }
    /// Root only, after every rank has published (the host's own barrier): sum the buses; returns the device
    /// pointer of the mixed `[(f64, f64); frames]`.
    pub fn reduce(&mut self, frames: usize, stream: *mut c_void) -> Result<*mut c_void> {
        self.check(unsafe { ffi::gb_bus_exchange_reduce(self.raw, frames, stream) })?;
        let (mut p, mut n) = (ptr::null_mut(), 0usize);
        self.check(unsafe { ffi::gb_bus_exchange_result(self.raw, &mut p, &mut n) })?;
        Ok(p)
    }
}
impl Drop for BusExchange {
    fn drop(&mut self) {
        unsafe { ffi::gb_bus_exchange_destroy(self.raw) }
    }
}
