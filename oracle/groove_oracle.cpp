// groove_oracle.cpp — CPU oracle for the Groove synthesis + effects hot path.
//
// *** TEST INFRASTRUCTURE ONLY.  Nothing in the product path (groove_b200/) may
// *** import, link or call this file.  Only tests/, __graft_entry__.smoke() and
// *** bench.py's cpu_baseline / --impl reference legs use it, as the checker
// *** and as the reported CPU baseline.
//
// PARITY STATUS: **parity unpinned** for the DSP bodies.  The reference
// snapshot does not contain the source of its oscillators, envelopes, filters,
// FM/sampler voices or time-based effects (they live in un-vendored path crates
// `../ensnare/*` / `groove-core` / `groove-entities`, no pinned version:
// Cargo.toml:26-30, entities/src/effects/mod.rs:8-11,
// entities/src/instruments/mod.rs:9-12) and there is no Rust toolchain here, so
// the reference cannot be run.  What IS pinned by surviving code/tests, and is
// restated and tested here:
//   * graph walk / mix / gain / chain / branch semantics  orchestration/src/orchestrator.rs:367-470,1444-1668
//   * render length ceil(beats*60/bpm*SR)                 orchestration/src/orchestrator.rs:1723-1737
//   * tuning ratio 2^((100*semis+cents)/1200)             settings/src/patches.rs:255-258,754-796
//   * patch -> voice parameter mapping (release:=decay …) settings/src/patches.rs:87-170
//   * PCM16 = trunc(x*32767), saturating, interleaved     orchestration/src/helpers.rs:78-91
//   * biquad math: RBJ cookbook, a0-normalised            doc/Audio-EQ-Cookbook.txt:35-43,74-198
//   * 24 dB low-pass = two cascaded 2nd-order sections    doc/filters004.txt:27,228-262
// Everything else follows docs/ORACLE_SPEC.md (our own written contract).
//
// Structure mirrors the reference: instruments expose tick(1)+value()
// (entities/src/instruments/metronome.rs:23-61), effects expose
// transform_audio(StereoSample) (orchestrator.rs:446-454), and render() walks
// the patch graph depth-first once per frame exactly like gather_audio
// (orchestrator.rs:367-470).  Single-threaded, f64, one frame at a time.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared (see oracle/Makefile).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <climits>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../include/groove_b200.h"

namespace {

constexpr double kTwo64 = 18446744073709551616.0;
constexpr double kTwoPi = 6.283185307179586476925286766559;
constexpr double kPi = 3.141592653589793238462643383279;
constexpr int64_t kHeld = INT64_C(1) << 60;    // "note still held"
constexpr int64_t kNever = -(INT64_C(1) << 60); // "never triggered"

struct Stereo {
  double l = 0.0, r = 0.0;
};

// ---------------------------------------------------------------- helpers ---
// Fractional cycles -> 64-bit fixed-point phase increment (2^64 == one cycle).
inline uint64_t cycles_to_q(double c) {
  c -= std::floor(c);
  double r = c * kTwo64;
  if (!(r < kTwo64)) return 0;
  return (uint64_t)r;
}
inline double pos_of(uint64_t q) { return (double)(q >> 12) * (1.0 / 4503599627370496.0); }  // top 52 bits
inline uint64_t splitmix64(uint64_t x) {
  x += UINT64_C(0x9E3779B97F4A7C15);
  x = (x ^ (x >> 30)) * UINT64_C(0xBF58476D1CE4E5B9);
  x = (x ^ (x >> 27)) * UINT64_C(0x94D049BB133111EB);
  return x ^ (x >> 31);
}
inline int64_t frames_of(double seconds, double sr) {
  if (!(seconds > 0.0)) return 0;
  return (int64_t)std::llround(seconds * sr);
}
inline double note_hz(int key) { return 440.0 * std::exp2(((double)key - 69.0) / 12.0); }
// FrequencyHz <-> percent: 25 Hz * 800^pct (25 Hz .. 20 kHz), patches.rs:150-152 names the mapping.
inline double pct_to_hz(double pct) { return 25.0 * std::exp2(pct * 9.6438561897747243); }  // log2(800)
inline double clamp01(double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); }
inline double denormalize_q(double v) { return v * v * 10.0 + 0.707; }

// Pan law (DCA): left = 1 - ((pan+1)/2)^2, right = 1 - ((pan-1)/2)^2.
inline void dca_gains(double gain, double pan, double* gl, double* gr) {
  double a = 0.5 * (pan + 1.0), b = 0.5 * (pan - 1.0);
  *gl = gain * (1.0 - a * a);
  *gr = gain * (1.0 - b * b);
}

// ------------------------------------------------------------- oscillator ---
// Naive ("pure algorithm", README.md:113-116) waveforms over a 64-bit
// fixed-point phase accumulator.
struct Oscillator {
  int waveform = GB_WAVE_SINE;
  uint64_t duty_q = UINT64_C(1) << 63;
  uint64_t seed = 0;
  uint64_t phase = 0;
  // Advance one frame.  Returns true when the cycle wrapped (for hard sync).
  bool tick(uint64_t dq, bool reset) {
    if (reset) {
      phase = 0;
      return false;
    }
    uint64_t old = phase;
    phase += dq;
    return phase < old;
  }
  double value(int64_t frame, uint64_t duty_override, bool use_override) const {
    const uint64_t half = UINT64_C(1) << 63;
    double p = pos_of(phase);
    switch (waveform) {
      case GB_WAVE_SINE: return std::sin(p * kTwoPi);
      case GB_WAVE_SQUARE: return phase < half ? 1.0 : -1.0;
      case GB_WAVE_PULSE_WIDTH: return phase < (use_override ? duty_override : duty_q) ? 1.0 : -1.0;
      case GB_WAVE_TRIANGLE: return phase < half ? 4.0 * p - 1.0 : 3.0 - 4.0 * p;
      case GB_WAVE_SAWTOOTH: return phase < half ? 2.0 * p : 2.0 * p - 2.0;
      case GB_WAVE_NOISE: return (double)(splitmix64(seed + (uint64_t)frame) >> 11) * (1.0 / 4503599627370496.0) - 1.0;
      case GB_WAVE_DEBUG_MAX: return 1.0;
      case GB_WAVE_DEBUG_MIN: return -1.0;
      default: return 0.0;
    }
  }
};

// --------------------------------------------------------------- envelope ---
// ADSR over integer frame counts.  Attack rises L_on -> 1 over Na frames with
// progress shape t(2-t); decay falls 1 -> S over Nd frames with (1-t)^2;
// release falls L_off -> 0 over Nr frames with (1-t)^2.
struct EnvShape {
  int64_t na = 0, nd = 0, nr = 0;
  double sustain = 1.0;
  void set(const gb_envelope_params& p, double sr) {
    na = frames_of(p.attack, sr);
    nd = frames_of(p.decay, sr);
    nr = frames_of(p.release, sr);
    sustain = clamp01(p.sustain);
  }
};
struct EnvState {
  double l_on = 0.0, l_off = 0.0;
};
inline double env_pre(const EnvShape& s, int64_t n_on, double l_on, int64_t n) {
  int64_t k = n - n_on;
  if (k < s.na) {
    double t = (double)k / (double)s.na;
    return l_on + (1.0 - l_on) * (t * (2.0 - t));
  }
  int64_t k2 = k - s.na;
  if (k2 < s.nd) {
    double u = 1.0 - (double)k2 / (double)s.nd;
    return s.sustain + (1.0 - s.sustain) * (u * u);
  }
  return s.sustain;
}
inline double env_level(const EnvShape& s, int64_t n_on, int64_t n_off, const EnvState& e, int64_t n) {
  if (n < n_on) return 0.0;
  if (n < n_off) return env_pre(s, n_on, e.l_on, n);
  int64_t k = n - n_off;
  if (k < s.nr) {
    double u = 1.0 - (double)k / (double)s.nr;
    return e.l_off * (u * u);
  }
  return 0.0;
}

// ------------------------------------------------- 24 dB low-pass (2 sections) ---
struct Lp24Coef {
  double b0[2], b1[2], b2[2], a1[2], a2[2];
};
struct Lp24Ripple {  // terms that depend on passband_ripple only
  double c0, c2, c1k, c3k;
  void set(double ripple) {
    double sg = std::sinh(ripple);
    double cg = std::cosh(ripple);
    cg *= cg;
    c0 = 1.0 / (cg - 0.85355339059327376220);
    c2 = 1.0 / (cg - 0.14644660940672623780);
    c1k = c0 * sg * 1.84775906502257351225;
    c3k = c2 * sg * 0.76536686473017954345;
  }
};
inline void lp24_coefficients(const Lp24Ripple& rp, double cutoff_hz, double sr, Lp24Coef* c) {
  double fc = cutoff_hz;
  if (fc > 0.49 * sr) fc = 0.49 * sr;
  if (fc < 1.0) fc = 1.0;
  double k = std::tan(kPi * fc / sr);
  double kk = k * k;
  double c1 = k * rp.c1k, c3 = k * rp.c3k;
  double a0 = 1.0 / (c1 + kk + rp.c0);
  c->a1[0] = 2.0 * (rp.c0 - kk) * a0;
  c->a2[0] = (c1 - kk - rp.c0) * a0;
  c->b0[0] = a0 * kk;
  c->b1[0] = 2.0 * c->b0[0];
  c->b2[0] = c->b0[0];
  a0 = 1.0 / (c3 + kk + rp.c2);
  c->a1[1] = 2.0 * (rp.c2 - kk) * a0;
  c->a2[1] = (c3 - kk - rp.c2) * a0;
  c->b0[1] = a0 * kk;
  c->b1[1] = 2.0 * c->b0[1];
  c->b2[1] = c->b0[1];
}
// Transposed direct form II, two sections in series; s = 4 state words.
inline double lp24_run(const Lp24Coef& c, double* s, double x) {
  double y1 = c.b0[0] * x + s[0];
  s[0] = c.b1[0] * x + c.a1[0] * y1 + s[1];
  s[1] = c.b2[0] * x + c.a2[0] * y1;
  double y2 = c.b0[1] * y1 + s[2];
  s[2] = c.b1[1] * y1 + c.a1[1] * y2 + s[3];
  s[3] = c.b2[1] * y1 + c.a2[1] * y2;
  return y2;
}

// ------------------------------------------------------------ voice store ---
// Host-side voice allocation, integer frames only.
struct VoiceSlot {
  int key = -1;
  int64_t on_frame = kNever;
  int64_t idle_at = kNever;  // first frame at which the voice is idle again
  bool held = false;
};
struct VoiceStore {
  std::vector<VoiceSlot> slots;
  // returns voice index; *was_playing tells whether it is a retrigger/steal of a sounding voice
  int note_on(int64_t f, int key, bool* was_playing) {
    int pick = -1;
    for (size_t i = 0; i < slots.size(); ++i)
      if (slots[i].key == key && f < slots[i].idle_at) { pick = (int)i; break; }
    if (pick < 0)
      for (size_t i = 0; i < slots.size(); ++i)
        if (f >= slots[i].idle_at) { pick = (int)i; break; }
    if (pick < 0) {
      pick = 0;
      for (size_t i = 1; i < slots.size(); ++i)
        if (slots[i].on_frame < slots[pick].on_frame) pick = (int)i;
    }
    *was_playing = f < slots[pick].idle_at;
    slots[pick].key = key;
    slots[pick].on_frame = f;
    slots[pick].idle_at = kHeld;
    slots[pick].held = true;
    return pick;
  }
  template <typename F>
  void note_off(int64_t f, int key, int64_t release_frames, F&& fn) {
    for (size_t i = 0; i < slots.size(); ++i)
      if (slots[i].key == key && slots[i].held) {
        slots[i].held = false;
        slots[i].idle_at = f + release_frames;
        fn((int)i);
      }
  }
};

// ---------------------------------------------------------------- entities ---
struct Entity {
  uint32_t uid = 0;
  int kind = 0;
  // Fan-out memo: a node reachable over several patch paths is evaluated ONCE per frame and its
  // output shared.  (The reference's walk would re-tick it once per path, orchestrator.rs:401-410 —
  // a DFS artifact that transposes the instrument; none of its 95 project fixtures has fan-out.)
  int64_t memo_frame = INT64_MIN;
  Stereo memo;
  virtual ~Entity() {}
  virtual bool is_instrument() const { return false; }
  // instruments
  virtual void tick(int64_t /*frame*/) {}
  virtual Stereo value() const { return Stereo(); }
  virtual void note_on(int64_t, int, int) {}
  virtual void note_off(int64_t, int) {}
  // effects
  virtual Stereo transform_audio(Stereo in) { return in; }
  // Controllable
  virtual int set_param(int /*index*/, double /*raw*/) { return GB_EINVAL; }
  virtual int control(int /*index*/, double /*value01*/) { return GB_EINVAL; }
  virtual int load_sample(uint8_t, const double*, size_t, int, double, double) { return GB_EINVAL; }
};

// -- Welsh subtractive voice ---------------------------------------------------
struct WelshVoice {
  int64_t n_on = kNever, n_off = kNever;
  EnvState amp, filt;
  bool reset_pending = false;
  Oscillator osc1, osc2, lfo;
  double cyc1 = 0.0, cyc2 = 0.0;  // base cycles/frame
  uint64_t d1 = 0, d2 = 0;
  double s[4] = {0, 0, 0, 0};
  Stereo out;
};
struct WelshSynth : Entity {
  gb_welsh_params p;
  double sr;
  EnvShape amp_shape, filt_shape;
  Lp24Ripple ripple;
  Lp24Coef fixed_coef;
  uint64_t lfo_dq = 0;
  double gl = 0, gr = 0;
  std::vector<WelshVoice> voices;
  VoiceStore store;
  Stereo sum;

  WelshSynth(const gb_welsh_params& pp, double sample_rate, uint32_t uid_) : p(pp), sr(sample_rate) {
    uid = uid_;
    kind = GB_INST_WELSH;
    uint32_t nv = p.voices ? p.voices : 8;
    voices.resize(nv);
    store.slots.resize(nv);
    amp_shape.set(p.amp_envelope, sr);
    filt_shape.set(p.filter_envelope, sr);
    ripple.set(p.filter_passband_ripple);
    lp24_coefficients(ripple, p.filter_cutoff_hz, sr, &fixed_coef);
    lfo_dq = cycles_to_q(p.lfo.frequency / sr);
    update_dca();
    for (uint32_t i = 0; i < nv; ++i) {
      WelshVoice& v = voices[i];
      v.osc1.waveform = p.oscillator_1.waveform;
      v.osc1.duty_q = cycles_to_q(std::min(clamp01(p.oscillator_1.pulse_width), 1.0 - 1.0 / 9007199254740992.0));
      v.osc1.seed = splitmix64(((uint64_t)uid << 32) ^ (uint64_t)(2 * i));
      v.osc2.waveform = p.oscillator_2.waveform;
      v.osc2.duty_q = cycles_to_q(std::min(clamp01(p.oscillator_2.pulse_width), 1.0 - 1.0 / 9007199254740992.0));
      v.osc2.seed = splitmix64(((uint64_t)uid << 32) ^ (uint64_t)(2 * i + 1));
      v.lfo.waveform = p.lfo.waveform;
      v.lfo.duty_q = cycles_to_q(std::min(clamp01(p.lfo.pulse_width), 1.0 - 1.0 / 9007199254740992.0));
      v.lfo.seed = splitmix64(((uint64_t)uid << 32) ^ UINT64_C(0x4C464F00) ^ (uint64_t)i);
    }
  }
  bool is_instrument() const override { return true; }
  void update_dca() {
    double pan = p.voice_dca.pan + p.dca.pan;
    pan = pan < -1.0 ? -1.0 : (pan > 1.0 ? 1.0 : pan);
    dca_gains(p.voice_dca.gain * p.dca.gain, pan, &gl, &gr);
  }
  double osc_cycles(const gb_oscillator_params& o, int key) const {
    double hz = o.fixed_frequency > 0.0 ? o.fixed_frequency : note_hz(key) * o.frequency_tune;
    return hz / sr;
  }
  bool playing(const WelshVoice& v, int64_t n) const { return n >= v.n_on && n < v.n_off + amp_shape.nr; }
  void note_on(int64_t f, int key, int) override {
    bool was;
    int i = store.note_on(f, key, &was);
    WelshVoice& v = voices[i];
    // `was` from the store is exact: idle_at == n_off + nr
    v.amp.l_on = was ? env_level(amp_shape, v.n_on, v.n_off, v.amp, f) : 0.0;
    v.filt.l_on = was ? env_level(filt_shape, v.n_on, v.n_off, v.filt, f) : 0.0;
    v.reset_pending = !was;
    v.n_on = f;
    v.n_off = kHeld;
    v.cyc1 = osc_cycles(p.oscillator_1, key);
    v.cyc2 = osc_cycles(p.oscillator_2, key);
    v.d1 = cycles_to_q(v.cyc1);
    v.d2 = cycles_to_q(v.cyc2);
  }
  void note_off(int64_t f, int key) override {
    store.note_off(f, key, amp_shape.nr, [&](int i) {
      WelshVoice& v = voices[i];
      v.amp.l_off = env_pre(amp_shape, v.n_on, v.amp.l_on, f);
      v.filt.l_off = env_pre(filt_shape, v.n_on, v.filt.l_on, f);
      v.n_off = f;
    });
  }
  void tick_voice(WelshVoice& v, int64_t n) {
    v.out = Stereo();
    if (!playing(v, n)) return;
    bool reset = v.reset_pending && n == v.n_on;
    if (reset) v.reset_pending = false;
    double l = 0.0;
    if (p.lfo_routing != GB_LFO_NONE) {
      v.lfo.tick(lfo_dq, reset);
      l = v.lfo.value(n, 0, false);
    }
    double ld = l * p.lfo_depth;
    uint64_t d1 = v.d1, d2 = v.d2;
    if (p.lfo_routing == GB_LFO_PITCH) {
      double f = std::exp2(ld);
      d1 = cycles_to_q(v.cyc1 * f);
      d2 = cycles_to_q(v.cyc2 * f);
    }
    bool wrapped = v.osc1.tick(d1, reset);
    v.osc2.tick(d2, reset || (p.oscillator_2_sync && wrapped));
    double o1, o2;
    if (p.lfo_routing == GB_LFO_PULSE_WIDTH) {
      const double top = 1.0 - 1.0 / 9007199254740992.0;
      double du1 = p.oscillator_1.pulse_width + 0.5 * ld;
      double du2 = p.oscillator_2.pulse_width + 0.5 * ld;
      du1 = du1 < 0.0 ? 0.0 : (du1 > top ? top : du1);
      du2 = du2 < 0.0 ? 0.0 : (du2 > top ? top : du2);
      o1 = v.osc1.value(n, cycles_to_q(du1), true);
      o2 = v.osc2.value(n, cycles_to_q(du2), true);
    } else {
      o1 = v.osc1.value(n, 0, false);
      o2 = v.osc2.value(n, 0, false);
    }
    double x = o1 * p.oscillator_mix + o2 * (1.0 - p.oscillator_mix);
    double y;
    if (p.filter_cutoff_end != 0.0) {
      double fe = env_level(filt_shape, v.n_on, v.n_off, v.filt, n);
      double pct = p.filter_cutoff_start + (1.0 - p.filter_cutoff_start) * p.filter_cutoff_end * fe;
      Lp24Coef c;
      lp24_coefficients(ripple, pct_to_hz(clamp01(pct)), sr, &c);
      y = lp24_run(c, v.s, x);
    } else if (p.lfo_routing == GB_LFO_FILTER_CUTOFF) {
      double pct = p.filter_cutoff_start * (1.0 + ld);
      Lp24Coef c;
      lp24_coefficients(ripple, pct_to_hz(clamp01(pct)), sr, &c);
      y = lp24_run(c, v.s, x);
    } else {
      y = lp24_run(fixed_coef, v.s, x);
    }
    double ae = env_level(amp_shape, v.n_on, v.n_off, v.amp, n);
    double amp_lfo = p.lfo_routing == GB_LFO_AMPLITUDE ? 0.5 * (1.0 + ld) : 0.5;
    double m = y * ae * amp_lfo;
    v.out.l = m * gl;
    v.out.r = m * gr;
  }
  void tick(int64_t n) override {
    sum = Stereo();
    for (auto& v : voices) {
      tick_voice(v, n);
      sum.l += v.out.l;
      sum.r += v.out.r;
    }
  }
  Stereo value() const override { return sum; }
  int set_param(int index, double raw) override {
    if (index == GB_CTL_INST_DCA_GAIN) p.dca.gain = raw;
    else if (index == GB_CTL_INST_DCA_PAN) p.dca.pan = raw;
    else return GB_EINVAL;
    update_dca();
    return 0;
  }
  int control(int index, double v) override {
    return set_param(index, index == GB_CTL_INST_DCA_PAN ? 2.0 * v - 1.0 : v);
  }
};

// -- FM voice -------------------------------------------------------------------
struct FmVoice {
  int64_t n_on = kNever, n_off = kNever;
  EnvState car, mod;
  bool reset_pending = false;
  Oscillator carrier, modulator;
  double cyc_c = 0.0;
  uint64_t dm = 0;
  Stereo out;
};
struct FmSynth : Entity {
  gb_fm_params p;
  double sr;
  EnvShape car_shape, mod_shape;
  double gl = 0, gr = 0;
  std::vector<FmVoice> voices;
  VoiceStore store;
  Stereo sum;
  FmSynth(const gb_fm_params& pp, double sample_rate, uint32_t uid_) : p(pp), sr(sample_rate) {
    uid = uid_;
    kind = GB_INST_FM;
    uint32_t nv = p.voices ? p.voices : 8;
    voices.resize(nv);
    store.slots.resize(nv);
    car_shape.set(p.carrier_envelope, sr);
    mod_shape.set(p.modulator_envelope, sr);
    dca_gains(p.dca.gain, p.dca.pan, &gl, &gr);
  }
  bool is_instrument() const override { return true; }
  void note_on(int64_t f, int key, int) override {
    bool was;
    int i = store.note_on(f, key, &was);
    FmVoice& v = voices[i];
    v.car.l_on = was ? env_level(car_shape, v.n_on, v.n_off, v.car, f) : 0.0;
    v.mod.l_on = was ? env_level(mod_shape, v.n_on, v.n_off, v.mod, f) : 0.0;
    v.reset_pending = !was;
    v.n_on = f;
    v.n_off = kHeld;
    v.cyc_c = note_hz(key) / sr;
    v.dm = cycles_to_q(v.cyc_c * p.ratio);
  }
  void note_off(int64_t f, int key) override {
    store.note_off(f, key, car_shape.nr, [&](int i) {
      FmVoice& v = voices[i];
      v.car.l_off = env_pre(car_shape, v.n_on, v.car.l_on, f);
      v.mod.l_off = env_pre(mod_shape, v.n_on, v.mod.l_on, f);
      v.n_off = f;
    });
  }
  void tick(int64_t n) override {
    sum = Stereo();
    for (auto& v : voices) {
      v.out = Stereo();
      if (!(n >= v.n_on && n < v.n_off + car_shape.nr)) continue;
      bool reset = v.reset_pending && n == v.n_on;
      if (reset) v.reset_pending = false;
      v.modulator.tick(v.dm, reset);
      double mval = std::sin(pos_of(v.modulator.phase) * kTwoPi);
      double menv = env_level(mod_shape, v.n_on, v.n_off, v.mod, n);
      double x = mval * menv * p.depth * p.beta;
      uint64_t dc = cycles_to_q(v.cyc_c * (1.0 + x));
      v.carrier.tick(dc, reset);
      double cval = std::sin(pos_of(v.carrier.phase) * kTwoPi);
      double m = cval * env_level(car_shape, v.n_on, v.n_off, v.car, n);
      v.out.l = m * gl;
      v.out.r = m * gr;
      sum.l += v.out.l;
      sum.r += v.out.r;
    }
  }
  Stereo value() const override { return sum; }
  int set_param(int index, double raw) override {
    if (index == GB_CTL_INST_DCA_GAIN) p.dca.gain = raw;
    else if (index == GB_CTL_INST_DCA_PAN) p.dca.pan = raw;
    else return GB_EINVAL;
    dca_gains(p.dca.gain, p.dca.pan, &gl, &gr);
    return 0;
  }
  int control(int index, double v) override {
    return set_param(index, index == GB_CTL_INST_DCA_PAN ? 2.0 * v - 1.0 : v);
  }
};

// -- Sampler / Drumkit -----------------------------------------------------------
struct SampleData {
  std::vector<double> frames;  // interleaved if channels == 2
  size_t n = 0;
  int channels = 1;
  double sr = 44100.0, root_hz = 0.0;
};
struct SampleVoice {
  int64_t n_on = kNever, n_end = kNever;  // plays for n in [n_on, n_end)
  uint64_t step_q = 0;                    // 32.32 fixed-point sample frames per output frame
  int sample = -1;
};
inline uint64_t step_to_q32(double step) {
  double r = step * 4294967296.0;
  if (!(r >= 1.0)) return 1;
  if (r > 1.8e19) r = 1.8e19;
  return (uint64_t)r;
}
inline int64_t frames_until_end(size_t len, uint64_t step_q) {
  // smallest k >= 0 with (k*step_q)>>32 >= len
  unsigned __int128 target = (unsigned __int128)len << 32;
  unsigned __int128 k = (target + step_q - 1) / step_q;
  if (k > (unsigned __int128)kHeld) return kHeld;
  return (int64_t)k;
}
struct SamplerBase : Entity {
  double sr;
  std::vector<SampleData> samples;
  std::vector<SampleVoice> voices;
  Stereo sum;
  bool is_instrument() const override { return true; }
  void tick(int64_t n) override {
    sum = Stereo();
    for (auto& v : voices) {
      if (!(n >= v.n_on && n < v.n_end)) continue;
      const SampleData& s = samples[v.sample];
      uint64_t idx = ((uint64_t)(n - v.n_on) * v.step_q) >> 32;
      if (idx >= s.n) continue;
      if (s.channels == 2) {
        sum.l += s.frames[2 * idx];
        sum.r += s.frames[2 * idx + 1];
      } else {
        sum.l += s.frames[idx];
        sum.r += s.frames[idx];
      }
    }
  }
  Stereo value() const override { return sum; }
};
struct Sampler : SamplerBase {
  gb_sampler_params p;
  VoiceStore store;
  Sampler(const gb_sampler_params& pp, double sample_rate, uint32_t uid_) : p(pp) {
    sr = sample_rate;
    uid = uid_;
    kind = GB_INST_SAMPLER;
    uint32_t nv = p.voices ? p.voices : 8;
    voices.resize(nv);
    store.slots.resize(nv);
  }
  int load_sample(uint8_t, const double* f, size_t n, int ch, double ssr, double root) override {
    SampleData d;
    d.frames.assign(f, f + n * (size_t)ch);
    d.n = n;
    d.channels = ch;
    d.sr = ssr;
    d.root_hz = root > 0.0 ? root : (p.root_hz > 0.0 ? p.root_hz : 440.0);
    samples.clear();
    samples.push_back(std::move(d));
    return 0;
  }
  void note_on(int64_t f, int key, int) override {
    if (samples.empty()) return;
    bool was;
    int i = store.note_on(f, key, &was);
    SampleVoice& v = voices[i];
    const SampleData& s = samples[0];
    v.sample = 0;
    v.step_q = step_to_q32(note_hz(key) / s.root_hz * (s.sr / sr));
    v.n_on = f;
    v.n_end = f + frames_until_end(s.n, v.step_q);
    store.slots[i].idle_at = v.n_end;
  }
  void note_off(int64_t f, int key) override {
    store.note_off(f, key, 0, [&](int i) {
      if (f < voices[i].n_end) voices[i].n_end = f;
    });
  }
};
struct Drumkit : SamplerBase {
  int key_to_voice[128];
  Drumkit(double sample_rate, uint32_t uid_) {
    sr = sample_rate;
    uid = uid_;
    kind = GB_INST_DRUMKIT;
    for (int& k : key_to_voice) k = -1;
  }
  int load_sample(uint8_t key, const double* f, size_t n, int ch, double ssr, double) override {
    if (key >= 128) return GB_EINVAL;
    SampleData d;
    d.frames.assign(f, f + n * (size_t)ch);
    d.n = n;
    d.channels = ch;
    d.sr = ssr;
    if (key_to_voice[key] < 0) {
      key_to_voice[key] = (int)voices.size();
      samples.push_back(std::move(d));
      SampleVoice v;
      v.sample = (int)samples.size() - 1;
      voices.push_back(v);
    } else {
      samples[voices[key_to_voice[key]].sample] = std::move(d);
    }
    return 0;
  }
  void note_on(int64_t f, int key, int) override {
    if (key < 0 || key >= 128 || key_to_voice[key] < 0) return;
    SampleVoice& v = voices[key_to_voice[key]];
    const SampleData& s = samples[v.sample];
    v.step_q = step_to_q32(s.sr / sr);
    v.n_on = f;
    v.n_end = f + frames_until_end(s.n, v.step_q);
  }
};
struct ToySource : Entity {
  gb_toy_source_params p;
  ToySource(const gb_toy_source_params& pp, uint32_t uid_) : p(pp) {
    uid = uid_;
    kind = GB_INST_TOY_SOURCE;
  }
  bool is_instrument() const override { return true; }
  Stereo value() const override {
    Stereo s;
    s.l = p.level_left;
    s.r = p.level_right;
    return s;
  }
};

// A bare Oscillator / Envelope as a device: the "oscillator" and "envelope" instrument types of the
// reference's older project fixtures (projects/demos/effects/filter-*.json, gain_*.json, bitcrusher_*.json,
// projects/demos/instruments/oscillator-*.json, envelope-adsr-linear.json; the enum variants are gone from
// settings/src/instruments.rs:24-39 in this snapshot).  parity unpinned: docs/ORACLE_SPEC.md §3a.
struct OscillatorSource : Entity {
  Oscillator osc;
  uint64_t dq = 0;
  bool started = false;
  double out = 0.0;
  OscillatorSource(const gb_oscillator_source_params& pp, double sr, uint32_t uid_) {
    uid = uid_;
    kind = GB_INST_OSCILLATOR;
    const double top = 1.0 - 1.0 / 9007199254740992.0;
    osc.waveform = pp.oscillator.waveform;
    osc.duty_q = cycles_to_q(std::min(clamp01(pp.oscillator.pulse_width), top));
    osc.seed = splitmix64((uint64_t)uid << 32);
    dq = cycles_to_q(pp.oscillator.frequency / sr);
  }
  bool is_instrument() const override { return true; }
  void tick(int64_t frame) override {  // free-running from the first rendered frame: phase(n) = n * dq
    osc.tick(dq, !started);
    started = true;
    out = osc.value(frame, 0, false);
  }
  Stereo value() const override {
    Stereo s;
    s.l = out;
    s.r = out;
    return s;
  }
};
struct EnvelopeSource : Entity {
  EnvShape shape;
  EnvState st;
  int64_t n_on = kNever, n_off = kNever;
  double out = 0.0;
  EnvelopeSource(const gb_envelope_source_params& pp, double sr, uint32_t uid_) {
    uid = uid_;
    kind = GB_INST_ENVELOPE;
    shape.set(pp.envelope, sr);
  }
  bool is_instrument() const override { return true; }
  void note_on(int64_t f, int, int) override {  // retrigger: the attack starts from the current level
    st.l_on = env_level(shape, n_on, n_off, st, f);
    n_on = f;
    n_off = kHeld;
  }
  void note_off(int64_t f, int) override {
    if (n_off != kHeld) return;
    st.l_off = env_pre(shape, n_on, st.l_on, f);
    n_off = f;
  }
  void tick(int64_t frame) override { out = env_level(shape, n_on, n_off, st, frame); }
  Stereo value() const override {
    Stereo s;
    s.l = out;
    s.r = out;
    return s;
  }
};

// -- effects ---------------------------------------------------------------------
struct Mixer : Entity {};
// SignalPassthroughController (settings/src/controllers.rs:110-111,181-187; source ABSENT): patched into a
// chain like an effect, passes the audio through; its last output is what control links read.
struct SignalPassthrough : Entity {};
struct Gain : Entity {
  double ceiling;
  Stereo transform_audio(Stereo in) override {
    in.l *= ceiling;
    in.r *= ceiling;
    return in;
  }
  int set_param(int i, double v) override {
    if (i != GB_CTL_GAIN_CEILING) return GB_EINVAL;
    ceiling = v;
    return 0;
  }
  int control(int i, double v) override { return set_param(i, v); }
};
struct Limiter : Entity {
  double mn, mx;
  static double ch(double x, double mn, double mx) {
    double a = std::fabs(x);
    a = a < mn ? mn : (a > mx ? mx : a);
    return std::signbit(x) ? -a : a;
  }
  Stereo transform_audio(Stereo in) override {
    in.l = ch(in.l, mn, mx);
    in.r = ch(in.r, mn, mx);
    return in;
  }
  int set_param(int i, double v) override {
    if (i == GB_CTL_LIMITER_MIN) mn = v;
    else if (i == GB_CTL_LIMITER_MAX) mx = v;
    else return GB_EINVAL;
    return 0;
  }
  int control(int i, double v) override { return set_param(i, v); }
};
struct Bitcrusher : Entity {
  double bits, c;
  void set_bits(double b) {
    bits = b;
    c = std::exp2(std::floor(b));
  }
  static double ch(double x, double c) {
    double a = std::floor(std::fabs(x) * 32767.0 / c) * c / 32767.0;
    return std::signbit(x) ? -a : a;
  }
  Stereo transform_audio(Stereo in) override {
    in.l = ch(in.l, c);
    in.r = ch(in.r, c);
    return in;
  }
  int set_param(int i, double v) override {
    if (i != GB_CTL_BITCRUSHER_BITS) return GB_EINVAL;
    set_bits(v);
    return 0;
  }
  int control(int i, double v) override { return set_param(i, std::floor(v * 16.0)); }
};
struct Compressor : Entity {
  double threshold, ratio;
  static double ch(double x, double th, double ratio) {
    double a = std::fabs(x);
    if (a > th) a = th + (a - th) * ratio;
    return std::signbit(x) ? -a : a;
  }
  Stereo transform_audio(Stereo in) override {
    in.l = ch(in.l, threshold, ratio);
    in.r = ch(in.r, threshold, ratio);
    return in;
  }
  int set_param(int i, double v) override {
    if (i == GB_CTL_COMPRESSOR_THRESHOLD) threshold = v;
    else if (i == GB_CTL_COMPRESSOR_RATIO) ratio = v;
    else return GB_EINVAL;
    return 0;
  }
  int control(int i, double v) override { return set_param(i, v); }
};

// RBJ cookbook biquad, Direct Form 1 (doc/Audio-EQ-Cookbook.txt:35-43), per channel.
struct BiquadCoef {
  double b0, b1, b2, a1, a2;  // a0-normalised
};
inline BiquadCoef rbj(int kind, double cutoff, double p2, double sr) {
  double fc = cutoff;
  if (fc > 0.49 * sr) fc = 0.49 * sr;
  if (fc < 1e-3) fc = 1e-3;
  double w0 = kTwoPi * fc / sr;
  double cs = std::cos(w0), sn = std::sin(w0);
  double b0, b1, b2, a0, a1, a2;
  switch (kind) {
    case GB_FX_LOW_PASS_12DB: {
      double q = p2 > 1e-6 ? p2 : 1e-6;
      double alpha = sn / (2.0 * q);
      b0 = (1.0 - cs) / 2.0; b1 = 1.0 - cs; b2 = (1.0 - cs) / 2.0;
      a0 = 1.0 + alpha; a1 = -2.0 * cs; a2 = 1.0 - alpha;
    } break;
    case GB_FX_HIGH_PASS_12DB: {
      double q = p2 > 1e-6 ? p2 : 1e-6;
      double alpha = sn / (2.0 * q);
      b0 = (1.0 + cs) / 2.0; b1 = -(1.0 + cs); b2 = (1.0 + cs) / 2.0;
      a0 = 1.0 + alpha; a1 = -2.0 * cs; a2 = 1.0 - alpha;
    } break;
    case GB_FX_BAND_PASS_12DB: {
      double bw = p2 > 1e-6 ? p2 : 1e-6;
      double alpha = sn * std::sinh(0.34657359027997264 * bw * w0 / sn);  // ln2/2
      b0 = alpha; b1 = 0.0; b2 = -alpha;
      a0 = 1.0 + alpha; a1 = -2.0 * cs; a2 = 1.0 - alpha;
    } break;
    case GB_FX_BAND_STOP_12DB: {
      double bw = p2 > 1e-6 ? p2 : 1e-6;
      double alpha = sn * std::sinh(0.34657359027997264 * bw * w0 / sn);
      b0 = 1.0; b1 = -2.0 * cs; b2 = 1.0;
      a0 = 1.0 + alpha; a1 = -2.0 * cs; a2 = 1.0 - alpha;
    } break;
    case GB_FX_ALL_PASS_12DB: {
      double q = p2 > 1e-6 ? p2 : 1e-6;
      double alpha = sn / (2.0 * q);
      b0 = 1.0 - alpha; b1 = -2.0 * cs; b2 = 1.0 + alpha;
      a0 = 1.0 + alpha; a1 = -2.0 * cs; a2 = 1.0 - alpha;
    } break;
    case GB_FX_PEAKING_EQ_12DB: {
      double A = std::pow(10.0, p2 / 40.0);
      double alpha = sn / (2.0 * 0.70710678118654752440);
      b0 = 1.0 + alpha * A; b1 = -2.0 * cs; b2 = 1.0 - alpha * A;
      a0 = 1.0 + alpha / A; a1 = -2.0 * cs; a2 = 1.0 - alpha / A;
    } break;
    case GB_FX_LOW_SHELF_12DB: {
      double A = std::pow(10.0, p2 / 40.0);
      double alpha = sn / 2.0 * 1.41421356237309504880;  // S = 1
      double t = 2.0 * std::sqrt(A) * alpha;
      b0 = A * ((A + 1.0) - (A - 1.0) * cs + t);
      b1 = 2.0 * A * ((A - 1.0) - (A + 1.0) * cs);
      b2 = A * ((A + 1.0) - (A - 1.0) * cs - t);
      a0 = (A + 1.0) + (A - 1.0) * cs + t;
      a1 = -2.0 * ((A - 1.0) + (A + 1.0) * cs);
      a2 = (A + 1.0) + (A - 1.0) * cs - t;
    } break;
    default: {  // GB_FX_HIGH_SHELF_12DB
      double A = std::pow(10.0, p2 / 40.0);
      double alpha = sn / 2.0 * 1.41421356237309504880;
      double t = 2.0 * std::sqrt(A) * alpha;
      b0 = A * ((A + 1.0) + (A - 1.0) * cs + t);
      b1 = -2.0 * A * ((A - 1.0) + (A + 1.0) * cs);
      b2 = A * ((A + 1.0) + (A - 1.0) * cs - t);
      a0 = (A + 1.0) - (A - 1.0) * cs + t;
      a1 = 2.0 * ((A - 1.0) - (A + 1.0) * cs);
      a2 = (A + 1.0) - (A - 1.0) * cs - t;
    } break;
  }
  BiquadCoef c;
  c.b0 = b0 / a0; c.b1 = b1 / a0; c.b2 = b2 / a0; c.a1 = a1 / a0; c.a2 = a2 / a0;
  return c;
}
struct Biquad : Entity {
  double sr, cutoff, p2;
  BiquadCoef c;
  double x1[2] = {0, 0}, x2[2] = {0, 0}, y1[2] = {0, 0}, y2[2] = {0, 0};
  void update() { c = rbj(kind, cutoff, p2, sr); }
  double ch(int i, double x) {
    double y = c.b0 * x + c.b1 * x1[i] + c.b2 * x2[i] - c.a1 * y1[i] - c.a2 * y2[i];
    x2[i] = x1[i]; x1[i] = x;
    y2[i] = y1[i]; y1[i] = y;
    return y;
  }
  Stereo transform_audio(Stereo in) override {
    in.l = ch(0, in.l);
    in.r = ch(1, in.r);
    return in;
  }
  int set_param(int i, double v) override {
    if (i == GB_CTL_FILTER_CUTOFF) cutoff = v;
    else if (i == GB_CTL_FILTER_PARAM2) p2 = v;
    else return GB_EINVAL;
    update();
    return 0;
  }
  int control(int i, double v) override {
    if (i == GB_CTL_FILTER_CUTOFF) return set_param(i, pct_to_hz(clamp01(v)));
    if (i == GB_CTL_FILTER_PARAM2) {
      switch (kind) {
        case GB_FX_BAND_PASS_12DB:
        case GB_FX_BAND_STOP_12DB: return set_param(i, v * 4.0);
        case GB_FX_PEAKING_EQ_12DB:
        case GB_FX_LOW_SHELF_12DB:
        case GB_FX_HIGH_SHELF_12DB: return set_param(i, (2.0 * v - 1.0) * 24.0);
        default: return set_param(i, denormalize_q(v));
      }
    }
    return GB_EINVAL;
  }
};
struct LowPass24 : Entity {
  double sr, cutoff, ripple_v;
  Lp24Ripple ripple;
  Lp24Coef c;
  double s[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  void update() {
    ripple.set(ripple_v);
    lp24_coefficients(ripple, cutoff, sr, &c);
  }
  Stereo transform_audio(Stereo in) override {
    in.l = lp24_run(c, s[0], in.l);
    in.r = lp24_run(c, s[1], in.r);
    return in;
  }
  int set_param(int i, double v) override {
    if (i == GB_CTL_FILTER_CUTOFF) cutoff = v;
    else if (i == GB_CTL_FILTER_PARAM2) ripple_v = v;
    else return GB_EINVAL;
    update();
    return 0;
  }
  int control(int i, double v) override {
    if (i == GB_CTL_FILTER_CUTOFF) return set_param(i, pct_to_hz(clamp01(v)));
    if (i == GB_CTL_FILTER_PARAM2) return set_param(i, denormalize_q(v));
    return GB_EINVAL;
  }
};

// Delay-line primitive: a ring of `d` frames per channel.
struct Ring {
  std::vector<double> buf;
  size_t ptr = 0;
  void resize(size_t d) {
    buf.assign(d, 0.0);
    ptr = 0;
  }
};
struct Delay : Entity {
  Ring ring[2];
  Delay(double seconds, double sr) {
    size_t d = (size_t)frames_of(seconds, sr);
    ring[0].resize(d);
    ring[1].resize(d);
  }
  static double ch(Ring& r, double x) {
    if (r.buf.empty()) return x;
    double out = r.buf[r.ptr];
    r.buf[r.ptr] = x;
    r.ptr = (r.ptr + 1) % r.buf.size();
    return out;
  }
  Stereo transform_audio(Stereo in) override {
    in.l = ch(ring[0], in.l);
    in.r = ch(ring[1], in.r);
    return in;
  }
};
// Chorus: `voices` equally spaced taps on one delay line, out = (1-w)*x + w*mean(taps).
struct Chorus : Entity {
  int nv;
  double wet;
  std::vector<int64_t> taps;
  std::vector<double> hist[2];  // hist[c][k] = x[n-1-k]
  size_t len = 0, head = 0;
  Chorus(const gb_chorus_params& p, double sr) {
    nv = (int)p.voices;
    if (nv < 1) nv = 1;
    if (nv > 64) nv = 64;
    wet = p.wet_dry_mix;
    int64_t d = frames_of(p.delay_seconds, sr);
    for (int i = 0; i < nv; ++i) taps.push_back(d * (i + 1) / nv);
    len = (size_t)std::max<int64_t>(d, 1);
    hist[0].assign(len, 0.0);
    hist[1].assign(len, 0.0);
  }
  double ch(int c, double x) {
    double acc = 0.0;
    for (int i = 0; i < nv; ++i) {
      int64_t t = taps[i];
      acc += t == 0 ? x : hist[c][(head + len - (size_t)t) % len];
    }
    double y = (1.0 - wet) * x + wet * (acc / (double)nv);
    hist[c][head] = x;
    return y;
  }
  Stereo transform_audio(Stereo in) override {
    // hist[c][(head - t) % len] = x[n - t] for t in 1..len
    Stereo o;
    o.l = ch(0, in.l);
    o.r = ch(1, in.r);
    head = (head + 1) % len;
    return o;
  }
  int set_param(int i, double v) override {
    if (i != GB_CTL_CHORUS_WET_DRY_MIX) return GB_EINVAL;
    wet = v;
    return 0;
  }
  int control(int i, double v) override { return set_param(i, v); }
};
// Schroeder reverberator: 4 parallel recirculating combs + 2 series all-passes.
struct Reverb : Entity {
  double attenuation;
  struct Line {
    Ring r;
    double g;
  };
  Line comb[2][4], ap[2][2];
  Reverb(const gb_reverb_params& p, double sr) {
    attenuation = p.attenuation;
    static const double comb_s[4] = {0.0297, 0.0371, 0.0411, 0.0437};
    static const double ap_delay[2] = {0.0050, 0.0017};
    static const double ap_decay[2] = {0.09683, 0.03292};
    double seconds = p.seconds > 1e-6 ? p.seconds : 1e-6;
    for (int c = 0; c < 2; ++c) {
      for (int i = 0; i < 4; ++i) {
        comb[c][i].r.resize((size_t)std::max<int64_t>(frames_of(comb_s[i], sr), 1));
        comb[c][i].g = std::pow(0.001, comb_s[i] / seconds);
      }
      for (int i = 0; i < 2; ++i) {
        ap[c][i].r.resize((size_t)std::max<int64_t>(frames_of(ap_delay[i], sr), 1));
        ap[c][i].g = std::pow(0.001, ap_delay[i] / ap_decay[i]);
      }
    }
  }
  // comb: y[n] = w[n-D], w[n] = x[n] + g*w[n-D]
  static double comb_tick(Line& l, double x) {
    double out = l.r.buf[l.r.ptr];
    l.r.buf[l.r.ptr] = x + l.g * out;
    l.r.ptr = (l.r.ptr + 1) % l.r.buf.size();
    return out;
  }
  // all-pass: y[n] = -g*x[n] + w[n-D], w[n] = x[n] + g*y[n]
  static double ap_tick(Line& l, double x) {
    double d = l.r.buf[l.r.ptr];
    double y = -l.g * x + d;
    l.r.buf[l.r.ptr] = x + l.g * y;
    l.r.ptr = (l.r.ptr + 1) % l.r.buf.size();
    return y;
  }
  double ch(int c, double x) {
    double xa = x * attenuation;
    double s = 0.0;
    for (int i = 0; i < 4; ++i) s += comb_tick(comb[c][i], xa);
    return ap_tick(ap[c][1], ap_tick(ap[c][0], s));
  }
  Stereo transform_audio(Stereo in) override {
    in.l = ch(0, in.l);
    in.r = ch(1, in.r);
    return in;
  }
  int set_param(int i, double v) override {
    if (i != GB_CTL_REVERB_ATTENUATION) return GB_EINVAL;
    attenuation = v;
    return 0;
  }
  int control(int i, double v) override { return set_param(i, v); }
};

}  // namespace

// ------------------------------------------------------------------ engine ---
struct go_engine {
  double sr = 44100.0;
  uint32_t next_uid = 2;
  int64_t pos = 0;
  bool finalized = false;
  std::map<uint32_t, std::unique_ptr<Entity>> store;
  std::map<uint32_t, std::vector<uint32_t>> patches;  // effect uid -> sources, in patch order
  std::vector<gb_event> events;                       // pending, sorted by frame (stable)
  struct Link { uint32_t src, dst; int index; };
  std::vector<Link> links;                            // control links from signal-passthrough nodes
  std::string err;
};
static std::string g_create_err;

static int fail(go_engine* e, int code, const char* msg) {
  if (e) e->err = msg;
  else g_create_err = msg;
  return code;
}

extern "C" {

int go_create(const gb_config* cfg, go_engine** out) {
  if (!cfg || !out) return fail(nullptr, GB_EINVAL, "null argument");
  if (cfg->abi_version != GB_ABI_VERSION) return fail(nullptr, GB_EINVAL, "ABI version mismatch");
  if (!(cfg->sample_rate > 0.0)) return fail(nullptr, GB_EINVAL, "sample_rate must be positive");
  go_engine* e = new go_engine();
  e->sr = cfg->sample_rate;
  auto m = std::make_unique<Mixer>();
  m->uid = GB_MAIN_MIXER;
  m->kind = GB_FX_MIXER;
  e->store[GB_MAIN_MIXER] = std::move(m);
  *out = e;
  return 0;
}
void go_destroy(go_engine* e) { delete e; }
const char* go_last_error(const go_engine* e) { return e ? e->err.c_str() : g_create_err.c_str(); }

int go_add_instrument(go_engine* e, int32_t kind, const void* params, size_t size, uint32_t* uid) {
  if (!e || !uid) return GB_EINVAL;
  if (e->finalized) return fail(e, GB_ESTATE, "engine is finalized");
  uint32_t id = e->next_uid;
  std::unique_ptr<Entity> ent;
  switch (kind) {
    case GB_INST_WELSH:
      if (!params || size != sizeof(gb_welsh_params)) return fail(e, GB_EINVAL, "bad welsh params size");
      ent = std::make_unique<WelshSynth>(*(const gb_welsh_params*)params, e->sr, id);
      break;
    case GB_INST_FM:
      if (!params || size != sizeof(gb_fm_params)) return fail(e, GB_EINVAL, "bad fm params size");
      ent = std::make_unique<FmSynth>(*(const gb_fm_params*)params, e->sr, id);
      break;
    case GB_INST_SAMPLER:
      if (!params || size != sizeof(gb_sampler_params)) return fail(e, GB_EINVAL, "bad sampler params size");
      ent = std::make_unique<Sampler>(*(const gb_sampler_params*)params, e->sr, id);
      break;
    case GB_INST_DRUMKIT:
      ent = std::make_unique<Drumkit>(e->sr, id);
      break;
    case GB_INST_TOY_SOURCE:
      if (!params || size != sizeof(gb_toy_source_params)) return fail(e, GB_EINVAL, "bad toy params size");
      ent = std::make_unique<ToySource>(*(const gb_toy_source_params*)params, id);
      break;
    case GB_INST_OSCILLATOR:
      if (!params || size != sizeof(gb_oscillator_source_params)) return fail(e, GB_EINVAL, "bad oscillator params size");
      ent = std::make_unique<OscillatorSource>(*(const gb_oscillator_source_params*)params, e->sr, id);
      break;
    case GB_INST_ENVELOPE:
      if (!params || size != sizeof(gb_envelope_source_params)) return fail(e, GB_EINVAL, "bad envelope params size");
      ent = std::make_unique<EnvelopeSource>(*(const gb_envelope_source_params*)params, e->sr, id);
      break;
    default:
      return fail(e, GB_EINVAL, "unknown instrument kind");
  }
  e->store[id] = std::move(ent);
  e->next_uid++;
  *uid = id;
  return 0;
}

int go_add_effect(go_engine* e, int32_t kind, const void* params, size_t size, uint32_t* uid) {
  if (!e || !uid) return GB_EINVAL;
  if (e->finalized) return fail(e, GB_ESTATE, "engine is finalized");
  uint32_t id = e->next_uid;
  std::unique_ptr<Entity> ent;
#define NEED(T) if (!params || size != sizeof(T)) return fail(e, GB_EINVAL, "bad effect params size")
  switch (kind) {
    case GB_FX_MIXER: ent = std::make_unique<Mixer>(); break;
    case GB_FX_SIGNAL_PASSTHROUGH: ent = std::make_unique<SignalPassthrough>(); break;
    case GB_FX_GAIN: {
      NEED(gb_gain_params);
      auto g = std::make_unique<Gain>();
      g->ceiling = ((const gb_gain_params*)params)->ceiling;
      ent = std::move(g);
    } break;
    case GB_FX_LIMITER: {
      NEED(gb_limiter_params);
      auto g = std::make_unique<Limiter>();
      g->mn = ((const gb_limiter_params*)params)->min;
      g->mx = ((const gb_limiter_params*)params)->max;
      ent = std::move(g);
    } break;
    case GB_FX_BITCRUSHER: {
      NEED(gb_bitcrusher_params);
      auto g = std::make_unique<Bitcrusher>();
      g->set_bits(((const gb_bitcrusher_params*)params)->bits);
      ent = std::move(g);
    } break;
    case GB_FX_COMPRESSOR: {
      NEED(gb_compressor_params);
      auto g = std::make_unique<Compressor>();
      g->threshold = ((const gb_compressor_params*)params)->threshold;
      g->ratio = ((const gb_compressor_params*)params)->ratio;
      ent = std::move(g);
    } break;
    case GB_FX_DELAY: {
      NEED(gb_delay_params);
      ent = std::make_unique<Delay>(((const gb_delay_params*)params)->seconds, e->sr);
    } break;
    case GB_FX_CHORUS: {
      NEED(gb_chorus_params);
      ent = std::make_unique<Chorus>(*(const gb_chorus_params*)params, e->sr);
    } break;
    case GB_FX_REVERB: {
      NEED(gb_reverb_params);
      ent = std::make_unique<Reverb>(*(const gb_reverb_params*)params, e->sr);
    } break;
    case GB_FX_LOW_PASS_12DB: case GB_FX_HIGH_PASS_12DB: case GB_FX_BAND_PASS_12DB:
    case GB_FX_BAND_STOP_12DB: case GB_FX_ALL_PASS_12DB: case GB_FX_PEAKING_EQ_12DB:
    case GB_FX_LOW_SHELF_12DB: case GB_FX_HIGH_SHELF_12DB: {
      NEED(gb_biquad_params);
      auto g = std::make_unique<Biquad>();
      g->kind = kind;
      g->sr = e->sr;
      g->cutoff = ((const gb_biquad_params*)params)->cutoff;
      g->p2 = ((const gb_biquad_params*)params)->param2;
      g->update();
      ent = std::move(g);
    } break;
    case GB_FX_LOW_PASS_24DB: {
      NEED(gb_lowpass24_params);
      auto g = std::make_unique<LowPass24>();
      g->sr = e->sr;
      g->cutoff = ((const gb_lowpass24_params*)params)->cutoff;
      g->ripple_v = ((const gb_lowpass24_params*)params)->passband_ripple;
      g->update();
      ent = std::move(g);
    } break;
    default:
      return fail(e, GB_EINVAL, "unknown effect kind");
  }
#undef NEED
  ent->uid = id;
  ent->kind = kind;
  e->store[id] = std::move(ent);
  e->next_uid++;
  *uid = id;
  return 0;
}

int go_load_sample(go_engine* e, uint32_t uid, uint8_t key, const double* frames, size_t n, int32_t channels,
                   double sample_rate, double root_hz) {
  if (!e || !frames || n == 0 || (channels != 1 && channels != 2)) return fail(e, GB_EINVAL, "bad sample");
  auto it = e->store.find(uid);
  if (it == e->store.end()) return fail(e, GB_ENOENT, "unknown uid");
  if (e->finalized) return fail(e, GB_ESTATE, "engine is finalized");
  int rc = it->second->load_sample(key, frames, n, channels, sample_rate, root_hz);
  if (rc) return fail(e, rc, "entity does not take samples");
  return 0;
}

int go_patch(go_engine* e, uint32_t src, uint32_t dst) {
  if (!e) return GB_EINVAL;
  if (e->finalized) return fail(e, GB_ESTATE, "engine is finalized");
  auto s = e->store.find(src), d = e->store.find(dst);
  if (s == e->store.end() || d == e->store.end()) return fail(e, GB_ENOENT, "unknown uid");
  if (d->second->is_instrument()) return fail(e, GB_EGRAPH, "input device is not an effect");
  if (src == dst) return fail(e, GB_EGRAPH, "cannot patch a device to itself");
  auto& v = e->patches[dst];
  if (std::find(v.begin(), v.end(), src) == v.end()) v.push_back(src);
  return 0;
}

// Orchestrator::link_control_by_name (orchestration/src/orchestrator.rs:207-234) for an audio-rate source.
int go_link_control(go_engine* e, uint32_t src, uint32_t dst, int32_t index) {
  if (!e) return GB_EINVAL;
  if (e->finalized) return fail(e, GB_ESTATE, "engine is finalized");
  auto s = e->store.find(src), d = e->store.find(dst);
  if (s == e->store.end() || d == e->store.end()) return fail(e, GB_ENOENT, "unknown uid");
  if (s->second->kind != GB_FX_SIGNAL_PASSTHROUGH) return fail(e, GB_EINVAL, "link source is not a signal-passthrough node");
  const int k = d->second->kind;
  if (!(k == GB_FX_GAIN || k == GB_FX_LIMITER || k == GB_FX_COMPRESSOR) || index < 0 || index > (k == GB_FX_GAIN ? 0 : 1))
    return fail(e, GB_EINVAL, "link target must be a gain, limiter or compressor parameter");
  for (auto& l : e->links)
    if (l.dst == dst) return fail(e, GB_EINVAL, "target already has a control link");
  e->links.push_back({src, dst, index});
  return 0;
}

int go_finalize(go_engine* e) {
  if (!e) return GB_EINVAL;
  // cycle check (the reference has none, orchestrator.rs:264; a cycle would never terminate)
  std::map<uint32_t, int> color;
  std::vector<std::pair<uint32_t, size_t>> st;
  st.push_back({GB_MAIN_MIXER, 0});
  color[GB_MAIN_MIXER] = 1;
  while (!st.empty()) {
    auto& top = st.back();
    auto it = e->patches.find(top.first);
    if (it == e->patches.end() || top.second >= it->second.size()) {
      color[top.first] = 2;
      st.pop_back();
      continue;
    }
    uint32_t c = it->second[top.second++];
    if (color[c] == 1) return fail(e, GB_EGRAPH, "patch graph has a cycle");
    if (color[c] == 0) {
      color[c] = 1;
      st.push_back({c, 0});
    }
  }
  e->finalized = true;
  return 0;
}

int go_push_events(go_engine* e, const gb_event* ev, size_t n) {
  if (!e || (!ev && n)) return GB_EINVAL;
  for (size_t i = 0; i < n; ++i) {
    if (ev[i].frame < e->pos) return fail(e, GB_EINVAL, "event frame is in the past");
    if (!e->store.count(ev[i].uid)) return fail(e, GB_ENOENT, "event targets unknown uid");
  }
  e->events.insert(e->events.end(), ev, ev + n);
  std::stable_sort(e->events.begin(), e->events.end(),
                   [](const gb_event& a, const gb_event& b) { return a.frame < b.frame; });
  return 0;
}

// One frame of Orchestrator::gather_audio (orchestration/src/orchestrator.rs:367-470).
static Stereo gather_frame(go_engine* e, int64_t frame) {
  struct Entry {
    bool collect;
    uint32_t uid;
    Stereo acc;
  };
  std::vector<Entry> stack;
  Stereo sum;
  stack.push_back({false, GB_MAIN_MIXER, Stereo()});
  while (!stack.empty()) {
    Entry en = stack.back();
    stack.pop_back();
    auto it = e->store.find(en.uid);
    if (it == e->store.end()) continue;
    Entity* ent = it->second.get();
    if (!en.collect) {
      if (ent->memo_frame == frame) {
        sum.l += ent->memo.l;
        sum.r += ent->memo.r;
      } else if (ent->is_instrument()) {
        ent->tick(frame);
        Stereo v = ent->value();
        ent->memo_frame = frame;
        ent->memo = v;
        sum.l += v.l;
        sum.r += v.r;
      } else {
        stack.push_back({true, en.uid, sum});
        sum = Stereo();
        auto p = e->patches.find(en.uid);
        if (p != e->patches.end())
          for (uint32_t src : p->second) stack.push_back({false, src, Stereo()});
      }
    } else {
      Stereo v = ent->transform_audio(sum);
      ent->memo_frame = frame;
      ent->memo = v;
      sum.l = en.acc.l + v.l;
      sum.r = en.acc.r + v.r;
    }
  }
  return sum;
}

int go_render_block(go_engine* e, double* out, size_t frames, size_t* done) {
  if (!e || (!out && frames)) return GB_EINVAL;
  if (!e->finalized) return fail(e, GB_ESTATE, "engine is not finalized");
  size_t ei = 0;
  for (size_t i = 0; i < frames; ++i) {
    int64_t n = e->pos;
    while (ei < e->events.size() && e->events[ei].frame <= n) {
      const gb_event& ev = e->events[ei++];
      Entity* ent = e->store[ev.uid].get();
      switch (ev.type) {
        case GB_EV_NOTE_ON: ent->note_on(n, ev.a, ev.b); break;
        case GB_EV_NOTE_OFF: ent->note_off(n, ev.a); break;
        case GB_EV_CONTROL: ent->control(ev.a, ev.value); break;
        case GB_EV_SET_PARAM: ent->set_param(ev.a, ev.value); break;
        default: break;
      }
    }
    // controllers work once per buffer (handle_work, orchestrator.rs:631-708; GB_CONTROL_PERIOD frames):
    // a signal-passthrough source hands the magnitude of its latest output to its link targets
    if (n > 0 && n % GB_CONTROL_PERIOD == 0) {
      for (auto& l : e->links) {
        Entity* src = e->store[l.src].get();
        if (src->memo_frame != n - 1) continue;  // not reached from the main mixer: it never rendered
        double v = std::fabs(0.5 * (src->memo.l + src->memo.r));
        e->store[l.dst]->control(l.index, v > 1.0 ? 1.0 : v);
      }
    }
    Stereo s = gather_frame(e, n);
    out[2 * i] = s.l;
    out[2 * i + 1] = s.r;
    e->pos++;
  }
  e->events.erase(e->events.begin(), e->events.begin() + (long)ei);
  if (done) *done = frames;
  return 0;
}

// orchestration/src/helpers.rs:78,90-91 — (sample * i16::MAX as f64) as i16 (truncate, saturate; NaN -> 0).
void go_pcm16(const double* in, int16_t* out, size_t n_values) {
  for (size_t i = 0; i < n_values; ++i) {
    double v = in[i] * 32767.0;
    int16_t q;
    if (v != v) q = 0;
    else if (v >= 32767.0) q = 32767;
    else if (v <= -32768.0) q = -32768;
    else q = (int16_t)v;
    out[i] = q;
  }
}

int go_render_pcm16(go_engine* e, int16_t* out, size_t frames, size_t* done) {
  std::vector<double> tmp(frames * 2);
  int rc = go_render_block(e, tmp.data(), frames, done);
  if (rc) return rc;
  go_pcm16(tmp.data(), out, frames * 2);
  return 0;
}

int64_t go_position(const go_engine* e) { return e ? e->pos : -1; }

// MMA (DLS level 2) concave / convex transforms — orchestration/src/util.rs:4-21.  Dead code in the
// reference snapshot (and not on this oracle's render path: see docs/ORACLE_SPEC.md §3 for the envelope
// shapes actually used), restated because the reference's own tests pin their bounds (util.rs:286-318).
double go_mma_concave(double x) {
  if (x > 1.0 - std::pow(10.0, -12.0 / 5.0)) return 1.0;
  return -(5.0 / 12.0) * std::log10(1.0 - x);
}
double go_mma_convex(double x) {
  if (x < std::pow(10.0, -12.0 / 5.0)) return 0.0;
  return 1.0 + (5.0 / 12.0) * std::log10(x);
}

// Known-answer helpers exported for tests (unit-level checks of the restated formulas).
double go_tune_ratio(int semitones, double cents) {  // settings/src/patches.rs:255-258
  return std::pow(2.0, ((double)semitones * 100.0 + cents) / 1200.0);
}
double go_note_hz(int key) { return note_hz(key); }
double go_pct_to_hz(double pct) { return pct_to_hz(pct); }
double go_envelope_level(const gb_envelope_params* p, double sr, int64_t n_on, int64_t n_off, int64_t n) {
  EnvShape s;
  s.set(*p, sr);
  EnvState st;
  st.l_on = 0.0;
  st.l_off = n_off < kHeld ? env_pre(s, n_on, 0.0, n_off) : 0.0;
  return env_level(s, n_on, n_off, st, n);
}
void go_lp24_coefficients(double cutoff, double ripple, double sr, double* out10) {
  Lp24Ripple rp;
  rp.set(ripple);
  Lp24Coef c;
  lp24_coefficients(rp, cutoff, sr, &c);
  for (int i = 0; i < 2; ++i) {
    out10[5 * i + 0] = c.b0[i]; out10[5 * i + 1] = c.b1[i]; out10[5 * i + 2] = c.b2[i];
    out10[5 * i + 3] = c.a1[i]; out10[5 * i + 4] = c.a2[i];
  }
}
void go_rbj_coefficients(int kind, double cutoff, double p2, double sr, double* out5) {
  BiquadCoef c = rbj(kind, cutoff, p2, sr);
  out5[0] = c.b0; out5[1] = c.b1; out5[2] = c.b2; out5[3] = c.a1; out5[4] = c.a2;
}

}  // extern "C"
