// engine.cu — host-side block scheduler + C ABI (include/groove_b200.h) of the B200 renderer.
//
// What the reference does per FRAME (Orchestrator::tick -> handle_work + gather_audio,
// orchestration/src/orchestrator.rs:856-877,367-470) this engine does per CHUNK of up to
// `max_block` frames:
//   1. gb_finalize() snapshots the patch graph once into a topologically sorted plan — the
//      reference author's own TODO (orchestrator.rs:357-359);
//   2. note events are resolved on the host into per-voice event lists (voice allocation only
//      needs integer frame arithmetic), control events into per-effect parameter segments;
//   3. all voices of one instrument type render in ONE kernel launch (voice_kernels.cuh), each
//      instrument into its node buffer; effects run in plan order (fx_kernels.cuh);
//   4. the main mixer's buffer is the result (f64 stereo in HBM), copied to the caller.
// There is no CPU fallback: every DSP operation happens in a CUDA kernel of this library.

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <chrono>

#include <cuda_runtime.h>
#include <unistd.h>

#include "../../include/groove_b200.h"
#include "fx_kernels.cuh"
#include "voice_kernels.cuh"

using namespace gbk;

namespace {

constexpr int kVoiceWarps = 8;  // warps per voice CTA
constexpr int kRestSmemMax = 96 * 1024;  // dynamic shared memory cap of welsh_rest_kernel (tiles + cached voice state)
constexpr uint32_t kDefaultMaxBlock = 1u << 16;
// welsh_solo_kernel: row staging tiles + parking columns + one state slot and one instrument record per warp
constexpr int kSoloSmemBytes = (int)SoloSmem<kVoiceWarps>::kBytes;
constexpr int kSoloTickets = 64;  // ticket words in front of the per-voice progress counters (one per launch of a chunk)

std::string g_create_err;

// Host staging is rotated over kStageSlots chunks so that the host can enqueue chunk i+1 .. i+3 while
// chunk i's uploads are still in flight (no stream synchronisation per chunk).
constexpr int kStageSlots = 4;

template <typename T>
struct DevBuf {  // growable device array with pinned host mirrors (one per staging slot) for uploads
  T* d = nullptr;
  T* h = nullptr;  // the mirror of the current staging slot
  T* hs[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
  size_t cap = 0;
  int slot = 0;
  ~DevBuf() { release(); }
  void release() {
    if (d) cudaFree(d);
    for (T*& p : hs) {
      if (p) cudaFreeHost(p);
      p = nullptr;
    }
    d = nullptr; h = nullptr; cap = 0;
  }
  void use_slot(int s) {
    slot = s;
    h = hs[s];
  }
  bool reserve(size_t n) {
    if (n <= cap) return true;
    size_t ncap = std::max<size_t>(n, cap * 2 + 16);
    cudaDeviceSynchronize();  // growth is rare; uploads from the old mirrors may still be in flight
    T* nd = nullptr;
    T* nh[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
    bool ok = cudaMalloc(&nd, ncap * sizeof(T)) == cudaSuccess;
    for (int i = 0; ok && i < kStageSlots; ++i) ok = cudaMallocHost(&nh[i], ncap * sizeof(T)) == cudaSuccess;
    if (!ok) {
      if (nd) cudaFree(nd);
      for (T* p : nh)
        if (p) cudaFreeHost(p);
      return false;
    }
    if (d) cudaFree(d);
    for (int i = 0; i < kStageSlots; ++i) {
      if (hs[i]) cudaFreeHost(hs[i]);
      hs[i] = nh[i];
    }
    d = nd; cap = ncap; h = hs[slot];
    return true;
  }
};

inline uint64_t h_cycles_to_q(double c) {
  c -= std::floor(c);
  double r = c * 18446744073709551616.0;
  if (!(r < 18446744073709551616.0)) return 0;
  return (uint64_t)r;
}
inline int64_t frames_of(double seconds, double sr) {
  if (!(seconds > 0.0)) return 0;
  return (int64_t)std::llround(seconds * sr);
}
inline double note_hz(int key) { return 440.0 * std::exp2(((double)key - 69.0) / 12.0); }
inline double pct_to_hz(double pct) { return 25.0 * std::exp2(pct * 9.6438561897747243); }
inline double clamp01(double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); }
inline double denormalize_q(double v) { return v * v * 10.0 + 0.707; }
inline void dca_gains(double gain, double pan, double* gl, double* gr) {
  double a = 0.5 * (pan + 1.0), b = 0.5 * (pan - 1.0);
  *gl = gain * (1.0 - a * a);
  *gr = gain * (1.0 - b * b);
}
// host copies of dsp.cuh's env_pre / env_level (levels at note events of a GB_INST_ENVELOPE device)
inline double h_env_pre(const EnvShape& s, int64_t n_on, double l_on, int64_t n) {
  const int64_t k = n - n_on;
  if (k < s.na) {
    const double t = (double)k / (double)s.na;
    return l_on + (1.0 - l_on) * (t * (2.0 - t));
  }
  const int64_t k2 = k - s.na;
  if (k2 < s.nd) {
    const double u = 1.0 - (double)k2 / (double)s.nd;
    return s.sustain + (1.0 - s.sustain) * (u * u);
  }
  return s.sustain;
}
inline double h_env_level(const EnvShape& s, int64_t n_on, int64_t n_off, double l_on, double l_off, int64_t n) {
  if (n < n_on) return 0.0;
  if (n < n_off) return h_env_pre(s, n_on, l_on, n);
  const int64_t k = n - n_off;
  if (k < s.nr) {
    const double u = 1.0 - (double)k / (double)s.nr;
    return l_off * (u * u);
  }
  return 0.0;
}
inline EnvShape make_shape(const gb_envelope_params& p, double sr) {
  EnvShape s;
  s.na = frames_of(p.attack, sr);
  s.nd = frames_of(p.decay, sr);
  s.nr = frames_of(p.release, sr);
  s.sustain = clamp01(p.sustain);
  s.inv_na = s.na > 0 ? 1.0 / (double)s.na : 0.0;
  s.inv_nd = s.nd > 0 ? 1.0 / (double)s.nd : 0.0;
  s.inv_nr = s.nr > 0 ? 1.0 / (double)s.nr : 0.0;
  return s;
}
inline Lp24Ripple make_ripple(double ripple) {
  Lp24Ripple r;
  double sg = std::sinh(ripple);
  double cg = std::cosh(ripple);
  cg *= cg;
  r.c0 = 1.0 / (cg - 0.85355339059327376220);
  r.c2 = 1.0 / (cg - 0.14644660940672623780);
  r.c1k = r.c0 * sg * 1.84775906502257351225;
  r.c3k = r.c2 * sg * 0.76536686473017954345;
  return r;
}
inline void host_lp24(const Lp24Ripple& rp, double cutoff, double sr, SecCoef* s1, SecCoef* s2) {
  double fc = cutoff;
  if (fc > 0.49 * sr) fc = 0.49 * sr;
  if (fc < 1.0) fc = 1.0;
  double k = std::tan(3.141592653589793238462643383279 * fc / sr);
  double kk = k * k;
  double c1 = k * rp.c1k, c3 = k * rp.c3k;
  double a0 = 1.0 / (c1 + kk + rp.c0);
  s1->a1 = 2.0 * (rp.c0 - kk) * a0;
  s1->a2 = (c1 - kk - rp.c0) * a0;
  s1->b0 = a0 * kk;
  a0 = 1.0 / (c3 + kk + rp.c2);
  s2->a1 = 2.0 * (rp.c2 - kk) * a0;
  s2->a2 = (c3 - kk - rp.c2) * a0;
  s2->b0 = a0 * kk;
}
// RBJ cookbook (doc/Audio-EQ-Cookbook.txt:74-198), a0-normalised.
BiquadCoefs host_rbj(int kind, double cutoff, double p2, double sr) {
  const double kTwoPi = 6.283185307179586476925286766559;
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
      double alpha = sn * std::sinh(0.34657359027997264 * bw * w0 / sn);
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
      double alpha = sn / 2.0 * 1.41421356237309504880;
      double t = 2.0 * std::sqrt(A) * alpha;
      b0 = A * ((A + 1.0) - (A - 1.0) * cs + t);
      b1 = 2.0 * A * ((A - 1.0) - (A + 1.0) * cs);
      b2 = A * ((A + 1.0) - (A - 1.0) * cs - t);
      a0 = (A + 1.0) + (A - 1.0) * cs + t;
      a1 = -2.0 * ((A - 1.0) + (A + 1.0) * cs);
      a2 = (A + 1.0) + (A - 1.0) * cs - t;
    } break;
    default: {
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
  BiquadCoefs c;
  c.b0 = b0 / a0; c.b1 = b1 / a0; c.b2 = b2 / a0; c.a1 = a1 / a0; c.a2 = a2 / a0;
  return c;
}
inline uint64_t step_to_q32(double step) {
  double r = step * 4294967296.0;
  if (!(r >= 1.0)) return 1;
  if (r > 1.8e19) r = 1.8e19;
  return (uint64_t)r;
}
inline int64_t frames_until_end(size_t len, uint64_t step_q) {
  unsigned __int128 target = (unsigned __int128)len << 32;
  unsigned __int128 k = (target + step_q - 1) / step_q;
  if (k > (unsigned __int128)kHeld) return kHeld;
  return (int64_t)k;
}

// ---- host-side voice allocation (integer frames only) -------------------------------------
struct Slot {
  int key = -1;
  int64_t on_frame = kNever;
  int64_t idle_at = kNever;
  bool held = false;
};
struct SlotStore {
  std::vector<Slot> slots;
  int note_on(int64_t f, int key) {
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
    slots[pick].key = key;
    slots[pick].on_frame = f;
    slots[pick].idle_at = kHeld;
    slots[pick].held = true;
    return pick;
  }
};

struct SampleDev {
  double* d = nullptr;
  size_t n = 0;
  int channels = 1;
  double sr = 44100.0, root_hz = 0.0;
};
struct SampleVoiceHost {
  int64_t n_on = kNever, n_end = kNever;
  uint64_t step_q = 0;
  int sample = -1;
};

struct Node {
  uint32_t uid = 0;
  int kind = 0;
  bool is_inst = false;
  std::vector<uint32_t> sources;  // uids, patch order
  int order = -1;                 // position in the plan (-1 = unreachable)
  int level = 0;                  // longest path from a leaf (sources and control-link sources are lower)
  Node* post = nullptr;           // IIR stage: the memoryless effect fused behind it (writes that node's buffer)
  bool fused_away = false;        // memoryless effect that runs as some IIR stage's post-op
  double2* buf = nullptr;         // chunk output (max_block frames)
  double2* scratch = nullptr;     // pre-reduction when > kMaxSources inputs / partial sums
  const double2** d_src_table = nullptr;  // device table of source buffers (> kMaxSources inputs)
  const double2** d_src_table_fused = nullptr;  // the same with split instruments replaced by their partial buffers
  const double2** d_src_table_fused_alt = nullptr;  // ... by their partial buffers of odd chunks (gb_engine::overlap)
  double2* scratch_alt = nullptr;  // split Welsh instrument: the partial buffers of odd chunks (gb_engine::overlap)
  const double2** d_src_table_vr[2] = {nullptr, nullptr};  // consumer of a voice-range engine: its other sources + the range partials, per chunk parity
  int n_src_vr = 0;
  bool fused_all_inst = false;  // every live source of this consumer is a split instrument whose partials it sums itself
  int n_src_fused = 0;
  int consumers = 0;           // plan nodes that read this node's buffer
  bool fuse_partials = false;  // split instrument whose only consumer sums its partials itself (no reduce pass)
  // --- instruments ---
  gb_welsh_params wp;
  gb_fm_params fp;
  gb_sampler_params sp;
  gb_toy_source_params tp;
  gb_oscillator_source_params osp;
  gb_envelope_source_params esp;
  struct EnvHost {  // GB_INST_ENVELOPE: note state tracked on the host (integer frames + two levels)
    int64_t n_on = kNever, n_off = kNever;
    double l_on = 0.0, l_off = 0.0;
  } envh;
  int table_index = -1;  // index in the welsh/fm instrument table
  int voice0 = 0, nvoices = 0;
  int partial_count = 0;  // partial output buffers in `scratch` (0 = renders straight into `buf`)
  bool unit_gain = false; // this chunk renders at unit DCA gain; a segmented OP_DCA pass applies gain/pan
  std::vector<std::vector<SampleVoiceHost>> done_plays;  // per sample voice: plays that ended inside this chunk
  int64_t release_frames = 0;
  SlotStore store;
  std::vector<SampleDev> samples;
  std::vector<SampleVoiceHost> svoices;
  int key_to_voice[128];
  // --- effects ---
  double p[4] = {0, 0, 0, 0};  // raw parameters in control-index order
  BiquadState* d_bq = nullptr;
  Lp24State* d_lp = nullptr;
  int delay_frames = 0;
  double2* hist[2] = {nullptr, nullptr};
  int hist_cur = 0;
  ChorusTaps taps;
  ReverbDesc rv;
  double* rv_comb_out = nullptr;  // [2][4][max_block]
  double* rv_ap_out[2] = {nullptr, nullptr};  // [2][max_block] each
};

}  // namespace

struct gb_engine {
  int device = 0;
  double sr = 44100.0;
  uint32_t max_block = kDefaultMaxBlock;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;               // device -> host result copies, overlapped with the next chunks
  cudaEvent_t stage_done[kStageSlots] = {};          // chunk i's work enqueued on `stream` (guards staging slot i % kStageSlots)
  cudaEvent_t d2d_done[kStageSlots] = {}, d2h_done[kStageSlots] = {};
  // Overlapped mixdown (engines whose instruments are all split Welsh instruments summed by their consumer, e.g.
  // BASELINE config 4): the grouped voice kernels run on `vstream` and write partial buffers that alternate
  // with the chunk's parity, while `stream` still sums / mixes / copies out the previous chunk — the HBM-bound
  // table sum then runs on the SM slots the voice kernels leave free instead of between them.
  cudaStream_t vstream = nullptr;
  cudaEvent_t v_done[2] = {}, p_done[2] = {}, call_start = nullptr;
  bool overlap = false;          // decided by gb_finalize
  bool overlap_enabled = true;   // GB_OVERLAP=0
  CtaWork* d_wwork_alt = nullptr;        // the grouped CTA list with the odd-chunk output buffers
  PartialDesc* d_partials_alt = nullptr; // e->partials with the odd-chunk buffers
  int parity = 0;                        // partial-buffer set of the chunk being enqueued
  // Voice-range resting chunks (welsh_rest_vr_kernel): when every grouped CTA rests, the voices are cut into
  // ranges of 14 regardless of instrument boundaries so that all SMs carry the same load.
  bool vr_ok = false;                    // decided by gb_finalize (needs `overlap`)
  int vr_class = -1;                     // the instruments' common rest class
  int vr_sweep_class = -1;               // the instruments' common sweep class (-1: sweeping chunks keep instrument CTAs)
  int vr_osc = 0;                        // welsh_rest_vr16_kernel's OSC: 1 = every instrument's oscillators are symmetric +-c waveforms, 2 = ... and oscillator 2 a square
  int vr_count = 0;                      // ranges (CTAs)
  VrWork* d_vr_work[2] = {nullptr, nullptr};
  bool chunk_vr = false;                 // this chunk's resting voices went through the ranges
  bool chunk_all_idle = false;           // this chunk: every grouped Welsh CTA idle (its partial buffer holds zeros)
  Rest16Table* d_rest16 = nullptr;       // per Welsh instrument: the tables of welsh_rest_vr16_kernel (16 frames per lane)
  double2* ring = nullptr;                           // pinned: kStageSlots x max_block frames (host-buffer renders)
  uint64_t chunk_seq = 0;
  bool finalized = false;
  int64_t pos = 0;
  // gb_set_lookahead: frames rendered ahead of what gb_render_block has handed out (host buffer)
  size_t lookahead = 0;
  std::vector<double> ahead;       // interleaved L,R
  size_t ahead_n = 0, ahead_off = 0;
  uint32_t next_uid = 2;
  int num_sms = 148;
  // Developer switches (A/B measurements, path-agreement tests): read from the environment ONCE, in
  // gb_create; nothing on the render path calls getenv.
  struct Options {
    int cta_target_mult = 2;      // GB_CTA_MULT: CTAs per SM the voice work lists aim for
    int vpc = 0;                  // GB_VPC: force the voices-per-CTA split (0 = automatic)
    double knot_max_rate = -1.0;  // GB_KNOT_MAX_RATE: < 0 = kKnotMaxRate; 0 = exact coefficients every frame
    bool lti = true;              // GB_LTI=0: no time-invariant blocks
    bool rest_kernel = true;      // GB_REST_KERNEL=0
    bool sweep_kernel = true;     // GB_SWEEP_KERNEL=0
    bool sync_kernels = true;     // GB_SYNC_KERNELS=0: hard-sync patches stay on the general kernel
    bool osc_sym = true;          // GB_OSC_SYM=0: no sign-bit oscillator form in welsh_rest_vr16_kernel
    bool rest16 = true;           // GB_REST16=0: voice-range chunks keep 8 frames per lane (welsh_rest_vr_kernel)
    int rest_vr = -1;             // GB_REST_VR: 1 = voice-range resting chunks whenever possible, 0 = never, -1 = when they even out the SM load
    int rest_nv = 2;              // GB_REST_NV=4: welsh_rest_kernel with four voices in lockstep per warp (one CTA per SM)
    bool rest_tp = true;          // GB_REST_TP=0: no time-parallel resting kernel (and no CTAs below 8 voices)
    int min_cut_voices = 256;     // GB_MIN_CUT_VOICES: chunk cuts only for engines with at least this many Welsh voices
  } opt;
  std::map<uint32_t, std::unique_ptr<Node>> nodes;
  std::vector<Node*> by_uid;      // nodes[uid].get(), O(1): every event of a chunk looks its target up
  std::vector<Node*> plan;  // reachable nodes, sources before consumers
  std::vector<gb_event> events;
  std::string err;

  // voice tables
  std::vector<WelshInst> h_winst;
  std::vector<FmInst> h_finst;
  WelshInst* d_winst = nullptr;
  FmInst* d_finst = nullptr;
  WelshVoice* d_wvoice = nullptr;
  FmVoice* d_fvoice = nullptr;
  int n_wvoice = 0, n_fvoice = 0;
  bool winst_dirty = true, finst_dirty = true;
  DevBuf<CtaWork> wwork, fwork;
  DevBuf<WarpItem> witems, fitems;  // solo-warp work items (instruments with fewer voices than a CTA has warps)
  int n_wwork = 0, n_fwork = 0;
  int n_wwork_grouped = 0;  // wwork[0 .. n_wwork_grouped) are grouped CTAs, the rest solo CTAs
  struct Link {                   // gb_link_control: signal-passthrough source -> effect parameter
    uint32_t src, dst;
    int index;
    SegParam* d_table = nullptr;  // the target's parameter table of the current chunk (device-built)
    double* d_state = nullptr;    // 2 x {control value, last l, last r, have-last}: ping-pong per chunk
    int cur = 0;
  };
  std::vector<Link> links;
  std::vector<Node*> wwork_node;  // instrument of each grouped Welsh CTA
  std::vector<char> wwork_zero;   // ... its output buffer currently holds zeros (idle CTAs are not launched)
  std::vector<char> wwork_zero_alt;  // ... the odd-chunk buffer (gb_engine::overlap)
  // solo Welsh items (instruments with fewer voices than a CTA has warps): with enough of them the host
  // classifies every (voice, sub-chunk) and welsh_solo_kernel walks the resulting job list (voice_kernels.cuh)
  bool solo_mega = false;
  std::vector<Node*> witem_node;  // instrument of each solo Welsh item
  std::vector<int> witem_slot;    // ... and its voice slot in that instrument
  struct SoloDirty { int lo = 0, hi = 0; };  // frames of the item's output buffer that may hold non-zero audio
  std::vector<SoloDirty> solo_dirty;
  DevBuf<SoloItem> sitems;
  DevBuf<SoloJob> sjobs;
  DevBuf<ZeroRange> szero;
  int* d_solo_sync = nullptr;     // [0, kSoloTickets) = job tickets (one per launch), [kSoloTickets + voice] = per-voice progress counter
  int solo_sub = kSoloSubDefault; // GB_SOLO_SUB: frames per sub-chunk (a multiple of kBlockFrames)
  int solo_lockstep = 0;          // GB_SOLO_LOCKSTEP: bit 0 = rest / sweep / exact jobs run their blocks in lockstep, bit 1 = general jobs too
                                  // (measured on config 5: 6.47 ms free-running, 6.68 / 7.84 ms with bit 0 / bits 0+1: profiles/r2_cfg5_scheduler_sweep.txt)
  bool fuse_post_ops = true;      // GB_FX_FUSE
  int fx_minb = 2;                // GB_FX_MINB: resident CTAs per SM the batched IIR kernels are compiled for (1: 174 registers, 2: 128, 3: 80)
  bool batch_fx = true;           // GB_FX_BATCH: independent effects of one kind and level share a launch
  DevBuf<FxDesc> fxdescs;
  bool solo_waves = false;        // GB_SOLO_WAVES=1: one launch per sub-chunk (stream order instead of the progress counters)
  int* d_solo_fault = nullptr;    // {flag, job, voice, need, have}: set by the kernel's dependency watchdog
  bool solo_ran = false;          // a job-list launch happened since the fault words were last checked
  int min_solo_items = 64;        // GB_SOLO_MIN: fewer solo items than this keep welsh_kernel<.., SOLO> (GB_SOLO_MIN=0: always the job list)
  DevBuf<int> widx;               // per chunk: grouped Welsh CTAs sorted into resting (4 variants) and general
  std::vector<int> widx_on_device;  // what widx.d currently holds (unchanged lists are not uploaded again)
  bool wev_empty_on_device = false, fev_empty_on_device = false;  // the event offset tables on the device are all zero
  DevBuf<VoiceEvent> wev, fev;
  DevBuf<int> wev_off, fev_off;
  DevBuf<SamplePlay> plays;
  DevBuf<SegParam> segs;         // per-chunk parameter segment tables of all effects / instrument DCAs
  DevBuf<PartialDesc> partials;  // instruments split over several CTAs (static after finalize)
  int n_partials = 0;
  DevBuf<PartialDesc> partials_nf;  // ... without those whose consumer sums the partials itself (fuse_partials)
  int n_partials_nf = 0;
  bool fused_sums = false;          // this chunk: consumers read partial buffers directly
  bool fused_sums_enabled = true;   // GB_FUSED_SUMS=0 switches the shortcut off (A/B measurements)
  bool chunk_cuts = true;           // GB_CHUNK_CUTS=0: chunks of max_block regardless of voice transitions
  std::vector<void*> allocations;  // everything cudaMalloc'ed at finalize/load time (slabs of dev_alloc, samples)
  char* slab_ptr = nullptr;        // bump pointer into the current slab
  size_t slab_left = 0, slab_next = 0;

  double2* last_out = nullptr;  // main mixer buffer of the last chunk
  size_t last_frames = 0;
  double2* d_full = nullptr;    // whole-call result (render_device / read_last), grown on demand
  size_t full_cap = 0;
  size_t full_frames = 0;       // frames of d_full that hold the last device-resident render
  short2* d_pcm = nullptr;
  size_t pcm_cap = 0;

  // stats
  gb_stats stats;
  bool timing = false;
  // CUDA-event pairs recorded around launches / render calls on the engine stream; resolved
  // lazily (gb_get_stats) so that timing never serialises the stream.
  struct TimedSpan { cudaEvent_t a, b; int what; bool closed; cudaStream_t st; };  // what: 0 = fx kernel, 1 = voice kernel, 2 = render call
  std::vector<TimedSpan> spans;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> event_pool;
};

namespace {

int fail(gb_engine* e, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (e) e->err = buf;
  else g_create_err = buf;
  return code;
}
#define CUDA_TRY(e, expr)                                                                        \
  do {                                                                                           \
    cudaError_t _rc = (expr);                                                                    \
    if (_rc != cudaSuccess)                                                                      \
      return fail(e, GB_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_rc), __FILE__, __LINE__); \
  } while (0)

// Plan-time device memory comes out of slabs (a batch of 8192 one-voice instruments would otherwise be 8192
// cudaMalloc / cudaFree pairs per engine): bump allocation, 256-byte aligned, freed with the engine.
template <typename T>
int dev_alloc(gb_engine* e, T** out, size_t count, bool zero = true) {
  const size_t bytes = (std::max<size_t>(count, 1) * sizeof(T) + 255) & ~(size_t)255;
  if (bytes > e->slab_left) {
    e->slab_next = std::min<size_t>(std::max<size_t>(2 * e->slab_next, (size_t)32 << 20), (size_t)1 << 30);  // 32 MB, doubling up to 1 GB
    const size_t slab = std::max(bytes, e->slab_next);
    void* p = nullptr;
    CUDA_TRY(e, cudaMalloc(&p, slab));
    // Zeroed once, here, and drained: a per-allocation memset on the engine stream is not ordered against the
    // synchronous table copies (legacy stream) that follow most allocations and could land after them.
    CUDA_TRY(e, cudaMemset(p, 0, slab));
    CUDA_TRY(e, cudaDeviceSynchronize());
    e->allocations.push_back(p);
    e->slab_ptr = (char*)p;
    e->slab_left = slab;
  }
  void* p = e->slab_ptr;
  e->slab_ptr += bytes;
  e->slab_left -= bytes;
  (void)zero;  // slabs are zeroed when they are created and never handed out twice
  *out = (T*)p;
  return 0;
}

Node* find(gb_engine* e, uint32_t uid) {  // uids are dense (1 = main mixer, then in order of creation)
  return uid < e->by_uid.size() ? e->by_uid[uid] : nullptr;
}
void index_node(gb_engine* e, Node* n) {
  if (n->uid >= e->by_uid.size()) e->by_uid.resize((size_t)n->uid + 1, nullptr);
  e->by_uid[n->uid] = n;
}

void welsh_inst_from_params(const Node& n, double sr, const gb_engine::Options& opt, WelshInst* I) {
  const gb_welsh_params& p = n.wp;
  const double top = 1.0 - 1.0 / 9007199254740992.0;
  memset(I, 0, sizeof *I);
  I->amp = make_shape(p.amp_envelope, sr);
  I->filt = make_shape(p.filter_envelope, sr);
  I->w1 = p.oscillator_1.waveform;
  I->w2 = p.oscillator_2.waveform;
  I->wl = p.lfo.waveform;
  I->sync = p.oscillator_2_sync ? 1 : 0;
  I->routing = p.lfo_routing;
  I->uid = (int)n.uid;
  I->voice0 = n.voice0;
  I->duty1_q = h_cycles_to_q(std::min(clamp01(p.oscillator_1.pulse_width), top));
  I->duty2_q = h_cycles_to_q(std::min(clamp01(p.oscillator_2.pulse_width), top));
  I->dutyl_q = h_cycles_to_q(std::min(clamp01(p.lfo.pulse_width), top));
  I->lfo_dq = h_cycles_to_q(p.lfo.frequency / sr);
  I->duty1 = p.oscillator_1.pulse_width;
  I->duty2 = p.oscillator_2.pulse_width;
  I->mix = p.oscillator_mix;
  I->depth = p.lfo_depth;
  if (p.filter_cutoff_end != 0.0) {
    I->filter_mode = FILTER_ENVELOPE;
    I->cut_a = p.filter_cutoff_start;
    I->cut_b = (1.0 - p.filter_cutoff_start) * p.filter_cutoff_end;
  } else if (p.lfo_routing == GB_LFO_FILTER_CUTOFF) {
    I->filter_mode = FILTER_LFO;
    I->cut_a = p.filter_cutoff_start;
    I->cut_b = 0.0;
  } else {
    I->filter_mode = FILTER_FIXED;
  }
  I->rp = make_ripple(p.filter_passband_ripple);
  host_lp24(I->rp, p.filter_cutoff_hz, sr, &I->fixed1, &I->fixed2);
  double pan = p.voice_dca.pan + p.dca.pan;
  pan = pan < -1.0 ? -1.0 : (pan > 1.0 ? 1.0 : pan);
  dca_gains(p.voice_dca.gain * p.dca.gain, pan, &I->gl, &I->gr);
  if (n.unit_gain) { I->gl = 1.0; I->gr = 1.0; }
  I->pi_over_sr = 3.141592653589793238462643383279 / sr;
  I->sr = sr;
  auto shape = [](int wf, uint64_t duty_q, OscShape* o) {
    memset(o, 0, sizeof *o);
    o->thresh = 1ull << 63;
    switch (wf) {
      case GB_WAVE_SINE: o->kind = 1; break;
      case GB_WAVE_NOISE: o->kind = 2; break;
      case GB_WAVE_SQUARE: o->b_lo = 1.0; o->b_hi = -1.0; break;
      case GB_WAVE_PULSE_WIDTH: o->thresh = duty_q; o->b_lo = 1.0; o->b_hi = -1.0; break;
      case GB_WAVE_TRIANGLE: o->a_lo = 4.0; o->b_lo = -1.0; o->a_hi = -4.0; o->b_hi = 3.0; break;
      case GB_WAVE_SAWTOOTH: o->a_lo = 2.0; o->b_lo = 0.0; o->a_hi = 2.0; o->b_hi = -2.0; break;
      case GB_WAVE_DEBUG_MAX: o->b_lo = 1.0; o->b_hi = 1.0; break;
      case GB_WAVE_DEBUG_MIN: o->b_lo = -1.0; o->b_hi = -1.0; break;
      default: break;  // none / debug-zero: 0
    }
  };
  shape(I->w1, I->duty1_q, &I->s1);
  shape(I->w2, I->duty2_q, &I->s2);
  shape(I->wl, I->dutyl_q, &I->sl);
  I->log2_25_over_sr = std::log2(25.0 / sr);
  I->u_min = 1.0 / sr;
  I->u_max = 0.49;
  I->knot_max_rate = opt.knot_max_rate >= 0.0 ? opt.knot_max_rate : kKnotMaxRate;
  for (int j = 0; j < kT; ++j) {
    uint64_t q = (uint64_t)j * I->lfo_dq;  // mod 2^64
    double ang = 6.283185307179586476925286766559 * ((double)q / 18446744073709551616.0);
    I->lfo_rot[j] = make_double2(std::cos(ang), std::sin(ang));
  }
  // oscillator pair with the mix and the [1,2) phase offset folded in (OscMix)
  auto fold = [](const OscShape& o, double mix, OscMix* m) {
    m->a_lo = mix * o.a_lo; m->b_lo = mix * (o.b_lo - o.a_lo);
    m->a_hi = mix * o.a_hi; m->b_hi = mix * (o.b_hi - o.a_hi);
  };
  fold(I->s1, I->mix, &I->m1);
  fold(I->s2, 1.0 - I->mix, &I->m2);
  // time-invariant stretches: the resting coefficient sets and what follows from them (LtiTable)
  {
    SecCoef c1 = I->fixed1, c2 = I->fixed2;
    if (I->filter_mode == FILTER_ENVELOPE) {
      double pct = I->cut_a + I->cut_b * I->filt.sustain;
      pct = pct < 0.0 ? 0.0 : (pct > 1.0 ? 1.0 : pct);
      host_lp24(I->rp, 25.0 * std::exp2(pct * 9.6438561897747243), sr, &c1, &c2);
    }
    for (int l = 0; l < 32; ++l) {
      uint64_t q = (uint64_t)(l * kT) * I->lfo_dq;
      double ang = 6.283185307179586476925286766559 * ((double)q / 18446744073709551616.0);
      I->lane_rot[l] = make_double2(std::cos(ang), std::sin(ang));
    }
    {
      uint64_t q = (uint64_t)kBlockFrames * I->lfo_dq;
      double ang = 6.283185307179586476925286766559 * ((double)q / 18446744073709551616.0);
      I->block_rot = make_double2(std::cos(ang), std::sin(ang));
    }
    I->steady_after = I->filter_mode == FILTER_ENVELOPE ? std::max(I->amp.na + I->amp.nd, I->filt.na + I->filt.nd)
                                                         : I->amp.na + I->amp.nd;
    I->amp_rest = 0.5 * I->amp.sustain;
    I->osc_flat = I->m1.a_lo == 0.0 && I->m1.a_hi == 0.0 && I->m2.a_lo == 0.0 && I->m2.a_hi == 0.0;
    auto table = [](const SecCoef& c, double (*g)[2], double (*mp)[4]) {
      double h00 = 1.0, h01 = 0.0, h10 = 0.0, h11 = 1.0;  // H = A^j, A = [[a1, 1], [a2, 0]]
      for (int j = 0; j < kT; ++j) {
        g[j][0] = h00; g[j][1] = h01;
        const double t00 = c.a1 * h00 + h10, t01 = c.a1 * h01 + h11;
        h10 = c.a2 * h00; h11 = c.a2 * h01;
        h00 = t00; h01 = t01;
      }
      mp[0][0] = h00; mp[0][1] = h01; mp[0][2] = h10; mp[0][3] = h11;
      for (int k = 1; k < 5; ++k) {
        const double* m = mp[k - 1];
        mp[k][0] = m[0] * m[0] + m[1] * m[2]; mp[k][1] = m[0] * m[1] + m[1] * m[3];
        mp[k][2] = m[2] * m[0] + m[3] * m[2]; mp[k][3] = m[2] * m[1] + m[3] * m[3];
      }
      mp[5][0] = mp[5][1] = mp[5][2] = mp[5][3] = 0.0;
    };
    I->lti.c1 = c1; I->lti.c2 = c2;
    auto scaled = [](const OscMix& m, double k) {
      OscMix r;
      r.a_lo = m.a_lo * k; r.b_lo = m.b_lo * k; r.a_hi = m.a_hi * k; r.b_hi = m.b_hi * k;
      return r;
    };
    I->m1b = scaled(I->m1, c1.b0);
    I->m2b = scaled(I->m2, c1.b0);
    I->m1bb = scaled(I->m1, c1.b0 * c2.b0);
    I->m2bb = scaled(I->m2, c1.b0 * c2.b0);
    table(c1, I->lti.g1, I->lti.mp1);
    table(c2, I->lti.g2, I->lti.mp2);
    auto square = [](const double* m, double* o) {
      o[0] = m[0] * m[0] + m[1] * m[2]; o[1] = m[0] * m[1] + m[1] * m[3];
      o[2] = m[2] * m[0] + m[3] * m[2]; o[3] = m[2] * m[1] + m[3] * m[3];
    };
    square(I->lti.mp1[4], I->mp32_1);  // the lane map to the 32nd power: one 256-frame block
    square(I->lti.mp2[4], I->mp32_2);
    for (int j = 0; j < kT; ++j) {
      I->lti.g1b[j][0] = I->lti.g1[j][0] * c2.b0;
      I->lti.g1b[j][1] = I->lti.g1[j][1] * c2.b0;
    }
    I->lti.inv_b0_2 = 1.0 / c2.b0;
    {  // released voice, filter envelope run out: cutoff back at cut_a (a fixed filter keeps its one set)
      SecCoef o1 = I->fixed1, o2 = I->fixed2;
      if (I->filter_mode == FILTER_ENVELOPE) {
        double pct = I->cut_a;
        pct = pct < 0.0 ? 0.0 : (pct > 1.0 ? 1.0 : pct);
        host_lp24(I->rp, 25.0 * std::exp2(pct * 9.6438561897747243), sr, &o1, &o2);
      }
      I->lti_off.c1 = o1; I->lti_off.c2 = o2;
      I->m1bb_off = scaled(I->m1, o1.b0 * o2.b0);
      I->m2bb_off = scaled(I->m2, o1.b0 * o2.b0);
      table(o1, I->lti_off.g1, I->lti_off.mp1);
      table(o2, I->lti_off.g2, I->lti_off.mp2);
      for (int j = 0; j < kT; ++j) {
        I->lti_off.g1b[j][0] = I->lti_off.g1[j][0] * o2.b0;
        I->lti_off.g1b[j][1] = I->lti_off.g1[j][1] * o2.b0;
      }
      I->lti_off.inv_b0_2 = 1.0 / o2.b0;
    }
    I->lti_ok = opt.lti ? 1 : 0;
    // welsh_rest_kernel variant: same preconditions as welsh_block_lti (see voice_kernels.cuh)
    const bool lin = I->s1.kind == 0 && I->s2.kind == 0 && !I->sync &&
                     (I->routing == LFO_NONE || (I->routing == LFO_AMPLITUDE && I->wl == W_SINE));
    // hard-sync patches take the same kernels in their SYNC instantiation (variants 4 / 5: general oscillator form)
    const bool lin_sync = I->s1.kind == 0 && I->s2.kind == 0 && I->sync && opt.sync_kernels &&
                          (I->routing == LFO_NONE || (I->routing == LFO_AMPLITUDE && I->wl == W_SINE));
    I->rest_class = -1;
    if (lin && I->lti_ok && (I->filter_mode == FILTER_FIXED || I->filter_mode == FILTER_ENVELOPE))
      I->rest_class = (I->routing == LFO_AMPLITUDE ? 2 : 0) + (I->osc_flat ? 1 : 0);
    else if (lin_sync && I->lti_ok && (I->filter_mode == FILTER_FIXED || I->filter_mode == FILTER_ENVELOPE))
      I->rest_class = 4 + (I->routing == LFO_AMPLITUDE ? 1 : 0);
    if (!opt.rest_kernel) I->rest_class = -1;
    // welsh_sweep_kernel: the knot path of a moving filter envelope; the frequency clamp must stay out of
    // reach (0.49 sr >= 20 kHz) so that the coefficient trajectory of a stage is smooth
    I->sweep_class = -1;
    if (lin && I->filter_mode == FILTER_ENVELOPE && I->knot_max_rate > 0.0 && 0.49 * sr >= 20000.0)
      I->sweep_class = (I->routing == LFO_AMPLITUDE ? 2 : 0) + (I->osc_flat ? 1 : 0);
    else if (lin_sync && I->filter_mode == FILTER_ENVELOPE && I->knot_max_rate > 0.0 && 0.49 * sr >= 20000.0)
      I->sweep_class = 4 + (I->routing == LFO_AMPLITUDE ? 1 : 0);
    if (!opt.sweep_kernel) I->sweep_class = -1;
    // welsh_exact_block (solo items): as above with exact coefficient sets at every frame, so no smoothness needed
    I->exact_class = lin && I->filter_mode == FILTER_ENVELOPE ? (I->routing == LFO_AMPLITUDE ? 1 : 0) : -1;
  }
}
void fm_inst_from_params(const Node& n, double sr, FmInst* I) {
  memset(I, 0, sizeof *I);
  I->car = make_shape(n.fp.carrier_envelope, sr);
  I->mod = make_shape(n.fp.modulator_envelope, sr);
  I->depth = n.fp.depth;
  I->beta = n.fp.beta;
  dca_gains(n.fp.dca.gain, n.fp.dca.pan, &I->gl, &I->gr);
  if (n.unit_gain) { I->gl = 1.0; I->gr = 1.0; }
  I->uid = (int)n.uid;
  I->voice0 = n.voice0;
}

bool span_begin(gb_engine* e, int what, cudaStream_t st = nullptr) {
  if (!e->timing) return false;
  if (!st) st = e->stream;
  std::pair<cudaEvent_t, cudaEvent_t> ev;
  if (!e->event_pool.empty()) {
    ev = e->event_pool.back();
    e->event_pool.pop_back();
  } else {
    if (cudaEventCreate(&ev.first) != cudaSuccess || cudaEventCreate(&ev.second) != cudaSuccess) return false;
  }
  if (cudaEventRecord(ev.first, st) != cudaSuccess) {  // timing is best effort: an untimed launch, not a failed one
    e->event_pool.push_back(ev);
    return false;
  }
  e->spans.push_back({ev.first, ev.second, what, false, st});
  return true;
}
void span_end(gb_engine* e, size_t index) {
  e->spans[index].closed = cudaEventRecord(e->spans[index].b, e->spans[index].st) == cudaSuccess;
}
void resolve_spans(gb_engine* e) {
  for (auto& sp : e->spans) {
    float ms = 0.f;
    if (!sp.closed || cudaEventSynchronize(sp.b) != cudaSuccess || cudaEventElapsedTime(&ms, sp.a, sp.b) != cudaSuccess)
      ms = 0.f;  // a span whose events could not be recorded contributes nothing
    if (sp.what == 1) e->stats.voice_kernel_ms += ms;
    else if (sp.what == 3) { e->stats.voice_kernel_ms += ms; e->stats.rest_kernel_ms += ms; }
    else if (sp.what == 4) { e->stats.voice_kernel_ms += ms; e->stats.sweep_kernel_ms += ms; }
    else if (sp.what == 5) { e->stats.voice_kernel_ms += ms; e->stats.solo_kernel_ms += ms; }
    else if (sp.what == 6) { e->stats.voice_kernel_ms += ms; e->stats.fm_kernel_ms += ms; }
    else if (sp.what == 0) e->stats.fx_kernel_ms += ms;
    else e->stats.render_ms += ms;
    e->event_pool.push_back({sp.a, sp.b});
  }
  e->spans.clear();
}

struct Launch {  // per-launch accounting (+ optional CUDA-event timing on the engine's stream)
  gb_engine* e;
  bool timed;
  size_t index;
  Launch(gb_engine* e_, bool voice, int special = 0, cudaStream_t st = nullptr) : e(e_) {  // special: 1 = resting, 2 = sweeping, 3 = solo job-list, 4 = FM kernel
    e->stats.kernel_launches++;
    if (voice) e->stats.voice_kernel_launches++;
    if (special == 1) e->stats.rest_kernel_launches++;
    if (special == 2) e->stats.sweep_kernel_launches++;
    if (special == 3) e->stats.solo_kernel_launches++;
    if (special == 4) e->stats.fm_kernel_launches++;
    timed = span_begin(e, special == 1 ? 3 : special == 2 ? 4 : special == 3 ? 5 : special == 4 ? 6 : voice ? 1 : 0, st);
    index = e->spans.size() - 1;
  }
  ~Launch() {
    if (timed) span_end(e, index);
  }
};

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Sum a node's inputs.  Up to kMaxSources go straight into the consuming kernel; more are
// pre-reduced (sum_table_kernel: four interleaved row sums, added in row order) into the node's scratch buffer.
// A plain sum node (mixer, signal-passthrough) with a source table needs no pass of its own: the table
// sum is written straight into its node buffer and *done is set.
int gather_sources(gb_engine* e, Node* n, int frames, SourceList* out, bool* done = nullptr) {
  std::vector<const double2*> ptrs;
  for (uint32_t s : n->sources) {
    Node* sn = find(e, s);
    if (sn && sn->order >= 0 && sn->buf) ptrs.push_back(sn->buf);
  }
  memset(out, 0, sizeof *out);
  if ((int)ptrs.size() <= kMaxSources) {
    out->n = (int)ptrs.size();
    for (int i = 0; i < out->n; ++i) out->p[i] = ptrs[i];
    return 0;
  }
  {
    // the pointer table was uploaded at finalize (sources never change afterwards)
    Launch l(e, false);
    const bool plain = done && (n->kind == GB_FX_MIXER || n->kind == GB_FX_SIGNAL_PASSTHROUGH);
    double2* dst = plain ? n->buf : n->scratch;
    if (e->fused_sums && e->overlap && e->chunk_all_idle && n->fused_all_inst)
      // a silent stretch: every partial buffer holds zeros (idle CTAs are not launched), and so does their sum
      CUDA_TRY(e, cudaMemsetAsync(dst, 0, (size_t)frames * sizeof(double2), e->stream));
    else if (e->fused_sums && e->chunk_vr && n->d_src_table_vr[e->parity])
      sum_table_kernel<<<cdiv(frames, kSumFrames), kSumRows * kSumFrames, 0, e->stream>>>(n->d_src_table_vr[e->parity], n->n_src_vr, dst, frames);
    else if (e->fused_sums && n->d_src_table_fused)
      sum_table_kernel<<<cdiv(frames, kSumFrames), kSumRows * kSumFrames, 0, e->stream>>>(
          e->parity && n->d_src_table_fused_alt ? n->d_src_table_fused_alt : n->d_src_table_fused, n->n_src_fused, dst, frames);
    else
      sum_table_kernel<<<cdiv(frames, kSumFrames), kSumRows * kSumFrames, 0, e->stream>>>(n->d_src_table, (int)ptrs.size(), dst, frames);
    if (plain) *done = true;
  }
  out->n = 1;
  out->p[0] = n->scratch;
  return 0;
}

void effect_apply_param(gb_engine* e, Node* n, int index, double raw) {
  if (index < 0 || index >= 4) return;
  n->p[index] = raw;
  (void)e;
}
// ControlValue (0..1) -> raw parameter, per effect kind (mirrors set_<field>(value.into()),
// proc-macros/src/control.rs:159-161).
double control_to_raw(const Node* n, int index, double v) {
  switch (n->kind) {
    case GB_FX_BITCRUSHER: return std::floor(v * 16.0);
    case GB_FX_LOW_PASS_24DB:
      return index == GB_CTL_FILTER_CUTOFF ? pct_to_hz(clamp01(v)) : denormalize_q(v);
    case GB_FX_LOW_PASS_12DB: case GB_FX_HIGH_PASS_12DB: case GB_FX_ALL_PASS_12DB:
      return index == GB_CTL_FILTER_CUTOFF ? pct_to_hz(clamp01(v)) : denormalize_q(v);
    case GB_FX_BAND_PASS_12DB: case GB_FX_BAND_STOP_12DB:
      return index == GB_CTL_FILTER_CUTOFF ? pct_to_hz(clamp01(v)) : v * 4.0;
    case GB_FX_PEAKING_EQ_12DB: case GB_FX_LOW_SHELF_12DB: case GB_FX_HIGH_SHELF_12DB:
      return index == GB_CTL_FILTER_CUTOFF ? pct_to_hz(clamp01(v)) : (2.0 * v - 1.0) * 24.0;
    case GB_INST_WELSH: case GB_INST_FM:
      return index == GB_CTL_INST_DCA_PAN ? 2.0 * v - 1.0 : v;
    default: return v;
  }
}
bool effect_accepts(const Node* n, int index) {
  switch (n->kind) {
    case GB_FX_GAIN: case GB_FX_BITCRUSHER: case GB_FX_REVERB: return index == 0;
    case GB_FX_LIMITER: case GB_FX_COMPRESSOR: return index == 0 || index == 1;
    case GB_FX_CHORUS: return index == GB_CTL_CHORUS_WET_DRY_MIX;
    case GB_FX_LOW_PASS_24DB: return index == 0 || index == 1;
    case GB_INST_WELSH: case GB_INST_FM: return index == 0 || index == 1;
    default:
      if (n->kind >= GB_FX_LOW_PASS_12DB && n->kind <= GB_FX_HIGH_SHELF_12DB) return index == 0 || index == 1;
      return false;
  }
}

// Kernel-ready values of an effect's (or instrument DCA's) current parameters, for one segment.
void seg_values(gb_engine* e, const Node* n, double* v) {
  for (int i = 0; i < 6; ++i) v[i] = 0.0;
  switch (n->kind) {
    case GB_FX_GAIN: v[0] = n->p[0]; break;
    case GB_FX_LIMITER: v[0] = n->p[0]; v[1] = n->p[1]; break;
    case GB_FX_BITCRUSHER: v[0] = std::exp2(std::floor(n->p[0])); break;
    case GB_FX_COMPRESSOR: v[0] = n->p[0]; v[1] = n->p[1]; break;
    case GB_FX_CHORUS: v[0] = n->p[GB_CTL_CHORUS_WET_DRY_MIX]; break;
    case GB_FX_REVERB: v[0] = n->p[0]; break;
    case GB_FX_LOW_PASS_24DB: {
      SecCoef s1, s2;
      Lp24Ripple rp = make_ripple(n->p[1]);
      host_lp24(rp, n->p[0], e->sr, &s1, &s2);
      v[0] = s1.b0; v[1] = s1.a1; v[2] = s1.a2; v[3] = s2.b0; v[4] = s2.a1; v[5] = s2.a2;
    } break;
    case GB_INST_WELSH: {
      double pan = n->wp.voice_dca.pan + n->wp.dca.pan;
      pan = pan < -1.0 ? -1.0 : (pan > 1.0 ? 1.0 : pan);
      dca_gains(n->wp.voice_dca.gain * n->wp.dca.gain, pan, &v[0], &v[1]);
    } break;
    case GB_INST_FM: dca_gains(n->fp.dca.gain, n->fp.dca.pan, &v[0], &v[1]); break;
    default:
      if (n->kind >= GB_FX_LOW_PASS_12DB && n->kind <= GB_FX_HIGH_SHELF_12DB) {
        BiquadCoefs c = host_rbj(n->kind, n->p[0], n->p[1], e->sr);
        v[0] = c.b0; v[1] = c.b1; v[2] = c.b2; v[3] = c.a1; v[4] = c.a2;
      }
      break;
  }
}

// Effects that can share a launch with other nodes of their level: 1 = memoryless (pointwise_batch_kernel),
// 2 = 24 dB low-pass, 3 = the biquad family; 0 = launched on their own.
int batch_class(const Node* n) {
  if (n->is_inst) return 0;
  switch (n->kind) {
    case GB_FX_GAIN: case GB_FX_LIMITER: case GB_FX_BITCRUSHER: case GB_FX_COMPRESSOR: return 1;
    case GB_FX_LOW_PASS_24DB: return 2;
    default: return n->kind >= GB_FX_LOW_PASS_12DB && n->kind <= GB_FX_HIGH_SHELF_12DB ? 3 : 0;
  }
}
int pointwise_op(const Node* n) {
  return n->kind == GB_FX_GAIN ? OP_GAIN : n->kind == GB_FX_LIMITER ? OP_LIMITER
       : n->kind == GB_FX_BITCRUSHER ? OP_BITCRUSHER : OP_COMPRESSOR;
}

// Run one effect node over the whole chunk; `segs`/`nseg` = its parameter segment table (device).
int run_effect(gb_engine* e, Node* n, const SourceList& src, int frames, const SegParam* segs, int nseg,
               int64_t chunk_pos, const PostOp* post = nullptr) {
  switch (n->kind) {
    case GB_FX_MIXER: case GB_FX_SIGNAL_PASSTHROUGH: {
      Launch l(e, false);
      pointwise_kernel<<<cdiv(frames, 256), 256, 0, e->stream>>>(src, n->buf, frames, OP_SUM, segs, nseg);
    } break;
    case GB_FX_GAIN: case GB_FX_LIMITER: case GB_FX_BITCRUSHER: case GB_FX_COMPRESSOR: {
      const int op = n->kind == GB_FX_GAIN ? OP_GAIN : n->kind == GB_FX_LIMITER ? OP_LIMITER
                   : n->kind == GB_FX_BITCRUSHER ? OP_BITCRUSHER : OP_COMPRESSOR;
      Launch l(e, false);
      pointwise_kernel<<<cdiv(frames, 256), 256, 0, e->stream>>>(src, n->buf, frames, op, segs, nseg);
    } break;
    case GB_FX_LOW_PASS_24DB: {
      Launch l(e, false);
      if (post) {
        FxDesc d;
        memset(&d, 0, sizeof d);
        d.src = src; d.out = n->post->buf; d.segs = segs; d.nseg = nseg; d.state = n->d_lp; d.post = *post;
        lp24_single_kernel<<<1, 32 * kFxWarps, 0, e->stream>>>(d, frames);
      } else {
        lp24_kernel<<<1, 32 * kFxWarps, 0, e->stream>>>(src, n->buf, frames, segs, nseg, n->d_lp);
      }
    } break;
    case GB_FX_CHORUS: {
      Launch l(e, false);
      chorus_kernel<<<cdiv(frames, 256), 256, 0, e->stream>>>(src, n->hist[n->hist_cur], std::max(n->delay_frames, 1),
                                                               n->taps, segs, nseg, n->buf, frames);
    } break;
    case GB_FX_REVERB: {
      int maxd = 0;
      for (int i = 0; i < 4; ++i) maxd = std::max(maxd, n->rv.comb_d[i]);
      {
        Launch l(e, false);
        dim3 grid(cdiv(maxd, 256), 8);
        reverb_comb_kernel<<<grid, 256, 0, e->stream>>>(src, n->rv, segs, nseg, n->rv_comb_out, frames, chunk_pos);
      }
      for (int stage = 0; stage < 2; ++stage) {
        Launch l(e, false);
        dim3 grid(cdiv(n->rv.ap_d[stage], 256), 2);
        reverb_allpass_kernel<<<grid, 256, 0, e->stream>>>(stage == 0 ? n->rv_comb_out : n->rv_ap_out[0],
                                                            stage == 0 ? 4 : 1, (size_t)frames, n->rv, stage,
                                                            n->rv_ap_out[stage], frames, chunk_pos);
      }
      {
        Launch l(e, false);
        interleave_kernel<<<cdiv(frames, 256), 256, 0, e->stream>>>(n->rv_ap_out[1], n->buf, frames);
      }
    } break;
    default:
      if (n->kind >= GB_FX_LOW_PASS_12DB && n->kind <= GB_FX_HIGH_SHELF_12DB) {
        Launch l(e, false);
        if (post) {
          FxDesc d;
          memset(&d, 0, sizeof d);
          d.src = src; d.out = n->post->buf; d.segs = segs; d.nseg = nseg; d.state = n->d_bq; d.post = *post;
          biquad_single_kernel<<<1, 32 * kFxWarps, 0, e->stream>>>(d, frames);
        } else {
          biquad_df1_kernel<<<1, 32 * kFxWarps, 0, e->stream>>>(src, n->buf, frames, segs, nseg, n->d_bq);
        }
      }
      break;
  }
  return 0;
}

}  // namespace

// ===================================================================== C ABI ===
extern "C" {

int gb_create(const gb_config* cfg, gb_engine** out) {
  if (!cfg || !out) return fail(nullptr, GB_EINVAL, "null argument");
  if (cfg->abi_version != GB_ABI_VERSION) return fail(nullptr, GB_EINVAL, "ABI version mismatch");
  if (!(cfg->sample_rate > 0.0)) return fail(nullptr, GB_EINVAL, "sample_rate must be positive");
  int count = 0;
  cudaError_t rc = cudaGetDeviceCount(&count);
  if (rc != cudaSuccess || count == 0)
    return fail(nullptr, GB_ENODEV, "no CUDA device available (%s); this library has no CPU fallback",
                rc == cudaSuccess ? "device count is 0" : cudaGetErrorString(rc));
  if (cfg->device < 0 || cfg->device >= count) return fail(nullptr, GB_ENODEV, "device ordinal out of range");
  if (cudaSetDevice(cfg->device) != cudaSuccess) return fail(nullptr, GB_ENODEV, "cudaSetDevice failed");
  std::unique_ptr<gb_engine> e(new gb_engine());
  e->device = cfg->device;
  e->sr = cfg->sample_rate;
  e->max_block = cfg->max_block ? cfg->max_block : kDefaultMaxBlock;
  if (e->max_block > (1u << 20)) e->max_block = 1u << 20;  // node buffers and the pinned ring are sized by it
  memset(&e->stats, 0, sizeof e->stats);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, e->device) == cudaSuccess) e->num_sms = prop.multiProcessorCount;
  if (const char* v = getenv("GB_CTA_MULT")) e->opt.cta_target_mult = std::max(1, atoi(v));
  if (const char* v = getenv("GB_VPC")) e->opt.vpc = std::max(1, atoi(v));
  if (const char* v = getenv("GB_KNOT_MAX_RATE")) e->opt.knot_max_rate = std::max(0.0, atof(v));
  if (const char* v = getenv("GB_LTI")) e->opt.lti = atoi(v) != 0;
  if (const char* v = getenv("GB_REST_KERNEL")) e->opt.rest_kernel = atoi(v) != 0;
  if (const char* v = getenv("GB_SWEEP_KERNEL")) e->opt.sweep_kernel = atoi(v) != 0;
  if (const char* v = getenv("GB_SYNC_KERNELS")) e->opt.sync_kernels = atoi(v) != 0;
  if (const char* v = getenv("GB_REST_TP")) e->opt.rest_tp = atoi(v) != 0;
  if (const char* v = getenv("GB_REST_NV")) e->opt.rest_nv = atoi(v) == 4 ? 4 : 2;
  if (const char* v = getenv("GB_OVERLAP")) e->overlap_enabled = atoi(v) != 0;
  if (const char* v = getenv("GB_REST_VR")) e->opt.rest_vr = atoi(v);
  if (const char* v = getenv("GB_REST16")) e->opt.rest16 = atoi(v) != 0;
  if (const char* v = getenv("GB_OSC_SYM")) e->opt.osc_sym = atoi(v) != 0;
  if (const char* v = getenv("GB_MIN_CUT_VOICES")) e->opt.min_cut_voices = std::max(1, atoi(v));
  if (const char* v = getenv("GB_SOLO_MIN")) e->min_solo_items = atoi(v);
  if (const char* v = getenv("GB_SOLO_WAVES")) e->solo_waves = atoi(v) != 0;
  if (const char* v = getenv("GB_FX_FUSE")) e->fuse_post_ops = atoi(v) != 0;
  if (const char* v = getenv("GB_FX_MINB")) e->fx_minb = atoi(v);
  if (const char* v = getenv("GB_FX_BATCH")) e->batch_fx = atoi(v) != 0;
  if (const char* v = getenv("GB_SOLO_LOCKSTEP")) e->solo_lockstep = atoi(v);
  if (const char* v = getenv("GB_SOLO_SUB")) e->solo_sub = std::max(1, atoi(v) / kBlockFrames) * kBlockFrames;
  if (const char* v = getenv("GB_FUSED_SUMS")) e->fused_sums_enabled = atoi(v) != 0;
  if (const char* v = getenv("GB_CHUNK_CUTS")) e->chunk_cuts = atoi(v) != 0;
  if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&e->vstream, cudaStreamNonBlocking) != cudaSuccess)
    return fail(nullptr, GB_ECUDA, "cudaStreamCreate failed");
  for (int i = 0; i < kStageSlots; ++i) {
    if (cudaEventCreateWithFlags(&e->stage_done[i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->d2d_done[i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->d2h_done[i], cudaEventDisableTiming) != cudaSuccess)
      return fail(nullptr, GB_ECUDA, "cudaEventCreate failed");
  }
  for (int i = 0; i < 2; ++i)
    if (cudaEventCreateWithFlags(&e->v_done[i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->p_done[i], cudaEventDisableTiming) != cudaSuccess)
      return fail(nullptr, GB_ECUDA, "cudaEventCreate failed");
  if (cudaEventCreateWithFlags(&e->call_start, cudaEventDisableTiming) != cudaSuccess)
    return fail(nullptr, GB_ECUDA, "cudaEventCreate failed");
  auto mixer = std::make_unique<Node>();
  mixer->uid = GB_MAIN_MIXER;
  mixer->kind = GB_FX_MIXER;
  index_node(e.get(), mixer.get());
  e->nodes[GB_MAIN_MIXER] = std::move(mixer);
  *out = e.release();
  return 0;
}

void gb_destroy(gb_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->vstream) cudaStreamSynchronize(e->vstream);  // a render that failed half-way may have left voice kernels enqueued
  if (e->stream) cudaStreamSynchronize(e->stream);
  if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);
  for (int i = 0; i < kStageSlots; ++i) {
    if (e->stage_done[i]) cudaEventDestroy(e->stage_done[i]);
    if (e->d2d_done[i]) cudaEventDestroy(e->d2d_done[i]);
    if (e->d2h_done[i]) cudaEventDestroy(e->d2h_done[i]);
  }
  if (e->ring) cudaFreeHost(e->ring);
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  if (e->vstream) cudaStreamDestroy(e->vstream);
  for (int i = 0; i < 2; ++i) {
    if (e->v_done[i]) cudaEventDestroy(e->v_done[i]);
    if (e->p_done[i]) cudaEventDestroy(e->p_done[i]);
  }
  if (e->call_start) cudaEventDestroy(e->call_start);
  for (void* p : e->allocations) cudaFree(p);
  if (e->d_full) cudaFree(e->d_full);
  if (e->d_pcm) cudaFree(e->d_pcm);
  resolve_spans(e);
  for (auto& ev : e->event_pool) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

const char* gb_last_error(const gb_engine* e) { return e ? e->err.c_str() : g_create_err.c_str(); }

int gb_add_instrument(gb_engine* e, int32_t kind, const void* params, size_t size, uint32_t* uid) {
  if (!e || !uid) return GB_EINVAL;
  if (e->finalized) return fail(e, GB_ESTATE, "engine is finalized");
  auto n = std::make_unique<Node>();
  n->uid = e->next_uid;
  n->kind = kind;
  n->is_inst = true;
  for (int& k : n->key_to_voice) k = -1;
  switch (kind) {
    case GB_INST_WELSH:
      if (!params || size != sizeof(gb_welsh_params)) return fail(e, GB_EINVAL, "bad welsh params size");
      n->wp = *(const gb_welsh_params*)params;
      n->nvoices = n->wp.voices ? (int)n->wp.voices : 8;
      n->release_frames = frames_of(n->wp.amp_envelope.release, e->sr);
      n->store.slots.resize(n->nvoices);
      break;
    case GB_INST_FM:
      if (!params || size != sizeof(gb_fm_params)) return fail(e, GB_EINVAL, "bad fm params size");
      n->fp = *(const gb_fm_params*)params;
      n->nvoices = n->fp.voices ? (int)n->fp.voices : 8;
      n->release_frames = frames_of(n->fp.carrier_envelope.release, e->sr);
      n->store.slots.resize(n->nvoices);
      break;
    case GB_INST_SAMPLER:
      if (!params || size != sizeof(gb_sampler_params)) return fail(e, GB_EINVAL, "bad sampler params size");
      n->sp = *(const gb_sampler_params*)params;
      n->nvoices = n->sp.voices ? (int)n->sp.voices : 8;
      n->store.slots.resize(n->nvoices);
      n->svoices.resize(n->nvoices);
      break;
    case GB_INST_DRUMKIT:
      break;
    case GB_INST_TOY_SOURCE:
      if (!params || size != sizeof(gb_toy_source_params)) return fail(e, GB_EINVAL, "bad toy params size");
      n->tp = *(const gb_toy_source_params*)params;
      break;
    case GB_INST_OSCILLATOR:
      if (!params || size != sizeof(gb_oscillator_source_params)) return fail(e, GB_EINVAL, "bad oscillator params size");
      n->osp = *(const gb_oscillator_source_params*)params;
      break;
    case GB_INST_ENVELOPE:
      if (!params || size != sizeof(gb_envelope_source_params)) return fail(e, GB_EINVAL, "bad envelope params size");
      n->esp = *(const gb_envelope_source_params*)params;
      break;
    default:
      return fail(e, GB_EINVAL, "unknown instrument kind %d", kind);
  }
  *uid = n->uid;
  index_node(e, n.get());
  e->nodes[n->uid] = std::move(n);
  e->next_uid++;
  return 0;
}

int gb_add_effect(gb_engine* e, int32_t kind, const void* params, size_t size, uint32_t* uid) {
  if (!e || !uid) return GB_EINVAL;
  if (e->finalized) return fail(e, GB_ESTATE, "engine is finalized");
  auto n = std::make_unique<Node>();
  n->uid = e->next_uid;
  n->kind = kind;
#define NEED(T) if (!params || size != sizeof(T)) return fail(e, GB_EINVAL, "bad effect params size for kind %d", kind)
  switch (kind) {
    case GB_FX_MIXER: case GB_FX_SIGNAL_PASSTHROUGH: break;
    case GB_FX_GAIN: NEED(gb_gain_params); n->p[0] = ((const gb_gain_params*)params)->ceiling; break;
    case GB_FX_LIMITER:
      NEED(gb_limiter_params);
      n->p[0] = ((const gb_limiter_params*)params)->min;
      n->p[1] = ((const gb_limiter_params*)params)->max;
      break;
    case GB_FX_BITCRUSHER: NEED(gb_bitcrusher_params); n->p[0] = ((const gb_bitcrusher_params*)params)->bits; break;
    case GB_FX_COMPRESSOR:
      NEED(gb_compressor_params);
      n->p[0] = ((const gb_compressor_params*)params)->threshold;
      n->p[1] = ((const gb_compressor_params*)params)->ratio;
      break;
    case GB_FX_DELAY:
      NEED(gb_delay_params);
      n->delay_frames = (int)frames_of(((const gb_delay_params*)params)->seconds, e->sr);
      break;
    case GB_FX_CHORUS: {
      NEED(gb_chorus_params);
      const gb_chorus_params* c = (const gb_chorus_params*)params;
      int nv = (int)c->voices;
      nv = nv < 1 ? 1 : (nv > 64 ? 64 : nv);
      int64_t d = frames_of(c->delay_seconds, e->sr);
      n->delay_frames = (int)d;
      n->taps.nv = nv;
      for (int i = 0; i < nv; ++i) n->taps.tap[i] = (int)(d * (i + 1) / nv);
      n->p[GB_CTL_CHORUS_WET_DRY_MIX] = c->wet_dry_mix;
    } break;
    case GB_FX_REVERB: {
      NEED(gb_reverb_params);
      const gb_reverb_params* r = (const gb_reverb_params*)params;
      static const double comb_s[4] = {0.0297, 0.0371, 0.0411, 0.0437};
      static const double ap_delay[2] = {0.0050, 0.0017};
      static const double ap_decay[2] = {0.09683, 0.03292};
      double seconds = r->seconds > 1e-6 ? r->seconds : 1e-6;
      memset(&n->rv, 0, sizeof n->rv);
      for (int i = 0; i < 4; ++i) {
        n->rv.comb_d[i] = (int)std::max<int64_t>(frames_of(comb_s[i], e->sr), 1);
        n->rv.comb_g[i] = std::pow(0.001, comb_s[i] / seconds);
      }
      for (int i = 0; i < 2; ++i) {
        n->rv.ap_d[i] = (int)std::max<int64_t>(frames_of(ap_delay[i], e->sr), 1);
        n->rv.ap_g[i] = std::pow(0.001, ap_delay[i] / ap_decay[i]);
      }
      n->p[0] = r->attenuation;
    } break;
    case GB_FX_LOW_PASS_24DB:
      NEED(gb_lowpass24_params);
      n->p[0] = ((const gb_lowpass24_params*)params)->cutoff;
      n->p[1] = ((const gb_lowpass24_params*)params)->passband_ripple;
      break;
    default:
      if (kind >= GB_FX_LOW_PASS_12DB && kind <= GB_FX_HIGH_SHELF_12DB) {
        NEED(gb_biquad_params);
        n->p[0] = ((const gb_biquad_params*)params)->cutoff;
        n->p[1] = ((const gb_biquad_params*)params)->param2;
      } else {
        return fail(e, GB_EINVAL, "unknown effect kind %d", kind);
      }
  }
#undef NEED
  *uid = n->uid;
  index_node(e, n.get());
  e->nodes[n->uid] = std::move(n);
  e->next_uid++;
  return 0;
}

int gb_load_sample(gb_engine* e, uint32_t uid, uint8_t key, const double* frames, size_t n_frames, int32_t channels,
                   double sample_rate, double root_hz) {
  if (!e) return GB_EINVAL;
  if (!frames || n_frames == 0 || (channels != 1 && channels != 2)) return fail(e, GB_EINVAL, "bad sample");
  Node* n = find(e, uid);
  if (!n) return fail(e, GB_ENOENT, "unknown uid %u", uid);
  if (n->kind != GB_INST_SAMPLER && n->kind != GB_INST_DRUMKIT) return fail(e, GB_EINVAL, "entity does not take samples");
  // the sample table is part of the plan (voice count, state blob layout): frozen by gb_finalize
  if (e->finalized) return fail(e, GB_ESTATE, "engine is finalized");
  if (n->kind == GB_INST_DRUMKIT && key >= 128) return fail(e, GB_EINVAL, "bad key");
  if (!(sample_rate > 0.0)) return fail(e, GB_EINVAL, "bad sample rate");
  cudaSetDevice(e->device);
  SampleDev s;
  s.n = n_frames;
  s.channels = channels;
  s.sr = sample_rate;
  size_t bytes = n_frames * (size_t)channels * sizeof(double);
  CUDA_TRY(e, cudaMalloc(&s.d, bytes));
  if (cudaMemcpy(s.d, frames, bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(s.d);
    return fail(e, GB_ECUDA, "sample upload failed");
  }
  e->allocations.push_back(s.d);
  e->stats.h2d_bytes += bytes;
  auto release = [&](SampleDev& old) {  // a replaced sample's buffer is freed now, not at destroy
    auto it = std::find(e->allocations.begin(), e->allocations.end(), (void*)old.d);
    if (it != e->allocations.end()) e->allocations.erase(it);
    cudaFree(old.d);
    old.d = nullptr;
  };
  if (n->kind == GB_INST_SAMPLER) {
    s.root_hz = root_hz > 0.0 ? root_hz : (n->sp.root_hz > 0.0 ? n->sp.root_hz : 440.0);
    for (SampleDev& old : n->samples) release(old);
    n->samples.clear();
    n->samples.push_back(s);
  } else {
    if (n->key_to_voice[key] < 0) {
      n->key_to_voice[key] = (int)n->svoices.size();
      n->samples.push_back(s);
      SampleVoiceHost v;
      v.sample = (int)n->samples.size() - 1;
      n->svoices.push_back(v);
    } else {
      SampleDev& slot = n->samples[(size_t)n->svoices[(size_t)n->key_to_voice[key]].sample];
      release(slot);
      slot = s;
    }
  }
  return 0;
}

int gb_patch(gb_engine* e, uint32_t src, uint32_t dst) {
  if (!e) return GB_EINVAL;
  if (e->finalized) return fail(e, GB_ESTATE, "engine is finalized");
  Node* s = find(e, src);
  Node* d = find(e, dst);
  if (!s || !d) return fail(e, GB_ENOENT, "unknown uid");
  if (d->is_inst) return fail(e, GB_EGRAPH, "input device is not an effect");
  if (src == dst) return fail(e, GB_EGRAPH, "cannot patch a device to itself");
  if (std::find(d->sources.begin(), d->sources.end(), src) == d->sources.end()) d->sources.push_back(src);
  return 0;
}

int gb_link_control(gb_engine* e, uint32_t src, uint32_t dst, int32_t index) {
  if (!e) return GB_EINVAL;
  if (e->finalized) return fail(e, GB_ESTATE, "engine is finalized");
  Node* s = find(e, src);
  Node* d = find(e, dst);
  if (!s || !d) return fail(e, GB_ENOENT, "unknown uid");
  if (s->kind != GB_FX_SIGNAL_PASSTHROUGH) return fail(e, GB_EINVAL, "link source is not a signal-passthrough node");
  if (!(d->kind == GB_FX_GAIN || d->kind == GB_FX_LIMITER || d->kind == GB_FX_COMPRESSOR) || !effect_accepts(d, index))
    return fail(e, GB_EINVAL, "link target must be a gain, limiter or compressor parameter");
  for (auto& l : e->links)
    if (l.dst == dst) return fail(e, GB_EINVAL, "target already has a control link");
  gb_engine::Link l;
  l.src = src; l.dst = dst; l.index = index;
  e->links.push_back(l);
  return 0;
}

// The instrument tables (coefficient sets, scan maps, rotation tables: welsh_inst_from_params) are part of the
// plan: built and uploaded by gb_finalize, and again before the next chunk whenever a parameter changed.
int upload_inst_tables(gb_engine* e) {
  if (e->winst_dirty && !e->h_winst.empty()) {
    for (Node* n : e->plan)
      if (n->kind == GB_INST_WELSH) welsh_inst_from_params(*n, e->sr, e->opt, &e->h_winst[(size_t)n->table_index]);
    CUDA_TRY(e, cudaMemcpyAsync(e->d_winst, e->h_winst.data(), e->h_winst.size() * sizeof(WelshInst),
                                cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));  // h_winst is pageable
    e->stats.h2d_bytes += e->h_winst.size() * sizeof(WelshInst);
  }
  e->winst_dirty = false;
  if (e->finst_dirty && !e->h_finst.empty()) {
    for (Node* n : e->plan)
      if (n->kind == GB_INST_FM) fm_inst_from_params(*n, e->sr, &e->h_finst[(size_t)n->table_index]);
    CUDA_TRY(e, cudaMemcpyAsync(e->d_finst, e->h_finst.data(), e->h_finst.size() * sizeof(FmInst),
                                cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    e->stats.h2d_bytes += e->h_finst.size() * sizeof(FmInst);
  }
  e->finst_dirty = false;
  return 0;
}

// Voice ranges for all-resting chunks (welsh_rest_vr_kernel).  Called by gb_finalize once the instrument tables
// exist (the rest classes are part of them).  Needs the overlapped-mixdown layout plus: one consumer for all
// instruments, instruments with even voice counts of at least one range, laid out back to back in the voice
// table, one common non-sync rest class.
int plan_voice_ranges(gb_engine* e) {
  e->vr_ok = false;
  if (!e->overlap) return 0;
  const size_t mb = e->max_block;
  int rc2;
  constexpr int kVrVoices = 14;
  Node* consumer = nullptr;
  int n_cons = 0, n_inst = 0, n_fused_src = 0, next_voice = 0, cls = -2, sweep_cls = -1;
  bool vr = e->opt.rest_vr != 0 && e->opt.rest_kernel;
  for (Node* n : e->plan) {
    if (!n->is_inst) {
      if (n->d_src_table_fused) { consumer = n; ++n_cons; }
      continue;
    }
    ++n_inst;
    const WelshInst& I = e->h_winst[(size_t)n->table_index];
    vr = vr && n->nvoices % 2 == 0 && n->nvoices >= kVrVoices && n->voice0 == next_voice && I.rest_class >= 0 &&
         I.rest_class < 4 && (cls == -2 || cls == I.rest_class);
    cls = I.rest_class;
    if (n_inst == 1) sweep_cls = I.sweep_class;
    else if (sweep_cls != I.sweep_class) sweep_cls = -1;
    next_voice = n->voice0 + n->nvoices;
  }
  if (vr && n_cons == 1) {
    for (uint32_t su : consumer->sources) {
      Node* sn = find(e, su);
      if (sn && sn->order >= 0 && sn->is_inst) ++n_fused_src;
    }
  }
  vr = vr && n_cons == 1 && n_fused_src == n_inst && next_voice == e->n_wvoice;
  const int slots = 2 * e->num_sms;
  const int ranges = (e->n_wvoice + kVrVoices - 1) / kVrVoices;
  // automatic: only where the ranges lower the voices on the fullest SM (4096 voices: 293 ranges of 14 put 28
  // voices on every SM, the 256 instrument CTAs of 16 put 32 on 108 SMs and 16 on the other 40)
  if (e->opt.rest_vr < 0) {
    int vpc_max = 0;
    for (int i = 0; i < e->n_wwork_grouped; ++i) vpc_max = std::max(vpc_max, e->wwork.h[i].nvoices);
    vr = vr && ranges <= slots &&
         cdiv(ranges, e->num_sms) * kVrVoices < cdiv(e->n_wwork_grouped, e->num_sms) * vpc_max && vpc_max > kTpMaxVoices;
  }
  if (vr) {
    std::vector<const Node*> inst_of((size_t)e->n_wvoice, nullptr);
    for (Node* n : e->plan)
      if (n->is_inst)
        for (int v = 0; v < n->nvoices; ++v) inst_of[(size_t)(n->voice0 + v)] = n;
    for (int par2 = 0; par2 < 2; ++par2) {
      double2* outs = nullptr;
      if ((rc2 = dev_alloc(e, &outs, (size_t)ranges * mb))) return rc2;
      std::vector<VrWork> vw((size_t)ranges);
      for (int r = 0; r < ranges; ++r) {
        VrWork& w = vw[(size_t)r];
        w.voice0 = r * kVrVoices;
        w.nvoices = std::min(kVrVoices, e->n_wvoice - w.voice0);
        const Node* a = inst_of[(size_t)w.voice0];
        const Node* b = inst_of[(size_t)(w.voice0 + w.nvoices - 1)];
        w.inst_a = a->table_index;
        w.inst_b = b->table_index;
        w.split = a == b ? w.nvoices : a->voice0 + a->nvoices - w.voice0;
        // the single-voice warps of the second wave of CTAs sit on the other two sub-partitions
        w.single0 = r < e->num_sms ? 6 : 4;
        w.out = outs + (size_t)r * mb;
      }
      if ((rc2 = dev_alloc(e, &e->d_vr_work[par2], vw.size(), false))) return rc2;
      CUDA_TRY(e, cudaMemcpy(e->d_vr_work[par2], vw.data(), vw.size() * sizeof(VrWork), cudaMemcpyHostToDevice));
      std::vector<const double2*> tab;
      for (uint32_t su : consumer->sources) {
        Node* sn = find(e, su);
        if (sn && sn->order >= 0 && sn->buf && !sn->is_inst) tab.push_back(sn->buf);
      }
      for (int r = 0; r < ranges; ++r) tab.push_back(vw[(size_t)r].out);
      if ((rc2 = dev_alloc(e, &consumer->d_src_table_vr[par2], tab.size(), false))) return rc2;
      CUDA_TRY(e, cudaMemcpy((void*)consumer->d_src_table_vr[par2], tab.data(), tab.size() * sizeof(double2*), cudaMemcpyHostToDevice));
      consumer->n_src_vr = (int)tab.size();
    }
    int osc = 0;
    {  // tables for 16 frames per lane: the same constructions as LtiTable / lane_rot with kT16
      std::vector<Rest16Table> tabs(e->h_winst.size());
      memset(tabs.data(), 0, tabs.size() * sizeof(Rest16Table));
      osc = (cls & 1) && e->opt.osc_sym ? 2 : 0;  // piecewise-constant oscillators only
      for (Node* n : e->plan) {
        if (!n->is_inst) continue;
        const WelshInst& I = e->h_winst[(size_t)n->table_index];
        const bool sym = I.m1bb.b_hi == -I.m1bb.b_lo && I.m2bb.b_hi == -I.m2bb.b_lo && I.m1bb.a_lo == 0.0 && I.m1bb.a_hi == 0.0 &&
                         I.m2bb.a_lo == 0.0 && I.m2bb.a_hi == 0.0;
        if (!sym) osc = 0;
        else if (osc == 2 && I.s2.thresh != (1ull << 63)) osc = 1;
        Rest16Table& R = tabs[(size_t)n->table_index];
        auto table = [](const SecCoef& c, double (*g)[2], double (*mp)[4]) {
          double h00 = 1.0, h01 = 0.0, h10 = 0.0, h11 = 1.0;  // H = A^j, A = [[a1, 1], [a2, 0]]
          for (int j = 0; j < kT16; ++j) {
            g[j][0] = h00; g[j][1] = h01;
            const double t00 = c.a1 * h00 + h10, t01 = c.a1 * h01 + h11;
            h10 = c.a2 * h00; h11 = c.a2 * h01;
            h00 = t00; h01 = t01;
          }
          mp[0][0] = h00; mp[0][1] = h01; mp[0][2] = h10; mp[0][3] = h11;
          for (int k = 1; k < 5; ++k) {
            const double* m = mp[k - 1];
            mp[k][0] = m[0] * m[0] + m[1] * m[2]; mp[k][1] = m[0] * m[1] + m[1] * m[3];
            mp[k][2] = m[2] * m[0] + m[3] * m[2]; mp[k][3] = m[2] * m[1] + m[3] * m[3];
          }
          mp[5][0] = mp[5][1] = mp[5][2] = mp[5][3] = 0.0;
        };
        table(I.lti.c1, R.g1b, R.mp1);
        table(I.lti.c2, R.g2, R.mp2);
        for (int j = 0; j < kT16; ++j) { R.g1b[j][0] *= I.lti.c2.b0; R.g1b[j][1] *= I.lti.c2.b0; }
        auto rot = [&](uint64_t steps) {
          const uint64_t q = steps * I.lfo_dq;  // mod 2^64
          const double ang = 6.283185307179586476925286766559 * ((double)q / 18446744073709551616.0);
          return make_double2(std::cos(ang), std::sin(ang));
        };
        for (int j = 0; j < kT16; ++j) R.lfo_rot[j] = rot((uint64_t)j);
        for (int l = 0; l < 32; ++l) R.lane_rot[l] = rot((uint64_t)(l * kT16));
        R.block_rot = rot((uint64_t)kBlock16);
      }
      if ((rc2 = dev_alloc(e, &e->d_rest16, tabs.size(), false))) return rc2;
      CUDA_TRY(e, cudaMemcpy(e->d_rest16, tabs.data(), tabs.size() * sizeof(Rest16Table), cudaMemcpyHostToDevice));
      CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_vr16_kernel<8, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
      CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_vr16_kernel<8, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
      CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_vr16_kernel<8, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
      CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_vr16_kernel<8, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
      CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_vr16_kernel<8, false, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
      CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_vr16_kernel<8, false, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
      CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_vr16_kernel<8, true, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
      CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_vr16_kernel<8, true, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
    }
    e->vr_ok = true;
    e->vr_sweep_class = sweep_cls >= 0 && sweep_cls < 4 && e->opt.sweep_kernel ? sweep_cls : -1;
    e->vr_osc = osc;
    e->vr_class = cls;
    e->vr_count = ranges;
  }
  return 0;
}


int gb_finalize(gb_engine* e) {
  if (!e) return GB_EINVAL;
  if (e->finalized) return fail(e, GB_ESTATE, "engine is already finalized");
  cudaSetDevice(e->device);
  // a finalize that failed half-way may be retried: start from a clean plan
  e->plan.clear();
  e->h_winst.clear();
  e->h_finst.clear();
  e->wwork_node.clear();
  for (auto& kv : e->nodes) {
    kv.second->consumers = 0;
    kv.second->order = -1;
    kv.second->fuse_partials = false;
    kv.second->partial_count = 0;
  }
  // Scheduling edges: a link target must run after its source.  Only sources that the main mixer reaches
  // through patch cables count (unpatched entities never render, orchestrator.rs:378-430).
  std::map<uint32_t, std::vector<uint32_t>> children;
  {
    std::map<uint32_t, bool> reach;
    std::vector<uint32_t> todo{GB_MAIN_MIXER};
    reach[GB_MAIN_MIXER] = true;
    while (!todo.empty()) {
      Node* n = find(e, todo.back());
      todo.pop_back();
      if (!n) continue;
      for (uint32_t c : n->sources)
        if (!reach[c]) { reach[c] = true; todo.push_back(c); }
    }
    for (auto& kv : e->nodes) {
      std::vector<uint32_t>& ch = children[kv.first];
      for (auto& l : e->links)
        if (l.dst == kv.first && reach[l.src] && reach[l.dst]) ch.push_back(l.src);
      for (uint32_t c : kv.second->sources) ch.push_back(c);
    }
  }
  // post-order DFS from the main mixer: sources before consumers; detects cycles.
  std::map<uint32_t, int> color;
  std::vector<std::pair<Node*, size_t>> st;
  Node* root = find(e, GB_MAIN_MIXER);
  st.push_back({root, 0});
  color[root->uid] = 1;
  e->plan.clear();
  while (!st.empty()) {
    auto& top = st.back();
    const std::vector<uint32_t>& kids = children[top.first->uid];
    if (top.second >= kids.size()) {
      color[top.first->uid] = 2;
      top.first->order = (int)e->plan.size();
      e->plan.push_back(top.first);
      st.pop_back();
      continue;
    }
    uint32_t cu = kids[top.second++];
    Node* c = find(e, cu);
    if (!c) continue;
    if (color[cu] == 1) return fail(e, GB_EGRAPH, "patch graph (with control links) has a cycle");
    if (color[cu] == 0) {
      color[cu] = 1;
      st.push_back({c, 0});
    }
  }
  // Levels: a node's sources (and control-link sources) sit on lower levels, so level order is a valid
  // execution order in which independent nodes of one kind are neighbours — render_chunk turns such runs
  // into ONE batched launch (parallel chains of a song, a batch of songs).
  for (Node* n : e->plan) {
    n->level = 0;
    n->post = nullptr;
    n->fused_away = false;
    for (uint32_t c : children[n->uid]) {
      Node* cn = find(e, c);
      if (cn && cn->order >= 0) n->level = std::max(n->level, cn->level + 1);
    }
  }
  std::stable_sort(e->plan.begin(), e->plan.end(), [](const Node* a, const Node* b) {
    return a->level != b->level ? a->level < b->level : (a->level > 0 && batch_class(a) < batch_class(b));
  });
  for (size_t i = 0; i < e->plan.size(); ++i) e->plan[i]->order = (int)i;
  const size_t mb = e->max_block;
  int wv = 0, fv = 0;
  for (Node* n : e->plan) {
    int rc = dev_alloc(e, &n->buf, mb);
    if (rc) return rc;
    if (n->kind == GB_INST_WELSH) {
      n->table_index = (int)e->h_winst.size();
      n->voice0 = wv;
      wv += n->nvoices;
      e->h_winst.emplace_back();
    } else if (n->kind == GB_INST_FM) {
      n->table_index = (int)e->h_finst.size();
      n->voice0 = fv;
      fv += n->nvoices;
      e->h_finst.emplace_back();
    }
    if (!n->is_inst) {
      if ((int)n->sources.size() > kMaxSources && (rc = dev_alloc(e, &n->scratch, mb))) return rc;
      if (n->kind >= GB_FX_LOW_PASS_12DB && n->kind <= GB_FX_HIGH_SHELF_12DB && (rc = dev_alloc(e, &n->d_bq, 1))) return rc;
      if (n->kind == GB_FX_LOW_PASS_24DB && (rc = dev_alloc(e, &n->d_lp, 1))) return rc;
      if (n->kind == GB_FX_DELAY || n->kind == GB_FX_CHORUS) {
        size_t len = (size_t)std::max(n->delay_frames, 1);
        if ((rc = dev_alloc(e, &n->hist[0], len)) || (rc = dev_alloc(e, &n->hist[1], len))) return rc;
      }
      if (n->kind == GB_FX_REVERB) {
        for (int ch = 0; ch < 2; ++ch) {
          for (int i = 0; i < 4; ++i)
            if ((rc = dev_alloc(e, &n->rv.comb_ring[ch][i], (size_t)n->rv.comb_d[i]))) return rc;
          for (int i = 0; i < 2; ++i)
            if ((rc = dev_alloc(e, &n->rv.ap_ring[ch][i], (size_t)n->rv.ap_d[i]))) return rc;
        }
        if ((rc = dev_alloc(e, &n->rv_comb_out, 8 * mb)) || (rc = dev_alloc(e, &n->rv_ap_out[0], 2 * mb)) ||
            (rc = dev_alloc(e, &n->rv_ap_out[1], 2 * mb)))
          return rc;
      }
    }
  }
  for (auto& l : e->links) {
    Node* sn = find(e, l.src);
    Node* dn = find(e, l.dst);
    if (!sn || !dn || sn->order < 0 || dn->order < 0) continue;
    int rc = dev_alloc(e, &l.d_table, mb / GB_CONTROL_PERIOD + 3);
    if (rc) return rc;
    if ((rc = dev_alloc(e, &l.d_state, 8))) return rc;
    const double init[8] = {dn->p[l.index], 0.0, 0.0, 0.0, dn->p[l.index], 0.0, 0.0, 0.0};
    CUDA_TRY(e, cudaMemcpy(l.d_state, init, sizeof init, cudaMemcpyHostToDevice));
  }
  // pinned ring for host-buffer renders (render_impl): allocated here so that no render call pays for it
  if (!e->ring) CUDA_TRY(e, cudaMallocHost(&e->ring, (size_t)kStageSlots * e->max_block * sizeof(double2)));
  // voice state: every voice starts idle
  e->n_wvoice = wv;
  e->n_fvoice = fv;
  if (wv) {
    int rc = dev_alloc(e, &e->d_wvoice, (size_t)wv, false);
    if (rc) return rc;
    if ((rc = dev_alloc(e, &e->d_winst, e->h_winst.size()))) return rc;
    std::vector<WelshVoice> init((size_t)wv);
    memset(init.data(), 0, init.size() * sizeof(WelshVoice));
    for (auto& v : init) { v.n_on = kNever; v.n_off = kNever; v.anchor = 0; v.knot_frame = kNever; }
    CUDA_TRY(e, cudaMemcpy(e->d_wvoice, init.data(), init.size() * sizeof(WelshVoice), cudaMemcpyHostToDevice));
  }
  if (fv) {
    int rc = dev_alloc(e, &e->d_fvoice, (size_t)fv, false);
    if (rc) return rc;
    if ((rc = dev_alloc(e, &e->d_finst, e->h_finst.size()))) return rc;
    std::vector<FmVoice> init((size_t)fv);
    memset(init.data(), 0, init.size() * sizeof(FmVoice));
    for (auto& v : init) { v.n_on = kNever; v.n_off = kNever; v.anchor = 0; }
    CUDA_TRY(e, cudaMemcpy(e->d_fvoice, init.data(), init.size() * sizeof(FmVoice), cudaMemcpyHostToDevice));
  }
  // CTA work lists.  vpc: aim for about two CTAs per SM across all voices of a type.
  auto plan_work = [&](int kind, int total_voices, DevBuf<CtaWork>& buf, DevBuf<WarpItem>& ibuf, int* count,
                       int* n_grouped) -> int {
    std::vector<CtaWork> work;
    std::vector<WarpItem> items;
    if (total_voices == 0) { *count = 0; *n_grouped = 0; return 0; }
    int target = std::max(1, e->num_sms * e->opt.cta_target_mult);
    const int wpc = kVoiceWarps;  // warps per CTA (4-warp CTAs with tighter register caps measured slower)
    int vpc = std::max(wpc, cdiv(total_voices, target));
    vpc = cdiv(vpc, wpc) * wpc;
    // Two voices per warp (lockstep pairs) beat twice the CTAs with one voice per warp as long as the pairs
    // still give about one CTA per SM: 2048 voices render in 28.6 ms as 128 CTAs of 16 against 30.5 ms as 256
    // CTAs of 8; 1024 voices in 20.0 ms as 128 CTAs of 8 against 27.5 ms as 64 of 16 (profiles/r2_strong_probe.txt).
    if (vpc == wpc && 20 * total_voices >= 17 * 2 * wpc * e->num_sms) vpc = 2 * wpc;
    // Few voices for the machine (a strong-scaling shard of 1/4 or 1/8): CTAs of two voice pairs whose resting
    // stretches go to welsh_rest_tp_kernel (its warps share a pair along time).  Measured per shard
    // (profiles/r2_strong_probe.txt): 1024 voices 17.9 ms as 256 such CTAs against 20.0 ms as 128 CTAs of 8 on the
    // per-warp kernel; 512 voices 11.5 ms as 128 CTAs of 4 (12.4 ms as 256 CTAs of 2: twice the partials to mix)
    // against 19.1 ms; 2048 voices stay with 16-voice CTAs (28.6 ms against 29.9 ms).
    // Small songs (< 256 voices) keep whole-instrument CTAs: there launches, not SM coverage, are the cost.
    if (kind == GB_INST_WELSH && e->opt.rest_tp && e->opt.rest_kernel && total_voices >= 256 &&
        cdiv(total_voices, target) <= 4)
      vpc = 4;
    if (e->opt.vpc > 0) vpc = e->opt.vpc;  // tests: force the voices-per-CTA split
    for (Node* n : e->plan) {
      if (n->kind != kind) continue;
      if (n->nvoices < wpc) {
        // solo warps: every voice is its own work item with its own output (partial) buffer
        n->partial_count = n->nvoices > 1 ? n->nvoices : 0;
        if (n->partial_count) {
          int rc = dev_alloc(e, &n->scratch, (size_t)n->partial_count * mb);
          if (rc) return rc;
        }
        for (int v = 0; v < n->nvoices; ++v) {
          WarpItem it;
          it.inst = n->table_index;
          it.voice = n->voice0 + v;
          it.out = n->partial_count ? n->scratch + (size_t)v * mb : n->buf;
          items.push_back(it);
          if (kind == GB_INST_WELSH) {
            e->witem_node.push_back(n);
            e->witem_slot.push_back(v);
          }
        }
        continue;
      }
      int ncta = cdiv(n->nvoices, vpc);
      n->partial_count = ncta > 1 ? ncta : 0;
      if (ncta > 1) {
        int rc = dev_alloc(e, &n->scratch, (size_t)ncta * mb);
        if (rc) return rc;
      }
      int per = cdiv(n->nvoices, ncta);
      int v = 0;
      for (int c = 0; c < ncta; ++c) {
        CtaWork w;
        w.inst = n->table_index;
        w.voice0 = n->voice0 + v;
        w.nvoices = std::min(per, n->nvoices - v);
        w.solo = 0;
        w.out = ncta == 1 ? n->buf : n->scratch + (size_t)c * mb;
        work.push_back(w);
        if (kind == GB_INST_WELSH) e->wwork_node.push_back(n);
        v += w.nvoices;
      }
    }
    *n_grouped = (int)work.size();
    for (size_t i = 0; i < items.size(); i += (size_t)wpc) {
      CtaWork w;
      w.inst = items[i].inst;
      w.voice0 = (int)i;
      w.nvoices = (int)std::min<size_t>((size_t)wpc, items.size() - i);
      w.solo = 1;
      w.out = nullptr;
      work.push_back(w);
    }
    if (!buf.reserve(work.size()) || !ibuf.reserve(items.size() + 1)) return fail(e, GB_ENOMEM, "out of memory");
    memcpy(buf.h, work.data(), work.size() * sizeof(CtaWork));
    CUDA_TRY(e, cudaMemcpy(buf.d, buf.h, work.size() * sizeof(CtaWork), cudaMemcpyHostToDevice));
    if (!items.empty()) {
      memcpy(ibuf.h, items.data(), items.size() * sizeof(WarpItem));
      CUDA_TRY(e, cudaMemcpy(ibuf.d, ibuf.h, items.size() * sizeof(WarpItem), cudaMemcpyHostToDevice));
    }
    *count = (int)work.size();
    return 0;
  };
  int rc;
  int fm_grouped = 0;
  e->wwork_node.clear();
  e->witem_node.clear();
  e->witem_slot.clear();
  if ((rc = plan_work(GB_INST_WELSH, wv, e->wwork, e->witems, &e->n_wwork, &e->n_wwork_grouped))) return rc;
  e->solo_mega = !e->witem_node.empty() && (int)e->witem_node.size() >= e->min_solo_items;
  e->solo_dirty.assign(e->witem_node.size(), gb_engine::SoloDirty());
  if (e->solo_mega && ((rc = dev_alloc(e, &e->d_solo_sync, (size_t)wv + kSoloTickets)) || (rc = dev_alloc(e, &e->d_solo_fault, 8)))) return rc;
  if ((rc = plan_work(GB_INST_FM, fv, e->fwork, e->fitems, &e->n_fwork, &fm_grouped))) return rc;
  {
    std::vector<PartialDesc> descs;
    for (Node* n : e->plan) {
      if (!(n->kind == GB_INST_WELSH || n->kind == GB_INST_FM) || n->partial_count == 0) continue;
      PartialDesc d;
      d.base = n->scratch; d.out = n->buf; d.stride = mb; d.count = n->partial_count; d.pad = 0;
      descs.push_back(d);
    }
    e->n_partials = (int)descs.size();
    if (!descs.empty()) {
      if (!e->partials.reserve(descs.size())) return fail(e, GB_ENOMEM, "out of memory");
      memcpy(e->partials.h, descs.data(), descs.size() * sizeof(PartialDesc));
      CUDA_TRY(e, cudaMemcpy(e->partials.d, e->partials.h, descs.size() * sizeof(PartialDesc), cudaMemcpyHostToDevice));
    }
    for (Node* n : e->plan)
      if (!n->is_inst)
        for (uint32_t su : n->sources) {
          Node* sn = find(e, su);
          if (sn && sn->order >= 0) sn->consumers++;
        }
    // A memoryless effect whose only source is an IIR stage that nothing else reads runs as that stage's
    // post-op: one read and one write per frame for the pair (GB_FX_FUSE=0 keeps them apart).
    if (e->fuse_post_ops)
      for (Node* n : e->plan) {
        if (batch_class(n) != 1 || n->sources.size() != 1) continue;
        Node* a = find(e, n->sources[0]);
        if (!a || a->order < 0 || batch_class(a) < 2 || a->consumers != 1 || a->post) continue;
        bool linked = false;
        for (auto& l : e->links) linked = linked || l.dst == n->uid || l.src == a->uid || l.dst == a->uid;
        if (linked) continue;
        a->post = n;
        n->fused_away = true;
      }
    for (Node* n : e->plan) {
      if (n->is_inst || (int)n->sources.size() <= kMaxSources) continue;
      std::vector<const double2*> ptrs, fused;
      for (uint32_t su : n->sources) {
        Node* sn = find(e, su);
        if (!(sn && sn->order >= 0 && sn->buf)) continue;
        ptrs.push_back(sn->buf);
        // a split instrument read by this node only: sum its partial buffers here instead of reducing
        // them into its node buffer first (one pass over HBM less)
        if ((sn->kind == GB_INST_WELSH || sn->kind == GB_INST_FM) && sn->partial_count > 0 && sn->consumers == 1) {
          sn->fuse_partials = true;
          for (int k = 0; k < sn->partial_count; ++k) fused.push_back(sn->scratch + (size_t)k * mb);
        } else {
          fused.push_back(sn->buf);
        }
      }
      int rc2 = dev_alloc(e, &n->d_src_table, ptrs.size());
      if (rc2) return rc2;
      CUDA_TRY(e, cudaMemcpy((void*)n->d_src_table, ptrs.data(), ptrs.size() * sizeof(double2*), cudaMemcpyHostToDevice));
      if (fused.size() != ptrs.size()) {
        if ((rc2 = dev_alloc(e, &n->d_src_table_fused, fused.size()))) return rc2;
        CUDA_TRY(e, cudaMemcpy((void*)n->d_src_table_fused, fused.data(), fused.size() * sizeof(double2*), cudaMemcpyHostToDevice));
        n->n_src_fused = (int)fused.size();
      }
    }
    {
      std::vector<PartialDesc> nf;
      for (Node* n : e->plan) {
        if (!(n->kind == GB_INST_WELSH || n->kind == GB_INST_FM) || n->partial_count == 0 || n->fuse_partials) continue;
        PartialDesc d;
        d.base = n->scratch; d.out = n->buf; d.stride = mb; d.count = n->partial_count; d.pad = 0;
        nf.push_back(d);
      }
      e->n_partials_nf = (int)nf.size();
      if (!nf.empty()) {
        if (!e->partials_nf.reserve(nf.size())) return fail(e, GB_ENOMEM, "out of memory");
        memcpy(e->partials_nf.h, nf.data(), nf.size() * sizeof(PartialDesc));
        CUDA_TRY(e, cudaMemcpy(e->partials_nf.d, e->partials_nf.h, nf.size() * sizeof(PartialDesc), cudaMemcpyHostToDevice));
      }
    }
  }
  // Overlapped mixdown: possible when every voice kernel output is a partial buffer that only a consumer's table
  // sum (or reduce_partials) reads — all instruments split Welsh instruments, no solo items, no control links.
  // Those buffers, the CTA list that points at them, and the consumers' tables then exist twice (chunk parity).
  e->overlap = false;
  if (e->overlap_enabled && e->n_wwork_grouped > 0 && e->n_wwork == e->n_wwork_grouped && e->links.empty()) {
    bool ok = true;
    for (Node* n : e->plan)
      if (n->is_inst) ok = ok && n->kind == GB_INST_WELSH && n->partial_count > 0 && n->fuse_partials;
    if (ok) {
      int rc2;
      for (Node* n : e->plan)
        if (n->is_inst && (rc2 = dev_alloc(e, &n->scratch_alt, (size_t)n->partial_count * mb))) return rc2;
      std::vector<CtaWork> alt(e->wwork.h, e->wwork.h + e->n_wwork);
      for (int i = 0; i < e->n_wwork; ++i) {
        const Node* n = e->wwork_node[(size_t)i];
        alt[(size_t)i].out = n->scratch_alt + (alt[(size_t)i].out - n->scratch);
      }
      if ((rc2 = dev_alloc(e, &e->d_wwork_alt, alt.size(), false))) return rc2;
      CUDA_TRY(e, cudaMemcpy(e->d_wwork_alt, alt.data(), alt.size() * sizeof(CtaWork), cudaMemcpyHostToDevice));
      std::vector<PartialDesc> pd(e->partials.h, e->partials.h + e->n_partials);
      {
        size_t k = 0;
        for (Node* n : e->plan) {
          if (!(n->kind == GB_INST_WELSH || n->kind == GB_INST_FM) || n->partial_count == 0) continue;
          pd[k++].base = n->scratch_alt;
        }
      }
      if ((rc2 = dev_alloc(e, &e->d_partials_alt, pd.size(), false))) return rc2;
      CUDA_TRY(e, cudaMemcpy(e->d_partials_alt, pd.data(), pd.size() * sizeof(PartialDesc), cudaMemcpyHostToDevice));
      for (Node* n : e->plan) {
        if (n->is_inst || !n->d_src_table_fused) continue;
        std::vector<const double2*> fused;
        for (uint32_t su : n->sources) {
          Node* sn = find(e, su);
          if (!(sn && sn->order >= 0 && sn->buf)) continue;
          if (sn->fuse_partials)
            for (int k = 0; k < sn->partial_count; ++k) fused.push_back(sn->scratch_alt + (size_t)k * mb);
          else
            fused.push_back(sn->buf);
        }
        n->fused_all_inst = true;
        for (uint32_t su : n->sources) {
          Node* sn = find(e, su);
          if (sn && sn->order >= 0 && sn->buf && !sn->fuse_partials) n->fused_all_inst = false;
        }
        if ((rc2 = dev_alloc(e, &n->d_src_table_fused_alt, fused.size(), false))) return rc2;
        CUDA_TRY(e, cudaMemcpy((void*)n->d_src_table_fused_alt, fused.data(), fused.size() * sizeof(double2*), cudaMemcpyHostToDevice));
      }
      e->overlap = true;
    }
  }
  e->parity = 0;
  const int welsh_smem_bytes = (int)(kVoiceWarps * kTileStride * sizeof(double2) + kParkWords * 32 * kVoiceWarps * sizeof(double));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_kernel<8, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, welsh_smem_bytes));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_kernel<8, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, welsh_smem_bytes));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_kernel<8, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_kernel<8, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_kernel<8, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_kernel<8, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_tp_kernel<8, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_tp_kernel<8, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_tp_kernel<8, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_tp_kernel<8, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_tp_kernel<8, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_tp_kernel<8, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_kernel<8, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_kernel<8, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_kernel<8, false, false, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_kernel<8, false, true, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_kernel<8, true, false, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_kernel<8, true, true, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_kernel<8, false, false, true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_rest_kernel<8, true, false, true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_sweep_kernel<8, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_sweep_kernel<8, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_sweep_kernel<8, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_sweep_kernel<8, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_sweep_kernel<8, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_sweep_kernel<8, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRestSmemMax));
  CUDA_TRY(e, cudaFuncSetAttribute(welsh_solo_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSoloSmemBytes));
  CUDA_TRY(e, cudaFuncSetAttribute(fm_kernel<kVoiceWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(kVoiceWarps * kTileStride * sizeof(double2))));
  {
    // staging buffers sized for a full chunk now: growing one inside a render call costs a device
    // synchronisation and fresh pinned host mirrors
    bool ok = e->wev_off.reserve((size_t)e->n_wvoice + 1) && e->fev_off.reserve((size_t)e->n_fvoice + 1) &&
              e->wev.reserve(4 * (size_t)e->n_wvoice + 64) && e->fev.reserve(4 * (size_t)e->n_fvoice + 64) &&
              e->widx.reserve((size_t)e->n_wwork_grouped + 1) && e->fxdescs.reserve(e->plan.size() + 1);
    if (ok && e->solo_mega) {
      const size_t kmax = (size_t)cdiv((int)e->max_block, e->solo_sub), n = e->witem_node.size();
      ok = e->sitems.reserve(n * kmax + 1) && e->sjobs.reserve(n * kmax / kVoiceWarps + kmax * SOLO_CLASSES + 1) &&
           e->szero.reserve(4 * n + 16);
    }
    if (!ok) return fail(e, GB_ENOMEM, "out of memory");
  }
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  e->finalized = true;
  e->winst_dirty = e->finst_dirty = true;
  {
    const int rc_up = upload_inst_tables(e);
    if (rc_up) return rc_up;
  }
  return plan_voice_ranges(e);
}

int gb_push_events(gb_engine* e, const gb_event* ev, size_t n) {
  if (!e || (!ev && n)) return GB_EINVAL;
  for (size_t i = 0; i < n; ++i) {
    if (ev[i].frame < e->pos) return fail(e, GB_EINVAL, "event frame %lld is in the past (position %lld)",
                                          (long long)ev[i].frame, (long long)e->pos);
    if (!find(e, ev[i].uid)) return fail(e, GB_ENOENT, "event targets unknown uid %u", ev[i].uid);
  }
  e->events.insert(e->events.end(), ev, ev + n);
  std::stable_sort(e->events.begin(), e->events.end(),
                   [](const gb_event& a, const gb_event& b) { return a.frame < b.frame; });
  return 0;
}

}  // extern "C"

namespace {

struct ControlPoint {
  int t;  // chunk-relative frame
  int index;
  double raw;
};

// Render one chunk of `frames` (<= max_block) frames starting at e->pos.  Events for the chunk are
// e->events[0 .. n_ev).
struct HostProfile {  // GB_HOST_PROFILE=1: wall time of render_chunk's host sections, to stderr
  bool on;
  std::chrono::steady_clock::time_point t;
  std::string line;
  HostProfile() : on(getenv("GB_HOST_PROFILE") != nullptr), t(std::chrono::steady_clock::now()) {}
  void mark(const char* what) {
    if (!on) return;
    auto n = std::chrono::steady_clock::now();
    char b[64];
    snprintf(b, sizeof b, " %s %.3f", what, std::chrono::duration<double, std::milli>(n - t).count());
    line += b;
    t = n;
  }
  ~HostProfile() { if (on) fprintf(stderr, "[host]%s\n", line.c_str()); }
};

int render_chunk(gb_engine* e, int frames, size_t n_ev) {
  const int64_t f0 = e->pos;
  HostProfile hp;
  // staging slot of this chunk: wait until the chunk that used it kStageSlots chunks ago has been consumed
  const int slot = (int)(e->chunk_seq % kStageSlots);
  if (e->chunk_seq >= (uint64_t)kStageSlots) CUDA_TRY(e, cudaEventSynchronize(e->stage_done[slot]));
  e->widx.use_slot(slot); e->wev.use_slot(slot); e->fev.use_slot(slot); e->wev_off.use_slot(slot);
  e->fev_off.use_slot(slot); e->plays.use_slot(slot); e->segs.use_slot(slot); e->fxdescs.use_slot(slot);
  // solo Welsh voices: note frames as they stand BEFORE this chunk's events (the sub-chunk classification
  // below replays the chunk's events on top of them)
  struct SoloPre { int64_t n_on, n_off; };
  std::vector<SoloPre> solo_pre;
  if (e->solo_mega) {
    solo_pre.resize(e->witem_node.size());
    for (size_t i = 0; i < solo_pre.size(); ++i) {
      const Node* n = e->witem_node[i];
      const Slot& sl = n->store.slots[(size_t)e->witem_slot[i]];
      solo_pre[i].n_on = sl.on_frame;
      solo_pre[i].n_off = sl.held ? kHeld : sl.idle_at - n->release_frames;
    }
  }
  hp.mark("pre");
  // ---- 1. resolve events ----
  std::vector<std::vector<VoiceEvent>> wlists((size_t)e->n_wvoice), flists((size_t)e->n_fvoice);
  std::map<Node*, std::vector<ControlPoint>> controls;
  std::map<Node*, std::vector<std::pair<int, Node::EnvHost>>> env_points;  // GB_INST_ENVELOPE: (chunk frame, state from there on)
  std::map<Node*, Node::EnvHost> env_entry;  // ... state at the start of the chunk
  for (Node* n : e->plan)
    if (n->kind == GB_INST_ENVELOPE) env_entry[n] = n->envh;
  bool any_w = false, any_f = false;
  for (size_t i = 0; i < n_ev; ++i) {
    const gb_event& ev = e->events[i];
    Node* n = find(e, ev.uid);
    if (!n || n->order < 0) continue;  // unpatched entities never render (orchestrator.rs:378-430)
    const int64_t f = std::max(ev.frame, f0);
    if (ev.type == GB_EV_NOTE_ON || ev.type == GB_EV_NOTE_OFF) {
      const bool on = ev.type == GB_EV_NOTE_ON;
      const int key = ev.a;
      if (n->kind == GB_INST_WELSH || n->kind == GB_INST_FM) {
        auto& lists = n->kind == GB_INST_WELSH ? wlists : flists;
        (n->kind == GB_INST_WELSH ? any_w : any_f) = true;
        if (on) {
          int v = n->store.note_on(f, key);
          VoiceEvent ve;
          ve.frame = f; ve.type = VEV_NOTE_ON; ve.pad = 0;
          if (n->kind == GB_INST_WELSH) {
            const gb_oscillator_params& o1 = n->wp.oscillator_1;
            const gb_oscillator_params& o2 = n->wp.oscillator_2;
            ve.cyc1 = (o1.fixed_frequency > 0.0 ? o1.fixed_frequency : note_hz(key) * o1.frequency_tune) / e->sr;
            ve.cyc2 = (o2.fixed_frequency > 0.0 ? o2.fixed_frequency : note_hz(key) * o2.frequency_tune) / e->sr;
          } else {
            ve.cyc1 = note_hz(key) / e->sr;
            ve.cyc2 = ve.cyc1 * n->fp.ratio;
          }
          lists[(size_t)(n->voice0 + v)].push_back(ve);
        } else {
          for (size_t s = 0; s < n->store.slots.size(); ++s) {
            Slot& sl = n->store.slots[s];
            if (sl.key == key && sl.held) {
              sl.held = false;
              sl.idle_at = f + n->release_frames;
              VoiceEvent ve;
              ve.frame = f; ve.type = VEV_NOTE_OFF; ve.pad = 0; ve.cyc1 = 0.0; ve.cyc2 = 0.0;
              lists[(size_t)n->voice0 + s].push_back(ve);
            }
          }
        }
      } else if (n->kind == GB_INST_SAMPLER) {
        if (n->samples.empty()) continue;
        if (on) {
          int v = n->store.note_on(f, key);
          const SampleDev& s = n->samples[0];
          SampleVoiceHost& sv = n->svoices[(size_t)v];
          // a retriggered voice's previous play ends here: keep it for this chunk if it overlaps it
          n->done_plays.resize(n->svoices.size());
          if (sv.sample >= 0 && sv.n_on < f && sv.n_end > f0) {
            SampleVoiceHost old = sv;
            old.n_end = std::min(old.n_end, f);
            n->done_plays[(size_t)v].push_back(old);
          }
          sv.sample = 0;
          sv.step_q = step_to_q32(note_hz(key) / s.root_hz * (s.sr / e->sr));
          sv.n_on = f;
          sv.n_end = f + frames_until_end(s.n, sv.step_q);
          n->store.slots[(size_t)v].idle_at = sv.n_end;
        } else {
          for (size_t s = 0; s < n->store.slots.size(); ++s) {
            Slot& sl = n->store.slots[s];
            if (sl.key == key && sl.held) {
              sl.held = false;
              sl.idle_at = f;
              if (f < n->svoices[s].n_end) n->svoices[s].n_end = f;
            }
          }
        }
      } else if (n->kind == GB_INST_ENVELOPE) {
        const EnvShape sh = make_shape(n->esp.envelope, e->sr);
        Node::EnvHost& h = n->envh;
        if (on) {
          h.l_on = h_env_level(sh, h.n_on, h.n_off, h.l_on, h.l_off, f);
          h.n_on = f;
          h.n_off = kHeld;
        } else {
          if (h.n_off != kHeld) continue;
          h.l_off = h_env_pre(sh, h.n_on, h.l_on, f);
          h.n_off = f;
        }
        env_points[n].push_back({(int)(f - f0), h});
      } else if (n->kind == GB_INST_DRUMKIT) {
        if (!on || key < 0 || key >= 128 || n->key_to_voice[key] < 0) continue;
        SampleVoiceHost& sv = n->svoices[(size_t)n->key_to_voice[key]];
        const SampleDev& s = n->samples[(size_t)sv.sample];
        n->done_plays.resize(n->svoices.size());
        if (sv.n_on < f && sv.n_end > f0 && sv.n_on > kNever) {
          SampleVoiceHost old = sv;
          old.n_end = std::min(old.n_end, f);
          n->done_plays[(size_t)n->key_to_voice[key]].push_back(old);
        }
        sv.step_q = step_to_q32(s.sr / e->sr);
        sv.n_on = f;
        sv.n_end = f + frames_until_end(s.n, sv.step_q);
      }
    } else if (ev.type == GB_EV_CONTROL || ev.type == GB_EV_SET_PARAM) {
      if (!effect_accepts(n, ev.a)) continue;
      ControlPoint cp;
      cp.t = (int)(f - f0);
      cp.index = ev.a;
      cp.raw = ev.type == GB_EV_CONTROL ? control_to_raw(n, ev.a, ev.value) : ev.value;
      controls[n].push_back(cp);
    }
  }
  e->events.erase(e->events.begin(), e->events.begin() + (long)n_ev);

  // Sampler retriggers inside a chunk do not split it: when a voice is restarted, its previous play is
  // truncated at the retrigger frame and kept in Node::done_plays for this chunk's play list (step 3),
  // so a voice restarted several times in one chunk contributes every one of its plays.

  // ---- 2. parameter segment tables: one per automated node, built on the host, one upload ----
  struct SegRange { size_t off; int n; };
  std::map<Node*, SegRange> seg_of;
  std::vector<SegParam> segs_h;
  auto apply_cp = [&](Node* n, const ControlPoint& cp) {
    if (n->kind == GB_INST_WELSH) {
      if (cp.index == GB_CTL_INST_DCA_GAIN) n->wp.dca.gain = cp.raw; else n->wp.dca.pan = cp.raw;
      e->winst_dirty = true;
    } else if (n->kind == GB_INST_FM) {
      if (cp.index == GB_CTL_INST_DCA_GAIN) n->fp.dca.gain = cp.raw; else n->fp.dca.pan = cp.raw;
      e->finst_dirty = true;
    } else {
      effect_apply_param(e, n, cp.index, cp.raw);
    }
  };
  auto push_seg = [&](Node* n, int t) {
    SegParam sp;
    sp.t0 = t; sp.pad = 0;
    seg_values(e, n, sp.v);
    segs_h.push_back(sp);
  };
  for (Node* n : e->plan) {
    const bool inst = n->kind == GB_INST_WELSH || n->kind == GB_INST_FM;
    if (n->is_inst && !inst) continue;
    if (!n->is_inst && (n->kind == GB_FX_MIXER || n->kind == GB_FX_DELAY || n->kind == GB_FX_SIGNAL_PASSTHROUGH))
      continue;  // no parameters
    std::vector<ControlPoint> cps;
    auto it = controls.find(n);
    if (it != controls.end()) cps = it->second;
    bool linked = false;
    for (auto& l : e->links) linked |= l.dst == n->uid && l.d_table;
    if (linked) {  // the table of a link target is built on the device (plan walk); host points apply up front
      for (auto& cp : cps) apply_cp(n, cp);
      continue;
    }
    std::stable_sort(cps.begin(), cps.end(), [](const ControlPoint& a, const ControlPoint& b) { return a.t < b.t; });
    size_t ci = 0;
    while (ci < cps.size() && cps[ci].t <= 0) apply_cp(n, cps[ci++]);  // events at the chunk start
    if (inst && ci == cps.size()) continue;  // no automation inside the chunk: the kernel applies the DCA
    SegRange r;
    r.off = segs_h.size();
    push_seg(n, 0);
    while (ci < cps.size()) {
      const int t = cps[ci].t;
      while (ci < cps.size() && cps[ci].t == t) apply_cp(n, cps[ci++]);
      push_seg(n, t);
    }
    r.n = (int)(segs_h.size() - r.off);
    seg_of[n] = r;
    if (inst) {  // render at unit gain, then one segmented gain/pan pass over the instrument's buffer
      n->unit_gain = true;
      (n->kind == GB_INST_WELSH ? e->winst_dirty : e->finst_dirty) = true;
    }
  }
  for (Node* n : e->plan) {  // envelope devices: one segment per note event of the chunk
    if (n->kind != GB_INST_ENVELOPE) continue;
    SegRange r;
    r.off = segs_h.size();
    auto seg = [&](int t, const Node::EnvHost& h) {
      SegParam sp;
      sp.t0 = t; sp.pad = 0;
      sp.v[0] = h.l_on; sp.v[1] = h.l_off; sp.v[2] = (double)h.n_on; sp.v[3] = (double)h.n_off; sp.v[4] = 0.0; sp.v[5] = 0.0;
      if (t == 0 && segs_h.size() > r.off) segs_h.back() = sp;  // an event on the chunk's first frame replaces the entry state
      else segs_h.push_back(sp);
    };
    seg(0, env_entry[n]);
    auto it = env_points.find(n);
    if (it != env_points.end())
      for (auto& pt : it->second) seg(pt.first, pt.second);
    r.n = (int)(segs_h.size() - r.off);
    seg_of[n] = r;
  }
  if (!segs_h.empty()) {
    if (!e->segs.reserve(segs_h.size())) return fail(e, GB_ENOMEM, "out of memory");
    memcpy(e->segs.h, segs_h.data(), segs_h.size() * sizeof(SegParam));
    CUDA_TRY(e, cudaMemcpyAsync(e->segs.d, e->segs.h, segs_h.size() * sizeof(SegParam), cudaMemcpyHostToDevice, e->stream));
    e->stats.h2d_bytes += segs_h.size() * sizeof(SegParam);
  }

  // ---- 3. voices ----
  auto upload_events = [&](std::vector<std::vector<VoiceEvent>>& lists, DevBuf<VoiceEvent>& evb, DevBuf<int>& offb,
                           bool any, bool* empty_on_device, cudaStream_t st) -> int {
    size_t nv = lists.size(), total = 0;
    // a chunk without note events needs the all-zero offset table: if that is what the device already
    // holds, nothing is uploaded (every small copy costs the stream several microseconds)
    if (!any && *empty_on_device && offb.cap >= nv + 1) return 0;
    *empty_on_device = !any;
    if (!offb.reserve(nv + 1)) return fail(e, GB_ENOMEM, "out of memory");
    if (any)
      for (auto& l : lists) total += l.size();
    if (!evb.reserve(total + 1)) return fail(e, GB_ENOMEM, "out of memory");
    size_t k = 0;
    for (size_t v = 0; v < nv; ++v) {
      offb.h[v] = (int)k;
      if (any)
        for (auto& x : lists[v]) evb.h[k++] = x;
    }
    offb.h[nv] = (int)k;
    CUDA_TRY(e, cudaMemcpyAsync(offb.d, offb.h, (nv + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    e->stats.h2d_bytes += (nv + 1) * sizeof(int);
    if (k) {
      CUDA_TRY(e, cudaMemcpyAsync(evb.d, evb.h, k * sizeof(VoiceEvent), cudaMemcpyHostToDevice, st));
      e->stats.h2d_bytes += k * sizeof(VoiceEvent);
    }
    return 0;
  };
  const size_t tile_bytes = (size_t)kVoiceWarps * kTileStride * sizeof(double2);
  hp.mark("events");
  // Overlapped mixdown: this chunk's grouped voice kernels go to vstream and write the partial buffers of the
  // chunk's parity; they wait for the table sum that read those buffers two chunks ago.
  const int par = e->overlap ? (int)(e->chunk_seq & 1) : 0;
  e->parity = par;
  e->chunk_vr = false;
  e->chunk_all_idle = false;
  cudaStream_t vs = e->overlap ? e->vstream : e->stream;
  if (e->overlap) {
    CUDA_TRY(e, cudaStreamWaitEvent(vs, e->call_start, 0));
    CUDA_TRY(e, cudaStreamWaitEvent(vs, e->p_done[par], 0));
  }
  if (e->n_wvoice) {
    if (e->winst_dirty) {
      if (e->overlap) CUDA_TRY(e, cudaStreamSynchronize(e->vstream));  // earlier chunks' kernels still read the table
      int rc0 = upload_inst_tables(e);
      if (rc0) return rc0;
    }
    int rc = upload_events(wlists, e->wev, e->wev_off, any_w, &e->wev_empty_on_device, vs);
    if (rc) return rc;
    {
      // Sort the grouped CTAs of this chunk: a CTA whose voices all rest for the whole chunk (note held
      // since before the chunk with both envelopes at their sustain levels, no note event inside it)
      // goes to welsh_rest_kernel, the others to welsh_kernel.  Host knowledge only: note frames are
      // integers tracked by the slot stores.
      const int ng = e->n_wwork_grouped, ns = e->n_wwork - e->n_wwork_grouped;
      constexpr int kVar = 6, kTp = kVar, kGen = 2 * kVar, kSw = 2 * kVar + 1;  // kernel variants: 0..3 = (lfo, flat oscillators), 4..5 = hard sync (lfo)
      std::vector<int> lists[3 * kVar + 1];  // 0..5 = resting variants, 6..11 = resting, time-parallel, 12 = general, 13..18 = sweeping (idle CTAs are not launched)
      size_t tp_voices_max = 0;
      e->wwork_zero.resize((size_t)ng, 0);
      e->wwork_zero_alt.resize((size_t)ng, 0);
      std::vector<char>& wzero = par ? e->wwork_zero_alt : e->wwork_zero;
      const CtaWork* wwd = par ? e->d_wwork_alt : e->wwork.d;
      const bool chunk_ok = frames % kBlockFrames == 0;
      size_t rest_voices_max = 0, sweep_voices_max = 0;
      uint64_t rest_voices = 0, sweep_voices = 0;
      for (int i = 0; i < ng; ++i) {
        const CtaWork& w = e->wwork.h[i];
        const Node* n = e->wwork_node[(size_t)i];
        const WelshInst& I = e->h_winst[(size_t)n->table_index];
        bool rest = chunk_ok && I.rest_class >= 0 &&
                    kVoiceWarps * kTileStride * sizeof(double2) + (size_t)w.nvoices * sizeof(RestState) <= (size_t)kRestSmemMax;
        for (int v = 0; rest && v < w.nvoices; ++v) {
          const Slot& sl = n->store.slots[(size_t)(w.voice0 - n->voice0 + v)];
          rest = wlists[(size_t)(w.voice0 + v)].empty() && sl.held && sl.on_frame > kNever &&
                 sl.on_frame + I.steady_after <= f0;
        }
        // idle: every voice silent for the whole chunk and no note event in it — the CTA's output is zeros
        bool idle = !rest;
        for (int v = 0; idle && v < w.nvoices; ++v) {
          const Slot& sl = n->store.slots[(size_t)(w.voice0 - n->voice0 + v)];
          idle = wlists[(size_t)(w.voice0 + v)].empty() && !sl.held && f0 >= sl.idle_at;
        }
        if (idle) {
          e->stats.idle_voice_samples += (uint64_t)w.nvoices * (uint64_t)frames;
          if (!wzero[(size_t)i]) {
            double2* zout = par ? n->scratch_alt + (w.out - n->scratch) : w.out;
            CUDA_TRY(e, cudaMemsetAsync(zout, 0, (size_t)e->max_block * sizeof(double2), vs));
            wzero[(size_t)i] = 1;
          }
          continue;
        }
        wzero[(size_t)i] = 0;
        // sweeping: every voice held since before the chunk, no note event in it, the filter envelope inside one
        // moving stage and the amplitude envelope inside one stage for the whole chunk, the cutoff slow enough
        // for coefficient knots (the bound uses the stage's steepest slope)
        bool sweep = !rest && chunk_ok && I.sweep_class >= 0 &&
                     kVoiceWarps * kTileStride * sizeof(double2) + (size_t)w.nvoices * sizeof(SweepState) <= (size_t)kRestSmemMax;
        auto stage_of = [](const EnvShape& sh, int64_t k) { return k < sh.na ? 0 : (k - sh.na < sh.nd ? 1 : 2); };
        for (int v = 0; sweep && v < w.nvoices; ++v) {
          const Slot& sl = n->store.slots[(size_t)(w.voice0 - n->voice0 + v)];
          sweep = wlists[(size_t)(w.voice0 + v)].empty() && sl.held && sl.on_frame > kNever && sl.on_frame <= f0;
          if (!sweep) break;
          const int64_t k0 = f0 - sl.on_frame, k1 = k0 + frames - 1;
          const int sa = stage_of(I.amp, k0), sf = stage_of(I.filt, k0);
          sweep = sa == stage_of(I.amp, k1) && sf == stage_of(I.filt, k1) && sf != 2;
          if (sweep) {
            const double slope = sf == 0 ? 2.0 * I.filt.inv_na : 2.0 * (1.0 - I.filt.sustain) * I.filt.inv_nd;
            sweep = std::fabs(I.cut_b) * slope <= I.knot_max_rate;
          }
        }
        if (sweep) {
          lists[kSw + I.sweep_class].push_back(i);
          sweep_voices_max = std::max(sweep_voices_max, (size_t)w.nvoices);
          sweep_voices += (uint64_t)w.nvoices;
          continue;
        }
        const bool tp = rest && e->opt.rest_tp && w.nvoices <= kTpMaxVoices;
        lists[rest ? (tp ? kTp : 0) + I.rest_class : kGen].push_back(i);
        if (rest) {
          if (tp) tp_voices_max = std::max(tp_voices_max, (size_t)w.nvoices);
          else rest_voices_max = std::max(rest_voices_max, (size_t)w.nvoices);
          rest_voices += (uint64_t)w.nvoices;
        }
      }
      {
        size_t launched = 0;
        for (auto& l : lists) launched += l.size();
        e->chunk_all_idle = ng > 0 && launched == 0;
      }
      if (ng) {
        if (e->widx.cap < (size_t)ng) e->widx_on_device.clear();  // a new device buffer holds nothing yet
        if (!e->widx.reserve((size_t)ng)) return fail(e, GB_ENOMEM, "out of memory");
        std::vector<int> flat;
        for (auto& l : lists) flat.insert(flat.end(), l.begin(), l.end());
        const size_t k = flat.size();
        if (k && flat != e->widx_on_device) {
          memcpy(e->widx.h, flat.data(), k * sizeof(int));
          CUDA_TRY(e, cudaMemcpyAsync(e->widx.d, e->widx.h, k * sizeof(int), cudaMemcpyHostToDevice, vs));
          e->stats.h2d_bytes += k * sizeof(int);
          e->widx_on_device = flat;
        }
      }
      const size_t rest_smem = (size_t)kVoiceWarps * kTileStride * sizeof(double2) + rest_voices_max * sizeof(RestState);
      size_t off = 0;
      // every grouped CTA rests in this chunk: the voice ranges take it (the consumer then sums the range partials)
      if (e->vr_ok && ng > 0 && lists[e->vr_class].size() + lists[kTp + e->vr_class].size() == (size_t)ng && e->fused_sums_enabled) {
        bool plain = true;
        for (Node* n : e->plan) plain = plain && !n->unit_gain;
        if (plain) {
          constexpr int kVrW = 8;
          Launch l(e, true, 1, vs);
          if (e->opt.rest16 && frames % kBlock16 == 0) {  // 16 frames per lane: half the scans and barriers per frame
            e->stats.rest_vr16_launches++;
            const size_t smem16 = (size_t)kVrW * kTile16Stride * sizeof(double2) + 14 * sizeof(RestState);
            switch (e->vr_class) {
              case 0: welsh_rest_vr16_kernel<kVrW, false, false><<<e->vr_count, 32 * kVrW, smem16, vs>>>(e->d_winst, e->d_rest16, e->d_wvoice, e->d_vr_work[par], f0, frames); break;
              case 1:
                if (e->vr_osc == 2) welsh_rest_vr16_kernel<kVrW, false, true, 2><<<e->vr_count, 32 * kVrW, smem16, vs>>>(e->d_winst, e->d_rest16, e->d_wvoice, e->d_vr_work[par], f0, frames);
                else if (e->vr_osc == 1) welsh_rest_vr16_kernel<kVrW, false, true, 1><<<e->vr_count, 32 * kVrW, smem16, vs>>>(e->d_winst, e->d_rest16, e->d_wvoice, e->d_vr_work[par], f0, frames);
                else welsh_rest_vr16_kernel<kVrW, false, true><<<e->vr_count, 32 * kVrW, smem16, vs>>>(e->d_winst, e->d_rest16, e->d_wvoice, e->d_vr_work[par], f0, frames);
                break;
              case 2: welsh_rest_vr16_kernel<kVrW, true, false><<<e->vr_count, 32 * kVrW, smem16, vs>>>(e->d_winst, e->d_rest16, e->d_wvoice, e->d_vr_work[par], f0, frames); break;
              default:
                if (e->vr_osc == 2) welsh_rest_vr16_kernel<kVrW, true, true, 2><<<e->vr_count, 32 * kVrW, smem16, vs>>>(e->d_winst, e->d_rest16, e->d_wvoice, e->d_vr_work[par], f0, frames);
                else if (e->vr_osc == 1) welsh_rest_vr16_kernel<kVrW, true, true, 1><<<e->vr_count, 32 * kVrW, smem16, vs>>>(e->d_winst, e->d_rest16, e->d_wvoice, e->d_vr_work[par], f0, frames);
                else welsh_rest_vr16_kernel<kVrW, true, true><<<e->vr_count, 32 * kVrW, smem16, vs>>>(e->d_winst, e->d_rest16, e->d_wvoice, e->d_vr_work[par], f0, frames);
                break;
            }
          } else {
            const size_t vr_smem = (size_t)kVrW * kTileStride * sizeof(double2) + 14 * sizeof(RestState);
            switch (e->vr_class) {
              case 0: welsh_rest_vr_kernel<kVrW, false, false><<<e->vr_count, 32 * kVrW, vr_smem, vs>>>(e->d_winst, e->d_wvoice, e->d_vr_work[par], f0, frames); break;
              case 1: welsh_rest_vr_kernel<kVrW, false, true><<<e->vr_count, 32 * kVrW, vr_smem, vs>>>(e->d_winst, e->d_wvoice, e->d_vr_work[par], f0, frames); break;
              case 2: welsh_rest_vr_kernel<kVrW, true, false><<<e->vr_count, 32 * kVrW, vr_smem, vs>>>(e->d_winst, e->d_wvoice, e->d_vr_work[par], f0, frames); break;
              default: welsh_rest_vr_kernel<kVrW, true, true><<<e->vr_count, 32 * kVrW, vr_smem, vs>>>(e->d_winst, e->d_wvoice, e->d_vr_work[par], f0, frames); break;
            }
          }
          e->stats.rest_ctas += (uint64_t)e->vr_count;
          e->stats.rest_vr_launches++;
          e->chunk_vr = true;
          lists[e->vr_class].clear();
          lists[kTp + e->vr_class].clear();
        }
      }
#define GB_REST_LAUNCH(CLS_, LFO_, FLAT_, SYNC_)                                                                     \
  if (!lists[CLS_].empty()) {                                                                                        \
    Launch l(e, true, 1, vs);                                                                                        \
    if (e->opt.rest_nv == 4)                                                                                         \
      welsh_rest_kernel<8, LFO_, FLAT_, SYNC_, 4><<<(int)lists[CLS_].size(), 32 * 8, rest_smem, vs>>>(                 \
          e->d_winst, e->d_wvoice, wwd, e->widx.d + off, f0, frames);                                                \
    else                                                                                                             \
      welsh_rest_kernel<8, LFO_, FLAT_, SYNC_><<<(int)lists[CLS_].size(), 32 * 8, rest_smem, vs>>>(                    \
          e->d_winst, e->d_wvoice, wwd, e->widx.d + off, f0, frames);                                                \
    off += lists[CLS_].size();                                                                                       \
    e->stats.rest_ctas += lists[CLS_].size();                                                                        \
  }
      GB_REST_LAUNCH(0, false, false, false)
      GB_REST_LAUNCH(1, false, true, false)
      GB_REST_LAUNCH(2, true, false, false)
      GB_REST_LAUNCH(3, true, true, false)
      GB_REST_LAUNCH(4, false, false, true)
      GB_REST_LAUNCH(5, true, false, true)
#undef GB_REST_LAUNCH
      const size_t tp_smem = (size_t)kVoiceWarps * kTileStride * sizeof(double2) +
                             tp_voices_max * (sizeof(TpState) + kVoiceWarps * sizeof(TpPriv));
#define GB_REST_TP_LAUNCH(CLS_, LFO_, FLAT_, SYNC_)                                                                  \
  if (!lists[kTp + CLS_].empty()) {                                                                                  \
    Launch l(e, true, 1, vs);                                                                                        \
    welsh_rest_tp_kernel<8, LFO_, FLAT_, SYNC_><<<(int)lists[kTp + CLS_].size(), 32 * 8, tp_smem, vs>>>(               \
        e->d_winst, e->d_wvoice, wwd, e->widx.d + off, f0, frames, (int)tp_voices_max);                              \
    off += lists[kTp + CLS_].size();                                                                                 \
    e->stats.rest_ctas += lists[kTp + CLS_].size();                                                                  \
    e->stats.rest_tp_launches++;                                                                                     \
  }
      GB_REST_TP_LAUNCH(0, false, false, false)
      GB_REST_TP_LAUNCH(1, false, true, false)
      GB_REST_TP_LAUNCH(2, true, false, false)
      GB_REST_TP_LAUNCH(3, true, true, false)
      GB_REST_TP_LAUNCH(4, false, false, true)
      GB_REST_TP_LAUNCH(5, true, false, true)
#undef GB_REST_TP_LAUNCH
      // every grouped CTA sweeps in this chunk: the voice ranges take it, as they take all-resting chunks
      if (e->vr_ok && e->vr_sweep_class >= 0 && ng > 0 && lists[kSw + e->vr_sweep_class].size() == (size_t)ng &&
          e->fused_sums_enabled) {
        bool plain = true;
        for (Node* n : e->plan) plain = plain && !n->unit_gain;
        if (plain) {
          constexpr int kVrW = 8;
          const size_t smem = (size_t)kVrW * kTileStride * sizeof(double2) + 14 * sizeof(SweepState);
          Launch l(e, true, 2, vs);
          switch (e->vr_sweep_class) {
            case 0: welsh_sweep_vr_kernel<kVrW, false, false><<<e->vr_count, 32 * kVrW, smem, vs>>>(e->d_winst, e->d_wvoice, e->d_vr_work[par], f0, frames); break;
            case 1: welsh_sweep_vr_kernel<kVrW, false, true><<<e->vr_count, 32 * kVrW, smem, vs>>>(e->d_winst, e->d_wvoice, e->d_vr_work[par], f0, frames); break;
            case 2: welsh_sweep_vr_kernel<kVrW, true, false><<<e->vr_count, 32 * kVrW, smem, vs>>>(e->d_winst, e->d_wvoice, e->d_vr_work[par], f0, frames); break;
            default: welsh_sweep_vr_kernel<kVrW, true, true><<<e->vr_count, 32 * kVrW, smem, vs>>>(e->d_winst, e->d_wvoice, e->d_vr_work[par], f0, frames); break;
          }
          e->stats.sweep_ctas += (uint64_t)e->vr_count;
          e->chunk_vr = true;
          lists[kSw + e->vr_sweep_class].clear();
        }
      }
      e->stats.rest_voice_samples += rest_voices * (uint64_t)frames;
      e->stats.sweep_voice_samples += sweep_voices * (uint64_t)frames;
      const size_t welsh_smem = (size_t)kVoiceWarps * kTileStride * sizeof(double2) + (size_t)kParkWords * 32 * kVoiceWarps * sizeof(double);
      if (!lists[kGen].empty()) {
        Launch l(e, true, 0, vs);
        welsh_kernel<8, 2, false><<<(int)lists[kGen].size(), 32 * 8, welsh_smem, vs>>>(
            e->d_winst, e->d_wvoice, wwd, e->witems.d, e->wev.d, e->wev_off.d, f0, frames, e->widx.d + off);
      }
      off += lists[kGen].size();
      const size_t sweep_smem = (size_t)kVoiceWarps * kTileStride * sizeof(double2) + sweep_voices_max * sizeof(SweepState);
#define GB_SWEEP_LAUNCH(CLS_, LFO_, FLAT_, SYNC_)                                                                    \
  if (!lists[kSw + CLS_].empty()) {                                                                                  \
    Launch l(e, true, 2, vs);                                                                                        \
    welsh_sweep_kernel<8, LFO_, FLAT_, SYNC_><<<(int)lists[kSw + CLS_].size(), 32 * 8, sweep_smem, vs>>>(              \
        e->d_winst, e->d_wvoice, wwd, e->widx.d + off, f0, frames);                                                  \
    off += lists[kSw + CLS_].size();                                                                                 \
    e->stats.sweep_ctas += lists[kSw + CLS_].size();                                                                 \
  }
      GB_SWEEP_LAUNCH(0, false, false, false)
      GB_SWEEP_LAUNCH(1, false, true, false)
      GB_SWEEP_LAUNCH(2, true, false, false)
      GB_SWEEP_LAUNCH(3, true, true, false)
      GB_SWEEP_LAUNCH(4, false, false, true)
      GB_SWEEP_LAUNCH(5, true, false, true)
#undef GB_SWEEP_LAUNCH
      if (ns && !e->solo_mega) {
        Launch l(e, true);
        welsh_kernel<8, 2, true><<<ns, 32 * 8, welsh_smem, e->stream>>>(
            e->d_winst, e->d_wvoice, e->wwork.d + ng, e->witems.d, e->wev.d, e->wev_off.d, f0, frames, nullptr);
      }
    }
    e->stats.voice_samples += (uint64_t)e->n_wvoice * (uint64_t)frames;
  }
  if (e->overlap) {  // the rest of the chunk (table sums, effects, copy-out) follows this chunk's voice kernels
    CUDA_TRY(e, cudaEventRecord(e->v_done[par], vs));
    CUDA_TRY(e, cudaStreamWaitEvent(e->stream, e->v_done[par], 0));
  }
  // Solo Welsh items through the job list (welsh_solo_kernel).  Prepared AFTER the FM launch below has been
  // enqueued, so that the host-side classification overlaps with GPU work of the same chunk.
  auto run_solo = [&]() -> int {
    if (!e->solo_mega) return 0;
    const int n_items = (int)e->witem_node.size();
    const int kSoloSub = e->solo_sub;  // frames per sub-chunk (classification granularity)
    const int K = cdiv(frames, kSoloSub);
    const bool dbg = getenv("GB_DEBUG") != nullptr;
    if (dbg) { fprintf(stderr, "[solo] f0=%lld frames=%d items=%d K=%d\n", (long long)f0, frames, n_items, K); fflush(stderr); }
    enum { C_IDLE = SOLO_CLASSES, NC = SOLO_CLASSES };
    std::vector<int8_t> cls((size_t)n_items * (size_t)K);
    auto classify = [&](const WelshInst& I, int64_t n_on, int64_t n_off, int64_t s0, int64_t s1, int64_t* valid_until) -> int {
      // *valid_until: the class holds for every later sub-chunk that ends at or before this frame
      *valid_until = s1;
      if (n_on <= kNever) { *valid_until = kHeld; return C_IDLE; }
      const int64_t idle_at = n_off >= kHeld ? (int64_t)kHeld : n_off + I.amp.nr;
      if (s0 >= idle_at) { *valid_until = kHeld; return C_IDLE; }
      if (s1 > idle_at || s0 < n_on || (n_off > s0 && n_off < s1) || (s1 - s0) % kBlockFrames) return SOLO_GENERAL;
      const bool released = s0 >= n_off;
      const int64_t base = released ? n_off : n_on;
      auto stage = [&](const EnvShape& sh, int64_t f, int64_t* next) {  // stage at f and the frame at which it ends
        const int64_t k = f - base;
        if (released) {
          if (k < sh.nr) { *next = base + sh.nr; return 3; }
          *next = kHeld; return 4;
        }
        if (k < sh.na) { *next = base + sh.na; return 0; }
        if (k - sh.na < sh.nd) { *next = base + sh.na + sh.nd; return 1; }
        *next = kHeld; return 2;
      };
      int64_t na_end, nf_end = kHeld;
      const int sa = stage(I.amp, s0, &na_end);
      int64_t until = std::min<int64_t>(std::min<int64_t>(na_end, idle_at), released ? (int64_t)kHeld : n_off);
      if (s1 > na_end) return SOLO_GENERAL;
      int result;
      if (I.filter_mode == FILTER_FIXED) {
        result = I.rest_class >= 0 && !I.sync ? SOLO_REST : SOLO_GENERAL;
      } else if (I.filter_mode != FILTER_ENVELOPE) {
        result = SOLO_GENERAL;
      } else {
        const int sf = stage(I.filt, s0, &nf_end);
        if (s1 > nf_end) return SOLO_GENERAL;
        until = std::min<int64_t>(until, nf_end);
        if (sf == 2 || sf == 4) {
          result = I.rest_class >= 0 && !I.sync ? SOLO_REST : SOLO_GENERAL;
        } else {
          const double slope = sf == 0 ? 2.0 * I.filt.inv_na : sf == 1 ? 2.0 * (1.0 - I.filt.sustain) * I.filt.inv_nd
                                                                       : 2.0 * I.filt.inv_nr;
          result = I.sweep_class >= 0 && !I.sync && std::fabs(I.cut_b) * slope <= I.knot_max_rate ? SOLO_SWEEP
                   : I.exact_class >= 0 ? SOLO_EXACT : SOLO_GENERAL;
        }
      }
      (void)sa;
      *valid_until = until;
      return result;
    };
    int counts[NC + 1] = {0, 0, 0, 0, 0};
    std::vector<int> bucket_n((size_t)K * NC, 0);
    for (int i = 0; i < n_items; ++i) {
      const Node* n = e->witem_node[(size_t)i];
      const WelshInst& I = e->h_winst[(size_t)n->table_index];
      const auto& evs = wlists[(size_t)(n->voice0 + e->witem_slot[(size_t)i])];
      int64_t n_on = solo_pre[(size_t)i].n_on, n_off = solo_pre[(size_t)i].n_off;
      size_t ei = 0;
      int8_t* c = cls.data() + (size_t)i * (size_t)K;
      int cur = -1;
      int64_t valid_until = 0;
      for (int k = 0; k < K; ++k) {
        const int64_t s0 = f0 + (int64_t)k * kSoloSub, s1 = std::min<int64_t>(s0 + kSoloSub, f0 + frames);
        int r;
        if (ei < evs.size() && evs[ei].frame < s1) {  // note events inside: the general class folds them
          while (ei < evs.size() && evs[ei].frame < s1) {
            if (evs[ei].type == VEV_NOTE_ON) { n_on = evs[ei].frame; n_off = kHeld; }
            else n_off = evs[ei].frame;
            ++ei;
          }
          r = SOLO_GENERAL;
          cur = -1;
        } else if (cur >= 0 && s1 <= valid_until) {
          r = cur;
        } else {
          r = classify(I, n_on, n_off, s0, s1, &valid_until);
          cur = r == SOLO_GENERAL ? -1 : r;
        }
        c[k] = (int8_t)r;
        counts[r]++;
        if (r != C_IDLE) bucket_n[(size_t)k * NC + (size_t)r]++;
      }
    }
    // buckets in job order: sub-chunk by sub-chunk; inside one the slowest class first
    static const int order[NC] = {SOLO_GENERAL, SOLO_EXACT, SOLO_SWEEP, SOLO_REST};
    std::vector<int> bucket_off((size_t)K * NC + 1, 0);
    {
      int acc = 0;
      for (int k = 0; k < K; ++k)
        for (int o = 0; o < NC; ++o) {
          bucket_off[(size_t)k * NC + (size_t)order[o]] = acc;
          acc += bucket_n[(size_t)k * NC + (size_t)order[o]];
        }
      bucket_off[(size_t)K * NC] = acc;
    }
    const int n_sitems = bucket_off[(size_t)K * NC];
    if (!e->sitems.reserve((size_t)n_sitems + 1)) return fail(e, GB_ENOMEM, "out of memory");
    e->sitems.use_slot(slot);
    {
      std::vector<int> fill(bucket_off.begin(), bucket_off.end() - 1);
      std::vector<int> done((size_t)n_items, 0);
      for (int k = 0; k < K; ++k)
        for (int i = 0; i < n_items; ++i) {
          const int r = cls[(size_t)i * (size_t)K + (size_t)k];
          if (r == C_IDLE) continue;
          SoloItem& si = e->sitems.h[fill[(size_t)k * NC + (size_t)r]++];
          si.item = i;
          si.need = done[(size_t)i]++;
        }
    }
    std::vector<SoloJob> jobs;
    std::vector<int> wave_first;  // first job of each sub-chunk (GB_SOLO_WAVES)
    for (int k = 0; k < K; ++k) {
      wave_first.push_back((int)jobs.size());
      for (int o = 0; o < NC; ++o) {
        const int r = order[o];
        const int first = bucket_off[(size_t)k * NC + (size_t)r], cnt = bucket_n[(size_t)k * NC + (size_t)r];
        for (int a = 0; a < cnt; a += kVoiceWarps) {
          SoloJob j;
          j.cls = r;
          j.t0 = k * kSoloSub;
          j.nframes = std::min(kSoloSub, frames - j.t0);
          j.first = first + a;
          j.n = std::min(kVoiceWarps, cnt - a);
          j.lockstep = e->solo_lockstep >> (r == SOLO_GENERAL ? 1 : 0) & 1;
          jobs.push_back(j);
        }
      }
    }
    wave_first.push_back((int)jobs.size());
    // idle stretches: the item's output must read zero there; only stretches that may hold old audio are written
    std::vector<ZeroRange> zr;
    for (int i = 0; i < n_items; ++i) {
      const int8_t* c = cls.data() + (size_t)i * (size_t)K;
      gb_engine::SoloDirty& d = e->solo_dirty[(size_t)i];
      int lo = INT32_MAX, hi = 0;
      for (int k = 0; k < K;) {
        const bool idle = c[k] == C_IDLE;
        int k1 = k;
        while (k1 < K && (c[k1] == C_IDLE) == idle) ++k1;
        const int a = k * kSoloSub, b = std::min(k1 * kSoloSub, frames);
        if (idle) {
          const int za = std::max(a, d.lo), zb = std::min(b, d.hi);
          if (za < zb) {
            ZeroRange z;
            z.p = e->witems.h[i].out + za;
            z.n = zb - za;
            zr.push_back(z);
          }
        } else {
          lo = std::min(lo, a);
          hi = std::max(hi, b);
        }
        k = k1;
      }
      // after this chunk [0, frames) holds what this chunk wrote; beyond it the old content stays
      int nlo = lo, nhi = hi;
      if (d.hi > frames) { nlo = std::min(nlo, std::max(d.lo, frames)); nhi = std::max(nhi, d.hi); }
      if (nlo >= nhi) { nlo = 0; nhi = 0; }
      d.lo = nlo; d.hi = nhi;
    }
    if (!zr.empty()) {
      e->szero.use_slot(slot);
      if (!e->szero.reserve(zr.size())) return fail(e, GB_ENOMEM, "out of memory");
      memcpy(e->szero.h, zr.data(), zr.size() * sizeof(ZeroRange));
      CUDA_TRY(e, cudaMemcpyAsync(e->szero.d, e->szero.h, zr.size() * sizeof(ZeroRange), cudaMemcpyHostToDevice, e->stream));
      e->stats.h2d_bytes += zr.size() * sizeof(ZeroRange);
      Launch l(e, false);
      zero_ranges_kernel<<<(int)zr.size(), 256, 0, e->stream>>>(e->szero.d);
    }
    e->stats.idle_voice_samples += (uint64_t)counts[C_IDLE] * (uint64_t)kSoloSub;
    if (jobs.empty()) return 0;
    e->sjobs.use_slot(slot);
    if (!e->sjobs.reserve(jobs.size())) return fail(e, GB_ENOMEM, "out of memory");
    memcpy(e->sjobs.h, jobs.data(), jobs.size() * sizeof(SoloJob));
    CUDA_TRY(e, cudaMemcpyAsync(e->sitems.d, e->sitems.h, (size_t)n_sitems * sizeof(SoloItem), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(e, cudaMemcpyAsync(e->sjobs.d, e->sjobs.h, jobs.size() * sizeof(SoloJob), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(e, cudaMemsetAsync(e->d_solo_sync, 0, ((size_t)e->n_wvoice + kSoloTickets) * sizeof(int), e->stream));
    e->stats.h2d_bytes += (size_t)n_sitems * sizeof(SoloItem) + jobs.size() * sizeof(SoloJob);
    if (dbg) {
      fprintf(stderr, "[solo] jobs=%zu sitems=%d counts=%d/%d/%d/%d idle=%d zr=%zu\n", jobs.size(), n_sitems, counts[0], counts[1],
              counts[2], counts[3], counts[C_IDLE], zr.size());
      fflush(stderr);
    }
    // one persistent launch walks the whole list; GB_SOLO_WAVES=1 launches once per sub-chunk instead (the
    // jobs of one sub-chunk are independent, stream order replaces the progress counters)
    {
      std::vector<std::pair<int, int>> launches;  // (first job, jobs)
      if (e->solo_waves && K <= kSoloTickets) {
        for (int k = 0; k < K; ++k)
          if (wave_first[(size_t)k + 1] > wave_first[(size_t)k])
            launches.push_back({wave_first[(size_t)k], wave_first[(size_t)k + 1] - wave_first[(size_t)k]});
      } else {
        launches.push_back({0, (int)jobs.size()});
      }
      int max_grid = 2 * e->num_sms;
      if (const char* v = getenv("GB_SOLO_GRID")) max_grid = std::max(1, atoi(v));
      for (size_t li = 0; li < launches.size(); ++li) {
        Launch l(e, true, 3);
        const int grid = std::min(launches[li].second, max_grid);
        welsh_solo_kernel<8><<<grid, 32 * 8, kSoloSmemBytes, e->stream>>>(
            e->d_winst, e->d_wvoice, e->witems.d, e->sitems.d, e->sjobs.d, launches[li].first, launches[li].second,
            e->d_solo_sync + li, e->d_solo_sync + kSoloTickets, e->d_solo_fault, e->wev.d, e->wev_off.d, f0, frames);
      }
      CUDA_TRY(e, cudaGetLastError());
      e->solo_ran = true;
      if (dbg) {
        cudaError_t q = cudaErrorNotReady;
        for (int it = 0; it < 800 && q == cudaErrorNotReady; ++it) { usleep(10000); q = cudaStreamQuery(e->stream); }
        fprintf(stderr, "[solo] f0=%lld frames=%d after wait: %s\n", (long long)f0, frames, cudaGetErrorString(q)); fflush(stderr);
        if (q == cudaErrorNotReady) _exit(3);
      }
    }
    e->stats.solo_jobs += jobs.size();
    e->stats.solo_voice_samples += (uint64_t)n_sitems * (uint64_t)kSoloSub;
    for (int c = 0; c < NC; ++c) e->stats.solo_class_items[c] += (uint64_t)counts[c];
    return 0;
  };
  hp.mark("welsh");
  if (e->n_fvoice) {
    if (e->finst_dirty) { int rc0 = upload_inst_tables(e); if (rc0) return rc0; }
    int rc = upload_events(flists, e->fev, e->fev_off, any_f, &e->fev_empty_on_device, e->stream);
    if (rc) return rc;
    {
      Launch l(e, true, 4);
      fm_kernel<kVoiceWarps><<<e->n_fwork, 32 * kVoiceWarps, tile_bytes, e->stream>>>(
          e->d_finst, e->d_fvoice, e->fwork.d, e->fitems.d, e->fev.d, e->fev_off.d, f0, frames);
    }
    e->stats.voice_samples += (uint64_t)e->n_fvoice * (uint64_t)frames;
  }
  hp.mark("fm");
  if (e->n_wvoice) {
    int rc = run_solo();
    if (rc) return rc;
  }
  hp.mark("solo");
  // samplers / drumkits: one launch per instrument over this chunk's plays (voice order, then time)
  {
    std::vector<SamplePlay> plays_h;
    std::vector<std::pair<Node*, std::pair<size_t, int>>> launches;
    auto add_play = [&](Node* n, const SampleVoiceHost& sv) {
      if (sv.sample < 0 || sv.n_end <= f0 || sv.n_on >= f0 + frames) return;
      const SampleDev& s = n->samples[(size_t)sv.sample];
      SamplePlay pl;
      pl.n_on = sv.n_on; pl.n_end = sv.n_end; pl.step_q = sv.step_q;
      pl.data = s.d; pl.len = s.n; pl.channels = s.channels; pl.pad = 0;
      plays_h.push_back(pl);
    };
    for (Node* n : e->plan) {
      if (n->kind != GB_INST_SAMPLER && n->kind != GB_INST_DRUMKIT) continue;
      const size_t k0 = plays_h.size();
      for (size_t v = 0; v < n->svoices.size(); ++v) {
        if (v < n->done_plays.size()) {
          for (const SampleVoiceHost& old : n->done_plays[v]) add_play(n, old);
          n->done_plays[v].clear();
        }
        add_play(n, n->svoices[v]);
      }
      launches.push_back({n, {k0, (int)(plays_h.size() - k0)}});
    }
    if (!plays_h.empty()) {
      if (!e->plays.reserve(plays_h.size())) return fail(e, GB_ENOMEM, "out of memory");
      memcpy(e->plays.h, plays_h.data(), plays_h.size() * sizeof(SamplePlay));
      CUDA_TRY(e, cudaMemcpyAsync(e->plays.d, e->plays.h, plays_h.size() * sizeof(SamplePlay), cudaMemcpyHostToDevice,
                                  e->stream));
      e->stats.h2d_bytes += plays_h.size() * sizeof(SamplePlay);
    }
    for (auto& L : launches) {
      Launch l(e, true);
      sampler_kernel<<<cdiv(frames, 256), 256, 0, e->stream>>>(e->plays.d + L.second.first, L.second.second,
                                                                L.first->buf, f0, frames);
      e->stats.voice_samples += (uint64_t)L.first->svoices.size() * (uint64_t)frames;
    }
  }
  // ---- 4. partial sums of instruments split over several CTAs: one launch for all of them ----
  // (instruments whose only consumer sums the partials itself skip this pass, unless one of them needs
  // its node buffer for a gain/pan automation pass in this chunk)
  e->fused_sums = e->fused_sums_enabled;
  for (Node* n : e->plan)
    if (n->fuse_partials && n->unit_gain) e->fused_sums = false;
  const int n_red = e->fused_sums ? e->n_partials_nf : e->n_partials;
  if (n_red) {
    Launch l(e, false);
    dim3 grid(cdiv(frames, 256), n_red);
    reduce_partials_kernel<<<grid, 256, 0, e->stream>>>(
        e->fused_sums ? e->partials_nf.d : (e->parity ? e->d_partials_alt : e->partials.d), frames);
  }
  // ---- 5. plan walk: toy sources, instrument DCA automation, effects ----
  // One launch per node, except runs of independent effects of one batch class on one level: those share
  // a launch (batch_class), and a memoryless effect fused behind an IIR stage has no launch of its own.
  auto seg_table = [&](Node* n, const SegParam** segs, int* nseg) {
    auto sr = seg_of.find(n);
    *segs = sr != seg_of.end() ? e->segs.d + sr->second.off : nullptr;
    *nseg = sr != seg_of.end() ? sr->second.n : 0;
  };
  auto batchable = [&](Node* n) {
    if (!e->batch_fx || batch_class(n) == 0 || n->fused_away) return false;
    int live = 0;
    for (uint32_t s : n->sources) {
      Node* sn = find(e, s);
      if (sn && sn->order >= 0 && sn->buf) ++live;
    }
    if (live > kMaxSources) return false;
    for (auto& l : e->links)
      if (l.dst == n->uid && l.d_table) return false;
    return true;
  };
  auto post_of = [&](Node* n) {
    PostOp po;
    po.segs = nullptr; po.nseg = 0; po.op = -1;
    if (n->post) {
      seg_table(n->post, &po.segs, &po.nseg);
      po.op = pointwise_op(n->post);
    }
    return po;
  };
  size_t fx_used = 0;
  for (size_t pi = 0; pi < e->plan.size(); ++pi) {
    Node* n = e->plan[pi];
    if (n->fused_away) continue;  // rendered by its source's launch
    if (batchable(n)) {
      size_t pj = pi + 1;
      while (pj < e->plan.size() && e->plan[pj]->level == n->level && batch_class(e->plan[pj]) == batch_class(n) &&
             (e->plan[pj]->fused_away || batchable(e->plan[pj])))
        ++pj;
      size_t cnt = 0;
      for (size_t q = pi; q < pj; ++q) cnt += e->plan[q]->fused_away ? 0 : 1;
      if (cnt >= 2) {
        if (!e->fxdescs.reserve(fx_used + cnt)) return fail(e, GB_ENOMEM, "out of memory");
        FxDesc* d0 = e->fxdescs.h + fx_used;
        size_t k = 0;
        for (size_t q = pi; q < pj; ++q) {
          Node* m = e->plan[q];
          if (m->fused_away) continue;
          FxDesc& d = d0[k++];
          memset(&d, 0, sizeof d);
          bool summed = false;
          int rc = gather_sources(e, m, frames, &d.src, &summed);
          if (rc) return rc;
          d.out = m->post ? m->post->buf : m->buf;
          seg_table(m, &d.segs, &d.nseg);
          d.state = m->d_lp ? (void*)m->d_lp : (void*)m->d_bq;
          d.post = post_of(m);
          d.op = batch_class(m) == 1 ? pointwise_op(m) : 0;
        }
        CUDA_TRY(e, cudaMemcpyAsync(e->fxdescs.d + fx_used, d0, cnt * sizeof(FxDesc), cudaMemcpyHostToDevice, e->stream));
        e->stats.h2d_bytes += cnt * sizeof(FxDesc);
        {
          Launch l(e, false);
          const FxDesc* dd = e->fxdescs.d + fx_used;
          if (batch_class(n) == 1) {
            dim3 grid(cdiv(frames, 256), (unsigned)cnt);
            pointwise_batch_kernel<<<grid, 256, 0, e->stream>>>(dd, frames);
          } else if (batch_class(n) == 2) {
            if (e->fx_minb == 1) lp24_batch_kernel<1><<<(int)cnt, 32 * kFxWarps, 0, e->stream>>>(dd, frames);
            else if (e->fx_minb == 2) lp24_batch_kernel<2><<<(int)cnt, 32 * kFxWarps, 0, e->stream>>>(dd, frames);
            else lp24_batch_kernel<3><<<(int)cnt, 32 * kFxWarps, 0, e->stream>>>(dd, frames);
          } else {
            if (e->fx_minb == 1) biquad_batch_kernel<1><<<(int)cnt, 32 * kFxWarps, 0, e->stream>>>(dd, frames);
            else if (e->fx_minb == 2) biquad_batch_kernel<2><<<(int)cnt, 32 * kFxWarps, 0, e->stream>>>(dd, frames);
            else biquad_batch_kernel<3><<<(int)cnt, 32 * kFxWarps, 0, e->stream>>>(dd, frames);
          }
        }
        fx_used += cnt;
        e->stats.fx_batched_nodes += cnt;
        pi = pj - 1;
        continue;
      }
    }
    const SegParam* segs;
    int nseg;
    seg_table(n, &segs, &nseg);
    if (n->is_inst) {
      if (n->kind == GB_INST_TOY_SOURCE) {
        Launch l(e, false);
        fill_kernel<<<cdiv(frames, 256), 256, 0, e->stream>>>(n->buf, frames, n->tp.level_left, n->tp.level_right);
      } else if (n->kind == GB_INST_OSCILLATOR) {
        const double top = 1.0 - 1.0 / 9007199254740992.0;
        OscSourceDesc d;
        d.waveform = n->osp.oscillator.waveform; d.pad = 0;
        d.dq = h_cycles_to_q(n->osp.oscillator.frequency / e->sr);
        d.duty_q = h_cycles_to_q(std::min(clamp01(n->osp.oscillator.pulse_width), top));
        d.seed = splitmix64((uint64_t)n->uid << 32);
        Launch l(e, true);
        oscillator_source_kernel<<<cdiv(frames, 256), 256, 0, e->stream>>>(d, n->buf, f0, frames);
        e->stats.voice_samples += (uint64_t)frames;
      } else if (n->kind == GB_INST_ENVELOPE) {
        Launch l(e, true);
        envelope_source_kernel<<<cdiv(frames, 256), 256, 0, e->stream>>>(make_shape(n->esp.envelope, e->sr), segs, nseg,
                                                                          n->buf, f0, frames);
        e->stats.voice_samples += (uint64_t)frames;
      } else if (n->unit_gain) {
        SourceList self;
        memset(&self, 0, sizeof self);
        self.n = 1;
        self.p[0] = n->buf;
        Launch l(e, false);
        pointwise_kernel<<<cdiv(frames, 256), 256, 0, e->stream>>>(self, n->buf, frames, OP_DCA, segs, nseg);
        n->unit_gain = false;  // the next chunk's instrument record carries the real gains again
        (n->kind == GB_INST_WELSH ? e->winst_dirty : e->finst_dirty) = true;
      }
      continue;
    }
    SourceList src;
    bool summed = false;
    int rc = gather_sources(e, n, frames, &src, &summed);
    if (rc) return rc;
    if (summed) continue;  // plain sum node: its buffer is already the table sum
    if (n->kind == GB_FX_DELAY) {
      {
        Launch l(e, false);
        delay_kernel<<<cdiv(frames, 256), 256, 0, e->stream>>>(src, n->hist[n->hist_cur], std::max(n->delay_frames, 1),
                                                                n->delay_frames, n->buf, frames);
      }
      if (n->delay_frames > 0) {
        Launch l(e, false);
        history_update_kernel<<<cdiv(n->delay_frames, 256), 256, 0, e->stream>>>(
            src, n->hist[n->hist_cur], n->hist[n->hist_cur ^ 1], n->delay_frames, frames);
        n->hist_cur ^= 1;
      }
      continue;
    }
    const SegParam* use_segs = segs;
    int use_nseg = nseg;
    for (auto& l : e->links) {
      if (l.dst != n->uid || !l.d_table) continue;
      Node* sn = find(e, l.src);
      const int first = (int)((GB_CONTROL_PERIOD - f0 % GB_CONTROL_PERIOD) % GB_CONTROL_PERIOD);
      const int nb = first < frames ? (frames - first + GB_CONTROL_PERIOD - 1) / GB_CONTROL_PERIOD : 0;
      const int ns = nb + (first == 0 ? 0 : 1);
      SegParam base;
      base.t0 = 0; base.pad = 0;
      seg_values(e, n, base.v);
      Launch lk(e, false);
      sidechain_table_kernel<<<cdiv(std::max(ns, 1), 128), 128, 0, e->stream>>>(
          sn->buf, frames, first, GB_CONTROL_PERIOD, std::max(ns, 1), base, l.index, l.d_table, l.d_state + 4 * l.cur,
          l.d_state + 4 * (l.cur ^ 1), (long long)f0);
      l.cur ^= 1;
      use_segs = l.d_table;
      use_nseg = std::max(ns, 1);
    }
    const PostOp po = post_of(n);
    rc = run_effect(e, n, src, frames, use_segs, use_nseg, f0, n->post ? &po : nullptr);
    if (rc) return rc;
    if (n->kind == GB_FX_CHORUS && n->delay_frames > 0) {
      Launch l(e, false);
      history_update_kernel<<<cdiv(n->delay_frames, 256), 256, 0, e->stream>>>(
          src, n->hist[n->hist_cur], n->hist[n->hist_cur ^ 1], n->delay_frames, frames);
      n->hist_cur ^= 1;
    }
  }
  CUDA_TRY(e, cudaGetLastError());
  Node* root = find(e, GB_MAIN_MIXER);
  e->last_out = root->buf;
  e->last_frames = (size_t)frames;
  CUDA_TRY(e, cudaEventRecord(e->stage_done[slot], e->stream));
  if (e->overlap) CUDA_TRY(e, cudaEventRecord(e->p_done[par], e->stream));
  hp.mark("fx");
  e->chunk_seq++;
  e->pos += frames;
  return 0;
}

enum OutMode { OUT_F64, OUT_PCM16, OUT_DEVICE };

// Where to end the chunk that starts at f0 (at most `limit` frames).  The resting and the sweeping
// kernels take a CTA only if its voices keep their state for the WHOLE chunk, so chunks are cut at the
// host-known transitions of the Welsh voices — note events and envelope stage boundaries (integer
// frames).  Transitions come in clusters (the voices of a chord, a beat); a chunk either ends just
// before a cluster (a pure stretch, left to the specialised kernels) or covers one cluster.  All cuts
// are multiples of kBlockFrames from f0; stretches shorter than kMinZone frames are not split off.
int64_t chunk_cut(gb_engine* e, int64_t f0, int64_t limit) {
  constexpr int64_t kMinZone = 16 * kBlockFrames;
  // cuts pay off when the Welsh voices dominate a chunk; with a handful of voices the extra launches cost
  // more than the specialised kernels save (a 16-voice song went from 28 to 210 launches)
  if (limit < 2 * kMinZone || e->n_wwork_grouped == 0 || !e->chunk_cuts || e->n_wvoice < e->opt.min_cut_voices) return limit;
  std::vector<int64_t> tr;
  const int64_t f1 = f0 + limit;
  for (const gb_event& ev : e->events) {  // sorted by frame
    if (ev.frame >= f1) break;
    if (ev.frame <= f0 || (ev.type != GB_EV_NOTE_ON && ev.type != GB_EV_NOTE_OFF)) continue;
    const Node* n = find(e, ev.uid);
    if (n && n->kind == GB_INST_WELSH && n->order >= 0) tr.push_back(ev.frame);
  }
  for (const Node* n : e->plan) {
    if (n->kind != GB_INST_WELSH || n->nvoices < kVoiceWarps) continue;
    const WelshInst& I = e->h_winst[(size_t)n->table_index];
    if (I.rest_class < 0 && I.sweep_class < 0) continue;
    const int64_t b[4] = {I.amp.na, I.amp.na + I.amp.nd, I.filt.na, I.filt.na + I.filt.nd};
    for (const Slot& sl : n->store.slots) {
      if (!sl.held) {  // releasing: the voice falls idle at idle_at
        if (sl.idle_at > f0 && sl.idle_at < f1) tr.push_back(sl.idle_at);
        continue;
      }
      if (sl.on_frame <= kNever) continue;
      for (int64_t d : b) {
        const int64_t t = sl.on_frame + d;
        if (t > f0 && t < f1) tr.push_back(t);
      }
    }
  }
  if (tr.empty()) return limit;
  std::sort(tr.begin(), tr.end());
  auto up = [&](int64_t t) { return (t - f0 + kBlockFrames - 1) / kBlockFrames * kBlockFrames; };
  const int64_t first = (tr.front() - f0) / kBlockFrames * kBlockFrames;  // whole blocks before the first transition
  if (first >= kMinZone) return first;
  // the chunk starts inside (or just before) a cluster: run to its end
  int64_t last = tr.front();
  for (int64_t t : tr) {
    if (t - last >= kMinZone) break;
    last = t;
  }
  const int64_t cut = up(last + 1);
  return cut >= limit || limit - cut < kMinZone ? limit : cut;
}

// Split the request into chunks of at most max_block frames.
int render_impl(gb_engine* e, void* out, size_t frames, size_t* done, OutMode mode) {
  if (!e) return GB_EINVAL;
  if (!e->finalized) return fail(e, GB_ESTATE, "engine is not finalized");
  if (frames && !out && mode != OUT_DEVICE) return fail(e, GB_EINVAL, "null output buffer");
  cudaSetDevice(e->device);
  // every mode gathers the chunks in d_full (device); host-buffer renders stream it back from there
  if (frames > e->full_cap) {
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    CUDA_TRY(e, cudaStreamSynchronize(e->copy_stream));
    if (e->d_full) cudaFree(e->d_full);
    e->d_full = nullptr;
    e->full_cap = 0;
    CUDA_TRY(e, cudaMalloc(&e->d_full, std::max<size_t>(frames, 1) * sizeof(double2)));
    e->full_cap = frames;
  }
  if (mode == OUT_F64 && !e->ring)  // normally allocated by gb_finalize
    CUDA_TRY(e, cudaMallocHost(&e->ring, (size_t)kStageSlots * e->max_block * sizeof(double2)));
  size_t produced = 0;
  const bool call_timed = span_begin(e, 2);
  if (e->overlap) CUDA_TRY(e, cudaEventRecord(e->call_start, e->stream));  // vstream starts after whatever precedes this call
  const size_t call_span = e->spans.size() - 1;
  // Chunks are enqueued without waiting for each other.  For host-buffer renders chunk i's result goes
  // d_full -> pinned ring slot on the copy stream while chunks i+1.. compute, and the host moves a ring
  // slot into the caller's (pageable) buffer just before the slot is reused.
  struct Pending { size_t off, n; int slot; };
  std::vector<Pending> pending;
  uint64_t out_seq = 0;
  auto drain_one = [&]() -> int {
    const Pending p = pending.front();
    pending.erase(pending.begin());
    CUDA_TRY(e, cudaEventSynchronize(e->d2h_done[p.slot]));
    memcpy((double*)out + 2 * p.off, e->ring + (size_t)p.slot * e->max_block, p.n * sizeof(double2));
    return 0;
  };
  while (produced < frames) {
    const int64_t f0 = e->pos;
    int64_t limit = (int64_t)std::min<size_t>(frames - produced, e->max_block);
    // long ragged chunks are cut at a block multiple (the remainder becomes its own small chunk), so
    // that resting voices can take welsh_rest_kernel, which renders whole blocks only
    if (limit >= 16 * kBlockFrames && limit % kBlockFrames) limit -= limit % kBlockFrames;
    limit = chunk_cut(e, f0, limit);
    // every event strictly before the chunk end belongs to this chunk (events never split a chunk:
    // parameters go through segment tables, sampler retriggers through per-chunk play lists)
    size_t n_ev = 0;
    while (n_ev < e->events.size() && e->events[n_ev].frame < f0 + limit) ++n_ev;
    int rc = render_chunk(e, (int)limit, n_ev);
    if (rc) return rc;
    const size_t n = (size_t)limit;
    CUDA_TRY(e, cudaMemcpyAsync(e->d_full + produced, e->last_out, n * sizeof(double2), cudaMemcpyDeviceToDevice,
                                e->stream));
    if (mode == OUT_F64) {
      const int rslot = (int)(out_seq++ % kStageSlots);
      if (pending.size() == (size_t)kStageSlots && (rc = drain_one())) return rc;
      CUDA_TRY(e, cudaEventRecord(e->d2d_done[rslot], e->stream));
      CUDA_TRY(e, cudaStreamWaitEvent(e->copy_stream, e->d2d_done[rslot], 0));
      CUDA_TRY(e, cudaMemcpyAsync(e->ring + (size_t)rslot * e->max_block, e->d_full + produced, n * sizeof(double2),
                                  cudaMemcpyDeviceToHost, e->copy_stream));
      CUDA_TRY(e, cudaEventRecord(e->d2h_done[rslot], e->copy_stream));
      pending.push_back({produced, n, rslot});
      e->stats.d2h_bytes += n * sizeof(double2);
    }
    produced += n;
  }
  while (!pending.empty()) {
    int rc = drain_one();
    if (rc) return rc;
  }
  if (mode != OUT_PCM16) CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  if (mode == OUT_PCM16 && frames) {
    if (frames > e->pcm_cap) {
      if (e->d_pcm) cudaFree(e->d_pcm);
      e->d_pcm = nullptr;
      e->pcm_cap = 0;
      CUDA_TRY(e, cudaMalloc(&e->d_pcm, frames * sizeof(short2)));
      e->pcm_cap = frames;
    }
    {
      Launch l(e, false);
      pcm16_kernel<<<cdiv((int)frames, 256), 256, 0, e->stream>>>(e->d_full, e->d_pcm, (int)frames);
    }
    CUDA_TRY(e, cudaMemcpyAsync(out, e->d_pcm, frames * sizeof(short2), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    e->stats.d2h_bytes += frames * sizeof(short2);
  }
  if (call_timed) span_end(e, call_span);
  if (e->solo_ran) {  // the job-list kernel's dependency watchdog (the stream is drained at this point)
    int fault[5] = {0, 0, 0, 0, 0};
    CUDA_TRY(e, cudaMemcpy(fault, e->d_solo_fault, sizeof fault, cudaMemcpyDeviceToHost));
    e->solo_ran = false;
    if (fault[0]) {
      CUDA_TRY(e, cudaMemset(e->d_solo_fault, 0, 8 * sizeof(int)));
      return fail(e, GB_ECUDA, "welsh_solo_kernel: job %d waited for voice %d to reach %d (stuck at %d)", fault[1], fault[2],
                  fault[3], fault[4]);
    }
  }
  e->full_frames = mode != OUT_F64 ? frames : 0;
  if (done) *done = produced;
  return 0;
}

}  // namespace

extern "C" {

int gb_render_block(gb_engine* e, double* out, size_t frames, size_t* done) {
  if (!e || !e->lookahead) return render_impl(e, out, frames, done, OUT_F64);
  if (frames && !out) return fail(e, GB_EINVAL, "null output buffer");
  // look-ahead mode: small caller buffers (the reference calls with 64 frames, orchestrator.rs:1696) are served
  // from a host buffer that one big render fills every `lookahead` frames
  size_t got = 0;
  while (got < frames) {
    if (e->ahead_off == e->ahead_n) {
      if (frames - got >= e->lookahead) {  // a big request goes straight through
        size_t d = 0;
        int rc = render_impl(e, out + 2 * got, frames - got, &d, OUT_F64);
        if (rc) return rc;
        got += d;
        break;
      }
      e->ahead_n = e->ahead_off = 0;
      size_t d = 0;
      int rc = render_impl(e, e->ahead.data(), e->lookahead, &d, OUT_F64);
      if (rc) return rc;
      e->ahead_n = d;
    }
    const size_t k = std::min(frames - got, e->ahead_n - e->ahead_off);
    memcpy(out + 2 * got, e->ahead.data() + 2 * e->ahead_off, k * 2 * sizeof(double));
    e->ahead_off += k;
    got += k;
  }
  if (done) *done = got;
  return 0;
}
int gb_set_lookahead(gb_engine* e, size_t frames) {
  if (!e) return GB_EINVAL;
  if (e->ahead_off != e->ahead_n) return fail(e, GB_ESTATE, "frames rendered ahead are still pending");
  e->lookahead = frames;
  e->ahead.assign(2 * frames, 0.0);
  e->ahead_n = e->ahead_off = 0;
  return 0;
}
int gb_render_pcm16(gb_engine* e, int16_t* out, size_t frames, size_t* done) {
  return render_impl(e, out, frames, done, OUT_PCM16);
}
int gb_render_device(gb_engine* e, size_t frames, size_t* done) {
  return render_impl(e, nullptr, frames, done, OUT_DEVICE);
}
int gb_read_last(gb_engine* e, double* out, size_t frames) {
  if (!e || !out) return GB_EINVAL;
  if (!e->d_full || frames > e->full_frames) return fail(e, GB_ESTATE, "no device-resident render of that size");
  cudaSetDevice(e->device);
  CUDA_TRY(e, cudaMemcpy(out, e->d_full, frames * sizeof(double2), cudaMemcpyDeviceToHost));
  e->stats.d2h_bytes += frames * sizeof(double2);
  return 0;
}
int gb_last_device_buffer(gb_engine* e, void** device_ptr, size_t* frames) {
  if (!e || !device_ptr || !frames) return GB_EINVAL;
  if (!e->d_full || e->full_frames == 0) return fail(e, GB_ESTATE, "no device-resident render");
  *device_ptr = e->d_full;
  *frames = e->full_frames;
  return 0;
}
int64_t gb_position(const gb_engine* e) { return e ? e->pos - (int64_t)(e->ahead_n - e->ahead_off) : -1; }

// ---- state save / restore -------------------------------------------------------------------
// Layout: header {magic, pos, n_wvoice, n_fvoice, n_nodes} then device voice tables, then per plan
// node its host allocation state and device effect state.  Only valid for an engine built by the
// same sequence of add/patch/finalize calls.
}  // extern "C"
namespace {
struct Blob {
  std::vector<uint8_t> data;
  template <typename T>
  void put(const T& v) {
    const uint8_t* p = (const uint8_t*)&v;
    data.insert(data.end(), p, p + sizeof(T));
  }
  void put_bytes(const void* p, size_t n) { data.insert(data.end(), (const uint8_t*)p, (const uint8_t*)p + n); }
};
struct Reader {
  const uint8_t* p;
  size_t left;
  bool ok = true;
  template <typename T>
  void get(T* v) { get_bytes(v, sizeof(T)); }
  void get_bytes(void* dst, size_t n) {
    if (n > left) { ok = false; return; }
    memcpy(dst, p, n);
    p += n; left -= n;
  }
};
int dev_to_blob(gb_engine* e, Blob& b, const void* d, size_t bytes) {
  std::vector<uint8_t> tmp(bytes);
  if (bytes) CUDA_TRY(e, cudaMemcpy(tmp.data(), d, bytes, cudaMemcpyDeviceToHost));
  b.put_bytes(tmp.data(), bytes);
  return 0;
}
int blob_to_dev(gb_engine* e, Reader& r, void* d, size_t bytes) {
  std::vector<uint8_t> tmp(bytes);
  r.get_bytes(tmp.data(), bytes);
  if (!r.ok) return fail(e, GB_EINVAL, "state blob truncated");
  if (bytes) CUDA_TRY(e, cudaMemcpy(d, tmp.data(), bytes, cudaMemcpyHostToDevice));
  return 0;
}
template <typename F>
int for_each_state_region(gb_engine* e, F&& fn) {
  int rc;
  if (e->n_wvoice && (rc = fn(e->d_wvoice, (size_t)e->n_wvoice * sizeof(WelshVoice)))) return rc;
  if (e->n_fvoice && (rc = fn(e->d_fvoice, (size_t)e->n_fvoice * sizeof(FmVoice)))) return rc;
  for (Node* n : e->plan) {
    if (n->d_bq && (rc = fn(n->d_bq, sizeof(BiquadState)))) return rc;
    if (n->d_lp && (rc = fn(n->d_lp, sizeof(Lp24State)))) return rc;
    if (n->hist[0]) {
      size_t len = (size_t)std::max(n->delay_frames, 1) * sizeof(double2);
      if ((rc = fn(n->hist[n->hist_cur], len))) return rc;
    }
    if (n->kind == GB_FX_REVERB) {
      for (int ch = 0; ch < 2; ++ch) {
        for (int i = 0; i < 4; ++i)
          if ((rc = fn(n->rv.comb_ring[ch][i], (size_t)n->rv.comb_d[i] * sizeof(double)))) return rc;
        for (int i = 0; i < 2; ++i)
          if ((rc = fn(n->rv.ap_ring[ch][i], (size_t)n->rv.ap_d[i] * sizeof(double)))) return rc;
      }
    }
  }
  for (auto& l : e->links)  // the current half of the ping-pong link state
    if (l.d_state && (rc = fn(l.d_state + 4 * l.cur, 4 * sizeof(double)))) return rc;
  return 0;
}
}  // namespace
extern "C" {

int gb_save_state(gb_engine* e, void* buf, size_t* size) {
  if (!e || !size) return GB_EINVAL;
  if (!e->finalized) return fail(e, GB_ESTATE, "engine is not finalized");
  if (e->ahead_off != e->ahead_n) return fail(e, GB_ESTATE, "look-ahead mode: frames rendered ahead have not been handed out yet");
  cudaSetDevice(e->device);
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  Blob b;
  uint64_t magic = 0x32305453424700ull;  // "GBST02"
  b.put(magic);
  b.put(e->pos);
  b.put(e->n_wvoice);
  b.put(e->n_fvoice);
  uint32_t nplan = (uint32_t)e->plan.size();
  b.put(nplan);
  for (Node* n : e->plan) {
    b.put(n->uid);
    b.put(n->p);
    b.put(n->wp.dca);
    b.put(n->fp.dca);
    uint32_t ns = (uint32_t)n->store.slots.size();
    b.put(ns);
    for (auto& s : n->store.slots) b.put(s);
    uint32_t nsv = (uint32_t)n->svoices.size();
    b.put(nsv);
    for (auto& s : n->svoices) b.put(s);
    b.put(n->envh);
  }
  uint64_t nev = e->events.size();
  b.put(nev);
  for (auto& ev : e->events) b.put(ev);
  int rc = for_each_state_region(e, [&](void* d, size_t bytes) { return dev_to_blob(e, b, d, bytes); });
  if (rc) return rc;
  if (!buf) {
    *size = b.data.size();
    return 0;
  }
  if (*size < b.data.size()) return fail(e, GB_EINVAL, "state buffer too small");
  memcpy(buf, b.data.data(), b.data.size());
  *size = b.data.size();
  return 0;
}

int gb_restore_state(gb_engine* e, const void* buf, size_t size) {
  if (!e || !buf) return GB_EINVAL;
  if (!e->finalized) return fail(e, GB_ESTATE, "engine is not finalized");
  cudaSetDevice(e->device);
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  Reader r{(const uint8_t*)buf, size};
  uint64_t magic = 0;
  int nw = 0, nf = 0;
  uint32_t nplan = 0;
  int64_t pos = 0;
  r.get(&magic); r.get(&pos); r.get(&nw); r.get(&nf); r.get(&nplan);
  if (!r.ok || magic != 0x32305453424700ull || nw != e->n_wvoice || nf != e->n_fvoice || nplan != e->plan.size())
    return fail(e, GB_EINVAL, "state blob does not match this engine");
  for (Node* n : e->plan) {
    uint32_t uid = 0, ns = 0, nsv = 0;
    r.get(&uid);
    if (!r.ok || uid != n->uid) return fail(e, GB_EINVAL, "state blob does not match this engine's plan");
    r.get(&n->p);
    r.get(&n->wp.dca);
    r.get(&n->fp.dca);
    r.get(&ns);
    if (ns != n->store.slots.size()) return fail(e, GB_EINVAL, "state blob voice-store mismatch");
    for (auto& s : n->store.slots) r.get(&s);
    r.get(&nsv);
    if (nsv != n->svoices.size()) return fail(e, GB_EINVAL, "state blob sample-voice mismatch");
    for (auto& s : n->svoices) r.get(&s);
    r.get(&n->envh);
  }
  uint64_t nev = 0;
  r.get(&nev);
  if (!r.ok || nev > r.left / sizeof(gb_event)) return fail(e, GB_EINVAL, "state blob truncated");
  e->events.resize((size_t)nev);
  for (auto& ev : e->events) r.get(&ev);
  int rc = for_each_state_region(e, [&](void* d, size_t bytes) { return blob_to_dev(e, r, d, bytes); });
  if (rc) return rc;
  e->pos = pos;
  e->ahead_n = e->ahead_off = 0;  // frames rendered ahead of the restored position are dropped
  e->winst_dirty = e->finst_dirty = true;
  return 0;
}

// ---- measurement ------------------------------------------------------------------------------
int gb_get_stats(gb_engine* e, gb_stats* out) {
  if (!e || !out) return GB_EINVAL;
  cudaSetDevice(e->device);
  resolve_spans(e);
  *out = e->stats;
  return 0;
}
int gb_reset_stats(gb_engine* e) {
  if (!e) return GB_EINVAL;
  cudaSetDevice(e->device);
  resolve_spans(e);
  memset(&e->stats, 0, sizeof e->stats);
  return 0;
}
int gb_set_timing(gb_engine* e, int32_t enabled) {
  if (!e) return GB_EINVAL;
  e->timing = enabled != 0;
  return 0;
}
int gb_measure_fma_peak(gb_engine* e, int32_t fp64, double* tflops) {
  if (!e || !tflops) return GB_EINVAL;
  cudaSetDevice(e->device);
  const int blocks = e->num_sms * 16, threads = 256, iters = 1 << 14;
  void* d = nullptr;
  CUDA_TRY(e, cudaMalloc(&d, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a, e->stream);
    if (fp64) fma_peak_kernel<double><<<blocks, threads, 0, e->stream>>>((double*)d, iters);
    else fma_peak_kernel<float><<<blocks, threads, 0, e->stream>>>((float*)d, iters);
    cudaEventRecord(b, e->stream);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d);
  CUDA_TRY(e, cudaGetLastError());
  double flops = (double)blocks * threads * (double)iters * 8.0 * 2.0;
  *tflops = flops / ((double)best * 1e-3) / 1e12;
  return 0;
}

// ---- multi-GPU bus exchange --------------------------------------------------------------------------
// One process per GPU.  Every rank owns an exchange buffer (its rendered stereo bus, f64 L/R interleaved) whose
// CUDA IPC handle the host side passes around (16-byte-per-frame buffers, one per rank); the root maps the
// peers' buffers and sums them with ONE kernel whose loads cross NVLink (peer_sum_kernel) — the transfer and
// the mix are the same pass, instead of an NCCL reduce that first moves and then adds.  All work is enqueued
// on the caller's stream (the host side brackets it with its own events / barrier).
struct gb_bus_exchange {
  int device = 0;
  size_t frames = 0;
  double2* mine = nullptr;
  double2* sum = nullptr;
  int n = 0, self = 0;
  PeerTable tab;
  std::vector<void*> opened;
};

int gb_bus_exchange_create(int32_t device, size_t frames, gb_bus_exchange** out) {
  if (!out || !frames) return fail(nullptr, GB_EINVAL, "gb_bus_exchange_create: bad arguments");
  if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, GB_ENODEV, "no CUDA device %d (groove_b200 has no CPU fallback)", device);
  auto x = std::make_unique<gb_bus_exchange>();
  x->device = device;
  x->frames = frames;
  memset(&x->tab, 0, sizeof x->tab);
  if (cudaMalloc(&x->mine, frames * sizeof(double2)) != cudaSuccess || cudaMalloc(&x->sum, frames * sizeof(double2)) != cudaSuccess) {
    if (x->mine) cudaFree(x->mine);
    return fail(nullptr, GB_ENOMEM, "gb_bus_exchange_create: out of device memory");
  }
  cudaMemset(x->mine, 0, frames * sizeof(double2));
  cudaMemset(x->sum, 0, frames * sizeof(double2));
  cudaDeviceSynchronize();
  *out = x.release();
  return 0;
}
void gb_bus_exchange_destroy(gb_bus_exchange* x) {
  if (!x) return;
  cudaSetDevice(x->device);
  for (void* p : x->opened) cudaIpcCloseMemHandle(p);
  if (x->mine) cudaFree(x->mine);
  if (x->sum) cudaFree(x->sum);
  delete x;
}
// handle: GB_IPC_HANDLE_BYTES bytes (a cudaIpcMemHandle_t) naming this rank's exchange buffer
int gb_bus_exchange_export(gb_bus_exchange* x, void* handle) {
  if (!x || !handle) return GB_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == GB_IPC_HANDLE_BYTES, "handle size");
  cudaIpcMemHandle_t h;
  cudaSetDevice(x->device);
  cudaError_t rc = cudaIpcGetMemHandle(&h, x->mine);
  if (rc != cudaSuccess) return fail(nullptr, GB_ECUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(rc));
  memcpy(handle, &h, sizeof h);
  return 0;
}
// handles: n x GB_IPC_HANDLE_BYTES in rank order (entry `self` is ignored: the local buffer is used)
int gb_bus_exchange_open(gb_bus_exchange* x, const void* handles, int32_t n, int32_t self) {
  if (!x || !handles || n < 1 || n > kMaxPeers || self < 0 || self >= n) return fail(nullptr, GB_EINVAL, "gb_bus_exchange_open: bad arguments");
  cudaSetDevice(x->device);
  for (void* p : x->opened) cudaIpcCloseMemHandle(p);
  x->opened.clear();
  memset(&x->tab, 0, sizeof x->tab);
  for (int r = 0; r < n; ++r) {
    if (r == self) { x->tab.p[r] = x->mine; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * GB_IPC_HANDLE_BYTES, sizeof h);
    void* p = nullptr;
    cudaError_t rc = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (rc != cudaSuccess) {
      cudaGetLastError();
      return fail(nullptr, GB_ECUDA, "cudaIpcOpenMemHandle (rank %d) failed: %s", r, cudaGetErrorString(rc));
    }
    x->opened.push_back(p);
    x->tab.p[r] = (const double2*)p;
  }
  x->n = n;
  x->self = self;
  return 0;
}
// copy the engine's last device-resident render (gb_render_device) into this rank's exchange buffer
int gb_bus_exchange_publish(gb_bus_exchange* x, gb_engine* e, size_t frames, void* stream) {
  if (!x || !e) return GB_EINVAL;
  if (frames > x->frames || frames > e->full_frames) return fail(e, GB_EINVAL, "gb_bus_exchange_publish: no device-resident render of %zu frames", frames);
  cudaSetDevice(x->device);
  CUDA_TRY(e, cudaMemcpyAsync(x->mine, e->d_full, frames * sizeof(double2), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}
// root: sum the ranks' buffers (remote ones read over NVLink) into the local result buffer
int gb_bus_exchange_reduce(gb_bus_exchange* x, size_t frames, void* stream) {
  if (!x || x->n < 1 || frames > x->frames) return fail(nullptr, GB_EINVAL, "gb_bus_exchange_reduce: exchange not opened");
  cudaSetDevice(x->device);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, x->device);
  const int grid = (int)std::min<size_t>((frames + 255) / 256, (size_t)sms * 8);
  peer_sum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x->tab, x->n, x->sum, frames);
  cudaError_t rc = cudaGetLastError();
  if (rc != cudaSuccess) return fail(nullptr, GB_ECUDA, "peer_sum_kernel launch failed: %s", cudaGetErrorString(rc));
  return 0;
}
int gb_bus_exchange_result(gb_bus_exchange* x, void** device_ptr, size_t* frames) {
  if (!x || !device_ptr) return GB_EINVAL;
  *device_ptr = x->sum;
  if (frames) *frames = x->frames;
  return 0;
}

}  // extern "C"
