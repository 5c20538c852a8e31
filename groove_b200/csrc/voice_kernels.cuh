// voice_kernels.cuh — instrument voices: Welsh subtractive synth, FM operator pair, sampler/drumkit.
//
// Work decomposition (replaces the reference's per-frame `tick(1)` + `value()` walk,
// orchestration/src/orchestrator.rs:401-410):
//   * one WARP per voice, marching over the render chunk in blocks of 32 lanes x T frames;
//     lane l owns T consecutive frames, so lane order == time order.
//   * everything that has a closed form per frame (oscillator phase, LFO, both ADSRs, the
//     per-frame 24 dB coefficient set) is evaluated independently per frame;
//   * the filter feedback is a time-varying 2x2 linear recurrence per section.  Each lane runs
//     its T frames from a zero state while tracking the homogeneous response, the 32 lane
//     aggregates are combined with a warp-shuffle scan of affine maps, and each lane then fixes
//     its T outputs up with 2 FMAs per frame;
//   * frequency-modulated phases (pitch LFO, FM carrier) are 64-bit integer prefix sums, scanned
//     the same way (exact, so the result does not depend on the association order);
//   * the W voices of a CTA are summed through shared memory, so per-voice audio never goes to
//     HBM: one 16-byte store per frame per CTA.
#pragma once

#include "dsp.cuh"

namespace gbk {

constexpr int kT = 8;                 // frames per lane
constexpr int kBlockFrames = 32 * kT; // frames per warp step
constexpr int kTileStride = kBlockFrames + kBlockFrames / 8;  // padded (16-byte units)

enum { VEV_NOTE_ON = 1, VEV_NOTE_OFF = 2 };

struct VoiceEvent {
  i64 frame;
  int type;
  int pad;
  double cyc1, cyc2;  // base cycles per frame of oscillator 1 / 2 (FM: carrier / unused)
};

// ------------------------------------------------------------------------ Welsh ---
// Every piecewise-linear waveform is  value = a*p + b  with (a,b) picked by one phase compare:
//   square  q<half ? +1 : -1      pulse  q<duty ? +1 : -1
//   triangle q<half ? 4p-1 : 3-4p  saw    q<half ? 2p : 2p-2      constants: a = 0
// which evaluates without a branch.  kind: 0 = affine, 1 = sine, 2 = noise.
struct OscShape {
  int kind, pad;
  u64 thresh;
  double a_lo, b_lo, a_hi, b_hi;
};
__device__ __forceinline__ double osc_affine(const OscShape& o, u64 q, u64 thresh) {
  const bool lo = q < thresh;
  return fma(lo ? o.a_lo : o.a_hi, pos_of(q), lo ? o.b_lo : o.b_hi);
}
// sine and noise oscillators: one shared out-of-line copy (the library sinpi is ~100 instructions, and the
// fast block evaluates up to three oscillators at each of its kT unrolled frames)
__device__ __noinline__ double osc_other_ool(int kind, u64 q, u64 seed, i64 frame) {
  if (kind == 1) return sinpi(2.0 * pos_of(q));
  return __ull2double_rn(splitmix64(seed + (u64)frame) >> 11) * (1.0 / 4503599627370496.0) - 1.0;
}
__device__ __forceinline__ double osc_eval(const OscShape& o, u64 q, u64 thresh, u64 seed, i64 frame) {
  if (o.kind == 0) return osc_affine(o, q, thresh);
  return osc_other_ool(o.kind, q, seed, frame);
}

// ---- range-specialised elementary functions for the cutoff -> coefficient map ----------------------
// The coefficient map only ever needs 2^y for y in [-20, 0] and sin/cos(pi u) for u in (0, 0.49].
// On those ranges plain Taylor/Horner forms reach 1e-16 with no special cases, and with the
// coefficients in constant memory every FMA takes its constant as a constant-bank operand (the
// library versions re-materialise each 64-bit constant with two UMOVs when registers are short).
__constant__ double kExp2C[13] = {  // (ln 2)^k / k!, k = 13 .. 1
    1.3691488853904128e-12, 2.5678435993488206e-11, 4.4455382718708116e-10, 7.054911620801123e-09,
    1.01780860092397e-07, 1.321548679014431e-06, 1.5252733804059841e-05, 0.0001540353039338161,
    0.0013333558146428443, 0.009618129107628477, 0.05550410866482158, 0.24022650695910072, 0.6931471805599453};
__constant__ double kSinC[9] = {  // (-1)^k / (2k+1)!, k = 8 .. 0
    2.8114572543455206e-15, -7.6471637318198164e-13, 1.6059043836821613e-10, -2.5052108385441720e-08,
    2.7557319223985893e-06, -1.9841269841269841e-04, 8.3333333333333332e-03, -1.6666666666666666e-01, 1.0};
__constant__ double kCosC[9] = {  // (-1)^k / (2k)!, k = 8 .. 0
    4.7794773323873853e-14, -1.1470745597729725e-11, 2.0876756987868100e-09, -2.7557319223985888e-07,
    2.4801587301587302e-05, -1.3888888888888889e-03, 4.1666666666666664e-02, -0.5, 1.0};

// 2^y, |y| < 1000 (no overflow/underflow handling needed by the caller's range)
__device__ __forceinline__ double exp2_ranged(double y) {
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: rounds y to an integer in the low word
  const double t = y + magic;
  const int n = __double2loint(t);
  const double f = y - (t - magic);  // [-0.5, 0.5]
  double p = kExp2C[0];
#pragma unroll
  for (int k = 1; k < 13; ++k) p = fma(p, f, kExp2C[k]);
  p = fma(p, f, 1.0);
  return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}
// sin(pi u), cos(pi u) for u in [0, 0.5]
__device__ __forceinline__ void sincospi_ranged(double u, double* s, double* c) {
  const bool hi = u > 0.25;
  const double t = hi ? 0.5 - u : u;  // [0, 0.25]
  const double x = t * kPi;
  const double z = x * x;
  double ps = kSinC[0], pc = kCosC[0];
#pragma unroll
  for (int k = 1; k < 9; ++k) {
    ps = fma(ps, z, kSinC[k]);
    pc = fma(pc, z, kCosC[k]);
  }
  ps *= x;
  *s = hi ? pc : ps;
  *c = hi ? ps : pc;
}
// 1/x for normal positive x: single-precision seed + two Newton steps (relative error 1e-7 -> 1e-14 -> < 1 ulp)
__device__ __forceinline__ double rcp_ranged(double x) {
  double r = (double)__frcp_rn((float)x);
  r = fma(fma(-x, r, 1.0), r, r);
  r = fma(fma(-x, r, 1.0), r, r);
  return r;
}

// Both sections' coefficients from u = fc/sr with ONE division: with s = sin(pi u), c = cos(pi u)
// and k = s/c, multiplying numerator and denominator of the k-form (dsp.cuh lp24_from_k) by c^2 gives
//   b0 = s^2/D, a1 = 2(c0 c^2 - s^2)/D, a2 = (c1k s c - s^2 - c0 c^2)/D, D = c1k s c + s^2 + c0 c^2.
__device__ __forceinline__ void lp24_from_u(const Lp24Ripple& rp, double u, SecCoef& s1, SecCoef& s2) {
  double s, c;
  sincospi_ranged(u, &s, &c);
  const double ss = s * s, cc = c * c, sc = s * c;
  const double n1 = rp.c0 * cc, n2 = rp.c2 * cc;
  const double d1 = fma(rp.c1k, sc, ss + n1), d2 = fma(rp.c3k, sc, ss + n2);
  const double r = rcp_ranged(d1 * d2);
  const double i1 = d2 * r, i2 = d1 * r;
  s1.b0 = ss * i1;
  s1.a1 = 2.0 * (n1 - ss) * i1;
  s1.a2 = (fma(rp.c1k, sc, -ss) - n1) * i1;
  s2.b0 = ss * i2;
  s2.a1 = 2.0 * (n2 - ss) * i2;
  s2.a2 = (fma(rp.c3k, sc, -ss) - n2) * i2;
}

// Both oscillators and their mix folded into one expression for the specialised blocks: with
// P = 1 + pos(q) taken straight from the mantissa trick (no subtraction) and (a, b) the affine piece
// selected by the phase compare,  mix1*(a1*pos1 + b1) + mix2*(a2*pos2 + b2) = A1*P1 + A2*P2 + (B1 + B2)
// where A = mix*a and B = mix*(b - a): 3 FP64 operations per frame instead of 6.
struct OscMix {
  double a_lo, b_lo, a_hi, b_hi;
};
// Time-invariant stretches of the 24 dB filter (fixed cutoff, or the filter envelope resting at its
// sustain level): the coefficient sets, the rows g[j] = (A^j)[0][*] that carry a lane's entry state to
// its j-th output, and the span maps A^(kT * 2^k) for the scan are per-instrument constants, computed
// once on the host (engine.cu, welsh_inst_from_params).
struct alignas(16) LtiTable {
  double g1[8][2], g2[8][2];    // kT rows per section
  double g1b[8][2];             // g1 scaled by section 2's b0 (welsh_rest_block feeds section 2 with b0*x directly)
  double inv_b0_2, pad_lti;     // 1 / c2.b0
  double mp1[6][4], mp2[6][4];  // row-major 2x2; entry 5 is the zero matrix (lti_scan_states)
  SecCoef c1, c2;
};

struct WelshInst {
  LtiTable lti;
  OscMix m1, m2;
  OscMix m1b, m2b;  // m1, m2 scaled by lti.c1.b0: the time-invariant blocks get section 1's b0*x straight from the selects
  OscMix m1bb, m2bb;  // ... and by lti.c2.b0 on top (welsh_rest_block: section 1 runs pre-scaled for section 2)
  double2 lfo_rot[8];  // (cos, sin)(2*pi*j*lfo_dq/2^64): rotation table for a sine LFO
  EnvShape amp, filt;
  int w1, w2, wl, sync, routing, filter_mode, uid, voice0;
  u64 duty1_q, duty2_q, dutyl_q, lfo_dq;
  double duty1, duty2;
  double mix, depth;
  double cut_a, cut_b;
  Lp24Ripple rp;
  SecCoef fixed1, fixed2;
  double gl, gr;
  double pi_over_sr, sr;
  OscShape s1, s2, sl;              // branch-free waveform descriptions (fast path)
  double log2_25_over_sr;           // fc/sr = exp2(pct*log2(800) + log2(25/sr))
  double u_min, u_max;              // clamp of fc/sr: [1/sr, 0.49]
  double knot_max_rate;             // cutoff motion (fraction of the log range per frame) up to which knots are used
  double2 lane_rot[32];             // (cos, sin)(2*pi*kT*l*lfo_dq/2^64): a lane's LFO angle relative to its block
  double2 block_rot;                // (cos, sin)(2*pi*kBlockFrames*lfo_dq/2^64): one block further
  int rest_class, sweep_class;      // welsh_rest_kernel / welsh_sweep_kernel variant (2*lfo_amp + zero_a), -1 = does not qualify
  i64 steady_after;                 // frames after note-on from which both envelopes rest at their sustain levels
  double amp_rest;                  // 0.5 * amp.sustain: the DCA input level of a resting voice (without LFO)
  int exact_class, pad_exact;       // welsh_exact_block applies (piecewise-linear oscillators, filter envelope): 0 / 1, else -1
  int lti_ok, osc_flat;             // osc_flat: both oscillators piecewise constant (OscMix slopes are 0); lti holds this instrument's resting coefficient sets (GB_LTI=0 disables the path)
  // the same tables for a RELEASED voice whose filter envelope has run out (cutoff back at cut_a) while the
  // amplitude envelope still releases; equal to lti / m1bb / m2bb for a fixed filter (welsh_solo_kernel)
  LtiTable lti_off;
  OscMix m1bb_off, m2bb_off;
  // lti.mp1[4]^2, lti.mp2[4]^2: the homogeneous map of a whole 256-frame block (welsh_rest_tp_kernel carries the
  // filter state from one warp's block to the next with it)
  double mp32_1[4], mp32_2[4];
};
enum { FILTER_FIXED = 0, FILTER_ENVELOPE = 1, FILTER_LFO = 2 };

struct alignas(16) WelshVoice {  // 208 bytes; field pairs at 16-byte offsets are moved as 128-bit words
  i64 n_on, n_off;
  double la_on, la_off, lf_on, lf_off;
  i64 anchor;       // frame at which p1/p2/pl are valid (closed-form path); LFO anchor always
  u64 p1, p2, pl;   // PITCH path: p1/p2 are the phases at (chunk position - 1)
  u64 d1, d2;
  double cyc1, cyc2;
  double s[4];
  i64 knot_frame;   // frame of the carried coefficient knot (fast path, COEF_KNOTS); kNever = none
  double knot[6];
};

// Two kinds of CTA work.  Grouped (solo == 0): `nvoices` voices of ONE instrument starting at voice
// `voice0`, processed in rounds of W and summed through shared memory into `out`.  Solo (solo == 1):
// up to W unrelated (instrument, voice) pairs, items[voice0 .. voice0 + nvoices), one per warp, each
// warp writing its own output buffer — for instruments with fewer voices than a CTA has warps
// (e.g. a batch of one-voice patch variants).
struct CtaWork {
  int inst;       // grouped: index into the instrument table
  int voice0;     // grouped: first voice (global index); solo: first WarpItem
  int nvoices;
  int solo;
  double2* out;   // grouped: chunk-relative output buffer (node buffer or a partial)
};
struct WarpItem {
  int inst;       // index into the instrument table
  int voice;      // global voice index
  double2* out;   // chunk-relative output buffer (node buffer, or a partial of a 2..W-1 voice instrument)
};

// A warp copies its 256-frame tile row (or zeros) to its own output buffer, coalesced.
__device__ __forceinline__ void warp_store_row(const double2* tile_row, bool any, double2* out, i64 fb, i64 f0,
                                               i64 f_end, int lane) {
#pragma unroll
  for (int i = 0; i < kT; ++i) {
    const int t = i * 32 + lane;
    const i64 n = fb + t;
    if (n < f_end) out[n - f0] = any ? tile_row[t + (t >> 3)] : make_double2(0.0, 0.0);
  }
}

struct Phases {
  u64 p1, p2, pl;
};

// Phases at frame m (m >= anchor-1) of an unmodulated voice, closed form.
__device__ __forceinline__ Phases welsh_phases_at(const WelshVoice& st, const WelshInst& I, i64 m) {
  Phases ph;
  i64 k = m - st.anchor;
  u64 uk = (u64)k;
  ph.pl = st.pl + uk * I.lfo_dq;
  ph.p1 = st.p1 + uk * st.d1;
  ph.p2 = st.p2 + uk * st.d2;
  if (I.sync && k > 0 && st.d1 != 0) {
    // frames since oscillator 1 last wrapped; if that is after the anchor, oscillator 2 restarted there
    u64 j = ph.p1 / st.d1;
    if (j < uk) ph.p2 = j * st.d2;
  }
  return ph;
}

template <bool PITCH>
__device__ __noinline__ WelshVoice welsh_fold(WelshVoice st, const WelshInst* Ip, VoiceEvent ev, bool* restart) {
  // by value + noinline: note events are rare, so the call sites stay small and the caller's
  // copy of the state stays in registers.  *restart = the oscillators restart (note-on of an idle voice)
  const WelshInst& I = *Ip;
  i64 f = ev.frame;
  *restart = false;
  if (ev.type == VEV_NOTE_ON) {
    bool was = f >= st.n_on && f < st.n_off + I.amp.nr;
    double la = 0.0, lf = 0.0;
    if (was) {
      la = env_level(I.amp, st.n_on, st.n_off, st.la_on, st.la_off, f);
      lf = env_level(I.filt, st.n_on, st.n_off, st.lf_on, st.lf_off, f);
      if (PITCH) {
        st.pl += (u64)(f - 1 - st.anchor) * I.lfo_dq;
      } else {
        Phases ph = welsh_phases_at(st, I, f - 1);
        st.p1 = ph.p1; st.p2 = ph.p2; st.pl = ph.pl;
      }
      st.anchor = f - 1;
    } else {
      st.pl = 0;
      if (!PITCH) { st.p1 = 0; st.p2 = 0; }
      st.anchor = f;
    }
    st.la_on = la; st.lf_on = lf;
    st.n_on = f; st.n_off = kHeld;
    st.cyc1 = ev.cyc1; st.cyc2 = ev.cyc2;
    st.d1 = cycles_to_q(ev.cyc1); st.d2 = cycles_to_q(ev.cyc2);
    *restart = !was;
  } else {
    st.la_off = env_pre(I.amp, st.n_on, st.la_on, f);
    st.lf_off = env_pre(I.filt, st.n_on, st.lf_on, f);
    st.n_off = f;
  }
  return st;
}

// Shared out-of-line copies of the heavy per-frame pieces of the general block: its loops are unrolled
// over the lane's kT frames, and with the waveform switch (library sinpi inside) and the coefficient map
// inlined at every frame one variant of the block was 9 k instructions — an instruction-cache problem
// wherever the warps of an SM sit in different blocks (solo-warp kernel).  The general block runs for
// note events, envelope stage boundaries and pitch-LFO voices; a call per frame is noise there.
__device__ __noinline__ double wave_value_ool(int wf, u64 q, u64 duty_q, u64 seed, i64 frame) {
  return wave_value(wf, q, duty_q, seed, frame);
}
__device__ __noinline__ void welsh_coef_exact_ool(const WelshInst* Ip, double pct, SecCoef* c1, SecCoef* c2);

// One warp step: kBlockFrames frames of one voice starting at frame fb (absolute), valid frames < f_end.
// `st` is the warp-uniform voice state at fb; on return it is the state at fb + kBlockFrames.
// EV (warp-uniform) = this voice has note events inside the block; the EV=false instantiation is
// the hot path and contains no event handling at all.
template <bool PITCH, bool EV>
__device__ __forceinline__ void welsh_block(WelshVoice& st, const WelshInst& I, const VoiceEvent* __restrict__ ev,
                                            int& ei, int e_end, i64 fb, i64 f_end, int lane, u64 seed1, u64 seed2,
                                            u64 seedl, double2* tile_row, bool accumulate) {
  const i64 c0 = fb + (i64)lane * kT;
  const i64 blk_end = fb + kBlockFrames;

  // ---- lane-local note state: fold the events that precede this lane's chunk ----
  WelshVoice ls = st;
  int li = ei;
  i64 next_ev = kHeld;
  bool rs_dummy;
  if (EV) {
    while (li < e_end && ev[li].frame < c0) {
      ls = welsh_fold<PITCH>(ls, &I, ev[li], &rs_dummy);
      ++li;
    }
    next_ev = li < e_end ? ev[li].frame : kHeld;
  }

  double yp[kT], g0[kT], g1[kT], sb0[kT], sa1[kT], sa2[kT], ampf[kT];
  unsigned play_bits = 0;

  // ---- PITCH: pre-passes for the modulated phase increments and their (segmented) scans ----
  double pf[PITCH ? kT : 1];
  unsigned reset_bits = 0;  // bit j = oscillators restart at frame c0+j
  u64 ent1 = 0, ent2 = 0;   // phases at c0-1
  if (PITCH) {
    WelshVoice ps = ls;
    int pi = li;
    i64 pnext = next_ev;
    SegSum a1; a1.sum = 0; a1.reset = 0;
    unsigned pplay = 0;
#pragma unroll
    for (int j = 0; j < kT; ++j) {
      i64 n = c0 + j;
      if (EV) {
        while (n == pnext) {
          bool rs;
          ps = welsh_fold<true>(ps, &I, ev[pi], &rs);
          if (rs) reset_bits |= 1u << j;
          ++pi;
          pnext = pi < e_end ? ev[pi].frame : kHeld;
        }
      }
      bool play = n >= ps.n_on && n < ps.n_off + I.amp.nr && n < f_end;
      double f = 1.0;
      if (play) {
        pplay |= 1u << j;
        u64 pl = ps.pl + (u64)(n - ps.anchor) * I.lfo_dq;
        double l = wave_value_ool(I.wl, pl, I.dutyl_q, seedl, n);
        f = exp2(l * I.depth);
        if (reset_bits & (1u << j)) { a1.sum = 0; a1.reset = 1; }
        else a1.sum += cycles_to_q(ps.cyc1 * f);
      }
      pf[j] = f;
    }
    SegSum inc1 = segsum_warp_scan(a1, lane);
    SegSum ex1;
    ex1.sum = __shfl_up_sync(0xffffffffu, inc1.sum, 1);
    ex1.reset = __shfl_up_sync(0xffffffffu, inc1.reset, 1);
    if (lane == 0) { ex1.sum = 0; ex1.reset = 0; }
    ent1 = ex1.reset ? ex1.sum : st.p1 + ex1.sum;
    u64 tot1 = inc1.reset ? inc1.sum : st.p1 + inc1.sum;
    tot1 = __shfl_sync(0xffffffffu, tot1, 31);
    // oscillator 2: restarts on note restarts and (hard sync) on oscillator-1 wraps
    SegSum a2; a2.sum = 0; a2.reset = 0;
    {
      WelshVoice qs = ls;
      int qi = li;
      i64 qnext = next_ev;
      u64 p1 = ent1;
#pragma unroll
      for (int j = 0; j < kT; ++j) {
        i64 n = c0 + j;
        if (EV) {
          while (n == qnext) {
            bool rs;
            qs = welsh_fold<true>(qs, &I, ev[qi], &rs);
            ++qi;
            qnext = qi < e_end ? ev[qi].frame : kHeld;
          }
        }
        if (pplay & (1u << j)) {
          bool rs = (reset_bits >> j) & 1u;
          u64 d1 = cycles_to_q(qs.cyc1 * pf[j]);
          bool wrapped;
          if (rs) { p1 = 0; wrapped = false; }
          else { p1 += d1; wrapped = p1 < d1; }
          if (rs || (I.sync && wrapped)) { a2.sum = 0; a2.reset = 1; }
          else a2.sum += cycles_to_q(qs.cyc2 * pf[j]);
        }
      }
    }
    SegSum inc2 = segsum_warp_scan(a2, lane);
    SegSum ex2;
    ex2.sum = __shfl_up_sync(0xffffffffu, inc2.sum, 1);
    ex2.reset = __shfl_up_sync(0xffffffffu, inc2.reset, 1);
    if (lane == 0) { ex2.sum = 0; ex2.reset = 0; }
    ent2 = ex2.reset ? ex2.sum : st.p2 + ex2.sum;
    u64 tot2 = inc2.reset ? inc2.sum : st.p2 + inc2.sum;
    tot2 = __shfl_sync(0xffffffffu, tot2, 31);
    st.p1 = tot1;  // running phases at blk_end-1 (note folds do not touch p1/p2 on this path)
    st.p2 = tot2;
  }

  // ---- pass 1: per-frame closed forms + section-1 response from a zero state ----
  Phases ph;
  if (PITCH) { ph.p1 = ent1; ph.p2 = ent2; ph.pl = 0; }
  else ph = welsh_phases_at(ls, I, c0 - 1);
  double ps0 = 0.0, ps1 = 0.0, h00 = 1.0, h01 = 0.0, h10 = 0.0, h11 = 1.0;
#pragma unroll
  for (int j = 0; j < kT; ++j) {
    i64 n = c0 + j;
    bool restart = false;
    if (EV) {
      while (n == next_ev) {
        bool rs;
        ls = welsh_fold<PITCH>(ls, &I, ev[li], &rs);
        restart |= rs;
        ++li;
        next_ev = li < e_end ? ev[li].frame : kHeld;
        if (!PITCH) ph = welsh_phases_at(ls, I, n - 1);
      }
    }
    bool play = n >= ls.n_on && n < ls.n_off + I.amp.nr && n < f_end;
    yp[j] = 0.0; g0[j] = 0.0; g1[j] = 0.0; ampf[j] = 0.0;
    sb0[j] = 0.0; sa1[j] = 0.0; sa2[j] = 0.0;
    if (play) {
      play_bits |= 1u << j;
      // LFO
      double ld = 0.0;
      if (I.routing != LFO_NONE) {
        u64 pl;
        if (PITCH) pl = ls.pl + (u64)(n - ls.anchor) * I.lfo_dq;
        else { ph.pl += I.lfo_dq; pl = ph.pl; }
        ld = wave_value_ool(I.wl, pl, I.dutyl_q, seedl, n) * I.depth;
      }
      // oscillators
      if (PITCH) {
        if (restart) { ph.p1 = 0; ph.p2 = 0; }
        else {
          u64 d1 = cycles_to_q(ls.cyc1 * pf[j]);
          ph.p1 += d1;
          bool wrapped = ph.p1 < d1;
          if (I.sync && wrapped) ph.p2 = 0;
          else ph.p2 += cycles_to_q(ls.cyc2 * pf[j]);
        }
      } else {
        ph.p1 += ls.d1;
        bool wrapped = ph.p1 < ls.d1;
        if (I.sync && wrapped) ph.p2 = 0;
        else ph.p2 += ls.d2;
      }
      u64 du1 = I.duty1_q, du2 = I.duty2_q;
      if (I.routing == LFO_PULSE_WIDTH) {
        const double top = 1.0 - 1.0 / 9007199254740992.0;
        double a = __dadd_rn(I.duty1, __dmul_rn(0.5, ld));
        double b = __dadd_rn(I.duty2, __dmul_rn(0.5, ld));
        a = a < 0.0 ? 0.0 : (a > top ? top : a);
        b = b < 0.0 ? 0.0 : (b > top ? top : b);
        du1 = cycles_to_q(a);
        du2 = cycles_to_q(b);
      }
      double o1 = wave_value_ool(I.w1, ph.p1, du1, seed1, n);
      double o2 = wave_value_ool(I.w2, ph.p2, du2, seed2, n);
      double x = o1 * I.mix + o2 * (1.0 - I.mix);
      // filter coefficients
      SecCoef c1 = I.fixed1, c2 = I.fixed2;
      if (I.filter_mode != FILTER_FIXED) {
        double pct;
        if (I.filter_mode == FILTER_ENVELOPE) {
          double fe = env_level(I.filt, ls.n_on, ls.n_off, ls.lf_on, ls.lf_off, n);
          pct = I.cut_a + I.cut_b * fe;
        } else {
          pct = I.cut_a * (1.0 + ld);
        }
        welsh_coef_exact_ool(&I, pct, &c1, &c2);  // clamps pct to [0,1] and fc to [1 Hz, 0.49 sr]
      }
      sb0[j] = c2.b0; sa1[j] = c2.a1; sa2[j] = c2.a2;
      // amplitude
      double ae = env_level(I.amp, ls.n_on, ls.n_off, ls.la_on, ls.la_off, n);
      ampf[j] = ae * (I.routing == LFO_AMPLITUDE ? 0.5 * (1.0 + ld) : 0.5);
      // section 1 from zero state, plus homogeneous response
      double bx = c1.b0 * x;
      double y = bx + ps0;
      yp[j] = y; g0[j] = h00; g1[j] = h01;
      double n0 = 2.0 * bx + c1.a1 * y + ps1;
      ps1 = bx + c1.a2 * y;
      ps0 = n0;
      double t00 = c1.a1 * h00 + h10, t01 = c1.a1 * h01 + h11;
      h10 = c1.a2 * h00; h11 = c1.a2 * h01;
      h00 = t00; h01 = t01;
    }
  }
  // ---- scan section 1, fix up, run section 2 from zero state ----
  double e0, e1, end0, end1;
  {
    Affine2 a;
    a.m00 = h00; a.m01 = h01; a.m10 = h10; a.m11 = h11; a.v0 = ps0; a.v1 = ps1;
    Affine2 inc = affine_warp_scan(a, lane);
    affine_lane_entry(inc, lane, st.s[0], st.s[1], e0, e1, end0, end1);
    st.s[0] = end0; st.s[1] = end1;
  }
  ps0 = 0.0; ps1 = 0.0; h00 = 1.0; h01 = 0.0; h10 = 0.0; h11 = 1.0;
#pragma unroll
  for (int j = 0; j < kT; ++j) {
    if (play_bits & (1u << j)) {
      double x = yp[j] + g0[j] * e0 + g1[j] * e1;  // true section-1 output
      double bx = sb0[j] * x;
      double y = bx + ps0;
      yp[j] = y; g0[j] = h00; g1[j] = h01;
      double n0 = 2.0 * bx + sa1[j] * y + ps1;
      ps1 = bx + sa2[j] * y;
      ps0 = n0;
      double t00 = sa1[j] * h00 + h10, t01 = sa1[j] * h01 + h11;
      h10 = sa2[j] * h00; h11 = sa2[j] * h01;
      h00 = t00; h01 = t01;
    }
  }
  {
    Affine2 a;
    a.m00 = h00; a.m01 = h01; a.m10 = h10; a.m11 = h11; a.v0 = ps0; a.v1 = ps1;
    Affine2 inc = affine_warp_scan(a, lane);
    affine_lane_entry(inc, lane, st.s[2], st.s[3], e0, e1, end0, end1);
    st.s[2] = end0; st.s[3] = end1;
  }
  // ---- amplitude, DCA, into the CTA tile ----
#pragma unroll
  for (int j = 0; j < kT; ++j) {
    double m = (yp[j] + g0[j] * e0 + g1[j] * e1) * ampf[j];
    int t = lane * kT + j;
    double2 o = make_double2(m * I.gl, m * I.gr);
    if (accumulate) {
      double2 p = tile_row[t + (t >> 3)];
      o.x += p.x; o.y += p.y;
    }
    tile_row[t + (t >> 3)] = o;
  }
  // ---- advance the warp-uniform note state over this block's events ----
  if (EV) {
    while (ei < e_end && ev[ei].frame < blk_end) {
      st = welsh_fold<PITCH>(st, &I, ev[ei], &rs_dummy);
      ++ei;
    }
  }
}

// ---- fast path ---------------------------------------------------------------------------------
// An ADSR restricted to one stage is a quadratic in the frame index: level = q0 + w*(q1 + q2*w),
// w = w0 + j*dw.  A lane whose kT frames lie inside one stage of each envelope evaluates them with
// 3 FMAs per frame and no integer work.
struct EnvSeg {
  double q0, q1, q2, w0, dw;
};
__device__ __forceinline__ bool env_segment(const EnvShape& s, i64 n_on, i64 n_off, double l_on, double l_off, i64 c0,
                                            EnvSeg& g) {
  g.q0 = 0.0; g.q1 = 0.0; g.q2 = 0.0; g.w0 = 0.0; g.dw = 0.0;
  const i64 last = c0 + (kT - 1);
  if (last < n_off) {
    const i64 k = c0 - n_on;
    if (k + (kT - 1) < s.na) {  // attack: l_on + (1-l_on) * t(2-t)
      g.w0 = (double)k * s.inv_na; g.dw = s.inv_na;
      g.q0 = l_on; g.q1 = 2.0 * (1.0 - l_on); g.q2 = -(1.0 - l_on);
      return true;
    }
    if (k < s.na) return false;
    const i64 k2 = k - s.na;
    if (k2 + (kT - 1) < s.nd) {  // decay: S + (1-S) * u^2
      g.w0 = 1.0 - (double)k2 * s.inv_nd; g.dw = -s.inv_nd;
      g.q0 = s.sustain; g.q2 = 1.0 - s.sustain;
      return true;
    }
    if (k2 < s.nd) return false;
    g.q0 = s.sustain;
    return true;
  }
  if (c0 >= n_off) {
    const i64 k = c0 - n_off;
    if (k + (kT - 1) < s.nr) {  // release: l_off * u^2
      g.w0 = 1.0 - (double)k * s.inv_nr; g.dw = -s.inv_nr;
      g.q2 = l_off;
      return true;
    }
    return k >= s.nr;  // silent tail
  }
  return false;
}
__device__ __forceinline__ double env_seg_at(const EnvSeg& g, int j) {
  double w = fma((double)j, g.dw, g.w0);
  return fma(w, fma(g.q2, w, g.q1), g.q0);
}

// Lane classification for the fast path: 0 = needs the general path, 1 = all kT frames idle,
// 2 = all kT frames sounding inside single envelope stages.
struct NoteWords {  // the part of a voice record that decides the path of a block
  i64 n_on, n_off;
  double la_on, la_off, lf_on, lf_off;
};
__device__ __forceinline__ int welsh_lane_class(const NoteWords& st, const WelshInst& I, i64 c0, i64 f_end,
                                                EnvSeg& amp, EnvSeg& filt) {
  const i64 last = c0 + (kT - 1);
  const i64 idle_at = st.n_off + I.amp.nr;
  if (c0 >= f_end || last < st.n_on || c0 >= idle_at) return 1;
  if (!(c0 >= st.n_on && last < idle_at && last < f_end)) return 0;
  if (!env_segment(I.amp, st.n_on, st.n_off, st.la_on, st.la_off, c0, amp)) return 0;
  if (I.filter_mode == FILTER_ENVELOPE &&
      !env_segment(I.filt, st.n_on, st.n_off, st.lf_on, st.lf_off, c0, filt))
    return 0;
  return 2;
}

// Exact coefficient set for cutoff fraction `pct` of the 25 Hz..20 kHz log range.
__device__ __forceinline__ void welsh_coef_exact(const WelshInst& I, double pct, SecCoef& c1, SecCoef& c2) {
  pct = pct < 0.0 ? 0.0 : (pct > 1.0 ? 1.0 : pct);
  double u = exp2_ranged(fma(pct, kLog2_800, I.log2_25_over_sr));
  u = u > I.u_max ? I.u_max : u;
  u = u < I.u_min ? I.u_min : u;
  lp24_from_u(I.rp, u, c1, c2);
}

__device__ __noinline__ void welsh_coef_exact_ool(const WelshInst* Ip, double pct, SecCoef* c1, SecCoef* c2) {
  welsh_coef_exact(*Ip, pct, *c1, *c2);
}

// Quadratic through (0,k0), (kT/2,k4), (kT,k8) in Newton form: c(j) = k0 + j*(d1 + (j - kT/2)*d2).
struct Quad {
  double k0, d1, d2;
};
__device__ __forceinline__ Quad quad_fit(double k0, double k4, double k8) {
  Quad q;
  q.k0 = k0;
  q.d1 = (k4 - k0) * (2.0 / kT);
  q.d2 = ((k8 - k4) * (2.0 / kT) - q.d1) * (1.0 / kT);
  return q;
}
__device__ __forceinline__ double quad_at(const Quad& q, int j) {
  return fma((double)j, fma((double)(j - kT / 2), q.d2, q.d1), q.k0);
}

// Largest cutoff motion (fraction of the log range per frame) for which the per-frame coefficient
// sets are taken from the quadratic through exact knots every kT/2 frames instead of being evaluated
// exactly: the interpolation error grows with the cube of the rate and is <= 1e-11 absolute here
// (docs/ORACLE_SPEC.md, "coefficient knots").
constexpr double kKnotMaxRate = 1.0e-5;  // default of WelshInst::knot_max_rate (GB_KNOT_MAX_RATE overrides; 0 = always exact)

enum { COEF_FIXED = 0, COEF_EXACT = 1, COEF_KNOTS = 2 };

// Fast path of the Welsh voice: same contract as welsh_block<false,false>, for blocks in which every
// lane is class 1 (idle) or 2 (sounding inside single envelope stages).  CMODE selects how the
// per-frame 24 dB coefficient sets are produced; COEF_KNOTS additionally needs every lane in class 2.
template <int CMODE, bool ALL_ON>
__device__ __forceinline__ void welsh_block_fast(WelshVoice& st, const WelshInst& I, i64 fb, int lane, int cls,
                                                 const EnvSeg& aseg, const EnvSeg& fseg, u64 seed1, u64 seed2,
                                                 u64 seedl, double2* tile_row, bool accumulate, double* knot_out) {
  const i64 c0 = fb + (i64)lane * kT;
  double yp[kT], g0[kT], g1[kT], ampf[kT];
  double sb0[CMODE == COEF_EXACT ? kT : 1], sa1[CMODE == COEF_EXACT ? kT : 1], sa2[CMODE == COEF_EXACT ? kT : 1];
  double ps0 = 0.0, ps1 = 0.0, h00 = 1.0, h01 = 0.0, h10 = 0.0, h11 = 1.0;
  const bool on = ALL_ON || cls == 2;  // ALL_ON: every lane sounds, no divergence at all
  if (!on) {
#pragma unroll
    for (int j = 0; j < kT; ++j) {
      yp[j] = 0.0; g0[j] = 0.0; g1[j] = 0.0; ampf[j] = 0.0;
      if (CMODE == COEF_EXACT) { sb0[j] = 0.0; sa1[j] = 0.0; sa2[j] = 0.0; }
    }
  }
  // ---- coefficient knots: exact at j = kT/2 and j = kT; j = 0 comes from the previous lane ----
  Quad qb1, qa11, qa21, qb2, qa12, qa22;
  if (CMODE == COEF_KNOTS) {
    SecCoef m1, m2, e1c, e2c;
    welsh_coef_exact(I, fma(I.cut_b, env_seg_at(fseg, kT / 2), I.cut_a), m1, m2);
    welsh_coef_exact(I, fma(I.cut_b, env_seg_at(fseg, kT), I.cut_a), e1c, e2c);
    SecCoef s1c, s2c;
    s1c.b0 = __shfl_up_sync(0xffffffffu, e1c.b0, 1); s1c.a1 = __shfl_up_sync(0xffffffffu, e1c.a1, 1);
    s1c.a2 = __shfl_up_sync(0xffffffffu, e1c.a2, 1); s2c.b0 = __shfl_up_sync(0xffffffffu, e2c.b0, 1);
    s2c.a1 = __shfl_up_sync(0xffffffffu, e2c.a1, 1); s2c.a2 = __shfl_up_sync(0xffffffffu, e2c.a2, 1);
    if (lane == 0) {
      if (st.knot_frame == fb) {
        s1c.b0 = st.knot[0]; s1c.a1 = st.knot[1]; s1c.a2 = st.knot[2];
        s2c.b0 = st.knot[3]; s2c.a1 = st.knot[4]; s2c.a2 = st.knot[5];
      } else {
        welsh_coef_exact(I, fma(I.cut_b, env_seg_at(fseg, 0), I.cut_a), s1c, s2c);
      }
    }
    qb1 = quad_fit(s1c.b0, m1.b0, e1c.b0); qa11 = quad_fit(s1c.a1, m1.a1, e1c.a1); qa21 = quad_fit(s1c.a2, m1.a2, e1c.a2);
    qb2 = quad_fit(s2c.b0, m2.b0, e2c.b0); qa12 = quad_fit(s2c.a1, m2.a1, e2c.a1); qa22 = quad_fit(s2c.a2, m2.a2, e2c.a2);
    if (lane == 31) {  // the knot at fb + kBlockFrames is the next block's first knot
      knot_out[0] = e1c.b0; knot_out[1] = e1c.a1; knot_out[2] = e1c.a2;
      knot_out[3] = e2c.b0; knot_out[4] = e2c.a1; knot_out[5] = e2c.a2;
    }
  }
  if (on) {
    Phases ph = welsh_phases_at(st, I, c0 - 1);
    // sine LFO: one sincospi per lane chunk, then an angle-addition rotation per frame
    double ls = 0.0, lc = 0.0;
    const bool lfo_on = I.routing != LFO_NONE;
    const bool lfo_sine = I.wl == W_SINE;
    if (lfo_on && lfo_sine) sincospi(2.0 * pos_of(ph.pl + I.lfo_dq), &ls, &lc);
#pragma unroll
    for (int j = 0; j < kT; ++j) {
      const i64 n = c0 + j;
      double ld = 0.0;
      if (lfo_on) {
        ph.pl += I.lfo_dq;
        double l = lfo_sine ? fma(ls, I.lfo_rot[j].x, lc * I.lfo_rot[j].y)
                            : osc_eval(I.sl, ph.pl, I.sl.thresh, seedl, n);
        ld = l * I.depth;
      }
      ph.p1 += st.d1;
      const bool wrapped = ph.p1 < st.d1;
      ph.p2 = (I.sync && wrapped) ? 0ull : ph.p2 + st.d2;
      u64 du1 = I.s1.thresh, du2 = I.s2.thresh;
      if (I.routing == LFO_PULSE_WIDTH) {
        const double top = 1.0 - 1.0 / 9007199254740992.0;
        double a = __dadd_rn(I.duty1, __dmul_rn(0.5, ld));
        double b = __dadd_rn(I.duty2, __dmul_rn(0.5, ld));
        a = a < 0.0 ? 0.0 : (a > top ? top : a);
        b = b < 0.0 ? 0.0 : (b > top ? top : b);
        if (I.w1 == W_PULSE) du1 = cycles_to_q(a);
        if (I.w2 == W_PULSE) du2 = cycles_to_q(b);
      }
      const double o1 = osc_eval(I.s1, ph.p1, du1, seed1, n);
      const double o2 = osc_eval(I.s2, ph.p2, du2, seed2, n);
      const double x = o1 * I.mix + o2 * (1.0 - I.mix);
      SecCoef c1 = I.fixed1;
      if (CMODE == COEF_EXACT) {
        SecCoef c2;
        const double pct = I.filter_mode == FILTER_ENVELOPE ? fma(I.cut_b, env_seg_at(fseg, j), I.cut_a)
                                                            : I.cut_a * (1.0 + ld);
        welsh_coef_exact_ool(&I, pct, &c1, &c2);  // one shared copy: this block is unrolled over kT frames
        sb0[j] = c2.b0; sa1[j] = c2.a1; sa2[j] = c2.a2;
      } else if (CMODE == COEF_KNOTS) {
        c1.b0 = quad_at(qb1, j); c1.a1 = quad_at(qa11, j); c1.a2 = quad_at(qa21, j);
      }
      ampf[j] = env_seg_at(aseg, j) * (I.routing == LFO_AMPLITUDE ? fma(0.5, ld, 0.5) : 0.5);
      const double bx = c1.b0 * x;
      const double y = bx + ps0;
      yp[j] = y; g0[j] = h00; g1[j] = h01;
      const double n0 = 2.0 * bx + c1.a1 * y + ps1;
      ps1 = bx + c1.a2 * y;
      ps0 = n0;
      const double t00 = c1.a1 * h00 + h10, t01 = c1.a1 * h01 + h11;
      h10 = c1.a2 * h00; h11 = c1.a2 * h01;
      h00 = t00; h01 = t01;
    }
  }
  double e0, e1, end0, end1;
  {
    Affine2 a;
    a.m00 = h00; a.m01 = h01; a.m10 = h10; a.m11 = h11; a.v0 = ps0; a.v1 = ps1;
    Affine2 inc = affine_warp_scan(a, lane);
    affine_lane_entry(inc, lane, st.s[0], st.s[1], e0, e1, end0, end1);
    st.s[0] = end0; st.s[1] = end1;
  }
  ps0 = 0.0; ps1 = 0.0; h00 = 1.0; h01 = 0.0; h10 = 0.0; h11 = 1.0;
  if (on) {
#pragma unroll
    for (int j = 0; j < kT; ++j) {
      SecCoef c2 = I.fixed2;
      if (CMODE == COEF_EXACT) { c2.b0 = sb0[j]; c2.a1 = sa1[j]; c2.a2 = sa2[j]; }
      else if (CMODE == COEF_KNOTS) { c2.b0 = quad_at(qb2, j); c2.a1 = quad_at(qa12, j); c2.a2 = quad_at(qa22, j); }
      const double x = yp[j] + g0[j] * e0 + g1[j] * e1;
      const double bx = c2.b0 * x;
      const double y = bx + ps0;
      yp[j] = y; g0[j] = h00; g1[j] = h01;
      const double n0 = 2.0 * bx + c2.a1 * y + ps1;
      ps1 = bx + c2.a2 * y;
      ps0 = n0;
      const double t00 = c2.a1 * h00 + h10, t01 = c2.a1 * h01 + h11;
      h10 = c2.a2 * h00; h11 = c2.a2 * h01;
      h00 = t00; h01 = t01;
    }
  }
  {
    Affine2 a;
    a.m00 = h00; a.m01 = h01; a.m10 = h10; a.m11 = h11; a.v0 = ps0; a.v1 = ps1;
    Affine2 inc = affine_warp_scan(a, lane);
    affine_lane_entry(inc, lane, st.s[2], st.s[3], e0, e1, end0, end1);
    st.s[2] = end0; st.s[3] = end1;
  }
#pragma unroll
  for (int j = 0; j < kT; ++j) {
    const double m = (yp[j] + g0[j] * e0 + g1[j] * e1) * ampf[j];
    const int t = lane * kT + j;
    double2 o = make_double2(m * I.gl, m * I.gr);
    if (accumulate) {
      double2 p = tile_row[t + (t >> 3)];
      o.x += p.x; o.y += p.y;
    }
    tile_row[t + (t >> 3)] = o;
  }
}

// Sum the W per-voice tiles of one block and store 16 bytes per frame.
template <int W>
__device__ __forceinline__ void cta_reduce_store(const double2* tiles, const int* s_active, double2* out, i64 fb,
                                                 i64 f0, i64 f_end) {
  for (int t = threadIdx.x; t < kBlockFrames; t += 32 * W) {
    i64 n = fb + t;
    if (n < f_end) {
      double l = 0.0, r = 0.0;
      int u = t + (t >> 3);
#pragma unroll
      for (int w = 0; w < W; ++w) {
        if (s_active[w]) {
          double2 v = tiles[w * kTileStride + u];
          l += v.x; r += v.y;
        }
      }
      out[n - f0] = make_double2(l, r);
    }
  }
}

// ---- the steady-state block of the common configuration --------------------------------------------
// Preconditions (checked by the caller, all warp-uniform): both oscillators piecewise linear
// (OscShape kind 0), no hard sync, LFO either unused or a sine routed to amplitude, cutoff driven by
// the filter envelope and moving <= kKnotMaxRate per frame, every lane sounding inside single
// envelope stages.  Everything the general fast path decides per frame is a compile-time constant
// here; the arithmetic is identical to welsh_block_fast<COEF_KNOTS, true>.
constexpr int kParkWords = 16;  // doubles parked per thread by welsh_block_simple (14) / welsh_exact_block (16)

// 1 + pos_of(q): the top 52 phase bits as the mantissa of a double in [1,2)
__device__ __forceinline__ double pos1_of(u64 q) {
  return __longlong_as_double((long long)(0x3FF0000000000000ull | (q >> 12)));
}
// ZERO_A: both waveforms are piecewise constant (square, pulse, debug levels, none) — no slope term.
template <bool ZERO_A>
__device__ __forceinline__ double osc_mix_eval(const OscMix& m1, u64 t1, u64 p1, const OscMix& m2, u64 t2, u64 p2) {
  const bool lo1 = p1 < t1, lo2 = p2 < t2;
  const double b = (lo1 ? m1.b_lo : m1.b_hi) + (lo2 ? m2.b_lo : m2.b_hi);
  if (ZERO_A) return b;
  return fma(lo1 ? m1.a_lo : m1.a_hi, pos1_of(p1), fma(lo2 ? m2.a_lo : m2.a_hi, pos1_of(p2), b));
}
// One frame of a transposed-DF2 section (b1 = 2 b0, b2 = b0): returns y, advances (s0, s1).
__device__ __forceinline__ double lp_step(double b0, double a1, double a2, double x, double& s0, double& s1) {
  const double bx = b0 * x;
  const double y = bx + s0;
  s0 = fma(a1, y, fma(2.0, bx, s1));
  s1 = fma(a2, y, bx);
  return y;
}

// the same with bx = b0 * x supplied by the caller
__device__ __forceinline__ double lp_step_bx(double bx, double a1, double a2, double& s0, double& s1) {
  const double y = bx + s0;
  s0 = fma(a1, y, fma(2.0, bx, s1));
  s1 = fma(a2, y, bx);
  return y;
}

template <bool LFO_AMP, bool ZERO_A>
__device__ __forceinline__ void welsh_block_simple(WelshVoice* vp, const WelshInst* Ip, i64 fb, int lane, EnvSeg aseg,
                                                EnvSeg fseg, double2* tile_row, double* park) {
  // `park` = this thread's column of a [kParkWords][blockDim.x] shared array: values that are only
  // needed after pass 1 wait there so that the oscillator constants fit in registers.
  const WelshInst& I = *Ip;
  const int pstride = blockDim.x;
  const i64 c0 = fb + (i64)lane * kT;
  // ---- coefficient knots ----
  Quad qb1, qa11, qa21;
  {
    SecCoef m1, m2, e1c, e2c, s1c, s2c;
    welsh_coef_exact(I, fma(I.cut_b, env_seg_at(fseg, kT / 2), I.cut_a), m1, m2);
    welsh_coef_exact(I, fma(I.cut_b, env_seg_at(fseg, kT), I.cut_a), e1c, e2c);
    s1c.b0 = __shfl_up_sync(0xffffffffu, e1c.b0, 1); s1c.a1 = __shfl_up_sync(0xffffffffu, e1c.a1, 1);
    s1c.a2 = __shfl_up_sync(0xffffffffu, e1c.a2, 1); s2c.b0 = __shfl_up_sync(0xffffffffu, e2c.b0, 1);
    s2c.a1 = __shfl_up_sync(0xffffffffu, e2c.a1, 1); s2c.a2 = __shfl_up_sync(0xffffffffu, e2c.a2, 1);
    if (lane == 0) {
      if (vp->knot_frame == fb) {
        s1c.b0 = vp->knot[0]; s1c.a1 = vp->knot[1]; s1c.a2 = vp->knot[2];
        s2c.b0 = vp->knot[3]; s2c.a1 = vp->knot[4]; s2c.a2 = vp->knot[5];
      } else {
        welsh_coef_exact(I, fma(I.cut_b, env_seg_at(fseg, 0), I.cut_a), s1c, s2c);
      }
    }
    __syncwarp();
    if (lane == 31) {
      vp->knot[0] = e1c.b0; vp->knot[1] = e1c.a1; vp->knot[2] = e1c.a2;
      vp->knot[3] = e2c.b0; vp->knot[4] = e2c.a1; vp->knot[5] = e2c.a2;
    }
    qb1 = quad_fit(s1c.b0, m1.b0, e1c.b0); qa11 = quad_fit(s1c.a1, m1.a1, e1c.a1); qa21 = quad_fit(s1c.a2, m1.a2, e1c.a2);
    const Quad qb2 = quad_fit(s2c.b0, m2.b0, e2c.b0), qa12 = quad_fit(s2c.a1, m2.a1, e2c.a1),
               qa22 = quad_fit(s2c.a2, m2.a2, e2c.a2);
    park[0 * pstride] = qb2.k0; park[1 * pstride] = qb2.d1; park[2 * pstride] = qb2.d2;
    park[3 * pstride] = qa12.k0; park[4 * pstride] = qa12.d1; park[5 * pstride] = qa12.d2;
    park[6 * pstride] = qa22.k0; park[7 * pstride] = qa22.d1; park[8 * pstride] = qa22.d2;
    // the amplitude envelope segment, with the DCA's 0.5 folded in
    park[9 * pstride] = 0.5 * aseg.q0; park[10 * pstride] = 0.5 * aseg.q1; park[11 * pstride] = 0.5 * aseg.q2;
    park[12 * pstride] = aseg.w0; park[13 * pstride] = aseg.dw;
  }
  // ---- phases at c0 - 1 (closed form), LFO base angle ----
  const u64 k = (u64)(c0 - 1 - vp->anchor);
  const u64 d1 = vp->d1, d2 = vp->d2;
  u64 p1 = vp->p1 + k * d1, p2 = vp->p2 + k * d2;
  double lsd = 0.0, lcd = 0.0;  // depth * (sin, cos) of the LFO angle at the lane's first frame
  if (LFO_AMP) {
    double ls, lc;
    sincos_phase(vp->pl + (k + 1) * I.lfo_dq, &ls, &lc);
    lsd = ls * I.depth; lcd = lc * I.depth;
  }
  // ---- pass 1: oscillators + section 1 from a zero state with its homogeneous response ----
  double yp[kT], g0[kT], g1[kT];
  double ps0 = 0.0, ps1 = 0.0, h00 = 1.0, h01 = 0.0, h10 = 0.0, h11 = 1.0;
  {
    const OscMix o1 = I.m1, o2 = I.m2;
    const u64 t1 = I.s1.thresh, t2 = I.s2.thresh;
#pragma unroll
    for (int j = 0; j < kT; ++j) {
      p1 += d1;
      p2 += d2;
      const double x = osc_mix_eval<ZERO_A>(o1, t1, p1, o2, t2, p2);
      const double b0 = quad_at(qb1, j), a1 = quad_at(qa11, j), a2 = quad_at(qa21, j);
      g0[j] = h00; g1[j] = h01;
      yp[j] = lp_step(b0, a1, a2, x, ps0, ps1);
      const double t00 = fma(a1, h00, h10), t01 = fma(a1, h01, h11);
      h10 = a2 * h00; h11 = a2 * h01;
      h00 = t00; h01 = t01;
    }
  }
  double e0, e1, end0, end1;
  {
    Affine2 a;
    a.m00 = h00; a.m01 = h01; a.m10 = h10; a.m11 = h11; a.v0 = ps0; a.v1 = ps1;
    affine_scan_states(a, lane, vp->s[0], vp->s[1], e0, e1, end0, end1);
  }
  const double ns0 = end0, ns1 = end1;
  // ---- pass 2: section 2 on the fixed-up section-1 output ----
  {
    Quad qb2, qa12, qa22;
    qb2.k0 = park[0 * pstride]; qb2.d1 = park[1 * pstride]; qb2.d2 = park[2 * pstride];
    qa12.k0 = park[3 * pstride]; qa12.d1 = park[4 * pstride]; qa12.d2 = park[5 * pstride];
    qa22.k0 = park[6 * pstride]; qa22.d1 = park[7 * pstride]; qa22.d2 = park[8 * pstride];
    ps0 = 0.0; ps1 = 0.0; h00 = 1.0; h01 = 0.0; h10 = 0.0; h11 = 1.0;
#pragma unroll
    for (int j = 0; j < kT; ++j) {
      const double b0 = quad_at(qb2, j), a1 = quad_at(qa12, j), a2 = quad_at(qa22, j);
      const double x = fma(g1[j], e1, fma(g0[j], e0, yp[j]));
      g0[j] = h00; g1[j] = h01;
      yp[j] = lp_step(b0, a1, a2, x, ps0, ps1);
      const double t00 = fma(a1, h00, h10), t01 = fma(a1, h01, h11);
      h10 = a2 * h00; h11 = a2 * h01;
      h00 = t00; h01 = t01;
    }
  }
  {
    Affine2 a;
    a.m00 = h00; a.m01 = h01; a.m10 = h10; a.m11 = h11; a.v0 = ps0; a.v1 = ps1;
    affine_scan_states(a, lane, vp->s[2], vp->s[3], e0, e1, end0, end1);
  }
  // ---- amplitude (envelope segment x LFO), DCA, CTA tile ----
  EnvSeg as;
  as.q0 = park[9 * pstride]; as.q1 = park[10 * pstride]; as.q2 = park[11 * pstride];
  as.w0 = park[12 * pstride]; as.dw = park[13 * pstride];
  const double gl = I.gl, gr = I.gr;
  double2* row = tile_row + lane * (kT + 1);
#pragma unroll
  for (int j = 0; j < kT; ++j) {
    double amp = env_seg_at(as, j);
    if (LFO_AMP) {
      const double2 rot = I.lfo_rot[j];
      amp *= fma(lsd, rot.x, fma(lcd, rot.y, 1.0));
    }
    const double m = fma(g1[j], e1, fma(g0[j], e0, yp[j])) * amp;
    const double2 p = row[j];  // the row holds the warp's earlier voices of this block (or zeros)
    row[j] = make_double2(fma(m, gl, p.x), fma(m, gr, p.y));
  }
  __syncwarp();
  if (lane == 0) {
    vp->s[0] = ns0; vp->s[1] = ns1; vp->s[2] = end0; vp->s[3] = end1;
    vp->knot_frame = fb + kBlockFrames;
  }
  __syncwarp();
}

// ---- the time-invariant block --------------------------------------------------------------------
// Preconditions (warp-uniform, checked by the caller): as welsh_block_simple, but the cutoff does not
// move over the block — the instrument's filter is fixed, or every lane's filter envelope rests at its
// sustain level.  Both coefficient sets are then the per-instrument constants of LtiTable: nothing is
// evaluated or interpolated per frame, the homogeneous response rows come from the table instead of
// being tracked per lane, and the scan moves 2-vectors under constant span maps.
//   NV        voices of the same instrument rendered in lockstep by this warp (1 or 2).  Two voices
//             give every dependent chain (recurrence, scan steps, table loads) an independent twin,
//             which is what hides the latencies at 4 warps per scheduler.
//   AMP_FLAT  the amplitude envelope rests too (level I.amp_rest); otherwise `aseg` is the lane's segment.
template <bool LFO_AMP, bool ZERO_A, bool AMP_FLAT, int NV>
__device__ __forceinline__ void welsh_block_lti(WelshVoice* const (&vp)[NV], const WelshInst* Ip, i64 fb, int lane,
                                                const EnvSeg& aseg, double2* tile_row) {
  static_assert(offsetof(WelshVoice, anchor) % 16 == 0 && offsetof(WelshVoice, p2) % 16 == 0 &&
                offsetof(WelshVoice, d1) % 16 == 0 && offsetof(WelshVoice, s) % 16 == 0 && sizeof(WelshVoice) % 16 == 0,
                "WelshVoice field pairs must sit on 16-byte boundaries");
  const WelshInst& I = *Ip;
  const LtiTable& L = I.lti;
  const i64 c0 = fb + (i64)lane * kT;
  u64 p1[NV], p2[NV], d1[NV], d2[NV];
  double lsd[NV], lcd[NV];  // (depth x level) * (sin, cos) of the LFO angle at the lane's first frame
  double s[NV][4];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const ulonglong2 ap = *reinterpret_cast<const ulonglong2*>(&vp[v]->anchor);  // anchor, p1
    const ulonglong2 pp = *reinterpret_cast<const ulonglong2*>(&vp[v]->p2);      // p2, pl
    const ulonglong2 dd = *reinterpret_cast<const ulonglong2*>(&vp[v]->d1);      // d1, d2
    const double2 sa = *reinterpret_cast<const double2*>(&vp[v]->s[0]), sb = *reinterpret_cast<const double2*>(&vp[v]->s[2]);
    s[v][0] = sa.x; s[v][1] = sa.y; s[v][2] = sb.x; s[v][3] = sb.y;
    const u64 k = (u64)(c0 - 1) - ap.x;
    d1[v] = dd.x; d2[v] = dd.y;
    p1[v] = ap.y + k * dd.x; p2[v] = pp.x + k * dd.y;
    lsd[v] = 0.0; lcd[v] = 0.0;
    if (LFO_AMP) {
      double ls, lc;
      sincos_phase(pp.y + (k + 1) * I.lfo_dq, &ls, &lc);
      const double dl = AMP_FLAT ? I.depth * I.amp_rest : I.depth;
      lsd[v] = ls * dl; lcd[v] = lc * dl;
    }
  }
  double yp[NV][kT];
  double ps0[NV], ps1[NV];
  {
    const OscMix o1 = I.m1b, o2 = I.m2b;  // pre-scaled by section 1's b0
    const u64 t1 = I.s1.thresh, t2 = I.s2.thresh;
    const double a1 = L.c1.a1, a2 = L.c1.a2;
#pragma unroll
    for (int v = 0; v < NV; ++v) { ps0[v] = 0.0; ps1[v] = 0.0; }
#pragma unroll
    for (int j = 0; j < kT; ++j) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        p1[v] += d1[v];
        p2[v] += d2[v];
        yp[v][j] = lp_step_bx(osc_mix_eval<ZERO_A>(o1, t1, p1[v], o2, t2, p2[v]), a1, a2, ps0[v], ps1[v]);
      }
    }
  }
  double e0[NV], e1[NV], end[NV][4];
#pragma unroll
  for (int v = 0; v < NV; ++v) lti_scan_states(ps0[v], ps1[v], L.mp1, lane, s[v][0], s[v][1], e0[v], e1[v], end[v][0], end[v][1]);
  {
    const double b0 = L.c2.b0, a1 = L.c2.a1, a2 = L.c2.a2;
#pragma unroll
    for (int v = 0; v < NV; ++v) { ps0[v] = 0.0; ps1[v] = 0.0; }
#pragma unroll
    for (int j = 0; j < kT; ++j) {
      const double2 g = *reinterpret_cast<const double2*>(L.g1[j]);
#pragma unroll
      for (int v = 0; v < NV; ++v)
        yp[v][j] = lp_step(b0, a1, a2, fma(g.y, e1[v], fma(g.x, e0[v], yp[v][j])), ps0[v], ps1[v]);
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) lti_scan_states(ps0[v], ps1[v], L.mp2, lane, s[v][2], s[v][3], e0[v], e1[v], end[v][2], end[v][3]);
  // ---- amplitude (envelope x LFO), DCA, into the warp's tile row (voices of a pair are summed first) ----
  EnvSeg as = aseg;
  if (!AMP_FLAT) { as.q0 *= 0.5; as.q1 *= 0.5; as.q2 *= 0.5; }
  const double arest = I.amp_rest;
  const double gl = I.gl, gr = I.gr;
  double2* row = tile_row + lane * (kT + 1);
#pragma unroll
  for (int j = 0; j < kT; ++j) {
    const double2 g = *reinterpret_cast<const double2*>(L.g2[j]);
    double2 rot = make_double2(0.0, 0.0);
    if (LFO_AMP) rot = I.lfo_rot[j];
    const double env = AMP_FLAT ? arest : env_seg_at(as, j);
    double m = 0.0;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      double amp = env;
      if (LFO_AMP) amp = AMP_FLAT ? fma(lsd[v], rot.x, fma(lcd[v], rot.y, arest))
                                  : env * fma(lsd[v], rot.x, fma(lcd[v], rot.y, 1.0));
      m = fma(fma(g.y, e1[v], fma(g.x, e0[v], yp[v][j])), amp, m);
    }
    const double2 p = row[j];  // the row holds the warp's earlier voices of this block (or zeros)
    row[j] = make_double2(fma(m, gl, p.x), fma(m, gr, p.y));
  }
  __syncwarp();
  if (lane == 0) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      *reinterpret_cast<double2*>(&vp[v]->s[0]) = make_double2(end[v][0], end[v][1]);
      *reinterpret_cast<double2*>(&vp[v]->s[2]) = make_double2(end[v][2], end[v][3]);
      vp[v]->knot_frame = kNever;  // no coefficient knot is carried out of a time-invariant block
    }
  }
  __syncwarp();
}

// Out-of-line forms of the specialised blocks for the solo-warp kernel.  There every warp of a CTA is a
// different instrument at a different stage of its note, so the warps of an SM are spread over many of
// the kernel's code paths at once; with every block inlined (and cloned per call site) the live code
// was 340 KB and 75 % of the stall samples were instruction-cache misses (`no_instructions`).  One
// shared copy per variant keeps the working set small; the grouped kernel, whose warps march in step,
// keeps the inlined forms (no ABI register saves on its hot path).
template <bool LFO_AMP, bool ZERO_A, bool AMP_FLAT>
__device__ __noinline__ void welsh_block_lti_ool(WelshVoice* vp, const WelshInst* Ip, i64 fb, int lane,
                                                 const EnvSeg* aseg, double2* tile_row) {
  WelshVoice* const one[1] = {vp};
  welsh_block_lti<LFO_AMP, ZERO_A, AMP_FLAT, 1>(one, Ip, fb, lane, *aseg, tile_row);
}
template <bool LFO_AMP, bool ZERO_A>
__device__ __noinline__ void welsh_block_simple_ool(WelshVoice* vp, const WelshInst* Ip, i64 fb, int lane,
                                                    const EnvSeg* aseg, const EnvSeg* fseg, double2* tile_row,
                                                    double* park) {
  welsh_block_simple<LFO_AMP, ZERO_A>(vp, Ip, fb, lane, *aseg, *fseg, tile_row, park);
}

// Each coefficient mode of the fast path is its own out-of-line function: separate register
// allocation and instruction footprint per mode, one call per 256-frame block.
template <int CMODE, bool ALL_ON>
__device__ __noinline__ void welsh_fast_call(WelshVoice* vp, const WelshInst* Ip, i64 fb, int lane, int cls,
                                             EnvSeg aseg, EnvSeg fseg, u64 seed1, u64 seed2, u64 seedl,
                                             double2* tile_row, bool accumulate) {
  WelshVoice st = *vp;
  welsh_block_fast<CMODE, ALL_ON>(st, *Ip, fb, lane, cls, aseg, fseg, seed1, seed2, seedl, tile_row, accumulate, vp->knot);
  __syncwarp();
  if (lane == 0) {  // everything but the knot words (lane 31 has just written those)
    vp->s[0] = st.s[0]; vp->s[1] = st.s[1]; vp->s[2] = st.s[2]; vp->s[3] = st.s[3];
    vp->knot_frame = CMODE == COEF_KNOTS ? fb + kBlockFrames : kNever;
  }
  __syncwarp();
}

// The general (event / pitch-LFO / stage-boundary) variants live behind one out-of-line call that
// works on the voice record in global memory, so the kernel's register allocation is set by the
// fast path; these variants run for a handful of blocks per note.
__device__ __noinline__ void welsh_block_general(int variant, WelshVoice* vp, const WelshInst* Ip,
                                                 const VoiceEvent* __restrict__ ev, int ei, int e_end, i64 fb,
                                                 i64 f_end, int lane, u64 seed1, u64 seed2, u64 seedl,
                                                 double2* tile_row, bool accumulate) {
  WelshVoice st = *vp;
  const WelshInst& I = *Ip;
  switch (variant) {
    case 0: welsh_block<false, false>(st, I, ev, ei, e_end, fb, f_end, lane, seed1, seed2, seedl, tile_row, accumulate); break;
    case 1: welsh_block<false, true>(st, I, ev, ei, e_end, fb, f_end, lane, seed1, seed2, seedl, tile_row, accumulate); break;
    case 2: welsh_block<true, false>(st, I, ev, ei, e_end, fb, f_end, lane, seed1, seed2, seedl, tile_row, accumulate); break;
    default: welsh_block<true, true>(st, I, ev, ei, e_end, fb, f_end, lane, seed1, seed2, seedl, tile_row, accumulate); break;
  }
  st.knot_frame = kNever;  // a carried coefficient knot is only valid between consecutive fast blocks
  __syncwarp();
  if (lane == 0) *vp = st;
  __syncwarp();
}

// grid = number of CtaWork items; block = 32 * W threads;
// dynamic smem = W * kTileStride double2 (tiles) + kParkWords * 32 * W doubles (parking columns).
// __launch_bounds__(.., 2): two CTAs (16 warps) per SM, i.e. at most 128 registers per thread.
template <int W, int MINB, bool SOLO>
__global__ void __launch_bounds__(32 * W, MINB) welsh_kernel(const WelshInst* __restrict__ insts,
                                                           WelshVoice* __restrict__ voices,
                                                           const CtaWork* __restrict__ work,
                                                           const WarpItem* __restrict__ items,
                                                           const VoiceEvent* __restrict__ events,
                                                           const int* __restrict__ ev_off, i64 f0, int nframes,
                                                           const int* __restrict__ idx) {
  extern __shared__ double2 smem_tiles[];
  __shared__ int s_active[W];
  __shared__ WelshInst sI[SOLO ? W : 1];  // instrument records: LDS instead of repeated global loads
  const CtaWork wk = work[idx ? idx[blockIdx.x] : blockIdx.x];  // idx: this chunk's subset of the work list
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool solo = SOLO;  // compile-time: the grouped kernel carries none of the solo plumbing
  const bool mine = !solo || warp < wk.nvoices;   // solo: does this warp have an item?
  WarpItem item;
  item.inst = wk.inst; item.voice = 0; item.out = nullptr;
  if (solo && mine) item = items[wk.voice0 + warp];
  if (!solo) {
    const int* src = reinterpret_cast<const int*>(insts + wk.inst);
    int* dst = reinterpret_cast<int*>(&sI[0]);
    for (int i = threadIdx.x; i < (int)(sizeof(WelshInst) / sizeof(int)); i += 32 * W) dst[i] = src[i];
  } else if (mine) {
    const int* src = reinterpret_cast<const int*>(insts + item.inst);
    int* dst = reinterpret_cast<int*>(&sI[warp]);
    for (int i = lane; i < (int)(sizeof(WelshInst) / sizeof(int)); i += 32) dst[i] = src[i];
  }
  __syncthreads();
  const WelshInst& I = sI[SOLO ? warp : 0];
  double2* tile_row = smem_tiles + warp * kTileStride;
  // per-thread parking column behind the tiles: [kParkWords][32 * W] doubles
  double* park = reinterpret_cast<double*>(smem_tiles + W * kTileStride) + threadIdx.x;
  const i64 f_end = f0 + nframes;
  const bool pitch = mine && I.routing == LFO_PITCH;
  // instrument qualifies for welsh_block_simple (see its preconditions)
  const bool lin_inst = mine && I.s1.kind == 0 && I.s2.kind == 0 && !I.sync &&
                        (I.routing == LFO_NONE || (I.routing == LFO_AMPLITUDE && I.wl == W_SINE));
  const bool simple_inst = lin_inst && I.filter_mode == FILTER_ENVELOPE;
  // instrument qualifies for welsh_block_lti whenever its cutoff rests (see its preconditions)
  const bool lti_inst = lin_inst && I.lti_ok && (I.filter_mode == FILTER_FIXED || I.filter_mode == FILTER_ENVELOPE);
  // voices this warp handles per block: grouped = warp, warp+W, ...; solo = its one item
  const int g_begin = solo ? 0 : warp;
  const int g_end = solo ? (mine ? 1 : 0) : wk.nvoices;
#pragma unroll 1
  for (i64 fb = f0; fb < f_end; fb += kBlockFrames) {
    bool any = false;
    // Voices are taken two at a time (g and g + W): when both rest — no note event for them in this
    // block, both envelopes at their sustain levels for the whole block — the pair is rendered in
    // lockstep by welsh_block_lti<.., 2>.  The test reads two words of each record.
#pragma unroll 1
    for (int g = g_begin; g < g_end; g += 2 * W) {
      const int nv = (!solo && g + W < g_end) ? 2 : 1;
      int vis[2];
      bool rest[2];
      vis[0] = solo ? item.voice : wk.voice0 + g;
      vis[1] = nv == 2 ? vis[0] + W : vis[0];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const longlong2 on_off = *reinterpret_cast<const longlong2*>(&voices[vis[h]].n_on);
        const i64 last = on_off.y < f_end ? on_off.y : f_end;
        rest[h] = lti_inst && fb >= on_off.x + I.steady_after && fb + kBlockFrames <= last;
        if (rest[h]) {  // ... and no note event inside this block (earlier ones are folded into the record)
          int ei = ev_off[vis[h]];
          const int e_end = ev_off[vis[h] + 1];
          while (ei < e_end && events[ei].frame < fb) ++ei;
          rest[h] = !(ei < e_end && events[ei].frame < fb + kBlockFrames);
        }
      }
      if (rest[0] || (nv == 2 && rest[1])) {
        if (!any) {  // the specialised blocks always accumulate into the warp's tile row
          double2* row = tile_row + lane * (kT + 1);
#pragma unroll
          for (int j = 0; j < kT; ++j) row[j] = make_double2(0.0, 0.0);
          any = true;
        }
        EnvSeg none;
        none.q0 = 0.0; none.q1 = 0.0; none.q2 = 0.0; none.w0 = 0.0; none.dw = 0.0;
        if constexpr (SOLO) {  // one voice per warp; shared out-of-line copies, general oscillator form only
          if (I.routing == LFO_NONE) welsh_block_lti_ool<false, false, true>(voices + vis[0], &I, fb, lane, &none, tile_row);
          else welsh_block_lti_ool<true, false, true>(voices + vis[0], &I, fb, lane, &none, tile_row);
        } else {
#define GB_LTI_REST(NV_, ARR_)                                                                                   \
  do {                                                                                                           \
    if (I.osc_flat) {                                                                                            \
      if (I.routing == LFO_NONE) welsh_block_lti<false, true, true, NV_>(ARR_, &I, fb, lane, none, tile_row);    \
      else welsh_block_lti<true, true, true, NV_>(ARR_, &I, fb, lane, none, tile_row);                           \
    } else {                                                                                                     \
      if (I.routing == LFO_NONE) welsh_block_lti<false, false, true, NV_>(ARR_, &I, fb, lane, none, tile_row);   \
      else welsh_block_lti<true, false, true, NV_>(ARR_, &I, fb, lane, none, tile_row);                          \
    }                                                                                                            \
  } while (0)
          if (nv == 2 && rest[0] && rest[1]) {
            WelshVoice* const two[2] = {voices + vis[0], voices + vis[1]};
            GB_LTI_REST(2, two);
            continue;
          }
          WelshVoice* const one[1] = {voices + (rest[0] ? vis[0] : vis[1])};
          GB_LTI_REST(1, one);
#undef GB_LTI_REST
        }
      }
#pragma unroll 1
      for (int h = 0; h < nv; ++h) {
        if (rest[h]) continue;
        const int vi = vis[h];
        WelshVoice* vp = voices + vi;
        NoteWords st;
        st.n_on = vp->n_on; st.n_off = vp->n_off;
        st.la_on = vp->la_on; st.la_off = vp->la_off; st.lf_on = vp->lf_on; st.lf_off = vp->lf_off;
        int ei = ev_off[vi];
        const int e_end = ev_off[vi + 1];
        while (ei < e_end && events[ei].frame < fb) ++ei;  // already folded into the record by earlier blocks
        const bool idle = fb >= st.n_off + I.amp.nr;
        const bool ev_here = ei < e_end && events[ei].frame < fb + kBlockFrames;
        if (idle && !ev_here) continue;
        // noise seeds are only needed off the specialised path
        const int local = vi - I.voice0;
        auto seed_of = [&](u64 salt) { return splitmix64(((u64)(unsigned)I.uid << 32) ^ salt); };
        bool fast = false;
        EnvSeg aseg, fseg;
        int cls = 0;
        if (!pitch && !ev_here) {
          cls = welsh_lane_class(st, I, fb + (i64)lane * kT, f_end, aseg, fseg);
          fast = __all_sync(0xffffffffu, cls != 0);
        }
        bool lti = false;
        if (fast && lti_inst) {
          // every lane sounding, and the cutoff at rest: fixed filter, or the filter envelope at its sustain level
          const bool rest = cls == 2 && (I.filter_mode == FILTER_FIXED ||
                                         (fseg.q1 == 0.0 && fseg.q2 == 0.0 && fseg.q0 == I.filt.sustain));
          lti = __all_sync(0xffffffffu, rest);
        }
        if ((lti || fast) && !any) {  // the specialised blocks always accumulate into the warp's tile row
          double2* row = tile_row + lane * (kT + 1);
  #pragma unroll
          for (int j = 0; j < kT; ++j) row[j] = make_double2(0.0, 0.0);
        }
        if (lti) {
          if constexpr (SOLO) {
            if (I.routing == LFO_NONE) welsh_block_lti_ool<false, false, false>(vp, &I, fb, lane, &aseg, tile_row);
            else welsh_block_lti_ool<true, false, false>(vp, &I, fb, lane, &aseg, tile_row);
          } else {
            WelshVoice* const one[1] = {vp};
            if (I.osc_flat) {
              if (I.routing == LFO_NONE) welsh_block_lti<false, true, false, 1>(one, &I, fb, lane, aseg, tile_row);
              else welsh_block_lti<true, true, false, 1>(one, &I, fb, lane, aseg, tile_row);
            } else {
              if (I.routing == LFO_NONE) welsh_block_lti<false, false, false, 1>(one, &I, fb, lane, aseg, tile_row);
              else welsh_block_lti<true, false, false, 1>(one, &I, fb, lane, aseg, tile_row);
            }
          }
        } else if (fast) {
          // the fast path changes only the filter state and the carried knot of the voice record
          const WelshInst* Ip = &I;  // shared-memory copy
          if (I.filter_mode == FILTER_FIXED) {
            if (__all_sync(0xffffffffu, cls == 2))
              welsh_fast_call<COEF_FIXED, true>(vp, Ip, fb, lane, cls, aseg, fseg, seed_of((u64)(2 * local)), seed_of((u64)(2 * local + 1)), seed_of(0x4C464F00ull ^ (u64)local), tile_row, true);
            else
              welsh_fast_call<COEF_FIXED, false>(vp, Ip, fb, lane, cls, aseg, fseg, seed_of((u64)(2 * local)), seed_of((u64)(2 * local + 1)), seed_of(0x4C464F00ull ^ (u64)local), tile_row, any);
          } else {
            // knots need every lane sounding and the cutoff moving slowly enough over the lane's frames
            bool smooth = false;
            if (I.filter_mode == FILTER_ENVELOPE && cls == 2) {
              const double w8 = fma((double)kT, fseg.dw, fseg.w0);
              const double r0 = fabs(fma(2.0 * fseg.q2, fseg.w0, fseg.q1)), r8 = fabs(fma(2.0 * fseg.q2, w8, fseg.q1));
              smooth = fabs(I.cut_b * fseg.dw) * fmax(r0, r8) <= I.knot_max_rate && I.knot_max_rate > 0.0;
            }
            if (__all_sync(0xffffffffu, smooth)) {
              if (SOLO && simple_inst) {
                if (I.routing == LFO_NONE) welsh_block_simple_ool<false, false>(vp, Ip, fb, lane, &aseg, &fseg, tile_row, park);
                else welsh_block_simple_ool<true, false>(vp, Ip, fb, lane, &aseg, &fseg, tile_row, park);
              } else if (!SOLO && simple_inst && I.osc_flat) {
                if (I.routing == LFO_NONE) welsh_block_simple<false, true>(vp, Ip, fb, lane, aseg, fseg, tile_row, park);
                else welsh_block_simple<true, true>(vp, Ip, fb, lane, aseg, fseg, tile_row, park);
              } else if (!SOLO && simple_inst) {
                if (I.routing == LFO_NONE) welsh_block_simple<false, false>(vp, Ip, fb, lane, aseg, fseg, tile_row, park);
                else welsh_block_simple<true, false>(vp, Ip, fb, lane, aseg, fseg, tile_row, park);
              } else
                welsh_fast_call<COEF_KNOTS, true>(vp, Ip, fb, lane, cls, aseg, fseg, seed_of((u64)(2 * local)), seed_of((u64)(2 * local + 1)), seed_of(0x4C464F00ull ^ (u64)local), tile_row, any);
            } else {
              welsh_fast_call<COEF_EXACT, false>(vp, Ip, fb, lane, cls, aseg, fseg, seed_of((u64)(2 * local)), seed_of((u64)(2 * local + 1)), seed_of(0x4C464F00ull ^ (u64)local), tile_row, any);
            }
          }
        } else {
          welsh_block_general((pitch ? 2 : 0) + (ev_here ? 1 : 0), vp, insts + item.inst, events, ei, e_end, fb,
                              f_end, lane, seed_of((u64)(2 * local)), seed_of((u64)(2 * local + 1)), seed_of(0x4C464F00ull ^ (u64)local), tile_row, any);
        }
        any = true;
      }
    }
    if (solo) {
      __syncwarp();
      if (mine) warp_store_row(tile_row, any, item.out, fb, f0, f_end, lane);
      __syncwarp();
    } else {
      if (lane == 0) s_active[warp] = any ? 1 : 0;
      __syncthreads();
      cta_reduce_store<W>(smem_tiles, s_active, wk.out, fb, f0, f_end);
      __syncthreads();
    }
  }
}

// Oscillator phases of a lane inside a 256-frame block whose base phases (P1, P2) are those of the frame
// before the block.  With hard sync oscillator 2 restarts on every frame at which oscillator 1 wraps
// (welsh_block / welsh_phases_at): a lane's start phase follows from j = floor(p1 / d1), the frames since the
// last wrap (if that is inside the block so far, p2 = j d2), taken from a double-precision estimate with a
// one-step correction instead of a 64-bit division.
template <bool SYNC>
__device__ __forceinline__ void osc_start_at(u64 P1, u64 P2, u64 d1, u64 d2, u64 k, u64& p1, u64& p2);
template <bool SYNC>
__device__ __forceinline__ void osc_lane_start(u64 P1, u64 P2, u64 d1, u64 d2, int lane, u64& p1, u64& p2) {
  osc_start_at<SYNC>(P1, P2, d1, d2, (u64)(lane * kT), p1, p2);
}
// ... the same k frames after the base frame (k = 0: the base phases themselves)
template <bool SYNC>
__device__ __forceinline__ void osc_start_at(u64 P1, u64 P2, u64 d1, u64 d2, u64 k, u64& p1, u64& p2) {
  p1 = P1 + k * d1;
  p2 = P2 + k * d2;
  if (SYNC && d1 != 0) {
    u64 j = (u64)__double2ull_rz(__ull2double_rn(p1) * rcp_ranged(__ull2double_rn(d1)));
    const u64 r = p1 - j * d1;
    if ((i64)r < 0) --j;          // estimate one too high (d1 < 2^63: a negative remainder is unambiguous)
    else if (r >= d1) ++j;        // ... or one too low
    if (j < k) p2 = j * d2;
  }
}
template <bool SYNC>
__device__ __forceinline__ void osc_advance(u64& p1, u64& p2, u64 d1, u64 d2) {
  p1 += d1;
  if (SYNC) p2 = p1 < d1 ? 0 : p2 + d2;
  else p2 += d2;
}

// ---- the resting-voice kernel ----------------------------------------------------------------------
// The host knows every voice's note frames, so per chunk it sorts the grouped CTAs: those whose voices
// all rest for the WHOLE chunk (note held since before the chunk, both envelopes at their sustain
// levels, no note event in the chunk, chunk a multiple of kBlockFrames) come here, the others go to
// welsh_kernel.  Here nothing is decided per block: the voice state lives in shared memory for the
// duration of the launch (phases advance by integer adds, the LFO by a rotation, the filter state
// straight from the scan), two voices run in lockstep per warp, and the only global traffic of the
// block loop is the CTA's 16-byte-per-frame output.
struct alignas(16) RestState {
  u64 p1, p2;     // oscillator phases at the frame before the current block
  u64 d1, d2;
  double s[4];    // filter state at the start of the current block
  double ls, lc;  // depth * level * (sin, cos) of the LFO angle at the first frame of the current block
};

// Exclusive form of lti_scan_states: the lanes' zero-state end vectors are shifted up by one lane with
// the span's entry state entering at lane 0, so the inclusive scan yields each lane's entry state
// directly.  NV voices of one instrument are scanned together: they share the span maps, so each step
// loads its matrix once.  sec = 0 / 1 picks the cached state words (s[2 sec], s[2 sec + 1]): lane 0 reads
// them as the span's entry state, lane 31 replaces them by the state after the span.
template <int NV, typename State>
__device__ __forceinline__ void lti_scan_entry_t(const double (&v0)[NV], const double (&v1)[NV], const double (*mp)[4],
                                                 int lane, State* const* rs, int sec, double (&e0)[NV],
                                                 double (&e1)[NV]) {
  double u0[NV], u1[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const double2 st = *reinterpret_cast<const double2*>(&rs[v]->s[2 * sec]);  // only lane 0 uses it
    u0[v] = shfl_up_f64(v0[v], 1);
    u1[v] = shfl_up_f64(v1[v], 1);
    u0[v] = lane == 0 ? st.x : u0[v];
    u1[v] = lane == 0 ? st.y : u1[v];
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int d = 1 << k;
    const double* m = mp[lane >= d ? k : 5];
    const double2 r0 = *reinterpret_cast<const double2*>(m), r1 = *reinterpret_cast<const double2*>(m + 2);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const double p0 = shfl_up_f64(u0[v], d), p1 = shfl_up_f64(u1[v], d);
      affine_vec_step(u0[v], u1[v], r0.x, r0.y, r1.x, r1.y, p0, p1);
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) { e0[v] = u0[v]; e1[v] = u1[v]; }
  __syncwarp();      // lane 0 has read the old state
  if (lane == 31) {  // state after the whole span: this lane's map applied to its entry state
    const double2 r0 = *reinterpret_cast<const double2*>(mp[0]), r1 = *reinterpret_cast<const double2*>(mp[0] + 2);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      double x0 = v0[v], x1 = v1[v];
      affine_vec_step(x0, x1, r0.x, r0.y, r1.x, r1.y, u0[v], u1[v]);
      *reinterpret_cast<double2*>(&rs[v]->s[2 * sec]) = make_double2(x0, x1);
    }
  }
}
template <int NV>
__device__ __forceinline__ void lti_scan_entry(const double (&v0)[NV], const double (&v1)[NV], const double (*mp)[4],
                                               int lane, RestState* const (&rs)[NV], int sec, double (&e0)[NV],
                                               double (&e1)[NV]) {
  lti_scan_entry_t<NV, RestState>(v0, v1, mp, lane, rs, sec, e0, e1);
}

template <bool LFO_AMP, bool ZERO_A, int NV, bool ACC, bool SYNC = false>
__device__ __forceinline__ void welsh_rest_block(RestState* const (&rs)[NV], const WelshInst& I, int lane,
                                                 double2* tile_row) {
  // Cached state is read where it is needed and written back as soon as its new value exists, so that
  // phases, filter states and the LFO phasor do not occupy registers across the whole block.
  const LtiTable& L = I.lti;
  double yp[NV][kT];
  double ps0[NV], ps1[NV];
  // ---- pass 1: oscillators + section 1 from a zero state; phases advance by one block ----
  {
    u64 p1[NV], p2[NV], d1[NV], d2[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const ulonglong2 pp = *reinterpret_cast<const ulonglong2*>(&rs[v]->p1);
      const ulonglong2 dd = *reinterpret_cast<const ulonglong2*>(&rs[v]->d1);
      d1[v] = dd.x; d2[v] = dd.y;
      osc_lane_start<SYNC>(pp.x, pp.y, dd.x, dd.y, lane, p1[v], p2[v]);
      ps0[v] = 0.0; ps1[v] = 0.0;
    }
    __syncwarp();  // every lane has read the block's base phases
    if (!SYNC && lane == 0) {
#pragma unroll
      for (int v = 0; v < NV; ++v)
        *reinterpret_cast<ulonglong2*>(&rs[v]->p1) =
            make_ulonglong2(p1[v] + (u64)kBlockFrames * d1[v], p2[v] + (u64)kBlockFrames * d2[v]);
    }
    // The zero-state run of section 1 is linear in its input, so it runs pre-scaled by both sections' b0:
    // its outputs are then section 2's b0*x (up to the entry-state term, whose rows g1b carry the same
    // factor) and only the lane's end vector is scaled back for the scan: 2 multiplies per lane instead
    // of 2 per frame.
    const OscMix o1 = I.m1bb, o2 = I.m2bb;
    const u64 t1 = I.s1.thresh, t2 = I.s2.thresh;
    const double a1 = L.c1.a1, a2 = L.c1.a2;
#pragma unroll
    for (int j = 0; j < kT; ++j) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        osc_advance<SYNC>(p1[v], p2[v], d1[v], d2[v]);
        yp[v][j] = lp_step_bx(osc_mix_eval<ZERO_A>(o1, t1, p1[v], o2, t2, p2[v]), a1, a2, ps0[v], ps1[v]);
      }
    }
    if (SYNC && lane == 31) {  // the last lane ends on the block's last frame: the next block's base phases
#pragma unroll
      for (int v = 0; v < NV; ++v) *reinterpret_cast<ulonglong2*>(&rs[v]->p1) = make_ulonglong2(p1[v], p2[v]);
    }
    const double inv = L.inv_b0_2;
#pragma unroll
    for (int v = 0; v < NV; ++v) { ps0[v] *= inv; ps1[v] *= inv; }
  }
  double e0[NV], e1[NV];
  lti_scan_entry<NV>(ps0, ps1, L.mp1, lane, rs, 0, e0, e1);
  // ---- pass 2: section 2 on the fixed-up section-1 output ----
  {
    const double a1 = L.c2.a1, a2 = L.c2.a2;
#pragma unroll
    for (int v = 0; v < NV; ++v) { ps0[v] = 0.0; ps1[v] = 0.0; }
#pragma unroll
    for (int j = 0; j < kT; ++j) {
      const double2 g = *reinterpret_cast<const double2*>(L.g1b[j]);
#pragma unroll
      for (int v = 0; v < NV; ++v)
        yp[v][j] = lp_step_bx(fma(g.y, e1[v], fma(g.x, e0[v], yp[v][j])), a1, a2, ps0[v], ps1[v]);
    }
  }
  lti_scan_entry<NV>(ps0, ps1, L.mp2, lane, rs, 1, e0, e1);
  // ---- LFO phasor of this lane; the cached one advances by one block ----
  double lsd[NV], lcd[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    lsd[v] = 0.0; lcd[v] = 0.0;
    if (LFO_AMP) {
      const double2 ph = *reinterpret_cast<const double2*>(&rs[v]->ls);
      const double2 r = I.lane_rot[lane];
      lsd[v] = fma(ph.x, r.x, ph.y * r.y);
      lcd[v] = fma(ph.y, r.x, -(ph.x * r.y));
    }
  }
  if (LFO_AMP) {
    __syncwarp();  // every lane has read the block's phasor
    if (lane == 0) {
      const double2 r = I.block_rot;  // lane 0's (lsd, lcd) is the block phasor itself (lane_rot[0] = (1, 0))
#pragma unroll
      for (int v = 0; v < NV; ++v)
        *reinterpret_cast<double2*>(&rs[v]->ls) =
            make_double2(fma(lsd[v], r.x, lcd[v] * r.y), fma(lcd[v], r.x, -(lsd[v] * r.y)));
    }
  }
  // ---- amplitude, DCA, into the warp's tile row (the voices of a pair are summed first) ----
  const double arest = I.amp_rest;
  const double gl = I.gl, gr = I.gr;
  double2* row = tile_row + lane * (kT + 1);
#pragma unroll
  for (int j = 0; j < kT; ++j) {
    const double2 g = *reinterpret_cast<const double2*>(L.g2[j]);
    double2 rot = make_double2(0.0, 0.0);
    if (LFO_AMP) rot = I.lfo_rot[j];
    double m = 0.0;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const double amp = LFO_AMP ? fma(lsd[v], rot.x, fma(lcd[v], rot.y, arest)) : arest;
      const double y = fma(g.y, e1[v], fma(g.x, e0[v], yp[v][j]));
      m = v == 0 ? y * amp : fma(y, amp, m);
    }
    if (ACC) {
      const double2 p = row[j];
      row[j] = make_double2(fma(m, gl, p.x), fma(m, gr, p.y));
    } else {
      row[j] = make_double2(m * gl, m * gr);
    }
  }
  __syncwarp();
}

// grid = number of resting CTAs of this variant; block = 32 * W threads;
// dynamic smem = W * kTileStride double2 (tiles) + max_voices RestState.  nframes is a multiple of kBlockFrames.
// NVW = 4: four voices in lockstep per warp, one CTA per SM (255 registers): CTAs of 32 voices, half the CTA
// partials for the mixdown at the same number of voices in flight per SM.
template <int W, bool LFO_AMP, bool ZERO_A, bool SYNC = false, int NVW = 2>
__global__ void __launch_bounds__(32 * W, NVW == 4 ? 1 : 2) welsh_rest_kernel(const WelshInst* __restrict__ insts,
                                                             WelshVoice* __restrict__ voices,
                                                             const CtaWork* __restrict__ work,
                                                             const int* __restrict__ idx, i64 f0, int nframes) {
  extern __shared__ double2 smem_tiles[];
  __shared__ int s_active[W];
  __shared__ WelshInst sI;
  const CtaWork wk = work[idx[blockIdx.x]];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const int* src = reinterpret_cast<const int*>(insts + wk.inst);
    int* dst = reinterpret_cast<int*>(&sI);
    for (int i = threadIdx.x; i < (int)(sizeof(WelshInst) / sizeof(int)); i += 32 * W) dst[i] = src[i];
  }
  __syncthreads();
  const WelshInst& I = sI;
  RestState* cache = reinterpret_cast<RestState*>(smem_tiles + W * kTileStride);
  for (int t = threadIdx.x; t < wk.nvoices; t += 32 * W) {
    const WelshVoice* vp = voices + wk.voice0 + t;
    RestState r;
    const u64 k = (u64)(f0 - 1 - vp->anchor);
    r.d1 = vp->d1; r.d2 = vp->d2;
    r.p1 = vp->p1 + k * r.d1; r.p2 = vp->p2 + k * r.d2;
    if (SYNC) r.p2 = welsh_phases_at(*vp, I, f0 - 1).p2;  // oscillator 2 restarted at oscillator 1's last wrap
    r.s[0] = vp->s[0]; r.s[1] = vp->s[1]; r.s[2] = vp->s[2]; r.s[3] = vp->s[3];
    r.ls = 0.0; r.lc = 0.0;
    if (LFO_AMP) {
      double ls, lc;
      sincos_phase(vp->pl + (k + 1) * I.lfo_dq, &ls, &lc);
      const double dl = I.depth * I.amp_rest;
      r.ls = ls * dl; r.lc = lc * dl;
    }
    cache[t] = r;
  }
  if (lane == 0) s_active[warp] = warp < wk.nvoices ? 1 : 0;
  __syncthreads();
  double2* tile_row = smem_tiles + warp * kTileStride;
  const i64 f_end = f0 + nframes;
#pragma unroll 1
  for (i64 fb = f0; fb < f_end; fb += kBlockFrames) {
    bool first = true;
#pragma unroll 1
    for (int g = warp; g < wk.nvoices; g += 2 * W) {
      bool took4 = false;
      if constexpr (NVW == 4) {
        if (g + 3 * W < wk.nvoices) {
          RestState* const four[4] = {cache + g, cache + g + W, cache + g + 2 * W, cache + g + 3 * W};
          if (first) welsh_rest_block<LFO_AMP, ZERO_A, 4, false, SYNC>(four, I, lane, tile_row);
          else welsh_rest_block<LFO_AMP, ZERO_A, 4, true, SYNC>(four, I, lane, tile_row);
          g += 2 * W;
          took4 = true;
        }
      }
      if (took4) {
      } else if (g + W < wk.nvoices) {
        RestState* const two[2] = {cache + g, cache + g + W};
        if (first) welsh_rest_block<LFO_AMP, ZERO_A, 2, false, SYNC>(two, I, lane, tile_row);
        else welsh_rest_block<LFO_AMP, ZERO_A, 2, true, SYNC>(two, I, lane, tile_row);
      } else {
        RestState* const one[1] = {cache + g};
        if (first) welsh_rest_block<LFO_AMP, ZERO_A, 1, false, SYNC>(one, I, lane, tile_row);
        else welsh_rest_block<LFO_AMP, ZERO_A, 1, true, SYNC>(one, I, lane, tile_row);
      }
      first = false;
    }
    __syncthreads();
    cta_reduce_store<W>(smem_tiles, s_active, wk.out, fb, f0, f_end);
    __syncthreads();
  }
  for (int t = threadIdx.x; t < wk.nvoices; t += 32 * W) {
    WelshVoice* vp = voices + wk.voice0 + t;
    vp->s[0] = cache[t].s[0]; vp->s[1] = cache[t].s[1]; vp->s[2] = cache[t].s[2]; vp->s[3] = cache[t].s[3];
    vp->knot_frame = kNever;
  }
}

// ---- the resting-voice kernel over voice ranges ------------------------------------------------------
// welsh_rest_kernel's CTAs belong to one instrument each (16 of its voices), so 4096 voices are 256 CTAs on the
// 296 CTA slots of 148 SMs: 108 SMs carry 32 voices, 40 carry 16, and the launch lasts as long as the full ones.
// When EVERY grouped CTA of an engine rests for a chunk (and all instruments mix into one consumer that sums
// the CTA partials itself), the chunk runs here instead: the engine's voices, in voice-table order, are cut
// into ranges of 2 W voices regardless of instrument boundaries (W = 7: 293 CTAs for 4096 voices, 28 voices on
// every SM).  A warp holds two consecutive voices — of one instrument, since instruments have even voice counts
// — and takes its instrument record from one of the CTA's two copies; the tile rows are already panned per
// warp, so the CTA sum mixes both instruments into the range's partial buffer.
// W = 8 spreads the same 14 voices over 8 warps (6 pairs + 2 single voices, `single0`), four warps of each CTA pair
// on every sub-partition: measured 0.943 ms per launch against 0.949 ms with 7 warps of two voices
// (profiles/r2f_vr_warps_ab.txt).
// Tried and dropped (profiles/r2f_vr_ring_experiment.txt): no CTA barrier at all — every warp writes its block's
// row into a 4-deep ring, counts itself in, and the LAST warp to arrive for a block reduces it — so that the warps
// drift out of phase: 1.11 ms per launch against 0.943 ms.  The warps of a CTA in the same phase share the
// instruction cache (the block is 3400 instructions); out of phase they do not.
struct alignas(16) VrWork {
  int inst_a, inst_b;  // instrument of the first voices / of the voices from `split` on (== inst_a if the range has one instrument)
  int split;           // voices of inst_a in this range (even)
  int voice0, nvoices; // global voice range
  int single0;         // warps single0 and single0 + 1 hold ONE voice each, the others two (>= W: every warp holds two)
  double2* out;
};

template <int W, bool LFO_AMP, bool ZERO_A>
__global__ void __launch_bounds__(32 * W, 2) welsh_rest_vr_kernel(const WelshInst* __restrict__ insts,
                                                                WelshVoice* __restrict__ voices,
                                                                const VrWork* __restrict__ work, i64 f0, int nframes) {
  extern __shared__ double2 smem_tiles[];
  __shared__ int s_active[W];
  __shared__ WelshInst sI[2];
  const VrWork wk = work[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const int* a = reinterpret_cast<const int*>(insts + wk.inst_a);
    const int* b = reinterpret_cast<const int*>(insts + wk.inst_b);
    int* dst = reinterpret_cast<int*>(&sI[0]);
    constexpr int kWords = (int)(sizeof(WelshInst) / sizeof(int));
    for (int i = threadIdx.x; i < 2 * kWords; i += 32 * W) dst[i] = i < kWords ? a[i] : b[i - kWords];
  }
  __syncthreads();
  RestState* cache = reinterpret_cast<RestState*>(smem_tiles + W * kTileStride);
  for (int t = threadIdx.x; t < wk.nvoices; t += 32 * W) {
    const WelshInst& I = sI[t >= wk.split ? 1 : 0];
    const WelshVoice* vp = voices + wk.voice0 + t;
    RestState r;
    const u64 k = (u64)(f0 - 1 - vp->anchor);
    r.d1 = vp->d1; r.d2 = vp->d2;
    r.p1 = vp->p1 + k * r.d1; r.p2 = vp->p2 + k * r.d2;
    r.s[0] = vp->s[0]; r.s[1] = vp->s[1]; r.s[2] = vp->s[2]; r.s[3] = vp->s[3];
    r.ls = 0.0; r.lc = 0.0;
    if (LFO_AMP) {
      double ls, lc;
      sincos_phase(vp->pl + (k + 1) * I.lfo_dq, &ls, &lc);
      const double dl = I.depth * I.amp_rest;
      r.ls = ls * dl; r.lc = lc * dl;
    }
    cache[t] = r;
  }
  // first voice of this warp: two per warp, except the two single-voice warps
  const int g = 2 * warp - min(max(warp - wk.single0, 0), 2);
  const bool pair = !(warp == wk.single0 || warp == wk.single0 + 1);
  if (lane == 0) s_active[warp] = g < wk.nvoices ? 1 : 0;
  __syncthreads();
  const WelshInst& I = sI[g >= wk.split ? 1 : 0];
  double2* tile_row = smem_tiles + warp * kTileStride;
  const i64 f_end = f0 + nframes;
#pragma unroll 1
  for (i64 fb = f0; fb < f_end; fb += kBlockFrames) {
    if (pair && g + 1 < wk.nvoices) {
      RestState* const two[2] = {cache + g, cache + g + 1};
      welsh_rest_block<LFO_AMP, ZERO_A, 2, false, false>(two, I, lane, tile_row);
    } else if (g < wk.nvoices) {
      RestState* const one[1] = {cache + g};
      welsh_rest_block<LFO_AMP, ZERO_A, 1, false, false>(one, I, lane, tile_row);
    }
    __syncthreads();
    cta_reduce_store<W>(smem_tiles, s_active, wk.out, fb, f0, f_end);
    __syncthreads();
  }
  for (int t = threadIdx.x; t < wk.nvoices; t += 32 * W) {
    WelshVoice* vp = voices + wk.voice0 + t;
    vp->s[0] = cache[t].s[0]; vp->s[1] = cache[t].s[1]; vp->s[2] = cache[t].s[2]; vp->s[3] = cache[t].s[3];
    vp->knot_frame = kNever;
  }
}

// ---- voice ranges with 16 frames per lane -------------------------------------------------------------
// A block of welsh_rest_block costs its warp about 1375 issue cycles per sub-partition plus a share of ~1950 cycles
// per block in which nobody issues (the two scans' shuffle -> FMA chains, the CTA barriers).  With
// 16 frames per lane a block is 512 frames: the scans, their matrix loads and the barriers come once per 512 frames
// instead of once per 256.  The zero-state outputs of the two passes (16 per voice) do not stay in registers: a
// pair's two values of a frame are exactly one 16-byte word, and that word lives in the warp's tile row — the
// slot the frame's panned output is written to at the end, when the pair's value is dead.  Tables for 16-frame
// lanes (rows g, span maps A^(16 2^k), LFO rotations) come from a per-instrument Rest16Table built at gb_finalize.
constexpr int kT16 = 16;
constexpr int kBlock16 = 32 * kT16;
constexpr int kTile16Stride = kBlock16 + kBlock16 / kT16;  // padded (16-byte units): lane stride kT16 + 1

struct alignas(16) Rest16Table {
  double g1b[kT16][2];     // section 1's entry-state rows, scaled by section 2's b0
  double g2[kT16][2];      // section 2's entry-state rows
  double mp1[6][4];        // (A1^16)^(2^k), k = 0..4; [5] = 0
  double mp2[6][4];
  double2 lfo_rot[kT16];   // (cos, sin) of j LFO steps
  double2 lane_rot[32];    // ... of 16 l LFO steps
  double2 block_rot;       // ... of 512 LFO steps
};

// Symmetric piecewise-constant waveforms (square, pulse: +c below the threshold, -c above): the waveform value is
// the constant with its SIGN BIT picked by the compare — one 32-bit select instead of a two-word double select —
// and for a square (threshold = half a turn) the phase's top bit IS that sign: no compare at all.
//   OSC = 1: both oscillators symmetric;  OSC = 2: ... and oscillator 2 is a square.  Same bits as osc_mix_eval.
template <int OSC>
__device__ __forceinline__ double osc_sym_eval(double c1, u64 t1, u64 p1, double c2, u64 t2, u64 p2) {
  const int c1h = __double2hiint(c1), c2h = __double2hiint(c2);
  const int h1 = p1 < t1 ? c1h : c1h ^ (int)0x80000000;
  int h2;
  if (OSC == 2) h2 = c2h ^ ((int)(p2 >> 32) & (int)0x80000000);
  else h2 = p2 < t2 ? c2h : c2h ^ (int)0x80000000;
  return __hiloint2double(h1, __double2loint(c1)) + __hiloint2double(h2, __double2loint(c2));
}

template <bool LFO_AMP, bool ZERO_A, int NV, int OSC = 0>
__device__ __forceinline__ void welsh_rest_block16(RestState* const (&rs)[NV], const WelshInst& I, const Rest16Table& R,
                                                   int lane, double2* tile_row) {
  static_assert(OSC == 0 || ZERO_A, "the sign-bit form is for piecewise-constant waveforms");
  const LtiTable& L = I.lti;
  double2* row = tile_row + lane * (kT16 + 1);
  double ps0[NV], ps1[NV];
  // ---- pass 1: oscillators + section 1 from a zero state (pre-scaled by both sections' b0, as welsh_rest_block) ----
  {
    u64 p1[NV], p2[NV], d1[NV], d2[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const ulonglong2 pp = *reinterpret_cast<const ulonglong2*>(&rs[v]->p1);
      const ulonglong2 dd = *reinterpret_cast<const ulonglong2*>(&rs[v]->d1);
      d1[v] = dd.x; d2[v] = dd.y;
      p1[v] = pp.x + (u64)(lane * kT16) * dd.x;
      p2[v] = pp.y + (u64)(lane * kT16) * dd.y;
      ps0[v] = 0.0; ps1[v] = 0.0;
    }
    __syncwarp();  // every lane has read the block's base phases
    if (lane == 0) {
#pragma unroll
      for (int v = 0; v < NV; ++v)
        *reinterpret_cast<ulonglong2*>(&rs[v]->p1) =
            make_ulonglong2(p1[v] + (u64)kBlock16 * d1[v], p2[v] + (u64)kBlock16 * d2[v]);
    }
    const OscMix o1 = I.m1bb, o2 = I.m2bb;
    const u64 t1 = I.s1.thresh, t2 = I.s2.thresh;
    const double a1 = L.c1.a1, a2 = L.c1.a2;
#pragma unroll
    for (int j = 0; j < kT16; ++j) {
      double y[2] = {0.0, 0.0};
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        osc_advance<false>(p1[v], p2[v], d1[v], d2[v]);
        const double x = OSC ? osc_sym_eval<OSC>(o1.b_lo, t1, p1[v], o2.b_lo, t2, p2[v])
                             : osc_mix_eval<ZERO_A>(o1, t1, p1[v], o2, t2, p2[v]);
        y[v] = lp_step_bx(x, a1, a2, ps0[v], ps1[v]);
      }
      row[j] = make_double2(y[0], y[1]);
    }
    const double inv = L.inv_b0_2;
#pragma unroll
    for (int v = 0; v < NV; ++v) { ps0[v] *= inv; ps1[v] *= inv; }
  }
  double e0[NV], e1[NV];
  lti_scan_entry_t<NV, RestState>(ps0, ps1, R.mp1, lane, rs, 0, e0, e1);
  // ---- pass 2: section 2 on the fixed-up section-1 output ----
  {
    const double a1 = L.c2.a1, a2 = L.c2.a2;
#pragma unroll
    for (int v = 0; v < NV; ++v) { ps0[v] = 0.0; ps1[v] = 0.0; }
    // the row words and table rows of the next two frames are in flight while a frame is computed
    double2 yq[3], gq[3];
    yq[0] = row[0]; gq[0] = *reinterpret_cast<const double2*>(R.g1b[0]);
    yq[1] = row[1]; gq[1] = *reinterpret_cast<const double2*>(R.g1b[1]);
#pragma unroll
    for (int j = 0; j < kT16; ++j) {
      if (j + 2 < kT16) { yq[(j + 2) % 3] = row[j + 2]; gq[(j + 2) % 3] = *reinterpret_cast<const double2*>(R.g1b[j + 2]); }
      const double2 g = gq[j % 3];
      const double2 yv = yq[j % 3];
      double y[2] = {yv.x, yv.y};
#pragma unroll
      for (int v = 0; v < NV; ++v)
        y[v] = lp_step_bx(fma(g.y, e1[v], fma(g.x, e0[v], y[v])), a1, a2, ps0[v], ps1[v]);
      row[j] = make_double2(y[0], y[1]);
    }
  }
  lti_scan_entry_t<NV, RestState>(ps0, ps1, R.mp2, lane, rs, 1, e0, e1);
  // ---- LFO phasor of this lane; the cached one advances by one block ----
  double lsd[NV], lcd[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    lsd[v] = 0.0; lcd[v] = 0.0;
    if (LFO_AMP) {
      const double2 ph = *reinterpret_cast<const double2*>(&rs[v]->ls);
      const double2 r = R.lane_rot[lane];
      lsd[v] = fma(ph.x, r.x, ph.y * r.y);
      lcd[v] = fma(ph.y, r.x, -(ph.x * r.y));
    }
  }
  if (LFO_AMP) {
    __syncwarp();  // every lane has read the block's phasor
    if (lane == 0) {
      const double2 r = R.block_rot;
#pragma unroll
      for (int v = 0; v < NV; ++v)
        *reinterpret_cast<double2*>(&rs[v]->ls) =
            make_double2(fma(lsd[v], r.x, lcd[v] * r.y), fma(lcd[v], r.x, -(lsd[v] * r.y)));
    }
  }
  // ---- amplitude, DCA: the pair's word of each frame becomes the frame's panned output ----
  const double arest = I.amp_rest;
  const double gl = I.gl, gr = I.gr;
  double2 yq[3], gq[3], rq[3];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    yq[j] = row[j]; gq[j] = *reinterpret_cast<const double2*>(R.g2[j]);
    rq[j] = LFO_AMP ? R.lfo_rot[j] : make_double2(0.0, 0.0);
  }
#pragma unroll
  for (int j = 0; j < kT16; ++j) {
    if (j + 2 < kT16) {
      yq[(j + 2) % 3] = row[j + 2]; gq[(j + 2) % 3] = *reinterpret_cast<const double2*>(R.g2[j + 2]);
      rq[(j + 2) % 3] = LFO_AMP ? R.lfo_rot[j + 2] : make_double2(0.0, 0.0);
    }
    const double2 g = gq[j % 3];
    const double2 rot = rq[j % 3];
    const double2 yv = yq[j % 3];
    const double yy[2] = {yv.x, yv.y};
    double m = 0.0;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const double amp = LFO_AMP ? fma(lsd[v], rot.x, fma(lcd[v], rot.y, arest)) : arest;
      const double y = fma(g.y, e1[v], fma(g.x, e0[v], yy[v]));
      m = v == 0 ? y * amp : fma(y, amp, m);
    }
    row[j] = make_double2(m * gl, m * gr);
  }
  __syncwarp();
}

// All W rows are summed unconditionally, in warp order: the rows of warps without voices (the last, partial range)
// are zeroed once at kernel start, so the loads do not hang on a per-warp flag (in cta_reduce_store that flag load
// -> compare -> predicated row load chain was 9 % of this kernel's stall samples, r2f_rest_vr16_hot_spots.txt).
template <int W>
__device__ __forceinline__ void cta_reduce_store16(const double2* tiles, double2* out, i64 fb, i64 f0) {
  for (int t = threadIdx.x; t < kBlock16; t += 32 * W) {
    const int u = t + (t >> 4);
    double2 v[W];
#pragma unroll
    for (int w = 0; w < W; ++w) v[w] = tiles[w * kTile16Stride + u];
    double l = 0.0, r = 0.0;
#pragma unroll
    for (int w = 0; w < W; ++w) { l += v[w].x; r += v[w].y; }
    out[fb + t - f0] = make_double2(l, r);
  }
}

// grid = ranges; block = 32 * W; dynamic smem = W * kTile16Stride double2 + 2 W RestState; nframes a multiple of 512
template <int W, bool LFO_AMP, bool ZERO_A, int OSC = 0>
__global__ void __launch_bounds__(32 * W, 2) welsh_rest_vr16_kernel(const WelshInst* __restrict__ insts,
                                                                  const Rest16Table* __restrict__ tabs,
                                                                  WelshVoice* __restrict__ voices,
                                                                  const VrWork* __restrict__ work, i64 f0, int nframes) {
  extern __shared__ double2 smem_tiles[];
  __shared__ WelshInst sI[2];
  __shared__ Rest16Table sR[2];
  const VrWork wk = work[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const int* a = reinterpret_cast<const int*>(insts + wk.inst_a);
    const int* b = reinterpret_cast<const int*>(insts + wk.inst_b);
    int* dst = reinterpret_cast<int*>(&sI[0]);
    constexpr int kWords = (int)(sizeof(WelshInst) / sizeof(int));
    for (int i = threadIdx.x; i < 2 * kWords; i += 32 * W) dst[i] = i < kWords ? a[i] : b[i - kWords];
    const int* ra = reinterpret_cast<const int*>(tabs + wk.inst_a);
    const int* rb = reinterpret_cast<const int*>(tabs + wk.inst_b);
    int* rdst = reinterpret_cast<int*>(&sR[0]);
    constexpr int kRWords = (int)(sizeof(Rest16Table) / sizeof(int));
    for (int i = threadIdx.x; i < 2 * kRWords; i += 32 * W) rdst[i] = i < kRWords ? ra[i] : rb[i - kRWords];
  }
  __syncthreads();
  RestState* cache = reinterpret_cast<RestState*>(smem_tiles + W * kTile16Stride);
  for (int t = threadIdx.x; t < wk.nvoices; t += 32 * W) {
    const WelshInst& I = sI[t >= wk.split ? 1 : 0];
    const WelshVoice* vp = voices + wk.voice0 + t;
    RestState r;
    const u64 k = (u64)(f0 - 1 - vp->anchor);
    r.d1 = vp->d1; r.d2 = vp->d2;
    r.p1 = vp->p1 + k * r.d1; r.p2 = vp->p2 + k * r.d2;
    r.s[0] = vp->s[0]; r.s[1] = vp->s[1]; r.s[2] = vp->s[2]; r.s[3] = vp->s[3];
    r.ls = 0.0; r.lc = 0.0;
    if (LFO_AMP) {
      double ls, lc;
      sincos_phase(vp->pl + (k + 1) * I.lfo_dq, &ls, &lc);
      const double dl = I.depth * I.amp_rest;
      r.ls = ls * dl; r.lc = lc * dl;
    }
    cache[t] = r;
  }
  const int g = 2 * warp - min(max(warp - wk.single0, 0), 2);
  const bool pair = !(warp == wk.single0 || warp == wk.single0 + 1);
  __syncthreads();
  const int which = g >= wk.split ? 1 : 0;
  const WelshInst& I = sI[which];
  const Rest16Table& R = sR[which];
  double2* tile_row = smem_tiles + warp * kTile16Stride;
  if (g >= wk.nvoices)  // a warp without voices: its row stays zero for the whole launch
    for (int i = lane; i < kTile16Stride; i += 32) tile_row[i] = make_double2(0.0, 0.0);
  const i64 f_end = f0 + nframes;
#pragma unroll 1
  for (i64 fb = f0; fb < f_end; fb += kBlock16) {
    if (pair && g + 1 < wk.nvoices) {
      RestState* const two[2] = {cache + g, cache + g + 1};
      welsh_rest_block16<LFO_AMP, ZERO_A, 2, OSC>(two, I, R, lane, tile_row);
    } else if (g < wk.nvoices) {
      RestState* const one[1] = {cache + g};
      welsh_rest_block16<LFO_AMP, ZERO_A, 1, OSC>(one, I, R, lane, tile_row);
    }
    __syncthreads();
    cta_reduce_store16<W>(smem_tiles, wk.out, fb, f0);
    __syncthreads();
  }
  for (int t = threadIdx.x; t < wk.nvoices; t += 32 * W) {
    WelshVoice* vp = voices + wk.voice0 + t;
    vp->s[0] = cache[t].s[0]; vp->s[1] = cache[t].s[1]; vp->s[2] = cache[t].s[2]; vp->s[3] = cache[t].s[3];
    vp->knot_frame = kNever;
  }
}

// ---- the resting-voice kernel, time-parallel ---------------------------------------------------------
// welsh_rest_kernel gives every warp its own voices and walks them through the chunk block by block: the
// chunk's 256 blocks are a serial chain per warp (1.7 us per block), so a shard of a few hundred voices —
// config 4 split over 8 GPUs — is latency-bound on a fraction of the SMs (profiles/r2_strong_probe.txt).
// Here the W warps of a CTA take W CONSECUTIVE blocks of the SAME voice pair instead: a resting voice is a
// time-invariant recurrence, so every warp runs its block from a zero entry state (the scan of
// welsh_rest_block, unchanged), publishes the block's end vector z_w, and after one CTA barrier each warp
// obtains its true block entry by the recurrence E_{w+1} = M^32 E_w + z_w (M^32: per-instrument constant, w
// steps for warp w) and adds M^lane E_w to its lanes' entry states (a 64-entry table built at kernel start).  Two barriers per pair
// and round (one per filter section); a chain of 256 blocks becomes 32 rounds.  CTAs hold 2..8 voices
// (pairs in sequence), so 512 voices are 256 CTAs.
struct alignas(16) TpState {
  u64 d1, d2;
  double s[2][4];  // filter state at the start of the round, double-buffered by round parity
};
struct alignas(16) TpPriv {  // a warp's own copy of the round-start phases and LFO phasor (advanced in step by all warps)
  u64 p1, p2;
  double ls, lc;
};
struct alignas(16) TpScratch {
  double s[4];    // zero-entry block end vectors of the two sections (written by the scan)
};

template <bool LFO_AMP, bool ZERO_A, int NV, bool ACC, bool SYNC, int W>
__device__ __forceinline__ void welsh_rest_block_tp(TpState* const (&ts)[NV], TpPriv* const (&pv)[NV], const WelshInst& I,
                                                    int warp, int lane, int nb, int par, TpScratch (*zs)[W],
                                                    const double2* brot, const double4* mlane, double2* tile_row) {
  const LtiTable& L = I.lti;
  double yp[NV][kT];
  double ps0[NV], ps1[NV];
  TpScratch* zr[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) zr[v] = &zs[v][warp];
  // ---- pass 1: oscillators + section 1 from a zero state ----
  {
    u64 p1[NV], p2[NV], d1[NV], d2[NV];
    const u64 koff = (u64)(warp * kBlockFrames + lane * kT);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const ulonglong2 pp = *reinterpret_cast<const ulonglong2*>(&pv[v]->p1);
      const ulonglong2 dd = *reinterpret_cast<const ulonglong2*>(&ts[v]->d1);
      d1[v] = dd.x; d2[v] = dd.y;
      osc_start_at<SYNC>(pp.x, pp.y, dd.x, dd.y, koff, p1[v], p2[v]);
      ps0[v] = 0.0; ps1[v] = 0.0;
    }
    const OscMix o1 = I.m1bb, o2 = I.m2bb;
    const u64 t1 = I.s1.thresh, t2 = I.s2.thresh;
    const double a1 = L.c1.a1, a2 = L.c1.a2;
#pragma unroll
    for (int j = 0; j < kT; ++j) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        osc_advance<SYNC>(p1[v], p2[v], d1[v], d2[v]);
        yp[v][j] = lp_step_bx(osc_mix_eval<ZERO_A>(o1, t1, p1[v], o2, t2, p2[v]), a1, a2, ps0[v], ps1[v]);
      }
    }
    const double inv = L.inv_b0_2;
#pragma unroll
    for (int v = 0; v < NV; ++v) { ps0[v] *= inv; ps1[v] *= inv; }
  }
  double e0[NV], e1[NV];
  // one filter section: zero-entry scan, publish, barrier, block entry by the W-step recurrence, lane correction;
  // warp 0 leaves the state after the round's nb blocks in the other parity's slot
  auto section = [&](const double (*mp)[4], const double* m32, int sec) {
    if (lane == 0) {
#pragma unroll
      for (int v = 0; v < NV; ++v) { zr[v]->s[2 * sec] = 0.0; zr[v]->s[2 * sec + 1] = 0.0; }
    }
    __syncwarp();
    lti_scan_entry_t<NV, TpScratch>(ps0, ps1, mp, lane, zr, sec, e0, e1);
    __syncthreads();
    const double2 r0 = *reinterpret_cast<const double2*>(m32), r1 = *reinterpret_cast<const double2*>(m32 + 2);
    // warp w needs the state after w blocks; warp 0 (whose own entry is the round's) runs all nb steps for the
    // state after the round
    const int steps = warp == 0 ? nb : warp;
    const double4 ml = mlane[sec * 32 + lane];  // M^lane of this section, row-major
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const double2 st = *reinterpret_cast<const double2*>(&ts[v]->s[par][2 * sec]);
      double x0 = st.x, x1 = st.y;
#pragma unroll
      for (int i = 0; i < W; ++i) {
        if (i < steps) {
          const double2 z = *reinterpret_cast<const double2*>(&zs[v][i].s[2 * sec]);
          double n0 = z.x, n1 = z.y;
          affine_vec_step(n0, n1, r0.x, r0.y, r1.x, r1.y, x0, x1);
          x0 = n0; x1 = n1;
        }
      }
      if (warp == 0 && lane == 0) *reinterpret_cast<double2*>(&ts[v]->s[par ^ 1][2 * sec]) = make_double2(x0, x1);
      const double c0 = warp == 0 ? st.x : x0, c1 = warp == 0 ? st.y : x1;  // this warp's block entry
      e0[v] = fma(ml.y, c1, fma(ml.x, c0, e0[v]));                           // + M^lane applied to it
      e1[v] = fma(ml.w, c1, fma(ml.z, c0, e1[v]));
    }
  };
  section(L.mp1, I.mp32_1, 0);
  // ---- pass 2: section 2 on the fixed-up section-1 output ----
  {
    const double a1 = L.c2.a1, a2 = L.c2.a2;
#pragma unroll
    for (int v = 0; v < NV; ++v) { ps0[v] = 0.0; ps1[v] = 0.0; }
#pragma unroll
    for (int j = 0; j < kT; ++j) {
      const double2 g = *reinterpret_cast<const double2*>(L.g1b[j]);
#pragma unroll
      for (int v = 0; v < NV; ++v)
        yp[v][j] = lp_step_bx(fma(g.y, e1[v], fma(g.x, e0[v], yp[v][j])), a1, a2, ps0[v], ps1[v]);
    }
  }
  section(L.mp2, I.mp32_2, 1);
  // ---- LFO phasor of this lane: the round's phasor turned by `warp` blocks and the lane's offset ----
  double lsd[NV], lcd[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    lsd[v] = 0.0; lcd[v] = 0.0;
    if (LFO_AMP) {
      const double2 ph = *reinterpret_cast<const double2*>(&pv[v]->ls);
      const double2 b = brot[warp], r = I.lane_rot[lane];
      const double bs = fma(ph.x, b.x, ph.y * b.y), bc = fma(ph.y, b.x, -(ph.x * b.y));
      lsd[v] = fma(bs, r.x, bc * r.y);
      lcd[v] = fma(bc, r.x, -(bs * r.y));
    }
  }
  // ---- amplitude, DCA, into the warp's tile row ----
  const double arest = I.amp_rest;
  const double gl = I.gl, gr = I.gr;
  double2* row = tile_row + lane * (kT + 1);
#pragma unroll
  for (int j = 0; j < kT; ++j) {
    const double2 g = *reinterpret_cast<const double2*>(L.g2[j]);
    double2 rot = make_double2(0.0, 0.0);
    if (LFO_AMP) rot = I.lfo_rot[j];
    double m = 0.0;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const double amp = LFO_AMP ? fma(lsd[v], rot.x, fma(lcd[v], rot.y, arest)) : arest;
      const double y = fma(g.y, e1[v], fma(g.x, e0[v], yp[v][j]));
      m = v == 0 ? y * amp : fma(y, amp, m);
    }
    if (ACC) {
      const double2 p = row[j];
      row[j] = make_double2(fma(m, gl, p.x), fma(m, gr, p.y));
    } else {
      row[j] = make_double2(m * gl, m * gr);
    }
  }
  __syncwarp();  // every lane has read the warp's round-start copies
  if (lane == 0) {  // ... which advance by the round's nb blocks
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      u64 n1, n2;
      osc_start_at<SYNC>(pv[v]->p1, pv[v]->p2, ts[v]->d1, ts[v]->d2, (u64)(nb * kBlockFrames), n1, n2);
      pv[v]->p1 = n1; pv[v]->p2 = n2;
      if (LFO_AMP) {
        const double2 ph = *reinterpret_cast<const double2*>(&pv[v]->ls);
        const double2 b = brot[nb];
        *reinterpret_cast<double2*>(&pv[v]->ls) = make_double2(fma(ph.x, b.x, ph.y * b.y), fma(ph.y, b.x, -(ph.x * b.y)));
      }
    }
  }
  __syncwarp();
}

// grid = number of resting CTAs of this variant with at most kTpMaxVoices voices; block = 32 * W threads;
// dynamic smem = W tile rows + max_voices TpState + W * max_voices TpPriv.  nframes is a multiple of kBlockFrames.
constexpr int kTpMaxVoices = 8;
template <int W, bool LFO_AMP, bool ZERO_A, bool SYNC = false>
__global__ void __launch_bounds__(32 * W, 2) welsh_rest_tp_kernel(const WelshInst* __restrict__ insts,
                                                                WelshVoice* __restrict__ voices,
                                                                const CtaWork* __restrict__ work,
                                                                const int* __restrict__ idx, i64 f0, int nframes,
                                                                int max_voices) {
  extern __shared__ double2 smem_tiles[];
  __shared__ WelshInst sI;
  __shared__ TpScratch zs[2][2][W];  // [pair parity][voice of the pair][warp]
  __shared__ double2 brot[W + 1];    // the LFO's block rotation to the k-th power
  __shared__ double4 mlane[64];      // [section][lane]: M^lane
  const CtaWork wk = work[idx[blockIdx.x]];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const int* src = reinterpret_cast<const int*>(insts + wk.inst);
    int* dst = reinterpret_cast<int*>(&sI);
    for (int i = threadIdx.x; i < (int)(sizeof(WelshInst) / sizeof(int)); i += 32 * W) dst[i] = src[i];
  }
  __syncthreads();
  const WelshInst& I = sI;
  TpState* cache = reinterpret_cast<TpState*>(smem_tiles + W * kTileStride);
  TpPriv* priv = reinterpret_cast<TpPriv*>(cache + max_voices) + warp * max_voices;
  for (int t = threadIdx.x; t < wk.nvoices * (W + 1); t += 32 * W) {
    const int v = t % wk.nvoices, who = t / wk.nvoices;  // who < W: that warp's private copy; who == W: the shared state
    const WelshVoice* vp = voices + wk.voice0 + v;
    const u64 k = (u64)(f0 - 1 - vp->anchor);
    if (who == W) {
      TpState r;
      r.d1 = vp->d1; r.d2 = vp->d2;
      r.s[0][0] = vp->s[0]; r.s[0][1] = vp->s[1]; r.s[0][2] = vp->s[2]; r.s[0][3] = vp->s[3];
      r.s[1][0] = 0.0; r.s[1][1] = 0.0; r.s[1][2] = 0.0; r.s[1][3] = 0.0;
      cache[v] = r;
    } else {
      TpPriv q;
      q.p1 = vp->p1 + k * vp->d1; q.p2 = vp->p2 + k * vp->d2;
      if (SYNC) q.p2 = welsh_phases_at(*vp, I, f0 - 1).p2;
      q.ls = 0.0; q.lc = 0.0;
      if (LFO_AMP) {
        double ls, lc;
        sincos_phase(vp->pl + (k + 1) * I.lfo_dq, &ls, &lc);
        const double dl = I.depth * I.amp_rest;
        q.ls = ls * dl; q.lc = lc * dl;
      }
      reinterpret_cast<TpPriv*>(cache + max_voices)[who * max_voices + v] = q;
    }
  }
  if (threadIdx.x < 64) {  // M^lane of both sections: the lane map to the lane's power, by its binary digits
    const int sec = threadIdx.x >> 5, l = threadIdx.x & 31;
    const double (*mp)[4] = sec ? I.lti.mp2 : I.lti.mp1;
    double a00 = 1.0, a01 = 0.0, a10 = 0.0, a11 = 1.0;
    for (int k = 0; k < 5; ++k)
      if ((l >> k) & 1) {
        const double* m = mp[k];
        const double b00 = m[0] * a00 + m[1] * a10, b01 = m[0] * a01 + m[1] * a11;
        const double b10 = m[2] * a00 + m[3] * a10, b11 = m[2] * a01 + m[3] * a11;
        a00 = b00; a01 = b01; a10 = b10; a11 = b11;
      }
    mlane[threadIdx.x] = make_double4(a00, a01, a10, a11);
  }
  if (threadIdx.x <= W) {
    double c = 1.0, sn = 0.0;  // (cos, sin) of k block rotations
    for (int k = 0; k < (int)threadIdx.x; ++k) {
      const double nc = c * I.block_rot.x - sn * I.block_rot.y, ns = sn * I.block_rot.x + c * I.block_rot.y;
      c = nc; sn = ns;
    }
    brot[threadIdx.x] = make_double2(c, sn);
  }
  __syncthreads();
  double2* tile_row = smem_tiles + warp * kTileStride;
  const i64 f_end = f0 + nframes;
  int par = 0, pairs = 0;
#pragma unroll 1
  for (int r0 = 0; r0 < nframes; r0 += W * kBlockFrames, par ^= 1) {
    const int nb = min(W, (nframes - r0) / kBlockFrames);
    bool first = true;
#pragma unroll 1
    for (int g = 0; g < wk.nvoices; g += 2, ++pairs) {
      TpScratch (*z)[W] = zs[pairs & 1];
      if (g + 1 < wk.nvoices) {
        TpState* const two[2] = {cache + g, cache + g + 1};
        TpPriv* const pq[2] = {priv + g, priv + g + 1};
        if (first) welsh_rest_block_tp<LFO_AMP, ZERO_A, 2, false, SYNC, W>(two, pq, I, warp, lane, nb, par, z, brot, mlane, tile_row);
        else welsh_rest_block_tp<LFO_AMP, ZERO_A, 2, true, SYNC, W>(two, pq, I, warp, lane, nb, par, z, brot, mlane, tile_row);
      } else {
        TpState* const one[1] = {cache + g};
        TpPriv* const pq[1] = {priv + g};
        if (first) welsh_rest_block_tp<LFO_AMP, ZERO_A, 1, false, SYNC, W>(one, pq, I, warp, lane, nb, par, z, brot, mlane, tile_row);
        else welsh_rest_block_tp<LFO_AMP, ZERO_A, 1, true, SYNC, W>(one, pq, I, warp, lane, nb, par, z, brot, mlane, tile_row);
      }
      first = false;
    }
    if (warp < nb) warp_store_row(tile_row, true, wk.out, f0 + r0 + (i64)warp * kBlockFrames, f0, f_end, lane);
    __syncwarp();
  }
  __syncthreads();  // warp 0's last state write
  for (int t = threadIdx.x; t < wk.nvoices; t += 32 * W) {
    WelshVoice* vp = voices + wk.voice0 + t;
    vp->s[0] = cache[t].s[par][0]; vp->s[1] = cache[t].s[par][1]; vp->s[2] = cache[t].s[par][2]; vp->s[3] = cache[t].s[par][3];
    vp->knot_frame = kNever;
  }
}

// ---- the sweeping-voice kernel ---------------------------------------------------------------------
// The moving-cutoff counterpart of welsh_rest_kernel: CTAs whose voices are all held for the whole chunk
// with the filter envelope inside ONE moving stage (attack or decay), the amplitude envelope inside one
// stage, no note event in the chunk and the cutoff moving slowly enough for coefficient knots.  Again
// the host decides (note frames and stage lengths are integers), nothing is classified per block, and
// the voice state lives in shared memory for the launch.  Per block a lane evaluates ONE exact
// coefficient set (at the end of its kT frames); with its two predecessors' sets — the lane's own start
// and the start of the lane before, carried across blocks — a one-sided quadratic gives the per-frame
// a1, a2 of both sections (error 8x that of welsh_block_simple's centred form: <= 1e-10 at the knot
// threshold), and b0 follows from the unity DC gain of each section: 4 b0 = 1 - a1 - a2.
struct alignas(16) SweepState {
  u64 p1, p2;     // oscillator phases at the frame before the current block
  u64 d1, d2;
  double s[4];    // filter state at the start of the current block
  double ls, lc;  // depth * (sin, cos) of the LFO angle at the first frame of the current block
  double aq0, aq1, aq2, aw, adw;  // amplitude stage (0.5 folded in): level = q0 + w (q1 + q2 w), w = aw + t adw, t = frame - f0
  double fq0, fq1, fq2, fw, fdw;  // filter-envelope stage, same form
  double kn[2][4];                // carried knots (a1, a2 of section 1, a1, a2 of section 2) at frames fb - kT and fb
};

// Stage parameters of a HELD note whose envelope stays inside one stage from frame f on (env_segment
// without the per-lane containment tests: the host has checked the whole chunk).
__device__ __forceinline__ void env_stage_held(const EnvShape& sh, i64 n_on, double l_on, i64 f, double& q0, double& q1,
                                               double& q2, double& w0, double& dw) {
  const i64 k = f - n_on;
  q1 = 0.0; q2 = 0.0; w0 = 0.0; dw = 0.0;
  if (k < sh.na) {
    w0 = (double)k * sh.inv_na; dw = sh.inv_na;
    q0 = l_on; q1 = 2.0 * (1.0 - l_on); q2 = -(1.0 - l_on);
  } else if (k - sh.na < sh.nd) {
    w0 = 1.0 - (double)(k - sh.na) * sh.inv_nd; dw = -sh.inv_nd;
    q0 = sh.sustain; q2 = 1.0 - sh.sustain;
  } else {
    q0 = sh.sustain;
  }
}
// a1, a2 of both sections for an UNCLAMPED cutoff fraction (the knot before a stage's first frame is
// the stage's own formula continued backwards; the frequency clamp still applies)
__device__ __forceinline__ void welsh_knot(const WelshInst& I, double pct, double (&k)[4]) {
  double u = exp2_ranged(fma(pct, kLog2_800, I.log2_25_over_sr));
  u = u > I.u_max ? I.u_max : u;
  u = u < I.u_min ? I.u_min : u;
  SecCoef c1, c2;
  lp24_from_u(I.rp, u, c1, c2);
  k[0] = c1.a1; k[1] = c1.a2; k[2] = c2.a1; k[3] = c2.a2;
}

template <bool LFO_AMP, bool ZERO_A, bool ACC, bool SYNC = false>
__device__ __forceinline__ void welsh_sweep_block(SweepState* rs, const WelshInst& I, int lane, int t0,
                                                  double2* tile_row) {
  // t0 = first frame of the block relative to the chunk start
  const int tl = t0 + lane * kT;  // the lane's first frame
  // ---- coefficient knots: exact at the lane's end, the two before it from the neighbours / the cache ----
  double d1q[4], d2q[4], k0q[4];
  {
    double e[4];
    const double w8 = fma((double)(tl + kT), rs->fdw, rs->fw);
    welsh_knot(I, fma(I.cut_b, fma(w8, fma(rs->fq2, w8, rs->fq1), rs->fq0), I.cut_a), e);
    const double4 c0 = *reinterpret_cast<const double4*>(rs->kn[0]), c1 = *reinterpret_cast<const double4*>(rs->kn[1]);
    const double cm2[4] = {c0.x, c0.y, c0.z, c0.w}, cm1[4] = {c1.x, c1.y, c1.z, c1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double sv = shfl_up_f64(e[i], 1), pv = shfl_up_f64(e[i], 2);
      sv = lane == 0 ? cm1[i] : sv;
      pv = lane == 0 ? cm2[i] : (lane == 1 ? cm1[i] : pv);
      k0q[i] = sv;
      d1q[i] = (e[i] - pv) * (1.0 / (2 * kT));
      d2q[i] = ((e[i] - sv) - (sv - pv)) * (1.0 / (2 * kT * kT));
    }
    __syncwarp();  // every lane has read the carried knots
    if (lane >= 30) *reinterpret_cast<double4*>(rs->kn[lane - 30]) = make_double4(e[0], e[1], e[2], e[3]);
  }
  auto coef = [&](int i, int j) { return fma((double)j, fma((double)j, d2q[i], d1q[i]), k0q[i]); };
  // ---- pass 1: oscillators + section 1 from a zero state with its homogeneous response ----
  double yp[kT], g0[kT], g1[kT];
  double ps0 = 0.0, ps1 = 0.0, h00 = 1.0, h01 = 0.0, h10 = 0.0, h11 = 1.0;
  {
    const ulonglong2 pp = *reinterpret_cast<const ulonglong2*>(&rs->p1);
    const ulonglong2 dd = *reinterpret_cast<const ulonglong2*>(&rs->d1);
    u64 p1, p2;
    osc_lane_start<SYNC>(pp.x, pp.y, dd.x, dd.y, lane, p1, p2);
    __syncwarp();  // every lane has read the block's base phases
    if (!SYNC && lane == 0)
      *reinterpret_cast<ulonglong2*>(&rs->p1) =
          make_ulonglong2(p1 + (u64)kBlockFrames * dd.x, p2 + (u64)kBlockFrames * dd.y);
    const OscMix o1 = I.m1, o2 = I.m2;
    const u64 t1 = I.s1.thresh, t2 = I.s2.thresh;
#pragma unroll
    for (int j = 0; j < kT; ++j) {
      osc_advance<SYNC>(p1, p2, dd.x, dd.y);
      const double x = osc_mix_eval<ZERO_A>(o1, t1, p1, o2, t2, p2);
      const double a1 = coef(0, j), a2 = coef(1, j);
      const double b0 = fma(-0.25, a1 + a2, 0.25);
      g0[j] = h00; g1[j] = h01;
      yp[j] = lp_step(b0, a1, a2, x, ps0, ps1);
      const double t00 = fma(a1, h00, h10), t01 = fma(a1, h01, h11);
      h10 = a2 * h00; h11 = a2 * h01;
      h00 = t00; h01 = t01;
    }
    if (SYNC && lane == 31) *reinterpret_cast<ulonglong2*>(&rs->p1) = make_ulonglong2(p1, p2);
  }
  double e0, e1, end0, end1;
  {
    Affine2 a;
    a.m00 = h00; a.m01 = h01; a.m10 = h10; a.m11 = h11; a.v0 = ps0; a.v1 = ps1;
    const double2 st = *reinterpret_cast<const double2*>(&rs->s[0]);  // only lane 0 uses it
    affine_scan_states(a, lane, st.x, st.y, e0, e1, end0, end1);
    __syncwarp();
    if (lane == 0) *reinterpret_cast<double2*>(&rs->s[0]) = make_double2(end0, end1);
  }
  // ---- pass 2: section 2 on the fixed-up section-1 output ----
  ps0 = 0.0; ps1 = 0.0; h00 = 1.0; h01 = 0.0; h10 = 0.0; h11 = 1.0;
#pragma unroll
  for (int j = 0; j < kT; ++j) {
    const double a1 = coef(2, j), a2 = coef(3, j);
    const double b0 = fma(-0.25, a1 + a2, 0.25);
    const double x = fma(g1[j], e1, fma(g0[j], e0, yp[j]));
    g0[j] = h00; g1[j] = h01;
    yp[j] = lp_step(b0, a1, a2, x, ps0, ps1);
    const double t00 = fma(a1, h00, h10), t01 = fma(a1, h01, h11);
    h10 = a2 * h00; h11 = a2 * h01;
    h00 = t00; h01 = t01;
  }
  {
    Affine2 a;
    a.m00 = h00; a.m01 = h01; a.m10 = h10; a.m11 = h11; a.v0 = ps0; a.v1 = ps1;
    const double2 st = *reinterpret_cast<const double2*>(&rs->s[2]);
    affine_scan_states(a, lane, st.x, st.y, e0, e1, end0, end1);
    __syncwarp();
    if (lane == 0) *reinterpret_cast<double2*>(&rs->s[2]) = make_double2(end0, end1);
  }
  // ---- LFO phasor of this lane; the cached one advances by one block ----
  double lsd = 0.0, lcd = 0.0;
  if (LFO_AMP) {
    const double2 ph = *reinterpret_cast<const double2*>(&rs->ls);
    const double2 r = I.lane_rot[lane];
    lsd = fma(ph.x, r.x, ph.y * r.y);
    lcd = fma(ph.y, r.x, -(ph.x * r.y));
    __syncwarp();
    if (lane == 0) {
      const double2 br = I.block_rot;
      *reinterpret_cast<double2*>(&rs->ls) = make_double2(fma(lsd, br.x, lcd * br.y), fma(lcd, br.x, -(lsd * br.y)));
    }
  }
  // ---- amplitude (stage x LFO), DCA, into the warp's tile row ----
  const double aq0 = rs->aq0, aq1 = rs->aq1, aq2 = rs->aq2, aw = rs->aw, adw = rs->adw;
  const double gl = I.gl, gr = I.gr;
  double2* row = tile_row + lane * (kT + 1);
#pragma unroll
  for (int j = 0; j < kT; ++j) {
    const double w = fma((double)(tl + j), adw, aw);
    double amp = fma(w, fma(aq2, w, aq1), aq0);
    if (LFO_AMP) {
      const double2 rot = I.lfo_rot[j];
      amp *= fma(lsd, rot.x, fma(lcd, rot.y, 1.0));
    }
    const double m = fma(g1[j], e1, fma(g0[j], e0, yp[j])) * amp;
    if (ACC) {
      const double2 p = row[j];
      row[j] = make_double2(fma(m, gl, p.x), fma(m, gr, p.y));
    } else {
      row[j] = make_double2(m * gl, m * gr);
    }
  }
  __syncwarp();
}

// welsh_sweep_block with the coefficient sets of BOTH sections evaluated exactly at every frame
// (welsh_coef_exact, as the general path does) instead of interpolated between knots: for filter-envelope
// stages whose cutoff moves faster than the knot threshold allows (short attacks, decays and releases).
// Section 2's a1, a2 of the lane's kT frames wait in the thread's parking column of shared memory
// (`park`, stride `pstride` doubles) while section 1 is scanned; b0 of section 2 follows from the unity
// DC gain, 4 b0 = 1 - a1 - a2 (an identity of lp24_from_u's forms).
template <bool LFO_AMP>
__device__ __forceinline__ void welsh_exact_block(SweepState* rs, const WelshInst& I, int lane, int t0, double2* tile_row,
                                                  double* park, int pstride) {
  const int tl = t0 + lane * kT;  // the lane's first frame, relative to the item start
  double yp[kT], g0[kT], g1[kT];
  double ps0 = 0.0, ps1 = 0.0, h00 = 1.0, h01 = 0.0, h10 = 0.0, h11 = 1.0;
  {
    const ulonglong2 pp = *reinterpret_cast<const ulonglong2*>(&rs->p1);
    const ulonglong2 dd = *reinterpret_cast<const ulonglong2*>(&rs->d1);
    const u64 k = (u64)(lane * kT);
    u64 p1 = pp.x + k * dd.x, p2 = pp.y + k * dd.y;
    __syncwarp();  // every lane has read the block's base phases
    if (lane == 0)
      *reinterpret_cast<ulonglong2*>(&rs->p1) =
          make_ulonglong2(p1 + (u64)kBlockFrames * dd.x, p2 + (u64)kBlockFrames * dd.y);
    const OscMix o1 = I.m1, o2 = I.m2;
    const u64 t1 = I.s1.thresh, t2 = I.s2.thresh;
    const double fq0 = rs->fq0, fq1 = rs->fq1, fq2 = rs->fq2, fw = rs->fw, fdw = rs->fdw;
    const double cut_a = I.cut_a, cut_b = I.cut_b;
#pragma unroll
    for (int j = 0; j < kT; ++j) {
      p1 += dd.x;
      p2 += dd.y;
      const double x = osc_mix_eval<false>(o1, t1, p1, o2, t2, p2);
      const double w = fma((double)(tl + j), fdw, fw);
      SecCoef c1, c2;
      welsh_coef_exact(I, fma(cut_b, fma(w, fma(fq2, w, fq1), fq0), cut_a), c1, c2);
      park[(2 * j) * pstride] = c2.a1;
      park[(2 * j + 1) * pstride] = c2.a2;
      g0[j] = h00; g1[j] = h01;
      yp[j] = lp_step(c1.b0, c1.a1, c1.a2, x, ps0, ps1);
      const double t00 = fma(c1.a1, h00, h10), t01 = fma(c1.a1, h01, h11);
      h10 = c1.a2 * h00; h11 = c1.a2 * h01;
      h00 = t00; h01 = t01;
    }
  }
  double e0, e1, end0, end1;
  {
    Affine2 a;
    a.m00 = h00; a.m01 = h01; a.m10 = h10; a.m11 = h11; a.v0 = ps0; a.v1 = ps1;
    const double2 st = *reinterpret_cast<const double2*>(&rs->s[0]);  // only lane 0 uses it
    affine_scan_states(a, lane, st.x, st.y, e0, e1, end0, end1);
    __syncwarp();
    if (lane == 0) *reinterpret_cast<double2*>(&rs->s[0]) = make_double2(end0, end1);
  }
  ps0 = 0.0; ps1 = 0.0; h00 = 1.0; h01 = 0.0; h10 = 0.0; h11 = 1.0;
#pragma unroll
  for (int j = 0; j < kT; ++j) {
    const double a1 = park[(2 * j) * pstride], a2 = park[(2 * j + 1) * pstride];
    const double b0 = fma(-0.25, a1 + a2, 0.25);
    const double x = fma(g1[j], e1, fma(g0[j], e0, yp[j]));
    g0[j] = h00; g1[j] = h01;
    yp[j] = lp_step(b0, a1, a2, x, ps0, ps1);
    const double t00 = fma(a1, h00, h10), t01 = fma(a1, h01, h11);
    h10 = a2 * h00; h11 = a2 * h01;
    h00 = t00; h01 = t01;
  }
  {
    Affine2 a;
    a.m00 = h00; a.m01 = h01; a.m10 = h10; a.m11 = h11; a.v0 = ps0; a.v1 = ps1;
    const double2 st = *reinterpret_cast<const double2*>(&rs->s[2]);
    affine_scan_states(a, lane, st.x, st.y, e0, e1, end0, end1);
    __syncwarp();
    if (lane == 0) *reinterpret_cast<double2*>(&rs->s[2]) = make_double2(end0, end1);
  }
  double lsd = 0.0, lcd = 0.0;
  if (LFO_AMP) {
    const double2 ph = *reinterpret_cast<const double2*>(&rs->ls);
    const double2 r = I.lane_rot[lane];
    lsd = fma(ph.x, r.x, ph.y * r.y);
    lcd = fma(ph.y, r.x, -(ph.x * r.y));
    __syncwarp();
    if (lane == 0) {
      const double2 br = I.block_rot;
      *reinterpret_cast<double2*>(&rs->ls) = make_double2(fma(lsd, br.x, lcd * br.y), fma(lcd, br.x, -(lsd * br.y)));
    }
  }
  const double aq0 = rs->aq0, aq1 = rs->aq1, aq2 = rs->aq2, aw = rs->aw, adw = rs->adw;
  const double gl = I.gl, gr = I.gr;
  double2* row = tile_row + lane * (kT + 1);
#pragma unroll
  for (int j = 0; j < kT; ++j) {
    const double w = fma((double)(tl + j), adw, aw);
    double amp = fma(w, fma(aq2, w, aq1), aq0);
    if (LFO_AMP) {
      const double2 rot = I.lfo_rot[j];
      amp *= fma(lsd, rot.x, fma(lcd, rot.y, 1.0));
    }
    const double m = fma(g1[j], e1, fma(g0[j], e0, yp[j])) * amp;
    row[j] = make_double2(m * gl, m * gr);
  }
  __syncwarp();
}

// grid = number of sweeping CTAs of this variant; block = 32 * W threads;
// dynamic smem = W * kTileStride double2 (tiles) + max_voices SweepState.  nframes is a multiple of kBlockFrames.
template <int W, bool LFO_AMP, bool ZERO_A, bool SYNC = false>
__global__ void __launch_bounds__(32 * W, 2) welsh_sweep_kernel(const WelshInst* __restrict__ insts,
                                                              WelshVoice* __restrict__ voices,
                                                              const CtaWork* __restrict__ work,
                                                              const int* __restrict__ idx, i64 f0, int nframes) {
  extern __shared__ double2 smem_tiles[];
  __shared__ int s_active[W];
  __shared__ WelshInst sI;
  const CtaWork wk = work[idx[blockIdx.x]];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const int* src = reinterpret_cast<const int*>(insts + wk.inst);
    int* dst = reinterpret_cast<int*>(&sI);
    for (int i = threadIdx.x; i < (int)(sizeof(WelshInst) / sizeof(int)); i += 32 * W) dst[i] = src[i];
  }
  __syncthreads();
  const WelshInst& I = sI;
  SweepState* cache = reinterpret_cast<SweepState*>(smem_tiles + W * kTileStride);
  for (int t = threadIdx.x; t < wk.nvoices; t += 32 * W) {
    const WelshVoice* vp = voices + wk.voice0 + t;
    SweepState r;
    const u64 k = (u64)(f0 - 1 - vp->anchor);
    r.d1 = vp->d1; r.d2 = vp->d2;
    r.p1 = vp->p1 + k * r.d1; r.p2 = vp->p2 + k * r.d2;
    if (SYNC) r.p2 = welsh_phases_at(*vp, I, f0 - 1).p2;  // oscillator 2 restarted at oscillator 1's last wrap
    r.s[0] = vp->s[0]; r.s[1] = vp->s[1]; r.s[2] = vp->s[2]; r.s[3] = vp->s[3];
    r.ls = 0.0; r.lc = 0.0;
    if (LFO_AMP) {
      double ls, lc;
      sincos_phase(vp->pl + (k + 1) * I.lfo_dq, &ls, &lc);
      r.ls = ls * I.depth; r.lc = lc * I.depth;
    }
    env_stage_held(I.amp, vp->n_on, vp->la_on, f0, r.aq0, r.aq1, r.aq2, r.aw, r.adw);
    r.aq0 *= 0.5; r.aq1 *= 0.5; r.aq2 *= 0.5;
    env_stage_held(I.filt, vp->n_on, vp->lf_on, f0, r.fq0, r.fq1, r.fq2, r.fw, r.fdw);
#pragma unroll
    for (int q = 0; q < 2; ++q) {  // knots at f0 - kT and f0: the stage's formula, continued backwards
      const double w = fma((double)((q - 1) * kT), r.fdw, r.fw);
      welsh_knot(I, fma(I.cut_b, fma(w, fma(r.fq2, w, r.fq1), r.fq0), I.cut_a), r.kn[q]);
    }
    cache[t] = r;
  }
  if (lane == 0) s_active[warp] = warp < wk.nvoices ? 1 : 0;
  __syncthreads();
  double2* tile_row = smem_tiles + warp * kTileStride;
  const i64 f_end = f0 + nframes;
  int t0 = 0;
#pragma unroll 1
  for (i64 fb = f0; fb < f_end; fb += kBlockFrames, t0 += kBlockFrames) {
    bool first = true;
#pragma unroll 1
    for (int g = warp; g < wk.nvoices; g += W) {
      if (first) welsh_sweep_block<LFO_AMP, ZERO_A, false, SYNC>(cache + g, I, lane, t0, tile_row);
      else welsh_sweep_block<LFO_AMP, ZERO_A, true, SYNC>(cache + g, I, lane, t0, tile_row);
      first = false;
    }
    __syncthreads();
    cta_reduce_store<W>(smem_tiles, s_active, wk.out, fb, f0, f_end);
    __syncthreads();
  }
  for (int t = threadIdx.x; t < wk.nvoices; t += 32 * W) {
    WelshVoice* vp = voices + wk.voice0 + t;
    vp->s[0] = cache[t].s[0]; vp->s[1] = cache[t].s[1]; vp->s[2] = cache[t].s[2]; vp->s[3] = cache[t].s[3];
    vp->knot_frame = kNever;
  }
}

// The sweeping kernel over voice ranges (VrWork, as welsh_rest_vr_kernel): chunks in which EVERY grouped CTA sweeps
// run as ranges of 14 voices regardless of instrument boundaries (28 voices on every SM instead of 32 on 108 SMs
// and 16 on the rest).  A warp takes voices w and w + W of its range, each with its own instrument record.
template <int W, bool LFO_AMP, bool ZERO_A>
__global__ void __launch_bounds__(32 * W, 2) welsh_sweep_vr_kernel(const WelshInst* __restrict__ insts,
                                                                 WelshVoice* __restrict__ voices,
                                                                 const VrWork* __restrict__ work, i64 f0, int nframes) {
  extern __shared__ double2 smem_tiles[];
  __shared__ int s_active[W];
  __shared__ WelshInst sI[2];
  const VrWork wk = work[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const int* a = reinterpret_cast<const int*>(insts + wk.inst_a);
    const int* b = reinterpret_cast<const int*>(insts + wk.inst_b);
    int* dst = reinterpret_cast<int*>(&sI[0]);
    constexpr int kWords = (int)(sizeof(WelshInst) / sizeof(int));
    for (int i = threadIdx.x; i < 2 * kWords; i += 32 * W) dst[i] = i < kWords ? a[i] : b[i - kWords];
  }
  __syncthreads();
  SweepState* cache = reinterpret_cast<SweepState*>(smem_tiles + W * kTileStride);
  for (int t = threadIdx.x; t < wk.nvoices; t += 32 * W) {
    const WelshInst& I = sI[t >= wk.split ? 1 : 0];
    const WelshVoice* vp = voices + wk.voice0 + t;
    SweepState r;
    const u64 k = (u64)(f0 - 1 - vp->anchor);
    r.d1 = vp->d1; r.d2 = vp->d2;
    r.p1 = vp->p1 + k * r.d1; r.p2 = vp->p2 + k * r.d2;
    r.s[0] = vp->s[0]; r.s[1] = vp->s[1]; r.s[2] = vp->s[2]; r.s[3] = vp->s[3];
    r.ls = 0.0; r.lc = 0.0;
    if (LFO_AMP) {
      double ls, lc;
      sincos_phase(vp->pl + (k + 1) * I.lfo_dq, &ls, &lc);
      r.ls = ls * I.depth; r.lc = lc * I.depth;
    }
    env_stage_held(I.amp, vp->n_on, vp->la_on, f0, r.aq0, r.aq1, r.aq2, r.aw, r.adw);
    r.aq0 *= 0.5; r.aq1 *= 0.5; r.aq2 *= 0.5;
    env_stage_held(I.filt, vp->n_on, vp->lf_on, f0, r.fq0, r.fq1, r.fq2, r.fw, r.fdw);
#pragma unroll
    for (int q = 0; q < 2; ++q) {  // knots at f0 - kT and f0: the stage's formula, continued backwards
      const double w = fma((double)((q - 1) * kT), r.fdw, r.fw);
      welsh_knot(I, fma(I.cut_b, fma(w, fma(r.fq2, w, r.fq1), r.fq0), I.cut_a), r.kn[q]);
    }
    cache[t] = r;
  }
  if (lane == 0) s_active[warp] = warp < wk.nvoices ? 1 : 0;
  __syncthreads();
  double2* tile_row = smem_tiles + warp * kTileStride;
  const i64 f_end = f0 + nframes;
  int t0 = 0;
#pragma unroll 1
  for (i64 fb = f0; fb < f_end; fb += kBlockFrames, t0 += kBlockFrames) {
    bool first = true;
#pragma unroll 1
    for (int g = warp; g < wk.nvoices; g += W) {
      const WelshInst& I = sI[g >= wk.split ? 1 : 0];
      if (first) welsh_sweep_block<LFO_AMP, ZERO_A, false, false>(cache + g, I, lane, t0, tile_row);
      else welsh_sweep_block<LFO_AMP, ZERO_A, true, false>(cache + g, I, lane, t0, tile_row);
      first = false;
    }
    __syncthreads();
    cta_reduce_store<W>(smem_tiles, s_active, wk.out, fb, f0, f_end);
    __syncthreads();
  }
  for (int t = threadIdx.x; t < wk.nvoices; t += 32 * W) {
    WelshVoice* vp = voices + wk.voice0 + t;
    vp->s[0] = cache[t].s[0]; vp->s[1] = cache[t].s[1]; vp->s[2] = cache[t].s[2]; vp->s[3] = cache[t].s[3];
    vp->knot_frame = kNever;
  }
}

// ---- the solo-voice kernel -------------------------------------------------------------------------
// Instruments with fewer voices than a CTA has warps contribute one (instrument, voice, output buffer)
// item per voice (a batch of one-voice patch variants — BASELINE config 5 — is 4096 such items per GPU).
// Every warp is then a different patch at a different stage of its note.  Left to classify its blocks on
// the device (welsh_kernel<.., SOLO>), the warps of an SM spread over all of the kernel's code paths at
// once: ncu showed 52 % of the stall samples as instruction-cache misses and another 22 % on the global
// voice records (profiles/r2_cfg5_before_*).  Here the HOST does the classification, as it already does
// for grouped CTAs: note frames and envelope stage lengths are integers, so for every sub-chunk of
// kSoloSubDefault frames it knows whether a voice is idle, rests (cutoff constant: welsh_solo_rest), sweeps
// inside one envelope stage slowly enough for coefficient knots (welsh_sweep_block), or needs the
// per-block general machinery (note events, stage boundaries, fast sweeps, non-linear oscillators).
// Items of one class and one sub-chunk are packed W to a job; ONE persistent launch hands the jobs out in
// list order (ticket counter), one per CTA.  The W warps of a job run the same class body and meet at a
// CTA barrier after every 256-frame block, so they stay within a block of each other and share their
// instruction fetches (the instruction cache holds one block body, not several).  Voice state lives in
// shared memory for the length of an item, and a voice's consecutive items are ordered by a per-voice
// progress counter in global memory (a job only ever waits for jobs earlier in the list, which running
// CTAs already hold: no deadlock).
constexpr int kSoloSubDefault = 8 * kBlockFrames;  // frames per sub-chunk (classification granularity; GB_SOLO_SUB overrides)
enum { SOLO_REST = 0, SOLO_SWEEP = 1, SOLO_GENERAL = 2, SOLO_EXACT = 3, SOLO_CLASSES = 4 };

struct SoloItem {
  int item;  // index into the WarpItem table
  int need;  // value of progress[voice] this item waits for (= the voice's earlier non-idle items of the chunk)
};
struct SoloJob {  // up to W items of one class and one sub-chunk: one per warp of a CTA, run in step
  int cls;      // SOLO_*
  int t0;       // first frame of the sub-chunk, relative to the chunk start
  int nframes;  // frames (REST / SWEEP / EXACT: a multiple of kBlockFrames)
  int first;    // first SoloItem
  int n;        // items (<= W)
  int lockstep; // the job's warps meet at a CTA barrier after every block
};

struct alignas(16) SoloRestState {
  u64 p1, p2;     // oscillator phases at the frame before the current block
  u64 d1, d2;
  double s[4];    // filter state at the start of the current block
  double ls, lc;  // depth * (sin, cos) of the LFO angle at the first frame of the current block
  double aq0, aq1, aq2, aw, adw;  // amplitude stage (0.5 folded in): level = q0 + w (q1 + q2 w), w = aw + t adw, t = frame - item start
};

// Stage parameters of an envelope that stays inside ONE stage from frame f on: attack / decay / sustain
// of a held note (as env_stage_held), or release / silence after the note-off.
__device__ __forceinline__ void env_stage_any(const EnvShape& sh, i64 n_on, i64 n_off, double l_on, double l_off, i64 f,
                                              double& q0, double& q1, double& q2, double& w0, double& dw) {
  if (f < n_off) {
    env_stage_held(sh, n_on, l_on, f, q0, q1, q2, w0, dw);
    return;
  }
  const i64 k = f - n_off;
  q0 = 0.0; q1 = 0.0; q2 = 0.0; w0 = 0.0; dw = 0.0;
  if (k < sh.nr) {  // release: l_off * u^2
    w0 = 1.0 - (double)k * sh.inv_nr; dw = -sh.inv_nr;
    q2 = l_off;
  }
}

// One block of a voice whose cutoff is constant (welsh_rest_block for one voice, with the amplitude
// envelope inside one stage instead of resting).  L / o1 / o2: the instrument's resting tables (lti,
// m1bb, m2bb) or those of a released voice (lti_off, ...).  t0 = the block's first frame relative to
// the item start.
template <bool LFO_AMP>
__device__ __forceinline__ void welsh_solo_rest_block(SoloRestState* rs, const WelshInst& I, const LtiTable& L,
                                                      const OscMix& o1, const OscMix& o2, int lane, int t0,
                                                      double2* tile_row) {
  double yp[kT];
  double ps0 = 0.0, ps1 = 0.0;
  {
    const ulonglong2 pp = *reinterpret_cast<const ulonglong2*>(&rs->p1);
    const ulonglong2 dd = *reinterpret_cast<const ulonglong2*>(&rs->d1);
    const u64 k = (u64)(lane * kT);
    u64 p1 = pp.x + k * dd.x, p2 = pp.y + k * dd.y;
    __syncwarp();  // every lane has read the block's base phases
    if (lane == 0)
      *reinterpret_cast<ulonglong2*>(&rs->p1) =
          make_ulonglong2(p1 + (u64)kBlockFrames * dd.x, p2 + (u64)kBlockFrames * dd.y);
    const u64 t1 = I.s1.thresh, t2 = I.s2.thresh;
    const double a1 = L.c1.a1, a2 = L.c1.a2;
#pragma unroll
    for (int j = 0; j < kT; ++j) {
      p1 += dd.x;
      p2 += dd.y;
      yp[j] = lp_step_bx(osc_mix_eval<false>(o1, t1, p1, o2, t2, p2), a1, a2, ps0, ps1);
    }
    ps0 *= L.inv_b0_2; ps1 *= L.inv_b0_2;
  }
  double e0, e1;
  {
    const double v0[1] = {ps0}, v1[1] = {ps1};
    double x0[1], x1[1];
    SoloRestState* const one[1] = {rs};
    lti_scan_entry_t<1, SoloRestState>(v0, v1, L.mp1, lane, one, 0, x0, x1);
    e0 = x0[0]; e1 = x1[0];
  }
  {
    const double a1 = L.c2.a1, a2 = L.c2.a2;
    ps0 = 0.0; ps1 = 0.0;
#pragma unroll
    for (int j = 0; j < kT; ++j) {
      const double2 g = *reinterpret_cast<const double2*>(L.g1b[j]);
      yp[j] = lp_step_bx(fma(g.y, e1, fma(g.x, e0, yp[j])), a1, a2, ps0, ps1);
    }
  }
  {
    const double v0[1] = {ps0}, v1[1] = {ps1};
    double x0[1], x1[1];
    SoloRestState* const one[1] = {rs};
    lti_scan_entry_t<1, SoloRestState>(v0, v1, L.mp2, lane, one, 1, x0, x1);
    e0 = x0[0]; e1 = x1[0];
  }
  double lsd = 0.0, lcd = 0.0;
  if (LFO_AMP) {
    const double2 ph = *reinterpret_cast<const double2*>(&rs->ls);
    const double2 r = I.lane_rot[lane];
    lsd = fma(ph.x, r.x, ph.y * r.y);
    lcd = fma(ph.y, r.x, -(ph.x * r.y));
    __syncwarp();
    if (lane == 0) {
      const double2 br = I.block_rot;
      *reinterpret_cast<double2*>(&rs->ls) = make_double2(fma(lsd, br.x, lcd * br.y), fma(lcd, br.x, -(lsd * br.y)));
    }
  }
  const double aq0 = rs->aq0, aq1 = rs->aq1, aq2 = rs->aq2, aw = rs->aw, adw = rs->adw;
  const double gl = I.gl, gr = I.gr;
  const int tl = t0 + lane * kT;
  double2* row = tile_row + lane * (kT + 1);
#pragma unroll
  for (int j = 0; j < kT; ++j) {
    const double2 g = *reinterpret_cast<const double2*>(L.g2[j]);
    const double w = fma((double)(tl + j), adw, aw);
    double amp = fma(w, fma(aq2, w, aq1), aq0);
    if (LFO_AMP) {
      const double2 rot = I.lfo_rot[j];
      amp *= fma(lsd, rot.x, fma(lcd, rot.y, 1.0));
    }
    const double m = fma(g.y, e1, fma(g.x, e0, yp[j])) * amp;
    row[j] = make_double2(m * gl, m * gr);
  }
  __syncwarp();
}

// One 256-frame block of one solo voice through the per-block classified paths of welsh_kernel (its
// SOLO = true body): time-invariant / fast / general.  Returns false when the voice is idle for the block.
__device__ __forceinline__ bool welsh_solo_general_block(const WelshInst& I, const WelshInst* gI, WelshVoice* voices,
                                                         int vi, const VoiceEvent* __restrict__ events,
                                                         const int* __restrict__ ev_off, i64 fb, i64 f_end, int lane,
                                                         double2* tile_row, double* park) {
  const bool pitch = I.routing == LFO_PITCH;
  const bool lin_inst = I.s1.kind == 0 && I.s2.kind == 0 && !I.sync &&
                        (I.routing == LFO_NONE || (I.routing == LFO_AMPLITUDE && I.wl == W_SINE));
  const bool simple_inst = lin_inst && I.filter_mode == FILTER_ENVELOPE;
  const bool lti_inst = lin_inst && I.lti_ok && (I.filter_mode == FILTER_FIXED || I.filter_mode == FILTER_ENVELOPE);
  WelshVoice* vp = voices + vi;
  NoteWords st;
  st.n_on = vp->n_on; st.n_off = vp->n_off;
  st.la_on = vp->la_on; st.la_off = vp->la_off; st.lf_on = vp->lf_on; st.lf_off = vp->lf_off;
  int ei = ev_off[vi];
  const int e_end = ev_off[vi + 1];
  while (ei < e_end && events[ei].frame < fb) ++ei;  // already folded into the record by earlier blocks
  const bool idle = fb >= st.n_off + I.amp.nr;
  const bool ev_here = ei < e_end && events[ei].frame < fb + kBlockFrames;
  if (idle && !ev_here) return false;
  const int local = vi - I.voice0;
  auto seed_of = [&](u64 salt) { return splitmix64(((u64)(unsigned)I.uid << 32) ^ salt); };
  bool fast = false;
  EnvSeg aseg, fseg;
  int cls = 0;
  if (!pitch && !ev_here) {
    cls = welsh_lane_class(st, I, fb + (i64)lane * kT, f_end, aseg, fseg);
    fast = __all_sync(0xffffffffu, cls != 0);
  }
  bool lti = false;
  if (fast && lti_inst) {
    const bool rest = cls == 2 && (I.filter_mode == FILTER_FIXED ||
                                   (fseg.q1 == 0.0 && fseg.q2 == 0.0 && fseg.q0 == I.filt.sustain));
    lti = __all_sync(0xffffffffu, rest);
  }
  if (lti || fast) {  // the specialised blocks accumulate into the warp's tile row
    double2* row = tile_row + lane * (kT + 1);
#pragma unroll
    for (int j = 0; j < kT; ++j) row[j] = make_double2(0.0, 0.0);
  }
  if (lti) {
    if (I.routing == LFO_NONE) welsh_block_lti_ool<false, false, false>(vp, &I, fb, lane, &aseg, tile_row);
    else welsh_block_lti_ool<true, false, false>(vp, &I, fb, lane, &aseg, tile_row);
  } else if (fast) {
    const WelshInst* Ip = &I;
    if (I.filter_mode == FILTER_FIXED) {
      if (__all_sync(0xffffffffu, cls == 2))
        welsh_fast_call<COEF_FIXED, true>(vp, Ip, fb, lane, cls, aseg, fseg, seed_of((u64)(2 * local)), seed_of((u64)(2 * local + 1)), seed_of(0x4C464F00ull ^ (u64)local), tile_row, true);
      else
        welsh_fast_call<COEF_FIXED, false>(vp, Ip, fb, lane, cls, aseg, fseg, seed_of((u64)(2 * local)), seed_of((u64)(2 * local + 1)), seed_of(0x4C464F00ull ^ (u64)local), tile_row, false);
    } else {
      bool smooth = false;
      if (I.filter_mode == FILTER_ENVELOPE && cls == 2) {
        const double w8 = fma((double)kT, fseg.dw, fseg.w0);
        const double r0 = fabs(fma(2.0 * fseg.q2, fseg.w0, fseg.q1)), r8 = fabs(fma(2.0 * fseg.q2, w8, fseg.q1));
        smooth = fabs(I.cut_b * fseg.dw) * fmax(r0, r8) <= I.knot_max_rate && I.knot_max_rate > 0.0;
      }
      if (__all_sync(0xffffffffu, smooth)) {
        if (simple_inst) {
          if (I.routing == LFO_NONE) welsh_block_simple_ool<false, false>(vp, Ip, fb, lane, &aseg, &fseg, tile_row, park);
          else welsh_block_simple_ool<true, false>(vp, Ip, fb, lane, &aseg, &fseg, tile_row, park);
        } else {
          welsh_fast_call<COEF_KNOTS, true>(vp, Ip, fb, lane, cls, aseg, fseg, seed_of((u64)(2 * local)), seed_of((u64)(2 * local + 1)), seed_of(0x4C464F00ull ^ (u64)local), tile_row, false);
        }
      } else {
        welsh_fast_call<COEF_EXACT, false>(vp, Ip, fb, lane, cls, aseg, fseg, seed_of((u64)(2 * local)), seed_of((u64)(2 * local + 1)), seed_of(0x4C464F00ull ^ (u64)local), tile_row, false);
      }
    }
  } else {
    welsh_block_general((pitch ? 2 : 0) + (ev_here ? 1 : 0), vp, gI, events, ei, e_end, fb, f_end, lane,
                        seed_of((u64)(2 * local)), seed_of((u64)(2 * local + 1)), seed_of(0x4C464F00ull ^ (u64)local), tile_row, false);
  }
  return true;
}

__device__ __forceinline__ int ld_acquire_i32(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_i32(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Shared-memory layout of welsh_solo_kernel (all dynamic, so that the out-of-line class bodies below address
// it as shared memory): W tile rows | kParkWords x 32 W parking doubles | W state slots | W instrument records.
constexpr int kSoloStateBytes = 256;  // per-warp slot for SoloRestState / SweepState
template <int W>
struct SoloSmem {
  static constexpr size_t kPark = (size_t)W * kTileStride * sizeof(double2);
  static constexpr size_t kSlots = kPark + (size_t)kParkWords * 32 * W * sizeof(double);
  static constexpr size_t kInsts = kSlots + (size_t)W * kSoloStateBytes;
  static constexpr size_t kBytes = kInsts + (size_t)W * sizeof(WelshInst);
  static __device__ __forceinline__ unsigned char* base() {
    extern __shared__ double2 smem_tiles[];
    return reinterpret_cast<unsigned char*>(smem_tiles);
  }
  static __device__ __forceinline__ double2* tile_row(int warp) {
    return reinterpret_cast<double2*>(base()) + warp * kTileStride;
  }
  static __device__ __forceinline__ double* park() { return reinterpret_cast<double*>(base() + kPark) + threadIdx.x; }
  static __device__ __forceinline__ unsigned char* slot(int warp) { return base() + kSlots + (size_t)warp * kSoloStateBytes; }
  static __device__ __forceinline__ WelshInst* inst(int warp) { return reinterpret_cast<WelshInst*>(base() + kInsts) + warp; }
};

// Phases, filter state, LFO phasor and amplitude stage of a voice at the first frame fs of an item
// (the common head of SoloRestState and SweepState), written by lane 0 straight into the warp's slot.
template <typename State>
__device__ __forceinline__ void solo_state_head(State* rs, const WelshInst& I, const WelshVoice* vp, i64 fs) {
  const u64 k = (u64)(fs - 1 - vp->anchor);
  const u64 d1 = vp->d1, d2 = vp->d2;
  rs->d1 = d1; rs->d2 = d2;
  rs->p1 = vp->p1 + k * d1; rs->p2 = vp->p2 + k * d2;
  rs->s[0] = vp->s[0]; rs->s[1] = vp->s[1]; rs->s[2] = vp->s[2]; rs->s[3] = vp->s[3];
  double ls = 0.0, lc = 0.0;
  if (I.routing == LFO_AMPLITUDE) {
    sincos_phase(vp->pl + (k + 1) * I.lfo_dq, &ls, &lc);
    ls *= I.depth; lc *= I.depth;
  }
  rs->ls = ls; rs->lc = lc;
  double q0, q1, q2, w, dw;
  env_stage_any(I.amp, vp->n_on, vp->n_off, vp->la_on, vp->la_off, fs, q0, q1, q2, w, dw);
  rs->aq0 = 0.5 * q0; rs->aq1 = 0.5 * q1; rs->aq2 = 0.5 * q2; rs->aw = w; rs->adw = dw;
}

// CTA barrier of a job's block loop: a named barrier (id 1), so that warps without an item can arrive from
// their own loop (solo_idle_item) while the others arrive from inside a class body.
template <int W>
__device__ __forceinline__ void solo_job_bar(bool lockstep) {
  if (lockstep) asm volatile("barrier.sync 1, %0;" ::"n"(32 * W) : "memory");
  else __syncwarp();
}
template <int W>
__device__ __noinline__ void solo_idle_item(int nframes) {
#pragma unroll 1
  for (int t = 0; t < nframes; t += kBlockFrames) solo_job_bar<W>(true);
}

// The three class bodies are out-of-line: each gets its own register allocation and the dispatcher stays
// a few dozen instructions.  They find their warp's tile row, state slot and instrument record in the
// kernel's dynamic shared memory.
template <int W>
__device__ __noinline__ void solo_rest_item(WelshVoice* vp, i64 fs, int nframes, double2* out, bool lockstep) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const WelshInst& I = *SoloSmem<W>::inst(warp);
  double2* tile_row = SoloSmem<W>::tile_row(warp);
  SoloRestState* rs = reinterpret_cast<SoloRestState*>(SoloSmem<W>::slot(warp));
  const bool released = fs >= vp->n_off;
  if (lane == 0) solo_state_head(rs, I, vp, fs);
  __syncwarp();
  const LtiTable& L = released ? I.lti_off : I.lti;
  const OscMix& o1 = released ? I.m1bb_off : I.m1bb;
  const OscMix& o2 = released ? I.m2bb_off : I.m2bb;
  const bool lfo = I.routing == LFO_AMPLITUDE;
  const i64 fe = fs + nframes;
  int t0 = 0;
#pragma unroll 1
  for (i64 fb = fs; fb < fe; fb += kBlockFrames, t0 += kBlockFrames) {
    if (lfo) welsh_solo_rest_block<true>(rs, I, L, o1, o2, lane, t0, tile_row);
    else welsh_solo_rest_block<false>(rs, I, L, o1, o2, lane, t0, tile_row);
    warp_store_row(tile_row, true, out, fb, fs, fe, lane);
    solo_job_bar<W>(lockstep);
  }
  if (lane == 0) {
    vp->s[0] = rs->s[0]; vp->s[1] = rs->s[1]; vp->s[2] = rs->s[2]; vp->s[3] = rs->s[3];
    vp->knot_frame = kNever;
  }
  __syncwarp();
}

template <int W, bool EXACT>
__device__ __noinline__ void solo_sweep_item(WelshVoice* vp, i64 fs, int nframes, double2* out, bool lockstep) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const WelshInst& I = *SoloSmem<W>::inst(warp);
  double2* tile_row = SoloSmem<W>::tile_row(warp);
  SweepState* rs = reinterpret_cast<SweepState*>(SoloSmem<W>::slot(warp));
  if (lane == 0) {
    solo_state_head(rs, I, vp, fs);
    double q0, q1, q2, w0, dw;
    env_stage_any(I.filt, vp->n_on, vp->n_off, vp->lf_on, vp->lf_off, fs, q0, q1, q2, w0, dw);
    rs->fq0 = q0; rs->fq1 = q1; rs->fq2 = q2; rs->fw = w0; rs->fdw = dw;
    if (!EXACT) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {  // knots at fs - kT and fs: the stage's formula, continued backwards
        const double w = fma((double)((q - 1) * kT), dw, w0);
        welsh_knot(I, fma(I.cut_b, fma(w, fma(q2, w, q1), q0), I.cut_a), rs->kn[q]);
      }
    }
  }
  __syncwarp();
  const bool lfo = I.routing == LFO_AMPLITUDE;
  const i64 fe = fs + nframes;
  double* park = SoloSmem<W>::park();
  int t0 = 0;
#pragma unroll 1
  for (i64 fb = fs; fb < fe; fb += kBlockFrames, t0 += kBlockFrames) {
    if (EXACT) {
      if (lfo) welsh_exact_block<true>(rs, I, lane, t0, tile_row, park, 32 * W);
      else welsh_exact_block<false>(rs, I, lane, t0, tile_row, park, 32 * W);
    } else if (lfo) welsh_sweep_block<true, false, false>(rs, I, lane, t0, tile_row);
    else welsh_sweep_block<false, false, false>(rs, I, lane, t0, tile_row);
    warp_store_row(tile_row, true, out, fb, fs, fe, lane);
    solo_job_bar<W>(lockstep);
  }
  if (lane == 0) {
    vp->s[0] = rs->s[0]; vp->s[1] = rs->s[1]; vp->s[2] = rs->s[2]; vp->s[3] = rs->s[3];
    vp->knot_frame = kNever;
  }
  __syncwarp();
}

template <int W>
__device__ __noinline__ void solo_general_item(const WelshInst* gI, WelshVoice* voices, int vi,
                                               const VoiceEvent* __restrict__ events, const int* __restrict__ ev_off,
                                               i64 fs, int nframes, i64 f_end, double2* out, bool lockstep) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const WelshInst& I = *SoloSmem<W>::inst(warp);
  double2* tile_row = SoloSmem<W>::tile_row(warp);
  double* park = SoloSmem<W>::park();
  const i64 fe = fs + nframes;
#pragma unroll 1
  for (i64 fb = fs; fb < fe; fb += kBlockFrames) {
    const bool any = welsh_solo_general_block(I, gI, voices, vi, events, ev_off, fb, f_end, lane, tile_row, park);
    __syncwarp();
    warp_store_row(tile_row, any, out, fb, fs, fe, lane);
    solo_job_bar<W>(lockstep);
  }
}

__device__ __forceinline__ int ld_relaxed_i32(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// One item of a job: make the instrument record resident, wait for the voice's previous item, run the class
// body, publish the voice's progress.  The wait is executed by all 32 lanes in step (same address, the value
// taken from lane 0), so the warp never diverges around the spin; it polls with relaxed loads and takes
// one acquire load at the end (an acquire invalidates the SM's L1: once per item, not once per poll).
// Returns the resident instrument.
template <int W>
__device__ __noinline__ int solo_run_item(int j, SoloJob job, int cur_inst, const WelshInst* __restrict__ insts,
                                          WelshVoice* __restrict__ voices, const WarpItem* __restrict__ items,
                                          const SoloItem* __restrict__ sitems, int* __restrict__ progress,
                                          int* __restrict__ fault, const VoiceEvent* __restrict__ events,
                                          const int* __restrict__ ev_off, i64 f0, i64 f_chunk_end) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const SoloItem si = sitems[job.first + warp];
  const WarpItem item = items[si.item];
  if (cur_inst != item.inst) {  // instrument record -> shared memory (kept while the warp stays on the instrument)
    const int4* src = reinterpret_cast<const int4*>(insts + item.inst);
    int4* dst = reinterpret_cast<int4*>(SoloSmem<W>::inst(warp));
    for (int i = lane; i < (int)(sizeof(WelshInst) / sizeof(int4)); i += 32) dst[i] = src[i];
  }
  {
    // watchdog: a dependency that does not resolve within ~0.5 s of SM clocks is a scheduling bug; report it
    // (job, voice, need, have) through `fault` and carry on, so that the host fails loudly instead of hanging
    const int* pp = progress + item.voice;
    const long long t_wait = clock64();
    for (;;) {
      const int have = __shfl_sync(0xffffffffu, ld_relaxed_i32(pp), 0);
      if (have >= si.need) break;
      __nanosleep(250);
      const int late = __shfl_sync(0xffffffffu, (int)(clock64() - t_wait > 1000000000ll), 0);
      if (late) {
        if (lane == 0 && atomicCAS(fault, 0, 1) == 0) { fault[1] = j; fault[2] = item.voice; fault[3] = si.need; fault[4] = have; }
        break;
      }
    }
    (void)ld_acquire_i32(pp);
  }
  __syncwarp();
  WelshVoice* vp = voices + item.voice;
  const i64 fs = f0 + job.t0;        // first frame of the item
  double2* out = item.out + job.t0;  // the voice's output for this sub-chunk
  if (job.cls == SOLO_REST) {
    solo_rest_item<W>(vp, fs, job.nframes, out, job.lockstep != 0);
  } else if (job.cls == SOLO_SWEEP) {
    solo_sweep_item<W, false>(vp, fs, job.nframes, out, job.lockstep != 0);
  } else if (job.cls == SOLO_EXACT) {
    solo_sweep_item<W, true>(vp, fs, job.nframes, out, job.lockstep != 0);
  } else {
    const i64 fe = fs + job.nframes;
    solo_general_item<W>(insts + item.inst, voices, item.voice, events, ev_off, fs, job.nframes,
                         fe < f_chunk_end ? fe : f_chunk_end, out, job.lockstep != 0);
  }
  __syncwarp();
  if (lane == 0) {
    __threadfence();
    st_release_i32(progress + item.voice, si.need + 1);
  }
  __syncwarp();
  return item.inst;
}

// grid = min(jobs, resident CTAs); block = 32 * W threads; dynamic smem = SoloSmem<W>::kBytes.
// Jobs [job0, job0 + n_jobs) are handed out through *ticket (zero at launch).
template <int W>
__global__ void __launch_bounds__(32 * W, 2) welsh_solo_kernel(const WelshInst* __restrict__ insts,
                                                             WelshVoice* __restrict__ voices,
                                                             const WarpItem* __restrict__ items,
                                                             const SoloItem* __restrict__ sitems,
                                                             const SoloJob* __restrict__ jobs, int job0, int n_jobs,
                                                             int* __restrict__ ticket, int* __restrict__ progress,
                                                             int* __restrict__ fault,
                                                             const VoiceEvent* __restrict__ events,
                                                             const int* __restrict__ ev_off, i64 f0, int chunk_frames) {
  static_assert(sizeof(SoloRestState) <= kSoloStateBytes && sizeof(SweepState) <= kSoloStateBytes, "state slot too small");
  static_assert(sizeof(WelshInst) % 16 == 0, "instrument records are copied and aligned as 16-byte words");
  __shared__ int s_job;
  const int warp = threadIdx.x >> 5;
  const i64 f_chunk_end = f0 + chunk_frames;
  int cur_inst = -1;
  for (;;) {
    __syncthreads();  // every warp is done with the previous job (and has read s_job)
    if (threadIdx.x == 0) s_job = atomicAdd(ticket, 1);
    __syncthreads();
    const int j = s_job;
    if (j >= n_jobs) break;
    const SoloJob job = jobs[job0 + j];
    if (warp < job.n)
      cur_inst = solo_run_item<W>(job0 + j, job, cur_inst, insts, voices, items, sitems, progress, fault, events, ev_off, f0,
                                  f_chunk_end);
    else if (job.lockstep)
      solo_idle_item<W>(job.nframes);
  }
}

// --------------------------------------------------------------------------- FM ---
struct FmInst {
  EnvShape car, mod;
  double depth, beta;
  double gl, gr;
  int uid, voice0;
};
struct FmVoice {
  i64 n_on, n_off;
  double lc_on, lc_off, lm_on, lm_off;
  i64 anchor;   // modulator phase anchor
  u64 pm, dm;   // modulator phase at anchor, increment
  u64 pc;       // carrier phase at (chunk position - 1)
  double cyc_c; // carrier base cycles per frame
  double pad;
};

__device__ __noinline__ FmVoice fm_fold(FmVoice st, const FmInst* Ip, VoiceEvent ev, bool* restart) {
  const FmInst& I = *Ip;
  i64 f = ev.frame;
  *restart = false;
  if (ev.type == VEV_NOTE_ON) {
    bool was = f >= st.n_on && f < st.n_off + I.car.nr;
    double lc = 0.0, lm = 0.0;
    if (was) {
      lc = env_level(I.car, st.n_on, st.n_off, st.lc_on, st.lc_off, f);
      lm = env_level(I.mod, st.n_on, st.n_off, st.lm_on, st.lm_off, f);
      st.pm += (u64)(f - 1 - st.anchor) * st.dm;
      st.anchor = f - 1;
    } else {
      st.pm = 0;
      st.anchor = f;
    }
    st.lc_on = lc; st.lm_on = lm;
    st.n_on = f; st.n_off = kHeld;
    st.cyc_c = ev.cyc1;
    st.dm = cycles_to_q(ev.cyc2);
    *restart = !was;
  } else {
    st.lc_off = env_pre(I.car, st.n_on, st.lc_on, f);
    st.lm_off = env_pre(I.mod, st.n_on, st.lm_on, f);
    st.n_off = f;
  }
  return st;
}

template <bool EV>
__device__ __forceinline__ void fm_block(FmVoice& st, const FmInst& I, const VoiceEvent* __restrict__ ev, int& ei,
                                         int e_end, i64 fb, i64 f_end, int lane, double2* tile_row,
                                         bool accumulate) {
  const i64 c0 = fb + (i64)lane * kT;
  const i64 blk_end = fb + kBlockFrames;
  FmVoice ls = st;
  int li = ei;
  i64 next_ev = kHeld;
  bool rs_dummy;
  if (EV) {
    while (li < e_end && ev[li].frame < c0) {
      ls = fm_fold(ls, &I, ev[li], &rs_dummy);
      ++li;
    }
    next_ev = li < e_end ? ev[li].frame : kHeld;
  }
  u64 dc[kT];
  double cenv[kT];
  unsigned play_bits = 0, reset_bits = 0;
  SegSum a; a.sum = 0; a.reset = 0;
#pragma unroll
  for (int j = 0; j < kT; ++j) {
    i64 n = c0 + j;
    if (EV) {
      while (n == next_ev) {
        bool rs;
        ls = fm_fold(ls, &I, ev[li], &rs);
        if (rs) reset_bits |= 1u << j;
        ++li;
        next_ev = li < e_end ? ev[li].frame : kHeld;
      }
    }
    bool play = n >= ls.n_on && n < ls.n_off + I.car.nr && n < f_end;
    dc[j] = 0; cenv[j] = 0.0;
    if (play) {
      play_bits |= 1u << j;
      u64 pm = ls.pm + (u64)(n - ls.anchor) * ls.dm;
      double mval = sinpi(2.0 * pos_of(pm));
      double menv = env_level(I.mod, ls.n_on, ls.n_off, ls.lm_on, ls.lm_off, n);
      double x = __dmul_rn(__dmul_rn(__dmul_rn(mval, menv), I.depth), I.beta);
      dc[j] = cycles_to_q(__dmul_rn(ls.cyc_c, __dadd_rn(1.0, x)));
      cenv[j] = env_level(I.car, ls.n_on, ls.n_off, ls.lc_on, ls.lc_off, n);
      if (reset_bits & (1u << j)) { a.sum = 0; a.reset = 1; }
      else a.sum += dc[j];
    }
  }
  SegSum inc = segsum_warp_scan(a, lane);
  SegSum ex;
  ex.sum = __shfl_up_sync(0xffffffffu, inc.sum, 1);
  ex.reset = __shfl_up_sync(0xffffffffu, inc.reset, 1);
  if (lane == 0) { ex.sum = 0; ex.reset = 0; }
  u64 pc = ex.reset ? ex.sum : st.pc + ex.sum;
  u64 tot = inc.reset ? inc.sum : st.pc + inc.sum;
  st.pc = __shfl_sync(0xffffffffu, tot, 31);
#pragma unroll
  for (int j = 0; j < kT; ++j) {
    double m = 0.0;
    if (play_bits & (1u << j)) {
      if (reset_bits & (1u << j)) pc = 0;
      else pc += dc[j];
      m = sin_phase(pc) * cenv[j];
    }
    int t = lane * kT + j;
    double2 o = make_double2(m * I.gl, m * I.gr);
    if (accumulate) {
      double2 p = tile_row[t + (t >> 3)];
      o.x += p.x; o.y += p.y;
    }
    tile_row[t + (t >> 3)] = o;
  }
  if (EV) {
    while (ei < e_end && ev[ei].frame < blk_end) {
      st = fm_fold(st, &I, ev[ei], &rs_dummy);
      ++ei;
    }
  }
}

// ---- FM fast path: blocks without note events in which every lane is idle or sounds inside single
// stages of both envelopes.  Envelopes are 3 FMAs per frame; the modulator sine advances by an
// angle-addition rotation (its phase increment is constant within a note); the carrier is exact:
// integer-summed phase (warp scan) and sinpi.
__device__ __forceinline__ int fm_lane_class(const FmVoice& st, const FmInst& I, i64 c0, i64 f_end, EnvSeg& car,
                                             EnvSeg& mod) {
  const i64 last = c0 + (kT - 1);
  const i64 idle_at = st.n_off + I.car.nr;
  if (c0 >= f_end || last < st.n_on || c0 >= idle_at) return 1;
  if (!(c0 >= st.n_on && last < idle_at && last < f_end)) return 0;
  if (!env_segment(I.car, st.n_on, st.n_off, st.lc_on, st.lc_off, c0, car)) return 0;
  if (!env_segment(I.mod, st.n_on, st.n_off, st.lm_on, st.lm_off, c0, mod)) return 0;
  return 2;
}
__device__ __forceinline__ void fm_block_fast(FmVoice& st, const FmInst& I, i64 fb, int lane, int cls,
                                              const EnvSeg& cseg, const EnvSeg& mseg, double2* tile_row,
                                              bool accumulate) {
  const i64 c0 = fb + (i64)lane * kT;
  const bool on = cls == 2;
  u64 dc[kT];
  u64 sum = 0;
#pragma unroll
  for (int j = 0; j < kT; ++j) dc[j] = 0;
  if (on) {
    const u64 pm0 = st.pm + (u64)(c0 - st.anchor) * st.dm;
    double s, c, sd, cd;
    sincospi(2.0 * pos_of(pm0), &s, &c);
    sincospi(2.0 * (__ull2double_rn(st.dm) * (1.0 / kTwo64)), &sd, &cd);
    const double depth = I.depth, beta = I.beta, cyc = st.cyc_c;
#pragma unroll
    for (int j = 0; j < kT; ++j) {
      const double menv = env_seg_at(mseg, j);
      const double x = __dmul_rn(__dmul_rn(__dmul_rn(s, menv), depth), beta);
      dc[j] = cycles_to_q(__dmul_rn(cyc, __dadd_rn(1.0, x)));
      sum += dc[j];
      const double ns = fma(s, cd, c * sd);
      c = fma(c, cd, -(s * sd));
      s = ns;
    }
  }
  // inclusive integer scan of the lane sums (no restarts on this path)
  u64 inc = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const u64 up = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += up;
  }
  u64 pc = st.pc + inc - sum;  // carrier phase at c0 - 1
  st.pc = __shfl_sync(0xffffffffu, st.pc + inc, 31);
#pragma unroll
  for (int j = 0; j < kT; ++j) {
    double m = 0.0;
    if (on) {
      pc += dc[j];
      m = sin_phase(pc) * env_seg_at(cseg, j);
    }
    const int t = lane * kT + j;
    double2 o = make_double2(m * I.gl, m * I.gr);
    if (accumulate) {
      double2 p = tile_row[t + (t >> 3)];
      o.x += p.x; o.y += p.y;
    }
    tile_row[t + (t >> 3)] = o;
  }
}

template <int W>
__global__ void __launch_bounds__(32 * W, 2) fm_kernel(const FmInst* __restrict__ insts, FmVoice* __restrict__ voices,
                                                     const CtaWork* __restrict__ work,
                                                     const WarpItem* __restrict__ items,
                                                     const VoiceEvent* __restrict__ events,
                                                     const int* __restrict__ ev_off, i64 f0, int nframes) {
  extern __shared__ double2 smem_tiles[];
  __shared__ int s_active[W];
  const CtaWork wk = work[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool solo = wk.solo != 0;
  const bool mine = !solo || warp < wk.nvoices;
  WarpItem item;
  item.inst = wk.inst; item.voice = 0; item.out = nullptr;
  if (solo && mine) item = items[wk.voice0 + warp];
  const FmInst& I = insts[item.inst];
  double2* tile_row = smem_tiles + warp * kTileStride;
  const i64 f_end = f0 + nframes;
  const int g_begin = solo ? 0 : warp;
  const int g_end = solo ? (mine ? 1 : 0) : wk.nvoices;
  // one voice per warp at most (solo items, instruments of up to W voices): the voice record and the event
  // cursor stay in registers for the whole chunk instead of going through global memory every block
  const bool resident = solo || wk.nvoices <= W;
  FmVoice st;
  int ei = 0, e_end = 0;
  if (resident && g_begin < g_end) {
    const int vi = solo ? item.voice : wk.voice0 + g_begin;
    st = voices[vi];
    ei = ev_off[vi];
    e_end = ev_off[vi + 1];
  }
#pragma unroll 1
  for (i64 fb = f0; fb < f_end; fb += kBlockFrames) {
    bool any = false;
#pragma unroll 1
    for (int g = g_begin; g < g_end; g += W) {
      const int vi = solo ? item.voice : wk.voice0 + g;
      if (!resident) {
        st = voices[vi];
        ei = ev_off[vi];
        e_end = ev_off[vi + 1];
      }
      while (ei < e_end && events[ei].frame < fb) ++ei;
      bool idle = fb >= st.n_off + I.car.nr;
      bool ev_here = ei < e_end && events[ei].frame < fb + kBlockFrames;
      if (idle && !ev_here) continue;
      if (ev_here) {
        fm_block<true>(st, I, events, ei, e_end, fb, f_end, lane, tile_row, any);
      } else {
        EnvSeg cseg, mseg;
        const int cls = fm_lane_class(st, I, fb + (i64)lane * kT, f_end, cseg, mseg);
        if (__all_sync(0xffffffffu, cls != 0)) fm_block_fast(st, I, fb, lane, cls, cseg, mseg, tile_row, any);
        else fm_block<false>(st, I, events, ei, e_end, fb, f_end, lane, tile_row, any);
      }
      any = true;
      if (!resident) {
        __syncwarp();
        if (lane == 0) voices[vi] = st;
        __syncwarp();
      }
    }
    if (solo) {
      __syncwarp();
      if (mine) warp_store_row(tile_row, any, item.out, fb, f0, f_end, lane);
      __syncwarp();
    } else {
      if (lane == 0) s_active[warp] = any ? 1 : 0;
      __syncthreads();
      cta_reduce_store<W>(smem_tiles, s_active, wk.out, fb, f0, f_end);
      __syncthreads();
    }
  }
  if (resident && g_begin < g_end && lane == 0) voices[solo ? item.voice : wk.voice0 + g_begin] = st;
}

// ------------------------------------------------------------ sampler / drumkit ---
// The host tracks every "play" (a voice sounding one sample from n_on until n_end) with integer
// frame arithmetic, so the device needs no persistent state: one thread per output frame sums the
// plays that cover it, in voice order.
struct SamplePlay {
  i64 n_on, n_end;
  u64 step_q;       // 32.32 sample frames per output frame
  const double* data;
  u64 len;          // frames
  int channels;
  int pad;
};

__global__ void __launch_bounds__(256) sampler_kernel(const SamplePlay* __restrict__ plays, int nplays,
                                                       double2* __restrict__ out, i64 f0, int nframes) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nframes) return;
  i64 n = f0 + t;
  double l = 0.0, r = 0.0;
  for (int p = 0; p < nplays; ++p) {
    SamplePlay pl = plays[p];
    if (n >= pl.n_on && n < pl.n_end) {
      u64 idx = ((u64)(n - pl.n_on) * pl.step_q) >> 32;
      if (idx < pl.len) {
        if (pl.channels == 2) {
          double2 v = reinterpret_cast<const double2*>(pl.data)[idx];
          l += v.x; r += v.y;
        } else {
          double v = __ldg(pl.data + idx);
          l += v; r += v;
        }
      }
    }
  }
  out[t] = make_double2(l, r);
}

}  // namespace gbk
