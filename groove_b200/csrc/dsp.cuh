// dsp.cuh — device-side DSP primitives shared by the voice and effect kernels.
//
// Formulas are the ones written down in docs/ORACLE_SPEC.md.  Quantities that
// decide a discontinuity (oscillator phase, pulse duty, envelope stage, sample
// index) are integers so that the closed-form / scanned evaluation here is
// bit-identical to a frame-by-frame accumulation; smooth quantities are f64.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace gbk {

typedef unsigned long long u64;
typedef long long i64;

constexpr i64 kHeld = 1ll << 60;
constexpr i64 kNever = -(1ll << 60);
constexpr double kTwo64 = 18446744073709551616.0;
constexpr double kLog2_800 = 9.6438561897747243;
constexpr double kPi = 3.141592653589793238462643383279;

// ---- waveforms (gb_waveform) ----
enum { W_NONE = 0, W_SINE, W_SQUARE, W_PULSE, W_TRIANGLE, W_SAW, W_NOISE, W_DZERO, W_DMAX, W_DMIN };
enum { LFO_NONE = 0, LFO_AMPLITUDE, LFO_PITCH, LFO_PULSE_WIDTH, LFO_FILTER_CUTOFF };

__host__ __device__ __forceinline__ u64 splitmix64(u64 x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// fractional cycles -> 2^64-scaled phase increment
__device__ __forceinline__ u64 cycles_to_q(double c) {
  c -= floor(c);
  double r = c * kTwo64;
  if (!(r < kTwo64)) return 0ull;
  return __double2ull_rz(r);
}
// phase -> position in [0,1): the top 52 bits, exactly (q >> 12) * 2^-52.  Built by placing them in
// the mantissa of a double in [1,2) — integer ops plus one DADD instead of a 64-bit I2F on the XU pipe.
__device__ __forceinline__ double pos_of(u64 q) {
  return __longlong_as_double((long long)(0x3FF0000000000000ull | (q >> 12))) - 1.0;
}

// Oscillator output for phase q.  `wf` is warp-uniform.
__device__ __forceinline__ double wave_value(int wf, u64 q, u64 duty_q, u64 seed, i64 frame) {
  const u64 half = 1ull << 63;
  switch (wf) {
    case W_SINE: return sinpi(2.0 * pos_of(q));
    case W_SQUARE: return q < half ? 1.0 : -1.0;
    case W_PULSE: return q < duty_q ? 1.0 : -1.0;
    case W_TRIANGLE: {
      double p = pos_of(q);
      return q < half ? 4.0 * p - 1.0 : 3.0 - 4.0 * p;
    }
    case W_SAW: {
      double p = pos_of(q);
      return q < half ? 2.0 * p : 2.0 * p - 2.0;
    }
    case W_NOISE:
      return __ull2double_rn(splitmix64(seed + (u64)frame) >> 11) * (1.0 / 4503599627370496.0) - 1.0;
    case W_DMAX: return 1.0;
    case W_DMIN: return -1.0;
    default: return 0.0;
  }
}

// ---- ADSR over integer frame counts ----
struct EnvShape {
  i64 na, nd, nr;
  double sustain;
  double inv_na, inv_nd, inv_nr;
};

__device__ __forceinline__ double env_pre(const EnvShape& s, i64 n_on, double l_on, i64 n) {
  i64 k = n - n_on;
  if (k < s.na) {
    double t = (double)k * s.inv_na;
    return l_on + (1.0 - l_on) * (t * (2.0 - t));
  }
  i64 k2 = k - s.na;
  if (k2 < s.nd) {
    double u = 1.0 - (double)k2 * s.inv_nd;
    return s.sustain + (1.0 - s.sustain) * (u * u);
  }
  return s.sustain;
}
__device__ __forceinline__ double env_level(const EnvShape& s, i64 n_on, i64 n_off, double l_on, double l_off,
                                            i64 n) {
  if (n < n_on) return 0.0;
  if (n < n_off) return env_pre(s, n_on, l_on, n);
  i64 k = n - n_off;
  if (k < s.nr) {
    double u = 1.0 - (double)k * s.inv_nr;
    return l_off * (u * u);
  }
  return 0.0;
}

// ---- 24 dB low-pass: two transposed-DF2 sections; only b0,a1,a2 are free (b1=2*b0, b2=b0) ----
struct Lp24Ripple {
  double c0, c2, c1k, c3k;
};
struct SecCoef {
  double b0, a1, a2;
};
// k = tan(pi*fc/sr)
__device__ __forceinline__ void lp24_from_k(const Lp24Ripple& rp, double k, SecCoef& s1, SecCoef& s2) {
  double kk = k * k;
  double c1 = k * rp.c1k, c3 = k * rp.c3k;
  double a0 = 1.0 / (c1 + kk + rp.c0);
  s1.a1 = 2.0 * (rp.c0 - kk) * a0;
  s1.a2 = (c1 - kk - rp.c0) * a0;
  s1.b0 = a0 * kk;
  a0 = 1.0 / (c3 + kk + rp.c2);
  s2.a1 = 2.0 * (rp.c2 - kk) * a0;
  s2.a2 = (c3 - kk - rp.c2) * a0;
  s2.b0 = a0 * kk;
}

// ---- 2x2 affine maps (state-space recurrences) and their warp scan ----
// s' = M s + v ;  M = [m00 m01; m10 m11]
struct Affine2 {
  double m00, m01, m10, m11, v0, v1;
};
// 64-bit shuffle as two explicit 32-bit shuffles (keeps the halves in a register pair).
__device__ __forceinline__ double shfl_up_f64(double x, int delta) {
  int lo = __shfl_up_sync(0xffffffffu, __double2loint(x), delta);
  int hi = __shfl_up_sync(0xffffffffu, __double2hiint(x), delta);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_idx_f64(double x, int src) {
  int lo = __shfl_sync(0xffffffffu, __double2loint(x), src);
  int hi = __shfl_sync(0xffffffffu, __double2hiint(x), src);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ Affine2 affine_shfl_up(const Affine2& a, int delta) {
  Affine2 r;
  r.m00 = shfl_up_f64(a.m00, delta);
  r.m01 = shfl_up_f64(a.m01, delta);
  r.m10 = shfl_up_f64(a.m10, delta);
  r.m11 = shfl_up_f64(a.m11, delta);
  r.v0 = shfl_up_f64(a.v0, delta);
  r.v1 = shfl_up_f64(a.v1, delta);
  return r;
}
// a <- a after e (e = the earlier map), in place.  Ordered so that every old word of `a` is dead
// by the time its register is overwritten: one temporary per matrix row, no copies.
__device__ __forceinline__ void affine_compose_inplace(Affine2& a, const Affine2& e) {
  a.v0 = fma(a.m01, e.v1, fma(a.m00, e.v0, a.v0));
  a.v1 = fma(a.m11, e.v1, fma(a.m10, e.v0, a.v1));
  const double t0 = a.m00 * e.m01;
  a.m00 = fma(a.m01, e.m10, a.m00 * e.m00);
  a.m01 = fma(a.m01, e.m11, t0);
  const double t1 = a.m10 * e.m01;
  a.m10 = fma(a.m11, e.m10, a.m10 * e.m00);
  a.m11 = fma(a.m11, e.m11, t1);
}
// Inclusive Kogge-Stone scan over the 32 lanes (lane order = time order).
__device__ __forceinline__ Affine2 affine_warp_scan(Affine2 a, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const Affine2 prev = affine_shfl_up(a, d);
    if (lane >= d) affine_compose_inplace(a, prev);
  }
  return a;
}
// Given the inclusive scan and the state (s0,s1) at the start of the warp's span: the state at the
// start of this lane's chunk (e0,e1) and after the whole 32-lane span (end0,end1).  The state after
// lane l is incl_l applied to s; lane l's entry state is that of lane l-1 (two words to shuffle).
__device__ __forceinline__ void affine_lane_entry(const Affine2& incl, int lane, double s0, double s1, double& e0,
                                                  double& e1, double& end0, double& end1) {
  const double t0 = fma(incl.m01, s1, fma(incl.m00, s0, incl.v0));
  const double t1 = fma(incl.m11, s1, fma(incl.m10, s0, incl.v1));
  const double u0 = shfl_up_f64(t0, 1), u1 = shfl_up_f64(t1, 1);
  e0 = lane == 0 ? s0 : u0;
  e1 = lane == 0 ? s1 : u1;
  end0 = shfl_idx_f64(t0, 31);
  end1 = shfl_idx_f64(t1, 31);
}

// v <- M p + v
__device__ __forceinline__ void affine_vec_step(double& v0, double& v1, double m00, double m01, double m10,
                                                double m11, double p0, double p1) {
  v0 = fma(m01, p1, fma(m00, p0, v0));
  v1 = fma(m11, p1, fma(m10, p0, v1));
}
// The scan the voice kernels need: lane aggregates `a` (lane order = time order) and the state (s0,s1)
// at the start of the warp's span in; the state at the start of this lane's chunk (e0,e1) and after
// the whole span (end0,end1) out.  The entry state is folded into lane 0's map first, so only the
// vector part of the inclusive scan is needed at the end: the last Kogge-Stone step moves and
// combines 2 words instead of 6.
__device__ __forceinline__ void affine_scan_states(Affine2 a, int lane, double s0, double s1, double& e0, double& e1,
                                                   double& end0, double& end1) {
  if (lane == 0) affine_vec_step(a.v0, a.v1, a.m00, a.m01, a.m10, a.m11, s0, s1);
#pragma unroll
  for (int d = 1; d < 16; d <<= 1) {
    const Affine2 prev = affine_shfl_up(a, d);
    if (lane >= d) affine_compose_inplace(a, prev);
  }
  {
    const double p0 = shfl_up_f64(a.v0, 16), p1 = shfl_up_f64(a.v1, 16);
    if (lane >= 16) affine_vec_step(a.v0, a.v1, a.m00, a.m01, a.m10, a.m11, p0, p1);
  }
  const double u0 = shfl_up_f64(a.v0, 1), u1 = shfl_up_f64(a.v1, 1);
  e0 = lane == 0 ? s0 : u0;
  e1 = lane == 0 ? s1 : u1;
  end0 = shfl_idx_f64(a.v0, 31);
  end1 = shfl_idx_f64(a.v1, 31);
}

// ---- time-invariant stretches: every lane's span has the same homogeneous map M = A^T, so the map of a
// 2^k-lane span is a per-instrument constant (mp[k] = M^(2^k), row-major; mp[5] = 0) and only the
// 2-vector of zero-state end states is scanned: 4 FMAs and 2 shuffled words per step.  Lanes that a
// Kogge-Stone step leaves alone read the zero matrix instead of being predicated off, so the FMAs
// update the vector in place with no selects or moves.
__device__ __forceinline__ void lti_scan_states(double v0, double v1, const double (*mp)[4], int lane, double s0,
                                                double s1, double& e0, double& e1, double& end0, double& end1) {
  {
    const double* m = mp[lane == 0 ? 0 : 5];
    const double2 r0 = *reinterpret_cast<const double2*>(m), r1 = *reinterpret_cast<const double2*>(m + 2);
    affine_vec_step(v0, v1, r0.x, r0.y, r1.x, r1.y, s0, s1);
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int d = 1 << k;
    const double p0 = shfl_up_f64(v0, d), p1 = shfl_up_f64(v1, d);
    const double* m = mp[lane >= d ? k : 5];
    const double2 r0 = *reinterpret_cast<const double2*>(m), r1 = *reinterpret_cast<const double2*>(m + 2);
    affine_vec_step(v0, v1, r0.x, r0.y, r1.x, r1.y, p0, p1);
  }
  const double u0 = shfl_up_f64(v0, 1), u1 = shfl_up_f64(v1, 1);
  e0 = lane == 0 ? s0 : u0;
  e1 = lane == 0 ? s1 : u1;
  end0 = shfl_idx_f64(v0, 31);
  end1 = shfl_idx_f64(v1, 31);
}

// sin and cos of 2*pi*q/2^64 straight from a phase integer: the quadrant from the top bits, the rest
// as a signed angle in [-pi/4, pi/4) for two Taylor forms (|error| < 1e-16; no slow path, no table).
__device__ __forceinline__ void sincos_phase(u64 q, double* s, double* c) {
  const u64 qr = q + (1ull << 61);  // round to the nearest quarter turn
  const int quad = (int)(qr >> 62);
  const i64 r = (i64)(qr & ((1ull << 62) - 1)) - (i64)(1ull << 61);
  const double x = (double)r * (2.0 * kPi / kTwo64);
  const double z = x * x;
  double ps = 2.8114572543455206e-15, pc = 4.7794773323873853e-14;
  ps = fma(ps, z, -7.6471637318198164e-13); pc = fma(pc, z, -1.1470745597729725e-11);
  ps = fma(ps, z, 1.6059043836821613e-10);  pc = fma(pc, z, 2.0876756987868100e-09);
  ps = fma(ps, z, -2.5052108385441720e-08); pc = fma(pc, z, -2.7557319223985888e-07);
  ps = fma(ps, z, 2.7557319223985893e-06);  pc = fma(pc, z, 2.4801587301587302e-05);
  ps = fma(ps, z, -1.9841269841269841e-04); pc = fma(pc, z, -1.3888888888888889e-03);
  ps = fma(ps, z, 8.3333333333333332e-03);  pc = fma(pc, z, 4.1666666666666664e-02);
  ps = fma(ps, z, -1.6666666666666666e-01); pc = fma(pc, z, -0.5);
  ps = fma(ps, z, 1.0);                     pc = fma(pc, z, 1.0);
  ps *= x;
  const double s0 = (quad & 1) ? pc : ps;
  const double c0 = (quad & 1) ? ps : pc;
  *s = (quad & 2) ? -s0 : s0;
  *c = ((quad + 1) & 2) ? -c0 : c0;
}

// sin(2*pi*q/2^64) straight from a phase integer, |error| < 2e-16: the signed phase is folded into
// [-1/4, 1/4] turn (sin(pi - x) = sin x, exact in two's complement), converted with the 1.5*2^52 magic
// constant (integer add + DADD instead of a 64-bit I2F on the XU pipe) and fed to the odd Taylor form up
// to x^21 (truncation 1.3e-18 at pi/2).  18 instructions against ~45 of the library sinpi.
__device__ __forceinline__ double sin_phase(u64 q) {
  long long s = (long long)q;
  s = ((s ^ (s << 1)) < 0) ? (long long)(0x8000000000000000ull - (u64)s) : s;   // |a| >= 1/4  ->  sign(a)/2 - a
  const double d = __longlong_as_double(0x4338000000000000ll + (s >> 11)) - 6755399441055744.0;  // (s >> 11), exactly
  const double x = d * (2.0 * kPi / 9007199254740992.0);                         // * 2 pi / 2^53
  const double z = x * x;
  double p = -1.9572941063391263e-20;                 // -1/21!
  p = fma(p, z, 8.2206352466243295e-18);              //  1/19!
  p = fma(p, z, -2.8114572543455206e-15);             // -1/17!
  p = fma(p, z, 7.6471637318198164e-13);              //  1/15!
  p = fma(p, z, -1.6059043836821613e-10);             // -1/13!
  p = fma(p, z, 2.5052108385441720e-08);              //  1/11!
  p = fma(p, z, -2.7557319223985893e-06);             // -1/9!
  p = fma(p, z, 1.9841269841269841e-04);              //  1/7!
  p = fma(p, z, -8.3333333333333332e-03);             // -1/5!
  p = fma(p, z, 1.6666666666666666e-01);              //  1/3!
  return fma(-x * z, p, x);                           // x - x^3 (1/3! - ...)
}

// Segmented u64 sum scan (phase accumulators with resets).
struct SegSum {
  u64 sum;
  int reset;  // 1 if a reset happened inside the span; sum then counts from the last reset
};
__device__ __forceinline__ SegSum segsum_compose(const SegSum& earlier, const SegSum& later) {
  SegSum r;
  r.reset = earlier.reset | later.reset;
  r.sum = later.reset ? later.sum : earlier.sum + later.sum;
  return r;
}
__device__ __forceinline__ SegSum segsum_warp_scan(SegSum a, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    SegSum prev;
    prev.sum = __shfl_up_sync(0xffffffffu, a.sum, d);
    prev.reset = __shfl_up_sync(0xffffffffu, a.reset, d);
    if (lane >= d) a = segsum_compose(prev, a);
  }
  return a;
}

}  // namespace gbk
