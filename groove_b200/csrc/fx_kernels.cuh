// fx_kernels.cuh — stream effects over stereo f64 chunks (double2 = one StereoSample).
//
// Each kernel replaces `transform_audio(StereoSample)` called once per frame
// (orchestration/src/orchestrator.rs:446-456) by one pass over the whole chunk:
//   * memoryless effects (mixer/gain/limiter/bitcrusher/compressor) are fused with the summation
//     of their inputs: 16 B read per source + 16 B write per frame;
//   * IIR filters are 2x2 affine recurrences scanned across a CTA (lane chunk -> warp shuffle
//     scan -> cross-warp combine through shared memory);
//   * delay-line effects are evaluated polyphase: a recirculating line of D frames is D
//     independent first-order recurrences, one thread each, with no scan at all.
#pragma once

#include "dsp.cuh"

namespace gbk {

constexpr int kMaxSources = 8;
struct SourceList {
  const double2* p[kMaxSources];
  int n;
};

__device__ __forceinline__ double2 sum_sources(const SourceList& s, int t) {
  double l = 0.0, r = 0.0;
#pragma unroll
  for (int k = 0; k < kMaxSources; ++k) {
    if (k < s.n) {
      double2 v = s.p[k][t];
      l += v.x; r += v.y;
    }
  }
  return make_double2(l, r);
}

enum { OP_SUM = 0, OP_GAIN, OP_LIMITER, OP_BITCRUSHER, OP_COMPRESSOR, OP_DCA };

__device__ __forceinline__ double op_apply(int op, double x, double a, double b) {
  switch (op) {
    case OP_GAIN: return x * a;
    case OP_LIMITER: {
      double m = fabs(x);
      m = m < a ? a : (m > b ? b : m);
      return copysign(m, x);
    }
    case OP_BITCRUSHER: {
      // a = 2^bits
      double m = __ddiv_rn(__dmul_rn(floor(__ddiv_rn(__dmul_rn(fabs(x), 32767.0), a)), a), 32767.0);
      return copysign(m, x);
    }
    case OP_COMPRESSOR: {
      double m = fabs(x);
      if (m > a) m = a + (m - a) * b;
      return copysign(m, x);
    }
    default: return x;
  }
}

// Parameter automation is piecewise constant in time.  A chunk's parameter history for one effect is
// a segment table in device memory: segment k holds the kernel-ready values that apply from chunk
// frame t0[k] (t0[0] == 0) until the next segment.  One launch covers the whole chunk.
struct SegParam {
  int t0;
  int pad;
  double v[6];
};
__device__ __forceinline__ int seg_find(const SegParam* __restrict__ segs, int n, int t) {
  int lo = 0, hi = n - 1;  // last k with segs[k].t0 <= t
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (segs[mid].t0 <= t) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

// out[t] = op(sum of sources[t]) with the parameters of the segment covering t.
// OP_DCA multiplies the two channels by v[0], v[1] (instrument gain/pan automation).
__global__ void __launch_bounds__(256) pointwise_kernel(SourceList src, double2* __restrict__ out, int frames,
                                                         int op, const SegParam* __restrict__ segs, int nseg) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= frames) return;
  double2 v = sum_sources(src, t);
  if (op == OP_SUM) {
    out[t] = v;
    return;
  }
  const SegParam& sp = segs[nseg > 1 ? seg_find(segs, nseg, t) : 0];
  if (op == OP_DCA) out[t] = make_double2(v.x * sp.v[0], v.y * sp.v[1]);
  else out[t] = make_double2(op_apply(op, v.x, sp.v[0], sp.v[1]), op_apply(op, v.y, sp.v[0], sp.v[1]));
}

// Sidechain (signal-passthrough source -> effect parameter, gb_link_control): the target's parameter table
// of this chunk is built on the device from the source's output.  Control boundaries are the absolute
// frames that are multiples of `period`; the value taken at boundary t is min(1, |mono|) of the source at
// t - 1 (for t = 0: the last frame of the previous chunk, carried in `state_in`).  state = {current
// control value, last l, last r, have-last flag}; `state_out` receives the values for the next chunk.
// first = chunk-relative frame of the first boundary (0 .. period-1); nseg = segments to write.
__global__ void __launch_bounds__(128) sidechain_table_kernel(const double2* __restrict__ src, int frames, int first,
                                                               int period, int nseg, SegParam base, int slot,
                                                               SegParam* __restrict__ table,
                                                               const double* __restrict__ state_in,
                                                               double* __restrict__ state_out, long long f0) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nseg) return;
  SegParam sp = base;
  double value;
  if (k == 0) {
    sp.t0 = 0;
    value = state_in[0];
    if (first == 0 && f0 > 0 && state_in[3] != 0.0) {
      const double m = fabs(0.5 * (state_in[1] + state_in[2]));
      value = m > 1.0 ? 1.0 : m;
    }
  } else {
    const int t = first + (k - (first == 0 ? 0 : 1)) * period;
    sp.t0 = t;
    const double2 v = src[t - 1];
    const double m = fabs(0.5 * (v.x + v.y));
    value = m > 1.0 ? 1.0 : m;
  }
  sp.v[slot] = value;
  table[k] = sp;
  if (k == nseg - 1) {
    const double2 last = src[frames - 1];
    state_out[0] = value;
    state_out[1] = last.x;
    state_out[2] = last.y;
    state_out[3] = 1.0;
  }
}

// out[t] = sum over a device-resident pointer table (any number of sources).  HBM-bound: 16 bytes per
// source per frame.  A block covers 64 frames with 4 thread rows; row r sums sources r, r+4, r+8, ...
// (4 loads in flight each, so 16 per frame) and row 0 adds the four row sums in row order — a fixed
// summation order, independent of the launch.
// The multi-GPU bus mixdown (gb_bus_exchange_reduce): the root GPU sums the ranks' stereo buses straight out of
// their HBM — ptrs[r] is rank r's exchange buffer, mapped into this process over CUDA IPC, so the loads of the
// remote rows travel over NVLink (P2P) while the kernel adds; n <= kMaxPeers loads in flight per frame, summed in
// rank order (a fixed order: the mix does not depend on timing).  16 B per frame per rank in, 16 B out.
constexpr int kMaxPeers = 16;
struct PeerTable { const double2* p[kMaxPeers]; };
__global__ void __launch_bounds__(256) peer_sum_kernel(PeerTable tab, int n, double2* __restrict__ out, size_t frames) {
  for (size_t t = (size_t)blockIdx.x * 256 + threadIdx.x; t < frames; t += (size_t)gridDim.x * 256) {
    double2 v[kMaxPeers];
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < n) v[r] = __ldcg(tab.p[r] + t);   // L2 only: peer data is read once
    double l = 0.0, rr = 0.0;
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < n) { l += v[r].x; rr += v[r].y; }
    out[t] = make_double2(l, rr);
  }
}

constexpr int kSumRows = 4, kSumFrames = 64;
__global__ void __launch_bounds__(kSumRows * kSumFrames) sum_table_kernel(const double2* const* __restrict__ ptrs, int n,
                                                                          double2* __restrict__ out, int frames) {
  __shared__ double2 part[kSumRows][kSumFrames];
  const int tx = threadIdx.x % kSumFrames, row = threadIdx.x / kSumFrames;
  const int t = blockIdx.x * kSumFrames + tx;
  double l = 0.0, r = 0.0;
  if (t < frames) {
    int k = row;
    for (; k + 3 * kSumRows < n; k += 4 * kSumRows) {
      const double2 a = ptrs[k][t], b = ptrs[k + kSumRows][t], c = ptrs[k + 2 * kSumRows][t], d = ptrs[k + 3 * kSumRows][t];
      l += a.x; r += a.y;
      l += b.x; r += b.y;
      l += c.x; r += c.y;
      l += d.x; r += d.y;
    }
    for (; k < n; k += kSumRows) {
      const double2 a = ptrs[k][t];
      l += a.x; r += a.y;
    }
  }
  part[row][tx] = make_double2(l, r);
  __syncthreads();
  if (row == 0 && t < frames) {
#pragma unroll
    for (int q = 1; q < kSumRows; ++q) {
      l += part[q][tx].x;
      r += part[q][tx].y;
    }
    out[t] = make_double2(l, r);
  }
}
// One launch for every instrument whose voices are split over several CTAs: blockIdx.y picks the
// instrument; its partial buffers are `count` planes `stride` frames apart.
struct PartialDesc {
  const double2* base;
  double2* out;
  size_t stride;
  int count;
  int pad;
};
__global__ void __launch_bounds__(256) reduce_partials_kernel(const PartialDesc* __restrict__ descs, int frames) {
  const PartialDesc d = descs[blockIdx.y];
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= frames) return;
  double l = 0.0, r = 0.0;
  for (int k = 0; k < d.count; ++k) {
    double2 v = d.base[(size_t)k * d.stride + t];
    l += v.x; r += v.y;
  }
  d.out[t] = make_double2(l, r);
}

// Zero a list of frame ranges (idle stretches of solo voices' output buffers): one CTA per range.
struct ZeroRange {
  double2* p;
  int n;
  int pad;
};
__global__ void __launch_bounds__(256) zero_ranges_kernel(const ZeroRange* __restrict__ ranges) {
  const ZeroRange r = ranges[blockIdx.x];
  for (int t = threadIdx.x; t < r.n; t += blockDim.x) r.p[t] = make_double2(0.0, 0.0);
}
__global__ void __launch_bounds__(256) fill_kernel(double2* __restrict__ out, int n, double l, double r) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = make_double2(l, r);
}

// ---- bare oscillator / envelope devices (GB_INST_OSCILLATOR, GB_INST_ENVELOPE) ----------------------
// Both are closed forms of the absolute frame, so they need no device state: the oscillator runs free
// from frame 0 (phase(n) = n * dq mod 2^64), the envelope's note state per segment of the chunk comes
// from the host (note frames are integers; v = {l_on, l_off, n_on, n_off}, frames exact in a double).
struct OscSourceDesc {
  int waveform, pad;
  u64 dq, duty_q, seed;
};
__global__ void __launch_bounds__(256) oscillator_source_kernel(OscSourceDesc d, double2* __restrict__ out, i64 f0,
                                                                 int frames) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= frames) return;
  const i64 n = f0 + t;
  const double v = wave_value(d.waveform, (u64)n * d.dq, d.duty_q, d.seed, n);
  out[t] = make_double2(v, v);
}
__global__ void __launch_bounds__(256) envelope_source_kernel(EnvShape sh, const SegParam* __restrict__ segs, int nseg,
                                                               double2* __restrict__ out, i64 f0, int frames) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= frames) return;
  const SegParam& sp = segs[nseg > 1 ? seg_find(segs, nseg, t) : 0];
  const double v = env_level(sh, (i64)sp.v[2], (i64)sp.v[3], sp.v[0], sp.v[1], f0 + t);
  out[t] = make_double2(v, v);
}

// f64 stereo -> 16-bit PCM: (x * 32767.0) as i16 — truncate toward zero, saturate, NaN -> 0
// (orchestration/src/helpers.rs:78,90-91).
__device__ __forceinline__ short pcm16_of(double x) {
  double v = x * 32767.0;
  if (v != v) return 0;
  if (v >= 32767.0) return 32767;
  if (v <= -32768.0) return -32768;
  return (short)__double2int_rz(v);
}
__global__ void __launch_bounds__(256) pcm16_kernel(const double2* __restrict__ in, short2* __restrict__ out, int n) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    double2 v = in[t];
    out[t] = make_short2(pcm16_of(v.x), pcm16_of(v.y));
  }
}

// ---------------------------------------------------------------- IIR sections ---
constexpr int kFxT = 8;
constexpr int kFxWarps = 8;
constexpr int kFxRound = kFxWarps * 32 * kFxT;  // frames per CTA round

struct BiquadCoefs {  // a0-normalised RBJ coefficients, Direct Form 1
  double b0, b1, b2, a1, a2;
};
struct BiquadState {  // per channel: x[n-1], x[n-2], y[n-1], y[n-2]
  double x1[2], x2[2], y1[2], y2[2];
};

// Cross-warp exclusive combine: every warp publishes its aggregate; warp w composes 0..w-1.
__device__ __forceinline__ void cta_entry_state(const Affine2& warp_total, Affine2* sh, int warp, int lane, double s0,
                                                double s1, double& w0, double& w1, double& end0, double& end1) {
  if (lane == 31) sh[warp] = warp_total;
  __syncthreads();
  double a = s0, b = s1;
  for (int w = 0; w < kFxWarps; ++w) {
    if (w == warp) { w0 = a; w1 = b; }
    Affine2 m = sh[w];
    double na = m.m00 * a + m.m01 * b + m.v0;
    double nb = m.m10 * a + m.m11 * b + m.v1;
    a = na; b = nb;
  }
  end0 = a; end1 = b;
  __syncthreads();
}

// A memoryless effect (gain, limiter, bitcrusher, compressor) fused behind an IIR stage: applied to the
// stage's output on its way to memory, with its own parameter segment table.  op < 0: none.
struct PostOp {
  const SegParam* segs;
  int nseg;
  int op;
};
__device__ __forceinline__ double2 post_apply(const PostOp& po, double2 v, int t) {
  if (po.op < 0) return v;
  const SegParam& sp = po.segs[po.nseg > 1 ? seg_find(po.segs, po.nseg, t) : 0];
  return make_double2(op_apply(po.op, v.x, sp.v[0], sp.v[1]), op_apply(po.op, v.y, sp.v[0], sp.v[1]));
}

// One CTA = one biquad effect over the whole chunk; coefficients (v[0..4] = b0,b1,b2,a1,a2) follow the
// segment table frame by frame.  kFxWarps warps; both channels per thread.  Input and output go through
// shared memory, so global loads and stores are coalesced 16-byte accesses.
__device__ __forceinline__ void biquad_df1_body(const SourceList& src, double2* __restrict__ out, int frames,
                                                const SegParam* __restrict__ segs, int nseg,
                                                BiquadState* __restrict__ state, const PostOp& po) {
  __shared__ Affine2 sh[2][kFxWarps];
  __shared__ double2 sx[kFxRound + 2 + (kFxRound + 2) / kFxT + 1];  // slot P(k) = k + k / kFxT: conflict-free per-lane runs
  auto P = [](int k) { return k + k / kFxT; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  BiquadState st = *state;
  for (int r0 = 0; r0 < frames; r0 += kFxRound) {
    // stage the round's input (coalesced), with the two history frames in front
    for (int i = threadIdx.x; i < kFxRound; i += blockDim.x) {
      int t = r0 + i;
      sx[P(i + 2)] = t < frames ? sum_sources(src, t) : make_double2(0.0, 0.0);
    }
    if (threadIdx.x == 0) {
      sx[P(0)] = make_double2(st.x2[0], st.x2[1]);
      sx[P(1)] = make_double2(st.x1[0], st.x1[1]);
    }
    __syncthreads();
    const int base = (warp * 32 + lane) * kFxT;  // round-relative first frame of this lane
    double yp[2][kFxT], g0[kFxT], g1[kFxT];
    double p0[2] = {0.0, 0.0}, p1[2] = {0.0, 0.0};
    double h00 = 1.0, h01 = 0.0, h10 = 0.0, h11 = 1.0;
    const int nvalid = frames - (r0 + base);
    int k = nvalid > 0 ? seg_find(segs, nseg, r0 + base) : 0;
    int next_t0 = k + 1 < nseg ? segs[k + 1].t0 : 0x7fffffff;
    double b0 = segs[k].v[0], b1 = segs[k].v[1], b2 = segs[k].v[2], a1 = segs[k].v[3], a2 = segs[k].v[4];
#pragma unroll
    for (int j = 0; j < kFxT; ++j) {
      if (j < nvalid) {
        while (r0 + base + j >= next_t0) {
          ++k;
          next_t0 = k + 1 < nseg ? segs[k + 1].t0 : 0x7fffffff;
          b0 = segs[k].v[0]; b1 = segs[k].v[1]; b2 = segs[k].v[2]; a1 = segs[k].v[3]; a2 = segs[k].v[4];
        }
        double2 x0 = sx[P(base + j + 2)], xm1 = sx[P(base + j + 1)], xm2 = sx[P(base + j)];
        double v[2] = {b0 * x0.x + b1 * xm1.x + b2 * xm2.x, b0 * x0.y + b1 * xm1.y + b2 * xm2.y};
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          double y = v[ch] - a1 * p0[ch] - a2 * p1[ch];
          yp[ch][j] = y;
          p1[ch] = p0[ch];
          p0[ch] = y;
        }
        // homogeneous: contribution of the entry state (y1,y2) to y[n] is row 0 of the running product
        double t00 = -a1 * h00 - a2 * h10, t01 = -a1 * h01 - a2 * h11;
        h10 = h00; h11 = h01;
        h00 = t00; h01 = t01;
        g0[j] = h00; g1[j] = h01;
      } else {
        yp[0][j] = 0.0; yp[1][j] = 0.0; g0[j] = 0.0; g1[j] = 0.0;
      }
    }
    double e0[2], e1[2], endv0[2], endv1[2];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      Affine2 a;
      a.m00 = h00; a.m01 = h01; a.m10 = h10; a.m11 = h11; a.v0 = p0[ch]; a.v1 = p1[ch];
      Affine2 inc = affine_warp_scan(a, lane);
      double w0, w1, ce0, ce1;
      cta_entry_state(inc, sh[ch], warp, lane, st.y1[ch], st.y2[ch], w0, w1, ce0, ce1);
      double d0, d1;
      affine_lane_entry(inc, lane, w0, w1, e0[ch], e1[ch], d0, d1);
      endv0[ch] = ce0; endv1[ch] = ce1;
    }
    // carry state to the next round
    int last = min(kFxRound, frames - r0);  // frames in this round
    double2 xl1 = sx[P(last + 1)], xl2 = sx[P(last)];
    __syncthreads();  // every thread has read its inputs: the staging buffer now takes the round's output
#pragma unroll
    for (int j = 0; j < kFxT; ++j) {
      double l = yp[0][j] + g0[j] * e0[0] + g1[j] * e1[0];
      double r = yp[1][j] + g0[j] * e0[1] + g1[j] * e1[1];
      sx[P(base + j + 2)] = make_double2(l, r);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < last; i += blockDim.x) out[r0 + i] = post_apply(po, sx[P(i + 2)], r0 + i);
    __syncthreads();
    st.x1[0] = xl1.x; st.x1[1] = xl1.y;
    st.x2[0] = xl2.x; st.x2[1] = xl2.y;
    st.y1[0] = endv0[0]; st.y1[1] = endv0[1];
    st.y2[0] = endv1[0]; st.y2[1] = endv1[1];
  }
  if (threadIdx.x == 0) *state = st;
}
__global__ void __launch_bounds__(32 * kFxWarps) biquad_df1_kernel(SourceList src, double2* __restrict__ out,
                                                                    int frames, const SegParam* __restrict__ segs,
                                                                    int nseg, BiquadState* __restrict__ state) {
  PostOp none;
  none.segs = nullptr; none.nseg = 0; none.op = -1;
  biquad_df1_body(src, out, frames, segs, nseg, state, none);
}

// 24 dB low-pass effect: two transposed-DF2 sections; coefficients (v[0..5] = b0,a1,a2 of section 1
// then of section 2) follow the segment table frame by frame.
struct Lp24State {
  double s[2][4];
};
__device__ __forceinline__ void lp24_body(const SourceList& src, double2* __restrict__ out, int frames,
                                          const SegParam* __restrict__ segs, int nseg, Lp24State* __restrict__ state,
                                          const PostOp& po) {
  __shared__ Affine2 sh[2][kFxWarps];
  __shared__ double2 sx[kFxRound + kFxRound / kFxT];  // a lane's kFxT frames sit kFxT + 1 slots from its neighbour's
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Lp24State st = *state;
  for (int r0 = 0; r0 < frames; r0 += kFxRound) {
    // stage the round's input: coalesced global loads, then each lane takes its kFxT consecutive frames
    for (int i = threadIdx.x; i < kFxRound; i += blockDim.x) {
      const int t = r0 + i;
      sx[i + i / kFxT] = t < frames ? sum_sources(src, t) : make_double2(0.0, 0.0);
    }
    __syncthreads();
    const int lbase = (warp * 32 + lane) * kFxT;
    const int base = r0 + lbase;
    const int nvalid = frames - base;
    double sig[2][kFxT];
#pragma unroll
    for (int j = 0; j < kFxT; ++j) {
      const double2 v = sx[lbase + lbase / kFxT + j];
      sig[0][j] = v.x; sig[1][j] = v.y;
    }
    const int k0 = nvalid > 0 ? seg_find(segs, nseg, base) : 0;
#pragma unroll
    for (int sec = 0; sec < 2; ++sec) {
      int k = k0;
      int next_t0 = k + 1 < nseg ? segs[k + 1].t0 : 0x7fffffff;
      double cb0 = segs[k].v[3 * sec], ca1 = segs[k].v[3 * sec + 1], ca2 = segs[k].v[3 * sec + 2];
      double g0[kFxT], g1[kFxT];
      double p0[2] = {0.0, 0.0}, p1[2] = {0.0, 0.0};
      double h00 = 1.0, h01 = 0.0, h10 = 0.0, h11 = 1.0;
#pragma unroll
      for (int j = 0; j < kFxT; ++j) {
        g0[j] = 0.0; g1[j] = 0.0;
        if (j < nvalid) {
          while (base + j >= next_t0) {
            ++k;
            next_t0 = k + 1 < nseg ? segs[k + 1].t0 : 0x7fffffff;
            cb0 = segs[k].v[3 * sec]; ca1 = segs[k].v[3 * sec + 1]; ca2 = segs[k].v[3 * sec + 2];
          }
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            double bx = cb0 * sig[ch][j];
            double y = bx + p0[ch];
            sig[ch][j] = y;
            double n0 = 2.0 * bx + ca1 * y + p1[ch];
            p1[ch] = bx + ca2 * y;
            p0[ch] = n0;
          }
          g0[j] = h00; g1[j] = h01;
          double t00 = ca1 * h00 + h10, t01 = ca1 * h01 + h11;
          h10 = ca2 * h00; h11 = ca2 * h01;
          h00 = t00; h01 = t01;
        }
      }
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        Affine2 a;
        a.m00 = h00; a.m01 = h01; a.m10 = h10; a.m11 = h11; a.v0 = p0[ch]; a.v1 = p1[ch];
        Affine2 inc = affine_warp_scan(a, lane);
        double w0, w1, ce0, ce1, e0, e1, d0, d1;
        cta_entry_state(inc, sh[ch], warp, lane, st.s[ch][2 * sec], st.s[ch][2 * sec + 1], w0, w1, ce0, ce1);
        affine_lane_entry(inc, lane, w0, w1, e0, e1, d0, d1);
        st.s[ch][2 * sec] = ce0; st.s[ch][2 * sec + 1] = ce1;
#pragma unroll
        for (int j = 0; j < kFxT; ++j) sig[ch][j] += g0[j] * e0 + g1[j] * e1;
      }
    }
    // (the barriers of cta_entry_state lie between the last input read and this write)
#pragma unroll
    for (int j = 0; j < kFxT; ++j) sx[lbase + lbase / kFxT + j] = make_double2(sig[0][j], sig[1][j]);
    __syncthreads();
    const int last = min(kFxRound, frames - r0);
    for (int i = threadIdx.x; i < last; i += blockDim.x) out[r0 + i] = post_apply(po, sx[i + i / kFxT], r0 + i);
    __syncthreads();
  }
  if (threadIdx.x == 0) *state = st;
}
__global__ void __launch_bounds__(32 * kFxWarps) lp24_kernel(SourceList src, double2* __restrict__ out, int frames,
                                                              const SegParam* __restrict__ segs, int nseg,
                                                              Lp24State* __restrict__ state) {
  PostOp none;
  none.segs = nullptr; none.nseg = 0; none.op = -1;
  lp24_body(src, out, frames, segs, nseg, state, none);
}

// ---- batched launches: independent effect nodes of one kind and one graph level in ONE launch ----------
// (a batch of songs, or the parallel chains of one song: config 1's chain copied 1024 times is 1024 CTAs
// here instead of 1024 one-CTA launches).  blockIdx.x (IIR stages) / blockIdx.y (pointwise) picks the node.
struct FxDesc {
  SourceList src;
  double2* out;
  const SegParam* segs;
  void* state;     // Lp24State / BiquadState of an IIR stage
  PostOp post;     // IIR stages: a memoryless effect fused behind the stage
  int nseg;
  int op;          // pointwise nodes: OP_*
};
template <int MINB>
__global__ void __launch_bounds__(32 * kFxWarps, MINB) lp24_batch_kernel(const FxDesc* __restrict__ descs, int frames) {
  const FxDesc& d = descs[blockIdx.x];
  lp24_body(d.src, d.out, frames, d.segs, d.nseg, static_cast<Lp24State*>(d.state), d.post);
}
template <int MINB>
__global__ void __launch_bounds__(32 * kFxWarps, MINB) biquad_batch_kernel(const FxDesc* __restrict__ descs, int frames) {
  const FxDesc& d = descs[blockIdx.x];
  biquad_df1_body(d.src, d.out, frames, d.segs, d.nseg, static_cast<BiquadState*>(d.state), d.post);
}
// a single IIR stage with a fused post-op (descriptor by value)
__global__ void __launch_bounds__(32 * kFxWarps) lp24_single_kernel(FxDesc d, int frames) {
  lp24_body(d.src, d.out, frames, d.segs, d.nseg, static_cast<Lp24State*>(d.state), d.post);
}
__global__ void __launch_bounds__(32 * kFxWarps) biquad_single_kernel(FxDesc d, int frames) {
  biquad_df1_body(d.src, d.out, frames, d.segs, d.nseg, static_cast<BiquadState*>(d.state), d.post);
}
__global__ void __launch_bounds__(256) pointwise_batch_kernel(const FxDesc* __restrict__ descs, int frames) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= frames) return;
  const FxDesc& d = descs[blockIdx.y];
  const double2 v = sum_sources(d.src, t);
  if (d.op == OP_SUM) {
    d.out[t] = v;
    return;
  }
  const SegParam& sp = d.segs[d.nseg > 1 ? seg_find(d.segs, d.nseg, t) : 0];
  d.out[t] = make_double2(op_apply(d.op, v.x, sp.v[0], sp.v[1]), op_apply(d.op, v.y, sp.v[0], sp.v[1]));
}



// ------------------------------------------------------------- delay-line effects ---
// `hist` holds the last `len` input frames before the chunk: hist[len-1] = x[chunk_start-1].
__device__ __forceinline__ double2 delayed_input(const SourceList& src, const double2* hist, int len, int t, int d) {
  int s = t - d;
  if (s >= 0) return sum_sources(src, s);
  return hist[len + s];
}
// Delay: y[n] = x[n-D]
__global__ void __launch_bounds__(256) delay_kernel(SourceList src, const double2* __restrict__ hist, int len, int d,
                                                     double2* __restrict__ out, int n) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = d == 0 ? sum_sources(src, t) : delayed_input(src, hist, len, t, d);
}
// Chorus: y[n] = (1-w) x[n] + w * mean_i x[n - taps[i]]
struct ChorusTaps {
  int tap[64];
  int nv;
};
__global__ void __launch_bounds__(256) chorus_kernel(SourceList src, const double2* __restrict__ hist, int len,
                                                      ChorusTaps taps, const SegParam* __restrict__ segs, int nseg,
                                                      double2* __restrict__ out, int frames) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= frames) return;
  const double wet = segs[nseg > 1 ? seg_find(segs, nseg, t) : 0].v[0];
  double2 x = sum_sources(src, t);
  double al = 0.0, ar = 0.0;
  for (int i = 0; i < taps.nv; ++i) {
    double2 v = taps.tap[i] == 0 ? x : delayed_input(src, hist, len, t, taps.tap[i]);
    al += v.x; ar += v.y;
  }
  double inv = (double)taps.nv;
  out[t] = make_double2(__dadd_rn(__dmul_rn(1.0 - wet, x.x), __dmul_rn(wet, __ddiv_rn(al, inv))),
                        __dadd_rn(__dmul_rn(1.0 - wet, x.y), __dmul_rn(wet, __ddiv_rn(ar, inv))));
}
// new_hist = last `len` frames of (hist ++ chunk input); written to a second buffer (ping-pong).
__global__ void __launch_bounds__(256) history_update_kernel(SourceList src, const double2* __restrict__ hist,
                                                              double2* __restrict__ new_hist, int len, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  int s = n - len + i;  // chunk-relative frame that lands in slot i
  new_hist[i] = s >= 0 ? sum_sources(src, s) : hist[i + n];
}

// Reverb.  Recirculating comb:  y[n] = w[n-D],  w[n] = a*x[n] + g*w[n-D]   (ring[n mod D] holds w)
// One thread per (channel, comb, residue); writes its comb's output into its own plane.
struct ReverbDesc {
  int comb_d[4];
  double comb_g[4];
  int ap_d[2];
  double ap_g[2];
  double* comb_ring[2][4];  // per channel, per comb: D doubles
  double* ap_ring[2][2];
};
__global__ void __launch_bounds__(256) reverb_comb_kernel(SourceList src, ReverbDesc d,
                                                           const SegParam* __restrict__ segs, int nseg,
                                                           double* __restrict__ comb_out /* [2][4][n] */, int n,
                                                           long long pos0) {
  const int t0 = 0, t1 = n;
  int comb = blockIdx.y & 3, ch = blockIdx.y >> 2;
  int D = d.comb_d[comb];
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= D) return;
  double g = d.comb_g[comb];
  double* ring = d.comb_ring[ch][comb];
  double* o = comb_out + ((size_t)(ch * 4 + comb)) * (size_t)n;
  // first frame >= t0 whose ring slot ((pos0 + t) mod D) is r
  int first = (int)((r - ((pos0 + t0) % D) + D) % D) + t0;
  double w = ring[r];
  for (int t = first; t < t1; t += D) {
    double2 x = sum_sources(src, t);
    const double attenuation = segs[nseg > 1 ? seg_find(segs, nseg, t) : 0].v[0];
    double xa = (ch == 0 ? x.x : x.y) * attenuation;
    o[t] = w;
    w = __dadd_rn(xa, __dmul_rn(g, w));
  }
  ring[r] = w;
}
// All-pass:  y[n] = -g*x[n] + w[n-D],  w[n] = x[n] + g*y[n].  Stage 0 reads the 4 comb planes.
__global__ void __launch_bounds__(256) reverb_allpass_kernel(const double* __restrict__ in /* planes */, int nplanes,
                                                              size_t plane_stride, ReverbDesc d, int stage,
                                                              double* __restrict__ out /* [2][n] */, int n,
                                                              long long pos0) {
  const int t0 = 0, t1 = n;
  int ch = blockIdx.y;
  int D = d.ap_d[stage];
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= D) return;
  double g = d.ap_g[stage];
  double* ring = d.ap_ring[ch][stage];
  const double* ip = in + (size_t)ch * (size_t)nplanes * plane_stride;
  double* o = out + (size_t)ch * (size_t)n;
  int first = (int)((r - ((pos0 + t0) % D) + D) % D) + t0;
  double w = ring[r];
  for (int t = first; t < t1; t += D) {
    double x = 0.0;
    for (int p = 0; p < nplanes; ++p) x += ip[(size_t)p * plane_stride + t];
    double y = __dadd_rn(__dmul_rn(-g, x), w);
    o[t] = y;
    w = __dadd_rn(x, __dmul_rn(g, y));
  }
  ring[r] = w;
}
__global__ void __launch_bounds__(256) interleave_kernel(const double* __restrict__ planes /* [2][n] */,
                                                          double2* __restrict__ out, int n) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = make_double2(planes[t], planes[(size_t)n + t]);
}

// ---------------------------------------------------------------- microbenchmark ---
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters) {
  T a0 = (T)threadIdx.x * (T)1e-3, a1 = a0 + (T)1, a2 = a0 + (T)2, a3 = a0 + (T)3;
  T a4 = a0 + (T)4, a5 = a0 + (T)5, a6 = a0 + (T)6, a7 = a0 + (T)7;
  const T m = (T)0.999999, c = (T)1e-6;
  for (int i = 0; i < iters; ++i) {
    a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
    a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace gbk
