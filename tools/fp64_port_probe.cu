// fp64_port_probe.cu — how the FP64 pipe and the issue port of an sm_100 sub-partition share cycles.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_port_probe tools/fp64_port_probe.cu && /tmp/fp64_port_probe
//
// Each warp runs a loop whose body is NF independent DFMAs interleaved with NI independent integer adds (volatile
// asm, so the SASS keeps the written order; 8 accumulators of each kind, i.e. a dependent distance of 8).  With W
// warps per sub-partition (one CTA of 4 W warps per SM) the probe prints the cycles one sub-partition spends per
// loop body, next to two models:
//   overlap   max(2 NF, NF + NI)   a DFMA holds the FP64 pipe for 2 cycles but the issue port for 1
//   blocking  2 NF + NI            a DFMA also holds the issue port for its second cycle
// welsh_rest_kernel issues 25.25 FP64 + 36.5 other instructions per voice-sample (profiles/r2_rest_kernel_facts.json);
// which model holds decides whether its floor is 62 or 87 cycles.
#include <cstdio>
#include <cuda_runtime.h>

template <int NF, int NI>
__global__ void probe(double* out, long long* cycles, int iters) {
  double f[8];
  unsigned a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[i] = 1.0 + threadIdx.x * 1e-9 + i; a[i] = threadIdx.x + i; }
  const double m = 0.999999, c = 1e-7;
  const unsigned k = blockIdx.x | 1;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it += 2) {
    // two bodies per trip: integer adds in the first, xors in the second (ptxas would fold two adds of one
    // accumulator into a single three-input IADD3)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      if constexpr (NF == 0) {
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          if (half == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[j & 7]) : "r"(k));
          else asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[j & 7]) : "r"(k));
        }
      } else {
#pragma unroll
        for (int j = 0; j < NF; ++j) {
          asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[j & 7]) : "d"(m), "d"(c));
#pragma unroll
          for (int q = j * NI / NF; q < (j + 1) * NI / NF; ++q) {  // the integer ops spread evenly between the DFMAs
            if (((q >> 3) + half) % 2 == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[q & 7]) : "r"(k));
            else asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[q & 7]) : "r"(k));
          }
        }
      }
    }
  }
  const long long t1 = clock64();
  double s = 0.0;
  unsigned u = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += f[i]; u += a[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + u;
  if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

template <int NF, int NI>
void run(int warps_per_smsp, double* out, long long* cyc) {
  const int iters = 20000, threads = 128 * warps_per_smsp, blocks = 148;
  probe<NF, NI><<<blocks, threads>>>(out, cyc, 100);
  probe<NF, NI><<<blocks, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  const int nw = blocks * threads / 32;
  long long* h = new long long[nw];
  cudaMemcpy(h, cyc, nw * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0.0;
  for (int i = 0; i < nw; ++i) mean += (double)h[i];
  mean /= nw;
  delete[] h;
  // the W warps of a sub-partition run concurrently: it completes W loop bodies in mean / iters cycles
  const double per_body_smsp = mean / iters / warps_per_smsp;
  const double overlap = NF * 2 > NF + NI ? NF * 2 : NF + NI, blocking = 2.0 * NF + NI;
  printf("NF %2d NI %2d warps/smsp %d: %.2f cycles per body per sub-partition (overlap model %.0f, blocking model %.0f)\n", NF, NI,
         warps_per_smsp, per_body_smsp, overlap, blocking);
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(double));
  cudaMalloc(&cyc, 148 * 32 * sizeof(long long));
  for (int w : {1, 2, 4, 8}) {
    run<8, 0>(w, out, cyc);
    run<0, 8>(w, out, cyc);
    run<8, 4>(w, out, cyc);
    run<8, 8>(w, out, cyc);
    run<8, 12>(w, out, cyc);
    run<8, 16>(w, out, cyc);
    run<8, 24>(w, out, cyc);
  }
  cudaError_t rc = cudaDeviceSynchronize();
  if (rc != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(rc)); return 1; }
  return 0;
}
