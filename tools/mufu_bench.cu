// Throughput probe for the softmax inner loop on sm_100a: how many 2^x per clock per SM can the MUFU pipe, the FMA-pipe
// polynomial (ex2_poly of common.cuh, scalar and packed f32x2) and mixes of both sustain?  Answers whether the
// attention kernel (55 % "XU busy" in ncu, whatever the structure) is at the hardware's exp2 rate or below it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mufu_bench tools/mufu_bench.cu && tools/bin/mufu_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "../versatile-diffusion_b200/csrc/common.cuh"
using namespace vdb;

// packed-pair polynomial 2^x: the Horner steps and the magic-number split run as f32x2 operations
__device__ __forceinline__ void ex2_poly2(float xa, float xb, float& ya, float& yb) {
  const unsigned long long magic = pack_f2(12582912.0f, 12582912.0f), nmagic = pack_f2(-12582912.0f, -12582912.0f);
  const unsigned long long one = pack_f2(1.f, 1.f), mone = pack_f2(-1.f, -1.f);
  const unsigned long long x = pack_f2(fmaxf(xa, -126.f), fmaxf(xb, -126.f));
  const unsigned long long r = add_f2(x, magic);
  const unsigned long long f = fma_f2(add_f2(r, nmagic), mone, x);   // x - (r - magic)
  unsigned long long p = fma_f2(pack_f2(0.0550886838f, 0.0550886838f), f, pack_f2(0.242604051f, 0.242604051f));
  p = fma_f2(p, f, pack_f2(0.693276242f, 0.693276242f));
  p = fma_f2(p, f, pack_f2(0.99992894f, 0.99992894f));
  (void)one;
  float pa, pb, ra, rb;
  unpack_f2(p, pa, pb); unpack_f2(r, ra, rb);
  ya = __int_as_float(__float_as_int(pa) + (__float_as_int(ra) << 23));
  yb = __int_as_float(__float_as_int(pb) + (__float_as_int(rb) << 23));
}

// MODE 0: MUFU only, 1: scalar poly only, 2: packed poly only, 3+k: of every 8 values, k (1..7) go through the packed poly
template <int MODE>
__global__ void __launch_bounds__(512) probe(float* out, int iters, float seed) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = seed * (threadIdx.x + i) - 3.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      constexpr int NPOLY = MODE >= 3 ? (MODE - 3 + 1) : 0;   // pairs: i/2 in [0, 8)
      const bool poly_pair = (MODE == 2) || (MODE >= 3 && (i / 2) < NPOLY);
      float a = v[i] * 0.999f - 0.5f, b = v[i + 1] * 0.999f - 0.5f;    // keeps the chain bounded: 2^x <= 1 for x <= 0
      if (MODE == 1) { a = ex2_poly(a); b = ex2_poly(b); }
      else if (poly_pair) ex2_poly2(a, b, a, b);
      else { a = ex2_mufu(a); b = ex2_mufu(b); }
      v[i] = a; v[i + 1] = b;
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) reinterpret_cast<long long*>(out)[gridDim.x * blockDim.x / 2 + 1] = t1 - t0;
}

template <int MODE>
static void run(const char* name, float* d, int sms) {
  const int iters = 4096, ctas_per_sm = 2, threads = 512;   // 32 warps per SM: the pipes, not latency, are the limit
  probe<MODE><<<sms * ctas_per_sm, threads>>>(d, 16, 1e-3f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<MODE><<<sms * ctas_per_sm, threads>>>(d, iters, 1e-3f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long cyc; cudaMemcpy(&cyc, reinterpret_cast<long long*>(d) + (sms * ctas_per_sm * threads / 2 + 1), 8, cudaMemcpyDeviceToHost);
  const double exps_per_sm = double(iters) * 16 * threads * ctas_per_sm;
  printf("%-44s %8.3f ms  %7.2f exp2/clk/SM (CTA-0 clock64: %lld cycles)  %7.1f Gexp2/s whole GPU\n", name, ms,
         exps_per_sm / double(cyc), cyc, exps_per_sm * sms / ms / 1e6);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  printf("%s, %d SMs\n", p.name, sms);
  float* d; cudaMalloc(&d, (size_t)sms * 2 * 512 * 4 + 64);
  run<0>("MUFU.EX2 only", d, sms);
  run<1>("FMA-pipe cubic (scalar ex2_poly)", d, sms);
  run<2>("FMA-pipe cubic (packed f32x2)", d, sms);
  run<3>("1 of 8 pairs on the packed polynomial", d, sms);
  run<4>("2 of 8", d, sms);
  run<5>("3 of 8", d, sms);
  run<6>("4 of 8", d, sms);
  run<7>("5 of 8", d, sms);
  run<8>("6 of 8", d, sms);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
  return 0;
}
