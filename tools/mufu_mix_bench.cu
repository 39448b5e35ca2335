// How fast can ONE warp (or a few) per SM sub-partition feed the MUFU pipe on sm_100a, alone and next to other work?
// The two-tile attention kernel serialises the exp2 phases of its two softmax warps per sub-partition; its role timeline
// (profiles/r02_attention_fa_timeline_v2.txt) shows such a warp needs 800 ns for 112 MUFU.EX2 + ~300 other instructions
// where the pipe alone would need 500 ns.  This probe separates the candidates:
//   mix 0: MUFU.EX2 only (32 independent registers)           mix 1: the softmax tile body (FFMA2 scale, EX2, FADD2 sum, bf16 pack)
//   mix 2: even warps MUFU only, odd warps the ALU part only  mix 3: ALU part only
// for W = 1, 2, 4 warps per sub-partition (block = 128 W threads, one block per SM).  Prints exp2 per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mufu_mix_bench tools/mufu_mix_bench.cu && tools/bin/mufu_mix_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "../versatile-diffusion_b200/csrc/common.cuh"
using namespace vdb;

template <int MIX>
__global__ void __launch_bounds__(512, 1) probe(float* out, unsigned* outp, int iters, float seed, long long* clk) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = -0.01f * (1.f + seed * ((threadIdx.x + i) & 15));
  const int warp = threadIdx.x >> 5;
  const bool mufu_role = (MIX == 0) || (MIX == 1) || (MIX == 2 && (warp & 4) == 0);     // (warp & 3) = sub-partition
  const bool alu_role = (MIX == 1) || (MIX == 3) || (MIX == 2 && (warp & 4) != 0);
  const unsigned long long sc2 = pack_f2(0.999f, 0.999f), nm2 = pack_f2(-0.01f, -0.01f);
  unsigned long long l2 = pack_f2(0.f, 0.f);
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      float a = v[i], b = v[i + 1];
      if (alu_role) unpack_f2(fma_f2(pack_f2(a, b), sc2, nm2), a, b);
      if (mufu_role) { a = ex2_mufu(a); b = ex2_mufu(b); a = a - 1.5f; b = b - 1.5f; }    // (keeps the chain in range; 2 FADD)
      if (alu_role) {
        l2 = add_f2(l2, pack_f2(a, b));
        acc ^= pack_bf16x2(a, b);
      }
      v[i] = a; v[i + 1] = b;
    }
  }
  const long long t1 = clock64();
  float la, lb;
  unpack_f2(l2, la, lb);
  float s = la + lb;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  outp[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MIX>
static void run(const char* name, int W, float* out, unsigned* outp, long long* clk) {
  const int iters = 2000, threads = 128 * W, blocks = 148;
  probe<MIX><<<blocks, threads>>>(out, outp, 10, 0.5f, clk);
  cudaDeviceSynchronize();
  probe<MIX><<<blocks, threads>>>(out, outp, iters, 0.5f, clk);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0;
  for (int i = 0; i < 148; ++i) c += h[i];
  c /= 148;
  const int mufu_warps = (MIX == 2) ? W * 2 : (MIX == 3 ? 0 : W * 4);   // warps per SM that issue MUFU
  const double exps = static_cast<double>(iters) * 32 * 32 * mufu_warps;
  const double instr_slots = static_cast<double>(iters) * 16 * ((MIX == 0) ? 4 : (MIX == 1 ? 8 : (MIX == 3 ? 4 : 4))) * W;   // per sub-partition, rough
  printf("%-44s W=%d  %8.0f clk/iter-block  exp2/clk/SM %6.2f  (issue slots/clk/SMSP ~%.2f)\n", name, W, c / iters, exps / c, instr_slots / c);
}

int main() {
  float* out; unsigned* outp; long long* clk;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&outp, 148 * 512 * 4); cudaMalloc(&clk, 148 * 8);
  for (int W : {1, 2, 4}) run<0>("mix0 MUFU + 2 FADD per pair only", W, out, outp, clk);
  for (int W : {1, 2, 4}) run<1>("mix1 softmax body (FFMA2, 2 EX2, FADD2, pack)", W, out, outp, clk);
  for (int W : {2, 4}) run<2>("mix2 half the warps MUFU, half ALU", W, out, outp, clk);
  for (int W : {1, 2}) run<3>("mix3 ALU part only", W, out, outp, clk);
  cudaError_t e = cudaGetLastError();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
