// vdb200 — persistent tcgen05 implicit-GEMM mainloop (sm_100a).
//
// One kernel family serves every GEMM-shaped stage of the Versatile-Diffusion sampling path:
//   * Linear layers and 1x1 convs on NHWC activations (CrossAttention.to_q/k/v/to_out,
//     FeedForward/GEGLU, SpatialTransformer.proj_in/out — reference lib/model_zoo/attention.py:37-64,
//     152-193, 221-266; AutoencoderKL AttnBlock q/k/v/proj_out — autokl_modules.py:150-202),
//   * 3x3 convolutions as implicit GEMM over 9 filter taps (ResBlock in_layers[2]/out_layers[3],
//     Downsample.op, Upsample.conv — openaimodel.py:89-274; VAE ResnetBlock/Downsample/Upsample —
//     autokl_modules.py:42-141), with the 1x1 skip_connection of a channel-changing ResBlock folded
//     in as extra K segments of the same accumulator.
//
// Structure: grid = #SMs (persistent, static round-robin over output tiles), 320 threads:
//   warp 0   : TMA producer  (cp.async.bulk.tensor 4D box loads of A, 2D box loads of W)
//   warp 1   : MMA issuer    (tcgen05.mma cta_group::1 kind::f16, 128 x BN x 16, fp32 accum in TMEM)
//   warps 2-9: epilogue      (tcgen05.ld -> smem transpose -> bias/act/residual -> coalesced bf16 stores)
// smem ring of STAGES x (A 128x64 bf16 | W BNx64 bf16), both 128B-swizzled K-major; TMEM holds two
// accumulator stages (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
#include "host_util.h"

namespace vdb {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // bf16 elements = 128 B = one swizzle row
constexpr int kMaxA = 6;     // A tensor maps per launch
constexpr int kMaxSeg = 12;  // K segments per launch
constexpr int kNumEpiWarps = 8;
constexpr int kNumEpiThreads = kNumEpiWarps * 32;
constexpr int kNumThreads = 64 + kNumEpiThreads;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr uint32_t kABytes = kBlockM * kBlockK * 2;

struct ASeg {
  int16_t tmap;  // which A tensor map
  int16_t dw;    // W-coordinate shift of this segment (filter tap / parity lattice)
  int16_t dh;    // H-coordinate shift
  int16_t nkb;   // number of 64-channel k-blocks
  int32_t c0;    // first channel coordinate
};

struct alignas(64) IgemmParams {
  CUtensorMap tmA[kMaxA];  // 4D (C, W, H, B) bf16, box (64, TW, TH, TB), SWIZZLE_128B
  CUtensorMap tmB;         // 2D (Ktot, N) bf16, box (64, BN), SWIZZLE_128B
  ASeg seg[kMaxSeg];
  int nseg;
  int kb_total;      // total k-blocks over all segments
  int ksplit;        // split-K factor (>=1); >1 => fp32 partial output
  int kb_per_split;  // ceil(kb_total / ksplit)
  int TW, TH, TB;    // M tile = TW*TH*TB = 128 output pixels
  int Wo, Ho, Bo;    // output pixel grid (GEMM view: Wo = M, Ho = Bo = 1)
  int tilesW, tilesH, tilesB, tilesN;
  int N;             // valid output columns (GEGLU: packed columns, output has N/2)
  // epilogue
  const float* bias;          // [bias_rows, N] fp32 or null
  long long bias_bstride;     // 0: shared bias row; else stride between per-batch rows
  int rows_per_batch;         // output pixels per batch element (for bias_bstride != 0)
  const __nv_bfloat16* resid; // [M, ldr] bf16 or null (added after activation)
  long long ldr;
  void* out;                  // bf16 or fp32 [M, ldo]
  long long ldo;
  int out_f32;                // 1: fp32 output
  int act;                    // 0 none, 1 silu, 2 gelu(erf), 3 quick_gelu, 4 geglu (packed halves)
  float alpha;                // out = act(alpha * (acc + bias)) + resid
  float* partial;             // split-K: [ksplit, M, N] fp32
  int epi_alt;                // 1: the two warps of a TMEM quarter swap chunk parity every tile (odd chunk counts)
  unsigned long long* timeline; // debug: per-tile role timestamps of CTA 0 (null = off)
};

enum { ACT_NONE = 0, ACT_SILU = 1, ACT_GELU = 2, ACT_QGELU = 3, ACT_GEGLU = 4 };

VDB_DEVINL unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define VDB_TL(slot, it) do { if (p.timeline && blockIdx.x == 0 && (it) < 8) p.timeline[(it) * 16 + (slot)] = gtime(); } while (0)
#define VDB_TLE(slot, it) do { if (warp == 2 && lane == 0) VDB_TL(slot, it); } while (0)

VDB_DEVINL float apply_act(float v, int act) {
  switch (act) {
    case ACT_SILU: return silu_f(v);
    case ACT_GELU: return gelu_erf_f(v);
    case ACT_QGELU: return quick_gelu_f(v);
    default: return v;
  }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(kNumThreads, 1) igemm_kernel(const __grid_constant__ IgemmParams p) {
  constexpr uint32_t kBBytes = BN * kBlockK * 2;
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static_assert(kBBytes % 1024 == 0, "B stage must keep 1024B alignment");
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "invalid UMMA N");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + STAGES * kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* sbias = reinterpret_cast<float*>(tmem_holder + 4);   // [BN] bias of the current output tile
  float* sstage = sbias + BN;                                  // kNumEpiWarps x [32][32] fp32 swizzled transposition tiles
  auto epi_bar_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(kNumEpiThreads) : "memory"); };   // the epilogue warps only

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kMaxA; ++i) tma_prefetch_desc(&p.tmA[i]);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kNumEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_launch_dependents();
  pdl_wait();   // everything above overlapped the previous kernel's tail; global inputs are valid from here

  const int tilesM = p.tilesW * p.tilesH * p.tilesB;
  const int num_tiles = tilesM * p.tilesN * p.ksplit;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m_idx = t % tilesM;
        const int rest = t / tilesM;
        const int n_idx = rest % p.tilesN;
        const int ks = rest / p.tilesN;
        const int wt = m_idx % p.tilesW;
        const int ht = (m_idx / p.tilesW) % p.tilesH;
        const int bt = m_idx / (p.tilesW * p.tilesH);
        const int w0 = wt * p.TW, h0 = ht * p.TH, b0 = bt * p.TB;
        const int n0 = n_idx * BN;
        const int kb_begin = ks * p.kb_per_split;
        const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
        int kb = 0;
        VDB_TL(0, (t - blockIdx.x) / gridDim.x);   // producer: starts issuing this tile
        for (int s = 0; s < p.nseg; ++s) {
          const ASeg sg = p.seg[s];
          if (kb + sg.nkb <= kb_begin) { kb += sg.nkb; continue; }
          for (int j = 0; j < sg.nkb; ++j, ++kb) {
            if (kb < kb_begin) continue;
            if (kb >= kb_end) break;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
            tma_load_4d(smemA + stage * kABytes, &p.tmA[sg.tmap], &full_bar[stage],
                        sg.c0 + j * kBlockK, w0 + sg.dw, h0 + sg.dh, b0);
            tma_load_2d(smemB + stage * kBBytes, &p.tmB, &full_bar[stage], kb * kBlockK, n0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (kb >= kb_end) break;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BN);
      uint32_t stage = 0, phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int ks = (t / tilesM) / p.tilesN;
        const int kb_begin = ks * p.kb_per_split;
        const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        VDB_TL(1, it);                             // MMA: wants the accumulator stage
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        VDB_TL(2, it);                             // MMA: got it
        const uint32_t tmem_d = tmem_base + as * 256;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = make_desc_sw128(smem_u32(smemA + stage * kABytes));
          const uint64_t bdesc = make_desc_sw128(smem_u32(smemB + stage * kBBytes));
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the swizzle atom: +2 in (addr >> 4) units
            umma_bf16_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[as]);
        VDB_TL(3, it);                             // MMA: all MMAs of the tile issued
      }
    }
  } else {
    // ------------------------------ epilogue ------------------------------
    // Eight warps: warps w and w+4 own the same TMEM lane quarter (w & 3) and alternate 32-column chunks, so each
    // SM sub-partition interleaves two epilogue warps (a single warp per sub-partition was measured to be
    // instruction-latency bound: ~770 ns per chunk).  tcgen05.ld hands each thread one ROW of 32 columns; the
    // warp writes the raw fp32 block to a private XOR-swizzled 4 KB shared-memory tile and reads it back
    // transposed, so every global access is 8 rows x 64 contiguous bytes per instruction (4 lanes per row), and
    // bias / activation / residual / bf16 conversion run on 8 fixed columns per lane (bias lives in registers).
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;      // which of the two warps of the quarter
    const int r = quarter * 32 + lane;
    const int tr_row = lane >> 2, tr_q = lane & 3;   // transposed role: rows tr_row + 8k, columns tr_q*8 .. +7
    float* stage = sstage + (warp - 2) * 1024;       // [32 rows][32 fp32], 16-byte chunk j of row i at (j ^ (i & 7))
    auto stage_write = [&](const uint32_t (&v)[32]) {
      float* srow = stage + lane * 32;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(srow + ((j ^ (lane & 7)) << 2)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    };
    auto stage_read = [&](int k, float (&o)[8]) {   // row tr_row + 8k, columns tr_q*8 .. +7
      const int row = k * 8 + tr_row;
      const float* srow = stage + row * 32;
      const float4 x0 = *reinterpret_cast<const float4*>(srow + (((2 * tr_q) ^ (row & 7)) << 2));
      const float4 x1 = *reinterpret_cast<const float4*>(srow + (((2 * tr_q + 1) ^ (row & 7)) << 2));
      o[0] = x0.x; o[1] = x0.y; o[2] = x0.z; o[3] = x0.w; o[4] = x1.x; o[5] = x1.y; o[6] = x1.z; o[7] = x1.w;
    };
    int it = 0;
    const float* sbias_src = nullptr;   // which bias row/offset currently sits in sbias
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int m_idx = t % tilesM;
      const int rest = t / tilesM;
      const int n_idx = rest % p.tilesN;
      const int ks = rest / p.tilesN;
      const int wt = m_idx % p.tilesW;
      const int ht = (m_idx / p.tilesW) % p.tilesH;
      const int bt = m_idx / (p.tilesW * p.tilesH);
      const int tw = r % p.TW;
      const int th = (r / p.TW) % p.TH;
      const int tb = r / (p.TW * p.TH);
      const int w = wt * p.TW + tw, h = ht * p.TH + th, b = bt * p.TB + tb;
      const bool row_ok = (w < p.Wo) && (h < p.Ho) && (b < p.Bo);
      const long long gp = (static_cast<long long>(b) * p.Ho + h) * p.Wo + w;
      const long long gp_first = (static_cast<long long>(bt * p.TB) * p.Ho + ht * p.TH) * p.Wo + wt * p.TW;
      const int n0 = n_idx * BN;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;

      // per-tile row bookkeeping for the transposed role (independent of the accumulator: done before the wait)
      long long gp_k[4];
      bool ok_k[4];
      {
        const unsigned okmask = __ballot_sync(0xffffffffu, row_ok);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int src = k * 8 + tr_row;
          const unsigned lo = __shfl_sync(0xffffffffu, static_cast<unsigned>(gp & 0xffffffffu), src);
          const unsigned hi = __shfl_sync(0xffffffffu, static_cast<unsigned>(static_cast<unsigned long long>(gp) >> 32), src);
          gp_k[k] = static_cast<long long>((static_cast<unsigned long long>(hi) << 32) | lo);
          ok_k[k] = (okmask >> src) & 1u;
        }
      }
      // bias tile -> shared memory (one global read per tile); one row serves the tile unless the bias is per-batch
      // and the tile's first / last rows belong to different batch items
      const long long gp_last = (static_cast<long long>(bt * p.TB + p.TB - 1) * p.Ho + ht * p.TH + p.TH - 1) * p.Wo +
                                wt * p.TW + p.TW - 1;
      const bool bias_uniform = p.bias && p.ksplit == 1 &&
                                (p.bias_bstride == 0 || gp_first / p.rows_per_batch == gp_last / p.rows_per_batch);
      if (bias_uniform) {
        // consecutive tiles of a CTA usually share the N tile (M is the fast tile index): reload only on change,
        // otherwise the ~0.7 us global-load latency + two barriers sit between every two tiles
        const float* brow = p.bias + (p.bias_bstride ? (gp_first / p.rows_per_batch) * p.bias_bstride : 0) + n0;
        if (brow != sbias_src) {              // uniform across the epilogue threads
          epi_bar_sync();                     // previous tile's readers are done with sbias
          for (int i = threadIdx.x - 64; i < BN; i += kNumEpiThreads) sbias[i] = (n0 + i < p.N) ? __ldg(brow + i) : 0.f;
          epi_bar_sync();
          sbias_src = brow;
        }
      }
      const float* bias_g = (p.bias && !bias_uniform)
                                ? p.bias + (p.bias_bstride ? (gp / p.rows_per_batch) * p.bias_bstride : 0) : nullptr;

      VDB_TLE(4, it);   // epilogue: waiting for the accumulator
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      VDB_TLE(5, it);   // epilogue: accumulator complete
      const uint32_t trow = tmem_base + as * 256 + (static_cast<uint32_t>(quarter * 32) << 16);

      if (p.ksplit > 1) {
        // fp32 partials, reduced (+bias/act/residual) by splitk_reduce_kernel
        const long long Mtot = static_cast<long long>(p.Bo) * p.Ho * p.Wo;
        float* dst = p.partial + (static_cast<long long>(ks) * Mtot + gp) * p.N + n0;
#pragma unroll 1
        for (int c = half; c < BN / 32; c += 2) {
          if (n0 + c * 32 >= p.N) break;
          uint32_t v[32];
          tmem_ld32(trow + c * 32, v);
          tmem_wait_ld();
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const int n = n0 + c * 32 + j;
              if (n + 3 < p.N) {
                *reinterpret_cast<float4*>(dst + c * 32 + j) =
                    make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
              } else {
                for (int q = 0; q < 4; ++q)
                  if (n + q < p.N) dst[c * 32 + j + q] = __uint_as_float(v[j + q]);
              }
            }
          }
        }
      } else if (p.act == ACT_GEGLU) {
        // packed tile: columns [0,BN/2) = value rows, [BN/2,BN) = gate rows of the same outputs
        constexpr int HALF = BN / 2;
        const int nout0 = n_idx * HALF;
        const int Nout = p.N / 2;
#pragma unroll 1
        for (int c = half; c < HALF / 32; c += 2) {
          uint32_t v[32];
          float a[4][8];
          tmem_ld32(trow + c * 32, v);
          tmem_wait_ld();
          stage_write(v);
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 4; ++k) stage_read(k, a[k]);
          __syncwarp();
          tmem_ld32(trow + HALF + c * 32, v);
          tmem_wait_ld();
          stage_write(v);
          __syncwarp();
          float bv[8], bg[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            bv[i] = p.bias ? sbias[c * 32 + tr_q * 8 + i] : 0.f;
            bg[i] = p.bias ? sbias[HALF + c * 32 + tr_q * 8 + i] : 0.f;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float g[8];
            stage_read(k, g);
            if (ok_k[k] && nout0 + c * 32 + tr_q * 8 + 7 < Nout) {
              float o[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = (a[k][i] + bv[i]) * gelu_fast_f(g[i] + bg[i]);
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + gp_k[k] * p.ldo + nout0 + c * 32 + tr_q * 8) =
                  make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
            }
          }
          __syncwarp();
        }
      } else {
        const int nchunks = min(BN / 32, (p.N - n0 + 31) / 32);
        auto chunk = [&](int c, const uint32_t (&v)[32], const uint4 (&rr)[4], bool fast) {
          const int nb = n0 + c * 32;
          if (fast) {
            stage_write(v);
            __syncwarp();
            if (c == half) VDB_TLE(8, it);
            float bb[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) bb[i] = p.bias ? sbias[c * 32 + tr_q * 8 + i] : 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float o[8];
              stage_read(k, o);
              if (ok_k[k]) {
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] += bb[i];
                if (p.alpha != 1.f) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) o[i] *= p.alpha;
                }
                if (p.act != ACT_NONE) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) o[i] = apply_act(o[i], p.act);
                }
                if (p.resid) {
                  const uint32_t w4[4] = {rr[k].x, rr[k].y, rr[k].z, rr[k].w};
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    const float2 x = unpack_bf16x2(w4[q]);
                    o[2 * q] += x.x;
                    o[2 * q + 1] += x.y;
                  }
                }
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + gp_k[k] * p.ldo + nb + tr_q * 8) =
                    make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
              }
              if (c == half && k == 0) VDB_TLE(9, it);
            }
            __syncwarp();   // the tile is rewritten by this warp's next chunk
            if (c == half) VDB_TLE(10, it);
          } else if (row_ok) {
            // slow path (fp32 output, partial last chunk, per-row bias): one row per thread
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nb + j < p.N) f[j] += bias_uniform ? sbias[c * 32 + j] : __ldg(bias_g + nb + j);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = apply_act(f[j] * p.alpha, p.act);
            if (p.resid) {
              const __nv_bfloat16* rs = p.resid + gp * p.ldr + nb;
#pragma unroll
              for (int j = 0; j < 32; ++j) if (nb + j < p.N) f[j] += __bfloat162float(rs[j]);
            }
            if (p.out_f32) {
              float* dst = reinterpret_cast<float*>(p.out) + gp * p.ldo + nb;
#pragma unroll
              for (int j = 0; j < 32; ++j) if (nb + j < p.N) dst[j] = f[j];
            } else {
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + gp * p.ldo + nb;
#pragma unroll
              for (int j = 0; j < 32; ++j) if (nb + j < p.N) dst[j] = __float2bfloat16(f[j]);
            }
          }
        };
        auto is_fast = [&](int c) { return (n0 + c * 32 + 32 <= p.N) && !p.out_f32 && (bias_uniform || !p.bias); };
        auto load_resid = [&](int c, uint4 (&rr)[4]) {
          if (p.resid && is_fast(c)) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (ok_k[k]) rr[k] = __ldg(reinterpret_cast<const uint4*>(p.resid + gp_k[k] * p.ldr + n0 + c * 32 + tr_q * 8));
          }
        };
        // this warp's chunks: half, half+2, ... (kept un-pipelined: double-buffering the 32-register TMEM chunk
        // pushed the kernel into spills and was measured slower)
        // odd chunk count (BN = 160: five): the warp that took three chunks on this tile takes two on the next one
        const int first = (p.epi_alt & (BN / 32) & 1) ? (half ^ (it & 1)) : half;
#pragma unroll 1
        for (int c = first; c < nchunks; c += 2) {
          uint4 rr[4];
          load_resid(c, rr);          // residual loads are in flight while the accumulator chunk is fetched
          uint32_t v[32];
          tmem_ld32(trow + c * 32, v);
          tmem_wait_ld();
          if (c == first) VDB_TLE(7, it);
          chunk(c, v, rr, is_fast(c));
        }
      }
      tc_fence_before();
      __syncwarp();
      VDB_TLE(6, it);   // epilogue: tile stored (this warp)
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// split-K reduction + epilogue: out[m, n] = act(alpha * (sum_s partial[s, m, n] + bias)) + resid
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int ksplit, long long M, int N,
                                     const float* __restrict__ bias, long long bias_bstride, int rows_per_batch,
                                     const __nv_bfloat16* __restrict__ resid, long long ldr, void* out,
                                     long long ldo, int out_f32, int act, float alpha) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = M * (N / 4);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / (N / 4);
    const int n = static_cast<int>(i % (N / 4)) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < ksplit; ++s) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(partial + (static_cast<long long>(s) * M + m) * N + n));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float f[4] = {acc.x, acc.y, acc.z, acc.w};
    if (bias) {
      const float* bp = bias + (bias_bstride ? (m / rows_per_batch) * bias_bstride : 0) + n;
      for (int q = 0; q < 4; ++q) f[q] += __ldg(bp + q);
    }
    for (int q = 0; q < 4; ++q) f[q] *= alpha;
    for (int q = 0; q < 4; ++q) f[q] = apply_act(f[q], act);
    if (resid) {
      for (int q = 0; q < 4; ++q) f[q] += __bfloat162float(resid[m * ldr + n + q]);
    }
    if (out_f32) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + m * ldo + n) = make_float4(f[0], f[1], f[2], f[3]);
    } else {
      uint2 o = make_uint2(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]));
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + m * ldo + n) = o;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
template <int BN, int STAGES>
static int launch_igemm(const IgemmParams& p, int num_tiles, cudaStream_t stream) {
  constexpr size_t smem = STAGES * (kABytes + BN * kBlockK * 2) + (2 * STAGES + 4) * 8 + 16 + BN * 4 + kNumEpiWarps * 4096 + 1024;
  static bool configured = false;
  if (!configured) {
    VDB_CUDA_CHECK(cudaFuncSetAttribute(igemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    prefer_max_smem(igemm_kernel<BN, STAGES>);
    configured = true;
  }
  const int grid = std::min(num_tiles, num_sms());
  VDB_CUDA_CHECK(launch_pdl(igemm_kernel<BN, STAGES>, dim3(grid), dim3(kNumThreads), smem, stream, p));
  count_launch();
  return VDB_OK;
}

static int pick_bn(int N, int act, int forced) {
  if (forced) return forced;
  if (act == ACT_GEGLU) return 256;
  if (N <= 64) return 64;
  if (N % 256 == 0) return 256;
  if (N % 160 == 0) return 160;
  if (N % 128 == 0) return 128;
  if (N <= 128) return 128;
  if (N <= 160) return 160;
  return 256;
}

static unsigned long long* g_timeline = nullptr;

struct IgemmEpilogue {
  const float* bias = nullptr;
  long long bias_bstride = 0;
  int rows_per_batch = 1;
  const void* resid = nullptr;
  long long ldr = 0;
  void* out = nullptr;
  long long ldo = 0;
  int out_f32 = 0;
  int act = 0;
  float alpha = 1.f;
};

// Finish IgemmParams (tiling, split-K, B map) and launch.
static int run_igemm(IgemmParams& p, const void* Wt, long long N, long long Ktot, long long ldw,
                     const IgemmEpilogue& e, int bn_forced, int ksplit_forced, void* workspace,
                     size_t ws_bytes, cudaStream_t stream) {
  int BN = pick_bn(static_cast<int>(N), e.act, bn_forced);
  if (!bn_forced && e.act != ACT_GEGLU && p.kb_total < 32) {
    // short K and a small MN grid (the 8x8 level): narrower tiles fill more SMs and need no split-K reduction pass
    // (M 512, N 1280, K 1280: 10.7 us with BN 64 vs 18.1 us with BN 256 + split-K 2, tools/bn_sweep.py)
    const int tm = p.tilesW * p.tilesH * p.tilesB;
    auto tiles = [&](int bn) { return tm * static_cast<int>((N + bn - 1) / bn); };
    while (BN > 64 && tiles(BN) * 2 <= num_sms()) BN = (BN == 256) ? 160 : (BN == 160 ? 128 : 64);
  }
  if (BN != 64 && BN != 128 && BN != 160 && BN != 256) return set_error(VDB_ERR_INVALID, "igemm: bad BN");
  if (e.act == ACT_GEGLU && (N % BN) != 0) return set_error(VDB_ERR_INVALID, "igemm: GEGLU needs N % 256 == 0");
  p.N = static_cast<int>(N);
  p.tilesN = static_cast<int>((N + BN - 1) / BN);
  p.bias = e.bias; p.bias_bstride = e.bias_bstride; p.rows_per_batch = e.rows_per_batch > 0 ? e.rows_per_batch : 1;
  p.resid = reinterpret_cast<const __nv_bfloat16*>(e.resid); p.ldr = e.ldr;
  p.out = e.out; p.ldo = e.ldo; p.out_f32 = e.out_f32; p.act = e.act; p.alpha = e.alpha;
  if (!e.out_f32 && (e.ldo % 8)) return set_error(VDB_ERR_INVALID, "igemm: ldo must be a multiple of 8 for bf16 out");
  if (e.resid && (e.ldr % 8)) return set_error(VDB_ERR_INVALID, "igemm: ldr must be a multiple of 8");
  int rc = make_tmap_2d(&p.tmB, Wt, static_cast<uint64_t>(Ktot), static_cast<uint64_t>(N),
                        static_cast<uint64_t>(ldw) * 2, kBlockK, BN);
  if (rc) return rc;
  const long long M = static_cast<long long>(p.Bo) * p.Ho * p.Wo;
  const int tilesM = p.tilesW * p.tilesH * p.tilesB;
  const int mn_tiles = tilesM * p.tilesN;
  // split-K heuristic: fill the machine when the MN grid is small and K is deep
  int ksplit = 1;
  if (ksplit_forced > 0) {
    ksplit = ksplit_forced;
  } else if (e.act != ACT_GEGLU && mn_tiles * 2 <= num_sms() && p.kb_total >= 16 && (N % 4) == 0) {
    ksplit = std::min(std::min(num_sms() / mn_tiles, p.kb_total / 8), 16);
    if (ksplit < 1) ksplit = 1;
  }
  if (ksplit > 1) {
    const size_t need = static_cast<size_t>(ksplit) * M * N * sizeof(float);
    if (workspace == nullptr || ws_bytes < need || e.act == ACT_GEGLU || (N % 4)) ksplit = 1;
  }
  p.ksplit = ksplit;
  p.kb_per_split = (p.kb_total + ksplit - 1) / ksplit;
  // drop empty trailing splits
  p.ksplit = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.partial = reinterpret_cast<float*>(workspace);
  p.timeline = g_timeline;
  static const int epi_alt = [] { const char* ev = getenv("VDB_EPI_ALT"); return (ev && ev[0] == '0') ? 0 : 1; }();
  p.epi_alt = epi_alt;
  const int num_tiles = mn_tiles * p.ksplit;
  switch (BN) {
    case 64: rc = launch_igemm<64, 8>(p, num_tiles, stream); break;
    case 128: rc = launch_igemm<128, 6>(p, num_tiles, stream); break;
    case 160: rc = launch_igemm<160, 5>(p, num_tiles, stream); break;
    default: rc = launch_igemm<256, 4>(p, num_tiles, stream); break;
  }
  if (rc) return rc;
  if (p.ksplit > 1) {
    const long long total = M * (N / 4);
    const int threads = 256;
    const int blocks = static_cast<int>(std::min<long long>((total + threads - 1) / threads, num_sms() * 8LL));
    VDB_PREFER_MAX_SMEM(splitk_reduce_kernel);
    VDB_CUDA_CHECK(launch_pdl(splitk_reduce_kernel, dim3(blocks), dim3(threads), 0, stream, (const float*)p.partial,
                              p.ksplit, M, static_cast<int>(N), p.bias, p.bias_bstride, p.rows_per_batch, p.resid,
                              p.ldr, p.out, p.ldo, p.out_f32, p.act, p.alpha));
    count_launch();
  }
  return VDB_OK;
}

static int pow2_ceil(int v) { int t = 1; while (t < v) t <<= 1; return t; }

// M tile = 128 output pixels as a (TW, TH, TB) box of the (W, H, B) pixel grid; box extents are
// powers of two so that TW*TH*TB == 128 (rows past the grid are zero-filled by TMA and masked on store).
static void set_tile_shape(IgemmParams& p, int Wo, int Ho, int Bo) {
  p.Wo = Wo; p.Ho = Ho; p.Bo = Bo;
  const int TW = std::min(pow2_ceil(Wo), 128);
  const int TH = std::min(pow2_ceil(Ho), 128 / TW);
  const int TB = 128 / (TW * TH);
  p.TW = TW; p.TH = TH; p.TB = TB;
  p.tilesW = (Wo + TW - 1) / TW;
  p.tilesH = (Ho + TH - 1) / TH;
  p.tilesB = (Bo + TB - 1) / TB;
}

}  // namespace vdb

using namespace vdb;

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

// debug aid (not part of the product ABI): device buffer of 16*8 u64 receiving CTA 0's per-tile role timestamps
void vdb_debug_igemm_timeline(void* buf) { g_timeline = reinterpret_cast<unsigned long long*>(buf); }

// out[M,N] = act(alpha * ([A | A2] @ W^T + bias)) + resid     (see include/vdb200.h)
int vdb_gemm_bf16(const void* A, long long M, long long K, long long lda, const void* A2, long long K2,
                  long long lda2, const void* W, long long N, long long ldw, const float* bias,
                  long long bias_bstride, long long rows_per_batch, const void* resid, long long ldr, void* out,
                  long long ldo, int out_f32, int act, float alpha, int bn, int ksplit, void* workspace,
                  size_t ws_bytes, void* stream) {
  if (!A || !W || !out || M <= 0 || N <= 0 || K <= 0) return set_error(VDB_ERR_INVALID, "gemm: null/empty argument");
  if ((K % 8) || (lda % 8) || (ldw % 8)) return set_error(VDB_ERR_INVALID, "gemm: K, lda, ldw must be multiples of 8");
  if (A2 && ((K % kBlockK) || (K2 % 8) || (lda2 % 8)))
    return set_error(VDB_ERR_INVALID, "gemm: two-source A needs K % 64 == 0 and K2, lda2 % 8 == 0");
  if (M > 0x7fffffffLL || N > 0x7fffffffLL) return set_error(VDB_ERR_INVALID, "gemm: dimension too large");
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  // GEMM view of the pixel grid: one row of M "pixels"; the box is always 128 rows (TMA zero-fills past M)
  p.Wo = static_cast<int>(M); p.Ho = 1; p.Bo = 1;
  p.TW = kBlockM; p.TH = 1; p.TB = 1;
  p.tilesW = static_cast<int>((M + kBlockM - 1) / kBlockM); p.tilesH = 1; p.tilesB = 1;
  int rc = make_tmap_4d(&p.tmA[0], A, K, M, 1, 1, lda * 2, lda * 2 * M, lda * 2 * M, kBlockK, p.TW, 1, 1);
  if (rc) return rc;
  p.seg[0] = ASeg{0, 0, 0, static_cast<int16_t>((K + kBlockK - 1) / kBlockK), 0};
  p.nseg = 1;
  p.kb_total = p.seg[0].nkb;
  if (A2) {
    rc = make_tmap_4d(&p.tmA[1], A2, K2, M, 1, 1, lda2 * 2, lda2 * 2 * M, lda2 * 2 * M, kBlockK, p.TW, 1, 1);
    if (rc) return rc;
    p.seg[1] = ASeg{1, 0, 0, static_cast<int16_t>((K2 + kBlockK - 1) / kBlockK), 0};
    p.nseg = 2;
    p.kb_total += p.seg[1].nkb;
  }
  for (int i = p.nseg; i < kMaxA; ++i) p.tmA[i] = p.tmA[0];
  IgemmEpilogue e;
  e.bias = bias; e.bias_bstride = bias_bstride; e.rows_per_batch = static_cast<int>(rows_per_batch);
  e.resid = resid; e.ldr = ldr; e.out = out; e.ldo = ldo; e.out_f32 = out_f32; e.act = act; e.alpha = alpha;
  return run_igemm(p, W, N, K + (A2 ? K2 : 0), ldw, e, bn, ksplit, workspace, ws_bytes,
                   reinterpret_cast<cudaStream_t>(stream));
}

// 3x3 convolution on NHWC bf16 as implicit GEMM.
//   mode 0: stride 1, pad 1                       (out H x W)
//   mode 1: stride 2, pad 1                       (out H/2 x W/2)      openaimodel.py:150-152
//   mode 2: stride 2, pad (0,1,0,1) then pad 0    (out H/2 x W/2)      autokl_modules.py:72-76
// Wt is [N, 9*C + Cs1 + Cs2] bf16 with K ordered (ky, kx, c) then the 1x1-skip columns.
// skip1/skip2: optional raw NHWC tensors at OUTPUT resolution whose 1x1 conv is accumulated too.
int vdb_conv3x3_bf16(const void* X, int B, int H, int Wd, int C, int mode, const void* Wt, int N, long long ldw,
                     const void* skip1, int Cs1, const void* skip2, int Cs2, const float* bias,
                     long long bias_bstride, const void* resid, long long ldr, void* out, long long ldo,
                     int out_f32, int act, int bn, int ksplit, void* workspace, size_t ws_bytes, void* stream) {
  if (!X || !Wt || !out || B <= 0 || H <= 0 || Wd <= 0 || C <= 0 || N <= 0)
    return set_error(VDB_ERR_INVALID, "conv3x3: null/empty argument");
  if ((C % kBlockK) || (ldw % 8)) return set_error(VDB_ERR_INVALID, "conv3x3: C must be a multiple of 64, ldw of 8");
  if ((skip1 && (Cs1 % kBlockK)) || (skip2 && (Cs2 % kBlockK)))
    return set_error(VDB_ERR_INVALID, "conv3x3: skip channels must be multiples of 64");
  if (mode < 0 || mode > 2) return set_error(VDB_ERR_INVALID, "conv3x3: bad mode");
  if (mode != 0 && ((H & 1) || (Wd & 1))) return set_error(VDB_ERR_UNSUPPORTED, "conv3x3: stride 2 needs even H, W");
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  const int Ho = mode ? H / 2 : H, Wo = mode ? Wd / 2 : Wd;
  set_tile_shape(p, Wo, Ho, B);
  const uint64_t eb = 2;
  const int nkb = C / kBlockK;
  int rc;
  int nmaps = 0;
  if (mode == 0) {
    rc = make_tmap_4d(&p.tmA[0], X, C, Wd, H, B, C * eb, (uint64_t)Wd * C * eb, (uint64_t)H * Wd * C * eb, kBlockK,
                      p.TW, p.TH, p.TB);
    if (rc) return rc;
    nmaps = 1;
    for (int t = 0; t < 9; ++t)
      p.seg[t] = ASeg{0, static_cast<int16_t>(t % 3 - 1), static_cast<int16_t>(t / 3 - 1), static_cast<int16_t>(nkb), 0};
  } else {
    // four parity sub-lattices of the input: X[b, 2*yo+py, 2*xo+px, c]
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        const uint8_t* base = reinterpret_cast<const uint8_t*>(X) + (static_cast<uint64_t>(py) * Wd + px) * C * eb;
        rc = make_tmap_4d(&p.tmA[py * 2 + px], base, C, Wo, Ho, B, 2ull * C * eb, 2ull * Wd * C * eb,
                          (uint64_t)H * Wd * C * eb, kBlockK, p.TW, p.TH, p.TB);
        if (rc) return rc;
      }
    nmaps = 4;
    for (int t = 0; t < 9; ++t) {
      const int ky = t / 3, kx = t % 3;
      int py, dy, px, dx;
      if (mode == 1) {  // input row = 2*yo + ky - 1
        py = (ky == 1) ? 0 : 1; dy = (ky == 0) ? -1 : 0;
        px = (kx == 1) ? 0 : 1; dx = (kx == 0) ? -1 : 0;
      } else {          // input row = 2*yo + ky (zero pad on bottom/right only)
        py = (ky == 1) ? 1 : 0; dy = (ky == 2) ? 1 : 0;
        px = (kx == 1) ? 1 : 0; dx = (kx == 2) ? 1 : 0;
      }
      p.seg[t] = ASeg{static_cast<int16_t>(py * 2 + px), static_cast<int16_t>(dx), static_cast<int16_t>(dy),
                      static_cast<int16_t>(nkb), 0};
    }
  }
  p.nseg = 9;
  p.kb_total = 9 * nkb;
  const void* sk[2] = {skip1, skip2};
  const int sc[2] = {Cs1, Cs2};
  for (int i = 0; i < 2; ++i) {
    if (!sk[i]) continue;
    rc = make_tmap_4d(&p.tmA[nmaps], sk[i], sc[i], Wo, Ho, B, sc[i] * eb, (uint64_t)Wo * sc[i] * eb,
                      (uint64_t)Ho * Wo * sc[i] * eb, kBlockK, p.TW, p.TH, p.TB);
    if (rc) return rc;
    p.seg[p.nseg] = ASeg{static_cast<int16_t>(nmaps), 0, 0, static_cast<int16_t>(sc[i] / kBlockK), 0};
    p.kb_total += sc[i] / kBlockK;
    ++p.nseg;
    ++nmaps;
  }
  for (int i = nmaps; i < kMaxA; ++i) p.tmA[i] = p.tmA[0];
  IgemmEpilogue e;
  e.bias = bias; e.bias_bstride = bias_bstride; e.rows_per_batch = Ho * Wo;
  e.resid = resid; e.ldr = ldr; e.out = out; e.ldo = ldo; e.out_f32 = out_f32; e.act = act; e.alpha = 1.f;
  return run_igemm(p, Wt, N, static_cast<long long>(p.kb_total) * kBlockK, ldw, e, bn, ksplit, workspace, ws_bytes,
                   reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
