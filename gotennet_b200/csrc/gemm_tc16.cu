// tcgen05 split-fp16 GEMM for sm_100a: fp32-accurate tensor-core GEMM at the 16-bit MMA rate
//     x' = x * 2^s (per-tensor power-of-two scale from the tensor's max |x|, so max |x'| is in [2^14, 2^15))
//     x_hi = fp16(x'),  x_lo = fp16(x' - x_hi)          (22 significant bits; the absolute error floor of the
//                                                        fp16 subnormals is 2^-40 of the tensor maximum)
//     D = (A_lo B_hi + A_hi B_lo + A_hi B_hi) * 2^-(sA+sB)        three kind::f16 MMAs, fp32 accumulation in TMEM
// (the reference runs its nn.Linear layers in strict fp32, scripts/train.py:16; three tf32 MMAs - gemm_tc.cu - cost
//  two 16-bit MMA slots each, this form costs one each: same accuracy class at twice the tensor-pipe throughput).
//
// Structure (persistent, warp specialised, one CTA per SM, 320 threads; 448 with eight epilogue warps for the
// epilogue-bound short-K shapes, template parameter EPIW) as in gemm_tc.cu:
//   warp 0      TMA producer: raw fp32 A tile + pre-split fp16 B_hi / B_lo tiles -> smem ring
//   warp 1      TMEM allocator + MMA issuer: 12 x tcgen05.mma.kind::f16 (M128/256 x N<=256 x K16) per 64-deep k-block
//   warps 2-5   epilogue: tcgen05.ld -> un-scale -> smem -> TMA store / fused bias + residual + SiLU side output
//   warps 6-9   converter: scales and splits the raw A tile IN PLACE into K-major 128B-swizzled fp16 hi / lo tiles.
//               For the weight-gradient GEMM (A given as [R][M], reduction over rows) the converter also transposes,
//               so every MMA of this file uses the plain K-major SWIZZLE_128B operand layout.
// B is always staged K-major: the pre-split kernels write fp16 hi / lo copies [N][K] (transposing when B is [K][N]).
#include <cuda_fp16.h>

#include "umma.cuh"

namespace goten {

namespace tc16 {

using namespace tc;

constexpr int BM = 128;          // UMMA M per CTA (cta_group::2: the pair computes M = 256)
constexpr int BK = 64;           // halves per k-block = one 128 B swizzle row
constexpr int MAX_STAGES = 4;
constexpr int NTHREADS = 320;
constexpr int EPI_WARP0 = 2, CONV_WARP0 = 6;
constexpr uint32_t A_BYTES = BM * BK * 4;   // raw fp32 tile = hi + lo fp16 tiles = 32 KB

struct Params {
  int M, N, K;
  int block_n;
  int n_mt, n_nt;
  int splits, kb_per_split, kb_total;
  int stages;
  float* C; int ldc;
  const float* bias;
  const float* add_src; int ld_add;
  float* act_out; int ld_act, act_lo, act_hi;
  int add_vec, c_vec, act_vec;
  int red_add;            // add_src == C: accumulate into C with a TMA reduction store
  int act_tma;            // SiLU side output through a second TMA store (tmAct)
  int a_tmem;             // K-major A: the converter writes the hi / lo tiles into TENSOR MEMORY (tcgen05.st) and the MMAs take
                          // A from there: no converted-A write to and no A read from shared memory (the short-K GEMMs are
                          // shared-memory-pipe bound).  TMEM: columns [0, 256) one accumulator stage, [256 + 64 s, ..) A stage s
  int dbg;                // GOTEN_GEMM_DBG (timing experiments only, results are wrong: 4 skip the operand split, 8 skip the output stores)
  float* partial;
  float* colsum;
  float* partial_colsum;
  const float* amax_a;    // device scalars: (a bound of) max |A| and max |B|
  const float* amax_b;
};

// B_RAW (weight-gradient form only): B is staged as raw fp32 [64 k][BNH n] like A and converted in place by the
// converter warps, so no pre-split copy of the (activation-sized) B operand is written to / re-read from HBM.
// EPIW: epilogue warps (4, or 8 = two per TMEM lane quadrant for the epilogue-bound short-K shapes: 448 threads put four
// warps on two of the SM's register-file partitions, i.e. 128 registers per thread instead of 168)
// CONVW: converter warps (4, or 8 for the weight-gradient forms: two threads per tile column, each transposing half of
// the 64 k values - those kernels are bound by the converter's strided shared-memory reads, not by the epilogue)
template <bool A_ROWS_ARE_K, bool B_RAW, int NCTA, int EPIW = 4, int CONVW = 4>
__global__ void __launch_bounds__(64 + 32 * (EPIW + CONVW), 1)
gemm16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
              const __grid_constant__ CUtensorMap tmBl, const __grid_constant__ CUtensorMap tmC,
              const __grid_constant__ CUtensorMap tmAct, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int BN = p.block_n;
  const int BNH = BN / NCTA;
  const int STAGES = p.stages;
  const uint32_t B_BYTES = (uint32_t)BNH * BK * 2;   // one fp16 B tile (hi or lo), <= 32 KB
  const uint32_t STAGE_BYTES = A_BYTES + 2 * B_BYTES;
  uint8_t* epi_smem = smem + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + 4 * 2 * 4096);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 4);
  const uint32_t bar_full = smem_u32(bars), bar_conv = smem_u32(bars + MAX_STAGES), bar_empty = smem_u32(bars + 2 * MAX_STAGES);
  const uint32_t bar_tfull = smem_u32(bars + 3 * MAX_STAGES), bar_tempty = smem_u32(bars + 3 * MAX_STAGES + 2);

  constexpr int CONV_WARP0 = EPI_WARP0 + EPIW;   // (shadows the 320-thread constant)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = NCTA == 2 ? cluster_ctarank() : 0u;
  const int unit = NCTA == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_units = NCTA == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_conv + 8 * s, CONVW * NCTA);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, EPIW * NCTA);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (NCTA == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  tc_fence_before();
  __syncthreads();  // (the pair also meets below: compute-sanitizer racecheck only models the CTA barrier)
  if (NCTA == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_tiles = p.n_mt * p.n_nt;
  const int n_items = n_tiles * p.splits;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      uint32_t it = 0;
      for (int w = unit; w < n_items; w += n_units) {
        const int split = w / n_tiles, tile = w % n_tiles;
        const int m0 = (tile / p.n_nt) * (BM * NCTA) + (int)rank * BM, n0 = (tile % p.n_nt) * BN + (int)rank * BNH;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t sbh = sa + A_BYTES, sbl = sbh + B_BYTES;
          const uint32_t bar = bar_full + 8 * s;
          mbar_arrive_expect_tx(bar, A_BYTES + 2 * B_BYTES);
          if (!A_ROWS_ARE_K) {
            // A[M][K]: two 128B-swizzled boxes of [128 rows][32 floats]
            tma_load_2d(sa, &tmA, bar, kb * BK, m0);
            tma_load_2d(sa + A_BYTES / 2, &tmA, bar, kb * BK + 32, m0);
          } else {
            // A[R][M]: one un-swizzled box of [64 k rows][128 floats]
            tma_load_2d(sa, &tmA, bar, m0, kb * BK);
            // (measured: an L2 prefetch of this tile 4 / 8 / 16 k-blocks ahead, cp.async.bulk.prefetch.tensor, makes the
            //  edge weight gradient 2 / 4 / 7 % slower - the ring is not waiting on DRAM latency)
          }
          if (B_RAW) {
            tma_load_2d(sbh, &tmBh, bar, n0, kb * BK);   // tmBh = fp32 map over B[R][N]: box [64 k][BNH n]
          } else {
            tma_load_2d(sbh, &tmBh, bar, kb * BK, n0);
            tma_load_2d(sbl, &tmBl, bar, kb * BK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    if (lane == 0 && rank == 0) {
      // c = f32 (bit 4), a = b = f16 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * NCTA) >> 4) << 24);
      uint32_t it = 0, tile_it = 0;
      for (int w = unit; w < n_items; w += n_units, ++tile_it) {
        const int split = w / n_tiles;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const bool a_tm = !A_ROWS_ARE_K && p.a_tmem;
        const bool one_acc = a_tm && BN > 192;   // 2 x 256 accumulator columns + the A stages exceed the 512 TMEM columns
        const uint32_t acc = one_acc ? 0u : (tile_it & 1), aph = one_acc ? (tile_it & 1) : ((tile_it >> 1) & 1);
        mbar_wait(bar_tempty + 8 * acc, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(bar_full + 8 * s, ph);
          mbar_wait(bar_conv + 8 * s, ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t sal = sa + A_BYTES / 2, sbh = sa + A_BYTES, sbl = sbh + B_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t a_hi = make_desc(sa + kk * 32, 16, 1024, 2), a_lo = make_desc(sal + kk * 32, 16, 1024, 2);
            const uint64_t b_hi = make_desc(sbh + kk * 32, 16, 1024, 2), b_lo = make_desc(sbl + kk * 32, 16, 1024, 2);
            if (a_tm) {
              // 192-column tiles: two accumulator stages in columns [0, 384) and a TWO-deep A ring behind them
              const uint32_t ta_hi = tmem_base + (BN == 192 ? 384 + (it & 1) * 64 : 256 + s * 64) + kk * 8, ta_lo = ta_hi + 32;
              if (NCTA == 2) {
                umma_f16_ts_2cta(d_tmem, ta_lo, b_hi, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
                umma_f16_ts_2cta(d_tmem, ta_hi, b_lo, idesc, 1u);
                umma_f16_ts_2cta(d_tmem, ta_hi, b_hi, idesc, 1u);
              } else {
                umma_f16_ts(d_tmem, ta_lo, b_hi, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
                umma_f16_ts(d_tmem, ta_hi, b_lo, idesc, 1u);
                umma_f16_ts(d_tmem, ta_hi, b_hi, idesc, 1u);
              }
            } else if (NCTA == 2) {
              umma_f16_2cta(d_tmem, a_lo, b_hi, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
              umma_f16_2cta(d_tmem, a_hi, b_lo, idesc, 1u);
              umma_f16_2cta(d_tmem, a_hi, b_hi, idesc, 1u);
            } else {
              umma_f16(d_tmem, a_lo, b_hi, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
              umma_f16(d_tmem, a_hi, b_lo, idesc, 1u);
              umma_f16(d_tmem, a_hi, b_hi, idesc, 1u);
            }
          }
          if (NCTA == 2) umma_commit_2cta(bar_empty + 8 * s); else umma_commit(bar_empty + 8 * s);
        }
        if (NCTA == 2) umma_commit_2cta(bar_tfull + 8 * acc); else umma_commit(bar_tfull + 8 * acc);
      }
    }
  } else if (warp >= CONV_WARP0) {
    // =============================== converter ==================================
    const bool a_tm = !A_ROWS_ARE_K && p.a_tmem;
    // 0..127 = tile row (M index) this thread produces; with A in tensor memory the row must sit in the TMEM lane
    // quadrant the warp may access (warp id mod 4)
    static_assert(CONVW == 4 || (CONVW == 8 && A_ROWS_ARE_K), "eight converter warps: weight-gradient forms only");
    constexpr int CF = CONVW / 4, KH = BK / CF, NCH = 8 / CF;   // threads per column, k values and 16 B chunks per thread
    const int cidx = threadIdx.x - CONV_WARP0 * 32;
    const int ct = a_tm ? (warp & 3) * 32 + lane : cidx % 128;
    const int hf = cidx / 128;   // which half of the k range (0 with four warps)
    __shared__ float cs_x[CONVW == 8 ? 128 : 1];   // column-sum hand-over between the two halves
    const float sA = scale_of(*p.amax_a);
    const float sB = scale_of(*p.amax_b);
    const uint32_t sw = (uint32_t)(ct & 7);
    uint32_t it = 0;
    for (int w = unit; w < n_items; w += n_units) {
      const int split = w / n_tiles, tile = w % n_tiles;
      const int m0 = (tile / p.n_nt) * (BM * NCTA) + (int)rank * BM;
      const bool do_cs = A_ROWS_ARE_K && (p.colsum != nullptr) && (tile % p.n_nt == 0);
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
      float csum = 0.f;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(bar_full + 8 * s, ph);
        uint8_t* a_raw = smem + s * STAGE_BYTES;
        // (explicit shared-space accesses: through the generic pointers the compiler emitted LD.E / ST.E)
        const uint32_t a_raw_s = smem_u32(a_raw);
        const uint32_t hi_row = a_raw_s + ct * 128, lo_row = hi_row + A_BYTES / 2;
        if (!A_ROWS_ARE_K && (p.dbg & 4)) {
        } else if (!A_ROWS_ARE_K) {
          // thread = tile row: its 64 floats live in row ct of the two raw boxes and are replaced by row ct of the
          // hi tile (first box) and of the lo tile (second box).  All loads are issued before the first store.
          float4 v[16];
#pragma unroll
          for (int c = 0; c < 16; ++c)
            v[c] = lds128(a_raw_s + (c >> 3) * (A_BYTES / 2) + ct * 128 + ((((uint32_t)c & 7) ^ sw) << 4));
          if (a_tm) {
            uint32_t hw[32], lw[32];   // packed half pairs: word j = K elements 2j, 2j + 1 of this row
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              split2(v[c].x * sA, v[c].y * sA, hw[2 * c], lw[2 * c]);
              split2(v[c].z * sA, v[c].w * sA, hw[2 * c + 1], lw[2 * c + 1]);
            }
            uint32_t ta = tmem_base + 256 + s * 64 + ((uint32_t)((warp & 3) * 32) << 16);
            if (BN == 192) {
              // two-deep A ring: slot it & 1 is free once the MMAs of k-block it - 2 have completed, which is what the
              // `empty` barrier of the shared-memory stage that k-block used reports
              ta = tmem_base + 384 + (it & 1) * 64 + ((uint32_t)((warp & 3) * 32) << 16);
              if (it >= 2) {
                const uint32_t it2 = it - 2;
                mbar_wait(bar_empty + 8 * (it2 % STAGES), (it2 / STAGES) & 1);
                tc_fence_after();
              }
            }
            GOTEN_STTM_X32(ta, hw);
            GOTEN_STTM_X32(ta + 32, lw);
            tmem_wait_st();
            tc_fence_before();
          } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) {  // 16 B chunk of 8 halves = floats 8c .. 8c+7
            uint4 h, l;
            split2(v[2 * c].x * sA, v[2 * c].y * sA, h.x, l.x);
            split2(v[2 * c].z * sA, v[2 * c].w * sA, h.y, l.y);
            split2(v[2 * c + 1].x * sA, v[2 * c + 1].y * sA, h.z, l.z);
            split2(v[2 * c + 1].z * sA, v[2 * c + 1].w * sA, h.w, l.w);
            const uint32_t off = (((uint32_t)c ^ sw) << 4);
            sts128u(hi_row + off, h);
            sts128u(lo_row + off, l);
          }
          }
        } else if (p.dbg & 4) {
        } else {
          // thread = MN column ct of the raw [64 k][128 mn] tile: read the column, then (after every converter
          // thread has read) write it as K-major row ct of the hi / lo tiles
          float v[KH];
#pragma unroll
          for (int k = 0; k < KH; ++k) v[k] = lds32(a_raw_s + (hf * KH + k) * 512 + ct * 4);
          conv_bar_sync<CONVW * 32>();
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            uint4 h, l;
            split2(v[8 * c + 0] * sA, v[8 * c + 1] * sA, h.x, l.x);
            split2(v[8 * c + 2] * sA, v[8 * c + 3] * sA, h.y, l.y);
            split2(v[8 * c + 4] * sA, v[8 * c + 5] * sA, h.z, l.z);
            split2(v[8 * c + 6] * sA, v[8 * c + 7] * sA, h.w, l.w);
            const uint32_t off = (((uint32_t)(hf * NCH + c) ^ sw) << 4);
            sts128u(hi_row + off, h);
            sts128u(lo_row + off, l);
          }
          if (do_cs) {
#pragma unroll
            for (int k = 0; k < KH; ++k) csum += v[k];
          }
          if (B_RAW) {
            // same transposing split for column ct of the raw B tile (BNH <= 128 columns, row pitch BNH floats)
            const uint32_t b_raw = a_raw_s + A_BYTES;
            const bool mine = ct < BNH;
            if (mine) {
#pragma unroll
              for (int k = 0; k < KH; ++k) v[k] = lds32(b_raw + (uint32_t)(hf * KH + k) * BNH * 4 + ct * 4);
            }
            conv_bar_sync<CONVW * 32>();
            if (mine) {
              const uint32_t bh_row = b_raw + ct * 128, bl_row = bh_row + B_BYTES;
#pragma unroll
              for (int c = 0; c < NCH; ++c) {
                uint4 h, l;
                split2(v[8 * c + 0] * sB, v[8 * c + 1] * sB, h.x, l.x);
                split2(v[8 * c + 2] * sB, v[8 * c + 3] * sB, h.y, l.y);
                split2(v[8 * c + 4] * sB, v[8 * c + 5] * sB, h.z, l.z);
                split2(v[8 * c + 6] * sB, v[8 * c + 7] * sB, h.w, l.w);
                const uint32_t off = (((uint32_t)(hf * NCH + c) ^ sw) << 4);
                sts128u(bh_row + off, h);
                sts128u(bl_row + off, l);
              }
            }
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (NCTA == 2) mbar_arrive_cluster(bar_conv + 8 * s, 0);
          else mbar_arrive(bar_conv + 8 * s);
        }
      }
      if (do_cs) {
        if (CONVW == 8) {   // the two halves of the k range meet (uniform branch: do_cs depends on the tile only)
          if (hf == 1) cs_x[ct] = csum;
          conv_bar_sync<CONVW * 32>();
          if (hf == 0) csum += cs_x[ct];
          conv_bar_sync<CONVW * 32>();
        }
        const int m = m0 + ct;
        if (m < p.M && hf == 0) {
          if (p.partial_colsum) p.partial_colsum[(size_t)split * p.M + m] = csum;
          else p.colsum[m] = csum;
        }
      }
    }
  } else {
    // =============================== epilogue (umma.cuh) ========================
    gemm_epilogue<NCTA, BM, EPI_WARP0, EPIW>(p, tmC, tmAct, epi_smem, bar_tfull, bar_tempty, tmem_base, warp, lane, unit, n_units, n_items,
                                       n_tiles, BN, rank, inv_scale_of(*p.amax_a), inv_scale_of(*p.amax_b),
                                       (!A_ROWS_ARE_K && p.a_tmem && BN > 192) ? 1 : 2);
  }

  tc_fence_before();
  __syncthreads();  // (the pair also meets below: compute-sanitizer racecheck only models the CTA barrier)
  if (NCTA == 2) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if (NCTA == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  }
}

// ---------------------------------------------------------------- operand preparation
// max |x| over a [rows][cols] matrix with leading dimension ld -> atomicMax on the (non-negative) float bits.
// Vector form: four independent 128-bit loads per thread and iteration (a pure read stream: memory-level parallelism
// is all that matters), one atomic per block.
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ in, int64_t ld, int64_t rows, int cols,
                                                     int vec, float* __restrict__ out) {
  __shared__ float red[8];
  float m = 0.f;
  if (vec) {
    const int c4 = cols >> 2;
    const int64_t total = rows * c4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ld == cols) {  // contiguous: no row arithmetic
      const float4* p = reinterpret_cast<const float4*>(in);
      for (; idx + 3 * stride < total; idx += 4 * stride) {
        const float4 a = p[idx], b = p[idx + stride], c = p[idx + 2 * stride], d = p[idx + 3 * stride];
        m = amax4(amax4(amax4(amax4(m, a.x, a.y, a.z, a.w), b.x, b.y, b.z, b.w), c.x, c.y, c.z, c.w), d.x, d.y, d.z, d.w);
      }
      for (; idx < total; idx += stride) {
        const float4 a = p[idx];
        m = amax4(m, a.x, a.y, a.z, a.w);
      }
    } else {
      for (; idx < total; idx += stride) {
        const int64_t r = idx / c4;
        const int c = (int)(idx - r * c4);
        const float4 v = *reinterpret_cast<const float4*>(in + r * ld + 4 * c);
        m = amax4(m, v.x, v.y, v.z, v.w);
      }
    }
  } else {
    const int64_t total = rows * cols;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = idx / cols;
      m = fmaxf(m, fabsf(in[r * ld + (idx - r * cols)]));
    }
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    m = warp_max(m);
    if (threadIdx.x == 0 && m > 0.f) atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(m));
  }
}

// B[N][K] (ld) -> fp16 hi / lo [N][Kp]: thread = 4 consecutive k
__global__ void split16_rows_kernel(const float* __restrict__ in, int ld, int rows, int cols, int Kp, int vec,
                                    const float* __restrict__ amax, __half* __restrict__ hi, __half* __restrict__ lo) {
  const float s = scale_of(*amax);
  const int c4 = (cols + 3) >> 2;
  const int64_t total = (int64_t)rows * c4;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(idx / c4), c = 4 * (int)(idx - (int64_t)r * c4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* ip = in + (int64_t)r * ld + c;
    if (vec && c + 3 < cols) v = *reinterpret_cast<const float4*>(ip);
    else {
      if (c + 0 < cols) v.x = ip[0];
      if (c + 1 < cols) v.y = ip[1];
      if (c + 2 < cols) v.z = ip[2];
      if (c + 3 < cols) v.w = ip[3];
    }
    uint2 h, l;
    split2(v.x * s, v.y * s, h.x, l.x);
    split2(v.z * s, v.w * s, h.y, l.y);
    *reinterpret_cast<uint2*>(hi + (int64_t)r * Kp + c) = h;   // Kp % 8 == 0 and c % 4 == 0: 8 B aligned, inside the row
    *reinterpret_cast<uint2*>(lo + (int64_t)r * Kp + c) = l;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Panel form for the short-K forward GEMMs (K <= 256, K-major A, weights pre-split, CTA pairs): A-STATIONARY.
// The converted A operand of a whole M block (256 rows x K <= 256: 128 lanes x 256 TMEM columns per CTA, hi + lo) is
// written ONCE into tensor memory and reused by every 128-column tile of the block, so per tile only the weight tiles
// cross shared memory: the default kernel re-stages and re-splits A for each column tile and is bound by the SM's
// shared-memory pipe.  TMEM: columns [0, 256) two 128-column accumulator stages, [256, 512) the panel.
// Shared memory: 4 raw A stages (one per k-block of the NEXT block, loaded while the current one is computed), a
// 4-deep ring of weight stages (hi + lo, 16 KB), the epilogue staging blocks.
//   warp 0   TMA producer (weights of the current block; raw A of the next block after the first tile's weights)
//   warp 1   leader CTA: MMA issuer ([d_tmem], [a_tmem], b_desc form); peer CTA: relays "my weight stage has landed"
//   warps 2-5 epilogue (umma.cuh, panel order)     warps 6-9 converter: raw A -> hi / lo -> tcgen05.st into the panel
constexpr int P_AR = 4, P_BS = 4, P_BN = 128, P_BNH = 64;
constexpr uint32_t P_B_BYTES = (uint32_t)P_BNH * BK * 2;   // one fp16 weight tile (hi or lo) of this CTA: 8 KB
constexpr uint32_t P_BSTAGE = 2 * P_B_BYTES;
constexpr int P_NBARS = 4 * P_AR + 3 * P_BS + 1 + 4;       // a_full, a_empty | b_full, b_empty, b_peer | pready[4] .. see below
constexpr size_t P_SMEM = 1024 + (size_t)P_AR * A_BYTES + (size_t)P_BS * P_BSTAGE + 4 * 2 * 4096 + 40 * 8 + 16;

__global__ void __launch_bounds__(NTHREADS, 1)
gemm16_panel_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
                    const __grid_constant__ CUtensorMap tmBl, const __grid_constant__ CUtensorMap tmC,
                    const __grid_constant__ CUtensorMap tmAct, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + (size_t)P_AR * A_BYTES;
  uint8_t* epi_smem = b_smem + (size_t)P_BS * P_BSTAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + 4 * 2 * 4096);
  const uint32_t bar_afull = smem_u32(bars), bar_aempty = bar_afull + 8 * P_AR;
  const uint32_t bar_bfull = bar_aempty + 8 * P_AR, bar_bempty = bar_bfull + 8 * P_BS, bar_bpeer = bar_bempty + 8 * P_BS;
  const uint32_t bar_pready = bar_bpeer + 8 * P_BS, bar_pfree = bar_pready + 8 * 4;
  const uint32_t bar_tfull = bar_pfree + 8, bar_tempty = bar_tfull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 36);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int unit = (int)(blockIdx.x >> 1), n_units = (int)(gridDim.x >> 1);
  const int KB = p.kb_total;   // <= 4

  if (threadIdx.x == 0) {
    for (int s = 0; s < P_AR; ++s) { mbar_init(bar_afull + 8 * s, 1); mbar_init(bar_aempty + 8 * s, 4); }
    for (int s = 0; s < P_BS; ++s) { mbar_init(bar_bfull + 8 * s, 1); mbar_init(bar_bempty + 8 * s, 1); mbar_init(bar_bpeer + 8 * s, 1); }
    for (int k = 0; k < 4; ++k) mbar_init(bar_pready + 8 * k, 8);   // four converter warps of both CTAs
    mbar_init(bar_pfree, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 8); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      uint32_t ait = 0, bit = 0;
      auto load_a = [&](int mb) {
        const int m0 = mb * (BM * 2) + (int)rank * BM;
        for (int kb = 0; kb < KB; ++kb, ++ait) {
          const uint32_t s = ait % P_AR, ph = (ait / P_AR) & 1;
          mbar_wait(bar_aempty + 8 * s, ph ^ 1);
          const uint32_t sa = smem_u32(a_smem + (size_t)s * A_BYTES), bar = bar_afull + 8 * s;
          mbar_arrive_expect_tx(bar, A_BYTES);
          tma_load_2d(sa, &tmA, bar, kb * BK, m0);
          tma_load_2d(sa + A_BYTES / 2, &tmA, bar, kb * BK + 32, m0);
        }
      };
      if (unit < p.n_mt) load_a(unit);
      for (int mb = unit; mb < p.n_mt; mb += n_units) {
        for (int nt = 0; nt < p.n_nt; ++nt) {
          const int n0 = nt * P_BN + (int)rank * P_BNH;
          for (int kb = 0; kb < KB; ++kb, ++bit) {
            const uint32_t s = bit % P_BS, ph = (bit / P_BS) & 1;
            mbar_wait(bar_bempty + 8 * s, ph ^ 1);
            const uint32_t sb = smem_u32(b_smem + (size_t)s * P_BSTAGE), bar = bar_bfull + 8 * s;
            mbar_arrive_expect_tx(bar, P_BSTAGE);
            tma_load_2d(sb, &tmBh, bar, kb * BK, n0);
            tma_load_2d(sb + P_B_BYTES, &tmBl, bar, kb * BK, n0);
          }
          // the raw A tiles of the NEXT block: their stages were released while this block's panel was written
          if (nt == 0 && mb + n_units < p.n_mt) load_a(mb + n_units);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // =============================== MMA issuer =================================
      const uint32_t idesc = (1u << 4) | ((uint32_t)(P_BN >> 3) << 17) | ((uint32_t)((BM * 2) >> 4) << 24);
      uint32_t bit = 0, tile_it = 0, mbi = 0;
      for (int mb = unit; mb < p.n_mt; mb += n_units, ++mbi) {
        for (int nt = 0; nt < p.n_nt; ++nt, ++tile_it) {
          const uint32_t acc = tile_it & 1, aph = (tile_it >> 1) & 1;
          mbar_wait(bar_tempty + 8 * acc, aph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * (uint32_t)P_BN;
          for (int kb = 0; kb < KB; ++kb, ++bit) {
            if (nt == 0) mbar_wait(bar_pready + 8 * kb, mbi & 1);
            const uint32_t s = bit % P_BS, ph = (bit / P_BS) & 1;
            mbar_wait(bar_bfull + 8 * s, ph);
            mbar_wait(bar_bpeer + 8 * s, ph);
            tc_fence_after();
            const uint32_t sbh = smem_u32(b_smem + (size_t)s * P_BSTAGE), sbl = sbh + P_B_BYTES;
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
              const uint64_t b_hi = make_desc(sbh + kk * 32, 16, 1024, 2), b_lo = make_desc(sbl + kk * 32, 16, 1024, 2);
              const uint32_t ta_hi = tmem_base + 256 + kb * 64 + kk * 8, ta_lo = ta_hi + 32;
              umma_f16_ts_2cta(d_tmem, ta_lo, b_hi, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
              umma_f16_ts_2cta(d_tmem, ta_hi, b_lo, idesc, 1u);
              umma_f16_ts_2cta(d_tmem, ta_hi, b_hi, idesc, 1u);
            }
            umma_commit_2cta(bar_bempty + 8 * s);
          }
          umma_commit_2cta(bar_tfull + 8 * acc);
        }
        umma_commit_2cta(bar_pfree);   // every MMA that reads this block's panel has completed
      }
    } else if (lane == 0) {
      // peer CTA: the leader's `full` barrier counts only the leader's own TMA bytes - report each landed stage
      uint32_t bit = 0;
      for (int mb = unit; mb < p.n_mt; mb += n_units)
        for (int i = 0; i < p.n_nt * KB; ++i, ++bit) {
          const uint32_t s = bit % P_BS, ph = (bit / P_BS) & 1;
          mbar_wait(bar_bfull + 8 * s, ph);
          mbar_arrive_cluster(bar_bpeer + 8 * s, 0);
        }
    }
  } else if (warp >= CONV_WARP0) {
    // =============================== converter ==================================
    const int ct = (warp & 3) * 32 + lane;   // tile row = TMEM lane (the warp's lane quadrant)
    const uint32_t sw = (uint32_t)(ct & 7);
    const float sA = scale_of(*p.amax_a);
    uint32_t ait = 0, mbi = 0;
    for (int mb = unit; mb < p.n_mt; mb += n_units, ++mbi) {
      if (mbi > 0) {   // the panel is free once the previous block's last MMA has completed
        mbar_wait(bar_pfree, (mbi - 1) & 1);
        tc_fence_after();
      }
      for (int kb = 0; kb < KB; ++kb, ++ait) {
        const uint32_t s = ait % P_AR, ph = (ait / P_AR) & 1;
        mbar_wait(bar_afull + 8 * s, ph);
        const uint32_t a_raw = smem_u32(a_smem + (size_t)s * A_BYTES);
        float4 v[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = lds128(a_raw + (c >> 3) * (A_BYTES / 2) + ct * 128 + ((((uint32_t)c & 7) ^ sw) << 4));
        uint32_t hw[32], lw[32];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          split2(v[c].x * sA, v[c].y * sA, hw[2 * c], lw[2 * c]);
          split2(v[c].z * sA, v[c].w * sA, hw[2 * c + 1], lw[2 * c + 1]);
        }
        const uint32_t ta = tmem_base + 256 + kb * 64 + ((uint32_t)((warp & 3) * 32) << 16);
        GOTEN_STTM_X32(ta, hw);
        GOTEN_STTM_X32(ta + 32, lw);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar_aempty + 8 * s);               // raw stage consumed (this CTA's producer)
          mbar_arrive_cluster(bar_pready + 8 * kb, 0);   // panel k-block written (leader's MMA issuer)
        }
      }
    }
  } else {
    // =============================== epilogue (umma.cuh) ========================
    gemm_epilogue<2, BM, EPI_WARP0>(p, tmC, tmAct, epi_smem, bar_tfull, bar_tempty, tmem_base, warp, lane, unit, n_units, 0, 0,
                                    P_BN, rank, inv_scale_of(*p.amax_a), inv_scale_of(*p.amax_b), 2, true);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  }
}

// B[K][N] (ld) -> fp16 hi / lo [N][Kp] through a 64 x 64 shared-memory tile (coalesced on both sides)
__global__ void __launch_bounds__(256)
split16_transpose_kernel(const float* __restrict__ in, int ld, int rows /*K*/, int cols /*N*/, int Kp,
                         const float* __restrict__ amax, __half* __restrict__ hi, __half* __restrict__ lo) {
  __shared__ float tile[64][65];
  const float s = scale_of(*amax);
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
  for (int y = ty; y < 64; y += 4) {
    const int r = r0 + y, c = c0 + tx;
    tile[y][tx] = (r < rows && c < cols) ? in[(int64_t)r * ld + c] * s : 0.f;
  }
  __syncthreads();
  // thread -> (output row c = c0 + oc, 8 consecutive k = r0 + 8 j ..)
  const int j = threadIdx.x & 7;
  for (int oc = threadIdx.x >> 3; oc < 64; oc += 32) {
    const int c = c0 + oc, r = r0 + 8 * j;
    if (c >= cols || r >= Kp) continue;
    uint4 h, l;
    split2(tile[8 * j + 0][oc], tile[8 * j + 1][oc], h.x, l.x);
    split2(tile[8 * j + 2][oc], tile[8 * j + 3][oc], h.y, l.y);
    split2(tile[8 * j + 4][oc], tile[8 * j + 5][oc], h.z, l.z);
    split2(tile[8 * j + 6][oc], tile[8 * j + 7][oc], h.w, l.w);
    *reinterpret_cast<uint4*>(hi + (int64_t)c * Kp + r) = h;   // r % 8 == 0, Kp % 8 == 0: 16 B aligned, inside the row
    *reinterpret_cast<uint4*>(lo + (int64_t)c * Kp + r) = l;
  }
}

// Same transposition without shared memory, for activation-sized B: thread = one output row c and 16 consecutive k;
// a warp reads 16 coalesced 128 B row segments and every lane writes whole 32 B sectors of its hi / lo rows.
__global__ void __launch_bounds__(256)
split16_transpose_direct_kernel(const float* __restrict__ in, int ld, int rows /*K*/, int cols /*N*/, int Kp,
                                const float* __restrict__ amax, __half* __restrict__ hi, __half* __restrict__ lo) {
  const float s = scale_of(*amax);
  const int c = blockIdx.y * 256 + threadIdx.x;
  const int r0 = blockIdx.x * 16;
  if (c >= cols) return;
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = (r0 + i < rows) ? in[(int64_t)(r0 + i) * ld + c] * s : 0.f;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    if (r0 + 8 * g >= Kp) break;
    uint4 h, l;
    split2(v[8 * g + 0], v[8 * g + 1], h.x, l.x);
    split2(v[8 * g + 2], v[8 * g + 3], h.y, l.y);
    split2(v[8 * g + 4], v[8 * g + 5], h.z, l.z);
    split2(v[8 * g + 6], v[8 * g + 7], h.w, l.w);
    *reinterpret_cast<uint4*>(hi + (int64_t)c * Kp + r0 + 8 * g) = h;
    *reinterpret_cast<uint4*>(lo + (int64_t)c * Kp + r0 + 8 * g) = l;
  }
}

// up to 16 contiguous tensors in one launch: blockIdx.y = tensor, grid-stride over its elements
struct AbsmaxMulti {
  const float* ptr[16];
  int64_t n[16];
};
__global__ void __launch_bounds__(256) absmax_multi_kernel(AbsmaxMulti a, float* __restrict__ out) {
  __shared__ float red[8];
  const float* __restrict__ p = a.ptr[blockIdx.y];
  const int64_t n = a.n[blockIdx.y];
  float m = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    const float4* p4 = reinterpret_cast<const float4*>(p);
    for (int64_t i = idx; i < (n >> 2); i += stride) {
      const float4 v = p4[i];
      m = amax4(m, v.x, v.y, v.z, v.w);
    }
    for (int64_t i = ((n >> 2) << 2) + idx; i < n; i += stride) m = fmaxf(m, fabsf(p[i]));
  } else {
    for (int64_t i = idx; i < n; i += stride) m = fmaxf(m, fabsf(p[i]));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    m = warp_max(m);
    if (threadIdx.x == 0 && m > 0.f) atomicMax(reinterpret_cast<unsigned int*>(out + blockIdx.y), __float_as_uint(m));
  }
}

}  // namespace tc16

int splitk_finish(const float* partial, const float* partial_cs, int splits, float* C, int ldc, int M, int N,
                  const float* bias, const float* add_src, int ld_add, float* act_out, int ld_act, int act_lo, int act_hi,
                  float* colsum, cudaStream_t st);

// 2-D fp32 map over P[rows][cols] (ld floats)
static bool make_map_f32(CUtensorMap* m, const float* P, int64_t ld, int64_t rows, int64_t cols, int box_cols, int box_rows,
                         CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(P), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// K-major fp16 operand P[rows][k] (ld halves): box = 64 k x box_rows rows, 128B swizzle
static bool make_map_f16(CUtensorMap* m, const __half* P, int64_t ld, int64_t rows, int64_t k, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(P), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct Tc16Plan {
  bool ok;
  bool a_rows_are_k;     // weight-gradient form: A given as [R][M]
  int block_n, n_mt, n_nt, splits, kb_total, kb_per_split;
  int ncta, stages;
  bool panel;            // A-stationary panel kernel (short K, several column tiles)
  bool b_raw;            // weight-gradient form with B staged raw and converted in the kernel
  int64_t kp;            // padded K of the pre-split B copies (multiple of 8 halves = 16 B)
  int64_t b_bytes;       // bytes of each pre-split B copy
  int64_t ws_bytes;
};

static Tc16Plan tc16_plan(int M, int N, int K, int trans_a, int trans_b) {
  Tc16Plan t{};
  t.ok = false;
  if (M <= 0 || N <= 0 || K <= 0) return t;
  if (trans_a && trans_b) return t;
  if (N < 16) return t;
  t.a_rows_are_k = trans_a != 0;
  if (t.a_rows_are_k && M % 32 != 0) return t;  // split-K partial stores are whole 32-row blocks
  t.block_n = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  {
    static int force_bn = -1;   // GOTEN_GEMM_BN=128: narrower tiles (timing experiments)
    if (force_bn < 0) { const char* e = getenv("GOTEN_GEMM_BN"); force_bn = e ? atoi(e) : 0; }
    if (force_bn == 128 && t.block_n == 256) t.block_n = 128;
    if (force_bn == 192 && t.block_n == 256 && N >= 512) t.block_n = 192;
  }
  static int force_ncta = -1;
  if (force_ncta < 0) { const char* e = getenv("GOTEN_GEMM_NCTA"); force_ncta = e ? atoi(e) : 0; }
  t.ncta = (force_ncta == 1) ? 1 : ((M >= 256 || force_ncta == 2) ? 2 : 1);
  const size_t stage_bytes = tc16::A_BYTES + 2 * (size_t)(t.block_n / t.ncta) * tc16::BK * 2;
  t.stages = (int)((232448 - 4 * 2 * 4096 - 1024 - 256) / stage_bytes);
  if (t.stages > tc16::MAX_STAGES) t.stages = tc16::MAX_STAGES;
  t.n_mt = (M + tc16::BM * t.ncta - 1) / (tc16::BM * t.ncta);
  t.n_nt = (N + t.block_n - 1) / t.block_n;
  t.kb_total = (K + tc16::BK - 1) / tc16::BK;
  t.splits = 1;
  if (t.a_rows_are_k) {
    // chains of accumulating MMAs are kept short (see gemm_tc.cu: the tensor core accumulates with truncation);
    // the fp32 partials are summed with round-to-nearest by the split-K reduction kernel
    constexpr int MAX_CHAIN_KB = 48;
    const int tiles = t.n_mt * t.n_nt;
    int want = (148 / t.ncta) / tiles;
    if (want < 1) want = 1;
    int per = (t.kb_total + want - 1) / want;
    if (per > MAX_CHAIN_KB) per = MAX_CHAIN_KB;
    if (per < 4) per = t.kb_total < 4 ? t.kb_total : 4;
    t.splits = (t.kb_total + per - 1) / per;
  }
  t.kb_per_split = (t.kb_total + t.splits - 1) / t.splits;
  t.splits = (t.kb_total + t.kb_per_split - 1) / t.kb_per_split;
  {
    // A-stationary panel kernel: K-major A with K <= 256 (the converted A of an M block fills the 256 spare TMEM
    // columns), at least two 128-column tiles to reuse it, CTA pairs.  GOTEN_GEMM_PANEL=0|1 forces it off / on.
    static int panel = -1;
    if (panel < 0) { const char* e = getenv("GOTEN_GEMM_PANEL"); panel = e ? atoi(e) : 0; }
    t.panel = panel && !t.a_rows_are_k && t.kb_total <= 4 && N >= 256 && M >= 512 && t.ncta == 2;
    if (t.panel) {
      t.block_n = tc16::P_BN;
      t.n_nt = (N + t.block_n - 1) / t.block_n;
      t.splits = 1;
      t.kb_per_split = t.kb_total;
    }
  }
  t.kp = ((int64_t)K + 7) & ~int64_t(7);
  // Weight-gradient form with a single row of M tiles: every B tile is consumed exactly once, so the kernel converts
  // it in place (no pre-split copy through HBM).  With several M tiles the kernel is shared-memory-bandwidth bound and
  // re-converting B per M tile costs more than the one pre-split pass (measured: 1792 x 256 x 301k, 0.94 -> 1.00 ms).
  // Needs one tile column per converter thread (BNH <= 128) and a TMA-addressable B (checked at launch).
  // GOTEN_GEMM_BRAW_MT: largest number of M tiles for which B is converted in the kernel.  Was 2 with four converter
  // warps (1792 x 256 x 301k: 0.94 -> 1.00 ms with B re-converted per M tile); with eight converter warps the in-kernel
  // form wins on every weight-gradient shape of the model (997 -> 850, 854 -> 710, 58.9 -> 52.9, 56.7 -> 51.1 us) and
  // the pre-split pass over the activation-sized B operand (0.1 ms and 0.6 GB of traffic per edge weight gradient) goes
  static int braw_mt = -1;
  if (braw_mt < 0) { const char* e = getenv("GOTEN_GEMM_BRAW_MT"); braw_mt = e ? atoi(e) : 16; }
  t.b_raw = t.a_rows_are_k && t.n_mt <= braw_mt && t.block_n / t.ncta <= 128;
  t.b_bytes = align256((int64_t)N * t.kp * 2);
  t.ws_bytes = 256 + 2 * t.b_bytes;
  if (t.splits > 1) t.ws_bytes += align256((int64_t)t.splits * ((int64_t)M * N + M) * 4);
  t.ok = true;
  return t;
}

int64_t gemm_tc16_workspace_bytes(int M, int N, int K, int trans_a, int trans_b) {
  Tc16Plan t = tc16_plan(M, N, K, trans_a, trans_b);
  return t.ok ? t.ws_bytes : 0;
}

static int launch_absmax(const float* P, int64_t ld, int64_t rows, int cols, float* slot, int sm_count, cudaStream_t st) {
  const int vec = (cols % 4 == 0 && ld % 4 == 0 && aligned16(P)) ? 1 : 0;
  const int64_t work = vec ? rows * (cols / 4) : rows * cols;
  int64_t grid = cdiv64(work, 256 * 4);
  if (grid > (int64_t)sm_count * 16) grid = (int64_t)sm_count * 16;
  if (grid < 1) grid = 1;
  tc16::absmax_kernel<<<(unsigned)grid, 256, 0, st>>>(P, ld, rows, cols, vec, slot);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

// a_amax / b_amax: optional device pointers to a known upper bound of max|A| / max|B| (any bound within a few
// powers of two of the true maximum keeps full accuracy); nullptr = computed here with one read pass.
int gemm_tc16(const float* A, int lda, int trans_a, const float* B, int ldb, int trans_b, float* C, int ldc, int M, int N,
              int K, const float* bias, const float* add_src, int ld_add, float* act_out, int ld_act, int act_lo,
              int act_hi, float* colsum, const float* a_amax, const float* b_amax, void* workspace,
              int64_t workspace_bytes, cudaStream_t st, bool* handled) {
  *handled = false;
  Tc16Plan t = tc16_plan(M, N, K, trans_a, trans_b);
  if (!t.ok || workspace == nullptr || workspace_bytes < t.ws_bytes) return 0;
  if (!aligned16(A) || lda % 4 != 0) return 0;
  if (t.b_raw && (!aligned16(B) || ldb % 4 != 0)) t.b_raw = false;
  if (colsum && !t.a_rows_are_k) return 0;
  if (get_encode() == nullptr) return 0;
  static int sm_count = 0, smem_optin = 0;
  if (sm_count == 0) {
    int dev = 0;
    GOTEN_CHECK_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    GOTEN_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return 0;
    sm_count = prop.multiProcessorCount;
    smem_optin = (int)prop.sharedMemPerBlockOptin;
  }

  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  float* amax = reinterpret_cast<float*>(ws);          // [0] A, [1] B
  __half* Bh = reinterpret_cast<__half*>(ws + 256);
  __half* Bl = reinterpret_cast<__half*>(ws + 256 + t.b_bytes);
  float* partial = nullptr;
  float* partial_cs = nullptr;
  if (t.splits > 1) {
    partial = reinterpret_cast<float*>(ws + 256 + 2 * t.b_bytes);
    partial_cs = partial + (size_t)t.splits * M * N;
  }

  // ---- scales: caller-supplied bounds are read in place; missing ones are measured into the workspace slots
  const float* pa = a_amax;
  const float* pb = b_amax;
  if (!pa || !pb) {
    GOTEN_CHECK_CUDA(cudaMemsetAsync(amax, 0, 8, st));
    if (!pa) {
      if (launch_absmax(A, lda, t.a_rows_are_k ? K : M, t.a_rows_are_k ? M : K, amax, sm_count, st)) return 1;
      pa = amax;
    }
    if (!pb) {
      const bool b_is_nk = !t.a_rows_are_k && trans_b;   // B[N][K]; otherwise B[K][N]
      if (launch_absmax(B, ldb, b_is_nk ? N : K, b_is_nk ? K : N, amax + 1, sm_count, st)) return 1;
      pb = amax + 1;
    }
  }

  // ---- B operand: scaled fp16 hi / lo copies, K-major [N][Kp] (weights); the weight-gradient form converts B in the kernel
  if (t.b_raw) {
  } else if (!t.a_rows_are_k && trans_b) {
    const int vec = (ldb % 4 == 0 && aligned16(B)) ? 1 : 0;
    const int64_t work = (int64_t)N * ((K + 3) / 4);
    int64_t grid = cdiv64(work, 256);
    if (grid > (int64_t)sm_count * 16) grid = (int64_t)sm_count * 16;
    tc16::split16_rows_kernel<<<(unsigned)grid, 256, 0, st>>>(B, ldb, N, K, (int)t.kp, vec, pb, Bh, Bl);
  } else {
    if (K >= 4096 && N >= 64) {  // long reduction (weight gradient): shared-memory-free form
      dim3 grid((unsigned)((K + 15) / 16), (unsigned)((N + 255) / 256));
      tc16::split16_transpose_direct_kernel<<<grid, 256, 0, st>>>(B, ldb, K, N, (int)t.kp, pb, Bh, Bl);
    } else {
      dim3 grid((unsigned)((N + 63) / 64), (unsigned)((K + 63) / 64));
      tc16::split16_transpose_kernel<<<grid, 256, 0, st>>>(B, ldb, K, N, (int)t.kp, pb, Bh, Bl);
    }
  }
  if (!t.b_raw) GOTEN_CHECK_LAUNCH();

  CUtensorMap mA, mBh, mBl, mC, mAct;
  bool ok;
  if (!t.a_rows_are_k) ok = make_map_f32(&mA, A, lda, M, K, 32, tc16::BM, CU_TENSOR_MAP_SWIZZLE_128B);
  else ok = make_map_f32(&mA, A, lda, K, M, tc16::BM, tc16::BK, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (t.b_raw) {
    ok = ok && make_map_f32(&mBh, B, ldb, K, N, t.block_n / t.ncta, tc16::BK, CU_TENSOR_MAP_SWIZZLE_NONE);
    mBl = mBh;
  } else {
    ok = ok && make_map_f16(&mBh, Bh, t.kp, N, K, t.block_n / t.ncta) && make_map_f16(&mBl, Bl, t.kp, N, K, t.block_n / t.ncta);
  }
  GOTEN_REQUIRE(ok, "cuTensorMapEncodeTiled failed (fp16 GEMM M=%d N=%d K=%d lda=%d)", M, N, K, lda);
  const bool red_add = add_src != nullptr && add_src == C && ld_add == ldc && act_out == nullptr && t.splits == 1;
  const bool act_tma = act_out != nullptr && t.splits == 1 && act_lo % 32 == 0 && act_hi > act_lo && aligned16(act_out) &&
                       ld_act % 4 == 0 && (add_src == nullptr || red_add);
  const bool fast_epi = t.splits > 1 || ((add_src == nullptr || red_add) && (act_out == nullptr || act_tma));
  if (t.splits > 1) ok = make_map_f32(&mC, partial, N, (int64_t)t.splits * M, N, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
  else if (fast_epi) {
    if (!aligned16(C) || ldc % 4 != 0) return 0;
    ok = make_map_f32(&mC, C, ldc, M, N, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
  } else mC = mA;
  GOTEN_REQUIRE(ok, "cuTensorMapEncodeTiled failed for the output (M=%d N=%d ldc=%d)", M, N, ldc);
  mAct = mC;
  if (act_tma && t.splits == 1) {
    ok = make_map_f32(&mAct, act_out, ld_act, M, act_hi - act_lo, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
    GOTEN_REQUIRE(ok, "cuTensorMapEncodeTiled failed for the activation output (M=%d cols=%d ld=%d)", M, act_hi - act_lo, ld_act);
  }

  tc16::Params p{};
  p.M = M; p.N = N; p.K = K;
  p.block_n = t.block_n; p.n_mt = t.n_mt; p.n_nt = t.n_nt;
  p.splits = t.splits; p.kb_per_split = t.kb_per_split; p.kb_total = t.kb_total;
  p.stages = t.stages;
  p.C = C; p.ldc = ldc; p.bias = bias; p.add_src = add_src; p.ld_add = ld_add;
  p.act_out = act_out; p.ld_act = ld_act; p.act_lo = act_lo; p.act_hi = act_hi;
  p.red_add = red_add ? 1 : 0;
  p.act_tma = act_tma ? 1 : 0;
  {
    static int atm = -1;   // GOTEN_GEMM_ATMEM=1: A operand through tensor memory (K-major A only)
    if (atm < 0) { const char* e = getenv("GOTEN_GEMM_ATMEM"); atm = e ? atoi(e) : 0; }
    p.a_tmem = (atm && !t.a_rows_are_k) ? 1 : 0;
  }
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("GOTEN_GEMM_DBG"); dbg = e ? atoi(e) : 0; } p.dbg = dbg; }
  p.add_vec = (add_src != nullptr && aligned16(add_src) && ld_add % 4 == 0) ? 1 : 0;
  p.c_vec = (aligned16(C) && ldc % 4 == 0) ? 1 : 0;
  p.act_vec = (act_out != nullptr && aligned16(act_out) && ld_act % 4 == 0 && act_lo % 4 == 0) ? 1 : 0;
  p.partial = partial; p.colsum = colsum; p.partial_colsum = partial_cs;
  p.amax_a = pa; p.amax_b = pb;
  if (t.splits > 1) { p.bias = nullptr; p.add_src = nullptr; p.act_out = nullptr; }

  const size_t smem = t.panel ? tc16::P_SMEM
                              : 1024 + (size_t)t.stages * (tc16::A_BYTES + 2 * (size_t)(t.block_n / t.ncta) * tc16::BK * 2) +
                                    4 * 2 * 4096 + (3 * tc16::MAX_STAGES + 4) * 8 + 16;
  GOTEN_REQUIRE((int)smem <= smem_optin, "tcgen05 fp16 GEMM needs %zu B of shared memory", smem);
  const int n_items = t.panel ? t.n_mt : t.n_mt * t.n_nt * t.splits;   // panel kernel: a work item is an M block
  const int max_units = sm_count / t.ncta;
  const int grid = (n_items < max_units ? n_items : max_units) * t.ncta;
  // short-K K-major shapes are epilogue-bound: they run with eight epilogue warps (K <= GOTEN_GEMM_EPI8_K, default 512;
  // GOTEN_GEMM_EPI8=0 switches the form off).  Measured, 4 -> 8 warps: edge projection 811 -> 779 us, node projections
  // 44.3 -> 41.1 / 50.9 -> 47.2 / 64.7 -> 54.7 (SiLU side output), EQ|EK 83.6 -> 79.4; long K unchanged.
  static int epi8_on = -1, epi8_k = 512;
  if (epi8_on < 0) {
    const char* e = getenv("GOTEN_GEMM_EPI8"); epi8_on = e ? atoi(e) : 1;
    const char* k = getenv("GOTEN_GEMM_EPI8_K"); if (k) epi8_k = atoi(k);
  }
  // eight converter warps for the weight-gradient form that converts BOTH operands in the kernel (b_raw: two
  // transposing passes per k-block): 82 -> 77, 94 -> 83, 60 -> 54.5, 240 -> 182 us on the node-sized shapes; the forms
  // with a pre-split B do not gain (GOTEN_GEMM_CONV8=0 off, =2 every weight-gradient form)
  static int conv8_mode = -1;
  if (conv8_mode < 0) { const char* e = getenv("GOTEN_GEMM_CONV8"); conv8_mode = e ? atoi(e) : 1; }
  const bool conv8 = conv8_mode == 2 || (conv8_mode == 1 && t.b_raw);
  const bool epi8 = epi8_on && !t.panel && !t.a_rows_are_k && !t.b_raw && K <= epi8_k && !p.a_tmem;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(tc16::NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)t.ncta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#define GOTEN_TC16_LAUNCH(RK, BR, NC)                                                                          \
  do {                                                                                                      \
    auto k = tc16::gemm16_kernel<RK, BR, NC>;                                                                 \
    static int smem_set = 0;                                                                                \
    if ((int)smem > smem_set) {                                                                             \
      GOTEN_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin));   \
      smem_set = smem_optin;                                                                                \
    }                                                                                                       \
    GOTEN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k, mA, mBh, mBl, mC, mAct, p));                                     \
  } while (0)
  if (t.panel) {
    auto k = tc16::gemm16_panel_kernel;
    static int smem_set_p = 0;
    if (!smem_set_p) {
      GOTEN_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin));
      smem_set_p = 1;
    }
    GOTEN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k, mA, mBh, mBl, mC, mAct, p));
  } else if (t.a_rows_are_k && t.ncta == 2 && conv8) {
    // weight-gradient forms with eight converter warps (448 threads)
#define GOTEN_TC16_LAUNCH8(BR)                                                                               \
  do {                                                                                                      \
    auto k = tc16::gemm16_kernel<true, BR, 2, 4, 8>;                                                         \
    static int smem_set_c8 = 0;                                                                             \
    if (!smem_set_c8) {                                                                                     \
      GOTEN_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 1024)); /* 512 B static */ \
      smem_set_c8 = 1;                                                                                      \
    }                                                                                                       \
    cfg.blockDim = dim3(448);                                                                               \
    GOTEN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k, mA, mBh, mBl, mC, mAct, p));                               \
  } while (0)
    if (t.b_raw) GOTEN_TC16_LAUNCH8(true); else GOTEN_TC16_LAUNCH8(false);
#undef GOTEN_TC16_LAUNCH8
  } else if (t.b_raw) {
    if (t.ncta == 2) GOTEN_TC16_LAUNCH(true, true, 2); else GOTEN_TC16_LAUNCH(true, true, 1);
  } else if (t.a_rows_are_k) {
    if (t.ncta == 2) GOTEN_TC16_LAUNCH(true, false, 2); else GOTEN_TC16_LAUNCH(true, false, 1);
  } else if (t.ncta == 2 && epi8) {
    // eight epilogue warps (448 threads)
    auto k = tc16::gemm16_kernel<false, false, 2, 8>;
    static int smem_set8 = 0;
    if (!smem_set8) {
      GOTEN_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin));
      smem_set8 = 1;
    }
    cfg.blockDim = dim3(448);
    GOTEN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k, mA, mBh, mBl, mC, mAct, p));
  } else {
    if (t.ncta == 2) GOTEN_TC16_LAUNCH(false, false, 2); else GOTEN_TC16_LAUNCH(false, false, 1);
  }
#undef GOTEN_TC16_LAUNCH
  GOTEN_CHECK_LAUNCH();
  if (t.splits > 1) {
    if (splitk_finish(partial, partial_cs, t.splits, C, ldc, M, N, bias, add_src, ld_add, act_out, ld_act, act_lo,
                      act_hi, colsum, st))
      return 1;
  }
  *handled = true;
  return 0;
}

}  // namespace goten

using namespace goten;

extern "C" {

int goten_absmax_multi(const float* const* ptrs, const int64_t* numel, int count, float* out, void* stream) {
  GOTEN_REQUIRE(count >= 0 && count <= 16, "goten_absmax_multi takes at most 16 tensors (got %d)", count);
  if (count == 0) return 0;
  tc16::AbsmaxMulti a{};
  int64_t nmax = 0;
  for (int i = 0; i < count; ++i) {
    a.ptr[i] = ptrs[i];
    a.n[i] = numel[i];
    nmax = numel[i] > nmax ? numel[i] : nmax;
  }
  int64_t gx = cdiv64(nmax, 256 * 16);
  gx = gx < 1 ? 1 : (gx > 64 ? 64 : gx);
  tc16::absmax_multi_kernel<<<dim3((unsigned)gx, (unsigned)count), 256, 0, as_stream(stream)>>>(a, out);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_absmax(const float* A, int64_t lda, int64_t M, int N, float* out, void* stream) {
  if (M <= 0 || N <= 0) return 0;
  int dev = 0, sms = 0;
  GOTEN_CHECK_CUDA(cudaGetDevice(&dev));
  GOTEN_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  return launch_absmax(A, lda, M, N, out, sms, as_stream(stream));
}

}  // extern "C"
