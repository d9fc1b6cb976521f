// tcgen05 / TMA / mbarrier PTX wrappers shared by the tensor-core GEMM kernels (sm_100a).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace goten {
namespace tc {

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded spin: a protocol bug traps (clean CUDA error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// ---- CTA-pair (cta_group::2) forms: one MMA spans two SMs (M = 256), each CTA stages its own 128 rows of A and
// half of the B tile; barriers that gate the issuing (leader) CTA receive remote arrives from the peer.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster.  Default (.release.cta)
// semantics as in CUTLASS' ClusterBarrier::arrive: the explicit .release.cluster form compiles to MEMBAR + ERRBAR
// and cost 17 % of all stall samples on the converter's critical path (profiles/r1f).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(bar),
      "r"(rank)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// completion of all prior MMAs -> arrive on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
// layout_type: 2 = SWIZZLE_128B (16 B chunks), 1 = SWIZZLE_128B_BASE32B (32 B chunks; the only swizzled
// layout tcgen05 accepts for MN-major 32-bit operands)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;
  return d;
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1)
               : "memory");
}
// global[tile] += smem tile (f32 add performed by the TMA unit at L2; every element is touched by exactly one
// reduction, so the result is as deterministic as a plain store)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// ---- split-fp16 operand helpers (gemm_tc16.cu, edge_fused.cu)
// power-of-two scale that maps a tensor maximum into [2^14, 2^15), and its inverse (both exact floats)
__device__ __forceinline__ float scale_of(float amax) {
  const int e = (int)((__float_as_uint(amax) >> 23) & 0xFF);
  if (amax == 0.f || e == 255) return 1.0f;
  int sb = 268 - e;  // biased exponent of 2^(14 - (e - 127))
  sb = sb < 1 ? 1 : (sb > 253 ? 253 : sb);
  return __uint_as_float((uint32_t)sb << 23);
}
__device__ __forceinline__ float inv_scale_of(float amax) {
  const float s = scale_of(amax);
  return __uint_as_float((uint32_t)(254 - (int)(__float_as_uint(s) >> 23)) << 23);
}

// hi/lo split of two scaled values -> packed half2 words
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
template <int NT = 128>
__device__ __forceinline__ void conv_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }
// A operand from tensor memory (lane = row of A, 32-bit column j = K elements 2j, 2j + 1), B from a shared-memory descriptor
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ts_2cta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// registers -> 32 consecutive 32-bit columns of this thread's TMEM lane; complete after tmem_wait_st()
#define GOTEN_STTM_X32(taddr, r)                                                                                      \
  asm volatile(                                                                                                       \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                                 \
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "                                      \
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"                              \
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),  \
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),    \
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),   \
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                                                    \
      : "memory")
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }


// ---------------------------------------------------------------------------------------------------------------
// Epilogue shared by the two tensor-core GEMM kernels (warps EPI_WARP0 .. EPI_WARP0 + 3 of a 320-thread CTA).
// Each warp owns a 32-row band of the tile.  Per 32-column chunk: tcgen05.ld (thread = row), un-scale (un_a * un_b;
// 1 for the 3xTF32 arm), + bias, write the 32x32 block into a 128B-swizzled smem buffer, then either
//   fast path : one TMA store of the block (plain / bias / split-K partial outputs), double buffered
//   fused path: read the block back row-wise and do coalesced global stores with the residual add and the SiLU side
//               output (8 lanes x float4 = one 128 B row segment, 4 rows per instruction).
// P needs: M, N, C, ldc, bias, add_src, ld_add, act_out, ld_act, act_lo, act_hi, add_vec, c_vec, act_vec, partial, red_add, act_tma.
// explicit shared-space vector accesses of the staging blocks (a generic pointer makes the compiler emit ST.E / LD.E)
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
// 32 consecutive fp32 accumulator columns of this thread's TMEM lane; completes at the next tcgen05.wait::ld
#define GOTEN_LDTM_X32(r, taddr)                                                                                      \
  asm volatile(                                                                                                       \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                       \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                       \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                       \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),    \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),        \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),       \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                     \
      : "r"(taddr))
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// One 32x32 block of the epilogue: registers (thread = row, 32 columns) -> un-scale + bias -> swizzled smem block ->
// TMA store(s) (fast) or the row-wise register path with residual / SiLU side output.  `blv`: the tile's bias, lane l
// holding columns n0 + 8 l .. 8 l + 7; column c of chunk ch sits in lane 4 ch + c / 8, slot c % 8.
template <int NBUF, class P>
__device__ __forceinline__ void epilogue_block(const P& p, const CUtensorMap& tmC, const CUtensorMap& tmAct, const uint32_t (&r)[32],
                                               const float (&blv)[8], int ch, int nc0, int row_base, int split,
                                               uint32_t my_buf, uint32_t& n_store, int lane, float un, bool fast) {
  if (p.dbg & 8) return;
  // NBUF staging blocks per warp (2: the TMA store of block k is read while block k + 1 is written; 1: the eight-warp
  // form, where the second warp of the scheduler covers the wait)
  const uint32_t buf = my_buf + (n_store % NBUF) * 4096;
  if (fast && n_store >= (uint32_t)NBUF) {
    if (lane == 0) bulk_wait_read<NBUF - 1>();
    __syncwarp();
  }
  const int l0 = ch * 4;
  const bool has_bias = p.bias != nullptr && !p.partial;
  if (!has_bias) {
    // (no bias: data / weight gradients, split-K partials) un-scale only
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float4 v;
      v.x = __uint_as_float(r[4 * c + 0]) * un; v.y = __uint_as_float(r[4 * c + 1]) * un;
      v.z = __uint_as_float(r[4 * c + 2]) * un; v.w = __uint_as_float(r[4 * c + 3]) * un;
      sts128(buf + lane * 128 + ((c ^ (lane & 7)) << 4), v);
    }
  } else if (nc0 + 32 <= p.N && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
    // the 32 bias values of the chunk are the same for every lane: eight broadcast 16-byte loads (one L1 sector each),
    // all issued before the first use - the 32 dependent SHFL -> FFMA pairs of the register-distributed form were the
    // epilogue warps' main stall (ncu source view, round 2).  (Hoisting these loads above the tcgen05.wait::ld of the
    // chunk was measured too: the eight live float4 spill, and every shape got 10 % slower; staging the tile's bias in
    // shared memory - two named barriers per tile among the four warps, LDS.128 broadcasts - cost 6 %.)
    float4 bq[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) bq[c] = __ldg(reinterpret_cast<const float4*>(p.bias + nc0) + c);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float4 v;
      v.x = fmaf(__uint_as_float(r[4 * c + 0]), un, bq[c].x); v.y = fmaf(__uint_as_float(r[4 * c + 1]), un, bq[c].y);
      v.z = fmaf(__uint_as_float(r[4 * c + 2]), un, bq[c].z); v.w = fmaf(__uint_as_float(r[4 * c + 3]), un, bq[c].w);
      sts128(buf + lane * 128 + ((c ^ (lane & 7)) << 4), v);
    }
  } else {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float4 v;
    v.x = fmaf(__uint_as_float(r[4 * c + 0]), un, __shfl_sync(0xffffffffu, blv[(4 * c + 0) & 7], l0 + (c >> 1)));
    v.y = fmaf(__uint_as_float(r[4 * c + 1]), un, __shfl_sync(0xffffffffu, blv[(4 * c + 1) & 7], l0 + (c >> 1)));
    v.z = fmaf(__uint_as_float(r[4 * c + 2]), un, __shfl_sync(0xffffffffu, blv[(4 * c + 2) & 7], l0 + (c >> 1)));
    v.w = fmaf(__uint_as_float(r[4 * c + 3]), un, __shfl_sync(0xffffffffu, blv[(4 * c + 3) & 7], l0 + (c >> 1)));
    sts128(buf + lane * 128 + ((c ^ (lane & 7)) << 4), v);
  }
  }
  if (fast) {
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      if (p.partial) tma_store_2d(&tmC, buf, nc0, split * p.M + row_base);
      else if (p.red_add) tma_reduce_add_2d(&tmC, buf, nc0, row_base);
      else tma_store_2d(&tmC, buf, nc0, row_base);
      bulk_commit();
    }
    ++n_store;
    if (p.act_tma && nc0 >= p.act_lo && nc0 < p.act_hi) {   // warp-uniform: SiLU side output of this block
      const uint32_t buf2 = my_buf + (n_store % NBUF) * 4096;   // (NBUF = 1: rewritten in place once the store has read it)
      if (n_store >= (uint32_t)NBUF) {
        if (lane == 0) bulk_wait_read<NBUF - 1>();
        __syncwarp();
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4 v = lds128(buf + lane * 128 + ((c ^ (lane & 7)) << 4));
        v.x = __fdividef(v.x, 1.0f + __expf(-v.x)); v.y = __fdividef(v.y, 1.0f + __expf(-v.y));
        v.z = __fdividef(v.z, 1.0f + __expf(-v.z)); v.w = __fdividef(v.w, 1.0f + __expf(-v.w));
        sts128(buf2 + lane * 128 + ((c ^ (lane & 7)) << 4), v);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tmAct, buf2, nc0 - p.act_lo, row_base);
        bulk_commit();
      }
      ++n_store;
    }
    return;
  }
  __syncwarp();
  const int cq = lane & 7, rsub = lane >> 3;
  const int n = nc0 + cq * 4;
  const bool vec_ok = (n + 3 < p.N);
  float4 addv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = i * 4 + rsub, m = row_base + rr;
    addv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.add_src && m < p.M) {
      const float* ap = p.add_src + (size_t)m * p.ld_add + n;
      if (vec_ok && p.add_vec) addv[i] = *reinterpret_cast<const float4*>(ap);
      else {
        if (n + 0 < p.N) addv[i].x = ap[0];
        if (n + 1 < p.N) addv[i].y = ap[1];
        if (n + 2 < p.N) addv[i].z = ap[2];
        if (n + 3 < p.N) addv[i].w = ap[3];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = i * 4 + rsub, m = row_base + rr;
    float4 v = lds128(buf + rr * 128 + ((cq ^ (rr & 7)) << 4));
    v.x += addv[i].x; v.y += addv[i].y; v.z += addv[i].z; v.w += addv[i].w;
    if (m < p.M) {
      float* cp = p.C + (size_t)m * p.ldc + n;
      if (vec_ok && p.c_vec) *reinterpret_cast<float4*>(cp) = v;
      else {
        if (n + 0 < p.N) cp[0] = v.x;
        if (n + 1 < p.N) cp[1] = v.y;
        if (n + 2 < p.N) cp[2] = v.z;
        if (n + 3 < p.N) cp[3] = v.w;
      }
      if (p.act_out && p.act_vec && n >= p.act_lo && n + 3 < p.act_hi) {  // whole float4 inside the SiLU range
        float4 a;
        a.x = __fdividef(v.x, 1.0f + __expf(-v.x)); a.y = __fdividef(v.y, 1.0f + __expf(-v.y));
        a.z = __fdividef(v.z, 1.0f + __expf(-v.z)); a.w = __fdividef(v.w, 1.0f + __expf(-v.w));
        *reinterpret_cast<float4*>(p.act_out + (size_t)m * p.ld_act + (n - p.act_lo)) = a;
      } else if (p.act_out) {
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          const int nn = n + qd;
          if (nn < p.N && nn >= p.act_lo && nn < p.act_hi)
            p.act_out[(size_t)m * p.ld_act + (nn - p.act_lo)] = vv[qd] / (1.0f + __expf(-vv[qd]));
        }
      }
    }
  }
  __syncwarp();
}

template <int NCTA, int BM, int EPI_WARP0, int EPIW = 4, class P>
__device__ __forceinline__ void gemm_epilogue(const P& p, const CUtensorMap& tmC, const CUtensorMap& tmAct, uint8_t* epi_smem,
                                              uint32_t bar_tfull,
                                              uint32_t bar_tempty, uint32_t tmem_base, int warp, int lane, int unit,
                                              int n_units, int n_items, int n_tiles, int BN, uint32_t rank, float un_a,
                                              float un_b, int acc_stages = 2, bool panel_order = false) {
  // =============================== epilogue ===================================
  // The four warps were the slowest stage of the short-K GEMMs (ncu source view, round 2: busy 87 % of the time, one
  // warp per scheduler running a serial chain tcgen05.ld -> bias LDG -> generic ST -> fence -> TMA per 32-column
  // chunk).  Now: the tile's bias is fetched before the accumulator wait, the TMEM load of chunk ch + 1 is in flight
  // while chunk ch is processed (two register buffers), the accumulator stage is handed back to the MMA warp as soon
  // as the last load has landed, and the staging block is written with st.shared.
  const int q = warp & 3;
  // EPIW = 8: two warps per TMEM lane quadrant (warp % 4), the second set takes the odd 32-column chunks; the 32 KB of
  // staging blocks then give one block per warp instead of two
  constexpr int NBUF = EPIW == 4 ? 2 : 1, CST = EPIW / 4;
  const int cs = (warp - EPI_WARP0) / 4;
  const uint32_t my_buf = smem_u32(epi_smem) + (uint32_t)(warp - EPI_WARP0) * NBUF * 4096;
  // red_add: the residual already sits in C (add_src == C): the tile is ADDED to it by a TMA reduction store, so the
  // residual never passes through the SM (the register path moves 4 KB per warp and round trip)
  // act_tma: the SiLU side output (columns [act_lo, act_hi), act_lo a multiple of 32) leaves through a second TMA store of
  // the same 32x32 block (tmAct is a map over act_out with act_hi - act_lo columns, so the range end clips for free)
  const bool fast = (p.add_src == nullptr || p.red_add) && (p.act_out == nullptr || p.act_tma);
  const float un = un_a * un_b;
  uint32_t tile_it = 0, n_store = 0;
  // default order: work items unit, unit + n_units, ... over (split, M tile, N tile) with N fastest; panel order (the
  // A-stationary kernel): M blocks unit, unit + n_units, ..., all N tiles of a block in turn
  for (int w = unit;; ++tile_it) {
    int split, m0, n0;
    if (!panel_order) {
      if (w >= n_items) break;
      const int tile = w % n_tiles;
      split = w / n_tiles;
      m0 = (tile / p.n_nt) * (BM * NCTA) + (int)rank * BM;
      n0 = (tile % p.n_nt) * BN;
      w += n_units;
    } else {
      const int mb = unit + (int)(tile_it / (uint32_t)p.n_nt) * n_units;
      if (mb >= p.n_mt) break;
      split = 0;
      m0 = mb * (BM * NCTA) + (int)rank * BM;
      n0 = (int)(tile_it % (uint32_t)p.n_nt) * BN;
    }
    // two accumulator stages alternate; with one (A operand resident in tensor memory) its barrier flips every tile
    const uint32_t acc = acc_stages == 2 ? (tile_it & 1) : 0u, aph = acc_stages == 2 ? ((tile_it >> 1) & 1) : (tile_it & 1);
    const int row_base = m0 + q * 32;
    const bool rows_live = row_base < p.M;
    const int nch = rows_live ? min(BN, p.N - n0 + 31) / 32 : 0;   // 32-column chunks with a live column
    float blv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) blv[k] = 0.f;
    if (p.bias && !p.partial && nch > 0) {
      const int nb = n0 + lane * 8;
      if (nb + 7 < p.N && (reinterpret_cast<uintptr_t>(p.bias + nb) & 15) == 0) {
        const float4 b0 = *reinterpret_cast<const float4*>(p.bias + nb), b1 = *reinterpret_cast<const float4*>(p.bias + nb + 4);
        blv[0] = b0.x; blv[1] = b0.y; blv[2] = b0.z; blv[3] = b0.w;
        blv[4] = b1.x; blv[5] = b1.y; blv[6] = b1.z; blv[7] = b1.w;
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (nb + k < p.N) blv[k] = p.bias[nb + k];
      }
    }
    mbar_wait(bar_tfull + 8 * acc, aph);
    tc_fence_after();
    const uint32_t tbase = tmem_base + acc * (uint32_t)BN + ((uint32_t)(q * 32) << 16);
    uint32_t rA[32], rB[32];
    bool released = false;
    auto release = [&]() {   // every tcgen05.ld of this tile has completed: the MMA warp may overwrite the stage
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (NCTA == 2) mbar_arrive_cluster(bar_tempty + 8 * acc, 0);
        else mbar_arrive(bar_tempty + 8 * acc);
      }
      released = true;
    };
    if (cs < nch) GOTEN_LDTM_X32(rA, tbase + cs * 32);
    for (int ch = cs; ch < nch; ch += 2 * CST) {
      tmem_wait_ld();
      if (ch + CST < nch) GOTEN_LDTM_X32(rB, tbase + (ch + CST) * 32);
      else release();
      epilogue_block<NBUF>(p, tmC, tmAct, rA, blv, ch, n0 + ch * 32, row_base, split, my_buf, n_store, lane, un, fast);
      if (ch + CST < nch) {
        tmem_wait_ld();
        if (ch + 2 * CST < nch) GOTEN_LDTM_X32(rA, tbase + (ch + 2 * CST) * 32);
        else release();
        epilogue_block<NBUF>(p, tmC, tmAct, rB, blv, ch + CST, n0 + (ch + CST) * 32, row_base, split, my_buf, n_store, lane, un,
                             fast);
      }
    }
    if (!released) release();
  }
  if (fast && lane == 0) bulk_wait_all();
}

}  // namespace tc

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline int64_t align256(int64_t x) { return (x + 255) & ~int64_t(255); }

}  // namespace goten
