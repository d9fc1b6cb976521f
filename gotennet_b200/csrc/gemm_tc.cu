// tcgen05 3xTF32 GEMM for sm_100a: fp32-accurate tensor-core GEMM
//     D = A_hi B_hi + A_hi B_lo + A_lo B_hi,   x_hi = rna_tf32(x), x_lo = x - x_hi (exact in fp32)
// (the reference runs its nn.Linear layers in strict fp32, scripts/train.py:16; a single TF32
// pass misses the 1e-4 parity bar, three passes sit at ~1e-6).
//
// Structure (persistent, warp specialised, one CTA per SM, 320 threads):
//   warp 0      TMA producer: A (raw fp32) + pre-split B_hi / B_lo tiles -> 128B-swizzled smem ring
//   warp 1      TMEM allocator + MMA issuer (one elected lane): 12 x tcgen05.mma.kind::tf32
//               (M128 x N<=256 x K8) per 32-deep k-block into a double-buffered TMEM accumulator
//   warps 2-5   epilogue: tcgen05.ld -> warp-private smem transpose -> coalesced stores with the
//               fused bias / residual-add / SiLU side output (or raw split-K partials)
//   warps 6-9   converter: splits the raw A tile in place into hi (tf32-rounded) and lo tiles
//               (generic-proxy smem writes + fence.proxy.async), and accumulates the column
//               sums of A for the fused bias gradient in the weight-gradient GEMM
// Operand layouts: K-major for forward / data-gradient GEMMs (A[M][K], B[N][K]); MN-major for the
// weight-gradient GEMM (A[R][M], B[R][N], reduction over rows R) via 3-D tensor maps.
#include "umma.cuh"

namespace goten {

namespace tc {

constexpr int BM = 128;          // UMMA M per CTA (cta_group::2: the pair computes M = 256)
constexpr int BK = 32;           // floats per k-block = one 128 B swizzle row
constexpr int MAX_STAGES = 4;    // smem ring depth is chosen by the host (Params::stages)
constexpr int NTHREADS = 320;
constexpr int EPI_WARP0 = 2, CONV_WARP0 = 6;

struct Params {
  int M, N, K;            // output M x N, reduction K
  int block_n;            // UMMA N (64 / 128 / 256)
  int n_mt, n_nt;         // tiles
  int splits, kb_per_split, kb_total;
  int stages;             // smem ring depth (<= MAX_STAGES)
  int dbg;                // GOTEN_GEMM_DBG experiment switches (timing only; results are wrong when set)
  float* C; int ldc;
  const float* bias;
  const float* add_src; int ld_add;
  float* act_out; int ld_act, act_lo, act_hi;
  int add_vec, c_vec, act_vec;
  // 16 B alignment of the add_src / C / act_out rows (vector epilogue accesses allowed)
  int red_add;            // add_src == C: accumulate into C with a TMA reduction store
  int act_tma;            // SiLU side output through a second TMA store (tmAct)
  float* partial;         // split-K partials [splits][M][N] (nullptr: direct epilogue)
  float* colsum;          // column sums of A (MN-major only), direct
  float* partial_colsum;  // [splits][M]
};

template <bool MN_MAJOR, int NCTA>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm3x_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
              const __grid_constant__ CUtensorMap tmBl, const __grid_constant__ CUtensorMap tmC,
              const __grid_constant__ CUtensorMap tmAct, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024 B alignment for the 128 B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int BN = p.block_n;                        // UMMA N (whole tile width)
  const int BNH = BN / NCTA;                       // B rows staged by this CTA
  const int STAGES = p.stages;
  const uint32_t A_BYTES = BM * BK * 4;            // 16 KB
  const uint32_t B_BYTES = (uint32_t)BNH * BK * 4; // <= 32 KB
  const uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  uint8_t* epi_smem = smem + STAGES * STAGE_BYTES;                    // 4 warps x 2 x [32 rows][128 B], 128B-swizzled
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + 4 * 2 * 4096);
  // bars: full[MAX_STAGES], conv[MAX_STAGES], empty[MAX_STAGES], tmem_full[2], tmem_empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 4);
  const uint32_t bar_full = smem_u32(bars), bar_conv = smem_u32(bars + MAX_STAGES), bar_empty = smem_u32(bars + 2 * MAX_STAGES);
  const uint32_t bar_tfull = smem_u32(bars + 3 * MAX_STAGES), bar_tempty = smem_u32(bars + 3 * MAX_STAGES + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = NCTA == 2 ? cluster_ctarank() : 0u;   // rank 0 = leader: issues the MMAs
  const int unit = NCTA == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;          // persistent work-unit id
  const int n_units = NCTA == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_conv + 8 * s, 4 * NCTA);   // one arrive per converter warp (of both CTAs: leader's barrier)
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 4 * NCTA);  // one arrive per epilogue warp (of both CTAs: leader's barrier)
    }
    fence_barrier_init();
  }
  if (warp == 1) {  // TMEM: 512 columns = two accumulator stages of up to 256 columns
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
          const uint32_t sbh = sa + 2 * A_BYTES, sbl = sbh + B_BYTES;
          const uint32_t bar = bar_full + 8 * s;
          mbar_arrive_expect_tx(bar, A_BYTES + 2 * B_BYTES);
          if (!MN_MAJOR) {
            tma_load_2d(sa, &tmA, bar, kb * BK, m0);
            tma_load_2d(sbh, &tmBh, bar, kb * BK, n0);
            tma_load_2d(sbl, &tmBl, bar, kb * BK, n0);
          } else {
            tma_load_3d(sa, &tmA, bar, 0, kb * BK, m0 / 32);
            tma_load_3d(sbh, &tmBh, bar, 0, kb * BK, n0 / 32);
            tma_load_3d(sbl, &tmBl, bar, 0, kb * BK, n0 / 32);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((MN_MAJOR ? 1u : 0u) << 15) |
                             ((MN_MAJOR ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * NCTA) >> 4) << 24);
      // K-major: 8-row groups 1024 B apart, k-step = +32 B.  MN-major: 32-wide MN chunks BK*128 B apart (LBO),
      // 8-deep k atoms 1024 B apart (SBO), k-step = +1024 B.
      // MN-major uses the 32B-base swizzle: k atoms are 4 rows (512 B) deep, two per K=8 instruction.
      const uint32_t lbo = MN_MAJOR ? BK * 128 : 16, sbo = MN_MAJOR ? 512 : 1024, kstep = MN_MAJOR ? 1024 : 32;
      const uint32_t lt = MN_MAJOR ? 1 : 2;
      uint32_t it = 0, tile_it = 0;
      for (int w = unit; w < n_items; w += n_units, ++tile_it) {
        const int split = w / n_tiles;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const uint32_t acc = tile_it & 1, aph = (tile_it >> 1) & 1;
        mbar_wait(bar_tempty + 8 * acc, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(bar_full + 8 * s, ph);
          mbar_wait(bar_conv + 8 * s, ph);   // converters of both CTAs: implies the peer's TMA data has landed too
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t sal = sa + A_BYTES, sbh = sa + 2 * A_BYTES, sbl = sbh + B_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t a_hi = make_desc(sa + kk * kstep, lbo, sbo, lt), a_lo = make_desc(sal + kk * kstep, lbo, sbo, lt);
            const uint64_t b_hi = make_desc(sbh + kk * kstep, lbo, sbo, lt), b_lo = make_desc(sbl + kk * kstep, lbo, sbo, lt);
            if (NCTA == 2) {
              umma_tf32_2cta(d_tmem, a_lo, b_hi, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
              if (!(p.dbg & 1)) umma_tf32_2cta(d_tmem, a_hi, b_lo, idesc, 1u);
              if (!(p.dbg & 2)) umma_tf32_2cta(d_tmem, a_hi, b_hi, idesc, 1u);
            } else {
              umma_tf32(d_tmem, a_lo, b_hi, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
              if (!(p.dbg & 1)) umma_tf32(d_tmem, a_hi, b_lo, idesc, 1u);
              if (!(p.dbg & 2)) umma_tf32(d_tmem, a_hi, b_hi, idesc, 1u);
            }
          }
          // smem stage reusable (in both CTAs) once these MMAs have read it
          if (NCTA == 2) umma_commit_2cta(bar_empty + 8 * s); else umma_commit(bar_empty + 8 * s);
        }
        if (NCTA == 2) umma_commit_2cta(bar_tfull + 8 * acc); else umma_commit(bar_tfull + 8 * acc);  // accumulator complete
      }
    }
  } else if (warp >= CONV_WARP0) {
    // =============================== converter ==================================
    const int ct = threadIdx.x - CONV_WARP0 * 32;  // 0..127
    uint32_t it = 0;
    for (int w = unit; w < n_items; w += n_units) {
      const int split = w / n_tiles, tile = w % n_tiles;
      const int m0 = (tile / p.n_nt) * (BM * NCTA) + (int)rank * BM;
      const bool do_cs = MN_MAJOR && (p.colsum != nullptr) && (tile % p.n_nt == 0);
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
      float csum = 0.f;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(bar_full + 8 * s, ph);
        uint8_t* a_raw = smem + s * STAGE_BYTES;
        uint8_t* a_lo = a_raw + A_BYTES;
        if (!MN_MAJOR) {
          // thread = tile row; rotate the 16 B chunk order so a quarter warp hits 8 distinct bank groups.
          // All eight loads are issued before the first store (the in-place stores would otherwise fence them).
          float4 v[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) v[c] = *reinterpret_cast<const float4*>(a_raw + ct * 128 + (((c + ct) & 7) << 4));
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int pc = (c + ct) & 7;
            float4 h, l;
            h.x = tf32_rna(v[c].x); h.y = tf32_rna(v[c].y); h.z = tf32_rna(v[c].z); h.w = tf32_rna(v[c].w);
            l.x = v[c].x - h.x; l.y = v[c].y - h.y; l.z = v[c].z - h.z; l.w = v[c].w - h.w;
            *reinterpret_cast<float4*>(a_raw + ct * 128 + pc * 16) = h;
            *reinterpret_cast<float4*>(a_lo + ct * 128 + pc * 16) = l;
          }
        } else {
          // thread = logical MN column (chunk = ct/32, col = ct%32); walks the 32 k rows of the chunk
          const int chunk = ct >> 5, col = ct & 31;
          const int cw = col >> 3, wi = col & 7;  // 32 B chunk index / word inside it (SWIZZLE_128B_ATOM_32B)
#pragma unroll
          for (int k0 = 0; k0 < BK; k0 += 8) {  // batches of 8 loads ahead of the in-place stores
            float v[8];
            int off[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int k = k0 + u;
              off[u] = chunk * (BK * 128) + k * 128 + ((cw ^ (k & 3)) << 5) + wi * 4;
              v[u] = *reinterpret_cast<const float*>(a_raw + off[u]);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float h = tf32_rna(v[u]);
              *reinterpret_cast<float*>(a_raw + off[u]) = h;
              *reinterpret_cast<float*>(a_lo + off[u]) = v[u] - h;
              csum += v[u];
            }
          }
        }
        fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) {
          if (NCTA == 2) mbar_arrive_cluster(bar_conv + 8 * s, 0);  // the leader's barrier gates the MMA issue
          else mbar_arrive(bar_conv + 8 * s);
        }
      }
      if (do_cs) {
        const int m = m0 + ct;
        if (m < p.M) {
          if (p.partial_colsum) p.partial_colsum[(size_t)split * p.M + m] = csum;
          else p.colsum[m] = csum;
        }
      }
    }
  } else {
    // =============================== epilogue (umma.cuh) ========================
    gemm_epilogue<NCTA, BM, EPI_WARP0>(p, tmC, tmAct, epi_smem, bar_tfull, bar_tempty, tmem_base, warp, lane, unit, n_units, n_items,
                                       n_tiles, BN, rank, 1.0f, 1.0f);
  }

  tc_fence_before();
  __syncthreads();  // (the pair also meets below: compute-sanitizer racecheck only models the CTA barrier)
  if (NCTA == 2) cluster_sync_all();   // the peer may still signal / read this CTA until here
  if (warp == 1) {
    tc_fence_after();
    if (NCTA == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  }
}

// B-operand preparation: hi = rna_tf32(b), lo = b - hi.
// Row-wise form ([R][C] -> [R][C]): 128-bit loads/stores, grid-stride, no integer division per element.
__global__ void split_tf32_vec4_kernel(const float* __restrict__ in, int ld, int rows, int cols4,
                                       float* __restrict__ hi, float* __restrict__ lo) {
  const int64_t total = (int64_t)rows * cols4;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(idx / cols4), c4 = (int)(idx - (int64_t)r * cols4);
    const float4 v = *reinterpret_cast<const float4*>(in + (int64_t)r * ld + 4 * c4);
    float4 h, l;
    h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
    reinterpret_cast<float4*>(hi)[idx] = h;
    reinterpret_cast<float4*>(lo)[idx] = l;
  }
}
// scalar row-wise form for unaligned shapes
__global__ void split_tf32_kernel(const float* __restrict__ in, int ld, int rows, int cols,
                                  float* __restrict__ hi, float* __restrict__ lo) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)rows * cols) return;
  const int r = (int)(idx / cols), c = (int)(idx - (int64_t)r * cols);
  const float v = in[(int64_t)r * ld + c];
  const float h = tf32_rna(v);
  hi[idx] = h;
  lo[idx] = v - h;
}
// transposing form ([R][C] -> [C][R]) through a 32x33 shared-memory tile: coalesced on both sides
__global__ void split_tf32_transpose_kernel(const float* __restrict__ in, int ld, int rows, int cols,
                                            float* __restrict__ hi, float* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int y = threadIdx.y; y < 32; y += blockDim.y) {
    const int r = r0 + y, c = c0 + threadIdx.x;
    tile[y][threadIdx.x] = (r < rows && c < cols) ? in[(int64_t)r * ld + c] : 0.f;
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += blockDim.y) {
    const int c = c0 + y, r = r0 + threadIdx.x;  // output row = input column
    if (c < cols && r < rows) {
      const float v = tile[threadIdx.x][y];
      const float h = tf32_rna(v);
      hi[(int64_t)c * rows + r] = h;
      lo[(int64_t)c * rows + r] = v - h;
    }
  }
}

}  // namespace tc

// split-K reduction + epilogue shared with the SIMT arm (defined in gemm_simt.cu)
int splitk_finish(const float* partial, const float* partial_cs, int splits, float* C, int ldc, int M, int N,
                  const float* bias, const float* add_src, int ld_add, float* act_out, int ld_act, int act_lo, int act_hi,
                  float* colsum, cudaStream_t st);

// K-major operand P[rows][k] (ld floats): box = 32 k x box_rows rows
static bool make_map_kmajor(CUtensorMap* m, const float* P, int64_t ld, int64_t rows, int k, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(P), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// MN-major operand P[r][mn] (ld floats), reduction over rows r: viewed as [mn/32][r][32], box = 32 x 32 r x box_mn/32
static bool make_map_mnmajor(CUtensorMap* m, const float* P, int ld, int r, int mn, int box_mn) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[3] = {32, (cuuint64_t)r, (cuuint64_t)(mn / 32)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, 128};
  cuuint32_t box[3] = {32, 32, (cuuint32_t)(box_mn / 32)};
  cuuint32_t es[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(P), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}


struct TcPlan {
  bool ok;
  bool mn_major;
  int block_n, n_mt, n_nt, splits, kb_total, kb_per_split;
  int ncta;              // 1, or 2 = CTA pair (tcgen05 cta_group::2, M = 256 per MMA)
  int stages;
  int64_t b_elems;       // elements of each pre-split B copy
  int64_t ws_bytes;
};

static TcPlan tc_plan(int M, int N, int K, int trans_a, int trans_b) {
  TcPlan t{};
  t.ok = false;
  if (M <= 0 || N <= 0 || K <= 0) return t;
  // (ta,tb) = (0,1) forward, (0,0) data gradient [B transposed during the split], (1,0) weight gradient
  if (trans_a && trans_b) return t;
  t.mn_major = trans_a != 0;
  if (t.mn_major && (M % 32 != 0 || N % 32 != 0)) return t;
  if (N < 16) return t;
  t.block_n = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  // CTA pairs halve the B bytes each SM stages and feeds to its tensor core per FLOP (the kernel is bound by
  // shared-memory bandwidth, profiles/r1d): used whenever there are enough 256-row tiles to fill the pairs.
  // GOTEN_GEMM_NCTA=1 forces single-CTA MMAs.
  static int force_ncta = -1;
  if (force_ncta < 0) { const char* e = getenv("GOTEN_GEMM_NCTA"); force_ncta = e ? atoi(e) : 0; }
  t.ncta = (force_ncta == 1) ? 1 : ((M >= 256 || force_ncta == 2) ? 2 : 1);
  const size_t stage_bytes = 2 * (size_t)tc::BM * tc::BK * 4 + 2 * (size_t)(t.block_n / t.ncta) * tc::BK * 4;
  t.stages = (int)((232448 - 4 * 2 * 4096 - 1024 - 256) / stage_bytes);  // 227 KB opt-in limit
  if (t.stages > tc::MAX_STAGES) t.stages = tc::MAX_STAGES;
  t.n_mt = (M + tc::BM * t.ncta - 1) / (tc::BM * t.ncta);
  t.n_nt = (N + t.block_n - 1) / t.block_n;
  t.kb_total = (K + tc::BK - 1) / tc::BK;
  t.splits = 1;
  if (t.mn_major) {
    // The tensor core accumulates with truncation: the relative error of a chain of n accumulating MMAs grows
    // like ~2e-8 n (measured: 1.7e-6 at K=256, 1.3e-5 at K=1792, 2e-4 at K=30k).  Weight-gradient GEMMs reduce
    // over 1e5..1e6 rows, so the reduction is cut into chains of <= MAX_CHAIN_KB k-blocks whose fp32 partials are
    // summed with round-to-nearest by the split-K reduction kernel.
    constexpr int MAX_CHAIN_KB = 48;
    const int tiles = t.n_mt * t.n_nt;
    int want = (148 / t.ncta) / tiles;  // at least one wave of (tile, split) work items
    if (want < 1) want = 1;
    int per = (t.kb_total + want - 1) / want;
    if (per > MAX_CHAIN_KB) per = MAX_CHAIN_KB;
    if (per < 8) per = t.kb_total < 8 ? t.kb_total : 8;
    t.splits = (t.kb_total + per - 1) / per;
  }
  t.kb_per_split = (t.kb_total + t.splits - 1) / t.splits;
  t.splits = (t.kb_total + t.kb_per_split - 1) / t.kb_per_split;  // no empty split
  t.b_elems = t.mn_major ? (int64_t)K * N : (int64_t)N * K;
  t.ws_bytes = 2 * align256(t.b_elems * 4);
  if (t.splits > 1) t.ws_bytes += align256((int64_t)t.splits * ((int64_t)M * N + M) * 4);
  t.ok = true;
  return t;
}

int64_t gemm_tc_workspace_bytes(int M, int N, int K, int trans_a, int trans_b) {
  TcPlan t = tc_plan(M, N, K, trans_a, trans_b);
  return t.ok ? t.ws_bytes : 0;
}

int gemm_tc(const float* A, int lda, int trans_a, const float* B, int ldb, int trans_b, float* C, int ldc, int M, int N,
            int K, const float* bias, const float* add_src, int ld_add, float* act_out, int ld_act, int act_lo,
            int act_hi, float* colsum, void* workspace, int64_t workspace_bytes, cudaStream_t st, bool* handled) {
  *handled = false;
  TcPlan t = tc_plan(M, N, K, trans_a, trans_b);
  if (!t.ok || workspace == nullptr || workspace_bytes < t.ws_bytes) return 0;
  if (!aligned16(A) || !aligned16(B) || lda % 4 != 0 || ldb % 4 != 0) return 0;
  if (colsum && !t.mn_major) return 0;
  if (get_encode() == nullptr) return 0;
  static int sm_count = 0, smem_optin = 0;
  if (sm_count == 0) {
    int dev = 0;
    GOTEN_CHECK_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    GOTEN_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return 0;  // tcgen05 needs sm_100
    sm_count = prop.multiProcessorCount;
    smem_optin = (int)prop.sharedMemPerBlockOptin;
  }

  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  float* Bh = reinterpret_cast<float*>(ws);
  float* Bl = reinterpret_cast<float*>(ws + align256(t.b_elems * 4));
  float* partial = nullptr;
  float* partial_cs = nullptr;
  if (t.splits > 1) {
    partial = reinterpret_cast<float*>(ws + 2 * align256(t.b_elems * 4));
    partial_cs = partial + (size_t)t.splits * M * N;
  }

  // ---- B operand: pre-split (and transpose for the data-gradient form) into the workspace
  int b_ld;
  {
    int rows, cols, transpose;
    if (t.mn_major) { rows = K; cols = N; transpose = 0; b_ld = N; }            // B[K][N] -> hi/lo [K][N]
    else if (trans_b) { rows = N; cols = K; transpose = 0; b_ld = K; }          // B[N][K] -> hi/lo [N][K]
    else { rows = K; cols = N; transpose = 1; b_ld = K; }                        // B[K][N] -> hi/lo [N][K]
    const int64_t tot = (int64_t)rows * cols;
    if (transpose) {
      dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32)), block(32, 8);
      tc::split_tf32_transpose_kernel<<<grid, block, 0, st>>>(B, ldb, rows, cols, Bh, Bl);
    } else if (cols % 4 == 0 && ldb % 4 == 0) {  // B is 16 B aligned (checked above)
      const int64_t tot4 = tot / 4;
      const unsigned grid = (unsigned)(cdiv64(tot4, 256) < (int64_t)sm_count * 16 ? cdiv64(tot4, 256) : (int64_t)sm_count * 16);
      tc::split_tf32_vec4_kernel<<<grid, 256, 0, st>>>(B, ldb, rows, cols / 4, Bh, Bl);
    } else {
      tc::split_tf32_kernel<<<(unsigned)cdiv64(tot, 256), 256, 0, st>>>(B, ldb, rows, cols, Bh, Bl);
    }
    GOTEN_CHECK_LAUNCH();
  }

  CUtensorMap mA, mBh, mBl, mC, mAct;
  bool ok;
  if (!t.mn_major) {
    ok = make_map_kmajor(&mA, A, lda, M, K, tc::BM) && make_map_kmajor(&mBh, Bh, b_ld, N, K, t.block_n / t.ncta) &&
         make_map_kmajor(&mBl, Bl, b_ld, N, K, t.block_n / t.ncta);
  } else {
    ok = make_map_mnmajor(&mA, A, lda, K, M, tc::BM) && make_map_mnmajor(&mBh, Bh, b_ld, K, N, t.block_n / t.ncta) &&
         make_map_mnmajor(&mBl, Bl, b_ld, K, N, t.block_n / t.ncta);
  }
  GOTEN_REQUIRE(ok, "cuTensorMapEncodeTiled failed (M=%d N=%d K=%d lda=%d)", M, N, K, lda);
  // output map for the TMA-store epilogue: 32 x 32 blocks of C (or of the split-K partial buffer)
  const bool red_add = add_src != nullptr && add_src == C && ld_add == ldc && act_out == nullptr && t.splits == 1;
  const bool act_tma = act_out != nullptr && t.splits == 1 && act_lo % 32 == 0 && act_hi > act_lo && aligned16(act_out) &&
                       ld_act % 4 == 0 && (add_src == nullptr || red_add);
  const bool fast_epi = t.splits > 1 || ((add_src == nullptr || red_add) && (act_out == nullptr || act_tma));
  if (t.splits > 1) ok = make_map_kmajor(&mC, partial, N, (int64_t)t.splits * M, N, 32);
  else if (fast_epi) {
    if (!aligned16(C) || ldc % 4 != 0) return 0;
    ok = make_map_kmajor(&mC, C, ldc, M, N, 32);
  } else mC = mA;  // unused by the fused epilogue path
  GOTEN_REQUIRE(ok, "cuTensorMapEncodeTiled failed for the output (M=%d N=%d ldc=%d)", M, N, ldc);
  mAct = mC;
  if (act_tma && t.splits == 1) {
    ok = make_map_kmajor(&mAct, act_out, ld_act, M, act_hi - act_lo, 32);
    GOTEN_REQUIRE(ok, "cuTensorMapEncodeTiled failed for the activation output (M=%d cols=%d ld=%d)", M, act_hi - act_lo, ld_act);
  }

  tc::Params p{};
  p.M = M; p.N = N; p.K = K;
  p.block_n = t.block_n; p.n_mt = t.n_mt; p.n_nt = t.n_nt;
  p.splits = t.splits; p.kb_per_split = t.kb_per_split; p.kb_total = t.kb_total;
  p.stages = t.stages;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("GOTEN_GEMM_DBG"); dbg = e ? atoi(e) : 0; }
    p.dbg = dbg;
  }
  p.C = C; p.ldc = ldc; p.bias = bias; p.add_src = add_src; p.ld_add = ld_add;
  p.act_out = act_out; p.ld_act = ld_act; p.act_lo = act_lo; p.act_hi = act_hi;
  p.red_add = red_add ? 1 : 0;
  p.act_tma = act_tma ? 1 : 0;
  p.add_vec = (add_src != nullptr && aligned16(add_src) && ld_add % 4 == 0) ? 1 : 0;
  p.c_vec = (aligned16(C) && ldc % 4 == 0) ? 1 : 0;
  p.act_vec = (act_out != nullptr && aligned16(act_out) && ld_act % 4 == 0 && act_lo % 4 == 0) ? 1 : 0;
  p.partial = partial; p.colsum = colsum; p.partial_colsum = partial_cs;
  if (t.splits > 1) { p.bias = nullptr; p.add_src = nullptr; p.act_out = nullptr; }  // applied by splitk_finish

  const size_t smem = 1024 + (size_t)t.stages * (2 * tc::BM * tc::BK * 4 + 2 * (size_t)(t.block_n / t.ncta) * tc::BK * 4) +
                      4 * 2 * 4096 + (3 * tc::MAX_STAGES + 4) * 8 + 16;
  GOTEN_REQUIRE((int)smem <= smem_optin, "tcgen05 GEMM needs %zu B of shared memory", smem);
  const int n_items = t.n_mt * t.n_nt * t.splits;
  const int max_units = sm_count / t.ncta;
  const int grid = (n_items < max_units ? n_items : max_units) * t.ncta;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(tc::NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)t.ncta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#define GOTEN_TC_LAUNCH(MN, NC)                                                                             \
  do {                                                                                                      \
    auto k = tc::gemm3x_kernel<MN, NC>;                                                                     \
    static int smem_set = 0; /* per instantiation: the attribute call takes a context lock, do it once */   \
    if ((int)smem > smem_set) {                                                                             \
      GOTEN_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin));   \
      smem_set = smem_optin;                                                                                \
    }                                                                                                       \
    GOTEN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k, mA, mBh, mBl, mC, mAct, p));                                     \
  } while (0)
  if (t.mn_major) {
    if (t.ncta == 2) GOTEN_TC_LAUNCH(true, 2); else GOTEN_TC_LAUNCH(true, 1);
  } else {
    if (t.ncta == 2) GOTEN_TC_LAUNCH(false, 2); else GOTEN_TC_LAUNCH(false, 1);
  }
#undef GOTEN_TC_LAUNCH
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
