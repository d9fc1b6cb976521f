// Fused edge kernel of the GATA block (forward): edge projections + attention + messages + aggregation in ONE kernel.
// Reference: representation/gotennet.py:406-407 (W_re, W_rs), :611 (gamma_t's linear part), :452-559 (message),
// :503 (PyG softmax), :613-640 (aggregate), :426-427 (residual).
//
// Until now a layer's forward ran  t @ [W_re | W_rs | gamma_t]^T  as a tcgen05 GEMM that wrote Ze [E][(S+2)C] to HBM
// (2.2 GB per layer at cfg2), then an attention kernel and a message kernel that read it back.  Here the projection
// tile never leaves the SM:
//
//   * TRANSPOSED product.  For a tile of <= 128 consecutive edges (whole targets: edges are sorted by target) and a
//     chunk of 128 projection columns the tensor core computes  D^T = We_chunk (128 x K) . t_tile^T (K x 128)  - the
//     same K-major 128B-swizzled fp16 hi / lo tiles as gemm_tc16.cu, only the two shared-memory descriptors are
//     swapped in the tcgen05.mma - so a TMEM LANE is a projection column (channel) and a TMEM COLUMN is an edge.
//   * The epilogue thread that owns a lane therefore owns one channel for all edges of the tile and walks them in
//     order: attention logits (W_re chunks; one warp = one head, reduced with shuffles) -> segment softmax per target
//     in shared memory -> messages (W_rs chunks): o = z x_j fc + alpha~ v_j, accumulated per target in registers and
//     flushed at the target's last edge.  The per-node scatter-sum is a running sum in one thread: no atomics, no
//     cross-thread reduction, bit-reproducible.  Neighbour values x_j[c], v_j[c], X_j[m][c] are coalesced 128 B warp
//     loads that hit L1/L2 (a tile's sources are a handful of atoms of one or two molecules).
//   * Ze is stored (coalesced, 512 B per edge and chunk) only from column `store_from` on: everything for training
//     (the backward kernels read the pre-activations), only gamma_t's columns for inference.
//
// Pipeline = gemm_tc16.cu's: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue, warps 6-9 converter (raw
// fp32 t tile -> fp16 hi / lo in place), 3-stage ring, double-buffered TMEM accumulator (2 x 128 columns).
// Contract (else the caller runs GEMM + goten_gata_fwd): head width C / H == 32, C % 128 == 0, K = C <= 256 ... any
// multiple of 64, in-degree <= 96, 16 B aligned rows.
#include <cuda_fp16.h>

#include "umma.cuh"

namespace goten {

namespace fused {

using namespace tc;

constexpr int BM = 128;          // edges per tile (TMEM columns)
constexpr int BNC = 128;         // projection columns per chunk (TMEM lanes)
constexpr int BK = 64;
constexpr int STAGES = 3;
constexpr int NTHREADS = 320;
constexpr int EPI_WARP0 = 2, CONV_WARP0 = 6;
constexpr uint32_t A_BYTES = BM * BK * 4;        // raw fp32 t tile of one k-block = its hi + lo fp16 tiles (32 KB)
constexpr uint32_t B_BYTES = BNC * BK * 2;       // one fp16 weight tile (hi or lo), 16 KB
constexpr uint32_t STAGE_BYTES = A_BYTES + 2 * B_BYTES;
constexpr int LMAXL = 15;

struct Params {
  int E, N, C, H, K, ldz, n_chunks, n_tiles, W;   // W: tile k starts at the first target whose edges begin at >= k*W
  int lmax, S, ND, NT, sep_dir, sep_tensor;
  int store_from;                                  // Ze columns >= store_from are written
  const float* bias;                               // [ldz]
  const float* amax_t;
  const float* amax_w;
  const float* h; const float* Xd; const float* qk; int ldqk; const float* x; const float* v;
  const float* Y; const float* fc; const float* kappa; const float* drop;
  const int32_t* tgt_ptr; const int32_t* src;
  float* Ze; float* alpha; float* h_out; float* Xd_out; float* xd_amax;
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 2, 128;" ::: "memory"); }

// first node a in [0, N] with tgt_ptr[a] >= val
__device__ __forceinline__ int lower_node(const int32_t* __restrict__ tgt_ptr, int N, int val) {
  int lo = 0, hi = N;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (tgt_ptr[mid] < val) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(NTHREADS, 1)
edge_fused_fwd_kernel(const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmWh,
                      const __grid_constant__ CUtensorMap tmWl, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // epilogue scalars of the current tile: STATIC shared arrays, so the compiler keeps them in the shared state space
  // (pointers carved out of the dynamic buffer degrade to generic LD.E with 64-bit address arithmetic: measured 32 M
  // generic loads and a 7-cycle issue interval in the epilogue warps)
  __shared__ int s_src[BM];                 // source node
  __shared__ int s_tgt[BM];                 // target node
  __shared__ int s_end[BM];                 // 1 = last edge of its target
  __shared__ float s_fc[BM];
  __shared__ float s_kap[BM];
  __shared__ float s_Y[BM * LMAXL];         // [128][L]
  __shared__ float s_al[BM * 16];           // [128][H] logits -> alpha~ (H <= 16)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);
  const uint32_t bar_full = smem_u32(bars), bar_conv = smem_u32(bars + STAGES), bar_empty = smem_u32(bars + 2 * STAGES);
  const uint32_t bar_tfull = smem_u32(bars + 3 * STAGES), bar_tempty = smem_u32(bars + 3 * STAGES + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb_total = p.K / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_conv + 8 * s, 4);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int a0 = lower_node(p.tgt_ptr, p.N, tile * p.W);
        const int e0 = p.tgt_ptr[a0];
        for (int ch = 0; ch < p.n_chunks; ++ch) {
          for (int kb = 0; kb < kb_total; ++kb, ++it) {
            const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(bar_empty + 8 * s, ph ^ 1);
            const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
            const uint32_t sbh = sa + A_BYTES, sbl = sbh + B_BYTES;
            const uint32_t bar = bar_full + 8 * s;
            mbar_arrive_expect_tx(bar, A_BYTES + 2 * B_BYTES);
            tma_load_2d(sa, &tmT, bar, kb * BK, e0);                  // t rows e0 .. e0+127: two [128][32 float] boxes
            tma_load_2d(sa + A_BYTES / 2, &tmT, bar, kb * BK + 32, e0);
            tma_load_2d(sbh, &tmWh, bar, kb * BK, ch * BNC);          // weight rows of the chunk, hi / lo
            tma_load_2d(sbl, &tmWl, bar, kb * BK, ch * BNC);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    if (lane == 0) {
      // c = f32 (bit 4), a = b = f16, both K-major; N (edges) >> 3 at bit 17, M (projection columns) >> 4 at bit 24
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BM >> 3) << 17) | ((uint32_t)(BNC >> 4) << 24);
      uint32_t it = 0, acc_it = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int ch = 0; ch < p.n_chunks; ++ch, ++acc_it) {
          const uint32_t acc = acc_it & 1, aph = (acc_it >> 1) & 1;
          mbar_wait(bar_tempty + 8 * acc, aph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * (uint32_t)BM;
          for (int kb = 0; kb < kb_total; ++kb, ++it) {
            const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(bar_full + 8 * s, ph);
            mbar_wait(bar_conv + 8 * s, ph);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
            const uint32_t sal = sa + A_BYTES / 2, sbh = sa + A_BYTES, sbl = sbh + B_BYTES;
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
              const uint64_t t_hi = make_desc(sa + kk * 32, 16, 1024, 2), t_lo = make_desc(sal + kk * 32, 16, 1024, 2);
              const uint64_t w_hi = make_desc(sbh + kk * 32, 16, 1024, 2), w_lo = make_desc(sbl + kk * 32, 16, 1024, 2);
              // D^T = W t^T: the weight tile is the M-side operand, the edge tile the N-side one
              umma_f16(d_tmem, w_lo, t_hi, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
              umma_f16(d_tmem, w_hi, t_lo, idesc, 1u);
              umma_f16(d_tmem, w_hi, t_hi, idesc, 1u);
            }
            umma_commit(bar_empty + 8 * s);
          }
          umma_commit(bar_tfull + 8 * acc);
        }
      }
    }
  } else if (warp >= CONV_WARP0) {
    // =============================== converter ==================================
    const int ct = threadIdx.x - CONV_WARP0 * 32;  // tile row (edge) this thread converts
    const float sA = scale_of(*p.amax_t);
    const uint32_t sw = (uint32_t)(ct & 7);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      for (int ch = 0; ch < p.n_chunks; ++ch) {
        for (int kb = 0; kb < kb_total; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(bar_full + 8 * s, ph);
          uint8_t* a_raw = smem + s * STAGE_BYTES;
          uint8_t* hi_row = a_raw + ct * 128;
          uint8_t* lo_row = hi_row + A_BYTES / 2;
          float4 vv[16];
#pragma unroll
          for (int c = 0; c < 16; ++c)
            vv[c] = *reinterpret_cast<const float4*>(a_raw + (c >> 3) * (A_BYTES / 2) + ct * 128 + ((((uint32_t)c & 7) ^ sw) << 4));
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 hh, ll;
            split2(vv[2 * c].x * sA, vv[2 * c].y * sA, hh.x, ll.x);
            split2(vv[2 * c].z * sA, vv[2 * c].w * sA, hh.y, ll.y);
            split2(vv[2 * c + 1].x * sA, vv[2 * c + 1].y * sA, hh.z, ll.z);
            split2(vv[2 * c + 1].z * sA, vv[2 * c + 1].w * sA, hh.w, ll.w);
            const uint32_t off = (((uint32_t)c ^ sw) << 4);
            *reinterpret_cast<uint4*>(hi_row + off) = hh;
            *reinterpret_cast<uint4*>(lo_row + off) = ll;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_conv + 8 * s);
        }
      }
    }
  } else {
    // =============================== epilogue ===================================
    const int q = warp & 3;                 // TMEM lane quarter of this warp
    const int et = threadIdx.x - EPI_WARP0 * 32;  // 0..127
    const int tl = q * 32 + lane;           // TMEM lane = column of the chunk
    const int C = p.C, H = p.H, L = (p.lmax + 1) * (p.lmax + 1) - 1, SC = p.S * C;
    const int SD_ = p.S * (C / H);          // value columns per head (gotennet.py:516-519)
    const float un = inv_scale_of(*p.amax_t) * inv_scale_of(*p.amax_w);
    const float* __restrict__ qkp = p.qk;
    const float* __restrict__ xp = p.x;
    const float* __restrict__ vp = p.v;
    const float* __restrict__ Xp = p.Xd;
    const float* __restrict__ hp = p.h;
    float xamx = 0.f;
    uint32_t acc_it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int a0 = lower_node(p.tgt_ptr, p.N, tile * p.W);
      const int a1 = (tile + 1 == p.n_tiles) ? p.N : lower_node(p.tgt_ptr, p.N, (tile + 1) * p.W);
      const int e0 = p.tgt_ptr[a0];
      const int ne = p.tgt_ptr[a1] - e0;    // <= 128
      // ---- per-edge scalars of the tile
      epi_bar();                             // previous tile fully consumed
      if (et < ne) {
        const int e = e0 + et;
        s_src[et] = p.src[e];
        s_fc[et] = p.fc[e];
        s_kap[et] = p.kappa[e];
        for (int m = 0; m < L; ++m) s_Y[et * LMAXL + m] = p.Y[(size_t)e * L + m];
      }
      for (int a = a0; a < a1; ++a) {             // (rare) targets without edges: outputs = inputs
        if (p.tgt_ptr[a + 1] != p.tgt_ptr[a]) continue;
        for (int c = et; c < C; c += 128) {
          p.h_out[(size_t)a * C + c] = p.h[(size_t)a * C + c];
          for (int m = 0; m < L; ++m) p.Xd_out[((size_t)m * p.N + a) * C + c] = p.Xd[((size_t)m * p.N + a) * C + c];
        }
      }
      for (int a = a0 + et; a < a1; a += 128) {   // targets of the tile mark their edges
        const int b0 = p.tgt_ptr[a] - e0, b1 = p.tgt_ptr[a + 1] - e0;
        for (int e = b0; e < b1; ++e) { s_tgt[e] = a; s_end[e] = (e == b1 - 1) ? 1 : 0; }
      }
      epi_bar();
      for (int ch = 0; ch < p.n_chunks; ++ch, ++acc_it) {
        const uint32_t acc = acc_it & 1, aph = (acc_it >> 1) & 1;
        const int col = ch * BNC + tl;       // projection column of this thread
        const float bcol = p.bias ? p.bias[col] : 0.f;
        const bool store = (ch * BNC >= p.store_from);
        // chunk kind
        const int kind = (ch * BNC < C) ? 0 : ((ch * BNC < C + SC) ? 1 : 2);   // 0 attention, 1 message, 2 store only
        const int scol = ch * BNC - C;       // message chunks: column inside [0, S*C)
        const int sidx = kind == 1 ? scol / C : 0;
        const int cc = kind == 1 ? (scol - sidx * C) + tl : (kind == 0 ? col : 0);   // channel in [0, C)
        // message chunk role: 0 scalar, 1 direction (degrees [mlo, mhi)), 2 tensor
        int role = 0, mlo = 0, mhi = 0;
        if (kind == 1 && sidx >= 1) {
          if (sidx <= p.ND) { role = 1; const int l = p.sep_dir ? sidx : 0; mlo = p.sep_dir ? l * l - 1 : 0; mhi = p.sep_dir ? (l + 1) * (l + 1) - 1 : L; }
          else { role = 2; const int l = p.sep_tensor ? sidx - p.ND : 0; mlo = p.sep_tensor ? l * l - 1 : 0; mhi = p.sep_tensor ? (l + 1) * (l + 1) - 1 : L; }
        }
        const int hd_v = kind == 1 ? (sidx * C + cc) / SD_ : 0;
        // per-chunk bases (node-row offsets fit 32 bits: N * S * C < 2^31 is checked on the host)
        const float* __restrict__ xc = xp + sidx * C + cc;
        const float* __restrict__ vc = vp + sidx * C + cc;
        const float* __restrict__ qc = qkp + cc;
        mbar_wait(bar_tfull + 8 * acc, aph);
        tc_fence_after();
        float acc_s = 0.f, accm[7];
#pragma unroll
        for (int m = 0; m < 7; ++m) accm[m] = 0.f;
        for (int g = 0; g < BM / 32; ++g) {
          if (g * 32 >= ne) break;           // CTA-uniform
          const uint32_t taddr = tmem_base + acc * (uint32_t)BM + g * 32 + ((uint32_t)(q * 32) << 16);
          uint32_t r[32];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          // All neighbour values of the 32 edges are requested before the first one is used (one L1/L2 round trip per
          // group instead of one per edge); edges past the tile's end read node 0 and are masked out below.
          if (store) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const int el = g * 32 + k;
              if (el < ne) p.Ze[(size_t)(e0 + el) * p.ldz + col] = __uint_as_float(r[k]) * un + bcol;
            }
          }
          if (kind == 0) {
            // attention logit partial of this channel; the warp's 32 channels are one head (C / H == 32)
            float kq[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const int el = g * 32 + k;
              const bool ok = el < ne;
              const int i = ok ? s_tgt[el] : 0, j = ok ? s_src[el] : 0;
              kq[k] = __ldg(qc + i * p.ldqk) * __ldg(qc + C + j * p.ldqk);
            }
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const int el = g * 32 + k;
              if (el >= ne) break;           // CTA-uniform
              float pr = kq[k] * siluf_(__uint_as_float(r[k]) * un + bcol);
              pr = warp_sum(pr);
              if (lane == 0) s_al[el * 16 + (cc >> 5)] = pr;
            }
          } else if (kind == 1) {
            float o[32];
            {
              float xv[32], vv[32];
#pragma unroll
              for (int k = 0; k < 32; ++k) {
                const int el = g * 32 + k;
                const int j = el < ne ? s_src[el] : 0;
                xv[k] = __ldg(xc + j * SC);
                vv[k] = __ldg(vc + j * SC);
              }
#pragma unroll
              for (int k = 0; k < 32; ++k) {
                const int el = g * 32 + k;
                const bool ok = el < ne;
                const float z = __uint_as_float(r[k]) * un + bcol;
                o[k] = ok ? z * xv[k] * s_fc[el] + s_al[el * 16 + hd_v] * vv[k] : 0.f;
              }
            }
            if (role == 0) {
#pragma unroll
              for (int k = 0; k < 32; ++k) {
                const int el = g * 32 + k;
                if (el >= ne) break;         // CTA-uniform
                acc_s += o[k];
                if (s_end[el]) {             // CTA-uniform: flush the target's sum
                  const int i = s_tgt[el];
                  p.h_out[(size_t)i * C + cc] = __ldg(hp + (size_t)i * C + cc) + acc_s;
                  acc_s = 0.f;
                }
              }
            } else {
              // one component at a time: Y (direction) or the gathered X_j (tensor) of all 32 edges, then the running sum
#pragma unroll 1
              for (int m = mlo; m < mhi; ++m) {
                const float* __restrict__ Xm = Xp + (size_t)m * p.N * C + cc;
                float f[32];
                if (role == 1) {
#pragma unroll
                  for (int k = 0; k < 32; ++k) { const int el = g * 32 + k; f[k] = el < ne ? s_Y[el * LMAXL + m] : 0.f; }
                } else {
#pragma unroll
                  for (int k = 0; k < 32; ++k) {
                    const int el = g * 32 + k;
                    const int j = el < ne ? s_src[el] : 0;
                    f[k] = __ldg(Xm + j * C);
                  }
                }
                float a_m = accm[m - mlo];
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                  const int el = g * 32 + k;
                  if (el >= ne) break;       // CTA-uniform
                  a_m = fmaf(f[k], o[k], a_m);
                  if (s_end[el]) {           // CTA-uniform: flush component m of this target
                    const int i = s_tgt[el];
                    const size_t o_ = ((size_t)m * p.N + i) * C + cc;
                    // direction chunks write X + dX first, the tensor chunks (later chunk index, same thread) add to it
                    const float base = (role == 1) ? __ldg(Xp + o_) : p.Xd_out[o_];
                    const float xo = base + a_m;
                    p.Xd_out[o_] = xo;
                    if (role == 2) xamx = fmaxf(xamx, fabsf(xo));
                    a_m = 0.f;
                  }
                }
                accm[m - mlo] = a_m;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
        // ---- after the last attention chunk: segment softmax per (target, head) of the tile
        if (kind == 0 && (ch + 1) * BNC >= C) {
          epi_bar();
          const int nt = a1 - a0;
          for (int pr = et; pr < nt * H; pr += 128) {
            const int a = a0 + pr / H, hd = pr % H;
            const int b0 = p.tgt_ptr[a] - e0, b1 = p.tgt_ptr[a + 1] - e0;
            float mx = -INFINITY;
            for (int e = b0; e < b1; ++e) mx = fmaxf(mx, s_al[e * 16 + hd]);
            float sum = 0.f;
            for (int e = b0; e < b1; ++e) sum += expf(s_al[e * 16 + hd] - mx);
            const float den = sum + 1e-16f;  // PyG softmax epsilon
            for (int e = b0; e < b1; ++e) {
              const float al = expf(s_al[e * 16 + hd] - mx) / den;
              p.alpha[(size_t)(e0 + e) * H + hd] = al;
              float at = al * s_kap[e];
              if (p.drop) at *= p.drop[(size_t)(e0 + e) * H + hd];   // attention dropout (gotennet.py:513)
              s_al[e * 16 + hd] = at;
            }
          }
          epi_bar();
        }
      }
    }
    amax_commit(p.xd_amax, xamx);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base));
  }
}

// weights [rows][K] fp32 -> scaled fp16 hi / lo copies [rows][K]
__global__ void presplit_rows_kernel(const float* __restrict__ W, int rows, int K, const float* __restrict__ amax,
                                     __half* __restrict__ hi, __half* __restrict__ lo) {
  const float s = scale_of(*amax);
  const int64_t total = (int64_t)rows * (K >> 2);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(W)[idx];
    uint2 h, l;
    split2(v.x * s, v.y * s, h.x, l.x);
    split2(v.z * s, v.w * s, h.y, l.y);
    reinterpret_cast<uint2*>(hi)[idx] = h;
    reinterpret_cast<uint2*>(lo)[idx] = l;
  }
}

static bool make_map(CUtensorMap* m, CUtensorMapDataType dt, const void* P, int64_t ld_bytes, int64_t rows, int64_t cols,
                     int box_cols, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  return enc(m, dt, 2, const_cast<void*>(P), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace fused
}  // namespace goten

using namespace goten;

extern "C" {

int64_t goten_gata_fused_workspace_bytes(int C, int ldz) { return 2 * align256((int64_t)ldz * C * 2) + 256; }

int goten_gata_fused_fwd(const float* t, const float* We, const float* be, const float* t_amax, const float* w_amax,
                         const float* h, const float* Xd, const float* qk, int ldqk, const float* x, const float* v,
                         const float* Y, const float* fc, const float* kappa, const float* drop,
                         const int32_t* tgt_ptr, const int32_t* src, int N, int64_t E, int C, int H, int lmax, int flags,
                         int max_deg_in, int ldz, int store_from, float* Ze, float* alpha, float* h_out, float* Xd_out,
                         float* xd_amax, void* workspace, int64_t workspace_bytes, int* handled, void* stream) {
  *handled = 0;
  // OPT-IN (GOTEN_EDGE_FUSED=1, read per call).  Parity-green, and Ze never reaches HBM in inference, but measured 5-9x
  // slower than GEMM + attention + message kernels on B200: the thread-per-channel epilogue executes ~100
  // instructions per (edge, 128-column chunk) with ONE warp per scheduler (14 % issue efficiency, ncu), see DESIGN.md.
  { const char* e = getenv("GOTEN_EDGE_FUSED"); if (!(e && atoi(e) == 1)) return 0; }
  if (N <= 0 || E <= 0 || E > 0x7fffffff) return 0;
  if ((int64_t)N * 15 * C >= 0x7fffffff || (int64_t)N * ldqk >= 0x7fffffff) return 0;   // 32-bit node-row offsets
  if (H <= 0 || C % H != 0 || C / H != 32 || H > 16) return 0;            // one epilogue warp = one head
  if (C % fused::BNC != 0 || C % fused::BK != 0 || ldz % fused::BNC != 0 || ldqk % 4 != 0) return 0;
  if (lmax < 1 || lmax > 3) return 0;
  if (lmax > 1 && (flags & 3) != 3) return 0;   // per-degree direction / tensor chunks (<= 7 components per chunk)
  if (max_deg_in < 1 || max_deg_in > 96) return 0;                        // a tile holds whole targets
  if (t_amax == nullptr || w_amax == nullptr || workspace == nullptr) return 0;
  if (workspace_bytes < goten_gata_fused_workspace_bytes(C, ldz)) return 0;
  if (!aligned16(t) || !aligned16(We) || !aligned16(Ze)) return 0;
  if (get_encode() == nullptr) return 0;
  static int sm_count = 0, smem_optin = 0, cc_major = 0;
  if (sm_count == 0) {
    int dev = 0;
    GOTEN_CHECK_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    GOTEN_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
    cc_major = prop.major;
    sm_count = prop.multiProcessorCount;
    smem_optin = (int)prop.sharedMemPerBlockOptin;
  }
  if (cc_major != 10) return 0;
  cudaStream_t st = as_stream(stream);
  const int sep_dir = (flags & 1) && lmax > 1, sep_tensor = (flags & 2) && lmax > 1;
  const int ND = sep_dir ? lmax : 1, NT = sep_tensor ? lmax : 1, S = 1 + ND + NT;
  if (ldz < (S + 1) * C) return 0;

  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const int64_t wbytes = align256((int64_t)ldz * C * 2);
  __half* Wh = reinterpret_cast<__half*>(ws);
  __half* Wl = reinterpret_cast<__half*>(ws + wbytes);
  {
    const int64_t work = (int64_t)ldz * (C / 4);
    int64_t grid = cdiv64(work, 256);
    if (grid > (int64_t)sm_count * 8) grid = (int64_t)sm_count * 8;
    fused::presplit_rows_kernel<<<(unsigned)grid, 256, 0, st>>>(We, ldz, C, w_amax, Wh, Wl);
    GOTEN_CHECK_LAUNCH();
  }
  CUtensorMap mT, mWh, mWl;
  bool ok = fused::make_map(&mT, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, t, (int64_t)C * 4, E, C, 32, fused::BM) &&
            fused::make_map(&mWh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, Wh, (int64_t)C * 2, ldz, C, 64, fused::BNC) &&
            fused::make_map(&mWl, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, Wl, (int64_t)C * 2, ldz, C, 64, fused::BNC);
  GOTEN_REQUIRE(ok, "cuTensorMapEncodeTiled failed (fused edge kernel, E=%lld C=%d ldz=%d)", (long long)E, C, ldz);

  fused::Params p{};
  p.E = (int)E; p.N = N; p.C = C; p.H = H; p.K = C; p.ldz = ldz; p.n_chunks = ldz / fused::BNC;
  p.W = fused::BM - (max_deg_in - 1);
  p.n_tiles = (int)((E + p.W - 1) / p.W);
  p.lmax = lmax; p.S = S; p.ND = ND; p.NT = NT; p.sep_dir = sep_dir; p.sep_tensor = sep_tensor;
  p.store_from = store_from;
  p.bias = be; p.amax_t = t_amax; p.amax_w = w_amax;
  p.h = h; p.Xd = Xd; p.qk = qk; p.ldqk = ldqk; p.x = x; p.v = v; p.Y = Y; p.fc = fc; p.kappa = kappa; p.drop = drop;
  p.tgt_ptr = tgt_ptr; p.src = src;
  p.Ze = Ze; p.alpha = alpha; p.h_out = h_out; p.Xd_out = Xd_out; p.xd_amax = xd_amax;

  const size_t smem = 1024 + (size_t)fused::STAGES * fused::STAGE_BYTES + (3 * fused::STAGES + 4) * 8 + 16;
  GOTEN_REQUIRE((int)smem + 18432 + 1024 <= smem_optin, "fused edge kernel needs %zu B of dynamic shared memory", smem);
  static int smem_set = 0;
  if ((int)smem > smem_set) {   // (18 KB of static shared memory ride beside the dynamic ring)
    GOTEN_CHECK_CUDA(cudaFuncSetAttribute(fused::edge_fused_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = (int)smem;
  }
  const int grid = p.n_tiles < sm_count ? p.n_tiles : sm_count;
  fused::edge_fused_fwd_kernel<<<grid, fused::NTHREADS, smem, st>>>(mT, mWh, mWl, p);
  GOTEN_CHECK_LAUNCH();
  *handled = 1;
  return 0;
}

}  // extern "C"
