/*
 * gotennet_b200 — C ABI of the B200-native GotenNet interaction path.
 *
 * The reference (sarpaykent/GotenNet) is 100 % Python: it has no FFI layer, its
 * boundary is the nn.Module API (representation/gotennet.py:366, :716, :956, :1026).
 * This header is therefore the boundary the reference WOULD bind if its hot path
 * were native: every entry point below names the reference statement(s) it
 * replaces (file:line relative to /root/reference/gotennet/models/).
 *
 * Conventions
 *   - plain C: raw DEVICE pointers, sizes, leading dimensions, a cudaStream_t
 *     passed as void*.  No torch types, no C++ exceptions across the ABI.
 *   - every function returns 0 on success, non-zero on failure;
 *     goten_last_error() returns a thread-local message.
 *   - all floating point is IEEE fp32; indices are int32 on the device side
 *     (edge_index is additionally emitted as int64 for the PyG-style API).
 *   - edges are stored sorted by (target, source).  `tgt_ptr[N+1]` is the CSR
 *     over targets; `src_ptr[N+1]` + `src_perm[E]` list edge ids grouped by
 *     source (ascending target) — the transposed view the backward needs.
 *   - steerable features are stored DEGREE-MAJOR: Xd[L][N][C] (the reference
 *     stores [N][L][C]); L = (lmax+1)^2-1.
 *   - per-layer edge GEMM output Ze[E][ldz]: columns [0,C) = pre-activation of
 *     W_re, [C,(S+1)C) = W_rs output ("filter"), [(S+1)C,(S+2)C) =
 *     pre-activation of gamma_t (non-last layers).
 *   - launches are asynchronous on `stream` unless stated otherwise.
 */
#ifndef GOTENNET_B200_H
#define GOTENNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GOTEN_ABI_VERSION 1

int goten_abi_version(void);
const char* goten_last_error(void);
/* number of CUDA kernels this library has launched so far in this process (bench.py's gpu_launches) */
int64_t goten_launch_count(void);
/* device properties the host uses to size launches: out[0]=SM count, out[1]=max smem/block optin, out[2]=cc major*10+minor */
int goten_device_info(int* out3);

/* ------------------------------------------------------------------ graph --
 * components/layers.py:1588-1590 (Distance.forward -> torch_cluster.radius_graph,
 * loop=True, max_num_neighbors=K, CUDA-build semantics: strict d^2 < r^2, same
 * molecule, first K sources in ascending index, grouped by target).
 *
 * Step 1 (count): molecule boundaries from the sorted `batch`, in-degree per
 * target, exclusive scan -> tgt_ptr.  SYNCHRONISES the stream once to return
 * E (the host must size the edge arrays) and the number of molecules.
 *   pos[N][3] f32, batch[N] i64 -> mol_ptr[N+1] i32 (first n_mol+1 entries valid),
 *   mol_of[N] i32, tgt_ptr[N+1] i32, *n_edges_out, *n_mol_out
 *   scratch: int32[ 2*N + 4096 ]                                              */
int goten_radius_graph_count(const float* pos, const int64_t* batch, int n_nodes, float cutoff,
                             int max_num_neighbors, int loop, int32_t* mol_ptr, int32_t* mol_of,
                             int32_t* tgt_ptr, int32_t* scratch, int64_t* n_edges_out,
                             int32_t* n_mol_out, void* stream);
/* Step 2 (fill): edge lists, int64 edge_index[2][E] (row0 = source, row1 = target),
 * out-degree per source, src_ptr (scan) and src_perm.  Bit-exact integers.      */
int goten_radius_graph_fill(const float* pos, const int32_t* mol_ptr, const int32_t* mol_of,
                            const int32_t* tgt_ptr, int n_nodes, int64_t n_edges, float cutoff,
                            int max_num_neighbors, int loop, int32_t* src, int32_t* tgt,
                            int64_t* edge_index, int32_t* deg_out, int32_t* src_ptr,
                            int32_t* src_perm, int32_t* scratch, void* stream);
/* Transposed view for an externally supplied, target-sorted edge list
 * (GotenNet.forward takes edge_index from the caller, gotennet.py:956):
 * deg_out, src_ptr, src_perm from src/tgt.  `order_by_src[E]` is a stable
 * argsort of src computed by the caller.                                        */
int goten_csr_from_sorted(const int32_t* src, const int32_t* tgt, const int32_t* order_by_src,
                          int n_nodes, int64_t n_edges, int32_t* tgt_ptr, int32_t* deg_out,
                          int32_t* src_ptr, int32_t* src_perm, int32_t* scratch, void* stream);

/* --------------------------------------------------------------- geometry --
 * layers.py:1591-1604 (edge_vec, edge_weight), gotennet.py:978-989 (unit vector,
 * out-degree gather), layers.py:805-869 (TensorInit, degrees 1..lmax<=3),
 * layers.py:744-746 + :149-152 (ExpNormalSmearing * CosineCutoff).
 *   -> r[E], u[E][3] (0 on self loops), Y[E][L], fc[E] (cosine cutoff),
 *      kappa[E] = (scale_edge ? sqrt(deg_out[src]) : 1) / sqrt(C)  (gotennet.py:506-511),
 *      phi[E][n_rbf]
 * If `edge_vec_in` != NULL the vectors are taken from it ([E][3], un-normalised,
 * GotenNet.forward's argument) instead of pos[src]-pos[tgt]; if `r_in` != NULL the
 * distances fed to the radial basis / cutoff are taken from it (edge_diff).
 * basis (layers.py:749-777): 0 = expnorm (means, betas; cosine cutoff folded in, :744-746),
 *   1 = BesselBasis sin(f r) / r (means = freqs; :349-358), 2 = GaussianRBF (means = offsets,
 *   betas = widths; :276-291).                                                     */
int goten_edge_geometry_fwd(const float* pos, const float* edge_vec_in, const float* r_in, const int32_t* src,
                            const int32_t* tgt, const int32_t* deg_out, int64_t n_edges, int lmax,
                            float cutoff, int n_rbf, int basis, const float* means, const float* betas,
                            int scale_edge, int n_atom_basis, float* r, float* u, float* Y,
                            float* fc, float* kappa, float* phi, void* stream);
/* d(loss)/d(edge_vec) from the gradients of phi, fc and Y (first-order forces,
 * outputs.py:365-375 needs d/dpos).  g_vec[E][3]; the caller scatters it to pos. */
int goten_edge_geometry_bwd(const float* r, const float* u, const int32_t* src, const int32_t* tgt,
                            int64_t n_edges, int lmax, float cutoff, int n_rbf, int basis,
                            const float* means, const float* betas, const float* g_phi, const float* g_fc,
                            const float* g_Y, const float* g_r, float* g_vec, void* stream);
/* scatter of per-edge vector gradients onto positions: g_pos[i] = sum_{e: src=i} g - sum_{e: tgt=i} g */
int goten_edge_vec_to_pos_bwd(const float* g_vec, const int32_t* tgt_ptr, const int32_t* src_ptr,
                              const int32_t* src_perm, int n_nodes, float* g_pos, void* stream);

/* ------------------------------------------------------------------- GEMM --
 * Every nn.Linear / Dense on the path (layers.py:523 -> F.linear), fp32 result
 * accuracy (the reference runs with TF32 disabled, scripts/train.py:16).
 *
 *   C[M][N] = opA(A) * opB(B) (+ bias[N])
 *   trans_a = 0: A is [M][K] (lda)      trans_a = 1: A is [K][M] (lda)
 *   trans_b = 0: B is [K][N] (ldb)      trans_b = 1: B is [N][K] (ldb)   (nn.Linear weight)
 * epilogue:
 *   act_out != NULL: additionally writes act_out[m][n - act_lo] = silu(C[m][n]) for
 *     act_lo <= n < act_hi (ld = ld_act); C keeps the pre-activation.
 *   add_src != NULL: C[m][n] += add_src[m][n] (ld = ld_add; may alias C) — residual /
 *     gradient-join fused into the GEMM.
 *   colsum != NULL (trans_a = 1 only): colsum[m] = sum_k A[k][m]  (bias gradient
 *     fused into the weight-gradient GEMM).
 * `workspace` (split-K partials): at least goten_gemm_workspace_bytes(...) bytes.
 * impl: 0 = auto (first arm that accepts the shape, in the order 3, 2, 1), 1 = fp32 SIMT,
 *       (a negative code -2 / -3 prefers that arm and falls back to 1 instead of failing)
 *       2 = tcgen05 3xTF32, 3 = tcgen05 split-fp16 (x*2^s = hi + lo in fp16, three
 *       kind::f16 MMAs, fp32 accumulation; 22 significant bits).
 * goten_gemm_scaled: same, with optional DEVICE pointers to an upper bound of max|A| /
 *   max|B| (within a few powers of two of the true maximum) used by arm 3 for its
 *   power-of-two operand scaling; NULL = measured inside the call by one read pass.     */
int64_t goten_gemm_workspace_bytes(int M, int N, int K, int trans_a, int trans_b);
int goten_gemm(const float* A, int lda, int trans_a, const float* B, int ldb, int trans_b,
               float* C, int ldc, int M, int N, int K, const float* bias, const float* add_src,
               int ld_add, float* act_out, int ld_act, int act_lo, int act_hi, float* colsum,
               void* workspace, int64_t workspace_bytes, int impl, void* stream);
int goten_gemm_scaled(const float* A, int lda, int trans_a, const float* B, int ldb, int trans_b,
                      float* C, int ldc, int M, int N, int K, const float* bias,
                      const float* add_src, int ld_add, float* act_out, int ld_act, int act_lo,
                      int act_hi, float* colsum, const float* a_amax, const float* b_amax,
                      void* workspace, int64_t workspace_bytes, int impl, void* stream);
/* Kernels that produce a large GEMM operand take an optional trailing `*_amax` DEVICE
 * pointer (t_amax, xd_amax, gze_amax, geq_amax, gek_amax, gp_amax, ctx_amax, gm_amax, out_amax): a running
 * max |value written| (atomic max on a non-negative float the caller zeroed), so the
 * operand needs no separate goten_absmax pass.  NULL disables it.
 * out[0] = max(out[0], max |A[m][n]|) over a [M][N] matrix (ld = lda); out must hold a
 * non-negative float (zero it first for a plain maximum).                                */
int goten_absmax(const float* A, int64_t lda, int64_t M, int N, float* out, void* stream);
/* The same for up to 16 CONTIGUOUS tensors in one launch (the weight matrices of a block):
 * out[i] = max(out[i], max |ptrs[i][0 .. numel[i])|); ptrs / numel are HOST arrays.          */
int goten_absmax_multi(const float* const* ptrs, const int64_t* numel, int count, float* out,
                       void* stream);
/* out[m][n] = g[m][n] * silu'(pre[m][n])  (Dense activation backward, layers.py:527-528) */
int goten_dsilu_mul(const float* g, int ldg, const float* pre, int ldp, float* out, int ldo,
                    int64_t M, int N, float* out_amax, void* stream);
/* column sums of a [M][N] matrix (bias gradients of non-fused cases) */
int goten_colsum(const float* A, int lda, int64_t M, int N, float* out, float* workspace,
                 int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------- init block --
 * F[E][2C] = phi * [W_ndp ; W_erp]^T + b is produced by goten_gemm.
 * NodeInit.message + aggregate (layers.py:1658-1675): self loops skipped,
 *   m[i][c] = sum_{e->i, j!=i} hnbr[j][c] * F[e][c] * fc[e]                      */
int goten_node_init_agg_fwd(const float* F, int ldf, const float* hnbr, const float* fc,
                            const int32_t* tgt_ptr, const int32_t* src, int n_nodes, int C,
                            float* m, void* stream);
/* gF[e][c] (cols [0,C) of gF) = g_m[i][c]*hnbr[j][c]*fc[e];  g_fc_acc[e] += sum_c g_m*hnbr*F (optional) */
int goten_node_init_agg_bwd_tgt(const float* g_m, const float* F, int ldf, const float* hnbr,
                                const float* fc, const int32_t* tgt_ptr, const int32_t* src,
                                int n_nodes, int C, float* gF, int ldgf, float* g_fc, void* stream);
/* g_hnbr[j][c] = sum_{e: src=j, tgt!=j} g_m[i][c]*F[e][c]*fc[e] */
int goten_node_init_agg_bwd_src(const float* g_m, const float* F, int ldf, const float* fc,
                                const int32_t* src_ptr, const int32_t* src_perm, const int32_t* tgt,
                                int n_nodes, int C, float* g_hnbr, void* stream);
/* Dense(2C->C)+LayerNorm then SiLU (layers.py:523-528 with norm='layer'):
 * y[n][c] = silu( LN(x[n][:]) * gamma + beta ), eps = 1e-5; saves mean/rstd.     */
int goten_ln_silu_fwd(const float* x, const float* gamma, const float* beta, int64_t n_rows, int C,
                      float eps, float* y, float* mean, float* rstd, void* stream);
int goten_ln_silu_bwd(const float* g_y, const float* x, const float* gamma, const float* beta,
                      const float* mean, const float* rstd, int64_t n_rows, int C, float* g_x,
                      float* g_gamma_part, float* g_beta_part, int n_part, void* stream);
/* Optional pre-norms of the GATA block (gotennet.py:306-315, :397-398).
 * goten_layernorm_*: plain nn.LayerNorm (same kernels as above without the SiLU).
 * goten_tensor_layernorm_*: TensorLayerNorm (layers.py:1497-1563) on Xd[L][N][C]: per degree
 *   and node, max-min normalisation of the channel norms; weight[C] (buffer, not trained). */
int goten_layernorm_fwd(const float* x, const float* gamma, const float* beta, int64_t n_rows, int C,
                        float eps, float* y, float* mean, float* rstd, void* stream);
int goten_layernorm_bwd(const float* g_y, const float* x, const float* gamma, const float* beta,
                        const float* mean, const float* rstd, int64_t n_rows, int C, float* g_x,
                        float* g_gamma_part, float* g_beta_part, int n_part, void* stream);
int goten_tensor_layernorm_fwd(const float* Xd, const float* weight, int n_nodes, int C, int lmax,
                               float* out, void* stream);
int goten_tensor_layernorm_bwd(const float* g_out, const float* Xd, const float* weight, int n_nodes,
                               int C, int lmax, float* g_X, void* stream);
/* EdgeInit.message (layers.py:1711): t[e][c] = (h[i][c] + h[j][c]) * F[e][C + c]  (self loops kept) */
int goten_edge_init_fwd(const float* h, const float* F, int ldf, int col0, const int32_t* src,
                        const int32_t* tgt, int64_t n_edges, int C, float* t, float* t_amax,
                        void* stream);
/* gF[e][col0+c] = g_t[e][c]*(h_i+h_j);  g_h[n][c] = sum_{e: tgt=n} g_t*F + sum_{e: src=n} g_t*F */
int goten_edge_init_bwd(const float* g_t, const float* h, const float* F, int ldf, int col0,
                        const int32_t* tgt_ptr, const int32_t* src, const int32_t* tgt,
                        const int32_t* src_ptr, const int32_t* src_perm, int n_nodes, int64_t n_edges,
                        int C, float* gF, int ldgf, float* g_h, void* stream);

/* ------------------------------------------------------------ GATA block --
 * GATA.message + softmax + aggregate + residual (gotennet.py:452-559, :503,
 * :613-640, :426-427), fused; one CTA per target node, neighbours read through
 * L2, segment softmax and scatter-sum done in-CTA (deterministic, no atomics).
 *   qk[N][ldqk]: cols [0,C) = q, [C,2C) = k      x[N][S*C], v[N][S*C]
 *   Xd[L][N][C]   Ze[E][ldz] (see top)           Y[E][L], fc[E], kappa[E]
 *   drop[E][H] (optional, NULL = none): attention dropout factors mask / (1 - p)
 *     (gotennet.py:513 F.dropout on the scaled attention weights); alpha[E][H] always
 *     holds the un-dropped softmax output
 * flags: bit0 = sep_dir, bit1 = sep_tensor.
 * outputs: h_out[N][C] = h + dh, Xd_out = Xd + dX, alpha[E][H] (normalised
 * attention weights, saved for the backward).                                  */
int goten_gata_fwd(const float* h, const float* Xd, const float* qk, int ldqk, const float* x,
                   const float* v, const float* Ze, int ldz, const float* Y, const float* fc,
                   const float* kappa, const float* drop, const int32_t* tgt_ptr, const int32_t* src,
                   int n_nodes,
                   int C, int H, int lmax, int flags, int max_deg_in, float* h_out, float* Xd_out,
                   float* alpha, float* xd_amax, void* stream);
/* Fused forward of the edge path (edge_fused.cu): the edge projections  t @ [W_re | W_rs | gamma_t]^T  (gotennet.py:
 * 406-407, :611), the attention softmax and the message aggregation of goten_gata_fwd in ONE kernel - the projection
 * tile stays in tensor memory (tcgen05, transposed product: TMEM lane = channel, column = edge) and is consumed by
 * the epilogue, so Ze is written only from column `store_from` on (0: everything, for the backward; (S+1)*C: only
 * gamma_t's pre-activations, for inference).  We [ldz][C], be [ldz]; t_amax / w_amax: device bounds of max|t| / max|We|.
 * *handled = 0 (and nothing launched) when the layout is outside the kernel's contract (head width 32, C % 128 == 0,
 * in-degree <= 96, per-degree chunks): the caller then runs goten_gemm_scaled + goten_gata_fwd.                       */
int64_t goten_gata_fused_workspace_bytes(int C, int ldz);
int goten_gata_fused_fwd(const float* t, const float* We, const float* be, const float* t_amax,
                         const float* w_amax, const float* h, const float* Xd, const float* qk, int ldqk,
                         const float* x, const float* v, const float* Y, const float* fc,
                         const float* kappa, const float* drop, const int32_t* tgt_ptr, const int32_t* src,
                         int n_nodes, int64_t n_edges, int C, int H, int lmax, int flags, int max_deg_in,
                         int ldz, int store_from, float* Ze, float* alpha, float* h_out, float* Xd_out,
                         float* xd_amax, void* workspace, int64_t workspace_bytes, int* handled,
                         void* stream);
/* backward, target-centric half: needs g_h[N][C], g_Xd[L][N][C] (gradients of the
 * block outputs).  Produces g_qk[:, 0:C) (dq), gZe[:, 0:(S+1)C) (d pre-act W_re, d filter),
 * da[E][H] (gradient of the attention logits) and, if non-NULL, the geometry
 * gradients g_fc[E] (accumulated) and g_Y[E][L] (accumulated).  gze_amax (optional,
 * device): running max |value written to gZe| (atomic max on a non-negative float; the
 * caller zeroes it), consumed by goten_gemm_scaled.                               */
int goten_gata_bwd_tgt(const float* g_h, const float* g_Xd, const float* Xd, const float* qk,
                       int ldqk, const float* x, const float* v, const float* Ze, int ldz,
                       const float* Y, const float* fc, const float* kappa, const float* drop,
                       const float* alpha,
                       const int32_t* tgt_ptr, const int32_t* src, int n_nodes, int C, int H,
                       int lmax, int flags, int max_deg_in, float* g_qk, int ldgqk, float* gZe,
                       int ldgz, float* da, float* g_fc, float* g_Y, float* gze_amax, void* stream);
/* backward, source-centric half: g_qk[:, C:2C) (dk), g_x, g_v [N][S*C] and
 * g_Xd_in[L][N][C] = g_Xd + sum over outgoing edges (residual included).
 * gx_amax / gv_amax (optional, device, zeroed by the caller): running max |g_x| / |g_v|,
 * the operand bounds of the gamma_s.1 / gamma_v.1 gradient GEMMs.               */
int goten_gata_bwd_src(const float* g_h, const float* g_Xd, const float* Xd, const float* qk,
                       int ldqk, const float* x, const float* v, const float* Ze, int ldz,
                       const float* Y, const float* fc, const float* kappa, const float* drop,
                       const float* alpha,
                       const float* da, const int32_t* src_ptr, const int32_t* src_perm,
                       const int32_t* tgt, int n_nodes, int C, int H, int lmax, int flags,
                       float* g_qk, int ldgqk, float* g_x, float* g_v, float* g_Xd_in, float* gx_amax,
                       float* gv_amax, void* stream);

/* ------------------------------------------------------------- HTR block --
 * GATA.edge_update + vector_rejection + residual (gotennet.py:351-364, :561-611,
 * :445): EQ/EK are [L][N][C] projections of the post-message X;
 *   w[e][c] = sum_l sum_m rej(EQ_i)^l_m rej(EK_j)^l_m ;  t_out = t + silu(zt) * w
 * zt = Ze[:, zt_col0 : zt_col0+C).  EQ / EK / g_EQ / g_EK rows have pitch ldp floats
 * (ldp = 2C with EK = EQ + C when both projections come from one GEMM with the stacked
 * weight [W_vq; W_vk]).  flags bits 2-3: gamma_w of the gated edge updates (0 identity,
 * 1 sigmoid "gated", 2 tanh "gatedt", 3 SiLU "act"; gotennet.py:283-289).
 * bit 4: gamma_t ends without activation ("mlp" edge update): zt is used as is.
 * bit 5: weight only (the "linw" family, gotennet.py:270-282): forward writes t_out[e][c] = w (C = evec_dim; Ze, t
 * unread, may be NULL); the backward halves take g_t_out = dL/dw and write g_EQ / g_EK (+ g_Y), no gZe.
 * flags: bit0 = sep_htr (rejection per degree),
 * bit1 = rejection enabled.                                                    */
int goten_htr_fwd(const float* EQ, const float* EK, int ldp, const float* Y, const float* Ze, int ldz,
                  int zt_col0, const float* t, const int32_t* tgt_ptr, const int32_t* src,
                  int n_nodes, int C, int lmax, int flags, float* t_out, float* t_amax,
                  void* stream);
/* target half: g_EQ[L][N][C], gZe[:, zt_col0..) = g_t_out * w * silu'(zt); optional g_Y (accumulated) */
int goten_htr_bwd_tgt(const float* g_t_out, const float* EQ, const float* EK, int ldp, const float* Y,
                      const float* Ze, int ldz, int zt_col0, const int32_t* tgt_ptr,
                      const int32_t* src, int n_nodes, int C, int lmax, int flags, float* g_EQ,
                      float* gZe, int ldgz, float* g_Y, float* gze_amax, float* geq_amax,
                      void* stream);
/* source half: g_EK[L][N][C] */
int goten_htr_bwd_src(const float* g_t_out, const float* EQ, const float* EK, int ldp, const float* Y,
                      const float* Ze, int ldz, int zt_col0, const int32_t* src_ptr,
                      const int32_t* src_perm, const int32_t* tgt, int n_nodes, int C, int lmax,
                      int flags, float* g_EK, float* gek_amax, void* stream);

/* ------------------------------------------------------------ EQFF block --
 * EQFF.forward (gotennet.py:728-748). P = X W_vu^T comes from goten_gemm.
 * ctx[N][2C] = [ h | sqrt(sum_m P^2 + eps) ]                                   */
int goten_eqff_ctx_fwd(const float* h, const float* P, int n_nodes, int C, int L, float eps,
                       float* ctx, float* ctx_amax, void* stream);
/* h_out = h + m[:, :C];  Xd_out[m][n][c] = Xd + m[n][C+c] * P[m][n][c]  (m = gamma_m output, [N][2C]) */
int goten_eqff_update_fwd(const float* h, const float* Xd, const float* P, const float* m,
                          int n_nodes, int C, int L, float* h_out, float* Xd_out, void* stream);
/* g_m[N][2C] = [ g_h_out | sum_m g_Xd_out * P ] */
int goten_eqff_update_bwd(const float* g_h_out, const float* g_Xd_out, const float* P, int n_nodes,
                          int C, int L, float* g_m, float* gm_amax, void* stream);
/* g_P = g_Xd_out * m2 + g_ctx[:, C:] * P / n ;  g_h = g_h_out + g_ctx[:, :C]  (n = ctx[:, C:]) */
int goten_eqff_ctx_bwd(const float* g_h_out, const float* g_Xd_out, const float* g_ctx,
                       const float* P, const float* m, const float* ctx, int n_nodes, int C, int L,
                       float* g_P, float* g_h, float* gp_amax, void* stream);

/* ------------------------------------------------------------ utilities --
 * out = a + b (gradient joins), layout change [N][L][C] <-> [L][N][C], row gather
 * (embedding lookup, gotennet.py:973 / layers.py:1665) and its deterministic
 * scatter-add backward.                                                        */
int goten_add(const float* a, const float* b, float* out, int64_t n, void* stream);
int goten_permute_nlc(const float* in, float* out, int n_nodes, int L, int C, int to_degree_major,
                      void* stream);
int goten_embedding_fwd(const float* table, const int64_t* idx, int64_t n, int C, float* out,
                        void* stream);
int goten_embedding_bwd(const float* g_out, const int64_t* idx, int64_t n, int C, int n_rows,
                        float* g_table, float* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------- read-out head --
 * Atomwise.forward (models/components/outputs.py:323-376): per-atom MLP output `raw`
 * [N][n_out] -> ScaleShift (components/layers.py:172-202, mean/stddev with 1 or n_out
 * entries, nullable) -> + atomref[z] (outputs.py:349-351, nullable) -> yi [N][n_out];
 * torch_scatter.scatter(yi, batch, reduce=mode) (outputs.py:354-355) -> y [n_mol][n_out]
 * as a deterministic segmented reduction over mol_ptr.  mode: 0 none, 1 sum, 2 mean.
 * goten_mol_ptr builds mol_ptr [n_mol+1] from the (sorted) int64 batch vector of a PyG
 * batch; unsorted_flag[0] (device, nullable) is set to 1 when the vector is not sorted.
 * goten_act_*: element-wise activations (layers.py:41-81, :619; 3 sigmoid, 4 tanh): kind 1 SiLU, 2 shifted
 * softplus.                                                                          */
int goten_mol_ptr(const int64_t* batch, int n_nodes, int n_mol, int32_t* mol_ptr, int32_t* unsorted_flag,
                  void* stream);
int goten_atomwise_reduce_fwd(const float* raw, const int64_t* z, const float* atomref, int atomref_rows,
                              const float* mean, const float* stddev, int n_stat, const int32_t* mol_ptr,
                              int n_nodes, int n_mol, int n_out, int mode, float* yi, float* y,
                              void* stream);
int goten_atomwise_reduce_bwd(const float* g_y, const float* g_yi, const float* stddev, int n_stat,
                              const int32_t* mol_ptr, int n_nodes, int n_mol, int n_out, int mode,
                              float* g_raw, void* stream);
int goten_act_fwd(int kind, const float* x, int64_t n, float* y, void* stream);
int goten_act_bwd(int kind, const float* g, const float* x, int64_t n, float* out, void* stream);
/* out = a * b + c;  g_a = g * b, g_b = g * a: the residual edge update t + gamma_t(t) * gamma_w(w) (gotennet.py:611,
 * :445) of the "linw" family, where gamma_w is a network on w (LayerNorm / activation / goten_gemm launches).       */
int goten_mul_add_fwd(const float* a, const float* b, const float* c, int64_t n, float* out, void* stream);
int goten_mul_add_bwd(const float* g, const float* a, const float* b, int64_t n, float* g_a, float* g_b, void* stream);

/* ------------------------------- equivariant read-out heads --
 * Element-wise / per-molecule stages of GatedEquivariantBlock (models/components/outputs.py:24-104: vmix [N][3][2*nv]
 * = mix_vectors(vectors) = [V | W]; ctx = [scalars | ||V||]; x = scalar_net(ctx) [N][nso+nv]; s_out = sact(x[:, :nso]),
 * v_out = x[:, nso:] * W), of Dipole (outputs.py:437-467: yi = mu_atom + pos * (stddev * q + mean), per-molecule sum
 * through goten_atomwise_reduce_*, optional magnitude = row norm) and of ElectronicSpatialExtentV2
 * (outputs.py:522-541: mass-weighted centroid per molecule, yi = |pos - c|^2 * x, y = per-molecule sum; `centroid`
 * [n_mol][4] keeps (c, total mass) for the backward).  sact: 0 none, 1 SiLU, 2 shifted softplus.  The dense layers of
 * the blocks are goten_gemm_scaled calls.  Norm gradients at a zero vector are 0 (torch semantics).               */
int goten_geb_ctx_fwd(const float* scalars, const float* vmix, int64_t n_nodes, int ns, int nv, float* ctx,
                      void* stream);
int goten_geb_ctx_bwd(const float* g_ctx, const float* vmix, int64_t n_nodes, int ns, int nv, float* g_scalars,
                      float* g_vmix, void* stream);
int goten_geb_gate_fwd(const float* x, const float* vmix, int64_t n_nodes, int nso, int nv, int sact, float* s_out,
                       float* v_out, void* stream);
int goten_geb_gate_bwd(const float* g_s, const float* g_v, const float* x, const float* vmix, int64_t n_nodes,
                       int nso, int nv, int sact, float* g_x, float* g_vmix, void* stream);
int goten_dipole_atom_fwd(const float* l1, const float* l0, const float* pos, float stddev, float mean,
                          int64_t n_nodes, float* yi, void* stream);
int goten_dipole_atom_bwd(const float* g_yi, const float* l0, const float* pos, float stddev, float mean,
                          int64_t n_nodes, float* g_l1, float* g_l0, float* g_pos, void* stream);
int goten_rownorm_fwd(const float* v, int64_t rows, int dim, float* y, void* stream);
int goten_rownorm_bwd(const float* g, const float* v, int64_t rows, int dim, float* g_v, void* stream);
int goten_ese_fwd(const float* x, const float* pos, const int64_t* z, const float* mass, int mass_rows,
                  const int32_t* mol_ptr, int n_mol, float* yi, float* y, float* centroid, void* stream);
int goten_ese_bwd(const float* g_y, const float* x, const float* pos, const int64_t* z, const float* mass,
                  int mass_rows, const int32_t* mol_ptr, int n_mol, const float* centroid, float* g_x,
                  float* g_pos, void* stream);

/* --------------------------------------------------------------- optimiser --
 * Training step of the reference on one FLAT fp32 parameter buffer
 * (models/goten_model.py:521-578: torch.optim.AdamW(eps=1e-7), weight decay on every
 * parameter; configs/trainer/default.yaml:10: gradient_clip_val 5.0, norm clipping).
 * goten_sumsq: out[0] = sum g[i]^2, deterministic (two stages, no atomics); `partial`
 *   holds goten_sumsq_workspace_floats() floats.
 * goten_adamw_step: g' = g * grad_scale * min(1, max_norm / (sqrt(sumsq[0]) * grad_scale + 1e-6))
 *   (max_norm <= 0 or sumsq NULL: no clipping), then the torch.optim.AdamW update of
 *   p, m (exp_avg), v (exp_avg_sq) in place.  bias_c1 = 1 - beta1^step, bias_c2 = 1 - beta2^step.
 *   The clip coefficient is formed on the device: no host read between backward and step. */
int goten_sumsq(const float* g, int64_t n, float* partial, float* out, void* stream);
int goten_sumsq_workspace_floats(void);
int goten_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1,
                     double beta2, double eps, double weight_decay, float bias_c1, float bias_c2,
                     float max_norm, const float* sumsq, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GOTENNET_B200_H */
